// Shared device-side definitions: elements, small dense algebra, constitutive laws.
// All arithmetic is FP64; nothing here touches tensor cores (the per-element contractions
// are 2x2 / 3x3 / 3x8).  Reference citations are relative to the tatva v0.11.1 tree.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tatva_b200.h"
#else
// Run-time compilation (NVRTC, csrc/user_law.cu): no host headers; the few names the device code needs.
typedef int int32_t;
typedef long long int64_t;
typedef unsigned short uint16_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
enum { TATVA_TRI3 = 0, TATVA_TET4 = 1, TATVA_HEX8 = 2, TATVA_QUAD4 = 3, TATVA_TRI6 = 4, TATVA_QUAD8 = 5, TATVA_LINE2 = 6, TATVA_LINE3 = 7 };
#endif

#define TATVA_HD __host__ __device__ __forceinline__
#define TATVA_D __device__ __forceinline__

namespace tatva {

constexpr int kBlock = 128;  // threads per CTA for element-per-thread kernels

// ---------------------------------------------------------------------------------------------
// Elements.  dNdr(q) is dN/dxi at quadrature point q, laid out [dim][npe] exactly like
// Element.shape_function_derivative (tatva/element/base.py:86-88).
// ---------------------------------------------------------------------------------------------

struct Tri3 {  // tatva/element/base.py:245-265
  static constexpr int dim = 2, gdim = 2, npe = 3, nq = 1, kind = TATVA_TRI3;
  static constexpr int max_nq = nq, rdim = 2;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static void N_at(const double* xi, double (&n)[npe]) { n[0] = 1.0 - xi[0] - xi[1]; n[1] = xi[0]; n[2] = xi[1]; }
  TATVA_HD static void dNdr_at(const double*, double (&d)[dim][npe]) { dNdr(0, d); }
  TATVA_HD static double weight(int) { return 0.5; }
  TATVA_HD static void N(int, double (&n)[npe]) {
    n[0] = 1.0 - 1.0 / 3 - 1.0 / 3;
    n[1] = 1.0 / 3;
    n[2] = 1.0 / 3;
  }
  TATVA_HD static void dNdr(int, double (&d)[dim][npe]) {
    d[0][0] = -1.0; d[0][1] = 1.0; d[0][2] = 0.0;
    d[1][0] = -1.0; d[1][1] = 0.0; d[1][2] = 1.0;
  }
};

struct Tet4 {  // tatva/element/base.py:448-472
  static constexpr int dim = 3, gdim = 3, npe = 4, nq = 1, kind = TATVA_TET4;
  static constexpr int max_nq = nq, rdim = 3;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static void N_at(const double* xi, double (&n)[npe]) { n[0] = 1.0 - xi[0] - xi[1] - xi[2]; n[1] = xi[0]; n[2] = xi[1]; n[3] = xi[2]; }
  TATVA_HD static void dNdr_at(const double*, double (&d)[dim][npe]) { dNdr(0, d); }
  TATVA_HD static double weight(int) { return 1.0 / 6; }
  TATVA_HD static void N(int, double (&n)[npe]) {
    n[0] = 1.0 - 0.25 - 0.25 - 0.25;
    n[1] = 0.25;
    n[2] = 0.25;
    n[3] = 0.25;
  }
  TATVA_HD static void dNdr(int, double (&d)[dim][npe]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      d[j][0] = -1.0;
#pragma unroll
      for (int n = 1; n < 4; ++n) d[j][n] = (n == j + 1) ? 1.0 : 0.0;
    }
  }
};

struct Hex8 {  // tatva/element/base.py:475-568
  static constexpr int dim = 3, gdim = 3, npe = 8, nq = 8, kind = TATVA_HEX8;
  static constexpr int max_nq = nq, rdim = 3;
  TATVA_HD static constexpr int num_q() { return nq; }
  // sign of reference node n along axis d (bottom face CCW, then top; :478-491); the 2x2x2
  // Gauss points are a * the same table (:493-513), all weights 1.
  TATVA_HD static constexpr double sgn(int n, int d) {
    return d == 0 ? (((n & 3) == 1 || (n & 3) == 2) ? 1.0 : -1.0)
         : d == 1 ? (((n & 3) >= 2) ? 1.0 : -1.0)
                  : ((n >= 4) ? 1.0 : -1.0);
  }
  TATVA_HD static double weight(int) { return 1.0; }
  TATVA_HD static void N(int q, double (&n)[npe]) {
    const double a = 0.57735026918962576451;  // 1/sqrt(3)
    const double x = a * sgn(q, 0), y = a * sgn(q, 1), z = a * sgn(q, 2);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      n[k] = 0.125 * (1.0 + sgn(k, 0) * x) * (1.0 + sgn(k, 1) * y) * (1.0 + sgn(k, 2) * z);
  }
  TATVA_HD static void dNdr(int q, double (&d)[dim][npe]) {
    const double a = 0.57735026918962576451;
    const double x = a * sgn(q, 0), y = a * sgn(q, 1), z = a * sgn(q, 2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double fx = 1.0 + sgn(k, 0) * x, fy = 1.0 + sgn(k, 1) * y, fz = 1.0 + sgn(k, 2) * z;
      d[0][k] = 0.125 * sgn(k, 0) * fy * fz;
      d[1][k] = 0.125 * sgn(k, 1) * fx * fz;
      d[2][k] = 0.125 * sgn(k, 2) * fx * fy;
    }
  }
  TATVA_HD static void N_at(const double* xi, double (&n)[npe]) {  // :515-529 at an arbitrary point
#pragma unroll
    for (int k = 0; k < 8; ++k) n[k] = 0.125 * (1.0 + sgn(k, 0) * xi[0]) * (1.0 + sgn(k, 1) * xi[1]) * (1.0 + sgn(k, 2) * xi[2]);
  }
  TATVA_HD static void dNdr_at(const double* xi, double (&d)[dim][npe]) {  // :531-568
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double fx = 1.0 + sgn(k, 0) * xi[0], fy = 1.0 + sgn(k, 1) * xi[1], fz = 1.0 + sgn(k, 2) * xi[2];
      d[0][k] = 0.125 * sgn(k, 0) * fy * fz;
      d[1][k] = 0.125 * sgn(k, 1) * fx * fz;
      d[2][k] = 0.125 * sgn(k, 2) * fx * fy;
    }
  }
};

struct Quad4 {  // tatva/element/base.py:331-366; 2x2 Gauss points, x fastest (:338-344)
  static constexpr int dim = 2, gdim = 2, npe = 4, nq = 4, kind = TATVA_QUAD4;
  static constexpr int max_nq = nq, rdim = 2;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static constexpr double sgn(int n, int d) { return d == 0 ? ((n == 1 || n == 2) ? 1.0 : -1.0) : ((n >= 2) ? 1.0 : -1.0); }
  TATVA_HD static double weight(int) { return 1.0; }
  TATVA_HD static void xi(int q, double& r, double& s) {
    const double a = 0.57735026918962576451;
    r = (q & 1) ? a : -a;
    s = (q & 2) ? a : -a;
  }
  TATVA_HD static void N_at(const double* x, double (&n)[npe]) {
    const double r = x[0], s = x[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) n[k] = 0.25 * (1.0 + sgn(k, 0) * r) * (1.0 + sgn(k, 1) * s);
  }
  TATVA_HD static void dNdr_at(const double* x, double (&d)[dim][npe]) {
    const double r = x[0], s = x[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      d[0][k] = 0.25 * sgn(k, 0) * (1.0 + sgn(k, 1) * s);
      d[1][k] = 0.25 * sgn(k, 1) * (1.0 + sgn(k, 0) * r);
    }
  }
  TATVA_HD static void N(int q, double (&n)[npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    N_at(x, n);
  }
  TATVA_HD static void dNdr(int q, double (&d)[dim][npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    dNdr_at(x, d);
  }
};

struct Tri6 {  // tatva/element/base.py:266-328; 3-point rule (:278-284)
  static constexpr int dim = 2, gdim = 2, npe = 6, nq = 3, kind = TATVA_TRI6;
  static constexpr int max_nq = nq, rdim = 2;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static double weight(int) { return 1.0 / 6.0; }
  TATVA_HD static void xi(int q, double& r, double& s) {
    r = (q == 1) ? 2.0 / 3.0 : 1.0 / 6.0;
    s = (q == 2) ? 2.0 / 3.0 : 1.0 / 6.0;
  }
  TATVA_HD static void N(int q, double (&n)[npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    N_at(x, n);
  }
  TATVA_HD static void dNdr(int q, double (&d)[dim][npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    dNdr_at(x, d);
  }
  TATVA_HD static void N_at(const double* x, double (&n)[npe]) {
    const double r = x[0], s = x[1];
    const double t = 1.0 - r - s;
    n[0] = t * (2 * t - 1); n[1] = r * (2 * r - 1); n[2] = s * (2 * s - 1);
    n[3] = 4 * r * t; n[4] = 4 * r * s; n[5] = 4 * s * t;
  }
  TATVA_HD static void dNdr_at(const double* x, double (&d)[dim][npe]) {
    const double r = x[0], s = x[1];
    const double t = 1.0 - r - s;
    d[0][0] = -(4 * t - 1); d[0][1] = 4 * r - 1; d[0][2] = 0.0; d[0][3] = 4 * (t - r); d[0][4] = 4 * s; d[0][5] = -4 * s;
    d[1][0] = -(4 * t - 1); d[1][1] = 0.0; d[1][2] = 4 * s - 1; d[1][3] = -4 * r; d[1][4] = 4 * r; d[1][5] = 4 * (t - s);
  }
};

struct Quad8 {  // tatva/element/base.py:366-445; 3x3 Gauss points, x fastest (:384-393)
  static constexpr int dim = 2, gdim = 2, npe = 8, nq = 9, kind = TATVA_QUAD8;
  static constexpr int max_nq = nq, rdim = 2;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static double w1(int i) { return i == 1 ? 8.0 / 9.0 : 5.0 / 9.0; }
  TATVA_HD static double x1(int i) { return i == 0 ? -0.77459666924148337704 : (i == 1 ? 0.0 : 0.77459666924148337704); }
  TATVA_HD static double weight(int q) { return w1(q / 3) * w1(q % 3); }
  TATVA_HD static void xi(int q, double& r, double& s) {
    r = x1(q % 3);
    s = x1(q / 3);
  }
  TATVA_HD static void N(int q, double (&n)[npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    N_at(x, n);
  }
  TATVA_HD static void dNdr(int q, double (&d)[dim][npe]) {
    double x[2];
    xi(q, x[0], x[1]);
    dNdr_at(x, d);
  }
  TATVA_HD static void N_at(const double* x, double (&n)[npe]) {
    const double r = x[0], s = x[1];
    n[0] = 0.25 * (1 - r) * (1 - s) * (-r - s - 1);
    n[1] = 0.25 * (1 + r) * (1 - s) * (r - s - 1);
    n[2] = 0.25 * (1 + r) * (1 + s) * (r + s - 1);
    n[3] = 0.25 * (1 - r) * (1 + s) * (-r + s - 1);
    n[4] = 0.5 * (1 - r * r) * (1 - s);
    n[5] = 0.5 * (1 + r) * (1 - s * s);
    n[6] = 0.5 * (1 - r * r) * (1 + s);
    n[7] = 0.5 * (1 - r) * (1 - s * s);
  }
  TATVA_HD static void dNdr_at(const double* x, double (&d)[dim][npe]) {
    const double r = x[0], s = x[1];
    d[0][0] = 0.25 * (-2 * r - s) * (s - 1); d[0][1] = 0.25 * (-2 * r + s) * (s - 1);
    d[0][2] = 0.25 * (2 * r + s) * (s + 1);  d[0][3] = 0.25 * (2 * r - s) * (s + 1);
    d[0][4] = r * (s - 1); d[0][5] = 0.5 - 0.5 * s * s; d[0][6] = -r * (s + 1); d[0][7] = 0.5 * s * s - 0.5;
    d[1][0] = 0.25 * (-r - 2 * s) * (r - 1); d[1][1] = 0.25 * (-r + 2 * s) * (r + 1);
    d[1][2] = 0.25 * (r + 1) * (r + 2 * s);  d[1][3] = 0.25 * (r - 1) * (r - 2 * s);
    d[1][4] = 0.5 * r * r - 0.5; d[1][5] = -s * (r + 1); d[1][6] = 0.5 - 0.5 * r * r; d[1][7] = s * (r - 1);
  }
};

// Line elements embedded in the plane (boundary integrals): `dim` is the width of a coordinate row, `gdim` the
// number of gradient components (1: the derivative along the arc length).  tatva/element/base.py:144-242.
struct Line2 {  // tatva/element/base.py:144-186
  static constexpr int dim = 2, gdim = 1, npe = 2, nq = 1, kind = TATVA_LINE2;
  static constexpr int max_nq = nq, rdim = 1;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static void N_at(const double* x, double (&n)[npe]) { n[0] = 0.5 * (1.0 - x[0]); n[1] = 0.5 * (1.0 + x[0]); }
  TATVA_HD static void dNdr_at(const double*, double (&d)[gdim][npe]) { d[0][0] = -0.5; d[0][1] = 0.5; }
  TATVA_HD static double weight(int) { return 2.0; }
  TATVA_HD static void N(int, double (&n)[npe]) { n[0] = 0.5; n[1] = 0.5; }
  TATVA_HD static void dNdr(int, double (&d)[gdim][npe]) { d[0][0] = -0.5; d[0][1] = 0.5; }
};

struct Line3 {  // tatva/element/base.py:189-242; nodes (-1, 1, 0), 3-point Gauss
  static constexpr int dim = 2, gdim = 1, npe = 3, nq = 3, kind = TATVA_LINE3;
  static constexpr int max_nq = nq, rdim = 1;
  TATVA_HD static constexpr int num_q() { return nq; }
  TATVA_HD static void N_at(const double* x, double (&n)[npe]) {
    const double r = x[0];
    n[0] = 0.5 * r * (r - 1.0);
    n[1] = 0.5 * r * (r + 1.0);
    n[2] = 1.0 - r * r;
  }
  TATVA_HD static void dNdr_at(const double* x, double (&d)[gdim][npe]) {
    const double r = x[0];
    d[0][0] = r - 0.5;
    d[0][1] = r + 0.5;
    d[0][2] = -2.0 * r;
  }
  TATVA_HD static double xi(int q) { return (q - 1) * 0.77459666924148337704; }  // sqrt(3/5)
  TATVA_HD static double weight(int q) { return q == 1 ? 8.0 / 9 : 5.0 / 9; }
  TATVA_HD static void N(int q, double (&n)[npe]) {
    const double r = xi(q);
    n[0] = 0.5 * r * (r - 1.0);
    n[1] = 0.5 * r * (r + 1.0);
    n[2] = 1.0 - r * r;
  }
  TATVA_HD static void dNdr(int q, double (&d)[gdim][npe]) {
    const double r = xi(q);
    d[0][0] = r - 0.5;
    d[0][1] = r + 0.5;
    d[0][2] = -2.0 * r;
  }
};

// ---------------------------------------------------------------------------------------------
// User-supplied quadrature rule (Element(quad_points, quad_weights), tatva/element/base.py:37-51): the points and
// weights sit in constant memory (installed stream-ordered before each launch from the plan's host copy) and the shape
// functions are evaluated at them.  The generic kernels take it as `Custom<El>`; the specialised kernels (modal Hex8,
// reference-space Tet4) carry their element's default rule and are not used with a custom one.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxQ = 64;
struct QuadRule {
  int nq;
  double xi[kMaxQ][3];
  double w[kMaxQ];
};
#ifdef __CUDACC__
static __constant__ QuadRule c_rule;  // one copy per translation unit; generic.cu installs and uses its own
#endif

template <class B>
struct Custom {
  static constexpr int dim = B::dim, gdim = B::gdim, npe = B::npe, kind = B::kind, max_nq = kMaxQ, rdim = B::rdim;
  static constexpr int nq = -1;  // not a compile-time constant: use num_q()
#ifdef __CUDA_ARCH__
  TATVA_HD static int num_q() { return c_rule.nq; }
  TATVA_HD static double weight(int q) { return c_rule.w[q]; }
  TATVA_HD static void N(int q, double (&n)[npe]) { B::N_at(c_rule.xi[q], n); }
  TATVA_HD static void dNdr(int q, double (&d)[gdim][npe]) { B::dNdr_at(c_rule.xi[q], d); }
#else  // the kernels run on the device only; the host twins exist so that __host__ __device__ templates compile
  TATVA_HD static int num_q() { return 0; }
  TATVA_HD static double weight(int) { return 0.0; }
  TATVA_HD static void N(int, double (&)[npe]) {}
  TATVA_HD static void dNdr(int, double (&)[gdim][npe]) {}
#endif
};

// ---------------------------------------------------------------------------------------------
// Nodal row gather.  A node's W doubles are contiguous; fetching them with 16-byte loads where alignment allows
// cuts the load instructions (and the L1 tag lookups of a scattered gather) from W to ceil(W/2).  Rows of 3 doubles
// alternate between 16-byte aligned and 8-off, so each row is one aligned double2 plus one double, picked by the
// parity of (node + base misalignment); no divergence, two selects per row.
// MEASURED SLOWER on B200 and therefore off (r01: Hex8 HVP 0.490 vs 0.453 ms, Tri3 HVP 20.7 vs 16.7 us, Tet4 energy
// 36.6 vs 32.8 us, Tet4 HVP unchanged): the gathers are latency- not instruction-bound and the extra selects and
// address arithmetic cost more issue slots than the saved loads.  Kept behind the macro as a record of the experiment.
// ---------------------------------------------------------------------------------------------
#ifndef TATVA_WIDE_GATHER
#define TATVA_WIDE_GATHER 0
#endif

template <int W>
TATVA_D void load_row(const double* __restrict__ src, int64_t node, double (&dst)[W]) {
#if TATVA_WIDE_GATHER
  const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(src) >> 3) & 1u);  // base is 8 mod 16 (uniform)
  if constexpr (W == 3) {
    const int64_t o = node * 3;
    const unsigned odd = ((unsigned)node + mis) & 1u;  // row start is 8 mod 16
    const double2 v = __ldg(reinterpret_cast<const double2*>(src + o + odd));
    const double s = __ldg(src + o + (odd ? 0 : 2));
    dst[0] = odd ? s : v.x;
    dst[1] = odd ? v.x : v.y;
    dst[2] = odd ? v.y : s;
    return;
  } else if constexpr (W == 2 || W == 4) {
    if (!mis) {
      const double2* q = reinterpret_cast<const double2*>(src + node * W);
#pragma unroll
      for (int k = 0; k < W / 2; ++k) {
        const double2 v = __ldg(q + k);
        dst[2 * k] = v.x;
        dst[2 * k + 1] = v.y;
      }
      return;
    }
  }
#endif
#pragma unroll
  for (int c = 0; c < W; ++c) dst[c] = __ldg(src + node * W + c);
}

// ---------------------------------------------------------------------------------------------
// d x d determinant / inverse (closed form; reference uses jnp.linalg.det / inv on J,
// tatva/element/base.py:92, :113)
// ---------------------------------------------------------------------------------------------

TATVA_HD double det_inv(const double (&A)[2][2], double (&Ai)[2][2]) {
  const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const double r = 1.0 / det;
  Ai[0][0] = A[1][1] * r;
  Ai[0][1] = -A[0][1] * r;
  Ai[1][0] = -A[1][0] * r;
  Ai[1][1] = A[0][0] * r;
  return det;
}

TATVA_HD double det_inv(const double (&A)[3][3], double (&Ai)[3][3]) {
  const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  const double r = 1.0 / det;
  Ai[0][0] = c00 * r;
  Ai[1][0] = c01 * r;
  Ai[2][0] = c02 * r;
  Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * r;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * r;
  Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * r;
  Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * r;
  Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * r;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * r;
  return det;
}

// Per-quadrature-point geometry: dNdX = inv(J) @ dNdr with J = dNdr @ X_e
// (tatva/element/base.py:111-113); returns det J (:92).
// Arc-length Jacobian of a line element: J = |dNdr @ X_e| (the reference's dot(Jvec, Jvec / |Jvec|),
// tatva/element/base.py:163-167, :218-224), dNdS = dNdr / J.
template <class El>
TATVA_HD double line_jacobian(int q, const double (&X)[El::npe][El::dim], double (&dNdr)[1][El::npe]) {
  El::dNdr(q, dNdr);
  double n2 = 0.0;
#pragma unroll
  for (int c = 0; c < El::dim; ++c) {
    double s = 0.0;
#pragma unroll
    for (int n = 0; n < El::npe; ++n) s += dNdr[0][n] * X[n][c];
    n2 += s * s;
  }
  return sqrt(n2);
}

template <class El>
TATVA_HD double geometry(int q, const double (&X)[El::npe][El::dim], double (&dNdX)[El::gdim][El::npe]) {
  if constexpr (El::gdim != El::dim) {
    const double J = line_jacobian<El>(q, X, dNdX);
#pragma unroll
    for (int n = 0; n < El::npe; ++n) dNdX[0][n] /= J;
    return J;
  } else {
    double dNdr[El::dim][El::npe];
    El::dNdr(q, dNdr);
    double J[El::dim][El::dim], Ji[El::dim][El::dim];
#pragma unroll
    for (int d = 0; d < El::dim; ++d)
#pragma unroll
      for (int c = 0; c < El::dim; ++c) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < El::npe; ++n) s += dNdr[d][n] * X[n][c];
        J[d][c] = s;
      }
    const double det = det_inv(J, Ji);
#pragma unroll
    for (int c = 0; c < El::dim; ++c)
#pragma unroll
      for (int n = 0; n < El::npe; ++n) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < El::dim; ++d) s += Ji[c][d] * dNdr[d][n];
        dNdX[c][n] = s;
      }
    return det;
  }
}

template <class El>
TATVA_HD double det_jacobian(int q, const double (&X)[El::npe][El::dim]) {
  if constexpr (El::gdim != El::dim) {
    double dNdr[1][El::npe];
    return line_jacobian<El>(q, X, dNdr);
  } else {
    double dNdr[El::dim][El::npe];
    El::dNdr(q, dNdr);
    double J[El::dim][El::dim];
#pragma unroll
    for (int d = 0; d < El::dim; ++d)
#pragma unroll
      for (int c = 0; c < El::dim; ++c) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < El::npe; ++n) s += dNdr[d][n] * X[n][c];
        J[d][c] = s;
      }
    if constexpr (El::dim == 2) {
      return J[0][0] * J[1][1] - J[0][1] * J[1][0];
    } else {
      return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
             J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Constitutive laws.  A law sees, per quadrature point, the gradients G[c][j] = d s_c / d x_j of
// every nodal component c (c < dpn) and — for components >= val_lo — the interpolated value.
// first()  fills the conjugate fluxes  A[c][j] = d psi / d G[c][j],  b[c] = d psi / d val[c];
// second() fills their directional derivative along (dG, dval).
// ---------------------------------------------------------------------------------------------

template <int DPN, int DIM>
struct QState {
  double G[DPN][DIM];
  double val[DPN];
};

template <int DIM>
struct LinearElastic {  // tests/test_sparse.py:20-38
  static constexpr int dim = DIM, dpn = DIM, val_lo = DIM, n_params = 2;
  static constexpr bool needs_u_for_hvp = false;  // quadratic energy: H does not depend on u
  double mu, lmbda;
  struct Cache {};
  using S = QState<dpn, dim>;
  TATVA_HD void prepare(const S&, Cache&) const {}
  TATVA_HD double psi(const S& s, const Cache&) const {
    double tr = 0.0, ee = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      tr += s.G[i][i];
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        const double e = 0.5 * (s.G[i][j] + s.G[j][i]);
        ee += e * e;
      }
    }
    return mu * ee + 0.5 * lmbda * tr * tr;
  }
  TATVA_HD void first(const S& s, const Cache&, S& f) const {
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) tr += s.G[i][i];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) f.G[i][j] = mu * (s.G[i][j] + s.G[j][i]) + (i == j ? lmbda * tr : 0.0);
  }
  TATVA_HD void second(const S&, const Cache& c, const S& ds, S& f) const { first(ds, c, f); }
};

struct NeoHookean {  // tests/test_sparse_tracer.py:103-115 (mu=500, lambda=1000 at :126)
  static constexpr int dim = 3, dpn = 3, val_lo = 3, n_params = 2;
  static constexpr bool needs_u_for_hvp = true;
  double mu, lmbda;
  struct Cache {
    double Fi[3][3];  // F^-1
    double lnJ;
    double I1;
  };
  using S = QState<3, 3>;
  TATVA_HD void prepare(const S& s, Cache& c) const {
    double F[3][3];
    double I1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        F[i][j] = s.G[i][j] + (i == j ? 1.0 : 0.0);
        I1 += F[i][j] * F[i][j];
      }
    const double J = det_inv(F, c.Fi);
    c.lnJ = log(J);
    c.I1 = I1;
  }
  TATVA_HD double psi(const S&, const Cache& c) const {
    return 0.5 * mu * (c.I1 - 3.0 - 2.0 * c.lnJ) + 0.5 * lmbda * c.lnJ * c.lnJ;
  }
  // P = mu (F - F^-T) + lambda lnJ F^-T
  TATVA_HD void first(const S& s, const Cache& c, S& f) const {
    const double k = lmbda * c.lnJ - mu;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) f.G[i][j] = mu * (s.G[i][j] + (i == j ? 1.0 : 0.0)) + k * c.Fi[j][i];
  }
  // dP = mu dG + (mu - lambda lnJ) F^-T dG^T F^-T + lambda tr(F^-1 dG) F^-T
  TATVA_HD void second(const S&, const Cache& c, const S& ds, S& f) const {
    double B[3][3];  // F^-1 dG
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) t += c.Fi[i][k] * ds.G[k][j];
        B[i][j] = t;
        if (i == j) tr += t;
      }
    const double k1 = mu - lmbda * c.lnJ, k2 = lmbda * tr;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;  // (B F^-1)[j][i]
#pragma unroll
        for (int k = 0; k < 3; ++k) t += B[j][k] * c.Fi[k][i];
        f.G[i][j] = mu * ds.G[i][j] + k1 * t + k2 * c.Fi[j][i];
      }
  }
};

// Config-5 two-field law (builder-defined AT2; no counterpart in the reference):
// psi = ((1-phi)^2 + k) psi_NH(grad u) + Gc (phi^2/(2 l) + l/2 |grad phi|^2); nodal state [ux,uy,uz,phi].
struct NeoHookeanPhaseField {
  static constexpr int dim = 3, dpn = 4, val_lo = 3, n_params = 5;
  static constexpr bool needs_u_for_hvp = true;
  double mu, lmbda, Gc, ell, k;
  struct Cache {
    NeoHookean::Cache nh;
    double psi_nh;
    double P[3][3];
  };
  using S = QState<4, 3>;
  TATVA_HD NeoHookean nh() const { return NeoHookean{mu, lmbda}; }
  TATVA_HD static void sub(const S& s, NeoHookean::S& t) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) t.G[i][j] = s.G[i][j];
  }
  TATVA_HD void prepare(const S& s, Cache& c) const {
    NeoHookean::S t, P;
    sub(s, t);
    const NeoHookean m = nh();
    m.prepare(t, c.nh);
    c.psi_nh = m.psi(t, c.nh);
    m.first(t, c.nh, P);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) c.P[i][j] = P.G[i][j];
  }
  TATVA_HD double psi(const S& s, const Cache& c) const {
    const double phi = s.val[3];
    const double g = (1.0 - phi) * (1.0 - phi) + k;
    const double gg = s.G[3][0] * s.G[3][0] + s.G[3][1] * s.G[3][1] + s.G[3][2] * s.G[3][2];
    return g * c.psi_nh + Gc * (phi * phi / (2.0 * ell) + 0.5 * ell * gg);
  }
  TATVA_HD void first(const S& s, const Cache& c, S& f) const {
    const double phi = s.val[3];
    const double g = (1.0 - phi) * (1.0 - phi) + k, dg = -2.0 * (1.0 - phi);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) f.G[i][j] = g * c.P[i][j];
#pragma unroll
    for (int j = 0; j < 3; ++j) f.G[3][j] = Gc * ell * s.G[3][j];
    f.val[3] = dg * c.psi_nh + Gc * phi / ell;
  }
  TATVA_HD void second(const S& s, const Cache& c, const S& ds, S& f) const {
    const double phi = s.val[3], dphi = ds.val[3];
    const double g = (1.0 - phi) * (1.0 - phi) + k, dg = -2.0 * (1.0 - phi);
    NeoHookean::S t, dt, dP;
    sub(s, t);
    sub(ds, dt);
    nh().second(t, c.nh, dt, dP);
    double PdG = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        f.G[i][j] = g * dP.G[i][j] + dg * dphi * c.P[i][j];
        PdG += c.P[i][j] * ds.G[i][j];
      }
#pragma unroll
    for (int j = 0; j < 3; ++j) f.G[3][j] = Gc * ell * ds.G[3][j];
    f.val[3] = dg * PdG + 2.0 * dphi * c.psi_nh + Gc * dphi / ell;
  }
};

// Sector-grouped scatter-add of the per-element nodal contributions of one warp: the values are re-dealt
// through shared memory so that consecutive lanes add the DPN consecutive doubles of one node (one 32-byte
// sector per group instead of DPN separate ones; the L2 atomic units work per sector).  All 32 lanes must call.
// doubles of shared memory per warp.  Per-lane row strides are ODD (values: S | 1 doubles, node ids: NPE | 1 ints) so
// that the lane-strided staging stores are free of bank conflicts (an even stride such as 12 or 24 doubles puts
// every 4th / 2nd lane on the same bank).
template <int NPE, int DPN>
TATVA_HD constexpr int grouped_scatter_words() {
  return 32 * ((NPE * DPN) | 1) + 16 * (NPE | 1);
}
template <int NPE, int DPN>
TATVA_D void grouped_scatter(double* __restrict__ y, const int (&nd)[NPE], const double (&Y)[NPE][DPN], bool valid,
                             double* warp_smem) {
  constexpr int S = NPE * DPN, SP = S | 1, NP = NPE | 1;
  const int lane = threadIdx.x & 31;
  int* snode = reinterpret_cast<int*>(warp_smem + 32 * SP);
#pragma unroll
  for (int n = 0; n < NPE; ++n) {
    snode[lane * NP + n] = valid ? nd[n] : -1;
#pragma unroll
    for (int c = 0; c < DPN; ++c) warp_smem[lane * SP + n * DPN + c] = Y[n][c];
  }
  __syncwarp();
  // the launcher may have put this grid behind the kernel that clears y (launch_behind_zero): everything up to here ran
  // while y was being cleared; the adds must wait for it.  A no-op for a normally launched grid.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int t = lane; t < 32 * S; t += 32) {
    const int j = t / S, r = t - j * S;
    const int node = snode[j * NP + r / DPN];
    if (node >= 0) atomicAdd(y + (int64_t)node * DPN + (r % DPN), warp_smem[j * SP + r]);
  }
  __syncwarp();
}
template <int NPE, int DPN>
TATVA_HD constexpr size_t grouped_scatter_smem(int warps) {
  return (size_t)warps * grouped_scatter_words<NPE, DPN>() * sizeof(double);
}

}  // namespace tatva

#ifndef __CUDACC_RTC__  // everything below is host code
// ---------------------------------------------------------------------------------------------
// The plan (opaque to C callers)
// ---------------------------------------------------------------------------------------------

struct tatva_plan {
  int element;
  int dim, gdim, npe, nq;  // dim: width of a coordinate row; gdim: gradient components (1 for line elements)
  int64_t n_nodes, n_elems;
  const double* coords;  // caller-owned device view (n_nodes, dim)
  const int32_t* conn;   // caller-owned device view (n_elems, npe)
  int flags;
  int variant;
  int zero_output;  // scatter-add entry points zero their output first (1) or accumulate (0)
  int pss;          // 1: the output is being cleared by the tatva_zero_release launched just before on the same stream — kernels that wait before their first add may be launched with programmatic stream serialization (sub-range calls, zero_y = 2)
  double* scratch;  // plan-owned: energy partials / row-sum partials
  int64_t scratch_len;
  double* weights;  // plan-owned (n_elems, nq) when TATVA_PLAN_CACHE_WEIGHTS
  // optional shared-memory staging tiles (caller-owned device views, see tatva_plan_set_tiles)
  const int32_t* tile_ptr;
  const int32_t* tile_nodes;
  const uint16_t* tile_conn;
  int tile_max_unique;
  // optional geometry cache of the Hex8 pair kernels (tatva_plan_cache_geometry): plan-owned, 64 doubles per element
  double* geo;
  int64_t geo_stride;  // elements per row of the cache (the full element list, also for a sub-range view)
  // optional node schedule of the warp-cooperative fused kernels (tatva_plan_set_node_schedule)
  const int32_t* ws_warp_nodes;
  const uint8_t* ws_warp_local;
  const int32_t* ws_tile_hdr;
  const int32_t* ws_tn_node;
  const int32_t* ws_ell_ptr;
  const uint16_t* ws_ell;
  // optional uniform background grid for point location (see tatva_plan_set_point_grid)
  int grid_nx, grid_ny;
  double grid_lo[2], grid_inv[2];  // bin = clamp(floor((x - lo) * inv), 0, n - 1)
  const int32_t* grid_ptr;         // caller-owned device views
  const int32_t* grid_elems;
  // optional user quadrature rule (tatva_plan_set_quadrature): host copy, installed in constant memory before a launch
  int custom;
  tatva::QuadRule rule;
};

namespace tatva {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: remember the opt-in per (call site, device), so a
// process driving several GPUs configures the kernel on each of them.
struct SmemOptIn {
  bool done[64] = {};
};
template <class K>
inline int opt_in_smem(K kernel, size_t bytes, SmemOptIn& flags) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && flags.done[dev]) return TATVA_OK;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  if (tracked) flags.done[dev] = true;
  return TATVA_OK;
}

inline int grid_for(int64_t n, int block = kBlock) { return (int)((n + block - 1) / block); }

#define TATVA_CUDA_TRY(expr)                \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

#define TATVA_LAUNCH_CHECK()                \
  do {                                      \
    cudaError_t _e = cudaPeekAtLastError(); \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

// y = 0 as a KERNEL that releases its dependent grid at once (griddepcontrol.launch_dependents): the element kernel
// launched behind it with programmatic stream serialization runs its gather and arithmetic while y is still being cleared
// and waits (griddepcontrol.wait) only before its first atomic add.  XLA does not zero results, so every scatter-add entry
// point clears its output; as a memset in front of the kernel that was 1-3 % of a Hex8 step and up to 15 % of a
// launch-sized one (Tri3 config 1).
static __global__ void __launch_bounds__(256) k_zero_release(double* __restrict__ y, int64_t n) {
#ifdef __CUDA_ARCH__
  asm volatile("griddepcontrol.launch_dependents;" ::);
#endif
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n2 = n >> 1;
  double2* y2 = reinterpret_cast<double2*>(y);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) y2[i] = make_double2(0.0, 0.0);
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = 0.0;
}

// Clear y[0..n) (when `zero`) and launch `kernel` behind it; the kernel must execute griddepcontrol.wait before it touches y.
template <class... KArgs, class... Args>
inline int launch_behind_zero(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int zero, double* y,
                              int64_t n, Args... args) {  // zero: 0 plain launch, 1 clear y and launch behind it, 2 launch behind a tatva_zero_release the caller issued
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (zero == 2) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  } else if (zero && n > 0) {
    if (reinterpret_cast<uintptr_t>(y) & 15) {  // not 16-byte aligned (a view into a larger array): plain memset, plain launch
      TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * n, st));
    } else {
      int64_t blocks = (n / 2 + 256 * 8 - 1) / (256 * 8);  // <= 8 double2 stores per thread, at most two CTAs per SM
      if (blocks > 296) blocks = 296;
      if (blocks < 1) blocks = 1;
      k_zero_release<<<(int)blocks, 256, 0, st>>>(y, n);
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 1;
    }
  }
  TATVA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
  return TATVA_OK;
}


// user-supplied laws compiled at run time (user_law.cu); material ids >= TATVA_USER_LAW_BASE
bool is_user_law(int material);
int user_law_info(int material, int* dpn);
int user_law_launch(const tatva_plan* p, int material, int what, const double* prm, int n_params, const double* u,
                    const double* v, const int32_t* map, double* out, cudaStream_t st);
int sum_partials(const double* partials, int n, double* out, cudaStream_t st);  // fixed-order sum (generic.cu)

// entry points implemented in the per-topic translation units
int hex8_nh_hvp_modal(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                      cudaStream_t st);
int hex8_nh_hvp_modal_lifted(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v_red,
                             const int32_t* map, double* y_red, cudaStream_t st, double* dot_partials = nullptr);

int hex8_nh_hvp_modal_dot(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                          double* dot_partials, cudaStream_t st);
int hex8_geometry_cache(const tatva_plan* p, double* geo, int64_t stride, cudaStream_t st);
int hex8_grad_modal(const tatva_plan* p, bool adjoint, const double* in, int nv, double* out, cudaStream_t st);
int hex8_weights_modal(const tatva_plan* p, double* out, cudaStream_t st);
int hex8_nh_residual_modal(const tatva_plan* p, double mu, double lmbda, const double* u, double* y, cudaStream_t st);
int hex8_nh_energy_modal_partials(const tatva_plan* p, double mu, double lmbda, const double* u, cudaStream_t st);
int tet4_nh_hvp_ref(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y, cudaStream_t st);
int tet4_nh_residual_ref(const tatva_plan* p, double mu, double lmbda, const double* u, double* y, cudaStream_t st);
int tet4_nh_tiled(const tatva_plan* p, bool hvp, double mu, double lmbda, const double* u, const double* v, double* y, cudaStream_t st);
int tet4_nh_wc(const tatva_plan* p, bool hvp, double mu, double lmbda, const double* u, const double* v, double* y, cudaStream_t st);

}  // namespace tatva
#endif  // !__CUDACC_RTC__
