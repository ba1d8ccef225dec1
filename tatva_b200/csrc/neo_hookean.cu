// Hex8 x compressible neo-Hookean matrix-free HVP — the headline kernel (config 3/4).
//
//   y = d/d eps  r(u + eps v),   r = dE/du,   E = sum_e sum_q W psi(grad u)           (README.md:93,
//   psi = mu/2 (I1 - 3 - 2 ln J) + lambda/2 (ln J)^2,  F = I + grad u                   tests/test_sparse_tracer.py:103-115)
//
// The kernel is bound by the FP64 pipe (AI ~ 60 flop/B), so the design minimises
// FP64 instructions per element rather than bytes:
//
//  * Modal form of the trilinear element.  Each nodal field is taken to its 7 non-constant
//    trilinear coefficients by an 8-point Walsh-Hadamard transform (24 adds); the reference-space
//    gradient at a 2x2x2 Gauss point (+-a,+-a,+-a) (tatva/element/base.py:493-513) is then 3 adds per
//    direction instead of an 8-term dot product with dN/dxi (:531-568), and the scatter is the
//    transposed accumulation into 7 modal residuals followed by one inverse transform.
//  * Everything is kept in reference space.  With J = dX/dxi, Fr = d(X+u)/dxi, A = Fr^-1,
//    K = J^-1, M = K^T K and Gv = dv/dxi:
//        W dP(grad v) K = W [ mu Gv M + (mu - lambda lnJ) (A Gv A)^T + lambda tr(A Gv) A^T ],
//        lnJ = log(det Fr / det J),  W = det J  (all quadrature weights are 1),
//    which needs two 3x3 inverses, four 3x3 products, one log per point and never forms dN/dX.
//  * One thread per element; gather by connectivity through the read-only path; scatter with
//    RED.ADD.F64.  Elements of a lexicographic / locality-sorted mesh put consecutive lanes on
//    consecutive nodes, so the 24-byte nodal rows of a warp share L1 lines.
#include "common.cuh"
#include "wc.cuh"

namespace tatva {

namespace {

constexpr double kA = 0.57735026918962576451;  // 1/sqrt(3)

// Gauss-point signs (tx, ty, tz, ty*tz, tx*tz, tx*ty) in the element's node order.  Read with a
// warp-uniform index so they reach the DFMAs as constant-bank / uniform-register operands: a DFMA with
// three distinct *vector*-register sources issues at 2/3 rate on B200 (tools/micro/dfma_operands.cu).
__constant__ double kSigns[8][6] = {
    {-1, -1, -1, +1, +1, +1}, {+1, -1, -1, +1, -1, -1}, {+1, +1, -1, -1, -1, +1}, {-1, +1, -1, -1, +1, -1},
    {-1, -1, +1, -1, -1, +1}, {+1, -1, +1, -1, +1, -1}, {+1, +1, +1, +1, +1, +1}, {-1, +1, +1, +1, -1, -1}};

// 8-point Walsh-Hadamard transform of the nodal values of one scalar field, in the element's node
// order (bottom CCW, top CCW).  Output: c[0..6] = {x, y, z, xy, yz, zx, xyz} coefficients scaled so
// that   d/dxi f = c_x + ty c_xy + tz c_zx + ty tz c_xyz   at the Gauss point a*(tx,ty,tz).
TATVA_D void to_modal(const double (&f)[8], double (&c)[7]) {
  // natural (bit) order: b = [sx>0] + 2 [sy>0] + 4 [sz>0]  <-  nodes 0,1,3,2,4,5,7,6
  const double s01 = f[0] + f[1], d01 = f[1] - f[0];
  const double s23 = f[3] + f[2], d23 = f[2] - f[3];
  const double s45 = f[4] + f[5], d45 = f[5] - f[4];
  const double s67 = f[7] + f[6], d67 = f[6] - f[7];
  // y stage
  const double ss0 = s01 + s23, ds0 = s23 - s01;  // (x-sum) y-sum / y-diff, lower face
  const double sd0 = d01 + d23, dd0 = d23 - d01;  // (x-diff)
  const double ss1 = s45 + s67, ds1 = s67 - s45;
  const double sd1 = d45 + d67, dd1 = d67 - d45;
  // z stage (the all-sum coefficient is not needed for gradients)
  c[0] = 0.125 * (sd0 + sd1);                // x
  c[1] = 0.125 * (ds0 + ds1);                // y
  c[2] = 0.125 * (ss1 - ss0);                // z
  c[3] = (0.125 * kA) * (dd0 + dd1);         // xy
  c[4] = (0.125 * kA) * (ds1 - ds0);         // yz
  c[5] = (0.125 * kA) * (sd1 - sd0);         // zx
  c[6] = (0.125 * kA * kA) * (dd1 - dd0);    // xyz
}

// transpose of to_modal: modal residuals r[0..6] -> nodal contributions
TATVA_D void from_modal(const double (&r)[7], double (&f)[8]) {
  const double x = 0.125 * r[0], y = 0.125 * r[1], z = 0.125 * r[2];
  const double xy = (0.125 * kA) * r[3], yz = (0.125 * kA) * r[4], zx = (0.125 * kA) * r[5];
  const double xyz = (0.125 * kA * kA) * r[6];
  // value at node with signs (sx,sy,sz):  sx x + sy y + sz z + sx sy xy + sy sz yz + sz sx zx + sx sy sz xyz
  // sz = -1 / +1 halves
  const double xm = x - zx, xp = x + zx;      // sx coefficient for sz = -1 / +1
  const double ym = y - yz, yp = y + yz;      // sy coefficient
  const double xym = xy - xyz, xyp = xy + xyz;  // sx sy coefficient
  // lower face (sz = -1): -z + sx xm + sy ym + sx sy xym
  f[0] = -z - xm - ym + xym;
  f[1] = -z + xm - ym - xym;
  f[2] = -z + xm + ym + xym;
  f[3] = -z - xm + ym - xym;
  f[4] = z - xp - yp + xyp;
  f[5] = z + xp - yp - xyp;
  f[6] = z + xp + yp + xyp;
  f[7] = z - xp + yp - xyp;
}

// reference gradient of a modal field at Gauss point with signs (TX,TY,TZ)
template <int TX, int TY, int TZ>
TATVA_D void ref_grad(const double (&c)[7], double (&g)[3]) {
  g[0] = (c[0] + TZ * c[5]) + TY * (c[3] + TZ * c[6]);
  g[1] = (c[1] + TZ * c[4]) + TX * (c[3] + TZ * c[6]);
  g[2] = (c[2] + TY * c[4]) + TX * (c[5] + TY * c[6]);
}

// transposed accumulation: r += ref_grad^T q
template <int TX, int TY, int TZ>
TATVA_D void ref_grad_T(const double (&q)[3], double (&r)[7]) {
  r[0] += q[0];
  r[1] += q[1];
  r[2] += q[2];
  r[3] += TY * q[0] + TX * q[1];
  r[4] += TZ * q[1] + TY * q[2];
  r[5] += TZ * q[0] + TX * q[2];
  r[6] += (TY * TZ) * q[0] + (TX * TZ) * q[1] + (TX * TY) * q[2];
}

struct Modal {
  double X[3][7];  // coordinates
  double x[3][7];  // coordinates + displacement
  double v[3][7];  // direction
};

template <int TX, int TY, int TZ>
TATVA_D void qp_hvp(const Modal& m, double mu, double lmbda, double (&R)[3][7]) {
  double J[3][3], K[3][3];  // J[d][c] = dX_c / dxi_d
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double g[3];
    ref_grad<TX, TY, TZ>(m.X[c], g);
    J[0][c] = g[0]; J[1][c] = g[1]; J[2][c] = g[2];
  }
  const double detJ = det_inv(J, K);
  double M[3][3];  // K^T K (symmetric)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      M[a][b] = K[0][a] * K[0][b] + K[1][a] * K[1][b] + K[2][a] * K[2][b];
      M[b][a] = M[a][b];
    }
  double Fr[3][3], A[3][3];  // Fr[i][d] = d x_i / d xi_d
#pragma unroll
  for (int i = 0; i < 3; ++i) ref_grad<TX, TY, TZ>(m.x[i], Fr[i]);
  const double detFr = det_inv(Fr, A);
  const double lnJ = log(detFr / detJ);
  double Gv[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) ref_grad<TX, TY, TZ>(m.v[i], Gv[i]);
  double B[3][3];  // A Gv
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int e = 0; e < 3; ++e) B[d][e] = A[d][0] * Gv[0][e] + A[d][1] * Gv[1][e] + A[d][2] * Gv[2][e];
  const double w1 = detJ * mu, w2 = detJ * (mu - lmbda * lnJ), w3 = detJ * lmbda * (B[0][0] + B[1][1] + B[2][2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double t1 = Gv[i][0] * M[0][d] + Gv[i][1] * M[1][d] + Gv[i][2] * M[2][d];
      const double t2 = B[d][0] * A[0][i] + B[d][1] * A[1][i] + B[d][2] * A[2][i];
      q[d] = w1 * t1 + w2 * t2 + w3 * A[d][i];
    }
    ref_grad_T<TX, TY, TZ>(q, R[i]);
  }
}

template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                  double lmbda, const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  Modal m;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], fv[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
      fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
      fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
    }
    to_modal(fX, m.X[c]);
    to_modal(fu, m.x[c]);
    to_modal(fv, m.v[c]);
#pragma unroll
    for (int k = 0; k < 7; ++k) m.x[c][k] += m.X[c][k];
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;

  qp_hvp<-1, -1, -1>(m, mu, lmbda, R);
  qp_hvp<+1, -1, -1>(m, mu, lmbda, R);
  qp_hvp<+1, +1, -1>(m, mu, lmbda, R);
  qp_hvp<-1, +1, -1>(m, mu, lmbda, R);
  qp_hvp<-1, -1, +1>(m, mu, lmbda, R);
  qp_hvp<+1, -1, +1>(m, mu, lmbda, R);
  qp_hvp<+1, +1, +1>(m, mu, lmbda, R);
  qp_hvp<-1, +1, +1>(m, mu, lmbda, R);

#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}


// ---------------------------------------------------------------------------------------------
// Rolled variant: the 8 Gauss points run in a loop with the signs as data.  a + t*b is one DFMA,
// so the FP64 instruction count matches the unrolled form, but nothing is hoisted across points
// (fewer live registers, 8x less code).  STAGE selects which modal fields live in shared memory
// (coefficient-major, one column per thread => conflict-free LDS.64):
//   0: none     1: X and v      2: X, v and x
// Reciprocals of the two determinants are folded into the three scalar weights so the inverses
// stay unscaled adjugates.
// ---------------------------------------------------------------------------------------------

template <class Ptr>
TATVA_D void ref_grad_s(Ptr c, int stride, double tx, double ty, double tz, double (&g)[3]) {
  const double c0 = c[0], c1 = c[stride], c2 = c[2 * stride], c3 = c[3 * stride], c4 = c[4 * stride],
               c5 = c[5 * stride], c6 = c[6 * stride];
  const double s36 = fma(tz, c6, c3);
  g[0] = fma(ty, s36, fma(tz, c5, c0));
  g[1] = fma(tx, s36, fma(tz, c4, c1));
  g[2] = fma(tx, fma(ty, c6, c5), fma(ty, c4, c2));
}

// Branch-free reciprocal and logarithm for the Gauss-point loops.  CUDA's `1.0 / x` and `log(x)` carry slow-path branches
// (denormals, infinities, NaNs) that cut the loop body into several basic blocks; ptxas schedules within a block, so the
// two independent Gauss points of a tx pair could not be interleaved across them and each warp sat in fixed-latency
// dependency stalls ("wait": 45 % of its time inside the loop, profiles/r02_hvp_ncu_stalls.md).  The arguments here are
// determinants of non-degenerate elements (finite, normal, and positive for the logarithm; an inverted element gives a
// NaN either way), so the special cases are not needed.
//   fast_rcp: MUFU.RCP64H seed (rel. error <= 2^-20) + two Newton steps                -> <= 1 ulp
//   log_pos : x = 2^e m, m in [sqrt(1/2), sqrt(2)), f = (m-1)/(m+1), log m = 2 f sum_k f^(2k) / (2k+1), k <= 10
//             (|f| <= 0.1716: truncation 1e-18), relative error ~2e-16 near x = 1 (e = 0), absolute ~1e-16 |e| ln 2 elsewhere.
// The host twins (kernel-arithmetic probes) use the C library.
TATVA_HD double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double t = fma(-x, r, 1.0);
  r = fma(r, t, r);
  t = fma(-x, r, 1.0);
  return fma(r, t, r);
#else
  return 1.0 / x;
#endif
}

TATVA_HD double log_pos(double x) {
#ifdef __CUDA_ARCH__
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;           // m in [1, 2)
  const int up = hi >= 0x3ff6a09f ? 1 : 0;       // m >= sqrt(2) (to 2^-20): halve it
  hi -= up << 20;
  e += up;
  const double m = __hiloint2double(hi, lo);
  const double f = (m - 1.0) * fast_rcp(m + 1.0);
  const double g = f * f;
  double p = 1.0 / 21.0;
  p = fma(p, g, 1.0 / 19.0);
  p = fma(p, g, 1.0 / 17.0);
  p = fma(p, g, 1.0 / 15.0);
  p = fma(p, g, 1.0 / 13.0);
  p = fma(p, g, 1.0 / 11.0);
  p = fma(p, g, 1.0 / 9.0);
  p = fma(p, g, 1.0 / 7.0);
  p = fma(p, g, 1.0 / 5.0);
  p = fma(p, g, 1.0 / 3.0);
  p = p * g;                                       // log m = 2 f (1 + p)
  const double two_f = f + f;
  return fma((double)e, 0.69314718055994530942, fma(two_f, p, two_f));
#else
  return log(x);
#endif
}

TATVA_HD void adjugate(const double (&A)[3][3], double (&C)[3][3], double& det) {
  C[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  C[1][0] = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  C[2][0] = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  det = A[0][0] * C[0][0] + A[0][1] * C[1][0] + A[0][2] * C[2][0];
  C[0][1] = A[0][2] * A[2][1] - A[0][1] * A[2][2];
  C[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0];
  C[2][1] = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  C[0][2] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
  C[1][2] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
  C[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
}

template <int STAGE, int MINB, int DEBUG = 0, int UNROLL = 1, int FOLD = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp_rolled(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                         double lmbda, const double* __restrict__ u, const double* __restrict__ v,
                         double* __restrict__ y) {
  constexpr int NS = STAGE == 0 ? 0 : (STAGE == 1 ? 2 : 3);  // staged fields
  extern __shared__ double sm[];                               // [NS][3][7][kBlock]
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  double rX[STAGE >= 1 ? 1 : 3][7], rv[STAGE >= 1 ? 1 : 3][7], rx[STAGE >= 2 ? 1 : 3][7];
  double* sX0 = sm + threadIdx.x;
  double* sv0 = sm + 21 * kBlock + threadIdx.x;
  double* sx0 = sm + 42 * kBlock + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], fv[8], mX[7], mx[7], mv[7];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if constexpr (DEBUG == 2) {  // timing experiment: no gather
        fX[n] = Hex8::sgn(n, c) * 0.01 + 1e-9 * nd[n];
        fu[n] = 1e-4 * (n + c) + 1e-10 * nd[n];
        fv[n] = 1e-3 * (n - c) + 1e-9 * nd[n];
      } else {
        fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
        fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
        fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
      }
    }
    to_modal(fX, mX);
    to_modal(fu, mx);
    to_modal(fv, mv);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      mx[k] += mX[k];
      if constexpr (STAGE >= 1) {
        sX0[(c * 7 + k) * kBlock] = mX[k];
        sv0[(c * 7 + k) * kBlock] = mv[k];
      } else {
        rX[c][k] = mX[k];
        rv[c][k] = mv[k];
      }
      if constexpr (STAGE >= 2) sx0[(c * 7 + k) * kBlock] = mx[k];
      else rx[c][k] = mx[k];
    }
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;

#pragma unroll UNROLL
  for (int q = 0; q < 8; ++q) {
    const double tx = kSigns[q][0], ty = kSigns[q][1], tz = kSigns[q][2];
    // Opaque per-iteration offset (always 0): the staged coefficients are re-read from shared memory at
    // every Gauss point instead of being hoisted out of the loop (and spilled), while loads within one
    // point stay freely schedulable (unlike `volatile`).
    int opaque = 0;
    asm volatile("" : "+r"(opaque));
    const double* sX = sX0 + opaque;
    const double* sv = sv0 + opaque;
    const double* sx = sx0 + opaque;
    double J[3][3], Kc[3][3], detJ;  // J[d][c] = dX_c/dxi_d ; Kc = adj(J) = detJ * K
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double g[3];
      if constexpr (STAGE >= 1) ref_grad_s(sX + c * 7 * kBlock, kBlock, tx, ty, tz, g);
      else ref_grad_s((const double*)rX[c], 1, tx, ty, tz, g);
      J[0][c] = g[0]; J[1][c] = g[1]; J[2][c] = g[2];
    }
    adjugate(J, Kc, detJ);
    double M[3][3];  // detJ^2 * K^T K
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) {
        M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
        M[b][a] = M[a][b];
      }
    double Fr[3][3], Ac[3][3], detF;  // Ac = adj(Fr) = detF * A
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (STAGE >= 2) ref_grad_s(sx + i * 7 * kBlock, kBlock, tx, ty, tz, Fr[i]);
      else ref_grad_s((const double*)rx[i], 1, tx, ty, tz, Fr[i]);
    }
    adjugate(Fr, Ac, detF);
    const double rJ = 1.0 / detJ, rF = 1.0 / detF;
    const double lnJ = log(detF * rJ);
    double Gv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (STAGE >= 1) ref_grad_s(sv + i * 7 * kBlock, kBlock, tx, ty, tz, Gv[i]);
      else ref_grad_s((const double*)rv[i], 1, tx, ty, tz, Gv[i]);
    }
    double B[3][3];  // Ac Gv = detF * A Gv
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int f = 0; f < 3; ++f) B[d][f] = Ac[d][0] * Gv[0][f] + Ac[d][1] * Gv[1][f] + Ac[d][2] * Gv[2][f];
    // W = detJ:  mu W Gv M/detJ^2 ; (mu - lambda lnJ) W (B Ac)^T/detF^2 ; lambda W tr(B) Ac^T/detF^2
    const double wF = detJ * rF * rF;
    const double w1 = mu * rJ, w2 = (mu - lmbda * lnJ) * wF, w3 = lmbda * wF * (B[0][0] + B[1][1] + B[2][2]);
    const double tyz = kSigns[q][3], txz = kSigns[q][4], txy = kSigns[q][5];
    if constexpr (FOLD) {  // fold the three weights into M and B: one 6-term chain per flux entry
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
          M[a][b] *= w1;
          M[b][a] = M[a][b];
        }
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int f = 0; f < 3; ++f) B[d][f] = (d == f) ? fma(w2, B[d][f], w3) : w2 * B[d][f];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double qv[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if constexpr (FOLD) {
          qv[d] = fma(B[d][0], Ac[0][i], fma(B[d][1], Ac[1][i], fma(B[d][2], Ac[2][i],
                  fma(Gv[i][0], M[0][d], fma(Gv[i][1], M[1][d], Gv[i][2] * M[2][d])))));
        } else {
          const double t1 = Gv[i][0] * M[0][d] + Gv[i][1] * M[1][d] + Gv[i][2] * M[2][d];
          const double t2 = B[d][0] * Ac[0][i] + B[d][1] * Ac[1][i] + B[d][2] * Ac[2][i];
          qv[d] = w1 * t1 + w2 * t2 + w3 * Ac[d][i];
        }
      }
      R[i][0] += qv[0];
      R[i][1] += qv[1];
      R[i][2] += qv[2];
      R[i][3] = fma(ty, qv[0], fma(tx, qv[1], R[i][3]));
      R[i][4] = fma(tz, qv[1], fma(ty, qv[2], R[i][4]));
      R[i][5] = fma(tz, qv[0], fma(tx, qv[2], R[i][5]));
      R[i][6] = fma(tyz, qv[0], fma(txz, qv[1], fma(txy, qv[2], R[i][6])));
    }
  }

  if constexpr (DEBUG == 1) {  // timing experiment: no scatter
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double f[8];
      from_modal(R[i], f);
#pragma unroll
      for (int n = 0; n < 8; ++n) s += f[n] * (n + 1 + i);
    }
    y[e] = s;
    return;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}

template <int STAGE, int MINB, int DEBUG = 0, int UNROLL = 1, int FOLD = 0>
int launch_rolled(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                  cudaStream_t st) {
  constexpr int NS = STAGE == 0 ? 0 : (STAGE == 1 ? 2 : 3);
  constexpr size_t smem = (size_t)NS * 21 * kBlock * sizeof(double);
  static SmemOptIn configured;
  if (smem > 48 * 1024) {
    const int rc = opt_in_smem(k_hex8_nh_hvp_rolled<STAGE, MINB, DEBUG, UNROLL, FOLD>, smem, configured);
    if (rc != TATVA_OK) return rc;
  }
  k_hex8_nh_hvp_rolled<STAGE, MINB, DEBUG, UNROLL, FOLD><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu,
                                                                               lmbda, u, v, y);
  return TATVA_OK;
}


// ---------------------------------------------------------------------------------------------
// v2 of the rolled HVP kernel: every scale factor is folded away.
//  * to_modal_raw / from_modal_raw are pure +-1 butterflies (no multiplies); the 1/8 of the shape
//    functions and the Gauss abscissa a live in the sign table (s = a t) and in one factor 1/512 applied to
//    the three scalar weights (gradients are carried as 8x their value; the tangent is homogeneous).
//  * 3x3 products are written k-outer so that consecutive DFMAs share their first operand (operand reuse).
// ---------------------------------------------------------------------------------------------
__constant__ double kScaledSigns[8][6] = {
#define TATVA_S(tx, ty, tz) {tx * kA, ty * kA, tz * kA, (ty * tz) * kA * kA, (tx * tz) * kA * kA, (tx * ty) * kA * kA}
    TATVA_S(-1, -1, -1), TATVA_S(+1, -1, -1), TATVA_S(+1, +1, -1), TATVA_S(-1, +1, -1),
    TATVA_S(-1, -1, +1), TATVA_S(+1, -1, +1), TATVA_S(+1, +1, +1), TATVA_S(-1, +1, +1)
#undef TATVA_S
};

TATVA_HD void to_modal_raw(const double (&f)[8], double (&h)[7]) {
  const double s01 = f[0] + f[1], d01 = f[1] - f[0];
  const double s23 = f[3] + f[2], d23 = f[2] - f[3];
  const double s45 = f[4] + f[5], d45 = f[5] - f[4];
  const double s67 = f[7] + f[6], d67 = f[6] - f[7];
  const double ss0 = s01 + s23, ds0 = s23 - s01, sd0 = d01 + d23, dd0 = d23 - d01;
  const double ss1 = s45 + s67, ds1 = s67 - s45, sd1 = d45 + d67, dd1 = d67 - d45;
  h[0] = sd0 + sd1;  // x
  h[1] = ds0 + ds1;  // y
  h[2] = ss1 - ss0;  // z
  h[3] = dd0 + dd1;  // xy
  h[4] = ds1 - ds0;  // yz
  h[5] = sd1 - sd0;  // zx
  h[6] = dd1 - dd0;  // xyz
}

TATVA_HD void from_modal_raw(const double (&r)[7], double (&f)[8]) {
  const double xm = r[0] - r[5], xp = r[0] + r[5];
  const double ym = r[1] - r[4], yp = r[1] + r[4];
  const double xym = r[3] - r[6], xyp = r[3] + r[6];
  const double a0 = xym - r[2], a1 = -xym - r[2];  // sx sy = +1 / -1 on the lower face
  const double b0 = xm + ym, b1 = xm - ym;
  f[0] = a0 - b0;
  f[1] = a1 + b1;
  f[2] = a0 + b0;
  f[3] = a1 - b1;
  const double c0 = xyp + r[2], c1 = r[2] - xyp;
  const double e0 = xp + yp, e1 = xp - yp;
  f[4] = c0 - e0;
  f[5] = c1 + e1;
  f[6] = c0 + e0;
  f[7] = c1 - e1;
}

TATVA_D void ref_grad8(const double (&h)[7], double sx, double sy, double sz, double (&g)[3]) {
  const double s36 = fma(sz, h[6], h[3]);
  g[0] = fma(sy, s36, fma(sz, h[5], h[0]));
  g[1] = fma(sx, s36, fma(sz, h[4], h[1]));
  g[2] = fma(sx, fma(sy, h[6], h[5]), fma(sy, h[4], h[2]));
}

// C = A * B (3x3), written k-outer: consecutive DFMAs share A[i][k]
TATVA_HD void mat3(const double (&A)[3][3], const double (&Bm)[3][3], double (&C)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * Bm[0][j];
#pragma unroll
  for (int k = 1; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) C[i][j] = fma(A[i][k], Bm[k][j], C[i][j]);
}

template <int MINB, int STAGE = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp_v2(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                     double lmbda, const double* __restrict__ u, const double* __restrict__ v,
                     double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  // STAGE 1: the modal coordinates and direction live in shared memory (coefficient-major, one column
  // per thread: conflict-free), which frees 84 registers => 3 CTAs (12 warps) per SM instead of 2.
  extern __shared__ double sm[];
  double* sX0 = sm + threadIdx.x;
  double* sv0 = sm + 21 * kBlock + threadIdx.x;
  double hX[STAGE ? 1 : 3][7], hx[3][7], hv[STAGE ? 1 : 3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], fv[8], tX[7], tv[7];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
      fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
      fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
    }
    to_modal_raw(fX, tX);
    to_modal_raw(fu, hx[c]);
    to_modal_raw(fv, tv);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      hx[c][k] += tX[k];
      if constexpr (STAGE) {
        sX0[(c * 7 + k) * kBlock] = tX[k];
        sv0[(c * 7 + k) * kBlock] = tv[k];
      } else {
        hX[c][k] = tX[k];
        hv[c][k] = tv[k];
      }
    }
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);

#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const double sx = kScaledSigns[q][0], sy = kScaledSigns[q][1], sz = kScaledSigns[q][2];
    int opaque = 0;  // always 0; keeps the staged loads inside the loop (see k_hex8_nh_hvp_rolled)
    asm volatile("" : "+r"(opaque));
    const double* sX = sX0 + opaque;
    const double* sv = sv0 + opaque;
    double J[3][3], Kc[3][3], detJ;  // 8 dX_c/dxi_d ; adj
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double g[3];
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sX[(c * 7 + k) * kBlock];
        ref_grad8(t, sx, sy, sz, g);
      } else {
        ref_grad8(hX[c], sx, sy, sz, g);
      }
      J[0][c] = g[0]; J[1][c] = g[1]; J[2][c] = g[2];
    }
    adjugate(J, Kc, detJ);
    double M[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) {
        M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
        M[b][a] = M[a][b];
      }
    double Fr[3][3], Ac[3][3], detF;
#pragma unroll
    for (int i = 0; i < 3; ++i) ref_grad8(hx[i], sx, sy, sz, Fr[i]);
    adjugate(Fr, Ac, detF);
    const double r = 1.0 / (detJ * detF);
    const double rJ = r * detF, rF = r * detJ;
    const double lnJ = log(detF * rJ);
    double Gv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sv[(i * 7 + k) * kBlock];
        ref_grad8(t, sx, sy, sz, Gv[i]);
      } else {
        ref_grad8(hv[i], sx, sy, sz, Gv[i]);
      }
    }
    double B[3][3];
    mat3(Ac, Gv, B);
    const double wF = detJ * rF * rF;
    const double w1 = mu_s * rJ, w2 = (mu_s - lm_s * lnJ) * wF, w3 = lm_s * wF * (B[0][0] + B[1][1] + B[2][2]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) {
        M[a][b] *= w1;
        M[b][a] = M[a][b];
      }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int f = 0; f < 3; ++f) B[d][f] = (d == f) ? fma(w2, B[d][f], w3) : w2 * B[d][f];
    double T1[3][3], T2[3][3];  // T1 = Gv M' ; T2 = B' Ac ; flux Q[i][d] = T1[i][d] + T2[d][i]
    mat3(Gv, M, T1);
    mat3(B, Ac, T2);
    const double syz = kScaledSigns[q][3], sxz = kScaledSigns[q][4], sxy = kScaledSigns[q][5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double q0 = T1[i][0] + T2[0][i], q1 = T1[i][1] + T2[1][i], q2 = T1[i][2] + T2[2][i];
      R[i][0] += q0;
      R[i][1] += q1;
      R[i][2] += q2;
      R[i][3] = fma(sy, q0, fma(sx, q1, R[i][3]));
      R[i][4] = fma(sz, q1, fma(sy, q2, R[i][4]));
      R[i][5] = fma(sz, q0, fma(sx, q2, R[i][5]));
      R[i][6] = fma(syz, q0, fma(sxz, q1, fma(sxy, q2, R[i][6])));
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal_raw(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}

// ---------------------------------------------------------------------------------------------
// v3: the two Gauss points of a tx = -/+ pair are processed together.  They share the xi-derivative and the
// (ty, tz) partial sums of every field (10 fused ops per field per pair instead of 16) and the modal
// accumulation (sums / differences of the two fluxes: 18 ops per component per pair instead of 24).
// ---------------------------------------------------------------------------------------------
// (sy, sz, sy*sz) for (ty, tz) = (-,-), (+,-), (-,+), (+,+); one initializer for the device table and its host twin
#define TATVA_PAIR_SIGNS \
  { {-kA, -kA, kA * kA}, {kA, -kA, -kA * kA}, {-kA, kA, -kA * kA}, {kA, kA, kA * kA} }
__constant__ double kPairSigns[4][3] = TATVA_PAIR_SIGNS;
static const double kPairSignsHost[4][3] = TATVA_PAIR_SIGNS;  // for the host probe (tatva_probe_hex8_nh_modal)

TATVA_HD void ref_grad8_pair(const double (&h)[7], double sy, double sz, double (&gm)[3], double (&gp)[3]) {
  const double s36 = fma(sz, h[6], h[3]);
  const double g0 = fma(sy, s36, fma(sz, h[5], h[0]));
  const double b1 = fma(sz, h[4], h[1]);
  const double s56 = fma(sy, h[6], h[5]);
  const double b2 = fma(sy, h[4], h[2]);
  gm[0] = g0;
  gp[0] = g0;
  gm[1] = fma(-kA, s36, b1);
  gp[1] = fma(kA, s36, b1);
  gm[2] = fma(-kA, s56, b2);
  gp[2] = fma(kA, s56, b2);
}

// flux Q[i][d] (scaled by 512) of one Gauss point from J (8 dX/dxi, [d][c]), Fr, Gv ([i][d])
#ifndef TATVA_FLUX_CHAINS
#define TATVA_FLUX_CHAINS 0
#endif
TATVA_HD void point_flux(const double (&J)[3][3], const double (&Fr)[3][3], const double (&Gv)[3][3], double mu_s,
                        double lm_s, double (&Q)[3][3]) {
  double Kc[3][3], detJ, Ac[3][3], detF;
  adjugate(J, Kc, detJ);
  double M[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
      M[b][a] = M[a][b];
    }
  adjugate(Fr, Ac, detF);
  const double r = fast_rcp(detJ * detF);
  const double rJ = r * detF, rF = r * detJ;
  const double lnJ = log_pos(detF * rJ);
  double B[3][3];
  mat3(Ac, Gv, B);
  const double wF = detJ * rF * rF;
  const double w1 = mu_s * rJ, w2 = (mu_s - lm_s * lnJ) * wF, w3 = lm_s * wF * (B[0][0] + B[1][1] + B[2][2]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      M[a][b] *= w1;
      M[b][a] = M[a][b];
    }
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int f = 0; f < 3; ++f) B[d][f] = (d == f) ? fma(w2, B[d][f], w3) : w2 * B[d][f];
#if TATVA_FLUX_CHAINS
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      Q[i][d] = fma(B[d][0], Ac[0][i], fma(B[d][1], Ac[1][i], fma(B[d][2], Ac[2][i],
                fma(Gv[i][0], M[0][d], fma(Gv[i][1], M[1][d], Gv[i][2] * M[2][d])))));
#else
  // Q = Gv M' + (B' Ac)^T accumulated k-outer: three consecutive DFMAs share one vector-register operand (operand
  // reuse), which is what a DFMA with three distinct register sources needs to issue at full rate
  // (tools/micro/fp64_mix.cu: 3.0 cycles with three distinct register sources, 2.4 with one of them reused, 2.0 with two).
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Q[i][d] = Gv[i][0] * M[0][d];
#pragma unroll
  for (int k = 1; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) Q[i][d] = fma(Gv[i][k], M[k][d], Q[i][d]);
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < 3; ++i) Q[i][d] = fma(B[d][k], Ac[k][i], Q[i][d]);
#endif
}

// Transposed reference gradient of a tx pair: the fluxes of the two Gauss points (xi = -a and +a) enter the 7 modal
// residuals through their sums and differences (18 fused ops per component instead of 24).
TATVA_HD void accumulate_pair(const double (&Qm)[3][3], const double (&Qp)[3][3], double sy, double sz, double syz,
                              double (&R)[3][7]) {
  const double asz = kA * sz, asy = kA * sy;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double S0 = Qp[i][0] + Qm[i][0], S1 = Qp[i][1] + Qm[i][1], S2 = Qp[i][2] + Qm[i][2];
    const double D1 = Qp[i][1] - Qm[i][1], D2 = Qp[i][2] - Qm[i][2];
    R[i][0] += S0;
    R[i][1] += S1;
    R[i][2] += S2;
    R[i][3] = fma(sy, S0, fma(kA, D1, R[i][3]));
    R[i][4] = fma(sz, S1, fma(sy, S2, R[i][4]));
    R[i][5] = fma(sz, S0, fma(kA, D2, R[i][5]));
    R[i][6] = fma(syz, S0, fma(asz, D1, fma(asy, D2, R[i][6])));
  }
}

// LIFT: v and y are REDUCED vectors and `map` (n_nodes*3 int32) sends a full DOF to its reduced index, or -1 for a
// DOF without a driver (Fixed): the homogeneous lift and reduce_adjoint of the Lifter folded into the gather / scatter.
// One aligned double2 + one double per 3-double row (see load_row in common.cuh), regardless of TATVA_WIDE_GATHER.
TATVA_D void wide_row(const double* __restrict__ src, int64_t node, double (&dst)[3]) {
  const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(src) >> 3) & 1u);
  const int64_t o = node * 3;
  const unsigned odd = ((unsigned)node + mis) & 1u;
  const double2 v = __ldg(reinterpret_cast<const double2*>(src + o + odd));
  const double s = __ldg(src + o + (odd ? 0 : 2));
  dst[0] = odd ? s : v.x;
  dst[1] = odd ? v.x : v.y;
  dst[2] = odd ? v.y : s;
}

// WIDE: node rows are fetched whole (one double2 + one double each) array by array, instead of component by component
// with 8-byte loads: 48 instead of 72 gather instructions per element, but ~400 more integer / select instructions.
// Measured slower (0.490 vs 0.453 ms, variant 28), so the default stays component-wise.
template <int MINB, int STAGE, int GROUPED = 0, bool LIFT = false, int WIDE = 0, int UNR = 1, int CPF = 0, int NDS = 0, int DOT = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp_v3(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                     double lmbda, const double* __restrict__ u, const double* __restrict__ v,
                     double* __restrict__ y, const int32_t* __restrict__ map = nullptr,
                     double* __restrict__ dot_partials = nullptr) {
  static_assert(!(LIFT && GROUPED), "the lifted scatter is per DOF");
  static_assert(!DOT || (STAGE >= 1 && !GROUPED), "the fused v.Hv needs the modal direction in shared memory");
  const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e0 < E;
  if (!GROUPED && !DOT && !valid) return;
  const int64_t e = valid ? e0 : E - 1;  // GROUPED: out-of-range lanes redo the last element and drop the result
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  if constexpr (CPF > 0) {
    // The connectivity is a pure stream and the first of the two dependent round trips of the gather: pull the lines
    // of the CTA that will run CPF CTAs from now (about one wave later on this SM) into L2 with one instruction.
    const int64_t ea = e + (int64_t)CPF * kBlock;
    if (ea < E) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const int4*>(conn) + 2 * ea));
  }
  extern __shared__ double sm[];
  double* sX0 = sm + threadIdx.x;
  double* sv0 = sm + 21 * kBlock + threadIdx.x;
  double* sx0 = sm + 42 * kBlock + threadIdx.x;
  double hX[STAGE ? 1 : 3][7], hx[STAGE >= 2 ? 1 : 3][7], hv[STAGE ? 1 : 3][7];
  if constexpr (WIDE) {
    double tX[3][7];
    {
      double row[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n) wide_row(coords, nd[n], row[n]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double f[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = row[n][c];
        to_modal_raw(f, tX[c]);
      }
    }
    {
      double row[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n) wide_row(u, nd[n], row[n]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double f[8], t[7];
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = row[n][c];
        to_modal_raw(f, t);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          t[k] += tX[c][k];
          if constexpr (STAGE >= 2) sx0[(c * 7 + k) * kBlock] = t[k];
          else hx[c][k] = t[k];
        }
      }
    }
    {
      double row[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if constexpr (LIFT) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int32_t m = __ldg(map + (int64_t)nd[n] * 3 + c);
            row[n][c] = m >= 0 ? __ldg(v + m) : 0.0;
          }
        } else {
          wide_row(v, nd[n], row[n]);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double f[8], t[7];
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = row[n][c];
        to_modal_raw(f, t);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          if constexpr (STAGE) sv0[(c * 7 + k) * kBlock] = t[k];
          else hv[c][k] = t[k];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        if constexpr (STAGE) sX0[(c * 7 + k) * kBlock] = tX[c][k];
        else hX[c][k] = tX[c][k];
      }
  } else {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], fv[8], tX[7], tv[7], tx_[7];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
      fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
      if constexpr (LIFT) {
        const int32_t m = __ldg(map + (int64_t)nd[n] * 3 + c);
        fv[n] = m >= 0 ? __ldg(v + m) : 0.0;
      } else {
        fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
      }
    }
    to_modal_raw(fX, tX);
    to_modal_raw(fu, tx_);
    to_modal_raw(fv, tv);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      tx_[k] += tX[k];
      if constexpr (STAGE >= 2) sx0[(c * 7 + k) * kBlock] = tx_[k];
      else hx[c][k] = tx_[k];
      if constexpr (STAGE) {
        sX0[(c * 7 + k) * kBlock] = tX[k];
        sv0[(c * 7 + k) * kBlock] = tv[k];
      } else {
        hX[c][k] = tX[k];
        hv[c][k] = tv[k];
      }
    }
  }
  }
  // NDS: the node ids are only needed again by the scatter; park them in shared memory across the Gauss-point loop
  // (8 registers less to carry through it: the 168-register / 3-CTA build then runs without spills).
  int4* snd = reinterpret_cast<int4*>(sm + (size_t)(STAGE == 0 ? 0 : (STAGE == 1 ? 42 : 63)) * kBlock) + threadIdx.x;
  if constexpr (NDS) {
    snd[0] = make_int4(nd[0], nd[1], nd[2], nd[3]);
    snd[kBlock] = make_int4(nd[4], nd[5], nd[6], nd[7]);
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
  const double mu_s = mu, lm_s = lmbda;  // already scaled by 1/512 on the host: used straight from the constant bank

#pragma unroll UNR
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
    int opaque = 0;
    asm volatile("" : "+r"(opaque));
    const double* sX = sX0 + opaque;
    const double* sv = sv0 + opaque;
    const double* sx = sx0 + opaque;
    double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3], Gvm[3][3], Gvp[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sX[(c * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, gm, gp);
      } else {
        ref_grad8_pair(hX[c], sy, sz, gm, gp);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (STAGE >= 2) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sx[(i * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, Frm[i], Frp[i]);
      } else {
        ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
      }
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sv[(i * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, Gvm[i], Gvp[i]);
      } else {
        ref_grad8_pair(hv[i], sy, sz, Gvm[i], Gvp[i]);
      }
    }
    double Qm[3][3], Qp[3][3];
    point_flux(Jm, Frm, Gvm, mu_s, lm_s, Qm);
    point_flux(Jp, Frp, Gvp, mu_s, lm_s, Qp);
    accumulate_pair(Qm, Qp, sy, sz, syz, R);
  }
  if constexpr (NDS) {
    const int4 a = snd[0], b = snd[kBlock];
    nd[0] = a.x; nd[1] = a.y; nd[2] = a.z; nd[3] = a.w;
    nd[4] = b.x; nd[5] = b.y; nd[6] = b.z; nd[7] = b.w;
  }
  // Programmatic dependent launch: when the launcher overlaps this grid with the kernel that zeroes y (launch_v3 with
  // pdl = true), everything above ran while y was still being cleared; the scatter must wait for it.  A no-op otherwise.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if constexpr (GROUPED) {
    // sector-grouped scatter: consecutive lanes add the 3 consecutive doubles of one node
    constexpr int NS = STAGE == 0 ? 0 : (STAGE == 1 ? 42 : 63);
    double* wsm = sm + (size_t)NS * kBlock + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<8, 3>();
    const int lane = threadIdx.x & 31;
    int* snode = reinterpret_cast<int*>(wsm + 32 * 25);  // odd strides (25 doubles, 9 ints): conflict-free staging
#pragma unroll
    for (int n = 0; n < 8; ++n) snode[lane * 9 + n] = valid ? nd[n] : -1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double f[8];
      from_modal_raw(R[i], f);
#pragma unroll
      for (int n = 0; n < 8; ++n) wsm[lane * 25 + n * 3 + i] = f[n];
    }
    __syncwarp();
    for (int t = lane; t < 32 * 24; t += 32) {
      const int j = t / 24, r = t - j * 24;
      const int node = snode[j * 9 + r / 3];
      if (node >= 0) atomicAdd(y + (int64_t)node * 3 + (r % 3), wsm[j * 25 + r]);
    }
  } else {
    // DOT: v . (H v) summed element by element on the way to the scatter: the 8-point transforms are transposes of each
    // other, so sum_n v_e[n] y_e[n] = sum_k modal(v)[k] R[k].  With a Lifter this is v_red . y_red exactly (a Fixed DOF has
    // v = 0, a Periodic image reads its master's entry and adds into it).  One partial per warp, fixed reduction order.
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (DOT) {
        // the opaque offset pins the 7 loads of this row behind the REDs of the previous one: hoisted to the top of the
        // epilogue (what ptxas does otherwise) they cost 21 live doubles next to R and 220 bytes of spill
        int opq = 0;
        asm volatile("" : "+r"(opq));
        const double* svr = sv0 + opq;
#pragma unroll
        for (int k = 0; k < 7; ++k) dot = fma(svr[(i * 7 + k) * kBlock], R[i][k], dot);
      }
      double f[8];
      from_modal_raw(R[i], f);
      if (DOT && !valid) continue;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if constexpr (LIFT) {
          const int32_t m = __ldg(map + (int64_t)nd[n] * 3 + i);
          if (m >= 0) atomicAdd(y + m, f[n]);
        } else {
          atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
        }
      }
    }
    if constexpr (DOT) {
      if (!valid) dot = 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
      // one partial per CTA (fixed order: warp 0..3): the final sum then reads 16 K values, not 64 K
      __shared__ double sdot[kBlock / 32];
      if ((threadIdx.x & 31) == 0) sdot[threadIdx.x >> 5] = dot;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = sdot[0];
#pragma unroll
        for (int w = 1; w < kBlock / 32; ++w) t += sdot[w];
        dot_partials[blockIdx.x] = t;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Geometry cache (r02): the HVP is bound by FP64 issue while HBM sits at 9 % of its peak, and a fifth of the loop's FP64
// instructions only re-derive what depends on the MESH alone — J = dX/dxi at the Gauss point, its adjugate and
// determinant, M = K^T K.  tatva_plan_cache_geometry computes them once per plan (as Operator(cache_weights=True) keeps
// det J, tatva/operator.py:119-130): per Gauss point 8 doubles
//     Mg = (1 / det J) adj(J)^T adj(J)  (6, symmetric: 00 01 02 11 12 22),  det J,  1 / det J
// in the scaled units of the pair kernels (J = 8 dX/dxi), stored element-fastest as double2 so that a warp reads 512
// contiguous bytes per load:  geo[((pq * 2 + s) * 4 + k) * stride + e],  pq = pair iteration, s = xi sign (-, +).
// 512 bytes per element per application (1.07 GB at 128^3) bought for 64 of 305 FP64 instructions per Gauss point.
// ---------------------------------------------------------------------------------------------
TATVA_HD void point_geometry(const double (&J)[3][3], double (&g)[8]) {
  double Kc[3][3], detJ;
  adjugate(J, Kc, detJ);
  const double rJ = fast_rcp(detJ);
  int t = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) g[t++] = (Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b]) * rJ;
  g[6] = detJ;
  g[7] = rJ;
}

__global__ void __launch_bounds__(kBlock) k_hex8_geometry(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                          int64_t E, double2* __restrict__ geo, int64_t stride) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  double hX[3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double f[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) f[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
    to_modal_raw(f, hX[c]);
  }
#pragma unroll 1
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1];
    double Jm[3][3], Jp[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      ref_grad8_pair(hX[c], sy, sz, gm, gp);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
    double g[8];
    point_geometry(Jm, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) geo[((pq * 2 + 0) * 4 + k) * stride + e] = make_double2(g[2 * k], g[2 * k + 1]);
    point_geometry(Jp, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) geo[((pq * 2 + 1) * 4 + k) * stride + e] = make_double2(g[2 * k], g[2 * k + 1]);
  }
}

// point_flux with the mesh-only part read from the cache: g = {Mg00, Mg01, Mg02, Mg11, Mg12, Mg22, det J, 1 / det J}
TATVA_HD void point_flux_geo(const double (&g)[8], const double (&Fr)[3][3], const double (&Gv)[3][3], double mu_s,
                            double lm_s, double (&Q)[3][3]) {
  double Ac[3][3], detF;
  adjugate(Fr, Ac, detF);
  const double rF = fast_rcp(detF);
  const double lnJ = log_pos(detF * g[7]);
  double B[3][3];
  mat3(Ac, Gv, B);
  const double wF = g[6] * rF * rF;
  const double w2 = (mu_s - lm_s * lnJ) * wF, w3 = lm_s * wF * (B[0][0] + B[1][1] + B[2][2]);
  double M[3][3];
  M[0][0] = mu_s * g[0];
  M[0][1] = M[1][0] = mu_s * g[1];
  M[0][2] = M[2][0] = mu_s * g[2];
  M[1][1] = mu_s * g[3];
  M[1][2] = M[2][1] = mu_s * g[4];
  M[2][2] = mu_s * g[5];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int f = 0; f < 3; ++f) B[d][f] = (d == f) ? fma(w2, B[d][f], w3) : w2 * B[d][f];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Q[i][d] = Gv[i][0] * M[0][d];
#pragma unroll
  for (int k = 1; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) Q[i][d] = fma(Gv[i][k], M[k][d], Q[i][d]);
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < 3; ++i) Q[i][d] = fma(B[d][k], Ac[k][i], Q[i][d]);
}

// The pair kernel on the cached geometry.  STAGE: 0 = modal x and v in registers, 1 = v staged in shared memory,
// 2 = x and v staged (42 doubles per thread).
// PF: 1 = streaming (evict-first) loads of the cache, 2 = + the thread's 32 cache rows prefetched into L2 at kernel start
// (one prefetch per 128-byte line: 12 warps x 16 KB in flight per SM instead of 12 x 4 KB)
template <int MINB, int STAGE, int PF = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp_geo(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu, double lmbda,
                      const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ y,
                      const double2* __restrict__ geo, int64_t stride) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  const double2* ge = geo + e;
  if constexpr (PF >= 2) {
    if ((threadIdx.x & 7) == 0) {
#pragma unroll
      for (int r = 0; r < 32; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(ge + r * stride));
    }
  }
  extern __shared__ double sm[];
  double* sv0 = sm + threadIdx.x;
  double* sx0 = sm + 21 * kBlock + threadIdx.x;
  double hx[STAGE >= 2 ? 1 : 3][7], hv[STAGE ? 1 : 3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fx[8], fv[8], tx_[7], tv[7];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fx[n] = __ldg(coords + (int64_t)nd[n] * 3 + c) + __ldg(u + (int64_t)nd[n] * 3 + c);  // deformed coordinates: ONE transform
      fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
    }
    to_modal_raw(fx, tx_);
    to_modal_raw(fv, tv);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      if constexpr (STAGE >= 2) sx0[(c * 7 + k) * kBlock] = tx_[k];
      else hx[c][k] = tx_[k];
      if constexpr (STAGE) sv0[(c * 7 + k) * kBlock] = tv[k];
      else hv[c][k] = tv[k];
    }
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
#pragma unroll 1
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
    int opaque = 0;
    asm volatile("" : "+r"(opaque));
    const double* sv = sv0 + opaque;
    const double* sx = sx0 + opaque;
    double gm[8], gp[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 a = PF ? __ldcs(ge + ((pq * 2 + 0) * 4 + k) * stride) : __ldg(ge + ((pq * 2 + 0) * 4 + k) * stride);
      const double2 b = PF ? __ldcs(ge + ((pq * 2 + 1) * 4 + k) * stride) : __ldg(ge + ((pq * 2 + 1) * 4 + k) * stride);
      gm[2 * k] = a.x; gm[2 * k + 1] = a.y;
      gp[2 * k] = b.x; gp[2 * k + 1] = b.y;
    }
    double Frm[3][3], Frp[3][3], Gvm[3][3], Gvp[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if constexpr (STAGE >= 2) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sx[(i * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, Frm[i], Frp[i]);
      } else {
        ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
      }
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sv[(i * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, Gvm[i], Gvp[i]);
      } else {
        ref_grad8_pair(hv[i], sy, sz, Gvm[i], Gvp[i]);
      }
    }
    double Qm[3][3], Qp[3][3];
    point_flux_geo(gm, Frm, Gvm, mu, lmbda, Qm);
    point_flux_geo(gp, Frp, Gvp, mu, lmbda, Qp);
    accumulate_pair(Qm, Qp, sy, sz, syz, R);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal_raw(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}

// ---------------------------------------------------------------------------------------------
// Operator.grad and its adjoint on Hex8 in modal form (r02; tatva/operator.py:379-397 -> element/base.py:99-115).
// The generic building block forms dN/dX = inv(J) dN/dxi (72 fused operations per point), J = dN/dxi X (72) and the
// gradient sum over the 8 nodes (24 per component) — 2 300 FP64 instructions per element for 3 components, which keeps
// a 1.27 GB streaming kernel at a third of the HBM roof.  Here, as in the pair kernels above: one Walsh-Hadamard transform
// per nodal field, reference gradients of a tx pair from the 7 modal coefficients (10 fused operations per field and
// pair), K = adj(J) / det J once per point, and grad u = (du/dxi) K: ~95 instructions per point.  The raw transforms carry
// 8 x the reference gradients on both J and du/dxi, which cancels in the product.  Gauss points of pair iteration pq, in
// the element's point order (kSigns): (-xi, +xi) = (0,1), (3,2), (4,5), (7,6).
// ---------------------------------------------------------------------------------------------
TATVA_D void warp_rows_store(double* __restrict__ dst, const double* st, int S, int CH, int count) {
  const int lane = threadIdx.x & 31;
  for (int t = lane; t < count * CH; t += 32) {
    const int j = t / CH;
    dst[t] = st[j * S + (t - j * CH)];
  }
}
// Full warps take the unrolled path: 8 independent 256-byte loads in flight per warp (the rolled loop has one load in
// flight at a time, its store waiting on it: a streaming read then runs at the latency, not the bandwidth, of HBM).
template <int CH>
TATVA_D void warp_rows_load(const double* __restrict__ src, double* st, int S, int count) {
  const int lane = threadIdx.x & 31;
  if (count == 32) {
#pragma unroll 8
    for (int i = 0; i < CH; ++i) {
      const int t = lane + 32 * i, j = t / CH;
      st[j * S + (t - j * CH)] = __ldcs(src + t);
    }
    return;
  }
  for (int t = lane; t < count * CH; t += 32) {
    const int j = t / CH;
    st[j * S + (t - j * CH)] = __ldg(src + t);
  }
}

// K = J^-1 (J[d][c] = dX_c / dxi_d, raw scaling) : K[j][d]
TATVA_HD void inverse_of(const double (&J)[3][3], double (&K)[3][3]) {
  double det;
  adjugate(J, K, det);
  const double r = fast_rcp(det);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) K[a][b] *= r;
}

// Operator.get_integration_weights on Hex8 (tatva/operator.py:172-192): W[e][q] = det J(xi_q) (all weights are 1), from
// the modal coordinates; the raw transform carries 8 J, so det(8 J) / 512.
__global__ void __launch_bounds__(kBlock) k_hex8_weights_modal(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                               int64_t E, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  double hX[3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double f[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) f[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
    to_modal_raw(f, hX[c]);
  }
  double w[8];
#pragma unroll
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1];
    const int qm = (pq == 0) ? 0 : (pq == 1) ? 3 : (pq == 2) ? 4 : 7, qp = (pq == 0) ? 1 : (pq == 1) ? 2 : (pq == 2) ? 5 : 6;
    double Jm[3][3], Jp[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      ref_grad8_pair(hX[c], sy, sz, gm, gp);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
    w[qm] = (1.0 / 512.0) * (Jm[0][0] * (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) + Jm[0][1] * (Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2]) + Jm[0][2] * (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]));
    w[qp] = (1.0 / 512.0) * (Jp[0][0] * (Jp[1][1] * Jp[2][2] - Jp[1][2] * Jp[2][1]) + Jp[0][1] * (Jp[1][2] * Jp[2][0] - Jp[1][0] * Jp[2][2]) + Jp[0][2] * (Jp[1][0] * Jp[2][1] - Jp[1][1] * Jp[2][0]));
  }
  double2* o = reinterpret_cast<double2*>(out + e * 8);  // 64-byte rows of a 256-byte aligned array
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = make_double2(w[2 * k], w[2 * k + 1]);
}

template <int NV, bool ADJOINT>
__global__ void __launch_bounds__(kBlock) k_hex8_grad_modal(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                            int64_t E, const double* __restrict__ in, double* __restrict__ out) {
  // forward: in = u (N, NV), out = grad (E, 8, NV, 3);  adjoint: in = g (E, 8, NV, 3), out = y (N, NV), accumulated
  extern __shared__ double sm[];
  constexpr int CH = 8 * NV * 3, S = CH | 1;
  const int lane = threadIdx.x & 31;
  double* st = sm + (size_t)(threadIdx.x >> 5) * 32 * S;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = e - lane;
  if constexpr (ADJOINT) {
    if (e0 < E) warp_rows_load<CH>(in + e0 * CH, st, S, (int)min((int64_t)32, E - e0));
    __syncwarp();
  }
  const bool valid = e < E;
  const int64_t ee = valid ? e : E - 1;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * ee);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * ee + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  double hX[3][7], hu[NV][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double f[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) f[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
    to_modal_raw(f, hX[c]);
  }
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    if constexpr (ADJOINT) {
#pragma unroll
      for (int k = 0; k < 7; ++k) hu[c][k] = 0.0;  // modal residuals
    } else {
      double f[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) f[n] = __ldg(in + (int64_t)nd[n] * NV + c);
      to_modal_raw(f, hu[c]);
    }
  }
  double* row = st + lane * S;
#pragma unroll 1
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
    const int qm = (pq == 0) ? 0 : (pq == 1) ? 3 : (pq == 2) ? 4 : 7, qp = (pq == 0) ? 1 : (pq == 1) ? 2 : (pq == 2) ? 5 : 6;
    double Jm[3][3], Jp[3][3], Km[3][3], Kp[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      ref_grad8_pair(hX[c], sy, sz, gm, gp);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
    inverse_of(Jm, Km);
    inverse_of(Jp, Kp);
    const double asz = kA * sz, asy = kA * sy;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if constexpr (!ADJOINT) {
        double gm[3], gp[3];
        ref_grad8_pair(hu[c], sy, sz, gm, gp);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // du/dX_j = sum_d K[j][d] du/dxi_d
          row[(qm * NV + c) * 3 + j] = fma(Km[j][2], gm[2], fma(Km[j][1], gm[1], Km[j][0] * gm[0]));
          row[(qp * NV + c) * 3 + j] = fma(Kp[j][2], gp[2], fma(Kp[j][1], gp[1], Kp[j][0] * gp[0]));
        }
      } else {
        double Qm[3], Qp[3];  // reference-space fluxes: Q[d] = sum_j g[j] K[j][d]
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Qm[d] = fma(row[(qm * NV + c) * 3 + 2], Km[2][d], fma(row[(qm * NV + c) * 3 + 1], Km[1][d], row[(qm * NV + c) * 3] * Km[0][d]));
          Qp[d] = fma(row[(qp * NV + c) * 3 + 2], Kp[2][d], fma(row[(qp * NV + c) * 3 + 1], Kp[1][d], row[(qp * NV + c) * 3] * Kp[0][d]));
        }
        // transpose of ref_grad8_pair (accumulate_pair, one component)
        const double S0 = Qp[0] + Qm[0], S1 = Qp[1] + Qm[1], S2 = Qp[2] + Qm[2];
        const double D1 = Qp[1] - Qm[1], D2 = Qp[2] - Qm[2];
        hu[c][0] += S0;
        hu[c][1] += S1;
        hu[c][2] += S2;
        hu[c][3] = fma(sy, S0, fma(kA, D1, hu[c][3]));
        hu[c][4] = fma(sz, S1, fma(sy, S2, hu[c][4]));
        hu[c][5] = fma(sz, S0, fma(kA, D2, hu[c][5]));
        hu[c][6] = fma(syz, S0, fma(asz, D1, fma(asy, D2, hu[c][6])));
      }
    }
  }
  if constexpr (!ADJOINT) {
    __syncwarp();
    if (e0 < E) warp_rows_store(out + e0 * CH, st, S, CH, (int)min((int64_t)32, E - e0));
  } else {
    __syncwarp();  // every lane is done with its staged g: the buffer becomes the scatter staging
    double Y[8][NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      double f[8];
      from_modal_raw(hu[c], f);
#pragma unroll
      for (int n = 0; n < 8; ++n) Y[n][c] = f[n];
    }
    grouped_scatter<8, NV>(out, nd, Y, valid, st);
  }
}

// ---------------------------------------------------------------------------------------------
// v4: the v3 arithmetic in a PERSISTENT kernel that hides the gather and scatter phases.
// ncu (profiles/r02_hvp_ncu_stalls.md): a v3 warp spends 29 % of its life in the gather / modal-transform prologue (41 %
// of that waiting on the dependent connectivity -> nodal-row round trips to L2 / HBM) and 9 % in the scatter epilogue
// (56 % of that draining its REDs before the CTA may exit); the FP64 pipe only saturates while BOTH warps of a scheduler
// are inside the Gauss-point loop.  Here every thread walks a grid-strided list of elements:
//   * the NEXT element's connectivity is loaded during the first pair iteration and its 24 nodal rows are prefetched
//     (PF = 1: into L2, PF = 2: into L1) during the second one, so the gather after the loop finds them on chip;
//   * the connectivity lines one more element ahead are prefetched into L2;
//   * the REDs of an element are fire-and-forget: the thread goes on with the next element, nothing drains until the end.
// ---------------------------------------------------------------------------------------------
template <int PF>
TATVA_D void prefetch_row(const double* p) {
  if constexpr (PF == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if constexpr (PF == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <int MINB, int PF, int DELAY = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_hvp_v4(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                     double lmbda, const double* __restrict__ u, const double* __restrict__ v,
                     double* __restrict__ y) {
  extern __shared__ double sm[];
  double* sX0 = sm + threadIdx.x;
  double* sv0 = sm + 21 * kBlock + threadIdx.x;
  int4* snd = reinterpret_cast<int4*>(sm + 42 * kBlock) + threadIdx.x;  // next connectivity: [2][kBlock] int4
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= E) return;
  const int4* c4 = reinterpret_cast<const int4*>(conn);
  if constexpr (DELAY > 0) {
    // All CTAs of a persistent grid start together and every element takes the same time, so the two warps of a
    // scheduler would sit in the same phase for the whole launch (both gathering, then both in the loop).  The second
    // half of the grid (the second CTA of every SM) starts DELAY cycles late and stays out of phase.
    if (blockIdx.x >= (gridDim.x >> 1)) {
      const long long t_start = clock64();
      while (clock64() - t_start < DELAY) {
      }
    }
  }
  int4 t0 = __ldg(c4 + 2 * e), t1 = __ldg(c4 + 2 * e + 1);
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);
#pragma unroll 1
  for (;;) {
    const int nd[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    double hx[3][7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double fX[8], fu[8], fv[8], tX[7], tv[7];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
        fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
        fv[n] = __ldg(v + (int64_t)nd[n] * 3 + c);
      }
      to_modal_raw(fX, tX);
      to_modal_raw(fu, hx[c]);
      to_modal_raw(fv, tv);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        hx[c][k] += tX[k];
        sX0[(c * 7 + k) * kBlock] = tX[k];
        sv0[(c * 7 + k) * kBlock] = tv[k];
      }
    }
    double R[3][7];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
    const int64_t en = e + stride;
    const bool more = en < E;
    int4 n0 = t0, n1 = t1;

#pragma unroll 1
    for (int pq = 0; pq < 4; ++pq) {
      const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
      int opaque = 0;
      asm volatile("" : "+r"(opaque));
      const double* sX = sX0 + opaque;
      const double* sv = sv0 + opaque;
      if (more) {
        if (pq == 0) {  // next element's connectivity (its lines were prefetched into L2 one element ago)
          n0 = __ldg(c4 + 2 * en);
          n1 = __ldg(c4 + 2 * en + 1);
          if (en + stride < E) asm volatile("prefetch.global.L2 [%0];" ::"l"(c4 + 2 * (en + stride)));
        } else if (pq == 1) {  // ... has arrived: park it in shared memory and prefetch its nodal rows
          snd[0] = n0;
          snd[kBlock] = n1;
          if constexpr (PF != 0) {
            const int nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              prefetch_row<PF>(coords + (int64_t)nn[n] * 3);
              prefetch_row<PF>(u + (int64_t)nn[n] * 3);
              prefetch_row<PF>(v + (int64_t)nn[n] * 3);
            }
          }
        }
      }
      double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3], Gvm[3][3], Gvp[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double gm[3], gp[3], t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sX[(c * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, gm, gp);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Jm[d][c] = gm[d];
          Jp[d][c] = gp[d];
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sv[(i * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, Gvm[i], Gvp[i]);
      }
      double Qm[3][3], Qp[3][3];
      point_flux(Jm, Frm, Gvm, mu_s, lm_s, Qm);
      point_flux(Jp, Frp, Gvp, mu_s, lm_s, Qp);
      accumulate_pair(Qm, Qp, sy, sz, syz, R);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double f[8];
      from_modal_raw(R[i], f);
#pragma unroll
      for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
    }
    if (!more) break;
    e = en;
    t0 = snd[0];
    t1 = snd[kBlock];
  }
}

// ---------------------------------------------------------------------------------------------
// v5: three warpgroups per SM that trade registers (setmaxnreg) so that two of them are always inside the Gauss-point
// loop while the third gathers / scatters.
//
// ncu of v3 (profiles/r02_hvp_ncu_stalls.md): with 244 registers only two warps fit per scheduler, and they fall into
// anti-phase — one runs the loop alone (it nearly saturates the FP64 pipe once the three-register-source DFMAs are
// counted at 3 cycles) while the other gathers; for ~19 % of the time NEITHER is in the loop and the pipe idles.  A third
// warp per scheduler needs <= 168 registers, which the loop cannot do without spilling (variant 26).  But only the loop
// needs many registers: the gather / modal transform / scatter phases need < 80.  So one persistent 384-thread CTA per
// SM holds three warpgroups which each cycle through
//     thin (THIN regs): scatter the previous tile, gather + transform the next one into shared memory
//     setmaxnreg.inc FAT  (blocks until another warpgroup has released its registers)
//     fat  (FAT regs):  the four tx-pair iterations, modal coefficients read from shared memory
//     setmaxnreg.dec THIN
// with 2 FAT + THIN = 3 x 168 = the CTA's register pool.  The gather is cp.async (global -> shared, no registers, ONE
// round trip for the 72 nodal values); the modal transform works in place on the thread's own shared-memory column,
// so no thread ever reads another thread's data and the kernel needs no barrier at all.
// ---------------------------------------------------------------------------------------------
constexpr int kV5Threads = 384;
constexpr int kV5Slots = 72;  // doubles per thread: 9 field components x 8 (7 modal coefficients + 1 spare)

TATVA_D void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

template <int FAT, int THIN, int XREG = 0>
__global__ void __launch_bounds__(kV5Threads, 1)
    k_hex8_nh_hvp_v5(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                     double lmbda, const double* __restrict__ u, const double* __restrict__ v,
                     double* __restrict__ y) {
  static_assert(2 * FAT + THIN == 3 * 168, "the CTA's register pool is 3 warpgroups x 168");
  extern __shared__ double sm[];
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(THIN));
  double* col = sm + threadIdx.x;  // slot (f, c, k) = col[((f * 3 + c) * 8 + k) * kV5Threads]
  const int wg = threadIdx.x >> 7;
  const int64_t n_tiles = (E + 127) >> 7;
  const int64_t tile_stride = (int64_t)gridDim.x * 3;
  const int4* c4 = reinterpret_cast<const int4*>(conn);
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);
  int* snd = reinterpret_cast<int*>(col + (size_t)(3 * 8 + 7) * kV5Threads);  // spare slot of (u, c = 0): node ids 0, 1
  // node ids are parked in the spare (k = 7) slots of the u field: ints (2n, 2n+1) in slot (1, n >> 1 ... ) -- see nd_slot
  auto nd_slot = [&](int n) -> int* { return reinterpret_cast<int*>(col + (size_t)((3 + (n >> 1)) * 8 + 7) * kV5Threads) + (n & 1); };
  (void)snd;
  bool have_prev = false, prev_valid = false;
#pragma unroll 1
  for (int64_t tile = (int64_t)blockIdx.x * 3 + wg; tile < n_tiles + tile_stride; tile += tile_stride) {
    // ---------------- thin: scatter the previous tile (its modal residuals sit in the X slots) ----------------
    if (have_prev) {
      int nd[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) nd[n] = *nd_slot(n);
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        double r[7], f[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) r[k] = col[(size_t)(i * 8 + k) * kV5Threads];
        from_modal_raw(r, f);
        if (prev_valid) {
#pragma unroll
          for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
        }
      }
    }
    if (tile >= n_tiles) break;  // uniform over the warpgroup
    // ---------------- thin: gather the next tile (cp.async, one round trip) and transform it in place ----------------
    const int64_t e0 = tile * 128 + (threadIdx.x & 127);
    const bool valid = e0 < E;
    const int64_t e = valid ? e0 : E - 1;
    {
      const int4 t0 = __ldg(c4 + 2 * e), t1 = __ldg(c4 + 2 * e + 1);
      const int nd[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int64_t o = (int64_t)nd[n] * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          cp_async8(col + (size_t)((0 + c) * 8 + n) * kV5Threads, coords + o + c);
          cp_async8(col + (size_t)((3 + c) * 8 + n) * kV5Threads, u + o + c);
          cp_async8(col + (size_t)((6 + c) * 8 + n) * kV5Threads, v + o + c);
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        double f[8], tX[7], t[7];
        double* cX = col + (size_t)((0 + c) * 8) * kV5Threads;
        double* cu = col + (size_t)((3 + c) * 8) * kV5Threads;
        double* cv = col + (size_t)((6 + c) * 8) * kV5Threads;
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = cX[(size_t)n * kV5Threads];
        to_modal_raw(f, tX);
#pragma unroll
        for (int k = 0; k < 7; ++k) cX[(size_t)k * kV5Threads] = tX[k];
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = cu[(size_t)n * kV5Threads];
        to_modal_raw(f, t);
#pragma unroll
        for (int k = 0; k < 7; ++k) cu[(size_t)k * kV5Threads] = t[k] + tX[k];
#pragma unroll
        for (int n = 0; n < 8; ++n) f[n] = cv[(size_t)n * kV5Threads];
        to_modal_raw(f, t);
#pragma unroll
        for (int k = 0; k < 7; ++k) cv[(size_t)k * kV5Threads] = t[k];
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) *nd_slot(n) = nd[n];
    }
    // ---------------- fat: the Gauss-point loop ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FAT));
    {
      double R[3][7];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
      double hx[XREG ? 3 : 1][7];  // XREG: the modal deformed coordinates stay in registers for the four iterations
      if constexpr (XREG) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 7; ++k) hx[i][k] = col[(size_t)((3 + i) * 8 + k) * kV5Threads];
      }
#pragma unroll 1
      for (int pq = 0; pq < 4; ++pq) {
        const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
        int opaque = 0;
        asm volatile("" : "+r"(opaque));
        const double* cc = col + opaque;
        double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3], Gvm[3][3], Gvp[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double gm[3], gp[3], t[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) t[k] = cc[(size_t)((0 + c) * 8 + k) * kV5Threads];
          ref_grad8_pair(t, sy, sz, gm, gp);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            Jm[d][c] = gm[d];
            Jp[d][c] = gp[d];
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double t[7];
          if constexpr (XREG) {
            ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
          } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) t[k] = cc[(size_t)((3 + i) * 8 + k) * kV5Threads];
            ref_grad8_pair(t, sy, sz, Frm[i], Frp[i]);
          }
#pragma unroll
          for (int k = 0; k < 7; ++k) t[k] = cc[(size_t)((6 + i) * 8 + k) * kV5Threads];
          ref_grad8_pair(t, sy, sz, Gvm[i], Gvp[i]);
        }
        double Qm[3][3], Qp[3][3];
        point_flux(Jm, Frm, Gvm, mu_s, lm_s, Qm);
        point_flux(Jp, Frp, Gvp, mu_s, lm_s, Qp);
        accumulate_pair(Qm, Qp, sy, sz, syz, R);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 7; ++k) col[(size_t)(i * 8 + k) * kV5Threads] = R[i][k];
    }
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(THIN));
    have_prev = true;
    prev_valid = valid;
  }
}

// ---------------------------------------------------------------------------------------------
// Residual in the same modal / reference-space form:
//   W P K = W [ mu Fr M + (lambda lnJ - mu) A^T ],   P = mu (F - F^-T) + lambda lnJ F^-T,  F = Fr K^T.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 2)
    k_hex8_nh_residual_modal(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                             double lmbda, const double* __restrict__ u, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  double rX[3][7], rx[3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
      fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
    }
    to_modal(fX, rX[c]);
    to_modal(fu, rx[c]);
#pragma unroll
    for (int k = 0; k < 7; ++k) rx[c][k] += rX[c][k];
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;

#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const double tx = kSigns[q][0], ty = kSigns[q][1], tz = kSigns[q][2];
    double J[3][3], Kc[3][3], detJ;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double g[3];
      ref_grad_s((const double*)rX[c], 1, tx, ty, tz, g);
      J[0][c] = g[0]; J[1][c] = g[1]; J[2][c] = g[2];
    }
    adjugate(J, Kc, detJ);
    double M[3][3];  // detJ^2 K^T K
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) {
        M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
        M[b][a] = M[a][b];
      }
    double Fr[3][3], Ac[3][3], detF;
#pragma unroll
    for (int i = 0; i < 3; ++i) ref_grad_s((const double*)rx[i], 1, tx, ty, tz, Fr[i]);
    adjugate(Fr, Ac, detF);
    const double rJ = 1.0 / detJ, rF = 1.0 / detF;
    const double lnJ = log(detF * rJ);
    const double w1 = mu * rJ, w2 = (lmbda * lnJ - mu) * detJ * rF;
    const double tyz = kSigns[q][3], txz = kSigns[q][4], txy = kSigns[q][5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double qv[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double t1 = Fr[i][0] * M[0][d] + Fr[i][1] * M[1][d] + Fr[i][2] * M[2][d];
        qv[d] = w1 * t1 + w2 * Ac[d][i];
      }
      R[i][0] += qv[0];
      R[i][1] += qv[1];
      R[i][2] += qv[2];
      R[i][3] = fma(ty, qv[0], fma(tx, qv[1], R[i][3]));
      R[i][4] = fma(tz, qv[1], fma(ty, qv[2], R[i][4]));
      R[i][5] = fma(tz, qv[0], fma(tx, qv[2], R[i][5]));
      R[i][6] = fma(tyz, qv[0], fma(txz, qv[1], fma(txy, qv[2], R[i][6])));
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}


// ---------------------------------------------------------------------------------------------
// Energy in the same modal / reference-space form: psi W with I1 = |F|^2 = tr(Fr M Fr^T), M = K^T K,
// J = det Fr / det J.  Per-CTA partial sums, finished by k_sum_rows_final (fixed order).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 3)
    k_hex8_nh_energy_modal(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                           double lmbda, const double* __restrict__ u, double* __restrict__ partials) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double energy = 0.0;
  if (e < E) {
    int nd[8];
    {
      const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
      const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
      nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
      nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
    }
    double rX[3][7], rx[3][7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double fX[8], fu[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
        fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
      }
      to_modal(fX, rX[c]);
      to_modal(fu, rx[c]);
#pragma unroll
      for (int k = 0; k < 7; ++k) rx[c][k] += rX[c][k];
    }
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      const double tx = kSigns[q][0], ty = kSigns[q][1], tz = kSigns[q][2];
      double J[3][3], Kc[3][3], detJ;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double g[3];
        ref_grad_s((const double*)rX[c], 1, tx, ty, tz, g);
        J[0][c] = g[0]; J[1][c] = g[1]; J[2][c] = g[2];
      }
      adjugate(J, Kc, detJ);
      double M[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
          M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
          M[b][a] = M[a][b];
        }
      double Fr[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ref_grad_s((const double*)rx[i], 1, tx, ty, tz, Fr[i]);
      const double detF = Fr[0][0] * (Fr[1][1] * Fr[2][2] - Fr[1][2] * Fr[2][1]) +
                          Fr[0][1] * (Fr[1][2] * Fr[2][0] - Fr[1][0] * Fr[2][2]) +
                          Fr[0][2] * (Fr[1][0] * Fr[2][1] - Fr[1][1] * Fr[2][0]);
      double I1 = 0.0;  // detJ^2 |F|^2
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d) I1 = fma(Fr[i][0] * M[0][d] + Fr[i][1] * M[1][d] + Fr[i][2] * M[2][d], Fr[i][d], I1);
      const double rJ = 1.0 / detJ;
      const double lnJ = log(detF * rJ);
      energy += detJ * (0.5 * mu * (I1 * rJ * rJ - 3.0 - 2.0 * lnJ) + 0.5 * lmbda * lnJ * lnJ);
    }
  }
  // block reduction (same scheme as the generic energy kernel)
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) energy += __shfl_down_sync(0xffffffffu, energy, o);
  if (lane == 0) sh[w] = energy;
  __syncthreads();
  if (w == 0) {
    energy = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_down_sync(0xffffffffu, energy, o);
    if (lane == 0) partials[blockIdx.x] = energy;
  }
}


// ---------------------------------------------------------------------------------------------
// Tet4 x neo-Hookean in the same reference-space form (one point, constant dN/dxi):
//   J = [X1-X0, X2-X0, X3-X0]^T,  Fr = J + [u1-u0, ...],  Gv = [v1-v0, ...],  W = det J / 6,
//   nodal forces f_1..3 = columns of Q = W dP K (resp. W P K), f_0 = -(f_1 + f_2 + f_3).
// About 290 FP64 instructions per element instead of ~450 for the generic template.
// ---------------------------------------------------------------------------------------------
TATVA_HD void point_flux_residual(const double (&J)[3][3], const double (&Fr)[3][3], double mu_s, double lm_s,
                                 double (&Q)[3][3]) {
  double Kc[3][3], detJ, Ac[3][3], detF;
  adjugate(J, Kc, detJ);
  double M[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
      M[b][a] = M[a][b];
    }
  adjugate(Fr, Ac, detF);
  const double r = fast_rcp(detJ * detF);
  const double rJ = r * detF, rF = r * detJ;
  const double lnJ = log_pos(detF * rJ);
  const double w1 = mu_s * rJ, w2 = (lm_s * lnJ - mu_s) * detJ * rF;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      Q[i][d] = fma(w1, Fr[i][0] * M[0][d] + Fr[i][1] * M[1][d] + Fr[i][2] * M[2][d], w2 * Ac[d][i]);
}

// Energy in the same raw-modal / tx-pair form.  Gradients are carried at 8x their value: I1 and lnJ are ratios and do
// not notice, the weight det J picks up 8^3, removed once at the end.
TATVA_HD double point_energy(const double (&J)[3][3], const double (&Fr)[3][3], double mu, double lmbda) {
  double Kc[3][3], detJ;
  adjugate(J, Kc, detJ);
  double M[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      M[a][b] = Kc[0][a] * Kc[0][b] + Kc[1][a] * Kc[1][b] + Kc[2][a] * Kc[2][b];
      M[b][a] = M[a][b];
    }
  const double detF = Fr[0][0] * (Fr[1][1] * Fr[2][2] - Fr[1][2] * Fr[2][1]) +
                      Fr[0][1] * (Fr[1][2] * Fr[2][0] - Fr[1][0] * Fr[2][2]) +
                      Fr[0][2] * (Fr[1][0] * Fr[2][1] - Fr[1][1] * Fr[2][0]);
  double I1 = 0.0;  // detJ^2 |F|^2
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) I1 = fma(Fr[i][0] * M[0][d] + Fr[i][1] * M[1][d] + Fr[i][2] * M[2][d], Fr[i][d], I1);
  const double rJ = fast_rcp(detJ);
  const double lnJ = log_pos(detF * rJ);
  return detJ * (0.5 * mu * (I1 * rJ * rJ - 3.0 - 2.0 * lnJ) + 0.5 * lmbda * lnJ * lnJ);
}

template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_energy_v3(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                        double lmbda, const double* __restrict__ u, double* __restrict__ partials) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double energy = 0.0;
  if (e < E) {
    int nd[8];
    {
      const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
      const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
      nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
      nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
    }
    double hX[3][7], hx[3][7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double fX[8], fu[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
        fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
      }
      to_modal_raw(fX, hX[c]);
      to_modal_raw(fu, hx[c]);
#pragma unroll
      for (int k = 0; k < 7; ++k) hx[c][k] += hX[c][k];
    }
#pragma unroll 1
    for (int pq = 0; pq < 4; ++pq) {
      const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1];
      double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double gm[3], gp[3];
        ref_grad8_pair(hX[c], sy, sz, gm, gp);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Jm[d][c] = gm[d];
          Jp[d][c] = gp[d];
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
      energy += point_energy(Jm, Frm, mu, lmbda) + point_energy(Jp, Frp, mu, lmbda);
    }
    energy *= 1.0 / 512.0;
  }
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) energy += __shfl_down_sync(0xffffffffu, energy, o);
  if (lane == 0) sh[w] = energy;
  __syncthreads();
  if (w == 0) {
    energy = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_down_sync(0xffffffffu, energy, o);
    if (lane == 0) partials[blockIdx.x] = energy;
  }
}

// Hex8 residual with the structure of the v3 HVP kernel: raw (unnormalised) modal coefficients, the two Gauss points
// of a tx pair evaluated together from shared partial sums, signs from the constant table, scalings folded into
// mu/512 and lambda/512, and (STAGE) the modal coordinates parked in shared memory between iterations.
template <int MINB, int STAGE>
__global__ void __launch_bounds__(kBlock, MINB)
    k_hex8_nh_residual_v3(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t E, double mu,
                          double lmbda, const double* __restrict__ u, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[8];
  {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  }
  extern __shared__ double sm[];
  double* sX0 = sm + threadIdx.x;
  double hX[STAGE ? 1 : 3][7], hx[3][7];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], tX[7];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      fX[n] = __ldg(coords + (int64_t)nd[n] * 3 + c);
      fu[n] = __ldg(u + (int64_t)nd[n] * 3 + c);
    }
    to_modal_raw(fX, tX);
    to_modal_raw(fu, hx[c]);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      hx[c][k] += tX[k];
      if constexpr (STAGE) sX0[(c * 7 + k) * kBlock] = tX[k];
      else hX[c][k] = tX[k];
    }
  }
  double R[3][7];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) R[i][k] = 0.0;
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);

#pragma unroll 1
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSigns[pq][0], sz = kPairSigns[pq][1], syz = kPairSigns[pq][2];
    int opaque = 0;
    asm volatile("" : "+r"(opaque));
    const double* sX = sX0 + opaque;
    double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      if constexpr (STAGE) {
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = sX[(c * 7 + k) * kBlock];
        ref_grad8_pair(t, sy, sz, gm, gp);
      } else {
        ref_grad8_pair(hX[c], sy, sz, gm, gp);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
    double Qm[3][3], Qp[3][3];
    point_flux_residual(Jm, Frm, mu_s, lm_s, Qm);
    point_flux_residual(Jp, Frp, mu_s, lm_s, Qp);
    accumulate_pair(Qm, Qp, sy, sz, syz, R);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // y may still be being cleared (launch_behind_zero)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal_raw(R[i], f);
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(y + (int64_t)nd[n] * 3 + i, f[n]);
  }
}

template <bool HVP>
__global__ void __launch_bounds__(kBlock) k_tet4_nh_ref(const double* __restrict__ coords,
                                                        const int32_t* __restrict__ conn, int64_t E, double mu,
                                                        double lmbda, const double* __restrict__ u,
                                                        const double* __restrict__ v, double* __restrict__ y) {
  extern __shared__ double sm[];
  const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e0 < E;
  const int64_t e = valid ? e0 : E - 1;
  int nd[4];
  {
    const int4 t = __ldg(reinterpret_cast<const int4*>(conn) + e);
    nd[0] = t.x; nd[1] = t.y; nd[2] = t.z; nd[3] = t.w;
  }
  double X[4][3], U[4][3], V[4][3];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    load_row<3>(coords, nd[n], X[n]);
    load_row<3>(u, nd[n], U[n]);
    if constexpr (HVP) load_row<3>(v, nd[n], V[n]);
  }
  double J[3][3], Fr[3][3], Gv[3][3], Q[3][3];  // J[d][c] = dX_c/dxi_d ; Fr[i][d] = dx_i/dxi_d
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      J[d][c] = X[d + 1][c] - X[0][c];
      Fr[c][d] = J[d][c] + (U[d + 1][c] - U[0][c]);
      if constexpr (HVP) Gv[c][d] = V[d + 1][c] - V[0][c];
    }
  if constexpr (HVP) point_flux(J, Fr, Gv, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  else point_flux_residual(J, Fr, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  double Y[4][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Y[1][i] = Q[i][0];
    Y[2][i] = Q[i][1];
    Y[3][i] = Q[i][2];
    Y[0][i] = -(Q[i][0] + Q[i][1] + Q[i][2]);
  }
  grouped_scatter<4, 3>(y, nd, Y, valid, sm + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<4, 3>());
}


// ---------------------------------------------------------------------------------------------
// Persistent variant with the connectivity prefetched one element ahead.  The one-point Tet4 kernels are latency
// bound (two dependent trips to memory per element: connectivity, then the nodal rows; ~300 FP64 instructions in
// between do not cover them).  Here every warp walks a strided list of 32-element groups and loads the NEXT group's
// connectivity before gathering the current one, so an element costs one trip instead of two.
// ---------------------------------------------------------------------------------------------
template <bool HVP>
__global__ void __launch_bounds__(kBlock) k_tet4_nh_pipe(const double* __restrict__ coords,
                                                         const int32_t* __restrict__ conn, int64_t E, double mu,
                                                         double lmbda, const double* __restrict__ u,
                                                         const double* __restrict__ v, double* __restrict__ y) {
  extern __shared__ double sm[];
  double* wsm = sm + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<4, 3>();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // elements per sweep, a multiple of 32
  int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  if (base >= E) return;  // whole warp
  const int4* c4 = reinterpret_cast<const int4*>(conn);
  const double mu_s = mu * (1.0 / 6.0), lm_s = lmbda * (1.0 / 6.0);
  int4 next = __ldg(c4 + ((base + lane < E) ? base + lane : E - 1));
#pragma unroll 1
  for (; base < E; base += stride) {
    const bool valid = base + lane < E;
    const int4 t = next;
    const int64_t nb = base + stride;
    if (nb < E) next = __ldg(c4 + ((nb + lane < E) ? nb + lane : E - 1));
    int nd[4] = {t.x, t.y, t.z, t.w};
    double X[4][3], U[4][3], V[4][3];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      load_row<3>(coords, nd[n], X[n]);
      load_row<3>(u, nd[n], U[n]);
      if constexpr (HVP) load_row<3>(v, nd[n], V[n]);
    }
    double J[3][3], Fr[3][3], Gv[3][3], Q[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J[d][c] = X[d + 1][c] - X[0][c];
        Fr[c][d] = J[d][c] + (U[d + 1][c] - U[0][c]);
        if constexpr (HVP) Gv[c][d] = V[d + 1][c] - V[0][c];
      }
    if constexpr (HVP) point_flux(J, Fr, Gv, mu_s, lm_s, Q);
    else point_flux_residual(J, Fr, mu_s, lm_s, Q);
    double Y[4][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Y[1][i] = Q[i][0];
      Y[2][i] = Q[i][1];
      Y[3][i] = Q[i][2];
      Y[0][i] = -(Q[i][0] + Q[i][1] + Q[i][2]);
    }
    grouped_scatter<4, 3>(y, nd, Y, valid, wsm);
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled variant: one CTA = one tile of kBlock consecutive elements.  The tile's unique nodes (plan-time table)
// are gathered ONCE, coalesced, into shared memory (coordinates, u, v: 9 doubles per node); elements then read
// their nodal rows with LDS through tile-local uint16 connectivity.  On the 6-tets-per-cell box a tile touches
// ~90 unique nodes for 512 node references, so the L1 sector traffic of the gather drops ~2.5x — the generic
// kernel is L1-bound there (ncu: l1tex throughput 78 %, 12 sectors per LDG.64 request).
// ---------------------------------------------------------------------------------------------
template <bool HVP>
__global__ void __launch_bounds__(kBlock) k_tet4_nh_tiled(const double* __restrict__ coords,
                                                          const int32_t* __restrict__ tile_ptr,
                                                          const int32_t* __restrict__ tile_nodes,
                                                          const uint16_t* __restrict__ tile_conn, int64_t E,
                                                          int max_unique, double mu, double lmbda,
                                                          const double* __restrict__ u, const double* __restrict__ v,
                                                          double* __restrict__ y) {
  extern __shared__ double sm[];
  const int n0 = __ldg(tile_ptr + blockIdx.x), nu = __ldg(tile_ptr + blockIdx.x + 1) - n0;
  double* sX = sm;
  double* sU = sm + 3 * max_unique;
  double* sV = sm + 6 * max_unique;
  int* sNode = reinterpret_cast<int*>(sm + (HVP ? 9 : 6) * max_unique);
  for (int t = threadIdx.x; t < nu; t += blockDim.x) sNode[t] = __ldg(tile_nodes + n0 + t);
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * nu; t += blockDim.x) {
    const int i = t / 3, c = t - 3 * i;
    const int64_t g = (int64_t)sNode[i] * 3 + c;
    sX[t] = __ldg(coords + g);
    sU[t] = __ldg(u + g);
    if constexpr (HVP) sV[t] = __ldg(v + g);
  }
  __syncthreads();
  const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e0 < E;
  const int64_t e = valid ? e0 : E - 1;
  int ln[4], nd[4];
  {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(tile_conn) + e);  // 4 x uint16
    ln[0] = t.x & 0xffff; ln[1] = t.x >> 16; ln[2] = t.y & 0xffff; ln[3] = t.y >> 16;
  }
  double X[4][3], U[4][3], V[4][3];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    nd[n] = sNode[ln[n]];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      X[n][c] = sX[ln[n] * 3 + c];
      U[n][c] = sU[ln[n] * 3 + c];
      if constexpr (HVP) V[n][c] = sV[ln[n] * 3 + c];
    }
  }
  double J[3][3], Fr[3][3], Gv[3][3], Q[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      J[d][c] = X[d + 1][c] - X[0][c];
      Fr[c][d] = J[d][c] + (U[d + 1][c] - U[0][c]);
      if constexpr (HVP) Gv[c][d] = V[d + 1][c] - V[0][c];
    }
  if constexpr (HVP) point_flux(J, Fr, Gv, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  else point_flux_residual(J, Fr, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  double Y[4][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Y[1][i] = Q[i][0];
    Y[2][i] = Q[i][1];
    Y[3][i] = Q[i][2];
    Y[0][i] = -(Q[i][0] + Q[i][1] + Q[i][2]);
  }
  // tile-local accumulation is not needed for correctness: the sector-grouped REDs go straight to y
  __syncthreads();  // everybody is done with the staged inputs; reuse the buffer for the grouped scatter
  grouped_scatter<4, 3>(y, nd, Y, valid, sm + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<4, 3>());
}

}  // namespace

template <int MINB, int STAGE, int GROUPED = 0, int WIDE = 0, int UNR = 1, int CPF = 0, int NDS = 0>
static int launch_v3(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                     cudaStream_t st, bool pdl = false) {
  static_assert(!(NDS && GROUPED), "the grouped scatter has its own staging area");
  constexpr size_t smem = ((size_t)(STAGE == 0 ? 0 : (STAGE == 1 ? 42 : 63)) * kBlock + (GROUPED ? (kBlock / 32) * grouped_scatter_words<8, 3>() : 0)) * sizeof(double) + (NDS ? 2 * kBlock * sizeof(int4) : 0);
  static SmemOptIn configured;
  if (smem > 48 * 1024) {
    const int rc = opt_in_smem(k_hex8_nh_hvp_v3<MINB, STAGE, GROUPED, false, WIDE, UNR, CPF, NDS>, smem, configured);
    if (rc != TATVA_OK) return rc;
  }
  if (pdl)  // clear y with our own kernel (or rely on the caller's, p->pss) and let the element kernel start behind it without waiting for it to finish
    return launch_behind_zero(k_hex8_nh_hvp_v3<MINB, STAGE, GROUPED, false, WIDE, UNR, CPF, NDS>, grid_for(p->n_elems), kBlock, smem, st, p->zero_output ? 1 : 2, y, p->n_nodes * 3,
                              p->coords, p->conn, p->n_elems, mu * (1.0 / 512.0), lmbda * (1.0 / 512.0), u, v, y, nullptr, nullptr);
  k_hex8_nh_hvp_v3<MINB, STAGE, GROUPED, false, WIDE, UNR, CPF, NDS><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu * (1.0 / 512.0), lmbda * (1.0 / 512.0), u, v, y);
  return TATVA_OK;
}

// persistent grid: as many CTAs as fit on the device at once (queried once per device)
template <class K>
static int resident_grid(K kernel, size_t smem, int64_t n_elems, int (&cache)[64], int* grid) {
  int dev = 0;
  TATVA_CUDA_TRY(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  int g = tracked ? cache[dev] : 0;
  if (g == 0) {
    int sms = 0, per_sm = 0;
    TATVA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TATVA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem));
    g = sms * (per_sm > 0 ? per_sm : 1);
    if (tracked) cache[dev] = g;
  }
  const int need = grid_for(n_elems);
  *grid = g < need ? g : need;
  return TATVA_OK;
}

template <int MINB, int PF, int DELAY = 0>
static int launch_v4(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                     cudaStream_t st) {
  constexpr size_t smem = (size_t)42 * kBlock * sizeof(double) + (size_t)2 * kBlock * sizeof(int4);
  static SmemOptIn configured;
  static int cache[64];
  int rc = opt_in_smem(k_hex8_nh_hvp_v4<MINB, PF, DELAY>, smem, configured);
  if (rc != TATVA_OK) return rc;
  int grid = 0;
  if ((rc = resident_grid(k_hex8_nh_hvp_v4<MINB, PF, DELAY>, smem, p->n_elems, cache, &grid)) != TATVA_OK) return rc;
  k_hex8_nh_hvp_v4<MINB, PF, DELAY><<<grid, kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y);
  return TATVA_OK;
}

template <int FAT, int THIN, int XREG = 0>
static int launch_v5(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                     cudaStream_t st) {
  constexpr size_t smem = (size_t)kV5Slots * kV5Threads * sizeof(double);  // 216 KB: one CTA per SM
  static SmemOptIn configured;
  int rc = opt_in_smem(k_hex8_nh_hvp_v5<FAT, THIN, XREG>, smem, configured);
  if (rc != TATVA_OK) return rc;
  int dev = 0, sms = 0;
  TATVA_CUDA_TRY(cudaGetDevice(&dev));
  TATVA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t tiles = (p->n_elems + 127) / 128;
  const int grid = (int)((tiles + 2) / 3 < sms ? (tiles + 2) / 3 : sms);
  k_hex8_nh_hvp_v5<FAT, THIN, XREG><<<grid, kV5Threads, smem, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y);
  return TATVA_OK;
}

template <int MINB, int STAGE, int PF = 0>
static int launch_geo(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y, cudaStream_t st) {
  constexpr size_t smem = (size_t)(STAGE == 0 ? 0 : (STAGE == 1 ? 21 : 42)) * kBlock * sizeof(double);
  static_assert(smem <= 48 * 1024, "fits the default shared-memory window");
  k_hex8_nh_hvp_geo<MINB, STAGE, PF><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu * (1.0 / 512.0), lmbda * (1.0 / 512.0), u, v, y,
                                                                             reinterpret_cast<const double2*>(p->geo), p->geo_stride);
  return TATVA_OK;
}

template <int NV, bool ADJOINT>
static int launch_grad_modal(const tatva_plan* p, const double* in, double* out, cudaStream_t st) {
  constexpr int CH = 8 * NV * 3, S = CH | 1;
  constexpr size_t stage = (size_t)32 * S, scat = (size_t)grouped_scatter_words<8, NV>();
  constexpr size_t smem = (size_t)(kBlock / 32) * (ADJOINT && scat > stage ? scat : stage) * sizeof(double);
  static SmemOptIn configured;
  if (smem > 48 * 1024) {
    const int rc = opt_in_smem(k_hex8_grad_modal<NV, ADJOINT>, smem, configured);
    if (rc != TATVA_OK) return rc;
  }
  k_hex8_grad_modal<NV, ADJOINT><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, in, out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int hex8_weights_modal(const tatva_plan* p, double* out, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(out) & 15) return TATVA_E_UNSUPPORTED;  // the caller falls back to the generic kernel
  k_hex8_weights_modal<<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// Operator.grad (adjoint = false: u (N, nv) -> (E, 8, nv, 3)) and its adjoint (g -> y (N, nv), accumulated into y) for
// nv <= 3; TATVA_E_UNSUPPORTED otherwise (the caller falls back to the generic building block).
int hex8_grad_modal(const tatva_plan* p, bool adjoint, const double* in, int nv, double* out, cudaStream_t st) {
  switch (nv) {
    case 1: return adjoint ? launch_grad_modal<1, true>(p, in, out, st) : launch_grad_modal<1, false>(p, in, out, st);
    case 2: return adjoint ? launch_grad_modal<2, true>(p, in, out, st) : launch_grad_modal<2, false>(p, in, out, st);
    case 3: return adjoint ? launch_grad_modal<3, true>(p, in, out, st) : launch_grad_modal<3, false>(p, in, out, st);
    default: return TATVA_E_UNSUPPORTED;
  }
}

int hex8_geometry_cache(const tatva_plan* p, double* geo, int64_t stride, cudaStream_t st) {
  k_hex8_geometry<<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, reinterpret_cast<double2*>(geo), stride);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int hex8_nh_hvp_modal(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                      cudaStream_t st) {
  // default (variant 0) with a zeroed output: the clearing of y overlaps the element kernel's gather and Gauss-point loop
  // (programmatic dependent launch); variant 26 is the same kernel behind a plain cudaMemsetAsync
  if ((p->zero_output || p->pss) && p->variant == 0 && !p->geo) {
    const int rc = launch_v3<3, 2>(p, mu, lmbda, u, v, y, st, true);
    if (rc != TATVA_OK) return rc;
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * p->n_nodes * 3, st));
  int rc = TATVA_OK;
  if (p->geo && (p->variant == 0 || (p->variant >= 50 && p->variant <= 58))) {
    switch (p->variant) {  // occupancy / staging points of the cached-geometry kernel
      case 50: rc = launch_geo<2, 0>(p, mu, lmbda, u, v, y, st); break;
      case 51: rc = launch_geo<2, 1>(p, mu, lmbda, u, v, y, st); break;
      case 52: rc = launch_geo<3, 1>(p, mu, lmbda, u, v, y, st); break;
      case 53: rc = launch_geo<3, 2>(p, mu, lmbda, u, v, y, st); break;
      case 54: rc = launch_geo<4, 2>(p, mu, lmbda, u, v, y, st); break;
      case 55: rc = launch_geo<3, 0>(p, mu, lmbda, u, v, y, st); break;
      case 56: rc = launch_geo<4, 1>(p, mu, lmbda, u, v, y, st); break;
      case 57: rc = launch_geo<3, 2, 1>(p, mu, lmbda, u, v, y, st); break;  // streaming loads
      case 58: rc = launch_geo<3, 2, 2>(p, mu, lmbda, u, v, y, st); break;  // + L2 prefetch of the thread's rows
      default: rc = launch_geo<3, 2>(p, mu, lmbda, u, v, y, st); break;
    }
    if (rc != TATVA_OK) return rc;
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  // measurement variants documented in DESIGN.md §3.1 (tatva_plan_set_variant); default = v3 pair kernel
  switch (p->variant) {
    case 2: k_hex8_nh_hvp<2><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y); break;
    case 3: rc = launch_rolled<0, 2>(p, mu, lmbda, u, v, y, st); break;
    case 8: rc = launch_rolled<0, 2, 1>(p, mu, lmbda, u, v, y, st); break;
    case 9: rc = launch_rolled<0, 2, 2>(p, mu, lmbda, u, v, y, st); break;
    case 15: rc = launch_rolled<0, 2, 0, 1, 1>(p, mu, lmbda, u, v, y, st); break;
    case 17: rc = launch_rolled<1, 3, 0, 1, 1>(p, mu, lmbda, u, v, y, st); break;
    case 22: rc = launch_v3<2, 0>(p, mu, lmbda, u, v, y, st); break;
    case 23: rc = launch_v3<2, 1>(p, mu, lmbda, u, v, y, st); break;
    case 28: rc = launch_v3<2, 1, 0, 1>(p, mu, lmbda, u, v, y, st); break;  // whole-row 16-byte gather (slower)
    case 27: rc = launch_v3<2, 1, 1>(p, mu, lmbda, u, v, y, st); break;
    case 25: rc = launch_v3<2, 2>(p, mu, lmbda, u, v, y, st); break;
    case 26: rc = launch_v3<3, 2>(p, mu, lmbda, u, v, y, st); break;
    case 31: rc = launch_v4<2, 0>(p, mu, lmbda, u, v, y, st); break;  // persistent, no prefetch
    case 32: rc = launch_v4<2, 1>(p, mu, lmbda, u, v, y, st); break;  // persistent + L2 prefetch of the next rows
    case 33: rc = launch_v4<2, 2>(p, mu, lmbda, u, v, y, st); break;  // persistent + L1 prefetch
    case 34: rc = launch_v4<2, 0, 2000>(p, mu, lmbda, u, v, y, st); break;  // persistent, second CTA of each SM out of phase
    case 35: rc = launch_v4<2, 0, 3500>(p, mu, lmbda, u, v, y, st); break;
    case 36: rc = launch_v4<2, 0, 5000>(p, mu, lmbda, u, v, y, st); break;
    case 37: rc = launch_v3<2, 1, 0, 0, 2>(p, mu, lmbda, u, v, y, st); break;  // two tx pairs per loop body
    case 43: rc = launch_v3<3, 2, 0, 0, 1, 0, 1>(p, mu, lmbda, u, v, y, st); break;  // 3 CTAs / SM, all staged, node ids parked
    case 44: rc = launch_v3<2, 1, 0, 0, 1, 0, 1>(p, mu, lmbda, u, v, y, st); break;  // 2 CTAs / SM, X and v staged, node ids parked
    case 45: rc = launch_v3<3, 1>(p, mu, lmbda, u, v, y, st); break;  // 3 CTAs / SM, X and v staged
    case 46: rc = launch_v3<4, 2>(p, mu, lmbda, u, v, y, st); break;  // 4 CTAs / SM (128 registers)
    case 38: rc = launch_v3<2, 1, 0, 0, 1, 296>(p, mu, lmbda, u, v, y, st); break;  // + connectivity prefetch one wave ahead
    case 39: rc = launch_v3<2, 1, 0, 0, 1, 592>(p, mu, lmbda, u, v, y, st); break;  // two waves ahead
    case 40: rc = launch_v5<216, 72>(p, mu, lmbda, u, v, y, st); break;  // rotating warpgroups (setmaxnreg)
    case 41: rc = launch_v5<208, 88>(p, mu, lmbda, u, v, y, st); break;
    case 42: rc = launch_v5<200, 104>(p, mu, lmbda, u, v, y, st); break;
    case 47: rc = launch_v5<232, 40, 1>(p, mu, lmbda, u, v, y, st); break;  // fat phase keeps modal x in registers
    case 48: rc = launch_v5<224, 56, 1>(p, mu, lmbda, u, v, y, st); break;
    case 49: rc = launch_v5<216, 72, 1>(p, mu, lmbda, u, v, y, st); break;
    case 16: k_hex8_nh_hvp_v2<2, 0><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y); break;
    case 20: k_hex8_nh_hvp_v2<3, 1><<<grid_for(p->n_elems), kBlock, 42 * kBlock * sizeof(double), st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y); break;
    // r02 default: pair kernel, all three modal fields staged, 168 registers -> 3 CTAs (12 warps) per SM: 0.397 ms at 128^3
    // (r01 default, variant 23: X and v staged, 244 registers, 2 CTAs per SM: 0.414 ms with the r02 loop body)
    default: rc = launch_v3<3, 2>(p, mu, lmbda, u, v, y, st); break;
  }
  if (rc != TATVA_OK) return rc;
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// y += H(u) v (y is NOT zeroed here) with the per-CTA partial sums of v . H v written to dot_partials (see DOT in
// k_hex8_nh_hvp_v3): the CG's p . A p for free.
int hex8_nh_hvp_modal_dot(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y,
                          double* dot_partials, cudaStream_t st) {
  constexpr size_t smem = (size_t)63 * kBlock * sizeof(double);
  static SmemOptIn configured;
  const int rc = opt_in_smem(k_hex8_nh_hvp_v3<3, 2, 0, false, 0, 1, 0, 0, 1>, smem, configured);
  if (rc != TATVA_OK) return rc;
  k_hex8_nh_hvp_v3<3, 2, 0, false, 0, 1, 0, 0, 1><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu * (1.0 / 512.0), lmbda * (1.0 / 512.0), u, v, y, nullptr, dot_partials);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int hex8_nh_hvp_modal_lifted(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v_red,
                             const int32_t* map, double* y_red, cudaStream_t st, double* dot_partials) {
  // all three modal fields staged, 3 CTAs per SM (the r02 default of the unconstrained kernel)
  constexpr size_t smem = (size_t)63 * kBlock * sizeof(double);
  static SmemOptIn configured[2];
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);
  if (dot_partials) {
    const int rc = opt_in_smem(k_hex8_nh_hvp_v3<3, 2, 0, true, 0, 1, 0, 0, 1>, smem, configured[1]);
    if (rc != TATVA_OK) return rc;
    k_hex8_nh_hvp_v3<3, 2, 0, true, 0, 1, 0, 0, 1><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu_s, lm_s, u, v_red, y_red, map, dot_partials);
  } else {
    const int rc = opt_in_smem(k_hex8_nh_hvp_v3<3, 2, 0, true>, smem, configured[0]);
    if (rc != TATVA_OK) return rc;
    k_hex8_nh_hvp_v3<3, 2, 0, true><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu_s, lm_s, u, v_red, y_red, map);
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int hex8_nh_residual_modal(const tatva_plan* p, double mu, double lmbda, const double* u, double* y, cudaStream_t st) {
  if (p->variant == 0) {  // default: y cleared under the kernel (launch_behind_zero)
    const int rc = launch_behind_zero(k_hex8_nh_residual_v3<2, 0>, grid_for(p->n_elems), kBlock, 0, st, p->zero_output != 0, y, p->n_nodes * 3,
                                      p->coords, p->conn, p->n_elems, mu, lmbda, u, y);
    if (rc != TATVA_OK) return rc;
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * p->n_nodes * 3, st));
  // variants (tatva_plan_set_variant), r01 at 128^3: 2 = first modal kernel, 8 rolled points (0.422 ms); 3 = pair
  // kernel, all in registers (0.382); 4 = pair kernel, modal coordinates staged, 2 CTAs / SM (0.394);
  // 5 = staged, 3 CTAs / SM (0.369, the r01 default).  r02, with the branch-free log / reciprocal: registers 0.3205,
  // staged at 3 CTAs 0.345 -> default = all in registers, 2 CTAs / SM
  switch (p->variant) {
    case 2: k_hex8_nh_residual_modal<<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, y); break;
    case 3: k_hex8_nh_residual_v3<2, 0><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, y); break;
    case 4: k_hex8_nh_residual_v3<2, 1><<<grid_for(p->n_elems), kBlock, 21 * kBlock * sizeof(double), st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, y); break;
    case 5: k_hex8_nh_residual_v3<3, 1><<<grid_for(p->n_elems), kBlock, 21 * kBlock * sizeof(double), st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, y); break;
    default: k_hex8_nh_residual_v3<2, 0><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, y); break;
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int hex8_nh_energy_modal_partials(const tatva_plan* p, double mu, double lmbda, const double* u, cudaStream_t st) {
  switch (p->variant) {  // 2 = first modal kernel (8 rolled points), 3 = pair kernel at 2 CTAs / SM, default = pair kernel, 3 CTAs / SM
    case 2: k_hex8_nh_energy_modal<<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, p->scratch); break;
    case 3: k_hex8_nh_energy_v3<2><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, p->scratch); break;
    default: k_hex8_nh_energy_v3<3><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, p->scratch); break;
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

static int tet4_nh_pipe(const tatva_plan* p, bool hvp, double mu, double lmbda, const double* u, const double* v, double* y,
                        cudaStream_t st) {
  constexpr size_t smem = grouped_scatter_smem<4, 3>(kBlock / 32);
  static int cache[2][64];
  int grid = 0, rc;
  if (hvp) {
    if ((rc = resident_grid(k_tet4_nh_pipe<true>, smem, p->n_elems, cache[1], &grid)) != TATVA_OK) return rc;
    k_tet4_nh_pipe<true><<<grid, kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y);
  } else {
    if ((rc = resident_grid(k_tet4_nh_pipe<false>, smem, p->n_elems, cache[0], &grid)) != TATVA_OK) return rc;
    k_tet4_nh_pipe<false><<<grid, kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mu, lmbda, u, nullptr, y);
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tet4_nh_hvp_ref(const tatva_plan* p, double mu, double lmbda, const double* u, const double* v, double* y, cudaStream_t st) {
  if (p->variant == 30) {
    if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * p->n_nodes * 3, st));
    return tet4_nh_pipe(p, true, mu, lmbda, u, v, y, st);
  }
  const int rc = launch_behind_zero(k_tet4_nh_ref<true>, grid_for(p->n_elems), kBlock, grouped_scatter_smem<4, 3>(kBlock / 32), st, p->zero_output != 0, y, p->n_nodes * 3,
                                    p->coords, p->conn, p->n_elems, mu, lmbda, u, v, y);
  if (rc != TATVA_OK) return rc;
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tet4_nh_residual_ref(const tatva_plan* p, double mu, double lmbda, const double* u, double* y, cudaStream_t st) {
  if (p->variant == 30) {
    if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * p->n_nodes * 3, st));
    return tet4_nh_pipe(p, false, mu, lmbda, u, nullptr, y, st);
  }
  const int rc = launch_behind_zero(k_tet4_nh_ref<false>, grid_for(p->n_elems), kBlock, grouped_scatter_smem<4, 3>(kBlock / 32), st, p->zero_output != 0, y, p->n_nodes * 3,
                                    p->coords, p->conn, p->n_elems, mu, lmbda, u, nullptr, y);
  if (rc != TATVA_OK) return rc;
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// The reference-space arithmetic of k_tet4_nh_ref as the element body of the warp-cooperative kernel (wc.cuh).
namespace {
template <int MINB>
struct Tet4NHRefBody {
  static constexpr int min_ctas = MINB;
  template <int MODE>
  TATVA_D static void run(const NeoHookean& m, const double (&X)[4][3], const double (&U)[4][3], const double (&V)[4][3],
                          double (&Y)[4][3]) {
    double J[3][3], Fr[3][3], Gv[3][3], Q[3][3];  // J[d][c] = dX_c/dxi_d ; Fr[i][d] = dx_i/dxi_d
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J[d][c] = X[d + 1][c] - X[0][c];
        Fr[c][d] = J[d][c] + (U[d + 1][c] - U[0][c]);
        if constexpr (MODE == MODE_HVP) Gv[c][d] = V[d + 1][c] - V[0][c];
      }
    if constexpr (MODE == MODE_HVP) point_flux(J, Fr, Gv, m.mu * (1.0 / 6.0), m.lmbda * (1.0 / 6.0), Q);
    else point_flux_residual(J, Fr, m.mu * (1.0 / 6.0), m.lmbda * (1.0 / 6.0), Q);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Y[1][i] = Q[i][0];
      Y[2][i] = Q[i][1];
      Y[3][i] = Q[i][2];
      Y[0][i] = -(Q[i][0] + Q[i][1] + Q[i][2]);
    }
  }
};
}  // namespace

int tet4_nh_wc(const tatva_plan* p, bool hvp, double mu, double lmbda, const double* u, const double* v, double* y,
               cudaStream_t st) {
  const NeoHookean m{mu, lmbda};
  // A/B variants (profiles/r02_tet4_node_schedule.jsonl): 32 = the other occupancy point, 33 = the element's own gather +
  // per-tile node sums, 34 = shuffle gather + per-warp sector-grouped scatter
  if (p->variant == 32) {
    if (hvp) return launch_fused_wc<Tet4, NeoHookean, MODE_HVP, Tet4NHRefBody<5>>(p, m, u, v, y, st);
    return launch_fused_wc<Tet4, NeoHookean, MODE_RESIDUAL, Tet4NHRefBody<4>>(p, m, u, v, y, st);
  }
  if (p->variant == 33) {
    if (hvp) return launch_fused_wc<Tet4, NeoHookean, MODE_HVP, Tet4NHRefBody<5>, false, true>(p, m, u, v, y, st);
    return launch_fused_wc<Tet4, NeoHookean, MODE_RESIDUAL, Tet4NHRefBody<5>, false, true>(p, m, u, v, y, st);
  }
  if (p->variant == 34) {
    if (hvp) return launch_fused_wc<Tet4, NeoHookean, MODE_HVP, Tet4NHRefBody<5>, true, false>(p, m, u, v, y, st);
    return launch_fused_wc<Tet4, NeoHookean, MODE_RESIDUAL, Tet4NHRefBody<5>, true, false>(p, m, u, v, y, st);
  }
  // measured: the HVP is faster at 4 CTAs per SM without spills, the residual at 5
  if (hvp) return launch_fused_wc<Tet4, NeoHookean, MODE_HVP, Tet4NHRefBody<4>>(p, m, u, v, y, st);
  return launch_fused_wc<Tet4, NeoHookean, MODE_RESIDUAL, Tet4NHRefBody<5>>(p, m, u, v, y, st);
}

int tet4_nh_tiled(const tatva_plan* p, bool hvp, double mu, double lmbda, const double* u, const double* v, double* y,
                  cudaStream_t st) {
  TATVA_CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * p->n_nodes * 3, st));
  const size_t stage = (size_t)((hvp ? 9 : 6) * p->tile_max_unique) * sizeof(double) + (size_t)p->tile_max_unique * sizeof(int);
  const size_t scat = grouped_scatter_smem<4, 3>(kBlock / 32);
  const size_t smem = ((stage > scat ? stage : scat) + 15) / 16 * 16;
  if (smem > 200 * 1024) return TATVA_E_UNSUPPORTED;
  static SmemOptIn configured[2];
  if (smem > 48 * 1024) {
    const int rc = hvp ? opt_in_smem(k_tet4_nh_tiled<true>, 200 * 1024, configured[1]) : opt_in_smem(k_tet4_nh_tiled<false>, 200 * 1024, configured[0]);
    if (rc != TATVA_OK) return rc;
  }
  const int grid = grid_for(p->n_elems);
  if (hvp) k_tet4_nh_tiled<true><<<grid, kBlock, smem, st>>>(p->coords, p->tile_ptr, p->tile_nodes, p->tile_conn, p->n_elems, p->tile_max_unique, mu, lmbda, u, v, y);
  else k_tet4_nh_tiled<false><<<grid, kBlock, smem, st>>>(p->coords, p->tile_ptr, p->tile_nodes, p->tile_conn, p->n_elems, p->tile_max_unique, mu, lmbda, u, v, y);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// ---------------------------------------------------------------------------------------------
// Host probe: the per-element arithmetic of the pair kernels (k_hex8_nh_hvp_v3, k_hex8_nh_residual_v3,
// k_hex8_nh_energy_v3) executed on the CPU with the SAME device functions compiled for the host (to_modal_raw,
// ref_grad8_pair, point_flux, point_flux_residual, point_energy, accumulate_pair, from_modal_raw), so the kernels'
// arithmetic can be checked against the oracle without a GPU.  Staging and gather / scatter are data movement and are
// not part of it.  mode: 0 energy (out[0]), 1 residual, 2 HVP (out[8][3], node-major like the nodal vectors).
// ---------------------------------------------------------------------------------------------
namespace {
void probe_hex8_pairs(int mode, const double* X, const double* u, const double* v, double mu, double lmbda, double* out) {
  double hX[3][7], hx[3][7], hv[3][7], hu[3][7];
  for (int c = 0; c < 3; ++c) {
    double fX[8], fu[8], fv[8];
    for (int n = 0; n < 8; ++n) {
      fX[n] = X[n * 3 + c];
      fu[n] = u[n * 3 + c];
      fv[n] = v ? v[n * 3 + c] : 0.0;
    }
    to_modal_raw(fX, hX[c]);
    to_modal_raw(fu, hx[c]);
    to_modal_raw(fv, hv[c]);
    for (int k = 0; k < 7; ++k) {
      hu[c][k] = hx[c][k];
      hx[c][k] += hX[c][k];
    }
  }
  double R[3][7] = {};
  double energy = 0.0;
  const double mu_s = mu * (1.0 / 512.0), lm_s = lmbda * (1.0 / 512.0);
  for (int pq = 0; pq < 4; ++pq) {
    const double sy = kPairSignsHost[pq][0], sz = kPairSignsHost[pq][1], syz = kPairSignsHost[pq][2];
    double Jm[3][3], Jp[3][3], Frm[3][3], Frp[3][3], Gvm[3][3], Gvp[3][3];
    for (int c = 0; c < 3; ++c) {
      double gm[3], gp[3];
      ref_grad8_pair(hX[c], sy, sz, gm, gp);
      for (int d = 0; d < 3; ++d) {
        Jm[d][c] = gm[d];
        Jp[d][c] = gp[d];
      }
    }
    for (int i = 0; i < 3; ++i) {
      ref_grad8_pair(hx[i], sy, sz, Frm[i], Frp[i]);
      ref_grad8_pair(hv[i], sy, sz, Gvm[i], Gvp[i]);
    }
    if (mode == 0) {
      energy += point_energy(Jm, Frm, mu, lmbda) + point_energy(Jp, Frp, mu, lmbda);
      continue;
    }
    double Qm[3][3], Qp[3][3];
    if (mode == 4 || mode == 5) {  // Operator.grad of the field u / integration weights, as k_hex8_grad_modal / k_hex8_weights_modal
      const int qm = (pq == 0) ? 0 : (pq == 1) ? 3 : (pq == 2) ? 4 : 7, qp = (pq == 0) ? 1 : (pq == 1) ? 2 : (pq == 2) ? 5 : 6;
      if (mode == 5) {
        double Kc[3][3], dm, dp;
        adjugate(Jm, Kc, dm);
        adjugate(Jp, Kc, dp);
        out[qm] = dm * (1.0 / 512.0);
        out[qp] = dp * (1.0 / 512.0);
        continue;
      }
      double Km[3][3], Kp[3][3];
      inverse_of(Jm, Km);
      inverse_of(Jp, Kp);
      for (int c = 0; c < 3; ++c) {
        double gm[3], gp[3];
        ref_grad8_pair(hu[c], sy, sz, gm, gp);
        for (int j = 0; j < 3; ++j) {
          out[(qm * 3 + c) * 3 + j] = fma(Km[j][2], gm[2], fma(Km[j][1], gm[1], Km[j][0] * gm[0]));
          out[(qp * 3 + c) * 3 + j] = fma(Kp[j][2], gp[2], fma(Kp[j][1], gp[1], Kp[j][0] * gp[0]));
        }
      }
      continue;
    }
    if (mode == 3) {  // the HVP through the geometry cache (point_geometry at plan time, point_flux_geo in the kernel)
      double gm8[8], gp8[8];
      point_geometry(Jm, gm8);
      point_geometry(Jp, gp8);
      point_flux_geo(gm8, Frm, Gvm, mu_s, lm_s, Qm);
      point_flux_geo(gp8, Frp, Gvp, mu_s, lm_s, Qp);
    } else if (mode == 2) {
      point_flux(Jm, Frm, Gvm, mu_s, lm_s, Qm);
      point_flux(Jp, Frp, Gvp, mu_s, lm_s, Qp);
    } else {
      point_flux_residual(Jm, Frm, mu_s, lm_s, Qm);
      point_flux_residual(Jp, Frp, mu_s, lm_s, Qp);
    }
    accumulate_pair(Qm, Qp, sy, sz, syz, R);
  }
  if (mode == 0) {
    out[0] = energy * (1.0 / 512.0);
    return;
  }
  if (mode == 4 || mode == 5) return;
  for (int i = 0; i < 3; ++i) {
    double f[8];
    from_modal_raw(R[i], f);
    for (int n = 0; n < 8; ++n) out[n * 3 + i] = f[n];
  }
}
// the body of k_tet4_nh_ref / k_tet4_nh_pipe / k_tet4_nh_tiled for one element (mode 1 residual, 2 HVP)
void probe_tet4_ref(int mode, const double* X, const double* u, const double* v, double mu, double lmbda, double* out) {
  double J[3][3], Fr[3][3], Gv[3][3], Q[3][3];
  for (int d = 0; d < 3; ++d)
    for (int c = 0; c < 3; ++c) {
      J[d][c] = X[(d + 1) * 3 + c] - X[c];
      Fr[c][d] = J[d][c] + (u[(d + 1) * 3 + c] - u[c]);
      Gv[c][d] = v ? v[(d + 1) * 3 + c] - v[c] : 0.0;
    }
  if (mode == 2) point_flux(J, Fr, Gv, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  else point_flux_residual(J, Fr, mu * (1.0 / 6.0), lmbda * (1.0 / 6.0), Q);
  for (int i = 0; i < 3; ++i) {
    out[3 + i] = Q[i][0];
    out[6 + i] = Q[i][1];
    out[9 + i] = Q[i][2];
    out[i] = -(Q[i][0] + Q[i][1] + Q[i][2]);
  }
}
}  // namespace

}  // namespace tatva

extern "C" int tatva_probe_tet4_nh_ref(int mode, const double* X, const double* u, const double* v, double mu, double lmbda,
                                       double* out) {
  if (!X || !u || !out || mode < 1 || mode > 2 || (mode == 2 && !v)) return TATVA_E_INVALID;
  tatva::probe_tet4_ref(mode, X, u, v, mu, lmbda, out);
  return TATVA_OK;
}

extern "C" int tatva_probe_hex8_nh_modal(int mode, const double* X, const double* u, const double* v, double mu, double lmbda,
                                         double* out) {
  if (!X || !u || !out || mode < 0 || mode > 5 || ((mode == 2 || mode == 3) && !v)) return TATVA_E_INVALID;
  tatva::probe_hex8_pairs(mode, X, u, v, mu, lmbda, out);
  return TATVA_OK;
}
