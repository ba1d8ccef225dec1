// Fused element kernels as TEMPLATES over (element, constitutive law): energy / residual / HVP (k_fused), the HVP with a
// Lifter folded in (k_hvp_lifted) and the Hessian diagonal (k_hessian_diag).  They live in a header because the same
// text is compiled twice: ahead of time for the built-in laws (generic.cu) and at run time, by NVRTC, for a law the user
// supplies as CUDA source (user_law.cpp; README.md:93 — in the reference the density is user code).
#pragma once
#include "common.cuh"

namespace tatva {

// ---- gather helpers ---------------------------------------------------------------------------

template <class El>
TATVA_D void load_conn(const int32_t* __restrict__ conn, int64_t e, int (&nd)[El::npe]) {
  if constexpr (El::npe == 4) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(conn) + e);
    nd[0] = t.x; nd[1] = t.y; nd[2] = t.z; nd[3] = t.w;
  } else if constexpr (El::npe == 8) {
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e);
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(conn) + 2 * e + 1);
    nd[0] = t0.x; nd[1] = t0.y; nd[2] = t0.z; nd[3] = t0.w;
    nd[4] = t1.x; nd[5] = t1.y; nd[6] = t1.z; nd[7] = t1.w;
  } else {
#pragma unroll
    for (int n = 0; n < El::npe; ++n) nd[n] = __ldg(conn + e * El::npe + n);
  }
}

template <int NPE, int W>
TATVA_D void gather_rows(const double* __restrict__ src, const int (&nd)[NPE], double (&dst)[NPE][W]) {
#pragma unroll
  for (int n = 0; n < NPE; ++n) load_row<W>(src, nd[n], dst[n]);
}

// ---- deterministic reductions -----------------------------------------------------------------

TATVA_D double block_sum(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;  // valid on thread 0
}

// ---- fused energy / residual / HVP ------------------------------------------------------------

enum { MODE_ENERGY = 0, MODE_RESIDUAL = 1, MODE_HVP = 2 };

template <class El, class Mat>
TATVA_HD void qp_state(const double (&dNdX)[El::dim][El::npe], const double (&N)[El::npe],
                      const double (&U)[El::npe][Mat::dpn], typename Mat::S& s) {
#pragma unroll
  for (int c = 0; c < Mat::dpn; ++c) {
#pragma unroll
    for (int j = 0; j < El::dim; ++j) {
      double t = 0.0;
#pragma unroll
      for (int n = 0; n < El::npe; ++n) t += dNdX[j][n] * U[n][c];
      s.G[c][j] = t;
    }
    if (c >= Mat::val_lo) {
      double t = 0.0;
#pragma unroll
      for (int n = 0; n < El::npe; ++n) t += N[n] * U[n][c];
      s.val[c] = t;
    }
  }
}

template <class El, class Mat, int MODE>
__global__ void __launch_bounds__(kBlock) k_fused(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                  int64_t E, Mat mat, const double* __restrict__ u,
                                                  const double* __restrict__ v, double* __restrict__ y,
                                                  double* __restrict__ partials) {
  static_assert(El::dim == Mat::dim, "element / law dimension mismatch");
  constexpr int dpn = Mat::dpn;
  extern __shared__ double sm_fused[];
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double energy = 0.0;
  int nd[El::npe];
  double Y[El::npe][dpn];
#pragma unroll
  for (int n = 0; n < El::npe; ++n) {
    nd[n] = 0;
#pragma unroll
    for (int c = 0; c < dpn; ++c) Y[n][c] = 0.0;
  }
  if (e < E) {
    load_conn<El>(conn, e, nd);
    double X[El::npe][El::dim], U[El::npe][dpn], V[El::npe][dpn];
    gather_rows(coords, nd, X);
    gather_rows(u, nd, U);
    if constexpr (MODE == MODE_HVP) gather_rows(v, nd, V);

#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dNdX[El::dim][El::npe], N[El::npe];
      const double W = geometry<El>(q, X, dNdX) * El::weight(q);
      El::N(q, N);
      typename Mat::S s, ds, f;
      typename Mat::Cache cache;
      qp_state<El, Mat>(dNdX, N, U, s);
      mat.prepare(s, cache);
      if constexpr (MODE == MODE_ENERGY) {
        energy += W * mat.psi(s, cache);
      } else {
        if constexpr (MODE == MODE_RESIDUAL) {
          mat.first(s, cache, f);
        } else {
          qp_state<El, Mat>(dNdX, N, V, ds);
          mat.second(s, cache, ds, f);
        }
#pragma unroll
        for (int n = 0; n < El::npe; ++n)
#pragma unroll
          for (int c = 0; c < dpn; ++c) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < El::dim; ++j) t += f.G[c][j] * dNdX[j][n];
            if (c >= Mat::val_lo) t += f.val[c] * N[n];
            Y[n][c] += W * t;
          }
      }
    }
  }
  if constexpr (MODE != MODE_ENERGY) {
    double* wsm = sm_fused + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<El::npe, dpn>();
    grouped_scatter<El::npe, dpn>(y, nd, Y, e < E, wsm);
  }
  if constexpr (MODE == MODE_ENERGY) {
    energy = block_sum(energy);
    if (threadIdx.x == 0) partials[blockIdx.x] = energy;
  }
}

// HVP with the Lifter folded in: v and y are REDUCED vectors, `map` (n_nodes*dpn int32) sends a full DOF to its
// reduced index or to -1 (no driver: the homogeneous lift is 0 there and the contribution is dropped).
template <class El, class Mat>
__global__ void __launch_bounds__(kBlock) k_hvp_lifted(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                       int64_t E, Mat mat, const double* __restrict__ u,
                                                       const double* __restrict__ v, const int32_t* __restrict__ map,
                                                       double* __restrict__ y) {
  constexpr int dpn = Mat::dpn;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim], U[El::npe][dpn], V[El::npe][dpn], Y[El::npe][dpn];
  gather_rows(coords, nd, X);
  gather_rows(u, nd, U);
#pragma unroll
  for (int n = 0; n < El::npe; ++n)
#pragma unroll
    for (int c = 0; c < dpn; ++c) {
      const int32_t m = __ldg(map + (int64_t)nd[n] * dpn + c);
      V[n][c] = m >= 0 ? __ldg(v + m) : 0.0;
      Y[n][c] = 0.0;
    }
#pragma unroll 1
  for (int q = 0; q < El::num_q(); ++q) {
    double dNdX[El::dim][El::npe], N[El::npe];
    const double W = geometry<El>(q, X, dNdX) * El::weight(q);
    El::N(q, N);
    typename Mat::S s, ds, f;
    typename Mat::Cache cache;
    qp_state<El, Mat>(dNdX, N, U, s);
    mat.prepare(s, cache);
    qp_state<El, Mat>(dNdX, N, V, ds);
    mat.second(s, cache, ds, f);
#pragma unroll
    for (int n = 0; n < El::npe; ++n)
#pragma unroll
      for (int c = 0; c < dpn; ++c) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < El::dim; ++j) t += f.G[c][j] * dNdX[j][n];
        if (c >= Mat::val_lo) t += f.val[c] * N[n];
        Y[n][c] += W * t;
      }
  }
#pragma unroll
  for (int n = 0; n < El::npe; ++n)
#pragma unroll
    for (int c = 0; c < dpn; ++c) {
      const int32_t m = __ldg(map + (int64_t)nd[n] * dpn + c);
      if (m >= 0) atomicAdd(y + m, Y[n][c]);
    }
}

// ---- Hessian diagonal (Jacobi preconditioner) -------------------------------------------------------
// diag[dpn*node_b + k] += sum_q W * (d2psi : unit_(b,k)) . unit_(b,k): the (b,k)/(b,k) entry of the element
// stiffness, with the geometry and the material state of a point computed once for all npe*dpn unit directions.
template <class El, class Mat>
__global__ void __launch_bounds__(kBlock) k_hessian_diag(const double* __restrict__ coords,
                                                         const int32_t* __restrict__ conn, int64_t E, Mat mat,
                                                         const double* __restrict__ u, double* __restrict__ diag) {
  constexpr int dpn = Mat::dpn;
  extern __shared__ double sm_fused[];
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nd[El::npe];
  double Y[El::npe][dpn];
#pragma unroll
  for (int n = 0; n < El::npe; ++n) {
    nd[n] = 0;
#pragma unroll
    for (int c = 0; c < dpn; ++c) Y[n][c] = 0.0;
  }
  if (e < E) {
    load_conn<El>(conn, e, nd);
    double X[El::npe][El::dim], U[El::npe][dpn];
    gather_rows(coords, nd, X);
    gather_rows(u, nd, U);
#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dNdX[El::dim][El::npe], N[El::npe];
      const double W = geometry<El>(q, X, dNdX) * El::weight(q);
      El::N(q, N);
      typename Mat::S s;
      typename Mat::Cache cache;
      qp_state<El, Mat>(dNdX, N, U, s);
      mat.prepare(s, cache);
#pragma unroll
      for (int b = 0; b < El::npe; ++b) {
#pragma unroll
        for (int k = 0; k < dpn; ++k) {
          typename Mat::S ds, f;
#pragma unroll
          for (int c = 0; c < dpn; ++c) {
#pragma unroll
            for (int j = 0; j < El::dim; ++j) ds.G[c][j] = (c == k) ? dNdX[j][b] : 0.0;
            ds.val[c] = (c == k) ? N[b] : 0.0;
          }
          mat.second(s, cache, ds, f);
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < El::dim; ++j) t += f.G[k][j] * dNdX[j][b];
          if (k >= Mat::val_lo) t += f.val[k] * N[b];
          Y[b][k] += W * t;
        }
      }
    }
  }
  double* wsm = sm_fused + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<El::npe, dpn>();
  grouped_scatter<El::npe, dpn>(diag, nd, Y, e < E, wsm);
}

// ---- CSR assembly -----------------------------------------------------------------------------
// Column (b,k) of the element stiffness is the element-local HVP with the unit direction
// "component k of node b"; rows (a,i) go to data[indptr[dpn*node_a + i] + pos[e,a,b] + k].

}  // namespace tatva
