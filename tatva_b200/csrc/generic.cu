// Generic element-per-thread kernels: the quadrature-loop building blocks behind
// Operator.grad / eval / integrate / map, and the fused energy / residual / HVP / CSR-assembly
// kernels for every (element, law) pair.  One thread owns one element: it gathers the element's
// nodal rows by connectivity (the `v[self.mesh.elements]` of tatva/operator.py:221), runs the
// quadrature loop in FP64 registers and scatter-adds with RED.F64 atomics (the transpose of that
// gather, which is what jax.grad produces in the reference).
#include <new>

#include "common.cuh"
#include "fused.cuh"
#include "wc.cuh"

namespace tatva {

// ---- Operator building blocks (runtime number of value components) -----------------------------

template <class El>
__global__ void __launch_bounds__(kBlock) k_weights(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                    int64_t E, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim];
  gather_rows(coords, nd, X);
#pragma unroll
  for (int q = 0; q < El::num_q(); ++q) out[e * El::num_q() + q] = det_jacobian<El>(q, X) * El::weight(q);
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_grad(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                 int64_t E, const double* __restrict__ u, int nv,
                                                 double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim];
  gather_rows(coords, nd, X);
#pragma unroll 1
  for (int q = 0; q < El::num_q(); ++q) {
    double dNdX[El::gdim][El::npe];
    geometry<El>(q, X, dNdX);
    for (int c = 0; c < nv; ++c) {
      double ue[El::npe];
#pragma unroll
      for (int n = 0; n < El::npe; ++n) ue[n] = __ldg(u + (int64_t)nd[n] * nv + c);
#pragma unroll
      for (int j = 0; j < El::gdim; ++j) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < El::npe; ++n) s += dNdX[j][n] * ue[n];
        out[((e * El::num_q() + q) * nv + c) * El::gdim + j] = s;
      }
    }
  }
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_grad_adjoint(const double* __restrict__ coords,
                                                         const int32_t* __restrict__ conn, int64_t E,
                                                         const double* __restrict__ g, int nv, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim];
  gather_rows(coords, nd, X);
#pragma unroll 1
  for (int q = 0; q < El::num_q(); ++q) {
    double dNdX[El::gdim][El::npe];
    geometry<El>(q, X, dNdX);
    for (int c = 0; c < nv; ++c) {
      double gj[El::gdim];
#pragma unroll
      for (int j = 0; j < El::gdim; ++j) gj[j] = __ldg(g + ((e * El::num_q() + q) * nv + c) * El::gdim + j);
#pragma unroll
      for (int n = 0; n < El::npe; ++n) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < El::gdim; ++j) s += gj[j] * dNdX[j][n];
        atomicAdd(y + (int64_t)nd[n] * nv + c, s);
      }
    }
  }
}

// ---- warp-staged variants: element-major (E, Q, v[, d]) arrays are written / read through shared memory ----
// Each lane produces (or consumes) the CH = nq*nv[*dim] contiguous doubles of its own element; staging them per
// warp turns 32 strided 8-byte accesses per instruction into one contiguous 32*CH-double block.  Row stride
// S = CH | 1 keeps the lane-strided side conflict-free.

TATVA_D void warp_block_store(double* __restrict__ dst, const double* st, int S, int CH, int count) {
  const int lane = threadIdx.x & 31;
  for (int t = lane; t < count * CH; t += 32) {
    const int j = t / CH;
    dst[t] = st[j * S + (t - j * CH)];
  }
}
// Loads go out in batches of 8 before the first staging store: a rolled load -> store loop keeps ONE 256-byte load in
// flight per warp, and a streaming read then runs at the latency, not the bandwidth, of HBM (Hex8 grad adjoint at 128^3:
// 0.55 -> 0.42 ms from this alone).
TATVA_D void warp_block_load(const double* __restrict__ src, double* st, int S, int CH, int count) {
  const int lane = threadIdx.x & 31;
  const int total = count * CH;
  const int q32 = 32 / CH, r32 = 32 - q32 * CH;
  int t = lane, j = lane / CH, r = lane - j * CH;
  for (; t + 7 * 32 < total; t += 8 * 32) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(src + t + 32 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      st[j * S + r] = v[k];
      j += q32;
      r += r32;
      if (r >= CH) {
        r -= CH;
        ++j;
      }
    }
  }
  for (; t < total; t += 32) {
    st[j * S + r] = __ldg(src + t);
    j += q32;
    r += r32;
    if (r >= CH) {
      r -= CH;
      ++j;
    }
  }
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_grad_staged(const double* __restrict__ coords,
                                                        const int32_t* __restrict__ conn, int64_t E,
                                                        const double* __restrict__ u, int nv, double* __restrict__ out) {
  extern __shared__ double sm_stage[];
  const int CH = El::num_q() * nv * El::gdim, S = CH | 1;
  const int lane = threadIdx.x & 31;
  double* st = sm_stage + (size_t)(threadIdx.x >> 5) * 32 * S;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) {
    int nd[El::npe];
    load_conn<El>(conn, e, nd);
    double X[El::npe][El::dim];
    gather_rows(coords, nd, X);
#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dNdX[El::gdim][El::npe];
      geometry<El>(q, X, dNdX);
      for (int c = 0; c < nv; ++c) {
        double ue[El::npe];
#pragma unroll
        for (int n = 0; n < El::npe; ++n) ue[n] = __ldg(u + (int64_t)nd[n] * nv + c);
#pragma unroll
        for (int j = 0; j < El::gdim; ++j) {
          double t = 0.0;
#pragma unroll
          for (int n = 0; n < El::npe; ++n) t += dNdX[j][n] * ue[n];
          st[lane * S + (q * nv + c) * El::gdim + j] = t;
        }
      }
    }
  }
  __syncwarp();
  const int64_t e0 = e - lane;
  if (e0 < E) warp_block_store(out + e0 * CH, st, S, CH, (int)min((int64_t)32, E - e0));
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_grad_adjoint_staged(const double* __restrict__ coords,
                                                                const int32_t* __restrict__ conn, int64_t E,
                                                                const double* __restrict__ g, int nv,
                                                                double* __restrict__ y) {
  extern __shared__ double sm_stage[];
  const int CH = El::num_q() * nv * El::gdim, S = CH | 1;
  const int lane = threadIdx.x & 31;
  double* st = sm_stage + (size_t)(threadIdx.x >> 5) * 32 * S;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = e - lane;
  if (e0 < E) warp_block_load(g + e0 * CH, st, S, CH, (int)min((int64_t)32, E - e0));
  __syncwarp();
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim];
  gather_rows(coords, nd, X);
#pragma unroll 1
  for (int q = 0; q < El::num_q(); ++q) {
    double dNdX[El::gdim][El::npe];
    geometry<El>(q, X, dNdX);
    for (int c = 0; c < nv; ++c) {
      double gj[El::gdim];
#pragma unroll
      for (int j = 0; j < El::gdim; ++j) gj[j] = st[lane * S + (q * nv + c) * El::gdim + j];
#pragma unroll
      for (int n = 0; n < El::npe; ++n) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < El::gdim; ++j) t += gj[j] * dNdX[j][n];
        atomicAdd(y + (int64_t)nd[n] * nv + c, t);
      }
    }
  }
}

// adjoint of grad with the per-node sums kept in registers over the quadrature loop (NV compile-time, <= 4):
// npe*NV REDs per element instead of nq*npe*NV, issued sector-grouped.
template <class El, int NV>
__global__ void __launch_bounds__(kBlock) k_grad_adjoint_acc(const double* __restrict__ coords,
                                                             const int32_t* __restrict__ conn, int64_t E,
                                                             const double* __restrict__ g, double* __restrict__ y) {
  extern __shared__ double sm_stage[];
  constexpr int CH = El::max_nq * NV * El::gdim, S = CH | 1;
  constexpr int SC = grouped_scatter_words<El::npe, NV>();  // grouped-scatter staging per warp
  constexpr int PER_WARP = (32 * S > SC) ? 32 * S : SC;
  const int lane = threadIdx.x & 31;
  double* st = sm_stage + (size_t)(threadIdx.x >> 5) * PER_WARP;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = e - lane;
  if (e0 < E) warp_block_load(g + e0 * CH, st, S, CH, (int)min((int64_t)32, E - e0));
  __syncwarp();
  int nd[El::npe];
  double Y[El::npe][NV];
#pragma unroll
  for (int n = 0; n < El::npe; ++n) {
    nd[n] = 0;
#pragma unroll
    for (int c = 0; c < NV; ++c) Y[n][c] = 0.0;
  }
  if (e < E) {
    load_conn<El>(conn, e, nd);
    double X[El::npe][El::dim];
    gather_rows(coords, nd, X);
#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dNdX[El::gdim][El::npe];
      geometry<El>(q, X, dNdX);
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        double gj[El::gdim];
#pragma unroll
        for (int j = 0; j < El::gdim; ++j) gj[j] = st[lane * S + (q * NV + c) * El::gdim + j];
#pragma unroll
        for (int n = 0; n < El::npe; ++n) {
          double t = Y[n][c];
#pragma unroll
          for (int j = 0; j < El::gdim; ++j) t = fma(gj[j], dNdX[j][n], t);
          Y[n][c] = t;
        }
      }
    }
  }
  __syncwarp();  // every lane is done reading its staged g before the buffer is reused
  grouped_scatter<El::npe, NV>(y, nd, Y, e < E, st);
}

// ---- one thread per (element, quadrature point) --------------------------------------------------------
// The (E, Q, ...) outputs are flat in t = e * nq + q, so thread t owns CH contiguous doubles and a warp owns 32 * CH:
// nq times more threads than element-per-thread (short dependent chains, full occupancy), the nq lanes of an
// element gather the same node rows (L1 broadcast), and the staging buffer is 32 * CH doubles per warp instead of
// 32 * nq * CH.  MEASURED SLOWER on B200 (r01, Hex8 64^3: grad 0.097 vs 0.086 ms, eval 0.038 vs 0.018, weights 0.026 vs
// 0.016): the nq-fold repeated gathers cost more L1 wavefronts than the shorter chains save.  Kept as plan variant 3.
template <class El>
__global__ void __launch_bounds__(kBlock) k_weights_qp(const double* __restrict__ coords,
                                                       const int32_t* __restrict__ conn, int64_t E,
                                                       double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * El::num_q()) return;
  const int64_t e = t / El::num_q();
  const int q = (int)(t - e * El::num_q());
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim];
  gather_rows(coords, nd, X);
  out[t] = det_jacobian<El>(q, X) * El::weight(q);
}

template <class El, bool GRAD>
__global__ void __launch_bounds__(kBlock) k_field_qp(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                     int64_t E, const double* __restrict__ u, int nv,
                                                     double* __restrict__ out) {
  extern __shared__ double sm_stage[];
  constexpr int G = GRAD ? El::gdim : 1;
  const int CH = nv * G, S = CH | 1;
  const int lane = threadIdx.x & 31;
  double* st = sm_stage + (size_t)(threadIdx.x >> 5) * 32 * S;
  const int64_t T = E * El::num_q();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    const int64_t e = t / El::num_q();
    const int q = (int)(t - e * El::num_q());
    int nd[El::npe];
    load_conn<El>(conn, e, nd);
    double B[G][El::npe];  // dNdX (GRAD) or N
    if constexpr (GRAD) {
      double X[El::npe][El::dim];
      gather_rows(coords, nd, X);
      geometry<El>(q, X, B);
    } else {
      El::N(q, B[0]);
    }
    for (int c = 0; c < nv; ++c) {
      double ue[El::npe];
#pragma unroll
      for (int n = 0; n < El::npe; ++n) ue[n] = __ldg(u + (int64_t)nd[n] * nv + c);
#pragma unroll
      for (int j = 0; j < G; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < El::npe; ++n) acc += B[j][n] * ue[n];
        st[lane * S + c * G + j] = acc;
      }
    }
  }
  __syncwarp();
  const int64_t t0 = t - lane;
  if (t0 < T) warp_block_store(out + t0 * CH, st, S, CH, (int)min((int64_t)32, T - t0));
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_eval_staged(const int32_t* __restrict__ conn, int64_t E,
                                                        const double* __restrict__ u, int nv, double* __restrict__ out) {
  extern __shared__ double sm_stage[];
  const int CH = El::num_q() * nv, S = CH | 1;
  const int lane = threadIdx.x & 31;
  double* st = sm_stage + (size_t)(threadIdx.x >> 5) * 32 * S;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) {
    int nd[El::npe];
    load_conn<El>(conn, e, nd);
    for (int c = 0; c < nv; ++c) {
      double ue[El::npe];
#pragma unroll
      for (int n = 0; n < El::npe; ++n) ue[n] = __ldg(u + (int64_t)nd[n] * nv + c);
#pragma unroll
      for (int q = 0; q < El::num_q(); ++q) {
        double N[El::npe];
        El::N(q, N);
        double t = 0.0;
#pragma unroll
        for (int n = 0; n < El::npe; ++n) t += N[n] * ue[n];
        st[lane * S + q * nv + c] = t;
      }
    }
  }
  __syncwarp();
  const int64_t e0 = e - lane;
  if (e0 < E) warp_block_store(out + e0 * CH, st, S, CH, (int)min((int64_t)32, E - e0));
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_eval(const int32_t* __restrict__ conn, int64_t E,
                                                 const double* __restrict__ u, int nv, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  for (int c = 0; c < nv; ++c) {
    double ue[El::npe];
#pragma unroll
    for (int n = 0; n < El::npe; ++n) ue[n] = __ldg(u + (int64_t)nd[n] * nv + c);
#pragma unroll
    for (int q = 0; q < El::num_q(); ++q) {
      double N[El::npe];
      El::N(q, N);
      double s = 0.0;
#pragma unroll
      for (int n = 0; n < El::npe; ++n) s += N[n] * ue[n];
      out[(e * El::num_q() + q) * nv + c] = s;
    }
  }
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_eval_adjoint(const int32_t* __restrict__ conn, int64_t E,
                                                         const double* __restrict__ g, int nv, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  for (int c = 0; c < nv; ++c) {
    double acc[El::npe];
#pragma unroll
    for (int n = 0; n < El::npe; ++n) acc[n] = 0.0;
#pragma unroll
    for (int q = 0; q < El::num_q(); ++q) {
      double N[El::npe];
      El::N(q, N);
      const double gq = __ldg(g + (e * El::num_q() + q) * nv + c);
#pragma unroll
      for (int n = 0; n < El::npe; ++n) acc[n] += N[n] * gq;
    }
#pragma unroll
    for (int n = 0; n < El::npe; ++n) atomicAdd(y + (int64_t)nd[n] * nv + c, acc[n]);
  }
}

template <class El>
__global__ void __launch_bounds__(kBlock) k_integrate_quad(const double* __restrict__ coords,
                                                           const int32_t* __restrict__ conn, int64_t E,
                                                           const double* __restrict__ cachedW,
                                                           const double* __restrict__ vals, int nv,
                                                           double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double W[El::max_nq];
  if (cachedW) {
#pragma unroll
    for (int q = 0; q < El::num_q(); ++q) W[q] = __ldg(cachedW + e * El::num_q() + q);
  } else {
    int nd[El::npe];
    load_conn<El>(conn, e, nd);
    double X[El::npe][El::dim];
    gather_rows(coords, nd, X);
#pragma unroll
    for (int q = 0; q < El::num_q(); ++q) W[q] = det_jacobian<El>(q, X) * El::weight(q);
  }
  for (int c = 0; c < nv; ++c) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < El::num_q(); ++q) s += __ldg(vals + (e * El::num_q() + q) * nv + c) * W[q];
    out[e * nv + c] = s;
  }
}

// gather / scatter over (element, node, component) with one thread per entry: coalesced on the
// element side, indexed on the nodal side.
__global__ void __launch_bounds__(256) k_gather(const int32_t* __restrict__ conn, int64_t total, int nv,
                                                const double* __restrict__ u, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t en = i / nv;
  const int c = (int)(i - en * nv);
  out[i] = __ldg(u + (int64_t)__ldg(conn + en) * nv + c);
}

// Four outputs per thread (stride 256: every store instruction still writes 2 KB contiguous per CTA), index arithmetic in
// 32 bits with a compile-time divisor for nv <= 4: the connectivity loads of the four go out together, then the four row
// loads — the one-output-per-thread kernel above holds one dependent load chain per thread and a 64-bit division.
template <int NVT>
__global__ void __launch_bounds__(256) k_gather4(const int32_t* __restrict__ conn, uint32_t total, int nv_rt,
                                                 const double* __restrict__ u, double* __restrict__ out) {
  const uint32_t nv = NVT ? (uint32_t)NVT : (uint32_t)nv_rt;
  const uint32_t base = blockIdx.x * 1024u + threadIdx.x;
  uint32_t en[4], c[4];
  int node[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = base + 256u * k;
    en[k] = i / nv;
    c[k] = i - en[k] * nv;
    node[k] = i < total ? __ldg(conn + en[k]) : 0;
  }
  double val[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) val[k] = __ldg(u + (int64_t)node[k] * nv + c[k]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = base + 256u * k;
    if (i < total) __stcs(out + i, val[k]);
  }
}

__global__ void __launch_bounds__(256) k_gather_adjoint(const int32_t* __restrict__ conn, int64_t total, int nv,
                                                        const double* __restrict__ g, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t en = i / nv;
  const int c = (int)(i - en * nv);
  atomicAdd(y + (int64_t)__ldg(conn + en) * nv + c, __ldg(g + i));
}
template <int NVT>
__global__ void __launch_bounds__(256) k_gather_adjoint4(const int32_t* __restrict__ conn, uint32_t total, int nv_rt,
                                                         const double* __restrict__ g, double* __restrict__ y) {
  const uint32_t nv = NVT ? (uint32_t)NVT : (uint32_t)nv_rt;
  const uint32_t base = blockIdx.x * 1024u + threadIdx.x;
  uint32_t c[4];
  int node[4];
  double val[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = base + 256u * k, en = i / nv;
    c[k] = i - en * nv;
    node[k] = i < total ? __ldg(conn + en) : -1;
    val[k] = i < total ? __ldcs(g + i) : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (node[k] >= 0) atomicAdd(y + (int64_t)node[k] * nv + c[k], val[k]);
}

// partial[b*nv + c] = sum over rows of chunk b; rows strided by blockDim
__global__ void __launch_bounds__(256) k_sum_rows_partial(const double* __restrict__ in, int64_t rows, int nv,
                                                          int64_t rows_per_block, double* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int c = 0; c < nv; ++c) {
    double s = 0.0;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) s += __ldg(in + r * nv + c);
    s = block_sum(s);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * nv + c] = s;
  }
}

__global__ void __launch_bounds__(256) k_sum_rows_final(const double* __restrict__ partial, int nblocks, int nv,
                                                        double* __restrict__ out) {
  for (int c = blockIdx.x; c < nv; c += gridDim.x) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(int64_t)b * nv + c];
    s = block_sum(s);
    if (threadIdx.x == 0) out[c] = s;
  }
}

// ---- Operator.interpolate: point location + one Newton step + shape functions at a free point ----------
// Shape functions of the plane elements at an arbitrary reference point (the element structs above only carry
// their quadrature points).  tatva/element/base.py:257-265, :286-328, :346-366, :395-445.
template <class El> struct PlaneShape;
template <> struct PlaneShape<Tri3> {
  TATVA_D static void N(double r, double s, double (&n)[3]) { n[0] = 1.0 - r - s; n[1] = r; n[2] = s; }
  TATVA_D static void dN(double, double, double (&d)[2][3]) {
    d[0][0] = -1.0; d[0][1] = 1.0; d[0][2] = 0.0;
    d[1][0] = -1.0; d[1][1] = 0.0; d[1][2] = 1.0;
  }
};
template <> struct PlaneShape<Quad4> {
  TATVA_D static void N(double r, double s, double (&n)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) n[k] = 0.25 * (1.0 + Quad4::sgn(k, 0) * r) * (1.0 + Quad4::sgn(k, 1) * s);
  }
  TATVA_D static void dN(double r, double s, double (&d)[2][4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      d[0][k] = 0.25 * Quad4::sgn(k, 0) * (1.0 + Quad4::sgn(k, 1) * s);
      d[1][k] = 0.25 * Quad4::sgn(k, 1) * (1.0 + Quad4::sgn(k, 0) * r);
    }
  }
};
template <> struct PlaneShape<Tri6> {
  TATVA_D static void N(double r, double s, double (&n)[6]) {
    const double t = 1.0 - r - s;
    n[0] = t * (2 * t - 1); n[1] = r * (2 * r - 1); n[2] = s * (2 * s - 1);
    n[3] = 4 * r * t; n[4] = 4 * r * s; n[5] = 4 * s * t;
  }
  TATVA_D static void dN(double r, double s, double (&d)[2][6]) {
    const double t = 1.0 - r - s;
    d[0][0] = -(4 * t - 1); d[0][1] = 4 * r - 1; d[0][2] = 0.0; d[0][3] = 4 * (t - r); d[0][4] = 4 * s; d[0][5] = -4 * s;
    d[1][0] = -(4 * t - 1); d[1][1] = 0.0; d[1][2] = 4 * s - 1; d[1][3] = -4 * r; d[1][4] = 4 * r; d[1][5] = 4 * (t - s);
  }
};
template <> struct PlaneShape<Quad8> {
  TATVA_D static void N(double r, double s, double (&n)[8]) {
    n[0] = 0.25 * (1 - r) * (1 - s) * (-r - s - 1);
    n[1] = 0.25 * (1 + r) * (1 - s) * (r - s - 1);
    n[2] = 0.25 * (1 + r) * (1 + s) * (r + s - 1);
    n[3] = 0.25 * (1 - r) * (1 + s) * (-r + s - 1);
    n[4] = 0.5 * (1 - r * r) * (1 - s);
    n[5] = 0.5 * (1 + r) * (1 - s * s);
    n[6] = 0.5 * (1 - r * r) * (1 + s);
    n[7] = 0.5 * (1 - r) * (1 - s * s);
  }
  TATVA_D static void dN(double r, double s, double (&d)[2][8]) {
    d[0][0] = 0.25 * (-2 * r - s) * (s - 1); d[0][1] = 0.25 * (-2 * r + s) * (s - 1);
    d[0][2] = 0.25 * (2 * r + s) * (s + 1);  d[0][3] = 0.25 * (2 * r - s) * (s + 1);
    d[0][4] = r * (s - 1); d[0][5] = 0.5 - 0.5 * s * s; d[0][6] = -r * (s + 1); d[0][7] = 0.5 * s * s - 0.5;
    d[1][0] = 0.25 * (-r - 2 * s) * (r - 1); d[1][1] = 0.25 * (-r + 2 * s) * (r + 1);
    d[1][2] = 0.25 * (r + 1) * (r + 2 * s);  d[1][3] = 0.25 * (r - 1) * (r - 2 * s);
    d[1][4] = 0.5 * r * r - 0.5; d[1][5] = -s * (r + 1); d[1][6] = 0.5 - 0.5 * r * r; d[1][7] = s * (r - 1);
  }
};
TATVA_D void first_quad_point(Tri3, double& r, double& s) { r = 1.0 / 3; s = 1.0 / 3; }
TATVA_D void first_quad_point(Quad4, double& r, double& s) { Quad4::xi(0, r, s); }
TATVA_D void first_quad_point(Tri6, double& r, double& s) { Tri6::xi(0, r, s); }
TATVA_D void first_quad_point(Quad8, double& r, double& s) { Quad8::xi(0, r, s); }

// Element search exactly as mesh.find_containing_polygons (tatva/mesh.py:294-388): the FIRST element, in element
// order, whose node loop (connectivity order) contains the point: bounding-box reject, then on-boundary
// (|cross| <= 1e-8 within the segment's box) OR an odd number of +x ray crossings.
template <class El>
TATVA_D bool loop_contains(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t e, double px,
                           double py, bool check_box) {
  constexpr int npe = El::npe;
  double vx[npe], vy[npe];
#pragma unroll
  for (int n = 0; n < npe; ++n) {
    const int64_t nd = __ldg(conn + e * npe + n);
    vx[n] = __ldg(coords + 2 * nd);
    vy[n] = __ldg(coords + 2 * nd + 1);
  }
  if (check_box) {
    double lx = vx[0], hx = vx[0], ly = vy[0], hy = vy[0];
#pragma unroll
    for (int n = 1; n < npe; ++n) {
      lx = fmin(lx, vx[n]); hx = fmax(hx, vx[n]); ly = fmin(ly, vy[n]); hy = fmax(hy, vy[n]);
    }
    if (!(px >= lx && px <= hx && py >= ly && py <= hy)) return false;
  }
  bool on_boundary = false;
  int crossings = 0;
#pragma unroll
  for (int n = 0; n < npe; ++n) {
    const double ax = vx[n], ay = vy[n], bx = vx[(n + 1) % npe], by = vy[(n + 1) % npe];
    const double cross = (bx - ax) * (py - ay) - (by - ay) * (px - ax);
    const bool on_seg = fmin(ax, bx) <= px && px <= fmax(ax, bx) && fmin(ay, by) <= py && py <= fmax(ay, by);
    on_boundary |= (fabs(cross) <= 1e-8) && on_seg;
    const bool y_cond = (ay <= py && by > py) || (by <= py && ay > py);
    if (y_cond && px < (bx - ax) * (py - ay) / (by - ay) + ax) ++crossings;
  }
  return on_boundary || (crossings & 1);
}

// operator.py:411-431: one Newton step from the first quadrature point, then N(xi) . u_e.  found < 0: NaN.
template <class El>
TATVA_D void interpolate_in_element(const double* __restrict__ coords, const int32_t* __restrict__ conn, int64_t found,
                                    const double* __restrict__ u, int nv, double px, double py,
                                    double* __restrict__ out_row) {
  constexpr int npe = El::npe;
  if (found < 0) {
    for (int c = 0; c < nv; ++c) out_row[c] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  int nd[npe];
  double X[npe][2];
#pragma unroll
  for (int n = 0; n < npe; ++n) {
    nd[n] = __ldg(conn + found * npe + n);
    X[n][0] = __ldg(coords + 2 * (int64_t)nd[n]);
    X[n][1] = __ldg(coords + 2 * (int64_t)nd[n] + 1);
  }
  double r0, s0, N0[npe], dN0[2][npe];
  first_quad_point(El{}, r0, s0);
  PlaneShape<El>::N(r0, s0, N0);
  PlaneShape<El>::dN(r0, s0, dN0);
  double x0 = 0.0, y0 = 0.0, a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;  // a_ij = d x_i / d xi_j
#pragma unroll
  for (int n = 0; n < npe; ++n) {
    x0 += N0[n] * X[n][0];
    y0 += N0[n] * X[n][1];
    a00 += dN0[0][n] * X[n][0];
    a01 += dN0[1][n] * X[n][0];
    a10 += dN0[0][n] * X[n][1];
    a11 += dN0[1][n] * X[n][1];
  }
  const double bx = px - x0, by = py - y0, det = a00 * a11 - a01 * a10;
  const double r = r0 + (a11 * bx - a01 * by) / det, s = s0 + (a00 * by - a10 * bx) / det;
  double N[npe];
  PlaneShape<El>::N(r, s, N);
  for (int c = 0; c < nv; ++c) {
    double t = 0.0;
#pragma unroll
    for (int n = 0; n < npe; ++n) t += N[n] * __ldg(u + (int64_t)nd[n] * nv + c);
    out_row[c] = t;
  }
}

// One thread per point, every element scanned: bounding boxes staged per CTA in shared memory, so the O(P E) scan
// reads each element's nodes once per CTA.  Used when the plan has no point grid.
template <class El>
__global__ void __launch_bounds__(128) k_interpolate(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                     int64_t E, const double* __restrict__ u, int nv,
                                                     const double* __restrict__ pts, int64_t P,
                                                     double* __restrict__ out, int32_t* __restrict__ elem) {
  constexpr int npe = El::npe, CH = 128;
  __shared__ double box[CH][4];
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < P;
  const double px = live ? pts[2 * p] : 0.0, py = live ? pts[2 * p + 1] : 0.0;
  int64_t found = -1;
  for (int64_t base = 0; base < E; base += CH) {
    __syncthreads();
    const int64_t e = base + threadIdx.x;
    if (e < E) {
      double lx = 1e308, ly = 1e308, hx = -1e308, hy = -1e308;
#pragma unroll
      for (int n = 0; n < npe; ++n) {
        const int64_t nd = __ldg(conn + e * npe + n);
        const double x = __ldg(coords + 2 * nd), y = __ldg(coords + 2 * nd + 1);
        lx = fmin(lx, x); hx = fmax(hx, x); ly = fmin(ly, y); hy = fmax(hy, y);
      }
      box[threadIdx.x][0] = lx; box[threadIdx.x][1] = hx; box[threadIdx.x][2] = ly; box[threadIdx.x][3] = hy;
    }
    __syncthreads();
    const int cnt = (int)((E - base < CH) ? (E - base) : CH);
    if (live && found < 0) {
      for (int k = 0; k < cnt; ++k) {
        if (!(px >= box[k][0] && px <= box[k][1] && py >= box[k][2] && py <= box[k][3])) continue;
        if (loop_contains<El>(coords, conn, base + k, px, py, false)) {
          found = base + k;
          break;
        }
      }
    }
  }
  if (!live) return;
  elem[p] = (int32_t)found;
  interpolate_in_element<El>(coords, conn, found, u, nv, px, py, out + p * nv);
}

// Same result through a uniform background grid (tatva_plan_set_point_grid): a bin lists, in ascending order, the
// elements whose bounding box overlaps it, so the first hit in the point's bin is the first hit overall.
struct PointGrid {
  int nx, ny;
  double lo[2], inv[2];
  const int32_t* ptr;
  const int32_t* elems;
};
TATVA_HD int grid_bin(double x, double lo, double inv, int n) {
  const double t = (x - lo) * inv;
  int b = t > 0.0 ? (t < (double)n ? (int)t : n - 1) : 0;
  return b;
}
template <class El>
__global__ void __launch_bounds__(128) k_interpolate_grid(const double* __restrict__ coords,
                                                          const int32_t* __restrict__ conn, PointGrid g,
                                                          const double* __restrict__ u, int nv,
                                                          const double* __restrict__ pts, int64_t P,
                                                          double* __restrict__ out, int32_t* __restrict__ elem) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double px = pts[2 * p], py = pts[2 * p + 1];
  int64_t found = -1;
  if (px == px && py == py) {  // NaN coordinates are in no element
    const int b = grid_bin(py, g.lo[1], g.inv[1], g.ny) * g.nx + grid_bin(px, g.lo[0], g.inv[0], g.nx);
    for (int k = __ldg(g.ptr + b), k1 = __ldg(g.ptr + b + 1); k < k1; ++k) {
      const int e = __ldg(g.elems + k);
      if (loop_contains<El>(coords, conn, e, px, py, true)) {
        found = e;
        break;
      }
    }
  }
  elem[p] = (int32_t)found;
  interpolate_in_element<El>(coords, conn, found, u, nv, px, py, out + p * nv);
}

template <class El, class Mat>
__global__ void __launch_bounds__(kBlock) k_csr(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                int64_t E, Mat mat, const double* __restrict__ u,
                                                const int32_t* __restrict__ indptr, const int32_t* __restrict__ pos,
                                                double* __restrict__ data) {
  constexpr int dpn = Mat::dpn;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int nd[El::npe];
  load_conn<El>(conn, e, nd);
  double X[El::npe][El::dim], U[El::npe][dpn];
  gather_rows(coords, nd, X);
  gather_rows(u, nd, U);
  int64_t row0[El::npe];  // indptr of the first DOF row of each node
#pragma unroll
  for (int a = 0; a < El::npe; ++a) row0[a] = (int64_t)nd[a] * dpn;

#pragma unroll 1
  for (int b = 0; b < El::npe; ++b) {
#pragma unroll 1
    for (int k = 0; k < dpn; ++k) {
      double col[El::npe][dpn];
#pragma unroll
      for (int a = 0; a < El::npe; ++a)
#pragma unroll
        for (int i = 0; i < dpn; ++i) col[a][i] = 0.0;
#pragma unroll 1
      for (int q = 0; q < El::num_q(); ++q) {
        double dNdX[El::dim][El::npe], N[El::npe];
        const double W = geometry<El>(q, X, dNdX) * El::weight(q);
        El::N(q, N);
        typename Mat::S s, ds, f;
        typename Mat::Cache cache;
        qp_state<El, Mat>(dNdX, N, U, s);
        mat.prepare(s, cache);
#pragma unroll
        for (int c = 0; c < dpn; ++c) {
#pragma unroll
          for (int j = 0; j < El::dim; ++j) {
            double t = 0.0;
#pragma unroll
            for (int n = 0; n < El::npe; ++n) t += (n == b && c == k) ? dNdX[j][n] : 0.0;
            ds.G[c][j] = t;
          }
          double t = 0.0;
#pragma unroll
          for (int n = 0; n < El::npe; ++n) t += (n == b && c == k) ? N[n] : 0.0;
          ds.val[c] = t;
        }
        mat.second(s, cache, ds, f);
#pragma unroll
        for (int a = 0; a < El::npe; ++a)
#pragma unroll
          for (int i = 0; i < dpn; ++i) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < El::dim; ++j) t += f.G[i][j] * dNdX[j][a];
            if (i >= Mat::val_lo) t += f.val[i] * N[a];
            col[a][i] += W * t;
          }
      }
#pragma unroll
      for (int a = 0; a < El::npe; ++a) {
        const int p = __ldg(pos + (e * El::npe + a) * El::npe + b) + k;
#pragma unroll
        for (int i = 0; i < dpn; ++i) atomicAdd(data + (int64_t)__ldg(indptr + row0[a] + i) + p, col[a][i]);
      }
    }
  }
}

// ---- CSR assembly with sector-grouped REDs --------------------------------------------------------
// Same arithmetic as k_csr, but the REDs of a warp are re-dealt through shared memory so that consecutive
// lanes add the dpn consecutive doubles of one (row, column-block): a RED group then touches one 32-byte
// sector instead of dpn.  The L2 atomic units work per sector (tools/micro/red_sector.cu: 212 G RED/s for
// scattered doubles, 420-490 G/s when 3-4 lanes share a sector), and assembly is bound by exactly that.
// Single-quadrature-point elements (the block of one column node fits in registers).
template <class El, class Mat, bool SYM>
__global__ void __launch_bounds__(kBlock) k_csr_grouped(const double* __restrict__ coords,
                                                        const int32_t* __restrict__ conn, int64_t E, Mat mat,
                                                        const double* __restrict__ u,
                                                        const int32_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ pos, double* __restrict__ data) {
  static_assert(El::max_nq == 1, "grouped assembly is implemented for single-point elements");
  constexpr int dpn = Mat::dpn, npe = El::npe, S = npe * dpn * dpn, NB = npe * dpn;
  constexpr int SP = S | 1, NBP = NB | 1;  // odd per-lane strides: conflict-free lane-strided stores
  extern __shared__ double sm_grp[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* sblk = sm_grp + (size_t)wib * (32 * SP + 16 * NBP);  // [32][SP] values, then [32][NBP] int32 bases
  int* sbase = reinterpret_cast<int*>(sblk + 32 * SP);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < E;
  int nd[npe];
  double dNdX[El::dim][npe], N[npe], W = 0.0;
  typename Mat::S s;
  typename Mat::Cache cache;
  if (valid) {
    load_conn<El>(conn, e, nd);
    double X[npe][El::dim], U[npe][dpn];
    gather_rows(coords, nd, X);
    gather_rows(u, nd, U);
    W = geometry<El>(0, X, dNdX) * El::weight(0);
    El::N(0, N);
    qp_state<El, Mat>(dNdX, N, U, s);
    mat.prepare(s, cache);
  }
#pragma unroll 1
  for (int b = 0; b < npe; ++b) {
    if (valid) {
#pragma unroll 1
      for (int k = 0; k < dpn; ++k) {
        typename Mat::S ds, f;
#pragma unroll
        for (int c = 0; c < dpn; ++c) {
#pragma unroll
          for (int j = 0; j < El::dim; ++j) {
            double t = 0.0;
#pragma unroll
            for (int n = 0; n < npe; ++n) t += (n == b && c == k) ? dNdX[j][n] : 0.0;
            ds.G[c][j] = t;
          }
          double t = 0.0;
#pragma unroll
          for (int n = 0; n < npe; ++n) t += (n == b && c == k) ? N[n] : 0.0;
          ds.val[c] = t;
        }
        mat.second(s, cache, ds, f);
#pragma unroll
        for (int a = 0; a < npe; ++a)
#pragma unroll
          for (int i = 0; i < dpn; ++i) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < El::dim; ++j) t += f.G[i][j] * dNdX[j][a];
            if (i >= Mat::val_lo) t += f.val[i] * N[a];
            sblk[lane * SP + (a * dpn + i) * dpn + k] = W * t;
          }
      }
      int nb_node = 0;  // nd[b] without dynamic indexing
#pragma unroll
      for (int n = 0; n < npe; ++n) nb_node = (n == b) ? nd[n] : nb_node;
#pragma unroll
      for (int a = 0; a < npe; ++a) {
        // SYM: only the upper triangle (row node <= column node) is accumulated; k_csr_mirror fills the rest
        const bool keep = !SYM || nd[a] <= nb_node;
        const int p = __ldg(pos + (e * npe + a) * npe + b);
#pragma unroll
        for (int i = 0; i < dpn; ++i) sbase[lane * NBP + a * dpn + i] = keep ? __ldg(indptr + (int64_t)nd[a] * dpn + i) + p : -1;
      }
    } else {
#pragma unroll
      for (int r = 0; r < NB; ++r) sbase[lane * NBP + r] = -1;
    }
    __syncwarp();
    for (int t = lane; t < 32 * S; t += 32) {
      const int j = t / S, r = t - j * S;
      const int base = sbase[j * NBP + r / dpn];
#ifdef TATVA_CSR_EXPERIMENT  // timing experiment: issue only every TATVA_CSR_EXPERIMENT-th RED group (wrong result)
      if (base >= 0 && ((t / dpn) % TATVA_CSR_EXPERIMENT) == 0) atomicAdd(data + (int64_t)base + (r % dpn), sblk[j * SP + r]);
#else
      if (base >= 0) atomicAdd(data + (int64_t)base + (r % dpn), sblk[j * SP + r]);
#endif
    }
    __syncwarp();
  }
}

// ---- tiled CSR assembly with on-chip combination of duplicate blocks ----------------------------------------
// ncu of k_csr_grouped at config 2 (profiles/r02_csr_ncu.md): 8300 instructions per element (12 generic tangent
// evaluations with select-built unit directions, a 144-entry staging / re-deal loop), LSU wavefronts 79 %, and 144 M
// REDs of which most are duplicates: the 16 (row node, column node) blocks of a tet are shared with its neighbours.
// Here a CTA owns a tile of kTile consecutive (locality-sorted) elements:
//   phase 1  one thread per element: geometry and the law's rank structure — for the isotropic laws below the tangent
//            block of nodes (a, b) is   K_ab = w1 (dNa . dNb) I + w2 g_b g_a^T + w3 g_a g_b^T   with per-element
//            vectors dN_a, g_a and three scalars (27 doubles per tet) — parked in shared memory;
//   phase 2  one thread per distinct block of the tile: sums the block's contributors from shared memory (plan-time
//            schedule, tatva_host_csr_tile_schedule); the sums are re-dealt through a chunk buffer so that ONE RED per entry
//            is issued with dpn consecutive lanes adding dpn consecutive doubles (one 32-byte sector).
// ~3 x fewer REDs and ~4 x fewer instructions than k_csr_grouped.  Laws: neo-Hookean (g = F^-T dN, w2 = W (mu -
// lambda lnJ), w3 = W lambda) and linear elasticity (g = dN, w2 = W mu, w3 = W lambda); w1 = W mu for both.
constexpr int kTile = 128;

template <class Mat>
struct RankLaw;
template <int DIM>
struct RankLaw<LinearElastic<DIM>> {
  template <int NPE>
  TATVA_HD static void eval(const LinearElastic<DIM>& m, double W, const double (&dN)[DIM][NPE], const double (&)[NPE][DIM],
                           double (&g)[DIM][NPE], double (&w)[3]) {
#pragma unroll
    for (int j = 0; j < DIM; ++j)
#pragma unroll
      for (int n = 0; n < NPE; ++n) g[j][n] = dN[j][n];
    w[0] = W * m.mu;
    w[1] = W * m.mu;
    w[2] = W * m.lmbda;
  }
};
template <>
struct RankLaw<NeoHookean> {
  template <int NPE>
  TATVA_HD static void eval(const NeoHookean& m, double W, const double (&dN)[3][NPE], const double (&U)[NPE][3], double (&g)[3][NPE],
                           double (&w)[3]) {
    double F[3][3], Fi[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double t = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int n = 0; n < NPE; ++n) t = fma(U[n][i], dN[j][n], t);
        F[i][j] = t;
      }
    const double lnJ = log(det_inv(F, Fi));
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int n = 0; n < NPE; ++n) g[c][n] = Fi[0][c] * dN[0][n] + Fi[1][c] * dN[1][n] + Fi[2][c] * dN[2][n];  // F^-T dN
    w[0] = W * m.mu;
    w[1] = W * (m.mu - m.lmbda * lnJ);
    w[2] = W * m.lmbda;
  }
};

template <class El, class Mat>
__global__ void __launch_bounds__(kTile) k_csr_tiled(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                     int64_t E, Mat mat, const double* __restrict__ u,
                                                     const int32_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_base,
                                                     const int32_t* __restrict__ blk_rowlen, const int32_t* __restrict__ blk_base_t,
                                                     const int32_t* __restrict__ blk_rowlen_t, const int32_t* __restrict__ con_ptr,
                                                     const uint32_t* __restrict__ con, double* __restrict__ data) {
  static_assert(El::max_nq == 1 && Mat::dpn == El::dim, "constant-gradient elements, one DOF per direction");
  constexpr int D = El::dim, NPE = El::npe, NV = 2 * NPE * D + 3, NVP = NV | 1;
  // [kTile][NVP], element-major with an odd stride: phase 1 (lane = element) and phase 2 (lanes read different fields of
  // the same or of neighbouring elements) are both free of bank conflicts; a field-major layout puts every field of one
  // element on the same bank, and the lanes of phase 2 mostly visit the same few elements (28 M conflicts, ncu).
  extern __shared__ double sm_tile[];
  const int64_t e = (int64_t)blockIdx.x * kTile + threadIdx.x;
  if (e < E) {
    int nd[NPE];
    load_conn<El>(conn, e, nd);
    double X[NPE][D], U[NPE][D], dN[D][NPE], g[D][NPE], w[3];
    gather_rows(coords, nd, X);
    gather_rows(u, nd, U);
    const double W = geometry<El>(0, X, dN) * El::weight(0);
    RankLaw<Mat>::template eval<NPE>(mat, W, dN, U, g, w);
    double* row = sm_tile + threadIdx.x * NVP;
#pragma unroll
    for (int a = 0; a < NPE; ++a)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        row[a * D + j] = dN[j][a];
        row[NPE * D + a * D + j] = g[j][a];
      }
#pragma unroll
    for (int k = 0; k < 3; ++k) row[2 * NPE * D + k] = w[k];
  }
  __syncthreads();
  // phase 2: one thread per distinct block sums ALL D x D entries over the block's contributors (each contributor's
  // vectors are read once), parks them in a chunk buffer, and the CTA re-deals the chunk so that D consecutive lanes add
  // D consecutive doubles of one CSR row (one 32-byte sector per RED group).
  constexpr int DD = D * D;
  double* buf = sm_tile + NVP * kTile;                      // [kTile][DD], lane-major with an odd stride: conflict-free
  int* sbase = reinterpret_cast<int*>(buf + DD * kTile);    // [kTile] block base position, -1 = none
  int* srl = sbase + kTile;                                 // [kTile] row length of the block's row node
  int* sbase_t = srl + kTile;                               // the same for the mirror block (b, a): K_ba = K_ab^T
  int* srl_t = sbase_t + kTile;
  const int b0 = __ldg(blk_ptr + blockIdx.x), nb = __ldg(blk_ptr + blockIdx.x + 1) - b0;
  // every WARP walks its own chunks of 32 blocks (the schedule lists the blocks by decreasing contributor count, so the
  // lanes of a warp loop about equally long) and re-deals them with warp-level synchronisation only
  const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
  double* wbuf = buf + wbase * DD;
  for (int chunk = wbase; chunk < nb; chunk += kTile) {
    const int bl = chunk + lane;
    double acc[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) acc[i][k] = 0.0;
    if (bl < nb) {
      const int d = b0 + bl;
      const int c0 = __ldg(con_ptr + d), c1 = __ldg(con_ptr + d + 1);
      // the block's CSR positions are fetched now and consumed after the contributor loop (latency hidden behind it)
      const int pb0 = __ldg(blk_base + d), pr0 = __ldg(blk_rowlen + d), pb1 = __ldg(blk_base_t + d), pr1 = __ldg(blk_rowlen_t + d);
      uint32_t src_next = __ldg(con + c0);
      for (int c = c0; c < c1; ++c) {
        const uint32_t src = src_next;
        if (c + 1 < c1) src_next = __ldg(con + c + 1);
        const double* row = sm_tile + (src >> 8) * NVP;
        const double* pa = row + ((src >> 4) & 15) * D;
        const double* pb = row + (src & 15) * D;
        double dNa[D], dNb[D], ga[D], gb[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
          dNa[j] = pa[j];
          dNb[j] = pb[j];
          ga[j] = pa[NPE * D + j];
          gb[j] = pb[NPE * D + j];
        }
        const double w1 = row[2 * NPE * D], w2 = row[2 * NPE * D + 1], w3 = row[2 * NPE * D + 2];
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) dot = fma(dNa[j], dNb[j], dot);
        const double a1 = w1 * dot;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double a2 = w2 * ga[k], a3 = w3 * gb[k];
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i][k] += fma(a2, gb[i], fma(a3, ga[i], (i == k) ? a1 : 0.0));
        }
      }
      sbase[threadIdx.x] = pb0;
      srl[threadIdx.x] = pr0;
      sbase_t[threadIdx.x] = pb1;
      srl_t[threadIdx.x] = pr1;
    } else {
      sbase[threadIdx.x] = -1;
      sbase_t[threadIdx.x] = -1;
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) wbuf[lane * DD + i * D + k] = acc[i][k];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < DD; ++j) {
      const int q = lane + j * 32, owner = q / DD, r = q - owner * DD, hi = r / D, lo = r - hi * D;
      // block (a, b): entry (i, k) = (hi, lo); mirror block (b, a): entry (k, i) = (hi, lo) holds acc[i = lo][k = hi]
      const int base = sbase[wbase + owner], base_t = sbase_t[wbase + owner];
      if (base >= 0) atomicAdd(data + (int64_t)base + (int64_t)hi * srl[wbase + owner] + lo, wbuf[q]);
      if (base_t >= 0) atomicAdd(data + (int64_t)base_t + (int64_t)hi * srl_t[wbase + owner] + lo, wbuf[owner * DD + lo * D + hi]);
    }
    __syncwarp();
  }
}

// Hessian diagonal through the laws' rank structure: K_aa(i, i) = w1 |dN_a|^2 + (w2 + w3) g_a[i]^2 per point — the
// geometry, F^-1 and one logarithm per point, then 3 + 3*dim fused operations per (node, point) instead of one full
// second-variation evaluation per (node, component) as in k_hessian_diag (Hex8 x neo-Hooke: 24 tangent evaluations per
// point).  Same element-per-thread gather and sector-grouped scatter.
template <class T>
struct has_rank_law : std::false_type {};
template <int DIM>
struct has_rank_law<LinearElastic<DIM>> : std::true_type {};
template <>
struct has_rank_law<NeoHookean> : std::true_type {};

template <class El, class Mat>
__global__ void __launch_bounds__(kBlock) k_hessian_diag_rank(const double* __restrict__ coords,
                                                              const int32_t* __restrict__ conn, int64_t E, Mat mat,
                                                              const double* __restrict__ u, double* __restrict__ diag) {
  static_assert(Mat::dpn == El::dim, "one DOF per direction");
  constexpr int D = El::dim, NPE = El::npe;
  extern __shared__ double sm_diag[];
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nd[NPE];
  double Y[NPE][D];
#pragma unroll
  for (int n = 0; n < NPE; ++n) {
    nd[n] = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) Y[n][c] = 0.0;
  }
  if (e < E) {
    load_conn<El>(conn, e, nd);
    double X[NPE][D], U[NPE][D];
    gather_rows(coords, nd, X);
    gather_rows(u, nd, U);
#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dN[D][NPE], g[D][NPE], w[3];
      const double W = geometry<El>(q, X, dN) * El::weight(q);
      RankLaw<Mat>::template eval<NPE>(mat, W, dN, U, g, w);
      const double w23 = w[1] + w[2];
#pragma unroll
      for (int a = 0; a < NPE; ++a) {
        double nn = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) nn = fma(dN[j][a], dN[j][a], nn);
        const double a1 = w[0] * nn;
#pragma unroll
        for (int i = 0; i < D; ++i) Y[a][i] += fma(w23 * g[i][a], g[i][a], a1);
      }
    }
  }
  double* wsm = sm_diag + (size_t)(threadIdx.x >> 5) * grouped_scatter_words<NPE, D>();
  grouped_scatter<NPE, D>(diag, nd, Y, e < E, wsm);
}

// Lower triangle from the upper one (the energy Hessian is symmetric): one warp per node row a, one lane per
// block (a, b) with b < a:  K[(a,i),(b,k)] = K[(b,k),(a,i)].  The position of a in row b is found by binary search.
__global__ void __launch_bounds__(128) k_csr_mirror(int64_t n_nodes, int dpn, const int32_t* __restrict__ indptr,
                                                    const int32_t* __restrict__ indices, double* __restrict__ data) {
  const int lane = threadIdx.x & 31;
  const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_nodes) return;
  const int r0 = __ldg(indptr + a * dpn);
  const int nb = (__ldg(indptr + a * dpn + 1) - r0) / dpn;
  for (int j = lane; j < nb; j += 32) {
    const int b = __ldg(indices + r0 + j * dpn) / dpn;
    if (b >= a) break;  // columns are sorted: the rest of the row is the upper triangle
    const int c0 = __ldg(indptr + (int64_t)b * dpn);
    int lo = 0, hi = (__ldg(indptr + (int64_t)b * dpn + 1) - c0) / dpn;
    while (lo < hi) {  // first block of row b whose node is >= a
      const int mid = (lo + hi) >> 1;
      if (__ldg(indices + c0 + mid * dpn) / dpn < (int)a) lo = mid + 1; else hi = mid;
    }
    for (int i = 0; i < dpn; ++i) {
      double* dst = data + (int64_t)__ldg(indptr + a * dpn + i) + j * dpn;
      for (int k = 0; k < dpn; ++k) dst[k] = data[(int64_t)__ldg(indptr + (int64_t)b * dpn + k) + lo * dpn + i];
    }
  }
}

// ---- CSR assembly by rows (no atomics, deterministic) ---------------------------------------------
// One warp owns the dpn CSR rows of one node a.  Phase 1: lane l takes the l-th element incident on a
// and computes the three columns (a,i) of its stiffness — by symmetry of the energy Hessian these are
// the rows (a,i) — into a shared-memory slab.  Phase 2: lane m takes the m-th block entry (a, b_m) of the
// row, sums the slabs of the elements that contain b_m in a fixed order and stores the dpn x dpn block
// with plain, row-contiguous stores.  Every stored entry of a mesh pattern is written exactly once.
// Single-quadrature-point elements only (Tri3, Tet4); others keep the atomic kernel.

template <class El, class Mat>
__global__ void __launch_bounds__(128) k_csr_rows(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                  int64_t n_nodes, Mat mat, const double* __restrict__ u,
                                                  const int32_t* __restrict__ indptr,
                                                  const int32_t* __restrict__ indices,
                                                  const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e,
                                                  double* __restrict__ data) {
  static_assert(El::max_nq == 1, "row-wise assembly is implemented for single-point elements");
  constexpr int dpn = Mat::dpn, npe = El::npe, SLAB = dpn * npe * dpn;
  extern __shared__ double sm_rows[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* slab = sm_rows + (size_t)wib * 32 * (SLAB + npe / 2 + 1);  // per warp: 32 slabs then 32 x npe node ids
  int* sconn = reinterpret_cast<int*>(slab + 32 * SLAB);
  const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (a >= n_nodes) return;
  const int ebeg = __ldg(n2e_ptr + a), eend = __ldg(n2e_ptr + a + 1);
  const int64_t row = a * dpn;
  const int r0 = __ldg(indptr + row);
  const int nb = (__ldg(indptr + row + 1) - r0) / dpn;

  for (int mb = 0; mb < nb; mb += 32) {  // block entries of the row, 32 at a time (one pass for usual meshes)
    const int m = mb + lane;
    const int bnode = (m < nb) ? __ldg(indices + r0 + m * dpn) / dpn : -1;
    double acc[dpn][dpn];
#pragma unroll
    for (int i = 0; i < dpn; ++i)
#pragma unroll
      for (int k = 0; k < dpn; ++k) acc[i][k] = 0.0;

    for (int eb = ebeg; eb < eend; eb += 32) {
      const int cnt = min(32, eend - eb);
      __syncwarp();
      if (lane < cnt) {
        const int64_t e = __ldg(n2e + eb + lane);
        int nd[npe];
        load_conn<El>(conn, e, nd);
        double X[npe][El::dim], U[npe][dpn];
        gather_rows(coords, nd, X);
        gather_rows(u, nd, U);
        double dNdX[El::dim][npe], N[npe];
        const double W = geometry<El>(0, X, dNdX) * El::weight(0);
        El::N(0, N);
        typename Mat::S s, ds, f;
        typename Mat::Cache cache;
        qp_state<El, Mat>(dNdX, N, U, s);
        mat.prepare(s, cache);
        double ga[El::dim], Na = 0.0;  // shape gradient / value of the row node inside this element
#pragma unroll
        for (int j = 0; j < El::dim; ++j) ga[j] = 0.0;
#pragma unroll
        for (int n = 0; n < npe; ++n) {
          const bool hit = (nd[n] == (int)a);
#pragma unroll
          for (int j = 0; j < El::dim; ++j) ga[j] = hit ? dNdX[j][n] : ga[j];
          Na = hit ? N[n] : Na;
          sconn[lane * npe + n] = nd[n];
        }
#pragma unroll
        for (int i = 0; i < dpn; ++i) {
#pragma unroll
          for (int c = 0; c < dpn; ++c) {
#pragma unroll
            for (int j = 0; j < El::dim; ++j) ds.G[c][j] = (c == i) ? ga[j] : 0.0;
            ds.val[c] = (c == i) ? Na : 0.0;
          }
          mat.second(s, cache, ds, f);
#pragma unroll
          for (int b = 0; b < npe; ++b)
#pragma unroll
            for (int k = 0; k < dpn; ++k) {
              double t = 0.0;
#pragma unroll
              for (int j = 0; j < El::dim; ++j) t += f.G[k][j] * dNdX[j][b];
              if (k >= Mat::val_lo) t += f.val[k] * N[b];
              slab[lane * SLAB + (i * npe + b) * dpn + k] = W * t;  // K_e[(b,k),(a,i)] = K_e[(a,i),(b,k)]
            }
        }
      }
      __syncwarp();
      if (m < nb) {
        for (int l = 0; l < cnt; ++l) {
#pragma unroll
          for (int b = 0; b < npe; ++b) {
            if (sconn[l * npe + b] == bnode) {
#pragma unroll
              for (int i = 0; i < dpn; ++i)
#pragma unroll
                for (int k = 0; k < dpn; ++k) acc[i][k] += slab[l * SLAB + (i * npe + b) * dpn + k];
            }
          }
        }
      }
    }
    if (m < nb) {
#pragma unroll
      for (int i = 0; i < dpn; ++i) {
        double* dst = data + (int64_t)__ldg(indptr + row + i) + (int64_t)m * dpn;
#pragma unroll
        for (int k = 0; k < dpn; ++k) dst[k] = acc[i][k];
      }
    }
  }
}

// ---- halo pack / unpack -----------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_pack(const double* __restrict__ src, const int64_t* __restrict__ idx,
                                              int64_t n, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __ldg(src + __ldg(idx + i));
}
__global__ void __launch_bounds__(256) k_unpack_set(const double* __restrict__ src, const int64_t* __restrict__ idx,
                                                    int64_t n, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[__ldg(idx + i)] = __ldg(src + i);
}
__global__ void __launch_bounds__(256) k_unpack_add(const double* __restrict__ src, const int64_t* __restrict__ idx,
                                                    int64_t n, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(dst + __ldg(idx + i), __ldg(src + i));
}

// ---- peer-memory halo over NVLink (no NCCL, no pack buffers) ------------------------------------------
// Local vectors live in symmetric (peer-mapped) memory; `peers[r]` is rank r's base pointer of the same vector.
// pull:  x_local[first_ghost + g]            = peers[owner[g]][owner_idx[g]]      (ghost fill, mpi.py:372-409)
// push:  peers[owner[g]][owner_idx[g]]      += y_local[first_ghost + g]           (ghost contributions to their
//        owners, mpi.py:479-516; RED over NVLink, performed at the owner's L2)
__global__ void __launch_bounds__(256) k_peer_pull(double* __restrict__ x_local, int64_t first_ghost, int64_t n_ghost,
                                                   const uint64_t* __restrict__ peers, const int32_t* __restrict__ owner,
                                                   const int64_t* __restrict__ owner_idx) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_ghost) return;
  const double* src = reinterpret_cast<const double*>(peers[owner[g]]);
  x_local[first_ghost + g] = __ldcv(src + owner_idx[g]);  // peer memory is not cached in the local L2; skip L1 too
}
__global__ void __launch_bounds__(256) k_peer_push_add(const double* __restrict__ y_local, int64_t first_ghost,
                                                       int64_t n_ghost, const uint64_t* __restrict__ peers,
                                                       const int32_t* __restrict__ owner,
                                                       const int64_t* __restrict__ owner_idx) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_ghost) return;
  double* dst = reinterpret_cast<double*>(peers[owner[g]]);
  atomicAdd_system(dst + owner_idx[g], y_local[first_ghost + g]);
}

// ---- lifter: reduced <-> full maps --------------------------------------------------------------
// lift:   out[i] = src[i] >= 0 ? u_red[src[i]] : (src[i] == -1 ? base[i] : consts[-(src[i] + 2)])
// adjoint: r_red[j] = sum of r_full over the full DOFs that read reduced DOF j (fixed order)
__global__ void __launch_bounds__(256) k_lift(const double* __restrict__ u_red, const int64_t* __restrict__ src,
                                              const double* __restrict__ consts, const double* __restrict__ base,
                                              int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // s >= 0: reduced index; s <= -2^40: entry -(s + 2^40) of the base vector (a Periodic slave may read its master's);
  // otherwise consts[-(s + 2)]
  constexpr int64_t kBaseOff = (int64_t)1 << 40;
  const int64_t s = __ldg(src + i);
  out[i] = s >= 0 ? __ldg(u_red + s) : (s <= -kBaseOff ? (base ? __ldg(base - (s + kBaseOff)) : 0.0) : __ldg(consts - (s + 2)));
}
__global__ void __launch_bounds__(256) k_reduce_adjoint(const double* __restrict__ r_full,
                                                        const int64_t* __restrict__ ptr,
                                                        const int64_t* __restrict__ list, int64_t n_red,
                                                        double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_red) return;
  double s = 0.0;
  for (int64_t k = __ldg(ptr + j); k < __ldg(ptr + j + 1); ++k) s += __ldg(r_full + __ldg(list + k));
  out[j] = s;
}

// ---- device-resident conjugate gradient: fused vector kernels, scalars never leave the device --------
// s[0] = r.r (current), s[1] = p.Ap, s[2] = r.r (next); partial sums in `part` (one slot per CTA, two lanes).
constexpr int kCgBlocks = 1184;  // 148 SMs x 8 CTAs

__global__ void __launch_bounds__(256) k_cg_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                                double* __restrict__ part) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += __ldg(a + i) * __ldg(b + i);
  s = block_sum(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// out[slot] = sum(part[0..nblocks))  (single CTA, fixed order)
__global__ void __launch_bounds__(256) k_cg_finish(const double* __restrict__ part, int nblocks, double* __restrict__ s,
                                                   int slot) {
  double t = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) t += part[b];
  t = block_sum(t);
  if (threadIdx.x == 0) s[slot] = t;
}
// alpha = s[0] / s[1];  x += alpha p;  r -= alpha Ap;  partial of r.r
__global__ void __launch_bounds__(256) k_cg_update(double* __restrict__ x, double* __restrict__ r,
                                                   const double* __restrict__ p, const double* __restrict__ Ap,
                                                   int64_t n, const double* __restrict__ s, double* __restrict__ part) {
  const double alpha = s[0] / s[1];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, __ldg(p + i), x[i]);
    const double ri = fma(-alpha, __ldg(Ap + i), r[i]);
    r[i] = ri;
    acc = fma(ri, ri, acc);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// k_cg_update / k_pcg_update on FULL-size vectors with Dirichlet rows masked out: map[i] < 0 marks a Fixed DOF (its
// residual entry stays 0, so p and x stay 0 there).  minv may be NULL.
__global__ void __launch_bounds__(256) k_cg_update_masked(double* __restrict__ x, double* __restrict__ r,
                                                          const double* __restrict__ p, const double* __restrict__ Ap,
                                                          const double* __restrict__ minv, const int32_t* __restrict__ map,
                                                          int64_t n, const double* __restrict__ s, double* __restrict__ part) {
  const double alpha = s[0] / s[1];
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, __ldg(p + i), x[i]);
    double ri = fma(-alpha, __ldg(Ap + i), r[i]);
    if (__ldg(map + i) < 0) ri = 0.0;
    r[i] = ri;
    rr = fma(ri, ri, rr);
    if (minv) rz = fma(ri * __ldg(minv + i), ri, rz);
  }
  rr = block_sum(rr);
  if (minv) {
    __syncthreads();
    rz = block_sum(rz);
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x] = minv ? rz : rr;
    if (minv) part[gridDim.x + blockIdx.x] = rr;
  }
}
// beta = s[2] / s[0];  p = r + beta p;  then s[0] <- s[2] (done by CTA 0 after its own work; readers use the
// value loaded at entry)
__global__ void __launch_bounds__(256) k_cg_direction(double* __restrict__ p, const double* __restrict__ r, int64_t n,
                                                      double* __restrict__ s) {
  const double beta = s[2] / s[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], __ldg(r + i));
}
__global__ void k_cg_roll(double* __restrict__ s) { s[0] = s[2]; }
// s[slot] = sum(part[0..nblocks)) in a fixed order (any nblocks, one CTA of 1024 threads); roll: s[0] <- s[2] first
__global__ void __launch_bounds__(1024) k_cg_finish_roll(const double* __restrict__ part, int nblocks, double* __restrict__ s,
                                                         int slot, int roll) {
  double t = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) t += part[b];
  t = block_sum(t);
  if (threadIdx.x == 0) {
    if (roll) s[0] = s[2];
    s[slot] = t;
  }
}
// beta = s[2] / s[0];  p = r + beta p  walking the vectors BACKWARDS (the tail of r, written last by k_cg_update, is
// still in L2), and zero = 0 written alongside (the next operator application accumulates into it: no memset launch).
// s[0] is NOT rolled here: the final-sum kernel of the next p.Ap does it (k_cg_finish_roll).
__global__ void __launch_bounds__(256) k_cg_direction_zero(double* __restrict__ p, const double* __restrict__ r,
                                                           const double* __restrict__ minv, int64_t n,
                                                           const double* __restrict__ s, double* __restrict__ zero) {
  const double beta = s[2] / s[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int64_t i = n - 1 - j;
    const double z = minv ? __ldg(r + i) * __ldg(minv + i) : __ldg(r + i);
    p[i] = fma(beta, p[i], z);
    if (zero) zero[i] = 0.0;
  }
}

// Jacobi-preconditioned variants: z = minv * r is never stored.  s[0] = r.z (current), s[2] = r.z (next), s[4] = r.r.
__global__ void __launch_bounds__(256) k_pcg_update(double* __restrict__ x, double* __restrict__ r,
                                                    const double* __restrict__ p, const double* __restrict__ Ap,
                                                    const double* __restrict__ minv, int64_t n,
                                                    const double* __restrict__ s, double* __restrict__ part) {
  const double alpha = s[0] / s[1];
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, __ldg(p + i), x[i]);
    const double ri = fma(-alpha, __ldg(Ap + i), r[i]);
    r[i] = ri;
    rr = fma(ri, ri, rr);
    rz = fma(ri * __ldg(minv + i), ri, rz);
  }
  rz = block_sum(rz);
  __syncthreads();
  rr = block_sum(rr);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = rz;
    part[gridDim.x + blockIdx.x] = rr;
  }
}
__global__ void __launch_bounds__(256) k_pcg_direction(double* __restrict__ p, const double* __restrict__ r,
                                                       const double* __restrict__ minv, int64_t n,
                                                       const double* __restrict__ s) {
  const double beta = s[2] / s[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], __ldg(r + i) * __ldg(minv + i));
}
// out[i] = 1 / d[i] (Jacobi), or 1 where d[i] is not a positive finite number
__global__ void __launch_bounds__(256) k_safe_reciprocal(const double* __restrict__ d, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = d[i];
    out[i] = (v > 0.0 && v < 1.79e308) ? 1.0 / v : 1.0;
  }
}
// p = minv * r ; partial of r.(minv r)
__global__ void __launch_bounds__(256) k_pcg_start(double* __restrict__ p, const double* __restrict__ r,
                                                   const double* __restrict__ minv, int64_t n, double* __restrict__ part) {
  double rz = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = __ldg(r + i), zi = ri * __ldg(minv + i);
    p[i] = zi;
    rz = fma(ri, zi, rz);
  }
  rz = block_sum(rz);
  if (threadIdx.x == 0) part[blockIdx.x] = rz;
}

// ---- FP64 FMA peak microbenchmark ---------------------------------------------------------------

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace tatva

// =================================================================================================
// C ABI
// =================================================================================================
using namespace tatva;

// staged kernels may need more than the default 48 KB of dynamic shared memory (opt-in, once per kernel)
constexpr size_t kStageMax = 160 * 1024;
template <class K>
static int allow_big_smem(K kernel, SmemOptIn& done) {
  return opt_in_smem(kernel, kStageMax, done);
}
#define STAGED_OPT_IN(KERNEL)                                   \
  {                                                             \
    static SmemOptIn done_[8];                                  \
    int rc_ = TATVA_OK;                                         \
    DISPATCH_ELEMENT(p, (rc_ = allow_big_smem(KERNEL<El>, done_[El::kind]))); \
    if (rc_ != TATVA_OK) return rc_;                            \
  }

#define DISPATCH_ELEMENT(p, CALL)                 \
  switch ((p)->element) {                         \
    case TATVA_TRI3: { using El = Tri3; CALL; } break; \
    case TATVA_TET4: { using El = Tet4; CALL; } break; \
    case TATVA_HEX8: { using El = Hex8; CALL; } break; \
    case TATVA_QUAD4: { using El = Quad4; CALL; } break; \
    case TATVA_TRI6: { using El = Tri6; CALL; } break; \
    case TATVA_QUAD8: { using El = Quad8; CALL; } break; \
    case TATVA_LINE2: { using El = Line2; CALL; } break; \
    case TATVA_LINE3: { using El = Line3; CALL; } break; \
    default: return TATVA_E_INVALID;              \
  }

// ---- user-supplied quadrature rules -------------------------------------------------------------------------
#define DISPATCH_CUSTOM(p, CALL)                 \
  switch ((p)->element) {                         \
    case TATVA_TRI3: { using El = Custom<Tri3>; CALL; } break; \
    case TATVA_TET4: { using El = Custom<Tet4>; CALL; } break; \
    case TATVA_HEX8: { using El = Custom<Hex8>; CALL; } break; \
    case TATVA_QUAD4: { using El = Custom<Quad4>; CALL; } break; \
    case TATVA_TRI6: { using El = Custom<Tri6>; CALL; } break; \
    case TATVA_QUAD8: { using El = Custom<Quad8>; CALL; } break; \
    case TATVA_LINE2: { using El = Custom<Line2>; CALL; } break; \
    case TATVA_LINE3: { using El = Custom<Line3>; CALL; } break; \
    default: return TATVA_E_INVALID;              \
  }

// Stream-ordered: the copy precedes the kernel that reads it.  Two plans with DIFFERENT custom rules must not launch
// concurrently on different streams (one constant-memory slot per process and device) — documented in the header.
static cudaError_t install_rule(const tatva_plan* p, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(tatva::c_rule, &p->rule, sizeof(QuadRule), 0, cudaMemcpyHostToDevice, st);
}

extern "C" {

const char* tatva_error_string(int code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case TATVA_OK: return "ok";
    case TATVA_E_INVALID: return "invalid argument";
    case TATVA_E_UNSUPPORTED: return "unsupported element/material combination";
    case TATVA_E_NOMEM: return "out of memory";
    case TATVA_E_NODEVICE: return "no CUDA device";
    default: return "unknown error";
  }
}

int tatva_abi_version(void) { return TATVA_B200_ABI_VERSION; }

int tatva_device_count(int* count) {
  if (!count) return TATVA_E_INVALID;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return (int)e;
  }
  return TATVA_OK;
}

int tatva_plan_create(tatva_plan_t** out, int element, int64_t n_nodes, int64_t n_elems, const double* d_coords,
                      const int32_t* d_conn, int flags, tatva_stream_t stream) {
  if (!out || !d_coords || !d_conn || n_nodes <= 0 || n_elems <= 0) return TATVA_E_INVALID;
  tatva_plan* p = new (std::nothrow) tatva_plan();
  if (!p) return TATVA_E_NOMEM;
  p->element = element;
  switch (element) {
    case TATVA_TRI3: p->dim = Tri3::dim; p->gdim = Tri3::gdim; p->npe = Tri3::npe; p->nq = Tri3::nq; break;
    case TATVA_TET4: p->dim = Tet4::dim; p->gdim = Tet4::gdim; p->npe = Tet4::npe; p->nq = Tet4::nq; break;
    case TATVA_HEX8: p->dim = Hex8::dim; p->gdim = Hex8::gdim; p->npe = Hex8::npe; p->nq = Hex8::nq; break;
    case TATVA_QUAD4: p->dim = Quad4::dim; p->gdim = Quad4::gdim; p->npe = Quad4::npe; p->nq = Quad4::nq; break;
    case TATVA_TRI6: p->dim = Tri6::dim; p->gdim = Tri6::gdim; p->npe = Tri6::npe; p->nq = Tri6::nq; break;
    case TATVA_QUAD8: p->dim = Quad8::dim; p->gdim = Quad8::gdim; p->npe = Quad8::npe; p->nq = Quad8::nq; break;
    case TATVA_LINE2: p->dim = Line2::dim; p->gdim = Line2::gdim; p->npe = Line2::npe; p->nq = Line2::nq; break;
    case TATVA_LINE3: p->dim = Line3::dim; p->gdim = Line3::gdim; p->npe = Line3::npe; p->nq = Line3::nq; break;
    default: delete p; return TATVA_E_INVALID;
  }
  p->n_nodes = n_nodes;
  p->n_elems = n_elems;
  p->coords = d_coords;
  p->conn = d_conn;
  p->flags = flags;
  p->variant = TATVA_VARIANT_DEFAULT;
  p->zero_output = 1;
  p->pss = 0;
  p->weights = nullptr;
  p->tile_ptr = nullptr;
  p->tile_nodes = nullptr;
  p->tile_conn = nullptr;
  p->tile_max_unique = 0;
  p->ws_warp_nodes = nullptr;
  p->geo = nullptr;
  p->geo_stride = 0;
  p->custom = 0;
  p->scratch_len = (int64_t)grid_for(n_elems) * (kBlock / 32) > 1024 * 64 ? (int64_t)grid_for(n_elems) * (kBlock / 32) : 1024 * 64;
  cudaError_t e = cudaMalloc(&p->scratch, sizeof(double) * p->scratch_len);
  if (e != cudaSuccess) { delete p; return (int)e; }
  if (flags & TATVA_PLAN_CACHE_WEIGHTS) {
    e = cudaMalloc(&p->weights, sizeof(double) * n_elems * p->nq);
    if (e != cudaSuccess) { cudaFree(p->scratch); delete p; return (int)e; }
    double* w = p->weights;
    p->weights = nullptr;  // compute (not read) the weights on the first pass
    int rc = tatva_op_integration_weights(p, w, stream);
    p->weights = w;
    if (rc != 0) { tatva_plan_destroy(p); return rc; }
  }
  *out = p;
  return TATVA_OK;
}

int tatva_plan_destroy(tatva_plan_t* p) {
  if (!p) return TATVA_OK;
  if (p->scratch) cudaFree(p->scratch);
  if (p->weights) cudaFree(p->weights);
  if (p->geo) cudaFree(p->geo);
  delete p;
  return TATVA_OK;
}

int tatva_plan_info(const tatva_plan_t* p, int* element, int* dim, int* npe, int* nq, int64_t* n_nodes,
                    int64_t* n_elems) {
  if (!p) return TATVA_E_INVALID;
  if (element) *element = p->element;
  if (dim) *dim = p->dim;
  if (npe) *npe = p->npe;
  if (nq) *nq = p->nq;
  if (n_nodes) *n_nodes = p->n_nodes;
  if (n_elems) *n_elems = p->n_elems;
  return TATVA_OK;
}

int tatva_plan_set_tiles(tatva_plan_t* p, const int32_t* d_tile_ptr, const int32_t* d_tile_nodes,
                         const uint16_t* d_tile_conn, int max_unique) {
  if (!p || (d_tile_ptr && (!d_tile_nodes || !d_tile_conn || max_unique <= 0))) return TATVA_E_INVALID;
  p->tile_ptr = d_tile_ptr;
  p->tile_nodes = d_tile_nodes;
  p->tile_conn = d_tile_conn;
  p->tile_max_unique = d_tile_ptr ? max_unique : 0;
  return TATVA_OK;
}

int tatva_plan_set_node_schedule(tatva_plan_t* p, const int32_t* d_warp_nodes, const uint8_t* d_warp_local,
                                 const int32_t* d_tile_hdr, const int32_t* d_tn_node, const int32_t* d_ell_ptr,
                                 const uint16_t* d_ell) {
  if (!p || (d_warp_nodes && (!d_warp_local || !d_tile_hdr || !d_tn_node || !d_ell_ptr || !d_ell))) return TATVA_E_INVALID;
  if (d_warp_nodes && p->npe > 8) return TATVA_E_UNSUPPORTED;
  p->ws_warp_nodes = d_warp_nodes;
  p->ws_warp_local = d_warp_local;
  p->ws_tile_hdr = d_tile_hdr;
  p->ws_tn_node = d_tn_node;
  p->ws_ell_ptr = d_ell_ptr;
  p->ws_ell = d_ell;
  return TATVA_OK;
}

// Uniform background grid for point location: bin (ix, iy) = (clamp(floor((x - lo_x) * inv_x)), ...), row-major
// iy * nx + ix; d_bin_elems[d_bin_ptr[b] .. d_bin_ptr[b+1]) = the elements whose bounding box overlaps bin b, ascending
// (tatva_host_build_point_grid).  Device views, caller-owned; NULL disables (every element is scanned).
int tatva_plan_set_point_grid(tatva_plan_t* p, int nx, int ny, const double* lo, const double* inv,
                              const int32_t* d_bin_ptr, const int32_t* d_bin_elems) {
  if (!p) return TATVA_E_INVALID;
  if (!d_bin_ptr) {
    p->grid_ptr = p->grid_elems = nullptr;
    return TATVA_OK;
  }
  if (!d_bin_elems || !lo || !inv || nx <= 0 || ny <= 0 || p->dim != 2) return TATVA_E_INVALID;
  p->grid_nx = nx;
  p->grid_ny = ny;
  p->grid_lo[0] = lo[0]; p->grid_lo[1] = lo[1];
  p->grid_inv[0] = inv[0]; p->grid_inv[1] = inv[1];
  p->grid_ptr = d_bin_ptr;
  p->grid_elems = d_bin_elems;
  return TATVA_OK;
}

int tatva_plan_rebind(tatva_plan_t* p, const double* d_coords, const int32_t* d_conn) {
  if (!p || !d_coords || !d_conn) return TATVA_E_INVALID;
  if ((p->weights || p->geo) && (d_coords != p->coords || (p->geo && d_conn != p->conn))) return TATVA_E_UNSUPPORTED;  // cached weights / geometry belong to the old mesh buffers
  p->coords = d_coords;
  p->conn = d_conn;
  return TATVA_OK;
}

// Mesh-only part of the Hex8 x neo-Hookean Gauss-point arithmetic, computed once and kept by the plan (64 doubles per
// element; see k_hex8_geometry).  enable = 0 frees it.  Allocates: call it at set-up time, not inside a captured region.
int tatva_plan_cache_geometry(tatva_plan_t* p, int enable, tatva_stream_t stream) {
  if (!p) return TATVA_E_INVALID;
  if (!enable) {
    if (p->geo) TATVA_CUDA_TRY(cudaFree(p->geo));
    p->geo = nullptr;
    p->geo_stride = 0;
    return TATVA_OK;
  }
  if (p->element != TATVA_HEX8 || p->custom) return TATVA_E_UNSUPPORTED;
  if (!p->geo) {
    TATVA_CUDA_TRY(cudaMalloc(&p->geo, sizeof(double) * 64 * (size_t)p->n_elems));
    p->geo_stride = p->n_elems;
  }
  return hex8_geometry_cache(p, p->geo, p->geo_stride, (cudaStream_t)stream);
}

int tatva_plan_set_variant(tatva_plan_t* p, int variant) {
  if (!p || variant < 0 || variant > 63) return TATVA_E_INVALID;
  p->variant = variant;
  return TATVA_OK;
}

int tatva_plan_set_quadrature(tatva_plan_t* p, int nq, const double* points, const double* weights, tatva_stream_t stream) {
  if (!p) return TATVA_E_INVALID;
  int rdim = 0, def_nq = 0;
  DISPATCH_ELEMENT(p, (rdim = El::rdim, def_nq = El::nq));
  if (nq == 0) {  // back to the element's default rule
    p->custom = 0;
    p->nq = def_nq;
  } else {
    if (!points || !weights || nq < 0 || nq > kMaxQ) return TATVA_E_INVALID;
    p->custom = 1;
    p->nq = nq;
    p->rule.nq = nq;
    for (int q = 0; q < kMaxQ; ++q) {
      p->rule.w[q] = q < nq ? weights[q] : 0.0;
      for (int d = 0; d < 3; ++d) p->rule.xi[q][d] = (q < nq && d < rdim) ? points[q * rdim + d] : 0.0;
    }
  }
  if (p->weights) {  // cached det J * w: recompute with the new rule
    cudaFree(p->weights);
    p->weights = nullptr;
    double* w = nullptr;
    TATVA_CUDA_TRY(cudaMalloc(&w, sizeof(double) * p->n_elems * p->nq));
    const int rc = tatva_op_integration_weights(p, w, stream);
    p->weights = w;
    if (rc != TATVA_OK) return rc;
  }
  return TATVA_OK;
}

int tatva_op_integration_weights(const tatva_plan_t* p, double* d_out, tatva_stream_t stream) {
  if (!p || !d_out) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->weights) {
    TATVA_CUDA_TRY(cudaMemcpyAsync(d_out, p->weights, sizeof(double) * p->n_elems * p->nq, cudaMemcpyDeviceToDevice, st));
    return TATVA_OK;
  }
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_weights<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_out)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->element == TATVA_HEX8 && p->variant == TATVA_VARIANT_DEFAULT) {
    const int rc = hex8_weights_modal(p, d_out, st);
    if (rc != TATVA_E_UNSUPPORTED) return rc;
  }
  if (p->nq > 1 && p->variant == 3) {  // thread per quadrature point: measured slower (see k_weights_qp)
    DISPATCH_ELEMENT(p, (k_weights_qp<El><<<grid_for(p->n_elems * p->nq), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_out)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  DISPATCH_ELEMENT(p, (k_weights<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_out)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_grad(const tatva_plan_t* p, const double* d_u, int nv, double* d_out, tatva_stream_t stream) {
  if (!p || !d_u || !d_out || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_grad<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_out)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->element == TATVA_HEX8 && nv <= 3 && p->variant == TATVA_VARIANT_DEFAULT) return hex8_grad_modal(p, false, d_u, nv, d_out, st);  // modal form (variant 2: the generic staged kernel)
  if (p->nq > 1 && p->variant == 3) {  // one thread per quadrature point: measured slower, kept as variant 3
    const size_t smem = (size_t)(kBlock / 32) * 32 * ((nv * p->gdim) | 1) * sizeof(double);
    if (smem <= 48 * 1024) {
      DISPATCH_ELEMENT(p, (k_field_qp<El, true><<<grid_for(p->n_elems * p->nq), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_out)));
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  {
    const size_t smem = (size_t)(kBlock / 32) * 32 * ((p->nq * nv * p->gdim) | 1) * sizeof(double);
    if (smem <= kStageMax && p->variant != TATVA_VARIANT_GENERIC) {
      STAGED_OPT_IN(k_grad_staged);
      DISPATCH_ELEMENT(p, (k_grad_staged<El><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_out)));
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  DISPATCH_ELEMENT(p, (k_grad<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_out)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_grad_adjoint(const tatva_plan_t* p, const double* d_g, int nv, double* d_y, tatva_stream_t stream) {
  if (!p || !d_g || !d_y || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * p->n_nodes * nv, st));
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_grad_adjoint<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_g, nv, d_y)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->element == TATVA_HEX8 && nv <= 3 && p->variant == TATVA_VARIANT_DEFAULT) return hex8_grad_modal(p, true, d_g, nv, d_y, st);
  if (nv <= 4 && p->variant != TATVA_VARIANT_GENERIC) {
    int rc = TATVA_OK;
#define ADJ_ACC(NV)                                                                                              \
  {                                                                                                              \
    constexpr int CH_ = El::max_nq * NV * El::gdim, S_ = CH_ | 1, SC_ = grouped_scatter_words<El::npe, NV>();            \
    constexpr size_t smem_ = (size_t)(kBlock / 32) * ((32 * S_ > SC_) ? 32 * S_ : SC_) * sizeof(double);         \
    static SmemOptIn done_;                                                                                      \
    rc = allow_big_smem(k_grad_adjoint_acc<El, NV>, done_);                                                      \
    if (rc == TATVA_OK)                                                                                          \
      k_grad_adjoint_acc<El, NV><<<grid_for(p->n_elems), kBlock, smem_, st>>>(p->coords, p->conn, p->n_elems, d_g, d_y); \
  }
    switch (nv) {
      case 1: DISPATCH_ELEMENT(p, ADJ_ACC(1)); break;
      case 2: DISPATCH_ELEMENT(p, ADJ_ACC(2)); break;
      case 3: DISPATCH_ELEMENT(p, ADJ_ACC(3)); break;
      default: DISPATCH_ELEMENT(p, ADJ_ACC(4)); break;
    }
#undef ADJ_ACC
    if (rc != TATVA_OK) return rc;
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  {
    const size_t smem = (size_t)(kBlock / 32) * 32 * ((p->nq * nv * p->gdim) | 1) * sizeof(double);
    if (smem <= kStageMax && p->variant != TATVA_VARIANT_GENERIC) {
      STAGED_OPT_IN(k_grad_adjoint_staged);
      DISPATCH_ELEMENT(p, (k_grad_adjoint_staged<El><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, d_g, nv, d_y)));
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  DISPATCH_ELEMENT(p, (k_grad_adjoint<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, d_g, nv, d_y)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_eval(const tatva_plan_t* p, const double* d_u, int nv, double* d_out, tatva_stream_t stream) {
  if (!p || !d_u || !d_out || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_eval<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->conn, p->n_elems, d_u, nv, d_out)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  if (p->nq > 1 && p->variant == 3) {  // one thread per quadrature point: measured slower, kept as variant 3
    const size_t smem = (size_t)(kBlock / 32) * 32 * (nv | 1) * sizeof(double);
    if (smem <= 48 * 1024) {
      DISPATCH_ELEMENT(p, (k_field_qp<El, false><<<grid_for(p->n_elems * p->nq), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_out)));
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  {
    const size_t smem = (size_t)(kBlock / 32) * 32 * ((p->nq * nv) | 1) * sizeof(double);
    if (smem <= kStageMax && p->variant != TATVA_VARIANT_GENERIC) {
      STAGED_OPT_IN(k_eval_staged);
      DISPATCH_ELEMENT(p, (k_eval_staged<El><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->conn, p->n_elems, d_u, nv, d_out)));
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  DISPATCH_ELEMENT(p, (k_eval<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->conn, p->n_elems, d_u, nv, d_out)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_eval_adjoint(const tatva_plan_t* p, const double* d_g, int nv, double* d_y, tatva_stream_t stream) {
  if (!p || !d_g || !d_y || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * p->n_nodes * nv, st));
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_eval_adjoint<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->conn, p->n_elems, d_g, nv, d_y)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  DISPATCH_ELEMENT(p, (k_eval_adjoint<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->conn, p->n_elems, d_g, nv, d_y)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_integrate_quad(const tatva_plan_t* p, const double* d_vals, int nv, double* d_out, tatva_stream_t stream) {
  if (!p || !d_vals || !d_out || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->custom) {
    TATVA_CUDA_TRY(install_rule(p, st));
    DISPATCH_CUSTOM(p, (k_integrate_quad<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, p->weights, d_vals, nv, d_out)));
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  DISPATCH_ELEMENT(p, (k_integrate_quad<El><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, p->weights, d_vals, nv, d_out)));
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// Operator.interpolate (tatva/operator.py:399-463) for the plane elements: values of the nodal field u (n_nodes, nv) at
// `n_points` physical points (n_points, 2) -> d_out (n_points, nv); d_elem (n_points) receives the containing
// element of every point, -1 (and NaN values) where the point lies outside the mesh.
int tatva_op_interpolate(const tatva_plan_t* p, const double* d_u, int nv, const double* d_points, int64_t n_points,
                         double* d_out, int32_t* d_elem, tatva_stream_t stream) {
  if (p && p->custom) return TATVA_E_UNSUPPORTED;  // the Newton start is the element's first DEFAULT point
  if (!p || !d_u || !d_points || !d_out || !d_elem || nv <= 0 || n_points <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)((n_points + 127) / 128);
  if (p->grid_ptr && p->variant != TATVA_VARIANT_GENERIC) {
    PointGrid g{p->grid_nx, p->grid_ny, {p->grid_lo[0], p->grid_lo[1]}, {p->grid_inv[0], p->grid_inv[1]}, p->grid_ptr, p->grid_elems};
    switch (p->element) {
      case TATVA_TRI3: k_interpolate_grid<Tri3><<<grid, 128, 0, st>>>(p->coords, p->conn, g, d_u, nv, d_points, n_points, d_out, d_elem); break;
      case TATVA_QUAD4: k_interpolate_grid<Quad4><<<grid, 128, 0, st>>>(p->coords, p->conn, g, d_u, nv, d_points, n_points, d_out, d_elem); break;
      case TATVA_TRI6: k_interpolate_grid<Tri6><<<grid, 128, 0, st>>>(p->coords, p->conn, g, d_u, nv, d_points, n_points, d_out, d_elem); break;
      case TATVA_QUAD8: k_interpolate_grid<Quad8><<<grid, 128, 0, st>>>(p->coords, p->conn, g, d_u, nv, d_points, n_points, d_out, d_elem); break;
      default: return TATVA_E_UNSUPPORTED;
    }
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  switch (p->element) {
    case TATVA_TRI3: k_interpolate<Tri3><<<grid, 128, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_points, n_points, d_out, d_elem); break;
    case TATVA_QUAD4: k_interpolate<Quad4><<<grid, 128, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_points, n_points, d_out, d_elem); break;
    case TATVA_TRI6: k_interpolate<Tri6><<<grid, 128, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_points, n_points, d_out, d_elem); break;
    case TATVA_QUAD8: k_interpolate<Quad8><<<grid, 128, 0, st>>>(p->coords, p->conn, p->n_elems, d_u, nv, d_points, n_points, d_out, d_elem); break;
    default: return TATVA_E_UNSUPPORTED;  // the reference's point search is two-dimensional (mesh.py:303)
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_gather(const tatva_plan_t* p, const double* d_u, int nv, double* d_out, tatva_stream_t stream) {
  if (!p || !d_u || !d_out || nv <= 0) return TATVA_E_INVALID;
  const int64_t total = p->n_elems * p->npe * nv;
  if (total < (int64_t)0xfffff000 && p->variant == TATVA_VARIANT_DEFAULT) {
    const int grid = (int)((total + 1023) / 1024);
    cudaStream_t st = (cudaStream_t)stream;
    switch (nv) {
      case 1: k_gather4<1><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_u, d_out); break;
      case 2: k_gather4<2><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_u, d_out); break;
      case 3: k_gather4<3><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_u, d_out); break;
      case 4: k_gather4<4><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_u, d_out); break;
      default: k_gather4<0><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_u, d_out); break;
    }
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  k_gather<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p->conn, total, nv, d_u, d_out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_gather_adjoint(const tatva_plan_t* p, const double* d_g, int nv, double* d_y, tatva_stream_t stream) {
  if (!p || !d_g || !d_y || nv <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * p->n_nodes * nv, st));
  const int64_t total = p->n_elems * p->npe * nv;
  if (total < (int64_t)0xfffff000 && p->variant == TATVA_VARIANT_DEFAULT) {
    const int grid = (int)((total + 1023) / 1024);
    switch (nv) {
      case 1: k_gather_adjoint4<1><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_g, d_y); break;
      case 2: k_gather_adjoint4<2><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_g, d_y); break;
      case 3: k_gather_adjoint4<3><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_g, d_y); break;
      case 4: k_gather_adjoint4<4><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_g, d_y); break;
      default: k_gather_adjoint4<0><<<grid, 256, 0, st>>>(p->conn, (uint32_t)total, nv, d_g, d_y); break;
    }
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  k_gather_adjoint<<<grid_for(total, 256), 256, 0, st>>>(p->conn, total, nv, d_g, d_y);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_op_sum_rows(tatva_plan_t* p, const double* d_in, int64_t rows, int nv, double* d_out, tatva_stream_t stream) {
  if (!p || !d_in || !d_out || rows <= 0 || nv <= 0 || nv > 64) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int nblocks = (int)((rows + 4095) / 4096);
  if (nblocks > 1024) nblocks = 1024;
  const int64_t rpb = (rows + nblocks - 1) / nblocks;
  k_sum_rows_partial<<<nblocks, 256, 0, st>>>(d_in, rows, nv, rpb, p->scratch);
  k_sum_rows_final<<<nv, 256, 0, st>>>(p->scratch, nblocks, nv, d_out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// ---- fused dispatch ------------------------------------------------------------------------------
}  // extern "C"

namespace tatva {
int sum_partials(const double* partials, int n, double* out, cudaStream_t st) {
  k_sum_rows_final<<<1, 256, 0, st>>>(partials, n, 1, out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
}  // namespace tatva

template <class El, class Mat, int MODE>
static int launch_fused(tatva_plan* p, const Mat& mat, const double* u, const double* v, double* out, cudaStream_t st) {
  const int grid = grid_for(p->n_elems);
  if (MODE == MODE_ENERGY) {
    k_fused<El, Mat, MODE><<<grid, kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mat, u, v, nullptr, p->scratch);
    k_sum_rows_final<<<1, 256, 0, st>>>(p->scratch, grid, 1, out);
  } else {
    constexpr size_t smem = grouped_scatter_smem<El::npe, Mat::dpn>(kBlock / 32);
    static_assert(smem <= 48 * 1024, "grouped scatter staging exceeds the default shared-memory window");
    // y is cleared under the kernel: grouped_scatter waits (griddepcontrol.wait) before its first atomic add
    const int rc = launch_behind_zero(k_fused<El, Mat, MODE>, grid, kBlock, smem, st, p->zero_output != 0, out, p->n_nodes * Mat::dpn,
                                      p->coords, p->conn, p->n_elems, mat, u, v, out, nullptr);
    if (rc != TATVA_OK) return rc;
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

template <int MODE>
static int dispatch_fused(tatva_plan* p, int material, const double* prm, int n_params, const double* u, const double* v,
                          double* out, cudaStream_t st) {
  if (!p || !prm || !u || !out) return TATVA_E_INVALID;
  if (MODE == MODE_HVP && !v) return TATVA_E_INVALID;
  const int el = p->element;
  if (is_user_law(material)) return user_law_launch(p, material, MODE, prm, n_params, u, v, nullptr, out, st);
  if (p->custom) {  // user quadrature rule: the generic template over Custom<El> (the specialised kernels carry the default rule)
    TATVA_CUDA_TRY(install_rule(p, st));
    if (material == TATVA_LINEAR_ELASTIC && n_params == 2) {
      if (el == TATVA_TRI3) return launch_fused<Custom<Tri3>, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_QUAD4) return launch_fused<Custom<Quad4>, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_TRI6) return launch_fused<Custom<Tri6>, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_QUAD8) return launch_fused<Custom<Quad8>, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_TET4) return launch_fused<Custom<Tet4>, LinearElastic<3>, MODE>(p, LinearElastic<3>{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_HEX8) return launch_fused<Custom<Hex8>, LinearElastic<3>, MODE>(p, LinearElastic<3>{prm[0], prm[1]}, u, v, out, st);
    } else if (material == TATVA_NEO_HOOKEAN && n_params == 2) {
      if (el == TATVA_TET4) return launch_fused<Custom<Tet4>, NeoHookean, MODE>(p, NeoHookean{prm[0], prm[1]}, u, v, out, st);
      if (el == TATVA_HEX8) return launch_fused<Custom<Hex8>, NeoHookean, MODE>(p, NeoHookean{prm[0], prm[1]}, u, v, out, st);
    } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD && n_params == 5) {
      const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
      if (el == TATVA_TET4) return launch_fused<Custom<Tet4>, NeoHookeanPhaseField, MODE>(p, m, u, v, out, st);
      if (el == TATVA_HEX8) return launch_fused<Custom<Hex8>, NeoHookeanPhaseField, MODE>(p, m, u, v, out, st);
    }
    return TATVA_E_UNSUPPORTED;
  }
  if constexpr (MODE != MODE_ENERGY) {
    // warp-cooperative kernels for the one-point simplices when the plan carries a node schedule
    if (p->ws_warp_nodes && p->variant != TATVA_VARIANT_GENERIC && p->variant != 30 && p->variant != 31) {
      if (material == TATVA_LINEAR_ELASTIC && n_params == 2) {
        if (el == TATVA_TRI3) return launch_fused_wc<Tri3, LinearElastic<2>, MODE, GenericBody<Tri3, LinearElastic<2>>>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
        if (el == TATVA_TET4) return launch_fused_wc<Tet4, LinearElastic<3>, MODE, GenericBody<Tet4, LinearElastic<3>>>(p, LinearElastic<3>{prm[0], prm[1]}, u, v, out, st);
      } else if (material == TATVA_NEO_HOOKEAN && n_params == 2) {
        if (el == TATVA_TET4) return tet4_nh_wc(p, MODE == MODE_HVP, prm[0], prm[1], u, v, out, st);
      } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD && n_params == 5 && el == TATVA_TET4) {
        // two-field law (32-byte nodal rows): shuffle gather + tile sums 0.0678 -> 0.0630 ms (HVP), 0.0475 -> 0.0453 (residual)
        // at config 5; variant 38 = the element's own gather + tile sums (no gain: 0.0678 / 0.0514)
        const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
        if (p->variant == 38) return launch_fused_wc<Tet4, NeoHookeanPhaseField, MODE, GenericBody<Tet4, NeoHookeanPhaseField>, false, true>(p, m, u, v, out, st);
        return launch_fused_wc<Tet4, NeoHookeanPhaseField, MODE, GenericBody<Tet4, NeoHookeanPhaseField>, true, true>(p, m, u, v, out, st);
      }
    }
  }
  if (material == TATVA_LINEAR_ELASTIC) {
    if (n_params != 2) return TATVA_E_INVALID;
    if (el == TATVA_TRI3) return launch_fused<Tri3, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
    if (el == TATVA_QUAD4) return launch_fused<Quad4, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
    if (el == TATVA_TRI6) return launch_fused<Tri6, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
    if (el == TATVA_QUAD8) return launch_fused<Quad8, LinearElastic<2>, MODE>(p, LinearElastic<2>{prm[0], prm[1]}, u, v, out, st);
    if (el == TATVA_TET4) return launch_fused<Tet4, LinearElastic<3>, MODE>(p, LinearElastic<3>{prm[0], prm[1]}, u, v, out, st);
    if (el == TATVA_HEX8) return launch_fused<Hex8, LinearElastic<3>, MODE>(p, LinearElastic<3>{prm[0], prm[1]}, u, v, out, st);
  } else if (material == TATVA_NEO_HOOKEAN) {
    if (n_params != 2) return TATVA_E_INVALID;
    if (el == TATVA_TET4) {
      if (MODE != MODE_ENERGY && p->tile_ptr && p->variant != TATVA_VARIANT_GENERIC && p->zero_output)
        return tet4_nh_tiled(p, MODE == MODE_HVP, prm[0], prm[1], u, v, out, st);
      if (MODE == MODE_HVP && p->variant != TATVA_VARIANT_GENERIC) return tet4_nh_hvp_ref(p, prm[0], prm[1], u, v, out, st);
      if (MODE == MODE_RESIDUAL && p->variant != TATVA_VARIANT_GENERIC) return tet4_nh_residual_ref(p, prm[0], prm[1], u, out, st);
      return launch_fused<Tet4, NeoHookean, MODE>(p, NeoHookean{prm[0], prm[1]}, u, v, out, st);
    }
    if (el == TATVA_HEX8) {
      if (MODE == MODE_HVP && p->variant != TATVA_VARIANT_GENERIC) return hex8_nh_hvp_modal(p, prm[0], prm[1], u, v, out, st);
      if (MODE == MODE_RESIDUAL && p->variant != TATVA_VARIANT_GENERIC) return hex8_nh_residual_modal(p, prm[0], prm[1], u, out, st);
      if (MODE == MODE_ENERGY && p->variant != TATVA_VARIANT_GENERIC) {
        int rc = hex8_nh_energy_modal_partials(p, prm[0], prm[1], u, st);
        if (rc != TATVA_OK) return rc;
        k_sum_rows_final<<<1, 256, 0, st>>>(p->scratch, grid_for(p->n_elems), 1, out);
        TATVA_LAUNCH_CHECK();
        return TATVA_OK;
      }
      return launch_fused<Hex8, NeoHookean, MODE>(p, NeoHookean{prm[0], prm[1]}, u, v, out, st);
    }
  } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD) {
    if (n_params != 5) return TATVA_E_INVALID;
    const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
    if (el == TATVA_TET4) return launch_fused<Tet4, NeoHookeanPhaseField, MODE>(p, m, u, v, out, st);
    if (el == TATVA_HEX8) return launch_fused<Hex8, NeoHookeanPhaseField, MODE>(p, m, u, v, out, st);
  } else {
    return TATVA_E_INVALID;
  }
  return TATVA_E_UNSUPPORTED;
}

extern "C" {

int tatva_energy(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u, double* d_energy,
                 tatva_stream_t stream) {
  return dispatch_fused<MODE_ENERGY>(p, material, params, n_params, d_u, nullptr, d_energy, (cudaStream_t)stream);
}
int tatva_residual(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u, double* d_r,
                   tatva_stream_t stream) {
  return dispatch_fused<MODE_RESIDUAL>(p, material, params, n_params, d_u, nullptr, d_r, (cudaStream_t)stream);
}
int tatva_hvp(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u, const double* d_v,
              double* d_y, tatva_stream_t stream) {
  return dispatch_fused<MODE_HVP>(p, material, params, n_params, d_u, d_v, d_y, (cudaStream_t)stream);
}

}  // extern "C"

// (element, law) pairs of the fused kernels, as a visitor: f(El{}, mat) for the pair named by the enums.
template <class F>
static int for_element_law(int el, int material, const double* prm, int n_params, F&& f, bool custom = false) {
#define TATVA_EL_LAW(ID, B, LAW) \
  if (el == ID) return custom ? f(Custom<B>{}, LAW) : f(B{}, LAW);
  if (material == TATVA_LINEAR_ELASTIC) {
    if (n_params != 2) return TATVA_E_INVALID;
    const LinearElastic<2> m2{prm[0], prm[1]};
    const LinearElastic<3> m3{prm[0], prm[1]};
    TATVA_EL_LAW(TATVA_TRI3, Tri3, m2)
    TATVA_EL_LAW(TATVA_QUAD4, Quad4, m2)
    TATVA_EL_LAW(TATVA_TRI6, Tri6, m2)
    TATVA_EL_LAW(TATVA_QUAD8, Quad8, m2)
    TATVA_EL_LAW(TATVA_TET4, Tet4, m3)
    TATVA_EL_LAW(TATVA_HEX8, Hex8, m3)
  } else if (material == TATVA_NEO_HOOKEAN) {
    if (n_params != 2) return TATVA_E_INVALID;
    const NeoHookean m{prm[0], prm[1]};
    TATVA_EL_LAW(TATVA_TET4, Tet4, m)
    TATVA_EL_LAW(TATVA_HEX8, Hex8, m)
  } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD) {
    if (n_params != 5) return TATVA_E_INVALID;
    const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
    TATVA_EL_LAW(TATVA_TET4, Tet4, m)
    TATVA_EL_LAW(TATVA_HEX8, Hex8, m)
  } else {
    return TATVA_E_INVALID;
  }
#undef TATVA_EL_LAW
  return TATVA_E_UNSUPPORTED;
}

extern "C" {

// Diagonal of the energy Hessian at u (length n_nodes * dofs_per_node), for Jacobi preconditioning.
int tatva_hessian_diag(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u,
                       double* d_diag, tatva_stream_t stream) {
  if (!p || !params || !d_u || !d_diag) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_user_law(material)) return user_law_launch(p, material, 3, params, n_params, d_u, nullptr, nullptr, d_diag, st);
  if (p->custom) TATVA_CUDA_TRY(install_rule(p, st));
  return for_element_law(p->element, material, params, n_params, [&](auto el, auto mat) -> int {
    using El = decltype(el);
    using Mat = decltype(mat);
    if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(d_diag, 0, sizeof(double) * p->n_nodes * Mat::dpn, st));
    constexpr size_t smem = grouped_scatter_smem<El::npe, Mat::dpn>(kBlock / 32);
    if constexpr (has_rank_law<Mat>::value) {
      if (p->variant != TATVA_VARIANT_GENERIC) {
        k_hessian_diag_rank<El, Mat><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mat, d_u, d_diag);
        TATVA_LAUNCH_CHECK();
        return TATVA_OK;
      }
    }
    k_hessian_diag<El, Mat><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mat, d_u, d_diag);
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }, p->custom != 0);
}

// HVP on the free DOFs of a Lifter in ONE kernel: y_red = reduce_adjoint(H(u_full) lift_0(v_red)).
int tatva_hvp_lifted(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u_full,
                     const double* d_v_red, const int32_t* d_dof_map, int64_t n_red, double* d_y_red,
                     tatva_stream_t stream) {
  if (!p || !params || !d_u_full || !d_v_red || !d_dof_map || !d_y_red || n_red <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(d_y_red, 0, sizeof(double) * n_red, st));
  if (is_user_law(material)) return user_law_launch(p, material, 4, params, n_params, d_u_full, d_v_red, d_dof_map, d_y_red, st);
  if (!p->custom && p->element == TATVA_HEX8 && material == TATVA_NEO_HOOKEAN && n_params == 2 && p->variant != TATVA_VARIANT_GENERIC)
    return hex8_nh_hvp_modal_lifted(p, params[0], params[1], d_u_full, d_v_red, d_dof_map, d_y_red, st);
  if (p->custom) TATVA_CUDA_TRY(install_rule(p, st));
  return for_element_law(p->element, material, params, n_params, [&](auto el, auto mat) -> int {
    using El = decltype(el);
    using Mat = decltype(mat);
    k_hvp_lifted<El, Mat><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mat, d_u_full, d_v_red, d_dof_map, d_y_red);
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }, p->custom != 0);
}

// The same with v_red . y_red computed on the way (CG: p . A p without a pass over the two vectors): the Hex8 x
// neo-Hookean kernel sums v_e . y_e per element before the scatter (one partial per CTA in the plan's scratch, then a
// fixed-order final sum); every other (element, law) pair runs the kernel above followed by a two-pass dot.  The scalar
// lands in d_scalars[slot]; with roll != 0 the final-sum kernel also copies d_scalars[2] to d_scalars[0] (the CG's
// "r.r <- next r.r" step of the previous iteration).  zero_y = 0 skips the memset (the caller's previous vector pass
// left y zeroed, see tatva_cg_direction_zero).
int tatva_hvp_lifted_dot(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u_full,
                         const double* d_v_red, const int32_t* d_dof_map, int64_t n_red, double* d_y_red, int zero_y,
                         double* d_partials, double* d_scalars, int slot, int roll, tatva_stream_t stream) {
  if (!p || !params || !d_u_full || !d_v_red || !d_dof_map || !d_y_red || !d_scalars || !d_partials || n_red <= 0 || slot < 0 || slot > 7) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_y) TATVA_CUDA_TRY(cudaMemsetAsync(d_y_red, 0, sizeof(double) * n_red, st));
  if (!p->custom && p->element == TATVA_HEX8 && material == TATVA_NEO_HOOKEAN && n_params == 2 && p->variant != TATVA_VARIANT_GENERIC) {
    const int grid = grid_for(p->n_elems);  // one partial per CTA
    if (grid > p->scratch_len) return TATVA_E_INVALID;
    const int rc = hex8_nh_hvp_modal_lifted(p, params[0], params[1], d_u_full, d_v_red, d_dof_map, d_y_red, st, p->scratch);
    if (rc != TATVA_OK) return rc;
    k_cg_finish_roll<<<1, 1024, 0, st>>>(p->scratch, grid, d_scalars, slot, roll);
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  tatva_plan q = *p;
  q.zero_output = 0;
  const int rc = tatva_hvp_lifted(&q, material, params, n_params, d_u_full, d_v_red, d_dof_map, n_red, d_y_red, stream);
  if (rc != TATVA_OK) return rc;
  k_cg_dot<<<kCgBlocks, 256, 0, st>>>(d_v_red, d_y_red, n_red, d_partials);
  k_cg_finish_roll<<<1, 1024, 0, st>>>(d_partials, kCgBlocks, d_scalars, slot, roll);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// Unconstrained counterpart of tatva_hvp_lifted_dot: y (+)= H(u) v and d_scalars[slot] = v . H v.  fuse_dot != 0 asks the
// Hex8 x neo-Hookean kernel to sum v_e . y_e itself; otherwise (and for every other pair) a two-pass dot follows.
int tatva_hvp_dot(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u, const double* d_v,
                  double* d_y, int zero_y, int fuse_dot, double* d_partials, double* d_scalars, int slot, int roll,
                  tatva_stream_t stream) {
  if (!p || !params || !d_u || !d_v || !d_y || !d_scalars || !d_partials || slot < 0 || slot > 7) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int dpn = material == TATVA_NEO_HOOKEAN_PHASE_FIELD ? 4 : p->dim;
  if (is_user_law(material)) user_law_info(material, &dpn);
      if (is_user_law(material)) user_law_info(material, &dpn);
  const int64_t n = p->n_nodes * dpn;
  if (zero_y) TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * n, st));
  if (fuse_dot && !p->custom && p->element == TATVA_HEX8 && material == TATVA_NEO_HOOKEAN && n_params == 2 && p->variant != TATVA_VARIANT_GENERIC) {
    const int grid = grid_for(p->n_elems);  // one partial per CTA
    if (grid > p->scratch_len) return TATVA_E_INVALID;
    const int rc = hex8_nh_hvp_modal_dot(p, params[0], params[1], d_u, d_v, d_y, p->scratch, st);
    if (rc != TATVA_OK) return rc;
    k_cg_finish_roll<<<1, 1024, 0, st>>>(p->scratch, grid, d_scalars, slot, roll);
    TATVA_LAUNCH_CHECK();
    return TATVA_OK;
  }
  tatva_plan q = *p;
  q.zero_output = 0;
  const int rc = dispatch_fused<MODE_HVP>(&q, material, params, n_params, d_u, d_v, d_y, st);
  if (rc != TATVA_OK) return rc;
  k_cg_dot<<<kCgBlocks, 256, 0, st>>>(d_v, d_y, n, d_partials);
  k_cg_finish_roll<<<1, 1024, 0, st>>>(d_partials, kCgBlocks, d_scalars, slot, roll);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

}  // extern "C"

// ---- host probe of the generic element body ---------------------------------------------------------
// The arithmetic of k_fused / k_hessian_diag for ONE element on the CPU, through the same element tables, geometry and
// constitutive-law functions compiled for the host: lets the CPU test suite check the kernels' formulas against the
// oracle.  mode 0 energy (out[0]), 1 residual, 2 HVP, 3 Hessian diagonal (out[npe][dpn]).  HOST pointers.
template <class El, class Mat>
static int probe_element(const Mat& mat, int mode, const double* Xp, const double* up, const double* vp, double* out) {
  constexpr int dpn = Mat::dpn;
  double X[El::npe][El::dim], U[El::npe][dpn], V[El::npe][dpn], Y[El::npe][dpn];
  for (int n = 0; n < El::npe; ++n) {
    for (int c = 0; c < El::dim; ++c) X[n][c] = Xp[n * El::dim + c];
    for (int c = 0; c < dpn; ++c) {
      U[n][c] = up[n * dpn + c];
      V[n][c] = vp ? vp[n * dpn + c] : 0.0;
      Y[n][c] = 0.0;
    }
  }
  double energy = 0.0;
  for (int q = 0; q < El::num_q(); ++q) {
    double dNdX[El::dim][El::npe], N[El::npe];
    const double W = geometry<El>(q, X, dNdX) * El::weight(q);
    El::N(q, N);
    typename Mat::S s, ds, f;
    typename Mat::Cache cache;
    qp_state<El, Mat>(dNdX, N, U, s);
    mat.prepare(s, cache);
    if (mode == 0) {
      energy += W * mat.psi(s, cache);
      continue;
    }
    if (mode == 4) {  // the rank-structured diagonal of k_hessian_diag_rank (laws with a RankLaw)
      if constexpr (has_rank_law<Mat>::value && Mat::dpn == El::dim) {
        double g[El::dim][El::npe], w[3];
        RankLaw<Mat>::template eval<El::npe>(mat, W, dNdX, U, g, w);
        for (int a = 0; a < El::npe; ++a) {
          double nn = 0.0;
          for (int j = 0; j < El::dim; ++j) nn = fma(dNdX[j][a], dNdX[j][a], nn);
          for (int i = 0; i < El::dim; ++i) Y[a][i] += fma((w[1] + w[2]) * g[i][a], g[i][a], w[0] * nn);
        }
        continue;
      } else {
        return TATVA_E_UNSUPPORTED;
      }
    }
    if (mode == 3) {
      for (int b = 0; b < El::npe; ++b)
        for (int k = 0; k < dpn; ++k) {
          for (int c = 0; c < dpn; ++c) {
            for (int j = 0; j < El::dim; ++j) ds.G[c][j] = (c == k) ? dNdX[j][b] : 0.0;
            ds.val[c] = (c == k) ? N[b] : 0.0;
          }
          mat.second(s, cache, ds, f);
          double t = 0.0;
          for (int j = 0; j < El::dim; ++j) t += f.G[k][j] * dNdX[j][b];
          if (k >= Mat::val_lo) t += f.val[k] * N[b];
          Y[b][k] += W * t;
        }
      continue;
    }
    if (mode == 1) {
      mat.first(s, cache, f);
    } else {
      qp_state<El, Mat>(dNdX, N, V, ds);
      mat.second(s, cache, ds, f);
    }
    for (int n = 0; n < El::npe; ++n)
      for (int c = 0; c < dpn; ++c) {
        double t = 0.0;
        for (int j = 0; j < El::dim; ++j) t += f.G[c][j] * dNdX[j][n];
        if (c >= Mat::val_lo) t += f.val[c] * N[n];
        Y[n][c] += W * t;
      }
  }
  if (mode == 0) {
    out[0] = energy;
  } else {
    for (int n = 0; n < El::npe; ++n)
      for (int c = 0; c < dpn; ++c) out[n * dpn + c] = Y[n][c];
  }
  return TATVA_OK;
}

extern "C" int tatva_probe_element(int element, int material, const double* params, int n_params, int mode, const double* X,
                                   const double* u, const double* v, double* out) {
  if (!params || !X || !u || !out || mode < 0 || mode > 4 || (mode == 2 && !v)) return TATVA_E_INVALID;
  return for_element_law(element, material, params, n_params, [&](auto el, auto mat) -> int {
    return probe_element<decltype(el), decltype(mat)>(mat, mode, X, u, v, out);
  });
}

// A sub-range that starts on a tile boundary (a multiple of 128 elements) sees the node schedule through offset views —
// the per-warp lists and per-tile headers are indexed by tile, the chunk and table arrays through the headers' absolute
// positions; a ragged END is fine (elements past it contribute zero rows).  Any other start drops the schedule.
static void sub_range_node_schedule(tatva_plan& sub, int64_t elem_begin) {
  if (!sub.ws_warp_nodes) return;
  if (elem_begin % 128 != 0) {
    sub.ws_warp_nodes = nullptr;
    return;
  }
  const int64_t tile0 = elem_begin / 128;
  sub.ws_warp_nodes += tile0 * 128;
  sub.ws_warp_local += elem_begin * sub.npe;
  sub.ws_tile_hdr += tile0 * 4;
}

extern "C" {
// y[0..n) = 0 by the kernel that releases its dependent grid at once (see launch_behind_zero): a caller that splits one
// application into several launches (the partitioned operator: interior and boundary elements on two streams) clears the
// output with this and passes zero_y = 2 to the FIRST sub-range launch it issues right behind it on the same stream.
int tatva_zero_release(double* d_y, int64_t n, tatva_stream_t stream) {
  if (!d_y || n < 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return TATVA_OK;
  if (reinterpret_cast<uintptr_t>(d_y) & 15) {
    TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * n, st));
    return TATVA_OK;
  }
  int64_t blocks = (n / 2 + 256 * 8 - 1) / (256 * 8);
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  k_zero_release<<<(int)blocks, 256, 0, st>>>(d_y, n);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// Element sub-range variants (overlap of halo exchange with interior elements): elements
// [elem_begin, elem_begin + elem_count) only; y is zeroed first iff zero_y != 0.
int tatva_hvp_elems(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u,
                    const double* d_v, double* d_y, int64_t elem_begin, int64_t elem_count, int zero_y,
                    tatva_stream_t stream) {
  if (!p || elem_begin < 0 || elem_count < 0 || elem_begin + elem_count > p->n_elems) return TATVA_E_INVALID;
  if (elem_count == 0) {
    if (zero_y == 1) {
      int dpn = material == TATVA_NEO_HOOKEAN_PHASE_FIELD ? 4 : p->dim;
      if (is_user_law(material)) user_law_info(material, &dpn);
      TATVA_CUDA_TRY(cudaMemsetAsync(d_y, 0, sizeof(double) * p->n_nodes * dpn, (cudaStream_t)stream));
    }
    return TATVA_OK;
  }
  tatva_plan sub = *p;
  sub.conn = p->conn + elem_begin * p->npe;
  sub.n_elems = elem_count;
  sub.zero_output = zero_y == 1 ? 1 : 0;
  sub.pss = zero_y == 2 ? 1 : 0;
  sub.tile_ptr = nullptr;  // staging tiles describe the whole element list, not a sub-range
  sub_range_node_schedule(sub, elem_begin);
  if (sub.geo) sub.geo += 2 * elem_begin;  // the cache is element-fastest: a sub-range is an offset view (same stride)
  return dispatch_fused<MODE_HVP>(&sub, material, params, n_params, d_u, d_v, d_y, (cudaStream_t)stream);
}
int tatva_residual_elems(tatva_plan_t* p, int material, const double* params, int n_params, const double* d_u,
                         double* d_r, int64_t elem_begin, int64_t elem_count, int zero_r, tatva_stream_t stream) {
  if (!p || elem_begin < 0 || elem_count < 0 || elem_begin + elem_count > p->n_elems) return TATVA_E_INVALID;
  if (elem_count == 0) {
    if (zero_r) {
      int dpn = material == TATVA_NEO_HOOKEAN_PHASE_FIELD ? 4 : p->dim;
      if (is_user_law(material)) user_law_info(material, &dpn);
      TATVA_CUDA_TRY(cudaMemsetAsync(d_r, 0, sizeof(double) * p->n_nodes * dpn, (cudaStream_t)stream));
    }
    return TATVA_OK;
  }
  tatva_plan sub = *p;
  sub.conn = p->conn + elem_begin * p->npe;
  sub.n_elems = elem_count;
  sub.zero_output = zero_r ? 1 : 0;
  sub.tile_ptr = nullptr;
  sub_range_node_schedule(sub, elem_begin);
  return dispatch_fused<MODE_RESIDUAL>(&sub, material, params, n_params, d_u, nullptr, d_r, (cudaStream_t)stream);
}
}  // extern "C"

template <class El, class Mat>
static int launch_csr(tatva_plan* p, const Mat& mat, const double* u, const int32_t* indptr, const int32_t* pos,
                      int64_t nnz, double* data, cudaStream_t st, const int32_t* indices = nullptr) {
  TATVA_CUDA_TRY(cudaMemsetAsync(data, 0, sizeof(double) * nnz, st));
  if constexpr (El::max_nq == 1) {
    if (p->variant != TATVA_VARIANT_GENERIC) {
      constexpr int S = El::npe * Mat::dpn * Mat::dpn, NB = El::npe * Mat::dpn;
      constexpr size_t smem = (size_t)(kBlock / 32) * (32 * (S | 1) + 16 * (NB | 1)) * sizeof(double);
      static SmemOptIn full, sym;
      if (smem > 48 * 1024) {
        int rc = opt_in_smem(k_csr_grouped<El, Mat, false>, smem, full);
        if (rc == TATVA_OK) rc = opt_in_smem(k_csr_grouped<El, Mat, true>, smem, sym);
        if (rc != TATVA_OK) return rc;
      }
      if (p->variant == 2 || indices == nullptr) {  // full assembly: every entry by REDs
        k_csr_grouped<El, Mat, false><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mat, u, indptr, pos, data);
      } else {  // upper triangle by REDs, lower triangle mirrored (symmetric energy Hessian)
        k_csr_grouped<El, Mat, true><<<grid_for(p->n_elems), kBlock, smem, st>>>(p->coords, p->conn, p->n_elems, mat, u, indptr, pos, data);
        k_csr_mirror<<<(int)((p->n_nodes + 3) / 4), 128, 0, st>>>(p->n_nodes, Mat::dpn, indptr, indices, data);
      }
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }
  }
  k_csr<El, Mat><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mat, u, indptr, pos, data);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

static int csr_dispatch(tatva_plan_t* p, int material, const double* prm, int n_params, const double* d_u,
                        const int32_t* d_indptr, const int32_t* d_indices, const int32_t* d_pos, int64_t nnz,
                        double* d_data, tatva_stream_t stream) {
  if (!p || !prm || !d_u || !d_indptr || !d_pos || !d_data || nnz <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int el = p->element;
  if (is_user_law(material)) return TATVA_E_UNSUPPORTED;  // sparse.jacfwd then runs the coloured route on the fused HVP
  if (p->custom) {  // user quadrature rule: the generic per-entry kernel over Custom<El>
    TATVA_CUDA_TRY(install_rule(p, st));
    return for_element_law(el, material, prm, n_params, [&](auto e, auto mat) -> int {
      using El = decltype(e);
      using Mat = decltype(mat);
      TATVA_CUDA_TRY(cudaMemsetAsync(d_data, 0, sizeof(double) * nnz, st));
      k_csr<El, Mat><<<grid_for(p->n_elems), kBlock, 0, st>>>(p->coords, p->conn, p->n_elems, mat, d_u, d_indptr, d_pos, d_data);
      TATVA_LAUNCH_CHECK();
      return TATVA_OK;
    }, true);
  }
  if (material == TATVA_LINEAR_ELASTIC && n_params == 2) {
    if (el == TATVA_TRI3) return launch_csr<Tri3, LinearElastic<2>>(p, LinearElastic<2>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_QUAD4) return launch_csr<Quad4, LinearElastic<2>>(p, LinearElastic<2>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_TRI6) return launch_csr<Tri6, LinearElastic<2>>(p, LinearElastic<2>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_QUAD8) return launch_csr<Quad8, LinearElastic<2>>(p, LinearElastic<2>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_TET4) return launch_csr<Tet4, LinearElastic<3>>(p, LinearElastic<3>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_HEX8) return launch_csr<Hex8, LinearElastic<3>>(p, LinearElastic<3>{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
  } else if (material == TATVA_NEO_HOOKEAN && n_params == 2) {
    if (el == TATVA_TET4) return launch_csr<Tet4, NeoHookean>(p, NeoHookean{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_HEX8) return launch_csr<Hex8, NeoHookean>(p, NeoHookean{prm[0], prm[1]}, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
  } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD && n_params == 5) {
    const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
    if (el == TATVA_TET4) return launch_csr<Tet4, NeoHookeanPhaseField>(p, m, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
    if (el == TATVA_HEX8) return launch_csr<Hex8, NeoHookeanPhaseField>(p, m, d_u, d_indptr, d_pos, nnz, d_data, st, d_indices);
  } else {
    return TATVA_E_INVALID;
  }
  return TATVA_E_UNSUPPORTED;
}

extern "C" {

int tatva_csr_assemble(tatva_plan_t* p, int material, const double* prm, int n_params, const double* d_u,
                       const int32_t* d_indptr, const int32_t* d_pos, int64_t nnz, double* d_data,
                       tatva_stream_t stream) {
  return csr_dispatch(p, material, prm, n_params, d_u, d_indptr, nullptr, d_pos, nnz, d_data, stream);
}
// Symmetric variant: REDs only for the upper triangle, then a mirror pass (needs the column indices).
// Tiled assembly with on-chip combination of duplicate blocks (k_csr_tiled): Tri3 / Tet4 with the neo-Hookean or the
// linear-elastic law.  `d_conn` is the element list the schedule was built for (tatva_host_csr_tile_schedule, tile = 128),
// normally a locality-sorted copy of the plan's connectivity.  TATVA_E_UNSUPPORTED for other (element, law) pairs.
int tatva_csr_assemble_tiled(tatva_plan_t* p, int material, const double* prm, int n_params, const double* d_u,
                             const int32_t* d_conn, const int32_t* d_blk_ptr, const int32_t* d_blk_base,
                             const int32_t* d_blk_rowlen, const int32_t* d_blk_base_t, const int32_t* d_blk_rowlen_t,
                             const int32_t* d_con_ptr, const uint32_t* d_con, int64_t nnz, double* d_data,
                             tatva_stream_t stream) {
  if (!p || !prm || !d_u || !d_conn || !d_blk_ptr || !d_blk_base || !d_blk_rowlen || !d_blk_base_t || !d_blk_rowlen_t || !d_con_ptr || !d_con || !d_data || nnz <= 0) return TATVA_E_INVALID;
  if (p->custom || n_params != 2) return TATVA_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)((p->n_elems + kTile - 1) / kTile);
#define TATVA_TILED(EL, MAT, ...)                                                                                     \
  {                                                                                                                   \
    constexpr size_t smem = (size_t)(((2 * EL::npe * EL::dim + 3) | 1) + EL::dim * EL::dim + 2) * kTile * sizeof(double);     \
    TATVA_CUDA_TRY(cudaMemsetAsync(d_data, 0, sizeof(double) * nnz, st));                                              \
    k_csr_tiled<EL, MAT><<<grid, kTile, smem, st>>>(p->coords, d_conn, p->n_elems, MAT __VA_ARGS__, d_u, d_blk_ptr, d_blk_base, \
                                                    d_blk_rowlen, d_blk_base_t, d_blk_rowlen_t, d_con_ptr, d_con, d_data); \
    TATVA_LAUNCH_CHECK();                                                                                             \
    return TATVA_OK;                                                                                                  \
  }
  if (material == TATVA_NEO_HOOKEAN && p->element == TATVA_TET4) TATVA_TILED(Tet4, NeoHookean, {prm[0], prm[1]})
  if (material == TATVA_LINEAR_ELASTIC && p->element == TATVA_TET4) TATVA_TILED(Tet4, LinearElastic<3>, {prm[0], prm[1]})
  if (material == TATVA_LINEAR_ELASTIC && p->element == TATVA_TRI3) TATVA_TILED(Tri3, LinearElastic<2>, {prm[0], prm[1]})
#undef TATVA_TILED
  return TATVA_E_UNSUPPORTED;
}

int tatva_csr_assemble_sym(tatva_plan_t* p, int material, const double* prm, int n_params, const double* d_u,
                           const int32_t* d_indptr, const int32_t* d_indices, const int32_t* d_pos, int64_t nnz,
                           double* d_data, tatva_stream_t stream) {
  if (!d_indices) return TATVA_E_INVALID;
  return csr_dispatch(p, material, prm, n_params, d_u, d_indptr, d_indices, d_pos, nnz, d_data, stream);
}

}  // extern "C"

template <class El, class Mat>
static int launch_csr_rows(tatva_plan* p, const Mat& mat, const double* u, const int32_t* indptr, const int32_t* indices,
                           const int32_t* n2e_ptr, const int32_t* n2e, double* data, cudaStream_t st) {
  constexpr int warps = 4, SLAB = Mat::dpn * El::npe * Mat::dpn;
  constexpr size_t smem = (size_t)warps * 32 * (SLAB + El::npe / 2 + 1) * sizeof(double);
  static SmemOptIn configured;
  if (smem > 48 * 1024) {
    const int rc = opt_in_smem(k_csr_rows<El, Mat>, smem, configured);
    if (rc != TATVA_OK) return rc;
  }
  const int grid = (int)((p->n_nodes + warps - 1) / warps);
  k_csr_rows<El, Mat><<<grid, warps * 32, smem, st>>>(p->coords, p->conn, p->n_nodes, mat, u, indptr, indices, n2e_ptr, n2e, data);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

extern "C" {

int tatva_csr_assemble_rows(tatva_plan_t* p, int material, const double* prm, int n_params, const double* d_u,
                            const int32_t* d_indptr, const int32_t* d_indices, const int32_t* d_n2e_ptr,
                            const int32_t* d_n2e, double* d_data, tatva_stream_t stream) {
  if (!p || !prm || !d_u || !d_indptr || !d_indices || !d_n2e_ptr || !d_n2e || !d_data) return TATVA_E_INVALID;
  if (p->custom) return TATVA_E_UNSUPPORTED;  // the row-wise kernel carries the default one-point rule
  cudaStream_t st = (cudaStream_t)stream;
  const int el = p->element;
  if (material == TATVA_LINEAR_ELASTIC && n_params == 2) {
    if (el == TATVA_TRI3) return launch_csr_rows<Tri3, LinearElastic<2>>(p, LinearElastic<2>{prm[0], prm[1]}, d_u, d_indptr, d_indices, d_n2e_ptr, d_n2e, d_data, st);
    if (el == TATVA_TET4) return launch_csr_rows<Tet4, LinearElastic<3>>(p, LinearElastic<3>{prm[0], prm[1]}, d_u, d_indptr, d_indices, d_n2e_ptr, d_n2e, d_data, st);
  } else if (material == TATVA_NEO_HOOKEAN && n_params == 2) {
    if (el == TATVA_TET4) return launch_csr_rows<Tet4, NeoHookean>(p, NeoHookean{prm[0], prm[1]}, d_u, d_indptr, d_indices, d_n2e_ptr, d_n2e, d_data, st);
  } else if (material == TATVA_NEO_HOOKEAN_PHASE_FIELD && n_params == 5) {
    const NeoHookeanPhaseField m{prm[0], prm[1], prm[2], prm[3], prm[4]};
    if (el == TATVA_TET4) return launch_csr_rows<Tet4, NeoHookeanPhaseField>(p, m, d_u, d_indptr, d_indices, d_n2e_ptr, d_n2e, d_data, st);
  } else {
    return TATVA_E_INVALID;
  }
  return TATVA_E_UNSUPPORTED;  // multi-point elements: use tatva_csr_assemble
}

int tatva_halo_pack(const double* s, const int64_t* idx, int64_t n, double* d, tatva_stream_t stream) {
  if (n == 0) return TATVA_OK;
  if (!s || !idx || !d || n < 0) return TATVA_E_INVALID;
  k_pack<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s, idx, n, d);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_halo_unpack_set(const double* s, const int64_t* idx, int64_t n, double* d, tatva_stream_t stream) {
  if (n == 0) return TATVA_OK;
  if (!s || !idx || !d || n < 0) return TATVA_E_INVALID;
  k_unpack_set<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s, idx, n, d);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_halo_unpack_add(const double* s, const int64_t* idx, int64_t n, double* d, tatva_stream_t stream) {
  if (n == 0) return TATVA_OK;
  if (!s || !idx || !d || n < 0) return TATVA_E_INVALID;
  k_unpack_add<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s, idx, n, d);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_peer_pull(double* d_x_local, int64_t first_ghost, int64_t n_ghost, const uint64_t* d_peer_ptrs,
                    const int32_t* d_owner, const int64_t* d_owner_idx, tatva_stream_t stream) {
  if (n_ghost == 0) return TATVA_OK;
  if (!d_x_local || !d_peer_ptrs || !d_owner || !d_owner_idx || n_ghost < 0) return TATVA_E_INVALID;
  k_peer_pull<<<grid_for(n_ghost, 256), 256, 0, (cudaStream_t)stream>>>(d_x_local, first_ghost, n_ghost, d_peer_ptrs, d_owner, d_owner_idx);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_peer_push_add(const double* d_y_local, int64_t first_ghost, int64_t n_ghost, const uint64_t* d_peer_ptrs,
                        const int32_t* d_owner, const int64_t* d_owner_idx, tatva_stream_t stream) {
  if (n_ghost == 0) return TATVA_OK;
  if (!d_y_local || !d_peer_ptrs || !d_owner || !d_owner_idx || n_ghost < 0) return TATVA_E_INVALID;
  k_peer_push_add<<<grid_for(n_ghost, 256), 256, 0, (cudaStream_t)stream>>>(d_y_local, first_ghost, n_ghost, d_peer_ptrs, d_owner, d_owner_idx);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_lift(const double* d_u_red, const int64_t* d_src, const double* d_consts, const double* d_base,
               int64_t n_full, double* d_out, tatva_stream_t stream) {
  if (n_full == 0) return TATVA_OK;
  if (!d_src || !d_out || n_full < 0) return TATVA_E_INVALID;
  k_lift<<<grid_for(n_full, 256), 256, 0, (cudaStream_t)stream>>>(d_u_red, d_src, d_consts, d_base, n_full, d_out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_reduce_adjoint(const double* d_r_full, const int64_t* d_ptr, const int64_t* d_list, int64_t n_red,
                         double* d_out, tatva_stream_t stream) {
  if (n_red == 0) return TATVA_OK;
  if (!d_r_full || !d_ptr || !d_list || !d_out || n_red < 0) return TATVA_E_INVALID;
  k_reduce_adjoint<<<grid_for(n_red, 256), 256, 0, (cudaStream_t)stream>>>(d_r_full, d_ptr, d_list, n_red, d_out);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// One CG iteration's vector work, in two calls around the operator application Ap = A p:
//   tatva_cg_after_matvec: s[1] = p.Ap ; alpha = s[0]/s[1] ; x += alpha p ; r -= alpha Ap ; s[2] = r.r ;
//                          beta = s[2]/s[0] ; p = r + beta p ; s[0] = s[2]
// `d_scalars` (>= 8 doubles) and `d_partials` (>= 1184 doubles) are caller-owned device buffers; s[0] must
// hold r.r on entry (tatva_cg_dot(r, r, ..., slot 0)).  Everything is stream-ordered: no host round trip.
int tatva_cg_dot(const double* d_a, const double* d_b, int64_t n, double* d_partials, double* d_scalars, int slot,
                 tatva_stream_t stream) {
  if (!d_a || !d_b || !d_partials || !d_scalars || n <= 0 || slot < 0 || slot > 7) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_cg_dot<<<kCgBlocks, 256, 0, st>>>(d_a, d_b, n, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, slot);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_cg_after_matvec(double* d_x, double* d_r, double* d_p, const double* d_Ap, int64_t n, double* d_partials,
                          double* d_scalars, tatva_stream_t stream) {
  if (!d_x || !d_r || !d_p || !d_Ap || !d_partials || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_cg_dot<<<kCgBlocks, 256, 0, st>>>(d_p, d_Ap, n, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 1);
  k_cg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, n, d_scalars, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 2);
  k_cg_direction<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, n, d_scalars);
  k_cg_roll<<<1, 1, 0, st>>>(d_scalars);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// Jacobi-preconditioned CG (M = diag(A), `d_minv` = 1/diag).  tatva_pcg_start: p = M^-1 r, s[0] = r.M^-1 r.
// tatva_pcg_after_matvec: as tatva_cg_after_matvec with z = M^-1 r folded in (z is never stored); leaves
// s[0] = r.z and s[4] = r.r of the new residual.  `d_scalars` >= 8 doubles, `d_partials` >= 2 * 1184 doubles.
int tatva_pcg_reciprocal(const double* d_diag, int64_t n, double* d_minv, tatva_stream_t stream) {
  if (!d_diag || !d_minv || n <= 0) return TATVA_E_INVALID;
  k_safe_reciprocal<<<kCgBlocks, 256, 0, (cudaStream_t)stream>>>(d_diag, n, d_minv);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_pcg_start(double* d_p, const double* d_r, const double* d_minv, int64_t n, double* d_partials,
                    double* d_scalars, tatva_stream_t stream) {
  if (!d_p || !d_r || !d_minv || !d_partials || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_pcg_start<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, d_minv, n, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 0);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_pcg_after_matvec(double* d_x, double* d_r, double* d_p, const double* d_Ap, const double* d_minv, int64_t n,
                           double* d_partials, double* d_scalars, tatva_stream_t stream) {
  if (!d_x || !d_r || !d_p || !d_Ap || !d_minv || !d_partials || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_cg_dot<<<kCgBlocks, 256, 0, st>>>(d_p, d_Ap, n, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 1);
  k_pcg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, d_minv, n, d_scalars, d_partials);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 2);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials + kCgBlocks, kCgBlocks, d_scalars, 4);
  k_pcg_direction<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, d_minv, n, d_scalars);
  k_cg_roll<<<1, 1, 0, st>>>(d_scalars);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// The same iteration split at its two dot products, for a distributed CG: the caller all-reduces d_scalars[1] after
// tatva_cg_dot(p, Ap, slot 1) and d_scalars[2] after tatva_cg_update, on the same stream, then calls
// tatva_cg_direction.  d_minv may be NULL (plain CG); with d_minv, s[0]/s[2] hold r.z and s[4] holds r.r.
int tatva_cg_update(double* d_x, double* d_r, const double* d_p, const double* d_Ap, const double* d_minv, int64_t n,
                    double* d_partials, double* d_scalars, tatva_stream_t stream) {
  if (!d_x || !d_r || !d_p || !d_Ap || !d_partials || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_minv) {
    k_pcg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, d_minv, n, d_scalars, d_partials);
    k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 2);
    k_cg_finish<<<1, 256, 0, st>>>(d_partials + kCgBlocks, kCgBlocks, d_scalars, 4);
  } else {
    k_cg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, n, d_scalars, d_partials);
    k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 2);
  }
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
int tatva_cg_direction(double* d_p, const double* d_r, const double* d_minv, int64_t n, double* d_scalars,
                       tatva_stream_t stream) {
  if (!d_p || !d_r || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_minv) k_pcg_direction<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, d_minv, n, d_scalars);
  else k_cg_direction<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, n, d_scalars);
  k_cg_roll<<<1, 1, 0, st>>>(d_scalars);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

// One CG iteration AFTER an operator application that already left p.Ap in d_scalars[1] (tatva_hvp_lifted_dot with
// roll = 1): alpha = s0/s1; x += alpha p; r -= alpha Ap; s2 = r.r (and s4 with d_minv); beta = s2/s0; p = r + beta p (or
// M^-1 r); d_zero (may be NULL) is cleared in the same pass for the next application.  s0 <- s2 is left to the next
// tatva_hvp[_lifted]_dot(..., roll = 1).  d_fixed_map (may be NULL; n int32, < 0 = Fixed DOF, e.g. Lifter.dof_map) runs the
// iteration on FULL-size vectors with the Dirichlet rows of r kept at zero, so that the unconstrained HVP kernel can be
// used instead of the lifted one when a lifter holds only Fixed constraints.
int tatva_cg_after_dot(double* d_x, double* d_r, double* d_p, const double* d_Ap, const double* d_minv, double* d_zero,
                       const int32_t* d_fixed_map, int64_t n, double* d_partials, double* d_scalars, tatva_stream_t stream) {
  if (!d_x || !d_r || !d_p || !d_Ap || !d_partials || !d_scalars || n <= 0) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_fixed_map) k_cg_update_masked<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, d_minv, d_fixed_map, n, d_scalars, d_partials);
  else if (d_minv) k_pcg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, d_minv, n, d_scalars, d_partials);
  else k_cg_update<<<kCgBlocks, 256, 0, st>>>(d_x, d_r, d_p, d_Ap, n, d_scalars, d_partials);
  if (d_minv) k_cg_finish<<<1, 256, 0, st>>>(d_partials + kCgBlocks, kCgBlocks, d_scalars, 4);
  k_cg_finish<<<1, 256, 0, st>>>(d_partials, kCgBlocks, d_scalars, 2);
  k_cg_direction_zero<<<kCgBlocks, 256, 0, st>>>(d_p, d_r, d_minv, n, d_scalars, d_zero);
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}

int tatva_fp64_peak_tflops(double* tflops, tatva_stream_t stream) {
  if (!tflops) return TATVA_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  TATVA_CUDA_TRY(cudaGetDevice(&dev));
  TATVA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* buf = nullptr;
  TATVA_CUDA_TRY(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_dfma<<<blocks, threads, 0, st>>>(buf, 64);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, st);
    k_dfma<<<blocks, threads, 0, st>>>(buf, iters);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(buf);
  TATVA_LAUNCH_CHECK();
  const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  return TATVA_OK;
}

}  // extern "C"
