// Warp-cooperative fused residual / HVP kernels (r02): gather through warp shuffles, per-tile node sums before the REDs.
// Header because two translation units instantiate it (generic.cu: any law through GenericBody; neo_hookean.cu: the
// reference-space Tet4 arithmetic).
#pragma once
#include "fused.cuh"

namespace tatva {

// ---- warp-cooperative fused residual / HVP (r02) -------------------------------------------------------------------
// The element-per-thread kernels of the one-point simplices are bound by L1 wavefronts, not by FP64 or bytes (ncu, Tet4
// HVP: l1tex data-pipe wavefronts 78 %, FP64 27 %): every lane gathers its own nodal rows (36 scattered 8-byte loads per
// tet, 12 sectors per request) although the 32 elements of a warp share ~28 distinct nodes, and every lane scatters its
// own 12 REDs although a 128-element tile holds ~96 distinct nodes for 512 node references.  Here
//   gather   (GW) lane l loads the rows of the warp's l-th distinct node ONCE (plan-time list, ascending ids: runs of
//            neighbouring rows) and the elements fetch their npe rows with warp shuffles (no bank conflicts); a node that
//            did not fit among the 32 is read by the element itself (`warp_local` = 255);
//   scatter  (SW) the tile's nodal contributions are parked in shared memory (element-major, odd stride); the distinct
//            nodes of the tile are walked in chunks of 32 (one per lane, by decreasing contributor count), each lane sums
//            its node's contributors (plan-time table, one coalesced row of 32 entries per step, staged in shared memory
//            with cp.async while the elements compute) and the sums are re-dealt so that dpn consecutive lanes add the
//            dpn consecutive doubles of one node: one RED per (distinct node, DOF) instead of one per (element, node, DOF).
// GW / SW = false fall back to the element's own gather / the per-warp sector-grouped scatter (A/B measurements).
constexpr int kWcEllStage = 2048;  // contributor entries of a tile staged in shared memory (larger tables are read in place)

TATVA_D void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

template <int NPE, int DPN, bool SW>
TATVA_HD constexpr size_t wc_smem_bytes() {
  constexpr int S = NPE * DPN, SP = S | 1;
  if (!SW) return grouped_scatter_smem<NPE, DPN>(kBlock / 32);
  return (((size_t)(kBlock + 1) * SP + (kBlock / 32) * 32 * DPN) * sizeof(double) + (kBlock / 32) * 32 * sizeof(int) + 15) / 16 * 16 + kWcEllStage * sizeof(uint16_t);
}

template <class El, class Mat, int MODE, class Body, bool GW = true, bool SW = true>
__global__ void __launch_bounds__(kBlock, Body::min_ctas) k_fused_wc(const double* __restrict__ coords, const int32_t* __restrict__ conn,
                                                     int64_t E, Mat mat, const double* __restrict__ u,
                                                     const double* __restrict__ v, double* __restrict__ y,
                                                     const int32_t* __restrict__ warp_nodes, const uint8_t* __restrict__ warp_local,
                                                     const int4* __restrict__ tile_hdr, const int32_t* __restrict__ tn_node,
                                                     const int32_t* __restrict__ ell_ptr, const uint16_t* __restrict__ ell) {
  static_assert(MODE == MODE_RESIDUAL || MODE == MODE_HVP, "scatter-add modes only");
  static_assert(kBlock == 128, "the node schedule is cut into tiles of 128 elements");
  constexpr int D = El::dim, NPE = El::npe, dpn = Mat::dpn, S = NPE * dpn, SP = S | 1;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ double sm_wc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sY = sm_wc;                                                                                    // [kBlock + 1][SP], row kBlock = 0
  double* wbuf = sY + (kBlock + 1) * SP + warp * (32 * dpn);                                             // [32 * dpn] per warp
  int* snode = reinterpret_cast<int*>(sY + (kBlock + 1) * SP + (kBlock / 32) * 32 * dpn) + warp * 32;    // [32] per warp
  constexpr size_t kEllOff = (((size_t)(kBlock + 1) * SP + (kBlock / 32) * 32 * dpn) * sizeof(double) + (kBlock / 32) * 32 * sizeof(int) + 15) / 16 * 16;  // cp.async moves 16-byte pieces
  uint16_t* sEll = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(sm_wc) + kEllOff);
  if constexpr (SW) {
    if (threadIdx.x < SP) sY[kBlock * SP + threadIdx.x] = 0.0;  // the row the empty table entries point at
  }
  // One CTA per tile.  (A resident grid walking the tiles with the next tile's index data prefetched was measured slower,
  // 0.049-0.050 vs 0.045-0.047 ms at config 2: the prefetched registers spill and a spill store waits for its load.)
  {
    const int64_t tile = blockIdx.x;
    const int64_t e = tile * kBlock + threadIdx.x;
    const bool valid = e < E;
    int4 hdr = make_int4(0, 0, 0, 0);  // {first chunk, chunks, first table entry, table entries}
    if constexpr (SW) hdr = __ldg(tile_hdr + tile);
    int my = -1;
    unsigned locw = 0;
    if constexpr (GW) {
      my = __ldg(warp_nodes + (tile * (kBlock / 32) + warp) * 32 + lane);
      if constexpr (NPE == 4) locw = valid ? __ldg(reinterpret_cast<const unsigned*>(warp_local) + e) : 0u;
    }
    // -- scatter table of the tile: on its way into shared memory while the elements compute
    const int ch0 = hdr.x, ch1 = hdr.x + hdr.y, ebase = hdr.z, ecount = hdr.w;
    int node_first = -1, o0_first = 0, o1_first = 0;
    if constexpr (SW) {
      if (ecount <= kWcEllStage)
        for (int i = threadIdx.x * 8; i < ecount; i += kBlock * 8) cp_async16(sEll + i, ell + ebase + i);  // rows of 32 entries: 16-byte pieces
      asm volatile("cp.async.commit_group;" ::);
      if (ch0 + warp < ch1) {  // this warp's first chunk: fetched now, used after the element arithmetic
        node_first = __ldg(tn_node + (int64_t)(ch0 + warp) * 32 + lane);
        o0_first = __ldg(ell_ptr + ch0 + warp) - ebase;
        o1_first = __ldg(ell_ptr + ch0 + warp + 1) - ebase;
      }
    }
    double X[NPE][D], U[NPE][dpn], V[NPE][dpn];
    int nd[NPE];
    if constexpr (GW) {
      // -- gather: one distinct node per lane, the elements pick their rows with shuffles
      double nX[D], nU[dpn], nV[dpn];
#pragma unroll
      for (int c = 0; c < D; ++c) nX[c] = 0.0;
#pragma unroll
      for (int c = 0; c < dpn; ++c) nU[c] = nV[c] = 0.0;
      if (my >= 0) {
        load_row<D>(coords, my, nX);
        load_row<dpn>(u, my, nU);
        if constexpr (MODE == MODE_HVP) load_row<dpn>(v, my, nV);
      }
      int loc[NPE];
      if constexpr (NPE == 4) {
#pragma unroll
        for (int n = 0; n < 4; ++n) loc[n] = (locw >> (8 * n)) & 255;
      } else {
#pragma unroll
        for (int n = 0; n < NPE; ++n) loc[n] = valid ? (int)__ldg(warp_local + e * NPE + n) : 0;
      }
      bool direct = false;
#pragma unroll
      for (int n = 0; n < NPE; ++n) direct |= loc[n] == 255;
#pragma unroll
      for (int n = 0; n < NPE; ++n) {
        const int src = loc[n] & 31;
#pragma unroll
        for (int c = 0; c < D; ++c) X[n][c] = __shfl_sync(kFull, nX[c], src);
#pragma unroll
        for (int c = 0; c < dpn; ++c) {
          U[n][c] = __shfl_sync(kFull, nU[c], src);
          if constexpr (MODE == MODE_HVP) V[n][c] = __shfl_sync(kFull, nV[c], src);
        }
      }
      const bool any_direct = __any_sync(kFull, direct);
      if (any_direct || !SW) {  // more than 32 distinct nodes in this warp: the overflow rows come straight from memory
        if (valid && (direct || !SW)) load_conn<El>(conn, e, nd);
        if (direct) {
#pragma unroll
          for (int n = 0; n < NPE; ++n)
            if (loc[n] == 255) {
              load_row<D>(coords, nd[n], X[n]);
              load_row<dpn>(u, nd[n], U[n]);
              if constexpr (MODE == MODE_HVP) load_row<dpn>(v, nd[n], V[n]);
            }
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < NPE; ++n) nd[n] = 0;
      if (valid) {
        load_conn<El>(conn, e, nd);
        gather_rows(coords, nd, X);
        gather_rows(u, nd, U);
        if constexpr (MODE == MODE_HVP) gather_rows(v, nd, V);
      }
    }
    // -- element arithmetic
    double Y[NPE][dpn];
#pragma unroll
    for (int n = 0; n < NPE; ++n)
#pragma unroll
      for (int c = 0; c < dpn; ++c) Y[n][c] = 0.0;
    if (valid) Body::template run<MODE>(mat, X, U, V, Y);
    if constexpr (!SW) {
      if (!valid) {
#pragma unroll
        for (int n = 0; n < NPE; ++n) nd[n] = 0;
      }
      grouped_scatter<NPE, dpn>(y, nd, Y, valid, sm_wc + (size_t)warp * grouped_scatter_words<NPE, dpn>());
    } else {
      // -- scatter: per-tile sums of the distinct nodes
      {
        double* row = sY + threadIdx.x * SP;
#pragma unroll
        for (int n = 0; n < NPE; ++n)
#pragma unroll
          for (int c = 0; c < dpn; ++c) row[n * dpn + c] = Y[n][c];
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      asm volatile("griddepcontrol.wait;" ::: "memory");  // y may still be being cleared (launch_behind_zero)
      const bool staged = ecount <= kWcEllStage;
      for (int ch = ch0 + warp; ch < ch1; ch += kBlock / 32) {
        int node = node_first, o0 = o0_first, o1 = o1_first;
        if (ch != ch0 + warp) {
          node = __ldg(tn_node + (int64_t)ch * 32 + lane);
          o0 = __ldg(ell_ptr + ch) - ebase;
          o1 = __ldg(ell_ptr + ch + 1) - ebase;
        }
        double acc[dpn];
#pragma unroll
        for (int k = 0; k < dpn; ++k) acc[k] = 0.0;
        // branch-free: an empty entry (0xFFFF) is clamped onto the zero row, so the loads of several steps can be in flight
#pragma unroll 4
        for (int o = o0; o < o1; o += 32) {
          int src = staged ? (int)sEll[o + lane] : (int)__ldg(ell + ebase + o + lane);
          src = src < (kBlock << 3) ? src : (kBlock << 3);
          const double* row = sY + (src >> 3) * SP + (src & 7) * dpn;
#pragma unroll
          for (int k = 0; k < dpn; ++k) acc[k] += row[k];
        }
        snode[lane] = node;
#pragma unroll
        for (int k = 0; k < dpn; ++k) wbuf[lane * dpn + k] = acc[k];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < dpn; ++j) {
          const int q = lane + 32 * j, owner = q / dpn, comp = q - owner * dpn;
          const int nn = snode[owner];
          if (nn >= 0) atomicAdd(y + (int64_t)nn * dpn + comp, wbuf[q]);
        }
        __syncwarp();
      }
    }
  }
}

// The arithmetic of k_fused for one element (any element / law pair): Y += sum_q W (first | second variation) . dN
template <class El, class Mat>
struct GenericBody {
  static constexpr int min_ctas = Mat::dpn > El::dim ? 3 : 4;  // register caps: 168 for a two-field law, 128 otherwise
  template <int MODE>
  TATVA_D static void run(const Mat& mat, const double (&X)[El::npe][El::dim], const double (&U)[El::npe][Mat::dpn],
                          const double (&V)[El::npe][Mat::dpn], double (&Y)[El::npe][Mat::dpn]) {
    constexpr int NPE = El::npe, dpn = Mat::dpn;
#pragma unroll 1
    for (int q = 0; q < El::num_q(); ++q) {
      double dNdX[El::dim][El::npe], N[El::npe];
      const double W = geometry<El>(q, X, dNdX) * El::weight(q);
      El::N(q, N);
      typename Mat::S s, ds, f;
      typename Mat::Cache cache;
      qp_state<El, Mat>(dNdX, N, U, s);
      mat.prepare(s, cache);
      if constexpr (MODE == MODE_RESIDUAL) {
        mat.first(s, cache, f);
      } else {
        qp_state<El, Mat>(dNdX, N, V, ds);
        mat.second(s, cache, ds, f);
      }
#pragma unroll
      for (int n = 0; n < NPE; ++n)
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
};

#ifndef __CUDACC_RTC__
template <class El, class Mat, int MODE, class Body, bool GW = true, bool SW = true>
static int launch_fused_wc(const tatva_plan* p, const Mat& mat, const double* u, const double* v, double* out, cudaStream_t st) {
  constexpr size_t smem = wc_smem_bytes<El::npe, Mat::dpn, SW>();
  static_assert(smem <= 48 * 1024, "tile staging exceeds the default shared-memory window");
  const int rc = launch_behind_zero(k_fused_wc<El, Mat, MODE, Body, GW, SW>, grid_for(p->n_elems), kBlock, smem, st, p->zero_output != 0, out, p->n_nodes * Mat::dpn,
                                    p->coords, p->conn, p->n_elems, mat, u, v, out, p->ws_warp_nodes, p->ws_warp_local, reinterpret_cast<const int4*>(p->ws_tile_hdr),
                                    p->ws_tn_node, p->ws_ell_ptr, p->ws_ell);
  if (rc != TATVA_OK) return rc;
  TATVA_LAUNCH_CHECK();
  return TATVA_OK;
}
#endif

}  // namespace tatva
