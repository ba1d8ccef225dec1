// Host-side setup for the coloured sparse-Jacobian path: sparsity pattern, distance-2 colouring and
// the element -> CSR position map.  Integer work that runs once per mesh; results are bit-exact with
// the reference's NumPy/SciPy code (tatva/sparse/_extraction.py:37-102, tatva/sparse/_coloring.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <utility>
#include <vector>

#include "../../include/tatva_b200.h"

namespace {

// node -> sorted unique neighbour nodes (nodes sharing an element, self included)
void node_adjacency(const int32_t* conn, int64_t n_elems, int npe, int64_t n_nodes, std::vector<int64_t>& ptr,
                    std::vector<int32_t>& adj) {
  std::vector<int64_t> n2e_ptr(n_nodes + 1, 0);
  for (int64_t i = 0; i < n_elems * npe; ++i) n2e_ptr[conn[i] + 1]++;
  for (int64_t n = 0; n < n_nodes; ++n) n2e_ptr[n + 1] += n2e_ptr[n];
  std::vector<int64_t> n2e(n2e_ptr[n_nodes]);
  {
    std::vector<int64_t> fill(n2e_ptr.begin(), n2e_ptr.end() - 1);
    for (int64_t e = 0; e < n_elems; ++e)
      for (int a = 0; a < npe; ++a) n2e[fill[conn[e * npe + a]]++] = e;
  }
  ptr.assign(n_nodes + 1, 0);
  std::vector<std::vector<int32_t>> rows(n_nodes);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t n = 0; n < n_nodes; ++n) {
    std::vector<int32_t>& r = rows[n];
    r.reserve((n2e_ptr[n + 1] - n2e_ptr[n]) * npe);
    for (int64_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
      const int32_t* el = conn + n2e[k] * npe;
      r.insert(r.end(), el, el + npe);
    }
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
  }
  for (int64_t n = 0; n < n_nodes; ++n) ptr[n + 1] = ptr[n] + (int64_t)rows[n].size();
  adj.resize(ptr[n_nodes]);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_nodes; ++n) std::copy(rows[n].begin(), rows[n].end(), adj.begin() + ptr[n]);
}

}  // namespace

extern "C" {

// tatva/sparse/_extraction.py:37-102.  DOF id = node * dpn + comp (:57-60); all (row, col) pairs of every
// element, unique + sorted by row * ncols + col (:74-79) == per row, the sorted DOFs of the sorted
// neighbour nodes.  Rows of nodes that appear in no element are empty.
int tatva_host_pattern_from_mesh(const int32_t* conn, int64_t n_elems, int npe, int64_t n_nodes, int dpn,
                                 int32_t* indptr, int32_t* indices, int64_t* nnz_out) {
  if (!conn || !indptr || !nnz_out || n_elems < 0 || npe <= 0 || n_nodes <= 0 || dpn <= 0) return TATVA_E_INVALID;
  for (int64_t i = 0; i < n_elems * npe; ++i)
    if (conn[i] < 0 || conn[i] >= n_nodes) return TATVA_E_INVALID;
  std::vector<int64_t> ptr;
  std::vector<int32_t> adj;
  node_adjacency(conn, n_elems, npe, n_nodes, ptr, adj);
  const int64_t nnz = ptr[n_nodes] * dpn * dpn;
  *nnz_out = nnz;
  if (nnz > INT32_MAX) return TATVA_E_INVALID;  // scipy would switch to int64 indices here
  indptr[0] = 0;
  for (int64_t n = 0; n < n_nodes; ++n) {
    const int64_t w = (ptr[n + 1] - ptr[n]) * dpn;
    for (int i = 0; i < dpn; ++i) indptr[n * dpn + i + 1] = (int32_t)(indptr[n * dpn + i] + w);
  }
  if (!indices) return TATVA_OK;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_nodes; ++n) {
    for (int i = 0; i < dpn; ++i) {
      int32_t* out = indices + indptr[n * dpn + i];
      for (int64_t k = ptr[n]; k < ptr[n + 1]; ++k)
        for (int c = 0; c < dpn; ++c) *out++ = adj[k] * dpn + c;
    }
  }
  return TATVA_OK;
}

// General pattern of tatva/sparse/_extraction.py:118-245 (pattern_from_compound): every element couples all the
// DOFs listed in its row of `elem_dofs` (n_elems, width; -1 = absent), and the DOFs in `diag` (n_diag) get their
// diagonal entry.  Result = sorted unique (row, col) pairs as CSR, identical to the reference's
// np.unique(row * n + col).  Two-call protocol: indices == NULL fills indptr and *nnz only.
int tatva_host_pattern_from_element_dofs(const int32_t* elem_dofs, int64_t n_elems, int width, const int32_t* diag,
                                         int64_t n_diag, int64_t n, int32_t* indptr, int32_t* indices, int64_t* nnz_out) {
  if (!indptr || !nnz_out || n <= 0 || n_elems < 0 || width < 0 || n_diag < 0 || (n_elems * width > 0 && !elem_dofs) ||
      (n_diag > 0 && !diag))
    return TATVA_E_INVALID;
  const int64_t total = n_elems * width;
  std::vector<int64_t> d2e_ptr(n + 1, 0);
  for (int64_t i = 0; i < total; ++i) {
    const int32_t d = elem_dofs[i];
    if (d >= n) return TATVA_E_INVALID;
    if (d >= 0) d2e_ptr[d + 1]++;
  }
  for (int64_t d = 0; d < n; ++d) d2e_ptr[d + 1] += d2e_ptr[d];
  std::vector<int32_t> d2e(d2e_ptr[n]);
  {
    std::vector<int64_t> fill(d2e_ptr.begin(), d2e_ptr.end() - 1);
    for (int64_t e = 0; e < n_elems; ++e)
      for (int a = 0; a < width; ++a) {
        const int32_t d = elem_dofs[e * width + a];
        if (d >= 0) d2e[fill[d]++] = (int32_t)e;
      }
  }
  std::vector<char> on_diag(n, 0);
  for (int64_t i = 0; i < n_diag; ++i) {
    if (diag[i] < 0 || diag[i] >= n) return TATVA_E_INVALID;
    on_diag[diag[i]] = 1;
  }
  std::vector<std::vector<int32_t>> rows(n);
#pragma omp parallel for schedule(dynamic, 2048)
  for (int64_t d = 0; d < n; ++d) {
    std::vector<int32_t>& r = rows[d];
    r.reserve((d2e_ptr[d + 1] - d2e_ptr[d]) * width + 1);
    int32_t last = -1;
    for (int64_t k = d2e_ptr[d]; k < d2e_ptr[d + 1]; ++k) {
      if (d2e[k] == last) continue;  // an element listing the DOF twice appears twice in a row: skip the repeat
      last = d2e[k];
      const int32_t* el = elem_dofs + (int64_t)last * width;
      for (int a = 0; a < width; ++a)
        if (el[a] >= 0) r.push_back(el[a]);
    }
    if (on_diag[d]) r.push_back((int32_t)d);
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
  }
  int64_t nnz = 0;
  for (int64_t d = 0; d < n; ++d) nnz += (int64_t)rows[d].size();
  *nnz_out = nnz;
  if (nnz > INT32_MAX) return TATVA_E_INVALID;
  indptr[0] = 0;
  for (int64_t d = 0; d < n; ++d) indptr[d + 1] = indptr[d] + (int32_t)rows[d].size();
  if (!indices) return TATVA_OK;
#pragma omp parallel for schedule(static)
  for (int64_t d = 0; d < n; ++d) std::copy(rows[d].begin(), rows[d].end(), indices + indptr[d]);
  return TATVA_OK;
}

// Block structure of a mesh pattern: the b rows of a node are identical and made of full, aligned b-wide column
// blocks, the diagonal block included (pattern_from_mesh with b DOFs per node; rows of nodes that belong to no element
// are empty).  Returns the largest such b in [2, 8], or 1.
static int detect_block_size(const int32_t* indptr, const int32_t* indices, int64_t n) {
  for (int b = 8; b >= 2; --b) {
    if (n % b) continue;
    bool ok = true;
    for (int64_t r = 0; r < n && ok; r += b) {
      const int32_t p0 = indptr[r], len = indptr[r + 1] - p0;
      if (len % b) ok = false;
      for (int j = 1; j < b && ok; ++j)
        ok = (indptr[r + j + 1] - indptr[r + j] == len) &&
             std::equal(indices + p0, indices + p0 + len, indices + indptr[r + j]);
      bool diagonal = (len == 0);
      for (int32_t k = 0; k < len && ok; k += b) {
        const int32_t c0 = indices[p0 + k];
        ok = (c0 % b == 0);
        diagonal |= (c0 == r);
        for (int j = 1; j < b && ok; ++j) ok = indices[p0 + k + j] == c0 + j;
      }
      ok = ok && diagonal;
    }
    if (ok) return b;
  }
  return 1;
}

// tatva/sparse/_coloring.py:270-283 -> :27-48 (pattern of A@A) -> :136-153 (first-fit, natural order).
// The squared graph is never materialised: the distance-2 neighbours of i are the columns of the rows
// named by row i.  `stamp[c] == i` marks colour c as used by a neighbour of i.
int tatva_host_distance2_colors(const int32_t* indptr, const int32_t* indices, int64_t n, int32_t* colors,
                                int32_t* n_colors) {
  if (!indptr || !indices || !colors || n <= 0) return TATVA_E_INVALID;
  std::fill(colors, colors + n, -1);
  std::vector<int64_t> stamp;
  stamp.reserve(256);
  int32_t maxc = -1;
  const int b = detect_block_size(indptr, indices, n);
  if (b > 1) {
    // Node-level sweep, identical to the per-DOF greedy: the b DOFs of a node see the same distance-2 set, and between
    // two of them only the node's own previous DOF gets coloured, so the forbidden set is built once per node and
    // the next DOF continues the search above the colour just given.  (b^2 fewer pattern entries are visited.)
    const int64_t n_nodes = n / b;
    for (int64_t r = 0; r < n_nodes; ++r) {
      const int32_t p0 = indptr[r * b], p1 = indptr[r * b + 1];
      if (p0 == p1) {  // node in no element: its DOFs have no neighbours at all, each takes colour 0
        for (int j = 0; j < b; ++j) colors[r * b + j] = 0;
        if (maxc < 0) maxc = 0;
        continue;
      }
      for (int32_t a = p0; a < p1; a += b) {
        const int64_t m = indices[a] / b;  // neighbour node
        const int32_t q0 = indptr[m * b], q1 = indptr[m * b + 1];
        for (int32_t t = q0; t < q1; t += b) {
          const int32_t* cc = colors + indices[t];  // the b DOFs of a node two hops away
          for (int j = 0; j < b; ++j) {
            const int32_t c = cc[j];
            if (c >= 0) {
              if ((size_t)c >= stamp.size()) stamp.resize(c + 1, -1);
              stamp[c] = r;
            }
          }
        }
      }
      int32_t c = 0;
      for (int j = 0; j < b; ++j) {
        while ((size_t)c < stamp.size() && stamp[c] == r) ++c;
        colors[r * b + j] = c;
        if (c > maxc) maxc = c;
        ++c;  // the colour just used is now forbidden, and every smaller one already was
      }
    }
    if (n_colors) *n_colors = maxc + 1;
    return TATVA_OK;
  }
  for (int64_t i = 0; i < n; ++i) {
    for (int32_t a = indptr[i]; a < indptr[i + 1]; ++a) {
      const int32_t k = indices[a];
      for (int32_t bb = indptr[k]; bb < indptr[k + 1]; ++bb) {
        const int32_t c = colors[indices[bb]];
        if (c >= 0) {
          if ((size_t)c >= stamp.size()) stamp.resize(c + 1, -1);
          stamp[c] = i;
        }
      }
    }
    int32_t c = 0;
    while ((size_t)c < stamp.size() && stamp[c] == i) ++c;
    colors[i] = c;
    if (c > maxc) maxc = c;
  }
  if (n_colors) *n_colors = maxc + 1;
  return TATVA_OK;
}

// node -> incident elements (CSR), elements ascending; two-call protocol (list == NULL: fill ptr only)
int tatva_host_node_to_elements(const int32_t* conn, int64_t n_elems, int npe, int64_t n_nodes, int32_t* ptr,
                                int32_t* list) {
  if (!conn || !ptr || n_elems <= 0 || npe <= 0 || n_nodes <= 0) return TATVA_E_INVALID;
  if (n_elems * npe > INT32_MAX) return TATVA_E_INVALID;
  std::fill(ptr, ptr + n_nodes + 1, 0);
  for (int64_t i = 0; i < n_elems * npe; ++i) {
    if (conn[i] < 0 || conn[i] >= n_nodes) return TATVA_E_INVALID;
    ptr[conn[i] + 1]++;
  }
  for (int64_t n = 0; n < n_nodes; ++n) ptr[n + 1] += ptr[n];
  if (!list) return TATVA_OK;
  std::vector<int32_t> fill(ptr, ptr + n_nodes);
  for (int64_t e = 0; e < n_elems; ++e)
    for (int a = 0; a < npe; ++a) {
      const int32_t n = conn[e * npe + a];
      // an element listing the same node twice contributes once
      if (fill[n] > ptr[n] && list[fill[n] - 1] == (int32_t)e) continue;
      list[fill[n]++] = (int32_t)e;
    }
  // compact (only needed if some element repeated a node)
  bool compact = false;
  for (int64_t n = 0; n < n_nodes && !compact; ++n) compact = fill[n] != ptr[n + 1];
  if (compact) {
    int32_t w = 0;
    std::vector<int32_t> np(n_nodes + 1, 0);
    for (int64_t n = 0; n < n_nodes; ++n) {
      for (int32_t k = ptr[n]; k < fill[n]; ++k) list[w++] = list[k];
      np[n + 1] = w;
    }
    std::copy(np.begin(), np.end(), ptr);
  }
  return TATVA_OK;
}

// Shared-memory staging tiles: consecutive groups of `tile_elems` elements (one CTA each), the sorted unique
// nodes each group touches and the connectivity re-expressed in tile-local node indices.  Two-call protocol:
// tile_nodes == NULL fills tile_ptr (n_tiles + 1) and *max_unique only.
int tatva_host_build_tiles(const int32_t* conn, int64_t n_elems, int npe, int tile_elems, int32_t* tile_ptr,
                           int32_t* tile_nodes, uint16_t* local_conn, int32_t* max_unique) {
  if (!conn || !tile_ptr || !max_unique || n_elems <= 0 || npe <= 0 || tile_elems <= 0) return TATVA_E_INVALID;
  const int64_t n_tiles = (n_elems + tile_elems - 1) / tile_elems;
  std::vector<int32_t> counts(n_tiles, 0);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t t = 0; t < n_tiles; ++t) {
    const int64_t e0 = t * tile_elems, e1 = std::min(n_elems, e0 + tile_elems);
    std::vector<int32_t> nodes(conn + e0 * npe, conn + e1 * npe);
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
    counts[t] = (int32_t)nodes.size();
    if (nodes.size() > 65535) bad |= 1;
    if (tile_nodes) {
      int32_t* dst = tile_nodes + tile_ptr[t];
      std::copy(nodes.begin(), nodes.end(), dst);
      for (int64_t i = e0 * npe; i < e1 * npe; ++i)
        local_conn[i] = (uint16_t)(std::lower_bound(nodes.begin(), nodes.end(), conn[i]) - nodes.begin());
    }
  }
  if (bad) return TATVA_E_INVALID;
  if (!tile_nodes) {
    tile_ptr[0] = 0;
    int32_t mx = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      tile_ptr[t + 1] = tile_ptr[t] + counts[t];
      mx = std::max(mx, counts[t]);
    }
    *max_unique = mx;
  }
  return TATVA_OK;
}

// Node schedule of the warp-cooperative fused kernels (k_fused_wc, r02): the element list is cut into warps of 32 and
// tiles of 128 consecutive elements (one CTA of 4 warps).
//   gather   warp_nodes [n_tiles * 128]  per warp of 32 elements its (up to) 32 most referenced distinct nodes, ascending,
//                                        -1 = unused lane: lane l loads the nodal rows of warp_nodes[32 w + l] ONCE
//            warp_local [n_elems * npe]  per element node: the lane that holds it, or 255 = not among the 32 (the
//                                        element then reads that row itself through the connectivity)
//   scatter  the distinct nodes of a tile by DECREASING contributor count, in chunks of 32 (one warp pass each):
//            ch_ptr  [n_tiles + 1]       first chunk of each tile
//            tn_node [32 * n_chunks]     node of (chunk, lane), -1 = none
//            ell_ptr [n_chunks + 1]      first contributor entry of each chunk (a multiple of 32)
//            ell     [n_ell]             entry [ell_ptr[k] + 32 r + lane] = r-th contributor of the lane's node:
//                                        (element slot in the tile) << 3 | local node, 0xFFFF = none; a chunk holds
//                                        32 x (largest count in the chunk) entries, so a warp reads them coalesced and
//                                        its lanes loop equally long
// `cap` > 0 cuts the contributor list of a node into entries of at most `cap` (each entry ends in its own atomic add): a
// tile's heaviest nodes otherwise keep one warp looping long after the others are done.
// Two calls (same cap): with tn_node == NULL ch_ptr, *n_chunks and *n_ell are filled (and warp_nodes / warp_local, if given); the
// second call fills tn_node, ell_ptr and ell.
int tatva_host_node_schedule(const int32_t* conn, int64_t n_elems, int npe, int cap, int32_t* warp_nodes, uint8_t* warp_local,
                             int32_t* ch_ptr, int64_t* n_chunks, int64_t* n_ell, int32_t* tn_node, int32_t* ell_ptr,
                             uint16_t* ell) {
  constexpr int kTileElems = 128;
  if (!conn || !ch_ptr || !n_chunks || !n_ell || n_elems <= 0 || npe <= 0 || npe > 8 || cap < 0) return TATVA_E_INVALID;
  if (tn_node && (!ell_ptr || !ell)) return TATVA_E_INVALID;
  const int64_t n_tiles = (n_elems + kTileElems - 1) / kTileElems;
  const int64_t n_warps = n_tiles * (kTileElems / 32);  // the last CTA reads the lists of all its warps
  if (n_elems * npe > INT32_MAX / 64) return TATVA_E_INVALID;
  if (warp_nodes && warp_local) {
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < n_warps; ++w) {
      int32_t* wn = warp_nodes + w * 32;
      const int64_t e0 = w * 32, e1 = std::min<int64_t>(n_elems, e0 + 32);
      if (e0 >= n_elems) {
        for (int l = 0; l < 32; ++l) wn[l] = -1;
        continue;
      }
      std::vector<std::pair<int32_t, int32_t>> cnt;  // (node, references)
      std::vector<int32_t> nodes(conn + e0 * npe, conn + e1 * npe);
      std::sort(nodes.begin(), nodes.end());
      for (size_t i = 0; i < nodes.size();) {
        size_t j = i;
        while (j < nodes.size() && nodes[j] == nodes[i]) ++j;
        cnt.emplace_back(nodes[i], (int32_t)(j - i));
        i = j;
      }
      if (cnt.size() > 32) {  // keep the 32 most referenced (ties: smaller node id), then back to ascending ids
        std::stable_sort(cnt.begin(), cnt.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
        cnt.resize(32);
        std::sort(cnt.begin(), cnt.end());
      }
      for (int l = 0; l < 32; ++l) wn[l] = l < (int)cnt.size() ? cnt[l].first : -1;
      for (int64_t i = e0 * npe; i < e1 * npe; ++i) {
        const auto it = std::lower_bound(cnt.begin(), cnt.end(), std::make_pair(conn[i], (int32_t)0),
                                         [](const auto& a, const auto& b) { return a.first < b.first; });
        warp_local[i] = (it != cnt.end() && it->first == conn[i]) ? (uint8_t)(it - cnt.begin()) : (uint8_t)255;
      }
    }
  }
  std::vector<int32_t> chunks(n_tiles, 0);
  std::vector<int64_t> ells(n_tiles, 0);
  // one tile: its (node, slot << 3 | local) references sorted by node, cut into entries of at most `cap` contributors
  // (0 = one entry per distinct node), entries by decreasing contributor count
  auto tile_entries = [&](int64_t t, std::vector<std::pair<int32_t, uint16_t>>& refs, std::vector<std::pair<int32_t, int32_t>>& seg) {
    const int64_t e0 = t * kTileElems, e1 = std::min<int64_t>(n_elems, e0 + kTileElems);
    refs.clear();
    seg.clear();
    for (int64_t e = e0; e < e1; ++e)
      for (int a = 0; a < npe; ++a) refs.emplace_back(conn[e * npe + a], (uint16_t)(((e - e0) << 3) | a));
    std::sort(refs.begin(), refs.end());
    for (size_t i = 0; i < refs.size();) {
      size_t j = i;
      while (j < refs.size() && refs[j].first == refs[i].first) ++j;
      const int32_t count = (int32_t)(j - i);
      const int32_t pieces = cap > 0 ? (count + cap - 1) / cap : 1;
      for (int32_t k = 0; k < pieces; ++k) {  // balanced pieces: sizes differ by at most one
        const int32_t lo = (int32_t)((int64_t)count * k / pieces), hi = (int32_t)((int64_t)count * (k + 1) / pieces);
        seg.emplace_back((int32_t)i + lo, hi - lo);
      }
      i = j;
    }
    std::stable_sort(seg.begin(), seg.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
  };
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, uint16_t>> refs;
    std::vector<std::pair<int32_t, int32_t>> seg;
#pragma omp for schedule(static)
    for (int64_t t = 0; t < n_tiles; ++t) {
      tile_entries(t, refs, seg);
      const int64_t nch = ((int64_t)seg.size() + 31) / 32;
      chunks[t] = (int32_t)nch;
      int64_t tot = 0;
      for (int64_t k = 0; k < nch; ++k) tot += 32 * (int64_t)seg[32 * k].second;
      ells[t] = tot;
    }
  }
  if (tn_node) {
    std::vector<int64_t> ell_first(n_tiles + 1, 0);
    for (int64_t t = 0; t < n_tiles; ++t) ell_first[t + 1] = ell_first[t] + ells[t];
#pragma omp parallel
    {
      std::vector<std::pair<int32_t, uint16_t>> refs;
      std::vector<std::pair<int32_t, int32_t>> seg;
#pragma omp for schedule(static)
      for (int64_t t = 0; t < n_tiles; ++t) {
        tile_entries(t, refs, seg);
        int64_t off = ell_first[t];
        for (int64_t k = 0; k < chunks[t]; ++k) {
          const int64_t ch = ch_ptr[t] + k;
          const int rows = seg[32 * k].second;
          ell_ptr[ch] = (int32_t)off;
          for (int l = 0; l < 32; ++l) {
            const size_t i = (size_t)(32 * k + l);
            tn_node[32 * ch + l] = i < seg.size() ? refs[seg[i].first].first : -1;
            for (int r = 0; r < rows; ++r)
              ell[off + 32 * r + l] = (i < seg.size() && r < seg[i].second) ? refs[seg[i].first + r].second : (uint16_t)0xFFFF;
          }
          off += 32 * (int64_t)rows;
        }
      }
    }
    ell_ptr[ch_ptr[n_tiles]] = (int32_t)ell_first[n_tiles];
    return TATVA_OK;
  }
  ch_ptr[0] = 0;
  int64_t tot = 0;
  for (int64_t t = 0; t < n_tiles; ++t) {
    ch_ptr[t + 1] = ch_ptr[t] + chunks[t];
    tot += ells[t];
  }
  *n_chunks = ch_ptr[n_tiles];
  *n_ell = tot;
  return TATVA_OK;
}

// Uniform background grid over the bounding box of a plane mesh, for point location.  Bin of a coordinate:
// clamp((int)((x - lo) * inv), 0, n - 1) — the same expression the kernel evaluates, and monotone in x, so an element
// is listed in every bin its bounding box can share a point with.  Two-call protocol: bin_elems == NULL fills lo,
// inv and bin_ptr (nx*ny + 1) only.  Elements are appended in ascending order.
static inline int grid_bin_host(double x, double lo, double inv, int n) {
  const double t = (x - lo) * inv;
  return t > 0.0 ? (t < (double)n ? (int)t : n - 1) : 0;
}
int tatva_host_build_point_grid(const double* coords, int64_t n_nodes, const int32_t* conn, int64_t n_elems, int npe,
                                int nx, int ny, double* lo, double* inv, int32_t* bin_ptr, int32_t* bin_elems) {
  if (!coords || !conn || !lo || !inv || !bin_ptr || n_nodes <= 0 || n_elems <= 0 || npe <= 0 || nx <= 0 || ny <= 0)
    return TATVA_E_INVALID;
  if (!bin_elems) {
    double l[2] = {coords[0], coords[1]}, h[2] = {coords[0], coords[1]};
    for (int64_t i = 0; i < n_nodes; ++i)
      for (int d = 0; d < 2; ++d) {
        l[d] = std::min(l[d], coords[2 * i + d]);
        h[d] = std::max(h[d], coords[2 * i + d]);
      }
    for (int d = 0; d < 2; ++d) {
      lo[d] = l[d];
      inv[d] = h[d] > l[d] ? (d == 0 ? nx : ny) / (h[d] - l[d]) : 0.0;
    }
  }
  const int64_t nb = (int64_t)nx * ny;
  std::vector<int32_t> fill;
  if (!bin_elems) std::fill(bin_ptr, bin_ptr + nb + 1, 0);
  else fill.assign(bin_ptr, bin_ptr + nb);
  for (int64_t e = 0; e < n_elems; ++e) {
    double l[2] = {1e308, 1e308}, h[2] = {-1e308, -1e308};
    for (int a = 0; a < npe; ++a) {
      const int32_t n = conn[e * npe + a];
      if (n < 0 || n >= n_nodes) return TATVA_E_INVALID;
      for (int d = 0; d < 2; ++d) {
        l[d] = std::min(l[d], coords[2 * (int64_t)n + d]);
        h[d] = std::max(h[d], coords[2 * (int64_t)n + d]);
      }
    }
    const int x0 = grid_bin_host(l[0], lo[0], inv[0], nx), x1 = grid_bin_host(h[0], lo[0], inv[0], nx);
    const int y0 = grid_bin_host(l[1], lo[1], inv[1], ny), y1 = grid_bin_host(h[1], lo[1], inv[1], ny);
    for (int iy = y0; iy <= y1; ++iy)
      for (int ix = x0; ix <= x1; ++ix) {
        const int64_t b = (int64_t)iy * nx + ix;
        if (!bin_elems) {
          if (bin_ptr[b + 1] == INT32_MAX) return TATVA_E_INVALID;
          bin_ptr[b + 1]++;
        } else {
          bin_elems[fill[b]++] = (int32_t)e;
        }
      }
  }
  if (!bin_elems) {
    int64_t total = 0;
    for (int64_t b = 0; b < nb; ++b) {
      total += bin_ptr[b + 1];
      if (total > INT32_MAX) return TATVA_E_INVALID;
      bin_ptr[b + 1] = (int32_t)total;
    }
  }
  return TATVA_OK;
}

// elem_pos[e, a, b] = offset of column dpn*conn[e,b] inside CSR row dpn*conn[e,a].
// The assembly kernels write entry (i, k) of the (a, b) block at indptr[dpn*node_a + i] + pos + k, so the pattern must
// be NODE-BLOCKED: in EVERY component row dpn*node_a + i the dpn columns dpn*node_b .. dpn*node_b + dpn-1 must sit
// contiguously at the same offset `pos`.  Patterns that are not (a per-component Periodic in lifter.augment_sparsity, a
// user-supplied pattern, a non-interleaved compound layout) are rejected with TATVA_E_INVALID; sparse.jacfwd then
// takes the coloured-JVP route of the reference (sparse/base.py:230-270) instead of the direct kernel.
int tatva_host_csr_element_positions(const int32_t* conn, int64_t n_elems, int npe, int dpn, const int32_t* indptr,
                                     const int32_t* indices, int32_t* elem_pos) {
  if (!conn || !indptr || !indices || !elem_pos || n_elems <= 0 || npe <= 0 || dpn <= 0) return TATVA_E_INVALID;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t e = 0; e < n_elems; ++e) {
    for (int a = 0; a < npe; ++a) {
      const int64_t row = (int64_t)conn[e * npe + a] * dpn;
      const int32_t* lo = indices + indptr[row];
      const int32_t* hi = indices + indptr[row + 1];
      for (int b = 0; b < npe; ++b) {
        const int32_t col = conn[e * npe + b] * dpn;
        const int32_t* it = std::lower_bound(lo, hi, col);
        int32_t pos = -1;
        if (it != hi && *it == col) {
          pos = (int32_t)(it - lo);
          for (int i = 0; i < dpn && pos >= 0; ++i) {  // every component row: same offset, dpn contiguous columns
            const int64_t r0 = indptr[row + i], r1 = indptr[row + i + 1];
            if (r0 + pos + dpn > r1) { pos = -1; break; }
            for (int k = 0; k < dpn; ++k)
              if (indices[r0 + pos + k] != col + k) { pos = -1; break; }
          }
        }
        if (pos < 0) bad |= 1;
        elem_pos[(e * npe + a) * npe + b] = pos;
      }
    }
  }
  return bad ? TATVA_E_INVALID : TATVA_OK;
}

// Combine schedule of the tiled CSR assembly (k_csr_tiled): the elements are cut into tiles of `tile` consecutive
// elements; within a tile every (row node, column node) block that several elements contribute to is listed ONCE, with
// its contributors, so that the kernel sums them on chip and issues one RED group per distinct block instead of one
// per element (config 2: 16 blocks per tet, 2.9-3.5 x fewer distinct blocks per 128-element tile).  The energy Hessian
// is symmetric, K_ba = K_ab^T, so only the blocks with row node <= column node are listed; each carries the position
// of its mirror block as well and the kernel adds the transposed sums there (half the on-chip work).
//   blk_ptr  [n_tiles + 1]  first block of each tile
//   blk_base [n_blk]        indptr[dpn * a] + elem_pos(a, b): position of the block in the first row of node a
//   blk_rowlen [n_blk]      length of the rows of node a (row i of the block sits i * rowlen further)
//   blk_base_t, blk_rowlen_t [n_blk]   the same for the mirror block (b, a); base -1 for a diagonal block
//   con_ptr  [n_blk + 1]    first contributor of each block
//   con      [n_con]        (element index inside the tile) << 8 | local row node << 4 | local column node
// Blocks of a tile are listed by DECREASING contributor count (see below).  Two calls: with blk_base == NULL only
// *n_blk and *n_con are returned (and blk_ptr filled).
int tatva_host_csr_tile_schedule(const int32_t* conn, int64_t n_elems, int npe, int dpn, int tile, const int32_t* indptr,
                                 const int32_t* elem_pos, int32_t* blk_ptr, int64_t* n_blk, int64_t* n_con, int32_t* blk_base,
                                 int32_t* blk_rowlen, int32_t* blk_base_t, int32_t* blk_rowlen_t, int32_t* con_ptr,
                                 uint32_t* con) {
  if (!conn || !indptr || !elem_pos || !blk_ptr || !n_blk || !n_con || n_elems <= 0 || npe <= 0 || npe > 16 || dpn <= 0 || tile <= 0 || tile > (1 << 20)) return TATVA_E_INVALID;
  const int64_t n_tiles = (n_elems + tile - 1) / tile;
  const bool fill = blk_base != nullptr;
  if (fill && (!blk_rowlen || !blk_base_t || !blk_rowlen_t || !con_ptr || !con)) return TATVA_E_INVALID;
  std::vector<int32_t> counts(n_tiles, 0);
  std::vector<int64_t> con_first;  // first contributor slot of each tile (fill pass)
  if (fill) {
    con_first.assign(n_tiles + 1, 0);
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int64_t e0 = t * tile, e1 = std::min<int64_t>(n_elems, e0 + tile);
      int64_t c = 0;
      for (int64_t e = e0; e < e1; ++e)
        for (int a = 0; a < npe; ++a)
          for (int b = 0; b < npe; ++b) c += conn[e * npe + a] <= conn[e * npe + b];
      con_first[t + 1] = con_first[t] + c;
    }
  }
#pragma omp parallel
  {
    std::vector<std::pair<int64_t, uint32_t>> keys;
    std::vector<std::pair<int32_t, int32_t>> order;
#pragma omp for schedule(dynamic, 16)
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int64_t e0 = t * tile, e1 = std::min<int64_t>(n_elems, e0 + tile);
      keys.clear();
      for (int64_t e = e0; e < e1; ++e)
        for (int a = 0; a < npe; ++a)
          for (int b = 0; b < npe; ++b) {
            const int32_t na = conn[e * npe + a], nb = conn[e * npe + b];
            if (na <= nb) keys.emplace_back(((int64_t)na << 32) | (uint32_t)nb, (uint32_t)((e - e0) << 8 | a << 4 | b));
          }
      std::sort(keys.begin(), keys.end());
      // blocks of the tile in order of DECREASING contributor count: the lanes of a warp then loop over similar
      // numbers of contributors (a diagonal block of the tet box has up to 24, most off-diagonal ones 2-6; in key
      // order every warp would run its longest lane's loop: ~3 x the instructions, measured)
      order.clear();
      for (size_t k = 0; k < keys.size();) {
        size_t k1 = k + 1;
        while (k1 < keys.size() && keys[k1].first == keys[k].first) ++k1;
        order.emplace_back(-(int32_t)(k1 - k), (int32_t)k);
        k = k1;
      }
      if (!fill) {
        counts[t] = (int32_t)order.size();
        continue;
      }
      std::stable_sort(order.begin(), order.end(), [](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) { return x.first < y.first; });
      int64_t bi = blk_ptr[t], ci = con_first[t];
      for (const auto& o : order) {
        const size_t k0 = (size_t)o.second, cnt = (size_t)(-o.first);
        const uint32_t src = keys[k0].second;
        const int64_t e = e0 + (src >> 8);
        const int a = (src >> 4) & 15, b = src & 15;
        const int64_t row = (int64_t)conn[e * npe + a] * dpn, rowt = (int64_t)conn[e * npe + b] * dpn;
        blk_base[bi] = indptr[row] + elem_pos[(e * npe + a) * npe + b];
        blk_rowlen[bi] = indptr[row + 1] - indptr[row];
        const bool diag = conn[e * npe + a] == conn[e * npe + b];
        blk_base_t[bi] = diag ? -1 : indptr[rowt] + elem_pos[(e * npe + b) * npe + a];
        blk_rowlen_t[bi] = indptr[rowt + 1] - indptr[rowt];
        con_ptr[bi] = (int32_t)ci;
        ++bi;
        for (size_t k = k0; k < k0 + cnt; ++k) con[ci++] = keys[k].second;
      }
    }
  }
  if (!fill) {
    int64_t total = 0, ncon = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      blk_ptr[t] = (int32_t)total;
      total += counts[t];
      if (total > INT32_MAX) return TATVA_E_INVALID;
    }
    blk_ptr[n_tiles] = (int32_t)total;
    for (int64_t e = 0; e < n_elems; ++e)
      for (int a = 0; a < npe; ++a)
        for (int b = 0; b < npe; ++b) ncon += conn[e * npe + a] <= conn[e * npe + b];
    if (ncon > INT32_MAX) return TATVA_E_INVALID;
    *n_blk = total;
    *n_con = ncon;
  } else {
    con_ptr[blk_ptr[n_tiles]] = (int32_t)con_first[n_tiles];
  }
  return TATVA_OK;
}

}  // extern "C"
