/*
 * tatva_b200 — C ABI of the B200-native element-level hot path of tatva.
 *
 * This header is the drop-in boundary: every entry point is `extern "C"`, takes plain
 * pointers and sizes, enqueues work on the caller's CUDA stream and returns an int
 * (0 = ok, >0 = cudaError_t, <0 = TATVA_E_*).  Hot calls do not allocate, do not
 * synchronise and do not throw.  Device buffers are caller-owned, FP64 row-major,
 * connectivity int32 (reference: tatva/mesh.py:205), CSR indptr/indices int32.
 * Output buffers need not be zero-initialised (XLA does not zero FFI results): calls that
 * scatter-add zero their output on-stream first.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the tatva
 * v0.11.1 tree).  INTEGRATION.md shows the XLA-FFI / ctypes stubs that bind these symbols.
 */
#ifndef TATVA_B200_H
#define TATVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TATVA_B200_ABI_VERSION 9  /* bumped on every signature change; the Python loader refuses a mismatch */

typedef struct tatva_plan tatva_plan_t; /* opaque: mesh views + scratch for one Operator */
typedef void* tatva_stream_t;           /* a cudaStream_t */

/* element kinds — tatva/element/base.py:245-265 (Tri3), :448-472 (Tetrahedron4), :475-568 (Hexahedron8),
 * :331-366 (Quad4), :266-328 (Tri6), :366-445 (Quad8), each with its default quadrature rule */
enum { TATVA_TRI3 = 0, TATVA_TET4 = 1, TATVA_HEX8 = 2, TATVA_QUAD4 = 3, TATVA_TRI6 = 4, TATVA_QUAD8 = 5,
       TATVA_LINE2 = 6, TATVA_LINE3 = 7 };

/* energy densities the configs name (user code in the reference, pinned by its tests):
 *   LINEAR_ELASTIC          psi = 1/2 sigma:eps           params {mu, lambda}          tests/test_sparse.py:20-38
 *   NEO_HOOKEAN             psi = mu/2 (I1-3-2lnJ) + lambda/2 (lnJ)^2   params {mu, lambda}   tests/test_sparse_tracer.py:103-115
 *   NEO_HOOKEAN_PHASE_FIELD ((1-phi)^2+k) psi_NH + Gc (phi^2/(2l) + l/2 |grad phi|^2),   params {mu, lambda, Gc, l, k}
 *                           nodal state interleaved [ux,uy,uz,phi] (tatva/compound/__init__.py:334-389)   */
enum { TATVA_LINEAR_ELASTIC = 0, TATVA_NEO_HOOKEAN = 1, TATVA_NEO_HOOKEAN_PHASE_FIELD = 2 };
#define TATVA_USER_LAW_BASE 1000 /* material ids handed out by tatva_law_register */

/* error codes (negative); positive return values are cudaError_t */
enum {
  TATVA_OK = 0,
  TATVA_E_INVALID = -1,     /* bad argument (null pointer, unknown kind, size <= 0) */
  TATVA_E_UNSUPPORTED = -2, /* element/material combination has no kernel */
  TATVA_E_NOMEM = -3,
  TATVA_E_NODEVICE = -4
};

/* plan flags */
enum {
  TATVA_PLAN_CACHE_WEIGHTS = 1 /* Operator(cache_weights=True), tatva/operator.py:119-130 */
};

/* kernel variant selector for tatva_hvp on (HEX8, NEO_HOOKEAN): 0 = tuned default */
enum { TATVA_VARIANT_DEFAULT = 0, TATVA_VARIANT_GENERIC = 1, TATVA_VARIANT_MODAL = 2 };

const char* tatva_error_string(int code);
int tatva_abi_version(void);
int tatva_device_count(int* count);

/* ---- plan lifetime ---------------------------------------------------------------------
 * Replaces Operator.__post_init__ (tatva/operator.py:112-130).  `d_coords` (n_nodes, dim)
 * f64 and `d_conn` (n_elems, npe) int32 are device pointers that must outlive the plan
 * (the plan keeps views, not copies).  Allocates the plan's scratch (energy partials,
 * cached weights when TATVA_PLAN_CACHE_WEIGHTS).                                           */
int tatva_plan_create(tatva_plan_t** plan, int element, int64_t n_nodes, int64_t n_elems,
                      const double* d_coords, const int32_t* d_conn, int flags,
                      tatva_stream_t stream);
int tatva_plan_destroy(tatva_plan_t* plan);
int tatva_plan_info(const tatva_plan_t* plan, int* element, int* dim, int* npe, int* nq,
                    int64_t* n_nodes, int64_t* n_elems);
/* Measurement switch, not part of the drop-in surface: 0 (default) = the tuned kernels; 1 = the generic element /
 * law templates everywhere (TATVA_VARIANT_GENERIC, the parity cross-check of every specialised kernel); other
 * values select alternative implementations kept for A/B timing, per kernel family (DESIGN.md sections 3.1-3.3):
 *   Hex8 x neo-Hookean HVP   2, 3, 8, 9, 15, 16, 17, 20, 22, 23, 25, 26, 27 (sector-grouped scatter), 28 (16-byte gathers),
 *                            50-56 = occupancy / staging points of the cached-geometry kernel (with a geometry cache)
 *   Hex8 residual / energy   2 = first modal kernel, 3 / 4 = pair kernel at other register / occupancy points
 *   Tet4 x neo-Hookean       30 = persistent kernel with connectivity prefetch, 31 = element-per-thread kernel even
 *                            when the plan carries a node schedule (tatva_plan_set_node_schedule)
 *   building blocks          2 = element-per-thread staged kernels (same as 0), 3 = one thread per quadrature point
 *   CSR assembly             2 = full assembly also where the symmetric entry point was called
 * Every variant computes the same result to rounding; unknown values fall back to the default.                  */
/* Re-point a plan at other device buffers of the SAME sizes (no allocation, no synchronisation): what a caller whose
 * runtime re-allocates buffers between executions needs (the XLA-FFI shim caches plans by sizes and rebinds per call).
 * TATVA_E_UNSUPPORTED for a plan with cached integration weights and different coordinates.                           */
int tatva_plan_rebind(tatva_plan_t* plan, const double* d_coords, const int32_t* d_conn);

int tatva_plan_set_variant(tatva_plan_t* plan, int variant);

/* Geometry cache for the Hex8 x neo-Hookean HVP (r02): everything the Gauss-point loop derives from the MESH alone —
 * (1 / det J) adj(J)^T adj(J), det J and 1 / det J at the 8 Gauss points, J = dX/dxi (tatva/element/base.py:90-93,
 * :99-115) — computed once and kept by the plan, 512 bytes per element, as Operator(cache_weights=True) keeps det J
 * (tatva/operator.py:119-130).  The HVP then reads it back (one coalesced 512-byte row per warp and load) instead of
 * re-deriving it: 64 of 305 FP64 instructions per Gauss point less, paid with HBM bandwidth the FP64-bound kernel leaves
 * idle.  enable = 0 frees the cache.  Hex8 with the default rule only (TATVA_E_UNSUPPORTED otherwise).  Allocates.   */
int tatva_plan_cache_geometry(tatva_plan_t* plan, int enable, tatva_stream_t stream);

/* User-supplied quadrature rule — Element(quad_points, quad_weights), tatva/element/base.py:37-51.  `points` is
 * (nq, reference dimension) row-major HOST memory, `weights` (nq) HOST memory, nq <= 64; nq = 0 restores the element's
 * default rule.  Every Operator building block (weights, grad, eval, integrate and their adjoints) and the fused
 * energy / residual / HVP / lifted HVP / Hessian diagonal / CSR assembly then run the GENERIC kernels with the rule in
 * constant memory (the modal Hex8 and reference-space Tet4 kernels are default-rule only).  One rule slot per process
 * and device: plans with different custom rules must not launch concurrently on different streams.
 * tatva_op_interpolate returns TATVA_E_UNSUPPORTED with a custom rule.                                              */
int tatva_plan_set_quadrature(tatva_plan_t* plan, int nq, const double* points, const double* weights,
                              tatva_stream_t stream);

/* ---- user-supplied energy densities (README.md:93: in the reference the density is user code differentiated by JAX) ----
 * `cuda_source` is the CUDA text of  `struct UserLaw { static constexpr int dim, dpn, val_lo, n_params; double prm[..];
 * struct Cache; using S = tatva::QState<dpn, dim>; prepare / psi / first / second }`  — the `Mat` interface of
 * csrc/common.cuh; tatva_b200/lawgen.py writes it from a density given once on symbols (forward-over-reverse AD of its
 * straight-line program).  The returned material id (>= TATVA_USER_LAW_BASE) is accepted by tatva_energy / _residual /
 * _hvp / _hvp_elems / _residual_elems / _hvp_lifted / _hessian_diag / _hvp_dot: at first use with a plan the source is
 * compiled by NVRTC into the same fused kernel templates as the built-in laws (per element kind and device, cached)
 * and launched on the caller's stream.  `params` of those calls fill prm[0..n_params).  tatva_csr_assemble returns
 * TATVA_E_UNSUPPORTED for a user law (sparse.jacfwd then runs one fused HVP per colour).  A failed compilation
 * returns TATVA_E_INVALID; its log is in tatva_law_compile_log.                                                    */
int tatva_law_register(const char* cuda_source, int dim, int dofs_per_node, int n_params, int uses_values,
                       int* material_id);
int tatva_law_compile_log(char* buf, int len);
/* Optional shared-memory staging tiles for gather-bound elements (Tet4 x neo-Hookean residual / HVP): tile t =
 * elements [128 t, 128 (t+1)); d_tile_nodes[d_tile_ptr[t] .. d_tile_ptr[t+1]) are its sorted unique nodes and
 * d_tile_conn (n_elems, npe) uint16 its connectivity in tile-local indices (tatva_host_build_tiles).  The CTA
 * gathers the unique nodes once, coalesced, into shared memory.  Device views, caller-owned; NULL disables.   */
int tatva_plan_set_tiles(tatva_plan_t* plan, const int32_t* d_tile_ptr, const int32_t* d_tile_nodes,
                         const uint16_t* d_tile_conn, int max_unique);

/* Node schedule of the warp-cooperative fused kernels (r02; Tri3 / Tet4 residual and HVP — the transpose of the gather
 * `v[self.mesh.elements]`, tatva/operator.py:221, done node-wise): every warp loads the nodal rows of its (up to) 32
 * distinct nodes ONCE, one node per lane, and the elements fetch them with warp shuffles; the nodal contributions of a
 * 128-element tile are summed per distinct node in shared memory and leave the SM as ONE atomic add per (node, DOF).
 * Arrays from tatva_host_node_schedule on the plan's element list; d_tile_hdr holds four int32 per tile, 16-byte aligned:
 * {ch_ptr[t], ch_ptr[t+1] - ch_ptr[t], ell_ptr[ch_ptr[t]], ell_ptr[ch_ptr[t+1]] - ell_ptr[ch_ptr[t]]}.  Device views,
 * caller-owned; NULL disables.                                                                                        */
int tatva_plan_set_node_schedule(tatva_plan_t* plan, const int32_t* d_warp_nodes, const uint8_t* d_warp_local,
                                 const int32_t* d_tile_hdr, const int32_t* d_tn_node, const int32_t* d_ell_ptr,
                                 const uint16_t* d_ell);

/* Optional uniform background grid for tatva_op_interpolate (plane meshes): bin (ix, iy), row-major iy * nx + ix, with
 * ix = clamp((int)((x - lo[0]) * inv[0]), 0, nx - 1); d_bin_elems[d_bin_ptr[b] .. d_bin_ptr[b+1]) lists, ascending,
 * the elements whose bounding box overlaps bin b (tatva_host_build_point_grid), so the first containing element of the
 * point's bin is the one mesh.find_containing_polygons returns (tatva/mesh.py:294-388).  Device views, caller-owned;
 * d_bin_ptr == NULL disables the grid (every element is scanned).                                              */
int tatva_plan_set_point_grid(tatva_plan_t* plan, int nx, int ny, const double* lo, const double* inv,
                              const int32_t* d_bin_ptr, const int32_t* d_bin_elems);

/* ---- quadrature-loop building blocks (generic path; any user energy on top) ------------ */

/* Operator.grad -> Element.gradient  (tatva/operator.py:379-397, tatva/element/base.py:99-115)
 * u (n_nodes, n_val) -> out (n_elems, nq, n_val, dim), out[e,q,i,j] = d u_i / d x_j.       */
int tatva_op_grad(const tatva_plan_t* plan, const double* d_u, int n_val, double* d_out,
                  tatva_stream_t stream);
/* adjoint of the above = what jax.grad makes of the gather at operator.py:221:
 * g (n_elems, nq, n_val, dim) -> y (n_nodes, n_val), y[n,i] = sum_{e,q,j} g[e,q,i,j] dNdX[e,q,j,n] */
int tatva_op_grad_adjoint(const tatva_plan_t* plan, const double* d_g, int n_val, double* d_y,
                          tatva_stream_t stream);
/* Operator.eval -> Element.interpolate (tatva/operator.py:358-377, element/base.py:95-97)
 * u (n_nodes, n_val) -> out (n_elems, nq, n_val)                                           */
int tatva_op_eval(const tatva_plan_t* plan, const double* d_u, int n_val, double* d_out,
                  tatva_stream_t stream);
int tatva_op_eval_adjoint(const tatva_plan_t* plan, const double* d_g, int n_val, double* d_y,
                          tatva_stream_t stream);
/* Operator.get_integration_weights (tatva/operator.py:172-192): out (n_elems, nq) = det(J) w_q (no abs) */
int tatva_op_integration_weights(const tatva_plan_t* plan, double* d_out, tatva_stream_t stream);
/* Operator._integrate_quad_array (tatva/operator.py:342-356):
 * vals (n_elems, nq, n_val) -> out (n_elems, n_val) = einsum("eq...,eq->e...", vals, W)    */
int tatva_op_integrate_quad(const tatva_plan_t* plan, const double* d_vals, int n_val,
                            double* d_out, tatva_stream_t stream);
/* Operator.interpolate (tatva/operator.py:399-463) with mesh.find_containing_polygons (tatva/mesh.py:294-388):
 * u (n_nodes, n_val) at n_points physical points (n_points, 2) -> out (n_points, n_val).  Plane elements only,
 * as in the reference.  d_elem (n_points, int32): containing element, -1 (and NaN values) outside the mesh.   */
int tatva_op_interpolate(const tatva_plan_t* plan, const double* d_u, int n_val, const double* d_points,
                         int64_t n_points, double* d_out, int32_t* d_elem, tatva_stream_t stream);
/* the gather `v[self.mesh.elements]` of Operator.map / map_over_elements
 * (tatva/operator.py:254-257, :296-299): u (n_nodes, n_val) -> out (n_elems, npe, n_val)    */
int tatva_op_gather(const tatva_plan_t* plan, const double* d_u, int n_val, double* d_out,
                    tatva_stream_t stream);
/* its transpose (scatter-add): g (n_elems, npe, n_val) -> y (n_nodes, n_val)                */
int tatva_op_gather_adjoint(const tatva_plan_t* plan, const double* d_g, int n_val, double* d_y,
                            tatva_stream_t stream);
/* sum over axis 0 of (n_rows, n_val) -> (n_val): the `jnp.sum(res, axis=0)` of Operator.integrate
 * (tatva/operator.py:319); deterministic two-pass tree, uses plan scratch                   */
int tatva_op_sum_rows(tatva_plan_t* plan, const double* d_in, int64_t n_rows, int n_val,
                      double* d_out, tatva_stream_t stream);

/* ---- fused energy / residual / HVP ------------------------------------------------------
 * E(u) = op.integrate(psi(op.grad(u)))           (README.md:93; tests/test_sparse.py:50-55)
 * r    = jax.grad(E)(u)                          (tests/test_sparse.py:79)
 * Hv   = jax.jvp(jax.grad(E), (u,), (v,))[1]     (tatva/sparse/base.py:264)
 * u, v, y: (n_nodes, dofs_per_node) with dofs_per_node = dim (4 for the phase-field law).
 * `params` is a HOST pointer to n_params doubles (copied into the launch).                 */
int tatva_energy(tatva_plan_t* plan, int material, const double* params, int n_params,
                 const double* d_u, double* d_energy, tatva_stream_t stream);
int tatva_residual(tatva_plan_t* plan, int material, const double* params, int n_params,
                   const double* d_u, double* d_r, tatva_stream_t stream);
int tatva_hvp(tatva_plan_t* plan, int material, const double* params, int n_params,
              const double* d_u, const double* d_v, double* d_y, tatva_stream_t stream);

/* Diagonal of the energy Hessian at u, diag[dpn*n + k] = d2E/du_{n,k}^2 (n_nodes * dofs_per_node): the Jacobi
 * preconditioner of the CG around the HVP (the reference ships no solver and
 * would obtain it as ColoredMatrix.diagonal() after sparse.jacfwd, tatva/sparse/base.py:37-105).           */
int tatva_hessian_diag(tatva_plan_t* plan, int material, const double* params, int n_params,
                       const double* d_u, double* d_diag, tatva_stream_t stream);

/* HVP with the Lifter folded into the kernel's gather / scatter (SURVEY.md section 8(f) row 1;
 * tatva/lifter/base.py:201-251, constraints.py:214-221, :312-318):
 *     y_red = reduce_adjoint( H(u_full) . lift_0(v_red) )
 * `d_u_full` is the lifted state (n_nodes * dpn), `d_v_red` / `d_y_red` live on the n_red free DOFs, and
 * `d_dof_map[i]` (n_nodes * dpn, int32) is the reduced DOF that drives full DOF i (its own for a free DOF, the
 * master's for a Periodic image), or -1 for a Fixed DOF (tangent value 0, contribution dropped).          */
int tatva_hvp_lifted(tatva_plan_t* plan, int material, const double* params, int n_params,
                     const double* d_u_full, const double* d_v_red, const int32_t* d_dof_map,
                     int64_t n_red, double* d_y_red, tatva_stream_t stream);
/* tatva_hvp_lifted that also leaves v_red . y_red in d_scalars[slot] (CG: p.Ap for free — the Hex8 x neo-Hookean kernel
 * sums it element by element before the scatter; other pairs add a two-pass dot).  roll != 0: d_scalars[0] <- d_scalars[2]
 * in the same final-sum kernel.  zero_y = 0: y_red is already zero (see tatva_cg_after_dot).  d_partials >= 1184 doubles. */
int tatva_hvp_lifted_dot(tatva_plan_t* plan, int material, const double* params, int n_params,
                         const double* d_u_full, const double* d_v_red, const int32_t* d_dof_map, int64_t n_red,
                         double* d_y_red, int zero_y, double* d_partials, double* d_scalars, int slot, int roll,
                         tatva_stream_t stream);

/* y[0..n) = 0 by a kernel that releases its dependent grid at once; with zero_y = 2 on the sub-range launch issued right
 * behind it on the same stream, that launch starts while y is still being cleared and waits only before its first add.  */
int tatva_zero_release(double* d_y, int64_t n, tatva_stream_t stream);
/* Element sub-range variants: only elements [elem_begin, elem_begin + elem_count) contribute, and the
 * output is zeroed first only if zero_out != 0.  They let the caller run the elements that touch ghost
 * nodes and the interior elements on different streams, so the halo exchange of tatva/mpi.py:372-409,
 * :479-516 overlaps the interior quadrature loop.  tatva_hvp_elems: zero_out = 2 says the output is being
 * cleared by the tatva_zero_release issued just before on the same stream (nothing is zeroed here; the Hex8 x
 * neo-Hookean kernel is launched behind it with programmatic stream serialization, other kernels in stream order). */
int tatva_hvp_elems(tatva_plan_t* plan, int material, const double* params, int n_params,
                    const double* d_u, const double* d_v, double* d_y, int64_t elem_begin,
                    int64_t elem_count, int zero_out, tatva_stream_t stream);
int tatva_residual_elems(tatva_plan_t* plan, int material, const double* params, int n_params,
                         const double* d_u, double* d_r, int64_t elem_begin, int64_t elem_count,
                         int zero_out, tatva_stream_t stream);

/* ---- coloured sparse Jacobian -> direct assembly into a fixed CSR pattern ---------------
 * Replaces sparse.jacfwd / colored_jacobian_batch / compute_rows_cols
 * (tatva/sparse/base.py:139-176, :230-270, :108-136): instead of n_colors HVPs and an
 * (N, n_colors) temporary, one kernel adds every element stiffness into `d_data` (nnz).
 * `d_elem_pos` (n_elems, npe, npe) int32: for element e and node pair (a,b), the offset of
 * column dpn*conn[e,b] inside CSR row dpn*conn[e,a] (same offset in the dpn rows of node a);
 * built once on the host by tatva_host_csr_element_positions.                              */
int tatva_csr_assemble(tatva_plan_t* plan, int material, const double* params, int n_params,
                       const double* d_u, const int32_t* d_indptr, const int32_t* d_elem_pos,
                       int64_t nnz, double* d_data, tatva_stream_t stream);

/* Tiled assembly (r02): a CTA owns a tile of 128 consecutive elements, evaluates each element's geometry and the
 * law's rank structure once into shared memory, then every DISTINCT (row node, column node) block of the tile is summed
 * from its contributors on chip and added to `d_data` with one RED group — ~3 x fewer REDs and ~4 x fewer instructions
 * than tatva_csr_assemble at config 2; only blocks with row node <= column node are summed, their transposes go to the
 * mirror blocks (the energy Hessian is symmetric).  Tri3 / Tet4 x {neo-Hookean, linear elastic}; TATVA_E_UNSUPPORTED otherwise.
 * The schedule comes from tatva_host_csr_tile_schedule (host, once per pattern); `d_conn` is the element list it was
 * built for (n_elems x npe, normally locality-sorted).  Replaces sparse/base.py:139-176, :230-270 like
 * tatva_csr_assemble.                                                                                                */
int tatva_host_csr_tile_schedule(const int32_t* conn, int64_t n_elems, int npe, int dofs_per_node, int tile,
                                 const int32_t* indptr, const int32_t* elem_pos, int32_t* blk_ptr, int64_t* n_blk,
                                 int64_t* n_con, int32_t* blk_base, int32_t* blk_rowlen, int32_t* blk_base_t,
                                 int32_t* blk_rowlen_t, int32_t* con_ptr, uint32_t* con);
int tatva_csr_assemble_tiled(tatva_plan_t* plan, int material, const double* params, int n_params, const double* d_u,
                             const int32_t* d_conn, const int32_t* d_blk_ptr, const int32_t* d_blk_base,
                             const int32_t* d_blk_rowlen, const int32_t* d_blk_base_t, const int32_t* d_blk_rowlen_t,
                             const int32_t* d_con_ptr, const uint32_t* d_con, int64_t nnz, double* d_data,
                             tatva_stream_t stream);

/* Same matrix, exploiting the symmetry of the energy Hessian: REDs only for the upper triangle (row node <=
 * column node), then a mirror pass fills the lower one (K[(a,i),(b,k)] = K[(b,k),(a,i)]).  Needs the CSR column
 * indices.  ~46 % fewer atomics than tatva_csr_assemble; for multi-point elements it falls back to full assembly. */
int tatva_csr_assemble_sym(tatva_plan_t* plan, int material, const double* params, int n_params,
                           const double* d_u, const int32_t* d_indptr, const int32_t* d_indices,
                           const int32_t* d_elem_pos, int64_t nnz, double* d_data, tatva_stream_t stream);

/* Same matrix, assembled BY ROWS without atomics (deterministic): one warp owns the CSR rows of one node,
 * computes the columns (a,i) of the stiffness of every incident element (= the rows, by symmetry of the
 * energy Hessian) and stores each dpn x dpn block once.  Needs the node -> elements table of
 * tatva_host_node_to_elements.  Single-quadrature-point elements (Tri3, Tet4); returns
 * TATVA_E_UNSUPPORTED otherwise (use tatva_csr_assemble).  `d_data` needs no zeroing.              */
int tatva_csr_assemble_rows(tatva_plan_t* plan, int material, const double* params, int n_params,
                            const double* d_u, const int32_t* d_indptr, const int32_t* d_indices,
                            const int32_t* d_n2e_ptr, const int32_t* d_n2e, double* d_data,
                            tatva_stream_t stream);

/* ---- halo exchange building blocks (tatva/mpi.py:372-409, :479-516) ---------------------
 * pack:        dst[k]        = src[idx[k]]      (send_buf = x_owned[nbr_send], mpi.py:400)
 * unpack_set:  dst[idx[k]]   = src[k]           (u_local.at[nbr_recv].set,     mpi.py:406-407)
 * unpack_add:  dst[idx[k]]  += src[k]           (owned.at[nbr_recv].add,       mpi.py:512-513;
 *                                                indices may repeat -> atomic)              */
int tatva_halo_pack(const double* d_src, const int64_t* d_idx, int64_t n, double* d_dst,
                    tatva_stream_t stream);
int tatva_halo_unpack_set(const double* d_src, const int64_t* d_idx, int64_t n, double* d_dst,
                          tatva_stream_t stream);
int tatva_halo_unpack_add(const double* d_src, const int64_t* d_idx, int64_t n, double* d_dst,
                          tatva_stream_t stream);

/* The exchange itself over an ncclComm_t (tatva/mpi.py:372-409 forward fill, :479-516 reverse add — there
 * one blocking mpi4jax.sendrecv per neighbour, :403-405 / :509-511): pack kernel -> ONE grouped ncclSend / ncclRecv with
 * every neighbour -> unpack kernel, all on `stream`, no allocation, no synchronisation.  `nccl_comm` is an ncclComm_t;
 * send_counts / recv_counts are HOST arrays with one entry per rank of the communicator (own rank: 0), the index arrays
 * list the neighbours' entries in rank order, d_send_buf / d_recv_buf hold sum(send_counts) / sum(recv_counts) doubles.
 * add = 0: d_dst[d_recv_idx[k]] = received (ghost refresh); add = 1: += (reverse add).  d_src may equal d_dst.
 * NCCL is opened with dlopen at first use (TATVA_E_UNSUPPORTED if the machine has none).                           */
int tatva_halo_exchange(void* nccl_comm, const double* d_src, const int64_t* d_send_idx, const int64_t* send_counts,
                        double* d_send_buf, double* d_recv_buf, const int64_t* recv_counts, const int64_t* d_recv_idx,
                        double* d_dst, int add, tatva_stream_t stream);
/* For a host without a communicator of its own: rank 0 draws the 128-byte id (tatva_halo_comm_unique_id), the host
 * ships it to the other ranks, every rank calls tatva_halo_comm_create (collective) with its device current.       */
int tatva_nccl_version(int* version);
int tatva_halo_comm_unique_id(void* id128);
int tatva_halo_comm_create(void** nccl_comm, const void* id128, int n_ranks, int rank);
int tatva_halo_comm_destroy(void* nccl_comm);

/* Peer-memory variant of the exchange (one box, NVLink / NVSwitch): local vectors live in peer-mapped memory,
 * `d_peer_ptrs[r]` is rank r's base address of the same vector, and each ghost entry g knows its owner rank and
 * its index in the owner's vector.  pull fills the ghosts with remote loads, push_add adds the ghost contributions
 * into the owners' rows with remote REDs — no pack buffers, no NCCL call.  The caller brackets them with a
 * cross-rank barrier (owners' data ready / all contributions landed).                                          */
int tatva_peer_pull(double* d_x_local, int64_t first_ghost, int64_t n_ghost, const uint64_t* d_peer_ptrs,
                    const int32_t* d_owner, const int64_t* d_owner_idx, tatva_stream_t stream);
int tatva_peer_push_add(const double* d_y_local, int64_t first_ghost, int64_t n_ghost,
                        const uint64_t* d_peer_ptrs, const int32_t* d_owner, const int64_t* d_owner_idx,
                        tatva_stream_t stream);

/* ---- lifter (tatva/lifter/base.py:201-251): reduced <-> full vectors with all constraints composed ----
 * The constraints (Fixed, Periodic, applied in order; lifter/constraints.py:214-221, :312-318) are composed
 * once on the host into one source table, so lift is ONE gather and reduce_adjoint ONE segmented sum:
 *   lift:            out[i] = src[i] >= 0 ? u_red[src[i]] : (src[i] == -1 ? base[i] (0 if base is NULL)
 *                                                                          : consts[-(src[i] + 2)])
 *   reduce_adjoint:  r_red[j] = sum_{k in [ptr[j], ptr[j+1])} r_full[list[k]]   (fixed order: deterministic) */
int tatva_lift(const double* d_u_red, const int64_t* d_src, const double* d_consts,
               const double* d_base, int64_t n_full, double* d_out, tatva_stream_t stream);
int tatva_reduce_adjoint(const double* d_r_full, const int64_t* d_ptr, const int64_t* d_list,
                         int64_t n_red, double* d_out, tatva_stream_t stream);

/* ---- device-resident CG around the matrix-free HVP (the reference has no solver:
 * its docstrings hand H v to an external one, tatva/mpi.py:594-595) ----------------------------------------
 * Scalars stay on the device: d_scalars[0] = r.r, [1] = p.Ap, [2] = next r.r (>= 8 doubles, slot in 0..7); d_partials holds
 * 1184 per-CTA partial sums.  Dots are two-pass with a fixed summation order (deterministic).
 *   tatva_cg_dot(a, b, ..., slot):  d_scalars[slot] = a.b
 *   tatva_cg_after_matvec:  alpha = s0/(p.Ap); x += alpha p; r -= alpha Ap; beta = (r.r)/s0; p = r + beta p; s0 = r.r */
int tatva_cg_dot(const double* d_a, const double* d_b, int64_t n, double* d_partials,
                 double* d_scalars, int slot, tatva_stream_t stream);
int tatva_cg_after_matvec(double* d_x, double* d_r, double* d_p, const double* d_Ap, int64_t n,
                          double* d_partials, double* d_scalars, tatva_stream_t stream);
/* Leaner iteration (r02): the operator application itself delivers p.Ap (tatva_hvp_lifted_dot) and the direction pass
 * clears the next application's output, so one iteration is  [tatva_hvp_lifted_dot(roll = 1)] + tatva_cg_after_dot:
 * 6 launches and 9 vector passes instead of 8 + memset and 11.  Same arithmetic and summation orders for x, r, p; p.Ap is
 * summed element by element instead of entry by entry (agrees to rounding).  d_minv / d_zero may be NULL.              */
int tatva_cg_after_dot(double* d_x, double* d_r, double* d_p, const double* d_Ap, const double* d_minv, double* d_zero,
                       const int32_t* d_fixed_map, int64_t n, double* d_partials, double* d_scalars,
                       tatva_stream_t stream);
/* d_fixed_map (may be NULL): n int32, < 0 marks a Fixed DOF (Lifter.dof_map) — the iteration then runs on FULL-size
 * vectors with those rows of r held at zero, around the unconstrained kernel:                                          */
int tatva_hvp_dot(tatva_plan_t* plan, int material, const double* params, int n_params, const double* d_u,
                  const double* d_v, double* d_y, int zero_y, int fuse_dot, double* d_partials, double* d_scalars,
                  int slot, int roll, tatva_stream_t stream);
/* Jacobi-preconditioned CG, M = diag(H) from tatva_hessian_diag; z = M^-1 r is folded into the vector kernels and
 * never stored.  d_scalars[0] = r.z, [2] = next r.z, [4] = r.r of the new residual; d_partials >= 2 * 1184 doubles.
 *   tatva_pcg_reciprocal:    minv = 1 / diag (1 where diag is not a positive finite number)
 *   tatva_pcg_start:         p = minv * r ; s0 = r.p
 *   tatva_pcg_after_matvec:  alpha = s0/(p.Ap); x += alpha p; r -= alpha Ap; beta = (r.z)/s0; p = z + beta p; s0 = r.z */
int tatva_pcg_reciprocal(const double* d_diag, int64_t n, double* d_minv, tatva_stream_t stream);
int tatva_pcg_start(double* d_p, const double* d_r, const double* d_minv, int64_t n,
                    double* d_partials, double* d_scalars, tatva_stream_t stream);
int tatva_pcg_after_matvec(double* d_x, double* d_r, double* d_p, const double* d_Ap,
                           const double* d_minv, int64_t n, double* d_partials, double* d_scalars,
                           tatva_stream_t stream);
/* The CG iteration split at its two dot products, for a CG distributed over ranks (SURVEY.md section 8(e):
 * "CG dot products: ncclAllReduce of 1-2 scalars"): the caller all-reduces d_scalars[1] after
 * tatva_cg_dot(p, Ap, .., slot 1) and d_scalars[2] (and [4] with a preconditioner) after tatva_cg_update, then
 * calls tatva_cg_direction; everything stays ordered on one stream.  d_minv may be NULL (no preconditioner).  */
int tatva_cg_update(double* d_x, double* d_r, const double* d_p, const double* d_Ap,
                    const double* d_minv, int64_t n, double* d_partials, double* d_scalars,
                    tatva_stream_t stream);
int tatva_cg_direction(double* d_p, const double* d_r, const double* d_minv, int64_t n,
                       double* d_scalars, tatva_stream_t stream);

/* ---- host-side setup (C++, no GPU needed) -----------------------------------------------
 * pattern_from_mesh / _create_sparse_structure (tatva/sparse/_extraction.py:37-102):
 * two-call protocol: pass indices == NULL to get nnz (indptr is filled), then call again.   */
int tatva_host_pattern_from_mesh(const int32_t* conn, int64_t n_elems, int npe, int64_t n_nodes,
                                 int dofs_per_node, int32_t* indptr, int32_t* indices,
                                 int64_t* nnz);
/* pattern_from_compound (tatva/sparse/_extraction.py:118-245) for any field layout: every element couples all the
 * DOFs of its row of elem_dofs (n_elems, width; -1 = absent); the DOFs listed in `diag` get a diagonal entry (fields
 * that are not nodal).  Sorted unique pairs as CSR; two-call protocol (indices == NULL: indptr and *nnz only).   */
int tatva_host_pattern_from_element_dofs(const int32_t* elem_dofs, int64_t n_elems, int width,
                                         const int32_t* diag, int64_t n_diag, int64_t n_dofs,
                                         int32_t* indptr, int32_t* indices, int64_t* nnz);

/* distance2_colors (tatva-coloring; in-tree spec tatva/sparse/_coloring.py:27-48,:136-153,:270-283):
 * greedy first-fit in natural order on the pattern of A@A.                                  */
int tatva_host_distance2_colors(const int32_t* indptr, const int32_t* indices, int64_t n,
                                int32_t* colors, int32_t* n_colors);
/* Background grid for point location on a plane mesh (HOST arrays).  Two-call protocol: bin_elems == NULL computes
 * lo / inv from the mesh bounding box and fills bin_ptr (nx*ny + 1 prefix sums); the second call fills bin_elems.  */
int tatva_host_build_point_grid(const double* coords, int64_t n_nodes, const int32_t* conn, int64_t n_elems,
                                int npe, int nx, int ny, double* lo, double* inv, int32_t* bin_ptr,
                                int32_t* bin_elems);

/* node -> incident elements in CSR form (ptr: n_nodes+1, list: sum of incidences); pass list == NULL
 * first to size it (ptr[n_nodes]).                                                                   */
int tatva_host_node_to_elements(const int32_t* conn, int64_t n_elems, int npe, int64_t n_nodes,
                                int32_t* ptr, int32_t* list);
int tatva_host_build_tiles(const int32_t* conn, int64_t n_elems, int npe, int tile_elems, int32_t* tile_ptr,
                           int32_t* tile_nodes, uint16_t* local_conn, int32_t* max_unique);
/* Host side of tatva_plan_set_node_schedule (layout in csrc/host.cpp).  Two calls: with tn_node == NULL it fills
 * warp_nodes (128 per tile of 128 elements), warp_local (n_elems * npe), ch_ptr (n_tiles + 1), *n_chunks and *n_ell;
 * the second call fills tn_node (32 * n_chunks), ell_ptr (n_chunks + 1) and ell (n_ell).  cap > 0: a node's
 * contributors are cut into entries of at most `cap` (balances the warps of a tile), 0 = one entry per node.         */
int tatva_host_node_schedule(const int32_t* conn, int64_t n_elems, int npe, int cap, int32_t* warp_nodes, uint8_t* warp_local,
                             int32_t* ch_ptr, int64_t* n_chunks, int64_t* n_ell, int32_t* tn_node, int32_t* ell_ptr,
                             uint16_t* ell);
int tatva_host_csr_element_positions(const int32_t* conn, int64_t n_elems, int npe,
                                     int dofs_per_node, const int32_t* indptr,
                                     const int32_t* indices, int32_t* elem_pos);

/* ---- host probes (no GPU): the kernels' per-element arithmetic executed on the CPU -----------------------------
 * The element tables, geometry, constitutive laws and the modal / pair functions of the Hex8 kernels are compiled for
 * host and device from the same source; these two entry points run one element through them on the CPU so that the
 * kernels' formulas can be verified against the oracle where no GPU is available.  All pointers are HOST pointers.
 *   tatva_probe_element:        generic element body (k_fused / k_hessian_diag): X (npe, dim), u, v (npe, dpn);
 *                               mode 0 energy -> out[0]; 1 residual, 2 HVP, 3 Hessian diagonal -> out (npe, dpn);
 *                               4 = the rank-structured diagonal (k_hessian_diag_rank; laws with a RankLaw only)
 *   tatva_probe_hex8_nh_modal:  the pair kernels of the Hex8 x neo-Hookean path (HVP v3, residual v3, energy v3):
 *                               X, u, v (8, 3); mode 0 energy, 1 residual, 2 HVP -> out[0] or out (8, 3)
 *                               mode 3: the HVP through the geometry-cache arithmetic (point_geometry + point_flux_geo);
 *                               mode 4: Operator.grad of u as k_hex8_grad_modal forms it -> out (8 points, 3, 3);
 *                               mode 5: integration weights as k_hex8_weights_modal -> out (8)
 *   tatva_probe_tet4_nh_ref:    the reference-space Tet4 x neo-Hookean kernels: X, u, v (4, 3); mode 1 residual,
 *                               2 HVP -> out (4, 3)                                                               */
int tatva_probe_element(int element, int material, const double* params, int n_params, int mode,
                        const double* X, const double* u, const double* v, double* out);
int tatva_probe_hex8_nh_modal(int mode, const double* X, const double* u, const double* v, double mu,
                              double lmbda, double* out);
int tatva_probe_tet4_nh_ref(int mode, const double* X, const double* u, const double* v, double mu,
                            double lmbda, double* out);

/* ---- measurement helper: sustained FP64 FMA rate of the device (DFMA microbenchmark) ---- */
int tatva_fp64_peak_tflops(double* tflops, tatva_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TATVA_B200_H */
