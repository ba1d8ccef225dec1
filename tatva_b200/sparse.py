"""Sparse Jacobians — mirror of tatva.sparse (tatva/sparse/base.py, _extraction.py, _coloring.py).

* `pattern_from_mesh` / `pattern_from_compound`: bit-exact CSR pattern (int32, sorted columns, int8
  ones), built by the C++ host routine `tatva_host_pattern_from_mesh`.
* `distance2_colors`: greedy first-fit distance-2 colouring in natural DOF order (the in-tree
  spec tatva/sparse/_coloring.py of the external `tatva-coloring` package), C++ host routine.
* `ColoredMatrix`: same fields as the reference (data, indptr, indices, shape, colors).
* `jacfwd` / `linearized_jacfwd`: when `fn` is a fused residual (`op.residual(material)`), the
  Jacobian is assembled by ONE kernel that adds every element stiffness straight into the fixed
  CSR pattern (`tatva_csr_assemble`) — no n_colors HVP sweeps, no (N, n_colors) temporary.  For any
  other `fn` the reference algorithm is kept: one forward-mode JVP per colour with a 0/1 seed and
  the decompression data[k] = J_c[row(k), colors[indices[k]]] (sparse/base.py:108-176, :230-270).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import Callable

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .mesh import Mesh, _np


def _i32p(a):
    return a.ctypes.data_as(_lib.c_i32p)


def pattern_arrays(elements, n_nodes: int, n_dofs_per_node: int) -> tuple[np.ndarray, np.ndarray]:
    """(indptr, indices) int32 of the element-coupling pattern; DOF id = node * dpn + comp."""
    conn = np.ascontiguousarray(_np(elements), dtype=np.int32)
    n_elems, npe = conn.shape
    n = n_nodes * n_dofs_per_node
    indptr = np.empty(n + 1, dtype=np.int32)
    nnz = C.c_int64()
    L = _lib.lib()
    _lib.check(L.tatva_host_pattern_from_mesh(_i32p(conn), n_elems, npe, n_nodes, n_dofs_per_node, _i32p(indptr), None, C.byref(nnz)), "pattern_from_mesh")
    indices = np.empty(nnz.value, dtype=np.int32)
    _lib.check(L.tatva_host_pattern_from_mesh(_i32p(conn), n_elems, npe, n_nodes, n_dofs_per_node, _i32p(indptr), _i32p(indices), C.byref(nnz)), "pattern_from_mesh")
    return indptr, indices


def pattern_from_mesh(mesh: Mesh, n_dofs_per_node: int) -> sp.csr_matrix:
    """tatva/sparse/_extraction.py:91-102."""
    n_nodes = mesh.coords.shape[0]
    indptr, indices = pattern_arrays(mesh.elements, n_nodes, n_dofs_per_node)
    n = n_nodes * n_dofs_per_node
    return sp.csr_matrix((np.ones(indices.shape[0], dtype=np.int8), indices, indptr), shape=(n, n))


def _pattern_from_element_dofs(elem_dofs: np.ndarray, diag: np.ndarray, n: int):
    """CSR (indptr, indices) of the pattern coupling, per element, all DOFs of its row (-1 = absent), plus diagonal
    entries for `diag`: the result of np.unique(row * n + col) over all pairs (_extraction.py:74-79, :207-216)."""
    L = _lib.lib()
    i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
    E, w = (int(elem_dofs.shape[0]), int(elem_dofs.shape[1])) if elem_dofs.size else (0, 0)
    indptr = np.zeros(n + 1, dtype=np.int32)
    nnz = C.c_int64()
    args = (i32(elem_dofs) if E * w else None, E, w, i32(diag) if diag.size else None, int(diag.size), n, i32(indptr))
    _lib.check(L.tatva_host_pattern_from_element_dofs(*args, None, C.byref(nnz)), "tatva_host_pattern_from_element_dofs")
    indices = np.empty(nnz.value, dtype=np.int32)
    if nnz.value:
        _lib.check(L.tatva_host_pattern_from_element_dofs(*args, i32(indices), C.byref(nnz)), "tatva_host_pattern_from_element_dofs")
    return indptr, indices


def pattern_from_compound(compound_cls, block_wise: bool = False):
    """tatva/sparse/_extraction.py:118-245: nodal fields coupled within elements, every other field
    diagonal.  The element -> DOF lists are assembled here; the sorted unique pair set is built in C++
    (`tatva_host_pattern_from_element_dofs`, one sorted row per DOF in parallel) instead of sorting all
    (npe * dpn)^2 * E pairs (38 s -> 0.5 s at config 5)."""
    from .compound import CompoundError, Nodal

    if compound_cls._mesh is None:
        raise CompoundError("Mesh must be set on Compound class to create sparsity pattern.")
    mesh = compound_cls._mesh
    n_nodes = mesh.coords.shape[0]
    elements = _np(mesh.elements).astype(np.int64)
    coupled, diagonal = [], []
    for name, f in compound_cls.fields:
        ft = f.field_type.get()
        idx = np.asarray(f.indices(slice(None)))
        if isinstance(ft, Nodal):
            n_items = len(ft.node_ids) if ft.node_ids is not None else n_nodes
            if n_items == 0:
                continue
            per = idx.size // n_items
            node_dofs = np.full((n_nodes, per), -1, dtype=np.int64)
            if ft.node_ids is None:
                node_dofs[:] = idx.reshape(n_nodes, per)
            else:
                ids = np.asarray(ft.node_ids, dtype=np.int64)
                ok = (ids >= 0) & (ids < n_nodes)
                node_dofs[ids[ok]] = idx.reshape(-1, per)[ok]
            coupled.append(node_dofs[elements].reshape(elements.shape[0], -1))
        else:
            import warnings

            warnings.warn(
                f"Custom space detected for field '{name}'. Only diagonal entries added to the sparsity pattern. "
                "Please provide your own sparsity pattern if cross-coupling is required.",
                UserWarning,
            )
            diagonal.append(idx.ravel())
    n = compound_cls.size
    ed = np.ascontiguousarray(np.concatenate(coupled, axis=1), dtype=np.int32) if coupled else np.zeros((0, 0), dtype=np.int32)
    dg = np.ascontiguousarray(np.concatenate(diagonal), dtype=np.int32) if diagonal else np.zeros(0, dtype=np.int32)
    indptr, indices = _pattern_from_element_dofs(ed, dg, n)
    full = sp.csr_matrix((np.ones(indices.shape[0], dtype=np.int8), indices, indptr), shape=(n, n))
    if not block_wise:
        return full
    slices, seen = [], set()
    for _, f in compound_cls.fields:
        s = getattr(f, "_root_slice", getattr(f, "_slice", None))
        if s is not None and (s.start, s.stop) not in seen:
            slices.append(s)
            seen.add((s.start, s.stop))
    slices.sort(key=lambda s: s.start)
    return [[full[a, b] for b in slices] for a in slices]


def distance2_colors(row_ptr, col_idx, n_dofs: int) -> np.ndarray:
    """Greedy distance-2 colouring (natural order, first fit) — tatva/sparse/_coloring.py:270-283."""
    indptr = np.ascontiguousarray(_np(row_ptr), dtype=np.int32)
    indices = np.ascontiguousarray(_np(col_idx), dtype=np.int32)
    colors = np.empty(n_dofs, dtype=np.int32)
    nc = C.c_int32()
    _lib.check(_lib.lib().tatva_host_distance2_colors(_i32p(indptr), _i32p(indices), n_dofs, _i32p(colors), C.byref(nc)), "distance2_colors")
    return colors


def distance2_color_and_seeds(row_ptr, col_idx, n_dofs: int):
    """tatva/sparse/_coloring.py:337-351: colours and the int32 0/1 seed matrix (n_colors, n_dofs)."""
    colors = distance2_colors(row_ptr, col_idx, n_dofs)
    seeds = (colors[None, :] == np.unique(colors)[:, None]).astype(np.int32)
    return colors, seeds


@dataclass(frozen=True)
class ColoredMatrix:
    """tatva/sparse/base.py:37-105."""

    data: object
    indptr: object
    indices: object
    shape: tuple
    colors: object

    @classmethod
    def from_csr(cls, csr_matrix: sp.csr_matrix, colors=None):
        indptr, indices = csr_matrix.indptr, csr_matrix.indices
        if colors is None:
            colors = distance2_colors(indptr, indices, csr_matrix.shape[0])
        return cls(data=np.asarray(csr_matrix.data), indptr=np.asarray(indptr), indices=np.asarray(indices), shape=tuple(csr_matrix.shape), colors=np.asarray(colors))

    def to_csr(self) -> sp.csr_matrix:
        return sp.csr_matrix((_np(self.data), _np(self.indices), _np(self.indptr)), shape=self.shape)

    def to_dense(self) -> np.ndarray:
        return self.to_csr().toarray()

    def _replace(self, **kw):
        return replace(self, **kw)


def compute_rows_cols(colored_matrix: ColoredMatrix):
    """tatva/sparse/base.py:108-136: row index and column colour of every stored entry."""
    indptr, indices, colors = _np(colored_matrix.indptr), _np(colored_matrix.indices), _np(colored_matrix.colors)
    rows = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return rows, colors[indices]


class _Assembler:
    """Direct CSR assembly plan for (operator, material, pattern): device copies of indptr and of the
    element -> CSR position table."""

    def __init__(self, op, material, colored_matrix: ColoredMatrix, by_rows: bool | None = None, symmetric: bool = False, tiled: bool | None = None):
        dpn = material.dofs_per_node(op.dim)
        indptr = np.ascontiguousarray(_np(colored_matrix.indptr), dtype=np.int32)
        indices = np.ascontiguousarray(_np(colored_matrix.indices), dtype=np.int32)
        if indptr.shape[0] - 1 != op.n_nodes * dpn:
            raise ValueError("sparsity pattern size does not match n_nodes * dofs_per_node")
        conn = np.ascontiguousarray(op.elements_fused.cpu().numpy(), dtype=np.int32)  # the order the fused kernels run in
        pos = np.empty((conn.shape[0], conn.shape[1], conn.shape[1]), dtype=np.int32)
        _lib.check(
            _lib.lib().tatva_host_csr_element_positions(_i32p(conn), conn.shape[0], conn.shape[1], dpn, _i32p(indptr), _i32p(indices), _i32p(pos)),
            "csr_element_positions (pattern must contain every element coupling, node-blocked)",
        )
        self.op, self.material, self.nnz = op, material, int(indices.shape[0])
        self.d_indptr = torch.as_tensor(indptr, device=op.device)
        # Tiled kernel with on-chip combination of duplicate blocks (r02 default where it exists): constant-gradient
        # elements with the neo-Hookean / linear-elastic law, default quadrature rule.
        can_tile = (
            op.element.kind in (_lib.TRI3, _lib.TET4) and not getattr(op, "_custom_rule", False) and not by_rows and not symmetric
            and ((material.material_id == _lib.NEO_HOOKEAN and op.element.kind == _lib.TET4) or material.material_id == _lib.LINEAR_ELASTIC)
        )
        self.tiled = can_tile if tiled is None else (bool(tiled) and can_tile)
        if tiled and not can_tile:
            raise NotImplementedError("tiled assembly: Tri3 / Tet4 with NeoHookean or LinearElastic, default rule")
        if self.tiled:
            self._build_tiles(op, conn, dpn, indptr, indices)
        # Row-wise (atomic-free, deterministic) kernel for single-point elements; per-entry atomics otherwise.
        self.by_rows = False if by_rows is None else bool(by_rows)
        if self.by_rows and op.nq != 1:
            raise NotImplementedError("row-wise assembly is implemented for single-point elements (Tri3, Tet4)")
        if self.by_rows:
            L = _lib.lib()
            ptr = np.empty(op.n_nodes + 1, dtype=np.int32)
            _lib.check(L.tatva_host_node_to_elements(_i32p(conn), conn.shape[0], conn.shape[1], op.n_nodes, _i32p(ptr), None), "node_to_elements")
            lst = np.empty(int(ptr[-1]), dtype=np.int32)
            _lib.check(L.tatva_host_node_to_elements(_i32p(conn), conn.shape[0], conn.shape[1], op.n_nodes, _i32p(ptr), _i32p(lst)), "node_to_elements")
            self.d_indices = torch.as_tensor(indices, device=op.device)
            self.d_n2e_ptr = torch.as_tensor(ptr, device=op.device)
            self.d_n2e = torch.as_tensor(lst[: int(ptr[-1])], device=op.device)
        else:
            self.d_pos = torch.as_tensor(pos, device=op.device)
            # symmetric = True: REDs for the upper triangle only + mirror pass (valid: all fused laws are energies)
            self.symmetric = bool(symmetric)
            if self.symmetric:
                self.d_indices = torch.as_tensor(indices, device=op.device)

    def _build_tiles(self, op, conn, dpn, indptr, indices):
        """Locality-sorted element list + the per-tile combine schedule (host C++, once per pattern)."""
        from .mesh import locality_order

        L = _lib.lib()
        perm = locality_order(op.coords.cpu().numpy(), conn)
        csort = np.ascontiguousarray(conn[perm], dtype=np.int32)
        E, npe = csort.shape
        pos = np.empty((E, npe, npe), dtype=np.int32)
        _lib.check(L.tatva_host_csr_element_positions(_i32p(csort), E, npe, dpn, _i32p(indptr), _i32p(indices), _i32p(pos)), "csr_element_positions")
        tile = 128
        n_tiles = (E + tile - 1) // tile
        blk_ptr = np.empty(n_tiles + 1, dtype=np.int32)
        n_blk, n_con = C.c_int64(), C.c_int64()
        args = (_i32p(csort), E, npe, dpn, tile, _i32p(indptr), _i32p(pos), _i32p(blk_ptr), C.byref(n_blk), C.byref(n_con))
        _lib.check(L.tatva_host_csr_tile_schedule(*args, None, None, None, None, None, None), "csr_tile_schedule")
        nb, nc = int(n_blk.value), int(n_con.value)
        base, rl, base_t, rl_t = (np.empty(nb, dtype=np.int32) for _ in range(4))
        con_ptr = np.empty(nb + 1, dtype=np.int32)
        con = np.empty(nc, dtype=np.uint32)
        _lib.check(L.tatva_host_csr_tile_schedule(*args, _i32p(base), _i32p(rl), _i32p(base_t), _i32p(rl_t), _i32p(con_ptr), con.ctypes.data_as(C.POINTER(C.c_uint32))), "csr_tile_schedule")
        dev = lambda a: torch.as_tensor(a.view(np.int32) if a.dtype == np.uint32 else a, device=op.device)  # noqa: E731
        self._tile = dict(conn=dev(csort), blk_ptr=dev(blk_ptr), blk_base=dev(base), blk_rowlen=dev(rl), blk_base_t=dev(base_t), blk_rowlen_t=dev(rl_t), con_ptr=dev(con_ptr), con=dev(con))
        self.tile_stats = dict(n_tiles=n_tiles, distinct_blocks_upper=nb, contributions_upper=nc, contributions_all=int(E * npe * npe), combine_ratio=nc / nb)

    def __call__(self, u, out=None) -> torch.Tensor:
        op = self.op
        uc = op._as_dev(u).contiguous()
        if out is None:
            out = torch.empty(self.nnz, dtype=torch.float64, device=op.device)
        prm, n = _lib.params_array(self.material.params())
        if self.tiled:
            t = self._tile
            op._call("tatva_csr_assemble_tiled", self.material.material_id, prm, n, uc.data_ptr(), t["conn"].data_ptr(), t["blk_ptr"].data_ptr(), t["blk_base"].data_ptr(),
                     t["blk_rowlen"].data_ptr(), t["blk_base_t"].data_ptr(), t["blk_rowlen_t"].data_ptr(), t["con_ptr"].data_ptr(), t["con"].data_ptr(), self.nnz, out.data_ptr())
            return out
        if self.by_rows:
            op._call("tatva_csr_assemble_rows", self.material.material_id, prm, n, uc.data_ptr(), self.d_indptr.data_ptr(), self.d_indices.data_ptr(),
                     self.d_n2e_ptr.data_ptr(), self.d_n2e.data_ptr(), out.data_ptr())
        elif self.symmetric:
            op._call("tatva_csr_assemble_sym", self.material.material_id, prm, n, uc.data_ptr(), self.d_indptr.data_ptr(), self.d_indices.data_ptr(), self.d_pos.data_ptr(), self.nnz, out.data_ptr())
        else:
            op._call("tatva_csr_assemble", self.material.material_id, prm, n, uc.data_ptr(), self.d_indptr.data_ptr(), self.d_pos.data_ptr(), self.nnz, out.data_ptr())
        return out


def assembler(op, material, colored_matrix: ColoredMatrix, by_rows: bool | None = None, symmetric: bool = False, tiled: bool | None = None) -> Callable:
    """u -> CSR data (nnz,) of d^2E/du^2 on the pattern of `colored_matrix` (one kernel).
    Default: element-per-thread kernel with sector-grouped FP64 REDs for every entry (0.46 ms at config 2).
    symmetric=True adds only the upper triangle by RED and mirrors the lower one (measured slower on B200: 0.61 ms —
    the RED loop is issue-bound, not L2-bound, so halving the active lanes does not pay for the mirror pass).
    by_rows=True selects the atomic-free row-wise kernel (Tri3, Tet4): bitwise reproducible, ~1.6x slower."""
    return _Assembler(op, material, colored_matrix, by_rows, symmetric, tiled)


def _coloured_columns(fn_jvp, u, colored_matrix, color_batch_size):
    """Reference algorithm (sparse/base.py:230-270) without the dense (N, n_colors) temporary: each
    colour's JVP column is decompressed straight into `data`."""
    colors = _np(colored_matrix.colors)
    n_colors = int(colors.max()) + 1
    rows, col_colors = compute_rows_cols(colored_matrix)
    dev = u.device
    d_colors = torch.as_tensor(colors, device=dev)
    order = np.argsort(col_colors, kind="stable")
    bounds = np.searchsorted(col_colors[order], np.arange(n_colors + 1))
    d_order = torch.as_tensor(order, device=dev)
    d_rows = torch.as_tensor(rows[order], device=dev)
    data = torch.zeros(rows.shape[0], dtype=u.dtype, device=dev)
    for c in range(n_colors):
        seed = (d_colors == c).to(u.dtype).reshape(u.shape)
        col = fn_jvp(seed).reshape(-1)
        lo, hi = int(bounds[c]), int(bounds[c + 1])
        data[d_order[lo:hi]] = col[d_rows[lo:hi]]
    return data


def _forward_mode(f, u):
    """seed -> J(u) seed with torch's forward-mode AD (dual tensors) — the jax.jvp of sparse/base.py:264.
    The kernels' autograd Functions implement `jvp`, so this dispatches to the tangent (HVP) kernels."""
    import torch.autograd.forward_ad as fwAD

    def jvp(seed):
        with fwAD.dual_level():
            out = f(fwAD.make_dual(u, seed))
            tangent = fwAD.unpack_dual(out).tangent
            return torch.zeros_like(out) if tangent is None else tangent.clone()

    return jvp


def _direct_assembler(fn, colored_matrix):
    """The one-kernel assembler for a fused residual, or None when the pattern is not node-blocked (the direct kernel
    writes (a, b) blocks at one offset in all dpn rows of node a; `tatva_host_csr_element_positions` verifies that).
    The caller then runs the reference's coloured algorithm: one fused HVP kernel per colour (sparse/base.py:230-270)."""
    if fn.material.material_id >= _lib.USER_LAW_BASE:  # run-time compiled laws have no assembly kernel: coloured route
        return None
    try:
        return _Assembler(fn.op, fn.material, colored_matrix)
    except (_lib.TatvaError, ValueError):
        return None


def jacfwd(fn: Callable, colored_matrix: ColoredMatrix, *, color_batch_size: int | None = None) -> Callable:
    """tatva/sparse/base.py:139-176.  `fn(u, *args)` returns the residual; the result is a new
    ColoredMatrix whose `data` holds d fn / d u on the pattern."""
    from .operator import FusedResidual

    if isinstance(fn, FusedResidual):
        asm = _direct_assembler(fn, colored_matrix)
        if asm is not None:
            return lambda u: replace(colored_matrix, data=asm(u))

    def _wrapped(u, *args, **kwargs):
        ut = u if isinstance(u, torch.Tensor) else torch.as_tensor(np.asarray(u), device="cuda")

        jvp = _forward_mode(lambda x: fn(x, *args, **kwargs), ut)
        return replace(colored_matrix, data=_coloured_columns(jvp, ut, colored_matrix, color_batch_size))

    return _wrapped


def linearized_jacfwd(fn: Callable, colored_matrix: ColoredMatrix, *, color_batch_size: int | None = None) -> Callable:
    """tatva/sparse/base.py:179-227: (primal, Jacobian) sharing the forward pass."""
    from .operator import FusedResidual

    if isinstance(fn, FusedResidual):
        asm = _direct_assembler(fn, colored_matrix)
        if asm is not None:
            return lambda u: (fn(u), replace(colored_matrix, data=asm(u)))

    def _wrapped(u, *args, **kwargs):
        ut = u if isinstance(u, torch.Tensor) else torch.as_tensor(np.asarray(u), device="cuda")
        f = lambda x: fn(x, *args, **kwargs)  # noqa: E731
        # torch.func.linearize traces with fake tensors, which cannot pass through the C-ABI kernels;
        # the primal is evaluated once and each colour is a forward-mode JVP.
        primal = f(ut)
        jvp = _forward_mode(f, ut)
        return primal, replace(colored_matrix, data=_coloured_columns(jvp, ut, colored_matrix, color_batch_size))

    return _wrapped


__all__ = [
    "ColoredMatrix",
    "jacfwd",
    "linearized_jacfwd",
    "pattern_from_mesh",
    "pattern_from_compound",
    "distance2_colors",
    "distance2_color_and_seeds",
    "assembler",
    "compute_rows_cols",
]
