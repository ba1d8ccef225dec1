"""Mesh container, structured generators and partition extraction — mirror of tatva/mesh.py.

`Mesh(coords, elements)` keeps the reference's two fields (mesh.py:53-66).  Arrays may be NumPy
or torch; `Operator` moves them to the device once.  Generators: `unit_square` / `rectangle`
(mesh.py:147-231) plus the 3-D boxes the configs need (`box_tet` follows the reference's test
helper tests/test_sparse_tracer.py:29-70; `box_hex` uses node id i + j(nx+1) + k(nx+1)(ny+1) and the
Hexahedron8 node order of element/base.py:478-491).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from enum import Enum
from typing import Any, NamedTuple

import numpy as np


class ElementType(Enum):
    TRIANGLE = "triangle"
    QUAD = "quad"
    TETRAHEDRON = "tetrahedron"
    HEXAHEDRON = "hexahedron"


class PartitionInfo(NamedTuple):
    """mesh.py:38-47."""

    nodes_local_to_global: np.ndarray
    n_owned_nodes: int


def _np(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    return np.asarray(a)


@dataclass(frozen=True)
class Mesh:
    coords: Any
    """(n_nodes, n_dim) float64"""
    elements: Any
    """(n_elements, nodes_per_element) int32"""

    def _replace(self, **changes):
        return replace(self, **changes)

    def set_coords(self, new_coords):
        return replace(self, coords=new_coords)

    # -- element sizes (mesh.py:87-144): setup-time host arithmetic -------------------------------
    def _element_circumdiameters(self) -> np.ndarray:
        """Circumscribed-sphere diameter of simplices (triangles in 2-D or embedded in 3-D, tetrahedra); for every
        other element the largest vertex-to-vertex distance."""
        X = _np(self.coords).astype(np.float64)[_np(self.elements)]  # (E, npe, dim)
        npe, dim = X.shape[1], X.shape[2]
        tiny = 1e-12
        if npe == 3:
            e01, e02, e12 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 2] - X[:, 1]
            edge_product = np.linalg.norm(e01, axis=1) * np.linalg.norm(e02, axis=1) * np.linalg.norm(e12, axis=1)
            if dim == 2:
                twice_area = np.abs(e01[:, 0] * e02[:, 1] - e01[:, 1] * e02[:, 0])
            else:
                twice_area = np.linalg.norm(np.cross(e01, e02), axis=1)
            return edge_product / np.maximum(twice_area, tiny)  # 2R = abc / (2 area)
        if npe == 4 and dim == 3:
            p, q, r = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]
            qr, rp, pq = np.cross(q, r), np.cross(r, p), np.cross(p, q)
            six_volume = np.abs(np.einsum("ei,ei->e", p, qr))
            w = (p * p).sum(1)[:, None] * qr + (q * q).sum(1)[:, None] * rp + (r * r).sum(1)[:, None] * pq
            return np.linalg.norm(w, axis=1) / np.maximum(six_volume, tiny)  # 2R = |p^2 (q x r) + ...| / (6 V)
        i, j = np.triu_indices(npe, k=1)
        return np.sqrt(((X[:, j] - X[:, i]) ** 2).sum(-1).max(axis=1))

    def hmin(self):
        """Smallest element diameter (mesh.py:138-140)."""
        return self._element_circumdiameters().min()

    def hmax(self):
        """Largest element diameter (mesh.py:142-144)."""
        return self._element_circumdiameters().max()

    # -- generators ---------------------------------------------------------------------------
    @classmethod
    def unit_square(cls, n_x, n_y, *, type=ElementType.TRIANGLE, dim=2):
        return cls.rectangle((0.0, 1.0), (0.0, 1.0), n_x, n_y, type=type, dim=dim)

    @classmethod
    def rectangle(cls, x, y, n_x, n_y, *, type=ElementType.TRIANGLE, dim=2):
        """mesh.py:156-231: node id = i (n_y+1) + j; triangles [n0,n1,n3],[n0,n3,n2]; quads CCW."""
        xv, yv = np.meshgrid(np.linspace(x[0], x[1], n_x + 1), np.linspace(y[0], y[1], n_y + 1), indexing="ij")
        coords = np.stack([xv.ravel(), yv.ravel()], axis=-1)
        i, j = np.meshgrid(np.arange(n_x), np.arange(n_y), indexing="ij")
        n0 = (i * (n_y + 1) + j).ravel()
        n1, n2, n3 = n0 + (n_y + 1), n0 + 1, n0 + (n_y + 1) + 1
        kind = ElementType(type)
        if kind is ElementType.TRIANGLE:
            el = np.stack([np.stack([n0, n1, n3], -1), np.stack([n0, n3, n2], -1)], axis=1).reshape(-1, 3)
        elif kind is ElementType.QUAD:
            el = np.stack([n0, n1, n3, n2], -1)
        else:
            raise NotImplementedError(f"Element type {type} not implemented.")
        if dim == 3:
            coords = np.hstack([coords, np.zeros((coords.shape[0], 1))])
        return cls(coords=coords, elements=el.astype(np.int32))

    @classmethod
    def box_tet(cls, lengths, nb_elems):
        """6 tetrahedra per cell, as tests/test_sparse_tracer.py:29-70."""
        (lx, ly, lz), (nx, ny, nz) = lengths, nb_elems
        xr = np.linspace(-lx / 2, lx / 2, nx + 1)
        yr = np.linspace(-ly / 2, ly / 2, ny + 1)
        zr = np.linspace(0, lz, nz + 1)
        Z, Y, X = np.meshgrid(zr, yr, xr, indexing="ij")
        nodes = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
        sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        n0 = (i * sx + j * sy + k * sz).ravel()
        n1, n2 = n0 + sx, n0 + sy
        n3, n4 = n2 + sx, n0 + sz
        n5, n6 = n4 + sx, n4 + sy
        n7 = n6 + sx
        corner_sets = [(n0, n1, n3, n7), (n0, n1, n7, n5), (n0, n5, n7, n4), (n0, n3, n2, n7), (n0, n2, n6, n7), (n0, n6, n4, n7)]
        tets = np.stack([np.stack(t, -1) for t in corner_sets], axis=1).reshape(-1, 4)
        return cls(coords=nodes, elements=tets.astype(np.int32))

    @classmethod
    def box_hex(cls, nb_elems, lengths=None):
        """Structured Hex8 box on [0,lx]x[0,ly]x[0,lz]; elements and nodes lexicographic, x fastest."""
        nx, ny, nz = (nb_elems,) * 3 if np.isscalar(nb_elems) else nb_elems
        if lengths is None:
            m = max(nx, ny, nz)
            lengths = (nx / m, ny / m, nz / m)
        xr, yr, zr = (np.linspace(0.0, L, n + 1) for L, n in zip(lengths, (nx, ny, nz)))
        Z, Y, X = np.meshgrid(zr, yr, xr, indexing="ij")
        nodes = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
        sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        n0 = (i * sx + j * sy + k * sz).ravel()
        el = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + sz, n0 + sz + sx, n0 + sz + sx + sy, n0 + sz + sy], -1)
        return cls(coords=nodes, elements=el.astype(np.int32))


def extract_local_mesh(mesh_global: Mesh, element_partition, part: int) -> tuple[Mesh, PartitionInfo]:
    """Local mesh of partition `part`, owned nodes first then ghosts (mesh.py:234-291).

    A node is owned by the smallest partition id among the elements touching it (:263-265);
    owned and ghost blocks are each ascending in global id (:258, :267-273)."""
    elements = _np(mesh_global.elements)
    coords = _np(mesh_global.coords)
    element_partition = np.asarray(element_partition)
    local_el = elements[element_partition == part]
    present = np.unique(local_el.ravel())
    owner = np.full(len(coords), element_partition.max() + 1, dtype=np.int32)
    for col in range(elements.shape[1]):
        np.minimum.at(owner, elements[:, col], element_partition)
    mine = owner[present] == part
    l2g = np.concatenate([present[mine], present[~mine]])
    g2l = np.full(len(coords), -1, dtype=np.int32)
    g2l[l2g] = np.arange(len(l2g), dtype=np.int32)
    return (
        Mesh(coords=coords[l2g], elements=g2l[local_el]),
        PartitionInfo(nodes_local_to_global=l2g, n_owned_nodes=int(mine.sum())),
    )


def block_partition(nb_elems, nparts):
    """Cartesian block partition of a structured box (x fastest element order): 2 -> 2x1x1,
    4 -> 2x2x1, 8 -> 2x2x2.  Returns an int32 partition id per element."""
    nx, ny, nz = nb_elems
    grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(nparts)
    if grid is None:
        raise ValueError("nparts must be 1, 2, 4 or 8")
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    px = (i * grid[0]) // nx
    py = (j * grid[1]) // ny
    pz = (k * grid[2]) // nz
    return (px + grid[0] * (py + grid[1] * pz)).ravel().astype(np.int32)


def _spread_bits(x: np.ndarray, dim: int) -> np.ndarray:
    """Insert dim-1 zero bits between the low 21 bits of x (Morton / Z-order interleave)."""
    x = x.astype(np.uint64) & np.uint64((1 << 21) - 1)
    if dim == 3:
        x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    else:
        x = (x | (x << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        x = (x | (x << np.uint64(2))) & np.uint64(0x3333333333333333)
        x = (x | (x << np.uint64(1))) & np.uint64(0x5555555555555555)
    return x


def locality_order(coords, elements) -> np.ndarray:
    """Permutation that sorts the elements along a Morton (Z-order) curve through their centroids —
    the locality-preserving ordering the element-per-thread kernels want: consecutive threads then touch
    nearby nodes, so the gathers of a warp share L1/L2 lines.  Returns `perm` with new element k = old
    element perm[k]."""
    c = _np(coords)
    el = _np(elements)
    dim = c.shape[1]
    cen = c[el].mean(axis=1)
    lo, hi = cen.min(axis=0), cen.max(axis=0)
    q = ((cen - lo) / np.where(hi > lo, hi - lo, 1.0) * ((1 << 21) - 1)).astype(np.uint64)
    key = np.zeros(len(el), dtype=np.uint64)
    for d in range(min(dim, 3)):
        key |= _spread_bits(q[:, d], 3 if dim >= 3 else 2) << np.uint64(d)
    return np.argsort(key, kind="stable")


def renumber_nodes_by_first_touch(elements) -> np.ndarray:
    """node_perm with new node k = old node node_perm[k], nodes numbered in the order the (already
    locality-sorted) elements first touch them; unreferenced nodes keep their relative order at the end."""
    el = _np(elements)
    flat = el.ravel()
    _, first = np.unique(flat, return_index=True)
    touched = flat[np.sort(first)]
    n = int(flat.max()) + 1
    rest = np.setdiff1d(np.arange(n), touched, assume_unique=False)
    return np.concatenate([touched, rest])


def reorder_mesh(mesh: Mesh, nodes: bool = True):
    """Locality-reordered copy of `mesh`: elements along a Morton curve and (optionally) nodes by first touch.
    Returns (new_mesh, elem_perm, node_perm): new element k = old elem_perm[k]; new node k = old node_perm[k]
    (nodal vectors move as u_new = u_old[node_perm], results back as r_old[node_perm] = r_new)."""
    c, el = _np(mesh.coords), _np(mesh.elements)
    elem_perm = locality_order(c, el)
    el_sorted = el[elem_perm]
    if not nodes:
        return Mesh(coords=c, elements=el_sorted), elem_perm, np.arange(c.shape[0])
    node_perm = renumber_nodes_by_first_touch(el_sorted)
    if node_perm.size < c.shape[0]:
        node_perm = np.concatenate([node_perm, np.arange(node_perm.size, c.shape[0])])
    inv = np.empty(c.shape[0], dtype=np.int64)
    inv[node_perm] = np.arange(c.shape[0])
    return Mesh(coords=c[node_perm], elements=inv[el_sorted].astype(el.dtype)), elem_perm, node_perm


def find_containing_polygons(points, polygons, device=None):
    """Index of the first polygon (vertex loops, shape (n_polygons, n_vertices, 2)) containing each 2-D point, -1 if
    none — tatva/mesh.py:294-388 (bounding box, then on-boundary OR odd ray crossings).  Runs the point-location
    part of the `tatva_op_interpolate` kernel on the polygons as a stand-alone mesh; loops of 3, 4, 6 or 8 vertices
    (the plane elements).  Returns an int32 CUDA tensor."""
    import ctypes as C

    import torch

    from . import _lib

    if not torch.cuda.is_available():
        raise _lib.TatvaError("find_containing_polygons needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    poly = torch.as_tensor(polygons, dtype=torch.float64, device=dev)
    pts = torch.as_tensor(points, dtype=torch.float64, device=dev).contiguous()
    if poly.ndim != 3 or poly.shape[2] != 2 or pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError("points must be (n_points, 2) and polygons (n_polygons, n_vertices, 2)")
    kinds = {3: _lib.TRI3, 4: _lib.QUAD4, 6: _lib.TRI6, 8: _lib.QUAD8}
    nv = int(poly.shape[1])
    if nv not in kinds:
        raise NotImplementedError("polygons with 3, 4, 6 or 8 vertices are supported")
    n_poly = int(poly.shape[0])
    out = torch.full((pts.shape[0],), -1, dtype=torch.int32, device=dev)
    if n_poly == 0 or pts.shape[0] == 0:
        return out
    coords = poly.reshape(-1, 2).contiguous()
    conn = torch.arange(n_poly * nv, dtype=torch.int32, device=dev).reshape(n_poly, nv)
    L = _lib.lib()
    handle = C.c_void_p()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.tatva_plan_create(C.byref(handle), kinds[nv], coords.shape[0], n_poly, coords.data_ptr(), conn.data_ptr(), 0, st), "tatva_plan_create")
        try:
            u = torch.zeros((coords.shape[0], 1), dtype=torch.float64, device=dev)
            vals = torch.empty((pts.shape[0], 1), dtype=torch.float64, device=dev)
            _lib.check(L.tatva_op_interpolate(handle, u.data_ptr(), 1, pts.data_ptr(), pts.shape[0], vals.data_ptr(), out.data_ptr(), st), "tatva_op_interpolate")
            torch.cuda.current_stream().synchronize()  # coords / conn are temporaries of this call
        finally:
            L.tatva_plan_destroy(handle)
    return out
