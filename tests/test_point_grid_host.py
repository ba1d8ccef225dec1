"""Host builder of the point-location background grid (tatva_host_build_point_grid): searching only the elements
listed in a point's bin must give the element mesh.find_containing_polygons returns (reference mesh.py:294-388),
for interior points, points on shared edges and nodes, and points outside the mesh."""
import numpy as np
import pytest

from oracle import tatva_oracle as orc
from tatva_b200 import _lib


def _bin(x, lo, inv, n):
    t = (x - lo) * inv
    return 0 if not t > 0 else (int(t) if t < n else n - 1)


@pytest.mark.parametrize("kind", ["tri3", "quad4", "tri6", "quad8"])
def test_grid_search_equals_full_scan(kind):
    L = _lib.lib()
    rng = np.random.default_rng(0)
    c, el = {"tri3": lambda: orc.mesh_unit_square_tri(13, 9), "quad4": lambda: orc.mesh_unit_square_quad(7, 11), "tri6": lambda: orc.mesh_second_order("tri6", 6, 5), "quad8": lambda: orc.mesh_second_order("quad8", 5, 6)}[kind]()
    inner = (c[:, 0] > 1e-9) & (c[:, 0] < 1 - 1e-9) & (c[:, 1] > 1e-9) & (c[:, 1] < 1 - 1e-9)
    c = np.ascontiguousarray(c + 0.02 * rng.uniform(-1, 1, c.shape) * inner[:, None])
    el = np.ascontiguousarray(el, dtype=np.int32)
    side = int(round(np.sqrt(len(el))))
    lo, inv = np.zeros(2), np.zeros(2)
    ptr = np.zeros(side * side + 1, dtype=np.int32)
    f64 = lambda a: a.ctypes.data_as(_lib.c_f64p)  # noqa: E731
    i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
    args = (f64(c), len(c), i32(el), len(el), el.shape[1], side, side, f64(lo), f64(inv), i32(ptr))
    assert L.tatva_host_build_point_grid(*args, None) == 0
    np.testing.assert_allclose(lo, c.min(axis=0))
    elems = np.empty(ptr[-1], dtype=np.int32)
    assert L.tatva_host_build_point_grid(*args, i32(elems)) == 0
    pts = np.concatenate([rng.uniform(-0.1, 1.1, size=(1500, 2)), c[:20], 0.5 * (c[el[:30, 0]] + c[el[:30, 1]]), c.max(axis=0)[None], c.min(axis=0)[None]])
    ref = orc.find_containing_polygons(pts, c[el])
    got = np.full(len(pts), -1)
    for i, (px, py) in enumerate(pts):
        b = _bin(py, lo[1], inv[1], side) * side + _bin(px, lo[0], inv[0], side)
        cand = elems[ptr[b] : ptr[b + 1]]
        assert np.all(np.diff(cand) > 0), "bin lists must be ascending"
        if len(cand):
            r = orc.find_containing_polygons(pts[i : i + 1], c[el[cand]])[0]
            got[i] = cand[r] if r >= 0 else -1
    np.testing.assert_array_equal(got, ref)
    assert (ref >= 0).sum() > 500 and (ref < 0).sum() > 100


def test_grid_builder_rejects_bad_input():
    L = _lib.lib()
    c = np.zeros((3, 2))
    el = np.array([[0, 1, 5]], dtype=np.int32)  # node out of range
    lo, inv, ptr = np.zeros(2), np.zeros(2), np.zeros(5, dtype=np.int32)
    rc = L.tatva_host_build_point_grid(c.ctypes.data_as(_lib.c_f64p), 3, el.ctypes.data_as(_lib.c_i32p), 1, 3, 2, 2, lo.ctypes.data_as(_lib.c_f64p), inv.ctypes.data_as(_lib.c_f64p), ptr.ctypes.data_as(_lib.c_i32p), None)
    assert rc != 0
