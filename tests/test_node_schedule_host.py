"""Host side of the warp-cooperative kernels (tatva_host_node_schedule): the gather lists reproduce the connectivity and
the per-tile contributor tables reproduce the scatter `np.add.at(y, elements, Y)` (the transpose of the reference's
gather, tatva/operator.py:221), on a structured tet box, on a random connectivity (overflow rows) and on ragged sizes."""
import ctypes as C

import numpy as np
import pytest

from oracle import tatva_oracle as orc
from tatva_b200 import _lib


def _schedule(conn, cap=0):
    L = _lib.lib()
    E, npe = conn.shape
    nt = (E + 127) // 128
    i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
    wn, wl, cp = np.empty(nt * 128, np.int32), np.empty(E * npe, np.uint8), np.empty(nt + 1, np.int32)
    nch, nell = C.c_int64(), C.c_int64()
    assert L.tatva_host_node_schedule(i32(conn), E, npe, cap, i32(wn), wl.ctypes.data_as(C.POINTER(C.c_uint8)), i32(cp), C.byref(nch), C.byref(nell), None, None, None) == 0
    tn, ep, ell = np.empty(32 * nch.value, np.int32), np.empty(nch.value + 1, np.int32), np.empty(nell.value, np.uint16)
    assert L.tatva_host_node_schedule(i32(conn), E, npe, cap, None, None, i32(cp), C.byref(nch), C.byref(nell), i32(tn), i32(ep), ell.ctypes.data_as(C.POINTER(C.c_uint16))) == 0
    return wn, wl.reshape(E, npe), cp, tn, ep, ell


def _check(conn, n_nodes, cap=0):
    E, npe = conn.shape
    wn, wl, cp, tn, ep, ell = _schedule(conn, cap)
    # gather: every row an element picks by lane is its own node; lists ascending; overflow only when the warp is full
    for w in range((E + 31) // 32):
        lst = wn[32 * w : 32 * w + 32]
        used = lst[lst >= 0]
        assert np.all(np.diff(used) > 0)
        sl = slice(32 * w, min(E, 32 * w + 32))
        loc, nodes = wl[sl].astype(int), conn[sl]
        picked = loc != 255
        assert np.array_equal(lst[loc[picked]], nodes[picked])
        if (~picked).any():
            assert len(used) == 32
    assert np.all(wn[32 * ((E + 31) // 32) :] == -1)
    # scatter: the contributor tables list every (element, local node) exactly once, under its own node
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(E, npe))
    ref = np.zeros(n_nodes)
    np.add.at(ref, conn, Y)
    out = np.zeros(n_nodes)
    seen = np.zeros((E, npe), dtype=int)
    assert ep[0] == 0 and ep[-1] == len(ell) and np.all(np.diff(ep) % 32 == 0)
    for t in range(len(cp) - 1):
        prev = None
        for ch in range(cp[t], cp[t + 1]):
            rows = (ep[ch + 1] - ep[ch]) // 32
            tab = ell[ep[ch] : ep[ch + 1]].reshape(rows, 32).astype(int)
            counts = (tab != 0xFFFF).sum(0)
            assert counts.max() == rows and np.all(np.diff(counts) <= 0)  # decreasing contributor count
            assert cap == 0 or rows <= cap
            assert prev is None or counts[0] <= prev
            prev = counts[-1]
            for lane in range(32):
                node = tn[32 * ch + lane]
                assert (node >= 0) == (counts[lane] > 0)
                for s in tab[: counts[lane], lane]:
                    e, a = 128 * t + (s >> 3), s & 7
                    assert conn[e, a] == node
                    seen[e, a] += 1
                    out[node] += Y[e, a]
    assert np.all(seen == 1)
    np.testing.assert_allclose(out, ref, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("n", [2, 5])
@pytest.mark.parametrize("cap", [0, 3, 6])
def test_structured_tet_box(n, cap):
    c, el = orc.mesh_box_tet((1.0, 1.0, 1.0), (n, n, n))
    _check(np.ascontiguousarray(el, dtype=np.int32), len(c), cap)


@pytest.mark.parametrize("E,npe,n_nodes", [(1, 3, 3), (129, 4, 40), (500, 4, 2000), (300, 8, 90)])
def test_random_connectivity_takes_the_overflow_path(E, npe, n_nodes):
    rng = np.random.default_rng(E)
    conn = np.stack([rng.choice(n_nodes, size=npe, replace=False) for _ in range(E)]).astype(np.int32)
    _check(conn, n_nodes)
    _check(conn, n_nodes, cap=2)


def test_invalid_arguments():
    L = _lib.lib()
    conn = np.zeros((4, 9), np.int32)
    n = C.c_int64()
    cp = np.zeros(2, np.int32)
    assert L.tatva_host_node_schedule(conn.ctypes.data_as(_lib.c_i32p), 4, 9, 0, None, None, cp.ctypes.data_as(_lib.c_i32p), C.byref(n), C.byref(n), None, None, None) != 0


def test_random_meshes_property():
    """Property test (hypothesis): for any connectivity — repeated nodes inside an element included — the schedule lists
    every (element, local node) reference exactly once under its node, the warp lists are consistent, and a cap bounds
    every table row count."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 400), st.integers(1, 8), st.integers(1, 60), st.integers(0, 5), st.integers(0, 2**31 - 1))
    def run(E, npe, n_nodes, cap, seed):
        rng = np.random.default_rng(seed)
        conn = rng.integers(0, n_nodes, size=(E, npe)).astype(np.int32)
        wn, wl, cp, tn, ep, ell = _schedule(conn, cap)
        for w in range((E + 31) // 32):
            lst = wn[32 * w : 32 * w + 32]
            sl = slice(32 * w, min(E, 32 * w + 32))
            loc, nodes = wl[sl].astype(int), conn[sl]
            picked = loc != 255
            assert np.array_equal(lst[loc[picked]], nodes[picked])
        seen = np.zeros((E, npe), dtype=int)
        for t in range(len(cp) - 1):
            for ch in range(cp[t], cp[t + 1]):
                rows = (ep[ch + 1] - ep[ch]) // 32
                assert cap == 0 or rows <= cap
                tab = ell[ep[ch] : ep[ch + 1]].reshape(rows, 32).astype(int)
                for lane in range(32):
                    for s_ in tab[:, lane]:
                        if s_ == 0xFFFF:
                            continue
                        e, a = 128 * t + (s_ >> 3), s_ & 7
                        assert conn[e, a] == tn[32 * ch + lane]
                        seen[e, a] += 1
        assert np.all(seen == 1)

    run()
