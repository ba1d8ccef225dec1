"""Host-side (C++) pattern, colouring and CSR position map: bit-exact against reference fixtures."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import tatva_oracle as orc
from tatva_b200 import _lib, sparse
from tatva_b200.mesh import Mesh


def test_library_exports_every_declared_symbol():
    import re, os

    L = _lib.lib()
    hdr = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "tatva_b200.h")).read()
    declared = set(re.findall(r"\b(tatva_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.tatva_abi_version() == _lib.ABI_VERSION


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (what cgo / an XLA-FFI shim / ctypes stubs
    bind) and a C program must link against libtatva_b200.so and call a GPU-free entry point."""
    import os, shutil, subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.abspath(os.path.join(os.path.dirname(_lib.__file__), ".."))
    lib = os.path.join(root, "tatva_b200", "libtatva_b200.so")
    _lib.lib()  # builds the library if it is missing
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "tatva_b200.h"\n'
        "int main(void) {\n"
        "  int32_t conn[6] = {0, 1, 2, 0, 2, 3}, indptr[9], indices[64]; int64_t nnz = 0;\n"
        "  if (tatva_host_pattern_from_mesh(conn, 2, 3, 4, 2, indptr, 0, &nnz) != TATVA_OK) return 2;\n"
        "  if (tatva_host_pattern_from_mesh(conn, 2, 3, 4, 2, indptr, indices, &nnz) != TATVA_OK) return 3;\n"
        '  printf("%d %lld %s\\n", tatva_abi_version(), (long long)nnz, tatva_error_string(TATVA_E_INVALID));\n'
        "  return 0;\n}\n"
    )
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(maxsplit=2)
    assert out[0] == str(_lib.ABI_VERSION) and out[1] == "56"  # two triangles sharing an edge, 2 DOFs per node: (3+4+3+4 node pairs) * 4


@pytest.mark.parametrize("name,dpn", [("tri3_8x8_d2", 2), ("tet4_3_d3", 3), ("tet4_2_d4", 4), ("hex8_3_d3", 3)])
def test_pattern_and_colours_bit_exact_with_reference(golden, name, dpn):
    conn, n_nodes = golden[f"sp_{name}_conn"], int(golden[f"sp_{name}_nnodes"])
    mesh = Mesh(coords=np.zeros((n_nodes, 3)), elements=conn)
    pat = sparse.pattern_from_mesh(mesh, dpn)
    assert pat.indptr.dtype == np.int32 and pat.indices.dtype == np.int32 and pat.data.dtype == np.int8
    np.testing.assert_array_equal(pat.indptr, golden[f"sp_{name}_indptr"])
    np.testing.assert_array_equal(pat.indices, golden[f"sp_{name}_indices"])
    colors = sparse.distance2_colors(pat.indptr, pat.indices, pat.shape[0])
    assert colors.dtype == np.int32
    np.testing.assert_array_equal(colors, golden[f"sp_{name}_colors"])
    cm = sparse.ColoredMatrix.from_csr(pat)
    np.testing.assert_array_equal(cm.colors, golden[f"sp_{name}_colors"])


def test_colouring_is_valid_distance2_on_general_pattern():
    """Ragged pattern with an isolated row: columns sharing a row must get different colours."""
    rng = np.random.default_rng(0)
    A = sps.random(60, 60, density=0.06, random_state=1, format="csr")
    A = ((A + A.T + sps.eye(60)) != 0).astype(np.int8).tolil()
    A[7, :] = 0
    A[:, 7] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    colors = sparse.distance2_colors(A.indptr, A.indices, 60)
    np.testing.assert_array_equal(colors, orc.distance2_colors(A.indptr.astype(np.int32), A.indices.astype(np.int32), 60))
    for i in range(60):
        cols = A.indices[A.indptr[i] : A.indptr[i + 1]]
        assert len(set(colors[cols])) == len(cols)


def test_pattern_with_unreferenced_nodes_and_repeated_elements():
    conn = np.array([[0, 1, 2], [0, 1, 2], [2, 3, 5]], dtype=np.int32)  # node 4 unused, element repeated
    ip, ix = sparse.pattern_arrays(conn, 6, 2)
    rp, rx = orc.pattern_from_mesh(conn, 6, 2)
    np.testing.assert_array_equal(ip, rp)
    np.testing.assert_array_equal(ix, rx)
    assert ip[9] == ip[8] == ip[10]  # rows of node 4 are empty


def test_csr_element_positions():
    c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2))
    dpn = 3
    ip, ix = sparse.pattern_arrays(el, len(c), dpn)
    pos = np.empty((el.shape[0], 4, 4), dtype=np.int32)
    rc = _lib.lib().tatva_host_csr_element_positions(
        el.ctypes.data_as(_lib.c_i32p), el.shape[0], 4, dpn, ip.ctypes.data_as(_lib.c_i32p), ix.ctypes.data_as(_lib.c_i32p), pos.ctypes.data_as(_lib.c_i32p)
    )
    assert rc == 0
    for e in (0, 5, el.shape[0] - 1):
        for a in range(4):
            for b in range(4):
                row = el[e, a] * dpn
                assert ix[ip[row] + pos[e, a, b]] == el[e, b] * dpn
                assert ix[ip[row + 2] + pos[e, a, b] + 2] == el[e, b] * dpn + 2


def test_csr_element_positions_reject_patterns_that_are_not_node_blocked():
    """A same-size pattern whose component rows of one node differ (here: one extra column in the SECOND component row
    of node 0, as a per-component Periodic adds through lifter.augment_sparsity) must be refused: the direct assembly
    kernels write (a, b) blocks at one offset in every component row of node a."""
    import scipy.sparse as sps

    c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2))
    dpn = 3
    ip, ix = sparse.pattern_arrays(el, len(c), dpn)
    n = dpn * len(c)
    A = sps.csr_matrix((np.ones(len(ix), dtype=np.int8), ix, ip), shape=(n, n)).tolil()
    row = 1  # second component row of node 0
    last = int(ix[ip[row + 1] - 1])
    missing = [j for j in range(last) if A[row, j] == 0]  # a column in front of at least one block of the row
    A[row, missing[0]] = 1  # shifts the offsets of the later blocks in this row only
    A = A.tocsr()
    A.sort_indices()
    ip2, ix2 = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    pos = np.empty((el.shape[0], 4, 4), dtype=np.int32)
    args = (el.ctypes.data_as(_lib.c_i32p), el.shape[0], 4, dpn)
    rc = _lib.lib().tatva_host_csr_element_positions(*args, ip2.ctypes.data_as(_lib.c_i32p), ix2.ctypes.data_as(_lib.c_i32p), pos.ctypes.data_as(_lib.c_i32p))
    assert rc == -1  # TATVA_E_INVALID
    assert (pos < 0).any()


def test_invalid_arguments_return_error_codes():
    L = _lib.lib()
    nnz = C.c_int64()
    bad = np.array([[0, 1, 9]], dtype=np.int32)
    ip = np.zeros(7, dtype=np.int32)
    assert L.tatva_host_pattern_from_mesh(bad.ctypes.data_as(_lib.c_i32p), 1, 3, 3, 2, ip.ctypes.data_as(_lib.c_i32p), None, C.byref(nnz)) == -1
    assert L.tatva_error_string(-1) == b"invalid argument"
    assert L.tatva_plan_destroy(None) == 0


@pytest.mark.parametrize("case", ["tri3_d1", "tri3_d3", "tet4_d4", "hex8_d3", "shuffled_isolated_d2", "shuffled_isolated_d3", "random"])
def test_colouring_node_level_sweep_is_bit_exact_with_the_per_dof_greedy(case):
    """The C++ colouring detects the block structure of mesh patterns (dpn identical rows per node) and sweeps node
    by node; the colours must equal the per-DOF first-fit of the in-tree reference algorithm
    (tatva/sparse/_coloring.py:27-48, :136-153, :270-283) restated in the oracle — also with nodes that belong to no
    element, shuffled node numbering, and patterns without block structure (generic path)."""
    import scipy.sparse as sps
    from oracle import tatva_oracle as orc

    rng = np.random.default_rng(0)
    if case == "random":
        A = sps.random(80, 80, density=0.06, random_state=1, format="csr")
        A = (A + A.T + sps.eye(80)).tocsr()
        A.sort_indices()
        ip, ix, n = A.indptr, A.indices, 80
    else:
        dpn = int(case[-1])
        if case.startswith("tri3"):
            c, el = orc.mesh_unit_square_tri(7, 5)
            nn = len(c)
        elif case.startswith("tet4"):
            c, el = orc.mesh_box_tet((1, 1, 1), (3, 3, 2))
            nn = len(c)
        elif case.startswith("hex8"):
            c, el = orc.mesh_box_hex(3)
            nn = len(c)
        else:
            c, el = orc.mesh_unit_square_tri(6, 6)
            nn = len(c) + 5
            el = rng.permutation(nn)[el]
        ip, ix = orc.pattern_from_mesh(el, nn, dpn)
        n = nn * dpn
    ref = orc.distance2_colors(ip, ix, n)
    got = sparse.distance2_colors(np.asarray(ip, dtype=np.int32), np.asarray(ix, dtype=np.int32), n)
    np.testing.assert_array_equal(got, ref)


def test_csr_tile_schedule_lists_every_upper_contribution_once():
    """Combine schedule of the tiled assembly (tatva_host_csr_tile_schedule): every (element, a, b) with row node <=
    column node appears exactly once, under the block that holds its CSR position and that of its mirror; blocks of a
    tile are ordered by decreasing contributor count."""
    c, el = orc.mesh_box_tet((1, 1, 1), (5, 4, 3))
    dpn = 3
    ip, ix = sparse.pattern_arrays(el, len(c), dpn)
    L = _lib.lib()
    i32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
    E, npe = el.shape
    pos = np.empty((E, npe, npe), dtype=np.int32)
    assert L.tatva_host_csr_element_positions(i32(el), E, npe, dpn, i32(ip), i32(ix), i32(pos)) == 0
    tile = 128
    nt = (E + tile - 1) // tile
    bp = np.empty(nt + 1, dtype=np.int32)
    nb, nc = C.c_int64(), C.c_int64()
    args = (i32(el), E, npe, dpn, tile, i32(ip), i32(pos), i32(bp), C.byref(nb), C.byref(nc))
    assert L.tatva_host_csr_tile_schedule(*args, None, None, None, None, None, None) == 0
    n = nb.value
    base, rl, base_t, rl_t = (np.empty(n, dtype=np.int32) for _ in range(4))
    cp = np.empty(n + 1, dtype=np.int32)
    con = np.empty(nc.value, dtype=np.uint32)
    assert L.tatva_host_csr_tile_schedule(*args, i32(base), i32(rl), i32(base_t), i32(rl_t), i32(cp), con.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    assert cp[0] == 0 and cp[-1] == nc.value and (np.diff(cp) > 0).all()
    upper = el[:, :, None] <= el[:, None, :]
    assert nc.value == int(upper.sum())
    seen = np.zeros((E, npe, npe), dtype=int)
    for t in range(nt):
        counts = np.diff(cp[bp[t] : bp[t + 1] + 1])
        assert (np.diff(counts) <= 0).all()
        for d in range(bp[t], bp[t + 1]):
            for k in range(cp[d], cp[d + 1]):
                s = int(con[k])
                e, a, b = t * tile + (s >> 8), (s >> 4) & 15, s & 15
                seen[e, a, b] += 1
                row, rowt = el[e, a] * dpn, el[e, b] * dpn
                assert base[d] == ip[row] + pos[e, a, b] and rl[d] == ip[row + 1] - ip[row]
                if el[e, a] == el[e, b]:
                    assert base_t[d] == -1
                else:
                    assert base_t[d] == ip[rowt] + pos[e, b, a] and rl_t[d] == ip[rowt + 1] - ip[rowt]
    assert np.array_equal(seen, upper.astype(int))


def test_halo_exchange_abi_rejects_bad_arguments_without_a_gpu():
    """tatva_halo_exchange / the communicator helpers (tatva/mpi.py:400-407, :505-513 as one C call over an ncclComm_t):
    argument checks that need neither a GPU nor a communicator; NCCL itself is opened lazily with dlopen."""
    import ctypes as C

    L = _lib.lib()
    ver = C.c_int(0)
    rc = L.tatva_nccl_version(C.byref(ver))
    assert rc in (0, _lib_unsupported()) and (rc != 0 or ver.value >= 20700)
    assert L.tatva_nccl_version(None) != 0
    cnt = (C.c_int64 * 2)(0, 0)
    assert L.tatva_halo_exchange(None, None, None, cnt, None, None, cnt, None, None, 0, None) != 0  # no communicator
    assert L.tatva_halo_comm_unique_id(None) != 0
    h = C.c_void_p()
    buf = (C.c_char * 128)()
    assert L.tatva_halo_comm_create(C.byref(h), buf, 0, 0) != 0  # n_ranks must be positive
    assert L.tatva_halo_comm_create(C.byref(h), buf, 2, 5) != 0  # rank out of range
    assert L.tatva_halo_comm_destroy(None) == 0
    assert L.tatva_zero_release(None, 8, None) != 0
    assert L.tatva_plan_cache_geometry(None, 1, None) != 0
    assert L.tatva_plan_set_node_schedule(None, None, None, None, None, None, None) != 0


def _lib_unsupported():
    import re

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "tatva_b200.h")).read()
    return int(re.search(r"TATVA_E_UNSUPPORTED\s*=\s*(-?\d+)", src).group(1))
