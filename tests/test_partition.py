"""Partitioning host logic: structured block builder == extract_local_mesh on the global mesh."""
import numpy as np
import pytest

from oracle import tatva_oracle as orc
from tatva_b200.distributed import structured_hex_block
from tatva_b200.mesh import Mesh, block_partition, extract_local_mesh


@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_structured_block_matches_extract_local_mesh(grid):
    n = 3
    nparts = grid[0] * grid[1] * grid[2]
    shape = (grid[0] * n, grid[1] * n, grid[2] * n)
    gm = Mesh.box_hex(shape)
    part = block_partition(shape, nparts)
    for r in range(nparts):
        ref_mesh, ref_info = extract_local_mesh(gm, part, r)
        mesh, info = structured_hex_block(n, grid, r, jitter=0.0)
        np.testing.assert_array_equal(info.nodes_local_to_global, ref_info.nodes_local_to_global)
        assert info.n_owned_nodes == ref_info.n_owned_nodes
        np.testing.assert_array_equal(mesh.elements, ref_mesh.elements)
        np.testing.assert_allclose(mesh.coords, ref_mesh.coords, atol=1e-15)


@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_structured_block_non_cubic_matches_extract_local_mesh(grid):
    """Strong scaling cuts a FIXED cube into non-cubic blocks (256^3 -> 128 x 256 x 256 at 2 GPUs)."""
    N = 4
    nparts = grid[0] * grid[1] * grid[2]
    per = (N // grid[0], N // grid[1], N // grid[2])
    gm = Mesh.box_hex((N, N, N))
    part = block_partition((N, N, N), nparts)
    for r in range(nparts):
        ref_mesh, ref_info = extract_local_mesh(gm, part, r)
        mesh, info = structured_hex_block(per, grid, r, jitter=0.0)
        np.testing.assert_array_equal(info.nodes_local_to_global, ref_info.nodes_local_to_global)
        assert info.n_owned_nodes == ref_info.n_owned_nodes
        np.testing.assert_array_equal(mesh.elements, ref_mesh.elements)
        np.testing.assert_allclose(mesh.coords, ref_mesh.coords, atol=1e-15)


def test_block_jitter_is_consistent_across_ranks():
    a, ia = structured_hex_block(3, (2, 1, 1), 0)
    b, ib = structured_hex_block(3, (2, 1, 1), 1)
    shared, ai, bi = np.intersect1d(ia.nodes_local_to_global, ib.nodes_local_to_global, return_indices=True)
    assert shared.size == 16
    np.testing.assert_array_equal(a.coords[ai], b.coords[bi])
    assert np.abs(a.coords[ai] - Mesh.box_hex((6, 3, 3)).coords[shared]).max() > 0


def test_extract_local_mesh_matches_oracle_restatement():
    c, el = orc.mesh_box_tet((1, 1, 1), (3, 2, 2))
    part = (np.arange(el.shape[0]) % 3).astype(np.int32)
    for r in range(3):
        m, info = extract_local_mesh(Mesh(coords=c, elements=el), part, r)
        cl, el_l, l2g, n_owned = orc.extract_local_mesh(c, el, part, r)
        np.testing.assert_array_equal(m.elements, el_l)
        np.testing.assert_array_equal(info.nodes_local_to_global, l2g)
        assert info.n_owned_nodes == n_owned


def test_locality_reordering_is_a_consistent_permutation():
    from tatva_b200.mesh import locality_order, reorder_mesh

    rng = np.random.default_rng(0)
    c, el = orc.mesh_box_tet((1, 1, 1), (8, 8, 8))
    shuffle_e, shuffle_n = rng.permutation(el.shape[0]), rng.permutation(c.shape[0])
    inv = np.empty_like(shuffle_n)
    inv[shuffle_n] = np.arange(len(shuffle_n))
    bad = Mesh(coords=c[shuffle_n], elements=inv[el[shuffle_e]].astype(np.int32))
    new, ep, npm = reorder_mesh(bad)
    assert sorted(ep.tolist()) == list(range(el.shape[0])) and sorted(npm.tolist()) == list(range(c.shape[0]))
    np.testing.assert_array_equal(new.coords[new.elements], bad.coords[bad.elements[ep]])  # same geometry per element
    # locality: consecutive elements are close in space, and node ids within an element are close
    cen = new.coords[new.elements].mean(axis=1)
    cen_bad = bad.coords[bad.elements].mean(axis=1)
    assert np.linalg.norm(np.diff(cen, axis=0), axis=1).mean() < 0.35 * np.linalg.norm(np.diff(cen_bad, axis=0), axis=1).mean()
    span = lambda e: (e.max(axis=1) - e.min(axis=1)).mean()  # noqa: E731
    assert span(new.elements) < 0.35 * span(bad.elements)
    # the oracle residual is invariant under the permutation
    mat = orc.NeoHookean(500.0, 1000.0)
    u = 0.01 * rng.normal(size=c.shape)
    r_bad = orc.residual("tet4", mat, bad.coords, bad.elements, u)
    r_new = orc.residual("tet4", mat, new.coords, new.elements, u[npm])
    np.testing.assert_allclose(r_new, r_bad[npm], rtol=1e-12, atol=1e-12)
    assert sorted(locality_order(c, el).tolist()) == list(range(el.shape[0]))


def test_structured_tet_block_matches_extract_local_mesh():
    from tatva_b200.distributed import structured_tet_block

    n, grid = 2, (2, 2, 1)
    shape = (grid[0] * n, grid[1] * n, grid[2] * n)
    m = max(shape)
    gm = Mesh.box_tet((shape[0] / m, shape[1] / m, shape[2] / m), shape)
    part = np.repeat(block_partition(shape, 4), 6)
    for r in range(4):
        ref_mesh, ref_info = extract_local_mesh(gm, part, r)
        mesh, info = structured_tet_block(n, grid, r, jitter=0.0)
        np.testing.assert_array_equal(info.nodes_local_to_global, ref_info.nodes_local_to_global)
        assert info.n_owned_nodes == ref_info.n_owned_nodes
        np.testing.assert_array_equal(mesh.elements, ref_mesh.elements)
        # box_tet is centred in x, y (tests/test_sparse_tracer.py:35-37); the block builder starts at the origin
        np.testing.assert_allclose(mesh.coords, ref_mesh.coords + np.array([shape[0] / m / 2, shape[1] / m / 2, 0.0]), atol=1e-15)


@pytest.mark.parametrize("name", ["tri2d", "tri3d", "quad", "tet", "hex"])
def test_mesh_hmin_hmax_match_reference(golden, name):
    """Mesh.hmin / hmax / _element_circumdiameters against the reference's outputs (tatva/mesh.py:87-144): both
    simplex formulas and the max-vertex-distance fallback."""
    from tatva_b200.mesh import Mesh

    m = Mesh(coords=golden[f"h_{name}_coords"], elements=golden[f"h_{name}_conn"])
    np.testing.assert_allclose(m._element_circumdiameters(), golden[f"h_{name}_diam"], rtol=1e-13)
    np.testing.assert_allclose(m.hmin(), golden[f"h_{name}_hmin"], rtol=1e-13)
    np.testing.assert_allclose(m.hmax(), golden[f"h_{name}_hmax"], rtol=1e-13)
