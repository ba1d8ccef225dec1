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
