"""Compound layouts — known answers of the reference's tests/test_compound.py."""
import numpy as np
import pytest
import torch

from tatva_b200.compound import Compound, FieldSize, field, stack_fields
from tatva_b200.mesh import Mesh


class SimpleState(Compound):
    u = field((2, 3))
    phi = field((2,))


@stack_fields("u", "v", axis=-1)
class StackedState(Compound):
    u = field((2, 2))
    v = field((2, 2))
    w = field((2,))


@pytest.mark.parametrize("state_cls", [SimpleState, StackedState])
def test_compound_size_matches_flat_array(state_cls):
    state = state_cls()
    assert state.arr.shape == (state_cls.size,)
    assert np.all(state.arr == 0)


def test_compound_field_access_and_assignment():
    """reference tests/test_compound.py:32-54."""
    state = SimpleState()
    assert SimpleState.size == 8
    u_val = np.arange(6.0).reshape(2, 3)
    phi_val = np.array([10.0, 20.0])
    state = state.at("u").set(u_val).at("phi").set(phi_val)
    np.testing.assert_array_equal(state.u, u_val)
    np.testing.assert_array_equal(state.phi, phi_val)
    np.testing.assert_array_equal(state.arr[:6], u_val.ravel())
    np.testing.assert_array_equal(state.arr[6:], phi_val)
    assert len(state) == 2
    assert [c.shape for c in state] == [(2, 3), (2,)]


def test_compound_index_helpers():
    """reference tests/test_compound.py:57-60."""
    np.testing.assert_array_equal(SimpleState.u[1], [3, 4, 5])
    np.testing.assert_array_equal(SimpleState.u[:, 1], [1, 4])
    np.testing.assert_array_equal(SimpleState.phi[1], [7])


def test_addition_and_unpacking():
    a = SimpleState(np.arange(8.0))
    b = SimpleState(np.arange(8.0) * 2)
    s = a + b
    np.testing.assert_array_equal(s.arr, np.arange(8.0) * 3)
    u, phi = s
    np.testing.assert_array_equal(u, a.u + b.u)
    np.testing.assert_array_equal(phi, a.phi + b.phi)


def test_stack_fields_access_and_indices():
    """reference tests/test_compound.py:83-95."""
    state = StackedState(np.arange(StackedState.size, dtype=np.float64))
    np.testing.assert_array_equal(state.u, [[0.0, 1.0], [4.0, 5.0]])
    np.testing.assert_array_equal(state.v, [[2.0, 3.0], [6.0, 7.0]])
    np.testing.assert_array_equal(state.w, [8.0, 9.0])
    np.testing.assert_array_equal(StackedState.u[1], [4, 5])
    np.testing.assert_array_equal(StackedState.v[0], [2, 3])


def test_auto_sizing_nodal_fields():
    """reference tests/test_compound.py:98-155: node-interleaved stacked block first, then the rest."""
    mesh = Mesh(coords=np.zeros((10, 2)), elements=None)

    class MyState(Compound, mesh=mesh):
        param1 = field(shape=(5,))
        u = field(shape=(FieldSize.AUTO, 3))
        phi = field(shape=(FieldSize.AUTO,))
        param2 = field(shape=(2,))

    state = MyState()
    assert [n for n, _ in MyState.fields] == ["param1", "u", "phi", "param2"]
    p1, u, phi, p2 = state
    assert (p1.shape, u.shape, phi.shape, p2.shape) == ((5,), (10, 3), (10,), (2,))
    assert MyState.size == 47 and state.arr.size == 47
    np.testing.assert_array_equal(MyState.u.indices(slice(None)), [i * 4 + j for i in range(10) for j in range(3)])
    np.testing.assert_array_equal(MyState.phi.indices(slice(None)), [i * 4 + 3 for i in range(10)])
    np.testing.assert_array_equal(MyState.param1.indices(slice(None)), np.arange(40, 45))
    np.testing.assert_array_equal(MyState.param2.indices(slice(None)), np.arange(45, 47))


def test_single_auto_field_is_not_stacked():
    """compound/__init__.py:184-188."""
    mesh = Mesh(coords=np.zeros((4, 3)), elements=None)

    class S(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 3))

    np.testing.assert_array_equal(S.u.indices(slice(None)), np.arange(12))
    assert S.u._slice == slice(0, 12)


def test_torch_views_are_zero_copy():
    mesh = Mesh(coords=np.zeros((5, 3)), elements=None)

    class S(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 3))
        phi = field(shape=(FieldSize.AUTO,))

    arr = torch.arange(20, dtype=torch.float64)
    s = S(arr)
    assert s.u.shape == (5, 3) and s.phi.shape == (5,)
    assert s.u.data_ptr() == arr.data_ptr() and s.phi.data_ptr() == arr.data_ptr() + 3 * 8
    np.testing.assert_array_equal(s.phi.numpy(), [3, 7, 11, 15, 19])
    s2 = s.at("phi").set(torch.zeros(5, dtype=torch.float64))
    assert float(s2.phi.abs().sum()) == 0 and float(s.phi.sum()) == 55
    np.testing.assert_array_equal(s2.u.numpy(), s.u.numpy())


def test_pattern_from_compound_matches_mesh_pattern_for_stacked_nodal():
    from oracle import tatva_oracle as orc
    from tatva_b200 import sparse

    c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2))
    mesh = Mesh(coords=c, elements=el)

    class S(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 3))
        phi = field(shape=(FieldSize.AUTO,))

    a = sparse.pattern_from_compound(S)
    b = sparse.pattern_from_mesh(mesh, 4)
    np.testing.assert_array_equal(a.indptr, b.indptr)
    np.testing.assert_array_equal(a.indices, b.indices)
    blocks = sparse.pattern_from_compound(S, block_wise=True)
    assert len(blocks) == 1 and blocks[0][0].shape == (S.size, S.size)


def test_pattern_from_compound_mixed_layout_equals_the_pair_list_construction():
    """A full nodal field, a nodal field on a node subset and a shared field: the C++ pattern builder must give
    exactly the unique (row, col) pairs the reference's construction produces (sparse/_extraction.py:118-245:
    nodal fields coupled within elements, every other field diagonal)."""
    import warnings

    import scipy.sparse as sp

    from oracle import tatva_oracle as orc
    from tatva_b200 import sparse
    from tatva_b200.compound import Compound, FieldSize, FieldType, Nodal, field
    from tatva_b200.mesh import Mesh

    c, el = orc.mesh_unit_square_tri(5, 4)
    mesh = Mesh(coords=c, elements=el.astype(np.int32))
    sub = np.array([0, 3, 7, 8, 20])

    class Mixed(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 2))
        lam = field(shape=(FieldSize.AUTO, 1), field_type=Nodal(node_ids=sub))
        g = field(shape=(3,), field_type=FieldType.SHARED)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "Custom space detected ..." for the shared field
        pat = sparse.pattern_from_compound(Mixed)
    fields = dict(Mixed.fields)
    nd_u = np.asarray(fields["u"].indices(slice(None))).reshape(len(c), 2)
    nd_l = np.full((len(c), 1), -1)
    nd_l[sub] = np.asarray(fields["lam"].indices(slice(None))).reshape(-1, 1)
    ed = np.concatenate([nd_u[el].reshape(len(el), -1), nd_l[el].reshape(len(el), -1)], axis=1)
    pairs = set()
    for row in ed:
        v = row[row >= 0]
        pairs.update((int(a), int(b)) for a in v for b in v)
    pairs.update((int(d), int(d)) for d in np.asarray(fields["g"].indices(slice(None))))
    rows, cols = zip(*sorted(pairs))
    ref = sp.csr_matrix((np.ones(len(pairs), dtype=np.int8), (rows, cols)), shape=(Mixed.size, Mixed.size))
    ref.sort_indices()
    assert pat.dtype == np.int8 and pat.indptr.dtype == np.int32
    np.testing.assert_array_equal(pat.indptr, ref.indptr)
    np.testing.assert_array_equal(pat.indices, ref.indices)


def test_pattern_from_compound_matches_the_reference_flat_and_block_wise(golden):
    """sparse.pattern_from_compound of the unmodified reference (sparse/_extraction.py:118-245) for a full nodal field, a
    nodal field on a node subset and a shared field: same CSR, and the same block-wise decomposition."""
    import warnings

    from oracle import tatva_oracle as orc
    from tatva_b200 import sparse
    from tatva_b200.compound import Compound, FieldSize, FieldType, Nodal, field
    from tatva_b200.mesh import Mesh

    c, el = orc.mesh_unit_square_tri(5, 4)
    mesh = Mesh(coords=c, elements=el.astype(np.int32))

    class Mixed(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 2))
        lam = field(shape=(FieldSize.AUTO, 1), field_type=Nodal(node_ids=golden["cpat_subset"]))
        g = field(shape=(3,), field_type=FieldType.SHARED)

    assert Mixed.size == int(golden["cpat_size"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pat = sparse.pattern_from_compound(Mixed)
        blocks = sparse.pattern_from_compound(Mixed, block_wise=True)
    np.testing.assert_array_equal(pat.indptr, golden["cpat_indptr"])
    np.testing.assert_array_equal(pat.indices, golden["cpat_indices"])
    ni, nj = (int(x) for x in golden["cpat_block_grid"])
    assert len(blocks) == ni and all(len(row) == nj for row in blocks)
    for i in range(ni):
        for j in range(nj):
            b = blocks[i][j].tocsr()
            b.sort_indices()
            assert tuple(b.shape) == tuple(golden[f"cpat_block_{i}{j}_shape"])
            np.testing.assert_array_equal(b.indptr, golden[f"cpat_block_{i}{j}_indptr"])
            np.testing.assert_array_equal(b.indices, golden[f"cpat_block_{i}{j}_indices"])
