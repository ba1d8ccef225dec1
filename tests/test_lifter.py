"""Lifter — known answers of the reference's tests/test_lifter.py and tests/test_lifter_residual.py (NumPy
path), plus the CUDA kernels against the NumPy path (gpu)."""
import numpy as np
import pytest
import scipy.sparse as sps

from tatva_b200.lifter import Fixed, Lifter, LifterError, Periodic, RuntimeValue, lifted


def test_lifter_without_constraints_roundtrips():
    lifter = Lifter(4)
    u = np.arange(4, dtype=np.float64)
    full = lifter.lift_from_zeros(u)
    np.testing.assert_array_equal(full, u)
    np.testing.assert_array_equal(lifter.reduce(full), u)
    assert lifter.constrained_dofs.size == 0


def _example():
    return Lifter(6, Fixed(np.array([0, 5], dtype=np.int32)), Periodic(dofs=np.array([2], dtype=np.int32), master_dofs=np.array([1], dtype=np.int32)))


def test_lifter_applies_dirichlet_and_periodic_constraints():
    """reference tests/test_lifter.py:27-43."""
    lifter = _example()
    u = np.array([10.0, 20.0, 30.0])
    full = lifter.lift_from_zeros(u)
    np.testing.assert_array_equal(full, [0.0, 10.0, 10.0, 20.0, 30.0, 0.0])
    np.testing.assert_array_equal(lifter.reduce(full), u)
    assert lifter.size_reduced == 3
    np.testing.assert_array_equal(lifter.free_dofs, [1, 3, 4])
    np.testing.assert_array_equal(lifter.constrained_dofs, [0, 2, 5])


def test_constraints_and_lifter_are_hashable():
    lifter = _example()
    assert all(isinstance(hash(x), int) for x in (lifter, *lifter.constraints))


def test_runtime_values():
    """reference tests/test_lifter.py:80-91."""
    lifter = Lifter(4, Fixed(np.array([0, 3], dtype=np.int32), RuntimeValue("top")))
    with pytest.raises(LifterError):
        lifter.lift_from_zeros(np.array([1.0, 2.0]))
    lhs = lifter.with_values({"top": np.array([1.0, 2.0])})
    rhs = lifter.at["top"].set(np.array([1.0, 2.0]))
    diff = lifter.with_values({"top": np.array([1.0, 3.0])})
    assert lhs == rhs and lhs != diff
    np.testing.assert_array_equal(lhs.lift_from_zeros(np.array([5.0, 6.0])), [1.0, 5.0, 6.0, 2.0])
    with pytest.raises(LifterError):
        lifter.with_values({"bottom": 1.0})


def test_lifted_decorator():
    """reference tests/test_lifter.py:94-117."""
    lifter = Lifter(4, Fixed(np.array([0, 3]), 0.0))
    assert lifted(lambda u: u.sum(), argnums=0)(lifter, np.array([10.0, 20.0])) == 30.0
    np.testing.assert_array_equal(lifted(lambda u: u * 2, argnums=0, output="primal")(lifter, np.array([10.0, 20.0])), [20.0, 40.0])
    with pytest.raises(LifterError):
        lifted(lambda u: u)(lifter, np.zeros(3))


def test_reduce_adjoint_known_answers():
    """reference tests/test_lifter_residual.py."""
    r = np.array([10.0, 20.0, 30.0, 40.0])
    per = Lifter(4, Periodic(dofs=np.array([2]), master_dofs=np.array([1])))
    np.testing.assert_array_equal(per.reduce_adjoint(r), [10.0, 50.0, 40.0])
    np.testing.assert_array_equal(Lifter(4, Fixed(np.array([0, 3]), 0.0)).reduce_adjoint(r), [20.0, 30.0])
    both = Lifter(4, Periodic(dofs=np.array([2]), master_dofs=np.array([1])), Fixed(np.array([3]), 0.0))
    np.testing.assert_array_equal(both.reduce_adjoint(r), [10.0, 50.0])
    out = lifted(lambda u: u * 2.0, argnums=0, output="dual")(per, np.array([1.0, 2.0, 3.0]))
    np.testing.assert_array_equal(out, [2.0, 8.0, 6.0])


def test_reduce_adjoint_is_the_adjoint_of_lift_and_chains_compose():
    rng = np.random.default_rng(0)
    n = 40
    lifter = Lifter(n, Fixed(np.array([0, 1, 39]), 2.5), Periodic(dofs=np.array([10, 11]), master_dofs=np.array([5, 6])), Periodic(dofs=np.array([20]), master_dofs=np.array([10])))
    u, r = rng.normal(size=lifter.size_reduced), rng.normal(size=n)
    lin = lifter.lift_from_zeros(u) - lifter.lift_from_zeros(np.zeros_like(u))
    assert abs(lin @ r - u @ lifter.reduce_adjoint(r)) < 1e-12
    full = lifter.lift_from_zeros(u)
    assert full[20] == full[10] == full[5] and full[0] == 2.5


def test_sparsity_adaptation():
    S = sps.csr_matrix(np.array([[1, 1, 0, 0], [1, 1, 1, 0], [0, 1, 1, 1], [0, 0, 1, 1]], dtype=np.int8))
    lifter = Lifter(4, Periodic(dofs=np.array([3]), master_dofs=np.array([0])))
    R = lifter.adapt_sparsity(S)
    assert R.shape == (3, 3)
    assert R[0, 2] != 0 and R[2, 0] != 0  # master 0 inherits the coupling of its slave 3 with dof 2


@pytest.mark.gpu
def test_cuda_kernels_match_numpy_path():
    import torch

    rng = np.random.default_rng(1)
    n = 10000
    slaves = np.arange(100, 200)
    lifter = Lifter(n, Fixed(np.arange(0, 50), rng.normal(size=50)), Periodic(dofs=slaves, master_dofs=slaves + 1000), Fixed(np.array([n - 1]), RuntimeValue("load", 3.0)))
    u, r, base = rng.normal(size=lifter.size_reduced), rng.normal(size=n), rng.normal(size=n)
    ut, rt, bt = (torch.as_tensor(a, device="cuda") for a in (u, r, base))
    np.testing.assert_array_equal(lifter.lift_from_zeros(ut).cpu().numpy(), lifter.lift_from_zeros(u))
    np.testing.assert_array_equal(lifter.lift(ut, bt).cpu().numpy(), lifter.lift(u, base))
    np.testing.assert_array_equal(lifter.reduce(rt).cpu().numpy(), lifter.reduce(r))
    np.testing.assert_allclose(lifter.reduce_adjoint(rt).cpu().numpy(), lifter.reduce_adjoint(r), rtol=1e-15, atol=1e-15)
    l2 = lifter.at["load"].set(7.0)
    assert float(l2.lift_from_zeros(ut)[-1]) == 7.0 and float(lifter.lift_from_zeros(ut)[-1]) == 3.0


@pytest.mark.gpu
def test_reduced_hvp_through_lifter_matches_oracle():
    """Dirichlet-constrained Hex8 neo-Hookean operator: reduce_adjoint(H(lift(u)) lift_lin(v)) vs the oracle."""
    import torch

    import tatva_b200
    from oracle import tatva_oracle as orc
    from tatva_b200 import element, materials

    rng = np.random.default_rng(2)
    c, el = orc.mesh_box_hex(6)
    c = c + 0.01 * rng.uniform(-1, 1, c.shape)
    fixed_nodes = np.where(c[:, 0] < 0.02)[0]
    load_nodes = np.where(c[:, 0] > 0.98)[0]
    lifter = Lifter(c.size, Fixed((fixed_nodes[:, None] * 3 + np.arange(3)).ravel()), Fixed(load_nodes * 3 + 2, 0.05))
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8())
    mat, omat = materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0)
    u_red, v_red = 0.01 * rng.normal(size=lifter.size_reduced), rng.normal(size=lifter.size_reduced)
    zero_bc = Lifter(c.size, Fixed(lifter.constrained_dofs))  # tangent directions vanish on constrained DOFs
    u_full = lifter.lift_from_zeros(torch.as_tensor(u_red, device="cuda"))
    v_full = zero_bc.lift_from_zeros(torch.as_tensor(v_red, device="cuda"))
    Hv = lifter.reduce_adjoint(op.hvp(mat)(u_full.view(-1, 3), v_full.view(-1, 3)).reshape(-1))
    ref = lifter.reduce_adjoint(orc.hvp("hex8", omat, c, el, lifter.lift_from_zeros(u_red).reshape(-1, 3), zero_bc.lift_from_zeros(v_red).reshape(-1, 3)).ravel())
    assert np.linalg.norm(Hv.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-12


def test_dof_map_is_the_homogeneous_lift_and_its_transpose():
    """`Lifter.dof_map` (consumed by tatva_hvp_lifted inside the HVP kernel): full DOF -> driving reduced DOF or -1,
    equivalent to homogeneous().lift_from_zeros and reduce_adjoint (reference lifter/base.py:201-251)."""
    import numpy as np
    from tatva_b200.lifter import Fixed, Lifter, Periodic

    rng = np.random.default_rng(0)
    lifter = Lifter(12, Fixed([0, 5], [1.0, -2.0]), Periodic([7, 9], [2, 3]), Fixed([11], 4.0))
    m = lifter.dof_map()
    assert m.dtype == np.int32 and m.shape == (12,)
    assert (m[[0, 5, 11]] == -1).all() and m[7] == m[2] and m[9] == m[3]
    v = rng.normal(size=lifter.size_reduced)
    np.testing.assert_array_equal(lifter.homogeneous().lift_from_zeros(v), np.where(m >= 0, v[np.maximum(m, 0)], 0.0))
    r = rng.normal(size=12)
    np.testing.assert_allclose(lifter.reduce_adjoint(r), np.bincount(m[m >= 0], weights=r[m >= 0], minlength=lifter.size_reduced))


def test_deprecated_aliases_warn_like_the_reference():
    import warnings
    from tatva_b200 import lifter as L

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert L.DirichletBC is L.Fixed and L.PeriodicMap is L.Periodic
    assert sum(issubclass(x.category, DeprecationWarning) for x in w) == 2


def test_lifter_matches_the_reference_on_a_constraint_chain(golden):
    """The UNMODIFIED reference lifter (lifter/base.py:201-251, constraints.py:184-318) on a Tri3 6x6 mesh with 2 DOFs per
    node: fixed bottom edge, runtime-valued and array-valued Fixed constraints on the top edge, left/right periodicity.
    Free DOFs, lift, lift_from_zeros, reduce, reduce_adjoint and the `lifted` decorator must reproduce its outputs;
    `dof_map` (what the fused HVP kernel consumes) is checked against the same data."""
    import numpy as np
    from tatva_b200.lifter import Fixed, Lifter, Periodic, RuntimeValue, lifted

    g = lambda k: golden[f"lift_{k}"]  # noqa: E731
    n, bottom, top, left, right = int(g("n")), g("bottom"), g("top"), g("left"), g("right")
    dofs = lambda nodes: (np.asarray(nodes)[:, None] * 2 + np.arange(2)).ravel()  # noqa: E731
    lifter = Lifter(
        n,
        Fixed(dofs(bottom), 0.0),
        Fixed(top * 2 + 1, RuntimeValue("top_uy")),
        Fixed(top * 2, 0.01 * np.arange(len(top))),
        Periodic(dofs=dofs(right), master_dofs=dofs(left)),
    ).with_values({"top_uy": 0.07})
    np.testing.assert_array_equal(lifter.free_dofs, g("free_dofs"))
    u_red, base, r_full = g("u_red"), g("base"), g("r_full")
    np.testing.assert_array_equal(lifter.lift_from_zeros(u_red), g("from_zeros"))
    np.testing.assert_array_equal(lifter.lift(u_red, base), g("on_base"))
    np.testing.assert_array_equal(lifter.reduce(base), g("reduce"))
    np.testing.assert_allclose(lifter.reduce_adjoint(r_full), g("reduce_adjoint"), rtol=1e-15, atol=1e-15)
    A = golden["lift_A"]
    np.testing.assert_allclose(lifted(lambda uf: A @ uf, argnums=0, output="dual")(lifter, u_red), golden["lifted_dual"], rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(lifted(lambda uf: uf * 2.0, argnums=0, output="primal")(lifter, u_red), golden["lifted_primal"])
    # sparsity adaptation on the mesh pattern: augmented by the periodic coupling, then reduced to the free DOFs
    from oracle import tatva_oracle as orc
    from tatva_b200 import sparse
    from tatva_b200.mesh import Mesh

    c, el = orc.mesh_unit_square_tri(6, 6)
    pat = sparse.pattern_from_mesh(Mesh(coords=c, elements=el), 2)
    for name, mtx in (("augmented", lifter.augment_sparsity(pat)), ("adapted", lifter.adapt_sparsity(pat))):
        mtx = mtx.tocsr()
        mtx.sort_indices()
        np.testing.assert_array_equal(mtx.indptr, golden[f"lift_sp_{name}_indptr"])
        np.testing.assert_array_equal(mtx.indices, golden[f"lift_sp_{name}_indices"])
    # dof_map == the homogeneous part of the reference's lift and the adjoint of its transpose chain
    m = lifter.dof_map()
    hom = lifter.homogeneous().lift_from_zeros(u_red)
    np.testing.assert_array_equal(hom, np.where(m >= 0, u_red[np.maximum(m, 0)], 0.0))
    np.testing.assert_allclose(np.bincount(m[m >= 0], weights=r_full[m >= 0], minlength=lifter.size_reduced), g("reduce_adjoint"), rtol=1e-15, atol=1e-15)


def test_periodic_chains_follow_the_sequential_semantics_of_the_reference():
    """Masters that still read the base vector (reference lifter/base.py:201-229 applies the constraints one after the
    other on `u_full`): a master that is itself a slave of the SAME Periodic hands over its OLD (base) value, and a
    master fixed by a LATER constraint hands over the base entry, not the later value."""
    from tatva_b200.lifter import Fixed, Lifter, Periodic

    n = 8
    # dofs [1, 2] follow masters [0, 1]: 1 <- 0 and 2 <- OLD 1 (gathered before the set)
    lifter = Lifter(n, Periodic([1, 2], [0, 1]), Periodic([5], [6]), Fixed([6], 3.0))
    free = np.asarray(lifter.free_dofs)
    assert list(free) == [0, 3, 4, 7]
    ur = np.array([10.0, 11.0, 12.0, 13.0])
    base = np.arange(100.0, 108.0)

    def reference_lift(u_red, u_full):
        u = np.array(u_full, copy=True)
        u[free] = u_red
        u[[1, 2]] = u[[0, 1]]
        u[[5]] = u[[6]]
        u[[6]] = 3.0
        return u

    np.testing.assert_array_equal(lifter.lift(ur, base), reference_lift(ur, base))
    np.testing.assert_array_equal(lifter.lift_from_zeros(ur), reference_lift(ur, np.zeros(n)))
    # and the outputs of the UNMODIFIED reference Lifter on the same chain (tests/golden/make_golden.py)
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
    np.testing.assert_array_equal(free, g["lift_chain_free_dofs"])
    np.testing.assert_array_equal(lifter.lift(ur, base), g["lift_chain_on_base"])
    np.testing.assert_array_equal(lifter.lift_from_zeros(ur), g["lift_chain_from_zeros"])
    np.testing.assert_array_equal(lifter.reduce_adjoint(np.arange(1.0, 9.0)), g["lift_chain_reduce_adjoint"])
    # the transpose drops what a base-reading slave collects
    r = np.arange(1.0, 9.0)
    np.testing.assert_array_equal(lifter.reduce_adjoint(r), np.array([r[0] + r[1], r[3], r[4], r[7]]))
    assert list(lifter.dof_map()) == [0, 0, -1, 1, 2, -1, -1, 3]
