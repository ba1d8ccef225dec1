"""Device-resident CG / Newton step around the HVP kernel (GPU): against SciPy on the oracle's matrices."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import torch

from oracle import tatva_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(n=5):
    import tatva_b200
    from tatva_b200 import element, materials
    from tatva_b200.lifter import Fixed, Lifter

    rng = np.random.default_rng(0)
    c, el = orc.mesh_box_hex(n)
    c = c + 0.01 * rng.uniform(-1, 1, c.shape)
    fixed = np.where(c[:, 2] < 0.02)[0]
    top = np.where(c[:, 2] > 0.98)[0]
    lifter = Lifter(c.size, Fixed((fixed[:, None] * 3 + np.arange(3)).ravel()), Fixed(top * 3 + 2, 0.08), Fixed((top[:, None] * 3 + np.arange(2)).ravel()))
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8())
    return c, el, lifter, op, materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0)


def _oracle_K(c, el, omat, u_full, lifter):
    ip, ix = orc.pattern_from_mesh(el, len(c), 3)
    data = orc.assemble_csr_data("hex8", omat, c, el, u_full.reshape(-1, 3), ip, ix)
    K = sps.csr_matrix((data, ix, ip), shape=(c.size, c.size))
    f = lifter.free_dofs
    return K[f][:, f]


@pytest.mark.parametrize("use_graph", [False, True])
def test_cg_solves_the_reduced_tangent_system(use_graph):
    from tatva_b200.solver import ConjugateGradient, ReducedOperator

    c, el, lifter, op, mat, omat = _problem()
    rng = np.random.default_rng(1)
    u_red = 0.005 * rng.normal(size=lifter.size_reduced)
    b = rng.normal(size=lifter.size_reduced)
    red = ReducedOperator(op, mat, lifter)
    red.set_state(torch.as_tensor(u_red, device="cuda"))
    cg = ConjugateGradient(red.matvec, lifter.size_reduced, "cuda", use_graph=use_graph)
    x, info = cg.solve(torch.as_tensor(b, device="cuda"), tol=1e-12, maxiter=2000, check_every=20)
    assert info["converged"], info
    K = _oracle_K(c, el, omat, lifter.lift_from_zeros(u_red), lifter)
    x_ref = spla.spsolve(K.tocsc(), b)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) / np.linalg.norm(x_ref) < 1e-9
    # a second solve reuses the captured graph
    x2, info2 = cg.solve(torch.as_tensor(2 * b, device="cuda"), tol=1e-12, maxiter=2000, check_every=20)
    assert info2["converged"] and np.linalg.norm(x2.cpu().numpy() - 2 * x_ref) / np.linalg.norm(x_ref) < 1e-8


@pytest.mark.parametrize("use_graph", [False, True])
def test_jacobi_pcg_matches_the_direct_solve_and_needs_fewer_iterations(use_graph):
    from tatva_b200.solver import ConjugateGradient, ReducedOperator

    c, el, lifter, op, mat, omat = _problem()
    # a stretched mesh makes the diagonal vary, so Jacobi has something to do
    rng = np.random.default_rng(2)
    u_red = 0.005 * rng.normal(size=lifter.size_reduced)
    b = rng.normal(size=lifter.size_reduced)
    red = ReducedOperator(op, mat, lifter)
    red.set_state(torch.as_tensor(u_red, device="cuda"))
    K = _oracle_K(c, el, omat, lifter.lift_from_zeros(u_red), lifter)
    diag = red.diagonal()
    np.testing.assert_allclose(diag.cpu().numpy(), K.diagonal(), rtol=1e-12)
    pcg = ConjugateGradient(red.matvec, lifter.size_reduced, "cuda", use_graph=use_graph, jacobi=True)
    pcg.set_diagonal(diag)
    x, info = pcg.solve(torch.as_tensor(b, device="cuda"), tol=1e-12, maxiter=2000, check_every=5)
    assert info["converged"], info
    x_ref = spla.spsolve(K.tocsc(), b)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) / np.linalg.norm(x_ref) < 1e-9
    # same iterates as SciPy-free textbook PCG on the oracle matrix (iteration count within the check granularity)
    M = 1.0 / K.diagonal()
    xr, r = np.zeros_like(b), b.copy()
    z = M * r
    p, rz, it = z.copy(), r @ z, 0
    while np.linalg.norm(r) > 1e-12 * np.linalg.norm(b) and it < 2000:
        Ap = K @ p
        a = rz / (p @ Ap)
        xr += a * p
        r -= a * Ap
        z = M * r
        rz, rz_old = r @ z, rz
        p = z + (rz / rz_old) * p
        it += 1
    assert abs(info["iterations"] - it) <= 5 + 0.1 * it, (info, it)
    # identity diagonal == plain CG
    pcg.set_diagonal(torch.ones_like(diag))
    x1, info1 = pcg.solve(torch.as_tensor(b, device="cuda"), tol=1e-12, maxiter=2000, check_every=5)
    cg = ConjugateGradient(red.matvec, lifter.size_reduced, "cuda", use_graph=False)
    x0, info0 = cg.solve(torch.as_tensor(b, device="cuda"), tol=1e-12, maxiter=2000, check_every=5)
    assert info1["iterations"] == info0["iterations"]
    assert np.linalg.norm((x1 - x0).cpu().numpy()) / np.linalg.norm(x_ref) < 1e-9


@pytest.mark.parametrize("case", ["hex8_nh", "hex8_nh_generic", "tet4_pf", "tri3_le"])
def test_lifter_fused_into_the_hvp_kernel_matches_lift_hvp_reduce(case):
    """`tatva_hvp_lifted`: Fixed + Periodic folded into the gather / scatter == lift_0 -> HVP -> reduce_adjoint
    (reference tatva/lifter/base.py:201-251), and == the oracle HVP pushed through the NumPy lifter."""
    import tatva_b200
    from tatva_b200 import element, materials
    from tatva_b200.lifter import Fixed, Lifter, Periodic
    from tatva_b200.solver import ReducedOperator

    rng = np.random.default_rng(5)
    if case.startswith("hex8"):
        c, el = orc.mesh_box_hex(4)
        cls, kind, dpn = element.Hexahedron8, "hex8", 3
        mat, omat = materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0)
    elif case == "tet4_pf":
        c, el = orc.mesh_box_tet((1, 1, 1), (4, 4, 4))
        cls, kind, dpn = element.Tetrahedron4, "tet4", 4
        prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
        mat, omat = materials.NeoHookeanPhaseField(*prm), orc.NeoHookeanPhaseField(*prm)
    else:
        c, el = orc.mesh_unit_square_tri(6, 6)
        cls, kind, dpn = element.Tri3, "tri3", 2
        mat, omat = materials.LinearElastic(0.38, 0.58), orc.LinearElastic(0.38, 0.58)
    x = c[:, 0]
    lo, hi = x.min(), x.max()
    left, right = np.where(np.isclose(x, lo))[0], np.where(np.isclose(x, hi))[0]
    # pair right-face nodes with the left-face nodes at the same transverse position
    key = lambda idx: np.lexsort(c[idx][:, ::-1][:, :-1].T) if c.shape[1] > 1 else np.arange(len(idx))  # noqa: E731
    left, right = left[key(left)], right[key(right)]
    assert np.allclose(c[left][:, 1:], c[right][:, 1:])
    bottom = np.where(np.isclose(c[:, -1], c[:, -1].min()))[0]
    bottom = np.setdiff1d(bottom, np.concatenate([left, right]))
    lifter = Lifter(
        c.shape[0] * dpn,
        Fixed((bottom[:, None] * dpn + np.arange(dpn)).ravel(), 0.0),
        Fixed(left * dpn, 0.01),
        Periodic((right[:, None] * dpn + np.arange(1, dpn)).ravel(), (left[:, None] * dpn + np.arange(1, dpn)).ravel()),
        Fixed(right * dpn, 0.02),
    )
    cj = c + 0.02 * rng.uniform(-1, 1, c.shape) * (c.shape[1] == 3)
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=cj, elements=el), cls())
    if case.endswith("generic"):
        op.set_variant(1)
    u_red = 0.01 * rng.normal(size=lifter.size_reduced)
    if dpn == 4:
        u_red = np.abs(u_red)
    v_red = rng.normal(size=lifter.size_reduced)
    fused, plain = ReducedOperator(op, mat, lifter, fused=True), ReducedOperator(op, mat, lifter, fused=False)
    assert fused.dof_map is not None and plain.dof_map is None
    for r in (fused, plain):
        r.set_state(torch.as_tensor(u_red, device="cuda"))
    v = torch.as_tensor(v_red, device="cuda")
    y1 = fused.matvec(v, torch.empty_like(v)).cpu().numpy()
    y0 = plain.matvec(v, torch.empty_like(v)).cpu().numpy()
    assert np.linalg.norm(y1 - y0) / np.linalg.norm(y0) < 1e-13
    # oracle: NumPy lifter around the oracle HVP
    hom = lifter.homogeneous()
    u_full = lifter.lift_from_zeros(u_red).reshape(-1, dpn)
    v_full = hom.lift_from_zeros(v_red).reshape(-1, dpn)
    Hv = orc.hvp_pf(kind, omat, cj, el, u_full, v_full) if dpn == 4 else orc.hvp(kind, omat, cj, el, u_full, v_full)
    ref = lifter.reduce_adjoint(Hv.ravel())
    assert np.linalg.norm(y1 - ref) / np.linalg.norm(ref) < 1e-12


@pytest.mark.parametrize("case", ["hex8_nh", "hex8_nh_generic", "tet4_pf"])
def test_hvp_lifted_dot_delivers_v_dot_Hv(case):
    """`tatva_hvp_lifted_dot`: the element-local sum of v_e . y_e (Hex8 x neo-Hookean kernel) and the two-pass dot of
    the other (element, law) pairs both equal v_red . (K_red v_red), with Fixed and Periodic constraints in the lifter;
    y accumulates into the caller's zeroed vector and scalars[0] <- scalars[2]."""
    import tatva_b200
    from tatva_b200 import element, materials
    from tatva_b200.lifter import Fixed, Lifter, Periodic
    from tatva_b200.solver import ReducedOperator

    rng = np.random.default_rng(11)
    if case.startswith("hex8"):
        c, el = orc.mesh_box_hex(7)  # 343 elements: 3 CTAs, the last one partial
        cls, dpn, mat = element.Hexahedron8, 3, materials.NeoHookean(500.0, 1000.0)
    else:
        c, el = orc.mesh_box_tet((1, 1, 1), (4, 4, 4))
        cls, dpn, mat = element.Tetrahedron4, 4, materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.1, 1e-6)
    x = c[:, 0]
    left, right = np.where(np.isclose(x, x.min()))[0], np.where(np.isclose(x, x.max()))[0]
    key = lambda idx: np.lexsort(c[idx][:, ::-1][:, :-1].T)  # noqa: E731
    left, right = left[key(left)], right[key(right)]
    bottom = np.setdiff1d(np.where(np.isclose(c[:, 2], 0.0))[0], np.concatenate([left, right]))
    lifter = Lifter(c.shape[0] * dpn, Fixed((bottom[:, None] * dpn + np.arange(dpn)).ravel(), 0.0),
                    Periodic((right[:, None] * dpn + np.arange(dpn)).ravel(), (left[:, None] * dpn + np.arange(dpn)).ravel()))
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c + 0.02 * rng.uniform(-1, 1, c.shape), elements=el), cls())
    if case.endswith("generic"):
        op.set_variant(1)
    red = ReducedOperator(op, mat, lifter)
    u_red = 0.01 * np.abs(rng.normal(size=lifter.size_reduced))
    red.set_state(torch.as_tensor(u_red, device="cuda"))
    v = torch.as_tensor(rng.normal(size=lifter.size_reduced), device="cuda")
    y_ref = red.matvec(v, torch.empty_like(v))
    y = torch.zeros_like(v)
    scalars = torch.tensor([1.0, 0.0, 7.0, 0, 0, 0, 0, 0], dtype=torch.float64, device="cuda")
    partials = torch.zeros(2 * 1184, dtype=torch.float64, device="cuda")
    red.matvec_dot(v, y, partials, scalars)
    assert float((y - y_ref).norm() / y_ref.norm()) < 1e-13
    want = float((v * y_ref).sum())
    s = scalars.cpu().numpy()
    assert abs(s[1] - want) <= 1e-12 * abs(want), (s[1], want)
    assert s[0] == 7.0 and s[2] == 7.0  # rolled


@pytest.mark.parametrize("jacobi", [False, True])
def test_cg_with_the_dot_fused_into_the_hvp_gives_the_same_iterates(jacobi):
    """r02 iteration (p.Ap from the HVP kernel, direction pass clears Ap, 6 launches) against the r01 iteration:
    the same iterates to 1e-13 after 40 steps, eager and graph."""
    from tatva_b200.solver import ConjugateGradient, ReducedOperator

    c, el, lifter, op, mat, omat = _problem(6)
    rng = np.random.default_rng(2)
    red = ReducedOperator(op, mat, lifter)
    red.set_state(torch.as_tensor(0.005 * rng.normal(size=lifter.size_reduced), device="cuda"))
    b = torch.as_tensor(rng.normal(size=lifter.size_reduced), device="cuda")
    n = lifter.size_reduced
    diag = red.diagonal() if jacobi else None
    out = []
    for fused, graph in ((False, False), (True, False), (True, True)):
        cg = ConjugateGradient(red.matvec, n, "cuda", use_graph=graph, jacobi=jacobi, matvec_dot=red.matvec_dot if fused else None)
        if jacobi:
            cg.set_diagonal(diag)
        x, info = cg.solve(b, tol=0.0, maxiter=40, check_every=40)
        assert info["iterations"] == 40
        out.append(x.cpu().numpy())
        x2, _ = cg.solve(b, tol=0.0, maxiter=40, check_every=40)  # a second solve starts from a clean state
        assert np.array_equal(x2.cpu().numpy(), out[-1]) or np.linalg.norm(x2.cpu().numpy() - out[-1]) / np.linalg.norm(out[-1]) < 1e-13
    for x in out[1:]:
        assert np.linalg.norm(x - out[0]) / np.linalg.norm(out[0]) < 1e-13


@pytest.mark.parametrize("fuse_dot", [False, True])
def test_masked_full_space_cg_equals_the_reduced_cg(fuse_dot):
    """Fixed-only lifter: CG on full-size vectors around the UNCONSTRAINED kernel, Dirichlet rows masked in the update
    pass (`MaskedOperator`), reproduces the reduced-space CG (lifted kernel) iterate for iterate."""
    from tatva_b200.solver import ConjugateGradient, MaskedOperator, ReducedOperator

    c, el, lifter, op, mat, omat = _problem(6)
    rng = np.random.default_rng(4)
    u_red = torch.as_tensor(0.005 * rng.normal(size=lifter.size_reduced), device="cuda")
    b = torch.as_tensor(rng.normal(size=lifter.size_reduced), device="cuda")
    red = ReducedOperator(op, mat, lifter)
    red.set_state(u_red)
    x_ref, _ = ConjugateGradient(red.matvec, lifter.size_reduced, "cuda", use_graph=False).solve(b, tol=0.0, maxiter=40, check_every=40)
    mo = MaskedOperator(op, mat, lifter, fuse_dot=fuse_dot)
    mo.set_state(u_red)
    for graph in (False, True):
        x_full, info = mo.solver(use_graph=graph).solve(mo.expand(b), tol=0.0, maxiter=40, check_every=40)
        assert info["iterations"] == 40
        assert float(x_full[mo.fixed_map < 0].abs().max()) == 0.0
        assert float((mo.restrict(x_full) - x_ref).norm() / x_ref.norm()) < 1e-13
    xs, info = mo.solver(use_graph=True).solve(mo.expand(b), tol=1e-11, maxiter=3000, check_every=20)
    assert info["converged"]
    K = _oracle_K(c, el, omat, lifter.lift_from_zeros(u_red.cpu().numpy()), lifter)
    x_dir = spla.spsolve(K.tocsc(), b.cpu().numpy())
    assert np.linalg.norm(mo.restrict(xs).cpu().numpy() - x_dir) / np.linalg.norm(x_dir) < 1e-8


def test_newton_with_jacobi_reaches_the_same_minimiser():
    from tatva_b200.solver import newton_solve

    c, el, lifter, op, mat, omat = _problem(4)
    u1, h1 = newton_solve(op, mat, lifter, tol=1e-10, cg_tol=1e-12)
    u2, h2 = newton_solve(op, mat, lifter, tol=1e-10, cg_tol=1e-12, jacobi=True)
    assert h2[-1]["residual_norm"] <= 1e-10 * h2[0]["residual_norm"] or h2[-1]["residual_norm"] < 1e-9
    assert float((u1 - u2).norm() / u1.norm()) < 1e-8


def test_newton_step_converges_to_the_oracle_minimiser():
    from tatva_b200.solver import newton_solve

    c, el, lifter, op, mat, omat = _problem(4)
    u, hist = newton_solve(op, mat, lifter, tol=1e-9, cg_tol=1e-11)
    assert hist[-1]["residual_norm"] <= 1e-9 * hist[0]["residual_norm"], hist
    assert len(hist) <= 10
    # oracle Newton with direct solves
    ur = np.zeros(lifter.size_reduced)
    for _ in range(12):
        uf = lifter.lift_from_zeros(ur)
        r = lifter.reduce_adjoint(orc.residual("hex8", omat, c, el, uf.reshape(-1, 3)).ravel())
        if np.linalg.norm(r) < 1e-10:
            break
        ur = ur - spla.spsolve(_oracle_K(c, el, omat, uf, lifter).tocsc(), r)
    assert np.linalg.norm(u.cpu().numpy() - ur) / np.linalg.norm(ur) < 1e-7
    e_gpu = float(op.energy(mat)(lifter.lift_from_zeros(u).view(-1, 3)))
    e_ref = orc.energy("hex8", omat, c, el, lifter.lift_from_zeros(ur).reshape(-1, 3))
    assert abs(e_gpu - e_ref) <= 1e-10 * abs(e_ref)


def test_cuda_lift_handles_periodic_chains_like_the_numpy_path():
    """`tatva_lift` with base-entry source codes: a slave whose master still reads the base vector (reference
    lifter/base.py:201-229, `lift_chain_*` fixtures) — CUDA gather kernel == NumPy path == reference outputs."""
    import os

    from tatva_b200.lifter import Fixed, Lifter, Periodic

    lifter = Lifter(8, Periodic([1, 2], [0, 1]), Periodic([5], [6]), Fixed([6], 3.0))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
    ur, base = np.array([10.0, 11.0, 12.0, 13.0]), np.arange(100.0, 108.0)
    out = lifter.lift(torch.as_tensor(ur, device="cuda"), torch.as_tensor(base, device="cuda"))
    np.testing.assert_array_equal(out.cpu().numpy(), g["lift_chain_on_base"])
    out0 = lifter.lift_from_zeros(torch.as_tensor(ur, device="cuda"))
    np.testing.assert_array_equal(out0.cpu().numpy(), g["lift_chain_from_zeros"])
    r = lifter.reduce_adjoint(torch.as_tensor(np.arange(1.0, 9.0), device="cuda"))
    np.testing.assert_array_equal(r.cpu().numpy(), g["lift_chain_reduce_adjoint"])
