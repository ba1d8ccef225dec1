"""GPU parity: direct CSR assembly and the coloured-JVP path against the oracle; the two-field
(compound) kernels; halo pack/unpack kernels."""
import numpy as np
import pytest
import scipy.sparse as sps
import torch

from oracle import tatva_oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return max(np.linalg.norm((a - b).ravel()) / (den or 1.0), np.abs(a - b).max() / (np.abs(b).max() or 1.0))


def _setup(kind, c, el):
    import tatva_b200
    from tatva_b200 import element

    cls = {"tri3": element.Tri3, "tet4": element.Tetrahedron4, "hex8": element.Hexahedron8}[kind]
    mesh = tatva_b200.Mesh(coords=c, elements=el)
    return mesh, tatva_b200.Operator(mesh, cls())


def test_sparse_matrix_tri3_reference_case():
    """reference tests/test_sparse.py:48-80: Tri3 8x8, mu=1, lambda=0, u=0; K_sparse == dense Hessian,
    linearized primal == grad."""
    from tatva_b200 import materials, sparse

    c, el = orc.mesh_unit_square_tri(8, 8)
    mesh, op = _setup("tri3", c, el)
    mat, omat = materials.LinearElastic(1.0, 0.0), orc.LinearElastic(1.0, 0.0)
    n = 2 * len(c)
    pat = sparse.pattern_from_mesh(mesh, 2)
    cm = sparse.ColoredMatrix.from_csr(pat)
    u0 = torch.zeros((len(c), 2), dtype=torch.float64, device="cuda")
    K = np.stack([orc.hvp("tri3", omat, c, el, np.zeros_like(c), e.reshape(-1, 2)).ravel() for e in np.eye(n)], axis=1)
    K_direct = sparse.jacfwd(op.residual(mat), cm)(u0)
    assert _rel(K_direct.to_dense(), K) < 1e-12
    # the reference algorithm (one JVP per colour) on a plain callable
    K_col = sparse.jacfwd(lambda x: op.residual(mat)(x), cm, color_batch_size=10)(u0)
    assert _rel(K_col.to_dense(), K) < 1e-12
    primal, K_lin = sparse.linearized_jacfwd(lambda x: op.residual(mat)(x), cm, color_batch_size=10)(u0)
    assert _rel(K_lin.to_dense(), K) < 1e-12
    assert float(primal.abs().max()) == 0.0
    primal2, K_lin2 = sparse.linearized_jacfwd(op.residual(mat), cm)(u0)
    assert _rel(K_lin2.data, K_direct.data.cpu().numpy()) == 0.0 and float(primal2.abs().max()) == 0.0


@pytest.mark.parametrize("kind,n", [("tet4", 4), ("hex8", 3), ("tri3", 12)])
def test_direct_assembly_matches_coloured_oracle(kind, n):
    from tatva_b200 import materials, sparse

    rng = np.random.default_rng(0)
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(n, n)
        mat, omat, dpn = materials.LinearElastic(0.38, 0.58), orc.LinearElastic(0.38, 0.58), 2
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1, 1, 1), (n, n, n))
        mat, omat, dpn = materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0), 3
    else:
        c, el = orc.mesh_box_hex(n)
        mat, omat, dpn = materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0), 3
    c = c + 0.1 / n * rng.uniform(-1, 1, c.shape)
    u = 0.02 * rng.normal(size=c.shape)
    mesh, op = _setup(kind, c, el)
    pat = sparse.pattern_from_mesh(mesh, dpn)
    cm = sparse.ColoredMatrix.from_csr(pat)
    asm = sparse.assembler(op, mat, cm)
    data = asm(u).cpu().numpy()
    # r02 default for Tri3 / Tet4: the tiled kernel (duplicate blocks combined on chip); the element-per-thread kernel
    # with one RED group per element block must give the same matrix
    assert asm.tiled == (kind != "hex8")
    if asm.tiled:
        assert asm.tile_stats["combine_ratio"] > 1.5
        plain = sparse.assembler(op, mat, cm, tiled=False)(u).cpu().numpy()
        assert _rel(plain, data) < 1e-13
    sym = sparse.assembler(op, mat, cm, symmetric=True)(u).cpu().numpy()  # upper triangle by RED + mirror pass
    assert _rel(sym, data) < 1e-13
    if kind != "hex8":  # both kernels: per-entry atomics (default) and the deterministic row-wise one
        rows1 = sparse.assembler(op, mat, cm, by_rows=True)(u).cpu().numpy()
        assert _rel(rows1, data) < 1e-13
        rows2 = sparse.assembler(op, mat, cm, by_rows=True)(u).cpu().numpy()
        assert np.array_equal(rows1, rows2), "row-wise assembly must be bitwise reproducible"
    # oracle 1: the reference algorithm, n_colors HVPs + decompression
    ndof = dpn * len(c)
    jvp = lambda seed: orc.hvp(kind, omat, c, el, u, seed.reshape(-1, dpn)).ravel()  # noqa: E731
    ref = orc.colored_jacobian_data(jvp, ndof, pat.indptr, pat.indices, np.asarray(cm.colors))
    assert _rel(data, ref) < 1e-12
    # oracle 2: direct element-stiffness assembly
    ref2 = orc.assemble_csr_data(kind, omat, c, el, u, pat.indptr, pat.indices)
    assert _rel(data, ref2) < 1e-12
    # symmetry and consistency with the matrix-free HVP at a size-independent level
    K = sps.csr_matrix((data, pat.indices, pat.indptr), shape=(ndof, ndof))
    v = rng.normal(size=ndof)
    Hv = op.hvp(mat)(u, v.reshape(-1, dpn)).cpu().numpy().ravel()
    assert _rel(K @ v, Hv) < 1e-12
    assert abs(K - K.T).max() < 1e-10 * abs(K).max()


@pytest.mark.parametrize("kind,law", [("tri3", "le"), ("tet4", "nh"), ("hex8", "nh"), ("hex8", "le"), ("tet4", "pf"), ("hex8", "pf")])
def test_hessian_diagonal_matches_the_oracle_matrix(kind, law):
    """`Operator.hessian_diagonal` (Jacobi preconditioner) == diagonal of the oracle's element-stiffness assembly; for
    the two-field law, e_i . H e_i by the oracle HVP on a sample of unit vectors."""
    from tatva_b200 import materials

    rng = np.random.default_rng(3)
    n = {"tri3": 10, "tet4": 4, "hex8": 3}[kind]
    c, el = {"tri3": lambda: orc.mesh_unit_square_tri(n, n), "tet4": lambda: orc.mesh_box_tet((1, 1, 1), (n, n, n)), "hex8": lambda: orc.mesh_box_hex(n)}[kind]()
    c = c + 0.1 / n * rng.uniform(-1, 1, c.shape)
    mesh, op = _setup(kind, c, el)
    dim = c.shape[1]
    if law == "pf":
        prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
        mat, omat = materials.NeoHookeanPhaseField(*prm), orc.NeoHookeanPhaseField(*prm)
        s = np.concatenate([0.02 * rng.normal(size=(len(c), 3)), rng.uniform(0, 0.8, size=(len(c), 1))], axis=1)
        d = op.hessian_diagonal(mat, torch.as_tensor(s, device="cuda")).cpu().numpy().ravel()
        idx = rng.choice(s.size, 40, replace=False)
        ref = np.array([orc.hvp_pf(kind, omat, c, el, s, np.eye(1, s.size, i).reshape(s.shape)).ravel()[i] for i in idx])
        assert _rel(d[idx], ref) < 1e-12
        return
    mat, omat = (materials.LinearElastic(0.38, 0.58), orc.LinearElastic(0.38, 0.58)) if law == "le" else (materials.NeoHookean(500.0, 1000.0), orc.NeoHookean(500.0, 1000.0))
    u = 0.02 * rng.normal(size=c.shape)
    ip, ix = orc.pattern_from_mesh(el, len(c), dim)
    data = orc.assemble_csr_data(kind, omat, c, el, u, ip, ix)
    ref = sps.csr_matrix((data, ix, ip), shape=(c.size, c.size)).diagonal()
    d = op.hessian_diagonal(mat, torch.as_tensor(u, device="cuda"))
    assert d.shape == u.shape
    assert _rel(d.reshape(-1), ref) < 1e-12


@pytest.mark.parametrize("kind,n", [("tet4", 4), ("hex8", 3)])
def test_phase_field_two_field_kernels(kind, n):
    """Config 5: compound (u, phi) state, node-interleaved [ux,uy,uz,phi]."""
    from tatva_b200 import materials, sparse
    from tatva_b200.compound import Compound, FieldSize, field

    rng = np.random.default_rng(1)
    c, el = orc.mesh_box_tet((1, 1, 1), (n, n, n)) if kind == "tet4" else orc.mesh_box_hex(n)
    c = c + 0.1 / n * rng.uniform(-1, 1, c.shape)
    mesh, op = _setup(kind, c, el)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    mat, omat = materials.NeoHookeanPhaseField(*prm), orc.NeoHookeanPhaseField(*prm)

    class State(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 3))
        phi = field(shape=(FieldSize.AUTO,))

    s = np.concatenate([0.02 * rng.normal(size=(len(c), 3)), rng.uniform(0, 0.8, size=(len(c), 1))], axis=1)
    t = rng.normal(size=s.shape)
    arr = torch.as_tensor(s.ravel(), device="cuda")
    st = State(arr)
    assert st.u.shape == (len(c), 3) and st.phi.shape == (len(c),)
    np.testing.assert_array_equal(st.phi.cpu().numpy(), s[:, 3])
    assert _rel(op.energy(mat)(arr), orc.energy_pf(kind, omat, c, el, s)) < 1e-12
    assert _rel(op.residual(mat)(arr).reshape(-1, 4), orc.residual_pf(kind, omat, c, el, s)) < 1e-12
    assert _rel(op.hvp(mat)(arr, torch.as_tensor(t.ravel(), device="cuda")).reshape(-1, 4), orc.hvp_pf(kind, omat, c, el, s, t)) < 1e-12
    # coupled Jacobian on pattern_from_compound
    pat = sparse.pattern_from_compound(State)
    cm = sparse.ColoredMatrix.from_csr(pat)
    data = sparse.assembler(op, mat, cm)(arr).cpu().numpy()
    K = sps.csr_matrix((data, pat.indices, pat.indptr), shape=pat.shape)
    assert _rel(K @ t.ravel(), orc.hvp_pf(kind, omat, c, el, s, t).ravel()) < 1e-12


def test_halo_kernels_and_single_rank_plan():
    from tatva_b200.mpi import ExchangePlan, _LocalLayout

    n = 1000
    rng = np.random.default_rng(0)
    perm = rng.permutation(n).astype(np.int32)
    layout = _LocalLayout(perm, 0, n, n, n, np.ones(n, dtype=bool), perm)
    plan = ExchangePlan(layout)
    x = torch.as_tensor(rng.normal(size=n), device="cuda")
    ul = plan.make_scatter_fwd_set()(x)
    np.testing.assert_array_equal(ul.cpu().numpy(), x.cpu().numpy()[perm])
    back = plan.make_scatter_rev_add(lambda u: u)(ul)
    np.testing.assert_array_equal(back.cpu().numpy(), x.cpu().numpy())
    # repeated indices accumulate
    from tatva_b200 import _lib

    idx = torch.as_tensor(rng.integers(0, 10, size=5000), device="cuda")
    vals = torch.as_tensor(rng.normal(size=5000), device="cuda")
    dst = torch.zeros(10, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().tatva_halo_unpack_add(vals.data_ptr(), idx.data_ptr(), 5000, dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ref = np.zeros(10)
    np.add.at(ref, idx.cpu().numpy(), vals.cpu().numpy())
    np.testing.assert_allclose(dst.cpu().numpy(), ref, rtol=1e-12)
