"""User-supplied energy densities (materials.UserLaw): written once on symbols, differentiated source-to-source
(lawgen), compiled at run time by NVRTC into the fused kernel templates.  Checked against `torch.func` autodiff of the
SAME density evaluated through a vmap-structured energy (the structure of reference operator.py:194-223), against the
built-in neo-Hookean kernels, and against the autograd route through the Operator building blocks."""
import numpy as np
import pytest
import torch

from oracle import tatva_oracle as orc
from test_gpu_parity import _assert_close, _case, _make_op, _rel

sp = pytest.importorskip("sympy")
pytestmark = pytest.mark.gpu


def mooney_rivlin(G, c1, c2, kappa, lib=sp):
    eye = sp.eye(3) if lib is sp else torch.eye(3, dtype=torch.float64, device=G.device)
    F = eye + G
    Cm = F.T @ F if lib is torch else F.T * F
    J = F.det() if lib is sp else torch.linalg.det(F)
    I1 = Cm.trace() if lib is sp else torch.trace(Cm)
    CC = (Cm * Cm).trace() if lib is sp else torch.trace(Cm @ Cm)
    I2 = (I1**2 - CC) / 2
    third = sp.Rational(1, 3) if lib is sp else 1.0 / 3.0
    return c1 * (J ** (-2 * third) * I1 - 3) + c2 * (J ** (-4 * third) * I2 - 3) + kappa / 2 * (J - 1) ** 2


def st_venant(G, mu, lmbda, lib=sp):
    eye = sp.eye(3) if lib is sp else torch.eye(3, dtype=torch.float64, device=G.device)
    F = eye + G
    E = ((F.T @ F if lib is torch else F.T * F) - eye) / 2
    trE = E.trace() if lib is sp else torch.trace(E)
    trEE = (E * E).trace() if lib is sp else torch.trace(E @ E)
    return lmbda / 2 * trE**2 + mu * trEE


def neo_hookean(G, mu, lmbda):
    F = sp.eye(3) + G
    lnJ = sp.log(F.det())
    return mu / 2 * ((F.T * F).trace() - 3 - 2 * lnJ) + lmbda / 2 * lnJ**2


def _torch_energy(kind, c, el, psi, prm):
    """E(u) with the structure of the reference: gather -> per (element, point) gradient -> psi -> weights -> sum."""
    dNdX, detJ = orc.geometry(kind, c, el)
    _, w = orc.quad_rule(kind)
    dNdX_t, W_t, conn = torch.as_tensor(dNdX), torch.as_tensor(detJ * w), torch.as_tensor(el.astype(np.int64))

    def E(u):
        G = torch.einsum("eqdn,eni->eqid", dNdX_t, u[conn])
        dens = torch.func.vmap(torch.func.vmap(lambda g: psi(g, *prm, lib=torch)))(G)
        return (dens * W_t).sum()

    return E


@pytest.mark.parametrize("kind,n", [("hex8", 5), ("tet4", 4)])
@pytest.mark.parametrize("law_name", ["mooney_rivlin", "st_venant"])
def test_user_law_matches_torch_func_on_the_same_density(kind, n, law_name):
    from tatva_b200 import materials

    psi, prm = {"mooney_rivlin": (mooney_rivlin, (120.0, 30.0, 900.0)), "st_venant": (st_venant, (80.0, 120.0))}[law_name]
    c, el, u, v, _ = _case(kind, n)
    law = materials.UserLaw.from_psi(psi, prm)
    op = _make_op(kind, c, el)
    E = _torch_energy(kind, c, el, psi, prm)
    ut, vt = torch.as_tensor(u), torch.as_tensor(v)
    e_ref = float(E(ut))
    r_ref = torch.func.grad(E)(ut).numpy()
    hv_ref = torch.func.jvp(torch.func.grad(E), (ut,), (vt,))[1].numpy()
    assert abs(float(op.energy(law)(u)) - e_ref) <= 1e-12 * abs(e_ref)
    _assert_close(op.residual(law)(u), r_ref)
    _assert_close(op.hvp(law)(u, v), hv_ref)
    # other parameter values reuse the compiled module
    law2 = law.with_params(tuple(2 * p for p in prm))
    assert law2.material_id == law.material_id
    _assert_close(op.hvp(law2)(u, v), 2 * hv_ref)


def test_user_neo_hookean_equals_the_built_in_kernels_and_assembles():
    """The reference's own density (tests/test_sparse_tracer.py:103-115) as a USER law: same energy / residual / HVP as
    the hand-written kernels and the oracle; sparse.jacfwd falls back to one fused HVP per colour and reproduces the
    directly assembled matrix; Hessian diagonal and the lifted HVP go through the run-time compiled kernels too."""
    import scipy.sparse as sps

    from tatva_b200 import materials, sparse
    from tatva_b200.lifter import Fixed, Lifter
    from tatva_b200.solver import ReducedOperator

    c, el, u, v, (_, omat) = _case("hex8", 4)
    op = _make_op("hex8", c, el)
    law = materials.UserLaw.from_psi(neo_hookean, (500.0, 1000.0))
    nh = materials.NeoHookean(500.0, 1000.0)
    assert abs(float(op.energy(law)(u)) - orc.energy("hex8", omat, c, el, u)) <= 1e-12 * abs(orc.energy("hex8", omat, c, el, u))
    _assert_close(op.residual(law)(u), orc.residual("hex8", omat, c, el, u))
    _assert_close(op.hvp(law)(u, v), orc.hvp("hex8", omat, c, el, u, v))
    pat = sparse.pattern_from_mesh(op.mesh, 3)
    cm = sparse.ColoredMatrix.from_csr(pat)
    K_user = sparse.jacfwd(op.residual(law), cm)(u)
    K_nh = sparse.jacfwd(op.residual(nh), cm)(u)
    d_user = K_user.data.cpu().numpy() if isinstance(K_user.data, torch.Tensor) else np.asarray(K_user.data)
    d_nh = K_nh.data.cpu().numpy() if isinstance(K_nh.data, torch.Tensor) else np.asarray(K_nh.data)
    assert _rel(d_user, d_nh) < 1e-12
    K = sps.csr_matrix((d_user, pat.indices, pat.indptr), shape=pat.shape)
    _assert_close(op.hessian_diagonal(law, torch.as_tensor(u, device="cuda")).reshape(-1), K.diagonal())
    fixed = np.where(c[:, 2] < 1e-9 + c[:, 2].min())[0]
    lifter = Lifter(c.size, Fixed((fixed[:, None] * 3 + np.arange(3)).ravel()))
    ra, rb = ReducedOperator(op, law, lifter), ReducedOperator(op, nh, lifter)
    u_red = lifter.reduce(torch.as_tensor(u.ravel(), device="cuda"))
    vr = torch.as_tensor(np.random.default_rng(0).normal(size=lifter.size_reduced), device="cuda")
    for r in (ra, rb):
        r.set_state(u_red)
    ya, yb = ra.matvec(vr, torch.empty_like(vr)), rb.matvec(vr, torch.empty_like(vr))
    assert float((ya - yb).norm() / yb.norm()) < 1e-12


def test_user_law_with_a_custom_quadrature_rule():
    from tatva_b200 import element, materials
    import tatva_b200

    c, el, u, v, _ = _case("hex8", 4)
    qp, qw = orc.gauss_rule("hex8", 3)
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8(quad_points=qp, quad_weights=qw))
    prm = (80.0, 120.0)
    law = materials.UserLaw.from_psi(st_venant, prm)
    with orc.custom_rule("hex8", qp, qw):
        E = _torch_energy("hex8", c, el, st_venant, prm)
    ut, vt = torch.as_tensor(u), torch.as_tensor(v)
    _assert_close(op.hvp(law)(u, v), torch.func.jvp(torch.func.grad(E), (ut,), (vt,))[1].numpy())


def test_a_broken_user_source_fails_loudly_with_the_compiler_log():
    from tatva_b200 import _lib, materials

    c, el, u, v, _ = _case("hex8", 3)
    op = _make_op("hex8", c, el)
    bad = materials.UserLaw("struct UserLaw { this is not CUDA };", (1.0,))
    with pytest.raises(_lib.TatvaError):
        op.energy(bad)(u)
    assert "error" in materials.UserLaw.compile_log()
