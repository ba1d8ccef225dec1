"""Second, independent oracle route: torch.func.jvp(torch.func.grad(E)) of a
vmap-structured energy that mirrors the reference program (gather -> vmap over elements and quadrature
points -> Element.gradient -> psi -> x detJ w -> sum; tatva/operator.py:194-223, element/base.py:99-115)
against the closed-form residual / HVP of oracle/tatva_oracle.py.  Also the only pin of the builder-defined
phase-field law (no counterpart in the reference)."""
import numpy as np
import pytest
import torch

from oracle import tatva_oracle as orc

torch.set_default_dtype(torch.float64)


def _energy_fn(kind, coords, conn, density, n_fields=1):
    qp, qw = orc.quad_rule(kind)
    dNdr = torch.as_tensor(np.stack([orc.shape_function_derivative(kind, x) for x in qp]))
    N = torch.as_tensor(np.stack([orc.shape_function(kind, x) for x in qp]))
    X = torch.as_tensor(coords)[torch.as_tensor(conn, dtype=torch.long)]
    qw = torch.as_tensor(qw)
    idx = torch.as_tensor(conn, dtype=torch.long)

    def per_quad(dn, n, w, ue, xe):
        J = dn @ xe
        dNdX = torch.linalg.inv(J) @ dn
        grad = torch.einsum("dn,nc->cd", dNdX, ue)
        val = torch.einsum("n,nc->c", n, ue)
        return density(grad, val) * torch.linalg.det(J) * w

    def per_element(ue, xe):
        return torch.vmap(per_quad, in_dims=(0, 0, 0, None, None))(dNdr, N, qw, ue, xe).sum()

    def E(u):
        return torch.vmap(per_element)(u[idx], X).sum()

    return E


def _nh_density(mu, lm):
    def psi(G, _val):
        F = torch.eye(3) + G[:3]
        lnJ = torch.log(torch.linalg.det(F))
        return 0.5 * mu * ((F * F).sum() - 3 - 2 * lnJ) + 0.5 * lm * lnJ**2

    return psi


def _le_density(mu, lm):
    def psi(G, _val):
        eps = 0.5 * (G + G.T)
        return mu * (eps * eps).sum() + 0.5 * lm * torch.trace(eps) ** 2

    return psi


def _pf_density(mu, lm, Gc, ell, k):
    nh = _nh_density(mu, lm)

    def psi(G, val):
        phi, gphi = val[3], G[3]
        return ((1 - phi) ** 2 + k) * nh(G, val) + Gc * (phi**2 / (2 * ell) + 0.5 * ell * (gphi * gphi).sum())

    return psi


def _rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))


@pytest.mark.parametrize("kind", ["tri3", "tet4", "hex8"])
def test_closed_forms_match_autodiff(kind):
    rng = np.random.default_rng(2)
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(4, 4)
        mat, dens = orc.LinearElastic(0.38, 0.58), _le_density(0.38, 0.58)
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2))
        mat, dens = orc.NeoHookean(500.0, 1000.0), _nh_density(500.0, 1000.0)
    else:
        c, el = orc.mesh_box_hex(3)
        mat, dens = orc.NeoHookean(500.0, 1000.0), _nh_density(500.0, 1000.0)
    c = c + 0.03 * rng.uniform(-1, 1, c.shape)
    u, v = 0.03 * rng.normal(size=c.shape), rng.normal(size=c.shape)
    E = _energy_fn(kind, c, el, dens)
    ut, vt = torch.as_tensor(u), torch.as_tensor(v)
    assert abs(float(E(ut)) - orc.energy(kind, mat, c, el, u)) <= 1e-13 * abs(float(E(ut)))
    r, Hv = torch.func.jvp(torch.func.grad(E), (ut,), (vt,))
    assert _rel(orc.residual(kind, mat, c, el, u), r.numpy()) < 1e-13
    assert _rel(orc.hvp(kind, mat, c, el, u, v), Hv.numpy()) < 1e-13


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_phase_field_closed_forms_match_autodiff(kind):
    rng = np.random.default_rng(4)
    c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2)) if kind == "tet4" else orc.mesh_box_hex(2)
    c = c + 0.03 * rng.uniform(-1, 1, c.shape)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    mat = orc.NeoHookeanPhaseField(*prm)
    s = np.concatenate([0.03 * rng.normal(size=(len(c), 3)), rng.uniform(0, 0.8, size=(len(c), 1))], axis=1)
    t = rng.normal(size=s.shape)
    E = _energy_fn(kind, c, el, _pf_density(*prm))
    st, tt = torch.as_tensor(s), torch.as_tensor(t)
    assert abs(float(E(st)) - orc.energy_pf(kind, mat, c, el, s)) <= 1e-13 * abs(float(E(st)))
    r, Hv = torch.func.jvp(torch.func.grad(E), (st,), (tt,))
    assert _rel(orc.residual_pf(kind, mat, c, el, s), r.numpy()) < 1e-13
    assert _rel(orc.hvp_pf(kind, mat, c, el, s, t), Hv.numpy()) < 1e-13
