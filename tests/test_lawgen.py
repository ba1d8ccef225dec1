"""Source-to-source AD of user energy densities (tatva_b200/lawgen.py): the generated psi / first / second programs,
compiled as plain C on the host, against SymPy's own symbolic derivatives of the same density (and, for the
neo-Hookean density of the reference's tests, against the oracle's closed forms)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from oracle import tatva_oracle as orc
from tatva_b200 import lawgen

sp = pytest.importorskip("sympy")


def neo_hookean(G, mu, lmbda):  # reference tests/test_sparse_tracer.py:103-115
    F = sp.eye(3) + G
    lnJ = sp.log(F.det())
    return mu / 2 * ((F.T * F).trace() - 3 - 2 * lnJ) + lmbda / 2 * lnJ**2


def mooney_rivlin(G, c1, c2, kappa):
    F = sp.eye(3) + G
    Cm = F.T * F
    J = F.det()
    I1 = Cm.trace()
    I2 = (I1**2 - (Cm * Cm).trace()) / 2
    return c1 * (J ** sp.Rational(-2, 3) * I1 - 3) + c2 * (J ** sp.Rational(-4, 3) * I2 - 3) + kappa / 2 * (J - 1) ** 2


def st_venant(G, mu, lmbda):
    F = sp.eye(3) + G
    E = (F.T * F - sp.eye(3)) / 2
    return lmbda / 2 * E.trace() ** 2 + mu * (E * E).trace()


def linear_elastic_2d(G, mu, lmbda):  # reference tests/test_sparse.py:20-38
    eps = (G + G.T) / 2
    return mu * (eps * eps).trace() + lmbda / 2 * eps.trace() ** 2


def _compile(law, tmp_path):
    src = tmp_path / "law.c"
    src.write_text(law.c_source())
    lib = tmp_path / "law.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(lib), str(src), "-lm"], check=True)
    L = C.CDLL(str(lib))
    L.law_psi.restype = C.c_double
    return L


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("psi,nprm,dim,prm", [(neo_hookean, 2, 3, (500.0, 1000.0)), (mooney_rivlin, 3, 3, (120.0, 30.0, 900.0)), (st_venant, 2, 3, (80.0, 120.0)), (linear_elastic_2d, 2, 2, (0.4, 0.6))])
def test_generated_law_matches_symbolic_derivatives(psi, nprm, dim, prm, tmp_path):
    law = lawgen.generate(psi, nprm, dim=dim)
    L = _compile(law, tmp_path)
    n = dim * dim
    Gs = sp.Matrix(dim, dim, lambda i, j: sp.Symbol(f"G{i}{j}", real=True))
    ps = [sp.Symbol(f"p{k}", real=True) for k in range(nprm)]
    expr = psi(Gs, *ps)
    flat = list(Gs)
    f_psi = sp.lambdify(flat + ps, expr, "numpy")
    f_P = sp.lambdify(flat + ps, [sp.diff(expr, g) for g in flat], "numpy")
    f_H = sp.lambdify(flat + ps, sp.hessian(expr, flat), "numpy")
    rng = np.random.default_rng(0)
    p = np.array(prm)
    for _ in range(5):
        G = 0.1 * rng.normal(size=n)
        dG = rng.normal(size=n)
        out = np.empty(n)
        assert abs(L.law_psi(_ptr(G), _ptr(p)) - f_psi(*G, *p)) <= 1e-13 * max(1.0, abs(f_psi(*G, *p)))
        L.law_first(_ptr(G), _ptr(p), _ptr(out))
        P = np.array(f_P(*G, *p), dtype=float)
        assert np.abs(out - P).max() <= 1e-12 * np.abs(P).max()
        L.law_second(_ptr(G), _ptr(dG), _ptr(p), _ptr(out))
        dP = np.array(f_H(*G, *p), dtype=float) @ dG
        assert np.abs(out - dP).max() <= 1e-12 * np.abs(dP).max()
    counts = law.op_counts()
    assert counts["psi"] <= counts["first"] <= counts["second"]


def test_generated_neo_hookean_equals_the_oracle_closed_forms(tmp_path):
    law = lawgen.generate(neo_hookean, 2)
    L = _compile(law, tmp_path)
    omat = orc.NeoHookean(500.0, 1000.0)
    rng = np.random.default_rng(1)
    G, dG, p, out = 0.1 * rng.normal(size=(3, 3)), rng.normal(size=(3, 3)), np.array([500.0, 1000.0]), np.empty(9)
    L.law_first(_ptr(G), _ptr(p), _ptr(out))
    assert np.abs(out.reshape(3, 3) - omat.P(G)).max() <= 1e-12 * np.abs(omat.P(G)).max()
    L.law_second(_ptr(G), _ptr(dG), _ptr(p), _ptr(out))
    assert np.abs(out.reshape(3, 3) - omat.dP(G, dG)).max() <= 1e-12 * np.abs(omat.dP(G, dG)).max()
    # forward-over-reverse stays within a small multiple of the density's own cost (no expression swell)
    c = law.op_counts()
    assert c["second"] <= 14 * c["psi"], c


def test_two_field_density_with_values(tmp_path):
    """A density that also reads nodal VALUES (phase-field style): inputs G (4x3) and val (4)."""

    def psi(G, V, mu, lmbda, Gc, ell, k):
        F = sp.eye(3) + G[:3, :]
        lnJ = sp.log(F.det())
        nh = mu / 2 * ((F.T * F).trace() - 3 - 2 * lnJ) + lmbda / 2 * lnJ**2
        phi, gphi = V[3], G[3, :]
        return ((1 - phi) ** 2 + k) * nh + Gc * (phi**2 / (2 * ell) + ell / 2 * (gphi * gphi.T)[0, 0])

    law = lawgen.generate(psi, 5, dim=3, dofs_per_node=4, uses_values=True)
    L = _compile(law, tmp_path)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    omat = orc.NeoHookeanPhaseField(*prm)
    rng = np.random.default_rng(2)
    G = 0.1 * rng.normal(size=(4, 3))
    V = np.array([0.0, 0.0, 0.0, 0.3])
    dGv, dV = rng.normal(size=(4, 3)), np.array([0.0, 0.0, 0.0, 0.7])
    inp, dinp, p, out = np.concatenate([G.ravel(), V]), np.concatenate([dGv.ravel(), dV]), np.array(prm), np.empty(16)
    assert abs(L.law_psi(_ptr(inp), _ptr(p)) - omat.psi(G[:3], V[3], G[3])) <= 1e-12 * abs(omat.psi(G[:3], V[3], G[3]))
    L.law_first(_ptr(inp), _ptr(p), _ptr(out))
    A, b, c = omat.first(G[:3], np.float64(V[3]), G[3])
    assert np.abs(out[:9].reshape(3, 3) - A).max() <= 1e-12 * np.abs(A).max() and abs(out[15] - b) <= 1e-12 * abs(b) and np.abs(out[9:12] - c).max() <= 1e-12
    L.law_second(_ptr(inp), _ptr(dinp), _ptr(p), _ptr(out))
    A, b, c = omat.second(G[:3], np.float64(V[3]), G[3], dGv[:3], np.float64(dV[3]), dGv[3])
    assert np.abs(out[:9].reshape(3, 3) - A).max() <= 1e-11 * np.abs(A).max() and abs(out[15] - b) <= 1e-11 * abs(b) and np.abs(out[9:12] - c).max() <= 1e-12
