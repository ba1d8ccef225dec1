"""The C/OpenMP oracle (cpu_baseline implementation) agrees with the NumPy oracle pinned to the reference."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import tatva_oracle as orc


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.mark.parametrize("kind,n", [("tri3", 6), ("tet4", 3), ("hex8", 4)])
def test_c_oracle_matches_numpy_oracle(kind, n):
    rng = np.random.default_rng(3)
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(n, n)
        mat, name, prm = orc.LinearElastic(0.4, 0.6), "linear_elastic", (0.4, 0.6)
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1, 1, 1), (n, n, n))
        mat, name, prm = orc.NeoHookean(500.0, 1000.0), "neo_hookean", (500.0, 1000.0)
    else:
        c, el = orc.mesh_box_hex(n)
        mat, name, prm = orc.NeoHookean(500.0, 1000.0), "neo_hookean", (500.0, 1000.0)
    c = c + 0.1 / n * rng.uniform(-1, 1, c.shape)
    u = 0.02 * rng.normal(size=c.shape)
    v = rng.normal(size=c.shape)
    assert abs(c_oracle.energy(kind, prm, c, el, u, name) - orc.energy(kind, mat, c, el, u)) <= 1e-13 * abs(orc.energy(kind, mat, c, el, u)) + 1e-18
    assert _rel(c_oracle.residual(kind, prm, c, el, u, name), orc.residual(kind, mat, c, el, u)) < 1e-13
    assert _rel(c_oracle.hvp(kind, prm, c, el, u, v, name), orc.hvp(kind, mat, c, el, u, v)) < 1e-13


@pytest.mark.parametrize("kind,n", [("tet4", 3), ("hex8", 3)])
def test_c_oracle_phase_field_matches_numpy_oracle(kind, n):
    """Two-field (u, phi) law of config 5: the C restatement against the NumPy one (itself checked against the
    reference's Operator on the stacked state, `pf_*` fixtures)."""
    rng = np.random.default_rng(4)
    c, el = orc.mesh_box_tet((1, 1, 1), (n, n, n)) if kind == "tet4" else orc.mesh_box_hex(n)
    c = c + 0.1 / n * rng.uniform(-1, 1, c.shape)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    mat = orc.NeoHookeanPhaseField(*prm)
    s = np.concatenate([0.02 * rng.normal(size=(len(c), 3)), rng.uniform(0, 0.8, size=(len(c), 1))], axis=1)
    t = rng.normal(size=s.shape)
    e_ref = orc.energy_pf(kind, mat, c, el, s)
    assert abs(c_oracle.energy_pf(kind, prm, c, el, s) - e_ref) <= 1e-13 * abs(e_ref)
    assert _rel(c_oracle.residual_pf(kind, prm, c, el, s), orc.residual_pf(kind, mat, c, el, s)) < 1e-13
    assert _rel(c_oracle.hvp_pf(kind, prm, c, el, s, t), orc.hvp_pf(kind, mat, c, el, s, t)) < 1e-13
