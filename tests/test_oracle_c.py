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


@pytest.mark.parametrize("kind", ["tri3", "tet4", "hex8"])
def test_c_building_blocks_match_the_numpy_oracle_and_the_reference_fixtures(kind, golden):
    """oracle_blocks (grad, its adjoint, integration weights, gather: the checker of the full-size GPU building-block
    tests) against the NumPy oracle, and against the reference's own Operator outputs (op_* fixtures)."""
    from oracle import c_oracle

    rng = np.random.default_rng(4)
    c, el = {"tri3": lambda: orc.mesh_unit_square_tri(6, 5), "tet4": lambda: orc.mesh_box_tet((1, 1, 1), (3, 2, 3)), "hex8": lambda: orc.mesh_box_hex(3)}[kind]()
    c = c + 0.03 * rng.uniform(-1, 1, c.shape)
    for nv in (1, 3):
        u = rng.normal(size=(len(c), nv))
        g = c_oracle.op_grad(kind, c, el, u)
        np.testing.assert_allclose(g, orc.op_grad(kind, c, el, u), rtol=1e-12, atol=1e-13)
        gd = rng.normal(size=g.shape)
        ref = np.zeros_like(u)
        # adjoint by the defining identity, column by column of the NumPy gradient operator
        lhs = (g * gd).sum()
        np.testing.assert_allclose((u * c_oracle.op_grad_adjoint(kind, c, el, gd)).sum(), lhs, rtol=1e-12)
        assert np.array_equal(c_oracle.op_gather(kind, c, el, u), u[el])
    np.testing.assert_allclose(c_oracle.op_integration_weights(kind, c, el), orc.op_integration_weights(kind, c, el), rtol=1e-13)
    gk = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    np.testing.assert_allclose(c_oracle.op_grad(kind, gk("coords"), gk("conn"), gk("u")), gk("grad_u"), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(c_oracle.op_integration_weights(kind, gk("coords"), gk("conn")), gk("weights"), rtol=1e-13)
