"""The kernels' per-element arithmetic on the CPU.

The element tables, geometry, constitutive laws and the modal / tx-pair functions of the Hex8 kernels are compiled for
host AND device from the same source (csrc/common.cuh, csrc/neo_hookean.cu); `tatva_probe_element` and
`tatva_probe_hex8_nh_modal` run one element through them on the host.  Summing the per-element results over a mesh
must reproduce the oracle's energy / residual / HVP / Hessian diagonal, which checks the formulas the GPU executes
without a GPU (gather, staging and scatter — pure data movement — are covered by the `-m gpu` parity tests)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import tatva_oracle as orc
from tatva_b200 import _lib

KIND = {"tri3": _lib.TRI3, "tet4": _lib.TET4, "hex8": _lib.HEX8, "quad4": _lib.QUAD4, "tri6": _lib.TRI6, "quad8": _lib.QUAD8}
f64 = lambda a: a.ctypes.data_as(_lib.c_f64p)  # noqa: E731


def _mesh(kind, rng):
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(4, 3)
    elif kind == "quad4":
        c, el = orc.mesh_unit_square_quad(3, 3)
    elif kind in ("tri6", "quad8"):
        c, el = orc.mesh_second_order(kind, 2, 2)
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1, 1, 1), (2, 2, 2))
    else:
        c, el = orc.mesh_box_hex(3)
    return c + 0.03 * rng.uniform(-1, 1, c.shape), el


def _assemble(kind, material_id, params, mode, c, el, u, v, dpn):
    """Sum of the probe's per-element results (what the kernel's scatter does)."""
    L = _lib.lib()
    prm, n_prm = _lib.params_array(params)
    npe = el.shape[1]
    out = np.zeros((c.shape[0], dpn)) if mode else 0.0
    buf = np.zeros(npe * dpn if mode else 1)
    for e in el:
        X, ue = np.ascontiguousarray(c[e]), np.ascontiguousarray(u[e])
        ve = np.ascontiguousarray(v[e]) if v is not None else None
        rc = L.tatva_probe_element(KIND[kind], material_id, prm, n_prm, mode, f64(X), f64(ue), f64(ve) if ve is not None else None, f64(buf))
        assert rc == 0, rc
        if mode:
            np.add.at(out, e, buf.reshape(npe, dpn))
        else:
            out += buf[0]
    return out


@pytest.mark.parametrize("kind,law", [("tri3", "le"), ("quad4", "le"), ("tri6", "le"), ("quad8", "le"), ("tet4", "le"), ("hex8", "le"), ("tet4", "nh"), ("hex8", "nh")])
def test_generic_element_body_matches_the_oracle(kind, law):
    rng = np.random.default_rng(0)
    c, el = _mesh(kind, rng)
    dim = c.shape[1]
    u, v = 0.02 * rng.normal(size=c.shape), rng.normal(size=c.shape)
    if law == "le":
        mid, prm, omat = _lib.LINEAR_ELASTIC, (0.38, 0.58), orc.LinearElastic(0.38, 0.58)
    else:
        mid, prm, omat = _lib.NEO_HOOKEAN, (500.0, 1000.0), orc.NeoHookean(500.0, 1000.0)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 0, c, el, u, None, dim), orc.energy(kind, omat, c, el, u), rtol=1e-13)
    r_ref = orc.residual(kind, omat, c, el, u)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 1, c, el, u, None, dim), r_ref, rtol=1e-11, atol=1e-13 * np.abs(r_ref).max())
    h_ref = orc.hvp(kind, omat, c, el, u, v)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 2, c, el, u, v, dim), h_ref, rtol=1e-11, atol=1e-13 * np.abs(h_ref).max())
    ip, ix = orc.pattern_from_mesh(el, len(c), dim)
    diag = sps.csr_matrix((orc.assemble_csr_data(kind, omat, c, el, u, ip, ix), ix, ip)).diagonal().reshape(-1, dim)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 3, c, el, u, None, dim), diag, rtol=1e-11)
    # r02: the rank-structured diagonal of k_hessian_diag_rank (both laws here have a RankLaw)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 4, c, el, u, None, dim), diag, rtol=1e-11)


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_phase_field_element_body_matches_the_oracle(kind):
    rng = np.random.default_rng(1)
    c, el = _mesh(kind, rng)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    omat = orc.NeoHookeanPhaseField(*prm)
    s = np.concatenate([0.02 * rng.normal(size=c.shape), rng.uniform(0, 0.8, size=(len(c), 1))], axis=1)
    t = rng.normal(size=s.shape)
    mid = _lib.NEO_HOOKEAN_PHASE_FIELD
    np.testing.assert_allclose(_assemble(kind, mid, prm, 0, c, el, s, None, 4), orc.energy_pf(kind, omat, c, el, s), rtol=1e-13)
    r_ref = orc.residual_pf(kind, omat, c, el, s)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 1, c, el, s, None, 4), r_ref, rtol=1e-11, atol=1e-13 * np.abs(r_ref).max())
    h_ref = orc.hvp_pf(kind, omat, c, el, s, t)
    np.testing.assert_allclose(_assemble(kind, mid, prm, 2, c, el, s, t, 4), h_ref, rtol=1e-11, atol=1e-13 * np.abs(h_ref).max())


def test_hex8_pair_kernels_arithmetic_matches_the_oracle():
    """The headline path: raw modal coefficients, tx-pair sharing, reference-space tangent, folded 1/512 scalings."""
    L = _lib.lib()
    rng = np.random.default_rng(2)
    c, el = orc.mesh_box_hex(4)
    c = c + 0.1 / 4 * rng.uniform(-1, 1, c.shape)
    t = 2 * np.pi
    u = 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 2]), np.sin(t * c[:, 2]) * np.cos(t * c[:, 0])], -1)
    v = rng.normal(size=c.shape)
    omat = orc.NeoHookean(500.0, 1000.0)
    energy, res, hv = 0.0, np.zeros_like(c), np.zeros_like(c)
    buf, e1 = np.zeros(24), np.zeros(1)
    for e in el:
        X, ue, ve = (np.ascontiguousarray(a[e]) for a in (c, u, v))
        assert L.tatva_probe_hex8_nh_modal(0, f64(X), f64(ue), None, 500.0, 1000.0, f64(e1)) == 0
        energy += e1[0]
        assert L.tatva_probe_hex8_nh_modal(1, f64(X), f64(ue), None, 500.0, 1000.0, f64(buf)) == 0
        np.add.at(res, e, buf.reshape(8, 3))
        assert L.tatva_probe_hex8_nh_modal(2, f64(X), f64(ue), f64(ve), 500.0, 1000.0, f64(buf)) == 0
        np.add.at(hv, e, buf.reshape(8, 3))
    np.testing.assert_allclose(energy, orc.energy("hex8", omat, c, el, u), rtol=1e-13)
    r_ref, h_ref = orc.residual("hex8", omat, c, el, u), orc.hvp("hex8", omat, c, el, u, v)
    np.testing.assert_allclose(res, r_ref, rtol=1e-10, atol=1e-13 * np.abs(r_ref).max())
    np.testing.assert_allclose(hv, h_ref, rtol=1e-10, atol=1e-13 * np.abs(h_ref).max())
    # r02: the HVP through the geometry-cache arithmetic, Operator.grad and the weights in modal form (k_hex8_nh_hvp_geo,
    # k_hex8_grad_modal, k_hex8_weights_modal)
    hv3, g, w = np.zeros_like(c), np.zeros((len(el), 8, 3, 3)), np.zeros((len(el), 8))
    gbuf, wbuf = np.zeros(72), np.zeros(8)
    for k, e in enumerate(el):
        X, ue, ve = (np.ascontiguousarray(a[e]) for a in (c, u, v))
        assert L.tatva_probe_hex8_nh_modal(3, f64(X), f64(ue), f64(ve), 500.0, 1000.0, f64(buf)) == 0
        np.add.at(hv3, e, buf.reshape(8, 3))
        assert L.tatva_probe_hex8_nh_modal(4, f64(X), f64(ue), None, 500.0, 1000.0, f64(gbuf)) == 0
        g[k] = gbuf.reshape(8, 3, 3)
        assert L.tatva_probe_hex8_nh_modal(5, f64(X), f64(ue), None, 500.0, 1000.0, f64(wbuf)) == 0
        w[k] = wbuf
    np.testing.assert_allclose(hv3, h_ref, rtol=1e-10, atol=1e-13 * np.abs(h_ref).max())
    g_ref = orc.op_grad("hex8", c, el, u)
    np.testing.assert_allclose(g, g_ref, rtol=1e-11, atol=1e-13 * np.abs(g_ref).max())
    np.testing.assert_allclose(w, orc.op_integration_weights("hex8", c, el), rtol=1e-13)


def test_tet4_reference_space_kernels_arithmetic_matches_the_oracle():
    """The default Tet4 x neo-Hookean kernels (one point, J = edge vectors, f_0 = -(f_1 + f_2 + f_3))."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    c, el = orc.mesh_box_tet((1, 1, 1), (3, 3, 3))
    c = c + 0.02 * rng.uniform(-1, 1, c.shape)
    u, v = 0.02 * rng.normal(size=c.shape), rng.normal(size=c.shape)
    omat = orc.NeoHookean(500.0, 1000.0)
    res, hv, buf = np.zeros_like(c), np.zeros_like(c), np.zeros(12)
    for e in el:
        X, ue, ve = (np.ascontiguousarray(a[e]) for a in (c, u, v))
        assert L.tatva_probe_tet4_nh_ref(1, f64(X), f64(ue), None, 500.0, 1000.0, f64(buf)) == 0
        np.add.at(res, e, buf.reshape(4, 3))
        assert L.tatva_probe_tet4_nh_ref(2, f64(X), f64(ue), f64(ve), 500.0, 1000.0, f64(buf)) == 0
        np.add.at(hv, e, buf.reshape(4, 3))
    r_ref, h_ref = orc.residual("tet4", omat, c, el, u), orc.hvp("tet4", omat, c, el, u, v)
    np.testing.assert_allclose(res, r_ref, rtol=1e-10, atol=1e-13 * np.abs(r_ref).max())
    np.testing.assert_allclose(hv, h_ref, rtol=1e-10, atol=1e-13 * np.abs(h_ref).max())


def test_probes_reject_bad_arguments():
    L = _lib.lib()
    z = np.zeros(24)
    assert L.tatva_probe_hex8_nh_modal(2, f64(z), f64(z), None, 1.0, 1.0, f64(z)) != 0  # HVP needs v
    assert L.tatva_probe_hex8_nh_modal(7, f64(z), f64(z), f64(z), 1.0, 1.0, f64(z)) != 0
    prm, n = _lib.params_array((1.0, 1.0))
    assert L.tatva_probe_element(_lib.TRI3, _lib.NEO_HOOKEAN, prm, n, 1, f64(z), f64(z), None, f64(z)) != 0  # no such pair
