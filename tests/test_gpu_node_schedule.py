"""Warp-cooperative residual / HVP kernels (`Operator(node_schedule=True)`, k_fused_wc): the transpose of the reference's
gather `v[self.mesh.elements]` (tatva/operator.py:221) done node-wise — one nodal-row load per distinct node of a warp,
one atomic add per distinct node of a 128-element tile.  Same results as the element-per-thread kernels and the oracle
at 1e-12, on the generator's element order, on a shuffled mesh (most rows then take the direct path) and at config size.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import tatva_oracle as orc
from test_gpu_parity import _assert_close, _case, _make_op, _material

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.mark.parametrize("kind,n", [("tri3", 8), ("tri3", 67), ("tet4", 3), ("tet4", 13)])
@pytest.mark.parametrize("shuffle", [False, True])
def test_node_schedule_kernels_match_the_oracle(kind, n, shuffle):
    c, el, u, v, (mname, omat) = _case(kind, n)
    if shuffle:
        el = el[np.random.default_rng(5).permutation(len(el))]
    op = _make_op(kind, c, el, node_schedule=True)
    assert op.node_schedule_stats["distinct_per_warp"] <= 32 and op._node_schedule is not None
    mat = _material(mname, omat)
    law = "linear_elastic" if kind == "tri3" else "neo_hookean"
    prm = (omat.mu, omat.lmbda)
    _assert_close(op.residual(mat)(u), c_oracle.residual(kind, prm, c, el, u, law), RTOL)
    _assert_close(op.hvp(mat)(u, v), c_oracle.hvp(kind, prm, c, el, u, v, law), RTOL)
    # the element-per-thread kernel on the same plan (variant 31 ignores the schedule)
    ref = op.hvp(mat)(u, v).clone()
    op.set_variant(31)
    _assert_close(op.hvp(mat)(u, v), ref.cpu().numpy(), RTOL)
    for variant in (32, 33, 34):  # A/B variants of the Tet4 kernel: occupancy, element's own gather, per-warp scatter
        op.set_variant(variant)
        _assert_close(op.hvp(mat)(u, v), ref.cpu().numpy(), RTOL)
        _assert_close(op.residual(mat)(u), c_oracle.residual(kind, prm, c, el, u, law), RTOL)
    op.set_variant(0)


@pytest.mark.parametrize("cap", [True, 1, 3])
def test_node_schedule_tet4_linear_elastic_and_contributor_caps(cap):
    from tatva_b200 import materials

    c, el, u, v, _ = _case("tet4", 7)
    op, op0 = _make_op("tet4", c, el, node_schedule=cap), _make_op("tet4", c, el, node_schedule=False)
    mat = materials.LinearElastic(0.38, 0.58)
    _assert_close(op.hvp(mat)(u, v), op0.hvp(mat)(u, v).cpu().numpy(), RTOL)
    _assert_close(op.residual(mat)(u), op0.residual(mat)(u).cpu().numpy(), RTOL)


def test_node_schedule_compound_phase_field_config5():
    """Config 5 at its parity size (n = 55): node-interleaved [ux, uy, uz, phi]; the default two-field kernel is the
    node-schedule one (shuffle gather + tile sums), variant 38 keeps the element's own gather, 31 is element-per-thread."""
    from tatva_b200 import materials

    rng = np.random.default_rng(1)
    c, el, u, _, _ = _case("tet4", 55)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    mat = materials.NeoHookeanPhaseField(*prm)
    op = _make_op("tet4", c, el)
    assert op._node_schedule is not None
    phi = 0.4 + 0.4 * np.sin(2 * np.pi * c[:, 0]) * np.cos(2 * np.pi * c[:, 1])
    s = np.concatenate([u, phi[:, None]], axis=1)
    t = rng.normal(size=s.shape)
    arr, tt = torch.as_tensor(s.ravel(), device="cuda"), torch.as_tensor(t.ravel(), device="cuda")
    ref_r, ref_h = c_oracle.residual_pf("tet4", prm, c, el, s), c_oracle.hvp_pf("tet4", prm, c, el, s, t)
    for variant in (0, 38, 31):
        op.set_variant(variant)
        _assert_close(op.residual(mat)(arr).reshape(-1, 4), ref_r, RTOL)
        _assert_close(op.hvp(mat)(arr, tt).reshape(-1, 4), ref_h, RTOL)


def test_node_schedule_config2_and_config1_sizes():
    for kind, n, law in (("tet4", 55, "neo_hookean"), ("tri3", 256, "linear_elastic")):
        c, el, u, v, (mname, omat) = _case(kind, n)
        op = _make_op(kind, c, el, node_schedule=True)
        mat = _material(mname, omat)
        prm = (omat.mu, omat.lmbda)
        _assert_close(op.residual(mat)(u), c_oracle.residual(kind, prm, c, el, u, law), RTOL)
        _assert_close(op.hvp(mat)(u, v), c_oracle.hvp(kind, prm, c, el, u, v, law), RTOL)


def test_node_schedule_auto_keeps_it_only_with_locality():
    """The default ("auto"): a Tet4 operator on the generator's element order carries the schedule, a shuffled element
    list (one reference per distinct node of a tile) drops it, other elements never build it."""
    c, el, u, v, (mname, omat) = _case("tet4", 9)
    assert _make_op("tet4", c, el)._node_schedule is not None
    op = _make_op("tet4", c, el[np.random.default_rng(0).permutation(len(el))])
    assert op._node_schedule is None and op.node_schedule_stats["references_per_entry"] < 3.0
    mat = _material(mname, omat)
    el2 = el[np.random.default_rng(0).permutation(len(el))]
    _assert_close(op.hvp(mat)(u, v), c_oracle.hvp("tet4", (omat.mu, omat.lmbda), c, el2, u, v, "neo_hookean"), RTOL)
    ch, eh, *_ = _case("hex8", 4)
    assert _make_op("hex8", ch, eh)._node_schedule is None


@pytest.mark.parametrize("law", ["nh", "pf"])
def test_element_sub_ranges_keep_the_schedule_on_tile_boundaries(law):
    """tatva_hvp_elems / tatva_residual_elems (what the partitioned operator launches): a range that starts on a multiple
    of 128 elements runs the node-schedule kernel through offset views (ragged end included), any other start falls back
    to the element-per-thread kernel; the pieces accumulate to the full result either way."""
    from tatva_b200 import _lib, materials

    c, el, u, v, (mname, omat) = _case("tet4", 6)  # 1296 elements: 10 full tiles + 16
    op = _make_op("tet4", c, el, node_schedule=True)
    if law == "pf":
        mat = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.1, 1e-6)
        rng = np.random.default_rng(2)
        u = np.concatenate([u, 0.4 + 0.3 * rng.uniform(size=(len(c), 1))], axis=1)
        v = rng.normal(size=u.shape)
    else:
        mat = _material(mname, omat)
    ut, vt = torch.as_tensor(u, device="cuda"), torch.as_tensor(v, device="cuda")
    ref_h, ref_r = op._raw_hvp(mat, ut, vt).clone(), op._raw_residual(mat, ut).clone()
    prm, npar = _lib.params_array(mat.params())
    E = el.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    for cuts in ([0, 384, E], [0, 200, 512, 1100, E], [0, 128, 256, E]):
        y, r = torch.full_like(ut, 3.0), torch.full_like(ut, 3.0)
        for k in range(len(cuts) - 1):
            b, n = cuts[k], cuts[k + 1] - cuts[k]
            _lib.check(op._L.tatva_hvp_elems(op._plan_fused, mat.material_id, prm, npar, ut.data_ptr(), vt.data_ptr(), y.data_ptr(), b, n, int(k == 0), st), "tatva_hvp_elems")
            _lib.check(op._L.tatva_residual_elems(op._plan_fused, mat.material_id, prm, npar, ut.data_ptr(), r.data_ptr(), b, n, int(k == 0), st), "tatva_residual_elems")
        _assert_close(y, ref_h.cpu().numpy(), RTOL)
        _assert_close(r, ref_r.cpu().numpy(), RTOL)
