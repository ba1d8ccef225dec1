"""GPU parity at the BASELINE.json config sizes, directly against the C oracle (oracle/tatva_oracle.c, the OpenMP
restatement of the reference's arithmetic, itself pinned to the NumPy oracle and through it to the reference's outputs).

Tolerance 1e-12 relative (l2 and max-norm), as the north star states.  Sizes: config 1 Tri3 256^2, config 2 Tet4 n = 55
(residual, HVP and the assembled CSR values), config 3 Hex8 128^3 (and 32^3), config 5 compound
(u, phi) Tet4 n = 55.  The C oracle finishes each of them in well under a second per call on the box's host cores.
"""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import tatva_oracle as orc
from test_gpu_parity import _assert_close, _case, _make_op, _material, _rel

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(autouse=True, scope="module")
def _all_host_cores():
    c_oracle.set_num_threads(os.cpu_count() or 1)
    yield


@pytest.mark.parametrize("n", [32, 128])
def test_config3_hex8_neo_hookean_vs_c_oracle(n):
    """Config 3 (Hex8 128^3, 6.44 M DOFs) and the 32^3 size of SURVEY §8(d): energy, residual, HVP."""
    c, el, u, v, (mname, omat) = _case("hex8", n)
    op = _make_op("hex8", c, el)
    mat = _material(mname, omat)
    prm = (omat.mu, omat.lmbda)
    e_ref = c_oracle.energy("hex8", prm, c, el, u)
    assert abs(float(op.energy(mat)(u)) - e_ref) <= RTOL * abs(e_ref)
    _assert_close(op.residual(mat)(u), c_oracle.residual("hex8", prm, c, el, u), RTOL)
    _assert_close(op.hvp(mat)(u, v), c_oracle.hvp("hex8", prm, c, el, u, v), RTOL)


def test_config3_hex8_every_hvp_variant_vs_c_oracle():
    """Every kernel variant kept for measurement computes the same HVP (Hex8 32^3)."""
    c, el, u, v, (mname, omat) = _case("hex8", 32)
    ref = c_oracle.hvp("hex8", (omat.mu, omat.lmbda), c, el, u, v)
    mat = _material(mname, omat)
    op = _make_op("hex8", c, el)
    for variant in op.hvp_variants():
        op.set_variant(variant)
        _assert_close(op.hvp(mat)(u, v), ref, RTOL)
    op.set_variant(0)


def test_config1_tri3_linear_elastic_256_vs_c_oracle():
    """Config 1: Tri3 256 x 256, plane-strain linear elasticity (E = 1, nu = 0.3)."""
    c, el, u, v, (mname, omat) = _case("tri3", 256)
    assert el.shape[0] == 131072 and c.shape[0] == 66049
    op = _make_op("tri3", c, el)
    mat = _material(mname, omat)
    prm = (omat.mu, omat.lmbda)
    e_ref = c_oracle.energy("tri3", prm, c, el, u, "linear_elastic")
    assert abs(float(op.energy(mat)(u)) - e_ref) <= RTOL * abs(e_ref)
    _assert_close(op.residual(mat)(u), c_oracle.residual("tri3", prm, c, el, u, "linear_elastic"), RTOL)
    _assert_close(op.hvp(mat)(u, v), c_oracle.hvp("tri3", prm, c, el, u, v, "linear_elastic"), RTOL)


def test_config2_tet4_residual_hvp_and_csr_values_vs_c_oracle():
    """Config 2: Tet4 box n = 55 (998 250 elements).  Residual and HVP against the C oracle; the assembled CSR `data`
    (23 036 814 values) against the REFERENCE ALGORITHM (sparse/base.py:139-176, :230-270) driven by the C oracle:
    one oracle HVP per colour with a 0/1 seed, then data[k] = J_c[row(k), colors[indices[k]]]."""
    from tatva_b200 import sparse

    c, el, u, v, (mname, omat) = _case("tet4", 55)
    assert el.shape[0] == 998250
    op = _make_op("tet4", c, el)
    mat = _material(mname, omat)
    prm = (omat.mu, omat.lmbda)
    _assert_close(op.residual(mat)(u), c_oracle.residual("tet4", prm, c, el, u), RTOL)
    _assert_close(op.hvp(mat)(u, v), c_oracle.hvp("tet4", prm, c, el, u, v), RTOL)
    pat = sparse.pattern_from_mesh(op.mesh, 3)
    assert pat.nnz == 23036814
    cm = sparse.ColoredMatrix.from_csr(pat)
    colors = np.asarray(cm.colors)
    n = 3 * c.shape[0]
    jvp = lambda seed: c_oracle.hvp("tet4", prm, c, el, u, seed.reshape(-1, 3)).ravel()  # noqa: E731
    ref = orc.colored_jacobian_data(jvp, n, pat.indptr, pat.indices, colors)
    data = sparse.assembler(op, mat, cm)(u).cpu().numpy()
    assert data.shape == ref.shape
    assert _rel(data, ref) <= RTOL
    rows = sparse.assembler(op, mat, cm, by_rows=True)(u).cpu().numpy()  # the bitwise-reproducible kernel
    assert _rel(rows, ref) <= RTOL
    # and the public reference-shaped entry point
    K = sparse.jacfwd(op.residual(mat), cm)(u)
    assert _rel(K.data.cpu().numpy() if isinstance(K.data, torch.Tensor) else K.data, ref) <= RTOL


def test_config5_compound_phase_field_tet4_55_vs_c_oracle():
    """Config 5 at its parity size (n = 55): the compound (u, phi) state, node-interleaved [ux,uy,uz,phi]."""
    from tatva_b200 import materials

    rng = np.random.default_rng(1)
    c, el, u, _, _ = _case("tet4", 55)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    mat = materials.NeoHookeanPhaseField(*prm)
    op = _make_op("tet4", c, el)
    phi = 0.4 + 0.4 * np.sin(2 * np.pi * c[:, 0]) * np.cos(2 * np.pi * c[:, 1])
    s = np.concatenate([u, phi[:, None]], axis=1)
    t = rng.normal(size=s.shape)
    arr, tt = torch.as_tensor(s.ravel(), device="cuda"), torch.as_tensor(t.ravel(), device="cuda")
    e_ref = c_oracle.energy_pf("tet4", prm, c, el, s)
    assert abs(float(op.energy(mat)(arr)) - e_ref) <= RTOL * abs(e_ref)
    _assert_close(op.residual(mat)(arr).reshape(-1, 4), c_oracle.residual_pf("tet4", prm, c, el, s), RTOL)
    _assert_close(op.hvp(mat)(arr, tt).reshape(-1, 4), c_oracle.hvp_pf("tet4", prm, c, el, s, t), RTOL)


@pytest.mark.parametrize("n", [5, 32])
def test_hex8_geometry_cache_variants_vs_c_oracle(n):
    """The Hex8 x neo-Hookean HVP on the plan's geometry cache (tatva_plan_cache_geometry, opt-in) and every
    occupancy / staging point of that kernel (variants 50-56), against the C oracle; the default operator runs the kernel
    that re-derives the geometry; an element sub-range reads the cache through an offset view."""
    import ctypes as C

    from tatva_b200 import _lib

    c, el, u, v, (mname, omat) = _case("hex8", n)
    ref = c_oracle.hvp("hex8", (omat.mu, omat.lmbda), c, el, u, v)
    mat = _material(mname, omat)
    op = _make_op("hex8", c, el, cache_geometry=True)
    _assert_close(op.hvp(mat)(u, v), ref, RTOL)
    assert op._geometry_cached
    for variant in (50, 51, 52, 53, 54, 55, 56):
        op.set_variant(variant)
        _assert_close(op.hvp(mat)(u, v), ref, RTOL)
    op.set_variant(0)
    op0 = _make_op("hex8", c, el)  # the default re-derives the geometry
    _assert_close(op0.hvp(mat)(u, v), ref, RTOL)
    assert not op0._geometry_cached
    # two element sub-ranges accumulate to the full result (what the partitioned operator launches)
    ut, vt = torch.as_tensor(u, device="cuda"), torch.as_tensor(v, device="cuda")
    y = torch.empty_like(ut)
    prm, npar = _lib.params_array(mat.params())
    E, cut = el.shape[0], el.shape[0] // 3 + 1
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(op._L.tatva_hvp_elems(op._plan_fused, mat.material_id, prm, npar, ut.data_ptr(), vt.data_ptr(), y.data_ptr(), 0, cut, 1, st), "tatva_hvp_elems")
    _lib.check(op._L.tatva_hvp_elems(op._plan_fused, mat.material_id, prm, npar, ut.data_ptr(), vt.data_ptr(), y.data_ptr(), cut, E - cut, 0, st), "tatva_hvp_elems")
    _assert_close(y, ref, RTOL)


@pytest.mark.parametrize("kind,n", [("hex8", 7), ("tet4", 9), ("tri3", 33)])
def test_output_views_at_an_8_byte_offset_take_the_plain_memset(kind, n):
    """The fused launches clear y with a kernel that stores 16 bytes at a time and start the element kernel behind it
    (launch_behind_zero); an output that is only 8-byte aligned (a view into a larger array) must fall back to the plain
    memset and give the same result, as must an output that already holds garbage."""
    c, el, u, v, (mname, omat) = _case(kind, n)
    op = _make_op(kind, c, el)
    mat = _material(mname, omat)
    ut, vt = torch.as_tensor(u, device="cuda"), torch.as_tensor(v, device="cuda")
    ref = op._raw_hvp(mat, ut, vt).clone()
    buf = torch.full((ut.numel() + 1,), 7.0, dtype=torch.float64, device="cuda")
    y = buf[1:].view_as(ut)
    assert y.data_ptr() % 16 == 8
    op._raw_hvp(mat, ut, vt, out=y)
    assert torch.equal(buf[:1].cpu(), torch.full((1,), 7.0, dtype=torch.float64))
    _assert_close(y, ref.cpu().numpy(), 1e-14)
    y2 = torch.full_like(ut, float("nan"))
    op._raw_hvp(mat, ut, vt, out=y2)
    _assert_close(y2, ref.cpu().numpy(), 1e-14)


def test_config3_building_blocks_at_128_vs_c_oracle():
    """Operator.grad / its adjoint / integration weights / gather on Hex8 128^3 (the modal kernels and k_gather4 at the
    BASELINE size) against the C oracle's generic formulas (J = dN/dxi X, dN/dX = inv(J) dN/dxi, element/base.py:90-115)."""
    c, el, u, v, _ = _case("hex8", 128)
    op = _make_op("hex8", c, el)
    ut = torch.as_tensor(u, device="cuda")
    g = op._k_grad(ut)
    _assert_close(g, c_oracle.op_grad("hex8", c, el, u), RTOL)
    _assert_close(op.get_integration_weights(), c_oracle.op_integration_weights("hex8", c, el), RTOL)
    gd = np.random.default_rng(5).normal(size=tuple(g.shape))
    del g
    _assert_close(op._k_grad_adj(torch.as_tensor(gd, device="cuda")), c_oracle.op_grad_adjoint("hex8", c, el, gd), RTOL)
    del gd
    assert np.array_equal(op._k_gather(ut).cpu().numpy(), c_oracle.op_gather("hex8", c, el, u))


def test_config2_building_blocks_vs_c_oracle():
    """The same for Tet4 n = 55 (generic staged kernels, batched staged loads in the adjoint)."""
    c, el, u, v, _ = _case("tet4", 55)
    op = _make_op("tet4", c, el)
    ut = torch.as_tensor(u, device="cuda")
    g = op._k_grad(ut)
    _assert_close(g, c_oracle.op_grad("tet4", c, el, u), RTOL)
    gd = np.random.default_rng(5).normal(size=tuple(g.shape))
    _assert_close(op._k_grad_adj(torch.as_tensor(gd, device="cuda")), c_oracle.op_grad_adjoint("tet4", c, el, gd), RTOL)
    _assert_close(op.get_integration_weights(), c_oracle.op_integration_weights("tet4", c, el), RTOL)
