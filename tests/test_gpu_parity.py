"""GPU parity tests: CUDA kernels (through the C ABI) against the CPU oracle and the reference fixtures.

Tolerance: <= 1e-12 relative (l2 and max-norm) for FP64 results, as BASELINE.json's north star states.
"""
import numpy as np
import pytest
import torch

from oracle import tatva_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    l2 = np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)
    mx = np.abs(a - b).max() / (np.abs(b).max() if np.abs(b).max() > 0 else 1.0)
    return max(l2, mx)


def _assert_close(a, b, tol=RTOL):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    assert a.shape == np.asarray(b).shape, (a.shape, np.asarray(b).shape)
    err = _rel(a, b)
    assert err <= tol, f"relative error {err:.3e} > {tol:.1e}"


def _tb():
    import tatva_b200
    from tatva_b200 import element, materials

    return tatva_b200, element, materials


ELEMS = {"tri3": "Tri3", "tet4": "Tetrahedron4", "hex8": "Hexahedron8"}


def _make_op(kind, coords, conn, **kw):
    tb, element, _ = _tb()
    return tb.Operator(tb.Mesh(coords=coords, elements=conn), getattr(element, ELEMS[kind])(), **kw)


def _smooth_u(x):
    if x.shape[1] == 2:
        return 0.05 * np.stack([np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]), np.sin(2 * np.pi * x[:, 1]) * np.cos(2 * np.pi * x[:, 0])], -1)
    return 0.05 * np.stack(
        [
            np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]),
            np.sin(2 * np.pi * x[:, 1]) * np.cos(2 * np.pi * x[:, 2]),
            np.sin(2 * np.pi * x[:, 2]) * np.cos(2 * np.pi * x[:, 0]),
        ],
        -1,
    )


def _case(kind, n, seed=0):
    """Synthetic inputs of SURVEY.md §8(d): jittered structured mesh, smooth u, Gaussian v."""
    rng = np.random.default_rng(seed)
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(n, n)
        mat = ("LinearElastic", orc.LinearElastic(*orc.lame_from_youngs_poisson_2d(1.0, 0.3)))
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1.0, 1.0, 1.0), (n, n, n))
        c = c + np.array([0.5, 0.5, 0.0])
        mat = ("NeoHookean", orc.NeoHookean(500.0, 1000.0))
    else:
        c, el = orc.mesh_box_hex(n)
        mat = ("NeoHookean", orc.NeoHookean(500.0, 1000.0))
    c = c + 0.1 * (1.0 / n) * rng.uniform(-1, 1, c.shape)
    u = _smooth_u(c)
    v = np.random.default_rng(1).normal(size=c.shape)
    return c, el, u, v, mat


def _material(name, omat):
    _, _, materials = _tb()
    return getattr(materials, name)(omat.mu, omat.lmbda)


# ---- building blocks against the reference's own outputs --------------------------------------


@pytest.mark.parametrize("kind", ["tri3", "tet4", "hex8"])
@pytest.mark.parametrize("cache_weights", [False, True])
def test_operator_blocks_match_reference_fixtures(golden, kind, cache_weights):
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    op = _make_op(kind, g("coords"), g("conn"), cache_weights=cache_weights)
    _assert_close(op.grad(g("u")), g("grad_u"))
    _assert_close(op.grad(g("s")), g("grad_s"))
    _assert_close(op.eval(g("u")), g("eval_u"))
    _assert_close(op.eval(g("s")), g("eval_s"))
    _assert_close(op.get_integration_weights(), g("weights"))
    _assert_close(op.integrate(g("s")), g("int_nodal_s"))
    _assert_close(op.integrate_per_element(g("u")), g("int_nodal_u_per_el"))
    _assert_close(op.integrate_per_element(g("quadvals")), g("int_quad_per_el"))


def test_operator_known_answers():
    """reference tests/test_operator.py:14-30, :113-143."""
    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    el = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    op = _make_op("tri3", nodes, el)
    _assert_close(op.eval(np.array([0.0, 1.0, 2.0, 3.0])), np.array([[1.0], [5.0 / 3.0]]), 1e-15)
    _assert_close(op.grad(nodes @ np.array([2.0, 3.0])), np.array([[[2.0, 3.0]], [[2.0, 3.0]]]), 1e-15)
    _assert_close(op.integrate_per_element(np.ones(4)), np.array([0.5, 0.5]), 1e-15)
    _assert_close(op.integrate(np.ones(4)), np.array(1.0), 1e-15)
    _assert_close(op.integrate_per_element(np.full((2, 1), 4.0)), np.array([2.0, 2.0]), 1e-15)
    _assert_close(op.integrate(3.0), np.array(3.0), 1e-15)


def test_map_matches_manual_loop():
    """reference tests/test_operator.py:52-110."""
    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    el = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    op = _make_op("tri3", nodes, el)
    vals = np.arange(4, dtype=np.float64)
    res = op.map(lambda xi, ev: ev.sum() + xi.sum())(vals)
    exp = np.array([[vals[e].sum() + 2.0 / 3] for e in el])
    _assert_close(res, exp, 1e-15)
    bias = np.array([10.0, 20.0])
    res = op.map(lambda xi, ev, b: ev.sum() + b + xi[0], element_quantity=(1,))(vals, bias)
    exp = np.array([[vals[e].sum() + bias[i] + 1.0 / 3] for i, e in enumerate(el)])
    _assert_close(res, exp, 1e-15)
    res = op.map_over_elements(lambda ev: ev.sum())(vals)
    _assert_close(res, np.array([vals[e].sum() for e in el]), 1e-15)


# ---- fused energy / residual / HVP against the oracle ------------------------------------------


@pytest.mark.parametrize("kind,n", [("tri3", 8), ("tri3", 64), ("tet4", 6), ("tet4", 12), ("hex8", 8), ("hex8", 16)])
def test_fused_energy_residual_hvp(kind, n):
    c, el, u, v, (mname, omat) = _case(kind, n)
    op = _make_op(kind, c, el)
    mat = _material(mname, omat)
    _assert_close(op.energy(mat)(u), orc.energy(kind, omat, c, el, u))
    _assert_close(op.residual(mat)(u), orc.residual(kind, omat, c, el, u))
    _assert_close(op.hvp(mat)(u, v), orc.hvp(kind, omat, c, el, u, v))


@pytest.mark.parametrize("variant", [1, 2])
def test_hex8_hvp_variants_agree_with_oracle(variant):
    c, el, u, v, (mname, omat) = _case("hex8", 12)
    op = _make_op("hex8", c, el)
    op.set_variant(variant)
    _assert_close(op.hvp(_material(mname, omat))(u, v), orc.hvp("hex8", omat, c, el, u, v))


def test_hex8_linear_elastic_3d():
    c, el, u, v, _ = _case("hex8", 6)
    omat = orc.LinearElastic(0.7, 1.3)
    _, _, materials = _tb()
    mat = materials.LinearElastic(0.7, 1.3)
    op = _make_op("hex8", c, el)
    _assert_close(op.energy(mat)(u), orc.energy("hex8", omat, c, el, u))
    _assert_close(op.residual(mat)(u), orc.residual("hex8", omat, c, el, u))
    _assert_close(op.hvp(mat)(u, v), orc.hvp("hex8", omat, c, el, u, v))


def test_hvp_is_linear_and_symmetric_at_scale():
    """Size-independent properties on a mesh too large for the NumPy oracle (Hex8 48^3):
    H(a v + b w) = a H v + b H w and <w, H v> = <v, H w>."""
    c, el, u, v, (mname, omat) = _case("hex8", 48)
    op = _make_op("hex8", c, el)
    H = op.hvp(_material(mname, omat))
    ut = torch.as_tensor(u, device="cuda")
    vt = torch.as_tensor(v, device="cuda")
    wt = torch.as_tensor(np.random.default_rng(5).normal(size=v.shape), device="cuda")
    Hv, Hw = H(ut, vt), H(ut, wt)
    comb = H(ut, 0.3 * vt - 1.7 * wt)
    assert _rel(comb.cpu().numpy(), (0.3 * Hv - 1.7 * Hw).cpu().numpy()) < 1e-12
    a, b = float((wt * Hv).sum()), float((vt * Hw).sum())
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))


# ---- autograd route: user energy on the building blocks == fused kernels ------------------------


def test_user_energy_autograd_matches_fused_kernels():
    c, el, u, v, (mname, omat) = _case("hex8", 6)
    op = _make_op("hex8", c, el)
    mat = _material(mname, omat)
    ut = torch.as_tensor(u, device="cuda").requires_grad_(True)
    vt = torch.as_tensor(v, device="cuda")

    def total_energy(uu):  # reference tests/test_sparse_tracer.py:139-144 written in torch
        G = op.grad(uu)
        F = torch.eye(3, dtype=G.dtype, device=G.device) + G
        lnJ = torch.log(torch.linalg.det(F))
        I1 = (F * F).sum(dim=(-1, -2))
        psi = 0.5 * omat.mu * (I1 - 3 - 2 * lnJ) + 0.5 * omat.lmbda * lnJ**2
        return op.integrate(psi)

    E = total_energy(ut)
    _assert_close(E, orc.energy("hex8", omat, c, el, u))
    (r,) = torch.autograd.grad(E, ut, create_graph=True)
    _assert_close(r, orc.residual("hex8", omat, c, el, u))
    (Hv,) = torch.autograd.grad(r, ut, vt)
    _assert_close(Hv, orc.hvp("hex8", omat, c, el, u, v), 1e-11)
    # fused chain through autograd
    ut2 = torch.as_tensor(u, device="cuda").requires_grad_(True)
    E2 = op.energy(mat)(ut2)
    (r2,) = torch.autograd.grad(E2, ut2, create_graph=True)
    (Hv2,) = torch.autograd.grad(r2, ut2, vt)
    _assert_close(r2, orc.residual("hex8", omat, c, el, u))
    _assert_close(Hv2, orc.hvp("hex8", omat, c, el, u, v))


def test_sorted_elements_give_the_same_fused_results_on_a_shuffled_mesh():
    """Operator(sort_elements=True): fused kernels run on a Morton-sorted copy of the connectivity; the
    (E, Q)-shaped building blocks keep the caller's element order."""
    from tatva_b200 import sparse

    rng = np.random.default_rng(3)
    c, el, u, v, (mname, omat) = _case("tet4", 8)
    el = el[rng.permutation(el.shape[0])]
    mat = _material(mname, omat)
    op0, op1 = _make_op("tet4", c, el), _make_op("tet4", c, el, sort_elements=True)
    _assert_close(op1.hvp(mat)(u, v), orc.hvp("tet4", omat, c, el, u, v))
    _assert_close(op1.residual(mat)(u), orc.residual("tet4", omat, c, el, u))
    _assert_close(op1.energy(mat)(u), orc.energy("tet4", omat, c, el, u))
    _assert_close(op1.grad(u), op0.grad(u).cpu().numpy(), 0.0)  # caller's element order, bit-identical
    cm = sparse.ColoredMatrix.from_csr(sparse.pattern_from_mesh(op1.mesh, 3))
    _assert_close(sparse.assembler(op1, mat, cm)(u), sparse.assembler(op0, mat, cm)(u).cpu().numpy(), 1e-13)


@pytest.mark.parametrize("kind", ["quad4", "tri6", "quad8"])
def test_second_order_and_quad_elements(golden, kind):
    """Quad4 / Tri6 / Quad8 through the same kernel templates: building blocks vs the reference's outputs,
    fused linear-elastic energy / residual / HVP / assembly vs the oracle."""
    from tatva_b200 import element, materials, sparse
    import tatva_b200

    cls = {"quad4": element.Quad4, "tri6": element.Tri6, "quad8": element.Quad8}[kind]
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, v = g("coords"), g("conn"), g("u"), g("v")
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), cls())
    _assert_close(op.grad(u), g("grad_u"))
    _assert_close(op.grad(g("s")), g("grad_s"))
    _assert_close(op.eval(u), g("eval_u"))
    _assert_close(op.get_integration_weights(), g("weights"))
    _assert_close(op.integrate(g("s")), g("int_nodal_s"))
    _assert_close(op.integrate_per_element(g("quadvals")), g("int_quad_per_el"))
    mu, lm = g("mat")
    mat, omat = materials.LinearElastic(mu, lm), orc.LinearElastic(mu, lm)
    _assert_close(op.energy(mat)(u), g("energy"))
    _assert_close(op.residual(mat)(u), orc.residual(kind, omat, c, el, u))
    _assert_close(op.hvp(mat)(u, v), orc.hvp(kind, omat, c, el, u, v))
    pat = sparse.pattern_from_mesh(op.mesh, 2)
    cm = sparse.ColoredMatrix.from_csr(pat)
    data = sparse.assembler(op, mat, cm)(u).cpu().numpy()
    _assert_close(data, orc.assemble_csr_data(kind, omat, c, el, u, pat.indptr, pat.indices))


@pytest.mark.parametrize("kind", ["line2", "line3"])
@pytest.mark.parametrize("variant", [0, 1])
def test_line_elements_on_a_curved_boundary(golden, kind, variant):
    """Line2 / Line3 kernels (arc-length Jacobian |dX/dxi|, derivative along the line, no spatial axis) against the
    reference's outputs, plus the adjoint identity <grad u, g> = <u, grad^T g> and a boundary-traction integral by
    autograd (the way the reference builds Neumann terms: op.integrate(t . op.eval(u)))."""
    from tatva_b200 import element
    import tatva_b200

    cls = {"line2": element.Line2, "line3": element.Line3}[kind]
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, s = g("coords"), g("conn"), g("u"), g("s")
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), cls())
    op.set_variant(variant)  # 1: plain kernels, 0: warp-staged ones
    _assert_close(op.grad(u), g("grad_u"))
    _assert_close(op.grad(s), g("grad_s"))
    _assert_close(op.eval(u), g("eval_u"))
    _assert_close(op.get_integration_weights(), g("weights"))
    _assert_close(op.integrate(s), g("int_nodal_s"))
    _assert_close(op.integrate_per_element(g("quadvals")), g("int_quad_per_el"))
    # larger closed curve: the oracle at a size with many CTAs, and the adjoint
    n = 5000
    t = np.linspace(0, 2 * np.pi, n, endpoint=False)
    r = 1.0 + 0.2 * np.cos(3 * t)
    ends = np.stack([r * np.cos(t), r * np.sin(t)], -1)
    if kind == "line2":
        cc, ee = ends, np.stack([np.arange(n), (np.arange(n) + 1) % n], -1).astype(np.int32)
    else:
        tm = t + np.pi / n
        rm = 1.0 + 0.2 * np.cos(3 * tm)
        cc = np.concatenate([ends, np.stack([rm * np.cos(tm), rm * np.sin(tm)], -1)])
        ee = np.stack([np.arange(n), (np.arange(n) + 1) % n, n + np.arange(n)], -1).astype(np.int32)
    rng = np.random.default_rng(0)
    uu = rng.normal(size=(cc.shape[0], 3))
    op2 = tatva_b200.Operator(tatva_b200.Mesh(coords=cc, elements=ee), cls())
    op2.set_variant(variant)
    G = op2.grad(uu)
    _assert_close(G, orc.op_grad(kind, cc, ee, uu))
    _assert_close(op2.get_integration_weights(), orc.op_integration_weights(kind, cc, ee))
    gq = torch.as_tensor(rng.normal(size=tuple(G.shape)), device="cuda")
    ut = torch.tensor(uu, device="cuda", requires_grad=True)
    (op2.grad(ut) * gq).sum().backward()
    ref_adj = np.zeros_like(uu)
    dNdX, _ = orc.geometry(kind, cc, ee)
    np.add.at(ref_adj, ee, np.einsum("eqn,eqv->env", dNdX[:, :, 0, :], gq.cpu().numpy()))
    _assert_close(ut.grad, ref_adj)
    # traction work W(u) = int t . u ds  ->  dW/du = consistent nodal forces; they sum to t * length
    trac = torch.tensor([0.3, -1.1, 0.7], dtype=torch.float64, device="cuda")
    ut = torch.tensor(uu, device="cuda", requires_grad=True)
    op2.integrate((op2.eval(ut) * trac).sum(-1)).backward()
    length = float(op2.get_integration_weights().sum())
    np.testing.assert_allclose(ut.grad.sum(0).cpu().numpy(), trac.cpu().numpy() * length, rtol=1e-12)


def test_line_element_needs_plane_coordinates():
    from tatva_b200 import element
    import tatva_b200

    with pytest.raises(ValueError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=np.zeros((3, 3)), elements=np.array([[0, 1]], dtype=np.int32)), element.Line2())
    with pytest.raises(ValueError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=np.random.default_rng(0).normal(size=(4, 3)), elements=np.array([[0, 1, 2]], dtype=np.int32)), element.Tri3())


@pytest.mark.parametrize("kind", ["tri3", "quad4"])
def test_interpolate_matches_reference(golden, kind):
    """Operator.interpolate against the reference's outputs (operator.py:399-463; tests/test_operator.py:145-166):
    interior points, points on shared edges and nodes (first containing element wins), points outside."""
    from tatva_b200 import element
    import tatva_b200

    g = lambda k: golden[f"interp_{kind}_{k}"]  # noqa: E731
    c, el = g("coords"), g("conn")
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), {"tri3": element.Tri3, "quad4": element.Quad4}[kind]())
    _assert_close(op.interpolate(g("u"), g("points")), g("values_u"))
    _assert_close(op.interpolate(g("s"), g("points")), g("values_s"))
    with pytest.raises(RuntimeError):
        op.interpolate(g("s"), g("outside"))
    # the element index the kernel reports == mesh.find_containing_polygons of the reference, outside points included
    allp = torch.as_tensor(np.concatenate([g("points"), g("outside")]), device="cuda")
    out = torch.empty((allp.shape[0], 1), dtype=torch.float64, device="cuda")
    elem = torch.empty(allp.shape[0], dtype=torch.int32, device="cuda")
    op._call("tatva_op_interpolate", torch.as_tensor(g("s"), device="cuda").data_ptr(), 1, allp.data_ptr(), allp.shape[0], out.data_ptr(), elem.data_ptr())
    np.testing.assert_array_equal(elem.cpu().numpy(), g("containing"))
    assert bool(torch.isnan(out[-3:]).all())
    # at scale against the oracle: many elements per point (several staging chunks), many points (several CTAs)
    rng = np.random.default_rng(4)
    cc, ee = (orc.mesh_unit_square_tri(24, 20) if kind == "tri3" else orc.mesh_unit_square_quad(21, 23))
    inner = (cc[:, 0] > 1e-9) & (cc[:, 0] < 1 - 1e-9) & (cc[:, 1] > 1e-9) & (cc[:, 1] < 1 - 1e-9)
    cc = cc + 0.01 * rng.uniform(-1, 1, cc.shape) * inner[:, None]
    pts = rng.uniform(0, 1, size=(1000, 2))
    uu = rng.normal(size=(cc.shape[0], 2, 2))
    op2 = tatva_b200.Operator(tatva_b200.Mesh(coords=cc, elements=ee), {"tri3": element.Tri3, "quad4": element.Quad4}[kind]())
    ref, idx = orc.op_interpolate(kind, cc, ee, uu, pts)
    assert (idx >= 0).all()
    got = op2.interpolate(uu, pts)  # background grid
    assert got.shape == (1000, 2, 2) and op2._point_grid is not None
    _assert_close(got, ref)
    op2.set_variant(1)  # full scan of every element
    full = op2.interpolate(uu, pts)
    assert torch.equal(full, got)


@pytest.mark.parametrize("kind", ["tri3", "quad4"])
def test_find_containing_polygons_matches_reference(golden, kind):
    """tatva.mesh.find_containing_polygons (mesh.py:294-388) as a stand-alone function."""
    from tatva_b200.mesh import find_containing_polygons

    g = lambda k: golden[f"interp_{kind}_{k}"]  # noqa: E731
    allp = np.concatenate([g("points"), g("outside")])
    got = find_containing_polygons(allp, g("coords")[g("conn")])
    np.testing.assert_array_equal(got.cpu().numpy(), g("containing"))


def test_interpolate_known_answer_two_triangles():
    """reference tests/test_operator.py:145-159."""
    from tatva_b200 import element
    import tatva_b200

    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=nodes, elements=tris), element.Tri3())
    pts = np.array([[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.5, 0.5]])
    np.testing.assert_allclose(op.interpolate(nodes.sum(axis=1), pts).cpu().numpy(), pts.sum(axis=1), rtol=1e-14)
    with pytest.raises(RuntimeError):
        op.interpolate(np.ones(4), np.array([[1.5, 0.5]]))


@pytest.mark.parametrize("kind", ["quad4", "quad8", "hex8"])
def test_project_quadrature_fields(kind):
    """Operator.project (operator.py:518-554): the reference's own tests (tests/test_operator_projection.py: linear
    scalar / vector / tensor fields are reproduced on Quad4 2x2, with scalar, coupled and default matrices), and the
    oracle's consistent mass matrix solved directly for a non-polynomial field."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    from tatva_b200 import element, sparse
    from tatva_b200.lifter import Fixed, Lifter
    import tatva_b200

    rng = np.random.default_rng(9)
    if kind == "quad4":
        c, el = orc.mesh_unit_square_quad(2, 2)
        cls = element.Quad4
    elif kind == "quad8":  # (the one-point / three-point triangle rules under-integrate N_a N_b: singular mass matrix)
        c, el = orc.mesh_second_order("quad8", 4, 3)
        cls = element.Quad8
    else:
        c, el = orc.mesh_box_hex(3)
        c = c + 0.03 * rng.uniform(-1, 1, c.shape)
        cls = element.Hexahedron8
    mesh = tatva_b200.Mesh(coords=c, elements=el)
    op = tatva_b200.Operator(mesh, cls())
    dim = c.shape[1]
    qp = op.quads()
    _assert_close(qp, orc.op_eval(kind, c, el, c))
    cm1 = sparse.ColoredMatrix.from_csr(sparse.pattern_from_mesh(mesh, 1))
    cmd = sparse.ColoredMatrix.from_csr(sparse.pattern_from_mesh(mesh, dim))
    N = c.shape[0]
    # linear fields are in the space: reproduced exactly
    lin = qp[:, :, 0] + 2.0 * qp[:, :, 1]
    for kw in (dict(colored_matrix=cm1), dict()):
        got = op.project(lin, **kw)
        assert got.shape == (N,)
        np.testing.assert_allclose(got.cpu().numpy(), c[:, 0] + 2.0 * c[:, 1], atol=1e-10)
    for cm in (cm1, cmd):
        got = op.project(qp, cm)
        assert got.shape == (N, dim)
        np.testing.assert_allclose(got.cpu().numpy(), c, atol=1e-10)
    T = torch.zeros(qp.shape[:2] + (2, 2), dtype=torch.float64, device="cuda")
    T[:, :, 0, 0], T[:, :, 1, 1] = qp[:, :, 0], qp[:, :, 1]
    got = op.project(T, cm1).cpu().numpy()
    assert got.shape == (N, 2, 2)
    np.testing.assert_allclose(got[:, 0, 0], c[:, 0], atol=1e-10)
    np.testing.assert_allclose(got[:, 1, 1], c[:, 1], atol=1e-10)
    np.testing.assert_allclose(got[:, 0, 1], 0.0, atol=1e-10)
    # a field outside the space: consistent mass matrix of the oracle, direct solve
    W = orc.op_integration_weights(kind, c, el)
    Nq = np.stack([orc.shape_function(kind, x) for x in orc.quad_rule(kind)[0]])
    Me = np.einsum("eq,qa,qb->eab", W, Nq, Nq)
    npe = el.shape[1]
    M = sps.coo_matrix((Me.ravel(), (np.repeat(el, npe, axis=1).ravel(), np.tile(el, (1, npe)).ravel())), shape=(N, N)).tocsc()
    fq = np.sin(3 * qp.cpu().numpy()[:, :, 0]) * np.exp(qp.cpu().numpy()[:, :, 1])
    bq = np.zeros(N)
    np.add.at(bq, el, np.einsum("eq,qa->ea", W * fq, Nq))
    ref = spla.spsolve(M, bq)
    _assert_close(op.project(fq), ref, tol=1e-10)
    _assert_close(op.project(fq, use_graph=True), ref, tol=1e-10)
    # Fixed DOFs through a lifter (utils.py:193-201, :233-236): reduced system + lifted solution
    fixed = np.where(c[:, 0] < 1e-9)[0]
    lifter = Lifter(N, Fixed(fixed, 0.25))
    free = lifter.free_dofs
    xr = spla.spsolve(M[free][:, free].tocsc(), bq[free])
    ref_l = np.full(N, 0.25)
    ref_l[free] = xr
    _assert_close(op.project(fq, lifter=lifter), ref_l, tol=1e-10)


def test_tiled_tet4_kernels_match_oracle():
    """Operator(stage_tiles=True): Tet4 neo-Hookean residual / HVP through the shared-memory staging tiles."""
    c, el, u, v, (mname, omat) = _case("tet4", 9)
    mat = _material(mname, omat)
    for kw in (dict(stage_tiles=True), dict(stage_tiles=True, sort_elements=True)):
        op = _make_op("tet4", c, el[np.random.default_rng(4).permutation(el.shape[0])] if "sort_elements" in kw else el, **kw)
        elx = np.asarray(op.mesh.elements)
        _assert_close(op.hvp(mat)(u, v), orc.hvp("tet4", omat, c, elx, u, v))
        _assert_close(op.residual(mat)(u), orc.residual("tet4", omat, c, elx, u))


def test_edge_cases_single_element_and_inverted_orientation():
    """One-element meshes, and an inverted element: the reference takes det J without abs
    (operator.py:172-192), so the weights (and integrals) are negative."""
    from tatva_b200 import materials

    X = orc.reference_nodes("tet4")
    el = np.array([[0, 1, 2, 3]], dtype=np.int32)
    op = _make_op("tet4", X, el)
    _assert_close(op.get_integration_weights(), np.array([[1.0 / 6.0]]), 1e-15)
    inv = np.array([[0, 2, 1, 3]], dtype=np.int32)  # swapped nodes: negative orientation
    op_inv = _make_op("tet4", X, inv)
    _assert_close(op_inv.get_integration_weights(), orc.op_integration_weights("tet4", X, inv), 1e-15)
    assert float(op_inv.integrate(np.ones(4))) < 0
    rng = np.random.default_rng(0)
    Xh = orc.reference_nodes("hex8") + 0.1 * rng.uniform(-1, 1, (8, 3))
    elh = np.arange(8, dtype=np.int32)[None, :]
    oph = _make_op("hex8", Xh, elh)
    u, v = 0.05 * rng.normal(size=(8, 3)), rng.normal(size=(8, 3))
    omat, mat = orc.NeoHookean(500.0, 1000.0), materials.NeoHookean(500.0, 1000.0)
    _assert_close(oph.hvp(mat)(u, v), orc.hvp("hex8", omat, Xh, elh, u, v))
    _assert_close(oph.residual(mat)(u), orc.residual("hex8", omat, Xh, elh, u))
    _assert_close(oph.energy(mat)(u), orc.energy("hex8", omat, Xh, elh, u))


def test_invalid_meshes_raise_like_the_reference():
    """operator.py:132-170 (__check_init__)."""
    import tatva_b200
    from tatva_b200 import element

    X = orc.reference_nodes("tri3")
    with pytest.raises(ValueError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.array([[0, 1, 5]], dtype=np.int32)), element.Tri3())
    with pytest.raises(ValueError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.array([[0, 1, -1]], dtype=np.int32)), element.Tri3())
    with pytest.raises(TypeError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.array([[0.0, 1.0, 2.0]])), element.Tri3())
    with pytest.raises(ValueError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.zeros((0, 3), dtype=np.int32)), element.Tri3())
    # a user quadrature rule is accepted (r02); more than 64 points is not
    op = tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.array([[0, 1, 2]], dtype=np.int32)), element.Tri3(quad_points=np.array([[0.2, 0.2]]), quad_weights=np.array([0.5])))
    _assert_close(op.get_integration_weights(), np.array([[0.5]]), 1e-15)
    with pytest.raises(NotImplementedError):
        tatva_b200.Operator(tatva_b200.Mesh(coords=X, elements=np.array([[0, 1, 2]], dtype=np.int32)), element.Tri3(quad_points=np.full((65, 2), 0.2), quad_weights=np.full(65, 0.5 / 65)))


# ---- user quadrature rules: Element(quad_points, quad_weights), reference element/base.py:37-51 -----------------


@pytest.mark.parametrize("kind", ["hex8", "tet4", "quad4"])
def test_custom_quadrature_rule_blocks_match_reference_fixtures(golden, kind):
    """Operator building blocks with the Hex8 3x3x3, Tet4 4-point and Quad4 3x3 rules against the reference's outputs."""
    tb, element, materials = _tb()
    g = lambda k: golden[f"cq_{kind}_{k}"]  # noqa: E731
    cls = {"hex8": element.Hexahedron8, "tet4": element.Tetrahedron4, "quad4": element.Quad4}[kind]
    for cache in (False, True):
        op = tb.Operator(tb.Mesh(coords=g("coords"), elements=g("conn")), cls(quad_points=g("qp"), quad_weights=g("qw")), cache_weights=cache)
        assert op.nq == len(g("qw"))
        _assert_close(op.grad(g("u")), g("grad_u"))
        _assert_close(op.eval(g("s")), g("eval_s"))
        _assert_close(op.get_integration_weights(), g("weights"))
        _assert_close(op.integrate(g("s")), g("int_nodal_s"))
        _assert_close(op.integrate_per_element(g("quadvals")), g("int_quad_per_el"))
    mat = materials.LinearElastic(*g("prm")) if kind == "quad4" else materials.NeoHookean(*g("prm"))
    _assert_close(op.energy(mat)(g("u")), g("energy"))
    _assert_close(op.residual(mat)(g("u")), g("residual_cs"))
    # autograd through the building blocks (adjoint kernels with the custom rule) gives the same residual
    ut = torch.as_tensor(g("u"), device="cuda").requires_grad_(True)
    mu, lm = (float(x) for x in g("prm"))
    G = op.grad(ut)
    if kind == "quad4":
        eps = 0.5 * (G + G.transpose(-1, -2))
        psi = mu * (eps * eps).sum((-1, -2)) + 0.5 * lm * eps.diagonal(dim1=-2, dim2=-1).sum(-1) ** 2
    else:
        F = G + torch.eye(3, dtype=torch.float64, device="cuda")
        lnJ = torch.log(torch.linalg.det(F))
        psi = 0.5 * mu * ((F * F).sum((-1, -2)) - 3 - 2 * lnJ) + 0.5 * lm * lnJ * lnJ
    (r,) = torch.autograd.grad(op.integrate(psi), ut)
    _assert_close(r, g("residual_cs"), 1e-11)


@pytest.mark.parametrize("kind,order,n", [("hex8", 3, 6), ("tet4", 2, 5), ("tri3", 2, 12)])
def test_custom_quadrature_rule_fused_kernels_vs_oracle(kind, order, n):
    """Fused energy / residual / HVP / Hessian diagonal / CSR assembly with a user rule (generic kernels, rule in
    constant memory) against the oracle run with the same rule, 1e-12."""
    import scipy.sparse as sps

    tb, element, materials = _tb()
    from tatva_b200 import sparse

    c, el, u, v, (mname, omat) = _case(kind, n)
    qp, qw = orc.gauss_rule(kind, order)
    cls = getattr(element, ELEMS[kind])
    op = tb.Operator(tb.Mesh(coords=c, elements=el), cls(quad_points=qp, quad_weights=qw))
    mat = _material(mname, omat)
    dpn = c.shape[1]
    with orc.custom_rule(kind, qp, qw):
        _assert_close(op.energy(mat)(u), orc.energy(kind, omat, c, el, u))
        _assert_close(op.residual(mat)(u), orc.residual(kind, omat, c, el, u))
        Hv = orc.hvp(kind, omat, c, el, u, v)
        _assert_close(op.hvp(mat)(u, v), Hv)
        pat = sparse.pattern_from_mesh(op.mesh, dpn)
        cm = sparse.ColoredMatrix.from_csr(pat)
        data = sparse.assembler(op, mat, cm)(u).cpu().numpy()
        ref = orc.assemble_csr_data(kind, omat, c, el, u, pat.indptr, pat.indices)
        assert _rel(data, ref) < 1e-12
        K = sps.csr_matrix((data, pat.indices, pat.indptr), shape=pat.shape)
        assert _rel(K @ v.ravel(), Hv.ravel()) < 1e-12
        _assert_close(op.hessian_diagonal(mat, torch.as_tensor(u, device="cuda")).reshape(-1), K.diagonal())
    # the default-rule operator on the same mesh gives a DIFFERENT (under-integrated) answer for Hex8: the rule is in use
    if kind == "hex8":
        op0 = _make_op(kind, c, el)
        assert _rel(op0.hvp(mat)(u, v).cpu().numpy(), Hv) > 1e-8


@pytest.mark.parametrize("kind", ["tri3", "tet4", "hex8"])
def test_fused_kernels_against_derivatives_of_the_reference_energy(golden, kind):
    """The CUDA energy / residual / HVP directly against quantities derived from the REFERENCE's own energy
    op.integrate(psi(op.grad(u))) (tests/golden/make_golden.py): E itself, dE/du by complex step (~1e-16), and
    w . H v by complex step x 8th-order central differences (~1e-13), tolerance 1e-12 — no oracle in between.
    Every Hex8 HVP kernel variant is held to the same fixture."""
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, v = g("coords"), g("conn"), g("u"), g("v")
    _, _, materials = _tb()
    mat = materials.LinearElastic(*g("mat")) if kind == "tri3" else materials.NeoHookean(*g("mat"))
    op = _make_op(kind, c, el)
    assert abs(float(op.energy(mat)(u)) - float(g("energy"))) <= 1e-12 * abs(float(g("energy")))
    r = op.residual(mat)(u).cpu().numpy()
    assert np.abs(r - g("residual_cs")).max() <= 1e-11 * np.abs(r).max()
    variants = op.hvp_variants() if kind == "hex8" else (0,)
    for variant in variants:
        op.set_variant(variant)
        Hv = op.hvp(mat)(u, v).cpu().numpy()
        wHv = np.einsum("kni,ni->k", g("hvp_probe_w"), Hv)
        np.testing.assert_allclose(wHv, g("hvp_probe_wHv_hi"), rtol=1e-12, err_msg=f"variant {variant}")
    op.set_variant(0)


def test_plan_rebind_points_a_plan_at_other_buffers_of_the_same_sizes():
    """`tatva_plan_rebind` (what the XLA-FFI shim does per call: XLA re-allocates buffers between executions)."""
    tb, element, materials = _tb()
    c, el, u, v, (mname, omat) = _case("hex8", 4)
    mat = _material(mname, omat)
    op = _make_op("hex8", c, el)
    ref = op.hvp(mat)(u, v).cpu().numpy()
    c2 = c + 0.01 * np.random.default_rng(3).uniform(-1, 1, c.shape)
    coords2 = torch.as_tensor(c2, device="cuda").contiguous()
    conn2 = torch.as_tensor(el, device="cuda").to(torch.int32).contiguous().clone()
    from tatva_b200 import _lib

    for plan in {id(p): p for p in (op._plan, op._plan_fused)}.values():
        _lib.check(op._L.tatva_plan_rebind(plan, coords2.data_ptr(), conn2.data_ptr()), "tatva_plan_rebind")
    got = op.hvp(mat)(u, v).cpu().numpy()
    want = _make_op("hex8", c2, el).hvp(mat)(u, v).cpu().numpy()
    assert _rel(got, want) < 1e-14 and _rel(got, ref) > 1e-6
    opw = _make_op("hex8", c, el, cache_weights=True)
    assert op._L.tatva_plan_rebind(opw._plan, coords2.data_ptr(), conn2.data_ptr()) == -2  # cached weights: unsupported
