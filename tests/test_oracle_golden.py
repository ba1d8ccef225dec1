"""Pin the CPU oracle (oracle/tatva_oracle.py) to the reference.

Fixtures in tests/golden/reference_golden.npz are outputs of the unmodified reference code
(tests/golden/make_golden.py); the known-answer values are those of the reference's own tests.
"""
import numpy as np
import pytest

from oracle import tatva_oracle as orc

KINDS = ["tri3", "tet4", "hex8"]
MATS = {"tri3": orc.LinearElastic, "tet4": orc.NeoHookean, "hex8": orc.NeoHookean}


@pytest.mark.parametrize("kind", KINDS)
def test_element_math_matches_reference(golden, kind):
    g = lambda k: golden[f"el_{kind}_{k}"]  # noqa: E731
    qp, qw = orc.quad_rule(kind)
    np.testing.assert_array_equal(qp, g("qp"))
    np.testing.assert_array_equal(qw, g("qw"))
    X, uv, us = g("X"), g("uv"), g("us")
    for q, xi in enumerate(qp):
        np.testing.assert_allclose(orc.shape_function(kind, xi), g("N")[q], rtol=0, atol=1e-15)
        np.testing.assert_allclose(orc.shape_function_derivative(kind, xi), g("dNdr")[q], rtol=0, atol=1e-15)
        J, detJ = orc.get_jacobian(kind, xi, X)
        np.testing.assert_allclose(J, g("J")[q], rtol=1e-15, atol=1e-15)
        np.testing.assert_allclose(detJ, g("detJ")[q], rtol=1e-14)
        np.testing.assert_allclose(orc.element_gradient(kind, xi, uv, X), g("grad_v")[q], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(orc.element_gradient(kind, xi, us, X), g("grad_s")[q], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(orc.element_interpolate(kind, xi, uv, X), g("interp_v")[q], rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("kind", KINDS)
def test_linear_fields_have_exact_gradients(kind):
    """reference tests/test_element.py:45-149 (scalar, vector, tensor; atol 1e-12)."""
    rng = np.random.default_rng(0)
    X = orc.reference_nodes(kind)
    dim = X.shape[1]
    a = np.array([2.0, 3.0, 4.0][:dim])
    A = rng.normal(size=(dim, dim))
    B = rng.normal(size=(2, 2, dim))
    for xi in orc.quad_rule(kind)[0]:
        np.testing.assert_allclose(orc.element_gradient(kind, xi, X @ a, X), a, atol=1e-12)
        np.testing.assert_allclose(orc.element_gradient(kind, xi, np.einsum("ij,kj->ki", A, X), X), A, atol=1e-12)
        np.testing.assert_allclose(orc.element_gradient(kind, xi, np.einsum("ijk,nk->nij", B, X), X), B, atol=1e-12)


def test_operator_known_answers():
    """reference tests/test_operator.py:14-30, :113-143."""
    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    el = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    np.testing.assert_allclose(orc.op_eval("tri3", nodes, el, np.array([0.0, 1.0, 2.0, 3.0])), [[1.0], [5.0 / 3.0]])
    np.testing.assert_allclose(orc.op_grad("tri3", nodes, el, nodes @ np.array([2.0, 3.0])), [[[2.0, 3.0]], [[2.0, 3.0]]])
    np.testing.assert_allclose(orc.op_integrate_per_element("tri3", nodes, el, np.ones(4)), [0.5, 0.5])
    np.testing.assert_allclose(orc.op_integrate("tri3", nodes, el, np.ones(4)), 1.0)
    np.testing.assert_allclose(orc.op_integrate_per_element("tri3", nodes, el, np.full((2, 1), 4.0)), [2.0, 2.0])
    np.testing.assert_allclose(orc.op_integrate("tri3", nodes, el, 3.0), 3.0)


@pytest.mark.parametrize("kind", KINDS)
def test_operator_matches_reference(golden, kind):
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, s = g("coords"), g("conn"), g("u"), g("s")
    kw = dict(rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(orc.op_grad(kind, c, el, u), g("grad_u"), **kw)
    np.testing.assert_allclose(orc.op_grad(kind, c, el, s), g("grad_s"), **kw)
    np.testing.assert_allclose(orc.op_eval(kind, c, el, u), g("eval_u"), **kw)
    np.testing.assert_allclose(orc.op_eval(kind, c, el, s), g("eval_s"), **kw)
    np.testing.assert_allclose(orc.op_integration_weights(kind, c, el), g("weights"), **kw)
    np.testing.assert_allclose(orc.op_integrate(kind, c, el, s), g("int_nodal_s"), **kw)
    np.testing.assert_allclose(orc.op_integrate_per_element(kind, c, el, u), g("int_nodal_u_per_el"), **kw)
    np.testing.assert_allclose(orc.op_integrate_per_element(kind, c, el, g("quadvals")), g("int_quad_per_el"), **kw)


@pytest.mark.parametrize("kind", ["tri3", "quad4", "tri6", "quad8"])
def test_interpolate_and_point_location_match_reference(golden, kind):
    """Operator.interpolate / find_containing_polygons (reference operator.py:399-463, mesh.py:294-388), including
    points on shared edges, on nodes and outside the mesh."""
    g = lambda k: golden[f"interp_{kind}_{k}"]  # noqa: E731
    c, el = g("coords"), g("conn")
    allp = np.concatenate([g("points"), g("outside")])
    np.testing.assert_array_equal(orc.find_containing_polygons(allp, c[el]), g("containing"))
    vals, idx = orc.op_interpolate(kind, c, el, g("u"), g("points"))
    np.testing.assert_allclose(vals, g("values_u"), rtol=1e-12, atol=1e-13)
    vals_s, _ = orc.op_interpolate(kind, c, el, g("s"), g("points"))
    np.testing.assert_allclose(vals_s, g("values_s"), rtol=1e-12, atol=1e-13)
    out, idx = orc.op_interpolate(kind, c, el, g("s"), g("outside"))
    assert np.all(idx == -1) and np.all(np.isnan(out))
    # reference tests/test_operator.py:145-159: a linear field is recovered on the two-triangle unit square
    if kind == "tri3":
        nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
        tris = np.array([[0, 1, 2], [0, 2, 3]])
        pts = np.array([[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.5, 0.5]])
        got, _ = orc.op_interpolate("tri3", nodes, tris, nodes.sum(axis=1), pts)
        np.testing.assert_allclose(got, pts.sum(axis=1))


def test_find_containing_polygons_includes_boundary_points():
    """reference tests/test_mesh.py:7-24: a point on the edge shared by two polygons belongs to the first one."""
    polygons = np.array([[[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]], [[1.0, 0.0], [2.0, 0.0], [2.0, 1.0], [1.0, 1.0]]])
    points = np.array([[0.5, 0.5], [1.0, 0.5], [1.5, 0.5]])
    np.testing.assert_array_equal(orc.find_containing_polygons(points, polygons), [0, 0, 1])


@pytest.mark.parametrize("kind", ["line2", "line3"])
def test_line_operator_matches_reference(golden, kind):
    """Line2 / Line3 on a curved polyline: arc-length Jacobian and derivative (reference element/base.py:144-242)."""
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, s = g("coords"), g("conn"), g("u"), g("s")
    kw = dict(rtol=1e-13, atol=1e-13)
    qp, qw = orc.quad_rule(kind)
    np.testing.assert_allclose(qp, g("qp"), atol=1e-15)
    np.testing.assert_allclose(qw, g("qw"), atol=1e-15)
    assert orc.op_grad(kind, c, el, u).shape == g("grad_u").shape  # (E, Q, 2): no spatial axis
    np.testing.assert_allclose(orc.op_grad(kind, c, el, u), g("grad_u"), **kw)
    np.testing.assert_allclose(orc.op_grad(kind, c, el, s), g("grad_s"), **kw)
    np.testing.assert_allclose(orc.op_eval(kind, c, el, u), g("eval_u"), **kw)
    np.testing.assert_allclose(orc.op_integration_weights(kind, c, el), g("weights"), **kw)
    np.testing.assert_allclose(orc.op_integrate(kind, c, el, s), g("int_nodal_s"), **kw)
    np.testing.assert_allclose(orc.op_integrate_per_element(kind, c, el, g("quadvals")), g("int_quad_per_el"), **kw)
    # per-point element functions agree with the vectorised ones
    for e in range(2):
        for q, xi in enumerate(qp):
            np.testing.assert_allclose(orc.element_gradient(kind, xi, u[el[e]], c[el[e]]), g("grad_u")[e, q], **kw)
            np.testing.assert_allclose(orc.get_jacobian(kind, xi, c[el[e]])[1] * qw[q], g("weights")[e, q], **kw)
    # the arc is 1.3 rad of radius 1.3: Line3 integrates its length to 4 digits, the chords of Line2 to 3
    length = orc.op_integration_weights(kind, c, el).sum()
    assert abs(length - 1.3 * 1.3) < (5e-3 if kind == "line2" else 2e-4)


@pytest.mark.parametrize("kind", KINDS)
def test_energy_residual_hvp_match_reference_energy_derivatives(golden, kind):
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, v = g("coords"), g("conn"), g("u"), g("v")
    mat = MATS[kind](*g("mat"))
    np.testing.assert_allclose(orc.energy(kind, mat, c, el, u), g("energy"), rtol=1e-13)
    r = orc.residual(kind, mat, c, el, u)
    np.testing.assert_allclose(r, g("residual_cs"), rtol=1e-11, atol=1e-12 * np.abs(r).max())
    Hv = orc.hvp(kind, mat, c, el, u, v)
    wHv = np.einsum("kni,ni->k", g("hvp_probe_w"), Hv)
    # complex step x 4th-order central difference: ~1e-9 relative
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv"), rtol=1e-7)
    # r02: complex step x 8th-order central difference at two steps (recorded estimate <= 2e-10 at the coarse step, /256
    # at the fine one): the oracle's HVP is pinned by the REFERENCE's own energy to 1e-12
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv_hi"), rtol=1e-12)


@pytest.mark.parametrize("name", ["hex3", "tri4"])
def test_layouts_and_routing_match_the_reference_plans(golden, name):
    """tatva.mpi._create_dof_layout and ExchangePlan routing tables of the UNMODIFIED reference, built for every rank
    of a partitioned mesh on a thread-based mpi4py stand-in (3 ranks on a Hex8 box with 3 DOFs per node, 4 ranks on
    a Tri3 square with 2): the oracle's all-ranks-at-once restatement must reproduce them bit for bit."""
    size, dpn = (int(x) for x in golden[f"mpi_{name}_size_dpn"])
    g = lambda r, k: golden[f"mpi_{name}_r{r}_{k}"]  # noqa: E731
    naturals = [g(r, "natural") for r in range(size)]
    masks = [g(r, "owned_mask") for r in range(size)]
    n_nat = golden[f"mpi_{name}_coords"].shape[0] * dpn
    # the natural maps themselves follow from extract_local_mesh, restated in the oracle
    for r in range(size):
        _, _, l2g_nodes, n_owned = orc.extract_local_mesh(golden[f"mpi_{name}_coords"], golden[f"mpi_{name}_conn"], golden[f"mpi_{name}_partition"], r)
        np.testing.assert_array_equal(orc.dof_map_from_node_map(l2g_nodes, dpn), naturals[r])
        assert masks[r][: n_owned * dpn].all() and not masks[r][n_owned * dpn :].any()
    layouts = orc.create_dof_layouts(naturals, masks, n_nat)
    plans = orc.exchange_routing(layouts)
    for r in range(size):
        off, n_owned, n_total, n_global = (int(x) for x in g(r, "offset_nowned_ntotal_nglobal"))
        L = layouts[r]
        assert (L["offset"], L["n_owned"], L["n_total"], L["n_global"]) == (off, n_owned, n_total, n_global)
        np.testing.assert_array_equal(L["local_to_global"], g(r, "l2g"))
        np.testing.assert_array_equal(plans[r]["self_send"], g(r, "self_send"))
        np.testing.assert_array_equal(plans[r]["self_recv"], g(r, "self_recv"))
        np.testing.assert_array_equal([nb["rank"] for nb in plans[r]["neighbors"]], g(r, "nbr_ranks"))
        for nb in plans[r]["neighbors"]:
            np.testing.assert_array_equal(nb["local_send_idx"], g(r, f"nbr{nb['rank']}_send"))
            np.testing.assert_array_equal(nb["recv_local_idx"], g(r, f"nbr{nb['rank']}_recv"))


def test_coloured_jacobian_matches_the_reference_jacfwd(golden):
    """sparse.jacfwd / colored_jacobian_batch / compute_rows_cols of the reference (sparse/base.py:108-176, :230-270),
    run unmodified on fn(u) = A u + 0.1 (A u)^2 over the Tri3 8x8 two-DOF pattern: the oracle's decompression and
    the product's coloured path (torch forward-mode AD, here on CPU tensors) must give the same `data`."""
    import scipy.sparse as sps
    import torch

    from tatva_b200 import sparse

    g = lambda k: golden[f"jac_{k}"]  # noqa: E731
    ip, ix, colors, u0 = g("indptr"), g("indices"), g("colors"), g("u0")
    n = len(u0)
    A = sps.csr_matrix((g("A_data"), ix, ip), shape=(n, n))
    rows, col_colors = orc.compute_rows_cols(ip, ix, colors)
    np.testing.assert_array_equal(rows, g("rows"))
    np.testing.assert_array_equal(col_colors, g("col_colors"))
    np.testing.assert_array_equal(orc.distance2_colors(ip, ix, n), colors)
    y = A @ u0
    jvp = lambda seed: (1 + 0.2 * y) * (A @ seed)  # noqa: E731
    np.testing.assert_allclose(orc.colored_jacobian_data(jvp, n, ip, ix, colors), g("data_batchNone"), rtol=1e-13, atol=1e-14)
    np.testing.assert_array_equal(g("data_batchNone"), g("data_batch5"))
    # the product's reference-algorithm path on a plain callable
    pat = sps.csr_matrix((np.ones(len(ix), dtype=np.int8), ix, ip), shape=(n, n))
    cm = sparse.ColoredMatrix.from_csr(pat)
    np.testing.assert_array_equal(cm.colors, colors)
    r2, c2 = sparse.compute_rows_cols(cm)
    np.testing.assert_array_equal(r2, g("rows"))
    np.testing.assert_array_equal(c2, g("col_colors"))
    At = torch.as_tensor(A.toarray())

    def fn(u):
        yy = At @ u
        return yy + 0.1 * yy * yy

    K = sparse.jacfwd(fn, cm, color_batch_size=5)(torch.as_tensor(u0))
    np.testing.assert_allclose(np.asarray(K.data), g("data_batchNone"), rtol=1e-13, atol=1e-14)
    primal, K2 = sparse.linearized_jacfwd(fn, cm)(torch.as_tensor(u0))
    np.testing.assert_allclose(np.asarray(K2.data), g("data_batchNone"), rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(np.asarray(primal), y + 0.1 * y * y, rtol=1e-14)


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_phase_field_energy_through_the_reference_operator(golden, kind):
    """Config 5: the builder-defined AT2 density evaluated with the REFERENCE's Operator on the stacked state
    [ux, uy, uz, phi]; the oracle's closed-form energy / residual / HVP must match it and its complex-step
    derivatives.  (The law itself has no counterpart in the reference; everything around it is pinned here.)"""
    g = lambda k: golden[f"pf_{kind}_{k}"]  # noqa: E731
    c, el, s, t = g("coords"), g("conn"), g("s"), g("t")
    mat = orc.NeoHookeanPhaseField(*g("params"))
    np.testing.assert_allclose(orc.energy_pf(kind, mat, c, el, s), g("energy"), rtol=1e-13)
    r = orc.residual_pf(kind, mat, c, el, s)
    np.testing.assert_allclose(r, g("residual_cs"), rtol=1e-10, atol=1e-12 * np.abs(r).max())
    wHv = np.einsum("kni,ni->k", g("hvp_probe_w"), orc.hvp_pf(kind, mat, c, el, s, t))
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv"), rtol=1e-7)
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv_hi"), rtol=1e-12)  # 8th-order probe, see make_golden.hvp_probe_hi


@pytest.mark.parametrize("kind", ["quad4", "tri6", "quad8"])
def test_energy_derivatives_of_the_other_plane_elements(golden, kind):
    """Quad4 / Tri6 / Quad8 with the reference's linear-elastic density (tests/test_sparse.py:20-38): oracle energy,
    residual and HVP against the reference's energy and its complex-step derivatives."""
    g = lambda k: golden[f"op_{kind}_{k}"]  # noqa: E731
    c, el, u, v = g("coords"), g("conn"), g("u"), g("v")
    mat = orc.LinearElastic(*g("mat"))
    np.testing.assert_allclose(orc.energy(kind, mat, c, el, u), g("energy"), rtol=1e-13)
    r = orc.residual(kind, mat, c, el, u)
    np.testing.assert_allclose(r, g("residual_cs"), rtol=1e-11, atol=1e-12 * np.abs(r).max())
    wHv = np.einsum("kni,ni->k", g("hvp_probe_w"), orc.hvp(kind, mat, c, el, u, v))
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv"), rtol=1e-9)  # quadratic energy: the difference quotient is exact
    np.testing.assert_allclose(wHv, g("hvp_probe_wHv_hi"), rtol=1e-12)


@pytest.mark.parametrize("name,dpn", [("tri3_8x8_d2", 2), ("tet4_3_d3", 3), ("tet4_2_d4", 4), ("hex8_3_d3", 3)])
def test_pattern_and_colouring_bit_exact(golden, name, dpn):
    conn, n_nodes = golden[f"sp_{name}_conn"], int(golden[f"sp_{name}_nnodes"])
    indptr, indices = orc.pattern_from_mesh(conn, n_nodes, dpn)
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    np.testing.assert_array_equal(indptr, golden[f"sp_{name}_indptr"])
    np.testing.assert_array_equal(indices, golden[f"sp_{name}_indices"])
    colors = orc.distance2_colors(indptr, indices, n_nodes * dpn)
    np.testing.assert_array_equal(colors, golden[f"sp_{name}_colors"])


def test_tri3_pattern_nnz_formula():
    """nnz = 4 (7 n^2 + 6 n + 1) for Tri3 n x n with 2 DOFs per node."""
    for n in (4, 8):
        c, el = orc.mesh_unit_square_tri(n, n)
        indptr, _ = orc.pattern_from_mesh(el, len(c), 2)
        assert indptr[-1] == 4 * (7 * n * n + 6 * n + 1)


def test_coloured_jacobian_equals_dense_hessian():
    """reference tests/test_sparse.py:48-80: Tri3 8x8, mu=1, lambda=0, at u=0."""
    c, el = orc.mesh_unit_square_tri(8, 8)
    mat = orc.LinearElastic(1.0, 0.0)
    n = 2 * len(c)
    u0 = np.zeros((len(c), 2))
    indptr, indices = orc.pattern_from_mesh(el, len(c), 2)
    colors = orc.distance2_colors(indptr, indices, n)
    jvp = lambda seed: orc.hvp("tri3", mat, c, el, u0, seed.reshape(-1, 2)).ravel()  # noqa: E731
    data = orc.colored_jacobian_data(jvp, n, indptr, indices, colors)
    K = np.stack([jvp(e) for e in np.eye(n)], axis=1)
    import scipy.sparse as sps

    Ks = sps.csr_matrix((data, indices, indptr), shape=(n, n)).toarray()
    np.testing.assert_allclose(Ks, K, rtol=1e-12, atol=1e-14)
    direct = orc.assemble_csr_data("tri3", mat, c, el, u0, indptr, indices)
    np.testing.assert_allclose(direct, data, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(orc.residual("tri3", mat, c, el, u0), 0.0)


@pytest.mark.parametrize("case,nparts", [("hex", 2), ("tri", 4)])
def test_extract_local_mesh_matches_reference(golden, case, nparts):
    c, el, part = golden[f"part_{case}_coords"], golden[f"part_{case}_conn"], golden[f"part_{case}_partition"]
    for r in range(nparts):
        cl, el_l, l2g, n_owned = orc.extract_local_mesh(c, el, part, r)
        np.testing.assert_array_equal(el_l, golden[f"part_{case}_r{r}_conn"])
        np.testing.assert_array_equal(l2g, golden[f"part_{case}_r{r}_l2g"])
        assert n_owned == int(golden[f"part_{case}_r{r}_nowned"])
        np.testing.assert_array_equal(cl, c[l2g])


def test_dof_range_block_distribution():
    """reference tests/test_allreduce_plan.py:31-40: 6 DOFs on 2 ranks -> [0,3), [3,6); mpi.py:714-726."""
    assert orc.dof_range(6, 2, 0) == (0, 3) and orc.dof_range(6, 2, 1) == (3, 6)
    assert orc.dof_range(5, 2, 0) == (0, 3) and orc.dof_range(5, 2, 1) == (3, 5)
    assert [orc.dof_range(7, 3, r) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]


def test_exchange_plan_known_answers():
    """reference tests/test_exchange_plan.py:97-166: 2 ranks, 2 nodes, 1 DOF per node."""
    nat = [np.array([0, 1], dtype=np.int32), np.array([1, 0], dtype=np.int32)]
    own = [np.array([True, False]), np.array([True, False])]
    layouts = orc.create_dof_layouts(nat, own, 2)
    np.testing.assert_array_equal(layouts[0]["local_to_global"], [0, 1])
    np.testing.assert_array_equal(layouts[1]["local_to_global"], [1, 0])
    plans = orc.exchange_routing(layouts)
    x = [np.array([10.0]), np.array([20.0])]
    ul = orc.scatter_fwd_set(plans, layouts, x)
    np.testing.assert_allclose(ul[0], [10.0, 20.0])
    np.testing.assert_allclose(ul[1], [20.0, 10.0])
    owned = orc.scatter_rev_add(plans, layouts, [2 * a for a in ul])
    np.testing.assert_allclose(owned[0], [40.0])
    np.testing.assert_allclose(owned[1], [80.0])


def test_exchange_plan_layout_two_dofs_per_node():
    """Nodal part of reference tests/test_exchange_plan.py:27-94 (u: 2 nodes x 2 DOFs):
    rank-contiguous owned blocks, ghosts resolved through the natural directory."""
    nat = [orc.dof_map_from_node_map([0, 1], 2), orc.dof_map_from_node_map([1, 0], 2)]
    own = [np.array([True, True, False, False])] * 2
    layouts = orc.create_dof_layouts(nat, own, 4)
    np.testing.assert_array_equal(layouts[0]["local_to_global"], [0, 1, 2, 3])
    np.testing.assert_array_equal(layouts[1]["local_to_global"], [2, 3, 0, 1])
    assert layouts[1]["offset"] == 2 and layouts[0]["n_global"] == 4


@pytest.mark.parametrize("kind", ["hex8", "tet4", "quad4"])
def test_oracle_with_a_custom_quadrature_rule_matches_the_reference(golden, kind):
    """Element(quad_points, quad_weights) (reference element/base.py:37-51): the oracle run under `custom_rule`
    reproduces the reference Operator with the Hex8 3x3x3, Tet4 4-point and Quad4 3x3 rules."""
    g = lambda k: golden[f"cq_{kind}_{k}"]  # noqa: E731
    c, el, u, s = g("coords"), g("conn"), g("u"), g("s")
    mat = orc.LinearElastic(*g("prm")) if kind == "quad4" else orc.NeoHookean(*g("prm"))
    with orc.custom_rule(kind, g("qp"), g("qw")):
        np.testing.assert_allclose(orc.op_grad(kind, c, el, u), g("grad_u"), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(orc.op_eval(kind, c, el, s), g("eval_s"), rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(orc.op_integration_weights(kind, c, el), g("weights"), rtol=1e-13)
        np.testing.assert_allclose(orc.op_integrate(kind, c, el, s), g("int_nodal_s"), rtol=1e-12)
        np.testing.assert_allclose(orc.op_integrate_per_element(kind, c, el, g("quadvals")), g("int_quad_per_el"), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(orc.energy(kind, mat, c, el, u), g("energy"), rtol=1e-13)
        r = orc.residual(kind, mat, c, el, u)
        assert np.linalg.norm(r - g("residual_cs")) <= 1e-12 * np.linalg.norm(r)
    # the override is gone
    assert len(orc.quad_rule(kind)[1]) == {"hex8": 8, "tet4": 1, "quad4": 4}[kind]
