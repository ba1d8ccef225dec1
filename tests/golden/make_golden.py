"""Generate golden fixtures by running the UNMODIFIED reference (tatva v0.11.1).

Run in the build container only (it needs /root/reference, which the GPU box lacks):

    python tests/golden/make_golden.py

JAX cannot be installed here, so the reference modules are executed eagerly on the NumPy
stand-in under tests/golden/_jaxshim (vmap / lax.map -> loops, jnp -> numpy).  Every array
written below is the output of reference code: tatva.element.*, tatva.Operator.{grad, eval,
integrate, integrate_per_element, get_integration_weights}, tatva.sparse.pattern_from_mesh,
tatva/sparse/_coloring.py:distance2_colors, tatva.mesh.extract_local_mesh, and (with an
in-process thread-based stand-in for mpi4py, tests/golden/_fakempi) tatva.mpi._create_dof_layout /
ExchangePlan routing tables for vectors and Hessian nonzeros (tests/golden/_fakempi_golden.py).

Derivative fixtures: the reference has no residual/HVP code (they are jax.grad / jax.jvp of a
user energy).  We differentiate the reference's *own* energy E(u) = op.integrate(psi(op.grad(u)))
with the complex-step method (exact to rounding for analytic E):
    r_i = Im E(u + i h e_i) / h,                   h = 1e-30
    (Hv)_i = Im[ r-free mixed form ] -> see hvp_complex_fd below (complex step x central FD).
The energy densities are written as in tests/test_sparse.py:20-38 and
tests/test_sparse_tracer.py:103-115.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TATVA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_jaxshim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import jax.numpy as jnp  # noqa: E402  (the shim)
from jax_autovmap import autovmap  # noqa: E402
from tatva import Mesh, Operator, element, sparse  # noqa: E402
from tatva.mesh import extract_local_mesh  # noqa: E402
from tatva_coloring import distance2_colors  # noqa: E402

sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import tatva_oracle as orc  # noqa: E402  (mesh generators only: inputs, not outputs)

KINDS = {"tri3": element.Tri3, "tet4": element.Tetrahedron4, "hex8": element.Hexahedron8}
MORE_KINDS = {"quad4": element.Quad4, "tri6": element.Tri6, "quad8": element.Quad8}


# -- user energies exactly as the reference tests write them -----------------------------
@autovmap(grad_u=2, mu=0, lmbda=0)
def strain_energy(grad_u, mu, lmbda):  # tests/test_sparse.py:20-38
    eps = 0.5 * (grad_u + grad_u.T)
    sig = 2 * mu * eps + lmbda * jnp.trace(eps) * jnp.eye(grad_u.shape[0])
    return 0.5 * jnp.einsum("ij,ij->", sig, eps)


@autovmap(grad_u=2, mu=0, lmbda=0)
def neo_hookean_density(grad_u, mu, lmbda):  # tests/test_sparse_tracer.py:103-115
    F = jnp.eye(3) + grad_u
    J = jnp.linalg.det(F)
    C = F.T @ F
    I1 = jnp.trace(C)
    return (mu / 2) * (I1 - 3 - 2 * jnp.log(J)) + (lmbda / 2) * (jnp.log(J)) ** 2


@autovmap(grad_u=2, phi=0, grad_phi=1, mu=0, lmbda=0, Gc=0, ell=0, k=0)
def phase_field_density(grad_u, phi, grad_phi, mu, lmbda, Gc, ell, k):
    """Config 5: builder-defined AT2 law on top of the reference's neo-Hookean density (no counterpart in the reference)."""
    return ((1 - phi) ** 2 + k) * neo_hookean_density(grad_u, mu, lmbda) + Gc * (phi**2 / (2 * ell) + 0.5 * ell * jnp.dot(grad_phi, grad_phi))


def jitter(coords, h, seed=0):
    rng = np.random.default_rng(seed)
    return coords + 0.1 * h * rng.uniform(-1, 1, coords.shape)


def smooth_u(coords):
    x = coords
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


def hvp_probe_hi(dE, x, v, ws, d):
    """w . H v to ~1e-13: the reference energy's complex-step directional derivative dE(x)[w] (exact to rounding)
    differentiated along v by the 8th-order central difference, at step d AND d/2 — the second value is the fixture,
    their difference the recorded error estimate (truncation ~ (d / R)^8 with R the distance to det F = 0, rounding
    ~ eps |dE| / d: both far below 1e-12 for the steps used here)."""
    cf = (4.0 / 5, -1.0 / 5, 4.0 / 105, -1.0 / 280)

    def diff(w, dd):
        return sum(cf[k] * (dE(x + (k + 1) * dd * v, w) - dE(x - (k + 1) * dd * v, w)) for k in range(4)) / dd

    a = np.array([diff(w, d) for w in ws])
    b = np.array([diff(w, d / 2) for w in ws])
    return b, np.abs(a - b) / np.abs(b)


def make_case(kind):
    if kind == "tri3":
        c, el = orc.mesh_unit_square_tri(4, 3)
        c = jitter(c, 0.25)
        mat = ("linear_elastic",) + orc.lame_from_youngs_poisson_2d(1.0, 0.3)
        psi = strain_energy
    elif kind == "tet4":
        c, el = orc.mesh_box_tet((1.0, 1.0, 1.0), (2, 2, 2))
        c = jitter(c, 0.5)
        mat = ("neo_hookean", 500.0, 1000.0)
        psi = neo_hookean_density
    else:
        c, el = orc.mesh_box_hex(3)
        c = jitter(c, 1.0 / 3)
        mat = ("neo_hookean", 500.0, 1000.0)
        psi = neo_hookean_density
    return c, el, mat, psi


def element_fixtures(out):
    """tatva.element.* on one distorted element per kind."""
    rng = np.random.default_rng(7)
    for kind, cls in {**KINDS, **MORE_KINDS}.items():
        el = cls()
        X = np.asarray(el._reference_nodes(), dtype=float)
        X = X + 0.15 * rng.uniform(-1, 1, X.shape)
        dim = X.shape[1]
        uv = rng.normal(size=(X.shape[0], dim))
        us = rng.normal(size=(X.shape[0],))
        qp = np.asarray(el.quad_points, dtype=float)
        out[f"el_{kind}_X"] = X
        out[f"el_{kind}_uv"] = uv
        out[f"el_{kind}_us"] = us
        out[f"el_{kind}_qp"] = qp
        out[f"el_{kind}_qw"] = np.asarray(el.quad_weights, dtype=float)
        out[f"el_{kind}_N"] = np.stack([el.shape_function(x) for x in qp])
        out[f"el_{kind}_dNdr"] = np.stack([el.shape_function_derivative(x) for x in qp])
        out[f"el_{kind}_J"] = np.stack([el.get_jacobian(x, X)[0] for x in qp])
        out[f"el_{kind}_detJ"] = np.stack([el.get_jacobian(x, X)[1] for x in qp])
        out[f"el_{kind}_grad_v"] = np.stack([el.gradient(x, uv, X) for x in qp])
        out[f"el_{kind}_grad_s"] = np.stack([el.gradient(x, us, X) for x in qp])
        out[f"el_{kind}_interp_v"] = np.stack([el.interpolate(x, uv, X) for x in qp])


def operator_fixtures(out):
    rng = np.random.default_rng(11)
    for kind, cls in KINDS.items():
        c, el, mat, psi = make_case(kind)
        op = Operator(Mesh(coords=c, elements=el), cls())
        dim = c.shape[1]
        u = smooth_u(c) + 0.01 * rng.normal(size=c.shape)
        s = rng.normal(size=(c.shape[0],))
        v = rng.normal(size=c.shape)
        p = f"op_{kind}_"
        out[p + "coords"], out[p + "conn"] = c, el
        out[p + "u"], out[p + "s"], out[p + "v"] = u, s, v
        out[p + "mat"] = np.array(mat[1:])
        out[p + "grad_u"] = op.grad(u)
        out[p + "grad_s"] = op.grad(s)
        out[p + "eval_u"] = op.eval(u)
        out[p + "eval_s"] = op.eval(s)
        out[p + "weights"] = op.get_integration_weights()
        out[p + "int_nodal_s"] = op.integrate(s)
        out[p + "int_nodal_u_per_el"] = op.integrate_per_element(u)
        q = rng.normal(size=(el.shape[0], len(op.element.quad_points), 2))
        out[p + "quadvals"] = q
        out[p + "int_quad_per_el"] = op.integrate_per_element(q)
        # op.integrate(<python scalar>) relies on XLA clamping the out-of-bounds gather
        # jnp.array([arg])[elements] (operator.py:335); NumPy raises instead, so that branch
        # is not recorded here (the oracle restates it as arg * sum(W)).

        def E(uu):
            return op.integrate(psi(op.grad(uu), mat[1], mat[2]))

        out[p + "energy"] = E(u)
        # residual by complex step on the reference's own energy
        h = 1e-30
        r = np.zeros(u.shape)
        for n in range(u.shape[0]):
            for i in range(dim):
                uc = u.astype(complex)
                uc[n, i] += 1j * h
                r[n, i] = np.imag(E(uc)) / h
        out[p + "residual_cs"] = r
        # w . H v  for a few probe directions w:  complex step (w) x 4th-order central FD (v)
        ws = rng.normal(size=(3,) + u.shape)
        hv = np.zeros(3)
        d = 1e-4

        def dE(uu, w):
            return np.imag(E(uu.astype(complex) + 1j * h * w)) / h

        for k in range(3):
            hv[k] = (-dE(u + 2 * d * v, ws[k]) + 8 * dE(u + d * v, ws[k]) - 8 * dE(u - d * v, ws[k]) + dE(u - 2 * d * v, ws[k])) / (12 * d)
        out[p + "hvp_probe_w"] = ws
        out[p + "hvp_probe_wHv"] = hv
        out[p + "hvp_probe_wHv_hi"], out[p + "hvp_probe_wHv_hi_err"] = hvp_probe_hi(dE, u, v, ws, 2e-3)


def custom_rule_fixtures(out):
    """Element(quad_points, quad_weights) — user quadrature rules (tatva/element/base.py:37-51) — through the
    reference's Operator: Hex8 with the 3x3x3 Gauss rule, Tet4 with the 4-point degree-2 rule, Quad4 with 3x3 Gauss."""
    rng = np.random.default_rng(17)
    cases = {
        "hex8": (element.Hexahedron8, orc.gauss_rule("hex8", 3), neo_hookean_density, (500.0, 1000.0)),
        "tet4": (element.Tetrahedron4, orc.gauss_rule("tet4", 2), neo_hookean_density, (500.0, 1000.0)),
        "quad4": (element.Quad4, orc.gauss_rule("quad4", 3), strain_energy, (0.4, 0.6)),
    }
    for kind, (cls, (qp, qw), psi, prm) in cases.items():
        if kind == "quad4":
            c, el = orc.mesh_unit_square_quad(3, 3)
            c = jitter(c, 1.0 / 3)
        else:
            c, el, _, _ = make_case(kind)
        op = Operator(Mesh(coords=c, elements=el), cls(quad_points=jnp.asarray(qp), quad_weights=jnp.asarray(qw)))
        u = (smooth_u(c) if c.shape[1] == 3 else 0.05 * np.stack([np.sin(2 * np.pi * c[:, 0]) * np.cos(2 * np.pi * c[:, 1]), np.sin(2 * np.pi * c[:, 1]) * np.cos(2 * np.pi * c[:, 0])], -1)) + 0.01 * rng.normal(size=c.shape)
        s = rng.normal(size=(c.shape[0],))
        p = f"cq_{kind}_"
        out[p + "coords"], out[p + "conn"], out[p + "qp"], out[p + "qw"] = c, el, qp, qw
        out[p + "u"], out[p + "s"], out[p + "prm"] = u, s, np.array(prm)
        out[p + "grad_u"] = op.grad(u)
        out[p + "eval_s"] = op.eval(s)
        out[p + "weights"] = op.get_integration_weights()
        out[p + "int_nodal_s"] = op.integrate(s)
        q = rng.normal(size=(el.shape[0], len(qw), 2))
        out[p + "quadvals"] = q
        out[p + "int_quad_per_el"] = op.integrate_per_element(q)

        def E(uu):
            return op.integrate(psi(op.grad(uu), prm[0], prm[1]))

        out[p + "energy"] = E(u)
        h = 1e-30
        r = np.zeros(u.shape)
        for n in range(u.shape[0]):
            for i in range(u.shape[1]):
                uc = u.astype(complex)
                uc[n, i] += 1j * h
                r[n, i] = np.imag(E(uc)) / h
        out[p + "residual_cs"] = r


def more_operator_fixtures(out):
    """Operator.* on Quad4 / Tri6 / Quad8 meshes (reference Operator, shimmed)."""
    rng = np.random.default_rng(13)
    for kind, cls in MORE_KINDS.items():
        if kind == "quad4":
            c, el = orc.mesh_unit_square_quad(3, 2)
        else:
            c, el = orc.mesh_second_order(kind, 2, 2)
        c = c + 0.02 * rng.uniform(-1, 1, c.shape)
        op = Operator(Mesh(coords=c, elements=el), cls())
        u = smooth_u(c) + 0.01 * rng.normal(size=c.shape)
        s = rng.normal(size=(c.shape[0],))
        p = f"op_{kind}_"
        mat = orc.lame_from_youngs_poisson_2d(1.0, 0.3)
        out[p + "coords"], out[p + "conn"], out[p + "u"], out[p + "s"] = c, el, u, s
        out[p + "v"] = rng.normal(size=c.shape)
        out[p + "mat"] = np.array(mat)
        out[p + "grad_u"], out[p + "grad_s"] = op.grad(u), op.grad(s)
        out[p + "eval_u"], out[p + "eval_s"] = op.eval(u), op.eval(s)
        out[p + "weights"] = op.get_integration_weights()
        out[p + "int_nodal_s"] = op.integrate(s)
        out[p + "int_nodal_u_per_el"] = op.integrate_per_element(u)
        q = rng.normal(size=(el.shape[0], len(op.element.quad_points), 2))
        out[p + "quadvals"] = q
        out[p + "int_quad_per_el"] = op.integrate_per_element(q)
        out[p + "energy"] = op.integrate(strain_energy(op.grad(u), mat[0], mat[1]))

        # derivatives of the reference's own energy, as for the three main kinds: complex-step residual and
        # w . H v probes (complex step in w, 4th-order central difference in v)
        def E(uu, op=op, mat=mat):
            return op.integrate(strain_energy(op.grad(uu), mat[0], mat[1]))

        h, dim, v = 1e-30, c.shape[1], out[p + "v"]
        r = np.zeros(u.shape)
        for n in range(u.shape[0]):
            for i in range(dim):
                uc = u.astype(complex)
                uc[n, i] += 1j * h
                r[n, i] = np.imag(E(uc)) / h
        out[p + "residual_cs"] = r
        ws = np.random.default_rng(31).normal(size=(3,) + u.shape)  # own stream: the inputs above keep their draws
        d = 1e-3
        dE = lambda uu, w: np.imag(E(uu.astype(complex) + 1j * h * w)) / h  # noqa: E731
        out[p + "hvp_probe_w"] = ws
        out[p + "hvp_probe_wHv"] = np.array([(-dE(u + 2 * d * v, w) + 8 * dE(u + d * v, w) - 8 * dE(u - d * v, w) + dE(u - 2 * d * v, w)) / (12 * d) for w in ws])
        out[p + "hvp_probe_wHv_hi"], out[p + "hvp_probe_wHv_hi_err"] = hvp_probe_hi(dE, u, v, ws, 2e-3)


def line_meshes():
    """Curved boundary polylines in the plane: a quarter arc of radius 1.3 (Line2: chords; Line3: end nodes then an
    off-chord midpoint, so |dX/dxi| varies along the element)."""
    n = 7
    t = np.linspace(0.1, 1.4, n + 1)
    ends = 1.3 * np.stack([np.cos(t), np.sin(t)], -1)
    tm = 0.5 * (t[:-1] + t[1:]) + 0.03
    mids = 1.3 * np.stack([np.cos(tm), np.sin(tm)], -1)
    line2 = (ends, np.stack([np.arange(n), np.arange(1, n + 1)], -1).astype(np.int32))
    line3 = (np.concatenate([ends, mids]), np.stack([np.arange(n), np.arange(1, n + 1), n + 1 + np.arange(n)], -1).astype(np.int32))
    return {"line2": line2, "line3": line3}


def line_fixtures(out):
    """tatva.element.Line2 / Line3 and tatva.Operator on line meshes (element/base.py:144-242)."""
    rng = np.random.default_rng(17)
    for kind, cls in {"line2": element.Line2, "line3": element.Line3}.items():
        c, el = line_meshes()[kind]
        op = Operator(Mesh(coords=c, elements=el), cls())
        u = rng.normal(size=c.shape)
        s = rng.normal(size=(c.shape[0],))
        p = f"op_{kind}_"
        out[p + "coords"], out[p + "conn"], out[p + "u"], out[p + "s"] = c, el, u, s
        out[p + "qp"], out[p + "qw"] = np.asarray(op.element.quad_points, dtype=float), np.asarray(op.element.quad_weights, dtype=float)
        out[p + "grad_u"] = op.grad(u)
        out[p + "grad_s"] = op.grad(s)
        out[p + "eval_u"] = op.eval(u)
        out[p + "weights"] = op.get_integration_weights()
        out[p + "int_nodal_s"] = op.integrate(s)
        q = rng.normal(size=(el.shape[0], len(op.element.quad_points), 2))
        out[p + "quadvals"] = q
        out[p + "int_quad_per_el"] = op.integrate_per_element(q)


def interpolate_fixtures(out):
    """Operator.interpolate (operator.py:399-463) and mesh.find_containing_polygons (mesh.py:294-388) on jittered
    Tri3 and Quad4 meshes: interior points, points on shared edges and on mesh nodes, and points outside."""
    from tatva.mesh import find_containing_polygons

    rng = np.random.default_rng(23)
    for kind, cls in {"tri3": element.Tri3, "quad4": element.Quad4, "tri6": element.Tri6, "quad8": element.Quad8}.items():
        if kind == "tri3":
            c, el = orc.mesh_unit_square_tri(5, 4)
        elif kind == "quad4":
            c, el = orc.mesh_unit_square_quad(4, 5)
        else:  # second-order elements: the polygon is the node loop in connectivity order, the Newton step is not exact
            c, el = orc.mesh_second_order(kind, 3, 3)
        interior = (c[:, 0] > 1e-9) & (c[:, 0] < 1 - 1e-9) & (c[:, 1] > 1e-9) & (c[:, 1] < 1 - 1e-9)
        c = c + 0.04 * rng.uniform(-1, 1, c.shape) * interior[:, None]
        pts = rng.uniform(0.02, 0.98, size=(40, 2))
        edge_mid = 0.5 * (c[el[::3, 0]] + c[el[::3, 1]])  # on an edge shared by two elements
        nodes = c[[0, 7, len(c) - 1]]
        inside = np.concatenate([pts, edge_mid, nodes])
        outside = np.array([[1.5, 0.5], [-0.2, 0.3], [0.5, 1.0001]])
        if kind in ("tri6", "quad8"):
            # keep only the points the reference itself locates (its node-loop polygons of second-order elements are
            # self-intersecting, so some interior points are in no polygon); Operator.interpolate raises otherwise
            found = np.asarray(find_containing_polygons(inside, c[el])) >= 0
            inside = inside[found]
        u = rng.normal(size=(c.shape[0], 3))
        s = rng.normal(size=(c.shape[0],))
        p = f"interp_{kind}_"
        out[p + "coords"], out[p + "conn"], out[p + "u"], out[p + "s"] = c, el, u, s
        out[p + "points"], out[p + "outside"] = inside, outside
        out[p + "containing"] = np.asarray(find_containing_polygons(np.concatenate([inside, outside]), c[el]))
        op = Operator(Mesh(coords=c, elements=el), cls())
        out[p + "values_u"] = np.asarray(op.interpolate(u, inside))
        out[p + "values_s"] = np.asarray(op.interpolate(s, inside))


def sparse_fixtures(out):
    cases = {
        "tri3_8x8_d2": (orc.mesh_unit_square_tri(8, 8), 2),  # tests/test_sparse.py:40-45
        "tet4_3_d3": (orc.mesh_box_tet((1, 1, 1), (3, 3, 3)), 3),
        "tet4_2_d4": (orc.mesh_box_tet((1, 1, 1), (2, 2, 2)), 4),
        "hex8_3_d3": (orc.mesh_box_hex(3), 3),
    }
    for name, ((c, el), dpn) in cases.items():
        pat = sparse.pattern_from_mesh(Mesh(coords=c, elements=el), dpn)
        colors = np.asarray(distance2_colors(pat.indptr, pat.indices, pat.shape[0]))
        out[f"sp_{name}_conn"] = el
        out[f"sp_{name}_nnodes"] = np.array(c.shape[0])
        out[f"sp_{name}_indptr"] = pat.indptr
        out[f"sp_{name}_indices"] = pat.indices
        out[f"sp_{name}_colors"] = colors
        assert pat.indptr.dtype == np.int32 and pat.indices.dtype == np.int32


def partition_fixtures(out):
    c, el = orc.mesh_box_hex((4, 2, 2))
    cx = c[el].mean(axis=1)[:, 0]
    part = (cx > cx.mean()).astype(np.int32)
    out["part_hex_coords"], out["part_hex_conn"], out["part_hex_partition"] = c, el, part
    for r in range(2):
        m, info = extract_local_mesh(Mesh(coords=c, elements=el), part, r)
        out[f"part_hex_r{r}_coords"] = np.asarray(m.coords)
        out[f"part_hex_r{r}_conn"] = np.asarray(m.elements)
        out[f"part_hex_r{r}_l2g"] = np.asarray(info.nodes_local_to_global)
        out[f"part_hex_r{r}_nowned"] = np.array(info.n_owned_nodes)
    c, el = orc.mesh_unit_square_tri(4, 4)
    cen = c[el].mean(axis=1)
    part = ((cen[:, 0] > 0.5).astype(np.int32) + 2 * (cen[:, 1] > 0.5).astype(np.int32)).astype(np.int32)
    out["part_tri_coords"], out["part_tri_conn"], out["part_tri_partition"] = c, el, part
    for r in range(4):
        m, info = extract_local_mesh(Mesh(coords=c, elements=el), part, r)
        out[f"part_tri_r{r}_conn"] = np.asarray(m.elements)
        out[f"part_tri_r{r}_l2g"] = np.asarray(info.nodes_local_to_global)
        out[f"part_tri_r{r}_nowned"] = np.array(info.n_owned_nodes)


def phase_field_fixtures(out):
    """The two-field energy of config 5 evaluated with the REFERENCE's Operator (grad of u, eval and grad of phi,
    integrate) on the compound-stacked state [ux, uy, uz, phi], with complex-step derivatives.  The density itself is
    ours (parity unpinned as a law); what this pins is everything around it: quadrature, gather, scatter, layout."""
    rng = np.random.default_rng(37)
    prm = (500.0, 1000.0, 2.7, 0.1, 1e-6)
    for kind, cls in {"tet4": element.Tetrahedron4, "hex8": element.Hexahedron8}.items():
        c, el = (orc.mesh_box_tet((1.0, 1.0, 1.0), (2, 2, 2)) if kind == "tet4" else orc.mesh_box_hex(2))
        c = c + 0.05 * rng.uniform(-1, 1, c.shape)
        op = Operator(Mesh(coords=c, elements=el), cls())
        st = np.concatenate([0.02 * rng.normal(size=c.shape), rng.uniform(0.0, 0.8, size=(c.shape[0], 1))], axis=1)
        t = rng.normal(size=st.shape)

        def E(ss, op=op):
            return op.integrate(phase_field_density(op.grad(ss[:, :3]), op.eval(ss[:, 3]), op.grad(ss[:, 3]), *prm))

        p = f"pf_{kind}_"
        out[p + "coords"], out[p + "conn"], out[p + "s"], out[p + "t"], out[p + "params"] = c, el, st, t, np.array(prm)
        out[p + "energy"] = E(st)
        h = 1e-30
        r = np.zeros(st.shape)
        for n in range(st.shape[0]):
            for i in range(4):
                sc = st.astype(complex)
                sc[n, i] += 1j * h
                r[n, i] = np.imag(E(sc)) / h
        out[p + "residual_cs"] = r
        ws = rng.normal(size=(3,) + st.shape)
        d = 1e-4
        dE = lambda ss, w: np.imag(E(ss.astype(complex) + 1j * h * w)) / h  # noqa: E731
        out[p + "hvp_probe_w"] = ws
        out[p + "hvp_probe_wHv"] = np.array([(-dE(st + 2 * d * t, w) + 8 * dE(st + d * t, w) - 8 * dE(st - d * t, w) + dE(st - 2 * d * t, w)) / (12 * d) for w in ws])
        out[p + "hvp_probe_wHv_hi"], out[p + "hvp_probe_wHv_hi_err"] = hvp_probe_hi(dE, st, t, ws, 1e-3)


def colored_jacobian_fixtures(out):
    """sparse.jacfwd / colored_jacobian_batch / compute_rows_cols (sparse/base.py:108-176, :230-270) on the Tri3 8x8
    two-DOF pattern with the reference colouring, for an analytic residual whose Jacobian has exactly that pattern:
    fn(u) = A u + 0.1 (A u)^2.  The shim's jax.jvp is a complex-step derivative, exact for this fn."""
    import scipy.sparse as sps

    rng = np.random.default_rng(41)
    c, el = orc.mesh_unit_square_tri(8, 8)
    mesh = Mesh(coords=jnp.asarray(c), elements=jnp.asarray(el))
    pat = sparse.pattern_from_mesh(mesh, 2)
    cm = sparse.ColoredMatrix.from_csr(pat)
    A = sps.csr_matrix((rng.normal(size=pat.indices.shape[0]), np.asarray(pat.indices), np.asarray(pat.indptr)), shape=pat.shape)
    u0 = rng.normal(size=pat.shape[0])

    def fn(u):
        y = A @ u
        return y + 0.1 * y * y

    for batch in (None, 5):
        K = sparse.jacfwd(fn, cm, color_batch_size=batch)(u0)
        out[f"jac_data_batch{batch}"] = np.asarray(K.data)
    out["jac_A_data"], out["jac_u0"] = A.data, u0
    out["jac_indptr"], out["jac_indices"], out["jac_colors"] = np.asarray(cm.indptr), np.asarray(cm.indices), np.asarray(cm.colors)
    rows, col_colors = sparse.base.compute_rows_cols(cm) if hasattr(sparse, "base") else (None, None)
    if rows is not None:
        out["jac_rows"], out["jac_col_colors"] = np.asarray(rows), np.asarray(col_colors)


def lifter_fixtures(out):
    """tatva.lifter (Lifter, Fixed, Periodic, RuntimeValue, lifted; lifter/base.py:201-251, constraints.py:184-318) run
    unmodified on a chain of constraints over the DOFs of a Tri3 6x6 mesh with 2 DOFs per node: a fixed bottom edge,
    a prescribed (runtime) top displacement, left/right periodicity.  Records free DOFs, lift / lift_from_zeros /
    reduce / reduce_adjoint of random vectors, and the lifted-decorator outputs."""
    from tatva.lifter import Fixed, Lifter, Periodic, RuntimeValue, lifted

    rng = np.random.default_rng(43)
    c, el = orc.mesh_unit_square_tri(6, 6)
    n = 2 * c.shape[0]
    bottom = np.where(c[:, 1] < 1e-12)[0]
    top = np.where(c[:, 1] > 1 - 1e-12)[0]
    inner = (c[:, 1] > 1e-12) & (c[:, 1] < 1 - 1e-12)
    left = np.where((c[:, 0] < 1e-12) & inner)[0]
    right = np.where((c[:, 0] > 1 - 1e-12) & inner)[0]
    left, right = left[np.argsort(c[left, 1])], right[np.argsort(c[right, 1])]
    dofs = lambda nodes: (np.asarray(nodes)[:, None] * 2 + np.arange(2)).ravel()  # noqa: E731
    lifter = Lifter(
        n,
        Fixed(jnp.asarray(dofs(bottom)), 0.0),
        Fixed(jnp.asarray(top * 2 + 1), RuntimeValue("top_uy")),
        Fixed(jnp.asarray(top * 2), jnp.asarray(0.01 * np.arange(len(top)))),
        Periodic(dofs=jnp.asarray(dofs(right)), master_dofs=jnp.asarray(dofs(left))),
    )
    lifter = lifter.with_values({"top_uy": 0.07})
    out["lift_n"], out["lift_bottom"], out["lift_top"], out["lift_left"], out["lift_right"] = np.array(n), bottom, top, left, right
    out["lift_free_dofs"] = np.asarray(lifter.free_dofs)
    u_red = rng.normal(size=lifter.size_reduced)
    base = rng.normal(size=n)
    r_full = rng.normal(size=n)
    out["lift_u_red"], out["lift_base"], out["lift_r_full"] = u_red, base, r_full
    out["lift_from_zeros"] = np.asarray(lifter.lift_from_zeros(jnp.asarray(u_red)))
    out["lift_on_base"] = np.asarray(lifter.lift(jnp.asarray(u_red), jnp.asarray(base)))
    out["lift_reduce"] = np.asarray(lifter.reduce(jnp.asarray(base)))
    out["lift_reduce_adjoint"] = np.asarray(lifter.reduce_adjoint(jnp.asarray(r_full)))
    A = rng.normal(size=(n, n))
    out["lift_A"] = A
    out["lifted_dual"] = np.asarray(lifted(lambda uf: jnp.asarray(A @ uf), argnums=0, output="dual")(lifter, jnp.asarray(u_red)))
    out["lifted_primal"] = np.asarray(lifted(lambda uf: uf * 2.0, argnums=0, output="primal")(lifter, jnp.asarray(u_red)))
    # chains: a master that is itself a slave of the same Periodic, and a master fixed by a LATER constraint
    chain = Lifter(8, Periodic(dofs=jnp.asarray([1, 2]), master_dofs=jnp.asarray([0, 1])), Periodic(dofs=jnp.asarray([5]), master_dofs=jnp.asarray([6])), Fixed(jnp.asarray([6]), 3.0))
    cu, cb, cr = np.array([10.0, 11.0, 12.0, 13.0]), np.arange(100.0, 108.0), np.arange(1.0, 9.0)
    out["lift_chain_free_dofs"] = np.asarray(chain.free_dofs)
    out["lift_chain_on_base"] = np.asarray(chain.lift(jnp.asarray(cu), jnp.asarray(cb)))
    out["lift_chain_from_zeros"] = np.asarray(chain.lift_from_zeros(jnp.asarray(cu)))
    out["lift_chain_reduce_adjoint"] = np.asarray(chain.reduce_adjoint(jnp.asarray(cr)))
    # sparsity adaptation (lifter/base.py:281-331, constraints.py:195-212): augmented by the periodic coupling, then
    # reduced to the free DOFs
    pat = sparse.pattern_from_mesh(Mesh(coords=jnp.asarray(c), elements=jnp.asarray(el)), 2)
    for name, m in (("augmented", lifter.augment_sparsity(pat)), ("adapted", lifter.adapt_sparsity(pat))):
        m = m.tocsr()
        m.sort_indices()
        out[f"lift_sp_{name}_indptr"], out[f"lift_sp_{name}_indices"] = np.asarray(m.indptr), np.asarray(m.indices)


def compound_pattern_fixtures(out):
    """sparse.pattern_from_compound (sparse/_extraction.py:118-245) of the unmodified reference for a mixed layout: a full
    nodal field, a nodal field on a node subset and a shared field (diagonal only), flat and block-wise."""
    import warnings

    from tatva.compound import Compound, FieldSize, field
    from tatva.compound.field_types import Nodal, Shared

    c, el = orc.mesh_unit_square_tri(5, 4)
    mesh = Mesh(coords=jnp.asarray(c), elements=jnp.asarray(el))
    sub = np.array([0, 3, 7, 8, 20])

    class Mixed(Compound, mesh=mesh):
        u = field(shape=(FieldSize.AUTO, 2))
        lam = field(shape=(FieldSize.AUTO, 1), field_type=Nodal(node_ids=jnp.asarray(sub)))
        g = field(shape=(3,), field_type=Shared())

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pat = sparse.pattern_from_compound(Mixed).tocsr()
        blocks = sparse.pattern_from_compound(Mixed, block_wise=True)
    pat.sort_indices()
    out["cpat_subset"], out["cpat_size"] = sub, np.array(Mixed.size)
    out["cpat_indptr"], out["cpat_indices"] = np.asarray(pat.indptr), np.asarray(pat.indices)
    out["cpat_block_grid"] = np.array([len(blocks), len(blocks[0])])
    for i, row in enumerate(blocks):
        for j, b in enumerate(row):
            b = b.tocsr()
            b.sort_indices()
            out[f"cpat_block_{i}{j}_shape"] = np.array(b.shape)
            out[f"cpat_block_{i}{j}_indptr"], out[f"cpat_block_{i}{j}_indices"] = np.asarray(b.indptr), np.asarray(b.indices)


def mesh_size_fixtures(out):
    """Mesh.hmin / hmax / _element_circumdiameters (mesh.py:87-144) on jittered meshes of every branch: triangles in
    2-D and embedded in 3-D, tetrahedra, and the max-vertex-distance fallback (quads, hexes)."""
    rng = np.random.default_rng(29)
    ct, et = orc.mesh_unit_square_tri(4, 3)
    ct = ct + 0.03 * rng.uniform(-1, 1, ct.shape)
    ct3 = np.concatenate([ct, 0.2 * np.sin(3 * ct[:, :1]) + 0.1 * ct[:, 1:2]], axis=1)
    cq, eq = orc.mesh_unit_square_quad(3, 4)
    cq = cq + 0.03 * rng.uniform(-1, 1, cq.shape)
    cT, eT = orc.mesh_box_tet((1.0, 0.7, 0.4), (3, 2, 2))
    cT = cT + 0.03 * rng.uniform(-1, 1, cT.shape)
    cH, eH = orc.mesh_box_hex((3, 2, 2))
    cH = cH + 0.03 * rng.uniform(-1, 1, cH.shape)
    for name, (c, el) in {"tri2d": (ct, et), "tri3d": (ct3, et), "quad": (cq, eq), "tet": (cT, eT), "hex": (cH, eH)}.items():
        m = Mesh(coords=jnp.asarray(c), elements=jnp.asarray(el))
        out[f"h_{name}_coords"], out[f"h_{name}_conn"] = c, el
        out[f"h_{name}_diam"] = np.asarray(m._element_circumdiameters())
        out[f"h_{name}_hmin"], out[f"h_{name}_hmax"] = np.asarray(m.hmin()), np.asarray(m.hmax())


def main():
    out: dict[str, np.ndarray] = {}
    element_fixtures(out)
    operator_fixtures(out)
    more_operator_fixtures(out)
    custom_rule_fixtures(out)
    sparse_fixtures(out)
    partition_fixtures(out)
    line_fixtures(out)
    interpolate_fixtures(out)
    mesh_size_fixtures(out)
    phase_field_fixtures(out)
    colored_jacobian_fixtures(out)
    lifter_fixtures(out)
    compound_pattern_fixtures(out)
    from _fakempi_golden import mpi_fixtures

    mpi_fixtures(out)
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
