"""CPU oracle: a NumPy restatement of tatva's element-level hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tatva_b200/`` imports this module; it may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs, and there only as the checker.  All paths below are relative to ``/root/reference``
(tatva v0.11.1).

Parity status (see DESIGN.md "Oracle"):
  * element maths, Operator.grad/eval/integrate/weights, pattern_from_mesh, the greedy
    distance-2 colouring, extract_local_mesh, _create_dof_layout and the ExchangePlan
    routing are PINNED: ``tests/golden/make_golden.py`` runs the unmodified reference
    modules (on a NumPy stand-in for the absent ``jax`` package) and the fixtures it wrote
    are compared against this file in ``tests/test_oracle_golden.py``, together with the
    known-answer values of the reference's own tests (tests/test_operator.py:113-143,
    tests/test_element.py:45-149, tests/test_exchange_plan.py, tests/test_allreduce_plan.py).
  * residual / HVP: the reference has no code for these (they are ``jax.grad`` /
    ``jax.jvp`` of a user energy, README.md:93).  The closed forms here are pinned to
    (i) complex-step derivatives of the *reference's own* ``Operator`` energy (fixtures)
    and (ii) ``torch.func.jvp(torch.func.grad(E))`` of a vmap-structured energy in
    tests/test_oracle_autodiff.py.
  * exact colour ids: the PyPI package ``tatva-coloring>=0.0.2`` (pyproject.toml:24) is a
    third-party dependency absent from /root/reference; we restate its in-tree predecessor
    tatva/sparse/_coloring.py:27-48,136-153,270-283 — "parity unpinned" against the PyPI
    package itself.
  * phase-field (AT2) energy of config 5: no such energy exists in the reference —
    "parity unpinned", defined here and checked by autodiff only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

# ----------------------------------------------------------------------------------------
# Elements (tatva/element/base.py)
# ----------------------------------------------------------------------------------------

_A = 1.0 / np.sqrt(3.0)

_HEX_SIGNS = np.array(
    [
        [-1.0, -1.0, -1.0],
        [1.0, -1.0, -1.0],
        [1.0, 1.0, -1.0],
        [-1.0, 1.0, -1.0],
        [-1.0, -1.0, 1.0],
        [1.0, -1.0, 1.0],
        [1.0, 1.0, 1.0],
        [-1.0, 1.0, 1.0],
    ]
)  # element/base.py:478-491 (reference nodes) and :496-507 (quad points = a * signs)

ELEMENT_INFO = {
    # kind: (dim, nodes per element, n quad points)
    "tri3": (2, 3, 1),
    "tet4": (3, 4, 1),
    "hex8": (3, 8, 8),
    "quad4": (2, 4, 4),
    "tri6": (2, 6, 3),
    "quad8": (2, 8, 9),
    # line elements embedded in the plane: dim = width of a coordinate row (the gradient has 1 component)
    "line2": (2, 2, 1),
    "line3": (2, 3, 3),
}
_LINES = ("line2", "line3")

_Q4_SIGNS = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])  # element/base.py:334-336
_B = np.sqrt(3.0 / 5.0)


def reference_nodes(kind: str) -> np.ndarray:
    if kind == "tri3":  # element/base.py:248-250
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    if kind == "tet4":  # element/base.py:451-455
        return np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    if kind == "hex8":  # element/base.py:478-491
        return _HEX_SIGNS.copy()
    if kind == "quad4":  # element/base.py:334-336
        return _Q4_SIGNS.copy()
    if kind == "tri6":  # element/base.py:272-276
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.0], [0.5, 0.5], [0.0, 0.5]])
    if kind == "quad8":  # element/base.py:369-382
        return np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0], [0.0, -1.0], [1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    if kind == "line2":  # element/base.py:147-149
        return np.array([[-1.0], [1.0]])
    if kind == "line3":  # element/base.py:192-194
        return np.array([[-1.0], [1.0], [0.0]])
    raise ValueError(kind)


_RULE_OVERRIDE: dict = {}


class custom_rule:
    """Context manager: run the oracle with a user quadrature rule for `kind`, as Element(quad_points, quad_weights)
    does in the reference (element/base.py:37-51): `with custom_rule("hex8", points, weights): ...`."""

    def __init__(self, kind, points, weights):
        self.kind, self.rule = kind, (np.asarray(points, dtype=np.float64), np.asarray(weights, dtype=np.float64))

    def __enter__(self):
        self.prev = _RULE_OVERRIDE.get(self.kind)
        _RULE_OVERRIDE[self.kind] = self.rule
        return self

    def __exit__(self, *exc):
        if self.prev is None:
            _RULE_OVERRIDE.pop(self.kind, None)
        else:
            _RULE_OVERRIDE[self.kind] = self.prev
        return False


def gauss_rule(kind: str, order: int) -> tuple[np.ndarray, np.ndarray]:
    """Tensor Gauss-Legendre rules with `order` points per direction for hex8 / quad4 (x fastest, like the reference's
    defaults), and the 4-point degree-2 rule for tet4 (order = 2): the rules of the custom-quadrature tests."""
    if kind in ("hex8", "quad4"):
        x, w = np.polynomial.legendre.leggauss(order)
        if kind == "quad4":
            return np.array([[x[j], x[i]] for i in range(order) for j in range(order)]), np.array([w[i] * w[j] for i in range(order) for j in range(order)])
        idx = [(i, j, k) for k in range(order) for j in range(order) for i in range(order)]
        return np.array([[x[i], x[j], x[k]] for i, j, k in idx]), np.array([w[i] * w[j] * w[k] for i, j, k in idx])
    if kind == "tet4" and order == 2:
        a, b = 0.58541019662496845446, 0.13819660112501051518
        return np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]]), np.full(4, 1.0 / 24)
    if kind == "tri3" and order == 2:
        return np.array([[1.0 / 6, 1.0 / 6], [2.0 / 3, 1.0 / 6], [1.0 / 6, 2.0 / 3]]), np.full(3, 1.0 / 6)
    raise ValueError((kind, order))


def quad_rule(kind: str) -> tuple[np.ndarray, np.ndarray]:
    """Default quadrature (points (Q,dim), weights (Q,)), or the rule installed by `custom_rule`."""
    if kind in _RULE_OVERRIDE:
        return _RULE_OVERRIDE[kind]
    if kind == "tri3":  # element/base.py:252-255
        return np.array([[1.0 / 3, 1.0 / 3]]), np.array([1.0 / 2])
    if kind == "tet4":  # element/base.py:457-460
        return np.array([[1.0 / 4, 1.0 / 4, 1.0 / 4]]), np.array([1.0 / 6])
    if kind == "hex8":  # element/base.py:493-513
        return _A * _HEX_SIGNS, np.ones(8)
    if kind == "quad4":  # element/base.py:338-344: meshgrid('xy') -> x fastest
        x = np.array([-_A, _A])
        return np.array([[x[j], x[i]] for i in range(2) for j in range(2)]), np.ones(4)
    if kind == "tri6":  # element/base.py:278-284
        return np.array([[1.0 / 6, 1.0 / 6], [2.0 / 3, 1.0 / 6], [1.0 / 6, 2.0 / 3]]), np.full(3, 1.0 / 6)
    if kind == "quad8":  # element/base.py:384-393
        x, w = np.array([-_B, 0.0, _B]), np.array([5.0 / 9, 8.0 / 9, 5.0 / 9])
        return np.array([[x[j], x[i]] for i in range(3) for j in range(3)]), np.kron(w, w)
    if kind == "line2":  # element/base.py:151-154
        return np.array([[0.0]]), np.array([2.0])
    if kind == "line3":  # element/base.py:196-199
        return np.array([[-_B], [0.0], [_B]]), np.array([5.0 / 9, 8.0 / 9, 5.0 / 9])
    raise ValueError(kind)


def shape_function(kind: str, xi: np.ndarray) -> np.ndarray:
    """N(xi), shape (npe,)."""
    if kind == "tri3":  # element/base.py:257-260
        return np.array([1.0 - xi[0] - xi[1], xi[0], xi[1]])
    if kind == "tet4":  # element/base.py:462-465
        return np.array([1.0 - xi[0] - xi[1] - xi[2], xi[0], xi[1], xi[2]])
    if kind == "hex8":  # element/base.py:515-529
        s = _HEX_SIGNS
        return 0.125 * (1 + s[:, 0] * xi[0]) * (1 + s[:, 1] * xi[1]) * (1 + s[:, 2] * xi[2])
    if kind == "quad4":  # element/base.py:346-350
        return 0.25 * (1 + _Q4_SIGNS[:, 0] * xi[0]) * (1 + _Q4_SIGNS[:, 1] * xi[1])
    if kind == "tri6":  # element/base.py:286-298
        r, t_, = xi[0], xi[1]
        t = 1.0 - r - t_
        return np.array([t * (2 * t - 1), r * (2 * r - 1), t_ * (2 * t_ - 1), 4 * r * t, 4 * r * t_, 4 * t_ * t])
    if kind == "quad8":  # element/base.py:395-407
        r, t = xi
        return np.array(
            [
                0.25 * (1 - r) * (1 - t) * (-r - t - 1),
                0.25 * (1 + r) * (1 - t) * (r - t - 1),
                0.25 * (1 + r) * (1 + t) * (r + t - 1),
                0.25 * (1 - r) * (1 + t) * (-r + t - 1),
                0.5 * (1 - r * r) * (1 - t),
                0.5 * (1 + r) * (1 - t * t),
                0.5 * (1 - r * r) * (1 + t),
                0.5 * (1 - r) * (1 - t * t),
            ]
        )
    if kind == "line2":  # element/base.py:156-158
        return np.array([0.5 * (1.0 - xi[0]), 0.5 * (1.0 + xi[0])])
    if kind == "line3":  # element/base.py:201-207
        r = xi[0]
        return np.array([0.5 * r * (r - 1.0), 0.5 * r * (r + 1.0), 1.0 - r * r])
    raise ValueError(kind)


def shape_function_derivative(kind: str, xi: np.ndarray) -> np.ndarray:
    """dN/dxi, shape (dim, npe)."""
    if kind == "tri3":  # element/base.py:262-265
        return np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]).T
    if kind == "tet4":  # element/base.py:467-472
        return np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]).T
    if kind == "hex8":  # element/base.py:531-568
        s = _HEX_SIGNS
        fx, fy, fz = 1 + s[:, 0] * xi[0], 1 + s[:, 1] * xi[1], 1 + s[:, 2] * xi[2]
        return 0.125 * np.stack([s[:, 0] * fy * fz, s[:, 1] * fx * fz, s[:, 2] * fx * fy])
    if kind == "quad4":  # element/base.py:352-366
        s = _Q4_SIGNS
        return 0.25 * np.stack([s[:, 0] * (1 + s[:, 1] * xi[1]), s[:, 1] * (1 + s[:, 0] * xi[0])])
    if kind == "tri6":  # element/base.py:300-328
        r, q = xi[0], xi[1]
        t = 1.0 - r - q
        return np.array(
            [
                [-(4 * t - 1), 4 * r - 1, 0.0, 4 * (t - r), 4 * q, -4 * q],
                [-(4 * t - 1), 0.0, 4 * q - 1, -4 * r, 4 * r, 4 * (t - q)],
            ]
        )
    if kind == "quad8":  # element/base.py:409-445
        r, t = xi
        dr = [0.25 * (-2 * r - t) * (t - 1), 0.25 * (-2 * r + t) * (t - 1), 0.25 * (2 * r + t) * (t + 1), 0.25 * (2 * r - t) * (t + 1), r * (t - 1), 0.5 - 0.5 * t * t, -r * (t + 1), 0.5 * t * t - 0.5]
        dt = [0.25 * (-r - 2 * t) * (r - 1), 0.25 * (-r + 2 * t) * (r + 1), 0.25 * (r + 1) * (r + 2 * t), 0.25 * (r - 1) * (r - 2 * t), 0.5 * r * r - 0.5, -t * (r + 1), 0.5 - 0.5 * r * r, t * (r - 1)]
        return np.array([dr, dt])
    if kind == "line2":  # element/base.py:160-162 (returned here as (1, npe))
        return np.array([[-0.5, 0.5]])
    if kind == "line3":  # element/base.py:209-214
        r = xi[0]
        return np.array([[r - 0.5, r + 0.5, -2.0 * r]])
    raise ValueError(kind)


def get_jacobian(kind: str, xi, X_e):
    """element/base.py:90-93: J = dNdr @ X_e, det J.  Line elements (:163-167, :216-222): the arc-length
    derivative dot(Jvec, Jvec / |Jvec|), returned twice."""
    if kind in _LINES:
        Jvec = shape_function_derivative(kind, xi)[0] @ X_e
        J = np.dot(Jvec, Jvec / np.linalg.norm(Jvec))
        return J, J
    J = shape_function_derivative(kind, xi) @ X_e
    return J, np.linalg.det(J)


def element_interpolate(kind: str, xi, u_e, X_e=None):
    """element/base.py:95-97."""
    return np.einsum("n,n...->...", shape_function(kind, xi), u_e)


def element_gradient(kind: str, xi, u_e, X_e):
    """element/base.py:99-115: value dims first, spatial dim last.  Line elements (:169-173, :224-228): dNdr / J,
    no spatial axis."""
    if kind in _LINES:
        J, _ = get_jacobian(kind, xi, X_e)
        return np.einsum("n,n...->...", shape_function_derivative(kind, xi)[0] / J, u_e)
    dNdr = shape_function_derivative(kind, xi)
    J = dNdr @ X_e
    dNdX = np.linalg.inv(J) @ dNdr
    return np.einsum("dn,n...->...d", dNdX, u_e)


# ----------------------------------------------------------------------------------------
# Operator (tatva/operator.py) — vectorised over (E, Q)
# ----------------------------------------------------------------------------------------


def _dNdr_all(kind: str) -> np.ndarray:
    qp, _ = quad_rule(kind)
    return np.stack([shape_function_derivative(kind, x) for x in qp])  # (Q, d, n)


def _N_all(kind: str) -> np.ndarray:
    qp, _ = quad_rule(kind)
    return np.stack([shape_function(kind, x) for x in qp])  # (Q, n)


def geometry(kind: str, coords: np.ndarray, conn: np.ndarray):
    """Per (e,q): dNdX (E,Q,d,n) and detJ (E,Q)   [element/base.py:107-114, :90-93]."""
    dNdr = _dNdr_all(kind)
    X_e = coords[conn]  # operator.py:221 gather
    if kind in _LINES:  # dNdX (E,Q,1,n) = dNdr / |dX/dxi|, "detJ" = |dX/dxi|
        Jvec = np.einsum("qn,enc->eqc", dNdr[:, 0, :], X_e)
        J = np.einsum("eqc,eqc->eq", Jvec, Jvec / np.linalg.norm(Jvec, axis=-1, keepdims=True))
        return dNdr[None, :, :, :] / J[:, :, None, None], J
    J = np.einsum("qdn,enc->eqdc", dNdr, X_e)
    detJ = np.linalg.det(J)
    dNdX = np.einsum("eqdc,qcn->eqdn", np.linalg.inv(J), dNdr)
    return dNdX, detJ


def op_integration_weights(kind, coords, conn):
    """operator.py:172-192: W[e,q] = det(J) * w_q   (no abs)."""
    _, w = quad_rule(kind)
    _, detJ = geometry(kind, coords, conn)
    return detJ * w[None, :]


def op_grad(kind, coords, conn, u):
    """operator.py:379-397 -> (E,Q,*value_dims,d)."""
    dNdX, _ = geometry(kind, coords, conn)
    g = np.einsum("eqdn,en...->eq...d", dNdX, u[conn])
    return g[..., 0] if kind in _LINES else g


def op_eval(kind, coords, conn, u):
    """operator.py:358-377 -> (E,Q,*value_dims)."""
    return np.einsum("qn,en...->eq...", _N_all(kind), u[conn])


def op_integrate_per_element(kind, coords, conn, arg):
    """operator.py:321-356 including its dispatch rule on arg.shape[0]."""
    E = conn.shape[0]
    W = op_integration_weights(kind, coords, conn)
    if np.isscalar(arg):
        # operator.py:335 evals jnp.array([arg]); XLA clamps the out-of-bounds gather, so every
        # node reads `arg` and the interpolant is the constant (shape (E,Q)).
        vals = np.full(W.shape, float(arg)) * _N_all(kind).sum(axis=1)[None, :]
    elif arg.shape[0] == E:
        vals = arg
    else:
        vals = op_eval(kind, coords, conn, arg)
    return np.einsum("eq...,eq->e...", vals, W)


def op_integrate(kind, coords, conn, arg):
    """operator.py:307-319."""
    return np.sum(op_integrate_per_element(kind, coords, conn, arg), axis=0)


def find_containing_polygons(points, polygons):
    """mesh.py:294-388: for every 2-D point the FIRST polygon (vertex loops in the given order) that contains it, -1
    if none.  Bounding-box reject, then boundary test (|cross| <= 1e-8, jnp.isclose's atol, on the segment's box)
    OR odd number of crossings of the +x ray."""
    points, polygons = np.asarray(points, dtype=np.float64), np.asarray(polygons, dtype=np.float64)
    out = np.full(points.shape[0], -1, dtype=np.int64)
    lo, hi = polygons.min(axis=1), polygons.max(axis=1)
    p1, p2 = polygons, np.roll(polygons, -1, axis=1)
    for i, (px, py) in enumerate(points):
        cand = np.where((px >= lo[:, 0]) & (px <= hi[:, 0]) & (py >= lo[:, 1]) & (py <= hi[:, 1]))[0]
        for j in cand:
            a, b = p1[j], p2[j]
            cross = (b[:, 0] - a[:, 0]) * (py - a[:, 1]) - (b[:, 1] - a[:, 1]) * (px - a[:, 0])
            on_seg = (np.minimum(a[:, 0], b[:, 0]) <= px) & (px <= np.maximum(a[:, 0], b[:, 0])) & (np.minimum(a[:, 1], b[:, 1]) <= py) & (py <= np.maximum(a[:, 1], b[:, 1]))
            on_boundary = np.any((np.abs(cross) <= 1e-8) & on_seg)
            y_cond = ((a[:, 1] <= py) & (b[:, 1] > py)) | ((b[:, 1] <= py) & (a[:, 1] > py))
            with np.errstate(divide="ignore", invalid="ignore"):
                x_int = (b[:, 0] - a[:, 0]) * (py - a[:, 1]) / (b[:, 1] - a[:, 1]) + a[:, 0]
            if on_boundary or (np.sum(y_cond & (px < x_int)) % 2 == 1):
                out[i] = j
                break
    return out


def op_interpolate(kind, coords, conn, u, points):
    """operator.py:399-463: containing element, then ONE Newton step from the first quadrature point
    (xi = xi0 - (dx/dxi)^-1 (x(xi0) - p); exact for affine elements), then N(xi) . u_e.  NaN rows where no element
    contains the point (the reference raises RuntimeError in that case when not traced)."""
    points = np.asarray(points, dtype=np.float64)
    idx = find_containing_polygons(points, coords[conn])
    xi0 = quad_rule(kind)[0][0]
    u = np.asarray(u, dtype=np.float64)
    out = np.full((points.shape[0],) + u.shape[1:], np.nan)
    for i, e in enumerate(idx):
        if e < 0:
            continue
        X_e = coords[conn[e]]
        x0 = shape_function(kind, xi0) @ X_e
        lhs = (shape_function_derivative(kind, xi0) @ X_e).T  # d x_i / d xi_j
        xi = xi0 + np.linalg.solve(lhs, -(x0 - points[i]))
        out[i] = np.einsum("n,n...->...", shape_function(kind, xi), u[conn[e]])
    return out, idx


# ----------------------------------------------------------------------------------------
# Constitutive laws used by the configs (user code in the reference, pinned by its tests)
# ----------------------------------------------------------------------------------------


def lame_from_youngs_poisson_2d(E, nu, plane_stress=False):
    """tests/test_sparse_benchmark.py:30-43."""
    mu = E / 2 / (1 + nu)
    lmbda = 2 * nu * mu / (1 - nu) if plane_stress else E * nu / (1 - 2 * nu) / (1 + nu)
    return mu, lmbda


class LinearElastic:
    """psi = 1/2 sigma:eps, sigma = 2 mu eps + lambda tr(eps) I   (tests/test_sparse.py:20-38)."""

    name = "linear_elastic"

    def __init__(self, mu, lmbda):
        self.mu, self.lmbda = float(mu), float(lmbda)

    def psi(self, G):
        eps = 0.5 * (G + np.swapaxes(G, -1, -2))
        tr = np.trace(eps, axis1=-2, axis2=-1)
        return self.mu * np.einsum("...ij,...ij->...", eps, eps) + 0.5 * self.lmbda * tr * tr

    def P(self, G):
        d = G.shape[-1]
        eps = 0.5 * (G + np.swapaxes(G, -1, -2))
        tr = np.trace(eps, axis1=-2, axis2=-1)
        return 2 * self.mu * eps + self.lmbda * tr[..., None, None] * np.eye(d)

    def dP(self, G, dG):
        return self.P(dG)


class NeoHookean:
    """psi = mu/2 (I1 - 3 - 2 ln J) + lambda/2 (ln J)^2, F = I + grad_u
    (tests/test_sparse_tracer.py:103-115; mu=500, lambda=1000 at :126)."""

    name = "neo_hookean"

    def __init__(self, mu, lmbda):
        self.mu, self.lmbda = float(mu), float(lmbda)

    def _F(self, G):
        return np.eye(G.shape[-1]) + G

    def psi(self, G):
        F = self._F(G)
        J = np.linalg.det(F)
        I1 = np.einsum("...ij,...ij->...", F, F)
        lnJ = np.log(J)
        return 0.5 * self.mu * (I1 - 3 - 2 * lnJ) + 0.5 * self.lmbda * lnJ * lnJ

    def P(self, G):
        F = self._F(G)
        FinvT = np.swapaxes(np.linalg.inv(F), -1, -2)
        lnJ = np.log(np.linalg.det(F))[..., None, None]
        return self.mu * (F - FinvT) + self.lmbda * lnJ * FinvT

    def dP(self, G, dG):
        F = self._F(G)
        Finv = np.linalg.inv(F)
        FinvT = np.swapaxes(Finv, -1, -2)
        lnJ = np.log(np.linalg.det(F))[..., None, None]
        FiG = Finv @ dG  # F^-1 dG
        tr = np.trace(FiG, axis1=-2, axis2=-1)[..., None, None]
        # F^-T dG^T F^-T = (F^-1 dG F^-1)^T
        return self.mu * dG + (self.mu - self.lmbda * lnJ) * np.swapaxes(FiG @ Finv, -1, -2) + self.lmbda * tr * FinvT


class NeoHookeanPhaseField:
    """Config 5 two-field density (builder-defined AT2; no counterpart in the reference):
        psi(grad_u, phi, grad_phi) = ((1-phi)^2 + k) psi_NH(grad_u) + Gc (phi^2/(2 l) + l/2 |grad_phi|^2)
    Nodal state is the Compound-stacked interleaving [ux, uy, uz, phi] per node
    (compound/__init__.py:334-389, pinned by tests/test_compound.py:134-147)."""

    name = "neo_hookean_phase_field"

    def __init__(self, mu, lmbda, Gc, ell, k):
        self.nh = NeoHookean(mu, lmbda)
        self.mu, self.lmbda = float(mu), float(lmbda)
        self.Gc, self.ell, self.k = float(Gc), float(ell), float(k)

    def psi(self, G, phi, gphi):
        g = (1 - phi) ** 2 + self.k
        return g * self.nh.psi(G) + self.Gc * (phi * phi / (2 * self.ell) + 0.5 * self.ell * np.einsum("...j,...j->...", gphi, gphi))

    def first(self, G, phi, gphi):
        """(dpsi/dG, dpsi/dphi, dpsi/dgphi)."""
        g = (1 - phi) ** 2 + self.k
        dg = -2 * (1 - phi)
        return (
            g[..., None, None] * self.nh.P(G),
            dg * self.nh.psi(G) + self.Gc * phi / self.ell,
            self.Gc * self.ell * gphi,
        )

    def second(self, G, phi, gphi, dG, dphi, dgphi):
        """Directional derivative of `first` along (dG, dphi, dgphi)."""
        g = (1 - phi) ** 2 + self.k
        dg = -2 * (1 - phi)
        P = self.nh.P(G)
        PdG = np.einsum("...ij,...ij->...", P, dG)
        return (
            g[..., None, None] * self.nh.dP(G, dG) + (dg * dphi)[..., None, None] * P,
            dg * PdG + 2 * dphi * self.nh.psi(G) + self.Gc * dphi / self.ell,
            self.Gc * self.ell * dgphi,
        )


# ----------------------------------------------------------------------------------------
# Energy, residual, HVP  (README.md:93; call sites sparse/base.py:264, :213)
# ----------------------------------------------------------------------------------------


def energy(kind, mat, coords, conn, u):
    """E(u) = op.integrate(psi(op.grad(u)))   (tests/test_sparse.py:50-55)."""
    psi = mat.psi(op_grad(kind, coords, conn, u))
    return op_integrate(kind, coords, conn, psi)


def _scatter_nodes(conn, contrib, n_nodes):
    """Transpose of the gather at operator.py:221: add (E,n,...) into (N,...)."""
    out = np.zeros((n_nodes,) + contrib.shape[2:], dtype=contrib.dtype)
    np.add.at(out, conn, contrib)
    return out


def residual(kind, mat, coords, conn, u):
    """r[n,i] = sum_e sum_q W[e,q] P_ij(grad u) dNdX[j,n]  = jax.grad(E)(u)."""
    dNdX, detJ = geometry(kind, coords, conn)
    _, w = quad_rule(kind)
    G = np.einsum("eqdn,eni->eqid", dNdX, u[conn])
    WP = mat.P(G) * (detJ * w)[..., None, None]
    return _scatter_nodes(conn, np.einsum("eqij,eqjn->eni", WP, dNdX), coords.shape[0])


def hvp(kind, mat, coords, conn, u, v):
    """Hv = d/d eps r(u + eps v) = jax.jvp(jax.grad(E), (u,), (v,))[1]."""
    dNdX, detJ = geometry(kind, coords, conn)
    _, w = quad_rule(kind)
    G = np.einsum("eqdn,eni->eqid", dNdX, u[conn])
    dG = np.einsum("eqdn,eni->eqid", dNdX, v[conn])
    WdP = mat.dP(G, dG) * (detJ * w)[..., None, None]
    return _scatter_nodes(conn, np.einsum("eqij,eqjn->eni", WdP, dNdX), coords.shape[0])


# Compound (u, phi) variants: nodal state s (N,4) = [ux,uy,uz,phi]


def energy_pf(kind, mat, coords, conn, s):
    G = op_grad(kind, coords, conn, s[:, :3])
    gphi = op_grad(kind, coords, conn, s[:, 3])
    phi = op_eval(kind, coords, conn, s[:, 3])
    return op_integrate(kind, coords, conn, mat.psi(G, phi, gphi))


def _pf_fields(kind, coords, conn, s):
    dNdX, detJ = geometry(kind, coords, conn)
    N = _N_all(kind)
    se = s[conn]
    G = np.einsum("eqdn,eni->eqid", dNdX, se[..., :3])
    gphi = np.einsum("eqdn,en->eqd", dNdX, se[..., 3])
    phi = np.einsum("qn,en->eq", N, se[..., 3])
    return dNdX, detJ, N, G, phi, gphi


def _pf_scatter(kind, coords, conn, dNdX, detJ, N, A, b, c):
    _, w = quad_rule(kind)
    W = detJ * w
    ru = np.einsum("eq,eqij,eqjn->eni", W, A, dNdX)
    rphi = np.einsum("eq,eq,qn->en", W, b, N) + np.einsum("eq,eqj,eqjn->en", W, c, dNdX)
    return _scatter_nodes(conn, np.concatenate([ru, rphi[..., None]], axis=-1), coords.shape[0])


def residual_pf(kind, mat, coords, conn, s):
    dNdX, detJ, N, G, phi, gphi = _pf_fields(kind, coords, conn, s)
    A, b, c = mat.first(G, phi, gphi)
    return _pf_scatter(kind, coords, conn, dNdX, detJ, N, A, b, c)


def hvp_pf(kind, mat, coords, conn, s, t):
    dNdX, detJ, N, G, phi, gphi = _pf_fields(kind, coords, conn, s)
    _, _, _, dG, dphi, dgphi = _pf_fields(kind, coords, conn, t)
    A, b, c = mat.second(G, phi, gphi, dG, dphi, dgphi)
    return _pf_scatter(kind, coords, conn, dNdX, detJ, N, A, b, c)


# ----------------------------------------------------------------------------------------
# Sparsity pattern, colouring, coloured Jacobian (tatva/sparse)
# ----------------------------------------------------------------------------------------


def pattern_from_mesh(conn, n_nodes, n_dofs_per_node):
    """sparse/_extraction.py:37-102: CSR (indptr, indices) int32 with sorted columns."""
    el = np.asarray(conn, dtype=np.int64)
    E, npe = el.shape
    nde = npe * n_dofs_per_node
    dofs = (el[..., None] * n_dofs_per_node + np.arange(n_dofs_per_node, dtype=np.int64)).reshape(E, -1)  # :57-60
    n = n_nodes * n_dofs_per_node
    rows = np.repeat(dofs, nde, axis=1).ravel()  # :62
    cols = np.tile(dofs, (1, nde)).ravel()  # :63
    lin = np.unique(rows * n + cols)  # :74-79
    r, c = lin // n, lin % n
    m = sps.csr_matrix((np.ones(r.shape[0], dtype=np.int8), (r, c)), shape=(n, n))  # :85-88
    return m.indptr.astype(np.int32), m.indices.astype(np.int32)


def pattern_from_compound_nodal(conn, n_nodes, n_dofs_per_node):
    """sparse/_extraction.py:118-245 for the all-full-Nodal stacked case (config 5): the stacked
    block makes flat DOF id = node*K + comp (compound/__init__.py:334-389), and every field's
    element DOFs are concatenated -> the same set of pairs as pattern_from_mesh with K DOFs."""
    return pattern_from_mesh(conn, n_nodes, n_dofs_per_node)


def distance2_colors(indptr, indices, n_dofs):
    """sparse/_coloring.py:270-283 -> get_distance2_adjacency (:27-48) -> greedy_coloring (:136-153):
    first-fit in natural order on the pattern of A@A (self included; self is uncoloured when visited)."""
    A = sps.csr_matrix((np.ones(len(indices), dtype=bool), indices, indptr), shape=(n_dofs, n_dofs))
    A2 = (A @ A).tocsr()
    A2.sort_indices()
    ip, ix = A2.indptr, A2.indices
    colors = -np.ones(n_dofs, dtype=np.int32)
    for i in range(n_dofs):
        used = set(colors[ix[ip[i] : ip[i + 1]]].tolist())
        c = 0
        while c in used:
            c += 1
        colors[i] = c
    return colors


def compute_rows_cols(indptr, indices, colors):
    """sparse/base.py:108-136."""
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    return rows, colors[indices]


def colored_jacobian_data(residual_jvp, n, indptr, indices, colors):
    """sparse/base.py:139-176 + :230-270: one JVP per colour with a 0/1 seed -> dense (N, n_colors)
    -> data[k] = J_c[row(k), colors[indices[k]]].  `residual_jvp(seed)` returns Hv flat."""
    n_colors = int(colors.max()) + 1
    Jc = np.empty((n, n_colors))
    for c in range(n_colors):
        seed = np.where(colors == c, 1.0, 0.0)
        Jc[:, c] = residual_jvp(seed)
    rows, cc = compute_rows_cols(indptr, indices, colors)
    return Jc[rows, cc]


def assemble_csr_data(kind, mat, coords, conn, u, indptr, indices, dofs_per_node=None, hvp_fn=None):
    """Direct assembly of the same matrix (what the CUDA kernel does): element stiffness
    K_e[(a,i),(b,k)] = sum_q W dNdX[j,a] dP_ij[e_k (x) dNdX[:,b]] added into the fixed CSR pattern."""
    dim = coords.shape[1]
    dNdX, detJ = geometry(kind, coords, conn)
    _, w = quad_rule(kind)
    W = detJ * w
    E, npe = conn.shape
    G = np.einsum("eqdn,eni->eqid", dNdX, u[conn])
    Ke = np.zeros((E, npe, dim, npe, dim))
    for b in range(npe):
        for k in range(dim):
            dG = np.zeros(G.shape)
            dG[..., k, :] = dNdX[..., :, b]
            dP = mat.dP(G, dG)
            Ke[:, :, :, b, k] = np.einsum("eq,eqij,eqja->eai", W, dP, dNdX)
    n = coords.shape[0] * dim
    dofs = (conn[..., None] * dim + np.arange(dim)).reshape(E, -1)
    rows = np.repeat(dofs, npe * dim, axis=1).ravel()
    cols = np.tile(dofs, (1, npe * dim)).ravel()
    K = sps.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    # project onto the fixed pattern
    P = sps.csr_matrix((np.arange(1, len(indices) + 1, dtype=np.float64), indices, indptr), shape=(n, n))
    K = K.tocoo()
    pos = np.asarray(P[K.row, K.col]).ravel().astype(np.int64) - 1
    data = np.zeros(len(indices))
    data[pos] = K.data
    return data


# ----------------------------------------------------------------------------------------
# Partitioning and exchange plans (tatva/mesh.py:234-291, tatva/mpi.py)
# ----------------------------------------------------------------------------------------


def extract_local_mesh(coords, conn, element_partition, part):
    """mesh.py:234-291 -> (coords_local, conn_local, nodes_local_to_global, n_owned_nodes)."""
    local_el = conn[element_partition == part]
    all_local = np.unique(local_el.ravel())  # :258
    owner = np.full(len(coords), element_partition.max() + 1, dtype=np.int32)
    for col in range(conn.shape[1]):
        np.minimum.at(owner, conn[:, col], element_partition)  # :263-265
    is_owned = owner[all_local] == part
    l2g = np.concatenate([all_local[is_owned], all_local[~is_owned]])  # :267-273
    g2l = np.full(len(coords), -1, dtype=np.int32)
    g2l[l2g] = np.arange(len(l2g), dtype=np.int32)
    return coords[l2g], g2l[local_el], l2g, int(is_owned.sum())


def dof_map_from_node_map(node_map, dpn):
    """compound/mpi.py:280-285."""
    return (np.asarray(node_map)[:, None] * dpn + np.arange(dpn)).ravel().astype(np.int32)


def create_dof_layouts(natural_maps, owned_masks, n_natural_global):
    """mpi.py:83-130 for all ranks at once (the MPI collectives become plain loops).
    Returns per rank dict(local_to_global, offset, n_owned, n_total, n_global)."""
    size = len(natural_maps)
    n_owned = [int(np.sum(m)) for m in owned_masks]
    n_global = sum(n_owned)
    offsets = np.concatenate([[0], np.cumsum(n_owned)])[:size]  # :94-99
    directory = np.full(n_natural_global, -1, dtype=np.int32)
    l2gs = []
    for r in range(size):
        l2g = np.full(natural_maps[r].size, -1, dtype=np.int32)
        oi = np.where(owned_masks[r])[0]
        l2g[oi] = offsets[r] + np.arange(n_owned[r], dtype=np.int32)  # :102-104
        directory[natural_maps[r][oi]] = np.maximum(directory[natural_maps[r][oi]], l2g[oi])  # :107-111 MAX
        l2gs.append(l2g)
    out = []
    for r in range(size):
        gi = np.where(~owned_masks[r])[0]
        l2gs[r][gi] = directory[natural_maps[r][gi]]  # :113-115
        out.append(
            dict(
                local_to_global=l2gs[r],
                offset=int(offsets[r]),
                n_owned=n_owned[r],
                n_total=int(natural_maps[r].size),
                n_global=n_global,
                owned_mask=owned_masks[r],
            )
        )
    return out


def exchange_routing(layouts):
    """mpi.py:168-234 for all ranks at once.  Returns per rank:
    dict(self_send, self_recv, neighbors=[dict(rank, local_send_idx, recv_local_idx)])."""
    size = len(layouts)
    ranges = [(L["offset"], L["offset"] + L["n_owned"]) for L in layouts]
    to_send = [[None] * size for _ in range(size)]
    for r in range(size):
        l2g = layouts[r]["local_to_global"]
        for nbr in range(size):
            if nbr == r:
                to_send[r][nbr] = np.array([], dtype=np.int32)
            else:
                rs, re = ranges[nbr]
                to_send[r][nbr] = np.where((l2g >= rs) & (l2g < re))[0].astype(np.int32)  # :196-198
    plans = []
    for r in range(size):
        l2g = layouts[r]["local_to_global"]
        nbrs = []
        for nbr in range(size):
            if nbr == r:
                continue
            s, rc = len(to_send[r][nbr]), len(to_send[nbr][r])
            if s == 0 and rc == 0:
                continue  # :204-208
            recv_global = layouts[nbr]["local_to_global"][to_send[nbr][r]]  # what nbr sends us (:211-216)
            nbrs.append(
                dict(
                    rank=nbr,
                    local_send_idx=to_send[r][nbr],
                    recv_local_idx=(recv_global - ranges[r][0]).astype(np.int32),  # :222
                )
            )
        ss = np.where(layouts[r]["owned_mask"])[0].astype(np.int32)  # :229-231
        plans.append(dict(self_send=ss, self_recv=(l2g[ss] - ranges[r][0]).astype(np.int32), neighbors=nbrs))
    return plans


def scatter_fwd_set(plans, layouts, x_owned):
    """mpi.py:372-409 for all ranks at once: owned -> local (ghost fill)."""
    out = []
    for r, (p, L) in enumerate(zip(plans, layouts)):
        u = np.zeros(L["n_total"])
        u[p["self_send"]] = x_owned[r][p["self_recv"]]
        for nb in p["neighbors"]:
            # neighbour sends x_owned[nbr][its recv_local_idx for us]; we store at our local_send_idx
            nbp = next(q for q in plans[nb["rank"]]["neighbors"] if q["rank"] == r)
            u[nb["local_send_idx"]] = x_owned[nb["rank"]][nbp["recv_local_idx"]]
        out.append(u)
    return out


def scatter_rev_add(plans, layouts, data_local):
    """mpi.py:479-516 for all ranks at once: local -> owned (ghost contributions added)."""
    out = []
    for r, (p, L) in enumerate(zip(plans, layouts)):
        owned = np.zeros(L["n_owned"])
        np.add.at(owned, p["self_recv"], data_local[r][p["self_send"]])
        for nb in p["neighbors"]:
            nbp = next(q for q in plans[nb["rank"]]["neighbors"] if q["rank"] == r)
            np.add.at(owned, nb["recv_local_idx"], data_local[nb["rank"]][nbp["local_send_idx"]])
        out.append(owned)
    return out


def dof_range(n, size, rank):
    """mpi.py:714-726."""
    base, rem = divmod(n, size)
    if rank < rem:
        start = rank * (base + 1)
        return start, start + base + 1
    start = rank * base + rem
    return start, start + base


# ----------------------------------------------------------------------------------------
# Synthetic inputs; meshes follow the reference generators
# ----------------------------------------------------------------------------------------


def mesh_unit_square_tri(nx, ny):
    """mesh.py:181-205 (Mesh.unit_square -> _rectangle_triangular)."""
    xv, yv = np.meshgrid(np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1), indexing="ij")
    coords = np.stack([xv.ravel(), yv.ravel()], axis=-1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    n0 = (i * (ny + 1) + j).ravel()
    n1, n2, n3 = n0 + (ny + 1), n0 + 1, n0 + (ny + 1) + 1
    el = np.stack([np.stack([n0, n1, n3], -1), np.stack([n0, n3, n2], -1)], axis=1).reshape(-1, 3)
    return coords, el.astype(np.int32)


def mesh_box_tet(lengths, nb):
    """tests/test_sparse_tracer.py:29-70 (6 tets per cell)."""
    (lx, ly, lz), (nx, ny, nz) = lengths, nb
    xr, yr, zr = np.linspace(-lx / 2, lx / 2, nx + 1), np.linspace(-ly / 2, ly / 2, ny + 1), np.linspace(0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zr, yr, xr, indexing="ij")
    nodes = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (i * sx + j * sy + k * sz).ravel()
    n1, n2 = n0 + sx, n0 + sy
    n3, n4 = n2 + sx, n0 + sz
    n5, n6 = n4 + sx, n4 + sy
    n7 = n6 + sx
    tets = np.stack(
        [
            np.stack([n0, n1, n3, n7], -1),
            np.stack([n0, n1, n7, n5], -1),
            np.stack([n0, n5, n7, n4], -1),
            np.stack([n0, n3, n2, n7], -1),
            np.stack([n0, n2, n6, n7], -1),
            np.stack([n0, n6, n4, n7], -1),
        ],
        axis=1,
    ).reshape(-1, 4)
    return nodes, tets.astype(np.int32)


def mesh_box_hex(n, length=1.0):
    """Unit-cube Hex8 box: node id i + j(n+1) + k(n+1)^2, element node order element/base.py:478-491."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    xr, yr, zr = (np.linspace(0, length * m / max(nx, ny, nz), m + 1) for m in (nx, ny, nz))
    Z, Y, X = np.meshgrid(zr, yr, xr, indexing="ij")
    nodes = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (i * sx + j * sy + k * sz).ravel()
    el = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + sz, n0 + sz + sx, n0 + sz + sx + sy, n0 + sz + sy], -1)
    return nodes, el.astype(np.int32)


def mesh_unit_square_quad(nx, ny):
    """mesh.py:207-231 (Mesh.unit_square(type="quad") -> _rectangle_quadrilateral)."""
    xv, yv = np.meshgrid(np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1), indexing="ij")
    coords = np.stack([xv.ravel(), yv.ravel()], axis=-1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    n0 = (i * (ny + 1) + j).ravel()
    return coords, np.stack([n0, n0 + (ny + 1), n0 + (ny + 1) + 1, n0 + 1], -1).astype(np.int32)


def mesh_second_order(kind, nx, ny):
    """Tri6 / Quad8 meshes (no generator in the reference): mid-edge nodes added to the linear mesh, node order
    of element/base.py:272-276 / :369-382."""
    base_c, base_el = mesh_unit_square_tri(nx, ny) if kind == "tri6" else mesh_unit_square_quad(nx, ny)
    edges = [(0, 1), (1, 2), (2, 0)] if kind == "tri6" else [(0, 1), (1, 2), (2, 3), (3, 0)]
    mid, coords, el = {}, [tuple(p) for p in base_c], []
    for e in base_el:
        extra = []
        for a, b in edges:
            key = (min(e[a], e[b]), max(e[a], e[b]))
            if key not in mid:
                mid[key] = len(coords)
                coords.append(tuple(0.5 * (base_c[e[a]] + base_c[e[b]])))
            extra.append(mid[key])
        el.append(list(e) + extra)
    return np.array(coords), np.array(el, dtype=np.int32)
