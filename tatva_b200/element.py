"""Element library — host-side mirror of tatva.element (tatva/element/base.py).

The classes keep the reference's names, attributes (`quad_points`, `quad_weights`) and per-point
methods (`shape_function`, `shape_function_derivative`, `get_jacobian`, `interpolate`, `gradient`,
`get_local_values`).  Here they are small NumPy helpers used at set-up time and by tests; the
quadrature loop itself runs in the CUDA kernels selected by `Element.kind`
(tatva_b200/csrc/common.cuh holds the same tables as device code).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from . import _lib


class Element(ABC):
    """Base class; value-based equality/hash like the reference (element/base.py:53-68)."""

    kind: int | None = None  # kernel id (TATVA_TRI3 / TATVA_TET4 / TATVA_HEX8) or None

    def __init__(self, quad_points=None, quad_weights=None):
        if quad_points is not None and quad_weights is not None:
            self.quad_points = np.asarray(quad_points, dtype=np.float64)
            self.quad_weights = np.asarray(quad_weights, dtype=np.float64)
            self._default_rule = False
        else:
            self.quad_points, self.quad_weights = self._default_quadrature()
            self._default_rule = True

    def __eq__(self, other):
        if type(self) is not type(other):
            return False
        return np.array_equal(self.quad_points, other.quad_points) and np.array_equal(
            self.quad_weights, other.quad_weights
        )

    def __hash__(self):
        return hash((type(self), self.quad_points.tobytes(), self.quad_weights.tobytes()))

    @abstractmethod
    def _reference_nodes(self) -> np.ndarray: ...

    @abstractmethod
    def _default_quadrature(self) -> tuple[np.ndarray, np.ndarray]: ...

    @abstractmethod
    def shape_function(self, xi) -> np.ndarray: ...

    @abstractmethod
    def shape_function_derivative(self, xi) -> np.ndarray: ...

    # element/base.py:90-93
    def get_jacobian(self, xi, nodal_coords):
        J = self.shape_function_derivative(xi) @ np.asarray(nodal_coords)
        return J, np.linalg.det(J)

    # element/base.py:95-97
    def interpolate(self, xi, nodal_values, nodal_coords=None):
        return np.einsum("n,n...->...", self.shape_function(xi), np.asarray(nodal_values))

    # element/base.py:99-115 — value dims first, spatial dim last
    def gradient(self, xi, nodal_values, nodal_coords):
        dNdr = self.shape_function_derivative(xi)
        J = dNdr @ np.asarray(nodal_coords)
        dNdX = np.linalg.solve(J, dNdr)
        return np.einsum("dn,n...->...d", dNdX, np.asarray(nodal_values))

    # element/base.py:117-141
    def get_local_values(self, xi, nodal_values, nodal_coords):
        J, detJ = self.get_jacobian(xi, nodal_coords)
        return (
            self.interpolate(xi, nodal_values, nodal_coords),
            self.gradient(xi, nodal_values, nodal_coords),
            detJ,
        )


class Tri3(Element):
    """3-node linear triangle (element/base.py:245-265)."""

    kind = _lib.TRI3

    def _reference_nodes(self):
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])

    def _default_quadrature(self):
        return np.array([[1.0 / 3, 1.0 / 3]]), np.array([1.0 / 2])

    def shape_function(self, xi):
        return np.array([1.0 - xi[0] - xi[1], xi[0], xi[1]])

    def shape_function_derivative(self, *_a, **_k):
        return np.array([[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]])


class Tetrahedron4(Element):
    """4-node linear tetrahedron (element/base.py:448-472)."""

    kind = _lib.TET4

    def _reference_nodes(self):
        return np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])

    def _default_quadrature(self):
        return np.array([[0.25, 0.25, 0.25]]), np.array([1.0 / 6])

    def shape_function(self, xi):
        return np.array([1.0 - xi[0] - xi[1] - xi[2], xi[0], xi[1], xi[2]])

    def shape_function_derivative(self, *_a, **_k):
        return np.array([[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]])


_HEX = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]],
    dtype=np.float64,
)


class Hexahedron8(Element):
    """8-node trilinear hexahedron, 2x2x2 Gauss rule (element/base.py:475-568)."""

    kind = _lib.HEX8

    def _reference_nodes(self):
        return _HEX.copy()

    def _default_quadrature(self):
        return _HEX / np.sqrt(3.0), np.ones(8)

    def shape_function(self, xi):
        return 0.125 * (1 + _HEX[:, 0] * xi[0]) * (1 + _HEX[:, 1] * xi[1]) * (1 + _HEX[:, 2] * xi[2])

    def shape_function_derivative(self, xi):
        fx, fy, fz = 1 + _HEX[:, 0] * xi[0], 1 + _HEX[:, 1] * xi[1], 1 + _HEX[:, 2] * xi[2]
        return 0.125 * np.stack([_HEX[:, 0] * fy * fz, _HEX[:, 1] * fx * fz, _HEX[:, 2] * fx * fy])


_Q4 = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64)


class Quad4(Element):
    """4-node bilinear quadrilateral, 2x2 Gauss rule with x fastest (element/base.py:331-366)."""

    kind = _lib.QUAD4

    def _reference_nodes(self):
        return _Q4.copy()

    def _default_quadrature(self):
        a = 1.0 / np.sqrt(3.0)
        x = np.array([-a, a])
        return np.array([[x[j], x[i]] for i in range(2) for j in range(2)]), np.ones(4)

    def shape_function(self, xi):
        return 0.25 * (1 + _Q4[:, 0] * xi[0]) * (1 + _Q4[:, 1] * xi[1])

    def shape_function_derivative(self, xi):
        return 0.25 * np.stack([_Q4[:, 0] * (1 + _Q4[:, 1] * xi[1]), _Q4[:, 1] * (1 + _Q4[:, 0] * xi[0])])


class Tri6(Element):
    """6-node quadratic triangle: 3 vertices then mid-edges (0-1, 1-2, 2-0); 3-point rule (element/base.py:266-328)."""

    kind = _lib.TRI6

    def _reference_nodes(self):
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.0], [0.5, 0.5], [0.0, 0.5]])

    def _default_quadrature(self):
        return np.array([[1.0 / 6, 1.0 / 6], [2.0 / 3, 1.0 / 6], [1.0 / 6, 2.0 / 3]]), np.full(3, 1.0 / 6)

    def shape_function(self, xi):
        r, s = xi[0], xi[1]
        t = 1.0 - r - s
        return np.array([t * (2 * t - 1), r * (2 * r - 1), s * (2 * s - 1), 4 * r * t, 4 * r * s, 4 * s * t])

    def shape_function_derivative(self, xi):
        r, s = xi[0], xi[1]
        t = 1.0 - r - s
        return np.array([[-(4 * t - 1), 4 * r - 1, 0.0, 4 * (t - r), 4 * s, -4 * s], [-(4 * t - 1), 0.0, 4 * s - 1, -4 * r, 4 * r, 4 * (t - s)]])


class Quad8(Element):
    """8-node serendipity quadrilateral: corners then mid-edges; 3x3 Gauss rule, x fastest (element/base.py:366-445)."""

    kind = _lib.QUAD8

    def _reference_nodes(self):
        return np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0], [0.0, -1.0], [1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])

    def _default_quadrature(self):
        b = np.sqrt(3.0 / 5.0)
        x, w = np.array([-b, 0.0, b]), np.array([5.0 / 9, 8.0 / 9, 5.0 / 9])
        return np.array([[x[j], x[i]] for i in range(3) for j in range(3)]), np.kron(w, w)

    def shape_function(self, xi):
        r, s = xi[0], xi[1]
        corner = [0.25 * (1 + a * r) * (1 + b * s) * (a * r + b * s - 1) for a, b in _Q4]
        return np.array(corner + [0.5 * (1 - r * r) * (1 - s), 0.5 * (1 + r) * (1 - s * s), 0.5 * (1 - r * r) * (1 + s), 0.5 * (1 - r) * (1 - s * s)])

    def shape_function_derivative(self, xi):
        r, s = xi[0], xi[1]
        dr = [0.25 * a * (1 + b * s) * (2 * a * r + b * s) for a, b in _Q4] + [-r * (1 - s), 0.5 * (1 - s * s), -r * (1 + s), -0.5 * (1 - s * s)]
        ds = [0.25 * b * (1 + a * r) * (a * r + 2 * b * s) for a, b in _Q4] + [-0.5 * (1 - r * r), -s * (1 + r), 0.5 * (1 - r * r), -s * (1 - r)]
        return np.array([dr, ds])


class _LineElement(Element):
    """1-D elements embedded in 2-D: the Jacobian is the arc-length derivative |dX/dxi| and the gradient is the
    derivative along the line (element/base.py:164-188, :216-243).  `Operator.grad` returns (E, Q, *v): no trailing
    spatial axis."""

    kind = None
    gradient_components = 1

    def get_jacobian(self, xi, nodal_coords):
        Jvec = self.shape_function_derivative(xi) @ np.asarray(nodal_coords)
        J = float(np.dot(Jvec, Jvec / np.linalg.norm(Jvec)))
        return J, J

    def gradient(self, xi, nodal_values, nodal_coords):
        J, _ = self.get_jacobian(xi, nodal_coords)
        return np.einsum("n,n...->...", self.shape_function_derivative(xi) / J, np.asarray(nodal_values))

    def get_local_values(self, xi, nodal_values, nodal_coords):
        J, detJ = self.get_jacobian(xi, nodal_coords)
        return self.interpolate(xi, nodal_values), self.gradient(xi, nodal_values, nodal_coords), detJ


class Line2(_LineElement):
    """2-node linear interval (element/base.py:144-188)."""

    kind = _lib.LINE2

    def _reference_nodes(self):
        return np.array([[-1.0], [1.0]])

    def _default_quadrature(self):
        return np.array([[0.0]]), np.array([2.0])

    def shape_function(self, xi):
        return np.array([0.5 * (1.0 - xi[0]), 0.5 * (1.0 + xi[0])])

    def shape_function_derivative(self, xi):
        return np.array([-0.5, 0.5])


class Line3(_LineElement):
    """3-node quadratic interval: end nodes then the midpoint (element/base.py:191-243)."""

    kind = _lib.LINE3

    def _reference_nodes(self):
        return np.array([[-1.0], [1.0], [0.0]])

    def _default_quadrature(self):
        b = np.sqrt(3.0 / 5.0)
        return np.array([[-b], [0.0], [b]]), np.array([5.0 / 9, 8.0 / 9, 5.0 / 9])

    def shape_function(self, xi):
        r = xi[0]
        return np.array([0.5 * r * (r - 1.0), 0.5 * r * (r + 1.0), 1.0 - r * r])

    def shape_function_derivative(self, xi):
        r = xi[0]
        return np.array([r - 0.5, r + 0.5, -2.0 * r])


__all__ = ["Element", "Line2", "Line3", "Tri3", "Tri6", "Quad4", "Quad8", "Tetrahedron4", "Hexahedron8"]
