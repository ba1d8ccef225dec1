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


__all__ = ["Element", "Tri3", "Tetrahedron4", "Hexahedron8"]
