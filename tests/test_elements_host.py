"""Host element classes (tatva_b200.element) against the reference's outputs and its exactness tests
(reference tests/test_element.py:45-149)."""
import numpy as np
import pytest

from tatva_b200 import element

CLASSES = {"tri3": element.Tri3, "tet4": element.Tetrahedron4, "hex8": element.Hexahedron8, "quad4": element.Quad4, "tri6": element.Tri6, "quad8": element.Quad8}


@pytest.mark.parametrize("kind", list(CLASSES))
def test_element_classes_match_reference_outputs(golden, kind):
    el = CLASSES[kind]()
    g = lambda k: golden[f"el_{kind}_{k}"]  # noqa: E731
    np.testing.assert_allclose(el.quad_points, g("qp"), atol=1e-15)
    np.testing.assert_allclose(el.quad_weights, g("qw"), atol=1e-15)
    X, uv, us = g("X"), g("uv"), g("us")
    for q, xi in enumerate(el.quad_points):
        np.testing.assert_allclose(el.shape_function(xi), g("N")[q], atol=1e-15)
        np.testing.assert_allclose(el.shape_function_derivative(xi), g("dNdr")[q], atol=1e-15)
        J, detJ = el.get_jacobian(xi, X)
        np.testing.assert_allclose(J, g("J")[q], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(detJ, g("detJ")[q], rtol=1e-13)
        np.testing.assert_allclose(el.gradient(xi, uv, X), g("grad_v")[q], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(el.gradient(xi, us, X), g("grad_s")[q], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(el.interpolate(xi, uv, X), g("interp_v")[q], rtol=1e-14, atol=1e-15)
        val, grad, dj = el.get_local_values(xi, uv, X)
        np.testing.assert_allclose(grad, g("grad_v")[q], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("cls", [element.Tri3, element.Tri6, element.Quad4, element.Quad8, element.Tetrahedron4, element.Hexahedron8])
def test_linear_fields_have_exact_gradients(cls):
    """reference tests/test_element.py:85-149."""
    el = cls()
    X = el._reference_nodes()
    dim = X.shape[1]
    rng = np.random.default_rng(0)
    A, B = rng.normal(size=(dim, dim)), rng.normal(size=(2, 2, dim))
    for xi in el.quad_points:
        np.testing.assert_allclose(el.gradient(xi, np.einsum("ij,kj->ki", A, X), X), A, atol=1e-12)
        np.testing.assert_allclose(el.gradient(xi, np.einsum("ijk,nk->nij", B, X), X), B, atol=1e-12)


@pytest.mark.parametrize("cls,coords", [(element.Line2, [[0.0, 0.0], [1.0, 0.0]]), (element.Line3, [[0.0, 0.0], [1.0, 0.0], [0.5, 0.0]])])
def test_line_elements_arc_length_derivative(cls, coords):
    """reference tests/test_element.py:45-82: u = 3 x along an x-aligned line -> derivative 3."""
    el = cls()
    X = np.array(coords)
    for xi in el.quad_points:
        np.testing.assert_allclose(el.gradient(xi, 3.0 * X[:, 0], X), 3.0, atol=1e-12)
    assert abs(sum(el.get_jacobian(xi, X)[1] * w for xi, w in zip(el.quad_points, el.quad_weights)) - 1.0) < 1e-14


def test_elements_are_value_hashable():
    assert element.Tri3() == element.Tri3() and hash(element.Quad8()) == hash(element.Quad8())
    assert element.Tri3() != element.Tri6()
