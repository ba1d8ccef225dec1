"""tatva_b200.utils on CPU tensors (the functions are device-agnostic torch code)."""
import numpy as np
import pytest
import torch

from tatva_b200 import utils


def test_virtual_work_to_residual_direct_and_decorator():
    """reference tatva/utils.py:39-115: residual = d fn / d test for fn linear in its first argument."""
    A = torch.as_tensor(np.random.default_rng(0).normal(size=(5, 5)))

    def work(test, u, scale=1.0):  # test . (A u) * scale
        return scale * (test * (A @ u)).sum()

    u = torch.as_tensor(np.random.default_rng(1).normal(size=5))
    res = utils.virtual_work_to_residual(work, test_size=5)
    torch.testing.assert_close(res(u), A @ u)
    torch.testing.assert_close(res(u, scale=2.0), 2 * (A @ u))

    @utils.virtual_work_to_residual(test_shape=(5,))
    def decorated(test, u):
        return (test * (A @ u)).sum()

    torch.testing.assert_close(decorated(u), A @ u)
    # a given test array is the evaluation point; the result does not depend on it for a linear functional
    res2 = utils.virtual_work_to_residual(work, test_arr=np.ones(5))
    torch.testing.assert_close(res2(u), A @ u)
    with pytest.raises(ValueError):
        utils.virtual_work_to_residual(work)


def test_residual_of_virtual_work_is_differentiable_again():
    """The tangent of the residual (what sparse.jacfwd differentiates) is available through autograd."""

    def work(test, u):
        return (test * u**3).sum()

    res = utils.virtual_work_to_residual(work, test_size=4)
    u = torch.arange(1.0, 5.0, dtype=torch.float64, requires_grad=True)
    r = res(u)
    torch.testing.assert_close(r.detach(), u.detach() ** 3)
    (hv,) = torch.autograd.grad(r, u, torch.ones(4, dtype=torch.float64))
    torch.testing.assert_close(hv, 3 * u.detach() ** 2)


def test_create_g2l_known_answer():
    """reference tatva/utils.py:265-280."""
    g2l = utils.create_g2l(np.array([10, 3, 7]))
    np.testing.assert_array_equal(g2l(np.array([3, 10, 5, 7])), [1, 0, -1, 2])


def test_make_project_function_validates_like_the_reference():
    with pytest.raises(ValueError):
        utils.make_project_function(4)
