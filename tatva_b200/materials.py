"""Energy densities with fused CUDA kernels (first and second variations in FP64 registers).

In the reference the density is user code differentiated by JAX; the ones the configs name are
pinned by the reference's tests.  Each class carries the kernel id and the parameter vector that
`tatva_energy / tatva_residual / tatva_hvp / tatva_csr_assemble` take.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class LinearElastic:
    """psi = 1/2 sigma:eps, sigma = 2 mu eps + lambda tr(eps) I  (reference tests/test_sparse.py:20-38)."""

    mu: float
    lmbda: float
    material_id = _lib.LINEAR_ELASTIC

    @classmethod
    def from_youngs_poisson_2d(cls, E, nu, plane_stress=False):
        """reference tests/test_sparse_benchmark.py:30-43."""
        mu = E / 2 / (1 + nu)
        lmbda = 2 * nu * mu / (1 - nu) if plane_stress else E * nu / (1 - 2 * nu) / (1 + nu)
        return cls(mu=mu, lmbda=lmbda)

    def params(self):
        return (self.mu, self.lmbda)

    def dofs_per_node(self, dim):
        return dim


@dataclass(frozen=True)
class NeoHookean:
    """psi = mu/2 (I1 - 3 - 2 ln J) + lambda/2 (ln J)^2, F = I + grad u
    (reference tests/test_sparse_tracer.py:103-115)."""

    mu: float
    lmbda: float
    material_id = _lib.NEO_HOOKEAN

    def params(self):
        return (self.mu, self.lmbda)

    def dofs_per_node(self, dim):
        return 3


@dataclass(frozen=True)
class NeoHookeanPhaseField:
    """Two-field AT2 law for the compound (u, phi) state, nodal layout [ux, uy, uz, phi]:
    psi = ((1-phi)^2 + k) psi_NH(grad u) + Gc (phi^2 / (2 l) + l/2 |grad phi|^2)."""

    mu: float
    lmbda: float
    Gc: float
    ell: float
    k: float = 1e-6
    material_id = _lib.NEO_HOOKEAN_PHASE_FIELD

    def params(self):
        return (self.mu, self.lmbda, self.Gc, self.ell, self.k)

    def dofs_per_node(self, dim):
        return 4
