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


class UserLaw:
    """A constitutive law the USER supplies — the reference's contract (README.md:93): the energy density is user code and
    the framework derives residual and Hessian-vector product from it.

        law = UserLaw.from_psi(lambda G, mu, lam: ..., params=(500.0, 1000.0))       # density written ONCE, on SymPy symbols
        op.energy(law)(u); op.residual(law)(u); op.hvp(law)(u, v); sparse.jacfwd(op.residual(law), cm)(u)

    `from_psi` differentiates the density's straight-line program source-to-source (tatva_b200/lawgen.py: reverse sweep for
    d psi / d G, forward sweep over both for its directional derivative — forward-over-reverse, what
    `jax.jvp(jax.grad(E))` evaluates) and emits a CUDA struct; `UserLaw(source, ...)` takes such a struct directly
    (the `Mat` interface of csrc/common.cuh: psi / first / second).  At first use with an Operator the source is compiled by
    NVRTC into the SAME fused kernel templates as the built-in laws (`k_fused<El, UserLaw, MODE>`) for the operator's
    element; there is no eager / autograd fallback.  `G[c][j] = d u_c / d x_j` (value component first, as Operator.grad)."""

    def __init__(self, source: str, params, *, dim: int = 3, dofs_per_node: int | None = None, uses_values: bool = False, generated=None):
        self.source = source
        self._params = tuple(float(p) for p in params)
        self.dim = int(dim)
        self.dpn = int(dim if dofs_per_node is None else dofs_per_node)
        self.uses_values = bool(uses_values)
        self.generated = generated  # lawgen.GeneratedLaw (operation counts, C twins) when built by from_psi
        self._id = None

    @classmethod
    def from_psi(cls, psi, params, *, dim: int = 3, dofs_per_node: int | None = None, uses_values: bool = False) -> "UserLaw":
        from . import lawgen

        g = lawgen.generate(psi, len(params), dim=dim, dofs_per_node=dofs_per_node, uses_values=uses_values)
        return cls(g.cuda_source(), params, dim=dim, dofs_per_node=dofs_per_node, uses_values=uses_values, generated=g)

    def with_params(self, params) -> "UserLaw":
        """The same compiled law with other parameter values (parameters are kernel arguments, not baked in)."""
        other = UserLaw(self.source, params, dim=self.dim, dofs_per_node=self.dpn, uses_values=self.uses_values, generated=self.generated)
        if len(other._params) != len(self._params):
            raise ValueError("with_params: the number of parameters is part of the compiled law")
        other._id = self.material_id
        return other

    @property
    def material_id(self) -> int:
        if self._id is None:
            import ctypes as C

            out = C.c_int()
            _lib.check(_lib.lib().tatva_law_register(self.source.encode(), self.dim, self.dpn, len(self._params), int(self.uses_values), C.byref(out)), "tatva_law_register")
            self._id = out.value
        return self._id

    def params(self):
        return self._params

    def dofs_per_node(self, dim):
        return self.dpn

    @staticmethod
    def compile_log() -> str:
        import ctypes as C

        buf = C.create_string_buffer(1 << 16)
        _lib.lib().tatva_law_compile_log(buf, len(buf))
        return buf.value.decode(errors="replace")
