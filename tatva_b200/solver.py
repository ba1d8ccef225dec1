"""Device-resident conjugate gradient and Newton step around the matrix-free HVP.

The reference ships no solver: its docstrings hand `H v` to an external one (tatva/mpi.py:594-595, :615-616).
BASELINE.json's config 3 is "matrix-free HVP inside a CG/Newton step", so this module provides that step in the
B200 style: every vector stays in HBM, the CG scalars (r.r, p.Ap) stay in device memory, one iteration is
[operator application] + `tatva_cg_after_matvec` (six small kernels, deterministic two-pass dots), and the whole
iteration is captured once in a CUDA graph and replayed; the host only looks at the residual norm every
`check_every` iterations.
"""
from __future__ import annotations

import math
from typing import Callable

import torch

from . import _lib


class ConjugateGradient:
    """Solve A x = b for a symmetric positive definite operator given as `matvec(p, out)` (writes A p into `out`
    in place, on torch's current stream, without allocating)."""

    def __init__(self, matvec: Callable, n: int, device, use_graph: bool = True, jacobi: bool = False, matvec_dot: Callable | None = None, fixed_map: torch.Tensor | None = None):
        """`matvec_dot(p, out, partials, scalars)` (optional, e.g. `ReducedOperator.matvec_dot`): an operator application
        that ACCUMULATES A p into a zeroed `out`, leaves p.Ap in scalars[1] and rolls scalars[0] <- scalars[2]
        (`tatva_hvp_lifted_dot`).  The iteration is then application + `tatva_cg_after_dot`: no separate p.Ap pass, no
        memset (the direction pass clears `out`), 6 launches instead of 9."""
        self.matvec, self.n, self.device = matvec, int(n), torch.device(device)
        self.matvec_dot = matvec_dot
        # fixed_map (int32, n entries, < 0 = Fixed DOF; needs matvec_dot): full-size vectors, Dirichlet rows of r held at 0
        self.fixed_map = fixed_map
        if fixed_map is not None and matvec_dot is None:
            raise ValueError("fixed_map needs the matvec_dot iteration")
        mk = lambda m: torch.zeros(m, dtype=torch.float64, device=self.device)  # noqa: E731
        self.x, self.r, self.p, self.Ap = mk(n), mk(n), mk(n), mk(n)
        self.scalars, self.partials = mk(8), mk(2 * 1184)
        # Jacobi preconditioner: 1 / diag(A), refreshed in place by `set_diagonal` (so a captured graph stays valid)
        self.minv = torch.ones(n, dtype=torch.float64, device=self.device) if jacobi else None
        self.use_graph = use_graph
        self._graph = None
        self._L = _lib.lib()

    def set_diagonal(self, diag: torch.Tensor) -> None:
        """minv <- 1 / diag (entries that are not positive and finite fall back to 1)."""
        if self.minv is None:
            raise ValueError("ConjugateGradient was built without jacobi=True")
        d = diag.reshape(-1)
        if d.numel() != self.n or not d.is_contiguous():
            raise ValueError("diagonal must be a contiguous vector of the operator's size")
        _lib.check(self._L.tatva_pcg_reciprocal(d.data_ptr(), self.n, self.minv.data_ptr(), self._stream()), "tatva_pcg_reciprocal")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dot(self, a, b, slot):
        _lib.check(self._L.tatva_cg_dot(a.data_ptr(), b.data_ptr(), self.n, self.partials.data_ptr(), self.scalars.data_ptr(), slot, self._stream()), "tatva_cg_dot")

    def _iteration(self):
        if self.matvec_dot is not None:
            self.matvec_dot(self.p, self.Ap, self.partials, self.scalars)
            _lib.check(
                self._L.tatva_cg_after_dot(self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), self.Ap.data_ptr(), self.minv.data_ptr() if self.minv is not None else None,
                                           self.Ap.data_ptr(), self.fixed_map.data_ptr() if self.fixed_map is not None else None, self.n, self.partials.data_ptr(), self.scalars.data_ptr(), self._stream()),
                "tatva_cg_after_dot",
            )
            return
        self.matvec(self.p, self.Ap)
        if self.minv is not None:
            _lib.check(
                self._L.tatva_pcg_after_matvec(self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), self.Ap.data_ptr(), self.minv.data_ptr(), self.n, self.partials.data_ptr(), self.scalars.data_ptr(), self._stream()),
                "tatva_pcg_after_matvec",
            )
            return
        _lib.check(
            self._L.tatva_cg_after_matvec(self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), self.Ap.data_ptr(), self.n, self.partials.data_ptr(), self.scalars.data_ptr(), self._stream()),
            "tatva_cg_after_matvec",
        )

    def _capture(self):
        # warm up on a side stream (lazy initialisation must not happen inside the capture), then capture
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            saved = [t.clone() for t in (self.x, self.r, self.p, self.scalars)]
            self._iteration()
            for t, s in zip((self.x, self.r, self.p, self.scalars), saved):
                t.copy_(s)
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._iteration()
        self._graph = g

    def solve(self, b: torch.Tensor, x0: torch.Tensor | None = None, tol: float = 1e-10, maxiter: int = 1000, check_every: int = 10):
        """Returns (x, info) with info = dict(iterations, residual_norm, converged).  Stops when
        ||r|| <= tol * ||b||."""
        with torch.cuda.device(self.device):
            if x0 is None:
                self.x.zero_()
                self.r.copy_(b)
                if self.fixed_map is not None:
                    self.r.mul_((self.fixed_map >= 0).to(torch.float64))
            elif self.fixed_map is not None:
                raise NotImplementedError("a starting guess with fixed_map")
            else:
                self.x.copy_(x0)
                self.matvec(self.x, self.Ap)
                torch.sub(b, self.Ap, out=self.r)
            self._dot(self.r if self.fixed_map is not None else b, self.r if self.fixed_map is not None else b, 3)
            if self.minv is not None:
                self._dot(self.r, self.r, 4)
                _lib.check(self._L.tatva_pcg_start(self.p.data_ptr(), self.r.data_ptr(), self.minv.data_ptr(), self.n, self.partials.data_ptr(), self.scalars.data_ptr(), self._stream()), "tatva_pcg_start")
                rr_slot = 4
            else:
                self.p.copy_(self.r)
                self._dot(self.r, self.r, 0)
                rr_slot = 0
            if self.matvec_dot is not None:  # the fused application rolls s[0] <- s[2] and accumulates into a zeroed Ap
                self.scalars[2:3].copy_(self.scalars[0:1])
                self.Ap.zero_()
            rr0, bb = (float(v) for v in self.scalars[[rr_slot, 3]].tolist())
            if bb == 0.0 or math.sqrt(rr0) <= tol * math.sqrt(bb):
                return self.x.clone(), dict(iterations=0, residual_norm=math.sqrt(rr0), converged=True)
            if self.use_graph and self._graph is None:
                saved = [t.clone() for t in (self.x, self.r, self.p, self.scalars)]
                self._capture()  # capturing replays nothing; the first graph launch is iteration 1
                for t, s in zip((self.x, self.r, self.p, self.scalars), saved):
                    t.copy_(s)
            it, rr = 0, rr0
            while it < maxiter:
                for _ in range(min(check_every, maxiter - it)):
                    if self._graph is not None:
                        self._graph.replay()
                    else:
                        self._iteration()
                    it += 1
                rr = float(self.scalars[rr_slot])
                if not math.isfinite(rr) or math.sqrt(rr) <= tol * math.sqrt(bb):
                    break
            return self.x.clone(), dict(iterations=it, residual_norm=math.sqrt(max(rr, 0.0)), converged=math.isfinite(rr) and math.sqrt(rr) <= tol * math.sqrt(bb))


class DistributedConjugateGradient:
    """CG over the ranks of a `PartitionedOperator`: A = H(u) restricted to the DOFs that are not pinned, every rank
    holding its owned block.  Per iteration: one distributed HVP (halo exchange inside) and two one-scalar
    all-reduces on the device; nothing returns to the host except the residual norm every
    `check_every` iterations, which is identical on all ranks, so they stop together.

    `pinned_owned`: bool mask (n_owned) of Dirichlet DOFs (their solution entries stay 0, like the homogeneous lift).
    """

    def __init__(self, pop, pinned_owned=None, jacobi_diagonal=None):
        import torch.distributed as dist

        self.pop, self.dist = pop, dist
        dev, n = pop.device, pop.n_owned
        self.n, self.device = n, dev
        mk = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        peer = pop.halo == "peer" and pop.comm.size > 1
        # p and Ap are LOCAL vectors (owned + ghosts): the operator refreshes the ghosts of p and assembles Ap
        self.p = pop.new_symmetric_vector() if peer else pop.new_local_vector()
        self.Ap = pop.new_symmetric_vector() if peer else pop.new_local_vector()
        self.x, self.r = mk(n), mk(n)
        self.scalars, self.partials = mk(8), mk(2 * 1184)
        self.free = None if pinned_owned is None else (~torch.as_tensor(pinned_owned, device=dev).bool()).to(torch.float64)
        self._L = _lib.lib()
        self.minv = None
        if jacobi_diagonal is not None:
            self.minv = mk(n)
            self.set_diagonal(jacobi_diagonal)
        self.u_local = None

    def set_diagonal(self, diag_owned: torch.Tensor) -> None:
        """Refresh the Jacobi preconditioner in place (minv <- 1 / diag on the owned DOFs)."""
        if self.minv is None:
            raise ValueError("DistributedConjugateGradient was built without a Jacobi diagonal")
        d = diag_owned.contiguous()
        if d.numel() != self.n:
            raise ValueError("the diagonal must cover the owned DOFs")
        _lib.check(self._L.tatva_pcg_reciprocal(d.data_ptr(), self.n, self.minv.data_ptr(), self._stream()), "tatva_pcg_reciprocal")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _allreduce(self, lo, hi):
        if self.pop.comm.size > 1:
            self.dist.all_reduce(self.scalars[lo:hi], group=self.pop.comm.group)

    def set_state(self, u_local: torch.Tensor):
        """`u_local`: the state with its ghost values in place (PartitionedOperator.fill_ghosts)."""
        self.u_local = u_local

    def _matvec(self):
        self.pop.hvp(self.u_local, self.p, self.Ap)
        if self.free is not None:
            self.Ap[: self.n].mul_(self.free)

    def solve(self, b_owned: torch.Tensor, tol: float = 1e-10, maxiter: int = 1000, check_every: int = 10):
        L, n, s = self._L, self.n, self.scalars
        mv = self.minv.data_ptr() if self.minv is not None else None
        rr_slot = 4 if self.minv is not None else 0
        with torch.cuda.device(self.device):
            self.x.zero_()
            self.r.copy_(b_owned)
            if self.free is not None:
                self.r.mul_(self.free)
            st = self._stream()
            _lib.check(L.tatva_cg_dot(self.r.data_ptr(), self.r.data_ptr(), n, self.partials.data_ptr(), s.data_ptr(), 5, st), "tatva_cg_dot")
            self._allreduce(5, 6)  # slot 5 = b.b (global)
            s[3:4].zero_()
            self.p.zero_()
            if self.minv is not None:
                _lib.check(L.tatva_pcg_start(self.p.data_ptr(), self.r.data_ptr(), mv, n, self.partials.data_ptr(), s.data_ptr(), st), "tatva_pcg_start")
                self._allreduce(0, 1)
                s[4:5].copy_(s[5:6])
            else:
                self.p[:n].copy_(self.r)
                s[0:1].copy_(s[5:6])
            bb = float(s[5])
            if bb == 0.0:
                return self.x.clone(), dict(iterations=0, residual_norm=0.0, converged=True)
            it, rr = 0, bb
            while it < maxiter:
                for _ in range(min(check_every, maxiter - it)):
                    self._matvec()
                    st = self._stream()
                    _lib.check(L.tatva_cg_dot(self.p.data_ptr(), self.Ap.data_ptr(), n, self.partials.data_ptr(), s.data_ptr(), 1, st), "tatva_cg_dot")
                    self._allreduce(1, 2)
                    _lib.check(L.tatva_cg_update(self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), self.Ap.data_ptr(), mv, n, self.partials.data_ptr(), s.data_ptr(), st), "tatva_cg_update")
                    if self.minv is not None:
                        self._allreduce(2, 5)  # r.z (slot 2) and r.r (slot 4) in one call; slot 3 is kept at 0
                    else:
                        self._allreduce(2, 3)
                    _lib.check(L.tatva_cg_direction(self.p.data_ptr(), self.r.data_ptr(), mv, n, s.data_ptr(), st), "tatva_cg_direction")
                    it += 1
                rr = float(s[rr_slot])
                if not math.isfinite(rr) or math.sqrt(rr) <= tol * math.sqrt(bb):
                    break
            ok = math.isfinite(rr) and math.sqrt(rr) <= tol * math.sqrt(bb)
            return self.x.clone(), dict(iterations=it, residual_norm=math.sqrt(max(rr, 0.0)), converged=ok)


def distributed_newton_solve(pop, u_local, pinned_owned=None, *, tol: float = 1e-8, max_newton: int = 20, cg_tol: float = 1e-10, cg_maxiter: int = 2000, jacobi: bool = False):
    """Newton's method on E(u) over the ranks of a `PartitionedOperator`.  `u_local` (local vector: symmetric memory
    for halo="peer") holds the initial state INCLUDING the prescribed values on the pinned DOFs, which stay fixed;
    every linear solve is a `DistributedConjugateGradient` on the free DOFs.  Updates `u_local` in place and returns
    (u_local, history); the residual norms are global (all-reduced), so every rank takes the same decisions."""
    import torch.distributed as dist

    n = pop.n_owned
    dev = pop.device
    peer = pop.halo == "peer" and pop.comm.size > 1
    r_local = pop.new_symmetric_vector() if peer else pop.new_local_vector()
    free = None if pinned_owned is None else (~torch.as_tensor(pinned_owned, device=dev).bool()).to(torch.float64)
    history, r0 = [], None
    cg, d_local = None, None
    for k in range(max_newton):
        pop.fill_ghosts(u_local)
        pop.residual(u_local, r_local)
        r = r_local[:n] * free if free is not None else r_local[:n].clone()
        rr = (r * r).sum().reshape(1)
        if pop.comm.size > 1:
            dist.all_reduce(rr, group=pop.comm.group)
        rn = math.sqrt(float(rr))
        r0 = rn if r0 is None else r0
        if rn <= tol * max(r0, 1e-300) or rn == 0.0:
            history.append(dict(newton=k, residual_norm=rn, cg_iterations=0))
            break
        if jacobi:
            d_local = pop.hessian_diagonal(u_local, d_local)
        if cg is None:
            cg = DistributedConjugateGradient(pop, pinned_owned=pinned_owned, jacobi_diagonal=d_local[:n] if jacobi else None)
        elif jacobi:
            cg.set_diagonal(d_local[:n])
        cg.set_state(u_local)
        du, info = cg.solve(-r, tol=cg_tol, maxiter=cg_maxiter)
        history.append(dict(newton=k, residual_norm=rn, cg_iterations=info["iterations"], cg_converged=info["converged"]))
        u_local[:n].add_(du)
    return u_local, history


class MaskedOperator:
    """The tangent of E on FULL-size vectors for a Lifter that holds only Fixed constraints: K v with v = 0 on the Fixed
    DOFs, the Fixed rows of the result masked by the CG's update pass (`fixed_map`).  It runs the UNCONSTRAINED HVP kernel
    (no index-map gathers in the element kernel: 0.397 ms against 0.481 ms for the lifted one at config 3); the solution
    on the free DOFs is `lifter.reduce(x_full)`.  `fuse_dot`: let the element kernel sum v . K v itself."""

    def __init__(self, op, material, lifter, fuse_dot: bool = False):
        import numpy as np

        self.op, self.material, self.lifter, self.fuse_dot = op, material, lifter, bool(fuse_dot)
        m = lifter.dof_map()
        free = np.asarray(lifter.free_dofs)
        if (m >= 0).sum() != lifter.size_reduced or not np.array_equal(m[free], np.arange(lifter.size_reduced)):
            raise NotImplementedError("MaskedOperator: only lifters made of Fixed constraints (use ReducedOperator otherwise)")
        self.fixed_map = torch.as_tensor(m.astype(np.int32), device=op.device)
        self.free = torch.as_tensor(free, device=op.device)
        self.u_full = torch.zeros(lifter.size, dtype=torch.float64, device=op.device)
        self.n = lifter.size

    def set_state(self, u_reduced: torch.Tensor):
        self.lifter.lift_from_zeros(u_reduced, out=self.u_full)

    def expand(self, b_reduced: torch.Tensor) -> torch.Tensor:
        """Right-hand side on the free DOFs -> full-size vector with zeros on the Fixed DOFs."""
        out = torch.zeros(self.n, dtype=torch.float64, device=self.op.device)
        out[self.free] = b_reduced
        return out

    def restrict(self, x_full: torch.Tensor) -> torch.Tensor:
        return x_full[self.free]

    def matvec(self, v_full: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        self.op._raw_hvp(self.material, self.u_full, v_full, out=out)
        out.mul_((self.fixed_map >= 0).to(torch.float64))
        return out

    def matvec_dot(self, v_full: torch.Tensor, out: torch.Tensor, partials: torch.Tensor, scalars: torch.Tensor) -> torch.Tensor:
        prm, n = _lib.params_array(self.material.params())
        self.op._call("tatva_hvp_dot", self.material.material_id, prm, n, self.u_full.data_ptr(), v_full.data_ptr(), out.data_ptr(), 0, int(self.fuse_dot),
                      partials.data_ptr(), scalars.data_ptr(), 1, 1)
        return out

    def solver(self, use_graph: bool = True, jacobi: bool = False) -> "ConjugateGradient":
        return ConjugateGradient(self.matvec, self.n, self.op.device, use_graph=use_graph, jacobi=jacobi, matvec_dot=self.matvec_dot, fixed_map=self.fixed_map)


class ReducedOperator:
    """The constrained tangent and residual of E(u) on the free DOFs of a Lifter:
        r_red(u_red) = reduce_adjoint(dE/du(lift(u_red))),    K_red v = reduce_adjoint(H(lift(u_red)) lift_0(v)),
    with work vectors preallocated so that `matvec` can sit inside a CUDA graph."""

    def __init__(self, op, material, lifter, fused: bool = True):
        self.op, self.material, self.lifter = op, material, lifter
        self.hom = lifter.homogeneous()
        # fused: lift_0 and reduce_adjoint happen inside the HVP kernel's gather / scatter (one launch per matvec)
        n_nodes, dim = op.mesh.coords.shape
        self.dof_map = lifter.dof_map(op.device) if fused and lifter.size == n_nodes * material.dofs_per_node(dim) else None
        dev = op.device
        self.u_full = torch.zeros(lifter.size, dtype=torch.float64, device=dev)
        self.v_full = torch.zeros(lifter.size, dtype=torch.float64, device=dev)
        self.y_full = torch.zeros(lifter.size, dtype=torch.float64, device=dev)

    def set_state(self, u_reduced: torch.Tensor):
        self.lifter.lift_from_zeros(u_reduced, out=self.u_full)

    def residual(self, out: torch.Tensor | None = None) -> torch.Tensor:
        r_full = self.op._raw_residual(self.material, self.u_full)
        return self.lifter.reduce_adjoint(r_full, out=out)

    def energy(self) -> float:
        return float(self.op._raw_energy(self.material, self.u_full))

    def diagonal(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """Jacobi diagonal of the reduced tangent: reduce_adjoint(diag H(u)).  Exact for Fixed constraints; for Periodic
        pairs it omits the coupling between a node and its image (zero unless they share an element), which a
        preconditioner does not need."""
        d_full = self.op.hessian_diagonal(self.material, self.u_full)
        return self.lifter.reduce_adjoint(d_full, out=out)

    def matvec_dot(self, v_reduced: torch.Tensor, out: torch.Tensor, partials: torch.Tensor, scalars: torch.Tensor) -> torch.Tensor:
        """out (already zero) += K_red v, scalars[1] = v . K_red v, scalars[0] <- scalars[2]: one element kernel + one
        final-sum kernel (`tatva_hvp_lifted_dot`).  Needs the fused lifter path (`dof_map`)."""
        if self.dof_map is None:
            raise ValueError("matvec_dot needs the fused lifter path (ReducedOperator(fused=True) on a nodal layout)")
        prm, n = _lib.params_array(self.material.params())
        self.op._call("tatva_hvp_lifted_dot", self.material.material_id, prm, n, self.u_full.data_ptr(), v_reduced.data_ptr(), self.dof_map.data_ptr(), out.numel(), out.data_ptr(), 0,
                      partials.data_ptr(), scalars.data_ptr(), 1, 1)
        return out

    def matvec(self, v_reduced: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if self.dof_map is not None:
            return self.op._raw_hvp_lifted(self.material, self.u_full, v_reduced.contiguous(), self.dof_map, out)
        self.hom.lift_from_zeros(v_reduced, out=self.v_full)
        self.op._raw_hvp(self.material, self.u_full, self.v_full, out=self.y_full)
        return self.lifter.reduce_adjoint(self.y_full, out=out)


def newton_solve(op, material, lifter, u0_reduced=None, *, tol: float = 1e-8, max_newton: int = 20, cg_tol: float = 1e-10, cg_maxiter: int = 2000, use_graph: bool = True, jacobi: bool = False):
    """Minimise E(lift(u_red)) by Newton's method; every linear solve is a matrix-free CG on the HVP kernel.
    Returns (u_reduced, history) with history = list of dict(newton, residual_norm, cg_iterations)."""
    dev = op.device
    red = ReducedOperator(op, material, lifter)
    n = lifter.size_reduced
    u = torch.zeros(n, dtype=torch.float64, device=dev) if u0_reduced is None else torch.as_tensor(u0_reduced, dtype=torch.float64, device=dev).clone()
    cg = ConjugateGradient(red.matvec, n, dev, use_graph=use_graph, jacobi=jacobi, matvec_dot=red.matvec_dot if red.dof_map is not None else None)
    r = torch.empty(n, dtype=torch.float64, device=dev)
    diag = torch.empty(n, dtype=torch.float64, device=dev) if jacobi else None
    history = []
    r0 = None
    for k in range(max_newton):
        red.set_state(u)
        red.residual(out=r)
        rn = float(r.norm())
        r0 = rn if r0 is None else r0
        if rn <= tol * max(r0, 1e-300) or rn == 0.0:
            history.append(dict(newton=k, residual_norm=rn, cg_iterations=0))
            break
        if jacobi:
            cg.set_diagonal(red.diagonal(out=diag))
        du, info = cg.solve(-r, tol=cg_tol, maxiter=cg_maxiter)
        history.append(dict(newton=k, residual_norm=rn, cg_iterations=info["iterations"], cg_converged=info["converged"]))
        u = u + du
    return u, history
