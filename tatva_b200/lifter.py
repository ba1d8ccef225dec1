"""Lifter — mirror of tatva.lifter (tatva/lifter/base.py, constraints.py, common.py).

Same surface as the reference: `Lifter(size, *constraints)` with `free_dofs`, `constrained_dofs`, `size`,
`size_reduced`, `lift`, `lift_from_zeros`, `reduce`, `reduce_adjoint`, `with_values`, `at[key].set(...)`,
`add`, `adapt_sparsity / augment_sparsity / reduce_sparsity`; constraints `Fixed`, `Periodic`;
`RuntimeValue`, `LifterError`, the `lifted` decorator.

B200 form: the reference applies the constraints one scatter at a time on the full vector
(lifter/base.py:201-229: `.at[free].set`, then every `apply_lift`; :235-251 the transposes in reverse).
Here the constraint chain is composed ONCE on the host into a source table (full DOF -> reduced DOF | constant
| base value), so on CUDA tensors `lift` is one gather kernel (`tatva_lift`) and `reduce_adjoint` one
deterministic segmented sum (`tatva_reduce_adjoint`); `reduce` is one pack kernel.  NumPy inputs take the
same composed tables through NumPy indexing (host-side set-up and tests).

`Lifter.adapt_layout` and `PeriodicMPI` (lifter/base.py:333-425, constraints.py:223-287) are the distributed
host logic: `comm` is a torch.distributed process group (see tatva_b200.mpi).
"""
from __future__ import annotations

from typing import Callable, Hashable
from uuid import uuid4

import numpy as np
import scipy.sparse as sps

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None

from . import _lib


BASE_OFF = 1 << 40  # source-table codes <= -BASE_OFF mean "entry -(code + BASE_OFF) of the base vector"


class LifterError(ValueError):
    """Problem with the lifter, e.g. a missing runtime value (lifter/common.py:41-42)."""


class RuntimeValue:
    """A constraint value provided at run time under `key` (lifter/common.py:45-64)."""

    def __init__(self, key: Hashable | None = None, default=None):
        self.key = key or uuid4()
        self.default = default

    def get_value(self, runtime_values):
        if self.key not in runtime_values:
            raise LifterError(f"Runtime value for (key={self.key}) not set on lifter")
        return runtime_values[self.key]


def _np(a):
    if torch is not None and isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


class Constraint:
    """Base class (lifter/constraints.py:58-136).  Identity-hashed like the reference."""

    def __init__(self, dofs):
        self.dofs = np.asarray(_np(dofs), dtype=np.int64).reshape(-1)
        self._constraint_id = uuid4()

    def __eq__(self, other):
        return type(self) is type(other) and self._constraint_id == other._constraint_id

    def __hash__(self):
        return hash((type(self), self._constraint_id))

    def augment_sparsity(self, sparsity):
        return sparsity

    def _runtime_specs(self):
        return ()

    def _resolve_indices(self, layout):
        """Re-express the constraint in the local indices of `layout` (only PeriodicMPI needs to)."""
        return self

    # symbolic application to the source table (kind, payload): see Lifter._compose
    def _compose(self, src, consts, runtime_values):
        raise NotImplementedError


class Fixed(Constraint):
    """Dirichlet values on `dofs` (lifter/constraints.py:290-318); `values` may be a RuntimeValue."""

    def __init__(self, dofs, values=None):
        super().__init__(dofs)
        self.values = values if values is not None else 0.0

    def _runtime_specs(self):
        return (self.values,) if isinstance(self.values, RuntimeValue) else ()

    def _compose(self, src, consts, runtime_values):
        vals = self.values.get_value(runtime_values) if isinstance(self.values, RuntimeValue) else self.values
        vals = np.broadcast_to(np.asarray(_np(vals), dtype=np.float64), self.dofs.shape)
        start = len(consts)
        consts.extend(vals.tolist())
        src[self.dofs] = -(start + np.arange(len(self.dofs), dtype=np.int64) + 2)


class Periodic(Constraint):
    """`dofs` follow `master_dofs` (lifter/constraints.py:184-221)."""

    def __init__(self, dofs, master_dofs):
        super().__init__(dofs)
        self.master_dofs = np.asarray(_np(master_dofs), dtype=np.int64).reshape(-1)

    def augment_sparsity(self, sparsity):
        """S_aug = M^T S M with M = I + (slave <- master) (lifter/constraints.py:195-212)."""
        n = sparsity.shape[0]
        rows = np.concatenate([np.arange(n), self.dofs])
        cols = np.concatenate([np.arange(n), self.master_dofs])
        M = sps.csr_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n, n))
        S = (M.T @ (sparsity.astype(np.int8) @ M)).tocsr()
        S.data = np.ones_like(S.data, dtype=np.int8)
        return S

    def _compose(self, src, consts, runtime_values):
        # RHS gathered before the set, like u.at[dofs].set(u[master]) (lifter/constraints.py:214-221).  A master that
        # still reads the base vector (it is constrained by a LATER constraint, or is itself a slave of this one) hands
        # its slave the code "base entry of the master" — what the reference's sequential application gives.
        src[self.dofs] = src[self.master_dofs]


def create_g2l(l2g):
    """Lookup natural-global -> local index, -1 where absent (tatva/utils.py:265-280)."""
    l2g = np.asarray(l2g)
    order = np.argsort(l2g)
    sorted_l2g = l2g[order]

    def lookup(global_indices):
        gi = np.asarray(global_indices)
        if sorted_l2g.size == 0:
            return np.full_like(gi, -1)
        pos = np.searchsorted(sorted_l2g, gi)
        ok = (pos < sorted_l2g.size) & (sorted_l2g[np.minimum(pos, sorted_l2g.size - 1)] == gi)
        out = np.full_like(gi, -1)
        out[ok] = order[pos[ok]]
        return out

    return lookup


class PeriodicMPI(Periodic):
    """Periodicity between natural-global DOF ids that may live on different ranks
    (lifter/constraints.py:223-287).  Masters of local slaves that this rank does not hold become extra
    ghost DOFs of the layout (`Lifter.adapt_layout`)."""

    def __init__(self, dofs, master_dofs, layout, *, comm=None):
        self._comm = comm
        d = np.asarray(_np(dofs), dtype=np.int64).reshape(-1)
        m = np.asarray(_np(master_dofs), dtype=np.int64).reshape(-1)
        lookup = create_g2l(layout.natural_l2g)
        local = lookup(d)
        here = local >= 0
        self._slave_natural_g, self._master_natural_g = d[here], m[here]
        self._extra_ghost_dofs = self._master_natural_g[lookup(self._master_natural_g) < 0]
        super().__init__(local[here], np.zeros(int(here.sum()), dtype=np.int64))

    def _resolve_indices(self, layout):
        lookup = create_g2l(layout.natural_l2g)
        s_loc, m_loc = lookup(self._slave_natural_g), lookup(self._master_natural_g)
        if (s_loc < 0).any() or (m_loc < 0).any():
            raise LifterError("PeriodicMPI: failed to resolve local indices for periodic DOFs; make sure all required masters were added as ghosts to the layout")
        out = type(self).__new__(type(self))
        out.__dict__ = dict(self.__dict__)
        out.dofs, out.master_dofs = s_loc.astype(np.int64), m_loc.astype(np.int64)
        return out


class _Indexer:
    def __init__(self, lifter):
        self.lifter = lifter

    def __getitem__(self, key):
        return _Setter(self.lifter, key)

    __call__ = __getitem__


class _Setter:
    def __init__(self, lifter, key):
        self.lifter, self.key = lifter, key

    def set(self, value):
        return self.lifter.with_values({self.key: value})


class Lifter:
    def __init__(self, size: int, /, *constraints: Constraint):
        self.size = int(size)
        self.constraints = tuple(constraints)
        specs = [s for c in self.constraints for s in c._runtime_specs()]
        self._runtime_keys = tuple(s.key for s in specs)
        self._runtime_values = {s.key: s.default for s in specs if s.default is not None}
        all_dofs = np.arange(self.size, dtype=np.int64)
        if self.constraints:
            self.constrained_dofs = np.unique(np.concatenate([c.dofs for c in self.constraints]))
            self.free_dofs = np.setdiff1d(all_dofs, self.constrained_dofs, assume_unique=True)
        else:
            self.constrained_dofs = np.array([], dtype=np.int32)
            self.free_dofs = all_dofs
        self.size_reduced = int(self.free_dofs.size)
        self._tables = None
        self._dev = {}

    # -- identity -----------------------------------------------------------------------------------
    def __hash__(self):
        return hash((self.size, self.constraints))

    def __eq__(self, other):
        if not (isinstance(other, Lifter) and self.size == other.size and self.constraints == other.constraints):
            return False
        a, b = self._runtime_values, other._runtime_values
        return a.keys() == b.keys() and all(a[k] is b[k] or np.array_equal(_np(a[k]), _np(b[k])) for k in a)

    @property
    def at(self):
        return _Indexer(self)

    def add(self, condition: Constraint):
        return type(self)(self.size, *self.constraints, condition)

    def with_values(self, updates: dict):
        for key in updates:
            if key not in self._runtime_keys:
                raise LifterError(f"There is no runtime value with key={key} in the lifter's constraints")
        new = type(self).__new__(type(self))
        new.__dict__ = dict(self.__dict__)
        new._runtime_values = {**self._runtime_values, **updates}
        new._tables, new._dev, new._homogeneous = None, {}, None
        return new

    # -- composed tables ------------------------------------------------------------------------------
    def _compose(self):
        """src[i] >= 0: reduced index; <= -BASE_OFF: entry -(src + BASE_OFF) of the base vector (initially its own);
        -2 ... : consts[-(src+2)].  inverse (ptr, list): for every reduced DOF the full DOFs that read it, ascending."""
        if self._tables is None:
            src = -(BASE_OFF + np.arange(self.size, dtype=np.int64))
            src[self.free_dofs] = np.arange(self.size_reduced, dtype=np.int64)
            consts: list[float] = []
            for c in self.constraints:
                c._compose(src, consts, self._runtime_values)
            readers = np.where(src >= 0)[0]
            order = np.argsort(src[readers], kind="stable")
            lst = readers[order].astype(np.int64)
            ptr = np.concatenate([[0], np.cumsum(np.bincount(src[readers], minlength=self.size_reduced))]).astype(np.int64)
            self._tables = (src, np.asarray(consts, dtype=np.float64), ptr, lst)
        return self._tables

    def _device_tables(self, device):
        if device not in self._dev:
            src, consts, ptr, lst = self._compose()
            t = lambda a: torch.as_tensor(a, device=device)  # noqa: E731
            self._dev[device] = (t(src), t(consts if consts.size else np.zeros(1)), t(ptr), t(lst), t(self.free_dofs.astype(np.int64)))
        return self._dev[device]

    def dof_map(self, device=None):
        """int32 table full DOF -> reduced DOF that drives it (-1: Fixed / untouched base entry): the homogeneous
        lift and its transpose as one index table, consumed by `tatva_hvp_lifted` inside the HVP kernel."""
        src = self._compose()[0]
        if self.size_reduced > np.iinfo(np.int32).max:
            raise LifterError("dof_map needs int32 reduced indices")
        m = np.where(src >= 0, src, -1).astype(np.int32)
        if device is None:
            return m
        key = ("dof_map", torch.device(device))
        if key not in self._dev:
            self._dev[key] = torch.as_tensor(m, device=device)
        return self._dev[key]

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    # -- reference API -----------------------------------------------------------------------------------
    def lift(self, u_reduced, u_full, out=None):
        """Full vector with free DOFs = u_reduced and every constraint applied (lifter/base.py:201-229).
        `out` (CUDA only) receives the result without allocating (CUDA-graph friendly)."""
        if torch is not None and isinstance(u_reduced, torch.Tensor) and u_reduced.is_cuda:
            src, consts, _, _, _ = self._device_tables(u_reduced.device)
            ur = u_reduced.contiguous()
            base = None if u_full is None else torch.as_tensor(u_full, device=ur.device, dtype=ur.dtype).contiguous()
            if out is None:
                out = torch.empty(self.size, dtype=ur.dtype, device=ur.device)
            with torch.cuda.device(ur.device):
                _lib.check(_lib.lib().tatva_lift(ur.data_ptr(), src.data_ptr(), consts.data_ptr(), base.data_ptr() if base is not None else None, self.size, out.data_ptr(), self._stream()), "tatva_lift")
            return out[: self._local_size] if self._local_size != self.size else out
        _lib.host_tables_only("Lifter.lift")
        src, consts, _, _ = self._compose()
        ur = _np(u_reduced)
        out = np.zeros(self.size, dtype=ur.dtype) if u_full is None else np.array(_np(u_full), dtype=ur.dtype, copy=True)
        base = None if u_full is None else np.array(_np(u_full), dtype=ur.dtype, copy=True)
        red = src >= 0
        out[red] = ur[src[red]]
        cst = (src <= -2) & (src > -BASE_OFF)
        out[cst] = consts[-(src[cst] + 2)]
        bas = src <= -BASE_OFF
        out[bas] = 0.0 if base is None else base[-(src[bas] + BASE_OFF)]
        return out[: self._local_size]  # extra ghost DOFs added by PeriodicMPI are not part of the result

    def lift_from_zeros(self, u_reduced, out=None):
        return self.lift(u_reduced, None, out=out) if out is not None else self.lift(u_reduced, None)

    def reduce(self, u_full):
        """u_full[free_dofs] (lifter/base.py:231-233)."""
        if torch is not None and isinstance(u_full, torch.Tensor) and u_full.is_cuda:
            free = self._device_tables(u_full.device)[4]
            uf = u_full.contiguous()
            out = torch.empty(self.size_reduced, dtype=uf.dtype, device=uf.device)
            with torch.cuda.device(uf.device):
                _lib.check(_lib.lib().tatva_halo_pack(uf.data_ptr(), free.data_ptr(), self.size_reduced, out.data_ptr(), self._stream()), "tatva_halo_pack")
            return out
        _lib.host_tables_only("Lifter.reduce")
        return _np(u_full)[self.free_dofs]

    def reduce_adjoint(self, r_full, out=None):
        """Reduced dual vector: constrained contributions folded back onto the DOFs that drive them
        (lifter/base.py:235-251: transposes in reverse order, then [free_dofs])."""
        if torch is not None and isinstance(r_full, torch.Tensor) and r_full.is_cuda:
            _, _, ptr, lst, _ = self._device_tables(r_full.device)
            rf = r_full.contiguous()
            if out is None:
                out = torch.empty(self.size_reduced, dtype=rf.dtype, device=rf.device)
            with torch.cuda.device(rf.device):
                _lib.check(_lib.lib().tatva_reduce_adjoint(rf.data_ptr(), ptr.data_ptr(), lst.data_ptr(), self.size_reduced, out.data_ptr(), self._stream()), "tatva_reduce_adjoint")
            return out
        _lib.host_tables_only("Lifter.reduce_adjoint")
        src, _, _, _ = self._compose()
        rf = _np(r_full)
        readers = src >= 0
        return np.bincount(src[readers], weights=rf[readers], minlength=self.size_reduced).astype(rf.dtype)

    def homogeneous(self) -> "Lifter":
        """Lifter with the same free DOFs whose constrained values are all zero: the lift of a tangent
        (Newton / CG direction), d(lift)/d(u_reduced)."""
        if getattr(self, "_homogeneous", None) is None:
            hom = []
            for c in self.constraints:
                hom.append(Fixed(c.dofs, 0.0) if isinstance(c, Fixed) else c)
            self._homogeneous = Lifter(self.size, *hom)
            self._homogeneous._nb_extra_ghost_dofs = getattr(self, "_nb_extra_ghost_dofs", None)  # same local size
        return self._homogeneous

    # -- distributed layouts (lifter/base.py:333-425) ------------------------------------------------------
    @property
    def _local_size(self) -> int:
        return self.size - (getattr(self, "_nb_extra_ghost_dofs", None) or 0)

    def adapt_layout(self, layout, comm):
        """(reduced layout of the free DOFs, lifter resolved against the possibly ghost-augmented layout).
        Free owned DOFs get new rank-contiguous global ids; free ghosts are resolved through the all-DOF
        global ids of the (augmented) full layout."""
        from .mpi import _LocalLayout, _as_comm, _create_dof_layout

        comm = _as_comm(comm)
        extra = [c._extra_ghost_dofs for c in self.constraints if hasattr(c, "_extra_ghost_dofs")]
        if extra:
            local_extra = np.unique(np.concatenate(extra))
            full = _create_dof_layout(
                np.concatenate([np.asarray(layout.natural_l2g), local_extra]).astype(np.int32),
                np.concatenate([np.asarray(layout.owned_mask), np.zeros(len(local_extra), dtype=bool)]),
                layout.n_global, comm,
            )
        else:
            local_extra, full = np.zeros(0, dtype=np.int64), layout
        lifter = type(self)(full.n_total, *(c._resolve_indices(full) for c in self.constraints)).with_values(self._runtime_values)
        if extra:
            lifter._nb_extra_ghost_dofs = int(len(local_extra))
        free = lifter.free_dofs
        owned_free = np.asarray(full.owned_mask)[free]
        g_free = np.asarray(full.local_to_global)[free]
        published = comm.allgather(g_free[owned_free])  # all-DOF global ids of every rank's owned free DOFs
        counts = [len(p) for p in published]
        offset = int(sum(counts[: comm.rank]))
        l2g = np.full(free.size, -1, dtype=np.int32)
        l2g[owned_free] = offset + np.arange(int(owned_free.sum()), dtype=np.int32)
        if (~owned_free).any():
            directory = np.full(full.n_global, -1, dtype=np.int32)
            start = 0
            for p in published:
                directory[p] = start + np.arange(len(p), dtype=np.int32)
                start += len(p)
            l2g[~owned_free] = directory[g_free[~owned_free]]
        reduced = _LocalLayout(l2g, offset, int(owned_free.sum()), int(free.size), int(sum(counts)), owned_free, g_free)
        return reduced, lifter

    # -- sparsity (lifter/base.py:281-331) ----------------------------------------------------------------
    def augment_sparsity(self, sparsity):
        if sparsity.shape[0] < self.size:
            extra = np.full(self.size - sparsity.shape[0], sparsity.indptr[-1], dtype=sparsity.indptr.dtype)
            sparsity = sps.csr_matrix((sparsity.data, sparsity.indices, np.concatenate([sparsity.indptr, extra])), shape=(self.size, self.size))
        for c in self.constraints:
            sparsity = c.augment_sparsity(sparsity)
        return sparsity

    def reduce_sparsity(self, sparsity):
        return sparsity[self.free_dofs][:, self.free_dofs]

    def adapt_sparsity(self, sparsity):
        return self.reduce_sparsity(self.augment_sparsity(sparsity))


def lifted(fn: Callable | None = None, *, argnums=0, output: str | None = None):
    """Decorator: `lifted(fn)(lifter, *args)` lifts the reduced arguments `argnums` before calling `fn` and
    optionally reduces the result ("primal": reduce, "dual": reduce_adjoint) — lifter/base.py:444-509."""
    argnums = (argnums,) if isinstance(argnums, int) else tuple(argnums)
    if fn is None:
        return lambda f: lifted(f, argnums=argnums, output=output)

    def lifted_fn(lifter: Lifter, *args, **kwargs):
        largs = list(args)
        for i, a in enumerate(args):
            if i not in argnums:
                continue
            if not hasattr(a, "shape"):
                raise LifterError(f"Argument {i} is not an Array and cannot be lifted by the lifter")
            if tuple(a.shape) != (lifter.size_reduced,):
                raise LifterError(f"Argument {i} has shape {tuple(a.shape)} but expected {(lifter.size_reduced,)} for lifting")
            largs[i] = lifter.lift_from_zeros(a)
        out = fn(*largs, **kwargs)
        if output:
            if not hasattr(out, "shape"):
                raise LifterError("Output is not an Array and cannot be reduced by the lifter")
            if output == "primal":
                return lifter.reduce(out)
            if output == "dual":
                return lifter.reduce_adjoint(out)
            raise LifterError(f"Invalid value for output: {output}")
        return out

    return lifted_fn


__all__ = ["Lifter", "lifted", "Constraint", "Fixed", "Periodic", "PeriodicMPI", "RuntimeValue", "LifterError", "create_g2l"]


def __getattr__(name: str):
    """Deprecated aliases of the reference (tatva/lifter/__init__.py): DirichletBC -> Fixed, PeriodicMap -> Periodic."""
    from warnings import warn

    if name == "DirichletBC":
        warn("`DirichletBC` is deprecated; use `Fixed` instead.", DeprecationWarning, stacklevel=2)
        return Fixed
    if name == "PeriodicMap":
        warn("`PeriodicMap` is deprecated; use `Periodic` instead.", DeprecationWarning, stacklevel=2)
        return Periodic
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
