"""Communication plans — mirror of tatva.mpi (tatva/mpi.py) on torch.distributed (NCCL over NVLink).

The reference talks MPI: set-up handshakes with mpi4py (`allgather`, `Alltoall`, `Sendrecv`,
`Allreduce(MAX)`), steady state with one blocking `mpi4jax.sendrecv` per neighbour inside the jitted
function (mpi.py:403-405, :509-511).  Here `comm` is a torch.distributed process group (one process
per GPU): the handshakes are object collectives, and the steady-state exchange is

    pack kernel  ->  ONE all_to_all_single (NCCL: a grouped ncclSend/ncclRecv over NVLink)  ->  unpack kernel

with the pack / unpack-set / unpack-add kernels of libtatva_b200.so on CUDA tensors.  Same public
surface as the reference: `ExchangePlan(layout, local_sparsity_pattern=None, comm=...)` with
`make_scatter_fwd_set`, `make_scatter_rev_add`, `owned_csr`, `rstart/rend/local_size/global_size`;
`AllreducePlan(global_size, global_sparsity_pattern=None, comm=...)` with `make_allgather`,
`make_allreduce_owned`; `_create_dof_layout`, `_dof_range`.

With CPU tensors packing is plain tensor indexing: that mode exists for the world_size-2 `gloo`
tests of the host logic only and must be switched on with TATVA_B200_HOST_TABLES=1 (tests/conftest.py
does); otherwise a CPU operand raises — GPU runs always go through the kernels.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Callable, NamedTuple

import numpy as np
import torch
import torch.distributed as dist
from scipy.sparse import csr_matrix

from . import _lib


class Comm:
    """The subset of an MPI communicator the plans need, on a torch.distributed group."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def allgather(self, obj):
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out


class SelfComm:
    """Single-process communicator (world size 1)."""

    rank, size, group = 0, 1, None

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allgather(self, obj):
        return [obj]


def _as_comm(comm):
    if comm is None:
        return Comm() if dist.is_initialized() else SelfComm()
    if isinstance(comm, (Comm, SelfComm)):
        return comm
    return Comm(comm)  # a torch.distributed ProcessGroup


class FieldGlobalInfo(NamedTuple):
    """compound/mpi.py:48-52: where a field lives in the natural (mesh-global) DOF numbering."""

    global_shape: tuple
    global_base_offset: int
    global_strides: tuple
    global_subset: np.ndarray | None = None


class _NeighborRoute(NamedTuple):
    """mpi.py:40-57 (_NeighborDofRoute / _NeighborNnzRoute)."""

    rank: int
    local_send_idx: np.ndarray
    recv_local_idx: np.ndarray
    send_size: int
    recv_size: int


class _HessianLayout(NamedTuple):
    owned_nnz: int
    owned_ptr: np.ndarray
    owned_indices: np.ndarray
    local_send_idx: np.ndarray
    recv_local_idx: np.ndarray
    neighbor_data: list


class _LocalLayout(NamedTuple):
    """mpi.py:71-80."""

    local_to_global: np.ndarray
    offset: int
    n_owned: int
    n_total: int
    n_global: int
    owned_mask: np.ndarray
    natural_l2g: np.ndarray


def _create_dof_layout(natural_dof_map, owned_mask, n_natural_global, comm) -> _LocalLayout:
    """mpi.py:83-130: rank-contiguous owned blocks (prefix sum of owned counts); ghosts resolved
    through a directory keyed by the natural (mesh-global) DOF id."""
    comm = _as_comm(comm)
    natural_dof_map = np.asarray(natural_dof_map, dtype=np.int32)
    owned_mask = np.asarray(owned_mask, dtype=bool)
    owned_idx = np.where(owned_mask)[0]
    n_owned = int(owned_idx.size)
    # one object collective instead of allreduce + allgather + Allreduce(MAX): every rank publishes
    # the natural ids of the DOFs it owns, in owned order
    published = comm.allgather(natural_dof_map[owned_idx])
    counts = [len(p) for p in published]
    n_global = int(sum(counts))
    offset = int(sum(counts[: comm.rank]))
    l2g = np.full(natural_dof_map.size, -1, dtype=np.int32)
    l2g[owned_idx] = offset + np.arange(n_owned, dtype=np.int32)
    ghost_idx = np.where(~owned_mask)[0]
    if ghost_idx.size:
        directory = np.full(n_natural_global, -1, dtype=np.int32)
        start = 0
        for p in published:
            ok = p >= 0
            directory[p[ok]] = (start + np.arange(len(p), dtype=np.int32))[ok]
            start += len(p)
        nat = natural_dof_map[ghost_idx]
        l2g[ghost_idx] = np.where(nat >= 0, directory[np.clip(nat, 0, None)], -1)
    return _LocalLayout(l2g, offset, n_owned, int(natural_dof_map.size), n_global, owned_mask, natural_dof_map)


def layout_from_compound(compound_cls, partition_info, comm):
    """compound/mpi.py:288-494 (`_layout_from_compound`): natural DOF ids and owned mask for every field
    of a Compound on a partitioned mesh, then `_create_dof_layout`.  Nodal (full and incomplete), Local
    and Shared fields follow the reference's rules; the returned info dict maps field name ->
    FieldGlobalInfo (global shape, base offset, strides, node subset), the data behind `Compound._g`."""
    from .compound import Local, Nodal, Shared

    comm = _as_comm(comm)
    l2g_nodes = np.asarray(partition_info.nodes_local_to_global)
    n_owned_nodes = int(partition_info.n_owned_nodes)
    natural = np.full(compound_cls.size, -1, dtype=np.int32)
    owned = np.zeros(compound_cls.size, dtype=bool)
    local_max = int(l2g_nodes.max()) if l2g_nodes.size else -1
    n_nodes_global = max(comm.allgather(local_max)) + 1
    cursor, done, info = 0, {}, {}

    def describe(f, n_items_global, base, subset):
        # a field is an affine window of its root block; per-item layout is the same locally and globally, so the
        # strides carry over and only the origin moves
        return FieldGlobalInfo((n_items_global, *f.shape[1:]), base + (f._base_offset - f._root_slice.start), tuple(f._strides), subset)

    for name, f in compound_cls.fields:
        sl = f._root_slice
        key = (sl.start, sl.stop)
        ft = f.field_type.get()
        root_shape = f._root_shape
        per_item = int(np.prod(root_shape[1:])) if len(root_shape) > 1 else 1
        size_local = sl.stop - sl.start
        if key in done:
            info[name] = describe(f, *done[key])
            continue
        g_subset = None
        if isinstance(ft, Local):
            sizes = comm.allgather(size_local)
            n_items_global = sum(sizes) // per_item
            natural[sl] = cursor + sum(sizes[: comm.rank]) + np.arange(size_local)
            owned[sl] = True
        elif isinstance(ft, Nodal):
            if ft.node_ids is not None:
                ids_local = np.asarray(ft.node_ids)
                sub_global = l2g_nodes[ids_local]
                g_subset = np.unique(np.concatenate(comm.allgather(sub_global)))
                n_items_global = len(g_subset)
                is_owned = np.isin(sub_global, l2g_nodes[:n_owned_nodes])
                owned[sl] = np.repeat(is_owned, per_item)
                pos = np.searchsorted(g_subset, sub_global)
                natural[sl] = cursor + np.repeat(pos, per_item) * per_item + np.tile(np.arange(per_item), len(sub_global))
            else:
                n_items_global = n_nodes_global
                natural[sl] = cursor + (l2g_nodes[:, None] * per_item + np.arange(per_item)).ravel()
                owned[sl.start : sl.start + n_owned_nodes * per_item] = True
        elif isinstance(ft, Shared):
            n_items_global = size_local // per_item
            natural[sl] = cursor + np.arange(size_local)
            if comm.rank == 0:
                owned[sl] = True
        else:
            raise TypeError(f"Unsupported field type: {type(ft)}")
        done[key] = (n_items_global, cursor, g_subset)
        info[name] = describe(f, *done[key])
        cursor += n_items_global * per_item
    return _create_dof_layout(natural, owned, cursor, comm), info


# ---------------------------------------------------------------------------------------------------
# pack / exchange / unpack
# ---------------------------------------------------------------------------------------------------


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pack(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    if idx.numel() == 0:
        return src.new_empty(0)
    if src.is_cuda:
        out = torch.empty(idx.numel(), dtype=src.dtype, device=src.device)
        _lib.check(_lib.lib().tatva_halo_pack(src.data_ptr(), idx.data_ptr(), idx.numel(), out.data_ptr(), _stream()), "tatva_halo_pack")
        return out
    _lib.host_tables_only("ExchangePlan pack")
    return src[idx]


def _unpack(dst: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor, add: bool) -> None:
    if idx.numel() == 0:
        return
    if dst.is_cuda:
        fn = _lib.lib().tatva_halo_unpack_add if add else _lib.lib().tatva_halo_unpack_set
        _lib.check(fn(vals.data_ptr(), idx.data_ptr(), idx.numel(), dst.data_ptr(), _stream()), "tatva_halo_unpack")
    else:
        _lib.host_tables_only("ExchangePlan unpack")
        if add:
            dst.index_add_(0, idx, vals)
        else:
            dst[idx] = vals


class _Router:
    """One direction of a plan: gather `send` entries per neighbour, exchange, scatter to `recv`."""

    def __init__(self, comm, neighbors, send_lists, recv_lists, self_send, self_recv):
        self.comm = comm
        self.send_splits = [0] * comm.size
        self.recv_splits = [0] * comm.size
        for nb, s, r in zip(neighbors, send_lists, recv_lists):
            self.send_splits[nb] = len(s)
            self.recv_splits[nb] = len(r)
        cat = lambda parts: np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, dtype=np.int64)  # noqa: E731
        self._send_np, self._recv_np = cat(list(send_lists)), cat(list(recv_lists))
        self._self_send_np, self._self_recv_np = np.asarray(self_send, dtype=np.int64), np.asarray(self_recv, dtype=np.int64)
        self._dev = {}

    def tables(self, device):
        if device not in self._dev:
            t = lambda a: torch.as_tensor(a, device=device)  # noqa: E731
            self._dev[device] = (t(self._send_np), t(self._recv_np), t(self._self_send_np), t(self._self_recv_np))
        return self._dev[device]

    def run(self, src: torch.Tensor, dst: torch.Tensor, add: bool):
        send_idx, recv_idx, self_send, self_recv = self.tables(src.device)
        # self part: dst[self_recv] (+)= src[self_send]
        _unpack(dst, self_recv, _pack(src, self_send), add)
        if self.comm.size > 1:
            send_buf = _pack(src, send_idx)
            recv_buf = torch.empty(int(sum(self.recv_splits)), dtype=src.dtype, device=src.device)
            dist.all_to_all_single(recv_buf, send_buf, self.recv_splits, self.send_splits, group=self.comm.group)
            _unpack(dst, recv_idx, recv_buf, add)
        return dst


class ExchangePlan:
    """Point-to-point ghost exchange plan — mpi.py:133-516."""

    def __init__(self, layout: _LocalLayout, local_sparsity_pattern: csr_matrix | None = None, *, comm=None):
        self._comm = _as_comm(comm)
        self._rank, self._size = self._comm.rank, self._comm.size
        self.layout = layout
        self._rstart = layout.offset
        self._rend = layout.offset + layout.n_owned
        self._global_size = layout.n_global
        self._precompute_routing_tables()
        self.hessian_layout = None
        if local_sparsity_pattern is not None:
            self.hessian_layout = self._precompute_nnz_routing_tables(local_sparsity_pattern)

    # mpi.py:168-234.  The discovery / negotiation / resolution handshake collapses into one object
    # all-gather of (range, {neighbour: global ids of the local DOFs that neighbour owns}).
    def _precompute_routing_tables(self):
        l2g = np.asarray(self.layout.local_to_global)
        all_ranges = self._comm.allgather((self._rstart, self._rend))
        to_send = {}
        for nbr, (rs, re) in enumerate(all_ranges):
            if nbr == self._rank:
                continue
            idx = np.where((l2g >= rs) & (l2g < re))[0].astype(np.int32)
            if idx.size:
                to_send[nbr] = idx
        wanted = self._comm.allgather({nbr: l2g[idx] for nbr, idx in to_send.items()})
        self._neighbor_dof_data = []
        for nbr in range(self._size):
            if nbr == self._rank:
                continue
            send_idx = to_send.get(nbr, np.zeros(0, dtype=np.int32))
            recv_global = wanted[nbr].get(self._rank, np.zeros(0, dtype=np.int32))
            if send_idx.size == 0 and len(recv_global) == 0:
                continue
            self._neighbor_dof_data.append(
                _NeighborRoute(nbr, send_idx, (np.asarray(recv_global) - self._rstart).astype(np.int32), int(send_idx.size), int(len(recv_global)))
            )
        owned_mask = np.asarray(self.layout.owned_mask)
        self._send_dof = np.where(owned_mask)[0].astype(np.int32)
        self._recv_dof = (l2g[self._send_dof] - self._rstart).astype(np.int32)
        nbrs = [d.rank for d in self._neighbor_dof_data]
        # forward (owned -> local): roles swap, send <- recv_local_idx, recv <- local_send_idx (mpi.py:381-389)
        self._fwd = _Router(self._comm, nbrs, [d.recv_local_idx for d in self._neighbor_dof_data], [d.local_send_idx for d in self._neighbor_dof_data], self._recv_dof, self._send_dof)
        self._rev = _Router(self._comm, nbrs, [d.local_send_idx for d in self._neighbor_dof_data], [d.recv_local_idx for d in self._neighbor_dof_data], self._send_dof, self._recv_dof)

    # mpi.py:236-336
    def _precompute_nnz_routing_tables(self, local_pattern) -> _HessianLayout:
        indptr, indices = np.asarray(local_pattern.indptr), np.asarray(local_pattern.indices)
        l2g = np.asarray(self.layout.local_to_global)
        l_row = np.repeat(np.arange(len(indptr) - 1, dtype=np.int32), np.diff(indptr))
        g_row, g_col = l2g[l_row], l2g[indices]
        valid = (g_row >= 0) & (g_col >= 0)
        l_nnz = np.where(valid)[0]
        g_row, g_col = g_row[valid], g_col[valid]
        all_ranges = self._comm.allgather((self._rstart, self._rend))
        send_to, coords_to = [], {}
        for nbr, (rs, re) in enumerate(all_ranges):
            m = (g_row >= rs) & (g_row < re)
            send_to.append(l_nnz[m].astype(np.int32))
            if nbr != self._rank and m.any():
                coords_to[nbr] = (g_row[m], g_col[m])
        received = self._comm.allgather(coords_to)
        mine = (g_row >= self._rstart) & (g_row < self._rend)
        rows, cols = [g_row[mine]], [g_col[mine]]
        nbrs = sorted(d for d in range(self._size) if d != self._rank and (len(send_to[d]) > 0 or self._rank in received[d]))
        recv_sizes = {}
        for nbr in nbrs:
            r, c = received[nbr].get(self._rank, (np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)))
            rows.append(np.asarray(r))
            cols.append(np.asarray(c))
            recv_sizes[nbr] = len(r)
        pairs = np.column_stack((np.concatenate(rows), np.concatenate(cols)))
        uniq, inverse = np.unique(pairs, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        counts = np.bincount(uniq[:, 0] - self._rstart, minlength=self.local_size)
        owned_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        n_self = len(rows[0])
        routes, cur = [], n_self
        for nbr in nbrs:
            n = recv_sizes[nbr]
            routes.append(_NeighborRoute(nbr, send_to[nbr], inverse[cur : cur + n].astype(np.int32), int(len(send_to[nbr])), int(n)))
            cur += n
        hl = _HessianLayout(int(uniq.shape[0]), owned_ptr, uniq[:, 1].astype(np.int32), send_to[self._rank], inverse[:n_self].astype(np.int32), routes)
        self._rev_nnz = _Router(self._comm, [d.rank for d in routes], [d.local_send_idx for d in routes], [d.recv_local_idx for d in routes], hl.local_send_idx, hl.recv_local_idx)
        return hl

    global_size = property(lambda self: self._global_size)
    rstart = property(lambda self: self._rstart)
    rend = property(lambda self: self._rend)
    local_size = property(lambda self: self._rend - self._rstart)

    @property
    def owned_nnz(self):
        if self.hessian_layout is None:
            raise ValueError("Hessian layout not initialized.")
        return self.hessian_layout.owned_nnz

    @property
    def owned_csr(self):
        if self.hessian_layout is None:
            raise ValueError("Hessian layout not initialized.")
        return self.hessian_layout.owned_ptr, self.hessian_layout.owned_indices

    def make_scatter_fwd_set(self) -> Callable:
        """x_owned -> u_local: ghost values fetched from their owners (mpi.py:372-409)."""
        n_local = len(self.layout.local_to_global)

        def fn(x_owned):
            x = torch.as_tensor(x_owned)
            return self._fwd.run(x.contiguous(), torch.zeros(n_local, dtype=x.dtype, device=x.device), add=False)

        return fn

    def make_scatter_rev_add(self, local_fn: Callable, is_hessian: bool = False) -> Callable:
        """args -> owned data: run `local_fn`, then add every rank's ghost contributions into the owner's
        rows (mpi.py:422-516).  With `is_hessian`, `local_fn` returns a ColoredMatrix and its `data` is routed
        by nonzero."""
        if is_hessian:
            if self.hessian_layout is None:
                raise ValueError("Hessian layout not initialized.")
            n_out, router = self.hessian_layout.owned_nnz, self._rev_nnz
        else:
            n_out, router = self.local_size, self._rev

        def fn(*args, **kwargs):
            result = local_fn(*args, **kwargs)
            data = torch.as_tensor(result.data if is_hessian else result).reshape(-1).contiguous()
            owned = router.run(data, torch.zeros(n_out, dtype=data.dtype, device=data.device), add=True)
            return replace(result, data=owned) if is_hessian else owned

        return fn


class AllreducePlan:
    """Replicated-vector plan — mpi.py:519-711: every rank holds the full vector, computes its element
    subset's contribution, and an all-reduce sums them; each rank keeps its block of rows."""

    def __init__(self, global_size: int, global_sparsity_pattern: csr_matrix | None = None, *, comm=None):
        self._comm = _as_comm(comm)
        self._rank, self._size = self._comm.rank, self._comm.size
        if global_sparsity_pattern is not None:
            assert global_sparsity_pattern.indptr.size - 1 == global_size, "global_sparsity_pattern shape does not match provided global_size"
        self._global_size = global_size
        self._rstart, self._rend = _dof_range(global_size, self._size, self._rank)
        if global_sparsity_pattern is not None:
            indptr, indices = np.asarray(global_sparsity_pattern.indptr), np.asarray(global_sparsity_pattern.indices)
            a, b = int(indptr[self._rstart]), int(indptr[self._rend])
            self._owned_nnz, self._owned_nnz_start = b - a, a
            self._owned_ptr = (indptr[self._rstart : self._rend + 1] - a).astype(np.int32)
            self._owned_indices = indices[a:b].astype(np.int32)
            self._plan_hessian = True
        else:
            self._owned_nnz, self._owned_nnz_start = 0, 0
            self._owned_ptr, self._owned_indices = np.array([0], dtype=np.int32), np.array([], dtype=np.int32)
            self._plan_hessian = False

    global_size = property(lambda self: self._global_size)
    rstart = property(lambda self: self._rstart)
    rend = property(lambda self: self._rend)
    local_size = property(lambda self: self._rend - self._rstart)
    owned_nnz = property(lambda self: self._owned_nnz)
    owned_csr = property(lambda self: (self._owned_ptr, self._owned_indices))

    def make_allgather(self) -> Callable:
        """x_owned -> full replicated vector (mpi.py:609-635; Allgatherv -> all_gather of ragged blocks)."""
        counts = [b - a for a, b in (_dof_range(self._global_size, self._size, r) for r in range(self._size))]

        def fn(x_owned):
            x = torch.as_tensor(x_owned).contiguous()
            if self._size == 1:
                return x.clone()
            parts = [torch.empty(c, dtype=x.dtype, device=x.device) for c in counts]
            dist.all_gather(parts, x, group=self._comm.group)
            return torch.cat(parts)

        return fn

    def make_allreduce_owned(self, local_fn: Callable, is_hessian: bool = False) -> Callable:
        """mpi.py:646-711."""
        if is_hessian and not self._plan_hessian:
            raise ValueError("AllreducePlan not initialized with Hessian sparsity pattern.")

        def reduce(t):
            t = torch.as_tensor(t).contiguous().clone()
            if self._size > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._comm.group)
            return t

        if not is_hessian:
            return lambda *a, **k: reduce(local_fn(*a, **k))[self._rstart : self._rend]

        def fn(*a, **k):
            from .sparse import ColoredMatrix

            result = local_fn(*a, **k)
            if not isinstance(result, ColoredMatrix):
                raise TypeError("local_fn must return a ColoredMatrix when is_hessian=True.")
            data = reduce(result.data)[self._owned_nnz_start : self._owned_nnz_start + self._owned_nnz]
            return replace(result, data=data, indices=self._owned_indices, indptr=self._owned_ptr, shape=(self._rend - self._rstart, self._global_size))

        return fn


def _dof_range(n: int, size: int, rank: int) -> tuple[int, int]:
    """mpi.py:714-726: block distribution, the first n % size ranks own one extra DOF."""
    base, rem = divmod(n, size)
    if rank < rem:
        start = rank * (base + 1)
        return start, start + base + 1
    start = rank * base + rem
    return start, start + base
