"""Element-partitioned operator: one mesh partition per GPU, halo exchange overlapped with the
interior quadrature loop.

This is the B200 form of the reference's distributed call stack (tatva/mpi.py:372-409, :479-516):

    x_owned --scatter_fwd--> u_local --local_fn (HVP over the rank's mesh)--> y_local --scatter_rev_add--> y_owned

with the reference's ownership and numbering rules (`extract_local_mesh`, mesh.py:234-291; owned nodes
first, so the owned block of a local vector IS the owned vector — no copy).  Per application:

    compute stream:  zero y_local ........ interior elements (touch no ghost node) ...................|
    comm stream:     pack owned v -> all_to_all (NVLink) -> ghost v | boundary elements | pack ghost y |
                     -> all_to_all -> atomic add into owned y ........................................|

Both element kernels and the unpack-add accumulate into y_local with FP64 atomics, so they may run
concurrently; the two streams join at the end.  Only the halo (a few hundred kB per face at 128^3 per
GPU) crosses NVLink; there is no other data-path collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .compound import Compound, FieldSize, field
from .mesh import Mesh, PartitionInfo
from .mpi import ExchangePlan, _as_comm
from .operator import Operator


def structured_hex_block(n, grid, rank, lengths=None, jitter=0.1, seed=0):
    """Local mesh of block `rank` of a (gx*n) x (gy*n) x (gz*n) Hex8 box split into gx x gy x gz blocks
    of n^3 elements (partition id = px + gx (py + gy pz)), WITHOUT materialising the global mesh.  `n` may be a
    triple (nx, ny, nz) of cells per block (strong scaling: a fixed global box cut into non-cubic blocks).

    Returns (Mesh, PartitionInfo) identical to `extract_local_mesh(global_mesh, block_partition, rank)`:
    nodes owned by the smallest touching partition id, owned nodes first, each group ascending in global
    node id (mesh.py:258-273); elements in global (x-fastest) order.  Coordinates carry the same
    deterministic jitter a global generator would apply (hash of the global node id)."""
    gx, gy, gz = grid
    nx, ny, nz = (n, n, n) if np.isscalar(n) else (int(a) for a in n)
    px, py, pz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
    NX, NY, NZ = gx * nx, gy * ny, gz * nz
    i0, j0, k0 = px * nx, py * ny, pz * nz
    ii, jj, kk = np.arange(i0, i0 + nx + 1), np.arange(j0, j0 + ny + 1), np.arange(k0, k0 + nz + 1)
    K, J, I = np.meshgrid(kk, jj, ii, indexing="ij")
    gid = (I + (NX + 1) * (J + (NY + 1) * K)).ravel()  # ascending along the local lexicographic order
    ghost = ((I == i0) & (px > 0)) | ((J == j0) & (py > 0)) | ((K == k0) & (pz > 0))
    ghost = ghost.ravel()
    order = np.concatenate([np.where(~ghost)[0], np.where(ghost)[0]])  # owned first, each ascending in gid
    new_id = np.empty(order.size, dtype=np.int32)
    new_id[order] = np.arange(order.size, dtype=np.int32)
    m = max(NX, NY, NZ)
    lengths = lengths or (NX / m, NY / m, NZ / m)
    g = gid[order]
    gi, gj, gk = g % (NX + 1), (g // (NX + 1)) % (NY + 1), g // ((NX + 1) * (NY + 1))
    coords = np.stack([gi * (lengths[0] / NX), gj * (lengths[1] / NY), gk * (lengths[2] / NZ)], axis=-1).astype(np.float64)
    if jitter:
        coords = coords + jitter * (lengths[0] / NX) * _hash_uniform(g, seed)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (i * sx + j * sy + k * sz).ravel()
    el = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + sz, n0 + sz + sx, n0 + sz + sx + sy, n0 + sz + sy], -1)
    return Mesh(coords=coords, elements=new_id[el]), PartitionInfo(nodes_local_to_global=g.astype(np.int64), n_owned_nodes=int((~ghost).sum()))


def structured_tet_block(n, grid, rank, lengths=None, jitter=0.1, seed=0):
    """Same block as `structured_hex_block`, each cell cut into the 6 tetrahedra of the reference's box helper
    (tests/test_sparse_tracer.py:29-70).  Every cell corner belongs to one of its tets, so node ownership and
    local numbering are those of the hex block."""
    mesh, info = structured_hex_block(n, grid, rank, lengths=lengths, jitter=jitter, seed=seed)
    h = np.asarray(mesh.elements)
    n0, n1, n2, n3, n4, n5, n6, n7 = h[:, 0], h[:, 1], h[:, 3], h[:, 2], h[:, 4], h[:, 5], h[:, 7], h[:, 6]
    corner_sets = [(n0, n1, n3, n7), (n0, n1, n7, n5), (n0, n5, n7, n4), (n0, n3, n2, n7), (n0, n2, n6, n7), (n0, n6, n4, n7)]
    tets = np.stack([np.stack(t, -1) for t in corner_sets], axis=1).reshape(-1, 4)
    return Mesh(coords=mesh.coords, elements=tets.astype(np.int32)), info


def _hash_uniform(gid, seed):
    """Deterministic U(-1,1)^3 per global node id (same value on every rank that holds the node)."""
    x = gid.astype(np.uint64)[:, None] * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :] + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


_ABI_COMMS = {}


def _abi_nccl_comm(comm, device):
    """An ncclComm_t of our own over the ranks of `comm` (torch does not hand out its communicators): rank 0 draws the
    id through the C ABI, torch.distributed ships it, every rank joins.  One per (group, device), kept for the process."""
    key = (id(comm.group), str(device))
    if key not in _ABI_COMMS:
        L = _lib.lib()
        buf = (C.c_char * 128)()
        if comm.rank == 0:
            _lib.check(L.tatva_halo_comm_unique_id(buf), "tatva_halo_comm_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(comm.group, 0) if comm.group is not None else 0, group=comm.group)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.tatva_halo_comm_create(C.byref(handle), box[0], comm.size, comm.rank), "tatva_halo_comm_create")
        _ABI_COMMS[key] = handle
    return _ABI_COMMS[key]


class PartitionedOperator:
    """Operator over one partition + its ExchangePlan + the overlapped distributed HVP / residual."""

    def __init__(self, local_mesh: Mesh, partition_info: PartitionInfo, element, material, comm=None, device=None, overlap=True, halo="nccl", use_graph=True):
        self.comm = _as_comm(comm)
        self.material = material
        self.info = partition_info
        conn = np.asarray(local_mesh.elements)
        n_owned_nodes = int(partition_info.n_owned_nodes)
        # boundary elements (touch a ghost node) first, interior after; stable => locality kept
        touches_ghost = (conn >= n_owned_nodes).any(axis=1)
        order = np.concatenate([np.where(touches_ghost)[0], np.where(~touches_ghost)[0]])
        self.n_boundary = int(touches_ghost.sum())
        self.mesh = Mesh(coords=local_mesh.coords, elements=conn[order])
        self.op = Operator(self.mesh, element, device=device)
        self.device = self.op.device
        self.dpn = material.dofs_per_node(self.op.dim)
        self.n_local = self.op.n_nodes * self.dpn
        self.n_owned = n_owned_nodes * self.dpn

        mesh_for_layout, dpn = self.mesh, self.dpn

        class _State(Compound, mesh=mesh_for_layout, partition_info=partition_info, comm=self.comm):
            s = field(shape=(FieldSize.AUTO, dpn))

        self.plan = ExchangePlan(_State.get_layout(), comm=self.comm)
        assert self.plan.local_size == self.n_owned
        self.n_global = self.plan.global_size
        self.overlap = bool(overlap) and self.comm.size > 1
        # high priority: the small exchange / boundary kernels must not queue behind the interior grid
        self._comm_stream = torch.cuda.Stream(device=self.device, priority=-1) if self.overlap else None
        self._prm = _lib.params_array(material.params())
        self.op._ensure_geometry(material.material_id)  # set-up work (allocates): not inside a captured application
        self._L = _lib.lib()
        # halo = "peer": vectors in symmetric (peer-mapped) memory, ghosts pulled / pushed by our own kernels over
        # NVLink with a device-side barrier on each side;
        # halo = "nccl": the same exchange entirely behind the C ABI (tatva_halo_exchange: pack kernel -> grouped
        # ncclSend / ncclRecv on an ncclComm_t of our own -> unpack kernel); "nccl_torch": torch's all_to_all_single.
        if halo not in ("nccl", "nccl_torch", "peer"):
            raise ValueError(f"unknown halo transport {halo!r}")
        self.halo = halo if self.comm.size > 1 else "nccl_torch"
        self._nccl = _abi_nccl_comm(self.comm, self.device) if self.halo == "nccl" else None
        self._xbuf = {}
        self._sym = {}
        if self.halo == "peer":
            self._setup_peer_tables()
        # One application with the peer halo is ~7 launches on two streams (memset, interior kernel | barrier, pull,
        # boundary kernel, push, barrier).  At config-5 size the kernels take 0.07 ms and issuing them from Python took
        # longer than running them (0.0985 ms at 8 GPUs in r01), so the whole application — both streams, the signal-pad
        # barriers included — is captured ONCE per (entry point, vectors) in a CUDA graph and replayed.  All ranks capture
        # at the same call (the warm-up application inside `_capture` is a collective like any other application).
        self.use_graph = bool(use_graph) and self.halo == "peer" and self.overlap
        self._graphs = {}

    # -- peer-memory halo ----------------------------------------------------------------------------------
    def _setup_peer_tables(self):
        l2g = np.asarray(self.plan.layout.local_to_global, dtype=np.int64)
        ranges = self.comm.allgather((self.plan.rstart, self.plan.rend))
        sizes = self.comm.allgather(self.n_local)
        self._sym_len = int(max(sizes))
        ghosts = l2g[self.n_owned :]
        starts = np.array([r[0] for r in ranges], dtype=np.int64)
        owner = (np.searchsorted(starts, ghosts, side="right") - 1).astype(np.int32)
        self._n_ghost = int(ghosts.size)
        self._ghost_owner = torch.as_tensor(owner, device=self.device)
        self._ghost_owner_idx = torch.as_tensor(ghosts - starts[owner], device=self.device)  # owned-first: owned block index == local index

    def new_symmetric_vector(self) -> torch.Tensor:
        """Local vector (n_local doubles, zero-filled) in peer-mapped memory; required for halo="peer"."""
        import torch.distributed._symmetric_memory as symm_mem

        with torch.cuda.device(self.device):
            buf = symm_mem.empty(self._sym_len, dtype=torch.float64, device=self.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group=self.comm.group if self.comm.group is not None else dist.group.WORLD)
        vec = buf[: self.n_local]
        self._sym[vec.data_ptr()] = (buf, hdl, torch.as_tensor(np.asarray(hdl.buffer_ptrs, dtype=np.uint64).view(np.int64), device=self.device))
        return vec

    def _peer(self, x):
        try:
            return self._sym[x.data_ptr()]
        except KeyError:
            raise ValueError('halo="peer" needs vectors from new_symmetric_vector()') from None

    def _peer_pull(self, x):
        _, hdl, ptrs = self._peer(x)
        hdl.barrier(channel=0)  # every owner's values are in place
        _lib.check(self._L.tatva_peer_pull(x.data_ptr(), self.n_owned, self._n_ghost, ptrs.data_ptr(), self._ghost_owner.data_ptr(), self._ghost_owner_idx.data_ptr(), torch.cuda.current_stream().cuda_stream), "tatva_peer_pull")

    def _peer_push(self, y):
        _, hdl, ptrs = self._peer(y)
        _lib.check(self._L.tatva_peer_push_add(y.data_ptr(), self.n_owned, self._n_ghost, ptrs.data_ptr(), self._ghost_owner.data_ptr(), self._ghost_owner_idx.data_ptr(), torch.cuda.current_stream().cuda_stream), "tatva_peer_push_add")
        hdl.barrier(channel=1)  # every contribution has landed in its owner's rows

    # -- building blocks ------------------------------------------------------------------------------
    def fill_ghosts(self, x_local: torch.Tensor) -> None:
        """In place: ghost entries of a local vector <- owners' values (scatter_fwd_set, mpi.py:372-409)."""
        if self.comm.size > 1:
            if self.halo == "peer":
                self._peer_pull(x_local)
            else:
                self._exchange(self.plan._fwd, x_local, x_local, add=False)

    def _elems(self, name, u, v, y, begin, count, zero):
        prm, n = self._prm
        args = (u.data_ptr(), v.data_ptr(), y.data_ptr()) if v is not None else (u.data_ptr(), y.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(getattr(self._L, name)(self.op._plan_fused, self.material.material_id, prm, n, *args, begin, count, zero, torch.cuda.current_stream().cuda_stream), name)

    def _apply(self, name, u_local, v_local, y_local):
        if self.use_graph and not torch.cuda.is_current_stream_capturing():
            key = (name, u_local.data_ptr(), v_local.data_ptr() if v_local is not None else 0, y_local.data_ptr())
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = self._capture(name, u_local, v_local, y_local)
            g.replay()
            return y_local
        return self._apply_eager(name, u_local, v_local, y_local)

    def _capture(self, name, u_local, v_local, y_local):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):  # warm-up outside the capture (lazy initialisation, shared-memory opt-ins)
            self._apply_eager(name, u_local, v_local, y_local)
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._apply_eager(name, u_local, v_local, y_local)
        return g

    def _apply_eager(self, name, u_local, v_local, y_local):
        E = self.op.n_elements
        # the boundary launch is rounded up to whole 128-element tiles (a few interior elements ride along): both element
        # ranges then start on a tile boundary and keep the plan's node schedule (tatva_hvp_elems)
        nb = min(E, -(-self.n_boundary // 128) * 128)
        if self.comm.size == 1:
            self._elems(name, u_local, v_local, y_local, 0, E, 1)
            return y_local
        x = v_local if v_local is not None else u_local  # the vector whose ghosts must be refreshed
        peer = self.halo == "peer"
        if not self.overlap:
            if peer:
                y_local.zero_()
                self._peer_pull(x)  # its barrier also orders every rank's zeroing before any push
                self._elems(name, u_local, v_local, y_local, 0, E, 0)
                self._peer_push(y_local)
            else:
                self._exchange(self.plan._fwd, x, x, add=False)
                self._elems(name, u_local, v_local, y_local, 0, E, 1)
                self._exchange(self.plan._rev, y_local, y_local, add=True)
            return y_local
        main, side = torch.cuda.current_stream(self.device), self._comm_stream
        # y is cleared by the releasing kernel of the C ABI and the interior launch goes out right behind it (zero_y = 2):
        # the Hex8 kernel then runs its gather and Gauss-point loop while y is still being cleared (programmatic dependent
        # launch), every other kernel simply follows in stream order
        with torch.cuda.device(self.device):
            _lib.check(self._L.tatva_zero_release(y_local.data_ptr(), y_local.numel(), main.cuda_stream), "tatva_zero_release")
        zeroed = torch.cuda.Event()
        zeroed.record(main)
        side.wait_stream(main)  # inputs are ready
        self._elems(name, u_local, v_local, y_local, nb, E - nb, 2 if name == "tatva_hvp_elems" else 0)  # interior, compute stream
        with torch.cuda.stream(side):
            if peer:
                side.wait_event(zeroed)  # the pull's barrier then orders every rank's zeroing before any push
                self._peer_pull(x)
                self._elems(name, u_local, v_local, y_local, 0, nb, 0)  # boundary elements
                self._peer_push(y_local)
            else:
                self._exchange(self.plan._fwd, x, x, add=False)
                side.wait_event(zeroed)
                self._elems(name, u_local, v_local, y_local, 0, nb, 0)  # boundary elements
                self._exchange(self.plan._rev, y_local, y_local, add=True)
        main.wait_stream(side)
        return y_local

    def _exchange(self, router, src, dst, add):
        """Neighbour part of a _Router.run.  On owned-first local vectors the self part is the identity
        (and would double the owned rows when adding in place), so it is skipped."""
        from .mpi import _pack, _unpack

        send_idx, recv_idx, _, _ = router.tables(src.device)
        if self._nccl is not None:
            key = id(router)
            if key not in self._xbuf:  # staging buffers and the HOST count arrays of this direction, once
                mk = lambda n: torch.empty(max(int(n), 1), dtype=torch.float64, device=src.device)  # noqa: E731
                cnt = lambda v: (C.c_int64 * len(v))(*[int(x) for x in v])  # noqa: E731
                self._xbuf[key] = (mk(sum(router.send_splits)), mk(sum(router.recv_splits)), cnt(router.send_splits), cnt(router.recv_splits))
            sbuf, rbuf, sc, rc = self._xbuf[key]
            with torch.cuda.device(src.device):
                _lib.check(self._L.tatva_halo_exchange(self._nccl, src.data_ptr(), send_idx.data_ptr(), sc, sbuf.data_ptr(), rbuf.data_ptr(), rc, recv_idx.data_ptr(), dst.data_ptr(), int(bool(add)),
                                                      torch.cuda.current_stream(src.device).cuda_stream), "tatva_halo_exchange")
            return
        send_buf = _pack(src, send_idx)
        recv_buf = torch.empty(int(sum(router.recv_splits)), dtype=src.dtype, device=src.device)
        dist.all_to_all_single(recv_buf, send_buf, router.recv_splits, router.send_splits, group=self.comm.group)
        _unpack(dst, recv_idx, recv_buf, add)

    # -- public -----------------------------------------------------------------------------------------
    def new_local_vector(self) -> torch.Tensor:
        return torch.zeros(self.n_local, dtype=torch.float64, device=self.device)

    def owned(self, x_local: torch.Tensor) -> torch.Tensor:
        return x_local[: self.n_owned]

    def hvp(self, u_local: torch.Tensor, v_local: torch.Tensor, y_local: torch.Tensor | None = None) -> torch.Tensor:
        """y_owned = (H(u) v)_owned.  `u_local` must already hold its ghost values (it changes once per
        Newton step: call fill_ghosts then); the ghosts of `v_local` are refreshed here.  Returns the local
        vector whose first n_owned entries are the assembled owned rows."""
        if y_local is None:
            y_local = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
        return self._apply("tatva_hvp_elems", u_local, v_local, y_local)

    def hessian_diagonal(self, u_local: torch.Tensor, d_local: torch.Tensor | None = None) -> torch.Tensor:
        """diag H(u) assembled on the owned DOFs (Jacobi preconditioner of the distributed CG): element kernel on the
        local mesh, then the reverse halo add.  `u_local` holds its ghost values."""
        peer = self.halo == "peer" and self.comm.size > 1
        if d_local is None:
            d_local = self.new_symmetric_vector() if peer else self.new_local_vector()
        self.op.hessian_diagonal(self.material, u_local, out=d_local)
        if self.comm.size > 1:
            if peer:
                self._peer(d_local)[1].barrier(channel=0)  # every rank has zeroed and filled its own rows
                self._peer_push(d_local)
            else:
                self._exchange(self.plan._rev, d_local, d_local, add=True)
        return d_local

    def residual(self, u_local: torch.Tensor, r_local: torch.Tensor | None = None) -> torch.Tensor:
        if r_local is None:
            r_local = torch.empty(self.n_local, dtype=torch.float64, device=self.device)
        return self._apply("tatva_residual_elems", u_local, None, r_local)
