"""Exchange / allreduce plans on 2 `gloo` ranks (CPU): known answers of the reference's
tests/test_exchange_plan.py and tests/test_allreduce_plan.py, plus a partitioned-mesh residual whose
assembled owned rows must equal the single-process oracle."""
import os
import socket
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.sparse import csr_matrix


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        globals()[fn_name](rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None))
    except Exception:  # noqa: BLE001
        q.put((rank, traceback.format_exc()))


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    errs = [f"rank {r}:\n{e}" for r, e in results if e]
    assert not errs, "\n".join(errs)


def _mock_state(rank, fields):
    from tatva_b200.compound import Compound
    from tatva_b200.mesh import Mesh, PartitionInfo

    l2g = np.array([0, 1] if rank == 0 else [1, 0], dtype=np.int32)
    info = PartitionInfo(nodes_local_to_global=l2g, n_owned_nodes=1)
    mesh = Mesh(coords=np.zeros((2, 1)), elements=np.zeros((0, 2), dtype=np.int32))
    ns = dict(fields)
    return type("MyState", (Compound,), ns, mesh=mesh, partition_info=info, comm=dist.group.WORLD)


# ---- bodies (run inside the spawned ranks) -----------------------------------------------------------


def _body_layout(rank, world):
    """reference tests/test_exchange_plan.py:27-94."""
    from tatva_b200.compound import FieldType, field
    from tatva_b200.mpi import ExchangePlan

    S = _mock_state(rank, {"u": field(shape=(2, 2), field_type=FieldType.NODAL), "s": field(shape=(1,), field_type=FieldType.SHARED), "v": field(shape=(1,), field_type=FieldType.LOCAL)})
    plan = ExchangePlan(S.get_layout(), comm=dist.group.WORLD)
    assert plan.global_size == 7
    if rank == 0:
        assert (plan.local_size, plan.rstart, plan.rend) == (4, 0, 4)
        np.testing.assert_array_equal(plan.layout.local_to_global, [0, 1, 4, 5, 2, 3])
        np.testing.assert_array_equal(plan.layout.owned_mask, [True, True, False, False, True, True])
    else:
        assert (plan.local_size, plan.rstart, plan.rend) == (3, 4, 7)
        np.testing.assert_array_equal(plan.layout.local_to_global, [4, 5, 0, 1, 2, 6])
        np.testing.assert_array_equal(plan.layout.owned_mask, [True, True, False, False, False, True])


def _body_communication(rank, world):
    """reference tests/test_exchange_plan.py:97-166."""
    from tatva_b200.compound import FieldType, field
    from tatva_b200.mpi import ExchangePlan

    S = _mock_state(rank, {"u": field(shape=(2, 1), field_type=FieldType.NODAL)})
    plan = ExchangePlan(S.get_layout(), comm=dist.group.WORLD)
    fwd = plan.make_scatter_fwd_set()
    u_local = fwd(torch.tensor([10.0 if rank == 0 else 20.0], dtype=torch.float64))
    np.testing.assert_allclose(u_local, [10.0, 20.0] if rank == 0 else [20.0, 10.0])
    rev = plan.make_scatter_rev_add(lambda u: u * 2.0)
    np.testing.assert_allclose(rev(u_local), [40.0] if rank == 0 else [80.0])


def _body_incomplete_nodal(rank, world):
    """reference tests/test_exchange_plan.py:169-245."""
    from tatva_b200.compound import Compound, FieldSize, FieldType, Nodal, field
    from tatva_b200.mesh import Mesh, PartitionInfo
    from tatva_b200.mpi import ExchangePlan

    if rank == 0:
        info, subset = PartitionInfo(np.array([0, 1], dtype=np.int32), 1), np.array([0], dtype=np.int32)
    else:
        info, subset = PartitionInfo(np.array([1, 2], dtype=np.int32), 2), np.array([1], dtype=np.int32)
    mesh = Mesh(coords=np.zeros((2, 1)), elements=np.zeros((0, 2), dtype=np.int32))

    class MyState(Compound, mesh=mesh, partition_info=info, comm=dist.group.WORLD):
        u = field(shape=(FieldSize.AUTO, 1), field_type=FieldType.NODAL)
        l = field(shape=(FieldSize.AUTO, 1), field_type=Nodal(node_ids=subset))  # noqa: E741

    plan = ExchangePlan(MyState.get_layout(), comm=dist.group.WORLD)
    assert plan.global_size == 5
    if rank == 0:
        assert plan.local_size == 2
        np.testing.assert_array_equal(plan.layout.local_to_global, [0, 2, 1])
        np.testing.assert_array_equal(plan.layout.owned_mask, [True, False, True])
    else:
        assert plan.local_size == 3
        np.testing.assert_array_equal(plan.layout.local_to_global, [2, 3, 4])
        np.testing.assert_array_equal(plan.layout.owned_mask, [True, True, True])


def _body_hessian(rank, world):
    """reference tests/test_exchange_plan.py:248-355: assembled owned CSR data [410, 320] / [230, 140]."""
    from dataclasses import replace

    from tatva_b200.compound import FieldType, field
    from tatva_b200.mpi import ExchangePlan
    from tatva_b200.sparse import ColoredMatrix

    S = _mock_state(rank, {"u": field(shape=(2, 1), field_type=FieldType.NODAL)})
    indptr, indices = np.array([0, 2, 4], dtype=np.int32), np.array([0, 1, 0, 1], dtype=np.int32)
    pattern = csr_matrix(([0.0] * 4, indices, indptr), shape=(2, 2))
    cm = ColoredMatrix.from_csr(csr_matrix(([1.0] * 4, indices, indptr), shape=(2, 2)))
    cm = replace(cm, data=torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64) * (10.0 if rank == 0 else 100.0))
    plan = ExchangePlan(S.get_layout(), local_sparsity_pattern=pattern, comm=dist.group.WORLD)
    out = plan.make_scatter_rev_add(lambda u: cm, is_hessian=True)(torch.zeros(2))
    assert isinstance(out, ColoredMatrix) and plan.owned_nnz == 2
    np.testing.assert_array_equal(plan.owned_csr[0], [0, 2])
    np.testing.assert_array_equal(plan.owned_csr[1], [0, 1])
    np.testing.assert_allclose(out.data, [410.0, 320.0] if rank == 0 else [230.0, 140.0])


def _body_allreduce(rank, world):
    """reference tests/test_allreduce_plan.py:24-129."""
    from tatva_b200.mpi import AllreducePlan
    from tatva_b200.sparse import ColoredMatrix

    plan = AllreducePlan(global_size=6, comm=dist.group.WORLD)
    full = plan.make_allgather()(torch.ones(plan.local_size, dtype=torch.float64) * (1.0 if rank == 0 else 40.0))
    np.testing.assert_allclose(full, [1.0, 1.0, 1.0, 40.0, 40.0, 40.0])
    u = torch.tensor([0.0, 1.0, 2.0, 3.0, 0.0, 0.0] if rank == 0 else [0.0, 0.0, 0.25, 0.25, 4.0, 5.0], dtype=torch.float64)
    res = plan.make_allreduce_owned(lambda x: x)(u)
    np.testing.assert_allclose(res, [0.0, 1.0, 2.25] if rank == 0 else [3.25, 4.0, 5.0])

    indptr, indices = np.array([0, 1, 2, 3, 4, 5]), np.array([0, 1, 2, 3, 4])
    pattern = csr_matrix((np.ones_like(indices), indices, indptr), shape=(5, 5))
    vals = [1.0, 1.0, 1.0, 1.0, 0.0] if rank == 0 else [0.0, 0.0, 0.0, 1.0, 1.0]
    cm = ColoredMatrix.from_csr(csr_matrix((vals, indices, indptr), shape=(5, 5)))
    plan = AllreducePlan(global_size=5, global_sparsity_pattern=pattern, comm=dist.group.WORLD)
    out = plan.make_allreduce_owned(lambda x: cm, is_hessian=True)(torch.zeros(6))
    if rank == 0:
        np.testing.assert_allclose(out.data, [1.0, 1.0, 1.0])
        np.testing.assert_array_equal(out.indices, [0, 1, 2])
        np.testing.assert_array_equal(out.indptr, [0, 1, 2, 3])
    else:
        np.testing.assert_allclose(out.data, [2.0, 1.0])
        np.testing.assert_array_equal(out.indices, [3, 4])
        np.testing.assert_array_equal(out.indptr, [0, 1, 2])


def _body_partitioned_residual(rank, world):
    """Hex8 box split in two: owned rows of scatter_rev_add(local oracle residual)(scatter_fwd(x)) equal
    the single-process oracle residual in the plan's global numbering."""
    from oracle import tatva_oracle as orc
    from tatva_b200.compound import Compound, FieldSize, field
    from tatva_b200.mesh import Mesh, extract_local_mesh
    from tatva_b200.mpi import ExchangePlan

    c, el = orc.mesh_box_hex((4, 3, 2))
    c = c + 0.02 * np.random.default_rng(0).uniform(-1, 1, c.shape)
    u = 0.02 * np.random.default_rng(1).normal(size=c.shape)
    mat = orc.NeoHookean(500.0, 1000.0)
    part = (c[el].mean(axis=1)[:, 0] > c[:, 0].mean()).astype(np.int32)
    lm, info = extract_local_mesh(Mesh(coords=c, elements=el), part, rank)

    class S(Compound, mesh=lm, partition_info=info, comm=dist.group.WORLD):
        u = field(shape=(FieldSize.AUTO, 3))

    plan = ExchangePlan(S.get_layout(), comm=dist.group.WORLD)
    l2g_nodes = info.nodes_local_to_global
    x_owned = torch.as_tensor(u[l2g_nodes[: info.n_owned_nodes]].ravel())
    u_local = plan.make_scatter_fwd_set()(x_owned)
    np.testing.assert_array_equal(u_local.numpy().reshape(-1, 3), u[l2g_nodes])
    local_res = lambda ul: torch.as_tensor(orc.residual("hex8", mat, lm.coords, lm.elements, ul.numpy().reshape(-1, 3)).ravel())  # noqa: E731
    r_owned = plan.make_scatter_rev_add(local_res)(u_local)
    r_ref = orc.residual("hex8", mat, c, el, u)[l2g_nodes[: info.n_owned_nodes]].ravel()
    np.testing.assert_allclose(r_owned.numpy(), r_ref, rtol=1e-12, atol=1e-12 * np.abs(r_ref).max())
    assert plan.global_size == c.size


def _body_lifter_adapt_layout(rank, world):
    """reference tests/test_exchange_plan.py:358-417 (a fixed node drops out of the reduced layout) and :27-94 with
    `Lifter.adapt_layout` in the chain as the reference's tests use it."""
    from tatva_b200.compound import FieldType, field
    from tatva_b200.lifter import Fixed, Lifter
    from tatva_b200.mpi import ExchangePlan

    S = _mock_state(rank, {"u": field(shape=(2, 1), field_type=FieldType.NODAL)})
    l2g_nodes = np.array([0, 1] if rank == 0 else [1, 0])
    fixed = np.where(l2g_nodes == 0)[0].astype(np.int32)  # constrain global node 0
    lifter = Lifter(S.size, Fixed(fixed, 0.0))
    layout_reduced, lifter_aug = lifter.adapt_layout(S.get_layout(), dist.group.WORLD)
    plan = ExchangePlan(layout_reduced, comm=dist.group.WORLD)
    assert plan.global_size == 1
    np.testing.assert_array_equal(plan.layout.local_to_global, [0])
    assert (plan.local_size, plan.rstart, plan.rend) == ((0, 0, 0) if rank == 0 else (1, 0, 1))
    # unconstrained lifter: adapt_layout is the identity on the layout
    S2 = _mock_state(rank, {"u": field(shape=(2, 2), field_type=FieldType.NODAL), "s": field(shape=(1,), field_type=FieldType.SHARED), "v": field(shape=(1,), field_type=FieldType.LOCAL)})
    red, _ = Lifter(S2.size).adapt_layout(S2.get_layout(), dist.group.WORLD)
    np.testing.assert_array_equal(red.local_to_global, [0, 1, 4, 5, 2, 3] if rank == 0 else [4, 5, 0, 1, 2, 6])
    assert red.n_global == 7


def _body_periodic_mpi(rank, world):
    """reference tests/test_periodic_mpi.py:24-92: a ghost slave whose master lives on the other rank."""
    import scipy.sparse as sps

    from tatva_b200.compound import Compound, FieldType, field
    from tatva_b200.lifter import Lifter, PeriodicMPI
    from tatva_b200.mesh import Mesh, PartitionInfo

    l2g = np.array([0, 1, 2] if rank == 0 else [1, 2, 0], dtype=np.int32)
    info = PartitionInfo(nodes_local_to_global=l2g, n_owned_nodes=1 if rank == 0 else 2)
    mesh = Mesh(coords=np.zeros((3, 1)), elements=np.zeros((0, 2), dtype=np.int32))

    class MyState(Compound, mesh=mesh, partition_info=info, comm=dist.group.WORLD):
        u = field(shape=(3, 1), field_type=FieldType.NODAL)

    layout = MyState.get_layout()
    cond = PeriodicMPI(np.array([1]), np.array([2]), layout, comm=dist.group.WORLD)  # global node 1 follows node 2
    lifter = Lifter(MyState.size, cond)
    layout_aug, lifter_aug = lifter.adapt_layout(layout, dist.group.WORLD)
    assert layout_aug.n_global == 2  # nodes 0 and 2 remain
    full = lifter_aug.lift_from_zeros(np.array([5.0, 7.0]))
    g = {int(n): float(v) for n, v in zip(l2g, full)}
    assert g[1] == g[2]
    if rank == 0:
        sp_ = sps.lil_matrix((3, 3), dtype=np.int8)
        for i, j in ((0, 1), (1, 0), (0, 0), (1, 1)):
            sp_[i, j] = 1
        aug = lifter_aug.augment_sparsity(sp_.tocsr())
        assert aug[0, 2] != 0 and aug[2, 0] != 0


def _body_global_views(rank, world):
    """reference tests/test_compound_mpi.py:22-80 (global indices, full and incomplete nodal fields) and :83-167
    (stacked fields: global indices and the gathered global data of `state._g`)."""
    from tatva_b200.compound import Compound, FieldSize, Nodal, field, stack_fields
    from tatva_b200.mesh import Mesh, PartitionInfo

    comm = dist.group.WORLD
    if rank == 0:
        l2g, n_owned, l_nodes = np.array([0, 1, 2], dtype=np.int32), 2, np.array([1])
    else:
        l2g, n_owned, l_nodes = np.array([2, 3, 1], dtype=np.int32), 2, np.array([2, 1])
    mesh = Mesh(coords=np.zeros((len(l2g), 1)), elements=np.zeros((0, 2), dtype=np.int32))

    class A(Compound, mesh=mesh, partition_info=PartitionInfo(l2g, n_owned), comm=comm):
        u = field(shape=(FieldSize.AUTO, 2))
        l = field(shape=(FieldSize.AUTO, 1), field_type=Nodal(node_ids=l_nodes))  # noqa: E741

    np.testing.assert_array_equal(A._g.u[0], [0, 1])
    np.testing.assert_array_equal(A._g.u[3, 1], [7])
    np.testing.assert_array_equal(A._g.l[1], [8])  # l lives on global nodes {1, 3}, after the 8 DOFs of u
    np.testing.assert_array_equal(A._g.l[3], [9])
    np.testing.assert_array_equal(A._g.l[:], [8, 9])
    np.testing.assert_array_equal(A._g.l[1:2], [8])
    np.testing.assert_array_equal(A._g.l[2:4], [9])
    np.testing.assert_array_equal(A._g.l[:2], [8])
    np.testing.assert_array_equal(A._g.l[2:], [9])
    np.testing.assert_array_equal(A._g.l[np.array([3, 1])], [9, 8])
    with pytest.raises(NotImplementedError):
        A._g.l[::2]
    with pytest.raises(IndexError):
        A._g.l[0]
    with pytest.raises(AttributeError):
        A._g.nope

    # stacked fields: 3 global nodes; rank 0 owns 0 and 1, rank 1 owns 2
    if rank == 0:
        l2g, n_owned = np.array([0, 1, 2], dtype=np.int32), 2
    else:
        l2g, n_owned = np.array([2, 0, 1], dtype=np.int32), 1
    mesh = Mesh(coords=np.zeros((3, 1)), elements=np.zeros((0, 2), dtype=np.int32))

    @stack_fields("u", "v")
    class B(Compound, mesh=mesh, partition_info=PartitionInfo(l2g, n_owned), comm=comm):
        u = field(shape=(FieldSize.AUTO, 1), field_type=Nodal())
        v = field(shape=(FieldSize.AUTO, 2), field_type=Nodal())

    np.testing.assert_array_equal(B._g.u[:], [0, 3, 6])
    np.testing.assert_array_equal(B._g.v[:], [1, 2, 4, 5, 7, 8])
    assert B._g.u[1, 0] == 3 and B._g.v[2, 1] == 8
    for as_torch in (False, True):
        if rank == 0:
            st = B(u=np.array([[10.0], [20.0], [-1.0]]), v=np.array([[1.0, 2.0], [3.0, 4.0], [-1.0, -1.0]]))  # last row: ghost
        else:
            st = B(u=np.array([[30.0], [-1.0], [-1.0]]), v=np.array([[5.0, 6.0], [-1.0, -1.0], [-1.0, -1.0]]))
        if as_torch:
            st = B(torch.as_tensor(st.arr))
        view = st._g
        g_u, g_v = view.u, view.v  # one all-reduce, cached in the view
        np.testing.assert_allclose(np.asarray(g_u), [[10.0], [20.0], [30.0]])
        np.testing.assert_allclose(np.asarray(g_v), [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
        assert isinstance(g_u, torch.Tensor) == as_torch

    class NoLayout(Compound):
        w = field(shape=(2,))

    with pytest.raises(ValueError):
        NoLayout._g


def _reference_plan_body(rank, world, name):
    """The product's layout + ExchangePlan (vector and Hessian-nnz routing) on `world` gloo ranks against the tables
    the UNMODIFIED reference builds for the same partitioned mesh (fixtures `mpi_*`, tests/golden/_fakempi_golden.py)."""
    from scipy.sparse import csr_matrix as csr

    from tatva_b200 import sparse
    from tatva_b200.mesh import Mesh, extract_local_mesh
    from tatva_b200.mpi import ExchangePlan, _create_dof_layout

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    size, dpn = (int(x) for x in G[f"mpi_{name}_size_dpn"])
    assert size == world
    g = lambda k: G[f"mpi_{name}_r{rank}_{k}"]  # noqa: E731
    c, el, part = G[f"mpi_{name}_coords"], G[f"mpi_{name}_conn"], G[f"mpi_{name}_partition"]
    mesh, info = extract_local_mesh(Mesh(coords=c, elements=el), part, rank)
    l2g_nodes = np.asarray(info.nodes_local_to_global)
    natural = (l2g_nodes[:, None] * dpn + np.arange(dpn)).ravel().astype(np.int32)
    np.testing.assert_array_equal(natural, g("natural"))
    owned = np.zeros(natural.size, dtype=bool)
    owned[: int(info.n_owned_nodes) * dpn] = True
    layout = _create_dof_layout(natural, owned, c.shape[0] * dpn, dist.group.WORLD)
    off, n_owned, n_total, n_global = (int(x) for x in g("offset_nowned_ntotal_nglobal"))
    assert (layout.offset, layout.n_owned, layout.n_total, layout.n_global) == (off, n_owned, n_total, n_global)
    np.testing.assert_array_equal(layout.local_to_global, g("l2g"))
    pat = sparse.pattern_from_mesh(mesh, dpn)
    np.testing.assert_array_equal(pat.indptr, g("pat_indptr"))
    np.testing.assert_array_equal(pat.indices, g("pat_indices"))
    plan = ExchangePlan(layout, csr((np.ones(pat.nnz), pat.indices, pat.indptr), shape=pat.shape), comm=dist.group.WORLD)
    np.testing.assert_array_equal(plan._send_dof, g("self_send"))
    np.testing.assert_array_equal(plan._recv_dof, g("self_recv"))
    np.testing.assert_array_equal([d.rank for d in plan._neighbor_dof_data], g("nbr_ranks"))
    for d in plan._neighbor_dof_data:
        np.testing.assert_array_equal(d.local_send_idx, g(f"nbr{d.rank}_send"))
        np.testing.assert_array_equal(d.recv_local_idx, g(f"nbr{d.rank}_recv"))
    h = plan.hessian_layout
    assert h.owned_nnz == int(g("h_owned_nnz"))
    np.testing.assert_array_equal(h.owned_ptr, g("h_owned_ptr"))
    np.testing.assert_array_equal(h.owned_indices, g("h_owned_indices"))
    np.testing.assert_array_equal(h.local_send_idx, g("h_self_send"))
    np.testing.assert_array_equal(h.recv_local_idx, g("h_self_recv"))
    np.testing.assert_array_equal([d.rank for d in h.neighbor_data], g("h_nbr_ranks"))
    for d in h.neighbor_data:
        np.testing.assert_array_equal(d.local_send_idx, g(f"h_nbr{d.rank}_send"))
        np.testing.assert_array_equal(d.recv_local_idx, g(f"h_nbr{d.rank}_recv"))
    # and the plan moves data the way those tables say: ghosts receive the owners' values
    x_owned = torch.as_tensor(np.arange(off, off + n_owned, dtype=np.float64))  # value = global DOF id
    u_local = plan.make_scatter_fwd_set()(x_owned)
    np.testing.assert_array_equal(u_local.numpy(), np.asarray(layout.local_to_global, dtype=np.float64))
    # the reference's own exchange functions on the same inputs (mpi.py:372-516, run on the stand-in): forward ghost
    # fill, reverse add of a local vector, reverse add of Hessian nonzeros
    from dataclasses import replace as dc_replace

    np.testing.assert_array_equal(plan.make_scatter_fwd_set()(torch.as_tensor(g("x_owned"))).numpy(), g("fwd_u_local"))
    contrib = torch.as_tensor(g("rev_contrib"))
    np.testing.assert_allclose(plan.make_scatter_rev_add(lambda: contrib)().numpy(), g("rev_owned"), rtol=1e-14, atol=1e-14)
    cm = sparse.ColoredMatrix.from_csr(pat)
    Kd = plan.make_scatter_rev_add(lambda: dc_replace(cm, data=torch.as_tensor(g("rev_nnz_vals"))), is_hessian=True)()
    np.testing.assert_allclose(np.asarray(Kd.data), g("rev_owned_nnz"), rtol=1e-14, atol=1e-14)


def _body_reference_compound_layout(rank, world):
    """The product's `layout_from_compound` and `Compound._g` on 3 gloo ranks against the UNMODIFIED reference
    (compound/mpi.py:288-494, fixtures `cmp_*`): stacked nodal fields, a nodal field on a node subset, a shared and
    a local field."""
    from tatva_b200.compound import Compound, FieldSize, Local, Nodal, Shared, field
    from tatva_b200.mesh import Mesh, extract_local_mesh

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    g = lambda k: G[f"cmp_r{rank}_{k}"]  # noqa: E731
    mesh, info = extract_local_mesh(Mesh(coords=G["cmp_coords"], elements=G["cmp_conn"]), G["cmp_partition"], rank)
    sub = g("subset_local_nodes")

    class S(Compound, mesh=mesh, partition_info=info, comm=dist.group.WORLD):
        u = field(shape=(FieldSize.AUTO, 3))
        p = field(shape=(FieldSize.AUTO,))
        lam = field(shape=(FieldSize.AUTO, 2), field_type=Nodal(node_ids=sub))
        g = field(shape=(2,), field_type=Shared())
        w = field(shape=(rank + 1, 2), field_type=Local())

    assert S.size == int(g("size"))
    L = S.get_layout()
    np.testing.assert_array_equal(L.natural_l2g, g("natural"))
    np.testing.assert_array_equal(L.owned_mask, g("owned_mask"))
    np.testing.assert_array_equal(L.local_to_global, g("l2g"))
    off, n_owned, n_total, n_global = (int(x) for x in g("offset_nowned_ntotal_nglobal"))
    assert (L.offset, L.n_owned, L.n_total, L.n_global) == (off, n_owned, n_total, n_global)
    for name in ("u", "p", "lam", "g", "w"):
        fi = S._global_field_info[name]
        np.testing.assert_array_equal(np.array(fi.global_shape), g(f"{name}_gshape"))
        np.testing.assert_array_equal(np.array([fi.global_base_offset, *fi.global_strides]), g(f"{name}_goffset_strides"))
        view = getattr(S._g, name)
        if fi.global_subset is None:
            got = view[(slice(None),) * len(fi.global_shape)]
        else:
            np.testing.assert_array_equal(fi.global_subset, g(f"{name}_gsubset"))
            got = view[fi.global_subset]
        np.testing.assert_array_equal(got, g(f"{name}_gindices"))


def _body_reference_periodic_mpi(rank, world):
    """Lifter.adapt_layout with PeriodicMPI + Fixed on 3 gloo ranks against the UNMODIFIED reference (lifter/base.py:
    333-425, constraints.py:223-287; fixtures `pmpi_*`): masters on another rank become extra ghosts, the reduced
    layout renumbers the free DOFs, the resolved lifter lifts like the reference's."""
    from tatva_b200.lifter import Fixed, Lifter, PeriodicMPI
    from tatva_b200.mesh import Mesh, extract_local_mesh
    from tatva_b200.mpi import _create_dof_layout

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    g = lambda k: G[f"pmpi_r{rank}_{k}"]  # noqa: E731
    dpn = 3
    c = G["pmpi_coords"]
    mesh, info = extract_local_mesh(Mesh(coords=c, elements=G["pmpi_conn"]), G["pmpi_partition"], rank)
    l2g_nodes = np.asarray(info.nodes_local_to_global)
    natural = (l2g_nodes[:, None] * dpn + np.arange(dpn)).ravel().astype(np.int32)
    owned = np.zeros(natural.size, dtype=bool)
    owned[: int(info.n_owned_nodes) * dpn] = True
    layout = _create_dof_layout(natural, owned, c.shape[0] * dpn, dist.group.WORLD)
    lifter = Lifter(layout.n_total, Fixed(g("fixed_local_dofs"), 0.25), PeriodicMPI(G["pmpi_slaves"], G["pmpi_masters"], layout, comm=dist.group.WORLD))
    reduced, lifter2 = lifter.adapt_layout(layout, dist.group.WORLD)
    off, n_owned, n_total, n_global = (int(x) for x in g("red_offset_nowned_ntotal_nglobal"))
    assert (reduced.offset, reduced.n_owned, reduced.n_total, reduced.n_global) == (off, n_owned, n_total, n_global)
    np.testing.assert_array_equal(reduced.local_to_global, g("red_l2g"))
    np.testing.assert_array_equal(reduced.owned_mask, g("red_owned_mask"))
    np.testing.assert_array_equal(reduced.natural_l2g, g("red_natural"))
    size, extra = (int(x) for x in g("lifter_size_extra"))
    assert lifter2.size == size and getattr(lifter2, "_nb_extra_ghost_dofs", 0) == extra
    np.testing.assert_array_equal(lifter2.free_dofs, g("free_dofs"))
    per = lifter2.constraints[1]
    np.testing.assert_array_equal(per.dofs, g("periodic_dofs"))
    np.testing.assert_array_equal(per.master_dofs, g("periodic_masters"))
    np.testing.assert_array_equal(lifter2.lift_from_zeros(g("u_red")), g("lift_from_zeros"))


def _body_reference_allreduce_plan(rank, world):
    """AllreducePlan on 4 gloo ranks against the UNMODIFIED reference (mpi.py:519-711, fixtures `arp_*`)."""
    from dataclasses import replace as dc_replace

    from tatva_b200 import sparse
    from tatva_b200.mesh import Mesh
    from tatva_b200.mpi import AllreducePlan

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    g = lambda k: G[f"arp_r{rank}_{k}"]  # noqa: E731
    pat = sparse.pattern_from_mesh(Mesh(coords=G["arp_coords"], elements=G["arp_conn"]), 2)
    plan = AllreducePlan(pat.shape[0], pat, comm=dist.group.WORLD)
    rs, re, nnz = (int(x) for x in g("range_nnz"))
    assert (plan.rstart, plan.rend, plan.owned_nnz) == (rs, re, nnz)
    np.testing.assert_array_equal(plan.owned_csr[0], g("owned_ptr"))
    np.testing.assert_array_equal(plan.owned_csr[1], g("owned_indices"))
    np.testing.assert_array_equal(np.asarray(plan.make_allgather()(torch.as_tensor(g("x_owned")))), g("allgather"))
    vec = torch.as_tensor(g("vec"))
    np.testing.assert_allclose(np.asarray(plan.make_allreduce_owned(lambda: vec)()), g("owned_vec"), rtol=1e-14, atol=1e-14)
    cm = sparse.ColoredMatrix.from_csr(pat)
    K = plan.make_allreduce_owned(lambda: dc_replace(cm, data=torch.as_tensor(g("vals"))), is_hessian=True)()
    np.testing.assert_allclose(np.asarray(K.data), g("K_data"), rtol=1e-14, atol=1e-14)
    assert tuple(K.shape) == tuple(g("K_shape"))
    np.testing.assert_array_equal(K.indptr, g("K_indptr"))
    np.testing.assert_array_equal(K.indices, g("K_indices"))


def _body_reference_plan_hex3(rank, world):
    _reference_plan_body(rank, world, "hex3")


def _body_reference_plan_tri4(rank, world):
    _reference_plan_body(rank, world, "tri4")


# ---- pytest entry points -----------------------------------------------------------------------------


@pytest.mark.parametrize(
    "body",
    ["_body_layout", "_body_communication", "_body_incomplete_nodal", "_body_hessian", "_body_allreduce", "_body_partitioned_residual", "_body_lifter_adapt_layout", "_body_periodic_mpi", "_body_global_views"],
)
def test_two_rank_gloo(body):
    _run(body, world=2)


@pytest.mark.parametrize("body,world", [("_body_reference_plan_hex3", 3), ("_body_reference_plan_tri4", 4), ("_body_reference_compound_layout", 3), ("_body_reference_periodic_mpi", 3), ("_body_reference_allreduce_plan", 4)])
def test_plans_match_the_reference_on_more_ranks(body, world):
    _run(body, world=world)


def test_dof_range_and_single_process_plan():
    from tatva_b200.mpi import AllreducePlan, _dof_range

    assert [_dof_range(7, 3, r) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    plan = AllreducePlan(global_size=4)
    assert (plan.rstart, plan.rend, plan.local_size) == (0, 4, 4)
    np.testing.assert_allclose(plan.make_allreduce_owned(lambda x: x * 3)(torch.arange(4.0)), [0, 3, 6, 9])
