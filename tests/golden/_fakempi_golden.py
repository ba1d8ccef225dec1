"""MPI fixtures: the UNMODIFIED tatva.mpi._create_dof_layout and ExchangePlan (routing tables for vectors and for
Hessian nonzeros) built for every rank of a partitioned mesh, on the thread-based mpi4py stand-in (_fakempi)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_fakempi"))

from mpi4py import run_ranks  # noqa: E402  (the stand-in)


def mpi_fixtures(out):
    compound_fixtures(out)
    periodic_mpi_fixtures(out)
    allreduce_plan_fixtures(out)
    from tatva import Mesh, sparse
    from tatva.mesh import extract_local_mesh
    from tatva.mpi import ExchangePlan, _create_dof_layout

    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import tatva_oracle as orc

    cases = {}
    c, el = orc.mesh_box_hex((4, 3, 2))
    cen = c[el].mean(axis=1)
    cases["hex3"] = (c, el, ((cen[:, 0] > 0.26).astype(np.int32) + (cen[:, 0] > 0.74).astype(np.int32)).astype(np.int32), 3, 3)
    c, el = orc.mesh_unit_square_tri(5, 4)
    cen = c[el].mean(axis=1)
    cases["tri4"] = (c, el, ((cen[:, 0] > 0.5).astype(np.int32) + 2 * (cen[:, 1] > 0.5).astype(np.int32)).astype(np.int32), 4, 2)
    for name, (c, el, part, size, dpn) in cases.items():
        n_nat = c.shape[0] * dpn
        out[f"mpi_{name}_coords"], out[f"mpi_{name}_conn"], out[f"mpi_{name}_partition"] = c, el, part
        out[f"mpi_{name}_size_dpn"] = np.array([size, dpn])

        def per_rank(comm, c=c, el=el, part=part, dpn=dpn, n_nat=n_nat):
            mesh, info = extract_local_mesh(Mesh(coords=c, elements=el), part, comm.rank)
            l2g_nodes = np.asarray(info.nodes_local_to_global)
            natural = (l2g_nodes[:, None] * dpn + np.arange(dpn)).ravel().astype(np.int32)
            owned = np.zeros(natural.size, dtype=bool)
            owned[: int(info.n_owned_nodes) * dpn] = True
            layout = _create_dof_layout(natural, owned, n_nat, comm)
            pat = sparse.pattern_from_mesh(mesh, dpn)
            plan = ExchangePlan(layout, pat, comm=comm)
            return layout, plan, (np.asarray(pat.indptr), np.asarray(pat.indices))

        n_global_dofs = None

        def per_rank_exchange(comm, c=c, el=el, part=part, dpn=dpn, n_nat=n_nat):
            """The reference's exchange functions at work: forward ghost fill, reverse add of a vector, reverse add of
            Hessian nonzeros (mpi.py:372-516), with inputs that are functions of global ids so every rank can build them."""
            from dataclasses import replace

            import jax.numpy as jnp

            layout, plan, (ip, ix) = per_rank(comm)
            l2g = np.asarray(layout.local_to_global)
            x_owned = np.sin(0.37 * np.arange(layout.offset, layout.offset + layout.n_owned))
            u_local = np.asarray(plan.make_scatter_fwd_set()(jnp.asarray(x_owned)))
            contrib = np.cos(0.11 * l2g + 0.5 * comm.rank)  # local (owned + ghost) contributions, rank dependent
            owned = np.asarray(plan.make_scatter_rev_add(lambda: jnp.asarray(contrib))())
            pat = sparse.pattern_from_mesh(extract_local_mesh(Mesh(coords=c, elements=el), part, comm.rank)[0], dpn)
            cm = sparse.ColoredMatrix.from_csr(pat)
            nnz_vals = np.sin(0.013 * np.arange(len(ix)) + comm.rank)
            owned_nnz = np.asarray(plan.make_scatter_rev_add(lambda: replace(cm, data=jnp.asarray(nnz_vals)), is_hessian=True)().data)
            return x_owned, u_local, contrib, owned, nnz_vals, owned_nnz

        for r, (x_owned, u_local, contrib, owned, nnz_vals, owned_nnz) in enumerate(run_ranks(size, per_rank_exchange)):
            p = f"mpi_{name}_r{r}_"
            out[p + "x_owned"], out[p + "fwd_u_local"] = x_owned, u_local
            out[p + "rev_contrib"], out[p + "rev_owned"] = contrib, owned
            out[p + "rev_nnz_vals"], out[p + "rev_owned_nnz"] = nnz_vals, owned_nnz

        res = run_ranks(size, per_rank)
        for r, (layout, plan, (ip, ix)) in enumerate(res):
            p = f"mpi_{name}_r{r}_"
            out[p + "natural"], out[p + "owned_mask"] = np.asarray(layout.natural_l2g), np.asarray(layout.owned_mask)
            out[p + "l2g"] = np.asarray(layout.local_to_global)
            out[p + "offset_nowned_ntotal_nglobal"] = np.array([layout.offset, layout.n_owned, layout.n_total, layout.n_global])
            out[p + "self_send"], out[p + "self_recv"] = np.asarray(plan._send_dof), np.asarray(plan._recv_dof)
            out[p + "nbr_ranks"] = np.array([d.rank for d in plan._neighbor_dof_data], dtype=np.int32)
            for d in plan._neighbor_dof_data:
                out[p + f"nbr{d.rank}_send"], out[p + f"nbr{d.rank}_recv"] = np.asarray(d.local_send_idx), np.asarray(d.recv_local_idx)
            h = plan.hessian_layout
            out[p + "pat_indptr"], out[p + "pat_indices"] = ip, ix
            out[p + "h_owned_nnz"] = np.array(h.owned_nnz)
            out[p + "h_owned_ptr"], out[p + "h_owned_indices"] = np.asarray(h.owned_ptr), np.asarray(h.owned_indices)
            out[p + "h_self_send"], out[p + "h_self_recv"] = np.asarray(h.local_send_idx), np.asarray(h.recv_local_idx)
            out[p + "h_nbr_ranks"] = np.array([d.rank for d in h.neighbor_data], dtype=np.int32)
            for d in h.neighbor_data:
                out[p + f"h_nbr{d.rank}_send"], out[p + f"h_nbr{d.rank}_recv"] = np.asarray(d.local_send_idx), np.asarray(d.recv_local_idx)


def compound_fixtures(out):
    """tatva.compound.mpi._layout_from_compound of the UNMODIFIED reference on 3 ranks: stacked full nodal fields, a
    nodal field on a node subset, a shared and a local field (compound/mpi.py:288-494), with the per-field global
    info behind `Compound._g`."""
    from tatva import Mesh
    from tatva.compound import Compound, FieldSize, field
    from tatva.compound.field_types import Local, Nodal, Shared
    from tatva.mesh import extract_local_mesh

    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import tatva_oracle as orc

    c, el = orc.mesh_box_hex((4, 3, 2))
    cen = c[el].mean(axis=1)
    part = ((cen[:, 0] > 0.26).astype(np.int32) + (cen[:, 0] > 0.74).astype(np.int32)).astype(np.int32)
    out["cmp_coords"], out["cmp_conn"], out["cmp_partition"] = c, el, part

    def per_rank(comm):
        mesh, info = extract_local_mesh(Mesh(coords=c, elements=el), part, comm.rank)
        sub = np.arange(0, mesh.coords.shape[0], 3)

        class S(Compound, mesh=mesh, partition_info=info, comm=comm):
            u = field(shape=(FieldSize.AUTO, 3))
            p = field(shape=(FieldSize.AUTO,))
            lam = field(shape=(FieldSize.AUTO, 2), field_type=Nodal(node_ids=sub))
            g = field(shape=(2,), field_type=Shared())
            w = field(shape=(comm.rank + 1, 2), field_type=Local())

        return S, sub

    for r, (S, sub) in enumerate(run_ranks(3, per_rank)):
        L = S.get_layout()
        p = f"cmp_r{r}_"
        out[p + "subset_local_nodes"] = sub
        out[p + "size"] = np.array(S.size)
        out[p + "natural"], out[p + "owned_mask"], out[p + "l2g"] = np.asarray(L.natural_l2g), np.asarray(L.owned_mask), np.asarray(L.local_to_global)
        out[p + "offset_nowned_ntotal_nglobal"] = np.array([L.offset, L.n_owned, L.n_total, L.n_global])
        for name, info in S._global_field_info.items():
            out[p + f"{name}_gshape"] = np.array(info.global_shape, dtype=np.int64)
            out[p + f"{name}_goffset_strides"] = np.array([info.global_base_offset, *info.global_strides], dtype=np.int64)
            if info.global_subset is not None:
                out[p + f"{name}_gsubset"] = np.asarray(info.global_subset)
            out[p + f"{name}_gindices"] = np.asarray(getattr(S._g, name)[(slice(None),) * len(info.global_shape)] if info.global_subset is None else getattr(S._g, name)[info.global_subset])


def periodic_mpi_fixtures(out):
    """Lifter.adapt_layout with PeriodicMPI and Fixed (lifter/base.py:333-425, constraints.py:223-287) run unmodified on
    3 ranks of the Hex8 4x3x2 box with 3 DOFs per node: the x = 1 face follows the x = 0 face (slaves and masters live on
    different ranks, so masters become extra ghosts), the z = 0 layer is fixed except on those two faces."""
    from tatva import Mesh
    from tatva.lifter import Fixed, Lifter, PeriodicMPI
    from tatva.mesh import extract_local_mesh
    from tatva.mpi import _create_dof_layout

    import jax.numpy as jnp

    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import tatva_oracle as orc

    dpn = 3
    c, el = orc.mesh_box_hex((4, 3, 2))
    cen = c[el].mean(axis=1)
    part = ((cen[:, 0] > 0.26).astype(np.int32) + (cen[:, 0] > 0.74).astype(np.int32)).astype(np.int32)
    x0 = np.where(c[:, 0] < 1e-12)[0]
    x1 = np.where(c[:, 0] > c[:, 0].max() - 1e-12)[0]
    key = lambda idx: np.lexsort((c[idx, 1], c[idx, 2]))  # noqa: E731
    x0, x1 = x0[key(x0)], x1[key(x1)]
    assert np.allclose(c[x0][:, 1:], c[x1][:, 1:])
    g_slaves = (x1[:, None] * dpn + np.arange(dpn)).ravel()
    g_masters = (x0[:, None] * dpn + np.arange(dpn)).ravel()
    out["pmpi_coords"], out["pmpi_conn"], out["pmpi_partition"] = c, el, part
    out["pmpi_slaves"], out["pmpi_masters"] = g_slaves, g_masters

    def per_rank(comm):
        mesh, info = extract_local_mesh(Mesh(coords=c, elements=el), part, comm.rank)
        l2g_nodes = np.asarray(info.nodes_local_to_global)
        natural = (l2g_nodes[:, None] * dpn + np.arange(dpn)).ravel().astype(np.int32)
        owned = np.zeros(natural.size, dtype=bool)
        owned[: int(info.n_owned_nodes) * dpn] = True
        layout = _create_dof_layout(natural, owned, c.shape[0] * dpn, comm)
        lc = np.asarray(mesh.coords)
        fixed_nodes = np.where((lc[:, 2] < 1e-12) & (lc[:, 0] > 1e-12) & (lc[:, 0] < c[:, 0].max() - 1e-12))[0]
        fixed = (fixed_nodes[:, None] * dpn + np.arange(dpn)).ravel()
        lifter = Lifter(layout.n_total, Fixed(jnp.asarray(fixed), 0.25), PeriodicMPI(jnp.asarray(g_slaves), jnp.asarray(g_masters), layout, comm=comm))
        reduced, lifter2 = lifter.adapt_layout(layout, comm)
        u_red = np.sin(1.0 + np.arange(lifter2.size_reduced) + 100.0 * comm.rank)
        return fixed, reduced, lifter2, u_red, np.asarray(lifter2.lift_from_zeros(jnp.asarray(u_red)))

    for r, (fixed, reduced, lifter2, u_red, lifted_full) in enumerate(run_ranks(3, per_rank)):
        p = f"pmpi_r{r}_"
        out[p + "fixed_local_dofs"] = fixed
        out[p + "red_l2g"], out[p + "red_owned_mask"], out[p + "red_natural"] = np.asarray(reduced.local_to_global), np.asarray(reduced.owned_mask), np.asarray(reduced.natural_l2g)
        out[p + "red_offset_nowned_ntotal_nglobal"] = np.array([reduced.offset, reduced.n_owned, reduced.n_total, reduced.n_global])
        out[p + "lifter_size_extra"] = np.array([lifter2.size, getattr(lifter2, "_nb_extra_ghost_dofs", 0)])
        out[p + "free_dofs"] = np.asarray(lifter2.free_dofs)
        per = lifter2.constraints[1]
        out[p + "periodic_dofs"], out[p + "periodic_masters"] = np.asarray(per.dofs), np.asarray(per.master_dofs)
        out[p + "u_red"], out[p + "lift_from_zeros"] = u_red, lifted_full


def allreduce_plan_fixtures(out):
    """AllreducePlan (mpi.py:519-711) of the unmodified reference on 4 ranks over the Tri3 5x4 global pattern with 2 DOFs
    per node: block ranges, owned CSR slices, allgather of the owned blocks, all-reduced owned vector and Hessian."""
    from dataclasses import replace

    import jax.numpy as jnp
    from tatva import Mesh, sparse
    from tatva.mpi import AllreducePlan

    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import tatva_oracle as orc

    c, el = orc.mesh_unit_square_tri(5, 4)
    pat = sparse.pattern_from_mesh(Mesh(coords=jnp.asarray(c), elements=jnp.asarray(el)), 2)
    n = pat.shape[0]
    out["arp_coords"], out["arp_conn"] = c, el

    def per_rank(comm):
        plan = AllreducePlan(n, pat, comm=comm)
        x_owned = np.sin(0.3 * np.arange(plan.rstart, plan.rend))
        full = np.asarray(plan.make_allgather()(x_owned))
        vec = np.cos(0.07 * np.arange(n) * (comm.rank + 1))
        owned_vec = np.asarray(plan.make_allreduce_owned(lambda: jnp.asarray(vec))())
        cm = sparse.ColoredMatrix.from_csr(pat)
        vals = np.sin(0.01 * np.arange(pat.indices.shape[0]) + comm.rank)
        K = plan.make_allreduce_owned(lambda: replace(cm, data=jnp.asarray(vals)), is_hessian=True)()
        return plan, x_owned, full, vec, owned_vec, vals, K

    for r, (plan, x_owned, full, vec, owned_vec, vals, K) in enumerate(run_ranks(4, per_rank)):
        p = f"arp_r{r}_"
        out[p + "range_nnz"] = np.array([plan.rstart, plan.rend, plan.owned_nnz])
        out[p + "owned_ptr"], out[p + "owned_indices"] = (np.asarray(a) for a in plan.owned_csr)
        out[p + "x_owned"], out[p + "allgather"] = x_owned, full
        out[p + "vec"], out[p + "owned_vec"] = vec, owned_vec
        out[p + "vals"], out[p + "K_data"], out[p + "K_shape"] = vals, np.asarray(K.data), np.array(K.shape)
        out[p + "K_indptr"], out[p + "K_indices"] = np.asarray(K.indptr), np.asarray(K.indices)
