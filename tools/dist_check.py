"""Multi-GPU parity check (run under torchrun): the distributed HVP / residual of a partitioned Hex8 box
must equal the single-GPU result on the global mesh, in the plan's global numbering."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import tatva_b200
from tatva_b200 import element, materials
from tatva_b200.distributed import PartitionedOperator, structured_hex_block
from tatva_b200.mesh import Mesh
from bench_dist import GRID, smooth_u

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
grid = GRID[world]
mat = materials.NeoHookean(500.0, 1000.0)
out = {}
HALOS = os.environ.get("TATVA_CHECK_HALOS", "nccl,peer").split(",")
for overlap, halo in [(o, h) for h in HALOS for o in (False, True)]:
    mesh, info = structured_hex_block(n, grid, rank)
    pop = PartitionedOperator(mesh, info, element.Hexahedron8(), mat, device=dev, overlap=overlap, halo=halo)
    l2g = info.nodes_local_to_global
    # global reference on every rank: same jittered coordinates via the block builder with a 1x1x1 grid
    shape = (grid[0] * n, grid[1] * n, grid[2] * n)
    gm = Mesh.box_hex(shape)
    # jitter must match structured_hex_block: rebuild from its hash
    from tatva_b200.distributed import _hash_uniform
    m = max(shape)
    gc = gm.coords + 0.1 * (shape[0] / m / shape[0]) * _hash_uniform(np.arange(gm.coords.shape[0]), 0)
    assert np.abs(gc[l2g] - mesh.coords).max() < 1e-14
    gop = tatva_b200.Operator(Mesh(coords=gc, elements=gm.elements), element.Hexahedron8(), device=dev)
    gu = smooth_u(gc)
    gv = np.random.default_rng(7).normal(size=gc.shape)
    ref_hvp = gop.hvp(mat)(gu, gv).cpu().numpy()
    ref_res = gop.residual(mat)(gu).cpu().numpy()
    if halo == "peer":
        u_local, v_local, y_buf = (pop.new_symmetric_vector() for _ in range(3))
        u_local.copy_(torch.as_tensor(gu[l2g].ravel(), device=dev))
        v_local.copy_(torch.as_tensor(gv[l2g].ravel(), device=dev))
    else:
        u_local = torch.as_tensor(gu[l2g].ravel(), device=dev)
        v_local = torch.as_tensor(gv[l2g].ravel(), device=dev)
        y_buf = None
    v_local[pop.n_owned:] = 0.0  # ghosts must come from the exchange
    y = pop.hvp(u_local, v_local, y_buf)
    torch.cuda.synchronize()
    no = info.n_owned_nodes
    e1 = np.abs(y[: pop.n_owned].cpu().numpy().reshape(-1, 3) - ref_hvp[l2g[:no]]).max() / np.abs(ref_hvp).max()
    if halo == "peer":
        u2, r_buf = pop.new_symmetric_vector(), pop.new_symmetric_vector()
        u2.copy_(u_local)
    else:
        u2, r_buf = u_local.clone(), None
    u2[pop.n_owned:] = 0.0
    r = pop.residual(u2, r_buf)
    torch.cuda.synchronize()
    e2 = np.abs(r[: pop.n_owned].cpu().numpy().reshape(-1, 3) - ref_res[l2g[:no]]).max() / np.abs(ref_res).max()
    out[f"halo={halo},overlap={overlap}"] = {"hvp_rel_err": float(e1), "residual_rel_err": float(e2), "n_boundary": pop.n_boundary, "n_global": pop.n_global}
    assert e1 < 1e-12 and e2 < 1e-12, (rank, overlap, e1, e2)
# ---- distributed CG (plain and Jacobi) on the free DOFs == single-GPU CG on the global mesh ----
# ("CG dot products: ncclAllReduce of 1-2 scalars"; uses the last pop: halo=HALOS[-1], overlap on)
from tatva_b200.solver import ConjugateGradient, DistributedConjugateGradient
nz0 = (shape[0] + 1) * (shape[1] + 1)  # global node ids of the z = 0 layer come first
g_pinned = np.zeros(gc.shape, dtype=bool)
g_pinned[:nz0] = True
gus = 0.02 * gu  # small strains: SPD tangent
gb = np.random.default_rng(21).normal(size=gc.shape)
gfree = torch.as_tensor((~g_pinned).ravel(), device=dev).to(torch.float64)
gu_t = torch.as_tensor(gus.ravel(), device=dev)
tmp = torch.empty_like(gu_t)


def g_matvec(p, o):
    gop._raw_hvp(mat, gu_t, p, out=tmp)
    torch.mul(tmp, gfree, out=o)
    return o


gcg = ConjugateGradient(g_matvec, gc.size, dev, use_graph=False)
x_ref, ginfo = gcg.solve(torch.as_tensor(gb.ravel(), device=dev) * gfree, tol=1e-11, maxiter=5000, check_every=10)
x_ref = x_ref.cpu().numpy().reshape(-1, 3)
peer = pop.halo == "peer"
us = pop.new_symmetric_vector() if peer else pop.new_local_vector()
us.copy_(torch.as_tensor(gus[l2g].ravel(), device=dev))
pinned_owned = g_pinned[l2g[:no]].ravel()
b_owned = torch.as_tensor(gb[l2g[:no]].ravel(), device=dev)
for jac in (False, True):
    diag = pop.hessian_diagonal(us)[: pop.n_owned].clone() if jac else None
    if jac:  # the assembled diagonal == the single-GPU one
        gd = gop.hessian_diagonal(mat, gu_t.view(-1, 3)).cpu().numpy()
        ed = np.abs(diag.cpu().numpy().reshape(-1, 3) - gd[l2g[:no]]).max() / np.abs(gd).max()
        assert ed < 1e-12, (rank, ed)
        out["hessian_diag_rel_err"] = float(ed)
    dcg = DistributedConjugateGradient(pop, pinned_owned=pinned_owned, jacobi_diagonal=diag)
    dcg.set_state(us)
    x_own, dinfo = dcg.solve(b_owned, tol=1e-11, maxiter=5000, check_every=10)
    ex = np.abs(x_own.cpu().numpy().reshape(-1, 3) - x_ref[l2g[:no]]).max() / np.abs(x_ref).max()
    assert dinfo["converged"] and ex < 1e-8, (rank, jac, dinfo, ex)
    out[f"distributed_cg_jacobi={jac}"] = {"rel_err_vs_single_gpu": float(ex), "iterations": dinfo["iterations"], "single_gpu_iterations": ginfo["iterations"]}
# ---- distributed Newton: same minimiser as the single-GPU Newton with a Lifter on the global mesh ----
from tatva_b200.solver import distributed_newton_solve, newton_solve
from tatva_b200.lifter import Fixed, Lifter
top = np.arange(gc.shape[0] - nz0, gc.shape[0])  # global node ids of the top layer
glift = Lifter(gc.size, Fixed((np.arange(nz0)[:, None] * 3 + np.arange(3)).ravel(), 0.0), Fixed(top * 3 + 2, 0.03), Fixed((top[:, None] * 3 + np.arange(2)).ravel(), 0.0))
u_red, ghist = newton_solve(gop, mat, glift, tol=1e-10, cg_tol=1e-12)
u_glob = glift.lift_from_zeros(u_red).cpu().numpy().reshape(-1, 3)
g_pin = np.zeros(gc.shape, dtype=bool)
g_pin[:nz0] = True
g_pin[top] = True
g_init = np.zeros(gc.shape)
g_init[top, 2] = 0.03
un = pop.new_symmetric_vector() if peer else pop.new_local_vector()
un.copy_(torch.as_tensor(g_init[l2g].ravel(), device=dev))
un, dhist = distributed_newton_solve(pop, un, pinned_owned=g_pin[l2g[:no]].ravel(), tol=1e-10, cg_tol=1e-12)
en = np.abs(un[: pop.n_owned].cpu().numpy().reshape(-1, 3) - u_glob[l2g[:no]]).max() / np.abs(u_glob).max()
assert en < 1e-7, (rank, en, dhist)
out["distributed_newton"] = {"rel_err_vs_single_gpu": float(en), "newton_steps": len(dhist), "single_gpu_newton_steps": len(ghist)}
# ---- public plan API on CUDA tensors across ranks (reference call stack mpi.py:372-409, :479-516, :609-711) ----
from tatva_b200.mpi import AllreducePlan
mesh, info = structured_hex_block(n, grid, rank)
pop = PartitionedOperator(mesh, info, element.Hexahedron8(), mat, device=dev, overlap=False)
l2g = info.nodes_local_to_global
shape = (grid[0] * n, grid[1] * n, grid[2] * n)
n_glob_nodes = (shape[0] + 1) * (shape[1] + 1) * (shape[2] + 1)
gvals = np.random.default_rng(11).normal(size=(n_glob_nodes, 3))
x_owned = torch.as_tensor(gvals[l2g[: info.n_owned_nodes]].ravel(), device=dev)
u_local = pop.plan.make_scatter_fwd_set()(x_owned)
assert np.array_equal(u_local.cpu().numpy().reshape(-1, 3), gvals[l2g]), "scatter_fwd_set mismatch"
owned = pop.plan.make_scatter_rev_add(lambda ul: ul)(u_local)
mult = np.zeros(n_glob_nodes)
counts = [None] * world
dist.all_gather_object(counts, l2g)
for lg in counts:
    mult[lg] += 1
ref_owned = (gvals * mult[:, None])[l2g[: info.n_owned_nodes]].ravel()
assert np.allclose(owned.cpu().numpy(), ref_owned, rtol=1e-14), "scatter_rev_add mismatch"
ap = AllreducePlan(global_size=1000, comm=dist.group.WORLD)
full = ap.make_allgather()(torch.full((ap.local_size,), float(rank + 1), device=dev, dtype=torch.float64))
assert full.numel() == 1000 and float(full[ap.rstart]) == rank + 1
red = ap.make_allreduce_owned(lambda x: x)(torch.arange(1000, device=dev, dtype=torch.float64) * (rank + 1))
assert np.allclose(red.cpu().numpy(), np.arange(1000)[ap.rstart : ap.rend] * (world * (world + 1) / 2))
out["plans_on_gpu"] = "ok"
# ---- config 5: compound (u, phi) Tet4 phase-field operator on a partitioned mesh (generic extract_local_mesh path) ----
from tatva_b200.mesh import extract_local_mesh, block_partition
nt = 6
gshape = (grid[0] * nt, grid[1] * nt, grid[2] * nt)
tm = Mesh.box_tet((gshape[0] / max(gshape), gshape[1] / max(gshape), gshape[2] / max(gshape)), gshape)
tc = tm.coords + np.array([0.5, 0.5, 0.0]) + 0.02 / nt * np.random.default_rng(5).uniform(-1, 1, tm.coords.shape)
tmesh = Mesh(coords=tc, elements=tm.elements)
tpart = np.repeat(block_partition(gshape, world), 6)  # 6 tets per cell, cells in x-fastest order
pf = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.1, 1e-6)
lm, linfo = extract_local_mesh(tmesh, tpart, rank)
ppop = PartitionedOperator(lm, linfo, element.Tetrahedron4(), pf, device=dev, overlap=True)
gs = np.concatenate([0.002 * np.random.default_rng(6).normal(size=tc.shape), np.random.default_rng(7).uniform(0, 0.8, size=(len(tc), 1))], axis=1)
gt = np.random.default_rng(8).normal(size=gs.shape)
gop5 = tatva_b200.Operator(tmesh, element.Tetrahedron4(), device=dev)
ref5 = gop5.hvp(pf)(gs, gt).cpu().numpy().reshape(-1, 4)
tl2g = linfo.nodes_local_to_global
s_local = torch.as_tensor(gs[tl2g].ravel(), device=dev)
t_local = torch.as_tensor(gt[tl2g].ravel(), device=dev)
t_local[ppop.n_owned:] = 0.0
y5 = ppop.hvp(s_local, t_local)
torch.cuda.synchronize()
e5 = np.abs(y5[: ppop.n_owned].cpu().numpy().reshape(-1, 4) - ref5[tl2g[: linfo.n_owned_nodes]]).max() / np.abs(ref5).max()
assert e5 < 1e-12, (rank, e5)
out["compound_pf_tet4_hvp_rel_err"] = float(e5)
allr = [None] * world
dist.all_gather_object(allr, out)
if rank == 0:
    print(json.dumps({"world": world, "n": n, "per_rank": allr}))
dist.barrier()
dist.destroy_process_group()
