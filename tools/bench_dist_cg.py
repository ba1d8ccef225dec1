"""One CG iteration around the distributed HVP (run under torchrun): Hex8 128^3 per GPU, z = 0 face pinned, neo-Hookean
tangent at a small smooth strain.  Per iteration: halo-overlapped HVP + vector kernels + two device-side scalar
all-reduces.  Prints one JSON line on rank 0 (time = max over ranks)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tatva_b200 import element, materials
from tatva_b200.distributed import PartitionedOperator, structured_hex_block
from tatva_b200.solver import DistributedConjugateGradient
from bench_dist import GRID, smooth_u

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
halo = sys.argv[2] if len(sys.argv) > 2 else "peer"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
mesh, info = structured_hex_block(n, GRID[world], rank)
mat = materials.NeoHookean(500.0, 1000.0)
pop = PartitionedOperator(mesh, info, element.Hexahedron8(), mat, device=dev, overlap=True, halo=halo)
c = np.asarray(mesh.coords)
peer = pop.halo == "peer"
u = pop.new_symmetric_vector() if peer else pop.new_local_vector()
u.copy_(torch.as_tensor(0.02 * smooth_u(c).ravel(), device=dev))
no = info.n_owned_nodes
pinned = np.zeros((no, 3), dtype=bool)
pinned[c[:no, 2] < 0.25 / (n * GRID[world][2])] = True
b = torch.as_tensor(np.random.default_rng(3 + rank).normal(size=pop.n_owned), device=dev)
out = {}
for jac in (False, True):
    diag = pop.hessian_diagonal(u)[: pop.n_owned].clone() if jac else None
    cg = DistributedConjugateGradient(pop, pinned_owned=pinned.ravel(), jacobi_diagonal=diag)
    cg.set_state(u)
    cg.solve(b, tol=0.0, maxiter=10, check_every=10)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    x, info_cg = cg.solve(b, tol=0.0, maxiter=iters, check_every=iters)
    a1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a0.elapsed_time(a1) / info_cg["iterations"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["pcg_jacobi" if jac else "cg"] = {"ms_per_iteration": round(float(ms), 4), "gdofs_per_iteration_rate": round(pop.n_global / float(ms) / 1e6, 3), "residual_norm": info_cg["residual_norm"]}
if rank == 0:
    print(json.dumps({"bench": "distributed CG iteration, Hex8 neo-Hookean", "n_gpus": world, "n_per_gpu": n, "dofs_global": pop.n_global, "halo": pop.halo, **out}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
