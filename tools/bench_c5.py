"""Config 5 at scale (run under torchrun): compound (u, phi) Tet4 phase-field operator, one 55^3-cell block
(998 250 tets, 4 DOFs per node) per GPU — 8 GPUs = the n = 110 box of SURVEY.md §8 — coupled residual and HVP
with the halo exchange overlapped.  Prints one JSON line on rank 0."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tatva_b200 import element, materials
from tatva_b200.distributed import PartitionedOperator, structured_tet_block
from bench_dist import GRID

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
halo = sys.argv[2] if len(sys.argv) > 2 else "peer"
mesh, info = structured_tet_block(n, GRID[world], rank)
mat = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.05, 1e-6)
pop = PartitionedOperator(mesh, info, element.Tetrahedron4(), mat, device=dev, overlap=True, halo=halo)
c = np.asarray(mesh.coords)
t = 2 * np.pi
s0 = np.concatenate([0.02 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 2]), np.sin(t * c[:, 2]) * np.cos(t * c[:, 0])], -1), 0.5 + 0.3 * np.sin(6 * c[:, :1])], axis=1)
mk = pop.new_symmetric_vector if (halo == "peer" and world > 1) else pop.new_local_vector
s, d, y = mk(), mk(), mk()
s.copy_(torch.as_tensor(s0.ravel(), device=dev))
d.copy_(torch.as_tensor(np.random.default_rng(1 + rank).normal(size=s0.shape).ravel(), device=dev))
pop.fill_ghosts(s)


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


ms_hvp = timed(lambda: pop.hvp(s, d, y))
ms_res = timed(lambda: pop.residual(s, y))
finite = bool(torch.isfinite(y[: pop.n_owned]).all())
if rank == 0:
    print(json.dumps({"config": "C5 compound (u,phi) Tet4 phase-field", "n_gpus": world, "cells_per_gpu": n**3, "tets_per_gpu": 6 * n**3, "dofs_global": pop.n_global, "halo": pop.halo,
                      "hvp_ms": round(ms_hvp, 4), "hvp_gdofs": round(pop.n_global / ms_hvp / 1e6, 3), "residual_ms": round(ms_res, 4), "residual_gdofs": round(pop.n_global / ms_res / 1e6, 3), "finite": finite}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
