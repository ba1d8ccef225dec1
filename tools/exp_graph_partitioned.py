"""EXPERIMENT for the next round (NOT yet run on a GPU): capture one partitioned operator application
(interior + boundary element kernels, peer pull / push, signal-pad barriers) in a CUDA graph and compare with the
eager launch sequence.  Motivation: at config-5 size the 8-GPU step (0.0985 ms) is bound by Python launch latency
while the kernels take 0.07 ms.  Run under torchrun like tools/bench_c5.py:

    torchrun --nproc-per-node 8 tools/exp_graph_partitioned.py 55 peer

Prints eager and graph times (max over ranks) and the max abs difference of the two results."""
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
mk = pop.new_symmetric_vector if (pop.halo == "peer" and world > 1) else pop.new_local_vector
s, d, y, y2 = mk(), mk(), mk(), mk()
s.copy_(torch.as_tensor(s0.ravel(), device=dev))
d.copy_(torch.as_tensor(np.random.default_rng(1 + rank).normal(size=s0.shape).ravel(), device=dev))
pop.fill_ghosts(s)


def timed(fn, reps=200, warm=10):
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


ms_eager = timed(lambda: pop.hvp(s, d, y))
out = {"n_gpus": world, "halo": pop.halo, "eager_ms": round(ms_eager, 4)}
try:
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        pop.hvp(s, d, y2)  # warm-up outside the capture
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pop.hvp(s, d, y2)
    ms_graph = timed(g.replay)
    out["graph_ms"] = round(ms_graph, 4)
    out["max_abs_diff"] = float((y2[: pop.n_owned] - y[: pop.n_owned]).abs().max())
except Exception as exc:  # noqa: BLE001  (experiment: report why the capture is not possible)
    out["graph_error"] = repr(exc)[:300]
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
