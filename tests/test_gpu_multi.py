"""Multi-GPU parity inside the GPU suite: two ranks (one process per GPU, NCCL) run the partitioned Hex8 operator and
compare it with the single-GPU result on the global mesh.  Skipped on a one-GPU box; `bench.py` runs the same check
before its timed region at every N > 1 and prints it as `parity` in its JSON line."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from tatva_b200 import materials
from bench_dist import distributed_parity
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{{lr}}")
dist.init_process_group("nccl", device_id=dev)
out = {{}}
for halo in ("nccl", "nccl_torch", "peer"):
    out[halo] = distributed_parity(10, rank, world, dev, materials.NeoHookean(500.0, 1000.0), halo=halo)
if rank == 0:
    print("PARITY " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_partitioned_operator_matches_single_gpu(tmp_path):
    import json

    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("PARITY ")][-1]
    out = json.loads(line[len("PARITY "):])
    for halo, r in out.items():
        assert r["hvp_rel_err"] <= 1e-12 and r["residual_rel_err"] <= 1e-12, (halo, r)
