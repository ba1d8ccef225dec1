#!/usr/bin/env python
"""Headline benchmark: matrix-free Hex8 neo-Hookean HVP, DOFs/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CELLS] [--impl ours|reference]

A "step" is one application y = H(u) v of the hot path over the whole mesh (config 3: Hex8 128^3,
6 440 067 DOFs per GPU).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MU, LMBDA = 500.0, 1000.0  # reference tests/test_sparse_tracer.py:126
METRIC = "hex8_neohookean_matrix_free_hvp_dofs_per_s"


def synthetic_inputs(n, rank=0):
    """jittered unit-cube Hex8 box, smooth displacement, Gaussian direction."""
    from tatva_b200.mesh import Mesh

    mesh = Mesh.box_hex(n)
    c = mesh.coords
    c = c + 0.1 * (1.0 / n) * np.random.default_rng(0).uniform(-1, 1, c.shape)
    two_pi = 2 * np.pi
    u = 0.05 * np.stack(
        [
            np.sin(two_pi * c[:, 0]) * np.cos(two_pi * c[:, 1]),
            np.sin(two_pi * c[:, 1]) * np.cos(two_pi * c[:, 2]),
            np.sin(two_pi * c[:, 2]) * np.cos(two_pi * c[:, 0]),
        ],
        -1,
    )
    v = np.random.default_rng(1 + rank).normal(size=c.shape)
    return c, mesh.elements, u, v


def algorithmic_bytes(n_nodes, n_elems):
    """each array touched once, y write-only: 8 (3 vec x 3 + 3) N + 4 x 8 E."""
    return 8 * (3 * 3 * n_nodes + 3 * n_nodes) + 4 * 8 * n_elems


FLOP_PER_ELEMENT = 7944  # SURVEY.md §8(d) nominal count for the Hex8 NH HVP


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): an NVML polling
    thread (2 ms period; the timed region can be as short as ~10 ms), nvidia-smi as a fallback."""

    REASONS = {
        "hw_slowdown": 0x8,
        "sw_power_cap": 0x4,
        "sw_thermal_slowdown": 0x20,
        "hw_thermal_slowdown": 0x40,
    }

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.power, self.reasons = [], [], set()
        self.max_mhz = None
        self._stop = None
        self._thread = None
        self._nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            try:
                return int(ids[self.gpu])
            except (ValueError, IndexError):
                return self.gpu
        return self.gpu

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, h)
        except Exception:  # noqa: BLE001
            self._nvml = None
            return
        self._stop = threading.Event()

        def poll():
            nv, hh = self._nvml
            while not self._stop.is_set():
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(hh, nv.NVML_CLOCK_SM)))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(hh) / 1000.0)
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hh)
                    for name, bit in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
                self._stop.wait(0.002)

        self._thread = threading.Thread(target=poll, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if not self.sm:
            return self._smi_once()
        return {
            "sm_mhz": statistics.median(self.sm),
            "sm_max_mhz": self.max_mhz,
            "samples": len(self.sm),
            "power_w_max": max(self.power) if self.power else None,
            "reasons": sorted(self.reasons),
            "source": "nvml, 2 ms polling during the timed region",
        }

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._physical_index())], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "samples": 1, "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "source": "nvidia-smi (single sample after the timed region)"}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}


def cpu_port_baseline(n_sample, threads=None):
    """Time the oracle (CPU restatement of the reference's arithmetic) on a bounded sample."""
    from oracle import tatva_oracle as orc

    c, el, u, v = synthetic_inputs(n_sample)
    mat = orc.NeoHookean(MU, LMBDA)
    try:
        from oracle import c_oracle

        fn = lambda: c_oracle.hvp("hex8", (MU, LMBDA), c, el, u, v)  # noqa: E731
        cores = c_oracle.num_threads()
        kind_note = "C/OpenMP port (oracle/tatva_oracle.c)"
    except Exception:
        fn = lambda: orc.hvp("hex8", mat, c, el, u, v)  # noqa: E731
        cores = 1
        kind_note = "NumPy port (oracle/tatva_oracle.py)"
    fn()
    reps, t_total = 0, 0.0
    while reps < 3 or (t_total < 10.0 and reps < 50):
        t0 = time.perf_counter()
        fn()
        t_total += time.perf_counter() - t0
        reps += 1
    t = t_total / reps
    return {
        "value": 3 * c.shape[0] / t,
        "unit": "DOF/s",
        "cores": cores,
        "kind": "port",
        "sample": f"Hex8 {n_sample}^3 ({3 * c.shape[0]} DOFs), mean of {reps} reps, {kind_note}; os.cpu_count()={os.cpu_count()}",
        "ms_per_step": t * 1e3,
    }


def run_reference(args):
    """Reference arm: the reference's JAX path cannot run (no jax in the image); the CPU port of its
    arithmetic is timed on the host cores instead, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.ref_n
    base = cpu_port_baseline(n_sample)
    ms = base.pop("ms_per_step")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": "DOF/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"Hex8 {args.n}^3 per GPU, neo-Hookean (mu=500, lambda=1000) matrix-free HVP",
            "sample": f"CPU port of the reference arithmetic on a {n_sample}^3 block of the same mesh family (bounded sample); "
                      "the reference's JAX path cannot run: jax is not installable in this image",
        },
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--n", type=int, default=128, help="cells per side of the per-GPU Hex8 block")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-n", type=int, default=48, help="cells per side of the CPU sample")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default=os.environ.get("TATVA_HALO", "peer"), choices=["nccl", "peer"], help="multi-GPU halo transport")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    from tatva_b200 import materials

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from bench_dist import DistributedHex8Problem  # multi-GPU decomposition + halo exchange

    prob = DistributedHex8Problem(args.n, rank, world, dev, materials.NeoHookean(MU, LMBDA), variant=args.variant, halo=args.halo)
    n_dofs_global = prob.n_dofs_global
    step = prob.step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    # kernel-only duration of the element kernel (per launch), events on the launching stream
    k_ms = prob.time_kernel_only(args.steps)
    # end-to-end through the public API with host buffers
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        prob.step_e2e()
    prob.e2e_finish()
    barrier()
    e0.record()
    n_e2e = max(3, min(args.steps, 20))
    for _ in range(n_e2e):
        prob.step_e2e()
    prob.e2e_finish()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total, ms_e2e, k_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, k_ms = (float(x) for x in t.tolist())
    ms_step = ms_total / args.steps
    ms_e2e_step = ms_e2e / n_e2e

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (1-GPU, 128^3)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_hvp_traffic.json")))
            if args.n == 128:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except (OSError, KeyError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        alg_bytes = algorithmic_bytes(prob.local_nodes, prob.local_elems)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        fp64_peak = prob.fp64_peak_tflops()
        fp64_ach = FLOP_PER_ELEMENT * prob.local_elems / (k_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC,
            "value": n_dofs_global / (ms_step * 1e-3),
            "unit": "DOF/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"Hex8 {args.n}^3 per GPU, neo-Hookean (mu=500, lambda=1000) matrix-free HVP, {prob.partition_desc}",
                "dofs_global": n_dofs_global,
                "dofs_per_gpu": 3 * prob.local_nodes,
                "l2_policy": "inputs larger than L2 (273 MB working set per GPU vs 126 MB L2)",
                "parallelism": prob.partition_desc,
            },
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": traffic,
                "peak_source": peak_src,
                "kernel": "k_hex8_nh_hvp",
                "kernel_ms": k_ms,
                "algorithmic_bytes": alg_bytes,
                "note": "binding roof is FP64 (AI ~ 61 flop/B), see fp64 block",
                "fp64": {
                    "achieved_tflops_nominal": fp64_ach,
                    "peak_tflops_measured_dfma": fp64_peak,
                    "frac": fp64_ach / fp64_peak if fp64_peak else None,
                    "flop_per_element_nominal": FLOP_PER_ELEMENT,
                },
            },
            "e2e": {
                "value": n_dofs_global / (ms_e2e_step * 1e-3),
                "unit": "DOF/s",
                "ms_per_step": ms_e2e_step,
                "h2d_bytes_per_step": prob.h2d_bytes,
                "d2h_bytes_per_step": prob.d2h_bytes,
            },
            "gpu_launches": prob.launches_per_step * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_baseline(args.ref_n)
            line["cpu_baseline"].pop("ms_per_step", None)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
