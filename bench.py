#!/usr/bin/env python
"""Headline benchmark: matrix-free Hex8 neo-Hookean HVP, DOFs/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CELLS] [--impl ours|reference] [--scaling weak|strong]

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


def _cpu_hvp_fn(n, threads=None):
    """(callable running ONE CPU HVP over the whole Hex8 n^3 mesh, n_dofs, cores, description).  The C/OpenMP oracle
    (oracle/tatva_oracle.c) on all host cores; torchrun exports OMP_NUM_THREADS=1 to its workers, so the thread count
    is set explicitly."""
    from oracle import tatva_oracle as orc

    c, el, u, v = synthetic_inputs(n)
    try:
        from oracle import c_oracle

        c_oracle.set_num_threads(threads or os.cpu_count() or 1)
        fn = lambda: c_oracle.hvp("hex8", (MU, LMBDA), c, el, u, v)  # noqa: E731
        cores = c_oracle.num_threads()
        note = "C/OpenMP port (oracle/tatva_oracle.c)"
    except Exception:  # noqa: BLE001
        mat = orc.NeoHookean(MU, LMBDA)
        fn = lambda: orc.hvp("hex8", mat, c, el, u, v)  # noqa: E731
        cores = 1
        note = "NumPy port (oracle/tatva_oracle.py)"
    return fn, 3 * c.shape[0], cores, note


def cpu_port_baseline(n, budget_s=12.0):
    """cpu_baseline of our own arm: the oracle timed on the host cores over the SAME mesh (Hex8 n^3), as many whole-mesh
    repetitions as fit in ~`budget_s` seconds (at least 3)."""
    fn, n_dofs, cores, note = _cpu_hvp_fn(n)
    fn()
    reps, t_total = 0, 0.0
    while reps < 3 or (t_total < budget_s and reps < 50):
        t0 = time.perf_counter()
        fn()
        t_total += time.perf_counter() - t0
        reps += 1
    t = t_total / reps
    return {
        "value": n_dofs / t,
        "unit": "DOF/s",
        "cores": cores,
        "kind": "port",
        "sample": f"Hex8 {n}^3 ({n_dofs} DOFs, the whole config-3 mesh), mean of {reps} reps, {note}; os.cpu_count()={os.cpu_count()}",
        "ms_per_step": t * 1e3,
    }


def run_reference(args):
    """Reference arm.  The reference's JAX path cannot run (jax is not installable in this image), so the
    CPU port of its arithmetic is timed on ALL host cores, on the same workload as our arm (the Hex8 `--n`^3 mesh,
    128^3 by default; `--scaling strong`: the fixed `--strong-n`^3 mesh), `--warmup` untimed and `--steps` timed
    whole-mesh applications.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.strong_n if args.scaling == "strong" else args.n
    fn, n_dofs, cores, note = _cpu_hvp_fn(n)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        fn()
        done += 1
        if time.perf_counter() - t0 > args.ref_budget_s:  # safety valve; `steps` below reports what was timed
            break
    t = (time.perf_counter() - t0) / done
    base = {"value": n_dofs / t, "unit": "DOF/s", "cores": cores, "kind": "port",
            "sample": f"Hex8 {n}^3 ({n_dofs} DOFs): the whole mesh of one GPU's workload, every step; {note}; os.cpu_count()={os.cpu_count()}"}
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": "DOF/s",
        "n_gpus": args.gpus,
        "steps": done,
        "warmup": args.warmup,
        "ms_per_step": t * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"Hex8 {n}^3, neo-Hookean (mu=500, lambda=1000) matrix-free HVP",
            "same_config": True,
            "note": "CPU port of the reference arithmetic (the reference's JAX path cannot run: jax is not installable in this image); "
                    "one host, all cores; at N > 1 GPUs the CPU arm still processes one GPU's mesh per step (there is one host)",
        },
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _timed_steps(step, steps, barrier, torch):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--n", type=int, default=128, help="cells per side of the per-GPU Hex8 block (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: a FIXED --strong-n^3 mesh cut into 1/2/4/8 blocks (config 4)")
    ap.add_argument("--strong-n", type=int, default=256, help="cells per side of the fixed global mesh for strong scaling")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="wall-clock cap of the reference arm's timed loop")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` (configs 1, 2, 5) and `strong_scaling` blocks")
    ap.add_argument("--halo", default=os.environ.get("TATVA_HALO", "peer"), choices=["nccl", "nccl_torch", "peer"], help="multi-GPU halo transport")
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

    from bench_dist import GRID, DistributedHex8Problem, distributed_parity, numa_pin, secondary_c5, secondary_single_gpu

    numa = numa_pin(local_rank)  # before any pinned host allocation
    mat = materials.NeoHookean(MU, LMBDA)
    grid = GRID[world]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def strong_block():
        N = args.strong_n
        return (N // grid[0], N // grid[1], N // grid[2])

    parity = None
    if world > 1:  # multi-GPU correctness as part of the run: distributed == single GPU on the global mesh
        parity = distributed_parity(12, rank, world, dev, mat, halo=args.halo)
        if parity["hvp_rel_err"] > 1e-12 or parity["residual_rel_err"] > 1e-12:
            raise SystemExit(f"multi-GPU parity check failed: {parity}")

    if args.scaling == "strong":
        prob = DistributedHex8Problem(strong_block(), rank, world, dev, mat, variant=args.variant, halo=args.halo, hashed_mesh=True)
        workload = f"Hex8 {args.strong_n}^3 FIXED global mesh (config 4), neo-Hookean (mu=500, lambda=1000) matrix-free HVP, {prob.partition_desc}"
    else:
        prob = DistributedHex8Problem(args.n, rank, world, dev, mat, variant=args.variant, halo=args.halo)
        workload = f"Hex8 {args.n}^3 per GPU, neo-Hookean (mu=500, lambda=1000) matrix-free HVP, {prob.partition_desc}"
    n_dofs_global, partition_desc = prob.n_dofs_global, prob.partition_desc
    step = prob.step

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = _timed_steps(step, args.steps, barrier, torch)
    # kernel-only duration of the element kernel (per launch), events on the launching stream
    k_ms = prob.time_kernel_only(args.steps)
    # end-to-end through the public API with host buffers
    barrier()
    for _ in range(3):
        prob.step_e2e()
    prob.e2e_finish()
    n_e2e = max(3, min(args.steps, 20))

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
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
    h2d, d2h = prob.h2d_bytes, prob.d2h_bytes
    local_nodes, local_elems, launches = prob.local_nodes, prob.local_elems, prob.launches_per_step
    fp64_peak = prob.fp64_peak_tflops() if rank == 0 else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # ---- extra blocks: strong scaling of config 4 and the secondary configurations ---------------------------------
    strong = secondary = None
    if not args.no_secondary:
        del prob, step
        torch.cuda.empty_cache()
        if args.scaling == "weak":
            sp = DistributedHex8Problem(strong_block(), rank, world, dev, mat, variant=args.variant, halo=args.halo, hashed_mesh=True)
            for _ in range(5):
                sp.step()
            k_strong = 20
            ms_s = torch.tensor([_timed_steps(sp.step, k_strong, barrier, torch)], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
            ms_s = float(ms_s) / k_strong
            strong = {"workload": f"Hex8 {args.strong_n}^3 FIXED global mesh (config 4): {sp.partition_desc}", "scaling": "strong", "n_gpus": world,
                      "dofs_global": sp.n_dofs_global, "steps": k_strong, "ms_per_step": ms_s, "value": sp.n_dofs_global / (ms_s * 1e-3), "unit": "DOF/s"}
            del sp
            torch.cuda.empty_cache()
        secondary = secondary_c5(rank, world, dev, args.halo)
        if rank == 0:
            secondary.update(secondary_single_gpu(dev, hbm_peak))
        barrier()

    if rank == 0:
        traffic, traffic_src, ncu_pipe = None, None, None
        for name in ("r02_hvp_traffic.json", "r01_hvp_traffic.json"):  # DRAM bytes per launch from the committed ncu --set full capture (1 GPU, 128^3)
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                if local_elems == 128**3:
                    traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                    traffic_src = f"profiles/{name} (ncu --set full of the same kernel and size; not re-measured in this run)"
                    ncu_pipe = {k: tr[k] for k in ("fp64_pipe_active_pct_at_2_cycles_per_instruction", "fp64_instructions_per_element_sass", "instructions_per_thread", "issue_active_pct", "registers_per_thread", "kernel_us_under_ncu") if k in tr} or None
                break
            except (OSError, KeyError, ValueError):
                continue
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        alg_bytes = algorithmic_bytes(local_nodes, local_elems)
        hbm_ach = alg_bytes / (k_ms * 1e-3) / 1e9
        flops = FLOP_PER_ELEMENT * local_elems
        fp64_ach = flops / (k_ms * 1e-3) / 1e12
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp64_spec = 148 * 64 * 2 * sm_max * 1e6 / 1e12  # 148 SMs x 64 DFMA/clk x 2 flop x clock
        fp64_frac = fp64_ach / fp64_spec
        hbm_frac = hbm_ach / hbm_peak
        line = {
            "metric": METRIC,
            "value": n_dofs_global / (ms_step * 1e-3),
            "unit": "DOF/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload,
                "dofs_global": n_dofs_global,
                "dofs_per_gpu": 3 * local_nodes,
                "l2_policy": "inputs larger than L2 (273 MB working set per GPU vs 126 MB L2)",
                "parallelism": partition_desc,
            },
            # achieved = max(HBM term, FP64 term); the binding roof of this operator is FP64 (AI ~ 61 flop/B)
            "roofline": {
                "bound": "fp64" if fp64_frac >= hbm_frac else "hbm",
                "achieved": fp64_ach if fp64_frac >= hbm_frac else hbm_ach,
                "peak": fp64_spec if fp64_frac >= hbm_frac else hbm_peak,
                "unit": "TFLOP/s" if fp64_frac >= hbm_frac else "GB/s",
                "frac": max(fp64_frac, hbm_frac),
                "traffic": traffic,
                "traffic_source": traffic_src,
                "kernel": "k_hex8_nh_hvp",
                "kernel_ms": k_ms,
                "algorithmic_bytes": alg_bytes,
                "algorithmic_flops": flops,
                "flop_per_element_nominal": FLOP_PER_ELEMENT,
                "peak_source": f"FP64: spec-derived 148 SM x 64 DFMA/clk x 2 x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP64 entry); HBM: {peak_src}",
                "fp64_peak_tflops_dfma_microbenchmark_this_run": fp64_peak,
                "fp64_frac_of_microbenchmark": fp64_ach / fp64_peak if fp64_peak else None,
                "hbm_achieved_gbs": hbm_ach,
                "hbm_peak_gbs": hbm_peak,
                "hbm_frac": hbm_frac,
                "ncu_same_kernel": ncu_pipe,
                "note": "flops are the NOMINAL textbook count of SURVEY §8(d) (7944 per element); the kernel executes ~2400 FP64 instructions per element (modal, reference-space form), so frac measures time against the textbook-work roof and can exceed 1; pipe occupancy is in ncu_same_kernel (profiles/r02_hvp_ncu_full_summary.md)",
            },
            "e2e": {
                "value": n_dofs_global / (ms_e2e_step * 1e-3),
                "unit": "DOF/s",
                "ms_per_step": ms_e2e_step,
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "h2d_gbs_per_gpu": h2d / (ms_e2e_step * 1e-3) / 1e9,
                "d2h_gbs_per_gpu": d2h / (ms_e2e_step * 1e-3) / 1e9,
                "host_numa": numa,
                "note": "PCIe-bound: u and v up, y down every step through pinned buffers, three streams, two buffer sets",
            },
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
        if strong is not None:
            line["strong_scaling"] = strong
        if secondary is not None:
            line["secondary"] = secondary
        if not args.no_cpu_baseline:
            n_cpu = args.n if args.scaling == "weak" else args.strong_n
            line["cpu_baseline"] = cpu_port_baseline(n_cpu)
            line["cpu_baseline"].pop("ms_per_step", None)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
