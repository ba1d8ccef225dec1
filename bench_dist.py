"""Problem set-up for bench.py: the per-GPU Hex8 block, its device buffers and the timed step.

N = 1: the whole 128^3 mesh on one GPU.  N = 2/4/8: weak scaling, one n^3 block per GPU of a
(2n x n x n) / (2n x 2n x n) / (2n)^3 box (config 4 at N = 8, n = 128), partitioned with the reference's
ownership rules and exchanged through tatva_b200.distributed.PartitionedOperator (NCCL over NVLink).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import tatva_b200
from tatva_b200 import _lib, element
from tatva_b200.distributed import PartitionedOperator, structured_hex_block

GRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def smooth_u(c):
    two_pi = 2 * np.pi
    return 0.05 * np.stack(
        [np.sin(two_pi * c[:, 0]) * np.cos(two_pi * c[:, 1]), np.sin(two_pi * c[:, 1]) * np.cos(two_pi * c[:, 2]), np.sin(two_pi * c[:, 2]) * np.cos(two_pi * c[:, 0])], -1
    )


class DistributedHex8Problem:
    def __init__(self, n, rank, world, device, material, variant=0, overlap=True, halo="nccl"):
        from bench import synthetic_inputs

        self.rank, self.world, self.device, self.material = rank, world, device, material
        if world == 1:
            c, el, u, v = synthetic_inputs(n, rank)
            self.op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8(), device=device)
            self.pop = None
            self.n_dofs_global = 3 * c.shape[0]
            self.partition_desc = "1 GPU, whole mesh"
            n_owned = u.size
            self.launches_per_step = 1
        else:
            grid = GRID[world]
            mesh, info = structured_hex_block(n, grid, rank)
            c, el = mesh.coords, mesh.elements
            self.pop = PartitionedOperator(mesh, info, element.Hexahedron8(), material, device=device, overlap=overlap, halo=halo)
            if halo == "peer":
                # all ranks must agree that peer-mapped memory works; otherwise everybody uses the NCCL path
                import torch.distributed as dist

                ok = torch.ones(1, device=device)
                try:
                    probe = self.pop.new_symmetric_vector()
                    self.pop._peer_pull(probe)
                    torch.cuda.synchronize()
                except Exception as exc:  # noqa: BLE001
                    ok.zero_()
                    self._peer_error = repr(exc)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok) == 0.0:
                    halo = "nccl"
                    self.pop = PartitionedOperator(mesh, info, element.Hexahedron8(), material, device=device, overlap=overlap, halo="nccl")
            self.op = self.pop.op
            u = smooth_u(c)  # ghost values consistent by construction (function of the shared coordinates)
            v = np.random.default_rng(1 + rank).normal(size=c.shape)
            self.n_dofs_global = self.pop.n_global
            how = "peer-memory halo over NVLink (own kernels + device barrier)" if halo == "peer" else "NCCL halo exchange"
            self.partition_desc = f"{world} GPUs, {grid[0]}x{grid[1]}x{grid[2]} blocks of {n}^3, {how} ({'overlapped' if overlap else 'serial'})"
            n_owned = self.pop.n_owned
            self.launches_per_step = 6  # pack, unpack-set, boundary + interior element kernels, pack, unpack-add
        if variant:
            self.op.set_variant(variant)
        self.local_nodes, self.local_elems = c.shape[0], el.shape[0]
        if self.pop is not None and halo == "peer":
            self.u, self.v, self.y = (self.pop.new_symmetric_vector() for _ in range(3))
            self.u.copy_(torch.as_tensor(u, device=device).reshape(-1))
            self.v.copy_(torch.as_tensor(v, device=device).reshape(-1))
        else:
            self.u = torch.as_tensor(u, device=device).reshape(-1)
            self.v = torch.as_tensor(v, device=device).reshape(-1)
            self.y = torch.empty_like(self.u)
        self.n_owned = n_owned
        # host-side pinned buffers for the end-to-end leg (owned entries only)
        self.h_u = torch.as_tensor(u).reshape(-1)[:n_owned].clone().pin_memory()
        self.h_v = torch.as_tensor(v).reshape(-1)[:n_owned].clone().pin_memory()
        self.h_y = torch.empty(n_owned, dtype=torch.float64).pin_memory()
        self.h2d_bytes = n_owned * 8 * 2
        self.d2h_bytes = n_owned * 8
        if self.pop is not None:
            self.pop.fill_ghosts(self.u)
            self.pop.fill_ghosts(self.v)

    def step(self):
        if self.pop is None:
            self.op._raw_hvp(self.material, self.u, self.v, out=self.y)
        else:
            self.pop.hvp(self.u, self.v, self.y)

    def step_e2e(self):
        """Public-API call with HOST buffers, every step: H2D of the owned u and v (pinned), (halo exchange +)
        HVP, D2H of the owned y.  Three streams and two device buffer sets pipeline consecutive steps: the
        upload of step i+1 and the download of step i-1 overlap the kernel of step i (PCIe is full duplex)."""
        if not hasattr(self, "_pipe"):
            mk = lambda: torch.cuda.Stream(device=self.device)  # noqa: E731
            self._pipe = dict(s_in=mk(), s_out=mk(), i=0, sets=[
                dict(u=self._clone_vec(self.u), v=self._clone_vec(self.v), y=self._clone_vec(self.u), in_done=torch.cuda.Event(), comp_done=torch.cuda.Event(), out_done=torch.cuda.Event())
                for _ in range(2)
            ])
        P = self._pipe
        B = P["sets"][P["i"] % 2]
        P["i"] += 1
        n = self.n_owned
        main = torch.cuda.current_stream(self.device)
        P["s_in"].wait_event(B["comp_done"])  # buffer set free again (no-op the first time round)
        with torch.cuda.stream(P["s_in"]):
            B["u"][:n].copy_(self.h_u, non_blocking=True)
            B["v"][:n].copy_(self.h_v, non_blocking=True)
            B["in_done"].record()
        main.wait_event(B["in_done"])
        main.wait_event(B["out_done"])  # previous download of this set's y finished
        if self.pop is None:
            y = self.op.hvp(self.material)(B["u"].view(-1, 3), B["v"].view(-1, 3)).reshape(-1)
            B["y_ref"] = y  # keep alive until downloaded
        else:
            self.pop.fill_ghosts(B["u"])
            y = self.pop.hvp(B["u"], B["v"], B["y"])
        B["comp_done"].record(main)
        P["s_out"].wait_event(B["comp_done"])
        with torch.cuda.stream(P["s_out"]):
            self.h_y.copy_(y[:n], non_blocking=True)
            B["out_done"].record()

    def _clone_vec(self, x):
        if self.pop is not None and self.pop.halo == "peer":
            y = self.pop.new_symmetric_vector()
            y.copy_(x)
            return y
        return x.clone()

    def e2e_finish(self):
        """Join the download stream into the current stream (so a following event covers the last D2H)."""
        if hasattr(self, "_pipe"):
            torch.cuda.current_stream(self.device).wait_stream(self._pipe["s_out"])

    def time_kernel_only(self, reps):
        """Average duration of the element kernel over the rank's whole local mesh (memset included),
        CUDA events on the launching stream, no exchange."""
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            self.op._raw_hvp(self.material, self.u, self.v, out=self.y)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def fp64_peak_tflops(self):
        out = C.c_double()
        _lib.check(_lib.lib().tatva_fp64_peak_tflops(C.byref(out), torch.cuda.current_stream().cuda_stream))
        return out.value
