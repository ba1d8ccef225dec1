"""Problem set-up for bench.py: the per-GPU Hex8 block, its device buffers and the timed step."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import tatva_b200
from tatva_b200 import _lib, element


class DistributedHex8Problem:
    def __init__(self, n, rank, world, device, material, variant=0):
        from bench import synthetic_inputs

        self.rank, self.world, self.device, self.material = rank, world, device, material
        if world > 1:
            raise NotImplementedError("multi-GPU decomposition is wired in bench_dist.py in a later commit")
        c, el, u, v = synthetic_inputs(n, rank)
        self.op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8(), device=device)
        if variant:
            self.op.set_variant(variant)
        self.local_nodes, self.local_elems = c.shape[0], el.shape[0]
        self.n_dofs_global = 3 * c.shape[0]
        self.partition_desc = "1 GPU, whole mesh"
        self.u = torch.as_tensor(u, device=device)
        self.v = torch.as_tensor(v, device=device)
        self.y = torch.empty_like(self.u)
        # host-side pinned buffers for the end-to-end leg
        self.h_u = torch.as_tensor(u).pin_memory()
        self.h_v = torch.as_tensor(v).pin_memory()
        self.h_y = torch.empty_like(self.h_u).pin_memory()
        self.h2d_bytes = self.h_u.numel() * 8 * 2
        self.d2h_bytes = self.h_y.numel() * 8
        self.launches_per_step = 1

    def step(self):
        self.op._raw_hvp(self.material, self.u, self.v, out=self.y)

    def step_e2e(self):
        """Public-API call with host buffers: H2D of u and v, HVP, D2H of y."""
        u = self.h_u.to(self.device, non_blocking=True)
        v = self.h_v.to(self.device, non_blocking=True)
        y = self.op.hvp(self.material)(u, v)
        self.h_y.copy_(y, non_blocking=True)

    def time_kernel_only(self, reps):
        """Average duration of one HVP call (memset + element kernel) on the launching stream."""
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
