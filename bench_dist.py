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


def numa_pin(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off BEFORE the pinned host buffers are allocated
    (first touch places them on that node): with one rank per GPU the end-to-end leg otherwise funnels every rank's
    H2D / D2H traffic through whichever socket the launcher started the process on.  Returns a small report."""
    import os

    info = {"numa_node": None, "cpus": None}
    try:
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id.lower()
    except Exception:  # noqa: BLE001
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis else device_index
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
            bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        except Exception:  # noqa: BLE001
            return info
    if len(bdf.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs uses 4
        bdf = bdf[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            cpus = set()
            for part in cpulist.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"] = len(cpus)
    except (OSError, ValueError):
        pass
    return info


class DistributedHex8Problem:
    def __init__(self, n, rank, world, device, material, variant=0, overlap=True, halo="nccl", grid=None, hashed_mesh=False):
        """`n`: cells per side of this rank's block (int) or a (nx, ny, nz) triple; `grid`: blocks per direction
        (default GRID[world]).  `hashed_mesh` builds the 1-GPU mesh with the block builder too (strong scaling: the
        same mesh family at every N)."""
        from bench import synthetic_inputs

        self.rank, self.world, self.device, self.material = rank, world, device, material
        if world == 1 and hashed_mesh:
            mesh, _ = structured_hex_block(n, (1, 1, 1), 0)
            c, el = mesh.coords, mesh.elements
            u, v = smooth_u(c), np.random.default_rng(1).normal(size=c.shape)
            self.op = tatva_b200.Operator(mesh, element.Hexahedron8(), device=device)
            self.pop = None
            self.n_dofs_global = 3 * c.shape[0]
            self.partition_desc = "1 GPU, whole mesh"
            n_owned = u.size
            self.launches_per_step = 2  # k_zero_release (clears y, releases its dependent at once) + the element kernel
        elif world == 1:
            c, el, u, v = synthetic_inputs(n, rank)
            self.op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8(), device=device)
            self.pop = None
            self.n_dofs_global = 3 * c.shape[0]
            self.partition_desc = "1 GPU, whole mesh"
            n_owned = u.size
            self.launches_per_step = 2  # k_zero_release (clears y, releases its dependent at once) + the element kernel
        else:
            grid = grid or GRID[world]
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
            shape = f"{n}^3" if np.isscalar(n) else "x".join(str(a) for a in n)
            self.partition_desc = f"{world} GPUs, {grid[0]}x{grid[1]}x{grid[2]} blocks of {shape}, {how} ({'overlapped' if overlap else 'serial'})"
            n_owned = self.pop.n_owned
            # own kernels per step: peer halo = pull, boundary + interior element kernels, push (replayed from one CUDA
            # graph); NCCL halo = pack, unpack-set, boundary + interior element kernels, pack, unpack-add
            self.launches_per_step = 4 if halo == "peer" else 6
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


# ------------------------------------------------------------------------------------------------------------------
# Multi-GPU correctness inside the bench run: distributed == single GPU on the global mesh (small n)
# ------------------------------------------------------------------------------------------------------------------
def distributed_parity(n, rank, world, device, material, halo="peer", grid=None):
    """The partitioned HVP / residual of a (gx n) x (gy n) x (gz n) box over `world` GPUs against the SAME box on this
    rank's GPU alone, compared on the owned rows in the plan's numbering (max-abs / max).  Every rank calls it; the
    returned errors are the maxima over ranks.  Also used by tests/test_gpu_multi.py."""
    import torch.distributed as dist

    from tatva_b200.distributed import _hash_uniform
    from tatva_b200.mesh import Mesh

    grid = grid or GRID[world]
    mesh, info = structured_hex_block(n, grid, rank)
    pop = PartitionedOperator(mesh, info, element.Hexahedron8(), material, device=device, overlap=True, halo=halo)
    l2g = info.nodes_local_to_global
    shape = (grid[0] * n, grid[1] * n, grid[2] * n)
    gm = Mesh.box_hex(shape)
    gc = gm.coords + 0.1 * (1.0 / max(shape)) * _hash_uniform(np.arange(gm.coords.shape[0]), 0)
    assert np.abs(gc[l2g] - mesh.coords).max() < 1e-14
    gop = tatva_b200.Operator(Mesh(coords=gc, elements=gm.elements), element.Hexahedron8(), device=device)
    gu, gv = smooth_u(gc), np.random.default_rng(7).normal(size=gc.shape)
    ref_hvp = gop.hvp(material)(gu, gv).cpu().numpy()
    ref_res = gop.residual(material)(gu).cpu().numpy()
    mk = pop.new_symmetric_vector if pop.halo == "peer" else pop.new_local_vector
    u_l, v_l, y_l, r_l = mk(), mk(), mk(), mk()
    u_l.copy_(torch.as_tensor(gu[l2g].ravel(), device=device))
    v_l.copy_(torch.as_tensor(gv[l2g].ravel(), device=device))
    v_l[pop.n_owned:] = 0.0  # the ghosts must come from the exchange
    y = pop.hvp(u_l, v_l, y_l)
    torch.cuda.synchronize()
    no = info.n_owned_nodes
    e_h = np.abs(y[: pop.n_owned].cpu().numpy().reshape(-1, 3) - ref_hvp[l2g[:no]]).max() / np.abs(ref_hvp).max()
    u2 = mk()
    u2.copy_(u_l)
    u2[pop.n_owned:] = 0.0
    r = pop.residual(u2, r_l)
    torch.cuda.synchronize()
    e_r = np.abs(r[: pop.n_owned].cpu().numpy().reshape(-1, 3) - ref_res[l2g[:no]]).max() / np.abs(ref_res).max()
    t = torch.tensor([e_h, e_r], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {
        "hvp_rel_err": float(t[0]), "residual_rel_err": float(t[1]), "tolerance": 1e-12,
        "what": f"distributed ({pop.halo} halo, overlapped) vs single-GPU on the global Hex8 {shape[0]}x{shape[1]}x{shape[2]} box, owned rows, max over ranks",
    }


# ------------------------------------------------------------------------------------------------------------------
# Secondary configurations (BASELINE.json configs 1, 2, 5) for the bench line's `secondary` block
# ------------------------------------------------------------------------------------------------------------------
def _timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _smooth(c):
    t = 2 * np.pi
    if c.shape[1] == 2:
        return 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 0])], -1)
    return smooth_u(c)


def secondary_single_gpu(device, hbm_peak_gbs):
    """Configs 1 and 2 on ONE GPU (rank 0): device-timed kernels, algorithmic bytes / time against the HBM peak."""
    from tatva_b200 import materials, sparse
    from tatva_b200.mesh import Mesh

    out = {}

    def rec(name, ms, alg_bytes, units, unit, **extra):
        out[name] = dict(ms=round(ms, 5), value=units / (ms * 1e-3), unit=unit + "/s", algorithmic_bytes=alg_bytes,
                         hbm_gbs=round(alg_bytes / ms / 1e6, 1), hbm_frac=round(alg_bytes / ms / 1e6 / hbm_peak_gbs, 4), **extra)

    with torch.cuda.device(device):
        # config 1: Tri3 256^2, plane-strain linear elasticity (E = 1, nu = 0.3): residual + matrix-free HVP
        m = Mesh.unit_square(256, 256)
        op = tatva_b200.Operator(m, element.Tri3(), device=device)
        mat = materials.LinearElastic.from_youngs_poisson_2d(1.0, 0.3)
        N, E = m.coords.shape[0], m.elements.shape[0]
        u = torch.as_tensor(_smooth(np.asarray(m.coords)), device=device)
        v = torch.as_tensor(np.random.default_rng(1).normal(size=m.coords.shape), device=device)
        y = torch.empty_like(u)
        note = "4.7 MB working set: L2-resident, launch-latency sized"
        rec("c1_tri3_le_hvp", _timeit(lambda: op._raw_hvp(mat, u, v, out=y), reps=200, warm=10), 8 * (2 * 2 * N + 2 * N) + 12 * E, 2 * N, "DOF", note=note)
        rec("c1_tri3_le_residual", _timeit(lambda: op._raw_residual(mat, u), reps=200, warm=10), 8 * (2 * 2 * N + 2 * N) + 12 * E, 2 * N, "DOF", note=note)
        del op
        # config 2: Tet4 box n = 55 (998 250 tets), neo-Hookean: residual + CSR assembly into the fixed pattern
        m = Mesh.box_tet((1.0, 1.0, 1.0), (55, 55, 55))
        c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / 55 * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
        m = Mesh(coords=c, elements=m.elements)
        op = tatva_b200.Operator(m, element.Tetrahedron4(), device=device)
        mat = materials.NeoHookean(500.0, 1000.0)
        N, E = c.shape[0], m.elements.shape[0]
        u = torch.as_tensor(smooth_u(c), device=device)
        v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device=device)
        y = torch.empty_like(u)
        how = "node-schedule kernel (k_fused_wc): per-warp distinct-node gather, per-tile node sums before the atomic adds" if op._node_schedule is not None else "element-per-thread kernel"
        op.set_variant(31)  # the element-per-thread kernel on the same plan
        ept_r, ept_h = _timeit(lambda: op._raw_residual(mat, u), reps=100, warm=10), _timeit(lambda: op._raw_hvp(mat, u, v, out=y), reps=100, warm=10)
        op.set_variant(0)
        rec("c2_tet4_nh_residual", _timeit(lambda: op._raw_residual(mat, u), reps=100, warm=10), 8 * (6 * N + 3 * N) + 16 * E, 3 * N, "DOF", fp64_floor_us=17.0, how=how, ms_element_per_thread=round(ept_r, 5))
        rec("c2_tet4_nh_hvp", _timeit(lambda: op._raw_hvp(mat, u, v, out=y), reps=100, warm=10), 8 * (9 * N + 3 * N) + 16 * E, 3 * N, "DOF", fp64_floor_us=17.0, how=how, ms_element_per_thread=round(ept_h, 5))
        pat = sparse.pattern_from_mesh(m, 3)
        cm = sparse.ColoredMatrix.from_csr(pat)
        asm = sparse.assembler(op, mat, cm)
        data = torch.empty(asm.nnz, dtype=torch.float64, device=device)
        rec("c2_tet4_nh_csr_assemble", _timeit(lambda: asm(u, out=data), reps=10), 8 * asm.nnz + 64 * E + 8 * 6 * N + 16 * E, asm.nnz, "nnz", nnz=asm.nnz, n_colors=int(np.asarray(cm.colors).max()) + 1)
        del asm, data, op
        # config 3 in context: one CG iteration around the HVP (Hex8 128^3, one Dirichlet face), CUDA graph
        from bench import synthetic_inputs
        from tatva_b200.lifter import Fixed, Lifter
        from tatva_b200.solver import ConjugateGradient, MaskedOperator, ReducedOperator

        c, el, u_, _ = synthetic_inputs(128)
        op = tatva_b200.Operator(Mesh(coords=c, elements=el), element.Hexahedron8(), device=device)
        fixed = np.where(c[:, 2] < 0.5 / 128)[0]
        lifter = Lifter(c.size, Fixed((fixed[:, None] * 3 + np.arange(3)).ravel()))
        n = lifter.size_reduced
        u_red = lifter.reduce(torch.as_tensor(0.02 * u_.ravel(), device=device))  # small strains: SPD tangent
        b = torch.as_tensor(np.random.default_rng(3).normal(size=n), device=device)

        def cg_ms(cg, rhs):
            cg.solve(rhs, tol=0.0, maxiter=20, check_every=20)  # warm-up + graph capture
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            x, info = cg.solve(rhs, tol=0.0, maxiter=100, check_every=50)
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / info["iterations"], x

        mo = MaskedOperator(op, mat, lifter)
        mo.set_state(u_red)
        ms_m, x_m = cg_ms(mo.solver(use_graph=True), mo.expand(b))
        red = ReducedOperator(op, mat, lifter)
        red.set_state(u_red)
        ms_r, x_r = cg_ms(ConjugateGradient(red.matvec, n, device, use_graph=True), b)
        out["c3_cg_iteration_hex8_128"] = dict(ms=round(ms_m, 5), value=n / (ms_m * 1e-3), unit="DOF/s", how="CUDA graph; full-size vectors around the unconstrained HVP kernel, Dirichlet rows masked in the update pass (MaskedOperator)",
                                               ms_reduced_space_lifted_kernel=round(ms_r, 5), iterate_rel_diff_after_100=float((mo.restrict(x_m) - x_r).norm() / x_r.norm()))
        diag = torch.empty(n, dtype=torch.float64, device=device)
        rec("c3_hessian_diagonal_hex8_128", _timeit(lambda: red.diagonal(out=diag), reps=5, warm=2), 8 * 9 * c.shape[0] + 32 * el.shape[0], n, "DOF",
            how="rank-structured diagonal (k_hessian_diag_rank): K_aa(i,i) = w1 |dN_a|^2 + (w2 + w3) g_a[i]^2; r01: 24 tangent evaluations per point, 1.72 ms")
        del mo, red, diag
        # user-supplied densities at config 3 (README.md:93: the density is user code): written once on symbols, compiled
        # at run time into the fused kernel template; beside them the built-in law through the same generic template and
        # the r01 route for a density without a kernel (autograd through the Operator building blocks)
        try:
            import sympy as sp

            def psi_nh(G, mu, lam):
                F = sp.eye(3) + G
                lnJ = sp.log(F.det())
                return mu / 2 * ((F.T * F).trace() - 3 - 2 * lnJ) + lam / 2 * lnJ**2

            def psi_mr(G, c1, c2, kappa):
                F = sp.eye(3) + G
                Cm = F.T * F
                J = F.det()
                I1 = Cm.trace()
                I2 = (I1**2 - (Cm * Cm).trace()) / 2
                return c1 * (J ** sp.Rational(-2, 3) * I1 - 3) + c2 * (J ** sp.Rational(-4, 3) * I2 - 3) + kappa / 2 * (J - 1) ** 2

            N, E = c.shape[0], el.shape[0]
            u = torch.as_tensor(u_, device=device)
            v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device=device)
            y = torch.empty_like(u)
            nb = 8 * 12 * N + 32 * E
            op.set_variant(1)
            rec("c3_hvp_generic_template_builtin_neo_hookean", _timeit(lambda: op._raw_hvp(mat, u, v, out=y), reps=10), nb, 3 * N, "DOF")
            ref = y.clone()
            op.set_variant(0)
            law_nh = materials.UserLaw.from_psi(psi_nh, (500.0, 1000.0))
            law_mr = materials.UserLaw.from_psi(psi_mr, (120.0, 30.0, 900.0))
            ms = _timeit(lambda: op._raw_hvp(law_nh, u, v, out=y), reps=10)
            rec("c3_hvp_user_law_neo_hookean", ms, nb, 3 * N, "DOF", rel_err_vs_builtin=float((y - ref).norm() / ref.norm()), ops=law_nh.generated.op_counts(),
                how="density written once on symbols -> source-to-source AD (lawgen) -> NVRTC -> k_fused<Hex8, UserLaw, HVP>")
            rec("c3_hvp_user_law_mooney_rivlin", _timeit(lambda: op._raw_hvp(law_mr, u, v, out=y), reps=10), nb, 3 * N, "DOF", ops=law_mr.generated.op_counts())

            def hv_autograd():
                uu = u.detach().requires_grad_(True)
                F = op.grad(uu) + torch.eye(3, dtype=torch.float64, device=device)
                lnJ = torch.log(torch.linalg.det(F))
                psi = 250.0 * ((F * F).sum((-1, -2)) - 3 - 2 * lnJ) + 500.0 * lnJ * lnJ
                (g,) = torch.autograd.grad(op.integrate(psi), uu, create_graph=True)
                (hv,) = torch.autograd.grad(g, uu, grad_outputs=v)
                return hv

            hv = hv_autograd()
            rec("c3_hvp_autograd_route_neo_hookean", _timeit(hv_autograd, reps=3, warm=1), nb, 3 * N, "DOF", rel_err_vs_builtin=float((hv - ref).norm() / ref.norm()),
                how="torch double backward through op.grad / op.integrate with (E, Q, 3, 3) temporaries in HBM: the r01 route for a density without a kernel")
            del hv
        except ImportError:
            out["c3_user_law"] = "sympy not available"
        del op
        torch.cuda.empty_cache()
    return out


def secondary_c5(rank, world, device, halo, n=55):
    """Config 5: compound (u, phi) Tet4 phase-field operator, one n^3-cell block (6 n^3 tets, 4 DOFs per node) per GPU
    (8 GPUs = the n = 110 box of SURVEY §8), coupled residual + HVP with the halo exchange.  All ranks call it."""
    import torch.distributed as dist

    from tatva_b200 import materials
    from tatva_b200.distributed import structured_tet_block

    mesh, info = structured_tet_block(n, GRID[world], rank)
    mat = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.05, 1e-6)
    pop = PartitionedOperator(mesh, info, element.Tetrahedron4(), mat, device=device, overlap=True, halo=halo)
    c = np.asarray(mesh.coords)
    s0 = np.concatenate([0.4 * smooth_u(c), 0.5 + 0.3 * np.sin(6 * c[:, :1])], axis=1)
    mk = pop.new_symmetric_vector if (pop.halo == "peer" and world > 1) else pop.new_local_vector
    s, d, y = mk(), mk(), mk()
    s.copy_(torch.as_tensor(s0.ravel(), device=device))
    d.copy_(torch.as_tensor(np.random.default_rng(1 + rank).normal(size=s0.shape).ravel(), device=device))
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
        ms = torch.tensor([a.elapsed_time(b) / reps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    ms_h = timed(lambda: pop.hvp(s, d, y))
    ms_r = timed(lambda: pop.residual(s, y))
    return {
        "c5_compound_tet4_pf": dict(n_gpus=world, tets_per_gpu=6 * n**3, dofs_global=int(pop.n_global), halo=pop.halo if world > 1 else None,
                                    hvp_ms=round(ms_h, 5), hvp_value=pop.n_global / (ms_h * 1e-3), residual_ms=round(ms_r, 5),
                                    residual_value=pop.n_global / (ms_r * 1e-3), unit="DOF/s", scaling="weak (one 55^3-cell block per GPU)")
    }
