"""Secondary kernels of SURVEY.md §8: timings at the configs' sizes (1 GPU).  One JSON line per kernel."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials, sparse
from tatva_b200.mesh import Mesh

HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None


def timeit(fn, reps=20, warm=3):
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


def smooth_u(c):
    t = 2 * np.pi
    if c.shape[1] == 2:
        return 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 0])], -1)
    return 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 2]), np.sin(t * c[:, 2]) * np.cos(t * c[:, 0])], -1)


def report(name, ms, alg_bytes, units, unit_name, **extra):
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), rate=units / ms * 1e3, unit=unit_name + "/s", alg_GBs=round(alg_bytes / ms / 1e6, 1), hbm_frac=round(alg_bytes / ms / 1e6 / HBM, 4), **extra)), flush=True)


def want(tag):
    return only is None or tag in only


# ---- config 2: Tet4 box n=55 (998 250 elements), neo-Hookean ----
if want("tet4"):
    m = Mesh.box_tet((1.0, 1.0, 1.0), (55, 55, 55))
    c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / 55 * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
    m = Mesh(coords=c, elements=m.elements)
    op = tatva_b200.Operator(m, element.Tetrahedron4())
    mat = materials.NeoHookean(500.0, 1000.0)
    N, E = c.shape[0], m.elements.shape[0]
    u = torch.as_tensor(smooth_u(c), device="cuda")
    v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device="cuda")
    y = torch.empty_like(u)
    report("tet4_nh_hvp_c2", timeit(lambda: op._raw_hvp(mat, u, v, out=y)), 8 * (9 * N + 3 * N) + 16 * E, 3 * N, "DOF", elems=E)
    report("tet4_nh_residual_c2", timeit(lambda: op._raw_residual(mat, u)), 8 * (6 * N + 3 * N) + 16 * E, 3 * N, "DOF")
    opt = tatva_b200.Operator(m, element.Tetrahedron4(), stage_tiles=True)
    report("tet4_nh_hvp_c2_smem_tiles", timeit(lambda: opt._raw_hvp(mat, u, v, out=y)), 8 * (9 * N + 3 * N) + 16 * E, 3 * N, "DOF", max_unique_per_tile=opt._tiles[3])
    report("tet4_nh_residual_c2_smem_tiles", timeit(lambda: opt._raw_residual(mat, u)), 8 * (6 * N + 3 * N) + 16 * E, 3 * N, "DOF")
    del opt
    report("tet4_nh_energy_c2", timeit(lambda: op._raw_energy(mat, u)), 8 * (3 * N + 3 * N) + 16 * E, 3 * N, "DOF")
    t0 = time.perf_counter()
    pat = sparse.pattern_from_mesh(m, 3)
    t1 = time.perf_counter()
    cm = sparse.ColoredMatrix.from_csr(pat)
    t2 = time.perf_counter()
    asm = sparse.assembler(op, mat, cm)
    t3 = time.perf_counter()
    data = torch.empty(asm.nnz, dtype=torch.float64, device="cuda")
    nnz = asm.nnz
    report("tet4_nh_csr_assemble_c2", timeit(lambda: asm(u, out=data), reps=10), 8 * nnz + 64 * E + 8 * 6 * N + 16 * E, nnz, "nnz", nnz=nnz, n_colors=int(cm.colors.max()) + 1,
           host_pattern_s=round(t1 - t0, 3), host_colouring_s=round(t2 - t1, 3), host_positions_s=round(t3 - t2, 3))
    asm3 = sparse.assembler(op, mat, cm, symmetric=True)
    data3 = torch.empty(nnz, dtype=torch.float64, device="cuda")
    report("tet4_nh_csr_assemble_symmetric_mirror_c2", timeit(lambda: asm3(u, out=data3), reps=10), 8 * nnz + 64 * E + 8 * 6 * N + 16 * E, nnz, "nnz", rel_diff=float((data - data3).norm() / data3.norm()))
    del data3, asm3
    asm2 = sparse.assembler(op, mat, cm, by_rows=True)
    data2 = torch.empty(nnz, dtype=torch.float64, device="cuda")
    report("tet4_nh_csr_assemble_rows_deterministic_c2", timeit(lambda: asm2(u, out=data2), reps=10), 8 * nnz + 64 * E + 8 * 6 * N + 16 * E, nnz, "nnz", rel_diff=float((data - data2).norm() / data2.norm()))
    del data2, asm2
    # the reference algorithm costs n_colors HVPs: report what that would be with our HVP kernel
    del data, asm

# ---- config 5: Tet4 n=55 compound (u, phi) ----
if want("pf"):
    m = Mesh.box_tet((1.0, 1.0, 1.0), (55, 55, 55))
    c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / 55 * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
    m = Mesh(coords=c, elements=m.elements)
    op = tatva_b200.Operator(m, element.Tetrahedron4())
    mat = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.05, 1e-6)
    N, E = c.shape[0], m.elements.shape[0]
    s = torch.as_tensor(np.concatenate([smooth_u(c), 0.5 + 0.3 * np.sin(6 * c[:, :1])], axis=1), device="cuda")
    t = torch.as_tensor(np.random.default_rng(1).normal(size=(N, 4)), device="cuda")
    y = torch.empty_like(s)
    report("tet4_pf_hvp_c5", timeit(lambda: op._raw_hvp(mat, s, t, out=y)), 8 * (12 * N + 3 * N) + 16 * E, 4 * N, "DOF")
    report("tet4_pf_residual_c5", timeit(lambda: op._raw_residual(mat, s)), 8 * (8 * N + 3 * N) + 16 * E, 4 * N, "DOF")

# ---- config 1: Tri3 256^2, linear elasticity ----
if want("tri3"):
    m = Mesh.unit_square(256, 256)
    op = tatva_b200.Operator(m, element.Tri3())
    mat = materials.LinearElastic.from_youngs_poisson_2d(1.0, 0.3)
    N, E = m.coords.shape[0], m.elements.shape[0]
    u = torch.as_tensor(smooth_u(m.coords), device="cuda")
    v = torch.as_tensor(np.random.default_rng(1).normal(size=m.coords.shape), device="cuda")
    y = torch.empty_like(u)
    report("tri3_le_hvp_c1", timeit(lambda: op._raw_hvp(mat, u, v, out=y), reps=100), 8 * (2 * 2 * N + 2 * N) + 12 * E, 2 * N, "DOF")
    report("tri3_le_residual_c1", timeit(lambda: op._raw_residual(mat, u), reps=100), 8 * (2 * 2 * N + 2 * N) + 12 * E, 2 * N, "DOF")

# ---- config 3 extras: Hex8 128^3 residual / energy, building blocks at 64^3 ----
if want("hex8"):
    from bench import synthetic_inputs
    c, el, u_, v_ = synthetic_inputs(128)
    op = tatva_b200.Operator(Mesh(coords=c, elements=el), element.Hexahedron8())
    mat = materials.NeoHookean(500.0, 1000.0)
    N, E = c.shape[0], el.shape[0]
    u = torch.as_tensor(u_, device="cuda")
    report("hex8_nh_residual_c3", timeit(lambda: op._raw_residual(mat, u)), 8 * (9 * N) + 32 * E, 3 * N, "DOF")
    report("hex8_nh_energy_c3", timeit(lambda: op._raw_energy(mat, u)), 8 * (6 * N) + 32 * E, 3 * N, "DOF")
    del op
    for nb in (64, 128):
        c, el, u_, v_ = synthetic_inputs(nb)
        op = tatva_b200.Operator(Mesh(coords=c, elements=el), element.Hexahedron8())
        N, E = c.shape[0], el.shape[0]
        u = torch.as_tensor(u_, device="cuda")
        g = op._k_grad(u)
        report(f"hex8_op_grad_{nb}", timeit(lambda: op._k_grad(u)), 8 * (6 * N + 72 * E) + 32 * E, E * 8, "qp")
        report(f"hex8_op_grad_adjoint_{nb}", timeit(lambda: op._k_grad_adj(g)), 8 * (6 * N + 72 * E) + 32 * E, E * 8, "qp")
        report(f"hex8_op_eval_{nb}", timeit(lambda: op._k_eval(u)), 8 * (3 * N + 24 * E) + 32 * E, E * 8, "qp")
        report(f"hex8_op_weights_{nb}", timeit(lambda: op.get_integration_weights()), 8 * (3 * N + 8 * E) + 32 * E, E * 8, "qp")
        report(f"hex8_op_gather_{nb}", timeit(lambda: op._k_gather(u)), 8 * (3 * N + 24 * E) + 32 * E, E * 8, "node-ref")
        del op, g

# ---- config 3 in context: one CG iteration around the HVP (Hex8 128^3, Dirichlet face), CUDA graph on/off ----
if want("cg"):
    from bench import synthetic_inputs
    from tatva_b200.lifter import Fixed, Lifter
    from tatva_b200.solver import ConjugateGradient, ReducedOperator
    c, el, u_, v_ = synthetic_inputs(128)
    op = tatva_b200.Operator(Mesh(coords=c, elements=el), element.Hexahedron8())
    mat = materials.NeoHookean(500.0, 1000.0)
    fixed = np.where(c[:, 2] < 0.5 / 128)[0]
    lifter = Lifter(c.size, Fixed((fixed[:, None] * 3 + np.arange(3)).ravel()))
    red = ReducedOperator(op, mat, lifter)
    red.set_state(lifter.reduce(torch.as_tensor(0.02 * u_.ravel(), device="cuda")))  # small strains: SPD tangent
    n = lifter.size_reduced
    b = torch.as_tensor(np.random.default_rng(3).normal(size=n), device="cuda")
    red_plain = ReducedOperator(op, mat, lifter, fused=False)  # lift kernel -> HVP -> segmented-sum kernel
    red_plain.set_state(lifter.reduce(torch.as_tensor(0.02 * u_.ravel(), device="cuda")))
    vr, yr = b.clone(), torch.empty_like(b)
    report("hvp_lifted_fused_hex8_128", timeit(lambda: red.matvec(vr, yr)), 8 * (9 * c.shape[0]) + 32 * el.shape[0] + 8 * 2 * n + 4 * 3 * c.shape[0], n, "DOF")
    report("hvp_lifted_3_kernels_hex8_128", timeit(lambda: red_plain.matvec(vr, yr)), 8 * (9 * c.shape[0]) + 32 * el.shape[0] + 8 * 2 * n + 4 * 3 * c.shape[0], n, "DOF")
    for graph, r_ in ((True, red_plain), (False, red), (True, red)):
        tag = "" if r_ is red else "_unfused_lifter"
        cg = ConjugateGradient(r_.matvec, n, "cuda", use_graph=graph)
        cg.solve(b, tol=0.0, maxiter=20, check_every=20)  # warm-up + capture
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        x, info = cg.solve(b, tol=0.0, maxiter=200, check_every=50)
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / info["iterations"]
        report(f"cg_iteration_hex8_128_{'graph' if graph else 'eager'}{tag}", ms, 8 * (12 * c.shape[0]) + 32 * el.shape[0] + 8 * 11 * n, n, "DOF", iterations=info["iterations"], residual_norm=info["residual_norm"])

    # r02: p.Ap summed inside the HVP kernel, direction pass clears Ap (no memset), 6 launches per iteration
    x_ref = x.clone()
    cgf = ConjugateGradient(red.matvec, n, "cuda", use_graph=True, matvec_dot=red.matvec_dot)
    cgf.solve(b, tol=0.0, maxiter=20, check_every=20)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    xf, info = cgf.solve(b, tol=0.0, maxiter=200, check_every=50)
    a1.record()
    torch.cuda.synchronize()
    report("cg_iteration_hex8_128_graph_fused_dot", a0.elapsed_time(a1) / info["iterations"], 8 * (12 * c.shape[0]) + 32 * el.shape[0] + 8 * 9 * n, n, "DOF", iterations=info["iterations"],
           residual_norm=info["residual_norm"], iterate_rel_diff_vs_unfused_after_200=float((xf - x_ref).norm() / x_ref.norm()))
    pcgf = ConjugateGradient(red.matvec, n, "cuda", use_graph=True, jacobi=True, matvec_dot=red.matvec_dot)
    # r02: Fixed-only lifter -> full-size vectors around the UNCONSTRAINED kernel, Dirichlet rows masked in the update pass
    from tatva_b200.solver import MaskedOperator
    for fuse in (False, True):
        mo = MaskedOperator(op, mat, lifter, fuse_dot=fuse)
        mo.set_state(lifter.reduce(torch.as_tensor(0.02 * u_.ravel(), device="cuda")))
        cgm = mo.solver(use_graph=True)
        bf = mo.expand(b)
        cgm.solve(bf, tol=0.0, maxiter=20, check_every=20)
        torch.cuda.synchronize()
        a0.record()
        xm, info = cgm.solve(bf, tol=0.0, maxiter=200, check_every=50)
        a1.record()
        torch.cuda.synchronize()
        report(f"cg_iteration_hex8_128_graph_masked_full_space{'_kernel_dot' if fuse else ''}", a0.elapsed_time(a1) / info["iterations"], 8 * (12 * c.shape[0]) + 32 * el.shape[0] + 8 * (11 - 2 * fuse) * c.size + 4 * c.size, n, "DOF",
               iterations=info["iterations"], iterate_rel_diff_vs_reduced_after_200=float((mo.restrict(xm) - x_ref).norm() / x_ref.norm()))
        del cgm, mo

    # Jacobi: cost of the diagonal kernel, of one preconditioned iteration, and iterations to 1e-8 with / without
    diag = torch.empty(n, dtype=torch.float64, device="cuda")
    report("hex8_nh_hessian_diag_c3", timeit(lambda: red.diagonal(out=diag), reps=5, warm=2), 8 * 9 * c.shape[0] + 32 * el.shape[0], n, "DOF")
    pcg = ConjugateGradient(red.matvec, n, "cuda", use_graph=True, jacobi=True)
    pcg.set_diagonal(diag)
    pcg.solve(b, tol=0.0, maxiter=20, check_every=20)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    x, info = pcg.solve(b, tol=0.0, maxiter=200, check_every=50)
    a1.record()
    torch.cuda.synchronize()
    report("pcg_jacobi_iteration_hex8_128_graph", a0.elapsed_time(a1) / info["iterations"], 8 * (12 * c.shape[0]) + 32 * el.shape[0] + 8 * 13 * n, n, "DOF", iterations=info["iterations"])
    pcgf.set_diagonal(diag)
    pcgf.solve(b, tol=0.0, maxiter=20, check_every=20)
    torch.cuda.synchronize()
    a0.record()
    x, info = pcgf.solve(b, tol=0.0, maxiter=200, check_every=50)
    a1.record()
    torch.cuda.synchronize()
    report("pcg_jacobi_iteration_hex8_128_graph_fused_dot", a0.elapsed_time(a1) / info["iterations"], 8 * (12 * c.shape[0]) + 32 * el.shape[0] + 8 * 11 * n, n, "DOF", iterations=info["iterations"])
    its = {}
    for name, solver in (("cg", cg), ("pcg_jacobi", pcg), ("cg_fused_dot", cgf)):
        t0 = time.perf_counter()
        x, info = solver.solve(b, tol=1e-8, maxiter=5000, check_every=25)
        torch.cuda.synchronize()
        its[name] = dict(iterations=info["iterations"], converged=info["converged"], wall_s=round(time.perf_counter() - t0, 3))
    print(json.dumps(dict(kernel="cg_vs_pcg_to_1e-8_hex8_128", **its)), flush=True)

# ---- user-supplied densities at config 3: run-time compiled fused kernels vs the generic template vs the autograd route ----
if want("userlaw"):
    import sympy as sp
    from bench import synthetic_inputs
    c, el, u_, v_ = synthetic_inputs(128)
    op = tatva_b200.Operator(Mesh(coords=c, elements=el), element.Hexahedron8())
    N, E = c.shape[0], el.shape[0]
    u, v = torch.as_tensor(u_, device="cuda"), torch.as_tensor(v_, device="cuda")
    y = torch.empty_like(u)
    nh = materials.NeoHookean(500.0, 1000.0)

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

    bytes_hvp = 8 * (12 * N) + 32 * E
    op.set_variant(1)
    report("hex8_nh_hvp_c3_generic_template_builtin_law", timeit(lambda: op._raw_hvp(nh, u, v, out=y)), bytes_hvp, 3 * N, "DOF")
    ref = y.clone()
    op.set_variant(0)
    t0 = time.perf_counter()
    law_nh = materials.UserLaw.from_psi(psi_nh, (500.0, 1000.0))
    law_mr = materials.UserLaw.from_psi(psi_mr, (120.0, 30.0, 900.0))
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    op._raw_hvp(law_nh, u, v, out=y)
    torch.cuda.synchronize()
    t_jit = time.perf_counter() - t0
    report("hex8_user_law_neo_hookean_hvp_c3", timeit(lambda: op._raw_hvp(law_nh, u, v, out=y)), bytes_hvp, 3 * N, "DOF", rel_err_vs_builtin=float((y - ref).norm() / ref.norm()),
           ops=law_nh.generated.op_counts(), codegen_s=round(t_gen, 2), nvrtc_first_call_s=round(t_jit, 2))
    report("hex8_user_law_neo_hookean_residual_c3", timeit(lambda: op._raw_residual(law_nh, u)), 8 * (9 * N) + 32 * E, 3 * N, "DOF")
    report("hex8_user_law_mooney_rivlin_hvp_c3", timeit(lambda: op._raw_hvp(law_mr, u, v, out=y)), bytes_hvp, 3 * N, "DOF", ops=law_mr.generated.op_counts())
    report("hex8_user_law_mooney_rivlin_residual_c3", timeit(lambda: op._raw_residual(law_mr, u)), 8 * (9 * N) + 32 * E, 3 * N, "DOF")
    # the r01 route for a density without a kernel: autograd (double backward) through the Operator building blocks,
    # (E, Q, 3, 3) temporaries in HBM
    def hv_autograd():
        uu = u.detach().requires_grad_(True)
        F = op.grad(uu) + torch.eye(3, dtype=torch.float64, device="cuda")
        lnJ = torch.log(torch.linalg.det(F))
        psi = 250.0 * ((F * F).sum((-1, -2)) - 3 - 2 * lnJ) + 500.0 * lnJ * lnJ
        (g,) = torch.autograd.grad(op.integrate(psi), uu, create_graph=True)
        (hv,) = torch.autograd.grad(g, uu, grad_outputs=v)
        return hv
    hv = hv_autograd()
    report("hex8_neo_hookean_hvp_c3_autograd_route", timeit(hv_autograd, reps=5, warm=2), bytes_hvp, 3 * N, "DOF", rel_err_vs_builtin=float((hv - ref).norm() / ref.norm()),
           peak_mem_GB=round(torch.cuda.max_memory_allocated() / 1e9, 2))
    del op

# ---- post-processing and boundary elements: interpolate (point location), project (mass CG), Line2 traction ----
if want("post"):
    rng = np.random.default_rng(0)
    m = Mesh.unit_square(256, 256)
    op = tatva_b200.Operator(m, element.Tri3())
    P = 100_000
    pts = torch.as_tensor(rng.uniform(0.001, 0.999, size=(P, 2)), device="cuda")
    un = torch.as_tensor(rng.normal(size=(m.coords.shape[0], 2)), device="cuda")
    t0 = time.perf_counter()
    op.interpolate(un, pts)
    torch.cuda.synchronize()
    ms = timeit(lambda: op.interpolate(un, pts), reps=5, warm=1)
    report("tri3_interpolate_100k_points_c1", ms, 8 * (4 * P) + 8 * 4 * m.coords.shape[0] + 12 * m.elements.shape[0], P, "point", elements=int(m.elements.shape[0]))
    mq = Mesh.unit_square(256, 256, type="quad")
    cq = np.asarray(mq.coords)
    opq = tatva_b200.Operator(mq, element.Quad4())
    qp = opq.quads()
    fq = torch.sin(3 * qp[:, :, 0]) * torch.exp(qp[:, :, 1])
    t0 = time.perf_counter()
    xq = opq.project(fq)
    torch.cuda.synchronize()
    t_proj = time.perf_counter() - t0
    print(json.dumps(dict(kernel="quad4_project_scalar_256x256", wall_ms=round(1e3 * t_proj, 2), nodes=int(cq.shape[0]), max_abs_err_vs_field=float((xq - torch.sin(3 * torch.as_tensor(cq[:, 0], device="cuda")) * torch.exp(torch.as_tensor(cq[:, 1], device="cuda"))).abs().max()))), flush=True)
    nb = 1_000_000
    tt = np.linspace(0, 2 * np.pi, nb, endpoint=False)
    cl = np.stack([np.cos(tt), np.sin(tt)], -1)
    opl = tatva_b200.Operator(Mesh(coords=cl, elements=np.stack([np.arange(nb), (np.arange(nb) + 1) % nb], -1).astype(np.int32)), element.Line2())
    ul = torch.as_tensor(rng.normal(size=(nb, 2)), device="cuda")
    report("line2_grad_1M", timeit(lambda: opl._k_grad(ul)), 8 * (4 * nb + 2 * nb) + 8 * nb, nb, "qp")
    report("line2_weights_1M", timeit(lambda: opl.get_integration_weights()), 8 * (2 * nb + nb) + 8 * nb, nb, "qp", length=float(opl.get_integration_weights().sum()))

# ---- locality: the same Tet4 config-2 mesh with elements AND nodes randomly shuffled, then re-sorted ----
if want("locality"):
    from tatva_b200.mesh import reorder_mesh
    rng = np.random.default_rng(0)
    m = Mesh.box_tet((1.0, 1.0, 1.0), (55, 55, 55))
    c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / 55 * rng.uniform(-1, 1, m.coords.shape)
    el = m.elements
    pe, pn = rng.permutation(el.shape[0]), rng.permutation(c.shape[0])
    inv = np.empty_like(pn); inv[pn] = np.arange(len(pn))
    shuffled = Mesh(coords=c[pn], elements=inv[el[pe]].astype(np.int32))
    mat = materials.NeoHookean(500.0, 1000.0)
    N, E = c.shape[0], el.shape[0]
    t0 = time.perf_counter(); resorted, ep, npm = reorder_mesh(shuffled); t_sort = time.perf_counter() - t0
    for name, mesh, kw in (("given_order", Mesh(coords=c, elements=el), {}), ("shuffled", shuffled, {}), ("shuffled_sort_elements", shuffled, dict(sort_elements=True)), ("reordered_elements_and_nodes", resorted, {})):
        op = tatva_b200.Operator(mesh, element.Tetrahedron4(), **kw)
        u = torch.as_tensor(smooth_u(_np_coords := np.asarray(mesh.coords)), device="cuda")
        v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device="cuda")
        y = torch.empty_like(u)
        report(f"tet4_nh_hvp_c2_{name}", timeit(lambda: op._raw_hvp(mat, u, v, out=y)), 8 * (9 * N + 3 * N) + 16 * E, 3 * N, "DOF", host_reorder_s=round(t_sort, 3))
        del op
