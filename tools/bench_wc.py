"""Warp-cooperative (node-schedule) residual / HVP kernels against the element-per-thread ones: config 2 (Tet4 n = 55,
neo-Hookean), config 5 (Tet4 n = 55, compound phase field), config 1 (Tri3 256^2), and the shuffled config-2 mesh with and
without locality sorting.  Every line carries the relative difference against the element-per-thread kernel."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials
from tatva_b200.mesh import Mesh


def timeit(fn, reps=100, warm=10):
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


def smooth(c):
    t = 2 * np.pi
    if c.shape[1] == 2:
        return 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 0])], -1)
    return 0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 2]), np.sin(t * c[:, 2]) * np.cos(t * c[:, 0])], -1)


def run(tag, mesh, el, mat, u, v, variants=(0,), cap=True, **kw):
    op0 = tatva_b200.Operator(mesh, el, node_schedule=False, **kw)
    t0 = time.perf_counter()
    op1 = tatva_b200.Operator(mesh, el, node_schedule=cap, **kw)
    host_s = time.perf_counter() - t0
    y = torch.empty_like(u)
    ref_h, ref_r = op0._raw_hvp(mat, u, v).clone(), op0._raw_residual(mat, u).clone()
    base_h, base_r = timeit(lambda: op0._raw_hvp(mat, u, v, out=y)), timeit(lambda: op0._raw_residual(mat, u))
    for var in variants:
        op1.set_variant(var)
        ms_h, ms_r = timeit(lambda: op1._raw_hvp(mat, u, v, out=y)), timeit(lambda: op1._raw_residual(mat, u))
        eh = float((op1._raw_hvp(mat, u, v) - ref_h).norm() / ref_h.norm())
        er = float((op1._raw_residual(mat, u) - ref_r).norm() / ref_r.norm())
        print(json.dumps({"case": tag, "variant": var, "elems": int(mesh.elements.shape[0]), "hvp_ms_element_per_thread": round(base_h, 4), "hvp_ms_node_schedule": round(ms_h, 4),
                          "residual_ms_element_per_thread": round(base_r, 4), "residual_ms_node_schedule": round(ms_r, 4), "hvp_rel_diff": eh, "residual_rel_diff": er,
                          "cap": int(cap), "schedule_host_s": round(host_s, 3), **op1.node_schedule_stats}), flush=True)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
m = Mesh.box_tet((1.0, 1.0, 1.0), (n, n, n))
rng = np.random.default_rng(0)
c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / n * rng.uniform(-1, 1, m.coords.shape)
mesh = Mesh(coords=c, elements=m.elements)
u = torch.as_tensor(smooth(c), device="cuda")
v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device="cuda")
for cap in (True, 4, 6):  # True: no contributor cap (the default)
    run("c2_tet4_nh", mesh, element.Tetrahedron4(), materials.NeoHookean(500.0, 1000.0), u, v, variants=(0, 32, 33, 34) if cap is True else (0,), cap=cap)
run("c2_tet4_nh_morton", mesh, element.Tetrahedron4(), materials.NeoHookean(500.0, 1000.0), u, v, sort_elements=True)
run("c2_tet4_le", mesh, element.Tetrahedron4(), materials.LinearElastic(0.38, 0.58), u, v)
phi = 0.4 + 0.4 * np.sin(2 * np.pi * c[:, 0]) * np.cos(2 * np.pi * c[:, 1])
s = torch.as_tensor(np.concatenate([smooth(c), phi[:, None]], axis=1), device="cuda")
t = torch.as_tensor(np.random.default_rng(2).normal(size=(c.shape[0], 4)), device="cuda")
run("c5_tet4_pf", mesh, element.Tetrahedron4(), materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.1, 1e-6), s, t, variants=(0, 38, 39))
perm = np.random.default_rng(3).permutation(m.elements.shape[0])
shuf = Mesh(coords=c, elements=m.elements[perm])
run("c2_tet4_nh_shuffled_elements", shuf, element.Tetrahedron4(), materials.NeoHookean(500.0, 1000.0), u, v)
run("c2_tet4_nh_shuffled_elements_morton", shuf, element.Tetrahedron4(), materials.NeoHookean(500.0, 1000.0), u, v, sort_elements=True)
m2 = Mesh.unit_square(256, 256)
c2 = m2.coords + 0.1 / 256 * rng.uniform(-1, 1, m2.coords.shape)
u2 = torch.as_tensor(smooth(c2), device="cuda")
v2 = torch.as_tensor(np.random.default_rng(1).normal(size=c2.shape), device="cuda")
run("c1_tri3_le", Mesh(coords=c2, elements=m2.elements), element.Tri3(), materials.LinearElastic(0.384615, 0.576923), u2, v2)
