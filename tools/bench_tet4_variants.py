"""Tet4 x neo-Hookean (config 2) HVP / residual: kernel variants side by side, each checked against the generic kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials
from tatva_b200.mesh import Mesh

variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 0, 30]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 55
m = Mesh.box_tet((1.0, 1.0, 1.0), (n, n, n))
c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / n * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
op = tatva_b200.Operator(Mesh(coords=c, elements=m.elements), element.Tetrahedron4())
mat = materials.NeoHookean(500.0, 1000.0)
t = 2 * np.pi
u = torch.as_tensor(0.05 * np.stack([np.sin(t * c[:, 0]) * np.cos(t * c[:, 1]), np.sin(t * c[:, 1]) * np.cos(t * c[:, 2]), np.sin(t * c[:, 2]) * np.cos(t * c[:, 0])], -1), device="cuda")
v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device="cuda")
y = torch.empty_like(u)


def timeit(fn, reps=50, warm=5):
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


op.set_variant(1)
ref_h, ref_r = op._raw_hvp(mat, u, v).clone(), op._raw_residual(mat, u).clone()
for var in variants:
    op.set_variant(var)
    ms_h = timeit(lambda: op._raw_hvp(mat, u, v, out=y))
    eh = float((op._raw_hvp(mat, u, v) - ref_h).norm() / ref_h.norm())
    ms_r = timeit(lambda: op._raw_residual(mat, u))
    er = float((op._raw_residual(mat, u) - ref_r).norm() / ref_r.norm())
    print(json.dumps({"variant": var, "n": n, "hvp_ms": round(ms_h, 4), "residual_ms": round(ms_r, 4), "hvp_rel_err": eh, "residual_rel_err": er}), flush=True)
