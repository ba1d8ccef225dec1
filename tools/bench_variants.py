"""Time every Hex8 neo-Hookean HVP kernel variant at 128^3 and check it against the generic kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials
from bench import synthetic_inputs

what = "hvp"
if len(sys.argv) > 1 and sys.argv[1] in ("hvp", "residual", "energy"):
    what = sys.argv.pop(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 15, 16, 17, 20, 22, 23, 25, 26, 27]
c, el, u, v = synthetic_inputs(n)
op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8(), cache_geometry=any(50 <= x <= 58 for x in variants))
mat = materials.NeoHookean(500.0, 1000.0)
ut, vt = torch.as_tensor(u, device="cuda"), torch.as_tensor(v, device="cuda")
y = torch.empty_like(ut)
run = {"hvp": lambda: op._raw_hvp(mat, ut, vt, out=y), "residual": lambda: y.copy_(op._raw_residual(mat, ut)), "energy": lambda: y.view(-1)[:1].copy_(op._raw_energy(mat, ut).reshape(1))}[what]
op.set_variant(1)
run()
ref = y.clone()
for var in variants:
    op.set_variant(var)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    err = float((y - ref).norm() / ref.norm()) if what != "energy" else float(abs(y.view(-1)[0] - ref.view(-1)[0]) / abs(ref.view(-1)[0]))
    print(json.dumps({"kernel": what, "variant": var, "ms": round(ms, 4), "gdof_s": round(3 * c.shape[0] / ms / 1e6, 3), "rel_err_vs_generic": err}))
