import sys, json
sys.path.insert(0,'.')
import numpy as np, torch, tatva_b200
from tatva_b200 import element, materials, sparse
from tatva_b200.mesh import Mesh
m = Mesh.box_tet((1.0, 1.0, 1.0), (55, 55, 55))
c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / 55 * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
m = Mesh(coords=c, elements=m.elements)
op = tatva_b200.Operator(m, element.Tetrahedron4())
mat = materials.NeoHookean(500.0, 1000.0)
t = 2*np.pi
u = torch.as_tensor(0.05*np.stack([np.sin(t*c[:,0])*np.cos(t*c[:,1]), np.sin(t*c[:,1])*np.cos(t*c[:,2]), np.sin(t*c[:,2])*np.cos(t*c[:,0])],-1), device="cuda")
pat = sparse.pattern_from_mesh(m, 3)
cm = sparse.ColoredMatrix.from_csr(pat)
res = {}
ref = None
for tiled in (False, True):
    asm = sparse.assembler(op, mat, cm, tiled=tiled)
    data = torch.empty(asm.nnz, dtype=torch.float64, device="cuda")
    for _ in range(3): asm(u, out=data)
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): asm(u, out=data)
    b.record(); torch.cuda.synchronize()
    res["tiled_ms" if tiled else "grouped_ms"] = a.elapsed_time(b)/10
    if tiled:
        res["rel_diff"] = float((data - ref).norm() / ref.norm()); res.update(asm.tile_stats)
    else:
        ref = data.clone()
res["tag"] = sys.argv[1] if len(sys.argv)>1 else ""
print(json.dumps(res))
