"""One Tet4 neo-Hookean HVP (config 2) through the node-schedule kernel, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials
from tatva_b200.mesh import Mesh

n = 55
m = Mesh.box_tet((1.0, 1.0, 1.0), (n, n, n))
c = m.coords + np.array([0.5, 0.5, 0.0]) + 0.1 / n * np.random.default_rng(0).uniform(-1, 1, m.coords.shape)
op = tatva_b200.Operator(Mesh(coords=c, elements=m.elements), element.Tetrahedron4(), node_schedule=True)
mat = materials.NeoHookean(500.0, 1000.0)
u = torch.as_tensor(0.01 * np.random.default_rng(2).normal(size=c.shape), device="cuda")
v = torch.as_tensor(np.random.default_rng(1).normal(size=c.shape), device="cuda")
y = torch.empty_like(u)
for _ in range(6):
    op._raw_hvp(mat, u, v, out=y)
torch.cuda.synchronize()
