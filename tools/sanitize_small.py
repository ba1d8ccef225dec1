"""Small run of the kernels added in the third session of round 2, for compute-sanitizer (memcheck / racecheck):
node-schedule residual / HVP (Tet4 neo-Hooke, linear elasticity, two-field; Tri3), element sub-ranges, Hex8 modal grad /
adjoint / weights, gather4, geometry cache, rank-structured diagonal, launch_behind_zero."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element, materials, _lib
from tatva_b200.mesh import Mesh

rng = np.random.default_rng(0)
m = Mesh.box_tet((1.0, 1.0, 1.0), (5, 5, 5))
c = m.coords + 0.02 * rng.uniform(-1, 1, m.coords.shape)
mesh = Mesh(coords=c, elements=m.elements)
op = tatva_b200.Operator(mesh, element.Tetrahedron4(), node_schedule=True)
u = torch.as_tensor(0.01 * rng.normal(size=c.shape), device="cuda")
v = torch.as_tensor(rng.normal(size=c.shape), device="cuda")
for mat in (materials.NeoHookean(500.0, 1000.0), materials.LinearElastic(0.38, 0.58)):
    for var in (0, 32, 33, 34, 31):
        op.set_variant(var)
        op._raw_hvp(mat, u, v)
        op._raw_residual(mat, u)
op.set_variant(0)
pf = materials.NeoHookeanPhaseField(500.0, 1000.0, 2.7, 0.1, 1e-6)
s = torch.as_tensor(np.concatenate([0.01 * rng.normal(size=c.shape), 0.5 + 0.2 * rng.uniform(size=(len(c), 1))], 1), device="cuda")
t = torch.as_tensor(rng.normal(size=(len(c), 4)), device="cuda")
for var in (0, 38, 31):
    op.set_variant(var)
    op._raw_hvp(pf, s, t)
    op._raw_residual(pf, s)
op.set_variant(0)
prm, npar = _lib.params_array(pf.params())
y = torch.empty_like(s)
E = m.elements.shape[0]
st = torch.cuda.current_stream().cuda_stream
for b, n, z in ((0, 256, 1), (256, E - 256, 0), (100, 300, 0)):
    _lib.check(op._L.tatva_hvp_elems(op._plan_fused, pf.material_id, prm, npar, s.data_ptr(), t.data_ptr(), y.data_ptr(), b, n, z, st))
op.hessian_diagonal(materials.NeoHookean(500.0, 1000.0), u)
m2 = Mesh.unit_square(19, 19)
op2 = tatva_b200.Operator(m2, element.Tri3(), node_schedule=True)
u2 = torch.as_tensor(0.01 * rng.normal(size=m2.coords.shape), device="cuda")
op2._raw_hvp(materials.LinearElastic(0.38, 0.58), u2, u2)
mh = Mesh.box_hex((5, 4, 3))
ch = mh.coords + 0.02 * rng.uniform(-1, 1, mh.coords.shape)
oph = tatva_b200.Operator(Mesh(coords=ch, elements=mh.elements), element.Hexahedron8(), cache_geometry=True)
nh = materials.NeoHookean(500.0, 1000.0)
uh = torch.as_tensor(0.01 * rng.normal(size=ch.shape), device="cuda")
vh = torch.as_tensor(rng.normal(size=ch.shape), device="cuda")
for var in (0, 53, 58, 26):
    oph.set_variant(var)
    oph._raw_hvp(nh, uh, vh)
oph.set_variant(0)
oph._raw_residual(nh, uh)
oph.hessian_diagonal(nh, uh)
buf = torch.empty(uh.numel() + 1, dtype=torch.float64, device="cuda")
oph._raw_hvp(nh, uh, vh, out=buf[1:].view_as(uh))
for nv in (1, 2, 3, 5):
    w = torch.as_tensor(rng.normal(size=(ch.shape[0], nv)), device="cuda")
    g = oph._k_grad(w)
    oph._k_grad_adj(g)
    oph._k_gather_adj(oph._k_gather(w))
oph.get_integration_weights()
torch.cuda.synchronize()
print("sanitize_small: done")
