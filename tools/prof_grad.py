"""Operator.grad and its adjoint on Hex8 128^3 (modal kernels), for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tatva_b200
from tatva_b200 import element
from bench import synthetic_inputs

c, el, u, v = synthetic_inputs(128)
op = tatva_b200.Operator(tatva_b200.Mesh(coords=c, elements=el), element.Hexahedron8())
ut = torch.as_tensor(u, device="cuda")
for _ in range(3):
    g = op._k_grad(ut)
    y = op._k_grad_adj(g)
torch.cuda.synchronize()
