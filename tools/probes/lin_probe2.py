"""Diagnostic: gradient sensitivity of the GNODE sibling case to 1e-7 relative perturbations of y0 (FFMA linear)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import rel_err
dev = torch.device("cuda:0")
os.environ["PEG_TC_NO_LINEAR"] = "1"
p = R.make_problem(n=140, h=32, e=0, L=3, T=4, t1=3, dt0=0.25, seed=17)
def run(cls, y0c, flags):
    vf = getattr(P, cls)(p.h, p.h, p.h, p.L, 0, p.n, key=0)
    with torch.no_grad():
        for mine, lp in zip(vf.gnn_layers, p.layers):
            mine.linear.weight.copy_(lp.weight); mine.linear.bias.copy_(lp.bias)
            mine.norm.weight.copy_(lp.norm_weight); mine.norm.bias.copy_(lp.norm_bias)
    vf = vf.to(dev); vf.flags = flags
    ts = p.ts.to(torch.float32).to(dev)
    ca = P.CubicInterpolation(ts, tuple(c.to(dev) for c in p.coeffs_adj))
    y0 = y0c.to(dev).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 3.0, 0.25, y0, ca, saveat=P.SaveAt(steps=True))
    (sol.ys[-1] * p.gyT.to(dev)).sum().backward()
    return sol.ys.detach(), y0.grad.detach()
for cls in ("GNODEVectorField", "GraphVectorField"):
    base = run(cls, p.y0, 1)
    for k in range(6):
        g = torch.Generator().manual_seed(k)
        yp = p.y0 * (1 + 1e-7 * torch.randn(p.y0.shape, generator=g))
        r = run(cls, yp, 1)
        print(cls, k, "ys", f"{rel_err(r[0], base[0]):.2e}", "gy0", f"{rel_err(r[1], base[1]):.2e}")
