"""Diagnostic (not a pytest test): tensor-core vs CUDA-core RMSNorm->Linear inside a full solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import rel_err

dev = torch.device("cuda:0")
for cls in ("GNODEVectorField", "GraphVectorField"):
    p = R.make_problem(n=140, h=32, e=0, L=3, T=4, t1=3, dt0=0.25, seed=17)
    outs = {}
    for name, env, env_b in (("fma_lin", "1", "1"), ("tc_lin", None, None), ("tc_lin2", None, "1"), ("tc_lin3", "1", None)):
        if env: os.environ["PEG_TC_NO_LINEAR"] = env
        else: os.environ.pop("PEG_TC_NO_LINEAR", None)
        vf = getattr(P, cls)(p.h, p.h, p.h, p.L, 0, p.n, key=0)
        with torch.no_grad():
            for mine, lp in zip(vf.gnn_layers, p.layers):
                mine.linear.weight.copy_(lp.weight); mine.linear.bias.copy_(lp.bias)
                mine.norm.weight.copy_(lp.norm_weight); mine.norm.bias.copy_(lp.norm_bias)
        vf = vf.to(dev); vf.flags = 1
        ts = p.ts.to(torch.float32).to(dev)
        ca = P.CubicInterpolation(ts, tuple(c.to(dev) for c in p.coeffs_adj))
        y0 = p.y0.to(dev).requires_grad_(True)
        sol = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 3.0, 0.25, y0, ca)
        torch.cuda.synchronize()
        if env_b: os.environ["PEG_TC_NO_LINEAR"] = env_b
        else: os.environ.pop("PEG_TC_NO_LINEAR", None)
        (sol.ys[-1] * p.gyT.to(dev)).sum().backward()
        torch.cuda.synchronize()
        outs[name] = (sol.ys[-1].detach(), y0.grad.detach(), [m.linear.weight.grad.detach() for m in vf.gnn_layers], [m.norm.weight.grad.detach() for m in vf.gnn_layers])
    for name in ("tc_lin", "tc_lin2", "tc_lin3"):
        a, b = outs[name], outs["fma_lin"]
        print(cls, name, "yT", f"{rel_err(a[0], b[0]):.2e}", "gy0", f"{rel_err(a[1], b[1]):.2e}", "gW", [f"{rel_err(x, y):.1e}" for x, y in zip(a[2], b[2])], "gnw", [f"{rel_err(x, y):.1e}" for x, y in zip(a[3], b[3])])
