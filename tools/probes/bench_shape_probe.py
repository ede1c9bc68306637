"""Diagnostic: one evaluation + VJP at a benchmarked shape against the fp64 oracle, every graph of the batch, every kernel family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import device_model, rel_err, product_grads_as_oracle

dev = torch.device("cuda:0")
n, h, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
seeds = [31 + i for i in range(B)]
ps = [R.make_problem(n=n, h=h, e=0, L=3, T=3, t1=2, dt0=0.5, seed=s) for s in seeds]
for p in ps[1:]:
    p.layers = ps[0].layers
refs = []
for p in ps:
    p64 = R.problem_to(p, torch.float64)
    y64 = p64.y0.clone().requires_grad_(True)
    ref = R.perm_equiv_vector_field(1.3, y64, R.CubicInterpolation(p64.ts, p64.coeffs_adj), p64.layers)
    (ref * p64.gyT).sum().backward()
    refs.append((ref.detach(), y64.grad))
ts = ps[0].ts.to(torch.float32).to(dev)
for name, flags in (("ffma", 0), ("bf16x2", 1), ("tf32x3", 17)):
    vf, term, _ = device_model(ps[0], dev, flags=flags)
    for mode in ("batched", "single"):
        groups = [list(range(B))] if mode == "batched" else [[b] for b in range(B)]
        line = []
        for grp in groups:
            cadj = P.CubicInterpolation(ts, tuple(torch.stack([ps[b].coeffs_adj[i] for b in grp]).to(dev) for i in range(4)))
            y = torch.stack([ps[b].y0 for b in grp]).to(dev).requires_grad_(True)
            dy = term(1.3, y, cadj)
            (dy * torch.stack([ps[b].gyT for b in grp]).to(dev)).sum().backward()
            for k, b in enumerate(grp):
                e = (y.grad[k].double().cpu() - refs[b][1]).abs() / refs[b][1].abs().max()
                bad_rows = (e.max(dim=1).values > 5e-5).nonzero().flatten()
                line.append("b%d dy %.1e gy %.1e badrows %d%s" % (b, rel_err(dy[k].detach(), refs[b][0]), float(e.max()), bad_rows.numel(),
                                                                  (" first %s" % bad_rows[:6].tolist()) if bad_rows.numel() else ""))
        print(f"n={n} h={h} B={B} {name:7s} {mode:8s} " + " | ".join(line), flush=True)
