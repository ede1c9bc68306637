"""Diagnostic (not a pytest test): tensor-core vs CUDA-core contraction on one evaluation, printing errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from perm_equiv_graph_neural_cdes_b200 import _lib
from oracle import reference_path as R
from tests.helpers import device_model, rel_err, product_grads_as_oracle

dev = torch.device("cuda:0")
for (n, h, e, L) in [(256, 32, 0, 1), (256, 32, 0, 2), (515, 32, 0, 2), (1000, 64, 0, 3), (300, 32, 3, 2), (1000, 64, 16, 3)]:
    p = R.make_problem(n=n, h=h, e=e, L=L, T=3, t1=2, dt0=0.5, seed=21)
    outs = {}
    for name, flags in (("ffma", 0), ("tc3", 1), ("tc1", 3)):
        vf, term, args = device_model(p, dev, flags=flags)
        y = p.y0.to(dev).requires_grad_(True)
        dy = term(1.3, y, args)
        (dy * p.gyT.to(dev)).sum().backward()
        torch.cuda.synchronize()
        fl = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
        outs[name] = (dy.detach(), y.grad.detach(), fl)
    for name in ("tc3", "tc1"):
        print(f"n={n} h={h} e={e} L={L} {name}: dy {rel_err(outs[name][0], outs['ffma'][0]):.2e}  gy {rel_err(outs[name][1], outs['ffma'][1]):.2e}"
              f"  gparams {rel_err(outs[name][2], outs['ffma'][2]):.2e}  finite={bool(torch.isfinite(outs[name][0]).all())}", flush=True)
