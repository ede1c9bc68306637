"""Diagnostic: split-K vs single-pass vs FFMA on one evaluation + VJP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import device_model, rel_err, product_grads_as_oracle
dev = torch.device("cuda:0")
for (n, h, e, L, T) in [(1000, 64, 0, 2, 3), (1000, 64, 0, 3, 3), (1000, 64, 0, 1, 3), (515, 32, 0, 2, 3)]:
    p = R.make_problem(n=n, h=h, e=e, L=L, T=T, t1=2, dt0=0.5, seed=8)
    outs = {}
    for name, flags, env in (("ffma", 0, None), ("tc_nosplit", 1, "1"), ("tc_split", 1, None), ("tc_split2", 1, None)):
        if env: os.environ["PEG_TC_NO_SPLITK"] = env
        else: os.environ.pop("PEG_TC_NO_SPLITK", None)
        vf, term, args = device_model(p, dev, flags=flags)
        y = p.y0.to(dev).requires_grad_(True)
        dy = term(1.3, y, args)
        (dy * p.gyT.to(dev)).sum().backward()
        torch.cuda.synchronize()
        fl = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
        outs[name] = (dy.detach(), y.grad.detach(), fl)
    for name in ("tc_nosplit", "tc_split", "tc_split2"):
        print(f"n={n} h={h} L={L} {name}: dy {rel_err(outs[name][0], outs['ffma'][0]):.2e} gy {rel_err(outs[name][1], outs['ffma'][1]):.2e} gparams {rel_err(outs[name][2], outs['ffma'][2]):.2e}", flush=True)
