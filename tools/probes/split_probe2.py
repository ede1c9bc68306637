"""Diagnostic: single-pass TC contraction vs FFMA under different settings (n=1000, h=64, L=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import device_model, rel_err, product_grads_as_oracle
dev = torch.device("cuda:0")
def run(p, flags):
    vf, term, args = device_model(p, dev, flags=flags)
    y = p.y0.to(dev).requires_grad_(True)
    dy = term(1.3, y, args)
    (dy * p.gyT.to(dev)).sum().backward()
    torch.cuda.synchronize()
    return dy.detach(), y.grad.detach()
for seed in (8, 21):
    p = R.make_problem(n=1000, h=64, e=0, L=2, T=3, t1=2, dt0=0.5, seed=seed)
    os.environ.pop("PEG_TC_NO_SPLITK", None); os.environ.pop("PEG_TC_NO_LINEAR", None)
    ref = run(p, 0)
    for name, env in (("split", {}), ("nosplit", {"PEG_TC_NO_SPLITK": "1"}), ("nosplit+ffma_linear", {"PEG_TC_NO_SPLITK": "1", "PEG_TC_NO_LINEAR": "1"}),
                      ("split+ffma_linear", {"PEG_TC_NO_LINEAR": "1"})):
        os.environ.pop("PEG_TC_NO_SPLITK", None); os.environ.pop("PEG_TC_NO_LINEAR", None)
        os.environ.update(env)
        for rep in range(2):
            o = run(p, 1)
            print(f"seed={seed} {name} rep{rep}: dy {rel_err(o[0], ref[0]):.2e} gy {rel_err(o[1], ref[1]):.2e}", flush=True)
