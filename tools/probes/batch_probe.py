"""Diagnostic: batched (B graphs in one call) vs one-graph-at-a-time evaluation + VJP, per kernel family / schedule knob."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from tests.helpers import rel_err

dev = torch.device("cuda:0")

def problem(n, h, B, T=3, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.rand((B, T, n, n), generator=g, device=dev) * (torch.rand((B, T, n, n), generator=g, device=dev) < 16.0 / n)
    A = A / A.sum(-1, keepdim=True).clamp_min(1e-3)
    ts = torch.arange(T, device=dev, dtype=torch.float32)
    y = torch.randn((B, n, h), generator=g, device=dev)
    gy = torch.randn((B, n, h), generator=g, device=dev)
    return ts, A, y, gy

def run(vf, pc, y, gy, t=1.3):
    vf.zero_grad()
    yy = y.detach().clone().requires_grad_(True)
    dy = vf(t, yy, pc)
    (dy * gy).sum().backward()
    return dy.detach(), yy.grad.detach(), torch.cat([p.grad.reshape(-1) for p in vf.parameters()])

for (n, h, B) in [(2048, 128, 3), (1024, 128, 2), (2048, 256, 2), (300, 64, 3)]:
    ts, A, y, gy = problem(n, h, B)
    pc = P.build_control(ts, A)
    pcs = [P.build_control(ts, A[b:b + 1]) for b in range(B)]
    for name, flags, env in [("ffma", 0, {}), ("tc-bf16", 1, {}), ("tc-bf16 nosplitk", 1, {"PEG_TC_NO_SPLITK": "1"}), ("tc-bf16 nolinear", 1, {"PEG_TC_NO_LINEAR": "1"}),
                             ("tc-bf16 nosplitk nolinear", 1, {"PEG_TC_NO_SPLITK": "1", "PEG_TC_NO_LINEAR": "1"}), ("tc-tf32x3", 17, {}), ("tc-tf32x3 nosplitk", 17, {"PEG_TC_NO_SPLITK": "1"})]:
        for k in ("PEG_TC_NO_SPLITK", "PEG_TC_NO_LINEAR"):
            os.environ.pop(k, None)
        os.environ.update(env)
        vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=1, flags=flags).to(dev)
        dyB, gB, gpB = run(vf, pc, y, gy)
        acc = torch.zeros_like(gpB)
        errs = []
        for b in range(B):
            dy1, g1, gp1 = run(vf, pcs[b], y[b:b + 1], gy[b:b + 1])
            acc += gp1
            errs.append("b%d dy %.1e gy %.1e" % (b, rel_err(dyB[b], dy1[0]), rel_err(gB[b], g1[0])))
        print(f"n={n} h={h} B={B} {name:28s} " + " | ".join(errs) + " | params %.1e" % rel_err(gpB, acc), flush=True)
