"""Smallest forward + adjoint solve on both kernel families (CUDA-core FFMA, tcgen05 + TMA), for compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/probes/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import device_model

dev = torch.device("cuda:0")
for flags, n, h, e in ((0, 20, 8, 3), (P._lib.PEG_FLAG_TENSOR_CORES, 129, 32, 2), (P._lib.PEG_FLAG_TENSOR_CORES, 300, 32, 0)):
    p = R.make_problem(n=n, h=h, e=e, L=2, T=3, t1=1, dt0=0.5, seed=1)
    vf, term, args = device_model(p, dev, flags=flags)
    y0 = p.y0.to(dev).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 1.0, 0.5, y0, args, saveat=P.SaveAt(t1=True))
    (sol.ys[-1] * p.gyT.to(dev)).sum().backward()
    torch.cuda.synchronize()
    print(f"flags={flags} n={n} steps={sol.stats['num_steps']} |yT|={float(sol.ys[-1].abs().max()):.4f} |g|={float(y0.grad.abs().max()):.4f} "
          f"launches={P._lib.lib().pegncde_launch_count()}", flush=True)
