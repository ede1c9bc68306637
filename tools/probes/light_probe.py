"""Diagnostic: LIGHT adjoint (two single-pass Frobenius products) vs the four-3xTF32-product adjoint -- per-evaluation
errors of the fusion-scalar gradients and of the state cotangent against the fp64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import perm_equiv_graph_neural_cdes_b200 as P
from oracle import reference_path as R
from tests.helpers import device_model, rel_err, product_grads_as_oracle

dev = torch.device("cuda:0")
for (n, h, e, L, seed) in [(1000, 64, 16, 3, 21), (1000, 64, 0, 3, 21), (515, 32, 0, 2, 21), (300, 32, 3, 2, 21), (129, 64, 8, 3, 21), (1000, 64, 0, 3, 5), (2048, 128, 0, 3, 7)]:
    p = R.make_problem(n=n, h=h, e=e, L=L, T=3, t1=2, dt0=0.5, seed=seed)
    p64 = R.problem_to(p, torch.float64)
    layers = R.params_to(p64.layers, requires_grad=True)
    q = R.Problem(p64.n, p64.h, p64.e, p64.L, p64.ts, p64.coeffs_adj, p64.x_coeffs, p64.y0, layers, p64.step_ts, p64.gyT)
    y64 = p64.y0.clone().requires_grad_(True)
    ca = R.CubicInterpolation(q.ts, q.coeffs_adj)
    if e > 0:
        ref = R.cde_wrapper_vector_field(1.3, y64, ca, R.CubicInterpolation(q.ts, q.x_coeffs), layers, h, e)
    else:
        ref = R.perm_equiv_vector_field(1.3, y64, ca, layers)
    (ref * p64.gyT).sum().backward()
    for mode in ("light", "full"):
        if mode == "light": os.environ["PEG_TC_ADJ_LIGHT"] = "1"
        else: os.environ.pop("PEG_TC_ADJ_LIGHT", None)
        vf, term, args = device_model(p, dev, flags=1)
        y = p.y0.to(dev).requires_grad_(True)
        dy = term(1.3, y, args)
        (dy * p.gyT.to(dev)).sum().backward()
        errs = []
        for l, (got, lp) in enumerate(zip(product_grads_as_oracle(vf), layers)):
            fus_ref = lp.fusion.grad
            per = ((got[0].double().cpu() - fus_ref).abs() / fus_ref.abs().max()).reshape(-1)
            errs.append("L%d fus max %.1e (p1 %.1e %.1e p2 %.1e %.1e) W %.1e" % (l, float(per.max()), per[0], per[1], per[2], per[3], rel_err(got[1], lp.weight.grad)))
        print(f"n={n} h={h} e={e} {mode:5s} gy {rel_err(y.grad, y64.grad):.1e} | " + " | ".join(errs), flush=True)
os.environ.pop("PEG_TC_ADJ_LIGHT", None)
