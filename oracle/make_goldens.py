"""Generates tests/golden/*.npz from the oracle (RESTATEMENT goldens: the real JAX/diffrax
reference cannot run in this image -- see oracle/regen_with_jax.py for the real-pinning path).

Each file holds, for one seeded problem (inputs are regenerated from the seed by the tests):
  yT64 / gy0_64 / gparams64 : fp64 truth of the oracle
  yT32                      : fp32 oracle (what the reference's dtype would give)
  cond                      : max |d yT / d y0| gain, recorded so tolerances can be read in context
  in_checksum               : sum of all inputs (guards against generator drift)
Run:  python -m oracle.make_goldens
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_path as R  # noqa: E402
from tests.helpers import GOLDEN_CASES  # noqa: E402


def input_checksum(p):
    s = float(p.y0.double().sum() + p.gyT.double().sum() + sum(c.double().sum() for c in p.coeffs_adj))
    if p.x_coeffs is not None:
        s += float(sum(c.double().sum() for c in p.x_coeffs))
    s += float(sum(t.double().sum() for lp in p.layers for t in lp.tensors()))
    return s


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, kw in GOLDEN_CASES.items():
        p32 = R.make_problem(**kw)
        p64 = R.problem_to(p32, torch.float64)
        yT64, gy064, g64 = R.run_forward_backward(p64)
        yT32, gy032, _ = R.run_forward_backward(p32)
        flat = np.concatenate([t.numpy().reshape(-1) for layer in g64 for t in layer])
        rel32 = float((yT32.double() - yT64).abs().max() / yT64.abs().max())
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"), yT64=yT64.numpy(), gy0_64=gy064.numpy(), gparams64=flat,
            yT32=yT32.numpy(), cond=float(gy064.abs().max()), in_checksum=input_checksum(p32), steps=len(p32.step_ts) - 1,
            rel32=rel32,
        )
        print(f"{name}: steps={len(p32.step_ts)-1} |yT|max={float(yT64.abs().max()):.4g} gain={float(gy064.abs().max()):.3g} "
              f"fp32-vs-fp64={rel32:.2e}")


if __name__ == "__main__":
    main()
