"""Re-derives the goldens from the REAL reference stack (jax + equinox + diffrax + the reference's own
src/models) wherever those packages are installed, and diffs them against the restatement in
oracle/reference_path.py.  It cannot run in this repository's build image (no jax/diffrax/equinox wheels, no
network) -- until it has been run somewhere, the THIRD-PARTY half of the oracle (diffrax / equinox arithmetic) stays "unpinned";
the reference's own code is pinned by oracle/pin_reference_source.py.

Usage (in an environment with the reference's dependencies, from the repo root):
    PYTHONPATH=/path/to/reference/src JAX_PLATFORMS=cpu python -m oracle.regen_with_jax
"""
import os
import sys

import numpy as np


def main():
    try:
        import diffrax
        import equinox as eqx
        import jax
        import jax.numpy as jnp
        from models.vector_fields import CDEWrapperVectorField, PermEquivGraphVectorField  # reference src/
    except Exception as exc:  # pragma: no cover - the whole point is that this cannot run here
        print(f"reference stack unavailable ({exc!r}); the third-party half of the oracle stays unpinned")
        return 2
    import torch

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import reference_path as R
    from tests.helpers import GOLDEN_CASES

    worst = 0.0
    for name, kw in GOLDEN_CASES.items():
        p = R.make_problem(**kw)
        widths = R.layer_widths(p.h, p.L, p.e, p.e > 0)
        vf = PermEquivGraphVectorField(p.h, p.h, widths[-1], p.L, p.e, p.n, key=jax.random.PRNGKey(0))
        # overwrite the Equinox leaves with the oracle's parameters
        for l, lp in enumerate(p.layers):
            fus, W, b, nw, nb = [jnp.asarray(t.numpy()) for t in lp.tensors()]
            for i in range(8):
                vf = eqx.tree_at(lambda m, l=l, i=i: getattr(m.gnn_layers[l], f"param{i+1}"), vf, fus[i])
            vf = eqx.tree_at(lambda m, l=l: m.gnn_layers[l].conv_layer.linear.weight, vf, W)
            vf = eqx.tree_at(lambda m, l=l: m.gnn_layers[l].conv_layer.linear.bias, vf, b)
            vf = eqx.tree_at(lambda m, l=l: m.gnn_layers[l].conv_layer.norm.weight, vf, nw)
            vf = eqx.tree_at(lambda m, l=l: m.gnn_layers[l].conv_layer.norm.bias, vf, nb)
        ts = jnp.asarray(p.ts.numpy())
        cadj = diffrax.CubicInterpolation(ts, tuple(jnp.asarray(c.numpy()) for c in p.coeffs_adj))
        if p.e > 0:
            cx = diffrax.CubicInterpolation(ts, tuple(jnp.asarray(c.numpy()) for c in p.x_coeffs))
            term, args = diffrax.ODETerm(CDEWrapperVectorField(vf, p.h)), [cadj, cx]
        else:
            term, args = diffrax.ODETerm(vf), cadj
        sol = diffrax.diffeqsolve(term, diffrax.Tsit5(), t0=ts[0], t1=ts[-1], dt0=kw["dt0"], y0=jnp.asarray(p.y0.numpy()),
                                  args=args, stepsize_controller=diffrax.ConstantStepSize(), saveat=diffrax.SaveAt(t1=True))
        yT = np.asarray(sol.ys[-1])
        ref = R.run_forward(R.problem_to(p, torch.float64)).numpy()
        err = float(np.abs(yT - ref).max() / np.abs(ref).max())
        worst = max(worst, err)
        print(f"{name}: steps real={int(sol.stats['num_steps'])} restated={len(p.step_ts)-1}  rel err vs restatement {err:.2e}")
    print("worst", worst)
    return 0 if worst < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
