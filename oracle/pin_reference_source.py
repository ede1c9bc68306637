"""Pins the oracle against the reference's OWN source files (test infrastructure, NOT product code).

The reference's vector-field modules are plain Python over a handful of jax.numpy / equinox calls
(`jnp.tile/diag/eye/full/sum/transpose/einsum/mean`, `jax.vmap`, `jax.nn.relu`, `jr.split/uniform`,
`eqx.Module`, `eqx.nn.Linear/RMSNorm`).  jax / equinox are not installable in this image, so this script
installs a small numpy-backed stand-in for exactly those third-party names into ``sys.modules``,
imports the UNMODIFIED reference files from ``/root/reference/src/models/vector_fields`` and executes
their ``__call__`` methods in fp64:

    layers.py                               ConvLayer, ConvEquivFusionLayer(._fusion), ConvEquivFusionDirectedLayer
    perm_equiv_graph_vector_field.py        PermEquivGraphVectorField.__call__
    cde_wrapper_vector_field.py             CDEWrapperVectorField.__call__
    graph_vector_field.py / gnode_vector_field.py / perm_equiv_dir_graph_vector_field.py   (sibling fields)

What this pins: every line of the reference's own code on the path (SURVEY 8a rows a4-a7 and the N3 siblings),
including its quirks (term-7 bug, directed term-4' pairing, `1.0 +` offset, residual, no final ReLU,
time-gradient scaling, the `[n,h,e,2]` wrapper reshape).  What stays restated: the third-party pieces --
`eqx.nn.Linear` / `RMSNorm` arithmetic (stand-ins below), diffrax's CubicInterpolation (the control objects handed
to the reference code are the oracle's) and the Tsit5 loop (the whole-solve fixture runs the oracle's
fixed-step Tsit5 over the REFERENCE's vector-field callable).

Outputs: ``tests/golden/refsrc_<case>.npz`` (small, committed).  ``tests/test_oracle.py`` checks the oracle's
restatement against them on CPU; the GPU parity tests compare the CUDA path with them directly.
The reference tree does not travel to the GPU box -- only these fixtures do.

Run (in the build container, where /root/reference exists):   python -m oracle.pin_reference_source
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SRC = os.environ.get("PEG_REFERENCE_SRC", "/root/reference/src")

# case name -> (oracle make_problem kwargs, evaluation times).  Small on purpose: the stand-in for jax.vmap is a Python loop.
REFSRC_CASES = {
    "nocontrol": dict(n=14, h=8, e=0, L=2, T=4, t1=3, dt0=0.25, seed=11),
    "control": dict(n=18, h=8, e=3, L=3, T=4, t1=3, dt0=0.25, seed=12),
    "ragged": dict(n=33, h=16, e=2, L=3, T=5, t1=2, dt0=0.2, seed=13),
    "wide_dynamic_range": dict(n=21, h=8, e=2, L=3, T=4, t1=3, dt0=0.25, seed=14, scale=1.0e4),
    # n >= 128 and widths that are multiples of 32: the shapes the tcgen05 kernels are selected for
    "tc_control": dict(n=130, h=32, e=2, L=3, T=4, t1=3, dt0=0.5, seed=15),
    "tc_nocontrol": dict(n=160, h=64, e=0, L=2, T=4, t1=3, dt0=0.5, seed=16),
}
EVAL_TIMES = (0.0, 0.37, 1.0, 1.61, 2.0)
DIRECTED_SEED_OFFSET = 1000


# --------------------------------------------------------------------------------------
# numpy stand-ins for the third-party names the reference files import
# --------------------------------------------------------------------------------------


def _install_shims():
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "_peg_numpy_standin", False):
        raise RuntimeError("a real jax is importable here: use oracle/regen_with_jax.py instead")

    jnp = types.ModuleType("jax.numpy")
    for name in ("tile", "diag", "eye", "full", "sum", "transpose", "einsum", "mean", "squeeze", "ones", "zeros", "sqrt",
                 "asarray", "array", "stack", "concatenate", "reshape", "float32", "float64", "ndarray", "broadcast_to"):
        setattr(jnp, name, getattr(np, name))
    jnp.concat = np.concatenate

    jnn = types.ModuleType("jax.nn")
    jnn.relu = lambda x: np.maximum(x, 0.0)

    jr = types.ModuleType("jax.random")
    jr.PRNGKey = lambda seed: np.random.SeedSequence(int(seed))

    def split(key, num=2):
        return list(key.spawn(int(num)))

    def uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
        return np.random.default_rng(key).uniform(minval, maxval, size=shape).astype(np.float64)

    jr.split, jr.uniform = split, uniform

    jax = types.ModuleType("jax")
    jax._peg_numpy_standin = True
    jax.numpy, jax.nn, jax.random = jnp, jnn, jr
    jax.Array = np.ndarray
    jax.vmap = lambda f, in_axes=0: (lambda x: np.stack([f(xi) for xi in x]))   # jax.vmap(f)(x): f over axis 0

    class Module:
        """equinox.Module: the reference assigns its fields in __init__; nothing else of the base class is used here."""

        def __init__(self, **kwargs):
            pass

    class Linear(Module):
        """eqx.nn.Linear(in, out, key=): y = W x + b, W [out, in], init U(+-1/sqrt(in))."""

        def __init__(self, in_features, out_features, use_bias=True, *, key, **kwargs):
            lim = 1.0 / np.sqrt(in_features)
            wkey, bkey = key.spawn(2)
            self.weight = np.random.default_rng(wkey).uniform(-lim, lim, size=(out_features, in_features))
            self.bias = np.random.default_rng(bkey).uniform(-lim, lim, size=(out_features,))

        def __call__(self, x, *, key=None):
            return self.weight @ x + self.bias

    class RMSNorm(Module):
        """eqx.nn.RMSNorm(shape): x * rsqrt(mean(x^2) + eps) * weight + bias, eps = 1e-5, weight 1, bias 0."""

        def __init__(self, shape, eps=1e-5, use_weight=True, use_bias=True, **kwargs):
            shape = (shape,) if isinstance(shape, int) else tuple(shape)
            self.eps = eps
            self.weight, self.bias = np.ones(shape), np.zeros(shape)

        def __call__(self, x, *, key=None):
            inv_rms = 1.0 / np.sqrt(np.mean(x * x) + self.eps)
            return self.weight * (x * inv_rms) + self.bias

    class MLP(Module):
        """eqx.nn.MLP(in_size, out_size, width_size, depth, key=): `depth` hidden Linear layers with ReLU, identity output."""

        def __init__(self, in_size, out_size, width_size, depth, *, key, **kwargs):
            sizes = [in_size] + [width_size] * depth + [out_size]
            keys = key.spawn(len(sizes) - 1)
            self.layers = [Linear(sizes[i], sizes[i + 1], key=keys[i]) for i in range(len(sizes) - 1)]

        def __call__(self, x, *, key=None):
            for i, layer in enumerate(self.layers):
                x = layer(x)
                if i < len(self.layers) - 1:
                    x = np.maximum(x, 0.0)
            return x

    eqx = types.ModuleType("equinox")
    eqx_nn = types.ModuleType("equinox.nn")
    eqx.Module = Module
    eqx.field = lambda *a, **k: None
    eqx_nn.Linear, eqx_nn.RMSNorm, eqx_nn.MLP = Linear, RMSNorm, MLP
    eqx.nn = eqx_nn

    jaxtyping = types.ModuleType("jaxtyping")
    jaxtyping.Array = np.ndarray

    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.nn": jnn, "jax.random": jr, "equinox": eqx,
                        "equinox.nn": eqx_nn, "jaxtyping": jaxtyping})


_SHIM_NAMES = ("jax", "jax.numpy", "jax.nn", "jax.random", "equinox", "equinox.nn", "jaxtyping", "diffrax", "refsrc_models",
               "refsrc_models.vector_fields", "refsrc_models.neural_nets")


def uninstall_shims():
    """Removes the stand-ins (and the reference modules imported on top of them) from sys.modules again."""
    if not getattr(sys.modules.get("jax"), "_peg_numpy_standin", False):
        return
    for name in list(sys.modules):
        if name in _SHIM_NAMES or name.startswith("refsrc_models."):    # a stand-in jax means every one of these is ours
            del sys.modules[name]


def load_reference_vector_fields():
    """Imports the reference's vector-field files by their real module names without running the package __init__ files
    (those import every model of the repository, most of which need diffrax / pydantic configs)."""
    _install_shims()
    sys.dont_write_bytecode = True     # never write __pycache__ into the (read-only) reference tree
    vf_dir = os.path.join(REFERENCE_SRC, "models", "vector_fields")
    if not os.path.isdir(vf_dir):
        raise FileNotFoundError(f"{vf_dir} not found: the reference tree exists only in the build container")
    pkg_models = types.ModuleType("refsrc_models")
    pkg_models.__path__ = [os.path.join(REFERENCE_SRC, "models")]
    pkg_vf = types.ModuleType("refsrc_models.vector_fields")
    pkg_vf.__path__ = [vf_dir]
    pkg_nn = types.ModuleType("refsrc_models.neural_nets")      # `from ..neural_nets import IdxEncoder`: dead import
    pkg_nn.IdxEncoder = type("IdxEncoder", (), {"__init__": lambda self, *a, **k: None})
    sys.modules.update({"refsrc_models": pkg_models, "refsrc_models.vector_fields": pkg_vf, "refsrc_models.neural_nets": pkg_nn})
    out = {}
    for mod in ("layers", "perm_equiv_graph_vector_field", "cde_wrapper_vector_field", "graph_vector_field",
                "gnode_vector_field", "perm_equiv_dir_graph_vector_field"):
        out[mod] = importlib.import_module(f"refsrc_models.vector_fields.{mod}")
    return out


def _install_diffrax_standin():
    """`diffrax` for the reference's solve wrappers: every name they use, backed by the ORACLE's restatement of the
    third-party arithmetic (so the model-level fixtures pin the reference's own glue code -- encoders, the arguments of the
    diffeqsolve call, read-outs -- not diffrax itself)."""
    import torch

    import oracle.reference_path as R

    dfx = types.ModuleType("diffrax")
    to_t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))

    class CubicInterpolation:
        def __init__(self, ts, coeffs):
            self.ctrl = R.CubicInterpolation(to_t(ts), tuple(to_t(c) for c in coeffs))

        def evaluate(self, t):
            return self.ctrl.evaluate(t).numpy()

        def derivative(self, t):
            return self.ctrl.derivative(t).numpy()

    def backward_hermite_coefficients(ts, ys):
        return tuple(c.numpy() for c in R.backward_hermite_coefficients(to_t(ts), to_t(ys)))

    class LinearInterpolation:
        """diffrax.LinearInterpolation(ts, ys), restated: piecewise linear between the knot values, left-continuous lookup
        (index = clip(searchsorted(ts, t, 'left') - 1, 0, T-2)) like CubicInterpolation."""

        def __init__(self, ts, ys):
            self.ts, self.ys = np.asarray(ts, dtype=np.float64), np.asarray(ys, dtype=np.float64)

        def _piece(self, t):
            i = int(np.clip(np.searchsorted(self.ts, float(t), side="left") - 1, 0, len(self.ts) - 2))
            return i, (self.ys[i + 1] - self.ys[i]) / (self.ts[i + 1] - self.ts[i])

        def evaluate(self, t):
            i, slope = self._piece(t)
            return self.ys[i] + slope * (float(t) - self.ts[i])

        def derivative(self, t):
            return self._piece(t)[1]

    def linear_interpolation(ts, ys):
        return np.asarray(ys)       # no NaNs to fill in

    class ODETerm:
        def __init__(self, vector_field):
            self.vector_field = vector_field

    class Tsit5:
        pass

    class ConstantStepSize:
        pass

    class PIDController:
        def __init__(self, rtol, atol):
            self.rtol, self.atol = rtol, atol

    class SaveAt:
        def __init__(self, ts=None, t1=False):
            self.ts, self.t1 = ts, t1

    class Solution:
        def __init__(self, ys, stats):
            self.ys, self.stats = ys, stats

    def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, stepsize_controller=None, saveat=None, **kwargs):
        assert isinstance(solver, Tsit5) and not kwargs, "the reference passes nothing else (default adjoint / max_steps)"
        f = lambda t, y: torch.from_numpy(np.asarray(terms.vector_field(float(t), y.numpy(), args), dtype=np.float64))
        y0 = to_t(y0)
        if isinstance(stepsize_controller, ConstantStepSize):
            table = R.constant_step_table(float(t0), float(t1), float(dt0))
            if saveat.ts is not None:     # evolving_out=True: dense output of every step that holds a save time
                ys, _, _ = R.tsit5_solve_adaptive(f, y0, float(t0), float(t1), save_ts=[float(t) for t in np.asarray(saveat.ts)],
                                                  forced_steps=table)
                return Solution(ys.numpy(), {"num_steps": len(table) - 1})
            assert saveat.t1
            return Solution(R.tsit5_solve_fixed(f, y0, table).numpy()[None], {"num_steps": len(table) - 1})
        assert isinstance(stepsize_controller, PIDController) and dt0 is None
        save_ts = None if saveat.ts is None else [float(t) for t in np.asarray(saveat.ts)]
        ys, table, stats = R.tsit5_solve_adaptive(f, y0, float(t0), float(t1), rtol=stepsize_controller.rtol,
                                                  atol=stepsize_controller.atol, dt0=None, save_ts=save_ts)
        ys = ys.numpy() if save_ts is not None else ys.numpy()[None]
        return Solution(ys, dict(stats, table=table))

    for name, obj in dict(CubicInterpolation=CubicInterpolation, backward_hermite_coefficients=backward_hermite_coefficients,
                          LinearInterpolation=LinearInterpolation, linear_interpolation=linear_interpolation,
                          ODETerm=ODETerm, Tsit5=Tsit5, ConstantStepSize=ConstantStepSize, PIDController=PIDController,
                          SaveAt=SaveAt, diffeqsolve=diffeqsolve).items():
        setattr(dfx, name, obj)
    dfx._peg_numpy_standin = True
    sys.modules["diffrax"] = dfx


def load_reference_models():
    """The reference's solve wrappers (src/models/{pgt_,tgb_,}graph_neural_cde.py), imported unmodified on top of the stand-ins."""
    mods = load_reference_vector_fields()
    _install_diffrax_standin()
    mods["gnode_floor_vector_field"] = importlib.import_module("refsrc_models.vector_fields.gnode_floor_vector_field")
    pkg_vf = sys.modules["refsrc_models.vector_fields"]        # what `from . import vector_fields` resolves to
    pkg_vf.CDEWrapperVectorField = mods["cde_wrapper_vector_field"].CDEWrapperVectorField
    pkg_vf.GNODEFloorVectorField = mods["gnode_floor_vector_field"].GNODEFloorVectorField
    sys.modules["refsrc_models"].vector_fields = pkg_vf
    for mod in ("pgt_graph_neural_cde", "tgb_graph_neural_cde", "graph_neural_cde"):
        mods[mod] = importlib.import_module(f"refsrc_models.{mod}")
    return mods


# --------------------------------------------------------------------------------------
# glue: oracle parameters / control objects -> reference objects
# --------------------------------------------------------------------------------------


class NumpyControl:
    """The oracle's CubicInterpolation (restated diffrax) behind the `.evaluate(t)` / `.derivative(t)` protocol, as numpy."""

    def __init__(self, ctrl):
        self.ctrl = ctrl

    def evaluate(self, t):
        return self.ctrl.evaluate(t).numpy()

    def derivative(self, t):
        return self.ctrl.derivative(t).numpy()


DIRECTED_FIELDS = ("param1", "param2", "param3", "param4", "param4_prime", "param5", "param5_prime", "param6",
                   "param6_prime", "param7", "param8")


def directed_fusion_tables(L: int, seed: int):
    """[L][11, 2] parameter tables for the directed layer, reference init distribution U(-1,1)/15 (layers.py:230-250)."""
    import torch

    g = torch.Generator().manual_seed(seed + DIRECTED_SEED_OFFSET)
    return [(torch.rand((11, 2), generator=g, dtype=torch.float64) * 2 - 1) / 15.0 for _ in range(L)]


def _set_conv(conv_layer, lp):
    conv_layer.linear.weight = lp.weight.numpy().copy()
    conv_layer.linear.bias = lp.bias.numpy().copy()
    conv_layer.norm.weight = lp.norm_weight.numpy().copy()
    conv_layer.norm.bias = lp.norm_bias.numpy().copy()


def build_reference_fields(mods, p64, dir_tables):
    """Reference modules constructed through their own __init__ and loaded with the oracle problem's parameters."""
    import jax.random as jr   # the stand-in

    import oracle.reference_path as R

    widths = R.layer_widths(p64.h, p64.L, p64.e, p64.e > 0)
    key = jr.PRNGKey(0)
    kw = dict(input_dim=p64.h, hidden_dim=p64.h, output_dim=widths[-1], num_layers=p64.L, data_embed_dim=p64.e, num_nodes=p64.n)
    pe = mods["perm_equiv_graph_vector_field"].PermEquivGraphVectorField(**kw, key=key)
    pd = mods["perm_equiv_dir_graph_vector_field"].PermEquivDirGraphVectorField(**kw, key=key)
    gv = mods["graph_vector_field"].GraphVectorField(**kw, key=key)
    gn = mods["gnode_vector_field"].GNODEVectorField(**kw, key=key)
    for l, lp in enumerate(p64.layers):
        for i in range(8):
            setattr(pe.gnn_layers[l], f"param{i + 1}", lp.fusion[i].numpy().copy())
        _set_conv(pe.gnn_layers[l].conv_layer, lp)
        for i, name in enumerate(DIRECTED_FIELDS):
            setattr(pd.gnn_layers[l], name, dir_tables[l][i].numpy().copy())
        _set_conv(pd.gnn_layers[l].conv_layer, lp)
        _set_conv(gv.gnn_layers[l], lp)
        _set_conv(gn.gnn_layers[l], lp)
    return pe, pd, gv, gn


# model-level cases: name -> (kind, oracle make_problem kwargs, extra sizes)
MODEL_CASES = {
    "pgt": ("pgt", dict(n=24, h=8, e=3, L=3, T=4, t1=3, dt0=0.1, seed=31), dict(data_dim=5, feature_dim=1)),
    "tgb_mlp": ("tgb", dict(n=20, h=8, e=4, L=2, T=3, t1=2, dt0=0.01, seed=32), dict(use_mlps=True)),
    "tgb_linear": ("tgb", dict(n=16, h=8, e=2, L=2, T=3, t1=2, dt0=0.01, seed=33), dict(use_mlps=False)),
    "dyn": ("dyn", dict(n=30, h=8, e=0, L=2, T=8, t1=5, dt0=0.1, seed=34, float_ts=True), dict()),
    # evolving_out=True (SaveAt(ts=ts) on the fixed-step path, pgt_graph_neural_cde.py:114-117 / tgb_graph_neural_cde.py:147-150,164-167)
    "pgt_evolving": ("pgt", dict(n=18, h=8, e=2, L=2, T=4, t1=3, dt0=0.1, seed=35), dict(data_dim=4, feature_dim=2, evolving_out=True)),
    "tgb_sequence": ("tgb", dict(n=14, h=8, e=2, L=2, T=3, t1=2, dt0=0.01, seed=36), dict(use_mlps=True, evolving_out=True, return_sequence=True)),
    # interpolation="linear" (pgt_graph_neural_cde.py:101-103): the knot VALUES [T,n,n,2] / [T,n,e,2] are the "coefficients"
    "pgt_linear": ("pgt", dict(n=18, h=8, e=2, L=2, T=4, t1=3, dt0=0.1, seed=37), dict(data_dim=4, feature_dim=1, interpolation="linear")),
}


def linear_knot_values(p64):
    """Knot values of the paths whose Hermite coefficients an oracle Problem holds: ``a`` of every piece plus the end point of the
    last piece -- the arrays a ``linear`` config hands to the model (dataset_configs.py builds them with diffrax.linear_interpolation)."""
    import torch

    def knots(coeffs, ts):
        d, c, b, a = coeffs
        h = (ts[-1] - ts[-2]).to(a.dtype)
        last = a[-1] + h * (b[-1] + h * (c[-1] + h * d[-1]))
        return torch.cat([a, last[None]], dim=0)

    return knots(p64.coeffs_adj, p64.ts), knots(p64.x_coeffs, p64.ts)


def model_inputs(name):
    """Seeded extra inputs of a model-level case (node features, raw node signals): shared by this script and the tests."""
    kind, kw, extra = MODEL_CASES[name]
    rng = np.random.default_rng(kw["seed"] + 7000)
    n, T = kw["n"], kw["T"]
    if kind == "pgt":
        return {"x0": rng.standard_normal((n, extra["data_dim"]))}
    if kind == "tgb":
        return {"x0": rng.standard_normal((n, n)), "x_data": 0.5 * rng.standard_normal((T, n, n))}
    return {"x0": rng.standard_normal((n, 1))}


def _linear_params(lin):
    return [lin.weight.copy(), lin.bias.copy()]


def _module_params(mod):
    """[(W, b), ...] of a stand-in MLP or Linear."""
    layers = mod.layers if hasattr(mod, "layers") else [mod]
    return [_linear_params(l) for l in layers]


def main_models():
    """Model-level fixtures: the reference's solve wrappers executed on the stand-ins (diffrax backed by the oracle)."""
    import types as _types

    import torch

    sys.path.insert(0, ROOT)
    import oracle.reference_path as R
    import jax.random as jr

    mods = load_reference_models()
    out_dir = os.path.join(ROOT, "tests", "golden")
    worst = 0.0
    for name, (kind, kw, extra) in MODEL_CASES.items():
        p64 = R.problem_to(R.make_problem(**kw), torch.float64)
        widths = R.layer_widths(p64.h, p64.L, p64.e, p64.e > 0)
        vf = mods["perm_equiv_graph_vector_field"].PermEquivGraphVectorField(
            input_dim=p64.h, hidden_dim=p64.h, output_dim=widths[-1], num_layers=p64.L, data_embed_dim=p64.e, num_nodes=p64.n, key=jr.PRNGKey(0))
        for l, lp in enumerate(p64.layers):
            for i in range(8):
                setattr(vf.gnn_layers[l], f"param{i + 1}", lp.fusion[i].numpy().copy())
            _set_conv(vf.gnn_layers[l].conv_layer, lp)
        inp = model_inputs(name)
        coeffs_adj = tuple(c.numpy() for c in p64.coeffs_adj)
        rec = {}
        tt = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
        lay = lambda ps: [(tt(W), tt(b)) for W, b in ps]
        if kind == "pgt":
            cfg = _types.SimpleNamespace(data_dim=extra["data_dim"], hidden_dim=p64.h, feature_dim=extra["feature_dim"], method="Tsit5", return_sequence=False)
            interp = extra.get("interpolation", "cubic")
            model = mods["pgt_graph_neural_cde"].PGTGraphNeuralCDE(cfg, vf, interp, jr.PRNGKey(kw["seed"]))
            ts = np.arange(kw["T"], dtype=np.int32)                                   # torch.arange in dataset_configs.py:1108
            ev = dict(evolving_out=True) if extra.get("evolving_out") else {}
            if interp == "linear":
                A_k, X_k = linear_knot_values(p64)
                rec["adj_knots"], rec["x_knots"] = A_k.numpy(), X_k.numpy()
                c_adj, x_coeffs = rec["adj_knots"], rec["x_knots"]
                ora_adj = (torch.zeros_like(p64.coeffs_adj[0]), torch.zeros_like(p64.coeffs_adj[0]), (A_k[1:] - A_k[:-1]), A_k[:-1])   # unit knot spacing
                ora_x = (torch.zeros_like(p64.x_coeffs[0]), torch.zeros_like(p64.x_coeffs[0]), (X_k[1:] - X_k[:-1]), X_k[:-1])
            else:
                c_adj = np.stack(coeffs_adj)
                x_coeffs = np.stack([c.numpy() for c in p64.x_coeffs])                # stacked [4, T-1, n, e, 2] like trainer_pgt.py:203
                ora_adj, ora_x = p64.coeffs_adj, p64.x_coeffs
            rec["out_global"] = model(ts, c_adj, x_coeffs, inp["x0"], **ev)
            rec["out_nodes"] = model(ts, c_adj, x_coeffs, inp["x0"], global_readout=False, **ev)
            enc, dec = _module_params(model.encoder), _module_params(model.decoder)
            ora = R.pgt_graph_neural_cde(p64.ts, ora_adj, ora_x, tt(inp["x0"]), lay(enc), lay(dec), p64.layers, p64.h, p64.e,
                                         evolving_out=bool(ev))
            err = float((ora - tt(rec["out_global"])).abs().max() / tt(rec["out_global"]).abs().max())
        elif kind == "tgb":
            seq = bool(extra.get("return_sequence"))
            cfg = _types.SimpleNamespace(hidden_dim=p64.h, use_mlps=extra["use_mlps"], method="Tsit5", return_sequence=seq)
            model = mods["tgb_graph_neural_cde"].TGBGraphNeuralCDE(cfg, vf, "cubic", jr.PRNGKey(kw["seed"]))
            ts = np.arange(kw["T"], dtype=np.int32)
            rec["out"] = model(ts, coeffs_adj, inp["x_data"], inp["x0"], None, **(dict(evolving_out=True) if extra.get("evolving_out") else {}))
            enc, dec = _module_params(model.encoder), _module_params(model.decoder)
            rec["data_encoder_W"], rec["data_encoder_b"] = _linear_params(model.data_encoder)
            ora = R.tgb_graph_neural_cde(p64.ts, p64.coeffs_adj, tt(inp["x_data"]), tt(inp["x0"]), lay(enc), lay(dec),
                                         (tt(rec["data_encoder_W"]), tt(rec["data_encoder_b"])), p64.layers, p64.h, p64.e,
                                         evolving_out=bool(extra.get("evolving_out")), return_sequence=seq)
            err = float((ora - tt(rec["out"])).abs().max() / tt(rec["out"]).abs().max())
        else:
            cfg = _types.SimpleNamespace(hidden_dim=p64.h, method="Tsit5", return_sequence=True)
            model = mods["graph_neural_cde"].GraphNeuralCDE(cfg, vf, "cubic", jr.PRNGKey(kw["seed"]))
            ts = p64.ts.numpy()
            rec["out"] = model(ts, coeffs_adj, inp["x0"])                              # evolving_out=True -> [T, n, 1]
            enc, dec = _module_params(model.initial_linear), _module_params(model.final_linear)
            ora, table = R.graph_neural_cde(p64.ts, p64.coeffs_adj, tt(inp["x0"]), lay(enc)[0], lay(dec)[0], p64.layers)
            rec["accepted_steps"] = len(table) - 1
            err = float((ora - tt(rec["out"])).abs().max() / tt(rec["out"]).abs().max())
        for i, (W, b) in enumerate(enc):
            rec[f"enc_W{i}"], rec[f"enc_b{i}"] = W, b
        for i, (W, b) in enumerate(dec):
            rec[f"dec_W{i}"], rec[f"dec_b{i}"] = W, b
        worst = max(worst, err)
        np.savez_compressed(os.path.join(out_dir, f"refsrc_model_{name}.npz"), **rec)
        print(f"refsrc_model_{name}: restated model vs reference source: {err:.2e}")
    print("worst (models)", worst)
    return worst


def load_reference_dataset_configs():
    """src/configs/dataset_configs.py imported unmodified.  Its heavy imports (torch_geometric, tgb, the reference's own
    `dataset` package) only serve type annotations and loaders that the control-path builders below never touch, so they are
    empty placeholder modules here."""
    load_reference_models()          # jax / equinox / diffrax stand-ins
    placeholders = {
        "torch_geometric": {}, "torch_geometric.data": {"Data": object, "TemporalData": object},
        "torch_geometric.utils": {"to_dense_adj": None}, "tgb": {}, "tgb.nodeproppred": {},
        "tgb.nodeproppred.dataset_pyg": {"PyGNodePropPredDataset": object},
        "dataset": {"ODEDataset": object, "misc": types.ModuleType("dataset.misc")},
        "dataset.tgb_dataset": {"SlidingWindowTemporalLoader": object},
    }
    added = []
    for name, attrs in placeholders.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m._peg_placeholder = True
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            added.append(name)
    import importlib.util

    path = os.path.join(REFERENCE_SRC, "configs", "dataset_configs.py")
    spec = importlib.util.spec_from_file_location("refsrc_models.dataset_configs", path)   # lives under the prefix uninstall_shims() clears
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    try:
        spec.loader.exec_module(mod)
    finally:
        for name in added:
            sys.modules.pop(name, None)
    return mod


def load_reference_function(rel_path, func_name, namespace):
    """One top-level function of a reference file, compiled from its own (unmodified) source lines -- for modules whose
    import pulls in the whole training stack (wandb, exca, optax, every config class): src/engine/trainer_pgt.py."""
    import ast

    path = os.path.join(REFERENCE_SRC, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func_name)
    module = ast.Module(body=[node], type_ignores=[])
    exec(compile(module, path, "exec"), namespace)
    return namespace[func_name]


def reference_pgt_mse_loss():
    """mse_loss of src/engine/trainer_pgt.py:45-66 on the numpy stand-in for jax.numpy."""
    _install_shims()
    import jax.numpy as jnp   # the stand-in

    return load_reference_function(os.path.join("engine", "trainer_pgt.py"), "mse_loss",
                                   {"jnp": jnp, "PGTGraphNeuralODE": type("PGTGraphNeuralODE", (), {})})


DATASET_CASE = dict(n=11, e=3, window=5, seed=41)     # England-like window: 5 snapshots, the last one is the label


def dataset_window(case=DATASET_CASE):
    """A seeded window of graph snapshots (objects with .x [n,e], .adj [n,n], .y [n]) like PGTDataSetCfg.process_window receives."""
    import torch

    import oracle.reference_path as R

    A = torch.from_numpy(R.synthetic_graph_path(case["n"], case["window"], case["seed"]))
    g = torch.Generator().manual_seed(case["seed"])
    return [types.SimpleNamespace(x=torch.randn((case["n"], case["e"]), generator=g, dtype=torch.float64), adj=A[k],
                                  y=torch.randn((case["n"],), generator=g, dtype=torch.float64)) for k in range(case["window"])]


def main_dataset():
    """The reference's control-path builders (what the trainers feed the models, trainer_pgt.py:201-207 / trainer.py:121-144)."""
    import torch

    sys.path.insert(0, ROOT)
    import oracle.reference_path as R

    mod = load_reference_dataset_configs()
    window = dataset_window()
    cfg = types.SimpleNamespace(interpolation="cubic")
    cfg.get_interpolation_coeffs = lambda ts, sig: mod.PGTDataSetCfg.get_interpolation_coeffs(cfg, ts, sig)
    d = mod.PGTDataSetCfg.process_window(cfg, window)                                   # dataset_configs.py:1103-1131
    rec = {"t": d["t"].numpy(), "true_y": d["true_y"].numpy(), "true_y0": d["true_y0"].numpy()}
    for i, nm in enumerate("dcba"):
        rec[f"graph_{nm}"] = d["graph_path_coeffs"][i].numpy()
        rec[f"x_{nm}"] = d["x_coeffs"][i].numpy()
    ts = torch.arange(len(window) - 1)
    A = torch.stack([w.adj for w in window[:-1]])
    x_t = torch.stack([w.x for w in window[:-1]])
    worst = 0.0
    for i, (mine_g, mine_x) in enumerate(zip(R.reference_layout_coeffs(ts, A), R.reference_layout_xcoeffs(ts, x_t))):
        worst = max(worst, float((mine_g - d["graph_path_coeffs"][i]).abs().max()), float((mine_x - d["x_coeffs"][i]).abs().max()))
    # dynamical-systems config: float time stamps, dataset_configs.py:147-173
    ode_cfg = types.SimpleNamespace(interpolation="cubic")
    ts_f = np.linspace(0.0, 5.0, 6)
    A_f = R.synthetic_graph_path(9, 6, DATASET_CASE["seed"] + 1)
    co = mod.ODEDataSetCfg.get_graph_interpolation_coeffs(ode_cfg, ts_f, A_f)
    for i, nm in enumerate("dcba"):
        rec[f"ode_graph_{nm}"] = np.asarray(co[i])
    for i, mine in enumerate(R.reference_layout_coeffs(torch.from_numpy(ts_f), torch.from_numpy(A_f))):
        worst = max(worst, float((mine - torch.from_numpy(np.asarray(co[i]))).abs().max()))
    # the loss the PGT trainer differentiates (trainer_pgt.py:45-66): the model's global read-out [feature_dim] reshaped to
    # (feature_dim, 1) against label [n] -- broadcast (1,1) - (n,) -> (1,n), kept as the reference computes it
    mse = reference_pgt_mse_loss()
    rng = np.random.default_rng(DATASET_CASE["seed"])
    y_pred, label = rng.standard_normal((1,)), rng.standard_normal((DATASET_CASE["n"],))
    rec["loss_y_pred"], rec["loss_label"] = y_pred, label
    rec["loss"] = mse(lambda t, a, x, x0: y_pred, (None, None, None, None, label))
    mine = R.pgt_mse_loss(torch.ones(1, 1, dtype=torch.float64), torch.from_numpy(y_pred).reshape(1, 1), torch.from_numpy(label))
    worst = max(worst, abs(float(mine) - float(rec["loss"])))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refsrc_dataset.npz"), **rec)
    print(f"refsrc_dataset: graph coeffs {tuple(rec['graph_d'].shape)} x coeffs {tuple(rec['x_d'].shape)}  restated layout vs reference source: {worst:.2e}")
    return worst


def main():
    import torch

    sys.path.insert(0, ROOT)
    import oracle.reference_path as R

    mods = load_reference_vector_fields()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    worst = 0.0
    for name, kw in REFSRC_CASES.items():
        p64 = R.problem_to(R.make_problem(**kw), torch.float64)
        dir_tables = directed_fusion_tables(p64.L, kw["seed"])
        pe, pd, gv, gn = build_reference_fields(mods, p64, dir_tables)
        cadj = R.CubicInterpolation(p64.ts, p64.coeffs_adj)
        ncadj = NumpyControl(cadj)
        y = p64.y0.numpy()
        rec = {}
        # layer-level: the reference's _fusion on the interpolated (A, A') of the first evaluation time after a knot
        adj = ncadj.evaluate(EVAL_TIMES[1])[..., -1]
        dadj = ncadj.derivative(EVAL_TIMES[1])[..., -1]
        fus, fus_dir = pe.gnn_layers[0]._fusion(adj, dadj), pd.gnn_layers[0]._fusion(adj, dadj)
        if p64.n <= 40:
            rec["fusion_t1"], rec["fusion_dir_t1"] = fus, fus_dir
        # checksums that see every entry with a different weight (kept for all sizes)
        wgt = np.cos(np.arange(p64.n * p64.n, dtype=np.float64)).reshape(p64.n, p64.n)
        rec["fusion_t1_wsum"], rec["fusion_dir_t1_wsum"] = float((fus * wgt).sum()), float((fus_dir * wgt).sum())
        times = [t for t in EVAL_TIMES if t <= float(p64.ts[-1])]
        if p64.n > 40:
            times = [times[0], times[1], times[3]]      # a knot, an interior point of the linear piece, one of a cubic piece
        rec["times"] = np.asarray(times)
        fields = {"perm_equiv": pe, "perm_equiv_dir": pd, "graph": gv, "gnode": gn}
        if p64.e > 0:
            # control shapes: every field behind the reference's CDEWrapperVectorField ([n, 2he] -> [n, h])
            ncx = NumpyControl(R.CubicInterpolation(p64.ts, p64.x_coeffs))
            Wrapper = mods["cde_wrapper_vector_field"].CDEWrapperVectorField
            for key, field in fields.items():
                if key == "gnode":      # the reference's GNODEVectorField keeps no data_embed_dim: it cannot sit behind the wrapper
                    continue
                wrapped_field = Wrapper(field, p64.h)
                rec[f"vf_{key}"] = np.stack([wrapped_field(t, y, [ncadj, ncx]) for t in times])
            wrapped = Wrapper(pe, p64.h)
            f_ref = lambda t, yy: torch.from_numpy(wrapped(float(t), yy.numpy(), [ncadj, ncx]))
        else:
            for key, field in fields.items():
                rec[f"vf_{key}"] = np.stack([field(t, y, ncadj) for t in times])
            f_ref = lambda t, yy: torch.from_numpy(pe(float(t), yy.numpy(), ncadj))
        # whole solve: the oracle's restated Tsit5 (third-party arithmetic) over the REFERENCE's vector-field callable
        rec["yT"] = R.tsit5_solve_fixed(f_ref, p64.y0, p64.step_ts).numpy()
        rec["steps"] = len(p64.step_ts) - 1
        from oracle.make_goldens import input_checksum

        rec["in_checksum"] = input_checksum(R.make_problem(**kw))
        # report how far the restatement is from the reference source (the CPU test asserts this stays < 1e-11)
        yT_restated = R.run_forward(p64).numpy()
        err = float(np.abs(yT_restated - rec["yT"]).max() / np.abs(rec["yT"]).max())
        worst = max(worst, err)
        np.savez_compressed(os.path.join(out_dir, f"refsrc_{name}.npz"), **rec)
        print(f"refsrc_{name}: n={p64.n} h={p64.h} e={p64.e} L={p64.L} steps={rec['steps']}  |yT|max={np.abs(rec['yT']).max():.4g}  "
              f"restatement vs reference source: {err:.2e}")
    print("worst", worst)
    worst = max(worst, main_models())
    worst = max(worst, main_dataset())
    uninstall_shims()
    return 0 if worst < 1e-11 else 1


if __name__ == "__main__":
    sys.exit(main())
