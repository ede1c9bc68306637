"""JAX host layer over the XLA FFI shim (``pegncde_ffi.cc``): ``fused_diffeqsolve`` with the reference's call-site
signature, wrapped in ``jax.custom_vjp`` so ``eqx.filter_value_and_grad`` (src/engine/trainer_pgt.py:346) works.

STATUS: NOT IMPORTABLE IN THIS REPOSITORY'S IMAGE (no jax / equinox / diffrax wheels, no network).  It is the
binding a maintainer adds in a JAX environment; the torch/ctypes host in the parent package is the layer that is
tested on B200, against the same C-ABI.  Nothing in tests/, bench.py or __graft_entry__ imports this module.
"""
import ctypes
import os

import jax
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_shim = ctypes.CDLL(os.path.join(_HERE, "libpegncde_ffi.so"))
_core = ctypes.CDLL(os.path.join(os.path.dirname(_HERE), "libpegncde.so"))
jax.ffi.register_ffi_target("peg_solve_fwd", jax.ffi.pycapsule(_shim.PegSolveFwd), platform="CUDA")
jax.ffi.register_ffi_target("peg_solve_bwd", jax.ffi.pycapsule(_shim.PegSolveBwd), platform="CUDA")
jax.ffi.register_ffi_target("peg_pack_adj", jax.ffi.pycapsule(_shim.PegPackAdj), platform="CUDA")
jax.ffi.register_ffi_target("peg_pack_x", jax.ffi.pycapsule(_shim.PegPackX), platform="CUDA")


def pack_params(vector_field):
    """Equinox PermEquivGraphVectorField pytree -> the flat buffer of pegncde.h (layer after layer:
    weight | bias | norm.weight | norm.bias | param1..param8)."""
    parts = []
    for layer in vector_field.gnn_layers:
        cl = layer.conv_layer
        parts += [cl.linear.weight.reshape(-1), cl.linear.bias, cl.norm.weight, cl.norm.bias]
        parts += [getattr(layer, f"param{i}") for i in range(1, 9)]
    return jnp.concatenate(parts).astype(jnp.float32)


def _call(name, out_types, *args, **attrs):
    return jax.ffi.ffi_call(name, out_types, vmap_method="sequential")(*args, **attrs)


def pack_control(ts, coeffs_adj, x_coeffs, dims):
    """Once per batch: the reference-layout arrays the trainer passes to the model (``coeffs_adj`` = (d, c, b, a), each
    ``[B, T-1, n, n, 2]``; ``x_coeffs`` likewise ``[B, T-1, n, e, 2]`` or None) -> the control tuple of ``make_fused_solve``
    (ts, adj_coef, adj_rowsum, adj_diag, adj_total, tch_coef, x_coef) = the fields of ``PegControl`` in pegncde.h."""
    B, n, e, T = dims["B"], dims["n"], dims["e"], dims["T"]
    ldn = (n + 31) // 32 * 32
    attrs = {k: np.int32(v) for k, v in dims.items()}
    f32 = lambda *shape: jax.ShapeDtypeStruct(shape, jnp.float32)
    outs = (f32(B, T - 1, 4 * ldn * ldn), f32(B, T - 1, 4, n), f32(B, T - 1, 4, n), f32(B, T - 1, 4), f32(B, T - 1, 3, n))
    adj = _call("peg_pack_adj", outs, *[jnp.asarray(c, jnp.float32) for c in coeffs_adj], **attrs)
    if e > 0:
        (x_coef,) = _call("peg_pack_x", (f32(B, T - 1, 3, n, 2 * e),), *[jnp.asarray(c, jnp.float32) for c in x_coeffs], **attrs)
    else:
        x_coef = jnp.zeros((1,), jnp.float32)      # placeholder operand: the shim passes NULL when e == 0
    ts_b = jnp.broadcast_to(jnp.asarray(ts, jnp.float32).reshape(-1, T)[:1] if jnp.ndim(ts) == 1 else jnp.asarray(ts, jnp.float32), (B, T))
    return (ts_b, *adj, x_coef)


def make_fused_solve(dims, step_ts, ws_bytes, store_elems):
    """dims = dict(B,n,h,e,L,T,flags); returns solve(params_flat, control_tuple, y0) -> y_ckpt with a custom VJP."""
    S = len(step_ts) - 1
    B, n, h = dims["B"], dims["n"], dims["h"]
    attrs = dict(step_ts=np.asarray(step_ts, np.float32), **{k: np.int32(v) for k, v in dims.items()})

    @jax.custom_vjp
    def solve(params, control, y0):
        return _fwd(params, control, y0)[0]

    def _fwd(params, control, y0):
        outs = (jax.ShapeDtypeStruct((S + 1, B, n, h), jnp.float32), jax.ShapeDtypeStruct((store_elems,), jnp.float32),
                jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8))
        y_ckpt, store, _ = _call("peg_solve_fwd", outs, params, *control, y0, **attrs)
        return y_ckpt, (params, control, y_ckpt, store)

    def _bwd(res, g_ckpt):
        params, control, y_ckpt, store = res
        outs = (jax.ShapeDtypeStruct((B, n, h), jnp.float32), jax.ShapeDtypeStruct(params.shape, jnp.float32),
                jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8))
        g_y0, g_params, _ = _call("peg_solve_bwd", outs, params, *control, y_ckpt, store, g_ckpt, **attrs)
        return g_params, jax.tree_util.tree_map(jnp.zeros_like, control), g_y0

    solve.defvjp(_fwd, _bwd)
    return solve
