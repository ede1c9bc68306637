"""diffrax-compatible term / solve API for the fused path.

Mirrors the call the reference makes
(src/models/pgt_graph_neural_cde.py:119-129, graph_neural_cde.py:94-104, tgb_graph_neural_cde.py:152-162)::

    diffeqsolve(terms=ODETerm(vf), solver=Tsit5(), t0=ts[0], t1=ts[-1], dt0=0.1, y0=y0,
                args=[control_adj, control_data], stepsize_controller=ConstantStepSize(),
                saveat=SaveAt(t1=True))  ->  Solution(.ys, .ts, .stats)

The whole integration (all steps x 7 stages x L layers) is enqueued by ONE C-ABI call
(``pegncde_solve_fwd``); reverse mode (``torch.autograd`` standing in for ``jax.custom_vjp``)
is ONE call of ``pegncde_solve_bwd`` -- the exact discrete adjoint of Tsit5, i.e. the
gradients diffrax's default ``RecursiveCheckpointAdjoint`` produces.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from ._lib import PEG_WS_SOLVE_BWD, PEG_WS_SOLVE_FWD, PEG_WS_STEP, check, lib
from .control import _stream_ptr
from .vector_field import CDEWrapperVectorField, PermEquivGraphVectorField, resolve_control, workspace


class Tsit5:
    """Marker for ``diffrax.Tsit5()`` -- the only solver the fused kernels implement."""


class ODETerm:
    def __init__(self, vector_field):
        self.vector_field = vector_field


class ConstantStepSize:
    """``diffrax.ConstantStepSize()``.  ``rule`` selects how the next step start is accumulated in fp32:
    "state" (t + dt0, diffrax >= 0.7) or "prev_diff" (t + (t - tprev), diffrax <= 0.6)."""

    def __init__(self, rule: str = "state"):
        self.rule = rule


@dataclass
class SaveAt:
    t1: bool = False
    steps: bool = False
    ts: Optional[torch.Tensor] = None


@dataclass
class Solution:
    ts: torch.Tensor
    ys: torch.Tensor
    stats: dict = field(default_factory=dict)


def constant_step_table(t0: float, t1: float, dt0: float, rule: str = "state", max_steps: int = 4096) -> np.ndarray:
    """fp32 step boundaries [S+1] under diffrax's ConstantStepSize + ``_clip_to_end``
    (``tnext > t1 - 1e-6 -> t1`` for fp32 times); raises like ``throw=True`` past ``max_steps``."""
    f = np.float32
    t0, t1, dt0 = f(t0), f(t1), f(dt0)
    if not dt0 > 0 or not t1 > t0:
        raise ValueError("need dt0 > 0 and t1 > t0")
    out = [t0]
    tprev, tnext = t0, f(t0 + dt0)
    while True:
        if tnext > f(t1 - f(1e-6)):
            tnext = t1
        out.append(tnext)
        if tnext >= t1:
            break
        if len(out) > max_steps:
            raise RuntimeError(f"max_steps={max_steps} reached")
        step = f(tnext - tprev) if rule == "prev_diff" else dt0
        tprev, tnext = tnext, f(tnext + step)
    return np.asarray(out, dtype=np.float32)


class _SolveFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, flat, pc, dims, step_ts, save_steps, store_stages):
        y0 = y0.contiguous()
        flat = flat.contiguous()
        S = len(step_ts) - 1
        l = lib()
        dev = y0.device
        nbytes = max(l.pegncde_workspace_bytes(dims, PEG_WS_SOLVE_FWD, S), l.pegncde_workspace_bytes(dims, PEG_WS_SOLVE_BWD, S))
        ws = workspace(dev, nbytes)
        y_ckpt = torch.empty((S + 1,) + tuple(y0.shape), dtype=torch.float32, device=dev)
        host_ts = np.ascontiguousarray(step_ts, dtype=np.float32)
        ctl = pc.struct()
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        store = None
        if store_stages and need_grad:
            nbytes_store = l.pegncde_stage_store_bytes(dims, S)
            capturing = torch.cuda.is_current_stream_capturing()
            free = torch.cuda.mem_get_info(dev)[0] if not capturing else float("inf")
            if nbytes_store < 0.5 * free:   # otherwise fall back to checkpoint-per-step + recompute
                store = torch.empty(nbytes_store // 4, dtype=torch.float32, device=dev)
        check(l.pegncde_solve_fwd(_stream_ptr(dev), dims, ctl, flat.data_ptr(),
                                  host_ts.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), S, y0.data_ptr(), None,
                                  y_ckpt.data_ptr(), store.data_ptr() if store is not None else None, ws.data_ptr(),
                                  ws.numel()), "pegncde_solve_fwd")
        ctx.save_for_backward(flat, y_ckpt)
        ctx.store = store
        ctx.pc, ctx.dims, ctx.host_ts, ctx.S, ctx.save_steps = pc, dims, host_ts, S, save_steps
        return y_ckpt if save_steps else y_ckpt[S]

    @staticmethod
    def backward(ctx, g_out):
        flat, y_ckpt = ctx.saved_tensors
        l = lib()
        dims, pc, S = ctx.dims, ctx.pc, ctx.S
        dev = flat.device
        nbytes = l.pegncde_workspace_bytes(dims, PEG_WS_SOLVE_BWD, S)
        ws = workspace(dev, nbytes)
        g_out = g_out.contiguous().to(torch.float32)
        if ctx.save_steps:
            g_ckpt = g_out
            g_yT = torch.zeros_like(y_ckpt[0])
        else:
            g_ckpt = None
            g_yT = g_out
        g_y0 = torch.empty_like(y_ckpt[0])
        g_flat = torch.zeros_like(flat)
        ctl = pc.struct()
        store = ctx.store
        check(l.pegncde_solve_bwd(_stream_ptr(dev), dims, ctl, flat.data_ptr(),
                                  ctx.host_ts.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), S, y_ckpt.data_ptr(),
                                  store.data_ptr() if store is not None else None, g_yT.data_ptr(), g_ckpt.data_ptr() if g_ckpt is not None else None, g_y0.data_ptr(),
                                  g_flat.data_ptr(), ws.data_ptr(), ws.numel()), "pegncde_solve_bwd")
        ctx.store = None
        return g_y0, g_flat, None, None, None, None, None


def _unwrap(term):
    vf = term.vector_field if isinstance(term, ODETerm) else term
    if isinstance(vf, CDEWrapperVectorField):
        return vf.vector_field, True
    if isinstance(vf, PermEquivGraphVectorField):
        return vf, False
    raise TypeError("the fused solve supports ODETerm(PermEquivGraphVectorField) and ODETerm(CDEWrapperVectorField(...))")


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, saveat: Optional[SaveAt] = None,
                stepsize_controller=None, max_steps: int = 4096, **unused) -> Solution:
    """Fused drop-in for ``diffrax.diffeqsolve`` on the hot path (Tsit5 + ConstantStepSize).

    ``y0``: ``[n,h]`` or batched ``[B,n,h]`` (the reference's ``jax.vmap(model)`` batch).  Adaptive
    controllers are driven step by step through :func:`tsit5_step` by the caller."""
    if not isinstance(solver, Tsit5):
        raise NotImplementedError("only Tsit5 is implemented by the fused kernels")
    controller = stepsize_controller or ConstantStepSize()
    if not isinstance(controller, ConstantStepSize):
        raise NotImplementedError("diffeqsolve here is the fixed-step path; use tsit5_step for adaptive control")
    if dt0 is None:
        raise ValueError("ConstantStepSize needs dt0")
    saveat = saveat or SaveAt(t1=True)
    if saveat.ts is not None:
        raise NotImplementedError("SaveAt(ts=...) needs Tsit5 dense output (adaptive path); use steps=True or t1=True")
    vf, wrapped = _unwrap(terms)
    if y0.device.type != "cuda":
        raise RuntimeError("the fused solve runs on CUDA only (no CPU fallback)")
    if wrapped:
        control_adj, control_data = args
    else:
        control_adj, control_data = (args[0] if isinstance(args, (list, tuple)) else args), None
    pc = resolve_control(control_adj, control_data, y0.device)
    dims = vf.dims_for(pc, with_wrapper=wrapped)
    step_ts = constant_step_table(float(t0), float(t1), float(dt0), controller.rule, max_steps)
    unb = y0.dim() == 2
    yb = (y0.unsqueeze(0) if unb else y0).to(torch.float32)
    if yb.shape[0] != pc.B:
        raise ValueError(f"state batch {yb.shape[0]} != control batch {pc.B}")
    out = _SolveFunction.apply(yb, vf.flat_params(), pc, dims, step_ts, bool(saveat.steps), bool(getattr(vf, "store_stages", True)))
    S = len(step_ts) - 1
    if saveat.steps:
        ys = out.squeeze(1) if unb else out
        ts_out = torch.from_numpy(step_ts.copy())
    else:
        ys = (out.squeeze(0) if unb else out).unsqueeze(0)  # diffrax: ys has a leading save axis of length 1
        ts_out = torch.tensor([float(step_ts[-1])])
    return Solution(ts=ts_out, ys=ys, stats={"num_steps": S, "num_accepted_steps": S, "num_rejected_steps": 0})


def tsit5_step(vf_term, t: float, dt: float, y: torch.Tensor, args, k1: Optional[torch.Tensor] = None):
    """One Tsit5 step through ``pegncde_step_fwd`` -> ``(y1, y_err, k7)`` for host-side adaptive
    controllers (PIDController) -- forward only."""
    vf, wrapped = _unwrap(vf_term)
    if wrapped:
        control_adj, control_data = args
    else:
        control_adj, control_data = (args[0] if isinstance(args, (list, tuple)) else args), None
    pc = resolve_control(control_adj, control_data, y.device)
    dims = vf.dims_for(pc, with_wrapper=wrapped)
    unb = y.dim() == 2
    yb = (y.unsqueeze(0) if unb else y).to(torch.float32).contiguous()
    l = lib()
    ws = workspace(y.device, l.pegncde_workspace_bytes(dims, PEG_WS_STEP, 1))
    k1_valid = k1 is not None
    k1b = (k1.unsqueeze(0) if (unb and k1_valid) else k1)
    k1b = k1b.contiguous().clone() if k1_valid else torch.empty_like(yb)
    y1, yerr, k7 = torch.empty_like(yb), torch.empty_like(yb), torch.empty_like(yb)
    flat = vf.flat_params().detach().contiguous()
    ctl = pc.struct()
    check(l.pegncde_step_fwd(_stream_ptr(y.device), dims, ctl, flat.data_ptr(), float(t), float(dt), yb.data_ptr(),
                             k1b.data_ptr(), 1 if k1_valid else 0, y1.data_ptr(), yerr.data_ptr(), k7.data_ptr(),
                             ws.data_ptr(), ws.numel()), "pegncde_step_fwd")
    if unb:
        return y1.squeeze(0), yerr.squeeze(0), k7.squeeze(0)
    return y1, yerr, k7
