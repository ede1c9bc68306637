"""CPU ORACLE (test infrastructure, NOT product code) -- reference-owned code pinned, third-party
solver arithmetic unpinned.

Literal restatement, in PyTorch-CPU, of the one hot path of
hits-mli/perm-equiv-graph-neural-cdes: the Tsit5 solve loop that evaluates the
permutation-equivariant graph vector field at every stage.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``perm_equiv_graph_neural_cdes_b200`` never does.

PINNED: the vector-field functions below (fusion, conv_layer, perm_equiv_vector_field,
cde_wrapper_vector_field, the directed / plain sibling fields) reproduce, to 2e-15, fp64
executions of the reference's UNMODIFIED source files on numpy stand-ins for jax.numpy / equinox
(``oracle/pin_reference_source.py`` -> ``tests/golden/refsrc_*.npz``, checked by
``tests/test_oracle.py``).
PARITY UNPINNED for the third-party pieces: the real stack (JAX + Equinox + diffrax) cannot be
imported in this image (no wheels, no network) and the reference ships no test, golden vector or
fixture for any function on this path (reference ``test/`` covers dataset utilities only).  The
diffrax / equinox pieces (Hermite coefficients, cubic interpolation, Tsit5, step-size controllers,
RMSNorm / Linear arithmetic) are restated from their published algorithms (diffrax >= 0.5, version
unpinned in reference ``environment.yaml:16``).  ``oracle/regen_with_jax.py`` re-derives the goldens
from the real packages wherever they are installed.  Every function follows the cited reference
lines op-for-op -- including materialising the fused n x n adjacency exactly like the reference.

All functions are dtype-generic (fp32 = what the reference computes in; fp64 = truth
the tolerance is set against) and differentiable with torch autograd, which stands in
for ``eqx.filter_value_and_grad`` through diffrax's default RecursiveCheckpointAdjoint
(exact discretise-then-optimise gradients).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# third-party: diffrax.backward_hermite_coefficients  (call sites
# src/configs/dataset_configs.py:168-170, 228-230, 1094-1096)
# --------------------------------------------------------------------------------------


def backward_hermite_coefficients(ts: torch.Tensor, ys: torch.Tensor):
    """Returns (d, c, b, a), each [T-1, ...], like diffrax.backward_hermite_coefficients.

    Per interval i (dt = t[i+1]-t[i], secant m_i = (y[i+1]-y[i])/dt): cubic with
    p(0)=y_i, p'(0)=m_{i-1} (m_0 for the first interval), p(dt)=y_{i+1}, p'(dt)=m_i.
    """
    ts = ts.to(ys.dtype)
    dt = (ts[1:] - ts[:-1]).reshape((-1,) + (1,) * (ys.dim() - 1))
    m = (ys[1:] - ys[:-1]) / dt
    b = torch.cat([m[:1], m[:-1]], dim=0)
    a = ys[:-1]
    c = 2.0 * (m - b) / dt
    d = -(m - b) / (dt * dt)
    return d, c, b, a


class CubicInterpolation:
    """diffrax.CubicInterpolation(ts, (d, c, b, a)) -- evaluate / derivative.

    Used at src/models/pgt_graph_neural_cde.py:105-107,
    src/models/graph_neural_cde.py:82-83.
    """

    def __init__(self, ts: torch.Tensor, coeffs):
        self.ts = ts
        self.d, self.c, self.b, self.a = coeffs

    def _interpret_t(self, t):
        # index = clip(searchsorted(ts, t, side="left") - 1, 0, T-2); frac = t - ts[index]
        tsf = self.ts.to(torch.float64)
        idx = int(torch.searchsorted(tsf, torch.as_tensor(float(t), dtype=torch.float64), right=False)) - 1
        idx = max(0, min(idx, self.ts.numel() - 2))
        frac = torch.as_tensor(t, dtype=self.a.dtype) - self.ts[idx].to(self.a.dtype)
        return idx, frac

    def evaluate(self, t):
        i, s = self._interpret_t(t)
        return self.a[i] + s * (self.b[i] + s * (self.c[i] + s * self.d[i]))

    def derivative(self, t):
        i, s = self._interpret_t(t)
        return self.b[i] + s * (2.0 * self.c[i] + 3.0 * s * self.d[i])


# --------------------------------------------------------------------------------------
# parameters (field names follow the reference pytree: layers.py:19-20, 66-74)
# --------------------------------------------------------------------------------------


@dataclass
class LayerParams:
    """One ConvEquivFusionLayer: param1..param8 (each [2]) + conv_layer.{linear,norm}."""

    fusion: torch.Tensor  # [8, 2]  rows = param1..param8
    weight: torch.Tensor  # [d_out, d_in]   conv_layer.linear.weight
    bias: torch.Tensor  # [d_out]          conv_layer.linear.bias
    norm_weight: torch.Tensor  # [d_in]    conv_layer.norm.weight
    norm_bias: torch.Tensor  # [d_in]      conv_layer.norm.bias

    def tensors(self):
        return [self.fusion, self.weight, self.bias, self.norm_weight, self.norm_bias]


def layer_widths(hidden_dim: int, num_layers: int, data_embed_dim: int, use_control: bool) -> List[int]:
    """[d_0, ..., d_L]: src/configs/vector_field_configs.py:66-76,100-108 and
    perm_equiv_graph_vector_field.py:47-61."""
    out = hidden_dim * data_embed_dim * 2 if use_control else hidden_dim
    return [hidden_dim] * num_layers + [out]


def init_params(widths: Sequence[int], seed: int, dtype=torch.float32, randomize_norm: bool = False) -> List[LayerParams]:
    """Reference init distributions (layers.py:86-100; eqx.nn.Linear U(+-1/sqrt(in));
    RMSNorm weight 1 / bias 0).  ``randomize_norm`` perturbs the norm affine so tests
    exercise those gradients away from the init point."""
    g = torch.Generator().manual_seed(seed)
    layers = []
    for l in range(len(widths) - 1):
        din, dout = widths[l], widths[l + 1]
        fusion = (torch.rand((8, 2), generator=g, dtype=torch.float64) * 2 - 1) / 15.0
        lim = 1.0 / math.sqrt(din)
        W = (torch.rand((dout, din), generator=g, dtype=torch.float64) * 2 - 1) * lim
        b = (torch.rand((dout,), generator=g, dtype=torch.float64) * 2 - 1) * lim
        nw = torch.ones(din, dtype=torch.float64)
        nb = torch.zeros(din, dtype=torch.float64)
        if randomize_norm:
            nw = nw + 0.2 * (torch.rand((din,), generator=g, dtype=torch.float64) * 2 - 1)
            nb = nb + 0.2 * (torch.rand((din,), generator=g, dtype=torch.float64) * 2 - 1)
        layers.append(LayerParams(*(x.to(dtype) for x in (fusion, W, b, nw, nb))))
    return layers


def params_to(layers: List[LayerParams], dtype=None, requires_grad=False) -> List[LayerParams]:
    out = []
    for lp in layers:
        ts = [x.detach().clone().to(dtype or x.dtype).requires_grad_(requires_grad) for x in lp.tensors()]
        out.append(LayerParams(*ts))
    return out


# --------------------------------------------------------------------------------------
# the vector field (reference, materialised)
# --------------------------------------------------------------------------------------


def fusion(adjacency: torch.Tensor, control_gradient: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """ConvEquivFusionLayer._fusion, src/models/vector_fields/layers.py:102-160.

    ``p`` is [8,2] = param1..param8.  Builds the n x n fused adjacency term by term in
    the reference's order, INCLUDING the term-7 quirk (both coefficients multiply
    sum(adjacency), layers.py:144-148)."""
    n = adjacency.shape[0]
    one = torch.ones((), dtype=adjacency.dtype)
    eye = torch.eye(n, dtype=adjacency.dtype)
    term_1 = (1.0 + p[0, 0]) * adjacency + (1.0 + p[0, 1]) * control_gradient
    term_2 = p[1, 0] * adjacency.t() + p[1, 1] * control_gradient.t()
    term_3 = p[2, 0] * torch.diag(torch.diag(adjacency)) + p[2, 1] * torch.diag(torch.diag(control_gradient))
    rs_a = adjacency.sum(dim=1)
    rs_c = control_gradient.sum(dim=1)
    # transpose(tile(rowsum,(n,1)))[i,j] = rowsum[i]
    term_4 = p[3, 0] / n * rs_a[:, None].expand(n, n) + p[3, 1] / n * rs_c[:, None].expand(n, n)
    # tile(rowsum,(n,1))[i,j] = rowsum[j]
    term_5 = p[4, 0] / n * rs_a[None, :].expand(n, n) + p[4, 1] / n * rs_c[None, :].expand(n, n)
    term_6 = p[5, 0] / n * torch.diag(rs_a) + p[5, 1] / n * torch.diag(rs_c)
    tot_a = adjacency.sum()
    term_7 = p[6, 0] / n**2 * (tot_a * one).expand(n, n) + p[6, 1] / n**2 * (tot_a * one).expand(n, n)
    term_8 = (p[7, 0] * tot_a + p[7, 1] * control_gradient.sum()) / n**2 * eye
    return term_1 + term_2 + term_3 + term_4 + term_5 + term_6 + term_7 + term_8


def rms_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """equinox.nn.RMSNorm(shape) applied per node (jax.vmap(self.norm), layers.py:45):
    x * rsqrt(mean(x^2) + eps) * weight + bias."""
    inv = torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps)
    return x * inv * weight + bias


def conv_layer(node_feats, adj_matrix, lp: LayerParams):
    """ConvLayer.__call__, layers.py:36-48: norm -> linear -> m + adj @ m."""
    z = rms_norm(node_feats, lp.norm_weight, lp.norm_bias)
    m = z @ lp.weight.t() + lp.bias
    return m + adj_matrix @ m


def perm_equiv_vector_field(t, y, control_adj: CubicInterpolation, layers: List[LayerParams]):
    """PermEquivGraphVectorField.__call__, perm_equiv_graph_vector_field.py:85-129
    (enc_idx=False; the enc_idx branch is unreachable in the reference)."""
    adj = control_adj.evaluate(t)[..., -1]
    adj_derivative = control_adj.derivative(t)[..., -1]
    t_gradient = control_adj.derivative(t)[..., 0]
    node_features = y
    for i, lp in enumerate(layers):
        fused = fusion(adj, adj_derivative, lp.fusion)  # layers.py:174
        node_features = conv_layer(node_features, fused, lp)  # layers.py:176
        if i < len(layers) - 1:
            node_features = torch.relu(node_features)
    t_gradient = t_gradient.mean(dim=0)  # [nodes]
    return t_gradient[:, None] * node_features


def fusion_directed(adjacency: torch.Tensor, control_gradient: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """ConvEquivFusionDirectedLayer._fusion, layers.py:256-345, term by term.  p [11, 2] rows = param1, param2, param3, param4,
    param4_prime, param5, param5_prime, param6, param6_prime, param7, param8 (the class's field order).  Quirks kept: term 4'
    pairs rowsum(A) with colsum(A') (``:288-292``), term 7 uses sum(adjacency) for both coefficients (``:317-321``)."""
    n = adjacency.shape[0]
    A, D = adjacency, control_gradient
    eye = torch.eye(n, dtype=A.dtype)
    ones = torch.ones((n, n), dtype=A.dtype)
    col = lambda X: X.sum(dim=0)   # jnp.sum(., axis=0)
    row = lambda X: X.sum(dim=1)   # jnp.sum(., axis=1)
    tile = lambda v: v[None, :].expand(n, n)          # jnp.tile(v, (n, 1)): every row is v
    term_1 = (1.0 + p[0, 0]) * A + (1.0 + p[0, 1]) * D
    term_2 = p[1, 0] * A.t() + p[1, 1] * D.t()
    term_3 = p[2, 0] * torch.diag(torch.diag(A)) + p[2, 1] * torch.diag(torch.diag(D))
    term_4 = p[3, 0] / n * tile(col(A)).t() + p[3, 1] / n * tile(col(D)).t()
    term_4p = p[4, 0] / n * tile(row(A)) + p[4, 1] / n * tile(col(D))
    term_5 = p[5, 0] / n * tile(col(A)) + p[5, 1] / n * tile(col(D))
    term_5p = p[6, 0] / n * tile(row(A)) + p[6, 1] / n * tile(row(D))
    term_6 = p[7, 0] / n * torch.diag(col(A)) + p[7, 1] / n * torch.diag(col(D))
    term_6p = p[8, 0] / n * torch.diag(row(A)) + p[8, 1] / n * torch.diag(row(D))
    term_7 = p[9, 0] / n**2 * ones * A.sum() + p[9, 1] / n**2 * ones * A.sum()
    term_8 = (p[10, 0] * A.sum() + p[10, 1] * D.sum()) / n**2 * eye
    return term_1 + term_2 + term_3 + term_4 + term_4p + term_5 + term_5p + term_6 + term_6p + term_7 + term_8


def perm_equiv_dir_vector_field(t, y, control_adj: CubicInterpolation, layers: List[LayerParams], fusions: Sequence[torch.Tensor]):
    """PermEquivDirGraphVectorField.__call__ (perm_equiv_dir_graph_vector_field.py:86-130, enc_idx=False); ``fusions[l]`` is
    the [11, 2] parameter table of layer l (``layers[l].fusion`` is ignored)."""
    value, deriv = control_adj.evaluate(t), control_adj.derivative(t)
    adj, adj_derivative, t_gradient = value[..., -1], deriv[..., -1], deriv[..., 0]
    z = y
    for i, (lp, pf) in enumerate(zip(layers, fusions)):
        z = conv_layer(z, fusion_directed(adj, adj_derivative, pf), lp)
        if i < len(layers) - 1:
            z = torch.relu(z)
    return t_gradient.mean(dim=0)[:, None] * z


def plain_graph_vector_field(t, y, control_adj: CubicInterpolation, layers: List[LayerParams], with_derivative: bool):
    """GraphVectorField.__call__ (graph_vector_field.py:80-115, enc_idx=False; message passing matrix A + A') and
    GNODEVectorField.__call__ (gnode_vector_field.py:57-81; A only): plain ConvLayers, ReLU between, time-gradient scale."""
    adj, adj_derivative = control_adj.evaluate(t), control_adj.derivative(t)
    mp = adj[..., -1] + adj_derivative[..., -1] if with_derivative else adj[..., -1]
    z = y
    for i, lp in enumerate(layers):
        z = conv_layer(z, mp, lp)
        if i < len(layers) - 1:
            z = torch.relu(z)
    t_gradient = adj_derivative[:, :, 0].mean(dim=0)
    return t_gradient[:, None] * z


def cde_wrapper_vector_field(t, y, control_adj, control_data, layers, hidden_dim, data_embed_dim):
    """CDEWrapperVectorField.__call__, cde_wrapper_vector_field.py:19-26."""
    out = perm_equiv_vector_field(t, y, control_adj, layers).reshape(-1, hidden_dim, data_embed_dim, 2)
    return torch.einsum("nmlk,nlk->nm", out, control_data.derivative(t))


# --------------------------------------------------------------------------------------
# third-party: diffrax Tsit5 + ConstantStepSize + diffeqsolve (fixed step)
# --------------------------------------------------------------------------------------

# Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the first column
# simplifying assumption" -- the tableau diffrax.Tsit5 uses.
TSIT5_C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
TSIT5_A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)
TSIT5_B = TSIT5_A[6] + (0.0,)
# b_sol - b_hat (error estimate weights); needed only by the adaptive path
TSIT5_BERR = (
    -0.00178001105222577714,
    -0.0008164344596567469,
    0.007880878010261995,
    -0.1447110071732629,
    0.5823571654525552,
    -0.45808210592918697,
    0.015151515151515152,
)


def constant_step_table(t0: float, t1: float, dt0: float, rule: str = "state", max_steps: int = 4096) -> np.ndarray:
    """Step boundaries [S+1] (fp32) of diffeqsolve(..., dt0, ConstantStepSize()).

    fp32 time accumulation with diffrax's end clipping (``_clip_to_end``:
    ``tnext > t1 - 1e-6  ->  t1`` for non-float64 times).  ``rule="state"`` is
    ``tnext = t + dt0`` (diffrax >= 0.7 keeps dt0 in the controller state);
    ``rule="prev_diff"`` is ``tnext = t + (t - tprev)`` (diffrax <= 0.6).  The two
    differ by ulps in the boundaries and at most by one ~1e-6-long final step."""
    f = np.float32
    t0, t1, dt0 = f(t0), f(t1), f(dt0)
    ts = [t0]
    tprev, tnext = t0, f(t0 + dt0)
    tol = f(1e-6)
    while True:
        if tnext > f(t1 - tol):
            tnext = t1
        ts.append(tnext)
        if tnext >= t1:
            break
        if len(ts) > max_steps:
            raise RuntimeError("max_steps reached (diffrax throw=True)")
        step = f(tnext - tprev) if rule == "prev_diff" else dt0
        tprev, tnext = tnext, f(tnext + step)
    return np.asarray(ts, dtype=np.float32)


def tsit5_solve_fixed(f, y0: torch.Tensor, step_ts: Sequence[float], save_all: bool = False):
    """Explicit 7-stage Tsit5 with FSAL over a fixed step table.

    k_i = f(t + c_i h, y + h sum_j a_ij k_j);  y1 = y + h sum_i b_i k_i  (b_7 = 0, so
    the 7th stage only feeds FSAL: it is k_1 of the next step).  Returns y(T) or the
    list of y at every step boundary."""
    dt_ = y0.dtype
    y = y0
    ys = [y0]
    k1 = None
    S = len(step_ts) - 1
    for s in range(S):
        t = float(step_ts[s])
        h_f = np.float32(step_ts[s + 1]) - np.float32(step_ts[s]) if dt_ == torch.float32 else float(step_ts[s + 1]) - float(step_ts[s])
        h = float(h_f)
        ks = [k1 if k1 is not None else f(t, y)]
        for i in range(1, 6):
            acc = ks[0] * TSIT5_A[i][0]
            for j in range(1, i):
                acc = acc + ks[j] * TSIT5_A[i][j]
            ks.append(f(t + TSIT5_C[i] * h, y + h * acc))
        acc = ks[0] * TSIT5_B[0]
        for j in range(1, 6):
            acc = acc + ks[j] * TSIT5_B[j]
        y = y + h * acc
        ys.append(y)
        k1 = f(float(step_ts[s + 1]), y) if s + 1 < S else None  # FSAL (7th stage)
    return ys if save_all else y


# --------------------------------------------------------------------------------------
# third-party: diffrax Tsit5 dense output + PIDController + initial step selection
# (call site src/models/graph_neural_cde.py:53-54, 86-104: dt0=None, PIDController(rtol=1e-3,
# atol=1e-6), SaveAt(ts=ts)).  Restated from diffrax (version unpinned) -- parity unpinned.
# --------------------------------------------------------------------------------------


def tsit5_dense_weights(theta: float):
    """b_i(theta), i = 1..7: Tsitouras' 4th-order interpolant in the factored form diffrax's
    ``_Tsit5Interpolation`` uses; y(t + theta h) = y + h sum_i b_i(theta) k_i."""
    t = float(theta)
    b1 = -1.0530884977290216 * t * (t - 1.3299890189751412) * (t * t - 1.4364028541716351 * t + 0.7139816917074209)
    b2 = 0.1017 * t * t * (t * t - 2.1966568338249754 * t + 1.2949852507374631)
    b3 = 2.490627285651252793 * t * t * (t * t - 2.38535645472061657 * t + 1.57803468208092486)
    b4 = -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * t * t
    b5 = 47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * t * t
    b6 = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t * t
    b7 = 2.5 * (t - 1.0) * (t - 0.6) * t * t
    return (b1, b2, b3, b4, b5, b6, b7)


def tsit5_step_full(f, t, h, y, k1):
    """One Tsit5 step -> (y1, y_err, [k1..k7]); k7 = f(t + h, y1) (FSAL)."""
    ks = [k1]
    for i in range(1, 6):
        acc = ks[0] * TSIT5_A[i][0]
        for j in range(1, i):
            acc = acc + ks[j] * TSIT5_A[i][j]
        ks.append(f(t + TSIT5_C[i] * h, y + h * acc))
    acc = ks[0] * TSIT5_B[0]
    for j in range(1, 6):
        acc = acc + ks[j] * TSIT5_B[j]
    y1 = y + h * acc
    ks.append(f(t + h, y1))
    err = ks[0] * TSIT5_BERR[0]
    for j in range(1, 7):
        err = err + ks[j] * TSIT5_BERR[j]
    return y1, h * err, ks


def _rms(x):
    return float(torch.sqrt(torch.mean(x.detach() ** 2)))


def pid_adapt(scaled_error, dt, keep_floor=1.0, safety=0.9, factormin=0.2, factormax=10.0, error_order=5, f=np.float32):
    """diffrax PIDController.adapt_step_size with pcoeff=0, icoeff=1, dcoeff=0 (its defaults, and what
    PIDController(rtol, atol) at graph_neural_cde.py:54 builds): keep = err < 1; factor = clip(safety * err^(-1/5),
    1 if accepted else factormin, factormax)."""
    err = f(scaled_error)
    keep = bool(err < f(1.0))
    with np.errstate(divide="ignore", over="ignore"):
        inv = f(1.0) / err
    if not np.isfinite(inv):
        inv = f(1.0) if np.isnan(inv) else f(np.finfo(np.float32).max)
    factor = f(safety) * f(inv ** f(1.0 / error_order))
    lo = f(keep_floor) if keep else f(factormin)
    factor = min(max(factor, lo), f(factormax))
    return keep, f(f(dt) * factor)


def clip_to_end(tprev, tnext, t1, keep, f=np.float32):
    """diffrax _clip_to_end (non-float64 times: tolerance 1e-6)."""
    if f(tnext) > f(f(t1) - f(1e-6)):
        return f(t1) if keep else f(f(tprev) + f(0.5) * f(f(t1) - f(tprev)))
    return f(tnext)


def select_initial_step(f_vf, t0, y0, f0, rtol, atol, error_order=5, f=np.float32):
    """diffrax _select_initial_step (Hairer, Norsett & Wanner, II.4 'Starting Step Size')."""
    scale = atol + y0.detach().abs() * rtol
    d0, d1 = f(_rms(y0 / scale)), f(_rms(f0 / scale))
    small = d0 < f(1e-5) or d1 < f(1e-5)
    h0 = f(1e-6) if small else f(f(0.01) * f(d0 / d1))
    f1 = f_vf(float(f(t0) + h0), y0 + float(h0) * f0)
    d2 = f(f(_rms((f1 - f0) / scale)) / h0)
    dmax = max(d1, d2)
    h1 = max(f(1e-6), f(h0 * f(1e-3))) if dmax <= f(1e-15) else f((f(0.01) / dmax) ** f(1.0 / error_order))
    return min(f(100.0) * h0, h1)


def tsit5_solve_adaptive(f_vf, y0, t0, t1, rtol=1e-3, atol=1e-6, dt0=None, save_ts=None, max_steps=4096, forced_steps=None):
    """diffeqsolve(ODETerm(f), Tsit5(), t0, t1, dt0, y0, stepsize_controller=PIDController(rtol, atol),
    saveat=SaveAt(ts=save_ts) or SaveAt(t1=True)) with fp32 time arithmetic.  Differentiable with autograd with the
    accepted step sizes held fixed (discretise-then-optimise).  ``forced_steps`` (a step table) bypasses the
    controller: every listed step is accepted -- used to compare implementations on an identical step sequence.
    Returns (ys [M, ...] or y(t1), accepted step table, stats)."""
    f = np.float32
    t0, t1 = f(t0), f(t1)
    y = y0
    k1 = f_vf(float(t0), y)
    if forced_steps is not None:
        dt = f(forced_steps[1]) - f(forced_steps[0])
    elif dt0 is None:
        dt = select_initial_step(f_vf, t0, y.detach(), k1.detach(), rtol, atol)
    else:
        dt = f(dt0)
    M = 0 if save_ts is None else len(save_ts)
    saves, mi, attempts, rejected = [], 0, 0, 0
    boundaries = [t0]
    tprev = t0
    tnext = clip_to_end(tprev, f(tprev + dt), t1, True) if forced_steps is None else f(forced_steps[1])
    while True:
        attempts += 1
        if attempts > max_steps:
            raise RuntimeError("max_steps reached (diffrax throw=True)")
        h = f(tnext - tprev)
        y1, yerr, ks = tsit5_step_full(f_vf, float(tprev), float(h), y, k1)
        if forced_steps is None:
            scale = atol + rtol * torch.maximum(y.detach().abs(), y1.detach().abs())
            keep, dt_new = pid_adapt(_rms(yerr / scale), h)
        else:
            keep = True
        if keep:
            while mi < M and f(save_ts[mi]) <= tnext:
                theta = f(min(max(f(f(save_ts[mi]) - tprev) / h, f(0.0)), f(1.0)))
                w = tsit5_dense_weights(float(theta))
                acc = ks[0] * w[0]
                for i in range(1, 7):
                    acc = acc + ks[i] * w[i]
                saves.append(y + float(h) * acc)
                mi += 1
            y, k1 = y1, ks[6]
            boundaries.append(tnext)
            tprev_new = tnext
        else:
            rejected += 1
            tprev_new = tprev
        if keep and tprev_new >= t1:
            break
        if forced_steps is None:
            tnext = clip_to_end(tprev_new, f(tprev_new + dt_new), t1, keep)
        else:
            tnext = f(forced_steps[len(boundaries)])
        tprev = tprev_new
    stats = {"num_steps": attempts, "num_accepted_steps": attempts - rejected, "num_rejected_steps": rejected}
    out = torch.stack(saves) if M else y
    return out, np.asarray(boundaries, dtype=np.float32), stats


# --------------------------------------------------------------------------------------
# solve wrappers + losses
# --------------------------------------------------------------------------------------


def solve_cde(step_ts, ts, coeffs_adj, x_coeffs, y0, layers, hidden_dim, data_embed_dim, save_all=False):
    """The diffeqsolve call of PGTGraphNeuralCDE.__call__ (pgt_graph_neural_cde.py:101-129),
    from y0 = encoder(x0) to ys[-1]; the encoder/decoder MLPs stay on the host side."""
    control_adj = CubicInterpolation(ts, coeffs_adj)
    if x_coeffs is not None:
        control_data = CubicInterpolation(ts, x_coeffs)
        f = lambda t, y: cde_wrapper_vector_field(t, y, control_adj, control_data, layers, hidden_dim, data_embed_dim)
    else:  # GraphNeuralCDE (graph_neural_cde.py:79-104): ODETerm(vector_field), no wrapper
        f = lambda t, y: perm_equiv_vector_field(t, y, control_adj, layers)
    return tsit5_solve_fixed(f, y0, step_ts, save_all=save_all)


def mlp(x, layers):
    """equinox.nn.MLP / equinox.nn.Linear applied per node (``jax.vmap``): ``layers`` = [(W, b), ...]; ReLU between the
    layers, identity after the last (a one-element list is a plain Linear)."""
    for i, (W, b) in enumerate(layers):
        x = x @ W.t() + b
        if i < len(layers) - 1:
            x = torch.relu(x)
    return x


def solve_cde_dense(step_ts, ts, coeffs_adj, x_coeffs, y0, layers, hidden_dim, data_embed_dim, save_ts):
    """The same call with ``SaveAt(ts=save_ts)`` (evolving_out=True, pgt_graph_neural_cde.py:114-115): fixed step table, every
    save time sampled with Tsit5's interpolant inside the step that holds it."""
    control_adj = CubicInterpolation(ts, coeffs_adj)
    if x_coeffs is not None:
        control_data = CubicInterpolation(ts, x_coeffs)
        f = lambda t, y: cde_wrapper_vector_field(t, y, control_adj, control_data, layers, hidden_dim, data_embed_dim)
    else:
        f = lambda t, y: perm_equiv_vector_field(t, y, control_adj, layers)
    ys, _, _ = tsit5_solve_adaptive(f, y0, float(step_ts[0]), float(step_ts[-1]), save_ts=[float(t) for t in save_ts], forced_steps=step_ts)
    return ys


def pgt_graph_neural_cde(ts, coeffs_adj, x_coeffs, x0, encoder, decoder, vf_layers, hidden_dim, data_embed_dim,
                         global_readout=True, dt0=0.1, evolving_out=False):
    """PGTGraphNeuralCDE.__call__, src/models/pgt_graph_neural_cde.py:78-136 (cubic interpolation, evolving_out=False):
    y0 = vmap(encoder)(x0); diffeqsolve(ODETerm(wrapped_vf), Tsit5, ts[0], ts[-1], dt0=0.1, ConstantStepSize, SaveAt(t1));
    vmap(decoder)(ys[-1]); sum over the nodes when ``global_readout``."""
    y0 = mlp(x0, encoder)
    step_ts = constant_step_table(float(ts[0]), float(ts[-1]), dt0)
    if evolving_out:   # SaveAt(ts=ts); the read-out still takes ys[-1] (:131), i.e. the dense-output sample at ts[-1]
        y_T = solve_cde_dense(step_ts, ts, coeffs_adj, x_coeffs, y0, vf_layers, hidden_dim, data_embed_dim, ts)[-1]
    else:
        y_T = solve_cde(step_ts, ts, coeffs_adj, x_coeffs, y0, vf_layers, hidden_dim, data_embed_dim)
    out = mlp(y_T, decoder)
    return out.sum(dim=0) if global_readout else out


def tgb_graph_neural_cde(ts, coeffs_adj, x_data, x0, encoder, decoder, data_encoder, vf_layers, hidden_dim, data_embed_dim, dt0=0.01,
                         evolving_out=False, return_sequence=False):
    """TGBGraphNeuralCDE.__call__, src/models/tgb_graph_neural_cde.py:96-171 (cubic, evolving_out=False,
    return_sequence=False): the node signal is embedded by ``data_encoder`` (:118), stacked behind a time channel
    (:120-125), turned into Hermite coefficients INSIDE the model (:130) and used as the second control; dt0 = 0.01 (:143)."""
    x_emb = mlp(x_data, [data_encoder])                                             # [T, n, e]
    tsf = ts.to(x_emb.dtype)
    x_path = torch.stack([tsf[:, None, None].expand_as(x_emb), x_emb], dim=-1)       # [T, n, e, 2]
    coeffs_data = backward_hermite_coefficients(tsf, x_path)
    y0 = mlp(x0, encoder)
    step_ts = constant_step_table(float(ts[0]), float(ts[-1]), dt0)
    if evolving_out:
        ys = solve_cde_dense(step_ts, ts, coeffs_adj, coeffs_data, y0, vf_layers, hidden_dim, data_embed_dim, ts)
    else:
        ys = solve_cde(step_ts, ts, coeffs_adj, coeffs_data, y0, vf_layers, hidden_dim, data_embed_dim)[None]
    return mlp(ys if return_sequence else ys[-1], decoder)      # :164-169


def graph_neural_cde(ts, coeffs_adj, x0, initial_linear, final_linear, vf_layers, rtol=1e-3, atol=1e-6):
    """GraphNeuralCDE.__call__, src/models/graph_neural_cde.py:60-113 (cubic, evolving_out=True, return_sequence=True):
    Linear(1 -> h) encoder, ODETerm(vector_field) without the CDE wrapper, dt0=None, PIDController(1e-3, 1e-6),
    SaveAt(ts=ts), Linear(h -> 1) on every saved state.  Returns ([T, n, 1], accepted step table)."""
    y0 = mlp(x0, [initial_linear])
    control_adj = CubicInterpolation(ts, coeffs_adj)
    f = lambda t, y: perm_equiv_vector_field(t, y, control_adj, vf_layers)
    ys, table, _ = tsit5_solve_adaptive(f, y0, float(ts[0]), float(ts[-1]), rtol=rtol, atol=atol, dt0=None,
                                        save_ts=[float(t) for t in ts])
    return mlp(ys, [final_linear]), table


def pgt_mse_loss(y_T, readout_w, label):
    """mse_loss of src/engine/trainer_pgt.py:45-66 with a linear stand-in decoder:
    y_pred = sum_nodes(decoder(y_T)) reshaped (feature_dim, 1) against label [n]
    (broadcast exactly like the reference: (1,1) - (n,) -> (1,n))."""
    out = y_T @ readout_w  # [n, feature_dim]
    y_pred = out.sum(dim=0).reshape(-1, 1)
    return ((y_pred - label) ** 2).mean()


# --------------------------------------------------------------------------------------
# synthetic inputs; shared by tests, smoke and bench
# --------------------------------------------------------------------------------------


def synthetic_graph_path(n: int, T: int, seed: int, scale: float = 1.0, density: Optional[float] = None) -> np.ndarray:
    """A_k [T, n, n] float64: sparse-ish dynamic weighted digraph, Bernoulli(min(1,16/n))
    mask with 10% of entries redrawn per knot, LogNormal(0,1) weights, self-loops."""
    rng = np.random.default_rng(seed)
    p = min(1.0, 16.0 / n) if density is None else density
    mask = rng.random((n, n)) < p
    w = rng.lognormal(0.0, 1.0, size=(n, n))
    out = np.empty((T, n, n), dtype=np.float64)
    for k in range(T):
        if k > 0:
            redraw = rng.random((n, n)) < 0.10
            mask = np.where(redraw, rng.random((n, n)) < p, mask)
            w = np.where(redraw, rng.lognormal(0.0, 1.0, size=(n, n)), w)
        a = np.where(mask, w, 0.0)
        np.fill_diagonal(a, 1.0)
        # row-normalise so the ODE stays O(1) over [0, T-1] (the reference feeds
        # normalised operators for the dynamical systems; PGT feeds raw weights)
        a = a / a.sum(axis=1, keepdims=True)
        out[k] = a * scale
    return out


def reference_layout_coeffs(ts: torch.Tensor, A: torch.Tensor):
    """get_graph_interpolation_coeffs (dataset_configs.py:147-173 / 1073-1100):
    X = stack([t broadcast, A], -1) -> backward_hermite_coefficients -> (d,c,b,a),
    each [T-1, n, n, 2] with the LAST axis interleaved (time, adjacency)."""
    t_index = ts.to(A.dtype)[:, None, None].expand(A.shape[0], A.shape[1], A.shape[2])
    X = torch.stack([t_index, A], dim=-1)
    return backward_hermite_coefficients(ts, X)


def reference_layout_xcoeffs(ts: torch.Tensor, x_t: torch.Tensor):
    """get_interpolation_coeffs for node signals (dataset_configs.py:1073-1100):
    x_t [T, n, e] -> (d,c,b,a) each [T-1, n, e, 2]."""
    t_index = ts.to(x_t.dtype)[:, None, None].expand(x_t.shape[0], x_t.shape[1], x_t.shape[2])
    X = torch.stack([t_index, x_t], dim=-1)
    return backward_hermite_coefficients(ts, X)


@dataclass
class Problem:
    """One graph trajectory worth of hot-path inputs, in the reference's layouts."""

    n: int
    h: int
    e: int  # 0 = no control wrapper
    L: int
    ts: torch.Tensor  # [T]
    coeffs_adj: tuple  # (d,c,b,a) each [T-1,n,n,2]
    x_coeffs: Optional[tuple]  # (d,c,b,a) each [T-1,n,e,2] or None
    y0: torch.Tensor  # [n,h]
    layers: List[LayerParams]
    step_ts: np.ndarray  # [S+1] fp32
    gyT: torch.Tensor  # cotangent [n,h]


def make_problem(n, h, e, L, T, t1, dt0, seed, dtype=torch.float32, scale=1.0, randomize_norm=True, float_ts=False, x_scale=0.3) -> Problem:
    g = torch.Generator().manual_seed(seed + 17)
    if float_ts:
        ts = torch.linspace(0.0, float(t1), T, dtype=torch.float64)
    else:
        ts = torch.arange(T, dtype=torch.float64) * (float(t1) / (T - 1))
    A = torch.from_numpy(synthetic_graph_path(n, T, seed, scale=scale))
    coeffs_adj = tuple(c.to(dtype) for c in reference_layout_coeffs(ts, A))
    x_coeffs = None
    if e > 0:
        x_t = x_scale * torch.randn((T, n, e), generator=g, dtype=torch.float64)
        x_coeffs = tuple(c.to(dtype) for c in reference_layout_xcoeffs(ts, x_t))
    y0 = torch.randn((n, h), generator=g, dtype=torch.float64).to(dtype)
    gyT = torch.randn((n, h), generator=g, dtype=torch.float64).to(dtype)
    widths = layer_widths(h, L, e, e > 0)
    layers = init_params(widths, seed, dtype=dtype, randomize_norm=randomize_norm)
    step_ts = constant_step_table(float(ts[0]), float(ts[-1]), dt0)
    return Problem(n, h, e, L, ts.to(dtype), coeffs_adj, x_coeffs, y0, layers, step_ts, gyT)


def problem_to(p: Problem, dtype) -> Problem:
    return Problem(
        p.n, p.h, p.e, p.L, p.ts.to(dtype), tuple(c.to(dtype) for c in p.coeffs_adj),
        None if p.x_coeffs is None else tuple(c.to(dtype) for c in p.x_coeffs),
        p.y0.to(dtype), params_to(p.layers, dtype), p.step_ts, p.gyT.to(dtype),
    )


def run_forward(p: Problem, save_all=False):
    return solve_cde(p.step_ts, p.ts, p.coeffs_adj, p.x_coeffs, p.y0, p.layers, p.h, p.e, save_all=save_all)


def run_forward_backward(p: Problem):
    """y_T, and the exact discrete-adjoint gradients of <gyT, y_T> w.r.t. y0 and every
    parameter leaf (what eqx.filter_value_and_grad returns through diffrax's default
    RecursiveCheckpointAdjoint)."""
    layers = params_to(p.layers, requires_grad=True)
    y0 = p.y0.detach().clone().requires_grad_(True)
    yT = solve_cde(p.step_ts, p.ts, p.coeffs_adj, p.x_coeffs, y0, layers, p.h, p.e)
    (yT * p.gyT).sum().backward()
    grads = [[t.grad if t.grad is not None else torch.zeros_like(t) for t in lp.tensors()] for lp in layers]
    return yT.detach(), y0.grad, grads
