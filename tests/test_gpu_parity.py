"""GPU parity tests proper: the CUDA path (through the C-ABI, via the host mirror of the reference
API) against the CPU oracle on the same seeded inputs, against the committed goldens, and -- at
BASELINE sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): Z_T within 1e-4 relative (max-norm) of the fp64 oracle in
fp32 mode; gradients within 1e-3 relative per leaf (they are sums of O(steps*stages*n) fp32 terms).
"""
import os

import numpy as np
import pytest
import torch

import perm_equiv_graph_neural_cdes_b200 as P
from perm_equiv_graph_neural_cdes_b200 import _lib
from oracle import reference_path as R
from tests.helpers import FUSED, GOLDEN_CASES, device_model, fused_flags, per_operator, product_grads_as_oracle, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_Y = 1e-4
TOL_G = 1e-3


def test_extension_is_loaded_and_counts_launches(cuda):
    before = _lib.lib().pegncde_launch_count()
    p = R.make_problem(n=12, h=8, e=0, L=2, T=4, t1=3, dt0=0.5, seed=0)
    vf, term, args = device_model(p, cuda)
    P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, p.y0.to(cuda), args)
    torch.cuda.synchronize()
    assert _lib.lib().pegncde_launch_count() > before


@pytest.mark.parametrize("case", ["tiny_nocontrol", "tiny_control", "ragged_n"])
def test_pack_control_matches_reference_layout(cuda, case):
    p = R.make_problem(**GOLDEN_CASES[case])
    xc = p.x_coeffs
    pc = P.pack_control(p.ts.to(cuda), tuple(c.to(cuda) for c in p.coeffs_adj), None if xc is None else tuple(c.to(cuda) for c in xc))
    d, c, b, a = p.coeffs_adj
    ref = torch.stack([a[..., 1], b[..., 1], c[..., 1], d[..., 1]], dim=1)  # [T-1,4,n,n]
    got = pc.dense_planes()[0].cpu()
    assert torch.equal(got, ref)                                  # byte-exact re-layout (un-tiled view)
    assert abs(float(pc.adj_coef[0].double().sum()) - float(ref.double().sum())) < 1e-6 * max(1.0, float(ref.abs().double().sum()))  # zero padding
    assert torch.allclose(pc.adj_rowsum[0].cpu(), ref.sum(-1), rtol=1e-5, atol=1e-6)
    assert torch.equal(pc.adj_diag[0].cpu(), torch.diagonal(ref, dim1=-2, dim2=-1))
    assert torch.allclose(pc.adj_total[0].cpu(), ref.sum((-1, -2)), rtol=1e-5, atol=1e-5)
    tch = torch.stack([b[..., 0].mean(1), c[..., 0].mean(1), d[..., 0].mean(1)], dim=1)
    assert torch.allclose(pc.tch_coef[0].cpu(), tch, rtol=1e-5, atol=1e-6)
    if xc is not None:
        xd, xcc, xb, xa = xc
        refx = torch.stack([xb, xcc, xd], dim=1).reshape(p.ts.numel() - 1, 3, p.n, 2 * p.e)
        assert torch.equal(pc.x_coef[0].cpu(), refx)


def _vf_oracle(p64, t, y):
    ca = R.CubicInterpolation(p64.ts, p64.coeffs_adj)
    if p64.e > 0:
        cx = R.CubicInterpolation(p64.ts, p64.x_coeffs)
        return R.cde_wrapper_vector_field(t, y, ca, cx, p64.layers, p64.h, p64.e)
    return R.perm_equiv_vector_field(t, y, ca, p64.layers)


@pytest.mark.parametrize("flags", [FUSED, 0], ids=["fused-small", "ffma"])
@pytest.mark.parametrize("case", ["tiny_nocontrol", "tiny_control", "ragged_n", "sir_like", "england_like"])
def test_vector_field_forward_and_vjp(cuda, case, flags):
    p = R.make_problem(**GOLDEN_CASES[case])
    p64 = R.problem_to(p, torch.float64)
    vf, term, args = device_model(p, cuda, flags=flags)
    t_lo, t_hi = float(p.ts[0]), float(p.ts[-1])
    # interior points, exact knots, and both out-of-range sides (index clipping)
    times = [t_lo + 0.37 * (t_hi - t_lo), float(p.ts[1]), t_lo, t_hi, t_lo - 0.25, t_hi + 0.25]
    for t in times:
        y = p.y0.to(cuda).requires_grad_(True)
        dy = term(t, y, args)
        layers = R.params_to(p64.layers, requires_grad=True)
        q = R.Problem(p64.n, p64.h, p64.e, p64.L, p64.ts, p64.coeffs_adj, p64.x_coeffs, p64.y0, layers, p64.step_ts, p64.gyT)
        y64 = p64.y0.clone().requires_grad_(True)
        ref = _vf_oracle(q, t, y64)
        assert rel_err(dy.detach(), ref.detach()) < 2e-5, (case, t)
        vf.zero_grad()
        (dy * p.gyT.to(cuda)).sum().backward()
        (ref * p64.gyT).sum().backward()
        assert rel_err(y.grad, y64.grad) < 5e-5, (case, t)
        for l, (got, lp) in enumerate(zip(product_grads_as_oracle(vf), layers)):
            for name, g, r in zip(("fusion", "W", "b", "nw", "nb"), got, lp.tensors()):
                assert rel_err(g, r.grad) < 2e-4, (case, t, l, name)


@pytest.mark.parametrize("flags", [FUSED, 0, _lib.PEG_FLAG_TENSOR_CORES, _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_TF32X3, _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_BF16X2],
                         ids=["fused-small", "ffma", "tcgen05", "tcgen05-tf32x3", "tcgen05-bf16x2"])
@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_solve_against_goldens(cuda, case, flags):
    g = np.load(os.path.join(GOLD, f"{case}.npz"))
    kw = GOLDEN_CASES[case]
    p = R.make_problem(**kw)
    if flags != FUSED and flags and p.n < 128:
        pytest.skip("tensor-core contraction is only selected for n >= 128")
    if flags == FUSED and p.n > 256:
        pytest.skip("the small-graph path covers n <= 256")
    vf, term, args = device_model(p, cuda, flags=flags)
    y0 = p.y0.to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), float(p.ts[0]), float(p.ts[-1]), kw["dt0"], y0, args,
                        stepsize_controller=P.ConstantStepSize(), saveat=P.SaveAt(t1=True))
    assert sol.stats["num_steps"] == int(g["steps"])
    yT = sol.ys[-1]
    slack = 4.0 * float(g["rel32"])  # the reference's own fp32 rounding noise on this problem
    assert rel_err(yT.detach(), g["yT64"]) < TOL_Y + slack, case
    (yT * p.gyT.to(cuda)).sum().backward()
    # gradient tolerance: 1e-3 per leaf (max-norm) on well-conditioned problems.  sir_like is the deliberately stiff case (knots
    # inside every step, |dyT/dy0| ~ 150).  Its exact gradient is piecewise smooth: the fp32 ORACLE reproduces the fp64 one to
    # 2.5e-5 (relative L2, y0 and parameters alike), but a 1e-6 relative perturbation of y0 flips ReLU masks and moves the fp64
    # oracle's own gradient by 3.9e-3 (y0) / 4.1e-3 (parameters) -- measured with oracle/reference_path.py, see DESIGN.md
    # "conditioning".  Two fp32 implementations with different rounding may sit on different sides of such a kink, so the bound
    # is a small multiple of that measured kink size: relative L2 <= 2e-2 for y0 AND for the flat parameter gradient (the
    # per-evaluation VJPs of this case are checked to 5e-5 in test_vector_field_forward_and_vjp[sir_like]).
    if float(g["cond"]) >= 50:
        assert torch.isfinite(y0.grad).all()
        rel_l2 = float((y0.grad.cpu().double() - torch.from_numpy(g["gy0_64"])).norm() / torch.from_numpy(g["gy0_64"]).norm())
        flat = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer]).cpu().double()
        ref = torch.from_numpy(g["gparams64"])
        rel_l2_p = float((flat - ref).norm() / ref.norm())
        print(f"\n[{case}] stiff-case gradient error, relative L2: y0 {rel_l2:.2e}, parameters {rel_l2_p:.2e}")
        assert rel_l2 < 2e-2 and rel_l2_p < 2e-2, (rel_l2, rel_l2_p)
        return
    tol_g = 2e-3 if (flags != FUSED and flags & _lib.PEG_FLAG_BF16X2) else TOL_G    # bf16x2: the separately stated looser-tolerance operand format
    assert rel_err(y0.grad, g["gy0_64"]) < tol_g, case
    flat = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer]).cpu().double().numpy()
    ref = g["gparams64"]
    # per-leaf relative error
    off = 0
    for layer in vf.gnn_layers:
        cl = layer.conv_layer
        for name, numel in (("fusion", 16), ("W", cl.linear.weight.numel()), ("b", cl.linear.bias.numel()),
                            ("nw", cl.norm.weight.numel()), ("nb", cl.norm.bias.numel())):
            a, b = flat[off:off + numel], ref[off:off + numel]
            assert np.abs(a - b).max() <= tol_g * max(np.abs(b).max(), 1e-12) + 1e-7, (case, name)
            off += numel


def test_save_steps_matches_oracle_trajectory(cuda):
    p = R.make_problem(n=20, h=8, e=3, L=3, T=4, t1=3, dt0=0.25, seed=1)
    ys_ref = torch.stack(R.run_forward(R.problem_to(p, torch.float64), save_all=True))
    vf, term, args = device_model(p, cuda)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.25, p.y0.to(cuda), args, saveat=P.SaveAt(steps=True))
    assert sol.ys.shape == ys_ref.shape
    assert rel_err(sol.ys, ys_ref) < TOL_Y
    # cotangents injected at every saved boundary
    y0 = p.y0.to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.25, y0, args, saveat=P.SaveAt(steps=True))
    w = torch.linspace(0.5, 1.5, sol.ys.shape[0], device=cuda)[:, None, None]
    (sol.ys * w).sum().backward()
    p64 = R.problem_to(p, torch.float64)
    y64 = p64.y0.clone().requires_grad_(True)
    ys = torch.stack(R.solve_cde(p64.step_ts, p64.ts, p64.coeffs_adj, p64.x_coeffs, y64, p64.layers, p64.h, p64.e, save_all=True))
    (ys * w.cpu().double()).sum().backward()
    assert rel_err(y0.grad, y64.grad) < TOL_G


def test_batch_of_independent_graphs(cuda):
    """B graphs with different control paths in one call == each solved alone (the reference's jax.vmap)."""
    ps = [R.make_problem(n=24, h=8, e=2, L=2, T=4, t1=3, dt0=0.5, seed=s) for s in (0, 1, 3)]
    base = ps[0]
    vf, term, _ = device_model(base, cuda)
    ts = base.ts.to(cuda)
    cadj = tuple(torch.stack([p.coeffs_adj[i] for p in ps]).to(cuda) for i in range(4))
    cx = tuple(torch.stack([p.x_coeffs[i] for p in ps]).to(cuda) for i in range(4))
    y0 = torch.stack([p.y0 for p in ps]).to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, y0, [P.CubicInterpolation(ts, cadj), P.CubicInterpolation(ts, cx)])
    gy = torch.stack([p.gyT for p in ps]).to(cuda)
    (sol.ys[-1] * gy).sum().backward()
    batched_grad = vf.flat_params().detach().clone() * 0
    batched_grad = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
    acc = torch.zeros_like(batched_grad)
    for i, p in enumerate(ps):
        vf.zero_grad()
        yi = p.y0.to(cuda).requires_grad_(True)
        args = [P.CubicInterpolation(ts, tuple(c.to(cuda) for c in p.coeffs_adj)), P.CubicInterpolation(ts, tuple(c.to(cuda) for c in p.x_coeffs))]
        si = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, yi, args)
        assert rel_err(sol.ys[-1][i].detach(), si.ys[-1].detach()) < 1e-6
        (si.ys[-1] * p.gyT.to(cuda)).sum().backward()
        assert rel_err(y0.grad[i], yi.grad) < 1e-5
        acc += torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
    assert rel_err(batched_grad, acc) < 1e-4  # parameter gradients are summed over the batch


def test_tsit5_step_and_error_estimate(cuda):
    p = R.make_problem(n=20, h=8, e=3, L=3, T=4, t1=3, dt0=0.1, seed=1)
    p64 = R.problem_to(p, torch.float64)
    vf, term, args = device_model(p, cuda)
    t, dt = 0.4, 0.3
    y1, yerr, k7 = P.tsit5_step(P.ODETerm(term), t, dt, p.y0.to(cuda), args)
    f = lambda tt, y: _vf_oracle(p64, tt, y)
    ks = [f(t, p64.y0)]
    for i in range(1, 7):
        acc = sum(ks[j] * R.TSIT5_A[i][j] for j in range(i))
        ks.append(f(t + R.TSIT5_C[i] * dt, p64.y0 + dt * acc))
    y1_ref = p64.y0 + dt * sum(ks[j] * R.TSIT5_B[j] for j in range(6))
    err_ref = dt * sum(ks[j] * R.TSIT5_BERR[j] for j in range(7))
    assert rel_err(y1, y1_ref) < 2e-5
    assert rel_err(k7, ks[6]) < 5e-5   # k7 = f(t+dt, y1) inherits (and amplifies) the rounding of y1
    assert rel_err(k7, _vf_oracle(p64, t + dt, y1.cpu().double())) < 2e-6   # the evaluation itself is exact
    assert float((yerr.cpu().double() - err_ref).abs().max()) < 1e-5 * float(y1_ref.abs().max())
    # FSAL: feeding k7 back as k1 of the next step reproduces a fresh evaluation
    y2a, _, _ = P.tsit5_step(P.ODETerm(term), t + dt, dt, y1, args, k1=k7)
    y2b, _, _ = P.tsit5_step(P.ODETerm(term), t + dt, dt, y1, args)
    assert rel_err(y2a, y2b) < 1e-6


def test_workspace_too_small_is_reported(cuda):
    p = R.make_problem(n=12, h=8, e=0, L=2, T=4, t1=3, dt0=0.5, seed=0)
    vf, term, args = device_model(p, cuda)
    pc = P.vector_field.resolve_control(args, None, cuda)
    dims = vf.dims_for(pc, with_wrapper=False)
    y = p.y0.to(cuda).unsqueeze(0).contiguous()
    dy = torch.empty_like(y)
    ws = torch.empty(256, dtype=torch.uint8, device=cuda)
    flat = vf.flat_params().detach()
    rc = _lib.lib().pegncde_vf_fwd(torch.cuda.current_stream().cuda_stream, dims, pc.struct(), flat.data_ptr(), 0.5,
                                   y.data_ptr(), dy.data_ptr(), ws.data_ptr(), ws.numel())
    assert rc == 3
    with pytest.raises(P.PegError):
        _lib.check(rc, "vf_fwd")


# ------------------------------------------------------------------------------------------------
# BASELINE-size properties (no CPU oracle at these sizes)
# ------------------------------------------------------------------------------------------------
def _device_problem(n, h, e, L, T, seed, device, B=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.from_numpy(R.synthetic_graph_path(n, T, seed)).to(torch.float32)
    ts = torch.arange(T, dtype=torch.float32)
    co = tuple(c.to(device) for c in R.reference_layout_coeffs(ts, A))
    xco = None
    if e > 0:
        x_t = 0.3 * torch.randn((T, n, e), generator=g)
        xco = tuple(c.to(device) for c in R.reference_layout_xcoeffs(ts, x_t))
    widths = R.layer_widths(h, L, e, e > 0)
    vf = P.PermEquivGraphVectorField(h, h, widths[-1], L, e, n, key=seed).to(device)
    y0 = torch.randn((n, h), generator=g).to(device)
    return ts.to(device), co, xco, vf, y0


@pytest.mark.parametrize("n,h,e", [(1000, 64, 16), (1000, 64, 0)])
def test_permutation_equivariance_at_twitter_size(cuda, n, h, e):
    """f(P Z, P A P^T, P X) = P f(Z, A, X) at C4 (Twitter) size: n=1000, h=64, e=16, L=3."""
    ts, co, xco, vf, y0 = _device_problem(n, h, e, 3, 3, 11, cuda)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5)).to(cuda)
    term = P.CDEWrapperVectorField(vf, h) if e > 0 else vf
    def run(co_, xco_, y_):
        ca = P.CubicInterpolation(ts, co_)
        args = [ca, P.CubicInterpolation(ts, xco_)] if e > 0 else ca
        return P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 2.0, 0.25, y_, args).ys[-1]
    out = run(co, xco, y0)
    cop = tuple(c[:, perm][:, :, perm].contiguous() for c in co)
    xcop = None if xco is None else tuple(c[:, perm].contiguous() for c in xco)
    outp = run(cop, xcop, y0[perm].contiguous())
    assert torch.isfinite(out).all()
    assert rel_err(outp, out[perm]) < 5e-5


TC = _lib.PEG_FLAG_TENSOR_CORES
FAST = _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_TF32_FAST
X3 = _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_TF32X3


@pytest.mark.parametrize("flags", [0, TC, X3, _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_BF16X2], ids=["ffma", "tcgen05", "tcgen05-tf32x3", "tcgen05-bf16x2"])
@pytest.mark.parametrize("n,h,e,L", [(1000, 64, 16, 3), (1000, 64, 0, 3), (515, 32, 0, 2), (300, 32, 3, 2), (129, 64, 8, 3)])
def test_vector_field_and_vjp_at_twitter_size_against_oracle(cuda, n, h, e, L, flags):
    """One evaluation + VJP at C4 (Twitter) size against the fp64 oracle (a single evaluation is cheap on CPU).
    flags=tcgen05 runs the n x n x d contractions on the tensor cores (3xTF32 split: same fp32 tolerance)."""
    p = R.make_problem(n=n, h=h, e=e, L=L, T=3, t1=2, dt0=0.5, seed=21)
    p64 = R.problem_to(p, torch.float64)
    vf, term, args = device_model(p, cuda, flags=flags)
    t = 1.3
    y = p.y0.to(cuda).requires_grad_(True)
    dy = term(t, y, args)
    (dy * p.gyT.to(cuda)).sum().backward()
    layers = R.params_to(p64.layers, requires_grad=True)
    q = R.Problem(p64.n, p64.h, p64.e, p64.L, p64.ts, p64.coeffs_adj, p64.x_coeffs, p64.y0, layers, p64.step_ts, p64.gyT)
    y64 = p64.y0.clone().requires_grad_(True)
    ref = _vf_oracle(q, t, y64)
    (ref * p64.gyT).sum().backward()
    assert rel_err(dy.detach(), ref.detach()) < 2e-5
    assert rel_err(y.grad, y64.grad) < 5e-5
    for l, (got, lp) in enumerate(zip(product_grads_as_oracle(vf), layers)):
        for name, g, r in zip(("fusion", "W", "b", "nw", "nb"), got, lp.tensors()):
            assert rel_err(g, r.grad) < 3e-4, (l, name)


def test_tf32_fast_mode_has_its_own_looser_tolerance(cuda):
    """PEG_FLAG_TF32_FAST (single-pass TF32, rna-rounded operands): stated tolerance 5e-3 on Z_T."""
    g = np.load(os.path.join(GOLD, "england_like.npz"))
    p = R.make_problem(**GOLDEN_CASES["england_like"])
    vf, term, args = device_model(p, cuda, flags=FAST)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.1, p.y0.to(cuda), args)
    err = rel_err(sol.ys[-1], g["yT64"])
    assert err < 5e-3, err


def test_stage_store_and_recompute_adjoints_agree(cuda):
    """solve_bwd with the forward's stored stage inputs (no recompute) == checkpoint-per-step + recompute."""
    p = R.make_problem(n=40, h=16, e=2, L=3, T=4, t1=3, dt0=0.25, seed=0)
    outs = []
    for store in (True, False):
        vf, term, args = device_model(p, cuda)
        vf.store_stages = store
        y0 = p.y0.to(cuda).requires_grad_(True)
        sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.25, y0, args)
        (sol.ys[-1] * p.gyT.to(cuda)).sum().backward()
        outs.append((y0.grad.clone(), torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])))
    assert rel_err(outs[0][0], outs[1][0]) < 1e-6
    assert rel_err(outs[0][1], outs[1][1]) < 1e-5


# ---------------------------------------------------------------------------------------------------
# adaptive path (configs 0 / 1 of BASELINE.json: GraphNeuralCDE, dt0=None, PIDController(1e-3, 1e-6), SaveAt(ts=ts))
# ---------------------------------------------------------------------------------------------------
def _oracle_vf(p):
    ca = R.CubicInterpolation(p.ts, p.coeffs_adj)
    if p.e > 0:
        cx = R.CubicInterpolation(p.ts, p.x_coeffs)
        return lambda t, y: R.cde_wrapper_vector_field(t, y, ca, cx, p.layers, p.h, p.e)
    return lambda t, y: R.perm_equiv_vector_field(t, y, ca, p.layers)


@pytest.mark.parametrize("n,h,e,L,T", [(60, 16, 0, 2, 12), (129, 32, 2, 2, 6)])
def test_adaptive_solve_with_dense_output_against_oracle(cuda, n, h, e, L, T):
    """PIDController + SaveAt(ts): (i) the CUDA path and the fp32 oracle, each running its own controller, accept
    (almost) the same step sequence; (ii) with the oracle FORCED onto the CUDA path's accepted step table (fp64), the
    dense-output samples and the exact discrete-adjoint gradients agree to the fixed-step tolerances."""
    p = R.make_problem(n=n, h=h, e=e, L=L, T=T, t1=4, dt0=0.1, seed=31)
    vf, term, args = device_model(p, cuda)
    t0, t1 = float(p.ts[0]), float(p.ts[-1])
    save_ts = p.ts.to(torch.float32)
    y0 = p.y0.to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), t0, t1, None, y0, args, stepsize_controller=P.PIDController(rtol=1e-3, atol=1e-6),
                        saveat=P.SaveAt(ts=save_ts))
    assert sol.ys.shape == (T, n, h) and torch.isfinite(sol.ys).all()
    table = sol.stats["step_ts"]
    assert sol.stats["num_accepted_steps"] == len(table) - 1 and table[0] == np.float32(t0) and table[-1] == np.float32(t1)
    G = torch.randn(sol.ys.shape, generator=torch.Generator().manual_seed(3))
    (sol.ys * G.to(cuda)).sum().backward()

    # (i) independent controllers: fp32 oracle
    ys32, table32, stats32 = R.tsit5_solve_adaptive(_oracle_vf(p), p.y0, t0, t1, save_ts=save_ts.numpy())
    # accept / reject decisions at scaled error ~ 1 are rounding-sensitive: counts agree approximately, not exactly
    assert abs(stats32["num_accepted_steps"] - sol.stats["num_accepted_steps"]) <= max(1, sol.stats["num_accepted_steps"] // 4)
    assert abs(stats32["num_steps"] - sol.stats["num_steps"]) <= max(2, sol.stats["num_steps"] // 4)
    assert abs(float(table32[1]) - float(table[1])) < 1e-3 * float(table[1])      # initial step-size heuristic
    assert rel_err(sol.ys, ys32) < 2e-2                                           # two rtol=1e-3 solves on (slightly) different step tables

    # (ii) same accepted steps, fp64 truth
    p64 = R.problem_to(p, torch.float64)
    layers = R.params_to(p64.layers, requires_grad=True)
    q = R.Problem(p64.n, p64.h, p64.e, p64.L, p64.ts, p64.coeffs_adj, p64.x_coeffs, p64.y0, layers, p64.step_ts, p64.gyT)
    y64 = p64.y0.clone().requires_grad_(True)
    ys64, table64, _ = R.tsit5_solve_adaptive(_oracle_vf(q), y64, t0, t1, save_ts=save_ts.numpy(), forced_steps=table)
    assert np.array_equal(table64, table)
    assert rel_err(sol.ys, ys64) < TOL_Y
    assert rel_err(sol.ys[0], p.y0) < 1e-6                                        # theta = 0 sample is y0
    (ys64 * G.double()).sum().backward()
    assert rel_err(y0.grad, y64.grad) < TOL_G
    for l, (got, lp) in enumerate(zip(product_grads_as_oracle(vf), layers)):
        for name, g, r in zip(("fusion", "W", "b", "nw", "nb"), got, lp.tensors()):
            assert rel_err(g, r.grad) < TOL_G, (l, name)


def test_adaptive_batch_steps_every_trajectory_on_its_own(cuda):
    """jax.vmap(model) over trajectories (loss_configs.py:44): each has its own accepted-step sequence; SaveAt(t1)."""
    ps = [R.make_problem(n=40, h=16, e=0, L=2, T=6, t1=3, dt0=0.1, seed=s, scale=sc) for s, sc in ((41, 1.0), (42, 3.0))]
    vf, term, _ = device_model(ps[0], cuda)
    ts = ps[0].ts.to(torch.float32).to(cuda)
    co = tuple(torch.stack([p.coeffs_adj[i] for p in ps]).to(cuda) for i in range(4))
    y0 = torch.stack([p.y0 for p in ps]).to(cuda)
    ctrl = P.PIDController(rtol=1e-3, atol=1e-6)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, None, y0, P.CubicInterpolation(ts, co), stepsize_controller=ctrl)
    assert sol.ys.shape == (1, 2, 40, 16)
    assert sol.stats["num_steps"][0] != sol.stats["num_steps"][1]      # the stiffer graph needs more steps
    for b, p in enumerate(ps):
        q = R.Problem(p.n, p.h, p.e, p.L, p.ts, p.coeffs_adj, None, p.y0, ps[0].layers, p.step_ts, p.gyT)
        # the batched call against the fp64 oracle forced onto trajectory b's own accepted step table
        yT, _, _ = R.tsit5_solve_adaptive(_oracle_vf(R.problem_to(q, torch.float64)), q.y0.double(), 0.0, 3.0,
                                          forced_steps=sol.stats["step_ts"][b])
        assert rel_err(sol.ys[0, b], yT) < TOL_Y
        # ... and against the same trajectory solved alone: bit-identical (the whole path, pack pre-pass included, is
        # deterministic, so every accept / reject decision repeats)
        single = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, None, y0[b], P.CubicInterpolation(ts, tuple(c[b] for c in co)),
                               stepsize_controller=ctrl)
        assert single.stats["num_steps"] == sol.stats["num_steps"][b]
        assert torch.equal(single.ys[0], sol.ys[0, b])


def test_graph_neural_cde_model_matches_reference_call(cuda):
    """model(ts, coeffs_adj, x0, evolving_out=True) of graph_neural_cde.py:60-113 -> [T, n, 1]."""
    p = R.make_problem(n=50, h=16, e=0, L=2, T=8, t1=5, dt0=0.1, seed=7, float_ts=True)
    vf, _, _ = device_model(p, cuda)
    model = P.GraphNeuralCDE(16, vf, seed=3).to(cuda)
    x0 = torch.randn(50, 1, generator=torch.Generator().manual_seed(1)).to(cuda)
    out = model(p.ts.to(torch.float32).to(cuda), tuple(c.to(cuda) for c in p.coeffs_adj), x0)
    assert out.shape == (8, 50, 1) and torch.isfinite(out).all()
    out.square().mean().backward()
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in model.parameters())


# ---------------------------------------------------------------------------------------------------
# learned node-signal control (TGB models): cotangent of the x coefficients through the solve
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, TC], ids=["ffma", "tcgen05"])
def test_solve_gradient_wrt_node_signal_coefficients(cuda, flags):
    """tgb_graph_neural_cde.py:118-137 builds coeffs_data inside the model, so d loss / d x_coeffs must come out of the
    solve: pegncde_solve_bwd's g_xcoef against autograd through the fp64 oracle (b, c, d arrays; `a` is unused by X')."""
    p = R.make_problem(n=130 if flags else 37, h=32 if flags else 16, e=2, L=2, T=5, t1=2, dt0=0.1, seed=13)
    vf, term, _ = device_model(p, cuda, flags=flags)
    ts = p.ts.to(torch.float32).to(cuda)
    xco = tuple(c.to(cuda).requires_grad_(True) for c in p.x_coeffs)
    args = [P.CubicInterpolation(ts, tuple(c.to(cuda) for c in p.coeffs_adj)), P.CubicInterpolation(ts, xco)]
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 2.0, 0.1, p.y0.to(cuda), args)
    (sol.ys[-1] * p.gyT.to(cuda)).sum().backward()
    p64 = R.problem_to(p, torch.float64)
    x64 = tuple(c.clone().requires_grad_(True) for c in p64.x_coeffs)
    yT = R.solve_cde(p64.step_ts, p64.ts, p64.coeffs_adj, x64, p64.y0, p64.layers, p64.h, p64.e)
    (yT * p64.gyT).sum().backward()
    assert rel_err(sol.ys[-1], yT) < TOL_Y
    for name, got, ref in zip("dcb", xco[:3], x64[:3]):
        # the time channel [..., 0] of the coefficients does not enter X' of the data channel
        assert rel_err(got.grad, ref.grad) < TOL_G, name
    assert xco[3].grad is None or float(xco[3].grad.abs().max()) == 0.0


def test_tgb_model_trains_its_data_encoder(cuda):
    """model(ts, coeffs_adj, x_data, x0, start_time) of tgb_graph_neural_cde.py:96-171: gradients reach data_encoder."""
    n, h, e, T = 48, 16, 4, 5
    p = R.make_problem(n=n, h=h, e=e, L=2, T=T, t1=T - 1, dt0=0.1, seed=23)
    vf, _, _ = device_model(p, cuda)
    model = P.TGBGraphNeuralCDE(h, vf, use_mlps=True, seed=2, dt0=0.05).to(cuda)
    g = torch.Generator().manual_seed(9)
    x_data = torch.randn(T, n, n, generator=g).to(cuda)
    x0 = torch.randn(n, n, generator=g).to(cuda)
    out = model(torch.arange(T, device=cuda), tuple(c.to(cuda) for c in p.coeffs_adj), x_data, x0, None)
    assert out.shape == (n, n) and torch.isfinite(out).all()
    out.square().mean().backward()
    for name, q in model.named_parameters():
        assert q.grad is not None and torch.isfinite(q.grad).all(), name
    assert float(model.data_encoder.weight.grad.abs().max()) > 0
    # finite-difference check of one data-encoder weight through the whole model
    w = model.data_encoder.weight
    idx = (1, 3)
    ga = float(w.grad[idx])
    def loss():
        with torch.no_grad():
            return float(model(torch.arange(T, device=cuda), tuple(c.to(cuda) for c in p.coeffs_adj), x_data, x0, None).double().square().mean())
    eps = 2e-2
    with torch.no_grad():
        w[idx] += eps; lp = loss(); w[idx] -= 2 * eps; lm = loss(); w[idx] += eps
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - ga) < 5e-2 * max(abs(fd), abs(ga), 1e-4), (fd, ga)


# ---------------------------------------------------------------------------------------------------
# sibling vector fields through the same kernels: GraphVectorField (A + A'), GNODEVectorField (A)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls,with_derivative", [("GraphVectorField", True), ("GNODEVectorField", False)])
@pytest.mark.parametrize("flags", [0, TC], ids=["ffma", "tcgen05"])
def test_sibling_vector_fields_against_oracle(cuda, cls, with_derivative, flags):
    # Few steps on purpose: the exact gradient of a ReLU network jumps when a hidden unit sits within rounding distance
    # of its kink (seed 17 with 12 steps has one such unit: a 1e-7 perturbation of y0 moves d loss / d y0 by 7e-3 in
    # the fp32 CUDA-core path too), and the chance of meeting one grows with units x stages x steps.
    p = R.make_problem(n=140 if flags else 33, h=32, e=0, L=3, T=4, t1=3, dt0=0.75, seed=19)
    vf = getattr(P, cls)(p.h, p.h, p.h, p.L, 0, p.n, key=0)
    with torch.no_grad():
        for mine, lp in zip(vf.gnn_layers, p.layers):
            mine.linear.weight.copy_(lp.weight); mine.linear.bias.copy_(lp.bias)
            mine.norm.weight.copy_(lp.norm_weight); mine.norm.bias.copy_(lp.norm_bias)
    vf = vf.to(cuda)
    vf.flags = per_operator(flags)
    assert sorted(k for k, _ in vf.named_parameters())[:2] == ["gnn_layers.0.linear.bias", "gnn_layers.0.linear.weight"]
    ts = p.ts.to(torch.float32).to(cuda)
    ca = P.CubicInterpolation(ts, tuple(c.to(cuda) for c in p.coeffs_adj))
    y0 = p.y0.to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 3.0, 0.75, y0, ca)
    (sol.ys[-1] * p.gyT.to(cuda)).sum().backward()
    p64 = R.problem_to(p, torch.float64)
    layers = R.params_to(p64.layers, requires_grad=True)
    c64 = R.CubicInterpolation(p64.ts, p64.coeffs_adj)
    y64 = p64.y0.clone().requires_grad_(True)
    yT = R.tsit5_solve_fixed(lambda t, y: R.plain_graph_vector_field(t, y, c64, layers, with_derivative), y64, R.constant_step_table(0.0, 3.0, 0.75))
    (yT * p64.gyT).sum().backward()
    assert rel_err(sol.ys[-1], yT) < TOL_Y
    assert rel_err(y0.grad, y64.grad) < TOL_G
    for mine, lp in zip(vf.gnn_layers, layers):
        assert rel_err(mine.linear.weight.grad, lp.weight.grad) < TOL_G
        assert rel_err(mine.linear.bias.grad, lp.bias.grad) < TOL_G
        assert rel_err(mine.norm.weight.grad, lp.norm_weight.grad) < TOL_G
        assert rel_err(mine.norm.bias.grad, lp.norm_bias.grad) < TOL_G


# ---------------------------------------------------------------------------------------------------
# control-path builder on the device: snapshots -> planes, Hermite coefficients fused
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,T,uniform", [(37, 5, True), (130, 4, False)])
def test_build_control_from_snapshots_matches_pack_of_host_coefficients(cuda, n, T, uniform):
    """pegncde_build_adj(A_k) == pegncde_pack_adj(backward_hermite_coefficients(stack([t, A_k]))) -- bit for bit on the planes
    (same fp32 operation order), and the solve on top agrees."""
    g = torch.Generator().manual_seed(n)
    ts = torch.arange(T, dtype=torch.float32) if uniform else torch.tensor([0.0, 0.7, 1.1, 2.5])
    A = torch.from_numpy(R.synthetic_graph_path(n, T, seed=n)).to(torch.float32)
    x_t = 0.3 * torch.randn(T, n, 2, generator=g)
    co = R.reference_layout_coeffs(ts, A)
    xco = R.reference_layout_xcoeffs(ts, x_t)
    ref = P.pack_control(ts.to(cuda), tuple(c.to(cuda) for c in co), tuple(c.to(cuda) for c in xco))
    got = P.build_control(ts.to(cuda), A.to(cuda), x_t.to(cuda))
    assert torch.equal(got.adj_coef, ref.adj_coef)
    assert torch.equal(got.adj_diag, ref.adj_diag)
    assert torch.allclose(got.adj_rowsum, ref.adj_rowsum, rtol=1e-5, atol=1e-6)
    assert torch.allclose(got.adj_total, ref.adj_total, rtol=1e-5, atol=1e-5)
    assert torch.allclose(got.tch_coef, ref.tch_coef, rtol=0, atol=1e-6)
    assert torch.equal(got.x_coef, ref.x_coef)
    # batched + the solve accepts the built control directly
    gb = P.build_control(ts.to(cuda), torch.stack([A, A]).to(cuda))
    assert torch.equal(gb.adj_coef[1], ref.adj_coef[0]) and gb.B == 2


def test_adaptive_solve_at_c1_heat_shape(cuda):
    """BASELINE.json configs[0] shape (configs/dynamical_systems/perm_equiv_gncde_config.yaml: n=400, hidden 16, 2 layers,
    knots every 5/79 on the time axis, no control wrapper, PIDController(1e-3, 1e-6), SaveAt(ts=ts)), first 40 knots:
    forward + gradient against the fp64 oracle forced onto the accepted step table.  A hundred-odd steps x 6 stages x 6400
    ReLU units make the exact gradient kink-sensitive (DESIGN.md "Conditioning": a 1e-6 relative perturbation of y0 moves the
    fp64 oracle's own gradient by 2.6e-3 in max-norm here), so the gradient is compared in the 2-norm."""
    full = R.make_problem(n=400, h=16, e=0, L=2, T=80, t1=5, dt0=0.1, seed=101, float_ts=True)
    T = 40
    p = R.Problem(full.n, full.h, 0, full.L, full.ts[:T], tuple(c[:T - 1] for c in full.coeffs_adj), None, full.y0, full.layers,
                  full.step_ts, full.gyT)
    t1 = float(p.ts[-1])
    vf, term, args = device_model(p, cuda)
    save_ts = p.ts.to(torch.float32)
    y0 = p.y0.to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, t1, None, y0, args, stepsize_controller=P.PIDController(rtol=1e-3, atol=1e-6),
                        saveat=P.SaveAt(ts=save_ts))
    assert sol.ys.shape == (T, 400, 16)
    G = torch.randn(sol.ys.shape, generator=torch.Generator().manual_seed(5)) / T
    (sol.ys * G.to(cuda)).sum().backward()
    p64 = R.problem_to(p, torch.float64)

    def oracle_grad(y_start):
        y = y_start.clone().requires_grad_(True)
        ys, _, _ = R.tsit5_solve_adaptive(_oracle_vf(p64), y, 0.0, t1, save_ts=save_ts.numpy(), forced_steps=sol.stats["step_ts"])
        (ys * G.double()).sum().backward()
        return ys.detach(), y.grad

    ys64, g64 = oracle_grad(p64.y0)
    assert rel_err(sol.ys, ys64) < TOL_Y
    # a flipped unit moves a few entries of the gradient by ~1e-2 of its max (and which unit flips depends on the last bit
    # of the BLAS summation order on the host): the 2-norm error stays tight, the max-norm error gets the loose bound
    g = y0.grad.detach().double().cpu()
    assert float((g - g64).norm() / g64.norm()) < 5e-3
    assert rel_err(g, g64) < 3e-2


# ---------------------------------------------------------------------------------------------------
# streamed control: host coefficient arrays, copy + pack of piece i+1 overlapped with the steps inside piece i
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["england_like", "sir_like", "ragged_n"])
def test_streamed_host_control_is_bit_identical_to_the_resident_one(cuda, case, monkeypatch):
    """pack_control on HOST tensors defers the adjacency planes; the fixed-step solve packs them piece by piece
    (pegncde_pack_adj_range) and runs the step table in segments.  Same bits as the all-at-once path, forward and backward
    (sir_like has knots inside the steps, ragged_n a partial last tile)."""
    p = R.make_problem(**GOLDEN_CASES[case])
    kw = GOLDEN_CASES[case]
    outs = []
    monkeypatch.setattr(P.solve, "STREAM_MIN_PIECE_BYTES", 0)     # these cases are tiny: force the streamed path
    for streamed in (False, True):
        vf, term, _ = device_model(p, cuda, flags=TC if p.n >= 128 else 0)
        place = (lambda t: t.to(torch.float32)) if streamed else (lambda t: t.to(torch.float32).to(cuda))
        ts = p.ts.to(torch.float32)
        cadj = P.CubicInterpolation(place(ts), tuple(place(c) for c in p.coeffs_adj))
        args = [cadj, P.CubicInterpolation(place(ts), tuple(place(c) for c in p.x_coeffs))] if p.e > 0 else cadj
        y0 = p.y0.to(cuda).requires_grad_(True)
        sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), float(p.ts[0]), float(p.ts[-1]), kw["dt0"], y0, args, saveat=P.SaveAt(steps=True))
        if streamed:
            assert cadj._packed.pending is None        # every piece has been packed by the end of the forward solve
        (sol.ys[-1] * p.gyT.to(cuda)).sum().backward()
        outs.append((sol.ys.detach().clone(), y0.grad.clone(), cadj._packed.adj_coef.clone(), cadj._packed.adj_rowsum.clone(),
                     cadj._packed.tch_coef.clone()))
    a, b = outs
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    assert torch.equal(a[0], b[0])
    assert rel_err(b[1], a[1]) < 1e-6      # parameter-gradient atomics aside, the adjoint sees identical inputs


def test_streamed_control_materialises_for_adaptive_and_single_evaluations(cuda):
    p = R.make_problem(n=40, h=16, e=0, L=2, T=5, t1=2, dt0=0.1, seed=3)
    vf, term, _ = device_model(p, cuda)
    ts = p.ts.to(torch.float32)
    host = P.CubicInterpolation(ts, tuple(c.to(torch.float32) for c in p.coeffs_adj))
    dev = P.CubicInterpolation(ts.to(cuda), tuple(c.to(torch.float32).to(cuda) for c in p.coeffs_adj))
    y = p.y0.to(cuda)
    assert torch.equal(term(0.7, y, host), term(0.7, y, dev))
    ctrl = P.PIDController(rtol=1e-3, atol=1e-6)
    host2 = P.CubicInterpolation(ts, tuple(c.to(torch.float32) for c in p.coeffs_adj))
    s1 = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 2.0, None, y, host2, stepsize_controller=ctrl)
    s2 = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 2.0, None, y, dev, stepsize_controller=ctrl)
    assert torch.equal(s1.ys, s2.ys)


# ---------------------------------------------------------------------------------------------------
# directed equivariant field: ConvEquivFusionDirectedLayer, 11 parameter pairs, row and column sums
# ---------------------------------------------------------------------------------------------------
DIR_FIELDS = ("param1", "param2", "param3", "param4", "param4_prime", "param5", "param5_prime", "param6", "param6_prime", "param7", "param8")


@pytest.mark.parametrize("flags", [0, TC], ids=["ffma", "tcgen05"])
def test_directed_vector_field_against_oracle(cuda, flags):
    p = R.make_problem(n=140 if flags else 33, h=32, e=0, L=2, T=4, t1=3, dt0=0.75, seed=29)
    vf = P.PermEquivDirGraphVectorField(p.h, p.h, p.h, p.L, 0, p.n, key=4)
    with torch.no_grad():
        for mine, lp in zip(vf.gnn_layers, p.layers):
            mine.conv_layer.linear.weight.copy_(lp.weight); mine.conv_layer.linear.bias.copy_(lp.bias)
            mine.conv_layer.norm.weight.copy_(lp.norm_weight); mine.conv_layer.norm.bias.copy_(lp.norm_bias)
    fus64 = [torch.stack([getattr(m, f).detach().double() for f in DIR_FIELDS]).requires_grad_(True) for m in vf.gnn_layers]
    vf = vf.to(cuda)
    vf.flags = per_operator(flags)
    assert vf.flat_params().numel() == sum(32 * 32 + 32 + 64 + 24 for _ in range(p.L))
    ts = p.ts.to(torch.float32).to(cuda)
    ca = P.CubicInterpolation(ts, tuple(c.to(cuda) for c in p.coeffs_adj))
    y0 = p.y0.to(cuda).requires_grad_(True)
    # one evaluation + VJP
    dy = vf(1.3, y0, ca)
    p64 = R.problem_to(p, torch.float64)
    layers = R.params_to(p64.layers, requires_grad=True)
    c64 = R.CubicInterpolation(p64.ts, p64.coeffs_adj)
    ref = R.perm_equiv_dir_vector_field(1.3, p64.y0, c64, layers, fus64)
    assert rel_err(dy, ref) < 2e-5
    # the solve, forward + exact adjoint, every leaf
    sol = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 3.0, 0.75, y0, ca)
    (sol.ys[-1] * p.gyT.to(cuda)).sum().backward()
    y64 = p64.y0.clone().requires_grad_(True)
    yT = R.tsit5_solve_fixed(lambda t, y: R.perm_equiv_dir_vector_field(t, y, c64, layers, fus64), y64, R.constant_step_table(0.0, 3.0, 0.75))
    (yT * p64.gyT).sum().backward()
    assert rel_err(sol.ys[-1], yT) < TOL_Y
    assert rel_err(y0.grad, y64.grad) < TOL_G
    for mine, lp, f64 in zip(vf.gnn_layers, layers, fus64):
        got = torch.stack([getattr(mine, f).grad for f in DIR_FIELDS])
        assert rel_err(got, f64.grad) < TOL_G
        assert rel_err(mine.conv_layer.linear.weight.grad, lp.weight.grad) < TOL_G
        assert rel_err(mine.conv_layer.norm.weight.grad, lp.norm_weight.grad) < TOL_G


def test_split_k_matches_the_single_pass_contraction(cuda, monkeypatch):
    """Launches far smaller than the GPU (one graph, n = 1000: 8 row blocks) run the contraction split-K (K slices on their
    own SMs, vector-atomic accumulation, epilogue-only second launch).  Same result as the single-pass kernel up to the
    fp32 summation order of the slices."""
    # seed 21: no hidden unit within fp32 rounding of its ReLU kink at t = 1.3 (seed 8 has one at |z| = 3.5e-7, where the
    # two summation orders legitimately pick different masks and d loss / d y differs by 7 % -- tools/probes/split_probe2.py)
    p = R.make_problem(n=1000, h=64, e=0, L=2, T=3, t1=2, dt0=0.5, seed=21)
    outs = []
    for no_split in ("1", None):
        if no_split:
            monkeypatch.setenv("PEG_TC_NO_SPLITK", no_split)
        else:
            monkeypatch.delenv("PEG_TC_NO_SPLITK", raising=False)
        vf, term, args = device_model(p, cuda, flags=TC)
        y = p.y0.to(cuda).requires_grad_(True)
        dy = term(1.3, y, args)
        (dy * p.gyT.to(cuda)).sum().backward()
        outs.append((dy.detach().clone(), y.grad.clone(), torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])))
    for a, b in zip(*outs):
        assert rel_err(a, b) < 5e-6


# ---------------------------------------------------------------------------------------------------
# the CUDA path against fixtures produced by executing the reference's OWN source files
# (oracle/pin_reference_source.py -> tests/golden/refsrc_*.npz; the reference tree is not needed at test time)
# ---------------------------------------------------------------------------------------------------
from oracle import pin_reference_source as PIN  # noqa: E402


def _refsrc_flag_cases():
    out = []
    for name, kw in PIN.REFSRC_CASES.items():
        out.append(pytest.param(name, 0, id=f"{name}-ffma"))
        if kw["n"] <= 256:
            out.append(pytest.param(name, FUSED, id=f"{name}-fused-small"))
        if kw["n"] >= 128:
            out.append(pytest.param(name, TC, id=f"{name}-tcgen05"))
            out.append(pytest.param(name, X3, id=f"{name}-tcgen05-tf32x3"))
    return out


@pytest.mark.parametrize("name,flags", _refsrc_flag_cases())
def test_cuda_path_against_reference_source_fixtures(cuda, name, flags):
    """PermEquivGraphVectorField / CDEWrapperVectorField / the sibling fields as the reference FILES compute them (fp64 execution
    of the unmodified sources, oracle/pin_reference_source.py), and the fixed-step solve over the reference's callable."""
    g = np.load(os.path.join(GOLD, f"refsrc_{name}.npz"))
    kw = PIN.REFSRC_CASES[name]
    p = R.make_problem(**kw)
    vf, term, args = device_model(p, cuda, flags=flags)
    y0 = p.y0.to(cuda)
    times = [float(t) for t in g["times"]]
    for k, t in enumerate(times):
        assert rel_err(term(t, y0, args), g["vf_perm_equiv"][k]) < 2e-5, (name, t)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), float(p.ts[0]), float(p.ts[-1]), kw["dt0"], y0, args,
                        stepsize_controller=P.ConstantStepSize(), saveat=P.SaveAt(t1=True))
    assert sol.stats["num_steps"] == int(g["steps"])
    assert rel_err(sol.ys[-1], g["yT"]) < TOL_Y, name
    # sibling fields on the same parameters (behind the wrapper on control shapes, like the fixtures)
    widths = R.layer_widths(p.h, p.L, p.e, p.e > 0)
    as_term = (lambda f: P.CDEWrapperVectorField(f, p.h)) if p.e > 0 else (lambda f: f)
    siblings = [("GraphVectorField", "vf_graph")] + ([("GNODEVectorField", "vf_gnode")] if p.e == 0 else [])
    for cls, key in siblings:
        sv = getattr(P, cls)(p.h, p.h, widths[-1], p.L, p.e, p.n, key=0)
        with torch.no_grad():
            for mine, lp in zip(sv.gnn_layers, p.layers):
                mine.linear.weight.copy_(lp.weight); mine.linear.bias.copy_(lp.bias)
                mine.norm.weight.copy_(lp.norm_weight); mine.norm.bias.copy_(lp.norm_bias)
        sv = sv.to(cuda)
        sv.flags = fused_flags() if flags == FUSED else per_operator(flags)
        for k, t in enumerate(times):
            assert rel_err(as_term(sv)(t, y0, args), g[key][k]) < 2e-5, (name, cls, t)
    dv = P.PermEquivDirGraphVectorField(p.h, p.h, widths[-1], p.L, p.e, p.n, key=0)
    tables = PIN.directed_fusion_tables(p.L, kw["seed"])
    with torch.no_grad():
        for mine, lp, tab in zip(dv.gnn_layers, p.layers, tables):
            mine.conv_layer.linear.weight.copy_(lp.weight); mine.conv_layer.linear.bias.copy_(lp.bias)
            mine.conv_layer.norm.weight.copy_(lp.norm_weight); mine.conv_layer.norm.bias.copy_(lp.norm_bias)
            for i, f in enumerate(PIN.DIRECTED_FIELDS):
                getattr(mine, f).copy_(tab[i].to(torch.float32))
    dv = dv.to(cuda)
    dv.flags = fused_flags() if flags == FUSED else per_operator(flags)
    for k, t in enumerate(times):
        assert rel_err(as_term(dv)(t, y0, args), g["vf_perm_equiv_dir"][k]) < 2e-5, (name, "directed", t)


def _load_linear_stack(mod, g, pre):
    """Copies fixture parameters <pre>_W{i} / <pre>_b{i} into a product MLP (``.layers``) or Linear."""
    layers = list(mod.layers) if hasattr(mod, "layers") else [mod]
    with torch.no_grad():
        for i, lin in enumerate(layers):
            lin.weight.copy_(torch.from_numpy(g[f"{pre}_W{i}"]).to(torch.float32))
            lin.bias.copy_(torch.from_numpy(g[f"{pre}_b{i}"]).to(torch.float32))
    assert f"{pre}_W{len(layers)}" not in g.files


@pytest.mark.parametrize("name", list(PIN.MODEL_CASES))
def test_models_against_reference_source_fixtures(cuda, name):
    """The host mirrors of the reference's solve wrappers (models.py) against fp64 executions of the reference's own
    pgt_ / tgb_ / graph_neural_cde.py (oracle/pin_reference_source.py: diffrax backed by the oracle's restatement)."""
    g = np.load(os.path.join(GOLD, f"refsrc_model_{name}.npz"))
    kind, kw, extra = PIN.MODEL_CASES[name]
    p = R.make_problem(**kw)
    vf, _, _ = device_model(p, cuda)
    inp = {k: torch.from_numpy(v).to(torch.float32).to(cuda) for k, v in PIN.model_inputs(name).items()}
    coeffs_adj = tuple(c.to(cuda) for c in p.coeffs_adj)
    if kind == "pgt":
        interp = extra.get("interpolation", "cubic")
        model = P.PGTGraphNeuralCDE(p.h, extra["data_dim"], extra["feature_dim"], vf, interp, seed=0).to(cuda)
        _load_linear_stack(model.encoder, g, "enc"); _load_linear_stack(model.decoder, g, "dec")
        ts = torch.arange(kw["T"], device=cuda)
        x_coeffs = tuple(c.to(cuda) for c in p.x_coeffs)
        if interp == "linear":   # the knot values are the "coefficients" of diffrax.LinearInterpolation
            coeffs_adj = torch.from_numpy(g["adj_knots"]).to(torch.float32).to(cuda)
            x_coeffs = torch.from_numpy(g["x_knots"]).to(torch.float32).to(cuda)
        ev = dict(evolving_out=True) if extra.get("evolving_out") else {}
        assert rel_err(model(ts, coeffs_adj, x_coeffs, inp["x0"], **ev), g["out_global"]) < TOL_Y
        assert rel_err(model(ts, coeffs_adj, x_coeffs, inp["x0"], global_readout=False, **ev), g["out_nodes"]) < TOL_Y
        if ev:   # reverse mode through the dense-output samples reaches the encoder and the vector field
            model(ts, coeffs_adj, x_coeffs, inp["x0"], **ev).sum().backward()
            assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in model.parameters())
    elif kind == "tgb":
        model = P.TGBGraphNeuralCDE(p.h, vf, use_mlps=extra["use_mlps"], seed=0, return_sequence=bool(extra.get("return_sequence"))).to(cuda)      # dt0 = 0.01 like the reference
        _load_linear_stack(model.encoder, g, "enc"); _load_linear_stack(model.decoder, g, "dec")
        with torch.no_grad():
            model.data_encoder.weight.copy_(torch.from_numpy(g["data_encoder_W"]).to(torch.float32))
            model.data_encoder.bias.copy_(torch.from_numpy(g["data_encoder_b"]).to(torch.float32))
        out = model(torch.arange(kw["T"], device=cuda), coeffs_adj, inp["x_data"], inp["x0"], None, evolving_out=bool(extra.get("evolving_out")))
        assert out.shape == g["out"].shape
        assert rel_err(out, g["out"]) < TOL_Y
    else:
        model = P.GraphNeuralCDE(p.h, vf, seed=0).to(cuda)
        _load_linear_stack(model.initial_linear, g, "enc"); _load_linear_stack(model.final_linear, g, "dec")
        out = model(p.ts.to(torch.float32).to(cuda), coeffs_adj, inp["x0"])
        assert out.shape == g["out"].shape
        # adaptive: PIDController(rtol=1e-3).  The fp64 fixture accepts 8 steps, an fp32 solve 9 (the CPU oracle in fp32 does the
        # same), and two accepted-step sequences differ by the controller tolerance: 1e-3 in fp32 on the CPU; bound 5e-3
        assert rel_err(out, g["out"]) < 5e-3


# ---------------------------------------------------------------------------------------------------
# The kernel instantiations bench.py measures (BASELINE.json configs[4]: n >= 1k, h = 128 / 256, batched graphs; configs[3]-style
# wide last layer) against the fp64 oracle: every graph of the batch has its own control path, parameter gradients are
# summed over the batch.  Both operand formats of the tcgen05 contraction are checked against the SAME tolerances:
# default = bf16x2 split, PEG_FLAG_TF32X3 = 3xTF32 split.
# ---------------------------------------------------------------------------------------------------
def _batched_problems(n, h, e, L, T, t1, dt0, seeds):
    ps = [R.make_problem(n=n, h=h, e=e, L=L, T=T, t1=t1, dt0=dt0, seed=s) for s in seeds]
    for p in ps[1:]:
        p.layers = ps[0].layers          # one parameter set for the whole batch (the reference's jax.vmap(model))
    return ps


def _batched_device_args(ps, cuda):
    ts = ps[0].ts.to(torch.float32).to(cuda)
    cadj = P.CubicInterpolation(ts, tuple(torch.stack([p.coeffs_adj[i] for p in ps]).to(cuda) for i in range(4)))
    if ps[0].e == 0:
        return cadj
    return [cadj, P.CubicInterpolation(ts, tuple(torch.stack([p.x_coeffs[i] for p in ps]).to(cuda) for i in range(4)))]


def _rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _entry_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return (a - b).abs() / b.abs().max().clamp_min(1e-300)


BF = _lib.PEG_FLAG_TENSOR_CORES | _lib.PEG_FLAG_BF16X2
OPERAND_FORMATS = [pytest.param(TC, id="fp16x2"), pytest.param(X3, id="tf32x3")]


@pytest.mark.parametrize("flags", OPERAND_FORMATS)
@pytest.mark.parametrize("n,h,e,B", [(2048, 128, 0, 3), (2048, 256, 0, 2), (1024, 128, 8, 2)], ids=["n2048_h128_B3", "n2048_h256_B2", "n1024_h128_e8_B2"])
def test_benchmarked_instantiations_one_evaluation_and_vjp(cuda, n, h, e, B, flags):
    """dy within 2e-5 (max-norm) for every graph of the batch.  The cotangent is compared entrywise with the ReLU kinks in mind:
    the field has 2 n h ~ 5e5 .. 1e6 hidden units, so a pre-activation within rounding distance (1e-6) of zero is an expected event;
    ONE flipped unit changes a few dozen rows of the exact cotangent by O(1e-2) and leaves every other entry untouched (measured
    on the GPU: 40 rows at n = 2048 / h = 128, 94 rows at h = 256, identically with 3xTF32 and with fp16x2 operands; the CPU oracle
    run in fp32 shows the same jumps on other seeds).  Hence: median entrywise error < 5e-6 and 90 % quantile < 5e-5 (the bulk is
    exact), at most 15 % of the rows above 5e-5 (a handful of flips), relative L2 < 2e-2; parameter gradients (sums over all
    nodes, a flip moves them by O(1/n)) within 3e-4 per leaf."""
    ps = _batched_problems(n, h, e, 3, 3, 2, 0.5, seeds=[31 + i for i in range(B)])
    vf, term, _ = device_model(ps[0], cuda, flags=flags)
    args = _batched_device_args(ps, cuda)
    t = 1.3
    y = torch.stack([p.y0 for p in ps]).to(cuda).requires_grad_(True)
    dy = term(t, y, args)
    (dy * torch.stack([p.gyT for p in ps]).to(cuda)).sum().backward()
    got = product_grads_as_oracle(vf)
    layers = R.params_to(R.params_to(ps[0].layers, torch.float64), requires_grad=True)
    flipped = 0
    for b, p in enumerate(ps):
        p64 = R.problem_to(p, torch.float64)
        q = R.Problem(p64.n, p64.h, p64.e, p64.L, p64.ts, p64.coeffs_adj, p64.x_coeffs, p64.y0, layers, p64.step_ts, p64.gyT)
        y64 = p64.y0.clone().requires_grad_(True)
        ref = _vf_oracle(q, t, y64)
        (ref * p64.gyT).sum().backward()       # parameter gradients accumulate over the batch in `layers`
        assert rel_err(dy[b].detach(), ref.detach()) < 2e-5, b
        err = _entry_err(y.grad[b], y64.grad)
        bad_rows = int((err.max(dim=1).values > 5e-5).sum())
        print(f"\n[n={n} h={h} e={e} graph {b}] dy {rel_err(dy[b].detach(), ref.detach()):.1e}  cotangent: max {float(err.max()):.1e} "
              f"q99.5 {float(err.flatten().quantile(0.995)):.1e} rows above 5e-5: {bad_rows}  rel L2 {_rel_l2(y.grad[b], y64.grad):.1e}")
        assert float(err.flatten().median()) < 5e-6 and float(err.flatten().quantile(0.9)) < 5e-5, b
        assert bad_rows <= 0.15 * n, (b, bad_rows)
        assert _rel_l2(y.grad[b], y64.grad) < 2e-2, b
        flipped += bad_rows
    worst = max(rel_err(g, r.grad) for g_l, lp in zip(got, layers) for g, r in zip(g_l, lp.tensors()))
    print(f"[n={n} h={h} e={e}] parameter gradients, worst leaf: {worst:.1e} (rows touched by ReLU flips: {flipped})")
    tol_p = 3e-4 if flipped == 0 else 5e-3      # a flipped unit also moves the gradients of the layers below it
    for l, (g_l, lp) in enumerate(zip(got, layers)):
        for name, g, r in zip(("fusion", "W", "b", "nw", "nb"), g_l, lp.tensors()):
            assert rel_err(g, r.grad) < tol_p, (l, name)


@pytest.mark.parametrize("flags", OPERAND_FORMATS)
def test_benchmarked_instantiation_whole_solve(cuda, flags):
    """A 5-step forward + adjoint solve at n=1024, h=128, B=2 (K = d_in = 128 in every tcgen05 kernel): Z_T within 1e-4 (max-norm).
    Gradients: over 5 x 6 evaluations of 2.6e5 hidden units ReLU masks DO flip at fp32 rounding distance -- the reference's own
    arithmetic in fp32 (the oracle run in fp32) differs from the fp64 oracle by 7e-4 relative L2 / 1e-2 max-norm on this problem
    -- so the yardstick is that fp32 run: relative L2 error <= max(1e-3, 3 x the fp32 oracle's), for y0 and for the parameters."""
    ps = _batched_problems(1024, 128, 0, 3, 3, 2, 0.1, seeds=[41, 42])
    t_end = 0.5
    vf, term, _ = device_model(ps[0], cuda, flags=flags)
    args = _batched_device_args(ps, cuda)
    y0 = torch.stack([p.y0 for p in ps]).to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, t_end, 0.1, y0, args)
    assert sol.stats["num_steps"] == 5
    (sol.ys[-1] * torch.stack([p.gyT for p in ps]).to(cuda)).sum().backward()
    got = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
    table = R.constant_step_table(0.0, t_end, 0.1)
    flat = {}
    for dt in (torch.float64, torch.float32):
        layers = R.params_to(R.params_to(ps[0].layers, dt), requires_grad=True)
        gy = []
        for b, p in enumerate(ps):
            pd = R.problem_to(p, dt)
            yd = pd.y0.clone().requires_grad_(True)
            yT = R.solve_cde(table, pd.ts, pd.coeffs_adj, None, yd, layers, pd.h, 0)
            (yT * pd.gyT).sum().backward()
            gy.append(yd.grad.double())
            if dt == torch.float64:
                assert rel_err(sol.ys[-1][b].detach(), yT.detach()) < TOL_Y, b
        flat[dt] = (gy, torch.cat([t.grad.reshape(-1) for lp in layers for t in lp.tensors()]).double())
    for b in range(len(ps)):
        ours, ref32 = _rel_l2(y0.grad[b], flat[torch.float64][0][b]), _rel_l2(flat[torch.float32][0][b], flat[torch.float64][0][b])
        print(f"\n[whole solve, graph {b}] y0-gradient rel L2: ours {ours:.1e}, fp32 oracle {ref32:.1e}")
        assert ours <= max(TOL_G, 3.0 * ref32), (b, ours, ref32)
    ours, ref32 = _rel_l2(got, flat[torch.float64][1]), _rel_l2(flat[torch.float32][1], flat[torch.float64][1])
    print(f"[whole solve] parameter-gradient rel L2: ours {ours:.1e}, fp32 oracle {ref32:.1e}")
    assert ours <= max(TOL_G, 3.0 * ref32), (ours, ref32)


@pytest.mark.parametrize("flags", OPERAND_FORMATS)
def test_batch_of_independent_graphs_on_tensor_cores(cuda, flags):
    """test_batch_of_independent_graphs at a tcgen05 shape (n = 300: three row blocks, ragged last one; B = 3)."""
    ps = _batched_problems(300, 64, 2, 2, 4, 3, 0.5, seeds=[0, 1, 3])
    vf, term, _ = device_model(ps[0], cuda, flags=flags)
    y0 = torch.stack([p.y0 for p in ps]).to(cuda).requires_grad_(True)
    sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, y0, _batched_device_args(ps, cuda))
    (sol.ys[-1] * torch.stack([p.gyT for p in ps]).to(cuda)).sum().backward()
    batched_grad = torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
    acc = torch.zeros_like(batched_grad)
    for i, p in enumerate(ps):
        vf.zero_grad()
        yi = p.y0.to(cuda).unsqueeze(0).requires_grad_(True)
        si = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, yi, _batched_device_args([p], cuda))
        assert rel_err(sol.ys[-1][i].detach(), si.ys[-1][0].detach()) < 1e-5
        (si.ys[-1][0] * p.gyT.to(cuda)).sum().backward()
        assert rel_err(y0.grad[i], yi.grad[0]) < 1e-4
        acc += torch.cat([t.reshape(-1) for layer in product_grads_as_oracle(vf) for t in layer])
    assert rel_err(batched_grad, acc) < 1e-4


def test_reused_adjacency_control_takes_the_node_signal_of_every_call(cuda):
    """Two training iterations with ONE adjacency control object and different node-signal controls (what the TGB model does at
    every step, tgb_graph_neural_cde.py:118-137): the adjacency planes are packed once, the node signal is never stale."""
    pa = R.make_problem(n=40, h=16, e=2, L=2, T=4, t1=3, dt0=0.5, seed=2)
    pb = R.make_problem(n=40, h=16, e=2, L=2, T=4, t1=3, dt0=0.5, seed=2, x_scale=0.9)
    vf, term, (cadj, cxa) = device_model(pa, cuda)
    cxb = P.CubicInterpolation(cadj.ts, tuple(c.to(cuda) for c in pb.x_coeffs))
    outs = []
    for cx in (cxa, cxb, cxa):
        y0 = pa.y0.to(cuda).requires_grad_(True)
        sol = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, y0, [cadj, cx])
        sol.ys[-1].sum().backward()
        outs.append((sol.ys[-1].detach().clone(), y0.grad.clone()))
    packed = cadj._packed
    assert packed is not None and packed.e == 0                      # the cache holds the adjacency part only
    fresh = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, pa.y0.to(cuda),
                          [P.CubicInterpolation(cadj.ts, (cadj.d, cadj.c, cadj.b, cadj.a)), cxb]).ys[-1]
    assert torch.equal(outs[1][0], fresh)                             # second iteration used ITS node signal
    assert not torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])
    assert cadj._packed is packed                                     # and the planes were not re-packed
    # a PackedControl handed in as the adjacency control behaves the same way
    again = P.diffeqsolve(P.ODETerm(term), P.Tsit5(), 0.0, 3.0, 0.5, pa.y0.to(cuda), [packed, cxb]).ys[-1]
    assert torch.equal(again, fresh)


# ---------------------------------------------------------------------------------------------------
# row-sharded mode: one graph spread over the ranks by rows, exchange over peer memory
# ---------------------------------------------------------------------------------------------------
def _rowshard_problem(n, h, L, T, B, seed, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.stack([torch.from_numpy(R.synthetic_graph_path(n, T, seed + b)).to(torch.float32) for b in range(B)])   # [B,T,n,n]
    ts = torch.arange(T, dtype=torch.float32)
    y0 = torch.randn((B, n, h), generator=g)
    gy = torch.randn((B, n, h), generator=g)
    return ts.to(device), A.to(device), y0.to(device), gy.to(device)


def _rowshard_reference(vf, ts, A, y0, gy, t1, dt0):
    """The whole graph on one GPU through the ordinary path."""
    pc = P.build_control(ts, A)
    vf.zero_grad()
    y = y0.clone().requires_grad_(True)
    yT = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, t1, dt0, y, pc).ys[-1]
    (yT * gy).sum().backward()
    return yT.detach(), y.grad.detach(), torch.cat([p.grad.reshape(-1) for p in vf.parameters()])


@pytest.mark.parametrize("flags", OPERAND_FORMATS)
@pytest.mark.parametrize("n,h,B", [(256, 64, 2), (384, 32, 1)])
def test_row_sharded_solve_on_one_rank_matches_the_ordinary_path(cuda, n, h, B, flags):
    """world = 1 exercises everything but the wires: rectangular strips, the transposed strip read like a direct one, V^T in the
    peer-visible buffers, the push / wait kernels (a rank is its own peer), epoch-parity double buffering."""
    from perm_equiv_graph_neural_cdes_b200 import rowshard as RS

    ts, A, y0, gy = _rowshard_problem(n, h, 3, 4, B, 3, cuda)
    vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=5, flags=per_operator(flags)).to(cuda)   # reference = the same tcgen05 kernels
    yT_ref, gy0_ref, gp_ref = _rowshard_reference(vf, ts, A, y0, gy, 1.0, 0.25)
    ctl = RS.RowShardedControl(ts, A, A.transpose(-1, -2).contiguous(), h, 3, flags=flags)
    dy = RS.vector_field_rowsharded(vf, ctl, 1.3, y0)
    assert rel_err(dy, vf(1.3, y0, P.build_control(ts, A))) < 1e-6
    vf.zero_grad()
    y = y0.clone().requires_grad_(True)
    yT = RS.diffeqsolve_rowsharded(vf, ctl, y, 0.0, 1.0, 0.25)
    (yT * gy).sum().backward()
    assert rel_err(yT.detach(), yT_ref) < 1e-6
    assert rel_err(y.grad, gy0_ref) < 1e-5
    assert rel_err(torch.cat([p.grad.reshape(-1) for p in vf.parameters()]), gp_ref) < 1e-5


def test_row_sharded_solve_replays_as_a_cuda_graph(cuda):
    """The exchange's epochs come from a device-side base that every call advances, so a captured forward + adjoint solve can be
    replayed: every replay must reproduce the eager result (an exchange that reused stale epochs would read V^T before it is
    written, or -- with an odd number of exchanges per call -- overwrite the half a peer still reads)."""
    from perm_equiv_graph_neural_cdes_b200 import rowshard as RS

    n, h, B = 256, 64, 2
    ts, A, y0, gy = _rowshard_problem(n, h, 3, 4, B, 7, cuda)
    vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=5).to(cuda)
    ctl = RS.RowShardedControl(ts, A, A.transpose(-1, -2).contiguous(), h, 3, flags=vf.flags)

    def step(dt0):
        vf.zero_grad(set_to_none=True)
        y = y0.detach().requires_grad_(True)
        yT = RS.diffeqsolve_rowsharded(vf, ctl, y, 0.0, 1.0, dt0, reduce_grads=False)
        (yT * gy).sum().backward()
        return yT.detach(), y.grad, torch.cat([p.grad.reshape(-1) for p in vf.parameters()])

    for dt0 in (1.0, 0.25):      # 1 step = 18 + 18 exchanges; dt0 = 1.0 also covers calls right after one another
        ref = [t.clone() for t in step(dt0)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(dt0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step(dt0)
        for _ in range(3):
            graph.replay()
            torch.cuda.synchronize()
            assert rel_err(out[0], ref[0]) < 1e-6
            assert rel_err(out[1], ref[1]) < 1e-5
            assert rel_err(out[2], ref[2]) < 1e-5
        assert int(ctl._bufs[4].view(torch.int32)[ctl.world].item()) == 0      # the exchange's error word
        eager = step(dt0)       # and an eager call after the replays still lines up with the epoch base
        assert rel_err(eager[0], ref[0]) < 1e-6


def _rowshard_rank(rank, world, port, n, h, B, flags, out):
    import torch.distributed as dist

    from perm_equiv_graph_neural_cdes_b200 import rowshard as RS

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ts, A, y0, gy = _rowshard_problem(n, h, 3, 4, B, 3, dev)
        vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=5, flags=flags).to(dev)
        r0, r1 = RS.row_range(n, rank, world)
        ctl = RS.RowShardedControl(ts, A[:, :, r0:r1].contiguous(), A[:, :, :, r0:r1].transpose(-1, -2).contiguous(), h, 3, flags=flags)
        y = y0[:, r0:r1].clone().requires_grad_(True)
        yT = RS.diffeqsolve_rowsharded(vf, ctl, y, 0.0, 1.0, 0.25)
        (yT * gy[:, r0:r1]).sum().backward()
        torch.cuda.synchronize()
        res = [yT.detach().cpu(), y.grad.cpu(), torch.cat([p.grad.reshape(-1) for p in vf.parameters()]).cpu()]
        if rank == 0:
            ref = [t.cpu() for t in _rowshard_reference(vf, ts, A, y0, gy, 1.0, 0.25)]
            res += ref
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("flags", OPERAND_FORMATS)
def test_row_sharded_solve_over_two_gpus(cuda, flags):
    """Two ranks, one graph: rows of Z_T and of the y0-cotangent from each rank, parameter gradients all-reduced, against the whole
    graph solved on one GPU (1e-6 / 1e-5: the same kernels, a different summation partition of the column sums)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import socket

    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n, h, B, world = 512, 64, 2, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rowshard_rank, args=(world, port, n, h, B, flags, out), nprocs=world, join=True)
    yT_ref, gy0_ref, gp_ref = out[0][3], out[0][4], out[0][5]
    yT = torch.cat([out[r][0] for r in range(world)], dim=1)
    gy0 = torch.cat([out[r][1] for r in range(world)], dim=1)
    assert rel_err(yT, yT_ref) < 1e-6
    assert rel_err(gy0, gy0_ref) < 1e-5
    for r in range(world):
        assert rel_err(out[r][2], gp_ref) < 1e-5, r


# ---------------------------------------------------------------------------------------------------
# batch-sharded data parallelism on real GPUs: NCCL all-reduce of the sharded gradients == the full-batch gradients
# ---------------------------------------------------------------------------------------------------
def _dp_rank(rank, world, port, out):
    import torch.distributed as dist

    from perm_equiv_graph_neural_cdes_b200 import dist as pdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n, h, B = 256, 64, 4
        ts, A, y0, gy = _rowshard_problem(n, h, 3, 4, B, 11, dev)
        vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=5).to(dev)

        def grads(sel):
            vf.zero_grad()
            y = y0[sel].clone().requires_grad_(True)
            yT = P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 1.0, 0.25, y, P.build_control(ts, A[sel])).ys[-1]
            (yT * gy[sel]).sum().backward()
            return yT.detach()

        b0, b1 = pdist.shard_range(B, rank, world)
        yT = grads(slice(b0, b1))
        flat = pdist.allreduce_gradients(list(vf.parameters()))        # ONE NCCL all-reduce of the flat buffer
        res = [flat.cpu(), yT.cpu()]
        if rank == 0:
            yT_full = grads(slice(0, B))
            res += [torch.cat([p.grad.reshape(-1) for p in vf.parameters()]).cpu(), yT_full.cpu()]
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_nccl_reduced_sharded_gradients_equal_the_full_batch_gradients(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import socket

    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_rank, args=(2, port, out), nprocs=2, join=True)
    full_g, full_y = out[0][2], out[0][3]
    assert torch.equal(out[0][0], out[1][0])                                   # every rank holds the same reduced buffer
    assert rel_err(out[0][0], full_g) < 1e-5
    assert rel_err(torch.cat([out[0][1], out[1][1]]), full_y) < 1e-6          # the shards are the rows of the full batch
