#!/usr/bin/env python
"""bench.py -- graph-trajectory solver steps/sec (fwd+bwd) of the fused sm_100a hot path.

One bench "step" = one forward + exact-adjoint backward solve of a batch of B graph trajectories
(S Tsit5 steps each) per GPU.  value = N_gpus * B * S / seconds_per_bench_step  (whole-job aggregate).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the batch of trajectories is sharded over the
ranks (weak scaling: B graphs per GPU) and the flat parameter-gradient buffer is all-reduced every step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> shape.  The C5 sweep points are BASELINE.json configs[4]; sir / england / twitter are configs[1..3].
WORKLOADS = {
    "sir": dict(n=100, h=32, e=3, L=3, T=120, t1=1.0, dt0=0.1, B=50, float_ts=True),
    "england": dict(n=129, h=64, e=8, L=3, T=4, t1=3.0, dt0=0.1, B=1),
    "twitter": dict(n=1000, h=64, e=16, L=3, T=9, t1=8.0, dt0=0.1, B=1),
    # graphs per GPU are chosen so that (row blocks of 128) x B is just under a multiple of the 148 SMs
    "sweep_n1024_h64": dict(n=1024, h=64, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=18),
    "sweep_n2048_h64": dict(n=2048, h=64, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=9),
    "sweep_n2048_h128": dict(n=2048, h=128, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=9),
    "sweep_n4096_h128": dict(n=4096, h=128, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=9),
    "sweep_n4096_h128_b4": dict(n=4096, h=128, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=4),
    "sweep_n4096_h256": dict(n=4096, h=256, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=4),
    "sweep_n8192_h128": dict(n=8192, h=128, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=2),
    "sweep_n16384_h128": dict(n=16384, h=128, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=1),
    "sweep_n16384_h256": dict(n=16384, h=256, e=0, L=3, T=9, t1=8.0, dt0=0.1, B=1),
}
DEFAULT_WORKLOAD = "sweep_n2048_h128"
METRIC = "graph-trajectory solver steps/sec (fwd+bwd)"
UNIT = "solver steps/s"


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (mean of one forward and one
    adjoint launch), from the committed `ncu --set full` captures of this workload (profiles/r01_traffic.json);
    None if no capture exists."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(path):
        ent = json.load(open(path)).get(workload)
        return ent["mean_bytes_per_launch"] if isinstance(ent, dict) else ent
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops_sustained"]), src="measured")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def synth_adjacency(n, T, seed, device):
    """Device-side synthetic dynamic weighted digraph [T,n,n]: Bernoulli(min(1,16/n)) mask with
    10% of entries redrawn per knot, LogNormal(0,1) weights, unit self-loops, row-normalised."""
    g = torch.Generator(device=device).manual_seed(seed)
    p = min(1.0, 16.0 / n)
    mask = torch.rand((n, n), generator=g, device=device) < p
    w = torch.exp(torch.randn((n, n), generator=g, device=device))
    out = torch.empty((T, n, n), device=device)
    eye = torch.eye(n, device=device, dtype=torch.bool)
    for k in range(T):
        if k > 0:
            redraw = torch.rand((n, n), generator=g, device=device) < 0.10
            mask = torch.where(redraw, torch.rand((n, n), generator=g, device=device) < p, mask)
            w = torch.where(redraw, torch.exp(torch.randn((n, n), generator=g, device=device)), w)
        a = torch.where(mask, w, torch.zeros_like(w))
        a = torch.where(eye, torch.ones_like(a), a)
        out[k] = a / a.sum(dim=1, keepdim=True)
    return out


def make_inputs(wl, seed, device, host_copy):
    """Inputs of one rank.  Reference layout: ts [T], coeffs_adj (d,c,b,a) each [B,T-1,n,n,2], x_coeffs, y0, cotangent -- plus the
    graph snapshots A_k [B,T,n,n] they were built from.  The reference-layout arrays are 32 n^2 (T-1) bytes per graph; past
    COEFF_BUDGET bytes (n = 16384: 69 GB) only the snapshots are kept and the control path is built on the device
    (pegncde_build_adj, bit-identical planes)."""
    import perm_equiv_graph_neural_cdes_b200 as P

    n, h, e, T, B = wl["n"], wl["h"], wl["e"], wl["T"], wl["B"]
    ts = torch.linspace(0.0, wl["t1"], T, device=device) if wl.get("float_ts") else torch.arange(T, device=device, dtype=torch.float32) * (wl["t1"] / (T - 1))
    coeff_bytes = 4 * B * (T - 1) * n * n * 2 * 4
    with_coeffs = coeff_bytes <= COEFF_BUDGET
    cadj = [torch.empty((B, T - 1, n, n, 2), device=device) for _ in range(4)] if with_coeffs else None
    snaps_dev = None if with_coeffs else torch.empty((B, T, n, n), device=device)
    snaps = torch.empty((B, T, n, n), dtype=torch.float32).pin_memory() if host_copy else None
    for b in range(B):
        A = synth_adjacency(n, T, seed * 1000 + b, device)
        if snaps is not None:
            snaps[b].copy_(A)
        if with_coeffs:
            X = torch.stack([ts[:, None, None].expand(T, n, n), A], dim=-1)
            for dst, src in zip(cadj, P.backward_hermite_coefficients(ts, X)):
                dst[b].copy_(src)
            del X
        else:
            snaps_dev[b].copy_(A)
        del A
    g = torch.Generator(device=device).manual_seed(seed + 77)
    xco, x_t = None, None
    if e > 0:
        x_t = 0.3 * torch.randn((B, T, n, e), generator=g, device=device)
        X = torch.stack([ts[None, :, None, None].expand(B, T, n, e), x_t], dim=-1)
        xco = [torch.stack([P.backward_hermite_coefficients(ts, X[b])[i] for b in range(B)]) for i in range(4)]
    y0 = torch.randn((B, n, h), generator=g, device=device)
    gy = torch.randn((B, n, h), generator=g, device=device)
    host = None
    if host_copy:
        pin = lambda t: t.cpu().pin_memory()
        host = dict(ts=pin(ts), cadj=[pin(c) for c in cadj] if with_coeffs else None, xco=None if xco is None else [pin(c) for c in xco], y0=pin(y0),
                    snaps=snaps, x_t=None if e == 0 else pin(x_t))
    return ts, cadj, xco, y0, gy, host, snaps_dev, x_t


COEFF_BUDGET = 40e9


def run_ours(args):
    import perm_equiv_graph_neural_cdes_b200 as P
    from perm_equiv_graph_neural_cdes_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    if args.t1 > 0:
        wl["t1_solve"] = args.t1
    n, h, e, L, T, B = wl["n"], wl["h"], wl["e"], wl["L"], wl["T"], wl["B"]
    flags = 0 if args.no_tensor_cores else _lib.PEG_FLAG_TENSOR_CORES
    if args.tf32_fast:
        flags |= _lib.PEG_FLAG_TF32_FAST
    if args.operands == "tf32x3":
        flags |= _lib.PEG_FLAG_TF32X3
    elif args.operands == "bf16x2":
        flags |= _lib.PEG_FLAG_BF16X2

    widths_out = 2 * h * e if e > 0 else h
    vf = P.PermEquivGraphVectorField(h, h, widths_out, L, e, n, key=1234).to(dev)
    vf.flags = flags
    term = P.ODETerm(P.CDEWrapperVectorField(vf, h) if e > 0 else vf)
    ts, cadj, xco, y0, gy, host, snaps_dev, x_t = make_inputs(wl, 1234 + rank, dev, host_copy=True)
    pc = P.pack_control(ts, cadj, xco) if cadj is not None else P.build_control(ts, snaps_dev, x_t)
    torch.cuda.synchronize()
    del cadj, xco, snaps_dev
    torch.cuda.empty_cache()
    t1_solve = wl.get("t1_solve", wl["t1"])
    step_ts = P.constant_step_table(0.0, t1_solve, wl["dt0"])
    S = len(step_ts) - 1
    l = _lib.lib()

    def solve_step(pc_, y0_, reduce=True):
        vf.zero_grad(set_to_none=True)
        y = y0_.detach().requires_grad_(True)
        sol = P.diffeqsolve(term, P.Tsit5(), 0.0, t1_solve, wl["dt0"], y, [pc_, None] if e > 0 else pc_,
                            stepsize_controller=P.ConstantStepSize(), saveat=P.SaveAt(t1=True))
        loss = (sol.ys[-1] * gy).sum()
        loss.backward()
        flat_g = torch.cat([p.grad.reshape(-1) for p in vf.parameters()])
        if world > 1 and reduce:
            torch.distributed.all_reduce(flat_g)
        return loss, flat_g

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value` ----------------
    # The C-ABI only enqueues kernels, so one whole forward + adjoint solve is captured once into a CUDA graph and
    # replayed: ~2.5k dependent launches per step are otherwise host-launch-bound for the smaller shapes.
    for _ in range(max(args.warmup, 1)):
        solve_step(pc, y0)
    barrier()
    graph = None
    l.pegncde_profile_enable(args.profile_stride)
    launches0 = l.pegncde_launch_count()
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            solve_step(pc, y0, reduce=False)   # warm the capture stream's workspace
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        l.pegncde_profile_enable(args.profile_stride)
        launches0 = l.pegncde_launch_count()
        with torch.cuda.graph(graph):
            g_loss, g_flat = solve_step(pc, y0, reduce=False)
        launches_per_step = l.pegncde_launch_count() - launches0
        for _ in range(2):
            graph.replay()
        barrier()

    def timed_step():
        if graph is not None:
            graph.replay()
            if world > 1:
                torch.distributed.all_reduce(g_flat)
        else:
            solve_step(pc, y0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        timed_step()
    e1.record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps
    barrier()
    clocks = sampler.summary()
    ms = e0.elapsed_time(e1)
    if graph is None:
        launches_per_step = (l.pegncde_launch_count() - launches0) // max(args.steps, 1)
    launches = launches_per_step * args.steps
    prof = {}
    for d, nm in ((0, "fwd"), (1, "bwd")):
        a, b_, c_, by, fl = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        l.pegncde_profile_read(d, a, b_, c_, by, fl)
        prof[nm] = dict(launches=a.value, timed=b_.value, ms=c_.value, bytes=by.value, flops=fl.value)
    l.pegncde_profile_enable(0)
    prof_steps = 1 if graph is not None else args.steps   # a replayed graph re-records the same event pairs
    t = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * B * S / (ms_per_step * 1e-3)

    # ---------------- end-to-end leg: host buffers through the public API ----------------
    def e2e_step():
        # host buffers go straight into the public API: pack_control keeps the (pinned) coefficient arrays on the host and the
        # solve streams them piece by piece (copy + pack of piece i+1 overlap the steps inside piece i)
        y0_d = host["y0"].to(dev, non_blocking=True)
        pc_ = P.pack_control(host["ts"], host["cadj"], host["xco"], device=dev)
        loss, flat_g = solve_step(pc_, y0_d)
        return float(loss.item()), flat_g.cpu()

    k_e2e = max(1, min(args.steps, args.e2e_steps))
    e2e_val, h2d = None, None
    if host["cadj"] is not None:
        h2d = sum(c.numel() * 4 for c in host["cadj"]) + host["y0"].numel() * 4 + host["ts"].numel() * 4
        if host["xco"] is not None:
            h2d += sum(c.numel() * 4 for c in host["xco"])
        e2e_step()
        barrier()
        e0.record()
        for _ in range(k_e2e):
            e2e_step()
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
        e2e_val = world * B * S / (float(t2.item()) / k_e2e * 1e-3)
    d2h = vf.flat_params().numel() * 4 + 4

    # same end-to-end step, but entering one stage earlier: the host ships the graph SNAPSHOTS A_k [B,T,n,n] and
    # the control path (Hermite coefficients + tiling) is built on the device by pegncde_build_adj
    def e2e_snap_step():
        ts_d = host["ts"].to(dev, non_blocking=True)
        A_d = host["snaps"].to(dev, non_blocking=True)
        x_d = None if host["x_t"] is None else host["x_t"].to(dev, non_blocking=True)
        y0_d = host["y0"].to(dev, non_blocking=True)
        pc_ = P.build_control(ts_d, A_d, x_d)
        loss, flat_g = solve_step(pc_, y0_d)
        return float(loss.item()), flat_g.cpu()

    e2e_snap_step()
    barrier()
    e0.record()
    for _ in range(k_e2e):
        e2e_snap_step()
    e1.record()
    barrier()
    t3 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t3, op=torch.distributed.ReduceOp.MAX)
    e2e_snap_val = world * B * S / (float(t3.item()) / k_e2e * 1e-3)
    h2d_snap = host["snaps"].numel() * 4 + host["y0"].numel() * 4 + host["ts"].numel() * 4 + (0 if host["x_t"] is None else host["x_t"].numel() * 4)

    # ---------------- roofline of the dominant kernel (the n x n x d contraction) ----------------
    pk = peaks()
    tot_ms = prof["fwd"]["ms"] + prof["bwd"]["ms"]
    tot_timed = prof["fwd"]["timed"] + prof["bwd"]["timed"]
    roof = None
    if tot_timed:
        by = (prof["fwd"]["bytes"] * prof["fwd"]["timed"] + prof["bwd"]["bytes"] * prof["bwd"]["timed"])
        fl = (prof["fwd"]["flops"] * prof["fwd"]["timed"] + prof["bwd"]["flops"] * prof["bwd"]["timed"])
        gbs = by / (tot_ms * 1e-3) / 1e9
        tfs = fl / (tot_ms * 1e-3) / 1e12
        # The planes stream from HBM once per launch; the tensor pipe executes `passes` tf32 MMAs per algorithmic product
        # (3xTF32 split = fp32 parity; 1 with --tf32-fast).  tf32 dense peak ~ 1/2 of the measured bf16 peak.
        tf32_peak = pk["bf16"] / 2.0
        passes = 1 if (flags & 2) or not (flags & 1) else 3
        # the opt-in LIGHT adjoint (PEG_TC_ADJ_LIGHT=1) runs two of its four products single-pass: 8 MMA passes instead of 12
        bwd_passes = 2.0 if (passes == 3 and os.environ.get("PEG_TC_ADJ_LIGHT")) else passes
        exec_tfs = (passes * prof["fwd"]["flops"] * prof["fwd"]["timed"] + bwd_passes * prof["bwd"]["flops"] * prof["bwd"]["timed"]) / (tot_ms * 1e-3) / 1e12
        # the bound is decided by ALGORITHMIC intensity (flops / bytes against the tf32 ridge); the executed tensor work
        # (passes x) is reported beside it: with 3xTF32 the adjoint (four products) is co-limited by the tensor pipe
        t_hbm, t_tc = by / (pk["hbm"] * 1e9), fl / (tf32_peak * 1e12)
        bound = "hbm" if t_hbm >= t_tc else "tensor"
        share = tot_ms * (prof["fwd"]["launches"] + prof["bwd"]["launches"]) / max(tot_timed, 1) / (ms_per_step * prof_steps)
        roof = {"bound": bound, "kernel": "k_tc_contract<fwd> + k_tc_contract<adjoint> (the n x n x d contraction)" if flags & 1 else "k_dual_contract (FFMA)",
                "achieved": gbs if bound == "hbm" else tfs, "peak": pk["hbm"] if bound == "hbm" else tf32_peak,
                "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": (gbs / pk["hbm"]) if bound == "hbm" else (tfs / tf32_peak),
                "peak_source": pk["src"] + (" hbm_gbs" if bound == "hbm" else " bf16_tflops_sustained/2 (tf32)"),
                "traffic": measured_traffic(args.workload), "achieved_gbs": gbs, "achieved_tflops": tfs,
                "tensor_passes": passes, "adjoint_tensor_passes": bwd_passes, "executed_tflops": exec_tfs, "executed_tensor_frac": exec_tfs / tf32_peak,
                "fwd_executed_tensor_frac": passes * prof["fwd"]["flops"] * prof["fwd"]["timed"] / max(prof["fwd"]["ms"], 1e-9) / 1e9 / tf32_peak,
                "bwd_executed_tensor_frac": bwd_passes * prof["bwd"]["flops"] * prof["bwd"]["timed"] / max(prof["bwd"]["ms"], 1e-9) / 1e9 / tf32_peak,
                "avg_launch_us": tot_ms / tot_timed * 1e3,
                "launches_timed": tot_timed, "share_of_step": share,
                "fwd_avg_us": prof["fwd"]["ms"] / max(prof["fwd"]["timed"], 1) * 1e3, "bwd_avg_us": prof["bwd"]["ms"] / max(prof["bwd"]["timed"], 1) * 1e3,
                "fwd_gbs": prof["fwd"]["bytes"] * prof["fwd"]["timed"] / max(prof["fwd"]["ms"], 1e-9) / 1e6,
                "bwd_gbs": prof["bwd"]["bytes"] * prof["bwd"]["timed"] / max(prof["bwd"]["ms"], 1e-9) / 1e6,
                "algorithmic_bytes_per_launch": {"fwd": prof["fwd"]["bytes"], "adjoint": prof["bwd"]["bytes"]},
                "algorithmic_flops_per_launch": {"fwd": prof["fwd"]["flops"], "adjoint": prof["bwd"]["flops"]},
                "l2_note": "planes %s L2 (126 MB)" % ("fit in" if 16.0 * n * n * B <= 126e6 else "exceed")}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(wl, args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" + (" (contraction: 3xTF32 tcgen05, fp32 accumulate)" if flags & 1 and not flags & 2 else
                               " (contraction: 1xTF32 tcgen05)" if flags & 2 else " (CUDA-core FFMA)"),
            "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "hidden": h, "data_embed_dim": e, "layers": L, "knots": T,
                       "solver": "Tsit5 fixed dt0=%g" % wl["dt0"], "solver_steps": S, "graphs_per_gpu": B,
                       "parallelism": "batch of trajectories sharded over %d GPU(s); NCCL all-reduce of %d param grads" % (world, vf.flat_params().numel()),
                       "l2": "inputs per GPU %.0f MB of coefficient planes (> L2 126 MB: %s)" % (pc.adj_coef.numel() * 4 / 1e6, pc.adj_coef.numel() * 4 > 126e6)},
            "e2e": {"value": e2e_val if e2e_val is not None else e2e_snap_val, "unit": UNIT,
                    "h2d_bytes_per_step": h2d if e2e_val is not None else h2d_snap, "d2h_bytes_per_step": d2h, "steps": k_e2e,
                    "input": "reference-layout coefficient arrays (d,c,b,a) [B,T-1,n,n,2] in pinned host memory handed to pack_control / diffeqsolve; the solve streams them: copy + pack of cubic piece i+1 overlap the steps inside piece i" if e2e_val is not None else
                             "graph snapshots A_k [B,T,n,n] from pinned host memory (the coefficient arrays of this workload exceed the 40 GB input budget)"},
            "e2e_from_snapshots": {"value": e2e_snap_val, "unit": UNIT, "h2d_bytes_per_step": h2d_snap, "d2h_bytes_per_step": d2h, "steps": k_e2e,
                                   "input": "graph snapshots A_k [B,T,n,n]; control path built on the device (pegncde_build_adj)"},
            "gpu_launches": int(launches), "cuda_graph": graph is not None, "host_ms_per_step": host_ms, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_problem(wl, seed, sample_steps):
    """The same workload shape for the CPU oracle, restricted to the first `sample_steps` solver steps."""
    from oracle import reference_path as R

    p = R.make_problem(n=wl["n"], h=wl["h"], e=wl["e"], L=wl["L"], T=wl["T"], t1=wl["t1"], dt0=wl["dt0"], seed=seed,
                       float_ts=bool(wl.get("float_ts")), randomize_norm=False)
    p.step_ts = p.step_ts[: sample_steps + 1]
    return p


def time_oracle(wl, sample_steps, repeats, threads):
    from oracle import reference_path as R

    torch.set_num_threads(threads)
    p = cpu_problem(wl, 1234, sample_steps)
    R.run_forward_backward(cpu_problem(wl, 1234, 1))  # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        R.run_forward_backward(p)
    dt = (time.perf_counter() - t0) / repeats
    return sample_steps / dt, dt


def cpu_baseline(wl, args):
    threads = os.cpu_count() or 1
    sample_steps = args.cpu_sample_steps
    v, dt = time_oracle(wl, sample_steps, 1, threads)
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "1 graph x %d solver steps fwd+bwd of the same workload shape, torch-CPU restatement (oracle), %.1f s" % (sample_steps, dt)}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (the oracle port -- JAX/diffrax are not installable here)
    with all host threads on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = dict(WORKLOADS[args.workload])
    threads = os.cpu_count() or 1
    from oracle import reference_path as R

    torch.set_num_threads(threads)
    sample_steps = args.cpu_sample_steps
    p = cpu_problem(wl, 1234, sample_steps)
    for _ in range(min(args.warmup, 1)):
        R.run_forward_backward(cpu_problem(wl, 1234, 1))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.run_forward_backward(p)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample_steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n": wl["n"], "hidden": wl["h"], "data_embed_dim": wl["e"], "layers": wl["L"],
                       "knots": wl["T"], "solver": "Tsit5 fixed dt0=%g" % wl["dt0"],
                       "solver_steps": len(R.constant_step_table(0.0, wl["t1"], wl["dt0"])) - 1, "graphs_per_gpu": wl["B"],
                       "parallelism": "host threads of one box (the reference has no multi-device path)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step = 1 graph x %d solver steps fwd+bwd (torch-CPU restatement of the reference path)" % sample_steps},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="graphs per GPU (default: the workload's)")
    ap.add_argument("--no-tensor-cores", action="store_true")
    ap.add_argument("--tf32-fast", action="store_true")
    ap.add_argument("--operands", default="fp16x2", choices=["fp16x2", "bf16x2", "tf32x3"],
                    help="operand format of the tcgen05 contraction: fp16x2 with block exponents (default), bf16x2 (looser tolerance), 3xTF32")
    ap.add_argument("--profile-stride", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-steps", type=int, default=24, help="solver steps of one graph the CPU baseline runs (~15 s of CPU work at the default workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--t1", type=float, default=0.0, help="profiling only: shorten the solve to [0, t1] (fewer solver steps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
