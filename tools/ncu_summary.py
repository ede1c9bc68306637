#!/usr/bin/env python
"""Summarises the two `ncu --set full` captures of tools/r2_profile.sh (forward / adjoint contraction launch) as text + the traffic JSON.
Usage: python tools/ncu_summary.py gpurun_out/<tag> profiles/r02   (writes <prefix>_ncu_summary.txt, <prefix>_traffic.json)"""
import csv, io, json, re, subprocess, sys

src, prefix = sys.argv[1], sys.argv[2]
WANT = re.compile(r"^(Kernel Name|Block Size|Grid Size|dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"gpu__time_duration\.sum|l1tex__data_pipe_lsu_wavefronts\.sum\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__throughput\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate\.pct|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.avg\.per_second|sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active)$")
TITLES = {0: "[forward contraction k_tc_contract<0,2>]", 1: "[adjoint contraction k_tc_contract<1,2> (4 accumulators, fused fusion-scalar gradients)]"}
ALGO = {0: 622854144, 1: 632291328}      # 9 * (16 n^2 + 8 n d), 9 * (16 n^2 + 12 n d) at n = 2048, d = 128
out, traffic = [], {}
def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
for k in (0, 1):
    raw = subprocess.run(["ncu", "-i", f"{src}/full_k_tc_contractILi{k}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out.append(TITLES[k])
    sel = sorted((h, v, u) for h, u, v in zip(hdr, units, vals) if WANT.match(h))
    out += [f"    {h} = {v} {u}" for h, v, u in sel]
    d = {h: (v, u) for h, v, u in sel}
    traffic[k] = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v.replace(",", ""))) for h, v in zip(hdr, vals)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v]
    tot = sum(v for _, v in st) or 1.0
    out.append("    warp-state samples (smsp__pcsamp_warps_issue_stalled_*, top entries):")
    out += [f"        {h:28s} {v:10.0f}  {v / tot:.3f}" for h, v in sorted(st, key=lambda x: -x[1])[:8]]
    out.append(f"    dram bytes read + written = {traffic[k] / 1e6:.1f} MB = {traffic[k] / ALGO[k]:.3f} x the algorithmic bytes ({ALGO[k] / 1e6:.1f} MB)")
    out.append("")
open(prefix + "_ncu_body.txt", "w").write("\n".join(out))
json.dump({"sweep_n2048_h128": {"operands": "fp16x2", "fwd_bytes_per_launch": traffic[0], "adjoint_bytes_per_launch": traffic[1],
                                "mean_bytes_per_launch": (traffic[0] + traffic[1]) / 2, "algorithmic_bytes_per_launch": {"fwd": ALGO[0], "adjoint": ALGO[1]},
                                "source": prefix + "_ncu_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch each)"}},
          open(prefix + "_traffic.json", "w"), indent=1)
print("\n".join(out))
