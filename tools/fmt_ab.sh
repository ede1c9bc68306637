#!/bin/bash
# A/B of the three operand formats on one workload (short bench runs).  Usage: bash tools/fmt_ab.sh <tag> [workload] [formats...]
T=${1:-fmt}; W=${2:-sweep_n2048_h128}; shift; shift; F=${@:-fp16x2 bf16x2 tf32x3}
O=gpurun_out/$T; mkdir -p $O
for f in $F; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --workload $W --operands $f > $O/bench_$f.json 2> $O/bench_$f.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$W $f value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as ex: print("$f failed:", ex)
PY
done
