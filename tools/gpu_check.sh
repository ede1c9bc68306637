#!/bin/bash
# GPU regression round: parity suite, smoke, default bench (short), in one gpurun call.  Usage: bash tools/gpu_check.sh [tag]
T=${1:-check}; O=gpurun_out/$T; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1) > $O/pytest_gpu.log 2>&1; grep -a "stiff-case\|passed\|failed\|Error\|error" $O/pytest_gpu.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 400 $B > $O/bench.json 2> $O/bench.err || tail -5 $O/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("bench value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex: print("bench failed:", ex)
PY
