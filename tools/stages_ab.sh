#!/bin/bash
# ring depths of the contraction kernel (PEG_TC_STAGES_A / _B) on the default workload
O=gpurun_out/${1:-stages}; mkdir -p $O
for cfg in "4 4" "4 2" "4 3" "2 2" "2 4"; do set -- $cfg
  PEG_TC_STAGES_A=$1 PEG_TC_STAGES_B=$2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-sweep --no-tensor-peaks > $O/b_$1_$2.json 2> $O/b_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b_$1_$2.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("SA $1 SB $2 value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as ex: print("$1 $2 failed:", ex)
PY
done
