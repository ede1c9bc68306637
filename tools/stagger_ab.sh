#!/bin/bash
# start-up stagger of converter group 1 (PEG_TC_STAGGER_NS) on the default workload, both 16-bit formats
O=gpurun_out/${1:-stagger}; mkdir -p $O
for f in fp16x2 bf16x2; do for ns in 0 600 1200 2000; do
  PEG_TC_STAGGER_NS=$ns timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-sweep --no-tensor-peaks --operands $f > $O/b_${f}_$ns.json 2> $O/b_${f}_$ns.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b_${f}_$ns.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$f stagger $ns value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3))
except Exception as ex: print("$f $ns failed:", ex)
PY
done; done
