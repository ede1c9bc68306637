#!/bin/bash
# timing A/B of experimental library builds for several operand formats.  Usage: bash tools/exp_ab2.sh <tag> "<libs>" "<formats>"
O=gpurun_out/$1; mkdir -p $O
for v in $2; do for f in $3; do
  L=""; [ "$v" != base ] && L="$PWD/tools/exp/libexp$v.so"
  PEGNCDE_LIB=$L timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --operands $f > $O/bench_${v}_$f.json 2> $O/bench_${v}_$f.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${v}_$f.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$v $f value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3))
except Exception as ex: print("$v $f failed:", ex)
PY
done; done
