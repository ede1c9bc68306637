#!/bin/bash
# timing A/B of experimental library builds (tools/exp/libexp*.so, same ABI) against the shipped one on the default workload
O=gpurun_out/${1:-exp}; mkdir -p $O
for v in base "$@"; do
  [ "$v" = "$1" ] && continue
  L=""; [ "$v" != base ] && L="$PWD/tools/exp/libexp$v.so"
  PEGNCDE_LIB=$L timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$v value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3))
except Exception as ex: print("$v failed:", ex)
PY
done
