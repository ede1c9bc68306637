#!/bin/bash
# A/B of the small-graph cluster kernels against the per-operator pipeline on the small configurations.  Usage: bash tools/small_ab.sh <tag> [workloads...]
T=${1:-smallab}; shift; O=gpurun_out/$T; mkdir -p $O
W=${@:-"england sir"}
for w in $W; do
  for off in 0 1; do
    B="python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-sweep --no-tensor-peaks"
    if [ $off = 1 ]; then export PEG_SMALL_OFF=1; else unset PEG_SMALL_OFF; fi
    timeout 200 $B > $O/bench_${w}_off$off.json 2> $O/bench_${w}_off$off.err || { echo "bench $w off=$off failed"; tail -5 $O/bench_${w}_off$off.err; continue; }
    python - <<PY
import json
d=json.loads(open("$O/bench_${w}_off$off.json").read().strip().splitlines()[-1])
print("$w", "per-operator" if $off else "fused-small", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), "launches", d["gpu_launches"], "grad_check", d.get("grad_check"))
PY
  done
done
