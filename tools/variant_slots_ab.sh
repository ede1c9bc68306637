#!/bin/bash
# Round-2 opener: A/B of -DPEG_TC_VARIANT_SLOTS=1 (per-variant A-slot barriers in the adjoint contraction, DESIGN.md (f) item 1;
# written in round 1, never run).  Run under gpurun:  bash tools/variant_slots_ab.sh
# Builds the variant next to the default library, runs the GPU parity suite and the default bench on both, restores the default.
set -u
O=gpurun_out/variant_slots; mkdir -p $O
C=perm_equiv_graph_neural_cdes_b200/csrc
L=perm_equiv_graph_neural_cdes_b200/libpegncde.so
cp $L /tmp/base.so
(cd $C && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPEG_TC_VARIANT_SLOTS=1 -shared -o /tmp/variant.so pegncde.cu peg_tc.cu) > $O/build.log 2>&1 || { echo "variant build failed"; tail -5 $O/build.log; exit 1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
for v in base variant; do
  cp /tmp/$v.so $L
  if [ $v = variant ]; then (timeout 300 python -m pytest tests -m gpu -q -x) > $O/pytest_$v.log 2>&1; tail -2 $O/pytest_$v.log; fi
  timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$v", "value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "frac", round(r["frac"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex: print("$v failed:", ex)
PY
done
cp /tmp/base.so $L
