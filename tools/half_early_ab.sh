#!/bin/bash
# A/B of PEG_TC_BWD_HALF_EARLY (adjoint converters issue part of the next item's plane loads before the slot wait)
O=gpurun_out/half_early; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
L=perm_equiv_graph_neural_cdes_b200/libpegncde.so
cp $L /tmp/base.so
for v in 1 2; do   # the variants are built on the box (nvcc is in the image); round 1 shipped them prebuilt
  (cd perm_equiv_graph_neural_cdes_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPEG_TC_BWD_HALF_EARLY=$v -shared -o /tmp/exp$v.so pegncde.cu peg_tc.cu) > $O/build_$v.log 2>&1
done
for v in base 1 2; do
  if [ $v = base ]; then cp /tmp/base.so $L; else cp /tmp/exp$v.so $L; fi
  timeout 200 $B > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("$v", "value", round(d["value"],1), "fwd_us", round(r["fwd_avg_us"],1), "bwd_us", round(r["bwd_avg_us"],1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
cp /tmp/base.so $L
