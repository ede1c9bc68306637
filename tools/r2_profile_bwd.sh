#!/bin/bash
# ncu --set full of one adjoint-contraction launch for the given operand formats.  Usage: bash tools/r2_profile_bwd.sh <tag> fmt...
T=$1; shift; O=gpurun_out/$T; mkdir -p $O
for f in "$@"; do
  B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --t1 0.2 --no-graph --operands $f"
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tc_contractILi1 -s 12 -c 1 -o $O/bwd_$f -f $B > $O/ncu_bwd_$f.log 2>&1
done
ls -la $O
