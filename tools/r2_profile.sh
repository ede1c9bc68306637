#!/bin/bash
# ncu evidence for the current kernels on the default bench workload.  Usage: bash tools/r2_profile.sh <tag> [extra bench args]
T=${1:-prof}; shift; O=gpurun_out/$T; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --t1 0.2 --no-graph $*"
timeout 300 $B > $O/bench_short.json 2> $O/bench_short.err || { echo "bench failed"; tail -5 $O/bench_short.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
for k in k_tc_contractILi0 k_tc_contractILi1; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 12 -c 1 -o $O/full_$k -f $B > $O/ncu_full_$k.log 2>&1
done
ls -la $O
