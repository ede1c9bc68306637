#!/bin/bash
# ncu evidence of the final kernels on the default bench workload (run under gpurun after bench.py has exited 0 without ncu)
mkdir -p gpurun_out/final4
O=gpurun_out/final4
(time python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --t1 0.2 --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file $O/launches_default.csv $B > $O/ncu_launches.log 2>&1
for k in k_tc_contractILi0 k_tc_contractILi1; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 12 -c 1 -o $O/full_$k -f $B > $O/ncu_full_$k.log 2>&1
done
ls -la $O
