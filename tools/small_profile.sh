#!/bin/bash
# per-launch durations of the small-graph configurations (england, sir): where a 2 ms solver step goes.  Usage: bash tools/small_profile.sh <tag>
T=${1:-small}; shift; O=gpurun_out/$T; mkdir -p $O; W=${@:-"england sir"}
for w in $W; do
  B="python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --t1 0.2 --no-graph --no-sweep --no-tensor-peaks"
  timeout 200 $B > $O/bench_$w.json 2> $O/bench_$w.err || { echo "bench $w failed"; tail -5 $O/bench_$w.err; continue; }
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 800 --csv --log-file $O/launches_$w.csv $B > $O/ncu_$w.log 2>&1
done
ls -la $O
