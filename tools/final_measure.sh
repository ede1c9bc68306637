#!/bin/bash
# Round-end measurement suite (run under gpurun): tests, bench lines, ncu launch list + full captures.
mkdir -p gpurun_out/final
O=gpurun_out/final
(time timeout 600 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
(time python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
(time python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_reference.json 2>&1
for w in sweep_n1024_h64 sweep_n2048_h64 sweep_n4096_h128 sweep_n4096_h256 sweep_n8192_h128 twitter england sir; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample-steps 2 > $O/bench_$w.json 2> $O/bench_$w.err || echo "bench $w failed"
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1]); r=d["roofline"] or {}
    print("$w", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(r.get("frac",0),3), "bound", r.get("bound"), "fwd_us", round(r.get("fwd_avg_us",0),1), "bwd_us", round(r.get("bwd_avg_us",0),1), "share", round(r.get("share_of_step",0),2), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex: print("$w parse failed", ex)
PY
done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --t1 0.2 --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file $O/launches_default.csv $B > $O/ncu_launches.log 2>&1
for k in k_tc_contractILi0 k_tc_contractILi1 k_tc_norm_linear k_tc_linear_bwd; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 12 -c 1 -o $O/full_$k -f $B > $O/ncu_full_$k.log 2>&1
done
ls -la $O | head -40
