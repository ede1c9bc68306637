#!/bin/bash
# A/B of the opt-in LIGHT adjoint (PEG_TC_ADJ_LIGHT=1) against the default adjoint: parity tests, per-evaluation gradient errors, bench
O=gpurun_out/light; mkdir -p $O
(time timeout 200 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
timeout 120 python tools/probes/light_probe.py > $O/probe.log 2>&1; cat $O/probe.log | tail -20
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
PEG_TC_ADJ_LIGHT=1 timeout 200 $B > $O/bench_light.json 2> $O/bench_light.err; tail -c 900 $O/bench_light.json
timeout 200 $B > $O/bench_full.json 2> $O/bench_full.err; tail -c 900 $O/bench_full.json
