#!/bin/bash
# usage: tools/exp.sh WORKLOAD T1 [ENVVAR=VALUE ...]   -> one summary line per run
w=$1; t1=$2; shift 2
env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --workload $w --no-cpu-baseline --t1 $t1 --e2e-steps 1 2>gpurun_out/err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', d['config']['workload'], 'value', round(d['value'],1), 'frac', round(r['frac'],3), 'fwd_us', round(r['fwd_avg_us'],1), 'bwd_us', round(r['bwd_avg_us'],1), 'share', round(r['share_of_step'],2), 'ms', round(d['ms_per_step'],1))
" || tail -3 gpurun_out/err.log
