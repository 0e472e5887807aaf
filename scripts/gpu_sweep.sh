#!/bin/bash
# usage: scripts/gpu_sweep.sh <tag> <workload> "<args A>" "<args B>" ...   -- one short device-resident bench line per argument set
tag=$1; wl=$2; shift 2
i=0
for a in "$@"; do
  timeout 150 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu --extra= --e2e-episodes 1 $a > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python -c "
import json,sys
try:
    j=json.loads(open('gpurun_out/${tag}_$i.json').read().strip().splitlines()[-1]); print('$a', '-> %.3e' % j['value'], 'kernel_ms %.2f' % j['roofline']['kernel_ms'], j['config']['kernel'])
except Exception as ex: print('$a', 'FAILED', open('gpurun_out/${tag}_$i.err').read()[-300:])
"
  i=$((i+1))
done
