#!/bin/bash
# usage: scripts/gpu_sweep.sh <workload> <lanes...>   -- decisions/s per lanes-per-env setting
w=$1; shift
for L in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu --workload $w --lanes $L 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w lanes', d['config']['lanes_per_env'], 'value %.3e' % d['value'], 'ms %.3f' % d['ms_per_step'])" || echo "$w lanes $L failed"
done
