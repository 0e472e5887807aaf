#!/bin/bash
# usage: scripts/gpu_sweep_cta.sh <workload> <warps per CTA...>   (0 = library choice)
w=$1; shift
for L in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu --workload $w --cta-warps $L 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w cta_warps $L', 'value %.3e' % d['value'], 'ms %.3f' % d['ms_per_step'])" || echo "$w cta_warps $L failed"
done
