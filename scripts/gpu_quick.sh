#!/bin/bash
# usage: scripts/gpu_quick.sh   -- GPU parity tests + short C2 / C4 bench lines (value, e2e, ms)
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'lanes', d['config']['lanes_per_env'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'ms %.3f' % d['ms_per_step'], 'aborted', d['config']['episodes_abandoned_rank0'], '/', d['config']['episodes_rank0'])" "$1"; }
python bench.py --steps 50 --warmup 3 --no-cpu 2>&1 | show c2
python bench.py --steps 50 --warmup 3 --no-cpu --lanes 16 2>&1 | show c2-16
python bench.py --steps 50 --warmup 3 --no-cpu --lanes 8 2>&1 | show c2-8
python bench.py --steps 5 --warmup 3 --no-cpu --workload c4 2>&1 | show c4
