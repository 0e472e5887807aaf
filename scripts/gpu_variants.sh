#!/bin/bash
# usage: scripts/gpu_variants.sh <workload> <lib...>  -- bench each prebuilt library variant (variants/*.so) on a workload
w=$1; shift
L=network-distributed-q-learning_b200/csrc/libswitchfl_b200.so
cp $L /tmp/lib_orig.so
for v in orig "$@"; do
  if [ $v != orig ]; then cp variants/$v $L; else cp /tmp/lib_orig.so $L; fi
  python bench.py --steps 5 --warmup 3 --no-cpu --workload $w 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $v lanes', d['config']['lanes_per_env'], 'value %.3e' % d['value'], 'ms %.3f' % d['ms_per_step'])" || echo "$w $v failed"
done
cp /tmp/lib_orig.so $L
