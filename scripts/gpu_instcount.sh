#!/bin/bash
# usage: scripts/gpu_instcount.sh <workload> <lanes...>  -- warp instructions, active threads/inst and duration of one k_run launch
w=$1; shift
for L in "$@"; do
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:k_run -s 3 -c 1 --csv python bench.py --steps 2 --warmup 3 --no-cpu --workload $w --lanes $L 2>/dev/null \
   | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
print('$w lanes $L', {r[-3]: r[-1] for r in rows})"
done
