#!/bin/bash
# usage: scripts/gpu_profile.sh <tag> [bench args...]
# plain run first (must exit 0), then the launch list and one full ncu capture of k_run; everything lands in gpurun_out/
tag=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_run -s 3 -c 1 -o gpurun_out/prof_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/plain_$tag.log | cut -c1-300
