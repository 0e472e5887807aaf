#!/bin/bash
# usage: scripts/gpu_profile.sh <tag> <workload> [bench args...]
# plain short run first (must exit 0), then the launch list and ONE full ncu capture of a warm k_run launch of that workload;
# everything lands in gpurun_out/.  Numbers printed under ncu are never bench values.
tag=$1; wl=$2; shift 2
ARGS="--workload $wl --steps 2 --warmup 2 --no-cpu --extra= --e2e-episodes 1 $@"
python bench.py $ARGS > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $ARGS > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_run -s 2 -c 1 -o gpurun_out/${tag}_prof \
    python bench.py $ARGS > gpurun_out/${tag}_ncu.log 2>&1
tail -c 300 gpurun_out/${tag}_plain.log; ls -la gpurun_out/${tag}_prof.ncu-rep
