#!/bin/bash
# usage: scripts/gpu_round.sh <tag>   -- GPU parity tests, the default bench line (as the driver runs it), then the ncu launch
# list and one full capture of the headline k_run on a short run of the same workload; everything lands in gpurun_out/
tag=$1
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${tag}_gputest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
SHORT="--workload c4 --steps 1 --warmup 1 --no-cpu --extra= --e2e-episodes 2"
python bench.py $SHORT > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $SHORT > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_run -s 1 -c 1 -o gpurun_out/${tag}_prof \
    python bench.py $SHORT > gpurun_out/${tag}_ncu.log 2>&1
tail -c 600 gpurun_out/${tag}_bench.json
