#!/bin/bash
# usage: scripts/gpu_round.sh <tag>   -- GPU parity tests, the default bench line (as the driver runs it), then the launch list
# and one full ncu capture of a warm headline k_run launch (same workload, same ticks per step, fewer steps)
tag=$1
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${tag}_gputest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
bash scripts/gpu_profile.sh ${tag}_c4 c4
tail -c 400 gpurun_out/${tag}_bench.json; cat gpurun_out/${tag}_gputest.log
