#!/bin/bash
# usage: scripts/gpu_quick.sh <tag> [pytest -k expr]  -- GPU parity tests + short bench lines of every workload (value, e2e, kernel ms)
tag=$1
timeout 400 python -m pytest tests -m gpu -q -x ${2:+-k "$2"} 2>&1 | tail -8 > gpurun_out/${tag}_gputest.log
timeout 150 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-episodes 8 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("c4 value %.3e e2e %.3e kernel_ms %.2f" % (j["value"], j["e2e"]["value"], j["roofline"]["kernel_ms"]))
    for k, v in j["extra"].items():
        print(k, "value %.3e kernel_ms %.2f" % (v["value"], v["kernel_ms"]))
except Exception as ex:
    print("bench failed:", ex)
PY
cat gpurun_out/${tag}_gputest.log
