#!/bin/bash
# usage: scripts/gpu_bench_quick.sh <tag> [extra bench args]  -- short bench lines only (no tests): value / kernel ms per workload
tag=$1; shift
timeout 150 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-episodes 4 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("main value %.3e e2e %.3e kernel_ms %.2f %s" % (j["value"], j["e2e"]["value"], j["roofline"]["kernel_ms"], j["config"]["kernel"]))
    for k, v in j["extra"].items():
        print(k, "value %.3e kernel_ms %.2f" % (v["value"], v["kernel_ms"]))
except Exception as ex:
    print("bench failed:", ex); print(open("gpurun_out/${tag}_bench.err").read()[-1500:])
PY
