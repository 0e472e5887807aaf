#!/bin/bash
# host BFS (numpy/python, railmap.distance_to) vs the k_distance_map kernel on the C4 map (50 targets, 100x100)
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
from __graft_entry__ import load_package
load_package()
import numpy as np
from switchfl_b200 import backend, mapgen, railmap
fx = mapgen.load_fixture("tests/golden/c4_synth100_t50.fixture.npz")
rm = backend.RailMap(fx)
W = fx["grid"].shape[1]
t0 = time.time(); host = np.stack([railmap.distance_to(fx["grid"], (int(c) // W, int(c) % W)) for c in rm.trains.targets]); th = time.time() - t0
backend.device_distance_map(fx["grid"], rm.trains.targets)            # warm-up (context, module load)
t0 = time.time(); dev = backend.device_distance_map(fx["grid"], rm.trains.targets); td = time.time() - t0
print(f"distance map, C4 map ({len(rm.trains.targets)} targets x 100x100x4): host BFS {th*1e3:.0f} ms, sfl_distance_map {td*1e3:.1f} ms "
      f"(H2D + kernel + D2H), equal: {np.array_equal(host, dev)}")
PY
