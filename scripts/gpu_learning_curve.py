#!/usr/bin/env python
"""Learning-quality sanity run (not a parity test): distributed Q-learning on a C3-class map (80x80, 15 trains, the
hyperparam_tuning.py hyper-parameters) for N episodes on many seeds at once; prints the mean number of arrived trains
and the mean cumulative reward per episode bucket.  The reference's curves (plot.ipynb) rise from ~6 to
~14.7 of 15 arrived trains over 10 000 episodes on its flatland map.

    gpurun -- python scripts/gpu_learning_curve.py [episodes] [envs] [c3|c4] > gpurun_out/learning_curve.txt
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package  # noqa: E402

load_package()
from switchfl_b200 import api, mapgen  # noqa: E402

n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
which = sys.argv[3] if len(sys.argv) > 3 else "c3"
fx = mapgen.c4_fixture() if which == "c4" else mapgen.c3_fixture(64)     # the C4 / C3 benchmark maps
env = api.ASyncSwitchEnv(api.RailEnv(fx), render_mode=None, max_steps=100_000, n_envs=B, q_cap=131072 if which == "c4" else 65536, ep_cap=256)
model = api.DistrQLearning(env=env, gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0, seed=64)
t0 = time.time()
model.learn(num_episodes=n_ep, out_dir=None, checkpoint_freq=10 ** 9, exploit_freq=None)
wall = time.time() - t0
m = model.metrics
arr, cum = m["arrived_trains"], m["cum_reward"]
T = env.rail_map.trains.T
print(f"map {fx['name']}: {fx['grid'].shape[0]}x{fx['grid'].shape[1]}, {T} trains, {env.rail_map.tab.S} switches; {B} environments (seeds 64..{64 + B - 1}), {n_ep} episodes each; "
      f"{model.total_decisions:.3e} decisions in {wall:.1f} s wall ({model.total_decisions / wall:.3e}/s incl. host logging)")
print(f"{'episodes':>14} {'arrived (mean)':>22} {'cum. reward (mean)':>20}")
edges = [0, 10, 50, 100, 250, 500, 1000, 2000, 3000, 5000, 10000]
for a, b in zip(edges[:-1], edges[1:]):
    if a >= n_ep:
        break
    b = min(b, n_ep)
    print(f"{a:>6}-{b:<7} {arr[:, a:b].mean():>22.2f} {cum[:, a:b].mean():>20.0f}")
r, a_, _ = model.test(out_dir=None, save_outputs=False, _batched=True)
print(f"greedy rollout after training: arrived {a_.mean():.2f} of {T} (mean over {B} envs), cum. reward {r.mean():.0f}")
