#!/usr/bin/env python
"""Dump a map fixture (+ optionally a full reference trace) from a GENUINE flatland install.

Run this where ``flatland-rl`` and the reference repository are importable -- NOT in the build container, which has
neither.  It writes the ``.fixture.npz`` this backend loads (mapgen.load_fixture) from the reference's own
``RailEnv`` after ``reset``, so a flatland-generated map (main.py:36-49) can be trained on the GPU:

    python tools/record_flatland_fixture.py --config config.ini --out my_map.fixture.npz
    # then:  [ENV] fixture = my_map.fixture.npz   in the config.ini given to  python -m switchfl_b200.cli

With ``--trace OUT.npz`` (and the reference repository on PYTHONPATH) it also runs the reference's ``learn()`` with
the tracing hooks of oracle/gen_golden.py and stores the golden-vector file; dropping both files into tests/golden/
turns every "[UPSTREAM-UNVERIFIED]" row of SURVEY.md Appendix B into a pinned one.
"""
import argparse
import configparser

import numpy as np


def build_rail_env(cfg):
    from flatland.envs.line_generators import sparse_line_generator
    from flatland.envs.malfunction_generators import MalfunctionParameters, ParamMalfunctionGen
    from flatland.envs.rail_env import RailEnv
    from flatland.envs.rail_generators import sparse_rail_generator
    seed = int(cfg["MISC"]["random_seed"])
    e = cfg["ENV"]
    mf = ParamMalfunctionGen(MalfunctionParameters(malfunction_rate=float(e["malfunction_rate"]), min_duration=int(e["min_duration"]),
                                                   max_duration=int(e["max_duration"])))
    return RailEnv(width=int(e["width"]), height=int(e["height"]),
                   rail_generator=sparse_rail_generator(max_num_cities=int(e["max_num_cities"]), grid_mode=True,
                                                        max_rails_between_cities=int(e["max_rails_between_cities"]),
                                                        max_rail_pairs_in_city=int(e["max_rail_pairs_in_city"]), seed=seed),
                   line_generator=sparse_line_generator(seed=seed), number_of_agents=int(e["number_of_agents"]),
                   malfunction_generator=mf), seed


def fixture_of(rail_env, cfg, seed):
    rail_env.reset(random_seed=seed)
    agents = sorted(rail_env.agents, key=lambda a: (a.initial_position, a.initial_direction))       # switch_env.py:104-119
    e = cfg["ENV"]
    return {
        "name": f"flatland_{e['width']}x{e['height']}_t{len(agents)}_s{seed}",
        "grid": np.asarray(rail_env.rail.grid, np.uint16),
        "init_pos": np.array([a.initial_position for a in agents], np.int32),
        "init_dir": np.array([int(a.initial_direction) for a in agents], np.int32),
        "target": np.array([a.target for a in agents], np.int32),
        "earliest_departure": np.array([a.earliest_departure for a in agents], np.int32),
        "latest_arrival": np.array([a.latest_arrival for a in agents], np.int32),
        "max_episode_steps": int(rail_env._max_episode_steps),
        "malfunction_rate": float(e["malfunction_rate"]), "min_duration": int(e["min_duration"]), "max_duration": int(e["max_duration"]),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, help="the reference's config.ini (hyperparam_tuning.py:51-78)")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    cfg = configparser.ConfigParser()
    cfg.read(args.config)
    rail_env, seed = build_rail_env(cfg)
    fx = fixture_of(rail_env, cfg, seed)
    np.savez_compressed(args.out, **{k: (np.array(v) if not isinstance(v, np.ndarray) else v) for k, v in fx.items()})
    print(f"wrote {args.out}: grid {fx['grid'].shape}, {len(fx['init_dir'])} trains, max_episode_steps {fx['max_episode_steps']}")


if __name__ == "__main__":
    main()
