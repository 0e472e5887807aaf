"""CPU: tools/record_flatland_fixture.py -- the recorder meant to run where flatland is installed -- exercised against the
fixture-backed stand-in RailEnv of oracle/trainsim.py: the fixture it writes is the fixture the env was built from, and
(in the build container, where /root/reference exists) the trace it records is the committed golden, bit for bit."""
import configparser
import importlib.util
import os
import sys

import numpy as np
import pytest

from oracle import trainsim
from switchfl_b200 import mapgen
from tests._util import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _recorder():
    spec = importlib.util.spec_from_file_location("record_flatland_fixture", os.path.join(ROOT, "tools", "record_flatland_fixture.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cfg(fx, g=None):
    c = configparser.ConfigParser()
    c["MISC"] = {"random_seed": "7"}
    c["ENV"] = {"width": str(fx["grid"].shape[1]), "height": str(fx["grid"].shape[0]), "malfunction_rate": repr(float(fx["malfunction_rate"])),
                "min_duration": str(fx["min_duration"]), "max_duration": str(fx["max_duration"])}
    if g is not None:
        c["MODEL"] = {k[3:]: repr(float(g[k])) for k in g if k.startswith("hp_")}
    return c


@pytest.mark.parametrize("name", ["c1_synth18", "slips24_t6", "c3_rail80_s64"])
def test_recorded_fixture_round_trips(name, tmp_path):
    rec = _recorder()
    fx, _ = load_golden(name)
    got = rec.fixture_of(trainsim.RailEnv(fx), _cfg(fx), 7)
    path = str(tmp_path / "x.fixture.npz")
    rec.save_fixture(path, got)
    back = mapgen.load_fixture(path)                                   # the loader the backend uses
    for k in ("grid", "init_pos", "init_dir", "target", "earliest_departure", "latest_arrival"):
        assert np.array_equal(back[k], fx[k]), k
    for k in ("max_episode_steps", "malfunction_rate", "min_duration", "max_duration"):
        assert back[k] == fx[k], k
    mapgen.check_fixture(back)


@pytest.mark.skipif(not os.path.isdir("/root/reference/switchfl"), reason="needs the reference repository (build container only)")
def test_recorded_trace_equals_the_committed_golden():
    rec = _recorder()
    for p in (os.path.join(ROOT, "oracle", "shim"), "/root/reference"):
        if p not in sys.path:
            sys.path.insert(0, p)
    fx, g = load_golden("c1_synth18")
    out = rec.record_trace(trainsim.RailEnv(fx), fx, _cfg(fx, g), int(g["seed"]), int(g["n_episodes"]))
    for k in ("dec_switch", "dec_train", "dec_action", "dec_reward", "dec_obs", "dec_sem", "tick_pos", "tick_state", "q_keys", "q_vals", "ep_cum_reward"):
        assert np.array_equal(out[k], g[k], equal_nan=k == "q_vals"), k
