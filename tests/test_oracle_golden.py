"""CPU: the standalone oracle must reproduce every golden vector recorded from the reference's own code.

Bars: integer fields bit-exact; rewards and Q-values bit-exact fp64 (same operator order)."""
import numpy as np
import pytest

from oracle.switchfl_oracle import SwitchFLOracle
from tests._util import golden_names, hparams, load_golden, q_dict, ref_tables


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_trace(name):
    fx, g = load_golden(name)
    tab = ref_tables(fx, g)
    o = SwitchFLOracle(fx, tab, seed=int(g["seed"]), **hparams(g))
    if "inject_events" in g:                                  # goldens recorded with an injected malfunction schedule
        o.rail_env.injected_malfunctions = {(int(t), int(h)): int(d) for t, h, d in g["inject_events"]}
    o.enable_trace()
    eps = o.learn(int(g["n_episodes"]))
    t = o.trace
    for k in ("dec_ep", "dec_tick", "dec_switch", "dec_train", "dec_action", "dec_next_switch", "dec_done", "dec_greedy"):
        assert np.array_equal(np.array(t[k], np.int32), g[k]), k
    assert np.array_equal(np.array(t["dec_obs"]).reshape(-1, 18), g["dec_obs"])
    assert np.array_equal(np.array(t["dec_mask"]).reshape(-1, 9), g["dec_mask"])
    assert np.array_equal(np.array(t["dec_reward"]), g["dec_reward"])
    assert np.array_equal(np.array(t["dec_arrived"], np.uint64), g["dec_arrived"])
    assert np.array_equal(np.array(t["dec_sem"]).reshape(g["dec_sem"].shape), g["dec_sem"])
    for k in ("tick_ep", "tick_tick", "tick_pos", "tick_dir", "tick_state", "tick_malf"):
        assert np.array_equal(np.array(t[k], np.int32).reshape(g[k].shape), g[k]), k
    assert [e["cum_reward"] for e in eps] == list(g["ep_cum_reward"])
    assert [e["arrived"] for e in eps] == list(g["ep_arrived"])
    assert [e["num_malfunctions"] for e in eps] == list(g["ep_num_malfunctions"])
    assert np.array_equal(np.array([e["delays"] for e in eps], np.float64), g["ep_delays"])
    assert o.q_table == q_dict(g["q_keys"], g["q_vals"])      # bit-exact fp64, identical key set


@pytest.mark.parametrize("name", golden_names())
def test_oracle_q_init(name):
    fx, g = load_golden(name)
    o = SwitchFLOracle(fx, ref_tables(fx, g), seed=int(g["seed"]), **hparams(g))
    o.reset(seed=o.seed)
    o.init_q_table()
    assert o.q_table == q_dict(g["qinit_keys"], g["qinit_vals"])


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replay_equals_free_run(name):
    """Feeding the recorded actions and malfunction events reproduces the same trajectory."""
    fx, g = load_golden(name)
    o = SwitchFLOracle(fx, ref_tables(fx, g), seed=int(g["seed"]), **hparams(g))
    ev = g["malf_events"]
    sched = {(int(t), int(h)): int(d) for (_, t, h, d) in ev}
    o.rail_env.injected_malfunctions = sched
    o.enable_trace()
    o.learn(int(g["n_episodes"]), replay_actions=g["dec_action"] | (g["dec_greedy"] << 6))
    assert np.array_equal(np.array(o.trace["tick_pos"], np.int32).reshape(g["tick_pos"].shape), g["tick_pos"])
    assert np.array_equal(np.array(o.trace["tick_malf"], np.int32).reshape(g["tick_malf"].shape), g["tick_malf"])
    assert np.array_equal(np.array(o.trace["dec_reward"]), g["dec_reward"])
    assert o.q_table == q_dict(g["q_keys"], g["q_vals"])


def test_f4_golden_takes_the_off_map_malfunction_to_stopped_transition():
    """Row F4: MALFUNCTION_OFF_MAP -> STOPPED when the counter runs out past the earliest
    departure and no valid movement is possible; the train is placed on its (occupied) entry cell."""
    _, g = load_golden("f4_offmap_7x7")
    st, pos = g["tick_state"], g["tick_pos"]
    k = np.nonzero((st[:-1, 1] == 2) & (st[1:, 1] == 4))[0]
    assert len(k), "transition not in the trace"
    k = int(k[0]) + 1
    assert pos[k, 1] == pos[k, 0] >= 0 and st[k, 0] == 5      # on train 0's cell, which is broken down there
