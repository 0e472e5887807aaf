"""Replay-parity checks shared by the CPU logic tests (tests/emul build) and the GPU tests (-m gpu).

The engine under test is fed the action stream and the malfunction events recorded from the REFERENCE's
own run (tests/golden) and must reproduce it: integer fields bit-exact, Q-values bit-exact fp64 when
lr_decay_rate == 1.0 and within 1e-12 relative otherwise (device pow vs glibc pow; BASELINE.json)."""
from __future__ import annotations

import numpy as np

from switchfl_b200 import backend
from tests._util import hparams, load_golden, q_dict


def unpad(pos, W):
    pos = np.asarray(pos, np.int64)
    Wp = W + 2
    return np.where(pos < 0, -1, (pos // Wp - 1) * W + (pos % Wp - 1))


def golden_events(g):
    ev = g["malf_events"]
    return np.unique(ev[:, 1:], axis=0) if len(ev) else np.zeros((0, 3), np.int32)


def replay_stream(g):
    """recorded actions, bit 6 set where the reference learner exploited (max_action consulted)"""
    return (g["dec_action"].astype(np.int8) | (g["dec_greedy"].astype(np.int8) << 6)).astype(np.int8)


def make_replay_engine(name, factory, n_envs=2):
    fx, g = load_golden(name)
    rm = backend.RailMap(fx)
    n_dec, n_tick, n_ep = len(g["dec_action"]), len(g["tick_ep"]), int(g["n_episodes"])
    ev = golden_events(g)
    q_cap = 1 << max(6, int(np.ceil(np.log2(len(g["q_keys"]) * 2 + 16))))
    eng = factory(rm, n_envs=n_envs, q_cap=q_cap, dec_cap=n_dec + 4, tick_cap=n_tick + 4, ep_cap=n_ep + 1,
                  act_cap=n_dec + 4, ev_cap=len(ev) + 2, trace_sem=True)
    eng.set_hparams(**hparams(g), seeds=np.arange(n_envs) + int(g["seed"]), episodes=n_ep)
    eng.set_replay([replay_stream(g)] * n_envs, [ev] * n_envs)
    eng.reset()
    eng.enable_q_init(True)
    return fx, g, rm, eng


def compare_trace(name, rm, g, dec, tick, sem):
    W = rm.tab.W
    assert len(dec) == len(g["dec_action"]) and len(tick) == len(g["tick_ep"]), (len(dec), len(tick))
    for k_mine, k_g in (("ep", "dec_ep"), ("tick", "dec_tick"), ("sw", "dec_switch"), ("train", "dec_train"),
                        ("action", "dec_action"), ("next_sw", "dec_next_switch"), ("done", "dec_done")):
        assert np.array_equal(dec[k_mine], g[k_g]), (name, k_mine)
    assert np.array_equal(dec["reward"].astype(np.float64), g["dec_reward"]), "rewards"
    assert np.array_equal(dec["arrived"], g["dec_arrived"]), "arrived"
    for i in range(len(dec)):
        obs = rm.key_to_obs(int(dec["key"][i]))
        gobs = tuple(int(x) for x in g["dec_obs"][i] if x != -9)
        assert obs == gobs, (name, i, obs, gobs)
        gm = [int(x) for x in g["dec_mask"][i] if x >= 0]
        assert [(int(dec["mask"][i]) >> a) & 1 for a in range(len(gm))] == gm, (name, i, "mask")
    if sem is not None:
        gs = g["dec_sem"]                        # (train, type, t0, t1); device record is {t0, t1, train, type}
        present = gs[:, :, 0] >= 0
        assert np.array_equal(sem[:, :, 2] >= 0, present), "semaphore presence"
        assert np.array_equal(sem[:, :, 2][present], gs[:, :, 0][present])
        assert np.array_equal(sem[:, :, 3][present], gs[:, :, 1][present])
        assert np.array_equal(sem[:, :, 0][present], gs[:, :, 2][present])
        assert np.array_equal(sem[:, :, 1][present], gs[:, :, 3][present])
    assert np.array_equal(unpad(tick["pos"], W), g["tick_pos"]), "positions"
    assert np.array_equal(tick["dir"], g["tick_dir"]), "directions"
    assert np.array_equal(tick["state"], g["tick_state"]), "states"
    assert np.array_equal(tick["malf"], g["tick_malf"]), "malfunction counters"


def check_replay(name, factory, n_envs=2, chunk=None):
    fx, g, rm, eng = make_replay_engine(name, factory, n_envs)
    total = len(g["tick_ep"]) + 8
    if chunk is None:
        eng.run(backend.MODE_REPLAY, total)
    else:                         # many short launches must equal one long one
        for _ in range(0, total + chunk, chunk):
            eng.run(backend.MODE_REPLAY, chunk)
    c = eng.counters()
    assert (c["err"] == 0).all(), c["err"]
    assert (c["halted"] == 1).all()
    assert (c["episodes"] == int(g["n_episodes"])).all()
    assert (c["decisions"] == len(g["dec_action"])).all()
    for env in range(n_envs):
        dec, tick, sem = eng.trace(env)
        compare_trace(name, rm, g, dec, tick, sem)
    n, log, delays = eng.episode_log()
    n_ep = int(g["n_episodes"])
    assert (n == n_ep).all()
    for env in range(n_envs):
        assert np.array_equal(log[env, :n_ep]["cum_reward"], g["ep_cum_reward"])
        assert np.array_equal(log[env, :n_ep]["arrived"], g["ep_arrived"])
        assert np.array_equal(log[env, :n_ep]["num_malfunctions"], g["ep_num_malfunctions"])
        assert np.array_equal(delays[env, :n_ep].astype(np.float64), g["ep_delays"])
    gq = q_dict(g["q_keys"], g["q_vals"])
    exact = float(g["hp_lr_decay_rate"]) == 1.0
    for env in range(n_envs):
        q = eng.export_q(env, include_init=True)
        assert set(q) == set(gq), (len(q), len(gq))
        for k, row in gq.items():
            if exact:
                assert q[k] == row, (k, q[k], row)
            else:
                assert np.allclose(q[k], row, rtol=1e-12, atol=0.0), (k, q[k], row)
    eng.close()
    return True


def check_against_oracle(factory, fx, hp, n_ep, seeds, max_steps=100_000, greedy_after=False, q_cap=4096):
    """The oracle free-runs learn() (and optionally one greedy test() rollout) per seed; the engine replays each
    env's action / malfunction stream and must agree on every decision, the episode metrics and the Q-table --
    including episodes cut short by max_steps (switch_env.py:652-657) and the rows test() inserts (distr_q.py:211)."""
    from oracle.switchfl_oracle import SwitchFLOracle
    rm = backend.RailMap(fx)
    B = len(seeds)
    oracles, acts, evs, n_learn = [], [], [], []
    kept = []
    for sd in seeds:
        o = SwitchFLOracle(fx, rm.tab, seed=int(sd), max_steps=max_steps, **hp)
        o.enable_trace()
        try:
            eps = o.learn(n_ep)
            n_dec_learn = len(o.trace["dec_action"])
            if greedy_after:
                o.episode = n_ep
                eps.append(o.test())
        except RuntimeError as ex:                 # the reference dies here too (observer.py:294-307); covered by its own test
            if "No train detected" not in str(ex):
                raise
            continue
        kept.append(sd)
        n_learn.append(n_dec_learn)
        o.eps = eps
        oracles.append(o)
        acts.append(o.replay_stream(0, n_learn[-1]))
        evs.append(o.malfunction_schedule())
    seeds, B = kept, len(kept)
    if not B:
        import pytest
        pytest.skip("the reference itself fails on this map for every seed (observer.py:294-307)")
    n_dec = max(len(o.trace["dec_action"]) for o in oracles)
    eng = factory(rm, n_envs=B, q_cap=q_cap, max_steps=max_steps, dec_cap=n_dec + 4, act_cap=n_dec + 4,
                  ev_cap=max(len(e) for e in evs) + 2, ep_cap=n_ep + 2)
    eng.set_hparams(**hp, seeds=np.asarray(seeds), episodes=n_ep)
    eng.set_replay(acts, evs)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_REPLAY, 1_000_000)
    eng.check_errors()
    assert (eng.counters()["halted"] == 1).all()
    if greedy_after:
        eng.set_hparams(**hp, seeds=np.asarray(seeds), episodes=1, episode_base=n_ep)
        eng.reset(keep_q=True, keep_interactions=True)
        eng.run(backend.MODE_GREEDY, 1_000_000)
        eng.check_errors()
    _, log, delays = eng.episode_log()
    for i, o in enumerate(oracles):
        dec, _, _ = eng.trace(i)
        t = o.trace
        first = n_learn[i] if greedy_after else 0                      # the trace buffers restart with sfl_reset
        assert len(dec) == len(t["dec_action"]) - first, (i, len(dec), len(t["dec_action"]), first)
        for k_mine, k_o in (("sw", "dec_switch"), ("train", "dec_train"), ("action", "dec_action"), ("next_sw", "dec_next_switch"),
                            ("tick", "dec_tick"), ("done", "dec_done")):
            assert np.array_equal(dec[k_mine], np.array(t[k_o][first:])), (i, k_mine)
        assert np.array_equal(dec["reward"].astype(np.float64), np.array(t["dec_reward"][first:])), i
        eps = o.eps[n_ep:] if greedy_after else o.eps
        for e_i, ep in enumerate(eps):
            assert log[i, e_i]["cum_reward"] == ep["cum_reward"] and log[i, e_i]["decisions"] == ep["decisions"], (i, e_i)
            assert log[i, e_i]["ticks"] == ep["ticks"], (i, e_i)
            if ep["decisions"]:                    # with no decision at all the reference has no post_step_info to count (distr_q.py:364)
                assert log[i, e_i]["arrived"] == ep["arrived"] == bin(int(log[i, e_i]["arrived_mask"])).count("1"), (i, e_i)
            assert log[i, e_i]["num_malfunctions"] == ep["num_malfunctions"], (i, e_i)
            assert list(delays[i, e_i]) == [int(x) for x in ep["delays"]], (i, e_i)
        q = eng.export_q(i, include_init=True)
        assert q == o.q_table, (i, len(q), len(o.q_table))
    eng.close()
