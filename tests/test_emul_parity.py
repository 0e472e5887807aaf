"""CPU: the kernel LOGIC (host build of the device sources, tests/emul) against the reference goldens."""
import os
import subprocess

import pytest

from switchfl_b200 import backend
from tests._parity import check_replay
from tests._util import golden_names

from tests.emulated import EmulEngine, EmulSwitchEnv, build_emul, emul_distance_map  # noqa: F401


@pytest.fixture(scope="session")
def emul_lib():
    return build_emul()


@pytest.mark.parametrize("name", golden_names())
def test_emul_replay_matches_reference(name, emul_lib):
    check_replay(name, EmulEngine, n_envs=2)


def test_emul_chunked_launches_equal_one_launch(emul_lib):
    check_replay("slips24_t6", EmulEngine, n_envs=1, chunk=7)


def test_emul_abandons_episode_exactly_where_the_reference_raises(emul_lib):
    """Congested C4-class map, free-running learn: some episodes reach observer.py:294-307 ("No train detected at
    active switch"), where the reference dies on an unbound local.  The kernel must flag the env and abandon the
    episode at exactly that decision: the oracle, fed the kernel's own action and malfunction stream, agrees on
    every decision before it and raises at the same one."""
    import numpy as np
    from oracle.switchfl_oracle import SwitchFLOracle
    from tests._util import load_golden
    fx, _ = load_golden("c4_synth100_t50")
    rm = backend.RailMap(fx)
    hp = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)
    B, n_ep, T = 8, 6, 50
    eng = EmulEngine(rm, n_envs=B, q_cap=16384, dec_cap=60000, tick_cap=5000, ep_cap=8)
    eng.set_hparams(**hp, seeds=np.arange(B) + 450565, episodes=n_ep)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100000)
    c = eng.counters()
    assert ((c["err"] & ~backend.ERR_NO_TRAIN_AT_SWITCH) == 0).all() and (c["halted"] == 1).all() and (c["episodes"] == n_ep).all()
    assert ((c["aborted"] > 0) == (c["err"] != 0)).all()
    hit = np.nonzero(c["aborted"])[0]
    assert len(hit), "this map/seed set is known to reach the reference's failure point"
    env = int(hit[0])
    dec, tick, _ = eng.trace(env)
    _, log, _ = eng.episode_log()
    # the malfunction draws are a function of (tick, train, seed): every episode sees the same schedule
    sched, prev = {}, np.zeros(T, np.int64)
    for i in range(int(log[env, 0]["ticks"])):
        m = tick["malf"][i]
        for h in range(T):
            if prev[h] == 0 and m[h] > 0:
                sched[(i + 1, h)] = int(m[h]) + 1
        prev = m
    o = SwitchFLOracle(fx, rm.tab, seed=1, **hp)
    o.rail_env.injected_malfunctions = sched
    o.enable_trace()
    with pytest.raises(RuntimeError, match="No train detected at active switch"):
        o.learn(n_ep, replay_actions=dec["action"])
    n = len(o.trace["dec_action"])
    assert 0 < n < len(dec)
    assert dec["ep"][n] == dec["ep"][n - 1] + 1, "the kernel starts the next episode right after the failing decision"
    for k_mine, k_o in (("ep", "dec_ep"), ("tick", "dec_tick"), ("sw", "dec_switch"), ("train", "dec_train"), ("next_sw", "dec_next_switch")):
        assert np.array_equal(dec[k_mine][:n], np.array(o.trace[k_o])), k_mine
    assert np.array_equal(dec["reward"][:n].astype(np.float64), np.array(o.trace["dec_reward"]))
    eng.close()


# ---------------------------------------------------------------------------------------------- the AEC protocol
def _aec_env(name, emul_lib, n_envs=1):
    """ASyncSwitchEnv on the host build, malfunction schedule of the golden bound as replay events."""
    from switchfl_b200 import api
    from tests._parity import golden_events
    from tests._util import load_golden
    fx, g = load_golden(name)
    ev = golden_events(g)
    env = EmulSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=n_envs, q_cap=64, ep_cap=2,
                        _engine_kwargs={"ev_cap": len(ev) + 2})
    env.engine.set_replay(None, [ev] * n_envs)
    return env, g


@pytest.mark.parametrize("name", ["loop_chord_7x7", "c1_synth18", "slips24_t6"])
def test_emul_aec_protocol_matches_reference(name, emul_lib):
    check_aec(*_aec_env(name, emul_lib))


def check_aec(env, g):
    """Drive reset / agent_iter / last / step exactly like distr_q.py:296-362 with the recorded actions: every
    observation, mask, reward, next switch and arrival list must be the reference's -- and a host-side learner running
    the reference's update rule on top of this env ends with the reference's Q-table."""
    import numpy as np
    from tests._util import hparams, q_dict
    hp = hparams(g)
    rm = env.rail_map
    q = rm.q_init_rows(hp["default_q"])                                  # distr_q.py:299-300
    n_act = {tuple(c): int(a) for c, a in zip(rm.tab.switch_cells, rm.tab.sw_A)}

    def row(obs):
        return q.setdefault(obs, [hp["default_q"]] * n_act[(obs[0], obs[1])])     # __check_entry, distr_q.py:47-57

    ninter = {a: 0 for a in env.possible_agents}
    i = 0
    for ep in range(int(g["n_episodes"])):
        env.reset(seed=int(g["seed"]))
        update_dict, at_dest, cum = {}, [], 0.0
        for agent in env.agent_iter():
            obs, R, term, trunc, info = env.last()
            assert not (term or trunc)
            train = info["active_train"]
            o = tuple(int(x) for x in obs)
            assert o == tuple(int(x) for x in g["dec_obs"][i] if x != -9), (i, o)
            assert list(info["action_mask"]) == [int(x) for x in g["dec_mask"][i] if x >= 0], i
            reward = R[train]
            assert reward == g["dec_reward"][i], i
            assert env.rail_env._elapsed_steps == g["dec_tick"][i] and agent == env.possible_agents[int(g["dec_switch"][i])]
            action = int(g["dec_action"][i])
            post = env.step(action)
            nxt = env.possible_agents.index("switch_%d-%d" % post["next_switch"])
            assert nxt == g["dec_next_switch"][i], i
            assert sum(1 << h for h in post["arrived_trains"]) == int(g["dec_arrived"][i]), i
            this_sw = env.possible_agents.index(agent)
            if (this_sw, train) in update_dict:                                                  # distr_q.py:329-338
                pobs, pact, pagent = update_dict.pop((this_sw, train))
                lr = hp["lr"] * hp["lr_decay_rate"] ** ninter[pagent]
                r_ = row(pobs)
                if agent != pagent:
                    r_[pact] = (1 - lr) * r_[pact] + lr * (reward + hp["gamma"] * max(row(o)))
                else:
                    r_[pact] = (1 - lr) * r_[pact] + lr * reward
            update_dict[(nxt, train)] = (o, action, agent)                                       # :340-342
            for h in post["arrived_trains"]:                                                     # :345-356
                if h not in at_dest:
                    at_dest.append(h)
                    for k in [k for k in update_dict if k[1] == h]:
                        pobs, pact, pagent = update_dict.pop(k)
                        lr = hp["lr"] * hp["lr_decay_rate"] ** ninter[pagent]
                        r_ = row(pobs)
                        r_[pact] = (1 - lr) * r_[pact] + lr * 1000.0
            cum += reward
            ninter[agent] += 1
            i += 1
        assert env.terminated or env.truncated
        assert cum == g["ep_cum_reward"][ep]
    assert i == len(g["dec_action"])
    gq = q_dict(g["q_keys"], g["q_vals"])
    exact = float(g["hp_lr_decay_rate"]) == 1.0
    touched = {k: v for k, v in q.items()}
    assert set(gq) <= set(touched) | set(gq)
    for k, v in gq.items():
        assert k in touched, k
        if exact:
            assert touched[k] == v, (k, touched[k], v)
        else:
            assert np.allclose(touched[k], v, rtol=1e-12, atol=0), (k, touched[k], v)


# ---------------------------------------------------------------------------------------------- edge cases vs the free-running oracle
HP_EDGE = dict(gamma=0.9, epsilon=0.6, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=1.0, default_q=0.5)


def edge_fixture(kind):
    from switchfl_b200 import mapgen
    if kind == "one_train":
        return mapgen.make_fixture(18, 1, 4, seed=3, num_cities=2, malfunction_rate=0.05, min_duration=2, max_duration=4)
    if kind == "max_trains":                                   # SFL_MAX_T = 64: both words of every train bit-set
        return mapgen.make_fixture(64, 64, 40, seed=3, num_cities=8, malfunction_rate=0.02, min_duration=3, max_duration=9, p_slip=0.4)
    raise KeyError(kind)


def test_emul_truncation_by_max_steps(emul_lib):
    """switch_env.py:652-657: the episode is truncated once more than max_steps decisions were taken."""
    from tests._parity import check_against_oracle
    from tests._util import load_golden
    fx, _ = load_golden("slips24_t6")
    check_against_oracle(EmulEngine, fx, HP_EDGE, 3, [5, 6], max_steps=25)


def test_emul_greedy_rollout_after_training(emul_lib):
    """distr_q.py:184-241 test(): greedy rollout, no updates, but default rows are inserted on lookup."""
    from tests._parity import check_against_oracle
    from tests._util import load_golden
    fx, _ = load_golden("slips24_t6")
    check_against_oracle(EmulEngine, fx, HP_EDGE, 2, [11, 12], greedy_after=True)


@pytest.mark.parametrize("kind", ["one_train", "max_trains"])
def test_emul_train_count_extremes(kind, emul_lib):
    from tests._parity import check_against_oracle
    check_against_oracle(EmulEngine, edge_fixture(kind), HP_EDGE, 1, [21, 22],
                         q_cap=65536 if kind == "max_trains" else 1024)


def fuzz_case(i):
    """Deterministic pseudo-random (map, timetable, hyper-parameters) case i."""
    import numpy as np
    from switchfl_b200 import mapgen
    r = np.random.RandomState(1000 + i)
    n = int(r.choice([12, 16, 20, 28, 36]))
    trains = int(r.randint(1, 10))
    fx = mapgen.make_fixture(n, trains, int(r.randint(2, 3 + n // 2)), seed=int(r.randint(10 ** 6)), num_cities=int(r.randint(2, 6)),
                             malfunction_rate=float(r.choice([0.0, 0.01, 0.08])), min_duration=int(r.randint(1, 4)),
                             max_duration=int(r.randint(4, 12)), p_slip=float(r.choice([0.0, 0.3, 0.8])))
    hp = dict(gamma=float(r.choice([1.0, 0.9, 0.5])), epsilon=float(r.choice([0.2, 0.5, 1.0])), epsilon_decay_rate=float(r.choice([1.0, 0.999, 0.9])),
              lr=float(r.choice([0.1, 0.5, 1.0])), lr_decay_rate=1.0, default_q=float(r.choice([0.0, -3.5, 200.0])))
    return fx, hp, [int(x) for x in r.randint(0, 10 ** 6, size=2)], int(r.choice([100_000, 100_000, 40]))


@pytest.mark.parametrize("i", range(12))
def test_emul_fuzz_maps_against_oracle(i, emul_lib):
    """Random maps / timetables / hyper-parameters: oracle free run vs engine replay, incl. a greedy rollout."""
    from tests._parity import check_against_oracle
    try:
        fx, hp, seeds, max_steps = fuzz_case(i)
    except ValueError as ex:                                    # "map too small for the requested number of trains"
        pytest.skip(str(ex))
    check_against_oracle(EmulEngine, fx, hp, 3, seeds, max_steps=max_steps,
                         greedy_after=True, q_cap=16384)


@pytest.mark.parametrize("name", golden_names())
def test_emul_distance_map_matches_vendored_reference(name, emul_lib):
    """Row F6: the relaxation of sfl_distance_map against the distance map recorded from the reference's vendored
    flatland_patch/distance_map.py (golden field ``dist``, per train) and against the host BFS."""
    import numpy as np
    from tests._util import load_golden
    fx, g = load_golden(name)
    rm = backend.RailMap(fx)
    d = emul_distance_map(fx["grid"], rm.trains.targets)
    assert np.array_equal(d, rm.trains.dist)
    assert np.array_equal(d[rm.trains.tgt_index], g["dist"])                 # golden: one map per train handle


def test_emul_malfunction_draws_have_the_flatland_distribution(emul_lib):
    """Row F5 in free-running mode (two-stage Philox draw): per (train, tick) an event with probability 1 - exp(-rate),
    duration uniform on {min..max} + 1 (ParamMalfunctionGen)."""
    import numpy as np
    from tests._util import load_golden
    fx, _ = load_golden("c1_synth18")                                     # rate 0.01, durations 5..15
    rm = backend.RailMap(fx)
    B, T = 8192, 2
    eng = EmulEngine(rm, n_envs=B, q_cap=64, ep_cap=2, tick_cap=60)
    eng.set_hparams(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0,
                    seeds=np.arange(B) * 7919 + 5, episodes=1)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100)
    tick = eng._download("trace_tick", B * 60 * T * 8).view(backend.TICK_DT).reshape(B, 60, T)
    n = eng.counters()["n_tick_logged"]
    m = tick["malf"].astype(int)
    valid = np.arange(60)[None, :, None] < n[:, None, None]
    prev = np.concatenate([np.zeros((B, 1, T), int), m[:, :-1]], axis=1)
    draws = ((prev == 0) & valid).sum()
    new = (prev == 0) & (m > 0) & valid
    p = 1.0 - np.exp(-0.01)
    assert abs(new.sum() / draws - p) < 4 * np.sqrt(p / draws)
    durs = np.bincount(m[new] + 1, minlength=17)
    assert durs[:6].sum() == 0 and len(durs) == 17                         # 5..15 + 1
    expect = new.sum() / 11
    assert (np.abs(durs[6:17] - expect) < 5 * np.sqrt(expect)).all()


def test_emul_reference_loop_with_the_drop_in_learner_methods(emul_lib):
    """distr_q.py:296-362 written against the drop-in classes only -- env.reset / agent_iter / last / step and
    model.update / max_q / max_action, actions replayed from the golden -- ends with the reference's Q-table."""
    import numpy as np
    from switchfl_b200 import api
    from tests._util import hparams, q_dict
    env, g = _aec_env("c1_synth18", emul_lib)
    hp = hparams(g)
    model = api.DistrQLearning(env=env, seed=int(g["seed"]), **hp)
    model.q_table = dict(env.rail_map.q_init_rows(hp["default_q"]))                             # __init_q_table (:299-300)
    ninter = {a: 0 for a in env.possible_agents}
    i = 0
    for ep in range(int(g["n_episodes"])):
        env.reset(seed=int(g["seed"]))
        update_dict, at_dest = {}, []
        for agent in env.agent_iter():
            obs, R, term, trunc, info = env.last()
            train = info["active_train"]
            action = int(g["dec_action"][i])
            if g["dec_greedy"][i]:                                                               # the exploit branch (:318-319)
                assert model.max_action(obs, agent, info["action_mask"]) == action
            post = env.step(action)
            nxt = "switch_%d-%d" % post["next_switch"]
            if (agent, train) in update_dict:
                pobs, pact, pagent = update_dict.pop((agent, train))
                model.update(pobs, pact, R[train], obs, pagent, agent, ninter)
            update_dict[(nxt, train)] = (obs, action, agent)
            for h in post["arrived_trains"]:
                if h not in at_dest:
                    at_dest.append(h)
                    for k in [k for k in update_dict if k[1] == h]:
                        pobs, pact, pagent = update_dict.pop(k)
                        model.update(pobs, pact, 1000.0, None, pagent, None, ninter)
            ninter[agent] += 1
            i += 1
    assert i == len(g["dec_action"])
    assert model.q_table == q_dict(g["q_keys"], g["q_vals"])
    assert model.eval(g["dec_obs"][0][g["dec_obs"][0] != -9], 0, env.possible_agents[int(g["dec_switch"][0])]) == \
        q_dict(g["q_keys"], g["q_vals"])[tuple(int(x) for x in g["dec_obs"][0] if x != -9)][0]


def check_reapply_q_init(engine_cls):
    """A second learn() call re-runs __init_q_table (distr_q.py:299-300), which ASSIGNS the optimistic rows (:156-158,
    :179-181): existing rows of init states go back to their initial values, every other row stays."""
    import numpy as np
    from tests._util import load_golden
    fx, _ = load_golden("slips24_t6")
    rm = backend.RailMap(fx)
    eng = engine_cls(rm, n_envs=3, q_cap=4096, ep_cap=8)
    eng.set_hparams(gamma=0.95, epsilon=0.5, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=1.0, default_q=[1.5, -2.0, 0.25], seeds=[4, 5, 6], episodes=5)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100000)
    eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
    before = [eng.export_q(i) for i in range(3)]
    eng.reapply_q_init()
    for i, dq in enumerate([1.5, -2.0, 0.25]):
        init = rm.q_init_rows(dq)
        after = eng.export_q(i)
        assert set(after) == set(before[i])
        changed = 0
        for k, row in before[i].items():
            want = init[k] if k in init else row
            assert after[k] == want, (i, k)
            changed += want != row
        assert changed > 0, "no learned init-state row to reset: the case is vacuous"
    eng.close()


def test_emul_reapply_q_init(emul_lib):
    check_reapply_q_init(EmulEngine)
