"""ORACLE TOOLING (test infrastructure) -- golden-vector generator.  Runs only in the build container.

Executes the reference's OWN ``switchfl`` code, unmodified, from /root/reference on top of
``oracle/shim`` (fixture-backed flatland subset = ``oracle/trainsim.py``) and records, per fixture:

  * ``ref_*``   the reference's port graph / action tables (row A0) as built by ``RailNetwork``,
  * ``dist``    the distance map of the VENDORED flatland_patch/distance_map.py (row F6) and the
                greedy shortest paths it yields,
  * ``qinit_*`` the Q-table right after ``__init_q_table`` (row Q4),
  * ``dec_*``   one record per switch-agent decision of ``DistrQLearning.learn`` (rows E2-E4, O1-O3,
                R1, Q1): episode, tick, switch, train, observation, mask, reward, action, next switch,
                arrived trains, semaphore table after the decision,
  * ``tick_*``  one record per ``rail_env.step`` (rows E5-E7, F1-F5),
  * ``malf_*``  the malfunction events (replay input),
  * ``q_*``     the final Q-table dict (rows Q2, Q3, Q6) and ``ep_*`` the per-episode metrics.

Outputs go to tests/golden/<fixture>.npz (+ the fixture itself as <fixture>.fixture.npz).  The GPU box
never runs this file: it has no /root/reference.

    python oracle/gen_golden.py [name ...]   # regenerate every golden (or the named ones)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shim"), REF]

from __graft_entry__ import load_package  # noqa: E402

load_package()
from switchfl_b200 import mapgen, railmap  # noqa: E402

from oracle import trainsim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

HPARAMS = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)


def fixtures():
    """(fixture, seed, episodes, hyper-parameters[, injected malfunctions {(tick, train): duration}])."""
    fx = mapgen.loop_chord_fixture()
    yield fx, 450565, 3, dict(HPARAMS)
    # Row F4, SURVEY Appendix B step 4 "MALF_OFF_MAP -> (done & ed) STOPPED": train 1 breaks down before departing; when its
    # counter runs out (tick 6, past its earliest departure) its entry cell is occupied by train 0, which broke down ON
    # that cell -- no valid movement, so train 1 goes STOPPED and is placed on the occupied cell.
    yield (mapgen.offmap_malfunction_fixture(), 450565, 2, dict(HPARAMS), {(3, 0): 12, (1, 1): 5})
    # C1-synthetic: test_model.py:14-63 parameters on the synthetic generator
    yield (mapgen.make_fixture(18, 2, 4, seed=450565 % 1000, num_cities=5, malfunction_rate=0.01, min_duration=5,
                               max_duration=15, name="c1_synth18"), 450565, 5, dict(HPARAMS))
    yield (mapgen.make_fixture(24, 6, 12, seed=11, num_cities=3, malfunction_rate=0.03, min_duration=2,
                               max_duration=6, name="slips24_t6", p_slip=0.7), 64, 4,
           dict(HPARAMS, lr_decay_rate=0.9999, gamma=0.95, default_q=1.5))
    yield (mapgen.make_fixture(40, 12, 30, seed=5, num_cities=4, malfunction_rate=0.0, name="synth40_t12",
                               p_slip=0.3), 65, 2, dict(HPARAMS))
    # C3 / C4: BASELINE.json configs[2] / configs[3] on the double-track generator -- the maps bench.py times
    yield (mapgen.c3_fixture(64), 64, 3, dict(HPARAMS))
    yield (mapgen.c4_fixture(), 7, 3, dict(HPARAMS))
    # Congested single-track maps of the same classes (edge cases: gridlock, forced stops, the reference's crash site)
    yield (mapgen.make_fixture(n=80, n_trains=15, n_chords=50, seed=64, num_cities=25, name="c3_synth80_s64", p_slip=0.3),
           64, 2, dict(HPARAMS))
    # C4-class: BASELINE.json configs[3] -- 100x100, 50 trains (T > 32: both mask words), 281 switches, malfunctions
    yield (mapgen.make_fixture(100, 50, 120, seed=0, num_cities=25, malfunction_rate=0.01, min_duration=5,
                               max_duration=15, name="c4_synth100_t50", p_slip=0.3), 7, 1, dict(HPARAMS))


def _load_vendored_distance_map():
    spec = importlib.util.spec_from_file_location("_ref_distance_map", os.path.join(REF, "flatland_patch", "distance_map.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_fixture(fx: dict, seed: int, n_episodes: int, hp: dict, inject=None, verbose=True, rail_env=None):
    """``rail_env``: a ready RailEnv to run the reference on (tools/record_flatland_fixture.py passes a GENUINE flatland
    one); default: the fixture-backed stand-in of oracle/trainsim.py."""
    import logging
    logging.disable(logging.INFO)
    from switchfl.distr_q import DistrQLearning
    from switchfl.switch_env import ASyncSwitchEnv
    from switchfl.utils.naming import name2switch_id

    if rail_env is None:
        rail_env = trainsim.RailEnv(fx)
    if inject is not None:
        rail_env.injected_malfunctions = dict(inject)         # the same schedule in every episode (replay input)
    env = ASyncSwitchEnv(rail_env, render_mode=None, max_steps=100_000)
    rn = env.rail_network
    out = {}

    # ------------------------------------------------------------------ A0: the reference's own tables
    tab = railmap.build_switch_tables(fx["grid"])
    names = rn.get_switch_names()
    assert names == tab.switch_names(), (names, tab.switch_names())
    pid = {p: i for i, p in enumerate(tab.port_ids)}
    W = tab.W
    ref_ports, ref_nbr, ref_dist, ref_prev, ref_nintra, ref_intra0 = [], [], [], [], [], []
    ref_act, ref_sw = [], []
    ref_rn, ref_rn_off = [], [0]
    for si, name in enumerate(names):
        sw = rn.get_switch_on_position(name2switch_id(name))
        ports = sw.get_port_nodes()
        ref_sw.append((len(ports), env.action_space(name).n))
        for p in ports:
            ref_ports.append((si, round((p[0] - int(p[0])) * 10), rn.map_direction(p)))
            nsw, nport = sw.port2neighbor[p]
            ref_nbr.append(pid[nport])
            ref_dist.append(rn.get_port_distance(p, nport))
            ref_rn.extend(rn.rail_graph.get_edge_data(p, nport)["rail_nodes"])
            ref_rn_off.append(len(ref_rn))
            prev = rn.rail_graph.nodes.data("rail_prev_node")[p]
            ref_prev.append(prev[0] * W + prev[1])
            intra = [e[1] for e in rn.rail_graph.edges(p) if (int(e[1][0]), int(e[1][1])) == sw.id and e[1] != nport]
            ref_nintra.append(len(intra))
            ref_intra0.append(pid[intra[0]] if intra else -1)
            assert pid[p] == len(ref_nbr) - 1, "global port order differs"
        for a, (pin, pout) in enumerate(sw.action_outcomes):
            ref_act.append((si, ports.index(pin), ports.index(pout), int(sw.actions[a][pin][1])))
    out["ref_switch"] = np.array(ref_sw, np.int32)
    out["ref_ports"] = np.array(ref_ports, np.int32)
    out["ref_port_nbr"] = np.array(ref_nbr, np.int32)
    out["ref_port_dist"] = np.array(ref_dist, np.int32)
    out["ref_port_prev"] = np.array(ref_prev, np.int32)
    out["ref_port_nintra"] = np.array(ref_nintra, np.int32)
    out["ref_port_intra0"] = np.array(ref_intra0, np.int32)
    out["ref_actions"] = np.array(ref_act, np.int32)
    out["ref_rail_nodes"] = np.array(ref_rn, np.int32).reshape(-1, 2)
    out["ref_rail_nodes_off"] = np.array(ref_rn_off, np.int32)

    # ------------------------------------------------------------------ F6 against the vendored patch
    vend = _load_vendored_distance_map()
    rail_env.reset(random_seed=seed)
    vdm = vend.DistanceMap(rail_env.agents, rail_env.height, rail_env.width)
    vdm.reset(rail_env.agents, rail_env.rail)
    vdist = vdm.get(rail_env.agents)
    own = rail_env.distance_map.get(rail_env.agents)
    assert np.array_equal(vdist, own), "oracle BFS differs from vendored flatland_patch/distance_map.py"
    vpaths = vdm.get_shortest_paths(agents=rail_env.agents)
    opaths = rail_env.distance_map.get_shortest_paths(agents=rail_env.agents)
    for h in vpaths:
        assert [(tuple(w.position), int(w.direction)) for w in vpaths[h]] == \
               [(tuple(w.position), int(w.direction)) for w in opaths[h]], "shortest path differs"
    dist_i = np.where(np.isinf(vdist), railmap.INF_DIST, vdist).astype(np.int32)
    out["dist"] = dist_i
    out["path_len"] = np.array([len(vpaths[h]) for h in sorted(vpaths)], np.int32)

    # ------------------------------------------------------------------ learn() with tracing
    model = DistrQLearning(env=env, seed=seed, **hp)
    T = rail_env.get_num_agents()
    NP = tab.NP
    dec = {k: [] for k in ("ep", "tick", "switch", "train", "obs", "mask", "reward", "action", "next_switch",
                           "arrived", "sem", "done", "greedy")}
    tick = {k: [] for k in ("ep", "tick", "pos", "dir", "state", "malf")}
    malf = []
    state = {"ep": -1, "qinit": None}

    orig_reset, orig_last, orig_step = env.reset, env.last, env.step
    orig_max_action = model.max_action

    def traced_max_action(*a, **k):          # distr_q.py:318-319: the exploit branch (it also inserts the row, :482)
        pending["greedy"] = 1
        return orig_max_action(*a, **k)
    model.max_action = traced_max_action
    sw_index = {name2switch_id(n): i for i, n in enumerate(names)}

    def traced_reset(seed=None, options=None):
        state["ep"] += 1
        hook_railenv()
        return orig_reset(seed=seed, options=options)

    def hook_railenv():
        if getattr(rail_env, "_traced", False):
            return
        rs = rail_env.step

        def traced_rail_step(actions):
            r = rs(actions)
            tick["ep"].append(state["ep"]); tick["tick"].append(rail_env._elapsed_steps)
            tick["pos"].append([(-1 if a.position is None else a.position[0] * W + a.position[1]) for a in rail_env.agents])
            tick["dir"].append([int(a.direction) for a in rail_env.agents])
            tick["state"].append([int(a.state) for a in rail_env.agents])
            tick["malf"].append([a.malfunction_handler.malfunction_down_counter for a in rail_env.agents])
            return r
        rail_env.step = traced_rail_step
        rail_env._traced = True

    pending = {}

    def traced_last(observe=True):
        if state["qinit"] is None and state["ep"] == 0 and model.q_table:
            state["qinit"] = {k: list(v) for k, v in model.q_table.items()}
        r = orig_last(observe)
        obs, rew, term, trunc, info = r
        pending.update(obs=np.array(obs), reward=float(rew[env.active_train]), mask=np.array(info["action_mask"]),
                       switch=sw_index[name2switch_id(env.agent_selection)], train=int(env.active_train),
                       tick=rail_env._elapsed_steps, done=bool(term or trunc), greedy=0)
        if term or trunc:   # learn() breaks without stepping
            pass
        return r

    def traced_step(action):
        post = orig_step(action)
        o = np.full(18, -9, np.int64); o[:len(pending["obs"])] = pending["obs"]
        m = np.full(9, -1, np.int8); m[:len(pending["mask"])] = pending["mask"]
        dec["ep"].append(state["ep"]); dec["tick"].append(pending["tick"]); dec["switch"].append(pending["switch"])
        dec["train"].append(pending["train"]); dec["obs"].append(o); dec["mask"].append(m)
        dec["reward"].append(pending["reward"]); dec["action"].append(int(action)); dec["greedy"].append(pending["greedy"])
        dec["next_switch"].append(sw_index[tuple(post["next_switch"])])
        dec["arrived"].append(sum(1 << int(h) for h in post["arrived_trains"]))
        dec["done"].append(int(env.terminated) | (int(env.truncated) << 1))
        sem = np.full((NP, 4), -1, np.int32)
        for p, (tr, typ, d, t0, t1) in rn.semaphores.items():
            assert d == rn.map_direction(p), "semaphore dir invariant (SURVEY 8a E3) violated"
            sem[pid[p]] = (tr, 0 if typ == "in" else 1, t0, t1)
        dec["sem"].append(sem)
        return post

    env.reset, env.last, env.step = traced_reset, traced_last, traced_step

    with tempfile.TemporaryDirectory() as tmp:
        model.learn(num_episodes=n_episodes, out_dir=tmp, checkpoint_freq=10 ** 9)
        ep_cum = np.load(os.path.join(tmp, "cum_reward.npz"))["x"]
        ep_arr = np.load(os.path.join(tmp, "arrived_trains.npz"))["x"]
        ep_del = np.load(os.path.join(tmp, "delays.npz"))["x"]
        ep_mal = np.load(os.path.join(tmp, "num_malfunctions.npz"))["x"]

    # malfunction events per episode are re-derived from the tick trace (counter jumps up)
    tk_ep, tk_t, tk_m = np.array(tick["ep"]), np.array(tick["tick"]), np.array(tick["malf"]).reshape(-1, T)
    prev_ep, prev = -1, None
    for i in range(len(tk_ep)):
        if tk_ep[i] != prev_ep:
            prev = np.zeros(T, np.int64); prev_ep = tk_ep[i]
        for h in range(T):
            # counter after the tick's decrement: an event of duration d shows as d-1 following 0
            if prev[h] == 0 and tk_m[i, h] > 0:
                malf.append((tk_ep[i], tk_t[i], h, tk_m[i, h] + 1))
        prev = tk_m[i]

    def pack_q(q):
        keys = np.full((len(q), 18), -9, np.int64)
        vals = np.full((len(q), 9), np.nan)
        for i, (k, v) in enumerate(q.items()):
            keys[i, :len(k)] = np.array(k, np.int64); vals[i, :len(v)] = v
        return keys, vals

    for k in ("ep", "tick", "switch", "train", "action", "next_switch", "done", "greedy"):
        out["dec_" + k] = np.array(dec[k], np.int32)
    out["dec_arrived"] = np.array(dec["arrived"], np.uint64)
    out["dec_obs"] = np.array(dec["obs"], np.int64).reshape(-1, 18)
    out["dec_mask"] = np.array(dec["mask"], np.int8).reshape(-1, 9)
    out["dec_reward"] = np.array(dec["reward"], np.float64)
    out["dec_sem"] = np.array(dec["sem"], np.int32).reshape(-1, NP, 4)
    for k in ("ep", "tick"):
        out["tick_" + k] = np.array(tick[k], np.int32)
    for k in ("pos", "dir", "state", "malf"):
        out["tick_" + k] = np.array(tick[k], np.int32).reshape(-1, T)
    out["malf_events"] = np.array(malf, np.int32).reshape(-1, 4)
    out["qinit_keys"], out["qinit_vals"] = pack_q(state["qinit"] or {})
    out["q_keys"], out["q_vals"] = pack_q(model.q_table)
    out["ep_cum_reward"] = np.asarray(ep_cum, np.float64)
    out["ep_arrived"] = np.asarray(ep_arr, np.int32)
    out["ep_delays"] = np.asarray(ep_del, np.float64).reshape(n_episodes, T)
    out["ep_num_malfunctions"] = np.asarray(ep_mal, np.int32)
    out["seed"] = np.int64(seed)
    if inject is not None:                                    # free-running consumers must inject the same schedule
        out["inject_events"] = np.array([(t, h, d) for (t, h), d in sorted(inject.items())], np.int32).reshape(-1, 3)
    out["n_episodes"] = np.int32(n_episodes)
    for k, v in hp.items():
        out["hp_" + k] = np.float64(v)
    if verbose:
        print(f"[{fx['name']}] S={tab.S} NP={NP} T={T} decisions={len(dec['ep'])} ticks={len(tick['ep'])} "
              f"malf_events={len(malf)} q_rows={len(model.q_table)} arrived/ep={list(ep_arr)} cum={list(np.round(ep_cum, 1))}")
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])
    for fx, seed, n_ep, hp, *rest in fixtures():
        if only and fx["name"] not in only:
            continue
        mapgen.check_fixture(fx)
        out = run_fixture(fx, seed, n_ep, hp, inject=rest[0] if rest else None)
        if fx["name"] == "f4_offmap_7x7":                     # the transition under test is really taken
            st = out["tick_state"][:, 1]
            assert any(a == 2 and b == 4 for a, b in zip(st[:-1], st[1:])), "MALFUNCTION_OFF_MAP -> STOPPED not taken"
            k = int(np.nonzero((st[:-1] == 2) & (st[1:] == 4))[0][0]) + 1
            assert out["tick_pos"][k, 0] == out["tick_pos"][k, 1] >= 0, "train 1 was not placed on the occupied entry cell"
        mapgen.save_fixture(os.path.join(GOLDEN_DIR, fx["name"] + ".fixture.npz"), fx)
        np.savez_compressed(os.path.join(GOLDEN_DIR, fx["name"] + ".npz"), **out)


if __name__ == "__main__":
    main()
