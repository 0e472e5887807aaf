"""CPU: the kernel LOGIC (host build of the device sources, tests/emul) against the reference goldens."""
import os
import subprocess

import pytest

from switchfl_b200 import backend
from tests._parity import check_replay
from tests._util import golden_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "network-distributed-q-learning_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul", "libsfl_emul.so")


def build_emul():
    srcs = [os.path.join(SRC, f) for f in ("sfl_api.cu", "sfl_core.cuh")] + [os.path.join(ROOT, "include", "switchfl_b200.h")]
    if not os.path.exists(EMUL) or any(os.path.getmtime(s) > os.path.getmtime(EMUL) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DSFL_HOST_EMUL", "-x", "c++",
                               "-I", os.path.join(ROOT, "include"), "-I", SRC, "-o", EMUL, os.path.join(SRC, "sfl_api.cu")])
    return EMUL


@pytest.fixture(scope="session")
def emul_lib():
    return build_emul()


@pytest.mark.parametrize("name", golden_names())
def test_emul_replay_matches_reference(name, emul_lib):
    check_replay(name, lambda rm, **kw: backend.Engine(rm, _emul_lib=emul_lib, **kw), n_envs=2)


def test_emul_chunked_launches_equal_one_launch(emul_lib):
    check_replay("slips24_t6", lambda rm, **kw: backend.Engine(rm, _emul_lib=emul_lib, **kw), n_envs=1, chunk=7)


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
    eng = backend.Engine(rm, n_envs=B, q_cap=16384, dec_cap=60000, tick_cap=5000, ep_cap=8, _emul_lib=emul_lib)
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
