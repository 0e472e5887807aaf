"""GPU (-m gpu): the CUDA path through the C-ABI against the reference goldens and against the oracle."""
import numpy as np
import pytest

from switchfl_b200 import backend, mapgen
from tests._parity import check_replay
from tests._util import golden_names, load_golden

pytestmark = pytest.mark.gpu


def gpu_engine(rm, **kw):
    return backend.Engine(rm, device="cuda:0", **kw)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_replay_matches_reference(name):
    check_replay(name, gpu_engine, n_envs=3)


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
@pytest.mark.parametrize("name", ["slips24_t6", "synth40_t12"])
def test_cuda_replay_any_lane_group_width(name, lanes):
    """The lanes-per-environment choice is scheduling only: every width reproduces the reference trace."""
    check_replay(name, lambda rm, **kw: gpu_engine(rm, lanes=lanes, **kw), n_envs=5)


def test_cuda_chunked_launches_equal_one_launch():
    check_replay("slips24_t6", gpu_engine, n_envs=2, chunk=5)


def test_cuda_replay_many_seeds_against_oracle():
    """64 environments with different seeds: the oracle free-runs, the CUDA path replays each stream."""
    from oracle.switchfl_oracle import SwitchFLOracle
    fx, g = load_golden("c1_synth18")
    rm = backend.RailMap(fx)
    hp = dict(gamma=0.9, epsilon=0.6, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=1.0, default_q=0.5)
    B, n_ep = 64, 3
    oracles, acts, evs = [], [], []
    for i in range(B):
        o = SwitchFLOracle(fx, rm.tab, seed=1000 + i, **hp)
        o.enable_trace()
        o.learn(n_ep)
        oracles.append(o)
        acts.append(o.replay_stream())
        evs.append(o.malfunction_schedule())
    n_dec = max(len(a) for a in acts)
    eng = gpu_engine(rm, n_envs=B, q_cap=1024, dec_cap=n_dec + 4, act_cap=n_dec + 4, ev_cap=max(len(e) for e in evs) + 2, ep_cap=n_ep + 1)
    eng.set_hparams(**hp, seeds=np.arange(B), episodes=n_ep)
    eng.set_replay(acts, evs)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_REPLAY, 100000)
    eng.check_errors()
    for i, o in enumerate(oracles):
        dec, _, _ = eng.trace(i)
        assert np.array_equal(dec["reward"].astype(np.float64), np.array(o.trace["dec_reward"])), i
        assert np.array_equal(dec["sw"], np.array(o.trace["dec_switch"])), i
        assert eng.export_q(i, include_init=True) == o.q_table, i
    eng.close()


def test_cuda_learn_4096_envs_properties():
    """BASELINE config C2 size: free-running learn on 4096 envs; size-independent invariants."""
    fx, _ = load_golden("c1_synth18")
    rm = backend.RailMap(fx)
    B = 4096
    eng = gpu_engine(rm, n_envs=B, q_cap=1024, ep_cap=8)
    eng.set_hparams(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0,
                    seeds=np.arange(B) + 450565, episodes=4)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100000)
    eng.check_errors()
    c = eng.counters()
    assert (c["halted"] == 1).all() and (c["episodes"] == 4).all()
    n, log, delays = eng.episode_log()
    assert (n == 4).all()
    assert (log["decisions"][:, :4].sum(axis=1) == c["decisions"]).all()           # bookkeeping closes
    assert (log["ticks"][:, :4].sum(axis=1) == c["ticks"]).all()
    assert (log["ticks"][:, :4] <= int(fx["max_episode_steps"])).all()
    assert (log["arrived"][:, :4] >= 0).all() and (log["arrived"][:, :4] <= 2).all()
    d, t = eng.total_decisions()
    assert d == int(c["decisions"].sum()) and t == int(c["ticks"].sum())
    # same seed -> same trajectory (determinism across launches)
    eng.reset()
    eng.run(backend.MODE_LEARN, 100000)
    c2 = eng.counters()
    assert np.array_equal(c2["decisions"], c["decisions"]) and np.array_equal(c2["ticks"], c["ticks"])
    # different seeds explore differently
    assert len(np.unique(c["decisions"])) > 1
    q = eng.export_q(0)
    assert all(np.isfinite(v).all() for v in q.values())
    eng.close()


def test_replay_underrun_is_reported_loudly():
    fx, _ = load_golden("loop_chord_7x7")
    rm = backend.RailMap(fx)
    eng = gpu_engine(rm, n_envs=2, act_cap=1)
    eng.set_hparams(episodes=1)
    eng.set_replay([[4], [4]])
    eng.reset()
    eng.run(backend.MODE_REPLAY, 100)
    with pytest.raises(RuntimeError, match="replay action stream exhausted"):
        eng.check_errors()
    eng.close()


@pytest.mark.parametrize("name", ["c1_synth18", "slips24_t6"])
def test_cuda_aec_protocol_matches_reference(name):
    """reset / agent_iter / last / step through the CUDA path (SFL_MODE_STEP), driven like distr_q.py:296-362."""
    from switchfl_b200 import api
    from tests._parity import golden_events
    from tests.test_emul_parity import check_aec
    fx, g = load_golden(name)
    ev = golden_events(g)
    env = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=1, device="cuda:0", q_cap=64, ep_cap=2,
                             _engine_kwargs={"ev_cap": len(ev) + 2})
    env.engine.set_replay(None, [ev])
    check_aec(env, g)
    env.engine.close()


def test_cuda_step_batch_lockstep_equals_single_env():
    """Five environments stepped in lockstep with a random masked policy each: env i behaves exactly like a
    single-environment run with the same seed and actions."""
    from switchfl_b200 import api
    fx, _ = load_golden("slips24_t6")
    B = 5
    env = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=B, device="cuda:0", q_cap=64, ep_cap=2)
    env.reset(seed=100)
    rng = np.random.default_rng(0)
    logs = [[] for _ in range(B)]
    for _ in range(4000):
        rec = env.last_batch()
        if not rec["pending"].any():
            break
        acts = np.full(B, -1)
        for i in range(B):
            if rec["pending"][i]:
                allowed = [a for a in range(16) if (int(rec["mask"][i]) >> a) & 1]
                acts[i] = rng.choice(allowed)
                logs[i].append((int(rec["sw"][i]), int(rec["train"][i]), int(rec["key"][i]), int(rec["mask"][i]), int(acts[i]),
                                int(rec["rewards"][i][rec["train"][i]])))
        env.step_batch(acts)
    assert all(len(l) > 10 for l in logs)
    for i in (0, 3):
        one = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=1, device="cuda:0", q_cap=64, ep_cap=2)
        one.reset(seed=100 + i)
        k = 0
        for agent in one.agent_iter():
            obs, R, term, trunc, info = one.last()
            sw, train, key, mask, action, reward = logs[i][k]
            assert agent == one.possible_agents[sw] and info["active_train"] == train and R[train] == reward
            assert tuple(int(x) for x in obs) == one.rail_map.key_to_obs(key)
            one.step(action)
            k += 1
        assert k == len(logs[i])
        one.engine.close()
    env.engine.close()


# ---------------------------------------------------------------------------------------------- edge cases vs the free-running oracle
def test_cuda_truncation_by_max_steps():
    from tests._parity import check_against_oracle
    from tests.test_emul_parity import HP_EDGE
    fx, _ = load_golden("slips24_t6")
    check_against_oracle(gpu_engine, fx, HP_EDGE, 3, [5, 6, 7], max_steps=25)


def test_cuda_greedy_rollout_after_training():
    from tests._parity import check_against_oracle
    from tests.test_emul_parity import HP_EDGE
    fx, _ = load_golden("slips24_t6")
    check_against_oracle(gpu_engine, fx, HP_EDGE, 2, [11, 12, 13], greedy_after=True)


@pytest.mark.parametrize("kind", ["one_train", "max_trains"])
def test_cuda_train_count_extremes(kind):
    from tests._parity import check_against_oracle
    from tests.test_emul_parity import HP_EDGE, edge_fixture
    check_against_oracle(gpu_engine, edge_fixture(kind), HP_EDGE, 1, [21, 22, 23], q_cap=65536 if kind == "max_trains" else 1024)


def test_cuda_replay_detects_a_wrong_greedy_action():
    """An action recorded as an exploit choice must be the argmax of the row on the device, else the env is flagged."""
    from tests._parity import make_replay_engine, replay_stream
    fx, g, rm, eng = make_replay_engine("c1_synth18", gpu_engine, n_envs=2)
    stream = replay_stream(g).copy()
    i = int(np.nonzero(g["dec_greedy"])[0][3])
    allowed = [a for a in range(9) if a < len([x for x in g["dec_mask"][i] if x >= 0]) and g["dec_mask"][i][a] == 1 and a != g["dec_action"][i]]
    if allowed:
        stream[i] = np.int8(allowed[0] | 0x40)
        eng.set_replay([replay_stream(g), stream], None)
        eng.run(backend.MODE_REPLAY, 100000)
        err = eng.counters()["err"]
        assert err[0] == 0 and err[1] & 128
    eng.close()


def test_q_table_full_is_reported_loudly():
    fx, _ = load_golden("slips24_t6")
    rm = backend.RailMap(fx)
    eng = gpu_engine(rm, n_envs=4, q_cap=8)
    eng.set_hparams(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0, episodes=2)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100000)
    with pytest.raises(RuntimeError, match="Q table full"):
        eng.check_errors()
    eng.close()


@pytest.mark.parametrize("i", [0, 3, 5, 7, 8, 11])
def test_cuda_fuzz_maps_against_oracle(i):
    """Random maps / timetables / hyper-parameters (the CPU suite runs the same cases on the host build)."""
    from tests._parity import check_against_oracle
    from tests.test_emul_parity import fuzz_case
    try:
        fx, hp, seeds, max_steps = fuzz_case(i)
    except ValueError as ex:
        pytest.skip(str(ex))
    check_against_oracle(gpu_engine, fx, hp, 3, seeds, max_steps=max_steps, greedy_after=True, q_cap=16384)


@pytest.mark.parametrize("name", ["c4_rail100_t50", "c4_synth100_t50"])
def test_cuda_learn_c4_size_properties(name):
    """BASELINE config C4 size (100x100, 50 trains, 8192 envs per GPU): free-running learn; size-independent invariants --
    on the benchmark map (double track: no episode is abandoned) and on the congested single-track map of round 1, where the
    reference itself can die in observer.py:294-307; those episodes are abandoned and counted."""
    fx, _ = load_golden(name)
    rm = backend.RailMap(fx)
    B, n_ep = 8192, 2
    eng = gpu_engine(rm, n_envs=B, q_cap=16384, ep_cap=4)
    hp = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)
    eng.set_hparams(**hp, seeds=np.arange(B) + 450565, episodes=n_ep)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 100000)
    eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
    c = eng.counters()
    assert (c["halted"] == 1).all() and (c["episodes"] == n_ep).all()
    assert ((c["aborted"] > 0) == (c["err"] != 0)).all() and (c["aborted"] <= n_ep).all()
    if name == "c4_rail100_t50":
        assert (c["aborted"] == 0).all() and (c["err"] == 0).all()
        assert c["forced_stops"].sum() < 0.2 * c["decisions"].sum() and c["arrived_trains"].sum() > 0.15 * 50 * n_ep * B
    n, log, delays = eng.episode_log()
    assert (n == n_ep).all()
    assert (log["decisions"][:, :n_ep].sum(axis=1) == c["decisions"]).all()          # bookkeeping closes
    assert (log["ticks"][:, :n_ep].sum(axis=1) == c["ticks"]).all()
    assert (log["ticks"][:, :n_ep] <= int(fx["max_episode_steps"])).all()
    full = (c["aborted"] == 0)
    assert (log["ticks"][full][:, :n_ep] == int(fx["max_episode_steps"])).all() | (log["arrived"][full][:, :n_ep] == 50).all()
    pop = np.array([[bin(int(m)).count("1") for m in row[:n_ep]] for row in log["arrived_mask"]])
    assert (pop == log["arrived"][:, :n_ep]).all()
    d, t = eng.total_decisions()
    assert d == int(c["decisions"].sum()) and t == int(c["ticks"].sum())
    first = c.copy()
    eng.reset()                                                                     # same seeds -> the same trajectories
    eng.run(backend.MODE_LEARN, 100000)
    c2 = eng.counters()
    assert np.array_equal(c2["decisions"], first["decisions"]) and np.array_equal(c2["aborted"], first["aborted"])
    eng.close()


def test_cuda_shared_table_mode_matches_the_host_build_properties():
    """Shared-table mode (extension): atomics on the device, integer accumulation -> bit-reproducible and equal to a
    single-environment run when all environments are identical."""
    fx, _ = load_golden("slips24_t6")
    rm = backend.RailMap(fx)
    hp = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)

    def run(seeds):
        eng = gpu_engine(rm, n_envs=len(seeds), q_cap=2, ep_cap=4, shared_q=True)
        eng.set_hparams(**hp, seeds=seeds, episodes=-1)
        eng.reset()
        eng.init_shared_q(0.0)
        for _ in range(5):
            eng.run(backend.MODE_LEARN, 64)
            eng.check_errors()
            eng.shared_q_sync()
        q = eng.shared_q_table()
        eng.close()
        return q
    q1 = run([7])
    assert np.array_equal(q1, run([7] * 64))
    qa, qb = run(list(range(1, 200))), run(list(range(1, 200)))
    assert np.array_equal(qa, qb) and np.isfinite(qa).all() and not np.array_equal(qa, q1)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_distance_map_matches_vendored_reference(name):
    """Row F6 / N1: the k_distance_map kernel against the vendored flatland_patch/distance_map.py (golden ``dist``)."""
    fx, g = load_golden(name)
    rm = backend.RailMap(fx)
    d = backend.device_distance_map(fx["grid"], rm.trains.targets)
    assert np.array_equal(d, rm.trains.dist)
    assert np.array_equal(d[rm.trains.tgt_index], g["dist"])


def test_cuda_distance_map_large_grid_uses_the_global_memory_variant():
    """A 240x240 grid (921 KB of states per target) does not fit shared memory."""
    from switchfl_b200 import railmap
    fx = mapgen.make_fixture(240, 6, 60, seed=9, num_cities=12)
    cells = [int(r) * 240 + int(c) for r, c in fx["target"]]
    d = backend.device_distance_map(fx["grid"], cells)
    for k, (r, c) in enumerate(fx["target"]):
        assert np.array_equal(d[k], railmap.distance_to(fx["grid"], (int(r), int(c)))), k


# (fixture, lanes, roomy): the small-batch configurations whose kernel instantiations are pinned below.  The last rows are
# the instantiations bench.py launches at full size (C2: 4096 envs -> G=16 / ONE / TH; C3: 4096 envs per map -> G=16,
# ROOMY; C4: 8192 envs -> G=32, not roomy); test_cuda_bench_kernel_variants_are_parity_tested checks that they really are.
VARIANT_CASES = [("c1_synth18", None, None), ("c1_synth18", 1, None), ("slips24_t6", None, None), ("slips24_t6", 4, None),
                 ("synth40_t12", 8, None), ("c4_synth100_t50", None, None),
                 ("c1_synth18", 16, None), ("c3_rail80_s64", 16, True), ("c3_rail80_s64", 16, False),
                 ("c4_rail100_t50", 32, False), ("c4_rail100_t50", 32, True)]
HP_VARIANT = dict(gamma=0.95, epsilon=0.5, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=0.9999, default_q=1.0)


def _variant_run(rm, cls, name, lanes, roomy, full, B=6, n_ep=3, **extra):
    """Free-running learn (n_ep episodes) then one greedy rollout; returns everything observable + the kernel names."""
    eng = cls(rm, n_envs=B, q_cap=32768 if "c4" in name else 4096, ep_cap=8, lanes=lanes, roomy=roomy,
              dec_cap=4 if full else 0, tick_cap=0, **extra)
    kernels = (eng.kernel_variant(backend.MODE_LEARN, traced=full), eng.kernel_variant(backend.MODE_GREEDY, traced=full))
    eng.set_hparams(**HP_VARIANT, seeds=np.arange(B) + 77, episodes=n_ep)
    eng.reset()
    eng.enable_q_init(True)
    eng.run(backend.MODE_LEARN, 1_000_000)
    eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
    c1 = eng.counters().copy()
    _, log1, d1 = eng.episode_log()
    eng.set_hparams(**HP_VARIANT, seeds=np.arange(B) + 77, episodes=1, episode_base=n_ep)
    eng.reset(keep_q=True, keep_interactions=True)
    eng.run(backend.MODE_GREEDY, 1_000_000)
    eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
    c2 = eng.counters().copy()
    _, log2, d2 = eng.episode_log()
    q = [eng.export_q(i) for i in range(B)]
    eng.close()
    return (c1, log1[:, :n_ep].copy(), d1[:, :n_ep].copy(), c2, log2[:, :1].copy(), d2[:, :1].copy(), q), kernels


def _assert_same_run(a, b):
    for k in ("decisions", "ticks", "train_ticks", "episodes", "err", "q_rows", "aborted"):
        assert np.array_equal(a[0][k], b[0][k]) and np.array_equal(a[3][k], b[3][k]), k
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    assert a[6] == b[6]


@pytest.mark.parametrize("name,lanes,roomy", VARIANT_CASES)
def test_cuda_production_kernels_equal_the_full_kernel_and_the_host_build(name, lanes, roomy):
    """The learn / greedy kernels are compile-time specialisations (no trace, no replay, single-pass train loops when
    T <= lanes, the 4-CTA "roomy" build).  Free-running learn and a greedy rollout through them must give exactly what
    (a) the full kernel -- the one the replay tests pin against the reference -- and (b) the same sources compiled for the
    host (tests/emul, which the CPU suite pins against the reference too) give for the same seeds: counters, episode
    logs, Q-tables."""
    from tests.emulated import EmulEngine
    fx, _ = load_golden(name)
    rm = backend.RailMap(fx)
    prod, k_prod = _variant_run(rm, backend.Engine, name, lanes, roomy, full=False)
    full, k_full = _variant_run(rm, backend.Engine, name, lanes, roomy, full=True)
    assert "KIND=learn" in k_prod[0] and "KIND=greedy" in k_prod[1] and "KIND=full" in k_full[0], (k_prod, k_full)
    if roomy is not None and lanes is not None and lanes >= 16 and "TH=0" in k_prod[0]:
        assert f"ROOMY={int(roomy)}" in k_prod[0], k_prod
    _assert_same_run(prod, full)
    host, _ = _variant_run(rm, EmulEngine, name, None, None, full=False)
    _assert_same_run(prod, host)


def _bench_variants():
    """Kernel instantiation of every bench.py workload at its full batch size (context only, no buffers)."""
    import bench
    out = {}
    for wl, (fixture, envs, q_cap, _) in bench.WORKLOADS.items():
        fxs = bench.workload_fixtures(bench.fixture_path(fixture))
        per_map = envs // len(fxs)
        eng = backend.Engine(backend.RailMap(fxs[0]), n_envs=per_map, q_cap=q_cap, ep_cap=4, shared_q=(wl == "c5"), bind=False,
                             **bench.engine_kwargs(len(fxs)))
        out[wl] = eng.kernel_variant(backend.MODE_LEARN)
        eng.close()
    return out


def test_cuda_bench_kernel_variants_are_parity_tested():
    """Every k_run instantiation bench.py times is one of the instantiations the parity tests above run (same template
    arguments; grid / block size are launch parameters, not code)."""
    tested = set()
    for name, lanes, roomy in VARIANT_CASES:
        fx, _ = load_golden(name)
        eng = backend.Engine(backend.RailMap(fx), n_envs=6, q_cap=4096, ep_cap=8, lanes=lanes, roomy=roomy, bind=False)
        tested.add(eng.kernel_variant(backend.MODE_LEARN))
        eng.close()
    fx, _ = load_golden("c4_rail100_t50")
    sq = backend.Engine(backend.RailMap(fx), n_envs=6, q_cap=2, shared_q=True, bind=False)        # pinned by the shared-table test below
    tested.add(sq.kernel_variant(backend.MODE_LEARN))
    sq.close()
    for wl, variant in _bench_variants().items():
        assert variant in tested, (wl, variant, sorted(tested))


def test_cuda_shared_table_kernel_equals_the_host_build_on_the_c4_map():
    """The SQ instantiation bench.py --workload c5 launches (large map, tail in HBM): integer accumulation makes the
    device run bit-reproducible, so it must equal the host build of the same sources exactly."""
    from tests.emulated import EmulEngine
    fx, _ = load_golden("c4_rail100_t50")
    rm = backend.RailMap(fx)
    hp = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)
    out = []
    for cls in (backend.Engine, EmulEngine):
        eng = cls(rm, n_envs=5, q_cap=2, ep_cap=4, shared_q=True)
        eng.set_hparams(**hp, seeds=np.arange(5) + 3, episodes=-1)
        eng.reset()
        eng.init_shared_q(0.0)
        for _ in range(4):
            eng.run(backend.MODE_LEARN, 128)
            eng.check_errors()
            eng.shared_q_sync()
        out.append((eng.shared_q_table(), eng.counters()["decisions"].copy()))
        eng.close()
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])


@pytest.mark.parametrize("name", ["c1_synth18", "slips24_t6"])
def test_cuda_free_running_learn_equals_the_host_build(name):
    """Free-running learn (Philox epsilon-greedy and malfunction draws included) on the device against the same sources
    compiled for the host (tests/emul): identical counters, episode logs and Q-tables for identical seeds."""
    from tests.emulated import EmulEngine
    fx, _ = load_golden(name)
    rm = backend.RailMap(fx)
    hp = dict(gamma=0.95, epsilon=0.5, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=1.0, default_q=0.0)
    B, n_ep = 12, 4
    out = []
    for cls in (backend.Engine, EmulEngine):
        eng = cls(rm, n_envs=B, q_cap=4096, ep_cap=8)
        eng.set_hparams(**hp, seeds=np.arange(B) + 31, episodes=n_ep)
        eng.reset()
        eng.enable_q_init(True)
        for _ in range(40):                                           # several launches: the streams do not depend on the cut
            eng.run(backend.MODE_LEARN, 97)
        eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
        c = eng.counters().copy()
        _, log, dl = eng.episode_log()
        out.append((c, log[:, :n_ep].copy(), dl[:, :n_ep].copy(), [eng.export_q(i) for i in range(B)]))
        eng.close()
    (c1, l1, d1, q1), (c2, l2, d2, q2) = out
    assert (c1["halted"] == 1).all()
    for k in ("decisions", "ticks", "train_ticks", "episodes", "err", "q_rows", "aborted"):
        assert np.array_equal(c1[k], c2[k]), k
    assert np.array_equal(l1, l2) and np.array_equal(d1, d2) and q1 == q2


def test_cuda_reapply_q_init():
    from tests.test_emul_parity import check_reapply_q_init
    check_reapply_q_init(backend.Engine)


def test_cuda_shared_table_overlapped_schedule_equals_the_serial_host_run():
    """run_shared on the device -- all-reduce + apply on a second stream while the next launch runs -- gives exactly the
    table the host build gives running the same schedule serially: the overlap changes timing, not results."""
    from tests.emulated import EmulEngine
    from tests.test_host_api import _shared_run
    fx, _ = load_golden("slips24_t6")
    rm = backend.RailMap(fx)
    seeds = list(range(1, 40))
    dev = _shared_run(rm, None, seeds, launches=7, ticks=48, overlapped=True, cls=backend.Engine)
    host = _shared_run(rm, None, seeds, launches=7, ticks=48, overlapped=True, cls=EmulEngine)
    assert np.array_equal(dev.counters()["decisions"], host.counters()["decisions"])
    assert np.array_equal(dev.shared_q_table(), host.shared_q_table())
    dev.close(); host.close()


def test_cuda_phase_timers_fill_the_reference_accumulators():
    """switch_env.py:67-73 / main.py:72-78: with ``phase_timers`` the instrumented (full) kernel splits its time into train
    ticks, observe, action selection, update and reset; the run itself is the run of the production kernel."""
    from switchfl_b200 import api
    fx, _ = load_golden("slips24_t6")
    hp = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0, seed=11)
    out = []
    for timers in (False, True):
        env = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=64, device="cuda:0", ep_cap=16, phase_timers=timers)
        m = api.DistrQLearning(env=env, **hp)
        m.learn(num_episodes=12, out_dir=None, checkpoint_freq=0)
        out.append((env, m))
    (e0, m0), (e1, m1) = out
    assert np.array_equal(m0.metrics["cum_reward"], m1.metrics["cum_reward"]) and m0.q_table == m1.q_table
    assert "KIND=full" in e1.engine.kernel_variant() and "KIND=learn" in e0.engine.kernel_variant()
    parts = [e1.flatland_step_time, e1.last_time, e1.action_selection_time, e1.update_time, e1.reset_time]
    assert all(p > 0 for p in parts) and sum(parts) <= e1.step_time * (1 + 1e-9)
    assert e0.step_time > 0 and e0.flatland_step_time == 0.0
    e0.engine.close(); e1.engine.close()


def test_cuda_learn_concurrently_equals_learning_in_turn():
    """api.learn_concurrently: one host thread + CUDA stream per learner (different maps); same results as learn() in turn."""
    from switchfl_b200 import api
    fxs = [load_golden(n)[0] for n in ("c1_synth18", "slips24_t6", "c3_rail80_s64")]

    def models():
        out = []
        for k, fx in enumerate(fxs):
            env = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=32, device="cuda:0", ep_cap=8)
            out.append(api.DistrQLearning(env=env, gamma=0.95, epsilon=0.5, epsilon_decay_rate=0.999, lr=0.2, lr_decay_rate=1.0, default_q=0.0, seed=50 + k))
        return out
    a, b = models(), models()
    for m in a:
        m.learn(num_episodes=6, out_dir=None, checkpoint_freq=0)
    api.learn_concurrently(b, 6)
    for x, y in zip(a, b):
        assert np.array_equal(x.metrics["cum_reward"], y.metrics["cum_reward"]) and np.array_equal(x.metrics["delays"], y.metrics["delays"])
        assert x.q_table == y.q_table and len(x.q_table) > 0
        x.env.engine.close(); y.env.engine.close()
