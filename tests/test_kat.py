"""CPU: the known-answer vectors of SURVEY.md Appendix C -- obtained there by running the reference's own code -- against
the host preprocessing (KAT-1), the oracle's semaphore / blocking / mask logic (KAT-2) and its Q-update (KAT-3), and the
reference's recorded port / action tables against ``railmap.build_switch_tables`` for every golden (row A0)."""
import numpy as np
import pytest

from oracle.switchfl_oracle import IN, OUT, SwitchFLOracle
from oracle.trainsim import TrainState as TS
from switchfl_b200 import mapgen, railmap
from tests._util import golden_names, load_golden

FWD, LEFT, RIGHT = railmap.MOVE_FORWARD, railmap.MOVE_LEFT, railmap.MOVE_RIGHT


def _pid(tab, r, c):
    """global port index of the reference PortId (r, c) given as floats like (1.3, 3.3)."""
    for i, p in enumerate(tab.port_ids):
        if abs(p[0] - r) < 1e-9 and abs(p[1] - c) < 1e-9:
            return i
    raise KeyError((r, c))


@pytest.fixture(scope="module")
def loop_chord():
    fx = mapgen.loop_chord_fixture()
    return fx, railmap.build_switch_tables(fx["grid"])


def test_kat1_port_graph_and_action_tables(loop_chord):
    """App. C KAT-1: switch names, port order (graph insertion order, NOT sorted by side), actions, neighbours, distances."""
    _, tab = loop_chord
    assert tab.switch_names() == ["switch_1-3", "switch_3-3"]
    assert list(tab.sw_P) == [3, 3] and list(tab.sw_A) == [5, 5]
    assert [tuple(round(x, 1) for x in p) for p in tab.port_ids] == [(1.3, 3.3), (1.1, 3.1), (1.4, 3.4), (3.2, 3.2), (3.3, 3.3), (3.1, 3.1)]
    W13, E13, S13, N33, W33, E33 = range(6)
    # a0..a3 per switch: (in, out, second train action) -> neighbour port, distance
    assert tab.actions_of(0) == [(0, 1, FWD), (0, 2, RIGHT), (1, 0, FWD), (2, 0, LEFT)]
    assert tab.actions_of(1) == [(0, 1, RIGHT), (1, 0, LEFT), (1, 2, FWD), (2, 1, FWD)]
    assert [int(tab.port_nbr[p]) for p in (E13, S13, W13)] == [E33, N33, W33]
    assert [int(tab.port_nbr[p]) for p in (W33, N33, E33)] == [W13, S13, E13]
    assert [int(tab.port_dist[p]) for p in (W13, E13, S13, N33, W33, E33)] == [5, 5, 1, 1, 5, 5]
    assert [int(x) for x in tab.port_dir] == [3, 1, 2, 0, 3, 1]                       # map_direction: .1->1 .2->0 .3->3 .4->2
    assert tab.rail_nodes[W13] == [(3, 2), (3, 1), (2, 1), (1, 2), (1, 1)]            # scrambled order (quirk A0)
    assert tab.rail_nodes[E13] == [(3, 4), (3, 5), (2, 5), (1, 5), (1, 4)] and tab.rail_nodes[S13] == [(2, 3)]


def test_kat1_non_square_grid_is_refused():
    g = np.zeros((5, 7), np.uint16)
    with pytest.raises(ValueError, match="non-square"):
        railmap.build_switch_tables(g)                                                # rail_graph.py:43-48 (App. A #17)


def _oracle(loop_chord, **hp):
    fx, tab = loop_chord
    o = SwitchFLOracle(fx, tab, **hp)
    o.reset(seed=1)
    o.semaphores = {}
    for a in o.agents:
        a.state_machine.state = TS.MOVING
    return o, tab


def test_kat2_transition_blocking_mask_extend(loop_chord):
    """App. C KAT-2 (rows E3 / E4 / E6 / O3)."""
    o, tab = _oracle(loop_chord)
    W13, S13, N33, W33 = _pid(tab, 1.3, 3.3), _pid(tab, 1.4, 3.4), _pid(tab, 3.2, 3.2), _pid(tab, 3.3, 3.3)
    o.rail_env._elapsed_steps = 10
    o.train_next_port[0], o.train_prev_port[0] = W13, None
    pin, pout, _ = o.sw_actions[0][1]                                                 # action_outcomes[1] of switch (1,3): W -> S
    assert o.transition_train(o.agents[0], pin, pout) == (1, N33)
    assert (o.train_next_port[0], o.train_prev_port[0], o.train_source_port[0]) == (N33, S13, W13)
    assert o.semaphores == {S13: [0, OUT, 2, 10, 13], N33: [0, IN, 0, 10, 12], W33: [0, OUT, 3, 10, 12]}
    # train 1 observing switch (3,3) from its W port at t = 11
    o.rail_env._elapsed_steps = 11
    t1 = o.agents[1]
    t1.position, t1.direction = (3, 2), 1
    o.train_next_port[1] = W33
    obs, mask, cur = o.observe(1, 1)
    assert obs[2:5] == [0, 1, 1] and list(mask) == [0, 0, 1, 0, 1] and cur == W33
    for t in range(11, 16):
        o.rail_env._elapsed_steps = t
        assert o.check_port_blocked(S13, N33, 1) == (t <= 13), t
    # extend_semaphores with train 0 in MALFUNCTION at t = 20: the three windows slide, nothing is added
    o.agents[0].state_machine.state = TS.MALFUNCTION
    o.train_next_port_dist[0] = 99
    o.rail_env._elapsed_steps = 20
    o.extend_semaphores()
    assert o.semaphores == {S13: [0, OUT, 2, 20, 23], N33: [0, IN, 0, 20, 22], W33: [0, OUT, 3, 20, 22]}


KAT3_S = (1, 3, 1, 1, 0, 4, 5, -1, -1, -1, -1, 0, -1, -1)
KAT3_S2 = (3, 3, 0, 1, 1, 4, 5, -1, -1, -1, -1, 1, -1, -1)
KAT3_STEPS = [(-3.0, KAT3_S2, 0, 1, "0x1.8d9999999999ap+5"), (-1304.0, KAT3_S, 0, 0, "-0x1.56ae147ae147bp+6"),
              (1000.0, None, 0, None, "0x1.6e5a1cac08310p+4")]


def test_kat3_q_update_bits(loop_chord):
    """App. C KAT-3 (rows Q1 / Q2): successive updates of Q[s][1], fp64 bit patterns; max_q ignores the mask, max_action
    honours it and takes the first maximum."""
    o, _ = _oracle(loop_chord, gamma=1.0, lr=0.1, lr_decay_rate=1.0, default_q=0.0, epsilon=0.5, epsilon_decay_rate=0.9997)
    o.q_table[KAT3_S2] = [0.0, 500.0, -7.25, 500.0, 3.0]
    for reward, nxt, prev_s, next_s, want in KAT3_STEPS:
        o.update(KAT3_S, 1, reward, nxt, prev_s, next_s)
        assert o.q_table[KAT3_S][1] == float.fromhex(want), (reward, o.q_table[KAT3_S][1].hex())
    assert max(o._row(KAT3_S2, 1)) == 500.0
    assert o.max_action(KAT3_S2, 1, np.array([0, 0, 1, 0, 1])) == 4 and o.max_action(KAT3_S2, 1, np.ones(5, int)) == 1
    assert 0.5 * 0.9997 ** 1000 == 0.37039243897165725


def test_kat3_host_learner_methods(loop_chord):
    """The drop-in DistrQLearning's host-side update / max_q / max_action (api.py) give the same bits."""
    from switchfl_b200 import api

    class _Env:                                               # just enough of the env surface for the host dict methods
        n_envs = 1

        def action_space(self, agent):
            return api.Discrete(5)
    m = api.DistrQLearning(env=_Env(), gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)
    m.q_table[KAT3_S2] = [0.0, 500.0, -7.25, 500.0, 3.0]
    ninter = {"switch_1-3": 0, "switch_3-3": 0}
    for reward, nxt, _, next_s, want in KAT3_STEPS:
        nxt_agent = None if next_s is None else ("switch_3-3" if next_s == 1 else "switch_1-3")
        m.update(KAT3_S, 1, reward, nxt, "switch_1-3", nxt_agent, ninter)
        assert m.q_table[KAT3_S][1] == float.fromhex(want)
    assert m.max_q(KAT3_S2, "switch_3-3") == 500.0 and m.max_action(KAT3_S2, "switch_3-3", [0, 0, 1, 0, 1]) == 4


@pytest.mark.parametrize("name", golden_names())
def test_switch_tables_equal_the_reference_tables(name):
    """Row A0: ``railmap.build_switch_tables`` against the tables the reference's RailNetwork built (golden ``ref_*``)."""
    fx, g = load_golden(name)
    t = railmap.build_switch_tables(fx["grid"])
    assert np.array_equal(np.stack([t.sw_P, t.sw_A], 1), g["ref_switch"])
    assert np.array_equal(np.stack([t.port_switch, t.port_side, t.port_dir], 1), g["ref_ports"])
    assert np.array_equal(t.port_nbr, g["ref_port_nbr"]) and np.array_equal(t.port_dist, g["ref_port_dist"])
    assert np.array_equal(t.port_prev_cell, g["ref_port_prev"])
    assert np.array_equal(t.port_n_intra, g["ref_port_nintra"])
    one = g["ref_port_nintra"] == 1                           # the first intra neighbour only matters on forced paths (rail_network.py:356)
    assert np.array_equal(t.port_intra0[one], g["ref_port_intra0"][one])
    s_of_act = np.repeat(np.arange(t.S), np.diff(t.sw_act0))
    assert np.array_equal(np.stack([s_of_act, t.act_in, t.act_out, t.act_move], 1), g["ref_actions"])
    off = g["ref_rail_nodes_off"]
    for p in range(t.NP):
        assert [tuple(x) for x in t.rail_nodes[p]] == [tuple(int(v) for v in rc) for rc in g["ref_rail_nodes"][off[p]:off[p + 1]]], p


KAT3_ROWS = [(0.0, 0.1, -3.0, 1.0, 500.0, 1), (float.fromhex("0x1.8d9999999999ap+5"), 0.1, -1304.0, 1.0, 0.0, 0),
             (float.fromhex("-0x1.56ae147ae147bp+6"), 0.1, 1000.0, 1.0, 0.0, 1)]


def check_device_q_update(lib=None):
    from switchfl_b200 import backend
    got = backend.device_q_update(KAT3_ROWS, lib=lib)
    assert [x.hex() for x in got] == ["0x1.8d9999999999ap+5", "-0x1.56ae147ae147bp+6", "0x1.6e5a1cac08310p+4"]
    rng = np.random.default_rng(3)                          # and against Python's own fp64 arithmetic on random operands
    rows = np.column_stack([rng.normal(0, 300, 4000), rng.uniform(0, 1, 4000), rng.normal(0, 900, 4000), rng.uniform(0, 1, 4000),
                            rng.normal(0, 500, 4000), rng.integers(0, 2, 4000)])
    got = backend.device_q_update(rows, lib=lib)
    want = [((1 - lr) * q + lr * (r + g * mq)) if b else ((1 - lr) * q + lr * r) for q, lr, r, g, mq, b in rows.tolist()]
    assert got.tolist() == want


def test_kat3_kernel_arithmetic_host_build():
    from tests.emulated import emul_library
    check_device_q_update(emul_library())


@pytest.mark.gpu
def test_kat3_kernel_arithmetic_on_the_device():
    """The same TD arithmetic the k_run kernels use, evaluated on the GPU: bit-exact with the reference's values."""
    check_device_q_update()
