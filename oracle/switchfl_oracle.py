"""ORACLE (test infrastructure, not product code) -- CPU restatement of SwitchFL's lockstep hot path.

Plain-Python restatement of the reference's switch environment + network-distributed tabular
Q-learning, one environment at a time, with ports as integer ids.  Every method cites the reference
file:line it follows (paths relative to /root/reference).  The train simulator underneath is
``oracle/trainsim.py`` (restated flatland subset, PARITY UNPINNED -- see that file's header).

Pinned by: ``tests/test_oracle_golden.py`` replays every golden vector under ``tests/golden/`` --
which ``oracle/gen_golden.py`` produced by running the reference's OWN switchfl code -- and requires
identical decisions, observations, masks, rewards, semaphore tables, tick trajectories, episode
metrics and final Q-table (bit-exact fp64).  The switchfl layer (rows A0, E1-E7, O1-O3, R1, Q1-Q6 of
SURVEY.md section 8a) is therefore pinned by reference code; rows F1-F5 are pinned only to the shared
restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; the product package never does.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Tuple

import numpy as np

from oracle import trainsim
from oracle.trainsim import RailEnvActions as RA
from oracle.trainsim import TrainState as TS

IN, OUT = 0, 1
MOVE_FORWARD, STOP_MOVING = RA.MOVE_FORWARD, RA.STOP_MOVING


class SwitchFLOracle:
    """env (switch_env.py) + learner (distr_q.py) for ONE environment.

    ``tab`` is any object exposing the port-graph tables of row A0 (H, W, switch_cells, sw_P, sw_A,
    sw_port0, sw_act0, port_dir, port_nbr, port_dist, port_prev_cell, port_n_intra, port_intra0, act_in,
    act_out, act_move, cell_switch, rail_nodes) -- the golden files carry the reference's own."""

    stop_penalty = 1300          # reward_func.py:21
    optimal_init = 500.0         # distr_q.py:44
    destination_bonus = 1000.0   # distr_q.py:45
    delay_threshold = 20         # observer.py:221

    def __init__(self, fixture: dict, tab, gamma=1.0, epsilon=0.4, epsilon_decay_rate=0.0, lr=0.4,
                 lr_decay_rate=0.0, default_q=0.0, seed=450565, max_steps=100_000):
        self.fx = fixture
        self.tab = tab
        self.W = int(tab.W)
        self.rail_env = trainsim.RailEnv(fixture)
        self.rail = self.rail_env.rail
        self.max_steps = max_steps
        self.S = len(tab.switch_cells)
        self.T = len(fixture["init_dir"])
        self.sw_ports: List[List[int]] = [list(range(int(tab.sw_port0[s]), int(tab.sw_port0[s + 1]))) for s in range(self.S)]
        self.sw_actions: List[List[Tuple[int, int, int]]] = []
        for s in range(self.S):
            p0 = int(tab.sw_port0[s])
            self.sw_actions.append([(p0 + int(tab.act_in[a]), p0 + int(tab.act_out[a]), RA(int(tab.act_move[a])))
                                    for a in range(int(tab.sw_act0[s]), int(tab.sw_act0[s + 1]))])
        self.n_actions = [int(a) for a in tab.sw_A]
        self.port_switch = [s for s in range(self.S) for _ in self.sw_ports[s]]
        self.port_dir = [int(x) for x in tab.port_dir]
        self.port_nbr = [int(x) for x in tab.port_nbr]
        self.port_dist = [int(x) for x in tab.port_dist]
        self.port_prev_cell = [int(x) for x in tab.port_prev_cell]
        self.cell_switch = {(r, c): s for s, (r, c) in enumerate(tab.switch_cells)}
        self.intra: List[List[int]] = None  # filled lazily from n_intra/intra0 (only len==1 case needs the id)
        # rail_network.py:113-128 -- prev/source ports are NOT cleared by RailNetwork.reset (rail_network.py:135-149)
        self.train_prev_port: List[Optional[int]] = [None] * self.T
        self.train_source_port: List[Optional[int]] = [None] * self.T
        # learner (distr_q.py:32-45)
        self.gamma = gamma
        self.initial_epsilon = epsilon
        self.epsilon_decay_rate = epsilon_decay_rate
        self.initial_lr = lr
        self.lr_decay_rate = lr_decay_rate
        self.default = [default_q]
        self.q_table: Dict[tuple, List[float]] = {}
        self.seed = seed
        self.agent_num_interactions = [0] * self.S
        self.total_decisions = 0
        self.trace = None

    # ------------------------------------------------------------------ env: reset (switch_env.py:93-158)
    def replay_stream(self, first: int = 0, last: Optional[int] = None) -> np.ndarray:
        """The recorded decisions as the engine's replay input: action | 0x40 where the learner exploited (max_action
        was consulted, which also inserts the row: distr_q.py:318-319, 482)."""
        a = np.array(self.trace["dec_action"][first:last], np.int8)
        g = np.array(self.trace["dec_greedy"][first:last], np.int8)
        return (a | (g << 6)).astype(np.int8)

    def malfunction_schedule(self) -> np.ndarray:
        """Every applied malfunction event (tick, train, duration) seen so far.  rail_env.reset(random_seed=seed)
        re-seeds flatland's RNG at every episode (switch_env.py:99), so the schedule repeats per episode and the
        union over episodes is THE schedule of this seed (replay input of the CUDA path)."""
        ev = set(self._malf_events) | set(self.rail_env.malfunction_events)
        return np.array(sorted(ev), np.int32).reshape(-1, 3)

    def reset(self, seed=None):
        self._malf_events = getattr(self, "_malf_events", set()) | set(self.rail_env.malfunction_events)
        self.rail_env.reset(random_seed=seed)
        agents = self.rail_env.agents
        keys = [a.initial_position + (a.initial_direction,) for a in agents]
        assert keys == sorted(keys), "fixture trains must be pre-sorted (switch_env.py:104-119)"
        self.agents = agents
        # rail_network.py:135-149
        self.train_next_port: List[Optional[int]] = [None] * self.T
        self.train_next_port_dist: List[Optional[int]] = [None] * self.T
        self.semaphores: Dict[int, list] = {}
        self.terminated = False
        self.truncated = False
        self.cumulative_rewards = [[0] * self.T for _ in range(self.S)]     # switch_env.py:130
        self.step_counter = 0
        self.train_action_plan: List[List[RA]] = [[] for _ in range(self.T)]
        self.rail_env_time = 0
        self.train_done = {h: False for h in range(self.T)}
        self.malfunctions = set()
        self.num_malfunctions = 0
        self.active_switch_agents: List[int] = []
        self.active_trains: List[int] = []
        self.prev_actions: List[Optional[RA]] = [None] * self.T
        self.train_to_last_node = [(None, self.compute_delay(a, a.initial_position, a.initial_direction, True))
                                   for a in agents]                          # switch_env.py:151-152
        self._init_ports()
        self._move_trains_to_switch()

    def _init_ports(self):
        """switch_env.py:507-568."""
        for train in self.agents:
            pos, d = train.position, train.direction
            if pos is None or d is None:
                pos, d = train.initial_position, train.initial_direction
            last_pos = train.old_position
            distance = 0
            while pos not in self.cell_switch:
                last_pos = pos
                nxt = self.rail.get_valid_move_actions_(d, pos)
                _, (pos, d), _, _ = self.rail.check_action_on_agent(nxt[0].action, (pos, d))
                distance += 1
            s = self.cell_switch[pos]
            port = None
            for p in self.sw_ports[s]:
                if self.port_prev_cell[p] == last_pos[0] * self.W + last_pos[1]:
                    port = p
                    break
            self.train_next_port[train.handle] = port
            self.train_next_port_dist[train.handle] = distance
        for train in self.agents:
            port = self.train_next_port[train.handle]
            self.semaphores[port] = [train.handle, IN, self.port_dir[port], train.earliest_departure - 2,
                                     train.earliest_departure + self.train_next_port_dist[train.handle]]

    # ------------------------------------------------------------------ observer.py
    def compute_delay(self, train, position, direction, earliest_departure=False):
        """observer.py:18-42 (float result: the distance map is float64 with inf fill)."""
        d = self.rail_env.distance_map.get(self.rail_env.agents)[train.handle, position[0], position[1], direction]
        if np.isinf(d):
            raise ValueError("Infinite distance to target encountered.")
        if earliest_departure:
            return train.earliest_departure - train.latest_arrival + d
        return self.rail_env._elapsed_steps - train.latest_arrival + d

    def check_port_blocked(self, next_port, out_port, me) -> bool:
        """observer.py:44-151, clause by clause."""
        sem = self.semaphores
        now = self.rail_env._elapsed_steps
        agents = self.agents

        def rule_next(port):   # observer.py:55-84 and :119-149
            r = sem.get(port)
            if r is None or r[0] == me or not (r[3] <= now <= r[4]):
                return False
            same = r[2] == self.port_dir[port]
            malf = agents[r[0]].state == TS.MALFUNCTION
            if r[1] == OUT:
                return same or malf
            return (not same) or malf

        def rule_out(port):    # observer.py:86-116
            r = sem.get(port)
            if r is None or r[0] == me or not (r[3] <= now <= r[4]):
                return False
            same = r[2] == self.port_dir[port]
            malf = agents[r[0]].state == TS.MALFUNCTION
            if r[1] == OUT:
                return (not same) or malf
            return same or malf

        if next_port is not None:
            if rule_next(next_port):
                return True
            return rule_out(out_port)
        return rule_next(out_port)

    def observe(self, s: int, train_h: int):
        """observer.py:246-308 -> (obs list, mask int8 array, current_port)."""
        train = self.agents[train_h]
        ports = self.sw_ports[s]
        sem, target, delay = [], [], []
        current_port = None
        for port in ports:
            blocked = self.check_port_blocked(self.port_nbr[port], port, train_h)
            sem.append(0 if blocked else 1)
            if self.train_next_port[train_h] == port:
                current_port = port
                dl = self.compute_delay(train, train.position, train.direction)
                avail = train.latest_arrival - train.earliest_departure          # observer.py:239-244
                delay.append(0 if dl <= 0 else (1 if dl <= avail * self.delay_threshold else 2))
                target.extend(train.target)
            else:
                delay.append(-1)
                target.extend([-1, -1])
        if current_port is None:
            raise RuntimeError("No train detected at active switch (observer.py:294-307 would raise UnboundLocalError)")
        r, c = self.tab.switch_cells[s]
        obs = [int(r), int(c), *sem, *[int(x) for x in target], *delay]
        # switch_agents.py:104-134
        mask = np.array([1 if (pin == current_port and sem[pout - ports[0]]) else 0
                         for (pin, pout, _) in self.sw_actions[s]] + [1], dtype=np.int8)
        return obs, mask, current_port

    # ------------------------------------------------------------------ rail_network.py
    def _delete_owned(self, port, h):
        for p in self.sw_ports[self.port_switch[port]]:
            r = self.semaphores.get(p)
            if r is not None and r[0] == h:
                del self.semaphores[p]

    def transition_semaphore(self, source, out_port, target, train):
        """rail_network.py:303-416, step by step."""
        sem, now, h = self.semaphores, self.rail_env._elapsed_steps, train.handle
        pd, pdir = self.port_dist, self.port_dir
        if train.state != TS.MALFUNCTION:                                       # :315-323
            self._delete_owned(self.train_next_port[h], h)
            if self.train_prev_port[h] is not None:
                self._delete_owned(self.train_prev_port[h], h)
        if out_port not in sem:                                                  # :326-334
            sem[out_port] = [h, OUT, pdir[out_port], now, now + 3]
        elif sem[out_port][1] == OUT or sem[out_port][3] > now:
            sem[out_port][0] = h; sem[out_port][3] = now; sem[out_port][4] = now + 3
        d_ot = pd[out_port]
        if target not in sem:                                                    # :336-344
            sem[target] = [h, IN, pdir[target], now, now + d_ot + 1]
        elif sem[target][1] == IN or sem[target][3] > now:
            sem[target][0] = h; sem[target][3] = now; sem[target][4] = now + d_ot + 1
        # :346-353 edges(target) minus the moving edge; intra-switch edges carry rail_nodes == [] (distance 0)
        if int(self.tab.port_n_intra[target]) == 1:                              # :356
            unique = int(self.tab.port_intra0[target])
            far = self.port_nbr[unique]                                          # :358-364 prox_list[0]
            if unique != source and unique != out_port and unique != target:     # :368-378
                rec = [h, OUT, pdir[unique], now, now + d_ot + 0 + 1]
                if unique not in sem or sem[unique][1] == OUT or sem[unique][3] > now:
                    sem[unique] = rec
            if unique not in sem or sem[unique][3] > now:                         # :380-388 (list == 'out' is never True)
                sem[unique] = [h, OUT, pdir[unique], now, now + d_ot + 0]
            for port in (unique, far):                                           # :390-402
                if port != source and port != out_port and port != unique:
                    rec = [h, IN, pdir[port], now, now + d_ot + 0 + pd[unique] + 1]
                    if port not in sem or sem[port][1] == IN or sem[port][3] > now:
                        sem[port] = rec
        for port in (target, out_port):                                          # :404-414 moving edge
            if port != source and port != out_port:
                rec = [h, OUT, pdir[port], now, now + pd[out_port] + 1]
                if port not in sem or sem[port][1] == OUT or sem[port][3] > now:
                    sem[port] = rec

    def transition_train(self, train, in_port, out_port):
        """rail_network.py:246-278."""
        assert self.port_switch[in_port] == self.port_switch[out_port]
        target = self.port_nbr[out_port]
        self.transition_semaphore(in_port, out_port, target, train)
        h = train.handle
        self.train_source_port[h] = in_port
        self.train_next_port[h] = target
        self.train_prev_port[h] = out_port
        return self.port_switch[target], target

    def extend_semaphores(self):
        """rail_network.py:229-244."""
        now = self.rail_env._elapsed_steps
        for train in self.agents:
            st = train.state
            if st == TS.STOPPED or st == TS.MALFUNCTION:
                for p, r in self.semaphores.items():
                    if r[0] == train.handle:
                        dist = r[4] - r[3]
                        r[3] = now
                        r[4] = now + dist
            if st == TS.MALFUNCTION:
                port = self.train_next_port[train.handle]
                if port not in self.semaphores:
                    self.semaphores[port] = [train.handle, IN, self.port_dir[port], now,
                                             now + self.train_next_port_dist[train.handle]]

    # ------------------------------------------------------------------ switch_env.py hot loops
    def _move_trains(self):
        """switch_env.py:296-401."""
        env, rail = self.rail_env, self.rail
        actions, expected = {}, {}
        for train in self.agents:
            h = train.handle
            if self.train_done[h]:
                continue
            plan = self.train_action_plan[h]
            if not plan:
                actions[h] = MOVE_FORWARD
            else:
                self.prev_actions[h] = plan[0]
                actions[h] = plan.pop(0)
            if train.position is not None:
                _, (npos, _), valid, _ = rail.check_action_on_agent(actions[h], (train.position, train.direction))
                expected[h] = (npos, True) if valid else (train.position, False)
        _, _, self.train_done, info = env.step(actions)
        if self.trace is not None:
            self._trace_tick()
        for train in self.agents:
            h = train.handle
            if h in expected:
                epos, valid = expected[h]
                if epos != train.position and valid and actions[h] != STOP_MOVING:
                    self.train_action_plan[h].insert(0, actions[h])
                    if epos in self.cell_switch:
                        self.train_next_port[h] = self.train_source_port[h]
            if self.train_done[h]:
                for p in [p for p, r in self.semaphores.items() if r[0] == h]:
                    del self.semaphores[p]
        for train in self.agents:
            if env._elapsed_steps == train.earliest_departure - 2:
                port = self.train_next_port[train.handle]
                self.semaphores[port] = [train.handle, IN, self.port_dir[port], train.earliest_departure - 2,
                                         train.earliest_departure + self.train_next_port_dist[train.handle]]
        self.extend_semaphores()
        self.rail_env_time += 1
        if self.train_done["__all__"]:
            self.terminated = True
        new = {h for h, v in info["malfunction"].items() if v != 0}
        self.num_malfunctions += len(new - self.malfunctions)
        self.malfunctions = new

    def _check_active_switch(self):
        """switch_env.py:427-485."""
        for train in self.agents:
            if train.position is None or train.state == TS.WAITING:
                continue
            h = train.handle
            plan = self.train_action_plan[h]
            nxt = plan[0] if plan else MOVE_FORWARD
            _, (npos, _), _, _ = self.rail.check_action_on_agent(nxt, (train.position, train.direction))
            s = self.cell_switch.get(npos)
            if s is None:
                continue
            st = train.state
            if st == TS.READY_TO_DEPART or st == TS.MOVING:
                self.active_switch_agents.append(s); self.active_trains.append(h)
            elif st in (TS.STOPPED, TS.MALFUNCTION) and self.prev_actions[h] == STOP_MOVING:
                self.active_switch_agents.append(s); self.active_trains.append(h)
            elif st in (TS.STOPPED, TS.MALFUNCTION):
                self.active_switch_agents.append(self.port_switch[self.train_next_port[h]]); self.active_trains.append(h)

    def _move_trains_to_switch(self):
        """switch_env.py:403-424."""
        while not self.active_switch_agents and not self.terminated:
            self._move_trains()
            self._check_active_switch()
        order = sorted(range(len(self.active_trains)), key=lambda i: self.active_trains[i])
        self.active_switch_agents = [self.active_switch_agents[i] for i in order]
        self.active_trains = [self.active_trains[i] for i in order]

    def reward_func(self, train, plan, port_blocked):
        """reward_func.py:23-78."""
        pos, d = train.position, train.direction
        for a in plan:
            if a != STOP_MOVING:
                _, (pos, d), _, _ = self.rail.check_action_on_agent(a, (pos, d))
        curr = self.compute_delay(train, pos, d)
        diff = self.train_to_last_node[train.handle][1] - curr
        if sum(port_blocked) == len(port_blocked) and not (len(port_blocked) == 1 and not port_blocked[0]):
            reward = diff
        elif plan[0] == STOP_MOVING:
            reward = diff - self.stop_penalty
        else:
            reward = diff
        return reward, curr

    def apply_action(self, s: int, h: int, action: int) -> int:
        """switch_env.py:203-294 (+ switch_agents.py:136-168)."""
        assert 0 <= action < self.n_actions[s], "Invalid action performed."        # switch_env.py:213-215
        acts = self.sw_actions[s]
        port_node = self.train_next_port[h]
        moving = None
        if action == len(acts):
            nta = [STOP_MOVING]
        elif acts[action][0] == port_node:
            nta = [MOVE_FORWARD, acts[action][2]]
            moving = h
        else:
            nta = [STOP_MOVING, STOP_MOVING]
        if nta[0] == STOP_MOVING:
            in_port = out_port = port_node
        else:
            in_port, out_port = acts[action][0], acts[action][1]
        train = self.agents[h]
        if moving is not None:
            next_switch, next_port = self.transition_train(train, in_port, out_port)
        else:
            next_switch, next_port = s, None
        plan = self.train_action_plan[h]
        if moving is not None and len(plan) > 0:                                   # :257-266
            nta.pop(0)
            if len(plan) > 1:
                plan = self.train_action_plan[h] = plan[:1]
            plan.extend(nta)
        elif moving is None:                                                       # :267-270
            plan.insert(0, STOP_MOVING)
        else:
            plan.extend(nta)
        if moving is not None:                                                     # :274-282
            blocked = [self.check_port_blocked(next_port, out_port, h)]
        else:
            blocked = []
            for (pin, pout, _) in acts:
                if pin == in_port:
                    blocked.append(self.check_port_blocked(self.port_nbr[pout], pout, h))
        reward, curr = self.reward_func(train, plan, blocked)
        self.cumulative_rewards[next_switch][h] = reward                           # :289
        self.train_to_last_node[h] = (s, curr)                                     # :291
        return next_switch

    def env_step(self, s: int, h: int, action: int):
        """switch_env.py:632-666."""
        next_switch = self.apply_action(s, h, action)
        if not self.active_switch_agents:
            self._move_trains_to_switch()
        self.step_counter += 1
        if self.step_counter > self.max_steps:
            self.truncated = True
        arrived = [t.handle for t in self.agents if t.position is None and t.arrival_time is not None]
        return next_switch, arrived

    # ------------------------------------------------------------------ learner (distr_q.py)
    def _row(self, state, s) -> List[float]:
        """distr_q.py:47-57 __check_entry."""
        k = tuple(state)
        row = self.q_table.get(k)
        if row is None:
            row = self.q_table[k] = self.default * self.n_actions[s]
        return row

    def max_action(self, state, s, mask) -> int:
        """distr_q.py:468-490."""
        row = self._row(state, s)
        a = int(np.argmax(row))
        if mask[a]:
            return a
        allowed = np.nonzero(mask)[0]
        return int(allowed[np.argmax(np.array(row)[allowed])])

    def update(self, state, action, reward, next_state, prev_s, next_s):
        """distr_q.py:419-447 (Python operator order, fp64)."""
        row = self._row(state, prev_s)
        lr = self.initial_lr * (self.lr_decay_rate ** self.agent_num_interactions[prev_s])
        if next_s != prev_s:
            mq = 0.0 if next_state is None else max(self._row(next_state, next_s))   # distr_q.py:449-466
            row[action] = (1 - lr) * row[action] + lr * (reward + self.gamma * mq)
        else:
            row[action] = (1 - lr) * row[action] + lr * reward

    def init_q_table(self):
        """distr_q.py:81-181."""
        env = self.rail_env
        for agent in self.agents:
            path = env.distance_map.get_shortest_paths(max_depth=None, agents=self.agents, agent_handle=agent.handle)[agent.handle]
            for i, wp in enumerate(path):
                s = self.cell_switch.get(tuple(wp.position))
                if s is None:
                    continue
                ports = self.sw_ports[s]
                P = len(ports)
                in_port = self._port_at(wp.position, wp.direction)
                sems = list(itertools.product([0, 1], repeat=P))[1:]
                tg = [-1] * (2 * P)
                for k, p in enumerate(ports):
                    if p == in_port:
                        tg[2 * k], tg[2 * k + 1] = agent.target
                states = []
                for sv in sems:
                    for lvl in range(3):
                        dl = [lvl if p == in_port else -1 for p in ports]
                        states.append((int(wp.position[0]), int(wp.position[1]), *sv, *tg, *dl))
                nxt = None
                for wp2 in path[i + 1:]:
                    if tuple(wp2.position) in self.cell_switch:
                        nxt = wp2
                        break
                best, opt = float("inf"), None
                for a, (pin, pout, _) in enumerate(self.sw_actions[s]):
                    if pin != in_port:
                        continue
                    if nxt is not None:
                        if self.port_nbr[pout] == self._port_at(nxt.position, nxt.direction) and self.port_dist[pout] < best:
                            best, opt = self.port_dist[pout], a
                    else:
                        for dpos, node in enumerate(self.tab.rail_nodes[pout]):
                            if tuple(node) == tuple(agent.target):
                                if dpos < best:
                                    best, opt = dpos, a
                                break
                if opt is None:
                    raise RuntimeError("q-init without optimal action (distr_q.py:158 would reuse a stale index)")
                for st in states:
                    row = self.q_table[st] = self.default * self.n_actions[s]
                    row[opt] = self.optimal_init if nxt is not None else self.destination_bonus

    def _port_at(self, position, direction) -> Optional[int]:
        """distr_q.py:95-96: in-port = position + map_inverse_direction(direction)/10."""
        side = {1: 3, 0: 4, 3: 1, 2: 2}[int(direction)]       # rail_network.py:292-301
        want_dir = {1: 1, 2: 0, 3: 3, 4: 2}[side]             # rail_network.py:280-290
        s = self.cell_switch[tuple(position)]
        for p in self.sw_ports[s]:
            if self.port_dir[p] == want_dir:
                return p
        return None

    def run_episode(self, rng, greedy=False, learn=True, replay_actions=None):
        """One pass of distr_q.py:296-366 (learn) or :195-224 (test).  Returns the episode metrics."""
        self.reset(seed=self.seed)
        if learn and not getattr(self, "_q_inited", False):
            self.init_q_table()                                                    # distr_q.py:299-300
            self._q_inited = True
        update_dict: Dict[Tuple[int, int], tuple] = {}
        at_dest: List[int] = []
        cum_reward, num_iter = 0.0, 0
        arrived: List[int] = []
        while not (self.terminated or self.truncated):                             # switch_env.py:616-622
            s = self.active_switch_agents.pop(0)
            h = self.active_trains.pop(0)
            obs, mask, _ = self.observe(s, h)
            reward = self.cumulative_rewards[s][h]
            tick = self.rail_env._elapsed_steps
            was_greedy = 0
            if replay_actions is not None:
                action = int(replay_actions[self.total_decisions])
                if action & 0x40:                                                  # recorded as an exploit choice
                    action &= 0x3F
                    was_greedy = 1
                    assert self.max_action(obs, s, mask) == action, "replayed greedy action differs from argmax"
            elif greedy:
                action = self.max_action(obs, s, mask)
                was_greedy = 1
            else:
                eps = self.initial_epsilon * (self.epsilon_decay_rate ** self.agent_num_interactions[s])
                if rng.random() < eps:                                             # distr_q.py:315-317
                    sd = int(rng.integers(0, np.iinfo(np.int32).max))
                    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(sd)))
                    valid = np.where(mask == 1)[0]
                    action = int(g.choice(valid)) if len(valid) else 0
                else:
                    action = self.max_action(obs, s, mask)
                    was_greedy = 1
            next_switch, arrived = self.env_step(s, h, action)
            if self.trace is not None:
                self._trace_decision(tick, s, h, obs, mask, reward, action, next_switch, arrived)
                self.trace["dec_greedy"].append(was_greedy)
            if learn:
                if (s, h) in update_dict:                                          # distr_q.py:329-338
                    pobs, pact, pagent = update_dict.pop((s, h))
                    self.update(pobs, pact, reward, obs, pagent, s)
                update_dict[(next_switch, h)] = (obs, action, s)                   # :340-342
                for tr in arrived:                                                 # :345-356
                    if tr not in at_dest:
                        at_dest.append(tr)
                        for (ua, ut), (uo, uact, uprev) in list(update_dict.items()):
                            if ut == tr:
                                self.update(uo, uact, self.destination_bonus, None, uprev, None)
                                del update_dict[(ua, ut)]
            cum_reward += reward
            num_iter += 1
            self.total_decisions += 1
            if learn:
                self.agent_num_interactions[s] += 1
        return dict(cum_reward=cum_reward, decisions=num_iter, arrived=len(arrived),
                    delays=[v[1] for v in self.train_to_last_node], num_malfunctions=self.num_malfunctions,
                    ticks=self.rail_env._elapsed_steps)

    def learn(self, num_episodes: int, replay_actions=None):
        """distr_q.py:244-379 without the file outputs; returns the list of per-episode metric dicts."""
        rng = np.random.default_rng(self.seed)
        self.episode = -1
        out = []
        for t in range(num_episodes):
            self.episode = t
            out.append(self.run_episode(rng, greedy=False, learn=True, replay_actions=replay_actions))
        return out

    def test(self):
        """distr_q.py:184-241: greedy rollout, no updates (still inserts default rows via max_action)."""
        return self.run_episode(None, greedy=True, learn=False)

    # ------------------------------------------------------------------ tracing (same layout as tests/golden)
    def enable_trace(self):
        self.trace = {k: [] for k in ("dec_ep", "dec_tick", "dec_switch", "dec_train", "dec_obs", "dec_mask", "dec_reward",
                                      "dec_action", "dec_next_switch", "dec_arrived", "dec_sem", "dec_done", "dec_greedy",
                                      "tick_ep", "tick_tick", "tick_pos", "tick_dir", "tick_state", "tick_malf")}

    def _trace_tick(self):
        t, W = self.trace, self.W
        t["tick_ep"].append(self.episode); t["tick_tick"].append(self.rail_env._elapsed_steps)
        t["tick_pos"].append([-1 if a.position is None else a.position[0] * W + a.position[1] for a in self.agents])
        t["tick_dir"].append([int(a.direction) for a in self.agents])
        t["tick_state"].append([int(a.state) for a in self.agents])
        t["tick_malf"].append([a.malfunction_handler.malfunction_down_counter for a in self.agents])

    def _trace_decision(self, tick, s, h, obs, mask, reward, action, next_switch, arrived):
        t = self.trace
        o = np.full(18, -9, np.int64); o[:len(obs)] = obs
        m = np.full(9, -1, np.int8); m[:len(mask)] = mask
        NP = len(self.port_dir)
        sem = np.full((NP, 4), -1, np.int32)
        for p, (tr, typ, d, t0, t1) in self.semaphores.items():
            sem[p] = (tr, typ, t0, t1)
        t["dec_ep"].append(self.episode); t["dec_tick"].append(tick); t["dec_switch"].append(s); t["dec_train"].append(h)
        t["dec_obs"].append(o); t["dec_mask"].append(m); t["dec_reward"].append(float(reward)); t["dec_action"].append(action)
        t["dec_next_switch"].append(next_switch); t["dec_arrived"].append(sum(1 << int(x) for x in arrived))
        t["dec_sem"].append(sem); t["dec_done"].append(int(self.terminated) | (int(self.truncated) << 1))
