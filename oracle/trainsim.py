"""ORACLE (test infrastructure, not product code) -- restated flatland-rl subset.

This file restates, from the published behaviour of the third-party package
``flatland-rl`` (listed UNPINNED in /root/reference/requirements.txt:5 and absent
from /root/reference and from this image), exactly the subset of the train
simulator that the reference's hot path calls into (SURVEY.md section 8c lists
every call site; rows F1-F6 of section 8a).  It is the substrate under BOTH

  * the reference's own ``switchfl`` code when it is executed here to produce
    golden vectors (``oracle/gen_golden.py`` imports /root/reference/switchfl on
    top of ``oracle/shim``, whose ``flatland.*`` modules re-export this file), and
  * the standalone oracle ``oracle/switchfl_oracle.py``.

PARITY UNPINNED at this boundary: the reference ships no test, fixture or golden
vector for anything that flows through flatland, and flatland itself cannot be
run here.  Everything in this file is therefore [UPSTREAM-UNVERIFIED]; it follows
SURVEY.md Appendix B so that the shim, the oracle and the CUDA kernel implement
one stated semantics, and a later correction is a one-place change.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from collections import deque, namedtuple
from enum import IntEnum
from typing import Dict, List, Optional, Tuple

import numpy as np


# --------------------------------------------------------------------------- enums
class RailEnvActions(IntEnum):
    DO_NOTHING = 0
    MOVE_LEFT = 1
    MOVE_FORWARD = 2
    MOVE_RIGHT = 3
    STOP_MOVING = 4

    def is_moving_action(self) -> bool:
        return self in (RailEnvActions.MOVE_LEFT, RailEnvActions.MOVE_FORWARD, RailEnvActions.MOVE_RIGHT)


class TrainState(IntEnum):
    WAITING = 0
    READY_TO_DEPART = 1
    MALFUNCTION_OFF_MAP = 2
    MOVING = 3
    STOPPED = 4
    MALFUNCTION = 5
    DONE = 6

    def is_off_map_state(self) -> bool:
        return self in (TrainState.WAITING, TrainState.READY_TO_DEPART, TrainState.MALFUNCTION_OFF_MAP)

    def is_on_map_state(self) -> bool:
        return self in (TrainState.MOVING, TrainState.STOPPED, TrainState.MALFUNCTION)


class Grid4TransitionsEnum(IntEnum):
    NORTH = 0
    EAST = 1
    SOUTH = 2
    WEST = 3


_DELTA = ((-1, 0), (0, 1), (1, 0), (0, -1))  # N, E, S, W


def get_new_position(position, movement):
    return (position[0] + _DELTA[movement][0], position[1] + _DELTA[movement][1])


Waypoint = namedtuple("Waypoint", ["position", "direction"])
RailEnvNextAction = namedtuple("RailEnvNextAction", ["action", "next_position", "next_direction"])


# --------------------------------------------------------------------------- F1: grid transitions
class RailGridTransitionMap:
    """uint16 transition grid; bit (15 - (4*heading + exit))."""

    def __init__(self, grid: np.ndarray):
        self.grid = np.asarray(grid, dtype=np.uint16)
        self.height, self.width = self.grid.shape

    def get_full_transitions(self, row, col) -> int:
        return int(self.grid[row, col])

    def get_transitions(self, configuration) -> Tuple[int, int, int, int]:
        (row, col), direction = configuration
        v = int(self.grid[row, col])
        nib = (v >> ((3 - int(direction)) * 4)) & 0xF
        return ((nib >> 3) & 1, (nib >> 2) & 1, (nib >> 1) & 1, nib & 1)

    def get_transition(self, cell_orientation, exit_direction) -> int:
        if len(cell_orientation) == 2:  # configuration form ((row, col), heading)
            (row, col), direction = cell_orientation
        else:
            row, col, direction = cell_orientation
        v = int(self.grid[row, col])
        return (v >> (15 - (4 * int(direction) + int(exit_direction)))) & 1

    def check_bounds(self, position) -> bool:
        return 0 <= position[0] < self.height and 0 <= position[1] < self.width

    def check_action_on_agent(self, action, configuration):
        """F1.  Returns (new_cell_valid, (new_pos, new_dir), transition_valid, action)."""
        position, direction = configuration
        direction = int(direction)
        trans = self.get_transitions((position, direction))
        n = trans[0] + trans[1] + trans[2] + trans[3]
        transition_valid = None
        new_direction = direction
        preprocessed = action
        if action == RailEnvActions.MOVE_LEFT:
            new_direction = direction - 1
            if n <= 1:
                transition_valid = False
        elif action == RailEnvActions.MOVE_RIGHT:
            new_direction = direction + 1
            if n <= 1:
                transition_valid = False
        new_direction %= 4
        if action == RailEnvActions.MOVE_FORWARD and n == 1:
            new_direction = trans.index(1)
            transition_valid = True
        new_position = get_new_position(position, new_direction)
        new_cell_valid = self.check_bounds(new_position) and self.get_full_transitions(*new_position) > 0
        if transition_valid is None:
            transition_valid = bool(trans[new_direction])
        if not transition_valid and action in (RailEnvActions.MOVE_LEFT, RailEnvActions.MOVE_RIGHT):
            preprocessed = RailEnvActions.MOVE_FORWARD
        return new_cell_valid, (new_position, new_direction), bool(transition_valid), preprocessed

    def get_valid_move_actions_(self, agent_direction, agent_position) -> List[RailEnvNextAction]:
        """Ordered (left, forward, right); a single transition is reported as MOVE_FORWARD."""
        agent_direction = int(agent_direction)
        trans = self.get_transitions((agent_position, agent_direction))
        n = sum(trans)
        out: List[RailEnvNextAction] = []
        if n == 1:
            nd = trans.index(1)
            out.append(RailEnvNextAction(RailEnvActions.MOVE_FORWARD, get_new_position(agent_position, nd), nd))
            return out
        for i, act in ((-1, RailEnvActions.MOVE_LEFT), (0, RailEnvActions.MOVE_FORWARD), (1, RailEnvActions.MOVE_RIGHT)):
            nd = (agent_direction + i) % 4
            if trans[nd]:
                out.append(RailEnvNextAction(act, get_new_position(agent_position, nd), nd))
        return out


# --------------------------------------------------------------------------- F6: distance map (own restatement)
class DistanceMap:
    """Restates /root/reference/flatland_patch/distance_map.py:62-242 (BFS + greedy shortest path).

    ``oracle/gen_golden.py`` cross-checks this against the vendored file itself.
    """

    def __init__(self, agents, env_height, env_width):
        self.env_height = env_height
        self.env_width = env_width
        self.distance_map = None
        self.agents = agents
        self.rail: Optional[RailGridTransitionMap] = None

    def reset(self, agents, rail):
        self.agents = agents
        self.rail = rail
        self.env_height = rail.height
        self.env_width = rail.width
        self.distance_map = None

    def get(self, agents=None) -> np.ndarray:
        if self.distance_map is None:
            self._compute(self.agents if agents is None else agents, self.rail)
        return self.distance_map

    def _compute(self, agents, rail):
        # distance_map.py:71-86 -- one BFS per distinct target, rows copied for repeats
        dm = np.full((len(agents), self.env_height, self.env_width, 4), np.inf)
        done: Dict[Tuple[int, int], int] = {}
        for i, agent in enumerate(agents):
            tgt = tuple(agent.target)
            if tgt in done:
                dm[i] = dm[done[tgt]]
            else:
                dm[i] = bfs_to_target(rail.grid, tgt)
                done[tgt] = i
        self.distance_map = dm

    def get_shortest_paths(self, max_depth=None, agents=None, agent_handle=None):
        # distance_map.py:170-242
        agents = agents if agents else self.agents
        out = {}
        for agent in agents:
            if agent_handle is not None and agent.handle != agent_handle:
                continue
            out[agent.handle] = shortest_path(self.rail, self.get(agents)[agent.handle], agent, max_depth)
        return out


def bfs_to_target(grid: np.ndarray, target) -> np.ndarray:
    """distance_map.py:88-167: reverse BFS over (cell, heading); unit edge cost; inf = unreachable."""
    H, W = grid.shape
    dist = np.full((H, W, 4), np.inf)
    dist[target[0], target[1], :] = 0
    visited = {(target[0], target[1], d) for d in range(4)}

    def neighbors(position, current_distance, enforce):
        res = []
        dirs = (0, 1, 2, 3) if enforce < 0 else ((enforce + 2) % 4,)
        for nd in dirs:
            nr, nc = position[0] + _DELTA[nd][0], position[1] + _DELTA[nd][1]
            if 0 <= nr < H and 0 <= nc < W:
                want = (nd + 2) % 4
                v = int(grid[nr, nc])
                for orient in range(4):
                    if (v >> (15 - (4 * orient + want))) & 1:
                        nd_ = min(dist[nr, nc, orient], current_distance + 1)
                        res.append((nr, nc, orient, nd_))
                        dist[nr, nc, orient] = nd_
        return res

    q = deque(neighbors(target, 0, -1))
    while q:
        node = q.popleft()
        nid = (node[0], node[1], node[2])
        if nid not in visited:
            visited.add(nid)
            for nb in neighbors((node[0], node[1]), node[3], node[2]):
                q.append(nb)
    return dist


def shortest_path(rail: RailGridTransitionMap, dist: np.ndarray, agent, max_depth=None) -> List[Waypoint]:
    """distance_map.py:195-232: greedy descent with strict '<' tie-break, candidates in L,F,R order."""
    if agent.state.is_off_map_state():
        position = agent.initial_position
    elif agent.state.is_on_map_state():
        position = agent.position
    elif agent.state == TrainState.DONE:
        position = agent.target
    else:
        return None
    direction = agent.direction
    path: List[Waypoint] = []
    distance = math.inf
    depth = 0
    while position != agent.target and (max_depth is None or depth < max_depth):
        best = None
        for na in rail.get_valid_move_actions_(direction, position):
            d = dist[na.next_position[0], na.next_position[1], na.next_direction]
            if d < distance:
                best = na
                distance = d
        path.append(Waypoint(position, direction))
        depth += 1
        if best is None:
            return path
        position = best.next_position
        direction = best.next_direction
    if max_depth is None or depth < max_depth:
        path.append(Waypoint(position, direction))
    return path


# --------------------------------------------------------------------------- F5: malfunctions
class MalfunctionParameters:
    def __init__(self, malfunction_rate=0.0, min_duration=0, max_duration=0):
        self.malfunction_rate = malfunction_rate
        self.min_duration = min_duration
        self.max_duration = max_duration


class ParamMalfunctionGen:
    """SURVEY Appendix B 'Malfunction draw (F5)'."""

    def __init__(self, parameters: MalfunctionParameters):
        self.mean_malfunction_rate = parameters.malfunction_rate
        self.min_number_of_steps_broken = parameters.min_duration
        self.max_number_of_steps_broken = parameters.max_duration

    def prob(self) -> float:
        r = self.mean_malfunction_rate
        return 0.0 if r <= 0 else 1.0 - math.exp(-r)

    def generate(self, np_random) -> int:
        if np_random.rand() < self.prob():
            return int(np_random.randint(self.min_number_of_steps_broken, self.max_number_of_steps_broken + 1)) + 1
        return 0


class _MalfunctionHandler:
    __slots__ = ("malfunction_down_counter", "num_malfunctions")

    def __init__(self):
        self.malfunction_down_counter = 0
        self.num_malfunctions = 0

    @property
    def in_malfunction(self):
        return self.malfunction_down_counter > 0


# --------------------------------------------------------------------------- agents
class _StateMachine:
    __slots__ = ("state", "previous_state")

    def __init__(self):
        self.state = TrainState.WAITING
        self.previous_state = None


class EnvAgent:
    """Field set of /root/reference/flatland_patch/agent_utils.py:68-105 that the hot path reads."""

    def __init__(self, handle, initial_position, initial_direction, target, earliest_departure, latest_arrival):
        self.handle = handle
        self.initial_position = tuple(int(x) for x in initial_position)
        self.initial_direction = int(initial_direction)
        self.target = tuple(int(x) for x in target)
        self.earliest_departure = int(earliest_departure)
        self.latest_arrival = int(latest_arrival)
        self.state_machine = _StateMachine()
        self.malfunction_handler = _MalfunctionHandler()
        self.reset()

    def reset(self):  # agent_utils.py:107-123
        self.position = None
        self.direction = self.initial_direction
        self.old_position = None
        self.old_direction = None
        self.moving = False
        self.arrival_time = None
        self.saved_action = None
        self.malfunction_handler = _MalfunctionHandler()
        self.state_machine = _StateMachine()

    @property
    def state(self) -> TrainState:
        return self.state_machine.state

    @state.setter
    def state(self, s):
        self.state_machine.state = s


# --------------------------------------------------------------------------- F3: motion check
def resolve_motion(olds: List[Optional[tuple]], news: List[Optional[tuple]]) -> List[bool]:
    """F3.  ``olds[i]``/``news[i]`` are agent i's current and tentative cells (None = off map).

    Returns motion_ok[i] = the agent wanted to move and may.  Rules:
    stationary occupants block their followers transitively; two agents exchanging cells are
    both blocked (with their followers); of several agents entering one cell the lowest
    handle wins and the others (with their followers) are blocked.
    """
    n = len(olds)
    src = [o if o is not None else (-1, i) for i, o in enumerate(olds)]
    dst = [news[i] if news[i] is not None else src[i] for i in range(n)]
    occupant = {src[i]: i for i in range(n)}
    wants = [dst[i] != src[i] for i in range(n)]
    blocked = [not w for w in wants]
    # swaps
    for i in range(n):
        if wants[i]:
            j = occupant.get(dst[i])
            if j is not None and j != i and wants[j] and dst[j] == src[i]:
                blocked[i] = True
    # same-destination: lowest handle wins
    entering: Dict[tuple, List[int]] = {}
    for i in range(n):
        if wants[i]:
            entering.setdefault(dst[i], []).append(i)
    for cell, lst in entering.items():
        if len(lst) > 1:
            # flatland skips the vote when the contended cell's own occupant is already blocked
            # (its followers are blocked through the chain rule below anyway)
            for i in lst[1:]:
                blocked[i] = True
    # chains: an agent whose destination is occupied by a blocked agent is blocked (fixed point)
    changed = True
    while changed:
        changed = False
        for i in range(n):
            if not blocked[i]:
                j = occupant.get(dst[i])
                if j is not None and j != i and blocked[j]:
                    blocked[i] = True
                    changed = True
    return [wants[i] and not blocked[i] for i in range(n)]


# --------------------------------------------------------------------------- F2/F4: RailEnv
class RailEnv:
    """Fixture-backed stand-in for flatland's RailEnv: the map, line and timetable come from a
    fixture dict (see ``oracle/fixtures.py``) instead of the sparse generators, which cannot be
    reproduced without the upstream source."""

    def __init__(self, fixture: dict, malfunction_generator: Optional[ParamMalfunctionGen] = None,
                 width=None, height=None, rail_generator=None, line_generator=None, number_of_agents=None):
        self.fixture = fixture
        self.rail = RailGridTransitionMap(fixture["grid"])
        self.height, self.width = self.rail.height, self.rail.width
        if malfunction_generator is None:
            malfunction_generator = ParamMalfunctionGen(MalfunctionParameters(
                float(fixture.get("malfunction_rate", 0.0)), int(fixture.get("min_duration", 0)),
                int(fixture.get("max_duration", 0))))
        self.malfunction_generator = malfunction_generator
        self._max_episode_steps = int(fixture["max_episode_steps"])
        self.remove_agents_at_target = True
        self.np_random = np.random.RandomState()
        self.agents: List[EnvAgent] = []
        self.distance_map = DistanceMap(self.agents, self.height, self.width)
        self._elapsed_steps = 0
        self.dones = {"__all__": False}
        self.malfunction_events: List[Tuple[int, int, int]] = []  # (tick, handle, duration) for the replay harness
        self.injected_malfunctions: Optional[Dict[Tuple[int, int], int]] = None
        self._make_agents()

    def _make_agents(self):
        f = self.fixture
        self.agents = [EnvAgent(i, f["init_pos"][i], f["init_dir"][i], f["target"][i],
                                f["earliest_departure"][i], f["latest_arrival"][i])
                       for i in range(len(f["init_dir"]))]

    def get_num_agents(self):
        return len(self.agents)

    def reset(self, regenerate_rail=True, regenerate_schedule=True, random_seed=None):
        if random_seed is not None:
            self.np_random = np.random.RandomState(random_seed)
        self._make_agents()
        self.distance_map.reset(self.agents, self.rail)
        self._elapsed_steps = 0
        self.dones = {i: False for i in range(len(self.agents))}
        self.dones["__all__"] = False
        self.malfunction_events = []
        return None, self.get_info_dict()

    def get_info_dict(self):
        return {
            "malfunction": {i: a.malfunction_handler.malfunction_down_counter for i, a in enumerate(self.agents)},
            "state": {i: a.state for i, a in enumerate(self.agents)},
        }

    # -- action preprocessing (Appendix B step 2)
    def _preprocess_action(self, action, agent: EnvAgent):
        action = RailEnvActions(action)
        if action == RailEnvActions.DO_NOTHING:
            if agent.state == TrainState.MOVING:
                action = RailEnvActions.MOVE_FORWARD
            elif agent.saved_action is not None:
                action = agent.saved_action
            else:
                action = RailEnvActions.STOP_MOVING
        if agent.state == TrainState.WAITING:
            action = RailEnvActions.DO_NOTHING
        pos, d = agent.position, agent.direction
        if pos is None:
            pos, d = agent.initial_position, agent.initial_direction
        if action in (RailEnvActions.MOVE_LEFT, RailEnvActions.MOVE_RIGHT):
            cell_ok, _, trans_ok, _ = self.rail.check_action_on_agent(action, (pos, d))
            if not (cell_ok and trans_ok):
                action = RailEnvActions.MOVE_FORWARD
        if action.is_moving_action():
            cell_ok, _, trans_ok, _ = self.rail.check_action_on_agent(action, (pos, d))
            if not (cell_ok and trans_ok):
                action = RailEnvActions.STOP_MOVING
        return action

    def step(self, action_dict: Dict[int, RailEnvActions]):
        """F2-F5, speed 1.0 for every train (SURVEY Appendix B 'RailEnv.step')."""
        self._elapsed_steps += 1
        agents = self.agents
        n = len(agents)
        olds: List[Optional[tuple]] = [None] * n
        news: List[Optional[tuple]] = [None] * n
        new_dirs = [0] * n
        pre = [RailEnvActions.DO_NOTHING] * n
        for i, agent in enumerate(agents):
            agent.old_position = agent.position
            agent.old_direction = agent.direction
            # F5: draw for every agent every tick; only applied when the counter is 0
            if self.injected_malfunctions is not None:
                dur = self.injected_malfunctions.get((self._elapsed_steps, i), 0)
            else:
                dur = self.malfunction_generator.generate(self.np_random)
            mh = agent.malfunction_handler
            if mh.malfunction_down_counter == 0 and dur > 0:
                mh.malfunction_down_counter = dur
                mh.num_malfunctions += 1
                self.malfunction_events.append((self._elapsed_steps, i, dur))
            action = self._preprocess_action(action_dict.get(i, RailEnvActions.DO_NOTHING), agent)
            if action.is_moving_action() and agent.saved_action is None and agent.state != TrainState.DONE:
                agent.saved_action = action
            update_allowed = (not mh.in_malfunction) and action != RailEnvActions.STOP_MOVING
            if agent.state == TrainState.DONE:
                npos, ndir = agent.position, agent.direction
            elif agent.position is None and agent.saved_action is not None:
                npos, ndir = agent.initial_position, agent.initial_direction
            elif agent.saved_action is not None and update_allowed:
                cell_ok, (p2, d2), trans_ok, _ = self.rail.check_action_on_agent(
                    agent.saved_action, (agent.position, agent.direction))
                if cell_ok and trans_ok:
                    npos, ndir = p2, d2
                else:
                    npos, ndir = agent.position, agent.direction
                action = agent.saved_action
            else:
                npos, ndir = agent.position, agent.direction
            olds[i], news[i], new_dirs[i], pre[i] = agent.position, npos, ndir, action

        motion_ok = resolve_motion(olds, news)

        all_done = True
        for i, agent in enumerate(agents):
            mh = agent.malfunction_handler
            movement_allowed = (not mh.in_malfunction) and motion_ok[i]
            action = pre[i]
            in_malf = mh.in_malfunction
            counter_complete = mh.malfunction_down_counter == 0
            ed_reached = self._elapsed_steps >= agent.earliest_departure
            stop_given = action == RailEnvActions.STOP_MOVING
            valid_move = action.is_moving_action() and movement_allowed
            conflict = not movement_allowed
            st = agent.state
            nxt = st
            if st == TrainState.WAITING:
                if in_malf:
                    nxt = TrainState.MALFUNCTION_OFF_MAP
                elif ed_reached:
                    nxt = TrainState.READY_TO_DEPART
            elif st == TrainState.READY_TO_DEPART:
                if in_malf:
                    nxt = TrainState.MALFUNCTION_OFF_MAP
                elif valid_move:
                    nxt = TrainState.MOVING
            elif st == TrainState.MALFUNCTION_OFF_MAP:
                # MALF_OFF_MAP -> (done & ed & valid) MOVING | (done & ed) STOPPED |
                # (done & not ed) WAITING.  A train whose entry cell is occupied when its off-map malfunction ends
                # therefore goes STOPPED -- an on-map state, so it is placed on its initial cell (position update below)
                if counter_complete:
                    if ed_reached:
                        nxt = TrainState.MOVING if valid_move else TrainState.STOPPED
                    else:
                        nxt = TrainState.WAITING
            elif st == TrainState.MOVING:
                if in_malf:
                    nxt = TrainState.MALFUNCTION
                elif stop_given or conflict:
                    nxt = TrainState.STOPPED
            elif st == TrainState.STOPPED:
                if in_malf:
                    nxt = TrainState.MALFUNCTION
                elif valid_move:
                    nxt = TrainState.MOVING
            elif st == TrainState.MALFUNCTION:
                if counter_complete and valid_move:
                    nxt = TrainState.MOVING
                elif counter_complete and (stop_given or conflict):
                    nxt = TrainState.STOPPED
            agent.state_machine.previous_state = st
            agent.state_machine.state = nxt

            movement_allowed = movement_allowed and nxt != TrainState.DONE
            if nxt.is_on_map_state():
                if st.is_off_map_state():
                    agent.position = agent.initial_position
                    agent.direction = agent.initial_direction
                elif movement_allowed:
                    agent.position = news[i]
                    agent.direction = new_dirs[i]
                    if agent.position == agent.target:
                        agent.state_machine.state = TrainState.DONE
            if agent.state == TrainState.DONE and agent.arrival_time is None:
                agent.arrival_time = self._elapsed_steps
                self.dones[i] = True
                if self.remove_agents_at_target:
                    agent.position = None
            all_done &= agent.state == TrainState.DONE
            if mh.malfunction_down_counter > 0:
                mh.malfunction_down_counter -= 1
            if agent.position is not None:
                agent.saved_action = None

        if all_done or self._elapsed_steps >= self._max_episode_steps:
            for i in range(n):
                self.dones[i] = True
            self.dones["__all__"] = True
        return None, None, dict(self.dones), self.get_info_dict()
