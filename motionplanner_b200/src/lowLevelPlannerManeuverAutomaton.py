"""Maneuver-automaton trajectory planner with the collision checks of each expansion batched on the GPU.

Drop-in for reference src/lowLevelPlannerManeuverAutomaton.py:
    lowLevelPlannerManeuverAutomaton(scenario, planning_problem, param, space_all, ref_traj, MA, search_horizon=2)
        -> (x, u, controller) | (None, None, None)                                                   (:19-92)
    collision_check(node, primitive, space, param, timepoint) -> bool                                (:95-125)
    goal_check / extract_control_inputs / transform_trajectory / expand_node / Node                  (:127-225)

Search semantics kept: best-first on accumulated squared distance to the reference trajectory, stable sort of the
queue at every iteration, lazy *use* of the collision decision when a node is popped, the "fix the first primitive
after `search_horizon` levels" heuristic (+1000 on disagreeing queue entries), goal test on the vehicle centre.
What changed is *when* the geometry is evaluated: when a node is expanded, all its children (one velocity bucket,
60-195 primitives x <= 11 occupancies) are decided by a single mpc_query and each child carries its bit; popping a
child then only applies the time guard of the reference (:99-109) and reads the bit.
"""
import operator

import numpy as np

from ..checker import PrimitiveChecker, SceneSet
from ..geometry import Point
from ..maneuverAutomaton.Controller import FeedbackController
from ..maneuverAutomaton.ManeuverAutomaton import automaton_of


_COST = operator.attrgetter('cost')      # stable sort key of the search queue (reference :54)


class Node:
    """search node: sequence of primitive indices, rear-axle trajectory 4 x (10 k + 1), accumulated cost.
    Children created by the batched expansion keep (parent, own segment) and build ``x`` on first use -- only nodes
    that are actually popped ever need the full trajectory."""
    __slots__ = ('primitives', '_x', 'cost', 'collides', '_parent', '_seg', 'goal_step')

    def __init__(self, primitives, x, cost, collides=None, parent=None, seg=None, goal_step=None):
        self.primitives = primitives
        self._x = x
        self.cost = cost
        self.collides = collides          # geometry bit of the last primitive (None: not evaluated yet)
        self._parent, self._seg = parent, seg
        self.goal_step = goal_step        # device-side goal test: step index, -1 = not reached, None = not evaluated

    @property
    def x(self):
        if self._x is None:
            seg = self._seg
            if not isinstance(seg, np.ndarray):        # device expansion: `seg` is the primitive, transform on first use
                px = self._parent.x
                seg = _rotation(px[3, -1]) @ seg.x + np.array([[px[0, -1]], [px[1, -1]], [0], [px[3, -1]]])
            self._x = np.concatenate((self._parent.x[:, :-1], seg), axis=1)
            self._parent = self._seg = None
        return self._x

    @x.setter
    def x(self, value):
        self._x = value


def _rotation(phi):
    """block_diag([[cos, -sin], [sin, cos]], eye(2)) of the reference (:44, :75), built directly (the scipy call costs
    100 us per expansion)"""
    c, s = np.cos(phi), np.sin(phi)
    return np.array([[c, -s, 0.0, 0.0], [s, c, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


def expand_node(node, primitive, ind, ref_traj, fixed, T, collides=None):
    """child of `node` that appends primitive `ind`: trajectory, cost against the reference trajectory (:192-214)"""
    primitives = node.primitives + [ind]
    seg = T @ primitive.x + np.array([[node.x[0, -1]], [node.x[1, -1]], [0], [node.x[3, -1]]])
    x = np.concatenate((node.x[:, :-1], seg), axis=1)
    first = x.shape[1] - primitive.x.shape[1]
    last = min(x.shape[1], ref_traj.shape[1])
    cols = range(first, last)
    cost = node.cost + np.sum((ref_traj[0:2, cols] - x[0:2, cols]) ** 2)
    if len(primitives) <= len(fixed) and primitives[-1] != fixed[len(primitives) - 1]:
        cost = cost + 1000
    return Node(primitives, x, cost, collides)


def _time_exceeded(ind, goals):
    return all(ind > g['time_end'] for g in goals)


def collision_check(node, primitive, space, param, timepoint, checker=None):
    """check if a motion primitive is collision free (True = free) -- the reference's call (:95-125) with the
    reference's arguments: `node` needs `.x` only (a reference `Node` works; nodes of the batched search also carry
    the bit decided at expansion), `primitive` is a primitive of a ManeuverAutomaton of this package, `space` the
    list of free-space geometries, `param` the reference's dictionary ('goal', 'steps', 'time_step'; nothing is
    added to it).  `checker` (extension) reuses staged data; without it the automaton is found through the primitive
    and (automaton, space) are staged once and cached."""
    ind = node.x.shape[1] - primitive.x.shape[1]
    if _time_exceeded(ind, param['goal']):
        return False
    bit = getattr(node, 'collides', None)
    if bit is not None:
        return not bit
    if checker is None:
        MA, idx = automaton_of(primitive)
        checker = _checker_for(MA, param['time_step'], space, timepoint)
    else:
        idx = checker.index_of[id(primitive)]
    x = node.x[:, ind]
    scenes = (0,) if timepoint else (0, 1)
    return not bool(checker.list_collides(x[0], x[1], x[3], ind, param['steps'], [idx], scenes)[0])


_CACHE = {}


def _checker_for(MA, time_step, space, timepoint, backend=None):
    """stage automaton + free space once per (automaton, space list) pair.  The key holds the identity of every
    geometry of the list (a list mutated in place is a different free space); the checker itself re-stages its scenes
    when another user of the engine has staged in between (PrimitiveChecker._ensure_staged)."""
    key = (id(MA), float(time_step), id(space), tuple(map(id, space)), bool(timepoint), id(backend))
    hit = _CACHE.get('k')
    if hit is not None and hit[0] == key and hit[1] is space:
        return hit[2]
    chk = PrimitiveChecker(MA, time_step, backend)
    scenes = SceneSet()
    if timepoint:
        scenes.add_scene_from_spaces(space)
    else:
        # interval occupancies: the reference intersects consecutive free spaces (:27-30); a convex set lies in
        # A n B iff it lies in A and in B, so the occupancy is tested against both steps instead
        scenes.add_scene_from_spaces(space[:-1])
        scenes.add_scene_from_spaces(space[1:])
    chk.set_scenes(scenes)
    _CACHE['k'] = (key, space, chk, list(space))        # the last entry keeps the geometries (their ids) alive
    return chk


def goal_check(node, primitive, param):
    """first step of the last primitive whose vehicle centre lies in a goal set inside its time window (:127-142)"""
    if getattr(node, 'goal_step', None) is not None:          # decided on the device together with the collision bits
        return (True, int(node.goal_step)) if node.goal_step >= 0 else (False, None)
    ind = node.x.shape[1] - primitive.x.shape[1]
    for i in range(ind, node.x.shape[1]):
        for goal in param['goal']:
            if goal['time_start'] <= i <= goal['time_end']:
                p = Point(node.x[0, i] + np.cos(node.x[3, i]) * param['b'],
                          node.x[1, i] + np.sin(node.x[3, i]) * param['b'])
                if goal['space'] is None or goal['space'].contains(p):
                    return True, i
    return False, None


def extract_control_inputs(node, primitives):
    """piecewise-constant input sequence of the node's primitives (:168-181)"""
    parts = [np.expand_dims(primitives[i].u, axis=1) @ np.ones((1, primitives[i].x.shape[1] - 1))
             for i in node.primitives]
    return np.concatenate(parts, axis=1) if parts else []


def transform_trajectory(x, param):
    """rear axle -> vehicle centre, in place (:183-190)"""
    for i in range(x.shape[1]):
        x[0, i] = x[0, i] + np.cos(x[3, i]) * param['b']
        x[1, i] = x[1, i] + np.sin(x[3, i]) * param['b']
    return x


def _children(checker, node, MA, bucket, ref_traj, fixed, param, scenes):
    """expand `node` by every primitive of velocity bucket `bucket`: one device query decides all of them, one batched
    product builds all child trajectories.  Same arithmetic per child as expand_node (reference :192-214)."""
    x = node.x
    ind = x.shape[1] - 1
    cand = MA.vel_dic[bucket]
    if not cand:
        return []
    depth = len(node.primitives) + 1
    if getattr(checker, 'device_expand', False):
        # bits, costs and goal steps of all children from one mpc_expand call; trajectories are built when popped
        bits, costs, goal = checker.bucket_expand(x[0, -1], x[1, -1], x[3, -1], node.cost, ind, param['steps'], bucket, scenes)
        if _time_exceeded(ind, param['goal']):
            bits = np.ones(len(cand), bool)
        prims = MA.primitives
        fix = fixed[depth - 1] if depth <= len(fixed) else None
        base = node.primitives
        return [Node(base + [i], None, (c + 1000 if fix is not None and i != fix else c), b, node, prims[i], g)
                for i, c, b, g in zip(cand, costs.tolist(), bits.tolist(), goal.tolist())]
    if _time_exceeded(ind, param['goal']):
        bits = np.ones(len(cand), bool)           # the reference rejects these before looking at geometry (:99-109)
    else:
        bits = checker.bucket_collides(x[0, -1], x[1, -1], x[3, -1], ind, param['steps'], bucket, scenes)
    T = _rotation(x[3, -1])
    off = np.array([[x[0, -1]], [x[1, -1]], [0], [x[3, -1]]])
    costs = np.empty(len(cand))
    segs = [None] * len(cand)
    for loc, PX in MA.bucket_stacks(bucket):
        X = T @ PX + off                                           # [m, 4, L]
        last = min(ind + PX.shape[2], ref_traj.shape[1])
        if last > ind:
            d = ref_traj[0:2, ind:last][None] - X[:, 0:2, :last - ind]
            # the reference sums `(ref_traj[0:2, index] - x[0:2, index])**2`, whose fancy-indexed operands are
            # Fortran-ordered: numpy adds x0, y0, x1, y1, ... -- reproduce that order so costs match to the last bit
            sq = np.ascontiguousarray(np.swapaxes(d ** 2, 1, 2))
            costs[loc] = node.cost + sq.reshape(len(loc), -1).sum(axis=1)
        else:
            costs[loc] = node.cost + 0.0
        for k, i in enumerate(loc):
            segs[i] = X[k]
    out = []
    for k, i in enumerate(cand):
        cost = costs[k]
        if depth <= len(fixed) and i != fixed[depth - 1]:
            cost = cost + 1000
        out.append(Node(node.primitives + [i], None, cost, bool(bits[k]), node, segs[k]))
    return out


def lowLevelPlannerManeuverAutomaton(scenario, planning_problem, param, space_all, ref_traj, MA, search_horizon=2,
                                     backend=None, device_expand=False):
    """plan a concrete trajectory for the given high-level plan using a maneuver automaton.

    device_expand=True moves the rest of the primitive-selection step to the GPU as well (mpc_expand: child costs and
    goal tests next to the collision bits); the default keeps them in numpy."""
    timepoint = MA.is_timepoint()
    checker = _checker_for(MA, param['time_step'], space_all, timepoint, backend)
    checker.device_expand = bool(device_expand)
    if device_expand:
        checker.enable_expand(ref_traj, param['goal'], param['b'])
    scenes = (0,) if timepoint else (0, 1)

    # initial state, shifted from the vehicle centre to the rear axle (:33-36)
    x = np.expand_dims(np.array([param['x0'][0], param['x0'][1], param['v_init'], param['orientation']]), 1)
    x[0, 0] = x[0, 0] - np.cos(x[3, 0]) * param['b']
    x[1, 0] = x[1, 0] - np.sin(x[3, 0]) * param['b']

    node = Node([], x, 0)
    cnt, fixed = 0, []
    queue = _children(checker, node, MA, int(MA._bucket(param['v_init'])), ref_traj, fixed, param, scenes)

    while len(queue) > 0:
        queue.sort(key=_COST)
        node = queue.pop(0)
        primitive = MA.primitives[node.primitives[-1]]

        if collision_check(node, primitive, space_all, param, timepoint, checker):

            # fix the first not yet fixed primitive once the search is `search_horizon` levels ahead of it (:64-69)
            if len(node.primitives) == cnt + search_horizon:
                for other in queue:
                    if cnt < len(other.primitives) and other.primitives[cnt] != node.primitives[cnt]:
                        other.cost = other.cost + 1000
                fixed.append(node.primitives[cnt])
                cnt = cnt + 1

            res, ind = goal_check(node, primitive, param)
            if res:
                if timepoint:
                    u = extract_control_inputs(node, MA.primitives)
                    t = param['time_step'] * np.arange(0, x.shape[1] + 1)
                    K = [np.zeros([u.shape[0], x.shape[0]]) for _ in range(u.shape[1])]
                    controller = FeedbackController(x, u, t, K)
                    return transform_trajectory(node.x[:, :ind + 1], param), u[:, :ind], controller
                raise NotImplementedError("interval-occupancy automata need the AROC controllers "
                                          "(reference loadAROCcontroller, out of scope: SURVEY.md section 2 #8)")

            queue.extend(_children(checker, node, MA, int(MA._bucket(primitive.x[2, -1])), ref_traj, fixed, param,
                                   scenes))

    return None, None, None
