"""Stand-alone maneuver-automaton planner (no decision module) with GPU-batched collision checks.

Drop-in for reference src/maneuverAutomatonPlannerStandalone.py:
    maneuverAutomatonPlannerStandalone(scenario, planning_problem, param, MA) -> (x, u, controller) | (None,)*3   (:19-96)
    get_goal_set (:98-149), free_space (:151-182), collision_check (:196-226), goal_check (:228-243), expand_node (:293-310)

The reference builds ``space[t] = road(safe_only, x0) - U obstacles_t`` with shapely ``difference`` for every time step
and then asks ``space[t].contains(occupancy)``.  Here the free space is kept in its decomposed form -- the road
boundary as segments plus the obstacle pieces of every step -- which is what the kernel evaluates
(DESIGN.md section 2); ``free_space`` therefore returns a ``FreeSpace`` handle instead of a list of polygons.
"""
from copy import deepcopy

import heapq

import numpy as np

from ..auxiliary.polygonLaneletNetwork import polygonLaneletNetwork
from ..checker import PrimitiveChecker, SceneSet, scenario_obstacle_arrays
from ..geometry import Point
from ..maneuverAutomaton.Controller import FeedbackController
from ..maneuverAutomaton.ManeuverAutomaton import automaton_of
from .lowLevelPlannerManeuverAutomaton import (Node, _rotation, extract_control_inputs, transform_trajectory)


class FreeSpace:
    """road boundary + per-step obstacle pieces; ``len()`` = number of time steps like the reference's list"""

    def __init__(self, road, step_obst_off, obst, obst_nv, origin):
        self.road, self.step_obst_off, self.obst, self.obst_nv, self.origin = road, step_obst_off, obst, obst_nv, origin

    def __len__(self):
        return len(self.step_obst_off) - 1

    def scene_set(self, drop_last=0, shift=0):
        n = len(self) - drop_last
        off = self.step_obst_off[shift:shift + n + 1] - self.step_obst_off[shift]
        lo, hi = self.step_obst_off[shift], self.step_obst_off[shift + n]
        return n, off, self.obst[lo:hi], self.obst_nv[lo:hi]


def _shape_geometry(shape):
    return shape.shapely_object


def get_goal_set(scenario, planning_problem):
    """list of goal dictionaries {'time_start','time_end','space','lane'} (reference :98-149)"""
    goal = []
    lanelets = {l.lanelet_id: l for l in scenario.lanelet_network.lanelets}
    t_init = planning_problem.initial_state.time_step
    of_goal = planning_problem.goal.lanelets_of_goal_position
    for i, gs in enumerate(planning_problem.goal.state_list):
        pos = getattr(gs, 'position', None)
        if pos is not None and hasattr(pos, 'shapes'):
            shapes = len(pos.shapes)
        elif of_goal is not None:
            shapes = len(list(of_goal.values())[i])
        else:
            shapes = 1
        for j in range(shapes):
            g = {'time_start': gs.time_step.start - t_init, 'time_end': gs.time_step.end - t_init}
            if pos is not None:
                g['space'] = _shape_geometry(pos.shapes[j] if hasattr(pos, 'shapes') else pos)
                if of_goal is None:
                    for l in scenario.lanelet_network.lanelets:
                        if _rings_touch(l.polygon.vertices, g['space']):
                            g['lane'] = l.lanelet_id
                            goal.append(deepcopy(g))
                else:
                    g['lane'] = list(of_goal.values())[i][j]
                    goal.append(deepcopy(g))
            else:
                if of_goal is None:
                    g['lane'], g['space'] = None, None
                else:
                    g['lane'] = list(of_goal.values())[i][j]
                    g['space'] = _shape_geometry(lanelets[g['lane']].polygon)
                goal.append(deepcopy(g))
    return goal


def _rings_touch(vertices, geom):
    """closed polygons share a point (host-side, used only to label goals with lanelets, reference :129-132)"""
    from ..geometry import point_in_rings, rings_of
    a = [tuple(map(float, p)) for p in vertices]
    if a[0] != a[-1]:
        a.append(a[0])
    for ring in rings_of(geom):
        if any(point_in_rings(p, [a]) >= 0 for p in ring) or any(point_in_rings(p, [ring]) >= 0 for p in a):
            return True
        for p0, p1 in zip(a[:-1], a[1:]):
            for q0, q1 in zip(ring[:-1], ring[1:]):
                d1 = (p1[0] - p0[0]) * (q0[1] - p0[1]) - (p1[1] - p0[1]) * (q0[0] - p0[0])
                d2 = (p1[0] - p0[0]) * (q1[1] - p0[1]) - (p1[1] - p0[1]) * (q1[0] - p0[0])
                d3 = (q1[0] - q0[0]) * (p0[1] - q0[1]) - (q1[1] - q0[1]) * (p0[0] - q0[0])
                d4 = (q1[0] - q0[0]) * (p1[1] - q0[1]) - (q1[1] - q0[1]) * (p1[0] - q0[0])
                if d1 * d2 < 0 and d3 * d4 < 0:
                    return True
    return False


def free_space(scenario, goal_set, x0):
    """free space for each time step 0..max goal time_end (reference :151-182), decomposed"""
    road = polygonLaneletNetwork(scenario, safe_only=True, x0=x0)
    steps = 0
    for goal in goal_set:
        steps = np.maximum(steps, goal['time_end'])
    off, obst, nv = scenario_obstacle_arrays(scenario, int(steps) + 1)
    return FreeSpace(road, off, obst, nv, np.asarray(x0[0:2], dtype=np.float64))


def _stage(MA, space, param, backend=None):
    chk = PrimitiveChecker(MA, param['time_step'], backend)
    scenes = SceneSet()
    if hasattr(space, 'scene_set'):                 # FreeSpace handle of this module's free_space()
        if param['timepoint']:
            n, off, obst, nv = space.scene_set()
            scenes.add_scene_explicit(n, off, obst, nv, space.road.segments, origin=space.origin)
        else:   # interval occupancies: inside space[i] n space[i+1]  <=>  inside both (reference :38-43)
            for shift in (0, 1):
                n, off, obst, nv = space.scene_set(drop_last=1, shift=shift)
                scenes.add_scene_explicit(n, off, obst, nv, space.road.segments, origin=space.origin)
    else:                                           # the reference's list of shapely(-like) polygons, one per step;
        scenes.add_scene_from_spaces(space)         # for interval automata the reference has intersected them already
    chk.set_scenes(scenes)
    return chk


_CACHE = {}


def _checker_for(MA, space, param):
    """single-call collision_check: automaton + free space staged once per (automaton, space) pair"""
    key = (id(MA), float(param['time_step']), id(space), bool(param['timepoint']))
    hit = _CACHE.get('k')
    if hit is not None and hit[0] == key and hit[1] is space:
        return hit[2]
    chk = _stage(MA, space, param)
    _CACHE['k'] = (key, space, chk)
    return chk


def collision_check(node, primitive, space, goal_set, param, checker=None):
    """check if a motion primitive is collision free (reference :196-226), called with the reference's arguments:
    `node` needs `.x` only, `space` is the FreeSpace handle of free_space() or a list of shapely-like polygons, `param`
    carries 'timepoint', 'steps', 'time_step' as the reference's planner sets them (:34-36).  Nodes of the batched
    search carry the bit decided at expansion."""
    ind = node.x.shape[1] - primitive.x.shape[1]
    if all(ind > g['time_end'] for g in goal_set):
        return False
    bit = getattr(node, 'collides', None)
    if bit is not None:
        return not bit
    if checker is None:
        MA, idx = automaton_of(primitive)
        checker = _checker_for(MA, space, param)
    else:
        idx = checker.index_of[id(primitive)]
    x = node.x[:, ind]
    # a FreeSpace staged for an interval automaton is two shifted scenes; a list of polygons is used as it is
    scenes = (0,) if param['timepoint'] or not hasattr(space, 'scene_set') else (0, 1)
    return not bool(checker.list_collides(x[0], x[1], x[3], ind, param['steps'], [idx], scenes)[0])


def goal_check(node, primitive, goal_set, param):
    """reference :228-243"""
    ind = node.x.shape[1] - primitive.x.shape[1]
    for i in range(ind, node.x.shape[1]):
        for goal in goal_set:
            if goal['time_start'] <= i <= goal['time_end']:
                p = Point(node.x[0, i] + np.cos(node.x[3, i]) * param['b'], node.x[1, i] + np.sin(node.x[3, i]) * param['b'])
                if goal['space'] is None or goal['space'].contains(p):
                    return True, i
    return False, None


def expand_node(node, primitive, ind, T, goal_set, collides=None):
    """child node; cost = smallest distance of the new states to a goal boundary (reference :293-310)"""
    primitives = node.primitives + [ind]
    seg = T @ primitive.x + np.array([[node.x[0, -1]], [node.x[1, -1]], [0], [node.x[3, -1]]])
    x = np.concatenate((node.x[:, :-1], seg), axis=1)
    cost = np.inf
    for j in range(node.x.shape[1] + 1, x.shape[1]):
        for goal in goal_set:
            if goal['space'] is not None:
                cost = min(cost, goal['space'].exterior.distance(Point(x[0, j], x[1, j])))
    return Node(primitives, x, cost, collides)


def _goal_rings(goal_set):
    """(a, b - a, max(|b - a|^2, 1e-300)) edge arrays of every goal's exterior ring, as `_Ring.distance` uses them"""
    out = []
    for goal in goal_set:
        if goal['space'] is not None:
            c = np.asarray(goal['space'].exterior.coords, dtype=np.float64).reshape(-1, 2)
            a, d = c[:-1], c[1:] - c[:-1]
            out.append((a, d, np.maximum(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1], 1e-300)))
    return out


def _children(checker, node, MA, bucket, goal_set, param, rings=None):
    """all children of `node` (one velocity bucket): one device query for the collision bits, one batched product for
    the trajectories, one vectorised pass for the costs.  Per child the arithmetic is that of expand_node (reference
    :293-310) -- tests/test_planner.py asserts bit equality -- and trajectories are assembled when a child is popped."""
    x = node.x
    ind = x.shape[1] - 1
    cand = MA.vel_dic[bucket]
    if not cand:
        return []
    scenes = (0,) if param['timepoint'] else (0, 1)
    if all(ind > g['time_end'] for g in goal_set):
        bits = np.ones(len(cand), bool)
    else:
        bits = checker.bucket_collides(x[0, -1], x[1, -1], x[3, -1], ind, param['steps'], bucket, scenes)
    if rings is None:
        rings = _goal_rings(goal_set)
    T = _rotation(x[3, -1])
    off = np.array([[x[0, -1]], [x[1, -1]], [0], [x[3, -1]]])
    costs = np.full(len(cand), np.inf)
    segs = [None] * len(cand)
    for loc, PX in MA.bucket_stacks(bucket):
        X = T @ PX + off                                           # [m, 4, L]
        # the reference's loop starts at node.x.shape[1] + 1, i.e. at state 2 of the appended primitive (:304)
        px, py = X[:, 0, 2:, None], X[:, 1, 2:, None]              # [m, L-2, 1]
        if px.shape[1]:
            for a, d, l2 in rings:
                t = np.clip(((px - a[:, 0]) * d[:, 0] + (py - a[:, 1]) * d[:, 1]) / l2, 0.0, 1.0)
                qx, qy = a[:, 0] + t * d[:, 0], a[:, 1] + t * d[:, 1]
                dist = np.sqrt((qx - px) ** 2 + (qy - py) ** 2)
                costs[loc] = np.minimum(costs[loc], dist.reshape(len(loc), -1).min(axis=1))
        for k, i in enumerate(loc):
            segs[i] = X[k]
    base = node.primitives
    return [Node(base + [i], None, c, b, node, seg)
            for i, c, b, seg in zip(cand, costs.tolist(), bits.tolist(), segs)]


def maneuverAutomatonPlannerStandalone(scenario, planning_problem, param, MA, backend=None, max_nodes=None):
    """solve the motion planning problem with a maneuver automaton.  ``max_nodes`` (extension) bounds the number of
    popped nodes; the reference searches until the queue is empty."""
    planning_problem = list(planning_problem.planning_problem_dict.values())[0]
    s0 = planning_problem.initial_state
    x0 = np.array([s0.position[0], s0.position[1], s0.velocity, s0.orientation])
    goal_set = get_goal_set(scenario, planning_problem)
    space = free_space(scenario, goal_set, x0)

    param['timepoint'] = MA.is_timepoint()
    param['steps'] = len(space)
    param['time_step'] = scenario.dt
    checker = _stage(MA, space, param, backend)

    x = np.expand_dims(x0, axis=1)
    x[0, 0] = x[0, 0] - np.cos(x[3, 0]) * param['b']
    x[1, 0] = x[1, 0] - np.sin(x[3, 0]) * param['b']

    # The reference keeps a list, sorts it (stably) by cost at every iteration and pops the head (:62-64).  Costs are
    # never changed after insertion here, so the list is always ordered by (cost, insertion number) and its head is the
    # minimum of that pair: a binary heap on (cost, insertion number) pops the very same nodes in the same order
    # (tests/test_planner.py::test_heap_pops_like_stable_sort) without re-sorting 10^5 entries per popped node.
    node = Node([], x, 0)
    rings = _goal_rings(goal_set)
    seq = 0
    queue = []
    for child in _children(checker, node, MA, int(MA._bucket(x0[2])), goal_set, param, rings):
        queue.append((child.cost, seq, child))
        seq += 1
    heapq.heapify(queue)
    popped = 0
    while len(queue) > 0:
        node = heapq.heappop(queue)[2]
        popped += 1
        if max_nodes is not None and popped > max_nodes:
            break
        primitive = MA.primitives[node.primitives[-1]]
        if collision_check(node, primitive, space, goal_set, param, checker):
            res, ind = goal_check(node, primitive, goal_set, param)
            if res:
                if not param['timepoint']:
                    raise NotImplementedError("interval-occupancy automata need the AROC controllers (out of scope)")
                u = extract_control_inputs(node, MA.primitives)
                t = param['time_step'] * np.arange(0, x.shape[1] + 1)
                K = [np.zeros([u.shape[0], x.shape[0]]) for _ in range(u.shape[1])]
                controller = FeedbackController(x, u, t, K)
                return transform_trajectory(node.x[:, :ind + 1], param), u[:, :ind], controller
            for child in _children(checker, node, MA, int(MA._bucket(primitive.x[2, -1])), goal_set, param, rings):
                heapq.heappush(queue, (child.cost, seq, child))
                seq += 1
    return None, None, None
