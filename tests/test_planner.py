"""The primitive-selection step (reference src/lowLevelPlannerManeuverAutomaton.py:19-92) behind its unchanged
signature: corridor given as one free-space polygon per time step (with a hole where an obstacle was subtracted),
reference trajectory, the regenerated automaton.  CPU: the oracle stands in for the engine; GPU: the product engine
must produce the identical plan, decision by decision."""
import os

import numpy as np
import pytest

from helpers import OracleBackend
from motionplanner_b200.geometry import Polygon
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from motionplanner_b200.src import lowLevelPlannerManeuverAutomaton as LL
from motionplanner_b200.vehicle.vehicleParameter import vehicleParameter
from oracle import exact as ex


@pytest.fixture(scope="module")
def MA():
    return load_maneuver_automaton()


def corridor_problem(v=10.0, steps=40, blocked=True):
    param = vehicleParameter()
    param.update({'time_step': 0.1, 'steps': steps, 'x0': np.array([0.0, 0.0]), 'v_init': v, 'orientation': 0.0,
                  'goal': [{'time_start': steps - 10, 'time_end': steps, 'space': None}]})
    shell = [(-10.0, -1.75), (200.0, -1.75), (200.0, 5.25), (-10.0, 5.25)]          # two lanes
    space = []
    for t in range(steps + 1):
        if blocked and 8 <= t <= 30:
            x0 = 22.0 + 2.0 * 0.1 * t                                                # slow car ahead in the ego lane
            hole = [(x0, -1.0), (x0, 1.0), (x0 + 4.5, 1.0), (x0 + 4.5, -1.0)]
            space.append(Polygon(shell, [hole]))
        else:
            space.append(Polygon(shell))
    t = np.arange(steps + 1) * 0.1
    ref = np.stack([v * t, np.zeros_like(t), np.full_like(t, v), np.zeros_like(t)])
    if blocked:                                                                       # reference line moves to lane 2
        ref[1] = np.clip((t - 0.5) * 3.5 / 1.0, 0.0, 3.5)
    return param, space, ref


def plan_is_inside(x_centre, space, param):
    """every ego rectangle of the returned centre-frame trajectory lies in the free space of its step (exact)"""
    for i in range(x_centre.shape[1]):
        car = ex.rectangle_vertices(param['length'], param['width'], x_centre[0, i], x_centre[1, i],
                                    np.cos(x_centre[3, i]), np.sin(x_centre[3, i]))
        rings = [list(space[i].exterior.coords)] + [list(h.coords) for h in space[i].interiors]
        if not ex.region_contains_convex(rings, car):
            return False
    return True


def run(MA, backend, **kw):
    param, space, ref = corridor_problem(**kw)
    LL._CACHE.clear()
    x, u, ctrl = LL.lowLevelPlannerManeuverAutomaton(None, None, param, space, ref, MA, backend=backend)
    return x, u, ctrl, param, space


def test_planner_cpu_free_corridor(MA):
    x, u, ctrl, param, space = run(MA, OracleBackend(), blocked=False)
    assert x is not None and x.shape[0] == 4 and u.shape == (2, x.shape[1] - 1)
    assert param['goal'][0]['time_start'] <= x.shape[1] - 1 <= param['goal'][0]['time_end']
    assert np.abs(x[1]).max() < 0.5                       # stays on the reference line
    assert plan_is_inside(x, space, param)
    # the reference hands the *initial state* to FeedbackController (src/lowLevelPlannerManeuverAutomaton.py:77-78,
    # `x` is never reassigned); kept as is
    assert ctrl.x_ref.shape == (4, 1) and ctrl.u_ref.shape[0] == 2 and len(ctrl.K) == ctrl.u_ref.shape[1]


def test_planner_cpu_lane_change_around_hole(MA):
    be = OracleBackend()
    x, u, ctrl, param, space = run(MA, be, blocked=True)
    assert x is not None
    assert x[1].max() > 2.0                               # it moved to the second lane
    assert plan_is_inside(x, space, param)
    assert len(be.log) >= 3                               # one batched query per expansion, not one per child


def test_collision_check_single_call_matches_batched(MA):
    """module-level collision_check(node, primitive, space, param, timepoint): same decision as the batched bit"""
    param, space, ref = corridor_problem(blocked=True)
    be = OracleBackend()
    LL._CACHE.clear()
    chk = LL._checker_for(MA, 0.1, space, True, be)
    x = np.array([[0.0 - 1.15], [0.0], [10.0], [0.0]])
    root = LL.Node([], x, 0)
    bucket = int(MA._bucket(10.0))
    kids = LL._children(chk, root, MA, bucket, ref, [], param, (0,))
    for k in kids[::17]:
        prim = MA.primitives[k.primitives[-1]]
        lazy = LL.Node(k.primitives, k.x, k.cost, None)
        assert LL.collision_check(lazy, prim, space, param, True, chk) == (not k.collides)
    # the reference's own call shape: a node with `primitives, x, cost` only (reference Node :217-225), no checker,
    # nothing planted in `param` -- the automaton is found through the primitive, (automaton, space) staged once
    class RefNode:
        def __init__(self, primitives, x, cost):
            self.primitives, self.x, self.cost = primitives, x, cost

    import motionplanner_b200.checker as CK
    keys_before = set(param)
    old = CK._ENGINES.copy()
    CK._ENGINES[('planner', int(os.environ.get('MPC_DEVICE', os.environ.get('LOCAL_RANK', '0'))))] = be
    try:
        LL._CACHE.clear()
        for k in kids[::23]:
            prim = MA.primitives[k.primitives[-1]]
            assert LL.collision_check(RefNode(k.primitives, k.x, k.cost), prim, space, param, True) == (not k.collides)
        staged = LL._CACHE['k'][2]
        # another user of the same engine stages something else in between: the cached checker must notice
        other = LL._checker_for(MA, 0.1, space[:5], True, be)
        assert other is not staged and be._scene_owner is other
        k = kids[3]
        assert LL.collision_check(RefNode(k.primitives, k.x, k.cost), MA.primitives[k.primitives[-1]], space, param, True,
                                  staged) == (not k.collides)
        assert be._scene_owner is staged
    finally:
        CK._ENGINES.clear()
        CK._ENGINES.update(old)
    assert set(param) == keys_before
    # time guard (reference :99-109): a node whose parent index lies beyond every goal's time_end is rejected
    late = LL.Node([kids[0].primitives[-1]], np.zeros((4, 61)), 0.0, False)
    assert LL.collision_check(late, MA.primitives[kids[0].primitives[-1]], space, param, True, chk) is False


@pytest.mark.gpu
def test_planner_gpu_identical_plan(MA):
    from motionplanner_b200.engine import CollisionEngine
    for blocked in (False, True):
        cpu = OracleBackend()
        x0, u0, _, _, _ = run(MA, cpu, blocked=blocked)
        eng = CollisionEngine(0)
        x1, u1, _, param, space = run(MA, eng, blocked=blocked)
        eng.close()
        assert x0 is not None and x1 is not None
        assert np.array_equal(x0, x1) and np.array_equal(u0, u1)
        assert plan_is_inside(x1, space, param)


# ---------------------------------------------------------------------------------------------------------------------
# stand-alone planner (reference src/maneuverAutomatonPlannerStandalone.py) on a bundled scenario with a goal region
# ---------------------------------------------------------------------------------------------------------------------

REF_SCENARIOS = '/root/reference/scenarios'


@pytest.mark.skipif(not os.path.isdir(REF_SCENARIOS), reason="reference scenarios not mounted (build container only)")
def test_standalone_planner_cpu_on_zam_zip(MA):
    """BASELINE.json configs[0] scenario (ZAM_Zip-1_19_T-1, goal = lanelet 24 in steps 84..85): the stand-alone planner
    finds a plan with one batched query per expansion and the plan passes the trajectory validator"""
    from motionplanner_b200.auxiliary import collisionChecker as CC
    from motionplanner_b200.crlite import CommonRoadFileReader
    from motionplanner_b200.src.maneuverAutomatonPlannerStandalone import (get_goal_set,
                                                                           maneuverAutomatonPlannerStandalone)
    scenario, pps = CommonRoadFileReader(os.path.join(REF_SCENARIOS, 'ZAM_Zip-1_19_T-1.xml')).open()
    goals = get_goal_set(scenario, list(pps.planning_problem_dict.values())[0])
    assert len(goals) == 1 and goals[0]['lane'] == 24 and (goals[0]['time_start'], goals[0]['time_end']) == (84, 85)
    param = vehicleParameter()
    be = OracleBackend()
    x, u, ctrl = maneuverAutomatonPlannerStandalone(scenario, pps, param, MA, backend=be, max_nodes=20000)
    assert param['steps'] == 86 and param['timepoint'] is True
    assert x is not None and x.shape == (4, 85) and u.shape == (2, 84)
    assert 20 < len(be.log) < 200                       # batched: one query per expansion
    flags, _ = CC.collision_flags(scenario, x, param, engine=OracleBackend())
    assert not flags.any()
    assert goals[0]['space'].contains(LL.Point(x[0, -1], x[1, -1]))


def test_batched_expansion_equals_per_child_expand_node(MA):
    """the vectorised child construction reproduces expand_node (reference :192-214) bit for bit"""
    param, space, ref = corridor_problem(blocked=True)
    be = OracleBackend()
    LL._CACHE.clear()
    chk = LL._checker_for(MA, 0.1, space, True, be)
    x = np.array([[0.3 - 1.15], [0.2], [10.0], [0.07]])
    root = LL.Node([], x, 1.25)
    for bucket, fixed in ((int(MA._bucket(10.0)), []), (int(MA._bucket(0.4)), [3]), (int(MA._bucket(29.6)), [])):
        kids = LL._children(chk, root, MA, bucket, ref, fixed, param, (0,))
        T = LL._rotation(x[3, -1])
        assert len(kids) == len(MA.vel_dic[bucket])
        for k, i in zip(kids, MA.vel_dic[bucket]):
            one = LL.expand_node(root, MA.primitives[i], i, ref, fixed, T)
            assert k.primitives == one.primitives and k.cost == one.cost
            assert np.array_equal(k.x, one.x)


@pytest.mark.skipif(not os.path.isdir(REF_SCENARIOS), reason="reference scenarios not mounted (build container only)")
def test_batch_harness_writes_reference_csv_format(MA, tmp_path):
    """evaluation/runAllScenarios.py mirror: CSV rows `file, comp_time, collision, tags, final_time` that the
    reference's evaluationScript.py (evaluation_matrics :7-22) can aggregate"""
    from motionplanner_b200.auxiliary import collisionChecker as CC
    from motionplanner_b200.evaluation.runAllScenarios import run_all
    validator = lambda sc, x, prm: not CC.collision_flags(sc, x, prm, engine=OracleBackend())[0].any()
    out, data = run_all(REF_SCENARIOS, str(tmp_path), max_nodes=20000, files=['ZAM_Zip-1_19_T-1.xml', 'USA_Lanker-1_8_T-1.xml'],
                        backend=OracleBackend(), validator=validator)
    rows = [l.strip().split(',') for l in open(out)]
    assert len(rows) == 2 and all(len(r) == 5 for r in rows)
    assert rows[0][0] == 'ZAM_Zip-1_19_T-1.xml' and rows[0][2] == 'no collision' and float(rows[0][4]) == 8.5
    assert rows[0][1].replace('.', '', 1).isdigit()            # the test evaluationScript.py uses for "solved"
    assert rows[1][1] == 'failed'


def test_interval_occupancies_use_two_consecutive_free_spaces():
    """AROC-style automaton (time-interval occupancies): the reference intersects consecutive free spaces
    (src/lowLevelPlannerManeuverAutomaton.py:25-30); here the occupancy is tested against both steps.  Checked
    against exact containment in both polygons, including the shortened horizon (len(space) - 1)."""
    from motionplanner_b200.maneuverAutomaton.ManeuverAutomaton import ManeuverAutomaton, MotionPrimitive
    rng = np.random.default_rng(3)
    prims = []
    for i in range(40):
        x = np.zeros((4, 11))
        x[2] = 2.0 if i < 30 else 2.4                  # the constructor needs at least two velocity buckets
        x[0] = np.linspace(0, 2.0, 11)
        occ = []
        for k in range(10):
            cx, cy = 0.2 * k + rng.uniform(-0.3, 0.3), rng.uniform(-1.5, 1.5)
            occ.append({'space': Polygon([(cx - 1, cy - 0.5), (cx + 1, cy - 0.5), (cx + 1, cy + 0.5), (cx - 1, cy + 0.5)]),
                        'starttime': float('%.1f' % (0.1 * k)), 'endtime': float('%.1f' % (0.1 * k + 0.1))})
        prims.append(MotionPrimitive(x, np.zeros((2, 10)), 1.0, occupancy=occ))
    MA2 = ManeuverAutomaton(prims)
    assert not MA2.is_timepoint()
    # free space shrinks and shifts over time
    space = [Polygon([(-2 + 0.1 * t, -2.2 + 0.05 * t), (6, -2.2 + 0.05 * t), (6, 2.4 - 0.08 * t), (-2 + 0.1 * t, 2.4 - 0.08 * t)])
             for t in range(9)]
    be = OracleBackend()
    LL._CACHE.clear()
    chk = LL._checker_for(MA2, 0.1, space, False, be)
    for t0 in (0, 3):
        bits = chk.bucket_collides(0.5, 0.1, 0.05, t0, 20, int(MA2._bucket(2.0)), (0, 1))
        c, s = np.cos(0.05), np.sin(0.05)
        for i, p in enumerate(prims[:30]):
            collide = False
            for o in p.occ:
                t = t0 + int(o['starttime'] / 0.1)
                if t > 20 or t >= len(space) - 1:              # reference guards with the intersected list
                    continue
                A = ex.affine_transform(list(o['space'].exterior.coords[:-1]), c, s, 0.5, 0.1)
                for sp in (space[t], space[t + 1]):
                    if not ex.region_contains_convex([list(sp.exterior.coords)], A):
                        collide = True
            assert bool(bits[i]) == collide, (t0, i)
        assert 0 < bits.sum() < len(bits)


# ---------------------------------------------------------------------------------------------------------------------
# whole planner runs on scenarios of the reference's own unit test (unitTests/unitTest.py:35-64: plan, then the plan
# must pass collisionChecker), from the pickled crlite fixtures (tools/make_golden.py unit) -- no /root/reference needed
# ---------------------------------------------------------------------------------------------------------------------
UNIT_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'unit_scenarios.pkl.gz')
# (scenario, bound on popped nodes, plan expected): the stand-alone planner (no decision module) solves the two ZAM
# scenarios; on the urban one the bounded search must fail identically on both backends
UNIT_RUNS = [("ZAM_Zip-1_19_T-1", 20000, True), ("ZAM_Tutorial-1_1_T-1", 20000, True), ("BEL_Zaventem-4_1_T-1", 1500, False)]


def unit_scenario(name):
    import gzip
    import pickle
    with gzip.open(UNIT_FIXTURE) as f:
        return pickle.loads(f.read())[name]


def plan_standalone(MA, name, max_nodes, backend):
    from motionplanner_b200.src.maneuverAutomatonPlannerStandalone import maneuverAutomatonPlannerStandalone
    scenario, pps = unit_scenario(name)
    param = vehicleParameter()
    x, u, _ = maneuverAutomatonPlannerStandalone(scenario, pps, param, MA, backend=backend, max_nodes=max_nodes)
    return scenario, param, x, u


def test_standalone_batched_children_equal_expand_node(MA):
    """the vectorised child construction of the stand-alone planner reproduces its expand_node (reference
    src/maneuverAutomatonPlannerStandalone.py:293-310) bit for bit: trajectories and goal-distance costs"""
    from motionplanner_b200.src import maneuverAutomatonPlannerStandalone as SA
    scenario, pps = unit_scenario("ZAM_Zip-1_19_T-1")
    pp = list(pps.planning_problem_dict.values())[0]
    goals = SA.get_goal_set(scenario, pp)
    assert any(g['space'] is not None for g in goals)

    class AllFree:
        def bucket_collides(self, x, y, phi, t0, limit, bucket, scenes):
            return np.zeros(len(MA.vel_dic[bucket]), bool)

    param = {'timepoint': True, 'steps': 86}
    s0 = pp.initial_state
    x = np.array([[s0.position[0] - 1.15], [s0.position[1]], [s0.velocity], [s0.orientation + 0.1]])
    root = SA.Node([], x, 0)
    for parent in (root, SA._children(AllFree(), root, MA, int(MA._bucket(s0.velocity)), goals, param)[5]):
        bucket = int(MA._bucket(parent.x[2, -1]))
        kids = SA._children(AllFree(), parent, MA, bucket, goals, param)
        T = SA._rotation(parent.x[3, -1])
        assert len(kids) == len(MA.vel_dic[bucket]) > 50
        for k, i in list(zip(kids, MA.vel_dic[bucket]))[::7]:
            one = SA.expand_node(parent, MA.primitives[i], i, T, goals)
            assert k.primitives == one.primitives and k.cost == one.cost and np.isfinite(k.cost)
            assert np.array_equal(k.x, one.x)


def test_standalone_collision_check_with_reference_shaped_arguments(MA):
    """stand-alone twin collision_check(node, primitive, space, goal_set, param) (reference
    src/maneuverAutomatonPlannerStandalone.py:196-226) called the way the reference calls it: a node with
    `primitives, x, cost` only, `space` = list of polygons, `param` with the keys the reference's planner sets"""
    import motionplanner_b200.checker as CK
    from motionplanner_b200.src import maneuverAutomatonPlannerStandalone as SA
    prm, space, ref = corridor_problem(blocked=True)
    param = {'timepoint': True, 'steps': len(space), 'time_step': 0.1, 'b': prm['b']}
    goal_set = [{'time_start': 30, 'time_end': 40, 'space': None, 'lane': None}]

    class RefNode:
        def __init__(self, primitives, x, cost):
            self.primitives, self.x, self.cost = primitives, x, cost

    be = OracleBackend()
    old = CK._ENGINES.copy()
    CK._ENGINES[('planner', int(os.environ.get('MPC_DEVICE', os.environ.get('LOCAL_RANK', '0'))))] = be
    try:
        SA._CACHE.clear()
        x = np.array([[0.0 - 1.15], [0.0], [10.0], [0.0]])
        bucket = int(MA._bucket(10.0))
        chk = LL._checker_for(MA, 0.1, space, True, be)
        want = chk.bucket_collides(x[0, 0], x[1, 0], x[3, 0], 0, len(space), bucket)
        assert 0 < want.sum() < len(want)
        T = SA._rotation(0.0)
        for k, i in list(enumerate(MA.vel_dic[bucket]))[::19]:
            node = SA.expand_node(SA.Node([], x, 0), MA.primitives[i], i, T, goal_set)
            got = SA.collision_check(RefNode(node.primitives, node.x, node.cost), MA.primitives[i], space, goal_set, param)
            assert got == (not want[k])
        assert set(param) == {'timepoint', 'steps', 'time_step', 'b'}
    finally:
        CK._ENGINES.clear()
        CK._ENGINES.update(old)


def test_heap_pops_like_stable_sort():
    """the stand-alone planner's heap on (cost, insertion number) pops what the reference's `queue.sort(key=cost);
    queue.pop(0)` pops (src/maneuverAutomatonPlannerStandalone.py:62-64), ties (inf costs are common) included"""
    import heapq
    rng = np.random.default_rng(0)
    ref, heap, seq = [], [], 0
    for it in range(300):
        for c in rng.choice([np.inf, 1.0, 2.5, 2.5, 7.0, float(rng.uniform(0, 9))], size=int(rng.integers(0, 6))):
            ref.append((float(c), seq))
            heapq.heappush(heap, (float(c), seq))
            seq += 1
        if ref:
            ref.sort(key=lambda e: e[0])
            assert ref.pop(0) == heapq.heappop(heap)


def test_unit_scenario_fixture_plans_on_cpu(MA):
    from motionplanner_b200.auxiliary import collisionChecker as CC
    be = OracleBackend()
    scenario, param, x, u = plan_standalone(MA, "ZAM_Zip-1_19_T-1", 20000, be)
    assert x is not None and x.shape == (4, 85) and u.shape == (2, 84)
    flags, _ = CC.collision_flags(scenario, x, param, engine=OracleBackend())
    assert not flags.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name,max_nodes,expect_plan", UNIT_RUNS, ids=[r[0] for r in UNIT_RUNS])
def test_reference_unit_test_scenarios_on_gpu(MA, name, max_nodes, expect_plan):
    """the search driven by the device's bits is the search driven by the oracle's bits -- same number of batched
    queries, same bits in every one of them, same plan -- and the plan passes ``collisionChecker(scenario, x, param)``
    (the drop-in validator, all steps in one device query), as the reference's unit test demands"""
    from motionplanner_b200.auxiliary.collisionChecker import collisionChecker
    from motionplanner_b200.engine import CollisionEngine

    class Logged(CollisionEngine):
        def query_fast(self, queries, mode=0, want_stats=False, **kw):
            mask, st = super().query_fast(queries, mode, want_stats, **kw)
            self.bits.append(mask.copy())
            return mask, st

        def query(self, queries, cand_idx=None, mode=0):
            bits, st = super().query(queries, cand_idx, mode)
            self.bits.append(np.packbits(bits, bitorder='little'))
            return bits, st

    cpu = OracleBackend()
    _, _, x0, u0 = plan_standalone(MA, name, max_nodes, cpu)
    eng = Logged(0)
    eng.bits = []
    try:
        scenario, param, x1, u1 = plan_standalone(MA, name, max_nodes, eng)
    finally:
        eng.close()
    assert (x0 is not None) == (x1 is not None) == expect_plan
    assert len(eng.bits) == len(cpu.log) > 10
    for got, (_, _, _, want) in zip(eng.bits, cpu.log):
        got = np.unpackbits(np.asarray(got).view(np.uint8), bitorder='little')[:len(want)]
        assert np.array_equal(got, want)
    if expect_plan:
        assert np.array_equal(x0, x1) and np.array_equal(u0, u1)
        assert collisionChecker(scenario, x1, param) is True
