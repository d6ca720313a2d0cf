"""BASELINE.json configs[0..2]: the bundled CommonRoad scenarios (tests/golden/scenarios.npz, extracted by
tools/make_golden.py) with the regenerated 13 702-primitive automaton.  Every collision decision the batched planner
would ask for along a bounded random descent is computed on the GPU and by the oracle and must be identical; the
trajectory validator is compared step by step the same way."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import OracleBackend, load_golden, random_walk, scenario_scene
from motionplanner_b200.checker import PrimitiveChecker, SceneSet, car_library, validate_trajectory
from motionplanner_b200.engine import MODE_PLANNER, MODE_VALIDATOR, CollisionEngine, make_queries
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from oracle import oracle as orc


@pytest.fixture(scope="module")
def MA():
    return load_maneuver_automaton()


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def test_all_bundled_scenarios_planner_decisions(MA, golden):
    gpu = PrimitiveChecker(MA, 0.1, CollisionEngine(0))
    cpu = PrimitiveChecker(MA, 0.1, OracleBackend())
    decisions = disagreements = collided = 0
    for k, (name, g) in enumerate(sorted(golden.items())):
        for chk in (gpu, cpu):
            chk.set_scenes(scenario_scene(g), MODE_PLANNER)
        a, _ = random_walk(gpu, MA, g, np.random.default_rng(k), expansions=10)
        b, _ = random_walk(cpu, MA, g, np.random.default_rng(k), expansions=10)
        assert len(a) == len(b), name
        for x, y in zip(a, b):
            decisions += len(x)
            disagreements += int((x != y).sum())
            collided += int(y.sum())
    print("scenarios %d decisions %d collided %d disagreements %d gpu stats %s" %
          (len(golden), decisions, collided, disagreements, gpu.stats))
    assert decisions > 50000 and 0 < collided < decisions
    assert disagreements == 0
    gpu.backend.close()


def test_us101_full_library(MA, golden):
    """configs[2]: USA_US101-16_2_T-1, the whole library (13 702 primitives) at ind in {0, 10, 20}"""
    g = golden['USA_US101-16_2_T-1']
    lay = MA.device_layout(0.1)
    eng, ora = CollisionEngine(0), OracleBackend()
    scene = scenario_scene(g, 'road').arrays()
    x0 = g['x0']
    q = make_queries(3)
    q['x'], q['y'] = x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15
    q['c'], q['s'] = np.cos(x0[3]), np.sin(x0[3])
    q['t0'], q['step_limit'] = [0, 10, 20], g['t_end']
    q['cand_set'], q['cand_cnt'] = -1, len(MA.primitives)
    cand = np.arange(len(MA.primitives), dtype=np.int32)
    out = []
    for be in (eng, ora):
        be.upload_library(lay['verts'], lay['nv'], lay['tidx'], lay['set_off'], lay['set_idx'])
        be.set_scenes(scene['scene_step_off'], scene['origins'], scene['step_obst_off'], scene['obst'], scene['obst_nv'],
                      scene['step_seg_begin'], scene['step_seg_end'], scene['segs'])
        out.append(be.query(q, cand, MODE_PLANNER))
    (got, st), (want, ost) = out
    assert st['pairs_nominal'] == ost['pairs_nominal'] > 10_000_000
    assert np.array_equal(got, want)
    assert 0 < want.mean() < 1
    eng.close()


def test_validator_on_scenarios(MA, golden):
    """collisionChecker semantics (closed-set intersects for obstacles, contains for the road) step by step along the
    random descents, GPU vs oracle"""
    eng, ora = CollisionEngine(0), OracleBackend()
    lib = car_library(4.3, 1.7)
    walker = PrimitiveChecker(MA, 0.1, OracleBackend())
    steps = flagged = 0
    for k, (name, g) in enumerate(sorted(golden.items())):
        if k % 4:
            continue
        walker.set_scenes(scenario_scene(g), MODE_PLANNER)
        _, traj = random_walk(walker, MA, g, np.random.default_rng(1000 + k), expansions=8)
        traj = traj[:, :g['nsteps']].copy()
        traj[0] += np.cos(traj[3]) * 1.15          # rear axle -> centre (reference transform_trajectory)
        traj[1] += np.sin(traj[3]) * 1.15
        res = []
        for be in (eng, ora):
            be.upload_library(lib['verts'], lib['nv'], lib['tidx'])
            s = SceneSet()
            s.add_scene_explicit(g['nsteps'], g['step_obst_off'], g['obst'], g['obst_nv'], g['road'], origin=g['x0'][0:2])
            s.stage(be)
            res.append(validate_trajectory(be, traj, 4.3, 1.7)[0])
        assert np.array_equal(res[0], res[1]), name
        steps += len(res[0])
        flagged += int(res[1].sum())
    assert steps > 500 and 0 < flagged < steps
    eng.close()


def test_batch_of_scenes_with_device_side_segment_sort(MA, golden):
    """300 scenario-shaped queries against 300 scenes staged in ONE mpc_set_scenes call (the C5 shape): more than 8192
    boundary segments in all, so the ranges are Morton-sorted by k_order_segs on the device instead of on the host"""
    from motionplanner_b200.engine import CollisionEngine, make_queries
    rng = np.random.default_rng(3)
    names = sorted(golden)
    lay = MA.device_layout(0.1)
    scenes = SceneSet()
    q = make_queries(300)
    draw = rng.integers(0, len(names), 300)
    for k, gi in enumerate(draw):
        g = golden[names[gi]]
        n = min(g['nsteps'], 31)
        off = g['step_obst_off'][:n + 1]
        segs = g['road_safe'] if len(g['road_safe']) else g['road']
        scenes.add_scene_explicit(n, off, g['obst'][:off[-1]], g['obst_nv'][:off[-1]], segs, origin=g['x0'][0:2])
        x0 = g['x0']
        bucket = int(MA._bucket(x0[2]))
        q[k] = (x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15, np.cos(x0[3]), np.sin(x0[3]), k, 0,
                g['t_end'], bucket, 0, len(MA.vel_dic[bucket]), 0, 0)
    arrays = scenes.arrays()
    assert len(arrays['segs']) >= 8192 and max(arrays['step_seg_end'] - arrays['step_seg_begin']) <= 4096
    eng = CollisionEngine(0)
    eng.upload_library(lay['verts'], lay['nv'], lay['tidx'], lay['set_off'], lay['set_idx'])
    eng.set_scenes(arrays['scene_step_off'], arrays['origins'], arrays['step_obst_off'], arrays['obst'],
                   arrays['obst_nv'], arrays['step_seg_begin'], arrays['step_seg_end'], arrays['segs'])
    got, st = eng.query(q)
    lib = dict(verts=lay['verts'], nv=lay['nv'], tidx=lay['tidx'], set_off=lay['set_off'], set_idx=lay['set_idx'])
    want, ost = orc.Problem(lib, arrays).check(q, None, orc.MODE_NO_EARLY_EXIT)
    assert np.array_equal(got, want) and st['pairs_nominal'] == ost['pairs_nominal']
    assert 0.2 < got.mean() < 0.95
    # the same scenes staged one at a time go through the host-side sort: same bits
    for k in (0, 17, 123, 299):
        one = SceneSet()
        g = golden[names[draw[k]]]
        n = min(g['nsteps'], 31)
        off = g['step_obst_off'][:n + 1]
        segs = g['road_safe'] if len(g['road_safe']) else g['road']
        one.add_scene_explicit(n, off, g['obst'][:off[-1]], g['obst_nv'][:off[-1]], segs, origin=g['x0'][0:2])
        a1 = one.arrays()
        eng.set_scenes(a1['scene_step_off'], a1['origins'], a1['step_obst_off'], a1['obst'], a1['obst_nv'],
                       a1['step_seg_begin'], a1['step_seg_end'], a1['segs'])
        qk = q[k:k + 1].copy()
        qk['scene'] = 0
        bits, _ = eng.query(qk)
        lo = int(q['cand_cnt'][:k].sum())
        assert np.array_equal(bits, got[lo:lo + int(q['cand_cnt'][k])])
    eng.close()
