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
