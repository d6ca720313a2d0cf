"""mpc_expand on the GPU against the CPU oracle: collision bits, child costs (bit for bit: same float64
operations in the same order on both sides) and goal steps of whole expansion batches; and the planner with the
expansion on the device."""
import numpy as np
import pytest

import test_planner as TP
from helpers import OracleBackend
from motionplanner_b200.checker import goal_arrays
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from motionplanner_b200.src import lowLevelPlannerManeuverAutomaton as LL
from oracle import oracle as orc
from test_expand_cpu import expansion_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def MA():
    return load_maneuver_automaton()


@pytest.fixture(scope="module")
def staged(MA):
    """engine + oracle stand-in with the automaton and the corridor scene staged"""
    from motionplanner_b200.engine import CollisionEngine
    eng, cpu = CollisionEngine(0), OracleBackend()
    param, space, ref = TP.corridor_problem(blocked=True)
    out = []
    for be in (eng, cpu):
        LL._CACHE.clear()
        chk = LL._checker_for(MA, 0.1, space, True, be)
        out.append(chk)
    yield eng, cpu, out[0], out[1], param, ref
    eng.close()


def test_expand_batches_match_the_oracle(MA, staged):
    eng, cpu, _, _, param, ref = staged
    states, nstates = MA.trajectory_layout()
    for be in (eng, cpu):
        be.upload_trajectories(states, nstates)
    rng = np.random.default_rng(5)
    total = hits = 0
    for trial in range(12):
        cases = [expansion_case(MA, rng, trial, ref) for _ in range(int(rng.integers(1, 6)))]
        goals = cases[0][2]
        rec, segs = goal_arrays(goals)
        q = np.concatenate([c[3] for c in cases])
        q['step_limit'] = param['steps']
        parent = np.concatenate([c[4] for c in cases])
        for be in (eng, cpu):
            be.set_guidance(ref[0:2], rec, segs, param['b'])
        b1, c1, g1, st = eng.expand(q, parent)
        b0, c0, g0, _ = cpu.expand(q, parent)
        assert np.array_equal(b0, b1)
        assert np.array_equal(c0, c1)                         # float64, identical operation order: bit for bit
        assert np.array_equal(g0, g1)
        assert st['kernels'] >= 2                             # k_expand + k_check (single-kernel query: refinement in its tail)
        total += len(c0)
        hits += int((g0 >= 0).sum())
    assert total > 2000 and hits > 50


def test_expand_explicit_lists_and_short_reference(MA, staged):
    eng, cpu, _, _, param, ref = staged
    lay = MA.device_layout(0.1)
    M = ref.shape[1]
    cand = np.array(MA.vel_dic[12][:40][::-1] + MA.vel_dic[3][:7], np.int32)
    rec, segs = goal_arrays([{'time_start': 5, 'time_end': 10 ** 6, 'space': None}])
    for be in (eng, cpu):
        be.upload_trajectories(*MA.trajectory_layout())
        be.set_guidance(ref[0:2], rec, segs, 1.15)
    q = np.zeros(7, orc.QUERY_DTYPE)
    parent = np.zeros((7, 2))
    for k, ind in enumerate((0, M - 12, M - 11, M - 5, M - 1, M, M + 7)):
        q[k] = (3.0 + ind, 0.4, np.cos(-0.05), np.sin(-0.05), 0, ind, param['steps'], -1, (k % 2) * 7, len(cand) - 7, 0, 0)
        parent[k] = (2.5 * k, -0.05)
    b1, c1, g1, _ = eng.expand(q, parent, cand)
    b0, c0, g0, _ = cpu.expand(q, parent, cand)
    assert np.array_equal(b0, b1) and np.array_equal(c0, c1) and np.array_equal(g0, g1)
    assert (c1[-(len(cand) - 7):] == parent[-1, 0]).all()     # no reference left: cost unchanged


def test_expand_requires_staging(MA):
    from motionplanner_b200.engine import CollisionEngine, MpcError
    eng = CollisionEngine(0)
    param, space, ref = TP.corridor_problem(blocked=False)
    LL._CACHE.clear()
    LL._checker_for(MA, 0.1, space, True, eng)
    q = np.zeros(1, orc.QUERY_DTYPE)
    q[0] = (0, 0, 1, 0, 0, 0, 10, 3, 0, 5, 0, 0)
    with pytest.raises(MpcError):
        eng.expand(q, [[0.0, 0.0]])
    eng.close()


def test_planner_device_expansion_same_plan(MA):
    from motionplanner_b200.engine import CollisionEngine
    for blocked in (False, True):
        plans = []
        for dev in (False, True):
            param, space, ref = TP.corridor_problem(blocked=blocked)
            LL._CACHE.clear()
            eng = CollisionEngine(0)
            x, u, _ = LL.lowLevelPlannerManeuverAutomaton(None, None, param, space, ref, MA, backend=eng, device_expand=dev)
            eng.close()
            assert x is not None
            plans.append((x, u))
        assert np.array_equal(plans[0][0], plans[1][0]) and np.array_equal(plans[0][1], plans[1][1])
