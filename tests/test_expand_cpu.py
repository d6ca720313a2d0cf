"""Primitive-selection step on the CPU: the oracle's restatement of expand_node's cost and of
goal_check (oracle/oracle_c.c orc_expand) against the reference formulas evaluated with numpy exactly as the reference
writes them (src/lowLevelPlannerManeuverAutomaton.py:199-212, :127-142), and the planner with the expansion routed
through that interface.

Float policy: the cost is float64 arithmetic in the reference's operation order.  On a numpy whose BLAS accumulates the
4x4 @ 4x11 product with fused multiply-adds (every OpenBLAS / MKL build for x86-64 and aarch64) the match is bit for
bit -- asserted below; TOL is the stated tolerance for a BLAS that rounds the product differently."""
import numpy as np
import pytest

import test_planner as TP
from helpers import OracleBackend
from motionplanner_b200.checker import goal_arrays
from motionplanner_b200.geometry import Polygon
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from motionplanner_b200.src import lowLevelPlannerManeuverAutomaton as LL
from oracle import oracle as orc

TOL = 1e-12


@pytest.fixture(scope="module")
def MA():
    return load_maneuver_automaton()


def expansion_case(MA, rng, trial, ref):
    """random parent + goal sets: (root node, bucket, param['goal'], query, parent record)"""
    ind = int(rng.integers(0, 60))
    x = np.zeros((4, ind + 1))
    x[:, -1] = [rng.uniform(0, 80), rng.uniform(-3, 3), rng.uniform(0, 30), rng.uniform(-0.3, 0.3)]
    root = LL.Node([], x, float(rng.uniform(0, 50)))
    bucket = int(MA._bucket(x[2, -1]))
    gx = x[0, -1] + rng.uniform(0, 15)
    ring = [(gx - 4, -6), (gx + 4, -6), (gx + 5, 5), (gx + 1, 1.5), (gx - 3, 6)]              # non-convex
    goals = [{'time_start': ind + 2, 'time_end': ind + 8, 'space': Polygon(ring)},
             {'time_start': ind + 9, 'time_end': ind + 9, 'space': None}]
    if trial % 3 == 0:
        goals = goals[:1]
    if trial % 5 == 4:
        goals = [{'time_start': ind + 3, 'time_end': ind + 30, 'space': Polygon(ring, [[(gx, -1), (gx + 1, -1), (gx + 1, 0), (gx, 0)]])}]
    phi = x[3, -1]
    q = np.zeros(1, orc.QUERY_DTYPE)
    q[0] = (x[0, -1], x[1, -1], np.cos(phi), np.sin(phi), 0, ind, 1000, bucket, 0, len(MA.vel_dic[bucket]), 0, 0)
    return root, bucket, goals, q, np.array([[root.cost, phi]])


def test_oracle_expand_equals_reference_formulas(MA):
    lay = MA.device_layout(0.1)
    states, nstates = MA.trajectory_layout()
    param, _, ref = TP.corridor_problem(blocked=True)
    rng = np.random.default_rng(0)
    total = hits = worst = 0
    for trial in range(25):
        root, bucket, goals, q, parent = expansion_case(MA, rng, trial, ref)
        prm = dict(param, goal=goals)
        rec, segs = goal_arrays(goals)
        cost, goal = orc.expand(states, nstates, ref[0:2], rec, segs, prm['b'], lay['set_off'], lay['set_idx'], q, None, parent)
        T = LL._rotation(root.x[3, -1])
        for k, i in enumerate(MA.vel_dic[bucket]):
            one = LL.expand_node(root, MA.primitives[i], i, ref, [], T)            # the reference arithmetic, per child
            res, gi = LL.goal_check(one, MA.primitives[i], prm)
            assert abs(one.cost - cost[k]) <= TOL * max(1.0, abs(one.cost))
            worst = max(worst, abs(one.cost - cost[k]))
            assert (gi if res else -1) == goal[k]
            hits += bool(res)
            total += 1
    assert total > 3000 and 100 < hits < total
    assert worst == 0.0          # bit for bit with this numpy / BLAS


def test_explicit_candidate_lists_and_short_reference(MA):
    """explicit candidate list; a parent so late that the reference trajectory covers only part of / none of the child"""
    lay = MA.device_layout(0.1)
    states, nstates = MA.trajectory_layout()
    param, _, ref = TP.corridor_problem(blocked=True)
    M = ref.shape[1]
    cand = np.array(MA.vel_dic[12][:40][::-1], np.int32)
    for ind in (0, M - 12, M - 11, M - 5, M - 1, M, M + 7):
        x = np.zeros((4, ind + 1))
        x[:, -1] = [3.0 + ind, 0.4, 12.0, -0.05]
        root = LL.Node([], x, 2.5)
        q = np.zeros(1, orc.QUERY_DTYPE)
        q[0] = (x[0, -1], x[1, -1], np.cos(-0.05), np.sin(-0.05), 0, ind, 1000, -1, 0, len(cand), 0, 0)
        goals = [{'time_start': 0, 'time_end': 10 ** 6, 'space': None}]
        rec, segs = goal_arrays(goals)
        cost, goal = orc.expand(states, nstates, ref[0:2], rec, segs, 1.15, lay['set_off'], lay['set_idx'], q, cand, [[2.5, -0.05]])
        T = LL._rotation(-0.05)
        for k, i in enumerate(cand):
            one = LL.expand_node(root, MA.primitives[i], i, ref, [], T)
            assert one.cost == cost[k]
            assert goal[k] == ind                      # space None: reached at the first step of the window


def test_planner_with_device_style_expansion_returns_the_same_plan(MA):
    """lowLevelPlannerManeuverAutomaton(device_expand=True) through the expand interface == the numpy expansion"""
    for blocked in (False, True):
        plans = []
        for dev in (False, True):
            param, space, ref = TP.corridor_problem(blocked=blocked)
            LL._CACHE.clear()
            be = OracleBackend()
            x, u, _ = LL.lowLevelPlannerManeuverAutomaton(None, None, param, space, ref, MA, backend=be, device_expand=dev)
            assert x is not None
            plans.append((x, u))
        assert np.array_equal(plans[0][0], plans[1][0]) and np.array_equal(plans[0][1], plans[1][1])
