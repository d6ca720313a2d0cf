"""Obstacle occupancies from the lane-following prediction (reference auxiliary/prediction.py):
route logic on a hand-built lanelet network, and the array path against the object path (occupancy_at_time +
Rectangle per step, the way the reference consumes the predicted scenario)."""
import os

import numpy as np
import pytest

from motionplanner_b200 import crlite
from motionplanner_b200.auxiliary import prediction as P
from motionplanner_b200.checker import scenario_obstacle_arrays


def straight(lid, x0, x1, y, n=6, width=3.5):
    xs = np.linspace(x0, x1, n)
    return crlite.Lanelet(lid, np.stack([xs, np.full(n, y + width / 2)], 1), np.stack([xs, np.full(n, y - width / 2)], 1))


def network():
    """lane 1 (x 0..50) -> lanes 2 (straight on) and 3 (drifting left); lane 4 parallel to lane 1"""
    l1, l2, l4 = straight(1, 0, 50, 0), straight(2, 50, 120, 0), straight(4, 0, 120, 3.5)
    xs = np.linspace(50, 120, 8)
    ys = np.linspace(0, 14, 8)
    l3 = crlite.Lanelet(3, np.stack([xs, ys + 1.75], 1), np.stack([xs, ys - 1.75], 1))
    l1.successor = [2, 3]
    l2.predecessor = l3.predecessor = [1]
    return crlite.LaneletNetwork([l1, l2, l3, l4])


def car(oid, x, y, v, phi=0.0, acc=None):
    kw = dict(position=np.array([x, y]), velocity=v, orientation=phi, time_step=0)
    if acc is not None:
        kw['acceleration'] = acc
    return crlite.Obstacle(oid, 'dynamic', 'car', crlite.Rectangle(4.5, 1.9), crlite.State(**kw))


def scenario(cars):
    return crlite.Scenario(0.1, 'test', network(), [], list(cars), [])


def test_constant_velocity_along_the_centre_line():
    sc = scenario([car(10, 5.0, 3.4, 10.0)])                      # on lane 4, slightly off centre
    routes = P.predict_routes(sc, 30, np.array([60.0, 0.0, 10.0, 0.0]))
    assert len(routes) == 1 and routes[0]['lanelets'] == [4] * 31
    r = routes[0]
    assert np.allclose(r['position'][1:, 0], 5.0 + 1.0 * np.arange(1, 31)) and np.allclose(r['position'][1:, 1], 3.5)
    assert np.array_equal(r['position'][0], [5.0, 3.4])            # step 0 is the measured state
    assert np.allclose(r['orientation'], 0.0)


def test_branching_at_successors_and_acceleration():
    sc = scenario([car(10, 30.0, 0.0, 10.0, acc=2.0)])
    routes = P.predict_routes(sc, 40, np.array([0.0, 3.5, 5.0, 0.0]))
    assert sorted(r['lanelets'][-1] for r in routes) == [2, 3] and all(len(r['orientation']) == 41 for r in routes)
    x = P.travelled_distance(10.0, 2.0, 0.1, 40)
    assert np.isclose(x[-1], 10.0 * 4 + 0.5 * 2.0 * 16)
    straight_on = [r for r in routes if r['lanelets'][-1] == 2][0]
    assert np.allclose(straight_on['position'][1:, 0], 30.0 + x[1:]) and np.allclose(straight_on['position'][1:, 1], 0.0)
    left = [r for r in routes if r['lanelets'][-1] == 3][0]
    k = left['lanelets'].index(3)
    assert np.allclose(left['orientation'][k:], np.arctan2(2.0, 10.0))
    assert np.array_equal(left['position'][:k], straight_on['position'][:k])          # common prefix on lane 1


def test_braking_to_standstill_and_far_vehicles():
    sc = scenario([car(10, 10.0, 3.5, 4.0, acc=-4.0), car(11, 100.0, 3.5, 0.0)])
    routes = P.predict_routes(sc, 30, np.array([60.0, 0.0, 0.0, 0.0]))              # standing ego: reach 0.5*9*3^2 = 40.5 m
    assert len(routes) == 1 and routes[0]['source'].obstacle_id == 11               # vehicle 10 cannot get there
    r = P.predict_routes(sc, 30, np.array([20.0, 0.0, 0.0, 0.0]))
    mine = [q for q in r if q['source'].obstacle_id == 10][0]
    assert np.isclose(mine['position'][-1, 0], mine['position'][12, 0])             # stopped after v/a = 1 s
    assert mine['position'][-1, 0] < 10.0 + 2.3


def test_vehicle_running_into_the_ego_from_behind_is_cut():
    sc = scenario([car(10, 2.0, 0.0, 20.0)])
    ego = np.array([40.0, 0.0, 5.0, 0.0])
    routes = P.predict_routes(sc, 30, ego)
    assert len(routes) == 1
    n = len(routes[0]['orientation'])
    assert 1 < n < 31
    assert routes[0]['position'][-1, 0] <= 40.0 - 5 - 2.25                          # stops short of the gap behind the ego
    assert routes[0]['position'][-1, 0] + 2.0 > 40.0 - 5 - 2.25


def test_prediction_replaces_the_obstacles_and_arrays_equal_the_object_path():
    cars = [car(10, 30.0, 0.1, 10.0, phi=0.02, acc=1.0), car(11, 5.0, 3.4, 14.0), car(12, 70.0, 3.5, 3.0, acc=-2.0)]
    sc = scenario(cars)
    ego = np.array([20.0, 0.0, 8.0, 0.0])
    out = P.prediction(sc, 35, ego)
    assert out is sc and len(sc.dynamic_obstacles) == 4                              # car 10 branches into two routes
    assert all(o.obstacle_id not in (10, 11, 12) for o in sc.dynamic_obstacles)
    off0, obst0, nv0 = scenario_obstacle_arrays(sc, 36)
    off1, obst1, nv1 = P.occupancy_arrays(sc._predicted_routes, 36)
    assert np.array_equal(off0, off1) and np.array_equal(nv0, nv1)
    assert np.array_equal(obst0[:, :4], obst1)                                       # bit for bit
    assert off1[-1] == sum(min(len(r['orientation']), 36) for r in sc._predicted_routes)


REF_SCENARIOS = '/root/reference/scenarios'


@pytest.mark.skipif(not os.path.isdir(REF_SCENARIOS), reason="reference scenarios not mounted (build container only)")
@pytest.mark.parametrize("name", ['ZAM_Zip-1_19_T-1.xml', 'USA_US101-16_2_T-1.xml', 'DEU_Flensburg-73_1_T-1.xml'])
def test_prediction_on_bundled_scenarios(name):
    path = os.path.join(REF_SCENARIOS, name)
    if not os.path.exists(path):
        pytest.skip("scenario not bundled")
    sc, pps = crlite.CommonRoadFileReader(path).open()
    s0 = list(pps.planning_problem_dict.values())[0].initial_state
    ego = np.array([s0.position[0], s0.position[1], s0.velocity, s0.orientation])
    n_before = len(sc.dynamic_obstacles)
    P.prediction(sc, 30, ego)
    routes = sc._predicted_routes
    assert n_before > 0 and len(routes) > 0
    off0, obst0, nv0 = scenario_obstacle_arrays(sc, 31)
    off1, obst1, nv1 = P.occupancy_arrays(routes, 31)
    assert np.array_equal(off0, off1) and np.array_equal(obst0[:, :4], obst1)
    for r in routes:                                                                 # every pose lies on its lanelet's centre line
        for k in range(1, len(r['orientation'])):
            l = sc.lanelet_network.find_lanelet_by_id(r['lanelets'][k])
            d = np.linalg.norm(l.center_vertices - r['position'][k], axis=1).min()
            seg = np.linalg.norm(np.diff(l.center_vertices, axis=0), axis=1).max()
            assert d <= seg
