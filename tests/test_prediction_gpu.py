"""prediction -> arrays -> mpc_set_scenes -> query on the GPU against the oracle: the CARLA-style cycle
(reference mainCARLA.py:386-406: predict, build the free space, plan) with the predicted occupancies staged directly"""
import numpy as np
import pytest

from helpers import OracleBackend
from motionplanner_b200.auxiliary import prediction as P
from motionplanner_b200.checker import PrimitiveChecker, SceneSet
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from test_prediction_cpu import car, scenario

pytestmark = pytest.mark.gpu


def test_predicted_occupancies_staged_directly():
    from motionplanner_b200.engine import CollisionEngine
    MA = load_maneuver_automaton()
    rng = np.random.default_rng(11)
    xs = [x for x in rng.uniform(0, 110, 40) if abs(x - 22.0) > 14.0][:9]           # nobody on top of the ego at step 0
    cars = [car(10 + k, float(x), float(rng.choice([0.0, 3.5])), float(rng.uniform(6, 12)), acc=float(rng.uniform(-1, 1)))
            for k, x in enumerate(xs)]
    sc = scenario(cars)
    ego = np.array([22.0, 0.0, 9.0, 0.0])
    routes = P.predict_routes(sc, 30, ego)
    off, obst, nv = P.occupancy_arrays(routes, 31)
    assert off[-1] > 150
    scenes = SceneSet()
    scenes.add_scene_explicit(31, off, obst, nv, None, origin=ego[0:2])
    eng, cpu = CollisionEngine(0), OracleBackend()
    bits = []
    for be in (eng, cpu):
        chk = PrimitiveChecker(MA, 0.1, be)
        chk.set_scenes(scenes)
        out = []
        for t0 in (0, 10, 20):
            for bucket in (int(MA._bucket(9.0)), int(MA._bucket(2.0)), int(MA._bucket(25.0))):
                out.append(chk.bucket_collides(ego[0] - 1.15 + 0.9 * t0, 0.0, 0.01 * t0, t0, 31, bucket))
        bits.append(np.concatenate(out))
    eng.close()
    assert np.array_equal(bits[0], bits[1])
    assert 0.05 < bits[0].mean() < 0.95
