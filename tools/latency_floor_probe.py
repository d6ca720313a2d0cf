#!/usr/bin/env python
"""fixed cost of a planner-sized query: the same call with step_limit = -1 (every CTA leaves at the time guard, the tail
still hands the result block over), next to the real query; USA_US101-16_2_T-1"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200 import workloads as W  # noqa: E402
from motionplanner_b200.checker import PrimitiveChecker, SceneSet  # noqa: E402
from motionplanner_b200.engine import MODE_PLANNER, CollisionEngine  # noqa: E402
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton  # noqa: E402

MA = load_maneuver_automaton()
g = W.load_golden_scenarios()["USA_US101-16_2_T-1"]
eng = CollisionEngine(0)
chk = PrimitiveChecker(MA, 0.1, eng)
s = SceneSet()
s.add_scene_explicit(g["nsteps"], g["step_obst_off"], g["obst"], g["obst_nv"], g["road_safe"], origin=g["x0"][0:2])
chk.set_scenes(s, MODE_PLANNER)
x0 = g["x0"]
x, y = x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15
bucket = int(MA._bucket(x0[2]))
for name, limit in (("real query", g["t_end"]), ("all CTAs leave at the guard", -1), ("real query", g["t_end"])):
    for _ in range(30):
        chk.bucket_collides(x, y, x0[3], 0, limit, bucket)
    lat = []
    for _ in range(500):
        t0 = time.perf_counter()
        chk.bucket_collides(x, y, x0[3], 0, limit, bucket)
        lat.append(time.perf_counter() - t0)
    print("%-30s p50 %.1f us  p10 %.1f us" % (name, 1e6 * np.percentile(lat, 50), 1e6 * np.percentile(lat, 10)))
# python-side share: the call without the C function
t0 = time.perf_counter()
for _ in range(2000):
    np.cos(x0[3]); np.sin(x0[3])
print("np.cos + np.sin per call %.2f us" % (1e6 * (time.perf_counter() - t0) / 2000))
eng.close()
