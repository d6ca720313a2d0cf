#!/usr/bin/env python
"""planning-step latency on real scenario shapes (golden fixtures): one velocity bucket (<=195 candidates x 11
occupancies) against the obstacles and road boundary of a bundled scenario, host doubles in -> bitmask out."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import load_golden, scenario_scene  # noqa: E402
from motionplanner_b200.checker import PrimitiveChecker  # noqa: E402
from motionplanner_b200.engine import CollisionEngine, make_queries  # noqa: E402
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton  # noqa: E402


def main():
    MA = load_maneuver_automaton()
    g = load_golden()
    eng = CollisionEngine(0)
    chk = PrimitiveChecker(MA, 0.1, eng)
    for name in ('ZAM_Zip-1_19_T-1', 'USA_US101-16_2_T-1', 'DEU_Flensburg-10_1_T-1', 'ZAM_ACC-1_2_S-1'):
        sc = g[name]
        t0 = time.perf_counter()
        chk.set_scenes(scenario_scene(sc))
        t_stage = time.perf_counter() - t0
        x0 = sc['x0']
        bucket = int(MA._bucket(x0[2]))
        args = (x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15, x0[3], 0, sc['t_end'], bucket)
        for _ in range(20):
            chk.bucket_collides(*args)
        lat = []
        for _ in range(300):
            t1 = time.perf_counter()
            bits = chk.bucket_collides(*args)
            lat.append(time.perf_counter() - t1)
        q = make_queries(1)
        q['x'], q['y'], q['c'], q['s'] = args[0], args[1], np.cos(args[2]), np.sin(args[2])
        q['t0'], q['step_limit'], q['cand_set'], q['cand_cnt'] = 0, sc['t_end'], bucket, len(MA.vel_dic[bucket])
        eng.prepare(q, None, 0)
        ms_total, ms_main, st = eng.run_prepared(100, flush_l2=False)
        nseg = len(sc['road_safe']) if len(sc['road_safe']) else len(sc['road'])
        print('%-24s cand %3d obst/step %4.1f segs %4d | stage scene %.2f ms | query p50 %.1f us p90 %.1f us | device total %.1f us '
              'k_check %.1f us | collide %d/%d refined %d' %
              (name, len(bits), len(sc['obst_nv']) / sc['nsteps'], nseg, 1e3 * t_stage, 1e6 * np.median(lat),
               1e6 * np.percentile(lat, 90), 1e3 * np.median(ms_total), 1e3 * np.median(ms_main), bits.sum(), len(bits),
               st['refined_fp64']))
    eng.close()


if __name__ == '__main__':
    main()
