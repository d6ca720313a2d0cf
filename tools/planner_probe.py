#!/usr/bin/env python
"""planner-level timing: lowLevelPlannerManeuverAutomaton on the two-lane corridor problem of tests/test_planner.py
(lane change around a slower car), GPU engine vs the CPU oracle standing in for the engine."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import test_planner as TP  # noqa: E402
from helpers import OracleBackend  # noqa: E402
from motionplanner_b200.engine import CollisionEngine  # noqa: E402
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton  # noqa: E402
from motionplanner_b200.src import lowLevelPlannerManeuverAutomaton as LL  # noqa: E402


def main():
    MA = load_maneuver_automaton()
    MA.device_layout(0.1)
    for name, make, dev in (("gpu", lambda: CollisionEngine(0), False), ("gpu+expand", lambda: CollisionEngine(0), True),
                            ("cpu-oracle", OracleBackend, False)):
        be = make()
        times, nq = [], 0
        for rep in range(12):
            param, space, ref = TP.corridor_problem(blocked=True)
            LL._CACHE.clear()
            t0 = time.perf_counter()
            chk = LL._checker_for(MA, 0.1, space, True, be)          # stages library + free space
            t1 = time.perf_counter()
            x, u, _ = LL.lowLevelPlannerManeuverAutomaton(None, None, param, space, ref, MA, backend=be, device_expand=dev)
            t2 = time.perf_counter()
            times.append((t1 - t0, t2 - t1))
            nq = chk.stats['queries']
        st, pl = np.median([t[0] for t in times[2:]]), np.median([t[1] for t in times[2:]])
        print("%-10s plan %s expansions %d | staging %.1f ms | search %.2f ms = %.0f us per expansion"
              % (name, None if x is None else x.shape, nq, 1e3 * st, 1e3 * pl, 1e6 * pl / max(nq, 1)))
        if hasattr(be, 'close'):
            be.close()


if __name__ == '__main__':
    main()
