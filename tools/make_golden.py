#!/usr/bin/env python
"""Builds tests/golden/scenarios.npz from the reference's bundled CommonRoad scenarios (/root/reference/scenarios).

The GPU box has no /root/reference, so everything the scenario parity tests need is extracted here, in the build
container, with this package's own reader (motionplanner_b200.crlite) and road-boundary extraction:

  per scenario: dt, initial state (x, y, v, phi), goal time window, goal polygon (if any), number of steps,
                boundary segments of the road (all lanelets: validator; same-direction lanelets reachable from the
                start: stand-alone planner), the convex obstacle pieces of every time step in float64.

Also writes tests/golden/unit_scenarios.pkl.gz: three of the scenarios the reference's own unit test plans on
(unitTests/unitTest.py:35-38), parsed by this package's reader and pickled as crlite objects, so that the whole planner
(not only single queries) can run on the GPU box.

Also (argument `all`) tests/golden/all_scenarios.pkl.xz: every bundled scenario as crlite objects.

Run from the repository root:  python tools/make_golden.py [npz|unit|all]
"""
import glob
import gzip
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from motionplanner_b200.auxiliary.polygonLaneletNetwork import polygonLaneletNetwork  # noqa: E402
from motionplanner_b200.checker import scenario_obstacle_arrays  # noqa: E402
from motionplanner_b200.crlite import CommonRoadFileReader  # noqa: E402


def main(src='/root/reference/scenarios', dst=os.path.join(ROOT, 'tests', 'golden', 'scenarios.npz')):
    out = {}
    names = []
    for f in sorted(glob.glob(os.path.join(src, '*.xml'))):
        name = os.path.basename(f)[:-4]
        scenario, pps = CommonRoadFileReader(f).open()
        pp = list(pps.planning_problem_dict.values())[0]
        s0 = pp.initial_state
        x0 = np.array([s0.position[0], s0.position[1], s0.velocity, s0.orientation])
        t_start = min(g.time_step.start for g in pp.goal.state_list) - s0.time_step
        t_end = max(g.time_step.end for g in pp.goal.state_list) - s0.time_step
        nsteps = int(t_end) + 1
        goal_ring = np.zeros((0, 2))
        g0 = pp.goal.state_list[0]
        if hasattr(g0, 'position') and hasattr(g0.position, 'vertices'):
            goal_ring = np.asarray(g0.position.vertices)
        road = polygonLaneletNetwork(scenario)
        try:
            road_safe = polygonLaneletNetwork(scenario, safe_only=True, x0=x0).segments
        except Exception:
            road_safe = np.zeros((0, 4))
        off, obst, nv = scenario_obstacle_arrays(scenario, nsteps)
        V = obst.shape[1]
        names.append(name)
        out[name + '/meta'] = np.array([scenario.dt, t_start, t_end, nsteps, len(scenario.obstacles), V], dtype=np.float64)
        out[name + '/x0'] = x0
        out[name + '/goal_ring'] = goal_ring
        out[name + '/road'] = road.segments
        out[name + '/road_safe'] = road_safe
        out[name + '/step_obst_off'] = off
        out[name + '/obst'] = obst
        out[name + '/obst_nv'] = nv
        print('%-32s steps %3d obstacles %2d pieces %5d road segs %4d safe %4d' %
              (name, nsteps, len(scenario.obstacles), len(nv), len(road.segments), len(road_safe)))
    out['names'] = np.array(names)
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst) / 1e6, 'MB')


UNIT_SCENARIOS = ("ZAM_Zip-1_19_T-1", "ZAM_Tutorial-1_1_T-1", "BEL_Zaventem-4_1_T-1")


def unit_scenarios(src='/root/reference/scenarios', dst=os.path.join(ROOT, 'tests', 'golden', 'unit_scenarios.pkl.gz')):
    out = {}
    for name in UNIT_SCENARIOS:
        out[name] = CommonRoadFileReader(os.path.join(src, name + '.xml')).open(lanelet_assignment=True)
    with gzip.GzipFile(dst, 'wb', mtime=0) as f:
        f.write(pickle.dumps(out, protocol=4))
    print('wrote', dst, os.path.getsize(dst) / 1e3, 'kB')


def all_scenarios(src='/root/reference/scenarios', dst=os.path.join(ROOT, 'tests', 'golden', 'all_scenarios.pkl.xz')):
    """all bundled scenarios as crlite objects (scenario, planning problem set), for whole-planner runs on the GPU box
    (BASELINE.json configs[1]; tests/test_all_scenarios_gpu.py)"""
    import lzma
    out = {}
    for f in sorted(glob.glob(os.path.join(src, '*.xml'))):
        out[os.path.basename(f)[:-4]] = CommonRoadFileReader(f).open(lanelet_assignment=True)
    with open(dst, 'wb') as f:
        f.write(lzma.compress(pickle.dumps(out, protocol=4), preset=6))
    print('wrote', dst, len(out), 'scenarios', os.path.getsize(dst) / 1e6, 'MB')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'npz+unit'
    if what in ('npz', 'npz+unit'):
        main()
    if what in ('unit', 'npz+unit'):
        unit_scenarios()
    if what == 'all':
        all_scenarios()
