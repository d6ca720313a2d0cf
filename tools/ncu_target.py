#!/usr/bin/env python
"""A bounded number of launches of one regime, for ncu captures (tools/capture_r02.sh):

    python tools/ncu_target.py c4|mixed|cull|full|planner|aroc [iters]

c4 / mixed / cull / full: one 64- (cull: 16-, full: 8-) query batch of the 4096 x 256 x 50 shape, `iters` timed
launches through mpc_run_prepared (plus its sizing run and the event-bracketed second pass).  planner: `iters`
planner-shaped queries (195 candidates x 11 occupancies) on USA_US101-16_2_T-1 and ZAM_ACC-1_2_S-1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from motionplanner_b200 import workloads as W  # noqa: E402
from motionplanner_b200.engine import (MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_PLANNER, CollisionEngine)  # noqa: E402


def stage(eng, lib, sc, org):
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
                   sc["step_seg_end"], sc["segs"])


def main():
    what = sys.argv[1]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    eng = CollisionEngine(0)
    if what in ("c4", "cull", "full"):
        lib, sc, org, q = W.c4_problem(nqueries=64)
        stage(eng, lib, sc, org)
        mode, n = {"c4": (MODE_PLANNER, 64), "cull": (MODE_NO_EARLY_EXIT, 16), "full": (MODE_NO_CULL | MODE_NO_EARLY_EXIT, 8)}[what]
        eng.prepare(q[:n], None, mode)
        ms, main_ms, st = eng.run_prepared(iters, flush_l2=True)
        print(what, "ms per batch", ms.mean(), "k_check", main_ms.mean(), st)
    elif what == "mixed":
        lib, sc, org, q = W.mixed_traffic_problem(nqueries=64)
        stage(eng, lib, sc, org)
        eng.prepare(q, None, MODE_PLANNER)
        ms, main_ms, st = eng.run_prepared(iters, flush_l2=True)
        print(what, "ms per batch", ms.mean(), "k_check", main_ms.mean(), st)
    elif what == "aroc":
        for ov in (4, 8):
            lib, sc, org, q = W.aroc_shaped_problem(P=4096, T=10, nqueries=16, obst_verts=ov)
            stage(eng, lib, sc, org)
            eng.prepare(q, None, MODE_PLANNER)
            ms, main_ms, st = eng.run_prepared(iters, flush_l2=True)
            print(what, ov, "ms per batch", ms.mean(), "k_check", main_ms.mean(), st)
    elif what in ("planner", "planner_acc"):
        from motionplanner_b200.checker import PrimitiveChecker, SceneSet
        from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
        MA = load_maneuver_automaton()
        golden = W.load_golden_scenarios()
        chk = PrimitiveChecker(MA, 0.1, eng)
        for name in (("USA_US101-16_2_T-1",) if what == "planner" else ("ZAM_ACC-1_2_S-1",)):
            g = golden[name]
            s = SceneSet()
            s.add_scene_explicit(g["nsteps"], g["step_obst_off"], g["obst"], g["obst_nv"],
                                 g["road_safe"] if len(g["road_safe"]) else g["road"], origin=g["x0"][0:2])
            chk.set_scenes(s, MODE_PLANNER)
            x0 = g["x0"]
            for _ in range(iters):
                bits = chk.bucket_collides(x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15, x0[3], 0, g["t_end"],
                                           int(MA._bucket(x0[2])))
            print(name, int(bits.sum()), len(bits))
    eng.close()


if __name__ == "__main__":
    main()
