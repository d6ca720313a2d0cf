#!/usr/bin/env python
"""runs the full-evaluation variant (MPC_MODE_NO_CULL | NO_EARLY_EXIT) of the C4 workload a few times (for ncu)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200 import workloads as W  # noqa: E402
from motionplanner_b200.engine import MODE_NO_CULL, MODE_NO_EARLY_EXIT, CollisionEngine  # noqa: E402

lib, sc, org, q = W.c4_problem(nqueries=2)
eng = CollisionEngine(0)
eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
               sc["step_seg_end"], sc["segs"])
eng.prepare(q, None, MODE_NO_CULL | MODE_NO_EARLY_EXIT)
ms_total, ms_main, st = eng.run_prepared(3, flush_l2=True)
print("k_check full evaluation ms", ms_main, "checks/s", 2 * 4096 * 256 * 50 / (ms_main.mean() * 1e-3))
eng.close()
