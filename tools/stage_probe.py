#!/usr/bin/env python
"""host-side cost of staging one C4 scene (mpc_set_scenes) and of one 16-query mpc_query, from pinned buffers"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200 import workloads as W
from motionplanner_b200.engine import CollisionEngine

lib, sc, org, q = W.c4_problem(nqueries=16)
eng = CollisionEngine(0)
eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
for k in ("obst", "obst_nv", "segs", "step_obst_off", "step_seg_begin", "step_seg_end", "scene_step_off"):
    sc[k] = eng.pinned_like(sc[k])
org = eng.pinned_like(org); q = eng.pinned_like(q)
def stage():
    eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"], sc["step_seg_end"], sc["segs"])
for _ in range(10): stage(); eng.query_mask(q, None, 0)
ts, tq = [], []
for _ in range(200):
    t0 = time.perf_counter(); stage(); t1 = time.perf_counter(); eng.query_mask(q, None, 0); t2 = time.perf_counter()
    ts.append(t1 - t0); tq.append(t2 - t1)
print("set_scenes p50 %.1f us   query(16) p50 %.1f us" % (1e6 * np.median(ts), 1e6 * np.median(tq)))
eng.close()
