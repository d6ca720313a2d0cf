#!/usr/bin/env python
"""where the end-to-end step of bench.py spends its time: mpc_set_scenes vs mpc_query, host vs device"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200.engine import CollisionEngine
from motionplanner_b200 import workloads as W

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = W.synthetic_library(4096, 50, seed=0)
scene = W.dense_traffic_scene(50, 8, 32, seed=1)
queries, origins = W.dense_traffic_queries(scene, 4096, nq, seed=2)
eng = CollisionEngine(0)
eng.upload_library(lib['verts'], lib['nv'], lib['tidx'], lib['set_off'], lib['set_idx'])
sc = {k: eng.pinned_like(v) for k, v in scene.items()}
org = eng.pinned_like(origins)
q = eng.pinned_like(queries)
def stage():
    eng.set_scenes(sc['scene_step_off'], org, sc['step_obst_off'], sc['obst'], sc['obst_nv'], sc['step_seg_begin'], sc['step_seg_end'], sc['segs'])
for _ in range(20):
    stage(); eng.query_mask(q)
ts, tq, dev = [], [], []
for _ in range(300):
    t0 = time.perf_counter(); stage(); t1 = time.perf_counter(); _, st = eng.query_mask(q); t2 = time.perf_counter()
    ts.append(t1 - t0); tq.append(t2 - t1); dev.append(st['ms_device'])
print("queries %d | set_scenes p50 %.1f us | query p50 %.1f us (device %.1f us) | step %.1f us" % (
    nq, 1e6 * np.median(ts), 1e6 * np.median(tq), 1e3 * np.median(dev), 1e6 * (np.median(ts) + np.median(tq))))
os.environ['MPC_DEBUG_TIMING'] = '1'
stage()
eng.close()
