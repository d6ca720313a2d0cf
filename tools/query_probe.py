import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from motionplanner_b200.engine import CollisionEngine
from motionplanner_b200 import workloads as W
for nq in (1, 64):
    lib = W.synthetic_library(4096, 50, seed=0)
    scene = W.dense_traffic_scene(50, 8, 32, seed=1)
    queries, origins = W.dense_traffic_queries(scene, 4096, nq, seed=2)
    eng = CollisionEngine(0)
    eng.upload_library(lib['verts'], lib['nv'], lib['tidx'], lib['set_off'], lib['set_idx'])
    eng.set_scenes(scene['scene_step_off'], origins, scene['step_obst_off'], scene['obst'], scene['obst_nv'], scene['step_seg_begin'], scene['step_seg_end'], scene['segs'])
    for _ in range(30): eng.query_mask(queries)
    print("nq", nq, file=sys.stderr)
    for _ in range(4): m, st = eng.query_mask(queries)
    print("device ms", st['ms_device'], file=sys.stderr)
    eng.close()
