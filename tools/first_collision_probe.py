#!/usr/bin/env python
"""distribution of the first colliding time slot over the candidates of C4 queries, and what it means for the early
return across time slots: live groups (any undecided lane) vs live candidates / 32 (perfect compaction)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200.engine import CollisionEngine
from motionplanner_b200 import workloads as W

P, T = 4096, 50
lib = W.synthetic_library(P, T, seed=0)
scene = W.dense_traffic_scene(T, 8, 32, seed=1)
queries, origins = W.dense_traffic_queries(scene, P, 8, seed=2)
eng = CollisionEngine(0)
eng.upload_library(lib['verts'], lib['nv'], lib['tidx'], lib['set_off'], lib['set_idx'])
eng.set_scenes(scene['scene_step_off'], origins, scene['step_obst_off'], scene['obst'], scene['obst_nv'],
               scene['step_seg_begin'], scene['step_seg_end'], scene['segs'])
first = np.full((len(queries), P), T, np.int32)
for k in range(T - 1, -1, -1):
    q = queries.copy(); q['step_limit'] = k
    bits, _ = eng.query(q)
    bits = bits.reshape(len(queries), P).astype(bool)
    first[bits] = k                      # collides using slots <= k
for lag in (0, 2, 4):
    f = first
    live_groups = live_c32 = tot = 0
    for j in range(T):
        live = f > (j - lag)            # undecided when slot j starts (results lag `lag` slots)
        g = live.reshape(len(queries), P // 32, 32)
        live_groups += g.any(axis=2).sum()
        live_c32 += np.ceil(live.sum(axis=1) / 32).sum()
        tot += len(queries) * (P // 32)
    print("lag", lag, "tasks total", tot, "live groups (caller order)", live_groups, "perfectly compacted groups", int(live_c32))
print("first-collision slot quantiles", np.percentile(first, [10, 25, 50, 75, 90, 99]), "never", (first == T).mean())
eng.close()
