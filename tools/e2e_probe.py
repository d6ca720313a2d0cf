#!/usr/bin/env python
"""where the host time of one producer-realistic cycle goes (refill, marshalling, mpc_submit, mpc_wait), C4 batch of 64"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from motionplanner_b200._cabi import MODE_NO_TIMING  # noqa: E402
from motionplanner_b200.engine import CollisionEngine  # noqa: E402
from motionplanner_b200.lanes import SCENE_KEYS, CollisionLanes  # noqa: E402

lib, scene, origins, poses = bench.build_c4(64, 0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pool = CollisionLanes(0, lanes=L)
pool.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
bufs = []
for _ in range(L):
    d = dict(scene, origins=origins)
    for k in SCENE_KEYS:
        d[k] = pool.pinned_like(np.ascontiguousarray(d[k]))
    bufs.append((d, pool.pinned_like(poses[0])))
variants = [scene["obst"] + np.array([0.03 * v, 0.0]) for v in range(4)]
T = dict(refill=0.0, marshal=0.0, submit=0.0, wait=0.0)
n = 2000
inflight = [False] * L
t_all = time.perf_counter()
for i in range(n):
    lane = i % L
    eng = pool.engines[lane]
    d, q = bufs[lane]
    t0 = time.perf_counter()
    if inflight[lane]:
        eng.wait(False)
    t1 = time.perf_counter()
    np.copyto(d["obst"], variants[i % 4])
    np.copyto(q.view(np.uint8), poses[(i // 4) % 4].view(np.uint8))
    t2 = time.perf_counter()
    args = CollisionEngine.scene_args(*[d[k] for k in SCENE_KEYS])
    t3 = time.perf_counter()
    eng.submit(args, q, MODE_NO_TIMING)
    t4 = time.perf_counter()
    inflight[lane] = True
    T["wait"] += t1 - t0; T["refill"] += t2 - t1; T["marshal"] += t3 - t2; T["submit"] += t4 - t3
for lane in range(L):
    pool.engines[lane].wait(False)
tot = time.perf_counter() - t_all
print("lanes %d: %.1f us per cycle; " % (L, 1e6 * tot / n) + ", ".join("%s %.1f us" % (k, 1e6 * v / n) for k, v in T.items()))
pool.close()
