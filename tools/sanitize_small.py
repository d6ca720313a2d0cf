#!/usr/bin/env python
"""small mixed workload for compute-sanitizer runs (memcheck / racecheck): multi-phase tiles, both modes, queues"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionplanner_b200 import workloads as W  # noqa: E402
from motionplanner_b200.engine import (MODE_COUNT_WORK, MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_VALIDATOR,  # noqa: E402
                                       CollisionEngine)

eng = CollisionEngine(0)
for seed, kw in ((0, {}), (1, dict(lib_verts=12, obst_verts=8)), (2, dict(P=64, max_obst=60))):
    lib, sc, org, q = W.adversarial_problem(seed, **kw)
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
                   sc["step_seg_end"], sc["segs"])
    base, _ = eng.query(q, None, 0)
    for mode in (MODE_VALIDATOR, MODE_NO_CULL, MODE_NO_EARLY_EXIT | MODE_COUNT_WORK):
        bits, st = eng.query(q, None, mode)
    print("seed", seed, "collide", int(base.sum()), "of", base.size)
lib, sc, org, q = W.c4_problem(P=256, slots=12, T=6, nqueries=2)
eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
               sc["step_seg_end"], sc["segs"])
bits, st = eng.query(q, None, 0)
print("c4-small collide", int(bits.sum()), "of", bits.size, st["pairs_nominal"])
eng.close()
print("sanitize workload done")
