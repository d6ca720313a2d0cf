import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['CUDA_LAUNCH_BLOCKING']='1'
import test_planner as TP
from motionplanner_b200.engine import CollisionEngine
from motionplanner_b200.checker import SceneSet
from motionplanner_b200 import workloads as W
param, space, ref = TP.corridor_problem(blocked=True)
eng = CollisionEngine(0)
lib, sc, org, q = W.adversarial_problem(1)
eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
s = SceneSet(); s.add_scene_from_spaces(space)
a = s.arrays()
print({k: (v.shape, v.dtype) for k, v in a.items()})
print(a['step_seg_begin'][:5], a['step_seg_end'][:5], a['step_obst_off'][:5], a['scene_step_off'])
try:
    s.stage(eng); print('stage ok')
except Exception as e: print('ERR', e)
