"""prints the measured float32 peaks of cuda:0 (mpc_measure_fp32_peak) as one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionplanner_b200.engine import CollisionEngine

eng = CollisionEngine(0)
print(json.dumps(dict(eng.measure_fp32_peak(), device=eng.device_info()["name"])))
eng.close()
