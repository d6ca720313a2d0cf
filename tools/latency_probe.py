#!/usr/bin/env python
"""planner-shaped query latency on bundled scenarios (the `latency` block of bench.py) as one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from motionplanner_b200.engine import CollisionEngine  # noqa: E402

print(json.dumps(bench.latency_block(lambda: CollisionEngine(0))))
