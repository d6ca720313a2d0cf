"""Randomised parity campaign: CUDA path (through the C ABI) against the float64 exact-sign oracle on a stream of seeded
problems whose shape parameters are themselves drawn at random -- near-contact obstacle placement (0, +-1e-15 ... +-1e-3
m), rectangles and general convex polygons up to the ABI's vertex limits, scenes up to several km from the origin,
ragged candidate ranges, both tie conventions (planner ``contains``: src/lowLevelPlannerManeuverAutomaton.py:122,
validator ``intersects``: auxiliary/collisionChecker.py:32).

The suite run takes a few seconds (MPC_CAMPAIGN_SECONDS, default 6; MPC_CAMPAIGN_SEED = first seed, default 50000).  A long run

    MPC_CAMPAIGN_SECONDS=120 MPC_CAMPAIGN_REPORT=gpurun_out/campaign.json python -m pytest tests/test_parity_campaign_gpu.py -m gpu -q

writes the totals (problems, decisions, disagreements, refined pairs, exact ties, near-tangent contacts) as one JSON
line; profiles/ keeps the one of the round.  Bar: zero disagreements."""
import json
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from motionplanner_b200 import workloads as W
from motionplanner_b200.engine import MODE_PLANNER, MODE_VALIDATOR, CollisionEngine
from oracle import oracle as orc


def draw_problem(seed):
    rng = np.random.default_rng(seed)
    general = rng.random() < 0.4
    far = rng.random() < 0.3
    kw = dict(P=int(rng.choice([32, 64, 96, 200, 320])), J=int(rng.integers(1, 9)), nsteps=int(rng.integers(2, 12)),
              max_obst=int(rng.choice([3, 10, 40])),
              lib_verts=int(rng.integers(5, 17)) if general else 4,
              obst_verts=int(rng.integers(5, 9)) if general and rng.random() < 0.7 else 4,
              offset=(float(rng.uniform(-6000, 6000)), float(rng.uniform(-6000, 6000))) if far else (0.0, 0.0),
              nqueries=int(rng.integers(1, 7)), with_region=bool(rng.random() < 0.8))
    return kw, W.adversarial_problem(int(seed), **kw)


def test_campaign_against_oracle():
    budget = float(os.environ.get("MPC_CAMPAIGN_SECONDS", "6"))
    eng = CollisionEngine(0)
    tot = dict(problems=0, decisions=0, disagreements=0, colliding=0, refined_fp64=0, exact_ties=0, near_tangent_lt_1e_9=0,
               parity_refined=0, pairs_nominal=0)
    bad = []
    t0 = time.perf_counter()
    seed = seed0 = int(os.environ.get("MPC_CAMPAIGN_SEED", "50000"))
    try:
        while time.perf_counter() - t0 < budget or tot["problems"] < 8:
            kw, (lib, sc, org, q) = draw_problem(seed)
            eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
            eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"],
                           sc["step_seg_begin"], sc["step_seg_end"], sc["segs"])
            pb = orc.Problem(lib, sc)
            for mode, omode in ((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT)):
                got, st = eng.query(q, None, mode)
                want, ost = pb.check(q, None, omode | orc.MODE_NO_EARLY_EXIT)
                nbad = int((got != want).sum())
                if nbad or st["pairs_nominal"] != ost["pairs_nominal"]:
                    bad.append((seed, mode, kw, nbad))
                tot["decisions"] += len(got)
                tot["disagreements"] += nbad
                tot["colliding"] += int(got.sum())
                tot["refined_fp64"] += st["refined_fp64"]
                tot["exact_ties"] += st["exact_ties"]
                tot["near_tangent_lt_1e_9"] += st["near_tangent"]
                tot["parity_refined"] += st["parity_refined"]
                tot["pairs_nominal"] += st["pairs_nominal"]
            tot["problems"] += 1
            seed += 1
    finally:
        eng.close()
    tot["seconds"] = round(time.perf_counter() - t0, 1)
    tot["seeds"] = [seed0, seed - 1]
    rep = os.environ.get("MPC_CAMPAIGN_REPORT")
    if rep:
        with open(rep, "w") as f:
            f.write(json.dumps(tot) + "\n")
    assert not bad, bad[:5]
    assert tot["refined_fp64"] > 0 and tot["decisions"] > 0
