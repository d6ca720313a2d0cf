"""The kernels' index invariants (work-queue indices, compaction ranks, tile extents, vertex counts), compiled in as
device asserts by `make -C motionplanner_b200/csrc debug` (libmpcheck_dbg.so, -DMPC_DEBUG_ASSERTS): a set of parity
problems that reaches every regime (multi-phase tiles, compaction, scout launches, warps over obstacles, single-kernel
tail, general polygons) runs through that build in a child process; a violated invariant traps the kernel and fails
the call.  compute-sanitizer is not available on the measurement pool; skipped when the debug
library has not been built."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "motionplanner_b200", "csrc", "libmpcheck_dbg.so")

CHILD = r'''
import numpy as np
from motionplanner_b200 import workloads as W
from motionplanner_b200.engine import MODE_PLANNER, MODE_VALIDATOR, MODE_NO_CULL, MODE_NO_EARLY_EXIT, CollisionEngine
from oracle import oracle as orc
eng = CollisionEngine(0)
n = 0
def run(lib, sc, org, q, cand=None, modes=((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT))):
    global n
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
                   sc["step_seg_end"], sc["segs"])
    pb = orc.Problem(lib, sc)
    for mode, omode in modes:
        got, st = eng.query(q, cand, mode)
        want, _ = pb.check(q, cand, omode | orc.MODE_NO_EARLY_EXIT)
        assert np.array_equal(got, want), (mode, int((got != want).sum()))
        n += len(got)
for seed in range(6):
    run(*W.adversarial_problem(seed))
run(*W.adversarial_problem(40, lib_verts=9, obst_verts=7))
run(*W.adversarial_problem(41, lib_verts=16, obst_verts=8, P=64))
run(*W.adversarial_problem(42, max_obst=400, P=40, J=9, nsteps=10))                    # multi-phase tiles
run(*W.adversarial_problem(43, P=1300, J=4, nsteps=6, max_obst=90, nqueries=4))        # sorted sets, block culling
lib, sc, org, q = W.c4_problem(P=4096, O_lanes=8, slots=16, T=12, nqueries=32)         # scout launch + compaction
run(lib, sc, org, q)
lib, sc, org, q = W.c4_problem(P=1024, O_lanes=8, slots=96, T=10, nqueries=24)         # early return across slots
run(lib, sc, org, q)
lib, sc, org, q = W.mixed_traffic_problem(P=2048, T=20, nqueries=8)
run(lib, sc, org, q)
lib, sc, org, q = W.c4_problem(P=512, O_lanes=8, slots=8, T=6, nqueries=2)
run(lib, sc, org, q, modes=((MODE_NO_CULL | MODE_NO_EARLY_EXIT, 0),))                 # k_check_full
eng.close()
print("debug-assert build: %d decisions identical to the oracle, no invariant violated" % n)
'''


@pytest.mark.skipif(not os.path.exists(DBG), reason="libmpcheck_dbg.so not built (make -C motionplanner_b200/csrc debug)")
def test_parity_problems_through_the_assert_build():
    env = dict(os.environ, MPC_LIB_PATH=DBG, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-400:], r.stderr[-2000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "no invariant violated" in r.stdout
