"""Several collision-check contexts on ONE GPU, each driven by its own host thread.

A planning cycle through the C ABI is `mpc_set_scenes` (host -> device copy of the predicted occupancies + staging
kernels) followed by `mpc_query` (k_check, k_refine, device -> host copy of the result bits); both calls return only
when their stream has drained, so a single context leaves the copy engines idle while k_check runs and the SMs idle
while the next scene is in flight.  Independent cycles (other vehicles, other start states, the next scenario of a batch
-- the same units BASELINE.json shards across GPUs) share nothing but the read-only primitive library, so they are dealt
round-robin to `lanes` contexts: every lane has its own stream, library replica (<= 30 MB) and scene buffers, and the
device overlaps the staging of one lane with the check of another.  ctypes releases the GIL during the calls.

Every decision is still made by the kernels of csrc/mpc_device.cuh; results do not depend on the lane a cycle ran on
(tests/test_gpu_parity.py::test_lanes_match_single_context).
Reference being replaced: the per-scenario loop of evaluation/runAllScenarios.py:56-99 around `collision_check`
(src/lowLevelPlannerManeuverAutomaton.py:95-125).
"""
import ctypes as C
import threading

import numpy as np

from .engine import MODE_PLANNER, QUERY_DTYPE, CollisionEngine, mask_words

__all__ = ["CollisionLanes"]

SCENE_KEYS = ("scene_step_off", "origins", "step_obst_off", "obst", "obst_nv", "step_seg_begin", "step_seg_end", "segs")


class CollisionLanes:
    def __init__(self, device=0, lanes=2):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.engines = []
        try:
            for _ in range(lanes):
                self.engines.append(CollisionEngine(device))
        except Exception:
            self.close()
            raise

    def __len__(self):
        return len(self.engines)

    def close(self):
        for e in self.engines:
            e.close()

    def upload_library(self, verts, nv, tidx, set_off=None, set_idx=None):
        for e in self.engines:
            e.upload_library(verts, nv, tidx, set_off, set_idx)

    def pinned_like(self, array):
        """page-locked copy; page-locked memory is visible to every context of the process"""
        return self.engines[0].pinned_like(array)

    def run_cycles(self, cycles, mode=MODE_PLANNER, cand_idx=None):
        """cycles: sequence of (scene, queries); scene = dict with the arguments of CollisionEngine.set_scenes, or None
        to keep what the lane staged last.  Returns [(packed mask copy, stats dict)] in the order of `cycles`.
        Cycle i runs on lane i % lanes; an exception in any lane is re-raised here after all lanes have stopped."""
        cycles = list(cycles)
        out = [None] * len(cycles)
        errors = []
        stop = threading.Event()
        # marshal every distinct scene / query array once (a producer that refills the same page-locked arrays hands
        # the same objects in again): the per-cycle interpreter time under the GIL is what the lanes compete for
        scene_args, query_args = {}, {}
        cand = None if cand_idx is None else np.ascontiguousarray(cand_idx, dtype=np.int32)
        for scene, queries in cycles:
            if scene is not None and id(scene) not in scene_args:
                scene_args[id(scene)] = CollisionEngine.scene_args(*[scene[k] for k in SCENE_KEYS])
            if id(queries) not in query_args:
                q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
                query_args[id(queries)] = (q, q.ctypes.data_as(C.c_void_p), mask_words(q))

        def work(lane):
            eng = self.engines[lane]
            try:
                for i in range(lane, len(cycles), len(self.engines)):
                    if stop.is_set():
                        return
                    scene, queries = cycles[i]
                    if scene is not None:
                        eng.set_scenes_args(scene_args[id(scene)])
                    q, qptr, nw = query_args[id(queries)]
                    if cand is None:
                        mask, st = eng.query_fast(q, mode, want_stats=True, nw=nw, qptr=qptr)
                        out[i] = (mask[:nw].copy(), {k: getattr(st, k) for k, _ in st._fields_})
                    else:
                        out[i] = eng.query_mask(q, cand, mode)
            except BaseException as exc:          # noqa: BLE001 -- handed to the caller below
                errors.append(exc)
                stop.set()

        if len(self.engines) == 1 or len(cycles) <= 1:
            work(0)
        else:
            threads = [threading.Thread(target=work, args=(k,), name="mpc-lane-%d" % k)
                       for k in range(min(len(self.engines), len(cycles)))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        return out
