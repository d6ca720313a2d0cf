"""Several collision-check contexts ("lanes") on ONE GPU, driven by ONE host thread.

A planning cycle through the C ABI is "new prediction in, collision bits out": `mpc_set_scenes` (host -> device copy of
the predicted occupancies + staging kernels) followed by `mpc_query` (k_check, k_refine, device -> host copy of the
result bits).  Done synchronously on one context, the copy engines idle while k_check runs and the SMs idle while the
next scene is in flight.  Independent cycles (other vehicles, other start states, the next scenario of a batch -- the
same units BASELINE.json shards across GPUs) share nothing but the read-only primitive library, so they are dealt
round-robin to `lanes` contexts: every lane has its own stream, library replica (<= 30 MB) and scene buffers.  The
driver loop uses the asynchronous pair `mpc_submit` / `mpc_wait` (include/mpc.h): it submits cycle i on lane i % lanes
and collects that lane's previous cycle just before, so up to `lanes` cycles are in flight and the device overlaps the
staging of one with the check of another.  Round 1 used one Python thread per lane; 8 ranks x 4 lane threads fought
for the interpreter lock and the cores (driver-measured 8-GPU efficiency 0.86) -- one thread per GPU needs neither.

Every decision is still made by the kernels of csrc/mpc_device.cuh; results do not depend on the lane a cycle ran on
(tests/test_gpu_parity.py::test_lanes_match_single_context).
Reference being replaced: the per-scenario loop of evaluation/runAllScenarios.py:56-99 around `collision_check`
(src/lowLevelPlannerManeuverAutomaton.py:95-125).
"""
import numpy as np

from .engine import MODE_PLANNER, CollisionEngine

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

    def run_cycles(self, cycles, mode=MODE_PLANNER, cand_idx=None, marshal_once=False, refill=None, want_stats=True,
                   mask_words_hint=None):
        """cycles: sequence of (scene, queries); scene = dict with the arguments of CollisionEngine.set_scenes, or None
        to keep what the lane staged last.  Returns [(packed mask copy, stats dict)] in the order of `cycles`.
        Cycle i runs on lane i % lanes.
        marshal_once: check / convert every distinct scene dict once for the whole call (a producer that refills the
                      same page-locked arrays and guarantees their layout) instead of per cycle.
        refill:       optional callable(i, scene, queries) invoked right before cycle i is submitted -- the place where
                      a producer writes its new prediction into the (page-locked) arrays of that cycle; its time is
                      part of the cycle.
        want_stats:   False skips the per-cycle statistics record (the second element of every result is None).
        mask_words_hint: number of result words of every cycle's batch when the caller knows it (all cycles ask for the
                      same candidate counts); counting them from the query records costs ~6 us per cycle."""
        cycles = list(cycles)
        out = [None] * len(cycles)
        L = len(self.engines)
        inflight = [None] * L
        cache = {}
        cand = None if cand_idx is None else np.ascontiguousarray(cand_idx, dtype=np.int32)
        try:
            for i, (scene, queries) in enumerate(cycles):
                lane = i % L
                eng = self.engines[lane]
                if inflight[lane] is not None:
                    out[inflight[lane]] = eng.wait(want_stats)
                    inflight[lane] = None
                if refill is not None:
                    refill(i, scene, queries)
                args = None
                if scene is not None:
                    if marshal_once:
                        args = cache.get(id(scene))
                        if args is None:
                            args = cache[id(scene)] = CollisionEngine.scene_args(*[scene[k] for k in SCENE_KEYS])
                    else:
                        args = CollisionEngine.scene_args(*[scene[k] for k in SCENE_KEYS])
                eng.submit(args, queries, mode, cand, mask_words_hint)
                inflight[lane] = i
            for lane in range(L):
                if inflight[lane] is not None:
                    out[inflight[lane]] = self.engines[lane].wait(want_stats)
                    inflight[lane] = None
        finally:
            for lane in range(L):                    # an exception above: drain what is still in flight
                if inflight[lane] is not None:
                    try:
                        self.engines[lane].wait()
                    except Exception:                # noqa: BLE001 -- the first error is the one reported
                        pass
        return out
