"""Host-side handle of the device-resident collision checker (one per GPU).

Thin numpy <-> C-ABI glue; every geometric decision is made by the CUDA kernels in csrc/mpc_device.cuh.
Reference being replaced: the per-occupancy shapely loop of ``collision_check``
(src/lowLevelPlannerManeuverAutomaton.py:95-125) and ``collisionChecker`` (auxiliary/collisionChecker.py:6-35).
"""
import ctypes as C
import weakref

import numpy as np

from . import _cabi
from ._cabi import (MODE_COUNT_WORK, MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_PLANNER, MODE_VALIDATOR, QUERY_DTYPE,
                    MpcError)

__all__ = ["CollisionEngine", "QUERY_DTYPE", "MODE_PLANNER", "MODE_VALIDATOR", "MODE_NO_CULL", "MODE_NO_EARLY_EXIT",
           "MODE_COUNT_WORK", "MpcError", "make_queries", "unpack_mask"]


_I4, _F8 = np.dtype(np.int32), np.dtype(np.float64)

# data address of an ndarray, remembered per array OBJECT: a producer refills the same (page-locked) arrays every cycle,
# and `a.ctypes.data` costs 1.1 us per array, nine arrays per cycle.  The entry holds a weak reference (a dead array's
# id may be reused: the reference then no longer resolves to the caller's array) and the shape (an in-place resize is
# the only way a live array's buffer moves); nothing is kept alive by the table.
_PTRS = {}


def _addr(a):
    if a is None or a.size == 0:
        return None
    ent = _PTRS.get(id(a))
    if ent is not None and ent[0]() is a and ent[1] == a.shape:
        return ent[2]
    if len(_PTRS) >= 256:
        _PTRS.clear()
    p = a.ctypes.data
    try:
        _PTRS[id(a)] = (weakref.ref(a), a.shape, p)
    except TypeError:                 # an ndarray subclass without weak references
        pass
    return p


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def make_queries(n):
    q = np.zeros(n, dtype=QUERY_DTYPE)
    q["c"] = 1.0
    q["step_limit"] = 2 ** 30
    return q


def mask_words(queries):
    return int(((queries["cand_cnt"].astype(np.int64) + 31) // 32).sum())


def unpack_mask(mask, queries):
    """packed device mask -> flat uint8 array (1 = candidate collides) in query order"""
    out = np.zeros(int(queries["cand_cnt"].sum()), dtype=np.uint8)
    w = o = 0
    for cnt in queries["cand_cnt"]:
        nw = (int(cnt) + 31) // 32
        bits = np.unpackbits(mask[w:w + nw].view(np.uint8), bitorder="little")[:cnt]
        out[o:o + cnt] = bits
        w += nw
        o += cnt
    return out


class CollisionEngine:
    """one context = one GPU + one stream (see include/mpc.h)"""

    def __init__(self, device=0):
        self._L = _cabi.load()
        self._ctx = C.c_void_p()
        rc = self._L.mpc_create(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self._L.mpc_last_error(self._ctx).decode() if self._ctx else "mpc_create failed"
            if self._ctx:
                self._L.mpc_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise MpcError("mpc_create(device=%d) -> %d: %s" % (device, rc, msg))
        self.device = int(device)
        self._keep = {}
        self._staged_library = self._scene_owner = None

    def close(self):
        if getattr(self, "_ctx", None):
            for ptr in getattr(self, '_pinned', []):
                self._L.mpc_free_pinned(self._ctx, ptr)
            self._pinned = []
            self._L.mpc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise MpcError("%s -> %d: %s" % (what, rc, self._L.mpc_last_error(self._ctx).decode()))

    def device_info(self):
        info = _cabi.DevInfo()
        self._check(self._L.mpc_device_info(self._ctx, C.byref(info)), "mpc_device_info")
        return dict(name=info.name.decode(), sm_count=info.sm_count, cc=(info.cc_major, info.cc_minor),
                    clock_khz=info.clock_khz, global_mem=info.global_mem, l2_bytes=info.l2_bytes,
                    abi_version=info.abi_version)

    def measure_fp32_peak(self):
        """measured float32 peaks of this device in TFLOP/s: scalar FFMA, packed FFMA2, the full-test instruction mix"""
        out = (C.c_double * 8)()
        self._check(self._L.mpc_measure_fp32_peak(self._ctx, C.byref(out)), "mpc_measure_fp32_peak")
        return dict(ffma_tflops=out[0], ffma2_tflops=out[1], sat_mix_ffma2_tflops=out[2], sat_mix_ffma_tflops=out[3],
                    fmnmx3_tflops=out[4], fmnmx_tflops=out[5], ffma_fmnmx_2to1_tflops=out[6], sm_mhz_seen=out[7])

    def pinned_like(self, array):
        """copy of `array` in page-locked host memory (mpc_alloc_pinned): host->device copies from it are real DMA"""
        a = np.ascontiguousarray(array)
        ptr = C.c_void_p()
        self._check(self._L.mpc_alloc_pinned(self._ctx, max(a.nbytes, 16), C.byref(ptr)), "mpc_alloc_pinned")
        buf = (C.c_char * max(a.nbytes, 16)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
        out[...] = a
        self._pinned = getattr(self, '_pinned', [])
        self._pinned.append(ptr)
        return out

    # -- library -------------------------------------------------------------------------------------------------
    def upload_library(self, verts, nv, tidx, set_off=None, set_idx=None):
        """verts [P,J,V,2] f8, nv [P,J] i4 (0 = unused slot), tidx [P,J] i4; optional candidate sets (CSR)"""
        verts = _arr(verts, np.float64)
        nv = _arr(nv, np.int32)
        tidx = _arr(tidx, np.int32)
        P, J, V = verts.shape[0], verts.shape[1], verts.shape[2]
        assert verts.shape == (P, J, V, 2) and nv.shape == (P, J) and tidx.shape == (P, J)
        nsets = 0
        if set_off is not None:
            set_off = _arr(set_off, np.int32)
            set_idx = _arr(set_idx, np.int32)
            nsets = len(set_off) - 1
        self._staged_library = None          # whoever staged through a cache key sets it again after this call
        self._check(self._L.mpc_upload_library(self._ctx, _cabi.ptr(verts), _cabi.ptr(nv), _cabi.ptr(tidx), P, J, V,
                                               _cabi.ptr(set_off) if nsets else None,
                                               _cabi.ptr(set_idx) if nsets else None, nsets), "mpc_upload_library")
        self.P, self.J = P, J

    # -- scenes --------------------------------------------------------------------------------------------------
    def set_scenes(self, scene_step_off, origins, step_obst_off, obst, obst_nv, step_seg_begin, step_seg_end, segs):
        self.set_scenes_args(self.scene_args(scene_step_off, origins, step_obst_off, obst, obst_nv, step_seg_begin,
                                             step_seg_end, segs))

    @staticmethod
    def scene_args(scene_step_off, origins, step_obst_off, obst, obst_nv, step_seg_begin, step_seg_end, segs):
        """checks the arrays and marshals them into the argument tuple of mpc_set_scenes / mpc_submit (the arrays are
        kept alive by the tuple).  Arrays that already have the dtype and layout of include/mpc.h are passed as they are
        (a producer refilling page-locked buffers): ~10 us; anything else is converted first."""
        def fit(a, dt, empty):
            if a is None:
                return np.zeros(empty, dt)
            if type(a) is np.ndarray and a.dtype == dt and a.flags.c_contiguous:
                return a
            return np.ascontiguousarray(a, dtype=dt)

        scene_step_off = fit(scene_step_off, _I4, 0)
        nscenes = len(scene_step_off) - 1
        origins = fit(origins, _F8, (0, 2))
        step_obst_off = fit(step_obst_off, _I4, 0)
        obst = fit(obst, _F8, (0, 4, 2))
        if obst.ndim != 3:
            obst = obst.reshape(-1, 4, 2)
        obst_nv = fit(obst_nv, _I4, 0)
        step_seg_begin = fit(step_seg_begin, _I4, 0)
        step_seg_end = fit(step_seg_end, _I4, 0)
        segs = fit(segs, _F8, (0, 4))
        nsteps = int(scene_step_off[-1])
        if not (origins.size == 2 * nscenes and len(step_obst_off) == nsteps + 1 and
                len(obst_nv) == obst.shape[0] == int(step_obst_off[-1]) and
                len(step_seg_begin) == len(step_seg_end) == nsteps and segs.size % 4 == 0):
            raise AssertionError("inconsistent scene arrays: %d scenes, %d steps, step_obst_off %d, obst %s, obst_nv %d, "
                                 "segment ranges %d / %d, segs %s" % (nscenes, nsteps, len(step_obst_off), obst.shape,
                                                                      len(obst_nv), len(step_seg_begin), len(step_seg_end), segs.shape))
        keep = (scene_step_off, origins, step_obst_off, obst, obst_nv, step_seg_begin, step_seg_end, segs)
        ptr = _addr
        return (nscenes, ptr(scene_step_off), ptr(origins), ptr(step_obst_off), ptr(obst),
                ptr(obst_nv), obst.shape[1], ptr(step_seg_begin), ptr(step_seg_end),
                ptr(segs), segs.size // 4, keep)

    def set_scenes_args(self, args):
        self._scene_owner = None             # whoever staged through a PrimitiveChecker sets it again after this call
        rc = self._L.mpc_set_scenes(self._ctx, *args[:11])
        if rc != 0:
            self._check(rc, "mpc_set_scenes")

    # -- queries -------------------------------------------------------------------------------------------------
    def query_mask(self, queries, cand_idx=None, mode=MODE_PLANNER):
        """returns (packed uint32 mask, stats dict): bit = 1 iff the candidate collides"""
        queries = _arr(queries, QUERY_DTYPE)
        cand = _arr(cand_idx, np.int32) if cand_idx is not None else np.zeros(0, np.int32)
        nw = mask_words(queries)
        mask = np.zeros(max(nw, 1), dtype=np.uint32)
        st = _cabi.Stats()
        self._check(self._L.mpc_query(self._ctx, len(queries), _cabi.ptr(queries), _cabi.ptr(cand), len(cand),
                                      int(mode), _cabi.ptr(mask), nw, C.byref(st)), "mpc_query")
        return mask[:nw], st.as_dict()

    def query_fast(self, queries, mode=MODE_PLANNER, want_stats=False, nw=None, qptr=None):
        """low-overhead path for the planner loop: `queries` is a ready QUERY_DTYPE array that references staged
        candidate sets only; returns the packed mask (a view into a reused buffer) and, optionally, the stats struct.
        `nw` (number of mask words of the batch) and `qptr` (ctypes pointer to `queries`) may be handed in by a caller
        that knows them: computing them costs 6 + 4 us of the ~50 us a planner-sized query takes"""
        if nw is None:
            nw = mask_words(queries)
        buf = getattr(self, '_mask_buf', None)
        if buf is None or len(buf) < nw:
            buf = self._mask_buf = np.zeros(max(nw, 64), dtype=np.uint32)
            self._mask_ptr = buf.ctypes.data_as(C.c_void_p)
        st = self._stats_buf = getattr(self, '_stats_buf', None) or _cabi.Stats()
        rc = self._L.mpc_query(self._ctx, len(queries), qptr if qptr is not None else queries.ctypes.data_as(C.c_void_p),
                               None, 0, int(mode),
                               self._mask_ptr, nw, C.byref(st) if want_stats else None)
        if rc != 0:
            self._check(rc, "mpc_query")
        return buf[:nw], (st if want_stats else None)

    # -- asynchronous cycle (mpc_submit / mpc_wait) ------------------------------------------------------------------
    def submit(self, scene_args, queries, mode=MODE_PLANNER, cand_idx=None, nw=None):
        """stage the scenes (`scene_args` = tuple of CollisionEngine.scene_args(...), or None to keep what is staged)
        and enqueue the query batch without waiting for the device; `wait()` returns its result.  Every array involved
        must stay alive and unchanged until then (this object keeps references)."""
        if not (type(queries) is np.ndarray and queries.dtype == QUERY_DTYPE and queries.flags.c_contiguous):
            queries = _arr(queries, QUERY_DTYPE)
        cand = _arr(cand_idx, np.int32) if cand_idx is not None else None
        if nw is None:
            nw = mask_words(queries)
        sa = scene_args[:11] if scene_args is not None else (0, None, None, None, None, None, 4, None, None, None, 0)
        rc = self._L.mpc_submit(self._ctx, *sa, len(queries), _addr(queries), _cabi.ptr(cand),
                                0 if cand is None else len(cand), int(mode))
        if rc != 0:
            self._check(rc, "mpc_submit")
        if scene_args is not None:
            self._scene_owner = None
        self._inflight = (scene_args, queries, cand, nw)

    def wait(self, want_stats=True):
        """result of the submitted batch: (packed mask [copy], stats dict or None)"""
        if getattr(self, "_inflight", None) is None:
            raise MpcError("wait() without a submitted batch")
        nw = self._inflight[3]
        mask = np.empty(max(nw, 1), dtype=np.uint32)
        st = _cabi.Stats() if want_stats else None
        rc = self._L.mpc_wait(self._ctx, _cabi.ptr(mask), nw, C.byref(st) if want_stats else None)
        self._inflight = None
        if rc != 0:
            self._check(rc, "mpc_wait")
        return mask[:nw], (st.as_dict() if want_stats else None)

    def query(self, queries, cand_idx=None, mode=MODE_PLANNER):
        """returns (uint8 [sum cand_cnt] 1 = collides, stats dict)"""
        queries = _arr(queries, QUERY_DTYPE)
        mask, st = self.query_mask(queries, cand_idx, mode)
        return unpack_mask(mask, queries), st

    # -- primitive-selection step ---------------------------------------------------------------------
    def upload_trajectories(self, states, nstates):
        """states [P, L, 4] = primitive.x transposed, nstates [P] = primitive.x.shape[1]; once per automaton"""
        states = _arr(states, np.float64)
        nstates = _arr(nstates, np.int32)
        assert states.ndim == 3 and states.shape[2] == 4 and nstates.shape == (states.shape[0],)
        self._check(self._L.mpc_upload_trajectories(self._ctx, _cabi.ptr(states), _cabi.ptr(nstates), states.shape[0],
                                                    states.shape[1]), "mpc_upload_trajectories")

    def set_guidance(self, ref_xy, goals, goal_segs, b):
        """ref_xy [2, M] = ref_traj[0:2]; goals: GOAL_DTYPE records; goal_segs [n, 4]; b = param['b']"""
        ref_xy = _arr(ref_xy, np.float64)
        goals = _arr(goals, _cabi.GOAL_DTYPE)
        goal_segs = _arr(goal_segs if goal_segs is not None else np.zeros((0, 4)), np.float64).reshape(-1, 4)
        assert ref_xy.ndim == 2 and ref_xy.shape[0] == 2
        self._check(self._L.mpc_set_guidance(self._ctx, _cabi.ptr(ref_xy), ref_xy.shape[1], _cabi.ptr(goals), len(goals),
                                             _cabi.ptr(goal_segs), len(goal_segs), float(b)), "mpc_set_guidance")

    def expand(self, queries, parent, cand_idx=None, mode=MODE_PLANNER):
        """collision bits + child cost + goal step of every candidate: (uint8 [n], float64 [n], int32 [n], stats)"""
        queries = _arr(queries, QUERY_DTYPE)
        cand = _arr(cand_idx, np.int32) if cand_idx is not None else np.zeros(0, np.int32)
        parent = _arr(parent, np.float64).reshape(len(queries), 2)
        nw = mask_words(queries)
        n = int(queries['cand_cnt'].sum())
        mask = np.zeros(max(nw, 1), dtype=np.uint32)
        cost = np.zeros(max(n, 1))
        goal = np.zeros(max(n, 1), np.int32)
        st = _cabi.Stats()
        self._check(self._L.mpc_expand(self._ctx, len(queries), _cabi.ptr(queries), _cabi.ptr(cand), len(cand),
                                       _cabi.ptr(parent), int(mode), _cabi.ptr(mask), nw, _cabi.ptr(cost),
                                       _cabi.ptr(goal), C.byref(st)), "mpc_expand")
        return unpack_mask(mask[:nw], queries), cost[:n], goal[:n], st.as_dict()

    # -- measurement (bench.py) ------------------------------------------------------------------------------------
    def prepare(self, queries, cand_idx=None, mode=MODE_PLANNER):
        queries = _arr(queries, QUERY_DTYPE)
        cand = _arr(cand_idx, np.int32) if cand_idx is not None else np.zeros(0, np.int32)
        self._check(self._L.mpc_prepare_query(self._ctx, len(queries), _cabi.ptr(queries), _cabi.ptr(cand), len(cand),
                                              int(mode)), "mpc_prepare_query")
        self._prepared = queries

    def run_prepared(self, iters, flush_l2=True):
        """returns (ms_total [iters], ms_main [iters], stats) -- device times from CUDA events around the kernels"""
        ms_total = np.zeros(max(iters, 1), np.float32)
        ms_main = np.zeros(max(iters, 1), np.float32)
        st = _cabi.Stats()
        self._check(self._L.mpc_run_prepared(self._ctx, int(iters), int(bool(flush_l2)), _cabi.ptr(ms_total),
                                             _cabi.ptr(ms_main), C.byref(st)), "mpc_run_prepared")
        return ms_total[:iters], ms_main[:iters], st.as_dict()

    def fetch(self):
        nw = mask_words(self._prepared)
        mask = np.zeros(max(nw, 1), dtype=np.uint32)
        self._check(self._L.mpc_fetch_mask(self._ctx, _cabi.ptr(mask), nw), "mpc_fetch_mask")
        return unpack_mask(mask[:nw], self._prepared)
