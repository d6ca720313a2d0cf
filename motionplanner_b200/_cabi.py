"""ctypes binding of include/mpc.h (csrc/libmpcheck.so).  Fails loudly if the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPC_LIB_PATH selects another build of the same library (csrc/libmpcheck_dbg.so: device asserts, `make debug`)
LIB_PATH = os.environ.get("MPC_LIB_PATH") or os.path.join(_HERE, "csrc", "libmpcheck.so")

MODE_PLANNER = 0
MODE_VALIDATOR = 1
MODE_NO_CULL = 2
MODE_NO_EARLY_EXIT = 4
MODE_COUNT_WORK = 8
MODE_NO_TIMING = 16

MAX_LIB_VERTS = 16
MAX_OBST_VERTS = 8

# mpc_query_desc
QUERY_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("c", "<f8"), ("s", "<f8"),
                        ("scene", "<i4"), ("t0", "<i4"), ("step_limit", "<i4"), ("cand_set", "<i4"),
                        ("cand_off", "<i4"), ("cand_cnt", "<i4"), ("reserved0", "<i4"), ("reserved1", "<i4")])
assert QUERY_DTYPE.itemsize == 64


class Stats(C.Structure):
    _fields_ = [("pairs_nominal", C.c_int64), ("refined_fp64", C.c_int64), ("exact_ties", C.c_int64),
                ("near_tangent", C.c_int64), ("parity_refined", C.c_int64), ("cull_tests", C.c_int64),
                ("axis_tests", C.c_int64), ("worklist_overflow", C.c_int64), ("box_skipped", C.c_int64),
                ("done_skipped", C.c_int64), ("tau", C.c_float),
                ("ms_device", C.c_float), ("ctas", C.c_int32), ("kernels", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DevInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("clock_khz", C.c_int32), ("global_mem", C.c_int64), ("l2_bytes", C.c_int32),
                ("abi_version", C.c_int32)]


class MpcError(RuntimeError):
    pass


_lib = None

EXPORTS = ["mpc_create", "mpc_destroy", "mpc_last_error", "mpc_device_info", "mpc_upload_library", "mpc_set_scenes",
           "mpc_query", "mpc_prepare_query", "mpc_run_prepared", "mpc_fetch_mask", "mpc_alloc_pinned", "mpc_free_pinned",
           "mpc_upload_trajectories", "mpc_set_guidance", "mpc_expand", "mpc_measure_fp32_peak", "mpc_submit", "mpc_wait"]


GOAL_DTYPE = np.dtype([("time_start", "<i4"), ("time_end", "<i4"), ("seg_begin", "<i4"), ("seg_end", "<i4")])


def load():
    """dlopen csrc/libmpcheck.so; raises ImportError (never falls back) when it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "motionplanner_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C motionplanner_b200/csrc`. There is no CPU fallback for the collision check." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint
    L.mpc_create.argtypes = [i32, C.POINTER(vp)]
    L.mpc_destroy.argtypes = [vp]
    L.mpc_destroy.restype = None
    L.mpc_last_error.argtypes = [vp]
    L.mpc_last_error.restype = C.c_char_p
    L.mpc_device_info.argtypes = [vp, C.POINTER(DevInfo)]
    L.mpc_upload_library.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, i32]
    L.mpc_set_scenes.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32]
    L.mpc_query.argtypes = [vp, i32, vp, vp, i32, u32, vp, i32, C.POINTER(Stats)]
    L.mpc_prepare_query.argtypes = [vp, i32, vp, vp, i32, u32]
    L.mpc_run_prepared.argtypes = [vp, i32, i32, vp, vp, C.POINTER(Stats)]
    L.mpc_fetch_mask.argtypes = [vp, vp, i32]
    L.mpc_alloc_pinned.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.mpc_free_pinned.argtypes = [vp, vp]
    L.mpc_upload_trajectories.argtypes = [vp, vp, vp, i32, i32]
    L.mpc_set_guidance.argtypes = [vp, vp, i32, vp, i32, vp, i32, C.c_double]
    L.mpc_expand.argtypes = [vp, i32, vp, vp, i32, vp, u32, vp, i32, vp, vp, C.POINTER(Stats)]
    L.mpc_submit.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, vp, vp, i32, u32]
    L.mpc_wait.argtypes = [vp, vp, i32, C.POINTER(Stats)]
    L.mpc_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double * 8)]
    for name in EXPORTS:
        if name not in ("mpc_destroy", "mpc_last_error"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def ptr(a):
    """address of a numpy array's data for a c_void_p argument (None for empty / missing).  A plain int: going through
    `a.ctypes.data_as` costs ~4 us per array, eight arrays per mpc_set_scenes"""
    return a.ctypes.data if a is not None and a.size else None
