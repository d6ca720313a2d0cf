"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the float64 CPU oracle (oracle/oracle_c.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``) may import this.
The argument arrays are the same plain numpy arrays the product's C ABI (include/mpc.h) takes, so a parity test
feeds both sides identical bytes.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_STRICT = 1          # obstacles tested with closed-set "intersects" (validator, collisionChecker.py:32)
MODE_NO_EARLY_EXIT = 4


class OrcQuery(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("c", C.c_double), ("s", C.c_double),
                ("scene", C.c_int32), ("t0", C.c_int32), ("step_limit", C.c_int32), ("cand_set", C.c_int32),
                ("cand_off", C.c_int32), ("cand_cnt", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32)]


QUERY_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("c", "<f8"), ("s", "<f8"),
                        ("scene", "<i4"), ("t0", "<i4"), ("step_limit", "<i4"), ("cand_set", "<i4"),
                        ("cand_off", "<i4"), ("cand_cnt", "<i4"), ("reserved0", "<i4"), ("reserved1", "<i4")])


class OrcStats(C.Structure):
    _fields_ = [("pairs_decided", C.c_longlong), ("pairs_nominal", C.c_longlong),
                ("near_tangent", C.c_longlong), ("exact_calls", C.c_longlong)]


class OrcProblem(C.Structure):
    _fields_ = [("lib_verts", C.c_void_p), ("lib_nv", C.c_void_p), ("lib_tidx", C.c_void_p),
                ("P", C.c_int), ("J", C.c_int), ("Vl", C.c_int),
                ("set_off", C.c_void_p), ("set_idx", C.c_void_p), ("nsets", C.c_int),
                ("nscenes", C.c_int), ("scene_step_off", C.c_void_p), ("step_obst_off", C.c_void_p),
                ("obst", C.c_void_p), ("obst_nv", C.c_void_p), ("Vo", C.c_int),
                ("step_seg_begin", C.c_void_p), ("step_seg_end", C.c_void_p), ("segs", C.c_void_p)]


def build():
    """compile oracle/liboracle.so (gcc)"""
    subprocess.run(["make", "-C", _HERE, "--quiet"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_orient2d.argtypes = [C.c_double] * 6
        L.orc_orient2d.restype = C.c_int
        L.orc_exact_calls.restype = C.c_longlong
        L.orc_convex_collide.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_segment_meets_interior.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_point_in_soup.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int]
        for f in (L.orc_check_batch, L.orc_check_batch_inner):
            f.argtypes = [C.POINTER(OrcProblem), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                          C.POINTER(OrcStats), C.c_int]
            f.restype = C.c_int
        L.orc_expand.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.orc_expand.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def orient2d(a, b, c):
    return lib().orc_orient2d(a[0], a[1], b[0], b[1], c[0], c[1])


def convex_collide(A, B, strict):
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    return bool(lib().orc_convex_collide(_p(A), len(A), _p(B), len(B), int(strict)))


def segment_meets_interior(P, seg):
    P = np.ascontiguousarray(P, dtype=np.float64)
    seg = np.ascontiguousarray(seg, dtype=np.float64).reshape(4)
    return bool(lib().orc_segment_meets_interior(_p(P), len(P), _p(seg)))


def point_in_soup(p, segs):
    segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
    return lib().orc_point_in_soup(float(p[0]), float(p[1]), _p(segs), len(segs))


class Problem:
    """holds the arrays alive and exposes ``check(queries, cand_idx, mode)``.

    library : dict(verts [P,J,Vl,2] f8, nv [P,J] i4, tidx [P,J] i4, set_off [nsets+1] i4, set_idx i4)
    scenes  : dict(scene_step_off [S+1] i4, step_obst_off [nsteps+1] i4, obst [nobst,Vo,2] f8, obst_nv [nobst] i4,
                   step_seg_begin [nsteps] i4, step_seg_end [nsteps] i4, segs [nseg,4] f8)
    """

    def __init__(self, library, scenes):
        g = lambda d, k, dt: np.ascontiguousarray(d[k], dtype=dt) if d.get(k) is not None else np.zeros(0, dt)
        self.verts = g(library, "verts", np.float64)
        self.nv = g(library, "nv", np.int32)
        self.tidx = g(library, "tidx", np.int32)
        self.set_off = g(library, "set_off", np.int32)
        self.set_idx = g(library, "set_idx", np.int32)
        P, J, Vl = self.verts.shape[0], self.verts.shape[1], self.verts.shape[2]
        self.scene_step_off = g(scenes, "scene_step_off", np.int32)
        self.step_obst_off = g(scenes, "step_obst_off", np.int32)
        self.obst = g(scenes, "obst", np.float64)
        self.obst_nv = g(scenes, "obst_nv", np.int32)
        self.step_seg_begin = g(scenes, "step_seg_begin", np.int32)
        self.step_seg_end = g(scenes, "step_seg_end", np.int32)
        self.segs = g(scenes, "segs", np.float64)
        Vo = self.obst.shape[1] if self.obst.ndim == 3 else 4
        pb = OrcProblem()
        pb.lib_verts, pb.lib_nv, pb.lib_tidx = _p(self.verts), _p(self.nv), _p(self.tidx)
        pb.P, pb.J, pb.Vl = P, J, Vl
        pb.set_off, pb.set_idx = _p(self.set_off), _p(self.set_idx)
        pb.nsets = max(len(self.set_off) - 1, 0)
        pb.nscenes = len(self.scene_step_off) - 1
        pb.scene_step_off, pb.step_obst_off = _p(self.scene_step_off), _p(self.step_obst_off)
        pb.obst, pb.obst_nv, pb.Vo = _p(self.obst), _p(self.obst_nv), Vo
        pb.step_seg_begin, pb.step_seg_end, pb.segs = _p(self.step_seg_begin), _p(self.step_seg_end), _p(self.segs)
        self.pb = pb

    def check(self, queries, cand_idx=None, mode=0, nthreads=0, inner_parallel=False):
        """returns (collide uint8 [sum cand_cnt], stats dict)"""
        queries = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        cand = np.ascontiguousarray(cand_idx, dtype=np.int32) if cand_idx is not None else np.zeros(0, np.int32)
        n = int(queries["cand_cnt"].sum())
        out = np.zeros(n, dtype=np.uint8)
        st = OrcStats()
        f = lib().orc_check_batch_inner if inner_parallel else lib().orc_check_batch
        rc = f(C.byref(self.pb), _p(queries), len(queries), _p(cand), int(mode), _p(out), C.byref(st), int(nthreads))
        assert rc == 0
        return out, dict(pairs_decided=st.pairs_decided, pairs_nominal=st.pairs_nominal,
                         near_tangent=st.near_tangent, exact_calls=st.exact_calls)


GOAL_DTYPE = np.dtype([("time_start", "<i4"), ("time_end", "<i4"), ("seg_begin", "<i4"), ("seg_end", "<i4")])


def expand(states, nstates, ref_xy, goals, goal_segs, b, set_off, set_idx, queries, cand_idx, parent):
    """cost and goal-hit step of every child (orc_expand): returns (cost float64 [n], goal_step int32 [n], -1 = none)"""
    states = np.ascontiguousarray(states, np.float64)
    nstates = np.ascontiguousarray(nstates, np.int32)
    ref_xy = np.ascontiguousarray(ref_xy, np.float64)
    goals = np.ascontiguousarray(goals, GOAL_DTYPE)
    goal_segs = np.ascontiguousarray(goal_segs if goal_segs is not None else np.zeros((0, 4)), np.float64)
    set_off = np.ascontiguousarray(set_off, np.int32)
    set_idx = np.ascontiguousarray(set_idx, np.int32)
    queries = np.ascontiguousarray(queries, QUERY_DTYPE)
    cand = np.ascontiguousarray(cand_idx, np.int32) if cand_idx is not None else np.zeros(0, np.int32)
    parent = np.ascontiguousarray(parent, np.float64).reshape(len(queries), 2)
    n = int(queries["cand_cnt"].sum())
    cost = np.zeros(n)
    goal = np.zeros(n, np.int32)
    rc = lib().orc_expand(_p(states), _p(nstates), states.shape[0], states.shape[1], _p(ref_xy), ref_xy.shape[1],
                          _p(goals), len(goals), _p(goal_segs), float(b), _p(set_off), _p(set_idx), _p(queries),
                          len(queries), _p(cand), _p(parent), _p(cost), _p(goal))
    assert rc == 0, rc
    return cost, goal
