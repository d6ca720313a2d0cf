"""Device-resident collision checker behind the reference's Python call surface.

Glue between Python objects (maneuver automaton, shapely-like free-space polygons, CommonRoad-like scenarios) and
the flat arrays of include/mpc.h.  Nothing geometric is decided here: this module flattens inputs, stages them once
and turns a batch of (parent state, candidate primitives) into one ``mpc_query``.

Reference call sites served:
  collision_check(node, primitive, space, param, timepoint)       src/lowLevelPlannerManeuverAutomaton.py:95-125
  collision_check(node, primitive, space, goal_set, param)        src/maneuverAutomatonPlannerStandalone.py:196-226
  collisionChecker(scenario, traj, param)                         auxiliary/collisionChecker.py:6-35
"""
import ctypes as C

import numpy as np

from ._cabi import MODE_NO_TIMING
from .engine import MODE_PLANNER, MODE_VALIDATOR, CollisionEngine, make_queries
from .geometry import ring_segments, rings_of

_ENGINES = {}


def get_engine(role='planner', device=None):
    """one engine (context + stream + device buffers) per role and device; created on first use"""
    import os
    if device is None:
        device = int(os.environ.get('MPC_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    key = (role, device)
    if key not in _ENGINES:
        _ENGINES[key] = CollisionEngine(device)
    return _ENGINES[key]


def close_engines():
    for e in _ENGINES.values():
        e.close()
    _ENGINES.clear()


# ---------------------------------------------------------------------------------------------------------------------
# obstacle shapes -> convex pieces
# ---------------------------------------------------------------------------------------------------------------------

def _is_convex(v):
    """all turns in one direction AND one revolution in total (a self-wrapping ring also turns one way throughout)"""
    d1 = np.roll(v, -1, axis=0) - v
    d2 = np.roll(v, -2, axis=0) - np.roll(v, -1, axis=0)
    c = d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]
    if not (np.all(c >= 0) or np.all(c <= 0)):
        return False
    if not np.any(c):
        return True           # all vertices collinear: a degenerate (zero-area) piece, dropped by the staging kernel
    turn = np.arctan2(c, d1[:, 0] * d2[:, 0] + d1[:, 1] * d2[:, 1]).sum()
    return bool(abs(abs(turn) - 2 * np.pi) < 1e-6)


def _area2(v):
    return float(np.sum(v[:, 0] * np.roll(v[:, 1], -1) - np.roll(v[:, 0], -1) * v[:, 1]))


def _triangulate(v):
    """ear clipping of a simple polygon (open vertex array) -> list of triangles.  The triangles must cover the
    polygon: a remainder that cannot be clipped (self-intersecting or numerically degenerate input) raises instead of
    being dropped -- an obstacle part that is silently not staged would be a missed collision."""
    v = np.asarray(v, dtype=np.float64)
    area = 0.5 * _area2(v)
    idx = list(range(len(v))) if area > 0 else list(range(len(v)))[::-1]
    tris = []

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    while len(idx) > 3:
        n = len(idx)
        for k in range(n):
            i0, i1, i2 = idx[k - 1], idx[k], idx[(k + 1) % n]
            a, b, c = v[i0], v[i1], v[i2]
            if cross(a, b, c) <= 0:
                continue
            ok = True
            for j in idx:
                if j in (i0, i1, i2):
                    continue
                p = v[j]
                if cross(a, b, p) >= 0 and cross(b, c, p) >= 0 and cross(c, a, p) >= 0:
                    ok = False
                    break
            if ok:
                tris.append(np.array([a, b, c]))
                idx.pop(k)
                break
        else:
            # no ear left: collinear runs are harmless (zero area), anything else is not a simple polygon
            rest = v[idx]
            if abs(_area2(rest)) > 1e-9 * max(abs(area), 1e-300):
                raise ValueError("obstacle polygon cannot be triangulated (not a simple polygon?): %d vertices left "
                                 "with area %.3g of %.3g" % (len(idx), 0.5 * abs(_area2(rest)), abs(area)))
            idx = []
    if len(idx) == 3:
        tris.append(v[idx])
    covered = sum(0.5 * abs(_area2(t)) for t in tris)
    if abs(covered - abs(area)) > 1e-9 * max(abs(area), 1e-300):
        raise ValueError("triangulation covers %.17g of the polygon's area %.17g" % (covered, abs(area)))
    return tris


def _circle_ring(center, radius, quad_segs=16):
    """the polygon the reference tests for a commonroad Circle: `shape.shapely_object` is Point(center).buffer(radius),
    a ring of 4*quad_segs vertices ON the circle, clockwise from angle 0 (shapely's default quad_segs=16; GEOS
    evaluates the same angles, last-place differences fall under the 1e-9 m policy)"""
    k = np.arange(4 * quad_segs)
    ang = -k * (np.pi / (2 * quad_segs))
    return np.stack([center[0] + radius * np.cos(ang), center[1] + radius * np.sin(ang)], -1)


def convex_pieces(shape, max_verts=8):
    """convex pieces (open vertex arrays) of a CommonRoad-like shape.  A ShapeGroup contributes its members
    (reference get_shapely_object unions them, auxiliary/collisionChecker.py:37-47; testing the members one by one is
    equivalent), a non-convex polygon its triangles, a circle the fans of its 64-gon."""
    if hasattr(shape, 'shapes'):
        out = []
        for s in shape.shapes:
            out.extend(convex_pieces(s, max_verts))
        return out
    if not hasattr(shape, 'vertices'):
        if hasattr(shape, 'radius') and hasattr(shape, 'center'):
            ring = _circle_ring(np.asarray(shape.center, dtype=np.float64), float(shape.radius))
            c = np.asarray(shape.center, dtype=np.float64)[None]
            step = max_verts - 2                   # centre + (step + 1) consecutive ring vertices per piece
            return [np.concatenate([c, ring[np.arange(i, min(i + step, len(ring)) + 1) % len(ring)]])
                    for i in range(0, len(ring), step)]
        raise NotImplementedError("shape %r has neither vertices nor centre/radius" % type(shape).__name__)
    v = np.asarray(shape.vertices, dtype=np.float64)
    if len(v) > 1 and np.array_equal(v[0], v[-1]):
        v = v[:-1]
    keep = np.ones(len(v), bool)
    keep[1:] = np.any(v[1:] != v[:-1], axis=1)
    v = v[keep]
    if len(v) < 3:
        return []
    if len(v) <= max_verts and _is_convex(v):
        return [v]
    return _triangulate(v)


def scenario_obstacle_arrays(scenario, nsteps, max_verts=8):
    """per-step convex obstacle pieces of a scenario: (step_obst_off [nsteps+1], obst [n,V,2], obst_nv [n]).
    ``obs.occupancy_at_time(i)`` exactly as the reference loops (auxiliary/collisionChecker.py:25-30)."""
    off, pieces = [0], []
    for i in range(nsteps):
        for obs in scenario.obstacles:
            occ = obs.occupancy_at_time(i)
            if occ is not None:
                pieces.extend(convex_pieces(occ.shape, max_verts))
        off.append(len(pieces))
    V = max([4] + [len(p) for p in pieces])
    obst = np.zeros((len(pieces), V, 2))
    nv = np.zeros(len(pieces), np.int32)
    for k, p in enumerate(pieces):
        obst[k, :len(p)] = p
        nv[k] = len(p)
    return np.array(off, np.int32), obst, nv


# ---------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------

class SceneSet:
    """flat arrays of one or more scenes (include/mpc.h: mpc_set_scenes)"""

    def __init__(self):
        self.scene_step_off = [0]
        self.origins = []
        self.step_obst_off = [0]
        self.obst, self.obst_nv = [], []
        self.seg_begin, self.seg_end, self.segs = [], [], []
        self._nseg = 0
        self._nobst = 0

    def add_scene_from_spaces(self, spaces):
        """planner mode: one region per time step given as shapely-like geometry (space_all[t]); no obstacles --
        they are already carved out of the regions (src/highLevelPlanner.py:1828-2003)"""
        first = None
        for g in spaces:
            rings = rings_of(g)
            s = ring_segments(rings)
            if first is None and len(s):
                first = s[0, 0:2].copy()
            self.seg_begin.append(self._nseg)
            self.seg_end.append(self._nseg + len(s))
            self.segs.append(s)
            self._nseg += len(s)
            self.step_obst_off.append(self._nobst)
        self.scene_step_off.append(self.scene_step_off[-1] + len(spaces))
        self.origins.append(first if first is not None else np.zeros(2))
        return len(self.origins) - 1

    def add_scene_explicit(self, nsteps, step_obst_off, obst, obst_nv, region_segments=None, origin=None):
        """explicit mode: obstacles per step plus one static region (road) for all steps, or none"""
        step_obst_off = np.asarray(step_obst_off, np.int64)
        self.step_obst_off.extend((step_obst_off[1:] + self._nobst).tolist())
        self.obst.append(np.asarray(obst, np.float64))
        self.obst_nv.append(np.asarray(obst_nv, np.int32))
        self._nobst += len(obst_nv)
        if region_segments is not None:
            s = np.asarray(region_segments, np.float64).reshape(-1, 4)
            b, e = self._nseg, self._nseg + len(s)
            self.segs.append(s)
            self._nseg += len(s)
        else:
            b = e = -1
        self.seg_begin.extend([b] * nsteps)
        self.seg_end.extend([e] * nsteps)
        self.scene_step_off.append(self.scene_step_off[-1] + nsteps)
        if origin is None:
            origin = obst[0, 0] if len(obst) else (region_segments[0, 0:2] if region_segments is not None and len(region_segments) else np.zeros(2))
        self.origins.append(np.asarray(origin, np.float64))
        return len(self.origins) - 1

    def arrays(self):
        V = max([4] + [o.shape[1] for o in self.obst if len(o)])
        obst = [np.pad(o, ((0, 0), (0, V - o.shape[1]), (0, 0))) for o in self.obst if len(o)]
        return dict(scene_step_off=np.array(self.scene_step_off, np.int32), origins=np.array(self.origins).reshape(-1, 2),
                    step_obst_off=np.array(self.step_obst_off, np.int32),
                    obst=np.concatenate(obst) if obst else np.zeros((0, V, 2)),
                    obst_nv=np.concatenate(self.obst_nv) if self.obst_nv else np.zeros(0, np.int32),
                    step_seg_begin=np.array(self.seg_begin, np.int32), step_seg_end=np.array(self.seg_end, np.int32),
                    segs=np.concatenate(self.segs) if self.segs else np.zeros((0, 4)))

    def stage(self, engine):
        a = self.arrays()
        engine.set_scenes(a['scene_step_off'], a['origins'], a['step_obst_off'], a['obst'], a['obst_nv'],
                          a['step_seg_begin'], a['step_seg_end'], a['segs'])
        return a


# ---------------------------------------------------------------------------------------------------------------------
# the checker used by the planners
# ---------------------------------------------------------------------------------------------------------------------

class PrimitiveChecker:
    """stages a maneuver automaton and a scene set, answers batched "which children collide" questions.

    ``backend`` is a CollisionEngine (the product) -- tests may pass any object with the same
    upload_library / set_scenes / query interface to cross-check decisions."""

    def __init__(self, MA, time_step, backend=None):
        self.MA = MA
        self.layout = MA.device_layout(time_step)
        self.backend = backend if backend is not None else get_engine('planner')
        lay = self.layout
        # the library is staged once per (engine, automaton, time step): replanning only re-stages the scene
        self._lib_key = (id(MA), float(time_step), len(MA.primitives))
        self._scene_arrays = None
        self._ensure_library()
        self.index_of = {id(p): i for i, p in enumerate(MA.primitives)}
        self._set_len = np.diff(lay['set_off'])
        self._set_len_list = [int(v) for v in self._set_len]
        self._qbuf = {}
        self.nscenes = 0
        self.mode = MODE_PLANNER
        self.stats = dict(queries=0, candidates=0, refined_fp64=0, near_tangent=0, exact_ties=0, pairs_nominal=0)

    def _ensure_library(self):
        be, lay = self.backend, self.layout
        if getattr(be, '_staged_library', None) != self._lib_key:
            be.upload_library(lay['verts'], lay['nv'], lay['tidx'], lay['set_off'], lay['set_idx'])
            be._staged_library = self._lib_key
            be._staged_library_ref = self.MA          # keeps id(MA) from being reused

    def set_scenes(self, scene_set, mode=MODE_PLANNER):
        self.scene_arrays = self._scene_arrays = scene_set.stage(self.backend)
        self.backend._scene_owner = self
        self.nscenes = len(scene_set.origins)
        self.mode = mode

    def _ensure_staged(self):
        """An engine holds ONE library and ONE scene set, and several checkers (another planner call, the single-call
        collision_check, a validator) may share the process-wide engine: before a query, whoever staged last must be
        this checker -- otherwise its own library / scenes go back on the device first."""
        be = self.backend
        if getattr(be, '_staged_library', None) != self._lib_key:
            self._ensure_library()
            be._scene_owner = None
        if getattr(be, '_scene_owner', None) is not self:
            if self._scene_arrays is None:
                raise RuntimeError("PrimitiveChecker.set_scenes has not been called")
            a = self._scene_arrays
            be.set_scenes(a['scene_step_off'], a['origins'], a['step_obst_off'], a['obst'], a['obst_nv'],
                          a['step_seg_begin'], a['step_seg_end'], a['segs'])
            be._scene_owner = self

    def _account(self, st, ncand):
        s = self.stats
        s['queries'] += 1
        s['candidates'] += ncand
        for k in ('refined_fp64', 'near_tangent', 'exact_ties', 'pairs_nominal'):
            s[k] += st.get(k, 0)

    def bucket_collides(self, x, y, phi, t0, step_limit, bucket, scenes=(0,)):
        """collision bits (True = collides) of every primitive of velocity bucket `bucket` appended at pose
        (x, y, phi) and time index t0; with several scenes a candidate collides if it does in any of them"""
        if getattr(self.backend, '_scene_owner', None) is not self or getattr(self.backend, '_staged_library', None) != self._lib_key:
            self._ensure_staged()
        cnt = self._set_len_list[bucket]
        ns = len(scenes)
        ent = self._qbuf.get(ns)
        if ent is None:
            q = make_queries(ns)
            ent = self._qbuf[ns] = (q, q.ctypes.data_as(C.c_void_p))     # data_as costs microseconds: once per buffer
        q, qptr = ent
        # one record assignment per scene (field-wise assignment costs 7 us); np.cos / np.sin like the reference
        # (src/lowLevelPlannerManeuverAutomaton.py:121) so that the pose is the one the reference transforms with
        c, sn = np.cos(phi), np.sin(phi)
        for k in range(ns):
            q[k] = (x, y, c, sn, scenes[k], t0, step_limit, bucket, 0, cnt, 0, 0)
        if hasattr(self.backend, 'query_fast'):
            nw = (cnt + 31) // 32
            mask, st = self.backend.query_fast(q, self.mode | MODE_NO_TIMING, want_stats=True, nw=ns * nw, qptr=qptr)
            s = self.stats
            s['queries'] += 1
            s['candidates'] += cnt * ns
            s['refined_fp64'] += st.refined_fp64
            s['near_tangent'] += st.near_tangent
            s['exact_ties'] += st.exact_ties
            s['pairs_nominal'] += st.pairs_nominal
            if ns == 1:
                return np.unpackbits(mask.view(np.uint8), count=cnt, bitorder='little').view(np.bool_)
            bits = np.unpackbits(mask.view(np.uint8).reshape(ns, nw * 4), axis=1, bitorder='little')[:, :cnt]
            return bits.any(axis=0)
        bits, st = self.backend.query(q, None, self.mode)
        self._account(st, cnt * ns)
        return bits.reshape(ns, cnt).any(axis=0)

    # -- primitive-selection step on the device --------------------------------------------------------
    def enable_expand(self, ref_traj, goals, b):
        """stage what mpc_expand needs: primitive trajectories (once per engine and automaton), the reference
        trajectory and the goal sets of this planning problem.  `goals` is the reference's param['goal'] list."""
        be = self.backend
        key = (id(self.MA), len(self.MA.primitives))
        if getattr(be, '_staged_trajectories', None) != key:
            states, nstates = self.MA.trajectory_layout()
            be.upload_trajectories(states, nstates)
            be._staged_trajectories = key
        rec, segs = goal_arrays(goals)
        be.set_guidance(np.ascontiguousarray(ref_traj[0:2]), rec, segs, b)
        self._pbuf = {}

    def bucket_expand(self, x, y, phi, cost, t0, step_limit, bucket, scenes=(0,)):
        """bucket_collides plus the cost expand_node gives every child and the step goal_check would return for it
        (-1: goal not reached): (bits, cost, goal_step)"""
        self._ensure_staged()
        cnt = int(self._set_len[bucket])
        ns = len(scenes)
        ent = self._qbuf.get(ns)
        if ent is None:
            q = make_queries(ns)
            ent = self._qbuf[ns] = (q, q.ctypes.data_as(C.c_void_p))
        q = ent[0]
        c, sn = np.cos(phi), np.sin(phi)
        for k in range(ns):
            q[k] = (x, y, c, sn, scenes[k], t0, step_limit, bucket, 0, cnt, 0, 0)
        parent = self._pbuf.get(ns)
        if parent is None:
            parent = self._pbuf[ns] = np.zeros((ns, 2))
        parent[:, 0], parent[:, 1] = cost, phi
        bits, costs, goal, st = self.backend.expand(q, parent, None, self.mode)
        self._account(st, cnt * ns)
        return bits.reshape(ns, cnt).any(axis=0), costs[:cnt], goal[:cnt]

    def list_collides(self, x, y, phi, t0, step_limit, prim_indices, scenes=(0,)):
        """same for an explicit list of primitive indices"""
        self._ensure_staged()
        idx = np.asarray(prim_indices, dtype=np.int32)
        q = make_queries(len(scenes))
        q['x'], q['y'], q['c'], q['s'] = x, y, np.cos(phi), np.sin(phi)
        q['t0'], q['step_limit'] = t0, step_limit
        q['cand_set'], q['cand_off'], q['cand_cnt'] = -1, 0, len(idx)
        q['scene'] = list(scenes)
        bits, st = self.backend.query(q, idx, self.mode)
        self._account(st, len(idx) * len(scenes))
        return bits.reshape(len(scenes), len(idx)).any(axis=0)


def goal_arrays(goals):
    """param['goal'] (list of dicts with time_start, time_end, space) -> (GOAL_DTYPE records, boundary segments [n, 4])"""
    from ._cabi import GOAL_DTYPE
    rec = np.zeros(len(goals), GOAL_DTYPE)
    parts, n = [], 0
    for k, g in enumerate(goals):
        if g['space'] is None:
            rec[k] = (int(g['time_start']), int(g['time_end']), -1, -1)
            continue
        s = ring_segments(rings_of(g['space']))
        rec[k] = (int(g['time_start']), int(g['time_end']), n, n + len(s))
        parts.append(s)
        n += len(s)
    return rec, (np.concatenate(parts) if parts else np.zeros((0, 4)))


def car_library(length, width):
    """one-primitive library holding the ego rectangle about its centre (commonroad Rectangle vertex order),
    used by the trajectory validator"""
    l2, w2 = 0.5 * length, 0.5 * width
    verts = np.array([[[[-l2, -w2], [-l2, w2], [l2, w2], [l2, -w2]]]], dtype=np.float64)
    return dict(verts=verts, nv=np.array([[4]], np.int32), tidx=np.array([[0]], np.int32))


def validate_trajectory(engine, traj, length, width, scene_index=0):
    """per-step collision bits of a centre-frame trajectory (4 x N) against a staged scene: the ego rectangle at
    step i is tested against the region and the obstacles of step i (reference auxiliary/collisionChecker.py:13-33).
    The library of `engine` must be car_library(length, width)."""
    n = traj.shape[1]
    q = make_queries(n)
    phi = np.mod(traj[3, :], 2 * np.pi)            # reference :17
    q['x'], q['y'], q['c'], q['s'] = traj[0, :], traj[1, :], np.cos(phi), np.sin(phi)
    q['scene'], q['t0'], q['cand_set'], q['cand_off'], q['cand_cnt'] = scene_index, np.arange(n), -1, 0, 1
    bits, st = engine.query(q, np.zeros(1, np.int32), MODE_VALIDATOR)
    return bits.astype(bool), st
