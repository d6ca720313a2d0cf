"""Maneuver-automaton data model -- same public attributes as reference maneuverAutomaton/ManeuverAutomaton.py:3-44,
plus the flat array view (`device_layout`) that is staged once in HBM (include/mpc.h: mpc_upload_library)."""
import weakref

import numpy as np

from ..geometry import polygon_vertices

# id(primitive) -> (weak reference to its automaton, index): lets the module-level collision_check(node, primitive, ...)
# of the planners find the staged library from the primitive alone, as the reference's signature demands
# (src/lowLevelPlannerManeuverAutomaton.py:95), without anything being planted in `param`
_PRIMITIVE_OWNER = {}


def automaton_of(primitive):
    """(automaton, index of `primitive` in automaton.primitives) for a primitive of a live ManeuverAutomaton"""
    hit = _PRIMITIVE_OWNER.get(id(primitive))
    MA = hit[0]() if hit is not None else None
    if MA is None or hit[1] >= len(MA.primitives) or MA.primitives[hit[1]] is not primitive:
        raise KeyError("motion primitive does not belong to a ManeuverAutomaton of this package (build the automaton "
                       "with motionplanner_b200.maneuverAutomaton.ManeuverAutomaton.ManeuverAutomaton)")
    return MA, hit[1]


def _fast_vertices(geom):
    """open vertex array of a polygon-like object (closing vertex and repeated vertices dropped)"""
    c = geom.exterior if hasattr(geom, 'exterior') else geom
    c = c if hasattr(c, '__array__') else c.coords
    v = np.asarray(c, dtype=np.float64).reshape(-1, 2)
    if len(v) > 1 and v[0, 0] == v[-1, 0] and v[0, 1] == v[-1, 1]:
        v = v[:-1]
    if len(v) > 1:
        keep = np.ones(len(v), bool)
        keep[1:] = (v[1:, 0] != v[:-1, 0]) | (v[1:, 1] != v[:-1, 1])
        if not keep.all():
            v = v[keep]
    return v


class MotionPrimitive:
    """one motion primitive: reference trajectory, inputs, duration, successors and occupancy sets"""

    def __init__(self, x, u, tFinal, successors=None, occupancy=None):
        self.x = x                      # states of the reference trajectory, 4 x (N+1): x, y, v, phi
        self.u = u                      # control inputs
        self.tFinal = tFinal            # duration
        self.successors = successors    # indices of primitives that may follow
        self.occ = occupancy            # [{'space': polygon, 'time': t} | {'space', 'starttime', 'endtime'}]


class ManeuverAutomaton:
    """velocity-bucketed library of motion primitives"""

    def __init__(self, primitives):
        self.primitives = primitives
        v_start = [p.x[2, 0] for p in primitives]
        uniq = np.sort(np.unique(v_start))
        self.v_diff = np.sort(np.diff(uniq))[0]
        nbuckets = 2 * np.ceil(max(v_start) / self.v_diff).astype(int)
        self.vel_dic = [[] for _ in range(nbuckets)]
        for i, p in enumerate(primitives):
            self.vel_dic[self._bucket(p.x[2, 0])].append(i)
        # a primitive can be followed by every primitive that starts in the bucket of its final velocity
        for p in self.primitives:
            p.successors = self.velocity2primitives(p.x[2, -1])
        self._layouts = {}
        self._register()

    def _register(self):
        def forget(dead, ids=[id(p) for p in self.primitives]):
            for i in ids:
                if _PRIMITIVE_OWNER.get(i, (None,))[0] is dead:
                    del _PRIMITIVE_OWNER[i]

        ref = weakref.ref(self, forget)
        for i, p in enumerate(self.primitives):
            _PRIMITIVE_OWNER[id(p)] = (ref, i)

    def _bucket(self, v):
        return np.round(v / self.v_diff).astype(int)

    def velocity2primitives(self, v):
        """all motion primitives starting at (the bucket of) velocity v"""
        return self.vel_dic[self._bucket(v)]

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop('_layouts', None)
        d.pop('_stacks', None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._layouts = {}
        self._register()

    def bucket_stacks(self, bucket):
        """reference trajectories of a velocity bucket stacked by length: [(positions in the bucket, [m, 4, L])];
        lets a planner expand all children of a node with one batched product instead of one per child"""
        cache = self.__dict__.setdefault('_stacks', {})
        if bucket not in cache:
            by_len = {}
            for k, i in enumerate(self.vel_dic[bucket]):
                by_len.setdefault(self.primitives[i].x.shape[1], []).append(k)
            cache[bucket] = [(np.array(ks), np.stack([self.primitives[self.vel_dic[bucket][k]].x for k in ks]))
                             for _, ks in sorted(by_len.items())]
        return cache[bucket]

    # ------------------------------------------------------------------------------------------------------------
    def is_timepoint(self):
        """time-point occupancies ('time') or time-interval occupancies ('starttime'/'endtime'); the reference looks
        at primitive 1 (src/lowLevelPlannerManeuverAutomaton.py:25)"""
        probe = self.primitives[1] if len(self.primitives) > 1 else self.primitives[0]
        return 'starttime' not in probe.occ[0]

    def trajectory_layout(self):
        """(states [P, L, 4], nstates [P]) for mpc_upload_trajectories: primitive.x transposed, padded to the longest"""
        if getattr(self, '_traj_layout', None) is None:
            L = max(p.x.shape[1] for p in self.primitives)
            states = np.zeros((len(self.primitives), L, 4))
            nstates = np.zeros(len(self.primitives), np.int32)
            for i, p in enumerate(self.primitives):
                states[i, :p.x.shape[1]] = p.x.T
                nstates[i] = p.x.shape[1]
            self._traj_layout = (states, nstates)
        return self._traj_layout

    def device_layout(self, time_step):
        """flat arrays for mpc_upload_library.

        Time indices are computed exactly like the reference does at check time,
        ``int(o['time'] / param['time_step'])`` (src/lowLevelPlannerManeuverAutomaton.py:116-119), in float64 on the
        host.  Slots are re-bucketed so that slot j means the same time index for every primitive; two occupancies of
        one primitive that truncate to the same index (AROC libraries) get separate slots.
        Candidate sets = the velocity buckets; a primitive's successors are exactly one bucket, so a search expansion
        references its candidates by bucket id."""
        key = float(time_step)
        if key in getattr(self, '_layouts', {}):
            return self._layouts[key]
        timepoint = self.is_timepoint()
        per_prim, slot_keys, vmax = [], set(), 3
        for p in self.primitives:
            seen, items = {}, []
            arr = getattr(p, '_occ_verts', None)
            for k, o in enumerate(p.occ or []):
                t = int((o['time'] if timepoint else o['starttime']) / time_step)
                rank = seen.get(t, 0)
                seen[t] = rank + 1
                v = arr[k] if arr is not None else _fast_vertices(o['space'])
                vmax = max(vmax, len(v))
                items.append(((t, rank), v))
                slot_keys.add((t, rank))
            per_prim.append(items)
        slots = {k: i for i, k in enumerate(sorted(slot_keys))}
        P, J = len(self.primitives), max(len(slots), 1)
        verts = np.zeros((P, J, vmax, 2))
        nv = np.zeros((P, J), np.int32)
        tidx = np.zeros((P, J), np.int32)
        for (t, _), j in slots.items():
            tidx[:, j] = t
        for i, items in enumerate(per_prim):
            for k, v in items:
                j = slots[k]
                verts[i, j, :len(v)] = v
                nv[i, j] = len(v)
        set_off = np.zeros(len(self.vel_dic) + 1, np.int32)
        set_off[1:] = np.cumsum([len(b) for b in self.vel_dic])
        set_idx = np.array([i for b in self.vel_dic for i in b], dtype=np.int32)
        # end pose of every primitive (child expansion) and bucket of its successors
        end = np.array([[p.x[0, -1], p.x[1, -1], p.x[3, -1]] for p in self.primitives])
        succ_bucket = np.array([self._bucket(p.x[2, -1]) for p in self.primitives], dtype=np.int32)
        lay = dict(verts=verts, nv=nv, tidx=tidx, set_off=set_off, set_idx=set_idx, end=end,
                   succ_bucket=succ_bucket, timepoint=timepoint)
        self._layouts[key] = lay
        return lay
