"""test helpers: an oracle-backed stand-in for CollisionEngine (same upload_library / set_scenes / query surface), the
golden scenario fixtures, and a bounded random expansion driver that produces planner-shaped queries."""
import os

import numpy as np

from motionplanner_b200.checker import PrimitiveChecker, SceneSet
from motionplanner_b200.engine import MODE_VALIDATOR
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scenarios.npz')


class OracleBackend:
    """CPU oracle behind the engine interface -- lets the host-side planner code run on CPU in tests and gives the
    expected bits for the GPU parity tests.  TEST ONLY."""

    def __init__(self):
        self.lib = self.sc = None
        self.log = []          # (queries, cand, mode, bits) of every call
        self._staged_library = self._scene_owner = None

    def upload_library(self, verts, nv, tidx, set_off=None, set_idx=None):
        self.lib = dict(verts=np.array(verts, np.float64), nv=np.array(nv, np.int32), tidx=np.array(tidx, np.int32),
                        set_off=None if set_off is None else np.array(set_off, np.int32),
                        set_idx=None if set_idx is None else np.array(set_idx, np.int32))
        self.pb = None
        self._staged_library = None

    def set_scenes(self, scene_step_off, origins, step_obst_off, obst, obst_nv, step_seg_begin, step_seg_end, segs):
        obst = np.asarray(obst, np.float64)
        if obst.ndim != 3:
            obst = obst.reshape(-1, 4, 2)
        self.sc = dict(scene_step_off=scene_step_off, step_obst_off=step_obst_off, obst=obst, obst_nv=obst_nv,
                       step_seg_begin=step_seg_begin, step_seg_end=step_seg_end, segs=segs)
        self.pb = None
        self._scene_owner = None

    def query(self, queries, cand_idx=None, mode=0):
        if self.pb is None:
            self.pb = orc.Problem(self.lib, self.sc)
        omode = (orc.MODE_STRICT if mode & MODE_VALIDATOR else 0) | orc.MODE_NO_EARLY_EXIT
        bits, st = self.pb.check(queries, cand_idx, omode)
        st = dict(st, refined_fp64=0, exact_ties=st['exact_calls'])
        self.log.append((np.array(queries), None if cand_idx is None else np.array(cand_idx), mode, bits.copy()))
        return bits, st


    # primitive-selection step (mpc_expand stand-in)
    def upload_trajectories(self, states, nstates):
        self.traj = (np.array(states, np.float64), np.array(nstates, np.int32))

    def set_guidance(self, ref_xy, goals, goal_segs, b):
        self.guidance = (np.array(ref_xy, np.float64), np.array(goals), np.array(goal_segs, np.float64).reshape(-1, 4), float(b))

    def expand(self, queries, parent, cand_idx=None, mode=0):
        bits, st = self.query(queries, cand_idx, mode)
        ref, goals, segs, b = self.guidance
        cost, goal = orc.expand(self.traj[0], self.traj[1], ref, goals, segs, b, self.lib['set_off'], self.lib['set_idx'],
                                queries, cand_idx, parent)
        return bits, cost, goal, st


def load_golden():
    from motionplanner_b200.workloads import load_golden_scenarios
    return load_golden_scenarios(GOLDEN)


def scenario_scene(g, region='road_safe'):
    s = SceneSet()
    segs = g[region] if len(g[region]) else g['road']
    s.add_scene_explicit(g['nsteps'], g['step_obst_off'], g['obst'], g['obst_nv'], segs, origin=g['x0'][0:2])
    return s


def random_walk(checker, MA, g, rng, expansions=12, b=1.15):
    """bounded random descent through the automaton from the scenario's initial state: every expansion asks the
    checker for the collision bits of one velocity bucket at the current pose (exactly the query the batched planner
    issues), then steps into a random free child (or a random child if all collide).
    Returns (list of bit arrays, visited rear-axle states [4 x n])."""
    x0 = g['x0']
    state = np.array([x0[0] - np.cos(x0[3]) * b, x0[1] - np.sin(x0[3]) * b, x0[2], x0[3]])
    t0 = 0
    bits_all, traj = [], [state.copy()]
    steps = g['t_end']
    for _ in range(expansions):
        if t0 > steps:
            break
        bucket = int(MA._bucket(state[2]))
        cand = MA.vel_dic[bucket]
        if not cand:
            break
        bits = checker.bucket_collides(state[0], state[1], state[3], t0, steps, bucket)
        bits_all.append(bits.copy())
        free = np.nonzero(~bits)[0]
        pick = cand[int(rng.choice(free))] if len(free) else cand[int(rng.integers(len(cand)))]
        prim = MA.primitives[pick]
        c, s = np.cos(state[3]), np.sin(state[3])
        for k in range(1, prim.x.shape[1]):
            traj.append(np.array([c * prim.x[0, k] - s * prim.x[1, k] + state[0],
                                  s * prim.x[0, k] + c * prim.x[1, k] + state[1], prim.x[2, k], prim.x[3, k] + state[3]]))
        state = traj[-1].copy()
        t0 += prim.x.shape[1] - 1
    return bits_all, np.array(traj).T
