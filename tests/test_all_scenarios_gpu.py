"""BASELINE.json configs[1]: the stand-alone maneuver-automaton planner (reference
src/maneuverAutomatonPlannerStandalone.py:19-96) run through the reference's batch harness
(evaluation/runAllScenarios.py:37-108, mirrored in motionplanner_b200/evaluation/runAllScenarios.py) on ALL 100 bundled
scenarios, every collision decision of the search diffed against the oracle.

Per scenario the search runs on the GPU engine with every batched query logged (the query records and the packed
result words).  The oracle then decides the very same queries (same library arrays, same scene arrays).  The host
side of the search is deterministic given the bits, so identical bits in every query mean an identical search and
an identical plan; on a subset of the scenarios the search is additionally re-run on the oracle alone and the two
plans and query sequences are compared directly.  Every plan found must pass the drop-in
``collisionChecker(scenario, x, param)`` (GPU, validator mode) and the oracle's validator.

Environment:  MPC_SCEN_MAX_NODES  cap on popped nodes per scenario (default 300; the committed campaign used more, see
                                  profiles/r02_all_scenarios.json); the reference has no cap and searches until the
                                  queue is empty
              MPC_SCEN_OUT        directory for computation_time_AutomatonStandalone.csv (reference column format) and
                                  all_scenarios.json
"""
import json
import lzma
import os
import pickle
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import OracleBackend
from motionplanner_b200.auxiliary import collisionChecker as CC
from motionplanner_b200.engine import MODE_PLANNER, CollisionEngine
from motionplanner_b200.evaluation.runAllScenarios import PLANNER, solve_loaded
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from oracle import oracle as orc

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'all_scenarios.pkl.xz')
MAX_NODES = int(os.environ.get('MPC_SCEN_MAX_NODES', '300'))
OUT_DIR = os.environ.get('MPC_SCEN_OUT')


class LoggedEngine(CollisionEngine):
    """records the scene arrays of every staging call and (query records, packed result words) of every query"""

    def reset_log(self):
        self.scene, self.calls, self.near_tangent, self.refined = None, [], 0, 0

    def set_scenes(self, *a):
        self.scene = [np.array(x) for x in a]
        return super().set_scenes(*a)

    def query_fast(self, queries, mode=0, want_stats=False, **kw):
        mask, st = super().query_fast(queries, mode, True, **kw)
        self.calls.append((np.array(queries), mask.copy(), int(mode)))
        self.near_tangent += st.near_tangent
        self.refined += st.refined_fp64
        return mask, st


def oracle_bits(lib, scene, calls):
    """the oracle's decisions for the logged queries, one flat uint8 array per call"""
    keys = ('scene_step_off', 'origins', 'step_obst_off', 'obst', 'obst_nv', 'step_seg_begin', 'step_seg_end', 'segs')
    sc = dict(zip(keys, scene))
    if np.asarray(sc['obst']).ndim != 3:
        sc['obst'] = np.asarray(sc['obst'], np.float64).reshape(-1, 4, 2)
    pb = orc.Problem(lib, sc)
    q = np.concatenate([c[0] for c in calls])
    bits, st = pb.check(q, None, orc.MODE_NO_EARLY_EXIT, nthreads=os.cpu_count() or 1)
    return np.split(bits, np.cumsum([int(c[0]['cand_cnt'].sum()) for c in calls])[:-1]), st


def test_standalone_planner_on_all_bundled_scenarios():
    with open(FIXTURE, 'rb') as f:
        scenarios = pickle.loads(lzma.decompress(f.read()))
    assert len(scenarios) == 100
    MA = load_maneuver_automaton()
    lay = MA.device_layout(0.1)
    lib = {k: lay[k] for k in ('verts', 'nv', 'tidx', 'set_off', 'set_idx')}
    eng = LoggedEngine(0)
    rows, per = [], []
    tot = dict(scenarios=0, queries=0, decisions=0, collided=0, disagreements=0, near_tangent_lt_1e_9=0, refined_fp64=0,
               plans=0, plans_pass_collisionChecker=0, plans_pass_oracle_validator=0, failed=0, errors=0,
               oracle_search_reruns=0, oracle_search_identical=0)
    t_all = time.time()
    for k, (name, (scenario, pps)) in enumerate(sorted(scenarios.items())):
        eng.reset_log()
        t0 = time.time()
        row, x, u, param = solve_loaded(name + '.xml', scenario, pps, MA, MAX_NODES, backend=eng)
        t_gpu = time.time() - t0
        rows.append(row)
        info = dict(name=name, row=row[1:3], queries=len(eng.calls), search_s=round(t_gpu, 3))
        tot['scenarios'] += 1
        if isinstance(row[1], str) and row[1] != 'failed':
            tot['errors'] += 1
        elif row[1] == 'failed':
            tot['failed'] += 1
        # ---- every decision of the search against the oracle
        if eng.calls:
            assert abs(scenario.dt - 0.1) < 1e-12
            want, _ = oracle_bits(lib, eng.scene, eng.calls)
            for (q, mask, mode), w in zip(eng.calls, want):
                assert mode & 1 == MODE_PLANNER
                got = np.unpackbits(mask.view(np.uint8), bitorder='little')
                nw = (q['cand_cnt'].astype(np.int64) + 31) // 32
                got = np.concatenate([got[32 * o:32 * o + c] for o, c in zip(np.cumsum(nw) - nw, q['cand_cnt'])])
                tot['decisions'] += len(w)
                tot['collided'] += int(w.sum())
                tot['disagreements'] += int((got != w).sum())
            tot['queries'] += len(eng.calls)
            tot['near_tangent_lt_1e_9'] += int(eng.near_tangent)
            tot['refined_fp64'] += int(eng.refined)
            info.update(decisions=int(sum(len(w) for w in want)), near_tangent=int(eng.near_tangent))
        # ---- plans: the drop-in validator on the GPU and the oracle's validator must both accept them
        if x is not None:
            tot['plans'] += 1
            tot['plans_pass_collisionChecker'] += int(row[2] == 'no collision')
            flags, _ = CC.collision_flags(scenario, x, param, engine=OracleBackend())
            tot['plans_pass_oracle_validator'] += int(not flags.any())
        # ---- on every fifth scenario: the whole search once more on the oracle alone
        if k % 5 == 0 and not (isinstance(row[1], str) and row[1] != 'failed'):
            cpu = OracleBackend()
            row2, x2, u2, _ = solve_loaded(name + '.xml', scenario, pps, MA, MAX_NODES, backend=cpu)
            same = (x is None) == (x2 is None) and len(cpu.log) == len(eng.calls)
            if same and x is not None:
                same = np.array_equal(x, x2) and np.array_equal(u, u2)
            tot['oracle_search_reruns'] += 1
            tot['oracle_search_identical'] += int(same)
            info['oracle_search_identical'] = bool(same)
        per.append(info)
    eng.close()
    tot['wall_s'] = round(time.time() - t_all, 1)
    tot['max_nodes'] = MAX_NODES
    print(json.dumps(tot))
    if OUT_DIR:
        os.makedirs(OUT_DIR, exist_ok=True)
        np.savetxt(os.path.join(OUT_DIR, 'computation_time_' + PLANNER + '.csv'), np.asarray(rows, dtype=object),
                   delimiter=",", fmt="%s")
        with open(os.path.join(OUT_DIR, 'all_scenarios.json'), 'w') as f:
            json.dump(dict(total=tot, scenarios=per), f, indent=1)
    assert tot['errors'] == 0, [r for r in rows if isinstance(r[1], str) and r[1] != 'failed']
    assert tot['decisions'] > 100000 and 0 < tot['collided'] < tot['decisions']
    assert tot['disagreements'] == 0
    assert tot['oracle_search_identical'] == tot['oracle_search_reruns'] > 0
    assert tot['plans_pass_collisionChecker'] == tot['plans_pass_oracle_validator'] == tot['plans']
