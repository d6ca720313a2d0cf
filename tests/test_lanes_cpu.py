"""Host-side scheduling of CollisionLanes.run_cycles (mpc_submit / mpc_wait driven by one thread) with stub engines:
cycle i runs on lane i % lanes, at most `lanes` cycles are in flight, a lane's previous cycle is collected before its
buffers are refilled, results come back in cycle order, and an error in one cycle still drains the others."""
import numpy as np
import pytest

from motionplanner_b200.engine import make_queries
from motionplanner_b200.lanes import SCENE_KEYS, CollisionLanes


class StubEngine:
    log = []

    def __init__(self, lane):
        self.lane, self.inflight, self.fail_on = lane, None, None

    def submit(self, scene_args, queries, mode, cand=None, nw=None):
        assert self.inflight is None, "submit on a lane whose previous cycle has not been collected"
        tag = int(queries["t0"][0])
        StubEngine.log.append(("submit", self.lane, tag))
        self.inflight = (tag, scene_args)

    def wait(self, want_stats=True):
        tag, scene_args = self.inflight
        self.inflight = None
        StubEngine.log.append(("wait", self.lane, tag))
        if tag == self.fail_on:
            raise RuntimeError("cycle %d failed" % tag)
        return np.array([tag], np.uint32), ({"tag": tag} if want_stats else None)

    def close(self):
        pass


def make_pool(n):
    pool = CollisionLanes.__new__(CollisionLanes)
    pool.engines = [StubEngine(k) for k in range(n)]
    StubEngine.log = []
    return pool


def scene():
    return dict(scene_step_off=np.array([0, 1], np.int32), origins=np.zeros((1, 2)), step_obst_off=np.array([0, 0], np.int32),
                obst=np.zeros((0, 4, 2)), obst_nv=np.zeros(0, np.int32), step_seg_begin=np.array([-1], np.int32),
                step_seg_end=np.array([-1], np.int32), segs=np.zeros((0, 4)))


def cycles(n, with_scene=True):
    out = []
    for i in range(n):
        q = make_queries(1)
        q["t0"] = i
        out.append((scene() if with_scene else None, q))
    return out


@pytest.mark.parametrize("lanes,n", [(1, 5), (3, 10), (4, 4), (4, 2)])
def test_round_robin_order_and_inflight_bound(lanes, n):
    pool = make_pool(lanes)
    refills = []
    res = pool.run_cycles(cycles(n), refill=lambda i, sc, q: refills.append((i, len([e for e in pool.engines if e.inflight]))))
    assert [int(r[0][0]) for r in res] == list(range(n)) and all(r[1]["tag"] == i for i, r in enumerate(res))
    subs = [e for e in StubEngine.log if e[0] == "submit"]
    assert [(e[1], e[2]) for e in subs] == [(i % lanes, i) for i in range(n)]
    # a lane's previous cycle is collected before its buffers are refilled: never more than lanes - 1 others in flight
    assert refills == [(i, min(i, lanes - 1)) for i in range(n)]
    assert all(e.inflight is None for e in pool.engines)


def test_marshal_once_reuses_the_argument_tuple_and_keep_scene():
    pool = make_pool(2)
    sc = scene()
    q = make_queries(1)
    seen = []
    for e in pool.engines:
        orig = e.submit
        e.submit = lambda a, qq, m, c=None, nw=None, _o=orig: (seen.append(a), _o(a, qq, m, c, nw))[1]
    pool.run_cycles([(sc, q)] * 4 + [(None, q)], marshal_once=True, want_stats=False)
    assert seen[0] is seen[1] is seen[2] is seen[3] and seen[4] is None
    assert len(seen[0]) == 12 and seen[0][0] == 1            # the tuple of CollisionEngine.scene_args
    pool.run_cycles([(sc, q)] * 2)
    assert seen[5] is not seen[6]                             # marshalled per cycle by default


def test_error_in_one_cycle_drains_the_others():
    pool = make_pool(3)
    pool.engines[1].fail_on = 4
    with pytest.raises(RuntimeError, match="cycle 4 failed"):
        pool.run_cycles(cycles(9))
    assert all(e.inflight is None for e in pool.engines)
    res = pool.run_cycles(cycles(3))                          # still usable afterwards
    assert [int(r[0][0]) for r in res] == [0, 1, 2]


def test_scene_keys_match_set_scenes_signature():
    import inspect
    from motionplanner_b200.engine import CollisionEngine
    assert tuple(inspect.signature(CollisionEngine.scene_args).parameters) == SCENE_KEYS
