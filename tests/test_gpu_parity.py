"""GPU parity tests proper: the CUDA path (through the C ABI, include/mpc.h) against the float64 exact-sign oracle on
the same seeded inputs.  Bar: bit-exact decisions (the path is boolean); the only tolerated source of disagreement
would be contacts closer than 1e-9 m, and those are counted -- the assertions below demand zero disagreements."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from motionplanner_b200 import workloads as W
from motionplanner_b200.engine import (MODE_COUNT_WORK, MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_PLANNER, MODE_VALIDATOR,
                                       CollisionEngine, MpcError, make_queries)
from oracle import oracle as orc


@pytest.fixture(scope="module")
def eng():
    e = CollisionEngine(0)
    yield e
    e.close()


def stage(eng, lib, sc, origins):
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib.get("set_off"), lib.get("set_idx"))
    eng.set_scenes(sc["scene_step_off"], origins, sc["step_obst_off"], sc["obst"], sc["obst_nv"],
                   sc["step_seg_begin"], sc["step_seg_end"], sc["segs"])


def both(eng, lib, sc, origins, q, cand=None, mode=0):
    stage(eng, lib, sc, origins)
    got, st = eng.query(q, cand, mode)
    omode = (orc.MODE_STRICT if mode & MODE_VALIDATOR else 0) | orc.MODE_NO_EARLY_EXIT
    want, ost = orc.Problem(lib, sc).check(q, cand, omode)
    return got, want, st, ost


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("mode", [MODE_PLANNER, MODE_VALIDATOR])
def test_adversarial_rectangles(eng, seed, mode):
    lib, sc, org, q = W.adversarial_problem(seed)
    got, want, st, ost = both(eng, lib, sc, org, q, None, mode)
    assert np.array_equal(got, want), "disagreements: %d" % (got != want).sum()
    assert st["pairs_nominal"] == ost["pairs_nominal"]
    assert 0 < want.mean() < 1                     # the problem exercises both outcomes
    assert st["refined_fp64"] > 0                  # near-contacts really went through the float64 pass


@pytest.mark.parametrize("seed", range(3))
def test_adversarial_far_from_origin(eng, seed):
    """|coordinates| like the bundled scenarios (up to 1.4 km): the scene origin keeps float32 local"""
    lib, sc, org, q = W.adversarial_problem(100 + seed, offset=(1421.0, -677.0))
    for mode in (MODE_PLANNER, MODE_VALIDATOR):
        got, want, st, _ = both(eng, lib, sc, org, q, None, mode)
        assert np.array_equal(got, want)
    # a badly chosen origin only widens the band, it must not change a decision
    got, want, st2, _ = both(eng, lib, sc, np.zeros((1, 2)), q, None, MODE_PLANNER)
    assert np.array_equal(got, want)
    assert st2["tau"] > st["tau"]


@pytest.mark.parametrize("lib_verts,obst_verts", [(4, 8), (8, 4), (12, 8), (16, 3)])
def test_adversarial_general_polygons(eng, lib_verts, obst_verts):
    for seed in range(2):
        lib, sc, org, q = W.adversarial_problem(200 + seed, lib_verts=lib_verts, obst_verts=obst_verts)
        for mode in (MODE_PLANNER, MODE_VALIDATOR):
            got, want, st, _ = both(eng, lib, sc, org, q, None, mode)
            assert np.array_equal(got, want), (lib_verts, obst_verts, seed, mode)


def test_mode_flags_do_not_change_decisions(eng):
    lib, sc, org, q = W.adversarial_problem(7, P=200, max_obst=70)
    stage(eng, lib, sc, org)
    base, st0 = eng.query(q, None, MODE_PLANNER)
    for m in (MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_NO_CULL | MODE_NO_EARLY_EXIT, MODE_COUNT_WORK,
              MODE_COUNT_WORK | MODE_NO_EARLY_EXIT | MODE_NO_CULL):
        got, st = eng.query(q, None, m)
        assert np.array_equal(got, base), m
    # instrumented launch: without culling or exits every nominal pair runs the full axis test
    assert st["axis_tests"] == st["pairs_nominal"]
    got, st = eng.query(q, None, MODE_COUNT_WORK | MODE_NO_EARLY_EXIT)
    assert st["cull_tests"] + st["box_skipped"] == st["pairs_nominal"] and 0 < st["axis_tests"] < st["pairs_nominal"]


def test_explicit_candidate_lists_and_sets(eng):
    lib, sc, org, q = W.adversarial_problem(11)
    stage(eng, lib, sc, org)
    P = lib["verts"].shape[0]
    rng = np.random.default_rng(0)
    cand = rng.permutation(P).astype(np.int32)[:70]
    q2 = q.copy()
    q2["cand_set"] = -1
    q2["cand_off"] = [0, 10, 33]
    q2["cand_cnt"] = [10, 23, 37]
    got, _ = eng.query(q2, cand, MODE_PLANNER)
    want, _ = orc.Problem(lib, sc).check(q2, cand, orc.MODE_NO_EARLY_EXIT)
    assert np.array_equal(got, want)
    # permutation equivariance: decisions follow the candidates
    q3 = q[:1].copy()
    q3["cand_set"] = -1
    q3["cand_off"] = 0
    q3["cand_cnt"] = P
    a, _ = eng.query(q3, np.arange(P, dtype=np.int32), MODE_PLANNER)
    perm = rng.permutation(P).astype(np.int32)
    b, _ = eng.query(q3, perm, MODE_PLANNER)
    assert np.array_equal(a[perm], b)


def test_time_index_guards(eng):
    """reference guards (src/lowLevelPlannerManeuverAutomaton.py:120,122): ind+time <= param['steps'] and
    ind+time < len(space); skipped occupancies count as free"""
    lib, sc, org, q = W.adversarial_problem(3)
    stage(eng, lib, sc, org)
    nsteps = int(sc["scene_step_off"][1])
    for t0, lim in ((nsteps, 99), (nsteps - 1, 99), (0, 0), (0, -1), (2, 3), (-3, 99)):
        qq = q.copy()
        qq["t0"], qq["step_limit"] = t0, lim
        got, _ = eng.query(qq, None, MODE_PLANNER)
        want, _ = orc.Problem(lib, sc).check(qq, None, orc.MODE_NO_EARLY_EXIT)
        assert np.array_equal(got, want), (t0, lim)
    qq = q.copy()
    qq["t0"] = nsteps
    got, _ = eng.query(qq, None, MODE_PLANNER)
    assert got.sum() == 0                      # everything beyond the horizon: nothing is tested


def test_empty_and_ragged_inputs(eng):
    lib, sc, org, q = W.adversarial_problem(4)
    stage(eng, lib, sc, org)
    # empty batch, empty candidate list
    got, st = eng.query(make_queries(0), None, MODE_PLANNER)
    assert got.size == 0
    qq = q.copy()
    qq["cand_cnt"] = [0, 5, 0]
    got, _ = eng.query(qq, None, MODE_PLANNER)
    want, _ = orc.Problem(lib, sc).check(qq, None, orc.MODE_NO_EARLY_EXIT)
    assert got.size == 5 and np.array_equal(got, want)
    # scene without obstacles and without region: everything free
    nsteps = 4
    sc0 = dict(scene_step_off=np.array([0, nsteps], np.int32), step_obst_off=np.zeros(nsteps + 1, np.int32),
               obst=np.zeros((0, 4, 2)), obst_nv=np.zeros(0, np.int32), step_seg_begin=np.full(nsteps, -1, np.int32),
               step_seg_end=np.full(nsteps, -1, np.int32), segs=np.zeros((0, 4)))
    stage(eng, lib, sc0, org)
    got, st = eng.query(q, None, MODE_PLANNER)
    assert got.sum() == 0 and st["pairs_nominal"] == 0
    # empty region at every step: nothing is contained (shapely: empty.contains(x) is False)
    sc0["step_seg_begin"][:] = 0
    sc0["step_seg_end"][:] = 0
    stage(eng, lib, sc0, org)
    got, _ = eng.query(q, None, MODE_PLANNER)
    want, _ = orc.Problem(lib, sc0).check(q, None, orc.MODE_NO_EARLY_EXIT)
    assert np.array_equal(got, want) and got.sum() > 0


def test_many_obstacles_and_segments_multi_phase(eng):
    """more obstacles / segments per step than one shared-memory tile holds"""
    rng = np.random.default_rng(5)
    lib, sc, org, q = W.adversarial_problem(9, P=64, nsteps=3, max_obst=4)
    nsteps = 3
    O = 700
    obst = W.rect_vertices(rng.uniform(-70, 70, (nsteps, O)), rng.uniform(-70, 70, (nsteps, O)),
                           rng.uniform(-3, 3, (nsteps, O)), 1.0, 0.5).reshape(-1, 4, 2)
    sc["obst"], sc["obst_nv"] = obst, np.full(nsteps * O, 4, np.int32)
    sc["step_obst_off"] = (np.arange(nsteps + 1) * O).astype(np.int32)
    # a many-sided ring (1200 segments) as the region
    n = 1200
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    ring = np.stack([75 * np.cos(ang), 75 * np.sin(ang)], -1)
    sc["segs"] = np.concatenate([ring, np.roll(ring, -1, axis=0)], axis=1)
    sc["step_seg_begin"] = np.zeros(nsteps, np.int32)
    sc["step_seg_end"] = np.full(nsteps, n, np.int32)
    for mode in (MODE_PLANNER, MODE_NO_EARLY_EXIT):
        got, want, st, ost = both(eng, lib, sc, org, q, None, mode)
        assert np.array_equal(got, want)
        assert st["pairs_nominal"] == ost["pairs_nominal"]
    assert 0 < want.mean() < 1


def test_multiple_scenes_in_one_batch(eng):
    parts = [W.adversarial_problem(300 + i, P=64, nsteps=5 + i, offset=(200.0 * i, -90.0 * i)) for i in range(4)]
    lib = parts[0][0]
    sso, soo, obst, onv, sb, se, segs, orgs, qs = [0], [0], [], [], [], [], [], [], []
    for i, (_, sc, org, q) in enumerate(parts):
        nseg0 = sum(len(s) for s in segs)
        sso.append(sso[-1] + int(sc["scene_step_off"][1]))
        soo.extend((sc["step_obst_off"][1:] + soo[-1]).tolist())
        obst.append(sc["obst"]); onv.append(sc["obst_nv"])
        sb.extend([b + nseg0 if b >= 0 else -1 for b in sc["step_seg_begin"]])
        se.extend([e + nseg0 if e >= 0 else -1 for e in sc["step_seg_end"]])
        segs.append(sc["segs"]); orgs.append(org[0])
        q = q.copy(); q["scene"] = i
        qs.append(q)
    sc = dict(scene_step_off=np.array(sso, np.int32), step_obst_off=np.array(soo, np.int32),
              obst=np.concatenate(obst), obst_nv=np.concatenate(onv), step_seg_begin=np.array(sb, np.int32),
              step_seg_end=np.array(se, np.int32), segs=np.concatenate(segs))
    q = np.concatenate(qs)
    got, want, st, ost = both(eng, lib, sc, np.array(orgs), q, None, MODE_PLANNER)
    assert np.array_equal(got, want)
    # per-scene results equal the single-scene results
    o = 0
    for i, (lib_i, sc_i, org_i, q_i) in enumerate(parts):
        w, _ = orc.Problem(lib, sc_i).check(q_i, None, orc.MODE_NO_EARLY_EXIT)
        n = int(q_i["cand_cnt"].sum())
        assert np.array_equal(got[o:o + n], w)
        o += n


def test_mixed_regime_full_size(eng):
    """the non-degenerate regime of bench.py (`regimes.mixed`: same 4096 x 256 x 50 shape, 35 % of the candidates are
    free and need every time step) at full size against the oracle, in every launch mode bench.py times it in, and
    through the asynchronous cycle (mpc_submit / mpc_wait)"""
    lib, sc, org, q = W.mixed_traffic_problem(nqueries=6)
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    want, ost = pb.check(q, None, 0, inner_parallel=True)
    assert 0.25 < 1.0 - want.mean() < 0.5                        # the regime: 30-50 % free candidates
    got, st = eng.query(q, None, MODE_PLANNER)
    assert st["pairs_nominal"] == 6 * 4096 * 50 * (256 + 4) == ost["pairs_nominal"]
    assert np.array_equal(got, want)
    got_ne, _ = eng.query(q, None, MODE_PLANNER | MODE_NO_EARLY_EXIT)          # `cull_only`
    assert np.array_equal(got_ne, want)
    got_full, _ = eng.query(q[:2], None, MODE_NO_CULL | MODE_NO_EARLY_EXIT)    # `full_evaluation` (k_check_full)
    assert np.array_equal(got_full, want[:2 * 4096])
    wantv, _ = pb.check(q, None, orc.MODE_STRICT, inner_parallel=True)
    gotv, _ = eng.query(q, None, MODE_VALIDATOR)
    assert np.array_equal(gotv, wantv)
    keys = ("scene_step_off", "origins", "step_obst_off", "obst", "obst_nv", "step_seg_begin", "step_seg_end", "segs")
    eng.submit(CollisionEngine.scene_args(*[dict(sc, origins=org)[k] for k in keys]), q, MODE_PLANNER)
    mask, _ = eng.wait()
    from motionplanner_b200.engine import unpack_mask
    assert np.array_equal(unpack_mask(mask, q), want)


@pytest.mark.parametrize("obst_verts", [4, 8])
def test_aroc_shaped_library_against_oracle(eng, obst_verts):
    """the many-vertex instantiations bench.py measures (`aroc_shaped`: 12-vertex convex occupancies over time intervals,
    two scenes per query as for interval automata; k_check<16,4> / <16,8>) against the oracle, default mode and full
    evaluation"""
    lib, sc, org, q = W.aroc_shaped_problem(P=1024, T=10, nqueries=4, obst_verts=obst_verts)
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    want, ost = pb.check(q, None, orc.MODE_NO_EARLY_EXIT)
    got, st = eng.query(q, None, MODE_PLANNER)
    assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    assert 0.0 < want.mean() < 1.0
    got_full, _ = eng.query(q[:2], None, MODE_NO_CULL | MODE_NO_EARLY_EXIT)
    assert np.array_equal(got_full, want[:2 * 1024])
    wantv, _ = pb.check(q, None, orc.MODE_STRICT | orc.MODE_NO_EARLY_EXIT)
    gotv, _ = eng.query(q, None, MODE_VALIDATOR)
    assert np.array_equal(gotv, wantv)


def test_c4_dense_traffic_full_size(eng):
    """BASELINE.json configs[3] at full size (4096 x 256 x 50 = 52.4 M pairs per query) against the oracle, plus
    size-independent properties: idempotence and invariance under an exact rigid motion of the whole problem"""
    lib, sc, org, q = W.c4_problem(nqueries=3)
    stage(eng, lib, sc, org)
    got, st = eng.query(q, None, MODE_PLANNER)
    want, ost = orc.Problem(lib, sc).check(q, None, 0, inner_parallel=True)     # reference early return: same decisions
    assert st["pairs_nominal"] == 3 * 4096 * 50 * (256 + 4) == ost["pairs_nominal"]
    assert np.array_equal(got, want)
    again, _ = eng.query(q, None, MODE_PLANNER)
    assert np.array_equal(got, again)
    # rotate the world by exactly 90 degrees about the origin and shift by integers: (x, y) -> (-y + 64, x - 32)
    rot = lambda a: np.stack([-a[..., 1] + 64.0, a[..., 0] - 32.0], -1)
    sc2 = dict(sc)
    sc2["obst"] = rot(sc["obst"])
    sc2["segs"] = np.concatenate([rot(sc["segs"][:, 0:2]), rot(sc["segs"][:, 2:4])], axis=1)
    q2 = q.copy()
    q2["x"], q2["y"] = -q["y"] + 64.0, q["x"] - 32.0
    q2["c"], q2["s"] = -q["s"], q["c"]
    stage(eng, lib, sc2, rot(org))
    got2, _ = eng.query(q2, None, MODE_PLANNER)
    # the rotation of the pose is exact only up to the rounding of the products, so compare through the oracle too
    want2, _ = orc.Problem(lib, sc2).check(q2, None, 0, inner_parallel=True)
    assert np.array_equal(got2, want2)
    assert (got2 != got).mean() < 1e-3


def test_error_behaviour(eng):
    e2 = CollisionEngine(0)
    with pytest.raises(MpcError):
        e2.query(make_queries(1), None, 0)                       # nothing staged
    lib, sc, org, q = W.adversarial_problem(1)
    stage(e2, lib, sc, org)
    bad = q.copy()
    bad["scene"] = 5
    with pytest.raises(MpcError):
        e2.query(bad, None, 0)
    bad = q.copy()
    bad["cand_set"] = -1
    with pytest.raises(MpcError):
        e2.query(bad, np.array([0, 1, 10 ** 6], np.int32), 0)
    tid = lib["tidx"].copy()
    tid[3, 2] += 1
    with pytest.raises(MpcError):
        e2.upload_library(lib["verts"], lib["nv"], tid)          # irregular slot must be re-slotted by the caller
    with pytest.raises(MpcError):
        CollisionEngine(99)
    e2.close()


def test_repeatability_under_load(eng):
    """the warp queues, shared-memory polygon records and atomics must not make results depend on scheduling:
    the same mid-size batch 25 times, in both cull modes, must give the identical mask every time"""
    lib, sc, org, q = W.c4_problem(P=1024, slots=24, T=20, nqueries=6)
    stage(eng, lib, sc, org)
    want, _ = orc.Problem(lib, sc).check(q, None, orc.MODE_NO_EARLY_EXIT, inner_parallel=True)
    for mode in (MODE_PLANNER, MODE_NO_EARLY_EXIT, MODE_NO_CULL):
        for _ in range(25 if mode != MODE_NO_CULL else 3):
            got, _ = eng.query(q, None, mode)
            assert np.array_equal(got, want)
    lib, sc, org, q = W.adversarial_problem(77, P=320, J=7, max_obst=90, nqueries=5)
    stage(eng, lib, sc, org)
    want, _ = orc.Problem(lib, sc).check(q, None, orc.MODE_NO_EARLY_EXIT)
    for _ in range(25):
        got, _ = eng.query(q, None, MODE_PLANNER)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("nring", [24, 300, 1500])
def test_parity_adversarial_vertex_alignments(eng, nring):
    """inside/outside parity of the occupancy centre when many boundary vertices share (exactly, or within 1e-12 ..
    1e-4 m) the centre's x or y: the float32 straddle tests then disagree with float64 for individual segments, and
    only their consistency between neighbouring segments keeps the crossing parity right.
    Star-shaped region with a star-shaped hole, both with `nring` vertices; every ray direction is exercised because
    the per-warp direction choice depends on the block layout."""
    rng = np.random.default_rng(nring)
    P, J = 160, 2
    verts = np.zeros((P, J, 4, 2))
    cx = rng.uniform(-30, 30, (P, J))
    cy = rng.uniform(-30, 30, (P, J))
    for p in range(P):
        for j in range(J):
            verts[p, j] = W.rect_vertices(np.array(cx[p, j]), np.array(cy[p, j]), np.array(rng.uniform(-3, 3)), 1.2, 0.6)[::-1]
    lib = dict(verts=verts, nv=np.full((P, J), 4, np.int32), tidx=np.tile(np.arange(J, dtype=np.int32), (P, 1)),
               set_off=np.array([0, P], np.int32), set_idx=np.arange(P, dtype=np.int32))
    pose = (3.25, -1.5)            # identity rotation: world centre = local centre + pose (up to one rounding)
    wx, wy = (verts[..., 0].mean(-1) + pose[0]).ravel(), (verts[..., 1].mean(-1) + pose[1]).ravel()

    def star(n, r0, r1, phase):
        ang = np.sort(rng.uniform(0, 2 * np.pi, n)) + phase
        r = rng.uniform(r0, r1, n)
        ring = np.stack([r * np.cos(ang) + pose[0], r * np.sin(ang) + pose[1]], -1)
        # align a third of the vertices with some centre coordinate, exactly or nearly
        # (the nearest centre coordinate is used, so the ring keeps its shape)
        for k in rng.choice(n, n // 3, replace=False):
            d = rng.choice([0.0, 1e-12, -1e-12, 1e-9, -1e-9, 1e-7, -1e-7, 1e-5, -1e-5, 1e-4, -1e-4])
            if rng.random() < 0.5:
                ring[k, 1] = wy[np.argmin(np.abs(wy - ring[k, 1]))] + d
            else:
                ring[k, 0] = wx[np.argmin(np.abs(wx - ring[k, 0]))] + d
        return ring

    outer, hole = star(nring, 38.0, 60.0, 0.0), star(nring, 4.0, 9.0, 0.3)
    seg = lambda r: np.concatenate([r, np.roll(r, -1, axis=0)], axis=1)
    segs = np.concatenate([seg(outer), seg(hole)])
    nsteps = 2
    sc = dict(scene_step_off=np.array([0, nsteps], np.int32), step_obst_off=np.zeros(nsteps + 1, np.int32),
              obst=np.zeros((0, 4, 2)), obst_nv=np.zeros(0, np.int32), step_seg_begin=np.zeros(nsteps, np.int32),
              step_seg_end=np.full(nsteps, len(segs), np.int32), segs=segs)
    q = make_queries(1)
    q['x'], q['y'], q['cand_set'], q['cand_cnt'] = pose[0], pose[1], 0, P
    for origin in (np.array([pose]), np.array([[pose[0] + 700.0, pose[1] - 900.0]])):    # tight and loose band
        got, want, st, ost = both(eng, lib, sc, origin, q, None, MODE_PLANNER)
        assert np.array_equal(got, want), "parity disagreements: %d" % (got != want).sum()
    assert 0 < want.mean() < 1


def test_dense_general_polygons_multi_phase_and_queue_pressure(eng):
    """8-gons on both sides, 300 obstacles per step (two tiles), candidates clustered on the obstacles so that most
    axis-0 tests pass: the per-warp queues fill and drain many times per tile"""
    rng = np.random.default_rng(8)
    lib, sc, org, q = W.adversarial_problem(500, P=128, J=3, nsteps=4, max_obst=5, lib_verts=8, obst_verts=8)
    nsteps, O = 4, 300
    obst = np.zeros((nsteps * O, 8, 2))
    nv = np.zeros(nsteps * O, np.int32)
    for k in range(nsteps * O):
        n = int(rng.integers(3, 9))
        poly = W.random_convex(rng, n, rng.uniform(1.0, 4.0)) + [rng.uniform(-10, 60), rng.uniform(-30, 30)]
        obst[k, :n] = poly if rng.random() < 0.5 else poly[::-1]
        nv[k] = n
    sc["obst"], sc["obst_nv"] = obst, nv
    sc["step_obst_off"] = (np.arange(nsteps + 1) * O).astype(np.int32)
    for mode in (MODE_PLANNER, MODE_VALIDATOR | MODE_NO_EARLY_EXIT, MODE_NO_CULL):
        got, want, st, ost = both(eng, lib, sc, org, q, None, mode)
        assert np.array_equal(got, want), mode
        assert st["pairs_nominal"] == ost["pairs_nominal"]
    got, st = eng.query(q, None, MODE_COUNT_WORK | MODE_NO_EARLY_EXIT)
    assert st["axis_tests"] > 0.01 * st["pairs_nominal"]          # the queue path really carried a lot of pairs


def test_refinement_list_overflow_is_recovered(monkeypatch):
    """more pairs inside the float32 band than the refinement list holds: the call grows the list and re-runs"""
    monkeypatch.setenv("MPC_WL_CAP", "8")
    e2 = CollisionEngine(0)
    monkeypatch.delenv("MPC_WL_CAP")
    lib, sc, org, q = W.adversarial_problem(7, P=200, max_obst=70)
    stage(e2, lib, sc, org)
    got, st = e2.query(q, None, MODE_PLANNER)
    want, _ = orc.Problem(lib, sc).check(q, None, orc.MODE_NO_EARLY_EXIT)
    assert np.array_equal(got, want)
    assert st["worklist_overflow"] == 1 and st["refined_fp64"] > 8
    got, st = e2.query(q, None, MODE_PLANNER)                       # the grown list is kept
    assert np.array_equal(got, want) and st["worklist_overflow"] == 0
    e2.close()


@pytest.mark.parametrize("mode", [MODE_PLANNER, MODE_VALIDATOR])
def test_adversarial_with_sorted_sets_and_block_culling(eng, mode):
    """candidate sets larger than 512 are kept spatially sorted on the device (results must come back in the caller's
    order) and steps with more than 32 obstacle slots are culled by block boxes: near-contact obstacles, whole-set
    queries, sub-range queries (which use the unsorted copy) and explicit lists must all agree with the oracle"""
    lib, sc, org, q = W.adversarial_problem(900 + mode, P=1300, J=4, nsteps=6, max_obst=90, nqueries=4)
    P = lib["verts"].shape[0]
    assert P // 2 > 512
    got, want, st, ost = both(eng, lib, sc, org, q, None, mode)           # whole sets: sorted path
    assert np.array_equal(got, want), (got != want).sum()
    assert st["pairs_nominal"] == ost["pairs_nominal"] and 0 < want.mean() < 1
    q2 = q.copy()
    q2["cand_off"], q2["cand_cnt"] = 17, P // 2 - 40                       # sub-ranges: caller's order
    got, want, _, _ = both(eng, lib, sc, org, q2, None, mode)
    assert np.array_equal(got, want)
    q3 = q.copy()
    rng = np.random.default_rng(1)
    cand = rng.permutation(P).astype(np.int32)
    q3["cand_set"], q3["cand_off"], q3["cand_cnt"] = -1, 0, P
    got, want, _, _ = both(eng, lib, sc, org, q3, cand, mode)
    assert np.array_equal(got, want)
    stage(eng, lib, sc, org)
    g2, st2 = eng.query(q, None, mode | MODE_COUNT_WORK | MODE_NO_EARLY_EXIT)
    assert st2["box_skipped"] > 0 and st2["cull_tests"] + st2["box_skipped"] + 0 >= 0


def test_early_return_across_time_slots_and_compaction(eng):
    """big batches run slot-major and skip tiles whose candidates already collide at another time step (the reference's
    early return); with a fully occupied library the undecided candidates are renumbered densely.  Every regime must
    give the oracle's bits and the exact nominal pair count: whole sorted sets (compaction), partial ranges and explicit
    lists (group-level skip), a library with empty slots (no compaction), validator mode, and no-exit mode."""
    lib, sc, org, q = W.c4_problem(nqueries=12)
    P = lib["verts"].shape[0]
    pb = orc.Problem(lib, sc)
    stage(eng, lib, sc, org)
    # (a) whole sets: 12 x 128 groups x 50 slots -> 128 groups per CTA, compaction on
    for mode, omode in ((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT)):
        got, st = eng.query(q, None, mode)
        want, ost = pb.check(q, None, omode, inner_parallel=True)
        assert np.array_equal(got, want)
        assert st["pairs_nominal"] == ost["pairs_nominal"] == 12 * P * 50 * (256 + 4)
    base = got
    noexit, st_ne = eng.query(q, None, MODE_VALIDATOR | MODE_NO_EARLY_EXIT)
    assert np.array_equal(noexit, base)
    cnt, st_c = eng.query(q, None, MODE_VALIDATOR | MODE_COUNT_WORK)
    assert np.array_equal(cnt, base)
    assert st_c["done_skipped"] > 0.3 * st_c["pairs_nominal"]                 # most of the work is never looked at
    assert st_c["done_skipped"] + st_c["box_skipped"] + st_c["cull_tests"] <= st_c["pairs_nominal"]
    # (b) ragged: partial ranges of the set, explicit lists, empty queries -- the group-level skip without compaction
    rng = np.random.default_rng(4)
    cand = rng.permutation(P).astype(np.int32)
    q2 = np.concatenate([q, q, q[:6]])
    for k in range(len(q2)):
        if k % 3 == 0:
            q2[k]["cand_off"], q2[k]["cand_cnt"] = 100 + k, P - 300 - 7 * k              # range inside the staged set
        elif k % 3 == 1:
            q2[k]["cand_set"], q2[k]["cand_off"], q2[k]["cand_cnt"] = -1, 5 * k, P - 5 * k   # explicit list
        if k == 7:
            q2[k]["cand_cnt"] = 0
    got, st = eng.query(q2, cand, MODE_PLANNER)
    want, ost = pb.check(q2, cand, 0, inner_parallel=True)
    assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    # (c) a library with holes: every 5th occupancy of every 3rd primitive is missing -> statistic needs the library
    lib2 = dict(lib, nv=lib["nv"].copy())
    lib2["nv"][::3, ::5] = 0
    stage(eng, lib2, sc, org)
    got, st = eng.query(q, None, MODE_PLANNER)
    want, ost = orc.Problem(lib2, sc).check(q, None, 0, inner_parallel=True)
    assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    # repeatability: which tiles are skipped depends on scheduling, the bits must not
    for _ in range(5):
        again, _ = eng.query(q, None, MODE_PLANNER)
        assert np.array_equal(again, got)


def test_early_return_with_multi_phase_tiles(eng):
    """768 obstacles per step do not fit one tile (cap 320): the step is processed in phases, and in a batch of more
    than one resident wave the group-level early return acts between them"""
    lib, sc, org, q = W.c4_problem(P=1024, O_lanes=8, slots=96, T=10, nqueries=24)
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    for mode, omode in ((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT)):
        got, st = eng.query(q, None, mode)
        want, ost = pb.check(q, None, omode, inner_parallel=True)
        assert np.array_equal(got, want)
        assert st["pairs_nominal"] == ost["pairs_nominal"] == 24 * 1024 * 10 * (768 + 4)
    assert st["ctas"] > 148 * 4
    noexit, _ = eng.query(q, None, MODE_VALIDATOR | MODE_NO_EARLY_EXIT)
    assert np.array_equal(noexit, got)
    assert 0.02 < got.mean() < 0.999


@pytest.mark.parametrize("seed", range(20))
def test_randomised_batch_shapes_against_oracle(eng, seed):
    """random batch shapes around the regime switches of the launch heuristics (groups per CTA, chunks per query, slot-
    major grid, early return with and without compaction, tiles in one or several phases): library size, horizon,
    obstacles per step, batch size, ragged candidate ranges, start indices and step limits are all drawn at random;
    bits and the nominal pair count must equal the oracle's in planner and validator mode"""
    rng = np.random.default_rng(1000 + seed)
    P = int(rng.choice([160, 500, 1024, 2080, 3000]))
    T = int(rng.integers(3, 14))
    slots = int(rng.choice([4, 12, 32, 45]))
    B = int(rng.choice([1, 3, 9, 20, 40]))
    lib, sc, org, q = W.c4_problem(P=P, O_lanes=8, slots=slots, T=T, nqueries=B)
    for k in range(B):
        r = rng.random()
        if r < 0.3:                                   # a range inside the (single) staged set
            lo = int(rng.integers(0, P // 2))
            q[k]["cand_off"], q[k]["cand_cnt"] = lo, int(rng.integers(0, P - lo + 1))
        elif r < 0.5:                                 # explicit list
            lo = int(rng.integers(0, P // 2))
            q[k]["cand_set"], q[k]["cand_off"], q[k]["cand_cnt"] = -1, lo, int(rng.integers(1, P - lo + 1))
        q[k]["t0"] = int(rng.integers(0, 3))
        q[k]["step_limit"] = int(rng.choice([T + 5, T - 2, 1]))
    cand = rng.permutation(P).astype(np.int32)
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    for mode, omode in ((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT)):
        got, st = eng.query(q, cand, mode)
        want, ost = pb.check(q, cand, omode | orc.MODE_NO_EARLY_EXIT)
        assert np.array_equal(got, want), (P, T, slots, B, mode)
        assert st["pairs_nominal"] == ost["pairs_nominal"], (P, T, slots, B, mode)


def test_compaction_with_general_polygons(eng):
    """the compaction path in the 8-vertex instantiations: 200 queries x 1024 candidates (32 groups per CTA) of an
    adversarial library of general convex polygons against near-touching general obstacles"""
    lib, sc, org, q = W.adversarial_problem(21, P=2048, J=8, nsteps=8, max_obst=12, lib_verts=8, obst_verts=8, nqueries=200,
                                            with_region=False)
    assert (lib["nv"] > 0).mean() < 1.0                    # the generator leaves holes: no compaction with this library
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    got, st = eng.query(q, None, MODE_PLANNER)
    want, ost = pb.check(q, None, orc.MODE_NO_EARLY_EXIT)
    assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    # fill the holes (copy the previous occupancy): now every slot is occupied and the compaction applies
    lib2 = dict(lib, verts=lib["verts"].copy(), nv=lib["nv"].copy())
    for p_, j_ in zip(*np.nonzero(lib2["nv"] == 0)):
        src = np.nonzero(lib["nv"][p_] > 0)[0]
        k = src[0] if len(src) else None
        if k is None:
            lib2["verts"][p_, j_, :3] = [[0, 0], [1, 0], [0, 1]]
            lib2["nv"][p_, j_] = 3
        else:
            lib2["verts"][p_, j_] = lib["verts"][p_, k]
            lib2["nv"][p_, j_] = lib["nv"][p_, k]
    stage(eng, lib2, sc, org)
    for mode, omode in ((MODE_PLANNER, 0), (MODE_VALIDATOR, orc.MODE_STRICT)):
        got, st = eng.query(q, None, mode)
        want, ost = orc.Problem(lib2, sc).check(q, None, omode | orc.MODE_NO_EARLY_EXIT)
        assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    assert st["ctas"] > 148 * 4
    assert 0.02 < got.mean() < 0.98


def test_lanes_match_single_context(eng):
    """independent cycles (scene + queries) dealt to several contexts on one GPU, one host thread each: the staging of
    one lane overlaps the check of another; every cycle's bits equal the oracle's and the single-context result"""
    from motionplanner_b200.engine import unpack_mask
    from motionplanner_b200.lanes import CollisionLanes
    parts = [W.adversarial_problem(700 + i, P=96, nsteps=4 + (i % 3), offset=(50.0 * i, 20.0 * i)) for i in range(7)]
    lib = parts[0][0]
    lanes = CollisionLanes(0, lanes=3)
    try:
        lanes.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib.get("set_off"), lib.get("set_idx"))
        cycles = [(dict(sc, origins=org), q) for _, sc, org, q in parts]
        for rep in range(3):                                  # repeated: lanes reuse their staged buffers and graphs
            res = lanes.run_cycles(cycles, MODE_PLANNER)
            for (_, sc, org, q), (mask, st) in zip(parts, res):
                want, _ = orc.Problem(lib, sc).check(q, None, orc.MODE_NO_EARLY_EXIT)
                assert np.array_equal(unpack_mask(mask, q), want)
        for (_, sc, org, q), (mask, _) in zip(parts, res):
            stage(eng, lib, sc, org)
            single, _ = eng.query(q, None, MODE_PLANNER)
            assert np.array_equal(unpack_mask(mask, q), single)
        # a failing cycle stops the lanes and surfaces as the same exception a single context raises
        bad = dict(cycles[2][0])
        bad["obst_nv"] = np.full_like(bad["obst_nv"], 99)
        with pytest.raises(MpcError):
            lanes.run_cycles(cycles[:2] + [(bad, cycles[2][1])] + cycles[3:], MODE_PLANNER)
        res2 = lanes.run_cycles(cycles, MODE_PLANNER)         # still usable afterwards
        assert all(np.array_equal(a[0], b[0]) for a, b in zip(res, res2))
    finally:
        lanes.close()


def test_scout_launch_regime(eng, monkeypatch):
    """batches whose time slots fill only a fraction of the resident wave launch the first slot of the order on its own
    (finer CTAs), the rest with its bits visible: bits and the nominal pair count equal the oracle's; forcing 0 / 2 / 3
    scout launches and other scout CTA sizes on the same batch changes nothing"""
    lib, sc, org, q = W.c4_problem(P=4096, O_lanes=8, slots=16, T=12, nqueries=32)
    rng = np.random.default_rng(5)
    q["t0"][5], q["step_limit"][7], q["cand_cnt"][9], q["cand_cnt"][11] = 2, 6, 1000, 0
    q["cand_off"][13], q["cand_cnt"][13] = 77, 3001
    stage(eng, lib, sc, org)
    pb = orc.Problem(lib, sc)
    want, ost = pb.check(q, None, orc.MODE_NO_EARLY_EXIT)
    got, st = eng.query(q, None, MODE_PLANNER)
    assert st["kernels"] == 4                                  # clear + scout + main + refine: the heuristic picked the regime
    assert np.array_equal(got, want) and st["pairs_nominal"] == ost["pairs_nominal"]
    wantv, _ = pb.check(q, None, orc.MODE_STRICT | orc.MODE_NO_EARLY_EXIT)
    gotv, _ = eng.query(q, None, MODE_VALIDATOR)
    assert np.array_equal(gotv, wantv)
    for scouts, sgpc in ((0, None), (2, None), (3, 8), (1, 40), (1, 128)):
        monkeypatch.setenv("MPC_SCOUTS", str(scouts))
        if sgpc is None:
            monkeypatch.delenv("MPC_SCOUT_GPC", raising=False)
        else:
            monkeypatch.setenv("MPC_SCOUT_GPC", str(sgpc))
        got2, st2 = eng.query(q, None, MODE_PLANNER)
        assert st2["kernels"] == 3 + scouts
        assert np.array_equal(got2, want) and st2["pairs_nominal"] == ost["pairs_nominal"], (scouts, sgpc)
        gotv2, _ = eng.query(q, None, MODE_VALIDATOR)
        assert np.array_equal(gotv2, wantv), (scouts, sgpc)
    # small problems with forced scouts (multi-phase tiles, general polygons, explicit lists)
    monkeypatch.setenv("MPC_SCOUTS", "2")
    monkeypatch.delenv("MPC_SCOUT_GPC", raising=False)
    for seed, kw in ((31, dict()), (32, dict(lib_verts=9, obst_verts=7)), (33, dict(max_obst=400, P=40, J=9, nsteps=10))):
        lib2, sc2, org2, q2 = W.adversarial_problem(seed, **kw)
        got3, want3, st3, ost3 = both(eng, lib2, sc2, org2, q2, None, MODE_PLANNER)
        assert np.array_equal(got3, want3) and st3["pairs_nominal"] == ost3["pairs_nominal"]
