"""Pins the fast float64 C oracle (oracle/oracle_c.c: float filter + exact expansion arithmetic) against the
definitional rational oracle (oracle/exact.py) on adversarial inputs."""
import math
import random

import numpy as np

from oracle import exact as ex
from oracle import oracle as orc
from test_oracle_kat import rand_convex, rand_star, CONVEX_KAT, REGION_KAT


def test_orient2d_adversarial():
    rng = random.Random(7)
    n_zero = n_exact0 = 0
    before = orc.lib().orc_exact_calls()
    for _ in range(20000):
        # points (nearly) on a common line, large offsets, one-ulp nudges
        ox, oy = rng.uniform(-1500, 1500), rng.uniform(-1500, 1500)
        dx, dy = rng.uniform(-30, 30), rng.uniform(-30, 30)
        t1, t2 = rng.uniform(-3, 3), rng.uniform(-3, 3)
        a = (ox, oy)
        b = (ox + t1 * dx, oy + t1 * dy)
        c = (ox + t2 * dx, oy + t2 * dy)
        k = rng.randint(0, 3)
        if k == 1:
            c = (math.nextafter(c[0], math.inf), c[1])
        elif k == 2:
            c = (c[0], math.nextafter(c[1], -math.inf))
        elif k == 3:   # exactly collinear on a dyadic lattice
            a, b, c = (ox := round(ox), oy := round(oy)), (ox + 3.0, oy + 1.5), (ox + 6.0 * 7, oy + 3.0 * 7)
        want = ex.orient(a, b, c)
        assert orc.orient2d(a, b, c) == want
        n_zero += want == 0
    assert n_zero > 1000
    assert orc.lib().orc_exact_calls() - before > 5000      # the exact branch really ran


def test_convex_collide_matches_exact():
    rng = random.Random(11)
    for name, A, B, closed, open_ in CONVEX_KAT:
        assert orc.convex_collide(ex.strip_ring(A), ex.strip_ring(B), True) is closed, name
        assert orc.convex_collide(ex.strip_ring(A), ex.strip_ring(B), False) is open_, name
    for _ in range(4000):
        span = rng.choice((2, 3, 6))
        A, B = rand_convex(rng, span), rand_convex(rng, span)
        if rng.random() < 0.5:     # move off the lattice by a rigid motion in float64: ties become near-ties
            th = rng.uniform(0, 2 * math.pi)
            tx, ty = rng.uniform(-1000, 1000), rng.uniform(-1000, 1000)
            A = ex.affine_transform(A, math.cos(th), math.sin(th), tx, ty)
            B = ex.affine_transform(B, math.cos(th), math.sin(th), tx, ty)
        assert orc.convex_collide(A, B, True) == ex.convex_intersects(A, B)
        assert orc.convex_collide(A, B, False) == ex.convex_interiors_overlap(A, B)


def test_segment_and_point_predicates_match_exact():
    rng = random.Random(5)
    for _ in range(3000):
        P = rand_convex(rng, 4)
        s0 = (rng.randint(0, 12) / 2.0, rng.randint(0, 12) / 2.0)
        s1 = (rng.randint(0, 12) / 2.0, rng.randint(0, 12) / 2.0)
        assert orc.segment_meets_interior(P, s0 + s1) == ex.segment_meets_interior(s0, s1, P)
    for _ in range(1500):
        S = rand_star(rng)
        segs = [S[i] + S[(i + 1) % len(S)] for i in range(len(S))]
        p = (rng.randint(0, 16) / 2.0, rng.randint(0, 16) / 2.0)
        want = {1: 1, 0: 2, -1: 0}[ex.point_in_rings(p, [S])]
        assert orc.point_in_soup(p, segs) == want


def _tiny_problem(rng, mode_strict):
    """one scene, 3 steps, lattice obstacles and a lattice road; library of rectangles/triangles"""
    P, J, Vl = 12, 3, 4
    verts = np.zeros((P, J, Vl, 2))
    nv = np.zeros((P, J), np.int32)
    tidx = np.zeros((P, J), np.int32)
    for p in range(P):
        for j in range(J):
            if rng.random() < 0.15:
                continue
            x0, y0 = rng.randint(-2, 6) / 2.0, rng.randint(-2, 6) / 2.0
            w, h = rng.randint(1, 4) / 2.0, rng.randint(1, 4) / 2.0
            r = [(x0, y0), (x0, y0 + h), (x0 + w, y0 + h), (x0 + w, y0)]      # clockwise like the reference car
            if rng.random() < 0.3:
                r = r[:3]
            verts[p, j, :len(r)] = r
            nv[p, j] = len(r)
            tidx[p, j] = j + (1 if rng.random() < 0.1 else 0)
    lib = dict(verts=verts, nv=nv, tidx=tidx, set_off=np.array([0, P], np.int32), set_idx=np.arange(P, dtype=np.int32))
    nsteps = 4
    obst, onv, off = [], [], [0]
    for t in range(nsteps):
        for _ in range(rng.randint(0, 4)):
            h = rand_convex(rng, 5)[:4]
            while len(h) < 3:
                h = rand_convex(rng, 5)[:4]
            v = np.zeros((4, 2))
            v[:len(h)] = h
            obst.append(v)
            onv.append(len(h))
        off.append(len(obst))
    road = [(0.0, 0.0), (8.0, 0.0), (8.0, 5.0), (0.0, 5.0)]
    segs = np.array([road[i] + road[(i + 1) % 4] for i in range(4)])
    scenes = dict(scene_step_off=np.array([0, nsteps], np.int32), step_obst_off=np.array(off, np.int32),
                  obst=np.array(obst).reshape(-1, 4, 2), obst_nv=np.array(onv, np.int32),
                  step_seg_begin=np.array([0, 0, -1, 0], np.int32), step_seg_end=np.array([4, 4, -1, 4], np.int32),
                  segs=segs)
    return lib, scenes, road


def _python_restatement(lib, scenes, road, q, strict):
    """reference loop (collision_check, src/lowLevelPlannerManeuverAutomaton.py:95-125) on exact predicates"""
    out = []
    nsteps = scenes["scene_step_off"][1]
    for p in range(lib["verts"].shape[0]):
        collide = False
        for j in range(lib["verts"].shape[1]):
            n = lib["nv"][p, j]
            if n == 0:
                continue
            t = q["t0"] + lib["tidx"][p, j]
            if t > q["step_limit"] or t >= nsteps:
                continue
            A = ex.affine_transform([tuple(v) for v in lib["verts"][p, j, :n]], q["c"], q["s"], q["x"], q["y"])
            ok = True
            if scenes["step_seg_begin"][t] >= 0:
                ok = ex.region_contains_convex([road], A)
            for o in range(scenes["step_obst_off"][t], scenes["step_obst_off"][t + 1]):
                B = [tuple(v) for v in scenes["obst"][o, :scenes["obst_nv"][o]]]
                hit = ex.convex_intersects(A, B) if strict else ex.convex_interiors_overlap(A, B)
                ok = ok and not hit
            if not ok:
                collide = True
                break
        out.append(collide)
    return np.array(out, np.uint8)


def test_check_batch_matches_python_restatement():
    rng = random.Random(21)
    for trial in range(30):
        strict = trial % 2
        lib, scenes, road = _tiny_problem(rng, strict)
        pb = orc.Problem(lib, scenes)
        qs = np.zeros(3, orc.QUERY_DTYPE)
        for i in range(3):
            th = rng.choice((0.0, math.pi / 2, rng.uniform(0, 6.28)))
            c, s = (1.0, 0.0) if th == 0.0 else ((0.0, 1.0) if th == math.pi / 2 else (math.cos(th), math.sin(th)))
            qs[i] = (rng.randint(0, 8) / 2.0, rng.randint(0, 6) / 2.0, c, s, 0, rng.randint(0, 2), rng.choice((2, 3, 9)),
                     0, 0, 12, 0, 0)
        got, st = pb.check(qs, None, mode=orc.MODE_STRICT if strict else 0)
        want = np.concatenate([_python_restatement(lib, scenes, road, qs[i], strict) for i in range(3)])
        assert (got == want).all(), trial
        got2, _ = pb.check(qs, None, mode=(orc.MODE_STRICT if strict else 0) | orc.MODE_NO_EARLY_EXIT, inner_parallel=True)
        assert (got2 == want).all()
        assert st["pairs_decided"] <= st["pairs_nominal"]
