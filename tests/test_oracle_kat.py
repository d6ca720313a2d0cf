"""Known-answer tests that pin the oracle (oracle/exact.py) -- answers derived by hand from the DE-9IM definitions of
shapely ``contains`` / ``intersects`` (reference call sites: auxiliary/collisionChecker.py:21,32,
src/lowLevelPlannerManeuverAutomaton.py:122) -- plus randomized agreement with an independent exact formulation
(rational clipping area / explicit vertex-edge incidence)."""
import random
from fractions import Fraction

import pytest

from oracle import exact as ex


def rect(x0, y0, x1, y1):
    return [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]


SQ = rect(0, 0, 1, 1)

# (name, A, B, closed sets intersect, open interiors overlap)
CONVEX_KAT = [
    ("disjoint", SQ, rect(2, 0, 3, 1), False, False),
    ("edge touch", SQ, rect(1, 0, 2, 1), True, False),
    ("partial edge touch", SQ, rect(1, 0.5, 2, 1.5), True, False),
    ("vertex touch", SQ, rect(1, 1, 2, 2), True, False),
    ("overlap", SQ, rect(0.5, 0.5, 1.5, 1.5), True, True),
    ("contained", SQ, rect(0.25, 0.25, 0.75, 0.75), True, True),
    ("identical", SQ, SQ, True, True),
    ("clockwise operand", SQ[::-1], rect(0.5, 0.5, 1.5, 1.5)[::-1], True, True),
    ("diamond vertex on edge", SQ, [(1, 0.5), (2, 0), (3, 0.5), (2, 1)], True, False),
    ("diamond vertex inside", SQ, [(0.9, 0.5), (2, 0), (3, 0.5), (2, 1)], True, True),
    ("diamond clear by a hair", SQ, [(1.0000000000000002, 0.5), (2, 0), (3, 0.5), (2, 1)], False, False),
    ("triangle edge through corner", SQ, [(1, 2), (2, 1), (3, 3)], False, False),
    ("triangle edge on corner", SQ, [(0, 2), (2, 0), (3, 3)], True, False),
    ("cross shape overlap", rect(-1, 0.4, 2, 0.6), rect(0.4, -1, 0.6, 2), True, True),
    ("closed ring input", SQ + [SQ[0]], rect(1, 0, 2, 1) + [(1, 0)], True, False),
    ("collinear extra vertex", [(0, 0), (0.5, 0), (1, 0), (1, 1), (0, 1)], rect(0.25, -1, 0.75, 0), True, False),
    ("far apart large coords", rect(1000, 1000, 1001, 1001), rect(-1000, -1000, -999, -999), False, False),
    ("one ulp overlap at 1e3", rect(1000, 0, 1001, 1), rect(1000.9999999999999, 0, 1002, 1), True, True),
    ("one ulp gap at 1e3", rect(1000, 0, 1001, 1), rect(1001.0000000000001, 0, 1002, 1), False, False),
]


@pytest.mark.parametrize("name,A,B,closed,open_", CONVEX_KAT, ids=[k[0] for k in CONVEX_KAT])
def test_convex_kat(name, A, B, closed, open_):
    assert ex.convex_intersects(A, B) is closed
    assert ex.convex_intersects(B, A) is closed
    assert ex.convex_interiors_overlap(A, B) is open_
    assert ex.convex_interiors_overlap(B, A) is open_


ROAD = rect(0, 0, 10, 4)
HOLE = rect(4, 1, 5, 2)
LSHAPE = [(0, 0), (6, 0), (6, 2), (2, 2), (2, 6), (0, 6)]

# (name, rings, P, region.contains(P))
REGION_KAT = [
    ("inside", [ROAD], rect(1, 1, 3, 2), True),
    ("touching edge from inside (Q1)", [ROAD], rect(1, 0, 3, 1), True),
    ("touching corner from inside", [ROAD], rect(0, 0, 2, 1), True),
    ("equal to region", [ROAD], ROAD, True),
    ("crossing boundary", [ROAD], rect(1, -0.5, 3, 1), False),
    ("outside", [ROAD], rect(1, 5, 3, 6), False),
    ("outside touching edge", [ROAD], rect(1, 4, 3, 5), False),
    ("bigger than region", [ROAD], rect(-1, -1, 11, 5), False),
    ("hole swallowed", [ROAD, HOLE], rect(3, 0.5, 6, 2.5), False),
    ("touching hole from outside", [ROAD, HOLE], rect(2, 1, 4, 2), True),
    ("overlapping hole", [ROAD, HOLE], rect(3.5, 1.2, 4.5, 1.8), False),
    ("inside hole", [ROAD, HOLE], rect(4.2, 1.2, 4.8, 1.8), False),
    ("equal to hole", [ROAD, HOLE], HOLE, False),
    ("hole given clockwise", [ROAD, HOLE[::-1]], rect(2, 1, 4, 2), True),
    ("multipolygon second part", [ROAD, rect(20, 0, 30, 4)], rect(21, 1, 23, 2), True),
    ("multipolygon spanning the gap", [ROAD, rect(20, 0, 30, 4)], rect(9, 1, 21, 2), False),
    ("L-shape arm", [LSHAPE], rect(3, 0.5, 5, 1.5), True),
    ("L-shape notch", [LSHAPE], rect(3, 3, 5, 5), False),
    ("L-shape reflex vertex touch", [LSHAPE], rect(0.5, 0.5, 2, 2), True),
    ("L-shape across the notch corner", [LSHAPE], rect(1, 1, 3, 3), False),
    ("L-shape triangle hugging the reflex corner", [LSHAPE], [(1, 2), (2, 2), (2, 1)], True),
    ("empty region", [], rect(0, 0, 1, 1), False),
    ("rotated car fits lane", [rect(0, 0, 100, 3.5)], ex.rectangle_vertices(4.3, 1.7, 50, 1.75, 0.995, 0.0998), True),
    ("rotated car pokes out", [rect(0, 0, 100, 3.5)], ex.rectangle_vertices(4.3, 1.7, 50, 1.75, 0.8, 0.6), False),
]


@pytest.mark.parametrize("name,rings,P,expect", REGION_KAT, ids=[k[0] for k in REGION_KAT])
def test_region_contains_kat(name, rings, P, expect):
    assert ex.region_contains_convex(rings, P) is expect
    assert ex.region_contains_convex(rings, P[::-1]) is expect


def test_point_in_rings_kat():
    assert ex.point_in_rings((0.5, 0.5), [SQ]) == 1
    assert ex.point_in_rings((1.5, 0.5), [SQ]) == -1
    assert ex.point_in_rings((1.0, 0.5), [SQ]) == 0
    assert ex.point_in_rings((1.0, 1.0), [SQ]) == 0
    assert ex.point_in_rings((-1.0, 1.0), [SQ]) == -1          # ray through a vertex row
    assert ex.point_in_rings((-1.0, 0.0), [SQ]) == -1
    assert ex.point_in_rings((1.0, 3.0), [LSHAPE]) == 1
    assert ex.point_in_rings((3.0, 3.0), [LSHAPE]) == -1
    assert ex.point_in_rings((4.5, 1.5), [ROAD, HOLE]) == -1
    assert ex.point_in_rings((1.0, 2.0), [[(0, 0), (2, 2), (0, 4)]]) == 1    # ray hits the apex vertex


def test_affine_and_rectangle_vertices():
    # 90 degrees: (x, y) -> (-y, x) + t   exactly representable
    pts = ex.affine_transform([(1.0, 0.0), (0.0, 2.0)], 0.0, 1.0, 10.0, 20.0)
    assert pts == [(10.0, 21.0), (8.0, 20.0)]
    v = ex.rectangle_vertices(4.0, 2.0, 1.0, 1.0, 1.0, 0.0)
    assert v == [(-1.0, 0.0), (-1.0, 2.0), (3.0, 2.0), (3.0, 0.0)]


# ----------------------------------------------------------------------------------------------------------------------
# randomized agreement between independent exact formulations on degenerate-rich inputs (half-integer lattice)
# ----------------------------------------------------------------------------------------------------------------------

def hull(points):
    pts = sorted(set(points))
    if len(pts) < 3:
        return None

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lo, up = [], []
    for p in pts:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], p) <= 0:
            lo.pop()
        lo.append(p)
    for p in reversed(pts):
        while len(up) >= 2 and cross(up[-2], up[-1], p) <= 0:
            up.pop()
        up.append(p)
    h = lo[:-1] + up[:-1]
    return h if len(h) >= 3 else None


def rand_convex(rng, span=6):
    while True:
        h = hull([(rng.randint(0, 2 * span) / 2.0, rng.randint(0, 2 * span) / 2.0) for _ in range(rng.randint(3, 7))])
        if h:
            return h


def seg_seg_closed(p, q, r, s):
    o1, o2, o3, o4 = ex.orient(p, q, r), ex.orient(p, q, s), ex.orient(r, s, p), ex.orient(r, s, q)
    if o1 * o2 < 0 and o3 * o4 < 0:
        return True

    def on(a, b, c):
        return ex.orient(a, b, c) == 0 and min(a[0], b[0]) <= c[0] <= max(a[0], b[0]) and \
            min(a[1], b[1]) <= c[1] <= max(a[1], b[1])

    return on(p, q, r) or on(p, q, s) or on(r, s, p) or on(r, s, q)


def closed_intersect_bruteforce(A, B):
    A, B = ex.ccw(A), ex.ccw(B)
    inA = lambda p: all(ex.orient(A[i], A[(i + 1) % len(A)], p) >= 0 for i in range(len(A)))
    inB = lambda p: all(ex.orient(B[i], B[(i + 1) % len(B)], p) >= 0 for i in range(len(B)))
    if any(inB(p) for p in A) or any(inA(p) for p in B):
        return True
    for i in range(len(A)):
        for k in range(len(B)):
            if seg_seg_closed(A[i], A[(i + 1) % len(A)], B[k], B[(k + 1) % len(B)]):
                return True
    return False


def test_random_convex_pairs_agree_with_clipping_area():
    rng = random.Random(1234)
    n_touch = n_overlap = 0
    for _ in range(1500):
        span = rng.choice((2, 3, 6))
        A, B = rand_convex(rng, span), rand_convex(rng, span)
        area = ex.clip_area2(A, B)
        open_ = ex.convex_interiors_overlap(A, B)
        closed = ex.convex_intersects(A, B)
        assert open_ == (area > 0)
        assert closed == closed_intersect_bruteforce(A, B)
        assert closed or not open_
        n_touch += closed and not open_
        n_overlap += open_
    assert n_touch > 50 and n_overlap > 200          # the lattice really produces the degenerate cases


def rand_star(rng, span=8):
    """random simple (star-shaped) polygon on the half-integer lattice"""
    c = (span / 2.0 + 0.25, span / 2.0 + 0.25)
    import math
    while True:
        pts = list({(rng.randint(0, 2 * span) / 2.0, rng.randint(0, 2 * span) / 2.0) for _ in range(rng.randint(4, 9))})
        if len(pts) < 4:
            continue
        pts.sort(key=lambda p: math.atan2(p[1] - c[1], p[0] - c[0]))
        # reject rings with collinear consecutive triples folding back / zero area
        if abs(ex.signed_area2(pts)) == 0:
            continue
        ok = all(ex.orient(c, pts[i], pts[(i + 1) % len(pts)]) > 0 for i in range(len(pts)))
        if ok:
            return pts


def test_random_region_contains_agrees_with_clipping_area():
    rng = random.Random(99)
    n_true = 0
    for _ in range(1200):
        S = rand_star(rng)
        P = rand_convex(rng, span=4)
        dx, dy = rng.randint(0, 6) / 2.0, rng.randint(0, 6) / 2.0
        P = [(x + dx, y + dy) for x, y in P]
        got = ex.region_contains_convex([S], P)
        want = ex.clip_area2(S, P) == abs(ex.signed_area2(P))
        assert got == want, (S, P)
        assert ex.simple_polygon_interior_overlaps_convex(S, P) == (ex.clip_area2(S, P) > 0)
        n_true += got
    assert n_true > 20


def test_simple_polygon_intersects_convex():
    assert ex.simple_polygon_intersects_convex(LSHAPE, rect(3, 3, 5, 5)) is False
    assert ex.simple_polygon_intersects_convex(LSHAPE, rect(2, 2, 5, 5)) is True       # touches the reflex vertex
    assert ex.simple_polygon_interior_overlaps_convex(LSHAPE, rect(2, 2, 5, 5)) is False
    assert ex.simple_polygon_intersects_convex(LSHAPE, rect(0.5, 0.5, 1, 1)) is True   # car inside
    assert ex.simple_polygon_intersects_convex(rect(0.5, 0.5, 1, 1), rect(-5, -5, 5, 5)) is True  # obstacle inside car
