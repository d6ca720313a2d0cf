"""TEST INFRASTRUCTURE ONLY -- definitional exact-arithmetic oracle (rationals).

This module is the ground truth of the parity suite.  Nothing under
``motionplanner_b200/`` may import it; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline leg do.

It restates, in exact rational arithmetic (``fractions.Fraction`` built from the
float64 inputs, so there is no rounding anywhere after the vertex coordinates
are fixed), the three third-party predicates the reference's hot path uses
(reference: SURVEY.md section 8 row R10):

* ``shapely.affinity.affine_transform(geom, [a, b, d, e, xoff, yoff])``
  (call sites ``src/lowLevelPlannerManeuverAutomaton.py:121``,
  ``src/maneuverAutomatonPlannerStandalone.py:222``,
  ``maneuverAutomaton/createManeuverAutomaton.py:84``) -- float64, per point
  ``x' = a*x + b*y + xoff`` evaluated left to right.
* ``space.contains(pgon)`` (``src/lowLevelPlannerManeuverAutomaton.py:122``,
  ``auxiliary/collisionChecker.py:21``) -- DE-9IM ``T*****FF*``: no point of
  ``pgon`` in the exterior of ``space`` and the interiors meet.
* ``obstacle.intersects(car)`` (``auxiliary/collisionChecker.py:32``,
  ``src/maneuverAutomatonPlannerStandalone.py:177``) -- closed sets share a point.

PARITY UNPINNED against shapely/GEOS itself: neither is installable in the build
container and the reference ships no golden vectors for these predicates
(SURVEY.md section 4 / 8c).  The oracle is instead pinned by
(i) hand-derived known-answer cases (``tests/test_oracle_kat.py``) and
(ii) an independent exact formulation (rational Sutherland-Hodgman clipping
area, ``clip_area`` below) that must agree with the orientation-sign formulation
on randomized degenerate inputs.
"""
from fractions import Fraction


def F(x):
    """exact rational value of a float64 (or int / Fraction)"""
    return x if isinstance(x, Fraction) else Fraction(x)


def orient(a, b, c):
    """exact sign of the 2x2 determinant |b-a, c-a| : +1 = c left of a->b, 0 = collinear"""
    d = (F(b[0]) - F(a[0])) * (F(c[1]) - F(a[1])) - (F(b[1]) - F(a[1])) * (F(c[0]) - F(a[0]))
    return (d > 0) - (d < 0)


def signed_area2(poly):
    """exact twice the signed area of a ring (list of (x, y)); > 0 = counter-clockwise"""
    s = Fraction(0)
    n = len(poly)
    for i in range(n):
        x0, y0 = F(poly[i][0]), F(poly[i][1])
        x1, y1 = F(poly[(i + 1) % n][0]), F(poly[(i + 1) % n][1])
        s += x0 * y1 - x1 * y0
    return s


def strip_ring(poly):
    """drop the closing duplicate vertex and consecutive duplicate vertices"""
    pts = [(float(p[0]), float(p[1])) for p in poly]
    out = []
    for p in pts:
        if not out or out[-1] != p:
            out.append(p)
    if len(out) > 1 and out[0] == out[-1]:
        out.pop()
    return out


def ccw(poly):
    """return the ring in counter-clockwise order (exact orientation test)"""
    poly = strip_ring(poly)
    return poly if signed_area2(poly) > 0 else poly[::-1]


# ----------------------------------------------------------------------------------------------------------------------
# convex / convex (separating-axis theorem with exact orientation signs)
# ----------------------------------------------------------------------------------------------------------------------

def _separated_by_edge_of(A, B, strict):
    """is there an edge of the CCW convex polygon A whose supporting line has all of B on its outer side?
    strict=True: B strictly outside (closed sets disjoint);  strict=False: B outside or on the line."""
    n = len(A)
    for i in range(n):
        a0, a1 = A[i], A[(i + 1) % n]
        if a0 == a1:
            continue
        if strict:
            if all(orient(a0, a1, q) < 0 for q in B):
                return True
        else:
            if all(orient(a0, a1, q) <= 0 for q in B):
                return True
    return False


def convex_intersects(A, B):
    """closed convex polygons share at least one point (shapely ``intersects``; validator semantics,
    reference auxiliary/collisionChecker.py:32)"""
    A, B = ccw(A), ccw(B)
    return not (_separated_by_edge_of(A, B, True) or _separated_by_edge_of(B, A, True))


def convex_interiors_overlap(A, B):
    """open interiors of two convex polygons share a point (planner semantics: the occupancy may touch an
    obstacle that was subtracted from the free space, reference src/lowLevelPlannerManeuverAutomaton.py:122 with
    space = road - obstacles, src/maneuverAutomatonPlannerStandalone.py:151-182)"""
    A, B = ccw(A), ccw(B)
    return not (_separated_by_edge_of(A, B, False) or _separated_by_edge_of(B, A, False))


# ----------------------------------------------------------------------------------------------------------------------
# segment / open convex polygon, point in ring soup, containment in a general polygonal region
# ----------------------------------------------------------------------------------------------------------------------

def segment_meets_interior(s0, s1, P):
    """closed segment s0-s1 has a point in the open interior of the convex polygon P"""
    P = ccw(P)
    n = len(P)
    for i in range(n):
        a0, a1 = P[i], P[(i + 1) % n]
        if a0 == a1:
            continue
        if orient(a0, a1, s0) <= 0 and orient(a0, a1, s1) <= 0:
            return False
    if (s0[0], s0[1]) == (s1[0], s1[1]):
        return True            # a point that no edge line excludes is interior
    o = [orient(s0, s1, p) for p in P]
    if all(v >= 0 for v in o) or all(v <= 0 for v in o):
        return False
    return True


def point_in_rings(p, rings):
    """even-odd rule over a list of rings.  returns +1 inside, 0 on a boundary, -1 outside (exact)"""
    px, py = F(p[0]), F(p[1])
    inside = False
    for ring in rings:
        ring = strip_ring(ring)
        n = len(ring)
        for i in range(n):
            a, b = ring[i], ring[(i + 1) % n]
            ax, ay, bx, by = F(a[0]), F(a[1]), F(b[0]), F(b[1])
            o = orient(a, b, p)
            if o == 0 and min(ax, bx) <= px <= max(ax, bx) and min(ay, by) <= py <= max(ay, by):
                return 0
            if (ay > py) != (by > py):
                # upward edge: crossing is to the right of p iff p is left of a->b
                if (by > ay and o > 0) or (by < ay and o < 0):
                    inside = not inside
    return 1 if inside else -1


def interior_point(P):
    """a point strictly inside the convex polygon P: the float64 vertex mean, exactly as the CUDA path computes
    it (the mean of a non-degenerate convex polygon's vertices is interior)"""
    P = strip_ring(P)
    n = len(P)
    return (sum(p[0] for p in P) / n, sum(p[1] for p in P) / n)


def region_contains_convex(rings, P):
    """``region.contains(P)`` for a polygonal region given by its boundary rings (exterior rings and holes of every
    part, any orientation) and a convex polygon P with non-empty interior.
    P is inside the closed region  <=>  no boundary segment meets int(P)  and  one interior point of P is inside."""
    for ring in rings:
        ring = strip_ring(ring)
        n = len(ring)
        for i in range(n):
            if segment_meets_interior(ring[i], ring[(i + 1) % n], P):
                return False
    return point_in_rings(interior_point(P), rings) > 0


def simple_polygon_intersects_convex(Q, P):
    """closed simple polygon Q (possibly non-convex) and closed convex polygon P share a point"""
    Q = strip_ring(Q)
    P = ccw(P)
    n = len(Q)
    for i in range(n):
        if segment_intersects_convex_closed(Q[i], Q[(i + 1) % n], P):
            return True
    # no boundary contact: P inside Q or disjoint
    return point_in_rings(P[0], [Q]) >= 0


def segment_intersects_convex_closed(s0, s1, P):
    """closed segment meets the closed convex polygon P (strict separation test)"""
    P = ccw(P)
    n = len(P)
    for i in range(n):
        a0, a1 = P[i], P[(i + 1) % n]
        if a0 == a1:
            continue
        if orient(a0, a1, s0) < 0 and orient(a0, a1, s1) < 0:
            return False
    if (s0[0], s0[1]) != (s1[0], s1[1]):
        o = [orient(s0, s1, p) for p in P]
        if all(v > 0 for v in o) or all(v < 0 for v in o):
            return False
        return True
    return True


def simple_polygon_interior_overlaps_convex(Q, P):
    """open interiors of a simple polygon Q and a convex polygon P share a point"""
    Q = strip_ring(Q)
    n = len(Q)
    for i in range(n):
        if segment_meets_interior(Q[i], Q[(i + 1) % n], P):
            return True
    return point_in_rings(interior_point(P), [Q]) > 0


# ----------------------------------------------------------------------------------------------------------------------
# independent formulation: exact clipping area (Sutherland-Hodgman in rationals)
# ----------------------------------------------------------------------------------------------------------------------

def clip_area2(S, P):
    """twice the exact area of S intersected with the convex polygon P (S: simple polygon, any orientation)."""
    S = ccw(S)
    P = ccw(P)
    out = [(F(x), F(y)) for x, y in S]
    n = len(P)
    for i in range(n):
        a = (F(P[i][0]), F(P[i][1]))
        b = (F(P[(i + 1) % n][0]), F(P[(i + 1) % n][1]))
        if a == b:
            continue
        inp, out = out, []
        m = len(inp)
        for k in range(m):
            cur, nxt = inp[k], inp[(k + 1) % m]
            dc = (b[0] - a[0]) * (cur[1] - a[1]) - (b[1] - a[1]) * (cur[0] - a[0])
            dn = (b[0] - a[0]) * (nxt[1] - a[1]) - (b[1] - a[1]) * (nxt[0] - a[0])
            if dc >= 0:
                out.append(cur)
            if (dc > 0 and dn < 0) or (dc < 0 and dn > 0):
                t = dc / (dc - dn)
                out.append((cur[0] + t * (nxt[0] - cur[0]), cur[1] + t * (nxt[1] - cur[1])))
        if not out:
            return Fraction(0)
    s = Fraction(0)
    m = len(out)
    for k in range(m):
        s += out[k][0] * out[(k + 1) % m][1] - out[(k + 1) % m][0] * out[k][1]
    return abs(s)


# ----------------------------------------------------------------------------------------------------------------------
# float64 vertex producers (bit-for-bit restatement of what the reference feeds GEOS)
# ----------------------------------------------------------------------------------------------------------------------

def affine_transform(pts, cosphi, sinphi, x, y):
    """shapely 2.0 ``affine_transform(geom, [cos, -sin, sin, cos, x, y])``: per point, Python float arithmetic,
    ``xp = a*x + b*y + xoff``; ``yp = d*x + e*y + yoff`` (reference src/lowLevelPlannerManeuverAutomaton.py:121)"""
    a, b, d, e = cosphi, -sinphi, sinphi, cosphi
    return [(a * px + b * py + x, d * px + e * py + y) for px, py in pts]


def rectangle_vertices(length, width, cx, cy, cosphi, sinphi):
    """commonroad ``Rectangle(length, width, center, orientation).shapely_object`` vertices:
    [(-l/2,-w/2), (-l/2,w/2), (l/2,w/2), (l/2,-w/2)] rotated then translated (homogeneous 3x3 product, i.e. the same
    ``c*x + (-s)*y + tx`` sums); used by reference auxiliary/collisionChecker.py:16-18."""
    l2, w2 = 0.5 * length, 0.5 * width
    return affine_transform([(-l2, -w2), (-l2, w2), (l2, w2), (l2, -w2)], cosphi, sinphi, cx, cy)
