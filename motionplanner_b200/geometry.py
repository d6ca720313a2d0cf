"""Minimal polygon containers with the attribute surface of the shapely objects the hot path touches.

The reference passes shapely geometries around (``o['space']``, ``space_all[t]``, ``goal['space']``); shapely is not a
dependency of this package.  These classes carry coordinates only -- no boolean operations -- and expose the same
attribute names (``exterior.coords``, ``interiors``, ``geoms``, ``is_empty``), so real shapely objects and these are
interchangeable wherever this package *reads* a geometry (``rings_of``).  All containment / intersection decisions on
the hot path are made on the GPU; the two host-side predicates kept here (point in polygon for the goal test,
reference src/lowLevelPlannerManeuverAutomaton.py:137-140; point to ring distance for the stand-alone cost,
src/maneuverAutomatonPlannerStandalone.py:304-308) operate on single points.
"""
from fractions import Fraction

import numpy as np


class _Ring:
    """closed coordinate sequence; built from a list of points or an [n,2] array (coords are materialised lazily)"""

    def __init__(self, coords):
        self._arr = None
        self._coords = None
        if isinstance(coords, np.ndarray):
            a = np.asarray(coords, dtype=np.float64).reshape(-1, 2)
            if len(a) and (a[0, 0] != a[-1, 0] or a[0, 1] != a[-1, 1]):
                a = np.concatenate([a, a[:1]])
            self._arr = a
        else:
            c = [(float(p[0]), float(p[1])) for p in coords]
            if c and c[0] != c[-1]:
                c.append(c[0])
            self._coords = c

    @property
    def coords(self):
        if self._coords is None:
            self._coords = [(float(x), float(y)) for x, y in self._arr]
        return self._coords

    def __array__(self, dtype=None, copy=None):
        a = self._arr if self._arr is not None else np.asarray(self._coords, dtype=np.float64).reshape(-1, 2)
        return a if dtype is None else a.astype(dtype)

    def distance(self, point):
        """euclidean distance from a Point to this ring (shapely LinearRing.distance)"""
        p = np.array([point.x, point.y])
        a = np.asarray(self.coords[:-1])
        b = np.asarray(self.coords[1:])
        d = b - a
        l2 = np.maximum((d * d).sum(1), 1e-300)
        t = np.clip(((p - a) * d).sum(1) / l2, 0.0, 1.0)
        q = a + t[:, None] * d
        return float(np.sqrt(((q - p) ** 2).sum(1)).min())


class Point:
    geom_type = 'Point'

    def __init__(self, x, y=None):
        if y is None:
            x, y = x
        self.x, self.y = float(x), float(y)

    @property
    def coords(self):
        return [(self.x, self.y)]


def _orient(a, b, c):
    d = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    e = 4e-16 * (abs((b[0] - a[0]) * (c[1] - a[1])) + abs((b[1] - a[1]) * (c[0] - a[0])))
    if d > e:
        return 1
    if d < -e:
        return -1
    F = Fraction
    d = (F(b[0]) - F(a[0])) * (F(c[1]) - F(a[1])) - (F(b[1]) - F(a[1])) * (F(c[0]) - F(a[0]))
    return (d > 0) - (d < 0)


def point_in_rings(p, rings):
    """+1 strictly inside, 0 on the boundary, -1 outside (even-odd over all rings, exact orientation signs)"""
    px, py = p
    inside = False
    for ring in rings:
        for a, b in zip(ring[:-1], ring[1:]):
            if a == b:
                continue
            o = _orient(a, b, p)
            if o == 0 and min(a[0], b[0]) <= px <= max(a[0], b[0]) and min(a[1], b[1]) <= py <= max(a[1], b[1]):
                return 0
            if (a[1] > py) != (b[1] > py):
                if (b[1] > a[1] and o > 0) or (b[1] < a[1] and o < 0):
                    inside = not inside
    return 1 if inside else -1


class Polygon:
    geom_type = 'Polygon'

    def __init__(self, shell=None, holes=None):
        if shell is None:
            shell = []
        self.exterior = _Ring(shell if isinstance(shell, np.ndarray) else list(shell))
        self.interiors = [_Ring(h) for h in (holes or [])]

    @property
    def is_empty(self):
        return len(np.asarray(self.exterior)) < 4

    def contains(self, other):
        """only Point arguments are decided on the host (goal test); polygons go through the GPU checker"""
        if isinstance(other, Point) or getattr(other, 'geom_type', None) == 'Point':
            return point_in_rings((other.x, other.y), rings_of(self)) > 0
        raise TypeError("host-side contains() is implemented for points only; use motionplanner_b200.checker")

    @property
    def area(self):
        a = _ring_area(self.exterior.coords)
        return abs(a) - sum(abs(_ring_area(h.coords)) for h in self.interiors)


class MultiPolygon:
    geom_type = 'MultiPolygon'

    def __init__(self, polygons=()):
        self.geoms = list(polygons)

    @property
    def is_empty(self):
        return all(g.is_empty for g in self.geoms)

    def contains(self, other):
        return any(g.contains(other) for g in self.geoms)


def _ring_area(c):
    a = np.asarray(c)
    return 0.5 * float(np.sum(a[:-1, 0] * a[1:, 1] - a[1:, 0] * a[:-1, 1])) if len(a) > 3 else 0.0


def rings_of(geom):
    """all boundary rings (closed coordinate lists) of a shapely-like geometry: Polygon (with holes), MultiPolygon or
    GeometryCollection; non-areal members and empty geometries contribute nothing (reference quirk Q8)"""
    if geom is None or getattr(geom, 'is_empty', False):
        return []
    if hasattr(geom, 'geoms'):
        out = []
        for g in geom.geoms:
            out.extend(rings_of(g))
        return out
    if hasattr(geom, 'exterior'):
        rings = [[(float(p[0]), float(p[1])) for p in geom.exterior.coords]]
        for h in geom.interiors:
            rings.append([(float(p[0]), float(p[1])) for p in h.coords])
        return [r for r in rings if len(r) >= 4]
    return []


def ring_segments(rings):
    """[n,4] float64 array (ax, ay, bx, by) of all ring edges"""
    segs = []
    for r in rings:
        a = np.asarray(r, dtype=np.float64)
        if len(a) >= 2:
            segs.append(np.concatenate([a[:-1], a[1:]], axis=1))
    return np.concatenate(segs) if segs else np.zeros((0, 4))


def polygon_vertices(geom):
    """open vertex list (no closing duplicate) of a simple polygon given as shapely-like object or array"""
    if hasattr(geom, 'exterior'):
        c = [(float(p[0]), float(p[1])) for p in geom.exterior.coords]
    else:
        c = [(float(p[0]), float(p[1])) for p in np.asarray(geom)]
    out = []
    for p in c:
        if not out or out[-1] != p:
            out.append(p)
    if len(out) > 1 and out[0] == out[-1]:
        out.pop()
    return out
