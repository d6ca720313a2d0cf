"""Third-party cross-check of the oracle on NON-degenerate inputs: OpenCV's rotated-rectangle intersection and
point-in-polygon test are independent implementations (float32 inside), so they are only consulted
where the configuration is clear of the decision boundary by > 1e-3 m.  shapely/GEOS themselves are not installable
here; this is the closest independently developed geometry library in the image.
(cv2.intersectConvexConvex was tried too and dropped: it returns area 0 for some deeply overlapping pairs.)"""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from motionplanner_b200 import workloads as W
from oracle import exact as ex
from oracle import oracle as orc


def test_rotated_rectangle_intersection_matches_opencv():
    rng = np.random.default_rng(7)
    n = 0
    for _ in range(2000):
        r1 = ((rng.uniform(-4, 4), rng.uniform(-4, 4)), (rng.uniform(1, 6), rng.uniform(0.5, 2.5)), rng.uniform(-180, 180))
        r2 = ((rng.uniform(-4, 4), rng.uniform(-4, 4)), (rng.uniform(1, 6), rng.uniform(0.5, 2.5)), rng.uniform(-180, 180))
        A, B = cv2.boxPoints(r1).astype(np.float64), cv2.boxPoints(r2).astype(np.float64)
        if abs(W._sat_sep(A, B)) < 1e-3:
            continue
        kind, _ = cv2.rotatedRectangleIntersection(r1, r2)          # 0 = none, 1 = partial, 2 = full
        assert ex.convex_interiors_overlap([tuple(p) for p in A], [tuple(p) for p in B]) == (kind != 0)
        n += 1
    assert n > 1800


def test_point_in_polygon_matches_opencv():
    rng = np.random.default_rng(11)
    n = 0
    for _ in range(300):
        m = int(rng.integers(5, 30))
        ang = np.sort(rng.uniform(0, 2 * math.pi, m))
        r = rng.uniform(2.0, 6.0, m)
        ring = np.stack([r * np.cos(ang), r * np.sin(ang)], -1)       # star-shaped, possibly non-convex
        segs = np.concatenate([ring, np.roll(ring, -1, axis=0)], axis=1)
        contour = ring.astype(np.float32).reshape(-1, 1, 2)
        for _ in range(20):
            p = rng.uniform(-7, 7, 2)
            d = cv2.pointPolygonTest(contour, (float(p[0]), float(p[1])), True)
            if abs(d) < 1e-3:
                continue
            assert (orc.point_in_soup(p, segs) == 1) == (d > 0)
            assert (ex.point_in_rings(tuple(p), [[tuple(v) for v in ring]]) > 0) == (d > 0)
            n += 1
    assert n > 5000
