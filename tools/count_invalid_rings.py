#!/usr/bin/env python
"""How exposed is the road region to the one step of the reference that is NOT restated -- `buffer(0)`?

The reference repairs every lanelet / bridge polygon that is not a valid (simple) ring with shapely's `buffer(0)`
before the union (auxiliary/polygonLaneletNetwork.py:83-85, 104-106, 125-127).  This package unions the rings as they
are, with the non-zero winding rule (motionplanner_b200/auxiliary/polygonLaneletNetwork.py::union_boundary).  The two
agree on every simple ring; on a self-intersecting ring `buffer(0)` keeps the lobes that wind with the ring's overall
orientation and drops the others, the non-zero rule keeps both.

This script walks all bundled scenarios (both walks the reference does: validator = all lanelets, planner =
safe_only from the initial position), tests every ring for self-intersections with exact orientation signs, and for
each invalid ring measures the area of its minority-orientation lobes (what `buffer(0)` would drop) and how much of
that area is NOT covered by any other polygon of the same union -- only that part can change the road region.

    python tools/count_invalid_rings.py [scenario dir] > profiles/r02_invalid_rings.json
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from motionplanner_b200.auxiliary.polygonLaneletNetwork import _dedupe_ring, _winding, road_polygons  # noqa: E402
from motionplanner_b200.crlite import CommonRoadFileReader  # noqa: E402
from motionplanner_b200.geometry import _orient  # noqa: E402


def self_intersections(v):
    """pairs (i, j) of non-adjacent edges of the closed ring v that share a point (exact signs on float candidates)"""
    n = len(v)
    a, b = v, np.roll(v, -1, axis=0)
    d = b - a
    out = []
    # candidate pairs by bounding boxes, then float orientation with a generous band, then exact
    xmin, xmax = np.minimum(a[:, 0], b[:, 0]), np.maximum(a[:, 0], b[:, 0])
    ymin, ymax = np.minimum(a[:, 1], b[:, 1]), np.maximum(a[:, 1], b[:, 1])
    for i in range(n):
        j = np.arange(i + 2, n if i > 0 else n - 1)                  # skip the two neighbours
        if not len(j):
            continue
        j = j[(xmin[j] <= xmax[i]) & (xmax[j] >= xmin[i]) & (ymin[j] <= ymax[i]) & (ymax[j] >= ymin[i])]
        for k in j:
            o1, o2 = _orient(tuple(a[i]), tuple(b[i]), tuple(a[k])), _orient(tuple(a[i]), tuple(b[i]), tuple(b[k]))
            o3, o4 = _orient(tuple(a[k]), tuple(b[k]), tuple(a[i])), _orient(tuple(a[k]), tuple(b[k]), tuple(b[i]))
            if o1 * o2 < 0 and o3 * o4 < 0:
                out.append((i, int(k), 'cross'))
            elif 0 in (o1, o2, o3, o4):
                def on(p, s0, s1):
                    return _orient(s0, s1, p) == 0 and min(s0[0], s1[0]) <= p[0] <= max(s0[0], s1[0]) and \
                        min(s0[1], s1[1]) <= p[1] <= max(s0[1], s1[1])
                if on(tuple(a[k]), tuple(a[i]), tuple(b[i])) or on(tuple(b[k]), tuple(a[i]), tuple(b[i])) or \
                        on(tuple(a[i]), tuple(a[k]), tuple(b[k])) or on(tuple(b[i]), tuple(a[k]), tuple(b[k])):
                    out.append((i, int(k), 'touch'))
    return out


def lobe_areas(v, others, grid=600):
    """area with positive / negative winding of ring v, and the part of the minority area no other polygon covers"""
    x0, y0, x1, y1 = v[:, 0].min(), v[:, 1].min(), v[:, 0].max(), v[:, 1].max()
    nx = grid
    ny = max(8, int(grid * (y1 - y0) / max(x1 - x0, 1e-9)))
    ny = min(ny, 4 * grid)
    xs = x0 + (np.arange(nx) + 0.5) * (x1 - x0) / nx
    ys = y0 + (np.arange(ny) + 0.5) * (y1 - y0) / ny
    pts = np.stack(np.meshgrid(xs, ys, indexing='ij'), -1).reshape(-1, 2)
    cell = (x1 - x0) / nx * (y1 - y0) / ny
    w = np.concatenate([_winding(pts[k:k + 20000], v) for k in range(0, len(pts), 20000)])
    pos, neg = float((w > 0).sum() * cell), float((w < 0).sum() * cell)
    minority = (w < 0) if pos >= neg else (w > 0)
    unc = minority.copy()
    if minority.any():
        sel = np.nonzero(minority)[0]
        cov = np.zeros(len(sel), bool)
        for o in others:
            cov |= _winding(pts[sel], o) != 0
        unc[sel] = ~cov
    return pos, neg, float(minority.sum() * cell), float(unc.sum() * cell), float(cell)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/scenarios'
    rep = dict(scenarios=0, rings=0, invalid_rings=0, scenarios_with_invalid=0, crossings=0, touches_only_rings=0,
               minority_lobe_area_m2=0.0, minority_lobe_area_not_covered_elsewhere_m2=0.0, details=[])
    for f in sorted(glob.glob(os.path.join(src, '*.xml'))):
        name = os.path.basename(f)[:-4]
        scenario, pps = CommonRoadFileReader(f).open()
        pp = list(pps.planning_problem_dict.values())[0]
        s0 = pp.initial_state
        x0 = np.array([s0.position[0], s0.position[1], s0.velocity, s0.orientation])
        rep['scenarios'] += 1
        bad_here = 0
        for walk, kw in (('validator', dict()), ('planner', dict(safe_only=True, x0=x0))):
            try:
                polys = [_dedupe_ring(np.asarray(p, float)) for p in road_polygons(scenario, **kw)]
            except Exception:          # noqa: BLE001
                continue
            for k, v in enumerate(polys):
                if len(v) < 3:
                    continue
                rep['rings'] += 1
                hits = self_intersections(v)
                if not hits:
                    continue
                cross = [h for h in hits if h[2] == 'cross']
                rep['invalid_rings'] += 1
                bad_here += 1
                rep['crossings'] += len(cross)
                d = dict(scenario=name, walk=walk, ring=k, vertices=len(v), crossings=len(cross), touches=len(hits) - len(cross))
                if cross:
                    pos, neg, minority, unc, cell = lobe_areas(v, [o for i, o in enumerate(polys) if i != k and len(o) >= 3])
                    d.update(area_pos_m2=pos, area_neg_m2=neg, minority_m2=minority, minority_not_covered_elsewhere_m2=unc, grid_cell_m2=cell)
                    rep['minority_lobe_area_m2'] += minority
                    rep['minority_lobe_area_not_covered_elsewhere_m2'] += unc
                else:
                    rep['touches_only_rings'] += 1
                rep['details'].append(d)
        rep['scenarios_with_invalid'] += int(bad_here > 0)
        print('%-34s invalid rings %d' % (name, bad_here), file=sys.stderr)
    print(json.dumps(rep, indent=1))


if __name__ == '__main__':
    main()
