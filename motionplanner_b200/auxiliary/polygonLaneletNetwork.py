"""Road region of a lanelet network as a boundary *segment soup*.

Reference: auxiliary/polygonLaneletNetwork.py:5-129 unions every lanelet polygon plus "bridge" polygons spanning
left/right neighbours and successor pairs (they close the slivers between neighbouring lanelets) with shapely.
The GPU checker never needs the union polygon itself, only its boundary (containment = no boundary segment meets the
occupancy's interior + one interior point inside, see csrc/mpc_kernels.cu), so this module

  1. enumerates exactly the polygons the reference unions (same breadth-first walk, same ``safe_only`` rules), and
  2. extracts the boundary of their union as segments: every edge is split where another polygon's edge crosses or
     touches it, and a piece is kept iff exactly one of its two sides is covered by some polygon.

Host preprocessing, float64, once per scenario.  Coordinates closer than 1e-6 m (the precision of CommonRoad files) are treated as the same vertex (GEOS
snaps at a comparable scale when it nodes the union); this is input preparation, not part of the collision decision.
"""
import numpy as np

from ..geometry import point_in_rings

TOL = 1e-6


class RoadRegion:
    """shapely-like handle of the road region: boundary segments + membership of a point"""
    geom_type = 'RoadRegion'

    def __init__(self, segments, polygons):
        self.segments = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
        self.polygons = polygons

    @property
    def is_empty(self):
        return len(self.segments) == 0

    def covers_point(self, p):
        return any(point_in_rings((float(p[0]), float(p[1])), [_closed(v)]) >= 0 for v in self.polygons)


def _closed(v):
    r = [tuple(map(float, q)) for q in v]
    if r[0] != r[-1]:
        r.append(r[0])
    return r


# ---------------------------------------------------------------------------------------------------------------------
# 1. the polygons the reference unions
# ---------------------------------------------------------------------------------------------------------------------

def lanelet2polygon(lanelet):
    """left bound followed by the reversed right bound (reference :71-87)"""
    return np.concatenate([lanelet.left_vertices, lanelet.right_vertices[::-1]])


def union_left(lanelet, lanelet_left):
    """polygon spanning a lanelet and its left neighbour (reference :89-108)"""
    left, right = lanelet_left.left_vertices, lanelet.right_vertices
    if lanelet.adj_left_same_direction:
        return np.concatenate([left, right[::-1]])
    return np.concatenate([left, right])


def union_successor(lanelet, lanelet_suc):
    """polygon spanning a lanelet and its successor (reference :110-129)"""
    return np.concatenate([lanelet.left_vertices, lanelet_suc.left_vertices, lanelet_suc.right_vertices[::-1],
                           lanelet.right_vertices[::-1]])


def road_polygons(scenario, safe_only=False, x0=None):
    """vertex arrays of all polygons whose union is the road (reference walk :29-67)"""
    net = scenario.lanelet_network
    lanelets = {l.lanelet_id: l for l in net.lanelets}
    if x0 is None:
        queue = [net.lanelets[0].lanelet_id]
    else:
        queue = list(net.find_lanelet_by_position([np.asarray(x0[0:2])])[0])
    visited, polys = [], []
    if not queue:
        return polys
    while queue:
        l = lanelets[queue.pop(0)]
        polys.append(lanelet2polygon(l))
        if l.adj_left is not None and l.adj_left not in visited and (not safe_only or l.adj_left_same_direction):
            polys.append(union_left(l, lanelets[l.adj_left]))
            queue.append(l.adj_left)
        if l.adj_right is not None and l.adj_right not in visited and (not safe_only or l.adj_right_same_direction):
            polys.append(union_left(lanelets[l.adj_right], l))
            queue.append(l.adj_right)
        for s in l.successor:
            if s not in visited:
                polys.append(union_successor(l, lanelets[s]))
                queue.append(s)
        for p in l.predecessor:
            if p not in visited and not safe_only:
                polys.append(union_successor(lanelets[p], l))
                queue.append(p)
        visited.append(l.lanelet_id)
        if not queue and not safe_only:
            queue.extend(i for i in lanelets if i not in visited)
    return polys


# ---------------------------------------------------------------------------------------------------------------------
# 2. boundary of the union
# ---------------------------------------------------------------------------------------------------------------------

def _cluster(points, tol=TOL):
    """replace every point by the representative (lowest index) of its tol-cluster (single linkage)"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    uniq, first, inv = np.unique(points, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    n = len(uniq)
    pairs = cKDTree(uniq).query_pairs(tol, output_type='ndarray')
    if len(pairs) == 0:
        return points[first][inv]
    g = coo_matrix((np.ones(len(pairs), np.int8), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    # representative of a cluster = its member that appears first in the input
    rep = np.full(lab.max() + 1, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(rep, lab, first)
    return points[rep[lab]][inv]


def _dedupe_ring(v):
    keep = np.ones(len(v), bool)
    keep[1:] = np.any(v[1:] != v[:-1], axis=1)
    v = v[keep]
    if len(v) > 1 and np.all(v[0] == v[-1]):
        v = v[:-1]
    return v


def _winding(points, poly):
    """winding number of every point w.r.t. the closed polygon (non-zero = covered)"""
    a, b = poly, np.roll(poly, -1, axis=0)
    px, py = points[:, 0][:, None], points[:, 1][:, None]
    ax, ay, bx, by = a[:, 0][None], a[:, 1][None], b[:, 0][None], b[:, 1][None]
    cross = (bx - ax) * (py - ay) - (by - ay) * (px - ax)
    up = (ay <= py) & (by > py) & (cross > 0)
    down = (ay > py) & (by <= py) & (cross < 0)
    return up.sum(1) - down.sum(1)


def union_boundary(polygons):
    """[n,4] boundary segments of the union of the given polygons.

    All work happens on a snapped vertex set: input vertices closer than TOL are merged, an edge that passes within
    TOL of another polygon's vertex is bent through that vertex, proper crossings add a vertex to both edges.  The
    refined rings then only meet at shared vertices, so a piece of edge is either entirely boundary or not, and which
    one is decided by the coverage (non-zero winding w.r.t. any refined ring) just left and just right of its
    midpoint."""
    polys = [np.asarray(p, dtype=np.float64) for p in polygons if len(p) >= 3]
    if not polys:
        return np.zeros((0, 4))
    lens = [len(p) for p in polys]
    allv = _cluster(np.concatenate(polys))
    polys, o = [], 0
    for n in lens:
        r = _dedupe_ring(allv[o:o + n])
        o += n
        if len(r) >= 3:
            polys.append(r)
    if not polys:
        return np.zeros((0, 4))
    A = np.concatenate(polys)
    B = np.concatenate([np.roll(p, -1, axis=0) for p in polys])
    E = len(A)
    D = B - A
    L = np.sqrt((D * D).sum(1))
    xmin, xmax = np.minimum(A[:, 0], B[:, 0]), np.maximum(A[:, 0], B[:, 0])
    ymin, ymax = np.minimum(A[:, 1], B[:, 1]), np.maximum(A[:, 1], B[:, 1])
    order = np.argsort(xmin, kind='stable')
    xs = xmin[order]
    hi = np.searchsorted(xs, xmax[order] + TOL, side='right')
    I, J = [], []
    for k in range(E):
        if hi[k] > k + 1:
            i = order[k]
            c = order[k + 1:hi[k]]
            c = c[(ymin[c] <= ymax[i] + TOL) & (ymax[c] >= ymin[i] - TOL)]
            if len(c):
                I.append(np.full(len(c), i))
                J.append(c)
    split_e, split_t, split_p = [], [], []
    if I:
        I, J = np.concatenate(I), np.concatenate(J)

        def sdist(i, P):                     # signed distance of P to the line of edge i (left positive)
            return (D[i, 0] * (P[:, 1] - A[i, 1]) - D[i, 1] * (P[:, 0] - A[i, 0])) / L[i]

        d1, d2 = sdist(I, A[J]), sdist(I, B[J])
        d3, d4 = sdist(J, A[I]), sdist(J, B[I])
        proper = (d1 * d2 < 0) & (d3 * d4 < 0) & (np.minimum(np.abs(d1), np.abs(d2)) > TOL) & \
                 (np.minimum(np.abs(d3), np.abs(d4)) > TOL)
        k = np.nonzero(proper)[0]
        if len(k):
            ti = d3[k] / (d3[k] - d4[k])
            tj = d1[k] / (d1[k] - d2[k])
            pts = A[I[k]] + ti[:, None] * D[I[k]]
            split_e += [I[k], J[k]]
            split_t += [ti, tj]
            split_p += [pts, pts]
        for edge, P, d in ((I, A[J], d1), (I, B[J], d2), (J, A[I], d3), (J, B[I], d4)):
            t = ((P[:, 0] - A[edge, 0]) * D[edge, 0] + (P[:, 1] - A[edge, 1]) * D[edge, 1]) / (L[edge] ** 2)
            m = TOL / L[edge]
            k = np.nonzero((np.abs(d) <= TOL) & (t > m) & (t < 1 - m))[0]
            if len(k):
                split_e.append(edge[k])
                split_t.append(t[k])
                split_p.append(P[k])
    # refined rings: original vertices plus the split vertices, everything snapped together
    if split_e:
        se, st_, sp = np.concatenate(split_e), np.concatenate(split_t), np.concatenate(split_p)
        snapped = _cluster(np.concatenate([A, sp]))
        A2, sp = snapped[:E], snapped[E:]
        idx = np.lexsort((st_, se))
        se, sp = se[idx], sp[idx]
        starts = np.searchsorted(se, np.arange(E), side='left')
        ends = np.searchsorted(se, np.arange(E), side='right')
    else:
        A2, sp, starts, ends = A, np.zeros((0, 2)), np.zeros(E, int), np.zeros(E, int)
    rings, e0 = [], 0
    for p in polys:
        pts = []
        for e in range(e0, e0 + len(p)):
            pts.append(A2[e:e + 1])
            if ends[e] > starts[e]:
                pts.append(sp[starts[e]:ends[e]])
        e0 += len(p)
        r = _dedupe_ring(np.concatenate(pts))
        if len(r) >= 3:
            rings.append(r)
    sa = np.concatenate(rings)
    sb = np.concatenate([np.roll(r, -1, axis=0) for r in rings])
    keep = np.any(sa != sb, axis=1)
    sa, sb = sa[keep], sb[keep]
    flip = (sa[:, 0] > sb[:, 0]) | ((sa[:, 0] == sb[:, 0]) & (sa[:, 1] > sb[:, 1]))
    segs = np.unique(np.concatenate([np.where(flip[:, None], sb, sa), np.where(flip[:, None], sa, sb)], axis=1), axis=0)
    # coverage just left and just right of every piece
    mid = 0.5 * (segs[:, 0:2] + segs[:, 2:4])
    d = segs[:, 2:4] - segs[:, 0:2]
    n = np.stack([-d[:, 1], d[:, 0]], 1) / np.sqrt((d * d).sum(1))[:, None]
    eps = 1e-8
    left, right = mid + eps * n, mid - eps * n
    cov_l = np.zeros(len(segs), bool)
    cov_r = np.zeros(len(segs), bool)
    for p in rings:
        x0, y0, x1, y1 = p[:, 0].min() - 1e-3, p[:, 1].min() - 1e-3, p[:, 0].max() + 1e-3, p[:, 1].max() + 1e-3
        sel = np.nonzero((mid[:, 0] >= x0) & (mid[:, 0] <= x1) & (mid[:, 1] >= y0) & (mid[:, 1] <= y1))[0]
        if len(sel):
            cov_l[sel] |= _winding(left[sel], p) != 0
            cov_r[sel] |= _winding(right[sel], p) != 0
    return segs[cov_l != cov_r]


def polygonLaneletNetwork(scenario, safe_only=False, x0=None):
    """road region of the lanelet network (drop-in for reference auxiliary/polygonLaneletNetwork.py:5); the result
    exposes ``.segments`` (what the GPU checker stages) instead of a shapely polygon"""
    polys = road_polygons(scenario, safe_only, x0)
    return RoadRegion(union_boundary(polys), polys)
