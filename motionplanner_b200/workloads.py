"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): plain numpy arrays in the layout of include/mpc.h.

Used by bench.py and the parity tests (both the CUDA path and the CPU oracle are fed these same arrays).
Vehicle constants: reference vehicle/vehicleParameter.py:3-15; primitive grids: reference
maneuverAutomaton/createManeuverAutomaton.py:15,18 (re-used as the value ranges of the synthetic library).
"""
import numpy as np

from ._cabi import QUERY_DTYPE

LENGTH, WIDTH, WHEELBASE = 4.3, 1.7, 2.3
B_REAR = WHEELBASE / 2
ACCELERATIONS = np.array([-8, -4, -2, -1.2, -0.8, -0.4, 0, 0.4, 0.8, 1.2, 2, 4, 8], dtype=np.float64)
ORIENTATIONS = np.array([-1, -0.8, -0.6, -0.4, -0.3, -0.2, -0.1, 0, 0.1, 0.2, 0.3, 0.4, 0.6, 0.8, 1], dtype=np.float64)

# car outline in the rear-axle frame, same vertex order as the reference (clockwise)
CAR = np.array([[-(LENGTH / 2 - B_REAR), -WIDTH / 2], [-(LENGTH / 2 - B_REAR), WIDTH / 2],
                [LENGTH / 2 + B_REAR, WIDTH / 2], [LENGTH / 2 + B_REAR, -WIDTH / 2]])


def _rk4_single_track(v0, acc, steer, steps, dt):
    """vectorised RK4 of the kinematic single-track model (reference createManeuverAutomaton.py:67-70)
    returns states [n, steps+1, 4] (x, y, v, phi) starting at the origin"""
    n = len(v0)
    x = np.zeros((n, steps + 1, 4))
    s = np.stack([np.zeros(n), np.zeros(n), v0, np.zeros(n)], axis=1)
    x[:, 0] = s
    tan_s = np.tan(steer) / WHEELBASE

    def f(s):
        return np.stack([s[:, 2] * np.cos(s[:, 3]), s[:, 2] * np.sin(s[:, 3]), acc, s[:, 2] * tan_s], axis=1)

    for i in range(steps):
        k1 = f(s)
        k2 = f(s + 0.5 * dt * k1)
        k3 = f(s + 0.5 * dt * k2)
        k4 = f(s + dt * k3)
        s = s + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        x[:, i + 1] = s
    return x


def rect_vertices(cx, cy, phi, length, width):
    """[n,4,2] rectangle corners, commonroad Rectangle vertex order"""
    l2, w2 = 0.5 * np.asarray(length), 0.5 * np.asarray(width)
    loc = np.stack([np.stack([-l2, -w2], -1), np.stack([-l2, w2], -1), np.stack([l2, w2], -1), np.stack([l2, -w2], -1)], -2)
    c, s = np.cos(phi)[..., None], np.sin(phi)[..., None]
    x = c * loc[..., 0] + (-s) * loc[..., 1] + np.asarray(cx)[..., None]
    y = s * loc[..., 0] + c * loc[..., 1] + np.asarray(cy)[..., None]
    return np.stack([x, y], -1)


def synthetic_library(P=4096, T=50, dt=0.1, seed=0, v_range=(0.0, 30.0), orientation_scale=1.0):
    """C4 library: P arcs of T occupancies (time offsets 0..T-1), car rectangle at every step.  `v_range` and
    `orientation_scale` narrow the start speeds / final headings (the mixed regime: primitives a planner would consider
    at the ego's speed instead of the whole speed range)."""
    rng = np.random.default_rng(seed)
    v0 = rng.uniform(v_range[0], v_range[1], P)
    acc = rng.choice(ACCELERATIONS, P)
    tf = T * dt
    acc = np.where(v0 + acc * tf < 0.0, -v0 / tf, acc)
    acc = np.where(v0 + acc * tf > 30.0, (30.0 - v0) / tf, acc)
    o = rng.choice(ORIENTATIONS, P) * orientation_scale
    dist = v0 * tf + 0.5 * acc * tf ** 2
    steer = np.where((o == 0) | (dist <= 1e-9), 0.0, np.arctan(WHEELBASE * o / np.maximum(dist, 1e-9)))
    steer = np.clip(steer, -1.0, 1.0)
    x = _rk4_single_track(v0, acc, steer, T - 1, dt)                      # [P, T, 4]
    c, s = np.cos(x[..., 3])[..., None], np.sin(x[..., 3])[..., None]
    vx = c * CAR[:, 0] + (-s) * CAR[:, 1] + x[..., 0][..., None]
    vy = s * CAR[:, 0] + c * CAR[:, 1] + x[..., 1][..., None]
    verts = np.stack([vx, vy], -1)                                        # [P, T, 4, 2]
    nv = np.full((P, T), 4, np.int32)
    tidx = np.tile(np.arange(T, dtype=np.int32), (P, 1))
    return dict(verts=verts, nv=nv, tidx=tidx, set_off=np.array([0, P], np.int32),
                set_idx=np.arange(P, dtype=np.int32), x=x)


def dense_traffic_scene(T=50, lanes=8, slots=32, dt=0.1, seed=1, lane_width=3.5, headway=(8.0, 20.0),
                        lane_speed=(5.0, 25.0)):
    """C4 scene: lanes x slots rectangles, constant velocity, straight multi-lane corridor as the region.
    Deviations from SURVEY.md 8(d), stated in bench.py's config: one cruise speed per LANE (U(lane_speed)) with +-1 m/s
    per car instead of U(5,25) per car, and heading sigma 0.01 instead of 0.03 -- with per-car speeds the cars of a lane
    drive through each other within the 5 s horizon."""
    rng = np.random.default_rng(seed)
    O = lanes * slots
    length = rng.uniform(3.5, 5.5, O)
    width = rng.uniform(1.6, 2.2, O)
    head = rng.uniform(headway[0], headway[1], (lanes, slots))
    x0 = np.cumsum(head + 5.0, axis=1).reshape(-1)
    y0 = (np.repeat(np.arange(lanes), slots) + 0.5) * lane_width + rng.normal(0.0, 0.15, O)
    # one cruise speed per lane with +-1 m/s per car, so the traffic stays a traffic for the horizon
    lane_v = rng.uniform(lane_speed[0], lane_speed[1], lanes)
    v = np.repeat(lane_v, slots) + rng.uniform(-1.0, 1.0, O)
    phi = rng.normal(0.0, 0.01, O)
    t = np.arange(T)[:, None] * dt
    cx = x0[None, :] + v[None, :] * np.cos(phi)[None, :] * t
    cy = y0[None, :] + v[None, :] * np.sin(phi)[None, :] * t
    obst = rect_vertices(cx, cy, np.broadcast_to(phi, (T, O)), length, width).reshape(T * O, 4, 2)
    xmin, xmax = -200.0, float(cx.max() + 200.0)
    ymin, ymax = 0.0, lanes * lane_width
    ring = [(xmin, ymin), (xmax, ymin), (xmax, ymax), (xmin, ymax)]
    segs = np.array([ring[i] + ring[(i + 1) % 4] for i in range(4)], dtype=np.float64)
    return dict(scene_step_off=np.array([0, T], np.int32), step_obst_off=(np.arange(T + 1) * O).astype(np.int32),
                obst=obst, obst_nv=np.full(T * O, 4, np.int32), step_seg_begin=np.zeros(T, np.int32),
                step_seg_end=np.full(T, 4, np.int32), segs=segs, extent=(xmin, xmax, ymin, ymax),
                x0=x0, y0=y0, v=v, length=length, lanes=lanes, slots=slots, lane_width=lane_width, lane_v=lane_v)


def mixed_traffic_problem(P=4096, O_lanes=8, slots=32, T=50, nqueries=1, seed=0, ego_lane=3):
    """the C4 shape (P x lanes*slots x T) in the regime a planner cares about: 30-50 % of the candidates are
    collision-free, so they need every time step (the dense C4 scene decides 99.8 % of its candidates by collision).
    Thinner traffic (headway U(25,60) m, lane speeds U(12,22) m/s, scene seed 101) and a library around the ego lane's
    cruise speed (+-4 m/s, final headings x0.35, library seed = `seed`).  38 % free at the default sizes (oracle)."""
    scene = dense_traffic_scene(T, O_lanes, slots, seed=101, headway=(25.0, 60.0), lane_speed=(12.0, 22.0))
    v = float(scene["lane_v"][ego_lane])
    lib = synthetic_library(P, T, seed=seed, v_range=(max(v - 4.0, 0.0), min(v + 4.0, 30.0)), orientation_scale=0.35)
    queries, origins = dense_traffic_queries(scene, P, nqueries)
    return lib, scene, origins, queries


def dense_traffic_queries(scene, P, n=1, seed=2, lanes=8, lane_width=3.5):
    """ego poses inside the traffic; query 0 is the un-jittered pose (lane 3 centre, mid-traffic)"""
    rng = np.random.default_rng(seed)
    q = np.zeros(n, dtype=QUERY_DTYPE)
    slots = scene["slots"]
    lane = 3
    xs = np.sort(scene["x0"][lane * slots:(lane + 1) * slots])
    k = slots // 2
    xm = 0.5 * (xs[k] + xs[k + 1]) - B_REAR               # rear axle such that the car sits mid-gap in lane 3
    for i in range(n):
        x = xm + (rng.uniform(-2.0, 2.0) if i else 0.0)
        y = (lane + 0.5) * lane_width + (rng.uniform(-0.5, 0.5) if i else 0.0)
        phi = rng.normal(0.0, 0.03) if i else 0.0
        q[i] = (x, y, np.cos(phi), np.sin(phi), 0, 0, 2 ** 30, 0, 0, P, 0, 0)
    origin = np.array([[xm, (lane + 0.5) * lane_width]])
    return q, origin


def c4_problem(P=4096, O_lanes=8, slots=32, T=50, nqueries=1):
    """the BASELINE.json configs[3] workload: (library, scenes, origins, queries)"""
    lib = synthetic_library(P, T)
    scene = dense_traffic_scene(T, O_lanes, slots)
    queries, origins = dense_traffic_queries(scene, P, nqueries)
    return lib, scene, origins, queries


def aroc_shaped_problem(P=4096, T=10, O_lanes=8, slots=32, nqueries=8, lib_verts=12, obst_verts=4, seed=7):
    """synthetic stand-in for an AROC automaton (reference maneuverAutomaton/loadAROCautomaton.py:110-137: occupancy sets
    are convex hulls of up to 12 vertices over time INTERVALS): P primitives x T interval occupancies, each a convex
    `lib_verts`-gon around the swept car; every query is tested against two consecutive scenes, as the reference
    intersects space[i] and space[i+1] for interval automata (src/lowLevelPlannerManeuverAutomaton.py:25-30).
    obst_verts = 4: rectangle obstacles (k_check<16,4>); 8: hexagonal obstacles (k_check<16,8>).  Returns (library, scenes, origins, queries) with 2 * nqueries query records."""
    rng = np.random.default_rng(seed)
    dt = 0.1
    base = synthetic_library(P, T + 1, seed=seed)
    x = base["x"]                                                       # [P, T+1, 4]
    mid = 0.5 * (x[:, :-1, :] + x[:, 1:, :])                            # interval centre pose
    sweep = np.hypot(x[:, 1:, 0] - x[:, :-1, 0], x[:, 1:, 1] - x[:, :-1, 1])
    # one vertex per angular stratum: strictly increasing, no coinciding vertices, no wrap-around
    ang = 2 * np.pi * (np.arange(lib_verts) + rng.uniform(0.1, 0.9, (P, T, lib_verts))) / lib_verts
    a = (0.5 * LENGTH + 0.5 * sweep + 0.3)[..., None]                   # semi-axes: car + travelled distance + uncertainty
    b = (0.5 * WIDTH + 0.3)
    ex, ey = a * np.cos(ang), b * np.sin(ang)
    c, s = np.cos(mid[..., 3])[..., None], np.sin(mid[..., 3])[..., None]
    cx = (mid[..., 0] + B_REAR * np.cos(mid[..., 3]))[..., None]
    cy = (mid[..., 1] + B_REAR * np.sin(mid[..., 3]))[..., None]
    verts = np.stack([c * ex - s * ey + cx, s * ex + c * ey + cy], -1)  # [P, T, V, 2]
    lib = dict(verts=verts, nv=np.full((P, T), lib_verts, np.int32), tidx=np.tile(np.arange(T, dtype=np.int32), (P, 1)),
               set_off=np.array([0, P], np.int32), set_idx=np.arange(P, dtype=np.int32))
    sc = dense_traffic_scene(T + 1, O_lanes, slots)
    O = O_lanes * slots
    obst = sc["obst"].reshape(T + 1, O, 4, 2)
    if obst_verts > 4:        # hexagons: the rectangle with a 0.3 m nose and tail (six strictly convex corners)
        r = obst[:-1]
        u = r[:, :, 3] - r[:, :, 0]
        u = u / np.linalg.norm(u, axis=-1, keepdims=True)
        tail, nose = 0.5 * (r[:, :, 0] + r[:, :, 1]) - 0.3 * u, 0.5 * (r[:, :, 2] + r[:, :, 3]) + 0.3 * u
        ob = np.zeros((T, O, 8, 2))
        ob[:, :, :6] = np.stack([r[:, :, 0], tail, r[:, :, 1], r[:, :, 2], nose, r[:, :, 3]], axis=2)
        nv = np.full(T * O, 6, np.int32)
        ob = ob.reshape(T * O, 8, 2)
    else:
        ob, nv = obst[:-1].reshape(T * O, 4, 2), np.full(T * O, 4, np.int32)
    # two scenes: steps 0..T-1 and the same shifted by one step (space[i] and space[i+1])
    ob2 = (np.concatenate([ob.reshape(T, O, -1, 2)[1:], ob.reshape(T, O, -1, 2)[-1:]])).reshape(T * O, -1, 2)
    scenes = dict(scene_step_off=np.array([0, T, 2 * T], np.int32), step_obst_off=(np.arange(2 * T + 1) * O).astype(np.int32),
                  obst=np.concatenate([ob, ob2]), obst_nv=np.concatenate([nv, nv]),
                  step_seg_begin=np.zeros(2 * T, np.int32), step_seg_end=np.full(2 * T, 4, np.int32), segs=sc["segs"],
                  slots=sc["slots"], x0=sc["x0"])
    q1, origin = dense_traffic_queries(scenes, P, nqueries)
    q = np.concatenate([q1, q1])
    q["scene"][nqueries:] = 1
    return lib, scenes, np.concatenate([origin, origin]), q


GOLDEN_SCENARIOS = None


def load_golden_scenarios(path=None):
    """the 100 bundled scenarios as flat arrays (tests/golden/scenarios.npz, extracted from the reference's XML files by
    tools/make_golden.py): road boundary segments, per-step obstacle pieces, initial state, goal window"""
    import os
    global GOLDEN_SCENARIOS
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenarios.npz")
    if GOLDEN_SCENARIOS is not None and GOLDEN_SCENARIOS[0] == path:
        return GOLDEN_SCENARIOS[1]
    z = np.load(path)
    out = {}
    for name in z["names"]:
        m = z[name + "/meta"]
        out[str(name)] = dict(dt=float(m[0]), t_start=int(m[1]), t_end=int(m[2]), nsteps=int(m[3]),
                              nobstacles=int(m[4]), x0=z[name + "/x0"], goal_ring=z[name + "/goal_ring"],
                              road=z[name + "/road"], road_safe=z[name + "/road_safe"],
                              step_obst_off=z[name + "/step_obst_off"], obst=z[name + "/obst"],
                              obst_nv=z[name + "/obst_nv"])
    GOLDEN_SCENARIOS = (path, out)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# adversarial parity problems: contacts within a few float64 ulps, mixed vertex counts, ragged steps, holes
# ---------------------------------------------------------------------------------------------------------------------

def random_convex(rng, n, radius):
    """n points in convex position: on a randomly oriented ellipse (semi-axes radius, 0.6..1 radius) at sorted angles"""
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    while np.min(np.diff(np.concatenate([ang, [ang[0] + 2 * np.pi]]))) < 0.05:      # no near-duplicate vertices
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    a, b, rot = radius, radius * rng.uniform(0.6, 1.0), rng.uniform(0, np.pi)
    x, y = a * np.cos(ang), b * np.sin(ang)
    return np.stack([np.cos(rot) * x - np.sin(rot) * y, np.sin(rot) * x + np.cos(rot) * y], -1)


def adversarial_problem(seed, P=96, J=5, nsteps=8, max_obst=10, lib_verts=4, obst_verts=4, offset=(0.0, 0.0),
                        nqueries=3, with_region=True):
    """small ragged problem whose obstacles are pushed into (near-)contact with transformed occupancies.

    Returns (library, scenes, origins, queries)."""
    rng = np.random.default_rng(seed)
    Vl = max(lib_verts, 4)
    verts = np.zeros((P, J, Vl, 2))
    nv = np.zeros((P, J), np.int32)
    for p in range(P):
        for j in range(J):
            if rng.random() < 0.1:
                continue
            n = 4 if lib_verts == 4 else int(rng.integers(3, lib_verts + 1))
            if lib_verts == 4 and rng.random() < 0.8:
                phi = rng.uniform(-1, 1)
                poly = rect_vertices(np.array(rng.uniform(0, 50)), np.array(rng.uniform(-20, 20)), np.array(phi),
                                     LENGTH, WIDTH)
                if rng.random() < 0.5:
                    poly = poly[::-1]
            else:
                poly = random_convex(rng, n, rng.uniform(1.0, 2.5)) + [rng.uniform(0, 50), rng.uniform(-20, 20)]
            verts[p, j, :len(poly)] = poly
            nv[p, j] = len(poly)
    tidx = np.tile(np.arange(J, dtype=np.int32), (P, 1))
    lib = dict(verts=verts, nv=nv, tidx=tidx, set_off=np.array([0, P // 2, P], np.int32),
               set_idx=np.arange(P, dtype=np.int32))

    ox, oy = offset
    q = np.zeros(nqueries, dtype=QUERY_DTYPE)
    for i in range(nqueries):
        phi = rng.choice([0.0, np.pi / 2, rng.uniform(-np.pi, np.pi)])
        c, s = (1.0, 0.0) if phi == 0.0 else ((0.0, 1.0) if phi == np.pi / 2 else (np.cos(phi), np.sin(phi)))
        cs = int(rng.integers(0, 2))
        cnt = P // 2
        q[i] = (ox + rng.uniform(-5, 5), oy + rng.uniform(-5, 5), c, s, 0, int(rng.integers(0, 3)),
                int(rng.choice([nsteps + 5, nsteps - 2])), cs, 0, cnt, 0, 0)

    def world(qi, p, j):
        a, b = q[qi]["c"], -q[qi]["s"]
        lv = verts[p, j, :nv[p, j]]
        return np.stack([a * lv[:, 0] + b * lv[:, 1] + q[qi]["x"], -b * lv[:, 0] + a * lv[:, 1] + q[qi]["y"]], -1)

    obst, onv, off = [], [], [0]
    for t in range(nsteps):
        for _ in range(int(rng.integers(0, max_obst + 1))):
            n = 4 if obst_verts == 4 else int(rng.integers(3, obst_verts + 1))
            if obst_verts == 4 and rng.random() < 0.7:
                B = rect_vertices(np.array(0.0), np.array(0.0), np.array(rng.uniform(-np.pi, np.pi)),
                                  rng.uniform(3, 6), rng.uniform(1.5, 2.3))
            else:
                B = random_convex(rng, n, rng.uniform(0.8, 3.0))
            kind = rng.random()
            qi = int(rng.integers(0, nqueries))
            j = t - q[qi]["t0"]
            p = int(rng.integers(0, P))
            if kind < 0.6 and 0 <= j < J and nv[p, j] > 0:
                # slide B along a random direction until it (nearly) touches the occupancy (p, j) of query qi
                A = world(qi, p, j)
                th = rng.uniform(0, 2 * np.pi)
                d = np.array([np.cos(th), np.sin(th)])
                B = B + A.mean(0) + 30.0 * d
                lo, hi = 0.0, 30.0
                for _ in range(60):                       # bisection on "separated" by projection on all axes
                    mid = 0.5 * (lo + hi)
                    if _sat_sep(A, B - mid * d) > 0:
                        lo = mid
                    else:
                        hi = mid
                nudge = rng.choice([0.0, 0.0, 1e-15, -1e-15, 1e-12, -1e-12, 1e-9, -1e-9, 1e-6, -1e-6, 1e-3, -1e-3])
                B = B - (lo + nudge) * d
            else:
                B = B + [ox + rng.uniform(-60, 60), oy + rng.uniform(-60, 60)]
            if rng.random() < 0.5:
                B = B[::-1]
            v = np.zeros((max(obst_verts, 4), 2))
            v[:len(B)] = B
            obst.append(v)
            onv.append(len(B))
        off.append(len(obst))
    Vo = max(obst_verts, 4)
    obst = np.array(obst).reshape(-1, Vo, 2) if obst else np.zeros((0, Vo, 2))

    segs, sb, se = [], [], []
    if with_region:
        # region: big box with a rectangular hole; every second step uses a second, smaller region; one step empty;
        # one step without constraint
        def ring_segs(r):
            return [r[i] + r[(i + 1) % len(r)] for i in range(len(r))]

        big = [(ox - 80.0, oy - 80.0), (ox + 85.0, oy - 80.0), (ox + 85.0, oy + 80.0), (ox - 80.0, oy + 80.0)]
        hole = [(ox + 6.0, oy + 4.0), (ox + 6.0, oy + 6.0), (ox + 9.0, oy + 6.0), (ox + 9.0, oy + 4.0)]
        small = [(ox - 42.0, oy - 39.0), (ox + 50.0, oy - 42.0), (ox + 52.0, oy + 40.0), (ox + 3.0, oy + 54.0), (ox - 41.0, oy + 38.0)]
        segs = ring_segs(big) + ring_segs(hole) + ring_segs(small)
        for t in range(nsteps):
            if t == 3:
                sb.append(-1); se.append(-1)
            elif t == 5:
                sb.append(4); se.append(4)
            elif t % 2:
                sb.append(8); se.append(13)
            else:
                sb.append(0); se.append(8)
    else:
        sb, se = [-1] * nsteps, [-1] * nsteps
    scenes = dict(scene_step_off=np.array([0, nsteps], np.int32), step_obst_off=np.array(off, np.int32), obst=obst,
                  obst_nv=np.array(onv, np.int32), step_seg_begin=np.array(sb, np.int32),
                  step_seg_end=np.array(se, np.int32), segs=np.array(segs, dtype=np.float64).reshape(-1, 4))
    origins = np.array([[ox, oy]])
    return lib, scenes, origins, q


def _sat_sep(A, B):
    """float64 signed separation of two convex polygons (any orientation); > 0 apart"""
    best = -np.inf
    for P, Q in ((A, B), (B, A)):
        area = 0.5 * np.sum(P[:, 0] * np.roll(P[:, 1], -1) - np.roll(P[:, 0], -1) * P[:, 1])
        e = np.roll(P, -1, axis=0) - P
        n = np.stack([e[:, 1], -e[:, 0]], -1) * (1.0 if area > 0 else -1.0)
        n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-300)
        d = (Q[None, :, :] - P[:, None, :]) @ n[:, :, None]
        best = max(best, d[..., 0].min(axis=1).max())
    return best
