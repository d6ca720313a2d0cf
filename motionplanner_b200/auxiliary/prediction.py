"""Obstacle occupancies from a constant-velocity lane-following prediction, emitted directly as the arrays
mpc_set_scenes ingests.

Mirror of reference auxiliary/prediction.py: `prediction(scenario, horizon, x_ego, ...)` (:15-174) replaces every
dynamic obstacle of the scenario by its predicted versions -- one per route through the successor lanelets, driven at
constant acceleration along the centre lines, cut where the vehicle would run into the ego vehicle from behind -- and
returns the scenario.  The reference then walks `occupancy_at_time(i)` of every obstacle and builds a shapely polygon
per step (auxiliary/collisionChecker.py:25-33, src/maneuverAutomatonPlannerStandalone.py:165-180); here

    routes = predict_routes(scenario, horizon, x_ego)           # positions + headings per step, no shape objects
    step_obst_off, obst, obst_nv = occupancy_arrays(routes, nsteps)     # [n, 4, 2] rectangles, ready for SceneSet

produce the same rectangles (same arithmetic as commonroad's Rectangle.rotate_translate_local, bit for bit: tested
against the object path) with a handful of numpy calls per cycle.

Not restated: the `shape_lanelet_assignment` of the predicted obstacles (:148-160), bookkeeping for the high-level
planner's lane logic that the collision path never reads; `overlapping_lanelets` is accepted and ignored.
"""
import numpy as np

from ..crlite import Obstacle, Rectangle, State, TrajectoryPrediction, make_valid_orientation


def distance_on_lanelet(lanelet, x, y):
    """arc length along the centre line of the point of the centre line closest to (x, y)   (reference :177-193)"""
    c = lanelet.center_vertices
    a, d = c[:-1], c[1:] - c[:-1]
    p = np.array([x, y], dtype=np.float64)
    n2 = np.maximum((d * d).sum(axis=1), 1e-300)
    t = np.clip(((p - a) * d).sum(axis=1) / n2, 0.0, 1.0)
    foot = a + t[:, None] * d
    k = int(np.argmin(np.linalg.norm(foot - p, axis=1)))        # first closest segment, like the strict `<` of the loop
    return float(lanelet.distance[k] + np.linalg.norm(foot[k] - a[k]))


def travelled_distance(velocity, acceleration, dt, horizon):
    """distance covered up to every step under constant acceleration, standing still once the speed is <= 0 (:58-66)"""
    vel = velocity + dt * acceleration * np.linspace(0, horizon, num=horizon + 1)
    x = np.zeros(len(vel))
    for i in range(len(vel) - 1):
        x[i + 1] = x[i] + vel[i] * dt + 0.5 * acceleration * dt ** 2 if vel[i] > 0 else x[i]
    return x


def _pose_on_lanelet(lanelet, dist):
    """(position, heading) at arc length `dist` of the centre line; None when no segment ends beyond it (:122-131)"""
    j = int(np.searchsorted(lanelet.distance[1:], dist, side='right'))       # first j with distance[j+1] > dist
    if j >= len(lanelet.distance) - 1:
        return None
    d = lanelet.center_vertices[j + 1, :] - lanelet.center_vertices[j, :]
    d = d / np.linalg.norm(d)
    return lanelet.center_vertices[j, :] + d * (dist - lanelet.distance[j]), float(np.arctan2(d[1], d[0]))


def predict_routes(scenario, horizon, x_ego, most_likely=True):
    """one entry per predicted trajectory: dict(position [n,2], orientation [n], velocity, length, width, lanelets [n],
    source = the original obstacle).  Step k of a route is time step k."""
    net = scenario.lanelet_network
    lanelets = {l.lanelet_id: l for l in net.lanelets}
    dt = scenario.dt
    ego_lanes = net.find_lanelet_by_position([np.asarray(x_ego[0:2], dtype=np.float64)])[0]
    ego_dist = [distance_on_lanelet(lanelets[i], x_ego[0], x_ego[1]) for i in ego_lanes]
    reach_ego = x_ego[2] * dt * horizon + 0.5 * 9 * (dt * horizon) ** 2          # :54
    routes = []
    for obs in scenario.dynamic_obstacles:
        s, shape = obs.initial_state, obs.obstacle_shape
        acc = getattr(s, 'acceleration', 0) or 0
        px, py = float(s.position[0]), float(s.position[1])
        x = travelled_distance(s.velocity, acc, dt, horizon)
        if np.sqrt((px - x_ego[0]) ** 2 + (py - x_ego[1]) ** 2) > reach_ego + x[-1]:     # cannot meet the ego (:69-70)
            continue
        first = State(position=np.array([px, py]), velocity=s.velocity, orientation=s.orientation, time_step=0)
        start = (net.find_most_likely_lanelet_by_state([first]) if most_likely
                 else net.find_lanelet_by_position([first.position])[0])
        # breadth-first over the routes: (poses so far, lanelet ids so far, arc-length offset, current lanelet)
        todo = [([(first.position, first.orientation)], [lid], -distance_on_lanelet(lanelets[lid], px, py), lid)
                for lid in start]
        while todo:
            poses, lanes, offset, lid = todo.pop(0)
            poses, lanes = list(poses), list(lanes)
            l = lanelets[lid]
            for i in range(len(poses), horizon + 1):
                dist = x[i] - offset
                behind_ego = any(lid == e and de > dist > de - 5 - shape.length / 2 for e, de in zip(ego_lanes, ego_dist))
                if behind_ego and len(poses) > 1:                   # would run into the ego from behind: cut here (:107-116)
                    routes.append(_route(poses, lanes, s.velocity, shape, obs))
                    break
                if dist > l.distance[-1]:                            # continue on every successor (:118-121)
                    for nxt in l.successor:
                        todo.append((poses, lanes, offset + l.distance[-1], nxt))
                    break
                pose = _pose_on_lanelet(l, dist)
                if pose is not None:
                    poses.append(pose)
                    lanes.append(lid)
                if i == horizon:
                    routes.append(_route(poses, lanes, s.velocity, shape, obs))
    return routes


def _route(poses, lanes, velocity, shape, source):
    return dict(position=np.array([p for p, _ in poses], dtype=np.float64).reshape(-1, 2),
                orientation=np.array([o for _, o in poses], dtype=np.float64), velocity=velocity,
                length=float(shape.length), width=float(shape.width), lanelets=list(lanes), source=source)


def prediction(scenario, horizon, x_ego, overlapping_lanelets=None, most_likely=True):
    """predict the future positions of the surrounding vehicles: the scenario with its dynamic obstacles replaced by
    the predicted ones (reference :15-174)"""
    routes = predict_routes(scenario, horizon, x_ego, most_likely)
    for obs in list(scenario.dynamic_obstacles):
        scenario.remove_obstacle(obs)
    for r in routes:
        states = [State(position=r['position'][k], velocity=r['velocity'], orientation=r['orientation'][k], time_step=k)
                  for k in range(len(r['orientation']))]
        shape = Rectangle(width=r['width'], length=r['length'])
        scenario.add_objects(Obstacle(scenario.generate_object_id(), 'dynamic', 'car', shape, states[0],
                                      TrajectoryPrediction(states, shape)))
    scenario._predicted_routes = routes
    return scenario


def occupancy_arrays(routes, nsteps):
    """(step_obst_off [nsteps+1] i4, obst [n, 4, 2] f8, obst_nv [n] i4): the rectangle every predicted vehicle occupies
    at every step, step-major, routes in order -- the arrays `scenario_obstacle_arrays(prediction(...))` builds through
    occupancy_at_time + Rectangle objects, computed for all steps of a route at once with the same operations
    (commonroad Rectangle: corners (-l/2,-w/2), (-l/2,w/2), (l/2,w/2), (l/2,-w/2) rotated by the wrapped orientation and
    translated)."""
    per_step = [[] for _ in range(nsteps)]
    for r in routes:
        n = min(len(r['orientation']), nsteps)
        if n <= 0:
            continue
        l2, w2 = 0.5 * r['length'], 0.5 * r['width']
        lx = np.array([-l2, -l2, l2, l2])
        ly = np.array([-w2, w2, w2, -w2])
        ang = np.array([make_valid_orientation(0.0 + a) for a in r['orientation'][:n]])
        c, s = np.cos(ang)[:, None], np.sin(ang)[:, None]
        tx, ty = (0.0 + r['position'][:n, 0])[:, None], (0.0 + r['position'][:n, 1])[:, None]
        vx = c * lx + (-s) * ly + tx
        vy = s * lx + c * ly + ty
        rect = np.stack([vx, vy], axis=2)                        # [n, 4, 2]
        for k in range(n):
            per_step[k].append(rect[k])
    off = np.zeros(nsteps + 1, np.int32)
    off[1:] = np.cumsum([len(p) for p in per_step])
    flat = [q for p in per_step for q in p]
    obst = np.array(flat, dtype=np.float64).reshape(-1, 4, 2) if flat else np.zeros((0, 4, 2))
    return off, obst, np.full(len(obst), 4, np.int32)
