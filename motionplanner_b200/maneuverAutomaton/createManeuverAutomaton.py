"""Generator of the bundled maneuver automaton (restates reference maneuverAutomaton/createManeuverAutomaton.py:14-98:
same acceleration / orientation / velocity grids, same feasibility rules, scipy ``solve_ivp`` with default RK45
tolerances and dense output sampled at the 0.1 s time points, car outline in the rear-axle frame).

The pickled result of the reference (``maneuverAutomaton.obj``) is a large blob that is not shipped with the sources;
``create_maneuver_automaton()`` regenerates it (13 702 primitives, 149 824 occupancy rectangles, about 10-20 s) and
``load_maneuver_automaton()`` caches the arrays in an ``.npz`` next to this file keyed on the generator inputs.
"""
import hashlib
import os
import pickle

import numpy as np
from scipy.integrate import solve_ivp

from ..geometry import Polygon
from ..vehicle.vehicleParameter import vehicleParameter
from .ManeuverAutomaton import ManeuverAutomaton, MotionPrimitive

ACCELERATIONS = [-8, -4, -2, -1.2, -0.8, -0.4, 0, 0.4, 0.8, 1.2, 2, 4, 8]
ORIENTATIONS = [-1, -0.8, -0.6, -0.4, -0.3, -0.2, -0.1, 0, 0.1, 0.2, 0.3, 0.4, 0.6, 0.8, 1]
V_START, V_END, V_DIFF = 0, 30, 0.4
T_FINAL, DT = 1, 0.1


def _plan(param, v_start=V_START, v_end=V_END, v_diff=V_DIFF, accelerations=ACCELERATIONS,
          orientations=ORIENTATIONS, t_final=T_FINAL, dt=DT):
    """enumerate (v_init, acc, steer, tFinal) of every feasible primitive in reference order"""
    plan = []
    v_init = v_start
    old = np.seterr(divide='ignore', invalid='ignore')     # tf == 0 branches divide by zero and are discarded below
    while v_init < v_end:
        for acc in accelerations:
            # stretch / shrink the primitive so that it ends on the velocity grid inside [v_start, v_end]
            if v_start <= v_init + acc * t_final <= v_end:
                tf = t_final
            elif v_init + acc * t_final < v_start:
                tf = np.round(((v_start - v_init) / acc) / dt) * dt
                acc = (v_start - v_init) / tf
            else:
                tf = np.round(((v_start - v_end) / acc) / dt) * dt
                acc = (v_start - v_end) / tf
            if tf > 0:
                for o in orientations:
                    dist = v_init * tf + 0.5 * acc * tf ** 2
                    if o == 0 or abs(dist) > 0:
                        steer = 0 if o == 0 else np.arctan(param['wheelbase'] * o / dist)
                        if abs(steer) < param['s_max']:
                            plan.append((v_init, acc, steer, tf))
        v_init = v_init + v_diff
    np.seterr(**old)
    return plan


def car_outline(param):
    """car rectangle in the rear-axle frame, reference vertex order (clockwise)"""
    rear, front, half = -(param['length'] / 2 - param['b']), param['length'] / 2 + param['b'], param['width'] / 2
    return [(rear, -half), (rear, half), (front, half), (front, -half)]


def _simulate(plan, param, dt=DT):
    """states [4 x n+1] of every primitive"""
    wb = param['wheelbase']
    out = []
    for v_init, acc, steer, tf in plan:
        def ode(t, x, u1, u2):
            return [x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u1, x[2] * np.tan(u2) / wb]
        sol = solve_ivp(ode, [0, tf], [0, 0, v_init, 0], args=(acc, steer), dense_output=True)
        t = np.linspace(0, tf, int(np.round(tf / dt)) + 1)
        out.append(sol.sol(t))
    return out


def _occupancy(x, car, dt=DT):
    """car outline at every state: shapely affine_transform(car, [cos, -sin, sin, cos, x, y]) = a*x + b*y + xoff per
    point, evaluated for all states at once (same arithmetic, element by element)"""
    car = np.asarray(car)
    c, s = np.cos(x[3])[:, None], np.sin(x[3])[:, None]
    px = c * car[:, 0] + (-s) * car[:, 1] + x[0][:, None]
    py = s * car[:, 0] + c * car[:, 1] + x[1][:, None]
    verts = np.stack([px, py], axis=-1)                       # [n, 4, 2]
    return [{'space': Polygon(verts[i]), 'time': dt * i} for i in range(x.shape[1])], verts


def create_maneuver_automaton(param=None, states=None, plan=None):
    """build the ManeuverAutomaton object (states/plan may come from the cache)"""
    param = param or vehicleParameter()
    plan = plan if plan is not None else _plan(param)
    states = states if states is not None else _simulate(plan, param)
    car = car_outline(param)
    prims = []
    for (v, acc, steer, tf), x in zip(plan, states):
        occ, verts = _occupancy(x, car)
        p = MotionPrimitive(x, np.array([acc, steer]), tf, occupancy=occ)
        p._occ_verts = verts            # the same vertices as an array: lets device_layout() skip 150 k Python objects
        prims.append(p)
    return ManeuverAutomaton(prims)


def _cache_path(param):
    h = hashlib.sha1(repr((sorted(param.items()), ACCELERATIONS, ORIENTATIONS, V_START, V_END, V_DIFF, T_FINAL, DT)).encode()).hexdigest()[:12]
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_cache')
    return d, os.path.join(d, 'automaton_%s.npz' % h)


def load_maneuver_automaton(param=None, use_cache=True):
    """regenerate (or load from the .npz cache) the automaton the reference pickles to maneuverAutomaton.obj"""
    param = param or vehicleParameter()
    d, path = _cache_path(param)
    plan = _plan(param)
    if use_cache and os.path.exists(path):
        z = np.load(path)
        flat, lens = z['flat'], z['lens']
        if len(lens) == len(plan):
            states, o = [], 0
            for n in lens:
                states.append(flat[:, o:o + n].copy())
                o += n
            return create_maneuver_automaton(param, states, plan)
    states = _simulate(plan, param)
    if use_cache:
        try:
            os.makedirs(d, exist_ok=True)
            np.savez_compressed(path, flat=np.concatenate(states, axis=1), lens=np.array([s.shape[1] for s in states]))
        except OSError:
            pass
    return create_maneuver_automaton(param, states, plan)


if __name__ == '__main__':
    MA = create_maneuver_automaton()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'maneuverAutomaton.obj'), 'wb') as f:
        pickle.dump(MA, f)
