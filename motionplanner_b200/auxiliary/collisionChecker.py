"""Post-hoc trajectory validator -- drop-in for reference auxiliary/collisionChecker.py:6-47.

``collisionChecker(scenario, traj, param) -> bool`` (True = collision-free): for every time step the ego rectangle
must lie inside the road polygon and must not intersect (closed sets) any obstacle occupancy of that step.
All N steps are decided by one GPU query (one candidate per step) instead of N*(O+1) shapely calls.
"""
import numpy as np

from ..checker import SceneSet, car_library, get_engine, scenario_obstacle_arrays, validate_trajectory
from .polygonLaneletNetwork import polygonLaneletNetwork

_ROAD_CACHE = {}


def _road_of(scenario):
    """the reference rebuilds the road polygon on every call (:10); the union is cached per scenario object here"""
    key = id(scenario)
    hit = _ROAD_CACHE.get(key)
    if hit is None or hit[0] is not scenario:
        _ROAD_CACHE.clear()
        hit = (scenario, polygonLaneletNetwork(scenario))
        _ROAD_CACHE[key] = hit
    return hit[1]


def collision_flags(scenario, traj, param, engine=None):
    """per-step flags (True = this step violates road containment or touches an obstacle) and device statistics"""
    engine = engine if engine is not None else get_engine('validator')
    traj = np.asarray(traj, dtype=np.float64)
    n = traj.shape[1]
    road = _road_of(scenario)
    off, obst, nv = scenario_obstacle_arrays(scenario, n)
    key = ('car', float(param['length']), float(param['width']))
    if getattr(engine, '_staged_library', None) != key:
        lib = car_library(param['length'], param['width'])
        engine.upload_library(lib['verts'], lib['nv'], lib['tidx'])
        try:
            engine._staged_library = key
        except AttributeError:
            pass
    scenes = SceneSet()
    scenes.add_scene_explicit(n, off, obst, nv, road.segments, origin=traj[0:2, 0])
    scenes.stage(engine)
    return validate_trajectory(engine, traj, param['length'], param['width'])


def collisionChecker(scenario, traj, param):
    """check if the planned trajectory collides with obstacles or with the road boundary"""
    flags, _ = collision_flags(scenario, traj, param)
    return not bool(flags.any())


def get_shapely_object(set):
    """reference :37-47 returns one shapely polygon (union of a ShapeGroup's members); this package consumes shapes
    piecewise, so the geometry container of the shape (or the list of member containers) is returned"""
    if hasattr(set, 'shapes'):
        return [s.shapely_object for s in set.shapes]
    return set.shapely_object
