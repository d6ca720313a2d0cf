"""CPU tests of the host-side mirror of the reference interface (no GPU: the oracle stands in for the engine)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, OracleBackend, load_golden, random_walk, scenario_scene
from motionplanner_b200.checker import PrimitiveChecker, SceneSet, car_library, convex_pieces, validate_trajectory
from motionplanner_b200.engine import MODE_PLANNER
from motionplanner_b200.geometry import MultiPolygon, Point, Polygon, ring_segments, rings_of
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import _plan, load_maneuver_automaton
from motionplanner_b200.maneuverAutomaton.ManeuverAutomaton import ManeuverAutomaton, MotionPrimitive
from motionplanner_b200.vehicle.vehicleParameter import vehicleParameter
from oracle import exact as ex


@pytest.fixture(scope="module")
def MA():
    return load_maneuver_automaton()


def test_library_matches_reference_counts(MA):
    """76 start velocities, 13 702 primitives, 149 824 occupancy rectangles, branching <= 195"""
    assert len(MA.primitives) == 13702
    assert sum(len(p.occ) for p in MA.primitives) == 149824
    assert sum(1 for b in MA.vel_dic if b) == 76
    assert max(len(b) for b in MA.vel_dic) == 195
    assert sum(1 for p in MA.primitives if len(p.occ) == 11) == 13453
    p = MA.primitives[0]
    assert p.x.shape[0] == 4 and p.occ[0]['time'] == 0.0 and MA.is_timepoint()
    # successors = the bucket of the final velocity (reference ManeuverAutomaton.py:24-27)
    for p in MA.primitives[::997]:
        assert p.successors is MA.velocity2primitives(p.x[2, -1])
    # car outline in the rear-axle frame at t = 0 (reference createManeuverAutomaton.py:77-80)
    v = np.array(p.occ[0]['space'].exterior.coords[:-1])
    assert np.allclose(v, [[-1.0, -0.85], [-1.0, 0.85], [3.3, 0.85], [3.3, -0.85]])


def test_device_layout_time_indices(MA):
    lay = MA.device_layout(0.1)
    assert lay['verts'].shape == (13702, 11, 4, 2)
    assert np.array_equal(lay['tidx'][0], np.arange(11))          # int(0.1*i/0.1) == i for i = 0..10 (quirk Q2)
    assert lay['set_off'][-1] == 13702 and len(lay['set_off']) == len(MA.vel_dic) + 1
    assert (lay['nv'] == 0).sum() == 13702 * 11 - 149824


def test_irregular_and_interval_library_is_reslotted():
    """AROC-style interval occupancies: int(0.6/0.1) == 5 == int(0.5/0.1) -> two slots with time index 5 (quirk Q2)"""
    sq = Polygon([(0, 0), (1, 0), (1, 1), (0, 1)])
    occ = [{'space': sq, 'starttime': 0.1 * 0, 'endtime': 0.1}, {'space': sq, 'starttime': 0.5, 'endtime': 0.6},
           {'space': sq, 'starttime': 0.6, 'endtime': 0.7}, {'space': sq, 'starttime': 0.7, 'endtime': 0.8}]
    x = np.zeros((4, 2))
    x[2] = [1.0, 1.0]
    prims = [MotionPrimitive(x.copy(), np.zeros(2), 0.8, occupancy=occ),
             MotionPrimitive(x.copy() + np.array([[0], [0], [0.4], [0]]), np.zeros(2), 0.8, occupancy=occ[:2])]
    MA2 = ManeuverAutomaton(prims)
    assert not MA2.is_timepoint()
    lay = MA2.device_layout(0.1)
    assert sorted(lay['tidx'][0].tolist()) == [0, 5, 5, 6]
    assert lay['nv'][0].tolist().count(4) == 4 and lay['nv'][1].tolist().count(4) == 2
    # every slot has one time index for all primitives that use it
    for j in range(lay['tidx'].shape[1]):
        assert len(set(lay['tidx'][lay['nv'][:, j] > 0, j].tolist())) <= 1


def test_geometry_containers():
    p = Polygon([(0, 0), (4, 0), (4, 4), (0, 4)], holes=[[(1, 1), (2, 1), (2, 2), (1, 2)]])
    assert len(rings_of(p)) == 2 and ring_segments(rings_of(p)).shape == (8, 4)
    assert p.contains(Point(3, 3)) and not p.contains(Point(1.5, 1.5)) and not p.contains(Point(4, 2))
    assert abs(p.area - 15.0) < 1e-12
    assert abs(p.exterior.distance(Point(2, 5)) - 1.0) < 1e-12
    mp = MultiPolygon([p, Polygon([(10, 0), (11, 0), (11, 1), (10, 1)])])
    assert len(rings_of(mp)) == 3 and mp.contains(Point(10.5, 0.5))
    assert rings_of(Polygon()) == [] and Polygon().is_empty


def test_convex_pieces_cover_nonconvex_polygon():
    class S:
        vertices = np.array([(0, 0), (6, 0), (6, 2), (2, 2), (2, 6), (0, 6), (0, 0)], float)
    pieces = convex_pieces(S)
    assert all(len(t) == 3 for t in pieces)
    area = sum(0.5 * abs((t[1] - t[0])[0] * (t[2] - t[0])[1] - (t[1] - t[0])[1] * (t[2] - t[0])[0]) for t in pieces)
    assert abs(area - 20.0) < 1e-12

    class G:
        shapes = [S, S]
    assert len(convex_pieces(G)) == 2 * len(pieces)


def test_golden_fixture_statistics():
    """obstacles per scenario 1 / 9 / 54, horizon 15 / 33 / 147"""
    g = load_golden()
    assert len(g) == 100
    nob = np.array([v['nobstacles'] for v in g.values()])
    hor = np.array([v['t_end'] for v in g.values()])
    assert (nob.min(), int(np.median(nob)), nob.max()) == (1, 9, 54)
    assert (hor.min(), int(np.median(hor)), hor.max()) == (15, 33, 147)
    assert all(len(v['road']) > 0 for v in g.values())


@pytest.mark.skipif(not os.path.isdir('/root/reference/scenarios'), reason="reference scenarios not mounted")
def test_golden_fixture_reproducible_from_reference():
    """the committed fixture equals what tools/make_golden.py extracts from the reference's XML (spot check)"""
    from motionplanner_b200.auxiliary.polygonLaneletNetwork import polygonLaneletNetwork
    from motionplanner_b200.checker import scenario_obstacle_arrays
    from motionplanner_b200.crlite import CommonRoadFileReader
    g = load_golden()
    for name in ('ZAM_Zip-1_19_T-1', 'USA_US101-16_2_T-1', 'ZAM_HW-1_1_S-1'):
        sc, pps = CommonRoadFileReader('/root/reference/scenarios/%s.xml' % name).open()
        off, obst, nv = scenario_obstacle_arrays(sc, g[name]['nsteps'])
        assert np.array_equal(obst, g[name]['obst']) and np.array_equal(off, g[name]['step_obst_off'])
        assert np.array_equal(polygonLaneletNetwork(sc).segments, g[name]['road'])


def test_road_boundary_is_watertight_on_fixture():
    import collections
    g = load_golden()
    for name in ('ZAM_Zip-1_19_T-1', 'USA_US101-8_4_T-1', 'DEU_Flensburg-10_1_T-1'):
        deg = collections.Counter()
        for s in g[name]['road']:
            deg[(s[0], s[1])] += 1
            deg[(s[2], s[3])] += 1
        assert all(v % 2 == 0 for v in deg.values()), name


def test_planner_queries_against_exact_oracle(MA):
    """the queries the batched planner issues, on ZAM_Zip-1_19 (configs[0]) -- the C oracle backend against the
    definitional rational oracle on a sample of decisions"""
    g = load_golden()['ZAM_Zip-1_19_T-1']
    be = OracleBackend()
    chk = PrimitiveChecker(MA, 0.1, be)
    chk.set_scenes(scenario_scene(g), MODE_PLANNER)
    bits, traj = random_walk(chk, MA, g, np.random.default_rng(0), expansions=3)
    assert len(bits) == 3 and traj.shape[0] == 4
    q, cand, mode, out = be.log[0]
    lay = MA.device_layout(0.1)
    cands = lay['set_idx'][lay['set_off'][q['cand_set'][0]]:lay['set_off'][q['cand_set'][0] + 1]]
    rings = [[(s[0], s[1]), (s[2], s[3])] for s in g['road_safe']]

    def inside_soup(p):          # even-odd over the segment soup with exact orientation signs
        inside = False
        for a, b in rings:
            if (a[1] > p[1]) != (b[1] > p[1]):
                o = ex.orient(a, b, p)
                if (b[1] > a[1] and o > 0) or (b[1] < a[1] and o < 0):
                    inside = not inside
        return inside
    rng = np.random.default_rng(1)
    for i in rng.choice(len(cands), 12, replace=False):
        p = cands[i]
        collide = False
        for j in range(11):
            if lay['nv'][p, j] == 0:
                continue
            t = int(q['t0'][0]) + int(lay['tidx'][p, j])
            if t > g['t_end'] or t >= g['nsteps']:
                continue
            A = ex.affine_transform([tuple(v) for v in lay['verts'][p, j, :4]], q['c'][0], q['s'][0], q['x'][0], q['y'][0])
            ok = not any(ex.segment_meets_interior(r[0], r[1], A) for r in rings)
            ok = ok and inside_soup(ex.interior_point(A))
            for o in range(g['step_obst_off'][t], g['step_obst_off'][t + 1]):
                B = [tuple(v) for v in g['obst'][o, :g['obst_nv'][o]]]
                ok = ok and not ex.convex_interiors_overlap(A, B)
            collide = collide or not ok
        assert bool(out[i]) == collide


def test_scene_args_marshalling_without_a_device():
    """CollisionEngine.scene_args checks and marshals the arrays of mpc_set_scenes once (lanes reuse the tuple every
    cycle); it needs no context, keeps the converted arrays alive and rejects inconsistent shapes"""
    from motionplanner_b200 import workloads as W
    from motionplanner_b200.engine import CollisionEngine
    lib, sc, org, q = W.adversarial_problem(3)
    keys = ("scene_step_off", "origins", "step_obst_off", "obst", "obst_nv", "step_seg_begin", "step_seg_end", "segs")
    scene = dict(sc, origins=org)
    args = CollisionEngine.scene_args(*[scene[k] for k in keys])
    assert len(args) == 12 and args[0] == 1 and args[6] == sc["obst"].shape[1] and args[10] == len(sc["segs"])
    keep = args[11]
    assert keep[3].dtype == np.float64 and keep[3].flags.c_contiguous and keep[4].dtype == np.int32
    assert args[4] == keep[3].ctypes.data                           # address of the kept obstacle array
    bad = dict(scene, obst_nv=sc["obst_nv"][:-1])
    with pytest.raises(AssertionError):
        CollisionEngine.scene_args(*[bad[k] for k in keys])
    none = dict(scene, obst=None, obst_nv=None, step_obst_off=np.zeros_like(sc["step_obst_off"]))
    a2 = CollisionEngine.scene_args(*[none[k] for k in keys])
    assert a2[4] is None and a2[5] is None                           # no obstacles: null pointers, as the ABI allows


def test_convexity_test_rejects_self_wrapping_rings_and_triangulation_covers():
    """`_is_convex` must not accept a ring that turns one way but wraps twice; `_triangulate` must cover the
    polygon or raise -- never drop a part of an obstacle silently"""
    from motionplanner_b200.checker import _is_convex, _triangulate, convex_pieces
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2.0]])
    assert _is_convex(sq) and _is_convex(sq[::-1])
    star = np.array([[np.cos(a), np.sin(a)] for a in np.arange(5) * (4 * np.pi / 5)])      # pentagram: turns left, twice round
    assert not _is_convex(star)
    L = np.array([[0, 0], [3, 0], [3, 1], [1, 1], [1, 3], [0, 3.0]])
    tris = _triangulate(L)
    area = sum(0.5 * abs((t[1, 0] - t[0, 0]) * (t[2, 1] - t[0, 1]) - (t[1, 1] - t[0, 1]) * (t[2, 0] - t[0, 0])) for t in tris)
    assert abs(area - 5.0) < 1e-12
    bow = np.array([[0, 0], [2, 2], [2, 0], [0, 2.0]])                                      # self-intersecting
    with pytest.raises(ValueError):
        _triangulate(bow)

    class Circle:
        center, radius = np.array([3.0, -1.0]), 2.0

    pieces = convex_pieces(Circle())
    assert all(len(p) <= 8 and _is_convex(p) for p in pieces)
    area = sum(0.5 * abs(np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])) for p in pieces)
    assert abs(area - 0.5 * 64 * 4.0 * np.sin(2 * np.pi / 64)) < 1e-12                      # area of the inscribed 64-gon


def test_primitive_registry_finds_the_automaton():
    from motionplanner_b200.maneuverAutomaton.ManeuverAutomaton import (ManeuverAutomaton, MotionPrimitive,
                                                                        automaton_of)
    prims = []
    for i in range(6):
        x = np.zeros((4, 11))
        x[2] = 2.0 if i < 3 else 2.4
        prims.append(MotionPrimitive(x, np.zeros(2), 1.0, occupancy=[{'space': None, 'time': 0.0}]))
    MA = ManeuverAutomaton(prims)
    assert automaton_of(prims[4]) == (MA, 4)
    import pickle
    MA2 = pickle.loads(pickle.dumps(MA))
    assert automaton_of(MA2.primitives[2]) == (MA2, 2)
    with pytest.raises(KeyError):
        automaton_of(MotionPrimitive(np.zeros((4, 11)), np.zeros(2), 1.0))


def test_invalid_ring_detector_and_committed_count():
    """tools/count_invalid_rings.py: exact self-intersection test on rings (the cases where the reference applies
    buffer(0), auxiliary/polygonLaneletNetwork.py:83-85), and the committed result for the 100 bundled scenarios:
    the lobes buffer(0) would drop are covered by other polygons of the same union everywhere"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("cir", os.path.join(os.path.dirname(__file__), "..", "tools", "count_invalid_rings.py"))
    cir = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cir)
    square = np.array([[0, 0], [2, 0], [2, 2], [0, 2.0]])
    assert cir.self_intersections(square) == []
    bow = np.array([[0, 0], [2, 2], [2, 0], [0, 2.0]])
    hits = cir.self_intersections(bow)
    assert len(hits) == 1 and hits[0][2] == 'cross'
    touch = np.array([[0, 0], [4, 0], [4, 2], [2, 0], [0, 2.0]])          # a vertex on a non-adjacent edge
    assert [h[2] for h in cir.self_intersections(touch)] == ['touch', 'touch']      # the two edges that meet in that vertex
    pos, neg, minority, unc, cell = cir.lobe_areas(bow, [np.array([[-1.0, -1], [3, -1], [3, 3], [-1, 3]])], grid=200)
    assert abs(pos - 1.0) < 0.05 and abs(neg - 1.0) < 0.05 and abs(minority - 1.0) < 0.05
    assert unc == 0.0                                                # another polygon of the union covers both lobes
    assert abs(cir.lobe_areas(bow, [square + 10.0], grid=200)[3] - minority) < 1e-12      # ... or none
    rep = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "r02_invalid_rings.json")))
    assert rep["scenarios"] == 100 and rep["invalid_rings"] > 0
    assert rep["minority_lobe_area_not_covered_elsewhere_m2"] == 0.0
