"""input guards of the C ABI that need a device: non-convex polygons are rejected (the separating-axis kernels
are only correct for convex pieces; a reflex edge would 'separate' what it must not -- an unsafe false negative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from motionplanner_b200 import workloads as W
from motionplanner_b200.engine import CollisionEngine, MpcError


def _stage(eng, sc, org):
    eng.set_scenes(sc["scene_step_off"], org, sc["step_obst_off"], sc["obst"], sc["obst_nv"], sc["step_seg_begin"],
                   sc["step_seg_end"], sc["segs"])


def test_nonconvex_library_polygon_is_rejected():
    eng = CollisionEngine(0)
    lib, sc, org, q = W.adversarial_problem(3, lib_verts=8)
    bad = lib["verts"].copy()
    p, j = np.argwhere(lib["nv"] >= 5)[0]
    n = lib["nv"][p, j]
    bad[p, j, 1] = bad[p, j, :n].mean(axis=0)                  # pull one vertex to the centre: a reflex corner
    with pytest.raises(MpcError, match="not convex"):
        eng.upload_library(bad, lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    star = lib["verts"].copy()                                 # pentagram: every turn to the left, two revolutions
    star[p, j, :5] = [[np.cos(a), np.sin(a)] for a in np.arange(5) * (4 * np.pi / 5)]
    nv = lib["nv"].copy()
    nv[p, j] = 5
    with pytest.raises(MpcError, match="convex"):
        eng.upload_library(star, nv, lib["tidx"], lib["set_off"], lib["set_idx"])
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])      # the context survives
    _stage(eng, sc, org)
    assert len(eng.query(q)[0]) == int(q["cand_cnt"].sum())
    eng.close()


def test_nonconvex_obstacle_piece_is_rejected_and_repeated_vertices_are_not():
    eng = CollisionEngine(0)
    lib, sc, org, q = W.adversarial_problem(4, obst_verts=8)
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    _stage(eng, sc, org)
    want, _ = eng.query(q)
    k = int(np.argmax(sc["obst_nv"] >= 5))
    bad = dict(sc, obst=sc["obst"].copy())
    bad["obst"][k, 1] = bad["obst"][k, :sc["obst_nv"][k]].mean(axis=0)
    with pytest.raises(MpcError, match="not convex"):
        _stage(eng, bad, org)
    with pytest.raises(MpcError):
        eng.query(q)                                           # no scenes staged after the failure
    # a repeated vertex (zero-length edge) is legal input: same decisions as without it
    rep = dict(sc, obst=sc["obst"].copy(), obst_nv=sc["obst_nv"].copy())
    k4 = int(np.argmax(sc["obst_nv"] <= 6))
    n = int(sc["obst_nv"][k4])
    rep["obst"][k4, 2:n + 1] = sc["obst"][k4, 1:n]
    rep["obst"][k4, 1] = sc["obst"][k4, 1]
    rep["obst_nv"][k4] = n + 1
    _stage(eng, rep, org)
    got, _ = eng.query(q)
    assert np.array_equal(got, want)
    eng.close()


def test_library_with_repeated_vertices_decides_like_the_clean_one():
    eng = CollisionEngine(0)
    lib, sc, org, q = W.adversarial_problem(5)
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
    _stage(eng, sc, org)
    want, st0 = eng.query(q)
    V = lib["verts"].shape[2]
    verts = np.zeros(lib["verts"].shape[:2] + (V + 1, 2))
    verts[:, :, :V] = lib["verts"]
    nv = lib["nv"].copy()
    for p, j in np.argwhere(lib["nv"] == 4)[::3]:
        verts[p, j, :5] = lib["verts"][p, j][[0, 1, 1, 2, 3]]
        nv[p, j] = 5
    eng.upload_library(verts, nv, lib["tidx"], lib["set_off"], lib["set_idx"])
    got, st1 = eng.query(q)
    assert np.array_equal(got, want)
    assert st1["refined_fp64"] == st0["refined_fp64"]         # no extra pairs pushed into the tie-break pass
    eng.close()
