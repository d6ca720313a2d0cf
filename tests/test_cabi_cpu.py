"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol include/mpc.h
declares, struct layouts match, and the product path fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from motionplanner_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _cabi.load()
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), "libmpcheck.so does not export %s" % n
    assert sorted(_cabi.EXPORTS) == names


def test_struct_layouts_match_header():
    assert _cabi.QUERY_DTYPE.itemsize == 64
    assert [_cabi.QUERY_DTYPE.fields[k][1] for k in ("x", "y", "c", "s", "scene", "t0", "step_limit", "cand_set",
                                                      "cand_off", "cand_cnt")] == [0, 8, 16, 24, 32, 36, 40, 44, 48, 52]
    assert C.sizeof(_cabi.Stats) == 10 * 8 + 4 * 4
    assert C.sizeof(_cabi.DevInfo) == 64 + 4 * 4 + 8 + 4 + 4


def test_mode_constants_match_header():
    src = open(os.path.join(ROOT, "include", "mpc.h")).read()
    for name, val in (("MPC_MODE_VALIDATOR", _cabi.MODE_VALIDATOR), ("MPC_MODE_NO_CULL", _cabi.MODE_NO_CULL),
                      ("MPC_MODE_NO_EARLY_EXIT", _cabi.MODE_NO_EARLY_EXIT), ("MPC_MODE_COUNT_WORK", _cabi.MODE_COUNT_WORK),
                      ("MPC_MODE_NO_TIMING", _cabi.MODE_NO_TIMING)):
        m = re.search(r"#define\s+%s\s+(\d+)u" % name, src)
        assert m and int(m.group(1)) == val
    assert int(re.search(r"#define\s+MPC_MAX_LIB_VERTS\s+(\d+)", src).group(1)) == _cabi.MAX_LIB_VERTS
    assert int(re.search(r"#define\s+MPC_MAX_OBST_VERTS\s+(\d+)", src).group(1)) == _cabi.MAX_OBST_VERTS


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    from motionplanner_b200.engine import CollisionEngine, MpcError
    with pytest.raises(MpcError) as e:
        CollisionEngine(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under motionplanner_b200/ may reference it"""
    pkg = os.path.join(ROOT, "motionplanner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_workloads_are_deterministic():
    from motionplanner_b200 import workloads as W
    a = W.c4_problem(P=64, slots=4, T=6)
    b = W.c4_problem(P=64, slots=4, T=6)
    assert np.array_equal(a[0]["verts"], b[0]["verts"]) and np.array_equal(a[1]["obst"], b[1]["obst"])
    assert a[0]["verts"].shape == (64, 6, 4, 2) and a[1]["obst"].shape == (6 * 32, 4, 2)
    l1 = W.adversarial_problem(5)
    l2 = W.adversarial_problem(5)
    assert np.array_equal(l1[1]["obst"], l2[1]["obst"])
