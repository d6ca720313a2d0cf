import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """skip reasons go into the summary even with -q: whether the shapely / GEOS replay of the oracle
    (tests/test_shapely_replay.py) ever ran on a box is something a reader of the log must be able to see"""
    seen = {}
    for rep in terminalreporter.stats.get("skipped", []):
        reason = rep.longrepr[2] if isinstance(rep.longrepr, tuple) and len(rep.longrepr) == 3 else str(rep.longrepr)
        key = (rep.nodeid.split("::")[0], reason)
        seen[key] = seen.get(key, 0) + 1
    for (where, reason), n in sorted(seen.items()):
        if "no CUDA device in this container" in reason:
            continue
        terminalreporter.write_line("SKIPPED x%d  %s  --  %s" % (n, where, reason))
    try:
        import shapely  # noqa: F401
        terminalreporter.write_line("shapely importable: the GEOS replay of the oracle (tests/test_shapely_replay.py) RAN")
    except Exception:
        terminalreporter.write_line("shapely NOT importable here: tests/test_shapely_replay.py skipped, oracle stays unpinned against GEOS")
