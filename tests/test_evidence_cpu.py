"""Evidence hygiene: what the documents and bench.py quote must exist and must belong to the
code in the tree.

  * profiles/r02_ncu_summary.json was captured from the device code that is in the tree now (bench.py reports
    `roofline.ncu.stale` otherwise) and carries the counts bench.py reads;
  * every profiles/ file named in DESIGN.md, README.md, INTEGRATION.md and profiles/README.md (round-2 names) exists;
  * the SASS extract shows the instructions the design claims (TMA bulk copies, mbarriers, packed float32x2
    arithmetic, three-input min / max, programmatic-dependent-launch waits) for the headline kernel;
  * both arms of bench.py describe the same workload (`config` equal), and the committed bench record says so."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_ncu_summary_belongs_to_the_device_code_in_the_tree():
    import bench
    s = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")))
    assert s["kernel_source_hash"] == bench.kernel_source_hash(), \
        "device code changed after the capture: re-run tools/capture_r02.sh + tools/profiles_from_capture.sh"
    assert s["headline"]["warp_instructions_per_batch"] > 1e7 and s["headline"]["dram_bytes_per_batch"] > 1e6
    assert s["full_evaluation"]["dram_bytes_per_launch"] > 1e7
    kinds = {l["kernel"].split("<")[0] for l in s["headline"]["launches"]}
    assert {"k_check", "k_refine"} <= kinds


def test_profiles_named_in_the_documents_exist():
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"`(?:profiles/)?(r02[a-z]?_[A-Za-z0-9_.]+\.(?:json|csv|txt))`", text)):
            if "<" in name or "*" in name:
                continue
            if not os.path.exists(os.path.join(ROOT, "profiles", name)):
                missing.append((doc, name))
    assert not missing, missing


def test_sass_extract_shows_what_the_design_claims():
    text = open(os.path.join(ROOT, "profiles", "r02_sass_extract.txt")).read()
    m = re.search(r"k_check<4, 4, false, false, false>\n\s+(REG:\d+ STACK:\d+[^\n]*)\n\s+([^\n]+)", text)
    assert m, "headline instantiation missing from the extract"
    counts = dict(re.findall(r"([A-Z0-9]+) (\d+)", m.group(2)))
    for op in ("UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMNMX3", "ACQBULK"):
        assert int(counts.get(op, 0)) > 0, op
    assert "REG:64" in m.group(1)
    assert "k_check_full<4, 4, false>" in text and "k_refine" in text and "k_clear" in text


def test_both_bench_arms_describe_the_same_workload():
    import argparse
    import bench
    args = argparse.Namespace(batches_per_step=100)
    assert bench.c4_config(args, 64) == bench.c4_config(args, 64)
    ours = json.load(open(os.path.join(ROOT, "profiles", "r02z_bench.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02z_bench_reference.json")))
    assert ours["config"] == ref["config"] and ref["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in ours, key
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(ours["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(ours["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(ours["cpu_baseline"])
    assert ours["roofline"]["ncu"]["stale"] is False
    assert ours["latency"]["planner_query_p50_us"] > 0 and ours["c5"]["disagreements_vs_oracle_rank0"] == 0
    assert set(ours["regimes"]) >= {"c4_dense", "mixed", "cull_only", "full_evaluation"}
