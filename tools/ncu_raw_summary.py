#!/usr/bin/env python
"""Key metrics of `ncu --page raw --csv` exports, one record per captured launch; and the summary file bench.py reads.

    python tools/ncu_raw_summary.py gpurun_out/<tag>/c4_raw.csv              # table
    python tools/ncu_raw_summary.py --summary <tag> gpurun_out/<tag>         # -> profiles/r02_ncu_summary.json (stdout)

Units are taken from the export's unit row and normalised (us, bytes)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_limit_regs"),
        ("launch__occupancy_limit_shared_mem", "occ_limit_smem"), ("launch__waves_per_multiprocessor", "waves_per_sm"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_inst_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_inst_pct"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_cycles_pct"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "pipe_alu_cycles_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct"),
        ("smsp__inst_executed.sum", "warp_instructions"), ("dram__bytes_read.sum", "dram_read_bytes"),
        ("dram__bytes_write.sum", "dram_write_bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleeping")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", ""), "id": int(d.get("ID", -1))}
        for k, name in KEYS:
            if k in d:
                v = num(d[k])
                if v is not None and (name == "us" or name.endswith("_bytes")):
                    v *= SCALE.get(u[k], 1.0)
                rec[name] = v
        out.append(rec)
    return out


def short(k):
    return k.replace("mpc::", "").replace("(mpc::KParams)", "").replace("(KParams)", "").replace("void ", "")


def summary(tag, folder):
    import hashlib
    h = hashlib.sha256()
    for f in ("mpc_device.cuh", "mpc_exact.cuh", "mpc_expand.cuh"):      # the device code
        h.update(open(os.path.join(ROOT, "motionplanner_b200", "csrc", f), "rb").read())
    out = {"capture": tag, "kernel_source_hash": h.hexdigest()[:16],
           "how": "tools/capture_r02.sh under gpurun: ncu --set full --clock-control none on tools/ncu_target.py <regime> 2; "
                  "per-launch numbers are serialised and cold-cache (counts of instructions and DRAM bytes do not depend on that)"}
    for reg in ("c4", "mixed", "cull", "full", "planner", "planner_acc", "aroc"):
        p = os.path.join(folder, reg + "_raw.csv")
        if not os.path.exists(p):
            continue
        recs = load(p)
        for r in recs:
            r["kernel"] = short(r["kernel"])
        checks = [r for r in recs if r["kernel"].startswith("k_check")]
        block = {"launches": recs}
        if reg in ("c4", "mixed", "cull"):
            # one batch = the k_check launches up to (not including) the first k_refine
            first = []
            for r in recs:
                if r["kernel"].startswith("k_refine"):
                    break
                first.append(r)
            block["warp_instructions_per_batch"] = sum(r["warp_instructions"] for r in first)
            block["dram_bytes_per_batch"] = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in first)
            block["k_check_us_per_batch_under_ncu"] = sum(r["us"] for r in first)
        if reg == "full" and checks:
            block["dram_bytes_per_launch"] = checks[0]["dram_read_bytes"] + checks[0]["dram_write_bytes"]
            block["warp_instructions_per_launch"] = checks[0]["warp_instructions"]
        out["headline" if reg == "c4" else ("full_evaluation" if reg == "full" else reg)] = block
    return out


if __name__ == "__main__":
    if sys.argv[1] == "--summary":
        print(json.dumps(summary(sys.argv[2], sys.argv[3]), indent=1))
    else:
        for r in load(sys.argv[1]):
            print(r["id"], short(r["kernel"]))
            print("   " + "  ".join("%s=%s" % (k, ("%.4g" % v) if isinstance(v, float) else v) for k, v in r.items() if k not in ("kernel", "id")))
