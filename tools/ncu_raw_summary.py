#!/usr/bin/env python
"""key metrics of an `ncu --page raw --csv` export, one row per captured launch

    python tools/ncu_raw_summary.py gpurun_out/r02f/c4_raw.csv [--json]"""
import csv
import json
import sys

KEYS = [("gpu__time_duration.sum", "us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__occupancy_limit_registers", "occ_reg", 1), ("launch__occupancy_limit_shared_mem", "occ_smem", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", 1),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaH%", 1),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fmacyc%", 1),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alucyc%", 1),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsuwf%", 1),
        ("smsp__inst_executed.sum", "winst", 1), ("dram__bytes_read.sum", "dramR_MB", 1e-6), ("dram__bytes_write.sum", "dramW_MB", 1e-6),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long", 1),
        ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short", 1),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar", 1),
        ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math", 1),
        ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait", 1),
        ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "st_nsel", 1),
        ("smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "st_disp", 1),
        ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio", 1),
        ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg", 1),
        ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "st_membar", 1),
        ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "st_noinst", 1),
        ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "st_branch", 1)]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name", "")[:60], "id": d.get("ID")}
        for k, name, sc in KEYS:
            if k in d:
                v = num(d[k])
                rec[name] = None if v is None else v * sc
        out.append(rec)
    return out, hdr, rows[1]


if __name__ == "__main__":
    recs, hdr, units = load(sys.argv[1])
    if "--json" in sys.argv:
        print(json.dumps(recs, indent=1))
    else:
        for r in recs:
            print(r["id"], r["kernel"])
            print("   " + "  ".join("%s=%s" % (k, ("%.4g" % v) if isinstance(v, float) else v) for k, v in r.items() if k not in ("kernel", "id")))
