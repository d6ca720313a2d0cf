#!/usr/bin/env python
"""by-region attribution of the captured k_check launches (instructions executed and stall samples per part of the
kernel source), from the gzipped `--page source` exports of tools/capture_r02.sh.  The regions are found by marker
comments in csrc/mpc_device.cuh, so the file follows the source.

    python tools/ncu_regions.py <tag> <all.sass from nvdisasm --print-line-info-inline>"""
import gzip
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "motionplanner_b200", "csrc", "mpc_device.cuh")
MARKS = [("prologue: CTA identity, query record", "__device__ __forceinline__ void check_body|__global__ void __launch_bounds__(K_THREADS, K_MIN_CTAS) k_check(const"),
         ("slot set-up: guards, step arrays, zeroing", "auto run_slot = [&]() {"),
         ("early return / compaction pre-pass", "const bool cmode ="),
         ("tile fetch (TMA) + wait", "for (int ph = 0; ph < nphase; ph++) {"),
         ("group set-up: candidate, library record, done test", "int nxt_prim = 0;"),
         ("polygon transform + park in shared memory", "float ax[VA], ay[VA], ln["),
         ("per-group state, queue lambdas", "const float acx = active ?"),
         ("obstacles: sub-block boxes, centre-line axis, full-test drains", "// ---- obstacles of this time step"),
         ("boundary segments: parity, near test, drains", "// ---- boundary segments of the free space at this time step"),
         ("fold + publish", "// ---- fold this phase into the group state"),
         ("epilogue: statistics, tail", "    };      // run_slot")]


def regions():
    lines = open(SRC).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "k_check(const" in l and "__global__" in l and "K_MIN_CTAS" in l and "full" not in l) + 1
    out = []
    pos = start
    for name, pat in MARKS:
        alts = pat.split("|")
        i = next((k for k in range(pos - 1, len(lines)) if any(a in lines[k] for a in alts)), None)
        if i is None:
            continue
        out.append([i + 1, None, name])
        pos = i + 1
    end = next(k for k in range(pos, len(lines)) if lines[k].startswith("}")) + 1
    for k in range(len(out)):
        out[k][1] = (out[k + 1][0] - 1) if k + 1 < len(out) else end
    return out


def main():
    tag, sass = sys.argv[1], sys.argv[2]
    regs = regions()
    arg = "--regions=" + ",".join("%d-%d:%s" % (a, b, n.replace(",", ";").replace(":", " ")) for a, b, n in regs)
    src_lines = open(SRC).read().split("\n")
    call_sites = [i + 1 for i, l in enumerate(src_lines) if l.strip() == "run_slot();"]
    rng = "--range=%d-%d" % (regs[0][0], regs[-1][1])
    exc = "--exclude=" + ",".join(str(c) for c in call_sites)
    folder = os.path.join(ROOT, "gpurun_out", tag)
    names = subprocess.run("grep -o '^.text._ZN3mpc7k_checkI[A-Za-z0-9_]*' %s" % sass, shell=True, capture_output=True, text=True).stdout.split()
    for reg in ("c4", "mixed", "planner", "planner_acc", "aroc"):
        p = os.path.join(folder, reg + "_source.csv.gz")
        if not os.path.exists(p):
            continue
        text = gzip.open(p, "rt").read().split("\n")
        starts = [i for i, l in enumerate(text) if l.startswith('"Kernel Name"')] + [len(text)]
        seen = set()
        for k in range(len(starts) - 1):
            head = text[starts[k]]
            if "k_check" not in head or "k_check_full" in head:
                continue
            body = text[starts[k]:starts[k + 1]]
            key = (head, sum(len(l) for l in body[2:12]))
            if key in seen:
                continue
            seen.add(key)
            m = re.search(r"k_check<\(int\)(\d+), \(int\)(\d+), \(bool\)(\d), \(bool\)(\d), \(bool\)(\d)>", head)
            if not m:
                continue
            mangled = "_ZN3mpc7k_checkILi%sELi%sELb%sELb%sELb%sEEEvNS_7KParamsE" % m.groups()
            with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
                f.write("\n".join(body) + "\n")
            print("== %s: launch %d of the capture, k_check<%s,%s,%s,%s,%s>" % ((reg, len(seen)) + m.groups()))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), sass, f.name, mangled, arg, rng, exc],
                               capture_output=True, text=True)
            print(r.stdout if r.returncode == 0 else r.stderr[-400:])
            os.unlink(f.name)


if __name__ == "__main__":
    main()
