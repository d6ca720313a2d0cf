#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of an ncu source-page CSV to source lines of the kernel body.

    cuobjdump -xelf all libmpcheck.so; nvdisasm --print-line-info-inline X.cubin > all.sass
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_by_line.py all.sass src.csv '<mangled kernel name>' [--regions a-b:name,...]

Every instruction is charged to the OUTERMOST line of its inline chain that lies inside the kernel body (so helper code
is charged to its call site in the kernel).  --range=a-b gives the body's line range (frames outside it -- the
__global__ wrapper that inlines the body -- are skipped), --exclude=n,m lines that are mere call sites of the body's
own top-level lambda."""
import collections
import csv
import re
import sys


def main():
    sass, src, kern = sys.argv[1:4]
    regions = []
    lo_hi, exclude = None, set()
    for a in sys.argv[4:]:
        if a.startswith('--range='):
            lo_hi = tuple(int(v) for v in a[len('--range='):].split('-'))
        if a.startswith('--exclude='):
            exclude = set(int(v) for v in a[len('--exclude='):].split(',') if v)
    for a in sys.argv[4:]:
        if a.startswith('--regions='):
            for item in a[len('--regions='):].split(','):
                rng, name = item.split(':')
                lo, hi = rng.split('-')
                regions.append((int(lo), int(hi), name))
    lines = open(sass).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':'))
    chain, inst = [], {}
    pending = []
    for l in lines[start + 1:]:
        if l.startswith('\t.section'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            pending.append((int(m.group(2)), int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            if pending:
                chain = pending
                pending = []
            frames = ([chain[0][0]] + [c[1] for c in chain if c[1] is not None]) if chain else [0]
            outer = frames[-1]
            if lo_hi is not None:       # the last entry of the chain is the outermost frame: walk inwards to the body
                for f in reversed(frames):
                    if lo_hi[0] <= f <= lo_hi[1] and f not in exclude:
                        outer = f
                        break
            inst[int(m.group(1), 16)] = (outer, chain[0][0] if chain else 0, m.group(2))
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    base = int(data[0][ia], 16)
    by, sp = collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in data:
        off = int(r[ia], 16) - base
        n, s = int(r[ie]), int(r[isamp])
        tot += n
        tots += s
        ln = inst.get(off, (0, 0, ''))[0]
        by[ln] += n
        sp[ln] += s
    print("instructions executed %d, samples %d" % (tot, tots))
    if regions:
        for lo, hi, name in regions:
            n = sum(v for k, v in by.items() if lo <= k <= hi)
            s = sum(v for k, v in sp.items() if lo <= k <= hi)
            print("%-28s lines %4d-%4d  %5.1f %% of instructions  %5.1f %% of samples" % (name, lo, hi, 100.0 * n / tot, 100.0 * s / tots))
    else:
        for ln, n in by.most_common(60):
            print("%5.1f %% inst  %5.1f %% samples  line %d" % (100.0 * n / tot, 100.0 * sp[ln] / tots, ln))


if __name__ == '__main__':
    main()
