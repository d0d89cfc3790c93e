#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (SASS rows with stall samples) with `nvdisasm -gi` line info of the
same cubin and aggregate instructions / stall samples per source line of a chosen function body.

usage: ncu_by_line.py <ncu_source.csv> <nvdisasm_gi.sass> <kernel-substring> <file-substring> <line_lo> <line_hi>
Each SASS instruction is attributed to the innermost frame of its inline chain that lies in
[line_lo, line_hi] of <file-substring> (e.g. the body of geo_tile)."""
import collections
import csv
import re
import sys


def main():
    src_csv, sass, kern, fsub, lo, hi = sys.argv[1:7]
    lo, hi = int(lo), int(hi)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    # parse the disassembly of the kernel's section
    insts = []
    chain = []
    active = False
    fresh = True
    for ln in open(sass):
        if ln.startswith("\t.section\t.text."):
            active = kern in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r"\s*/\*[0-9a-f]+\*/", ln):
            insts.append(list(chain))
            fresh = True
    if len(insts) != len(data):
        print(f"warning: {len(insts)} disassembled instructions vs {len(data)} ncu rows", file=sys.stderr)
    n = min(len(insts), len(data))
    agg = collections.defaultdict(lambda: collections.Counter())
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for i in range(n):
        key = None
        for f, l in insts[i]:
            if fsub in f and lo <= l <= hi:
                key = l
                break
        if key is None:
            key = -1
        r = data[i]
        a = agg[key]
        a["inst"] += int(r[ix["Instructions Executed"]])
        a["samples"] += int(r[ix["# Samples"]])
        a["static"] += 1
        for s in stall_cols:
            a[s] += int(r[ix[s]])
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["samples"] for a in agg.values())
    print(f"total warp-inst {ti}  samples {ts}")
    print(f"{'line':>6} {'inst%':>6} {'smp%':>6} {'static':>6}  top stalls")
    for k in sorted(agg):
        a = agg[k]
        st = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:4]
        print(f"{k:6d} {100*a['inst']/ti:6.2f} {100*a['samples']/ts:6.2f} {a['static']:6d}  " +
              " ".join(f"{n}:{100*v/max(a['samples'],1):.0f}%" for v, n in st if v))


if __name__ == "__main__":
    main()
