#!/usr/bin/env python3
"""Executed instructions (or stall samples) of one kernel grouped by the OUTERMOST source line of a given
file -- i.e. inlined helpers are charged to the kernel statement that called them -- and optionally summed
over line ranges ("phases").

usage: ncu_phase.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING FILE_SUBSTRING [column] [name=lo-hi ...]
e.g.   ncu_phase.py prof.ncu-rep engine.cubin recon_inter_kernel recon_inter.cuh "Instructions Executed" stage=336-352 bucket=353-400
"""
import os
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kname, fsub = sys.argv[1:5]
    col = sys.argv[5] if len(sys.argv) > 5 else "Instructions Executed"
    phases = []
    for a in sys.argv[6:]:
        n, r = a.split("=")
        lo, hi = r.split("-")
        phases.append((n, int(lo), int(hi)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    counts, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur, hdr = r[1], None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if cur and kname in cur and hdr and len(r) > hdr.index(col):
            counts.append((r[1].strip(), int(r[hdr.index(col)])))
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, in_k, chain = [], False, []
    fresh = True
    for l in dis:
        if l.startswith("\t.section\t.text."):
            in_k = (os.environ.get("CUBIN_KERNEL") or kname) in l
            continue
        if l.startswith("\t.section"):
            in_k = False
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            fresh = True
            outer = [c for c in chain if fsub in c[0]]
            lines.append((outer[-1] if outer else (chain[-1] if chain else ("?", 0)), m.group(1)))
    if len(lines) != len(counts):
        print(f"warning: {len(lines)} disassembled vs {len(counts)} profiled instructions", file=sys.stderr)
    per = collections.Counter()
    tot = 0
    for (loc, _), (_, n) in zip(lines, counts):
        per[loc] += n
        tot += n
    print(f"total {col}: {tot}")
    if phases:
        rest = tot
        for name, lo, hi in phases:
            s = sum(n for (f, ln), n in per.items() if fsub in f and lo <= ln <= hi)
            rest -= s
            print(f"{100 * s / tot:6.2f}%  {s:12d}  {name} ({lo}-{hi})")
        print(f"{100 * rest / tot:6.2f}%  {rest:12d}  (other)")
    else:
        for (f, ln), n in sorted(per.items()):
            print(f"{100 * n / tot:6.2f}%  {n:12d}  {f}:{ln}")


if __name__ == "__main__":
    main()
