#!/usr/bin/env python3
"""Per-source-line executed-instruction totals of one kernel from an ncu report.

ncu's CSV export of the source page carries metrics only in the SASS view, so this joins
  ncu -i REP --page source --csv --print-source sass     (per-instruction counts)
with
  nvdisasm --print-line-info CUBIN                        (instruction -> file:line)
by instruction order inside the kernel's .text section.

usage: ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [top_n] [column]
(column: "Instructions Executed" by default; "# Samples" gives the warp-stall samples = where the time goes)
"""
import os
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    col = sys.argv[5] if len(sys.argv) > 5 else "Instructions Executed"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    counts, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            hdr = None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if cur and kname in cur and hdr and len(r) > hdr.index(col):
            counts.append((r[1].strip(), int(r[hdr.index(col)])))
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, in_k, loc = [], False, ("?", 0)
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
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            lines.append((loc, m.group(1)))
    if len(lines) != len(counts):
        print(f"warning: {len(lines)} disassembled vs {len(counts)} profiled instructions", file=sys.stderr)
    per = collections.Counter()
    tot = 0
    for (loc, _), (_, n) in zip(lines, counts):
        per[loc] += n
        tot += n
    print(f"total {col}: {tot}")
    for loc, n in per.most_common(top):
        print(f"{100 * n / tot:6.2f}%  {n:12d}  {loc[0]}:{loc[1]}")


if __name__ == "__main__":
    main()
