#!/usr/bin/env python3
"""Static SASS mnemonic histogram per kernel of an object / cubin (evidence for the instruction choices:
packed s16x2 min/max/add, VABSDIFF4, IDP.4A, PRMT, SYNCS mbarrier ops, vector LDG/STG widths ...).
usage: sass_hist.py OBJ [top_n]"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 28
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
hist, name = {}, None
for l in out:
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m and name:
        hist[name][m.group(1)] += 1
for k, c in hist.items():
    n = sum(c.values())
    print(f"== {k}: {n} instructions")
    print("   " + "  ".join(f"{op} {v}" for op, v in c.most_common(top)))
