#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (no GPU needed).

usage: sass_lines.py CUBIN KERNEL_SUBSTRING [file_substring]
Joins `nvdisasm --print-line-info` line markers with the instructions that follow them; with
inlining every instruction is attributed to the innermost source line.  Useful for sizing the
class bodies of recon_inter before spending GPU time (dynamic counts need tools/ncu_lines.py)."""
import collections
import re
import subprocess
import sys


def main():
    cubin, kname = sys.argv[1:3]
    fsub = sys.argv[3] if len(sys.argv) > 3 else ""
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
    in_k, loc = False, ("?", 0)
    per = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    total = 0
    for l in dis:
        if l.startswith("\t.section\t.text."):
            in_k = kname in l
            continue
        if l.startswith("\t.section"):
            in_k = False
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            per[loc] += 1
            ops[loc][m.group(2).split(".")[0]] += 1
            total += 1
    print("total", total)
    for (f, ln), n in sorted(per.items()):
        if fsub in f:
            top = " ".join(f"{k}:{v}" for k, v in ops[(f, ln)].most_common(5))
            print(f"{f}:{ln}\t{n}\t{top}")


if __name__ == "__main__":
    main()
