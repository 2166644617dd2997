#!/usr/bin/env python3
"""Print SASS with the scheduling control fields decoded (Volta+ 128-bit encoding):
stall count, yield, write-barrier index, read-barrier index, wait-barrier mask.
usage: sass_ctrl.py CUBIN KERNEL_SUBSTRING [start_hex end_hex]"""
import re
import subprocess
import sys

cubin, kname = sys.argv[1:3]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout.splitlines()
on = False
i = 0
while i < len(out):
    l = out[i]
    if "Function :" in l:
        on = kname in l
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if on and m and i + 1 < len(out):
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
        addr = int(m.group(1), 16)
        if m2 and lo <= addr <= hi:
            hw = int(m2.group(1), 16)
            stall = (hw >> 41) & 0xf
            yld = (hw >> 45) & 1
            wr = (hw >> 46) & 7
            rd = (hw >> 49) & 7
            wait = (hw >> 52) & 0x3f
            ctl = f"st{stall:2d} {'Y' if yld == 0 else ' '} W{wr if wr != 7 else '-'} R{rd if rd != 7 else '-'} wait{wait:06b}"
            print(f"{addr:05x} {ctl}  {m.group(2)[:90]}")
        i += 2
        continue
    i += 1
