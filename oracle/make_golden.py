#!/usr/bin/env python3
"""Regenerates tests/golden/f26_frames.md5 (per-frame md5 of the reference decoder's I420 output)
by running oracle/_ref/p264dec_ref -- the UNMODIFIED reference CLI built by oracle/Makefile -- on
bin/f26.264.  Only runs where /root/reference is mounted; the md5 list is committed."""
import hashlib
import subprocess
import sys
import tempfile
from pathlib import Path

here = Path(__file__).resolve().parent
ref = here / "_ref"
W, H = 352, 288
with tempfile.TemporaryDirectory() as td:
    yuv = Path(td) / "f26.yuv"
    subprocess.check_call([str(ref / "p264dec_ref"), "-d", str(ref / "f26.264"), str(yuv)], stderr=subprocess.DEVNULL)
    data = yuv.read_bytes()
fs = W * H * 3 // 2
assert len(data) % fs == 0
whole = hashlib.md5(data).hexdigest()
lines = [hashlib.md5(data[i : i + fs]).hexdigest() for i in range(0, len(data), fs)]
out = here.parent / "tests" / "golden" / "f26_frames.md5"
out.write_text("\n".join(lines) + "\n")
(here.parent / "tests" / "golden" / "f26_yuv.md5").write_text(whole + "\n")
print(f"{len(lines)} frames, whole-file md5 {whole}", file=sys.stderr)
