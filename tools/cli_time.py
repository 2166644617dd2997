#!/usr/bin/env python3
"""Wall time of the two command-line decoders on bin/f26.264 (BASELINE.json configs[0] / [1]): the reference's own
CLI built in oracle/_ref and tools/p264dec_b200.c on the drop-in library.  Needs a GPU for the second."""
import hashlib
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
src = REF / "f26.264"


def run(exe, out, n=3):
    best = None
    for _ in range(n):
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), "-d", str(globals()["src"]), out], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode:
            sys.exit(f"{exe} failed: {r.stderr[-400:]}")
        best = dt if best is None else min(best, dt)
    return best, hashlib.md5(Path(out).read_bytes()).hexdigest()


t_ref, m_ref = run(REF / "p264dec_ref", "/tmp/f26_ref.yuv")
t_gpu, m_gpu = run(ROOT / "p264decoder_b200" / "lib" / "p264dec_b200", "/tmp/f26_b200.yuv")
print(f"reference CLI : {t_ref:.3f} s  ({300 / t_ref:.0f} CIF frames/s)  md5 {m_ref}")
print(f"B200 CLI      : {t_gpu:.3f} s  ({300 / t_gpu:.0f} CIF frames/s, incl. process start + CUDA context)  md5 {m_gpu}")
print("byte-identical" if m_ref == m_gpu else "DIFFERENT")

# N copies of the stream at once: the reference = N processes of its CLI on the host cores (no output file),
# the B200 build = one p264dec_multi process, one engine lane per stream
import os
cores = os.cpu_count() or 1
for n in (16, 64, 256):
    t0 = time.perf_counter()
    procs = [subprocess.Popen([str(REF / "p264dec_ref"), "-d", str(src)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(min(n, 64))]
    for p in procs:
        p.wait()
    t_ref_n = (time.perf_counter() - t0) * (n / min(n, 64))
    r = subprocess.run([str(ROOT / "p264decoder_b200" / "lib" / "p264dec_multi"), "-n", str(n), str(src)], capture_output=True, text=True)
    fps = [l for l in r.stderr.splitlines() if "decoding speed" in l]
    steady = [l for l in r.stderr.splitlines() if "after the first step" in l]
    print(f"{n:3d} streams: reference CLI x {n} on {cores} cores {300 * n / t_ref_n:8.0f} frames/s   p264dec_multi {fps[0].split(':')[1].strip() if fps else r.stderr[-200:]}"
          f" whole process, {steady[0].split(':')[1].strip() if steady else '?'} after CUDA context + engine creation")


# the same on a synthetic 1080p stream (BASELINE.json configs[2] as a real bitstream, tools/make_stream.py)
sys.path.insert(0, str(ROOT / "tools"))
import make_stream  # noqa: E402

src = Path("/tmp/synth1080.264")
n_pic = 48
make_stream.make(src, "1080p", n_pic)
t_ref, m_ref = run(REF / "p264dec_ref", "/tmp/s1080_ref.yuv", n=2)
t_gpu, m_gpu = run(ROOT / "p264decoder_b200" / "lib" / "p264dec_b200", "/tmp/s1080_b200.yuv", n=2)
print(f"1080p synthetic stream, {n_pic} pictures: reference CLI {n_pic / t_ref:.1f} pictures/s, B200 CLI {n_pic / t_gpu:.1f} pictures/s whole process; "
      + ("byte-identical" if m_ref == m_gpu else "DIFFERENT"))
for n in (16, 64):
    t0 = time.perf_counter()
    procs = [subprocess.Popen([str(REF / "p264dec_ref"), "-d", str(src)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(min(n, cores))]
    for p in procs:
        p.wait()
    t_ref_n = (time.perf_counter() - t0) * (n / min(n, cores))
    r = subprocess.run([str(ROOT / "p264decoder_b200" / "lib" / "p264dec_multi"), "-n", str(n), str(src)], capture_output=True, text=True)
    steady = [l for l in r.stderr.splitlines() if "after the first step" in l]
    print(f"{n:3d} x 1080p streams: reference CLI on {cores} cores {n_pic * n / t_ref_n:7.1f} pictures/s   p264dec_multi "
          f"{steady[0].split(':')[1].strip() if steady else r.stderr[-200:]} after CUDA context + engine creation")
