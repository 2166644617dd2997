#!/usr/bin/env python3
"""Per-step cycle trace of one deblock CTA (needs a GPU).

usage: P264B200_TRACE=<ticket> python tools/dbf_trace.py [lanes]
Runs a few 1080p steps, then prints for every warp of the traced CTA the mean cycles per lockstep step spent
(a) waiting in the barrier, (b) in the vertical-edge pass, (c) in the transpose + horizontal-edge pass,
(d) in the tail (read-back, stores), and the mean step period.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import p264decoder_b200 as P  # noqa: E402


def main():
    lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mb_w, mb_h = 120, 68
    eng = P.Engine(mb_w, mb_h, n_slots=2, lanes=lanes, stage_steps=2)
    gens = [P.Synth(mb_w, mb_h, seed=100 + l, first_intra=0, intra_pct=0) for l in range(lanes)]
    keep = []
    for step in range(2):
        for l, gsyn in enumerate(gens):
            f = gsyn.next()
            keep.append(f)
            eng.stage(step, l, f.syntax())
    for it in range(4):
        eng.recon_step(it & 1)
    eng.sync()
    lib = P.load_library()
    lib.p264b200_debug_trace.restype = C.c_int
    lib.p264b200_debug_trace.argtypes = [C.c_void_p, C.c_size_t]
    buf = np.zeros((9, 320, 6), dtype=np.int64)
    rc = lib.p264b200_debug_trace(buf.ctypes.data, buf.nbytes)
    assert rc == 0, rc
    n_steps = mb_w + 7
    print(f"ticket {os.environ.get('P264B200_TRACE')}: mean cycles per lockstep step (marks: 0 after barrier A, 1 before barrier B, 2 after it, 3 end of step)")
    t0 = buf[:, :n_steps, 0]
    print("first..last barrier exit (warp 0):", int(t0[0, n_steps - 1] - t0[0, 0]), " per step:", round(float(t0[0, n_steps - 1] - t0[0, 0]) / (n_steps - 1)))
    for w in range(9):
        lo, hi = (w, w + mb_w) if w < 8 else (0, n_steps)
        m = [buf[w, lo + 1:hi - 1, k] for k in range(4)]
        nxt0 = buf[w, lo + 2:hi, 0]
        nxt3 = buf[w, lo + 2:hi, 3]   # taken right before the NEXT step's barrier A
        print(f"{'warp %d' % w if w < 8 else 'I/O   '}: V-phase {np.mean(m[1] - m[0]):6.0f}  barrier B {np.mean(m[2] - m[1]):6.0f}  H-phase {np.mean(nxt3 - m[2]):6.0f}"
              f"  barrier A {np.mean(nxt0 - nxt3):6.0f}  period {np.mean(nxt0 - m[0]):6.0f}")
    lib.p264b200_debug_cta_times.restype = C.c_int
    lib.p264b200_debug_cta_times.argtypes = [C.c_void_p, C.c_size_t]
    ct = np.zeros((2048, 4), dtype=np.int64)
    assert lib.p264b200_debug_cta_times(ct.ctypes.data, ct.nbytes) == 0
    groups = (mb_h + 7) // 8
    t00 = ct[:2 * groups * ((lanes + 3) // 4), 0].min()
    print("per CTA of quad 0 (us since the first CTA entered): ticket role grp  entry  first-step  last-step  exit")
    for tk in range(2 * groups):
        e = (ct[tk] - t00) / 1000.0
        print(f"  {tk:3d} {'luma' if tk % 2 == 0 else 'chro'} {tk // 2:2d}   {e[0]:8.1f} {e[1]:8.1f} {e[2]:8.1f} {e[3]:8.1f}")


if __name__ == "__main__":
    main()
