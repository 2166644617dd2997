#!/usr/bin/env python3
"""Per-step cycle trace of one deblock CTA (needs a GPU and a library built with `make -C p264decoder_b200/csrc TRACE=1`).

usage: P264B200_TRACE=<ticket> python tools/dbf_trace.py [lanes]
Runs a few 1080p steps, then prints for every warp of the traced CTA the mean cycles per lockstep step spent
(a) in the vertical-edge pass and the hand-off to the row below, (b) waiting for the rows above, (c) in the transpose +
horizontal-edge pass, and the mean period per macroblock.
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
    print(f"ticket {os.environ.get('P264B200_TRACE')}: mean cycles per macroblock (marks: 0 top of the iteration, 1 before waiting for the rows above, 2 after the wait, 3 end)")
    for w in range(8):
        m = [buf[w, 2:mb_w - 2, k] for k in range(4)]
        nxt0 = buf[w, 3:mb_w - 1, 0]
        print(f"warp {w}: V-phase+hand-off {np.mean(m[1] - m[0]):6.0f}  wait-top {np.mean(m[2] - m[1]):6.0f}  H-phase {np.mean(m[3] - m[2]):6.0f}"
              f"  period {np.mean(nxt0 - m[0]):6.0f}   start {int(buf[w, 0, 0] - buf[0, 0, 0]):8d}  end {int(buf[w, mb_w - 1, 3] - buf[0, 0, 0]):8d}")
    if os.environ.get("DBF_DETAIL"):
        x0 = int(os.environ["DBF_DETAIL"])
        base = buf[0, x0, 0]
        print(f"timeline (cycles since warp 0 started macroblock {x0}): per warp, per macroblock: top / wait-start / wait-end / end / side info arrived / rows arrived")
        for w in range(8):
            print(f" warp {w}: " + "  ".join(f"x={x}:" + "/".join(str(int(buf[w, x, k] - base)) for k in range(6)) for x in range(x0 - 2, x0 + 6)))
    lib.p264b200_debug_cta_times.restype = C.c_int
    lib.p264b200_debug_cta_times.argtypes = [C.c_void_p, C.c_size_t]
    ct = np.zeros((2048, 4), dtype=np.int64)
    assert lib.p264b200_debug_cta_times(ct.ctypes.data, ct.nbytes) == 0
    groups = (mb_h + 7) // 8
    t00 = ct[:2 * groups * ((lanes + 3) // 4), 0].min()
    print("per CTA of quad 0 (us since the first CTA entered): ticket role grp  entry  first-MB  last-MB  exit")
    for tk in range(2 * groups):
        e = (ct[tk] - t00) / 1000.0
        print(f"  {tk:3d} {'luma' if tk % 2 == 0 else 'chro'} {tk // 2:2d}   {e[0]:8.1f} {e[1]:8.1f} {e[2]:8.1f} {e[3]:8.1f}")


if __name__ == "__main__":
    main()
