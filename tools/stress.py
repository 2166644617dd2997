#!/usr/bin/env python3
"""Randomised parity stress (needs a GPU): random geometries, lane counts and stream options for a wall-clock budget,
every reconstructed picture compared byte for byte with the CPU oracle.  Aimed at the schedule-dependent parts
(deblock mbarrier rings + in/out warps, intra run wavefront, tickets):  stress.py [seconds] [seed] [big]
(`big`: up to 130x70 macroblocks and 48 lanes -- more deblock CTAs than fit on the machine at once)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import p264decoder_b200 as P  # noqa: E402
import _oracle as O  # noqa: E402


def run(budget=60.0, seed=1, big=False):
    """returns (configurations, pictures) checked inside the wall-clock budget; raises AssertionError on the first mismatch"""
    rng = np.random.default_rng(seed)
    t0, cases, pictures = time.time(), 0, 0
    while time.time() - t0 < budget:
        mb_w, mb_h = (int(rng.integers(40, 131)), int(rng.integers(20, 71))) if big else (int(rng.integers(1, 40)), int(rng.integers(1, 40)))
        lanes = int(rng.choice([17, 24, 33, 48])) if big else int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 13, 16]))
        n_refs = int(rng.choice([1, 1, 2, 4]))
        kw = dict(n_refs=n_refs, seed=int(rng.integers(1, 1 << 30)), intra_pct=int(rng.choice([0, 0, 3, 10, 40, 100])),
                  sweep_offsets=int(rng.integers(0, 2)), coded_pct=int(rng.choice([0, 10, 25, 60])), skip_pct=int(rng.choice([0, 5, 50])),
                  sub8x8=int(rng.integers(0, 2)), mv_range=int(rng.choice([2, 16, 64])), confine_mv=int(rng.integers(0, 2)),
                  deblock=int(rng.choice([1, 1, 1, 0])), first_intra=int(rng.integers(0, 2)), qp_min=int(rng.choice([0, 20])), qp_max=int(rng.choice([40, 51])),
                  qp_step=3, chroma_qp_index_offset=int(rng.integers(-4, 5)))
        n_slots = n_refs + 1
        eng = P.Engine(mb_w, mb_h, n_slots=n_slots, lanes=lanes)
        syns = [P.Synth(mb_w, mb_h, **dict(kw, seed=kw["seed"] + 7 * l)) for l in range(lanes)]
        rings = [O.OracleFrames(mb_w, mb_h, n_slots) for _ in range(lanes)]
        if not kw["first_intra"]:
            for l in range(lanes):
                for s in range(n_slots):
                    pic = P.smooth_picture(16 * mb_w, 16 * mb_h, seed=l * 10 + s)
                    eng.upload(l, s, *pic)
                    rings[l].set(s, *pic)
        for i in range(2 if big else int(rng.integers(2, 6))):
            frames = [s.next() for s in syns]
            for l, fr in enumerate(frames):
                eng.stage(0, l, fr.syntax())
            eng.recon_step(0, lanes)
            eng.sync()
            for l, fr in enumerate(frames):
                want = rings[l].recon(fr)
                got = eng.download(l, fr.hdr.dst_slot)
                for name, g, w in zip("YUV", got, want):
                    if not np.array_equal(g, w):
                        raise AssertionError(f"MISMATCH: {mb_w}x{mb_h} MBs, {lanes} lanes, picture {i} lane {l} plane {name}, options {kw}")
                pictures += 1
        eng.close()
        cases += 1
    return cases, pictures


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    t0 = time.time()
    try:
        cases, pictures = run(budget, seed, big)
    except AssertionError as e:
        print(e)
        sys.exit(1)
    print(f"stress ok: {cases} random configurations, {pictures} pictures bit-exact in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
