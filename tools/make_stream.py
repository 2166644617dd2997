#!/usr/bin/env python3
"""Write a synthetic H.264 stream (BASELINE.json configs[2], bitstream variant) that the unmodified reference decoder
and this repo's decoders both decode:  make_stream.py OUT.264 [--size 1080p|4k|cif] [--pictures N] [--seed S]
(csrc/host/synth.cc -> csrc/host/writer.cc; no GPU needed)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import p264decoder_b200 as P  # noqa: E402

SIZES = {"1080p": (120, 68), "4k": (240, 135), "cif": (22, 18)}


def make(path, size="1080p", pictures=48, seed=264):
    mb_w, mb_h = SIZES[size]
    syn = P.Synth(mb_w, mb_h, n_refs=1, seed=seed, sub8x8=0, first_intra=1, intra_period=6, confine_mv=1, qp_min=22, qp_max=34, qp_step=2,
                  max_level=5, coded_pct=25, mv_range=16, skip_pct=5, intra_pct=3)
    wr = P.Writer(mb_w, mb_h)
    for _ in range(pictures):
        wr.put(syn.next_syntax())
    data = wr.data()
    Path(path).write_bytes(data)
    return len(data)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--size", default="1080p", choices=list(SIZES))
    ap.add_argument("--pictures", type=int, default=48)
    ap.add_argument("--seed", type=int, default=264)
    a = ap.parse_args()
    n = make(a.out, a.size, a.pictures, a.seed)
    print(f"{a.out}: {a.pictures} pictures, {n} bytes")
