"""GPU: closed-GOP splitter + ordered merge (include/p264b200_host.h, p264b200_gopdec_*).  One stream is cut at its
IDR pictures, the GOPs are decoded concurrently on different lanes of one engine and the pictures must come back in
stream order, byte-identical to the reference decoder's serial decode (decoder/decoder.c:43-64,745-806)."""
import hashlib

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lanes", [2, 1])
def test_f26_split_over_lanes_in_order(lanes):
    path = O.f26_path()
    if path is None:
        pytest.skip("oracle/_ref/f26.264 not present")
    golden = O.f26_frame_md5s()
    data = np.fromfile(path, dtype=np.uint8)
    assert [n for _, n in P.gop_scan(data)] == [250, 50]
    n = 0
    for pic, w, h in P.decode_annexb_gops(data, lanes=lanes):
        assert (w, h) == (352, 288)
        assert hashlib.md5(pic.tobytes()).hexdigest() == golden[n], f"picture {n} (stream order) with {lanes} lane(s)"
        n += 1
    assert n == 300


def test_written_1080p_stream_with_intra_period_over_four_lanes():
    """a synthetic 1080p stream with an IDR every 3 pictures (written by csrc/host/writer.cc as a real CAVLC stream):
    5 closed GOPs over 4 lanes -- lane 0 decodes GOPs 0 and 4 back to back -- against the serial drop-in decode"""
    mb_w, mb_h, n_pic = 120, 68, 14
    syn = P.Synth(mb_w, mb_h, n_refs=1, seed=77, first_intra=1, intra_period=3, sub8x8=0, max_level=3, coded_pct=20, qp_min=28, qp_max=28, qp_step=0)
    wr = P.Writer(mb_w, mb_h)
    for _ in range(n_pic):
        wr.put(syn.next_syntax())
    stream = np.frombuffer(wr.data(), dtype=np.uint8).copy()
    gops = P.gop_scan(stream)
    assert [n for _, n in gops] == [3, 3, 3, 3, 2]
    serial = [np.concatenate([p.ravel() for p in yuv]) for yuv in P.decode_annexb(stream)]
    assert len(serial) == n_pic
    got = [pic for pic, w, h in P.decode_annexb_gops(stream, lanes=4)]
    assert len(got) == n_pic
    for i, (a, b) in enumerate(zip(got, serial)):
        assert np.array_equal(a, b), f"picture {i}"
