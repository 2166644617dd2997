"""GPU: the multi-stream decoder (include/p264b200_host.h, p264b200_multi_*) -- several copies of
bin/f26.264, cut to different lengths, decoded concurrently on one engine; every picture of every
stream must carry the md5 the reference's own C decoder produced for that frame (tests/golden)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = pytest.mark.gpu


class MultiCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_streams", C.c_int32), ("n_threads", C.c_int32), ("reserved", C.c_int32 * 5)]


def _nal_cuts(data):
    """byte offsets of the start codes (where a stream may be cut without splitting a NAL unit)"""
    cuts, i = [], 0
    while True:
        i = data.find(b"\x00\x00\x00\x01", i)
        if i < 0:
            return cuts
        cuts.append(i)
        i += 4


@pytest.mark.parametrize("threads", [1, 4])
def test_multi_stream_f26_bit_exact(threads):
    path = O.f26_path()
    if not path.exists():
        pytest.skip("oracle/_ref/f26.264 not present")
    golden = O.f26_frame_md5s()
    data = path.read_bytes()
    cuts = _nal_cuts(data)
    # 5 streams: the whole file, and prefixes ending right before NAL 40, 90, 150, 260 (the last one crosses
    # the second IDR at frame 250): lanes end at different steps and stay idle afterwards
    ends = [len(data)] + [cuts[k] for k in (40, 90, 150, 260)]
    bufs = [C.create_string_buffer(data[:e], e) for e in ends]
    lib = P.load_library()
    lib.p264b200_multi_open.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MultiCfg)]
    lib.p264b200_multi_set_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.p264b200_multi_step.argtypes = [C.c_void_p, C.c_void_p]
    lib.p264b200_multi_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.p264b200_multi_picture.restype = C.c_void_p
    lib.p264b200_multi_close.argtypes = [C.c_void_p]
    cfg = MultiCfg(device=0, n_streams=len(bufs), n_threads=threads)
    m = C.c_void_p()
    assert lib.p264b200_multi_open(C.byref(m), C.byref(cfg)) == 0
    for s, b in enumerate(bufs):
        assert lib.p264b200_multi_set_stream(m, s, b, len(b)) == 0
    counts = [0] * len(bufs)
    produced = (C.c_uint8 * len(bufs))()
    total = 0
    while True:
        r = lib.p264b200_multi_step(m, produced)
        assert r >= 0, r
        if r == 0:
            break
        total += r
        assert r == sum(produced)
        for s in range(len(bufs)):
            w, h = C.c_int(), C.c_int()
            p = lib.p264b200_multi_picture(m, s, C.byref(w), C.byref(h))
            if not produced[s]:
                assert not p
                continue
            assert (w.value, h.value) == (352, 288)
            pic = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(352 * 288 * 3 // 2,))
            assert hashlib.md5(pic.tobytes()).hexdigest() == golden[counts[s]], f"stream {s} frame {counts[s]}"
            counts[s] += 1
    lib.p264b200_multi_close(m)
    assert counts[0] == 300 and total == sum(counts)
    assert counts[1] < counts[2] < counts[3] < counts[4] < 300 and counts[4] > 250


def test_multi_stream_mixed_content():
    """Three DIFFERENT streams of the same coded size in one batch: bin/f26.264 and two written synthetic CIF streams
    (other QPs, deblock offsets, intra macroblocks); every picture against its own reference."""
    from test_bitstream_writer import make_stream

    path = O.f26_path()
    if not path.exists():
        pytest.skip("oracle/_ref/f26.264 not present")
    golden = O.f26_frame_md5s()
    f26 = path.read_bytes()
    s1, want1 = make_stream(22, 18, 40, seed=501, intra_pct=6, sweep_offsets=1)
    s2, want2 = make_stream(22, 18, 25, seed=502, intra_pct=0, chroma_qp_index_offset=2)
    streams = [f26, s1, s2]
    bufs = [C.create_string_buffer(d, len(d)) for d in streams]
    lib = P.load_library()
    lib.p264b200_multi_open.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MultiCfg)]
    lib.p264b200_multi_set_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.p264b200_multi_step.argtypes = [C.c_void_p, C.c_void_p]
    lib.p264b200_multi_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.p264b200_multi_picture.restype = C.c_void_p
    lib.p264b200_multi_close.argtypes = [C.c_void_p]
    cfg = MultiCfg(device=0, n_streams=3, n_threads=2)
    m = C.c_void_p()
    assert lib.p264b200_multi_open(C.byref(m), C.byref(cfg)) == 0
    for s, b in enumerate(bufs):
        assert lib.p264b200_multi_set_stream(m, s, b, len(streams[s])) == 0
    counts = [0, 0, 0]
    produced = (C.c_uint8 * 3)()
    fsz = 352 * 288 * 3 // 2
    while lib.p264b200_multi_step(m, produced) > 0:
        for s in range(3):
            if not produced[s]:
                continue
            p = lib.p264b200_multi_picture(m, s, None, None)
            pic = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(fsz,)).tobytes()
            if s == 0:
                assert hashlib.md5(pic).hexdigest() == golden[counts[0]], f"f26 picture {counts[0]}"
            else:
                assert pic == (want1 if s == 1 else want2)[counts[s]], f"stream {s} picture {counts[s]}"
            counts[s] += 1
    lib.p264b200_multi_close(m)
    assert counts == [300, 40, 25]
