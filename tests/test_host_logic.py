"""CPU: host-side logic -- CAVLC code tables against the standard's tables (golden parsed from the
reference's core/vlc.h), Annex-B splitting and emulation-prevention removal against the
reference's own p264_nal_decode, parser error behaviour, synthetic generator invariants."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

ROOT = Path(__file__).resolve().parents[1]


def test_cavlc_tables_match_standard():
    lib = P.load_library()
    gold = json.loads((ROOT / "tests" / "golden" / "cavlc_tables.json").read_text())

    def entry(kind, table, sym):
        l, b = C.c_int(), C.c_int()
        assert lib.p264b200_cavlc_table_entry(kind, table, sym, C.byref(l), C.byref(b)) == 0
        return [b.value, l.value]

    for t in range(4):
        for s in range(68):
            if (s & 3) <= (s >> 2):
                assert entry(0, t, s) == gold["coeff_token"][t][s], (t, s)
    for s in range(20):
        if (s & 3) <= (s >> 2):
            assert entry(1, 0, s) == gold["coeff_token"][4][s], s
    for t in range(15):
        for s in range(16 - t):
            assert entry(2, t, s) == gold["total_zeros"][t][s], (t, s)
    for t in range(3):
        for s in range(4 - t):
            assert entry(3, t, s) == gold["total_zeros_dc"][t][s], (t, s)
    for t in range(7):
        for s in range([2, 3, 4, 5, 6, 7, 15][t]):
            assert entry(4, t, s) == gold["run_before"][t][s], (t, s)


def test_annexb_split_and_unescape():
    raw = bytes([0, 0, 0, 1, 0x67, 1, 2, 0, 0, 3, 1, 9, 0, 0, 1, 0x68, 5, 0, 0, 3, 0, 0, 0, 0, 1, 0x65, 7, 0, 0, 3])
    nals = list(P.split_annexb(np.frombuffer(raw, np.uint8)))
    assert [(t, r) for t, r, _ in nals] == [(7, 3), (8, 3), (5, 3)]
    assert nals[0][2].tolist() == [1, 2, 0, 0, 1, 9]
    # every zero in front of the next 01 belongs to its start code (p264decoder.c:259-301), and the
    # reference only unescapes while src < end-3 (core/core.c:317): the trailing 00 00 03 survives
    assert nals[1][2].tolist() == [5, 0, 0, 3]
    assert nals[2][2].tolist() == [7, 0, 0, 3]


@pytest.mark.skipif(not O.have_ref(), reason="needs oracle/_ref")
def test_nal_decode_equals_reference():
    ref = O.ref()

    class Nal(C.Structure):
        _fields_ = [("i_ref_idc", C.c_int), ("i_type", C.c_int), ("i_payload", C.c_int), ("p_payload", C.c_void_p)]

    rng = np.random.default_rng(5)
    lib = P.load_library()
    for n in [1, 2, 3, 4, 5, 9, 64, 1000]:
        for _ in range(20):
            data = rng.choice(np.array([0, 0, 0, 3, 1, 7], np.uint8), size=n).astype(np.uint8)
            buf_r, buf_o = np.zeros(n + 8, np.uint8), np.zeros(n + 8, np.uint8)
            nal = Nal(0, 0, 0, buf_r.ctypes.data)
            ref.p264_nal_decode(C.byref(nal), data.ctypes.data, n)
            ty, ri = C.c_int(), C.c_int()
            m = lib.p264b200_nal_unescape(data.ctypes.data, n, buf_o.ctypes.data, C.byref(ty), C.byref(ri))
            assert (m, ty.value, ri.value) == (nal.i_payload, nal.i_type, nal.i_ref_idc)
            assert np.array_equal(buf_r[:m], buf_o[:m])


def test_parser_rejects_garbage_without_crashing():
    p = P.Parser()
    rng = np.random.default_rng(1)
    # slice before any parameter set
    with pytest.raises(P.P264Error):
        p.nal(5, 3, rng.integers(0, 256, 100, dtype=np.uint8))
    assert p.nal(6, 0, np.zeros(4, np.uint8)) is None  # SEI ignored
    with pytest.raises(P.P264Error):
        p.nal(2, 0, np.zeros(4, np.uint8))  # data partitioning unsupported (decoder/decoder.c:790-795)
    for _ in range(50):
        try:
            p.nal(int(rng.integers(1, 9)), 3, rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8))
        except P.P264Error:
            pass


def test_synth_stream_invariants():
    syn = P.Synth(9, 7, n_refs=3, seed=9, intra_pct=10)
    last_qp = None
    for i in range(6):
        fr = syn.next()
        h, m = fr.hdr, fr.mbs
        assert (h.mb_w, h.mb_h) == (9, 7) and h.dst_slot == i % 4
        assert h.slice_type == (P.SLICE_I if i == 0 else P.SLICE_P)
        assert h.num_ref == min(i, 3)
        assert sorted({h.ref_slot[k] for k in range(h.num_ref)} | {h.dst_slot}) == sorted({(i - k) % 4 for k in range(h.num_ref + 1)})
        intra = m["mb_type"] <= P.MB_I16x16
        assert intra.sum() == h.n_intra
        assert (m["ref"][intra] == -1).all() and (m["mv"][intra] == 0).all()
        assert (m["ref"][~intra] >= 0).all() and (m["ref"][~intra] < max(h.num_ref, 1)).all()
        assert (m["coef_off"] % 8 == 0).all() and h.n_coef % 8 == 0
        # qp_dbf follows the reference's last-QP rule
        for mb in m:
            coded = mb["mb_type"] == P.MB_I16x16 or mb["luma_mask"] or mb["cbp_chroma"]
            if not coded and last_qp is not None:
                assert mb["qp_dbf"] == last_qp
            elif coded:
                assert mb["qp_dbf"] == mb["qp"]
            last_qp = int(mb["qp_dbf"])
    # determinism
    a, b = P.Synth(5, 4, seed=3), P.Synth(5, 4, seed=3)
    for _ in range(3):
        fa, fb = a.next(), b.next()
        assert fa.mbs.tobytes() == fb.mbs.tobytes() and fa.coefs.tobytes() == fb.coefs.tobytes()


def test_empty_and_tiny_pictures_through_oracle():
    # 1x1 macroblock pictures: every neighbour unavailable
    syn = P.Synth(1, 1, n_refs=1, seed=2, intra_pct=50)
    ring = O.OracleFrames(1, 1, 2)
    for _ in range(5):
        y, u, v = ring.recon(syn.next())
        assert y.shape == (16, 16)


def test_bench_parity_gate_helpers():
    """bench.py's parity gate (host side): lane selection and the oracle replay worker that re-derives a lane's
    reference ring from the staged pictures, cycled the way the timed loops cycle them."""
    import bench

    lanes = bench.pick_check_lanes(256, 8)
    assert len(lanes) == 8 and lanes[0] == 0 and lanes[-1] == 255 and 3 in lanes and 4 in lanes
    assert bench.pick_check_lanes(1, 8) == [0] and bench.pick_check_lanes(2, 8) == [0, 1] and bench.pick_check_lanes(9, 0) == []
    mb_w, mb_h, n_slots, T = 5, 4, 2, 2
    syn = P.Synth(mb_w, mb_h, n_refs=1, seed=5, first_intra=0)
    frames = [syn.next() for _ in range(T)]
    job = (mb_w, mb_h, n_slots, [11, 12], [(bytes(f.hdr), f.mbs.tobytes(), f.coefs.tobytes()) for f in frames], [3, 6])
    got = bench._replay_worker(job)
    ring = O.OracleFrames(mb_w, mb_h, n_slots)
    for s, seed in enumerate([11, 12]):
        ring.set(s, *P.smooth_picture(16 * mb_w, 16 * mb_h, seed=seed))
    for cp_i, cp in enumerate([3, 6]):
        for i in range(0 if cp_i == 0 else 3, cp):
            ring.recon(frames[i % T])
        want = [b"".join(np.ascontiguousarray(p).tobytes() for p in fr) for fr in ring.frames]
        assert got[cp_i] == want


def test_gop_scan_f26():
    """closed-GOP splitter, host side: bin/f26.264 has IDR pictures at 0 and 250; every GOP must begin at
    the parameter sets in front of its IDR slice and the picture counts must add up to the 300 frames"""
    path = O.f26_path()
    if path is None:
        pytest.skip("oracle/_ref/f26.264 not built")
    data = np.fromfile(path, dtype=np.uint8)
    gops = P.gop_scan(data)
    assert [n for _, n in gops] == [250, 50]
    assert gops[0][0] == 0
    for off, _ in gops:
        assert bytes(data[off : off + 4]) == b"\x00\x00\x00\x01" and (data[off + 4] & 0x1f) in (6, 7)   # SEI / SPS first


@pytest.mark.parametrize("kw", [dict(intra_pct=10, n_refs=2), dict(intra_pct=0, sub8x8=0, max_level=2000, qp_min=0, qp_max=51, coded_pct=60), dict(first_intra=1),
                                dict(coded_pct=0, skip_pct=60), dict(first_intra=0)])
def test_wire_format_v2_round_trip(kw):
    """FrameSyntax v2 (compact wire format): pack -> reference unpack reproduces every field a kernel reads -- all vectors,
    the record tails, every coefficient -- and is at least 2.5x smaller on the bench-shaped stream."""
    mb_w, mb_h = 13, 7
    syn = P.Synth(mb_w, mb_h, seed=17, **kw)
    for _ in range(4):
        fr = syn.next()
        fs = fr.syntax()
        v2, blob = P.pack_v2(fs)
        assert v2.blob_bytes <= len(blob) and v2.blob_bytes % 16 == 0
        back = P.unpack_v2(v2)
        a, b = fr.mbs.copy(), back.mbs.copy()
        inter = a["mb_type"] > P.MB_I16x16
        for arr in (a, b):
            arr["reserved"] = 0
            arr["i4_mode"][inter] = 0
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        assert np.array_equal(fr.coefs[: fr.hdr.n_coef], back.coefs[: fr.hdr.n_coef])
        v1_bytes = fr.mbs.nbytes + 2 * fr.hdr.n_coef
        if kw.get("max_level", 8) <= 127:
            assert v2.flags & 1
        if kw == dict(first_intra=0):   # the bench-shaped stream: all partition shapes, 25 % coded blocks
            assert v2.blob_bytes * 2.5 < v1_bytes


def test_gop_scan_written_stream_and_lane_streams_parse():
    """closed-GOP splitter on a written stream with an IDR every 3 pictures (host only): the scan finds every GOP with its picture
    count, each GOP begins at its parameter sets, and a GOP cut out of the stream parses on its own to the same FrameSyntax the
    serial parse produced for those pictures (an IDR resets the parser's DPB: decoder/decoder.c:43-64)"""
    mb_w, mb_h, n_pic = 8, 6, 11
    syn = P.Synth(mb_w, mb_h, n_refs=1, seed=3, first_intra=1, intra_period=3, sub8x8=0, max_level=3)
    wr = P.Writer(mb_w, mb_h)
    for _ in range(n_pic):
        wr.put(syn.next_syntax())
    stream = np.frombuffer(wr.data(), dtype=np.uint8).copy()
    gops = P.gop_scan(stream)
    assert [n for _, n in gops] == [3, 3, 3, 2]
    serial = list(P.Parser(verbose=False).parse_stream(stream))
    assert len(serial) == n_pic
    first = 0
    for gi, (off, n) in enumerate(gops):
        end = gops[gi + 1][0] if gi + 1 < len(gops) else len(stream)
        assert (stream[off + 4] & 0x1f) == 7
        alone = list(P.Parser(verbose=False).parse_stream(stream[off:end]))
        assert len(alone) == n
        for a, b in zip(alone, serial[first : first + n]):
            assert np.array_equal(a.mbs.view(np.uint8), b.mbs.view(np.uint8)) and np.array_equal(a.coefs, b.coefs)
            assert a.hdr.slice_type == b.hdr.slice_type and a.hdr.num_ref == b.hdr.num_ref
        first += n
