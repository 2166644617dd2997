"""GPU: configs 3/4/5 of BASELINE.json at parity-test sizes -- synthetic P streams (all partition
shapes, all 16 quarter-pel phases, QP sweep, multi-reference, deblock offsets, intra MBs) through
the C-ABI engine, byte-compared with the CPU oracle picture by picture."""
import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = pytest.mark.gpu


def _compare(got, want, fr, what):
    for name, g, w, s in zip("YUV", got, want, (16, 8, 8)):
        bad = np.argwhere(g != w)
        if len(bad):
            mbs = sorted({(int(r) // s, int(c) // s) for r, c in bad})[:6]
            kinds = [(my, mx, int(fr.mbs["mb_type"][my * fr.hdr.mb_w + mx])) for my, mx in mbs]
            raise AssertionError(f"{what} plane {name}: {len(bad)} samples differ; first MBs (y,x,type) {kinds}")


def _run_stream(mb_w, mb_h, n_frames, lanes=1, **kw):
    n_slots = kw.get("n_refs", 1) + 1
    eng = P.Engine(mb_w, mb_h, n_slots=n_slots, lanes=lanes)
    syns = [P.Synth(mb_w, mb_h, **dict(kw, seed=kw.get("seed", 7) + 101 * l)) for l in range(lanes)]
    rings = [O.OracleFrames(mb_w, mb_h, n_slots) for _ in range(lanes)]
    if not kw.get("first_intra", 1):
        for l in range(lanes):
            for s in range(n_slots):
                pic = P.smooth_picture(16 * mb_w, 16 * mb_h, seed=l * 10 + s)
                eng.upload(l, s, *pic)
                rings[l].set(s, *pic)
    for i in range(n_frames):
        frames = [s.next() for s in syns]
        for l, fr in enumerate(frames):
            eng.stage(0, l, fr.syntax())
        eng.recon_step(0, lanes)
        eng.sync()
        for l, fr in enumerate(frames):
            want = rings[l].recon(fr)
            got = eng.download(l, fr.hdr.dst_slot)
            _compare(got, want, fr, f"picture {i} lane {l}")
    eng.close()


def test_small_all_features():
    _run_stream(8, 6, 8, n_refs=1, seed=1, intra_pct=10, sweep_offsets=1)


def test_multi_ref_chroma_offset():
    _run_stream(11, 9, 8, n_refs=4, seed=3, intra_pct=5, sweep_offsets=1, chroma_qp_index_offset=-3)


def test_full_qp_range_and_big_levels():
    # exercises the int16 wrap-around points of dequant / IDCT (levels far beyond real streams)
    _run_stream(6, 5, 10, n_refs=2, seed=4, intra_pct=30, max_level=2000, qp_min=0, qp_max=51, qp_step=3, coded_pct=60)


def test_unconfined_mvs_far_outside_picture():
    # MVs up to +-200 samples: beyond the reference's 32-sample border (undefined there); the engine
    # and the oracle both implement the standard's unrestricted-MV replication
    _run_stream(5, 4, 6, n_refs=1, seed=5, mv_range=200, confine_mv=0, first_intra=0)


def test_no_deblock_and_no_residual():
    _run_stream(7, 5, 4, n_refs=1, seed=6, deblock=0, coded_pct=0, skip_pct=50)


def test_lanes_batched():
    _run_stream(9, 7, 5, lanes=5, n_refs=2, seed=11, intra_pct=8)


@pytest.mark.parametrize("mb_w,mb_h", [(1, 1), (2, 1), (1, 3), (3, 2), (1, 9), (9, 1), (5, 17)])
def test_tiny_and_ragged_geometries(mb_w, mb_h):
    # pictures smaller than one recon_inter tile (8x8 MBs), narrower than the deblock ring (4 MBs), a single
    # macroblock row / column, and 17 rows = two full deblock row groups + one row
    _run_stream(mb_w, mb_h, 4, lanes=3, n_refs=2, seed=31 + mb_w * 7 + mb_h, intra_pct=15, sweep_offsets=1, mv_range=6)


def test_tall_picture_many_lanes_dense_intra():
    # 19 macroblock rows = three deblock row groups (8 + 8 + 3), 9 lanes = two full stream quads + one lane,
    # 40 % intra macroblocks = long intra runs with intra neighbours above (the per-macroblock wait path)
    _run_stream(7, 19, 3, lanes=9, n_refs=1, seed=19, intra_pct=40, sweep_offsets=1)


def test_1080p_stream():
    _run_stream(120, 68, 3, n_refs=1, seed=264, intra_pct=2)


def test_1080p_bench_shaped_many_lanes_cycled_staging():
    """The shape bench.py times: many 1080p lanes (more deblock CTAs than fit on the machine at once), syntax pre-staged in
    HBM, two staged pictures per lane cycled over several steps without intermediate syncs; every lane's final reference
    ring must equal the oracle's."""
    mb_w, mb_h, lanes, T, steps = 120, 68, 72, 2, 5
    eng = P.Engine(mb_w, mb_h, n_slots=2, lanes=lanes, stage_steps=T)
    rings = [O.OracleFrames(mb_w, mb_h, 2) for _ in range(lanes)]
    staged = [[None] * lanes for _ in range(T)]
    for l in range(lanes):
        syn = P.Synth(mb_w, mb_h, n_refs=1, seed=900 + l, first_intra=0, confine_mv=1, intra_pct=0, coded_pct=25, max_level=8, mv_range=16,
                      sub8x8=1, skip_pct=5, qp_min=20, qp_max=40, qp_step=2)
        for s in range(2):
            pic = P.smooth_picture(16 * mb_w, 16 * mb_h, seed=(l * 2 + s) % 4)
            eng.upload(l, s, *pic)
            rings[l].set(s, *pic)
        for t in range(T):
            staged[t][l] = syn.next()
            eng.stage(t, l, staged[t][l].syntax())
    for i in range(steps):
        eng.recon_step(i % T, lanes)
    eng.sync()
    check = sorted(set([0, 1, 5, 35, 36, lanes - 5, lanes - 1]))   # the oracle needs ~30 ms per 1080p picture
    for l in check:
        for i in range(steps):
            rings[l].recon(staged[i % T][l])
        for slot in range(2):
            _compare(eng.download(l, slot), rings[l].frames[slot], staged[0][l], f"lane {l} slot {slot} after {steps} steps")
    eng.close()


def test_4k_multiref_two_pictures():
    _run_stream(240, 135, 2, n_refs=4, seed=2160, sweep_offsets=1, first_intra=0)


def test_4k_multiref_many_lanes():
    """configs[3] at a batched shape: 16 lanes x 4K x 4 references, three unsynchronised steps over pre-staged syntax
    (16 lanes x 17 row groups x 2 roles = 136 deblock CTAs, multi-reference windows from four ring slots)."""
    mb_w, mb_h, lanes, refs, steps = 240, 135, 16, 4, 3
    n_slots = refs + 1
    eng = P.Engine(mb_w, mb_h, n_slots=n_slots, lanes=lanes, stage_steps=steps)
    rings = {}
    staged = [[None] * lanes for _ in range(steps)]
    check = [0, 3, 4, 9, lanes - 1]
    pics = [P.smooth_picture(16 * mb_w, 16 * mb_h, seed=s) for s in range(3)]
    for l in range(lanes):
        syn = P.Synth(mb_w, mb_h, n_refs=refs, seed=4000 + l, first_intra=0, sweep_offsets=1, intra_pct=1 if l == 3 else 0)
        for s in range(n_slots):
            eng.upload(l, s, *pics[(l + s) % 3])
        if l in check:
            rings[l] = O.OracleFrames(mb_w, mb_h, n_slots)
            for s in range(n_slots):
                rings[l].set(s, *pics[(l + s) % 3])
        for t in range(steps):
            staged[t][l] = syn.next()
            eng.stage(t, l, staged[t][l].syntax())
    for t in range(steps):
        eng.recon_step(t, lanes)
    eng.sync()
    for l in check:
        for t in range(steps):
            rings[l].recon(staged[t][l])
        for slot in range(n_slots):
            _compare(eng.download(l, slot), rings[l].frames[slot], staged[0][l], f"4K lane {l} slot {slot} after {steps} steps")
    eng.close()


def test_randomised_stress_time_boxed():
    """tools/stress.py for a fixed wall-clock budget: random geometries, lane counts, reference counts and stream options --
    the check aimed at the schedule-dependent parts (tickets, mbarrier rings, progress words, class buckets)."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
    import stress

    cases, pictures = stress.run(budget=35.0, seed=2)
    assert cases >= 3 and pictures >= 20
    cases, pictures = stress.run(budget=25.0, seed=3, big=True)
    assert cases >= 1


def test_wire_format_v2_staging_matches_oracle():
    """FrameSyntax v2: packed pictures laid back to back in one host arena (one H2D copy), expanded on the device, alternating
    with v1 staging of the other step slot; every picture byte-identical to the oracle."""
    mb_w, mb_h, lanes, steps = 11, 9, 5, 6
    eng = P.Engine(mb_w, mb_h, n_slots=3, lanes=lanes, stage_steps=2)
    rings = [O.OracleFrames(mb_w, mb_h, 3) for _ in range(lanes)]
    syns = [P.Synth(mb_w, mb_h, n_refs=2, seed=500 + l, intra_pct=10 if l != 2 else 100, sweep_offsets=1, max_level=8 if l else 900) for l in range(lanes)]
    lib = P.load_library()
    for s in range(steps):
        frames = [sy.next() for sy in syns]
        if s % 3 == 2:
            for l, fr in enumerate(frames):
                eng.stage(s % 2, l, fr.syntax())
        else:
            need = [lib.p264b200_pack_v2_bound(mb_w, mb_h, fr.hdr.n_coef) for fr in frames]
            raw = np.zeros(sum(need) + 16, np.uint8)
            o = (-raw.ctypes.data) % 16
            v2s, at = [], o
            for fr, nd in zip(frames, need):
                v2, _ = P.pack_v2(fr.syntax(), raw[at : at + nd])
                v2s.append(v2)
                at += v2.blob_bytes if s % 2 == 0 else nd      # even steps: tightly back to back (one copy); odd: gaps (per-lane copies)
            eng.stage_v2(s % 2, v2s)
        eng.recon_step(s % 2, lanes)
        eng.sync()
        for l, fr in enumerate(frames):
            _compare(eng.download(l, fr.hdr.dst_slot), rings[l].recon(fr), fr, f"v2 step {s} lane {l}")
    eng.close()


def test_wire_format_v2_rejects_inconsistent_section_offsets():
    """the engine checks a packed picture's section table on the host before anything is copied: EINVAL, nothing launched"""
    import ctypes as C

    mb_w, mb_h = 6, 4
    eng = P.Engine(mb_w, mb_h, n_slots=2, lanes=1)
    fr = P.Synth(mb_w, mb_h, seed=9, first_intra=0).next()
    v2, blob = P.pack_v2(fr.syntax())
    lib = P.load_library()
    for field, value in (("off_mv", v2.off_mask + 16), ("off_level", v2.blob_bytes + 16), ("off_offs", v2.off_offs + 4), ("blob_bytes", 1 << 30)):
        bad = P.FrameSyntaxV2.from_buffer_copy(v2)
        setattr(bad, field, value)
        arr = (P.FrameSyntaxV2 * 1)(bad)
        assert lib.p264b200_stage_frames_v2(eng._e, 0, 1, arr) == -1, field   # P264B200_EINVAL
    eng.stage_v2(0, [v2])            # the untouched picture still stages and reconstructs
    eng.recon_step(0, 1)
    ring = O.OracleFrames(mb_w, mb_h, 2)
    _compare(eng.download(0, fr.hdr.dst_slot), ring.recon(fr), fr, "after the rejected calls")
    eng.close()


def test_batched_stage_and_packed_download_match_per_picture_path():
    """p264b200_stage_frames / p264b200_frames_download (pack kernel + one contiguous D2H) against the
    per-picture calls, pipelined over several steps without intermediate syncs."""
    import ctypes as C

    mb_w, mb_h, lanes, steps = 10, 6, 6, 4
    lib = P.load_library()
    eng = P.Engine(mb_w, mb_h, n_slots=2, lanes=lanes, stage_steps=2)
    rings = [O.OracleFrames(mb_w, mb_h, 2) for _ in range(lanes)]
    syns = [P.Synth(mb_w, mb_h, n_refs=1, seed=77 + l, intra_pct=5) for l in range(lanes)]
    W, H = 16 * mb_w, 16 * mb_h
    fsz = W * H * 3 // 2
    outs = [np.zeros(lanes * fsz, np.uint8) for _ in range(steps)]
    keep, want = [], []
    for s in range(steps):
        frames = [sy.next() for sy in syns]
        keep.append(frames)  # host buffers must stay alive until the sync
        arr = (P.FrameSyntax * lanes)(*[f.syntax() for f in frames])
        slots = (C.c_int32 * lanes)(*[f.hdr.dst_slot for f in frames])
        assert lib.p264b200_stage_frames(eng._e, s % 2, lanes, arr) == 0
        assert lib.p264b200_recon_step(eng._e, s % 2, lanes) == 0
        assert lib.p264b200_frames_download(eng._e, lanes, slots, outs[s].ctypes.data, fsz) == 0
        want.append([np.concatenate([p.ravel() for p in rings[l].recon(frames[l])]) for l in range(lanes)])
    eng.sync()
    for s in range(steps):
        for l in range(lanes):
            assert np.array_equal(outs[s][l * fsz : (l + 1) * fsz], want[s][l]), f"step {s} lane {l}"
    eng.close()
