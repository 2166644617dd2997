"""CPU: the oracle restatement against the reference's OWN reconstruction (oracle/_ref MB-feed
harness) on synthetic P streams that bin/f26.264 does not cover: sub-8x8 partitions, several
reference frames, QP sweep 20..40 (both dequant branches), deblock offsets, chroma QP offset,
intra MBs inside P pictures and an all-intra first picture."""
import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libp264ref.so not built (needs /root/reference)")

CASES = [
    dict(mb_w=8, mb_h=6, n_refs=1, seed=1, sub8x8=1, intra_pct=0, max_level=3),
    dict(mb_w=8, mb_h=6, n_refs=1, seed=2, sub8x8=1, intra_pct=10, max_level=3, sweep_offsets=1),
    dict(mb_w=11, mb_h=9, n_refs=4, seed=3, sub8x8=1, intra_pct=5, max_level=2, sweep_offsets=1, chroma_qp_index_offset=-3),
    dict(mb_w=6, mb_h=5, n_refs=2, seed=4, sub8x8=0, intra_pct=30, max_level=2, chroma_qp_index_offset=4, qp_min=12, qp_max=48, qp_step=3),
    dict(mb_w=5, mb_h=4, n_refs=1, seed=5, sub8x8=1, intra_pct=0, max_level=2, deblock=0, coded_pct=60),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c['seed']}")
def test_oracle_equals_reference_reconstruction(case):
    case = dict(case)
    mb_w, mb_h = case.pop("mb_w"), case.pop("mb_h")
    n_slots = case["n_refs"] + 1
    syn = P.Synth(mb_w, mb_h, first_intra=1, confine_mv=1, **case)
    ring = O.OracleFrames(mb_w, mb_h, n_slots)
    feed = O.RefFeed(mb_w, mb_h, n_slots)
    O.oracle().orc_hc_overrun_count(1)
    try:
        for i in range(9):
            fr = syn.next()
            want = feed.recon(fr)
            got = ring.recon(fr)
            for name, g, w in zip("YUV", got, want):
                bad = np.argwhere(g != w)
                assert len(bad) == 0, f"picture {i} plane {name}: {len(bad)} samples differ, first at {bad[:4].tolist()}"
            # the real reference is only an oracle while mc_hc stays inside its clip table
            assert O.oracle().orc_hc_overrun_count(0) == 0, "generator left the reference's defined envelope"
    finally:
        feed.close()
