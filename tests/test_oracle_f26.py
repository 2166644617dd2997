"""CPU: host parser + oracle restatement reproduce the reference decoder's output on bin/f26.264
(config 1 of BASELINE.json) -- this is what pins the oracle and the parser."""
import hashlib

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O


@pytest.fixture(scope="module")
def f26_frames():
    path = O.f26_path()
    if path is None:
        pytest.skip("oracle/_ref/f26.264 not built (needs /root/reference; run make -C oracle ref)")
    data = np.fromfile(path, dtype=np.uint8)
    return list(P.Parser(verbose=False).parse_stream(data))


def test_f26_syntax_statistics(f26_frames):
    # the coverage numbers SURVEY.md 8(c) probed from an instrumented reference build
    assert len(f26_frames) == 300
    types = np.concatenate([f.mbs["mb_type"] for f in f26_frames])
    assert (types == P.MB_I4x4).sum() == 2977
    assert (types == P.MB_I16x16).sum() == 2387
    assert (types == P.MB_P_L0).sum() == 80147
    assert (types == P.MB_P_8x8).sum() == 10073
    assert (types == P.MB_P_SKIP).sum() == 23216
    assert [f.hdr.slice_type for f in f26_frames].count(P.SLICE_I) == 2
    mv = np.concatenate([f.mbs["mv"].reshape(-1, 2) for f in f26_frames])
    assert (mv[:, 0].min(), mv[:, 0].max(), mv[:, 1].min(), mv[:, 1].max()) == (-215, 240, -181, 108)


def test_f26_oracle_matches_reference_golden(f26_frames):
    golden = O.f26_frame_md5s()
    ring = O.OracleFrames(22, 18, 2)
    whole = hashlib.md5()
    for i, fr in enumerate(f26_frames):
        y, u, v = ring.recon(fr)
        assert O.i420_md5(y, u, v) == golden[i], f"frame {i} differs from the reference decoder"
        for p in (y, u, v):
            whole.update(p.tobytes())
    assert whole.hexdigest() == "a482adab07324894b443e081a84ee1df"
    assert O.oracle().orc_hc_overrun_count(0) == 0  # f26 never leaves the reference's clip table
