"""CPU: the bitstream writer (csrc/host/writer.cc) against the UNMODIFIED reference decoder.

A synthetic stream (bitstream variant: first picture intra, one reference frame,
partitions >= 8x8) is written as a real Annex-B file;
  * the reference's own CLI (oracle/_ref/p264dec_ref, built from /root/reference) decodes the file, and its YUV must
    equal the CPU oracle's reconstruction of the original FrameSyntax  -> pins writer + oracle against the real decoder;
  * our host parser reads the same file back, and the oracle's reconstruction of what it parsed must be identical
    -> pins the parser on syntax f26.264 does not contain (other sizes, QPs, deblock offsets, chroma QP offset).
"""
import hashlib
import subprocess

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

REF_CLI = O.ROOT / "oracle" / "_ref" / "p264dec_ref"


def make_stream(mb_w, mb_h, n_pictures, seed, **kw):
    """-> (annexb bytes, [tight I420 bytes per picture as the oracle reconstructs the original syntax])"""
    # Residual energy is kept moderate and an intra picture recurs: on content that saturates to 0 / 255 the reference's
    # mc_hc reads past its 416-entry clip table (core/clip1.h:25-36, undefined behaviour) and its output is
    # no oracle any more -- with max_level 8 / QP up to 40 a 1080p stream first differs (3 samples) in picture 10.
    opts = dict(n_refs=1, seed=seed, sub8x8=0, first_intra=1, intra_period=6, confine_mv=1, qp_min=22, qp_max=34, qp_step=2, max_level=5,
                coded_pct=25, mv_range=16, skip_pct=5)
    opts.update(kw)
    syn = P.Synth(mb_w, mb_h, **opts)
    wr = P.Writer(mb_w, mb_h, opts.get("chroma_qp_index_offset", 0))
    ring = O.OracleFrames(mb_w, mb_h, 2)
    want = []
    for _ in range(n_pictures):
        fr = syn.next()
        wr.put(fr.syntax())
        planes = ring.recon(fr)
        want.append(b"".join(p.tobytes() for p in planes))
    data = wr.data()
    wr.close()
    return data, want


@pytest.mark.parametrize("mb_w,mb_h,n,kw", [
    (6, 5, 6, dict(intra_pct=10, sweep_offsets=1)),
    (11, 9, 5, dict(intra_pct=0, skip_pct=30, chroma_qp_index_offset=-2)),
    (22, 18, 8, dict(intra_pct=5, coded_pct=40)),
    (120, 68, 9, dict(intra_pct=3)),          # BASELINE.json configs[2] at full size, as a real bitstream
])
def test_reference_cli_decodes_written_stream(tmp_path, mb_w, mb_h, n, kw):
    if not REF_CLI.exists():
        pytest.skip("oracle/_ref/p264dec_ref not built (needs /root/reference)")
    data, want = make_stream(mb_w, mb_h, n, seed=5 + mb_w, **kw)
    src, out = tmp_path / "s.264", tmp_path / "s.yuv"
    src.write_bytes(data)
    r = subprocess.run([str(REF_CLI), "-d", str(src), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    got = out.read_bytes()
    fsz = 16 * mb_w * 16 * mb_h * 3 // 2
    assert len(got) == n * fsz, f"reference decoded {len(got) // fsz} of {n} pictures"
    for i in range(n):
        assert got[i * fsz:(i + 1) * fsz] == want[i], f"picture {i}: reference CLI output differs from the oracle's reconstruction of the syntax"


def test_parser_reads_written_stream_back():
    mb_w, mb_h, n = 9, 7, 6
    data, want = make_stream(mb_w, mb_h, n, seed=77, intra_pct=8, sweep_offsets=1, chroma_qp_index_offset=3)
    parser = P.Parser(pinned=False, verbose=False)
    ring = O.OracleFrames(mb_w, mb_h, 2)
    got = []
    for ty, ri, payload in P.split_annexb(np.frombuffer(data, np.uint8)):
        fs = parser.nal(ty, ri, payload)
        if fs is None:
            continue
        planes = ring.recon(P.Frame.from_syntax(fs))
        got.append(b"".join(p.tobytes() for p in planes))
    assert len(got) == n
    for i in range(n):
        assert hashlib.md5(got[i]).hexdigest() == hashlib.md5(want[i]).hexdigest(), f"picture {i}"
