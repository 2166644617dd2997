"""GPU: BASELINE.json configs[2] as a REAL bitstream -- a synthetic 1080p stream written by csrc/host/writer.cc,
decoded (a) by the unmodified reference CLI on the host and (b) by host parser + GPU reconstruction through the
drop-in API and through the multi-stream decoder; the YUV must be byte-identical."""
import ctypes as C
import subprocess

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O
from test_bitstream_writer import REF_CLI, make_stream
from test_gpu_multi import MultiCfg

pytestmark = pytest.mark.gpu


def test_1080p_written_stream_gpu_equals_reference_cli(tmp_path):
    if not REF_CLI.exists():
        pytest.skip("oracle/_ref/p264dec_ref not present")
    mb_w, mb_h, n = 120, 68, 14
    data, _ = make_stream(mb_w, mb_h, n, seed=1080, intra_pct=3)
    src, out = tmp_path / "s.264", tmp_path / "s.yuv"
    src.write_bytes(data)
    r = subprocess.run([str(REF_CLI), "-d", str(src), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    ref = np.fromfile(out, np.uint8)
    fsz = 16 * mb_w * 16 * mb_h * 3 // 2
    assert len(ref) == n * fsz
    # (a) drop-in single-stream path
    k = 0
    for y, u, v in P.decode_annexb(np.frombuffer(data, np.uint8)):
        got = np.concatenate([y.ravel(), u.ravel(), v.ravel()])
        assert np.array_equal(got, ref[k * fsz:(k + 1) * fsz]), f"drop-in decode, picture {k}"
        k += 1
    assert k == n
    # (b) multi-stream decoder, three lanes on the same stream
    lib = P.load_library()
    lib.p264b200_multi_open.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MultiCfg)]
    lib.p264b200_multi_set_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.p264b200_multi_step.argtypes = [C.c_void_p, C.c_void_p]
    lib.p264b200_multi_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.p264b200_multi_picture.restype = C.c_void_p
    lib.p264b200_multi_close.argtypes = [C.c_void_p]
    buf = C.create_string_buffer(data, len(data))
    cfg = MultiCfg(device=0, n_streams=3, n_threads=3)
    m = C.c_void_p()
    assert lib.p264b200_multi_open(C.byref(m), C.byref(cfg)) == 0
    for s in range(3):
        assert lib.p264b200_multi_set_stream(m, s, buf, len(data)) == 0
    k = 0
    while lib.p264b200_multi_step(m, None) > 0:
        for s in range(3):
            w, h = C.c_int(), C.c_int()
            p = lib.p264b200_multi_picture(m, s, C.byref(w), C.byref(h))
            assert p and (w.value, h.value) == (1920, 1088)
            pic = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(fsz,))
            assert np.array_equal(pic, ref[k * fsz:(k + 1) * fsz]), f"multi-stream decode, stream {s} picture {k}"
        k += 1
    lib.p264b200_multi_close(m)
    assert k == n
