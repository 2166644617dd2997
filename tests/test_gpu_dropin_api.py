"""GPU: the reference's public API (p264.h:379-382) served by the engine: p264_decoder_open /
p264_nal_decode / p264_decoder_decode / p264_decoder_close and the CLI with the reference's
grammar, both byte-identical to the reference decoder on bin/f26.264."""
import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


class Image(C.Structure):
    _fields_ = [("i_csp", C.c_int), ("i_plane", C.c_int), ("i_stride", C.c_int * 4), ("plane", C.c_void_p * 4)]


class Picture(C.Structure):
    _fields_ = [("i_type", C.c_int), ("i_qpplus1", C.c_int), ("i_pts", C.c_int64), ("i_width", C.c_int), ("i_height", C.c_int), ("img", Image)]


class Nal(C.Structure):
    _fields_ = [("i_ref_idc", C.c_int), ("i_type", C.c_int), ("i_payload", C.c_int), ("p_payload", C.c_void_p)]


def test_p264_api_decodes_f26_bit_exact():
    path = O.f26_path()
    if path is None:
        pytest.skip("oracle/_ref/f26.264 not present")
    lib = P.load_library()
    lib.p264_decoder_open.restype = C.c_void_p
    lib.p264_decoder_open.argtypes = [C.c_void_p]
    lib.p264_decoder_decode.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Picture)), C.POINTER(Nal)]
    lib.p264_decoder_close.argtypes = [C.c_void_p]
    lib.p264_nal_decode.argtypes = [C.POINTER(Nal), C.c_void_p, C.c_int]
    param = (C.c_uint8 * 1024)()
    lib.p264_param_default(C.byref(param))
    h = lib.p264_decoder_open(C.byref(param))
    assert h
    data = np.fromfile(path, dtype=np.uint8)
    payload = np.zeros(len(data), np.uint8)
    nal = Nal(0, 0, 0, payload.ctypes.data)
    golden = O.f26_frame_md5s()
    n = 0
    for raw in P.raw_nals(data):
        raw = np.ascontiguousarray(raw)
        assert lib.p264_nal_decode(C.byref(nal), raw.ctypes.data, len(raw)) == 0
        pic = C.POINTER(Picture)()
        assert lib.p264_decoder_decode(h, C.byref(pic), C.byref(nal)) == 0
        if pic:
            p = pic.contents
            assert (p.i_width, p.i_height, p.img.i_plane) == (352, 288, 3)
            assert (p.img.i_stride[0], p.img.i_stride[1]) == (352 + 64, (352 + 64) // 2)  # the reference's padded strides
            m = hashlib.md5()
            for c in range(3):
                w, hh = (352, 288) if c == 0 else (176, 144)
                plane = np.ctypeslib.as_array(C.cast(p.img.plane[c], C.POINTER(C.c_uint8)), shape=(hh * p.img.i_stride[c],))
                rows = plane[: hh * p.img.i_stride[c]].reshape(hh, p.img.i_stride[c])[:, :w]
                m.update(np.ascontiguousarray(rows).tobytes())
            assert m.hexdigest() == golden[n], f"picture {n}"
            n += 1
    assert n == 300
    lib.p264_decoder_close(h)


def test_cli_matches_reference_md5(tmp_path):
    path = O.f26_path()
    exe = ROOT / "p264decoder_b200" / "lib" / "p264dec_b200"
    if path is None or not exe.exists():
        pytest.skip("f26.264 or CLI not present")
    out = tmp_path / "f26.yuv"
    r = subprocess.run([str(exe), "-d", str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "decoded total 300 frames" in r.stderr
    assert hashlib.md5(out.read_bytes()).hexdigest() == "a482adab07324894b443e081a84ee1df"


def test_unmodified_reference_cli_relinked_against_the_engine(tmp_path):
    """The drop-in claim end to end: the reference's OWN CLI source (p264decoder.c:164-381 + core/mdate.c, compiled
    unmodified against its own p264.h by oracle/Makefile where /root/reference is mounted) linked against
    libp264b200.so instead of the reference's decoder/ and core/ objects, decoding bin/f26.264 on the GPU."""
    path = O.f26_path()
    exe = O.REF_DIR / "p264dec_relinked"
    if path is None or not exe.exists():
        pytest.skip("oracle/_ref/p264dec_relinked not built (needs /root/reference at build time)")
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libp264b200.so" in ldd and "libp264ref" not in ldd
    out = tmp_path / "f26_relinked.yuv"
    r = subprocess.run([str(exe), "-d", str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert "decoded total 300 frames" in r.stderr
    assert hashlib.md5(out.read_bytes()).hexdigest() == "a482adab07324894b443e081a84ee1df"
