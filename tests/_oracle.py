"""Test-only bindings of the checkers: oracle/liboracle.so (our CPU restatement) and, when it
was built in this container, oracle/_ref/libp264ref.so (the unmodified reference + harness).
Nothing under p264decoder_b200/ imports this module."""
import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np

import p264decoder_b200 as P

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
GOLDEN = ROOT / "tests" / "golden"

_orc = None
_ref = None


def oracle():
    global _orc
    if _orc is None:
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            subprocess.check_call(["make", "-C", str(ORACLE_DIR), "liboracle.so"])
        _orc = C.CDLL(str(so))
        _orc.orc_recon_frame_flat.argtypes = [C.POINTER(P.FrameHdr), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]
        _orc.orc_hc_overrun_count.restype = C.c_long
        assert _orc.orc_sizeof_mb() == 96 and _orc.orc_sizeof_hdr() == C.sizeof(P.FrameHdr)
    return _orc


def have_ref():
    return (REF_DIR / "libp264ref.so").exists()


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(str(REF_DIR / "libp264ref.so"))
        _ref.ref_feed_open.restype = C.c_void_p
        _ref.ref_feed_open.argtypes = [C.c_int] * 3
        _ref.ref_feed_close.argtypes = [C.c_void_p]
        _ref.ref_feed_write.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        _ref.ref_feed_read.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        _ref.ref_feed_frame.argtypes = [C.c_void_p, C.POINTER(P.FrameHdr), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _ref.ref_dec_open.restype = C.c_void_p
        _ref.ref_dec_close.argtypes = [C.c_void_p]
        _ref.ref_dec_nal.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)] * 2
    return _ref


class OracleFrames:
    """Frame ring of tight I420 planes reconstructed by the CPU oracle."""

    def __init__(self, mb_w, mb_h, n_slots):
        self.W, self.H, self.n = 16 * mb_w, 16 * mb_h, n_slots
        self.frames = [
            [np.full((self.H, self.W), 128, np.uint8), np.full((self.H // 2, self.W // 2), 128, np.uint8), np.full((self.H // 2, self.W // 2), 128, np.uint8)]
            for _ in range(n_slots)
        ]
        self.ptrs = (C.c_void_p * (3 * n_slots))(*[pl.ctypes.data for f in self.frames for pl in f])

    def recon(self, frame: "P.Frame", deblock=True):
        fs = frame.syntax()
        rc = oracle().orc_recon_frame_flat(C.byref(fs.hdr), fs.mbs, fs.coefs, self.ptrs, self.n, int(deblock))
        assert rc == 0
        return self.frames[frame.hdr.dst_slot]

    def set(self, slot, y, u, v):
        for dst, src in zip(self.frames[slot], (y, u, v)):
            dst[...] = src


class RefFeed:
    """The reference's own reconstruction driven from FrameSyntax (MB-feed oracle)."""

    def __init__(self, mb_w, mb_h, n_slots):
        self.W, self.H = 16 * mb_w, 16 * mb_h
        self.h = ref().ref_feed_open(mb_w, mb_h, n_slots)
        assert self.h

    def close(self):
        if self.h:
            ref().ref_feed_close(self.h)
            self.h = None

    def set(self, slot, y, u, v):
        y, u, v = (np.ascontiguousarray(a) for a in (y, u, v))
        ref().ref_feed_write(self.h, slot, y.ctypes.data, u.ctypes.data, v.ctypes.data)

    def get(self, slot):
        y = np.empty((self.H, self.W), np.uint8)
        u = np.empty((self.H // 2, self.W // 2), np.uint8)
        v = np.empty_like(u)
        ref().ref_feed_read(self.h, slot, y.ctypes.data, u.ctypes.data, v.ctypes.data)
        return y, u, v

    def recon(self, frame: "P.Frame", deblock=True):
        fs = frame.syntax()
        rc = ref().ref_feed_frame(self.h, C.byref(fs.hdr), fs.mbs, fs.coefs, int(deblock), 1)
        assert rc == 0
        return self.get(frame.hdr.dst_slot)


def i420_md5(y, u, v):
    m = hashlib.md5()
    for p in (y, u, v):
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


def f26_path():
    """bin/f26.264 is the reference's only test vector; oracle/Makefile copies it next to the
    compiled reference (git-ignored, travels to the GPU box with the snapshot)."""
    p = REF_DIR / "f26.264"
    return p if p.exists() else None


def f26_frame_md5s():
    return (GOLDEN / "f26_frames.md5").read_text().split()
