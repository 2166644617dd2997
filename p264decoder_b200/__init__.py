"""ctypes binding of libp264b200.so -- the B200-native H.264 macroblock reconstruction engine.

The product is the C-ABI shared library (``include/*.h``); this module is the thin Python
mirror used by the tests, ``bench.py`` and ``__graft_entry__``.  It never computes anything
itself and there is no CPU fallback: if the library or a CUDA device is missing the calls
raise.

Reference surface mirrored here (all paths relative to the reference tree):
  * ``Parser``  -> NAL switch + slice/MB parse of decoder/decoder.c:745-806, decoder/macroblock.c:488-592
  * ``Engine``  -> the per-picture reconstruction of decoder/decoder.c:623-661
  * ``Decoder`` -> p264_decoder_open/decode/close (p264.h:379-382)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libp264b200.so"

MB_I4x4, MB_I16x16, MB_P_L0, MB_P_8x8, MB_P_SKIP = range(5)
SLICE_P, SLICE_I = 0, 2

# numpy view of struct p264b200_mb (96 bytes)
MB_DTYPE = np.dtype(
    [
        ("mv", "<i2", (16, 2)),
        ("ref", "i1", (4,)),
        ("mb_type", "u1"),
        ("qp", "u1"),
        ("qp_dbf", "u1"),
        ("cbp_chroma", "u1"),
        ("luma_mask", "<u2"),
        ("i16_mode", "u1"),
        ("chroma_mode", "u1"),
        ("i4_mode", "u1", (8,)),
        ("coef_off", "<u4"),
        ("chroma_mask", "u1"),
        ("part", "u1"),
        ("sub_part", "u1", (4,)),
        ("reserved", "u1", (2,)),
    ]
)
assert MB_DTYPE.itemsize == 96


class FrameHdr(C.Structure):
    _fields_ = [
        ("mb_w", C.c_int32),
        ("mb_h", C.c_int32),
        ("slice_type", C.c_int32),
        ("deblock", C.c_int32),
        ("alpha_c0_offset", C.c_int32),
        ("beta_offset", C.c_int32),
        ("chroma_qp_index_offset", C.c_int32),
        ("num_ref", C.c_int32),
        ("ref_slot", C.c_int32 * 16),
        ("dst_slot", C.c_int32),
        ("n_intra", C.c_int32),
        ("n_coef", C.c_uint32),
        ("reserved", C.c_int32 * 4),
    ]


class FrameSyntax(C.Structure):
    _fields_ = [("hdr", FrameHdr), ("mbs", C.c_void_p), ("coefs", C.c_void_p)]


class FrameSyntaxV2(C.Structure):
    _fields_ = [("hdr", FrameHdr), ("blob", C.c_void_p), ("blob_bytes", C.c_uint32), ("flags", C.c_uint32), ("off_hdr", C.c_uint32),
                ("off_offs", C.c_uint32), ("off_mv", C.c_uint32), ("off_mask", C.c_uint32), ("off_level", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class EngineCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("lanes", C.c_int32),
        ("mb_w", C.c_int32),
        ("mb_h", C.c_int32),
        ("n_slots", C.c_int32),
        ("coef_capacity", C.c_uint32),
        ("stage_steps", C.c_int32),
        ("flags", C.c_uint32),
    ]


class SynthCfg(C.Structure):
    _fields_ = [
        ("mb_w", C.c_int32),
        ("mb_h", C.c_int32),
        ("n_refs", C.c_int32),
        ("seed", C.c_uint64),
        ("qp_min", C.c_int32),
        ("qp_max", C.c_int32),
        ("qp_step", C.c_int32),
        ("coded_pct", C.c_int32),
        ("max_level", C.c_int32),
        ("mv_range", C.c_int32),
        ("sub8x8", C.c_int32),
        ("intra_pct", C.c_int32),
        ("skip_pct", C.c_int32),
        ("deblock", C.c_int32),
        ("sweep_offsets", C.c_int32),
        ("chroma_qp_index_offset", C.c_int32),
        ("confine_mv", C.c_int32),
        ("first_intra", C.c_int32),
        ("intra_period", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class P264Error(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libp264b200.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise P264Error(f"{LIB_PATH} is missing: run `make -C p264decoder_b200/csrc` (no CPU fallback exists)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, u8p = C.c_void_p, C.c_int, C.c_void_p
    sig = {
        "p264b200_last_error": (C.c_char_p, []),
        "p264b200_abi_version": (i32, []),
        "p264b200_device_count": (i32, []),
        "p264b200_engine_create": (i32, [C.POINTER(vp), C.POINTER(EngineCfg)]),
        "p264b200_engine_destroy": (None, [vp]),
        "p264b200_engine_geometry": (i32, [vp] + [C.POINTER(C.c_int32)] * 4),
        "p264b200_stage_frame": (i32, [vp, i32, i32, C.POINTER(FrameSyntax)]),
        "p264b200_recon_step": (i32, [vp, i32, i32]),
        "p264b200_stage_frames": (i32, [vp, i32, i32, C.POINTER(FrameSyntax)]),
        "p264b200_stage_frames_v2": (i32, [vp, i32, i32, C.POINTER(FrameSyntaxV2)]),
        "p264b200_pack_v2_bound": (C.c_size_t, [i32, i32, C.c_uint32]),
        "p264b200_pack_v2": (i32, [C.POINTER(FrameSyntax), u8p, C.c_size_t, C.POINTER(FrameSyntaxV2)]),
        "p264b200_unpack_v2": (i32, [C.POINTER(FrameSyntaxV2), u8p, u8p]),
        "p264b200_frames_download": (i32, [vp, i32, C.POINTER(C.c_int32), u8p, C.c_size_t]),
        "p264b200_frames_md5": (i32, [vp, i32, C.POINTER(C.c_int32), u8p]),
        "p264b200_frame_device_planes": (i32, [vp, i32, i32, C.POINTER(C.c_void_p * 3)]),
        "p264b200_recon_frame": (i32, [vp, i32, C.POINTER(FrameSyntax)]),
        "p264b200_frame_upload": (i32, [vp, i32, i32, u8p, i32, u8p, u8p, i32]),
        "p264b200_frame_download": (i32, [vp, i32, i32, u8p, i32, u8p, u8p, i32]),
        "p264b200_engine_sync": (i32, [vp]),
        "p264b200_engine_stream": (vp, [vp]),
        "p264b200_timer_start": (i32, [vp]),
        "p264b200_timer_stop": (i32, [vp, C.POINTER(C.c_float)]),
        "p264b200_profile_enable": (i32, [vp, i32]),
        "p264b200_profile_read": (i32, [vp, C.POINTER(C.c_float * 8), C.POINTER(C.c_uint64 * 8)]),
        "p264b200_engine_launches": (C.c_uint64, [vp]),
        "p264b200_host_alloc": (vp, [C.c_size_t]),
        "p264b200_host_free": (None, [vp]),
        "p264b200_parser_open": (vp, [i32, i32]),
        "p264b200_parser_close": (None, [vp]),
        "p264b200_parser_nal": (i32, [vp, i32, i32, u8p, i32, C.POINTER(FrameSyntax), C.POINTER(i32)]),
        "p264b200_parser_geometry": (i32, [vp] + [C.POINTER(i32)] * 3),
        "p264b200_annexb_next": (i32, [u8p, C.c_size_t] + [C.POINTER(C.c_size_t)] * 3),
        "p264b200_nal_unescape": (i32, [u8p, i32, u8p, C.POINTER(i32), C.POINTER(i32)]),
        "p264b200_cavlc_table_entry": (i32, [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]),
        "p264b200_gop_scan": (i32, [u8p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32), i32]),
        "p264b200_gopdec_open": (i32, [C.POINTER(vp), i32, i32, i32, u8p, C.c_size_t]),
        "p264b200_gopdec_close": (None, [vp]),
        "p264b200_gopdec_gops": (i32, [vp]),
        "p264b200_gopdec_next": (i32, [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]),
        "p264b200_synth_default": (None, [C.POINTER(SynthCfg), i32, i32]),
        "p264b200_synth_open": (vp, [C.POINTER(SynthCfg)]),
        "p264b200_synth_close": (None, [vp]),
        "p264b200_synth_next": (i32, [vp, C.POINTER(FrameSyntax)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc < 0:
        msg = load_library().p264b200_last_error().decode(errors="replace")
        raise P264Error(f"{what} failed with {rc}: {msg}")


def split_annexb(data: np.ndarray):
    """Yield (nal_type, nal_ref_idc, payload ndarray) for every NAL of an Annex-B byte stream."""
    lib = load_library()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    pos, start, size = C.c_size_t(0), C.c_size_t(), C.c_size_t()
    buf = np.empty(len(data) + 16, dtype=np.uint8)
    while lib.p264b200_annexb_next(data.ctypes.data, len(data), C.byref(pos), C.byref(start), C.byref(size)):
        ty, ri = C.c_int(), C.c_int()
        n = lib.p264b200_nal_unescape(data.ctypes.data + start.value, size.value, buf.ctypes.data, C.byref(ty), C.byref(ri))
        if n < 0:
            continue
        yield ty.value, ri.value, buf[:n].copy()


def raw_nals(data: np.ndarray):
    """Yield the raw (still escaped, header byte included) NAL units of an Annex-B stream."""
    lib = load_library()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    pos, start, size = C.c_size_t(0), C.c_size_t(), C.c_size_t()
    while lib.p264b200_annexb_next(data.ctypes.data, len(data), C.byref(pos), C.byref(start), C.byref(size)):
        yield data[start.value : start.value + size.value]


class Frame:
    """Host copy of one FrameSyntax (owns its arrays)."""

    def __init__(self, hdr: FrameHdr, mbs: np.ndarray, coefs: np.ndarray):
        self.hdr, self.mbs, self.coefs = hdr, mbs, coefs

    @classmethod
    def from_syntax(cls, fs: FrameSyntax) -> "Frame":
        h = FrameHdr.from_buffer_copy(fs.hdr)
        n = h.mb_w * h.mb_h
        mbs = np.frombuffer(C.string_at(fs.mbs, n * 96), dtype=MB_DTYPE).copy()
        coefs = np.frombuffer(C.string_at(fs.coefs, h.n_coef * 2), dtype=np.int16).copy() if h.n_coef else np.zeros(8, np.int16)
        return cls(h, mbs, coefs)

    def syntax(self) -> FrameSyntax:
        fs = FrameSyntax()
        fs.hdr = self.hdr
        fs.mbs = self.mbs.ctypes.data
        fs.coefs = self.coefs.ctypes.data
        return fs


def pack_v2(fs: FrameSyntax, dst: np.ndarray | None = None):
    """FrameSyntax v1 -> (FrameSyntaxV2, backing uint8 array): the compact wire format (no GPU needed).  `dst` (optional)
    is a 16-byte aligned uint8 buffer of at least p264b200_pack_v2_bound bytes to pack into."""
    lib = load_library()
    need = lib.p264b200_pack_v2_bound(fs.hdr.mb_w, fs.hdr.mb_h, fs.hdr.n_coef)
    if dst is None:
        raw = np.empty(need + 16, np.uint8)
        o = (-raw.ctypes.data) % 16
        dst = raw[o : o + need]
    out = FrameSyntaxV2()
    _check(lib.p264b200_pack_v2(C.byref(fs), dst.ctypes.data, len(dst), C.byref(out)), "p264b200_pack_v2")
    return out, dst


def unpack_v2(v2: FrameSyntaxV2) -> "Frame":
    """reference expander (host): FrameSyntaxV2 -> Frame"""
    lib = load_library()
    n = v2.hdr.mb_w * v2.hdr.mb_h
    mbs = np.zeros(n, dtype=MB_DTYPE)
    coefs = np.zeros(max(8, v2.hdr.n_coef), np.int16)
    _check(lib.p264b200_unpack_v2(C.byref(v2), mbs.ctypes.data, coefs.ctypes.data), "p264b200_unpack_v2")
    return Frame(FrameHdr.from_buffer_copy(v2.hdr), mbs, coefs[: max(8, v2.hdr.n_coef)])


class Parser:
    """Host syntax front-end (no GPU needed)."""

    def __init__(self, pinned: bool = False, verbose: bool = False):
        self._lib = load_library()
        self._p = self._lib.p264b200_parser_open(int(pinned), int(verbose))
        if not self._p:
            raise P264Error("p264b200_parser_open failed (pinned buffers need a CUDA device)")

    def close(self):
        if self._p:
            self._lib.p264b200_parser_close(self._p)
            self._p = None

    __del__ = close

    def nal(self, nal_type: int, nal_ref_idc: int, payload: np.ndarray):
        """Returns a FrameSyntax (pointing at parser-owned memory) when a picture completed, else None."""
        fs, got = FrameSyntax(), C.c_int(0)
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        rc = self._lib.p264b200_parser_nal(self._p, nal_type, nal_ref_idc, payload.ctypes.data, len(payload), C.byref(fs), C.byref(got))
        if rc < 0:
            raise P264Error(f"p264b200_parser_nal failed with {rc}")
        return fs if got.value else None

    def geometry(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._lib.p264b200_parser_geometry(self._p, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def parse_stream(self, data: np.ndarray):
        """Yield Frame objects (copies) for every picture of an Annex-B stream."""
        for ty, ri, payload in split_annexb(data):
            fs = self.nal(ty, ri, payload)
            if fs is not None:
                yield Frame.from_syntax(fs)


class Synth:
    """Synthetic P-frame stream generator (BASELINE.json configs 3-5); see csrc/host/synth.cc."""

    def __init__(self, mb_w, mb_h, **kw):
        self._lib = load_library()
        self.cfg = SynthCfg()
        self._lib.p264b200_synth_default(C.byref(self.cfg), mb_w, mb_h)
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise TypeError(f"unknown synth option {k}")
            setattr(self.cfg, k, v)
        self._s = self._lib.p264b200_synth_open(C.byref(self.cfg))
        if not self._s:
            raise P264Error("p264b200_synth_open: bad configuration")

    def close(self):
        if getattr(self, "_s", None):
            self._lib.p264b200_synth_close(self._s)
            self._s = None

    __del__ = close

    def next_syntax(self) -> FrameSyntax:
        """Next picture, pointing at generator-owned memory (valid until the next call)."""
        fs = FrameSyntax()
        _check(self._lib.p264b200_synth_next(self._s, C.byref(fs)), "p264b200_synth_next")
        return fs

    def next(self) -> Frame:
        return Frame.from_syntax(self.next_syntax())


class Writer:
    """Annex-B bitstream writer (csrc/host/writer.cc): FrameSyntax pictures -> a stream the unmodified
    reference decoder can decode.  Test / benchmark infrastructure."""

    def __init__(self, mb_w, mb_h, chroma_qp_index_offset=0):
        lib = self._lib = load_library()
        lib.p264b200_writer_open.restype = C.c_void_p
        lib.p264b200_writer_open.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.p264b200_writer_close.argtypes = [C.c_void_p]
        lib.p264b200_writer_put.argtypes = [C.c_void_p, C.POINTER(FrameSyntax)]
        lib.p264b200_writer_data.restype = C.c_void_p
        lib.p264b200_writer_data.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        self._w = lib.p264b200_writer_open(mb_w, mb_h, chroma_qp_index_offset)
        if not self._w:
            raise P264Error("p264b200_writer_open failed")

    def put(self, fs: "FrameSyntax") -> int:
        n = self._lib.p264b200_writer_put(self._w, C.byref(fs))
        if n < 0:
            raise P264Error(f"p264b200_writer_put failed ({n}): the picture is outside what the stock decoder can decode")
        return n

    def data(self) -> bytes:
        n = C.c_size_t()
        p = self._lib.p264b200_writer_data(self._w, C.byref(n))
        return C.string_at(p, n.value)

    def close(self):
        if getattr(self, "_w", None):
            self._lib.p264b200_writer_close(self._w)
            self._w = None

    __del__ = close


def smooth_picture(width, height, seed=0):
    """Bounded smooth I420 test picture (samples in [64,192]) used to seed reference slots: keeps the
    reference's mc_hc inside its 416-entry clip table (core/clip1.h:25-36)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    ph = rng.uniform(0, 6.28, 6)
    fx, fy = rng.uniform(0.01, 0.06, 3), rng.uniform(0.01, 0.06, 3)
    img = 128 + 24 * np.sin(fx[0] * xx + ph[0]) + 20 * np.sin(fy[0] * yy + ph[1]) + 16 * np.sin(fx[1] * xx + fy[1] * yy + ph[2])
    y = np.clip(img + rng.integers(-3, 4, img.shape), 64, 192).astype(np.uint8)
    c = lambda a, b: np.clip(128 + 30 * np.sin(fx[2] * xx[::2, ::2] * a + fy[2] * yy[::2, ::2] * b + ph[3]) + rng.integers(-2, 3, (height // 2, width // 2)), 64, 192).astype(np.uint8)
    return y, c(1.0, 0.5), c(0.4, 1.2)


class Engine:
    """GPU reconstruction engine: `lanes` independent streams per launch."""

    def __init__(self, mb_w, mb_h, n_slots=2, lanes=1, stage_steps=1, coef_capacity=0, device=0):
        self._lib = load_library()
        self.cfg = EngineCfg(device, lanes, mb_w, mb_h, n_slots, coef_capacity, stage_steps, 0)
        self._e = C.c_void_p()
        _check(self._lib.p264b200_engine_create(C.byref(self._e), C.byref(self.cfg)), "p264b200_engine_create")
        self.width, self.height = 16 * mb_w, 16 * mb_h
        self.lanes, self.n_slots, self.stage_steps = lanes, n_slots, stage_steps

    def close(self):
        if getattr(self, "_e", None):
            self._lib.p264b200_engine_destroy(self._e)
            self._e = None

    __del__ = close

    def stage(self, step: int, lane: int, fs: FrameSyntax):
        _check(self._lib.p264b200_stage_frame(self._e, step, lane, C.byref(fs)), "p264b200_stage_frame")

    def stage_v2(self, step: int, frames_v2):
        """lanes [0, len) of one step from packed (v2) pictures"""
        arr = (FrameSyntaxV2 * len(frames_v2))(*frames_v2)
        _check(self._lib.p264b200_stage_frames_v2(self._e, step, len(frames_v2), arr), "p264b200_stage_frames_v2")

    def recon_step(self, step: int, n_lanes: int | None = None):
        _check(self._lib.p264b200_recon_step(self._e, step, n_lanes or self.lanes), "p264b200_recon_step")

    def recon_frame(self, fs: FrameSyntax, lane: int = 0):
        _check(self._lib.p264b200_recon_frame(self._e, lane, C.byref(fs)), "p264b200_recon_frame")

    def sync(self):
        _check(self._lib.p264b200_engine_sync(self._e), "p264b200_engine_sync")

    def upload(self, lane: int, slot: int, y: np.ndarray, u: np.ndarray, v: np.ndarray):
        y, u, v = (np.ascontiguousarray(a, dtype=np.uint8) for a in (y, u, v))
        _check(self._lib.p264b200_frame_upload(self._e, lane, slot, y.ctypes.data, y.shape[1], u.ctypes.data, v.ctypes.data, u.shape[1]), "p264b200_frame_upload")
        self.sync()

    def download(self, lane: int, slot: int):
        y = np.empty((self.height, self.width), np.uint8)
        u = np.empty((self.height // 2, self.width // 2), np.uint8)
        v = np.empty_like(u)
        _check(self._lib.p264b200_frame_download(self._e, lane, slot, y.ctypes.data, self.width, u.ctypes.data, v.ctypes.data, self.width // 2), "p264b200_frame_download")
        self.sync()
        return y, u, v

    def md5(self, slots):
        """hex MD5 of the tight I420 picture in ring slot slots[l] of every lane l < len(slots), computed on the device"""
        n = len(slots)
        out = np.zeros(16 * n, np.uint8)
        arr = (C.c_int32 * n)(*slots)
        _check(self._lib.p264b200_frames_md5(self._e, n, arr, out.ctypes.data), "p264b200_frames_md5")
        self.sync()
        return [out[16 * l : 16 * l + 16].tobytes().hex() for l in range(n)]

    def device_planes(self, lane: int, slot: int):
        """device addresses of sample (0, 0) of the Y, U, V planes of a ring slot (zero-copy output)"""
        pl = (C.c_void_p * 3)()
        _check(self._lib.p264b200_frame_device_planes(self._e, lane, slot, C.byref(pl)), "p264b200_frame_device_planes")
        return [int(p) for p in pl]

    def timer_start(self):
        _check(self._lib.p264b200_timer_start(self._e), "p264b200_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self._lib.p264b200_timer_stop(self._e, C.byref(ms)), "p264b200_timer_stop")
        return ms.value

    def profile_enable(self, on: bool = True):
        _check(self._lib.p264b200_profile_enable(self._e, int(on)), "p264b200_profile_enable")

    def profile_read(self):
        ms, n = (C.c_float * 8)(), (C.c_uint64 * 8)()
        _check(self._lib.p264b200_profile_read(self._e, C.byref(ms), C.byref(n)), "p264b200_profile_read")
        names = ["recon_inter", "recon_intra", "deblock", "border", "deblock_bs"]
        return {k: (ms[i], n[i]) for i, k in enumerate(names)}

    @property
    def launches(self) -> int:
        return int(self._lib.p264b200_engine_launches(self._e))

    @property
    def stream(self) -> int:
        return int(self._lib.p264b200_engine_stream(self._e) or 0)


def gop_scan(data: np.ndarray, max_gops: int = 4096):
    """[(byte offset, pictures)] of every closed GOP of an Annex-B stream (no GPU needed)."""
    lib = load_library()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    begin, pics = (C.c_size_t * max_gops)(), (C.c_int32 * max_gops)()
    n = lib.p264b200_gop_scan(data.ctypes.data, len(data), begin, pics, max_gops)
    _check(n, "p264b200_gop_scan")
    return [(int(begin[i]), int(pics[i])) for i in range(min(n, max_gops))]


def decode_annexb_gops(data: np.ndarray, lanes: int, device: int = 0, threads: int = 0):
    """One Annex-B stream cut at its IDR pictures, closed GOP g decoded on lane g mod `lanes` of one batched engine;
    yields the tight I420 pictures (one uint8 array each) in STREAM order."""
    lib = load_library()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    d = C.c_void_p()
    _check(lib.p264b200_gopdec_open(C.byref(d), device, lanes, threads, data.ctypes.data, len(data)), "p264b200_gopdec_open")
    try:
        pic, w, h = C.c_void_p(), C.c_int(), C.c_int()
        while True:
            r = lib.p264b200_gopdec_next(d, C.byref(pic), C.byref(w), C.byref(h))
            _check(r, "p264b200_gopdec_next")
            if r == 0:
                break
            yield np.frombuffer(C.string_at(pic.value, w.value * h.value * 3 // 2), dtype=np.uint8), w.value, h.value
    finally:
        lib.p264b200_gopdec_close(d)


def decode_annexb(data: np.ndarray, device: int = 0):
    """Host entropy decode + GPU reconstruction of a whole Annex-B stream.

    Yields tight I420 frames (y, u, v) -- the Python spelling of the Decode() loop of
    p264decoder.c:164-381 on top of the frame-level ABI.
    """
    parser = Parser(pinned=False, verbose=False)
    engine = None
    for ty, ri, payload in split_annexb(data):
        fs = parser.nal(ty, ri, payload)
        if fs is None:
            continue
        if engine is None or (engine.cfg.mb_w, engine.cfg.mb_h) != (fs.hdr.mb_w, fs.hdr.mb_h):
            mb_w, mb_h, ring = parser.geometry()
            engine = Engine(mb_w, mb_h, n_slots=ring, device=device)
        engine.recon_frame(fs)
        yield engine.download(0, fs.hdr.dst_slot)
