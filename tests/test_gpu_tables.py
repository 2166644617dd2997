"""GPU: known-answer tests of the x264-style function-pointer tables (include/p264_b200_tables.h)
against the reference's OWN tables (oracle/_ref/libp264ref.so: p264_dct_init, p264_mc_init, ...)
on seeded random blocks -- SURVEY.md 8(c)(i).  Covers the table-only slots too (8x8 transform,
bi-pred averages)."""
import ctypes as C

import numpy as np
import pytest

import p264decoder_b200 as P
import _oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.have_ref(), reason="needs oracle/_ref/libp264ref.so")]

VP = C.c_void_p
F_ADD = C.CFUNCTYPE(None, VP, C.c_int, VP)
F_COEF = C.CFUNCTYPE(None, VP)
F_DEQ = C.CFUNCTYPE(None, VP, VP, C.c_int)
F_PRED = C.CFUNCTYPE(None, VP, C.c_int)
F_DBF = C.CFUNCTYPE(None, VP, C.c_int, C.c_int, C.c_int, VP)
F_DBFI = C.CFUNCTYPE(None, VP, C.c_int, C.c_int, C.c_int)
F_AVG = C.CFUNCTYPE(None, VP, C.c_int, VP, C.c_int)
F_AVGW = C.CFUNCTYPE(None, VP, C.c_int, VP, C.c_int, C.c_int)
F_CMP = C.CFUNCTYPE(C.c_int, VP, C.c_int, VP, C.c_int)
F_MCL = C.CFUNCTYPE(None, VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
F_MCC = C.CFUNCTYPE(None, VP, C.c_int, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


def table(lib, init, n, *pre):
    t = (VP * n)()
    getattr(lib, init)(*pre, C.byref(t)) if pre else getattr(lib, init)(C.byref(t))
    return t


@pytest.fixture(scope="module")
def libs():
    ours, ref = P.load_library(), O.ref()
    assert ours.p264b200_tables_ready() == 0
    return ours, ref


def rnd_coefs(rng, shape, big=False):
    c = rng.integers(-600, 600, shape).astype(np.int16)
    if big:
        c = rng.integers(-32768, 32767, shape).astype(np.int16)
    c[rng.random(shape) < 0.5] = 0
    return c


def test_dct_tables(libs):
    ours, ref = libs
    to, tr = table(ours, "p264_dct_init", 14, 0), table(ref, "p264_dct_init", 14, 0)
    rng = np.random.default_rng(1)
    for slot, nblk, size in [(1, 1, 4), (3, 4, 8), (5, 16, 16)]:  # add4x4 / add8x8 / add16x16_idct
        for it in range(12):
            pix = rng.integers(0, 256, (24, 32)).astype(np.uint8)
            c = rnd_coefs(rng, (nblk, 4, 4), big=it >= 8)
            a, b = pix.copy(), pix.copy()
            F_ADD(tr[slot])(a.ctypes.data + 4 * 32 + 8, 32, c.copy().ctypes.data)
            F_ADD(to[slot])(b.ctypes.data + 4 * 32 + 8, 32, c.copy().ctypes.data)
            assert np.array_equal(a, b), (slot, it)
    for slot, nblk in [(7, 1), (9, 4)]:  # add8x8_idct8 / add16x16_idct8 (table-only High-profile path)
        for it in range(8):
            pix = rng.integers(0, 256, (24, 32)).astype(np.uint8)
            c = rnd_coefs(rng, (nblk, 8, 8))
            a, b = pix.copy(), pix.copy()
            F_ADD(tr[slot])(a.ctypes.data + 4 * 32 + 8, 32, c.copy().ctypes.data)
            F_ADD(to[slot])(b.ctypes.data + 4 * 32 + 8, 32, c.copy().ctypes.data)
            assert np.array_equal(a, b), (slot, it)
    for slot, n in [(11, 16), (12, 4), (13, 4)]:  # idct4x4dc, dct2x2dc, idct2x2dc
        for it in range(8):
            c = rnd_coefs(rng, (n,), big=it >= 4)
            a, b = c.copy(), c.copy()
            F_COEF(tr[slot])(a.ctypes.data)
            F_COEF(to[slot])(b.ctypes.data)
            assert np.array_equal(a, b), (slot, it)
    for slot in (0, 2, 4, 6, 8, 10):  # forward transforms are encoder-only
        assert not to[slot]


def test_dequant_tables(libs):
    ours, ref = libs
    tq = table(ours, "p264_quant_init", 6, None, 0)
    assert not tq[0] and not tq[1] and not tq[2] and not tq[3]
    ref.ref_dequant4_table.argtypes = [C.c_int] * 4
    mf4 = np.array([[[ref.ref_dequant4_table(0, q, y, x) for x in range(4)] for y in range(4)] for q in range(6)], np.int32)
    scale8 = np.array([[20, 18, 32, 19, 25, 24], [22, 19, 35, 21, 28, 26], [26, 23, 42, 24, 33, 31], [28, 25, 45, 26, 35, 33], [32, 28, 51, 30, 40, 38], [36, 32, 58, 34, 46, 43]])
    scan = np.array([0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1]).reshape(4, 4)
    mf8 = np.array([[[16 * scale8[q][scan[y & 3][x & 3]] for x in range(8)] for y in range(8)] for q in range(6)], np.int32)
    rng = np.random.default_rng(2)
    for qp in range(52):
        for big in (False, True):
            c = rnd_coefs(rng, (16,), big)
            a, b = c.copy(), c.copy()
            ref.ref_dequant_4x4(a.ctypes.data, qp)
            F_DEQ(tq[4])(b.ctypes.data, mf4.ctypes.data, qp)
            assert np.array_equal(a, b), ("dequant_4x4", qp)
            c = rnd_coefs(rng, (64,), big)
            a, b = c.copy(), c.copy()
            ref.ref_dequant_8x8(a.ctypes.data, qp)
            F_DEQ(tq[5])(b.ctypes.data, mf8.ctypes.data, qp)
            assert np.array_equal(a, b), ("dequant_8x8", qp)
            for n, rf, of in ((16, ref.ref_dequant_4x4_dc, ours.p264_mb_dequant_4x4_dc), (4, ref.ref_dequant_2x2_dc, ours.p264_mb_dequant_2x2_dc)):
                c = rnd_coefs(rng, (n,), big)
                a, b = c.copy(), c.copy()
                rf(a.ctypes.data, qp)
                of(C.c_void_p(b.ctypes.data), C.c_void_p(mf4.ctypes.data), qp)
                assert np.array_equal(a, b), (n, qp)


def test_predict_tables(libs):
    ours, ref = libs
    rng = np.random.default_rng(3)
    for init, n, size in [("p264_predict_16x16_init", 7, 16), ("p264_predict_8x8c_init", 7, 8), ("p264_predict_4x4_init", 12, 4)]:
        to, tr = table(ours, init, n, 0), table(ref, init, n, 0)
        for mode in range(n):
            for it in range(4):
                pix = rng.integers(0, 256, (40, 48)).astype(np.uint8)
                a, b = pix.copy(), pix.copy()
                F_PRED(tr[mode])(a.ctypes.data + 8 * 48 + 8, 48)
                F_PRED(to[mode])(b.ctypes.data + 8 * 48 + 8, 48)
                assert np.array_equal(a, b), (init, mode)
    t8 = table(ours, "p264_predict_8x8_init", 12, 0)
    assert not any(t8)  # Intra-8x8 is unreachable in the reference decoder; slots stay empty


def test_deblock_tables(libs):
    ours, ref = libs
    to, tr = table(ours, "p264_deblock_init", 8, 0), table(ref, "p264_deblock_init", 8, 0)
    rng = np.random.default_rng(4)
    # our slots run the PACKED s16x2 filters of swar.cuh on the device (the instructions of the frame kernel)
    for it in range(160):
        base = rng.integers(60, 200)
        pix = np.clip(base + rng.integers(-12, 13, (40, 48)), 0, 255).astype(np.uint8)
        if it % 3 == 0:
            pix = rng.integers(0, 256, (40, 48)).astype(np.uint8)
        alpha, beta = int(rng.integers(0, 256)), int(rng.integers(0, 19))
        if it >= 40:
            # near-threshold lines: a step of about alpha across the edge (both directions), texture of about beta beside it,
            # so |p0-q0| - alpha, |p1-p0| - beta, |p2-p0| - beta and |p0-q0| - ((alpha>>2)+2) all change sign inside one call
            alpha, beta = int(rng.integers(2, 80)), int(rng.integers(1, 19))
            step = alpha + rng.integers(-2, 3, (40, 48)) if it % 2 else (alpha >> 2) + 2 + rng.integers(-2, 3, (40, 48))
            rr, cc = np.mgrid[0:40, 0:48]
            tex = rng.integers(-1, 2, (40, 48)) * (beta + rng.integers(-1, 2, (40, 48))) // (1 + (it % 3))
            lo = 0 if it % 5 else -30   # some blocks near 0 / 255 for the clip8 of p0 + delta
            pix = np.clip(base + lo * 4 + ((rr >= 12) ^ (cc >= 12)) * step + tex, 0, 255).astype(np.uint8)
        tc = rng.integers(-1, 12, 4).astype(np.int8)
        if it >= 100:
            tc = rng.integers(0, 26, 4).astype(np.int8)   # tc0 table tops out at 25
        for slot in range(4):
            a, b = pix.copy(), pix.copy()
            F_DBF(tr[slot])(a.ctypes.data + 12 * 48 + 12, 48, alpha, beta, tc.ctypes.data)
            F_DBF(to[slot])(b.ctypes.data + 12 * 48 + 12, 48, alpha, beta, tc.ctypes.data)
            assert np.array_equal(a, b), ("normal", slot, it)
        for slot in range(4, 8):
            a, b = pix.copy(), pix.copy()
            F_DBFI(tr[slot])(a.ctypes.data + 12 * 48 + 12, 48, alpha, beta)
            F_DBFI(to[slot])(b.ctypes.data + 12 * 48 + 12, 48, alpha, beta)
            assert np.array_equal(a, b), ("intra", slot, it)


def test_mc_tables(libs):
    ours, ref = libs
    n_mc = 3 + 10 + 10
    to, tr = table(ours, "p264_mc_init", n_mc, 0), table(ref, "p264_mc_init", n_mc, 0)
    rng = np.random.default_rng(5)
    wh = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 2), (2, 4), (2, 2)]
    for i, (w, h) in enumerate(wh):  # bi-pred average slots (table-only: B slices are unsupported)
        d0, s0 = rng.integers(0, 256, (20, 24)).astype(np.uint8), rng.integers(0, 256, (20, 24)).astype(np.uint8)
        a, b = d0.copy(), d0.copy()
        F_AVG(tr[3 + i])(a.ctypes.data, 24, s0.ctypes.data, 24)
        F_AVG(to[3 + i])(b.ctypes.data, 24, s0.ctypes.data, 24)
        assert np.array_equal(a, b), ("avg", w, h)
        for wt in (32, 5, 60, -20, 90):
            a, b = d0.copy(), d0.copy()
            F_AVGW(tr[13 + i])(a.ctypes.data, 24, s0.ctypes.data, 24, wt)
            F_AVGW(to[13 + i])(b.ctypes.data, 24, s0.ctypes.data, 24, wt)
            assert np.array_equal(a, b), ("avg_weight", w, h, wt)
    # mc_luma / mc_chroma: reference on its real filtered planes vs ours on the integer plane only
    mb_w, mb_h = 6, 5
    W, H = 16 * mb_w, 16 * mb_h
    feed = O.RefFeed(mb_w, mb_h, 2)
    y, u, v = P.smooth_picture(W, H, seed=3)
    y = np.clip(y.astype(int) + rng.integers(-20, 21, y.shape), 48, 200).astype(np.uint8)
    feed.set(0, y, u, v)
    ypad = np.pad(y, 32, mode="edge")
    upad = np.pad(u, 16, mode="edge")
    ref.ref_mc_luma.argtypes = [VP] + [C.c_int] * 7 + [VP, C.c_int]
    ref.ref_mc_chroma.argtypes = [VP] + [C.c_int] * 8 + [VP, C.c_int]
    for w, h in wh[:7]:
        for it in range(24):
            bx, by = int(rng.integers(0, W - w + 1)) & ~3, int(rng.integers(0, H - h + 1)) & ~3
            mvx = int(np.clip(rng.integers(-80, 80), 4 * (-24 - bx), 4 * (W + 24 - w - bx) - 1))
            mvy = int(np.clip(rng.integers(-80, 80), 4 * (-24 - by), 4 * (H + 24 - h - by) - 1))
            a, b = np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8)
            ref.ref_mc_luma(feed.h, 0, bx, by, mvx, mvy, w, h, a.ctypes.data, 16)
            srcs = (VP * 4)(ypad.ctypes.data + (32 + by) * ypad.shape[1] + 32 + bx, 0, 0, 0)
            F_MCL(to[0])(C.addressof(srcs), ypad.shape[1], b.ctypes.data, 16, mvx, mvy, w, h)
            assert np.array_equal(a, b), ("mc_luma", w, h, mvx & 3, mvy & 3)
            cw, ch, cx, cy = w // 2, h // 2, bx // 2, by // 2
            a, b = np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8)
            ref.ref_mc_chroma(feed.h, 0, 1, cx, cy, mvx, mvy, cw, ch, a.ctypes.data, 8)
            F_MCC(to[2])(upad.ctypes.data + (16 + cy) * upad.shape[1] + 16 + cx, upad.shape[1], b.ctypes.data, 8, mvx, mvy, cw, ch)
            assert np.array_equal(a, b), ("mc_chroma", cw, ch, mvx & 7, mvy & 7)
    feed.close()


def test_pixel_ssd_table(libs):
    ours, ref = libs
    n = 7 + 7 + 7 + 4 + 7
    to, tr = table(ours, "p264_pixel_init", n, 0), table(ref, "p264_pixel_init", n, 0)
    rng = np.random.default_rng(6)
    for i in range(7):
        a, b = rng.integers(0, 256, (20, 24)).astype(np.uint8), rng.integers(0, 256, (20, 24)).astype(np.uint8)
        assert F_CMP(tr[7 + i])(a.ctypes.data, 24, b.ctypes.data, 24) == F_CMP(to[7 + i])(a.ctypes.data, 24, b.ctypes.data, 24)
    assert not to[0] and not to[14] and not to[21]  # sad / satd / sa8d: encoder-only
