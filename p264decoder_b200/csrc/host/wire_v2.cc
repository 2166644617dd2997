// FrameSyntax v2: the compact wire format of the host -> device cut (ABI version 2).
//
// v1 (include/p264b200_recon.h) ships 96 bytes per macroblock and 16 int16 slots per coded block -- 2.09 MB per dense
// synthetic 1080p picture, mostly zeros and repeated vectors, which is what bounds the end-to-end path once the PCIe
// link (one GPU) or the host's aggregate DMA rate (eight GPUs) is saturated.  v2 ships
//   * the 32-byte tail of the v1 record (everything except mv[16]) per macroblock,
//   * one vector per PARTITION (the shape is re-derived from mv[16] itself, so records whose informational
//     part / sub_part fields disagree with their vectors still round-trip),
//   * per coded block a 16-bit significance mask + its non-zero levels (int8 when every level of the picture fits),
// and the engine expands it back into the v1 staging layout on the device (csrc/cuda/expand_v2.cuh), so every
// reconstruction kernel is unchanged and v1 stays accepted.  This file is the host side: packer + reference unpacker.
#include <cstring>

#include "../../../include/p264b200_host.h"

namespace {

inline uint32_t mv_word(const p264b200_mb &m, int b)
{
    uint32_t w;
    memcpy(&w, m.mv[b], 4);
    return w;
}
// shape code: bits 0-1 macroblock (0 one vector, 1 two rows of 16x8, 2 two columns of 8x16, 3 per quadrant),
// bits 2+2q.. quadrant q (0 one vector, 1 8x4 top/bottom, 2 4x8 left/right, 3 four vectors)
int shape_of(const p264b200_mb &m)
{
    uint32_t v[16];
    for (int b = 0; b < 16; b++) v[b] = mv_word(m, b);
    bool all = true, rows = true, cols = true;
    for (int b = 0; b < 16; b++) {
        all &= v[b] == v[0];
        rows &= v[b] == v[(b >> 3) * 8];
        cols &= v[b] == v[((b & 3) >> 1) * 2];
    }
    if (all) return 0;
    if (rows) return 1;
    if (cols) return 2;
    int code = 3;
    for (int q = 0; q < 4; q++) {
        const int b0 = 8 * (q >> 1) + 2 * (q & 1);
        const uint32_t a = v[b0], b = v[b0 + 1], c = v[b0 + 4], d = v[b0 + 5];
        int qc = 3;
        if (a == b && a == c && a == d)
            qc = 0;
        else if (a == b && c == d)
            qc = 1;
        else if (a == c && b == d)
            qc = 2;
        code |= qc << (2 + 2 * q);
    }
    return code;
}
const int kQuadMvs[4] = {1, 2, 2, 4};
int shape_mvs(int code)
{
    const int s = code & 3;
    if (s == 0) return 1;
    if (s != 3) return 2;
    int n = 0;
    for (int q = 0; q < 4; q++) n += kQuadMvs[(code >> (2 + 2 * q)) & 3];
    return n;
}
// index of block b's vector inside the macroblock's vector run
int shape_index(int code, int b)
{
    const int bx = b & 3, by = b >> 2, s = code & 3;
    if (s == 0) return 0;
    if (s == 1) return by >> 1;
    if (s == 2) return bx >> 1;
    const int q = (by >> 1) * 2 + (bx >> 1);
    int base = 0;
    for (int k = 0; k < q; k++) base += kQuadMvs[(code >> (2 + 2 * k)) & 3];
    const int qc = (code >> (2 + 2 * q)) & 3;
    return base + (qc == 0 ? 0 : qc == 1 ? (by & 1) : qc == 2 ? (bx & 1) : (by & 1) * 2 + (bx & 1));
}
// first block (in raster order) that uses vector k of the run
int shape_first_block(int code, int k)
{
    for (int b = 0; b < 16; b++)
        if (shape_index(code, b) == k) return b;
    return 0;
}
inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// 16-slot blocks of one macroblock's v1 chunk, in chunk order; the chroma DC group (8 slots) is reported with n = 8
template <class F>
void for_each_block(const p264b200_mb &m, F f)
{
    size_t off = m.coef_off;
    if (m.mb_type == P264B200_MB_I16x16) f(off, 16), off += 16;
    for (int b = 0; b < 16; b++)
        if (m.luma_mask >> b & 1) f(off, 16), off += 16;
    if (m.cbp_chroma) {
        f(off, 8), off += 8;
        for (int i = 0; i < 8; i++)
            if (m.chroma_mask >> i & 1) f(off, 16), off += 16;
    }
}

}  // namespace

extern "C" {

size_t p264b200_pack_v2_bound(int mb_w, int mb_h, uint32_t n_coef)
{
    const size_t n_mb = (size_t)mb_w * mb_h;
    // headers + two offsets + 16 vectors per macroblock, one mask per 8 coefficient slots + every slot as a 16-bit level
    return align16(n_mb * 32) + align16(n_mb * 8) + align16(n_mb * 64) + align16(((size_t)n_coef / 8 + n_mb) * 2) + align16((size_t)n_coef * 2) + 64;
}

int p264b200_pack_v2(const p264b200_frame_syntax *fs, uint8_t *dst, size_t cap, p264b200_frame_syntax_v2 *out)
{
    if (!fs || !dst || !out || !fs->mbs || ((uintptr_t)dst & 15)) return P264B200_EINVAL;
    const p264b200_frame_hdr &h = fs->hdr;
    const size_t n_mb = (size_t)h.mb_w * h.mb_h;
    if (cap < p264b200_pack_v2_bound(h.mb_w, h.mb_h, h.n_coef)) return P264B200_ENOMEM;
    // pass 1: do all levels fit int8?
    bool fit8 = true;
    for (size_t i = 0; i < (size_t)h.n_coef && fit8; i++) fit8 = fs->coefs[i] >= -128 && fs->coefs[i] <= 127;
    memset(out, 0, sizeof(*out));
    out->hdr = h;
    out->flags = fit8 ? P264B200_V2_LEVELS8 : 0;
    uint8_t *p = dst;
    out->off_hdr = 0;
    uint8_t *hdrs = p;
    p += align16(n_mb * 32);
    out->off_offs = (uint32_t)(p - dst);
    uint32_t *offs = reinterpret_cast<uint32_t *>(p);   // [n_mb][2]: first mask index, first level index
    p += align16(n_mb * 8);
    out->off_mv = (uint32_t)(p - dst);
    uint32_t *mvs = reinterpret_cast<uint32_t *>(p);
    size_t n_mv = 0;
    for (size_t i = 0; i < n_mb; i++) {
        const p264b200_mb &m = fs->mbs[i];
        uint8_t *hd = hdrs + 32 * i;
        memcpy(hd, reinterpret_cast<const uint8_t *>(&m) + 64, 32);
        if (!P264B200_IS_INTRA(m.mb_type)) {
            const int code = shape_of(m), n = shape_mvs(code);
            const uint32_t mv_off = (uint32_t)n_mv;
            memcpy(hd + (offsetof(p264b200_mb, i4_mode) - 64), &mv_off, 4);   // i4_mode[0..3] is unused by inter macroblocks
            const uint16_t c16 = (uint16_t)code;
            memcpy(hd + (offsetof(p264b200_mb, reserved) - 64), &c16, 2);
            for (int k = 0; k < n; k++) mvs[n_mv++] = mv_word(m, shape_first_block(code, k));
        }
    }
    p += align16(n_mv * 4);
    out->off_mask = (uint32_t)(p - dst);
    uint16_t *masks = reinterpret_cast<uint16_t *>(p);
    // masks first (their count is needed to place the level stream)
    size_t n_mask = 0, n_lvl = 0;
    for (size_t i = 0; i < n_mb; i++) {
        offs[2 * i] = (uint32_t)n_mask, offs[2 * i + 1] = (uint32_t)n_lvl;
        for_each_block(fs->mbs[i], [&](size_t off, int n) {
            unsigned mask = 0;
            for (int k = 0; k < n; k++)
                if (fs->coefs[off + k]) mask |= 1u << k, n_lvl++;
            masks[n_mask++] = (uint16_t)mask;
        });
    }
    p += align16(n_mask * 2);
    out->off_level = (uint32_t)(p - dst);
    size_t li = 0;
    for (size_t i = 0; i < n_mb; i++)
        for_each_block(fs->mbs[i], [&](size_t off, int n) {
            for (int k = 0; k < n; k++) {
                const int16_t v = fs->coefs[off + k];
                if (!v) continue;
                if (fit8)
                    reinterpret_cast<int8_t *>(p)[li++] = (int8_t)v;
                else
                    reinterpret_cast<int16_t *>(p)[li++] = v;
            }
        });
    p += align16(li * (fit8 ? 1 : 2));
    out->blob = dst;
    out->blob_bytes = (uint32_t)(p - dst);
    return P264B200_OK;
}

// Reference expander (what csrc/cuda/expand_v2.cuh does on the device); mbs: n_mb records, coefs: hdr.n_coef levels.
// Inter records come back with i4_mode zeroed and `reserved` holding the shape code; everything a kernel reads is exact.
int p264b200_unpack_v2(const p264b200_frame_syntax_v2 *in, p264b200_mb *mbs, int16_t *coefs)
{
    if (!in || !in->blob || !mbs || (in->hdr.n_coef && !coefs)) return P264B200_EINVAL;
    const size_t n_mb = (size_t)in->hdr.mb_w * in->hdr.mb_h;
    const uint8_t *hdrs = in->blob + in->off_hdr;
    const uint32_t *offs = reinterpret_cast<const uint32_t *>(in->blob + in->off_offs);
    const uint32_t *mvs = reinterpret_cast<const uint32_t *>(in->blob + in->off_mv);
    const uint16_t *masks = reinterpret_cast<const uint16_t *>(in->blob + in->off_mask);
    const uint8_t *lv = in->blob + in->off_level;
    const bool fit8 = in->flags & P264B200_V2_LEVELS8;
    if (in->hdr.n_coef) memset(coefs, 0, (size_t)in->hdr.n_coef * 2);
    for (size_t i = 0; i < n_mb; i++) {
        p264b200_mb &m = mbs[i];
        memset(&m, 0, 64);
        memcpy(reinterpret_cast<uint8_t *>(&m) + 64, hdrs + 32 * i, 32);
        if (!P264B200_IS_INTRA(m.mb_type)) {
            uint32_t mv_off;
            uint16_t code;
            memcpy(&mv_off, m.i4_mode, 4);
            memcpy(&code, m.reserved, 2);
            memset(m.i4_mode, 0, 4);
            for (int b = 0; b < 16; b++) memcpy(m.mv[b], &mvs[mv_off + shape_index(code, b)], 4);
        }
        size_t mi = offs[2 * i], li = offs[2 * i + 1];
        for_each_block(m, [&](size_t off, int n) {
            const unsigned mask = masks[mi++];
            for (int k = 0; k < n; k++)
                if (mask >> k & 1) {
                    coefs[off + k] = fit8 ? (int16_t) reinterpret_cast<const int8_t *>(lv)[li] : reinterpret_cast<const int16_t *>(lv)[li];
                    li++;
                }
        });
    }
    return P264B200_OK;
}

}  // extern "C"
