// Bitstream writer: FrameSyntax pictures -> a real H.264 Annex-B byte stream (Baseline, CAVLC).
//
// Test / benchmark infrastructure ("the variant that is also emitted as a real
// bitstream for the stock CLI"): the synthetic streams of synth.cc become files that the UNMODIFIED
// reference decoder decodes, so the whole chain -- host parser + GPU reconstruction -- is pinned against
// the reference's own CLI at full 1080p size, not only through the macroblock-feed harness.  It is the
// exact inverse of parser.cc (same syntax order, same predictors: 7.3.5 macroblock layer, 8.4.1.3 motion
// vector prediction, 9.2 CAVLC with the nC context), restricted to what the stock decoder can decode
// (SURVEY 8a quirks): one reference frame, partitions >= 8x8, one slice per picture, mb_qp_delta = 0.
// P_SKIP records (whose synthetic vectors are not the inferred skip vectors) are written as P_L0 16x16
// without residual: a different mb_type, the same reconstruction.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../../include/p264b200_host.h"
#include "cavlc.h"

namespace {

using namespace p264b200;

const uint8_t kCbpIntra[48] = {47, 31, 15, 0,  23, 27, 29, 30, 7,  11, 13, 14, 39, 43, 45, 46,
                               16, 3,  5,  10, 12, 19, 21, 26, 28, 35, 37, 42, 44, 1,  2,  4,
                               8,  17, 18, 20, 24, 6,  9,  22, 25, 32, 33, 34, 36, 40, 38, 41};
const uint8_t kCbpInter[48] = {0,  16, 1,  2,  4,  8,  32, 3,  5,  10, 12, 15, 47, 7,  11, 13,
                               14, 6,  9,  31, 35, 37, 42, 44, 33, 34, 36, 40, 39, 43, 45, 46,
                               17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41};
const uint8_t kZx[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
const uint8_t kZy[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};

inline int median3(int a, int b, int c)
{
    int mn = std::min(a, std::min(b, c)), mx = std::max(a, std::max(b, c));
    return a + b + c - mn - mx;
}

// RBSP bit writer
struct BitWriter {
    std::vector<uint8_t> bytes;
    uint32_t acc = 0;
    int n = 0;
    void put(uint32_t v, int bits)
    {
        for (int i = bits - 1; i >= 0; i--) {
            acc = (acc << 1) | ((v >> i) & 1);
            if (++n == 8) {
                bytes.push_back((uint8_t)acc);
                acc = 0, n = 0;
            }
        }
    }
    void ue(uint32_t v)
    {
        v++;
        int len = 0;
        while ((v >> len) > 1) len++;
        put(0, len);
        put(v, len + 1);
    }
    void se(int v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
    void trailing()
    {
        put(1, 1);
        while (n) put(0, 1);
    }
};

}  // namespace

struct p264b200_writer {
    int mb_w, mb_h, chroma_qp_off;
    int frame_num = 0, idr_id = 0;
    bool headers_written = false;
    std::vector<uint8_t> out;
    // neighbour grids, same meaning and update rules as parser.cc
    std::vector<uint8_t> nnz_y, nnz_c[2];
    std::vector<int8_t> imode, ref4;
    std::vector<int16_t> mv4;
    char err[160] = {0};
};

namespace {

typedef p264b200_writer W;

void emit_nal(W *w, int ref_idc, int type, const BitWriter &bw)
{
    static const uint8_t sc[4] = {0, 0, 0, 1};
    w->out.insert(w->out.end(), sc, sc + 4);
    w->out.push_back((uint8_t)((ref_idc << 5) | type));
    int zeros = 0;
    for (uint8_t b : bw.bytes) {
        if (zeros == 2 && b <= 3) {
            w->out.push_back(3);
            zeros = 0;
        }
        zeros = b == 0 ? zeros + 1 : 0;
        w->out.push_back(b);
    }
}

int predict_nnz(const uint8_t *grid, int stride, int x, int y)
{
    const bool a = x > 0, b = y > 0;
    const int na = a ? grid[y * stride + x - 1] : 0, nb = b ? grid[(y - 1) * stride + x] : 0;
    if (a && b) return (na + nb + 1) >> 1;
    return a ? na : (b ? nb : 0);
}

// 8.4.1.3, identical to Parser::predict_mv
void predict_mv(const W *w, int x4, int y4, int w4, int ref, int shape, int part_idx, int mvp[2])
{
    const int s4 = 4 * w->mb_w;
    auto cell_ref = [&](int x, int y) -> int {
        if (x < 0 || y < 0 || x >= s4 || y >= 4 * w->mb_h) return -2;
        return w->ref4[y * s4 + x];
    };
    auto cell_mv = [&](int x, int y, int c) -> int {
        if (x < 0 || y < 0 || x >= s4 || y >= 4 * w->mb_h) return 0;
        return w->ref4[y * s4 + x] == -2 ? 0 : w->mv4[(y * s4 + x) * 2 + c];
    };
    const int ax = x4 - 1, ay = y4, bx = x4, by = y4 - 1;
    int cx = x4 + w4, cy = y4 - 1;
    int ra = cell_ref(ax, ay), rb = cell_ref(bx, by), rc = cell_ref(cx, cy);
    if (rc == -2) {
        cx = x4 - 1;
        rc = cell_ref(cx, cy);
    }
    const int mva[2] = {cell_mv(ax, ay, 0), cell_mv(ax, ay, 1)};
    const int mvb[2] = {cell_mv(bx, by, 0), cell_mv(bx, by, 1)};
    const int mvc[2] = {cell_mv(cx, cy, 0), cell_mv(cx, cy, 1)};
    if (shape == 1) {
        if (part_idx == 0 && rb == ref) {
            mvp[0] = mvb[0], mvp[1] = mvb[1];
            return;
        }
        if (part_idx != 0 && ra == ref) {
            mvp[0] = mva[0], mvp[1] = mva[1];
            return;
        }
    } else if (shape == 2) {
        if (part_idx == 0 && ra == ref) {
            mvp[0] = mva[0], mvp[1] = mva[1];
            return;
        }
        if (part_idx != 0 && rc == ref) {
            mvp[0] = mvc[0], mvp[1] = mvc[1];
            return;
        }
    }
    const int cnt = (ra == ref) + (rb == ref) + (rc == ref);
    if (cnt == 1) {
        const int *m = ra == ref ? mva : (rb == ref ? mvb : mvc);
        mvp[0] = m[0], mvp[1] = m[1];
    } else if (cnt == 0 && rb == -2 && rc == -2 && ra != -2) {
        mvp[0] = mva[0], mvp[1] = mva[1];
    } else {
        mvp[0] = median3(mva[0], mvb[0], mvc[0]);
        mvp[1] = median3(mva[1], mvb[1], mvc[1]);
    }
}

void fill_motion(W *w, int mbx, int mby, int bx, int by, int bw, int bh, int ref, int mvx, int mvy)
{
    const int s4 = 4 * w->mb_w;
    for (int y = by; y < by + bh; y++)
        for (int x = bx; x < bx + bw; x++) {
            const int g = (4 * mby + y) * s4 + 4 * mbx + x;
            w->ref4[g] = (int8_t)ref;
            w->mv4[g * 2] = (int16_t)mvx;
            w->mv4[g * 2 + 1] = (int16_t)mvy;
        }
}

void put_code(BitWriter &bw, int kind, int table, int sym)
{
    int len = 0, bits = 0;
    cavlc_table_entry(kind, table, sym, &len, &bits);
    bw.put((uint32_t)bits, len);
}

// 9.2: one residual block, levels[0..max_coeff-1] in scan order; returns total_coeff or -1 (level out of range)
int cavlc_write_block(BitWriter &bw, int nC, int max_coeff, const int16_t *levels)
{
    int lev[16], run[16], total = 0;
    // reverse scan order: highest frequency first; run[k] = zeros between coefficient k and the next lower one
    int last = -1, highest = -1;
    for (int i = max_coeff - 1; i >= 0; i--)
        if (levels[i]) {
            lev[total] = levels[i];
            if (total > 0)
                run[total - 1] = last - i - 1;
            else
                highest = i;
            last = i;
            total++;
        }
    int t1s = 0;
    while (t1s < total && t1s < 3 && (lev[t1s] == 1 || lev[t1s] == -1)) t1s++;
    if (nC < 0)
        put_code(bw, 1, 0, total * 4 + t1s);
    else
        put_code(bw, 0, nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3, total * 4 + t1s);
    if (total == 0) return 0;
    const int zeros_total = highest + 1 - total;  // total_zeros: zeros below the highest-frequency coefficient
    run[total - 1] = last;                        // zeros before the lowest-frequency coefficient (implied, never written)
    for (int i = 0; i < t1s; i++) bw.put(lev[i] < 0 ? 1 : 0, 1);
    int suffix_len = (total > 10 && t1s < 3) ? 1 : 0;
    for (int i = t1s; i < total; i++) {
        int code = lev[i] > 0 ? 2 * lev[i] - 2 : -2 * lev[i] - 1;
        if (i == t1s && t1s < 3) code -= 2;
        if (suffix_len == 0) {
            if (code < 14)
                bw.put(1, code + 1);
            else if (code < 30) {
                bw.put(1, 15);
                bw.put((uint32_t)(code - 14), 4);
            } else if (code < 30 + 4096) {
                bw.put(1, 16);
                bw.put((uint32_t)(code - 30), 12);
            } else
                return -1;
        } else {
            if (code < (15 << suffix_len)) {
                bw.put(1, (code >> suffix_len) + 1);
                bw.put((uint32_t)(code & ((1 << suffix_len) - 1)), suffix_len);
            } else if (code < (15 << suffix_len) + 4096) {
                bw.put(1, 16);
                bw.put((uint32_t)(code - (15 << suffix_len)), 12);
            } else
                return -1;
        }
        if (suffix_len == 0) suffix_len = 1;
        const int a = lev[i] < 0 ? -lev[i] : lev[i];
        if (a > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }
    if (total < max_coeff) {
        if (nC < 0)
            put_code(bw, 3, total - 1, zeros_total);
        else
            put_code(bw, 2, total - 1, zeros_total);
    }
    int zeros_left = zeros_total;
    for (int i = 0; i < total - 1 && zeros_left > 0; i++) {
        put_code(bw, 4, (zeros_left > 7 ? 7 : zeros_left) - 1, run[i]);
        zeros_left -= run[i];
    }
    return total;
}

int fail(W *w, const char *what, int mb_xy)
{
    snprintf(w->err, sizeof(w->err), "p264b200_writer: %s (macroblock %d)", what, mb_xy);
    fprintf(stderr, "%s\n", w->err);
    return P264B200_EINVAL;
}

int write_mb(W *w, BitWriter &bw, const p264b200_frame_syntax *fs, int mb_xy, bool pslice, int slice_qp, int *last_qp)
{
    const int mbx = mb_xy % w->mb_w, mby = mb_xy / w->mb_w;
    const int s4 = 4 * w->mb_w, s2 = 2 * w->mb_w;
    const p264b200_mb &m = fs->mbs[mb_xy];
    const bool intra = P264B200_IS_INTRA(m.mb_type), i16 = m.mb_type == P264B200_MB_I16x16, i4 = m.mb_type == P264B200_MB_I4x4;
    if (!pslice && !intra) return fail(w, "inter macroblock in an I picture", mb_xy);
    if (pslice) bw.ue(0);  // mb_skip_run: nothing is skipped (see the header comment)

    // coded_block_pattern from the masks
    int cbp_luma = 0;
    for (int b = 0; b < 16; b++)
        if (m.luma_mask >> b & 1) cbp_luma |= 1 << ((b >> 3) * 2 + ((b & 3) >> 1));
    if (i16) cbp_luma = m.luma_mask ? 15 : 0;
    if (m.cbp_chroma > 2) return fail(w, "cbp_chroma out of range", mb_xy);

    if (intra) {
        const int t = i4 ? 0 : 1 + m.i16_mode + 4 * m.cbp_chroma + (cbp_luma ? 12 : 0);
        bw.ue((uint32_t)(t + (pslice ? 5 : 0)));
        if (i4)
            for (int i = 0; i < 16; i++) {
                const int x = 4 * mbx + kZx[i], y = 4 * mby + kZy[i], b = kZy[i] * 4 + kZx[i];
                const int ma = x > 0 ? w->imode[y * s4 + x - 1] : -1, mb = y > 0 ? w->imode[(y - 1) * s4 + x] : -1;
                int pred = std::min(ma, mb);
                if (pred < 0) pred = 2;
                const int mode = (m.i4_mode[b >> 1] >> ((b & 1) * 4)) & 15;
                if (mode > 8) return fail(w, "intra 4x4 mode out of range", mb_xy);
                if (mode == pred)
                    bw.put(1, 1);
                else {
                    bw.put(0, 1);
                    bw.put((uint32_t)(mode < pred ? mode : mode - 1), 3);
                }
                w->imode[y * s4 + x] = (int8_t)mode;
            }
        bw.ue(m.chroma_mode);
    } else {
        int part = m.mb_type == P264B200_MB_P_SKIP ? P264B200_D_16x16 : m.part;
        if (m.mb_type == P264B200_MB_P_8x8) part = P264B200_D_8x8;
        for (int i = 0; i < 4; i++)
            if (m.ref[i] != 0) return fail(w, "the bitstream variant has one reference frame", mb_xy);
        bw.ue((uint32_t)part);  // 0 16x16, 1 16x8, 2 8x16, 3 8x8
        if (part == P264B200_D_8x8) {
            for (int i = 0; i < 4; i++) {
                if (m.sub_part[i] != P264B200_SUB_8x8) return fail(w, "sub-8x8 partitions are not decodable by the stock parser", mb_xy);
                bw.ue(0);
            }
            for (int i = 0; i < 4; i++) {
                const int bx = 2 * (i & 1), by = 2 * (i >> 1), b = by * 4 + bx;
                int mvp[2];
                predict_mv(w, 4 * mbx + bx, 4 * mby + by, 2, 0, 0, 0, mvp);
                bw.se(m.mv[b][0] - mvp[0]);
                bw.se(m.mv[b][1] - mvp[1]);
                fill_motion(w, mbx, mby, bx, by, 2, 2, 0, m.mv[b][0], m.mv[b][1]);
            }
        } else {
            const int nparts = part == P264B200_D_16x16 ? 1 : 2;
            const int pw = part == P264B200_D_8x16 ? 2 : 4, ph = part == P264B200_D_16x8 ? 2 : 4;
            for (int i = 0; i < nparts; i++) {
                const int bx = part == P264B200_D_8x16 ? 2 * i : 0, by = part == P264B200_D_16x8 ? 2 * i : 0, b = by * 4 + bx;
                const int shape = part == P264B200_D_16x8 ? 1 : (part == P264B200_D_8x16 ? 2 : 0);
                int mvp[2];
                predict_mv(w, 4 * mbx + bx, 4 * mby + by, pw, 0, shape, i, mvp);
                bw.se(m.mv[b][0] - mvp[0]);
                bw.se(m.mv[b][1] - mvp[1]);
                fill_motion(w, mbx, mby, bx, by, pw, ph, 0, m.mv[b][0], m.mv[b][1]);
            }
        }
    }
    if (!i16) {
        const int cbp = cbp_luma | (m.cbp_chroma << 4);
        const uint8_t *tab = i4 ? kCbpIntra : kCbpInter;
        int c = 0;
        while (c < 48 && tab[c] != cbp) c++;
        if (c == 48) return fail(w, "coded_block_pattern not representable", mb_xy);
        bw.ue((uint32_t)c);
    }
    const bool coded = cbp_luma > 0 || m.cbp_chroma > 0 || i16;
    if (coded) {
        if (m.qp != slice_qp) return fail(w, "mb_qp_delta != 0 is not supported by the writer", mb_xy);
        bw.se(0);
        const int16_t *cf = fs->coefs + m.coef_off;
        if (i16) {
            const int nC = predict_nnz(w->nnz_y.data(), s4, 4 * mbx, 4 * mby);
            if (cavlc_write_block(bw, nC, 16, cf) < 0) return fail(w, "level out of range", mb_xy);
            cf += 16;
        }
        const int16_t *blk[16];
        for (int b = 0; b < 16; b++) {
            blk[b] = nullptr;
            if (m.luma_mask >> b & 1) blk[b] = cf, cf += 16;
        }
        static const int16_t zeros[16] = {0};
        for (int i = 0; i < 16; i++) {
            const int bx = kZx[i], by = kZy[i], b = by * 4 + bx, gx = 4 * mbx + bx, gy = 4 * mby + by;
            int tot = 0;
            if (cbp_luma & (1 << (i / 4))) {
                const int nC = predict_nnz(w->nnz_y.data(), s4, gx, gy);
                const int16_t *src = blk[b] ? blk[b] : zeros;
                tot = i16 ? cavlc_write_block(bw, nC, 15, src + 1) : cavlc_write_block(bw, nC, 16, src);
                if (tot < 0) return fail(w, "level out of range", mb_xy);
                if (i16 && src[0] != 0) return fail(w, "Intra16x16 AC block with a DC slot", mb_xy);
            }
            w->nnz_y[gy * s4 + gx] = (uint8_t)tot;
        }
        const int16_t *cdc = cf;
        if (m.cbp_chroma) {
            cf += 8;
            if (cavlc_write_block(bw, -1, 4, cdc) < 0 || cavlc_write_block(bw, -1, 4, cdc + 4) < 0) return fail(w, "level out of range", mb_xy);
        }
        for (int c = 0; c < 2; c++)
            for (int i = 0; i < 4; i++) {
                const int gx = 2 * mbx + (i & 1), gy = 2 * mby + (i >> 1);
                int tot = 0;
                if (m.cbp_chroma & 2) {
                    const int16_t *src = zeros;
                    if (m.chroma_mask >> (c * 4 + i) & 1) src = cf, cf += 16;
                    const int nC = predict_nnz(w->nnz_c[c].data(), s2, gx, gy);
                    tot = cavlc_write_block(bw, nC, 15, src + 1);
                    if (tot < 0) return fail(w, "level out of range", mb_xy);
                }
                w->nnz_c[c][gy * s2 + gx] = (uint8_t)tot;
            }
        *last_qp = m.qp;
    } else {
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) w->nnz_y[(4 * mby + y) * s4 + 4 * mbx + x] = 0;
        for (int c = 0; c < 2; c++)
            for (int i = 0; i < 4; i++) w->nnz_c[c][(2 * mby + (i >> 1)) * s2 + 2 * mbx + (i & 1)] = 0;
    }
    if (!i4)
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) w->imode[(4 * mby + y) * s4 + 4 * mbx + x] = 2;
    if (intra)
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) {
                const int g = (4 * mby + y) * s4 + 4 * mbx + x;
                w->ref4[g] = -1;
                w->mv4[g * 2] = w->mv4[g * 2 + 1] = 0;
            }
    return 0;
}

}  // namespace

extern "C" {

p264b200_writer *p264b200_writer_open(int mb_w, int mb_h, int chroma_qp_index_offset)
{
    if (mb_w < 1 || mb_h < 1 || mb_w > 1024 || mb_h > 1024 || chroma_qp_index_offset < -12 || chroma_qp_index_offset > 12) return nullptr;
    cavlc_init();
    p264b200_writer *w = new (std::nothrow) p264b200_writer;
    if (!w) return nullptr;
    w->mb_w = mb_w, w->mb_h = mb_h, w->chroma_qp_off = chroma_qp_index_offset;
    const size_t n = (size_t)mb_w * mb_h;
    w->nnz_y.assign(n * 16, 0);
    w->nnz_c[0].assign(n * 4, 0);
    w->nnz_c[1].assign(n * 4, 0);
    w->imode.assign(n * 16, 2);
    w->ref4.assign(n * 16, -2);
    w->mv4.assign(n * 32, 0);
    return w;
}

void p264b200_writer_close(p264b200_writer *w) { delete w; }

int p264b200_writer_put(p264b200_writer *w, const p264b200_frame_syntax *fs)
{
    if (!w || !fs || !fs->mbs || fs->hdr.mb_w != w->mb_w || fs->hdr.mb_h != w->mb_h) return P264B200_EINVAL;
    const p264b200_frame_hdr &h = fs->hdr;
    const bool idr = h.slice_type == P264B200_SLICE_I;
    if (!idr && !w->headers_written) return P264B200_EINVAL;  // a stream starts with an intra picture
    if (!idr && h.num_ref != 1) return P264B200_EINVAL;
    const size_t before = w->out.size();
    if (idr) {
        BitWriter sps;
        sps.put(66, 8);   // Baseline
        sps.put(0, 8);    // constraint flags + reserved
        sps.put(51, 8);   // level
        sps.ue(0);        // sps id
        sps.ue(0);        // log2_max_frame_num - 4
        sps.ue(2);        // pic_order_cnt_type 2: nothing per slice
        sps.ue(1);        // num_ref_frames
        sps.put(0, 1);    // gaps_in_frame_num_value_allowed
        sps.ue((uint32_t)(w->mb_w - 1));
        sps.ue((uint32_t)(w->mb_h - 1));
        sps.put(1, 1);    // frame_mbs_only
        sps.put(1, 1);    // direct_8x8_inference
        sps.put(0, 1);    // no cropping
        sps.put(0, 1);    // no VUI
        sps.trailing();
        emit_nal(w, 3, 7, sps);
        BitWriter pps;
        pps.ue(0), pps.ue(0);
        pps.put(0, 1);    // CAVLC
        pps.put(0, 1);    // pic_order_present
        pps.ue(0);        // one slice group
        pps.ue(0), pps.ue(0);   // num_ref_idx_l0/l1_default_active - 1
        pps.put(0, 1);    // weighted_pred
        pps.put(0, 2);    // weighted_bipred_idc
        pps.se(0);        // pic_init_qp - 26
        pps.se(0);        // pic_init_qs - 26
        pps.se(w->chroma_qp_off);
        pps.put(1, 1);    // deblocking_filter_control_present
        pps.put(0, 1);    // constrained_intra_pred
        pps.put(0, 1);    // redundant_pic_cnt_present
        pps.trailing();
        emit_nal(w, 3, 8, pps);
        w->headers_written = true;
        w->frame_num = 0;
    }
    // slice QP = QP of the picture's macroblocks (uniform: mb_qp_delta stays 0)
    int slice_qp = -1;
    const int n_mb = w->mb_w * w->mb_h;
    for (int i = 0; i < n_mb && slice_qp < 0; i++) {
        const p264b200_mb &m = fs->mbs[i];
        if (m.mb_type == P264B200_MB_I16x16 || m.luma_mask || m.cbp_chroma) slice_qp = m.qp;
    }
    if (slice_qp < 0) slice_qp = fs->mbs[0].qp;
    BitWriter bw;
    bw.ue(0);                           // first_mb_in_slice
    bw.ue(idr ? 2u : 0u);               // slice_type I / P
    bw.ue(0);                           // pps id
    bw.put((uint32_t)(w->frame_num & 15), 4);
    if (idr) bw.ue((uint32_t)(w->idr_id++ & 0xffff));
    if (!idr) {
        bw.put(0, 1);                   // num_ref_idx_active_override
        bw.put(0, 1);                   // ref_pic_list_reordering_flag_l0
    }
    if (idr) {
        bw.put(0, 1);                   // no_output_of_prior_pics
        bw.put(0, 1);                   // long_term_reference_flag
    } else
        bw.put(0, 1);                   // adaptive_ref_pic_marking_mode
    bw.se(slice_qp - 26);
    bw.ue(h.deblock ? 0u : 1u);         // disable_deblocking_filter_idc
    if (h.deblock) {
        // the raw values the reference adds to QP un-doubled (decoder/decoder.c:177-178)
        bw.se(h.alpha_c0_offset);
        bw.se(h.beta_offset);
    }
    std::fill(w->ref4.begin(), w->ref4.end(), (int8_t)-2);
    int last_qp = slice_qp;
    for (int i = 0; i < n_mb; i++) {
        const int r = write_mb(w, bw, fs, i, !idr, slice_qp, &last_qp);
        if (r < 0) {
            w->out.resize(before);
            return r;
        }
    }
    bw.trailing();
    emit_nal(w, idr ? 3 : 2, idr ? 5 : 1, bw);
    w->frame_num++;
    return (int)(w->out.size() - before);
}

const uint8_t *p264b200_writer_data(const p264b200_writer *w, size_t *bytes)
{
    if (!w) return nullptr;
    if (bytes) *bytes = w->out.size();
    return w->out.data();
}

}  // extern "C"
