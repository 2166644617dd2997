// Synthetic FrameSyntax generator for the BASELINE.json configs 3-5 (synthetic 1080p / 4K P-frame
// streams: random quarter-pel MVs over all partition shapes, random residuals, QP sweep 20..40,
// optional multi-reference and deblock-offset sweep).  It produces exactly the buffers the host
// parser would produce for such a stream, so the GPU engine, the CPU oracle and the reference's
// own reconstruction (oracle/ref_harness.c) can all be driven from the same bytes.
//
// The generator mirrors the parts of the host side that shape those buffers:
//   * qp_dbf follows the last-QP rule of core/macroblock.c:1247-1252 (state kept across frames);
//   * intra prediction modes respect neighbour availability (decoder/macroblock.c:635-753);
//   * MVs can be confined so that every referenced sample lies inside the reference's 32-sample
//     border (core/frame.c:42,62-63) -- required when the real reference is the checker.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../../include/p264b200_host.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    int below(int n) { return (int)(next() % (uint64_t)n); }          // [0, n)
    int range(int lo, int hi) { return lo + below(hi - lo + 1); }      // [lo, hi]
    bool pct(int p) { return below(100) < p; }
};

}  // namespace

struct p264b200_synth {
    p264b200_synth_cfg cfg;
    Rng rng;
    int frame = 0;
    int last_qp = 26;
    std::vector<p264b200_mb> mbs;
    std::vector<int16_t> coefs;
    p264b200_synth(const p264b200_synth_cfg &c) : cfg(c), rng(c.seed) {}
};

namespace {

int rand_level(Rng &r, int max_level)
{
    // two-sided geometric: small magnitudes dominate
    int m = 1;
    while (m < max_level && r.pct(40)) m++;
    return r.pct(50) ? m : -m;
}

// fills 16 zig-zag slots with 1..4 non-zero levels among the first `span` positions, from slot `first`
void rand_block(Rng &r, int16_t *dst, int first, int span, int max_level)
{
    memset(dst, 0, 32);
    const int n = r.range(1, 4);
    for (int i = 0; i < n; i++) dst[first + r.below(span)] = (int16_t)rand_level(r, max_level);
    bool any = false;
    for (int i = first; i < 16; i++) any |= dst[i] != 0;
    if (!any) dst[first] = 1;
}

void set_motion(p264b200_mb &m, int bx, int by, int w, int h, int ref, int mvx, int mvy)
{
    for (int y = by; y < by + h; y++)
        for (int x = bx; x < bx + w; x++) {
            m.mv[y * 4 + x][0] = (int16_t)mvx;
            m.mv[y * 4 + x][1] = (int16_t)mvy;
            m.ref[(y >> 1) * 2 + (x >> 1)] = (int8_t)ref;
        }
}

}  // namespace

extern "C" {

void p264b200_synth_default(p264b200_synth_cfg *c, int mb_w, int mb_h)
{
    memset(c, 0, sizeof(*c));
    c->mb_w = mb_w;
    c->mb_h = mb_h;
    c->n_refs = 1;
    c->seed = 264;
    c->qp_min = 20;
    c->qp_max = 40;
    c->qp_step = 2;
    c->coded_pct = 25;
    c->max_level = 8;
    c->mv_range = 16;
    c->sub8x8 = 1;
    c->intra_pct = 0;
    c->skip_pct = 5;
    c->deblock = 1;
    c->confine_mv = 1;
    c->first_intra = 1;
}

p264b200_synth *p264b200_synth_open(const p264b200_synth_cfg *cfg)
{
    if (!cfg || cfg->mb_w < 1 || cfg->mb_h < 1 || cfg->n_refs < 1 || cfg->n_refs > 16 || cfg->qp_min < 0 ||
        cfg->qp_max > 51 || cfg->qp_min > cfg->qp_max || cfg->max_level < 1)
        return nullptr;
    p264b200_synth *s = new (std::nothrow) p264b200_synth(*cfg);
    if (!s) return nullptr;
    s->mbs.resize((size_t)cfg->mb_w * cfg->mb_h);
    s->coefs.reserve((size_t)cfg->mb_w * cfg->mb_h * 64);
    return s;
}

void p264b200_synth_close(p264b200_synth *s) { delete s; }

int p264b200_synth_next(p264b200_synth *s, p264b200_frame_syntax *out)
{
    if (!s || !out) return P264B200_EINVAL;
    const p264b200_synth_cfg &c = s->cfg;
    Rng &r = s->rng;
    const int n_slots = c.n_refs + 1;
    const bool iframe = c.first_intra && (s->frame == 0 || (c.intra_period > 0 && s->frame % c.intra_period == 0));
    const int steps = c.qp_step > 0 ? (c.qp_max - c.qp_min) / c.qp_step + 1 : 1;
    const int qp = c.qp_min + (c.qp_step > 0 ? (s->frame % steps) * c.qp_step : 0);
    const int W = 16 * c.mb_w, H = 16 * c.mb_h;
    // pictures since the last intra picture bound the list (an IDR empties the DPB)
    const int since = c.first_intra && c.intra_period > 0 ? s->frame % c.intra_period : s->frame;
    const int num_ref = iframe ? 0 : (since < c.n_refs ? (since > 0 ? since : 1) : c.n_refs);
    s->coefs.clear();
    int n_intra = 0;

    for (int mby = 0; mby < c.mb_h; mby++)
        for (int mbx = 0; mbx < c.mb_w; mbx++) {
            p264b200_mb &m = s->mbs[(size_t)mby * c.mb_w + mbx];
            memset(&m, 0, sizeof(m));
            m.qp = (uint8_t)qp;
            m.coef_off = (uint32_t)s->coefs.size();
            const bool intra = iframe || r.pct(c.intra_pct);
            bool i16 = false;
            if (intra) {
                n_intra++;
                const bool left = mbx > 0, top = mby > 0;
                i16 = r.pct(50);
                m.mb_type = i16 ? P264B200_MB_I16x16 : P264B200_MB_I4x4;
                // only modes whose neighbours exist (DC always legal)
                int legal16[4], n16 = 0, legalc[4], nc = 0;
                legal16[n16++] = 2;
                legalc[nc++] = 0;
                if (top) legal16[n16++] = 0, legalc[nc++] = 2;
                if (left) legal16[n16++] = 1, legalc[nc++] = 1;
                if (top && left) legal16[n16++] = 3, legalc[nc++] = 3;
                m.i16_mode = (uint8_t)legal16[r.below(n16)];
                m.chroma_mode = (uint8_t)legalc[r.below(nc)];
                if (!i16)
                    for (int b = 0; b < 16; b++) {
                        const int mode = r.below(9);
                        m.i4_mode[b >> 1] |= (uint8_t)(mode << ((b & 1) * 4));
                    }
                for (int i = 0; i < 4; i++) m.ref[i] = -1;
            } else if (r.pct(c.skip_pct)) {
                m.mb_type = P264B200_MB_P_SKIP;
            } else {
                m.mb_type = P264B200_MB_P_L0;
            }
            if (!intra) {
                // partition shape: 16x16, 16x8, 8x16, 8x8 (+ sub shapes when enabled)
                const int shape = m.mb_type == P264B200_MB_P_SKIP ? 0 : r.below(4);
                m.part = (uint8_t)shape;
                if (shape == 3) m.mb_type = P264B200_MB_P_8x8;
                auto draw = [&](int bx, int by, int w, int h, int ref) {
                    int mvx = r.range(-4 * c.mv_range, 4 * c.mv_range + 3);
                    int mvy = r.range(-4 * c.mv_range, 4 * c.mv_range + 3);
                    if (c.confine_mv) {
                        // block origin after the integer shift must stay in [-24, W+24-w]: with the
                        // 6-tap halo (-2..+3) and the second half-pel plane's +1 that is inside +-32
                        const int px = 16 * mbx + 4 * bx, py = 16 * mby + 4 * by;
                        const int lo_x = 4 * (-24 - px), hi_x = 4 * (W + 24 - 4 * w - px);
                        const int lo_y = 4 * (-24 - py), hi_y = 4 * (H + 24 - 4 * h - py);
                        if (mvx < lo_x) mvx = lo_x + (mvx & 3);
                        if (mvx > hi_x) mvx = hi_x - 4 + (mvx & 3);
                        if (mvy < lo_y) mvy = lo_y + (mvy & 3);
                        if (mvy > hi_y) mvy = hi_y - 4 + (mvy & 3);
                    }
                    set_motion(m, bx, by, w, h, ref, mvx, mvy);
                };
                if (shape == 0)
                    draw(0, 0, 4, 4, m.mb_type == P264B200_MB_P_SKIP ? 0 : r.below(num_ref));
                else if (shape == 1)
                    for (int i = 0; i < 2; i++) draw(0, 2 * i, 4, 2, r.below(num_ref));
                else if (shape == 2)
                    for (int i = 0; i < 2; i++) draw(2 * i, 0, 2, 4, r.below(num_ref));
                else
                    for (int i = 0; i < 4; i++) {
                        const int ox = 2 * (i & 1), oy = 2 * (i >> 1), ref = r.below(num_ref);
                        const int sub = c.sub8x8 ? r.below(4) : 0;
                        m.sub_part[i] = (uint8_t)sub;
                        if (sub == P264B200_SUB_8x8)
                            draw(ox, oy, 2, 2, ref);
                        else if (sub == P264B200_SUB_8x4)
                            for (int j = 0; j < 2; j++) draw(ox, oy + j, 2, 1, ref);
                        else if (sub == P264B200_SUB_4x8)
                            for (int j = 0; j < 2; j++) draw(ox + j, oy, 1, 2, ref);
                        else
                            for (int j = 0; j < 4; j++) draw(ox + (j & 1), oy + (j >> 1), 1, 1, ref);
                    }
            }
            // residual
            int cbp_luma_any = 0;
            if (m.mb_type != P264B200_MB_P_SKIP) {
                int16_t blk[16];
                if (i16) {
                    rand_block(r, blk, 0, 6, c.max_level);
                    s->coefs.insert(s->coefs.end(), blk, blk + 16);
                }
                for (int b = 0; b < 16; b++)
                    if (r.pct(c.coded_pct)) {
                        rand_block(r, blk, i16 ? 1 : 0, 8, c.max_level);
                        s->coefs.insert(s->coefs.end(), blk, blk + 16);
                        m.luma_mask |= (uint16_t)(1 << b);
                        cbp_luma_any = 1;
                    }
                const int pc = r.below(100);
                m.cbp_chroma = (uint8_t)(pc < c.coded_pct ? 2 : pc < 2 * c.coded_pct ? 1 : 0);
                if (m.cbp_chroma) {
                    int16_t dc[8];
                    for (int i = 0; i < 8; i++) dc[i] = (int16_t)(r.pct(50) ? rand_level(r, c.max_level) : 0);
                    s->coefs.insert(s->coefs.end(), dc, dc + 8);
                    if (m.cbp_chroma == 2)
                        for (int i = 0; i < 8; i++)
                            if (r.pct(50)) {
                                rand_block(r, blk, 1, 6, c.max_level);
                                s->coefs.insert(s->coefs.end(), blk, blk + 16);
                                m.chroma_mask |= (uint8_t)(1 << i);
                            }
                }
            }
            // last-QP rule for the deblocker (core/macroblock.c:1247-1252)
            if (!i16 && !cbp_luma_any && m.cbp_chroma == 0)
                m.qp_dbf = (uint8_t)s->last_qp;
            else
                m.qp_dbf = m.qp;
            s->last_qp = m.qp_dbf;
        }

    p264b200_frame_hdr &h = out->hdr;
    memset(&h, 0, sizeof(h));
    h.mb_w = c.mb_w;
    h.mb_h = c.mb_h;
    h.slice_type = iframe ? P264B200_SLICE_I : P264B200_SLICE_P;
    h.deblock = c.deblock;
    if (c.sweep_offsets) {
        h.alpha_c0_offset = r.range(-6, 6);
        h.beta_offset = r.range(-6, 6);
    }
    h.chroma_qp_index_offset = c.chroma_qp_index_offset;
    h.num_ref = num_ref;
    for (int i = 0; i < num_ref; i++) h.ref_slot[i] = ((s->frame - 1 - i) % n_slots + n_slots) % n_slots;
    h.dst_slot = s->frame % n_slots;
    h.n_intra = n_intra;
    if (s->coefs.empty()) s->coefs.resize(8, 0);
    h.n_coef = (uint32_t)s->coefs.size();
    out->mbs = s->mbs.data();
    out->coefs = s->coefs.data();
    s->frame++;
    return P264B200_OK;
}

}  // extern "C"
