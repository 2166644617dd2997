// Host syntax front-end, see parser.h.  Baseline profile: CAVLC, I and P slices, one slice
// per picture, frame pictures -- the feature set the reference decodes.  Where the
// reference deviates from the standard in a way that changes the reconstructed picture of
// a stream it CAN decode (QP not accumulated, last-QP leak, un-doubled deblock offsets,
// emulation-prevention boundary) the deviation is mirrored; its parser bugs on syntax it
// cannot decode meaningfully (multi-ref partitions, sub-8x8 MV storage) are not: those
// follow the standard (SURVEY.md 8a "Quirks").
#include "parser.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cavlc.h"

namespace p264b200 {

namespace {

// Table 9-4 coded_block_pattern mapping for me(v), Intra4x4 / Inter columns
const uint8_t kCbpIntra[48] = {47, 31, 15, 0,  23, 27, 29, 30, 7,  11, 13, 14, 39, 43, 45, 46,
                               16, 3,  5,  10, 12, 19, 21, 26, 28, 35, 37, 42, 44, 1,  2,  4,
                               8,  17, 18, 20, 24, 6,  9,  22, 25, 32, 33, 34, 36, 40, 38, 41};
const uint8_t kCbpInter[48] = {0,  16, 1,  2,  4,  8,  32, 3,  5,  10, 12, 15, 47, 7,  11, 13,
                               14, 6,  9,  31, 35, 37, 42, 44, 33, 34, 36, 40, 39, 43, 45, 46,
                               17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41};
// 4x4 block index (bitstream order) -> position inside the MB (6.4.3 inverse 4x4 luma block scan)
const uint8_t kZx[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
const uint8_t kZy[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};

inline int median3(int a, int b, int c)
{
    int mn = std::min(a, std::min(b, c)), mx = std::max(a, std::max(b, c));
    return a + b + c - mn - mx;
}

void *default_alloc(size_t n) { return malloc(n); }
void default_free(void *p) { free(p); }

}  // namespace

int nal_unescape(const uint8_t *src, int size, uint8_t *dst, int *nal_type, int *nal_ref_idc)
{
    if (size < 1) return -1;
    const uint8_t *end = src + size;
    uint8_t *d = dst;
    *nal_type = src[0] & 0x1f;
    *nal_ref_idc = (src[0] >> 5) & 3;
    src++;
    while (src < end) {
        if (src < end - 3 && src[0] == 0 && src[1] == 0 && src[2] == 3) {
            *d++ = 0;
            *d++ = 0;
            src += 3;
            continue;
        }
        *d++ = *src++;
    }
    return (int)(d - dst);
}

Parser::Parser(alloc_fn a, free_fn f) : alloc_(a ? a : default_alloc), free_(f ? f : default_free)
{
    cavlc_init();
    memset(&hdr_, 0, sizeof(hdr_));
    memset(list0_, 0, sizeof(list0_));
}

Parser::~Parser()
{
    if (mbs_) free_(mbs_);
    if (coefs_) free_(coefs_);
}

// ---------------------------------------------------------------- parameter sets
// decoder/set.c:37-168
int Parser::read_sps(BitReader &br)
{
    Sps s;
    s.profile_idc = br.read(8);
    br.read(3);  // constraint_set0..2
    br.skip(5);
    s.level_idc = br.read(8);
    int id = br.ue();
    if (br.eof() || id < 0 || id >= 32) return P264B200_EBITSTREAM;
    s.id = id;
    // ue() returns -1 (or a huge value) on a damaged code: every field is range-checked before it sizes a shift or a loop
    const int l2fn = br.ue();
    if (l2fn < 0 || l2fn > 12) return P264B200_EBITSTREAM;
    s.log2_max_frame_num = l2fn + 4;
    s.poc_type = br.ue();
    if (s.poc_type < 0) return P264B200_EBITSTREAM;
    if (s.poc_type == 0) {
        const int l2poc = br.ue();
        if (l2poc < 0 || l2poc > 12) return P264B200_EBITSTREAM;
        s.log2_max_poc_lsb = l2poc + 4;
    } else if (s.poc_type == 1) {
        s.delta_pic_order_always_zero = br.read1();
        br.se();
        br.se();
        int n = br.ue();
        if (n < 0 || n > 255) return P264B200_EBITSTREAM;
        for (int i = 0; i < n; i++) br.se();
    } else if (s.poc_type > 2) {
        sps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    s.num_ref_frames = br.ue();
    br.read1();  // gaps_in_frame_num_value_allowed
    s.mb_w = br.ue() + 1;
    s.mb_h = br.ue() + 1;
    s.frame_mbs_only = br.read1();
    if (!s.frame_mbs_only) br.read1();
    br.read1();  // direct_8x8_inference
    if (br.read1()) {
        for (int i = 0; i < 4; i++) {
            s.crop[i] = br.ue();
            if (s.crop[i] < 0) return P264B200_EBITSTREAM;
        }
    }
    br.read1();  // vui_parameters_present: skipped like decoder/set.c:136-144
    if (br.eof()) {
        fprintf(stderr, "incomplete SPS\n");
        sps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    if (s.num_ref_frames < 0 || s.num_ref_frames > 16 || s.mb_w <= 0 || s.mb_h <= 0 || s.mb_w > 1024 ||
        s.mb_h > 1024) {
        sps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    sps_[id] = s;
    if (verbose)
        fprintf(stderr, "p264_sps_read: sps:0x%x profile:%d/%d poc:%d ref:%d %xx%d crop:%d-%d-%d-%d\n", s.id,
                s.profile_idc, s.level_idc, s.poc_type, s.num_ref_frames, s.mb_w, s.mb_h, s.crop[0], s.crop[1],
                s.crop[2], s.crop[3]);
    return id;
}

// decoder/set.c:171-270
int Parser::read_pps(BitReader &br)
{
    Pps p;
    int id = br.ue();
    if (br.eof() || id < 0 || id >= 256) {
        fprintf(stderr, "id invalid\n");
        return P264B200_EBITSTREAM;
    }
    p.id = id;
    p.sps_id = br.ue();
    if (p.sps_id < 0 || p.sps_id >= 32) {
        pps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    p.cabac = br.read1();
    p.pic_order = br.read1();
    p.num_slice_groups = br.ue() + 1;
    if (p.num_slice_groups < 1 || p.num_slice_groups > 8) {
        pps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    if (p.num_slice_groups > 1) {
        fprintf(stderr, "FMO unsupported\n ");
        p.slice_group_map_type = br.ue();
        if (p.slice_group_map_type == 0) {
            for (int i = 0; i < p.num_slice_groups; i++) br.ue();
        } else if (p.slice_group_map_type == 2) {
            for (int i = 0; i < p.num_slice_groups; i++) {
                br.ue();
                br.ue();
            }
        } else if (p.slice_group_map_type >= 3 && p.slice_group_map_type <= 5) {
            br.read1();
            br.ue();
        } else if (p.slice_group_map_type == 6) {
            br.ue();
        }
    }
    p.num_ref_idx_l0 = br.ue() + 1;
    p.num_ref_idx_l1 = br.ue() + 1;
    if (p.num_ref_idx_l0 < 1 || p.num_ref_idx_l0 > 32 || p.num_ref_idx_l1 < 1 || p.num_ref_idx_l1 > 32) {
        pps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    p.weighted_pred = br.read1();
    p.weighted_bipred = br.read(2);
    p.pic_init_qp = br.se() + 26;
    p.pic_init_qs = br.se() + 26;
    p.chroma_qp_index_offset = br.se();
    p.deblocking_filter_control = br.read1();
    p.constrained_intra_pred = br.read1();
    p.redundant_pic_cnt = br.read1();
    if (br.eof()) {
        fprintf(stderr, "incomplete PPS\n");
        pps_[id].id = -1;
        return P264B200_EBITSTREAM;
    }
    pps_[id] = p;
    if (verbose)
        fprintf(stderr,
                "p264_sps_read: pps:0x%x sps:0x%x %s slice_groups=%d ref0:%d ref1:%d QP:%d QS:%d QC=%d DFC:%d CIP:%d "
                "RPC:%d\n",
                p.id, p.sps_id, p.cabac ? "CABAC" : "CAVLC", p.num_slice_groups, p.num_ref_idx_l0, p.num_ref_idx_l1,
                p.pic_init_qp, p.pic_init_qs, p.chroma_qp_index_offset, p.deblocking_filter_control,
                p.constrained_intra_pred, p.redundant_pic_cnt);
    return id;
}

// ------------------------------------------------------------------ slice header
// decoder/decoder.c:70-301
int Parser::slice_header(BitReader &br, int nal_type, int nal_ref_idc, SliceHeader &sh)
{
    const bool idr = nal_type == 5;
    sh.first_mb = br.ue();
    sh.type = br.ue();
    if (sh.type >= 5) sh.type -= 5;
    sh.pps_id = br.ue();
    if (br.eof() || sh.pps_id < 0 || sh.pps_id >= 256 || pps_[sh.pps_id].id == -1) {
        fprintf(stderr, "invalid pps_id %d in slice header\n", sh.pps_id);
        return P264B200_EBITSTREAM;
    }
    const Pps *pps = &pps_[sh.pps_id];
    const Sps *sps = &sps_[pps->sps_id];
    if (sps->id == -1) return P264B200_EBITSTREAM;

    sh.frame_num = br.read(sps->log2_max_frame_num);
    if (!sps->frame_mbs_only) {
        sh.field_pic = br.read1();
        if (sh.field_pic) br.read1();
    }
    sh.idr_pic_id = idr ? br.ue() : 0;
    if (sps->poc_type == 0) {
        br.read(sps->log2_max_poc_lsb);
        if (pps->pic_order && !sh.field_pic) br.se();
    } else if (sps->poc_type == 1 && !sps->delta_pic_order_always_zero) {
        br.se();
        if (pps->pic_order && !sh.field_pic) br.se();
    }
    if (pps->redundant_pic_cnt) sh.redundant_pic_cnt = br.ue();
    if (sh.type == 1) br.read1();  // direct_spatial_mv_pred
    sh.num_ref_idx_l0_active = pps->num_ref_idx_l0;
    if (sh.type == 0 || sh.type == 3 || sh.type == 1) {
        if (br.read1()) {
            sh.num_ref_idx_l0_active = br.ue() + 1;
            if (sh.type == 1) br.ue();
        }
    }
    if (br.eof()) return P264B200_EBITSTREAM;

    if (sh.type != 0 && sh.type != 2) {
        fprintf(stderr, "slice unsupported yet \n");
        return P264B200_EUNSUP;
    }
    if (pps->cabac || sh.field_pic || !sps->frame_mbs_only) return P264B200_EUNSUP;
    if (sh.type == 0 && pps->weighted_pred) return P264B200_EUNSUP;  // pred_weight_table is never parsed
    if (sh.num_ref_idx_l0_active < 1 || sh.num_ref_idx_l0_active > 16) return P264B200_EBITSTREAM;

    // ref_pic_list_reordering: parsed and ignored (decoder/decoder.c:196-257, decoder/lists.c:146-149)
    if (sh.type != 2) {
        if (br.read1()) {
            for (int guard = 0; guard < 64; guard++) {
                int idc = br.ue();
                if (idc == 3) break;
                if (idc > 3 || idc < 0) {
                    fprintf(stderr, "wrong reordering of pic nums idc\n");
                    return P264B200_EBITSTREAM;
                }
                br.ue();
            }
        }
    }
    // dec_ref_pic_marking (decoder/decoder.c:265-301)
    if (nal_ref_idc != 0) {
        if (idr) {
            sh.no_output_of_prior_pics = br.read1();
            sh.long_term_reference_flag = br.read1();
        } else {
            sh.adaptive_ref_pic_marking = br.read1();
            if (sh.adaptive_ref_pic_marking) {
                for (int guard = 0; guard < 64; guard++) {
                    int cmd = br.ue();
                    if (cmd == 0) break;
                    if (cmd > 6 || cmd < 0) {
                        fprintf(stderr, "wrong memory mangement control operation\n");
                        break;
                    }
                    if (cmd != 5) br.ue();
                }
            }
        }
    }
    sh.qp_delta = br.se();
    if (pps->deblocking_filter_control) {
        sh.disable_deblocking_filter_idc = br.ue();
        if (sh.disable_deblocking_filter_idc != 1) {
            sh.alpha_c0_offset = br.se();  // NOT doubled: decoder/decoder.c:177-178 + core/frame.c:476-478
            sh.beta_offset = br.se();
        }
    }
    if (pps->num_slice_groups > 1 && pps->slice_group_map_type >= 3 && pps->slice_group_map_type <= 5)
        return P264B200_EUNSUP;
    if (pps->num_slice_groups > 1) return P264B200_EUNSUP;

    // activate parameter sets (decoder/decoder.c:379-398: re-init whenever either pointer changes)
    if (asps_ != sps || apps_ != pps || mb_w_ != sps->mb_w || mb_h_ != sps->mb_h ||
        ring_n_ != sps->num_ref_frames + 1) {
        asps_ = sps;
        apps_ = pps;
        if (context_init()) return P264B200_ENOMEM;
    }
    return 0;
}

// decoder/decoder.c:304-343
int Parser::context_init()
{
    mb_w_ = asps_->mb_w;
    mb_h_ = asps_->mb_h;
    if (verbose) fprintf(stderr, "p264: %dx%d\n", 16 * mb_w_, 16 * mb_h_);
    const size_t n = (size_t)mb_w_ * mb_h_;
    if (n > mbs_cap_) {
        if (mbs_) free_(mbs_);
        mbs_ = (p264b200_mb *)alloc_(n * sizeof(p264b200_mb));
        mbs_cap_ = mbs_ ? n : 0;
        if (!mbs_) return P264B200_ENOMEM;
    }
    ring_n_ = asps_->num_ref_frames + 1;
    ring_.assign(ring_n_, RingEntry{0, 0, -1, 0});
    for (int i = 0; i < ring_n_; i++) ring_[i].slot = i;
    ring_used_ = 0;
    nnz_y_.assign(n * 16, 0);
    nnz_c_[0].assign(n * 4, 0);
    nnz_c_[1].assign(n * 4, 0);
    imode_.assign(n * 16, 2);
    ref4_.assign(n * 16, -2);
    mv4_.assign(n * 32, 0);
    geometry_changed_ = true;
    return 0;
}

int Parser::ensure_coef(size_t need)
{
    if (need <= coef_cap_) return 0;
    size_t cap = std::max(need, coef_cap_ ? coef_cap_ * 2 : (size_t)mb_w_ * mb_h_ * 64 + 4096);
    int16_t *p = (int16_t *)alloc_(cap * sizeof(int16_t));
    if (!p) return P264B200_ENOMEM;
    if (coefs_) {
        memcpy(p, coefs_, coef_n_ * sizeof(int16_t));
        free_(coefs_);
    }
    coefs_ = p;
    coef_cap_ = cap;
    return 0;
}

// decoder/lists.c:72-143: list 0 = short-term references by descending PicNum, then long-term
void Parser::lists_init(const SliceHeader &sh)
{
    n_list0_ = 0;
    if (sh.type == 2) return;
    const int max_frame_num = 1 << asps_->log2_max_frame_num;
    std::vector<int> shorts, longs;
    for (int i = 1; i < ring_n_; i++) {
        RingEntry &r = ring_[i];
        if (r.ref_type == 1) {
            r.pic_num = r.frame_num > sh.frame_num ? r.frame_num - max_frame_num : r.frame_num;
            shorts.push_back(i);
        } else if (r.ref_type == 2)
            longs.push_back(i);
    }
    std::stable_sort(shorts.begin(), shorts.end(), [&](int a, int b) { return ring_[a].pic_num > ring_[b].pic_num; });
    for (int i : shorts)
        if (n_list0_ < 16) list0_[n_list0_++] = i;
    for (int i : longs)
        if (n_list0_ < 16) list0_[n_list0_++] = i;
}

// decoder/lists.c:152-228 sliding window + ring rotation
void Parser::marking(int nal_type, const SliceHeader &sh)
{
    if (nal_type == 5) {
        if (sh.no_output_of_prior_pics)
            for (int i = 1; i < ring_n_; i++) ring_[i].ref_type = 0;
        ring_[0].ref_type = sh.long_term_reference_flag ? 2 : 1;
        if (ring_n_ > 1) std::swap(ring_[0], ring_[1]);
        ring_used_ = 2;
        return;
    }
    if (sh.adaptive_ref_pic_marking) {
        // the reference prints and returns before rotating (decoder/lists.c:183-187)
        printf("not support adaptive ref marking yet");
        return;
    }
    ring_[0].ref_type = 1;
    int i;
    if (ring_used_ < ring_n_) {
        i = ring_used_++;
    } else {
        for (i = ring_used_ - 1; i >= 0; i--)
            if (ring_[i].ref_type == 1) {
                ring_[i].ref_type = 0;
                break;
            }
        if (i < 0) i = 0;
    }
    RingEntry fdec = ring_[i];
    for (int j = i; j > 0; j--) ring_[j] = ring_[j - 1];
    ring_[0] = fdec;
}

// ------------------------------------------------------------------- predictors
int Parser::predict_nnz(const uint8_t *grid, int stride, int x, int y) const
{
    // core/macroblock.c:53-65: average of left/top counts, unavailable neighbours dropped
    const bool a = x > 0, b = y > 0;
    const int na = a ? grid[y * stride + x - 1] : 0, nb = b ? grid[(y - 1) * stride + x] : 0;
    if (a && b) return (na + nb + 1) >> 1;
    return a ? na : (b ? nb : 0);
}

// H.264 8.4.1.3 / core/macroblock.c:87-175.  (x4,y4) picture-wide 4x4 coordinates of the partition's
// top-left block, w4 its width; shape: 0 none, 1 = 16x8, 2 = 8x16 (directional rules).
void Parser::predict_mv(int x4, int y4, int w4, int ref, int shape, int part_idx, int mvp[2]) const
{
    const int s4 = 4 * mb_w_;
    auto cell_ref = [&](int x, int y) -> int {
        if (x < 0 || y < 0 || x >= s4 || y >= 4 * mb_h_) return -2;
        return ref4_[y * s4 + x];
    };
    auto cell_mv = [&](int x, int y, int c) -> int {
        if (x < 0 || y < 0 || x >= s4 || y >= 4 * mb_h_) return 0;
        return ref4_[y * s4 + x] == -2 ? 0 : mv4_[(y * s4 + x) * 2 + c];
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

void Parser::fill_motion(p264b200_mb &m, int mbx, int mby, int bx, int by, int w, int h, int ref, int mvx, int mvy)
{
    const int s4 = 4 * mb_w_;
    for (int y = by; y < by + h; y++)
        for (int x = bx; x < bx + w; x++) {
            m.mv[y * 4 + x][0] = (int16_t)mvx;
            m.mv[y * 4 + x][1] = (int16_t)mvy;
            m.ref[(y >> 1) * 2 + (x >> 1)] = (int8_t)ref;
            const int g = (4 * mby + y) * s4 + 4 * mbx + x;
            ref4_[g] = (int8_t)ref;
            mv4_[g * 2] = (int16_t)mvx;
            mv4_[g * 2 + 1] = (int16_t)mvy;
        }
}

// ---------------------------------------------------------------- macroblock layer
// decoder/macroblock.c:265-301 + core/macroblock.c:40-51
int Parser::mb_intra_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, bool i4x4)
{
    const int s4 = 4 * mb_w_;
    if (i4x4) {
        memset(m.i4_mode, 0, sizeof(m.i4_mode));
        for (int i = 0; i < 16; i++) {
            const int x = 4 * mbx + kZx[i], y = 4 * mby + kZy[i];
            const int ma = x > 0 ? imode_[y * s4 + x - 1] : -1;
            const int mb = y > 0 ? imode_[(y - 1) * s4 + x] : -1;
            int pred = std::min(ma, mb);
            if (pred < 0) pred = 2;
            int mode;
            if (br.read1())
                mode = pred;
            else {
                const int rem = br.read(3);
                mode = rem >= pred ? rem + 1 : rem;
            }
            imode_[y * s4 + x] = (int8_t)mode;
            const int b = kZy[i] * 4 + kZx[i];
            m.i4_mode[b >> 1] |= (uint8_t)(mode << ((b & 1) * 4));
        }
    }
    const int cm = br.ue();
    if (cm < 0 || cm > 3) return P264B200_EBITSTREAM;
    m.chroma_mode = (uint8_t)cm;
    // modes that would read samples the picture does not have: the reference reads stale
    // padding there (undefined content), so such streams are rejected instead of guessed
    const bool left = mbx > 0, top = mby > 0;
    if ((cm == 1 && !left) || (cm == 2 && !top) || (cm == 3 && !(left && top))) return P264B200_EBITSTREAM;
    if (!i4x4) {
        const int lm = m.i16_mode;
        if ((lm == 0 && !top) || (lm == 1 && !left) || (lm == 3 && !(left && top))) return P264B200_EBITSTREAM;
    }
    return 0;
}

// decoder/macroblock.c:304-339 (P_L0 16x16 / 16x8 / 8x16)
int Parser::mb_inter_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, const SliceHeader &sh)
{
    const int nparts = m.part == P264B200_D_16x16 ? 1 : 2;
    const int w = m.part == P264B200_D_8x16 ? 2 : 4, h = m.part == P264B200_D_16x8 ? 2 : 4;
    int refs[2] = {0, 0};
    if (sh.num_ref_idx_l0_active > 1)
        for (int i = 0; i < nparts; i++) {
            refs[i] = br.te(sh.num_ref_idx_l0_active - 1);
            if (refs[i] < 0 || refs[i] >= n_list0_) return P264B200_EBITSTREAM;
        }
    for (int i = 0; i < nparts; i++) {
        const int bx = m.part == P264B200_D_8x16 ? 2 * i : 0, by = m.part == P264B200_D_16x8 ? 2 * i : 0;
        int mvp[2];
        const int shape = m.part == P264B200_D_16x8 ? 1 : (m.part == P264B200_D_8x16 ? 2 : 0);
        predict_mv(4 * mbx + bx, 4 * mby + by, w, refs[i], shape, i, mvp);
        const int mvx = mvp[0] + br.se(), mvy = mvp[1] + br.se();
        fill_motion(m, mbx, mby, bx, by, w, h, refs[i], mvx, mvy);
    }
    return 0;
}

// decoder/macroblock.c:342-408 (P_8x8), sub-partition MV storage per the standard
int Parser::mb_sub_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, const SliceHeader &sh, bool ref0)
{
    for (int i = 0; i < 4; i++) {
        const int t = br.ue();
        if (t < 0 || t > 3) {
            fprintf(stderr, "invalid i_sub_partition\n");
            return P264B200_EBITSTREAM;
        }
        m.sub_part[i] = (uint8_t)t;  // 0 8x8, 1 8x4, 2 4x8, 3 4x4 == P264B200_SUB_*
    }
    int refs[4] = {0, 0, 0, 0};
    if (sh.num_ref_idx_l0_active > 1 && !ref0)
        for (int i = 0; i < 4; i++) {
            refs[i] = br.te(sh.num_ref_idx_l0_active - 1);
            if (refs[i] < 0 || refs[i] >= n_list0_) return P264B200_EBITSTREAM;
        }
    for (int i = 0; i < 4; i++) {
        const int ox = 2 * (i & 1), oy = 2 * (i >> 1);
        const int sw = (m.sub_part[i] == P264B200_SUB_8x8 || m.sub_part[i] == P264B200_SUB_8x4) ? 2 : 1;
        const int sh4 = (m.sub_part[i] == P264B200_SUB_8x8 || m.sub_part[i] == P264B200_SUB_4x8) ? 2 : 1;
        const int n = (2 / sw) * (2 / sh4);
        for (int j = 0; j < n; j++) {
            int bx, by;
            if (sw == 2)
                bx = ox, by = oy + j;  // 8x8 (j=0) or 8x4
            else if (sh4 == 2)
                bx = ox + j, by = oy;  // 4x8
            else
                bx = ox + (j & 1), by = oy + (j >> 1);
            int mvp[2];
            predict_mv(4 * mbx + bx, 4 * mby + by, sw, refs[i], 0, 0, mvp);
            const int mvx = mvp[0] + br.se(), mvy = mvp[1] + br.se();
            fill_motion(m, mbx, mby, bx, by, sw, sh4, refs[i], mvx, mvy);
        }
    }
    return 0;
}

// decoder/macroblock.c:895-933 + core/macroblock.c:235-252
void Parser::mb_skip(p264b200_mb &m, int mbx, int mby)
{
    const int s4 = 4 * mb_w_, x4 = 4 * mbx, y4 = 4 * mby;
    int mv[2] = {0, 0};
    const int ra = x4 > 0 ? ref4_[y4 * s4 + x4 - 1] : -2;
    const int rb = y4 > 0 ? ref4_[(y4 - 1) * s4 + x4] : -2;
    bool zero = ra == -2 || rb == -2;
    if (!zero) {
        const int16_t *ma = &mv4_[(y4 * s4 + x4 - 1) * 2], *mb = &mv4_[((y4 - 1) * s4 + x4) * 2];
        zero = (ra == 0 && ma[0] == 0 && ma[1] == 0) || (rb == 0 && mb[0] == 0 && mb[1] == 0);
    }
    if (!zero) predict_mv(x4, y4, 4, 0, 0, 0, mv);
    m.mb_type = P264B200_MB_P_SKIP;
    m.part = P264B200_D_16x16;
    fill_motion(m, mbx, mby, 0, 0, 4, 4, 0, mv[0], mv[1]);
}

// decoder/macroblock.c:410-486: CAVLC residual of one MB, packed into the coefficient stream
int Parser::mb_residual(BitReader &br_in, p264b200_mb &m, int mbx, int mby, int cbp_luma)
{
    // (the reader is worked on as a local copy: its position and cached window then live in registers across the ~24 one-bit
    // "empty block" probes of a macroblock instead of going through memory each time; written back on every exit)
    BitReader br = br_in;
    struct WriteBack {
        BitReader &dst;
        const BitReader &src;
        ~WriteBack() { dst = src; }
    } write_back{br_in, br};
    const int s4 = 4 * mb_w_, s2 = 2 * mb_w_;
    int16_t luma[16][16];  // raster block index
    int16_t dc[16], cdc[2][4], cac[8][16];
    int tot_luma[16];
    // (blocks are cleared only when they are about to be read: three quarters of the 8x8s carry nothing)
    const bool i16 = m.mb_type == P264B200_MB_I16x16;

    if (i16) {
        memset(dc, 0, sizeof(dc));
        const int nC = predict_nnz(nnz_y_.data(), s4, 4 * mbx, 4 * mby);
        if (cavlc_read_block(br, nC, 16, dc) < 0) return P264B200_EBITSTREAM;
    }
    {
        // non-zero counts of the 4x4 neighbourhood: nz[(y + 1) * 5 + (x + 1)], row / column -1 = the macroblocks above / to the left
        // (-1 = unavailable), so that the predictor of a block is two byte loads from the stack
        int8_t nz[25];
        uint8_t *grid = nnz_y_.data() + (size_t)(4 * mby) * s4 + 4 * mbx;
        for (int k = 0; k < 4; k++) {
            nz[k + 1] = mby > 0 ? (int8_t)grid[-s4 + k] : (int8_t)-1;
            nz[5 * (k + 1)] = mbx > 0 ? (int8_t)grid[k * s4 - 1] : (int8_t)-1;
        }
        for (int i = 0; i < 16; i++) {
            const int bx = kZx[i], by = kZy[i], b = by * 4 + bx;
            int tot = 0;
            if (cbp_luma & (1 << (i / 4))) {
                const int na = nz[5 * (by + 1) + bx], nb = nz[5 * by + bx + 1];
                const int nC = (na >= 0 && nb >= 0) ? (na + nb + 1) >> 1 : na >= 0 ? na : nb >= 0 ? nb : 0;
                if (!cavlc_skip_empty(br, nC)) {
                    memset(luma[b], 0, sizeof(luma[b]));
                    tot = i16 ? cavlc_read_block(br, nC, 15, luma[b] + 1) : cavlc_read_block(br, nC, 16, luma[b]);
                    if (tot < 0) return P264B200_EBITSTREAM;
                }
            }
            nz[5 * (by + 1) + bx + 1] = (int8_t)tot;
            tot_luma[b] = tot;
        }
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) grid[y * s4 + x] = (uint8_t)nz[5 * (y + 1) + x + 1];
    }
    int tot_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (m.cbp_chroma & 3) {
        memset(cdc, 0, sizeof(cdc));
        if (cavlc_read_block(br, -1, 4, cdc[0]) < 0 || cavlc_read_block(br, -1, 4, cdc[1]) < 0)
            return P264B200_EBITSTREAM;
    }
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < 4; i++) {
            const int gx = 2 * mbx + (i & 1), gy = 2 * mby + (i >> 1);
            int tot = 0;
            if (m.cbp_chroma & 2) {
                const int nC = predict_nnz(nnz_c_[c].data(), s2, gx, gy);
                if (!cavlc_skip_empty(br, nC)) {
                    memset(cac[c * 4 + i], 0, sizeof(cac[0]));
                    tot = cavlc_read_block(br, nC, 15, cac[c * 4 + i] + 1);
                    if (tot < 0) return P264B200_EBITSTREAM;
                }
            }
            nnz_c_[c][gy * s2 + gx] = (uint8_t)tot;
            tot_c[c * 4 + i] = tot;
        }

    // pack (layout documented in include/p264b200_recon.h)
    if (ensure_coef(coef_n_ + 16 + 16 * 16 + 8 + 8 * 16)) return P264B200_ENOMEM;
    m.coef_off = (uint32_t)coef_n_;
    int16_t *o = coefs_ + coef_n_;
    if (i16) {
        memcpy(o, dc, 32);
        o += 16;
    }
    m.luma_mask = 0;
    for (int b = 0; b < 16; b++)
        if (tot_luma[b] > 0) {
            m.luma_mask |= (uint16_t)(1 << b);
            memcpy(o, luma[b], 32);
            o += 16;
        }
    m.chroma_mask = 0;
    if (m.cbp_chroma) {
        memcpy(o, cdc[0], 8);
        memcpy(o + 4, cdc[1], 8);
        o += 8;
        for (int i = 0; i < 8; i++)
            if (tot_c[i] > 0) {
                m.chroma_mask |= (uint8_t)(1 << i);
                memcpy(o, cac[i], 32);
                o += 16;
            }
    }
    coef_n_ = (size_t)(o - coefs_);
    return 0;
}

// the per-frame side arrays of p264_macroblock_cache_save (core/macroblock.c:1234-1340)
void Parser::mb_finish(p264b200_mb &m, int mbx, int mby, int cbp_luma)
{
    const int s4 = 4 * mb_w_, s2 = 2 * mb_w_;
    const bool intra = P264B200_IS_INTRA(m.mb_type);
    if (m.mb_type != P264B200_MB_I16x16 && cbp_luma == 0 && m.cbp_chroma == 0) {
        m.qp_dbf = (uint8_t)last_qp_;
        // no residual was parsed: neighbours see zero counts (decoder/macroblock.c:576-587)
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) nnz_y_[(4 * mby + y) * s4 + 4 * mbx + x] = 0;
        for (int c = 0; c < 2; c++)
            for (int i = 0; i < 4; i++) nnz_c_[c][(2 * mby + (i >> 1)) * s2 + 2 * mbx + (i & 1)] = 0;
        m.luma_mask = 0;
        m.chroma_mask = 0;
        m.coef_off = (uint32_t)coef_n_;
    } else
        m.qp_dbf = m.qp;
    last_qp_ = m.qp_dbf;
    if (m.mb_type != P264B200_MB_I4x4)
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) imode_[(4 * mby + y) * s4 + 4 * mbx + x] = 2;
    if (intra) {
        n_intra_++;
        for (int b = 0; b < 16; b++) m.mv[b][0] = m.mv[b][1] = 0;
        for (int i = 0; i < 4; i++) m.ref[i] = -1;
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) {
                const int g = (4 * mby + y) * s4 + 4 * mbx + x;
                ref4_[g] = -1;
                mv4_[g * 2] = mv4_[g * 2 + 1] = 0;
            }
    }
}

// decoder/decoder.c:502-593 + decoder/macroblock.c:488-592
int Parser::slice_data(BitReader &br, const SliceHeader &sh)
{
    const int n_mb = mb_w_ * mb_h_;
    std::fill(ref4_.begin(), ref4_.end(), (int8_t)-2);
    coef_n_ = 0;
    n_intra_ = 0;
    int skip_run = -1;
    const int base_qp = apps_->pic_init_qp + sh.qp_delta;
    for (int mb_xy = 0; mb_xy < n_mb; mb_xy++) {
        const int mbx = mb_xy % mb_w_, mby = mb_xy / mb_w_;
        p264b200_mb &m = mbs_[mb_xy];
        memset(&m, 0, sizeof(m));
        bool have_mb = false;
        if (skip_run < 1) {
            bool read_type = true;
            if (sh.type != 2 && skip_run == -1) {
                skip_run = br.ue();
                if (skip_run < 0) return P264B200_EBITSTREAM;
                if (skip_run > 0) read_type = false;
            }
            if (read_type) {
                have_mb = true;
                int t = br.ue();
                if (t < 0) return P264B200_EBITSTREAM;
                bool intra = true;
                if (sh.type == 0) {
                    if (t < 5)
                        intra = false;
                    else
                        t -= 5;
                }
                int cbp_luma = 0;
                bool ref0 = false;
                if (intra) {
                    if (t == 0)
                        m.mb_type = P264B200_MB_I4x4;
                    else if (t < 25) {
                        m.mb_type = P264B200_MB_I16x16;
                        m.i16_mode = (uint8_t)((t - 1) % 4);
                        m.cbp_chroma = (uint8_t)(((t - 1) / 4) % 3);
                        cbp_luma = t > 12 ? 15 : 0;
                    } else if (t == 25) {
                        fprintf(stderr, "unsupport i_pcm mb\n");
                        return P264B200_EUNSUP;
                    } else {
                        fprintf(stderr, "invalid mb type %d \n", t);
                        return P264B200_EBITSTREAM;
                    }
                    int r = mb_intra_pred(br, m, mbx, mby, m.mb_type == P264B200_MB_I4x4);
                    if (r < 0) return r;
                } else {
                    if (n_list0_ < 1) return P264B200_EBITSTREAM;
                    int r;
                    if (t <= 2) {
                        m.mb_type = P264B200_MB_P_L0;
                        m.part = (uint8_t)t;  // 0 16x16, 1 16x8, 2 8x16 == P264B200_D_*
                        r = mb_inter_pred(br, m, mbx, mby, sh);
                    } else {
                        m.mb_type = P264B200_MB_P_8x8;
                        m.part = P264B200_D_8x8;
                        ref0 = (t == 4);
                        r = mb_sub_pred(br, m, mbx, mby, sh, ref0);
                    }
                    if (r < 0) return r;
                }
                if (m.mb_type != P264B200_MB_I16x16) {
                    const int c = br.ue();
                    if (c < 0 || c >= 48) {
                        fprintf(stderr, "invalid cbp\n");
                        return P264B200_EBITSTREAM;
                    }
                    const int cbp = m.mb_type == P264B200_MB_I4x4 ? kCbpIntra[c] : kCbpInter[c];
                    cbp_luma = cbp & 15;
                    m.cbp_chroma = (uint8_t)(cbp >> 4);
                }
                if (cbp_luma > 0 || m.cbp_chroma > 0 || m.mb_type == P264B200_MB_I16x16) {
                    // decoder/macroblock.c:568: delta applied to the slice QP, never accumulated
                    const int qp = br.se() + base_qp;
                    if (qp < 0 || qp > 51) return P264B200_EBITSTREAM;
                    m.qp = (uint8_t)qp;
                    int r = mb_residual(br, m, mbx, mby, cbp_luma);
                    if (r < 0) return r;
                } else {
                    if (base_qp < 0 || base_qp > 51) return P264B200_EBITSTREAM;
                    m.qp = (uint8_t)base_qp;
                }
                mb_finish(m, mbx, mby, cbp_luma);
            }
        }
        if (skip_run > 0) {
            mb_skip(m, mbx, mby);
            m.qp = (uint8_t)last_qp_;
            mb_finish(m, mbx, mby, 0);
            skip_run--;
        } else if (have_mb) {
            skip_run = -1;
        }
    }
    return 0;
}

int Parser::slice(int nal_type, int nal_ref_idc, BitReader &br, p264b200_frame_syntax *out, int *got_frame)
{
    SliceHeader sh;
    if (nal_type == 5) {
        // p264_slice_idr (decoder/decoder.c:43-64)
        for (auto &r : ring_) r.ref_type = 0;
    }
    int r = slice_header(br, nal_type, nal_ref_idc, sh);
    if (r < 0) {
        fprintf(stderr, "p264: p264_slice_header_decode failed\n");
        return r;
    }
    if (sh.first_mb != 0) return P264B200_EUNSUP;  // one slice per picture (decoder/decoder.c:516-523)
    ring_[0].frame_num = sh.frame_num;
    lists_init(sh);
    if (sh.redundant_pic_cnt != 0) return 0;
    r = slice_data(br, sh);
    if (r < 0) {
        fprintf(stderr, "p264: p264_slice_data_decode failed\n");
        return r;
    }
    memset(&hdr_, 0, sizeof(hdr_));
    hdr_.mb_w = mb_w_;
    hdr_.mb_h = mb_h_;
    hdr_.slice_type = sh.type;
    hdr_.deblock = !(apps_->deblocking_filter_control && sh.disable_deblocking_filter_idc == 1);
    hdr_.alpha_c0_offset = sh.alpha_c0_offset;
    hdr_.beta_offset = sh.beta_offset;
    hdr_.chroma_qp_index_offset = apps_->chroma_qp_index_offset;
    hdr_.num_ref = n_list0_;
    for (int i = 0; i < n_list0_; i++) hdr_.ref_slot[i] = ring_[list0_[i]].slot;
    hdr_.dst_slot = ring_[0].slot;
    hdr_.n_intra = n_intra_;
    hdr_.n_coef = (uint32_t)coef_n_;
    out->hdr = hdr_;
    out->mbs = mbs_;
    out->coefs = coefs_;
    *got_frame = 1;
    marking(nal_type, sh);
    return 0;
}

int Parser::nal(int nal_type, int nal_ref_idc, const uint8_t *payload, int size, p264b200_frame_syntax *out,
                int *got_frame)
{
    *got_frame = 0;
    BitReader br(payload, (size_t)(size < 0 ? 0 : size));
    switch (nal_type) {
    case 7: {
        int r = read_sps(br);
        if (r < 0) fprintf(stderr, "p264: p264_sps_read failed\n");
        return r < 0 ? r : 0;
    }
    case 8: {
        int r = read_pps(br);
        if (r < 0) fprintf(stderr, "p264: p264_pps_read failed\n");
        return r < 0 ? r : 0;
    }
    case 5:
    case 1: {
        int r = slice(nal_type, nal_ref_idc, br, out, got_frame);
        if (r < 0) fprintf(stderr, "p264: p264_slice_decode failed\n");
        return r;
    }
    case 2:
    case 3:
    case 4: fprintf(stderr, "partitioned stream unsupported\n"); return P264B200_EUNSUP;
    default: return 0;
    }
}

}  // namespace p264b200
