// Host syntax front-end: NAL payload -> FrameSyntax (include/p264b200_recon.h).
// Re-implements, from the H.264 syntax tables, what the reference does in
// decoder/set.c (SPS/PPS), decoder/decoder.c:70-301,368-593 (slice header, MB loop),
// decoder/macroblock.c:72-597 (MB layer), decoder/lists.c (DPB ring) and
// core/macroblock.c:40-252,870-1340 (predictors + the per-frame side arrays), but
// emits packed per-frame buffers instead of reconstructing macroblock by macroblock.
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/p264b200_recon.h"
#include "bitreader.h"

namespace p264b200 {

struct Sps {
    int id = -1;
    int profile_idc = 0, level_idc = 0;
    int log2_max_frame_num = 4;
    int poc_type = 0, log2_max_poc_lsb = 4;
    int delta_pic_order_always_zero = 0;
    int num_ref_frames = 1;
    int mb_w = 0, mb_h = 0;
    int frame_mbs_only = 1;
    int crop[4] = {0, 0, 0, 0};
};

struct Pps {
    int id = -1;
    int sps_id = 0;
    int cabac = 0, pic_order = 0;
    int num_slice_groups = 1, slice_group_map_type = 0;
    int num_ref_idx_l0 = 1, num_ref_idx_l1 = 1;
    int weighted_pred = 0, weighted_bipred = 0;
    int pic_init_qp = 26, pic_init_qs = 26;
    int chroma_qp_index_offset = 0;
    int deblocking_filter_control = 0, constrained_intra_pred = 0, redundant_pic_cnt = 0;
};

struct SliceHeader {
    int first_mb = 0, type = 0, pps_id = 0, frame_num = 0, idr_pic_id = 0;
    int field_pic = 0;
    int redundant_pic_cnt = 0;
    int num_ref_idx_l0_active = 1;
    int qp_delta = 0;
    int disable_deblocking_filter_idc = 0, alpha_c0_offset = 0, beta_offset = 0;
    int no_output_of_prior_pics = 0, long_term_reference_flag = 0, adaptive_ref_pic_marking = 0;
};

typedef void *(*alloc_fn)(size_t);
typedef void (*free_fn)(void *);

class Parser {
public:
    Parser(alloc_fn a = nullptr, free_fn f = nullptr);
    ~Parser();

    // One NAL unit (payload already unescaped, header byte removed: p264_nal_t semantics).
    // Returns <0 on error; *got_frame = 1 when `out` describes a complete picture.
    int nal(int nal_type, int nal_ref_idc, const uint8_t *payload, int size, p264b200_frame_syntax *out,
            int *got_frame);

    int mb_w() const { return mb_w_; }
    int mb_h() const { return mb_h_; }
    int ring_size() const { return ring_n_; }
    bool geometry_changed_reset() {
        bool g = geometry_changed_;
        geometry_changed_ = false;
        return g;
    }
    int verbose = 1;  // print the reference's SPS/PPS lines to stderr

private:
    int read_sps(BitReader &br);
    int read_pps(BitReader &br);
    int slice(int nal_type, int nal_ref_idc, BitReader &br, p264b200_frame_syntax *out, int *got_frame);
    int slice_header(BitReader &br, int nal_type, int nal_ref_idc, SliceHeader &sh);
    int slice_data(BitReader &br, const SliceHeader &sh);
    int context_init();
    void lists_init(const SliceHeader &sh);
    void marking(int nal_type, const SliceHeader &sh);

    // macroblock layer
    int mb_intra_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, bool i4x4);
    int mb_inter_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, const SliceHeader &sh);
    int mb_sub_pred(BitReader &br, p264b200_mb &m, int mbx, int mby, const SliceHeader &sh, bool ref0);
    int mb_residual(BitReader &br, p264b200_mb &m, int mbx, int mby, int cbp_luma);
    void mb_skip(p264b200_mb &m, int mbx, int mby);
    void mb_finish(p264b200_mb &m, int mbx, int mby, int cbp_luma);

    // predictors
    void predict_mv(int x4, int y4, int w4, int ref, int shape, int part_idx, int mvp[2]) const;
    int predict_nnz(const uint8_t *grid, int stride, int x, int y) const;
    void fill_motion(p264b200_mb &m, int mbx, int mby, int bx, int by, int w, int h, int ref, int mvx, int mvy);

    int ensure_coef(size_t need);

    alloc_fn alloc_;
    free_fn free_;

    Sps sps_[32];
    Pps pps_[256];
    const Sps *asps_ = nullptr;
    const Pps *apps_ = nullptr;
    bool geometry_changed_ = false;

    int mb_w_ = 0, mb_h_ = 0;
    // frame ring (decoder/lists.c): ring_[0] = slot being decoded, ring_[1..] = references newest first
    struct RingEntry {
        int slot;
        int ref_type;  // 0 unused, 1 short, 2 long
        int frame_num;
        int pic_num;
    };
    std::vector<RingEntry> ring_;
    int ring_n_ = 0, ring_used_ = 0;
    int list0_[16];
    int n_list0_ = 0;

    int last_qp_ = 0;  // core/macroblock.c:1247-1252: never reset per slice/frame
    int slice_qp_ = 26;

    // per-picture output buffers
    p264b200_frame_hdr hdr_;
    p264b200_mb *mbs_ = nullptr;
    size_t mbs_cap_ = 0;
    int16_t *coefs_ = nullptr;
    size_t coef_cap_ = 0, coef_n_ = 0;
    int n_intra_ = 0;

    // per-picture neighbour grids
    std::vector<uint8_t> nnz_y_, nnz_c_[2];
    std::vector<int8_t> imode_;   // 4x4 grid of intra4x4 modes (2 = DC for non-I4x4 MBs)
    std::vector<int8_t> ref4_;    // 4x4 grid: -2 unavailable / not decoded yet, -1 intra, >=0 list-0 index
    std::vector<int16_t> mv4_;    // 4x4 grid, 2 per cell
};

// Removes emulation prevention bytes exactly like p264_nal_decode (core/core.c:306-331),
// including its `src < end - 3` boundary condition.  Returns payload size.
int nal_unescape(const uint8_t *src, int size, uint8_t *dst, int *nal_type, int *nal_ref_idc);

}  // namespace p264b200
