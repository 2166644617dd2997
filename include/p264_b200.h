/*
 * p264_b200.h -- the reference's public decode API, served by the B200 engine.
 *
 * A program written against the reference's p264.h links against libp264b200.so unchanged:
 * the entry points below keep the reference's names, argument meaning, ownership rules and
 * error behaviour, and the structs keep its field order (binary layout).  When the
 * reference's own p264.h has already been included this file only adds the prototypes of
 * the internal table constructors and declares nothing twice.
 *
 *   entry point                 replaces (reference file:line)
 *   -------------------------   ----------------------------------------
 *   p264_param_default          core/core.c:41        (p264.h:266)
 *   p264_picture_alloc/clean    core/core.c:183,253   (p264.h:300-305)
 *   p264_nal_encode/decode      core/core.c:258,306   (p264.h:347-351)
 *   p264_decoder_open           decoder/decoder.c:675 (p264.h:379)
 *   p264_decoder_decode         decoder/decoder.c:745 (p264.h:381)
 *   p264_decoder_close          decoder/decoder.c:812 (p264.h:380)
 *   p264_*_init (tables)        decoder/decoder.c:702-711, see p264_b200_tables.h
 *
 * Differences a caller can observe:
 *   * p264_decoder_open returns NULL (and says why on stderr) when no CUDA device is usable --
 *     there is no CPU reconstruction path in this library;
 *   * pic->img.plane[] point into a pinned host mirror of the device frame with the
 *     reference's stride (coded width + 64) and, like the reference, stay valid until the
 *     next p264_decoder_decode call that outputs a picture;
 *   * the device is chosen with the environment variable P264B200_DEVICE (default 0).
 */
#ifndef P264_B200_H
#define P264_B200_H

#include <stdarg.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef _P264_H /* the reference header was not included: provide the same declarations */
#define _P264_H 1
#define P264_BUILD 42

typedef struct p264_t p264_t; /* opaque decoder handle */

/* cpu capability bits (ignored by this implementation; kept so callers compile) */
#define P264_CPU_MMX 0x000001
#define P264_CPU_MMXEXT 0x000002
#define P264_CPU_SSE 0x000004
#define P264_CPU_SSE2 0x000008
#define P264_CPU_3DNOW 0x000010
#define P264_CPU_3DNOWEXT 0x000020
#define P264_CPU_ALTIVEC 0x000040

/* colour spaces */
#define P264_CSP_MASK 0x00ff
#define P264_CSP_NONE 0x0000
#define P264_CSP_I420 0x0001
#define P264_CSP_I422 0x0002
#define P264_CSP_I444 0x0003
#define P264_CSP_YV12 0x0004
#define P264_CSP_YUYV 0x0005
#define P264_CSP_RGB 0x0006
#define P264_CSP_BGR 0x0007
#define P264_CSP_BGRA 0x0008
#define P264_CSP_VFLIP 0x1000

/* picture types */
#define P264_TYPE_AUTO 0x0000
#define P264_TYPE_IDR 0x0001
#define P264_TYPE_I 0x0002
#define P264_TYPE_P 0x0003
#define P264_TYPE_BREF 0x0004
#define P264_TYPE_B 0x0005

/* log levels */
#define P264_LOG_NONE (-1)
#define P264_LOG_ERROR 0
#define P264_LOG_WARNING 1
#define P264_LOG_INFO 2
#define P264_LOG_DEBUG 3

/* misc encoder-era constants that p264_param_default writes */
#define P264_ANALYSE_I4x4 0x0001
#define P264_ANALYSE_I8x8 0x0002
#define P264_ANALYSE_PSUB16x16 0x0010
#define P264_ANALYSE_PSUB8x8 0x0020
#define P264_ANALYSE_BSUB16x16 0x0100
#define P264_DIRECT_PRED_TEMPORAL 2
#define P264_ME_HEX 1
#define P264_CQM_FLAT 0

typedef struct {
    int i_start, i_end;
    int b_force_qp;
    int i_qp;
    float f_bitrate_factor;
} p264_zone_t;

/* Parameter block.  Layout (field order and types) is that of p264.h:119-244; the decoder
 * reads only cpu, i_csp, i_bframe, b_cabac and i_frame_reference from it. */
typedef struct {
    unsigned int cpu;
    int i_threads;
    int i_width, i_height, i_csp, i_level_idc, i_frame_total;
    struct {
        int i_sar_height, i_sar_width, i_overscan;
        int i_vidformat, b_fullrange, i_colorprim, i_transfer, i_colmatrix, i_chroma_loc;
    } vui;
    int i_fps_num, i_fps_den;
    int i_frame_reference, i_keyint_max, i_keyint_min, i_scenecut_threshold;
    int i_bframe, b_bframe_adaptive, i_bframe_bias, b_bframe_pyramid;
    int b_deblocking_filter, i_deblocking_filter_alphac0, i_deblocking_filter_beta;
    int b_cabac, i_cabac_init_idc;
    int i_cqm_preset;
    char *psz_cqm_file;
    uint8_t cqm_4iy[16], cqm_4ic[16], cqm_4py[16], cqm_4pc[16], cqm_8iy[64], cqm_8py[64];
    void (*pf_log)(void *, int i_level, const char *psz, va_list);
    void *p_log_private;
    int i_log_level, b_visualize;
    struct {
        unsigned int intra, inter;
        int b_transform_8x8, b_weighted_bipred, i_direct_mv_pred, i_chroma_qp_offset;
        int i_me_method, i_me_range, i_mv_range, i_subpel_refine, b_chroma_me, b_bframe_rdo;
        int b_mixed_references, i_trellis, b_fast_pskip, b_psnr;
    } analyse;
    struct {
        int i_qp_constant, i_qp_min, i_qp_max, i_qp_step;
        int b_cbr, i_bitrate, i_rf_constant;
        float f_rate_tolerance;
        int i_vbv_max_bitrate, i_vbv_buffer_size;
        float f_vbv_buffer_init, f_ip_factor, f_pb_factor;
        int b_stat_write;
        char *psz_stat_out;
        int b_stat_read;
        char *psz_stat_in;
        char *psz_rc_eq;
        float f_qcompress, f_qblur, f_complexity_blur;
        p264_zone_t *zones;
        int i_zones;
        char *psz_zones;
    } rc;
    int b_aud, b_repeat_headers;
} p264_param_t;

/* p264.h:271-296 */
typedef struct {
    int i_csp;
    int i_plane;
    int i_stride[4];
    uint8_t *plane[4];
} p264_image_t;

typedef struct {
    int i_type;
    int i_qpplus1;
    int64_t i_pts;
    int i_width, i_height; /* out: coded size = 16 * macroblocks, cropping ignored (decoder/decoder.c:313-314) */
    p264_image_t img;
} p264_picture_t;

/* p264.h:311-341 */
enum nal_unit_type_e {
    NAL_UNKNOWN = 0,
    NAL_SLICE = 1,
    NAL_SLICE_DPA = 2,
    NAL_SLICE_DPB = 3,
    NAL_SLICE_DPC = 4,
    NAL_SLICE_IDR = 5,
    NAL_SEI = 6,
    NAL_SPS = 7,
    NAL_PPS = 8,
    NAL_AUD = 9
};
enum nal_priority_e { NAL_PRIORITY_DISPOSABLE = 0, NAL_PRIORITY_LOW = 1, NAL_PRIORITY_HIGH = 2, NAL_PRIORITY_HIGHEST = 3 };

typedef struct {
    int i_ref_idc;
    int i_type;
    int i_payload;
    uint8_t *p_payload; /* caller-owned buffer, filled by p264_nal_decode */
} p264_nal_t;

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
void p264_param_default(p264_param_t *param);
void p264_picture_alloc(p264_picture_t *pic, int i_csp, int i_width, int i_height);
void p264_picture_clean(p264_picture_t *pic);
int p264_nal_encode(void *p_data, int *pi_data, int b_annexeb, p264_nal_t *nal);
int p264_nal_decode(p264_nal_t *nal, void *p_data, int i_data);

p264_t *p264_decoder_open(p264_param_t *param);
void p264_decoder_close(p264_t *h);
/* one NAL per call; returns 0 or <0; *pp_pic is NULL or the decoder-owned output picture */
int p264_decoder_decode(p264_t *h, p264_picture_t **pp_pic, p264_nal_t *nal);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#endif /* _P264_H */

#ifdef __cplusplus
}
#endif
#endif /* P264_B200_H */
