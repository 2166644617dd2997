/*
 * p264_b200_tables.h -- the reference's x264-style function-pointer tables, bound to the GPU.
 *
 * The reference reaches every pixel primitive through tables filled by p264_*_init
 * (decoder/decoder.c:702-711).  This library exports the same constructors with the same
 * struct layouts; each slot on the decode path is a host shim that stages the block in device
 * memory, launches the matching sm_100a device routine on that single block (the same device
 * code the batched frame kernels inline) and copies the result back.  The per-block pointer
 * ABI is a TEST surface (known-answer tests against the reference's own tables); production
 * traffic uses the frame-level ABI of p264b200_recon.h.
 *
 *   table / constructor              reference declaration      slots bound here
 *   p264_dct_function_t   dct_init   core/dct.h:76-100          add4x4/8x8/16x16_idct, add8x8/16x16_idct8,
 *                                                               idct4x4dc, dct2x2dc, idct2x2dc
 *   p264_quant_function_t quant_init core/quant.h:27-36         dequant_4x4, dequant_8x8
 *   p264_mb_dequant_4x4_dc/_2x2_dc   core/quant.h:40-41         (plain functions)
 *   p264_mc_functions_t   mc_init    core/mc.h:34-52            mc_luma, get_ref, mc_chroma, avg[10], avg_weight[10]
 *   p264_predict_t[7/7/12] *_init    core/predict.h:27,109-111  all 26 slots
 *   p264_predict8x8_t[12]            core/predict.h:28,112      NULL (Intra-8x8 is unreachable: decoder/set.c:230-242)
 *   p264_deblock_function_t          core/frame.h:75-87         all 8 slots
 *   p264_pixel_function_t            core/pixel.h:64-71         ssd[7]; sad/satd/sa8d/mbcmp NULL (encoder-only)
 * Encoder-only slots (forward DCT, quant cores) are NULL: the reference's decoder never calls them.
 */
#ifndef P264_B200_TABLES_H
#define P264_B200_TABLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef _DCT_H
typedef struct {
    void (*sub4x4_dct)(int16_t dct[4][4], uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2);
    void (*add4x4_idct)(uint8_t *p_dst, int i_dst, int16_t dct[4][4]);
    void (*sub8x8_dct)(int16_t dct[4][4][4], uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2);
    void (*add8x8_idct)(uint8_t *p_dst, int i_dst, int16_t dct[4][4][4]);
    void (*sub16x16_dct)(int16_t dct[16][4][4], uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2);
    void (*add16x16_idct)(uint8_t *p_dst, int i_dst, int16_t dct[16][4][4]);
    void (*sub8x8_dct8)(int16_t dct[8][8], uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2);
    void (*add8x8_idct8)(uint8_t *p_dst, int i_dst, int16_t dct[8][8]);
    void (*sub16x16_dct8)(int16_t dct[4][8][8], uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2);
    void (*add16x16_idct8)(uint8_t *p_dst, int i_dst, int16_t dct[4][8][8]);
    void (*dct4x4dc)(int16_t d[4][4]);
    void (*idct4x4dc)(int16_t d[4][4]);
    void (*dct2x2dc)(int16_t d[2][2]);
    void (*idct2x2dc)(int16_t d[2][2]);
} p264_dct_function_t;
#endif

#ifndef _QUANT_H
typedef struct {
    void (*quant_8x8_core)(int16_t dct[8][8], int quant_mf[8][8], int i_qbits, int f);
    void (*quant_4x4_core)(int16_t dct[4][4], int quant_mf[4][4], int i_qbits, int f);
    void (*quant_4x4_dc_core)(int16_t dct[4][4], int i_quant_mf, int i_qbits, int f);
    void (*quant_2x2_dc_core)(int16_t dct[2][2], int i_quant_mf, int i_qbits, int f);
    void (*dequant_4x4)(int16_t dct[4][4], int dequant_mf[6][4][4], int i_qp);
    void (*dequant_8x8)(int16_t dct[8][8], int dequant_mf[6][8][8], int i_qp);
} p264_quant_function_t;
#endif

#ifndef _MC_H
typedef struct {
    void (*mc_luma)(uint8_t **src, int i_src_stride, uint8_t *dst, int i_dst_stride, int mvx, int mvy, int i_width, int i_height);
    uint8_t *(*get_ref)(uint8_t **src, int i_src_stride, uint8_t *dst, int *i_dst_stride, int mvx, int mvy, int i_width, int i_height);
    void (*mc_chroma)(uint8_t *src, int i_src_stride, uint8_t *dst, int i_dst_stride, int mvx, int mvy, int i_width, int i_height);
    void (*avg[10])(uint8_t *dst, int i_dst, uint8_t *src, int i_src);
    void (*avg_weight[10])(uint8_t *dst, int i_dst, uint8_t *src, int i_src, int i_weight);
} p264_mc_functions_t;
#endif

#ifndef _PREDICT_H
typedef void (*p264_predict_t)(uint8_t *src, int i_stride);
typedef void (*p264_predict8x8_t)(uint8_t *src, int i_stride, int i_neighbor);
#endif

#ifndef _FRAME_H
typedef void (*p264_deblock_inter_t)(uint8_t *pix, int stride, int alpha, int beta, int8_t *tc0);
typedef void (*p264_deblock_intra_t)(uint8_t *pix, int stride, int alpha, int beta);
typedef struct {
    p264_deblock_inter_t deblock_v_luma, deblock_h_luma, deblock_v_chroma, deblock_h_chroma;
    p264_deblock_intra_t deblock_v_luma_intra, deblock_h_luma_intra, deblock_v_chroma_intra, deblock_h_chroma_intra;
} p264_deblock_function_t;
#endif

#ifndef _PIXEL_H
typedef int (*p264_pixel_cmp_t)(uint8_t *, int, uint8_t *, int);
typedef struct {
    p264_pixel_cmp_t sad[7], ssd[7], satd[7], sa8d[4], mbcmp[7];
} p264_pixel_function_t;
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
struct p264_t;
void p264_dct_init(int cpu, p264_dct_function_t *dctf);
void p264_quant_init(struct p264_t *h, int cpu, p264_quant_function_t *pf);
void p264_mb_dequant_4x4_dc(int16_t dct[4][4], int dequant_mf[6][4][4], int i_qscale);
void p264_mb_dequant_2x2_dc(int16_t dct[2][2], int dequant_mf[6][4][4], int i_qscale);
void p264_mc_init(int cpu, p264_mc_functions_t *pf);
void p264_predict_16x16_init(int cpu, p264_predict_t pf[7]);
void p264_predict_8x8c_init(int cpu, p264_predict_t pf[7]);
void p264_predict_4x4_init(int cpu, p264_predict_t pf[12]);
void p264_predict_8x8_init(int cpu, p264_predict8x8_t pf[12]);
void p264_deblock_init(int cpu, p264_deblock_function_t *pf);
void p264_pixel_init(int cpu, p264_pixel_function_t *pixf);
/* 0 when the table shims can run (a CUDA device is present), else P264B200_ENODEV */
int p264b200_tables_ready(void);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif
