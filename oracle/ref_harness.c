/*
 * ref_harness.c -- TEST INFRASTRUCTURE.  Driver code (ours) that is compiled TOGETHER WITH the
 * untouched reference sources taken in place from /root/reference (see oracle/Makefile) into
 * oracle/_ref/libp264ref.so.  It contains no reference code; it only calls the reference's
 * non-static functions so that the real implementation can be used
 *   (i)  as the "MB-feed oracle": p264_slice_decode's sequence (decoder/decoder.c:598-664)
 *        re-created with p264_macroblock_cache_load / p264_macroblock_decode /
 *        p264_macroblock_cache_save / p264_frame_deblocking_filter / border + half-pel passes,
 *        fed from the same FrameSyntax buffers the GPU engine consumes (bypasses the stock
 *        parser and its multi-ref / sub-8x8 bugs);
 *   (ii) as the hot-path-only CPU baseline (bench.py cpu_baseline.kind = "reference").
 * The primitive function tables (p264_dct_init, p264_mc_init, ...) are exported by the same
 * .so straight from the reference objects and are called from the tests through ctypes.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "core/core.h"
#include "decoder/macroblock.h"

#include "p264b200_recon.h"

/* non-static reference functions without a public prototype */
void p264_decoder_context_init(p264_t *h);  /* decoder/decoder.c:304 */
void p264_decoder_context_clean(p264_t *h); /* decoder/decoder.c:346 */
void p264_macroblock_init(p264_t *h);       /* decoder/decoder.c:490 */

#define API __attribute__((visibility("default")))

typedef struct ref_feed {
    p264_t *h;
    int n_slots;
    p264_frame_t *slot[17];
} ref_feed;

static const uint8_t zx[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
static const uint8_t zy[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};

API ref_feed *ref_feed_open(int mb_w, int mb_h, int n_slots)
{
    p264_param_t param;
    ref_feed *f;
    p264_t *h;
    int i;
    if (n_slots < 1 || n_slots > 17) return NULL;
    p264_param_default(&param);
    param.cpu = 0;
    param.i_frame_reference = n_slots > 1 ? n_slots - 1 : 1;
    h = p264_decoder_open(&param);
    if (!h) return NULL;
    f = calloc(1, sizeof(*f));
    f->h = h;
    f->n_slots = n_slots;
    /* fabricate the active parameter sets that p264_decoder_context_init reads */
    h->sps_array[0].i_id = 0;
    h->sps_array[0].i_mb_width = mb_w;
    h->sps_array[0].i_mb_height = mb_h;
    h->sps_array[0].i_num_ref_frames = n_slots - 1;
    h->sps_array[0].i_log2_max_frame_num = 16;
    h->sps_array[0].b_frame_mbs_only = 1;
    h->pps_array[0].i_id = 0;
    h->pps_array[0].i_sps_id = 0;
    h->pps_array[0].i_pic_init_qp = 26;
    h->pps_array[0].b_deblocking_filter_control = 1;
    for (i = 0; i < 6; i++) h->pps_array[0].scaling_list[i] = p264_cqm_flat16;
    h->sh.sps = &h->sps_array[0];
    h->sh.pps = &h->pps_array[0];
    {
        FILE *saved = stderr; /* context_init prints the size; keep test output quiet */
        (void)saved;
    }
    p264_decoder_context_init(h);
    for (i = 0; i < n_slots; i++) f->slot[i] = h->frames.reference[i];
    return f;
}

API void ref_feed_close(ref_feed *f)
{
    if (!f) return;
    p264_decoder_close(f->h);
    free(f);
}

/* seed a frame slot with pixels and make it usable as a reference
 * (decoder/decoder.c:644-649: border, half-pel planes, border of those) */
API void ref_feed_write(ref_feed *f, int slot, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    p264_frame_t *fr = f->slot[slot];
    const uint8_t *src[3] = {y, u, v};
    const int W = 16 * f->h->sps->i_mb_width, H = 16 * f->h->sps->i_mb_height;
    int c, r;
    for (c = 0; c < 3; c++) {
        const int w = c ? W / 2 : W, hh = c ? H / 2 : H;
        for (r = 0; r < hh; r++) memcpy(fr->plane[c] + r * fr->i_stride[c], src[c] + r * w, w);
    }
    p264_frame_expand_border(fr);
    p264_frame_filter(0, fr);
    p264_frame_expand_border_filtered(fr);
}

API void ref_feed_read(ref_feed *f, int slot, uint8_t *y, uint8_t *u, uint8_t *v)
{
    p264_frame_t *fr = f->slot[slot];
    uint8_t *dst[3] = {y, u, v};
    const int W = 16 * f->h->sps->i_mb_width, H = 16 * f->h->sps->i_mb_height;
    int c, r;
    for (c = 0; c < 3; c++) {
        const int w = c ? W / 2 : W, hh = c ? H / 2 : H;
        for (r = 0; r < hh; r++) memcpy(dst[c] + r * w, fr->plane[c] + r * fr->i_stride[c], w);
    }
}

static void feed_mb(p264_t *h, const p264b200_mb *m, const int16_t *coefs)
{
    const int16_t *cf = coefs + m->coef_off;
    int i, k;
    int off[16], o = 0;

    h->mb.i_qp = m->qp;
    h->mb.i_cbp_chroma = m->cbp_chroma;
    h->mb.i_cbp_luma = m->luma_mask ? 15 : 0;
    h->mb.i_chroma_pred_mode = m->chroma_mode;
    h->mb.i_intra16x16_pred_mode = m->i16_mode;

    switch (m->mb_type) {
    case P264B200_MB_I4x4: h->mb.i_type = I_4x4; break;
    case P264B200_MB_I16x16: h->mb.i_type = I_16x16; break;
    case P264B200_MB_P_8x8: h->mb.i_type = P_8x8; break;
    default: h->mb.i_type = P_L0; break;
    }
    if (m->mb_type == P264B200_MB_P_8x8) {
        static const int sub[4] = {D_L0_8x8, D_L0_8x4, D_L0_4x8, D_L0_4x4};
        h->mb.i_partition = D_8x8;
        for (i = 0; i < 4; i++) h->mb.i_sub_partition[i] = sub[m->sub_part[i] & 3];
    } else {
        static const int part[4] = {D_16x16, D_16x8, D_8x16, D_8x8};
        h->mb.i_partition = part[m->part & 3];
    }

    if (m->mb_type == P264B200_MB_I16x16) {
        for (k = 0; k < 16; k++) h->dct.luma16x16_dc[k] = cf[k];
        cf += 16;
    }
    for (i = 0; i < 16; i++) { /* chunks are stored in raster-bit order */
        off[i] = o;
        if (m->luma_mask & (1 << i)) o += 16;
    }
    for (i = 0; i < 16; i++) { /* i = bitstream (z) order */
        const int b = zy[i] * 4 + zx[i];
        const int coded = (m->luma_mask >> b) & 1;
        h->mb.cache.non_zero_count[p264_scan8[i]] = coded;
        if (coded) {
            if (m->mb_type == P264B200_MB_I16x16)
                for (k = 0; k < 15; k++) h->dct.block[i].residual_ac[k] = cf[off[b] + 1 + k];
            else
                for (k = 0; k < 16; k++) h->dct.block[i].luma4x4[k] = cf[off[b] + k];
        }
        h->mb.cache.intra4x4_pred_mode[p264_scan8[i]] = (m->i4_mode[b >> 1] >> ((b & 1) * 4)) & 15;
        if (!P264B200_IS_INTRA(m->mb_type)) {
            h->mb.cache.ref[0][p264_scan8[i]] = m->ref[(b >> 3) * 2 + ((b & 3) >> 1)];
            h->mb.cache.mv[0][p264_scan8[i]][0] = m->mv[b][0];
            h->mb.cache.mv[0][p264_scan8[i]][1] = m->mv[b][1];
        }
    }
    cf += o;
    for (i = 0; i < 8; i++) h->mb.cache.non_zero_count[p264_scan8[16 + i]] = 0;
    if (m->cbp_chroma) {
        for (k = 0; k < 4; k++) {
            h->dct.chroma_dc[0][k] = cf[k];
            h->dct.chroma_dc[1][k] = cf[4 + k];
        }
        cf += 8;
        for (i = 0; i < 8; i++)
            if (m->chroma_mask & (1 << i)) {
                for (k = 0; k < 15; k++) h->dct.block[16 + i].residual_ac[k] = cf[1 + k];
                h->mb.cache.non_zero_count[p264_scan8[16 + i]] = 1;
                cf += 16;
            }
    }

    p264_macroblock_decode(h);
    if (m->mb_type == P264B200_MB_P_SKIP) h->mb.i_type = P_SKIP;

    /* make p264_macroblock_cache_save (core/macroblock.c:1247-1252) record exactly qp_dbf */
    h->mb.i_qp = m->qp_dbf;
    h->mb.i_last_qp = m->qp_dbf;
    p264_macroblock_cache_save(h);
}

/* one picture through the reference's own reconstruction, p264_slice_decode steps [3]-[4] */
API int ref_feed_frame(ref_feed *f, const p264b200_frame_hdr *hd, const p264b200_mb *mbs, const int16_t *coefs,
                       int do_deblock, int do_filter)
{
    p264_t *h = f->h;
    int i, mb_xy;
    const int n_mb = hd->mb_w * hd->mb_h;
    if (hd->mb_w != h->sps->i_mb_width || hd->mb_h != h->sps->i_mb_height) return -1;

    h->sh.i_type = hd->slice_type == P264B200_SLICE_I ? SLICE_TYPE_I : SLICE_TYPE_P;
    h->sh.i_first_mb = 0;
    h->sh.i_alpha_c0_offset = hd->alpha_c0_offset;
    h->sh.i_beta_offset = hd->beta_offset;
    h->sh.i_disable_deblocking_filter_idc = hd->deblock ? 0 : 1;
    h->pps->i_chroma_qp_index_offset = hd->chroma_qp_index_offset;
    h->fdec = f->slot[hd->dst_slot];
    h->fenc = h->fdec;
    h->i_ref0 = hd->slice_type == P264B200_SLICE_I ? 0 : hd->num_ref;
    h->i_ref1 = 0;
    for (i = 0; i < h->i_ref0; i++) h->fref0[i] = f->slot[hd->ref_slot[i]];

    for (mb_xy = 0; mb_xy < n_mb; mb_xy++) {
        p264_macroblock_init(h);
        p264_macroblock_cache_load(h, mb_xy % hd->mb_w, mb_xy / hd->mb_w);
        feed_mb(h, &mbs[mb_xy], coefs);
    }
    if (hd->deblock && do_deblock) p264_frame_deblocking_filter(h, h->sh.i_type);
    if (do_filter) {
        p264_frame_expand_border(h->fdec);
        p264_frame_filter(0, h->fdec);
        p264_frame_expand_border_filtered(h->fdec);
    }
    return 0;
}

/* whole-stream decode through the reference's public API (p264.h:379-382), tight I420 out.
 * NAL splitting is done by the caller; this mirrors Decode() of p264decoder.c:164-381. */
typedef struct ref_dec {
    p264_t *h;
    p264_nal_t nal;
    int cap;
} ref_dec;

API ref_dec *ref_dec_open(void)
{
    p264_param_t param;
    ref_dec *d = calloc(1, sizeof(*d));
    p264_param_default(&param);
    param.cpu = 0;
    d->h = p264_decoder_open(&param);
    d->cap = 1 << 20;
    d->nal.p_payload = malloc(d->cap);
    return d;
}
API void ref_dec_close(ref_dec *d)
{
    if (!d) return;
    p264_decoder_close(d->h);
    free(d->nal.p_payload);
    free(d);
}
/* returns 1 and fills y/u/v (tight) when the NAL completed a picture, 0 otherwise, <0 on error */
API int ref_dec_nal(ref_dec *d, const uint8_t *nal_bytes, int size, uint8_t *y, uint8_t *u, uint8_t *v, int *w, int *hgt)
{
    p264_picture_t *pic = NULL;
    int r, c, row;
    if (size + 16 > d->cap) {
        d->cap = size * 2 + 16;
        d->nal.p_payload = realloc(d->nal.p_payload, d->cap);
    }
    p264_nal_decode(&d->nal, (void *)nal_bytes, size);
    r = p264_decoder_decode(d->h, &pic, &d->nal);
    if (r < 0) return r;
    if (!pic) return 0;
    *w = pic->i_width;
    *hgt = pic->i_height;
    if (y) {
        uint8_t *dst[3] = {y, u, v};
        for (c = 0; c < 3; c++) {
            const int ww = c ? pic->i_width / 2 : pic->i_width, hh = c ? pic->i_height / 2 : pic->i_height;
            for (row = 0; row < hh; row++)
                memcpy(dst[c] + row * ww, pic->img.plane[c] + row * pic->img.i_stride[c], ww);
        }
    }
    return 1;
}

/* helpers for the primitive KATs that need reference-side state */
API int ref_dequant4_table(int list, int q, int y, int x)
{
    /* flat-16 table built by p264_cqm_init (core/set.c:69-107) inside a scratch context */
    static ref_feed *f;
    if (!f) f = ref_feed_open(1, 1, 2);
    return f->h->dequant4_mf[list][q][y][x];
}
API void ref_dequant_4x4(int16_t d[16], int qp)
{
    static ref_feed *f;
    if (!f) f = ref_feed_open(1, 1, 2);
    f->h->quantf.dequant_4x4((int16_t(*)[4])d, f->h->dequant4_mf[CQM_4IY], qp);
}
API void ref_dequant_8x8(int16_t d[64], int qp)
{
    static ref_feed *f;
    if (!f) f = ref_feed_open(1, 1, 2);
    f->h->quantf.dequant_8x8((int16_t(*)[8])d, f->h->dequant8_mf[CQM_8IY], qp);
}
API void ref_dequant_4x4_dc(int16_t d[16], int qp)
{
    static ref_feed *f;
    if (!f) f = ref_feed_open(1, 1, 2);
    p264_mb_dequant_4x4_dc((int16_t(*)[4])d, f->h->dequant4_mf[CQM_4IY], qp);
}
API void ref_dequant_2x2_dc(int16_t d[4], int qp)
{
    static ref_feed *f;
    if (!f) f = ref_feed_open(1, 1, 2);
    p264_mb_dequant_2x2_dc((int16_t(*)[2])d, f->h->dequant4_mf[CQM_4IC], qp);
}

/* mc_luma / mc_chroma of the reference on a frame slot prepared by ref_feed_write (integer plane +
 * the three half-pel planes of p264_frame_filter), block at picture position (bx,by) */
API void ref_mc_luma(ref_feed *f, int slot, int bx, int by, int mvx, int mvy, int w, int h, uint8_t *dst, int dstride)
{
    p264_frame_t *fr = f->slot[slot];
    uint8_t *src[4];
    int i;
    for (i = 0; i < 4; i++) src[i] = fr->filtered[i] + by * fr->i_stride[0] + bx;
    f->h->mc.mc_luma(src, fr->i_stride[0], dst, dstride, mvx, mvy, w, h);
}
API void ref_mc_chroma(ref_feed *f, int slot, int plane, int bx, int by, int mvx, int mvy, int w, int h, uint8_t *dst, int dstride)
{
    p264_frame_t *fr = f->slot[slot];
    f->h->mc.mc_chroma(fr->plane[plane] + by * fr->i_stride[plane] + bx, fr->i_stride[plane], dst, dstride, mvx, mvy, w, h);
}
