/*
 * p264b200_recon.h -- frame-level C-ABI of the B200 macroblock reconstruction engine.
 *
 * This is the real hot-path boundary: the host (entropy decode, MV prediction)
 * fills one FrameSyntax per picture and hands it to the GPU, which runs everything
 * the reference does below decoder/macroblock.c:599 for that picture:
 *
 *   reference call                                     replaced by
 *   ------------------------------------------------   -------------------------------
 *   p264_macroblock_decode      decoder/macroblock.c:755   recon_inter / recon_intra kernels
 *   p264_macroblock_decode_skip decoder/macroblock.c:895   recon_inter kernel (P_SKIP = 16x16, no residual)
 *   p264_mb_mc                  core/macroblock.c:633      recon_inter kernel (MC from mv[16]/ref[4])
 *   p264_frame_deblocking_filter core/frame.c:490          deblock kernel
 *   p264_frame_expand_border    core/frame.c:205           border kernel
 *   p264_frame_filter           core/mc.c:409              eliminated (6-tap on the fly)
 *
 * Plain C, no torch / CUDA types in any signature.  Every entry point returns
 * 0 on success or a negative P264B200_E* code and never aborts the host.
 */
#ifndef P264B200_RECON_H
#define P264B200_RECON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define P264B200_ABI_VERSION 2   /* 2: FrameSyntax v2 (compact wire format) added; every v1 entry point is unchanged */

/* error codes */
#define P264B200_OK          0
#define P264B200_EINVAL     (-1)  /* bad argument */
#define P264B200_ENODEV     (-2)  /* no usable CUDA device: there is NO CPU fallback */
#define P264B200_ECUDA      (-3)  /* CUDA runtime error (see p264b200_last_error) */
#define P264B200_ENOMEM     (-4)
#define P264B200_EUNSUP     (-5)  /* syntax the reference cannot decode either */
#define P264B200_EBITSTREAM (-6)

/* macroblock classes (the subset core/macroblock.h:42-65 that the decoder reaches) */
enum {
    P264B200_MB_I4x4   = 0,
    P264B200_MB_I16x16 = 1,
    P264B200_MB_P_L0   = 2,   /* 16x16 / 16x8 / 8x16 */
    P264B200_MB_P_8x8  = 3,
    P264B200_MB_P_SKIP = 4
};
#define P264B200_IS_INTRA(t) ((t) <= P264B200_MB_I16x16)

enum { P264B200_SLICE_P = 0, P264B200_SLICE_I = 2 };

/* partition shapes, informational (MC is driven by mv[]/ref[] alone) */
enum { P264B200_D_16x16 = 0, P264B200_D_16x8 = 1, P264B200_D_8x16 = 2, P264B200_D_8x8 = 3 };
enum { P264B200_SUB_8x8 = 0, P264B200_SUB_8x4 = 1, P264B200_SUB_4x8 = 2, P264B200_SUB_4x4 = 3 };

/*
 * One macroblock of side information: exactly the fields p264_macroblock_decode and
 * p264_frame_deblocking_filter read from h->mb / the per-frame arrays written by
 * p264_macroblock_cache_save (core/macroblock.c:1234-1340).  96 bytes, all 4x4-block
 * indexed fields are in RASTER order inside the MB (b = 4*y + x), not the bitstream
 * z-order.
 */
typedef struct p264b200_mb {
    int16_t  mv[16][2];    /* qpel list-0 MV of every 4x4 block; 0 for intra                    */
    int8_t   ref[4];       /* list-0 index of every 8x8 (raster); -1 for intra                   */
    uint8_t  mb_type;      /* P264B200_MB_*                                                      */
    uint8_t  qp;           /* luma QP used for dequant (decoder/macroblock.c:568,578)            */
    uint8_t  qp_dbf;       /* QP the deblocker sees: last-QP rule of core/macroblock.c:1247-1252 */
    uint8_t  cbp_chroma;   /* 0 none, 1 DC only, 2 DC+AC                                         */
    uint16_t luma_mask;    /* bit b: luma 4x4 block b has total_coeff>0 and 16 coefs in the stream */
    uint8_t  i16_mode;     /* raw intra16x16 mode 0..3 (V,H,DC,P)                                */
    uint8_t  chroma_mode;  /* raw intra chroma mode 0..3 (DC,H,V,P)                              */
    uint8_t  i4_mode[8];   /* 16 nibbles, raster order, low nibble first: raw 4x4 mode 0..8      */
    uint32_t coef_off;     /* offset of this MB's chunk in the coefficient stream, int16 units   */
    uint8_t  chroma_mask;  /* bit i: chroma AC block i (0-3 Cb, 4-7 Cr, raster) present          */
    uint8_t  part;         /* P264B200_D_*                                                       */
    uint8_t  sub_part[4];  /* P264B200_SUB_* per 8x8                                             */
    uint8_t  reserved[2];
} p264b200_mb;

/*
 * Coefficient stream: int16, zig-zag order exactly as the CAVLC reader produced
 * them (decoder/dec_cavlc.c:1514-1520) truncated to 16 bit the way the reference's
 * unscan does (decoder/macroblock.c:605-630).  Chunk of one MB, in this order:
 *   [I16x16 only]         16  luma DC levels
 *   for each set bit b of luma_mask, ascending:  16 levels
 *                          (I16x16: slot 0 is 0 and slots 1..15 are the 15 AC levels)
 *   [cbp_chroma != 0]     4 Cb DC levels, 4 Cr DC levels (raster 2x2)
 *   for each set bit i of chroma_mask, ascending: 16 levels (slot 0 = 0, 1..15 AC)
 * Every chunk starts on an 8-element (16-byte) boundary.
 */

typedef struct p264b200_frame_hdr {
    int32_t  mb_w, mb_h;
    int32_t  slice_type;              /* P264B200_SLICE_*                                     */
    int32_t  deblock;                 /* 0 = filter off (disable_deblocking_filter_idc == 1)   */
    int32_t  alpha_c0_offset;         /* as the reference uses them: the raw, UN-doubled        */
    int32_t  beta_offset;             /*   slice header values (decoder/decoder.c:177-178)      */
    int32_t  chroma_qp_index_offset;
    int32_t  num_ref;                 /* size of list 0                                         */
    int32_t  ref_slot[16];            /* list-0 index -> frame-ring slot                        */
    int32_t  dst_slot;                /* ring slot this picture is reconstructed into           */
    int32_t  n_intra;                 /* number of intra MBs (0 lets the engine skip the wavefront) */
    uint32_t n_coef;                  /* int16 elements in the coefficient stream               */
    int32_t  reserved[4];
} p264b200_frame_hdr;

typedef struct p264b200_frame_syntax {
    p264b200_frame_hdr hdr;
    const p264b200_mb *mbs;           /* mb_w*mb_h records, raster                             */
    const int16_t     *coefs;         /* n_coef levels                                         */
} p264b200_frame_syntax;

/*
 * FrameSyntax v2 -- the compact wire format (ABI version 2).  Same picture, fewer bytes over PCIe: the 32-byte tail of every
 * macroblock record (all fields but mv[16]), one vector per PARTITION, and per coded block a 16-bit significance mask +
 * its non-zero levels (int8 when every level of the picture fits).  A dense synthetic 1080p P picture shrinks from 2.09 MB
 * to about 0.65 MB.  The engine expands it into the v1 staging layout on the device, so reconstruction is identical
 * (csrc/host/wire_v2.cc documents the sections; p264b200_pack_v2 produces them from a v1 FrameSyntax).
 */
#define P264B200_V2_LEVELS8 1u        /* flags: the level stream is int8 */
typedef struct p264b200_frame_syntax_v2 {
    p264b200_frame_hdr hdr;           /* as v1 (n_coef = size of the EXPANDED coefficient stream)           */
    const uint8_t *blob;              /* the packed picture, 16-byte aligned                                 */
    uint32_t blob_bytes;
    uint32_t flags;
    uint32_t off_hdr, off_offs, off_mv, off_mask, off_level;   /* section offsets inside blob (16-byte aligned) */
    uint32_t reserved[3];
} p264b200_frame_syntax_v2;

/* ------------------------------------------------------------------------- *
 * Engine: L independent lanes (streams / closed GOPs) reconstructed per launch
 * ------------------------------------------------------------------------- */
typedef struct p264b200_engine p264b200_engine;

typedef struct p264b200_engine_cfg {
    int32_t  device;                  /* CUDA ordinal                                          */
    int32_t  lanes;                   /* independent streams batched per launch                */
    int32_t  mb_w, mb_h;
    int32_t  n_slots;                 /* frame ring size per lane = num_ref_frames + 1          */
    uint32_t coef_capacity;           /* int16 per lane per frame (0 = dense worst case)        */
    int32_t  stage_steps;             /* frames per lane that can be pre-staged in HBM (>=1)    */
    uint32_t flags;
} p264b200_engine_cfg;

const char *p264b200_last_error(void);
int  p264b200_abi_version(void);
int  p264b200_device_count(void);

int  p264b200_engine_create(p264b200_engine **out, const p264b200_engine_cfg *cfg);
void p264b200_engine_destroy(p264b200_engine *e);

/* geometry of the padded device frame store (for upload/download helpers) */
int  p264b200_engine_geometry(const p264b200_engine *e, int32_t *luma_stride, int32_t *chroma_stride,
                              int32_t *width, int32_t *height);

/* Copy one picture's syntax from HOST memory into staging step `step` of lane `lane`.
 * Asynchronous when the source is pinned.  Uploads, kernels and downloads run on three internal
 * streams ordered by events, so H2D of step n+1, reconstruction of step n and D2H of step n-1
 * overlap; a staging step may be re-staged as soon as the call returns (the copy waits on the device
 * for the reconstruction that last used it), but the HOST buffers must stay untouched until
 * p264b200_engine_sync. */
int  p264b200_stage_frame(p264b200_engine *e, int step, int lane, const p264b200_frame_syntax *fs);

/* Batched form: lanes [0, n) of one step from an array of n FrameSyntax (one call instead of n). */
int  p264b200_stage_frames(p264b200_engine *e, int step, int n, const p264b200_frame_syntax *fs);

/* v2 (compact) form of p264b200_stage_frames: copies the packed pictures (one copy for the whole step when the blobs lie
 * back to back in host memory, each starting on the next 16-byte boundary) and expands them on the device into the
 * same staging area p264b200_stage_frames fills.  Asynchronous; same rules for the host buffers. */
int  p264b200_stage_frames_v2(p264b200_engine *e, int step, int n, const p264b200_frame_syntax_v2 *fs);
/* host side of v2 (no GPU needed): pack a v1 FrameSyntax into dst (>= p264b200_pack_v2_bound bytes, 16-byte aligned),
 * and the reference expander the device kernel is tested against */
size_t p264b200_pack_v2_bound(int mb_w, int mb_h, uint32_t n_coef);
int  p264b200_pack_v2(const p264b200_frame_syntax *fs, uint8_t *dst, size_t cap, p264b200_frame_syntax_v2 *out);
int  p264b200_unpack_v2(const p264b200_frame_syntax_v2 *in, p264b200_mb *mbs, int16_t *coefs);

/* Reconstruct staging step `step` for lanes [0, n_lanes): MC + IDCT + intra + deblock + border.
 * Asynchronous; inputs are already resident in HBM. */
int  p264b200_recon_step(p264b200_engine *e, int step, int n_lanes);

/* Convenience for the serial decoder: stage + reconstruct one picture on one lane. */
int  p264b200_recon_frame(p264b200_engine *e, int lane, const p264b200_frame_syntax *fs);

/* Picture transfer, tight (unpadded) I420 planes on the host side. Asynchronous on
 * the engine stream; call p264b200_engine_sync before reading host memory. */
int  p264b200_frame_upload(p264b200_engine *e, int lane, int slot,
                           const uint8_t *y, int y_stride, const uint8_t *u, const uint8_t *v, int c_stride);
int  p264b200_frame_download(p264b200_engine *e, int lane, int slot,
                             uint8_t *y, int y_stride, uint8_t *u, uint8_t *v, int c_stride);

/* Batched download: picture of lane l (ring slot slots[l]) into dst + l * picture_bytes as a tight
 * I420 image (Y, then U, then V); picture_bytes >= width*height*3/2. */
int  p264b200_frames_download(p264b200_engine *e, int n, const int32_t *slots, uint8_t *dst, size_t picture_bytes);

/* Output without a device -> host copy of the samples:
 *  - p264b200_frame_device_planes: the DEVICE addresses of a reconstructed picture's planes (sample 0,0 of Y, U, V; the
 *    padded frame store, strides from p264b200_engine_geometry) for consumers that stay on the GPU (display, transcode,
 *    analysis).  Valid until the ring slot is reconstructed into again; order your work after p264b200_engine_stream.
 *  - p264b200_frames_md5: MD5 (RFC 1321) of the tight I420 image (Y, U, V back to back) of lane l's slots[l], computed on the
 *    device, 16 bytes per lane into `digests` (host memory) -- the reference's own check ("decode, md5 the YUV") at 16 bytes
 *    of PCIe traffic per picture.  Asynchronous like the downloads; a verification path (one thread per picture). */
int  p264b200_frame_device_planes(p264b200_engine *e, int lane, int slot, void *planes[3]);
int  p264b200_frames_md5(p264b200_engine *e, int n, const int32_t *slots, uint8_t *digests);

int  p264b200_engine_sync(p264b200_engine *e);

/* CUDA stream the engine launches on (cudaStream_t as void*), for external event timing */
void *p264b200_engine_stream(p264b200_engine *e);

/* Device-side event timing on the engine stream (milliseconds). */
int  p264b200_timer_start(p264b200_engine *e);
int  p264b200_timer_stop(p264b200_engine *e, float *ms);
/* Per-kernel accumulated device time since the last reset (CUDA events around each launch;
 * only collected when profiling is enabled, which serialises nothing but adds events). */
int  p264b200_profile_enable(p264b200_engine *e, int on);
int  p264b200_profile_read(p264b200_engine *e, float ms_out[8], uint64_t launches_out[8]);
/* number of kernel launches issued by the engine since creation */
uint64_t p264b200_engine_launches(const p264b200_engine *e);

/* Diagnostics (library built with `make TRACE=1` only; otherwise the buffers stay zero): with P264B200_TRACE=<ticket> in
 * the environment the deblock CTA that drew that ticket records clock64() marks per row warp and macroblock
 * ([9][320 macroblocks][6 marks] int64); this copies them out (tools/dbf_trace.py). */
int  p264b200_debug_trace(void *dst, size_t bytes);
/* ... and every deblock CTA records %globaltimer (ns) at entry / first macroblock / last macroblock / exit: [2048 tickets][4] int64 */
int  p264b200_debug_cta_times(void *dst, size_t bytes);

/* pinned host memory helpers for the callers that pack FrameSyntax */
void *p264b200_host_alloc(size_t bytes);
void  p264b200_host_free(void *p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* P264B200_RECON_H */
