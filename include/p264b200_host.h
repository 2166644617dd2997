/*
 * p264b200_host.h -- host-side (CPU) half of the drop-in: the syntax front-end that turns
 * NAL units into FrameSyntax buffers, plus small bitstream utilities.  No GPU needed.
 *
 * Replaces, on the host, decoder/set.c, decoder/decoder.c:70-301,368-593,
 * decoder/macroblock.c:72-597, decoder/dec_cavlc.c, decoder/lists.c and the predictor /
 * cache parts of core/macroblock.c:40-252,870-1340 of the reference.
 */
#ifndef P264B200_HOST_H
#define P264B200_HOST_H

#include "p264b200_recon.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct p264b200_parser p264b200_parser;

/* pinned != 0: FrameSyntax buffers live in page-locked memory (needs a CUDA device) */
p264b200_parser *p264b200_parser_open(int pinned, int verbose);
void p264b200_parser_close(p264b200_parser *p);

/* One NAL unit in p264_nal_t form (header byte stripped, emulation prevention removed).
 * On a completed picture *got_frame = 1 and *out points at parser-owned buffers that stay
 * valid until the next call.  Returns 0 or a negative P264B200_E* code. */
int p264b200_parser_nal(p264b200_parser *p, int nal_type, int nal_ref_idc, const uint8_t *payload, int size,
                        p264b200_frame_syntax *out, int *got_frame);
int p264b200_parser_geometry(const p264b200_parser *p, int *mb_w, int *mb_h, int *ring_size);

/* Annex-B byte-stream splitter (replaces the start-code FSM of p264decoder.c:259-301, without
 * its 3 MB NAL limit).  *pos is the scan cursor; on return 1, [*nal_start, *nal_start+*nal_size)
 * is the next NAL unit (header byte included, trailing zero bytes of the next start code
 * excluded the same way the reference CLI excludes them).  Returns 0 at end of buffer. */
int p264b200_annexb_next(const uint8_t *buf, size_t size, size_t *pos, size_t *nal_start, size_t *nal_size);

/* p264_nal_decode semantics (core/core.c:306-331) on raw pointers */
int p264b200_nal_unescape(const uint8_t *src, int size, uint8_t *dst, int *nal_type, int *nal_ref_idc);

/* raw CAVLC code tables for self-checks: kind 0 coeff_token[nC class], 1 chroma-DC coeff_token,
 * 2 total_zeros[total_coeff-1], 3 chroma-DC total_zeros, 4 run_before[min(zeros_left,7)-1] */
int p264b200_cavlc_table_entry(int kind, int table, int sym, int *len, int *bits);

/* ------------------------------------------------------------------------- *
 * Multi-stream decoder: N independent Annex-B byte streams of the
 * same coded size, one engine lane each.  Per step every stream parses up to its next complete
 * picture on a pool of host threads (what decoder/decoder.c:598-622 does for one stream), then ONE
 * batched stage + reconstruction + download serves all lanes.  The reference has no counterpart:
 * its CLI (p264decoder.c:164-381) decodes one stream on one core.  Needs a CUDA device.
 * ------------------------------------------------------------------------- */
typedef struct p264b200_multi p264b200_multi;
typedef struct p264b200_multi_cfg {
    int32_t device;       /* CUDA ordinal */
    int32_t n_streams;    /* 1..256 */
    int32_t n_threads;    /* parser threads, 0 = one per host core (capped at n_streams) */
    int32_t reserved[5];
} p264b200_multi_cfg;

int  p264b200_multi_open(p264b200_multi **out, const p264b200_multi_cfg *cfg);
void p264b200_multi_close(p264b200_multi *m);
/* whole byte stream of stream s; the buffer must stay valid and unchanged until close */
int  p264b200_multi_set_stream(p264b200_multi *m, int s, const uint8_t *annexb, size_t bytes);
/* Decode the next picture of every stream that still has one.  Returns the number of pictures
 * produced (0 = every stream has ended) or a negative P264B200_E* code; produced[s] (optional,
 * n_streams bytes) tells which streams delivered a picture in this step. */
int  p264b200_multi_step(p264b200_multi *m, uint8_t *produced);
/* Tight I420 picture (Y, U, V back to back, coded size) of stream s from the last step: pinned host
 * memory owned by the decoder, valid until the next step.  NULL if the stream produced nothing. */
const uint8_t *p264b200_multi_picture(const p264b200_multi *m, int s, int *width, int *height);

/* ------------------------------------------------------------------------- *
 * Closed-GOP splitter + ordered merge: ONE Annex-B stream cut at its IDR pictures
 * (decoder/decoder.c:43-64: an IDR empties the DPB, so every closed GOP decodes on its own), GOP g decoded on
 * lane g mod L of one batched engine, pictures delivered in stream order.  Across GPUs: rank r of N opens the
 * GOPs g with g mod N == r (p264b200_gop_scan gives the byte ranges) -- no exchange between ranks.
 * ------------------------------------------------------------------------- */
/* byte offset and picture count of every closed GOP (each begins at the parameter sets directly in front of its
 * IDR slice); returns the number of GOPs in the stream, fills at most max_gops entries */
int  p264b200_gop_scan(const uint8_t *annexb, size_t bytes, size_t *gop_begin, int32_t *gop_pictures, int max_gops);
typedef struct p264b200_gopdec p264b200_gopdec;
/* returns the number of lanes in use (<= lanes, <= GOPs) or a negative P264B200_E* code; the byte stream is copied */
int  p264b200_gopdec_open(p264b200_gopdec **out, int device, int lanes, int threads, const uint8_t *annexb, size_t bytes);
void p264b200_gopdec_close(p264b200_gopdec *d);
int  p264b200_gopdec_gops(const p264b200_gopdec *d);
/* next picture IN STREAM ORDER as a tight I420 image owned by the decoder (valid until the next call):
 * 1 = delivered, 0 = end of stream, < 0 = error */
int  p264b200_gopdec_next(p264b200_gopdec *d, const uint8_t **picture, int *width, int *height);

/* ------------------------------------------------------------------------- *
 * Synthetic stream generator (BASELINE.json configs 3-5): emits FrameSyntax for a
 * stream of random P pictures (optionally after one intra picture).  Deterministic
 * for a given cfg.  Not part of the reference (which ships no generator or tests).
 * ------------------------------------------------------------------------- */
typedef struct p264b200_synth p264b200_synth;
typedef struct p264b200_synth_cfg {
    int32_t  mb_w, mb_h;
    int32_t  n_refs;            /* reference frames (ring = n_refs + 1 slots), ref_idx uniform per partition */
    uint64_t seed;
    int32_t  qp_min, qp_max, qp_step;  /* slice QP cycles qp_min, qp_min+step, ... per picture */
    int32_t  coded_pct;         /* % of 4x4 luma blocks with residual                                   */
    int32_t  max_level;         /* |level| bound                                                        */
    int32_t  mv_range;          /* integer MV part uniform in +-mv_range samples, all 16 quarter phases */
    int32_t  sub8x8;            /* allow 8x4 / 4x8 / 4x4 sub-partitions                                 */
    int32_t  intra_pct;         /* % intra MBs inside P pictures                                        */
    int32_t  skip_pct;          /* % P_SKIP                                                             */
    int32_t  deblock;
    int32_t  sweep_offsets;     /* random alpha/beta offsets in [-6,6] per picture                      */
    int32_t  chroma_qp_index_offset;
    int32_t  confine_mv;        /* keep referenced samples inside the reference's 32-sample border      */
    int32_t  first_intra;       /* picture 0 is an all-intra picture                                    */
    int32_t  intra_period;      /* > 0 (with first_intra): every intra_period-th picture is all-intra   */
    int32_t  reserved[3];
} p264b200_synth_cfg;

void p264b200_synth_default(p264b200_synth_cfg *cfg, int mb_w, int mb_h);
p264b200_synth *p264b200_synth_open(const p264b200_synth_cfg *cfg);
void p264b200_synth_close(p264b200_synth *s);
/* next picture; *out points at generator-owned buffers valid until the next call */
int  p264b200_synth_next(p264b200_synth *s, p264b200_frame_syntax *out);

/* ------------------------------------------------------------------------- *
 * Bitstream writer (test / benchmark infrastructure): turns FrameSyntax
 * pictures into a real Baseline / CAVLC Annex-B stream that the UNMODIFIED reference decoder can
 * decode -- the inverse of the parser.  Restricted to what the stock decoder handles: the first
 * picture intra, one reference frame, partitions >= 8x8, one QP per picture.  P_SKIP records are
 * written as P_L0 16x16 without residual (same reconstruction).  No GPU needed.
 * ------------------------------------------------------------------------- */
typedef struct p264b200_writer p264b200_writer;
p264b200_writer *p264b200_writer_open(int mb_w, int mb_h, int chroma_qp_index_offset);
void p264b200_writer_close(p264b200_writer *w);
/* appends one picture (SPS + PPS first when it is an intra picture); returns the bytes appended or < 0 */
int  p264b200_writer_put(p264b200_writer *w, const p264b200_frame_syntax *fs);
const uint8_t *p264b200_writer_data(const p264b200_writer *w, size_t *bytes);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
