// In-loop deblocking filter as a macroblock-row wavefront.
//
// Replaces p264_frame_deblocking_filter (core/frame.c:490-643), deblock_edge (:472-488) and the
// eight edge filters (:302-470).  The reference filters macroblocks in raster order, vertical
// edges then horizontal edges per macroblock; MB(x,y) therefore depends on MB(x-1,y) (whole MB)
// and on MB(x+1,y-1) (its left-edge filter rewrites columns 13..15 of MB(x,y-1), which MB(x,y)'s
// top-edge filter reads).  A picture-wide "all vertical, then all horizontal" pass is NOT
// bit-exact, so the kernel keeps the reference order: row y may filter MB x once row y-1 has
// published progress >= min(x+2, mb_w).
//
// Work decomposition (v2).  Luma and chroma are independent given the boundary strengths, so
// they run as separate warps with separate progress flags.  A warp owns one macroblock row of
// TWO lanes (streams): threads 0..15 work on stream 2p, threads 16..31 on stream 2p+1, always in
// the same code path, so there is no luma/chroma divergence and every thread is busy:
//   luma warp  : thread = one of the 16 lines crossing the edge direction
//   chroma warp: thread = one of the 8 Cb + 8 Cr lines
// Vertical edges (filter along a row) are done entirely in registers: the thread keeps its row of
// the current macroblock (one 128-bit load) plus the 4 right-most samples of the previous
// macroblock, carried from iteration to iteration.  Horizontal edges need the transpose, which
// goes through a small shared-memory tile (row-wise 32-bit stores, column-wise byte loads, both
// bank-conflict free).  Boundary strengths are derived one per thread and exchanged with warp
// shuffles.  Only samples that changed are written back; the right-most 4 columns stay in
// registers until the next macroblock's left edge has been filtered.
#pragma once
#include "common.cuh"

namespace p264b200 {

// ---- scalar line filters (also used by the one-block table shims in blockops.cu) -------------
// bS < 4 luma filter on one line (core/frame.c:310-338); v = p3 p2 p1 p0 q0 q1 q2 q3
__device__ __forceinline__ void dbf_luma_normal(int v[8], int alpha, int beta, int tc0)
{
    const int p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        int tc = tc0;
        if (abs(p2 - p0) < beta) {
            v[2] = p1 + clip3i(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0);
            tc++;
        }
        if (abs(q2 - q0) < beta) {
            v[5] = q1 + clip3i(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0);
            tc++;
        }
        const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        v[3] = clip8i(p0 + delta);
        v[4] = clip8i(q0 - delta);
    }
}
// bS == 4 luma filter on one line (core/frame.c:390-431)
__device__ __forceinline__ void dbf_luma_strong(int v[8], int alpha, int beta)
{
    const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        if (abs(p0 - q0) < ((alpha >> 2) + 2)) {
            if (abs(p2 - p0) < beta) {
                v[3] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
                v[2] = (p2 + p1 + p0 + q0 + 2) >> 2;
                v[1] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
            } else
                v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
            if (abs(q2 - q0) < beta) {
                v[4] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
                v[5] = (p0 + q0 + q1 + q2 + 2) >> 2;
                v[6] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
            } else
                v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
        } else {
            v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
            v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
        }
    }
}
// chroma line, normal (core/frame.c:360-373) / strong (:446-459); v = p1 p0 q0 q1
__device__ __forceinline__ void dbf_chroma(int v[4], int alpha, int beta, int bs, int tc)
{
    const int p1 = v[0], p0 = v[1], q0 = v[2], q1 = v[3];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        if (bs < 4) {
            const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
            v[1] = clip8i(p0 + delta);
            v[2] = clip8i(q0 - delta);
        } else {
            v[1] = (2 * p1 + p0 + q1 + 2) >> 2;
            v[2] = (2 * q1 + q0 + p1 + 2) >> 2;
        }
    }
}

// ---- branch-free (predicated) forms used by the frame kernel ----------------------------------
// one luma edge on a line held in registers; bs == 0 leaves the line untouched
__device__ __forceinline__ void luma_edge(int &p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int &q3, int bs,
                                          int alpha, int beta, uint32_t tc0_packed, bool any_strong)
{
    const int tc0 = (int)((tc0_packed >> (8 * ((bs - 1) & 3))) & 0xff);
    const bool f = bs != 0 && abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
    const bool fn = f && bs < 4;
    const int avg = (p0 + q0 + 1) >> 1;
    const int np1 = p1 + clip3i(((p2 + avg) >> 1) - p1, -tc0, tc0);
    const int nq1 = q1 + clip3i(((q2 + avg) >> 1) - q1, -tc0, tc0);
    const int tc = tc0 + (int)ap + (int)aq;
    const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    int r1 = (fn && ap) ? np1 : p1, r0 = fn ? clip8i(p0 + delta) : p0;
    int s0 = fn ? clip8i(q0 - delta) : q0, s1 = (fn && aq) ? nq1 : q1;
    int r2 = p2, s2 = q2;
    if (any_strong) {  // warp-uniform: some line of this edge is an intra macroblock edge
        const bool fs = f && bs == 4;
        const bool sm = abs(p0 - q0) < ((alpha >> 2) + 2);
        const bool sp = fs && sm && ap, sq = fs && sm && aq;
        const int wp0 = (2 * p1 + p0 + q1 + 2) >> 2, wq0 = (2 * q1 + q0 + p1 + 2) >> 2;
        r0 = fs ? (sp ? (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3 : wp0) : r0;
        r1 = sp ? (p2 + p1 + p0 + q0 + 2) >> 2 : r1;
        r2 = sp ? (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3 : r2;
        s0 = fs ? (sq ? (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3 : wq0) : s0;
        s1 = sq ? (p0 + q0 + q1 + q2 + 2) >> 2 : s1;
        s2 = sq ? (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3 : s2;
    }
    p2 = r2, p1 = r1, p0 = r0, q0 = s0, q1 = s1, q2 = s2;
}
__device__ __forceinline__ void chroma_edge(int p1, int &p0, int &q0, int q1, int bs, int alpha, int beta, uint32_t tc0_packed)
{
    const int tc = (int)((tc0_packed >> (8 * ((bs - 1) & 3))) & 0xff) + 1;
    const bool f = bs != 0 && abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    const int n0 = bs < 4 ? clip8i(p0 + delta) : (2 * p1 + p0 + q1 + 2) >> 2;
    const int m0 = bs < 4 ? clip8i(q0 - delta) : (2 * q1 + q0 + p1 + 2) >> 2;
    p0 = f ? n0 : p0;
    q0 = f ? m0 : q0;
}

// packed-parameter variants (see DeblockSide)
__device__ __forceinline__ void luma_edge_p(int &p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int &q3, int bs,
                                            uint32_t prm, bool any_strong)
{
    const int alpha = prm & 0xff, beta = (prm >> 8) & 31;
    const int tc0 = (int)((prm >> (8 + 5 * bs)) & 31) & (bs < 4 ? 31 : 0);
    const bool f = bs != 0 && abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
    const bool fn = f && bs < 4;
    const int avg = (p0 + q0 + 1) >> 1;
    const int np1 = p1 + clip3i(((p2 + avg) >> 1) - p1, -tc0, tc0);
    const int nq1 = q1 + clip3i(((q2 + avg) >> 1) - q1, -tc0, tc0);
    const int tc = tc0 + (int)ap + (int)aq;
    const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    int r1 = (fn && ap) ? np1 : p1, r0 = fn ? clip8i(p0 + delta) : p0;
    int s0 = fn ? clip8i(q0 - delta) : q0, s1 = (fn && aq) ? nq1 : q1;
    int r2 = p2, s2 = q2;
    if (any_strong) {  // warp-uniform: some line of this edge is an intra macroblock edge
        const bool fs = f && bs == 4;
        const bool sm = abs(p0 - q0) < ((alpha >> 2) + 2);
        const bool sp = fs && sm && ap, sq = fs && sm && aq;
        const int wp0 = (2 * p1 + p0 + q1 + 2) >> 2, wq0 = (2 * q1 + q0 + p1 + 2) >> 2;
        r0 = fs ? (sp ? (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3 : wp0) : r0;
        r1 = sp ? (p2 + p1 + p0 + q0 + 2) >> 2 : r1;
        r2 = sp ? (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3 : r2;
        s0 = fs ? (sq ? (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3 : wq0) : s0;
        s1 = sq ? (p0 + q0 + q1 + q2 + 2) >> 2 : s1;
        s2 = sq ? (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3 : s2;
    }
    p2 = r2, p1 = r1, p0 = r0, q0 = s0, q1 = s1, q2 = s2;
}
__device__ __forceinline__ void chroma_edge_p(int p1, int &p0, int &q0, int q1, int bs, uint32_t prm)
{
    const int alpha = prm & 0xff, beta = (prm >> 8) & 31;
    const int tc = (int)((prm >> (8 + 5 * bs)) & 31) + 1;
    const bool f = bs != 0 && abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    const int n0 = bs < 4 ? clip8i(p0 + delta) : (2 * p1 + p0 + q1 + 2) >> 2;
    const int m0 = bs < 4 ? clip8i(q0 - delta) : (2 * q1 + q0 + p1 + 2) >> 2;
    p0 = f ? n0 : p0;
    q0 = f ? m0 : q0;
}

// alpha / beta / packed tc0[3] for an averaged QP (core/frame.c:476-483)
struct EdgeParams {
    int alpha, beta;
    uint32_t tc0;
};
__device__ __forceinline__ EdgeParams edge_params(int qp, int alpha_off, int beta_off)
{
    EdgeParams e;
    const int ia = clip3i(qp + alpha_off, 0, 51);
    e.alpha = c_alpha[ia];
    e.beta = c_beta[clip3i(qp + beta_off, 0, 51)];
    e.tc0 = *reinterpret_cast<const uint32_t *>(c_tc0[ia]);
    return e;
}

// boundary strength of one 4-sample segment (core/frame.c:535-581); m = current MB, n = neighbour
// across the edge (== m for inner edges)
__device__ __forceinline__ int boundary_strength(const p264b200_mb *__restrict__ m, const p264b200_mb *__restrict__ n, int dir,
                                                 int e, int seg)
{
    const int tm = __ldg(&m->mb_type), tn = __ldg(&n->mb_type);
    if (P264B200_IS_INTRA(tm) || P264B200_IS_INTRA(tn)) return e == 0 ? 4 : 3;
    const int x = dir == 0 ? e : seg, y = dir == 0 ? seg : e;
    const int xn = (x - (dir == 0)) & 3, yn = (y - (dir == 1)) & 3;
    const int bq = y * 4 + x, bp = yn * 4 + xn;
    const unsigned mq = __ldg(&m->luma_mask), mp = __ldg(&n->luma_mask);
    if (((mq >> bq) | (mp >> bp)) & 1) return 2;
    const int rq = __ldg(&m->ref[(bq >> 3) * 2 + ((bq & 3) >> 1)]), rp = __ldg(&n->ref[(bp >> 3) * 2 + ((bp & 3) >> 1)]);
    const int vq = __ldg(reinterpret_cast<const int *>(m->mv[bq])), vp = __ldg(reinterpret_cast<const int *>(n->mv[bp]));
    const int dx = abs((int)(short)(vq & 0xffff) - (int)(short)(vp & 0xffff)), dy = abs((vq >> 16) - (vp >> 16));
    return (rq != rp || dx >= 4 || dy >= 4) ? 1 : 0;
}

// Boundary strengths are pure syntax (no sample dependency), so they are derived by a fully
// parallel pre-pass and not inside the wavefront's dependent chain.  Per macroblock: four 32-bit
// words, word `seg` = bytes e=0..3, byte = bS(vertical edge e, segment seg) | bS(horizontal edge e,
// segment seg) << 4; plus one word qp | qp_left << 8 | qp_top << 16 (the QPs the deblocker sees).
// A luma thread (line i) and a chroma thread (line l) each need exactly one bS word:
// seg = i >> 2 resp. l >> 1, for both edge directions.
// The same pre-pass also resolves the filter parameters (core/frame.c:476-483) of the three edge kinds a
// macroblock has -- left MB edge, top MB edge, inner edges -- for luma and chroma:
//   bits 0-7 alpha, 8-12 beta, 13-17 / 18-22 / 23-27 tc0[bS-1] for bS = 1..3
struct DeblockSide {
    uint32_t bs[4];
    uint32_t luma[4];    // [left, top, inner, unused]
    uint32_t chroma[4];
};
__device__ __forceinline__ uint32_t pack_edge_params(int qp, int alpha_off, int beta_off)
{
    const int ia = clip3i(qp + alpha_off, 0, 51);
    return (uint32_t)c_alpha[ia] | ((uint32_t)c_beta[clip3i(qp + beta_off, 0, 51)] << 8) | ((uint32_t)c_tc0[ia][0] << 13) |
           ((uint32_t)c_tc0[ia][1] << 18) | ((uint32_t)c_tc0[ia][2] << 23);
}

constexpr int kLS = 20;                  // luma transpose tile: 20 rows (-4..15) x 16 cols, 20-byte rows
constexpr int kLHalf = 20 * kLS + 16;    // bytes per stream half (+16 shifts the second half's banks)
constexpr int kCS2 = 12;                 // chroma transpose tile: per plane 10 rows (-2..7) x 8 cols, 12-byte rows
constexpr int kCPlane = 10 * kCS2 + 8;
constexpr int kCHalf = 2 * kCPlane + 16;

#ifdef P264B200_DEFINE_KERNELS

// one thread per (macroblock, segment): 8 boundary strengths -> one word; thread seg 0 also writes the QP word
__global__ void __launch_bounds__(256) deblock_bs_kernel(const FrameDesc *__restrict__ descs, Geometry g)
{
    const FrameDesc &fd = descs[blockIdx.y];
    if (!fd.deblock) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int mb_xy = t >> 2, seg = t & 3;
    if (mb_xy >= g.mb_w * g.mb_h) return;
    const int mbx = mb_xy % g.mb_w, mby = mb_xy / g.mb_w;
    const p264b200_mb *m = fd.mbs + mb_xy;
    uint32_t w = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        int bv = 0, bh = 0;
        if (e > 0 || mbx > 0) bv = boundary_strength(m, e > 0 ? m : m - 1, 0, e, seg);
        if (e > 0 || mby > 0) bh = boundary_strength(m, e > 0 ? m : m - g.mb_w, 1, e, seg);
        w |= (uint32_t)(bv | (bh << 4)) << (8 * e);
    }
    fd.dbf_bs[mb_xy].bs[seg] = w;
    if (seg < 3) {
        // seg 0/1/2 -> left / top / inner edge parameters, luma QP average and mapped chroma QP average
        const int qp = m->qp_dbf, qn = seg == 0 ? (mbx > 0 ? m[-1].qp_dbf : qp) : seg == 1 ? (mby > 0 ? m[-g.mb_w].qp_dbf : qp) : qp;
        const int off = fd.chroma_qp_off;
        const int qc = c_chroma_qp[clip3i(qp + off, 0, 51)], qcn = c_chroma_qp[clip3i(qn + off, 0, 51)];
        fd.dbf_bs[mb_xy].luma[seg] = pack_edge_params((qp + qn + 1) >> 1, fd.alpha_off, fd.beta_off);
        fd.dbf_bs[mb_xy].chroma[seg] = pack_edge_params((qc + qcn + 1) >> 1, fd.alpha_off, fd.beta_off);
    }
}

// A CTA owns kDbfRows consecutive macroblock rows (one warp each) of one stream pair and one role.
// Inside the CTA the hand-off between row r and row r+1 (progress flag AND the four sample rows that
// cross the boundary) goes through shared memory; only every kDbfRows-th row boundary uses the global
// progress words + a device-scope fence.  A global hand-off costs several microseconds per macroblock
// step (measured), a shared-memory one ~100 cycles, and the wavefront pays it on every step.
constexpr int kDbfRows = 8;
constexpr int kDbfRing = 8;  // macroblocks a producer row may run ahead of the slots its consumer still reads

struct DbfSmem {
    uint8_t tile[kDbfRows][2 * kLHalf];                 // per-warp transpose tiles
    uint4 ring[kDbfRows][kDbfRing][2][4];               // [producer warp][slot][stream half][sample row 12..15] (luma)
    volatile int progress[kDbfRows];                    // macroblocks finished by each warp
    int ticket;
};

__device__ __forceinline__ void wait_smem(volatile int *flag, int need, int lane)
{
    if (lane == 0) {
        // exponential back-off: rows that have not started yet must not steal issue slots from working warps
        int spins = 0;
        while (*flag < need) __nanosleep(++spins < 64 ? 20 : 1000);
    }
    __threadfence_block();
    __syncwarp();
}
__device__ __forceinline__ void wait_global(const int *prog, int need, bool poll)
{
    if (poll) {
        int spins = 0;
        while (ld_acquire(prog) < need) __nanosleep(++spins < 32 ? 20 : 1000);
    }
    __syncwarp();
}

struct RowCtx {
    int w, row;
    bool top_smem, top_glob, bottom_smem, bottom_glob;
};

__device__ __forceinline__ int ub(uint32_t w, int k) { return (int)__byte_perm(w, 0, 0x4440 + k); }  // byte k, zero-extended
__device__ __forceinline__ uint32_t pk4(int a, int b, int c, int d) { return (uint32_t)(a | (b << 8) | (c << 16) | (d << 24)); }

// shared-memory progress is counted in half macroblocks: 2x+1 = vertical edges of MB x done (so the
// previous MB's columns 12..15 are final), 2x+2 = MB x done.  Row r+1 may filter the top edge of MB x as
// soon as row r has finished the VERTICAL edges of MB x+1 -- a lag of 1.5 instead of 2 macroblocks.
__device__ __forceinline__ void publish_smem(volatile int *flag, int v, int lane)
{
    __threadfence_block();
    __syncwarp();
    if (lane == 0) *flag = v;
}

// ------------------------------------------------------------------------------ luma rows
__device__ __forceinline__ void deblock_luma_row(DbfSmem &sm, const RowCtx rc, const FrameDesc &fd, const Geometry &g, int half,
                                                 int i, bool act)
{
    const int lane = threadIdx.x & 31, row = rc.row, w = rc.w;
    uint8_t *T = sm.tile[w] + half * kLHalf;  // this half's tile, row r at T + (r + 4) * kLS
    int *prog = fd.row_progress + g.mb_h;     // [1]: luma deblock wavefront (CTA boundaries only, in macroblocks)
    uint8_t *grow = fd.cur[0] + (ptrdiff_t)(16 * row + i) * g.y_stride;  // this thread's sample row
    uint8_t *gtop = fd.cur[0] + (ptrdiff_t)(16 * row - 4 + (i & 3)) * g.y_stride;
    const DeblockSide *side = fd.dbf_bs + (size_t)row * g.mb_w;
    uint32_t left = 0;                        // columns -4..-1 of this row (previous MB's 12..15)
    uint4 own = make_uint4(0, 0, 0, 0), prm = make_uint4(0, 0, 0, 0);
    uint32_t bsw = 0;
    if (act) {
        own = __ldcg(reinterpret_cast<const uint4 *>(grow));
        bsw = __ldg(&side[0].bs[i >> 2]);
        prm = __ldg(reinterpret_cast<const uint4 *>(side[0].luma));
    }
    const bool poll_g = act && i == 0 && rc.top_glob;
    // rows 13..15 are finished (and written) by the row below when it lives in this CTA
    const bool mine = !(rc.bottom_smem && i >= 13);
    const bool to_ring = rc.bottom_smem && i >= 12;
    bool dirty_prev = false;  // the previous MB changed samples that are still only in `left`

    for (int mbx = 0; mbx < g.mb_w; mbx++) {
        // ---- prefetch the next macroblock's row, strengths and parameters (no dependency on other rows)
        uint4 nxt = make_uint4(0, 0, 0, 0), prm_n = make_uint4(0, 0, 0, 0);
        uint32_t bsw_n = 0;
        if (act && mbx + 1 < g.mb_w) {
            nxt = __ldcg(reinterpret_cast<const uint4 *>(grow + 16 * (mbx + 1)));
            bsw_n = __ldg(&side[mbx + 1].bs[i >> 2]);
            prm_n = __ldg(reinterpret_cast<const uint4 *>(side[mbx + 1].luma));
        }
        const uint32_t bv = bsw & 0x0f0f0f0fu, bh = (bsw >> 4) & 0x0f0f0f0fu;  // byte e = bS of edge e
        const unsigned any_v = __ballot_sync(0xffffffffu, bv != 0), any_h = __ballot_sync(0xffffffffu, bh != 0);

        // ---- vertical edges: the whole row lives in registers, independent of the row above
        int v[20];
        {
            const uint32_t wd[5] = {left, own.x, own.y, own.z, own.w};
#pragma unroll
            for (int k = 0; k < 20; k++) v[k] = ub(wd[k >> 2], k & 3);
        }
        if (any_v) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int bs = (bv >> (8 * e)) & 0xff;
                const unsigned need = __ballot_sync(0xffffffffu, bs != 0);
                if (!need) continue;
                const unsigned strong = __ballot_sync(0xffffffffu, bs == 4);
                luma_edge_p(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3], v[4 * e + 4], v[4 * e + 5], v[4 * e + 6],
                            v[4 * e + 7], bs, e == 0 ? prm.x : prm.z, strong != 0);
            }
        }
        // columns -4..-1 are final now (the previous MB's horizontal edges were filtered already)
        if (mbx > 0) {
            const uint32_t lw = pk4(v[0], v[1], v[2], v[3]);
            if (act && mine && (dirty_prev || any_v)) __stcg(reinterpret_cast<uint32_t *>(grow + 16 * mbx - 4), lw);
            if (to_ring) reinterpret_cast<uint32_t *>(&sm.ring[w][(mbx - 1) % kDbfRing][half][i - 12])[3] = lw;
        }
        if (rc.bottom_smem) publish_smem(&sm.progress[w], 2 * mbx + 1, lane);
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = pk4(v[4 + 4 * k], v[5 + 4 * k], v[6 + 4 * k], v[7 + 4 * k]);

        // ---- the rows above become readable once row-1 has done the vertical edges of the next macroblock
        uint4 top = make_uint4(0, 0, 0, 0);
        if (rc.top_smem) {
            wait_smem(&sm.progress[w - 1], min(2 * mbx + 3, 2 * g.mb_w), lane);
            if (i < 4) top = sm.ring[w - 1][mbx % kDbfRing][half][i];
        } else {
            wait_global(prog + row - 1, min(mbx + 2, g.mb_w), poll_g);
            if (act && rc.top_glob && i < 4 && any_h) top = __ldcg(reinterpret_cast<const uint4 *>(gtop + 16 * mbx));
        }
        unsigned top_edge = 0;
        if (any_h) {
            // transpose through shared memory, horizontal edges, transpose back
#pragma unroll
            for (int k = 0; k < 4; k++) *reinterpret_cast<uint32_t *>(T + (i + 4) * kLS + 4 * k) = r[k];
            if (i < 4) {
                *reinterpret_cast<uint32_t *>(T + i * kLS + 0) = top.x;
                *reinterpret_cast<uint32_t *>(T + i * kLS + 4) = top.y;
                *reinterpret_cast<uint32_t *>(T + i * kLS + 8) = top.z;
                *reinterpret_cast<uint32_t *>(T + i * kLS + 12) = top.w;
            }
            __syncwarp();
            int c[20];
#pragma unroll
            for (int k = 0; k < 20; k++) c[k] = T[k * kLS + i];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int bs = (bh >> (8 * e)) & 0xff;
                const unsigned need = __ballot_sync(0xffffffffu, bs != 0);
                if (e == 0) top_edge = need;
                if (!need) continue;
                const unsigned strong = __ballot_sync(0xffffffffu, bs == 4);
                luma_edge_p(c[4 * e], c[4 * e + 1], c[4 * e + 2], c[4 * e + 3], c[4 * e + 4], c[4 * e + 5], c[4 * e + 6],
                            c[4 * e + 7], bs, e == 0 ? prm.y : prm.z, strong != 0);
            }
#pragma unroll
            for (int k = 1; k < 20; k++) T[k * kLS + i] = (uint8_t)c[k];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) r[k] = *reinterpret_cast<const uint32_t *>(T + (i + 4) * kLS + 4 * k);
            if (i < 4) {
                top.x = *reinterpret_cast<const uint32_t *>(T + i * kLS + 0);
                top.y = *reinterpret_cast<const uint32_t *>(T + i * kLS + 4);
                top.z = *reinterpret_cast<const uint32_t *>(T + i * kLS + 8);
                top.w = *reinterpret_cast<const uint32_t *>(T + i * kLS + 12);
            }
            __syncwarp();
        }
        // rows -3..-1 of the macroblock above: always ours to write when that row handed them over in
        // shared memory, otherwise only when the top edge changed them
        if (act && i >= 1 && i < 4 && (rc.top_smem || (rc.top_glob && top_edge)))
            __stcg(reinterpret_cast<uint4 *>(gtop + 16 * mbx), top);
        // columns 0..11 are final for this row; 12..15 wait for the next MB's left edge
        const bool last = mbx == g.mb_w - 1;
        if (act && mine && (any_v | any_h)) {
            __stcg(reinterpret_cast<uint2 *>(grow + 16 * mbx), make_uint2(r[0], r[1]));
            __stcg(reinterpret_cast<uint32_t *>(grow + 16 * mbx + 8), r[2]);
            if (last) __stcg(reinterpret_cast<uint32_t *>(grow + 16 * mbx + 12), r[3]);
        }
        if (rc.bottom_smem) {
            // the consumer must have finished the macroblock that used this ring slot before
            if (mbx >= kDbfRing) wait_smem(&sm.progress[w + 1], 2 * (mbx - kDbfRing) + 2, lane);
            if (i >= 12) {
                uint32_t *slot = reinterpret_cast<uint32_t *>(&sm.ring[w][mbx % kDbfRing][half][i - 12]);
                slot[0] = r[0], slot[1] = r[1], slot[2] = r[2];
                if (last) slot[3] = r[3];
            }
        }
        left = r[3];
        own = nxt;
        bsw = bsw_n;
        prm = prm_n;
        dirty_prev = (any_v | any_h) != 0;
        if (rc.bottom_glob) __threadfence();
        publish_smem(&sm.progress[w], 2 * mbx + 2, lane);
        if (rc.bottom_glob && act && i == 0) st_release(prog + row, mbx + 1);
    }
}

// ---------------------------------------------------------------------------- chroma rows
// ring slot reuse for chroma: uint4 ring[..][half][plane*2 + (row 6|7)] holds 8 samples in .x/.y
__device__ __forceinline__ void deblock_chroma_row(DbfSmem &sm, const RowCtx rc, const FrameDesc &fd, const Geometry &g, int half,
                                                   int i, bool act)
{
    const int lane = threadIdx.x & 31, row = rc.row, w = rc.w;
    const int pl = i >> 3, l = i & 7;          // plane (0 Cb, 1 Cr), line
    uint8_t *Th = sm.tile[w] + half * kCHalf;
    uint8_t *T = Th + pl * kCPlane;            // row r at T + (r + 2) * kCS2
    int *prog = fd.row_progress + 2 * g.mb_h;  // [2]: chroma deblock wavefront (CTA boundaries only)
    uint8_t *grow = fd.cur[1 + pl] + (ptrdiff_t)(8 * row + l) * g.c_stride;
    // threads 0..3 of a half also move the two rows above: (plane i>>1, row -2 + (i&1))
    uint8_t *gtop = fd.cur[1 + ((i >> 1) & 1)] + (ptrdiff_t)(8 * row - 2 + (i & 1)) * g.c_stride;
    const DeblockSide *side = fd.dbf_bs + (size_t)row * g.mb_w;
    uint32_t left = 0;
    uint2 own = make_uint2(0, 0);
    uint4 prm = make_uint4(0, 0, 0, 0);
    uint32_t bsw = 0;
    if (act) {
        own = __ldcg(reinterpret_cast<const uint2 *>(grow));
        bsw = __ldg(&side[0].bs[l >> 1]);
        prm = __ldg(reinterpret_cast<const uint4 *>(side[0].chroma));
    }
    const bool poll_g = act && i == 0 && rc.top_glob;
    const bool mine = !(rc.bottom_smem && l == 7);  // row 7 is finished by the row below inside a CTA
    const bool to_ring = rc.bottom_smem && l >= 6;
    bool dirty_prev = false;

    for (int mbx = 0; mbx < g.mb_w; mbx++) {
        uint2 nxt = make_uint2(0, 0);
        uint4 prm_n = make_uint4(0, 0, 0, 0);
        uint32_t bsw_n = 0;
        if (act && mbx + 1 < g.mb_w) {
            nxt = __ldcg(reinterpret_cast<const uint2 *>(grow + 8 * (mbx + 1)));
            bsw_n = __ldg(&side[mbx + 1].bs[l >> 1]);
            prm_n = __ldg(reinterpret_cast<const uint4 *>(side[mbx + 1].chroma));
        }
        // only even luma edges (0 and 2) touch chroma (core/frame.c:597,620)
        const int bv0 = bsw & 0xf, bv2 = (bsw >> 16) & 0xf, bh0 = (bsw >> 4) & 0xf, bh2 = (bsw >> 20) & 0xf;
        const unsigned any_v = __ballot_sync(0xffffffffu, (bv0 | bv2) != 0), any_h = __ballot_sync(0xffffffffu, (bh0 | bh2) != 0);
        int v[12];
        {
            const uint32_t wd[3] = {left, own.x, own.y};
#pragma unroll
            for (int k = 0; k < 12; k++) v[k] = ub(wd[k >> 2], k & 3);
        }
        if (any_v) {
            chroma_edge_p(v[2], v[3], v[4], v[5], bv0, prm.x);
            chroma_edge_p(v[6], v[7], v[8], v[9], bv2, prm.z);
        }
        if (mbx > 0) {
            const uint32_t lw = pk4(v[0], v[1], v[2], v[3]);
            if (act && mine && (dirty_prev || any_v)) __stcg(reinterpret_cast<uint32_t *>(grow + 8 * mbx - 4), lw);
            if (to_ring) sm.ring[w][(mbx - 1) % kDbfRing][half][pl * 2 + (l - 6)].y = lw;
        }
        if (rc.bottom_smem) publish_smem(&sm.progress[w], 2 * mbx + 1, lane);
        uint32_t r0 = pk4(v[4], v[5], v[6], v[7]), r1 = pk4(v[8], v[9], v[10], v[11]);

        uint2 top = make_uint2(0, 0);
        if (rc.top_smem) {
            wait_smem(&sm.progress[w - 1], min(2 * mbx + 3, 2 * g.mb_w), lane);
            if (i < 4) {
                const uint4 t4 = sm.ring[w - 1][mbx % kDbfRing][half][i];  // i = plane*2 + (row -2 | -1)
                top = make_uint2(t4.x, t4.y);
            }
        } else {
            wait_global(prog + row - 1, min(mbx + 2, g.mb_w), poll_g);
            if (act && rc.top_glob && i < 4 && any_h) top = __ldcg(reinterpret_cast<const uint2 *>(gtop + 8 * mbx));
        }
        unsigned top_edge = 0;
        if (any_h) {
            *reinterpret_cast<uint32_t *>(T + (l + 2) * kCS2) = r0;
            *reinterpret_cast<uint32_t *>(T + (l + 2) * kCS2 + 4) = r1;
            uint8_t *Tt = Th + ((i >> 1) & 1) * kCPlane + (i & 1) * kCS2;
            if (i < 4) {
                *reinterpret_cast<uint32_t *>(Tt) = top.x;
                *reinterpret_cast<uint32_t *>(Tt + 4) = top.y;
            }
            __syncwarp();
            int c[8];
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = T[k * kCS2 + l];  // column l of this plane, rows -2..5
            chroma_edge_p(c[0], c[1], c[2], c[3], bh0, prm.y);
            chroma_edge_p(c[4], c[5], c[6], c[7], bh2, prm.z);
            T[1 * kCS2 + l] = (uint8_t)c[1];
            T[2 * kCS2 + l] = (uint8_t)c[2];
            T[5 * kCS2 + l] = (uint8_t)c[5];
            T[6 * kCS2 + l] = (uint8_t)c[6];
            __syncwarp();
            r0 = *reinterpret_cast<const uint32_t *>(T + (l + 2) * kCS2);
            r1 = *reinterpret_cast<const uint32_t *>(T + (l + 2) * kCS2 + 4);
            if (i < 4) top = make_uint2(*reinterpret_cast<const uint32_t *>(Tt), *reinterpret_cast<const uint32_t *>(Tt + 4));
            top_edge = __ballot_sync(0xffffffffu, bh0 != 0);
            __syncwarp();
        }
        // row -1 of the macroblock above (p0 of the top edge), plane (i>>1)
        if (act && (i & 1) && i < 4 && (rc.top_smem || (rc.top_glob && top_edge)))
            __stcg(reinterpret_cast<uint2 *>(gtop + 8 * mbx), top);
        const bool last = mbx == g.mb_w - 1;
        if (act && mine && (any_v | any_h)) {
            __stcg(reinterpret_cast<uint32_t *>(grow + 8 * mbx), r0);
            if (last) __stcg(reinterpret_cast<uint32_t *>(grow + 8 * mbx + 4), r1);
        }
        if (rc.bottom_smem) {
            if (mbx >= kDbfRing) wait_smem(&sm.progress[w + 1], 2 * (mbx - kDbfRing) + 2, lane);
            if (l >= 6) {
                uint4 &slot = sm.ring[w][mbx % kDbfRing][half][pl * 2 + (l - 6)];
                slot.x = r0;
                if (last) slot.y = r1;
            }
        }
        left = r1;
        own = nxt;
        bsw = bsw_n;
        prm = prm_n;
        dirty_prev = (any_v | any_h) != 0;
        if (rc.bottom_glob) __threadfence();
        publish_smem(&sm.progress[w], 2 * mbx + 2, lane);
        if (rc.bottom_glob && act && i == 0) st_release(prog + row, mbx + 1);
    }
}

// grid: 2 roles x ceil(n_lanes/2) stream pairs x ceil(mb_h/kDbfRows) row groups, handed out by ticket in
// dependency order (the group above of the same pair and role always has a smaller ticket)
__global__ void __launch_bounds__(32 * kDbfRows, 4) deblock_kernel(const FrameDesc *__restrict__ descs, Geometry g, int n_lanes,
                                                                   int *ticket, int dbg)
{
    __shared__ __align__(16) DbfSmem sm;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) sm.ticket = atomicAdd(ticket, 1);
    if (threadIdx.x < kDbfRows) sm.progress[threadIdx.x] = 0;
    __syncthreads();
    const int t = sm.ticket;
    const int groups = (g.mb_h + kDbfRows - 1) / kDbfRows;
    const int role = t & 1, u = t >> 1;
    const int pair = u / groups, grp = u % groups;
    RowCtx rc;
    rc.w = w;
    rc.row = grp * kDbfRows + w;
    if (rc.row >= g.mb_h) return;
    rc.top_smem = w > 0;
    rc.top_glob = w == 0 && rc.row > 0;
    rc.bottom_smem = w + 1 < kDbfRows && rc.row + 1 < g.mb_h;
    rc.bottom_glob = !rc.bottom_smem && rc.row + 1 < g.mb_h;
    const int half = lane >> 4, i = lane & 15;
    const int stream = 2 * pair + half;
    const FrameDesc &fd = descs[min(stream, n_lanes - 1)];
    const bool act = stream < n_lanes && fd.deblock != 0;
    (void)dbg;
    if (role == 0)
        deblock_luma_row(sm, rc, fd, g, half, i, act);
    else
        deblock_chroma_row(sm, rc, fd, g, half, i, act);
}
#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
