// In-loop deblocking filter as a macroblock-row wavefront.
//
// Replaces p264_frame_deblocking_filter (core/frame.c:490-643), deblock_edge (:472-488) and the
// eight edge filters (:302-470).  The reference filters macroblocks in raster order, vertical
// edges then horizontal edges per macroblock; MB(x,y) therefore depends on MB(x-1,y) (whole MB)
// and on MB(x+1,y-1) (its left-edge filter rewrites columns 13..15 of MB(x,y-1), which MB(x,y)'s
// top-edge filter reads).  A picture-wide "all vertical, then all horizontal" pass is NOT
// bit-exact, so the kernel keeps the reference order as a two-macroblock-lag row wavefront.
//
// Work decomposition (v3; v2 spent 40 % of its issue slots spinning on shared-memory flags and most of
// the rest on one-sample-per-register arithmetic and byte shuffling):
//  * TWO sample lines per register: every tap (p3..q3) is held as s16x2, the filters are the packed
//    forms in swar.cuh (VABSDIFF4 / VIADD.16 / VIMNMX.S16x2 / VIADDMNMX.RELU), so one thread filters
//    two rows (vertical edges) or two columns (horizontal edges) per instruction stream;
//  * a warp owns one macroblock row of FOUR lanes (streams): 8 threads per stream, always in the
//    same code path.  Luma and chroma are independent given the boundary strengths and run as
//    separate warps (role = CTA) with separate progress flags;
//  * a CTA owns kDbfRows consecutive macroblock rows and runs them in LOCKSTEP: two __syncthreads per
//    macroblock step (vertical edges | horizontal edges), warp w works on macroblock (step - w): a
//    ONE-macroblock lag between rows.  Waiting warps sit in the barrier instead of polling.  The four sample rows that cross a row boundary are handed down through a
//    small shared-memory ring; only every kDbfRows-th row boundary goes through global memory
//    (progress word + acquire/release), handled by a ninth "I/O" warp so that no filtering warp ever
//    executes a fence or polls;
//  * vertical edges are filtered in registers straight from two 16-byte row loads; the transpose for
//    the horizontal edges is a shared-memory tile written as rows and read as 16-bit column pairs;
//  * all global traffic is full 16-byte (luma) / 8-byte (chroma) rows: a macroblock's rows are stored
//    once, after the next macroblock's left edge has finalised their last three columns.
#pragma once
#include "common.cuh"
#include "swar.cuh"

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


constexpr int kDbfRows = 8;     // macroblock rows (warps) per CTA
constexpr int kDbfQuad = 4;     // lanes (streams) per warp, 8 threads each
constexpr int kDbfRing = 4;     // hand-off slots per producer row (the consumer runs two macroblocks behind)
constexpr int kDbfTile = 272;   // bytes per (warp, stream) transpose tile: 16 rows x 16 B, + 16 B bank skew
constexpr int kDbfSlot = 80;    // bytes per (slot, stream): 4 rows x 16 B, + 16 B bank skew

// Optional per-step cycle trace of ONE CTA (engine debug knob P264B200_TRACE=<ticket>): [warp][step][marks]:
// 0 after the first barrier of a step, 1 before the second, 2 after it, 3 at the end of the step
constexpr int kDbfTraceSteps = 320;
__device__ long long g_dbf_trace[kDbfRows + 1][kDbfTraceSteps][6];
__device__ long long g_dbf_cta_ns[2048][4];  // per ticket: %globaltimer at kernel entry, first step, last step, exit
__device__ __forceinline__ long long dbf_now_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void dbf_mark(bool on, int w, int s, int k)
{
    if (on && s < kDbfTraceSteps && (threadIdx.x & 31) == 0) g_dbf_trace[w][s][k] = clock64();
}

struct DbfSmem {
    uint8_t tile[kDbfRows][kDbfQuad][kDbfTile];              // luma: row r at 16r; chroma: plane p row r at 64p + 8r
    uint8_t ring[kDbfRows][kDbfRing][kDbfQuad][kDbfSlot];    // luma: rows 12..15 at 16k; chroma: plane p rows 6,7 at 16p + 8k
    uint8_t topring[kDbfRing][kDbfQuad][kDbfSlot];            // same layout: rows above warp 0, fetched from global memory by the I/O warp
    int ticket;
};

// Edge parameters of one direction of one macroblock, resolved once per thread: the macroblock edge (left or
// top) and the three inner edges share alpha / beta per kind; only tc0 and the on/off switch follow bS.
struct DbfDir {
    uint32_t prm_mb, prm_in;       // packed parameters (see DeblockSide) of the macroblock edge / the inner edges
    uint32_t na_mb, nb_mb;         // -alpha, -beta in both fields
    uint32_t na_in, nb_in;
};
__device__ __forceinline__ DbfDir dbf_dir(uint32_t prm_mb, uint32_t prm_in)
{
    DbfDir d;
    d.prm_mb = prm_mb, d.prm_in = prm_in;
    d.na_mb = swar::rep2(-(int)(prm_mb & 0xff)), d.nb_mb = swar::rep2(-(int)((prm_mb >> 8) & 31));
    d.na_in = swar::rep2(-(int)(prm_in & 0xff)), d.nb_in = swar::rep2(-(int)((prm_in >> 8) & 31));
    return d;
}
// constants of edge e for boundary strength bs: filtering is switched off by alpha = 0 (bS 0; for luma also bS 4,
// which the strong filter handles)
__device__ __forceinline__ swar::EdgeK dbf_edge_k(const DbfDir &d, bool mb_edge, int bs, bool luma)
{
    const uint32_t prm = mb_edge ? d.prm_mb : d.prm_in;
    swar::EdgeK k;
    const bool on = luma ? (unsigned)(bs - 1) < 3u : bs != 0;
    k.n_alpha = on ? (mb_edge ? d.na_mb : d.na_in) : 0u;
    k.n_beta = mb_edge ? d.nb_mb : d.nb_in;
    k.tc0 = ((prm >> (8 + 5 * bs)) & 31) * swar::kOnes;
    return k;
}

// The four luma edges of one direction on two lines, v = 4 samples before the macroblock + its 16 samples.
// bs4: bS of edge e in bits 8e..8e+3.  The bS < 4 filters are straight-line code with no votes or branches
// between the edges, so the instruction scheduler can overlap them (edge e+1 needs edge e only through one tap).
__device__ __forceinline__ void dbf_luma_dir(uint32_t *v, uint32_t bs4, const DbfDir &d)
{
    if (!__any_sync(0xffffffffu, bs4 != 0)) return;
    const int bs0 = bs4 & 0xf;
    if (__any_sync(0xffffffffu, bs0 == 4)) {
        // intra macroblock edge: the strong filter, only on the lines that have bS 4 (the normal one is off there)
        const int alpha = d.prm_mb & 0xff;
        swar::EdgeK k4;
        k4.n_alpha = bs0 == 4 ? d.na_mb : 0u, k4.n_beta = d.nb_mb, k4.tc0 = 0;
        swar::luma_strong(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], k4, alpha);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const swar::EdgeK k = dbf_edge_k(d, e == 0, (bs4 >> (8 * e)) & 0xf, true);
        swar::luma_normal(v[4 * e + 1], v[4 * e + 2], v[4 * e + 3], v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], k);
    }
}
// The two chroma edges of one direction on two lines, v = p1 p0 | q0 q1 . . q0' q1' ...: edge 0 at v[2], edge 2 at v[6]
__device__ __forceinline__ void dbf_chroma_dir(uint32_t *v, int bs0, int bs2, const DbfDir &d)
{
    if (!__any_sync(0xffffffffu, (bs0 | bs2) != 0)) return;
    const swar::EdgeK k0 = dbf_edge_k(d, true, bs0, false), k2 = dbf_edge_k(d, false, bs2, false);
    uint32_t p0 = v[1], q0 = v[2];
    swar::chroma_edge2(v[0], p0, q0, v[3], k0, false);
    swar::chroma_edge2(v[4], v[5], v[6], v[7], k2, false);
    if (__any_sync(0xffffffffu, bs0 == 4)) {
        uint32_t sp0 = v[1], sq0 = v[2];
        swar::chroma_edge2(v[0], sp0, sq0, v[3], k0, true);
        if (bs0 == 4) p0 = sp0, q0 = sq0;
    }
    v[1] = p0, v[2] = q0;
}

template <int N> struct DbfVec;
template <> struct DbfVec<4> { typedef uint4 type; };
template <> struct DbfVec<2> { typedef uint2 type; };
__device__ __forceinline__ void vec_get(const uint4 &v, uint32_t *r) { r[0] = v.x, r[1] = v.y, r[2] = v.z, r[3] = v.w; }
__device__ __forceinline__ void vec_get(const uint2 &v, uint32_t *r) { r[0] = v.x, r[1] = v.y; }
__device__ __forceinline__ void vec_set(uint4 &v, const uint32_t *r) { v = make_uint4(r[0], r[1], r[2], r[3]); }
__device__ __forceinline__ void vec_set(uint2 &v, const uint32_t *r) { v = make_uint2(r[0], r[1]); }

// The macroblock rows of one CTA for one role.  C = false: luma (16 rows x 16 B per macroblock, 4 edges per
// direction, 4 rows handed down); C = true: Cb and Cr (per plane 8 rows x 8 B, 2 edges, 2 rows handed down,
// threads 0..3 of a stream on Cb, 4..7 on Cr).
template <bool C>
__device__ __forceinline__ void deblock_rows(DbfSmem &sm, const FrameDesc *__restrict__ descs, const Geometry &g, int n_lanes, int quad,
                                             int grp, bool trace, bool times_on, int tk)
{
    constexpr int NW = C ? 2 : 4;    // 32-bit words per sample row of a macroblock
    constexpr int RB = 4 * NW;       // bytes per row
    constexpr int NR = C ? 8 : 16;   // rows per plane
    constexpr int TR = C ? 2 : 4;    // rows handed down to the macroblock row below (per plane)
    constexpr int NP = 4 + 4 * NW;   // taps along a row incl. the 4 samples left of the macroblock
    constexpr int NQ = TR + NR;      // taps down a column incl. the rows above
    typedef typename DbfVec<NW>::type Vec;
    using swar::prmt;

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, t = lane & 7;
    const int pl = C ? t >> 2 : 0, j = C ? t & 3 : t;  // plane; row pair (vertical edges) = column pair (horizontal edges)
    const int row = grp * kDbfRows + w;
    const bool row_ok = row < g.mb_h;
    const bool has_top = row > 0;
    const bool top_smem = w > 0;
    const bool bottom_smem = w + 1 < kDbfRows && row + 1 < g.mb_h;
    const int stream = kDbfQuad * quad + sub;
    const FrameDesc &fd = descs[min(stream, n_lanes - 1)];
    const bool act = row_ok && stream < n_lanes && fd.deblock != 0;
    const int stride = C ? g.c_stride : g.y_stride;
    // this thread's two sample rows, and (threads 0..3) the row above it moves between global and shared memory:
    // luma row -4 + t; chroma plane t >> 1, row -2 + (t & 1)
    uint8_t *grow = fd.cur[C ? 1 + pl : 0] + (ptrdiff_t)(NR * row + 2 * j) * stride;
    uint8_t *gtop = C ? fd.cur[1 + ((t >> 1) & 1)] + (ptrdiff_t)(NR * row - 2 + (t & 1)) * stride : fd.cur[0] + (ptrdiff_t)(NR * row - 4 + (t & 3)) * stride;
    const int top_off = C ? 16 * ((t >> 1) & 1) + 8 * (t & 1) : 16 * (t & 3);  // of that row inside a slot
    const DeblockSide *side = fd.dbf_bs + (size_t)row * g.mb_w;
    uint8_t *T = sm.tile[w][sub] + (C ? 64 * pl : 0);
    // rows this thread stores itself / hands to the row below through the ring
    const bool store_a = !(bottom_smem && 2 * j > NR - TR), store_b = !(bottom_smem && 2 * j + 1 > NR - TR);
    const bool to_ring = bottom_smem && 2 * j >= NR - TR;
    const int ring_off = (C ? 16 * pl : 0) + RB * (2 * j - (NR - TR));

    uint32_t prev[2][NW], cur[2][NW], nxt[2][NW];
    uint32_t bsw = 0, bsw_n = 0;
    uint4 prm = make_uint4(0, 0, 0, 0), prm_n = prm;
#pragma unroll
    for (int k = 0; k < NW; k++) prev[0][k] = prev[1][k] = cur[0][k] = cur[1][k] = nxt[0][k] = nxt[1][k] = 0;
    if (act) {
        vec_get(__ldcg(reinterpret_cast<const Vec *>(grow)), cur[0]);
        vec_get(__ldcg(reinterpret_cast<const Vec *>(grow + stride)), cur[1]);
        bsw = __ldg(&side[0].bs[j >> (C ? 0 : 1)]);
        prm = __ldg(reinterpret_cast<const uint4 *>(C ? side[0].chroma : side[0].luma));
    }

    // Lockstep schedule, one-macroblock lag: step s = [barrier] vertical edges of MB (s - w) [barrier] horizontal
    // edges of MB (s - w).  The left-edge filter of MB x+1 finalises the last columns of MB x in the first half of
    // a step, so the row below can filter the top edge of its MB x in the second half of the same step.
    const int n_steps = g.mb_w + (kDbfRows - 1);
#pragma unroll 1
    for (int s = 0; s < n_steps; s++) {
        const int x = s - w;
        const bool on = x >= 0 && x < g.mb_w && row_ok;
        const bool last = x == g.mb_w - 1;
        dbf_mark(trace, w, s, 3);
        __syncthreads();
        dbf_mark(trace, w, s, 0);
        if (times_on && threadIdx.x == 0 && (s == 0 || s == n_steps - 1)) g_dbf_cta_ns[tk][s == 0 ? 1 : 2] = dbf_now_ns();
        if (on) {
            // ---- prefetch the next macroblock's rows, strengths and parameters (and, once per 128-byte line, pull
            // the next line into L2 so that those loads do not wait on HBM inside the dependent chain)
            if (act && (x & (128 / RB - 1)) == 0 && x + 128 / RB < g.mb_w) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + RB * x + 128));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + stride + RB * x + 128));
            }
            if (act && !last) {
                bsw_n = ldg_now_u32(&side[x + 1].bs[j >> (C ? 0 : 1)]);
                {
                    // (two loads with no unused component: the register of an unused one would be recycled as scratch
                    // while the prefetch is still in flight, and that write has to wait out the whole memory latency)
                    const uint32_t *pp = C ? side[x + 1].chroma : side[x + 1].luma;
                    const uint2 xy = ldg_now_v2(pp);
                    prm_n.x = xy.x, prm_n.y = xy.y, prm_n.z = ldg_now_u32(pp + 2);
                }
                // (this row's samples were written by the previous kernel; nobody else touches them before we do)
                Vec va, vb;
                ldg_now(grow + RB * (x + 1), va);
                ldg_now(grow + stride + RB * (x + 1), vb);
                vec_get(va, nxt[0]);
                vec_get(vb, nxt[1]);
            }
            // ---- vertical edges: taps of this thread's two rows, two rows per register
            {
                uint32_t P[NP];
#pragma unroll
                for (int wd = 0; wd < 1 + NW; wd++) {
                    const uint32_t wa = wd == 0 ? prev[0][NW - 1] : cur[0][wd - 1], wb = wd == 0 ? prev[1][NW - 1] : cur[1][wd - 1];
                    const uint32_t t01 = prmt(wa, wb, 0x5140), t23 = prmt(wa, wb, 0x7362);
                    P[4 * wd + 0] = prmt(t01, 0, 0x4140);
                    P[4 * wd + 1] = prmt(t01, 0, 0x4342);
                    P[4 * wd + 2] = prmt(t23, 0, 0x4140);
                    P[4 * wd + 3] = prmt(t23, 0, 0x4342);
                }
                const DbfDir dv = dbf_dir(prm.x, prm.z);
                if (C)
                    dbf_chroma_dir(P + 2, bsw & 0xf, (bsw >> 16) & 0xf, dv);
                else
                    dbf_luma_dir(P, bsw & 0x0f0f0f0fu, dv);
#pragma unroll
                for (int wd = 0; wd < 1 + NW; wd++) {
                    const uint32_t t01 = prmt(P[4 * wd + 0], P[4 * wd + 1], 0x6240), t23 = prmt(P[4 * wd + 2], P[4 * wd + 3], 0x6240);
                    const uint32_t wa = prmt(t01, t23, 0x5410), wb = prmt(t01, t23, 0x7632);
                    if (wd == 0)
                        prev[0][NW - 1] = wa, prev[1][NW - 1] = wb;
                    else
                        cur[0][wd - 1] = wa, cur[1][wd - 1] = wb;
                }
            }
            // ---- the previous macroblock's rows are final now (its last columns just saw this left edge)
            if (x > 0) {
                Vec va, vb;
                vec_set(va, prev[0]);
                vec_set(vb, prev[1]);
                if (act && store_a) *reinterpret_cast<Vec *>(grow + RB * (x - 1)) = va;
                if (act && store_b) *reinterpret_cast<Vec *>(grow + stride + RB * (x - 1)) = vb;
                if (to_ring) {
                    uint8_t *slot = sm.ring[w][(x - 1) & (kDbfRing - 1)][sub] + ring_off;
                    *reinterpret_cast<Vec *>(slot) = va;
                    *reinterpret_cast<Vec *>(slot + RB) = vb;
                }
            }
            // ---- transpose through shared memory: rows in, 16-bit column pairs out
            {
                Vec va, vb;
                vec_set(va, cur[0]);
                vec_set(vb, cur[1]);
                *reinterpret_cast<Vec *>(T + RB * (2 * j)) = va;
                *reinterpret_cast<Vec *>(T + RB * (2 * j + 1)) = vb;
            }
        }
        dbf_mark(trace, w, s, 1);
        __syncthreads();
        dbf_mark(trace, w, s, 2);
        if (!on) continue;
        uint8_t *slot_top = top_smem ? sm.ring[w - 1][x & (kDbfRing - 1)][sub] : sm.topring[x & (kDbfRing - 1)][sub];
        uint8_t *topp = slot_top + (C ? 16 * pl : 0);
        {
            uint32_t Q[NQ];
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const uint8_t *src = k < TR ? topp + RB * k + 2 * j : T + RB * (k - TR) + 2 * j;
                Q[k] = (k < TR && !has_top) ? 0u : prmt(*reinterpret_cast<const uint16_t *>(src), 0, 0x4140);
            }
            const DbfDir dh = dbf_dir(prm.y, prm.z);
            if (C) {
                dbf_chroma_dir(Q, (bsw >> 4) & 0xf, (bsw >> 20) & 0xf, dh);
                if (has_top) *reinterpret_cast<uint16_t *>(topp + RB * 1 + 2 * j) = (uint16_t)prmt(Q[1], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 0 + 2 * j) = (uint16_t)prmt(Q[2], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 3 + 2 * j) = (uint16_t)prmt(Q[5], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 4 + 2 * j) = (uint16_t)prmt(Q[6], 0, 0x4420);
            } else {
                dbf_luma_dir(Q, (bsw >> 4) & 0x0f0f0f0fu, dh);
#pragma unroll
                for (int k = 1; k < NQ - 1; k++) {
                    uint8_t *dst = k < TR ? topp + RB * k + 2 * j : T + RB * (k - TR) + 2 * j;
                    if (k >= TR || has_top) *reinterpret_cast<uint16_t *>(dst) = (uint16_t)prmt(Q[k], 0, 0x4420);
                }
            }
        }
        __syncwarp();
        // ---- back to rows: these wait in registers for the next macroblock's left edge
        vec_get(*reinterpret_cast<const Vec *>(T + RB * (2 * j)), prev[0]);
        vec_get(*reinterpret_cast<const Vec *>(T + RB * (2 * j + 1)), prev[1]);
        // the rows above are finished: luma rows -3..-1, chroma row -1 (rows -4 / -2 are only read)
        if (has_top && act && t < 4 && (C ? (t & 1) : (t != 0)))
            *reinterpret_cast<Vec *>(gtop + RB * x) = *reinterpret_cast<const Vec *>(slot_top + top_off);
        if (last) {
            Vec va, vb;
            vec_set(va, prev[0]);
            vec_set(vb, prev[1]);
            if (act && store_a) *reinterpret_cast<Vec *>(grow + RB * x) = va;
            if (act && store_b) *reinterpret_cast<Vec *>(grow + stride + RB * x) = vb;
            if (to_ring) {
                uint8_t *slot = sm.ring[w][x & (kDbfRing - 1)][sub] + ring_off;
                *reinterpret_cast<Vec *>(slot) = va;
                *reinterpret_cast<Vec *>(slot + RB) = vb;
            }
        }
#pragma unroll
        for (int k = 0; k < NW; k++) cur[0][k] = nxt[0][k], cur[1][k] = nxt[1][k];
        bsw = bsw_n;
        prm = prm_n;
    }
}

// The I/O warp of a CTA (warp kDbfRows): everything that touches the global progress words, so that no
// filtering warp ever executes a fence or polls.  Per lockstep step s it
//  * (consumer side, row group > 0) keeps the rows above warp 0 two macroblocks ahead in sm.topring: waits for
//    the row group above to have stored macroblock s+2, loads its last rows, and drops them into the ring one
//    step later (warp 0 reads slot s at step s);
//  * (producer side) publishes how many macroblocks of the CTA's last row are in global memory.  The barrier
//    orders that row's stores before this warp's fence + release (fence cumulativity), one step behind.
template <bool C>
__device__ __forceinline__ void deblock_io_warp(DbfSmem &sm, const FrameDesc *__restrict__ descs, const Geometry &g, int n_lanes, int quad,
                                                int grp, bool trace)
{
    constexpr int RB = C ? 8 : 16, NR = C ? 8 : 16;
    typedef typename DbfVec<RB / 4>::type Vec;
    const int lane = threadIdx.x & 31, sub = (lane >> 2) & 3, k = lane & 3;  // lanes 0..15: (stream, row above)
    const int row0 = grp * kDbfRows;                 // first macroblock row of the CTA
    const int last_w = kDbfRows - 1, row_last = row0 + last_w;
    const bool consumer = grp > 0, producer = row_last + 1 < g.mb_h;
    const int stream = kDbfQuad * quad + sub;
    const FrameDesc &fd = descs[min(stream, n_lanes - 1)];
    const bool act = lane < 16 && stream < n_lanes && fd.deblock != 0;
    int *prog = descs[kDbfQuad * quad].row_progress + (C ? 2 : 1) * g.mb_h;  // one progress word per (quad, role, row)
    const int stride = C ? g.c_stride : g.y_stride;
    // luma: row -4 + k; chroma: plane k >> 1, row -2 + (k & 1)
    const uint8_t *gtop = C ? fd.cur[1 + (k >> 1)] + (ptrdiff_t)(NR * row0 - 2 + (k & 1)) * stride : fd.cur[0] + (ptrdiff_t)(NR * row0 - 4 + k) * stride;
    const int top_off = C ? 16 * (k >> 1) + 8 * (k & 1) : 16 * k;
    int seen = 0;
    auto wait_for = [&](int need) {
        if (seen < need) {
            if (lane == 0) {
                int spins = 0;
                while ((seen = ld_acquire(prog + row0 - 1)) < need) __nanosleep(++spins < 16 ? 40 : 400);
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
        }
    };
    Vec pend;
    {
        uint32_t z[4] = {0, 0, 0, 0};
        vec_set(pend, z);
    }
    if (consumer) {
        // macroblocks 0 and 1 before the first step
        wait_for(min(2, g.mb_w));
        if (act) {
            *reinterpret_cast<Vec *>(sm.topring[0][sub] + top_off) = __ldcg(reinterpret_cast<const Vec *>(gtop));
            if (g.mb_w > 1) *reinterpret_cast<Vec *>(sm.topring[1][sub] + top_off) = __ldcg(reinterpret_cast<const Vec *>(gtop + RB));
        }
    }
    const int n_steps = g.mb_w + (kDbfRows - 1);
    int peek = 0;  // progress word read asynchronously during the previous step
#pragma unroll 1
    for (int s = 0; s <= n_steps; s++) {
        dbf_mark(trace, kDbfRows, s, 3);
        __syncthreads();  // s == n_steps: the extra barrier after the loop of the filtering warps
        dbf_mark(trace, kDbfRows, s, 0);
        if (consumer && s >= 1 && s + 1 < g.mb_w && act) *reinterpret_cast<Vec *>(sm.topring[(s + 1) & (kDbfRing - 1)][sub] + top_off) = pend;
        if (producer) {
            // the last row worked on macroblock x7 during step s-1: macroblocks [0, x7) are stored, all after the last.
            // (before this step's loads are issued, so that the fence does not wait for them)
            const int x7 = s - 1 - last_w;
            if (x7 >= 1 && x7 < g.mb_w && lane == 0) {
                __threadfence();
                st_release(prog + row_last, x7 == g.mb_w - 1 ? g.mb_w : x7);
            }
        }
        if (consumer && s + 2 < g.mb_w) {
            seen = max(seen, __shfl_sync(0xffffffffu, peek, 0));
            wait_for(min(s + 3, g.mb_w));  // normally satisfied by the value peeked one step ago
            __threadfence();               // acquire side for the peeked (relaxed) value
            if (act) pend = __ldcg(reinterpret_cast<const Vec *>(gtop + RB * (s + 2)));
            if (lane == 0) peek = ld_relaxed(prog + row0 - 1);
        }
        if (s < n_steps) {
            dbf_mark(trace, kDbfRows, s, 1);
            __syncthreads();
            dbf_mark(trace, kDbfRows, s, 2);
        }
    }
}

// grid: 2 roles x ceil(n_lanes/4) stream quads x ceil(mb_h/kDbfRows) row groups, handed out by ticket in
// dependency order (the group above of the same quad and role always has a smaller ticket)
__global__ void __launch_bounds__(32 * (kDbfRows + 1), 2) deblock_kernel(const FrameDesc *__restrict__ descs, Geometry g, int n_lanes, int *ticket, int trace_ticket)
{
    __shared__ __align__(16) DbfSmem sm;
    if (threadIdx.x == 0) sm.ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int tk = sm.ticket;
    const int groups = (g.mb_h + kDbfRows - 1) / kDbfRows;
    const int role = tk & 1, u = tk >> 1;
    const int quad = u / groups, grp = u % groups;
    const bool trace = tk == trace_ticket;
    const bool times = trace_ticket >= 0 && tk < 2048 && threadIdx.x == 0;
    if (times) g_dbf_cta_ns[tk][0] = dbf_now_ns();
    if (threadIdx.x >= 32 * kDbfRows) {
        if (role == 0)
            deblock_io_warp<false>(sm, descs, g, n_lanes, quad, grp, trace);
        else
            deblock_io_warp<true>(sm, descs, g, n_lanes, quad, grp, trace);
        return;
    }
    if (role == 0)
        deblock_rows<false>(sm, descs, g, n_lanes, quad, grp, trace, trace_ticket >= 0 && tk < 2048, tk);
    else
        deblock_rows<true>(sm, descs, g, n_lanes, quad, grp, trace, trace_ticket >= 0 && tk < 2048, tk);
    if (times) g_dbf_cta_ns[tk][3] = dbf_now_ns();
    __syncthreads();  // lets the I/O warp publish the last row's final macroblock
}
#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
