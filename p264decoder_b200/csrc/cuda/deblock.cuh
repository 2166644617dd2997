// In-loop deblocking filter as a macroblock-row wavefront.
//
// Replaces p264_frame_deblocking_filter (core/frame.c:490-643), deblock_edge (:472-488) and the
// eight edge filters (:302-470).  The reference filters macroblocks in raster order, vertical
// edges then horizontal edges per macroblock; MB(x,y) therefore depends on MB(x-1,y) (whole MB)
// and on MB(x+1,y-1) (its left-edge filter rewrites columns 13..15 of MB(x,y-1), which MB(x,y)'s
// top-edge filter reads).  A picture-wide "all vertical, then all horizontal" pass is NOT
// bit-exact, so the kernel keeps the reference order as a row wavefront with a one-macroblock lag.
//
// Work decomposition (v4):
//  * TWO sample lines per register: every tap (p3..q3) is held as s16x2, the filters are the packed
//    forms in swar.cuh (VABSDIFF4 / VIADD.16 / VIMNMX.S16x2 / VIADDMNMX.RELU), so one thread filters
//    two rows (vertical edges) or two columns (horizontal edges) per instruction stream;
//  * a warp owns one macroblock row of FOUR lanes (streams): 8 threads per stream, always in the
//    same code path.  Luma and chroma are independent given the boundary strengths and run as
//    separate CTAs (roles) with separate progress words; a CTA draws its role from its arrival order on
//    its SM so that every SM runs one luma and one chroma CTA;
//  * a CTA owns kDbfRows consecutive macroblock rows.  Each row warp runs at its own pace; the only
//    cross-warp dependency -- the top edge of macroblock x needs the last rows of macroblock x of the
//    row above after the left edge of ITS macroblock x+1 -- goes through a shared-memory ring of
//    kDbfRing slots guarded by full/empty mbarriers (one arriving lane, hardware-suspended try_wait:
//    no polling, no CTA-wide barrier).  The left macroblock edge is filtered first and the finished
//    macroblock handed down before the inner edges, which halves the row-to-row lag of the wavefront;
//  * only every kDbfRows-th row boundary goes through global memory (progress word per row, role and
//    stream quad).  Two helper warps own those words: the "in" warp polls the row group above (relaxed
//    loads, one fence per hand-over -- an acquire load per spin would flush the SM's L1 each time) and
//    keeps ring[0] up to kDbfRing macroblocks ahead; the "out" thread publishes the CTA's last row.
//    No filtering warp ever executes a fence or polls;
//  * vertical edges are filtered in registers straight from two 16-byte row loads (prefetched one
//    macroblock ahead, the next 128-byte line pulled into L2); the transpose for the horizontal edges
//    is a private shared-memory tile written as rows and read as 16-bit column pairs;
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

// bS of one 4-sample segment from values already in registers (same rule as boundary_strength above)
__device__ __forceinline__ int bs_of(bool intra_any, int strong, unsigned coded, int rq, int rp, int vq, int vp)
{
    if (intra_any) return strong;
    if (coded & 1u) return 2;
    const int dx = abs((int)(short)(vq & 0xffff) - (int)(short)(vp & 0xffff)), dy = abs((vq >> 16) - (vp >> 16));
    return (rq != rp || dx >= 4 || dy >= 4) ? 1 : 0;
}

// One thread per macroblock: its record (five 16-byte loads) plus the last column / row of the left / top
// neighbour give all 32 boundary strengths and the six edge-parameter words; three 16-byte stores.
__global__ void __launch_bounds__(128) deblock_bs_kernel(const FrameDesc *__restrict__ descs, Geometry g)
{
    const FrameDesc &fd = descs[blockIdx.y];
    if (!fd.deblock) return;
    const int mb_xy = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb_xy >= g.mb_w * g.mb_h) return;
    const int mby = mb_xy / g.mb_w, mbx = mb_xy - mby * g.mb_w;
    const uint4 *rec = reinterpret_cast<const uint4 *>(fd.mbs + mb_xy);
    int mv[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 q = __ldg(rec + i);
        mv[4 * i] = (int)q.x, mv[4 * i + 1] = (int)q.y, mv[4 * i + 2] = (int)q.z, mv[4 * i + 3] = (int)q.w;
    }
    const uint4 h = __ldg(rec + 4);  // ref[4] | mb_type qp qp_dbf cbp | luma_mask ...
    const int type = h.y & 0xff, qp = (h.y >> 16) & 0xff;
    const unsigned mask = h.z & 0xffff;
    const bool intra = P264B200_IS_INTRA(type);
    int ref[4];
#pragma unroll
    for (int i = 0; i < 4; i++) ref[i] = (int)(signed char)(h.x >> (8 * i));

    // left neighbour: blocks 3, 7, 11, 15 and 8x8 quadrants 1, 3; top neighbour: blocks 12..15, quadrants 2, 3
    int lmv[4] = {0, 0, 0, 0}, tmv[4] = {0, 0, 0, 0}, lref[2] = {0, 0}, tref[2] = {0, 0}, lqp = qp, tqp = qp;
    unsigned lmask = 0, tmask = 0;
    bool lintra = false, tintra = false;
    if (mbx > 0) {
        const uint4 *nr = rec - 6;
#pragma unroll
        for (int i = 0; i < 4; i++) lmv[i] = (int)__ldg(reinterpret_cast<const uint32_t *>(nr + i) + 3);
        const uint4 nh = __ldg(nr + 4);
        lref[0] = (int)(signed char)(nh.x >> 8), lref[1] = (int)(signed char)(nh.x >> 24);
        lintra = P264B200_IS_INTRA(nh.y & 0xff);
        lqp = (nh.y >> 16) & 0xff;
        lmask = nh.z & 0xffff;
    }
    if (mby > 0) {
        const uint4 *nr = rec - 6 * g.mb_w;
        const uint4 q = __ldg(nr + 3);
        tmv[0] = (int)q.x, tmv[1] = (int)q.y, tmv[2] = (int)q.z, tmv[3] = (int)q.w;
        const uint4 nh = __ldg(nr + 4);
        tref[0] = (int)(signed char)(nh.x >> 16), tref[1] = (int)(signed char)(nh.x >> 24);
        tintra = P264B200_IS_INTRA(nh.y & 0xff);
        tqp = (nh.y >> 16) & 0xff;
        tmask = nh.z & 0xffff;
    }

    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int seg = 0; seg < 4; seg++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int bv = 0, bh = 0;
            {
                // vertical edge e, rows 4*seg..: q block (e, seg), p block to its left
                const int bq = 4 * seg + e;
                if (e > 0)
                    bv = bs_of(intra, 3, (mask >> bq) | (mask >> (bq - 1)), ref[(bq >> 3) * 2 + ((bq & 3) >> 1)],
                               ref[((bq - 1) >> 3) * 2 + (((bq - 1) & 3) >> 1)], mv[bq], mv[bq - 1]);
                else if (mbx > 0)
                    bv = bs_of(intra || lintra, 4, (mask >> bq) | (lmask >> (bq + 3)), ref[(bq >> 3) * 2], lref[seg >> 1], mv[bq], lmv[seg]);
            }
            {
                // horizontal edge e, columns 4*seg..: q block (seg, e), p block above it
                const int bq = 4 * e + seg;
                if (e > 0)
                    bh = bs_of(intra, 3, (mask >> bq) | (mask >> (bq - 4)), ref[(bq >> 3) * 2 + ((bq & 3) >> 1)],
                               ref[((bq - 4) >> 3) * 2 + (((bq - 4) & 3) >> 1)], mv[bq], mv[bq - 4]);
                else if (mby > 0)
                    bh = bs_of(intra || tintra, 4, (mask >> bq) | (tmask >> (12 + seg)), ref[seg >> 1], tref[seg >> 1], mv[bq], tmv[seg]);
            }
            w[seg] |= (uint32_t)(bv | (bh << 4)) << (8 * e);
        }
    }
    // left / top / inner edge parameters: luma QP average and mapped chroma QP average (core/frame.c:476-483,600-601)
    const int off = fd.chroma_qp_off;
    const int qc = c_chroma_qp[clip3i(qp + off, 0, 51)];
    uint32_t pl[3], pc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int qn = k == 0 ? lqp : k == 1 ? tqp : qp;
        const int qcn = c_chroma_qp[clip3i(qn + off, 0, 51)];
        pl[k] = pack_edge_params((qp + qn + 1) >> 1, fd.alpha_off, fd.beta_off);
        pc[k] = pack_edge_params((qc + qcn + 1) >> 1, fd.alpha_off, fd.beta_off);
    }
    uint4 *out = reinterpret_cast<uint4 *>(fd.dbf_bs + mb_xy);
    out[0] = make_uint4(w[0], w[1], w[2], w[3]);
    out[1] = make_uint4(pl[0], pl[1], pl[2], 0);
    out[2] = make_uint4(pc[0], pc[1], pc[2], 0);
}


constexpr int kDbfRows = 8;     // macroblock rows (filtering warps) per CTA
constexpr int kDbfQuad = 4;     // lanes (streams) per warp, 8 threads each
constexpr int kDbfRing = 4;     // hand-off slots per producer row
constexpr int kDbfTile = 272;   // bytes per (warp, stream) transpose tile: 16 rows x 16 B, + 16 B bank skew
constexpr int kDbfSlot = 80;    // bytes per (slot, stream): 4 rows x 16 B, + 16 B bank skew
constexpr int kDbfSmSlots = 256;  // per-SM arrival counters (indexed by %smid) behind the tickets
constexpr int kDbfWarps = kDbfRows + 2;  // + the "in" and "out" warps that own the global progress words

// Optional per-macroblock cycle trace of ONE CTA (engine debug knob P264B200_TRACE=<ticket>): [warp][x][marks]:
// 0 top of the iteration, 1 before waiting for the rows above, 2 after that wait, 3 end of the iteration
constexpr int kDbfTraceSteps = 320;
__device__ long long g_dbf_trace[12][kDbfTraceSteps][6];
__device__ long long g_dbf_cta_ns[2048][4];  // per ticket: %globaltimer at kernel entry, first macroblock, last macroblock, exit (row warp 0)
__device__ __forceinline__ long long dbf_now_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// (compiled in only with -DP264B200_DBF_TRACE: the four predicate tests per macroblock were 3 % of the kernel's instructions)
#ifdef P264B200_DBF_TRACE
constexpr bool kDbfTraceOn = true;
#else
constexpr bool kDbfTraceOn = false;
#endif
__device__ __forceinline__ void dbf_mark(bool on, int w, int s, int k)
{
    if (kDbfTraceOn && on && s < kDbfTraceSteps && (threadIdx.x & 31) == 0) g_dbf_trace[w][s][k] = clock64();
}

// ---- shared-memory transaction barriers (mbarrier), one arriving thread per phase -------------------
// Slot k of a ring is used for macroblocks k, k + R, k + 2R, ...; use number n = x / R.  The consumer waits
// on `full` with parity n & 1, the producer on `empty` with parity (n & 1) ^ 1 (passes at once for n = 0:
// a fresh barrier reports its "previous" phase as complete).  arrive = release.cta, try_wait = acquire.cta.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// (the suspend-time hint keeps a waiting warp asleep in hardware instead of re-issuing the probe: without it the
// probe loop was 12 % of all executed instructions)
constexpr uint32_t kDbfSuspendNs = 20000;
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DBF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DBF_DONE;\n"
        "bra DBF_WAIT;\n"
        "DBF_DONE:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity), "r"(kDbfSuspendNs)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ int lds_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release(int *p, int v)
{
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

template <int ROWS>
struct DbfSmem {
    uint8_t tile[ROWS][kDbfQuad][kDbfTile];                  // luma: row r at 16r; chroma: plane p row r at 64p + 8r
    uint8_t ring[ROWS + 1][kDbfRing][kDbfQuad][kDbfSlot];    // ring[0]: rows above warp 0 (filled by the in warp from global memory);
                                                                 // ring[1 + w]: last rows of warp w's macroblocks.  luma: rows 12..15 at 16k;
                                                                 // chroma: plane p rows 6,7 at 16p + 8k
    uint64_t full[ROWS + 1][kDbfRing];                       // ring[i][k] holds the rows of its next macroblock
    uint64_t empty[ROWS + 1][kDbfRing];                      // the consumer is done with ring[i][k]
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

// The four luma edges of one direction on two lines, v = 4 samples before the macroblock + its 16 samples, split
// into the macroblock edge (after which the neighbouring macroblock is final and can be handed on) and the three
// inner edges.  bs4: bS of edge e in bits 8e..8e+3.  The bS < 4 filters are straight-line code with no votes or
// branches between the edges, so the instruction scheduler can overlap them (edge e+1 needs edge e only through one tap).
__device__ __forceinline__ void dbf_luma_edge0(uint32_t *v, uint32_t bs4, const DbfDir &d)
{
    const int bs0 = bs4 & 0xf;
    if (!__any_sync(0xffffffffu, bs0 != 0)) return;
    if (__any_sync(0xffffffffu, bs0 == 4)) {
        // intra macroblock edge: the strong filter, only on the lines that have bS 4 (the normal one is off there)
        const int alpha = d.prm_mb & 0xff;
        swar::EdgeK k4;
        k4.n_alpha = bs0 == 4 ? d.na_mb : 0u, k4.n_beta = d.nb_mb, k4.tc0 = 0;
        swar::luma_strong(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], k4, alpha);
    }
    const swar::EdgeK k = dbf_edge_k(d, true, bs0, true);
    swar::luma_normal(v[1], v[2], v[3], v[4], v[5], v[6], k);
}
__device__ __forceinline__ void dbf_luma_inner(uint32_t *v, uint32_t bs4, const DbfDir &d)
{
    if (!__any_sync(0xffffffffu, (bs4 & 0x0f0f0f00u) != 0)) return;
#pragma unroll
    for (int e = 1; e < 4; e++) {
        const swar::EdgeK k = dbf_edge_k(d, false, (bs4 >> (8 * e)) & 0xf, true);
        swar::luma_normal(v[4 * e + 1], v[4 * e + 2], v[4 * e + 3], v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], k);
    }
}
// The two chroma edges of one direction on two lines, v = p1 p0 | q0 q1 . . q0' q1' ...: edge 0 at v[2], edge 2 at v[6]
__device__ __forceinline__ void dbf_chroma_edge0(uint32_t *v, int bs0, const DbfDir &d)
{
    if (!__any_sync(0xffffffffu, bs0 != 0)) return;
    const swar::EdgeK k0 = dbf_edge_k(d, true, bs0, false);
    uint32_t p0 = v[1], q0 = v[2];
    swar::chroma_edge2(v[0], p0, q0, v[3], k0, false);
    if (__any_sync(0xffffffffu, bs0 == 4)) {
        uint32_t sp0 = v[1], sq0 = v[2];
        swar::chroma_edge2(v[0], sp0, sq0, v[3], k0, true);
        if (bs0 == 4) p0 = sp0, q0 = sq0;
    }
    v[1] = p0, v[2] = q0;
}
__device__ __forceinline__ void dbf_chroma_inner(uint32_t *v, int bs2, const DbfDir &d)
{
    if (!__any_sync(0xffffffffu, bs2 != 0)) return;
    const swar::EdgeK k2 = dbf_edge_k(d, false, bs2, false);
    swar::chroma_edge2(v[4], v[5], v[6], v[7], k2, false);
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
template <bool C, int ROWS>
__device__ __forceinline__ void deblock_rows(DbfSmem<ROWS> &sm, const FrameDesc *__restrict__ descs, const Geometry &g, int n_lanes, int quad,
                                             int grp, bool trace, bool times_on, int tk)
{
    constexpr int NW = C ? 2 : 4;    // 32-bit words per sample row of a macroblock
    constexpr int RB = 4 * NW;       // bytes per row
    constexpr int NR = C ? 8 : 16;   // rows per plane
    constexpr int TR = C ? 2 : 4;    // rows handed down to the macroblock row below (per plane)
    constexpr int NP = 4 + 4 * NW;   // taps along a row incl. the 4 samples left of the macroblock
    constexpr int NQ = TR + NR;      // taps down a column incl. the rows above
    constexpr int RM = kDbfRing - 1;
    typedef typename DbfVec<NW>::type Vec;
    using swar::prmt;

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, t = lane & 7;
    const int pl = C ? t >> 2 : 0, j = C ? t & 3 : t;  // plane; row pair (vertical edges) = column pair (horizontal edges)
    const int row = grp * ROWS + w;
    if (row >= g.mb_h) return;       // nobody waits for a row outside the picture
    const bool has_top = row > 0;
    const bool bottom_smem = w + 1 < ROWS && row + 1 < g.mb_h;     // the row below is a warp of this CTA
    const bool publishes = w + 1 == ROWS && row + 1 < g.mb_h;      // ... or the first row of the next CTA
    const int stream = kDbfQuad * quad + sub;
    const FrameDesc &fd = descs[min(stream, n_lanes - 1)];
    const bool act = stream < n_lanes && fd.deblock != 0;
    const int stride = C ? g.c_stride : g.y_stride;
    // this thread's two sample rows, and (threads 0..3) the row above it moves from shared to global memory:
    // luma row -4 + t; chroma plane t >> 1, row -2 + (t & 1)
    uint8_t *grow = fd.cur[C ? 1 + pl : 0] + (ptrdiff_t)(NR * row + 2 * j) * stride;
    uint8_t *gtop = C ? fd.cur[1 + ((t >> 1) & 1)] + (ptrdiff_t)(NR * row - 2 + (t & 1)) * stride : fd.cur[0] + (ptrdiff_t)(NR * row - 4 + (t & 3)) * stride;
    const int top_off = C ? 16 * ((t >> 1) & 1) + 8 * (t & 1) : 16 * (t & 3);  // of that row inside a slot
    const DeblockSide *side = fd.dbf_bs + (size_t)row * g.mb_w;
    uint8_t *T = sm.tile[w][sub] + (C ? 64 * pl : 0);
    // rows this thread stores itself / hands to the row below through the ring
    const bool store_a = !(bottom_smem && 2 * j > NR - TR), store_b = !(bottom_smem && 2 * j + 1 > NR - TR);   // (a publishing row stores everything itself)
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

    // Macroblock m of this row is final except for what the row below does to its last rows: store it, and
    // pass those rows down (shared-memory ring inside the CTA, global memory + the out warp across CTAs).
    auto hand_off = [&](int m) {
        Vec va, vb;
        vec_set(va, prev[0]);
        vec_set(vb, prev[1]);
        if (act && store_a) *reinterpret_cast<Vec *>(grow + RB * m) = va;
        if (act && store_b) *reinterpret_cast<Vec *>(grow + stride + RB * m) = vb;
        if (bottom_smem || publishes) {
            // (the CTA's last row hands over to the out thread the same way, without data: ring index ROWS)
            mbar_wait(&sm.empty[w + 1][m & RM], ((m / kDbfRing) & 1) ^ 1);
            if (to_ring) {
                uint8_t *slot = sm.ring[w + 1][m & RM][sub] + ring_off;
                *reinterpret_cast<Vec *>(slot) = va;
                *reinterpret_cast<Vec *>(slot + RB) = vb;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.full[w + 1][m & RM]);
        }
    };

    // Each warp runs at its own pace; the only cross-warp dependency is the top edge of macroblock x, which needs
    // the last rows of macroblock x of the row above AFTER the left edge of its macroblock x+1 (a one-macroblock lag).
#pragma unroll 1
    for (int x = 0; x < g.mb_w; x++) {
        const bool last = x == g.mb_w - 1;
        dbf_mark(trace, w, x, 0);
        if (kDbfTraceOn && times_on && threadIdx.x == 0 && (x == 0 || last)) g_dbf_cta_ns[tk][x == 0 ? 1 : 2] = dbf_now_ns();
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
        // ---- vertical edges: taps of this thread's two rows, two rows per register.  The macroblock edge goes first:
        // it finalises the previous macroblock (its last columns), whose hand-off is what the row below waits for.
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
                dbf_chroma_edge0(P + 2, bsw & 0xf, dv);
            else
                dbf_luma_edge0(P, bsw & 0x0f0f0f0fu, dv);
            {
                const uint32_t t01 = prmt(P[0], P[1], 0x6240), t23 = prmt(P[2], P[3], 0x6240);
                prev[0][NW - 1] = prmt(t01, t23, 0x5410), prev[1][NW - 1] = prmt(t01, t23, 0x7632);
            }
            if (x > 0) hand_off(x - 1);
            if (C)
                dbf_chroma_inner(P + 2, (bsw >> 16) & 0xf, dv);
            else
                dbf_luma_inner(P, bsw & 0x0f0f0f0fu, dv);
#pragma unroll
            for (int wd = 1; wd < 1 + NW; wd++) {
                const uint32_t t01 = prmt(P[4 * wd + 0], P[4 * wd + 1], 0x6240), t23 = prmt(P[4 * wd + 2], P[4 * wd + 3], 0x6240);
                cur[0][wd - 1] = prmt(t01, t23, 0x5410), cur[1][wd - 1] = prmt(t01, t23, 0x7632);
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
        __syncwarp();
        dbf_mark(trace, w, x, 1);
        if (has_top) mbar_wait(&sm.full[w][x & RM], (x / kDbfRing) & 1);
        dbf_mark(trace, w, x, 2);
        uint8_t *slot_top = sm.ring[w][x & RM][sub];
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
                dbf_chroma_edge0(Q, (bsw >> 4) & 0xf, dh);
                dbf_chroma_inner(Q, (bsw >> 20) & 0xf, dh);
                if (has_top) *reinterpret_cast<uint16_t *>(topp + RB * 1 + 2 * j) = (uint16_t)prmt(Q[1], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 0 + 2 * j) = (uint16_t)prmt(Q[2], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 3 + 2 * j) = (uint16_t)prmt(Q[5], 0, 0x4420);
                *reinterpret_cast<uint16_t *>(T + RB * 4 + 2 * j) = (uint16_t)prmt(Q[6], 0, 0x4420);
            } else {
                dbf_luma_edge0(Q, (bsw >> 4) & 0x0f0f0f0fu, dh);
                dbf_luma_inner(Q, (bsw >> 4) & 0x0f0f0f0fu, dh);
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
        if (has_top) {
            if (act && t < 4 && (C ? (t & 1) : (t != 0))) *reinterpret_cast<Vec *>(gtop + RB * x) = *reinterpret_cast<const Vec *>(slot_top + top_off);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[w][x & RM]);
        }
        if (last) hand_off(x);
        dbf_mark(trace, w, x, 3);
        if (kDbfTraceOn && trace) {
            // diagnostics only: split the wait for the prefetched side info from the wait for the prefetched rows
            uint32_t d;
            asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(bsw_n), "r"(prm_n.x ^ prm_n.z));
            if (d == 0x12345678u) g_dbf_trace[0][0][5] = d;
            dbf_mark(trace, w, x, 4);
            asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(nxt[0][0]), "r"(nxt[1][NW - 1]));
            if (d == 0x12345678u) g_dbf_trace[0][0][5] = d;
            dbf_mark(trace, w, x, 5);
        }
#pragma unroll
        for (int k = 0; k < NW; k++) cur[0][k] = nxt[0][k], cur[1][k] = nxt[1][k];
        bsw = bsw_n;
        prm = prm_n;
    }
}

// The "in" warp of a CTA (row group > 0): keeps ring[0] -- the rows above warp 0, which the row group above
// stored to global memory -- up to kDbfRing macroblocks ahead of warp 0.  It alone polls the global progress
// word of the row above, so no filtering warp ever spins on global memory.
template <bool C, int ROWS>
__device__ __forceinline__ void deblock_in_warp(DbfSmem<ROWS> &sm, const FrameDesc *__restrict__ descs, const Geometry &g, int n_lanes, int quad,
                                                int grp)
{
    constexpr int RB = C ? 8 : 16, NR = C ? 8 : 16, RM = kDbfRing - 1;
    typedef typename DbfVec<RB / 4>::type Vec;
    const int lane = threadIdx.x & 31, sub = (lane >> 2) & 3, k = lane & 3;  // lanes 0..15: (stream, row above)
    const int row0 = grp * ROWS;                 // first macroblock row of the CTA
    const int stream = kDbfQuad * quad + sub;
    const FrameDesc &fd = descs[min(stream, n_lanes - 1)];
    const bool act = lane < 16 && stream < n_lanes && fd.deblock != 0;
    const int *prog = descs[kDbfQuad * quad].row_progress + (C ? 2 : 1) * g.mb_h + row0 - 1;  // one progress word per (quad, role, row)
    const int stride = C ? g.c_stride : g.y_stride;
    // luma: row -4 + k; chroma: plane k >> 1, row -2 + (k & 1)
    const uint8_t *gtop = C ? fd.cur[1 + (k >> 1)] + (ptrdiff_t)(NR * row0 - 2 + (k & 1)) * stride : fd.cur[0] + (ptrdiff_t)(NR * row0 - 4 + k) * stride;
    const int top_off = C ? 16 * (k >> 1) + 8 * (k & 1) : 16 * k;
    int seen = 0;
#pragma unroll 1
    for (int x = 0; x < g.mb_w; x++) {
        if (seen <= x) {
            if (lane == 0) {
                // relaxed polls (an acquire load invalidates the SM's whole L1 on every spin), one fence on success
                int spins = 0;
                while ((seen = ld_relaxed(prog)) <= x) __nanosleep(++spins < 8 ? 100 : 400);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncwarp();   // orders the other lanes' loads after lane 0's fence
            seen = __shfl_sync(0xffffffffu, seen, 0);
        }
        Vec v;
        {
            uint32_t z[4] = {0, 0, 0, 0};
            vec_set(v, z);
        }
        if (act) v = __ldcg(reinterpret_cast<const Vec *>(gtop + RB * x));
        mbar_wait(&sm.empty[0][x & RM], ((x / kDbfRing) & 1) ^ 1);
        if (act) *reinterpret_cast<Vec *>(sm.ring[0][x & RM][sub] + top_off) = v;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[0][x & RM]);
    }
}

// The "out" thread of a CTA that has a row group below it: publishes how many macroblocks of the CTA's last row
// are in global memory.  It is the consumer of that row's (data-less) hand-off ring: the row's lanes store,
// __syncwarp, lane 0 arrives (release.cta); this thread's wait acquires, and its release.gpu store of the progress
// word is cumulative over those stores.  Whatever has arrived meanwhile is published in one go.
template <int ROWS>
__device__ __forceinline__ void deblock_out_thread(DbfSmem<ROWS> &sm, const FrameDesc *__restrict__ descs, const Geometry &g, int quad, int grp, bool chroma)
{
    constexpr int RM = kDbfRing - 1;
    const int row_last = grp * ROWS + ROWS - 1;
    if (row_last + 1 >= g.mb_h) return;
    int *prog = descs[kDbfQuad * quad].row_progress + (chroma ? 2 : 1) * g.mb_h + row_last;
    int m = 0;
    while (m < g.mb_w) {
        mbar_wait(&sm.full[ROWS][m & RM], (m / kDbfRing) & 1);
        int hi = m + 1;
        while (hi < g.mb_w && hi < m + kDbfRing && mbar_test(&sm.full[ROWS][hi & RM], (hi / kDbfRing) & 1)) hi++;
        for (int k = m; k < hi; k++) mbar_arrive(&sm.empty[ROWS][k & RM]);
        st_release(prog, hi);   // release.gpu is cumulative over what the waits above made visible
        m = hi;
    }
}

// grid: 2 roles x ceil(n_lanes/4) stream quads x ceil(mb_h/ROWS) row groups.  A CTA draws its work at run time:
//  * the role from its arrival order on its SM (first luma, second chroma, ...), so that co-resident CTAs are
//    one luma + one chroma -- luma is the heavier role, and two luma CTAs on one SM would pace every chain below them;
//  * (row group, quad) from a per-role ticket, row-group-major: the group above of the same quad and role always
//    has a smaller ticket, i.e. is resident or finished.  A role whose tickets are used up falls back to the other.
// sync: [0] intra ticket (other kernel), [1] luma ticket, [2] chroma ticket, [4 + smid] arrivals per SM
// ROWS = macroblock rows (filtering warps) per CTA, MINB = CTAs per SM the register budget is set for (engine knob P264B200_DBF_VARIANT)
template <int ROWS, int MINB>
__global__ void __launch_bounds__(32 * (ROWS + 2), MINB) deblock_kernel(const FrameDesc *__restrict__ descs, Geometry g, int n_lanes, int *sync, int trace_ticket)
{
    __shared__ __align__(16) DbfSmem<ROWS> sm;
    const int groups = (g.mb_h + ROWS - 1) / ROWS;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        const int per_role = (int)(gridDim.x >> 1);
        int role = atomicAdd(sync + 4 + (smid & (kDbfSmSlots - 1)), 1) & 1;
        int u = atomicAdd(sync + 1 + role, 1);
        if (u >= per_role) {
            role ^= 1;
            u = atomicAdd(sync + 1 + role, 1);
        }
        sm.ticket = 2 * u + role;
        for (int i = 0; i < (ROWS + 1) * kDbfRing; i++) {
            mbar_init(&sm.full[0][0] + i, 1);
            mbar_init(&sm.empty[0][0] + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tk = sm.ticket;
    const int role = tk & 1, u = tk >> 1;
    // row-group-major over the stream quads: when there are more CTAs than fit on the machine, the resident ones are
    // the upper row groups of EVERY quad (all busy) rather than whole chains of a few quads (lower groups idle)
    const int quads = (int)(gridDim.x >> 1) / groups;
    const int grp = u / quads, quad = u - grp * quads;
    const bool trace = tk == trace_ticket;
    const bool times = trace_ticket >= 0 && tk < 2048 && threadIdx.x == 0;
    if (kDbfTraceOn && times) g_dbf_cta_ns[tk][0] = dbf_now_ns();
    const int w = threadIdx.x >> 5;
    if (w == ROWS) {
        if (grp > 0) {
            if (role == 0)
                deblock_in_warp<false, ROWS>(sm, descs, g, n_lanes, quad, grp);
            else
                deblock_in_warp<true, ROWS>(sm, descs, g, n_lanes, quad, grp);
        }
        return;
    }
    if (w == ROWS + 1) {
        if ((threadIdx.x & 31) == 0) deblock_out_thread<ROWS>(sm, descs, g, quad, grp, role != 0);
        return;
    }
    if (role == 0)
        deblock_rows<false, ROWS>(sm, descs, g, n_lanes, quad, grp, trace, trace_ticket >= 0 && tk < 2048, tk);
    else
        deblock_rows<true, ROWS>(sm, descs, g, n_lanes, quad, grp, trace, trace_ticket >= 0 && tk < 2048, tk);
    if (kDbfTraceOn && times) g_dbf_cta_ns[tk][3] = dbf_now_ns();
}
#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
