// Intra macroblock reconstruction as a macroblock-row wavefront.
//
// Replaces the intra branches of p264_macroblock_decode (decoder/macroblock.c:771-831,853-890),
// valid_intra16x16/4x4/8x8c_mode (decoder/macroblock.c:635-753), the predictors of
// core/predict.c:55-638 and the neighbour flags of core/macroblock.c:1210-1231.
//
// Work item = a RUN of horizontally adjacent intra macroblocks, walked left to right by one warp (the left
// neighbour is then the warp's own previous macroblock).  Inter macroblocks were written by recon_inter before
// this kernel starts, so a macroblock only has to wait for those of its top-left / top / top-right neighbours
// that are intra themselves: per-macroblock "done" words (value = the launch's epoch, so nothing is ever
// cleared), polled with relaxed loads + one fence.  An I picture degenerates to the classic row wavefront
// (run = row, every macroblock waits for the row above to be two macroblocks ahead); the few intra
// macroblocks of a P picture have almost no intra neighbours and reconstruct in parallel instead of
// queueing behind a row-serial progress counter (5 % intra macroblocks: 5.9 ms -> see profiles/README.md).
// Runs are listed in raster order by intra_runs_kernel and handed out by an atomic ticket, run-index-major
// over the lanes: every dependency of a run has a smaller ticket, which makes the spin-wait deadlock-free
// whatever order the hardware schedules CTAs in.
#pragma once
#include "common.cuh"

namespace p264b200 {

constexpr int kIntraWarpsPerCta = 1;
constexpr int kTS = 24;  // luma tile row stride: rows -1..15, cols -1..19
constexpr int kCS = 12;  // chroma tile row stride: rows -1..7, cols -1..7

struct __align__(16) IntraSmem {
    p264b200_mb mb;     // the macroblock record (every field is read many times, across warp barriers)
    uint8_t y[17 * kTS];
    uint8_t c[2][9 * kCS];
    short res_y[16][16];
    short res_c[8][16];
    uint8_t edge[16];  // l3 l2 l1 l0 lt t0..t7
};

__device__ __forceinline__ uint8_t &TY(IntraSmem &s, int r, int c) { return s.y[(r + 1) * kTS + (c + 1)]; }
__device__ __forceinline__ uint8_t &TC(IntraSmem &s, int p, int r, int c) { return s.c[p][(r + 1) * kCS + (c + 1)]; }

// One sample of an Intra4x4 prediction (H.264 8.3.1.2 == core/predict.c:366-638).
// e[] = {l3,l2,l1,l0,lt,t0,...,t7}: left p[-1,k] = e[3-k], top p[k,-1] = e[5+k].
__device__ __forceinline__ int pred4x4_sample(int mode, int x, int y, const uint8_t *e)
{
#define TOPP(k) ((int)e[5 + (k)])
#define LEFTP(k) ((int)e[3 - (k)])
    switch (mode) {
    case 0: return TOPP(x);
    case 1: return LEFTP(y);
    case 2: return (LEFTP(0) + LEFTP(1) + LEFTP(2) + LEFTP(3) + TOPP(0) + TOPP(1) + TOPP(2) + TOPP(3) + 4) >> 3;
    case 9: return (LEFTP(0) + LEFTP(1) + LEFTP(2) + LEFTP(3) + 2) >> 2;
    case 10: return (TOPP(0) + TOPP(1) + TOPP(2) + TOPP(3) + 2) >> 2;
    case 11: return 128;
    case 3:  // diagonal down-left
        if (x == 3 && y == 3) return (TOPP(6) + 3 * TOPP(7) + 2) >> 2;
        return (TOPP(x + y) + 2 * TOPP(x + y + 1) + TOPP(x + y + 2) + 2) >> 2;
    case 4: {  // diagonal down-right
        const int i = 4 + x - y;
        return (e[i - 1] + 2 * e[i] + e[i + 1] + 2) >> 2;
    }
    case 5: {  // vertical-right
        const int z = 2 * x - y, k = x - (y >> 1);
        if (z >= 0) return (z & 1) ? (TOPP(k - 2) + 2 * TOPP(k - 1) + TOPP(k) + 2) >> 2 : (TOPP(k - 1) + TOPP(k) + 1) >> 1;
        if (z == -1) return (LEFTP(0) + 2 * LEFTP(-1) + TOPP(0) + 2) >> 2;
        return (LEFTP(y - 1) + 2 * LEFTP(y - 2) + LEFTP(y - 3) + 2) >> 2;
    }
    case 6: {  // horizontal-down
        const int z = 2 * y - x, k = y - (x >> 1);
        if (z >= 0) return (z & 1) ? (LEFTP(k - 2) + 2 * LEFTP(k - 1) + LEFTP(k) + 2) >> 2 : (LEFTP(k - 1) + LEFTP(k) + 1) >> 1;
        if (z == -1) return (LEFTP(0) + 2 * LEFTP(-1) + TOPP(0) + 2) >> 2;
        return (TOPP(x - 1) + 2 * TOPP(x - 2) + TOPP(x - 3) + 2) >> 2;
    }
    case 7: {  // vertical-left
        const int k = x + (y >> 1);
        return (y & 1) ? (TOPP(k) + 2 * TOPP(k + 1) + TOPP(k + 2) + 2) >> 2 : (TOPP(k) + TOPP(k + 1) + 1) >> 1;
    }
    default: {  // 8: horizontal-up
        const int z = x + 2 * y, k = y + (x >> 1);
        if (z > 5) return LEFTP(3);
        if (z == 5) return (LEFTP(2) + 3 * LEFTP(3) + 2) >> 2;
        return (z & 1) ? (LEFTP(k) + 2 * LEFTP(k + 1) + LEFTP(k + 2) + 2) >> 2 : (LEFTP(k) + LEFTP(k + 1) + 1) >> 1;
    }
    }
#undef TOPP
#undef LEFTP
}

// plane predictors (core/predict.c:159-193 luma n=16, :329-361 chroma n=8); warp-uniform setup
template <int N>
__device__ __forceinline__ void plane_params(const uint8_t *tile, int ts, int &i00, int &b, int &c)
{
    // tile points at sample (0,0); top row at -ts, left column at -1, corner at -ts-1
    int H = 0, V = 0;
    constexpr int h = N / 2;
#pragma unroll
    for (int i = 0; i < h; i++) {
        H += (i + 1) * ((int)tile[-ts + h + i] - (int)tile[-ts + h - 2 - i]);
        V += (i + 1) * ((int)tile[(h + i) * ts - 1] - (int)tile[(h - 2 - i) * ts - 1]);
    }
    const int a = 16 * ((int)tile[(N - 1) * ts - 1] + (int)tile[-ts + N - 1]);
    if (N == 16) {
        b = (5 * H + 32) >> 6;
        c = (5 * V + 32) >> 6;
        i00 = a - 7 * b - 7 * c + 16;
    } else {
        b = (17 * H + 16) >> 5;
        c = (17 * V + 16) >> 5;
        i00 = a - 3 * b - 3 * c + 16;
    }
}

// what recon_intra_mb needs of the FrameDesc, held in registers (the descriptor lives in global memory and would
// be re-read after every warp barrier)
struct IntraCtx {
    uint8_t *cur[3];
    const int16_t *coefs;
    int chroma_qp_off;
};

static __device__ void recon_intra_mb(IntraSmem &s, const IntraCtx &fd, const Geometry &g, const p264b200_mb &m, int mbx,
                               int mby, int lane)
{
    const bool has_left = mbx > 0, has_top = mby > 0;
    const bool has_tr = mby > 0 && mbx < g.mb_w - 1, has_tl = mbx > 0 && mby > 0;
    uint8_t *gy = fd.cur[0] + (ptrdiff_t)16 * mby * g.y_stride + 16 * mbx;
    uint8_t *gc[2] = {fd.cur[1] + (ptrdiff_t)8 * mby * g.c_stride + 8 * mbx,
                      fd.cur[2] + (ptrdiff_t)8 * mby * g.c_stride + 8 * mbx};
    const int qp = m.qp;
    const int qpc = c_chroma_qp[clip3i(qp + fd.chroma_qp_off, 0, 51)];
    const bool i16 = m.mb_type == P264B200_MB_I16x16;

    // ---- neighbours into the tiles (L2 loads: written by other SMs in this or the previous kernel); all four
    // loads are issued before the first store so that their latencies overlap
    {
        const int c0 = lane - 1;
        const bool ok0 = lane < 21 && has_top && (c0 >= 0 || has_tl) && (c0 < 16 || has_tr);
        const bool ok1 = lane < 16 && has_left;
        const int p2 = lane / 9, c2 = lane % 9 - 1;
        const bool ok2 = lane < 18 && has_top && (c2 >= 0 || has_tl);
        const int p3 = (lane >> 3) & 1, r3 = lane & 7;
        const bool ok3 = lane < 16 && has_left;
        uint8_t v0 = 128, v1 = 128, v2 = 128, v3 = 128;
        if (ok0) v0 = __ldcg(gy - g.y_stride + c0);
        if (ok1) v1 = __ldcg(gy + lane * g.y_stride - 1);
        if (ok2) v2 = __ldcg(gc[p2] - g.c_stride + c2);
        if (ok3) v3 = __ldcg(gc[p3] + r3 * g.c_stride - 1);
        if (lane < 21) TY(s, -1, c0) = v0;
        if (lane < 16) TY(s, lane, -1) = v1;
        if (lane < 18) TC(s, p2, -1, c2) = v2;
        if (lane < 16) TC(s, p3, r3, -1) = v3;
    }

    // ---- residual samples of all 24 blocks, independent of the prediction
    const int16_t *cf = fd.coefs + m.coef_off;
    const int n_luma = __popc(m.luma_mask);
    if (lane < 16) {
        const int b = lane;
        int d[16], r[16];
        const bool coded = m.luma_mask >> b & 1;
        if (coded)
            unscan_dequant(cf + (i16 ? 16 : 0) + 16 * __popc(m.luma_mask & ((1u << b) - 1)), qp, d);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) d[i] = 0;
        }
        if (i16) {
            int dc[16];
            luma_dc(cf, qp, dc);
            int v = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) v = (i == b) ? dc[i] : v;
            d[0] = v;
        }
        if (coded || i16) {
            idct4x4_core(d, r);
#pragma unroll
            for (int i = 0; i < 16; i++) s.res_y[b][i] = (short)r[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) s.res_y[b][i] = 0;
        }
    } else if (lane < 24) {
        const int cb = lane - 16;
        int d[16], r[16];
        if (m.cbp_chroma) {
            const int16_t *cc = cf + (i16 ? 16 : 0) + 16 * n_luma;
            int dc[4];
            chroma_dc(cc + 4 * (cb >> 2), qpc, dc);
            if (m.chroma_mask >> cb & 1)
                unscan_dequant(cc + 8 + 16 * __popc(m.chroma_mask & ((1u << cb) - 1)), qpc, d);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) d[i] = 0;
            }
            int v = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) v = (i == (cb & 3)) ? dc[i] : v;
            d[0] = v;
            idct4x4_core(d, r);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) s.res_c[cb][i] = (short)r[i];
    }
    __syncwarp();

    // ---- luma prediction + residual
    if (i16) {
        int mode = m.i16_mode;  // valid_intra16x16_mode (decoder/macroblock.c:635-667)
        if (mode == 2) mode = has_tl ? 2 : has_left ? 4 : has_top ? 5 : 6;
        int dc = 128, i00 = 0, pb = 0, pc = 0;
        if (mode == 2 || mode == 4 || mode == 5) {
            int st = 0, sl = 0;
            for (int i = 0; i < 16; i++) {
                st += TY(s, -1, i);
                sl += TY(s, i, -1);
            }
            dc = mode == 2 ? (st + sl + 16) >> 5 : mode == 4 ? (sl + 8) >> 4 : (st + 8) >> 4;
        } else if (mode == 3)
            plane_params<16>(&TY(s, 0, 0), kTS, i00, pb, pc);
        // lane -> row (lane>>1), 8 samples from column 8*(lane&1)
        const int r = lane >> 1, c0 = 8 * (lane & 1);
        int pred[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = c0 + k;
            pred[k] = mode == 0 ? TY(s, -1, c) : mode == 1 ? TY(s, r, -1) : mode == 3 ? clip8i((i00 + pb * c + pc * r) >> 5) : dc;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = c0 + k;
            TY(s, r, c) = (uint8_t)clip8i(pred[k] + s.res_y[(r >> 2) * 4 + (c >> 2)][(r & 3) * 4 + (c & 3)]);
        }
    } else {
        const uint8_t zx[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
        const uint8_t zy[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
        for (int i = 0; i < 16; i++) {
            const int bx = (i & 1) | ((i >> 1) & 2), by = ((i >> 1) & 1) | ((i >> 2) & 2);
            (void)zx;
            (void)zy;
            const int b = by * 4 + bx;
            // availability at the time the block is coded (core/macroblock.c:1210-1231)
            const bool a_left = bx > 0 || has_left, a_top = by > 0 || has_top;
            const bool a_tl = (bx > 0 && by > 0) ? true : (bx > 0) ? has_top : (by > 0) ? has_left : has_tl;
            const bool a_tr = by == 0 ? (bx < 3 ? has_top : has_tr) : !(bx == 3 || (bx == 1 && (by & 1)));
            if (lane < 13) {
                int v;
                if (lane < 4)
                    v = a_left ? TY(s, 4 * by + 3 - lane, 4 * bx - 1) : 128;
                else if (lane == 4)
                    v = a_tl ? TY(s, 4 * by - 1, 4 * bx - 1) : 128;
                else if (lane < 9)
                    v = a_top ? TY(s, 4 * by - 1, 4 * bx + lane - 5) : 128;
                else
                    v = a_tr ? TY(s, 4 * by - 1, 4 * bx + lane - 5) : (a_top ? TY(s, 4 * by - 1, 4 * bx + 3) : 128);
                s.edge[lane] = (uint8_t)v;
            }
            __syncwarp();
            int mode = (m.i4_mode[b >> 1] >> ((b & 1) * 4)) & 15;  // valid_intra4x4_mode (decoder/macroblock.c:669-719)
            if (mode == 2) mode = (a_left && a_top) ? 2 : a_left ? 9 : a_top ? 10 : 11;
            if (lane < 16) {
                const int x = lane & 3, y = lane >> 2;
                const int p = pred4x4_sample(mode, x, y, s.edge);
                TY(s, 4 * by + y, 4 * bx + x) = (uint8_t)clip8i(p + s.res_y[b][lane]);
            }
            __syncwarp();
        }
    }

    // ---- chroma prediction + residual (decoder/macroblock.c:853-890)
    {
        int mode = m.chroma_mode;  // valid_intra8x8c_mode (decoder/macroblock.c:721-753)
        if (mode == 0) mode = has_tl ? 0 : has_left ? 4 : has_top ? 5 : 6;
        const int p = lane >> 4, r = (lane >> 1) & 7, c0 = 4 * (lane & 1);  // 4 samples per lane
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0, i00 = 0, pb = 0, pc = 0;
        if (mode == 0 || mode == 4 || mode == 5) {
            for (int i = 0; i < 4; i++) {
                s0 += TC(s, p, -1, i);
                s1 += TC(s, p, -1, 4 + i);
                s2 += TC(s, p, i, -1);
                s3 += TC(s, p, 4 + i, -1);
            }
        } else if (mode == 3)
            plane_params<8>(&TC(s, p, 0, 0), kCS, i00, pb, pc);
        int dcq;  // DC of this lane's quadrant (core/predict.c:212-297)
        {
            const int q = (r >> 2) * 2 + (c0 >> 2);
            if (mode == 0)
                dcq = q == 0 ? (s0 + s2 + 4) >> 3 : q == 1 ? (s1 + 2) >> 2 : q == 2 ? (s3 + 2) >> 2 : (s1 + s3 + 4) >> 3;
            else if (mode == 4)
                dcq = (r >> 2) ? (s3 + 2) >> 2 : (s2 + 2) >> 2;
            else if (mode == 5)
                dcq = (c0 >> 2) ? (s1 + 2) >> 2 : (s0 + 2) >> 2;
            else
                dcq = 128;
        }
        int pred[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + k;
            pred[k] = mode == 1 ? TC(s, p, r, -1) : mode == 2 ? TC(s, p, -1, c) : mode == 3 ? clip8i((i00 + pb * c + pc * r) >> 5) : dcq;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + k;
            TC(s, p, r, c) = (uint8_t)clip8i(pred[k] + s.res_c[p * 4 + (r >> 2) * 2 + (c >> 2)][(r & 3) * 4 + (c & 3)]);
        }
    }
    __syncwarp();

    // ---- write the macroblock back (L2, other SMs read it as a neighbour)
    for (int w = lane; w < 64; w += 32) {
        const int r = w >> 2, c = 4 * (w & 3);
        const uint32_t v = TY(s, r, c) | (TY(s, r, c + 1) << 8) | (TY(s, r, c + 2) << 16) | ((uint32_t)TY(s, r, c + 3) << 24);
        __stcg(reinterpret_cast<uint32_t *>(gy + r * g.y_stride + c), v);
    }
    {
        const int p = lane >> 4, r = (lane >> 1) & 7, c = 4 * (lane & 1);
        const uint32_t v = TC(s, p, r, c) | (TC(s, p, r, c + 1) << 8) | (TC(s, p, r, c + 2) << 16) | ((uint32_t)TC(s, p, r, c + 3) << 24);
        __stcg(reinterpret_cast<uint32_t *>(gc[p] + r * g.c_stride + c), v);
    }
}

#ifdef P264B200_DEFINE_KERNELS
// Per lane: work[0] = number of runs, work[1 + i] = first macroblock of run i (raster order),
// work[1 + n_mb + mb] = epoch of the launch that last reconstructed intra macroblock mb.
constexpr int kRunThreads = 1024;
__global__ void __launch_bounds__(kRunThreads) intra_runs_kernel(const FrameDesc *__restrict__ descs, Geometry g)
{
    __shared__ int warp_sum[kRunThreads / 32];
    const FrameDesc &fd = descs[blockIdx.x];
    int *work = fd.intra_work;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (fd.n_intra == 0) {
        if (tid == 0) work[0] = 0;
        return;
    }
    const int n_mb = g.mb_w * g.mb_h;
    const int per = (n_mb + kRunThreads - 1) / kRunThreads, m0 = tid * per, m1 = min(n_mb, m0 + per);
    auto starts_run = [&](int mb) {
        if (!P264B200_IS_INTRA(__ldg(&fd.mbs[mb].mb_type))) return false;
        return mb % g.mb_w == 0 || !P264B200_IS_INTRA(__ldg(&fd.mbs[mb - 1].mb_type));
    };
    int cnt = 0;
    for (int mb = m0; mb < m1; mb++) cnt += starts_run(mb);
    // exclusive scan of cnt over the block
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += u;
        }
        warp_sum[lane] = v;
    }
    __syncthreads();
    int pos = incl - cnt + (wid ? warp_sum[wid - 1] : 0);
    for (int mb = m0; mb < m1; mb++)
        if (starts_run(mb)) work[1 + pos++] = mb;
    if (tid == kRunThreads - 1) work[0] = pos;
}

__global__ void __launch_bounds__(32, 32) recon_intra_kernel(const FrameDesc *__restrict__ descs, Geometry g, int n_lanes, int *ticket, int epoch)
{
    __shared__ IntraSmem s;
    __shared__ int s_ticket;
    const int lane = threadIdx.x;
    if (lane == 0) s_ticket = atomicAdd(ticket, 1);
    __syncwarp();
    const int t = s_ticket;
    const int run = t / n_lanes;
    const FrameDesc &fd = descs[t - run * n_lanes];
    const int n_mb = g.mb_w * g.mb_h;
    const int *work = fd.intra_work;
    if (run >= work[0]) return;
    int *done = fd.intra_work + 1 + n_mb;
    const int mb0 = work[1 + run], row = mb0 / g.mb_w;
    const p264b200_mb *mbs = fd.mbs;
    IntraCtx ctx;
    ctx.cur[0] = fd.cur[0], ctx.cur[1] = fd.cur[1], ctx.cur[2] = fd.cur[2];
    ctx.coefs = fd.coefs;
    ctx.chroma_qp_off = fd.chroma_qp_off;

    for (int mb = mb0, mbx = mb0 - row * g.mb_w; mbx < g.mb_w && P264B200_IS_INTRA(mbs[mb].mb_type); mb++, mbx++) {
        if (row > 0) {
            // lanes 0..2: top-left, top, top-right; only intra neighbours are reconstructed by this kernel
            const int nx = mbx - 1 + lane;
            bool waited = false;
            if (lane < 3 && nx >= 0 && nx < g.mb_w) {
                const int nb = mb - g.mb_w - 1 + lane;
                if (P264B200_IS_INTRA(mbs[nb].mb_type)) {
                    unsigned ns = 32;
                    while (ld_relaxed(done + nb) != epoch) {
                        __nanosleep(ns);
                        if (ns < 512) ns <<= 1;
                    }
                    waited = true;
                }
            }
            if (waited) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            __syncwarp();
        }
        if (lane < 6) reinterpret_cast<uint4 *>(&s.mb)[lane] = __ldg(reinterpret_cast<const uint4 *>(mbs + mb) + lane);
        __syncwarp();
        recon_intra_mb(s, ctx, g, s.mb, mbx, row, lane);
        __syncwarp();
        if (lane == 0) st_release(done + mb, epoch);  // release.gpu, cumulative over the other lanes' stores (__syncwarp)
    }
}

#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
