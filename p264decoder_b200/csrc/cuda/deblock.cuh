// In-loop deblocking filter as a macroblock-row wavefront.
//
// Replaces p264_frame_deblocking_filter (core/frame.c:490-643), deblock_edge (:472-488) and the
// eight edge filters (:302-470).  The reference filters macroblocks in raster order, vertical
// edges then horizontal edges per macroblock; MB(x,y) therefore depends on MB(x-1,y) (whole MB)
// and on MB(x+1,y-1) (its left-edge filter rewrites columns 13..15 of MB(x,y-1), which MB(x,y)'s
// top-edge filter reads).  A picture-wide "all vertical, then all horizontal" pass is NOT
// bit-exact, so the kernel keeps the reference order: one warp per macroblock row, row y may
// filter MB x once row y-1 has published progress >= min(x+2, mb_w).
//
// Inside a macroblock the warp is edge-parallel: lanes 0..15 own the 16 luma lines crossing the
// current edge direction, lanes 16..23 / 24..31 the 8 Cb / Cr lines; the 32 boundary strengths
// (2 directions x 4 edges x 4 segments) are derived one per lane.
#pragma once
#include "common.cuh"

namespace p264b200 {

constexpr int kDS = 24;  // luma tile stride: rows -4..15, cols -4..15
constexpr int kDC = 12;  // chroma tile stride: rows -4..7, cols -4..7

struct DeblockSmem {
    uint8_t y[20 * kDS];
    uint8_t c[2][12 * kDC];
    uint8_t bs[32];
};

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

__device__ __forceinline__ void deblock_mb(DeblockSmem &s, const FrameDesc &fd, const Geometry &g,
                                           const p264b200_mb *mbs_row, int mbx, int mby, int lane)
{
    const p264b200_mb &m = mbs_row[mbx];
    // ---- boundary strengths, one per lane: lane = dir*16 + edge*4 + segment (core/frame.c:535-581)
    int qp_left = m.qp_dbf, qp_top = m.qp_dbf;
    {
        const int dir = lane >> 4, e = (lane >> 2) & 3, i = lane & 3;
        int bs = 0;
        const bool have = e > 0 || (dir == 0 ? mbx > 0 : mby > 0);
        if (have) {
            const p264b200_mb &n = e > 0 ? m : (dir == 0 ? mbs_row[mbx - 1] : mbs_row[mbx - g.mb_w]);
            if (P264B200_IS_INTRA(m.mb_type) || P264B200_IS_INTRA(n.mb_type))
                bs = e == 0 ? 4 : 3;
            else {
                const int x = dir == 0 ? e : i, y = dir == 0 ? i : e;
                const int xn = (x - (dir == 0)) & 3, yn = (y - (dir == 1)) & 3;
                const int bq = y * 4 + x, bp = yn * 4 + xn;
                if ((m.luma_mask >> bq & 1) || (n.luma_mask >> bp & 1))
                    bs = 2;
                else if (mb_ref8(m, bq) != mb_ref8(n, bp) || abs(m.mv[bq][0] - n.mv[bp][0]) >= 4 ||
                         abs(m.mv[bq][1] - n.mv[bp][1]) >= 4)
                    bs = 1;
            }
        }
        s.bs[lane] = (uint8_t)bs;
        if (mbx > 0) qp_left = mbs_row[mbx - 1].qp_dbf;
        if (mby > 0) qp_top = mbs_row[mbx - g.mb_w].qp_dbf;
        if (!__any_sync(0xffffffffu, bs != 0)) return;
    }

    uint8_t *gy = fd.cur[0] + (ptrdiff_t)16 * mby * g.y_stride + 16 * mbx;
    uint8_t *gc[2] = {fd.cur[1] + (ptrdiff_t)8 * mby * g.c_stride + 8 * mbx,
                      fd.cur[2] + (ptrdiff_t)8 * mby * g.c_stride + 8 * mbx};
    // ---- load tiles (words; L2 loads because neighbouring rows' warps rewrite these samples)
    for (int w = lane; w < 100; w += 32) {
        const int r = w / 5 - 4, cw = w % 5 - 1;
        *reinterpret_cast<uint32_t *>(&s.y[(r + 4) * kDS + 4 * (cw + 1)]) =
            __ldcg(reinterpret_cast<const uint32_t *>(gy + (ptrdiff_t)r * g.y_stride + 4 * cw));
    }
    for (int w = lane; w < 60; w += 32) {
        const int p = w / 30, q = w % 30, r = q / 3 - 2, cw = q % 3 - 1;
        *reinterpret_cast<uint32_t *>(&s.c[p][(r + 4) * kDC + 4 * (cw + 1)]) =
            __ldcg(reinterpret_cast<const uint32_t *>(gc[p] + (ptrdiff_t)r * g.c_stride + 4 * cw));
    }
    __syncwarp();

    const int qp = m.qp_dbf;
    const int off = fd.chroma_qp_off;
    const int qpc_self = c_chroma_qp[clip3i(qp + off, 0, 51)];
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const int qpn = dir == 0 ? qp_left : qp_top;
        if (lane < 16) {
            // luma line `lane`: dir 0 -> row, filter across columns; dir 1 -> column, filter across rows
            const int xs = dir == 0 ? 1 : kDS, ys = dir == 0 ? kDS : 1;
            uint8_t *line = &s.y[4 * kDS + 4] + lane * ys;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int bs = s.bs[dir * 16 + e * 4 + (lane >> 2)];
                if (bs == 0) continue;
                const int q = e == 0 ? (qp + qpn + 1) >> 1 : qp;
                const int ia = clip3i(q + fd.alpha_off, 0, 51);
                const int alpha = c_alpha[ia], beta = c_beta[clip3i(q + fd.beta_off, 0, 51)];
                uint8_t *px = line + 4 * e * xs;
                int v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = px[(k - 4) * xs];
                if (bs < 4)
                    dbf_luma_normal(v, alpha, beta, c_tc0[ia][bs - 1]);
                else
                    dbf_luma_strong(v, alpha, beta);
#pragma unroll
                for (int k = 1; k < 7; k++) px[(k - 4) * xs] = (uint8_t)v[k];
            }
        } else {
            const int p = (lane - 16) >> 3, l = (lane - 16) & 7;
            const int xs = dir == 0 ? 1 : kDC, ys = dir == 0 ? kDC : 1;
            uint8_t *line = &s.c[p][4 * kDC + 4] + l * ys;
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
                const int bs = s.bs[dir * 16 + e * 4 + (l >> 1)];
                if (bs == 0) continue;
                const int qc = e == 0 ? (qpc_self + c_chroma_qp[clip3i(qpn + off, 0, 51)] + 1) >> 1 : qpc_self;
                const int ia = clip3i(qc + fd.alpha_off, 0, 51);
                const int alpha = c_alpha[ia], beta = c_beta[clip3i(qc + fd.beta_off, 0, 51)];
                uint8_t *px = line + 2 * e * xs;
                int v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = px[(k - 2) * xs];
                dbf_chroma(v, alpha, beta, bs, bs < 4 ? c_tc0[ia][bs - 1] + 1 : 0);
                px[-xs] = (uint8_t)v[1];
                px[0] = (uint8_t)v[2];
            }
        }
        __syncwarp();
    }

    // ---- write back rows -3..15 x words -1..3 (luma), rows -1..7 x words -1..1 (chroma)
    for (int w = lane; w < 95; w += 32) {
        const int r = w / 5 - 3, cw = w % 5 - 1;
        if ((r < 0 && mby == 0) || (cw < 0 && mbx == 0)) continue;
        __stcg(reinterpret_cast<uint32_t *>(gy + (ptrdiff_t)r * g.y_stride + 4 * cw),
               *reinterpret_cast<const uint32_t *>(&s.y[(r + 4) * kDS + 4 * (cw + 1)]));
    }
    for (int w = lane; w < 54; w += 32) {
        const int p = w / 27, q = w % 27, r = q / 3 - 1, cw = q % 3 - 1;
        if ((r < 0 && mby == 0) || (cw < 0 && mbx == 0)) continue;
        __stcg(reinterpret_cast<uint32_t *>(gc[p] + (ptrdiff_t)r * g.c_stride + 4 * cw),
               *reinterpret_cast<const uint32_t *>(&s.c[p][(r + 4) * kDC + 4 * (cw + 1)]));
    }
}

#ifdef P264B200_DEFINE_KERNELS
__global__ void __launch_bounds__(32) deblock_kernel(const FrameDesc *__restrict__ descs, Geometry g, int *ticket)
{
    __shared__ DeblockSmem s;
    __shared__ int s_ticket;
    const int lane = threadIdx.x;
    if (lane == 0) s_ticket = atomicAdd(ticket, 1);
    __syncwarp();
    const int t = s_ticket;
    const int lane_id = t / g.mb_h, row = t % g.mb_h;
    const FrameDesc &fd = descs[lane_id];
    if (!fd.deblock) return;
    int *prog = fd.row_progress + g.mb_h;  // second half: deblock wavefront
    const p264b200_mb *mbs_row = fd.mbs + (size_t)row * g.mb_w;

    for (int mbx = 0; mbx < g.mb_w; mbx++) {
        if (row > 0) {
            const int need = min(mbx + 2, g.mb_w);
            if (lane == 0)
                while (ld_acquire(prog + row - 1) < need) __nanosleep(32);
            __syncwarp();
        }
        deblock_mb(s, fd, g, mbs_row, mbx, row, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(prog + row, mbx + 1);
    }
}

#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
