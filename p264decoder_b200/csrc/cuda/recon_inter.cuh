// Inter macroblock reconstruction: on-the-fly quarter-pel luma / eighth-pel chroma MC from the
// integer reference plane + fused dequant / 4x4 inverse transform / residual add.
//
// Replaces p264_mb_mc (core/macroblock.c:506-524,633-717), mc_luma + the half-pel planes of
// p264_frame_filter (core/mc.c:172-266,409-451), motion_compensation_chroma (core/mc.c:303-334)
// and the inter branch of p264_macroblock_decode (decoder/macroblock.c:832-890).
//
// Work decomposition: a CTA owns kMbPerCta consecutive macroblocks of one lane.  Threads
// [0, 16*kMbPerCta) each own one luma 4x4 block (MC window 9x9, transform in registers);
// threads [16*kMbPerCta, 24*kMbPerCta) each own one chroma 4x4 block (four 2x2 MC cells).
// Every sample depends only on its own 4x4 block's (ref, mv), so no partition walk is needed.
#pragma once
#include "common.cuh"

namespace p264b200 {

constexpr int kMbPerCta = 8;
constexpr int kInterThreads = 24 * kMbPerCta;

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * (b + e) + 20 * (c + d) + f; }

// 12 consecutive samples starting at p (any alignment) as three packed words
__device__ __forceinline__ void load_row12(const uint8_t *p, uint32_t &a, uint32_t &b, uint32_t &c)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(addr & ~uintptr_t(3));
    const int sh = (int)(addr & 3) * 8;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
    a = __funnelshift_r(w0, w1, sh);
    b = __funnelshift_r(w1, w2, sh);
    c = __funnelshift_r(w2, w3, sh);
}
__device__ __forceinline__ int byte_of(uint32_t w, int i) { return (int)((w >> (8 * i)) & 0xff); }

// Quarter-pel luma prediction of one 4x4 block.  `src` points at the integer sample the MV's
// integer part selects (already clamped into the padded plane).  H.264 8.4.2.2.1 with the
// reference's rounding points: b,h = clip((tap+16)>>5), j = clip((tap(tap)+512)>>10),
// quarter positions = (s1+s2+1)>>1 of the two neighbours mc_luma picks (core/mc.c:244-257).
__device__ __forceinline__ void mc_luma_4x4(const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[4])
{
    const int dx = fx == 3, dy = fy == 3;
    // window rows -2..6, columns -2..9 of `src`
    uint32_t wa[9], wb[9], wc[9];
#pragma unroll
    for (int r = 0; r < 9; r++) load_row12(src + (r - 2) * stride - 2, wa[r], wb[r], wc[r]);

    const bool need_h = fx != 0, need_v = fy != 0;
    const bool need_j = need_h && need_v && (fx == 2 || fy == 2);
    // horizontal 6-tap intermediates, un-rounded, for the rows that are used
    int hm[9][4];
    if (__any_sync(__activemask(), need_h)) {
#pragma unroll
        for (int r = 0; r < 9; r++) {
            int p[9];
#pragma unroll
            for (int k = 0; k < 4; k++) p[k] = byte_of(wa[r], k), p[4 + k] = byte_of(wb[r], k);
            p[8] = byte_of(wc[r], 0);
#pragma unroll
            for (int c = 0; c < 4; c++) hm[r][c] = tap6(p[c], p[c + 1], p[c + 2], p[c + 3], p[c + 4], p[c + 5]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 9; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) hm[r][c] = 0;
    }
    // the column used by the vertical half-sample (x or x+1) moved to fixed byte lanes
    uint32_t vcol[9];
#pragma unroll
    for (int r = 0; r < 9; r++) vcol[r] = __funnelshift_r(wa[r], wb[r], 8 * (2 + dx));

#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint32_t o = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // integer sample: G, G(x+1) for fx==3 & fy==0, G(y+1) for fy==3 & fx==0
            const uint32_t grow_a = (dy && fx == 0) ? wa[r + 3] : wa[r + 2];
            const uint32_t grow_b = (dy && fx == 0) ? wb[r + 3] : wb[r + 2];
            const int gsh = 8 * (2 + ((dx && fy == 0) ? 1 : 0));
            const int g = byte_of(__funnelshift_r(grow_a, grow_b, gsh), c);
            const int bm = dy ? hm[r + 3][c] : hm[r + 2][c];
            const int bq = clip8i((bm + 16) >> 5);
            int hq = 0, jq = 0;
            if (need_v) {
                hq = clip8i((tap6(byte_of(vcol[r], c), byte_of(vcol[r + 1], c), byte_of(vcol[r + 2], c),
                                  byte_of(vcol[r + 3], c), byte_of(vcol[r + 4], c), byte_of(vcol[r + 5], c)) +
                             16) >>
                            5);
                if (need_j)
                    jq = clip8i((tap6(hm[r][c], hm[r + 1][c], hm[r + 2][c], hm[r + 3][c], hm[r + 4][c], hm[r + 5][c]) + 512) >> 10);
            }
            int X, Y;
            if (!need_h && !need_v)
                X = Y = g;
            else if (!need_v) {
                X = bq;
                Y = (fx & 1) ? g : bq;
            } else if (!need_h) {
                X = hq;
                Y = (fy & 1) ? g : hq;
            } else if (fx == 2 && fy == 2)
                X = Y = jq;
            else if (fx == 2) {
                X = jq;
                Y = bq;
            } else if (fy == 2) {
                X = jq;
                Y = hq;
            } else {
                X = bq;
                Y = hq;
            }
            o |= (uint32_t)((X + Y + 1) >> 1) << (8 * c);
        }
        out[r] = o;
    }
}

// eighth-pel bilinear chroma prediction of one 2x2 cell (core/mc.c:303-334)
__device__ __forceinline__ void mc_chroma_2x2(const uint8_t *__restrict__ src, int stride, int dx, int dy, int o[4])
{
    const int cA = (8 - dx) * (8 - dy), cB = dx * (8 - dy), cC = (8 - dx) * dy, cD = dx * dy;
    int p[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) p[r][c] = __ldg(src + r * stride + c);
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++)
            o[r * 2 + c] = (cA * p[r][c] + cB * p[r][c + 1] + cC * p[r + 1][c] + cD * p[r + 1][c + 1] + 32) >> 6;
}

#ifdef P264B200_DEFINE_KERNELS
__global__ void __launch_bounds__(kInterThreads) recon_inter_kernel(const FrameDesc *__restrict__ descs, Geometry g)
{
    const FrameDesc &fd = descs[blockIdx.y];
    if (fd.slice_type != P264B200_SLICE_P) return;
    const int n_mb = g.mb_w * g.mb_h;
    const int tid = threadIdx.x;
    const bool luma = tid < 16 * kMbPerCta;
    const int mb_local = luma ? (tid >> 4) : ((tid - 16 * kMbPerCta) >> 3);
    const int mb_xy = blockIdx.x * kMbPerCta + mb_local;
    if (mb_xy >= n_mb) return;
    const p264b200_mb &m = fd.mbs[mb_xy];
    if (P264B200_IS_INTRA(m.mb_type)) return;
    const int mbx = mb_xy % g.mb_w, mby = mb_xy / g.mb_w;

    if (luma) {
        const int b = tid & 15, bx = b & 3, by = b >> 2;
        const int ref = mb_ref8(m, b);
        const int mvx = m.mv[b][0], mvy = m.mv[b][1];
        // integer position, clamped so the 9x9 window (plus word alignment slack) stays inside
        // the 32-sample border; beyond the clamp every tap sees replicated edge samples anyway
        const int x0 = clip3i(16 * mbx + 4 * bx + (mvx >> 2), -16, g.width + 8);
        const int y0 = clip3i(16 * mby + 4 * by + (mvy >> 2), -16, g.height + 8);
        uint32_t px[4];
        mc_luma_4x4(fd.ref[ref][0] + (ptrdiff_t)y0 * g.y_stride + x0, g.y_stride, mvx & 3, mvy & 3, px);
        if (m.luma_mask >> b & 1) {
            const int idx = __popc(m.luma_mask & ((1u << b) - 1));
            residual4x4(fd.coefs + m.coef_off + 16 * idx, m.qp, false, 0, px);
        }
        uint8_t *dst = fd.cur[0] + (ptrdiff_t)(16 * mby + 4 * by) * g.y_stride + 16 * mbx + 4 * bx;
#pragma unroll
        for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(dst + r * g.y_stride) = px[r];
    } else {
        const int cb = (tid - 16 * kMbPerCta) & 7, plane = 1 + (cb >> 2), i = cb & 3;
        const int cx = i & 1, cy = i >> 1;  // 4x4 chroma block inside the 8x8
        uint32_t px[4] = {0, 0, 0, 0};
#pragma unroll
        for (int s = 0; s < 4; s++) {
            // 2x2 cell s of this chroma block <-> luma 4x4 block (2cx + s&1, 2cy + s>>1)
            const int lb = (2 * cy + (s >> 1)) * 4 + 2 * cx + (s & 1);
            const int ref = mb_ref8(m, lb);
            const int mvx = m.mv[lb][0], mvy = m.mv[lb][1];
            const int x0 = clip3i(8 * mbx + 4 * cx + 2 * (s & 1) + (mvx >> 3), -8, g.width / 2 + 4);
            const int y0 = clip3i(8 * mby + 4 * cy + 2 * (s >> 1) + (mvy >> 3), -8, g.height / 2 + 4);
            int o[4];
            mc_chroma_2x2(fd.ref[ref][plane] + (ptrdiff_t)y0 * g.c_stride + x0, g.c_stride, mvx & 7, mvy & 7, o);
            const int r0 = 2 * (s >> 1), c0 = 2 * (s & 1);
            px[r0] |= (uint32_t)(o[0] | (o[1] << 8)) << (8 * c0);
            px[r0 + 1] |= (uint32_t)(o[2] | (o[3] << 8)) << (8 * c0);
        }
        if (m.cbp_chroma) {
            const int qpc = c_chroma_qp[clip3i(m.qp + fd.chroma_qp_off, 0, 51)];
            const int16_t *cf = fd.coefs + m.coef_off + 16 * __popc(m.luma_mask);
            int dc[4];
            chroma_dc(cf + 4 * (plane - 1), qpc, dc);
            const int16_t *ac = nullptr;
            if (m.chroma_mask >> cb & 1) ac = cf + 8 + 16 * __popc(m.chroma_mask & ((1u << cb) - 1));
            residual4x4(ac, qpc, true, dc[i], px);
        }
        uint8_t *dst = fd.cur[plane] + (ptrdiff_t)(8 * mby + 4 * cy) * g.c_stride + 8 * mbx + 4 * cx;
#pragma unroll
        for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(dst + r * g.c_stride) = px[r];
    }
}

#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
