// Inter macroblock reconstruction: on-the-fly quarter-pel luma / eighth-pel chroma MC from the
// integer reference plane + fused dequant / 4x4 inverse transform / residual add.
//
// Replaces p264_mb_mc (core/macroblock.c:506-524,633-717), mc_luma + the half-pel planes of
// p264_frame_filter (core/mc.c:172-266,409-451), motion_compensation_chroma (core/mc.c:303-334)
// and the inter branch of p264_macroblock_decode (decoder/macroblock.c:832-890).
//
// Work decomposition: a CTA owns kMbPerCta consecutive macroblocks of one lane.  Threads
// [0, 16*kMbPerCta) each own one luma 4x4 block, threads [16*kMbPerCta, 24*kMbPerCta) one chroma
// 4x4 block (four 2x2 MC cells).  Every sample depends only on its own 4x4 block's (ref, mv), so
// no partition walk is needed.
//
// v2 (instruction-bound in v1, DRAM traffic already == algorithmic bytes):
//  * the CTA's 128 luma blocks are counting-sorted by interpolation class (copy / H only / V only /
//    diagonal / centre) so that a warp runs one class and skips the filter stages it does not need
//    with warp-uniform branches instead of paying for the union of all 16 phases;
//  * 6-tap filters are byte dot products: horizontal taps = 2 x dp4a on funnel-shifted words, vertical
//    taps = byte transpose (PRMT) + 2 x dp4a, centre taps = dp2a on packed 16-bit intermediates; the
//    +16 rounding rides in the dp4a accumulator, which also absorbs the centre's +512 exactly
//    (sum of taps = 32, 32 * 16 = 512);
//  * chroma bilinear samples are one PRMT + one dp4a each.
#pragma once
#include "common.cuh"

namespace p264b200 {

constexpr int kMbPerCta = 8;
constexpr int kLumaThreads = 16 * kMbPerCta;
constexpr int kInterThreads = 24 * kMbPerCta;

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(int a, int b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(int a, int b, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int kTapA = 0x1414FB01;   // bytes ( 1, -5, 20, 20)
constexpr int kTapB = 0x000001FB;   // bytes (-5,  1,  0,  0)
constexpr int kTapOdd0 = 0x14FB0100;  // bytes (0, 1 | -5, 20): rows (r-1, r) then (r+1, r+2) of an odd-aligned 6-tap
constexpr int kTapOdd1 = 0x0001FB14;  // bytes (20, -5 | 1, 0)

// 6-tap over 6 consecutive bytes starting at byte k (0..3) of the 12-byte string (w0, w1, w2), + acc
__device__ __forceinline__ int tap6_bytes(uint32_t w0, uint32_t w1, uint32_t w2, int k, int acc)
{
    const uint32_t a = k == 0 ? w0 : __funnelshift_r(w0, w1, 8 * k);
    const uint32_t b = k == 0 ? w1 : (k == 3 ? __funnelshift_r(w1, w2, 24) : (w1 >> (8 * k)));
    return dp4a_us(b, kTapB, dp4a_us(a, kTapA, acc));
}

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
// 8 consecutive samples (enough for the vertical filter's 4 columns at x or x+1)
__device__ __forceinline__ void load_row8(const uint8_t *p, uint32_t &a, uint32_t &b)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(addr & ~uintptr_t(3));
    const int sh = (int)(addr & 3) * 8;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    a = __funnelshift_r(w0, w1, sh);
    b = __funnelshift_r(w1, w2, sh);
}
__device__ __forceinline__ int byte_of(uint32_t w, int i) { return (int)((w >> (8 * i)) & 0xff); }

// interpolation class of a quarter-pel phase: 0 copy, 1 horizontal only, 2 vertical only,
// 3 diagonal (b and h, no centre), 4 centre j involved
__device__ __forceinline__ int mc_class(int fx, int fy)
{
    if (fx == 0) return fy == 0 ? 0 : 2;
    if (fy == 0) return 1;
    return (fx == 2 || fy == 2) ? 4 : 3;
}

// Quarter-pel luma prediction of one 4x4 block.  `src` points at the integer sample the MV's
// integer part selects (already clamped into the padded plane).  H.264 8.4.2.2.1 with the
// reference's rounding points: b,h = clip((tap+16)>>5), j = clip((tap(tap)+512)>>10),
// quarter positions = (s1+s2+1)>>1 of the two neighbours mc_luma picks (core/mc.c:244-257).
// Stages are skipped per warp (the caller groups threads by class), never per thread.
__device__ __forceinline__ void mc_luma_4x4(const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[4])
{
    const int dx = fx == 3, dy = fy == 3;
    const bool need_h = fx != 0, need_v = fy != 0;
    const bool need_j = need_h && need_v && (fx == 2 || fy == 2);
    const unsigned am = __activemask();
    const bool w_h = __any_sync(am, need_h), w_v = __any_sync(am, need_v), w_j = __any_sync(am, need_j);

    // window rows 0..8 = picture rows -2..6, words: a = cols -2..1, b = cols 2..5, c = cols 6..9
    uint32_t wa[9], wb[9], wc[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        wa[r] = wb[r] = wc[r] = 0;
        const bool row_needed = w_v || (r >= 2 && r <= 5);
        if (row_needed) {
            if (w_h)
                load_row12(src + (r - 2) * stride - 2, wa[r], wb[r], wc[r]);
            else
                load_row8(src + (r - 2) * stride - 2, wa[r], wb[r]);
        }
    }

    // horizontal 6-tap + 16 for the rows in use: all 9 for the centre, rows 2..6 otherwise
    int hm[9][4];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const bool row_needed = w_h && (w_j || (r >= 2 && r <= 6 && (w_v || r <= 5)));
#pragma unroll
        for (int c = 0; c < 4; c++) hm[r][c] = row_needed ? tap6_bytes(wa[r], wb[r], wc[r], c, 16) : 0;
    }

    // vertical half samples h at column x (or x+1): transpose the 9x4 byte block, then dp4a down each column
    uint32_t hw[4] = {0, 0, 0, 0};  // packed rows of h
    if (w_v) {
        int hq[4][4];
        uint32_t vc[9];
#pragma unroll
        for (int r = 0; r < 9; r++) vc[r] = __funnelshift_r(wa[r], wb[r], 8 * (2 + dx));
        uint32_t col[4][2];
#pragma unroll
        for (int g4 = 0; g4 < 2; g4++) {
            const uint32_t t0 = __byte_perm(vc[4 * g4 + 0], vc[4 * g4 + 1], 0x5140), t1 = __byte_perm(vc[4 * g4 + 2], vc[4 * g4 + 3], 0x5140);
            const uint32_t t2 = __byte_perm(vc[4 * g4 + 0], vc[4 * g4 + 1], 0x7362), t3 = __byte_perm(vc[4 * g4 + 2], vc[4 * g4 + 3], 0x7362);
            col[0][g4] = __byte_perm(t0, t1, 0x5410);
            col[1][g4] = __byte_perm(t0, t1, 0x7632);
            col[2][g4] = __byte_perm(t2, t3, 0x5410);
            col[3][g4] = __byte_perm(t2, t3, 0x7632);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t tail = byte_of(vc[8], c);  // row 8 of this column
#pragma unroll
            for (int r = 0; r < 4; r++) hq[r][c] = tap6_bytes(col[c][0], col[c][1], tail, r, 16) >> 5;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) hw[r] = pack4_sat_u8(hq[r][0], hq[r][1], hq[r][2], hq[r][3]);
    }

    // centre samples j: 6-tap down the (already +16) horizontal intermediates, two rows per dp2a
    uint32_t jw[4] = {0, 0, 0, 0};  // packed rows of j
    if (w_j) {
        int jq[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int pk[4];  // rows (0,1) (2,3) (4,5) (6,7) as s16x2
#pragma unroll
            for (int k = 0; k < 4; k++) pk[k] = (int)__byte_perm((uint32_t)hm[2 * k][c], (uint32_t)hm[2 * k + 1][c], 0x5410);
            const int j0 = dp2a_lo_ss(pk[2], kTapB, dp2a_hi_ss(pk[1], kTapA, dp2a_lo_ss(pk[0], kTapA, 0)));
            const int j2 = dp2a_lo_ss(pk[3], kTapB, dp2a_hi_ss(pk[2], kTapA, dp2a_lo_ss(pk[1], kTapA, 0)));
            const int j1 = dp2a_hi_ss(pk[3], kTapOdd1, dp2a_lo_ss(pk[2], kTapOdd1, dp2a_hi_ss(pk[1], kTapOdd0, dp2a_lo_ss(pk[0], kTapOdd0, 0))));
            const int j3 = dp2a_hi_ss(hm[8][c], kTapOdd1, dp2a_lo_ss(pk[3], kTapOdd1, dp2a_hi_ss(pk[2], kTapOdd0, dp2a_lo_ss(pk[1], kTapOdd0, 0))));
            jq[0][c] = j0 >> 10;
            jq[1][c] = j1 >> 10;
            jq[2][c] = j2 >> 10;
            jq[3][c] = j3 >> 10;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) jw[r] = pack4_sat_u8(jq[r][0], jq[r][1], jq[r][2], jq[r][3]);
    }

    // quarter positions: rounded average of the two samples mc_luma would pick (core/mc.c:244-257),
    // on packed rows; the operand choice depends only on the phase
    const bool x_is_j = need_j, x_is_b = need_h && !need_j, x_is_h = !need_h && need_v;
    const bool y_is_g = (need_h != need_v) && (((need_h ? fx : fy) & 1) != 0);
    const bool y_is_b = need_j && fx == 2 && fy != 2, y_is_h = need_h && need_v && fx != 2 && (fy == 2 || !need_j);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        // integer sample: G, G(x+1) for fx==3 & fy==0, G(y+1) for fy==3 & fx==0
        const uint32_t grow_a = (dy && fx == 0) ? wa[r + 3] : wa[r + 2];
        const uint32_t grow_b = (dy && fx == 0) ? wb[r + 3] : wb[r + 2];
        const uint32_t gw = __funnelshift_r(grow_a, grow_b, 8 * (2 + ((dx && fy == 0) ? 1 : 0)));
        uint32_t bw = 0;
        if (w_h) {
            const int *hr = dy ? hm[r + 3] : hm[r + 2];
            bw = pack4_sat_u8((dy ? hm[r + 3][0] : hm[r + 2][0]) >> 5, (dy ? hm[r + 3][1] : hm[r + 2][1]) >> 5,
                              (dy ? hm[r + 3][2] : hm[r + 2][2]) >> 5, (dy ? hm[r + 3][3] : hm[r + 2][3]) >> 5);
            (void)hr;
        }
        const uint32_t X = x_is_j ? jw[r] : x_is_b ? bw : x_is_h ? hw[r] : gw;
        const uint32_t Y = y_is_g ? gw : y_is_b ? bw : y_is_h ? hw[r] : X;
        out[r] = avg4_u8(X, Y);
    }
}

// eighth-pel bilinear chroma prediction of one 2x2 cell (core/mc.c:303-334): one dp4a per sample
__device__ __forceinline__ void mc_chroma_2x2(const uint8_t *__restrict__ src, int stride, int dx, int dy, int o[4])
{
    const uint32_t wgt = (uint32_t)((8 - dx) * (8 - dy)) | ((uint32_t)(dx * (8 - dy)) << 8) | ((uint32_t)((8 - dx) * dy) << 16) |
                         ((uint32_t)(dx * dy) << 24);
    uint32_t row[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src + r * stride);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(addr & ~uintptr_t(3));
        row[r] = __funnelshift_r(__ldg(w), __ldg(w + 1), (int)(addr & 3) * 8);
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        o[r * 2 + 0] = dp4a_uu(__byte_perm(row[r], row[r + 1], 0x5410), wgt, 32) >> 6;
        o[r * 2 + 1] = dp4a_uu(__byte_perm(row[r], row[r + 1], 0x6521), wgt, 32) >> 6;
    }
}

#ifdef P264B200_DEFINE_KERNELS
__global__ void __launch_bounds__(kInterThreads, 6) recon_inter_kernel(const FrameDesc *__restrict__ descs, Geometry g)
{
    constexpr int kKeys = 11;           // (coded ? 0 : 5) + class, 10 = nothing to do
    __shared__ int s_cnt[4][12];        // [luma warp][key], then exclusive start of (key, warp)
    __shared__ uint8_t s_perm[kLumaThreads];
    const FrameDesc &fd = descs[blockIdx.y];
    if (fd.slice_type != P264B200_SLICE_P) return;
    const int n_mb = g.mb_w * g.mb_h;
    const int tid = threadIdx.x;
    const bool luma = tid < kLumaThreads;
    const int mb_base = blockIdx.x * kMbPerCta;

    // ---- counting sort of the CTA's luma blocks by (has residual, interpolation class): a warp then
    // runs one filter class, and the dequant/IDCT code only runs in the warps that hold coded blocks
    int key = kKeys - 1, rank = 0;
    const int lane = tid & 31, wid = tid >> 5;
    if (luma) {
        const int mb_xy = mb_base + (tid >> 4), b = tid & 15;
        if (mb_xy < n_mb) {
            const p264b200_mb &m = fd.mbs[mb_xy];
            if (!P264B200_IS_INTRA(m.mb_type))
                key = mc_class(m.mv[b][0] & 3, m.mv[b][1] & 3) + ((m.luma_mask >> b & 1) ? 0 : 5);
        }
        int cnt = 0;
#pragma unroll
        for (int c = 0; c < kKeys; c++) {
            const unsigned msk = __ballot_sync(0xffffffffu, key == c);
            if (key == c) rank = __popc(msk & ((1u << lane) - 1));
            if (lane == c) cnt = __popc(msk);
        }
        if (lane < kKeys) s_cnt[wid][lane] = cnt;
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive prefix over (key major, warp minor), 44 entries handled by lanes 0..10
        int c4[4] = {0, 0, 0, 0}, tot = 0;
        if (lane < kKeys) {
#pragma unroll
            for (int w2 = 0; w2 < 4; w2++) c4[w2] = s_cnt[w2][lane], tot += c4[w2];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int base = incl - tot;
        if (lane < kKeys) {
#pragma unroll
            for (int w2 = 0; w2 < 4; w2++) s_cnt[w2][lane] = base, base += c4[w2];
        }
    }
    __syncthreads();
    if (luma) s_perm[s_cnt[wid][key] + rank] = (uint8_t)tid;
    __syncthreads();

    if (luma) {
        const int item = s_perm[tid];
        const int mb_xy = mb_base + (item >> 4), b = item & 15;
        if (mb_xy >= n_mb) return;
        const p264b200_mb &m = fd.mbs[mb_xy];
        if (P264B200_IS_INTRA(m.mb_type)) return;
        const int mbx = mb_xy % g.mb_w, mby = mb_xy / g.mb_w;
        const int bx = b & 3, by = b >> 2;
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
        const int mb_xy = mb_base + ((tid - kLumaThreads) >> 3);
        if (mb_xy >= n_mb) return;
        const p264b200_mb &m = fd.mbs[mb_xy];
        if (P264B200_IS_INTRA(m.mb_type)) return;
        const int mbx = mb_xy % g.mb_w, mby = mb_xy / g.mb_w;
        const int cb = (tid - kLumaThreads) & 7, plane = 1 + (cb >> 2), i = cb & 3;
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
