// Inter macroblock reconstruction: on-the-fly quarter-pel luma / eighth-pel chroma MC from the
// integer reference plane + fused dequant / 4x4 inverse transform / residual add.
//
// Replaces p264_mb_mc (core/macroblock.c:506-524,633-717), mc_luma + the half-pel planes of
// p264_frame_filter (core/mc.c:172-266,409-451), motion_compensation_chroma (core/mc.c:303-334)
// and the inter branch of p264_macroblock_decode (decoder/macroblock.c:832-890).
//
// v4 work decomposition (r02; v3 = thread per 4x4 block everywhere, 439 warp-instructions per macroblock, L1 data pipe 85 %):
//  * a CTA owns a tile of 8x8 macroblocks (128x128 luma samples) of one lane; the 64 macroblock
//    records are staged in shared memory once (coalesced 16-byte loads);
//  * classification is per 8x8 QUADRANT (one thread each, 4 per macroblock instead of 16 + 8): a column of two 4x4
//    blocks with one vector becomes a 4x8 STRIP (every partition except 8x4 / 4x4 sub-partitions), the rest stay
//    4x4 blocks ("half" items: they run the strip body and keep its first four rows); items are counted per
//    interpolation class (copy / H / V / diagonal / centre+b / centre+h, core/mc.c:244-257) with one shared-memory
//    atomic each and placed back to back after a barrier, the two strips of a one-vector quadrant side by side (the
//    lanes of a 16-wide partition then read the same cache lines in one load instruction -- the L1 data pipe charges per
//    distinct line, tools/ubench/ldg_width.cu); chroma is listed per quadrant (both planes in one item);
//  * a strip loads its 13 window rows ONCE for 8 output rows (a 4x4 block loads 9 for 4): 28 % fewer window loads
//    and horizontal taps in every class with a vertical filter, and the lanes of a partition still share cache lines;
//  * the warps of the CTA draw class-pure chunks of 32 items from the lists, heaviest class first, through a
//    shared ticket -- no warp idles at the barrier behind a long class body;
//  * predictions go to a shared-memory picture tile; blocks that carry residual are compacted into a
//    second list so dequant + inverse transform runs with full warps, on the tile;
//  * the finished tile leaves with 16-byte (luma) / 8-byte (chroma) coalesced stores.
//  Filter arithmetic: 6-tap filters are byte dot products.  The four outputs of an aligned 12-byte string use four
//  tap constants shifted by one byte each (9 IDP.4A, no funnel shifts); vertical taps run on byte-transposed
//  columns, the centre on packed 16-bit intermediates (IDP.2A); +16 / +512 rounding rides in the accumulators.
#pragma once
#include "common.cuh"

namespace p264b200 {

constexpr int kTileW = 8;          // macroblocks per CTA tile row; the tile height TH (4, 8 or 16 rows) is a kernel template parameter


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

// A block's window starts at byte `base + sh/8`; base is 4-byte aligned and the row stride is a
// multiple of 4, so one (aligned pointer, shift) pair serves every row.
struct RowPtr {
    const uint32_t *w;  // aligned word holding the first window byte of row 0
    int sh;             // 8 * (byte offset inside that word)
    int stride4;        // row stride in 32-bit words
};
__device__ __forceinline__ RowPtr row_ptr(const uint8_t *p, int stride)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    RowPtr r;
    r.w = reinterpret_cast<const uint32_t *>(addr & ~uintptr_t(3));
    r.sh = (int)(addr & 3) * 8;
    r.stride4 = stride >> 2;
    return r;
}
// 9 window bytes of row r (enough for four 6-tap outputs) as a, b and byte 0 of c
__device__ __forceinline__ void load_win9(const RowPtr &rp, int r, uint32_t &a, uint32_t &b, uint32_t &c)
{
    const uint32_t *w = rp.w + r * rp.stride4;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    a = __funnelshift_r(w0, w1, rp.sh);
    b = __funnelshift_r(w1, w2, rp.sh);
    c = w2 >> rp.sh;
}
// 4 window bytes of row r
__device__ __forceinline__ uint32_t load_win4(const RowPtr &rp, int r)
{
    const uint32_t *w = rp.w + r * rp.stride4;
    return __funnelshift_r(__ldg(w), __ldg(w + 1), rp.sh);
}
__device__ __forceinline__ int byte_of(uint32_t w, int i) { return (int)((w >> (8 * i)) & 0xff); }

// interpolation class of a quarter-pel phase:
//   0 copy, 1 horizontal only, 2 vertical only, 3 diagonal (b and h, no centre),
//   4 centre j (+ b when fy is odd), fx == 2,  5 centre j + h, fy == 2 and fx odd
enum { kMcCopy = 0, kMcH = 1, kMcV = 2, kMcDiag = 3, kMcCentreB = 4, kMcCentreH = 5, kMcClasses = 6 };
__device__ __forceinline__ int mc_class(int fx, int fy)
{
    if (fx == 0) return fy == 0 ? kMcCopy : kMcV;
    if (fy == 0) return kMcH;
    if (fx == 2) return kMcCentreB;
    return fy == 2 ? kMcCentreH : kMcDiag;
}

// four packed rows of vertical half samples h = clip((tapV + 16) >> 5) from the 9 window rows vc[0..8]
// (4 columns each): byte transpose, then two dp4a down each column
__device__ __forceinline__ void vfilter4(const uint32_t vc[9], uint32_t hw[4])
{
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
    int hq[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t tail = byte_of(vc[8], c);  // row 8 of this column
#pragma unroll
        for (int r = 0; r < 4; r++) hq[r][c] = tap6_bytes(col[c][0], col[c][1], tail, r, 16) >> 5;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) hw[r] = pack4_sat_u8(hq[r][0], hq[r][1], hq[r][2], hq[r][3]);
}

// four packed rows of centre samples j = clip((tapV(tapH) + 512) >> 10) from the 9x4 horizontal
// intermediates hm (each already + 16): two rows per dp2a
__device__ __forceinline__ void centre4(const int hm[9][4], uint32_t jw[4])
{
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

// Quarter-pel luma prediction of one 4x4 block, specialised per interpolation class.  `src` points at
// the integer sample the MV's integer part selects (already clamped into the padded plane).
// H.264 8.4.2.2.1 with the reference's rounding points: b,h = clip((tap+16)>>5),
// j = clip((tap(tap)+512)>>10), quarter positions = (s1+s2+1)>>1 of the two neighbours mc_luma picks
// (core/mc.c:244-257).  (fx, fy) must belong to class CLS.
template <int CLS>
__device__ __forceinline__ void mc_luma_cls(const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[4])
{
    const int dx = fx == 3, dy = fy == 3;
    if (CLS == kMcCopy) {
        const RowPtr rp = row_ptr(src, stride);
#pragma unroll
        for (int r = 0; r < 4; r++) out[r] = load_win4(rp, r);
    } else if (CLS == kMcH) {
        // b on rows 0..3; quarter phases average with G (fx 1) or G(x+1) (fx 3)
        const RowPtr rp = row_ptr(src - 2, stride);
        const int gsh = 16 + 8 * dx;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t a, b, c;
            load_win9(rp, r, a, b, c);
            const uint32_t bw = pack4_sat_u8(tap6_bytes(a, b, c, 0, 16) >> 5, tap6_bytes(a, b, c, 1, 16) >> 5,
                                             tap6_bytes(a, b, c, 2, 16) >> 5, tap6_bytes(a, b, c, 3, 16) >> 5);
            const uint32_t gw = __funnelshift_r(a, b, gsh);
            out[r] = avg4_u8(bw, fx == 2 ? bw : gw);
        }
    } else if (CLS == kMcV) {
        // h on columns 0..3; quarter phases average with G (fy 1) or G(y+1) (fy 3)
        const RowPtr rp = row_ptr(src - 2 * stride, stride);
        uint32_t vc[9], hw[4];
#pragma unroll
        for (int r = 0; r < 9; r++) vc[r] = load_win4(rp, r);
        vfilter4(vc, hw);
#pragma unroll
        for (int r = 0; r < 4; r++) out[r] = avg4_u8(hw[r], fy == 2 ? hw[r] : (dy ? vc[r + 3] : vc[r + 2]));
    } else if (CLS == kMcDiag) {
        // (b at row y + dy, h at column x + dx) averaged
        const RowPtr rp = row_ptr(src - 2 * stride - 2, stride);
        const int csh = 16 + 8 * dx;
        uint32_t vc[9], bw[5], hw[4];
#pragma unroll
        for (int r = 0; r < 9; r++) {
            uint32_t a, b, c;
            load_win9(rp, r, a, b, c);
            vc[r] = __funnelshift_r(a, b, csh);
            if (r >= 2 && r <= 6)
                bw[r - 2] = pack4_sat_u8(tap6_bytes(a, b, c, 0, 16) >> 5, tap6_bytes(a, b, c, 1, 16) >> 5,
                                         tap6_bytes(a, b, c, 2, 16) >> 5, tap6_bytes(a, b, c, 3, 16) >> 5);
        }
        vfilter4(vc, hw);
#pragma unroll
        for (int r = 0; r < 4; r++) out[r] = avg4_u8(dy ? bw[r + 1] : bw[r], hw[r]);
    } else {
        // centre: horizontal intermediates of all 9 rows, j down the columns
        const RowPtr rp = row_ptr(src - 2 * stride - 2, stride);
        const int csh = 16 + 8 * dx;
        int hm[9][4];
        uint32_t vc[9], jw[4];
#pragma unroll
        for (int r = 0; r < 9; r++) {
            uint32_t a, b, c;
            load_win9(rp, r, a, b, c);
            if (CLS == kMcCentreH) vc[r] = __funnelshift_r(a, b, csh);
#pragma unroll
            for (int k = 0; k < 4; k++) hm[r][k] = tap6_bytes(a, b, c, k, 16);
        }
        centre4(hm, jw);
        if (CLS == kMcCentreB) {
            // fx == 2: j alone (fy 2) or averaged with b of row y (fy 1) / y+1 (fy 3)
            uint32_t bw[5];
#pragma unroll
            for (int r = 0; r < 5; r++) bw[r] = pack4_sat_u8(hm[r + 2][0] >> 5, hm[r + 2][1] >> 5, hm[r + 2][2] >> 5, hm[r + 2][3] >> 5);
#pragma unroll
            for (int r = 0; r < 4; r++) out[r] = avg4_u8(jw[r], fy == 2 ? jw[r] : (dy ? bw[r + 1] : bw[r]));
        } else {
            // fy == 2, fx odd: j averaged with h of column x (fx 1) / x+1 (fx 3)
            uint32_t hw[4];
            vfilter4(vc, hw);
#pragma unroll
            for (int r = 0; r < 4; r++) out[r] = avg4_u8(jw[r], hw[r]);
        }
    }
}

// class dispatch (warp-uniform when the caller buckets blocks by class)
__device__ __forceinline__ void mc_luma_dispatch(int cls, const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[4])
{
    switch (cls) {
    case kMcCopy: mc_luma_cls<kMcCopy>(src, stride, fx, fy, out); break;
    case kMcH: mc_luma_cls<kMcH>(src, stride, fx, fy, out); break;
    case kMcV: mc_luma_cls<kMcV>(src, stride, fx, fy, out); break;
    case kMcDiag: mc_luma_cls<kMcDiag>(src, stride, fx, fy, out); break;
    case kMcCentreB: mc_luma_cls<kMcCentreB>(src, stride, fx, fy, out); break;
    default: mc_luma_cls<kMcCentreH>(src, stride, fx, fy, out); break;
    }
}
// any phase (the one-block table shims in blockops.cu)
__device__ __forceinline__ void mc_luma_4x4(const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[4])
{
    mc_luma_dispatch(mc_class(fx, fy), src, stride, fx, fy, out);
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
// a 4x4 chroma block whose four 2x2 cells share one MV: 5 rows x 5 samples, one dp4a per sample
__device__ __forceinline__ void mc_chroma_4x4(const uint8_t *__restrict__ src, int stride, int dx, int dy, uint32_t out[4])
{
    const uint32_t wgt = (uint32_t)((8 - dx) * (8 - dy)) | ((uint32_t)(dx * (8 - dy)) << 8) | ((uint32_t)((8 - dx) * dy) << 16) |
                         ((uint32_t)(dx * dy) << 24);
    const RowPtr rp = row_ptr(src, stride);
    uint32_t lo[5], hi[5];  // samples 0..3 and 1..4 of each row
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint32_t *w = rp.w + r * rp.stride4;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
        lo[r] = __funnelshift_r(w0, w1, rp.sh);
        hi[r] = __funnelshift_r(lo[r], w1 >> rp.sh, 8);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int v[4];
        // sample c: bytes (A, B, C, D) = (row r col c, row r col c+1, row r+1 col c, row r+1 col c+1)
        v[0] = dp4a_uu(__byte_perm(lo[r], lo[r + 1], 0x5410), wgt, 32) >> 6;
        v[1] = dp4a_uu(__byte_perm(lo[r], lo[r + 1], 0x6521), wgt, 32) >> 6;
        v[2] = dp4a_uu(__byte_perm(lo[r], lo[r + 1], 0x7632), wgt, 32) >> 6;
        v[3] = dp4a_uu(__byte_perm(hi[r], hi[r + 1], 0x7632), wgt, 32) >> 6;
        out[r] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    }
}

// ------------------------------------------------------------------ 4x8 strips
// Tap constants of the four outputs k = 0..3 of an ALIGNED 12-byte string (a = bytes 0..3, b = 4..7, c = 8..11):
// out_k = sum_i tap[i] * byte[k + i], i.e. the 6-tap (1,-5,20,20,-5,1) shifted by k bytes inside the dot-product operands.
constexpr int kT0 = 0x1414FB01, kU0 = 0x000001FB;                    // a ( 1,-5,20,20)  b (-5, 1, 0, 0)
constexpr int kT1 = 0x14FB0100, kU1 = 0x0001FB14;                    // a ( 0, 1,-5,20)  b (20,-5, 1, 0)
constexpr int kT2 = (int)0xFB010000, kU2 = 0x01FB1414;               // a ( 0, 0, 1,-5)  b (20,20,-5, 1)
constexpr int kT3 = 0x01000000, kU3 = (int)0xFB1414FB, kV3 = 0x00000001;  // a (0,0,0,1) b (-5,20,20,-5) c (1,0,0,0)

// the four 6-tap sums of the string (a, b, c) + acc; c contributes its byte 0 only
__device__ __forceinline__ void tap6x4(uint32_t a, uint32_t b, uint32_t c, int acc, int o[4])
{
    o[0] = dp4a_us(b, kU0, dp4a_us(a, kT0, acc));
    o[1] = dp4a_us(b, kU1, dp4a_us(a, kT1, acc));
    o[2] = dp4a_us(b, kU2, dp4a_us(a, kT2, acc));
    o[3] = dp4a_us(c, kV3, dp4a_us(b, kU3, dp4a_us(a, kT3, acc)));
}
__device__ __forceinline__ uint32_t pack4_shr5(const int o[4]) { return pack4_sat_u8(o[0] >> 5, o[1] >> 5, o[2] >> 5, o[3] >> 5); }

// eight packed rows of vertical half samples h = clip((tapV + 16) >> 5) from the 13 window rows vc[0..12] (4 columns
// each): byte transpose of rows 0..11 into column words, row 12 is picked out of vc[12] by the dot product itself
__device__ __forceinline__ void vfilter8(const uint32_t vc[13], uint32_t hw[8])
{
    uint32_t col[4][3];
#pragma unroll
    for (int g4 = 0; g4 < 3; g4++) {
        const uint32_t t0 = __byte_perm(vc[4 * g4 + 0], vc[4 * g4 + 1], 0x5140), t1 = __byte_perm(vc[4 * g4 + 2], vc[4 * g4 + 3], 0x5140);
        const uint32_t t2 = __byte_perm(vc[4 * g4 + 0], vc[4 * g4 + 1], 0x7362), t3 = __byte_perm(vc[4 * g4 + 2], vc[4 * g4 + 3], 0x7362);
        col[0][g4] = __byte_perm(t0, t1, 0x5410);
        col[1][g4] = __byte_perm(t0, t1, 0x7632);
        col[2][g4] = __byte_perm(t2, t3, 0x5410);
        col[3][g4] = __byte_perm(t2, t3, 0x7632);
    }
    int hq[8][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int o[4];
        tap6x4(col[c][0], col[c][1], col[c][2], 16, o);
#pragma unroll
        for (int r = 0; r < 4; r++) hq[r][c] = o[r] >> 5;
        // rows 4..7: the last tap of row 7 is row 12 = byte c of vc[12]
        hq[4][c] = dp4a_us(col[c][2], kU0, dp4a_us(col[c][1], kT0, 16)) >> 5;
        hq[5][c] = dp4a_us(col[c][2], kU1, dp4a_us(col[c][1], kT1, 16)) >> 5;
        hq[6][c] = dp4a_us(col[c][2], kU2, dp4a_us(col[c][1], kT2, 16)) >> 5;
        hq[7][c] = dp4a_us(vc[12], 1 << (8 * c), dp4a_us(col[c][2], kU3, dp4a_us(col[c][1], kT3, 16))) >> 5;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) hw[r] = pack4_sat_u8(hq[r][0], hq[r][1], hq[r][2], hq[r][3]);
}

// eight packed rows of centre samples j = clip((tapV(tapH) + 512) >> 10) from the horizontal intermediates (each already
// + 16) of the 13 window rows, paired into s16x2 words as they are produced: pk[c][k] = rows (2k, 2k+1) of column c,
// pk[c][6] = row 12 in the low half.  Even output rows take three dp2a, odd ones four.
__device__ __forceinline__ void centre8(const int pk[4][7], uint32_t jw[8])
{
    int jq[8][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            jq[2 * m][c] = dp2a_lo_ss(pk[c][m + 2], kTapB, dp2a_hi_ss(pk[c][m + 1], kTapA, dp2a_lo_ss(pk[c][m], kTapA, 0))) >> 10;
            jq[2 * m + 1][c] = dp2a_hi_ss(pk[c][m + 3], kTapOdd1,
                                          dp2a_lo_ss(pk[c][m + 2], kTapOdd1, dp2a_hi_ss(pk[c][m + 1], kTapOdd0, dp2a_lo_ss(pk[c][m], kTapOdd0, 0)))) >> 10;
        }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) jw[r] = pack4_sat_u8(jq[r][0], jq[r][1], jq[r][2], jq[r][3]);
}

// Quarter-pel luma prediction of one 4-wide, 8-tall strip (two vertically adjacent 4x4 blocks with one vector); same
// arithmetic as mc_luma_cls, but the 13 window rows are loaded and horizontally filtered once for the eight output rows.
// `cls` is WARP-UNIFORM in the frame kernel (class-pure chunks), so the branches below do not diverge; the classes with a
// vertical filter share two bodies (V + diagonal, centre+b + centre+h) to keep the kernel inside the instruction cache
// (six fully specialised strip bodies + six block bodies were 114 KB of code: 37 % of the issue slots starved).
__device__ __forceinline__ void mc_luma_strip(const uint8_t *__restrict__ src, int stride, int fx, int fy, int cls, uint32_t out[8])
{
    const int dx = fx == 3, dy = fy == 3;
    if (cls == kMcCopy) {
        const RowPtr rp = row_ptr(src, stride);
#pragma unroll
        for (int r = 0; r < 8; r++) out[r] = load_win4(rp, r);
    } else if (cls == kMcH) {
        // b on rows 0..7; quarter phases average with G (fx 1) or G(x+1) (fx 3)
        const RowPtr rp = row_ptr(src - 2, stride);
        const int gsh = 16 + 8 * dx;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint32_t a, b, c;
            load_win9(rp, r, a, b, c);
            int o[4];
            tap6x4(a, b, c, 16, o);
            const uint32_t bw = pack4_shr5(o);
            const uint32_t gw = __funnelshift_r(a, b, gsh);
            out[r] = avg4_u8(bw, fx == 2 ? bw : gw);
        }
    } else if (cls == kMcV || cls == kMcDiag) {
        // h on the integer columns x (V) / x + dx (diagonal); the diagonal averages it with b of row y + dy
        const bool diag = cls == kMcDiag;
        uint32_t vc[13], bw[13], hw[8];
        if (diag) {
            const RowPtr rp = row_ptr(src - 2 * stride - 2, stride);
            const int csh = 16 + 8 * dx;
#pragma unroll
            for (int r = 0; r < 13; r++) {
                uint32_t a, b, c;
                load_win9(rp, r, a, b, c);
                vc[r] = __funnelshift_r(a, b, csh);
                if (r >= 2 && r <= 10) {
                    int o[4];
                    tap6x4(a, b, c, 16, o);
                    bw[r] = pack4_shr5(o);
                }
            }
        } else {
            const RowPtr rp = row_ptr(src - 2 * stride, stride);
#pragma unroll
            for (int r = 0; r < 13; r++) bw[r] = vc[r] = load_win4(rp, r);
        }
        vfilter8(vc, hw);
        // V: quarter phases average with G of row y (fy 1) / y + 1 (fy 3) = bw[r + 2 + dy]; same index for the diagonal's b rows
#pragma unroll
        for (int r = 0; r < 8; r++) out[r] = avg4_u8(hw[r], (!diag && fy == 2) ? hw[r] : (dy ? bw[r + 3] : bw[r + 2]));
    } else {
        // centre j from the horizontal intermediates of all 13 rows; + b of row y + dy (fx == 2) or + h of column x + dx (fy == 2)
        const bool ch = cls == kMcCentreH;
        const RowPtr rp = row_ptr(src - 2 * stride - 2, stride);
        const int csh = 16 + 8 * dx;
        int pk[4][7], hold[4];
        uint32_t aux[13], jw[8];   // centre+h: the integer columns; centre+b: the b rows
#pragma unroll
        for (int r = 0; r < 13; r++) {
            uint32_t a, b, c;
            load_win9(rp, r, a, b, c);
            int o[4];
            tap6x4(a, b, c, 16, o);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (r == 12)
                    pk[k][6] = o[k];
                else if (r & 1)
                    pk[k][r >> 1] = (int)__byte_perm((uint32_t)hold[k], (uint32_t)o[k], 0x5410);
                else
                    hold[k] = o[k];
            }
            if (ch)
                aux[r] = __funnelshift_r(a, b, csh);
            else if (r >= 2 && r <= 10)
                aux[r] = pack4_shr5(o);
        }
        centre8(pk, jw);
        if (ch) {
            uint32_t hw[8];
            vfilter8(aux, hw);
#pragma unroll
            for (int r = 0; r < 8; r++) out[r] = avg4_u8(jw[r], hw[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) out[r] = avg4_u8(jw[r], fy == 2 ? jw[r] : (dy ? aux[r + 3] : aux[r + 2]));
        }
    }
}
// any phase (the table shims in blockops.cu use it for partitions at least 8 rows tall)
__device__ __forceinline__ void mc_luma_4x8(const uint8_t *__restrict__ src, int stride, int fx, int fy, uint32_t out[8])
{
    mc_luma_strip(src, stride, fx, fy, mc_class(fx, fy), out);
}

constexpr int kYPitch = 16 * kTileW + 8, kCPitch = 8 * kTileW + 8;

#ifdef P264B200_DEFINE_KERNELS
// Work lists of a tile, in the order the prediction pass walks them (the longest bodies first).  Luma: one list per
// interpolation class holding 4x8 strips and -- flagged "half" -- the 4x4 blocks of 8x4 / 4x4 sub-partitions, which run the
// strip body and keep its first four rows (12 % of the samples of the bench workload; one body instead of two per class);
// chroma: quadrants (both planes) with per-cell vectors / one vector.
enum { kLsCH = 0, kLsCB, kLsDiag, kLsV, kLsH, kLsChromaCell, kLsChromaOne, kLsCopy, kLsCount };
template <int TH>
struct InterSmem {
    static constexpr int kMbs = kTileW * TH;
    static constexpr int kLumaCap = 16 * kMbs, kQuadCap = 4 * kMbs;
    static constexpr int kResCap = 24 * kMbs;
    p264b200_mb mb[kMbs];                           // the tile's macroblock records
    // picture tile.  Rows are padded by 8 bytes: with a pitch of exactly 32 (16) banks every row of a 4x4 block falls
    // into the same bank, so the 16 blocks of one macroblock were a 4-way conflict on every tile access
    uint8_t y[16 * TH][kYPitch];                    // luma
    uint8_t c[2][8 * TH][kCPitch];                  // Cb / Cr
    // Work lists, compact: the six luma lists lie back to back in `luma` (counted first, placed after a barrier), because
    // every kilobyte of shared memory is a kilobyte less L1 for the reference windows (worst-case-sized lists: 13 KB per CTA)
    uint16_t luma[kLumaCap];                        // luma item = mb << 5 | b << 1 | half
    uint16_t chroma[2][kQuadCap];                   // [per-cell vectors, one vector]: quadrant = mb << 2 | q
    uint16_t res[kResCap];                          // residual work: full blocks (luma: mb << 4 | b, chroma: 16 * kMbs + (mb << 3 | cb)) from the
                                                    //   front, DC-only chroma blocks from the back
    const uint8_t *ref[kMaxRefs][3];
    int cnt[kLsCount];                              // items per list
    int nres[2];                                    // full / DC-only residual blocks
    int ticket;                                     // next chunk of the prediction pass
};
// storage order of the luma lists inside InterSmem::luma = processing order with the copy list last
__device__ __forceinline__ int luma_slot(int li) { return li == kLsCopy ? 5 : li; }
__device__ __forceinline__ int luma_list(int cls) { return cls == kMcCopy ? kLsCopy : kMcCentreH - cls; }
__device__ __forceinline__ int list_class(int li) { return li == kLsCopy ? kMcCopy : kMcCentreH - li; }

// one luma item: strip or 4x4 block (half) of interpolation class cls (warp-uniform)
template <class SM>
__device__ __forceinline__ void predict_luma(SM &sm, const Geometry &g, int mbx0, int mby0, int e, int cls)
{
    const int mb = e >> 5, b = (e >> 1) & 15;
    const bool half = e & 1;
    const p264b200_mb &m = sm.mb[mb];
    const int lx = 16 * (mb & (kTileW - 1)) + 4 * (b & 3), ly = 16 * (mb / kTileW) + 4 * (b >> 2);  // position inside the tile
    const int mvx = m.mv[b][0], mvy = m.mv[b][1];
    // integer position, clamped so the window (plus word alignment slack) stays inside the 32-sample border;
    // beyond the clamp every tap sees replicated edge samples anyway
    const int x0 = clip3i(16 * mbx0 + lx + (mvx >> 2), -16, g.width + 8);
    const int y0 = clip3i(16 * mby0 + ly + (mvy >> 2), -16, g.height + 8);
    uint32_t px[8];
    mc_luma_strip(sm.ref[mb_ref8(m, b)][0] + (ptrdiff_t)y0 * g.y_stride + x0, g.y_stride, mvx & 3, mvy & 3, cls, px);
#pragma unroll
    for (int r = 0; r < 8; r++)
        if (r < 4 || !half) *reinterpret_cast<uint32_t *>(&sm.y[ly + r][lx]) = px[r];
}

// one 8x8 luma quadrant's motion = one 4x4 chroma block of each plane
template <class SM>
__device__ __forceinline__ void predict_chroma(SM &sm, const Geometry &g, int mbx0, int mby0, int e, bool one)
{
    const int mb = e >> 2, i = e & 3;
    const p264b200_mb &m = sm.mb[mb];
    const int cx = i & 1, cy = i >> 1;
    const int lx = 8 * (mb & (kTileW - 1)) + 4 * cx, ly = 8 * (mb / kTileW) + 4 * cy;
    const int lb0 = 8 * cy + 2 * cx;  // top-left luma 4x4 block of the quadrant
    const int ref = m.ref[2 * cy + cx];
    if (one) {
        const int mvx = m.mv[lb0][0], mvy = m.mv[lb0][1];
        const int x0 = clip3i(8 * mbx0 + lx + (mvx >> 3), -8, g.width / 2 + 4);
        const int y0 = clip3i(8 * mby0 + ly + (mvy >> 3), -8, g.height / 2 + 4);
        const ptrdiff_t off = (ptrdiff_t)y0 * g.c_stride + x0;
#pragma unroll 1
        for (int plane = 0; plane < 2; plane++) {
            uint32_t px[4];
            mc_chroma_4x4(sm.ref[ref][1 + plane] + off, g.c_stride, mvx & 7, mvy & 7, px);
#pragma unroll
            for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(&sm.c[plane][ly + r][lx]) = px[r];
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < 4; s++) {
            // 2x2 cell s of this chroma block <-> luma 4x4 block lb0 + (s&1) + 4*(s>>1)
            const int lb = lb0 + (s & 1) + 4 * (s >> 1);
            const int mvx = m.mv[lb][0], mvy = m.mv[lb][1];
            const int x0 = clip3i(8 * mbx0 + lx + 2 * (s & 1) + (mvx >> 3), -8, g.width / 2 + 4);
            const int y0 = clip3i(8 * mby0 + ly + 2 * (s >> 1) + (mvy >> 3), -8, g.height / 2 + 4);
            const ptrdiff_t off = (ptrdiff_t)y0 * g.c_stride + x0;
#pragma unroll
            for (int plane = 0; plane < 2; plane++) {
                int o[4];
                mc_chroma_2x2(sm.ref[ref][1 + plane] + off, g.c_stride, mvx & 7, mvy & 7, o);
                uint8_t *t = &sm.c[plane][ly + 2 * (s >> 1)][lx + 2 * (s & 1)];
                *reinterpret_cast<uint16_t *>(t) = (uint16_t)(o[0] | (o[1] << 8));
                *reinterpret_cast<uint16_t *>(t + kCPitch) = (uint16_t)(o[2] | (o[3] << 8));
            }
        }
    }
}

// THREADS per CTA / MINB = CTAs per SM the register budget is set for (engine knob P264B200_INTER_VARIANT, see engine.cu)
template <int THREADS, int MINB, int TH>
__global__ void __launch_bounds__(THREADS, MINB) recon_inter_kernel(const FrameDesc *__restrict__ descs, Geometry g, int dbg)
{
    // dbg (engine knob P264B200_DBG, timing experiments only -- the pictures are wrong when set): bit 0 no prediction, bit 2 no residual
    extern __shared__ __align__(16) uint8_t inter_smem_raw[];   // sizeof(InterSmem) > 48 KB: dynamic, opted in by the engine
    typedef InterSmem<TH> SM;
    constexpr int kMbs = SM::kMbs;
    SM &sm = *reinterpret_cast<SM *>(inter_smem_raw);
    const FrameDesc &fd = descs[blockIdx.z];
    if (fd.slice_type != P264B200_SLICE_P) return;
    const int tid = threadIdx.x;
    const int mbx0 = blockIdx.x * kTileW, mby0 = blockIdx.y * TH;

    // ---- stage the tile's macroblock records (6 x 16 bytes each); outside the picture = "intra" = skipped
    for (int i = tid; i < kMbs * 6; i += THREADS) {
        const int ly = i / (6 * kTileW), rest = i - ly * (6 * kTileW);  // one tile row = 8 consecutive records
        const int mbx = mbx0 + rest / 6, mby = mby0 + ly;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (mbx < g.mb_w && mby < g.mb_h) v = __ldg(reinterpret_cast<const uint4 *>(fd.mbs + (size_t)mby * g.mb_w + mbx0) + rest);
        reinterpret_cast<uint4 *>(sm.mb)[i] = v;
    }
    static_assert(3 * kMaxRefs + kLsCount + 1 <= THREADS && (4 * kMbs) % 32 == 0, "setup assumes one pass, classification whole warps");
    if (tid < 3 * kMaxRefs) {
        const int r = tid / 3;
        sm.ref[r][tid - 3 * r] = r < fd.num_ref ? fd.ref[r][tid - 3 * r] : nullptr;
    } else if (tid < 3 * kMaxRefs + kLsCount) {
        sm.cnt[tid - 3 * kMaxRefs] = 0;
    } else if (tid == 3 * kMaxRefs + kLsCount) {
        sm.nres[0] = sm.nres[1] = 0;
        sm.ticket = THREADS / 32;   // the first chunk of every warp is its own index
    }
    __syncthreads();

    // ---- classify, one thread per 8x8 quadrant: strips / blocks by interpolation class, chroma by "one vector for the
    // quadrant", and the blocks that carry residual (one position range per quadrant from a warp scan + one atomic per warp).
    const int lane = tid & 31;
    static_assert(4 * kMbs <= THREADS, "one quadrant per thread: its luma items wait in registers for the placement pass");
    uint32_t slot[4] = {0, 0, 0, 0};
    if (tid < 4 * kMbs) {
        const int qi = tid, mb = qi >> 2, q = qi & 3;
        const p264b200_mb &m = sm.mb[mb];
        const bool inter = !P264B200_IS_INTRA(m.mb_type);
        const int lb0 = 8 * (q >> 1) + 2 * (q & 1);
        unsigned lm = 0, cfull = 0, cdc = 0;   // luma blocks (bits 0, 1, 4, 5 = lb0, lb0 + 1, lb0 + 4, lb0 + 5) / chroma planes with full / DC-only residual
        if (inter) {
            if (m.mb_type != P264B200_MB_P_SKIP) {
                // pull this macroblock's coefficient chunk towards the SM while the prediction pass runs (the residual pass
                // would otherwise wait out the HBM latency behind the barrier): lines q, q + 4, ... of the chunk
                const int n16 = 16 * (__popc(m.luma_mask) + __popc(m.chroma_mask)) + (m.cbp_chroma ? 8 : 0);
                const char *cp = reinterpret_cast<const char *>(fd.coefs + m.coef_off);
                for (int o = 128 * q; o < 2 * n16; o += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + o));
            }
            const int2 v0 = *reinterpret_cast<const int2 *>(m.mv[lb0]), v1 = *reinterpret_cast<const int2 *>(m.mv[lb0 + 4]);
            const bool one = v0.x == v0.y && v0.x == v1.x && v0.x == v1.y;
            // luma items are COUNTED here and placed after the barrier (slot = 1 << 31 | list << 27 | position << 16 | item)
            if (one) {
                // one vector for the quadrant (every partition of 8x8 and up): its two strips sit side by side in the list, so
                // that the lanes of a 16-wide partition read the same cache lines in the same load instruction
                const int l0 = luma_list(mc_class(v0.x & 3, (v0.x >> 16) & 3));
                const uint32_t pos = (uint32_t)atomicAdd(&sm.cnt[l0], 2);
                slot[0] = 0x80000000u | (uint32_t)l0 << 27 | pos << 16 | (uint32_t)(mb << 5 | lb0 << 1);
                slot[1] = slot[0] + (1u << 16) + 2u;
            } else {
#pragma unroll
                for (int sx = 0; sx < 2; sx++) {
                    const int top = sx ? v0.y : v0.x, bot = sx ? v1.y : v1.x;
                    const int l0 = luma_list(mc_class(top & 3, (top >> 16) & 3));
                    const uint32_t e0 = (uint32_t)(mb << 5 | (lb0 + sx) << 1);
                    if (top == bot) {
                        slot[2 * sx] = 0x80000000u | (uint32_t)l0 << 27 | (uint32_t)atomicAdd(&sm.cnt[l0], 1) << 16 | e0;
                    } else {
                        const int l1 = luma_list(mc_class(bot & 3, (bot >> 16) & 3));
                        slot[2 * sx] = 0x80000000u | (uint32_t)l0 << 27 | (uint32_t)atomicAdd(&sm.cnt[l0], 1) << 16 | e0 | 1u;
                        slot[2 * sx + 1] = 0x80000000u | (uint32_t)l1 << 27 | (uint32_t)atomicAdd(&sm.cnt[l1], 1) << 16 | (e0 + (4u << 1)) | 1u;
                    }
                }
            }
            const int lc = one ? kLsChromaOne : kLsChromaCell;
            sm.chroma[lc - kLsChromaCell][atomicAdd(&sm.cnt[lc], 1)] = (uint16_t)qi;
            lm = (m.luma_mask >> lb0) & 0x33u;
            if (m.cbp_chroma) {
                cfull = ((m.chroma_mask >> q) & 1u) | (((m.chroma_mask >> (4 + q)) & 1u) << 1);
                cdc = cfull ^ 3u;
            }
        }
        // residual work lists: full blocks from the front, DC-only chroma blocks from the back
        const int mine = (__popc(lm) + __popc(cfull)) | (__popc(cdc) << 16);
        int scan = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, scan, d);
            if (lane >= d) scan += t;
        }
        int base = 0;
        if (lane == 31) base = atomicAdd(&sm.nres[0], scan & 0xffff) | (atomicAdd(&sm.nres[1], scan >> 16) << 16);
        base = __shfl_sync(0xffffffffu, base, 31) + scan - mine;
        int pf = base & 0xffff, pd = SM::kResCap - 1 - (base >> 16);
        while (lm) {
            const int k = __ffs(lm) - 1;
            lm &= lm - 1;
            sm.res[pf++] = (uint16_t)(mb << 4 | (lb0 + k));
        }
#pragma unroll
        for (int plane = 0; plane < 2; plane++) {
            const int cb = 4 * plane + q;
            if (cfull >> plane & 1) sm.res[pf++] = (uint16_t)(16 * kMbs + (mb << 3 | cb));
            if (cdc >> plane & 1) sm.res[pd--] = (uint16_t)(mb << 3 | cb);
        }
    }
    __syncthreads();

    // ---- placement: the luma lists back to back in storage order (luma_slot)
    if (tid < 4 * kMbs) {
        int off[6];
        off[0] = 0;
#pragma unroll
        for (int k = 1; k < 6; k++) off[k] = off[k - 1] + sm.cnt[k - 1 == 5 ? kLsCopy : k - 1];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (slot[k] >> 31) {
                const int sl = luma_slot((slot[k] >> 27) & 7);
                int o = 0;
#pragma unroll
                for (int j = 1; j < 6; j++) o = sl == j ? off[j] : o;
                sm.luma[o + ((slot[k] >> 16) & 0x7ff)] = (uint16_t)(slot[k] & 0xffff);
            }
    }
    __syncthreads();

    // ---- prediction: the warps draw class-pure chunks of 32 items by ticket, heaviest lists first.  Lanes 0..7 hold the
    // chunk range [start, end) of list `lane`; a ticket's list is the number of lists that end at or before it.
    if (!(dbg & 1)) {
        const int n_l = lane < kLsCount ? sm.cnt[lane] : 0;
        // where list `lane` starts inside its storage array
        int soff = 0;
        if (lane < kLsCount && lane != kLsChromaCell && lane != kLsChromaOne) {
            const int sl = luma_slot(lane);
#pragma unroll
            for (int k = 0; k < 5; k++) soff += k < sl ? sm.cnt[k] : 0;
        }
        int end = (n_l + 31) >> 5;
        const int nch = end;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, end, d);
            if (lane >= d) end += t;
        }
        const int start = end - nch;
        int ticket = tid >> 5;
        for (;;) {
            const int li = __popc(__ballot_sync(0xffffffffu, lane < kLsCount && end <= ticket));
            if (li >= kLsCount) break;
            int t = 0;
            if (lane == 0) t = atomicAdd(&sm.ticket, 1);   // drawn ahead: the atomic's latency hides behind the chunk
            const int idx = 32 * (ticket - __shfl_sync(0xffffffffu, start, li)) + lane;
            const int so = __shfl_sync(0xffffffffu, soff, li);
            if (idx < __shfl_sync(0xffffffffu, n_l, li)) {
                if (li == kLsChromaCell || li == kLsChromaOne)
                    predict_chroma(sm, g, mbx0, mby0, sm.chroma[li - kLsChromaCell][idx], li == kLsChromaOne);
                else
                    predict_luma(sm, g, mbx0, mby0, sm.luma[so + idx], list_class(li));
            }
            ticket = __shfl_sync(0xffffffffu, t, 0);
        }
    }
    __syncthreads();

    // ---- residual on the tile: blocks with coefficients get dequant + inverse transform ...
    const int nres = (dbg & 4) ? 0 : sm.nres[0], ndc = (dbg & 4) ? 0 : sm.nres[1];
#pragma unroll 1
    for (int idx = tid; idx < nres; idx += THREADS) {
        const int k = sm.res[idx];
        const bool chroma = k >= 16 * kMbs;
        const int q = k & (16 * kMbs - 1);
        const int mb = chroma ? q >> 3 : q >> 4;
        const p264b200_mb &m = sm.mb[mb];
        const int tmx = mb & (kTileW - 1), tmy = mb / kTileW;
        uint8_t *t;
        int pitch, qp, dcv = 0;
        const int16_t *lvl;
        if (!chroma) {
            const int b = q & 15;
            t = &sm.y[16 * tmy + 4 * (b >> 2)][16 * tmx + 4 * (b & 3)];
            pitch = kYPitch;
            qp = m.qp;
            lvl = fd.coefs + m.coef_off + 16 * __popc(m.luma_mask & ((1u << b) - 1));
        } else {
            const int cb = q & 7, plane = cb >> 2, i = cb & 3;
            t = &sm.c[plane][8 * tmy + 4 * (i >> 1)][8 * tmx + 4 * (i & 1)];
            pitch = kCPitch;
            qp = c_chroma_qp[clip3i(m.qp + fd.chroma_qp_off, 0, 51)];
            const int16_t *cf = fd.coefs + m.coef_off + 16 * __popc(m.luma_mask);
            int dc[4];
            chroma_dc(cf + 4 * plane, qp, dc);
            dcv = i == 0 ? dc[0] : i == 1 ? dc[1] : i == 2 ? dc[2] : dc[3];
            lvl = cf + 8 + 16 * __popc(m.chroma_mask & ((1u << cb) - 1));
        }
        uint32_t px[4];
#pragma unroll
        for (int r = 0; r < 4; r++) px[r] = *reinterpret_cast<const uint32_t *>(t + r * pitch);
        residual4x4(lvl, qp, chroma, dcv, px);
#pragma unroll
        for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(t + r * pitch) = px[r];
    }
    // ---- ... chroma blocks with a DC term only add the constant (dc + 32) >> 6 (what add4x4_idct makes of a lone DC)
#pragma unroll 1
    for (int idx = tid; idx < ndc; idx += THREADS) {
        const int q = sm.res[SM::kResCap - 1 - idx];
        const int mb = q >> 3, cb = q & 7, plane = cb >> 2, i = cb & 3;
        const p264b200_mb &m = sm.mb[mb];
        uint8_t *t = &sm.c[plane][8 * (mb / kTileW) + 4 * (i >> 1)][8 * (mb & (kTileW - 1)) + 4 * (i & 1)];
        const int qpc = c_chroma_qp[clip3i(m.qp + fd.chroma_qp_off, 0, 51)];
        int dc[4];
        chroma_dc(fd.coefs + m.coef_off + 16 * __popc(m.luma_mask) + 4 * plane, qpc, dc);
        const int r0 = ((i == 0 ? dc[0] : i == 1 ? dc[1] : i == 2 ? dc[2] : dc[3]) + 32) >> 6;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t p = *reinterpret_cast<const uint32_t *>(t + r * kCPitch);
            *reinterpret_cast<uint32_t *>(t + r * kCPitch) =
                pack4_sat_u8((int)(p & 0xff) + r0, (int)((p >> 8) & 0xff) + r0, (int)((p >> 16) & 0xff) + r0, (int)(p >> 24) + r0);
        }
    }
    __syncthreads();

    // ---- the tile leaves with coalesced stores; intra / outside macroblocks are not ours
#pragma unroll
    for (int i = tid; i < 16 * TH * kTileW; i += THREADS) {
        const int row = i >> 3, seg = i & 7;  // 128 rows x 8 macroblock-wide segments
        const int mb = (row >> 4) * kTileW + seg;
        if (!P264B200_IS_INTRA(sm.mb[mb].mb_type)) {
            const uint2 a = *reinterpret_cast<const uint2 *>(&sm.y[row][16 * seg]), b = *reinterpret_cast<const uint2 *>(&sm.y[row][16 * seg + 8]);
            *reinterpret_cast<uint4 *>(fd.cur[0] + (ptrdiff_t)(16 * mby0 + row) * g.y_stride + 16 * (mbx0 + seg)) = make_uint4(a.x, a.y, b.x, b.y);
        }
    }
#pragma unroll
    for (int i = tid; i < 2 * 8 * TH * kTileW; i += THREADS) {
        const int plane = i / (8 * TH * kTileW), row = (i >> 3) & (8 * TH - 1), seg = i & 7;  // 2 planes x 8 TH rows x 8 segments of 8 samples
        const int mb = (row >> 3) * kTileW + seg;
        if (!P264B200_IS_INTRA(sm.mb[mb].mb_type))
            *reinterpret_cast<uint2 *>(fd.cur[1 + plane] + (ptrdiff_t)(8 * mby0 + row) * g.c_stride + 8 * (mbx0 + seg)) =
                *reinterpret_cast<const uint2 *>(&sm.c[plane][row][8 * seg]);
    }
}
#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
