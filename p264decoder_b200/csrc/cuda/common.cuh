// Shared device-side definitions of the reconstruction engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/p264b200_recon.h"

namespace p264b200 {

constexpr int kLumaPad = 32;    // same envelope as core/frame.c:42,62-63
constexpr int kChromaPad = 16;
constexpr int kMaxRefs = 16;

// Geometry of the padded frame store (identical for every lane / slot of an engine)
struct Geometry {
    int mb_w, mb_h;
    int width, height;          // 16*mb_w, 16*mb_h
    int y_stride, c_stride;     // bytes per row, multiples of 128 / 64
    int y_rows, c_rows;         // padded rows
    size_t y_plane, c_plane;    // bytes per padded plane
};

// One picture of one lane, as the kernels see it (built on the host per step, lives in HBM)
struct FrameDesc {
    const p264b200_mb *mbs;
    const int16_t *coefs;
    uint8_t *cur[3];                    // plane origins (sample 0,0) of the picture being reconstructed
    const uint8_t *ref[kMaxRefs][3];    // list-0 reference plane origins
    int *row_progress;                  // [3][mb_h]: (unused), luma deblock, chroma deblock wavefronts
    struct DeblockSide *dbf_bs;         // per-MB boundary strengths (deblock_bs_kernel -> deblock_kernel)
    int *intra_work;                    // [1 + 2*n_mb]: run count, run starts, per-MB done epochs (recon_intra.cuh)
    int slice_type, deblock, alpha_off, beta_off, chroma_qp_off, n_intra, num_ref, pad_;
};

// ---- tables (core/set.c:27-35, core/macroblock.h:210-218, core/frame.c:262-291) -------------
static __constant__ uint8_t c_dq_scale[6][3] = {{10, 13, 16}, {11, 14, 18}, {13, 16, 20},
                                         {14, 18, 23}, {16, 20, 25}, {18, 23, 29}};
static __constant__ uint8_t c_chroma_qp[52] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                                        18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33,
                                        34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
static __constant__ uint8_t c_alpha[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,   0,   0,   4,   4,
                                    5,  6,  7,  8,  9,  10, 12, 13, 15, 17, 20, 22, 25, 28,  32,  36,  40,  45,
                                    50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
static __constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  0,  2,  2,
                                   2, 3, 3, 3, 3, 4, 4, 4, 6,  6,  7,  7,  8,  8,  9,  9,  10, 10,
                                   11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
static __constant__ __align__(16) uint8_t c_tc0[52][4] = {
    {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 0, 0},   {0, 0, 0, 0},   {0, 0, 0, 0},   {0, 0, 0, 0},
    {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 0, 0},   {0, 0, 0, 0},   {0, 0, 0, 0},   {0, 0, 0, 0},
    {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 0, 0},  {0, 0, 1, 0},   {0, 0, 1, 0},   {0, 0, 1, 0},   {0, 0, 1, 0},
    {0, 1, 1, 0},  {0, 1, 1, 0},  {1, 1, 1, 0},  {1, 1, 1, 0},   {1, 1, 1, 0},   {1, 1, 1, 0},   {1, 1, 2, 0},
    {1, 1, 2, 0},  {1, 1, 2, 0},  {1, 1, 2, 0},  {1, 2, 3, 0},   {1, 2, 3, 0},   {2, 2, 3, 0},   {2, 2, 4, 0},
    {2, 3, 4, 0},  {2, 3, 4, 0},  {3, 3, 5, 0},  {3, 4, 6, 0},   {3, 4, 6, 0},   {4, 5, 7, 0},   {4, 5, 8, 0},
    {4, 6, 9, 0},  {5, 7, 10, 0}, {6, 8, 11, 0}, {6, 8, 13, 0},  {7, 10, 14, 0}, {8, 11, 16, 0}, {9, 12, 18, 0},
    {10, 13, 20, 0}, {11, 15, 23, 0}, {13, 17, 25, 0}};
// zig-zag scan position -> raster index 4*y+x (decoder/macroblock.c:602-603)
static __constant__ uint8_t c_zz[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};

__device__ __forceinline__ int clip3i(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip8i(int v) { return min(max(v, 0), 255); }

// clamp four ints to 0..255 and pack them, v0 in the low byte: two I2IP (cvt.pack.sat) instructions
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3)
{
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
    return d;
}
// per-byte rounded average (a + b + 1) >> 1 of four packed u8
__device__ __forceinline__ uint32_t avg4_u8(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu); }

__device__ __forceinline__ int mb_ref8(const p264b200_mb &m, int b) { return m.ref[(b >> 3) * 2 + ((b & 3) >> 1)]; }

// ------------------------------------------------------------------ coefficient path
// unscan (decoder/macroblock.c:605-622) + dequant_4x4 (core/quant.c:72-99, flat lists) of one block.
// lvl: 16 levels in zig-zag order; d: raster, int16 wrap-around preserved.
__device__ __forceinline__ void unscan_dequant(const int16_t *__restrict__ lvl, int qp, int d[16])
{
    const int4 lo = *reinterpret_cast<const int4 *>(lvl), hi = *reinterpret_cast<const int4 *>(lvl + 8);
    int v[16];
    v[0] = (short)(lo.x & 0xffff), v[1] = lo.x >> 16, v[2] = (short)(lo.y & 0xffff), v[3] = lo.y >> 16;
    v[4] = (short)(lo.z & 0xffff), v[5] = lo.z >> 16, v[6] = (short)(lo.w & 0xffff), v[7] = lo.w >> 16;
    v[8] = (short)(hi.x & 0xffff), v[9] = hi.x >> 16, v[10] = (short)(hi.y & 0xffff), v[11] = hi.y >> 16;
    v[12] = (short)(hi.z & 0xffff), v[13] = hi.z >> 16, v[14] = (short)(hi.w & 0xffff), v[15] = hi.w >> 16;
    const int rem = qp % 6, qbits = qp / 6 - 4;
    const int s0 = 16 * c_dq_scale[rem][0], s1 = 16 * c_dq_scale[rem][1], s2 = 16 * c_dq_scale[rem][2];
    // raster positions in zig-zag order: {0,1,4,8,5,2,3,6,9,12,13,10,7,11,14,15}
    const int zz[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
    if (qbits >= 0) {
        // qp >= 24 (uniform over a picture in practice): the left shift folds into the multiplier, the int16
        // truncation sees the same low 16 bits
        const int m0 = (int)((unsigned)s0 << qbits), m1 = (int)((unsigned)s1 << qbits), m2 = (int)((unsigned)s2 << qbits);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int r = zz[i], x = r & 3, y = r >> 2;
            const int mf = ((x & 1) + (y & 1)) == 0 ? m0 : (((x & 1) + (y & 1)) == 1 ? m1 : m2);
            d[r] = (short)(v[i] * mf);
        }
    } else {
        const int sh = -qbits, rnd = 1 << (sh - 1);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int r = zz[i], x = r & 3, y = r >> 2;
            const int mf = ((x & 1) + (y & 1)) == 0 ? s0 : (((x & 1) + (y & 1)) == 1 ? s1 : s2);
            d[r] = (short)((v[i] * mf + rnd) >> sh);
        }
    }
}

// 4x4 inverse core transform of add4x4_idct (core/dct.c:205-247): rows then columns, int16
// truncation at the same points; r = residual samples before the add.
__device__ __forceinline__ void idct4x4_core(const int d[16], int r[16])
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s02 = d[i * 4 + 0] + d[i * 4 + 2], d02 = d[i * 4 + 0] - d[i * 4 + 2];
        const int s13 = d[i * 4 + 1] + (d[i * 4 + 3] >> 1), d13 = (d[i * 4 + 1] >> 1) - d[i * 4 + 3];
        t[i * 4 + 0] = (short)(s02 + s13);
        t[i * 4 + 1] = (short)(d02 + d13);
        t[i * 4 + 2] = (short)(d02 - d13);
        t[i * 4 + 3] = (short)(s02 - s13);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s02 = t[0 * 4 + i] + t[2 * 4 + i], d02 = t[0 * 4 + i] - t[2 * 4 + i];
        const int s13 = t[1 * 4 + i] + (t[3 * 4 + i] >> 1), d13 = (t[1 * 4 + i] >> 1) - t[3 * 4 + i];
        r[0 * 4 + i] = (short)((s02 + s13 + 32) >> 6);
        r[1 * 4 + i] = (short)((d02 + d13 + 32) >> 6);
        r[2 * 4 + i] = (short)((d02 - d13 + 32) >> 6);
        r[3 * 4 + i] = (short)((s02 - s13 + 32) >> 6);
    }
}
// pred/out: 4 rows of 4 packed u8 samples; dst = clip_uint8(dst + d) (core/dct.c:243)
__device__ __forceinline__ void idct4x4_add(const int d[16], uint32_t px[4])
{
    int r[16];
    idct4x4_core(d, r);
#pragma unroll
    for (int y = 0; y < 4; y++)
        px[y] = pack4_sat_u8((int)(px[y] & 0xff) + r[y * 4 + 0], (int)((px[y] >> 8) & 0xff) + r[y * 4 + 1],
                             (int)((px[y] >> 16) & 0xff) + r[y * 4 + 2], (int)(px[y] >> 24) + r[y * 4 + 3]);
}

// chroma DC: dct2x2dc (core/dct.c:55-68) then p264_mb_dequant_2x2_dc (core/quant.c:138-159)
__device__ __forceinline__ void chroma_dc(const int16_t *__restrict__ lvl, int qpc, int dc[4])
{
    const int a = lvl[0], b = lvl[1], c = lvl[2], e = lvl[3];
    const int t00 = a + b, t10 = a - b, t01 = c + e, t11 = c - e;
    int v[4] = {(short)(t00 + t01), (short)(t10 + t11), (short)(t00 - t01), (short)(t10 - t11)};
    const int qbits = qpc / 6 - 5, mf = 16 * c_dq_scale[qpc % 6][0];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int t = qbits >= 0 ? v[i] * (int)((unsigned)mf << qbits) : ((v[i] * mf) >> (-qbits));
        dc[i] = (short)t;
    }
}

// luma DC of I16x16: idct4x4dc (core/dct.c:104-136) then p264_mb_dequant_4x4_dc (core/quant.c:161-191)
__device__ __forceinline__ void luma_dc(const int16_t *__restrict__ lvl, int qp, int dc[16])
{
    int d[16], t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = 0;
    const int zz[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
#pragma unroll
    for (int i = 0; i < 16; i++) d[zz[i]] = lvl[i];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s01 = d[0 + i] + d[4 + i], d01 = d[0 + i] - d[4 + i];
        const int s23 = d[8 + i] + d[12 + i], d23 = d[8 + i] - d[12 + i];
        t[0 + i] = (short)(s01 + s23);
        t[4 + i] = (short)(s01 - s23);
        t[8 + i] = (short)(d01 - d23);
        t[12 + i] = (short)(d01 + d23);
    }
    const int qbits = qp / 6 - 6, mf = 16 * c_dq_scale[qp % 6][0];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s01 = t[i * 4 + 0] + t[i * 4 + 1], d01 = t[i * 4 + 0] - t[i * 4 + 1];
        const int s23 = t[i * 4 + 2] + t[i * 4 + 3], d23 = t[i * 4 + 2] - t[i * 4 + 3];
        int o[4] = {(short)(s01 + s23), (short)(s01 - s23), (short)(d01 - d23), (short)(d01 + d23)};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = qbits >= 0 ? o[k] * (int)((unsigned)mf << qbits) : ((o[k] * mf + (1 << (-qbits - 1))) >> (-qbits));
            dc[i * 4 + k] = (short)v;
        }
    }
}

// residual of one 4x4 block onto px: lvl == nullptr means "all AC zero"; dc_splice replaces d[0]
__device__ __forceinline__ void residual4x4(const int16_t *__restrict__ lvl, int qp, bool has_dc, int dcv, uint32_t px[4])
{
    int d[16];
    if (lvl)
        unscan_dequant(lvl, qp, d);
    else {
#pragma unroll
        for (int i = 0; i < 16; i++) d[i] = 0;
    }
    if (has_dc) d[0] = dcv;
    idct4x4_add(d, px);
}

// prefetch loads that must be ISSUED where they are written (the compiler is free to sink an ordinary
// read-only load down to its first use, which puts the whole memory latency back into the dependent chain)
__device__ __forceinline__ uint4 ldg_now_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_now_v2(const void *p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void ldg_now(const void *p, uint4 &v) { v = ldg_now_v4(p); }
__device__ __forceinline__ void ldg_now(const void *p, uint2 &v) { v = ldg_now_v2(p); }
__device__ __forceinline__ uint32_t ldg_now_u32(const void *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wavefront flag helpers (global memory, visible across SMs)
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace p264b200
