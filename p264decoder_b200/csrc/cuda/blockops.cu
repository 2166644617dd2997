// Single-block device entry points behind the reference's function-pointer tables
// (include/p264_b200_tables.h).  Each shim stages the caller's block (plus exactly the
// neighbouring samples the reference routine would read) into a small device scratch tile,
// runs the SAME device functions the batched frame kernels inline (for the deblocking slots: the packed
// two-lines-per-register filters of swar.cuh, not the scalar forms), and copies the result back.
// A test surface: one launch + two copies per call; never used by the frame path.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../../include/p264_b200_tables.h"
#include "common.cuh"
#include "deblock.cuh"
#include "recon_inter.cuh"
#include "recon_intra.cuh"

using namespace p264b200;

namespace {

constexpr int TS = 64;            // scratch tile stride
constexpr int TO = 16 * TS + 16;  // tile origin (sample 0,0)

struct Scratch {
    uint8_t pix[TS * TS];
    uint8_t pix2[TS * TS];
    int16_t coef[256];
    int tab[6 * 64];
    int args[16];
    int result[4];
};

enum Op {
    OP_ADD_IDCT4 = 1,  // args: nblk (1,4,16)
    OP_ADD_IDCT8,      // args: nblk (1,4)
    OP_IDCT4DC,
    OP_IDCT2DC,
    OP_DEQUANT4,  // args: qp
    OP_DEQUANT8,
    OP_DEQUANT4DC,
    OP_DEQUANT2DC,
    OP_MC_LUMA,    // args: w,h,fx,fy
    OP_MC_CHROMA,  // args: w,h,dx,dy
    OP_AVG,        // args: w,h
    OP_AVG_WEIGHT, // args: w,h,weight
    OP_PRED16,     // args: mode
    OP_PRED8C,
    OP_PRED4,
    OP_DBF_LUMA,   // args: dir(0 = filter across columns "h", 1 = across rows "v"), alpha, beta, tc0[4] or intra flag
    OP_DBF_CHROMA,
    OP_SSD,        // args: w,h
};

// add8x8_idct8 (core/dct.c:321-367): 8x8 inverse transform, dct[0][0] += 32, >> 6, int16 stores
__device__ void add_idct8(uint8_t *dst, int stride, int16_t *c)
{
    int t[64];
    for (int i = 0; i < 64; i++) t[i] = c[i];
    t[0] = (short)(t[0] + 32);
#define IDCT8_1D(S, D)                                                      \
    {                                                                       \
        const int a0 = S(0) + S(4), a2 = S(0) - S(4);                       \
        const int a4 = (S(2) >> 1) - S(6), a6 = (S(6) >> 1) + S(2);         \
        const int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;   \
        const int a1 = -S(3) + S(5) - S(7) - (S(7) >> 1);                   \
        const int a3 = S(1) + S(7) - S(3) - (S(3) >> 1);                    \
        const int a5 = -S(1) + S(7) + S(5) + (S(5) >> 1);                   \
        const int a7 = S(3) + S(5) + S(1) + (S(1) >> 1);                    \
        const int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2);                 \
        const int b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);                 \
        D(0, b0 + b7);                                                      \
        D(1, b2 + b5);                                                      \
        D(2, b4 + b3);                                                      \
        D(3, b6 + b1);                                                      \
        D(4, b6 - b1);                                                      \
        D(5, b4 - b3);                                                      \
        D(6, b2 - b5);                                                      \
        D(7, b0 - b7);                                                      \
    }
    for (int i = 0; i < 8; i++) {
#define S(x) t[i * 8 + x]
#define D(x, v) t[i * 8 + x] = (short)(v)
        IDCT8_1D(S, D)
#undef S
#undef D
    }
    for (int i = 0; i < 8; i++) {
#define S(x) t[x * 8 + i]
#define D(x, v) dst[i + x * stride] = (uint8_t)clip8i(dst[i + x * stride] + ((v) >> 6))
        IDCT8_1D(S, D)
#undef S
#undef D
    }
#undef IDCT8_1D
}

__global__ void blockop_kernel(Scratch *s, int op)
{
    const int t = threadIdx.x;
    uint8_t *p = s->pix + TO;
    uint8_t *q = s->pix2 + TO;
    const int *a = s->args;
    switch (op) {
    case OP_ADD_IDCT4: {
        const int n = a[0];
        if (t < n) {
            // block order of add8x8_idct / add16x16_idct (core/dct.c:249-263)
            const int x = (n == 1) ? 0 : 4 * ((t & 1) + 2 * ((t >> 2) & 1)), y = (n == 1) ? 0 : 4 * (((t >> 1) & 1) + 2 * (t >> 3));
            int d[16];
            for (int i = 0; i < 16; i++) d[i] = s->coef[t * 16 + i];
            uint32_t px[4];
            for (int r = 0; r < 4; r++) px[r] = *reinterpret_cast<uint32_t *>(p + (y + r) * TS + x);
            idct4x4_add(d, px);
            for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(p + (y + r) * TS + x) = px[r];
        }
        break;
    }
    case OP_ADD_IDCT8:
        if (t < a[0]) add_idct8(p + 8 * (t >> 1) * TS + 8 * (t & 1), TS, s->coef + 64 * t);
        break;
    case OP_IDCT4DC:
        if (t == 0) {
            // idct4x4dc alone (no dequant): reuse luma_dc's butterflies on raster input
            int d[16], u[16];
            for (int i = 0; i < 16; i++) d[i] = s->coef[i];
            for (int i = 0; i < 4; i++) {
                const int s01 = d[i] + d[4 + i], d01 = d[i] - d[4 + i], s23 = d[8 + i] + d[12 + i], d23 = d[8 + i] - d[12 + i];
                u[i] = (short)(s01 + s23), u[4 + i] = (short)(s01 - s23), u[8 + i] = (short)(d01 - d23), u[12 + i] = (short)(d01 + d23);
            }
            for (int i = 0; i < 4; i++) {
                const int s01 = u[i * 4] + u[i * 4 + 1], d01 = u[i * 4] - u[i * 4 + 1], s23 = u[i * 4 + 2] + u[i * 4 + 3], d23 = u[i * 4 + 2] - u[i * 4 + 3];
                s->coef[i * 4] = (short)(s01 + s23), s->coef[i * 4 + 1] = (short)(s01 - s23);
                s->coef[i * 4 + 2] = (short)(d01 - d23), s->coef[i * 4 + 3] = (short)(d01 + d23);
            }
        }
        break;
    case OP_IDCT2DC:
        if (t == 0) {
            const int c0 = s->coef[0], c1 = s->coef[1], c2 = s->coef[2], c3 = s->coef[3];
            const int t00 = c0 + c1, t10 = c0 - c1, t01 = c2 + c3, t11 = c2 - c3;
            s->coef[0] = (short)(t00 + t01), s->coef[1] = (short)(t10 + t11), s->coef[2] = (short)(t00 - t01), s->coef[3] = (short)(t10 - t11);
        }
        break;
    case OP_DEQUANT4:
    case OP_DEQUANT8: {
        // core/quant.c:72-136 with the CALLER's dequant_mf table
        const int n = op == OP_DEQUANT4 ? 16 : 64, qp = a[0], mf = qp % 6, qb = qp / 6 - (op == OP_DEQUANT4 ? 4 : 6);
        for (int i = t; i < n; i += 32) {
            int v = s->coef[i] * s->tab[mf * n + i];
            v = qb >= 0 ? (int)((unsigned)v << qb) : ((v + (1 << (-qb - 1))) >> (-qb));
            s->coef[i] = (short)v;
        }
        break;
    }
    case OP_DEQUANT4DC:
    case OP_DEQUANT2DC: {
        // core/quant.c:138-191
        const int n = op == OP_DEQUANT4DC ? 16 : 4, qp = a[0], qb = qp / 6 - (op == OP_DEQUANT4DC ? 6 : 5), m = s->tab[(qp % 6) * 16];
        if (t < n) {
            int v;
            if (qb >= 0)
                v = s->coef[t] * (int)((unsigned)m << qb);
            else if (op == OP_DEQUANT4DC)
                v = (s->coef[t] * m + (1 << (-qb - 1))) >> (-qb);
            else
                v = (s->coef[t] * m) >> (-qb);
            s->coef[t] = (short)v;
        }
        break;
    }
    case OP_MC_LUMA: {
        const int w4 = a[0] >> 2, h4 = a[1] >> 2;
        if ((h4 & 1) == 0) {
            // partitions at least 8 rows tall: the 4x8 strip bodies of the frame kernel
            if (t < w4 * (h4 >> 1)) {
                const int bx = 4 * (t % w4), by = 8 * (t / w4);
                uint32_t px[8];
                mc_luma_4x8(p + by * TS + bx, TS, a[2], a[3], px);
                for (int r = 0; r < 8; r++) *reinterpret_cast<uint32_t *>(q + (by + r) * TS + bx) = px[r];
            }
        } else if (t < w4 * h4) {
            const int bx = 4 * (t % w4), by = 4 * (t / w4);
            uint32_t px[4];
            mc_luma_4x4(p + by * TS + bx, TS, a[2], a[3], px);
            for (int r = 0; r < 4; r++) *reinterpret_cast<uint32_t *>(q + (by + r) * TS + bx) = px[r];
        }
        break;
    }
    case OP_MC_CHROMA: {
        const int w2 = a[0] >> 1, h2 = a[1] >> 1;
        if (t < w2 * h2) {
            const int bx = 2 * (t % w2), by = 2 * (t / w2);
            int o[4];
            mc_chroma_2x2(p + by * TS + bx, TS, a[2], a[3], o);
            q[by * TS + bx] = (uint8_t)o[0], q[by * TS + bx + 1] = (uint8_t)o[1];
            q[(by + 1) * TS + bx] = (uint8_t)o[2], q[(by + 1) * TS + bx + 1] = (uint8_t)o[3];
        }
        break;
    }
    case OP_AVG:  // core/mc.c:76-88
        for (int i = t; i < a[0] * a[1]; i += 32) {
            const int x = i % a[0], y = i / a[0];
            p[y * TS + x] = (uint8_t)((p[y * TS + x] + q[y * TS + x] + 1) >> 1);
        }
        break;
    case OP_AVG_WEIGHT:  // core/mc.c:110-135
        for (int i = t; i < a[0] * a[1]; i += 32) {
            const int x = i % a[0], y = i / a[0];
            p[y * TS + x] = (uint8_t)clip8i((p[y * TS + x] * a[2] + q[y * TS + x] * (64 - a[2]) + 32) >> 6);
        }
        break;
    case OP_PRED16:
    case OP_PRED8C: {
        const int N = op == OP_PRED16 ? 16 : 8, mode = a[0];
        // canonical numbering differs: 16x16 V0 H1 DC2 P3, chroma DC0 H1 V2 P3 (core/predict.h:30-54)
        const int m = op == OP_PRED16 ? mode : (mode == 0 ? 2 : mode == 2 ? 0 : mode);
        int i00 = 0, pb = 0, pc = 0;
        if (m == 3) {
            if (N == 16)
                plane_params<16>(p, TS, i00, pb, pc);
            else
                plane_params<8>(p, TS, i00, pb, pc);
        }
        uint8_t out[8];
        for (int k = 0; k < N * N / 32; k++) {
            const int i = t * (N * N / 32) + k, x = i % N, y = i / N;
            int v;
            if (m == 0)
                v = p[-TS + x];
            else if (m == 1)
                v = p[y * TS - 1];
            else if (m == 3)
                v = clip8i((i00 + pb * x + pc * y) >> 5);
            else if (N == 16) {
                int st = 0, sl = 0;
                for (int j = 0; j < 16; j++) st += p[-TS + j], sl += p[j * TS - 1];
                v = m == 2 ? (st + sl + 16) >> 5 : m == 4 ? (sl + 8) >> 4 : m == 5 ? (st + 8) >> 4 : 128;
            } else {
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int j = 0; j < 4; j++) s0 += p[-TS + j], s1 += p[-TS + 4 + j], s2 += p[j * TS - 1], s3 += p[(4 + j) * TS - 1];
                const int qd = (y >> 2) * 2 + (x >> 2);
                if (m == 2)
                    v = qd == 0 ? (s0 + s2 + 4) >> 3 : qd == 1 ? (s1 + 2) >> 2 : qd == 2 ? (s3 + 2) >> 2 : (s1 + s3 + 4) >> 3;
                else if (m == 4)
                    v = (y >> 2) ? (s3 + 2) >> 2 : (s2 + 2) >> 2;
                else if (m == 5)
                    v = (x >> 2) ? (s1 + 2) >> 2 : (s0 + 2) >> 2;
                else
                    v = 128;
            }
            out[k] = (uint8_t)v;
        }
        __syncwarp();
        for (int k = 0; k < N * N / 32; k++) {
            const int i = t * (N * N / 32) + k;
            p[(i / N) * TS + i % N] = out[k];
        }
        break;
    }
    case OP_PRED4: {
        __shared__ uint8_t e[16];
        if (t < 4) e[t] = p[(3 - t) * TS - 1];
        if (t == 4) e[4] = p[-TS - 1];
        if (t >= 5 && t < 13) e[t] = p[-TS + t - 5];
        __syncwarp();
        int v = 0;
        if (t < 16) v = pred4x4_sample(a[0], t & 3, t >> 2, e);
        __syncwarp();
        if (t < 16) p[(t >> 2) * TS + (t & 3)] = (uint8_t)v;
        break;
    }
    case OP_DBF_LUMA:
        // the PACKED two-lines-per-register filters of swar.cuh, i.e. the instructions the frame kernel executes
        // (VABSDIFF4 / VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.RELU): thread t filters lines 2t and 2t+1, which lie in
        // the same 4-line segment and therefore share tc0
        if (t < 8) {
            const int xs = a[0] == 0 ? 1 : TS, ys = a[0] == 0 ? TS : 1;
            uint8_t *pa = p + (2 * t) * ys, *pb = pa + ys;
            uint32_t v[8];
            for (int k = 0; k < 8; k++) v[k] = (uint32_t)pa[(k - 4) * xs] | ((uint32_t)pb[(k - 4) * xs] << 16);
            if (a[3]) {
                const swar::EdgeK k4 = swar::edge_k(a[1], a[2], 0);
                swar::luma_strong(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], k4, a[1]);
            } else {
                const int tc0 = a[4 + (t >> 1)];
                if (tc0 >= 0) {
                    const swar::EdgeK k = swar::edge_k(a[1], a[2], tc0);
                    swar::luma_normal(v[1], v[2], v[3], v[4], v[5], v[6], k);
                }
            }
            for (int k = 1; k < 7; k++) pa[(k - 4) * xs] = (uint8_t)(v[k] & 0xff), pb[(k - 4) * xs] = (uint8_t)((v[k] >> 16) & 0xff);
        }
        break;
    case OP_DBF_CHROMA:
        if (t < 4) {
            const int xs = a[0] == 0 ? 1 : TS, ys = a[0] == 0 ? TS : 1;
            uint8_t *pa = p + (2 * t) * ys, *pb = pa + ys;
            uint32_t v[4];
            for (int k = 0; k < 4; k++) v[k] = (uint32_t)pa[(k - 2) * xs] | ((uint32_t)pb[(k - 2) * xs] << 16);
            const int tc = a[4 + t];
            if (a[3] || tc > 0) {
                // chroma_edge2 takes tc0 and adds the +1 of core/frame.c:351-377 itself; the table passes tc = tc0 + 1
                const swar::EdgeK k = swar::edge_k(a[1], a[2], a[3] ? 0 : tc - 1);
                swar::chroma_edge2(v[0], v[1], v[2], v[3], k, a[3] != 0);
            }
            pa[-xs] = (uint8_t)(v[1] & 0xff), pb[-xs] = (uint8_t)((v[1] >> 16) & 0xff);
            pa[0] = (uint8_t)(v[2] & 0xff), pb[0] = (uint8_t)((v[2] >> 16) & 0xff);
        }
        break;
    case OP_SSD: {
        int acc = 0;
        for (int i = t; i < a[0] * a[1]; i += 32) {
            const int d = (int)p[(i / a[0]) * TS + i % a[0]] - (int)q[(i / a[0]) * TS + i % a[0]];
            acc += d * d;
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (t == 0) s->result[0] = acc;
        break;
    }
    }
}

// ------------------------------------------------------------------ host side
struct Ctx {
    std::mutex mu;
    Scratch *h = nullptr, *d = nullptr;
    cudaStream_t st = nullptr;
    bool ok = false, tried = false;
    bool init()
    {
        if (tried) return ok;
        tried = true;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
            cudaGetLastError();
            fprintf(stderr, "p264: function tables need a CUDA device (no CPU fallback in this library)\n");
            return false;
        }
        ok = cudaMallocHost(&h, sizeof(Scratch)) == cudaSuccess && cudaMalloc(&d, sizeof(Scratch)) == cudaSuccess &&
             cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        if (ok) memset(h, 0, sizeof(Scratch));
        return ok;
    }
    void run(int op)
    {
        cudaMemcpyAsync(d, h, sizeof(Scratch), cudaMemcpyHostToDevice, st);
        blockop_kernel<<<1, 32, 0, st>>>(d, op);
        cudaMemcpyAsync(h, d, sizeof(Scratch), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
    }
};
Ctx g;

struct Lock {
    std::unique_lock<std::mutex> l;
    bool ok;
    Lock() : l(g.mu), ok(g.init())
    {
        if (!ok) {
            fprintf(stderr, "p264: table call without a usable CUDA device -- aborting the call\n");
        }
    }
};

// copy a w x h region around (x0,y0) relative to the tile origin
void put(uint8_t *tile, const uint8_t *src, int stride, int x0, int y0, int w, int h)
{
    for (int y = 0; y < h; y++) memcpy(tile + TO + (y0 + y) * TS + x0, src + (ptrdiff_t)(y0 + y) * stride + x0, w);
}
void get(const uint8_t *tile, uint8_t *dst, int stride, int x0, int y0, int w, int h)
{
    for (int y = 0; y < h; y++) memcpy(dst + (ptrdiff_t)(y0 + y) * stride + x0, tile + TO + (y0 + y) * TS + x0, w);
}

// ---- dct
void add_idct4_n(uint8_t *dst, int stride, const int16_t *c, int n)
{
    Lock k;
    if (!k.ok) return;
    const int sz = n == 1 ? 4 : n == 4 ? 8 : 16;
    put(g.h->pix, dst, stride, 0, 0, sz, sz);
    memcpy(g.h->coef, c, n * 32);
    g.h->args[0] = n;
    g.run(OP_ADD_IDCT4);
    get(g.h->pix, dst, stride, 0, 0, sz, sz);
}
void t_add4x4_idct(uint8_t *d, int s, int16_t c[4][4]) { add_idct4_n(d, s, &c[0][0], 1); }
void t_add8x8_idct(uint8_t *d, int s, int16_t c[4][4][4]) { add_idct4_n(d, s, &c[0][0][0], 4); }
void t_add16x16_idct(uint8_t *d, int s, int16_t c[16][4][4]) { add_idct4_n(d, s, &c[0][0][0], 16); }
void add_idct8_n(uint8_t *dst, int stride, int16_t *c, int n)
{
    Lock k;
    if (!k.ok) return;
    const int sz = n == 1 ? 8 : 16;
    put(g.h->pix, dst, stride, 0, 0, sz, sz);
    memcpy(g.h->coef, c, n * 128);
    g.h->args[0] = n;
    g.run(OP_ADD_IDCT8);
    get(g.h->pix, dst, stride, 0, 0, sz, sz);
}
void t_add8x8_idct8(uint8_t *d, int s, int16_t c[8][8]) { add_idct8_n(d, s, &c[0][0], 1); }
void t_add16x16_idct8(uint8_t *d, int s, int16_t c[4][8][8]) { add_idct8_n(d, s, &c[0][0][0], 4); }
void coef_op(int16_t *c, int n, int op, const int *tab, int ntab, int qp)
{
    Lock k;
    if (!k.ok) return;
    memcpy(g.h->coef, c, n * 2);
    if (tab) memcpy(g.h->tab, tab, ntab * sizeof(int));
    g.h->args[0] = qp;
    g.run(op);
    memcpy(c, g.h->coef, n * 2);
}
void t_idct4x4dc(int16_t d[4][4]) { coef_op(&d[0][0], 16, OP_IDCT4DC, nullptr, 0, 0); }
void t_dct2x2dc(int16_t d[2][2]) { coef_op(&d[0][0], 4, OP_IDCT2DC, nullptr, 0, 0); }
void t_dequant_4x4(int16_t d[4][4], int mf[6][4][4], int qp) { coef_op(&d[0][0], 16, OP_DEQUANT4, &mf[0][0][0], 96, qp); }
void t_dequant_8x8(int16_t d[8][8], int mf[6][8][8], int qp) { coef_op(&d[0][0], 64, OP_DEQUANT8, &mf[0][0][0], 384, qp); }

// ---- mc
void mc_luma_impl(uint8_t **src, int sstride, uint8_t *dst, int dstride, int mvx, int mvy, int w, int h)
{
    Lock k;
    if (!k.ok) return;
    // only the integer plane src[0] is read: the half-pel planes H/V/HV of the reference
    // (core/mc.c:409-451) are recomputed on the fly from it
    const uint8_t *s0 = src[0] + (ptrdiff_t)(mvy >> 2) * sstride + (mvx >> 2);
    put(g.h->pix, s0, sstride, -2, -2, w + 6, h + 6);
    g.h->args[0] = w, g.h->args[1] = h, g.h->args[2] = mvx & 3, g.h->args[3] = mvy & 3;
    g.run(OP_MC_LUMA);
    get(g.h->pix2, dst, dstride, 0, 0, w, h);
}
void t_mc_luma(uint8_t **src, int ss, uint8_t *dst, int ds, int mvx, int mvy, int w, int h) { mc_luma_impl(src, ss, dst, ds, mvx, mvy, w, h); }
uint8_t *t_get_ref(uint8_t **src, int ss, uint8_t *dst, int *ds, int mvx, int mvy, int w, int h)
{
    mc_luma_impl(src, ss, dst, *ds, mvx, mvy, w, h);  // always materialised in dst (see header)
    return dst;
}
void t_mc_chroma(uint8_t *src, int ss, uint8_t *dst, int ds, int mvx, int mvy, int w, int h)
{
    Lock k;
    if (!k.ok) return;
    const uint8_t *s0 = src + (ptrdiff_t)(mvy >> 3) * ss + (mvx >> 3);
    put(g.h->pix, s0, ss, 0, 0, w + 1, h + 1);
    g.h->args[0] = w, g.h->args[1] = h, g.h->args[2] = mvx & 7, g.h->args[3] = mvy & 7;
    g.run(OP_MC_CHROMA);
    get(g.h->pix2, dst, ds, 0, 0, w, h);
}
void avg_impl(uint8_t *dst, int ds, uint8_t *src, int ss, int w, int h, int weight, bool weighted)
{
    Lock k;
    if (!k.ok) return;
    put(g.h->pix, dst, ds, 0, 0, w, h);
    put(g.h->pix2, src, ss, 0, 0, w, h);
    g.h->args[0] = w, g.h->args[1] = h, g.h->args[2] = weight;
    g.run(weighted ? OP_AVG_WEIGHT : OP_AVG);
    get(g.h->pix, dst, ds, 0, 0, w, h);
}
const int kAvgW[10] = {16, 16, 8, 8, 8, 4, 4, 4, 2, 2}, kAvgH[10] = {16, 8, 16, 8, 4, 8, 4, 2, 4, 2};
template <int I>
void t_avg(uint8_t *d, int ds, uint8_t *s, int ss) { avg_impl(d, ds, s, ss, kAvgW[I], kAvgH[I], 0, false); }
template <int I>
void t_avgw(uint8_t *d, int ds, uint8_t *s, int ss, int w) { avg_impl(d, ds, s, ss, kAvgW[I], kAvgH[I], w, true); }

// ---- intra prediction: stage the neighbours the reference routine reads (core/predict.c)
template <int N, int OPC, int MODE>
void t_pred(uint8_t *src, int stride)
{
    Lock k;
    if (!k.ok) return;
    const int tr = (N == 4) ? 4 : 0;
    put(g.h->pix, src, stride, -1, -1, N + 1 + tr, 1);  // top-left, top, (top-right)
    put(g.h->pix, src, stride, -1, 0, 1, N);            // left
    g.h->args[0] = MODE;
    g.run(OPC);
    get(g.h->pix, src, stride, 0, 0, N, N);
}

// ---- deblock (core/frame.c:302-470): v = horizontal edge (filter across rows), h = vertical edge
void dbf_impl(uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0, bool chroma, bool vdir)
{
    Lock k;
    if (!k.ok) return;
    const int n = chroma ? 8 : 16, reach = chroma ? 2 : 4;
    if (vdir)
        put(g.h->pix, pix, stride, 0, -reach, n, 2 * reach);
    else
        put(g.h->pix, pix, stride, -reach, 0, 2 * reach, n);
    g.h->args[0] = vdir ? 1 : 0, g.h->args[1] = alpha, g.h->args[2] = beta, g.h->args[3] = tc0 ? 0 : 1;
    for (int i = 0; i < 4; i++) g.h->args[4 + i] = tc0 ? tc0[i] : 0;
    g.run(chroma ? OP_DBF_CHROMA : OP_DBF_LUMA);
    if (vdir)
        get(g.h->pix, pix, stride, 0, -reach, n, 2 * reach);
    else
        get(g.h->pix, pix, stride, -reach, 0, 2 * reach, n);
}
void t_dbf_v_luma(uint8_t *p, int s, int a, int b, int8_t *tc) { dbf_impl(p, s, a, b, tc, false, true); }
void t_dbf_h_luma(uint8_t *p, int s, int a, int b, int8_t *tc) { dbf_impl(p, s, a, b, tc, false, false); }
void t_dbf_v_chroma(uint8_t *p, int s, int a, int b, int8_t *tc) { dbf_impl(p, s, a, b, tc, true, true); }
void t_dbf_h_chroma(uint8_t *p, int s, int a, int b, int8_t *tc) { dbf_impl(p, s, a, b, tc, true, false); }
void t_dbf_v_luma_i(uint8_t *p, int s, int a, int b) { dbf_impl(p, s, a, b, nullptr, false, true); }
void t_dbf_h_luma_i(uint8_t *p, int s, int a, int b) { dbf_impl(p, s, a, b, nullptr, false, false); }
void t_dbf_v_chroma_i(uint8_t *p, int s, int a, int b) { dbf_impl(p, s, a, b, nullptr, true, true); }
void t_dbf_h_chroma_i(uint8_t *p, int s, int a, int b) { dbf_impl(p, s, a, b, nullptr, true, false); }

// ---- pixel: SSD only (PSNR / parity utility); core/pixel.c:79-115
const int kPixW[7] = {16, 16, 8, 8, 8, 4, 4}, kPixH[7] = {16, 8, 16, 8, 4, 8, 4};
template <int I>
int t_ssd(uint8_t *a, int sa, uint8_t *b, int sb)
{
    Lock k;
    if (!k.ok) return -1;
    put(g.h->pix, a, sa, 0, 0, kPixW[I], kPixH[I]);
    put(g.h->pix2, b, sb, 0, 0, kPixW[I], kPixH[I]);
    g.h->args[0] = kPixW[I], g.h->args[1] = kPixH[I];
    g.run(OP_SSD);
    return g.h->result[0];
}

}  // namespace

extern "C" {

int p264b200_tables_ready(void)
{
    std::unique_lock<std::mutex> l(g.mu);
    return g.init() ? 0 : P264B200_ENODEV;
}

void p264_dct_init(int, p264_dct_function_t *f)
{
    memset(f, 0, sizeof(*f));  // forward transforms: encoder-only, left NULL
    f->add4x4_idct = t_add4x4_idct;
    f->add8x8_idct = t_add8x8_idct;
    f->add16x16_idct = t_add16x16_idct;
    f->add8x8_idct8 = t_add8x8_idct8;
    f->add16x16_idct8 = t_add16x16_idct8;
    f->idct4x4dc = t_idct4x4dc;
    f->dct2x2dc = t_dct2x2dc;  // the reference installs the same routine in both slots (core/dct.c:401-402)
    f->idct2x2dc = t_dct2x2dc;
}

void p264_quant_init(struct p264_t *, int, p264_quant_function_t *f)
{
    memset(f, 0, sizeof(*f));  // quant cores: encoder-only, left NULL
    f->dequant_4x4 = t_dequant_4x4;
    f->dequant_8x8 = t_dequant_8x8;
}
void p264_mb_dequant_4x4_dc(int16_t d[4][4], int mf[6][4][4], int qp) { coef_op(&d[0][0], 16, OP_DEQUANT4DC, &mf[0][0][0], 96, qp); }
void p264_mb_dequant_2x2_dc(int16_t d[2][2], int mf[6][4][4], int qp) { coef_op(&d[0][0], 4, OP_DEQUANT2DC, &mf[0][0][0], 96, qp); }

void p264_mc_init(int, p264_mc_functions_t *f)
{
    f->mc_luma = t_mc_luma;
    f->get_ref = t_get_ref;
    f->mc_chroma = t_mc_chroma;
    f->avg[0] = t_avg<0>, f->avg[1] = t_avg<1>, f->avg[2] = t_avg<2>, f->avg[3] = t_avg<3>, f->avg[4] = t_avg<4>;
    f->avg[5] = t_avg<5>, f->avg[6] = t_avg<6>, f->avg[7] = t_avg<7>, f->avg[8] = t_avg<8>, f->avg[9] = t_avg<9>;
    f->avg_weight[0] = t_avgw<0>, f->avg_weight[1] = t_avgw<1>, f->avg_weight[2] = t_avgw<2>, f->avg_weight[3] = t_avgw<3>;
    f->avg_weight[4] = t_avgw<4>, f->avg_weight[5] = t_avgw<5>, f->avg_weight[6] = t_avgw<6>, f->avg_weight[7] = t_avgw<7>;
    f->avg_weight[8] = t_avgw<8>, f->avg_weight[9] = t_avgw<9>;
}

void p264_predict_16x16_init(int, p264_predict_t pf[7])
{
    pf[0] = t_pred<16, OP_PRED16, 0>, pf[1] = t_pred<16, OP_PRED16, 1>, pf[2] = t_pred<16, OP_PRED16, 2>, pf[3] = t_pred<16, OP_PRED16, 3>;
    pf[4] = t_pred<16, OP_PRED16, 4>, pf[5] = t_pred<16, OP_PRED16, 5>, pf[6] = t_pred<16, OP_PRED16, 6>;
}
void p264_predict_8x8c_init(int, p264_predict_t pf[7])
{
    pf[0] = t_pred<8, OP_PRED8C, 0>, pf[1] = t_pred<8, OP_PRED8C, 1>, pf[2] = t_pred<8, OP_PRED8C, 2>, pf[3] = t_pred<8, OP_PRED8C, 3>;
    pf[4] = t_pred<8, OP_PRED8C, 4>, pf[5] = t_pred<8, OP_PRED8C, 5>, pf[6] = t_pred<8, OP_PRED8C, 6>;
}
void p264_predict_4x4_init(int, p264_predict_t pf[12])
{
    pf[0] = t_pred<4, OP_PRED4, 0>, pf[1] = t_pred<4, OP_PRED4, 1>, pf[2] = t_pred<4, OP_PRED4, 2>, pf[3] = t_pred<4, OP_PRED4, 3>;
    pf[4] = t_pred<4, OP_PRED4, 4>, pf[5] = t_pred<4, OP_PRED4, 5>, pf[6] = t_pred<4, OP_PRED4, 6>, pf[7] = t_pred<4, OP_PRED4, 7>;
    pf[8] = t_pred<4, OP_PRED4, 8>, pf[9] = t_pred<4, OP_PRED4, 9>, pf[10] = t_pred<4, OP_PRED4, 10>, pf[11] = t_pred<4, OP_PRED4, 11>;
}
void p264_predict_8x8_init(int, p264_predict8x8_t pf[12])
{
    for (int i = 0; i < 12; i++) pf[i] = nullptr;  // Intra-8x8 is unreachable in the reference decoder
}

void p264_deblock_init(int, p264_deblock_function_t *f)
{
    f->deblock_v_luma = t_dbf_v_luma;
    f->deblock_h_luma = t_dbf_h_luma;
    f->deblock_v_chroma = t_dbf_v_chroma;
    f->deblock_h_chroma = t_dbf_h_chroma;
    f->deblock_v_luma_intra = t_dbf_v_luma_i;
    f->deblock_h_luma_intra = t_dbf_h_luma_i;
    f->deblock_v_chroma_intra = t_dbf_v_chroma_i;
    f->deblock_h_chroma_intra = t_dbf_h_chroma_i;
}

void p264_pixel_init(int, p264_pixel_function_t *f)
{
    memset(f, 0, sizeof(*f));  // SAD / SATD / SA8D: encoder-only, never called by the decoder
    f->ssd[0] = t_ssd<0>, f->ssd[1] = t_ssd<1>, f->ssd[2] = t_ssd<2>, f->ssd[3] = t_ssd<3>;
    f->ssd[4] = t_ssd<4>, f->ssd[5] = t_ssd<5>, f->ssd[6] = t_ssd<6>;
}

}  // extern "C"
