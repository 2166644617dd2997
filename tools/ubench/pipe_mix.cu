// Which of the packed-integer instructions share an issue pipe?  Two independent instruction streams are
// interleaved in one warp (ILP 4 + 4); if both ops sit on the same pipe the pair costs 2 + 2 cycles per
// SMSP, on different pipes ~2.  nvcc -arch=sm_100a -o pipe_mix pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define OP_prmt(a, b, c) a = __byte_perm(a, b, 0x5140)
#define OP_shf(a, b, c) a = __funnelshift_r(a, b, 7)
#define OP_lop3(a, b, c) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c))
#define OP_iadd3(a, b, c) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b))
#define OP_imad(a, b, c) a = a * b + c
#define OP_dp4a(a, b, c) a = __dp4a((int)a, (int)b, (int)c)
#define OP_vabsdiff4(a, b, c) a = __vabsdiffu4(a, b)
#define OP_viadd16(a, b, c) a = __vadd2(a, b)
#define OP_vimnmx16(a, b, c) a = __vmaxs2(a, b) ^ c
#define OP_vimnmx3(a, b, c) a = __vimax3_s16x2(a, b, c)
#define OP_viaddmnmx(a, b, c) a = __viaddmin_s16x2_relu(a, b, c)
#define OP_umulhi(a, b, c) a = __umulhi(a, 0x08000000u) + b
#define OP_mulhi(a, b, c) a = (uint32_t)__mulhi((int)a, 1 << 27) ^ c
#define OP_imadwide(a, b, c) { unsigned long long w = (unsigned long long)a * 0x08000000ull + b; a = (uint32_t)(w >> 32) ^ (uint32_t)w; }
#define OP_i2ip(a, b, c) { uint32_t t; asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c)); a = t; }

#define KERN(n1, n2)                                                                                   \
    __global__ void k_##n1##_##n2(uint32_t *out, long long *cyc, int iters)                            \
    {                                                                                                  \
        uint32_t v[4], u[4], b = out[1] | 1, c = out[2] | 3;                                           \
        for (int i = 0; i < 4; i++) v[i] = out[3] + threadIdx.x + i, u[i] = out[4] + threadIdx.x * 3 + i; \
        __syncthreads();                                                                               \
        long long t0 = clock64();                                                                      \
        for (int it = 0; it < iters; it++) {                                                           \
            _Pragma("unroll") for (int r = 0; r < 16; r++) {                                           \
                _Pragma("unroll") for (int i = 0; i < 4; i++) { OP_##n1(v[i], b, c); OP_##n2(u[i], c, b); } \
            }                                                                                          \
        }                                                                                              \
        long long t1 = clock64();                                                                      \
        uint32_t s = 0;                                                                                \
        for (int i = 0; i < 4; i++) s ^= v[i] ^ u[i];                                                  \
        if (s == 0x12345678) out[0] = s;                                                               \
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;                                     \
    }
#define PAIRS(X) X(prmt, prmt) X(prmt, shf) X(prmt, lop3) X(prmt, iadd3) X(prmt, imad) X(prmt, dp4a) X(prmt, vabsdiff4) X(prmt, viadd16) \
    X(prmt, vimnmx16) X(prmt, vimnmx3) X(prmt, viaddmnmx) X(prmt, i2ip) X(imad, dp4a) X(imad, viadd16) X(imad, vimnmx3) X(imad, vabsdiff4) \
    X(imad, lop3) X(imad, iadd3) X(imad, shf) X(viadd16, vimnmx3) X(vabsdiff4, vimnmx3) X(lop3, iadd3) X(umulhi, umulhi) X(prmt, umulhi) X(imad, umulhi) X(mulhi, mulhi) X(prmt, mulhi) X(prmt, imadwide) X(imad, imadwide)
PAIRS(KERN)

int main()
{
    uint32_t *d;
    long long *dc;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    cudaMalloc(&dc, 8);
    const int iters = 256;
    printf("%-24s %s\n", "pair (4 + 4 chains)", "cycles per instruction per SMSP, 4 warps");
#define ROW(n1, n2)                                                                                    \
    {                                                                                                  \
        for (int rep = 0; rep < 2; rep++) { k_##n1##_##n2<<<1, 128>>>(d, dc, iters); cudaDeviceSynchronize(); } \
        long long c;                                                                                   \
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);                                                 \
        printf("%-24s %.2f\n", #n1 " + " #n2, (double)c / (16.0 * 8 * iters));                         \
    }
    PAIRS(ROW)
    return 0;
}
