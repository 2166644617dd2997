// Latency / single-warp issue rate / SM throughput of the packed-integer instructions the
// reconstruction kernels are built from (sm_100a).  nvcc -arch=sm_100a -o int_simd int_simd.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define OPS(X)                                                                                         \
    X(lop3, a = (a & b) ^ c)                                                                           \
    X(iadd3, a = a + b + c)                                                                            \
    X(shf, a = __funnelshift_r(a, b, 8) )                                                              \
    X(prmt, a = __byte_perm(a, b, 0x5140))                                                             \
    X(imad, a = a * b + c)                                                                             \
    X(dp4a, a = __dp4a((int)a, (int)b, (int)c))                                                        \
    X(vabsdiff4, a = __vabsdiffu4(a, b))                                                               \
    X(viadd16x2, a = __vadd2(a, b))                                                                    \
    X(vimnmx16x2, a = __vmaxs2(a, b))                                                                  \
    X(vimnmx3_16x2, a = __vimax3_s16x2(a, b, c))                                                       \
    X(viaddmnmx16x2, a = __viaddmin_s16x2_relu(a, b, c))                                               \
    X(i2ip, { uint32_t t; asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(a), "r"(b), "r"(c)); a = t; }) \
    X(imnmx, a = max((int)a, (int)b))                                                                  \
    X(sel, a = (a > c) ? b : a)

#define KERN(name, expr)                                                                               \
    template <int ILP> __global__ void k_##name(uint32_t *out, long long *cyc, int iters)              \
    {                                                                                                  \
        uint32_t v[ILP], b = out[1] | 1, c = out[2] | 3;                                               \
        for (int i = 0; i < ILP; i++) v[i] = out[3] + threadIdx.x + i;                                 \
        __syncthreads();                                                                               \
        long long t0 = clock64();                                                                      \
        for (int it = 0; it < iters; it++) {                                                           \
            _Pragma("unroll") for (int r = 0; r < 16; r++) {                                           \
                _Pragma("unroll") for (int i = 0; i < ILP; i++) { uint32_t a = v[i]; expr; v[i] = a; } \
            }                                                                                          \
        }                                                                                              \
        long long t1 = clock64();                                                                      \
        uint32_t s = 0;                                                                                \
        for (int i = 0; i < ILP; i++) s ^= v[i];                                                       \
        if (s == 0x12345678) out[0] = s;                                                               \
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;                                     \
    }
OPS(KERN)

template <typename K> double run(K k, int warps, int iters, uint32_t *d, long long *dc)
{
    k<<<1, 32 * warps>>>(d, dc, iters);
    cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(d, dc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    return (double)c;
}

int main()
{
    uint32_t *d;
    long long *dc;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    cudaMalloc(&dc, 8);
    const int iters = 256;
    printf("%-16s %10s %14s %16s %16s\n", "op", "latency", "1 warp ILP8", "4 warps ILP8", "16 warps ILP8");
    printf("%-16s %10s %14s %16s %16s\n", "", "cyc/op", "cyc/warp-instr", "cyc/instr/SMSP", "cyc/instr/SMSP");
#define ROW(name, expr)                                                                                \
    {                                                                                                  \
        const double n1 = 16.0 * iters, n8 = 16.0 * 8 * iters;                                         \
        double lat = run(k_##name<1>, 1, iters, d, dc) / n1;                                           \
        double w1 = run(k_##name<8>, 1, iters, d, dc) / n8;                                            \
        double w4 = run(k_##name<8>, 4, iters, d, dc) / n8;                                            \
        double w16 = run(k_##name<8>, 16, iters, d, dc) / (n8 * 4);                                    \
        printf("%-16s %10.2f %14.2f %16.2f %16.2f\n", #name, lat, w1, w4, w16);                        \
    }
    OPS(ROW)
    return 0;
}
