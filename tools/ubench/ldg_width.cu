// What does the L1 data pipe charge for a motion-compensation window row: three scattered 32-bit loads or two 64-bit ones?
// Every lane fetches the 12 (16) bytes of one window row per iteration; groups of 4 neighbouring lanes read the same row at
// consecutive 4-byte offsets (the strips of a 16-wide partition), the rows of different groups are random inside a region
// that stays L1-resident.  Reports cycles per warp-row on one SM with all SMs busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldg_width ldg_width.cu && ./ldg_width
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(384) k(const uint8_t *buf, int region, int stride, int iters, unsigned *sink, long long *cyc)
{
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 2;
    unsigned rng = (blockIdx.x * 7919u + grp * 104729u + 13u) * 2654435761u;
    const uint8_t *base = buf + (size_t)(blockIdx.x % 64) * region;
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        rng = rng * 1664525u + 1013904223u;
        const int row = (rng >> 10) % (region / stride - 16), x = ((rng >> 3) & 63) + 4 * (lane & 3);
        const uint8_t *p = base + (size_t)row * stride + x;
#pragma unroll
        for (int r = 0; r < 13; r++) {
            const uint8_t *q = p + r * stride;
            if (MODE == 0) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>((uintptr_t)q & ~(uintptr_t)3);
                const int sh = ((uintptr_t)q & 3) * 8;
                const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                acc += __funnelshift_r(w0, w1, sh) ^ __funnelshift_r(w1, w2, sh) ^ (w2 >> sh);
            } else {
                const uint2 *w = reinterpret_cast<const uint2 *>((uintptr_t)q & ~(uintptr_t)7);
                int sh = ((uintptr_t)q & 7) * 8;
                const uint2 a = __ldg(w), b = __ldg(w + 1);
                const bool hi = sh >= 32;
                const uint32_t t0 = hi ? a.y : a.x, t1 = hi ? b.x : a.y, t2 = hi ? b.y : b.x;
                sh &= 31;
                acc += __funnelshift_r(t0, t1, sh) ^ __funnelshift_r(t1, t2, sh) ^ (t2 >> sh);
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    const int region = 40 * 1024, stride = 2048, iters = 400;   // 20 rows x 2 KB per CTA region -> wait: rows are 2 KB apart, region holds 20 of them
    uint8_t *buf;
    unsigned *sink;
    long long *cyc;
    const size_t bytes = (size_t)64 * 4 * 1024 * 1024;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 3, bytes);
    cudaMalloc(&sink, 4);
    cudaMalloc(&cyc, 8 * 4096);
    const int big = 4 * 1024 * 1024;   // 2048 rows of 2 KB: random rows mostly miss L1 -> use a small window instead
    for (int reg : {64 * 1024, 256 * 1024, big}) {
        for (int mode = 0; mode < 2; mode++) {
            const int grid = 148 * 3;
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0)
                    k<0><<<grid, 384>>>(buf, reg, stride, iters, sink, cyc);
                else
                    k<1><<<grid, 384>>>(buf, reg, stride, iters, sink, cyc);
                cudaDeviceSynchronize();
            }
            long long h[444];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (auto v : h) avg += (double)v;
            avg /= grid;
            // per SM: 3 CTAs x 12 warps x iters x 13 rows in `avg` cycles
            printf("region %7d B  %s : %.2f cycles per warp-row per SM (%.0f cycles per CTA)\n", reg, mode ? "2 x LDG.64" : "3 x LDG.32", avg / (3.0 * 12 * iters * 13), avg);
        }
    }
    (void)region;
    return 0;
}
