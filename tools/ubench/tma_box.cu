// How many small 2-D TMA boxes (cp.async.bulk.tensor) at pseudo-random positions can one SM pull per cycle?
// The question behind recon_inter's window staging: a motion-compensation window is (w+5) x (h+5) bytes
// (21x21 for a 16x16 partition, 9x9 for a 4x4 one), i.e. a box of 16 or 32 bytes x 3..21 rows per partition.
// Every warp keeps D boxes in flight (own mbarrier per stage), lane 0 issues, all lanes read the box back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_box tma_box.cu && ./tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x)                                                                     \
    do {                                                                          \
        cudaError_t e_ = (x);                                                     \
        if (e_ != cudaSuccess) {                                                  \
            printf("%s: %s\n", #x, cudaGetErrorString(e_));                       \
            exit(1);                                                              \
        }                                                                         \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 10000;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_box(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kMaxWarps = 16, kMaxDepth = 8, kSlot = 1024;

__global__ void __launch_bounds__(32 * kMaxWarps) tma_kernel(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int bx, int by, int depth, int iters, int W, int H,
                                                            int planes, int spread, unsigned *sink, long long *cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[kMaxWarps][kMaxDepth];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CUtensorMap *map = gmap ? gmap : &pmap;
    uint8_t *mine = smem + (size_t)w * kMaxDepth * kSlot;
    if (lane == 0)
        for (int d = 0; d < depth; d++) mbar_init(&bars[w][d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t bytes = (uint32_t)(bx * by);
    uint32_t rng = (blockIdx.x * 977u + w * 131u + 7u) * 2654435761u;
    // every CTA works inside its own region of the plane set (like a tile of macroblocks whose vectors spread +-`spread`)
    const int plane = blockIdx.x % planes;
    const int cx = (int)((blockIdx.x * 613u) % (unsigned)(W - 2 * spread - 64)) + spread, cy = (int)((blockIdx.x * 389u) % (unsigned)(H - 2 * spread - 64)) + spread;
    auto issue = [&](int d) {
        rng = rng * 1664525u + 1013904223u;
        const int x = (cx + (int)((rng >> 8) % (unsigned)(2 * spread + 1)) - spread) & ~15,   /* box origins must be 16-byte aligned: a misaligned x is an illegal instruction */ y = cy + (int)((rng >> 20) % (unsigned)(2 * spread + 1)) - spread;
        mbar_expect(&bars[w][d], bytes);
        tma_box(mine + d * kSlot, map, x, y, plane, &bars[w][d]);
    };
    unsigned acc = 0;
    const long long t0 = clock64();
    if (lane == 0)
        for (int d = 0; d < depth; d++) issue(d);
    for (int i = 0; i < iters; i++) {
        const int d = i % depth;
        mbar_wait(&bars[w][d], (i / depth) & 1);
        acc += *reinterpret_cast<const uint32_t *>(mine + d * kSlot + 4 * (lane % (bytes / 4)));
        __syncwarp();
        if (lane == 0 && i + depth < iters) issue(d);
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const bool use_global = argc > 1 && argv[1][0] == 'g';
    const bool quick = argc > 2;
    CUtensorMap *gmap = nullptr;
    CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
    const int W = 2048, H = 1152, planes = 512;
    uint8_t *buf;
    CK(cudaMalloc(&buf, (size_t)W * H * planes));
    CK(cudaMemset(buf, 1, (size_t)W * H * planes));
    unsigned *sink;
    long long *cyc;
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&cyc, 8 * 4096));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn;
    int dev_clock_khz = 0;
    CK(cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxWarps * kMaxDepth * kSlot));
    printf("box_x box_y warps ctas/SM depth spread | boxes/us/SM  cycles/box/SM  rows/cycle/SM  GB/s(chip)\n");
    const int boxes[][2] = {{48, 21}, {32, 21}, {32, 13}, {32, 9}, {16, 21}, {16, 13}, {16, 9}, {16, 5}, {16, 3}};
    for (auto &b : boxes) {
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
        cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        cuuint32_t box[3] = {(cuuint32_t)b[0], (cuuint32_t)b[1], 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("encode failed %d\n", (int)r);
            return 1;
        }
        const int cfgs[][4] = {{1, 1, 1, 24}, {4, 1, 4, 24}, {8, 2, 4, 24}, {8, 1, 8, 24}, {16, 1, 8, 24}, {16, 1, 8, 200}};
        for (auto &c : cfgs) {
            const int warps = quick ? 1 : c[0], per_sm = c[1], depth = quick ? 1 : c[2], spread = c[3], iters = quick ? 4 : 2000;
            const int grid = quick ? 1 : 148 * per_sm;
            const size_t smem = (size_t)warps * kMaxDepth * kSlot;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0), cudaEventCreate(&e1);
            CK(cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice));
            tma_kernel<<<grid, 32 * warps, smem>>>(map, use_global ? gmap : nullptr, b[0], b[1], depth, 200, W, H, planes, spread, sink, cyc);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            tma_kernel<<<grid, 32 * warps, smem>>>(map, use_global ? gmap : nullptr, b[0], b[1], depth, iters, W, H, planes, spread, sink, cyc);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> hc(grid);
            CK(cudaMemcpy(hc.data(), cyc, 8 * grid, cudaMemcpyDeviceToHost));
            double avg_cyc = 0;
            for (auto v : hc) avg_cyc += (double)v;
            avg_cyc /= grid;
            const double boxes_sm = (double)warps * per_sm * iters;
            printf("%5d %5d %5d %7d %5d %6d | %10.1f %13.2f %13.3f %10.0f   (%.3f ms, %.0f cycles)\n", b[0], b[1], warps, per_sm, depth, spread,
                   boxes_sm / (ms * 1e3), avg_cyc / boxes_sm, boxes_sm * b[1] / avg_cyc, boxes_sm * 148 * b[0] * b[1] / (ms * 1e-3) / 1e9, ms, avg_cyc);
        }
    }
    return 0;
}
