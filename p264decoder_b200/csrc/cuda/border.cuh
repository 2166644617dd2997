// Border expansion of a reconstructed picture: 32 luma / 16 chroma samples of edge replication
// so that motion vectors may point outside the picture (p264_frame_expand_border,
// core/frame.c:183-213).  The half-pel planes of p264_frame_filter / expand_border_filtered
// (core/mc.c:409-451, core/frame.c:215-222) do not exist here: MC filters on the fly.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace p264b200 {

// One VB-byte vector of the border region of one plane (VB = 16 for luma: width and pad are multiples of 16;
// VB = 8 for chroma: the width is only a multiple of 8).  idx enumerates: top band, bottom band, then the left
// and right pads of every picture row.
template <int VB>
__device__ __forceinline__ void border_plane_vec(uint8_t *plane, int stride, int W, int H, int pad, int idx)
{
    typedef typename std::conditional<VB == 16, uint4, uint2>::type Vec;
    const int rv = (W + 2 * pad) / VB;       // vectors per padded row
    const int band = pad * rv;               // vectors in the top (or bottom) band
    int x, y;                                // sample coordinates of the vector's first byte
    if (idx < 2 * band) {
        const int b = idx >= band;
        const int k = idx - b * band, r = k / rv;
        y = b ? H + r : -pad + r;
        x = -pad + VB * (k - r * rv);
    } else {
        const int k = idx - 2 * band, pv = pad / VB, sv = 2 * pv;  // side vectors per row (left + right)
        y = k / sv;
        const int j = k - y * sv;
        x = j < pv ? -pad + VB * j : W + VB * (j - pv);
    }
    const int sy = min(max(y, 0), H - 1);
    const uint8_t *src = plane + (ptrdiff_t)sy * stride;
    Vec v;
    if (x >= 0 && x + VB <= W)
        v = *reinterpret_cast<const Vec *>(src + x);
    else {
        const uint32_t e = 0x01010101u * (uint32_t)src[x < 0 ? 0 : W - 1];
        uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
        for (int i = 0; i < VB / 4; i++) w[i] = e;
    }
    *reinterpret_cast<Vec *>(plane + (ptrdiff_t)y * stride + x) = v;
}

template <int VB>
__host__ __device__ __forceinline__ int border_vecs(int W, int H, int pad) { return 2 * pad * ((W + 2 * pad) / VB) + H * (2 * pad / VB); }
__host__ __device__ __forceinline__ int border_threads(int width, int height)
{
    return border_vecs<16>(width, height, kLumaPad) + 2 * border_vecs<8>(width / 2, height / 2, kChromaPad);
}

#ifdef P264B200_DEFINE_KERNELS
__global__ void __launch_bounds__(256) border_kernel(const FrameDesc *__restrict__ descs, Geometry g, uint8_t *y,
                                                     uint8_t *u, uint8_t *v)
{
    uint8_t *pl[3] = {y, u, v};
    if (descs) {
        const FrameDesc &fd = descs[blockIdx.y];
        pl[0] = fd.cur[0], pl[1] = fd.cur[1], pl[2] = fd.cur[2];
    }
    const int ny = border_vecs<16>(g.width, g.height, kLumaPad);
    const int nc = border_vecs<8>(g.width / 2, g.height / 2, kChromaPad);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ny)
        border_plane_vec<16>(pl[0], g.y_stride, g.width, g.height, kLumaPad, idx);
    else if (idx < ny + nc)
        border_plane_vec<8>(pl[1], g.c_stride, g.width / 2, g.height / 2, kChromaPad, idx - ny);
    else if (idx < ny + 2 * nc)
        border_plane_vec<8>(pl[2], g.c_stride, g.width / 2, g.height / 2, kChromaPad, idx - ny - nc);
}

// Output packing for the batched download: padded planes of `n` pictures -> tight I420 images in one
// contiguous device buffer (Y, U, V back to back), so that the host copy is a single large DMA instead
// of three pitched 2-D copies per picture.  128-bit loads and stores (16 samples per thread).
struct PackSrc {
    const uint8_t *plane[3];
};
constexpr int kPackMaxLanes = 256;
struct PackSel {
    uint8_t slot[kPackMaxLanes];  // ring slot to read for every lane (kernel argument, no staging copy)
};
// table: [lane * n_slots + slot] plane origins, filled once at engine creation
__global__ void __launch_bounds__(256) pack_i420_kernel(const PackSrc *__restrict__ table, int n_slots, PackSel sel, Geometry g,
                                                        uint8_t *__restrict__ dst, size_t picture_bytes)
{
    const PackSrc s = table[blockIdx.y * n_slots + sel.slot[blockIdx.y]];
    uint8_t *out = dst + (size_t)blockIdx.y * picture_bytes;
    const int yv = (g.width >> 4) * g.height;                 // uint4 per luma plane
    const int cv = (g.width >> 5) * (g.height >> 1);          // uint4 per chroma plane
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < yv + 2 * cv; i += gridDim.x * blockDim.x) {
        if (i < yv) {
            const int wv = g.width >> 4, r = i / wv, c = i % wv;
            reinterpret_cast<uint4 *>(out)[i] = *reinterpret_cast<const uint4 *>(s.plane[0] + (size_t)r * g.y_stride + 16 * c);
        } else {
            const int j = i - yv, p = j >= cv, k = j - p * cv, wv = g.width >> 5, r = k / wv, c = k % wv;
            reinterpret_cast<uint4 *>(out + (size_t)g.width * g.height + (size_t)p * (g.width / 2) * (g.height / 2))[k] =
                *reinterpret_cast<const uint4 *>(s.plane[1 + p] + (size_t)r * g.c_stride + 16 * c);
        }
    }
}
#endif  // P264B200_DEFINE_KERNELS

}  // namespace p264b200
