// Border expansion of a reconstructed picture: 32 luma / 16 chroma samples of edge replication
// so that motion vectors may point outside the picture (p264_frame_expand_border,
// core/frame.c:183-213).  The half-pel planes of p264_frame_filter / expand_border_filtered
// (core/mc.c:409-451, core/frame.c:215-222) do not exist here: MC filters on the fly.
#pragma once
#include "common.cuh"

namespace p264b200 {

__device__ __forceinline__ void border_plane_word(uint8_t *plane, int stride, int W, int H, int pad, int idx)
{
    // idx enumerates the 32-bit words of the border region of one plane
    const int rw = (W + 2 * pad) >> 2;       // words per padded row
    const int band = pad * rw;               // words in the top (or bottom) band
    int x, y;                                // sample coordinates of the word's first byte
    if (idx < 2 * band) {
        const int b = idx >= band;
        const int k = idx - b * band;
        y = b ? H + k / rw : -pad + k / rw;
        x = -pad + 4 * (k % rw);
    } else {
        const int k = idx - 2 * band, sw = pad >> 1;  // side words per row (left + right)
        y = k / sw;
        const int j = k % sw;
        x = j < (pad >> 2) ? -pad + 4 * j : W + 4 * (j - (pad >> 2));
    }
    const int sy = min(max(y, 0), H - 1);
    const uint8_t *src = plane + (ptrdiff_t)sy * stride;
    uint32_t v;
    if (x >= 0 && x + 3 < W)
        v = *reinterpret_cast<const uint32_t *>(src + x);
    else
        v = 0x01010101u * (uint32_t)src[x < 0 ? 0 : W - 1];
    *reinterpret_cast<uint32_t *>(plane + (ptrdiff_t)y * stride + x) = v;
}

__device__ __forceinline__ int border_words(int W, int H, int pad) { return 2 * pad * ((W + 2 * pad) >> 2) + H * (pad >> 1); }

#ifdef P264B200_DEFINE_KERNELS
__global__ void __launch_bounds__(256) border_kernel(const FrameDesc *__restrict__ descs, Geometry g, uint8_t *y,
                                                     uint8_t *u, uint8_t *v)
{
    uint8_t *pl[3] = {y, u, v};
    if (descs) {
        const FrameDesc &fd = descs[blockIdx.y];
        pl[0] = fd.cur[0], pl[1] = fd.cur[1], pl[2] = fd.cur[2];
    }
    const int ny = border_words(g.width, g.height, kLumaPad);
    const int nc = border_words(g.width / 2, g.height / 2, kChromaPad);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ny)
        border_plane_word(pl[0], g.y_stride, g.width, g.height, kLumaPad, idx);
    else if (idx < ny + nc)
        border_plane_word(pl[1], g.c_stride, g.width / 2, g.height / 2, kChromaPad, idx - ny);
    else if (idx < ny + 2 * nc)
        border_plane_word(pl[2], g.c_stride, g.width / 2, g.height / 2, kChromaPad, idx - ny - nc);
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
