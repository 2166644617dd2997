// FrameSyntax v2 -> v1 staging layout, on the device (include/p264b200_recon.h "FrameSyntax v2", csrc/host/wire_v2.cc).
// One thread per macroblock: the 96-byte record is rebuilt from the 32-byte tail + its partition vectors, the
// coefficient chunk from the significance masks + non-zero levels.  Runs on the upload stream right behind the copy
// of the packed pictures, so the reconstruction kernels see exactly what p264b200_stage_frames would have staged.
#pragma once
#include "common.cuh"

namespace p264b200 {

struct V2Desc {
    const uint8_t *blob;                                       // device copy of the packed picture
    uint32_t off_hdr, off_offs, off_mv, off_mask, off_level;
    uint32_t flags, n_coef, pad;
};

#ifdef P264B200_DEFINE_KERNELS
__device__ __forceinline__ int v2_shape_index(int code, int b)
{
    const int bx = b & 3, by = b >> 2, s = code & 3;
    if (s == 0) return 0;
    if (s == 1) return by >> 1;
    if (s == 2) return bx >> 1;
    const int q = (by >> 1) * 2 + (bx >> 1);
    int base = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int qc = (code >> (2 + 2 * k)) & 3;
        if (k < q) base += qc == 0 ? 1 : qc == 3 ? 4 : 2;
    }
    const int qc = (code >> (2 + 2 * q)) & 3;
    return base + (qc == 0 ? 0 : qc == 1 ? (by & 1) : qc == 2 ? (bx & 1) : (by & 1) * 2 + (bx & 1));
}

__global__ void __launch_bounds__(128) expand_v2_kernel(const V2Desc *__restrict__ v2, const FrameDesc *__restrict__ descs, int n_mb)
{
    const V2Desc &vd = v2[blockIdx.y];
    const FrameDesc &fd = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mb) return;
    const uint4 *hp = reinterpret_cast<const uint4 *>(vd.blob + vd.off_hdr) + 2 * i;
    uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
    // tail layout (bytes 64..95 of p264b200_mb): h0.x ref[4] | h0.y type qp qp_dbf cbp | h0.z luma_mask i16 chroma_mode | h0.w i4_mode[0..3]
    //                                             h1.x i4_mode[4..7] | h1.y coef_off | h1.z chroma_mask part sub_part[0..1] | h1.w sub_part[2..3] reserved[2]
    const int type = h0.y & 0xff;
    uint4 *rec = reinterpret_cast<uint4 *>(const_cast<p264b200_mb *>(fd.mbs) + i);
    uint32_t mv[16];
    if (!P264B200_IS_INTRA(type)) {
        const uint32_t *mvs = reinterpret_cast<const uint32_t *>(vd.blob + vd.off_mv) + h0.w;
        const int code = h1.w >> 16;
        h0.w = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) mv[b] = __ldg(mvs + v2_shape_index(code, b));
    } else {
#pragma unroll
        for (int b = 0; b < 16; b++) mv[b] = 0;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) rec[k] = make_uint4(mv[4 * k], mv[4 * k + 1], mv[4 * k + 2], mv[4 * k + 3]);
    rec[4] = h0;
    rec[5] = h1;

    // coefficient chunk: blocks in v1 order, zero-filled, non-zero levels scattered into their slots
    const unsigned luma_mask = h0.z & 0xffff, cbp_chroma = h0.y >> 24, chroma_mask = h1.z & 0xff;
    const int n_luma = (type == P264B200_MB_I16x16 ? 1 : 0) + __popc(luma_mask);
    const int n_blocks = n_luma + (cbp_chroma ? 1 + __popc(chroma_mask) : 0);
    if (!n_blocks) return;
    const uint2 off = __ldg(reinterpret_cast<const uint2 *>(vd.blob + vd.off_offs) + i);
    const uint16_t *masks = reinterpret_cast<const uint16_t *>(vd.blob + vd.off_mask) + off.x;
    const uint8_t *lv = vd.blob + vd.off_level;
    const bool fit8 = vd.flags & P264B200_V2_LEVELS8;
    int16_t *dst = const_cast<int16_t *>(fd.coefs) + h1.y;     // chunks start on 16-byte boundaries
    uint32_t li = off.y;
#pragma unroll 1
    for (int k = 0; k < n_blocks; k++) {
        const bool dc8 = cbp_chroma && k == n_luma;            // the 8-slot chroma DC group
        unsigned mask = __ldg(masks + k);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        d4[0] = make_uint4(0, 0, 0, 0);
        if (!dc8) d4[1] = make_uint4(0, 0, 0, 0);
        while (mask) {
            const int s = __ffs(mask) - 1;
            mask &= mask - 1;
            dst[s] = fit8 ? (int16_t) reinterpret_cast<const int8_t *>(lv)[li] : reinterpret_cast<const int16_t *>(lv)[li];
            li++;
        }
        dst += dc8 ? 8 : 16;
    }
}
#endif

}  // namespace p264b200
