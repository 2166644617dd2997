// MD5 of reconstructed pictures on the device (RFC 1321), for verification without a device -> host copy of the
// pictures: the reference's own test is "decode bin/f26.264, md5 the YUV" (SURVEY.md 4), and a 1080p picture is 3.1 MB
// of PCIe traffic against 16 bytes of digest.  One thread per picture walks the tight I420 image (Y, then U, then V,
// row by row out of the padded planes) -- MD5 is a serial chain, so the parallelism is across lanes, not inside a
// picture: a verification path (about 25 ms per 1080p picture and thread), never on the reconstruction path.
#pragma once
#include "border.cuh"
#include "common.cuh"

namespace p264b200 {

#ifdef P264B200_DEFINE_KERNELS
__device__ __forceinline__ uint32_t md5_rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

__device__ void md5_block(uint32_t st[4], const uint32_t w[16])
{
    // T[i] = floor(2^32 * |sin(i + 1)|)
    const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
        0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
        0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
        0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
        0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
        0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    const int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16)
            f = (b & c) | (~b & d), g = i;
        else if (i < 32)
            f = (d & b) | (~d & c), g = (5 * i + 1) & 15;
        else if (i < 48)
            f = b ^ c ^ d, g = (3 * i + 5) & 15;
        else
            f = c ^ (b | ~d), g = (7 * i) & 15;
        const uint32_t t = d;
        d = c;
        c = b;
        b = b + md5_rotl(a + f + K[i] + w[g], S[i >> 4][i & 3]);
        a = t;
    }
    st[0] += a, st[1] += b, st[2] += c, st[3] += d;
}

// grid: ceil(n / 32) CTAs of 32 threads; thread = one lane's picture (ring slot sel.slot[lane]); digests: [n][16] bytes.
// Widths are multiples of 16 (luma) / 8 (chroma), so rows are whole 32-bit words but a 64-byte MD5 block may end inside a row.
__global__ void __launch_bounds__(32) md5_i420_kernel(const PackSrc *__restrict__ table, int n_slots, PackSel sel, Geometry g, int n, uint8_t *__restrict__ digests)
{
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n) return;
    const PackSrc s = table[lane * n_slots + sel.slot[lane]];
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint32_t w[16];
    int fill = 0;   // words in w
#pragma unroll 1
    for (int p = 0; p < 3; p++) {
        const int width = p ? g.width / 2 : g.width, height = p ? g.height / 2 : g.height, stride = p ? g.c_stride : g.y_stride;
#pragma unroll 1
        for (int y = 0; y < height; y++) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(s.plane[p] + (size_t)y * stride);
#pragma unroll 1
            for (int x = 0; x < width / 4; x++) {
                w[fill & 15] = row[x];
                if (++fill == 16) {
                    md5_block(st, w);
                    fill = 0;
                }
            }
        }
    }
    // padding: 0x80, zeros, 64-bit bit length
    const unsigned long long bits = 8ull * ((unsigned long long)g.width * g.height * 3 / 2);
    w[fill++] = 0x80u;
    if (fill > 14) {
        while (fill < 16) w[fill++] = 0;
        md5_block(st, w);
        fill = 0;
    }
    while (fill < 14) w[fill++] = 0;
    w[14] = (uint32_t)bits, w[15] = (uint32_t)(bits >> 32);
    md5_block(st, w);
    uint32_t *out = reinterpret_cast<uint32_t *>(digests + 16 * lane);
    out[0] = st[0], out[1] = st[1], out[2] = st[2], out[3] = st[3];
}
#endif

}  // namespace p264b200
