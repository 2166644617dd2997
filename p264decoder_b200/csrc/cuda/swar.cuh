// Two-samples-per-register (s16x2) forms of the deblocking edge filters (core/frame.c:302-470).
//
// A register holds the same tap (p2, p1, p0, q0, ...) of TWO neighbouring sample lines, one per
// 16-bit field, values 0..255.  Both lines always belong to the same 4-sample (luma) / 2-sample
// (chroma) segment, so they share bS, alpha, beta and tc0.  The arithmetic uses the sm_100a packed
// 16-bit integer instructions (VIADD.16, VIMNMX[3].S16x2, VIADDMNMX.S16x2.RELU), VABSDIFF4.U8 for the
// |a-b| tests and plain 32-bit adds where both fields provably stay in 0..65535 (no borrow into the
// neighbouring field).  Every formula keeps the reference's rounding and clipping points.
//
// The file also compiles as plain C++ (no CUDA) with bit-accurate emulations of the intrinsics, which
// is how tests/test_swar_filters.py checks the packed filters against the scalar ones line by line.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SWAR_FN __device__ __forceinline__
#else
#define SWAR_FN static inline
#endif

namespace p264b200 {
namespace swar {

#if defined(__CUDA_ARCH__)
SWAR_FN uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }                  // per-field a + b (wraps)
SWAR_FN uint32_t vabsdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }          // per-byte |a - b|
SWAR_FN uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }                  // per-field signed max
SWAR_FN uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
SWAR_FN uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
SWAR_FN uint32_t vaddmin_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }  // max(min(a+b, c), 0)
SWAR_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
#else
SWAR_FN uint32_t f2(int lo, int hi) { return (uint32_t)(lo & 0xffff) | ((uint32_t)(hi & 0xffff) << 16); }
SWAR_FN int slo(uint32_t a) { return (int16_t)(a & 0xffff); }
SWAR_FN int shi(uint32_t a) { return (int16_t)(a >> 16); }
SWAR_FN uint32_t vadd2(uint32_t a, uint32_t b) { return f2(slo(a) + slo(b), shi(a) + shi(b)); }
SWAR_FN uint32_t vabsdiff4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
SWAR_FN int mx(int a, int b) { return a > b ? a : b; }
SWAR_FN int mn(int a, int b) { return a < b ? a : b; }
SWAR_FN uint32_t vmax2(uint32_t a, uint32_t b) { return f2(mx(slo(a), slo(b)), mx(shi(a), shi(b))); }
SWAR_FN uint32_t vmin2(uint32_t a, uint32_t b) { return f2(mn(slo(a), slo(b)), mn(shi(a), shi(b))); }
SWAR_FN uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
SWAR_FN uint32_t vaddmin_relu(uint32_t a, uint32_t b, uint32_t c)
{
    // the hardware adds with 16-bit wrap-around before the min / max
    return f2(mx(mn((int16_t)(slo(a) + slo(b)), slo(c)), 0), mx(mn((int16_t)(shi(a) + shi(b)), shi(c)), 0));
}
SWAR_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int sel = (s >> (4 * i)) & 0xf;
        int byte = (int)((v >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= (uint32_t)byte << (8 * i);
    }
    return r;
}
#endif

constexpr uint32_t kOnes = 0x00010001u, kBias256 = 0x01000100u, kFF = 0x00ff00ffu;

SWAR_FN uint32_t rep2(int v) { return (uint32_t)(v & 0xffff) * kOnes; }           // v in both fields
SWAR_FN uint32_t sign_mask(uint32_t x) { return prmt(x, 0, 0xbb99); }             // field < 0 ? 0xffff : 0
SWAR_FN uint32_t sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }  // one LOP3

// Per-edge constants shared by every line of the edge (scalar work, once per thread and edge).
struct EdgeK {
    uint32_t n_alpha, n_beta;  // -alpha, -beta in both fields (alpha 0 disables the edge)
    uint32_t tc0;              // tc0 in both fields
};
SWAR_FN EdgeK edge_k(int alpha, int beta, int tc0)
{
    EdgeK k;
    k.n_alpha = rep2(-alpha);
    k.n_beta = rep2(-beta);
    k.tc0 = rep2(tc0);
    return k;
}

// filterSamplesFlag of both lines: 0xffff where |p0-q0| < alpha && |p1-p0| < beta && |q1-q0| < beta
SWAR_FN uint32_t filter_mask(uint32_t p1, uint32_t p0, uint32_t q0, uint32_t q1, const EdgeK &k)
{
    const uint32_t x0 = vadd2(vabsdiff4(p0, q0), k.n_alpha);
    const uint32_t x1 = vadd2(vabsdiff4(p1, p0), k.n_beta);
    const uint32_t x2 = vadd2(vabsdiff4(q1, q0), k.n_beta);
    return sign_mask(vmax3(x0, x1, x2));
}

// delta = clip3((((q0-p0)<<2) + (p1-q1) + 4) >> 3, -tc, tc) applied to p0 / q0 with the 0..255 clip;
// returns the new p0, q0 (unselected).  tc: both fields, 0..~27.
SWAR_FN void delta_pq(uint32_t p1, uint32_t p0, uint32_t q0, uint32_t q1, uint32_t tc, uint32_t &np0, uint32_t &nq0)
{
    // biased by 2048 so that both fields stay positive through the 32-bit arithmetic
    uint32_t u = p1 + 0x08040804u - q1;
    u = u + 4 * q0 - 4 * p0;
    const uint32_t d = (u >> 3) & 0x1fff1fffu;                         // delta + 256
    const uint32_t dc = vmin2(vmax2(d, kBias256 - tc), kBias256 + tc);  // clipped, still + 256
    np0 = vaddmin_relu(p0, vadd2(dc, 0xff00ff00u), kFF);                // clip8(p0 + delta)
    nq0 = vaddmin_relu(q0, vadd2(0x02000200u - dc, 0xff00ff00u), kFF);  // clip8(q0 - delta)
}

// bS < 4 luma filter (core/frame.c:302-341) on two lines.  p3/q3 are not used by this branch.
SWAR_FN void luma_normal(uint32_t p2, uint32_t &p1, uint32_t &p0, uint32_t &q0, uint32_t &q1, uint32_t q2, const EdgeK &k)
{
    const uint32_t f = filter_mask(p1, p0, q0, q1, k);
    const uint32_t ap = sign_mask(vadd2(vabsdiff4(p2, p0), k.n_beta));
    const uint32_t aq = sign_mask(vadd2(vabsdiff4(q2, q0), k.n_beta));
    const uint32_t avg = ((p0 + q0 + kOnes) >> 1) & 0x01ff01ffu;
    const uint32_t lo = kBias256 - k.tc0, hi = kBias256 + k.tc0;
    // p1 += clip3(((p2 + avg) >> 1) - p1, -tc0, tc0), same for q1; the +256 bias keeps the difference positive
    const uint32_t vp = ((p2 + avg) >> 1) & 0x01ff01ffu, vq = ((q2 + avg) >> 1) & 0x01ff01ffu;
    const uint32_t np1 = p1 + vmin2(vmax2(vp + kBias256 - p1, lo), hi) - kBias256;
    const uint32_t nq1 = q1 + vmin2(vmax2(vq + kBias256 - q1, lo), hi) - kBias256;
    const uint32_t tc = k.tc0 + (ap & kOnes) + (aq & kOnes);
    uint32_t np0, nq0;
    delta_pq(p1, p0, q0, q1, tc, np0, nq0);
    p1 = sel(f & ap, np1, p1);
    q1 = sel(f & aq, nq1, q1);
    p0 = sel(f, np0, p0);
    q0 = sel(f, nq0, q0);
}

// chroma filter on two lines: bS < 4 (core/frame.c:351-377, tc = tc0 + 1) or bS == 4 (:443-462)
SWAR_FN void chroma_edge2(uint32_t p1, uint32_t &p0, uint32_t &q0, uint32_t q1, const EdgeK &k, bool strong)
{
    const uint32_t f = filter_mask(p1, p0, q0, q1, k);
    uint32_t np0, nq0;
    if (strong) {
        np0 = ((2 * p1 + p0 + q1 + 2 * kOnes) >> 2) & 0x00ff00ffu;
        nq0 = ((2 * q1 + q0 + p1 + 2 * kOnes) >> 2) & 0x00ff00ffu;
    } else {
        delta_pq(p1, p0, q0, q1, k.tc0 + kOnes, np0, nq0);
    }
    p0 = sel(f, np0, p0);
    q0 = sel(f, nq0, q0);
}

// bS == 4 luma filter (core/frame.c:387-433) on two lines
SWAR_FN void luma_strong(uint32_t p3, uint32_t &p2, uint32_t &p1, uint32_t &p0, uint32_t &q0, uint32_t &q1, uint32_t &q2, uint32_t q3,
                         const EdgeK &k, int alpha)
{
    const uint32_t f = filter_mask(p1, p0, q0, q1, k);
    const uint32_t ap = sign_mask(vadd2(vabsdiff4(p2, p0), k.n_beta));
    const uint32_t aq = sign_mask(vadd2(vabsdiff4(q2, q0), k.n_beta));
    const uint32_t sm = sign_mask(vadd2(vabsdiff4(p0, q0), rep2(-((alpha >> 2) + 2))));  // |p0-q0| < (alpha>>2)+2
    const uint32_t m8 = 0x00ff00ffu;
    const uint32_t wp0 = ((2 * p1 + p0 + q1 + 2 * kOnes) >> 2) & m8, wq0 = ((2 * q1 + q0 + p1 + 2 * kOnes) >> 2) & m8;
    const uint32_t sp0 = ((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4 * kOnes) >> 3) & m8;
    const uint32_t sp1 = ((p2 + p1 + p0 + q0 + 2 * kOnes) >> 2) & m8;
    const uint32_t sp2 = ((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4 * kOnes) >> 3) & m8;
    const uint32_t sq0 = ((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4 * kOnes) >> 3) & m8;
    const uint32_t sq1 = ((p0 + q0 + q1 + q2 + 2 * kOnes) >> 2) & m8;
    const uint32_t sq2 = ((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4 * kOnes) >> 3) & m8;
    const uint32_t fp = f & sm & ap, fq = f & sm & aq;
    const uint32_t r0 = sel(fp, sp0, wp0), s0 = sel(fq, sq0, wq0);
    p2 = sel(fp, sp2, p2);
    p1 = sel(fp, sp1, p1);
    q2 = sel(fq, sq2, q2);
    q1 = sel(fq, sq1, q1);
    p0 = sel(f, r0, p0);
    q0 = sel(f, s0, q0);
}

}  // namespace swar
}  // namespace p264b200
