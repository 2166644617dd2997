// Host-side check of the packed (two lines per register) deblocking filters in
// p264decoder_b200/csrc/cuda/swar.cuh against a scalar restatement of core/frame.c:302-470.
// Built and run by tests/test_swar_filters.py (no GPU needed: swar.cuh emulates the intrinsics).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../p264decoder_b200/csrc/cuda/swar.cuh"
using namespace p264b200::swar;

static int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static int clip8(int v) { return clip3(v, 0, 255); }

// core/frame.c:310-338, v = p3 p2 p1 p0 q0 q1 q2 q3
static void luma_normal_ref(int v[8], int alpha, int beta, int tc0)
{
    const int p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        int tc = tc0;
        if (abs(p2 - p0) < beta) { v[2] = p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0); tc++; }
        if (abs(q2 - q0) < beta) { v[5] = q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0); tc++; }
        const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        v[3] = clip8(p0 + delta);
        v[4] = clip8(q0 - delta);
    }
}
// core/frame.c:390-431
static void luma_strong_ref(int v[8], int alpha, int beta)
{
    const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        if (abs(p0 - q0) < ((alpha >> 2) + 2)) {
            if (abs(p2 - p0) < beta) {
                v[3] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
                v[2] = (p2 + p1 + p0 + q0 + 2) >> 2;
                v[1] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
            } else
                v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
            if (abs(q2 - q0) < beta) {
                v[4] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
                v[5] = (p0 + q0 + q1 + q2 + 2) >> 2;
                v[6] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
            } else
                v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
        } else {
            v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
            v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
        }
    }
}
// core/frame.c:360-373 / :446-459, v = p1 p0 q0 q1, tc = tc0 + 1
static void chroma_ref(int v[4], int alpha, int beta, int bs, int tc)
{
    const int p1 = v[0], p0 = v[1], q0 = v[2], q1 = v[3];
    if (abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta) {
        if (bs < 4) {
            const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
            v[1] = clip8(p0 + delta);
            v[2] = clip8(q0 - delta);
        } else {
            v[1] = (2 * p1 + p0 + q1 + 2) >> 2;
            v[2] = (2 * q1 + q0 + p1 + 2) >> 2;
        }
    }
}

static uint64_t rng_s = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
    rng_s ^= rng_s << 13, rng_s ^= rng_s >> 7, rng_s ^= rng_s << 17;
    return (uint32_t)(rng_s >> 32);
}
static void rand_line(int v[8], int mode)
{
    if (mode == 0)
        for (int i = 0; i < 8; i++) v[i] = rnd() & 255;
    else {
        // smooth-ish lines so that the alpha / beta tests pass often; extremes included
        int base = mode == 2 ? (rnd() & 1 ? 0 : 255) : (int)(rnd() & 255);
        const int spread = 1 + rnd() % (mode == 3 ? 40 : 8);
        for (int i = 0; i < 8; i++) v[i] = clip8(base + (int)(rnd() % (2 * spread + 1)) - spread);
        if (rnd() & 1)
            for (int i = 4; i < 8; i++) v[i] = clip8(v[i] + (int)(rnd() % 41) - 20);
    }
}

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    static const int alpha_t[] = {0, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255};
    long bad = 0;
    for (long it = 0; it < n && bad < 10; it++) {
        int a[8], b[8];
        rand_line(a, rnd() & 3);
        rand_line(b, rnd() & 3);
        const int alpha = alpha_t[rnd() % (sizeof(alpha_t) / sizeof(int))], beta = rnd() % 19, tc0 = rnd() % 26;
        uint32_t r[8];
        for (int i = 0; i < 8; i++) r[i] = (uint32_t)a[i] | ((uint32_t)b[i] << 16);
        const int kind = rnd() % 4;
        int ra[8], rb[8];
        memcpy(ra, a, sizeof(a)), memcpy(rb, b, sizeof(b));
        const EdgeK k = edge_k(alpha, beta, tc0);
        if (kind == 0) {
            luma_normal_ref(ra, alpha, beta, tc0), luma_normal_ref(rb, alpha, beta, tc0);
            luma_normal(r[1], r[2], r[3], r[4], r[5], r[6], k);
        } else if (kind == 1) {
            luma_strong_ref(ra, alpha, beta), luma_strong_ref(rb, alpha, beta);
            luma_strong(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], k, alpha);
        } else {
            const int bs = kind == 2 ? 1 + rnd() % 3 : 4;
            chroma_ref(ra + 2, alpha, beta, bs, tc0 + 1), chroma_ref(rb + 2, alpha, beta, bs, tc0 + 1);
            chroma_edge2(r[2], r[3], r[4], r[5], k, bs == 4);
        }
        for (int i = 0; i < 8; i++)
            if ((int)(r[i] & 0xffff) != ra[i] || (int)(r[i] >> 16) != rb[i]) {
                printf("MISMATCH kind %d alpha %d beta %d tc0 %d tap %d: got (%u,%u) want (%d,%d)\n", kind, alpha, beta, tc0, i,
                       r[i] & 0xffff, r[i] >> 16, ra[i], rb[i]);
                bad++;
                break;
            }
    }
    printf("%s: %ld cases, %ld mismatches\n", bad ? "FAIL" : "OK", n, bad);
    return bad != 0;
}
