// CAVLC residual block reader (H.264 clause 9.2), host side.
// Replaces p264dec_read_residual_block_cavlc (decoder/dec_cavlc.c:1371-1524): same
// outputs (levels in zig-zag scan order, total_coeff), built from the standard's
// code tables (Tables 9-5, 9-7, 9-8, 9-9, 9-10) expanded into flat prefix lookups
// at start-up instead of the reference's hand-split per-table readers.
#include "cavlc.h"

#include <cstring>
#include <mutex>

namespace p264b200 {

namespace {

// Table 9-5 coeff_token, indexed [nC class][4*total_coeff + trailing_ones]: code length / code word.
const uint8_t kTokLen[4][68] = {
    {1,  0,  0,  0,  6,  2,  0,  0,  8,  6,  3,  0,  9,  8,  7,  5,  10, 9,  8,  6,  11, 10, 9,
     7,  13, 11, 10, 8,  13, 13, 11, 9,  13, 13, 13, 10, 14, 14, 13, 11, 14, 14, 14, 13, 15, 15,
     14, 14, 15, 15, 15, 14, 16, 15, 15, 15, 16, 16, 16, 15, 16, 16, 16, 16, 16, 16, 16, 16},
    {2,  0,  0,  0,  6,  2,  0,  0,  6,  5,  3,  0,  7,  6,  6,  4,  8,  6,  6,  4,  8,  7,  7,
     5,  9,  8,  8,  6,  11, 9,  9,  6,  11, 11, 11, 7,  12, 11, 11, 9,  12, 12, 12, 11, 12, 12,
     12, 11, 13, 13, 13, 12, 13, 13, 13, 13, 13, 14, 13, 13, 14, 14, 14, 13, 14, 14, 14, 14},
    {4,  0,  0,  0,  6,  4,  0,  0,  6,  5,  4,  0,  6,  5,  5,  4,  7,  5,  5,  4,  7,  5,  5,
     4,  7,  6,  6,  4,  7,  6,  6,  4,  8,  7,  7,  5,  8,  8,  7,  6,  9,  8,  8,  7,  9,  9,
     8,  8,  9,  9,  9,  8,  10, 9,  9,  9,  10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10},
    {6, 0, 0, 0, 6, 6, 0, 0, 6, 6, 6, 0, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
     6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6}};
const uint8_t kTokBits[4][68] = {
    {1,  0,  0,  0,  5,  1,  0,  0,  7,  4,  1,  0,  7,  6,  5,  3,  7,  6,  5,  3,  7,  6,  5,
     4,  15, 6,  5,  4,  11, 14, 5,  4,  8,  10, 13, 4,  15, 14, 9,  4,  11, 10, 13, 12, 15, 14,
     9,  12, 11, 10, 13, 8,  15, 1,  9,  12, 11, 14, 13, 8,  7,  10, 9,  12, 4,  6,  5,  8},
    {3,  0,  0,  0,  11, 2,  0,  0,  7,  7,  3,  0,  7,  10, 9,  5,  7,  6,  5,  4,  4,  6,  5,
     6,  7,  6,  5,  8,  15, 6,  5,  4,  11, 14, 13, 4,  15, 10, 9,  4,  11, 14, 13, 12, 8,  10,
     9,  8,  15, 14, 13, 12, 11, 10, 9,  12, 7,  11, 6,  8,  9,  8,  10, 1,  7,  6,  5,  4},
    {15, 0,  0,  0,  15, 14, 0,  0,  11, 15, 13, 0,  8,  12, 14, 12, 15, 10, 11, 11, 11, 8,  9,
     10, 9,  14, 13, 9,  8,  10, 9,  8,  15, 14, 13, 13, 11, 14, 10, 12, 15, 10, 13, 12, 11, 14,
     9,  12, 8,  10, 13, 8,  13, 7,  9,  12, 9,  12, 11, 10, 5,  8,  7,  6,  1,  4,  3,  2},
    {3,  0,  0,  0,  0,  1,  0,  0,  4,  5,  6,  0,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18,
     19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41,
     42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63}};
// nC == -1 (chroma DC), [4*total_coeff + trailing_ones]
const uint8_t kTokDcLen[20] = {2, 0, 0, 0, 6, 1, 0, 0, 6, 6, 3, 0, 6, 7, 7, 6, 6, 8, 8, 7};
const uint8_t kTokDcBits[20] = {1, 0, 0, 0, 7, 1, 0, 0, 4, 6, 1, 0, 3, 3, 2, 5, 2, 3, 2, 0};

// Tables 9-7 / 9-8 total_zeros for 4x4 blocks, [total_coeff-1][total_zeros]
const uint8_t kTzLen[15][16] = {{1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9},
                                {3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6},
                                {4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6},
                                {5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5},
                                {4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5},
                                {6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6},
                                {6, 5, 3, 3, 3, 2, 3, 4, 3, 6},
                                {6, 4, 5, 3, 2, 2, 3, 3, 6},
                                {6, 6, 4, 2, 2, 3, 2, 5},
                                {5, 5, 3, 2, 2, 2, 4},
                                {4, 4, 3, 3, 1, 3},
                                {4, 4, 2, 1, 3},
                                {3, 3, 1, 2},
                                {2, 2, 1},
                                {1, 1}};
const uint8_t kTzBits[15][16] = {{1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1},
                                 {7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0},
                                 {5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0},
                                 {3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0},
                                 {5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0},
                                 {1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0},
                                 {1, 1, 5, 4, 3, 3, 2, 1, 1, 0},
                                 {1, 1, 1, 3, 3, 2, 2, 1, 0},
                                 {1, 0, 1, 3, 2, 1, 1, 1},
                                 {1, 0, 1, 3, 2, 1, 1},
                                 {0, 1, 1, 2, 1, 3},
                                 {0, 1, 1, 1, 1},
                                 {0, 1, 1, 1},
                                 {0, 1, 1},
                                 {0, 1}};
const uint8_t kTzCount[15] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
// Table 9-9 total_zeros for chroma DC 2x2, [total_coeff-1][total_zeros]
const uint8_t kTzDcLen[3][4] = {{1, 2, 3, 3}, {1, 2, 2, 0}, {1, 1, 0, 0}};
const uint8_t kTzDcBits[3][4] = {{1, 1, 1, 0}, {1, 1, 0, 0}, {1, 0, 0, 0}};
// Table 9-10 run_before, [min(zeros_left,7)-1][run_before]
const uint8_t kRunLen[7][15] = {{1, 1},
                                {1, 2, 2},
                                {2, 2, 2, 2},
                                {2, 2, 2, 3, 3},
                                {2, 2, 3, 3, 3, 3},
                                {2, 3, 3, 3, 3, 3, 3},
                                {3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11}};
const uint8_t kRunBits[7][15] = {{1, 0},
                                 {1, 1, 0},
                                 {3, 2, 1, 0},
                                 {3, 2, 1, 1, 0},
                                 {3, 2, 3, 2, 1, 0},
                                 {3, 0, 1, 3, 2, 5, 4},
                                 {7, 6, 5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1}};
const uint8_t kRunCount[7] = {2, 3, 4, 5, 6, 7, 15};

// flat prefix lookups: index = next kBits bits, value = (symbol << 8) | length, 0 = invalid
struct Lut {
    int bits;
    uint16_t *tab;
};
uint16_t g_tok[4][1 << 16];
uint16_t g_tok8[4][1 << 8];   // codes of length <= 8 only
uint16_t g_tok_dc[1 << 8];
uint16_t g_tz[15][1 << 9];
uint16_t g_tz_dc[3][1 << 3];
uint16_t g_run[7][1 << 11];
std::once_flag g_once;

void fill(uint16_t *tab, int tab_bits, int sym, int len, int code)
{
    if (len == 0) return;
    const int lo = code << (tab_bits - len), n = 1 << (tab_bits - len);
    for (int i = 0; i < n; i++) tab[lo + i] = (uint16_t)((sym << 8) | len);
}

void build()
{
    for (int t = 0; t < 4; t++)
        for (int s = 0; s < 68; s++)
            if ((s & 3) <= (s >> 2)) {
                fill(g_tok[t], 16, s, kTokLen[t][s], kTokBits[t][s]);
                if (kTokLen[t][s] <= 8) fill(g_tok8[t], 8, s, kTokLen[t][s], kTokBits[t][s]);
            }
    for (int s = 0; s < 20; s++)
        if ((s & 3) <= (s >> 2)) fill(g_tok_dc, 8, s, kTokDcLen[s], kTokDcBits[s]);
    for (int t = 0; t < 15; t++)
        for (int s = 0; s < kTzCount[t]; s++) fill(g_tz[t], 9, s, kTzLen[t][s], kTzBits[t][s]);
    for (int t = 0; t < 3; t++)
        for (int s = 0; s < 4 - t; s++) fill(g_tz_dc[t], 3, s, kTzDcLen[t][s], kTzDcBits[t][s]);
    for (int t = 0; t < 7; t++)
        for (int s = 0; s < kRunCount[t]; s++) fill(g_run[t], 11, s, kRunLen[t][s], kRunBits[t][s]);
}

inline int lookup(BitReader &br, const uint16_t *tab, int bits)
{
    const uint16_t e = tab[br.show(bits)];
    if (!e) return -1;
    br.skip(e & 0xff);
    return e >> 8;
}

}  // namespace

EmptyToken g_empty_token[5];

void cavlc_init()
{
    std::call_once(g_once, [] {
        build();
        for (int t = 0; t < 4; t++) g_empty_token[t] = EmptyToken{(uint8_t)kTokLen[t][0], (uint8_t)kTokBits[t][0]};
        g_empty_token[4] = EmptyToken{(uint8_t)kTokDcLen[0], (uint8_t)kTokDcBits[0]};
    });
}

// raw table access for the table self-check (tests compare these with the standard's code words)
int cavlc_table_entry(int kind, int table, int sym, int *len, int *bits)
{
    switch (kind) {
    case 0:
        if (table < 0 || table > 3 || sym < 0 || sym >= 68) return -1;
        *len = kTokLen[table][sym], *bits = kTokBits[table][sym];
        return 0;
    case 1:
        if (sym < 0 || sym >= 20) return -1;
        *len = kTokDcLen[sym], *bits = kTokDcBits[sym];
        return 0;
    case 2:
        if (table < 0 || table > 14 || sym < 0 || sym >= kTzCount[table]) return -1;
        *len = kTzLen[table][sym], *bits = kTzBits[table][sym];
        return 0;
    case 3:
        if (table < 0 || table > 2 || sym < 0 || sym >= 4 - table) return -1;
        *len = kTzDcLen[table][sym], *bits = kTzDcBits[table][sym];
        return 0;
    case 4:
        if (table < 0 || table > 6 || sym < 0 || sym >= kRunCount[table]) return -1;
        *len = kRunLen[table][sym], *bits = kRunBits[table][sym];
        return 0;
    }
    return -1;
}

// nC: -1 chroma DC, else predicted non-zero count (0..16).  max_coeff: 16, 15 or 4.
// levels[0..max_coeff-1] receive the coefficients in scan order (zeros elsewhere).
// Returns total_coeff or -1 on an invalid code.
int cavlc_read_block(BitReader &br, int nC, int max_coeff, int16_t *levels)
{
    int tok;
    if (nC < 0)
        tok = lookup(br, g_tok_dc, 8);
    else {
        // nC class boundaries of Table 9-5 (decoder/dec_cavlc.c:1389-1400 uses the same split)
        const int cls = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
        // codes of up to 8 bits (the common ones) from a 512-byte table that stays in L1; the 128 KB flat table only for the rest
        const uint16_t e = g_tok8[cls][br.show(8)];
        if (e) {
            br.skip(e & 0xff);
            tok = e >> 8;
        } else
            tok = lookup(br, g_tok[cls], 16);
    }
    if (tok < 0) return -1;
    const int total = tok >> 2, t1s = tok & 3;
    if (total == 0) return 0;
    if (total > max_coeff) return -1;

    int lev[16];
    int suffix_len = (total > 10 && t1s < 3) ? 1 : 0;
    int i = 0;
    for (; i < t1s; i++) lev[i] = 1 - 2 * (int)br.read1();
    for (; i < total; i++) {
        int prefix = 0;
        const uint32_t w = br.show(25);
        if (w >> 9) {  // at most 15 zeros before the 1: the whole prefix sits in the window
            prefix = __builtin_clz(w) - 7;
            br.skip(prefix + 1);
        } else
            while (br.read1() == 0) {
                if (++prefix > 32 || br.eof()) return -1;
            }
        int suffix_size = suffix_len;
        if (prefix == 14 && suffix_len == 0) suffix_size = 4;
        if (prefix >= 15) suffix_size = prefix - 3;
        int code = ((prefix < 15 ? prefix : 15) << suffix_len) + (suffix_size ? (int)br.read(suffix_size) : 0);
        if (prefix >= 15 && suffix_len == 0) code += 15;
        if (prefix >= 16) code += (1 << (prefix - 3)) - 4096;
        if (i == t1s && t1s < 3) code += 2;
        // dec_cavlc.c:1376 keeps levels in int16_t before widening them again
        lev[i] = (int16_t)((code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1);
        if (suffix_len == 0) suffix_len = 1;
        const int a = lev[i] < 0 ? -lev[i] : lev[i];
        if (a > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }

    int zeros_left = 0;
    if (total < max_coeff) {
        zeros_left = (nC < 0) ? lookup(br, g_tz_dc[total - 1], 3) : lookup(br, g_tz[total - 1], 9);
        if (zeros_left < 0) return -1;
    }
    int run[16];
    for (i = 0; i < total - 1; i++) {
        if (zeros_left > 0) {
            const int r = lookup(br, g_run[(zeros_left > 7 ? 7 : zeros_left) - 1], 11);
            if (r < 0 || r > zeros_left) return -1;
            run[i] = r;
        } else
            run[i] = 0;
        zeros_left -= run[i];
    }
    run[total - 1] = zeros_left;

    int pos = -1;
    for (i = total - 1; i >= 0; i--) {
        pos += run[i] + 1;
        if (pos >= max_coeff) return -1;
        levels[pos] = (int16_t)lev[i];
    }
    return total;
}

}  // namespace p264b200
