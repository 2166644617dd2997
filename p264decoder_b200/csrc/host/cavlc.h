// CAVLC residual reader, see cavlc.cc
#pragma once
#include <cstdint>
#include "bitreader.h"

namespace p264b200 {
void cavlc_init();
int cavlc_read_block(BitReader &br, int nC, int max_coeff, int16_t *levels);

// coeff_token of an EMPTY block (TotalCoeff 0) per nC class [0..3] and for chroma DC [4]: about half of the residual
// blocks of a coded 8x8 are empty, and for those the whole block is this one short code
struct EmptyToken {
    uint8_t len, bits;
};
extern EmptyToken g_empty_token[5];
inline bool cavlc_skip_empty(BitReader &br, int nC)
{
    const EmptyToken t = g_empty_token[nC < 0 ? 4 : nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3];
    if (br.show(t.len) != t.bits) return false;
    br.skip(t.len);
    return true;
}
int cavlc_table_entry(int kind, int table, int sym, int *len, int *bits);
}  // namespace p264b200
