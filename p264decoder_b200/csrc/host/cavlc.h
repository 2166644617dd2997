// CAVLC residual reader, see cavlc.cc
#pragma once
#include <cstdint>
#include "bitreader.h"

namespace p264b200 {
void cavlc_init();
int cavlc_read_block(BitReader &br, int nC, int max_coeff, int16_t *levels);
int cavlc_table_entry(int kind, int table, int sym, int *len, int *bits);
}  // namespace p264b200
