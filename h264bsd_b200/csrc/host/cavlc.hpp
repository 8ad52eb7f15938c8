// cavlc.hpp -- CAVLC residual block decoding (ITU-T H.264 clause 9.2) for the host parser.
//
// Replaces h264bsdDecodeResidualBlockCavlc (h264bsd_cavlc.c:749-916).  The code tables are
// the standard's Tables 9-5, 9-7, 9-8, 9-9 and 9-10 written as (length, code) pairs and
// expanded into prefix look-up tables at start-up; tests/test_cavlc_tables.py cross-checks
// every prefix against the reference's own look-up arrays.
#pragma once
#include <cstdint>
#include "bits.hpp"

namespace b200 {

struct CavlcResult {
    int totalCoeff;      // number of non-zero levels, <0 on error
    uint32_t coeffMap;   // bit i set: zig-zag position i (relative to the block's first coded position) non-zero
    uint32_t sumAbs;     // sum of the magnitudes of the levels (bounds what the inverse transform can produce)
};

void cavlcInit();

// nC: >=0 luma/chroma-AC context, -1 chroma DC.  maxNumCoeff 16, 15 or 4.
// Writes levels (zig-zag order) into out[0..maxNumCoeff-1]; `out` must be pre-zeroed.
CavlcResult cavlcResidualBlock(BitReader &br, int16_t *out, int nC, int maxNumCoeff);

}  // namespace b200
