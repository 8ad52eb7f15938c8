// pool_geom.hpp -- frame-pool geometry and per-launch job descriptors shared by host and device code
// (layout rationale in device_common.cuh)
#pragma once
#include <cstdint>
#include "h264bsd_b200_tape.h"

namespace b200 {

constexpr int kPadY = 32;
constexpr int kPadC = 16;
// TMA tile loads need the innermost start coordinate 16-byte aligned (measured on B200: an unaligned
// start raises an illegal-instruction fault), so a box starts at the window origin rounded DOWN to 16 and is
// 15 bytes wider than the window: luma (16+5)+15 -> 48 x 21, chroma (8+1)+15 -> 32 x 9 x 2 planes.
constexpr int kLumaWin = 21, kChromaWin = 9;
constexpr int kLumaBoxW = 48, kLumaBoxH = 21;
constexpr int kChromaBoxW = 32, kChromaBoxH = 9;

struct PoolGeom {
    int W, H;                 // luma size in pels (coded size)
    int widthMbs, heightMbs, nMbs;
    int pitchY, pitchC;
    int rowsY, rowsC;         // rows incl. border
    unsigned long long offCb, offCr;   // plane offsets inside a frame
    unsigned long long frameStride;
    int numSlots, nStreams;
    unsigned invWidthMbs;     // ceil(2^31 / widthMbs): mb / widthMbs == __umulhi(2 * mb + 1, invWidthMbs) for mb < 65536 and every
                              // width from 1 (see mbRowOf in device_common.cuh)
};

// what one stream contributes to one launch (one picture)
struct StreamJob {
    const b200_mb_rec *recs;  // nMbs records of this picture
    const int16_t *coefs;     // this picture's coefficient pool
    const uint16_t *order;    // list entries: nR zero-motion runs (two entries each: first address, length), nC single plain
                              // copies, nA other pass-A macroblocks, nB pass-B macroblocks in wavefront order (b200_tape.mbOrder)
    uint16_t curSlot;
    uint16_t nR, nC, nA, nB;
    uint16_t nE;              // spatially concealed macroblocks: after the nB entries, concealment order
    uint16_t pad[2];
};

static_assert(sizeof(StreamJob) == 40, "StreamJob is copied to the device as is");
static_assert(sizeof(b200_mb_rec) == B200_MB_REC_BYTES, "record layout");

}  // namespace b200
