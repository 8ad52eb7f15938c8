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

// launch parameters of the reconstruction kernels (recon_kernel.cuh, conceal_kernel.cuh)
struct ReconParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;     // nStreams
    uint32_t *done;            // nStreams * nMbs completion flags (pass B only)
    uint32_t *ticket;          // CTA ticket counter of pass B (zeroed before launch)
    uint32_t *errors;          // [0] IDCT range errors (h264bsd_transform.c:183-188)
    uint32_t serial;           // value that marks "done in this launch"
    uint32_t chunksB;          // pass B: warp tasks (chunkB entries) per stream
    uint32_t chunkB;           // pass B: list entries per warp task
    uint32_t chunkA;           // pass A: list entries per warp (<= kChunkA)
    uint32_t copyRuns;         // copy pass: runs per warp task (<= kCopyRunsPerTask)
    uint32_t chunksA;          // pass A: virtual CTAs per stream (kReconWarps * kChunkA entries each)
    uint32_t virtualCtasA;     // chunksA * nStreams
    uint32_t chunksC;          // copy pass: warp tasks of 32 single copies per stream
    uint32_t chunksQ;          // copy pass: warp tasks of copyRuns zero-motion runs per stream (they come first)
};

// geometry of a pool of nStreams x numSlots frames of widthMbs x heightMbs macroblocks (layout in device_common.cuh)
inline PoolGeom makePoolGeom(uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t nStreams) {
    PoolGeom g;
    g.widthMbs = (int)widthMbs; g.heightMbs = (int)heightMbs; g.nMbs = (int)(widthMbs * heightMbs);
    g.W = 16 * (int)widthMbs; g.H = 16 * (int)heightMbs;
    g.pitchY = g.W + 2 * kPadY;
    g.pitchC = (g.W / 2 + 2 * kPadC + 15) & ~15;
    g.rowsY = g.H + 2 * kPadY;
    g.rowsC = g.H / 2 + 2 * kPadC;
    g.offCb = (unsigned long long)g.pitchY * g.rowsY;
    g.offCr = g.offCb + (unsigned long long)g.pitchC * g.rowsC;
    g.frameStride = (g.offCr + (unsigned long long)g.pitchC * g.rowsC + 255) & ~255ull;
    g.numSlots = (int)numSlots; g.nStreams = (int)nStreams;
    g.invWidthMbs = (unsigned)((0x80000000ull + widthMbs - 1) / widthMbs);
    return g;
}

static_assert(sizeof(StreamJob) == 40, "StreamJob is copied to the device as is");
static_assert(sizeof(b200_mb_rec) == B200_MB_REC_BYTES, "record layout");

}  // namespace b200
