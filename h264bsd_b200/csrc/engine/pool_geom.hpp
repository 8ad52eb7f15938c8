// pool_geom.hpp -- frame-pool geometry and per-launch job descriptors shared by host and device code.
//
// Frame layout in HBM ("strips"): a plane is stored as vertical strips one macroblock wide, every strip row 16 bytes:
//     luma   strip s (s = macroblock column + 2), row r (r = y + 32):  16 pels            at  (s * rowsY + r) * 16
//     chroma strip s,                             row r (r = y/2 + 16): 8 Cb | 8 Cr pels   at  offC + (s * rowsC + r) * 16
// so a macroblock is 256 contiguous bytes of luma plus 128 contiguous bytes of chroma (whole 128-byte lines, nothing shared
// with a neighbour), vertically adjacent macroblocks follow each other in memory, and a column of n zero-motion macroblocks is
// one burst of 256 n + 128 n bytes.  The picture is surrounded by a replicated border of two macroblocks (32 luma pels) on
// every side, which turns the reference's per-coordinate clamp (h264bsdFillBlock, h264bsd_reconstruct.c:2244-2367) into a
// rectangular fetch at a clamped origin.
//
// A TMA tensor map over dims (x in strip: 16, strip, row, frame) with strides (1, rowsY * 16, 16, frameStride) -- the strip
// stride larger than the row stride; measured on B200 with tools/probe/strip_probe.cu -- delivers a box {16, nx, nr, 1} into
// shared memory as a RASTER window of pitch 16 nx: the de-stripping costs nothing, the box starts at any row, and because a
// box always starts at x = 0 of a strip the 16-byte start-alignment rule of tile loads is met by construction.
#pragma once
#include <cstdint>
#include "h264bsd_b200_tape.h"

namespace b200 {

constexpr int kPadY = 32;      // luma border in pels (two macroblocks)
constexpr int kPadC = 16;      // chroma border in pels
constexpr int kPadMbs = 2;     // border in strips / macroblocks

struct PoolGeom {
    int W, H;                 // luma size in pels (coded size)
    int widthMbs, heightMbs, nMbs;
    int strips;               // widthMbs + 4
    int rowsY, rowsC;         // rows of a strip incl. border: H + 64, H / 2 + 32
    unsigned long long offC;  // chroma strips inside a frame
    unsigned long long frameStride;
    int numSlots, nStreams;
    unsigned invWidthMbs;     // ceil(2^31 / widthMbs): mb / widthMbs == __umulhi(2 * mb + 1, invWidthMbs) for mb < 65536 and every
                              // width from 1 (see mbRowOf in frame_addr.cuh)
};

// what one stream contributes to one launch (one picture)
struct StreamJob {
    const b200_mb_rec *recs;  // nMbs records of this picture, raster order
    const int16_t *coefs;     // this picture's coefficient pool
    const uint16_t *orderE;   // the nE spatially concealed macroblocks in concealment order (b200_tape.mbOrder for this picture)
    uint16_t curSlot;
    uint16_t nB;              // intra-predicted macroblocks (0: the intra pass has nothing to do for this stream)
    uint16_t nE;
    uint16_t pad;
};

// launch parameters of the reconstruction kernels (recon_kernel.cuh, conceal_kernel.cuh)
struct ReconParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;     // nStreams
    uint32_t *done;            // nStreams * heightMbs row-progress words of pass B: serial << 16 | macroblocks finished
    uint32_t *ticket;          // CTA ticket counter of pass B (zeroed before launch)
    uint32_t *ticketA;         // ticket counters of the two pass-A instances (zeroed before launch)
    uint32_t *multiCount;      // pass A: entries in multiList (zeroed before launch)
    uint32_t *multiList;       // pass A: the picture's macroblocks with several partitions, stream << 16 | address (nStreams * nMbs entries)
    uint32_t *errors;          // [0] IDCT range errors (h264bsd_transform.c:183-188)
    uint32_t serial;           // value that marks "done in this launch"
    uint32_t chunkRows;        // pass A: macroblocks of one column per warp task (<= 32)
    uint32_t chunksPerCol;     // pass A: ceil(heightMbs / chunkRows)
    uint32_t totalChunks;      // pass A: chunksPerCol * widthMbs * nStreams
};

// geometry of a pool of nStreams x numSlots frames of widthMbs x heightMbs macroblocks
inline PoolGeom makePoolGeom(uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t nStreams) {
    PoolGeom g;
    g.widthMbs = (int)widthMbs; g.heightMbs = (int)heightMbs; g.nMbs = (int)(widthMbs * heightMbs);
    g.W = 16 * (int)widthMbs; g.H = 16 * (int)heightMbs;
    g.strips = (int)widthMbs + 2 * kPadMbs;
    g.rowsY = g.H + 2 * kPadY;
    g.rowsC = g.H / 2 + 2 * kPadC;
    g.offC = (unsigned long long)g.strips * g.rowsY * 16;
    g.frameStride = (g.offC + (unsigned long long)g.strips * g.rowsC * 16 + 255) & ~255ull;
    g.numSlots = (int)numSlots; g.nStreams = (int)nStreams;
    g.invWidthMbs = (unsigned)((0x80000000ull + widthMbs - 1) / widthMbs);
    return g;
}

static_assert(sizeof(StreamJob) == 32, "StreamJob is copied to the device as is");
static_assert(sizeof(b200_mb_rec) == B200_MB_REC_BYTES, "record layout");

}  // namespace b200
