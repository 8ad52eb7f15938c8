// device_common.cuh -- shared device-side definitions of the B200 reconstruction engine.
//
// Frame pool layout (one pool per batch; every stream of a batch has the same geometry):
//   frame f = stream * numSlots + slot lives at pool + f * frameStride and holds three planes
//   stored WITH a replicated border so that the reference decoder's "clamp every coordinate into
//   the picture" fetch (h264bsdFillBlock, h264bsd_reconstruct.c:2244-2367) becomes a plain
//   rectangular TMA box load at a clamped origin:
//     Y : pitchY = W + 64 bytes, H + 64 rows, picture origin at (32, 32)
//     Cb: pitchC = align16(W/2 + 32) bytes, H/2 + 32 rows, picture origin at (16, 16);  Cr same
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "h264bsd_b200_tape.h"
#include "pool_geom.hpp"
#include "frame_addr.cuh"
#include "device_ptx.cuh"

namespace b200 {


// ---- inter-warp completion flags (one 32-bit word per stream x macroblock) -----------------------
// Watchdog: a wait that lasts implausibly long (seconds) records a code and gives up, so that a protocol bug
// shows up as a reported error (Batch::watchdog) instead of a hung GPU.
__device__ uint32_t gWatchdog[4];
constexpr unsigned long long kWatchdogNs = 4000000000ull;  // 4 s
__device__ __forceinline__ void waitFlag(const uint32_t *p, uint32_t serial) {
    unsigned ns = 20, spins = 0;
    unsigned long long t0 = 0;
    while (ldAcquire(p) != serial) {
        __nanosleep(ns);
        if (ns < 640) ns <<= 1;
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = globalTimerNs();
            if (!t0) t0 = now;
            else if (now - t0 > kWatchdogNs) { atomicAdd(&gWatchdog[0], 1u); break; }
        }
    }
}

// progress of a macroblock row: serial << 16 | macroblocks finished
__device__ __forceinline__ uint32_t waitRow(const uint32_t *p, uint32_t serial16, uint32_t need) {
    unsigned ns = 20, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        const uint32_t v = ldAcquire(p);
        if ((v >> 16) == serial16 && (v & 0xFFFFu) >= need) return v & 0xFFFFu;
        __nanosleep(ns);
        if (ns < 640) ns <<= 1;
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = globalTimerNs();
            if (!t0) t0 = now;
            else if (now - t0 > kWatchdogNs) { atomicAdd(&gWatchdog[0], 1u); return need; }
        }
    }
}

// ---- TMA / mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!mbarTryWait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            const unsigned long long now = globalTimerNs();
            if (!t0) t0 = now;
            else if (now - t0 > kWatchdogNs) { atomicAdd(&gWatchdog[1], 1u); break; }
        }
    }
}
// ---- constant tables ------------------------------------------------------------------------------
// 4x4 block order inside a macroblock (h264bsdBlockX/Y, intra_prediction.c:86-89), in 4-pel units
__device__ __constant__ uint8_t cBlkX[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
__device__ __constant__ uint8_t cBlkY[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
// raster 4x4 position -> block index (mb4x4Index, deblocking.c:124-125)
__device__ __constant__ uint8_t cRasterToBlk[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
__device__ __constant__ uint8_t cQpC[52] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                                            18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33,
                                            34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
__device__ __constant__ uint8_t cAlpha[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,   0,   0,   0,   4,   4,
                                              5,  6,  7,  8,  9,  10, 12, 13, 15, 17, 20, 22, 25,  28,  32,  36,  40,  45,
                                              50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
__device__ __constant__ uint8_t cBeta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  2,  2,  2,  3,  3,  3,  3,  4,  4,  4,
                                             6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
__device__ __constant__ uint8_t cTc0[52][4] = {
    {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0},
    {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 1, 0},
    {0, 0, 1, 0}, {0, 0, 1, 0}, {0, 0, 1, 0}, {0, 1, 1, 0}, {0, 1, 1, 0}, {1, 1, 1, 0}, {1, 1, 1, 0}, {1, 1, 1, 0}, {1, 1, 1, 0},
    {1, 1, 2, 0}, {1, 1, 2, 0}, {1, 1, 2, 0}, {1, 1, 2, 0}, {1, 2, 3, 0}, {1, 2, 3, 0}, {2, 2, 3, 0}, {2, 2, 4, 0}, {2, 3, 4, 0},
    {2, 3, 4, 0}, {3, 3, 5, 0}, {3, 4, 6, 0}, {3, 4, 6, 0}, {4, 5, 7, 0}, {4, 5, 8, 0}, {4, 6, 9, 0}, {5, 7, 10, 0}, {6, 8, 11, 0},
    {6, 8, 13, 0}, {7, 10, 14, 0}, {8, 11, 16, 0}, {9, 12, 18, 0}, {10, 13, 20, 0}, {11, 15, 23, 0}, {13, 17, 25, 0}};

}  // namespace b200
