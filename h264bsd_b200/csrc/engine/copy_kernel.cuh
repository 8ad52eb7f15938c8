// copy_kernel.cuh -- the copy pass of the reconstruction (plain-copy macroblocks, zero-motion runs).
// Like conceal_kernel.cuh it needs nothing but the frame addressing helpers, shuffles and plain / read-only loads, so the
// same source also compiles for the host (tests/emu/).
#pragma once
#include "frame_addr.cuh"

namespace b200 {

// =====================================================================================================
// copy pass: the macroblocks the host classified as plain copies (P_Skip / P_L0_16x16, no residual, vector integer
// for luma and chroma -- two thirds of a typical P picture).  h264bsdPredictSamples degenerates to h264bsdFillBlock
// (reconstruct.c:1852, :2244) and h264bsdWriteOutputBlocks to a store: 384 bytes in, 384 bytes out, no shared memory.
// Horizontal runs of zero-vector copies are moved as whole row segments; the other plain copies one macroblock at a time: a
// warp owns 32 consecutive list entries, lane j fetches entry j's address, reference slot and vector (one dependent-load
// chain per 32 macroblocks), then the warp copies four macroblocks per step, loads before stores.
// =====================================================================================================
constexpr int kCopyWarps = 8;
constexpr int kCopyUnroll = 4;
constexpr int kCopyRunsPerTask = 16;   // most zero-motion runs per warp task (ReconParams::copyRuns)

#define B200_COPY_NAME reconCopyKernel
#define B200_COPY_BOUNDS __launch_bounds__(kCopyWarps * 32)
#define B200_COPY_STEPS 2
#include "copy_kernel_body.inc"
#undef B200_COPY_NAME
#undef B200_COPY_BOUNDS
#undef B200_COPY_STEPS

// Two unmeasured variants for an A/B run (B200_COPY_VARIANT=1 / 2, Batch::create; the kernel above is the default and the one
// every number in DESIGN.md was taken with).  The copy pass is latency-bound with too few bytes in flight:
// variant 1 trades registers for a fourth resident CTA per SM (ptxas: 64 registers, 28 bytes of spills), variant 2 issues the
// loads of four steps instead of two before the first store (80 registers).
#define B200_COPY_NAME reconCopyKernelOcc4
#define B200_COPY_BOUNDS __launch_bounds__(kCopyWarps * 32, 4)
#define B200_COPY_STEPS 2
#include "copy_kernel_body.inc"
#undef B200_COPY_NAME
#undef B200_COPY_BOUNDS
#undef B200_COPY_STEPS

#define B200_COPY_NAME reconCopyKernelDeep
#define B200_COPY_BOUNDS __launch_bounds__(kCopyWarps * 32)
#define B200_COPY_STEPS 4
#include "copy_kernel_body.inc"
#undef B200_COPY_NAME
#undef B200_COPY_BOUNDS
#undef B200_COPY_STEPS

}  // namespace b200
