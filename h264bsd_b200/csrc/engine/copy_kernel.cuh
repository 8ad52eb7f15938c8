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

__global__ void __launch_bounds__(kCopyWarps * 32) reconCopyKernel(const ReconParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cp = lane >> 4, cr = (lane >> 1) & 7, cc = (lane & 1) * 4;
    const uint32_t tasksPerStream = p.chunksQ + p.chunksC;
    const uint32_t totalTasks = tasksPerStream * (uint32_t)g.nStreams;
    for (uint32_t t = blockIdx.x * kCopyWarps + warp; t < totalTasks; t += gridDim.x * kCopyWarps) {
        const uint32_t s = t / tasksPerStream, task = t - s * tasksPerStream;
        const StreamJob job = p.jobs[s];
        const uint32_t frameBase = s * (uint32_t)g.numSlots;
        uint8_t *cur = framePtr(p.pool, g, frameBase + job.curSlot);
        if (task < p.chunksQ) {
            // ---- runs: 2..32 macroblocks side by side, zero vector, one reference frame.  A run is 16 luma rows of 16 * len
            // bytes and 2 x 8 chroma rows of 8 * len bytes at the same offset in the reference and the current frame; the warp
            // walks them as 16-byte (8-byte) chunks in row-major order, so every row segment is one contiguous burst
            const uint32_t e0 = task * p.copyRuns;
            if (e0 >= job.nR) continue;
            const int n = (int)min(p.copyRuns, (uint32_t)job.nR - e0);
            uint32_t mOff = 0, mOffC = 0, mLen = 1;
            long long mDelta = 0;   // reference frame - current frame
            if (lane < n) {
                const uint32_t mb = __ldg(job.order + 2u * (e0 + lane));   // (address, length) pairs
                mLen = __ldg(job.order + 2u * (e0 + lane) + 1u);
                const uint32_t refSlots = __ldg(reinterpret_cast<const uint32_t *>(job.recs + mb) + 4);
                const int mby = mbRowOf(mb, g), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
                mOff = (uint32_t)((mby * 16 + kPadY) * g.pitchY + mbx * 16 + kPadY);
                mOffC = (uint32_t)((mby * 8 + kPadC) * g.pitchC + mbx * 8 + kPadC);
                mDelta = ((long long)(refSlots & 0xFF) - (long long)job.curSlot) * (long long)g.frameStride;
            }
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                const uint32_t len = __shfl_sync(0xffffffffu, mLen, i);
                uint8_t *dY = cur + __shfl_sync(0xffffffffu, mOff, i);
                uint8_t *dC = cur + g.offCb + __shfl_sync(0xffffffffu, mOffC, i);
                const long long delta = __shfl_sync(0xffffffffu, mDelta, i);
                const uint32_t inv = 65536u / len + 1u;        // chunk / len == (chunk * inv) >> 16 for chunk < 512, len <= 32
                const uint32_t chunks = 16u * len;
#pragma unroll 1
                for (uint32_t c0 = 0; c0 < chunks; c0 += 64) {
                    uint4 a[2];
                    uint2 b[2];
                    size_t oy[2], oc[2];
                    bool ok[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const uint32_t c = c0 + 32u * u + lane;
                        ok[u] = c < chunks;
                        const uint32_t row = (c * inv) >> 16, col = c - row * len;
                        oy[u] = (size_t)row * g.pitchY + col * 16;
                        oc[u] = (row >> 3) * (size_t)(g.offCr - g.offCb) + (size_t)(row & 7) * g.pitchC + col * 8;
                        if (ok[u]) {
                            a[u] = __ldg(reinterpret_cast<const uint4 *>(dY + delta + oy[u]));
                            b[u] = __ldg(reinterpret_cast<const uint2 *>(dC + delta + oc[u]));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++)
                        if (ok[u]) {
                            *reinterpret_cast<uint4 *>(dY + oy[u]) = a[u];
                            *reinterpret_cast<uint2 *>(dC + oc[u]) = b[u];
                        }
                }
            }
            continue;
        }
        const uint32_t e0 = (task - p.chunksQ) * 32u;
        if (e0 >= job.nC) continue;
        const int n = (int)min(32u, (uint32_t)job.nC - e0);
        // lane j: where entry j's source lies (clamped like issueWindow: a block wholly outside the picture on an axis equals
        // the block at the clamped origin because the border is a replication)
        uint32_t mMb = 0;
        unsigned long long mSrcY = 0, mSrcC = 0;
        if (lane < n) {
            mMb = __ldg(job.order + 2u * job.nR + e0 + lane);
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(job.recs + mMb);
            const uint32_t refSlots = __ldg(rw + 4), mvv = __ldg(rw + 8);
            const int mvx = (int)(int16_t)(mvv & 0xFFFF), mvy = (int)(int16_t)(mvv >> 16);
            const int mby = mbRowOf(mMb, g), mbx = (int)(mMb - (uint32_t)mby * g.widthMbs);
            const int x = clip3(-kPadY, g.W + kPadY - 16, mbx * 16 + (mvx >> 2)), y = clip3(-kPadY, g.H + kPadY - 16, mby * 16 + (mvy >> 2));
            const int cx = clip3(-kPadC, g.W / 2 + kPadC - 8, mbx * 8 + (mvx >> 3)), cy = clip3(-kPadC, g.H / 2 + kPadC - 8, mby * 8 + (mvy >> 3));
            const unsigned long long ref = (unsigned long long)(frameBase + (refSlots & 0xFF)) * g.frameStride;
            mSrcY = ref + (unsigned long long)(y + kPadY) * g.pitchY + (x + kPadY);
            mSrcC = ref + g.offCb + (unsigned long long)(cy + kPadC) * g.pitchC + (cx + kPadC);
        }
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += kCopyUnroll) {
            uint2 pv[kCopyUnroll];
            uint32_t pc[kCopyUnroll];
            uint32_t mbs[kCopyUnroll];
#pragma unroll
            for (int u = 0; u < kCopyUnroll; u++) {
                const int i = min(i0 + u, n - 1);   // a short tail repeats the last entry (same bytes, same place)
                mbs[u] = __shfl_sync(0xffffffffu, mMb, i);
                const unsigned long long sy = __shfl_sync(0xffffffffu, mSrcY, i), sc = __shfl_sync(0xffffffffu, mSrcC, i);
                const uint8_t *srcY = p.pool + sy + (size_t)r8 * g.pitchY + c8;
                const uint8_t *srcC = p.pool + sc + (cp ? g.offCr - g.offCb : 0ull) + (size_t)cr * g.pitchC + cc;
                // the vector is a multiple of two luma pels / one chroma pel: 2-byte aligned luma, 1-byte aligned chroma
                const uint32_t ay = (uint32_t)(sy + c8) & 3u, ac = (uint32_t)(sc + cc) & 3u;
                const uint32_t *wy = reinterpret_cast<const uint32_t *>(srcY - ay);
                const uint32_t *wc = reinterpret_cast<const uint32_t *>(srcC - ac);
                if (ay == 0) {                       // warp-uniform: every lane has the same vector
                    pv[u] = make_uint2(__ldg(wy), __ldg(wy + 1));
                } else {
                    const uint32_t w0 = __ldg(wy), w1 = __ldg(wy + 1), w2 = __ldg(wy + 2);
                    pv[u] = make_uint2(__funnelshift_r(w0, w1, ay * 8), __funnelshift_r(w1, w2, ay * 8));
                }
                if (ac == 0) pc[u] = __ldg(wc);
                else pc[u] = __funnelshift_r(__ldg(wc), __ldg(wc + 1), ac * 8);
            }
#pragma unroll
            for (int u = 0; u < kCopyUnroll; u++) {
                const uint32_t mb = mbs[u];
                const int mby = mbRowOf(mb, g), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
                *reinterpret_cast<uint2 *>(lumaAt(cur, g, mbx * 16 + c8, mby * 16 + r8)) = pv[u];
                *reinterpret_cast<uint32_t *>(chromaAt(cur, g, cp, mbx * 8 + cc, mby * 8 + cr)) = pc[u];
            }
        }
    }
}

}  // namespace b200
