// copy_bulk_kernel.cuh -- the zero-motion runs of the copy pass moved by the bulk-copy engine instead of through registers.
// EXPERIMENTAL, off by default (B200_COPY_BULK=1 selects it; Batch::launchPicture): written after round 1's GPU time was spent,
// checked on the host by emulation (tests/emu/) only, never timed.  Why it exists: reconCopyKernel keeps two steps of 384 bytes
// per warp in flight (18 KB per SM at 24 resident warps) where 6.5 TB/s at ~800 ns wants about 35 KB per SM -- it reaches 32 % of
// the HBM peak, latency-bound.  Registers cannot hold more; shared memory can: here a warp keeps
// kBulkAhead + 1 pieces of up to 16 macroblocks (6 KB each) in flight without a register, rows as whole bursts.
//
// A run (reference h264bsdPredictSamples -> h264bsdFillBlock, reconstruct.c:1852, :2244, and h264bsdWriteOutputBlocks,
// image.c:81-344, for a P_Skip / P_L0_16x16 macroblock without residual and with a zero vector) is the same rectangle in the
// reference frame and in the current frame: 16 luma rows of 16 * len bytes, 2 x 8 chroma rows of 8 * len bytes.  It is cut into
// pieces of at most kBulkPieceMbs macroblocks; of a piece, lane r < 16 moves luma row r and lane 16 + 8 p + r row r of chroma
// plane p with one cp.async.bulk global -> shared (completion on the piece's mbarrier) and one shared -> global (bulk group).
// cp.async.bulk wants 16-byte aligned addresses and sizes: luma rows always are; chroma rows only from an even macroblock
// column to an even one, so an odd first / last macroblock's chroma (8 bytes per row) goes through a register of lanes
// 0..15 / 16..31.  The single plain copies (non-zero vectors, unaligned sources) stay with reconCopyKernel.
#pragma once
#include "frame_addr.cuh"
#include "device_ptx.cuh"
#include "device_common.cuh"

#ifndef B200_DYNAMIC_SMEM
#define B200_DYNAMIC_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#endif

namespace b200 {

constexpr int kBulkWarps = 4;
constexpr int kBulkBufs = 4;         // staging buffers per warp
constexpr int kBulkAhead = 2;        // pieces whose loads are issued ahead of the piece being stored (< kBulkBufs - 1, see below)
constexpr int kBulkPieceMbs = 16;    // macroblocks per piece at most: a run of 17..32 is two pieces
constexpr int kBulkRunsPerTask = 16; // runs per warp task (a task's pieces must fit 32 lanes' worth of metadata: 2 per run)
constexpr int kBulkBufBytes = 384 * kBulkPieceMbs;

struct BulkWarpSmem {
    __align__(128) uint8_t buf[kBulkBufs][kBulkBufBytes];
    __align__(8) uint64_t bar[kBulkBufs];
};

// what a lane moves of one piece
struct BulkPiece {
    uint8_t *dst;          // current frame: this lane's row at the piece's first (luma) / first even (chroma) macroblock
    long long delta;       // reference frame - current frame
    uint32_t bytes;        // of this lane's bulk row (0: nothing, e.g. a chroma row of a one-macroblock piece)
    uint32_t smemOff;      // where the row is staged
    uint32_t total;        // bytes of all 32 rows (what the mbarrier expects)
    uint8_t *edge;         // current frame: 8 chroma bytes of an odd first (lanes 0..15) / last (16..31) macroblock, or nullptr
};

__global__ void __launch_bounds__(kBulkWarps * 32) reconCopyBulkKernel(const ReconParams p) {
    B200_DYNAMIC_SMEM(bulkSmemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BulkWarpSmem &sm = reinterpret_cast<BulkWarpSmem *>(bulkSmemRaw)[warp];
    const PoolGeom &g = p.g;
    if (lane == 0) {
        for (int b = 0; b < kBulkBufs; b++) mbarInit(&sm.bar[b], 1);
        fenceMbarInit();
    }
    __syncwarp();
    const uint32_t runsPerTask = min(p.copyRuns, (uint32_t)kBulkRunsPerTask);
    const uint32_t totalTasks = p.chunksQ * (uint32_t)g.nStreams;
    uint32_t q = 0;   // pieces this warp has issued so far: piece q is staged in buffer q % kBulkBufs, phase (q / kBulkBufs) & 1
    uint32_t w = 0;   // pieces this warp has stored so far
    bool dead = false;
    for (uint32_t t = blockIdx.x * kBulkWarps + warp; t < totalTasks; t += gridDim.x * kBulkWarps) {
        const uint32_t s = t / p.chunksQ, task = t - s * p.chunksQ;
        const StreamJob job = p.jobs[s];
        const uint32_t e0 = task * runsPerTask;
        if (e0 >= job.nR) continue;
        const int n = (int)min(runsPerTask, (uint32_t)job.nR - e0);
        uint8_t *cur = framePtr(p.pool, g, s * (uint32_t)g.numSlots + job.curSlot);
        // lane j: run j of the task (address, length, reference frame) and where its pieces start in the task's piece sequence
        uint32_t mX = 0, mRowY = 0, mRowC = 0, mLen = 0, mPieces = 0;
        long long mDelta = 0;
        if (lane < n) {
            const uint32_t mb = __ldg(job.order + 2u * (e0 + lane));
            mLen = __ldg(job.order + 2u * (e0 + lane) + 1u);
            const uint32_t refSlots = __ldg(reinterpret_cast<const uint32_t *>(job.recs + mb) + 4);
            const int mby = mbRowOf(mb, g);
            mX = mb - (uint32_t)mby * g.widthMbs;
            mRowY = (uint32_t)((mby * 16 + kPadY) * g.pitchY + kPadY);
            mRowC = (uint32_t)((mby * 8 + kPadC) * g.pitchC + kPadC);
            mDelta = ((long long)(refSlots & 0xFF) - (long long)job.curSlot) * (long long)g.frameStride;
            mPieces = (mLen + kBulkPieceMbs - 1) / kBulkPieceMbs;
        }
        uint32_t mEnd = mPieces;   // inclusive prefix sum of the piece counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, mEnd, d);
            if (lane >= d) mEnd += v;
        }
        const uint32_t nPieces = __shfl_sync(0xffffffffu, mEnd, 31);
        // piece j of the task, as seen by this lane
        auto pieceAt = [&](uint32_t j) {
            const int r = __popc(__ballot_sync(0xffffffffu, mEnd <= j));              // the run piece j belongs to
            const uint32_t first = j - (__shfl_sync(0xffffffffu, mEnd, r) - __shfl_sync(0xffffffffu, mPieces, r));
            const uint32_t len = __shfl_sync(0xffffffffu, mLen, r) - first * kBulkPieceMbs;
            const uint32_t plen = min(len, (uint32_t)kBulkPieceMbs);
            const uint32_t x0 = __shfl_sync(0xffffffffu, mX, r) + first * kBulkPieceMbs, x1 = x0 + plen;
            const uint32_t rowY = __shfl_sync(0xffffffffu, mRowY, r), rowC = __shfl_sync(0xffffffffu, mRowC, r);
            const uint32_t xs = x0 + (x0 & 1u);                                        // chroma: first even column
            const uint32_t lenE = (x1 & ~1u) > xs ? (x1 & ~1u) - xs : 0u;              // macroblocks of the aligned part
            const bool lead = (x0 & 1u) != 0, trail = (x1 & 1u) != 0 && x1 - 1 >= xs;
            BulkPiece pc;
            pc.delta = __shfl_sync(0xffffffffu, mDelta, r);
            pc.total = 256u * plen + 128u * lenE;
            const int cpl = (lane >> 3) & 1, crow = lane & 7;
            uint8_t *cbase = cur + (cpl ? g.offCr : g.offCb) + rowC + (size_t)crow * g.pitchC;
            if (lane < 16) {
                pc.dst = cur + rowY + (size_t)lane * g.pitchY + x0 * 16u;
                pc.bytes = 16u * plen;
                pc.smemOff = (uint32_t)lane * pc.bytes;
                pc.edge = lead ? cbase + x0 * 8u : nullptr;
            } else {
                pc.dst = cbase + xs * 8u;
                pc.bytes = 8u * lenE;
                pc.smemOff = 256u * plen + (uint32_t)(lane - 16) * pc.bytes;
                pc.edge = trail ? cbase + (x1 - 1) * 8u : nullptr;
            }
            return pc;
        };
        auto issue = [&](uint32_t j) {
            // buffer q % kBulkBufs was last read by the stores of piece q - kBulkBufs; this lane has committed one bulk group per
            // stored piece (w of them), so at most w - (q - kBulkBufs) - 1 = kBulkBufs - 1 - (q - w) of its groups may still be
            // reading.  q - w <= kBulkAhead here, so waiting for "at most kBulkBufs - 1 - kBulkAhead pending" always suffices
            bulkWaitRead<kBulkBufs - 1 - kBulkAhead>();
            __syncwarp();                                  // every lane's stores have let go of the buffer
            const BulkPiece pc = pieceAt(j);
            const uint32_t b = q % kBulkBufs;
            if (lane == 0) mbarExpectTx(&sm.bar[b], pc.total);
            __syncwarp();
            if (pc.bytes) bulkLoad(sm.buf[b] + pc.smemOff, pc.dst + pc.delta, pc.bytes, &sm.bar[b]);
            q++;
        };
        uint32_t issued = 0;
        for (; issued < nPieces && issued < (uint32_t)kBulkAhead; issued++) issue(issued);
#pragma unroll 1
        for (uint32_t j = 0; j < nPieces; j++) {
            if (issued < nPieces) issue(issued++);
            const BulkPiece pc = pieceAt(j);
            uint2 e = make_uint2(0u, 0u);
            if (pc.edge) e = __ldg(reinterpret_cast<const uint2 *>(pc.edge + pc.delta));
            const uint32_t b = w % kBulkBufs, parity = (w / kBulkBufs) & 1u;
            // the piece has landed -- or, should the protocol be wrong on hardware, the watchdog ends this warp's work: the error
            // is reported (h264bsdB200BatchWatchdog) instead of a hung GPU
            bool landed = false;
            unsigned spins = 0;
            unsigned long long t0 = 0;
            while (!(landed = mbarTryWait(&sm.bar[b], parity))) {
                if ((++spins & 255u) == 0) {
                    const unsigned long long now = globalTimerNs();
                    if (!t0) t0 = now;
                    else if (now - t0 > kWatchdogNs) break;
                }
            }
            if (__ballot_sync(0xffffffffu, landed) == 0u) {
                if (lane == 0) atomicAdd(&gWatchdog[1], 1u);
                dead = true;
                break;
            }
            if (pc.bytes) bulkStore(pc.dst, sm.buf[b] + pc.smemOff, pc.bytes);
            bulkCommit();
            if (pc.edge) *reinterpret_cast<uint2 *>(pc.edge) = e;
            w++;
        }
        if (dead) break;
    }
    bulkWaitAll();   // the stores have reached memory before the warp leaves
}

}  // namespace b200
