// kernels_emu.cpp -- TEST INFRASTRUCTURE.  The shipped sources of concealKernel, reconCopyKernel, strengthKernel and deblockKernel
// (h264bsd_b200/csrc/engine/conceal_kernel.cuh, copy_kernel.cuh, deblock_kernel.cuh) compiled for the host with warp_emu.hpp and exposed to the tests:
//   emu_geom()     the pool geometry the engine would use (makePoolGeom)
//   emu_conceal()  one launch of concealKernel over a pool of nStreams streams in host memory
//   emu_copy()     one launch of reconCopyKernel, the grid the engine would use cut down to `blocks`
//   emu_deblock()  strengthKernel + deblockKernel over a pool of nStreams streams (Batch::launchPicture's deblock half)
#include "warp_emu.hpp"
#include "copy_kernel.cuh"
#include "conceal_kernel.cuh"
#include "deblock_kernel.cuh"

using namespace b200;

extern "C" void emu_geom(uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint64_t out[8]) {
    const PoolGeom g = makePoolGeom(widthMbs, heightMbs, numSlots, 1);
    out[0] = (uint64_t)g.pitchY; out[1] = (uint64_t)g.pitchC; out[2] = (uint64_t)g.rowsY; out[3] = (uint64_t)g.rowsC;
    out[4] = g.offCb; out[5] = g.offCr; out[6] = g.frameStride; out[7] = (uint64_t)kPadY | ((uint64_t)kPadC << 32);
}

extern "C" void emu_conceal(uint8_t *pool, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t curSlot,
                            const b200_mb_rec *recs, const uint16_t *order, uint32_t nR, uint32_t nC, uint32_t nA, uint32_t nB, uint32_t nE,
                            uint32_t nStreams) {
    // every stream of the batch gets the same job (its own frames: stream s owns slots [s * numSlots, (s + 1) * numSlots))
    StreamJob job;
    std::memset(&job, 0, sizeof job);
    job.recs = recs; job.order = order; job.curSlot = (uint16_t)curSlot;
    job.nR = (uint16_t)nR; job.nC = (uint16_t)nC; job.nA = (uint16_t)nA; job.nB = (uint16_t)nB; job.nE = (uint16_t)nE;
    std::vector<StreamJob> jobs(nStreams, job);
    ReconParams p;
    std::memset(&p, 0, sizeof p);
    p.pool = pool;
    p.g = makePoolGeom(widthMbs, heightMbs, numSlots, nStreams);
    p.jobs = jobs.data();
    // the engine's launch: ceil(nStreams / kConcealWarps) blocks of kConcealWarps warps
    warp_emu::runGrid((nStreams + kConcealWarps - 1) / kConcealWarps, kConcealWarps * 32, [&]() { concealKernel(p); });
}

extern "C" void emu_copy(uint8_t *pool, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t curSlot,
                         const b200_mb_rec *recs, const uint16_t *order, uint32_t nR, uint32_t nC, uint32_t nStreams, uint32_t copyRuns,
                         uint32_t blocks) {
    StreamJob job;
    std::memset(&job, 0, sizeof job);
    job.recs = recs; job.order = order; job.curSlot = (uint16_t)curSlot;
    job.nR = (uint16_t)nR; job.nC = (uint16_t)nC;
    std::vector<StreamJob> jobs(nStreams, job);
    ReconParams p;
    std::memset(&p, 0, sizeof p);
    p.pool = pool;
    p.g = makePoolGeom(widthMbs, heightMbs, numSlots, nStreams);
    p.jobs = jobs.data();
    p.copyRuns = copyRuns;                                   // as Batch::launchPicture sets them
    p.chunksC = (nC + 31) / 32;
    p.chunksQ = (nR + copyRuns - 1) / copyRuns;
    warp_emu::runGrid(blocks, kCopyWarps * 32, [&]() { reconCopyKernel(p); });   // a persistent grid: tasks are strided over it
}

// The in-loop filter of one picture: boundary strengths, then the ticketed wavefront filter -- parameters as Batch::create /
// Batch::launchPicture set them up.  Returns the watchdog count of the flag waits (0 = nobody waited in vain).
extern "C" uint32_t emu_deblock(uint8_t *pool, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t curSlot,
                                const b200_mb_rec *recs, uint32_t nStreams, uint32_t filterChunk, uint32_t blocks) {
    StreamJob job;
    std::memset(&job, 0, sizeof job);
    job.recs = recs; job.curSlot = (uint16_t)curSlot;
    std::vector<StreamJob> jobs(nStreams, job);
    DeblockParams dp;
    std::memset(&dp, 0, sizeof dp);
    dp.g = makePoolGeom(widthMbs, heightMbs, numSlots, nStreams);
    const uint32_t nMbs = (uint32_t)dp.g.nMbs, total = nStreams * nMbs;
    // wavefront order: x + 2y ascending (Batch::create)
    std::vector<std::pair<uint32_t, uint32_t>> keyed(nMbs);
    for (uint32_t mb = 0; mb < nMbs; mb++) keyed[mb] = {mb % widthMbs + 2 * (mb / widthMbs), mb};
    std::sort(keyed.begin(), keyed.end());
    std::vector<uint16_t> order(nMbs);
    for (uint32_t i = 0; i < nMbs; i++) order[i] = (uint16_t)keyed[i].second;
    std::vector<uint32_t> done(total, 0), bsWords((size_t)total * 4, 0), counters(8, 0);
    std::vector<uint8_t> work(total, 0);
    dp.pool = pool; dp.jobs = jobs.data(); dp.order = order.data(); dp.done = done.data();
    dp.ticket = counters.data() + 1; dp.serial = 1; dp.totalTickets = total;
    dp.bsWords = bsWords.data(); dp.work = work.data();
    dp.workCount = reinterpret_cast<unsigned long long *>(counters.data() + 4);
    dp.filterChunk = filterChunk;
    gWatchdog[0] = 0;
    const uint32_t chunks = (nMbs + kDeblockWarps * 32 - 1) / (kDeblockWarps * 32) * nStreams;
    warp_emu::runGrid(std::min(chunks, blocks), kDeblockWarps * 32, [&]() { strengthKernel(dp); });
    const uint32_t ctas = (total + kDeblockWarps * filterChunk - 1) / (kDeblockWarps * filterChunk);
    warp_emu::runGrid(std::min(ctas, blocks), kDeblockWarps * 32, [&]() { deblockKernel(dp); });
    return gWatchdog[0];
}
