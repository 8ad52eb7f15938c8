// kernels_emu.cpp -- TEST INFRASTRUCTURE.  The shipped sources of concealKernel, reconCopyKernel, strengthKernel and deblockKernel
// (h264bsd_b200/csrc/engine/conceal_kernel.cuh, copy_kernel.cuh, deblock_kernel.cuh) compiled for the host with warp_emu.hpp and exposed to the tests:
//   emu_geom()     the pool geometry the engine would use (makePoolGeom)
//   emu_conceal()  one launch of concealKernel over a pool of nStreams streams in host memory
//   emu_copy()     one launch of reconCopyKernel, the grid the engine would use cut down to `blocks`; copyRuns with bit 31 set:
//                  the B200_COPY_BULK=1 sequence instead (reconCopyBulkKernel for the runs, reconCopyKernel for the single copies);
//                  bit 30: reconCopyKernelDeep (B200_COPY_VARIANT=2)
//   emu_deblock()  strengthKernel + deblockKernel over a pool of nStreams streams (Batch::launchPicture's deblock half)
//   emu_engine_*() the whole per-picture launch sequence of Batch::launchPicture over a persistent pool
#include "warp_emu.hpp"
#include "recon_kernel_emu.cuh"   // = recon_kernel.cuh with the dynamic shared array declared plain extern (made by the test)
#include "copy_kernel.cuh"
#include "copy_bulk_kernel.cuh"
#include "conceal_kernel.cuh"
#include "deblock_kernel.cuh"

namespace b200 {
alignas(128) uint8_t interSmemRaw[sizeof(InterWarpSmem) * kReconWarps];   // the dynamic shared memory of reconInterKernel
alignas(128) uint8_t bulkSmemRaw[sizeof(BulkWarpSmem) * kBulkWarps];      // ... of reconCopyBulkKernel
}

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

// the copy pass with B200_COPY_BULK=1 (Batch::launchPicture): runs by the bulk kernel, single copies by a launch without run tasks
static void launchCopyBulk(const ReconParams &rp, uint32_t nR, uint32_t nC, uint32_t blocks) {
    ReconParams rq = rp, rs = rp;
    rq.copyRuns = std::min<uint32_t>(rp.copyRuns, kBulkRunsPerTask);
    rq.chunksQ = (nR + rq.copyRuns - 1) / rq.copyRuns;
    rs.chunksQ = 0;
    if (nR) {
        const uint32_t ctas = (rq.chunksQ * (uint32_t)rp.g.nStreams + kBulkWarps - 1) / kBulkWarps;
        warp_emu::runGrid(std::min(ctas, blocks), kBulkWarps * 32, [&]() { reconCopyBulkKernel(rq); });
    }
    if (nC) {
        const uint32_t ctas = (rs.chunksC * (uint32_t)rp.g.nStreams + kCopyWarps - 1) / kCopyWarps;
        warp_emu::runGrid(std::min(ctas, blocks), kCopyWarps * 32, [&]() { reconCopyKernel(rs); });
    }
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
    const bool bulk = (copyRuns & 0x80000000u) != 0, deep = (copyRuns & 0x40000000u) != 0;   // bit 30: B200_COPY_VARIANT=2
    copyRuns &= 0x3FFFFFFFu;
    p.copyRuns = copyRuns;                                   // as Batch::launchPicture sets them
    p.chunksC = (nC + 31) / 32;
    p.chunksQ = (nR + copyRuns - 1) / copyRuns;
    if (bulk) launchCopyBulk(p, nR, nC, blocks);
    else if (deep) warp_emu::runGrid(blocks, kCopyWarps * 32, [&]() { reconCopyKernelDeep(p); });
    else warp_emu::runGrid(blocks, kCopyWarps * 32, [&]() { reconCopyKernel(p); });   // a persistent grid: tasks are strided over it
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

// ---- the whole engine: pool, flags and counters live across pictures as in Batch; one call = Batch::launchPicture ----
struct EmuEngine {
    PoolGeom g;
    std::vector<uint8_t> poolRaw;
    uint8_t *pool = nullptr;
    std::vector<uint32_t> doneRecon, doneDeblock, bsWords, counters;
    std::vector<uint8_t> work;
    std::vector<uint16_t> order;
    CUtensorMap lumaMap, chromaMap;
    uint32_t serial = 0;
};

extern "C" EmuEngine *emu_engine_create(uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots, uint32_t nStreams) {
    EmuEngine *e = new EmuEngine();
    e->g = makePoolGeom(widthMbs, heightMbs, numSlots, nStreams);
    const PoolGeom &g = e->g;
    const unsigned long long nFrames = (unsigned long long)nStreams * numSlots;
    e->poolRaw.assign(nFrames * g.frameStride + 256, 128);                     // Batch::create: cudaMemset(pool, 128)
    e->pool = e->poolRaw.data() + ((256 - (reinterpret_cast<uintptr_t>(e->poolRaw.data()) & 255)) & 255);
    const size_t total = (size_t)nStreams * g.nMbs;
    e->doneRecon.assign(total, 0); e->doneDeblock.assign(total, 0); e->bsWords.assign(total * 4, 0); e->work.assign(total, 0);
    e->counters.assign(8, 0);
    std::vector<std::pair<uint32_t, uint32_t>> keyed(g.nMbs);
    for (int mb = 0; mb < g.nMbs; mb++) keyed[mb] = {(uint32_t)(mb % g.widthMbs + 2 * (mb / g.widthMbs)), (uint32_t)mb};
    std::sort(keyed.begin(), keyed.end());
    e->order.resize(g.nMbs);
    for (int i = 0; i < g.nMbs; i++) e->order[i] = (uint16_t)keyed[i].second;
    // the tiled views of Batch::create: luma {x, y, frame} box 48 x 21, chroma {x, y, plane, frame} box 32 x 9 x 2
    CUtensorMap &l = e->lumaMap, &c = e->chromaMap;
    std::memset(&l, 0, sizeof l); std::memset(&c, 0, sizeof c);
    l.base = e->pool; l.rank = 3;
    l.dims[0] = (uint64_t)g.pitchY; l.dims[1] = (uint64_t)g.rowsY; l.dims[2] = nFrames;
    l.strides[0] = (uint64_t)g.pitchY; l.strides[1] = g.frameStride;
    l.box[0] = kLumaBoxW; l.box[1] = kLumaBoxH; l.box[2] = 1;
    c.base = e->pool + g.offCb; c.rank = 4;
    c.dims[0] = (uint64_t)g.pitchC; c.dims[1] = (uint64_t)g.rowsC; c.dims[2] = 2; c.dims[3] = nFrames;
    c.strides[0] = (uint64_t)g.pitchC; c.strides[1] = (uint64_t)g.pitchC * g.rowsC; c.strides[2] = g.frameStride;
    c.box[0] = kChromaBoxW; c.box[1] = kChromaBoxH; c.box[2] = 2; c.box[3] = 1;
    return e;
}
extern "C" void emu_engine_destroy(EmuEngine *e) { delete e; }
extern "C" uint8_t *emu_engine_pool(EmuEngine *e) { return e->pool; }

// One picture of every stream (all streams replay the same work-list).  chunkA / chunkB / copyRuns / filterChunk: the engine's
// tuning knobs; blocks: cap on the grid of the persistent kernels.  Returns watchdog[0] | watchdog[1] << 8 | IDCT errors << 16.
extern "C" uint32_t emu_engine_picture(EmuEngine *e, const b200_mb_rec *recs, const b200_mb_rec *filterRecs, const int16_t *coefs, const uint16_t *order, uint32_t curSlot,
                                       uint32_t nR, uint32_t nC, uint32_t nA, uint32_t nB, uint32_t nE, int recon, int deblock,
                                       uint32_t chunkA, uint32_t chunkB, uint32_t copyRuns, uint32_t filterChunk, uint32_t blocks) {
    const PoolGeom &g = e->g;
    StreamJob job;
    std::memset(&job, 0, sizeof job);
    job.recs = recs; job.coefs = coefs; job.order = order; job.curSlot = (uint16_t)curSlot;
    job.nR = (uint16_t)nR; job.nC = (uint16_t)nC; job.nA = (uint16_t)nA; job.nB = (uint16_t)nB; job.nE = (uint16_t)nE;
    std::vector<StreamJob> jobs(g.nStreams, job);
    job.recs = filterRecs ? filterRecs : recs;               // what the filter kernels get (Batch::buildJobs)
    std::vector<StreamJob> jobsFilter(g.nStreams, job);
    const uint32_t total = (uint32_t)g.nStreams * (uint32_t)g.nMbs;
    e->serial++;
    gWatchdog[0] = gWatchdog[1] = 0;
    if (recon) {
        ReconParams rp;
        std::memset(&rp, 0, sizeof rp);
        rp.pool = e->pool; rp.g = g; rp.jobs = jobs.data(); rp.done = e->doneRecon.data();
        rp.ticket = e->counters.data() + 0; rp.errors = e->counters.data() + 2; rp.serial = e->serial;
        rp.chunkB = chunkB; rp.chunksB = (nB + chunkB - 1) / chunkB;
        const bool bulk = (copyRuns & 0x80000000u) != 0;   // the B200_COPY_BULK=1 sequence
        copyRuns &= 0x7FFFFFFFu;
        rp.chunkA = chunkA; rp.copyRuns = copyRuns;
        rp.chunksA = (nA + kReconWarps * chunkA - 1) / (kReconWarps * chunkA);
        rp.virtualCtasA = rp.chunksA * (uint32_t)g.nStreams;
        rp.chunksC = (nC + 31) / 32;
        rp.chunksQ = (nR + copyRuns - 1) / copyRuns;
        if ((nC || nR) && bulk) launchCopyBulk(rp, nR, nC, blocks);
        else if (nC || nR) {
            const uint32_t ctas = ((rp.chunksC + rp.chunksQ) * (uint32_t)g.nStreams + kCopyWarps - 1) / kCopyWarps;
            warp_emu::runGrid(std::min(ctas, blocks), kCopyWarps * 32, [&]() { reconCopyKernel(rp); });
        }
        if (nA) warp_emu::runGrid(std::min(rp.virtualCtasA, blocks), kReconWarps * 32, [&]() { reconInterKernel(rp, e->lumaMap, e->chromaMap); });
        if (nB) warp_emu::runGrid((rp.chunksB * (uint32_t)g.nStreams + kReconWarps - 1) / kReconWarps, kReconWarps * 32, [&]() { reconIntraKernel(rp); });
        if (nE) warp_emu::runGrid(((uint32_t)g.nStreams + kConcealWarps - 1) / kConcealWarps, kConcealWarps * 32, [&]() { concealKernel(rp); });
    }
    if (deblock) {
        DeblockParams dp;
        std::memset(&dp, 0, sizeof dp);
        dp.pool = e->pool; dp.g = g; dp.jobs = jobsFilter.data(); dp.order = e->order.data(); dp.done = e->doneDeblock.data();
        dp.ticket = e->counters.data() + 1; dp.serial = e->serial; dp.totalTickets = total;
        dp.bsWords = e->bsWords.data(); dp.work = e->work.data();
        dp.workCount = reinterpret_cast<unsigned long long *>(e->counters.data() + 4);
        dp.filterChunk = filterChunk;
        const uint32_t chunks = ((uint32_t)g.nMbs + kDeblockWarps * 32 - 1) / (kDeblockWarps * 32) * (uint32_t)g.nStreams;
        warp_emu::runGrid(std::min(chunks, blocks), kDeblockWarps * 32, [&]() { strengthKernel(dp); });
        const uint32_t ctas = (total + kDeblockWarps * filterChunk - 1) / (kDeblockWarps * filterChunk);
        warp_emu::runGrid(std::min(ctas, blocks), kDeblockWarps * 32, [&]() { deblockKernel(dp); });
    }
    {
        BorderParams bp;
        bp.pool = e->pool; bp.g = g; bp.jobs = jobs.data();
        const long long borderTasks = (g.H + 31) / 32 + 2 * ((g.H / 2 + 31) / 32) + 2 * ((g.pitchY + 127) / 128) + 4 * ((g.pitchC + 127) / 128);
        const long long tasks = borderTasks * g.nStreams;
        warp_emu::runGrid((unsigned)((tasks + 7) / 8), 256, [&]() { borderKernel(bp); });
    }
    e->counters[0] = e->counters[1] = 0;       // Batch::launchPicture: the ticket counters start from zero again
    return gWatchdog[0] | (gWatchdog[1] << 8) | (e->counters[2] << 16);
}
