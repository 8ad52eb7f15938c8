// conceal_emu.cpp -- TEST INFRASTRUCTURE.  The shipped sources of concealKernel and reconCopyKernel
// (h264bsd_b200/csrc/engine/conceal_kernel.cuh, copy_kernel.cuh) compiled for the host with warp_emu.hpp and exposed to the tests:
//   emu_geom()     the pool geometry the engine would use (makePoolGeom)
//   emu_conceal()  one launch of concealKernel over a pool of nStreams streams in host memory
//   emu_copy()     one launch of reconCopyKernel, the grid the engine would use cut down to `blocks`
#include "warp_emu.hpp"
#include "copy_kernel.cuh"
#include "conceal_kernel.cuh"

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
    for (uint32_t b = 0; b < (nStreams + kConcealWarps - 1) / kConcealWarps; b++)
        warp_emu::runBlock(b, kConcealWarps, [&]() { concealKernel(p); });
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
    gridDim.x = blocks;                                      // a persistent grid: tasks are strided over it
    for (uint32_t b = 0; b < blocks; b++)
        warp_emu::runBlock(b, kCopyWarps, [&]() { reconCopyKernel(p); });
}
