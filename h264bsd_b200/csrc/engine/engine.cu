// engine.cu -- host side of the B200 reconstruction engine: HBM frame pool, tape upload, per-picture
// launch sequence (reconstruct -> in-loop filter -> border replication), read-back.
//
// One Batch = N independent streams of identical geometry resident on one GPU.  Pictures with the same
// decode index are processed together: one launch per stage covers all streams.  There is no CPU
// fallback anywhere in this file: without a usable CUDA device every entry point fails loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include "recon_kernel.cuh"
#include "conceal_kernel.cuh"
#include "deblock_kernel.cuh"
#include "engine.hpp"

namespace b200 {

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            std::fprintf(stderr, "h264bsd_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
                         cudaGetErrorString(e_));                                                    \
            return false;                                                                            \
        }                                                                                            \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn getEncodeTiled() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return nullptr;
    return (EncodeTiledFn)fn;
}

int deviceCount() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

Batch::~Batch() { destroy(); }

void Batch::destroy() {
    if (!created_) return;
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (copyStream_) cudaStreamSynchronize(copyStream_);
    if (uploadStream_) cudaStreamSynchronize(uploadStream_);
    if (auxStream_) cudaStreamSynchronize(auxStream_);
    for (auto &t : tapes_) {
        if (t.owned) { cudaFree(t.recs); cudaFree(t.coefs); cudaFree(t.order); }
    }
    cudaFree(dBsWords_); cudaFree(dWork_); cudaFree(dPack_[0]); cudaFree(dPack_[1]);
    cudaFree(dConvertAll_); cudaFree(dFrameStage_); cudaFree(dMirror_);
    for (cudaEvent_t e : mirrorEv_) if (e) cudaEventDestroy(e);
    cudaFree(pool_); cudaFree(dDoneRecon_); cudaFree(dDoneDeblock_); cudaFree(dCounters_); cudaFree(dMultiList_);
    cudaFree(dJobs_); cudaFree(dStage_[0]); cudaFree(dStage_[1]); cudaFree(dConvert_); cudaFree(dSlots_);
    if (hStage_[0]) cudaFreeHost(hStage_[0]);
    if (hStage_[1]) cudaFreeHost(hStage_[1]);
    for (cudaEvent_t e : evPool_) cudaEventDestroy(e);
    for (auto &f : fences_) cudaEventDestroy(f.second);
    for (cudaEvent_t e : fenceFree_) cudaEventDestroy(e);
    for (cudaEvent_t e : {syncEv_, forkEv_, joinEv_, evA_, evB_, stageEv_[0], stageEv_[1], packEv_[0], packEv_[1], packedEv_[0], packedEv_[1]})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {copyStream_, uploadStream_, auxStream_, stream_})
        if (st) cudaStreamDestroy(st);
    // every pointer, capacity, event and stream back to its default: a second create() on this object (a stream that activates
    // another sequence parameter set, api.cpp LegacyDecoder::configure) must not see anything of the first
    created_ = false;
    resetState();
}

void Batch::resetState() {
    created_ = false; device_ = 0; numSms_ = 0; stream_ = nullptr; evA_ = evB_ = nullptr; g_ = PoolGeom{}; pool_ = nullptr;
    dDoneRecon_ = dDoneDeblock_ = dCounters_ = dSlots_ = dBsWords_ = dMultiList_ = nullptr; dWork_ = nullptr;
    strengthBlocks_ = 0; serial_ = 0; passABlocks_ = deblockBlocks_ = intraBlocks_ = 0; chunkRows_ = 32; chunksPerCol_ = 1;
    syncEv_ = forkEv_ = joinEv_ = nullptr; jobsCap_ = 0; dConvertAll_ = nullptr; uploadStream_ = nullptr;
    fences_.clear(); fenceFree_.clear(); auxStream_ = nullptr; tapes_.clear(); dJobs_ = nullptr; jobsFilterAt_ = 0; numPics_ = 0;
    jobsDirty_ = true; hStage_[0] = hStage_[1] = nullptr; dStage_[0] = dStage_[1] = nullptr; stageCap_[0] = stageCap_[1] = 0;
    stageEv_[0] = stageEv_[1] = nullptr; stageIdx_ = 0; dConvert_ = nullptr; convertCap_ = 0; dFrameStage_ = nullptr; frameStageCap_ = 0;
    dMirror_ = nullptr; mirrorEv_.clear(); mirrorBusy_.clear();
    launches_ = 0; d2hBytes_ = 0; h2dBytes_ = 0; timing_ = false; evPool_.clear(); evUsed_ = 0; evStage_.clear(); picMaxA_.clear();
    dPack_[0] = dPack_[1] = nullptr; packEv_[0] = packEv_[1] = nullptr; packedEv_[0] = packedEv_[1] = nullptr; copyStream_ = nullptr;
    packUsed_[0] = packUsed_[1] = false; packIdx_ = 0; picMaxB_.clear(); picMaxE_.clear();
}

static bool encodeStripMap(EncodeTiledFn enc, CUtensorMap *m, uint8_t *base, const PoolGeom &g, int rows, unsigned long long nFrames, int nx, int nr) {
    // dims (x in strip, strip, row, frame); the strip stride is larger than the row stride on purpose (pool_geom.hpp)
    cuuint64_t dims[4] = {16, (cuuint64_t)g.strips, (cuuint64_t)rows, nFrames};
    cuuint64_t strides[3] = {(cuuint64_t)rows * 16, 16, g.frameStride};
    cuuint32_t box[4] = {16, (cuuint32_t)nx, (cuuint32_t)nr, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { std::fprintf(stderr, "h264bsd_b200: tensor map (%d strips x %d rows) failed (%d)\n", nx, nr, (int)r); return false; }
    return true;
}

bool Batch::create(int device, uint32_t nStreams, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots) {
    int n = deviceCount();
    if (n <= 0) {
        std::fprintf(stderr, "h264bsd_b200: no CUDA device visible -- this engine has no CPU fallback\n");
        return false;
    }
    if (device < 0 || device >= n || !nStreams || !widthMbs || !heightMbs || !numSlots || numSlots > 32) return false;
    // macroblock addresses are 16 bits on the device (order lists); checked before anything is allocated
    // (and pass A packs a window's row into 16 bits: rows incl. border < 65536)
    if ((unsigned long long)widthMbs * heightMbs > 65535ull || widthMbs > 4000 || heightMbs > 4000) return false;
    if (created_) destroy();
    device_ = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        std::fprintf(stderr, "h264bsd_b200: device sm_%d%d lacks TMA; this engine targets sm_100a\n", prop.major, prop.minor);
        return false;
    }
    numSms_ = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CK(cudaEventCreate(&evA_));
    CK(cudaEventCreate(&evB_));
    created_ = true;

    g_ = makePoolGeom(widthMbs, heightMbs, numSlots, nStreams);
    PoolGeom &g = g_;
    const unsigned long long nFrames = (unsigned long long)nStreams * numSlots;
    CK(cudaMalloc(&pool_, nFrames * g.frameStride));
    CK(cudaMemsetAsync(pool_, 128, nFrames * g.frameStride, stream_));

    // row-progress words of the intra pass and of the filter (serial << 16 | macroblocks finished), boundary strengths
    const size_t rowBytes = sizeof(uint32_t) * (size_t)nStreams * g.heightMbs;
    CK(cudaMalloc(&dDoneRecon_, rowBytes));
    CK(cudaMalloc(&dDoneDeblock_, rowBytes));
    CK(cudaMemsetAsync(dDoneRecon_, 0, rowBytes, stream_));
    CK(cudaMemsetAsync(dDoneDeblock_, 0, rowBytes, stream_));
    CK(cudaMalloc(&dBsWords_, sizeof(uint32_t) * 4 * (size_t)nStreams * g.nMbs));
    CK(cudaMalloc(&dWork_, (size_t)nStreams * g.nMbs));
    // counters: [0] pass-B tickets, [1] filter tickets, [2] [3] tickets of the two pass-A instances, [4] entries in the list of
    // macroblocks with several partitions (zeroed per picture); [8] IDCT range errors, [10..11] macroblocks with filter work
    // (running totals)
    CK(cudaMalloc(&dCounters_, sizeof(uint32_t) * 16));
    CK(cudaMemsetAsync(dCounters_, 0, sizeof(uint32_t) * 16, stream_));
    if (nStreams > 65535u) return false;   // (list entries are stream << 16 | address)
    CK(cudaMalloc(&dMultiList_, sizeof(uint32_t) * (size_t)nStreams * g.nMbs));
    CK(cudaMalloc(&dSlots_, sizeof(uint32_t) * nStreams));
    serial_ = 0;

    // TMA descriptors over the whole pool, one per box shape of pass A
    EncodeTiledFn enc = getEncodeTiled();
    if (!enc) {
        std::fprintf(stderr, "h264bsd_b200: cuTensorMapEncodeTiled unavailable\n");
        return false;
    }
    static_assert(sizeof(PassAMaps) == sizeof(CUtensorMap) * 10, "Batch::maps_ holds a PassAMaps");
    {
        PassAMaps *m = reinterpret_cast<PassAMaps *>(maps_);
        for (int nx = 1; nx <= 3; nx++)
            for (int v = 0; v < 2; v++)
                if (!encodeStripMap(enc, &m->luma[nx - 1][v], pool_, g, g.rowsY, nFrames, nx, v ? 21 : 16)) return false;
        for (int nx = 1; nx <= 2; nx++)
            for (int v = 0; v < 2; v++)
                if (!encodeStripMap(enc, &m->chroma[nx - 1][v], pool_ + g.offC, g, g.rowsC, nFrames, nx, v ? 9 : 8)) return false;
    }
    int occA = 1, occD = 0, occS = 0, occB = 0;
    CK(cudaFuncSetAttribute(passAKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(PassAWarpSmem) * kPassAWarps)));
    CK(cudaFuncSetAttribute(passAMultiKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(MultiWarpSmem) * kPassAWarps)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, passAKernel, kPassAWarps * 32, sizeof(PassAWarpSmem) * kPassAWarps));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occD, deblockKernel, kDeblockWarps * 32, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, strengthKernel, kDeblockWarps * 32, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, reconIntraKernel, kReconWarps * 32, 0));
    intraBlocks_ = std::max(1, occB) * numSms_;
    strengthBlocks_ = std::max(1, occS) * numSms_;
    passABlocks_ = std::max(1, occA) * numSms_;
    deblockBlocks_ = std::max(1, occD) * numSms_;
    // pass A: a warp task is a column piece of at most kStageMbs macroblocks (what a warp's copy staging buffer holds); pieces of a
    // column are made equally long
    chunksPerCol_ = (heightMbs + kStageMbs - 1) / kStageMbs;
    chunkRows_ = (heightMbs + chunksPerCol_ - 1) / chunksPerCol_;
    tapes_.assign(nStreams, DevTape());
    CK(cudaStreamCreateWithFlags(&auxStream_, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&joinEv_, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&forkEv_, cudaEventDisableTiming));
    CK(cudaStreamSynchronize(stream_));
    return true;
}

bool Batch::uploadTape(uint32_t stream, const b200_tape *t) {
    if (!created_) return false;
    CK(cudaSetDevice(device_));
    if (!uploadTapeOn(stream, t, stream_)) return false;
    jobsDirty_ = true;
    return true;
}

bool Batch::uploadTapeOn(uint32_t stream, const b200_tape *t, cudaStream_t st) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || !t) return false;
    if (t->widthMbs != (uint32_t)g_.widthMbs || t->heightMbs != (uint32_t)g_.heightMbs || t->numSlots > (uint32_t)g_.numSlots) return false;
    CK(cudaSetDevice(device_));
    DevTape &d = tapes_[stream];
    const size_t orderBytes = (size_t)t->numOrder * sizeof(uint16_t);
    // re-use the device arrays of a previous upload when they are large enough (no cudaMalloc in steady state)
    if (!d.owned || d.capRecs < t->mbRecBytes || d.capCoefs < t->coefBytes || d.capOrder < orderBytes) {
        CK(cudaStreamSynchronize(stream_));
        if (d.owned) { cudaFree(d.recs); cudaFree(d.coefs); cudaFree(d.order); }
        d = DevTape();
        d.capRecs = t->mbRecBytes + t->mbRecBytes / 8 + 256;
        d.capCoefs = t->coefBytes + t->coefBytes / 4 + 256;
        d.capOrder = orderBytes + 256;
        CK(cudaMalloc(&d.recs, d.capRecs));
        CK(cudaMalloc(&d.coefs, d.capCoefs));
        CK(cudaMalloc(&d.order, d.capOrder));
        d.owned = true;
    }
    d.recBytes = t->mbRecBytes; d.coefBytes = t->coefBytes; d.orderBytes = orderBytes;
    CK(cudaMemcpyAsync(d.recs, t->mbRecs, t->mbRecBytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d.coefs, t->coefs, t->coefBytes, cudaMemcpyHostToDevice, st));
    if (orderBytes) CK(cudaMemcpyAsync(d.order, t->mbOrder, orderBytes, cudaMemcpyHostToDevice, st));
    d.pics.assign(t->pics, t->pics + t->numPics);
    h2dBytes_ += t->mbRecBytes + t->coefBytes + orderBytes;
    return true;
}

// Streaming form of uploadTape: the work-list of pictures [firstPic, firstPic + numPics) of one stream, on the upload
// stream, so that the H2D of later pictures overlaps the decode (and the D2H of the output) of earlier ones.  The call
// with firstPic == 0 (re)sizes the device arrays and takes the picture headers; uploadFence(p) then makes every picture
// below p wait for what has been queued so far.
bool Batch::uploadTapeRange(uint32_t stream, const b200_tape *t, uint32_t firstPic, uint32_t numPics) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || !t || firstPic + numPics > t->numPics || !numPics) return false;
    if (t->widthMbs != (uint32_t)g_.widthMbs || t->heightMbs != (uint32_t)g_.heightMbs || t->numSlots > (uint32_t)g_.numSlots) return false;
    CK(cudaSetDevice(device_));
    if (!uploadStream_) CK(cudaStreamCreateWithFlags(&uploadStream_, cudaStreamNonBlocking));
    DevTape &d = tapes_[stream];
    const size_t orderBytes = (size_t)t->numOrder * sizeof(uint16_t);
    if (firstPic == 0) {
        if (!d.owned || d.capRecs < t->mbRecBytes || d.capCoefs < t->coefBytes || d.capOrder < orderBytes) {
            CK(cudaStreamSynchronize(stream_));
            CK(cudaStreamSynchronize(uploadStream_));
            if (d.owned) { cudaFree(d.recs); cudaFree(d.coefs); cudaFree(d.order); }
            d = DevTape();
            d.capRecs = t->mbRecBytes + t->mbRecBytes / 8 + 256;
            d.capCoefs = t->coefBytes + t->coefBytes / 4 + 256;
            d.capOrder = orderBytes + 256;
            CK(cudaMalloc(&d.recs, d.capRecs));
            CK(cudaMalloc(&d.coefs, d.capCoefs));
            CK(cudaMalloc(&d.order, d.capOrder));
            d.owned = true;
        }
        d.recBytes = t->mbRecBytes; d.coefBytes = t->coefBytes; d.orderBytes = orderBytes;
        d.pics.assign(t->pics, t->pics + t->numPics);
        jobsDirty_ = true;
    } else if (!d.owned || d.pics.size() != t->numPics) {
        return false;
    }
    const uint32_t last = firstPic + numPics;
    const uint64_t r0 = t->pics[firstPic].mbRecOffset, r1 = last < t->numPics ? t->pics[last].mbRecOffset : t->mbRecBytes;
    const uint64_t c0 = t->pics[firstPic].coefOffset, c1 = last < t->numPics ? t->pics[last].coefOffset : t->coefBytes;
    const size_t o0 = (size_t)t->pics[firstPic].orderOffset * sizeof(uint16_t), o1 = (size_t)(last < t->numPics ? t->pics[last].orderOffset : t->numOrder) * sizeof(uint16_t);
    if (r1 > r0) CK(cudaMemcpyAsync(d.recs + r0, t->mbRecs + r0, r1 - r0, cudaMemcpyHostToDevice, uploadStream_));
    if (c1 > c0) CK(cudaMemcpyAsync(d.coefs + c0, t->coefs + c0, c1 - c0, cudaMemcpyHostToDevice, uploadStream_));
    if (o1 > o0) CK(cudaMemcpyAsync(d.order + o0, reinterpret_cast<const uint8_t *>(t->mbOrder) + o0, o1 - o0, cudaMemcpyHostToDevice, uploadStream_));
    h2dBytes_ += (r1 - r0) + (c1 - c0) + (o1 - o0);
    return true;
}

bool Batch::uploadFence(uint32_t throughPic) {
    if (!created_ || !uploadStream_) return false;
    CK(cudaSetDevice(device_));
    cudaEvent_t e = nullptr;
    if (!fenceFree_.empty()) { e = fenceFree_.back(); fenceFree_.pop_back(); }
    else CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventRecord(e, uploadStream_));
    fences_.push_back({throughPic, e});
    return true;
}

// give every other stream its OWN copy in HBM of stream `src`'s tape (device-to-device)
bool Batch::replicateTape(uint32_t src) {
    if (!created_ || src >= (uint32_t)g_.nStreams || !tapes_[src].recs) return false;
    CK(cudaSetDevice(device_));
    const DevTape &s = tapes_[src];
    for (uint32_t i = 0; i < (uint32_t)g_.nStreams; i++) {
        if (i == src) continue;
        DevTape &d = tapes_[i];
        if (d.owned) { cudaFree(d.recs); cudaFree(d.coefs); cudaFree(d.order); }
        d = DevTape();
        CK(cudaMalloc(&d.recs, s.recBytes + 256));
        CK(cudaMalloc(&d.coefs, s.coefBytes + 256));
        CK(cudaMalloc(&d.order, s.orderBytes + 256));
        d.owned = true;
        d.recBytes = s.recBytes; d.coefBytes = s.coefBytes; d.orderBytes = s.orderBytes;
        CK(cudaMemcpyAsync(d.recs, s.recs, s.recBytes, cudaMemcpyDeviceToDevice, stream_));
        CK(cudaMemcpyAsync(d.coefs, s.coefs, s.coefBytes, cudaMemcpyDeviceToDevice, stream_));
        if (s.orderBytes) CK(cudaMemcpyAsync(d.order, s.order, s.orderBytes, cudaMemcpyDeviceToDevice, stream_));
        d.pics = s.pics;
    }
    jobsDirty_ = true;
    return true;
}

bool Batch::buildJobs() {
    uint32_t np = 0xFFFFFFFFu;
    for (const auto &t : tapes_) {
        if (!t.recs) return false;
        np = std::min<uint32_t>(np, (uint32_t)t.pics.size());
    }
    numPics_ = np;
    picMaxB_.assign(np, 0);
    picMaxE_.assign(np, 0);
    picMaxA_.assign(np, 0);
    jobsFilterAt_ = (size_t)np * g_.nStreams;
    std::vector<StreamJob> jobs(2 * jobsFilterAt_);
    for (uint32_t k = 0; k < np; k++)
        for (int s = 0; s < g_.nStreams; s++) {
            const DevTape &t = tapes_[s];
            StreamJob &j = jobs[(size_t)k * g_.nStreams + s];
            j.recs = reinterpret_cast<const b200_mb_rec *>(t.recs + t.pics[k].mbRecOffset);
            j.coefs = reinterpret_cast<const int16_t *>(t.coefs + t.pics[k].coefOffset);
            j.orderE = reinterpret_cast<const uint16_t *>(t.order) + t.pics[k].orderOffset;
            j.curSlot = (uint16_t)t.pics[k].curSlot;
            j.nB = (uint16_t)t.pics[k].numPassB;
            j.nE = (uint16_t)t.pics[k].numConceal;
            j.pad = 0;
            picMaxE_[k] = std::max<uint32_t>(picMaxE_[k], j.nE);
            picMaxB_[k] = std::max<uint32_t>(picMaxB_[k], j.nB);
            picMaxA_[k] = std::max<uint32_t>(picMaxA_[k], t.pics[k].numPassA);
            // what the filter kernels get: the same job, with the picture's filter records where it has any
            StreamJob &jf = jobs[jobsFilterAt_ + (size_t)k * g_.nStreams + s];
            jf = j;
            if (t.pics[k].filterRecOffset) jf.recs = reinterpret_cast<const b200_mb_rec *>(t.recs + t.pics[k].filterRecOffset);
        }
    // (cudaFree synchronises the whole device, queued uploads included: keep the table when it is large enough)
    if (jobsCap_ < jobs.size()) {
        cudaFree(dJobs_);
        dJobs_ = nullptr;
        CK(cudaMalloc(&dJobs_, sizeof(StreamJob) * jobs.size() + 64));
        jobsCap_ = jobs.size();
    }
    CK(cudaStreamSynchronize(stream_));   // nothing in flight may still read the previous table
    CK(cudaMemcpyAsync(dJobs_, jobs.data(), sizeof(StreamJob) * jobs.size(), cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
    jobsDirty_ = false;
    return true;
}

cudaEvent_t Batch::nextEvent() {
    if (evUsed_ == evPool_.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        evPool_.push_back(e);
        evStage_.push_back(-1);
    }
    return evPool_[evUsed_++];
}

void Batch::kernelTiming(bool enable) {
    timing_ = enable;
    evUsed_ = 0;
}

bool Batch::kernelTimes(float ms[6], uint32_t *launchesPerStage) {
    CK(cudaSetDevice(device_));
    CK(cudaStreamSynchronize(stream_));
    for (int i = 0; i < 6; i++) ms[i] = 0.f;
    uint32_t n[6] = {0, 0, 0, 0, 0, 0};
    for (size_t i = 1; i < evUsed_; i++) {
        const int st = evStage_[i];
        if (st < 0) continue;
        float d = 0.f;
        CK(cudaEventElapsedTime(&d, evPool_[i - 1], evPool_[i]));
        ms[st] += d;
        n[st]++;
    }
    if (launchesPerStage) for (int i = 0; i < 6; i++) launchesPerStage[i] = n[i];
    evUsed_ = 0;
    return true;
}

bool Batch::launchPicture(const StreamJob *dJobs, const StreamJob *dJobsFilter, uint32_t maxA, uint32_t maxB, uint32_t maxE, bool recon, bool deblock) {
    serial_++;
    auto mark = [&](int stageEnded) {
        if (!timing_) return;
        cudaEvent_t e = nextEvent();
        evStage_[evUsed_ - 1] = stageEnded;
        cudaEventRecord(e, stream_);
    };
    mark(-1);
    // The boundary strengths read records only: outside the per-kernel timing mode they run on a second stream next to
    // pass A.  Fork after the previous picture's border, join before the filter.
    const bool fork = !timing_ && recon && deblock && auxStream_ != nullptr;
    ReconParams rp;
    if (recon) {
        rp.pool = pool_; rp.g = g_; rp.jobs = dJobs; rp.done = dDoneRecon_;
        rp.ticket = dCounters_ + 0; rp.ticketA = dCounters_ + 2; rp.errors = dCounters_ + 8; rp.serial = serial_;
        rp.multiCount = dCounters_ + 4; rp.multiList = dMultiList_;
        rp.chunkRows = chunkRows_;
        rp.chunksPerCol = chunksPerCol_;
        rp.totalChunks = chunksPerCol_ * (uint32_t)g_.widthMbs * (uint32_t)g_.nStreams;
    }
    DeblockParams dp;
    if (deblock) {
        dp.pool = pool_; dp.g = g_; dp.jobs = dJobsFilter; dp.done = dDoneDeblock_;
        dp.ticket = dCounters_ + 1; dp.serial = serial_;
        dp.totalTickets = (uint32_t)g_.heightMbs * (((uint32_t)g_.nStreams + 1u) / 2u);
        dp.bsWords = dBsWords_; dp.work = dWork_;
        dp.workCount = reinterpret_cast<unsigned long long *>(dCounters_ + 10);
    }
    auto launchStrength = [&](cudaStream_t st) {
        const uint32_t chunks = ((uint32_t)g_.nMbs + kDeblockWarps * 32 - 1) / (kDeblockWarps * 32) * (uint32_t)g_.nStreams;
        strengthKernel<<<std::min<uint32_t>(chunks, (uint32_t)strengthBlocks_), kDeblockWarps * 32, 0, st>>>(dp);
        launches_++;
    };
    if (fork) {
        CK(cudaEventRecord(forkEv_, stream_));
        CK(cudaStreamWaitEvent(auxStream_, forkEv_, 0));
        launchStrength(auxStream_);
        CK(cudaEventRecord(joinEv_, auxStream_));
    }
    if (recon) {
        if (maxA) {     // (an IDR picture has nothing for pass A)
            const uint32_t ctas = (rp.totalChunks + kPassAWarps - 1) / kPassAWarps;
            const PassAMaps &maps = *reinterpret_cast<const PassAMaps *>(maps_);
            passAKernel<<<std::min<uint32_t>(ctas, (uint32_t)passABlocks_), kPassAWarps * 32, sizeof(PassAWarpSmem) * kPassAWarps, stream_>>>(rp, maps);
            mark(0);
            passAMultiKernel<<<std::min<uint32_t>(ctas, (uint32_t)passABlocks_), kPassAWarps * 32, sizeof(MultiWarpSmem) * kPassAWarps, stream_>>>(rp, maps);
            launches_ += 2;
            mark(5);
        }
        if (maxB) {
            const uint32_t ctasB = ((uint32_t)g_.heightMbs * (uint32_t)g_.nStreams + kReconWarps - 1) / kReconWarps;
            reconIntraKernel<<<std::min<uint32_t>(ctasB, (uint32_t)intraBlocks_), kReconWarps * 32, 0, stream_>>>(rp);
            launches_++;
            mark(3);
        }
        if (maxE) {
            // error path only: macroblocks that never arrived, estimated from their neighbours once everything else of the
            // picture is reconstructed (and before its filter)
            concealKernel<<<((uint32_t)g_.nStreams + kConcealWarps - 1) / kConcealWarps, kConcealWarps * 32, 0, stream_>>>(rp);
            launches_++;
            mark(3);
        }
    }
    if (deblock) {
        if (fork) {
            CK(cudaStreamWaitEvent(stream_, joinEv_, 0));
        } else {
            launchStrength(stream_);
            mark(4);
        }
        const uint32_t ctas = (dp.totalTickets + kDeblockWarps - 1) / kDeblockWarps;
        deblockKernel<<<std::min<uint32_t>(ctas, (uint32_t)deblockBlocks_), kDeblockWarps * 32, 0, stream_>>>(dp);
        launches_++;
        mark(1);
    }
    {
        BorderParams bp;
        bp.pool = pool_; bp.g = g_; bp.jobs = dJobs;
        const long long tasks = (long long)borderTasksPerStream(g_) * g_.nStreams;
        borderKernel<<<(int)((tasks + 7) / 8), 256, 0, stream_>>>(bp);
        launches_++;
        mark(2);
    }
    CK(cudaMemsetAsync(dCounters_, 0, 5 * sizeof(uint32_t), stream_));
    CK(cudaGetLastError());
    return true;
}

bool Batch::decodePicture(uint32_t k) {
    if (!created_) return false;
    CK(cudaSetDevice(device_));
    if (jobsDirty_ && !buildJobs()) return false;
    if (k >= numPics_) return false;
    // streamed uploads: this picture needs every fence up to the first one that covers it
    while (!fences_.empty()) {
        const bool covers = fences_.front().first > k;
        CK(cudaStreamWaitEvent(stream_, fences_.front().second, 0));
        fenceFree_.push_back(fences_.front().second);
        fences_.pop_front();
        if (covers) break;
    }
    return launchPicture(dJobs_ + (size_t)k * g_.nStreams, dJobs_ + jobsFilterAt_ + (size_t)k * g_.nStreams, picMaxA_[k], picMaxB_[k], picMaxE_[k], true, true);
}

bool Batch::debugStage(uint32_t k, bool recon, bool deblock) {
    if (!created_) return false;
    CK(cudaSetDevice(device_));
    if (jobsDirty_ && !buildJobs()) return false;
    if (k >= numPics_) return false;
    return launchPicture(dJobs_ + (size_t)k * g_.nStreams, dJobs_ + jobsFilterAt_ + (size_t)k * g_.nStreams, picMaxA_[k], picMaxB_[k], picMaxE_[k], recon, deblock);
}

bool Batch::run(uint32_t first, uint32_t count) {
    for (uint32_t k = first; k < first + count; k++)
        if (!decodePicture(k)) return false;
    return true;
}

bool Batch::sync() {
    if (!created_) return false;
    CK(cudaSetDevice(device_));
    // blocking-sync events: the waiting host thread sleeps instead of spinning (the host cores are busy parsing)
    if (!syncEv_) CK(cudaEventCreateWithFlags(&syncEv_, cudaEventBlockingSync | cudaEventDisableTiming));
    CK(cudaEventRecord(syncEv_, stream_));
    CK(cudaEventSynchronize(syncEv_));
    if (copyStream_) {
        CK(cudaEventRecord(syncEv_, copyStream_));
        CK(cudaEventSynchronize(syncEv_));
    }
    if (uploadStream_) {
        CK(cudaEventRecord(syncEv_, uploadStream_));
        CK(cudaEventSynchronize(syncEv_));
    }
    return true;
}

bool Batch::timerStart() { CK(cudaSetDevice(device_)); CK(cudaEventRecord(evA_, stream_)); return true; }
bool Batch::timerStop(float *ms) {
    CK(cudaSetDevice(device_));
    CK(cudaEventRecord(evB_, stream_));
    CK(cudaEventSynchronize(evB_));
    CK(cudaEventElapsedTime(ms, evA_, evB_));
    return true;
}

// streaming (single picture, host buffers): the legacy API path
bool Batch::submitHostPicture(uint32_t stream, const b200_pic_hdr &hdr, const b200_mb_rec *recs, const int16_t *coefs, const uint16_t *order,
                              const b200_mb_rec *filterRecs) {
    if (!created_ || g_.nStreams != 1 || stream != 0) return false;
    CK(cudaSetDevice(device_));
    const size_t recBytes = (size_t)g_.nMbs * sizeof(b200_mb_rec);
    const size_t coefBytes = (size_t)hdr.numCoefBlocks * B200_COEF_BLOCK_BYTES;
    const size_t orderBytes = (size_t)hdr.numConceal * sizeof(uint16_t);     // the concealment order
    const size_t need = recBytes + coefBytes + orderBytes + (filterRecs ? recBytes + 256 : 0) + 1024;
    const int b = stageIdx_ ^= 1;
    if (stageCap_[b] < need) {
        // growing a staging buffer: make sure nothing in flight still reads it
        CK(cudaStreamSynchronize(stream_));
        if (hStage_[b]) cudaFreeHost(hStage_[b]);
        cudaFree(dStage_[b]);
        hStage_[b] = nullptr; dStage_[b] = nullptr;
        stageCap_[b] = 0;
        const size_t cap = need + need / 2;
        CK(cudaMallocHost(&hStage_[b], cap));
        CK(cudaMalloc(&dStage_[b], cap));
        stageCap_[b] = cap;
        if (!stageEv_[b]) CK(cudaEventCreateWithFlags(&stageEv_[b], cudaEventDisableTiming));
    } else if (stageEv_[b]) {
        CK(cudaEventSynchronize(stageEv_[b]));  // the picture that last used this buffer has been consumed
    }
    uint8_t *h = hStage_[b];
    StreamJob job;
    const size_t recOff = 256, orderOff = (recOff + recBytes + 255) & ~(size_t)255, coefOff = (orderOff + orderBytes + 255) & ~(size_t)255;
    job.recs = reinterpret_cast<const b200_mb_rec *>(dStage_[b] + recOff);
    job.coefs = reinterpret_cast<const int16_t *>(dStage_[b] + coefOff);
    job.orderE = reinterpret_cast<const uint16_t *>(dStage_[b] + orderOff);
    job.curSlot = (uint16_t)hdr.curSlot;
    job.nB = (uint16_t)hdr.numPassB;
    job.nE = (uint16_t)hdr.numConceal;
    job.pad = 0;
    // the job the filter kernels get sits 64 bytes behind: the same, unless the picture has records of its own for the filter
    StreamJob jobF = job;
    size_t end = coefOff + coefBytes;
    if (filterRecs) {
        const size_t fOff = (end + 255) & ~(size_t)255;
        jobF.recs = reinterpret_cast<const b200_mb_rec *>(dStage_[b] + fOff);
        std::memcpy(h + fOff, filterRecs, recBytes);
        end = fOff + recBytes;
    }
    std::memcpy(h, &job, sizeof job);
    std::memcpy(h + 64, &jobF, sizeof jobF);
    std::memcpy(h + recOff, recs, recBytes);
    if (orderBytes) std::memcpy(h + orderOff, order, orderBytes);
    if (coefBytes) std::memcpy(h + coefOff, coefs, coefBytes);
    CK(cudaMemcpyAsync(dStage_[b], h, end, cudaMemcpyHostToDevice, stream_));
    h2dBytes_ += end;
    if (!launchPicture(reinterpret_cast<const StreamJob *>(dStage_[b]), reinterpret_cast<const StreamJob *>(dStage_[b] + 64), hdr.numPassA, hdr.numPassB, hdr.numConceal, true, true)) return false;
    CK(cudaEventRecord(stageEv_[b], stream_));
    return true;
}

bool Batch::ensureFrameStage(size_t bytes) {
    if (frameStageCap_ >= bytes) return true;
    CK(cudaStreamSynchronize(stream_));
    cudaFree(dFrameStage_);
    dFrameStage_ = nullptr; frameStageCap_ = 0;
    CK(cudaMalloc(&dFrameStage_, bytes));
    frameStageCap_ = bytes;
    return true;
}

// strip layout -> planar: streams [firstStream, firstStream + nStreams), the frame slot of each stream's job (or `slot`)
void Batch::launchPack(cudaStream_t st, const StreamJob *jobs, uint32_t slot, uint32_t firstStream, uint32_t nStreams, uint8_t *out, size_t outStride,
                       int cropX, int cropY, int cropW, int cropH, int nv12) {
    PackParams pp;
    pp.pool = pool_ + (unsigned long long)firstStream * g_.numSlots * g_.frameStride;
    pp.g = g_; pp.jobs = jobs; pp.slot = slot; pp.out = out; pp.outStride = outStride;
    pp.cropX = cropX; pp.cropY = cropY; pp.cropW = cropW; pp.cropH = cropH; pp.nv12 = nv12;
    const int units = g_.widthMbs * g_.H;
    dim3 grid((unsigned)std::min(64, (units + 255) / 256), nStreams);
    packKernel<<<grid, 256, 0, st>>>(pp);
    launches_++;
}

bool Batch::readFrame(uint32_t stream, uint32_t slot, uint8_t *dst) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || slot >= (uint32_t)g_.numSlots) return false;
    CK(cudaSetDevice(device_));
    const size_t fb = frameBytes();
    if (!ensureFrameStage(fb)) return false;
    launchPack(stream_, nullptr, slot, stream, 1, dFrameStage_, fb, 0, 0, g_.W, g_.H, 0);
    CK(cudaMemcpyAsync(dst, dFrameStage_, fb, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    d2hBytes_ += fb;
    return true;
}

bool Batch::mirrorFrameAsync(uint32_t slot, uint8_t *dst) {
    if (!created_ || g_.nStreams != 1 || slot >= (uint32_t)g_.numSlots || !dst) return false;
    CK(cudaSetDevice(device_));
    const size_t fb = frameBytes();
    if (!dMirror_) {
        CK(cudaMalloc(&dMirror_, fb * g_.numSlots));
        mirrorEv_.assign(g_.numSlots, nullptr);
        mirrorBusy_.assign(g_.numSlots, 0);
        for (auto &e : mirrorEv_) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (!copyStream_) CK(cudaStreamCreateWithFlags(&copyStream_, cudaStreamNonBlocking));
        if (!packedEv_[0]) CK(cudaEventCreateWithFlags(&packedEv_[0], cudaEventDisableTiming));
    }
    // (the slot's previous copy has been waited for by the caller or is overwritten in stream order: the staging is per slot and
    // a slot is decoded into again only after its picture has left the DPB)
    if (mirrorBusy_[slot]) CK(cudaStreamWaitEvent(stream_, mirrorEv_[slot], 0));
    launchPack(stream_, nullptr, slot, 0, 1, dMirror_ + fb * slot, fb, 0, 0, g_.W, g_.H, 0);
    CK(cudaEventRecord(packedEv_[0], stream_));
    CK(cudaStreamWaitEvent(copyStream_, packedEv_[0], 0));
    CK(cudaMemcpyAsync(dst, dMirror_ + fb * slot, fb, cudaMemcpyDeviceToHost, copyStream_));   // next to the following picture's kernels
    CK(cudaEventRecord(mirrorEv_[slot], copyStream_));
    mirrorBusy_[slot] = 1;
    d2hBytes_ += fb;
    return true;
}

bool Batch::waitMirror(uint32_t slot) {
    if (!created_ || slot >= mirrorEv_.size() || !mirrorBusy_[slot]) return false;
    CK(cudaSetDevice(device_));
    CK(cudaEventSynchronize(mirrorEv_[slot]));
    return true;
}

// test hook: put an I420 picture into a slot (and replicate its border) so a stage can be checked in isolation
bool Batch::writeFrame(uint32_t stream, uint32_t slot, const uint8_t *src) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || slot >= (uint32_t)g_.numSlots) return false;
    CK(cudaSetDevice(device_));
    const size_t fb = frameBytes();
    if (!ensureFrameStage(fb + 256)) return false;
    uint8_t *f = pool_ + ((unsigned long long)stream * g_.numSlots + slot) * g_.frameStride;
    CK(cudaMemcpyAsync(dFrameStage_, src, fb, cudaMemcpyHostToDevice, stream_));
    unpackKernel<<<64, 256, 0, stream_>>>(f, g_, dFrameStage_);
    // border: a one-stream BorderParams view whose "stream 0" is this frame (the job sits behind the picture in the staging)
    StreamJob job;
    std::memset(&job, 0, sizeof job);
    StreamJob *dJob = reinterpret_cast<StreamJob *>(dFrameStage_ + ((fb + 63) & ~(size_t)63));
    CK(cudaMemcpyAsync(dJob, &job, sizeof job, cudaMemcpyHostToDevice, stream_));
    BorderParams bp;
    bp.pool = f; bp.g = g_; bp.g.nStreams = 1; bp.jobs = dJob;
    borderKernel<<<(borderTasksPerStream(g_) + 7) / 8, 256, 0, stream_>>>(bp);
    CK(cudaStreamSynchronize(stream_));
    return true;
}

bool Batch::convertFrame(uint32_t stream, uint32_t slot, int mode, uint32_t *dstHost) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || slot >= (uint32_t)g_.numSlots || mode < 0 || mode > 2) return false;
    CK(cudaSetDevice(device_));
    const size_t bytes = (size_t)g_.W * g_.H * 4;
    if (convertCap_ < bytes) {
        CK(cudaStreamSynchronize(stream_));
        cudaFree(dConvert_);
        dConvert_ = nullptr; convertCap_ = 0;
        CK(cudaMalloc(&dConvert_, bytes));
        convertCap_ = bytes;
    }
    const uint8_t *f = pool_ + ((unsigned long long)stream * g_.numSlots + slot) * g_.frameStride;
    dim3 grid((g_.W / 8 + 255) / 256, g_.H, 1);
    convertFrameKernel<<<grid, 256, 0, stream_>>>(f, g_, mode, dConvert_, 0, 0, 0, 0, g_.W, g_.H);
    launches_++;
    CK(cudaMemcpyAsync(dstHost, dConvert_, bytes, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    d2hBytes_ += bytes;
    return true;
}

// device-only conversion of one frame `reps` times (config 5 throughput); returns elapsed ms
bool Batch::convertBench(uint32_t stream, uint32_t slot, int mode, int reps, float *ms) {
    if (!created_ || stream >= (uint32_t)g_.nStreams || slot >= (uint32_t)g_.numSlots) return false;
    CK(cudaSetDevice(device_));
    const size_t bytes = (size_t)g_.W * g_.H * 4;
    if (convertCap_ < bytes) {
        CK(cudaStreamSynchronize(stream_));
        cudaFree(dConvert_);
        dConvert_ = nullptr; convertCap_ = 0;
        CK(cudaMalloc(&dConvert_, bytes));
        convertCap_ = bytes;
    }
    const uint8_t *f = pool_ + ((unsigned long long)stream * g_.numSlots + slot) * g_.frameStride;
    dim3 grid((g_.W / 8 + 255) / 256, g_.H, 1);
    CK(cudaEventRecord(evA_, stream_));
    for (int i = 0; i < reps; i++) convertFrameKernel<<<grid, 256, 0, stream_>>>(f, g_, mode, dConvert_, 0, 0, 0, 0, g_.W, g_.H);
    CK(cudaEventRecord(evB_, stream_));
    CK(cudaEventSynchronize(evB_));
    CK(cudaEventElapsedTime(ms, evA_, evB_));
    launches_ += reps;
    return true;
}

// config 5 at batch size: frame `slot` of EVERY stream -> 32-bit pixels in one launch (blockIdx.z = stream), `reps` times
bool Batch::convertBenchAll(uint32_t slot, int mode, int reps, float *ms) {
    if (!created_ || slot >= (uint32_t)g_.numSlots || reps <= 0) return false;
    CK(cudaSetDevice(device_));
    const size_t pixels = (size_t)g_.W * g_.H;
    if (!dConvertAll_) CK(cudaMalloc(&dConvertAll_, pixels * 4 * g_.nStreams));
    const uint8_t *f = pool_ + (unsigned long long)slot * g_.frameStride;
    dim3 grid((g_.W / 8 + 255) / 256, g_.H, g_.nStreams);
    const unsigned long long inStride = (unsigned long long)g_.numSlots * g_.frameStride;
    CK(cudaEventRecord(evA_, stream_));
    for (int i = 0; i < reps; i++)
        convertFrameKernel<<<grid, 256, 0, stream_>>>(f, g_, mode, dConvertAll_, inStride, (unsigned long long)pixels, 0, 0, g_.W, g_.H);
    CK(cudaEventRecord(evB_, stream_));
    CK(cudaEventSynchronize(evB_));
    CK(cudaEventElapsedTime(ms, evA_, evB_));
    launches_ += reps;
    return true;
}

// h264bsdConvertTo{RGBA,BGRA,YCbCrA} for caller-owned host buffers (contiguous I420 in, u32 pixels out).  The device buffers
// are kept between calls (grown when a larger picture comes): a caller converts every picture of a stream with the same size.
bool convertHostI420(int mode, uint32_t width, uint32_t height, const uint8_t *yuv, uint32_t *out) {
    if (deviceCount() <= 0) {
        std::fprintf(stderr, "h264bsd_b200: no CUDA device visible -- this engine has no CPU fallback\n");
        return false;
    }
    if (!width || !height || (width & 3)) return false;
    const size_t inBytes = (size_t)width * height * 3 / 2, outBytes = (size_t)width * height * 4;
    static thread_local uint8_t *dIn = nullptr;
    static thread_local uint32_t *dOut = nullptr;
    static thread_local size_t capIn = 0, capOut = 0;
    if (capIn < inBytes) { cudaFree(dIn); dIn = nullptr; capIn = 0; CK(cudaMalloc(&dIn, inBytes)); capIn = inBytes; }
    if (capOut < outBytes) { cudaFree(dOut); dOut = nullptr; capOut = 0; CK(cudaMalloc(&dOut, outBytes)); capOut = outBytes; }
    CK(cudaMemcpy(dIn, yuv, inBytes, cudaMemcpyHostToDevice));
    const uint8_t *yPlane = dIn, *cbPlane = dIn + (size_t)width * height, *crPlane = dIn + (size_t)width * height * 5 / 4;
    const int W = (int)width, pitchC = W / 2;
    const bool wide = (W % 8 == 0) && ((uintptr_t)cbPlane % 4 == 0) && ((uintptr_t)crPlane % 4 == 0) && (pitchC % 4 == 0);
    if (wide) {
        dim3 grid((W / 8 + 255) / 256, height);
        convertKernelT<8><<<grid, 256, 0, nullptr>>>(yPlane, W, cbPlane, crPlane, pitchC, W, mode, dOut);
    } else {
        dim3 grid((W / 4 + 255) / 256, height);
        convertKernelT<4><<<grid, 256, 0, nullptr>>>(yPlane, W, cbPlane, crPlane, pitchC, W, mode, dOut);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dOut, outBytes, cudaMemcpyDeviceToHost));
    return true;
}

// picture k's frame of every stream (frame slot of that picture), de-stripped into one contiguous device staging buffer and
// sent to the host in a single transfer: dst + s * strideBytes receives stream s's I420 (or NV12) picture -- the coded size, or
// the cropping rectangle (cropW x cropH luma pels at (cropX, cropY))
bool Batch::readPictureAllEx(uint32_t k, uint8_t *dst, size_t strideBytes, int cropX, int cropY, int cropW, int cropH, int nv12) {
    if (!created_ || k >= numPics_) return false;
    if (!cropW) { cropX = cropY = 0; cropW = g_.W; cropH = g_.H; }
    if (cropX < 0 || cropY < 0 || cropW <= 0 || cropH <= 0 || cropX + cropW > g_.W || cropY + cropH > g_.H || ((cropX | cropY | cropW | cropH) & 1)) return false;
    CK(cudaSetDevice(device_));
    const size_t fb = frameBytes(), ob = (size_t)cropW * cropH * 3 / 2;
    if (strideBytes < ob) return false;
    if (!dPack_[0]) {
        CK(cudaMalloc(&dPack_[0], fb * g_.nStreams));
        CK(cudaMalloc(&dPack_[1], fb * g_.nStreams));
        CK(cudaEventCreateWithFlags(&packEv_[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&packEv_[1], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&packedEv_[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&packedEv_[1], cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&copyStream_, cudaStreamNonBlocking));
    }
    const int b = packIdx_ ^= 1;
    // the transfer that last used this staging buffer must be done before it is overwritten (device-side wait)
    if (packUsed_[b]) CK(cudaStreamWaitEvent(stream_, packEv_[b], 0));
    launchPack(stream_, dJobs_ + (size_t)k * g_.nStreams, 0, 0, (uint32_t)g_.nStreams, dPack_[b], ob, cropX, cropY, cropW, cropH, nv12);
    // the transfer runs on its own stream so that the next picture's kernels overlap it
    CK(cudaEventRecord(packedEv_[b], stream_));
    CK(cudaStreamWaitEvent(copyStream_, packedEv_[b], 0));
    if (strideBytes == ob) {
        CK(cudaMemcpyAsync(dst, dPack_[b], ob * g_.nStreams, cudaMemcpyDeviceToHost, copyStream_));
    } else {
        CK(cudaMemcpy2DAsync(dst, strideBytes, dPack_[b], ob, ob, g_.nStreams, cudaMemcpyDeviceToHost, copyStream_));
    }
    CK(cudaEventRecord(packEv_[b], copyStream_));
    packUsed_[b] = true;
    d2hBytes_ += ob * g_.nStreams;
    return true;
}
bool Batch::readPictureAll(uint32_t k, uint8_t *dst, size_t strideBytes) { return readPictureAllEx(k, dst, strideBytes, 0, 0, 0, 0, 0); }

// number of streams whose frame in slots[s] differs from stream 0's frame in slots[0]
int Batch::compareStreams(const uint32_t *slots) {
    if (!created_) return -1;
    if (cudaSetDevice(device_) != cudaSuccess) return -1;
    if (g_.nStreams == 1) return 0;
    uint32_t *dMis = nullptr;
    if (cudaMalloc(&dMis, sizeof(uint32_t) * g_.nStreams) != cudaSuccess) return -1;
    cudaMemsetAsync(dMis, 0, sizeof(uint32_t) * g_.nStreams, stream_);
    cudaMemcpyAsync(dSlots_, slots, sizeof(uint32_t) * g_.nStreams, cudaMemcpyHostToDevice, stream_);
    dim3 grid(64, g_.nStreams - 1);
    compareKernel<<<grid, 256, 0, stream_>>>(pool_, g_, dSlots_, dMis);
    std::vector<uint32_t> mis(g_.nStreams);
    cudaMemcpyAsync(mis.data(), dMis, sizeof(uint32_t) * g_.nStreams, cudaMemcpyDeviceToHost, stream_);
    if (cudaStreamSynchronize(stream_) != cudaSuccess) { cudaFree(dMis); return -1; }
    cudaFree(dMis);
    int bad = 0;
    for (int s = 1; s < g_.nStreams; s++) bad += mis[s] != 0;
    return bad;
}

uint32_t Batch::watchdog(int which) {
    uint32_t v[4] = {0, 0, 0, 0};
    if (!created_) return 0;
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    cudaMemcpyFromSymbol(v, gWatchdog, sizeof v);
    return v[which & 3];
}

uint64_t Batch::deblockWorkMbs() {
    unsigned long long v = 0;
    if (!created_) return 0;
    cudaSetDevice(device_);
    if (auxStream_) cudaStreamSynchronize(auxStream_);
    cudaMemcpyAsync(&v, dCounters_ + 10, sizeof v, cudaMemcpyDeviceToHost, stream_);
    cudaStreamSynchronize(stream_);
    return v;
}

uint32_t Batch::idctErrors() {
    uint32_t v = 0;
    if (!created_) return 0;
    cudaSetDevice(device_);
    cudaMemcpyAsync(&v, dCounters_ + 8, sizeof v, cudaMemcpyDeviceToHost, stream_);
    cudaStreamSynchronize(stream_);
    return v;
}

}  // namespace b200
