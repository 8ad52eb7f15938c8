// engine.hpp -- host-visible interface of the B200 reconstruction engine (see engine.cu)
#pragma once
#include <cstdint>
#include <atomic>
#include <deque>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "pool_geom.hpp"
#include "h264bsd_b200_tape.h"

namespace b200 {

int deviceCount();
bool convertHostI420(int mode, uint32_t width, uint32_t height, const uint8_t *yuv, uint32_t *out);

class Batch {
public:
    Batch() = default;
    ~Batch();
    Batch(const Batch &) = delete;
    Batch &operator=(const Batch &) = delete;

    bool create(int device, uint32_t nStreams, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots);
    void destroy();

    bool uploadTape(uint32_t stream, const b200_tape *t);
    // the same on a caller-owned CUDA stream; safe to call from several host threads at once for DISTINCT streams (the caller
    // waits for `st` before it re-uses the tape's memory, and calls tapesChanged() once when all uploads are queued)
    bool uploadTapeOn(uint32_t stream, const b200_tape *t, cudaStream_t st);
    void tapesChanged() { jobsDirty_ = true; }
    bool replicateTape(uint32_t srcStream);
    bool uploadTapeRange(uint32_t stream, const b200_tape *t, uint32_t firstPic, uint32_t numPics);
    bool uploadFence(uint32_t throughPic);
    bool convertBenchAll(uint32_t slot, int mode, int reps, float *ms);
    uint64_t deblockWorkMbs();   // macroblocks with a non-zero boundary strength since creation
    bool decodePicture(uint32_t k);                // picture k of every stream
    bool run(uint32_t first, uint32_t count);
    bool debugStage(uint32_t k, bool recon, bool deblock);
    bool sync();
    bool timerStart();
    bool timerStop(float *ms);

    bool submitHostPicture(uint32_t stream, const b200_pic_hdr &hdr, const b200_mb_rec *recs, const int16_t *coefs, const uint16_t *order,
                           const b200_mb_rec *filterRecs = nullptr);
    bool readFrame(uint32_t stream, uint32_t slot, uint8_t *dst);
    // one-stream batches (the legacy API): start the de-stripped copy of frame `slot` into page-locked `dst` behind the work queued
    // so far and return at once; waitMirror(slot) blocks until that copy has landed
    bool mirrorFrameAsync(uint32_t slot, uint8_t *dst);
    bool waitMirror(uint32_t slot);
    // picture k's frame of EVERY stream -> dst + s * strideBytes (asynchronous; dst should be pinned; sync() to wait)
    bool readPictureAll(uint32_t k, uint8_t *dst, size_t strideBytes);
    bool writeFrame(uint32_t stream, uint32_t slot, const uint8_t *src);
    bool convertFrame(uint32_t stream, uint32_t slot, int mode, uint32_t *dstHost);
    bool convertBench(uint32_t stream, uint32_t slot, int mode, int reps, float *ms);
    int compareStreams(const uint32_t *slots);
    // per-stage device time (CUDA events on the engine's stream around every launch)
    void kernelTiming(bool enable);
    bool kernelTimes(float ms[6], uint32_t *launchesPerStage);  // recon pass A (first instance), deblock filter, border, recon pass B, boundary strengths, pass A (second instance: several partitions)
    // picture k's frame of every stream, cropped / as NV12 (see packKernel); cropW == 0: the coded size
    bool readPictureAllEx(uint32_t k, uint8_t *dst, size_t strideBytes, int cropX, int cropY, int cropW, int cropH, int nv12);
    uint32_t idctErrors();
    uint32_t watchdog(int which);  // 0: flag waits that gave up, 1: TMA waits that gave up

    const PoolGeom &geom() const { return g_; }
    uint32_t numPics() const { return numPics_; }
    uint64_t launches() const { return launches_; }
    uint64_t h2dBytes() const { return h2dBytes_.load(); }
    uint64_t d2hBytes() const { return d2hBytes_; }
    size_t frameBytes() const { return (size_t)g_.nMbs * 384; }
    const std::vector<b200_pic_hdr> &pics(uint32_t stream) const { return tapes_[stream].pics; }
    int device() const { return device_; }

private:
    void resetState();   // destroy() only: every member back to its default (nothing dangles after a re-create)
    struct DevTape {
        uint8_t *recs = nullptr, *coefs = nullptr, *order = nullptr;
        size_t recBytes = 0, coefBytes = 0, orderBytes = 0, capRecs = 0, capCoefs = 0, capOrder = 0;
        bool owned = false;
        std::vector<b200_pic_hdr> pics;
    };
    bool buildJobs();
    bool launchPicture(const StreamJob *dJobs, const StreamJob *dJobsFilter, uint32_t maxA, uint32_t maxB, uint32_t maxE, bool recon, bool deblock);
    bool ensureFrameStage(size_t bytes);
    void launchPack(cudaStream_t st, const StreamJob *jobs, uint32_t slot, uint32_t firstStream, uint32_t nStreams, uint8_t *out, size_t outStride,
                    int cropX, int cropY, int cropW, int cropH, int nv12);
    std::vector<uint32_t> picMaxA_, picMaxB_, picMaxE_;

    bool created_ = false;
    int device_ = 0, numSms_ = 0;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t evA_ = nullptr, evB_ = nullptr;
    PoolGeom g_{};
    uint8_t *pool_ = nullptr;
    CUtensorMap maps_[10];             // PassAMaps (recon_kernel.cuh): luma [nx 1..3][16 / 21 rows], chroma [nx 1..2][8 / 9 rows]
    uint32_t *dDoneRecon_ = nullptr, *dDoneDeblock_ = nullptr, *dCounters_ = nullptr, *dSlots_ = nullptr, *dBsWords_ = nullptr, *dMultiList_ = nullptr;
    uint8_t *dWork_ = nullptr;
    int strengthBlocks_ = 0;
    uint32_t serial_ = 0;
    int passABlocks_ = 0, deblockBlocks_ = 0, intraBlocks_ = 0;
    uint32_t chunkRows_ = 32, chunksPerCol_ = 1;
    cudaEvent_t syncEv_ = nullptr, forkEv_ = nullptr, joinEv_ = nullptr;
    size_t jobsCap_ = 0;
    uint32_t *dConvertAll_ = nullptr;
    cudaStream_t uploadStream_ = nullptr;
    std::deque<std::pair<uint32_t, cudaEvent_t>> fences_;   // (pictures below this index, upload-stream event)
    std::vector<cudaEvent_t> fenceFree_;
    cudaStream_t auxStream_ = nullptr;   // boundary strengths next to pass A
    std::vector<DevTape> tapes_;
    StreamJob *dJobs_ = nullptr;        // per picture and stream; the second half of the table is what the filter kernels get
    size_t jobsFilterAt_ = 0;           // (the same jobs, except where a picture has records of its own for the filter)
    uint32_t numPics_ = 0;
    bool jobsDirty_ = true;
    // streaming staging (legacy single-stream API)
    uint8_t *hStage_[2] = {nullptr, nullptr}, *dStage_[2] = {nullptr, nullptr};
    size_t stageCap_[2] = {0, 0};
    cudaEvent_t stageEv_[2] = {nullptr, nullptr};
    int stageIdx_ = 0;
    uint32_t *dConvert_ = nullptr;
    size_t convertCap_ = 0;
    uint8_t *dFrameStage_ = nullptr;    // one picture, planar: readFrame / writeFrame
    uint8_t *dMirror_ = nullptr;        // numSlots pictures, planar: mirrorFrameAsync
    std::vector<cudaEvent_t> mirrorEv_; // per slot: its copy to the host has landed
    std::vector<char> mirrorBusy_;
    size_t frameStageCap_ = 0;
    uint64_t launches_ = 0, d2hBytes_ = 0;
    std::atomic<uint64_t> h2dBytes_{0};
    bool timing_ = false;
    std::vector<cudaEvent_t> evPool_;
    size_t evUsed_ = 0;
    std::vector<int> evStage_;  // stage id of the interval that ENDS at event i (or -1)
    cudaEvent_t nextEvent();
    uint8_t *dPack_[2] = {nullptr, nullptr};
    cudaEvent_t packEv_[2] = {nullptr, nullptr}, packedEv_[2] = {nullptr, nullptr};
    cudaStream_t copyStream_ = nullptr;
    bool packUsed_[2] = {false, false};
    int packIdx_ = 0;
};

}  // namespace b200
