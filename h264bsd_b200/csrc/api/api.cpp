// api.cpp -- the C-ABI of libh264bsd_b200.so.
//
//  (1) the preserved single-stream API of oneam/h264bsd (include/h264bsd_decoder.h): the host
//      syntax decoder drives a one-stream Batch; every pel comes from the GPU.
//  (2) the batched API (include/h264bsd_b200.h): many independent streams per GPU.
//
// There is no CPU pixel path: h264bsdInit / h264bsdB200BatchCreate fail (and say so on stderr) when
// no CUDA device is usable.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "h264bsd_decoder.h"
#include "h264bsd_util.h"
#include "h264bsd_b200.h"
#include "../host/stream_decoder.hpp"
#include "../engine/engine.hpp"

using namespace b200;

namespace {

constexpr uint32_t kMagic = 0xB200264Du;

static int envDevice() {
    const char *e = std::getenv("H264BSD_B200_DEVICE");
    return e ? std::atoi(e) : 0;
}

struct LegacyDecoder : public PictureSink {
    StreamDecoder dec;
    Batch batch;
    std::vector<uint8_t *> hostFrames;  // pinned mirror per frame slot: what NextOutputPicture hands out
    uint32_t *conv = nullptr;           // pinned conversion buffer (h264bsd_storage.h: conversionBuffer)
    size_t convBytes = 0;
    bool failed = false;

    explicit LegacyDecoder(bool noReorder) : dec(this, noReorder) {}
    ~LegacyDecoder() override {
        freeHost();
    }
    void freeHost() {
        for (uint8_t *p : hostFrames) cudaFreeHost(p);
        hostFrames.clear();
        if (conv) cudaFreeHost(conv);
        conv = nullptr;
        convBytes = 0;
    }
    bool configure(uint32_t w, uint32_t h, uint32_t slots) override {
        batch.destroy();
        freeHost();
        if (!batch.create(envDevice(), 1, w, h, slots)) { failed = true; return false; }
        hostFrames.assign(slots, nullptr);
        for (uint32_t i = 0; i < slots; i++)
            if (cudaMallocHost(&hostFrames[i], batch.frameBytes()) != cudaSuccess) { failed = true; return false; }
        return true;
    }
    bool submitPicture(const b200_pic_hdr &hdr, const b200_mb_rec *recs, const int16_t *coefs, const uint16_t *order,
                       const b200_mb_rec *filterRecs) override {
        if (!batch.submitHostPicture(0, hdr, recs, coefs, order, filterRecs)) { failed = true; return false; }
        // the picture's way to the host starts now, behind its kernels; h264bsdNextOutputPicture only waits for it
        if (hdr.curSlot < hostFrames.size() && !batch.mirrorFrameAsync(hdr.curSlot, hostFrames[hdr.curSlot])) { failed = true; return false; }
        return true;
    }
};

LegacyDecoder *self(storage_t *s) {
    if (!s || s->u.b200.magic != kMagic) return nullptr;
    return static_cast<LegacyDecoder *>(s->u.b200.engine);
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// preserved API
// ---------------------------------------------------------------------------------------------
u32 h264bsdInit(storage_t *pStorage, u32 noOutputReordering) {
    if (!pStorage) return HANTRO_NOK;
    std::memset(pStorage, 0, sizeof *pStorage);
    if (deviceCount() <= 0) {
        std::fprintf(stderr, "h264bsd_b200: h264bsdInit: no CUDA device visible; this build has no CPU pixel path\n");
        return HANTRO_NOK;
    }
    LegacyDecoder *d = new (std::nothrow) LegacyDecoder(noOutputReordering != 0);
    if (!d) return HANTRO_NOK;
    pStorage->u.b200.engine = d;
    pStorage->u.b200.magic = kMagic;
    return HANTRO_OK;
}

u32 h264bsdDecode(storage_t *pStorage, u8 *byteStrm, u32 len, u32 picId, u32 *readBytes) {
    LegacyDecoder *d = self(pStorage);
    if (!d || !byteStrm || !readBytes) return H264BSD_ERROR;
    u32 r;
    try {
        r = d->dec.decode(byteStrm, len, picId, readBytes);
    } catch (const std::exception &) {     // an allocation that failed (or a size nothing could satisfy): no exception leaves the C-ABI
        return H264BSD_MEMALLOC_ERROR;
    }
    if (d->failed) return H264BSD_MEMALLOC_ERROR;
    return r;
}

void h264bsdShutdown(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    if (!d) return;
    delete d;
    pStorage->u.b200.engine = nullptr;
    pStorage->u.b200.magic = 0;
}

u8 *h264bsdNextOutputPicture(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs) {
    LegacyDecoder *d = self(pStorage);
    if (!d) return nullptr;
    const OutPic *o = d->dec.nextOutput();
    if (!o) return nullptr;
    if (picId) *picId = o->picId;
    if (isIdrPic) *isIdrPic = o->isIdr;
    if (numErrMbs) *numErrMbs = o->numErrMbs;
    if ((size_t)o->slot >= d->hostFrames.size()) return nullptr;
    if (!d->batch.waitMirror((uint32_t)o->slot)) return nullptr;
    return d->hostFrames[o->slot];
}

static u32 *nextConverted(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs, int mode) {
    LegacyDecoder *d = self(pStorage);
    if (!d) return nullptr;
    const OutPic *o = d->dec.nextOutput();
    if (!o) return nullptr;
    if (picId) *picId = o->picId;
    if (isIdrPic) *isIdrPic = o->isIdr;
    if (numErrMbs) *numErrMbs = o->numErrMbs;
    const size_t bytes = d->batch.frameBytes() / 384 * 256 * 4;
    if (d->convBytes < bytes) {
        if (d->conv) cudaFreeHost(d->conv);
        d->conv = nullptr;
        if (cudaMallocHost(&d->conv, bytes) != cudaSuccess) return nullptr;
        d->convBytes = bytes;
    }
    if (!d->batch.convertFrame(0, (uint32_t)o->slot, mode, d->conv)) return nullptr;
    return d->conv;
}
u32 *h264bsdNextOutputPictureRGBA(storage_t *s, u32 *a, u32 *b, u32 *c) { return nextConverted(s, a, b, c, 0); }
u32 *h264bsdNextOutputPictureBGRA(storage_t *s, u32 *a, u32 *b, u32 *c) { return nextConverted(s, a, b, c, 1); }
u32 *h264bsdNextOutputPictureYCbCrA(storage_t *s, u32 *a, u32 *b, u32 *c) { return nextConverted(s, a, b, c, 2); }

u32 h264bsdPicWidth(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    return d && d->dec.activeSps() ? d->dec.activeSps()->widthMbs : 0;
}
u32 h264bsdPicHeight(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    return d && d->dec.activeSps() ? d->dec.activeSps()->heightMbs : 0;
}
u32 h264bsdVideoRange(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    const Sps *s = d ? d->dec.activeSps() : nullptr;
    return (s && s->vuiPresent && s->vui.videoSignalTypePresent && s->vui.videoFullRange) ? 1 : 0;
}
u32 h264bsdMatrixCoefficients(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    const Sps *s = d ? d->dec.activeSps() : nullptr;
    if (s && s->vuiPresent && s->vui.videoSignalTypePresent && s->vui.colourDescriptionPresent) return s->vui.matrixCoefficients;
    return 2;
}
void h264bsdCroppingParams(storage_t *pStorage, u32 *croppingFlag, u32 *left, u32 *width, u32 *top, u32 *height) {
    LegacyDecoder *d = self(pStorage);
    const Sps *s = d ? d->dec.activeSps() : nullptr;
    if (s && s->cropping) {
        *croppingFlag = 1;
        *left = 2 * s->cropLeft;
        *width = 16 * s->widthMbs - 2 * (s->cropLeft + s->cropRight);
        *top = 2 * s->cropTop;
        *height = 16 * s->heightMbs - 2 * (s->cropTop + s->cropBottom);
    } else {
        *croppingFlag = 0; *left = 0; *width = 0; *top = 0; *height = 0;
    }
}
void h264bsdSampleAspectRatio(storage_t *pStorage, u32 *sarWidth, u32 *sarHeight) {
    LegacyDecoder *d = self(pStorage);
    const Sps *s = d ? d->dec.activeSps() : nullptr;
    u32 w = 1, h = 1;
    if (s && s->vuiPresent && s->vui.aspectRatioPresent) {
        // Table E-1 (decoder.c:1010-1055)
        static const u32 kSar[14][2] = {{0, 0}, {1, 1}, {12, 11}, {10, 11}, {16, 11}, {40, 33}, {24, 11},
                                       {20, 11}, {32, 11}, {80, 33}, {18, 11}, {15, 11}, {64, 33}, {160, 99}};
        u32 idc = s->vui.aspectRatioIdc;
        if (idc < 14) { w = kSar[idc][0]; h = kSar[idc][1]; }
        else if (idc == 255) { w = s->vui.sarWidth; h = s->vui.sarHeight; if (!w || !h) w = h = 0; }
        else { w = h = 0; }
    }
    *sarWidth = w;
    *sarHeight = h;
}
u32 h264bsdCheckValidParamSets(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    return d && d->dec.validParamSets() ? 1 : 0;
}
void h264bsdFlushBuffer(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    if (d) d->dec.flushBuffer();
}
u32 h264bsdProfile(storage_t *pStorage) {
    LegacyDecoder *d = self(pStorage);
    return d && d->dec.activeSps() ? d->dec.activeSps()->profileIdc : 0;
}
storage_t *h264bsdAlloc(void) { return (storage_t *)std::malloc(sizeof(storage_t)); }
void h264bsdFree(storage_t *pStorage) { std::free(pStorage); }

void h264bsdConvertToRGBA(u32 width, u32 height, u8 *data, u32 *pOutput) { convertHostI420(0, width, height, data, pOutput); }
void h264bsdConvertToBGRA(u32 width, u32 height, u8 *data, u32 *pOutput) { convertHostI420(1, width, height, data, pOutput); }
void h264bsdConvertToYCbCrA(u32 width, u32 height, u8 *data, u32 *pOutput) { convertHostI420(2, width, height, data, pOutput); }

// ---------------------------------------------------------------------------------------------
// batched API
// ---------------------------------------------------------------------------------------------
int h264bsdB200DeviceCount(void) { return deviceCount(); }

void *h264bsdB200HostAlloc(size_t bytes) {
    void *p = nullptr;
    if (deviceCount() <= 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void h264bsdB200HostFree(void *p) { if (p) cudaFreeHost(p); }
// page-lock a parsed tape's arrays so that uploads run at full PCIe speed (no-op without a device)
int h264bsdB200PinTape(b200_tape *t) {
    if (!t || deviceCount() <= 0) return -1;
    if (t->pinned == 1) return 0;
    // the whole allocations, not just the bytes in use: a tape that is re-used for a longer stream stays page-locked as long as
    // it fits (tape_builder.cpp unpins before an array has to move)
    int rc = 0;
    if (t->mbRecs && t->capRecs) rc |= cudaHostRegister(t->mbRecs, t->capRecs, cudaHostRegisterPortable) != cudaSuccess;
    if (!rc && t->coefs && t->capCoefs) rc |= cudaHostRegister(t->coefs, t->capCoefs, cudaHostRegisterPortable) != cudaSuccess;
    if (!rc && t->mbOrder && t->capOrder) rc |= cudaHostRegister(t->mbOrder, t->capOrder, cudaHostRegisterPortable) != cudaSuccess;
    if (rc) {
        cudaGetLastError();
        cudaHostUnregister(t->mbRecs); cudaHostUnregister(t->coefs); cudaHostUnregister(t->mbOrder);   // whatever did register
        cudaGetLastError();
    }
    t->pinned = rc ? 0 : 1;
    return rc ? -1 : 0;
}
void h264bsdB200UnpinTape(b200_tape *t) {
    if (!t || t->pinned != 1) return;
    if (t->mbRecs && t->capRecs) cudaHostUnregister(t->mbRecs);
    if (t->coefs && t->capCoefs) cudaHostUnregister(t->coefs);
    if (t->mbOrder && t->capOrder) cudaHostUnregister(t->mbOrder);
    cudaGetLastError();
    t->pinned = 0;
}

b200_batch *h264bsdB200BatchCreate(int device, uint32_t nStreams, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots) {
    Batch *b = new (std::nothrow) Batch();
    if (!b) return nullptr;
    if (!b->create(device, nStreams, widthMbs, heightMbs, numSlots)) {
        delete b;
        return nullptr;
    }
    return reinterpret_cast<b200_batch *>(b);
}
void h264bsdB200BatchDestroy(b200_batch *h) { delete reinterpret_cast<Batch *>(h); }

#define B(h) reinterpret_cast<Batch *>(h)

// ---- parse + upload in one go (the end-to-end path): every worker thread parses a stream into ITS re-used page-locked tape,
// queues the work-list's upload on ITS copy stream, waits for the copy and takes the next stream.  The host holds `threads`
// tapes instead of one per stream, nothing is issued from the caller's thread, and parsing overlaps the H2D copies.
struct b200_pu_worker {
    b200_tape *tape[2] = {nullptr, nullptr};     // two tapes: the upload of one overlaps the parse into the other
    cudaEvent_t done[2] = {nullptr, nullptr};    // upload of tape[j] finished
    bool busy[2] = {false, false};
    cudaStream_t st = nullptr;
};
struct b200_pu_pool {             // lives as long as the process: a worker slot keeps its tapes and stream across jobs
    std::vector<b200_pu_worker> w;
};
struct b200_pu_job {
    std::thread th;
    int failed = 0;
};
static void puRun(Batch *b, b200_pu_pool *pool, uint32_t n, const uint8_t *const *streams, const size_t *lens, uint32_t flags, uint32_t threads, int *failedOut) {
    std::atomic<uint32_t> next(0), failed(0);
    auto work = [&](uint32_t slot) {
        cudaSetDevice(b->device());
        b200_pu_worker &w = pool->w[slot];
        if (!w.st && (cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) != cudaSuccess ||
                      cudaEventCreateWithFlags(&w.done[0], cudaEventDisableTiming) != cudaSuccess ||
                      cudaEventCreateWithFlags(&w.done[1], cudaEventDisableTiming) != cudaSuccess)) { failed.fetch_add(n); return; }
        int j = 0;
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            if (w.busy[j]) { cudaEventSynchronize(w.done[j]); w.busy[j] = false; }   // the copy out of this tape two streams ago
            w.tape[j] = h264bsdB200ReparseStream(w.tape[j], streams[i], lens[i], flags);
            if (!w.tape[j] || w.tape[j]->status != 0) { failed.fetch_add(1); continue; }
            if (w.tape[j]->pinned != 1) h264bsdB200PinTape(w.tape[j]);     // first use, or an array had to grow
            if (!b->uploadTapeOn(i, w.tape[j], w.st) || cudaEventRecord(w.done[j], w.st) != cudaSuccess) { failed.fetch_add(1); continue; }
            w.busy[j] = true;
            j ^= 1;
        }
        // the job ends when every work-list has arrived
        if (cudaStreamSynchronize(w.st) != cudaSuccess) failed.fetch_add(1);
        w.busy[0] = w.busy[1] = false;
    };
    std::vector<std::thread> pool_;
    for (uint32_t t = 1; t < threads; t++) pool_.emplace_back(work, t);
    work(0);
    for (auto &t : pool_) t.join();
    b->tapesChanged();
    *failedOut = (int)failed.load();
}
int h264bsdB200BatchUploadTape(b200_batch *h, uint32_t stream, const b200_tape *tape) { return h && B(h)->uploadTape(stream, tape) ? 0 : -1; }
int h264bsdB200BatchUploadTapeRange(b200_batch *h, uint32_t stream, const b200_tape *tape, uint32_t firstPic, uint32_t numPics) {
    return h && B(h)->uploadTapeRange(stream, tape, firstPic, numPics) ? 0 : -1;
}
int h264bsdB200BatchUploadTapesRange(b200_batch *h, const b200_tape *const *tapes, uint32_t nStreams, uint32_t firstPic, uint32_t numPics) {
    if (!h || !tapes) return -1;
    for (uint32_t s = 0; s < nStreams; s++)
        if (!B(h)->uploadTapeRange(s, tapes[s], firstPic, numPics)) return -1;
    return B(h)->uploadFence(firstPic + numPics) ? 0 : -1;
}
int h264bsdB200BatchUploadFence(b200_batch *h, uint32_t throughPic) { return h && B(h)->uploadFence(throughPic) ? 0 : -1; }
int h264bsdB200BatchReplicateTape(b200_batch *h, uint32_t srcStream) { return h && B(h)->replicateTape(srcStream) ? 0 : -1; }
int h264bsdB200BatchDecodePicture(b200_batch *h, uint32_t picIndex) { return h && B(h)->decodePicture(picIndex) ? 0 : -1; }
int h264bsdB200BatchRun(b200_batch *h, uint32_t firstPic, uint32_t numPics) { return h && B(h)->run(firstPic, numPics) ? 0 : -1; }
int h264bsdB200BatchSync(b200_batch *h) { return h && B(h)->sync() ? 0 : -1; }
int h264bsdB200BatchTimerStart(b200_batch *h) { return h && B(h)->timerStart() ? 0 : -1; }
int h264bsdB200BatchTimerStop(b200_batch *h, float *ms) { return h && ms && B(h)->timerStop(ms) ? 0 : -1; }
int h264bsdB200BatchReadFrame(b200_batch *h, uint32_t stream, uint32_t slot, uint8_t *dst) { return h && B(h)->readFrame(stream, slot, dst) ? 0 : -1; }
int h264bsdB200BatchReadPictureAll(b200_batch *h, uint32_t picIndex, uint8_t *dst, size_t strideBytes) { return h && B(h)->readPictureAll(picIndex, dst, strideBytes) ? 0 : -1; }
int h264bsdB200BatchWriteFrame(b200_batch *h, uint32_t stream, uint32_t slot, const uint8_t *src) { return h && B(h)->writeFrame(stream, slot, src) ? 0 : -1; }
int h264bsdB200BatchConvertFrame(b200_batch *h, uint32_t stream, uint32_t slot, int mode, uint32_t *dst) { return h && B(h)->convertFrame(stream, slot, mode, dst) ? 0 : -1; }
int h264bsdB200BatchConvertBenchAll(b200_batch *h, uint32_t slot, int mode, int reps, float *ms) { return h && B(h)->convertBenchAll(slot, mode, reps, ms) ? 0 : -1; }
int h264bsdB200BatchConvertBench(b200_batch *h, uint32_t stream, uint32_t slot, int mode, int reps, float *ms) { return h && B(h)->convertBench(stream, slot, mode, reps, ms) ? 0 : -1; }
int h264bsdB200BatchCompareStreams(b200_batch *h, const uint32_t *slots) { return h ? B(h)->compareStreams(slots) : -1; }
int h264bsdB200BatchDebugStage(b200_batch *h, uint32_t picIndex, int recon, int deblock) { return h && B(h)->debugStage(picIndex, recon != 0, deblock != 0) ? 0 : -1; }
uint64_t h264bsdB200BatchDeblockWorkMbs(b200_batch *h) { return h ? B(h)->deblockWorkMbs() : 0; }
uint32_t h264bsdB200BatchIdctErrors(b200_batch *h) { return h ? B(h)->idctErrors() : 0; }
void h264bsdB200BatchKernelTiming(b200_batch *h, int enable) { if (h) B(h)->kernelTiming(enable != 0); }
int h264bsdB200BatchKernelTimes(b200_batch *h, float *ms6, uint32_t *launches6) { return h && ms6 && B(h)->kernelTimes(ms6, launches6) ? 0 : -1; }
uint32_t h264bsdB200BatchWatchdog(b200_batch *h, int which) { return h ? B(h)->watchdog(which) : 0; }
uint64_t h264bsdB200BatchLaunches(b200_batch *h) { return h ? B(h)->launches() : 0; }
uint64_t h264bsdB200BatchH2DBytes(b200_batch *h) { return h ? B(h)->h2dBytes() : 0; }
uint64_t h264bsdB200BatchD2HBytes(b200_batch *h) { return h ? B(h)->d2hBytes() : 0; }
uint32_t h264bsdB200BatchNumPics(b200_batch *h) { return h ? B(h)->numPics() : 0; }
b200_pu_pool *h264bsdB200ParseUploadPoolCreate(uint32_t threads) {
    b200_pu_pool *p = new (std::nothrow) b200_pu_pool();
    if (p) p->w.resize(threads ? threads : 1);
    return p;
}
void h264bsdB200ParseUploadPoolDestroy(b200_pu_pool *p) {
    if (!p) return;
    for (auto &w : p->w) {
        if (w.st) cudaStreamSynchronize(w.st);
        for (int j = 0; j < 2; j++) {
            if (w.tape[j]) h264bsdB200FreeTape(w.tape[j]);      // (unpins)
            if (w.done[j]) cudaEventDestroy(w.done[j]);
        }
        if (w.st) cudaStreamDestroy(w.st);
    }
    delete p;
}
b200_pu_job *h264bsdB200BatchParseUploadBegin(b200_batch *h, b200_pu_pool *pool, uint32_t n, const uint8_t *const *streams, const size_t *lens,
                                              uint32_t flags) {
    if (!h || !pool || !n || !streams || !lens || pool->w.empty()) return nullptr;
    b200_pu_job *j = new (std::nothrow) b200_pu_job();
    if (!j) return nullptr;
    Batch *b = B(h);
    const uint32_t threads = (uint32_t)std::min<size_t>(pool->w.size(), n);
    j->th = std::thread([=]() { puRun(b, pool, n, streams, lens, flags, threads, &j->failed); });
    return j;
}
int h264bsdB200BatchParseUploadWait(b200_pu_job *j) {
    if (!j) return -1;
    j->th.join();
    const int f = j->failed;
    delete j;
    return f;
}
int h264bsdB200BatchReadPictureAllEx(b200_batch *h, uint32_t picIndex, uint8_t *dst, size_t strideBytes, uint32_t cropX, uint32_t cropY,
                                     uint32_t cropW, uint32_t cropH, int nv12) {
    return h && B(h)->readPictureAllEx(picIndex, dst, strideBytes, (int)cropX, (int)cropY, (int)cropW, (int)cropH, nv12) ? 0 : -1;
}
#undef B

}  // extern "C"
