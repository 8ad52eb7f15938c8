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
#include <new>
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
    u32 r = d->dec.decode(byteStrm, len, picId, readBytes);
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
    if (!d->batch.readFrame(0, (uint32_t)o->slot, d->hostFrames[o->slot])) return nullptr;
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
    int rc = 0;
    if (t->mbRecBytes) rc |= cudaHostRegister(t->mbRecs, t->mbRecBytes, cudaHostRegisterPortable) != cudaSuccess;
    if (t->coefBytes) rc |= cudaHostRegister(t->coefs, t->coefBytes, cudaHostRegisterPortable) != cudaSuccess;
    rc |= cudaHostRegister(t->mbOrder, (size_t)t->numPics * t->widthMbs * t->heightMbs * 2, cudaHostRegisterPortable) != cudaSuccess;
    if (rc) cudaGetLastError();
    t->pinned = rc ? 0 : 1;
    return rc ? -1 : 0;
}
void h264bsdB200UnpinTape(b200_tape *t) {
    if (!t || t->pinned != 1) return;
    cudaHostUnregister(t->mbRecs);
    cudaHostUnregister(t->coefs);
    cudaHostUnregister(t->mbOrder);
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
#undef B

}  // extern "C"
