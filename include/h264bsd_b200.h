/*
 * h264bsd_b200.h -- batched C-ABI of the B200 H.264 Baseline reconstruction engine.
 *
 * Lives next to -- not instead of -- the preserved single-stream API (h264bsd_decoder.h).
 * Plain pointers and sizes only.  A "tape" is the pre-parsed work-list of one stream
 * (h264bsd_b200_tape.h); a "batch" is a set of independent streams resident on one GPU, the
 * unit of data parallelism (streams are sharded statically, no collective).
 */
#ifndef H264BSD_B200_H
#define H264BSD_B200_H

#include <stddef.h>
#include <stdint.h>
#include "h264bsd_b200_tape.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host only (no GPU needed) ------------------------------------------------------- */

/* Parse a whole Annex-B byte stream (all NAL/CAVLC/MV-prediction/DPB work the reference does in
 * h264bsdDecode, h264bsd_decoder.c:152-515) into a tape.  The input is not modified.
 * tape->status != 0 if the parse stopped on a decoder error.  NULL only on allocation failure.
 * noOutputReordering: bit 0 = the flag of h264bsdInit; bit 1 (B200_PARSE_RESILIENT) = carry on after H264BSD_ERROR the way
 * a player does: macroblocks missing from a picture are concealed at the next access unit boundary
 * (h264bsd_decoder.c:226-262, h264bsd_conceal.c) and show up in b200_pic_hdr.numErrMbs. */
#define B200_PARSE_RESILIENT 2u
b200_tape *h264bsdB200ParseStream(const uint8_t *stream, size_t len, uint32_t noOutputReordering);
/* same, re-using `tape`'s arrays (NULL: allocate a new tape).  Steady-state parsing then touches no fresh pages and a
 * page-locked tape stays page-locked unless an array had to grow (tape->pinned == 2: pin again). */
b200_tape *h264bsdB200ReparseStream(b200_tape *tape, const uint8_t *stream, size_t len, uint32_t noOutputReordering);
/* ReparseStream for n independent streams on `threads` host threads (tapes[i] may be NULL); returns the number of failures */
int h264bsdB200ReparseStreams(b200_tape **tapes, uint32_t n, const uint8_t *const *streams, const size_t *lens,
                              uint32_t noOutputReordering, uint32_t threads);
/* the same in the background: Begin returns at once; Wait joins and returns the number of failures (the arrays passed to
 * Begin must stay alive until Wait) */
typedef struct b200_parse_job b200_parse_job;
b200_parse_job *h264bsdB200ReparseStreamsBegin(b200_tape **tapes, uint32_t n, const uint8_t *const *streams, const size_t *lens,
                                               uint32_t noOutputReordering, uint32_t threads);
int h264bsdB200ReparseStreamsWait(b200_parse_job *job);
void h264bsdB200FreeTape(b200_tape *tape);

/* ---- GPU (fail loudly -- NULL / -1 and a message on stderr -- when no CUDA device is usable) ---- */

typedef struct b200_batch b200_batch;

int h264bsdB200DeviceCount(void);

/* page-locked host memory for read-backs, and page-locking of a parsed tape for uploads */
void *h264bsdB200HostAlloc(size_t bytes);
void h264bsdB200HostFree(void *p);
int h264bsdB200PinTape(b200_tape *tape);
void h264bsdB200UnpinTape(b200_tape *tape);

/* nStreams independent streams of one geometry on GPU `device`; numSlots frame slots per stream
 * (tape->numSlots = dpbSize+1, what h264bsdInitDpb allocates: h264bsd_dpb.c:1014-1034). */
b200_batch *h264bsdB200BatchCreate(int device, uint32_t nStreams, uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots);
void h264bsdB200BatchDestroy(b200_batch *batch);

/* copy a parsed tape into HBM as stream `stream`'s work-list (host -> device) */
int h264bsdB200BatchUploadTape(b200_batch *batch, uint32_t stream, const b200_tape *tape);
/* give every other stream its own HBM copy of stream `srcStream`'s work-list (device -> device) */
int h264bsdB200BatchReplicateTape(b200_batch *batch, uint32_t srcStream);
/* streaming upload: the work-list of pictures [firstPic, firstPic+numPics) of one stream on a separate copy stream (call
 * with firstPic == 0 first: it sizes the device arrays and takes the picture headers); UploadFence(p) makes the decode of
 * every picture below p wait for everything queued so far, so that the H2D of later pictures overlaps earlier decodes */
int h264bsdB200BatchUploadTapeRange(b200_batch *batch, uint32_t stream, const b200_tape *tape, uint32_t firstPic, uint32_t numPics);
int h264bsdB200BatchUploadFence(b200_batch *batch, uint32_t throughPic);
/* UploadTapeRange for streams 0..nStreams-1 (tapes[s] -> stream s) followed by UploadFence(firstPic + numPics) */
int h264bsdB200BatchUploadTapesRange(b200_batch *batch, const b200_tape *const *tapes, uint32_t nStreams, uint32_t firstPic, uint32_t numPics);

/* Parse + upload in one go (the end-to-end path): `threads` host threads (the pool's size) each parse a stream into their own
 * re-used page-locked tape -- everything h264bsdDecode does up to the pixel path (h264bsd_decoder.c:152-515,
 * h264bsd_slice_data.c:131-220 without the pels) -- and queue the work-list's upload as stream i of `batch` on their own copy
 * stream, so that parsing, H2D copies and (on another batch) the GPU's work overlap and the host holds one tape per thread instead
 * of one per stream.  Begin returns at once; Wait joins and returns the number of streams that failed.  The arrays passed to
 * Begin must stay alive until Wait; `batch` must not be decoding meanwhile (alternate between two batches). */
typedef struct b200_pu_pool b200_pu_pool;
typedef struct b200_pu_job b200_pu_job;
b200_pu_pool *h264bsdB200ParseUploadPoolCreate(uint32_t threads);
void h264bsdB200ParseUploadPoolDestroy(b200_pu_pool *pool);
b200_pu_job *h264bsdB200BatchParseUploadBegin(b200_batch *batch, b200_pu_pool *pool, uint32_t n, const uint8_t *const *streams,
                                              const size_t *lens, uint32_t flags);
int h264bsdB200BatchParseUploadWait(b200_pu_job *job);

/* reconstruct + in-loop filter + border for picture `picIndex` of EVERY stream (asynchronous).
 * Replaces, per macroblock, h264bsdDecodeMacroblock's pixel half (macroblock_layer.c:965-1131) and,
 * per picture, h264bsdFilterPicture (deblocking.c:575-640). */
int h264bsdB200BatchDecodePicture(b200_batch *batch, uint32_t picIndex);
int h264bsdB200BatchRun(b200_batch *batch, uint32_t firstPic, uint32_t numPics);
int h264bsdB200BatchSync(b200_batch *batch);
uint32_t h264bsdB200BatchNumPics(b200_batch *batch);

/* CUDA-event timing on the engine's own stream */
int h264bsdB200BatchTimerStart(b200_batch *batch);
int h264bsdB200BatchTimerStop(b200_batch *batch, float *ms);

/* frame slot -> contiguous I420 of the coded size (widthMbs*heightMbs*384 bytes), as
 * h264bsdNextOutputPicture hands out (decoder.c:599-623); and the reverse (test hook) */
int h264bsdB200BatchReadFrame(b200_batch *batch, uint32_t stream, uint32_t slot, uint8_t *dst);
/* picture `picIndex` of EVERY stream -> dst + s * strideBytes, one packed transfer (asynchronous: call
 * h264bsdB200BatchSync before reading dst; dst should come from h264bsdB200HostAlloc) */
int h264bsdB200BatchReadPictureAll(b200_batch *batch, uint32_t picIndex, uint8_t *dst, size_t strideBytes);
/* the same for a cropping rectangle (h264bsdCroppingParams, decoder.c:887-921; cropW x cropH luma pels at (cropX, cropY), all
 * even; cropW == 0: the coded size) and / or as NV12 (one interleaved chroma plane): cropping, pitch stripping and format
 * conversion happen in the kernel that stages the pictures for the transfer; dst + s * strideBytes receives cropW * cropH * 3 / 2
 * bytes per stream */
int h264bsdB200BatchReadPictureAllEx(b200_batch *batch, uint32_t picIndex, uint8_t *dst, size_t strideBytes, uint32_t cropX,
                                     uint32_t cropY, uint32_t cropW, uint32_t cropH, int nv12);
int h264bsdB200BatchWriteFrame(b200_batch *batch, uint32_t stream, uint32_t slot, const uint8_t *src);
/* h264bsdConvertTo{RGBA(0),BGRA(1),YCbCrA(2)} of a frame slot (decoder.c:1163-1370) into host memory */
int h264bsdB200BatchConvertFrame(b200_batch *batch, uint32_t stream, uint32_t slot, int mode, uint32_t *dst);
int h264bsdB200BatchConvertBench(b200_batch *batch, uint32_t stream, uint32_t slot, int mode, int reps, float *ms);
/* the same for frame `slot` of every stream in one launch (device output only): the YUV->ARGB kernel at batch size */
int h264bsdB200BatchConvertBenchAll(b200_batch *batch, uint32_t slot, int mode, int reps, float *ms);
/* number of streams whose frame in slots[s] differs from stream 0's frame in slots[0]; <0 on error */
int h264bsdB200BatchCompareStreams(b200_batch *batch, const uint32_t *slots);
/* run only some stages of a picture (test hook) */
int h264bsdB200BatchDebugStage(b200_batch *batch, uint32_t picIndex, int recon, int deblock);
/* blocks whose residual left [-512,511] since creation (h264bsd_transform.c:183-188 error return) */
uint32_t h264bsdB200BatchIdctErrors(b200_batch *batch);
/* macroblocks that had at least one non-zero boundary strength (the ones the in-loop filter touches) since creation */
uint64_t h264bsdB200BatchDeblockWorkMbs(b200_batch *batch);
/* per-stage device time: CUDA events around every launch on the engine's stream.  ms6/launches6 = {reconstruct pass A, first
 * instance (copies, one partition, I_PCM), in-loop filter, border, reconstruct pass B (intra), boundary strengths, pass A second
 * instance (several partitions)}; reading resets the accumulators */
void h264bsdB200BatchKernelTiming(b200_batch *batch, int enable);
int h264bsdB200BatchKernelTimes(b200_batch *batch, float *ms6, uint32_t *launches6);
/* waits inside the kernels that gave up (0: macroblock-flag waits, 1: TMA waits); non-zero = engine bug */
uint32_t h264bsdB200BatchWatchdog(b200_batch *batch, int which);
uint64_t h264bsdB200BatchLaunches(b200_batch *batch);
uint64_t h264bsdB200BatchH2DBytes(b200_batch *batch);
uint64_t h264bsdB200BatchD2HBytes(b200_batch *batch);

#ifdef __cplusplus
}
#endif

#endif /* H264BSD_B200_H */
