/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin tracing shim linked into oracle/_ref/libh264bsd_ref.so together with the
 * UNMODIFIED reference objects.  It adds no arithmetic of its own:
 *
 *   refShimFilterPicture()  stands in for the call at h264bsd_decoder.c:475; it snapshots
 *                           the picture before the in-loop filter (the "pre-deblock tap")
 *                           and the per-MB state the filter reads, then calls the real
 *                           h264bsdFilterPicture (h264bsd_deblocking.c:575).
 *   ref_decode_stream()     the decode loop of posix/test_h264bsd.c:127-183 with frames
 *                           written at FULL coded size (picSizeInMbs*384 B), not the
 *                           cropped/truncated savePic format.
 *   ref_convert()           h264bsdConvertToRGBA/BGRA/YCbCrA (h264bsd_decoder.c:1163-1370).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "h264bsd_decoder.h"
#include "h264bsd_util.h"
#include "h264bsd_deblocking.h"

/* one record per macroblock, written at every pre-deblock tap (decode order) */
typedef struct {
    int32_t mbType;
    int32_t qpY;
    int32_t sliceId;
    int32_t disableDeblockingFilterIdc;
    int32_t filterOffsetA, filterOffsetB;
    int32_t chromaQpIndexOffset;
    int16_t totalCoeff[27];
    uint8_t intra4x4PredMode[16];
    int32_t refPic[4];
    int32_t refSlot[4]; /* refAddr resolved to an index in the tap's frame-pointer table */
    int16_t mv[16][2];
} ref_mb_tap_t;

static uint8_t *g_pre;          /* pre-deblock frames, decode order */
static size_t g_pre_cap, g_pre_len;
static ref_mb_tap_t *g_mbtap;   /* per-MB taps, decode order */
static size_t g_mbtap_cap, g_mbtap_len;
static uint32_t g_tap_pics;
static int g_resilient;            /* carry on after H264BSD_ERROR (what a player does; the posix test program exits) */
static uint32_t g_err_mbs[4096];   /* numErrMbs of every output picture of the last ref_decode_stream call */
static uint32_t g_err_count;

static uint32_t g_video_info[4];   /* h264bsdVideoRange, h264bsdMatrixCoefficients, sample aspect ratio w, h at the end of the last decode */

static int g_no_reordering;        /* h264bsdInit's noOutputReordering argument for the next ref_decode_stream */

void ref_set_resilient(int on) { g_resilient = on; }
void ref_set_no_reordering(int on) { g_no_reordering = on; }
void ref_video_info(uint32_t out[4]) { memcpy(out, g_video_info, sizeof g_video_info); }
uint32_t ref_err_mbs(uint32_t *dst, uint32_t cap)
{
    uint32_t n = g_err_count < cap ? g_err_count : cap;
    memcpy(dst, g_err_mbs, n * sizeof(uint32_t));
    return g_err_count;
}

void refShimFilterPicture(image_t *image, mbStorage_t *pMb)
{
    size_t nmb = (size_t)image->width * image->height;
    size_t bytes = nmb * 384;
    if (g_pre && g_pre_len + bytes <= g_pre_cap) {
        memcpy(g_pre + g_pre_len, image->data, bytes);
        g_pre_len += bytes;
    }
    if (g_mbtap && g_mbtap_len + nmb <= g_mbtap_cap) {
        for (size_t i = 0; i < nmb; i++) {
            ref_mb_tap_t *t = g_mbtap + g_mbtap_len + i;
            const mbStorage_t *m = pMb + i;
            t->mbType = (int32_t)m->mbType;
            t->qpY = (int32_t)m->qpY;
            t->sliceId = (int32_t)m->sliceId;
            t->disableDeblockingFilterIdc = (int32_t)m->disableDeblockingFilterIdc;
            t->filterOffsetA = m->filterOffsetA;
            t->filterOffsetB = m->filterOffsetB;
            t->chromaQpIndexOffset = m->chromaQpIndexOffset;
            memcpy(t->totalCoeff, m->totalCoeff, sizeof t->totalCoeff);
            memcpy(t->intra4x4PredMode, m->intra4x4PredMode, 16);
            for (int k = 0; k < 4; k++) {
                t->refPic[k] = (int32_t)m->refPic[k];
                t->refSlot[k] = -1;
            }
            for (int k = 0; k < 16; k++) {
                t->mv[k][0] = m->mv[k].hor;
                t->mv[k][1] = m->mv[k].ver;
            }
        }
        g_mbtap_len += nmb;
    }
    g_tap_pics++;
    h264bsdFilterPicture(image, pMb);
}

/* Decode a whole Annex-B stream.  post: output-order frames after the in-loop filter;
 * pre: decode-order frames before it; mbtap: decode-order per-MB state.  Any of the three
 * may be NULL.  Returns number of output pictures, or -1 on decoder error. */
int ref_decode_stream(const uint8_t *stream, size_t len,
                      uint8_t *post, size_t post_cap,
                      uint8_t *pre, size_t pre_cap,
                      void *mbtap, size_t mbtap_cap_records,
                      uint32_t *info /* [0]=widthMbs [1]=heightMbs [2]=cropFlag [3..6]=crop l,w,t,h [7]=tap pics */)
{
    storage_t *dec = h264bsdAlloc();
    uint8_t *buf = (uint8_t *)malloc(len ? len : 1);
    if (!dec || !buf) return -1;
    memcpy(buf, stream, len);  /* decode mutates its input: byte_stream.c:193-233 */
    if (h264bsdInit(dec, g_no_reordering ? HANTRO_TRUE : HANTRO_FALSE) != HANTRO_OK) return -1;

    g_pre = pre; g_pre_cap = pre_cap; g_pre_len = 0;
    g_mbtap = (ref_mb_tap_t *)mbtap; g_mbtap_cap = mbtap_cap_records; g_mbtap_len = 0;
    g_tap_pics = 0;
    g_err_count = 0;

    uint8_t *p = buf;
    uint32_t left = (uint32_t)len, rb = 0;
    size_t post_len = 0;
    int npics = 0, err = 0;
    while (left > 0) {
        uint32_t r = h264bsdDecode(dec, p, left, 0, &rb);
        p += rb; left -= rb;
        if (r == H264BSD_PIC_RDY) {
            uint32_t picId, isIdr, nErr;
            uint8_t *pic;
            while ((pic = h264bsdNextOutputPicture(dec, &picId, &isIdr, &nErr)) != NULL) {
                size_t bytes = (size_t)dec->picSizeInMbs * 384;
                if (post && post_len + bytes <= post_cap) memcpy(post + post_len, pic, bytes);
                post_len += bytes;
                if (g_err_count < 4096) g_err_mbs[g_err_count++] = nErr;
                npics++;
            }
        } else if (r == H264BSD_HDRS_RDY) {
            if (info) {
                info[0] = h264bsdPicWidth(dec);
                info[1] = h264bsdPicHeight(dec);
                h264bsdCroppingParams(dec, &info[2], &info[3], &info[4], &info[5], &info[6]);
            }
        } else if (r == H264BSD_ERROR && g_resilient && rb) {
            continue;
        } else if (r == H264BSD_ERROR || r == H264BSD_PARAM_SET_ERROR) {
            err = 1;
            break;
        }
    }
    if (!err) {
        /* drain pictures still held for reordering (h264bsd_decoder.c: h264bsdFlushBuffer) */
        uint32_t picId, isIdr, nErr;
        uint8_t *pic;
        h264bsdFlushBuffer(dec);
        while ((pic = h264bsdNextOutputPicture(dec, &picId, &isIdr, &nErr)) != NULL) {
            size_t bytes = (size_t)dec->picSizeInMbs * 384;
            if (post && post_len + bytes <= post_cap) memcpy(post + post_len, pic, bytes);
            post_len += bytes;
            if (g_err_count < 4096) g_err_mbs[g_err_count++] = nErr;
            npics++;
        }
    }
    if (info) info[7] = g_tap_pics;
    memset(g_video_info, 0, sizeof g_video_info);
    if (dec->activeSps) {
        g_video_info[0] = h264bsdVideoRange(dec);
        g_video_info[1] = h264bsdMatrixCoefficients(dec);
        h264bsdSampleAspectRatio(dec, &g_video_info[2], &g_video_info[3]);
    }
    g_pre = NULL; g_mbtap = NULL;
    h264bsdShutdown(dec);
    h264bsdFree(dec);
    free(buf);
    return err ? -1 : npics;
}

/* mode 0 RGBA, 1 BGRA, 2 YCbCrA; width/height in pixels (coded size) */
void ref_convert(int mode, uint32_t width, uint32_t height, uint8_t *yuv, uint32_t *out)
{
    if (mode == 0) h264bsdConvertToRGBA(width, height, yuv, out);
    else if (mode == 1) h264bsdConvertToBGRA(width, height, yuv, out);
    else h264bsdConvertToYCbCrA(width, height, yuv, out);
}

size_t ref_sizeof_storage(void) { return sizeof(storage_t); }
size_t ref_sizeof_mbtap(void) { return sizeof(ref_mb_tap_t); }
