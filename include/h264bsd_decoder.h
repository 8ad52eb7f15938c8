/* h264bsd_decoder.h -- the single-stream C API of oneam/h264bsd, preserved: same symbols, same
 * signatures, same return codes (reference: src/h264bsd_decoder.h:45-93; export lists
 * wasm/Rakefile:11-25, win/h264bsd.def).  Behind it CAVLC/NAL parsing runs on the host and every
 * pel is produced by the B200 engine; there is no CPU pixel path.  Each prototype names the
 * reference definition it replaces (file:line under src/). */
#ifndef H264BSD_B200_DECODER_H
#define H264BSD_B200_DECODER_H

#include "basetype.h"
#include "h264bsd_storage.h"

#ifdef __cplusplus
extern "C" {
#endif

/* h264bsd_decoder.h:45-52 */
enum {
    H264BSD_RDY,
    H264BSD_PIC_RDY,
    H264BSD_HDRS_RDY,
    H264BSD_ERROR,
    H264BSD_PARAM_SET_ERROR,
    H264BSD_MEMALLOC_ERROR
};

u32 h264bsdInit(storage_t *pStorage, u32 noOutputReordering);                              /* decoder.c:90   */
u32 h264bsdDecode(storage_t *pStorage, u8 *byteStrm, u32 len, u32 picId, u32 *readBytes);  /* decoder.c:152  */
void h264bsdShutdown(storage_t *pStorage);                                                 /* decoder.c:534  */

u8 *h264bsdNextOutputPicture(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs);        /* decoder.c:599 */
u32 *h264bsdNextOutputPictureRGBA(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs);   /* decoder.c:648 */
u32 *h264bsdNextOutputPictureBGRA(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs);   /* decoder.c:690 */
u32 *h264bsdNextOutputPictureYCbCrA(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs); /* decoder.c:732 */

u32 h264bsdPicWidth(storage_t *pStorage);            /* decoder.c:771 */
u32 h264bsdPicHeight(storage_t *pStorage);           /* decoder.c:797 */
u32 h264bsdVideoRange(storage_t *pStorage);          /* decoder.c:875 */
u32 h264bsdMatrixCoefficients(storage_t *pStorage);  /* decoder.c:906 */
void h264bsdCroppingParams(storage_t *pStorage, u32 *croppingFlag, u32 *left, u32 *width, u32 *top, u32 *height); /* decoder.c:941 */
void h264bsdSampleAspectRatio(storage_t *pStorage, u32 *sarWidth, u32 *sarHeight);                                /* decoder.c:993 */
u32 h264bsdCheckValidParamSets(storage_t *pStorage); /* decoder.c:859 */
void h264bsdFlushBuffer(storage_t *pStorage);        /* decoder.c:834 */
u32 h264bsdProfile(storage_t *pStorage);             /* decoder.c:1073 */

storage_t *h264bsdAlloc(void);            /* decoder.c:1110 */
void h264bsdFree(storage_t *pStorage);    /* decoder.c:1133 */

void h264bsdConvertToRGBA(u32 width, u32 height, u8 *data, u32 *pOutput);    /* decoder.c:1163 */
void h264bsdConvertToBGRA(u32 width, u32 height, u8 *data, u32 *pOutput);    /* decoder.c:1244 */
void h264bsdConvertToYCbCrA(u32 width, u32 height, u8 *data, u32 *pOutput);  /* decoder.c:1324 */

#ifdef __cplusplus
}
#endif

#endif
