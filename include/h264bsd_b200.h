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
 * tape->status != 0 if the parse stopped on a decoder error.  NULL only on allocation failure. */
b200_tape *h264bsdB200ParseStream(const uint8_t *stream, size_t len, uint32_t noOutputReordering);
void h264bsdB200FreeTape(b200_tape *tape);

#ifdef __cplusplus
}
#endif

#endif /* H264BSD_B200_H */
