/* decode_file.c -- command-line decoder over the preserved single-stream API.
 *
 * Plays the role of the reference's posix/test_h264bsd.c (its decode loop at lines 130-179): read an Annex-B
 * file, feed it to h264bsdDecode, collect pictures with h264bsdNextOutputPicture, optionally write raw I420 or
 * compare with a .yuv file.  It includes only the public headers in include/ and links against
 * libh264bsd_b200.so, so it is also the link-compatibility check for the drop-in boundary.
 *
 *   cc -Iinclude examples/decode_file.c -Lh264bsd_b200 -lh264bsd_b200 -Wl,-rpath,$PWD/h264bsd_b200 -o decode_file
 *   ./decode_file [-o out.yuv] [-c expected.yuv] [-r repeat] in.h264
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "h264bsd_decoder.h"
#include "h264bsd_util.h"

static u8 *slurp(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    u8 *buf = (u8 *)malloc((size_t)n + 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { perror("read"); exit(2); }
    fclose(f);
    *len = (size_t)n;
    return buf;
}

int main(int argc, char **argv)
{
    const char *outPath = NULL, *cmpPath = NULL, *inPath = NULL;
    int repeat = 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-o") && i + 1 < argc) outPath = argv[++i];
        else if (!strcmp(argv[i], "-c") && i + 1 < argc) cmpPath = argv[++i];
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) repeat = atoi(argv[++i]);
        else inPath = argv[i];
    }
    if (!inPath) { fprintf(stderr, "usage: %s [-o out.yuv] [-c expected.yuv] [-r n] in.h264\n", argv[0]); return 2; }

    size_t streamLen = 0, cmpLen = 0, cmpPos = 0;
    u8 *stream = slurp(inPath, &streamLen);
    u8 *expected = cmpPath ? slurp(cmpPath, &cmpLen) : NULL;
    FILE *out = outPath ? fopen(outPath, "wb") : NULL;
    u32 pics = 0, mismatches = 0;

    for (int rep = 0; rep < repeat; rep++) {
        storage_t *dec = h264bsdAlloc();
        if (!dec || h264bsdInit(dec, HANTRO_FALSE) != HANTRO_OK) { fprintf(stderr, "h264bsdInit failed\n"); return 1; }
        u8 *p = stream;
        u32 len = (u32)streamLen;
        u32 frameBytes = 0;
        while (len > 0) {
            u32 readBytes = 0;
            u32 res = h264bsdDecode(dec, p, len, 0, &readBytes);
            p += readBytes;
            len -= readBytes;
            if (res == H264BSD_HDRS_RDY) {
                frameBytes = h264bsdPicWidth(dec) * h264bsdPicHeight(dec) * 384;
            } else if (res == H264BSD_PIC_RDY) {
                u32 picId, isIdr, numErr;
                u8 *pic;
                while ((pic = h264bsdNextOutputPicture(dec, &picId, &isIdr, &numErr)) != NULL) {
                    pics++;
                    if (out && rep == 0) fwrite(pic, 1, frameBytes, out);
                    if (expected && rep == 0) {
                        if (cmpPos + frameBytes > cmpLen || memcmp(pic, expected + cmpPos, frameBytes)) mismatches++;
                        cmpPos += frameBytes;
                    }
                }
            } else if (res == H264BSD_ERROR || res == H264BSD_PARAM_SET_ERROR) {
                fprintf(stderr, "decode error %u at offset %zu\n", res, (size_t)(p - stream));
            } else if (res == H264BSD_MEMALLOC_ERROR) {
                fprintf(stderr, "out of memory (or no CUDA device)\n");
                return 1;
            }
        }
        h264bsdShutdown(dec);
        h264bsdFree(dec);
    }
    if (out) fclose(out);
    printf("%u pictures decoded", pics);
    if (expected) printf(", %u mismatching%s", mismatches, cmpPos != cmpLen ? " (length differs)" : "");
    printf("\n");
    free(stream);
    free(expected);
    return (mismatches || (expected && cmpPos != cmpLen)) ? 1 : 0;
}
