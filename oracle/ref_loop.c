/* oracle/ref_loop.c -- TEST INFRASTRUCTURE ONLY (CPU baseline of bench.py).
 * The reference decoder's own decode loop (posix/test_h264bsd.c:127-183) run on T host
 * threads, one independent decoder instance per thread, input re-copied before every pass
 * (decode mutates it), no file output.  Prints one JSON line.
 *   usage: ref_loop in.h264 threads seconds_min                                          */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <pthread.h>
#include "h264bsd_decoder.h"
#include "h264bsd_util.h"

static uint8_t *g_stream; static size_t g_len; static double g_min_s;
typedef struct { uint64_t mbs; uint64_t pics; double secs; int err; } res_t;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *worker(void *arg)
{
    res_t *r = (res_t *)arg;
    uint8_t *buf = malloc(g_len);
    storage_t *dec = h264bsdAlloc();
    double t0 = now();
    do {
        memcpy(buf, g_stream, g_len);
        if (h264bsdInit(dec, HANTRO_FALSE) != HANTRO_OK) { r->err = 1; break; }
        uint8_t *p = buf; uint32_t left = (uint32_t)g_len, rb = 0;
        while (left > 0) {
            uint32_t st = h264bsdDecode(dec, p, left, 0, &rb);
            p += rb; left -= rb;
            if (st == H264BSD_PIC_RDY) {
                uint32_t a, b, c;
                if (h264bsdNextOutputPicture(dec, &a, &b, &c)) { r->pics++; r->mbs += dec->picSizeInMbs; }
            } else if (st == H264BSD_ERROR || st == H264BSD_PARAM_SET_ERROR) { r->err = 1; break; }
        }
        h264bsdShutdown(dec);
    } while (!r->err && now() - t0 < g_min_s);
    r->secs = now() - t0;
    h264bsdFree(dec); free(buf);
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s in.h264 threads seconds\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END); g_len = ftell(f); fseek(f, 0, SEEK_SET);
    g_stream = malloc(g_len); if (fread(g_stream, 1, g_len, f) != g_len) return 1; fclose(f);
    int T = atoi(argv[2]); g_min_s = atof(argv[3]);
    pthread_t *th = malloc(sizeof(pthread_t) * T); res_t *rs = calloc(T, sizeof(res_t));
    double t0 = now();
    for (int i = 0; i < T; i++) pthread_create(&th[i], NULL, worker, &rs[i]);
    for (int i = 0; i < T; i++) pthread_join(th[i], NULL);
    double wall = now() - t0;
    uint64_t mbs = 0, pics = 0; int err = 0;
    for (int i = 0; i < T; i++) { mbs += rs[i].mbs; pics += rs[i].pics; err |= rs[i].err; }
    printf("{\"threads\": %d, \"wall_s\": %.4f, \"mbs\": %llu, \"pics\": %llu, \"mb_per_s\": %.1f, \"err\": %d}\n",
           T, wall, (unsigned long long)mbs, (unsigned long long)pics, mbs / wall, err);
    return err;
}
