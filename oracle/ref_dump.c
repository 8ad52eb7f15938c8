/* oracle/ref_dump.c -- TEST INFRASTRUCTURE ONLY.
 * CLI around ref_decode_stream() (ref_shim.c): writes the reference decoder's frames at
 * full coded size.   usage: ref_dump in.h264 post.yuv [pre.yuv]                        */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

int ref_decode_stream(const uint8_t *, size_t, uint8_t *, size_t, uint8_t *, size_t,
                      void *, size_t, uint32_t *);

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.h264 post.yuv [pre.yuv]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *s = malloc(n);
    if (fread(s, 1, n, f) != (size_t)n) return 1;
    fclose(f);
    size_t cap = (size_t)512 << 20;
    uint8_t *post = malloc(cap), *pre = argc > 3 ? malloc(cap) : NULL;
    uint32_t info[8] = {0};
    int np = ref_decode_stream(s, n, post, cap, pre, cap, NULL, 0, info);
    if (np < 0) { fprintf(stderr, "decode error\n"); return 1; }
    size_t bytes = (size_t)np * info[0] * info[1] * 384;
    f = fopen(argv[2], "wb"); fwrite(post, 1, bytes, f); fclose(f);
    if (pre) { f = fopen(argv[3], "wb"); fwrite(pre, 1, (size_t)info[7] * info[0] * info[1] * 384, f); fclose(f); }
    printf("%d pictures, %ux%u MBs, crop=%u (%u,%u,%u,%u)\n", np, info[0], info[1], info[2],
           info[3], info[4], info[5], info[6]);
    return 0;
}
