/* h264bsd_storage.h -- the decoder instance object of the preserved API.
 *
 * Callers allocate storage_t themselves (on the stack in posix/test_h264bsd.c:129, through
 * h264bsdAlloc in the wasm / iOS / Windows shells), so its SIZE and ALIGNMENT are ABI: 4648 bytes,
 * 8-byte aligned on LP64, exactly the reference's struct (src/h264bsd_storage.h:75-152).  The B200
 * engine keeps its state behind one pointer; the rest of the block is reserved and zeroed by
 * h264bsdInit.  Front ends only ever pass the address around. */
#ifndef H264BSD_B200_STORAGE_H
#define H264BSD_B200_STORAGE_H
#include "basetype.h"

#define H264BSD_STORAGE_BYTES 4648

typedef struct storage {
    union {
        struct {
            void *engine;  /* b200 decoder instance, owned by the library */
            u32 magic;
        } b200;
        unsigned long long align_[H264BSD_STORAGE_BYTES / 8];
        unsigned char bytes_[H264BSD_STORAGE_BYTES];
    } u;
} storage_t;

#endif
