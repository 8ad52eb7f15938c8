/* h264bsd_util.h -- the status macros front ends use with the preserved API
 * (posix/test_h264bsd.c:129-134 tests h264bsdInit() against HANTRO_OK and passes HANTRO_FALSE;
 * values as in the reference's src/h264bsd_util.h:54-59). */
#ifndef H264BSD_B200_UTIL_H
#define H264BSD_B200_UTIL_H
#define HANTRO_OK 0
#define HANTRO_NOK 1
#define HANTRO_TRUE 1
#define HANTRO_FALSE 0
#endif
