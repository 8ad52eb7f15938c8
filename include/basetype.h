/* basetype.h -- scalar type names used by the preserved h264bsd C API (cf. the reference's
 * src/basetype.h:29-34).  Own header of the B200 engine; only the names callers rely on. */
#ifndef H264BSD_B200_BASETYPE_H
#define H264BSD_B200_BASETYPE_H
#include <stdint.h>
#include <stddef.h>
typedef uint8_t u8;
typedef int8_t i8;
typedef uint16_t u16;
typedef int16_t i16;
typedef uint32_t u32;
typedef int32_t i32;
#endif
