/*
 * xpng.h — drop-in replacement of the reference's public header (reference xpng.h:1-20).
 *
 * Same constants, same xpng_t layout, same six entry points with the same meaning and the same
 * return convention (_Bool: 0 = success, 1 = failure).  The reference header pulls its typedefs
 * from "until_fork/until_fork.h" (until_fork.h:30-35, :45); the compatible subset is declared here
 * so that callers written against the reference compile unchanged.
 *
 * Implementation: xpng_b200/host/xpng_file.c (plain C) on top of the CUDA C ABI in xpng_b200.h.
 * The `T` (thread count) argument of the *_T variants is accepted and ignored: the work is spread
 * over the GPU's SMs, not host threads.  There is no CPU fallback; without a CUDA device every
 * call returns 1.
 */
#ifndef XPNG_H_B200
#define XPNG_H_B200
#include <stdbool.h>
#include <stdint.h>

#ifndef XPNG_UNTIL_FORK_TYPES
#define XPNG_UNTIL_FORK_TYPES
typedef int8_t s7_t;   typedef uint8_t u8_t;     /* until_fork.h:30 */
typedef int16_t s15_t; typedef uint16_t u16_t;   /* until_fork.h:31 */
typedef int32_t s31_t; typedef uint32_t u32_t;   /* until_fork.h:32 */
typedef int64_t s63_t; typedef uint64_t u64_t;   /* until_fork.h:33 */
#endif
#ifndef CHECK
#define CHECK __attribute__((warn_unused_result))   /* until_fork.h:37 */
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define XPNG_COMPRESSION_TYPE_FAST 1           /* xpng.h:5 */
#define XPNG_COMPRESSION_TYPE_SLOW 2           /* xpng.h:6 */
#define XPNG_COMPRESSION_TYPE_EXJPEG 3         /* xpng.h:7 */
#define XPNG_COMPRESSION_TYPE_UNCOMPRESSED 7   /* xpng.h:8 */

/* Interleaved 8-bit RGB (A = 0) or RGBA (A = 1), row-major, s = w*h*(3+A) bytes.  xpng.h:10 */
typedef struct xpng_t { u8_t *p; u64_t w, h, s; _Bool A; } xpng_t;

CHECK _Bool xpng_store(u64_t mode, const xpng_t *pm, const char *xpng);            /* xpng.h:12 */
CHECK _Bool xpng_load(const char *xpng, xpng_t *pm);                               /* xpng.h:13 */
CHECK _Bool xpng_from_jpg(const char *jpg, const char *xpng);                      /* xpng.h:15 */
CHECK _Bool xpng_store_T(u64_t T, u64_t mode, const xpng_t *pm, const char *xpng); /* xpng.h:17 */
CHECK _Bool xpng_load_T(u64_t T, const char *xpng, xpng_t *pm);                    /* xpng.h:18 */
CHECK _Bool xpng_from_jpg_T(u64_t T, const char *jpg, const char *xpng);           /* xpng.h:20 */

#ifdef __cplusplus
}
#endif
#endif
