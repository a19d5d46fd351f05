/*
 * seven.h — the `.7` raw-pixel container (reference 7/seven.h:3-4, 7/libseven.c:3-36).
 * File = two little-endian u32 { (w-1) | 7<<24, (h-1) | A<<24 } followed by w*h*(3+A) raw bytes.
 * Both functions return 0 on success, 1 on failure.  load_7 mallocs pm->p (caller frees).
 */
#ifndef SEVEN_H_B200
#define SEVEN_H_B200
#include "xpng.h"
#ifdef __cplusplus
extern "C" {
#endif
_Bool store_7(const xpng_t *pm, const char *fn);   /* 7/seven.h:3 */
_Bool load_7(const char *fn, xpng_t *pm);          /* 7/seven.h:4 */
#ifdef __cplusplus
}
#endif
#endif
