/*
 * png7.h — PNG <-> `.7` conversion of the exchange scheme (reference tool 7/seven.c:39-79, built on libpng's
 * simplified API there; here zlib only).  Host-side C, no GPU work.  All functions return 0 on success, 1 on failure.
 *   png_load   : 7/seven.c:44-59 (png_image_begin_read_from_file ... png_image_finish_read): 8-bit RGB, or RGBA when the
 *                file has alpha / tRNS; 16-bit ("linear") files are refused like 7/seven.c:48; pm->p is malloc'd.
 *   png_store  : 7/seven.c:63-71 (png_image_write_to_file).
 *   ppm_load / ppm_store : binary PPM (P6, maxval 255), the pixmap format of the reference's older front end
 *                (ancestor/gray.c:667-682) and of the sintel frames; RGB only.
 *   seven_main : the `seven --to_7 | --to_png` command line including normalize_RGBA (7/seven.c:4-37) and the usage text.
 */
#ifndef PNG7_H_B200
#define PNG7_H_B200
#include "seven.h"
#ifdef __cplusplus
extern "C" {
#endif
_Bool png_load(const char *fn, xpng_t *pm);
_Bool png_store(const xpng_t *pm, const char *fn);
_Bool ppm_load(const char *fn, xpng_t *pm);
_Bool ppm_store(const xpng_t *pm, const char *fn);
int seven_main(int argc, char **argv);
#ifdef __cplusplus
}
#endif
#endif
