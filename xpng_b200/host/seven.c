/* seven.c — `.7` container I/O, restating reference 7/libseven.c:3-36 (host-side C, no GPU work). */
#include "seven.h"
#include <stdio.h>
#include <stdlib.h>
#include <sys/stat.h>

#define DIM_MAX (1u << 24)

_Bool store_7(const xpng_t *pm, const char *fn) {
    /* validation as 7/libseven.c:5-8 */
    if (!pm || !pm->p || !fn) return 1;
    if (pm->w == 0 || pm->w > DIM_MAX || pm->h == 0 || pm->h > DIM_MAX) return 1;
    if (pm->s != pm->w * pm->h * (3 + (u64_t)pm->A)) return 1;
    const u32_t hdr[2] = { (u32_t)(pm->w - 1) | (7u << 24), (u32_t)(pm->h - 1) | ((u32_t)pm->A << 24) };
    FILE *f = fopen(fn, "wb");
    if (!f) return 1;
    _Bool bad = fwrite(hdr, 1, 8, f) != 8 || fwrite(pm->p, 1, pm->s, f) != pm->s;
    bad |= fclose(f) != 0;
    return bad;
}

_Bool load_7(const char *fn, xpng_t *pm) {
    struct stat st;
    if (!fn || !pm || stat(fn, &st) != 0) return 1;
    const u64_t fsize = (u64_t)st.st_size;
    if (fsize < 11) return 1;                                   /* 7/libseven.c:21 */
    FILE *f = fopen(fn, "rb");
    u32_t hdr[2];
    if (!f) return 1;
    if (fread(hdr, 1, 8, f) != 8) { fclose(f); return 1; }
    pm->w = (hdr[0] & 0xFFFFFFu) + 1; pm->h = (hdr[1] & 0xFFFFFFu) + 1; pm->A = (hdr[1] >> 24) & 1;
    pm->s = pm->w * pm->h * (3 + (u64_t)pm->A); pm->p = NULL;
    if (pm->s + 8 != fsize || (hdr[0] >> 24) != 7) { fclose(f); return 1; }   /* 7/libseven.c:30 */
    pm->p = malloc(pm->s);
    if (!pm->p) { fclose(f); return 1; }
    if (fread(pm->p, 1, pm->s, f) != pm->s) { fclose(f); free(pm->p); pm->p = NULL; return 1; }
    return fclose(f) != 0;
}
