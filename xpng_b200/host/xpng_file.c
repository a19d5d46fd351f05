/*
 * xpng_file.c — the reference's file-at-a-time API (xpng.h:12-20) implemented in C on top of the
 * CUDA C ABI (xpng_b200.h).  Mirrors xpng_store_T (libxpng.c:723-789) and xpng_load_T
 * (libxpng.c:963-997): same validation, same return convention, same one-line stdout report, same
 * ownership (xpng_load mallocs pm->p).  All pixel work happens on the GPU; with no CUDA device the
 * calls fail (return 1) — there is deliberately no CPU path.
 */
#include "xpng.h"
#include "xpng_b200.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <unistd.h>

/* The reference API carries no handle and is re-entrant (no global mutable state, SURVEY 8(b) "Threading"): concurrent
 * callers each take a codec context from a small pool (created on demand, returned after the call), so calls from
 * different host threads run side by side on the GPU instead of queueing behind one lock. */
#define POOL_MAX 16
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_free = PTHREAD_COND_INITIALIZER;
static xpngb_ctx *g_idle[POOL_MAX];
static int g_nidle = 0, g_ncreated = 0;

static xpngb_ctx *ctx_get(void) {
    xpngb_ctx *ctx = NULL;
    pthread_mutex_lock(&g_lock);
    for (;;) {
        if (g_nidle) { ctx = g_idle[--g_nidle]; break; }
        if (g_ncreated < POOL_MAX) { g_ncreated++; break; }        /* create outside the lock */
        pthread_cond_wait(&g_free, &g_lock);
    }
    pthread_mutex_unlock(&g_lock);
    if (ctx) return ctx;
    const char *d = getenv("XPNG_DEVICE");
    if (xpngb_create(&ctx, d ? atoi(d) : 0)) {
        fprintf(stderr, "xpng: no usable CUDA device (this build has no CPU path)\n");
        pthread_mutex_lock(&g_lock); g_ncreated--; pthread_cond_signal(&g_free); pthread_mutex_unlock(&g_lock);
        return NULL;
    }
    return ctx;
}
static void ctx_put(xpngb_ctx *ctx) {
    if (!ctx) return;
    pthread_mutex_lock(&g_lock);
    g_idle[g_nidle++] = ctx;
    pthread_cond_signal(&g_free);
    pthread_mutex_unlock(&g_lock);
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_REALTIME, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void report(const char *what, u64_t T, double secs, u64_t npx) {
    /* line shape of libxpng.c:761 / :986 */
    const int t = T ? (int)T : (int)sysconf(_SC_NPROCESSORS_ONLN);   /* the reference resolves T = 0 to the online cores (libxpng.c:147) and prints that */
    printf("%s, %3d thread%c: %5lu MPx/s\n", what, t, t > 1 ? 's' : ' ', (unsigned long)((npx / 1e6) / (secs > 0 ? secs : 1e-9)));
}

_Bool xpng_store_T(u64_t T, u64_t mode, const xpng_t *pm, const char *fn) {
    if (!pm || !fn) return 1;
    /* libxpng.c:729-731 */
    if (pm->w > (1u << 24) || !pm->w || !(mode == 1 || mode == 2 || mode == 7) || pm->h > (1u << 24) || !pm->h ||
        pm->w * pm->h * (3 + (u64_t)pm->A) != pm->s || pm->p == NULL) return 1;
    _Bool bad = 1;
    uint8_t *out = NULL;
    xpngb_ctx *ctx = ctx_get();
    /* the reference's clock covers the encode work up to the joined threads, not the file write (libxpng.c:727, :760);
     * here it covers the whole codec call (transfers included) but not the one-time CUDA device initialisation */
    const double t0 = now_s();
    if (ctx) {
        xpngb_image im = { pm->w, pm->h, 0, pm->A ? 1u : 0u, 0 };
        const uint64_t cap = xpngb_encode_bound(&im, 1);
        uint64_t off = 0, size = 0;
        out = malloc(cap);
        if (out && !xpngb_encode(ctx, (int)mode, &im, 1, pm->p, pm->s, 0, out, cap, 0, &off, &size)) {
            const double secs = now_s() - t0;
            FILE *f = fopen(fn, "wb");
            if (f) {
                bad = fwrite(out + off, 1, size, f) != size;
                bad |= fclose(f) != 0;
                /* the reference prints only when tiles were coded (libxpng.c:738, :751 return earlier) */
                const _Bool single = size == 11 + (u64_t)im.A && (out[off + 7] & 2);
                if (!bad && im.mode != 7 && !single) report("encode", T, secs, pm->w * pm->h);
                else if (!bad && im.mode == 7 && mode != 7 && pm->s > 4) report("encode", T, secs, pm->w * pm->h);
            }
        } else if (out) fprintf(stderr, "xpng: %s\n", xpngb_last_error(ctx));
    }
    ctx_put(ctx);
    free(out);
    return bad;
}

_Bool xpng_store(u64_t mode, const xpng_t *pm, const char *fn) { return xpng_store_T(0, mode, pm, fn); }   /* libxpng.c:791-794 */

_Bool xpng_load_T(u64_t T, const char *fn, xpng_t *pm) {
    struct stat st;
    if (!fn || !pm || stat(fn, &st) != 0) return 1;
    const uint64_t fsize = (uint64_t)st.st_size;
    FILE *f = fopen(fn, "rb");
    if (!f) return 1;
    uint8_t *file = malloc(fsize + 16);
    if (!file || fread(file, 1, fsize, f) != fsize) { fclose(f); free(file); return 1; }
    fclose(f);
    xpngb_image im;
    memset(&im, 0, sizeof im);
    if (xpngb_peek(file, fsize, &im)) { free(file); return 1; }   /* libxpng.c:969-972 */
    pm->w = im.w; pm->h = im.h; pm->A = im.A != 0; pm->s = im.w * im.h * (3 + (u64_t)im.A);
    pm->p = malloc(pm->s + 16);
    if (!pm->p) { free(file); return 1; }
    _Bool bad = 1;
    xpngb_ctx *ctx = ctx_get();
    const double t0 = now_s();   /* the reference starts its clock after f_read (libxpng.c:967); device initialisation is not codec work */
    if (ctx) {
        const uint64_t off = 0;
        im.offset = 0;
        bad = xpngb_decode(ctx, &im, 1, file, fsize, 0, &off, &fsize, pm->p, pm->s, 0) != 0;
        if (bad) fprintf(stderr, "xpng: %s\n", xpngb_last_error(ctx));
    }
    ctx_put(ctx);
    const _Bool single = fsize == 11 + (u64_t)im.A && (file[7] & 2);
    if (!bad && im.mode != 7 && !single) report("decode", T, now_s() - t0, pm->w * pm->h);
    free(file);
    if (bad) { free(pm->p); pm->p = NULL; }
    return bad;
}

_Bool xpng_load(const char *fn, xpng_t *pm) { return xpng_load_T(0, fn, pm); }   /* libxpng.c:999-1002 */

/* libxpng.c:1004-1014: a stub in the reference as well */
_Bool xpng_from_jpg_T(u64_t T, const char *jpg, const char *xpng) {
    (void)T; (void)jpg; (void)xpng;
    puts("\nNot Implemented.\n");
    return 1;
}
_Bool xpng_from_jpg(const char *jpg, const char *xpng) { return xpng_from_jpg_T(0, jpg, xpng); }
