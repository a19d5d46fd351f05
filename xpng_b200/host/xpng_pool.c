/*
 * xpng_pool.c — frame batches sharded over the GPUs of one box, in C, without NCCL.
 *
 * The reference's only parallel strategy is a tile cursor plus an ordered concatenation of the results
 * (libxpng.c:146-151 fan-out, :764-769 ordered fwrite, :982 offset chain on the way back).  The B200 analogue:
 * frames are independent, so a batch is cut into contiguous shards (frame i -> shard floor(i * S / n)), every shard
 * is coded on its own device with no data-path exchange, and the ONLY exchange is the table of compressed sizes,
 * from which every participant derives the same offset table by an exclusive scan.
 *
 * Two forms of the same thing:
 *   xpngb_pool_*    one process, one host thread and one codec context per device (host buffers in, host buffers out)
 *   xpngb_gather_*  one process per device (torchrun-style launch): the size tables meet in a POSIX shared-memory
 *                   segment; no sockets, no NCCL, no Python.
 */
#define _GNU_SOURCE
#include "xpng_b200.h"
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

void xpngb_shard_range(uint32_t n, uint32_t nshards, uint32_t shard, uint32_t *first, uint32_t *count) {
    /* frame i belongs to shard floor(i * nshards / n): shard k owns [ceil(k n / S), ceil((k + 1) n / S)) */
    const uint64_t S = nshards ? nshards : 1;
    const uint64_t lo = ((uint64_t)shard * n + S - 1) / S, hi = ((uint64_t)(shard + 1) * n + S - 1) / S;
    *first = (uint32_t)lo;
    *count = (uint32_t)(hi - lo);
}

void xpngb_packed_offsets(const uint64_t *sizes, uint32_t n, uint64_t *offsets, uint64_t *total) {
    uint64_t off = 0;
    for (uint32_t i = 0; i < n; i++) { offsets[i] = off; off += (sizes[i] + 15) & ~15ull; }
    if (total) *total = off;
}

/* ------------------------------------------------------------------------------------------------ pool */
struct xpngb_pool {
    uint32_t ndev;
    xpngb_ctx **ctx;
    char err[600];
};

typedef struct job {
    xpngb_pool *pool;
    uint32_t shard;
    int encode, level, rc;
    xpngb_image *imgs; uint32_t first, count;
    const uint8_t *in; uint64_t in_size;            /* pixels (encode) / files (decode) */
    const uint64_t *file_offsets, *file_sizes;      /* decode */
    uint8_t *out; uint64_t out_cap;                 /* files (encode) / pixels (decode) */
    uint64_t region_base, region_cap;               /* encode: this shard's slice of `out` */
    uint64_t *out_offsets, *out_sizes;
    char err[520];
} job;

int xpngb_pool_create(xpngb_pool **out, const int *devices, uint32_t ndev) {
    if (!out || !ndev) return 1;
    *out = NULL;
    xpngb_pool *p = calloc(1, sizeof *p);
    if (!p) return 1;
    p->ctx = calloc(ndev, sizeof *p->ctx);
    if (!p->ctx) { free(p); return 1; }
    p->ndev = ndev;
    for (uint32_t k = 0; k < ndev; k++)
        if (xpngb_create(&p->ctx[k], devices ? devices[k] : (int)k)) { xpngb_pool_destroy(p); return 1; }
    *out = p;
    return 0;
}

void xpngb_pool_destroy(xpngb_pool *p) {
    if (!p) return;
    for (uint32_t k = 0; k < p->ndev; k++) if (p->ctx[k]) xpngb_destroy(p->ctx[k]);
    free(p->ctx);
    free(p);
}

uint32_t xpngb_pool_size(const xpngb_pool *p) { return p ? p->ndev : 0; }
const char *xpngb_pool_last_error(const xpngb_pool *p) { return p ? p->err : "no pool"; }
xpngb_ctx *xpngb_pool_context(const xpngb_pool *p, uint32_t k) { return p && k < p->ndev ? p->ctx[k] : NULL; }

static void *job_run(void *arg) {
    job *j = arg;
    xpngb_ctx *ctx = j->pool->ctx[j->shard];
    j->rc = 0; j->err[0] = 0;
    if (!j->count) return NULL;
    xpngb_image *sub = malloc(j->count * sizeof *sub);
    if (!sub) { j->rc = 1; snprintf(j->err, sizeof j->err, "out of memory"); return NULL; }
    memcpy(sub, j->imgs + j->first, j->count * sizeof *sub);
    if (j->encode) {
        /* the shard's pixels as a view of the caller's buffer (offsets are multiples of 16, so the view stays aligned) */
        uint64_t lo = ~0ull, hi = 0;
        for (uint32_t i = 0; i < j->count; i++) {
            const uint64_t a = sub[i].offset, b = a + sub[i].w * sub[i].h * (3 + (sub[i].A ? 1 : 0));
            if (a < lo) lo = a;
            if (b > hi) hi = b;
        }
        if (hi > j->in_size) { j->rc = 1; snprintf(j->err, sizeof j->err, "pixels exceed the buffer"); free(sub); return NULL; }
        for (uint32_t i = 0; i < j->count; i++) sub[i].offset -= lo;
        j->rc = xpngb_encode(ctx, j->level, sub, j->count, j->in + lo, hi - lo, 0, j->out + j->region_base, j->region_cap, 0,
                             j->out_offsets + j->first, j->out_sizes + j->first);
        for (uint32_t i = 0; i < j->count; i++) {
            j->out_offsets[j->first + i] += j->region_base;        /* position in the caller's arena */
            j->imgs[j->first + i].A = sub[i].A; j->imgs[j->first + i].mode = sub[i].mode;
        }
    } else {
        uint64_t flo = ~0ull, fhi = 0, plo = ~0ull;
        for (uint32_t i = 0; i < j->count; i++) {
            const uint64_t a = j->file_offsets[j->first + i], b = a + j->file_sizes[j->first + i];
            if (a < flo) flo = a;
            if (b > fhi) fhi = b;
            if (sub[i].offset < plo) plo = sub[i].offset;
        }
        uint64_t *fo = malloc(j->count * sizeof *fo);
        if (!fo || fhi > j->in_size || plo > j->out_cap) { j->rc = 1; snprintf(j->err, sizeof j->err, "bad file table"); free(fo); free(sub); return NULL; }
        flo &= ~15ull;
        for (uint32_t i = 0; i < j->count; i++) { fo[i] = j->file_offsets[j->first + i] - flo; sub[i].offset -= plo; }
        /* pixels: the shard writes [plo, next shard's plo); the capacity handed down ends at the caller's buffer end */
        j->rc = xpngb_decode(ctx, sub, j->count, j->in + flo, fhi - flo, 0, fo, j->file_sizes + j->first, j->out + plo, j->out_cap - plo, 0);
        for (uint32_t i = 0; i < j->count; i++) {
            xpngb_image *d = &j->imgs[j->first + i];
            d->w = sub[i].w; d->h = sub[i].h; d->A = sub[i].A; d->mode = sub[i].mode;
        }
        free(fo);
    }
    if (j->rc) snprintf(j->err, sizeof j->err, "shard %u: %s", j->shard, xpngb_last_error(ctx));
    free(sub);
    return NULL;
}

static int pool_run(xpngb_pool *p, job *jobs) {
    pthread_t *th = calloc(p->ndev, sizeof *th);
    if (!th) return 1;
    int rc = 0;
    uint32_t started = 0;
    for (uint32_t k = 1; k < p->ndev; k++, started++)
        if (pthread_create(&th[k], NULL, job_run, &jobs[k])) { rc = 1; snprintf(p->err, sizeof p->err, "pthread_create failed"); break; }
    if (!rc) job_run(&jobs[0]);                      /* shard 0 on the calling thread */
    for (uint32_t k = 1; k <= started; k++) pthread_join(th[k], NULL);
    for (uint32_t k = 0; k < p->ndev && !rc; k++)
        if (jobs[k].rc) { rc = 1; snprintf(p->err, sizeof p->err, "%s", jobs[k].err); }
    free(th);
    return rc;
}

int xpngb_pool_encode(xpngb_pool *p, int level, xpngb_image *imgs, uint32_t n, const void *pixels, uint64_t pixels_size,
                      void *out, uint64_t out_cap, uint64_t *out_offsets, uint64_t *out_sizes) {
    if (!p) return 1;
    p->err[0] = 0;
    if (!imgs || !pixels || !out || !out_offsets || !out_sizes) { snprintf(p->err, sizeof p->err, "null argument"); return 1; }
    if (!n) return 0;
    for (uint32_t i = 0; i < n; i++)
        if (!imgs[i].w || !imgs[i].h || imgs[i].w > (1u << 24) || imgs[i].h > (1u << 24)) { snprintf(p->err, sizeof p->err, "image %u: bad dimensions", i); return 1; }
    job *jobs = calloc(p->ndev, sizeof *jobs);
    if (!jobs) return 1;
    uint64_t region = 0;
    for (uint32_t k = 0; k < p->ndev; k++) {
        job *j = &jobs[k];
        j->pool = p; j->shard = k; j->encode = 1; j->level = level; j->imgs = imgs;
        xpngb_shard_range(n, p->ndev, k, &j->first, &j->count);
        j->in = pixels; j->in_size = pixels_size; j->out = out; j->out_cap = out_cap;
        j->region_base = region; j->region_cap = xpngb_encode_bound(imgs + j->first, j->count);
        region += j->region_cap;
        j->out_offsets = out_offsets; j->out_sizes = out_sizes;
    }
    int rc = 0;
    if (region > out_cap) { snprintf(p->err, sizeof p->err, "output buffer must hold xpngb_encode_bound() = %llu bytes", (unsigned long long)region); rc = 1; }
    if (!rc) rc = pool_run(p, jobs);
    free(jobs);
    return rc;
}

int xpngb_pool_decode(xpngb_pool *p, xpngb_image *imgs, uint32_t n, const void *files, uint64_t files_size,
                      const uint64_t *file_offsets, const uint64_t *file_sizes, void *pixels, uint64_t pixels_cap) {
    if (!p) return 1;
    p->err[0] = 0;
    if (!imgs || !files || !file_offsets || !file_sizes || !pixels) { snprintf(p->err, sizeof p->err, "null argument"); return 1; }
    if (!n) return 0;
    job *jobs = calloc(p->ndev, sizeof *jobs);
    if (!jobs) return 1;
    for (uint32_t k = 0; k < p->ndev; k++) {
        job *j = &jobs[k];
        j->pool = p; j->shard = k; j->encode = 0; j->imgs = imgs;
        xpngb_shard_range(n, p->ndev, k, &j->first, &j->count);
        j->in = files; j->in_size = files_size; j->file_offsets = file_offsets; j->file_sizes = file_sizes;
        j->out = pixels; j->out_cap = pixels_cap;
    }
    const int rc = pool_run(p, jobs);
    free(jobs);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ gather */
#define GATHER_MAGIC 0x78706e6762673031ull   /* "xpngbg01" */

typedef struct gather_hdr {
    _Atomic uint64_t magic;
    uint32_t world, max_items;
    _Atomic uint32_t attached;       /* ranks that have mapped this segment: a segment that is already full is a stale one */
    _Atomic uint64_t ready[64];      /* ready[r] = last round rank r has published */
} gather_hdr;

struct xpngb_gather {
    gather_hdr *h;
    uint64_t *slots;                 /* [2][max_items] */
    size_t bytes;
    uint32_t rank, world, max_items;
    uint64_t round;
    char name[96];
    int creator;
};

static double mono_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int xpngb_gather_open(xpngb_gather **out, const char *name, uint32_t rank, uint32_t world, uint32_t max_items) {
    if (!out || !name || !world || world > 64 || rank >= world || !max_items) return 1;
    *out = NULL;
    xpngb_gather *g = calloc(1, sizeof *g);
    if (!g) return 1;
    snprintf(g->name, sizeof g->name, "/xpngb_%s", name);
    for (char *c = g->name + 1; *c; c++) if (*c == '/') *c = '_';
    g->rank = rank; g->world = world; g->max_items = max_items;
    g->bytes = sizeof(gather_hdr) + 2ull * max_items * sizeof(uint64_t);
    const double t_open = mono_s();
    int fd;
again:
    fd = -1;
    if (rank == 0) {
        shm_unlink(g->name);                       /* a stale segment of an earlier run with the same name */
        fd = shm_open(g->name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)g->bytes) != 0) { close(fd); fd = -1; }
        g->creator = 1;
    } else {
        const double t0 = mono_s();
        while (fd < 0 && mono_s() - t0 < 120.0) {   /* rank 0 creates; the others wait for a segment of full size */
            fd = shm_open(g->name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) != 0 || (size_t)st.st_size < g->bytes) { close(fd); fd = -1; }
            }
            if (fd < 0) usleep(1000);
        }
    }
    if (fd < 0) { free(g); return 1; }
    void *m = mmap(NULL, g->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { free(g); return 1; }
    g->h = m; g->slots = (uint64_t *)((uint8_t *)m + sizeof(gather_hdr));
    if (rank == 0) {
        g->h->world = world; g->h->max_items = max_items;
        for (uint32_t r = 0; r < 64; r++) atomic_store(&g->h->ready[r], 0);
        atomic_store(&g->h->attached, 1u);
        atomic_store(&g->h->magic, GATHER_MAGIC);
    } else {
        const double t0 = mono_s();
        while (atomic_load(&g->h->magic) != GATHER_MAGIC) { if (mono_s() - t0 > 120.0) { munmap(m, g->bytes); free(g); return 1; } sched_yield(); }
        /* a segment every rank has already joined belongs to an earlier run that rank 0 is about to replace: try again */
        if (g->h->world != world || g->h->max_items != max_items || atomic_fetch_add(&g->h->attached, 1u) >= world) {
            munmap(m, g->bytes);
            if (mono_s() - t_open > 120.0) { free(g); return 1; }
            usleep(2000);
            goto again;
        }
    }
    *out = g;
    return 0;
}

/* Every rank contributes the sizes of ITS shard of n items (xpngb_shard_range(n, world, rank)); on return every rank
 * holds all n sizes and the packed offset table derived from them.  Collective: all ranks call it the same number of times. */
int xpngb_gather_sizes(xpngb_gather *g, uint32_t n, const uint64_t *local_sizes, uint64_t *all_sizes, uint64_t *all_offsets) {
    if (!g || n > g->max_items || !all_sizes) return 1;
    const uint64_t round = ++g->round;
    uint64_t *slot = g->slots + (round & 1) * (uint64_t)g->max_items;
    uint32_t first, count;
    xpngb_shard_range(n, g->world, g->rank, &first, &count);
    if (count && !local_sizes) return 1;
    for (uint32_t i = 0; i < count; i++) slot[first + i] = local_sizes[i];
    atomic_store_explicit(&g->h->ready[g->rank], round, memory_order_release);
    const double t0 = mono_s();
    for (uint32_t r = 0; r < g->world; r++) {
        uint32_t spins = 0;
        while (atomic_load_explicit(&g->h->ready[r], memory_order_acquire) < round) {
            if (++spins > 2000) { sched_yield(); if (mono_s() - t0 > 300.0) return 1; }
        }
    }
    memcpy(all_sizes, slot, (size_t)n * sizeof(uint64_t));
    if (all_offsets) xpngb_packed_offsets(all_sizes, n, all_offsets, NULL);
    return 0;
}

void xpngb_gather_close(xpngb_gather *g) {
    if (!g) return;
    if (g->h) munmap(g->h, g->bytes);
    if (g->creator) shm_unlink(g->name);
    free(g);
}
