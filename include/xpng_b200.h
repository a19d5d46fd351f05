/*
 * xpng_b200.h — batch / in-memory C ABI of the B200-native xPNG codec.
 *
 * The reference (alantudyk/xPNG) exposes only a file-at-a-time API (xpng.h:12-20, re-declared in
 * include/xpng.h and implemented on top of this interface).  This header adds what a GPU needs to
 * be measured and sharded: N images in, N `.xpng` byte strings out (and back), with the pixels and
 * the compressed bytes in HOST or DEVICE memory.  Plain C types only; no CUDA or torch types.
 *
 * Reference interfaces replaced:
 *   xpngb_encode  <->  xpng_store_T   (libxpng.c:723-789) minus the fopen/fwrite
 *   xpngb_decode  <->  xpng_load_T    (libxpng.c:963-997) minus the f_read/malloc
 *   xpngb_peek    <->  header parse   (libxpng.c:969-973)
 * All functions return 0 on success and non-zero on failure (the reference's _Bool convention).
 */
#ifndef XPNG_B200_H
#define XPNG_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xpngb_ctx xpngb_ctx;

/* One image of a batch.  `offset` is the byte offset of its pixels (interleaved 8-bit RGB or RGBA,
 * row-major, w*h*(3+A) bytes) inside the pixel buffer; it must be a multiple of 16. */
typedef struct xpngb_image {
    uint64_t w, h;
    uint64_t offset;
    uint32_t A;        /* 1 = RGBA */
    uint32_t mode;     /* out (encode): level actually written (1, 2, 7); out (decode): level found */
} xpngb_image;

/* Create / destroy a codec context bound to CUDA device `device`.  Fails (non-zero) when no usable
 * CUDA device exists: there is no CPU fallback. */
int xpngb_create(xpngb_ctx **ctx, int device);
void xpngb_destroy(xpngb_ctx *ctx);
const char *xpngb_last_error(const xpngb_ctx *ctx);

/* Upper bound of the output bytes xpngb_encode needs for these images (16-byte padded files). */
uint64_t xpngb_encode_bound(const xpngb_image *imgs, uint32_t n);

/* Encode n images at `level` (1, 2 or 7).
 *   pixels / pixels_on_device : base of the pixel buffer (host or device memory), `pixels_size` bytes
 *   out / out_on_device       : receives the .xpng files, file i at out_offsets[i], out_sizes[i] bytes
 * imgs[i].A and imgs[i].mode are updated to what the file header says (alpha may be stripped by the
 * normalisation of libxpng.c:688-721; a level may fall back to 7). */
int xpngb_encode(xpngb_ctx *ctx, int level, xpngb_image *imgs, uint32_t n,
                 const void *pixels, uint64_t pixels_size, int pixels_on_device,
                 void *out, uint64_t out_cap, int out_on_device,
                 uint64_t *out_offsets, uint64_t *out_sizes);

/* Read w, h, A, mode from the first 8 bytes of a .xpng file held in host memory. */
int xpngb_peek(const void *file, uint64_t size, xpngb_image *img);

/* Decode n files.  imgs[i] must carry w, h, A (from xpngb_peek) and `offset` = where image i's
 * pixels go inside `pixels` (multiple of 16).  files / pixels may live on host or device. */
int xpngb_decode(xpngb_ctx *ctx, xpngb_image *imgs, uint32_t n,
                 const void *files, uint64_t files_size, int files_on_device,
                 const uint64_t *file_offsets, const uint64_t *file_sizes,
                 void *pixels, uint64_t pixels_cap, int pixels_on_device);

/* Traversal-order operations of the exchange scheme (Mirroring_and_Rotating/tool.c:3-127; op codes in the order of
 * its option table, tool.c:133).  TL and TR are empty functions in the reference and therefore plain copies. */
enum { XPNGB_OP_R90 = 0, XPNGB_OP_R270 = 1, XPNGB_OP_MV = 2, XPNGB_OP_MH = 3, XPNGB_OP_MVH = 4, XPNGB_OP_TL = 5, XPNGB_OP_TR = 6 };

/* Apply `op` to n pixmaps: image i is read at src + imgs[i].offset and written at dst + imgs[i].offset (same byte
 * count, so the same offsets serve both buffers; src and dst must not overlap).  imgs[i].w and .h are swapped for the
 * quarter turns.  Replaces op_mv/op_mh/op_mvh/op_r90/op_r270 (tool.c:3-119). */
int xpngb_transform(xpngb_ctx *ctx, int op, xpngb_image *imgs, uint32_t n,
                    const void *src, uint64_t size, int src_on_device, void *dst, int dst_on_device);

/* xpngb_encode of the images as `op` would leave them — the "change the traversal order instead of transforming the
 * pixmap" idea of Mirroring_and_Rotating/README.md:1-4: the files are byte-identical to `tool --op` followed by
 * `xpng -level`, the caller's pixels are not touched, imgs[i].w/.h report the encoded orientation. */
int xpngb_encode_oriented(xpngb_ctx *ctx, int level, int op, xpngb_image *imgs, uint32_t n,
                          const void *pixels, uint64_t pixels_size, int pixels_on_device,
                          void *out, uint64_t out_cap, int out_on_device,
                          uint64_t *out_offsets, uint64_t *out_sizes);

/* Device time of the kernels of the last encode/decode call on this context, in milliseconds
 * (CUDA events on the context's stream; excludes host<->device copies). */
float xpngb_last_kernel_ms(const xpngb_ctx *ctx);
/* Number of kernel launches issued by the last call. */
uint32_t xpngb_last_launches(const xpngb_ctx *ctx);
/* Per-kernel timing: while on, every launch is bracketed by CUDA events and synchronised (so the call
 * gets slower); xpngb_profile_report writes "kernel_name total_ms launches\n" lines and returns the
 * length.  Turning profiling on or off clears the table. */
void xpngb_profile(xpngb_ctx *ctx, int on);
uint32_t xpngb_profile_report(const xpngb_ctx *ctx, char *buf, uint32_t cap);
/* The CUDA stream (cudaStream_t as void*) the context launches on. */
void *xpngb_stream(const xpngb_ctx *ctx);

/* ---- Frame batches sharded over the GPUs of one box (host code in C, no NCCL).  The reference's parallel strategy is a
 * tile cursor plus an ordered concatenation (libxpng.c:146-151, :764-769; offset chain :982); frames are independent, so a
 * batch is cut into contiguous shards, every shard is coded on its own device with no data-path exchange, and the one
 * exchange is the table of compressed sizes, from which every participant derives the same offsets. */

/* Frame i of n belongs to shard floor(i * nshards / n); shard k owns [first, first + count). */
void xpngb_shard_range(uint32_t n, uint32_t nshards, uint32_t shard, uint32_t *first, uint32_t *count);
/* Exclusive scan of 16-byte padded sizes: where file i starts when the n files are written back to back. */
void xpngb_packed_offsets(const uint64_t *sizes, uint32_t n, uint64_t *offsets, uint64_t *total);

/* One process: a pool holds one codec context per device and drives each from its own host thread. */
typedef struct xpngb_pool xpngb_pool;
int xpngb_pool_create(xpngb_pool **pool, const int *devices /* NULL: 0..ndev-1 */, uint32_t ndev);
void xpngb_pool_destroy(xpngb_pool *pool);
uint32_t xpngb_pool_size(const xpngb_pool *pool);
xpngb_ctx *xpngb_pool_context(const xpngb_pool *pool, uint32_t k);
const char *xpngb_pool_last_error(const xpngb_pool *pool);
/* xpngb_encode / xpngb_decode of n images in HOST memory, shard k on pool device k.  Encode: `out` must hold
 * xpngb_encode_bound(imgs, n) bytes; shard k writes its files into its own slice of `out`, out_offsets[i] is the
 * position of file i inside `out`, and the files are byte-identical to a one-device call. */
int xpngb_pool_encode(xpngb_pool *pool, int level, xpngb_image *imgs, uint32_t n, const void *pixels, uint64_t pixels_size,
                      void *out, uint64_t out_cap, uint64_t *out_offsets, uint64_t *out_sizes);
int xpngb_pool_decode(xpngb_pool *pool, xpngb_image *imgs, uint32_t n, const void *files, uint64_t files_size,
                      const uint64_t *file_offsets, const uint64_t *file_sizes, void *pixels, uint64_t pixels_cap);

/* One process per device (torchrun-style launch): the size tables of the ranks meet in a POSIX shared-memory segment.
 * `name` must be the same on every rank and unique per job (e.g. the rendezvous port); rank 0 creates the segment. */
typedef struct xpngb_gather xpngb_gather;
int xpngb_gather_open(xpngb_gather **g, const char *name, uint32_t rank, uint32_t world, uint32_t max_items);
/* Collective: rank r passes the sizes of its shard (xpngb_shard_range(n, world, r)); every rank receives all n sizes and
 * the packed offsets derived from them. */
int xpngb_gather_sizes(xpngb_gather *g, uint32_t n, const uint64_t *local_sizes, uint64_t *all_sizes, uint64_t *all_offsets);
void xpngb_gather_close(xpngb_gather *g);

/* Reversible YCoCg-R lifting of n RGB triples on the device (Tell_Me_Why/YCoCg-R.c:22,:31): a side
 * component, NOT part of the .xpng bit stream.  rgb: 3*n bytes; ycc: 3*n int16 (Y, Co, Cg). */
int xpngb_ycocg_forward(xpngb_ctx *ctx, const uint8_t *rgb_host, int16_t *ycc_host, uint64_t n);
int xpngb_ycocg_inverse(xpngb_ctx *ctx, const int16_t *ycc_host, uint8_t *rgb_host, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
