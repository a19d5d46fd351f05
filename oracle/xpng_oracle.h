/*
 * xpng_oracle.h — CPU restatement of the xPNG encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xpng_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker.  It is a from-scratch, single-threaded, in-memory restatement of the
 * algorithm in the reference's libxpng.c; every function cites the reference lines it follows
 * (paths relative to the reference tree).  Parity is PINNED: tests/test_oracle_pin.py compares it
 * byte-for-byte with files produced by the unmodified reference build (oracle/_ref, recipe in
 * oracle/Makefile) and with the committed golden vectors in tests/golden/.
 */
#ifndef XPNG_ORACLE_H
#define XPNG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xo_tile_t { uint64_t x, y, w, h; } xo_tile_t;

/* Tile grid of a W x H image with pxsz bytes per pixel (libxpng.c:51-83).  Writes up to cap tiles
 * in row-major tile order; returns the tile count. */
uint64_t xo_tile_grid(uint64_t W, uint64_t H, int pxsz, xo_tile_t *tiles, uint64_t cap);

/* Alpha normalisation (libxpng.c:688-721).  in: w*h*(3+A) bytes.  out: capacity >= input size.
 * Returns the output alpha flag (0/1) and writes the output byte count to *s_out. */
int xo_normalize(const uint8_t *in, uint64_t w, uint64_t h, int A, uint8_t *out, uint64_t *s_out);

/* Predictor/transform selection of one tile (libxpng.c:92-140).  tile points at the tile's first
 * pixel inside an image whose rows are bpr bytes apart. */
unsigned xo_select_predictor(const uint8_t *tile, uint64_t w, uint64_t h, uint64_t bpr, int pxsz);

/* Mode-1 front end of one tile (libxpng.c:497-532): the ten symbol streams (cx0..cx8, alpha)
 * concatenated into `streams` (capacity >= 2*w*h), their lengths, the 512-bin histogram block
 * (F[ctx*16+nl], alpha at F[256+v]) and the residual bit stream k (words, capacity >= w*h + 4).
 * Returns the number of 32-bit words of k. */
uint64_t xo_m1_front(const uint8_t *tile, uint64_t w, uint64_t h, uint64_t bpr, int pxsz,
                     unsigned pr, uint8_t *streams, uint32_t lens[10], uint32_t F[512],
                     uint32_t *kwords);

/* v2 entropy block (libxpng.c:307-427 / :429-493).  F is modified (normalised) as in the
 * reference.  Encode returns the block size in bytes; decode returns the bytes consumed and
 * writes the symbol count to *n_out. */
uint64_t xo_block_v2_encode(uint32_t *F, unsigned nsym, const uint8_t *in, uint64_t n,
                            uint8_t *out, int prob_bits);
uint64_t xo_block_v2_decode(const uint8_t *in, uint8_t *out, uint64_t *n_out);

/* One tile -> tile blob.  `out` must hold 8*w*h + 4096 bytes.  Return blob size (bytes). */
uint64_t xo_encode_tile_m1(const uint8_t *tile, uint64_t w, uint64_t h, uint64_t bpr, int pxsz,
                           uint8_t *out);
uint64_t xo_encode_tile_m2(const uint8_t *tile, uint64_t w, uint64_t h, uint64_t bpr,
                           uint8_t *out);

/* Whole image -> .xpng file bytes (libxpng.c:723-789).  mode in {1,2,7}.  out capacity must be
 * >= 8 + w*h*(3+A).  Returns the file size, or 0 on a validation failure. */
uint64_t xo_encode(int mode, const uint8_t *px, uint64_t w, uint64_t h, int A, uint8_t *out);

/* Header peek (libxpng.c:969-973): returns 0 on success. */
int xo_peek(const uint8_t *file, uint64_t n, uint64_t *w, uint64_t *h, int *A, int *mode);

/* .xpng file bytes -> pixels (libxpng.c:963-997).  px capacity w*h*(3+A).  Returns 0 on success. */
int xo_decode(const uint8_t *file, uint64_t n, uint8_t *px);

/* Reversible YCoCg-R lifting, the side experiment of Tell_Me_Why/YCoCg-R.c:10-34 (NOT on the
 * .xpng path).  Co and Cg are 9-bit signed. */
void xo_ycocg_r_fwd(int R, int G, int B, int *Y, int *Co, int *Cg);
void xo_ycocg_r_inv(int Y, int Co, int Cg, int *R, int *G, int *B);

#ifdef __cplusplus
}
#endif
#endif
