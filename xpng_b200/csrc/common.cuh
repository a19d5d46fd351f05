// common.cuh — shared device/host definitions for the xPNG B200 kernels (sm_100a).
// Bit-exact integer arithmetic of the reference's hot path; citations are libxpng.c:line.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xpb {

constexpr int SEG = 4096;            // pixels of a tile's raster sequence handled by one front-end CTA
constexpr int FRONT_THREADS = 256;   // SEG / FRONT_THREADS = 16 consecutive entries per thread in phase 2
constexpr int PPT = SEG / FRONT_THREADS;
constexpr int SEG_BITS_BYTES = SEG * 3;  // worst case 24 bits per pixel
constexpr uint32_t TILE_AREA = 444u * 444u;  // libxpng.c:49

// One tile of one image of the batch (host-built, libxpng.c:51-83).
struct TileDesc {
    uint64_t src_off;   // address of the tile's first pixel relative to the `px` kernel argument (absolute when px == nullptr)
    uint64_t px_off;    // running sum of (16-padded pixel count + 256): base of this tile's 1-byte-per-pixel scratch slices
    uint64_t str_off;   // byte offset of the tile's slice in the stream scratch
    uint64_t blk_off;   // byte offset of the tile's slice in the entropy-block scratch
    uint64_t row_off;   // running sum of tile heights: base of this tile's per-row tables
    uint32_t w, h;      // tile size in pixels
    uint32_t bpr;       // bytes per image row
    uint32_t npx;       // w*h
    uint32_t img;       // image index in the batch
    uint32_t seg0;      // first global segment index
    uint32_t nseg;      // ceil(npx / SEG)
    uint32_t pxsz;      // 3 or 4
    uint32_t x0, y0;    // tile origin inside the image (pixels)
    uint32_t tix;       // tile index inside the image
    uint32_t pad;
};

// One image of the batch.
struct ImageDesc {
    uint64_t px_off;    // byte offset of the image's pixels in the pixel buffer (16-aligned)
    uint64_t raw_size;  // w*h*pxsz
    uint32_t w, h;
    uint32_t pxsz;
    uint32_t tile0, ntiles;
    uint32_t mode;      // requested / effective level (1, 2, 7); 0x100 flag = whole-image single colour
    uint32_t pad[2];
};

// Per-segment results of the encode front end.
struct SegInfo {
    uint16_t cnt[9];    // symbols this segment put into context chunk c (segment-first symbol excluded)
    uint16_t nvalid;    // coded (non-skipped) pixels in the segment
    uint32_t nbits;     // residual bits produced by the segment
    uint8_t has_valid, first_nl, last_nl, pad;
    uint32_t pad2;
};
static_assert(sizeof(SegInfo) == 32, "SegInfo layout");

// Per-segment placement computed by the per-tile scan.
struct SegPlace {
    uint32_t pos[9];    // destination index inside stream c of this segment's chunk c
    uint32_t first_pos; // destination index (in stream first_ctx) of the segment-first symbol
    uint32_t first_ctx;
    uint32_t pad;
    uint64_t bit_off;   // bit offset of this segment's residual bits inside the tile's k stream
    uint64_t pad2;
};
static_assert(sizeof(SegPlace) == 64, "SegPlace layout");

constexpr int MAX_STREAMS = 17;      // mode 1: 9 contexts + alpha; mode 2: 9 contexts + 8 value streams

// Per-tile encode state.
struct TileState {
    uint32_t pr;                 // predictor code (libxpng.c:139)
    uint32_t kind;               // 0 raw, 1 coded mode-1, 2 coded mode-2 RGB, 3 single colour, 4 grey rANS, 5 grey raw
    uint32_t kbits_lo, kbits_hi; // total bits of the side/residual bit stream (incl. first pixel)
    uint32_t len[MAX_STREAMS];   // symbols per stream
    uint32_t soff[MAX_STREAMS];  // byte offset of stream inside the tile's stream slice (16-aligned)
    uint32_t breg[MAX_STREAMS];  // byte offset of the stream's block scratch region inside the tile's block slice
    uint32_t boff[MAX_STREAMS];  // byte offset of the finished block (== breg for v2 blocks, end-aligned for v1)
    uint32_t bsize[MAX_STREAMS]; // final block size in bytes
    uint32_t pbits[MAX_STREAMS]; // level 2: bits this block adds to the tile's shared side stream (table or raw symbols)
    uint32_t pbo[MAX_STREAMS];   // level 2: bit offset of that piece inside the side stream
    uint32_t btype[MAX_STREAMS]; // level 2: block type 0..4
    uint32_t size;               // final tile blob size
    uint32_t grey_pick;          // chosen grey predictor (mode 2)
    uint64_t out_off;            // byte offset of the blob in the output arena
};

// Tile classes of level 2 (libxpng.c:655: single colour, then grey, then RGB)
enum TileClass { TC_RGB = 0, TC_SINGLE = 1, TC_GREY = 2, TC_NONE = 3 };

// ---------------------------------------------------------------- scalar helpers

__host__ __device__ __forceinline__ uint32_t bitlen32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return 32u - (uint32_t)__clz((int)v);
#else
    return v ? 32u - (uint32_t)__builtin_clz(v) : 0u;
#endif
}
// zig-zag of the residual wrapped to int8 (libxpng.c:20)
__host__ __device__ __forceinline__ uint32_t zz8(int v) {
    int s = (int)(int8_t)v;
    return (uint32_t)((s << 1) ^ (s >> 31)) & 0xFFu;
}
__host__ __device__ __forceinline__ int unzz(uint32_t u) { return (int)(u >> 1) ^ -(int)(u & 1u); }   // :21
__host__ __device__ __forceinline__ int pred_avg2(int L, int U) { return (L + U + 1) >> 1; }              // :27
__host__ __device__ __forceinline__ int pred_grad3(int L, int U, int UL) { return (3 * L + 3 * U - 2 * UL + 2) >> 2; }  // :29

// Argmin of the four candidate costs, ties to the lowest (libxpng.c:133-139).
__host__ __device__ __forceinline__ uint32_t pick_predictor(const uint32_t c[4], uint32_t w, uint32_t h, uint32_t pxsz) {
    if (w < 4 || h < 4) return 0;
    uint32_t m = 0, r = c[0];
    if (c[1] < r) { m = 1; r = c[1]; }
    if (c[2] < r) { m = 2; r = c[2]; }
    if (c[3] < r) { m = 3; r = c[3]; }
    return (pxsz & 4u) | m;
}

#ifdef __CUDACC__
// Unaligned little-endian accessors (tile blobs follow raw tiles of odd size, so nothing is aligned).
__device__ __forceinline__ uint32_t ld32u(const uint8_t* p) {
    if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) return *reinterpret_cast<const uint32_t*>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ void st32u(uint8_t* p, uint32_t v) {
    if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) { *reinterpret_cast<uint32_t*>(p) = v; return; }
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ uint64_t ld64u(const uint8_t* p) { return (uint64_t)ld32u(p) | ((uint64_t)ld32u(p + 4) << 32); }

// Four consecutive RGB pixels (12 bytes) at ANY byte alignment as 0x00BBGGRR each: four aligned 32-bit loads and
// three funnel shifts instead of twelve byte loads.  Touches the aligned words covering p .. p + 15 (callers keep
// 4 bytes of slack; every pixel buffer of the codec has it).
__device__ __forceinline__ void ld_rgb4(const uint8_t* p, uint32_t px[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    const uint32_t a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2), a3 = __ldg(q + 3);
    const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh);
    px[0] = w0 & 0xFFFFFFu; px[1] = (w0 >> 24) | ((w1 & 0xFFFFu) << 8); px[2] = (w1 >> 16) | ((w2 & 0xFFu) << 16); px[3] = w2 >> 8;
}

// Load one pixel as 0x00BBGGRR / 0xAABBGGRR.
template <int PXSZ>
__device__ __forceinline__ uint32_t ld_pixel(const uint8_t* p) {
    if (PXSZ == 4) return *reinterpret_cast<const uint32_t*>(p);   // RGBA images are 16-byte aligned in the batch
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}
#endif

}  // namespace xpb
