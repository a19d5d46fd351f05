// dec_m1.cuh — level-1 decode: tile offsets, block parsing, v2 rANS decode (libxpng.c:429-493),
// alpha plane, context walk (nl = *cx[nl]++, libxpng.c:803), row prefix tables and the wavefront
// un-predict (libxpng.c:796-830).  Also the stored (level 7) and raw-tile copies.
#pragma once
#include "common.cuh"

namespace xpb {

enum DecErr { DEC_OK = 0, DEC_BAD_TILE_CHAIN = 1, DEC_BAD_BLOCK = 2, DEC_BAD_COUNTS = 3, DEC_BAD_HEADER = 4 };

struct DecImage {
    uint64_t file_off, file_size;   // inside the input buffer
    uint64_t px_off;                // output pixel offset
    uint32_t w, h, pxsz, mode;      // mode: 1, 2, 7; 0x100 = whole-image single colour
    uint32_t tile0, ntiles;
    uint32_t pad[2];
};

struct DecBlock { uint32_t off; uint32_t n; uint32_t type; uint32_t soff; };   // off: from blob start; soff: in stream slice

struct DecTile {
    uint64_t blob_off;       // inside the input buffer
    uint32_t size, m;        // blob size, tile type byte
    uint32_t bsz;            // side/residual bit stream bytes incl. its size word
    uint32_t nsym;           // total context symbols (coded pixels - 1 + ... ) = sum n[0..8]
    DecBlock blk[MAX_STREAMS];
    uint32_t bitpos[MAX_STREAMS];   // mode 2: bit offset of each block's table / raw symbols in the shared stream
    uint32_t pad[2];
};

__device__ __forceinline__ void dec_fail(int* err, int code) { atomicCAS(err, 0, code); }
__host__ __device__ __forceinline__ uint32_t align16u_dec(uint32_t v) { return (v + 15u) & ~15u; }

// ------------------------------------------------------------------------------------------------
// Tile offsets: serial chain over the 24-bit tile sizes of one image (libxpng.c:982).
// ------------------------------------------------------------------------------------------------
__global__ void k_dec_tile_offsets(const DecImage* imgs, const uint8_t* in, DecTile* dt, uint32_t nimg, int* err) {
    const uint32_t im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im >= nimg) return;
    const DecImage I = imgs[im];
    if ((I.mode & 0xFF) == 7 || (I.mode & 0x100)) return;
    uint64_t off = 8;
    for (uint32_t k = 0; k < I.ntiles; k++) {
        DecTile* d = dt + I.tile0 + k;
        if (off + 4 > I.file_size) { dec_fail(err, DEC_BAD_TILE_CHAIN); d->size = 0; d->m = 0xFE; d->blob_off = I.file_off; continue; }
        const uint32_t w0 = ld32u(in + I.file_off + off);
        const uint32_t size = w0 & 0xFFFFFFu;
        d->blob_off = I.file_off + off; d->size = size; d->m = w0 >> 24;
        if (size < 4 || off + size > I.file_size) { dec_fail(err, DEC_BAD_TILE_CHAIN); d->m = 0xFE; }
        off += size < 4 ? 4 : size;
    }
}

// ------------------------------------------------------------------------------------------------
// Level-1 tile parse: bit-stream size, the 9 (10) v2 block headers, stream offsets.
// ------------------------------------------------------------------------------------------------
__global__ void k_dec_parse_m1(const TileDesc* tiles, const DecImage* imgs, const uint8_t* in, DecTile* dt, uint32_t ntiles, int* err) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const TileDesc t = tiles[tile];
    const DecImage I = imgs[t.img];
    if (I.mode != 1) return;
    DecTile* d = dt + tile;
    if (d->m == 0xFE) return;
    if (d->m == 0) { if (d->size != t.npx * t.pxsz + 4) { dec_fail(err, DEC_BAD_TILE_CHAIN); d->m = 0xFE; } return; }
    const uint8_t* blob = in + d->blob_off;
    const int nblk = t.pxsz == 4 ? 10 : 9;
    bool bad = d->size < 8 || (d->m >> 4) != 1 || (((d->m >> 2) & 1) != (t.pxsz == 4));
    uint32_t bsz = bad ? 4 : ld32u(blob + 4);
    if (bsz < 4 || (uint64_t)4 + bsz > d->size) bad = true;
    uint32_t off = 4 + bsz, soff = 0, nsym = 0;
    for (int c = 0; c < nblk && !bad; c++) {
        if (off + 4 > d->size) { bad = true; break; }
        const uint32_t w0 = ld32u(blob + off), type = w0 >> 24;
        uint32_t csz = w0 & 0xFFFFFFu, n = 0;
        if (type == 0) csz = 4;
        else {
            if (type > 4 || csz < 8 || off + csz > d->size) { bad = true; break; }
            n = ld32u(blob + off + 4) & 0xFFFFFFu;
            if (type >= 3 && csz < 28) { bad = true; break; }
        }
        if (c < 9) { nsym += n; if (nsym > t.npx) { bad = true; break; } }
        else if (n != t.npx - 1) { bad = true; break; }
        d->blk[c] = DecBlock{ off, n, type, c < 9 ? soff : 0u };
        if (c < 9) soff += align16u_dec(n);
        off += csz;
    }
    if (!bad && t.pxsz == 3 && nsym != t.npx - 1) bad = true;
    if (bad) { dec_fail(err, DEC_BAD_BLOCK); d->m = 0xFE; return; }
    d->bsz = bsz; d->nsym = nsym;
}

// ------------------------------------------------------------------------------------------------
// v2 block decoder: one lane per block.  Small alphabets (<= 16 symbols) search the cumulative
// table held in registers; the 256-symbol alpha alphabet uses a coarse 256-entry start table plus a
// short linear scan over cum[] in shared memory.
// ------------------------------------------------------------------------------------------------
struct BitR {   // MSB-first reader over unaligned LE words; zero past `end` (libxpng.c:12)
    const uint8_t* base; const uint8_t* end; uint32_t pos;
    __device__ __forceinline__ uint32_t get(uint32_t c) {
        uint32_t v = 0;
        while (c) {
            const uint32_t off = pos & 31u; uint32_t take = 32u - off; if (take > c) take = c;
            const uint8_t* p = base + 4ull * (pos >> 5);
            const uint32_t word = (p + 4 <= end) ? ld32u(p) : 0u;
            const uint32_t bits = (word >> (32u - off - take)) & (take == 32u ? 0xFFFFFFFFu : ((1u << take) - 1u));
            v = take == 32u ? bits : ((v << take) | bits);
            pos += take; c -= take;
        }
        return v;
    }
};

struct RansDecArgs {
    const TileDesc* tiles;
    const DecImage* imgs;
    const DecTile* dt;
    const uint8_t* in;
    uint8_t* streams;     // ctx streams at stream_slice
    uint8_t* alpha;       // alpha symbols at px_off
    uint32_t ntiles;
    uint32_t c0, nc;
    int* err;
};

// decode step shared by all variants: x' = f*(x>>pb) + slot - start, then renormalise from *rp backwards
#define XPB_RENORM_BACK(x)                                                                 \
    if ((x) < (1ull << 31)) { const uint32_t wv = wa; wa = wb; if (rp > lo) rp -= 4; \
        wb = (rp - 8 >= lo) ? ld32u(rp - 8) : 0u; (x) = ((x) << 32) | wv; }

template <int LANES>
__global__ void __launch_bounds__(LANES) k_dec_rans_v2_small(RansDecArgs A) {
    __shared__ uint32_t fs[16 * LANES];   // [sym][lane]: start | freq << 16
    const uint32_t id = blockIdx.x * LANES + threadIdx.x;
    if (id >= A.nc * A.ntiles) return;
    const uint32_t c = A.c0 + id / A.ntiles, tile = id % A.ntiles;
    const TileDesc t = A.tiles[tile];
    if (A.imgs[t.img].mode != 1) return;
    const DecTile* d = A.dt + tile;
    if (d->m == 0 || d->m == 0xFE) return;
    const DecBlock b = d->blk[c];
    const uint8_t* blk = A.in + d->blob_off + b.off;
    uint8_t* out = A.streams + t.str_off + b.soff;
    const uint32_t n = b.n;
    if (b.type == 0 || n == 0) return;
    const uint32_t w1 = ld32u(blk + 4), v2 = w1 >> 24;
    const uint32_t csz = ld32u(blk) & 0xFFFFFFu;
    // this kernel only decodes context streams (c < 9): symbols above 8 can only come from corrupt data and are
    // stored as 0, because the context walk uses them as lane indices
    if (b.type == 1) {   // run of one symbol
        const uint32_t v4 = (v2 > 8u ? 0u : v2) * 0x01010101u;
        for (uint32_t k = 0; k < (n + 3) / 4; k++) reinterpret_cast<uint32_t*>(out)[k] = v4;
        return;
    }
    if (b.type == 2) {   // plain v2-bit symbols
        BitR r{ blk + 8, blk + csz, 0 };
        for (uint32_t k = 0; k < n; k++) { const uint32_t v = r.get(v2); out[k] = (uint8_t)(v > 8u ? 0u : v); }
        return;
    }
    const uint32_t N = v2 + 2, w2 = ld32u(blk + 8); const int pb = (int)(w2 >> 24);
    const uint32_t tabw = w2 & 0xFFFFFFu;
    if (N > 16 || pb < 10 || pb > 15 || 8 + 4ull * tabw > csz || tabw < 5) { for (uint32_t k = 0; k < n; k++) out[k] = 0; return; }
    const uint8_t* tab = blk + 8 + 4ull * tabw; const uint8_t* lo = blk + 12;
    uint32_t cum[17];
    {
        BitR r{ tab, blk + csz, 0 };
        cum[0] = 0;
        for (uint32_t i = 0; i < 16; i++) {
            uint32_t f = 0;
            if (i < N) f = b.type == 3 ? r.get((uint32_t)pb) : (r.get(1) ? r.get((uint32_t)pb) : 0u);
            cum[i + 1] = cum[i] + f;
            fs[i * LANES + threadIdx.x] = (cum[i] & 0xFFFFu) | (f << 16);
        }
    }
    // thresholds: symbol s = #{ i in 1..15 : slot >= cum[i] } restricted to symbols with freq > 0.
    // cum is non-decreasing, zero-width symbols share a threshold with their successor, so the count of
    // thresholds <= slot lands on the last symbol whose start <= slot, i.e. the one with non-zero width.
    uint32_t th[15];
#pragma unroll
    for (int i = 0; i < 15; i++) th[i] = (uint32_t)(i + 1) < N ? cum[i + 1] : 0xFFFFFFFFu;
    const uint32_t mask = (1u << pb) - 1u;
    const uint8_t* rp = tab - 16;
    uint64_t x0 = ld64u(rp), x1 = ld64u(rp + 8);
    uint32_t wa = (rp - 4 >= lo) ? ld32u(rp - 4) : 0u, wb = (rp - 8 >= lo) ? ld32u(rp - 8) : 0u;
    // note: XPB_RENORM_BACK consumes wa (= word at rp-4), then moves rp down by one word.
    auto step = [&](uint64_t& x) -> uint32_t {
        const uint32_t slot = (uint32_t)x & mask;
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 15; i++) s += slot >= th[i];
        const uint32_t e = fs[s * LANES + threadIdx.x];
        x = (uint64_t)(e >> 16) * (x >> pb) + slot - (e & 0xFFFFu);
        XPB_RENORM_BACK(x)
        return s > 8u ? 0u : s;
    };
    int64_t i = (int64_t)n - 1;
    // head: until (i + 1) is a multiple of 4
    for (; i >= 0 && ((i + 1) & 3); i--) out[i] = (uint8_t)((i & 1) ? step(x1) : step(x0));
    for (; i >= 3; i -= 4) {
        const uint32_t s3 = step(x1), s2 = step(x0), s1 = step(x1), s0 = step(x0);
        *reinterpret_cast<uint32_t*>(out + i - 3) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
}

// 256-symbol alphabets (alpha plane, PB = 15): one warp-lane per block, tables in shared memory.
template <int LANES>
__global__ void __launch_bounds__(LANES) k_dec_rans_v2_big(RansDecArgs A) {
    extern __shared__ __align__(16) uint8_t smem_big[];
    // per lane: cum[257] as u16 (pitch 258 -> odd word pitch 129, conflict-free) + coarse[256] u8
    uint16_t* cumS = reinterpret_cast<uint16_t*>(smem_big) + threadIdx.x * 258;
    uint8_t* coarse = smem_big + LANES * 258 * 2 + threadIdx.x * 260;
    const uint32_t tile = blockIdx.x * LANES + threadIdx.x;
    if (tile >= A.ntiles) return;
    const TileDesc t = A.tiles[tile];
    if (A.imgs[t.img].mode != 1 || t.pxsz != 4) return;
    const DecTile* d = A.dt + tile;
    if (d->m == 0 || d->m == 0xFE) return;
    const DecBlock b = d->blk[9];
    const uint8_t* blk = A.in + d->blob_off + b.off;
    uint8_t* out = A.alpha + t.px_off;
    const uint32_t n = b.n;
    if (b.type == 0 || n == 0) return;
    const uint32_t w1 = ld32u(blk + 4), v2 = w1 >> 24, csz = ld32u(blk) & 0xFFFFFFu;
    if (b.type == 1) { const uint32_t v4 = v2 * 0x01010101u; for (uint32_t k = 0; k < (n + 3) / 4; k++) reinterpret_cast<uint32_t*>(out)[k] = v4; return; }
    if (b.type == 2) { BitR r{ blk + 8, blk + csz, 0 }; for (uint32_t k = 0; k < n; k++) out[k] = (uint8_t)r.get(v2); return; }
    const uint32_t N = v2 + 2, w2 = ld32u(blk + 8); const int pb = (int)(w2 >> 24);
    const uint32_t tabw = w2 & 0xFFFFFFu;
    if (N > 256 || pb < 10 || pb > 15 || 8 + 4ull * tabw > csz || tabw < 5) { for (uint32_t k = 0; k < n; k++) out[k] = 0; return; }
    const uint8_t* tab = blk + 8 + 4ull * tabw; const uint8_t* lo = blk + 12;
    {
        BitR r{ tab, blk + csz, 0 };
        uint32_t acc = 0;
        for (uint32_t i = 0; i < 256; i++) {
            cumS[i] = (uint16_t)acc;
            uint32_t f = 0;
            if (i < N) f = b.type == 3 ? r.get((uint32_t)pb) : (r.get(1) ? r.get((uint32_t)pb) : 0u);
            acc += f; if (acc > (1u << pb)) acc = 1u << pb;
        }
        cumS[256] = (uint16_t)(acc >= 65536u ? 65535u : acc);   // 2^15 at most (pb <= 15)
        // coarse[k] = symbol containing slot k << (pb-8)
        uint32_t s = 0;
        for (uint32_t k = 0; k < 256; k++) {
            const uint32_t slot = k << (pb - 8);
            while (s < 255 && (uint32_t)cumS[s + 1] <= slot) s++;
            coarse[k] = (uint8_t)s;
        }
    }
    const uint32_t mask = (1u << pb) - 1u;
    const uint8_t* rp = tab - 16;
    uint64_t x0 = ld64u(rp), x1 = ld64u(rp + 8);
    uint32_t wa = (rp - 4 >= lo) ? ld32u(rp - 4) : 0u, wb = (rp - 8 >= lo) ? ld32u(rp - 8) : 0u;
    auto step = [&](uint64_t& x) -> uint32_t {
        const uint32_t slot = (uint32_t)x & mask;
        uint32_t s = coarse[slot >> (pb - 8)];
        while (s < 255 && (uint32_t)cumS[s + 1] <= slot) s++;
        const uint32_t start = cumS[s], f = (uint32_t)cumS[s + 1] - start;
        x = (uint64_t)f * (x >> pb) + slot - start;
        XPB_RENORM_BACK(x)
        return s;
    };
    int64_t i = (int64_t)n - 1;
    for (; i >= 0 && ((i + 1) & 3); i--) out[i] = (uint8_t)((i & 1) ? step(x1) : step(x0));
    for (; i >= 3; i -= 4) {
        const uint32_t s3 = step(x1), s2 = step(x0), s1 = step(x1), s0 = step(x0);
        *reinterpret_cast<uint32_t*>(out + i - 3) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
}

// ------------------------------------------------------------------------------------------------
// Context walk: nl_{i+1} = next unread symbol of stream nl_i.  Inherently serial per tile; the kernels
// (k_dec_walk_smem, k_dec_walk_lat) live in dec_rans_lat.cuh, this is their argument block.
// ------------------------------------------------------------------------------------------------
struct WalkArgs {
    const TileDesc* tiles;
    const DecImage* imgs;
    const DecTile* dt;
    const uint8_t* streams;
    uint8_t* nlseq;          // per tile at px_off: nl of every coded pixel in raster order
    uint32_t ntiles;
    uint32_t mode;
};

// ------------------------------------------------------------------------------------------------
// Alpha plane (RGBA): un-zig-zag + prefix sums (column 0 down, then along each row), per-row counts
// of coded (alpha != 0) pixels.  One CTA per tile.
// ------------------------------------------------------------------------------------------------
struct RowInfo { uint32_t idx; uint32_t bit; };   // per tile row: index into nlseq and bit offset of its first coded pixel

struct AlphaArgs {
    const TileDesc* tiles;
    const DecImage* imgs;
    const DecTile* dt;
    const uint8_t* in;
    uint8_t* alpha;        // in: symbols (index raster-1) ; out: alpha plane values (index raster), in place via plane
    uint8_t* plane;        // alpha plane per tile at px_off (index raster)
    uint32_t* rowcnt;      // per tile row (at row_off): coded pixels in the row
};

__global__ void __launch_bounds__(256) k_dec_alpha(AlphaArgs A) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const TileDesc t = A.tiles[tile];
    if (A.imgs[t.img].mode != 1 || t.pxsz != 4) return;
    const DecTile* d = A.dt + tile;
    if (d->m == 0 || d->m == 0xFE) return;
    const uint8_t* sym = A.alpha + t.px_off;     // symbol of raster r at sym[r-1]
    uint8_t* pl = A.plane + t.px_off;
    // The first pixel is stored as 32 bits MSB-first: R,G,B,A -> LE word bytes are A,B,G,R.
    const uint32_t fp = ld32u(A.in + d->blob_off + 8);
    const uint32_t alpha0 = fp & 0xFFu;
    // column 0: alpha(0,y) = alpha0 + sum_{yy<=y} unzz(sym[yy*w - 1])  (mod 256), block scan over rows
    if (tid == 0) carry_s = alpha0;
    __syncthreads();
    for (uint32_t y0 = 0; y0 < t.h; y0 += 256) {
        const uint32_t y = y0 + tid;
        int v = (y < t.h && y > 0) ? unzz(sym[(uint64_t)y * t.w - 1]) : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
        if (lane == 31) wsum[wid] = (uint32_t)inc;
        __syncthreads();
        int pre = (int)carry_s;
        for (uint32_t k = 0; k < wid; k++) pre += (int)wsum[k];
        if (y < t.h) pl[(uint64_t)y * t.w] = (uint8_t)(pre + inc);
        __syncthreads();
        if (tid == 255) carry_s = (uint32_t)(pre + inc);
        __syncthreads();
    }
    // rows: warp per row, scan along x in chunks of 32
    for (uint32_t y = wid; y < t.h; y += 8) {
        int run = pl[(uint64_t)y * t.w];
        uint32_t cnt = 0;
        if (lane == 0 && run != 0 && y > 0) cnt = 1;   // (0,0) is never counted as a coded symbol
        for (uint32_t x0 = 1; x0 < t.w; x0 += 32) {
            const uint32_t x = x0 + lane;
            int v = x < t.w ? unzz(sym[(uint64_t)y * t.w + x - 1]) : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
            const uint8_t a = (uint8_t)(run + inc);
            if (x < t.w) { pl[(uint64_t)y * t.w + x] = a; cnt += a != 0; }
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) A.rowcnt[t.row_off + y] = cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// Copies: stored images (level 7), raw tiles (m = 0) and whole-image single colour.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dec_copy(const TileDesc* tiles, const DecImage* imgs, const DecTile* dt, const uint8_t* in, uint8_t* px) {
    const uint32_t tile = blockIdx.x;
    const TileDesc t = tiles[tile];
    const DecImage I = imgs[t.img];
    uint8_t* dst = px + t.src_off;
    const uint32_t rowb = t.w * t.pxsz;
    if ((I.mode & 0xFF) == 7) return;   // stored images are copied by k_load7
    if (I.mode & 0x100) {   // libxpng.c:976-980
        const uint8_t* col = in + I.file_off + 8;
        for (uint32_t y = threadIdx.x >> 5; y < t.h; y += 8)
            for (uint32_t k = threadIdx.x & 31; k < rowb; k += 32) dst[(uint64_t)y * t.bpr + k] = col[k % t.pxsz];
        return;
    }
    const DecTile* d = dt + tile;
    if (d->m == 0) {        // libxpng.c:846
        const uint8_t* src = in + d->blob_off + 4;
        for (uint32_t y = threadIdx.x >> 5; y < t.h; y += 8)
            for (uint32_t k = threadIdx.x & 31; k < rowb; k += 32) dst[(uint64_t)y * t.bpr + k] = src[(uint64_t)y * rowb + k];
    } else if (d->m == 0xFF) {   // single-colour tile (level 2, libxpng.c:916-927)
        const uint8_t* col = in + d->blob_off + 4;
        for (uint32_t y = threadIdx.x >> 5; y < t.h; y += 8)
            for (uint32_t k = threadIdx.x & 31; k < rowb; k += 32) dst[(uint64_t)y * t.bpr + k] = col[k % 3];
    }
}

}  // namespace xpb
