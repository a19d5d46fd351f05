// enc_m2.cuh — level-2 encode specifics: tile classification (single colour / grey / RGB,
// libxpng.c:628-643, :583-626), grey candidate planes, backward 2-state rANS with v1 blocks
// (libxpng.c:160-260), the shared side bit stream, size decisions and assembly (libxpng.c:645-686).
#pragma once
#include "common.cuh"
#include "enc_front.cuh"
#include "enc_back.cuh"

namespace xpb {

// ------------------------------------------------------------------------------------------------
// Tile classification.  One CTA per tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_m2_classify(const TileDesc* __restrict__ tiles, const ImageDesc* __restrict__ imgs,
                                                     uint8_t* __restrict__ tclass) {
    const uint32_t tile = blockIdx.x;
    const TileDesc t = tiles[tile];
    if (imgs[t.img].mode != 2) { if (threadIdx.x == 0) tclass[tile] = TC_NONE; return; }
    const uint8_t* src = reinterpret_cast<const uint8_t*>(t.src_off);
    const uint32_t f0 = src[0], f1 = src[1], f2 = src[2];
    int single = 1, grey = 1;
    const uint32_t first = f0 | (f1 << 8) | (f2 << 16), wq = 3 * t.w >= 16 ? (3 * t.w - 16) / 12 + 1 : 0;   // 16-byte windows inside the row
    for (uint32_t y = threadIdx.x >> 5; y < t.h; y += 8) {
        if (!__any_sync(0xffffffffu, single | grey)) break;              // this warp has seen a pixel that rules out both classes: the tile is RGB
        const uint8_t* row = src + (uint64_t)y * t.bpr;
        for (uint32_t g = threadIdx.x & 31; g < wq; g += 32) {           // four pixels per lane and iteration (ld_rgb4)
            uint32_t v[4]; ld_rgb4(row + 12 * g, v);
#pragma unroll
            for (int k = 0; k < 4; k++) { single &= v[k] == first; grey &= v[k] == (v[k] & 0xFFu) * 0x010101u; }
        }
        for (uint32_t x = (wq << 2) + (threadIdx.x & 31); x < t.w; x += 32) {
            const uint32_t a = row[3 * x], b = row[3 * x + 1], c = row[3 * x + 2];
            single &= (a == f0) & (b == f1) & (c == f2);
            grey &= (a == b) & (b == c);
        }
    }
    single = __syncthreads_and(single);
    grey = __syncthreads_and(grey);
    if (threadIdx.x == 0) tclass[tile] = single ? TC_SINGLE : (grey ? TC_GREY : TC_RGB);
}

// tile_skip[tile] = 1 unless the tile takes the RGB path
__global__ void k_m2_skipmask(const uint8_t* __restrict__ tclass, uint8_t* __restrict__ skip, uint32_t ntiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) skip[i] = tclass[i] != TC_RGB;
}

// ------------------------------------------------------------------------------------------------
// Grey tiles: four candidate residual planes (left, up, avg2, grad3; row 0 left, column 0 up) and
// their 256-bin histograms (libxpng.c:597-604).  One CTA per segment.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t grey_plane_pitch(uint32_t npx) { return (npx + 15u) & ~15u; }

__global__ void __launch_bounds__(256) k_m2_grey_front(const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ seg_tile,
                                                       const uint8_t* __restrict__ tclass, uint8_t* __restrict__ streams,
                                                       uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[4 * 256];
    const uint32_t gseg = blockIdx.x, tile = seg_tile[gseg], tid = threadIdx.x;
    if (tclass[tile] != TC_GREY) return;
    const TileDesc t = tiles[tile];
    for (int k = tid; k < 1024; k += 256) sh[k] = 0;
    __syncthreads();
    const uint32_t r0 = (gseg - t.seg0) * SEG, r1 = min(r0 + (uint32_t)SEG, t.npx);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(t.src_off);
    uint8_t* planes = streams + t.str_off;
    const uint32_t pitch = grey_plane_pitch(t.npx);
    for (uint32_t i = r0 + tid; i < r1; i += 256) {
        if (i == 0) continue;
        const uint32_t y = i / t.w, x = i - y * t.w;
        const uint8_t* p = src + (uint64_t)y * t.bpr + 3ull * x;
        const int v = p[0];
        int c[4];
        if (y == 0) c[0] = c[1] = c[2] = c[3] = v - p[-3];
        else if (x == 0) c[0] = c[1] = c[2] = c[3] = v - p[-(int64_t)t.bpr];
        else {
            const int L = p[-3], U = p[-(int64_t)t.bpr], UL = p[-(int64_t)t.bpr - 3];
            c[0] = v - L; c[1] = v - U; c[2] = v - pred_avg2(L, U); c[3] = v - pred_grad3(L, U, UL);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t u = zz8(c[k]);
            planes[(uint64_t)k * pitch + i - 1] = (uint8_t)u;
            atomicAdd(&sh[k * 256 + u], 1u);
        }
    }
    __syncthreads();
    for (int k = tid; k < 1024; k += 256) { const uint32_t v = sh[k]; if (v) atomicAdd(hist + (uint64_t)tile * HIST_STRIDE_M2 + k, v); }
}

// ------------------------------------------------------------------------------------------------
// v1 block encoder (libxpng.c:160-260): symbols are consumed from the END of the stream, the
// renormalisation words are stored at descending addresses, so the finished block
// [hdr][n][state0][state1][words] sits at the end of the stream's scratch region.  The frequency
// table (or the raw symbols of a type-2 block) is a separate bit-string "piece" that k_assemble_m2
// splices into the tile's shared side stream.
// Stream numbering: RGB tile c = 0..8 contexts (9 symbols), 9..16 values of nl = 1..8; grey tile
// c = 0..3 candidate planes (256 symbols, PB 15).
// ------------------------------------------------------------------------------------------------
constexpr int TAB_WORDS = 136;   // 256 symbols x 16 bits = 128 words (+ slack)

struct RansV1Args {
    const TileDesc* tiles;
    TileState* state;
    const uint32_t* hist;
    const uint8_t* tclass;
    const uint8_t* streams;
    uint8_t* blocks;
    uint32_t* tabs;            // [ntiles][17][TAB_WORDS]
    uint32_t ntiles;
    uint32_t c0, nc;
    uint32_t grey;             // 1: grey candidates, 0: RGB streams
    uint32_t nmin;             // skip alphabets of <= nmin symbols (done by the launch with the smaller table)
    const uint32_t* order = nullptr; const uint32_t* total = nullptr;   // pair encoders: sorted work list instead of the id range
};

__device__ __constant__ const uint16_t M2_NSYM[17] = { 9, 9, 9, 9, 9, 9, 9, 9, 9, 8, 64, 8, 16, 32, 64, 128, 256 };

template <int NSYM, int LANES>
__global__ void __launch_bounds__(LANES) k_rans_v1(RansV1Args A) {
    extern __shared__ __align__(16) uint4 etab[];   // [NSYM][LANES]
    const uint32_t id = blockIdx.x * LANES + threadIdx.x;
    if (id >= A.nc * A.ntiles) return;
    const uint32_t c = A.c0 + id / A.ntiles, tile = id % A.ntiles;
    const uint8_t cls = A.tclass[tile];
    if (cls != (A.grey ? TC_GREY : TC_RGB)) return;
    const TileDesc t = A.tiles[tile];
    TileState* st = A.state + tile;
    uint32_t N, n, rsize; int pb; const uint32_t* F; const uint8_t* in; uint8_t* region;
    if (A.grey) {
        N = 256; pb = 15; n = t.npx - 1;
        F = A.hist + (uint64_t)tile * HIST_STRIDE_M2 + c * 256;
        in = A.streams + t.str_off + (uint64_t)c * grey_plane_pitch(t.npx);
        rsize = align16u(2 * n + 1024);
        region = A.blocks + t.blk_off + (uint64_t)c * rsize;
        st->breg[c] = c * rsize;
    } else {
        N = M2_NSYM[c]; pb = 14; n = st->len[c];
        F = A.hist + (uint64_t)tile * HIST_STRIDE_M2 + (c < 9 ? HIST_CTX + c * 16 : HIST_VAL + VAL_OFF[c - 8]);
        in = A.streams + t.str_off + st->soff[c];
        rsize = align16u(2 * n + 256);
        region = A.blocks + t.blk_off + st->breg[c];
    }
    if (N > (uint32_t)NSYM || N <= A.nmin) return;   // handled by the launch with the other table size
    uint32_t* rend = reinterpret_cast<uint32_t*>(region + rsize);
    uint32_t* tab = A.tabs + ((uint64_t)tile * 17 + c) * TAB_WORDS;
    const uint32_t nbit = bitlen32(N - 1);
    auto finish = [&](uint32_t* start, uint32_t type, uint32_t pbits) {
        st->boff[c] = st->breg[c] + (uint32_t)((uint8_t*)start - region);
        st->bsize[c] = (uint32_t)((uint8_t*)rend - (uint8_t*)start);
        st->btype[c] = type; st->pbits[c] = pbits;
    };
    if (n == 0) { rend[-1] = 4; finish(rend - 1, 0, 0); return; }                         // libxpng.c:167
    uint32_t used = 0;
    for (uint32_t i = 0; i < N; i++) used += F[i] != 0;
    if (used == 1) {                                                                     // libxpng.c:169-172
        rend[-2] = 8u | (1u << 24); rend[-1] = n | ((uint32_t)in[n - 1] << 24);
        finish(rend - 2, 1, 0); return;
    }
    uint32_t cum[NSYM + 1];
    normalise_freqs(F, cum, N, n, pb);
    uint4* E = etab + threadIdx.x;
    for (uint32_t i = 0; i < N; i++) E[i * LANES] = make_encsym(cum[i + 1] - cum[i], cum[i], pb);

    // backward pass; renormalisation words go down from the region end (libxpng.c:215-245)
    uint32_t* wp = rend;
    uint64_t x0 = 1ull << 31, x1 = 1ull << 31;
    auto step = [&](uint64_t& x, uint32_t sym) {
        const uint4 e = E[sym * LANES];
        const uint64_t xmax = (uint64_t)(e.w & 0xFFFFu) << (63 - pb);
        if (x >= xmax) { *--wp = (uint32_t)x; x >>= 32; }
        const uint64_t rcp = (uint64_t)e.x | ((uint64_t)e.y << 32);
        const uint64_t q = __umul64hi(x, rcp) >> (e.w >> 16);
        x += (e.z & 0xFFFFu) + q * (e.z >> 16);
    };
    int64_t i = (int64_t)n - 1;
    for (; i >= 0 && ((i + 1) & 15); i--) { if (i & 1) step(x1, in[i]); else step(x0, in[i]); }
    if (i >= 15) {
        const uint4* in16 = reinterpret_cast<const uint4*>(in);
        uint4 nxt = in16[i >> 4];
        for (; i >= 15; i -= 16) {
            const uint4 v = nxt;
            if (i >= 31) nxt = in16[(i >> 4) - 1];
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 15; k >= 1; k -= 2) {
                const uint32_t s1 = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu, s0 = (w[(k - 1) >> 2] >> (8 * ((k - 1) & 3))) & 0xFFu;
                step(x1, s1);
                step(x0, s0);
            }
        }
    }
    wp -= 4; wp[0] = (uint32_t)x0; wp[1] = (uint32_t)(x0 >> 32); wp[2] = (uint32_t)x1; wp[3] = (uint32_t)(x1 >> 32);   // :245
    const uint32_t payload = (uint32_t)((uint8_t*)rend - (uint8_t*)wp);
    uint32_t tab_bits = (N - used) + used * ((uint32_t)pb + 1);
    const bool sparse = tab_bits < N * (uint32_t)pb;
    if (!sparse) tab_bits = N * (uint32_t)pb;
    if ((uint64_t)tab_bits + 8ull * payload >= (uint64_t)nbit * n) {                      // :250-254 raw symbols
        BitW r{ 0, 0, reinterpret_cast<uint32_t*>(region) };
        for (uint32_t k = 0; k < n; k++) r.put(nbit, in[k]);
        r.end();
        rend[-2] = 8u | (2u << 24); rend[-1] = n;
        finish(rend - 2, 2, nbit * n);
        return;
    }
    BitW b{ 0, 0, tab };                                                                 // :256-257
    for (uint32_t k = 0; k < N; k++) {
        const uint32_t f = cum[k + 1] - cum[k];
        if (!sparse) b.put((uint32_t)pb, f);
        else if (f) b.put((uint32_t)pb + 1, f + (1u << pb));
        else b.put(1, 0);
    }
    b.end();
    wp -= 2; wp[0] = (payload + 8) | ((3u + (uint32_t)sparse) << 24); wp[1] = n;          // :258-259
    finish(wp, 3 + sparse, tab_bits);
}

// ------------------------------------------------------------------------------------------------
// Per-tile size decisions (libxpng.c:606-626, :672-678).  One thread per tile.
// ------------------------------------------------------------------------------------------------
__global__ void k_m2_finish(const TileDesc* __restrict__ tiles, const uint8_t* __restrict__ tclass, TileState* __restrict__ state,
                            uint32_t ntiles) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const TileDesc t = tiles[tile];
    TileState* st = state + tile;
    const uint8_t cls = tclass[tile];
    if (cls == TC_NONE) return;
    if (cls == TC_SINGLE) { st->kind = 3; st->size = 8; return; }
    if (cls == TC_GREY) {
        uint32_t best = 0, bb = 0, br = 0;
        for (uint32_t i = 0; i < 4; i++) {
            const uint32_t bsz = 4 + ((8 + st->pbits[i] + 31) / 32) * 4, rsz = st->bsize[i];
            if (i == 0 || bsz + rsz < bb + br) { best = i; bb = bsz; br = rsz; }          // strict <, ties to the lowest
        }
        if (bb + br >= t.npx) { st->kind = 5; st->size = t.npx + 4; }
        else { st->kind = 4; st->size = bb + br + 4; st->grey_pick = best; }
        return;
    }
    uint32_t bit = 24, rsz = 0;
    for (int c = 0; c < 17; c++) { st->pbo[c] = bit; bit += st->pbits[c]; rsz += st->bsize[c]; }
    st->kbits_lo = bit; st->kbits_hi = 0;
    const uint32_t bsz = 4 + ((bit + 31) / 32) * 4;
    if ((uint64_t)bsz + rsz >= 3ull * t.npx) { st->kind = 0; st->size = 3 * t.npx + 4; }
    else { st->kind = 2; st->size = bsz + rsz + 4; }
}

// ------------------------------------------------------------------------------------------------
// Assembly, level 2.  One CTA per tile.
// ------------------------------------------------------------------------------------------------
struct AssembleM2Args {
    AssembleArgs a;
    const uint8_t* tclass;
    const uint32_t* tabs;
    const uint8_t* streams;
};

__global__ void __launch_bounds__(256) k_assemble_m2(AssembleM2Args M) {
    __shared__ uint64_t sbit[18];
    __shared__ uint32_t snb[18];
    __shared__ const uint32_t* sptr[18];
    __shared__ uint32_t fpw;
    const AssembleArgs& A = M.a;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const TileDesc t = A.tiles[tile];
    const ImageDesc I = A.imgs[t.img];
    const ImageOut O = A.outs[t.img];
    const TileState* st = A.state + tile;
    uint8_t* file = A.out + O.off;
    const uint8_t* src = A.px + t.src_off;
    const uint32_t part = blockIdx.y, nparts = gridDim.y;        // RGB tiles: the side-stream gather and the 17 block copies are sliced over gridDim.y CTAs
    if (part && ((O.mode & 0x100) || (O.mode & 0xFF) == 7 || st->kind != 2)) return;
    if (assemble_common(t, I, O, file, src)) return;
    uint8_t* blob = file + st->out_off;
    const uint8_t* bsrc = A.blocks + t.blk_off;
    switch (st->kind) {
    case 0:   // raw tile (libxpng.c:675-677)
        if (tid == 0) st32u(blob, st->size);
        copy_tile_rows(blob + 4, (uint64_t)t.w * 3, src, t);
        return;
    case 3:   // single colour (libxpng.c:637-640)
        if (tid == 0) { st32u(blob, (255u << 24) | 8u); blob[4] = src[0]; blob[5] = src[1]; blob[6] = src[2]; blob[7] = 0; }
        return;
    case 5:   // raw grey plane (libxpng.c:615-619)
        if (tid == 0) st32u(blob, st->size + (5u << 27));
        for (uint32_t y = tid >> 5; y < t.h; y += 8)
            for (uint32_t x = tid & 31; x < t.w; x += 32) blob[4 + (uint64_t)y * t.w + x] = src[(uint64_t)y * t.bpr + 3ull * x];
        return;
    case 4: { // grey, one v1 block (libxpng.c:622-624)
        const uint32_t g = st->grey_pick;
        const uint32_t bits = 8 + st->pbits[g], words = (bits + 31) / 32;
        if (tid == 0) {
            st32u(blob, st->size + (2u << 28) + (g << 24)); st32u(blob + 4, 4 + 4 * words);
            fpw = (uint32_t)src[0] << 24;
            sbit[0] = 0; snb[0] = 8; sptr[0] = &fpw;
            sbit[1] = 8; snb[1] = st->pbits[g];
            sptr[1] = st->btype[g] == 2 ? reinterpret_cast<const uint32_t*>(bsrc + st->breg[g]) : M.tabs + ((uint64_t)tile * 17 + g) * TAB_WORDS;
        }
        __syncthreads();
        for (uint32_t wi = tid; wi < words; wi += 256) st32u(blob + 8 + 4ull * wi, gather_word(wi, 2, sbit, snb, sptr));
        copy_bytes(blob + 8 + 4ull * words, bsrc + st->boff[g], st->bsize[g]);
        return;
    }
    default: break;
    }
    // RGB: [hdr][bsz][shared side stream][9 context blocks][8 value blocks]   (libxpng.c:680-683)
    const uint32_t bits = st->kbits_lo, words = (bits + 31) / 32;
    if (tid == 0) {
        if (part == 0) { st32u(blob, st->size + (1u << 28) + (st->pr << 24)); st32u(blob + 4, 4 + 4 * words); }
        fpw = ((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 8);
        sbit[0] = 0; snb[0] = 24; sptr[0] = &fpw;
    }
    if (tid < 17) {
        sbit[tid + 1] = st->pbo[tid]; snb[tid + 1] = st->pbits[tid];
        sptr[tid + 1] = st->btype[tid] == 2 ? reinterpret_cast<const uint32_t*>(bsrc + st->breg[tid]) : M.tabs + ((uint64_t)tile * 17 + tid) * TAB_WORDS;
    }
    __syncthreads();
    for (uint32_t wi = part * 256 + tid; wi < words; wi += 256 * nparts) st32u(blob + 8 + 4ull * wi, gather_word(wi, 18, sbit, snb, sptr));
    uint8_t* dst = blob + 8 + 4ull * words;
    for (int c = 0; c < 17; c++) { copy_bytes_part(dst, bsrc + st->boff[c], st->bsize[c], part, nparts); dst += st->bsize[c]; }
}

}  // namespace xpb
