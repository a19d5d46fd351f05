// enc_front.cuh — encode front end: predictor selection, per-pixel residuals, context split,
// residual bit packing.  Restates libxpng.c:92-140 (pp_rgbx) and :497-532 (m1e_*) as data-parallel
// kernels.  A tile's raster sequence is cut into SEG-pixel segments; one CTA per segment.
#pragma once
#include "common.cuh"

namespace xpb {

// ------------------------------------------------------------------------------------------------
// Predictor cost: sum over the 4x4 lattice (x,y == 3 mod 4) of bitlen(OR of zig-zagged residuals)
// for {avg2, avg2+G, grad3, grad3+G}.  grid = (ntiles, PP_SPLIT); costs must be zeroed.
// ------------------------------------------------------------------------------------------------
constexpr int PP_SPLIT = 4;
constexpr uint32_t FRONT2_MAXW = 670;   // widest tile the staged RGB front end (enc_front2.cuh) takes; wider ones are one-tile images

template <int PXSZ>
__device__ __forceinline__ void predictor_cost_tile(const TileDesc& t, const uint8_t* __restrict__ px,
                                                    uint32_t* __restrict__ cost4) {
    const uint32_t nx = t.w >> 2, ny = t.h >> 2, total = nx * ny;
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const uint8_t* base = px + t.src_off;
    for (uint32_t idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
        const uint32_t j = idx / nx, i = idx - j * nx;
        const uint8_t* p = base + (uint64_t)(3 + 4 * j) * t.bpr + (uint64_t)(3 + 4 * i) * PXSZ;
        const uint32_t cur = ld_pixel<PXSZ>(p);
        if (PXSZ == 4 && (cur >> 24) == 0) continue;   // libxpng.c:121
        const uint32_t lf = ld_pixel<PXSZ>(p - PXSZ), up = ld_pixel<PXSZ>(p - t.bpr), ul = ld_pixel<PXSZ>(p - t.bpr - PXSZ);
        int a[3], g[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int v = (cur >> (8 * c)) & 255, L = (lf >> (8 * c)) & 255, U = (up >> (8 * c)) & 255, UL = (ul >> (8 * c)) & 255;
            a[c] = v - pred_avg2(L, U);
            g[c] = v - pred_grad3(L, U, UL);
        }
        c0 += bitlen32(zz8(a[0]) | zz8(a[1]) | zz8(a[2]));
        c1 += bitlen32(zz8(a[0] - a[1]) | zz8(a[1]) | zz8(a[2] - a[1]));
        c2 += bitlen32(zz8(g[0]) | zz8(g[1]) | zz8(g[2]));
        c3 += bitlen32(zz8(g[0] - g[1]) | zz8(g[1]) | zz8(g[2] - g[1]));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(cost4 + 0, c0);
        if (c1) atomicAdd(cost4 + 1, c1);
        if (c2) atomicAdd(cost4 + 2, c2);
        if (c3) atomicAdd(cost4 + 3, c3);
    }
}

__global__ void __launch_bounds__(256) k_predictor_cost(const TileDesc* __restrict__ tiles, const uint8_t* __restrict__ px,
                                                        uint32_t* __restrict__ costs) {
    const TileDesc t = tiles[blockIdx.x];
    if (t.w < 4 || t.h < 4) return;
    if (t.pxsz == 4) predictor_cost_tile<4>(t, px, costs + 4 * blockIdx.x);
    else predictor_cost_tile<3>(t, px, costs + 4 * blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Block-wide scans used by the front end (256 threads = 8 warps).
// ------------------------------------------------------------------------------------------------
struct ScanA { uint32_t sum; uint32_t last; };   // sum: nbits | nvalid<<17 ; last: 0x10|nl of the last coded pixel, 0 if none

__device__ __forceinline__ ScanA combineA(ScanA a, ScanA b) { return ScanA{ a.sum + b.sum, (b.last & 0x10u) ? b.last : a.last }; }

// Exclusive scan over the block; returns the exclusive prefix and writes the block total.
__device__ __forceinline__ ScanA block_scan_A(ScanA v, ScanA* warp_tot /*[8]*/, ScanA& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ScanA inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        ScanA n{ __shfl_up_sync(0xffffffffu, inc.sum, o), __shfl_up_sync(0xffffffffu, inc.last, o) };
        if (lane >= o) inc = combineA(n, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    ScanA pre{ 0, 0 }, tot{ 0, 0 };
#pragma unroll
    for (int k = 0; k < FRONT_THREADS / 32; k++) {
        const ScanA w = warp_tot[k];
        if (k < wid) pre = combineA(pre, w);
        tot = combineA(tot, w);
    }
    total = tot;
    ScanA exl{ __shfl_up_sync(0xffffffffu, inc.sum, 1), __shfl_up_sync(0xffffffffu, inc.last, 1) };
    if (lane == 0) exl = ScanA{ 0, 0 };
    __syncthreads();
    return combineA(pre, exl);
}

// Nine 16-bit counters in three 64-bit words.
struct Cnt9 { unsigned long long a, b, c; };
__device__ __forceinline__ Cnt9 operator+(Cnt9 x, Cnt9 y) { return Cnt9{ x.a + y.a, x.b + y.b, x.c + y.c }; }
__device__ __forceinline__ uint32_t cnt9_get(const Cnt9& v, uint32_t k) {
    const unsigned long long w = k < 4 ? v.a : (k < 8 ? v.b : v.c);
    return (uint32_t)(w >> (16 * (k & 3))) & 0xFFFFu;
}
__device__ __forceinline__ void cnt9_inc(Cnt9& v, uint32_t k) {
    const unsigned long long one = 1ull << (16 * (k & 3));
    v.a += k < 4 ? one : 0ull; v.b += (k >= 4 && k < 8) ? one : 0ull; v.c += k >= 8 ? one : 0ull;
}

__device__ __forceinline__ Cnt9 block_scan_cnt9(Cnt9 v, Cnt9* warp_tot /*[8]*/, Cnt9& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Cnt9 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Cnt9 n{ __shfl_up_sync(0xffffffffu, inc.a, o), __shfl_up_sync(0xffffffffu, inc.b, o), __shfl_up_sync(0xffffffffu, inc.c, o) };
        if (lane >= o) inc = n + inc;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    Cnt9 pre{ 0, 0, 0 }, tot{ 0, 0, 0 };
#pragma unroll
    for (int k = 0; k < FRONT_THREADS / 32; k++) {
        const Cnt9 w = warp_tot[k];
        if (k < wid) pre = pre + w;
        tot = tot + w;
    }
    total = tot;
    Cnt9 exl{ __shfl_up_sync(0xffffffffu, inc.a, 1), __shfl_up_sync(0xffffffffu, inc.b, 1), __shfl_up_sync(0xffffffffu, inc.c, 1) };
    if (lane == 0) exl = Cnt9{ 0, 0, 0 };
    __syncthreads();
    return pre + exl;
}

// ------------------------------------------------------------------------------------------------
// Front end, modes 1 and 2.  One CTA per segment.
//   phase 1 (interleaved, coalesced): residual / zig-zag / nl / packed field per pixel, alpha symbol.
//   phase 2 (blocked, 16 consecutive entries per thread): context keys, stable 9-way split into the
//           segment's chunks, 81-bin histogram, residual bits (mode 1) or value bytes (mode 2).
//   phase 3: coalesced copy-out of the segment's chunks and bits.
// The first coded pixel of a segment does not know its context (the last coded pixel before the
// segment); it is reported in SegInfo and placed by k_tile_scan.
// ------------------------------------------------------------------------------------------------
struct FrontShared {
    uint8_t nl[SEG];                        // 0xFF = not coded
    uint32_t fld[(SEG / 16) * 20];          // 16 fields per thread, 80-byte pitch (conflict-free LDS.128)
    uint32_t bits[SEG_BITS_BYTES / 4];      // mode 1: residual bits; mode 2: value bytes
    uint8_t sym[SEG];                       // context chunks, concatenated
    uint32_t hist[9 * 16];
    uint32_t hist2[576];                    // mode 1: alpha histogram [256]; mode 2: value histograms (VAL_OFF)
    ScanA wa[FRONT_THREADS / 32];
    Cnt9 wc[FRONT_THREADS / 32];
    uint32_t chunk_start[9];
    uint32_t vchunk_start[9];
    uint32_t first_nl;
    uint32_t vbytes;                        // mode 2: bytes in the value area
};

// Mode-2 value alphabets (libxpng.c:669): nl=1..8 -> 8,64,8,16,32,64,128,256 symbols, packed back to back.
__device__ __constant__ const uint16_t VAL_OFF[10] = { 0, 0, 8, 72, 80, 96, 128, 192, 320, 576 };

struct FrontArgs {
    const TileDesc* tiles;
    const uint32_t* seg_tile;     // global segment -> tile
    const uint8_t* px;
    const uint32_t* costs;        // [ntiles][4]
    const uint8_t* tile_skip;     // mode 2: 1 = tile handled elsewhere (single colour / grey); may be null
    SegInfo* seginfo;
    uint8_t* sym_area;            // [nseg_total][SEG]
    uint8_t* bits_area;           // [nseg_total][SEG_BITS_BYTES]
    uint8_t* alpha_area;          // per tile at px_off: raster-order alpha symbols (index raster-1)
    uint32_t* hist;               // per tile: HIST_STRIDE u32
    uint16_t* vcnt;               // mode 2: [nseg_total][9] value-chunk element counts
    uint32_t rgba_only;           // 1: RGB tiles are handled by k_front2 (enc_front2.cuh), this launch only takes the RGBA ones
};

constexpr int HIST_CTX = 0;        // [9][16] context histograms
constexpr int HIST_ALPHA = 256;    // [256] alpha (mode 1)
constexpr int HIST_VAL = 256;      // mode 2: value histograms for nl = 1..8 at HIST_VAL + VAL_OFF[nl]
constexpr int HIST_STRIDE_M1 = 512;
constexpr int HIST_STRIDE_M2 = 1024;   // 256 ctx + 576 value bins, or 4 x 256 grey candidate bins

template <int MODE, int PXSZ>
__device__ __forceinline__ void front_segment(const FrontArgs& A, FrontShared& S, const TileDesc& t, uint32_t tile, uint32_t gseg) {
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t j = gseg - t.seg0;
    const uint32_t r0 = j * SEG, r1 = min(r0 + (uint32_t)SEG, t.npx);
    const uint32_t pr = pick_predictor(A.costs + 4 * tile, t.w, t.h, MODE == 2 ? 3u : t.pxsz);
    const bool Y = (pr >> 1) & 1, G = pr & 1;
    constexpr int HSTRIDE = MODE == 1 ? HIST_STRIDE_M1 : HIST_STRIDE_M2;

    for (uint32_t k = tid; k < SEG_BITS_BYTES / 4; k += FRONT_THREADS) S.bits[k] = 0;
    if (tid < 144) S.hist[tid] = 0;
    if (MODE == 2 || PXSZ == 4) for (uint32_t k = tid; k < 576; k += FRONT_THREADS) S.hist2[k] = 0;
    __syncthreads();

    // ---------------- phase 1
    const uint8_t* base = A.px + t.src_off;
    uint32_t i = r0 + tid;
    uint32_t y = i / t.w, x = i - y * t.w;
#pragma unroll 1
    for (int it = 0; it < PPT; it++, i += FRONT_THREADS) {
        const bool in = i < r1;
        const uint8_t* p = base + (uint64_t)y * t.bpr + (uint64_t)x * PXSZ;
        uint32_t cur = 0, up = 0;
        if (in) { cur = ld_pixel<PXSZ>(p); if (y) up = ld_pixel<PXSZ>(p - t.bpr); }
        uint32_t lf = __shfl_up_sync(0xffffffffu, cur, 1), ul = __shfl_up_sync(0xffffffffu, up, 1);
        if (lane == 0 && in && x) { lf = ld_pixel<PXSZ>(p - PXSZ); ul = y ? ld_pixel<PXSZ>(p - t.bpr - PXSZ) : 0; }
        bool skip = !in || i == 0;
        if (PXSZ == 4 && in && i) {   // alpha plane symbol, libxpng.c:497-502
            const int ap = x ? (int)(lf >> 24) : (int)(up >> 24);
            const uint32_t v = zz8((int)(cur >> 24) - ap);
            A.alpha_area[t.px_off + i - 1] = (uint8_t)v;
            atomicAdd(&S.hist2[v], 1u);
            if ((cur >> 24) == 0) skip = true;
        }
        int r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int v = (cur >> (8 * c)) & 255, L = (lf >> (8 * c)) & 255, U = (up >> (8 * c)) & 255, UL = (ul >> (8 * c)) & 255;
            const int pd = y == 0 ? L : (x == 0 ? U : (Y ? pred_grad3(L, U, UL) : pred_avg2(L, U)));
            r[c] = v - pd;
        }
        if (G && x && y) { r[0] -= r[1]; r[2] -= r[1]; }
        const uint32_t u0 = zz8(r[0]), u1 = zz8(r[1]), u2 = zz8(r[2]);
        const uint32_t nl = bitlen32(u0 | u1 | u2);
        const uint32_t li = it * FRONT_THREADS + tid;
        S.nl[li] = skip ? 0xFFu : (uint8_t)nl;
        S.fld[(li >> 4) * 20 + (li & 15)] = MODE == 1 ? ((u0 << (2 * nl)) | (u1 << nl) | u2) : ((u0 << 16) | (u1 << 8) | u2);
        x += FRONT_THREADS;
        while (x >= t.w) { x -= t.w; y++; }
    }
    __syncthreads();

    // ---------------- phase 2
    uint32_t nlw[4];
    {
        const uint4 q = *reinterpret_cast<const uint4*>(&S.nl[tid * 16]);
        nlw[0] = q.x; nlw[1] = q.y; nlw[2] = q.z; nlw[3] = q.w;
    }
    uint32_t fld[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 f = *reinterpret_cast<const uint4*>(&S.fld[tid * 20 + q * 4]);
        fld[q * 4 + 0] = f.x; fld[q * 4 + 1] = f.y; fld[q * 4 + 2] = f.z; fld[q * 4 + 3] = f.w;
    }
    ScanA mine{ 0, 0 };
    Cnt9 vc{ 0, 0, 0 };   // mode 2: per-nl value counts
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const uint32_t nl = (nlw[e >> 2] >> (8 * (e & 3))) & 0xFFu;
        if (nl != 0xFFu) {
            mine.sum += (MODE == 1 ? 3 * nl : 0) + (1u << 17); mine.last = 0x10u | nl;
            if (MODE == 2) cnt9_inc(vc, nl);
        }
    }
    ScanA totA;
    const ScanA preA = block_scan_A(mine, S.wa, totA);

    // context keys and per-thread counts
    Cnt9 cc{ 0, 0, 0 };
    {
        uint32_t pl = (preA.last & 0x10u) ? (preA.last & 0xFu) : 15u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint32_t nl = (nlw[e >> 2] >> (8 * (e & 3))) & 0xFFu;
            if (nl != 0xFFu) { if (pl != 15u) cnt9_inc(cc, pl); pl = nl; }
        }
    }
    Cnt9 totC;
    Cnt9 pos = block_scan_cnt9(cc, S.wc, totC);
    if (tid < 9) {
        uint32_t s = 0;
        for (uint32_t c = 0; c < tid; c++) s += cnt9_get(totC, c);
        S.chunk_start[tid] = s;
    }
    Cnt9 vpos{ 0, 0, 0 };
    if (MODE == 2) {
        Cnt9 totV;
        vpos = block_scan_cnt9(vc, S.wc, totV);
        if (tid < 9) {   // value chunks in bytes: nl 1,2 -> 1 byte per pixel, nl >= 3 -> 3 bytes
            uint32_t s = 0;
            for (uint32_t c = 1; c < tid; c++) s += cnt9_get(totV, c) * (c < 3 ? 1u : 3u);
            S.vchunk_start[tid] = s;
            A.vcnt[(uint64_t)gseg * 9 + tid] = (uint16_t)cnt9_get(totV, tid);
            if (tid == 8) S.vbytes = s + 3 * cnt9_get(totV, 8);
        }
    }
    __syncthreads();
    {
        uint32_t pl = (preA.last & 0x10u) ? (preA.last & 0xFu) : 15u;
        const uint32_t bit0 = preA.sum & 0x1FFFFu;
        unsigned long long acc = 0; uint32_t nacc = bit0 & 31u, widx = bit0 >> 5; bool first_word = true;
        uint8_t* vbytes = reinterpret_cast<uint8_t*>(S.bits);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint32_t nl = (nlw[e >> 2] >> (8 * (e & 3))) & 0xFFu;
            if (nl == 0xFFu) continue;
            if (pl != 15u) {
                const uint32_t p = S.chunk_start[pl] + cnt9_get(pos, pl);
                cnt9_inc(pos, pl);
                S.sym[p] = (uint8_t)nl;
                atomicAdd(&S.hist[pl * 16 + nl], 1u);
            } else S.first_nl = nl;
            pl = nl;
            if (MODE == 1) {
                if (nl) {
                    acc = (acc << (3 * nl)) | fld[e]; nacc += 3 * nl;
                    if (nacc >= 32) {
                        nacc -= 32;
                        const uint32_t wv = (uint32_t)(acc >> nacc);
                        if (first_word) { atomicOr(&S.bits[widx], wv); first_word = false; } else S.bits[widx] = wv;
                        widx++;
                    }
                }
            } else if (nl) {
                const uint32_t u0 = fld[e] >> 16, u1 = (fld[e] >> 8) & 255u, u2 = fld[e] & 255u;
                const uint32_t vp = S.vchunk_start[nl] + cnt9_get(vpos, nl) * (nl < 3 ? 1u : 3u);
                cnt9_inc(vpos, nl);
                uint32_t* hv = S.hist2 + VAL_OFF[nl];
                if (nl == 1) { const uint32_t v = (u0 << 2) | (u1 << 1) | u2; vbytes[vp] = (uint8_t)v; atomicAdd(hv + v, 1u); }
                else if (nl == 2) { const uint32_t v = (u0 << 4) | (u1 << 2) | u2; vbytes[vp] = (uint8_t)v; atomicAdd(hv + v, 1u); }
                else {
                    vbytes[vp] = (uint8_t)u0; vbytes[vp + 1] = (uint8_t)u1; vbytes[vp + 2] = (uint8_t)u2;
                    atomicAdd(hv + u0, 1u); atomicAdd(hv + u1, 1u); atomicAdd(hv + u2, 1u);
                }
            }
        }
        if (MODE == 1 && nacc && (nacc != (bit0 & 31u) || !first_word))
            atomicOr(&S.bits[widx], (uint32_t)(acc << (32 - nacc)));
    }
    __syncthreads();

    // ---------------- phase 3
    const uint32_t nvalid = totA.sum >> 17, nbits = totA.sum & 0x1FFFFu;
    const uint32_t nsym = nvalid ? nvalid - 1 : 0;
    {
        uint4* dst = reinterpret_cast<uint4*>(A.sym_area + (uint64_t)gseg * SEG);
        const uint4* src = reinterpret_cast<const uint4*>(S.sym);
        for (uint32_t k = tid; k < (nsym + 15) / 16; k += FRONT_THREADS) dst[k] = src[k];
        uint32_t nbytes = MODE == 1 ? ((nbits + 31) / 32) * 4 : 0;
        if (MODE == 2) nbytes = S.vbytes;
        uint4* bd = reinterpret_cast<uint4*>(A.bits_area + (uint64_t)gseg * SEG_BITS_BYTES);
        const uint4* bs = reinterpret_cast<const uint4*>(S.bits);
        for (uint32_t k = tid; k < (nbytes + 15) / 16; k += FRONT_THREADS) bd[k] = bs[k];
    }
    if (tid < 144) { const uint32_t v = S.hist[tid]; if (v) atomicAdd(A.hist + (uint64_t)tile * HSTRIDE + HIST_CTX + tid, v); }
    if (MODE == 1 && PXSZ == 4) { const uint32_t v = S.hist2[tid]; if (v) atomicAdd(A.hist + (uint64_t)tile * HSTRIDE + HIST_ALPHA + tid, v); }
    if (MODE == 2) for (uint32_t k = tid; k < 576; k += FRONT_THREADS) { const uint32_t v = S.hist2[k]; if (v) atomicAdd(A.hist + (uint64_t)tile * HSTRIDE + HIST_VAL + k, v); }
    if (tid == 0) {
        SegInfo si;
#pragma unroll
        for (int c = 0; c < 9; c++) si.cnt[c] = (uint16_t)cnt9_get(totC, c);
        si.nvalid = (uint16_t)nvalid; si.nbits = nbits;
        si.has_valid = nvalid != 0; si.first_nl = nvalid ? (uint8_t)S.first_nl : 0; si.last_nl = (uint8_t)(totA.last & 0xFu);
        si.pad = 0; si.pad2 = 0;
        A.seginfo[gseg] = si;
    }
}

template <int MODE>
#ifndef XPB_FRONT_MINB
#define XPB_FRONT_MINB 4   /* 4 CTAs per SM: 64 registers per thread; measured 1.44x over the unconstrained 128-register build on batches */
#endif
__global__ void __launch_bounds__(FRONT_THREADS, XPB_FRONT_MINB) k_front(FrontArgs A) {
    __shared__ __align__(16) FrontShared S;
    const uint32_t gseg = blockIdx.x, tile = A.seg_tile[gseg];
    const TileDesc t = A.tiles[tile];
    if (MODE == 2 && A.tile_skip && A.tile_skip[tile]) return;
    if (A.rgba_only && t.pxsz != 4 && t.w <= FRONT2_MAXW) return;
    if (MODE == 1 && t.pxsz == 4) front_segment<MODE, 4>(A, S, t, tile, gseg);
    else front_segment<MODE, 3>(A, S, t, tile, gseg);
}

}  // namespace xpb
