// dec_rans_pair.cuh — throughput rANS block decoders for batches (v2: libxpng.c:429-493, v1: :262-301).
//
// The two states of a block share one word pointer, but each state's recurrence only needs to know WHICH word it
// takes when it renormalises: state A (the one that decodes first in a pair of symbols) takes word k, state B takes
// word k + (A renormalised).  So a block is given to a PAIR of lanes (lane h owns state h), 16 blocks per warp: all 32
// lanes carry a useful recurrence, one ballot per round publishes the renormalisation bits, and both candidate words
// of a round are read from a shared-memory ring before the round starts (k is known from the previous round).
// Symbol search is a compare-accumulate over per-lane thresholds held in registers (alphabets of at most 16 symbols;
// larger alphabets: per-block cumulative table + 256-entry coarse index in shared memory).  The renormalisation
// words of the 16 blocks are staged by the WHOLE warp, 32 consecutive ALIGNED words of one block per refill, copied
// global -> shared by cp.async (no register ever waits for global memory; any number of blocks per check); a block's
// byte misalignment is removed by a funnel shift of neighbouring ring words when the chain reads them, off its
// dependent path.  Copies issued at one check are awaited at the next, eight rounds later.
// Blocks are handed out from a work list sorted by symbol count (k_pd_*), so that the 16 blocks of a warp end together.
//
// The warp-per-block kernels of dec_rans_lat.cuh (41 cycles per symbol, but one useful lane in 32) stay the latency
// variant for calls with few blocks; this is the one for batches (about 1.4 warp instructions per symbol).
#pragma once
#include "common.cuh"
#include "dec_m1.cuh"
#include "dec_back.cuh"
#include "dec_rans_lat.cuh"

namespace xpb {

constexpr uint32_t PD_BLK = 16;        // blocks per warp
constexpr uint32_t PD_WARPS = 4;       // independent warps per CTA
constexpr uint32_t PD_RING = 128;      // staged words per block
constexpr uint32_t PD_LOW = 50;        // refill when fewer than this many staged words are ahead of the chain
constexpr uint32_t PD_BUCKETS = 2048;  // work items are bucketed by symbol count >> 10, longest first
enum { PD_S8 = 0, PD_S16 = 1, PD_BIG = 2, PD_WALK = 3, PD_NCLASS = 4 };   // PD_WALK: the coded tiles themselves, by context symbols (dec_walk3.cuh)

struct PdWork {
    uint32_t* hist;     // [PD_NCLASS][PD_BUCKETS]: counts (k_pd_count), then fill cursors (k_pd_fill)
    uint32_t* start;    // [PD_NCLASS][PD_BUCKETS]: first list position of each bucket
    uint32_t* total;    // [PD_NCLASS]: list lengths
    uint32_t* order;    // [PD_NCLASS][cap]: tile | stream << 24
    uint32_t cap;
};

// Which decoder a block of a coded tile needs: a class, -1 nothing to do, -2 run / raw fill.
__device__ __forceinline__ int pd_class(const TileDesc& t, uint32_t mode, const DecTile* d, uint32_t c, uint32_t& n) {
    n = 0;
    if (mode == 1) {
        if (d->m == 0 || d->m == 0xFE) return -1;
        if (c > 9 || (c == 9 && t.pxsz != 4)) return -1;
        const DecBlock b = d->blk[c]; n = b.n;
        if (b.type == 0 || n == 0) return -1;
        if (b.type < 3) return -2;
        return c == 9 ? PD_BIG : PD_S8;
    }
    if (mode == 2) {
        const uint32_t kind = d->m >> 4;
        if (d->m == 0xFE || d->m == 0xFF || d->m == 0 || (kind == 2 && (d->m & 8))) return -1;
        const bool grey = kind == 2;
        if (c > 16 || (grey && c != 0)) return -1;
        const DecBlock b = d->blk[c]; n = b.n;
        if (b.type == 0 || n == 0) return -1;
        if (b.type < 3) return -2;
        const uint32_t N = grey ? 256u : (uint32_t)DEC_M2_NSYM[c];
        return N <= 9 ? PD_S8 : (N <= 16 ? PD_S16 : PD_BIG);
    }
    return -1;
}
__device__ __forceinline__ uint32_t pd_bucket(uint32_t n) { return PD_BUCKETS - 1u - min(n >> 10, PD_BUCKETS - 1u); }

// Work lists of the tiles coded at `mode`: count, scan (one CTA per class), fill.  hist must be zero before k_pd_count.
__global__ void __launch_bounds__(256) k_pd_count(const TileDesc* tiles, const DecImage* imgs, const DecTile* dt, uint32_t ntiles, uint32_t mode, PdWork W) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x, tile = id / 17u, c = id - tile * 17u;
    if (tile >= ntiles) return;
    const TileDesc t = tiles[tile];
    if (imgs[t.img].mode != mode) return;
    uint32_t n; const int cls = pd_class(t, mode, dt + tile, c, n);
    if (cls >= 0) atomicAdd(W.hist + cls * PD_BUCKETS + pd_bucket(n), 1u);
    if (c == 0 && coded_tile(dt + tile)) atomicAdd(W.hist + PD_WALK * PD_BUCKETS + pd_bucket(dt[tile].nsym), 1u);
}
__global__ void __launch_bounds__(1024) k_pd_scan(PdWork W) {
    __shared__ uint32_t wsum[32];
    const uint32_t cls = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* h = W.hist + cls * PD_BUCKETS;
    const uint32_t a = h[2 * tid], b = h[2 * tid + 1];
    uint32_t inc = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    uint32_t pre = 0;
    for (uint32_t k = 0; k < wid; k++) pre += wsum[k];
    const uint32_t ex = pre + inc - (a + b);
    W.start[cls * PD_BUCKETS + 2 * tid] = ex; W.start[cls * PD_BUCKETS + 2 * tid + 1] = ex + a;
    h[2 * tid] = 0; h[2 * tid + 1] = 0;
    if (tid == 1023) W.total[cls] = pre + inc;
}
__global__ void __launch_bounds__(256) k_pd_fill(const TileDesc* tiles, const DecImage* imgs, const DecTile* dt, uint32_t ntiles, uint32_t mode, PdWork W) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x, tile = id / 17u, c = id - tile * 17u;
    if (tile >= ntiles) return;
    const TileDesc t = tiles[tile];
    if (imgs[t.img].mode != mode) return;
    uint32_t n; const int cls = pd_class(t, mode, dt + tile, c, n);
    if (c == 0 && coded_tile(dt + tile)) {
        const uint32_t bk = PD_WALK * PD_BUCKETS + pd_bucket(dt[tile].nsym);
        const uint32_t pos = W.start[bk] + atomicAdd(W.hist + bk, 1u);
        if (pos < W.cap) W.order[(uint64_t)PD_WALK * W.cap + pos] = tile;
    }
    if (cls < 0) return;
    const uint32_t bk = cls * PD_BUCKETS + pd_bucket(n);
    const uint32_t pos = W.start[bk] + atomicAdd(W.hist + bk, 1u);
    if (pos < W.cap) W.order[(uint64_t)cls * W.cap + pos] = tile | (c << 24);
}

// Run (type 1) and raw (type 2) blocks of the tiles coded at `mode`: one warp per (tile, stream), parallel fills.
struct PdFillArgs {
    const TileDesc* tiles; const DecImage* imgs; const DecTile* dt; const uint8_t* in;
    uint8_t* streams; uint8_t* alpha; uint32_t ntiles, mode;
};
__global__ void __launch_bounds__(128) k_pd_fill_blocks(PdFillArgs A) {
    const uint32_t id = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31, tile = id / 17u, c = id - tile * 17u;
    if (tile >= A.ntiles) return;
    const TileDesc t = A.tiles[tile];
    if (A.imgs[t.img].mode != A.mode) return;
    const DecTile* d = A.dt + tile;
    uint32_t n;
    if (pd_class(t, A.mode, d, c, n) != -2) return;
    const DecBlock b = d->blk[c];
    const uint8_t* blob = A.in + d->blob_off; const uint8_t* blk = blob + b.off;
    if (A.mode == 1) {
        uint8_t* out = c == 9 ? A.alpha + t.px_off : A.streams + t.str_off + b.soff;
        const uint32_t v2 = ld32u(blk + 4) >> 24, csz = ld32u(blk) & 0xFFFFFFu;
        const bool ctx = c < 9;
        if (b.type == 1) lat_fill_run(out, n, ctx && v2 > 8u ? 0u : v2, lane);
        else lat_fill_raw(out, n, v2, blk + 8, blk + csz, 0, lane, ctx);
    } else {
        const bool grey = (d->m >> 4) == 2;
        uint8_t* out = A.streams + t.str_off + b.soff;
        const uint32_t N = grey ? 256u : (uint32_t)DEC_M2_NSYM[c];
        const bool ctx = !grey && c < 9;
        if (b.type == 1) { const uint32_t v = ld32u(blk + 4) >> 24; lat_fill_run(out, n, ctx && v > 8u ? 0u : v, lane); }
        else lat_fill_raw(out, n, bitlen32(N - 1), blob + 8, blob + 4 + d->bsz, d->bitpos[c], lane, ctx);
    }
}

struct PairDecArgs {
    const TileDesc* tiles; const DecImage* imgs; const DecTile* dt; const uint8_t* in;
    uint8_t* streams; uint8_t* alpha;
    const uint32_t* order; const uint32_t* total;   // this class's work list and its length
    int* err;
};

// Shared memory per warp of the large-alphabet variant: cumulative table (u16, pitch 260) + coarse index per block.
constexpr uint32_t PD_CUM_PITCH = 260;
constexpr uint32_t PD_BIG_BYTES = PD_BLK * (PD_CUM_PITCH * 2 + 256);

// VER 2: level-1 v2 blocks (symbols n-1 .. 0, words downwards); VER 1: level-2 v1 blocks (forward).
// NTH > 0: alphabets of at most NTH + 1 symbols, thresholds in registers; NTH == 0: up to 256 symbols, tables in shared memory.
template <int VER, int NTH>
__global__ void __launch_bounds__(PD_WARPS * 32) k_dec_rans_pair(PairDecArgs A) {
    constexpr bool BIG = NTH == 0;
    constexpr int NTHR = BIG ? 1 : NTH;
    __shared__ uint32_t s_ring[PD_WARPS][(PD_RING + 2) * PD_BLK];   // aligned words; slots 0 and 1 are mirrored behind the ring: words k + 1, k + 2 never need a wrap
    __shared__ uint4 s_src[PD_WARPS][PD_BLK];                       // word source: aligned base (lo, hi), 8 * misalignment, words
    __shared__ uint32_t s_fs[BIG ? 1 : PD_WARPS][BIG ? 1 : (NTH + 1) * PD_BLK];   // [symbol][block]: start | freq << 16
    extern __shared__ __align__(16) uint8_t s_big[];                // BIG: [PD_WARPS][PD_BIG_BYTES]
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane & 1, blk = lane >> 1;
    const uint32_t total = *A.total;
    const uint32_t item = (blockIdx.x * PD_WARPS + wid) * PD_BLK + blk;
    if (item - blk >= total) return;                                // warp-uniform
    const bool exists = item < total;
    const uint32_t entry = exists ? A.order[item] : 0u;
    const uint32_t tile = entry & 0xFFFFFFu, c = entry >> 24;
    const TileDesc t = A.tiles[tile];
    const DecTile* d = A.dt + tile;
    const DecBlock b = d->blk[c];
    const uint8_t* blob = A.in + d->blob_off; const uint8_t* bp = blob + b.off;
    uint32_t* ring = s_ring[wid];
    uint16_t* cumS = reinterpret_cast<uint16_t*>(s_big + wid * PD_BIG_BYTES) + blk * PD_CUM_PITCH;
    uint8_t* coarse = s_big + wid * PD_BIG_BYTES + PD_BLK * PD_CUM_PITCH * 2 + blk * 256;

    // ---- block geometry
    uint32_t n = exists ? b.n : 0u, N, tab_bit = 0, nwords = 0, sh = 0; int pb;
    const uint8_t* tabp; const uint8_t* tab_end; const uint8_t* wfirst; uint8_t* out;
    uint64_t x0 = 1ull << 31, x1 = 1ull << 31;
    bool ok = exists;
    if (VER == 2) {
        const uint32_t csz = ld32u(bp) & 0xFFFFFFu, w1 = ld32u(bp + 4), w2 = ld32u(bp + 8), tabw = w2 & 0xFFFFFFu;
        N = (w1 >> 24) + 2; pb = (int)(w2 >> 24);
        if (N > (BIG ? 256u : (uint32_t)NTH + 1u) || pb < 10 || pb > 15 || 8 + 4ull * tabw > csz || tabw < 5) ok = false;
        tabp = ok ? bp + 8 + 4ull * tabw : bp; tab_end = ok ? bp + csz : bp;
        const uint8_t* sp = tabp - 16;                               // state0, state1 (libxpng.c:467)
        if (ok) { x0 = ld64u(sp); x1 = ld64u(sp + 8); nwords = (uint32_t)((sp - (bp + 12)) / 4); }
        wfirst = ok ? sp - 4 : bp;
        out = c == 9 ? A.alpha + t.px_off : A.streams + t.str_off + b.soff;
    } else {
        const bool grey = (d->m >> 4) == 2;
        N = grey ? 256u : (uint32_t)DEC_M2_NSYM[c]; pb = grey ? 15 : 14;
        const uint32_t bsize = ld32u(bp) & 0xFFFFFFu;
        tabp = blob + 8; tab_end = blob + 4 + d->bsz; tab_bit = d->bitpos[c];
        if (bsize < 24) ok = false;
        if (ok) { x0 = ld64u(bp + 8); x1 = ld64u(bp + 16); nwords = (bsize - 24) / 4; }
        wfirst = bp + 24;
        out = A.streams + t.str_off + b.soff;
    }
    // ---- frequency table -> thresholds (registers) and [symbol][block] table, or the shared cumulative table
    const bool sparse = b.type == 4;
    uint32_t th[NTHR];
    if (ok) {
        BitR r{ tabp, tab_end, tab_bit };
        uint32_t acc = 0;
        if (!BIG) {
#pragma unroll
            for (int i = 0; i <= NTHR; i++) {
                uint32_t f = 0;
                if ((uint32_t)i < N) f = !sparse ? r.get((uint32_t)pb) : (r.get(1) ? r.get((uint32_t)pb) : 0u);
                if (h == 0) s_fs[BIG ? 0 : wid][i * PD_BLK + blk] = (acc & 0xFFFFu) | (f << 16);
                acc += f;
                if (i < NTHR) th[i] = (uint32_t)(i + 1) < N ? acc : 0xFFFFFFFFu;
            }
        } else {
            for (uint32_t i = h; i < 256; i += 2) cumS[i] = 0;     // (both lanes run the bit scan; the pair shares the stores)
            for (uint32_t i = 0; i < N; i++) {
                if ((i & 1u) == h) cumS[i] = (uint16_t)acc;
                acc += !sparse ? r.get((uint32_t)pb) : (r.get(1) ? r.get((uint32_t)pb) : 0u);
                if (acc > (1u << pb)) acc = (1u << pb) + 1u;          // corrupt table: refused below
            }
            for (uint32_t i = N + h; i <= 256; i += 2) cumS[i] = (uint16_t)min(acc, 0xFFFFu);
            th[0] = 0;
        }
        if (acc != (1u << pb)) ok = false;                          // a valid table sums to 2^PROB_BITS exactly (libxpng.c:316-329)
    }
    if (BIG) {
        __syncwarp();
        if (ok) {   // coarse[k] = symbol containing slot k << (pb - 8); lane h fills the odd / even half
            uint32_t s = 0;
            for (uint32_t k = 0; k < 256; k++) {
                const uint32_t slot = k << (pb - 8);
                while (s < 255 && (uint32_t)cumS[s + 1] <= slot) s++;
                if ((k & 1u) == h) coarse[k] = (uint8_t)s;
            }
        }
    }
    if (exists && !ok) {   // refused block: zeros, and the call fails
        if (h == 0) dec_fail(A.err, DEC_BAD_BLOCK);
        for (uint32_t k = h; k < n; k += 2) out[k] = 0;
        n = 0; nwords = 0;
    }
    {   // word source descriptor (WordSrc of dec_rans_lat.cuh)
        const uintptr_t a = reinterpret_cast<uintptr_t>(wfirst);
        sh = (uint32_t)(a & 3u) * 8u;
        const uintptr_t ab = a & ~(uintptr_t)3;
        if (h == 0) s_src[wid][blk] = make_uint4((uint32_t)ab, (uint32_t)((uint64_t)ab >> 32), sh, nwords);
    }
    __syncwarp();

    // ---- ring staging by the whole warp: slot j of a block holds its aligned word j (VER 1: base[j]; VER 2: base[1 - j],
    // the words run downwards); stream word k is the funnel shift of slots k and k + 1
    uint32_t staged = 0, k = 0;
    const uint32_t nraw = nwords + 1u;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    auto stage = [&](const uint32_t bi) {                          // 32 aligned words of block bi from its `staged` on (all lanes)
        const uint4 src = s_src[wid][bi];
        const uint32_t* base = reinterpret_cast<const uint32_t*>(((uint64_t)src.y << 32) | src.x);
        const uint32_t s0 = __shfl_sync(0xffffffffu, staged, 2 * bi);
        const uint32_t nw = src.w, j = s0 + lane;
        const uint32_t* p;                                          // clamped to the words the block owns (slots past them are never consumed by a valid stream)
        if (VER == 1) { const uint32_t jmax = src.z ? nw : (nw ? nw - 1u : 0u); p = base + min(j, jmax); }
        else p = base + 1 - (int64_t)min(j, nw);
        const uint32_t slot = j & (PD_RING - 1u);
        const uint32_t dst = ring_s + (slot * PD_BLK + bi) * 4u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(p) : "memory");
        if (slot < 2u) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + PD_RING * PD_BLK * 4u), "l"(p) : "memory");
        if (blk == bi) staged += 32;
    };
    for (uint32_t rep = 0; rep < 2; rep++)
        for (uint32_t bi = 0; bi < PD_BLK; bi++) stage(bi);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    // ---- the recurrences
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t xlo = h ? (uint32_t)x1 : (uint32_t)x0, xhi = h ? (uint32_t)(x1 >> 32) : (uint32_t)(x0 >> 32);
    const bool isA = VER == 1 ? h == 0 : h == 1;                    // in full groups: the state that decodes first
    const char* ringb = reinterpret_cast<const char*>(ring + blk);
    const uint32_t* fsb = s_fs[BIG ? 0 : wid] + blk;
    // thresholds two per register (16-bit lanes; 0x8000 = never reached: slots are below 2^15): the comparisons are monotone
    // (cumulative table), so the symbol is the NUMBER of thresholds <= slot: one subtraction per pair, one popcount
    constexpr int NP = (NTHR + 1) / 2;
    uint32_t tp[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const uint32_t lo = th[2 * q] > 0x8000u ? 0x8000u : th[2 * q];
        const uint32_t hi = (2 * q + 1 < NTHR) ? (th[(2 * q + 1 < NTHR) ? 2 * q + 1 : 0] > 0x8000u ? 0x8000u : th[(2 * q + 1 < NTHR) ? 2 * q + 1 : 0]) : 0x8000u;
        tp[q] = lo | (hi << 16);
    }
    const uint32_t otmask = 1u << (lane ^ 1u), pairmask = 3u << (lane & 30u);
    auto round = [&](const bool act, const bool first) -> uint32_t {   // one symbol of my state; first: my state precedes the partner's in this round
        const uint32_t kb = (k & (PD_RING - 1u)) * (PD_BLK * 4u);
        const uint32_t r0 = *reinterpret_cast<const uint32_t*>(ringb + kb), r1 = *reinterpret_cast<const uint32_t*>(ringb + kb + PD_BLK * 4u),
                       r2 = *reinterpret_cast<const uint32_t*>(ringb + kb + PD_BLK * 8u);
        const uint32_t c0 = VER == 1 ? __funnelshift_r(r0, r1, sh) : __funnelshift_r(r1, r0, sh);   // stream words k and k + 1
        const uint32_t c1 = VER == 1 ? __funnelshift_r(r1, r2, sh) : __funnelshift_r(r2, r1, sh);
        const uint32_t slot = xlo & mask;
        uint32_t s = 0, f, bias;
        if (!BIG) {
            const uint32_t s2 = (slot * 0x00010001u) | 0x80008000u;
            uint32_t r = 0, r1 = 0;                                   // two accumulation chains: half the dependent depth
#pragma unroll
            for (int q = 0; q < NP; q++) { const uint32_t bit = ((s2 - tp[q]) >> q) & (0x80008000u >> q); if (q & 1) r1 |= bit; else r |= bit; }
            s = (uint32_t)__popc(r | r1);
            const uint32_t e = fsb[s * PD_BLK];
            f = e >> 16; bias = slot - (e & 0xFFFFu);
        } else {
            s = coarse[slot >> (pb - 8)];
            uint32_t lo = cumS[s], hi = cumS[s + 1];
            while (hi <= slot && s < 255u) { s++; lo = hi; hi = cumS[s + 1]; }
            f = hi - lo; bias = slot - lo;
        }
        const uint32_t qlo = __funnelshift_r(xlo, xhi, pb), qhi = xhi >> pb;
        const uint64_t tt = (uint64_t)f * qlo + bias;
        const uint32_t nlo = (uint32_t)tt, nhi = f * qhi + (uint32_t)(tt >> 32);
        const bool p = act && (nhi | (nlo & 0x80000000u)) == 0;     // x < 2^31: renormalise (libxpng.c:295, :486)
        const uint32_t bal = __ballot_sync(0xffffffffu, p);
        const uint32_t w = (first || !(bal & otmask)) ? c0 : c1;    // the partner's word comes first when it decodes first and renormalised
        if (act) { xlo = p ? w : nlo; xhi = p ? nlo : nhi; }
        k += (uint32_t)__popc(bal & pairmask);
        return s;
    };
    // words are staged ahead of the chain: the copies requested at the previous check (eight rounds ago) are complete and
    // made visible to the warp, then every block with fewer than PD_LOW words ahead gets its next 32
    auto check = [&]() {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        uint32_t needy = __ballot_sync(0xffffffffu, h == 0 && staged < nraw && staged < k + PD_LOW);
        while (needy) {
            const uint32_t bi = (__ffs(needy) - 1) >> 1;
            needy &= needy - 1;
            stage(bi);
        }
    };
    // a group = 8 symbols of the block = 4 rounds; lane h keeps the symbols at offsets h, h + 2, h + 4, h + 6 of the group
    auto emit8 = [&](uint32_t m, uint8_t* dst, bool on) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, m, 1);
        const uint32_t e = h ? o : m, od = h ? m : o;               // even lane's word, odd lane's word
        const uint32_t word = h ? __byte_perm(e, od, 0x7362) : __byte_perm(e, od, 0x5140);
        if (on) *reinterpret_cast<uint32_t*>(dst + 4 * h) = word;
    };
    const uint32_t head = VER == 2 ? (n & 7u) : 0u, G = (n - head) >> 3, tail = VER == 1 ? (n & 7u) : 0u;
    uint32_t Gmax = G;
#pragma unroll
    for (int o = 16; o; o >>= 1) Gmax = max(Gmax, __shfl_xor_sync(0xffffffffu, Gmax, o));
    if (VER == 2) {   // symbols n-1 .. n-head one at a time: symbol i belongs to state i & 1
#pragma unroll 1
        for (uint32_t j = 0; j < 7; j++) {
            const uint32_t i = n - 1u - j;
            const bool act = j < head && (i & 1u) == h;
            const uint32_t s = round(act, true);
            if (act) out[i] = (uint8_t)s;
        }
    }
#pragma unroll 1
    for (uint32_t g = 0; g < Gmax; g += 2) {
        check();
#pragma unroll
        for (uint32_t gg = 0; gg < 2; gg++) {
            const uint32_t gi = g + gg;
            const bool on = gi < G;                                  // a block that has run out of full groups keeps its state (tail symbols follow)
            uint32_t m = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t s = round(on, isA);
                m |= s << (8 * (VER == 1 ? r : 3 - r));
            }
            uint8_t* dst = VER == 1 ? out + 8ull * gi : out + (uint64_t)(n - head) - 8ull - 8ull * gi;
            emit8(m, dst, on);
        }
    }
    if (VER == 1) {   // the last n & 7 symbols one at a time
        check();
#pragma unroll 1
        for (uint32_t j = 0; j < 7; j++) {
            const bool act = j < tail && (j & 1u) == h;
            const uint32_t s = round(act, true);
            if (act) out[8ull * G + j] = (uint8_t)s;
        }
    }
}

}  // namespace xpb
