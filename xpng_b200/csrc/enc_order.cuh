// enc_order.cuh — work lists of the pair-lane rANS encoders (enc_rans_lat.cuh), sorted by symbol count.
//
// A warp of the pair encoders runs until the longest of its 16 blocks is done, and a launch until its longest warp is:
// handing out blocks in tile order wastes most lanes (the streams of a tile differ in length by two orders of
// magnitude) and may start a long chain last.  One CTA per chunk counting-sorts the chunk's blocks by length, longest
// first, into one list per table size (alphabets of at most 16 symbols / larger ones), so the 16 blocks of a warp end
// together and CTAs are dispatched in order of decreasing length.  Order inside a bucket is arbitrary: blocks are
// independent, the bytes do not depend on it.
#pragma once
#include "common.cuh"
#include "enc_m2.cuh"

namespace xpb {

constexpr uint32_t EO_BUCKETS = 4096;    // bucket = symbol count >> 9 (tiles hold at most 443556 pixels, 3 symbols each)

struct EncOrderArgs {
    const TileDesc* tiles; const TileState* state; const uint8_t* tclass;   // tclass: level 2 only
    uint32_t ntiles, mode;
    uint32_t* order;       // [2][cap]: tile | stream << 24
    uint32_t* total;       // [2]
    uint32_t cap;
};

__device__ __forceinline__ int eo_class(const EncOrderArgs& A, uint32_t tile, uint32_t c, uint32_t& n) {
    const TileState* st = A.state + tile;
    n = st->len[c];
    if (A.mode == 1) {
        if (c < 9) return 0;
        return (c == 9 && A.tiles[tile].pxsz == 4) ? 1 : -1;
    }
    if (A.tclass[tile] != TC_RGB) return -1;
    return M2_NSYM[c] <= 16 ? 0 : 1;
}

__global__ void __launch_bounds__(1024) k_enc_order(EncOrderArgs A) {
    __shared__ uint32_t hist[2][EO_BUCKETS];
    __shared__ uint32_t wsum[2][32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t per = A.mode == 1 ? 10u : 17u, items = per * A.ntiles;
    for (uint32_t k = tid; k < 2 * EO_BUCKETS; k += 1024) (&hist[0][0])[k] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < items; i += 1024) {
        const uint32_t tile = i / per, c = i - tile * per;
        uint32_t n; const int cls = eo_class(A, tile, c, n);
        if (cls >= 0) atomicAdd(&hist[cls][EO_BUCKETS - 1u - min(n >> 9, EO_BUCKETS - 1u)], 1u);
    }
    __syncthreads();
    // exclusive scan of both histograms: four buckets per thread, warp scan, warp totals
#pragma unroll
    for (int cls = 0; cls < 2; cls++) {
        uint32_t v[4], s = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { v[q] = hist[cls][4 * tid + q]; s += v[q]; }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) wsum[cls][wid] = inc;
        __syncthreads();
        uint32_t pre = 0;
        for (uint32_t k = 0; k < wid; k++) pre += wsum[cls][k];
        uint32_t ex = pre + inc - s;
#pragma unroll
        for (int q = 0; q < 4; q++) { hist[cls][4 * tid + q] = ex; ex += v[q]; }
        if (tid == 1023) A.total[cls] = ex;
    }
    __syncthreads();
    for (uint32_t i = tid; i < items; i += 1024) {
        const uint32_t tile = i / per, c = i - tile * per;
        uint32_t n; const int cls = eo_class(A, tile, c, n);
        if (cls < 0) continue;
        const uint32_t pos = atomicAdd(&hist[cls][EO_BUCKETS - 1u - min(n >> 9, EO_BUCKETS - 1u)], 1u);
        if (pos < A.cap) A.order[(uint64_t)cls * A.cap + pos] = tile | (c << 24);
    }
}

}  // namespace xpb
