// dec_walk3.cuh — context walk for batches: THREE tiles per warp (nl_{i+1} = next unread symbol of stream nl_i,
// libxpng.c:803).
//
// The walk of a tile is one shuffle per symbol between nine lanes (k_dec_walk_ring, dec_rans_lat.cuh): a warp that
// walks one tile issues every instruction for 9 useful lanes, and ncu shows the batch decoders issue-bound on exactly
// this kernel (11.6 warp instructions per symbol, 66 % issue-slot utilisation).  Here lanes 9g + c own stream c of the
// warp's tile g (g = 0..2): the same instruction stream advances three independent chains, the window pop is
// predicated instead of selected, and a lane publishes its head symbol already offset by 9g, so the shuffle result
// IS the next source lane.  Tiles come from a list sorted by symbol count (class PD_WALK of dec_rans_pair.cuh), so the
// three tiles of a warp finish together.  Rings, top-up and refill follow k_dec_walk_ring (one refill of 32 chunks for
// one stream per half block, loaded during one half and stored at the start of the next; a stream that is about
// to run dry is served on the spot).
#pragma once
#include "common.cuh"
#include "dec_m1.cuh"
#include "dec_rans_lat.cuh"
#include "dec_rans_pair.cuh"

namespace xpb {

constexpr uint32_t W3_CH = 128, W3_LOW = 64, W3_CRIT = 12;   // ring chunks per stream; refill / serve-now thresholds (chunks ahead)
constexpr int W3_STEPS = 72;                                 // steps per block: 9 top-up periods of 8, 8 stores of 9 symbols

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_dec_walk3(WalkArgs A, const uint32_t* __restrict__ order, const uint32_t* __restrict__ total) {
    __shared__ uint32_t wring[WARPS][27][W3_CH];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ntot = *total;
    const uint32_t first = (blockIdx.x * WARPS + wid) * 3u;
    if (first >= ntot) return;                                  // warp-uniform
    const bool lane_on = lane < 27;
    const uint32_t g = lane_on ? lane / 9u : 2u, gb = 9u * g, c = lane_on ? lane - gb : 0u;
    const bool exists = lane_on && first + g < ntot;
    const uint32_t tile = exists ? (order[first + g] & 0xFFFFFFu) : (order[first] & 0xFFFFFFu);
    const TileDesc t = A.tiles[tile];
    const DecTile* d = A.dt + tile;
    uint8_t* out = A.nlseq + t.px_off;
    asm volatile("" : "+l"(out));                            // held in registers: rebuilt from the parameter bank at every store group, the add waited for the constant load (13 % of the stall samples, ncu r03c)
    const uint32_t m = exists ? d->nsym : 0u;
    uint32_t mmax = m;
#pragma unroll
    for (int o = 16; o; o >>= 1) mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    const uint32_t nch = exists ? (d->blk[c].n + 7) / 8 : 0u;    // 8-symbol chunks of my stream
    const uint2* gsrc = reinterpret_cast<const uint2*>(A.streams + t.str_off + d->blk[c].soff);
    uint32_t* myring = wring[wid][lane_on ? lane : 0];
    uint32_t st = 0, ch = 0;
    // one refill of the stream owned by lane p: 32 chunks from chunk index st[p] (all lanes)
    auto fetch = [&](uint32_t p, uint2& r) {
        const uint2* base = reinterpret_cast<const uint2*>(__shfl_sync(0xffffffffu, (unsigned long long)gsrc, p));
        const uint32_t s0 = __shfl_sync(0xffffffffu, st, p), np = __shfl_sync(0xffffffffu, nch, p);
        const uint32_t k = s0 + lane;
        r = k < np ? __ldg(base + k) : make_uint2(0u, 0u);
    };
    auto store = [&](uint32_t p, const uint2 r) {
        const uint32_t s0 = __shfl_sync(0xffffffffu, st, p);
        wring[wid][p][(s0 + lane) & (W3_CH - 1)] = walk_pack8(r);
        if (lane == p) st += 32;
    };
    auto needy = [&](uint32_t ahead) -> uint32_t { return __ballot_sync(0xffffffffu, lane_on && st < nch && st - ch <= ahead); };
    for (uint32_t nb = needy(W3_LOW); nb; nb = needy(W3_LOW)) {   // initial fill
        const uint32_t p = __ffs(nb) - 1;
        uint2 r; fetch(p, r); store(p, r);
    }
    __syncwarp();
    auto chunk = [&](uint32_t k) -> uint32_t { const uint32_t v = myring[k & (W3_CH - 1)]; return k < nch ? v : 0u; };
    uint32_t wlo = chunk(0), whi = chunk(1), cnt = 16, nbuf = chunk(2);
    ch = 3;
    uint32_t info = (wlo & 0xFu) + gb, cur = gb, nm = 0, nx = 0;
    uint32_t mk[9];
#pragma unroll
    for (int j = 0; j < 9; j++) { mk[j] = (lane_on && c == (uint32_t)j) ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    uint32_t pend0 = 32, pend1 = 32;                            // streams whose refills are in flight (32 = none)
    uint2 preg0 = make_uint2(0u, 0u), preg1 = make_uint2(0u, 0u);
    for (uint32_t pos = 0; pos < mmax; pos += W3_STEPS) {
        // between blocks (warp-uniform): land the refills loaded during the previous block, serve streams about to run dry,
        // start the next two refills
        if (pend0 < 32) store(pend0, preg0);
        if (pend1 < 32) store(pend1, preg1);
        __syncwarp();
        for (uint32_t nb = needy(W3_CRIT); nb; nb = needy(W3_CRIT)) {
            const uint32_t p = __ffs(nb) - 1;
            uint2 r; fetch(p, r); store(p, r);
            __syncwarp();
        }
        {
            uint32_t nb = needy(W3_LOW);
            pend0 = nb ? __ffs(nb) - 1 : 32u;
            if (nb) { fetch(pend0, preg0); nb &= nb - 1; }
            pend1 = nb ? __ffs(nb) - 1 : 32u;
            if (nb) fetch(pend1, preg1);
        }
        uint32_t keep = 0;
#pragma unroll
        for (int s = 0; s < W3_STEPS; s++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);   // the chains: one shuffle per symbol (and tile)
            if (lane == cur) {                                          // the owner pops while the shuffle is in flight
                wlo = __funnelshift_r(wlo, whi, 4); whi >>= 4; cnt--;
            }
            info = (wlo & 0xFu) + gb;
            keep = (s % 9 == 0) ? (got & mk[0]) : (keep | (got & mk[s % 9]));
            cur = got;
            if ((s & 7) == 1) {
                // every 8 steps, branch-free: windows at <= 8 symbols append the next 8.  cnt >= 1 holds at every check (16 at
                // start; a refilled window has >= 9 and at most 8 are popped until the next check), so `info` is never stale.
                nm = cnt <= 8u ? 0xFFFFFFFFu : 0u;
                const unsigned long long add = (unsigned long long)(nbuf & nm) << (4u * min(cnt, 8u));
                wlo |= (uint32_t)add; whi |= (uint32_t)(add >> 32);
                cnt += nm & 8u;
                nx = chunk(ch);                                         // consumed four steps later
                ch += nm & 1u;
            }
            if ((s & 7) == 5) nbuf = (nx & nm) | (nbuf & ~nm);
            if (s % 9 == 8) {                                           // lane c of a group holds symbol c of these nine
                const uint32_t idx = pos + (uint32_t)(s - 8) + c;
                if (lane_on && idx < m) out[idx] = (uint8_t)(keep - gb);
            }
        }
    }
}

}  // namespace xpb
