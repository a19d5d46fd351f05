// api.cu — host orchestration and the C ABI (include/xpng_b200.h) of the B200 xPNG codec.
// Everything here is plumbing: tile tables (libxpng.c:51-83), scratch layout, kernel launches, host<->device
// copies.  All codec arithmetic lives in the kernels.
//
// Execution model.  A context owns LANES: a lane is a set of CUDA streams plus its own scratch buffers.  A call
// cuts its batch into chunks of whole images and issues every chunk on its own lane WITHOUT waiting for it; only at
// the end does the host collect the per-chunk size tables (encode) or error flags (decode).  The serial chains of the
// bit stream (rANS recurrences, context walk) leave most of the machine idle while they run; with several chunks
// in flight the data-parallel kernels of one chunk fill the SMs under the chains of another, and with host buffers the
// H2D / D2H copies of one chunk overlap the kernels of its neighbours.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/xpng_b200.h"
#include "common.cuh"
#include "enc_front.cuh"
#include "enc_front2.cuh"
#include "enc_back.cuh"
#include "enc_m2.cuh"
#include "enc_rans_lat.cuh"
#include "enc_order.cuh"
#include "dec_m1.cuh"
#include "dec_back.cuh"
#include "dec_rans_lat.cuh"
#include "dec_rans_pair.cuh"
#include "dec_walk3.cuh"
#include "misc.cuh"
#include "orient.cuh"

using namespace xpb;

struct DevBuf { void* p = nullptr; size_t cap = 0; };
struct PinBuf { void* p = nullptr; size_t cap = 0; };

struct xpngb_ctx {
    int device = 0;
    xpngb_ctx* root = nullptr;                 // the context the caller holds: error text, launch counter, profile table, tunables
    std::vector<xpngb_ctx*> lanes;             // root only: further lanes (lane 0 is the root itself), created on demand
    static constexpr int NSIDE = 5;
    cudaStream_t stream = nullptr, side[NSIDE] = {}, cur = nullptr;   // main stream, side streams for independent chains, stream of the next launch
    cudaStream_t hi = nullptr;                 // high-priority twin of the main stream: serial-chain kernels (LAUNCH_HI)
    cudaEvent_t ev_hi = nullptr;
    bool enc_sorted = true;                    // root: XPNGB_ENC_SORT=0: pair encoders take blocks in tile order (A/B)
    bool use_prio = true;                      // root: XPNGB_PRIO=0 switches the priorities off (A/B)
    cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, pe0 = nullptr, pe1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_stage = nullptr;            // decode: this lane's chunk has finished its context-decode (or walk) phase (staggered waves)
    uint32_t dec_stagger1 = 0, dec_stagger2 = 0, dec_stagger_at = 0;   // root: XPNGB_DEC_STAGGER1/2 = S: chunk k of a level-1/2 batch decode starts when chunk k - S has passed its stage (0: all at once); XPNGB_DEC_STAGGER_AT: 0 contexts decoded, 1 walked
    int profile = 0;   // 1: per-kernel CUDA-event timing accumulated (serialises the launches); 2: also print to stderr;
                       // 3: timeline — start/end of every launch relative to the call's start, launches NOT serialised (stderr)
    int lane_id = 0;
    struct TlRow { const char* name; int lane; cudaEvent_t a, b; };
    std::vector<TlRow> tl;
    struct ProfRow { const char* name; double ms; uint32_t count; };
    std::vector<ProfRow> prof;
    char err[512] = { 0 };
    float last_ms = 0.f;
    uint32_t launches = 0;
    uint64_t max_chunk_px = 1ull << 30;       // XPNGB_CHUNK_MPIX: upper bound of a chunk (bounds the scratch of a lane)
    uint32_t pipe_lanes = 8;                  // XPNGB_PIPE_LANES: chunks in flight per call
    uint64_t pipe_min_px = 32ull << 20;       // XPNGB_PIPE_MIN_MPIX: no chunk smaller than this
    bool front_v1 = false;                    // XPNGB_FRONT=1: first-generation front end for RGB tiles too (A/B, tests)
    bool walk_global = false, walk_ring = false;   // XPNGB_WALK=global | ring: force a walk variant (tests, A/B)
    uint32_t direct_max_tiles = 148;  // level-2 decode: tiles per call up to which the 64 KiB direct tables are used (3 chains per SM stay resident)
    uint32_t v2_direct_max_tiles = ~0u;  // level-1 decode: same trade for the 16 KiB context tables
    uint32_t unr_multi_max_tiles = 592, unr_force = 0;   // un-predict: tiles per call up to which a tile gets 8 warps; XPNGB_UNR_NW forces a variant (A/B)
    uint32_t s16_lat_max_tiles = 1776;  // level-2 decode, pair family: calls up to this many tiles decode the 16-symbol value stream (the longest chain) warp-per-block
    uint32_t lat_max_blocks = 12000;  // decode: entropy blocks per call up to which the warp-per-block (latency) kernels are used (measured crossover, profiles/)
    uint32_t enc_lat_max_blocks = ~0u; // encode: the pair-lane encoders serve every batch size (chunks in flight keep their launches resident)
    // device scratch
    DevBuf pixels, norm, files, arena, tiles, imgs, seg_tile, costs, hist, seginfo, place, vplace, vcnt, sym_area,
        bits_area, alpha, streams, blocks, state, outs, flags, skip, dimgs, dtiles, plane, nlseq, rowcnt, rowbits, eord,
        rows, edge, errflag, hdr, offs, m2a, m2b, tclass, tabs, ccnt, cbit, resv, oriented, odesc, pdw;
    PinBuf pin_a, pin_b, pin_o;   // pin_o: orientation descriptors (their upload may still be in flight when an encode reuses pin_a)
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->root->err, sizeof ctx->root->err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)
#define FAIL(...)                                                            \
    do {                                                                     \
        snprintf(ctx->root->err, sizeof ctx->root->err, __VA_ARGS__);        \
        return 1;                                                            \
    } while (0)
#define LAUNCH(kernel, grid, block, smem, ...)                                                      \
    do {                                                                                            \
        xpngb_ctx* r_ = ctx->root;                                                                  \
        if (r_->profile == 3) tl_mark(ctx, #kernel, 0);                                             \
        else if (r_->profile) cudaEventRecord(ctx->pe0, ctx->cur);                                  \
        kernel<<<grid, block, smem, ctx->cur>>>(__VA_ARGS__);                                       \
        r_->launches++;                                                                             \
        CK(cudaGetLastError());                                                                     \
        if (r_->profile == 3) tl_mark(ctx, #kernel, 1);                                             \
        else if (r_->profile) {                                                                          \
            float ms_ = 0; cudaEventRecord(ctx->pe1, ctx->cur); cudaEventSynchronize(ctx->pe1);     \
            cudaEventElapsedTime(&ms_, ctx->pe0, ctx->pe1);                                         \
            prof_add(r_, #kernel, ms_);                                                             \
            if (r_->profile > 1) fprintf(stderr, "[xpngb] %-28s %9.3f ms\n", #kernel, ms_);       \
        }                                                                                           \
    } while (0)

// Independent serial chains (e.g. the value-stream blocks of level 2 while the context streams are walked)
// run on the side stream: FORK makes it wait for everything launched so far, JOIN makes the main stream wait for it.
// Chain kernels (rANS recurrences, context walks, the tiny serial scans between them) keep a few warps busy for a long
// time; the data-parallel kernels of the other chunks in flight fill the machine.  The block scheduler serves streams
// of higher priority first, so chains go to high-priority streams: their CTAs become resident as soon as they are
// launched instead of queueing behind hundreds of thousands of front-end CTAs.
// level 0: the longest chains of a call (side stream 0: the 16-symbol value streams of level 2), level 1: other chains.
static cudaStream_t prio_stream(const xpngb_ctx* ctx, int level) {
    cudaStream_t s = nullptr;
    int lo = 0, hi = 0;   // numerically lower = served first
    if (ctx->root->use_prio && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess && hi != lo)
        cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, (hi + level < lo) ? hi + level : hi);
    else cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    return s;
}
static cudaStream_t side_of(xpngb_ctx* ctx, int k) {
    if (!ctx->side[k]) ctx->side[k] = prio_stream(ctx, k == 0 ? 0 : 1);   // side streams only ever carry chains
    return ctx->side[k];
}
#define LAUNCH_HI(kernel, grid, block, smem, ...)                                                   \
    do {                                                                                            \
        cudaStream_t base_ = ctx->cur;                                                              \
        const bool tw_ = base_ == ctx->stream && ctx->root->use_prio && (ctx->root->profile == 0 || ctx->root->profile == 3);       \
        if (tw_) {                                                                                  \
            if (!ctx->hi) ctx->hi = prio_stream(ctx, 1);                                               \
            CK(cudaEventRecord(ctx->ev_hi, base_)); CK(cudaStreamWaitEvent(ctx->hi, ctx->ev_hi, 0)); ctx->cur = ctx->hi; \
        }                                                                                           \
        LAUNCH(kernel, grid, block, smem, __VA_ARGS__);                                             \
        if (tw_) { CK(cudaEventRecord(ctx->ev_hi, ctx->hi)); CK(cudaStreamWaitEvent(base_, ctx->ev_hi, 0)); ctx->cur = base_; } \
    } while (0)
#define FORK_SIDE(k) FORK_FROM(ctx->stream, k)
#define FORK_FROM(src, k) do { CK(cudaEventRecord(ctx->ev_fork, (src))); CK(cudaStreamWaitEvent(side_of(ctx, k), ctx->ev_fork, 0)); ctx->cur = ctx->side[k]; } while (0)
#define BACK_TO_MAIN() do { ctx->cur = ctx->stream; } while (0)
#define JOIN_SIDE(k) do { ctx->cur = ctx->stream; CK(cudaEventRecord(ctx->ev_join[k], ctx->side[k])); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0)); } while (0)

// Timeline mode: two events around a launch on its own stream (no host wait); tl_dump prints them after the call.
static void tl_mark(xpngb_ctx* ctx, const char* name, int end) {
    xpngb_ctx* r = ctx->root;
    if (!end) {
        xpngb_ctx::TlRow row{ name, ctx->lane_id, nullptr, nullptr };
        cudaEventCreate(&row.a); cudaEventCreate(&row.b);
        cudaEventRecord(row.a, ctx->cur);
        r->tl.push_back(row);
    } else cudaEventRecord(r->tl.back().b, ctx->cur);
}
static void tl_dump(xpngb_ctx* ctx, const char* what) {
    if (ctx->profile != 3) return;
    cudaDeviceSynchronize();
    for (auto& row : ctx->tl) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev0, row.a); cudaEventElapsedTime(&b, ctx->ev0, row.b);
        fprintf(stderr, "TL %s %d %s %.3f %.3f\n", what, row.lane, row.name, a, b);
        cudaEventDestroy(row.a); cudaEventDestroy(row.b);
    }
    ctx->tl.clear();
}
static void prof_add(xpngb_ctx* ctx, const char* name, float ms) {
    for (auto& r : ctx->prof) if (!strcmp(r.name, name)) { r.ms += ms; r.count++; return; }
    ctx->prof.push_back({ name, ms, 1 });
}

static int ensure(xpngb_ctx* ctx, DevBuf& b, size_t n) {
    n += 64;   // tail slack for vector over-reads
    if (b.cap >= n) return 0;
    if (b.p) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = n + n / 8;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
static int ensure_pin(xpngb_ctx* ctx, PinBuf& b, size_t n) {
    if (b.cap >= n) return 0;
    if (b.p) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFreeHost(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = n + n / 8 + 4096;
    CK(cudaMallocHost(&b.p, want));
    b.cap = want;
    return 0;
}
#define ENSURE(buf, n) do { if (ensure(ctx, ctx->buf, (n))) return 1; } while (0)

// ------------------------------------------------------------------------------------------------
// Tile grid (libxpng.c:51-83): base tile 444x444 (or full width/height when the image is thinner),
// the remainder of each axis goes to the first tile, or to the first two when it exceeds half a tile.
// ------------------------------------------------------------------------------------------------
static uint64_t axis_cut(uint64_t extent, uint64_t base, uint64_t* first, uint64_t* second) {
    uint64_t n = extent / base, rem = extent % base;
    *first = base + rem; *second = base;
    if (rem > base / 2) { n++; *second = (base + rem) / 2; *first = *second + ((base + rem) & 1); }
    return n;
}

struct Plan {
    std::vector<ImageDesc> imgs;
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> seg_tile;
    uint64_t px_total = 0, row_total = 0, str_total = 0, blk_total = 0;
    bool any_rgba = false;
    bool any_wide = false;     // a tile wider than the staged front end takes (one-tile images up to 197136 x 1)
};

// Appends the tiles of one image.  base = absolute device address of its pixels.
static void plan_image(Plan& P, uint64_t base, uint64_t W, uint64_t H, uint32_t pxsz, uint32_t mode, int sfac, int bfac) {
    ImageDesc I{};
    I.px_off = base; I.raw_size = W * H * pxsz; I.w = (uint32_t)W; I.h = (uint32_t)H; I.pxsz = pxsz;
    I.tile0 = (uint32_t)P.tiles.size(); I.mode = mode;
    uint64_t nw = 1, nh = 1, w0 = W, w1 = W, wb = W, h0 = H, h1 = H, hb = H;
    if (W * H > TILE_AREA) {
        if (W < 444) { wb = W; hb = TILE_AREA / W; }
        else if (H < 444) { hb = H; wb = TILE_AREA / H; }
        else wb = hb = 444;
        nw = axis_cut(W, wb, &w0, &w1);
        nh = axis_cut(H, hb, &h0, &h1);
    }
    const uint32_t img = (uint32_t)P.imgs.size();
    uint64_t y = 0; uint32_t tix = 0;
    for (uint64_t i = 0; i < nh; i++) {
        const uint64_t th = i == 0 ? h0 : (i == 1 ? h1 : hb);
        uint64_t x = 0;
        for (uint64_t j = 0; j < nw; j++) {
            const uint64_t tw = j == 0 ? w0 : (j == 1 ? w1 : wb);
            TileDesc t{};
            t.src_off = base + (y * W + x) * pxsz;
            t.px_off = P.px_total; t.row_off = P.row_total;
            t.w = (uint32_t)tw; t.h = (uint32_t)th; t.bpr = (uint32_t)(W * pxsz); t.npx = (uint32_t)(tw * th);
            t.img = img; t.seg0 = (uint32_t)P.seg_tile.size(); t.nseg = (t.npx + SEG - 1) / SEG; t.pxsz = pxsz;
            t.x0 = (uint32_t)x; t.y0 = (uint32_t)y; t.tix = tix++;
            const uint64_t slice = ((uint64_t)t.npx + 15) / 16 * 16 + 256;
            t.str_off = P.str_total; t.blk_off = P.blk_total;
            P.px_total += slice; P.row_total += th;
            P.str_total += slice * sfac; P.blk_total += slice * bfac + 8192;
            for (uint32_t s = 0; s < t.nseg; s++) P.seg_tile.push_back((uint32_t)P.tiles.size());
            if (t.w > FRONT2_MAXW) P.any_wide = true;
            P.tiles.push_back(t);
            x += tw;
        }
        y += th;
    }
    I.ntiles = (uint32_t)P.tiles.size() - I.tile0;
    if (pxsz == 4) P.any_rgba = true;
    P.imgs.push_back(I);
}

static int upload_plan(xpngb_ctx* ctx, const Plan& P) {
    const size_t nb_t = P.tiles.size() * sizeof(TileDesc), nb_i = P.imgs.size() * sizeof(ImageDesc), nb_s = P.seg_tile.size() * 4;
    ENSURE(tiles, nb_t); ENSURE(imgs, nb_i); ENSURE(seg_tile, nb_s);
    if (ensure_pin(ctx, ctx->pin_a, nb_t + nb_i + nb_s)) return 1;
    uint8_t* h = (uint8_t*)ctx->pin_a.p;
    memcpy(h, P.tiles.data(), nb_t); memcpy(h + nb_t, P.imgs.data(), nb_i); memcpy(h + nb_t + nb_i, P.seg_tile.data(), nb_s);
    CK(cudaMemcpyAsync(ctx->tiles.p, h, nb_t, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->imgs.p, h + nb_t, nb_i, cudaMemcpyHostToDevice, ctx->stream));
    if (nb_s) CK(cudaMemcpyAsync(ctx->seg_tile.p, h + nb_t + nb_i, nb_s, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Level 2 host side
// ------------------------------------------------------------------------------------------------
// Shared memory per entropy-block CTA = cum[] + tables + word ring.  Tables are sized per launch so that every
// chain of a frame is resident at once (one level: 4 B x 2^PROB_BITS; two levels: 1 KiB + 2^PROB_BITS bytes).
constexpr uint32_t lat_smem(uint32_t lut_bytes) { return LAT_CUM_WORDS * 4 + lut_bytes + LAT_RING_WORDS * 4; }
constexpr uint32_t LUT_ONE_12 = 4u << 12, LUT_ONE_14 = 4u << 14, LUT_TWO_12 = 1024 + (1u << 12), LUT_TWO_14 = 1024 + (1u << 14), LUT_TWO_15 = 1024 + (1u << 15);

static void m2_set_attributes() {
    auto k_big = k_rans_v1<256, 32>;
    cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 * 16);
}

// classification, RGB front end (contexts + value streams), grey candidates, backward rANS, size decisions
static int m2_encode_tiles(xpngb_ctx* ctx, const Plan& P, uint32_t ntiles, uint32_t nseg, bool lat) {
    const TileDesc* d_tiles = (const TileDesc*)ctx->tiles.p;
    const uint32_t* d_seg_tile = (const uint32_t*)ctx->seg_tile.p;
    ENSURE(tclass, ntiles); ENSURE(skip, ntiles); ENSURE(tabs, (size_t)ntiles * 17 * TAB_WORDS * 4);
    ENSURE(vplace, nseg * sizeof(SegPlace)); ENSURE(vcnt, (size_t)nseg * 9 * 2);
    (void)P;
    LAUNCH(k_m2_classify, ntiles, 256, 0, d_tiles, (const ImageDesc*)ctx->imgs.p, (uint8_t*)ctx->tclass.p);
    LAUNCH(k_m2_skipmask, (ntiles + 255) / 256, 256, 0, (const uint8_t*)ctx->tclass.p, (uint8_t*)ctx->skip.p, ntiles);
    FrontArgs fa{ d_tiles, d_seg_tile, nullptr, (const uint32_t*)ctx->costs.p, (const uint8_t*)ctx->skip.p, (SegInfo*)ctx->seginfo.p,
                  (uint8_t*)ctx->sym_area.p, (uint8_t*)ctx->bits_area.p, nullptr, (uint32_t*)ctx->hist.p, (uint16_t*)ctx->vcnt.p, 0u };
    if (ctx->root->front_v1) LAUNCH(k_front<2>, nseg, FRONT_THREADS, 0, fa);
    else {                                                     // level 2 codes RGB tiles only
        LAUNCH(k_front2<2>, nseg, FRONT_THREADS, 0, fa);
        fa.rgba_only = 1;
        if (P.any_wide) LAUNCH(k_front<2>, nseg, FRONT_THREADS, 0, fa);
    }
    TileScanArgs ta{ d_tiles, (const SegInfo*)ctx->seginfo.p, (const uint32_t*)ctx->costs.p, (const uint16_t*)ctx->vcnt.p,
                     (SegPlace*)ctx->place.p, (SegPlace*)ctx->vplace.p, (uint32_t*)ctx->hist.p, (TileState*)ctx->state.p,
                     (const uint8_t*)ctx->skip.p, ntiles };
    LAUNCH_HI(k_tile_scan<2>, (ntiles + 3) / 4, 128, 0, ta);
    CompactArgs ca{ d_tiles, d_seg_tile, (const SegInfo*)ctx->seginfo.p, (const SegPlace*)ctx->place.p, (const SegPlace*)ctx->vplace.p,
                    (const uint16_t*)ctx->vcnt.p, (const TileState*)ctx->state.p, (const uint8_t*)ctx->sym_area.p,
                    (const uint8_t*)ctx->bits_area.p, (uint8_t*)ctx->streams.p, (const uint8_t*)ctx->skip.p };
    LAUNCH(k_compact<2>, nseg, 256, 0, ca);
    LAUNCH(k_m2_grey_front, nseg, 256, 0, d_tiles, d_seg_tile, (const uint8_t*)ctx->tclass.p, (uint8_t*)ctx->streams.p, (uint32_t*)ctx->hist.p);
    RansV1Args ra{ d_tiles, (TileState*)ctx->state.p, (const uint32_t*)ctx->hist.p, (const uint8_t*)ctx->tclass.p,
                   (const uint8_t*)ctx->streams.p, (uint8_t*)ctx->blocks.p, (uint32_t*)ctx->tabs.p, ntiles, 0, 17, 0, 0 };
    if (lat) {
        auto k_rans_v1_pair_16 = k_rans_v1_pair<16>; auto k_rans_v1_pair_256 = k_rans_v1_pair<256>;
        const uint32_t ecap = 17u * ntiles;
        if (ctx->root->enc_sorted) {                   // blocks by decreasing length, one list per table size (enc_order.cuh)
            ENSURE(eord, (size_t)(2 * ecap + 2) * 4);
            uint32_t* eo = (uint32_t*)ctx->eord.p;
            LAUNCH_HI(k_enc_order, 1, 1024, 0, EncOrderArgs{ d_tiles, (const TileState*)ctx->state.p, (const uint8_t*)ctx->tclass.p, ntiles, 2u, eo + 2, eo, ecap });
            ra.order = eo + 2; ra.total = eo;
        }
        // alphabets above 16 symbols and the grey candidates go to a side stream that forks BEFORE the main launch but is
        // fed AFTER it: the small-alphabet kernel holds the longest chains, and in a batch its CTAs must be placed first
        // (the 98 KiB CTAs of the other kernel would otherwise take the shared memory and delay them)
        FORK_SIDE(0);
        BACK_TO_MAIN();
        LAUNCH_HI(k_rans_v1_pair_16, (17 * ntiles + PAIR_BLK - 1) / PAIR_BLK, 32, 17 * PAIR_BLK * 24, ra);
        ctx->cur = ctx->side[0];
        RansV1Args rb = ra; rb.c0 = 9; rb.nc = 8; rb.nmin = 16;
        if (ra.order) { rb.order = ra.order + ecap; rb.total = ra.total + 1; }
        LAUNCH_HI(k_rans_v1_pair_256, (8 * ntiles + PAIR_BLK - 1) / PAIR_BLK, 32, 257 * PAIR_BLK * 24, rb);
        rb.c0 = 0; rb.nc = 4; rb.nmin = 0; rb.grey = 1; rb.order = nullptr; rb.total = nullptr;
        LAUNCH_HI(k_rans_v1_pair_256, (4 * ntiles + PAIR_BLK - 1) / PAIR_BLK, 32, 257 * PAIR_BLK * 24, rb);
        JOIN_SIDE(0);
    } else {
    auto k_rans_v1_lane_16 = k_rans_v1<16, 128>; auto k_rans_v1_lane_256 = k_rans_v1<256, 32>;
    LAUNCH(k_rans_v1_lane_16, (17 * ntiles + 127) / 128, 128, 16 * 128 * 16, ra);
    ra.c0 = 9; ra.nc = 8; ra.nmin = 16;
    LAUNCH(k_rans_v1_lane_256, (8 * ntiles + 31) / 32, 32, 256 * 32 * 16, ra);
    ra.c0 = 0; ra.nc = 4; ra.nmin = 0; ra.grey = 1;
    LAUNCH(k_rans_v1_lane_256, (4 * ntiles + 31) / 32, 32, 256 * 32 * 16, ra);
    }
    LAUNCH_HI(k_m2_finish, (ntiles + 127) / 128, 128, 0, d_tiles, (const uint8_t*)ctx->tclass.p, (TileState*)ctx->state.p, ntiles);
    return 0;
}
// CTAs per tile for the assembly kernels: a single frame has fewer tiles than SMs, so each tile's copies are sliced
static unsigned assemble_split(uint32_t ntiles) { const uint32_t y = 592u / (ntiles ? ntiles : 1u); return y < 1u ? 1u : (y > 8u ? 8u : y); }
static int m2_assemble(xpngb_ctx* ctx, const AssembleArgs& aa, uint32_t ntiles) {
    AssembleM2Args ma{ aa, (const uint8_t*)ctx->tclass.p, (const uint32_t*)ctx->tabs.p, (const uint8_t*)ctx->streams.p };
    LAUNCH(k_assemble_m2, dim3(ntiles, assemble_split(ntiles)), 256, 0, ma);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Context and lanes
// ------------------------------------------------------------------------------------------------
static void lane_free(xpngb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < xpngb_ctx::NSIDE; k++) if (ctx->side[k]) cudaStreamSynchronize(ctx->side[k]);
    DevBuf* all[] = { &ctx->pixels, &ctx->norm, &ctx->files, &ctx->arena, &ctx->tiles, &ctx->imgs, &ctx->seg_tile, &ctx->costs,
                      &ctx->hist, &ctx->seginfo, &ctx->place, &ctx->vplace, &ctx->vcnt, &ctx->sym_area, &ctx->bits_area, &ctx->alpha,
                      &ctx->streams, &ctx->blocks, &ctx->state, &ctx->outs, &ctx->flags, &ctx->skip, &ctx->dimgs, &ctx->dtiles,
                      &ctx->plane, &ctx->nlseq, &ctx->rowcnt, &ctx->rowbits, &ctx->rows, &ctx->edge, &ctx->errflag, &ctx->hdr,
                      &ctx->offs, &ctx->m2a, &ctx->m2b, &ctx->tclass, &ctx->tabs, &ctx->ccnt, &ctx->cbit, &ctx->resv, &ctx->oriented,
                      &ctx->odesc, &ctx->pdw, &ctx->eord };
    for (DevBuf* b : all) if (b->p) cudaFree(b->p);
    if (ctx->pin_a.p) cudaFreeHost(ctx->pin_a.p);
    if (ctx->pin_b.p) cudaFreeHost(ctx->pin_b.p);
    if (ctx->pin_o.p) cudaFreeHost(ctx->pin_o.p);
    if (ctx->hi) { cudaStreamSynchronize(ctx->hi); cudaStreamDestroy(ctx->hi); }
    cudaEvent_t evs[] = { ctx->ev0, ctx->ev1, ctx->ev_fork, ctx->pe0, ctx->pe1, ctx->ev_done, ctx->ev_hi, ctx->ev_stage };
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (int k = 0; k < xpngb_ctx::NSIDE; k++) { if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]); if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// A lane: streams, events and (empty, grow-only) scratch.  root == nullptr creates the root lane.
static xpngb_ctx* lane_new(int device, xpngb_ctx* root) {
    xpngb_ctx* ctx = new xpngb_ctx();
    ctx->device = device; ctx->root = root ? root : ctx;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_hi, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
              cudaEventCreate(&ctx->pe0) == cudaSuccess && cudaEventCreate(&ctx->pe1) == cudaSuccess;
    for (int k = 0; ok && k < xpngb_ctx::NSIDE; k++)   // side streams are created on first use (side_of): the device has at most
        ok = cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming) == cudaSuccess;   // 32 hardware queues, unused streams would alias used ones
    if (!ok) { lane_free(ctx); return nullptr; }
    ctx->cur = ctx->stream;
    return ctx;
}
// Lane i of a context (0 = the root itself).
static xpngb_ctx* lane_get(xpngb_ctx* root, uint32_t i) {
    if (i == 0) return root;
    while (root->lanes.size() < i) {
        xpngb_ctx* l = lane_new(root->device, root);
        if (!l) return nullptr;
        root->lanes.push_back(l);
        l->lane_id = (int)root->lanes.size();
    }
    return root->lanes[i - 1];
}

extern "C" int xpngb_create(xpngb_ctx** out, int device) {
    if (!out) return 1;
    *out = nullptr;
    // A batch call keeps up to pipe_lanes x 3 streams busy: hosts that code batches export CUDA_DEVICE_MAX_CONNECTIONS=32
    // before their first CUDA call (the driver's default of 8 hardware queues makes the streams of a call
    // wait on each other).  It is not forced here: 32 queues add ~2.5 s to every process start, which a one-image CLI run
    // should not pay (tools/ctx_time.py).
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return 1;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    xpngb_ctx* ctx = lane_new(device, nullptr);
    if (!ctx) return 1;
    if (const char* e = getenv("XPNGB_PROFILE")) ctx->profile = atoi(e) == 3 ? 3 : (atoi(e) ? 2 : 0);
    if (const char* e = getenv("XPNGB_CHUNK_MPIX")) { const long v = atol(e); if (v > 0) ctx->max_chunk_px = (uint64_t)v << 20; }
    if (const char* e = getenv("XPNGB_PIPE_LANES")) { const long v = atol(e); if (v > 0 && v <= 64) ctx->pipe_lanes = (uint32_t)v; }
    if (const char* e = getenv("XPNGB_PIPE_MIN_MPIX")) { const long v = atol(e); if (v > 0) ctx->pipe_min_px = (uint64_t)v << 20; }
    { auto k_big = k_rans_v2<256, 32>; cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 * 16); }
    m2_set_attributes();
    if (const char* e = getenv("XPNGB_LAT_MAX_BLOCKS")) ctx->lat_max_blocks = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_S16_LAT_MAX_TILES")) ctx->s16_lat_max_tiles = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_DEC_STAGGER1")) ctx->dec_stagger1 = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_DEC_STAGGER2")) ctx->dec_stagger2 = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_DEC_STAGGER_AT")) ctx->dec_stagger_at = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_ENC_LAT_MAX_BLOCKS")) ctx->enc_lat_max_blocks = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_V2_DIRECT_MAX_TILES")) ctx->v2_direct_max_tiles = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_WALK")) { ctx->walk_global = !strcmp(e, "global"); ctx->walk_ring = !strcmp(e, "ring"); }
    if (const char* e = getenv("XPNGB_DIRECT_MAX_TILES")) ctx->direct_max_tiles = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_UNR_NW")) ctx->unr_force = (uint32_t)atol(e);
    if (const char* e = getenv("XPNGB_FRONT")) ctx->front_v1 = atoi(e) == 1;
    if (const char* e = getenv("XPNGB_PRIO")) ctx->use_prio = atoi(e) != 0;
    if (const char* e = getenv("XPNGB_ENC_SORT")) ctx->enc_sorted = atoi(e) != 0;
    { auto p = k_rans_v2_pair<256>; cudaFuncSetAttribute(p, cudaFuncAttributeMaxDynamicSharedMemorySize, 257 * PAIR_BLK * 24); }
    { auto p = k_rans_v1_pair<256>; cudaFuncSetAttribute(p, cudaFuncAttributeMaxDynamicSharedMemorySize, 257 * PAIR_BLK * 24); }
    cudaFuncSetAttribute(k_dec_unpredict_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 32 * UNR_PITCH);
    cudaFuncSetAttribute(k_dec_unpredict_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 32 * UNR_PITCH);
    cudaFuncSetAttribute(k_dec_walk_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (WALK_SMEM_MAX_SYMS / 8 + 32) * 4);
    cudaFuncSetAttribute(k_dec_rans_v1_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, lat_smem(LUT_ONE_14));
    { auto p = k_dec_rans_pair<1, 0>; cudaFuncSetAttribute(p, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_WARPS * PD_BIG_BYTES); }
    { auto p = k_dec_rans_pair<2, 0>; cudaFuncSetAttribute(p, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_WARPS * PD_BIG_BYTES); }
    *out = ctx;
    return 0;
}

extern "C" void xpngb_destroy(xpngb_ctx* ctx) {
    if (!ctx) return;
    for (xpngb_ctx* l : ctx->lanes) lane_free(l);
    ctx->lanes.clear();
    lane_free(ctx);
}

extern "C" const char* xpngb_last_error(const xpngb_ctx* ctx) { return ctx ? ctx->err : "no context"; }
extern "C" float xpngb_last_kernel_ms(const xpngb_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }
extern "C" uint32_t xpngb_last_launches(const xpngb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* xpngb_stream(const xpngb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" void xpngb_profile(xpngb_ctx* ctx, int on) { if (ctx) { ctx->profile = on == 3 ? 3 : (on ? (ctx->profile == 2 ? 2 : 1) : 0); ctx->prof.clear(); } }
extern "C" uint32_t xpngb_profile_report(const xpngb_ctx* ctx, char* buf, uint32_t cap) {
    uint32_t n = 0;
    if (!ctx || !buf || !cap) return 0;
    buf[0] = 0;
    for (const auto& r : ctx->prof) {
        const int k = snprintf(buf + n, cap - n, "%s %.6f %u\n", r.name, r.ms, r.count);
        if (k < 0 || (uint32_t)k >= cap - n) break;
        n += (uint32_t)k;
    }
    return n;
}

extern "C" uint64_t xpngb_encode_bound(const xpngb_image* imgs, uint32_t n) {
    uint64_t t = 0;
    for (uint32_t i = 0; i < n; i++) t += (8 + imgs[i].w * imgs[i].h * (3 + (imgs[i].A ? 1 : 0)) + 15) & ~15ull;
    return t;
}

extern "C" int xpngb_peek(const void* file, uint64_t size, xpngb_image* img) {
    if (!file || !img || size < 8) return 1;
    uint32_t h[2]; memcpy(h, file, 8);
    img->w = (h[0] & 0xFFFFFFu) + 1; img->h = (h[1] & 0xFFFFFFu) + 1; img->A = (h[1] >> 24) & 1; img->mode = h[0] >> 24;
    return !(img->mode == 1 || img->mode == 2 || img->mode == 7);   // libxpng.c:972
}

// Cut n images into chunks of whole images: at most pipe_lanes chunks of roughly equal pixel count, none below
// pipe_min_px (so small calls stay one chunk) and none above max_chunk_px (scratch bound).  Returns the chunk starts + n.
static std::vector<uint32_t> cut_chunks(const xpngb_ctx* ctx, const xpngb_image* imgs, uint32_t n) {
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += imgs[i].w * imgs[i].h;
    uint64_t want = (total + ctx->pipe_lanes - 1) / ctx->pipe_lanes;
    if (want < ctx->pipe_min_px) want = ctx->pipe_min_px;
    if (want > ctx->max_chunk_px) want = ctx->max_chunk_px;
    // k chunks of equal share: chunk c ends with the image that brings the running pixel count to (c + 1) / k of the total.
    // (Cutting whenever `want` pixels are exceeded leaves a small extra chunk when the images do not divide evenly, and a
    // chunk beyond the lanes waits for a lane: a whole chain latency for a handful of images.)
    uint64_t k = (total + want - 1) / want;
    if (k < 1) k = 1;
    if (k > n) k = n;
    std::vector<uint32_t> cuts{ 0 };
    uint64_t px = 0, c = 1;
    for (uint32_t i = 0; i < n && c < k; i++) {
        px += imgs[i].w * imgs[i].h;
        typedef unsigned __int128 u128;   // w, h <= 2^24 each: the products exceed 64 bits for absurd batches only, but must not wrap
        if ((u128)px * k >= (u128)c * total && i + 1 < n) {
            cuts.push_back(i + 1);
            while (c < k && (u128)px * k >= (u128)c * total) c++;
        }
    }
    cuts.push_back(n);
    return cuts;
}

// ------------------------------------------------------------------------------------------------
// Encode
// ------------------------------------------------------------------------------------------------
// Whole-image single colour at level 2 (libxpng.c:741-753), decided on the device for RGB batches so that the host
// never waits for the scan.
__global__ void k_apply_scan(ImageDesc* imgs, const uint32_t* flags, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && imgs[i].mode == 2 && !(flags[i] & SCAN_NOT_SINGLE)) imgs[i].mode = 2 | 0x100;
}

// One chunk in flight on a lane: what the host needs to finish it.
struct EncChunk {
    xpngb_ctx* lane = nullptr;
    xpngb_image* imgs = nullptr; uint32_t n = 0;
    uint64_t* out_offsets = nullptr; uint64_t* out_sizes = nullptr;
    std::vector<uint32_t> pxsz;
    uint64_t out_base = 0, out_end = 0;     // region of the device arena: [out_base, out_end) is this chunk's bound, files are packed from out_base
    bool finished = false; uint64_t used_end = 0;
};

static int encode_finish(xpngb_ctx* ctx, EncChunk& C);

// Launch everything for one chunk on lane `ctx`; nothing here waits for the device unless the chunk holds RGBA images
// (alpha normalisation changes the pixel size, which the tile plan depends on).
static int encode_issue(xpngb_ctx* ctx, int level, bool lat, const uint8_t* dpix, uint8_t* dout, EncChunk& C) {
    xpngb_image* imgs = C.imgs; const uint32_t n = C.n;
    ctx->cur = ctx->stream;
    bool any_rgba = false;
    for (uint32_t i = 0; i < n; i++) any_rgba |= imgs[i].A != 0;
    std::vector<uint32_t> flags(n, SCAN_NOT_SINGLE);
    std::vector<uint64_t> addr(n);
    std::vector<uint32_t>& pxsz = C.pxsz; pxsz.resize(n);
    for (uint32_t i = 0; i < n; i++) { addr[i] = (uint64_t)(dpix + imgs[i].offset); pxsz[i] = imgs[i].A ? 4 : 3; }
    const unsigned scan_chunks = n >= 148 ? 8u : (1184u / n > 592u ? 592u : 1184u / n);   // ~8 CTAs per SM over the whole chunk
    if (any_rgba) {
        // ---- normalisation / whole-image scans: the one place where the host waits inside a chunk
        std::vector<ImageDesc> S(n);
        for (uint32_t i = 0; i < n; i++) {
            ImageDesc I{}; I.px_off = addr[i]; I.w = (uint32_t)imgs[i].w; I.h = (uint32_t)imgs[i].h; I.pxsz = pxsz[i];
            I.raw_size = imgs[i].w * imgs[i].h * pxsz[i];
            S[i] = I;
        }
        const size_t nb = n * sizeof(ImageDesc);
        ENSURE(imgs, nb); ENSURE(flags, n * 4);
        if (ensure_pin(ctx, ctx->pin_a, nb)) return 1;
        memcpy(ctx->pin_a.p, S.data(), nb);
        CK(cudaMemcpyAsync(ctx->imgs.p, ctx->pin_a.p, nb, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->flags.p, 0, n * 4, ctx->stream));
        LAUNCH(k_image_scan, dim3(scan_chunks, n), 256, 0, (const ImageDesc*)ctx->imgs.p, (uint32_t*)ctx->flags.p);
        CK(cudaMemcpyAsync(flags.data(), ctx->flags.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        // apply normalisation (libxpng.c:688-721)
        uint64_t need = 0;
        for (uint32_t i = 0; i < n; i++) if (imgs[i].A && ((flags[i] & SCAN_DIRTY) || !(flags[i] & SCAN_TRANSLUCENT)))
            need += (imgs[i].w * imgs[i].h * 4 + 15) & ~15ull;
        if (need) {
            ENSURE(norm, need);
            uint64_t o = 0;
            for (uint32_t i = 0; i < n; i++) {
                if (!imgs[i].A) continue;
                const uint64_t npx = imgs[i].w * imgs[i].h;
                uint8_t* dst = (uint8_t*)ctx->norm.p + o;
                const unsigned grid = (unsigned)((npx + 1023) / 1024 > 1184 ? 1184 : (npx + 1023) / 1024);
                if (flags[i] & SCAN_DIRTY) LAUNCH(k_alpha_zero, grid, 256, 0, (const uint32_t*)addr[i], (uint32_t*)dst, npx);
                else if (!(flags[i] & SCAN_TRANSLUCENT)) { LAUNCH(k_alpha_strip, grid, 256, 0, (const uint32_t*)addr[i], dst, npx); pxsz[i] = 3; }
                else continue;
                addr[i] = (uint64_t)dst; o += (npx * 4 + 15) & ~15ull;
            }
        }
    }
    // ---- effective level per image (libxpng.c:735, :741-755).  Without RGBA images the single-colour test of level 2
    // is applied on the device (k_apply_scan); the host only needs to know that level-2 work exists.
    std::vector<uint32_t> mode(n);
    bool any1 = false, any2 = false;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t s = imgs[i].w * imgs[i].h * pxsz[i];
        uint32_t m = (uint32_t)level;
        if (s <= 4) m = 7;
        if (m == 2 && any_rgba && !(flags[i] & SCAN_NOT_SINGLE)) m = 2 | 0x100;
        else if (m == 2 && pxsz[i] == 4) m = 1;
        if ((m == 1) && pxsz[i] == 4 && (imgs[i].w < 4 || imgs[i].h < 4))
            FAIL("image %u: RGBA tiles thinner than 4 pixels are not encodable at level 1/2 (the reference crashes here)", i);
        mode[i] = m; any1 |= m == 1; any2 |= m == 2;
    }
    if (any1 && any2) {
        // level 2 with some images that keep alpha (coded at level 1, libxpng.c:755): the chunk is regrouped into the two
        // families and each family is encoded in ONE pass (files of a family are adjacent in the arena; the offset table,
        // not the arena order, is the contract).  Only chunks with RGBA images get here, and they have waited already.
        uint64_t base = C.out_base;
        for (int fam = 0; fam < 2; fam++) {
            std::vector<uint32_t> idx;
            for (uint32_t i = 0; i < n; i++) if ((mode[i] == 1) == (fam == 0)) idx.push_back(i);
            if (idx.empty()) continue;
            std::vector<xpngb_image> sub(idx.size());
            std::vector<uint64_t> so(idx.size()), ss(idx.size());
            for (size_t k = 0; k < idx.size(); k++) sub[k] = imgs[idx[k]];
            EncChunk S; S.lane = ctx; S.imgs = sub.data(); S.n = (uint32_t)idx.size(); S.out_offsets = so.data(); S.out_sizes = ss.data();
            S.out_base = base; S.out_end = C.out_end;
            if (encode_issue(ctx, fam == 0 ? 1 : 2, lat, dpix, dout, S) || encode_finish(ctx, S)) return 1;
            base = S.used_end;
            for (size_t k = 0; k < idx.size(); k++) { imgs[idx[k]] = sub[k]; C.out_offsets[idx[k]] = so[k]; C.out_sizes[idx[k]] = ss[k]; }
        }
        C.finished = true; C.used_end = base;
        CK(cudaEventRecord(ctx->ev_done, ctx->stream));
        return 0;
    }
    // ---- plan
    Plan P;
    const int sfac = any2 ? 4 : 1, bfac = any2 ? 8 : 4;
    for (uint32_t i = 0; i < n; i++) plan_image(P, addr[i], imgs[i].w, imgs[i].h, pxsz[i], mode[i], sfac, bfac);
    if (P.tiles.size() > (1u << 24) || P.seg_tile.size() > (1u << 28)) FAIL("too many tiles for one chunk (%zu)", P.tiles.size());
    const uint32_t ntiles = (uint32_t)P.tiles.size(), nseg = (uint32_t)P.seg_tile.size();
    if (upload_plan(ctx, P)) return 1;
    ENSURE(outs, n * sizeof(ImageOut)); ENSURE(state, ntiles * sizeof(TileState));
    CK(cudaMemsetAsync(ctx->state.p, 0, ntiles * sizeof(TileState), ctx->stream));
    const TileDesc* d_tiles = (const TileDesc*)ctx->tiles.p;
    const ImageDesc* d_imgs = (const ImageDesc*)ctx->imgs.p;
    const uint32_t* d_seg_tile = (const uint32_t*)ctx->seg_tile.p;
    if (any2 && !any_rgba) {
        ENSURE(flags, n * 4);
        CK(cudaMemsetAsync(ctx->flags.p, 0, n * 4, ctx->stream));
        LAUNCH(k_image_scan, dim3(scan_chunks, n), 256, 0, d_imgs, (uint32_t*)ctx->flags.p);
        LAUNCH(k_apply_scan, (n + 127) / 128, 128, 0, (ImageDesc*)ctx->imgs.p, (const uint32_t*)ctx->flags.p, n);
    }

    if (any1 || any2) {
        const int hstride = any2 ? HIST_STRIDE_M2 : HIST_STRIDE_M1;
        ENSURE(costs, ntiles * 16); ENSURE(hist, (size_t)ntiles * hstride * 4); ENSURE(seginfo, nseg * sizeof(SegInfo));
        ENSURE(place, nseg * sizeof(SegPlace)); ENSURE(sym_area, (size_t)nseg * SEG); ENSURE(bits_area, (size_t)nseg * SEG_BITS_BYTES + 64);
        ENSURE(streams, P.str_total); ENSURE(blocks, P.blk_total);
        if (P.any_rgba) ENSURE(alpha, P.px_total);
        CK(cudaMemsetAsync(ctx->costs.p, 0, ntiles * 16, ctx->stream));
        CK(cudaMemsetAsync(ctx->hist.p, 0, (size_t)ntiles * hstride * 4, ctx->stream));
        LAUNCH(k_predictor_cost, dim3(ntiles, PP_SPLIT), 256, 0, d_tiles, (const uint8_t*)nullptr, (uint32_t*)ctx->costs.p);
    }
    if (any1) {
        FrontArgs fa{ d_tiles, d_seg_tile, nullptr, (const uint32_t*)ctx->costs.p, nullptr, (SegInfo*)ctx->seginfo.p,
                      (uint8_t*)ctx->sym_area.p, (uint8_t*)ctx->bits_area.p, (uint8_t*)ctx->alpha.p, (uint32_t*)ctx->hist.p, nullptr, 0u };
        if (ctx->root->front_v1) LAUNCH(k_front<1>, nseg, FRONT_THREADS, 0, fa);
        else {
            bool any_rgb = false;
            for (uint32_t i = 0; i < n; i++) any_rgb |= pxsz[i] == 3 && mode[i] == 1;
            if (any_rgb) LAUNCH(k_front2<1>, nseg, FRONT_THREADS, 0, fa);
            fa.rgba_only = 1;
            if (P.any_rgba || P.any_wide) LAUNCH(k_front<1>, nseg, FRONT_THREADS, 0, fa);
        }
        TileScanArgs ta{ d_tiles, (const SegInfo*)ctx->seginfo.p, (const uint32_t*)ctx->costs.p, nullptr, (SegPlace*)ctx->place.p, nullptr,
                         (uint32_t*)ctx->hist.p, (TileState*)ctx->state.p, nullptr, ntiles };
        LAUNCH_HI(k_tile_scan<1>, (ntiles + 3) / 4, 128, 0, ta);
        CompactArgs ca{ d_tiles, d_seg_tile, (const SegInfo*)ctx->seginfo.p, (const SegPlace*)ctx->place.p, nullptr, nullptr,
                        (const TileState*)ctx->state.p, (const uint8_t*)ctx->sym_area.p, (const uint8_t*)ctx->bits_area.p,
                        (uint8_t*)ctx->streams.p, nullptr };
        LAUNCH(k_compact<1>, nseg, 256, 0, ca);
        RansV2Args ra{ d_tiles, (TileState*)ctx->state.p, (uint32_t*)ctx->hist.p, (const uint8_t*)ctx->streams.p,
                       (const uint8_t*)ctx->alpha.p, (uint8_t*)ctx->blocks.p, ntiles, 0, 9 };
        if (lat) {
            auto k_rans_v2_pair_16 = k_rans_v2_pair<16>; auto k_rans_v2_pair_256 = k_rans_v2_pair<256>;
            const uint32_t ecap = 10u * ntiles;
            if (ctx->root->enc_sorted) {               // blocks by decreasing length (enc_order.cuh)
                ENSURE(eord, (size_t)(2 * ecap + 2) * 4);
                uint32_t* eo = (uint32_t*)ctx->eord.p;
                LAUNCH_HI(k_enc_order, 1, 1024, 0, EncOrderArgs{ d_tiles, (const TileState*)ctx->state.p, nullptr, ntiles, 1u, eo + 2, eo, ecap });
                ra.order = eo + 2; ra.total = eo;
            }
            if (P.any_rgba) {                          // the alpha blocks are independent of the context blocks: side stream
                FORK_SIDE(0);                          // forks before, is fed after the main launch (see m2_encode_tiles)
                BACK_TO_MAIN();
            }
            LAUNCH_HI(k_rans_v2_pair_16, (9 * ntiles + PAIR_BLK - 1) / PAIR_BLK, 32, 17 * PAIR_BLK * 24, ra);
            if (P.any_rgba) {
                RansV2Args rb = ra; rb.c0 = 9; rb.nc = 1;
                if (ra.order) { rb.order = ra.order + ecap; rb.total = ra.total + 1; }
                ctx->cur = ctx->side[0];
                LAUNCH_HI(k_rans_v2_pair_256, (ntiles + PAIR_BLK - 1) / PAIR_BLK, 32, 257 * PAIR_BLK * 24, rb);
                JOIN_SIDE(0);
            }
        } else {
            auto k_rans_v2_lane_9 = k_rans_v2<9, 128>; auto k_rans_v2_lane_256 = k_rans_v2<256, 32>;
            LAUNCH(k_rans_v2_lane_9, (9 * ntiles + 127) / 128, 128, 9 * 128 * 16, ra);
            if (P.any_rgba) {
                ra.c0 = 9; ra.nc = 1;
                LAUNCH(k_rans_v2_lane_256, (ntiles + 31) / 32, 32, 256 * 32 * 16, ra);
            }
        }
    }
    if (any2) {
        if (m2_encode_tiles(ctx, P, ntiles, nseg, lat)) return 1;
    }
    LAUNCH_HI(k_image_sizes, (n + 127) / 128, 128, 0, d_imgs, d_tiles, (TileState*)ctx->state.p, (ImageOut*)ctx->outs.p, n);
    LAUNCH_HI(k_image_offsets, 1, 1, 0, (ImageOut*)ctx->outs.p, n, C.out_base);
    {
        AssembleArgs aa{ d_imgs, d_tiles, (const TileState*)ctx->state.p, (const ImageOut*)ctx->outs.p, (const SegInfo*)ctx->seginfo.p,
                         (const SegPlace*)ctx->place.p, nullptr, (const uint8_t*)ctx->bits_area.p, (const uint8_t*)ctx->blocks.p, dout };
        if (any2) { if (m2_assemble(ctx, aa, ntiles)) return 1; }
        else LAUNCH(k_assemble_m1, dim3(ntiles, assemble_split(ntiles)), 256, 0, aa);
    }
    if (ensure_pin(ctx, ctx->pin_b, n * sizeof(ImageOut))) return 1;
    CK(cudaMemcpyAsync(ctx->pin_b.p, ctx->outs.p, n * sizeof(ImageOut), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_done, ctx->stream));
    return 0;
}

// Wait for the chunk's size table and hand the results to the caller's arrays.
static int encode_finish(xpngb_ctx* ctx, EncChunk& C) {
    CK(cudaEventSynchronize(ctx->ev_done));
    if (C.finished) return 0;
    const ImageOut* ho = (const ImageOut*)ctx->pin_b.p;
    uint64_t end = C.out_base;
    for (uint32_t i = 0; i < C.n; i++) {
        C.out_offsets[i] = ho[i].off; C.out_sizes[i] = ho[i].size;
        C.imgs[i].A = C.pxsz[i] == 4; C.imgs[i].mode = ho[i].mode & 0xFF;
        end = ho[i].off + ((ho[i].size + 15) & ~15ull);
    }
    if (end > C.out_end) FAIL("output region too small: need %llu bytes, have %llu", (unsigned long long)(end - C.out_base), (unsigned long long)(C.out_end - C.out_base));
    C.finished = true; C.used_end = end;
    return 0;
}

extern "C" int xpngb_encode(xpngb_ctx* ctx, int level, xpngb_image* imgs, uint32_t n, const void* pixels, uint64_t pixels_size,
                            int pixels_on_device, void* out, uint64_t out_cap, int out_on_device, uint64_t* out_offsets,
                            uint64_t* out_sizes) {
    if (!ctx) return 1;
    ctx->err[0] = 0; ctx->launches = 0; ctx->last_ms = 0.f; ctx->cur = ctx->stream;
    if (!imgs || !pixels || !out || !out_offsets || !out_sizes) FAIL("null argument");
    if (!(level == 1 || level == 2 || level == 7)) FAIL("level must be 1, 2 or 7");          // libxpng.c:729
    if (n == 0) return 0;
    for (uint32_t i = 0; i < n; i++) {
        const xpngb_image& m = imgs[i];
        if (!m.w || !m.h || m.w > (1u << 24) || m.h > (1u << 24)) FAIL("image %u: bad dimensions", i);   // libxpng.c:729-730
        if (m.offset & 15) FAIL("image %u: pixel offset must be a multiple of 16", i);
        const uint64_t bytes = m.w * m.h * (3 + (m.A ? 1 : 0));   // w, h <= 2^24: no overflow
        if (bytes > pixels_size || m.offset > pixels_size - bytes) FAIL("image %u: pixels exceed the buffer", i);
    }
    CK(cudaSetDevice(ctx->device));
    const uint64_t bound = xpngb_encode_bound(imgs, n);
    if (out_on_device && out_cap < bound) FAIL("device output buffer must hold xpngb_encode_bound() = %llu bytes", (unsigned long long)bound);
    const uint8_t* dpix = (const uint8_t*)pixels;
    if (!pixels_on_device) { ENSURE(pixels, pixels_size); dpix = (const uint8_t*)ctx->pixels.p; }
    uint8_t* dout = (uint8_t*)out;
    if (!out_on_device) { ENSURE(arena, bound); dout = (uint8_t*)ctx->arena.p; }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    bool any_rgba = false;
    for (uint32_t i = 0; i < n; i++) any_rgba |= imgs[i].A != 0;
    if (level == 7 && !any_rgba) {
        // stored files need no tiles: header + copy
        if (!pixels_on_device) CK(cudaMemcpyAsync(ctx->pixels.p, pixels, pixels_size, cudaMemcpyHostToDevice, ctx->stream));
        uint64_t used = 0;
        std::vector<ImageDesc> I(n); std::vector<uint64_t> offs(n);
        for (uint32_t i = 0; i < n; i++) {
            I[i] = ImageDesc{}; I[i].px_off = (uint64_t)(dpix + imgs[i].offset); I[i].w = (uint32_t)imgs[i].w; I[i].h = (uint32_t)imgs[i].h;
            I[i].pxsz = 3; I[i].raw_size = imgs[i].w * imgs[i].h * 3;
            offs[i] = used; out_offsets[i] = used; out_sizes[i] = 8 + I[i].raw_size; used += (out_sizes[i] + 15) & ~15ull;
            imgs[i].mode = 7;
        }
        ENSURE(imgs, n * sizeof(ImageDesc)); ENSURE(offs, n * 8);
        if (ensure_pin(ctx, ctx->pin_a, n * (sizeof(ImageDesc) + 8))) return 1;
        memcpy(ctx->pin_a.p, I.data(), n * sizeof(ImageDesc)); memcpy((uint8_t*)ctx->pin_a.p + n * sizeof(ImageDesc), offs.data(), n * 8);
        CK(cudaMemcpyAsync(ctx->imgs.p, ctx->pin_a.p, n * sizeof(ImageDesc), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->offs.p, (uint8_t*)ctx->pin_a.p + n * sizeof(ImageDesc), n * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_store7, dim3(296, n), 256, 0, (const ImageDesc*)ctx->imgs.p, (const uint64_t*)ctx->offs.p, dout);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        if (!out_on_device) {
            if (used > out_cap) { CK(cudaStreamSynchronize(ctx->stream)); FAIL("output buffer too small: need %llu bytes, have %llu", (unsigned long long)used, (unsigned long long)out_cap); }
            CK(cudaMemcpyAsync(out, dout, used, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
        return 0;
    }
    // ---- chunks on lanes
    const std::vector<uint32_t> cuts = cut_chunks(ctx, imgs, n);
    const uint32_t nchunks = (uint32_t)cuts.size() - 1;
    uint64_t tiles_est = 0;
    for (uint32_t i = 0; i < n; i++) tiles_est += (imgs[i].w * imgs[i].h + TILE_AREA - 1) / TILE_AREA;
    // kernel family by the size of the CALL (what is in flight at once), not of a chunk
    const bool lat = (level == 2 ? 17 * tiles_est <= (uint64_t)ctx->enc_lat_max_blocks * 5 / 2 : 9 * tiles_est <= ctx->enc_lat_max_blocks);
    std::vector<EncChunk> chunks(nchunks);
    uint64_t host_packed = 0;   // host output: files are packed across chunks
    int rc = 0;
    auto harvest = [&](uint32_t ci) -> int {   // finish chunk ci; host output: start its D2H copy on its lane
        EncChunk& C = chunks[ci];
        xpngb_ctx* lane = C.lane;
        if (encode_finish(lane, C)) return 1;
        if (!out_on_device) {
            const uint64_t len = C.used_end - C.out_base;
            if (host_packed + len > out_cap) FAIL("output buffer too small: need more than %llu bytes, have %llu", (unsigned long long)(host_packed + len), (unsigned long long)out_cap);
            CK(cudaMemcpyAsync((uint8_t*)out + host_packed, dout + C.out_base, len, cudaMemcpyDeviceToHost, lane->stream));
            for (uint32_t i = 0; i < C.n; i++) C.out_offsets[i] = C.out_offsets[i] - C.out_base + host_packed;
            host_packed += len;
        }
        return 0;
    };
    uint64_t region = 0;
    uint32_t harvested = 0;
    for (uint32_t ci = 0; ci < nchunks && !rc; ci++) {
        const uint32_t li = ci % ctx->pipe_lanes;
        while (!rc && ci >= ctx->pipe_lanes && harvested <= ci - ctx->pipe_lanes) rc = harvest(harvested++);   // the lane's previous chunk must be done
        if (rc) break;
        xpngb_ctx* lane = lane_get(ctx, li);
        if (!lane) { snprintf(ctx->err, sizeof ctx->err, "cannot create pipeline lane %u", li); rc = 1; break; }
        EncChunk& C = chunks[ci];
        C.lane = lane; C.imgs = imgs + cuts[ci]; C.n = cuts[ci + 1] - cuts[ci];
        C.out_offsets = out_offsets + cuts[ci]; C.out_sizes = out_sizes + cuts[ci];
        C.out_base = region; region += xpngb_encode_bound(C.imgs, C.n); C.out_end = region;
        if (lane != ctx) { if (cudaStreamWaitEvent(lane->stream, ctx->ev0, 0) != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaStreamWaitEvent failed"); rc = 1; break; } }
        if (!pixels_on_device) {   // the chunk's pixel range (images of a chunk are adjacent in every sane layout; any layout is correct)
            uint64_t lo = ~0ull, hi = 0;
            for (uint32_t i = 0; i < C.n; i++) {
                const uint64_t a = C.imgs[i].offset, b = a + C.imgs[i].w * C.imgs[i].h * (3 + (C.imgs[i].A ? 1 : 0));
                if (a < lo) lo = a; if (b > hi) hi = b;
            }
            if (cudaMemcpyAsync((uint8_t*)ctx->pixels.p + lo, (const uint8_t*)pixels + lo, hi - lo, cudaMemcpyHostToDevice, lane->stream) != cudaSuccess) {
                snprintf(ctx->err, sizeof ctx->err, "host to device copy failed"); rc = 1; break;
            }
        }
        rc = encode_issue(lane, level, lat, dpix, dout, C);
    }
    while (!rc && harvested < nchunks) rc = harvest(harvested++);
    // join: the root stream waits for every lane, so that ev1 closes the whole call
    for (uint32_t li = 1; li < ctx->pipe_lanes && li <= ctx->lanes.size(); li++) {
        xpngb_ctx* lane = ctx->lanes[li - 1];
        cudaEventRecord(lane->ev_done, lane->stream);
        cudaStreamWaitEvent(ctx->stream, lane->ev_done, 0);
    }
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) { snprintf(ctx->err, sizeof ctx->err, "device error: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
    if (rc) { for (xpngb_ctx* l : ctx->lanes) cudaStreamSynchronize(l->stream); return 1; }
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    tl_dump(ctx, "enc");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Traversal-order operations (Mirroring_and_Rotating/tool.c): mirror / quarter-turn n pixmaps
// ------------------------------------------------------------------------------------------------
static int orient_launch(xpngb_ctx* ctx, int op, xpngb_image* imgs, uint32_t n, const uint8_t* dsrc, uint8_t* ddst) {
    std::vector<OrientDesc> D(n);
    uint64_t ctas = 0;
    for (uint32_t i = 0; i < n; i++) {
        OrientDesc& d = D[i];
        d.src = (uint64_t)(dsrc + imgs[i].offset); d.dst = (uint64_t)(ddst + imgs[i].offset);
        d.w = (uint32_t)imgs[i].w; d.h = (uint32_t)imgs[i].h; d.pxsz = imgs[i].A ? 4 : 3; d.op = (uint32_t)op;
        d.tiles_x = (d.w + ORI_T - 1) / ORI_T; d.first_cta = (uint32_t)ctas;
        ctas += (uint64_t)d.tiles_x * ((d.h + ORI_T - 1) / ORI_T);
        if (ctas > 0x7FFFFFFFull) FAIL("too many pixels for one orientation launch");
    }
    const size_t nb = n * sizeof(OrientDesc);
    ENSURE(odesc, nb);
    if (ensure_pin(ctx, ctx->pin_o, nb)) return 1;
    memcpy(ctx->pin_o.p, D.data(), nb);
    CK(cudaMemcpyAsync(ctx->odesc.p, ctx->pin_o.p, nb, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_orient, (unsigned)ctas, 256, 0, (const OrientDesc*)ctx->odesc.p, n);
    if (op == OP_R90 || op == OP_R270) for (uint32_t i = 0; i < n; i++) { const uint64_t t = imgs[i].w; imgs[i].w = imgs[i].h; imgs[i].h = t; }
    return 0;
}

static int orient_check(xpngb_ctx* ctx, int op, const xpngb_image* imgs, uint32_t n, uint64_t size) {
    if (op < OP_R90 || op > OP_TR) FAIL("unknown orientation op %d", op);
    for (uint32_t i = 0; i < n; i++) {
        const xpngb_image& m = imgs[i];
        if (!m.w || !m.h || m.w > (1u << 24) || m.h > (1u << 24)) FAIL("image %u: bad dimensions", i);
        if (m.offset & 15) FAIL("image %u: pixel offset must be a multiple of 16", i);
        const uint64_t bytes = m.w * m.h * (3 + (m.A ? 1 : 0));
        if (bytes > size || m.offset > size - bytes) FAIL("image %u: pixels exceed the buffer", i);
    }
    return 0;
}

extern "C" int xpngb_transform(xpngb_ctx* ctx, int op, xpngb_image* imgs, uint32_t n, const void* src, uint64_t size, int src_on_device,
                               void* dst, int dst_on_device) {
    if (!ctx) return 1;
    ctx->err[0] = 0; ctx->launches = 0; ctx->last_ms = 0.f; ctx->cur = ctx->stream;
    if (!imgs || !src || !dst) FAIL("null argument");
    if (src == dst) FAIL("in-place orientation is not supported");
    if (!src_on_device == !dst_on_device) {   // same side: the two byte ranges must be disjoint
        const uintptr_t a = (uintptr_t)src, b = (uintptr_t)dst;
        if (a < b + size && b < a + size) FAIL("source and destination pixel buffers overlap");
    }
    if (orient_check(ctx, op, imgs, n, size)) return 1;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const uint8_t* dsrc = (const uint8_t*)src;
    if (!src_on_device) {
        ENSURE(pixels, size);
        CK(cudaMemcpyAsync(ctx->pixels.p, src, size, cudaMemcpyHostToDevice, ctx->stream));
        dsrc = (const uint8_t*)ctx->pixels.p;
    }
    uint8_t* ddst = (uint8_t*)dst;
    if (!dst_on_device) { ENSURE(oriented, size); ddst = (uint8_t*)ctx->oriented.p; }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (orient_launch(ctx, op, imgs, n, dsrc, ddst)) return 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    if (!dst_on_device) CK(cudaMemcpyAsync(dst, ddst, size, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return 0;
}

extern "C" int xpngb_encode_oriented(xpngb_ctx* ctx, int level, int op, xpngb_image* imgs, uint32_t n, const void* pixels,
                                     uint64_t pixels_size, int pixels_on_device, void* out, uint64_t out_cap, int out_on_device,
                                     uint64_t* out_offsets, uint64_t* out_sizes) {
    if (!ctx) return 1;
    ctx->err[0] = 0; ctx->cur = ctx->stream;
    if (!imgs || !pixels) FAIL("null argument");
    if (orient_check(ctx, op, imgs, n, pixels_size)) return 1;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const uint8_t* dsrc = (const uint8_t*)pixels;
    if (!pixels_on_device) {
        ENSURE(pixels, pixels_size);
        CK(cudaMemcpyAsync(ctx->pixels.p, pixels, pixels_size, cudaMemcpyHostToDevice, ctx->stream));
        dsrc = (const uint8_t*)ctx->pixels.p;
    }
    ENSURE(oriented, pixels_size);
    if (orient_launch(ctx, op, imgs, n, dsrc, (uint8_t*)ctx->oriented.p)) return 1;
    const uint32_t l0 = ctx->launches;
    const int rc = xpngb_encode(ctx, level, imgs, n, ctx->oriented.p, pixels_size, 1, out, out_cap, out_on_device, out_offsets, out_sizes);
    ctx->launches += l0;
    if (rc && (op == OP_R90 || op == OP_R270))   // failed: hand the descriptors back in the caller's orientation
        for (uint32_t i = 0; i < n; i++) { const uint64_t t = imgs[i].w; imgs[i].w = imgs[i].h; imgs[i].h = t; }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Decode
// ------------------------------------------------------------------------------------------------
// Work lists of the pair decoders (dec_rans_pair.cuh): two sets, one per level family present in the chunk.
static size_t pdw_bytes(uint32_t cap) { return (size_t)(2 * PD_NCLASS * PD_BUCKETS + PD_NCLASS + 5) * 4 + (size_t)PD_NCLASS * cap * 4; }
static PdWork pdw_at(void* base, uint32_t cap) {
    uint32_t* p = (uint32_t*)base;
    PdWork W; W.hist = p; W.start = p + PD_NCLASS * PD_BUCKETS; W.total = p + 2 * PD_NCLASS * PD_BUCKETS; W.order = W.total + PD_NCLASS + 5; W.cap = cap;
    return W;
}

// Launch the decode of one chunk on lane `ctx`.  hdr: the two header words of each file.  `lat`: warp-per-block chain
// kernels (few blocks in the whole call) instead of the pair decoders.  Never waits for the device.
static int decode_issue(xpngb_ctx* ctx, bool lat, xpngb_image* imgs, uint32_t n, const uint32_t* hdr, const uint8_t* din,
                        const uint64_t* file_offsets, const uint64_t* file_sizes, uint8_t* dpx, uint64_t pixels_cap, uint32_t call_tiles) {
    ctx->cur = ctx->stream;
    Plan P;
    std::vector<DecImage> DI(n);
    bool any1 = false, any2 = false;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t w = (hdr[2 * i] & 0xFFFFFFu) + 1, h = (hdr[2 * i + 1] & 0xFFFFFFu) + 1;
        const uint32_t A = (hdr[2 * i + 1] >> 24) & 1, mode = hdr[2 * i] >> 24, pxsz = 3 + A;
        if (!(mode == 1 || mode == 2 || mode == 7)) FAIL("file: unknown mode %u", mode);   // libxpng.c:972
        if (imgs[i].w && (imgs[i].w != w || imgs[i].h != h)) FAIL("file: header %llux%llu does not match the descriptor",
                                                                   (unsigned long long)w, (unsigned long long)h);
        imgs[i].w = w; imgs[i].h = h; imgs[i].A = A; imgs[i].mode = mode;
        const uint64_t s = w * h * pxsz;
        if (imgs[i].offset & 15) FAIL("file: pixel offset must be a multiple of 16");
        if (s > pixels_cap || imgs[i].offset > pixels_cap - s) FAIL("file: pixels exceed the output buffer");
        uint32_t m = mode;
        if (mode == 7) { if (file_sizes[i] < 8 + s) FAIL("file: truncated stored image"); }
        else if (file_sizes[i] == 11 + A && ((hdr[2 * i + 1] >> 24) & 2)) m = mode | 0x100;      // libxpng.c:976
        // dimensions come from an untrusted header and size every scratch buffer below: a tile is at most 666 x 666 pixels and
        // its blob at least 8 bytes, so a coded file shorter than that cannot hold the image it claims
        else if (file_sizes[i] < 8 + 8 * (w * h / (666ull * 666ull))) FAIL("file: %llu bytes cannot hold a %llux%llu image",
                                                                           (unsigned long long)file_sizes[i], (unsigned long long)w, (unsigned long long)h);
        any1 |= m == 1; any2 |= m == 2;
        DecImage& D = DI[i];
        D.file_off = file_offsets[i]; D.file_size = file_sizes[i]; D.px_off = imgs[i].offset; D.w = (uint32_t)w; D.h = (uint32_t)h;
        D.pxsz = pxsz; D.mode = m;
    }
    for (uint32_t i = 0; i < n; i++) {
        plan_image(P, (uint64_t)(dpx + imgs[i].offset), DI[i].w, DI[i].h, DI[i].pxsz, DI[i].mode, any2 ? 4 : 1, 0);
        DI[i].tile0 = P.imgs[i].tile0; DI[i].ntiles = P.imgs[i].ntiles;
    }
    if (P.tiles.size() > (1u << 24) || P.seg_tile.size() > (1u << 28)) FAIL("too many tiles for one chunk (%zu)", P.tiles.size());
    const uint32_t ntiles = (uint32_t)P.tiles.size();
    if (upload_plan(ctx, P)) return 1;
    ENSURE(dimgs, n * sizeof(DecImage)); ENSURE(dtiles, ntiles * sizeof(DecTile)); ENSURE(errflag, 4);
    if (ensure_pin(ctx, ctx->pin_b, n * (sizeof(DecImage) + 8) + 64)) return 1;
    memcpy(ctx->pin_b.p, DI.data(), n * sizeof(DecImage));
    CK(cudaMemcpyAsync(ctx->dimgs.p, ctx->pin_b.p, n * sizeof(DecImage), cudaMemcpyHostToDevice, ctx->stream));
    bool any7 = false;
    for (uint32_t i = 0; i < n; i++) any7 |= DI[i].mode == 7;
    if (any7) {   // where each stored image's file starts (k_load7), ~0 = not a stored image
        uint64_t* cf = (uint64_t*)((uint8_t*)ctx->pin_b.p + n * sizeof(DecImage));
        for (uint32_t i = 0; i < n; i++) cf[i] = DI[i].mode == 7 ? DI[i].file_off : ~0ull;
        ENSURE(offs, n * 8);
        CK(cudaMemcpyAsync(ctx->offs.p, cf, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->errflag.p, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->dtiles.p, 0, ntiles * sizeof(DecTile), ctx->stream));
    const TileDesc* d_tiles = (const TileDesc*)ctx->tiles.p;
    const DecImage* d_imgs = (const DecImage*)ctx->dimgs.p;
    DecTile* d_dt = (DecTile*)ctx->dtiles.p;
    int* d_err = (int*)ctx->errflag.p;
    const uint32_t nseg = (uint32_t)P.seg_tile.size();
    const uint32_t* d_seg_tile = (const uint32_t*)ctx->seg_tile.p;
    bool side_busy[xpngb_ctx::NSIDE] = {};
    const uint32_t pd_cap = 17u * ntiles;
    if (any1 || any2) {
        ENSURE(streams, P.str_total); ENSURE(nlseq, P.px_total); ENSURE(rows, P.row_total * sizeof(RowInfo));
        ENSURE(rowcnt, P.row_total * 4); ENSURE(edge, P.row_total * 16);
        ENSURE(ccnt, (size_t)nseg * 9 * 4); ENSURE(cbit, (size_t)nseg * 4); ENSURE(resv, (P.px_total + 4 * P.row_total) * 4);
        if (P.any_rgba) { ENSURE(alpha, P.px_total); ENSURE(plane, P.px_total); }
        if (!lat) { ENSURE(pdw, 2 * pdw_bytes(pd_cap)); CK(cudaMemsetAsync(ctx->pdw.p, 0, 2 * pdw_bytes(pd_cap), ctx->stream)); }
        LAUNCH_HI(k_dec_tile_offsets, (n + 127) / 128, 128, 0, d_imgs, din, d_dt, n, d_err);
    }
    // Stream plan of one chunk.  Main stream: the level-2 family (or the only family).  side[2]: the level-1 family
    // when both are present (config 0: the corpus mixes RGB and RGBA files).  side[0], side[1]: level-2
    // value streams.  side[3]: the alpha plane.  Each family's chain is parse -> context rANS -> context walk.
    uint32_t maxpx = 0;
    for (const TileDesc& t : P.tiles) if (t.npx > maxpx) maxpx = t.npx;
    const uint32_t wsm = (maxpx / 8 + 32) * 4;          // nibble-packed streams of the largest tile + pad words
    PdWork Wl[3] = {};                              // work lists of the pair decoders / the batch walk, per level family
    auto launch_walk = [&](uint32_t mode) -> int {
        WalkArgs wa{ d_tiles, d_imgs, d_dt, (const uint8_t*)ctx->streams.p, (uint8_t*)ctx->nlseq.p, ntiles, mode };
        if (!lat && !ctx->root->walk_ring && !ctx->root->walk_global) {   // batches: three tiles per warp, longest tiles first
            const PdWork& W = Wl[mode];
            LAUNCH_HI(k_dec_walk3<2>, (ntiles + 5) / 6, 64, 0, wa, (const uint32_t*)(W.order + (size_t)PD_WALK * W.cap), (const uint32_t*)(W.total + PD_WALK));
            return 0;
        }
        // the shared-memory walk only pays when every tile's CTA of the CALL is resident at once (no waves): 227 KB per SM, 148 SMs
        const uint32_t resident = 148u * ((227u * 1024u) / (wsm + 1024u));
        xpngb_ctx* r = ctx->root;
        if (call_tiles <= resident && maxpx <= WALK_SMEM_MAX_SYMS && !r->walk_ring && !r->walk_global) LAUNCH_HI(k_dec_walk_smem<0>, ntiles, 32, wsm, wa);
        else if (r->walk_global) {                     // XPNGB_WALK=global: the older variant with refills straight from global memory
            if (ntiles <= 592) LAUNCH_HI(k_dec_walk_lat<1>, ntiles, 32, 0, wa);
            else LAUNCH_HI(k_dec_walk_lat<4>, (ntiles + 3) / 4, 128, 0, wa);
        }
        else if (call_tiles <= 592) LAUNCH_HI(k_dec_walk_ring<1>, ntiles, 32, 0, wa);
        else LAUNCH_HI(k_dec_walk_ring<4>, (ntiles + 3) / 4, 128, 0, wa);
        return 0;
    };
    // pair decoders of one level family: work lists, run / raw fills, then one launch per class
    auto pd_lists = [&](uint32_t mode, PdWork& W) -> int {
        W = pdw_at((uint8_t*)ctx->pdw.p + (mode == 1 ? 0 : pdw_bytes(pd_cap)), pd_cap);
        const unsigned g = (17u * ntiles + 255) / 256;
        LAUNCH_HI(k_pd_count, g, 256, 0, d_tiles, d_imgs, (const DecTile*)d_dt, ntiles, mode, W);
        LAUNCH_HI(k_pd_scan, PD_NCLASS, 1024, 0, W);
        LAUNCH_HI(k_pd_fill, g, 256, 0, d_tiles, d_imgs, (const DecTile*)d_dt, ntiles, mode, W);
        return 0;
    };
    auto pd_args = [&](const PdWork& W, int cls) {
        return PairDecArgs{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, (uint8_t*)ctx->alpha.p, W.order + (size_t)cls * W.cap, W.total + cls, d_err };
    };
    auto pd_grid = [&](uint32_t per_tile) { return (unsigned)(((uint64_t)per_tile * ntiles + PD_BLK * PD_WARPS - 1) / (PD_BLK * PD_WARPS)); };
    if (any1) {
        const bool own_stream = any2;                   // level-1 family next to a level-2 family
        cudaStream_t f1 = own_stream ? side_of(ctx, 2) : ctx->stream;
        if (own_stream) { FORK_SIDE(2); side_busy[2] = true; }
        LAUNCH_HI(k_dec_parse_m1, (ntiles + 127) / 128, 128, 0, d_tiles, d_imgs, din, d_dt, ntiles, d_err);
        RansDecArgs ra{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, (uint8_t*)ctx->alpha.p, ntiles, 0, 9, d_err };
        AlphaArgs al{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->alpha.p, (uint8_t*)ctx->plane.p, (uint32_t*)ctx->rowcnt.p };
        if (lat) {
            if (P.any_rgba) {                             // the alpha plane is independent of the context walk
                FORK_FROM(f1, 3); side_busy[3] = true;
                RansDecArgs rb = ra; rb.c0 = 9; rb.nc = 1;
                auto k_dec_rans_v2_lat_alpha = k_dec_rans_v2_lat;
                LAUNCH_HI(k_dec_rans_v2_lat_alpha, ntiles, 32, lat_smem(LUT_TWO_15), rb, LUT_TWO_15);
                LAUNCH(k_dec_alpha, ntiles, 256, 0, al);
                LAUNCH(k_dec_rows_rgba, ntiles, 32, 0, d_tiles, d_imgs, (const DecTile*)d_dt, (const uint32_t*)ctx->rowcnt.p, (RowInfo*)ctx->rows.p);
            }
            ctx->cur = f1;
            const uint32_t lut12 = call_tiles <= ctx->root->v2_direct_max_tiles ? LUT_ONE_12 : LUT_TWO_12;
            LAUNCH_HI(k_dec_rans_v2_lat, 9 * ntiles, 32, lat_smem(lut12), ra, lut12);
        } else {
            PdWork& W = Wl[1];
            if (pd_lists(1, W)) return 1;
            PdFillArgs fa{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, (uint8_t*)ctx->alpha.p, ntiles, 1 };
            LAUNCH_HI(k_pd_fill_blocks, (17u * ntiles + 3) / 4, 128, 0, fa);
            if (P.any_rgba) {
                FORK_FROM(f1, 3); side_busy[3] = true;
                auto k_dec_rans_pair_v2_alpha = k_dec_rans_pair<2, 0>;
                LAUNCH_HI(k_dec_rans_pair_v2_alpha, pd_grid(1), PD_WARPS * 32, PD_WARPS * PD_BIG_BYTES, pd_args(W, PD_BIG));
                LAUNCH(k_dec_alpha, ntiles, 256, 0, al);
                LAUNCH(k_dec_rows_rgba, ntiles, 32, 0, d_tiles, d_imgs, (const DecTile*)d_dt, (const uint32_t*)ctx->rowcnt.p, (RowInfo*)ctx->rows.p);
            }
            ctx->cur = f1;
            auto k_dec_rans_pair_v2_ctx = k_dec_rans_pair<2, 8>;
            LAUNCH_HI(k_dec_rans_pair_v2_ctx, pd_grid(9), PD_WARPS * 32, 0, pd_args(W, PD_S8));
        }
        if (!ctx->root->dec_stagger_at && !own_stream) CK(cudaEventRecord(ctx->ev_stage, ctx->stream));
        if (launch_walk(1)) return 1;
        BACK_TO_MAIN();
    }
    if (any2) {
        LAUNCH_HI(k_dec_parse_m2, (ntiles + 127) / 128, 128, 0, d_tiles, d_imgs, din, d_dt, ntiles, d_err);
        if (lat) {
            // value streams on the side streams (joined before the residual kernels); LAT_M2_ORDER: 0..2 alphabets of at
            // most 16 symbols (direct table, 64 KiB), 3..7 larger alphabets (two-level), 8..16 contexts, 17 grey plane
            // a batch with more chains than fit next to 64 KiB tables (3 per SM) trades the shorter dependent step of the
            // direct table for residency: two-level tables (17 KiB) keep 12 chains per SM
            const bool direct = call_tiles <= ctx->root->direct_max_tiles;
            const uint32_t lut16 = direct ? LUT_ONE_14 : LUT_TWO_14;
            RansV1LatArgs la{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, ntiles, 0, 3, lut16, 0u, ~0u, d_err };
            auto k_dec_rans_v1_lat_values16 = k_dec_rans_v1_lat; auto k_dec_rans_v1_lat_values256 = k_dec_rans_v1_lat;
            auto k_dec_rans_v1_lat_grey = k_dec_rans_v1_lat; auto k_dec_rans_v1_lat_ctx = k_dec_rans_v1_lat;   // names for the profile report
            auto k_dec_rans_v1_lat_ctx_short = k_dec_rans_v1_lat;
            FORK_SIDE(0); side_busy[0] = true;
            LAUNCH_HI(k_dec_rans_v1_lat_values16, 3 * ntiles, 32, lat_smem(lut16), la);
            FORK_SIDE(1); side_busy[1] = true;
            la.j0 = 3; la.nj = 5; la.lut_bytes = LUT_TWO_14;
            LAUNCH_HI(k_dec_rans_v1_lat_values256, 5 * ntiles, 32, lat_smem(LUT_TWO_14), la);
            la.j0 = 17; la.nj = 1; la.lut_bytes = LUT_TWO_15;
            LAUNCH_HI(k_dec_rans_v1_lat_grey, ntiles, 32, lat_smem(LUT_TWO_15), la);
            BACK_TO_MAIN();
            // context streams: the few long ones (they bound the walk's start on real images) get the 64 KiB direct table,
            // the many short ones a two-level table on a side stream, so that everything stays resident
            constexpr uint32_t CTX_LONG = 24576;
            la.j0 = 8; la.nj = 9; la.lut_bytes = LUT_TWO_14; la.n_lo = 0; la.n_hi = direct ? CTX_LONG : ~0u;
            if (direct) {
                FORK_SIDE(4); side_busy[4] = true;
                LAUNCH_HI(k_dec_rans_v1_lat_ctx_short, 9 * ntiles, 32, lat_smem(LUT_TWO_14), la);
                BACK_TO_MAIN();
                la.lut_bytes = LUT_ONE_14; la.n_lo = CTX_LONG; la.n_hi = ~0u;
                LAUNCH_HI(k_dec_rans_v1_lat_ctx, 9 * ntiles, 32, lat_smem(LUT_ONE_14), la);
                JOIN_SIDE(4); side_busy[4] = false;       // the walk needs every context stream
            } else LAUNCH_HI(k_dec_rans_v1_lat_ctx, 9 * ntiles, 32, lat_smem(LUT_TWO_14), la);
        } else {
            // the value streams (16-symbol alphabet: the longest chains of a tile; large alphabets) start first on the side
            // streams; contexts and the 8-symbol value streams on the main stream, followed by the walk
            PdWork& W = Wl[2];
            if (pd_lists(2, W)) return 1;
            PdFillArgs fa{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, (uint8_t*)ctx->alpha.p, ntiles, 2 };
            auto k_dec_rans_pair_v1_s16 = k_dec_rans_pair<1, 15>; auto k_dec_rans_pair_v1_big = k_dec_rans_pair<1, 0>; auto k_dec_rans_pair_v1_s8 = k_dec_rans_pair<1, 8>;
            FORK_SIDE(0); side_busy[0] = true;
            if (call_tiles <= ctx->root->s16_lat_max_tiles) {
                // few enough of the longest chains to be resident all at once as warp-per-block chains with a two-level table
                // (about 55 cycles per symbol, against about 170 per symbol and state of the pair decoder)
                RansV1LatArgs la{ d_tiles, d_imgs, d_dt, din, (uint8_t*)ctx->streams.p, ntiles, 0, 1, LUT_TWO_14, 0u, ~0u, d_err };
                auto k_dec_rans_v1_lat_st4 = k_dec_rans_v1_lat;
                LAUNCH_HI(k_dec_rans_v1_lat_st4, ntiles, 32, lat_smem(LUT_TWO_14), la);
            } else
            LAUNCH_HI(k_dec_rans_pair_v1_s16, pd_grid(1), PD_WARPS * 32, 0, pd_args(W, PD_S16));
            FORK_SIDE(1); side_busy[1] = true;
            LAUNCH_HI(k_dec_rans_pair_v1_big, pd_grid(5), PD_WARPS * 32, PD_WARPS * PD_BIG_BYTES, pd_args(W, PD_BIG));
            BACK_TO_MAIN();
            LAUNCH_HI(k_pd_fill_blocks, (17u * ntiles + 3) / 4, 128, 0, fa);   // run / raw blocks (contexts among them: before the walk)
            LAUNCH_HI(k_dec_rans_pair_v1_s8, pd_grid(11), PD_WARPS * 32, 0, pd_args(W, PD_S8));
        }
        if (!ctx->root->dec_stagger_at) CK(cudaEventRecord(ctx->ev_stage, ctx->stream));
        if (launch_walk(2)) return 1;
    }
    if (ctx->root->dec_stagger_at || !(any1 || any2)) CK(cudaEventRecord(ctx->ev_stage, ctx->stream));
    // batches: RGB tiles with word-aligned rows get the row-pitched residual plane and k_dec_unpredict_rgb
    const uint32_t pitched = ((!lat || call_tiles > ctx->root->unr_multi_max_tiles) && !ctx->root->unr_force) ? 1u : 0u;
    bool all_pitched = true;
    for (const TileDesc& t : P.tiles) all_pitched &= tile_pitched(t);
    if (any1 || any2) {
        if (side_busy[2]) JOIN_SIDE(2);                   // the other family's nl sequences
        ChunkArgs ch{ d_tiles, d_seg_tile, d_imgs, d_dt, (const uint8_t*)ctx->nlseq.p, (const uint8_t*)ctx->streams.p, din,
                      (uint32_t*)ctx->ccnt.p, (uint32_t*)ctx->cbit.p, (uint32_t*)ctx->resv.p, ntiles, d_err, pitched };
        LAUNCH(k_dec_chunk_hist, nseg, 256, 0, ch);
        LAUNCH_HI(k_dec_chunk_scan, (ntiles + 3) / 4, 128, 0, ch);
        for (int k = 0; k < xpngb_ctx::NSIDE; k++) if (k != 2 && side_busy[k]) JOIN_SIDE(k);
        if (any1) LAUNCH(k_dec_residuals<1>, nseg, 256, 0, ch);
        if (any2) { LAUNCH(k_dec_residuals<2>, nseg, 256, 0, ch); LAUNCH(k_dec_residuals_grey, nseg, 256, 0, ch); }
        UnpredArgs ua{ d_tiles, d_imgs, d_dt, din, (const uint32_t*)ctx->resv.p, (const uint8_t*)ctx->plane.p, (const RowInfo*)ctx->rows.p,
                       (uint4*)ctx->edge.p, 0, pitched };
        uint32_t maxw = 0;
        for (const TileDesc& t : P.tiles) if (t.w > maxw) maxw = t.w;
        // 16 warps per tile when the call has fewer tiles than SMs (shortest band pipeline), 8 up to a few per SM; beyond
        // that one warp per tile: bands of a tile depend on each other, so extra warps only wait on the band above, and
        // with thousands of tiles the tiles themselves are the parallelism
        uint32_t nw = call_tiles <= 148 ? 16u : (call_tiles <= ctx->root->unr_multi_max_tiles ? 8u : 1u);
        if (ctx->root->unr_force) nw = ctx->root->unr_force;
        if (pitched) {   // batches: one warp per RGB tile on the row-pitched residual plane; whatever it cannot take goes on below
            LAUNCH(k_dec_unpredict_rgb, ntiles, 32, 0, ua);
            if (all_pitched) nw = 0;
        }
        if (nw == 0) { }
        else if (nw == 16) LAUNCH(k_dec_unpredict_rows<16>, ntiles, 16 * 32, 16 * 2 * 32 * UNR_PITCH, ua);
        else if (nw == 8) LAUNCH(k_dec_unpredict_rows<8>, ntiles, 8 * 32, 8 * 2 * 32 * UNR_PITCH, ua);
        else if (nw == 4) LAUNCH(k_dec_unpredict_rows<4>, ntiles, 4 * 32, 4 * 2 * 32 * UNR_PITCH, ua);
        else if (nw == 2) LAUNCH(k_dec_unpredict_rows<2>, ntiles, 2 * 32, 2 * 2 * 32 * UNR_PITCH, ua);
        else LAUNCH(k_dec_unpredict_rows<1>, ntiles, 32, 2 * 32 * UNR_PITCH, ua);
        ua.min_w = UNR_MAXW;
        if (maxw > UNR_MAXW) LAUNCH(k_dec_unpredict, ntiles, UNP_THREADS, 0, ua);   // very wide, flat tiles only
        if (any2) LAUNCH(k_dec_grey_raw, ntiles, 256, 0, d_tiles, d_imgs, (const DecTile*)d_dt, din);
    }
    if (any7) LAUNCH(k_load7, dim3(296, n), 256, 0, (const ImageDesc*)ctx->imgs.p, (const uint64_t*)ctx->offs.p, din);   // stored images: flat copies
    LAUNCH(k_dec_copy, ntiles, 256, 0, d_tiles, d_imgs, (const DecTile*)d_dt, din, (uint8_t*)nullptr);
    // the error flag travels to pinned memory behind the DecImage table
    CK(cudaMemcpyAsync((uint8_t*)ctx->pin_b.p + n * (sizeof(DecImage) + 8), d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

extern "C" int xpngb_decode(xpngb_ctx* ctx, xpngb_image* imgs, uint32_t n, const void* files, uint64_t files_size, int files_on_device,
                            const uint64_t* file_offsets, const uint64_t* file_sizes, void* pixels, uint64_t pixels_cap,
                            int pixels_on_device) {
    if (!ctx) return 1;
    ctx->err[0] = 0; ctx->launches = 0; ctx->last_ms = 0.f; ctx->cur = ctx->stream;
    if (!imgs || !files || !file_offsets || !file_sizes || !pixels) FAIL("null argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < n; i++) {
        if (file_sizes[i] < 11 || file_sizes[i] > files_size || file_offsets[i] > files_size - file_sizes[i]) FAIL("file %u: bad offset/size", i);
    }
    const uint8_t* din = (const uint8_t*)files;
    if (!files_on_device) { ENSURE(files, files_size); din = (const uint8_t*)ctx->files.p; }
    // ---- headers (libxpng.c:969-973)
    std::vector<uint32_t> hdr(2 * n);
    if (!files_on_device) {
        for (uint32_t i = 0; i < n; i++) memcpy(&hdr[2 * i], (const uint8_t*)files + file_offsets[i], 8);
    } else {
        ENSURE(offs, n * 8); ENSURE(hdr, n * 8);
        if (ensure_pin(ctx, ctx->pin_a, n * 8)) return 1;
        memcpy(ctx->pin_a.p, file_offsets, n * 8);
        CK(cudaMemcpyAsync(ctx->offs.p, ctx->pin_a.p, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_gather_headers, (n + 127) / 128, 128, 0, din, (const uint64_t*)ctx->offs.p, n, (uint32_t*)ctx->hdr.p);
        CK(cudaMemcpyAsync(hdr.data(), ctx->hdr.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    uint8_t* dpx = (uint8_t*)pixels;
    if (!pixels_on_device) { ENSURE(pixels, pixels_cap); dpx = (uint8_t*)ctx->pixels.p; }
    // chunks are cut by the header dimensions (validated per chunk in decode_issue)
    std::vector<xpngb_image> dims(n);
    uint64_t tiles_est = 0; bool l2 = false;
    for (uint32_t i = 0; i < n; i++) {
        dims[i].w = (hdr[2 * i] & 0xFFFFFFu) + 1; dims[i].h = (hdr[2 * i + 1] & 0xFFFFFFu) + 1;
        const uint32_t mode = hdr[2 * i] >> 24;
        if (mode == 1 || mode == 2) tiles_est += (dims[i].w * dims[i].h + TILE_AREA - 1) / TILE_AREA;
        l2 |= mode == 2;
    }
    // level 2: above ~600 tiles the pair family wins, with the 16-symbol value streams warp-per-block while they all fit
    // (s16_lat_max_tiles); level 1: the warp-per-block family up to ~1300 tiles (measured at 125 / 250 / 500 frames, profiles/)
    const bool lat = l2 ? 17 * tiles_est <= (uint64_t)ctx->lat_max_blocks * 5 / 6 : 9 * tiles_est <= ctx->lat_max_blocks;
    const std::vector<uint32_t> cuts = cut_chunks(ctx, dims.data(), n);
    const uint32_t nchunks = (uint32_t)cuts.size() - 1;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = 0;
    std::vector<xpngb_ctx*> used(nchunks, nullptr);
    std::vector<uint32_t> cn(nchunks, 0);
    // staggered waves (batches only): the latency-bound context decode of one wave runs under the issue-bound walk of the wave before
    uint32_t stagger = lat ? 0u : (l2 ? ctx->dec_stagger2 : ctx->dec_stagger1);
    if (stagger >= ctx->pipe_lanes) stagger = 0;
    auto chunk_done = [&](uint32_t ci) -> int {   // wait for chunk ci, read its error flag
        xpngb_ctx* lane = used[ci];
        CK(cudaStreamSynchronize(lane->stream));
        int herr = 0; memcpy(&herr, (uint8_t*)lane->pin_b.p + cn[ci] * (sizeof(DecImage) + 8), 4);
        if (herr) FAIL("corrupt .xpng data (decoder error %d)", herr);
        return 0;
    };
    uint32_t done = 0;
    for (uint32_t ci = 0; ci < nchunks && !rc; ci++) {
        const uint32_t li = ci % ctx->pipe_lanes;
        while (!rc && ci >= ctx->pipe_lanes && done <= ci - ctx->pipe_lanes) rc = chunk_done(done++);
        if (rc) break;
        xpngb_ctx* lane = lane_get(ctx, li);
        if (!lane) { snprintf(ctx->err, sizeof ctx->err, "cannot create pipeline lane %u", li); rc = 1; break; }
        used[ci] = lane; cn[ci] = cuts[ci + 1] - cuts[ci];
        const uint32_t i0 = cuts[ci], m = cn[ci];
        if (lane != ctx && cudaStreamWaitEvent(lane->stream, ctx->ev0, 0) != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaStreamWaitEvent failed"); rc = 1; break; }
        if (!files_on_device) {
            uint64_t lo = ~0ull, hi = 0;
            for (uint32_t i = i0; i < i0 + m; i++) { if (file_offsets[i] < lo) lo = file_offsets[i]; if (file_offsets[i] + file_sizes[i] > hi) hi = file_offsets[i] + file_sizes[i]; }
            if (cudaMemcpyAsync((uint8_t*)ctx->files.p + lo, (const uint8_t*)files + lo, hi - lo, cudaMemcpyHostToDevice, lane->stream) != cudaSuccess) {
                snprintf(ctx->err, sizeof ctx->err, "host to device copy failed"); rc = 1; break;
            }
        }
        if (stagger && ci >= stagger && cudaStreamWaitEvent(lane->stream, used[ci - stagger]->ev_stage, 0) != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaStreamWaitEvent failed"); rc = 1; break; }
        rc = decode_issue(lane, lat, imgs + i0, m, hdr.data() + 2 * i0, din, file_offsets + i0, file_sizes + i0, dpx, pixels_cap, (uint32_t)(tiles_est > 0xFFFFFFFFull ? 0xFFFFFFFFu : tiles_est));
        if (!rc && !pixels_on_device) {
            uint64_t lo = ~0ull, hi = 0;
            for (uint32_t i = i0; i < i0 + m; i++) {
                const uint64_t a = imgs[i].offset, b = a + imgs[i].w * imgs[i].h * (3 + imgs[i].A);
                if (a < lo) lo = a; if (b > hi) hi = b;
            }
            if (cudaMemcpyAsync((uint8_t*)pixels + lo, dpx + lo, hi - lo, cudaMemcpyDeviceToHost, lane->stream) != cudaSuccess) {
                snprintf(ctx->err, sizeof ctx->err, "device to host copy failed"); rc = 1;
            }
        }
    }
    // every issued chunk is waited for, also after an error (the caller's buffers must be quiet when we return)
    for (; done < nchunks; done++) {
        if (!used[done]) continue;
        if (rc) cudaStreamSynchronize(used[done]->stream);
        else rc = chunk_done(done);
    }
    for (uint32_t li = 1; li <= ctx->lanes.size(); li++) {
        xpngb_ctx* lane = ctx->lanes[li - 1];
        cudaEventRecord(lane->ev_done, lane->stream);
        cudaStreamWaitEvent(ctx->stream, lane->ev_done, 0);
    }
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (rc) return 1;
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    tl_dump(ctx, "dec");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// YCoCg-R side component
// ------------------------------------------------------------------------------------------------
extern "C" int xpngb_ycocg_forward(xpngb_ctx* ctx, const uint8_t* rgb, int16_t* ycc, uint64_t n) {
    if (!ctx || !rgb || !ycc) return 1;
    CK(cudaSetDevice(ctx->device));
    ENSURE(m2a, 3 * n); ENSURE(m2b, 6 * n);
    CK(cudaMemcpyAsync(ctx->m2a.p, rgb, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_ycocg_fwd, 1184, 256, 0, (const uint8_t*)ctx->m2a.p, (int16_t*)ctx->m2b.p, n);
    CK(cudaMemcpyAsync(ycc, ctx->m2b.p, 6 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int xpngb_ycocg_inverse(xpngb_ctx* ctx, const int16_t* ycc, uint8_t* rgb, uint64_t n) {
    if (!ctx || !rgb || !ycc) return 1;
    CK(cudaSetDevice(ctx->device));
    ENSURE(m2a, 3 * n); ENSURE(m2b, 6 * n);
    CK(cudaMemcpyAsync(ctx->m2b.p, ycc, 6 * n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_ycocg_inv, 1184, 256, 0, (const int16_t*)ctx->m2b.p, (uint8_t*)ctx->m2a.p, n);
    CK(cudaMemcpyAsync(rgb, ctx->m2a.p, 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
