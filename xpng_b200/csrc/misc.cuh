// misc.cuh — whole-image helpers: alpha normalisation scan/fix (libxpng.c:688-721), whole-image
// single-colour test (libxpng.c:741-753), stored (level 7) files, header gather, YCoCg-R side kernel.
#pragma once
#include "common.cuh"

namespace xpb {

enum ScanFlags { SCAN_DIRTY = 1, SCAN_TRANSLUCENT = 2, SCAN_NOT_SINGLE = 4 };

// grid = (chunks, nimg).  Pixels are compared in their NORMALISED form (alpha == 0 -> 0x00000000), which
// is what the reference sees after normalize_RGBA, so one pass answers all three questions.
__global__ void __launch_bounds__(256) k_image_scan(const ImageDesc* __restrict__ imgs, uint32_t* __restrict__ flags) {
    const ImageDesc I = imgs[blockIdx.y];
    const uint64_t npx = (uint64_t)I.w * I.h;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(I.px_off);
    uint32_t f = 0;
    if (I.pxsz == 4) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        uint32_t first = q[0]; if ((first >> 24) == 0) first = 0;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (uint64_t)gridDim.x * blockDim.x) {
            uint32_t v = q[i];
            const uint32_t a = v >> 24;
            if (a == 0 && v != 0) { f |= SCAN_DIRTY; v = 0; }
            if (a != 255) f |= SCAN_TRANSLUCENT;
            if (v != first) f |= SCAN_NOT_SINGLE;
        }
    } else {
        const uint32_t first = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        const uint64_t nq = 3 * npx >= 16 ? (3 * npx - 16) / 12 + 1 : 0;   // groups of four pixels whose 16-byte window stays inside the image
        for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nq; g += (uint64_t)gridDim.x * blockDim.x) {
            uint32_t v[4]; ld_rgb4(p + 12 * g, v);
            if ((v[0] != first) | (v[1] != first) | (v[2] != first) | (v[3] != first)) f |= SCAN_NOT_SINGLE;
        }
        if (blockIdx.x == 0 && (nq << 2) + threadIdx.x < npx) {  // tail (at most 7 pixels)
            const uint64_t i = (nq << 2) + threadIdx.x;
            const uint32_t v = (uint32_t)p[3 * i] | ((uint32_t)p[3 * i + 1] << 8) | ((uint32_t)p[3 * i + 2] << 16);
            if (v != first) f |= SCAN_NOT_SINGLE;
        }
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags + blockIdx.y, f);
}

// dst[i] = alpha ? src[i] : 0   (libxpng.c:699-706)
__global__ void __launch_bounds__(256) k_alpha_zero(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t npx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = src[i];
        dst[i] = (v >> 24) ? v : 0u;
    }
}
// RGBA -> RGB when every alpha is 255 (libxpng.c:711-718)
__global__ void __launch_bounds__(256) k_alpha_strip(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t npx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = src[i];
        dst[3 * i] = (uint8_t)v; dst[3 * i + 1] = (uint8_t)(v >> 8); dst[3 * i + 2] = (uint8_t)(v >> 16);
    }
}

// Stored file (level 7, libxpng.c:738-739): 8-byte header + raw bytes.  grid = (chunks, nimg).
struct StoreOut { uint64_t off; };
__global__ void __launch_bounds__(256) k_store7(const ImageDesc* __restrict__ imgs, const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out) {
    const ImageDesc I = imgs[blockIdx.y];
    const uint8_t* src = reinterpret_cast<const uint8_t*>(I.px_off);
    uint8_t* file = out + out_off[blockIdx.y];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st32u(file, (I.w - 1) | (7u << 24));
        st32u(file + 4, (I.h - 1) | ((I.pxsz - 3) << 24));
    }
    const uint64_t nw = I.raw_size >> 3;   // file + 8 is 8-aligned (files start 16-aligned), src is 16-aligned
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src);
    unsigned long long* d8 = reinterpret_cast<unsigned long long*>(file + 8);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) d8[i] = s8[i];
    if (blockIdx.x == 0) for (uint64_t i = (nw << 3) + threadIdx.x; i < I.raw_size; i += blockDim.x) file[8 + i] = src[i];
}

// Stored file -> pixels (level 7 decode, libxpng.c:974): a flat copy.  grid = (chunks, nimg); only images with
// copy_from[i] != ~0 are copied (byte offset of the file inside `in`).  Source (file + 8) is 8-aligned.
__global__ void __launch_bounds__(256) k_load7(const ImageDesc* __restrict__ imgs, const uint64_t* __restrict__ copy_from, const uint8_t* __restrict__ in) {
    const uint64_t from = copy_from[blockIdx.y];
    if (from == ~0ull) return;
    const ImageDesc I = imgs[blockIdx.y];
    uint8_t* dst = reinterpret_cast<uint8_t*>(I.px_off);
    const uint8_t* src = in + from + 8;
    const uint64_t nw = I.raw_size >> 3;
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src);
    unsigned long long* d8 = reinterpret_cast<unsigned long long*>(dst);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) d8[i] = __ldg(s8 + i);
    if (blockIdx.x == 0) for (uint64_t i = (nw << 3) + threadIdx.x; i < I.raw_size; i += blockDim.x) dst[i] = src[i];
}

// First two header words of n files -> hdr[n][2]
__global__ void k_gather_headers(const uint8_t* __restrict__ files, const uint64_t* __restrict__ offs, uint32_t n, uint32_t* __restrict__ hdr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hdr[2 * i] = ld32u(files + offs[i]); hdr[2 * i + 1] = ld32u(files + offs[i] + 4);
}

// YCoCg-R lifting (Tell_Me_Why/YCoCg-R.c:22, :31) — side component, not on the .xpng path.
__global__ void __launch_bounds__(256) k_ycocg_fwd(const uint8_t* __restrict__ rgb, int16_t* __restrict__ ycc, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int R = rgb[3 * i], G = rgb[3 * i + 1], B = rgb[3 * i + 2];
        const int co = R - B, t = B + (co >> 1), cg = G - t, y = t + (cg >> 1);
        ycc[3 * i] = (int16_t)y; ycc[3 * i + 1] = (int16_t)co; ycc[3 * i + 2] = (int16_t)cg;
    }
}
__global__ void __launch_bounds__(256) k_ycocg_inv(const int16_t* __restrict__ ycc, uint8_t* __restrict__ rgb, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int y = ycc[3 * i], co = ycc[3 * i + 1], cg = ycc[3 * i + 2];
        const int t = y - (cg >> 1), g = cg + t, b = t - (co >> 1), r = b + co;
        rgb[3 * i] = (uint8_t)r; rgb[3 * i + 1] = (uint8_t)g; rgb[3 * i + 2] = (uint8_t)b;
    }
}

}  // namespace xpb
