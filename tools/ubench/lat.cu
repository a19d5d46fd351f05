// Latency micro-benchmarks for the serial-chain kernels (one warp, dependent chains, clock64).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 8192
__global__ void k_shfl(uint32_t* out, long long* cyc) {
    uint32_t v = (threadIdx.x * 7 + 3) & 31;   // permutation step
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __shfl_sync(0xffffffffu, v, v);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[0] = t1 - t0;
}
__global__ void k_shfl_and(uint32_t* out, long long* cyc) {
    uint32_t v = (threadIdx.x * 7 + 3) & 31 | 0x40;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __shfl_sync(0xffffffffu, v, v & 15) + 0;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[1] = t1 - t0;
}
__global__ void k_lds(uint32_t* out, long long* cyc) {
    __shared__ uint32_t s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = ((i * 37 + 11) & 1023) * 4;
    __syncwarp();
    uint32_t off = threadIdx.x * 4;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) off = *(volatile uint32_t*)((char*)s + off);
    long long t1 = clock64();
    out[threadIdx.x] = off; if (!threadIdx.x) cyc[2] = t1 - t0;
}
__global__ void k_lds64_and(uint32_t* out, long long* cyc) {
    __shared__ unsigned long long s[64];
    for (int i = threadIdx.x; i < 64; i += 32) s[i] = (unsigned long long)(((i * 5 + 3) & 15) * 8) | 0xABCD00ull;
    __syncwarp();
    uint32_t off = (threadIdx.x & 15) * 8;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { unsigned long long w = *(volatile unsigned long long*)((char*)s + off); off = (uint32_t)w & 0xF8u; }
    long long t1 = clock64();
    out[threadIdx.x] = off; if (!threadIdx.x) cyc[3] = t1 - t0;
}
__global__ void k_ldg(const uint32_t* g, uint32_t* out, long long* cyc) {
    uint32_t off = threadIdx.x;
    for (int i = 0; i < 64; i++) off = g[off];   // warm L1
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) off = g[off];
    long long t1 = clock64();
    out[threadIdx.x] = off; if (!threadIdx.x) cyc[4] = t1 - t0;
}
__global__ void k_imadwide(uint32_t* out, long long* cyc, uint32_t a) {
    unsigned long long x = threadIdx.x + 12345;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (unsigned long long)(uint32_t)x * a + (x >> 32);
    long long t1 = clock64();
    out[threadIdx.x] = (uint32_t)x; if (!threadIdx.x) cyc[5] = t1 - t0;
}
__global__ void k_mulhi64(uint32_t* out, long long* cyc, unsigned long long a) {
    unsigned long long x = threadIdx.x + 0x123456789ull;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __umul64hi(x, a) + 0x9000000000000000ull;
    long long t1 = clock64();
    out[threadIdx.x] = (uint32_t)x; if (!threadIdx.x) cyc[6] = t1 - t0;
}
// umul64hi with the four partial products independent of each other (the compiler's own sequence chains
// xh*rl -> xl*rh+that -> xh*rh+that, three dependent IMAD.WIDE)
__device__ __forceinline__ unsigned long long mulhi_par(unsigned long long x, uint32_t rl, uint32_t rh) {
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    const unsigned long long A = (unsigned long long)xh * rl, B = (unsigned long long)xl * rh, Cc = (unsigned long long)xl * rl,
                             D = (unsigned long long)xh * rh;
    const unsigned long long mid = (unsigned long long)(uint32_t)A + (uint32_t)B + (Cc >> 32);
    return D + (A >> 32) + (B >> 32) + (mid >> 32);
}
__global__ void k_mulhi_par(uint32_t* out, long long* cyc, unsigned long long a) {
    unsigned long long x = threadIdx.x + 0x123456789ull;
    const uint32_t rl = (uint32_t)a, rh = (uint32_t)(a >> 32);
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = mulhi_par(x, rl, rh) + 0x9000000000000000ull;
    long long t1 = clock64();
    out[threadIdx.x] = (uint32_t)x; if (!threadIdx.x) cyc[12] = t1 - t0;
}
// same with add.cc chains written out
__device__ __forceinline__ unsigned long long mulhi_ptx(unsigned long long x, uint32_t rl, uint32_t rh) {
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint32_t dl, dh;
    asm("{\n\t.reg .u32 al, ah, bl, bh, ch, m, c1;\n\t"
        "mul.lo.u32 al, %3, %4;\n\tmul.hi.u32 ah, %3, %4;\n\t"       // A = xh*rl
        "mul.lo.u32 bl, %2, %5;\n\tmul.hi.u32 bh, %2, %5;\n\t"       // B = xl*rh
        "mul.hi.u32 ch, %2, %4;\n\t"                                   // hi(C) = hi(xl*rl)
        "mul.lo.u32 %0, %3, %5;\n\tmul.hi.u32 %1, %3, %5;\n\t"       // D = xh*rh
        "add.cc.u32 m, al, bl;\n\taddc.u32 c1, 0, 0;\n\t"
        "add.cc.u32 m, m, ch;\n\taddc.u32 c1, c1, 0;\n\t"
        "add.cc.u32 %0, %0, ah;\n\taddc.u32 %1, %1, 0;\n\t"
        "add.cc.u32 bh, bh, c1;\n\taddc.u32 c1, 0, 0;\n\t"
        "add.cc.u32 %0, %0, bh;\n\taddc.u32 %1, %1, c1;\n\t}"
        : "=&r"(dl), "=&r"(dh) : "r"(xl), "r"(xh), "r"(rl), "r"(rh));
    return ((unsigned long long)dh << 32) | dl;
}
__global__ void k_mulhi_ptx(uint32_t* out, long long* cyc, unsigned long long a) {
    unsigned long long x = threadIdx.x + 0x123456789ull;
    const uint32_t rl = (uint32_t)a, rh = (uint32_t)(a >> 32);
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = mulhi_ptx(x, rl, rh) + 0x9000000000000000ull;
    long long t1 = clock64();
    out[threadIdx.x] = (uint32_t)x; if (!threadIdx.x) cyc[13] = t1 - t0;
}
__global__ void k_mulhi_check(unsigned long long a, uint32_t* bad) {
    unsigned long long x = (threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + blockIdx.x * 0xD1B54A32D192ED03ull;
    const uint32_t rl = (uint32_t)a, rh = (uint32_t)(a >> 32);
    for (int i = 0; i < 4096; i++) {
        const unsigned long long w = __umul64hi(x, a);
        if (mulhi_par(x, rl, rh) != w || mulhi_ptx(x, rl, rh) != w) atomicAdd(bad, 1);
        x = x * 6364136223846793005ull + 1442695040888963407ull + w;
    }
}
__global__ void k_redux(uint32_t* out, long long* cyc) {
    uint32_t v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __reduce_or_sync(0xffffffffu, (threadIdx.x == (v & 31)) ? (v + 1) : 0);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[7] = t1 - t0;
}
__global__ void k_ballot(uint32_t* out, long long* cyc) {
    uint32_t v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __ballot_sync(0xffffffffu, threadIdx.x >= (v & 31));
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[8] = t1 - t0;
}
__global__ void k_iadd(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = (v ^ a) + (v >> 3);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[9] = t1 - t0;
}
__global__ void k_popc(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t v = threadIdx.x + a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __popc(v) + a;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (!threadIdx.x) cyc[10] = t1 - t0;
}
// switch-based dynamic dispatch (9 cases), chain through the switch
__global__ void k_switch(const uint8_t* g, uint32_t* out, long long* cyc) {
    __shared__ uint8_t s[9][256];
    for (int i = threadIdx.x; i < 9 * 256; i += 32) s[i / 256][i % 256] = g[i];
    __syncwarp();
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0, p8 = 0, c = 0;
    long long t0 = clock64();
    for (int i = 0; i < 1024; i++) {
        switch (c) {
            case 0: c = s[0][p0++ & 255]; break; case 1: c = s[1][p1++ & 255]; break; case 2: c = s[2][p2++ & 255]; break;
            case 3: c = s[3][p3++ & 255]; break; case 4: c = s[4][p4++ & 255]; break; case 5: c = s[5][p5++ & 255]; break;
            case 6: c = s[6][p6++ & 255]; break; case 7: c = s[7][p7++ & 255]; break; default: c = s[8][p8++ & 255]; break;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = c + p0 + p1 + p2 + p3 + p4 + p5 + p6 + p7 + p8; if (!threadIdx.x) cyc[11] = (t1 - t0) * (N / 1024);
}
int main() {
    uint32_t* out; long long* cyc; uint32_t* g; uint8_t* g8;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 128 * 8); cudaMalloc(&g, 4096 * 4); cudaMalloc(&g8, 9 * 256);
    uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = (i * 37 + 11) & 1023;
    cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    uint8_t h8[9 * 256]; for (int i = 0; i < 9 * 256; i++) h8[i] = (uint8_t)((i * 7 + i / 256 + (i >> 3)) % 9);
    cudaMemcpy(g8, h8, sizeof h8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
        k_shfl<<<1, 32>>>(out, cyc); k_shfl_and<<<1, 32>>>(out, cyc); k_lds<<<1, 32>>>(out, cyc); k_lds64_and<<<1, 32>>>(out, cyc);
        k_ldg<<<1, 32>>>(g, out, cyc); k_imadwide<<<1, 32>>>(out, cyc, 77777); k_mulhi64<<<1, 32>>>(out, cyc, 0xF123456789ABCDEFull);
        k_redux<<<1, 32>>>(out, cyc); k_ballot<<<1, 32>>>(out, cyc); k_iadd<<<1, 32>>>(out, cyc, 5); k_popc<<<1, 32>>>(out, cyc, 3);
        k_switch<<<1, 32>>>(g8, out, cyc);
        k_mulhi_par<<<1, 32>>>(out, cyc, 0xF123456789ABCDEFull); k_mulhi_ptx<<<1, 32>>>(out, cyc, 0xF123456789ABCDEFull);
        cudaDeviceSynchronize();
    }
    const char* names[] = { "shfl(idx=v)", "shfl(idx=v&15)", "lds32 chase", "lds64+and chase", "ldg L1 chase", "imad.wide+shift", "umul64hi+add",
                            "redux.or(sel)", "ballot(cmp)", "xor+shr+add (3 alu)", "popc+add", "switch9+lds", "mulhi64 parallel (C)", "mulhi64 parallel (PTX)" };
    {
        uint32_t* bad; cudaMallocManaged(&bad, 4); *bad = 0;
        k_mulhi_check<<<64, 256>>>(0xF123456789ABCDEFull, bad); k_mulhi_check<<<64, 256>>>(0x8000000000000001ull, bad);
        k_mulhi_check<<<64, 256>>>(0xFFFFFFFFFFFFFFFFull, bad); cudaDeviceSynchronize();
        printf("mulhi variants mismatches: %u\n", *bad);
    }
    for (int i = 0; i < 14; i++) printf("%-22s %7.2f cycles/iter\n", names[i], (double)cyc[i] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
