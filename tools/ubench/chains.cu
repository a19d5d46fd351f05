// Micro-benchmarks of the serial-chain inner loops on synthetic data (cycles per symbol / step).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define RING 2048
// ------------------------------------------------------------------ rANS decode chain, one-level table
// entry layout: bias (14) | sym (4) << 14 | freq (14) << 18
template <int ACTIVE>   // lanes that run the chain (1 or 32)
__global__ void k_dec_d1(const uint32_t* lut_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* ring = sm + (1u << pb);
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < (1u << pb); i += 32) lut[i] = lut_g[i];
    for (uint32_t i = lane; i < RING; i += 32) ring[i] = words[i % nwords];
    __syncwarp();
    if (lane >= ACTIVE) return;
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t alo = 0x12345678u, ahi = 0x1u, blo = 0x9abcdef0u, bhi = 0x2u, kb = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    long long t0 = clock64();
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            uint32_t ea, eb;
            { const uint32_t slot = alo & mask; ea = lut[slot];
              const uint32_t qlo = __funnelshift_r(alo, ahi, pb), qhi = ahi >> pb, f = ea >> 18, b = ea & 0x3FFFu;
              const unsigned long long t = (unsigned long long)f * qlo + b; alo = (uint32_t)t; ahi = f * qhi + (uint32_t)(t >> 32); }
            { const uint32_t slot = blo & mask; eb = lut[slot];
              const uint32_t qlo = __funnelshift_r(blo, bhi, pb), qhi = bhi >> pb, f = eb >> 18, b = eb & 0x3FFFu;
              const unsigned long long t = (unsigned long long)f * qlo + b; blo = (uint32_t)t; bhi = f * qhi + (uint32_t)(t >> 32); }
            { const bool p = (ahi | (alo & 0x80000000u)) == 0; const uint32_t w = ring[kb]; ahi = p ? alo : ahi; alo = p ? w : alo; kb = p ? ((kb + 1) & (RING - 1)) : kb; }
            { const bool p = (bhi | (blo & 0x80000000u)) == 0; const uint32_t w = ring[kb]; bhi = p ? blo : bhi; blo = p ? w : blo; kb = p ? ((kb + 1) & (RING - 1)) : kb; }
            keep |= (ea & mk[j]) | (eb & mk[j + 1]);
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; out[n] = (uint8_t)(alo + blo + ahi + bhi); }
}


// D2: ring candidates for a round are loaded at the start of the round (k known from the previous
// round), so the renormalisation word never waits for the other state's predicate.
__global__ void k_dec_d2(const uint32_t* lut_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* ring = sm + (1u << pb);
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < (1u << pb); i += 32) lut[i] = lut_g[i];
    for (uint32_t i = lane; i < RING + 8; i += 32) ring[i] = words[(i % RING) % nwords];
    __syncwarp();
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t alo = 0x12345678u, ahi = 0x1u, blo = 0x9abcdef0u, bhi = 0x2u, kb = 0;   // kb: byte offset into the ring
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    const char* ringb = reinterpret_cast<const char*>(ring);
    long long t0 = clock64();
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb), c1 = *reinterpret_cast<const uint32_t*>(ringb + kb + 4);
            uint32_t ea, eb;
            { const uint32_t slot = alo & mask; ea = lut[slot];
              const uint32_t qlo = __funnelshift_r(alo, ahi, pb), qhi = ahi >> pb, f = ea >> 18, b = ea & 0x3FFFu;
              const unsigned long long t = (unsigned long long)f * qlo + b; alo = (uint32_t)t; ahi = f * qhi + (uint32_t)(t >> 32); }
            { const uint32_t slot = blo & mask; eb = lut[slot];
              const uint32_t qlo = __funnelshift_r(blo, bhi, pb), qhi = bhi >> pb, f = eb >> 18, b = eb & 0x3FFFu;
              const unsigned long long t = (unsigned long long)f * qlo + b; blo = (uint32_t)t; bhi = f * qhi + (uint32_t)(t >> 32); }
            const bool pa = (ahi | (alo & 0x80000000u)) == 0, pbb = (bhi | (blo & 0x80000000u)) == 0;
            const uint32_t wb = pa ? c1 : c0;
            ahi = pa ? alo : ahi; alo = pa ? c0 : alo;
            bhi = pbb ? blo : bhi; blo = pbb ? wb : blo;
            kb = (kb + (pa ? 4u : 0u) + (pbb ? 4u : 0u)) & (RING * 4 - 1);
            keep |= (ea & mk[j]) | (eb & mk[j + 1]);
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; out[n] = (uint8_t)(alo + blo + ahi + bhi); }
}


__global__ void k_dec_d2one(const uint32_t* lut_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* ring = sm + (1u << pb);
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < (1u << pb); i += 32) lut[i] = lut_g[i];
    for (uint32_t i = lane; i < RING + 8; i += 32) ring[i] = words[(i % RING) % nwords];
    __syncwarp();
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t alo = 0x12345678u, ahi = 0x1u, blo = 0x9abcdef0u, bhi = 0x2u, kb = 0;   // kb: byte offset into the ring
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    const char* ringb = reinterpret_cast<const char*>(ring);
    long long t0 = clock64();
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb), c1 = *reinterpret_cast<const uint32_t*>(ringb + kb + 4);
            uint32_t ea, eb;
            { const uint32_t slot = alo & mask; ea = lut[slot];
              const uint32_t qlo = __funnelshift_r(alo, ahi, pb), qhi = ahi >> pb, f = ea >> 18, b = ea & 0x3FFFu;
              const unsigned long long t = (unsigned long long)f * qlo + b; alo = (uint32_t)t; ahi = f * qhi + (uint32_t)(t >> 32); }
            eb = 0;
            const bool pa = (ahi | (alo & 0x80000000u)) == 0, pbb = false;
            const uint32_t wb = pa ? c1 : c0;
            ahi = pa ? alo : ahi; alo = pa ? c0 : alo;
            bhi = pbb ? blo : bhi; blo = pbb ? wb : blo;
            kb = (kb + (pa ? 4u : 0u) + (pbb ? 4u : 0u)) & (RING * 4 - 1);
            keep |= (ea & mk[j]) | (eb & mk[j + 1]);
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; out[n] = (uint8_t)(alo + blo + ahi + bhi); }
}



// D4: as D2, but chain B is skewed by half a round in program order: B's table lookup is issued before A's
// multiply, and B's multiply before A's renormalisation.
__global__ void k_dec_d4(const uint32_t* lut_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* ring = sm + (1u << pb);
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < (1u << pb); i += 32) lut[i] = lut_g[i];
    for (uint32_t i = lane; i < RING + 8; i += 32) ring[i] = words[(i % RING) % nwords];
    __syncwarp();
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t alo = 0x12345678u, ahi = 0x1u, blo = 0x9abcdef0u, bhi = 0x2u, kb = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) { mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    const char* ringb = reinterpret_cast<const char*>(ring);
    long long t0 = clock64();
    // prologue: A's lookup
    uint32_t ea = lut[alo & mask];
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            // B lookup issued first (its state is final since the previous round)
            const uint32_t eb = lut[blo & mask];
            // A multiply + renorm (word = ring[kb])
            {   const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb);
                const uint32_t qlo = __funnelshift_r(alo, ahi, pb), qhi = ahi >> pb, f = ea >> 18, b = ea & 0x3FFFu;
                const unsigned long long t = (unsigned long long)f * qlo + b; alo = (uint32_t)t; ahi = f * qhi + (uint32_t)(t >> 32);
                const bool pa = (ahi | (alo & 0x80000000u)) == 0;
                ahi = pa ? alo : ahi; alo = pa ? c0 : alo; kb = (kb + (pa ? 4u : 0u)) & (RING * 4 - 1); }
            keep |= ea & mk[j];
            // A's next lookup
            ea = lut[alo & mask];
            // B multiply + renorm
            {   const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb);
                const uint32_t qlo = __funnelshift_r(blo, bhi, pb), qhi = bhi >> pb, f = eb >> 18, b = eb & 0x3FFFu;
                const unsigned long long t = (unsigned long long)f * qlo + b; blo = (uint32_t)t; bhi = f * qhi + (uint32_t)(t >> 32);
                const bool pq = (bhi | (blo & 0x80000000u)) == 0;
                bhi = pq ? blo : bhi; blo = pq ? c0 : blo; kb = (kb + (pq ? 4u : 0u)) & (RING * 4 - 1); }
            keep |= eb & mk[j + 1];
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; out[n] = (uint8_t)(alo + blo + ahi + bhi + ea); }
}


// D5: pair-SIMT decode: even lanes own state A, odd lanes state B (one instruction stream for both);
// a ballot per round tells everyone who renormalised; B's word is ring[k + pA].
template <int SPEC>
__global__ void k_dec_d5(const uint32_t* lut_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* ring = sm + (1u << pb);
    const uint32_t lane = threadIdx.x, h = lane & 1;
    for (uint32_t i = lane; i < (1u << pb); i += 32) lut[i] = lut_g[i];
    for (uint32_t i = lane; i < RING + 8; i += 32) ring[i] = words[(i % RING) % nwords];
    __syncwarp();
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t xlo = h ? 0x9abcdef0u : 0x12345678u, xhi = h ? 0x2u : 0x1u, kb = 0;
    uint32_t mk[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { mk[j] = (lane >> 1) == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    const char* ringb = reinterpret_cast<const char*>(ring);
    long long t0 = clock64();
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb), c1 = *reinterpret_cast<const uint32_t*>(ringb + kb + 4);
            const uint32_t slot = xlo & mask; const uint32_t e = lut[slot];
            const uint32_t qlo = __funnelshift_r(xlo, xhi, pb), qhi = xhi >> pb, f = e >> 18, b = e & 0x3FFFu;
            const unsigned long long t = (unsigned long long)f * qlo + b; const uint32_t lo = (uint32_t)t, hi = f * qhi + (uint32_t)(t >> 32);
            const bool p = (hi | (lo & 0x80000000u)) == 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, p);
            const uint32_t pa = bal & 1u, pq = (bal >> 1) & 1u;
            const uint32_t w = (h & pa) ? c1 : c0;
            xhi = p ? lo : hi; xlo = p ? w : lo;
            kb = (kb + 4u * (pa + pq)) & (RING * 4 - 1);
            keep |= e & mk[j];
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane < 2) { cyc[0] = t1 - t0; out[n + lane] = (uint8_t)(xlo + xhi); }
}

// D3: the renormalisation predicate comes from a second table: x' = f*q + b < 2^31  <=>  qhi == 0 && qlo <= T[slot],
// T[slot] = floor((2^31 - 1 - b) / f): the predicate no longer waits for the 64-bit multiply, and the next slot only
// needs the LOW word of x' (a 32-bit IMAD).
__global__ void k_dec_d3(const uint32_t* lut_g, const uint32_t* lutT_g, const uint32_t* words, uint32_t nwords, uint32_t n, int pb, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint32_t* lutT = sm + (1u << pb); uint32_t* ring = sm + (2u << pb);
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < (1u << pb); i += 32) { lut[i] = lut_g[i]; lutT[i] = lutT_g[i]; }
    for (uint32_t i = lane; i < RING + 8; i += 32) ring[i] = words[(i % RING) % nwords];
    __syncwarp();
    const uint32_t mask = (1u << pb) - 1u;
    uint32_t alo = 0x12345678u, ahi = 0x1u, blo = 0x9abcdef0u, bhi = 0x2u, kb = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) { mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    const char* ringb = reinterpret_cast<const char*>(ring);
    long long t0 = clock64();
    for (uint32_t i = 0; i + 32 <= n; i += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const uint32_t c0 = *reinterpret_cast<const uint32_t*>(ringb + kb), c1 = *reinterpret_cast<const uint32_t*>(ringb + kb + 4);
            uint32_t ea, eb; bool pa, pq;
            uint32_t alo2, ahi2, blo2, bhi2;
            { const uint32_t slot = alo & mask; ea = lut[slot]; const uint32_t T = lutT[slot];
              const uint32_t qlo = __funnelshift_r(alo, ahi, pb), qhi = ahi >> pb, f = ea >> 18, b = ea & 0x3FFFu;
              pa = (qhi == 0) & (qlo <= T);
              alo2 = f * qlo + b;                                             // low word: 32-bit IMAD
              ahi2 = __umulhi(f, qlo) + f * qhi + (alo2 < b ? 1u : 0u); }     // high word, off the slot path
            { const uint32_t slot = blo & mask; eb = lut[slot]; const uint32_t T = lutT[slot];
              const uint32_t qlo = __funnelshift_r(blo, bhi, pb), qhi = bhi >> pb, f = eb >> 18, b = eb & 0x3FFFu;
              pq = (qhi == 0) & (qlo <= T);
              blo2 = f * qlo + b;
              bhi2 = __umulhi(f, qlo) + f * qhi + (blo2 < b ? 1u : 0u); }
            const uint32_t wb = pa ? c1 : c0;
            ahi = pa ? alo2 : ahi2; alo = pa ? c0 : alo2;
            bhi = pq ? blo2 : bhi2; blo = pq ? wb : blo2;
            kb = (kb + (pa ? 4u : 0u) + (pq ? 4u : 0u)) & (RING * 4 - 1);
            keep |= (ea & mk[j]) | (eb & mk[j + 1]);
        }
        out[i + lane] = (uint8_t)((keep >> 14) & 15u);
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; out[n] = (uint8_t)(alo + blo + ahi + bhi); }
}

// issue-rate probes: independent instructions in one warp
__global__ void k_issue(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t v[8]; for (int i = 0; i < 8; i++) v[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < 4096; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (v[i] ^ a) + 1;   // LOP3+IADD fused? keep 2 ops
    }
    long long t1 = clock64();
    uint32_t s = 0; for (int i = 0; i < 8; i++) s += v[i];
    out[threadIdx.x] = s; if (!threadIdx.x) cyc[0] = t1 - t0;
}
__global__ void k_issue_mix(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t v[8]; for (int i = 0; i < 8; i++) v[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < 4096; it++) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) { v[i] = v[i] * a + 7; v[i + 1] = (v[i + 1] >> 3) ^ a; }   // IMAD | SHF+LOP
    }
    long long t1 = clock64();
    uint32_t s = 0; for (int i = 0; i < 8; i++) s += v[i];
    out[threadIdx.x] = s; if (!threadIdx.x) cyc[0] = t1 - t0;
}

// ------------------------------------------------------------------ context walk variants
// W_shfl: lane c owns stream c (64-bit window, prefetched next chunk), chain = one SHFL per step.
__global__ void k_walk_shfl(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    const uint32_t lane = threadIdx.x;
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint2* src = reinterpret_cast<const uint2*>(streams + soff[c]);
    uint32_t wlo = 0, whi = 0, nlo = 0, nhi = 0;
    if (nch > 0) { const uint2 q = __ldg(src); wlo = q.x; whi = q.y; }
    if (nch > 1) { const uint2 q = __ldg(src + 1); nlo = q.x; nhi = q.y; }
    uint32_t ch = 2, cnt = 8, info = wlo & 0xFFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            const bool own = lane == cur;
            const bool ref = own && cnt == 1u;
            const uint32_t plo = __funnelshift_r(wlo, whi, 8), phi = whi >> 8;
            wlo = own ? plo : wlo; whi = own ? phi : whi; cnt -= own ? 1u : 0u;
            if (__any_sync(0xffffffffu, ref)) {
                wlo = ref ? nlo : wlo; whi = ref ? nhi : whi; cnt = ref ? 8u : cnt;
                const uint32_t doload = (ref && ch < nch) ? 1u : 0u;
                nlo = ref ? 0u : nlo; nhi = ref ? 0u : nhi;
                asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.nc.v2.u32 {%0, %1}, [%2];\n}" : "+r"(nlo), "+r"(nhi) : "l"(src + ch), "r"(doload));
                ch += ref ? 1u : 0u;
            }
            info = wlo & 0xFFu;
            keep |= got & mk[j];
            cur = got;
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}


template <int VAR>
__global__ void k_walk_var(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    const uint32_t lane = threadIdx.x;
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint2* src = reinterpret_cast<const uint2*>(streams + soff[c]);
    uint32_t wlo = 0, whi = 0, nlo = 0, nhi = 0;
    if (nch > 0) { const uint2 q = __ldg(src); wlo = q.x; whi = q.y; }
    if (nch > 1) { const uint2 q = __ldg(src + 1); nlo = q.x; nhi = q.y; }
    uint32_t ch = 2, cnt = 8, info = wlo & 0xFFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            if (VAR >= 1) {
                const bool own = lane == cur;
                const uint32_t plo = __funnelshift_r(wlo, whi, 8), phi = whi >> 8;
                wlo = own ? plo : wlo; whi = own ? phi : whi;
                if (VAR >= 2) {
                    const bool ref = own && cnt == 1u;
                    cnt -= own ? 1u : 0u;
                    if (VAR == 2) {
                        if (__any_sync(0xffffffffu, ref)) {
                            wlo = ref ? nlo : wlo; whi = ref ? nhi : whi; cnt = ref ? 8u : cnt;
                            const uint32_t doload = (ref && ch < nch) ? 1u : 0u;
                            nlo = ref ? 0u : nlo; nhi = ref ? 0u : nhi;
                            asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.nc.v2.u32 {%0, %1}, [%2];\n}" : "+r"(nlo), "+r"(nhi) : "l"(src + ch), "r"(doload));
                            ch += ref ? 1u : 0u;
                        }
                    } else if (VAR == 3) {   // refill check only every 4th step, windows keep >= 4 spare (not exact; timing only)
                        if ((j & 3) == 3 && __any_sync(0xffffffffu, cnt <= 4u)) { wlo = cnt <= 4u ? nlo : wlo; whi = cnt <= 4u ? nhi : whi; cnt = cnt <= 4u ? 8u : cnt; }
                    }
                }
                info = wlo & 0xFFu;
            } else info = (info + 1) & 7u;
            keep |= got & mk[j];
            cur = VAR == 0 ? (got & 7u) : got;
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}


__device__ __forceinline__ uint32_t walk_pack8(uint2 q) {
    uint32_t a = (q.x | (q.x >> 4)) & 0x00FF00FFu; a = (a | (a >> 8)) & 0xFFFFu;
    uint32_t b = (q.y | (q.y >> 4)) & 0x00FF00FFu; b = (b | (b >> 8)) & 0xFFFFu;
    return a | (b << 16);
}
template <int VAR>
__global__ void k_walk_nib(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    const uint32_t lane = threadIdx.x;
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint2* src = reinterpret_cast<const uint2*>(streams + soff[c]);
    auto chunk = [&](uint32_t k) -> uint2 { return k < nch ? __ldg(src + k) : make_uint2(0u, 0u); };
    uint32_t wlo = walk_pack8(chunk(0)), whi = walk_pack8(chunk(1)), cnt = 16;
    uint32_t nbuf = walk_pack8(chunk(2));
    uint2 raw = chunk(3);
    uint32_t ch = 4;
    uint32_t info = wlo & 0xFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u;
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            const bool own = lane == cur;
            const uint32_t plo = __funnelshift_r(wlo, whi, 4), phi = whi >> 4;
            wlo = own ? plo : wlo; whi = own ? phi : whi; cnt -= own ? 1u : 0u;
            info = wlo & 0xFu;
            keep |= got & mk[j];
            cur = got;
            if ((j & 7) == 7) {
                const bool need = cnt <= 8u;
                const unsigned long long add = (unsigned long long)nbuf << (4u * min(cnt, 8u));
                wlo |= need ? (uint32_t)add : 0u; whi |= need ? (uint32_t)(add >> 32) : 0u;
                info = wlo & 0xFu;
                cnt += need ? 8u : 0u;
                if (VAR == 0) {
                    nbuf = need ? walk_pack8(raw) : nbuf;
                    const uint32_t doload = (need && ch < nch) ? 1u : 0u;
                    raw.x = need ? 0u : raw.x; raw.y = need ? 0u : raw.y;
                    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.nc.v2.u32 {%0, %1}, [%2];\n}"
                                 : "+r"(raw.x), "+r"(raw.y) : "l"(src + ch), "r"(doload));
                    ch += need ? 1u : 0u;
                } else { nbuf = need ? (nbuf * 5u + 1u) & 0x33333333u : nbuf; }   // VAR 1: no loads (timing only)
            }
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}


template <int LD>
__global__ void k_walk_nib2(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    const uint32_t lane = threadIdx.x;
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint2* src = reinterpret_cast<const uint2*>(streams + soff[c]);
    auto chunk = [&](uint32_t k) -> uint2 { return k < nch ? __ldg(src + k) : make_uint2(0u, 0u); };
    uint32_t wlo = walk_pack8(chunk(0)), whi = walk_pack8(chunk(1)), cnt = 16;
    uint32_t nbuf = walk_pack8(chunk(2));
    uint2 raw = chunk(3);
    uint32_t ch = 4;
    uint32_t info = wlo & 0xFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) { mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    if (lane < 9) for (uint32_t k = 0; k < 4; k++) asm volatile("prefetch.global.L1 [%0];" :: "l"((const char*)src + 128 * k));
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            const bool own = lane == cur;
            const uint32_t plo = __funnelshift_r(wlo, whi, 4), phi = whi >> 4;
            wlo = own ? plo : wlo; whi = own ? phi : whi; cnt -= own ? 1u : 0u;
            info = wlo & 0xFu;
            keep |= got & mk[j];
            cur = got;
            if ((j & 7) == 7) {
                const uint32_t nm = cnt <= 8u ? 0xFFFFFFFFu : 0u;            // need mask
                const unsigned long long add = (unsigned long long)(nbuf & nm) << (4u * min(cnt, 8u));
                wlo |= (uint32_t)add; whi |= (uint32_t)(add >> 32);
                info = wlo & 0xFu;
                cnt += nm & 8u;
                const uint32_t pk = walk_pack8(raw);
                nbuf = (pk & nm) | (nbuf & ~nm);
                const uint32_t doload = nm & (ch < nch ? 1u : 0u);
                raw.x &= ~nm; raw.y &= ~nm;
                if (LD == 0) asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.nc.v2.u32 {%0, %1}, [%2];\n @q prefetch.global.L1 [%2 + 512];\n}"
                             : "+r"(raw.x), "+r"(raw.y) : "l"(src + ch), "r"(doload));
                if (LD == 1) asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.ca.v2.u32 {%0, %1}, [%2];\n @q prefetch.global.L1 [%2 + 512];\n}"
                             : "+r"(raw.x), "+r"(raw.y) : "l"(src + ch), "r"(doload));
                if (LD == 2) asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.ca.v2.u32 {%0, %1}, [%2];\n}"
                             : "+r"(raw.x), "+r"(raw.y) : "l"(src + ch), "r"(doload));
                ch += nm & 1u;
            }
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}


__global__ void k_walk_nib3(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    const uint32_t lane = threadIdx.x;
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint2* src = reinterpret_cast<const uint2*>(streams + soff[c]);
    auto chunk = [&](uint32_t k) -> uint2 { return k < nch ? __ldg(src + k) : make_uint2(0u, 0u); };
    uint32_t wlo = walk_pack8(chunk(0)), whi = walk_pack8(chunk(1)), cnt = 16;
    uint32_t nbuf = walk_pack8(chunk(2));
    uint2 raw = chunk(3);
    uint32_t ch = 4, nm = 0;
    uint32_t info = wlo & 0xFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) { mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            const bool own = lane == cur;
            const uint32_t plo = __funnelshift_r(wlo, whi, 4), phi = whi >> 4;
            wlo = own ? plo : wlo; whi = own ? phi : whi; cnt -= own ? 1u : 0u;
            info = wlo & 0xFu;
            keep |= got & mk[j];
            cur = got;
            if ((j & 7) == 1) {          // append (cnt >= 1 here, so info is unaffected)
                nm = cnt <= 8u ? 0xFFFFFFFFu : 0u;
                const unsigned long long add = (unsigned long long)(nbuf & nm) << (4u * min(cnt, 8u));
                wlo |= (uint32_t)add; whi |= (uint32_t)(add >> 32);
                cnt += nm & 8u;
            }
            if ((j & 7) == 5) {          // advance the prefetch for the lanes that appended
                const uint32_t pk = walk_pack8(raw);
                nbuf = (pk & nm) | (nbuf & ~nm);
                const uint32_t doload = nm & (ch < nch ? 1u : 0u);
                raw.x &= ~nm; raw.y &= ~nm;
                asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.global.nc.v2.u32 {%0, %1}, [%2];\n}"
                             : "+r"(raw.x), "+r"(raw.y) : "l"(src + ch), "r"(doload));
                ch += nm & 1u;
            }
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}


// nib4: the tile's nine streams are packed to nibbles in shared memory up front (whole warp), the walk
// refills its windows with one shared load (8 symbols = one word) -- no global latency inside the walk.
__global__ void k_walk_nib4(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    extern __shared__ uint32_t sm[];
    const uint32_t lane = threadIdx.x;
    // word offsets of the packed streams
    uint32_t wo[10]; wo[0] = 0;
    for (int k = 0; k < 9; k++) wo[k + 1] = wo[k] + (slen[k] + 7) / 8 + 1;
    for (int k = 0; k < 9; k++) {
        const uint2* s = reinterpret_cast<const uint2*>(streams + soff[k]);
        const uint32_t nw = (slen[k] + 7) / 8;
        for (uint32_t i = lane; i < nw; i += 32) sm[wo[k] + i] = walk_pack8(__ldg(s + i));
        if (lane == 0) sm[wo[k] + nw] = 0;
    }
    __syncwarp();
    const uint32_t c = lane < 9 ? lane : 0;
    const uint32_t nch = lane < 9 ? (slen[c] + 7) / 8 : 0;
    const uint32_t* src = sm + wo[c];
    auto chunk = [&](uint32_t k) -> uint32_t { return k < nch ? src[k] : 0u; };
    uint32_t wlo = chunk(0), whi = chunk(1), cnt = 16, nbuf = chunk(2), ch = 3;
    uint32_t info = wlo & 0xFu, cur = 0;
    uint32_t mk[32];
#pragma unroll
    for (int j = 0; j < 32; j++) { mk[j] = lane == (uint32_t)j ? 0xFFFFFFFFu : 0u; asm volatile("" : "+r"(mk[j])); }
    long long t0 = clock64();
    for (uint32_t pos = 0; pos < m; pos += 32) {
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t got = __shfl_sync(0xffffffffu, info, cur);
            const bool own = lane == cur;
            const uint32_t plo = __funnelshift_r(wlo, whi, 4), phi = whi >> 4;
            wlo = own ? plo : wlo; whi = own ? phi : whi; cnt -= own ? 1u : 0u;
            info = wlo & 0xFu;
            keep |= got & mk[j];
            cur = got;
            if ((j & 7) == 1) {
                const uint32_t nm = cnt <= 8u ? 0xFFFFFFFFu : 0u;
                const unsigned long long add = (unsigned long long)(nbuf & nm) << (4u * min(cnt, 8u));
                wlo |= (uint32_t)add; whi |= (uint32_t)(add >> 32);
                cnt += nm & 8u;
                const uint32_t nx = src[min(ch, nch)];      // slot nch holds 0
                nbuf = (nx & nm) | (nbuf & ~nm);
                ch += nm & 1u;
            }
        }
        if (pos + lane < m) out[pos + lane] = (uint8_t)keep;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
}

// W_smem: single lane, per-context 64-bit window in shared memory whose bytes are the byte offsets of
// the next window (sym * 8); 0xFF = empty -> refill from the stream.  Chain = LDS + LOP.
__global__ void k_walk_smem(const uint8_t* streams, const uint32_t* soff, const uint32_t* slen, uint32_t m, uint8_t* out, long long* cyc) {
    __shared__ unsigned long long W[16];
    __shared__ uint32_t P[16];
    const uint32_t lane = threadIdx.x;
    if (lane < 16) { W[lane] = ~0ull; P[lane] = 0; }
    __syncwarp();
    if (lane) return;
    auto refill = [&](uint32_t off) -> unsigned long long {
        const uint32_t c = off >> 3, p = P[c]; P[c] = p + 7;
        unsigned long long w = 0xFF00000000000000ull;
        for (int k = 0; k < 7; k++) { const uint32_t s = (c < 9 && p + k < slen[c]) ? streams[soff[c] + p + k] : 0; w |= (unsigned long long)(s * 8u) << (8 * k); }
        return w;
    };
    long long t0 = clock64();
    uint32_t off = 0, acc = 0;
    for (uint32_t pos = 0; pos < m; pos += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned long long w = *(volatile unsigned long long*)((char*)W + off);
            uint32_t b = (uint32_t)w & 0xFFu;
            if (b == 0xFFu) { w = refill(off); b = (uint32_t)w & 0xFFu; }
            *(volatile unsigned long long*)((char*)W + off) = (w >> 8) | 0xFF00000000000000ull;
            acc = (acc >> 8) | (b << 21);
            off = b;
        }
        *reinterpret_cast<uint32_t*>(out + pos) = acc;
    }
    long long t1 = clock64();
    cyc[0] = t1 - t0;
}

// ------------------------------------------------------------------ rANS encode chain (2 states, branch-free)
__global__ void k_enc_e1(const uint8_t* sym, uint32_t n, const uint4* tab_g, int pb, uint32_t* wout, long long* cyc, int nsym) {
    __shared__ uint4 tab[16];
    const uint32_t lane = threadIdx.x;
    if (lane < 16) tab[lane] = tab_g[lane];
    __syncwarp();
    if (lane) return;
    unsigned long long x0 = 1ull << 31, x1 = 1ull << 31;
    uint32_t* wp = wout;
    long long t0 = clock64();
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(sym);
    for (uint32_t i = 0; i + 4 <= n; i += 4) {
        const uint32_t v = s4[i >> 2];
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            const uint4 e0 = tab[(v >> (8 * k)) & 15u], e1 = tab[(v >> (8 * k + 8)) & 15u];
            {   // x = rcp lo, y = rcp hi, z = bias | cmpl << 16, w = xmax_hi | shift << 24 ... keep simple
                const bool p = (uint32_t)(x0 >> 32) >= (e0.w & 0xFFFFFFu); *wp = (uint32_t)x0; wp += p; x0 = p ? x0 >> 32 : x0;
                const unsigned long long q = __umul64hi(x0, (unsigned long long)e0.x | ((unsigned long long)e0.y << 32)) >> (e0.w >> 24);
                x0 += (e0.z & 0xFFFFu) + q * (e0.z >> 16); }
            {   const bool p = (uint32_t)(x1 >> 32) >= (e1.w & 0xFFFFFFu); *wp = (uint32_t)x1; wp += p; x1 = p ? x1 >> 32 : x1;
                const unsigned long long q = __umul64hi(x1, (unsigned long long)e1.x | ((unsigned long long)e1.y << 32)) >> (e1.w >> 24);
                x1 += (e1.z & 0xFFFFu) + q * (e1.z >> 16); }
        }
    }
    long long t1 = clock64();
    cyc[0] = t1 - t0; wp[0] = (uint32_t)x0; wp[1] = (uint32_t)x1; cyc[1] = wp - wout;
}


// ------------------------------------------------------------------ E2: two lanes per block (lane h owns state h)
__device__ __forceinline__ unsigned long long mulw(uint32_t a, uint32_t b) {
    unsigned long long r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
// high 64 bits of a 64 x 64 product: four independent wide multiplies, then a three-level carry tree
__device__ __forceinline__ unsigned long long mulhi64_tree(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const unsigned long long p00 = mulw(a0, b0), p01 = mulw(a0, b1), p10 = mulw(a1, b0), p11 = mulw(a1, b1);
    const unsigned long long mid = (p00 >> 32) + (uint32_t)p01 + (uint32_t)p10;
    return p11 + (p01 >> 32) + (p10 >> 32) + (mid >> 32);
}
__device__ __forceinline__ unsigned long long mulhi64_flat(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const unsigned long long p00 = (unsigned long long)a0 * b0, p01 = (unsigned long long)a0 * b1, p10 = (unsigned long long)a1 * b0,
                             p11 = (unsigned long long)a1 * b1;
    const unsigned long long mid = (p00 >> 32) + (uint32_t)p01 + (uint32_t)p10;
    return p11 + (p01 >> 32) + (p10 >> 32) + (mid >> 32);
}
// table entry: x = rcp lo, y = rcp hi, z = bias | cmpl << 16, w = xmax_hi (freq << (31 - pb)) ; shift in a second array
template <int MULHI>
__global__ void k_enc_e2(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    uint32_t wp = 0;
    const uint32_t npairs = n / 2;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    uint32_t pend_w = 0, pend_bits = 0;     // previous step: word and the pair's (p0, p1)
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 8) {
        const uint4 v = in16[k >> 3];
        const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu;
            const uint4 e = tab[s]; const uint32_t sh = shf[s];
            // store of the previous step (its ballot is a full step old)
            {
                const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0;
                if (mine) wout[wp + (h ? p0 : 0u)] = pend_w;
                wp += p0 + p1;
            }
            const bool p = xhi >= e.w;
            pend_w = xlo;
            xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
            const uint32_t bal = __ballot_sync(0xffffffffu, p);
            pend_bits = (bal >> (lane & 30u)) & 3u;
            unsigned long long q;
            if (MULHI == 0) q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x);
            else if (MULHI == 2) q = mulhi64_tree(xlo, xhi, e.x, e.y);
            else q = mulhi64_flat(xlo, xhi, e.x, e.y);
            q >>= sh;
            const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
            const unsigned long long x = q * (e.z >> 16) + t;
            xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
        }
    }
    { const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0; if (mine) wout[wp + (h ? p0 : 0u)] = pend_w; wp += p0 + p1; }
    long long t1 = clock64();
    if (lane < 2) { wout[wp + 2 * lane] = xlo; wout[wp + 2 * lane + 1] = xhi; }
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = wp; }
}

template <int MULHI>
__global__ void k_enc_e3(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    uint32_t wp = 0;
    const uint32_t npairs = n / 2;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    uint32_t pend_w = 0, pend_bits = 0;     // previous step: word and the pair's (p0, p1)
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 8) {
        const uint4 v = in16[k >> 3];
        const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
        uint4 ev[8]; uint32_t shv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu; ev[j] = tab[s]; shv[j] = shf[s]; }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 e = ev[j]; const uint32_t sh = shv[j];
            // store of the previous step (its ballot is a full step old)
            {
                const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0;
                if (mine) wout[wp + (h ? p0 : 0u)] = pend_w;
                wp += p0 + p1;
            }
            const bool p = xhi >= e.w;
            pend_w = xlo;
            xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
            const uint32_t bal = __ballot_sync(0xffffffffu, p);
            pend_bits = (bal >> (lane & 30u)) & 3u;
            unsigned long long q;
            if (MULHI == 0) q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x);
            else if (MULHI == 2) q = mulhi64_tree(xlo, xhi, e.x, e.y);
            else q = mulhi64_flat(xlo, xhi, e.x, e.y);
            q >>= sh;
            const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
            const unsigned long long x = q * (e.z >> 16) + t;
            xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
        }
    }
    { const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0; if (mine) wout[wp + (h ? p0 : 0u)] = pend_w; wp += p0 + p1; }
    long long t1 = clock64();
    if (lane < 2) { wout[wp + 2 * lane] = xlo; wout[wp + 2 * lane + 1] = xhi; }
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = wp; }
}

template <int MULHI>
__global__ void k_enc_e4(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    uint32_t wp = 0;
    const uint32_t npairs = n / 2;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    uint32_t pend_w = 0, pend_bal = 0; const uint32_t bsh = lane & 30u;     // previous step: word and the pair's (p0, p1)
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 8) {
        const uint4 v = in16[k >> 3];
        const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
        uint4 ev[8]; uint32_t shv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu; ev[j] = tab[s]; shv[j] = shf[s]; }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 e = ev[j]; const uint32_t sh = shv[j];
            // store of the previous step (its ballot is a full step old)
            {
                const uint32_t pend_bits = pend_bal >> bsh;
                const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0;
                if (mine) wout[wp + (h ? p0 : 0u)] = pend_w;
                wp += p0 + p1;
            }
            const bool p = xhi >= e.w;
            pend_w = xlo;
            xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
            pend_bal = __ballot_sync(0xffffffffu, p);
            unsigned long long q;
            if (MULHI == 0) q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x);
            else if (MULHI == 2) q = mulhi64_tree(xlo, xhi, e.x, e.y);
            else q = mulhi64_flat(xlo, xhi, e.x, e.y);
            q >>= sh;
            const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
            const unsigned long long x = q * (e.z >> 16) + t;
            xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
        }
    }
    { const uint32_t pend_bits = pend_bal >> bsh; const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0; if (mine) wout[wp + (h ? p0 : 0u)] = pend_w; wp += p0 + p1; }
    long long t1 = clock64();
    if (lane < 2) { wout[wp + 2 * lane] = xlo; wout[wp + 2 * lane + 1] = xhi; }
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = wp; }
}


template <int MULHI>
__global__ void k_enc_e5(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    uint32_t wp = 0;
    asm volatile("" : "+l"(wout));
    const uint32_t npairs = n / 2;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    uint32_t pend_w = 0, pend_bal = 0; const uint32_t bsh = lane & 30u;     // previous step: word and the pair's (p0, p1)
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 8) {
        const uint4 v = in16[k >> 3];
        const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
        uint4 ev[8]; uint32_t shv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu; ev[j] = tab[s]; shv[j] = shf[s]; }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 e = ev[j]; const uint32_t sh = shv[j];
            // store of the previous step (its ballot is a full step old)
            {
                const uint32_t pend_bits = pend_bal >> bsh;
                const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0;
                if (mine) wout[wp + (h ? p0 : 0u)] = pend_w;
                wp += p0 + p1;
            }
            const bool p = xhi >= e.w;
            pend_w = xlo;
            xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
            pend_bal = __ballot_sync(0xffffffffu, p);
            unsigned long long q;
            if (MULHI == 0) q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x);
            else if (MULHI == 2) q = mulhi64_tree(xlo, xhi, e.x, e.y);
            else q = mulhi64_flat(xlo, xhi, e.x, e.y);
            q >>= sh;
            const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
            const unsigned long long x = q * (e.z >> 16) + t;
            xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
        }
    }
    { const uint32_t pend_bits = pend_bal >> bsh; const uint32_t p0 = pend_bits & 1u, p1 = (pend_bits >> 1) & 1u, mine = h ? p1 : p0; if (mine) wout[wp + (h ? p0 : 0u)] = pend_w; wp += p0 + p1; }
    long long t1 = clock64();
    if (lane < 2) { wout[wp + 2 * lane] = xlo; wout[wp + 2 * lane + 1] = xhi; }
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = wp; }
}



// E6: lanes fully independent: private word list + renormalisation flag bitmap per lane (merged afterwards)
__global__ void k_enc_e6(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, uint32_t* flags_out, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    const uint32_t npairs = n / 2;
    uint32_t* priv = wout + (size_t)lane * npairs;   // private list
    uint32_t* fl = flags_out + (size_t)lane * (npairs / 32 + 1);
    uint32_t cnt = 0;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 32) {
        uint32_t flags = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint4 v = in16[(k >> 3) + g];
            const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                    (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu;
                const uint4 e = tab[s]; const uint32_t sh = shf[s];
                const bool p = xhi >= e.w;
                priv[cnt] = xlo; cnt += p ? 1u : 0u;
                flags |= (p ? 1u : 0u) << (g * 8 + j);
                xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
                const unsigned long long q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x) >> sh;
                const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
                const unsigned long long x = q * (e.z >> 16) + t;
                xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
            }
        }
        fl[k >> 5] = flags;
    }
    long long t1 = clock64();
    priv[cnt] = xlo; priv[cnt + 1] = xhi;
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = cnt; }
}

template <int VAR>
__global__ void k_enc_e7(const uint8_t* sym, uint32_t n, const uint4* tab_g, const uint32_t* sh_g, uint32_t* wout, uint32_t* flags_out, long long* cyc) {
    __shared__ uint4 tab[16];
    __shared__ uint32_t shf[16];
    const uint32_t lane = threadIdx.x, h = lane & 1;
    if (lane < 16) { tab[lane] = tab_g[lane]; shf[lane] = sh_g[lane]; }
    __syncwarp();
    uint32_t xlo = 0x80000000u, xhi = 0;
    const uint32_t npairs = n / 2;
    uint32_t* priv = wout + (size_t)lane * npairs;   // private list
    uint32_t* fl = flags_out + (size_t)lane * (npairs / 32 + 1);
    uint32_t cnt = 0;
    const uint4* in16 = reinterpret_cast<const uint4*>(sym);
    long long t0 = clock64();
    for (uint32_t k = 0; k < npairs; k += 32) {
        uint32_t flags = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint4 v = in16[(k >> 3) + g];
            const uint32_t u[4] = { (h ? v.x >> 8 : v.x) & 0x00FF00FFu, (h ? v.y >> 8 : v.y) & 0x00FF00FFu, (h ? v.z >> 8 : v.z) & 0x00FF00FFu,
                                    (h ? v.w >> 8 : v.w) & 0x00FF00FFu };
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t s = (u[j >> 1] >> (16 * (j & 1))) & 0xFFu;
                const uint4 e = tab[s]; const uint32_t sh = shf[s];
                const bool p = xhi >= e.w;
                if (VAR == 7) { priv[k + g * 8 + j] = xlo; flags |= (p ? 1u : 0u) << (g * 8 + j); }
                else if (VAR == 5) { if (p) priv[cnt] = xlo; cnt += p ? 1u : 0u; flags |= (p ? 1u : 0u) << (g * 8 + j); }
                else if (VAR == 6) { if (p) priv[cnt] = xlo; cnt += p ? 1u : 0u; }
                else if (VAR != 1) { priv[cnt] = xlo; cnt += p ? 1u : 0u; flags |= (p ? 1u : 0u) << (g * 8 + j); }
                xlo = p ? xhi : xlo; xhi = p ? 0u : xhi;
                unsigned long long q;
                if (VAR == 2) q = (((unsigned long long)xhi << 32) | xlo) >> 12;          // no mulhi
                else if (VAR == 3) q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x) >> 5;   // const shift
                else q = __umul64hi(((unsigned long long)xhi << 32) | xlo, ((unsigned long long)e.y << 32) | e.x) >> sh;
                const unsigned long long t = (((unsigned long long)xhi << 32) | xlo) + (e.z & 0xFFFFu);
                const unsigned long long x = VAR == 4 ? (q + t) : (q * (e.z >> 16) + t);
                xlo = (uint32_t)x; xhi = (uint32_t)(x >> 32);
            }
        }
        fl[k >> 5] = flags;
    }
    long long t1 = clock64();
    priv[cnt] = xlo; priv[cnt + 1] = xhi;
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = cnt; }
}

__global__ void k_imadwide2(uint32_t* out, long long* cyc, uint32_t a) {
    unsigned long long x = threadIdx.x + 12345;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 8192; i++) x = (unsigned long long)(uint32_t)x * a + x;
    long long t1 = clock64();
    out[threadIdx.x] = (uint32_t)x + (uint32_t)(x >> 32); if (!threadIdx.x) cyc[0] = t1 - t0;
}
__global__ void k_imad32(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t x = threadIdx.x + 12345;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 8192; i++) x = x * a + 7;
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[0] = t1 - t0;
}
static uint32_t g_shift;
static uint4 make_encsym(uint32_t freq, uint32_t start, int pb) {
    uint64_t rcp; uint32_t shift, bias;
    if (freq < 2) { rcp = ~0ull; shift = 0; bias = start + (1u << pb) - 1; }
    else { uint32_t sh = 0; while (freq > (1u << sh)) sh++;
        const uint64_t x1 = 1ull << (sh + 31), t1 = x1 / freq, x0 = (uint64_t)(freq - 1) + ((x1 % freq) << 32), t0 = x0 / freq;
        rcp = t0 + (t1 << 32); shift = sh - 1; bias = start; }
    const uint32_t cmpl = (1u << pb) - freq;
    g_shift = shift; return make_uint4((uint32_t)rcp, (uint32_t)(rcp >> 32), bias | (cmpl << 16), freq << (31 - pb));
}

int main() {
    const int pb = 12; const uint32_t n = 1 << 17;
    // alphabet like the synthetic 4K frame: 9 symbols, skewed
    const double pr[9] = { 0.001, 0.01, 0.04, 0.22, 0.52, 0.18, 0.02, 0.005, 0.004 };
    uint32_t cum[10]; cum[0] = 0; { double a = 0; for (int i = 0; i < 9; i++) { a += pr[i]; cum[i + 1] = (uint32_t)(a * (1 << pb) + 0.5); } cum[9] = 1 << pb; }
    std::vector<uint32_t> lut(1 << pb);
    for (int s = 0; s < 9; s++) for (uint32_t i = cum[s]; i < cum[s + 1]; i++) lut[i] = (i - cum[s]) | (s << 14) | ((cum[s + 1] - cum[s]) << 18);
    std::vector<uint32_t> words(1 << 16); srand(1); for (auto& w : words) w = (uint32_t)rand() * 2654435761u + rand();
    uint32_t *d_lut, *d_words; uint8_t* d_out; long long* cyc;
    cudaMalloc(&d_lut, lut.size() * 4); cudaMalloc(&d_words, words.size() * 4); cudaMalloc(&d_out, n * 4 + 64); cudaMallocManaged(&cyc, 64);
    cudaMemcpy(d_lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_words, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
    const size_t sm = (1u << pb) * 4 + RING * 4;
    for (int r = 0; r < 2; r++) { k_dec_d1<32><<<1, 32, sm>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
    printf("dec_d1 32 lanes      %7.2f cycles/symbol\n", (double)cyc[0] / n);
    for (int r = 0; r < 2; r++) { k_dec_d1<1><<<1, 32, sm>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
    printf("dec_d1 1 lane        %7.2f cycles/symbol\n", (double)cyc[0] / n);
    for (int r = 0; r < 2; r++) { k_dec_d2<<<1, 32, sm + 64>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
    printf("dec_d2               %7.2f cycles/symbol\n", (double)cyc[0] / n);
    {   std::vector<uint32_t> lutT(1 << pb);
        for (int s2 = 0; s2 < 9; s2++) for (uint32_t i = cum[s2]; i < cum[s2 + 1]; i++) lutT[i] = (uint32_t)((0x7FFFFFFFull - (i - cum[s2])) / (cum[s2 + 1] - cum[s2]));
        uint32_t* d_lutT; cudaMalloc(&d_lutT, lutT.size() * 4); cudaMemcpy(d_lutT, lutT.data(), lutT.size() * 4, cudaMemcpyHostToDevice);
        std::vector<uint8_t> o2(n), o3(n);
        k_dec_d2<<<1, 32, sm + 64>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); cudaMemcpy(o2.data(), d_out, n, cudaMemcpyDeviceToHost);
        for (int r = 0; r < 2; r++) { k_dec_d2one<<<1, 32, sm + 64>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
        printf("dec_d2 one state     %7.2f cycles/step\n", (double)cyc[0] / (n / 2));
        cudaFuncSetAttribute(k_dec_d3, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * sm + 64);
        for (int r = 0; r < 2; r++) { k_dec_d3<<<1, 32, 2 * sm + 64>>>(d_lut, d_lutT, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(o3.data(), d_out, n, cudaMemcpyDeviceToHost);
        { std::vector<uint8_t> o4(n);
          for (int r = 0; r < 2; r++) { k_dec_d4<<<1, 32, sm + 64>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
          cudaMemcpy(o4.data(), d_out, n, cudaMemcpyDeviceToHost);
          printf("dec_d4 (skewed)      %7.2f cycles/symbol %s\n", (double)cyc[0] / n, o4 == o2 ? "same symbols as d2" : "MISMATCH vs d2"); }
        { std::vector<uint8_t> o5(n);
          for (int r = 0; r < 2; r++) { k_dec_d5<0><<<1, 32, sm + 64>>>(d_lut, d_words, 4096, n, pb, d_out, cyc); cudaDeviceSynchronize(); }
          cudaMemcpy(o5.data(), d_out, n, cudaMemcpyDeviceToHost);
          printf("dec_d5 (pair SIMT)   %7.2f cycles/symbol %s\n", (double)cyc[0] / n, o5 == o2 ? "same symbols as d2" : "MISMATCH vs d2"); }
        printf("dec_d3 (pred table)  %7.2f cycles/symbol %s\n", (double)cyc[0] / n, o2 == o3 ? "same symbols as d2" : "MISMATCH vs d2"); }
    for (int r = 0; r < 2; r++) { k_issue<<<1, 32>>>((uint32_t*)d_out, cyc, 5); cudaDeviceSynchronize(); }
    printf("issue 8 indep (xor,add)   %7.2f cycles per 16 ops (SASS-count dependent)\n", (double)cyc[0] / 4096);
    for (int r = 0; r < 2; r++) { k_issue_mix<<<1, 32>>>((uint32_t*)d_out, cyc, 5); cudaDeviceSynchronize(); }
    printf("issue mix 4 imad + 4(shf,lop) %7.2f cycles per 12 ops\n", (double)cyc[0] / 4096);
    // ---- walk: Markov streams from a random nl sequence with the same marginals
    const uint32_t m = 1 << 17;
    std::vector<uint8_t> seq(m); { for (auto& s : seq) { double u = rand() / (RAND_MAX + 1.0), a = 0; int k = 0; for (; k < 8; k++) { a += pr[k]; if (u < a) break; } s = (uint8_t)k; } }
    std::vector<std::vector<uint8_t>> st(9); { uint8_t pl = 0; for (auto s : seq) { st[pl].push_back(s); pl = s; } }
    std::vector<uint8_t> flat; uint32_t soff[16] = { 0 }, slen[16] = { 0 };
    for (int c = 0; c < 9; c++) { soff[c] = (uint32_t)flat.size(); slen[c] = (uint32_t)st[c].size(); flat.insert(flat.end(), st[c].begin(), st[c].end()); while (flat.size() % 16) flat.push_back(0); }
    flat.resize(flat.size() + 64);
    uint8_t* d_st; uint32_t *d_soff, *d_slen; cudaMalloc(&d_st, flat.size()); cudaMalloc(&d_soff, 64); cudaMalloc(&d_slen, 64);
    cudaMemcpy(d_st, flat.data(), flat.size(), cudaMemcpyHostToDevice); cudaMemcpy(d_soff, soff, 64, cudaMemcpyHostToDevice); cudaMemcpy(d_slen, slen, 64, cudaMemcpyHostToDevice);
    std::vector<uint8_t> back(m);
    for (int r = 0; r < 2; r++) { k_walk_shfl<<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    printf("walk_shfl            %7.2f cycles/step  %s\n", (double)cyc[0] / m, back == seq ? "ok" : "MISMATCH");
    for (int r = 0; r < 2; r++) { k_walk_var<0><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_var0 (shfl+keep)        %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_var<1><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_var1 (+pop)             %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_var<2><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_var2 (+vote refill)     %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_var<3><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_var3 (vote every 4th)   %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_nib<0><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    printf("walk_nib (production)        %7.2f cycles/step %s\n", (double)cyc[0] / m, back == seq ? "ok" : "MISMATCH");
    for (int r = 0; r < 2; r++) { k_walk_nib<1><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_nib (no loads)          %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_nib2<0><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    printf("walk_nib2 (nc + prefetch L1) %7.2f cycles/step %s\n", (double)cyc[0] / m, back == seq ? "ok" : "MISMATCH");
    for (int r = 0; r < 2; r++) { k_walk_nib2<1><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_nib2 (ca + prefetch L1) %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_nib2<2><<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    printf("walk_nib2 (ca, no prefetch)  %7.2f cycles/step\n", (double)cyc[0] / m);
    for (int r = 0; r < 2; r++) { k_walk_nib3<<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    printf("walk_nib3 (spread refill)    %7.2f cycles/step %s\n", (double)cyc[0] / m, back == seq ? "ok" : "MISMATCH");
    cudaFuncSetAttribute(k_walk_nib4, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int r = 0; r < 2; r++) { k_walk_nib4<<<1, 32, 100 * 1024>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    printf("walk_nib4 (smem streams)     %7.2f cycles/step %s\n", (double)cyc[0] / m, back == seq ? "ok" : "MISMATCH");
    for (int r = 0; r < 2; r++) { k_walk_smem<<<1, 32>>>(d_st, d_soff, d_slen, m, d_out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(back.data(), d_out, m, cudaMemcpyDeviceToHost);
    { bool ok = true; for (uint32_t i = 0; i < m; i++) ok &= back[i] == seq[i]; printf("walk_smem            %7.2f cycles/step  %s\n", (double)cyc[0] / m, ok ? "ok" : "MISMATCH"); }
    // ---- encode
    std::vector<uint4> tab(16), tab2(16); std::vector<uint32_t> shv(16);
    for (int s = 0; s < 9; s++) { tab2[s] = make_encsym(cum[s + 1] - cum[s], cum[s], pb); shv[s] = g_shift; tab[s] = tab2[s]; tab[s].w = ((cum[s + 1] - cum[s]) << (31 - pb)) >> 8 | (g_shift << 24); }
    uint4* d_tab; cudaMalloc(&d_tab, 256); cudaMemcpy(d_tab, tab.data(), 256, cudaMemcpyHostToDevice);
    uint8_t* d_sym; cudaMalloc(&d_sym, m); cudaMemcpy(d_sym, seq.data(), m, cudaMemcpyHostToDevice);
    uint32_t* d_w; cudaMalloc(&d_w, m * 4);
    for (int r = 0; r < 2; r++) { k_enc_e1<<<1, 32>>>(d_sym, m, d_tab, pb, d_w, cyc, 9); cudaDeviceSynchronize(); }
    printf("enc_e1               %7.2f cycles/symbol (%lld words)\n", (double)cyc[0] / m, cyc[1]);
    {   // host reference (libxpng.c:362-392) for the E2 kernels
        std::vector<uint32_t> ref; unsigned long long x[2] = { 1ull << 31, 1ull << 31 };
        for (uint32_t i = 0; i < (m & ~1u); i++) {
            const int st_ = i & 1; const uint32_t sy = seq[i], f = cum[sy + 1] - cum[sy];
            if (x[st_] >= ((unsigned long long)f << (63 - pb))) { ref.push_back((uint32_t)x[st_]); x[st_] >>= 32; }
            const uint4 e = tab2[sy]; const unsigned long long rcp = ((unsigned long long)e.y << 32) | e.x;
            const unsigned long long q = (unsigned long long)(((unsigned __int128)x[st_] * rcp) >> 64) >> shv[sy];
            x[st_] += (e.z & 0xFFFF) + q * (e.z >> 16);
        }
        ref.push_back((uint32_t)x[0]); ref.push_back((uint32_t)(x[0] >> 32)); ref.push_back((uint32_t)x[1]); ref.push_back((uint32_t)(x[1] >> 32));
        uint4* d_tab2; uint32_t* d_sh; cudaMalloc(&d_tab2, 256); cudaMalloc(&d_sh, 64);
        cudaMemcpy(d_tab2, tab2.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(d_sh, shv.data(), 64, cudaMemcpyHostToDevice);
        std::vector<uint32_t> got(ref.size());
        for (int r = 0; r < 2; r++) { k_enc_e2<0><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e2 umul64hi      %7.2f cycles/symbol (%lld words, ref %zu) %s\n", (double)cyc[0] / m, cyc[1], ref.size() - 4, got == ref ? "ok" : "MISMATCH");
        for (int r = 0; r < 2; r++) { k_enc_e3<0><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e3 prefetch tab  %7.2f cycles/symbol (%lld words) %s\n", (double)cyc[0] / m, cyc[1], got == ref ? "ok" : "MISMATCH");
        for (int r = 0; r < 2; r++) { k_enc_e4<0><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e4 deferred vote %7.2f cycles/symbol (%lld words) %s\n", (double)cyc[0] / m, cyc[1], got == ref ? "ok" : "MISMATCH");
        for (int r = 0; r < 2; r++) { k_enc_e4<2><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e4 tree mulhi    %7.2f cycles/symbol (%lld words) %s\n", (double)cyc[0] / m, cyc[1], got == ref ? "ok" : "MISMATCH");
        for (int r = 0; r < 2; r++) { k_enc_e5<0><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e5 pinned ptr    %7.2f cycles/symbol (%lld words) %s\n", (double)cyc[0] / m, cyc[1], got == ref ? "ok" : "MISMATCH");
        for (int r = 0; r < 2; r++) { k_enc_e5<2><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        printf("enc_e5 pinned + tree %7.2f cycles/symbol\n", (double)cyc[0] / m);
        { uint32_t *d_w2, *d_fl; cudaMalloc(&d_w2, (size_t)32 * (m / 2 + 8) * 4); cudaMalloc(&d_fl, 32 * (m / 64 + 8) * 4);
          for (int r = 0; r < 2; r++) { k_enc_e6<<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc); cudaDeviceSynchronize(); }
          const char* nm[8] = { "full", "no stores", "no mulhi", "const shift", "no q*cmpl", "pred store", "pred st noflag", "store every step" };
          for (int var = 0; var < 8; var++) { for (int r = 0; r < 2; r++) {
              if (var == 7) k_enc_e7<7><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc);
              if (var == 5) k_enc_e7<5><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc); if (var == 6) k_enc_e7<6><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc);
              if (var == 0) k_enc_e7<0><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc); if (var == 1) k_enc_e7<1><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc);
              if (var == 2) k_enc_e7<2><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc); if (var == 3) k_enc_e7<3><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc);
              if (var == 4) k_enc_e7<4><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w2, d_fl, cyc); cudaDeviceSynchronize(); }
            printf("enc_e7 %-12s %7.2f cycles/symbol\n", nm[var], (double)cyc[0] / m); }
          printf("enc_e6 private lists %7.2f cycles/symbol (%lld words lane0)\n", (double)cyc[0] / m, cyc[1]); }
        for (int r = 0; r < 2; r++) { k_imadwide2<<<1, 32>>>(d_w, cyc, 77777); cudaDeviceSynchronize(); }
        printf("imad.wide chain      %7.2f cycles/op\n", (double)cyc[0] / 8192);
        for (int r = 0; r < 2; r++) { k_imad32<<<1, 32>>>(d_w, cyc, 77777); cudaDeviceSynchronize(); }
        printf("imad32 chain         %7.2f cycles/op\n", (double)cyc[0] / 8192);
        for (int r = 0; r < 2; r++) { k_enc_e2<1><<<1, 32>>>(d_sym, m, d_tab2, d_sh, d_w, cyc); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_w, ref.size() * 4, cudaMemcpyDeviceToHost);
        printf("enc_e2 flat mulhi    %7.2f cycles/symbol (%lld words) %s\n", (double)cyc[0] / m, cyc[1], got == ref ? "ok" : "MISMATCH");
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
