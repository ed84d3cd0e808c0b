// Experimental low-register-pressure GF(2^233) multiplier (see scripts/mulbench.cu).
#pragma once
#include "../dv-pari_b200/csrc/gf233.cuh"

namespace dvp {

#ifdef __CUDACC__
#ifndef DVP_MW_VOLATILE
#define DVP_MW_VOLATILE volatile
#endif
__device__ __forceinline__ uint64_t mw(uint32_t a, uint32_t b) {
    uint64_t r;
    asm DVP_MW_VOLATILE("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint64_t mwa(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r;
    asm DVP_MW_VOLATILE("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}

// 32x32 -> 64 carry-less product, products issued in a fixed order
__device__ __forceinline__ uint64_t clmul32_v2(uint32_t a, uint32_t b) {
    const uint32_t a0 = a & 0x11111111u, a1 = a & 0x22222222u, a2 = a & 0x44444444u, a3 = a & 0x88888888u;
    const uint32_t b0 = b & 0x11111111u, b1 = b & 0x22222222u, b2 = b & 0x44444444u, b3 = b & 0x88888888u;
    uint64_t z0 = mwa(a1, b3, mw(a0, b0)) ^ mw(a2, b2) ^ mw(a3, b1);
    uint64_t z1 = mwa(a2, b3, mw(a0, b1)) ^ mwa(a3, b2, mw(a1, b0));
    uint64_t x = (z0 & 0x5555555555555555ull) | (z1 & 0xaaaaaaaaaaaaaaaaull);
    uint64_t z2 = mwa(a3, b3, mw(a0, b2)) ^ mw(a1, b1) ^ mw(a2, b0);
    uint64_t z3 = mw(a0, b3) ^ mw(a1, b2) ^ mw(a2, b1) ^ mw(a3, b0);
    uint64_t y = (z2 & 0x5555555555555555ull) | (z3 & 0xaaaaaaaaaaaaaaaaull);
    return (x & 0x3333333333333333ull) | (y & 0xccccccccccccccccull);
}

// c[0..3] ^= (a0 + a1 X)(b0 + b1 X), X = x^32 (Karatsuba, 3 word products)
__device__ __forceinline__ void mul2w_acc(uint32_t *c, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const uint64_t lo = clmul32_v2(a0, b0);
    const uint64_t hi = clmul32_v2(a1, b1);
    const uint64_t mid = clmul32_v2(a0 ^ a1, b0 ^ b1) ^ lo ^ hi;
    c[0] ^= (uint32_t)lo;
    c[1] ^= (uint32_t)(lo >> 32) ^ (uint32_t)mid;
    c[2] ^= (uint32_t)hi ^ (uint32_t)(mid >> 32);
    c[3] ^= (uint32_t)(hi >> 32);
}

// c[0..7] ^= A(4 words) * B(4 words):  lo at 0, hi at 4, (lo + hi + mid) at 2
__device__ __forceinline__ void mul4w_acc(uint32_t *c, const uint32_t *a, const uint32_t *b) {
    uint32_t t[4] = {0, 0, 0, 0};
    mul2w_acc(t, a[0], a[1], b[0], b[1]); // lo
    c[0] ^= t[0]; c[1] ^= t[1]; c[2] ^= t[2] ^ t[0]; c[3] ^= t[3] ^ t[1]; c[4] ^= t[2]; c[5] ^= t[3];
    uint32_t u[4] = {0, 0, 0, 0};
    mul2w_acc(u, a[2], a[3], b[2], b[3]); // hi
    c[2] ^= u[0]; c[3] ^= u[1]; c[4] ^= u[2] ^ u[0]; c[5] ^= u[3] ^ u[1]; c[6] ^= u[2]; c[7] ^= u[3];
    mul2w_acc(c + 2, a[0] ^ a[2], a[1] ^ a[3], b[0] ^ b[2], b[1] ^ b[3]);
}

__device__ __forceinline__ gf gf_mul_v2(const gf &a, const gf &b) {
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = 0;
    mul4w_acc(t, a.v, b.v); // lo
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c[i] = t[i];
        c[i + 4] = t[i + 4] ^ t[i];
        c[i + 8] = t[i + 4];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = 0;
    mul4w_acc(t, a.v + 4, b.v + 4); // hi
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c[i + 4] ^= t[i];
        c[i + 8] ^= t[i + 4] ^ t[i];
        c[i + 12] = t[i + 4];
    }
    uint32_t sa[4], sb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sa[i] = a.v[i] ^ a.v[i + 4];
        sb[i] = b.v[i] ^ b.v[i + 4];
    }
    mul4w_acc(c + 4, sa, sb);
    return gf_reduce(c);
}
#endif

} // namespace dvp
