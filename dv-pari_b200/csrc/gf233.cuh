// GF(2^233) = GF(2)[x]/(x^233 + x^74 + 1) for sm_100a: the base field of sect233k1.
//
// Replaces the field layer of crate xs233-sys (reached from /root/reference/src/curve.rs:13-137).
// There is no carry-less multiply instruction on the SM, so a 32x32 carry-less product is built
// from 16 IMAD.WIDE on operands masked to every 4th bit (column sums <= 8 never carry into the
// next 4-bit field) recombined with LOP3; 8x8 words use three levels of Karatsuba (27 word
// products).  Squaring is a mask/IMAD bit-spread ladder.  Elements are 8 x u32, little-endian,
// always fully reduced (bits >= 233 zero).
#pragma once
#include <stdint.h>

namespace dvp {

struct gf {
    uint32_t v[8];
};

__host__ __device__ __forceinline__ gf gf_zero() {
    gf r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
__host__ __device__ __forceinline__ gf gf_one() {
    gf r = gf_zero();
    r.v[0] = 1;
    return r;
}
__host__ __device__ __forceinline__ gf gf_add(const gf &a, const gf &b) {
    gf r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = a.v[i] ^ b.v[i];
    return r;
}
__host__ __device__ __forceinline__ bool gf_is_zero(const gf &a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.v[i];
    return t == 0;
}
__host__ __device__ __forceinline__ bool gf_eq(const gf &a, const gf &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
}
// r = c ? a : b (branch-free)
__host__ __device__ __forceinline__ gf gf_sel(bool c, const gf &a, const gf &b) {
    gf r;
    uint32_t m = c ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    return r;
}

// 32x32 -> 64 carry-less product ("holes" trick: every 4th bit, 16 integer products).
__host__ __device__ __forceinline__ uint64_t clmul32(uint32_t a, uint32_t b) {
    const uint32_t a0 = a & 0x11111111u, a1 = a & 0x22222222u, a2 = a & 0x44444444u, a3 = a & 0x88888888u;
    const uint32_t b0 = b & 0x11111111u, b1 = b & 0x22222222u, b2 = b & 0x44444444u, b3 = b & 0x88888888u;
    // pairs whose bit offsets differ by 4 never reach a column sum of 16, so they may share an
    // integer accumulate; the remaining terms are xor-ed (garbage bits are masked off below).
    uint64_t z0 = ((uint64_t)a0 * b0 + (uint64_t)a1 * b3) ^ ((uint64_t)a2 * b2) ^ ((uint64_t)a3 * b1);
    uint64_t z1 = ((uint64_t)a0 * b1 + (uint64_t)a2 * b3) ^ ((uint64_t)a1 * b0 + (uint64_t)a3 * b2);
    uint64_t z2 = ((uint64_t)a0 * b2 + (uint64_t)a3 * b3) ^ ((uint64_t)a1 * b1) ^ ((uint64_t)a2 * b0);
    uint64_t z3 = ((uint64_t)a0 * b3) ^ ((uint64_t)a1 * b2) ^ ((uint64_t)a2 * b1) ^ ((uint64_t)a3 * b0);
    // bit-select merge: class 0/2 live on even bits, 1/3 on odd bits
    uint64_t x = (z0 & 0x5555555555555555ull) | (z1 & 0xaaaaaaaaaaaaaaaaull);
    uint64_t y = (z2 & 0x5555555555555555ull) | (z3 & 0xaaaaaaaaaaaaaaaaull);
    return (x & 0x3333333333333333ull) | (y & 0xccccccccccccccccull);
}

// 2x2 words (Karatsuba, 3 products): r[0..3] = a[0..1] * b[0..1]
__host__ __device__ __forceinline__ void clmul_2w(uint32_t r[4], const uint32_t a[2], const uint32_t b[2]) {
    uint64_t lo = clmul32(a[0], b[0]);
    uint64_t hi = clmul32(a[1], b[1]);
    uint64_t mid = clmul32(a[0] ^ a[1], b[0] ^ b[1]) ^ lo ^ hi;
    r[0] = (uint32_t)lo;
    r[1] = (uint32_t)(lo >> 32) ^ (uint32_t)mid;
    r[2] = (uint32_t)hi ^ (uint32_t)(mid >> 32);
    r[3] = (uint32_t)(hi >> 32);
}
// 4x4 words: r[0..7]
__host__ __device__ __forceinline__ void clmul_4w(uint32_t r[8], const uint32_t a[4], const uint32_t b[4]) {
    uint32_t lo[4], hi[4], mid[4], sa[2], sb[2];
    clmul_2w(lo, a, b);
    clmul_2w(hi, a + 2, b + 2);
    sa[0] = a[0] ^ a[2]; sa[1] = a[1] ^ a[3];
    sb[0] = b[0] ^ b[2]; sb[1] = b[1] ^ b[3];
    clmul_2w(mid, sa, sb);
#pragma unroll
    for (int i = 0; i < 4; i++) mid[i] ^= lo[i] ^ hi[i];
    r[0] = lo[0]; r[1] = lo[1];
    r[2] = lo[2] ^ mid[0]; r[3] = lo[3] ^ mid[1];
    r[4] = hi[0] ^ mid[2]; r[5] = hi[1] ^ mid[3];
    r[6] = hi[2]; r[7] = hi[3];
}

// c[0..15] (466 significant bits) -> reduced element.  x^233 = x^74 + 1.
__host__ __device__ __forceinline__ gf gf_reduce(uint32_t c[16]) {
    // word i (i >= 8) holds x^(32 i + k): fold to bit offsets 32 i - 233 = 32 (i-8) + 23 and
    // 32 i - 159 = 32 (i-5) + 1
#pragma unroll
    for (int i = 15; i >= 8; i--) {
        const uint32_t t = c[i];
        c[i - 8] ^= t << 23;
        c[i - 7] ^= t >> 9;
        c[i - 5] ^= t << 1;
        c[i - 4] ^= t >> 31;
    }
    const uint32_t t = c[7] >> 9; // bits 233..255
    c[0] ^= t;
    c[2] ^= t << 10; // x^74 = word 2, bit 10
    c[3] ^= t >> 22;
    c[7] &= 0x1ffu;
    gf r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}

__host__ __device__ __forceinline__ gf gf_mul_portable(const gf &a, const gf &b) {
    uint32_t lo[8], hi[8], mid[8], sa[4], sb[4], c[16];
    clmul_4w(lo, a.v, b.v);
    clmul_4w(hi, a.v + 4, b.v + 4);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sa[i] = a.v[i] ^ a.v[i + 4];
        sb[i] = b.v[i] ^ b.v[i + 4];
    }
    clmul_4w(mid, sa, sb);
#pragma unroll
    for (int i = 0; i < 8; i++) mid[i] ^= lo[i] ^ hi[i];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c[i] = lo[i];
        c[i + 4] = lo[i + 4] ^ mid[i];
        c[i + 8] = hi[i] ^ mid[i + 4];
        c[i + 12] = hi[i + 4];
    }
    return gf_reduce(c);
}

#ifdef __CUDACC__
// Device form of the same product: the 16 class products of a word pair are issued as explicit mul.wide / mad.wide
// in a fixed order and the Karatsuba levels accumulate into their output words (scripts/mulbench.cu: 752 instead of
// 768 LOP3 per multiplication and 2-4 % more multiplications per second in both the inlined and the out-of-line form).
__device__ __forceinline__ uint64_t gf_mw(uint32_t a, uint32_t b) {
    uint64_t r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint64_t gf_mwa(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r;
    asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t clmul32_dev(uint32_t a, uint32_t b) {
    const uint32_t a0 = a & 0x11111111u, a1 = a & 0x22222222u, a2 = a & 0x44444444u, a3 = a & 0x88888888u;
    const uint32_t b0 = b & 0x11111111u, b1 = b & 0x22222222u, b2 = b & 0x44444444u, b3 = b & 0x88888888u;
    const uint64_t z0 = gf_mwa(a1, b3, gf_mw(a0, b0)) ^ gf_mw(a2, b2) ^ gf_mw(a3, b1);
    const uint64_t z1 = gf_mwa(a2, b3, gf_mw(a0, b1)) ^ gf_mwa(a3, b2, gf_mw(a1, b0));
    const uint64_t x = (z0 & 0x5555555555555555ull) | (z1 & 0xaaaaaaaaaaaaaaaaull);
    const uint64_t z2 = gf_mwa(a3, b3, gf_mw(a0, b2)) ^ gf_mw(a1, b1) ^ gf_mw(a2, b0);
    const uint64_t z3 = gf_mw(a0, b3) ^ gf_mw(a1, b2) ^ gf_mw(a2, b1) ^ gf_mw(a3, b0);
    const uint64_t y = (z2 & 0x5555555555555555ull) | (z3 & 0xaaaaaaaaaaaaaaaaull);
    return (x & 0x3333333333333333ull) | (y & 0xccccccccccccccccull);
}
// c[0..3] ^= (a0 + a1 X)(b0 + b1 X), X = x^32
__device__ __forceinline__ void clmul_2w_acc(uint32_t *c, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const uint64_t lo = clmul32_dev(a0, b0);
    const uint64_t hi = clmul32_dev(a1, b1);
    const uint64_t mid = clmul32_dev(a0 ^ a1, b0 ^ b1) ^ lo ^ hi;
    c[0] ^= (uint32_t)lo;
    c[1] ^= (uint32_t)(lo >> 32) ^ (uint32_t)mid;
    c[2] ^= (uint32_t)hi ^ (uint32_t)(mid >> 32);
    c[3] ^= (uint32_t)(hi >> 32);
}
// c[0..7] ^= A(4 words) * B(4 words)
__device__ __forceinline__ void clmul_4w_acc(uint32_t *c, const uint32_t *a, const uint32_t *b) {
    uint32_t t[4] = {0, 0, 0, 0};
    clmul_2w_acc(t, a[0], a[1], b[0], b[1]);
    c[0] ^= t[0]; c[1] ^= t[1]; c[2] ^= t[2] ^ t[0]; c[3] ^= t[3] ^ t[1]; c[4] ^= t[2]; c[5] ^= t[3];
    uint32_t u[4] = {0, 0, 0, 0};
    clmul_2w_acc(u, a[2], a[3], b[2], b[3]);
    c[2] ^= u[0]; c[3] ^= u[1]; c[4] ^= u[2] ^ u[0]; c[5] ^= u[3] ^ u[1]; c[6] ^= u[2]; c[7] ^= u[3];
    clmul_2w_acc(c + 2, a[0] ^ a[2], a[1] ^ a[3], b[0] ^ b[2], b[1] ^ b[3]);
}
__device__ __forceinline__ gf gf_mul_dev(const gf &a, const gf &b) {
    uint32_t c[16], t[8];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = 0;
    clmul_4w_acc(t, a.v, b.v);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c[i] = t[i];
        c[i + 4] = t[i + 4] ^ t[i];
        c[i + 8] = t[i + 4];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = 0;
    clmul_4w_acc(t, a.v + 4, b.v + 4);
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
    clmul_4w_acc(c + 4, sa, sb);
    return gf_reduce(c);
}
#endif
} // namespace dvp
#include "gf233_mul2.cuh" // gf_mul_dev2: the same product from 32-bit IMAD only (two streams), the device path of gf_mul
namespace dvp {

__host__ __device__ __forceinline__ gf gf_mul(const gf &a, const gf &b) {
#if defined(__CUDA_ARCH__) && defined(DVP_GF_MUL_WIDE)
    return gf_mul_dev(a, b); // the IMAD.WIDE form, kept for A/B runs (make NVFLAGS+=-DDVP_GF_MUL_WIDE)
#elif defined(__CUDA_ARCH__)
    return gf_mul_dev2(a, b);
#else
    return gf_mul_portable(a, b);
#endif
}

// 16 bits -> 32 bits with zeros interleaved; each step is one mask and one multiply-add
__host__ __device__ __forceinline__ uint32_t spread16(uint32_t x) {
    uint32_t t;
    t = x & 0x0000ff00u; x = t * 255u + x;   // x = lo | hi << 8
    t = x & 0x00f000f0u; x = t * 15u + x;
    t = x & 0x0c0c0c0cu; x = t * 3u + x;
    t = x & 0x22222222u; x = t + x;
    return x;
}

__host__ __device__ __forceinline__ gf gf_sqr(const gf &a) {
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[2 * i] = spread16(a.v[i] & 0xffffu);
        c[2 * i + 1] = spread16(a.v[i] >> 16);
    }
    return gf_reduce(c);
}

__host__ __device__ inline gf gf_sqr_n(gf a, int n) {
    for (int i = 0; i < n; i++) a = gf_sqr(a);
    return a;
}

// Itoh-Tsujii, chain 1,2,3,6,7,14,28,29,58,116,232.  0 -> 0.
__host__ __device__ inline gf gf_inv(const gf &a) {
    gf b1 = a;
    gf b2 = gf_mul(gf_sqr(b1), b1);
    gf b3 = gf_mul(gf_sqr(b2), b1);
    gf b6 = gf_mul(gf_sqr_n(b3, 3), b3);
    gf b7 = gf_mul(gf_sqr(b6), b1);
    gf b14 = gf_mul(gf_sqr_n(b7, 7), b7);
    gf b28 = gf_mul(gf_sqr_n(b14, 14), b14);
    gf b29 = gf_mul(gf_sqr(b28), b1);
    gf b58 = gf_mul(gf_sqr_n(b29, 29), b29);
    gf b116 = gf_mul(gf_sqr_n(b58, 58), b58);
    gf b232 = gf_mul(gf_sqr_n(b116, 116), b116);
    return gf_sqr(b232);
}

} // namespace dvp
