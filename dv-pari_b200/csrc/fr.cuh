// Fr: the 232-bit prime scalar field of sect233k1 on sm_100a, 8 x u32 limbs, Montgomery R = 2^256.
//
// Bit-compatible with the reference's in-memory `Fr = Fp256<MontBackend<FqConfig,4>>`
// (/root/reference/src/curve.rs:16-22): 4 x u64 little-endian limbs == 8 x u32 little-endian limbs,
// value * 2^256 mod p, always fully reduced.  p = 2^231 + delta with delta < 2^115, so limbs 4..6
// of p are zero and the Montgomery reduction step only multiplies by 4 limbs plus a shift.
#pragma once
#include <stdint.h>

namespace dvp {

struct fr {
    uint32_t v[8];
};

// p = 0x80000000_00000000_00000000_00069d5b_b915bcd4_6efb1ad5_f173abdf
#define DVP_FR_P0 0xf173abdfu
#define DVP_FR_P1 0x6efb1ad5u
#define DVP_FR_P2 0xb915bcd4u
#define DVP_FR_P3 0x00069d5bu
#define DVP_FR_P7 0x00000080u
#define DVP_FR_NP0 0x8c382fe1u /* -p^-1 mod 2^32 */

__host__ __device__ __forceinline__ fr fr_zero() {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
// R mod p
__host__ __device__ __forceinline__ fr fr_one() {
    fr r;
    r.v[0] = 0x3373abdfu; r.v[1] = 0xc318337eu; r.v[2] = 0x1037c69eu; r.v[3] = 0x489471e2u;
    r.v[4] = 0xfffff2c5u; r.v[5] = 0xffffffffu; r.v[6] = 0xffffffffu; r.v[7] = 0x0000007fu;
    return r;
}
// R^2 mod p
__host__ __device__ __forceinline__ fr fr_r2() {
    fr r;
    r.v[0] = 0x09468bb6u; r.v[1] = 0x1710ac10u; r.v[2] = 0xdb9a5b86u; r.v[3] = 0xf7e3eb91u;
    r.v[4] = 0xb5b58a0au; r.v[5] = 0x93c813eeu; r.v[6] = 0xbebed802u; r.v[7] = 0x00000059u;
    return r;
}
__host__ __device__ __forceinline__ bool fr_is_zero(const fr &a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.v[i];
    return t == 0;
}
__host__ __device__ __forceinline__ bool fr_eq(const fr &a, const fr &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
}

__host__ __device__ __forceinline__ uint32_t fr_p_limb(int i) {
    return i == 0 ? DVP_FR_P0 : i == 1 ? DVP_FR_P1 : i == 2 ? DVP_FR_P2 : i == 3 ? DVP_FR_P3 : i == 7 ? DVP_FR_P7 : 0u;
}

// t >= p ?
__host__ __device__ __forceinline__ bool fr_geq_p(const uint32_t t[8]) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        const uint32_t pi = fr_p_limb(i);
        if (t[i] > pi) return true;
        if (t[i] < pi) return false;
    }
    return true;
}
__host__ __device__ __forceinline__ void fr_sub_p(uint32_t t[8]) {
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)t[i] - fr_p_limb(i) - br;
        t[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
}

__host__ __device__ __forceinline__ fr fr_add(const fr &a, const fr &b) {
    fr r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    if (fr_geq_p(r.v)) fr_sub_p(r.v);
    return r;
}
__host__ __device__ __forceinline__ fr fr_sub(const fr &a, const fr &b) {
    fr r;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (br) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)r.v[i] + fr_p_limb(i);
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return r;
}
__host__ __device__ __forceinline__ fr fr_neg(const fr &a) { return fr_sub(fr_zero(), a); }

// One Montgomery reduction step on t[0..9]: t = (t + m p) / 2^32 with m = t0 * np0.
// p has only limbs 0..3 and 7 non-zero.
__host__ __device__ __forceinline__ void fr_redc_step(uint32_t t[10]) {
    const uint32_t m = t[0] * DVP_FR_NP0;
    uint64_t c = (uint64_t)m * DVP_FR_P0 + t[0];
    c >>= 32;
    c += (uint64_t)m * DVP_FR_P1 + t[1]; t[0] = (uint32_t)c; c >>= 32;
    c += (uint64_t)m * DVP_FR_P2 + t[2]; t[1] = (uint32_t)c; c >>= 32;
    c += (uint64_t)m * DVP_FR_P3 + t[3]; t[2] = (uint32_t)c; c >>= 32;
    c += t[4]; t[3] = (uint32_t)c; c >>= 32;
    c += t[5]; t[4] = (uint32_t)c; c >>= 32;
    c += t[6]; t[5] = (uint32_t)c; c >>= 32;
    c += (uint64_t)m * DVP_FR_P7 + t[7]; t[6] = (uint32_t)c; c >>= 32;
    c += t[8]; t[7] = (uint32_t)c; c >>= 32;
    t[8] = t[9] + (uint32_t)c;
    t[9] = 0;
}

// CIOS Montgomery product, fully reduced
__host__ __device__ __forceinline__ fr fr_mul(const fr &a, const fr &b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[j] * b.v[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        fr_redc_step(t);
    }
    if (t[8] || fr_geq_p(t)) fr_sub_p(t);
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return r;
}
__host__ __device__ __forceinline__ fr fr_sqr(const fr &a) { return fr_mul(a, a); }

// Montgomery form -> canonical integer (< p), i.e. a * R^-1 mod p
__host__ __device__ __forceinline__ void fr_to_canonical(uint32_t out[8], const fr &a) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = a.v[i];
    t[8] = 0; t[9] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) fr_redc_step(t);
    if (t[8] || fr_geq_p(t)) fr_sub_p(t);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = t[i];
}
__host__ __device__ __forceinline__ fr fr_from_canonical(const uint32_t in[8]) {
    fr t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = in[i];
    return fr_mul(t, fr_r2());
}
__host__ __device__ inline fr fr_from_u64(uint64_t x) {
    uint32_t c[8] = {(uint32_t)x, (uint32_t)(x >> 32), 0, 0, 0, 0, 0, 0};
    return fr_from_canonical(c);
}

// a^(p-2); 0 -> 0
__host__ __device__ inline fr fr_inv(const fr &a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = fr_p_limb(i);
    e[0] -= 2;
    fr acc = fr_one(), base = a;
    for (int i = 0; i < 232; i++) {
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fr_mul(acc, base);
        base = fr_sqr(base);
    }
    return acc;
}

} // namespace dvp
