// Fr on 8 x 29-bit limbs for the ECFFT butterflies (sm_100a).
//
// The scalar field of sect233k1 is a 232-bit prime (/root/reference/src/curve.rs:16-22), so a value fits 8 limbs of
// 29 bits exactly.  Limb products are 58 bits wide: a 64-bit column takes the 16 products of a two-term dot
// product plus the Montgomery reduction terms without a single carry, so every partial product is one IMAD.WIDE
// with its free 64-bit accumulate and the carry chain of the 32-bit CIOS form (one IADD3 pair per product, the
// ALU-pipe bound of fr_mul) disappears.  One reduction serves both products of a butterfly row.
//
// The reduction divides by 2^232 (8 steps of 29 bits).  The memory format is ark's Montgomery form with R = 2^256
// (fr.cuh), so one operand of every product must carry an extra factor 2^-24: the ECFFT matrices are constants of
// the domain and are stored pre-scaled (fr29_prescale), which makes  dot2(m0', x0, m1', x1) = (m0 x0 + m1 x1) / R
// exactly the Montgomery-form result, fully reduced -- bit-identical to fr_add(fr_mul(m0,x0), fr_mul(m1,x1)).
#pragma once
#include "fr.cuh"

namespace dvp {

struct fr29 {
    uint32_t l[8];
};

#define DVP_M29 0x1fffffffu
// p = 2^231 + delta in 29-bit limbs: limbs 4..6 are zero, limb 7 = 2^28
#define DVP_P29_0 0x1173abdfu
#define DVP_P29_1 0x17d8d6afu
#define DVP_P29_2 0x056f351bu
#define DVP_P29_3 0x0d3ab772u
#define DVP_NP29 0x0c382fe1u /* -p^-1 mod 2^29 */

// 2^232 mod p as a "Montgomery" constant: fr_mul(a, this) = a * 2^-24 (value-wise), the pre-scaling of a matrix entry
__host__ __device__ __forceinline__ fr fr_two232() {
    fr r;
    r.v[0] = 0x0e8c5421u; r.v[1] = 0x9104e52au; r.v[2] = 0x46ea432bu; r.v[3] = 0xfff962a4u;
    r.v[4] = 0xffffffffu; r.v[5] = 0xffffffffu; r.v[6] = 0xffffffffu; r.v[7] = 0x0000007fu;
    return r;
}

// 8 x 32-bit little-endian (value < 2^232) -> 8 x 29-bit limbs
__host__ __device__ __forceinline__ fr29 fr29_from_fr(const fr &a) {
    fr29 r;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int bit = 29 * k, w = bit >> 5, s = bit & 31;
        uint32_t v = a.v[w] >> s;
        if (s > 3 && w + 1 < 8) v |= a.v[w + 1] << (32 - s);
        r.l[k] = v & DVP_M29;
    }
    return r;
}
// normalised limbs (each < 2^29, value < 2^232) -> 8 x 32-bit
__host__ __device__ __forceinline__ fr fr_from_fr29(const fr29 &a) {
    fr r;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int bit = 32 * w, k = bit / 29, o = bit - 29 * k; // word w starts inside limb k at offset o
        uint32_t v = a.l[k] >> o;
        if (k + 1 < 8) v |= a.l[k + 1] << (29 - o);
        if (k + 2 < 8 && 58 - o < 32) v |= a.l[k + 2] << (58 - o);
        r.v[w] = v;
    }
    return r;
}
__host__ __device__ __forceinline__ fr29 fr29_prescale(const fr &m) { return fr29_from_fr(fr_mul(m, fr_two232())); }

// (m0 x0 + m1 x1) / 2^232 mod p, fully reduced; all operands normalised (limbs < 2^29, values < p)
__host__ __device__ __forceinline__ fr29 fr29_dot2(const fr29 &m0, const fr29 &x0, const fr29 &m1, const fr29 &x1) {
    uint64_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c[i + j] += (uint64_t)m0.l[i] * x0.l[j];
            c[i + j] += (uint64_t)m1.l[i] * x1.l[j];
        }
    // Montgomery reduction, 29 bits per step; column sums stay below 2^63
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = ((uint32_t)c[i] * DVP_NP29) & DVP_M29;
        c[i] += (uint64_t)m * DVP_P29_0;
        c[i + 1] += (uint64_t)m * DVP_P29_1;
        c[i + 2] += (uint64_t)m * DVP_P29_2;
        c[i + 3] += (uint64_t)m * DVP_P29_3;
        c[i + 7] += (uint64_t)m << 28;
        c[i + 1] += c[i] >> 29; // the low 29 bits of c[i] are zero now
    }
    // carry propagation of columns 8..15; the value is < 2p < 2^233, the top limb keeps the extra bit
    fr29 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t t = c[8 + k] + carry;
        r.l[k] = k < 7 ? (uint32_t)t & DVP_M29 : (uint32_t)t;
        carry = t >> 29;
    }
    // conditional subtraction of p
    uint32_t d[8], borrow = 0;
    const uint32_t pl[8] = {DVP_P29_0, DVP_P29_1, DVP_P29_2, DVP_P29_3, 0u, 0u, 0u, 0x10000000u};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t t = r.l[k] - pl[k] - borrow;
        borrow = t >> 31;
        d[k] = k < 7 ? t & DVP_M29 : t;
    }
    if (!borrow) {
#pragma unroll
        for (int k = 0; k < 8; k++) r.l[k] = d[k];
    }
    return r;
}

// sum_{t < N} m_t x_t / 2^232 mod p for N <= 3 products, fully reduced (column sums stay below 2^63; the result of the
// reduction is below 2.5 p, hence two conditional subtractions)
template <int N>
__host__ __device__ __forceinline__ fr29 fr29_dotn(const fr29 *m, const fr29 *x) {
    uint64_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
#pragma unroll
    for (int t = 0; t < N; t++)
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) c[i + j] += (uint64_t)m[t].l[i] * x[t].l[j];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t q = ((uint32_t)c[i] * DVP_NP29) & DVP_M29;
        c[i] += (uint64_t)q * DVP_P29_0;
        c[i + 1] += (uint64_t)q * DVP_P29_1;
        c[i + 2] += (uint64_t)q * DVP_P29_2;
        c[i + 3] += (uint64_t)q * DVP_P29_3;
        c[i + 7] += (uint64_t)q << 28;
        c[i + 1] += c[i] >> 29;
    }
    fr29 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t t = c[8 + k] + carry;
        r.l[k] = k < 7 ? (uint32_t)t & DVP_M29 : (uint32_t)t;
        carry = t >> 29;
    }
    const uint32_t pl[8] = {DVP_P29_0, DVP_P29_1, DVP_P29_2, DVP_P29_3, 0u, 0u, 0u, 0x10000000u};
#pragma unroll
    for (int rep = 0; rep < (N > 1 ? 2 : 1); rep++) {
        uint32_t d[8], borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t t = r.l[k] - pl[k] - borrow;
            borrow = t >> 31;
            d[k] = k < 7 ? t & DVP_M29 : t;
        }
        if (!borrow) {
#pragma unroll
            for (int k = 0; k < 8; k++) r.l[k] = d[k];
        }
    }
    return r;
}
// x0 + m x1 / 2^232 mod p, fully reduced (m pre-scaled like a matrix entry, all operands normalised and < p): the row of a
// normalised butterfly (1, m).  The addend enters the accumulator as x0 * 2^232, i.e. on columns 8..15, so the row costs one
// product and one reduction; the reduction's result is below p^2 / 2^232 + p + x0 < 2.5 p, hence two conditional subtractions.
__host__ __device__ __forceinline__ fr29 fr29_muladd(const fr29 &m, const fr29 &x1, const fr29 &x0) {
    uint64_t c[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[i] = 0;
        c[8 + i] = x0.l[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) c[i + j] += (uint64_t)m.l[i] * x1.l[j];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t q = ((uint32_t)c[i] * DVP_NP29) & DVP_M29;
        c[i] += (uint64_t)q * DVP_P29_0;
        c[i + 1] += (uint64_t)q * DVP_P29_1;
        c[i + 2] += (uint64_t)q * DVP_P29_2;
        c[i + 3] += (uint64_t)q * DVP_P29_3;
        c[i + 7] += (uint64_t)q << 28;
        c[i + 1] += c[i] >> 29;
    }
    fr29 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t t = c[8 + k] + carry;
        r.l[k] = k < 7 ? (uint32_t)t & DVP_M29 : (uint32_t)t;
        carry = t >> 29;
    }
    const uint32_t pl[8] = {DVP_P29_0, DVP_P29_1, DVP_P29_2, DVP_P29_3, 0u, 0u, 0u, 0x10000000u};
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
        uint32_t d[8], borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t t = r.l[k] - pl[k] - borrow;
            borrow = t >> 31;
            d[k] = k < 7 ? t & DVP_M29 : t;
        }
        if (!borrow) {
#pragma unroll
            for (int k = 0; k < 8; k++) r.l[k] = d[k];
        }
    }
    return r;
}
// ---- semi-reduced rows --------------------------------------------------------------------------------------------
// Between the levels of an extend the values only need to stay below 2^232 (< 2p, limbs normalised): the columns of the
// next product still cannot overflow, and a full reduction is done once, by the last level.  After the Montgomery
// reduction a row is below 2^233 + 2^232 (muladd: p + 2^232 + p; dot2: 3p), so with k = floor(y / 2^231) read off the
// top limb, y - max(k - 1, 0) p lies in (0, 2^232): one branch-free pass instead of two conditional subtractions.
__host__ __device__ __forceinline__ fr29 fr29_finish_semi(uint64_t (&c)[16]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t q = ((uint32_t)c[i] * DVP_NP29) & DVP_M29;
        c[i] += (uint64_t)q * DVP_P29_0;
        c[i + 1] += (uint64_t)q * DVP_P29_1;
        c[i + 2] += (uint64_t)q * DVP_P29_2;
        c[i + 3] += (uint64_t)q * DVP_P29_3;
        c[i + 7] += (uint64_t)q << 28;
        c[i + 1] += c[i] >> 29;
    }
    fr29 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint64_t t = c[8 + k] + carry;
        r.l[k] = k < 7 ? (uint32_t)t & DVP_M29 : (uint32_t)t;
        carry = t >> 29;
    }
    const uint32_t k = r.l[7] >> 28, kk = k ? k - 1 : 0; // k <= 4, so kk p[j] < 3 * 2^29 and every t below is > -2^31
    const uint32_t pl[4] = {DVP_P29_0, DVP_P29_1, DVP_P29_2, DVP_P29_3};
    int32_t borrow = 0; // 0 or negative
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const int32_t t = (int32_t)r.l[j] - (j < 4 ? (int32_t)(kk * pl[j]) : 0) + borrow;
        r.l[j] = (uint32_t)t & DVP_M29;
        borrow = t >> 29; // arithmetic shift: floor(t / 2^29)
    }
    r.l[7] = (uint32_t)((int32_t)r.l[7] - (int32_t)(kk << 28) + borrow);
    return r;
}
// x0 + m x1 / 2^232, semi-reduced; m pre-scaled and < p, x0 and x1 semi-reduced
__host__ __device__ __forceinline__ fr29 fr29_muladd_semi(const fr29 &m, const fr29 &x1, const fr29 &x0) {
    uint64_t c[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[i] = 0;
        c[8 + i] = x0.l[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) c[i + j] += (uint64_t)m.l[i] * x1.l[j];
    return fr29_finish_semi(c);
}
// (m0 x0 + m1 x1) / 2^232, semi-reduced; m0, m1 pre-scaled and < p, x0 and x1 semi-reduced
__host__ __device__ __forceinline__ fr29 fr29_dot2_semi(const fr29 &m0, const fr29 &x0, const fr29 &m1, const fr29 &x1) {
    uint64_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c[i + j] += (uint64_t)m0.l[i] * x0.l[j];
            c[i + j] += (uint64_t)m1.l[i] * x1.l[j];
        }
    return fr29_finish_semi(c);
}
// a + b mod p on normalised limbs (both < p)
__host__ __device__ __forceinline__ fr29 fr29_add(const fr29 &a, const fr29 &b) {
    fr29 r;
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t t = a.l[k] + b.l[k] + carry;
        r.l[k] = k < 7 ? t & DVP_M29 : t;
        carry = k < 7 ? t >> 29 : 0;
    }
    const uint32_t pl[8] = {DVP_P29_0, DVP_P29_1, DVP_P29_2, DVP_P29_3, 0u, 0u, 0u, 0x10000000u};
    uint32_t d[8], borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t t = r.l[k] - pl[k] - borrow;
        borrow = t >> 31;
        d[k] = k < 7 ? t & DVP_M29 : t;
    }
    if (!borrow) {
#pragma unroll
        for (int k = 0; k < 8; k++) r.l[k] = d[k];
    }
    return r;
}

} // namespace dvp
