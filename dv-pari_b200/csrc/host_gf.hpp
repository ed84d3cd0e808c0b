// Host-side GF(2^233)/K-233 helpers of the product library (NOT the oracle): the O(1)-sized tails of
// the device algorithms -- the final ~232-step double-and-add over the per-bit partial sums an MSM
// leaves behind, and the 30-byte xsk233 encoding of the result (xsk233_encode, reached from
// /root/reference/src/curve.rs:93-100).  PCLMULQDQ when the build has it, the portable
// IMAD-style multiplier of gf233.cuh otherwise.
#pragma once
#include <cstring>
#include "k233.cuh"
#if defined(__PCLMUL__)
#include <immintrin.h>
#include <wmmintrin.h>
#endif

namespace dvp {
namespace host {

#if defined(__PCLMUL__)
// c[0..7] (466 bits, little-endian u64) -> reduced element; x^233 = x^74 + 1 on 64-bit words
inline gf hreduce64(uint64_t c[8]) {
    // word i (i >= 4) holds x^(64 i + k); 64 i - 233 = 64 (i - 4) + 23 and 64 i - 159 = 64 (i - 3) + 33
    for (int i = 7; i >= 4; i--) {
        const uint64_t t = c[i];
        c[i - 4] ^= t << 23;
        c[i - 3] ^= (t >> 41) ^ (t << 33);
        c[i - 2] ^= t >> 31;
    }
    const uint64_t t = c[3] >> 41; // bits 233..255
    c[0] ^= t;
    c[1] ^= t << 10; // x^74 = word 1, bit 10
    c[3] &= (1ull << 41) - 1;
    gf r;
    std::memcpy(r.v, c, 32);
    return r;
}
#endif
inline gf hmul(const gf &a, const gf &b) {
#if defined(__PCLMUL__)
    // 4 x 4 words of 64 bits, two Karatsuba levels: 9 carry-less multiplications
    uint64_t A[4], B[4];
    std::memcpy(A, a.v, 32);
    std::memcpy(B, b.v, 32);
    auto mul2 = [](uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t r[4]) {
        const __m128i x = _mm_set_epi64x((long long)a1, (long long)a0), y = _mm_set_epi64x((long long)b1, (long long)b0);
        const __m128i lo = _mm_clmulepi64_si128(x, y, 0x00), hi = _mm_clmulepi64_si128(x, y, 0x11);
        const __m128i sx = _mm_xor_si128(x, _mm_srli_si128(x, 8)), sy = _mm_xor_si128(y, _mm_srli_si128(y, 8));
        __m128i mid = _mm_clmulepi64_si128(sx, sy, 0x00);
        mid = _mm_xor_si128(mid, _mm_xor_si128(lo, hi));
        r[0] = (uint64_t)_mm_cvtsi128_si64(lo);
        r[1] = (uint64_t)_mm_extract_epi64(lo, 1) ^ (uint64_t)_mm_cvtsi128_si64(mid);
        r[2] = (uint64_t)_mm_cvtsi128_si64(hi) ^ (uint64_t)_mm_extract_epi64(mid, 1);
        r[3] = (uint64_t)_mm_extract_epi64(hi, 1);
    };
    uint64_t lo[4], hi[4], mid[4], c[8];
    mul2(A[0], A[1], B[0], B[1], lo);
    mul2(A[2], A[3], B[2], B[3], hi);
    mul2(A[0] ^ A[2], A[1] ^ A[3], B[0] ^ B[2], B[1] ^ B[3], mid);
    for (int i = 0; i < 4; i++) mid[i] ^= lo[i] ^ hi[i];
    c[0] = lo[0]; c[1] = lo[1];
    c[2] = lo[2] ^ mid[0]; c[3] = lo[3] ^ mid[1];
    c[4] = hi[0] ^ mid[2]; c[5] = hi[1] ^ mid[3];
    c[6] = hi[2]; c[7] = hi[3];
    return hreduce64(c);
#else
    return gf_mul(a, b);
#endif
}
inline gf hsqr(const gf &a) {
#if defined(__PCLMUL__)
    uint64_t A[4], c[8];
    std::memcpy(A, a.v, 32);
    for (int i = 0; i < 4; i++) {
        const __m128i x = _mm_cvtsi64_si128((long long)A[i]);
        const __m128i p = _mm_clmulepi64_si128(x, x, 0);
        c[2 * i] = (uint64_t)_mm_cvtsi128_si64(p);
        c[2 * i + 1] = (uint64_t)_mm_extract_epi64(p, 1);
    }
    return hreduce64(c);
#else
    return gf_sqr(a);
#endif
}
inline gf hsqr_n(gf a, int n) {
    for (int i = 0; i < n; i++) a = hsqr(a);
    return a;
}
inline gf hinv(const gf &a) {
    gf b1 = a;
    gf b2 = hmul(hsqr(b1), b1);
    gf b3 = hmul(hsqr(b2), b1);
    gf b6 = hmul(hsqr_n(b3, 3), b3);
    gf b7 = hmul(hsqr(b6), b1);
    gf b14 = hmul(hsqr_n(b7, 7), b7);
    gf b28 = hmul(hsqr_n(b14, 14), b14);
    gf b29 = hmul(hsqr(b28), b1);
    gf b58 = hmul(hsqr_n(b29, 29), b29);
    gf b116 = hmul(hsqr_n(b58, 58), b58);
    gf b232 = hmul(hsqr_n(b116, 116), b116);
    return hsqr(b232);
}

// Lopez-Dahab projective accumulator (x = X/Z, y = Y/Z^2); Z = 0 is infinity.  a = 0, b = 1.
struct LdPt {
    gf X, Y, Z;
};
inline LdPt ld_inf() {
    LdPt r;
    r.X = gf_one();
    r.Y = gf_zero();
    r.Z = gf_zero();
    return r;
}
inline LdPt ld_dbl(const LdPt &p) {
    if (gf_is_zero(p.Z)) return p;
    gf z2 = hsqr(p.Z), x2 = hsqr(p.X);
    LdPt o;
    o.Z = hmul(z2, x2);
    gf z4 = hsqr(z2), x4 = hsqr(x2);
    o.X = gf_add(x4, z4);
    gf t = gf_add(hsqr(p.Y), z4);
    o.Y = gf_add(hmul(z4, o.Z), hmul(o.X, t));
    return o;
}
inline LdPt ld_add_affine(const LdPt &p, const AffPt &q) {
    if (pt_is_inf(q)) return p;
    if (gf_is_zero(p.Z)) {
        LdPt r;
        r.X = q.x; r.Y = q.y; r.Z = gf_one();
        return r;
    }
    gf z2 = hsqr(p.Z);
    gf A = gf_add(p.Y, hmul(q.y, z2));
    gf B = gf_add(p.X, hmul(q.x, p.Z));
    if (gf_is_zero(B)) {
        if (gf_is_zero(A)) {
            LdPt r;
            r.X = q.x; r.Y = q.y; r.Z = gf_one();
            return ld_dbl(r);
        }
        return ld_inf();
    }
    gf Cc = hmul(p.Z, B);
    gf D = hmul(hsqr(B), Cc);
    LdPt o;
    o.Z = hsqr(Cc);
    gf E = hmul(A, Cc);
    o.X = gf_add(gf_add(hsqr(A), D), E);
    gf F = gf_add(o.X, hmul(q.x, o.Z));
    gf G = hmul(hsqr(o.Z), gf_add(q.x, q.y));
    o.Y = gf_add(hmul(gf_add(E, o.Z), F), G);
    return o;
}
inline AffPt ld_to_affine(const LdPt &p) {
    if (gf_is_zero(p.Z)) return pt_inf();
    gf zi = hinv(p.Z);
    AffPt r;
    r.x = hmul(p.X, zi);
    r.y = hmul(p.Y, hsqr(zi));
    return r;
}
inline AffPt aff_add(const AffPt &p, const AffPt &q) {
    gf d;
    int kind = pair_classify(p, q, d);
    if (kind >= 2) return pair_finish(p, q, kind, d);
    gf di = hinv(d);
    gf num = (kind == 1) ? p.y : gf_add(p.y, q.y);
    gf lam = hmul(num, di);
    if (kind == 1) lam = gf_add(lam, p.x);
    AffPt r;
    r.x = gf_add(gf_add(hsqr(lam), lam), gf_add(p.x, q.x));
    r.y = gf_add(gf_add(hmul(lam, gf_add(p.x, r.x)), r.x), p.y);
    return r;
}

// The K-233 point behind CurvePoint::generator() (xsk233_generator, /root/reference/src/curve.rs:84-91):
// the NIST / SEC 2 sect233k1 base point (restated convention).
inline AffPt k233_generator() {
    static const uint32_t gx[8] = {0xEFAD6126u, 0x0A4C9D6Eu, 0x19C26BF5u, 0x149563A4u, 0x29F22FF4u, 0x7E731AF1u, 0x32BA853Au, 0x00000172u};
    static const uint32_t gy[8] = {0x56FAE6A3u, 0x56E0C110u, 0xF18AEB9Bu, 0x27A8CD9Bu, 0x555A67C4u, 0x19B7F70Fu, 0x537DECE8u, 0x000001DBu};
    AffPt p;
    for (int i = 0; i < 8; i++) {
        p.x.v[i] = gx[i];
        p.y.v[i] = gy[i];
    }
    return p;
}

// xsk233_encode of the group element P + N for P in E[r] (or infinity -> neutral -> zeros):
// w = (y + 1)/x of P + N = (y + x + 1)/x of P, 233 bits little-endian.
inline void encode30(uint8_t out[30], const AffPt &p) {
    std::memset(out, 0, 30);
    if (pt_is_inf(p)) return;
    gf t = gf_add(gf_add(p.y, p.x), gf_one());
    gf w = hmul(t, hinv(p.x));
    for (int i = 0; i < 30; i++) out[i] = (uint8_t)(w.v[i >> 2] >> (8 * (i & 3)));
}

} // namespace host
} // namespace dvp
