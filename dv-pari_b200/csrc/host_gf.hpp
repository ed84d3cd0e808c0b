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

inline gf hmul(const gf &a, const gf &b) {
#if defined(__PCLMUL__)
    uint64_t A[4], B[4], c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::memcpy(A, a.v, 32);
    std::memcpy(B, b.v, 32);
    for (int i = 0; i < 4; i++) {
        __m128i ai = _mm_cvtsi64_si128((long long)A[i]);
        for (int j = 0; j < 4; j++) {
            __m128i p = _mm_clmulepi64_si128(ai, _mm_cvtsi64_si128((long long)B[j]), 0);
            c[i + j] ^= (uint64_t)_mm_cvtsi128_si64(p);
            c[i + j + 1] ^= (uint64_t)_mm_extract_epi64(p, 1);
        }
    }
    uint32_t w[16];
    std::memcpy(w, c, 64);
    return gf_reduce(w);
#else
    return gf_mul(a, b);
#endif
}
inline gf hsqr(const gf &a) { return gf_sqr(a); }
inline gf hsqr_n(gf a, int n) {
    for (int i = 0; i < n; i++) a = gf_sqr(a);
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
