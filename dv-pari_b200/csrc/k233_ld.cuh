// Lopez-Dahab projective points on sect233k1 (a = 0, b = 1): x = X/Z, y = Y/Z^2, Z = 0 is infinity.
//
// Used where a chain of dependent additions is too short to amortise a batched inversion: the two reduction
// levels of the MSM (msm.cu, sum_b (b+1) B_b) are regular trees of depth <= 7 over a few thousand points, and an
// inversion-free addition (14M + 5S, ~27 us of single-warp latency) beats affine + inversion (~75 us) per level.
// The addition is complete over E[r] u {inf} (equal and opposite operands are detected by cross-multiplication),
// so the affine result is the exact group element -- the same bytes as any other evaluation order.
#pragma once
#include "k233.cuh"

namespace dvp {

struct LdPt {
    gf X, Y, Z;
};

__host__ __device__ __forceinline__ LdPt ld_infinity() {
    LdPt r;
    r.X = gf_one();
    r.Y = gf_zero();
    r.Z = gf_zero();
    return r;
}
__host__ __device__ __forceinline__ LdPt ld_from_affine(const AffPt &p) {
    LdPt r;
    r.X = p.x;
    r.Y = p.y;
    r.Z = gf_one();
    if (pt_is_inf(p)) r = ld_infinity();
    return r;
}

template <class MUL> __host__ __device__ __forceinline__ LdPt ld_double_t(const LdPt &p, MUL mul) {
    if (gf_is_zero(p.Z)) return p;
    const gf z2 = gf_sqr(p.Z), x2 = gf_sqr(p.X);
    LdPt o;
    o.Z = mul(z2, x2);
    const gf z4 = gf_sqr(z2), x4 = gf_sqr(x2);
    o.X = gf_add(x4, z4);
    const gf t = gf_add(gf_sqr(p.Y), z4);
    o.Y = gf_add(mul(z4, o.Z), mul(o.X, t));
    return o;
}

// P1 + P2.  With A = X1 Z2, B = X2 Z1, G = Y1 Z2^2, H = Y2 Z1^2, E = A + B, I = G + H, F = Z1 Z2 E:
//   Z3 = F^2,  X3 = I^2 + I F + E^2 F,  Y3 = I F (A E F + X3) + Z3 (X3 + G E^2)
template <class MUL> __host__ __device__ __forceinline__ LdPt ld_add_t(const LdPt &p1, const LdPt &p2, MUL mul) {
    if (gf_is_zero(p1.Z)) return p2;
    if (gf_is_zero(p2.Z)) return p1;
    const gf A = mul(p1.X, p2.Z), B = mul(p2.X, p1.Z);
    const gf G = mul(p1.Y, gf_sqr(p2.Z)), H = mul(p2.Y, gf_sqr(p1.Z));
    const gf E = gf_add(A, B), I = gf_add(G, H);
    if (gf_is_zero(E)) {
        if (gf_is_zero(I)) return ld_double_t(p1, mul); // same point
        return ld_infinity();                            // opposite points
    }
    const gf F = mul(mul(p1.Z, p2.Z), E);
    const gf E2 = gf_sqr(E), IF = mul(I, F);
    LdPt o;
    o.Z = gf_sqr(F);
    o.X = gf_add(gf_add(gf_sqr(I), IF), mul(E2, F));
    const gf t1 = gf_add(mul(mul(A, E), F), o.X);
    const gf t2 = gf_add(o.X, mul(G, E2));
    o.Y = gf_add(mul(IF, t1), mul(o.Z, t2));
    return o;
}

struct GfMulInline {
    __host__ __device__ __forceinline__ gf operator()(const gf &a, const gf &b) const { return gf_mul(a, b); }
};
__host__ __device__ __forceinline__ LdPt ld_add(const LdPt &a, const LdPt &b) { return ld_add_t(a, b, GfMulInline()); }
__host__ __device__ __forceinline__ LdPt ld_double(const LdPt &a) { return ld_double_t(a, GfMulInline()); }
// zinv = 1/Z (any value for infinity)
__host__ __device__ __forceinline__ AffPt ld_to_affine_with(const LdPt &p, const gf &zinv) {
    if (gf_is_zero(p.Z)) return pt_inf();
    AffPt r;
    r.x = gf_mul(p.X, zinv);
    r.y = gf_mul(p.Y, gf_sqr(zinv));
    return r;
}

} // namespace dvp
