// sect233k1 affine point layer for sm_100a: y^2 + xy = x^3 + 1 over GF(2^233).
//
// Replaces the group layer of crate xs233-sys as used by multi_scalar_mul
// (/root/reference/src/curve.rs:141-158).  All bulk work runs on ordinary K-233 points of the
// prime-order subgroup E[r]; the xsk233 <-> K-233 map (Q = P + N) lives in the codec only.
// A point is 64 bytes (x then y, 8 x u32 each); (0,0) marks the point at infinity -- x = 0 occurs
// on the curve only for N = (0,1), which is never a member of E[r].
#pragma once
#include "gf233.cuh"

namespace dvp {

struct __align__(16) AffPt {
    gf x, y;
};

__host__ __device__ __forceinline__ AffPt pt_inf() {
    AffPt p;
    p.x = gf_zero();
    p.y = gf_zero();
    return p;
}
__host__ __device__ __forceinline__ bool pt_is_inf(const AffPt &p) { return gf_is_zero(p.x); }
__host__ __device__ __forceinline__ AffPt pt_neg(const AffPt &p) {
    AffPt r;
    r.x = p.x;
    r.y = gf_add(p.x, p.y);
    return r;
}

#ifdef __CUDACC__
__device__ __forceinline__ gf gf_load(const gf *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    gf r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void gf_store(gf *p, const gf &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
__device__ __forceinline__ AffPt pt_load(const AffPt *p) {
    AffPt r;
    r.x = gf_load(&p->x);
    r.y = gf_load(&p->y);
    return r;
}
__device__ __forceinline__ void pt_store(AffPt *p, const AffPt &v) {
    gf_store(&p->x, v.x);
    gf_store(&p->y, v.y);
}
#endif

// One affine addition P1 + P2, complete over E[r] u {inf}, split around the shared inversion:
//   kind 0: chord        d = x1 + x2, lambda = (y1 + y2)/d
//   kind 1: tangent      d = x1,      lambda = x1 + y1/d
//   kind 2: result = P1 (P2 = inf)    kind 3: result = P2 (P1 = inf)    kind 4: result = inf (P2 = -P1)
// For kinds 2..4 the denominator is 1 so that it can sit in a batched (Montgomery-trick) inversion.
__host__ __device__ __forceinline__ int pair_classify(const AffPt &p1, const AffPt &p2, gf &d) {
    const bool i1 = pt_is_inf(p1), i2 = pt_is_inf(p2);
    d = gf_add(p1.x, p2.x);
    int kind = 0;
    if (i1 | i2) {
        kind = i2 ? 2 : 3;
        d = gf_one();
    } else if (gf_is_zero(d)) {
        if (gf_eq(p1.y, p2.y)) {
            kind = 1;
            d = p1.x;
        } else {
            kind = 4;
            d = gf_one();
        }
    }
    return kind;
}
// finish with dinv = 1/d
__host__ __device__ __forceinline__ AffPt pair_finish(const AffPt &p1, const AffPt &p2, int kind, const gf &dinv) {
    if (kind == 2) return p1;
    if (kind == 3) return p2;
    if (kind == 4) return pt_inf();
    gf num = (kind == 1) ? p1.y : gf_add(p1.y, p2.y);
    gf lam = gf_mul(num, dinv);
    if (kind == 1) lam = gf_add(lam, p1.x);
    AffPt r;
    // x3 = lam^2 + lam + x1 + x2 (a = 0); for the tangent x1 + x2 = 0
    r.x = gf_add(gf_add(gf_sqr(lam), lam), gf_add(p1.x, p2.x));
    // y3 = lam (x1 + x3) + x3 + y1
    r.y = gf_add(gf_add(gf_mul(lam, gf_add(p1.x, r.x)), r.x), p1.y);
    return r;
}

} // namespace dvp
