// xsk233 30-byte point codec on the device.
//
// Replaces xsk233_decode / xsk233_encode of crate xs233-sys as reached through
// CurvePoint::{from_bytes,to_bytes} (/root/reference/src/curve.rs:93-109) and the bulk decode in
// read_point_vec_from_file (/root/reference/src/io_utils.rs:187-239).
//
// Model (T. Pornin, ePrint 2022/1325): the group is { P + N : P in E[r] }, N = (0,1), and an element
// Q = (x, y) is encoded as w = (y + 1)/x (w^2 + w = x + 1/x on K-233).  We keep the E[r]
// representative P = Q + N, whose w-coordinate is w + 1:  y_P = x_P (w + 1) + 1.
// The neutral (P = infinity) encodes as zero.  Byte-parity with upstream is unpinned (no vectors in
// the reference), the group law is not.
#pragma once
#include "k233.cuh"

namespace dvp {

// trace over GF(2) for x^233 + x^74 + 1: Tr(a) = a_0 + a_159
__host__ __device__ __forceinline__ uint32_t gf_trace(const gf &a) { return (a.v[0] ^ (a.v[4] >> 31)) & 1u; }

// z with z^2 + z = a, valid when Tr(a) = 0: z = sum_{i=0}^{116} a^(4^i)
__host__ __device__ inline gf gf_halftrace(const gf &a) {
    gf z = a;
#pragma unroll 1
    for (int i = 0; i < 116; i++) z = gf_add(gf_sqr(gf_sqr(z)), a);
    return z;
}

__host__ __device__ inline gf gf_from_le30(const uint8_t *in, bool &ok) {
    gf w = gf_zero();
    for (int i = 0; i < 30; i++) w.v[i >> 2] |= (uint32_t)in[i] << (8 * (i & 3));
    ok = (w.v[7] >> 9) == 0;
    return w;
}
__host__ __device__ inline void gf_to_le30(uint8_t *out, const gf &w) {
    for (int i = 0; i < 30; i++) out[i] = (uint8_t)(w.v[i >> 2] >> (8 * (i & 3)));
}

// returns false if the bytes do not name a group element
__host__ __device__ inline bool xsk233_decode_pt(const uint8_t *in, AffPt &p) {
    p = pt_inf();
    bool ok;
    const gf w = gf_from_le30(in, ok);
    if (!ok) return false;
    if (gf_is_zero(w)) return true; // neutral
    const gf d = gf_add(gf_sqr(w), w);
    if (gf_is_zero(d)) return false; // w = 1: x = 1, a point of order 4
    const gf e = gf_sqr(gf_inv(d));
    if (gf_trace(e)) return false; // x^2 + d x + 1 has no root
    const gf f = gf_halftrace(e);
    const gf x1 = gf_mul(d, f), x2 = gf_add(x1, d);
    if (gf_trace(x1)) return false; // not a double: outside E[r] u (E[r] + N)
    // T1 = (x1, x1 w + 1).  T1 lies in E[r] + N iff its halves are not doubles, i.e. Tr(x_half) = 1,
    // with x_half^2 = y + (lam + 1) x, lam^2 + lam = x.
    const gf y1 = gf_add(gf_mul(x1, w), gf_one());
    const gf lam = gf_halftrace(x1);
    const gf u2 = gf_add(y1, gf_mul(gf_add(lam, gf_one()), x1));
    const gf xp = gf_trace(u2) ? x2 : x1;
    p.x = xp;
    p.y = gf_add(gf_mul(xp, gf_add(w, gf_one())), gf_one());
    return true;
}

__host__ __device__ inline void xsk233_encode_pt(uint8_t *out, const AffPt &p) {
    gf w = gf_zero();
    if (!pt_is_inf(p)) w = gf_mul(gf_add(gf_add(p.y, p.x), gf_one()), gf_inv(p.x));
    gf_to_le30(out, w);
}

// complete affine addition with its own inversion (O(1) uses only)
__host__ __device__ inline AffPt pt_add_slow(const AffPt &a, const AffPt &b) {
    gf d;
    const int kind = pair_classify(a, b, d);
    return pair_finish(a, b, kind, kind < 2 ? gf_inv(d) : d);
}

} // namespace dvp
