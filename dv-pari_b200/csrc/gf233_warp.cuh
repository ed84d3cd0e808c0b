// Warp-cooperative GF(2^233) arithmetic for the latency-bound steps of the MSM (sm_100a).
//
// A field multiplication done by ONE thread is a ~1 600-instruction dependent program: 1.9 us for a lone warp, and the
// Itoh-Tsujii inversion that every tree round of the MSM waits for is ten of them plus the squarings (46 us).  When a
// round is small there is nothing to overlap that latency with.  Here the 32 lanes of a warp share ONE operation:
//   multiplication   the 27 word products of the three Karatsuba levels go to 27 lanes (one 32x32 carry-less product
//                    each, gf233.cuh), every lane spreads its product over the output words it feeds (three mask
//                    stages, the transposes of the operand selection), REDUX.XOR sums the 16 output words over the
//                    warp, every lane reduces mod x^233 + x^74 + 1
//   x -> x^(2^k)     GF(2)-linear: lane l fetches the table row of byte l (30 lanes), REDUX.XOR folds them
//   inversion        Itoh-Tsujii (chain 1,2,3,6,7,14,28,29,58,116,232) on those two
// All lanes pass the same operands and receive the same result (warp-uniform values); every lane of the warp must call.
#pragma once
#include "gf233.cuh"

namespace dvp {

// per-lane constants of the cooperative multiplication: lane l < 27 is leaf (d2, d1, d0), l = 9 d2 + 3 d1 + d0, of the
// Karatsuba tree; digit 0 = low halves, 1 = high halves, 2 = their sum.  L_k / H_k = all-ones masks for digit 0 / 1.
struct WarpMulCtx {
    uint32_t nl[3], nh[3]; // operand selection: take the low half unless digit = 1, the high half unless digit = 0
    uint32_t L[3], H[3];   // placement: also at offset 0 (digit 0) / also at offset 2h (digit 1)
    uint32_t live;         // all-ones for lanes 0..26
};
__device__ __forceinline__ WarpMulCtx warp_mul_ctx() {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t d[3] = {lane % 3, (lane / 3) % 3, lane / 9};
    WarpMulCtx c;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        c.nl[k] = d[k] != 1 ? 0xffffffffu : 0u;
        c.nh[k] = d[k] != 0 ? 0xffffffffu : 0u;
        c.L[k] = d[k] == 0 ? 0xffffffffu : 0u;
        c.H[k] = d[k] == 1 ? 0xffffffffu : 0u;
    }
    c.live = lane < 27 ? 0xffffffffu : 0u;
    return c;
}

// the operand word of this lane's leaf: three halvings of the 8 words
__device__ __forceinline__ uint32_t warp_mul_pick(const gf &a, const WarpMulCtx &c) {
    uint32_t u[4], v[2];
#pragma unroll
    for (int j = 0; j < 4; j++) u[j] = (a.v[j] & c.nl[2]) ^ (a.v[j + 4] & c.nh[2]);
#pragma unroll
    for (int j = 0; j < 2; j++) v[j] = (u[j] & c.nl[1]) ^ (u[j + 2] & c.nh[1]);
    return (v[0] & c.nl[0]) ^ (v[1] & c.nh[0]);
}

// lane src_lane's element to every lane
__device__ __forceinline__ gf gf_bcast(const gf &v, int src_lane) {
    gf r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(0xffffffffu, v.v[k], src_lane);
    return r;
}

// a * b, operands and result warp-uniform
__device__ __forceinline__ gf gf_mul_warp(const gf &a, const gf &b, const WarpMulCtx &c) {
    const uint32_t x = warp_mul_pick(a, c) & c.live, y = warp_mul_pick(b, c);
    const uint64_t p = clmul32_dev(x, y);
    const uint32_t p0 = (uint32_t)p, p1 = (uint32_t)(p >> 32);
    // a level with half size h turns (lo, hi, mid) into lo (1 + X^h) + hi (X^h + X^2h) + mid X^h: every leaf lands at
    // offset h, the lo leaves also at 0, the hi leaves also at 2h
    uint32_t r[4], s[8], t[16];
    r[0] = p0 & c.L[0];
    r[1] = p0 ^ (p1 & c.L[0]);
    r[2] = p1 ^ (p0 & c.H[0]);
    r[3] = p1 & c.H[0];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        s[j] = r[j] & c.L[1];
        s[2 + j] = r[j] ^ (r[2 + j] & c.L[1]);
        s[4 + j] = r[2 + j] ^ (r[j] & c.H[1]);
        s[6 + j] = r[2 + j] & c.H[1];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        t[j] = s[j] & c.L[2];
        t[4 + j] = s[j] ^ (s[4 + j] & c.L[2]);
        t[8 + j] = s[4 + j] ^ (s[j] & c.H[2]);
        t[12 + j] = s[4 + j] & c.H[2];
    }
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = __reduce_xor_sync(0xffffffffu, t[k]);
    return gf_reduce(w);
}

// x^(2^k) through the byte-indexed table of that k (30 x 256 rows of 32 bytes, msm.cu): one row per lane
__device__ __forceinline__ gf gf_msqr_tab_warp(const gf &x, const gf *__restrict__ tab) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t pos = lane < 30 ? lane : 0;
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) word |= x.v[j] & ((pos >> 2) == (uint32_t)j ? 0xffffffffu : 0u);
    const uint32_t b = (word >> (8 * (pos & 3))) & 255u;
    const uint4 *row = reinterpret_cast<const uint4 *>(tab + pos * 256 + b);
    uint4 lo = __ldg(row), hi = __ldg(row + 1);
    if (lane >= 30) lo = hi = make_uint4(0, 0, 0, 0);
    gf r;
    r.v[0] = __reduce_xor_sync(0xffffffffu, lo.x);
    r.v[1] = __reduce_xor_sync(0xffffffffu, lo.y);
    r.v[2] = __reduce_xor_sync(0xffffffffu, lo.z);
    r.v[3] = __reduce_xor_sync(0xffffffffu, lo.w);
    r.v[4] = __reduce_xor_sync(0xffffffffu, hi.x);
    r.v[5] = __reduce_xor_sync(0xffffffffu, hi.y);
    r.v[6] = __reduce_xor_sync(0xffffffffu, hi.z);
    r.v[7] = __reduce_xor_sync(0xffffffffu, hi.w);
    return r;
}

// 1/a (0 -> 0); tabs = the five tables for k = 7, 14, 29, 58, 116, `stride` rows apart
__device__ __forceinline__ gf gf_inv_warp(const gf &a, const gf *__restrict__ tabs, size_t stride, const WarpMulCtx &c) {
    const gf b2 = gf_mul_warp(gf_sqr(a), a, c);
    const gf b3 = gf_mul_warp(gf_sqr(b2), a, c);
    const gf b6 = gf_mul_warp(gf_sqr(gf_sqr(gf_sqr(b3))), b3, c);
    const gf b7 = gf_mul_warp(gf_sqr(b6), a, c);
    const gf b14 = gf_mul_warp(gf_msqr_tab_warp(b7, tabs), b7, c);
    const gf b28 = gf_mul_warp(gf_msqr_tab_warp(b14, tabs + stride), b14, c);
    const gf b29 = gf_mul_warp(gf_sqr(b28), a, c);
    const gf b58 = gf_mul_warp(gf_msqr_tab_warp(b29, tabs + 2 * stride), b29, c);
    const gf b116 = gf_mul_warp(gf_msqr_tab_warp(b58, tabs + 3 * stride), b58, c);
    const gf b232 = gf_mul_warp(gf_msqr_tab_warp(b116, tabs + 4 * stride), b116, c);
    return gf_sqr(b232);
}

} // namespace dvp
