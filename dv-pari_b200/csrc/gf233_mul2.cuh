// GF(2^233) product built from 32-bit multiply-adds only ("two streams").
//
// Why: on sm_100a an IMAD.WIDE takes the FMA-heavy pipe for 4 clocks AND an ALU issue slot, while a 32-bit IMAD
// dual-issues with LOP3 (profiles/README.md, r2c).  The low word of a 32x32 carry-less product comes from 16 IMAD
// on operands masked to every 4th bit; the HIGH word is the low word of the product of the bit-reversed operands
// (rev(A) * rev(B) = rev63(A*B)), so one word product is 32 IMAD and no IMAD.WIDE.  The reversed halves are carried
// through the Karatsuba levels in a second accumulator set and turned round once at the end (16 BREV).
// Measured (scripts/mulbench.cu, profiles/README.md r2e): 1.66e10 products/s against 1.45e10 for the IMAD.WIDE form;
// taking the reversed class words by BREV of the straight ones instead of masking is slower (BREV: 8 clocks per warp).
// Splitting the product into two calls of one out-of-line half (same code for both streams, 15 KB instead of 30) costs
// 11 % (1.48e10).
// In a low-word stream the k-th 4-bit field holds at most k+1 partial bits, so ANY two class products may share an
// integer accumulate (the only field that can reach 16 is the top one, whose carry leaves the word).
#pragma once

namespace dvp {
#ifdef __CUDACC__

struct w2 {
    uint32_t s, r; // straight word, bit-reversed word
};
__device__ __forceinline__ w2 operator^(const w2 &x, const w2 &y) { return w2{x.s ^ y.s, x.r ^ y.r}; }

__device__ __forceinline__ uint32_t gf_ml(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t gf_mla(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// low 32 bits of the carry-less product a * b
__device__ __forceinline__ uint32_t clmul32_lo(uint32_t a, uint32_t b) {
    const uint32_t a0 = a & 0x11111111u, a1 = a & 0x22222222u, a2 = a & 0x44444444u, a3 = a & 0x88888888u;
    const uint32_t b0 = b & 0x11111111u, b1 = b & 0x22222222u, b2 = b & 0x44444444u, b3 = b & 0x88888888u;
    const uint32_t p0 = gf_mla(a1, b3, gf_ml(a0, b0)), q0 = gf_mla(a3, b1, gf_ml(a2, b2));
    const uint32_t p1 = gf_mla(a2, b3, gf_ml(a0, b1)), q1 = gf_mla(a3, b2, gf_ml(a1, b0));
    const uint32_t p2 = gf_mla(a3, b3, gf_ml(a0, b2)), q2 = gf_mla(a2, b0, gf_ml(a1, b1));
    const uint32_t p3 = gf_mla(a1, b2, gf_ml(a0, b3)), q3 = gf_mla(a3, b0, gf_ml(a2, b1));
    const uint32_t x = ((p0 ^ q0) & 0x55555555u) | ((p1 ^ q1) & 0xaaaaaaaau);
    const uint32_t y = ((p2 ^ q2) & 0x55555555u) | ((p3 ^ q3) & 0xaaaaaaaau);
    return (x & 0x33333333u) | (y & 0xccccccccu);
}
// word product: .s = low word, .r = high word bit-reversed
__device__ __forceinline__ w2 clmul32_2s(const w2 &a, const w2 &b) { return w2{clmul32_lo(a.s, b.s), clmul32_lo(a.r, b.r)}; }

// cs[0..2] / cr[1..3] ^= (a0 + a1 X)(b0 + b1 X): cs[w] collects low words landing on word w, cr[w] the reversed high words
__device__ __forceinline__ void clmul_2w_acc2(uint32_t *cs, uint32_t *cr, const w2 &a0, const w2 &a1, const w2 &b0, const w2 &b1) {
    const w2 lo = clmul32_2s(a0, b0);
    const w2 hi = clmul32_2s(a1, b1);
    const w2 mid = clmul32_2s(a0 ^ a1, b0 ^ b1) ^ lo ^ hi;
    cs[0] ^= lo.s;
    cr[1] ^= lo.r;
    cs[1] ^= mid.s;
    cr[2] ^= mid.r;
    cs[2] ^= hi.s;
    cr[3] ^= hi.r;
}
// cs/cr[0..7] ^= A(4 words) * B(4 words)
__device__ __forceinline__ void clmul_4w_acc2(uint32_t *cs, uint32_t *cr, const w2 *a, const w2 *b) {
    uint32_t ts[4] = {0, 0, 0, 0}, tr[4] = {0, 0, 0, 0};
    clmul_2w_acc2(ts, tr, a[0], a[1], b[0], b[1]);
    cs[0] ^= ts[0]; cs[1] ^= ts[1]; cs[2] ^= ts[2] ^ ts[0]; cs[3] ^= ts[1]; cs[4] ^= ts[2];
    cr[1] ^= tr[1]; cr[2] ^= tr[2]; cr[3] ^= tr[3] ^ tr[1]; cr[4] ^= tr[2]; cr[5] ^= tr[3];
    uint32_t us[4] = {0, 0, 0, 0}, ur[4] = {0, 0, 0, 0};
    clmul_2w_acc2(us, ur, a[2], a[3], b[2], b[3]);
    cs[2] ^= us[0]; cs[3] ^= us[1]; cs[4] ^= us[2] ^ us[0]; cs[5] ^= us[1]; cs[6] ^= us[2];
    cr[3] ^= ur[1]; cr[4] ^= ur[2]; cr[5] ^= ur[3] ^ ur[1]; cr[6] ^= ur[2]; cr[7] ^= ur[3];
    clmul_2w_acc2(cs + 2, cr + 2, a[0] ^ a[2], a[1] ^ a[3], b[0] ^ b[2], b[1] ^ b[3]);
}
__device__ __forceinline__ gf gf_mul_dev2(const gf &a, const gf &b) {
    w2 A[8], B[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        A[i] = w2{a.v[i], __brev(a.v[i])};
        B[i] = w2{b.v[i], __brev(b.v[i]) << 1}; // rev(A) * (rev(B) x) mod x^32 = the high word of A*B, reversed
    }
    uint32_t cs[16], cr[16], ts[8], tr[8];
#pragma unroll
    for (int i = 0; i < 16; i++) cs[i] = cr[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) ts[i] = tr[i] = 0;
    clmul_4w_acc2(ts, tr, A, B);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cs[i] = ts[i]; cs[i + 4] = ts[i + 4] ^ ts[i]; cs[i + 8] = ts[i + 4];
        cr[i] = tr[i]; cr[i + 4] = tr[i + 4] ^ tr[i]; cr[i + 8] = tr[i + 4];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) ts[i] = tr[i] = 0;
    clmul_4w_acc2(ts, tr, A + 4, B + 4);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cs[i + 4] ^= ts[i]; cs[i + 8] ^= ts[i + 4] ^ ts[i]; cs[i + 12] = ts[i + 4];
        cr[i + 4] ^= tr[i]; cr[i + 8] ^= tr[i + 4] ^ tr[i]; cr[i + 12] = tr[i + 4];
    }
    w2 sa[4], sb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sa[i] = A[i] ^ A[i + 4];
        sb[i] = B[i] ^ B[i + 4];
    }
    clmul_4w_acc2(cs + 4, cr + 4, sa, sb);
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = cs[i] ^ __brev(cr[i]);
    return gf_reduce(c);
}

#endif
} // namespace dvp
