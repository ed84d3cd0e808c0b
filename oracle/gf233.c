/* ORACLE (test infrastructure only). See gf233.h for provenance. */
#include "gf233.h"
#include <string.h>
#if defined(__PCLMUL__)
#include <immintrin.h>
#include <wmmintrin.h>
#endif

const gf_t GF_ZERO = {{0, 0, 0, 0}};
const gf_t GF_ONE = {{1, 0, 0, 0}};
static int g_portable = 0;
void gf_set_portable(int on) { g_portable = on; }

void gf_add(gf_t *r, const gf_t *a, const gf_t *b) {
    for (int i = 0; i < 4; i++) r->w[i] = a->w[i] ^ b->w[i];
}
int gf_is_zero(const gf_t *a) { return (a->w[0] | a->w[1] | a->w[2] | a->w[3]) == 0; }
int gf_eq(const gf_t *a, const gf_t *b) { return memcmp(a->w, b->w, 32) == 0; }

/* c[0..7] (466 significant bits) -> reduced mod x^233 + x^74 + 1 */
static inline __attribute__((always_inline)) void gf_reduce(gf_t *r, uint64_t c[8]) {
#pragma GCC unroll 4
    for (int i = 7; i >= 4; i--) {
        uint64_t t = c[i];
        /* x^(64i+k) = x^(64(i-4)+23+k) + x^(64(i-3)+33+k) */
        c[i - 4] ^= t << 23;
        c[i - 3] ^= t >> 41;
        c[i - 3] ^= t << 33;
        c[i - 2] ^= t >> 31;
    }
    uint64_t t = c[3] >> 41; /* bits 233.. of the low half */
    c[0] ^= t;
    c[1] ^= t << 10;
    c[3] &= (1ULL << 41) - 1;
    r->w[0] = c[0]; r->w[1] = c[1]; r->w[2] = c[2]; r->w[3] = c[3];
}

static void clmul64_portable(uint64_t *lo, uint64_t *hi, uint64_t a, uint64_t b) {
    uint64_t l = 0, h = 0;
    for (int i = 0; i < 64; i++) {
        if ((a >> i) & 1) {
            l ^= b << i;
            if (i) h ^= b >> (64 - i);
        }
    }
    *lo = l;
    *hi = h;
}

static void mul_portable(uint64_t c[8], const gf_t *a, const gf_t *b) {
    memset(c, 0, 64);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            uint64_t l, h;
            clmul64_portable(&l, &h, a->w[i], b->w[j]);
            c[i + j] ^= l;
            c[i + j + 1] ^= h;
        }
}

#if defined(__PCLMUL__)
/* 128 x 128 -> 256 bits carry-less, Karatsuba: 3 PCLMULQDQ */
static inline __attribute__((always_inline)) void clmul128(__m128i *lo, __m128i *hi, __m128i a, __m128i b) {
    const __m128i t0 = _mm_clmulepi64_si128(a, b, 0x00), t2 = _mm_clmulepi64_si128(a, b, 0x11);
    const __m128i am = _mm_xor_si128(a, _mm_srli_si128(a, 8)), bm = _mm_xor_si128(b, _mm_srli_si128(b, 8));
    const __m128i t1 = _mm_xor_si128(_mm_clmulepi64_si128(am, bm, 0x00), _mm_xor_si128(t0, t2));
    *lo = _mm_xor_si128(t0, _mm_slli_si128(t1, 8));
    *hi = _mm_xor_si128(t2, _mm_srli_si128(t1, 8));
}
static void mul_pclmul(uint64_t c[8], const gf_t *a, const gf_t *b) {
    /* two levels of Karatsuba: 9 products instead of 16 */
    const __m128i a0 = _mm_loadu_si128((const __m128i *)&a->w[0]), a1 = _mm_loadu_si128((const __m128i *)&a->w[2]);
    const __m128i b0 = _mm_loadu_si128((const __m128i *)&b->w[0]), b1 = _mm_loadu_si128((const __m128i *)&b->w[2]);
    __m128i p0l, p0h, p2l, p2h, p1l, p1h;
    clmul128(&p0l, &p0h, a0, b0);
    clmul128(&p2l, &p2h, a1, b1);
    clmul128(&p1l, &p1h, _mm_xor_si128(a0, a1), _mm_xor_si128(b0, b1));
    p1l = _mm_xor_si128(p1l, _mm_xor_si128(p0l, p2l));
    p1h = _mm_xor_si128(p1h, _mm_xor_si128(p0h, p2h));
    const __m128i r0 = p0l, r1 = _mm_xor_si128(p0h, p1l), r2 = _mm_xor_si128(p2l, p1h), r3 = p2h;
    _mm_storeu_si128((__m128i *)&c[0], r0);
    _mm_storeu_si128((__m128i *)&c[2], r1);
    _mm_storeu_si128((__m128i *)&c[4], r2);
    _mm_storeu_si128((__m128i *)&c[6], r3);
}
#endif

void gf_mul(gf_t *r, const gf_t *a, const gf_t *b) {
    uint64_t c[8];
#if defined(__PCLMUL__)
    if (!g_portable) mul_pclmul(c, a, b);
    else
#endif
        mul_portable(c, a, b);
    gf_reduce(r, c);
}

static uint64_t spread32(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x << 2)) & 0x3333333333333333ULL;
    x = (x | (x << 1)) & 0x5555555555555555ULL;
    return x;
}
#if defined(__x86_64__)
#include <immintrin.h>
/* squaring = bit spreading: one PDEP per 32 input bits (BMI2), chosen at run time */
static int g_bmi2 = -1;
__attribute__((target("bmi2"))) static void sqr_pdep(gf_t *r, const gf_t *a) {
    uint64_t c[8];
    for (int i = 0; i < 4; i++) {
        c[2 * i] = _pdep_u64(a->w[i] & 0xffffffffu, 0x5555555555555555ULL);
        c[2 * i + 1] = _pdep_u64(a->w[i] >> 32, 0x5555555555555555ULL);
    }
    gf_reduce(r, c);
}
#endif
void gf_sqr(gf_t *r, const gf_t *a) {
    uint64_t c[8];
#if defined(__x86_64__)
    if (g_bmi2 < 0) g_bmi2 = __builtin_cpu_supports("bmi2") ? 1 : 0;
    if (g_bmi2 && !g_portable) {
        sqr_pdep(r, a);
        return;
    }
#endif
#if defined(__PCLMUL__)
    if (!g_portable) {
        __m128i a01 = _mm_loadu_si128((const __m128i *)&a->w[0]), a23 = _mm_loadu_si128((const __m128i *)&a->w[2]);
        __m128i k0 = _mm_clmulepi64_si128(a01, a01, 0x00), k2 = _mm_clmulepi64_si128(a01, a01, 0x11);
        __m128i k4 = _mm_clmulepi64_si128(a23, a23, 0x00), k6 = _mm_clmulepi64_si128(a23, a23, 0x11);
        c[0] = (uint64_t)_mm_cvtsi128_si64(k0); c[1] = (uint64_t)_mm_extract_epi64(k0, 1);
        c[2] = (uint64_t)_mm_cvtsi128_si64(k2); c[3] = (uint64_t)_mm_extract_epi64(k2, 1);
        c[4] = (uint64_t)_mm_cvtsi128_si64(k4); c[5] = (uint64_t)_mm_extract_epi64(k4, 1);
        c[6] = (uint64_t)_mm_cvtsi128_si64(k6); c[7] = (uint64_t)_mm_extract_epi64(k6, 1);
        gf_reduce(r, c);
        return;
    }
#endif
    for (int i = 0; i < 4; i++) {
        c[2 * i] = spread32((uint32_t)a->w[i]);
        c[2 * i + 1] = spread32((uint32_t)(a->w[i] >> 32));
    }
    gf_reduce(r, c);
}
void gf_sqr_n(gf_t *r, const gf_t *a, int n) {
    gf_t t = *a;
    for (int i = 0; i < n; i++) gf_sqr(&t, &t);
    *r = t;
}

/* Itoh-Tsujii: a^(2^233-2) = (a^(2^232-1))^2, chain 1,2,3,6,7,14,28,29,58,116,232 */
void gf_inv(gf_t *r, const gf_t *a) {
    gf_t b1 = *a, b2, b3, b6, b7, b14, b28, b29, b58, b116, b232, t;
    gf_sqr(&t, &b1);          gf_mul(&b2, &t, &b1);
    gf_sqr(&t, &b2);          gf_mul(&b3, &t, &b1);
    gf_sqr_n(&t, &b3, 3);     gf_mul(&b6, &t, &b3);
    gf_sqr(&t, &b6);          gf_mul(&b7, &t, &b1);
    gf_sqr_n(&t, &b7, 7);     gf_mul(&b14, &t, &b7);
    gf_sqr_n(&t, &b14, 14);   gf_mul(&b28, &t, &b14);
    gf_sqr(&t, &b28);         gf_mul(&b29, &t, &b1);
    gf_sqr_n(&t, &b29, 29);   gf_mul(&b58, &t, &b29);
    gf_sqr_n(&t, &b58, 58);   gf_mul(&b116, &t, &b58);
    gf_sqr_n(&t, &b116, 116); gf_mul(&b232, &t, &b116);
    gf_sqr(r, &b232);
}

void gf_sqrt(gf_t *r, const gf_t *a) { gf_sqr_n(r, a, 232); }

int gf_trace(const gf_t *a) {
    gf_t acc = *a, t = *a;
    for (int i = 1; i < 233; i++) {
        gf_sqr(&t, &t);
        gf_add(&acc, &acc, &t);
    }
    /* the trace lies in GF(2) */
    return (int)(acc.w[0] & 1);
}

void gf_halftrace(gf_t *r, const gf_t *a) {
    gf_t z = *a;
    for (int i = 1; i <= 116; i++) {
        gf_sqr_n(&z, &z, 2);
        gf_add(&z, &z, a);
    }
    *r = z;
}

void gf_to_le30(uint8_t out[30], const gf_t *a) {
    for (int i = 0; i < 30; i++) out[i] = (uint8_t)(a->w[i >> 3] >> (8 * (i & 7)));
}
int gf_from_le30(gf_t *r, const uint8_t in[30]) {
    memset(r->w, 0, 32);
    for (int i = 0; i < 30; i++) r->w[i >> 3] |= (uint64_t)in[i] << (8 * (i & 7));
    return (r->w[3] >> 41) == 0;
}
