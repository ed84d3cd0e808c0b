/* ORACLE (test infrastructure only). See fr.h for provenance. */
#include "fr.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

const uint64_t FR_P[4] = {0x6efb1ad5f173abdfULL, 0x00069d5bb915bcd4ULL, 0x0ULL, 0x0000008000000000ULL};
static const uint64_t FR_NP0 = 0xa2918b898c382fe1ULL; /* -p^-1 mod 2^64 */
const fr_t FR_ZERO = {{0, 0, 0, 0}};
const fr_t FR_ONE = {{0xc318337e3373abdfULL, 0x489471e21037c69eULL, 0xfffffffffffff2c5ULL, 0x0000007fffffffffULL}};
static const fr_t FR_R2 = {{0x1710ac1009468bb6ULL, 0xf7e3eb91db9a5b86ULL, 0x93c813eeb5b58a0aULL, 0x00000059bebed802ULL}};

static int geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > FR_P[i]) return 1;
        if (a[i] < FR_P[i]) return 0;
    }
    return 1;
}
static void sub_p(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - FR_P[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

void fr_add(fr_t *r, const fr_t *a, const fr_t *b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* p < 2^232 so no carry out of 256 bits */
    if (geq_p(t)) sub_p(t);
    memcpy(r->l, t, 32);
}
void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) {
    u128 br = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + FR_P[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}
void fr_neg(fr_t *r, const fr_t *a) { fr_sub(r, &FR_ZERO, a); }

/* CIOS Montgomery product, R = 2^256 */
void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_NP0;
        c = (u128)m * FR_P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * FR_P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_p(t)) sub_p(t);
    memcpy(r->l, t, 32);
}
void fr_sqr(fr_t *r, const fr_t *a) { fr_mul(r, a, a); }

int fr_is_zero(const fr_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a->l, b->l, 32) == 0; }

void fr_from_canonical(fr_t *r, const uint64_t c[4]) {
    fr_t t;
    memcpy(t.l, c, 32);
    fr_mul(r, &t, &FR_R2);
}
void fr_to_canonical(uint64_t c[4], const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}}, t;
    fr_mul(&t, a, &one);
    memcpy(c, t.l, 32);
}
void fr_from_u64(fr_t *r, uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    fr_from_canonical(r, c);
}

void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e) {
    fr_t acc = FR_ONE, base = *a;
    while (e) {
        if (e & 1) fr_mul(&acc, &acc, &base);
        fr_sqr(&base, &base);
        e >>= 1;
    }
    *r = acc;
}

/* a^(p-2) */
void fr_inv(fr_t *r, const fr_t *a) {
    uint64_t e[4];
    memcpy(e, FR_P, 32);
    e[0] -= 2; /* low limb of p is > 2, no borrow */
    fr_t acc = FR_ONE, base = *a;
    for (int i = 0; i < 232; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(&acc, &acc, &base);
        fr_sqr(&base, &base);
    }
    *r = acc;
}

void fr_from_be32_mod(fr_t *r, const uint8_t b[32]) {
    /* value = hi * 2^128 + lo, both 128-bit < p; combine in the field */
    uint64_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; i++) {
        hi[1] = (hi[1] << 8) | b[i];
        hi[0] = (hi[0] << 8) | b[8 + i];
        lo[1] = (lo[1] << 8) | b[16 + i];
        lo[0] = (lo[0] << 8) | b[24 + i];
    }
    fr_t H, L, S;
    uint64_t two128[4] = {0, 0, 1, 0};
    fr_from_canonical(&H, hi);
    fr_from_canonical(&L, lo);
    fr_from_canonical(&S, two128);
    fr_mul(&H, &H, &S);
    fr_add(r, &H, &L);
}

void fr_to_le29(uint8_t out[29], const fr_t *a) {
    uint64_t c[4];
    fr_to_canonical(c, a);
    for (int i = 0; i < 29; i++) out[i] = (uint8_t)(c[i >> 3] >> (8 * (i & 7)));
}
int fr_from_le29(fr_t *r, const uint8_t in[29]) {
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 29; i++) c[i >> 3] |= (uint64_t)in[i] << (8 * (i & 7));
    if (geq_p(c)) return 0;
    fr_from_canonical(r, c);
    return 1;
}

void fr_batch_inv(fr_t *v, size_t n) {
    if (!n) return;
    fr_t *pre = (fr_t *)malloc(n * sizeof(fr_t));
    fr_t acc = FR_ONE;
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!fr_is_zero(&v[i])) fr_mul(&acc, &acc, &v[i]);
    }
    fr_inv(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (fr_is_zero(&v[i])) continue;
        fr_t t;
        fr_mul(&t, &acc, &pre[i]);
        fr_mul(&acc, &acc, &v[i]);
        v[i] = t;
    }
    free(pre);
}

/* s0 = sum k_i, s1 = sum i * k_i over Montgomery-form scalars: the closed form of an MSM over points A + i Q
 * (tests of the full-size configurations) */
void fr_sum_weighted(const fr_t *k, size_t n, fr_t *s0, fr_t *s1) {
    fr_t a0 = FR_ZERO, a1 = FR_ZERO;
#pragma omp parallel
    {
        fr_t l0 = FR_ZERO, l1 = FR_ZERO, idx, t;
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < n; i++) {
            fr_from_u64(&idx, (uint64_t)i);
            fr_mul(&t, &idx, &k[i]);
            fr_add(&l0, &l0, &k[i]);
            fr_add(&l1, &l1, &t);
        }
#pragma omp critical
        {
            fr_add(&a0, &a0, &l0);
            fr_add(&a1, &a1, &l1);
        }
    }
    *s0 = a0;
    *s1 = a1;
}
