/* ORACLE (test infrastructure only). See k233.h for provenance. */
#include "k233.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* NIST K-233 generator (FIPS 186-4 D.1.3.2 / SEC 2 sect233k1) */
static k233_pt g_gen;
static int g_gen_ready = 0;
static void hex_to_gf(gf_t *r, const char *hex) {
    /* big-endian hex string, up to 60 digits */
    memset(r->w, 0, 32);
    size_t n = strlen(hex);
    for (size_t i = 0; i < n; i++) {
        char c = hex[n - 1 - i];
        uint64_t v = (c >= '0' && c <= '9') ? (uint64_t)(c - '0') : (uint64_t)((c | 32) - 'a' + 10);
        r->w[i >> 4] |= v << (4 * (i & 15));
    }
}
static const k233_pt *gen(void) {
    if (!g_gen_ready) {
        hex_to_gf(&g_gen.x, "017232BA853A7E731AF129F22FF4149563A419C26BF50A4C9D6EEFAD6126");
        hex_to_gf(&g_gen.y, "01DB537DECE819B7F70F555A67C427A8CD9BF18AEB9B56E0C11056FAE6A3");
        g_gen.inf = 0;
        g_gen_ready = 1;
    }
    return &g_gen;
}
void k233_generator(k233_pt *r) { *r = *gen(); }

int k233_on_curve(const k233_pt *p) {
    if (p->inf) return 1;
    gf_t l, r, t;
    gf_sqr(&l, &p->y);
    gf_mul(&t, &p->x, &p->y);
    gf_add(&l, &l, &t); /* y^2 + xy */
    gf_sqr(&t, &p->x);
    gf_mul(&r, &t, &p->x);
    gf_add(&r, &r, &GF_ONE); /* x^3 + 1 */
    return gf_eq(&l, &r);
}
void k233_neg(k233_pt *r, const k233_pt *p) {
    *r = *p;
    if (!p->inf) gf_add(&r->y, &p->x, &p->y);
}
int k233_eq(const k233_pt *p, const k233_pt *q) {
    if (p->inf || q->inf) return p->inf && q->inf;
    return gf_eq(&p->x, &q->x) && gf_eq(&p->y, &q->y);
}
void k233_dbl(k233_pt *r, const k233_pt *p) {
    if (p->inf || gf_is_zero(&p->x)) { /* 2*(0,1) = infinity */
        r->inf = 1;
        r->x = GF_ZERO;
        r->y = GF_ZERO;
        return;
    }
    gf_t l, t, x3, y3;
    gf_inv(&t, &p->x);
    gf_mul(&l, &t, &p->y);
    gf_add(&l, &l, &p->x); /* lambda = x + y/x */
    gf_sqr(&x3, &l);
    gf_add(&x3, &x3, &l); /* a = 0 */
    gf_sqr(&y3, &p->x);
    gf_add(&t, &l, &GF_ONE);
    gf_mul(&t, &t, &x3);
    gf_add(&y3, &y3, &t); /* x^2 + (lambda+1) x3 */
    r->x = x3;
    r->y = y3;
    r->inf = 0;
}
void k233_add(k233_pt *r, const k233_pt *p, const k233_pt *q) {
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    if (gf_eq(&p->x, &q->x)) {
        if (gf_eq(&p->y, &q->y)) { k233_dbl(r, p); return; }
        r->inf = 1; r->x = GF_ZERO; r->y = GF_ZERO; /* q = -p */
        return;
    }
    gf_t l, t, x3, y3, dx;
    gf_add(&dx, &p->x, &q->x);
    gf_add(&t, &p->y, &q->y);
    gf_inv(&l, &dx);
    gf_mul(&l, &l, &t);
    gf_sqr(&x3, &l);
    gf_add(&x3, &x3, &l);
    gf_add(&x3, &x3, &dx);
    gf_add(&t, &p->x, &x3);
    gf_mul(&y3, &l, &t);
    gf_add(&y3, &y3, &x3);
    gf_add(&y3, &y3, &p->y);
    r->x = x3;
    r->y = y3;
    r->inf = 0;
}

/* ---- Lopez-Dahab projective (x = X/Z, y = Y/Z^2), HMV "Guide to ECC" Alg. 3.24/3.25, a=0 b=1 ---- */
typedef struct { gf_t X, Y, Z; } ld_pt; /* Z == 0: infinity */

static void ld_dbl(ld_pt *r, const ld_pt *p) {
    if (gf_is_zero(&p->Z)) { *r = *p; return; }
    gf_t z2, x2, z4, x4, y2, t;
    ld_pt o;
    gf_sqr(&z2, &p->Z);
    gf_sqr(&x2, &p->X);
    gf_mul(&o.Z, &z2, &x2);
    gf_sqr(&z4, &z2);
    gf_sqr(&x4, &x2);
    gf_add(&o.X, &x4, &z4);
    gf_sqr(&y2, &p->Y);
    gf_add(&y2, &y2, &z4);
    gf_mul(&t, &o.X, &y2);
    gf_mul(&o.Y, &z4, &o.Z);
    gf_add(&o.Y, &o.Y, &t);
    *r = o;
}
static void ld_from_affine(ld_pt *r, const k233_pt *p) {
    if (p->inf) { r->X = GF_ONE; r->Y = GF_ZERO; r->Z = GF_ZERO; return; }
    r->X = p->x; r->Y = p->y; r->Z = GF_ONE;
}
static void ld_add_mixed(ld_pt *r, const ld_pt *p, const k233_pt *q) {
    if (q->inf) { *r = *p; return; }
    if (gf_is_zero(&p->Z)) { ld_from_affine(r, q); return; }
    gf_t A, B, C, D, E, F, G, z2, t;
    ld_pt o;
    gf_sqr(&z2, &p->Z);
    gf_mul(&t, &q->y, &z2);
    gf_add(&A, &p->Y, &t);
    gf_mul(&t, &q->x, &p->Z);
    gf_add(&B, &p->X, &t);
    if (gf_is_zero(&B)) {
        if (gf_is_zero(&A)) { ld_pt qq; ld_from_affine(&qq, q); ld_dbl(r, &qq); return; }
        r->X = GF_ONE; r->Y = GF_ZERO; r->Z = GF_ZERO;
        return;
    }
    gf_mul(&C, &p->Z, &B);
    gf_sqr(&t, &B);
    gf_mul(&D, &t, &C); /* a = 0 */
    gf_sqr(&o.Z, &C);
    gf_mul(&E, &A, &C);
    gf_sqr(&t, &A);
    gf_add(&o.X, &t, &D);
    gf_add(&o.X, &o.X, &E);
    gf_mul(&t, &q->x, &o.Z);
    gf_add(&F, &o.X, &t);
    gf_add(&t, &q->x, &q->y);
    gf_sqr(&G, &o.Z);
    gf_mul(&G, &G, &t);
    gf_add(&t, &E, &o.Z);
    gf_mul(&o.Y, &t, &F);
    gf_add(&o.Y, &o.Y, &G);
    *r = o;
}
static void ld_to_affine(k233_pt *r, const ld_pt *p) {
    if (gf_is_zero(&p->Z)) { r->inf = 1; r->x = GF_ZERO; r->y = GF_ZERO; return; }
    gf_t zi, zi2;
    gf_inv(&zi, &p->Z);
    gf_sqr(&zi2, &zi);
    gf_mul(&r->x, &p->X, &zi);
    gf_mul(&r->y, &p->Y, &zi2);
    r->inf = 0;
}

/* width-4 NAF, left-to-right, LD accumulator */
void k233_mul_bytes(k233_pt *r, const k233_pt *p, const uint8_t *k, size_t klen) {
    uint64_t e[5] = {0, 0, 0, 0, 0};
    if (klen > 32) klen = 32;
    for (size_t i = 0; i < klen; i++) e[i >> 3] |= (uint64_t)k[i] << (8 * (i & 7));
    int8_t naf[264];
    int len = 0;
    while (e[0] | e[1] | e[2] | e[3] | e[4]) {
        int d = 0;
        if (e[0] & 1) {
            d = (int)(e[0] & 15);
            if (d >= 8) d -= 16;
            /* e -= d */
            if (d > 0) {
                e[0] -= (uint64_t)d; /* low 4 bits >= d: no borrow */
            } else {
                uint64_t add = (uint64_t)(-d);
                for (int i = 0; i < 5 && add; i++) {
                    uint64_t s = e[i] + add;
                    add = s < e[i];
                    e[i] = s;
                }
            }
        }
        naf[len++] = (int8_t)d;
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
        e[4] >>= 1;
    }
    if (p->inf || len == 0) { r->inf = 1; r->x = GF_ZERO; r->y = GF_ZERO; return; }
    k233_pt tab[4], p2, nq; /* 1P 3P 5P 7P */
    tab[0] = *p;
    k233_dbl(&p2, p);
    for (int i = 1; i < 4; i++) k233_add(&tab[i], &tab[i - 1], &p2);
    ld_pt acc;
    acc.X = GF_ONE; acc.Y = GF_ZERO; acc.Z = GF_ZERO;
    for (int i = len - 1; i >= 0; i--) {
        ld_dbl(&acc, &acc);
        int d = naf[i];
        if (d > 0) ld_add_mixed(&acc, &acc, &tab[d >> 1]);
        else if (d < 0) { k233_neg(&nq, &tab[(-d) >> 1]); ld_add_mixed(&acc, &acc, &nq); }
    }
    ld_to_affine(r, &acc);
}

void k233_mul_fr_wnaf(k233_pt *r, const k233_pt *p, const fr_t *k) {
    /* fr_to_le_bytes (curve.rs:162-182): canonical limbs, LE bytes, truncate to 30, strip trailing zeros */
    uint64_t c[4];
    uint8_t b[32];
    fr_to_canonical(c, k);
    for (int i = 0; i < 32; i++) b[i] = (uint8_t)(c[i >> 3] >> (8 * (i & 7)));
    size_t len = 30;
    while (len && b[len - 1] == 0) len--;
    k233_mul_bytes(r, p, b, len);
}

/*
 * tau-adic scalar multiplication on the Koblitz curve -- what the reference's xsk233_mul_frob (curve.rs:118, "frob")
 * does inside xs233: the Frobenius map tau(x, y) = (x^2, y^2) satisfies tau^2 + tau + 2 = 0 on K-233 (a = 0, mu = -1)
 * and acts on E[r] as multiplication by lambda (lambda^2 + lambda + 2 = 0 mod r), so k P = sum_i u_i tau^i(P) for a
 * tau-adic expansion of k: no doublings, only squarings.  Restated from J. Solinas, "Efficient arithmetic on Koblitz
 * curves" (2000): width-4 TNAF (digits +-1 +-3 +-5 +-7, density 1/5, representatives alpha_u of u mod tau^4 below).
 * The reduction of k to r0 + r1 tau with |r0|, |r1| < 2^118 is done GLV-style with a reduced basis (V1, V2) of the
 * lattice { (a, c) : a + c lambda = 0 mod r } instead of Solinas' rounding in Z[tau]; any such pair gives the same point.
 * Constants generated and checked with Python big integers (tests/test_oracle_k233.py cross-checks against the wNAF).
 */
typedef unsigned __int128 u128;
typedef __int128 i128;
static const uint64_t TAU_G1[5] = {0x751327740a80ea96ull, 0x34a059f450a5cbbdull, 0xe974eb7675acc618ull, 0x805b961da3b46581ull, 0x000000000000064aull}; /* floor(2^384 V2C / r) */
static const uint64_t TAU_G2[5] = {0xa8d31ce74a1c19ddull, 0x39da825e09899e5eull, 0x79966d7dcb1ecea9ull, 0xae5af5c6dc2d5428ull, 0x0000000000001105ull}; /* floor(2^384 V1C / r) */
#define U128(hi, lo) (((u128)(hi) << 64) | (u128)(lo))
/* V1 = (V1A, V1C), V2 = (V2A, V2C), two's complement mod 2^128; V1A V2C - V2A V1C = r */
#define TAU_V1A U128(0x000325402dcb0ed1ull, 0xda32c0f4ba75bb3bull)
#define TAU_V1C U128(0x000882d72d7ae36eull, 0x16aa143ccb36bee6ull)
#define TAU_V2A U128(0xfff21f91d2d547f5ull, 0xacde987b24083d6full)
#define TAU_V2C U128(0x000325402dcb0ed1ull, 0xda32c0f4ba75bb3bull)

/* floor(k g / 2^384) for k of 4 limbs and g of 5 limbs (the result fits 128 bits) */
static u128 mul_shift384(const uint64_t k[4], const uint64_t g[5]) {
    uint64_t prod[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 5; j++) {
            const u128 t = (u128)k[i] * g[j] + prod[i + j] + carry;
            prod[i + j] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        prod[i + 5] = carry;
    }
    return U128(prod[7], prod[6]);
}
/* width-4 TNAF of the canonical scalar k (4 limbs): digits in {0, +-1, +-3, +-5, +-7}, returns the length (<= 240) */
static int tau_recode(int8_t *dig, const uint64_t k[4]) {
    const u128 q1 = mul_shift384(k, TAU_G1), q2 = mul_shift384(k, TAU_G2);
    /* (k, 0) - q1 V1 + q2 V2, arithmetic mod 2^128: the true values are below 2^118 in absolute value */
    i128 r0 = (i128)(U128(k[1], k[0]) - q1 * TAU_V1A + q2 * TAU_V2A);
    i128 r1 = (i128)(q2 * TAU_V2C - q1 * TAU_V1C);
    /* alpha_u = beta + gamma tau for u = 1, 3, 5, 7:  1,  -3 - tau,  -1 - tau,  1 - tau;  t_4 = 10 (tau mod tau^4) */
    static const int beta[4] = {1, -3, -1, 1}, gamma[4] = {0, -1, -1, -1};
    int len = 0;
    while ((r0 != 0 || r1 != 0) && len < 250) {
        int u = 0;
        if (r0 & 1) {
            u = (int)((uint64_t)(r0 + r1 * 10) & 15);
            if (u >= 8) u -= 16;
            const int a = u > 0 ? u : -u, sg = u > 0 ? 1 : -1;
            r0 -= sg * beta[a >> 1];
            r1 -= sg * gamma[a >> 1];
        }
        dig[len++] = (int8_t)u;
        /* (r0 + r1 tau) / tau = (r1 - r0/2) - (r0/2) tau */
        const i128 half = r0 / 2; /* r0 is even here */
        r0 = r1 - half;
        r1 = -half;
    }
    return len;
}
static void ld_frobenius(ld_pt *p) {
    gf_sqr(&p->X, &p->X);
    gf_sqr(&p->Y, &p->Y);
    gf_sqr(&p->Z, &p->Z);
}
/* r = a + b for affine a, b with x_a != x_b, given inv = 1/(x_a + x_b) */
static void aff_add_with_inv(k233_pt *r, const k233_pt *a, const gf_t *bx, const gf_t *by, const gf_t *inv) {
    gf_t l, t, dx, x3, y3;
    gf_add(&dx, &a->x, bx);
    gf_add(&t, &a->y, by);
    gf_mul(&l, &t, inv);
    gf_sqr(&x3, &l);
    gf_add(&x3, &x3, &l);
    gf_add(&x3, &x3, &dx);
    gf_add(&t, &a->x, &x3);
    gf_mul(&y3, &l, &t);
    gf_add(&y3, &y3, &x3);
    gf_add(&y3, &y3, &a->y);
    r->x = x3;
    r->y = y3;
    r->inf = 0;
}
static void k233_mul_tau_ld(ld_pt *acc, const k233_pt *p, const fr_t *k) {
    acc->X = GF_ONE; acc->Y = GF_ZERO; acc->Z = GF_ZERO;
    uint64_t c[4];
    fr_to_canonical(c, k);
    if (p->inf || !(c[0] | c[1] | c[2] | c[3])) return;
    int8_t dig[256];
    const int len = tau_recode(dig, c);
    /* table: P, alpha_3 P = tau^2 P - P, alpha_5 P = -(P + tau P), alpha_7 P = P - tau P; the two denominators
       x_P + x_tauP and x_P + x_tau2P are inverted together.  They vanish only for points outside E[r] \ {inf}. */
    k233_pt tab[4], T1, T2, nT, nP;
    T1.inf = T2.inf = 0;
    gf_sqr(&T1.x, &p->x); gf_sqr(&T1.y, &p->y);
    gf_sqr(&T2.x, &T1.x); gf_sqr(&T2.y, &T1.y);
    gf_t d1, d2, d12, i12, i1, i2;
    gf_add(&d1, &p->x, &T1.x);
    gf_add(&d2, &p->x, &T2.x);
    if (gf_is_zero(&d1) || gf_is_zero(&d2)) { /* not a point of E[r]: plain double-and-add */
        k233_pt t;
        k233_mul_fr_wnaf(&t, p, k);
        ld_from_affine(acc, &t);
        return;
    }
    gf_mul(&d12, &d1, &d2);
    gf_inv(&i12, &d12);
    gf_mul(&i1, &i12, &d2);
    gf_mul(&i2, &i12, &d1);
    tab[0] = *p;
    k233_neg(&nP, p);
    aff_add_with_inv(&tab[1], &T2, &nP.x, &nP.y, &i2);  /* tau^2 P - P */
    aff_add_with_inv(&tab[2], p, &T1.x, &T1.y, &i1);    /* P + tau P, negated below */
    k233_neg(&tab[2], &tab[2]);
    k233_neg(&nT, &T1);
    aff_add_with_inv(&tab[3], p, &nT.x, &nT.y, &i1);    /* P - tau P */
    for (int i = len - 1; i >= 0; i--) {
        ld_frobenius(acc);
        const int d = dig[i];
        if (d > 0) ld_add_mixed(acc, acc, &tab[d >> 1]);
        else if (d < 0) { k233_pt nq; k233_neg(&nq, &tab[(-d) >> 1]); ld_add_mixed(acc, acc, &nq); }
    }
}
/* general Lopez-Dahab addition (Lange-Doche 2005 form, 13M + 4S), with the exceptional cases handled in front */
static void ld_add(ld_pt *r, const ld_pt *p, const ld_pt *q) {
    if (gf_is_zero(&p->Z)) { *r = *q; return; }
    if (gf_is_zero(&q->Z)) { *r = *p; return; }
    gf_t A, B, C, D, E, F, G, H, I, J, t, u;
    ld_pt o;
    gf_mul(&A, &p->X, &q->Z);
    gf_mul(&B, &q->X, &p->Z);
    gf_sqr(&t, &q->Z);
    gf_mul(&G, &p->Y, &t);
    gf_sqr(&t, &p->Z);
    gf_mul(&H, &q->Y, &t);
    gf_add(&E, &A, &B);
    gf_add(&I, &G, &H);
    if (gf_is_zero(&E)) { /* same x: P = Q or P = -Q */
        if (gf_is_zero(&I)) { ld_dbl(r, p); return; }
        r->X = GF_ONE; r->Y = GF_ZERO; r->Z = GF_ZERO;
        return;
    }
    gf_sqr(&C, &A);
    gf_sqr(&D, &B);
    gf_add(&F, &C, &D);
    gf_mul(&J, &I, &E);
    gf_mul(&t, &p->Z, &q->Z);
    gf_mul(&o.Z, &F, &t);
    gf_add(&t, &H, &D);
    gf_mul(&t, &A, &t);
    gf_add(&u, &C, &G);
    gf_mul(&u, &B, &u);
    gf_add(&o.X, &t, &u);
    gf_mul(&t, &A, &J);
    gf_mul(&u, &F, &G);
    gf_add(&t, &t, &u);
    gf_mul(&t, &t, &F);
    gf_add(&u, &J, &o.Z);
    gf_mul(&u, &u, &o.X);
    gf_add(&o.Y, &t, &u);
    *r = o;
}
void k233_mul_fr_tau(k233_pt *r, const k233_pt *p, const fr_t *k) {
    ld_pt acc;
    k233_mul_tau_ld(&acc, p, k);
    ld_to_affine(r, &acc);
}
/* point_scalar_mul (curve.rs:113-126): the reference calls xsk233_mul_frob, i.e. the tau-adic ladder */
void k233_mul_fr(k233_pt *r, const k233_pt *p, const fr_t *k) { k233_mul_fr_tau(r, p, k); }

void xsk233_encode(uint8_t out[30], const k233_pt *p) {
    if (p->inf) { memset(out, 0, 30); return; }
    /* Q = P + N has w(Q) = w(P) + 1 = (y + 1 + x)/x */
    gf_t t, xi, w;
    gf_add(&t, &p->y, &p->x);
    gf_add(&t, &t, &GF_ONE);
    gf_inv(&xi, &p->x);
    gf_mul(&w, &t, &xi);
    gf_to_le30(out, &w);
}

int xsk233_decode(k233_pt *p, const uint8_t in[30]) {
    gf_t w, d, e, f, x1, x2, y1, lam, u2, t;
    p->inf = 1; p->x = GF_ZERO; p->y = GF_ZERO;
    if (!gf_from_le30(&w, in)) return 0;
    if (gf_is_zero(&w)) return 1; /* neutral */
    gf_sqr(&d, &w);
    gf_add(&d, &d, &w); /* d = w^2 + w + a, a = 0 */
    if (gf_is_zero(&d)) return 0; /* w = 1: x = 1, an order-4 point */
    gf_inv(&t, &d);
    gf_sqr(&e, &t); /* b/d^2 */
    if (gf_trace(&e)) return 0;
    gf_halftrace(&f, &e);
    gf_mul(&x1, &d, &f);   /* roots of x^2 + d x + 1 */
    gf_add(&x2, &x1, &d);
    if (gf_trace(&x1)) return 0; /* not in 2E: outside E[r] u (E[r]+N) */
    /* T1 = (x1, x1 w + 1) has w-coordinate w.  T1 is in E[r]+N iff its halves are not doubles. */
    gf_mul(&y1, &x1, &w);
    gf_add(&y1, &y1, &GF_ONE);
    gf_halftrace(&lam, &x1);            /* lam^2 + lam = x1 (+a) */
    gf_add(&t, &lam, &GF_ONE);
    gf_mul(&t, &t, &x1);
    gf_add(&u2, &y1, &t);               /* u^2 = y + (lam+1) x, u = x(half) */
    gf_t xp;
    if (gf_trace(&u2)) xp = x2;         /* T1 = Q in the coset: P = Q + N has x = 1/x1 = x2 */
    else xp = x1;                       /* T1 in E[r]: Q = -T1 + N, P = -T1, x = x1 */
    /* P has w-coordinate w + 1: y = x (w+1) + 1 */
    gf_add(&t, &w, &GF_ONE);
    gf_mul(&p->y, &xp, &t);
    gf_add(&p->y, &p->y, &GF_ONE);
    p->x = xp;
    p->inf = 0;
    return 1;
}

void k233_msm(k233_pt *r, const fr_t *scalars, const k233_pt *points, size_t n, int nthreads) {
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    (void)nthreads;
    k233_pt *part = (k233_pt *)malloc((size_t)nt * sizeof(k233_pt));
    for (int t = 0; t < nt; t++) { part[t].inf = 1; part[t].x = GF_ZERO; part[t].y = GF_ZERO; }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        ld_pt acc;
        acc.X = GF_ONE; acc.Y = GF_ZERO; acc.Z = GF_ZERO;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (long i = 0; i < (long)n; i++) {
            /* the products stay projective (xs233 keeps extended coordinates too): no inversion per point */
            ld_pt s;
            k233_mul_tau_ld(&s, &points[i], &scalars[i]);
            ld_add(&acc, &acc, &s);
        }
        ld_to_affine(&part[tid], &acc);
    }
    k233_pt acc = part[0];
    for (int t = 1; t < nt; t++) k233_add(&acc, &acc, &part[t]);
    *r = acc;
    free(part);
}

/* ---- bulk helpers for fixtures and benchmarks (OpenMP) ---- */
/* out[i] = P0 + i * Q, i < n  (cheap synthetic SRS: one affine addition per point) */
void k233_chain_points(k233_pt *out, size_t n, const k233_pt *p0, const k233_pt *q) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    size_t chunk = (n + (size_t)nt - 1) / (size_t)nt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1)
#endif
    for (int t = 0; t < nt; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) continue;
        uint8_t kb[8];
        for (int i = 0; i < 8; i++) kb[i] = (uint8_t)((uint64_t)lo >> (8 * i));
        k233_pt cur;
        k233_mul_bytes(&cur, q, kb, 8);
        k233_add(&cur, &cur, p0);
        for (size_t i = lo; i < hi; i++) {
            out[i] = cur;
            k233_add(&cur, &cur, q);
        }
    }
}
/* out[i] = k_i * P (k_i Montgomery Fr): point_scalar_mul(_gen) in bulk (curve.rs:113-137, srs.rs:126-160) */
void k233_mul_batch(k233_pt *out, const k233_pt *p, const fr_t *k, size_t n) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (long i = 0; i < (long)n; i++) k233_mul_fr(&out[i], p, &k[i]);
}
void xsk233_encode_batch(uint8_t *out30, const k233_pt *pts, size_t n) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long i = 0; i < (long)n; i++) xsk233_encode(out30 + 30 * (size_t)i, &pts[i]);
}
/* returns the index of the first invalid encoding, or -1 */
long xsk233_decode_batch(k233_pt *pts, const uint8_t *in30, size_t n) {
    long bad = -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long i = 0; i < (long)n; i++) {
        if (!xsk233_decode(&pts[i], in30 + 30 * (size_t)i)) {
#ifdef _OPENMP
#pragma omp critical
#endif
            if (bad < 0 || i < bad) bad = i;
        }
    }
    return bad;
}
