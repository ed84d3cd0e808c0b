/* ORACLE (test infrastructure only). See ecfft.h for provenance. */
#include "ecfft.h"
#include <assert.h>
#include <stdlib.h>
#include <string.h>

static void fr_from_dec(fr_t *r, const char *s) {
    fr_t ten, acc = FR_ZERO, d;
    fr_from_u64(&ten, 10);
    for (; *s; s++) {
        fr_mul(&acc, &acc, &ten);
        fr_from_u64(&d, (uint64_t)(*s - '0'));
        fr_add(&acc, &acc, &d);
    }
    *r = acc;
}

/* short Weierstrass y^2 = x^3 + A x + B over Fr, affine, never the point at infinity here */
typedef struct { fr_t x, y; } swpt;

static void sw_dbl(swpt *r, const swpt *p, const fr_t *A) {
    fr_t l, t, x3, y3, three;
    fr_from_u64(&three, 3);
    fr_sqr(&t, &p->x);
    fr_mul(&t, &t, &three);
    fr_add(&t, &t, A);
    fr_add(&l, &p->y, &p->y);
    fr_inv(&l, &l);
    fr_mul(&l, &l, &t);
    fr_sqr(&x3, &l);
    fr_sub(&x3, &x3, &p->x);
    fr_sub(&x3, &x3, &p->x);
    fr_sub(&t, &p->x, &x3);
    fr_mul(&y3, &l, &t);
    fr_sub(&y3, &y3, &p->y);
    r->x = x3;
    r->y = y3;
}
/* p != +-q */
static void sw_add_with_inv(swpt *r, const swpt *p, const swpt *q, const fr_t *dxinv) {
    fr_t l, t, x3, y3;
    fr_sub(&t, &q->y, &p->y);
    fr_mul(&l, &t, dxinv);
    fr_sqr(&x3, &l);
    fr_sub(&x3, &x3, &p->x);
    fr_sub(&x3, &x3, &q->x);
    fr_sub(&t, &p->x, &x3);
    fr_mul(&y3, &l, &t);
    fr_sub(&y3, &y3, &p->y);
    r->x = x3;
    r->y = y3;
}
static void sw_add(swpt *r, const swpt *p, const swpt *q) {
    fr_t dx;
    fr_sub(&dx, &q->x, &p->x);
    assert(!fr_is_zero(&dx));
    fr_inv(&dx, &dx);
    sw_add_with_inv(r, p, q, &dx);
}

/* v^(2^e) */
static void fr_pow2k(fr_t *r, const fr_t *v, int e) {
    fr_t t = *v;
    for (int i = 0; i < e; i++) fr_sqr(&t, &t);
    *r = t;
}

static ecfft_domain *domain_new(int log_n2, int with_matrices);
ecfft_domain *ecfft_domain_new(int log_n2) { return domain_new(log_n2, 1); }
ecfft_domain *ecfft_domain_new_light(int log_n2) { return domain_new(log_n2, 0); }

static ecfft_domain *domain_new(int log_n2, int with_matrices) {
    assert(log_n2 >= 1 && log_n2 <= 28);
    ecfft_domain *d = (ecfft_domain *)calloc(1, sizeof(*d));
    const size_t N = (size_t)1 << log_n2;
    d->log_n2 = log_n2;
    d->n2 = N;
    d->levels = log_n2 - 1;
    /* constants: /root/reference/src/ec_fft.rs:209-229 */
    fr_t A, B;
    swpt G, C;
    fr_from_dec(&A, "2125753088427212854352924174339172498722499297750753614229533284661082");
    fr_from_dec(&B, "3303427382072851929105738691313541325219445842218525662544269869787589");
    fr_from_dec(&G.x, "1969398527398874941115360315313056361667745675958024267654083765592400");
    fr_from_dec(&G.y, "917696706299601920847965073366118878832337776859300472447868491055982");
    fr_from_dec(&C.x, "1557215852494830750811239888869886110709986867282698163663807961412586");
    fr_from_dec(&C.y, "2302954593454110051167704558708330032236229062988890422530712548754008");
    (void)B;
    /* generator of the order-N subgroup: 28 - log_n2 doublings (ec_fft.rs:121-124) */
    swpt g = G;
    for (int i = 0; i < 28 - log_n2; i++) sw_dbl(&g, &g, &A);

    /* ---- top-layer leaves: leaf_i = x(C + i g), natural order (ec_fft.rs:157-162) ---- */
    fr_t **layer = (fr_t **)calloc((size_t)log_n2 + 1, sizeof(fr_t *));
    layer[0] = (fr_t *)malloc(N * sizeof(fr_t));
    {
        const size_t BL = N < 1024 ? N : 1024; /* block: base_k + T_j with one batched inversion */
        swpt *T = (swpt *)malloc(BL * sizeof(swpt)); /* T[j] = j g for j >= 1 */
        if (BL > 1) T[1] = g;
        if (BL > 2) sw_dbl(&T[2], &g, &A);
        for (size_t j = 3; j < BL; j++) sw_add(&T[j], &T[j - 1], &g);
        swpt step = g; /* BL * g, only needed (and only finite) when there is more than one block */
        if (N > BL) sw_add(&step, &T[BL - 1], &g);
        fr_t *den = (fr_t *)malloc(BL * sizeof(fr_t));
        swpt base = C;
        for (size_t k = 0; k < N; k += BL) {
            layer[0][k] = base.x;
            for (size_t j = 1; j < BL; j++) {
                fr_sub(&den[j], &T[j].x, &base.x);
                assert(!fr_is_zero(&den[j]));
            }
            if (BL > 1) fr_batch_inv(den + 1, BL - 1);
            for (size_t j = 1; j < BL; j++) {
                swpt s;
                sw_add_with_inv(&s, &base, &T[j], &den[j]);
                layer[0][k + j] = s.x;
            }
            if (k + BL < N) sw_add(&base, &base, &step);
        }
        free(den);
        free(T);
    }
    d->leaves = layer[0];

    /* ---- isogeny chain and lower layers ---- */
    d->x0 = (fr_t *)malloc((size_t)log_n2 * sizeof(fr_t));
    d->t = (fr_t *)malloc((size_t)log_n2 * sizeof(fr_t));
    fr_t Ak = A;
    swpt gk = g;
    for (int k = 0; k < log_n2; k++) {
        const size_t Nk = N >> k; /* order of gk, size of layer k */
        swpt K = gk;
        for (size_t o = Nk; o > 2; o >>= 1) sw_dbl(&K, &K, &Ak);
        assert(fr_is_zero(&K.y)); /* order-2 point */
        fr_t x0 = K.x, t, three, five;
        fr_from_u64(&three, 3);
        fr_from_u64(&five, 5);
        fr_sqr(&t, &x0);
        fr_mul(&t, &t, &three);
        fr_add(&t, &t, &Ak); /* t = 3 x0^2 + A */
        d->x0[k] = x0;
        d->t[k] = t;
        /* next layer: psi(x) = x + t/(x - x0) on the first half of this layer */
        const size_t half = Nk >> 1;
        layer[k + 1] = (fr_t *)malloc(half * sizeof(fr_t));
        fr_t *den = (fr_t *)malloc(half * sizeof(fr_t));
        for (size_t i = 0; i < half; i++) fr_sub(&den[i], &layer[k][i], &x0);
        fr_batch_inv(den, half);
#pragma omp parallel for schedule(static) if (half >= 4096)
        for (size_t i = 0; i < half; i++) {
            fr_t q;
            fr_mul(&q, &t, &den[i]);
            fr_add(&layer[k + 1][i], &layer[k][i], &q);
        }
        free(den);
        /* image of the generator: X = x + t/(x-x0), Y = y (1 - t/(x-x0)^2); A' = A - 5t */
        if (Nk > 2) {
            fr_t dx, q, q2;
            fr_sub(&dx, &gk.x, &x0);
            fr_inv(&dx, &dx);
            fr_mul(&q, &t, &dx);
            fr_mul(&q2, &q, &dx);
            swpt ng;
            fr_add(&ng.x, &gk.x, &q);
            fr_mul(&q2, &q2, &gk.y);
            fr_sub(&ng.y, &gk.y, &q2);
            gk = ng;
            fr_mul(&q, &five, &t);
            fr_sub(&Ak, &Ak, &q);
        }
    }
    d->last = (fr_t *)malloc(2 * sizeof(fr_t));
    /* layer log_n2 has one leaf; the D / D' chains end one layer earlier with two leaves */
    d->last[0] = layer[log_n2 - 1][0];
    d->last[1] = layer[log_n2 - 1][1];

    /* ---- extend matrices, level k: sub-problem size m = n >> k on the even leaves of layer k ---- */
    const size_t n = N >> 1;
    d->dec = (fr_t **)calloc((size_t)d->levels + 1, sizeof(fr_t *));
    d->rec = (fr_t **)calloc((size_t)d->levels + 1, sizeof(fr_t *));
    for (int k = 0; with_matrices && k < d->levels; k++) {
        const size_t m = n >> k, h = m >> 1;
        const fr_t *L = layer[k];
        const fr_t x0 = d->x0[k];
        int e = 0; /* h = 2^e */
        while (((size_t)1 << e) < h) e++;
        d->dec[k] = (fr_t *)malloc(h * 4 * sizeof(fr_t));
        d->rec[k] = (fr_t *)malloc(h * 4 * sizeof(fr_t));
        fr_t *det = (fr_t *)malloc(h * sizeof(fr_t));
        fr_t *sv = (fr_t *)malloc(h * 2 * sizeof(fr_t));
#pragma omp parallel for schedule(static) if (h >= 1024)
        for (size_t j = 0; j < h; j++) {
            /* source pair (even leaves 2j, 2j+m), target pair (odd leaves 2j+1, 2j+1+m) */
            const fr_t s[2] = {L[2 * j], L[2 * j + m]}, tg[2] = {L[2 * j + 1], L[2 * j + 1 + m]};
            fr_t vs[2], vt[2];
            for (int q = 0; q < 2; q++) {
                /* v(x)^(h-1) = prod_{i<e} v^(2^i) */
                fr_t b, acc = FR_ONE;
                fr_sub(&b, &s[q], &x0);
                for (int i = 0; i < e; i++) { fr_mul(&acc, &acc, &b); fr_sqr(&b, &b); }
                vs[q] = acc;
                acc = FR_ONE;
                fr_sub(&b, &tg[q], &x0);
                for (int i = 0; i < e; i++) { fr_mul(&acc, &acc, &b); fr_sqr(&b, &b); }
                vt[q] = acc;
            }
            fr_t *R = &d->rec[k][4 * j];
            R[0] = vt[0]; fr_mul(&R[1], &tg[0], &vt[0]);
            R[2] = vt[1]; fr_mul(&R[3], &tg[1], &vt[1]);
            /* decompose = inverse of [[v0, s0 v0],[v1, s1 v1]] = 1/det [[s1 v1, -s0 v0],[-v1, v0]] */
            fr_t ds;
            fr_sub(&ds, &s[1], &s[0]);
            fr_mul(&det[j], &vs[0], &vs[1]);
            fr_mul(&det[j], &det[j], &ds);
            sv[2 * j] = vs[0];
            sv[2 * j + 1] = vs[1];
        }
        fr_batch_inv(det, h);
#pragma omp parallel for schedule(static) if (h >= 1024)
        for (size_t j = 0; j < h; j++) {
            const fr_t s0 = L[2 * j], s1 = L[2 * j + m];
            fr_t *M = &d->dec[k][4 * j], tmp;
            fr_mul(&tmp, &s1, &sv[2 * j + 1]); fr_mul(&M[0], &tmp, &det[j]);
            fr_mul(&tmp, &s0, &sv[2 * j]);     fr_mul(&tmp, &tmp, &det[j]); fr_neg(&M[1], &tmp);
            fr_mul(&tmp, &sv[2 * j + 1], &det[j]); fr_neg(&M[2], &tmp);
            fr_mul(&M[3], &sv[2 * j], &det[j]);
        }
        free(det);
        free(sv);
    }
    /* keep the lower layers for the chain-rule helpers */
    d->dec[d->levels] = NULL;
    d->rec[d->levels] = (fr_t *)layer; /* stash the layer table (freed in ecfft_domain_free) */
    return d;
}

static fr_t **layers_of(const ecfft_domain *d) { return (fr_t **)d->rec[d->levels]; }

void ecfft_domain_free(ecfft_domain *d) {
    if (!d) return;
    fr_t **layer = layers_of(d);
    for (int k = 0; k <= d->log_n2; k++) free(layer[k]);
    free(layer);
    for (int k = 0; k < d->levels; k++) { free(d->dec[k]); free(d->rec[k]); } /* NULL in a light domain */
    free(d->dec); free(d->rec); free(d->x0); free(d->t); free(d->last);
    free(d);
}

static void extend_rec(const ecfft_domain *d, int k, size_t m, const fr_t *in, fr_t *out) {
    if (m == 1) { out[0] = in[0]; return; }
    const size_t h = m >> 1;
    fr_t *p = (fr_t *)malloc(m * sizeof(fr_t)), *q = (fr_t *)malloc(m * sizeof(fr_t));
    for (size_t j = 0; j < h; j++) {
        const fr_t *M = &d->dec[k][4 * j];
        fr_t a, b;
        fr_mul(&a, &M[0], &in[j]); fr_mul(&b, &M[1], &in[j + h]); fr_add(&p[j], &a, &b);
        fr_mul(&a, &M[2], &in[j]); fr_mul(&b, &M[3], &in[j + h]); fr_add(&p[j + h], &a, &b);
    }
    extend_rec(d, k + 1, h, p, q);
    extend_rec(d, k + 1, h, p + h, q + h);
    for (size_t j = 0; j < h; j++) {
        const fr_t *M = &d->rec[k][4 * j];
        fr_t a, b;
        fr_mul(&a, &M[0], &q[j]); fr_mul(&b, &M[1], &q[j + h]); fr_add(&out[j], &a, &b);
        fr_mul(&a, &M[2], &q[j]); fr_mul(&b, &M[3], &q[j + h]); fr_add(&out[j + h], &a, &b);
    }
    free(p);
    free(q);
}
void ecfft_extend(const ecfft_domain *d, const fr_t *in, fr_t *out) { extend_rec(d, 0, d->n2 >> 1, in, out); }

void ecfft_vanish_at(const ecfft_domain *d, int shift, const fr_t *x, fr_t *out) {
    const size_t n = d->n2 >> 1;
    fr_t acc = FR_ONE, xk = *x;
    int e = 0;
    while (((size_t)1 << e) < n) e++;
    for (int k = 0; k < d->levels; k++) {
        /* |S^k| = n >> k; factor v(x)^(|S^k|/2) */
        fr_t v, pw, q;
        fr_sub(&v, &xk, &d->x0[k]);
        fr_pow2k(&pw, &v, e - k - 1);
        fr_mul(&acc, &acc, &pw);
        fr_inv(&q, &v);
        fr_mul(&q, &q, &d->t[k]);
        fr_add(&xk, &xk, &q);
    }
    fr_t lastf;
    fr_sub(&lastf, &xk, &d->last[shift]);
    fr_mul(out, &acc, &lastf);
}

void ecfft_vanish_derivative_on_roots(const ecfft_domain *d, int shift, fr_t *out) {
    const size_t n = d->n2 >> 1;
    fr_t **layer = layers_of(d);
    int e = 0;
    while (((size_t)1 << e) < n) e++;
    fr_t *cur = (fr_t *)malloc(n * sizeof(fr_t)), *nxt = (fr_t *)malloc(n * sizeof(fr_t));
    cur[0] = FR_ONE; /* |S| = 1: Z' = 1 */
    for (int k = d->levels - 1; k >= 0; k--) {
        const size_t m = n >> k, h = m >> 1;
        fr_t *den = (fr_t *)malloc(m * sizeof(fr_t));
        for (size_t j = 0; j < m; j++) fr_sub(&den[j], &layer[k][2 * j + shift], &d->x0[k]);
        fr_t *inv = (fr_t *)malloc(m * sizeof(fr_t));
        memcpy(inv, den, m * sizeof(fr_t));
        fr_batch_inv(inv, m);
#pragma omp parallel for schedule(static) if (m >= 4096)
        for (size_t j = 0; j < m; j++) {
            /* Z'_S(s) = v(s)^h psi'(s) Z'_psi(S)(psi(s)), psi' = 1 - t/(s-x0)^2 */
            fr_t pw, dpsi, q;
            fr_pow2k(&pw, &den[j], e - k - 1);
            fr_sqr(&q, &inv[j]);
            fr_mul(&q, &q, &d->t[k]);
            fr_sub(&dpsi, &FR_ONE, &q);
            fr_mul(&pw, &pw, &dpsi);
            fr_mul(&nxt[j], &pw, &cur[j % h]);
        }
        free(den);
        free(inv);
        fr_t *sw = cur; cur = nxt; nxt = sw;
    }
    memcpy(out, cur, n * sizeof(fr_t));
    free(cur);
    free(nxt);
}

void ecfft_vanish_on_other(const ecfft_domain *d, int shift, fr_t *out) {
    const size_t n = d->n2 >> 1;
    fr_t **layer = layers_of(d);
    int e = 0;
    while (((size_t)1 << e) < n) e++;
    fr_t *cur = (fr_t *)malloc(n * sizeof(fr_t)), *nxt = (fr_t *)malloc(n * sizeof(fr_t));
    fr_sub(&cur[0], &d->last[1 - shift], &d->last[shift]);
    for (int k = d->levels - 1; k >= 0; k--) {
        const size_t m = n >> k, h = m >> 1;
#pragma omp parallel for schedule(static) if (m >= 4096)
        for (size_t j = 0; j < m; j++) {
            fr_t v, pw;
            fr_sub(&v, &layer[k][2 * j + (1 - shift)], &d->x0[k]);
            fr_pow2k(&pw, &v, e - k - 1);
            fr_mul(&nxt[j], &pw, &cur[j % h]);
        }
        fr_t *sw = cur; cur = nxt; nxt = sw;
    }
    memcpy(out, cur, n * sizeof(fr_t));
    free(cur);
    free(nxt);
}

/* evaluate_poly_at_alpha_using_barycentric_weights (/root/reference/src/ec_fft.rs:455-491) for several points:
 * out[q] = Z_D(x_q) * sum_i evals[i] * w_i / (x_q - d_i), w_i = 1/Z'_D(d_i), D = the even leaves.  O(n) per point
 * (plus the weights once); works on a light domain.  No x_q may lie in D. */
void ecfft_bary_eval(const ecfft_domain *d, const fr_t *evals, const fr_t *xs, size_t nx, fr_t *out) {
    const size_t n = d->n2 >> 1;
    fr_t *w = (fr_t *)malloc(n * sizeof(fr_t));
    ecfft_vanish_derivative_on_roots(d, 0, w);
    fr_batch_inv(w, n);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t q = 0; q < nx; q++) {
        fr_t *den = (fr_t *)malloc(n * sizeof(fr_t));
        for (size_t i = 0; i < n; i++) fr_sub(&den[i], &xs[q], &d->leaves[2 * i]);
        fr_batch_inv(den, n);
        fr_t acc = FR_ZERO, t, z;
        for (size_t i = 0; i < n; i++) {
            fr_mul(&t, &w[i], &den[i]);
            fr_mul(&t, &t, &evals[i]);
            fr_add(&acc, &acc, &t);
        }
        ecfft_vanish_at(d, 0, &xs[q], &z);
        fr_mul(&out[q], &acc, &z);
        free(den);
    }
    free(w);
}
