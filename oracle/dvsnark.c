/* ORACLE (test infrastructure only). See dvsnark.h for provenance. */
#include "dvsnark.h"
#include <stdlib.h>
#include <string.h>
#include "blake3.h"

static void eval_terms(const r1cs_t *r, int which, size_t row, const fr_t *w, fr_t *out) {
    /* R1CSInstance::eval_row, gnark_r1cs.rs:273-280 */
    fr_t acc = FR_ZERO, t;
    for (uint32_t p = r->rowptr[which][row]; p < r->rowptr[which][row + 1]; p++) {
        fr_mul(&t, &r->coeffs[r->coeff[which][p]], &w[r->wire[which][p]]);
        fr_add(&acc, &acc, &t);
    }
    *out = acc;
}
/* evaluate_monomial_basis_poly, gnark_r1cs.rs:391-399 */
static void monomial_eval(const fr_t *pub, size_t k, const fr_t *x, fr_t *out) {
    fr_t pw = FR_ONE, acc = FR_ZERO, t;
    for (size_t j = 0; j < k; j++) {
        fr_mul(&t, &pub[j], &pw);
        fr_add(&acc, &acc, &t);
        fr_mul(&pw, &pw, x);
    }
    *out = acc;
}

long r1cs_eval(const r1cs_t *r, const ecfft_domain *dom, const fr_t *w, fr_t *a, fr_t *b, fr_t *c, fr_t *iv) {
    long bad = -1;
    for (size_t row = 0; row < r->n; row++) {
        fr_t cw = FR_ZERO, lhs;
        a[row] = FR_ZERO;
        b[row] = FR_ZERO;
        if (row < r->nrows) {
            eval_terms(r, 0, row, w, &a[row]);
            eval_terms(r, 1, row, w, &b[row]);
            eval_terms(r, 2, row, w, &cw);
        }
        monomial_eval(w + 1, r->k, &dom->leaves[2 * row], &iv[row]);
        fr_sub(&c[row], &cw, &iv[row]); /* C' = C - D */
        fr_mul(&lhs, &a[row], &b[row]);
        if (!fr_eq(&lhs, &cw) && bad < 0) bad = (long)row; /* a*b == c + i */
    }
    return bad;
}

void dv_transcript_alpha(const uint8_t commit_p[30], const fr_t *pub, size_t k, fr_t *alpha) {
    uint8_t h_srs[32], h_circ[32], h_wit[32], h_pub[32], buf[64], ct[32], rt[32], root[32];
    blake3_hash_small(buf, 0, h_srs);  /* hashing of the SRS is commented out upstream: empty buffer */
    blake3_hash_small(buf, 0, h_circ); /* likewise for the circuit */
    blake3_hash_small(commit_p, 30, h_wit);
    uint8_t *pb = (uint8_t *)malloc(29 * k + 1);
    for (size_t j = 0; j < k; j++) fr_to_le29(pb + 29 * j, &pub[j]);
    blake3_hash_small(pb, 29 * k, h_pub);
    free(pb);
    memcpy(buf, h_srs, 32); memcpy(buf + 32, h_circ, 32);
    blake3_hash_small(buf, 64, ct);
    memcpy(buf, h_wit, 32); memcpy(buf + 32, h_pub, 32);
    blake3_hash_small(buf, 64, rt);
    memcpy(buf, ct, 32); memcpy(buf + 32, rt, 32);
    blake3_hash_small(buf, 64, root);
    memset(root + 28, 0, 4);
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 32; i++) c[i >> 3] |= (uint64_t)root[i] << (8 * (i & 7));
    fr_from_canonical(alpha, c);
}

static fr_t *fr_vec(size_t n) { return (fr_t *)malloc((n ? n : 1) * sizeof(fr_t)); }

/* The discrete logs of the SRS (compute_srs_matrices, srs.rs:112-167, before the fixed-base multiplications):
 * sc_m[nwires], sc_q[n], sc_k[4n]; bar_wts / z_vals2inv (n each, may be NULL) are the prover precomputes. */
void dv_setup_scalars(const r1cs_t *r, const ecfft_domain *dom, const trapdoor_t *td, fr_t *sc_m, fr_t *sc_q, fr_t *sc_k,
                      fr_t *bar_wts, fr_t *z_vals2inv) {
    const size_t n = r->n;
    fr_t zt[2], delta2;
    ecfft_vanish_at(dom, 0, &td->tau, &zt[0]);
    ecfft_vanish_at(dom, 1, &td->tau, &zt[1]);
    fr_sqr(&delta2, &td->delta);
    fr_t *bw[2], *zo_inv[2], *lt[2];
    for (int sh = 0; sh < 2; sh++) {
        bw[sh] = fr_vec(n);
        ecfft_vanish_derivative_on_roots(dom, sh, bw[sh]);
        fr_batch_inv(bw[sh], n); /* barycentric weights 1/Z'_S(s_i), ec_fft.rs:284-335 */
        zo_inv[sh] = fr_vec(n);
        ecfft_vanish_on_other(dom, sh, zo_inv[sh]);
        fr_batch_inv(zo_inv[sh], n); /* prepare_z_inv, srs.rs:292-304 */
        lt[sh] = fr_vec(n);
        for (size_t i = 0; i < n; i++) fr_sub(&lt[sh][i], &td->tau, &dom->leaves[2 * i + sh]);
        fr_batch_inv(lt[sh], n);
        for (size_t i = 0; i < n; i++) { /* L_i(tau) = Z(tau) / ((tau - s_i) Z'(s_i)), ec_fft.rs:340-390 */
            fr_mul(&lt[sh][i], &lt[sh][i], &zt[sh]);
            fr_mul(&lt[sh][i], &lt[sh][i], &bw[sh][i]);
        }
    }
    if (bar_wts) memcpy(bar_wts, bw[0], n * sizeof(fr_t));
    if (z_vals2inv) memcpy(z_vals2inv, zo_inv[0], n * sizeof(fr_t)); /* 1 / Z_D(d'_i) */
    /* unified-domain basis, ec_fft.rs:424-450 */
    fr_t *ltl = fr_vec(2 * n);
    for (size_t i = 0; i < n; i++) {
        fr_mul(&ltl[2 * i], &lt[0][i], &zt[1]);
        fr_mul(&ltl[2 * i], &ltl[2 * i], &zo_inv[1][i]);
        fr_mul(&ltl[2 * i + 1], &lt[1][i], &zt[0]);
        fr_mul(&ltl[2 * i + 1], &ltl[2 * i + 1], &zo_inv[0][i]);
    }
    /* accumulate_m_values over the rows with the Vandermonde block appended, srs.rs:53-84 */
    fr_t *mv = fr_vec(r->nwires);
    for (size_t j = 0; j < r->nwires; j++) mv[j] = FR_ZERO;
    for (size_t i = 0; i < n; i++) {
        fr_t sc[3], t;
        sc[0] = lt[0][i];
        fr_mul(&sc[1], &lt[0][i], &td->delta);
        fr_mul(&sc[2], &lt[0][i], &delta2);
        if (i < r->nrows)
            for (int which = 0; which < 3; which++)
                for (uint32_t p = r->rowptr[which][i]; p < r->rowptr[which][i + 1]; p++) {
                    fr_mul(&t, &r->coeffs[r->coeff[which][p]], &sc[which]);
                    fr_add(&mv[r->wire[which][p]], &mv[r->wire[which][p]], &t);
                }
        fr_t pw = FR_ONE; /* O row gains (wire 1+j, -d_i^j), gnark_r1cs.rs:357-383 */
        for (size_t j = 0; j < r->k; j++) {
            fr_mul(&t, &pw, &sc[2]);
            fr_sub(&mv[1 + j], &mv[1 + j], &t);
            fr_mul(&pw, &pw, &dom->leaves[2 * i]);
        }
    }
    /* compute_srs_matrices, srs.rs:112-167 */
    for (size_t j = 0; j < r->nwires; j++) fr_mul(&sc_m[j], &mv[j], &td->epsilon);
    fr_t f;
    fr_mul(&f, &zt[0], &delta2);
    fr_mul(&f, &f, &td->epsilon);
    for (size_t i = 0; i < n; i++) fr_mul(&sc_q[i], &f, &lt[1][i]);
    for (size_t i = 0; i < n; i++) {
        sc_k[i] = lt[0][i];
        fr_mul(&sc_k[n + i], &lt[0][i], &td->delta);
    }
    for (size_t j = 0; j < 2 * n; j++) fr_mul(&sc_k[2 * n + j], &ltl[j], &delta2);
    free(mv); free(ltl);
    for (int sh = 0; sh < 2; sh++) { free(bw[sh]); free(zo_inv[sh]); free(lt[sh]); }
}

srs_t *dv_setup(const r1cs_t *r, const ecfft_domain *dom, const trapdoor_t *td) {
    const size_t n = r->n;
    srs_t *s = (srs_t *)calloc(1, sizeof(*s));
    s->bar_wts = fr_vec(n);
    s->z_vals2inv = fr_vec(n);
    fr_t *sc_m = fr_vec(r->nwires), *sc_q = fr_vec(n), *sc_k = fr_vec(4 * n);
    dv_setup_scalars(r, dom, td, sc_m, sc_q, sc_k, s->bar_wts, s->z_vals2inv);
    k233_pt G;
    k233_generator(&G);
    s->g_m = (k233_pt *)malloc(r->nwires * sizeof(k233_pt));
    k233_mul_batch(s->g_m, &G, sc_m, r->nwires);
    s->g_q = (k233_pt *)malloc(n * sizeof(k233_pt));
    k233_mul_batch(s->g_q, &G, sc_q, n);
    s->g_k = (k233_pt *)malloc(4 * n * sizeof(k233_pt));
    k233_mul_batch(s->g_k, &G, sc_k, 4 * n);
    free(sc_m); free(sc_q); free(sc_k);
    return s;
}
void dv_srs_free(srs_t *s) {
    if (!s) return;
    free(s->g_m); free(s->g_q); free(s->g_k); free(s->z_vals2inv); free(s->bar_wts);
    free(s);
}

long dv_prove(const r1cs_t *r, const ecfft_domain *dom, const srs_t *srs, const fr_t *w, uint8_t proof[118],
              fr_t *stages, int nthreads) {
    const size_t n = r->n;
    fr_t *buf = stages ? stages : fr_vec(13 * n);
    fr_t *a = buf, *b = buf + n, *c = buf + 2 * n, *iv = buf + 3 * n;
    fr_t *a2 = buf + 4 * n, *b2 = buf + 5 * n, *c2 = buf + 6 * n, *i2 = buf + 7 * n;
    fr_t *q = buf + 8 * n, *ka = buf + 9 * n, *kb = buf + 10 * n, *kr = buf + 11 * n;
    long ret = 0;
    long bad = r1cs_eval(r, dom, w, a, b, c, iv);
    if (bad >= 0) { ret = 1 + bad; goto done; }
    k233_pt msm_gm, msm_q, commit, kzg;
    k233_msm(&msm_gm, w, srs->g_m, r->nwires, nthreads);            /* proving.rs:462-463 */
    ecfft_extend(dom, a, a2);                                        /* proving.rs:410-422 */
    ecfft_extend(dom, b, b2);
    ecfft_extend(dom, c, c2);
    ecfft_extend(dom, iv, i2);
    fr_t *r2 = fr_vec(n);
    for (size_t i = 0; i < n; i++) {                                 /* proving.rs:492-509 */
        fr_t t;
        fr_mul(&r2[i], &a2[i], &b2[i]);
        fr_sub(&r2[i], &r2[i], &i2[i]);
        fr_sub(&t, &r2[i], &c2[i]);
        fr_mul(&q[i], &t, &srs->z_vals2inv[i]);
    }
    k233_msm(&msm_q, q, srs->g_q, n, nthreads);                      /* proving.rs:511-512 */
    k233_add(&commit, &msm_q, &msm_gm);
    xsk233_encode(proof, &commit);
    fr_t alpha;
    dv_transcript_alpha(proof, w + 1, r->k, &alpha);
    for (size_t i = 0; i < 2 * n; i++)
        if (fr_eq(&alpha, &dom->leaves[i])) { ret = -2; free(r2); goto done; }
    /* barycentric evaluation at alpha, ec_fft.rs:455-491 */
    fr_t za, a0 = FR_ZERO, b0 = FR_ZERO, i0 = FR_ZERO, r0, t;
    ecfft_vanish_at(dom, 0, &alpha, &za);
    fr_t *ai = fr_vec(n), *di2 = fr_vec(n);
    for (size_t i = 0; i < n; i++) {
        fr_sub(&ai[i], &alpha, &dom->leaves[2 * i]);
        fr_sub(&di2[i], &dom->leaves[2 * i + 1], &alpha);
    }
    fr_batch_inv(ai, n);
    fr_batch_inv(di2, n);
    for (size_t i = 0; i < n; i++) {
        fr_t wi;
        fr_mul(&wi, &srs->bar_wts[i], &ai[i]);
        fr_mul(&t, &a[i], &wi);  fr_add(&a0, &a0, &t);
        fr_mul(&t, &b[i], &wi);  fr_add(&b0, &b0, &t);
        fr_mul(&t, &iv[i], &wi); fr_add(&i0, &i0, &t);
    }
    fr_mul(&a0, &a0, &za); fr_mul(&b0, &b0, &za); fr_mul(&i0, &i0, &za);
    fr_mul(&r0, &a0, &b0);
    fr_sub(&r0, &r0, &i0);
    /* K scalars, proving.rs:599-654; denom_inv = 1/(d_i - alpha) = -1/(alpha - d_i) */
    for (size_t i = 0; i < n; i++) {
        fr_t dinv, rv;
        fr_neg(&dinv, &ai[i]);
        fr_sub(&t, &a[i], &a0); fr_mul(&ka[i], &t, &dinv);
        fr_sub(&t, &b[i], &b0); fr_mul(&kb[i], &t, &dinv);
        fr_mul(&rv, &a[i], &b[i]);
        fr_sub(&rv, &rv, &iv[i]);
        fr_sub(&t, &rv, &r0);    fr_mul(&kr[2 * i], &t, &dinv);
        fr_sub(&t, &r2[i], &r0); fr_mul(&kr[2 * i + 1], &t, &di2[i]);
    }
    k233_msm(&kzg, ka, srs->g_k, 4 * n, nthreads);                   /* k_a | k_b | k_r are contiguous */
    xsk233_encode(proof + 30, &kzg);
    fr_to_le29(proof + 60, &a0);
    fr_to_le29(proof + 89, &b0);
    free(ai); free(di2); free(r2);
done:
    if (!stages) free(buf);
    return ret;
}

int dv_verify(const trapdoor_t *td, const fr_t *pub, size_t k, const uint8_t proof[118]) {
    k233_pt P, K, G, lhs, t1, t2;
    int ok = xsk233_decode(&P, proof) != 0;
    ok &= xsk233_decode(&K, proof + 30) != 0;
    fr_t a0, b0, alpha, i0, r0, u0, v0, d2, t;
    ok &= fr_from_le29(&a0, proof + 60);
    ok &= fr_from_le29(&b0, proof + 89);
    if (!ok) return 0;
    dv_transcript_alpha(proof, pub, k, &alpha);
    monomial_eval(pub, k, &alpha, &i0);
    fr_mul(&r0, &a0, &b0);
    fr_sub(&r0, &r0, &i0);
    fr_sqr(&d2, &td->delta);
    fr_mul(&t, &td->delta, &b0);
    fr_add(&u0, &a0, &t);
    fr_mul(&t, &d2, &r0);
    fr_add(&u0, &u0, &t);
    fr_mul(&u0, &u0, &td->epsilon);
    fr_sub(&v0, &td->tau, &alpha);
    fr_mul(&v0, &v0, &td->epsilon);
    k233_generator(&G);
    k233_mul_fr(&t1, &K, &v0);
    k233_mul_fr(&t2, &G, &u0);
    k233_add(&lhs, &t1, &t2);
    return k233_eq(&lhs, &P);
}
