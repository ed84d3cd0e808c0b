/*
 * ORACLE (test infrastructure only -- never linked into the product library).
 *
 * ECFFT over Fr on the curve-derived domain of DV-Pari.  In the reference the transform lives in
 * crate `ecfft` (git alpenlabs/ecfft rev 9c6cac7, Cargo.toml:39, Cargo.lock:322-324), which is NOT
 * vendored under /root/reference.  What is restated here:
 *   - the domain: curve, generator of order 2^28 and coset offset hard-coded at
 *     /root/reference/src/ec_fft.rs:205-229; leaf_i = x(C + i*G_N), G_N = 2^(28-log2 N) * G_28
 *     (ec_fft.rs:93-170); D = even leaves, D' = odd leaves (ec_fft.rs:179-189);
 *   - FFTree::extend(evals, Moiety::S1) (call site /root/reference/src/proving.rs:410-422):
 *     the unique degree < n interpolant of evaluations on D, evaluated on D'.  Restated from the
 *     ECFFT construction (Ben-Sasson, Carmon, Kopparty, Levit, "Elliptic Curve Fast Fourier
 *     Transform", 2021): 2-isogeny psi(x) = x + t/(x - x0), decomposition
 *     P(s) = (P0(psi(s)) + s P1(psi(s))) v(s)^(m/2-1).
 * The outputs are mathematically unique, so parity is pinned by comparing with brute-force Lagrange
 * interpolation (tests/test_oracle_ecfft.py), exactly what the reference's own test does
 * (ec_fft.rs:883-907).
 */
#ifndef DVP_ORACLE_ECFFT_H
#define DVP_ORACLE_ECFFT_H
#include "fr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int log_n2;      /* log2 of the number of leaves N = 2n */
    size_t n2;       /* N */
    fr_t *leaves;    /* N leaves of the top layer, natural order */
    /* level k (k = 0 .. log2(n) - 1) handles sub-problems of size m = n >> k */
    int levels;
    fr_t **dec;      /* dec[k]: (m/2) 2x2 matrices, row-major 4 Fr each: source pairs -> (P0, P1) */
    fr_t **rec;      /* rec[k]: (m/2) 2x2 matrices: (P0, P1) -> target pairs */
    /* isogeny chain of the top tree, level k: psi_k(x) = x + t_k/(x - x0_k) */
    fr_t *x0, *t;    /* log_n2 entries each */
    fr_t *last;      /* the single leaf of the bottom layer reached from leaf 0 of D resp. D' */
} ecfft_domain;

/* Build leaves, isogenies and the extend matrices (D -> D') for N = 2^log_n2 leaves. */
ecfft_domain *ecfft_domain_new(int log_n2);
/* leaves, isogeny chain and lower layers only (no extend matrices): enough for the chain-rule helpers and
 * ecfft_bary_eval at sizes where the matrices would take minutes on one core */
ecfft_domain *ecfft_domain_new_light(int log_n2);
void ecfft_domain_free(ecfft_domain *d);
/* out[i] = value at D'[i] of the interpolant of in[] on D; n = N/2 elements */
void ecfft_extend(const ecfft_domain *d, const fr_t *in, fr_t *out);
/* Vanishing polynomial of D (shift = 0) or D' (shift = 1) at x, via Z_S(x) = v(x)^(|S|/2) Z_psi(S)(psi(x)) */
void ecfft_vanish_at(const ecfft_domain *d, int shift, const fr_t *x, fr_t *out);
/* Z'_S(s_i) for all roots s_i of S = D (shift 0) or D' (shift 1): n values */
void ecfft_vanish_derivative_on_roots(const ecfft_domain *d, int shift, fr_t *out);
/* Z_S evaluated on the n points of the other half-domain */
void ecfft_vanish_on_other(const ecfft_domain *d, int shift, fr_t *out);

/* O(n) barycentric evaluation of the interpolant of evals on D at nx points outside D (ec_fft.rs:455-491) */
void ecfft_bary_eval(const ecfft_domain *d, const fr_t *evals, const fr_t *xs, size_t nx, fr_t *out);

#ifdef __cplusplus
}
#endif
#endif
