/*
 * ORACLE (test infrastructure only -- never linked into the product library).
 *
 * CPU restatement of the DV-Pari protocol layer of the reference:
 *   r1cs_eval     get_matrix_evaluations_from_witness  /root/reference/src/proving.rs:348-403
 *                 (+ eval_row gnark_r1cs.rs:273-280, Vandermonde C' = C - D gnark_r1cs.rs:333-386,
 *                  evaluate_monomial_basis_poly gnark_r1cs.rs:391-399)
 *   transcript    Transcript::{public_input_hash,witness_commitment_hash,output}  proving.rs:71-198
 *   dv_setup      SRS::verifier_runs_setup / compute_srs_matrices / accumulate_m_values
 *                 /root/reference/src/srs.rs:53-167,177-361 (artifacts only, no files)
 *   dv_prove      Proof::prove  /root/reference/src/proving.rs:426-688
 *   dv_verify     SRS::verify   /root/reference/src/srs.rs:374-428
 * Z_D(alpha), barycentric weights and Z_D on D' come from the chain rule of ecfft.h instead of the
 * vanish/exit/enter pipeline (same field elements; validated against brute force in the tests).
 */
#ifndef DVP_ORACLE_DVSNARK_H
#define DVP_ORACLE_DVSNARK_H
#include "ecfft.h"
#include "k233.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Sparse R1CS in file order (gnark_r1cs.rs:1-20): three CSR matrices over a shared coefficient table. */
typedef struct {
    size_t nrows;        /* rows present in the dump */
    size_t n;            /* padded constraint count = nrows.next_power_of_two() (gnark_r1cs.rs:291) */
    size_t k;            /* number of public inputs */
    size_t nwires;       /* 1 + k + #private */
    const uint32_t *rowptr[3]; /* L, R, O: nrows + 1 entries each */
    const uint32_t *wire[3];
    const uint32_t *coeff[3];
    const fr_t *coeffs;  /* Montgomery */
    size_t ncoeffs;
} r1cs_t;

typedef struct { fr_t tau, delta, epsilon; } trapdoor_t;

typedef struct {
    k233_pt *g_m;  /* nwires */
    k233_pt *g_q;  /* n */
    k233_pt *g_k;  /* 4n: g_k0 | g_k1 | g_k2 (proving.rs:666-673) */
    fr_t *z_vals2inv; /* n: 1 / Z_D(d'_i) */
    fr_t *bar_wts;    /* n: 1 / Z_D'(d_i) */
} srs_t;

/* a, b, c (= C'w), i on D.  Returns -1 if every row satisfies a*b = c + i, else the first bad row. */
long r1cs_eval(const r1cs_t *r, const ecfft_domain *dom, const fr_t *assignment, fr_t *a, fr_t *b, fr_t *c, fr_t *i);
/* alpha from the 30-byte commitment and the public inputs */
void dv_transcript_alpha(const uint8_t commit_p[30], const fr_t *pub, size_t k, fr_t *alpha);
/* the discrete logs of the SRS points (srs.rs:112-167 before the fixed-base multiplications) */
void dv_setup_scalars(const r1cs_t *r, const ecfft_domain *dom, const trapdoor_t *td, fr_t *sc_m, fr_t *sc_q, fr_t *sc_k,
                      fr_t *bar_wts, fr_t *z_vals2inv);
/* allocate and fill the SRS and prover precomputes */
srs_t *dv_setup(const r1cs_t *r, const ecfft_domain *dom, const trapdoor_t *td);
void dv_srs_free(srs_t *s);
/* proof118 = commit_p (30) | kzg_k (30) | a0 (29, LE) | b0 (29, LE).  Returns 0, or 1 + first bad row,
 * or -2 if alpha falls in the domain.  If stages != NULL it receives 12 vectors of n Fr:
 * a b c i a' b' c' i' q k_a k_b and 2n for k_r is written to stages + 11 n (so 13 n total). */
long dv_prove(const r1cs_t *r, const ecfft_domain *dom, const srs_t *srs, const fr_t *assignment, uint8_t proof118[118],
              fr_t *stages, int nthreads);
int dv_verify(const trapdoor_t *td, const fr_t *pub, size_t k, const uint8_t proof118[118]);

#ifdef __cplusplus
}
#endif
#endif
