/*
 * ORACLE (test infrastructure only -- never linked into the product library).
 *
 * sect233k1 (NIST K-233): y^2 + xy = x^3 + 1 over GF(2^233), cofactor 4, prime
 * subgroup order r = p of fr.h, and the `xsk233` prime-order group that the
 * reference reaches through crate `xs233-sys =0.2.0` (not vendored):
 *   call sites /root/reference/src/curve.rs:13,72,81,87,97,106,118,134.
 *
 * Restated model (T. Pornin, "Efficient and Complete Formulas for Binary Curves",
 * ePrint 2022/1325): with N = (0,1) the point of order 2, the group is
 * { P + N : P in E[r] } with law Q1 (+) Q2 = Q1 + Q2 + N and neutral N.  A group
 * element Q = (x,y) is encoded as w = (y + 1)/x (the sqrt(S/T) coordinate of xs233's
 * extended (X:S:Z:T) form), 233 bits little-endian in 30 bytes; the neutral encodes
 * as 30 zero bytes.  Since (+)_i k_i (.) Q_i = (sum_i k_i (Q_i + N)) + N, every bulk
 * computation runs on the ordinary K-233 points P_i = Q_i + N in E[r].
 *
 * PARITY UNPINNED for the 30-byte codec: the reference holds no byte-level vector
 * for xsk233_encode/decode (only round-trips, curve.rs:236-248, io_utils.rs:253-267),
 * and the crate source is unavailable offline.  The group law itself is pinned
 * against OpenSSL NID_sect233k1 in tests/test_oracle_k233.py.
 */
#ifndef DVP_ORACLE_K233_H
#define DVP_ORACLE_K233_H
#include "fr.h"
#include "gf233.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Affine K-233 point; inf != 0 is the point at infinity (x,y ignored). */
typedef struct { gf_t x, y; int inf; } k233_pt;

void k233_generator(k233_pt *r); /* the standard NIST generator (in E[r]) */

int  k233_on_curve(const k233_pt *p);
void k233_neg(k233_pt *r, const k233_pt *p);
void k233_add(k233_pt *r, const k233_pt *p, const k233_pt *q); /* complete, affine */
void k233_dbl(k233_pt *r, const k233_pt *p);
int  k233_eq(const k233_pt *p, const k233_pt *q);
/* k * P, k = little-endian bytes (xsk233_mul_frob argument convention, curve.rs:113-126) */
void k233_mul_bytes(k233_pt *r, const k233_pt *p, const uint8_t *k, size_t klen);
/* k * P with k an Fr in Montgomery form: fr_to_le_bytes + point_scalar_mul (curve.rs:113-126,162-182) */
void k233_mul_fr(k233_pt *r, const k233_pt *p, const fr_t *k);
/* the two implementations behind it: tau-adic width-4 TNAF (the reference's xsk233_mul_frob; the default) and
 * width-4 NAF double-and-add (kept as the cross-check) */
void k233_mul_fr_tau(k233_pt *r, const k233_pt *p, const fr_t *k);
void k233_mul_fr_wnaf(k233_pt *r, const k233_pt *p, const fr_t *k);

/* xsk233 codec on the E[r] representative P = Q + N of the group element Q */
void xsk233_encode(uint8_t out[30], const k233_pt *p);
int  xsk233_decode(k233_pt *p, const uint8_t in[30]); /* non-zero on success (curve.rs:107) */

/* multi_scalar_mul (curve.rs:141-158): per-point scalar mul, then sum. nthreads<=0: all cores */
void k233_msm(k233_pt *r, const fr_t *scalars, const k233_pt *points, size_t n, int nthreads);

/* bulk helpers for fixtures and benchmarks (OpenMP over all cores) */
void k233_chain_points(k233_pt *out, size_t n, const k233_pt *p0, const k233_pt *q); /* out[i] = p0 + i q */
void k233_mul_batch(k233_pt *out, const k233_pt *p, const fr_t *k, size_t n);         /* out[i] = k_i p */
void xsk233_encode_batch(uint8_t *out30, const k233_pt *pts, size_t n);
long xsk233_decode_batch(k233_pt *pts, const uint8_t *in30, size_t n); /* first invalid index or -1 */

#ifdef __cplusplus
}
#endif
#endif
