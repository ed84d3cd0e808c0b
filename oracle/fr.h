/*
 * ORACLE (test infrastructure only -- never linked into the product library).
 *
 * Fr: the 232-bit prime scalar field of sect233k1, restated from the reference's
 * `Fr = Fp256<MontBackend<FqConfig,4>>` (/root/reference/src/curve.rs:16-22):
 *   modulus p = 0x8000000000000000000000000000069D5BB915BCD46EFB1AD5F173ABDF
 *   in-memory form = 4 x u64 little-endian limbs of (v * 2^256 mod p), fully reduced.
 * The arithmetic itself lives in ark-ff 0.5.0 (Cargo.lock:38-42), which is not
 * vendored; this is the textbook CIOS Montgomery algorithm, pinned by the constants
 * verified in SURVEY.md section 4.3 (R mod p, R^2 mod p, -p^-1 mod 2^64) and by
 * tests/test_oracle_fr.py against Python big integers.
 */
#ifndef DVP_ORACLE_FR_H
#define DVP_ORACLE_FR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } fr_t; /* Montgomery form, < p */

extern const fr_t FR_ZERO, FR_ONE /* = R mod p */;
extern const uint64_t FR_P[4];

void fr_add(fr_t *r, const fr_t *a, const fr_t *b);
void fr_sub(fr_t *r, const fr_t *a, const fr_t *b);
void fr_neg(fr_t *r, const fr_t *a);
void fr_mul(fr_t *r, const fr_t *a, const fr_t *b);
void fr_sqr(fr_t *r, const fr_t *a);
void fr_inv(fr_t *r, const fr_t *a);              /* 0 -> 0 */
void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e);
int  fr_is_zero(const fr_t *a);
int  fr_eq(const fr_t *a, const fr_t *b);
void fr_from_u64(fr_t *r, uint64_t v);
/* canonical (non-Montgomery) 4-limb integer <-> Montgomery; input reduced mod p */
void fr_from_canonical(fr_t *r, const uint64_t c[4]);
void fr_to_canonical(uint64_t c[4], const fr_t *a);
/* 32-byte big-endian, reduced mod p: Fr::from_be_bytes_mod_order (gnark_r1cs.rs:283-288) */
void fr_from_be32_mod(fr_t *r, const uint8_t b[32]);
/* 29-byte canonical little-endian (io_utils.rs:127, proving.rs:153-156) */
void fr_to_le29(uint8_t out[29], const fr_t *a);
int  fr_from_le29(fr_t *r, const uint8_t in[29]); /* returns 0 if >= p */
/* ark_ff::batch_inversion semantics: zeros stay zero */
void fr_batch_inv(fr_t *v, size_t n);
/* s0 = sum k_i, s1 = sum i k_i (closed form of an MSM over the points A + i Q) */
void fr_sum_weighted(const fr_t *k, size_t n, fr_t *s0, fr_t *s1);

#ifdef __cplusplus
}
#endif
#endif
