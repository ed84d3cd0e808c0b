/*
 * ORACLE (test infrastructure only -- never linked into the product library).
 *
 * GF(2^233) = GF(2)[x]/(x^233 + x^74 + 1), the base field of sect233k1.
 * In the reference this arithmetic is inside crate `xs233-sys =0.2.0`
 * (T. Pornin's xs233 C library; /root/reference/Cargo.toml:38, Cargo.lock:902-905),
 * which is NOT vendored under /root/reference. This file restates the published
 * algorithms (NIST FIPS 186 trinomial reduction, Itoh-Tsujii inversion, half-trace)
 * and is pinned against OpenSSL's BN_GF2m_* in tests/test_oracle_gf233.py.
 *
 * Elements: 4 x u64 little-endian, bit i of the polynomial = bit (i%64) of w[i/64];
 * always fully reduced (bits >= 233 are zero).
 */
#ifndef DVP_ORACLE_GF233_H
#define DVP_ORACLE_GF233_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t w[4]; } gf_t;

extern const gf_t GF_ZERO, GF_ONE;

void gf_set_portable(int on); /* 1 = force the bit-serial multiplier (cross-check) */
void gf_add(gf_t *r, const gf_t *a, const gf_t *b);
void gf_mul(gf_t *r, const gf_t *a, const gf_t *b);
void gf_sqr(gf_t *r, const gf_t *a);
void gf_sqr_n(gf_t *r, const gf_t *a, int n);
void gf_inv(gf_t *r, const gf_t *a); /* 0 -> 0 */
void gf_sqrt(gf_t *r, const gf_t *a);
int  gf_trace(const gf_t *a);
void gf_halftrace(gf_t *r, const gf_t *a); /* r^2 + r = a when trace(a) == 0 */
int  gf_is_zero(const gf_t *a);
int  gf_eq(const gf_t *a, const gf_t *b);
/* 30-byte little-endian; decode returns 0 if any of the top 7 bits is set */
void gf_to_le30(uint8_t out[30], const gf_t *a);
int  gf_from_le30(gf_t *r, const uint8_t in[30]);

#ifdef __cplusplus
}
#endif
#endif
