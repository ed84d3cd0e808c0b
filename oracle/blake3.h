/*
 * ORACLE (test infrastructure only).  BLAKE3 (crate blake3 1.8.2, Cargo.lock:188-192, not vendored)
 * restated from the published specification for inputs of at most one chunk (1024 bytes) -- the
 * transcript of /root/reference/src/proving.rs:71-198 never hashes more than 64 bytes.
 * Pinned against the Python `blake3` package and the reference's KAT (gnark_r1cs.rs:497-504) in
 * tests/test_oracle_protocol.py.
 */
#ifndef DVP_ORACLE_BLAKE3_H
#define DVP_ORACLE_BLAKE3_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* returns 0 on success, -1 if len > 1024 */
int blake3_hash_small(const uint8_t *in, size_t len, uint8_t out[32]);
#ifdef __cplusplus
}
#endif
#endif
