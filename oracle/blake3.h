/*
 * ORACLE (test infrastructure only).  BLAKE3 (crate blake3 1.8.2, Cargo.lock:188-192, not vendored)
 * restated from the published specification (any length: 1024-byte chunks, binary tree of parent
 * nodes) -- Transcript::public_input_hash (/root/reference/src/proving.rs:149-161) hashes 29 bytes per public input.
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
/* blake3::hash for any length (chunk tree, recursive form); always returns 0 */
int blake3_hash(const uint8_t *in, size_t len, uint8_t out[32]);
int blake3_hash_small(const uint8_t *in, size_t len, uint8_t out[32]); /* same function, historical name */
#ifdef __cplusplus
}
#endif
#endif
