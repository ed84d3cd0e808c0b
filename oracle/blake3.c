/* ORACLE (test infrastructure only). See blake3.h. */
#include "blake3.h"
#include <string.h>

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { CHUNK_START = 1, CHUNK_END = 2, ROOT = 8 };

static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
static void compress(const uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint64_t counter, uint32_t flags,
                     uint32_t out[8]) {
    uint32_t m[16], s[16], t[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) |
               ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32);
    s[14] = block_len;
    s[15] = flags;
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) out[i] = s[i] ^ s[i + 8];
}
enum { PARENT = 4 };

/* chaining value of the chunk with index `counter` (1..1024 bytes, or 0 bytes for the empty input); `last_flags` is
 * OR-ed into the final block (ROOT when the chunk is the whole message) */
static void chunk_cv(const uint8_t *in, size_t len, uint64_t counter, uint32_t last_flags, uint32_t out[8]) {
    uint32_t cv[8];
    memcpy(cv, IV, sizeof cv);
    size_t off = 0;
    uint32_t flags = CHUNK_START;
    while (len - off > 64) {
        compress(cv, in + off, 64, counter, flags, cv);
        flags = 0;
        off += 64;
    }
    uint8_t block[64];
    memset(block, 0, 64);
    if (len - off) memcpy(block, in + off, len - off);
    compress(cv, block, (uint32_t)(len - off), counter, flags | CHUNK_END | last_flags, out);
}
static void parent_cv(const uint32_t l[8], const uint32_t r[8], uint32_t extra, uint32_t out[8]) {
    uint8_t block[64];
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 4; b++) {
            block[4 * i + b] = (uint8_t)(l[i] >> (8 * b));
            block[32 + 4 * i + b] = (uint8_t)(r[i] >> (8 * b));
        }
    compress(IV, block, 64, 0, PARENT | extra, out);
}
/* recursive form of the BLAKE3 tree: the left subtree takes the largest power of two of chunks that leaves at least
 * one byte on the right (spec section 2.1) */
static void subtree_cv(const uint8_t *in, size_t len, uint64_t first_chunk, int is_root, uint32_t out[8]) {
    if (len <= 1024) {
        chunk_cv(in, len, first_chunk, is_root ? ROOT : 0, out);
        return;
    }
    size_t left = 1024;
    while (2 * left < len) left *= 2;
    uint32_t l[8], r[8];
    subtree_cv(in, left, first_chunk, 0, l);
    subtree_cv(in + left, len - left, first_chunk + left / 1024, 0, r);
    parent_cv(l, r, is_root ? ROOT : 0, out);
}
/* blake3::hash for any length */
int blake3_hash(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    subtree_cv(in, len, 0, 1, cv);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)cv[i]; out[4 * i + 1] = (uint8_t)(cv[i] >> 8);
        out[4 * i + 2] = (uint8_t)(cv[i] >> 16); out[4 * i + 3] = (uint8_t)(cv[i] >> 24);
    }
    return 0;
}
/* historical name (inputs of one chunk); now the same function */
int blake3_hash_small(const uint8_t *in, size_t len, uint8_t out[32]) { return blake3_hash(in, len, out); }
