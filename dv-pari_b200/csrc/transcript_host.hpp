// Fiat-Shamir transcript of the prover on the host (product code, not the oracle).
//
// Mirrors Transcript::{srs_hash,circuit_info_hash,witness_commitment_hash,public_input_hash,output}
// (/root/reference/src/proving.rs:71-198): alpha = LE(blake3(blake3(H_srs|H_circ) | blake3(H_wit|H_pub)))
// with bytes 28..31 cleared; H_srs = H_circ = blake3("") because upstream hashes empty buffers.
// BLAKE3 here covers inputs of at most one chunk (1024 bytes): k public inputs use 29 k bytes.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace dvp {
namespace host {

struct Blake3Small {
    static inline uint32_t ror(uint32_t v, int s) { return (v >> s) | (v << (32 - s)); }
    static inline void quarter(uint32_t *v, int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] += v[b] + x; v[d] = ror(v[d] ^ v[a], 16);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 12);
        v[a] += v[b] + y; v[d] = ror(v[d] ^ v[a], 8);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 7);
    }
    static void compress(uint32_t chain[8], const uint8_t *block, uint32_t len, uint32_t flags) {
        static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        static const int sigma[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        uint32_t msg[16], v[16];
        for (int i = 0; i < 16; i++) std::memcpy(&msg[i], block + 4 * i, 4); // little-endian host
        for (int i = 0; i < 8; i++) v[i] = chain[i];
        for (int i = 0; i < 4; i++) v[8 + i] = iv[i];
        v[12] = 0; v[13] = 0; v[14] = len; v[15] = flags;
        for (int round = 0; round < 7; round++) {
            for (int col = 0; col < 4; col++) quarter(v, col, 4 + col, 8 + col, 12 + col, msg[2 * col], msg[2 * col + 1]);
            for (int dg = 0; dg < 4; dg++)
                quarter(v, dg, 4 + (dg + 1) % 4, 8 + (dg + 2) % 4, 12 + (dg + 3) % 4, msg[8 + 2 * dg], msg[9 + 2 * dg]);
            uint32_t nx[16];
            for (int i = 0; i < 16; i++) nx[i] = msg[sigma[i]];
            std::memcpy(msg, nx, sizeof msg);
        }
        for (int i = 0; i < 8; i++) chain[i] = v[i] ^ v[i + 8];
    }
    // returns false if the input is longer than one chunk
    static bool hash(const uint8_t *data, size_t len, uint8_t out[32]) {
        if (len > 1024) return false;
        uint32_t chain[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                             0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        uint32_t flags = 1; // CHUNK_START
        size_t pos = 0;
        while (len - pos > 64) {
            compress(chain, data + pos, 64, flags);
            flags = 0;
            pos += 64;
        }
        uint8_t last[64] = {0};
        if (len > pos) std::memcpy(last, data + pos, len - pos);
        compress(chain, last, (uint32_t)(len - pos), flags | 2 | 8); // CHUNK_END | ROOT
        std::memcpy(out, chain, 32);
        return true;
    }
};

// commit_p: 30-byte encoding; pub_le29: k x 29 bytes canonical little-endian.  alpha_le32: 32 bytes LE.
inline bool transcript_alpha(const uint8_t commit_p[30], const uint8_t *pub_le29, size_t k, uint8_t alpha_le32[32]) {
    uint8_t h_empty[32], h_wit[32], h_pub[32], two[64], compile_time[32], run_time[32];
    if (!Blake3Small::hash(nullptr, 0, h_empty)) return false;
    if (!Blake3Small::hash(commit_p, 30, h_wit)) return false;
    if (!Blake3Small::hash(pub_le29, 29 * k, h_pub)) return false;
    std::memcpy(two, h_empty, 32); std::memcpy(two + 32, h_empty, 32);
    Blake3Small::hash(two, 64, compile_time);
    std::memcpy(two, h_wit, 32); std::memcpy(two + 32, h_pub, 32);
    Blake3Small::hash(two, 64, run_time);
    std::memcpy(two, compile_time, 32); std::memcpy(two + 32, run_time, 32);
    Blake3Small::hash(two, 64, alpha_le32);
    std::memset(alpha_le32 + 28, 0, 4); // 224-bit challenge, proving.rs:186-190
    return true;
}

} // namespace host
} // namespace dvp
