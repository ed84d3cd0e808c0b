// Fiat-Shamir transcript of the prover on the host (product code, not the oracle).
//
// Mirrors Transcript::{srs_hash,circuit_info_hash,witness_commitment_hash,public_input_hash,output}
// (/root/reference/src/proving.rs:71-198): alpha = LE(blake3(blake3(H_srs|H_circ) | blake3(H_wit|H_pub)))
// with bytes 28..31 cleared; H_srs = H_circ = blake3("") because upstream hashes empty buffers.
// BLAKE3 is complete (any input length, chunk tree), so any number of public inputs hashes like upstream.
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
    // One chunk (at most 1024 bytes, counter = its index): the chaining value after all blocks but the last, and the
    // last block's (words, length, flags) left pending so that the caller can still add ROOT to it.
    struct Pending {
        uint32_t cv[8];
        uint8_t block[64];
        uint32_t len, flags;
        uint64_t counter;
    };
    static void compress_ctr(uint32_t chain[8], const uint8_t *block, uint32_t len, uint64_t counter, uint32_t flags) {
        static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        static const int sigma[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        uint32_t msg[16], v[16];
        for (int i = 0; i < 16; i++) std::memcpy(&msg[i], block + 4 * i, 4);
        for (int i = 0; i < 8; i++) v[i] = chain[i];
        for (int i = 0; i < 4; i++) v[8 + i] = iv[i];
        v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = len; v[15] = flags;
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
    static void init_iv(uint32_t cv[8]) {
        static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        std::memcpy(cv, iv, sizeof iv);
    }
    static Pending chunk(const uint8_t *data, size_t len, uint64_t index) {
        Pending p;
        init_iv(p.cv);
        uint32_t flags = 1; // CHUNK_START
        size_t pos = 0;
        while (len - pos > 64) {
            compress_ctr(p.cv, data + pos, 64, index, flags);
            flags = 0;
            pos += 64;
        }
        std::memset(p.block, 0, 64);
        if (len > pos) std::memcpy(p.block, data + pos, len - pos);
        p.len = (uint32_t)(len - pos);
        p.flags = flags | 2; // CHUNK_END
        p.counter = index;
        return p;
    }
    static Pending parent(const uint32_t left[8], const uint32_t right[8]) {
        Pending p;
        init_iv(p.cv);
        std::memcpy(p.block, left, 32);
        std::memcpy(p.block + 32, right, 32);
        p.len = 64;
        p.flags = 4; // PARENT
        p.counter = 0;
        return p;
    }
    static void finish(const Pending &p, uint32_t extra_flags, uint32_t out[8]) {
        std::memcpy(out, p.cv, 32);
        compress_ctr(out, p.block, p.len, p.counter, p.flags | extra_flags);
    }
    // BLAKE3 of any length (blake3::hash, the call of Transcript::public_input_hash, proving.rs:149-161: 29 k bytes for
    // k public inputs): chunks of 1024 bytes, chaining values merged as a binary tree (left subtrees are the largest
    // powers of two), ROOT on the last compression.
    static bool hash(const uint8_t *data, size_t len, uint8_t out[32]) {
        uint32_t stack[64][8];
        int depth = 0;
        uint64_t index = 0;
        size_t pos = 0;
        while (len - pos > 1024) {
            uint32_t cv[8];
            finish(chunk(data + pos, 1024, index), 0, cv);
            pos += 1024;
            index++;
            // a completed subtree for every trailing one bit... merge while the chunk count is even
            for (uint64_t total = index; (total & 1) == 0; total >>= 1) {
                uint32_t merged[8];
                finish(parent(stack[depth - 1], cv), 0, merged);
                std::memcpy(cv, merged, 32);
                depth--;
            }
            std::memcpy(stack[depth++], cv, 32);
        }
        Pending node = chunk(data + pos, len - pos, index);
        while (depth > 0) {
            uint32_t cv[8];
            finish(node, 0, cv);
            node = parent(stack[--depth], cv);
        }
        uint32_t root[8];
        finish(node, 8, root); // ROOT
        std::memcpy(out, root, 32);
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
