// Artifact formats either side of the prover path (host code of the product, no device work):
//   Fr vector files    u64 LE count | 29-byte LE canonical elements      /root/reference/src/io_utils.rs:27-66,113-165
//   SP1 / gnark dumps  u32 nbCoeffs | 32-byte BE coefficients | u32 nbRows | rows of (nL nR nO | terms)
//                      Term = (u32 wire_id, u32 coeff_id), little-endian  /root/reference/src/gnark_r1cs.rs:1-20,121-185
//   witness files      u32 BE count | 32-byte BE elements                 /root/reference/src/gnark_r1cs.rs:58-77,188-199
//   SP1 public input   blake3(raw u64 LE), top 4 bytes cleared, as BE int /root/reference/src/gnark_r1cs.rs:218-236
//   FFTree files       "FFTR\0\0\0\0" | u64 len | node = u32 sections, u32 pad | 24-byte section metas | blobs,
//                      section 12 = the child tree as a nested node           /root/reference/src/tree_io.rs:1-15
// Everything lands in the layouts the C ABI takes: Montgomery limbs (Vec<Fr>) and CSR per matrix.
// (Point-vector files are u64 count | 30-byte encodings: their payload goes to dvp_srs_load unchanged.)
#include <cstring>
#include "../../include/dvpari.h"
#include "fr.cuh"
#include "transcript_host.hpp"

using namespace dvp;

namespace {
// value of 32 big-endian bytes mod p, as Montgomery limbs (Fr::from_be_bytes_mod_order)
fr fr_from_be32_mod(const uint8_t *b) {
    // split v = hi * 2^224 + lo with lo < 2^224 < p and hi < 2^32: both canonical
    uint32_t lo[8], hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < 8; w++) {
        const uint8_t *q = b + 28 - 4 * w; // word w (little-endian order) sits at bytes 28-4w .. 31-4w
        lo[w] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    hi[0] = lo[7];
    lo[7] = 0;
    // 2^224 in Montgomery form = fr_from_canonical of the limb pattern with bit 224 set
    uint32_t t224[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    const fr m224 = fr_from_canonical(t224);
    return fr_add(fr_from_canonical(lo), fr_mul(fr_from_canonical(hi), m224));
}
} // namespace

extern "C" {

int dvp_fr_from_le29(const uint8_t *in, size_t n, uint64_t *out_mont) {
    if ((!in || !out_mont) && n) return DVP_ERR_BAD_ARG;
    for (size_t i = 0; i < n; i++) {
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint8_t buf[32] = {0};
        memcpy(buf, in + 29 * i, 29);
        memcpy(c, buf, 32);
        if (fr_geq_p(c)) return DVP_ERR_BAD_ARG; // not canonical
        const fr m = fr_from_canonical(c);
        memcpy(out_mont + 4 * i, m.v, 32);
    }
    return DVP_OK;
}

int dvp_fr_to_le29(const uint64_t *in_mont, size_t n, uint8_t *out) {
    if ((!in_mont || !out) && n) return DVP_ERR_BAD_ARG;
    for (size_t i = 0; i < n; i++) {
        fr a;
        memcpy(a.v, in_mont + 4 * i, 32);
        uint32_t c[8];
        fr_to_canonical(c, a);
        uint8_t buf[32];
        memcpy(buf, c, 32);
        memcpy(out + 29 * i, buf, 29);
    }
    return DVP_OK;
}

int dvp_fr_from_be32_mod_order(const uint8_t *in, size_t n, uint64_t *out_mont) {
    if ((!in || !out_mont) && n) return DVP_ERR_BAD_ARG;
    for (size_t i = 0; i < n; i++) {
        const fr m = fr_from_be32_mod(in + 32 * i);
        memcpy(out_mont + 4 * i, m.v, 32);
    }
    return DVP_OK;
}

int dvp_sp1_public_input(uint64_t raw, uint64_t out_mont[4]) {
    if (!out_mont) return DVP_ERR_BAD_ARG;
    uint8_t le[8], h[32];
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(raw >> (8 * i));
    host::Blake3Small::hash(le, 8, h);
    memset(h, 0, 4); // the first four bytes of the big-endian number are masked off: 224 bits remain
    const fr m = fr_from_be32_mod(h);
    memcpy(out_mont, m.v, 32);
    return DVP_OK;
}

// blake3::hash and the Fiat-Shamir challenge of Transcript::output (proving.rs:137-197) as the prover computes them
// (host code; exported so that the transcript can be checked without a device)
int dvp_blake3(const uint8_t *data, size_t len, uint8_t out32[32]) {
    if (!out32 || (!data && len)) return DVP_ERR_BAD_ARG;
    host::Blake3Small::hash(data, len, out32);
    return DVP_OK;
}
int dvp_transcript_alpha(const uint8_t commit_p30[30], const uint64_t *public_mont, size_t k, uint64_t alpha_mont[4]) {
    if (!commit_p30 || !alpha_mont || (!public_mont && k)) return DVP_ERR_BAD_ARG;
    std::vector<uint8_t> pub29(29 * k + 1);
    for (size_t j = 0; j < k; j++) {
        fr x;
        memcpy(x.v, public_mont + 4 * j, 32);
        uint32_t c[8];
        fr_to_canonical(c, x);
        for (int i = 0; i < 29; i++) pub29[29 * j + i] = (uint8_t)(c[i >> 2] >> (8 * (i & 3)));
    }
    uint8_t al[32];
    if (!host::transcript_alpha(commit_p30, pub29.data(), k, al)) return DVP_ERR_INTERNAL;
    uint32_t alc[8];
    memcpy(alc, al, 32);
    const fr a = fr_from_canonical(alc);
    memcpy(alpha_mont, a.v, 32);
    return DVP_OK;
}

// First pass over a dump: sizes.  Returns DVP_ERR_BAD_ARG if the buffer is truncated or inconsistent.
int dvp_r1cs_dump_sizes(const uint8_t *buf, size_t len, size_t *ncoeffs, size_t *nrows, size_t nnz[3], size_t *max_wire) {
    if (!buf || !ncoeffs || !nrows || !nnz || !max_wire) return DVP_ERR_BAD_ARG;
    size_t pos = 0;
    auto rd32 = [&](uint32_t &v) -> bool {
        if (pos + 4 > len) return false;
        memcpy(&v, buf + pos, 4);
        pos += 4;
        return true;
    };
    uint32_t nc, nr;
    if (!rd32(nc)) return DVP_ERR_BAD_ARG;
    if (pos + (size_t)nc * 32 > len) return DVP_ERR_BAD_ARG;
    pos += (size_t)nc * 32;
    if (!rd32(nr)) return DVP_ERR_BAD_ARG;
    nnz[0] = nnz[1] = nnz[2] = 0;
    size_t mw = 0;
    for (uint32_t r = 0; r < nr; r++) {
        uint32_t cnt[3];
        for (int w = 0; w < 3; w++)
            if (!rd32(cnt[w])) return DVP_ERR_BAD_ARG;
        for (int w = 0; w < 3; w++) {
            if (pos + (size_t)cnt[w] * 8 > len) return DVP_ERR_BAD_ARG;
            for (uint32_t t = 0; t < cnt[w]; t++) {
                uint32_t wire, cid;
                memcpy(&wire, buf + pos, 4);
                memcpy(&cid, buf + pos + 4, 4);
                pos += 8;
                if (cid >= nc) return DVP_ERR_BAD_ARG;
                if (wire > mw) mw = wire;
            }
            nnz[w] += cnt[w];
        }
    }
    *ncoeffs = nc;
    *nrows = nr;
    *max_wire = mw;
    return DVP_OK;
}

// Second pass: fill the CSR arrays (rowptr[w]: nrows + 1, wire[w] / coeff[w]: nnz[w]) and the coefficient table.
// The capacities are the sizes the first pass reported; the walk is validated again (every read against len, every
// write against the capacities), so a buffer that changed or is truncated since dvp_r1cs_dump_sizes is an error, not an
// out-of-bounds access.
int dvp_r1cs_dump_parse(const uint8_t *buf, size_t len, size_t ncoeffs, size_t nrows, const size_t nnz[3],
                        uint32_t *const rowptr[3], uint32_t *const wire[3], uint32_t *const coeff[3],
                        uint64_t *coeffs_mont) {
    if (!buf || !nnz || !rowptr || !wire || !coeff || !coeffs_mont) return DVP_ERR_BAD_ARG;
    for (int w = 0; w < 3; w++)
        if (!rowptr[w] || (nnz[w] && (!wire[w] || !coeff[w]))) return DVP_ERR_BAD_ARG;
    size_t pos = 0;
    uint32_t nc, nr;
    if (len < 4) return DVP_ERR_BAD_ARG;
    memcpy(&nc, buf + pos, 4);
    pos += 4;
    if (nc != ncoeffs || (size_t)nc * 32 > len - pos) return DVP_ERR_BAD_ARG;
    for (uint32_t i = 0; i < nc; i++) {
        const fr m = fr_from_be32_mod(buf + pos);
        memcpy(coeffs_mont + 4 * (size_t)i, m.v, 32);
        pos += 32;
    }
    if (len - pos < 4) return DVP_ERR_BAD_ARG;
    memcpy(&nr, buf + pos, 4);
    pos += 4;
    if (nr != nrows) return DVP_ERR_BAD_ARG;
    size_t fill[3] = {0, 0, 0};
    for (int w = 0; w < 3; w++) rowptr[w][0] = 0;
    for (uint32_t r = 0; r < nr; r++) {
        uint32_t cnt[3];
        if (len - pos < 12) return DVP_ERR_BAD_ARG;
        memcpy(cnt, buf + pos, 12);
        pos += 12;
        for (int w = 0; w < 3; w++) {
            if ((size_t)cnt[w] > nnz[w] - fill[w] || (size_t)cnt[w] * 8 > len - pos) return DVP_ERR_BAD_ARG;
            for (uint32_t t = 0; t < cnt[w]; t++) {
                uint32_t cid;
                memcpy(&wire[w][fill[w]], buf + pos, 4);
                memcpy(&cid, buf + pos + 4, 4);
                if (cid >= nc) return DVP_ERR_BAD_ARG;
                coeff[w][fill[w]] = cid;
                pos += 8;
                fill[w]++;
            }
            rowptr[w][r + 1] = (uint32_t)fill[w];
        }
    }
    for (int w = 0; w < 3; w++)
        if (fill[w] != nnz[w]) return DVP_ERR_BAD_ARG;
    return DVP_OK;
}

// ---- FFTR tree files (tree_io.rs).  The blobs are ark-serialize "compressed" output: a Vec<T> / BinaryTree<T> is a
// u64 LE count followed by the elements, an Fr is 29 bytes LE canonical, a Mat2x2<Fr> is its four entries row-major.
// A BinaryTree is the heap array of crate ecfft: 2m entries for m leaves, entry 0 unused, the leaves in the upper half.
int dvp_fftree_file_sections(const uint8_t *file, size_t len, size_t depth, uint64_t off[13], uint64_t slen[13]) {
    if (!file || !off || !slen || len < 16) return DVP_ERR_BAD_ARG;
    if (memcmp(file, "FFTR\0\0\0\0", 8) != 0) return DVP_ERR_BAD_ARG; // "not an FFTR file", tree_io.rs:225,427
    uint64_t total;
    memcpy(&total, file + 8, 8);
    if (total > len - 16) return DVP_ERR_BAD_ARG;
    size_t base = 16, size = (size_t)total; // the node occupies file[base, base + size)
    for (size_t level = 0;; level++) {
        if (size < 8) return DVP_ERR_BAD_ARG;
        uint32_t count;
        memcpy(&count, file + base, 4);
        if (count > 13 || (size_t)8 + 24 * (size_t)count > size) return DVP_ERR_BAD_ARG;
        for (int i = 0; i < 13; i++) off[i] = slen[i] = 0;
        for (uint32_t s = 0; s < count; s++) {
            const uint8_t *m = file + base + 8 + 24 * (size_t)s;
            uint64_t o, l;
            memcpy(&o, m + 8, 8);
            memcpy(&l, m + 16, 8);
            if (m[0] > 12) return DVP_ERR_BAD_ARG; // "unknown section id", tree_io.rs:69
            if (o > size || l > size - o) return DVP_ERR_BAD_ARG;
            off[m[0]] = base + o; // offsets in the file are relative to the node (tree_io.rs:193-203)
            slen[m[0]] = l;
        }
        if (level == depth) return DVP_OK;
        if (!slen[12]) return DVP_ERR_BAD_ARG; // no subtree that deep
        base = (size_t)off[12];
        size = (size_t)slen[12];
    }
}

// FLeaves of the node `depth` subtrees below the root: *n_leaves = leaves of that tree (f.leaves().len());
// leaves_mont (n_leaves x 4, or NULL for the count alone) = f.leaves() as Montgomery limbs.
int dvp_fftree_file_leaves(const uint8_t *file, size_t len, size_t depth, size_t *n_leaves, uint64_t *leaves_mont) {
    if (!n_leaves) return DVP_ERR_BAD_ARG;
    uint64_t off[13], slen[13];
    const int rc = dvp_fftree_file_sections(file, len, depth, off, slen);
    if (rc) return rc;
    if (slen[0] < 8) return DVP_ERR_BAD_ARG; // "missing section", tree_io.rs:139
    uint64_t cnt;
    memcpy(&cnt, file + off[0], 8);
    if (cnt > (slen[0] - 8) / 29 || cnt * 29 + 8 != slen[0] || (cnt & (cnt - 1))) return DVP_ERR_BAD_ARG;
    *n_leaves = (size_t)(cnt / 2);
    if (!leaves_mont) return DVP_OK;
    return dvp_fr_from_le29(file + off[0] + 8 + 29 * (size_t)(cnt / 2), (size_t)(cnt / 2), leaves_mont);
}

// RecombineMatrices (which = 1) / DecomposeMatrices (which = 2) of that node: *count matrices (the whole heap array,
// entry 0 included), out (count x 16 u64, or NULL) = their entries as Montgomery limbs, row-major.
int dvp_fftree_file_matrices(const uint8_t *file, size_t len, size_t depth, int which, size_t *count, uint64_t *out) {
    if (!count || (which != 1 && which != 2)) return DVP_ERR_BAD_ARG;
    uint64_t off[13], slen[13];
    const int rc = dvp_fftree_file_sections(file, len, depth, off, slen);
    if (rc) return rc;
    if (slen[which] < 8) return DVP_ERR_BAD_ARG;
    uint64_t cnt;
    memcpy(&cnt, file + off[which], 8);
    if (cnt > (slen[which] - 8) / 116 || cnt * 116 + 8 != slen[which]) return DVP_ERR_BAD_ARG;
    *count = (size_t)cnt;
    if (!out) return DVP_OK;
    return dvp_fr_from_le29(file + off[which] + 8, 4 * (size_t)cnt, out);
}

} // extern "C"
