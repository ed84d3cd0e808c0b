// sect233k1 multi-scalar multiplication for sm_100a.
//
// Replaces curve::multi_scalar_mul (/root/reference/src/curve.rs:141-158; call sites
// src/proving.rs:463,512,680).  The reference does one tau-adic scalar multiplication per point and
// a tree sum; the group element is unique, so any algorithm matches bit for bit.  Here:
//
//   recode      Montgomery Fr -> canonical 232-bit integer -> W signed digits (windows of even width)
//   sort        global counting sort of the (point, window) entries by bucket (histogram, scan, scatter).
//               Resident SRS slots carry tables T[j] = 2^(off_j) P, so all windows share ONE bucket set.
//   plan        per round one launch: scans over the segment lengths + one (a, b, out) descriptor per addition
//   accumulate  every bucket is a segment; segments are tree-reduced in rounds of independent
//               affine additions sharing one inversion (Montgomery trick, hierarchical), on 2 lanes (streams)
//   reduce      sum_b (b+1) B_b per virtual window without a serial running sum: rows/columns of the bucket
//               matrix (level A), then per-bit subset sums (level B); batched-affine rounds while large,
//               inversion-free Lopez-Dahab trees (one block per segment) for the rest
//   tail        the host folds the per-bit partial sums with one double-and-add pass (PCLMULQDQ)
//
// Everything between the scalars and the partial sums stays on the device; a single 8-byte-per-lane
// read-back (entries and longest bucket) sizes the rounds.  Also here: the batched fixed-base multiplication
// (mulgen) and the table builder, both expressed as tree rounds.
#include "msm.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
#include "../../include/dvpari.h"
#include "fr.cuh"
#include "host_gf.hpp"
#include "gf233_warp.cuh"
#include "k233_ld.cuh"

namespace dvp {

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) {                                                                    \
            fprintf(stderr, "[dvpari] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__,  \
                    __LINE__);                                                                      \
            return DVP_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

static inline uint32_t cdiv(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        p = nullptr;
        return DVP_ERR_OOM;
    }
    cap = want;
    return 0;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

__global__ void k_latency_probe(int mode, int iters, const gf *__restrict__ tabs, uint32_t *__restrict__ sink);
// microseconds per operation for a single warp running a dependent chain
int latency_probe(MsmEngine &E, int mode, int iters, float *us_per_op) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (E.hb.reserve(64)) return DVP_ERR_OOM;
    k_latency_probe<<<1, 32, 0, E.stream>>>(mode, 2, E.msqr_tabs.as<gf>(), E.hb.as<uint32_t>());
    cudaEventRecord(e0, E.stream);
    k_latency_probe<<<1, 32, 0, E.stream>>>(mode, iters, E.msqr_tabs.as<gf>(), E.hb.as<uint32_t>());
    cudaEventRecord(e1, E.stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) return DVP_ERR_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *us_per_op = ms * 1e3f / iters;
    return 0;
}

int choose_window_bits(size_t n) {
    // adds ~ n*W + tail(2^(c-1)*W); the tail rounds are latency-bound, so stay a little below the
    // arithmetic optimum
    // (plain layout on the persistent path, profiles/r2w_cwin_sweep_plain.log: 2^12 / 2^13 points c = 9 0.84 / 0.93 ms
    // against c = 8 0.91 / 1.03; 2^15 / 2^16 c = 13 1.21 / 1.53 against c = 11 1.29 / 1.70)
    if (n <= (1u << 10)) return 6;
    if (n <= (1u << 13)) return 9;
    if (n <= (1u << 18)) return 13;
    if (n <= (1u << 21)) return 15;
    return 16;
}

// ------------------------------------------------------------------------------------------------
// recode + histogram
// ------------------------------------------------------------------------------------------------
// Window j covers bits [off_j, off_j + width_j) with width_j = base + (j < rem), off_j = j*base + min(j, rem): the 233
// scalar bits are spread evenly over the W windows so that no window is nearly empty (a short top window would put
// all n entries into a handful of buckets and double the tree depth).  Two bucket layouts:
//   separate bucket sets (uniform = 0): key = j*nb + d - 1, the point of an entry is P_i;
//   one shared bucket set (uniform = 1, precomputed tables T[j] = 2^(off_j) P): key = d - 1, the point is T[j][i].
__global__ void k_recode_count(const uint32_t *__restrict__ scalars, uint32_t n, int base, int rem, int W, uint32_t nb,
                               int uniform, uint32_t *__restrict__ keys, uint32_t *__restrict__ seg_len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr a;
    const uint4 *q = reinterpret_cast<const uint4 *>(scalars + (size_t)i * 8);
    uint4 lo = q[0], hi = q[1];
    a.v[0] = lo.x; a.v[1] = lo.y; a.v[2] = lo.z; a.v[3] = lo.w;
    a.v[4] = hi.x; a.v[5] = hi.y; a.v[6] = hi.z; a.v[7] = hi.w;
    uint32_t k[9];
    fr_to_canonical(k, a);
    k[8] = 0;
    uint32_t carry = 0;
    for (int j = 0; j < W; j++) {
        const int c = base + (j < rem ? 1 : 0);
        const int bit = j * base + min(j, rem), w = bit >> 5, sh = bit & 31;
        const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
        uint32_t raw = 0;
        if (w < 8) {
            uint64_t two = (uint64_t)k[w] | ((uint64_t)k[w + 1] << 32);
            raw = (uint32_t)(two >> sh) & mask;
        }
        raw += carry;
        uint32_t d, neg;
        if (raw > half) {
            d = (1u << c) - raw;
            neg = 1;
            carry = 1;
        } else {
            d = raw;
            neg = 0;
            carry = 0;
        }
        uint32_t out = 0xffffffffu;
        if (d) {
            const uint32_t key = (uniform ? 0u : (uint32_t)j * nb) + d - 1;
            out = (key << 1) | neg;
            atomicAdd(&seg_len[key], 1u);
        }
        keys[(size_t)j * n + i] = out;
    }
}

// counting sort, second half: entry = index of the point in the source list | negate << 31, where the source
// list is the point range itself (ebase_stride = 0) or the tables (window j starts at j * ebase_stride)
__global__ void k_scatter(const uint32_t *__restrict__ keys, uint32_t n, size_t total, uint32_t ebase0,
                          uint32_t ebase_stride, uint32_t *__restrict__ cursor, uint32_t *__restrict__ entries) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const uint32_t k = keys[e];
    if (k == 0xffffffffu) return;
    const uint32_t j = (uint32_t)(e / n), i = (uint32_t)(e - (size_t)j * n);
    const uint32_t pos = atomicAdd(&cursor[k >> 1], 1u);
    entries[pos] = (ebase0 + j * ebase_stride + i) | ((k & 1u) << 31);
}

// per lane: entries and longest bucket of the segment range [bounds[l], bounds[l+1]); grid = (blocks, lanes),
// out[2l + 1] must be zero on entry
__global__ void k_lane_info(const uint32_t *__restrict__ len, const uint32_t *__restrict__ start,
                            const uint32_t *__restrict__ bounds, uint32_t *__restrict__ out) {
    const uint32_t l = blockIdx.y, s0 = bounds[l], s1 = bounds[l + 1];
    uint32_t mx = 0;
    for (uint32_t s = s0 + blockIdx.x * blockDim.x + threadIdx.x; s < s1; s += gridDim.x * blockDim.x) mx = max(mx, len[s]);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&out[2 * l + 1], mx);
    if (blockIdx.x == 0 && threadIdx.x == 0) out[2 * l] = start[s1] - start[s0];
}

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan of the bucket sizes: start[s] and a copy as scatter cursors (counting-sort
// offsets; start[nseg] = number of entries).  Three launches: tile sums, scan of the tile sums, offsets.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t block_reduce_u64(uint64_t v, uint64_t *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? sh[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    v = sh[0];
    __syncthreads();
    return v;
}

__global__ void k_scan1(const uint32_t *__restrict__ len, uint32_t nseg, uint64_t *__restrict__ blk_sum) {
    __shared__ uint64_t sh[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        s += idx < nseg ? len[idx] : 0;
    }
    s = block_reduce_u64(s, sh);
    if (threadIdx.x == 0) blk_sum[blockIdx.x] = s;
}

// single block: exclusive scan of blk_sum[0..nblk) in place; blk_sum[nblk] = total
__global__ void k_scan2(uint64_t *__restrict__ blk_sum, uint32_t nblk) {
    __shared__ uint64_t sh[SCAN_THREADS];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblk; base += SCAN_THREADS) {
        const uint32_t idx = base + threadIdx.x;
        const uint64_t v = idx < nblk ? blk_sum[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_THREADS; o <<= 1) {
            uint64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const uint64_t incl = sh[threadIdx.x];
        if (idx < nblk) blk_sum[idx] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) blk_sum[nblk] = carry;
}

__global__ void k_scan3(const uint32_t *__restrict__ len, uint32_t nseg, const uint64_t *__restrict__ blk_sum,
                        uint32_t *__restrict__ start, uint32_t *__restrict__ cursor) {
    __shared__ uint64_t sh[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t L[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        L[k] = idx < nseg ? len[idx] : 0;
        s += L[k];
    }
    // exclusive scan of the per-thread sums across the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = sh[lane];
        uint64_t wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        sh[lane] = wi - w;
    }
    __syncthreads();
    uint64_t run = blk_sum[blockIdx.x] + sh[wid] + (incl - s);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        if (idx < nseg) {
            start[idx] = (uint32_t)run;
            cursor[idx] = (uint32_t)run;
        }
        run += L[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) start[nseg] = (uint32_t)blk_sum[gridDim.x];
}

// ------------------------------------------------------------------------------------------------
// tree rounds
// ------------------------------------------------------------------------------------------------
// An entry names a point of the source list: index in bits 0..30, bit 31 = negate (round 0 only).
__device__ __forceinline__ AffPt fetch_entry(const AffPt *__restrict__ src, uint32_t e) {
    AffPt p = pt_load(&src[e & 0x7fffffffu]);
    if (e >> 31) p.y = gf_add(p.y, p.x);
    return p;
}
template <bool INDEXED>
__device__ __forceinline__ AffPt fetch_pt(const AffPt *__restrict__ src, const uint32_t *__restrict__ ent, uint32_t pos) {
    return fetch_entry(src, INDEXED ? ent[pos] : pos);
}

// ------------------------------------------------------------------------------------------------
// plan of one round, one launch: exclusive scans of (L/2, (L+1)/2) over the segment lengths, the
// segment tables of the next round, and one descriptor (a, b, out) per addition so that the two
// passes never search for their segment.  Blocks publish their aggregates and every block sums its
// predecessors' (no chain): the grid is small enough to be co-resident.
//   info[0] = max L, info[1] = additions of this round, info[2] = points after this round
// ------------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 256;

// info[3] is a ticket counter (cleared with info): a block's position in the scan is the order in which it STARTED,
// so the blocks it waits for are running or finished whatever order the hardware dispatches them in (decoupled
// look-back; no assumption of ascending dispatch or of a co-resident grid).
// The plan also carries the last point of every odd segment over to the next round (dst != nullptr).
template <bool INDEXED>
__global__ void __launch_bounds__(PLAN_THREADS)
    k_plan(const uint32_t *__restrict__ len, const uint32_t *__restrict__ in_start, const uint32_t *__restrict__ ent,
           uint32_t nseg, uint32_t items, uint64_t *__restrict__ blk_sum, uint32_t *__restrict__ blk_flag, uint32_t epoch,
           uint32_t *__restrict__ out_start, uint32_t *__restrict__ new_len, uint4 *__restrict__ desc,
           uint32_t *__restrict__ info, const AffPt *__restrict__ src, AffPt *__restrict__ dst) {
    __shared__ uint64_t sh[PLAN_THREADS / 32];
    __shared__ uint64_t sh_base;
    __shared__ uint32_t sh_vb;
    if (threadIdx.x == 0) sh_vb = atomicAdd(&info[3], 1u);
    __syncthreads();
    const uint32_t vb = sh_vb; // virtual block id
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = (vb * PLAN_THREADS + threadIdx.x) * items;
    // phase 1: aggregate of this block
    uint64_t s = 0;
    uint32_t mx = 0;
    for (uint32_t k = 0; k < items; k++) {
        const uint32_t idx = base + k;
        const uint32_t L = idx < nseg ? len[idx] : 0;
        s += (uint64_t)(L >> 1) | ((uint64_t)((L + 1) >> 1) << 32);
        mx = max(mx, L);
    }
    uint64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(&info[0], mx);
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint64_t w = lane < PLAN_THREADS / 32 ? sh[lane] : 0;
        uint64_t wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < PLAN_THREADS / 32) sh[lane] = wi - w; // exclusive offsets of the warps
        if (lane == PLAN_THREADS / 32 - 1) {
            blk_sum[vb] = wi; // block aggregate
            __threadfence();
            atomicExch(&blk_flag[vb], epoch);
        }
        // phase 2: sum of the aggregates of the blocks that started earlier
        uint64_t p = 0;
        for (uint32_t b = lane; b < vb; b += 32) {
            while (atomicAdd(&blk_flag[b], 0u) != epoch) __nanosleep(20);
            __threadfence();
            p += *reinterpret_cast<volatile uint64_t *>(&blk_sum[b]);
        }
        for (int o = 16; o > 0; o >>= 1) p += __shfl_down_sync(0xffffffffu, p, o);
        if (lane == 0) sh_base = p;
    }
    __syncthreads();
    // phase 3: outputs
    uint64_t run = sh_base + sh[wid] + (incl - s);
    for (uint32_t k = 0; k < items; k++) {
        const uint32_t idx = base + k;
        uint32_t L = 0, in = 0;
        if (idx < nseg) {
            L = len[idx];
            in = in_start[idx];
            out_start[idx] = (uint32_t)(run >> 32);
            new_len[idx] = (L + 1) >> 1;
            if (dst && (L & 1)) pt_store(&dst[(uint32_t)(run >> 32) + (L >> 1)], fetch_pt<INDEXED>(src, ent, in + L - 1));
        }
        const uint32_t nt = L >> 1, ts = (uint32_t)run, os = (uint32_t)(run >> 32);
        // descriptors, written by the whole warp for one segment at a time
        uint32_t todo = __ballot_sync(0xffffffffu, nt > 0);
        while (todo) {
            const int src_lane = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t nt_b = __shfl_sync(0xffffffffu, nt, src_lane), ts_b = __shfl_sync(0xffffffffu, ts, src_lane);
            const uint32_t in_b = __shfl_sync(0xffffffffu, in, src_lane), os_b = __shfl_sync(0xffffffffu, os, src_lane);
            for (uint32_t j = lane; j < nt_b; j += 32) {
                uint32_t a = in_b + 2 * j, b = a + 1;
                if (INDEXED) {
                    a = ent[a];
                    b = ent[b];
                }
                desc[ts_b + j] = make_uint4(a, b, os_b + j, 0);
            }
        }
        run += (uint64_t)(L >> 1) | ((uint64_t)((L + 1) >> 1) << 32);
    }
    if (vb == gridDim.x - 1 && threadIdx.x == PLAN_THREADS - 1) {
        out_start[nseg] = (uint32_t)(run >> 32);
        info[1] = (uint32_t)run;
        info[2] = (uint32_t)(run >> 32);
    }
}

// One shared copy of the 1.2k-instruction multiplier for the pass-2 loop: four inlined copies are 86 KB
// of code (I-cache misses were 10 % of the stalls) and push the kernel to 255 registers; called out of
// line the loop fits 128 registers, i.e. twice the resident warps.
__device__ __noinline__ gf gf_mul_call(const gf a, const gf b) { return gf_mul(a, b); }

// ------------------------------------------------------------------------------------------------
// table-driven inversion: x -> x^(2^k) is GF(2)-linear, so for k in {7,14,29,58,116} it is 30 byte-indexed
// lookups (one 32-byte row per input byte) xor-ed together.  Itoh-Tsujii then needs 8 single squarings,
// 5 table passes and 10 multiplications instead of 232 squarings: the inversion sits on the critical
// path of every round of the tree, so its latency (not its throughput) is what matters.
// ------------------------------------------------------------------------------------------------
constexpr int MSQ_TABLES = 5;
__device__ __constant__ int MSQ_K[MSQ_TABLES] = {7, 14, 29, 58, 116};
constexpr size_t MSQ_TABLE_ELEMS = 30 * 256; // gf rows per table

__global__ void k_build_msqr_tables(gf *__restrict__ tabs) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= MSQ_TABLES * MSQ_TABLE_ELEMS) return;
    const uint32_t t = id / MSQ_TABLE_ELEMS, r = id % MSQ_TABLE_ELEMS, pos = r >> 8, v = r & 255;
    gf x = gf_zero();
    x.v[pos >> 2] = v << (8 * (pos & 3));
    if (pos == 29) x.v[7] &= 0x1ffu; // bits >= 233 do not exist
    gf_store(&tabs[id], gf_sqr_n(x, MSQ_K[t]));
}
__device__ __forceinline__ gf gf_msqr_tab(const gf &x, const gf *__restrict__ tab) {
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // the 30 row fetches are independent: issue them in batches of 10 (20 x 16 bytes in flight) before folding, so the
    // pass costs ~3 L2 round trips instead of one per row -- this sits on the latency-critical path of every round
#pragma unroll
    for (int b0 = 0; b0 < 30; b0 += 10) {
        uint4 lo[10], hi[10];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const int pos = b0 + k;
            const uint32_t b = (x.v[pos >> 2] >> (8 * (pos & 3))) & 255u;
            const uint4 *row = reinterpret_cast<const uint4 *>(tab + pos * 256 + b);
            lo[k] = __ldg(row);
            hi[k] = __ldg(row + 1);
        }
#pragma unroll
        for (int k = 0; k < 10; k++) {
            acc[0] ^= lo[k].x; acc[1] ^= lo[k].y; acc[2] ^= lo[k].z; acc[3] ^= lo[k].w;
            acc[4] ^= hi[k].x; acc[5] ^= hi[k].y; acc[6] ^= hi[k].z; acc[7] ^= hi[k].w;
        }
    }
    gf r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = acc[i];
    return r;
}
// chain 1,2,3,6,7,14,28,29,58,116,232; 0 -> 0
__device__ __noinline__ gf gf_inv_tab(const gf &a, const gf *__restrict__ tabs) {
    const gf *t7 = tabs, *t14 = tabs + MSQ_TABLE_ELEMS, *t29 = tabs + 2 * MSQ_TABLE_ELEMS,
             *t58 = tabs + 3 * MSQ_TABLE_ELEMS, *t116 = tabs + 4 * MSQ_TABLE_ELEMS;
    const gf b2 = gf_mul(gf_sqr(a), a);
    const gf b3 = gf_mul(gf_sqr(b2), a);
    const gf b6 = gf_mul(gf_sqr(gf_sqr(gf_sqr(b3))), b3);
    const gf b7 = gf_mul(gf_sqr(b6), a);
    const gf b14 = gf_mul(gf_msqr_tab(b7, t7), b7);
    const gf b28 = gf_mul(gf_msqr_tab(b14, t14), b14);
    const gf b29 = gf_mul(gf_sqr(b28), a);
    const gf b58 = gf_mul(gf_msqr_tab(b29, t29), b29);
    const gf b116 = gf_mul(gf_msqr_tab(b58, t58), b58);
    const gf b232 = gf_mul(gf_msqr_tab(b116, t116), b116);
    return gf_sqr(b232);
}

// the same inversion shared by the 32 lanes of a warp (gf233_warp.cuh): ~6 us instead of 46
__device__ __forceinline__ gf gf_inv_tab_warp(const gf &a, const gf *__restrict__ tabs, const WarpMulCtx &c) {
    return gf_inv_warp(a, tabs, MSQ_TABLE_ELEMS, c);
}

// A tree round small enough to be pure latency: one WARP per addition, no batching -- the warp inverts its own
// denominator cooperatively and finishes the addition with cooperative multiplications (one launch, ~10 us).
__global__ void __launch_bounds__(128)
    k_round_warp(const AffPt *__restrict__ src, const uint32_t *__restrict__ info, const uint4 *__restrict__ desc,
                 AffPt *__restrict__ dst, const gf *__restrict__ tabs) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= info[1]) return; // whole warp
    const WarpMulCtx c = warp_mul_ctx();
    const uint4 de = desc[t];
    const AffPt p1 = fetch_entry(src, de.x), p2 = fetch_entry(src, de.y);
    gf d;
    const int kind = pair_classify(p1, p2, d); // warp-uniform: every lane holds the same points
    AffPt r;
    if (kind >= 2) {
        r = pair_finish(p1, p2, kind, d);
    } else {
        const gf dinv = gf_inv_tab_warp(d, tabs, c);
        gf lam = gf_mul_warp(kind == 1 ? p1.y : gf_add(p1.y, p2.y), dinv, c);
        if (kind == 1) lam = gf_add(lam, p1.x);
        r.x = gf_add(gf_add(gf_sqr(lam), lam), gf_add(p1.x, p2.x));
        r.y = gf_add(gf_add(gf_mul_warp(lam, gf_add(p1.x, r.x), c), r.x), p1.y);
    }
    if ((threadIdx.x & 31) == 0) pt_store(&dst[de.z], r);
}

// ---- operand staging for the pass-2 loops: cp.async (LDGSTS) gathers the NEXT addition's two points, its prefix
// product and the descriptor after it into the thread's own shared-memory slots while the current addition's four
// products run.  No register is live across the out-of-line multiplier calls for it (a prefetch into registers is
// spilled around the calls, and the spill store waits for the load), and with 4 warps per scheduler every warp that
// sits on a global load is a quarter of the issue slots (ncu: 12 % of the warp samples were long-scoreboard stalls).
// Two slot sets by parity: the copies of addition k-1 never land in the set addition k is read from.
constexpr int STG_CHUNKS = 11;                       // 16-byte chunks: 4 + 4 (points), 2 (prefix), 1 (descriptor)
constexpr size_t STG_BYTES = 2 * STG_CHUNKS * 256 * sizeof(uint4); // per block of 256 threads
__device__ __forceinline__ void cp_async16(uint4 *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// slot c of set b of thread tid (chunk-major: a warp's 16-byte accesses to one chunk are contiguous, conflict-free)
#define STG_SLOT(stage, b, c, tid) ((stage) + ((b) * STG_CHUNKS + (c)) * 256 + (tid))
__device__ __forceinline__ void stage_issue(uint4 *stage, int b, uint32_t tid, const AffPt *src, const uint4 &de,
                                            const gf *prefix_t, const uint4 *desc_next) {
    const uint4 *a = reinterpret_cast<const uint4 *>(&src[de.x & 0x7fffffffu]);
    const uint4 *c = reinterpret_cast<const uint4 *>(&src[de.y & 0x7fffffffu]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cp_async16(STG_SLOT(stage, b, i, tid), a + i);
        cp_async16(STG_SLOT(stage, b, 4 + i, tid), c + i);
    }
    if (prefix_t) {
        const uint4 *q = reinterpret_cast<const uint4 *>(prefix_t);
        cp_async16(STG_SLOT(stage, b, 8, tid), q);
        cp_async16(STG_SLOT(stage, b, 9, tid), q + 1);
    }
    if (desc_next) cp_async16(STG_SLOT(stage, b, 10, tid), desc_next);
    cp_async_commit();
}
// pass 1 needs the x coordinates only (slots 0-1 and 4-5)
__device__ __forceinline__ void stage_issue_x(uint4 *stage, int b, uint32_t tid, const AffPt *src, const uint4 &de,
                                              const uint4 *desc_next) {
    const uint4 *a = reinterpret_cast<const uint4 *>(&src[de.x & 0x7fffffffu].x);
    const uint4 *c = reinterpret_cast<const uint4 *>(&src[de.y & 0x7fffffffu].x);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        cp_async16(STG_SLOT(stage, b, i, tid), a + i);
        cp_async16(STG_SLOT(stage, b, 4 + i, tid), c + i);
    }
    if (desc_next) cp_async16(STG_SLOT(stage, b, 10, tid), desc_next);
    cp_async_commit();
}
__device__ __forceinline__ gf stage_gf(const uint4 *stage, int b, int c, uint32_t tid) {
    const uint4 u = *STG_SLOT(stage, b, c, tid), v = *STG_SLOT(stage, b, c + 1, tid);
    gf r;
    r.v[0] = u.x; r.v[1] = u.y; r.v[2] = u.z; r.v[3] = u.w;
    r.v[4] = v.x; r.v[5] = v.y; r.v[6] = v.z; r.v[7] = v.w;
    return r;
}
__device__ __forceinline__ void stage_points(const uint4 *stage, int b, uint32_t tid, const uint4 &de, AffPt &p1, AffPt &p2) {
    p1.x = stage_gf(stage, b, 0, tid);
    p1.y = stage_gf(stage, b, 2, tid);
    p2.x = stage_gf(stage, b, 4, tid);
    p2.y = stage_gf(stage, b, 6, tid);
    if (de.x >> 31) p1.y = gf_add(p1.y, p1.x); // entry = index | negate << 31 (fetch_entry)
    if (de.y >> 31) p2.y = gf_add(p2.y, p2.x);
}
// the additions that need no field product: kind 2 -> P1, 3 -> P2, 4 -> infinity
__device__ __forceinline__ AffPt pair_degenerate(const AffPt &p1, const AffPt &p2, int kind) {
    return kind == 2 ? p1 : kind == 3 ? p2 : pt_inf();
}

// ------------------------------------------------------------------------------------------------
// All tree rounds of a lane in ONE persistent kernel (cooperative launch: the grid is co-resident, one or two
// blocks per SM).  A round is  plan | grid barrier | pass 1 -> inversion -> pass 2 | grid barrier:  the additions of a
// round are dealt evenly to the threads (chains of ceil(tasks / threads): no wave quantisation), and every WARP inverts
// the totals of its own 32 chains in registers (butterfly product tree by shuffles, one warp-cooperative inversion at
// its root), so between the two passes nothing is exchanged and no warp waits for another.  The number of rounds is
// decided on the device (the longest segment the plan saw).  The blocks of the OTHER lane's kernel are resident on the
// same SMs: a lane's latency-bound stretches (plan, inversion, barriers, the small last rounds) are filled with the
// other lane's multiplications warp by warp.
// Everything one block writes and another reads inside the kernel is loaded with ld.global.cg (L1 is not coherent).
// ------------------------------------------------------------------------------------------------
struct AccArgs {
    const AffPt *src0;              // points named by round 0 (SRS / tables, or the buckets for a reduction level)
    const uint32_t *ent;            // round-0 index list (entry = index | negate << 31), or nullptr: positions
    const uint32_t *start0, *len0;  // round-0 segment tables (nseg + 1 / nseg entries)
    uint32_t nseg;
    AffPt *pp[2];                   // ping-pong point lists of the later rounds
    uint32_t *seg_start[2], *seg_len[2];
    uint4 *desc;
    gf *prefix, *tot, *tot_inv;
    unsigned long long *blk_sum;    // one per block
    uint32_t *ctl;                  // [0] barrier, [1] abort flag, [8 + r] longest segment at the plan of round r
    uint32_t *result;               // [0] rounds run, [1] additions of round 0, [2] additions of all rounds (low 32 bits)
    unsigned long long *times;      // profiling: globaltimer of block 0 after each phase, 6 per round (or nullptr)
    AffPt *dst;                     // finalize: dst[s] = the point of segment s (or infinity); nullptr: leave the list
    const gf *tabs;
    int max_rounds;                 // run at most this many rounds (a fixed count for the reduction levels)
    int tiny_max;                   // a round of at most this many additions runs one warp per addition
    // the plan of ALL rounds made before the launch (k_pa_*), or nullptr: every round plans itself
    const uint32_t *pl_ctl;         // see PL_* below
    uint32_t *pl_start, *pl_len;    // tables of round r (r >= 1) at + r * pl_stride
    uint32_t pl_stride;
};

// ------------------------------------------------------------------------------------------------
// The plan of all rounds at once.  Segment lengths halve (rounding up) from round to round whatever the points are, so
// the segment tables and the descriptors of EVERY round follow from the round-0 lengths alone: three small launches
// before the persistent kernel (block aggregates of the per-round (additions, points) pairs | scan over the blocks,
// round count, per-round totals | tables and descriptors of all rounds) replace the two plan phases and two of the three
// grid barriers of every round inside it; what is left of a round's plan in the kernel is the copy of the odd
// segments' last points.  Up to PL_K rounds (segments of up to 2^PL_K points); longer segments (degenerate inputs)
// leave mode 0 and the kernel plans round by round as before.
// pl_ctl: [0] mode (1: plan valid) [1] rounds [2] longest segment [4 + r] additions of round r
//         [20 + r] first descriptor of round r [40] additions of all rounds
// ------------------------------------------------------------------------------------------------
constexpr int PL_K = 12;
constexpr int PL_THREADS = 256;
constexpr uint32_t PL_MAX_BLOCKS = 1024;
constexpr int PL_CTL_WORDS = 64;

__device__ __forceinline__ void pl_thread_sums(const uint32_t *__restrict__ len0, uint32_t nseg, uint32_t base, uint32_t items,
                                               unsigned long long s[PL_K], uint32_t &mx) {
#pragma unroll
    for (int r = 0; r < PL_K; r++) s[r] = 0;
    mx = 0;
    for (uint32_t k = 0; k < items; k++) {
        const uint32_t idx = base + k;
        uint32_t L = idx < nseg ? len0[idx] : 0;
        mx = max(mx, L);
#pragma unroll
        for (int r = 0; r < PL_K; r++) {
            s[r] += (unsigned long long)(L >> 1) | ((unsigned long long)((L + 1) >> 1) << 32);
            L = (L + 1) >> 1;
        }
    }
}

__global__ void __launch_bounds__(PL_THREADS)
    k_pa_aggr(const uint32_t *__restrict__ len0, uint32_t nseg, uint32_t items, unsigned long long *__restrict__ blk_sum,
              uint32_t *__restrict__ pl_ctl) {
    __shared__ unsigned long long sh[PL_K][PL_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long s[PL_K];
    uint32_t mx;
    pl_thread_sums(len0, nseg, (blockIdx.x * PL_THREADS + threadIdx.x) * items, items, s, mx);
#pragma unroll
    for (int r = 0; r < PL_K; r++) {
        unsigned long long v = s[r];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[r][wid] = v;
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(&pl_ctl[2], mx);
    __syncthreads();
    if (threadIdx.x < PL_K) {
        unsigned long long v = 0;
        for (int w = 0; w < PL_THREADS / 32; w++) v += sh[threadIdx.x][w];
        blk_sum[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v; // [round][block]
    }
}

// one block: blk_sum[r][b] -> exclusive prefix over b (in place), totals, round count, descriptor offsets.
// One warp per round (the rounds' scans are independent), 32 blocks per step with a running carry (coalesced).
__global__ void __launch_bounds__(PL_THREADS)
    k_pa_scan(unsigned long long *__restrict__ blk_sum, uint32_t nblk, int max_rounds, uint32_t *__restrict__ pl_ctl) {
    __shared__ unsigned long long tot[PL_K];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = (int)wid; r < PL_K; r += PL_THREADS / 32) {
        unsigned long long *row = blk_sum + (size_t)r * nblk;
        unsigned long long carry = 0;
        for (uint32_t b0 = 0; b0 < nblk; b0 += 32) {
            const uint32_t b = b0 + lane;
            const unsigned long long v = b < nblk ? row[b] : 0;
            unsigned long long incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            if (b < nblk) row[b] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) tot[r] = carry;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t maxlen = pl_ctl[2];
        int R = 0;
        while (R < 32 && (1u << R) < maxlen) R++;
        const uint32_t mode = R <= PL_K ? 1u : 0u;
        R = min(R, max_rounds);
        uint32_t off = 0, all = 0;
        for (int r = 0; r < PL_K; r++) {
            const uint32_t nt = (uint32_t)tot[r];
            pl_ctl[4 + r] = nt;
            pl_ctl[20 + r] = off;
            if (r < R) off += nt, all += nt;
        }
        pl_ctl[40] = all;
        pl_ctl[1] = mode ? (uint32_t)R : 0u;
        pl_ctl[0] = mode;
    }
}

template <bool INDEXED>
__global__ void __launch_bounds__(PL_THREADS)
    k_pa_emit(const uint32_t *__restrict__ len0, const uint32_t *__restrict__ start0, const uint32_t *__restrict__ ent,
              uint32_t nseg, uint32_t items, const unsigned long long *__restrict__ blk_off,
              const uint32_t *__restrict__ pl_ctl, uint32_t *__restrict__ pl_start, uint32_t *__restrict__ pl_len,
              uint32_t stride, uint4 *__restrict__ desc) {
    __shared__ unsigned long long sh[PL_K][PL_THREADS / 32];
    if (pl_ctl[0] != 1u) return;
    const int R = (int)pl_ctl[1];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = (blockIdx.x * PL_THREADS + threadIdx.x) * items;
    unsigned long long run[PL_K];
    {
        unsigned long long s[PL_K];
        uint32_t mx;
        pl_thread_sums(len0, nseg, base, items, s, mx);
#pragma unroll
        for (int r = 0; r < PL_K; r++) {
            unsigned long long incl = s[r];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            if (lane == 31) sh[r][wid] = incl;
            run[r] = incl - s[r]; // exclusive inside the warp
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PL_K; r++) {
            unsigned long long w = 0;
            for (uint32_t q = 0; q < wid; q++) w += sh[r][q];
            run[r] += w + blk_off[(size_t)r * gridDim.x + blockIdx.x];
        }
    }
    // Tables by every thread for its own segments; descriptors by the whole warp over the flat list of its 32 segments'
    // additions in a round (a lane finds the segment of its position by a 5-step search over the lanes' offsets): the
    // stores are coalesced and every iteration is independent of the others.  (A warp working through one segment at a
    // time waited for the index loads of round 0: 160 us at 2^20 points; a thread per segment issued 32 partial-sector
    // stores per instruction: 180 us.)
    for (uint32_t k = 0; k < items; k++) {
        const uint32_t idx = base + k;
        uint32_t L = 0, in_r = 0;
        if (idx < nseg) {
            L = len0[idx];
            in_r = start0[idx];
        }
#pragma unroll
        for (int r = 0; r < PL_K; r++) {
            if (r < R) { // (R is uniform)
                const uint32_t nt = L >> 1, ts = pl_ctl[20 + r] + (uint32_t)run[r], os = (uint32_t)(run[r] >> 32);
                if (idx < nseg) {
                    pl_start[(size_t)(r + 1) * stride + idx] = os;
                    pl_len[(size_t)(r + 1) * stride + idx] = (L + 1) >> 1;
                }
                uint32_t incl = nt;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += v;
                }
                const uint32_t excl = incl - nt, total = __shfl_sync(0xffffffffu, incl, 31);
                for (uint32_t f0 = 0; f0 < total; f0 += 32) {
                    const uint32_t f = f0 + lane;
                    uint32_t sg = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t e = __shfl_sync(0xffffffffu, excl, sg + step);
                        if (e <= f) sg += step;
                    }
                    const uint32_t j = f - __shfl_sync(0xffffffffu, excl, sg);
                    const uint32_t in_b = __shfl_sync(0xffffffffu, in_r, sg), os_b = __shfl_sync(0xffffffffu, os, sg);
                    const uint32_t ts_b = __shfl_sync(0xffffffffu, ts, sg);
                    if (f < total) {
                        uint32_t a = in_b + 2 * j, b = a + 1;
                        if (INDEXED && r == 0) {
                            a = ent[a];
                            b = ent[b];
                        }
                        desc[ts_b + j] = make_uint4(a, b, os_b + j, 0);
                    }
                }
                run[r] += (unsigned long long)nt | ((unsigned long long)((L + 1) >> 1) << 32);
                L = (L + 1) >> 1;
                in_r = os;
            }
        }
    }
}
#ifndef ACC_THREADS_N
#define ACC_THREADS_N 256
#endif
constexpr int ACC_THREADS = ACC_THREADS_N; // <= 256 (the staging slots are laid out for 256 threads)
#ifndef ACC_MINB
#define ACC_MINB 1
#endif
constexpr int ACC_MAX_ROUNDS = 48;
constexpr int ACC_TIME_WORDS = 6 * (ACC_MAX_ROUNDS + 1) + 2; // [last] = kernel start

__device__ __forceinline__ gf gf_load_cg(const gf *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldcg(q), b = __ldcg(q + 1);
    gf r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ AffPt fetch_entry_cg(const AffPt *src, uint32_t e) {
    const AffPt *q = &src[e & 0x7fffffffu];
    AffPt p;
    p.x = gf_load_cg(&q->x);
    p.y = gf_load_cg(&q->y);
    if (e >> 31) p.y = gf_add(p.y, p.x);
    return p;
}
__device__ __forceinline__ unsigned long long acc_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// sense-free grid barrier on a counter that only grows; gives up (abort flag) instead of spinning for ever
__device__ __forceinline__ bool acc_barrier(uint32_t *ctl, uint32_t &target) {
    __shared__ uint32_t sh_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&ctl[0], 1u);
        const unsigned long long t0 = acc_now();
        uint32_t ab = 0;
        while (atomicAdd(&ctl[0], 0u) < target) {
            __nanosleep(40);
            if (atomicAdd(&ctl[1], 0u) != 0u || acc_now() - t0 > 4000000000ull) { // 4 s: a lost block, not a slow one
                atomicExch(&ctl[1], 1u);
                ab = 1;
                break;
            }
        }
        __threadfence();
        sh_abort = ab;
    }
    __syncthreads();
    return sh_abort == 0;
}

// ---- the three steps of a warp's share of a round (32 chains of B additions), shared by the persistent kernel and the
// fused per-round kernel of the separate-launch path.  Lanes without a task keep the total 1.
// pass 1: denominators and per-thread prefix products; operands staged one addition ahead (a prefetch into registers
// would be spilled around the out-of-line product, and the spill store waits for the load)
__device__ __forceinline__ gf chain_pass1(uint4 *stage, uint32_t tid, uint32_t lane, uint32_t base, uint32_t B, uint32_t ntasks,
                                          const AffPt *cur, const uint4 *desc, gf *prefix) {
    gf acc = gf_one();
    if (base + lane >= ntasks) return acc;
    uint32_t t = base + lane;
    uint4 de = __ldcg(&desc[t]);
    stage_issue_x(stage, 0, tid, cur, de, (B > 1 && t + 32 < ntasks) ? &desc[t + 32] : nullptr);
#pragma unroll 1
    for (uint32_t k = 0; k < B && t < ntasks; k++, t += 32) {
        cp_async_wait_all();
        const gf x1 = stage_gf(stage, k & 1, 0, tid), x2 = stage_gf(stage, k & 1, 4, tid);
        const uint4 de0 = de;
        if (k + 1 < B && t + 32 < ntasks) {
            de = *STG_SLOT(stage, k & 1, 10, tid);
            stage_issue_x(stage, (k + 1) & 1, tid, cur, de, (k + 2 < B && t + 64 < ntasks) ? &desc[t + 64] : nullptr);
        }
        gf d = gf_add(x1, x2);
        if (gf_is_zero(x1) | gf_is_zero(x2)) d = gf_one();
        else if (gf_is_zero(d)) {
            const AffPt p1 = fetch_entry_cg(cur, de0.x), p2 = fetch_entry_cg(cur, de0.y);
            d = gf_eq(p1.y, p2.y) ? x1 : gf_one();
        }
        if (B > 1) {
            gf_store(&prefix[t], acc);
            acc = gf_mul_call(acc, d);
        } else {
            acc = d; // a chain of one: its prefix is 1
        }
    }
    return acc;
}
// The warp inverts its own 32 thread totals, in registers: five butterfly levels up (every lane keeps its sibling's
// sub-product), one cooperative inversion of the warp product, five levels back down.  No inversion leaves the warp, so
// the warps of a round never wait for each other between the two passes.  All 32 lanes must call.
__device__ __forceinline__ gf warp_invert_totals(const gf &acc, const gf *__restrict__ tabs) {
    gf S[5], P = acc;
#pragma unroll
    for (int k = 0; k < 5; k++) {
#pragma unroll
        for (int q = 0; q < 8; q++) S[k].v[q] = __shfl_xor_sync(0xffffffffu, P.v[q], 1 << k);
        P = gf_mul_call(P, S[k]);
    }
    const WarpMulCtx wc = warp_mul_ctx();
    gf inv = gf_inv_tab_warp(P, tabs, wc);
#pragma unroll
    for (int k = 4; k >= 0; k--) inv = gf_mul_call(inv, S[k]);
    return inv;
}
// pass 2: the same additions backwards with the inverse of the thread's total, operands staged one addition ahead
__device__ __forceinline__ void chain_pass2(uint4 *stage, uint32_t tid, uint32_t lane, uint32_t base, uint32_t B, uint32_t ntasks,
                                            const AffPt *cur, const uint4 *desc, const gf *prefix, gf inv, AffPt *out) {
    if (base + lane >= ntasks) return;
    int k = (int)min(B - 1, (ntasks - 1 - base - lane) >> 5); // the thread's last task
    uint32_t t = base + (uint32_t)k * 32 + lane;
    uint4 de = __ldcg(&desc[t]);
    stage_issue(stage, k & 1, tid, cur, de, B > 1 ? &prefix[t] : nullptr, k > 0 ? &desc[t - 32] : nullptr);
#pragma unroll 1
    for (; k >= 0; k--, t -= 32) {
        cp_async_wait_all();
        AffPt p1, p2;
        stage_points(stage, k & 1, tid, de, p1, p2);
        gf dinv = inv;
        if (B > 1) dinv = stage_gf(stage, k & 1, 8, tid);
        const uint32_t dz = de.z;
        if (k > 0) {
            de = *STG_SLOT(stage, k & 1, 10, tid);
            stage_issue(stage, (k - 1) & 1, tid, cur, de, &prefix[t - 32], k > 1 ? &desc[t - 64] : nullptr);
        }
        gf d;
        const int kind = pair_classify(p1, p2, d);
        if (B > 1) {
            dinv = gf_mul_call(inv, dinv);
            if (k) inv = gf_mul_call(inv, d);
        }
        AffPt q;
        if (kind >= 2) {
            q = pair_degenerate(p1, p2, kind);
        } else {
            // chord / tangent: lambda = num/d (+ x1 for the tangent), see pair_finish
            gf lam = gf_mul_call(kind == 1 ? p1.y : gf_add(p1.y, p2.y), dinv);
            if (kind == 1) lam = gf_add(lam, p1.x);
            q.x = gf_add(gf_add(gf_sqr(lam), lam), gf_add(p1.x, p2.x));
            q.y = gf_add(gf_add(gf_mul_call(lam, gf_add(p1.x, q.x)), q.x), p1.y);
        }
        pt_store(&out[dz], q);
    }
}
// One round of the separate-launch path in ONE kernel: every warp runs pass 1, inverts its own 32 totals and runs pass 2
// (no thr_total / thr_inv arrays, no hierarchical inversion launches between the passes).
template <int B>
__global__ void __launch_bounds__(256, 1)
    k_round_fused(const AffPt *__restrict__ src, const uint32_t *__restrict__ info, const uint4 *__restrict__ desc,
                  gf *__restrict__ prefix, AffPt *__restrict__ dst, const gf *__restrict__ tabs) {
    extern __shared__ uint4 stage[];
    const uint32_t ntasks = info[1];
    const uint32_t tid = threadIdx.x, gtid = blockIdx.x * blockDim.x + tid;
    const uint32_t lane = gtid & 31, warp = gtid >> 5;
    const uint32_t base = warp * (32u * B);
    if (base >= ntasks) return; // the whole warp
    const gf acc = chain_pass1(stage, tid, lane, base, B, ntasks, src, desc, prefix);
    const gf inv = warp_invert_totals(acc, tabs);
    chain_pass2(stage, tid, lane, base, B, ntasks, src, desc, prefix, inv, dst);
}

__global__ void __launch_bounds__(ACC_THREADS, ACC_MINB) k_accumulate(const AccArgs A) {
    extern __shared__ uint4 stage[]; // STG_BYTES: the pass-2 operand slots
    __shared__ unsigned long long sh[ACC_THREADS / 32];
    __shared__ unsigned long long sh_base, sh_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nthreads = gridDim.x * ACC_THREADS, nwarps = nthreads >> 5;
    const uint32_t gtid = blockIdx.x * ACC_THREADS + tid, gw = gtid >> 5;
    uint32_t bar = 0;
    const AffPt *cur = A.src0;
    const uint32_t *cur_ent = A.ent, *in_start = A.start0, *in_len = A.len0;
    const uint32_t nseg = A.nseg;
    const uint32_t items = (nseg + nthreads - 1) / nthreads;
    const uint32_t seg_base = gtid * items;
    uint32_t adds_r0 = 0, adds_all = 0;
    int r = 0;
    const bool prof = A.times != nullptr && blockIdx.x == 0 && tid == 0;
    if (prof) A.times[ACC_TIME_WORDS - 1] = acc_now();
    // planned ahead?  (the same answer in every thread of the grid; -1: every round plans itself)
    const int pre_rounds = (A.pl_ctl != nullptr && __ldcg(&A.pl_ctl[0]) == 1u) ? (int)__ldcg(&A.pl_ctl[1]) : -1;
    for (; r < A.max_rounds; r++) {
        const int o = (r + 1) & 1;
        uint32_t *out_start = A.seg_start[o], *new_len = A.seg_len[o];
        AffPt *out = A.pp[r & 1];
        uint32_t doff = 0; // this round's descriptors start at A.desc + doff
        uint32_t ntasks = 0;
        if (pre_rounds >= 0) {
            // ---- planned before the launch: this round's tables and descriptors exist; carry the odd segments' last
            // points over (the passes of this round write other positions of `out`, the next round reads them)
            if (r >= pre_rounds) break;
            ntasks = __ldcg(&A.pl_ctl[4 + r]);
            doff = __ldcg(&A.pl_ctl[20 + r]);
            out_start = A.pl_start + (size_t)(r + 1) * A.pl_stride;
            new_len = A.pl_len + (size_t)(r + 1) * A.pl_stride;
            for (uint32_t k = 0; k < items; k++) {
                const uint32_t idx = seg_base + k;
                if (idx >= nseg) break;
                const uint32_t L = __ldcg(&in_len[idx]);
                if (L & 1) {
                    const uint32_t in = __ldcg(&in_start[idx]), os = __ldcg(&out_start[idx]);
                    const uint32_t e = cur_ent ? cur_ent[in + L - 1] : in + L - 1;
                    pt_store(&out[os + (L >> 1)], fetch_entry_cg(cur, e));
                }
            }
        } else {
        // ---- plan, step 1: block aggregates of (additions, points after the round) and the longest segment
        unsigned long long s = 0;
        uint32_t mx = 0;
        for (uint32_t k = 0; k < items; k++) {
            const uint32_t idx = seg_base + k;
            const uint32_t L = idx < nseg ? __ldcg(&in_len[idx]) : 0;
            s += (unsigned long long)(L >> 1) | ((unsigned long long)((L + 1) >> 1) << 32);
            mx = max(mx, L);
        }
        unsigned long long incl = s;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
        if (lane == 0 && mx) atomicMax(&A.ctl[8 + r], mx);
        if (lane == 31) sh[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = lane < ACC_THREADS / 32 ? sh[lane] : 0;
            unsigned long long wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += t;
            }
            if (lane < ACC_THREADS / 32) sh[lane] = wi - w; // exclusive offsets of the warps
            if (lane == ACC_THREADS / 32 - 1) A.blk_sum[blockIdx.x] = wi;
        }
        if (!acc_barrier(A.ctl, bar)) return;
        if (prof) A.times[6 * r + 0] = acc_now();
        const uint32_t maxlen = __ldcg(&A.ctl[8 + r]);
        if (maxlen <= 1) break; // every segment holds at most one point
        // ---- plan, step 2: this block's offset and the totals
        if (wid == 0) {
            unsigned long long p = 0, tot = 0;
            for (uint32_t b = lane; b < gridDim.x; b += 32) {
                const unsigned long long v = __ldcg(&A.blk_sum[b]);
                tot += v;
                if (b < blockIdx.x) p += v;
            }
            for (int d = 16; d > 0; d >>= 1) {
                p += __shfl_down_sync(0xffffffffu, p, d);
                tot += __shfl_down_sync(0xffffffffu, tot, d);
            }
            if (lane == 0) {
                sh_base = p;
                sh_total = tot;
            }
        }
        __syncthreads();
        ntasks = (uint32_t)sh_total;
        // ---- plan, step 3: next segment tables, one descriptor per addition, odd leftovers carried over
        {
            unsigned long long run = sh_base + sh[wid] + (incl - s);
            for (uint32_t k = 0; k < items; k++) {
                const uint32_t idx = seg_base + k;
                uint32_t L = 0, in = 0;
                if (idx < nseg) {
                    L = __ldcg(&in_len[idx]);
                    in = __ldcg(&in_start[idx]);
                    out_start[idx] = (uint32_t)(run >> 32);
                    new_len[idx] = (L + 1) >> 1;
                    if (L & 1) {
                        const uint32_t e = cur_ent ? cur_ent[in + L - 1] : in + L - 1;
                        pt_store(&out[(uint32_t)(run >> 32) + (L >> 1)], fetch_entry_cg(cur, e));
                    }
                }
                const uint32_t nt = L >> 1, ts = (uint32_t)run, os = (uint32_t)(run >> 32);
                uint32_t todo = __ballot_sync(0xffffffffu, nt > 0);
                while (todo) {
                    const int sl = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t nt_b = __shfl_sync(0xffffffffu, nt, sl), ts_b = __shfl_sync(0xffffffffu, ts, sl);
                    const uint32_t in_b = __shfl_sync(0xffffffffu, in, sl), os_b = __shfl_sync(0xffffffffu, os, sl);
                    for (uint32_t j = lane; j < nt_b; j += 32) {
                        uint32_t a = in_b + 2 * j, b = a + 1;
                        if (cur_ent) {
                            a = cur_ent[a];
                            b = cur_ent[b];
                        }
                        A.desc[ts_b + j] = make_uint4(a, b, os_b + j, 0);
                    }
                }
                run += (unsigned long long)(L >> 1) | ((unsigned long long)((L + 1) >> 1) << 32);
            }
        }
            if (!acc_barrier(A.ctl, bar)) return;
            if (prof) A.times[6 * r + 1] = acc_now();
        }
        if (r == 0) adds_r0 = ntasks;
        adds_all += ntasks;
        // ---- pass 1: every thread chains B additions; warp gw owns the additions [gw 32 B, (gw + 1) 32 B)
        // chain length: the additions are dealt evenly to all threads (longer chains on fewer warps were measured
        // 30-100 % slower: an addition of pass 2 is ~30 us of dependent loads and products at low occupancy)
        const uint32_t B = (ntasks + nthreads - 1) / nthreads;
        const uint32_t base = gw * 32u * B;
        if (ntasks <= (uint32_t)A.tiny_max) {
            // ---- a tiny round: one WARP per addition, everything cooperative, no batching
            const WarpMulCtx wc = warp_mul_ctx();
            for (uint32_t t = gw; t < ntasks; t += nwarps) {
                const uint4 de = __ldcg(&A.desc[doff + t]);
                const AffPt p1 = fetch_entry_cg(cur, de.x), p2 = fetch_entry_cg(cur, de.y);
                gf d;
                const int kind = pair_classify(p1, p2, d); // warp-uniform
                AffPt q;
                if (kind >= 2) {
                    q = pair_finish(p1, p2, kind, d);
                } else {
                    const gf dinv = gf_inv_tab_warp(d, A.tabs, wc);
                    gf lam = gf_mul_warp(kind == 1 ? p1.y : gf_add(p1.y, p2.y), dinv, wc);
                    if (kind == 1) lam = gf_add(lam, p1.x);
                    q.x = gf_add(gf_add(gf_sqr(lam), lam), gf_add(p1.x, p2.x));
                    q.y = gf_add(gf_add(gf_mul_warp(lam, gf_add(p1.x, q.x), wc), q.x), p1.y);
                }
                if (lane == 0) pt_store(&out[de.z], q);
            }
        } else if (base < ntasks) { // the whole warp or none of it
            const gf acc = chain_pass1(stage, tid, lane, base, B, ntasks, cur, A.desc + doff, A.prefix);
            if (prof) A.times[6 * r + 2] = acc_now();
            const gf inv = warp_invert_totals(acc, A.tabs);
            if (prof) A.times[6 * r + 3] = acc_now();
            chain_pass2(stage, tid, lane, base, B, ntasks, cur, A.desc + doff, A.prefix, inv, out);
        }
        if (!acc_barrier(A.ctl, bar)) return;
        if (prof) A.times[6 * r + 4] = acc_now();
        cur = out;
        cur_ent = nullptr;
        in_start = out_start;
        in_len = new_len;
    }
    // ---- after the last round every segment holds 0 or 1 points
    if (A.dst) {
        for (uint32_t sgm = gtid; sgm < nseg; sgm += nthreads) {
            AffPt p = pt_inf();
            if (__ldcg(&in_len[sgm])) {
                const uint32_t pos = __ldcg(&in_start[sgm]);
                p = fetch_entry_cg(cur, cur_ent ? cur_ent[pos] : pos);
            }
            pt_store(&A.dst[sgm], p);
        }
    }
    if (gtid == 0) {
        A.result[0] = (uint32_t)r;
        A.result[1] = adds_r0;
        A.result[2] = adds_all;
    }
    if (prof) A.times[6 * r + 5] = acc_now();
}

// pass 1: per task form the denominator and chain a per-thread prefix product
template <int B>
__global__ void __launch_bounds__(256)
    k_pass1(const AffPt *__restrict__ src, const uint32_t *__restrict__ info, const uint4 *__restrict__ desc,
            gf *__restrict__ prefix, gf *__restrict__ thr_total) {
    const uint32_t ntasks = info[1];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = gtid & 31, warp = gtid >> 5;
    const uint32_t base = warp * (32u * B);
    if (base >= ntasks) return;
    gf acc = gf_one();
    // the next task's descriptor and x coordinates are in flight while the current product is computed
    uint32_t t = base + lane;
    uint4 de = make_uint4(0, 0, 0, 0);
    gf x1 = gf_zero(), x2 = gf_zero();
    if (t < ntasks) {
        de = desc[t];
        x1 = gf_load(&src[de.x & 0x7fffffffu].x);
        x2 = gf_load(&src[de.y & 0x7fffffffu].x);
    }
#pragma unroll 1
    for (int k = 0; k < B; k++) {
        if (t >= ntasks) break;
        const uint32_t tn = t + 32;
        uint4 den = make_uint4(0, 0, 0, 0);
        gf x1n = gf_zero(), x2n = gf_zero();
        if (k + 1 < B && tn < ntasks) {
            den = desc[tn];
            x1n = gf_load(&src[den.x & 0x7fffffffu].x);
            x2n = gf_load(&src[den.y & 0x7fffffffu].x);
        }
        gf d = gf_add(x1, x2);
        if (gf_is_zero(x1) | gf_is_zero(x2)) d = gf_one();
        else if (gf_is_zero(d)) {
            const AffPt p1 = fetch_entry(src, de.x), p2 = fetch_entry(src, de.y);
            d = gf_eq(p1.y, p2.y) ? x1 : gf_one();
        }
        gf_store(&prefix[t], acc);
        acc = gf_mul(acc, d);
        t = tn;
        de = den;
        x1 = x1n;
        x2 = x2n;
    }
    gf_store(&thr_total[gtid], acc);
}

// pass 2: walk the same tasks backwards with the inverse of the thread total, finish the additions
template <int B, int MINB>
__global__ void __launch_bounds__(256, MINB)
    k_pass2(const AffPt *__restrict__ src, const uint32_t *__restrict__ info, const uint4 *__restrict__ desc,
            const gf *__restrict__ prefix, const gf *__restrict__ thr_inv, AffPt *__restrict__ dst) {
    extern __shared__ uint4 stage[];
    const uint32_t ntasks = info[1];
    const uint32_t tid = threadIdx.x, gtid = blockIdx.x * blockDim.x + tid;
    const uint32_t lane = gtid & 31, warp = gtid >> 5;
    const uint32_t base = warp * (32u * B);
    if (base + lane >= ntasks) return;
    gf inv = gf_load(&thr_inv[gtid]);
    int k = (int)min((uint32_t)(B - 1), (ntasks - 1 - base - lane) >> 5); // the thread's last task
    uint32_t t = base + (uint32_t)k * 32 + lane;
    uint4 de = desc[t];
    stage_issue(stage, k & 1, tid, src, de, &prefix[t], k > 0 ? &desc[t - 32] : nullptr);
#pragma unroll 1
    for (; k >= 0; k--, t -= 32) {
        cp_async_wait_all();
        AffPt p1, p2;
        stage_points(stage, k & 1, tid, de, p1, p2);
        const gf pre = stage_gf(stage, k & 1, 8, tid);
        const uint32_t dz = de.z;
        if (k > 0) { // addition k-1 into the other set; its descriptor came with addition k
            de = *STG_SLOT(stage, k & 1, 10, tid);
            stage_issue(stage, (k - 1) & 1, tid, src, de, &prefix[t - 32], k > 1 ? &desc[t - 64] : nullptr);
        }
        gf d;
        const int kind = pair_classify(p1, p2, d);
        const gf dinv = gf_mul_call(inv, pre);
        if (k) inv = gf_mul_call(inv, d);
        AffPt r;
        if (kind >= 2) {
            r = pair_degenerate(p1, p2, kind);
        } else {
            // chord / tangent: lambda = num/d (+ x1 for the tangent), see pair_finish
            gf lam = gf_mul_call(kind == 1 ? p1.y : gf_add(p1.y, p2.y), dinv);
            if (kind == 1) lam = gf_add(lam, p1.x);
            r.x = gf_add(gf_add(gf_sqr(lam), lam), gf_add(p1.x, p2.x));
            r.y = gf_add(gf_add(gf_mul_call(lam, gf_add(p1.x, r.x)), r.x), p1.y);
        }
        pt_store(&dst[dz], r);
    }
}

// after the last round every segment holds 0 or 1 points
template <bool INDEXED>
__global__ void k_finalize(const AffPt *__restrict__ src, const uint32_t *__restrict__ ent,
                           const uint32_t *__restrict__ in_start, const uint32_t *__restrict__ len, uint32_t nseg,
                           AffPt *__restrict__ dst) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    AffPt p = pt_inf();
    if (len[s]) p = fetch_pt<INDEXED>(src, ent, in_start[s]);
    pt_store(&dst[s], p);
}

// single-warp latency probe: iters dependent operations (mode 0 Itoh-Tsujii by squarings, 1 table-driven, 2 gf_mul,
// 3 warp-cooperative gf_mul, 4 warp-cooperative table-driven inversion)
__global__ void k_latency_probe(int mode, int iters, const gf *__restrict__ tabs, uint32_t *__restrict__ sink) {
    gf a;
    // modes 3, 4 are the warp-cooperative forms: their operands are warp-uniform
    const uint32_t seed = mode >= 3 ? 7u : threadIdx.x;
    for (int k = 0; k < 8; k++) a.v[k] = seed * 2654435761u + k * 40503u + 1;
    a.v[7] &= 0x1ff;
    gf b = a;
    const WarpMulCtx wc = warp_mul_ctx();
    for (int i = 0; i < iters; i++) {
        if (mode == 0) a = gf_inv(a);
        else if (mode == 1) a = gf_inv_tab(a, tabs);
        else if (mode == 2) a = gf_mul(a, b);
        else if (mode == 3) a = gf_mul_warp(a, b, wc);
        else a = gf_inv_tab_warp(a, tabs, wc);
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= a.v[k];
    if (s == 0x12345678u) sink[0] = s;
}

// self-test of the warp-cooperative primitives: one warp per element; op 0 a*b, 1 1/a
__global__ void k_selftest_warp(int op, const gf *__restrict__ a, const gf *__restrict__ b, gf *__restrict__ out, uint32_t n,
                                const gf *__restrict__ tabs) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const WarpMulCtx c = warp_mul_ctx();
    const gf x = gf_load(&a[i]);
    const gf r = op == 0 ? gf_mul_warp(x, gf_load(&b[i]), c) : gf_inv_tab_warp(x, tabs, c);
    if ((threadIdx.x & 31) == 0) gf_store(&out[i], r);
}
int selftest_warp(MsmEngine &E, int op, const void *d_a, const void *d_b, void *d_out, size_t n) {
    k_selftest_warp<<<cdiv(n, 4), 128, 0, E.stream>>>(op, (const gf *)d_a, (const gf *)d_b, (gf *)d_out, (uint32_t)n,
                                                      E.msqr_tabs.as<gf>());
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// hierarchical batched inversion of n non-zero field elements
// ------------------------------------------------------------------------------------------------
// The number of live elements is only known on the device: n0 = whole warps of pass-1 threads that own
// at least one task (info[1] tasks, `unit` = 32 B tasks per warp), then divided by the fan-ins above.
__device__ __forceinline__ uint32_t binv_count(const uint32_t *__restrict__ info, uint32_t unit, uint32_t div1,
                                               uint32_t div2) {
    uint32_t n = ((info[1] + unit - 1) / unit) * 32u;
    n = (n + div1 - 1) / div1;
    return (n + div2 - 1) / div2;
}
// Batched inversion in ONE launch: a warp owns G <= 32 consecutive elements and does the whole Montgomery trick on them
// cooperatively (gf233_warp.cuh) -- prefix products, the inverse of its total, the walk back.  Lane j keeps element j,
// prefix j and result j (the operands of a cooperative product are warp-uniform, so they are broadcast by shuffles).
// 3 cooperative products per element and one cooperative inversion per warp: (3 G - 1) 0.36 us + 6.7 us of latency
// whatever the batch size, against ~70 us per level of the thread-per-group kernels below.
__global__ void __launch_bounds__(128)
    k_binv_coop(const gf *__restrict__ in, gf *__restrict__ out, const uint32_t *__restrict__ info, uint32_t unit,
                uint32_t div1, uint32_t div2, uint32_t G, const gf *__restrict__ tabs) {
    const uint32_t n = binv_count(info, unit, div1, div2);
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t lo = w * G;
    if (lo >= n) return; // whole warp
    const int cnt = (int)min(G, n - lo);
    const WarpMulCtx c = warp_mul_ctx();
    gf mine = gf_one(), mypre = gf_one(), myout = gf_zero();
    if ((int)lane < cnt) mine = gf_load(&in[lo + lane]);
    gf acc = gf_bcast(mine, 0);
#pragma unroll 1
    for (int j = 1; j < cnt; j++) {
        if ((int)lane == j) mypre = acc;
        acc = gf_mul_warp(acc, gf_bcast(mine, j), c);
    }
    gf inv = gf_inv_tab_warp(acc, tabs, c);
#pragma unroll 1
    for (int j = cnt - 1; j > 0; j--) {
        const gf o = gf_mul_warp(inv, gf_bcast(mypre, j), c);
        if ((int)lane == j) myout = o;
        inv = gf_mul_warp(inv, gf_bcast(mine, j), c);
    }
    if (lane == 0) myout = inv;
    if ((int)lane < cnt) gf_store(&out[lo + lane], myout);
}
__global__ void k_binv_up(const gf *__restrict__ in, const uint32_t *__restrict__ info, uint32_t unit, uint32_t div1,
                          uint32_t div2, uint32_t G, gf *__restrict__ pre, gf *__restrict__ tot) {
    const uint32_t n = binv_count(info, unit, div1, div2);
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * G;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + G);
    gf acc = gf_one();
    for (uint32_t i = lo; i < hi; i++) {
        gf_store(&pre[i], acc);
        acc = gf_mul(acc, gf_load(&in[i]));
    }
    gf_store(&tot[g], acc);
}
__global__ void k_binv_down(const gf *__restrict__ in, const uint32_t *__restrict__ info, uint32_t unit, uint32_t div1,
                            uint32_t div2, uint32_t G, const gf *__restrict__ pre, const gf *__restrict__ tot_inv,
                            gf *__restrict__ out) {
    const uint32_t n = binv_count(info, unit, div1, div2);
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * G;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + G);
    gf inv = gf_load(&tot_inv[g]);
    for (uint32_t i = hi; i-- > lo;) {
        const gf v = gf_load(&in[i]);
        gf_store(&out[i], gf_mul(inv, gf_load(&pre[i])));
        if (i > lo) inv = gf_mul(inv, v);
    }
}

// ------------------------------------------------------------------------------------------------
// index lists for the two reduction levels (all segments have equal, power-of-two length)
// ------------------------------------------------------------------------------------------------
// level A: window w has nb = R*m buckets b = hi*m + lo.  Entries: nb row-major then nb column-major.
__global__ void k_gen_level_a(uint32_t W, uint32_t nb, uint32_t lm, uint32_t *__restrict__ ent) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= W * 2 * nb) return;
    const uint32_t w = e / (2 * nb), r = e % (2 * nb);
    const uint32_t m = 1u << lm, R = nb >> lm;
    uint32_t b;
    if (r < nb) b = r;
    else {
        const uint32_t q = r - nb, lo = q / R, hi = q % R;
        b = hi * m + lo;
    }
    ent[e] = w * nb + b;
}
// level B: per window lm column-bit subsets (m/2 each), lr row-bit subsets (R/2 each), then all m columns.
// Source layout rc[w*(R+m) + hi] (rows) and rc[w*(R+m) + R + lo] (columns).
__global__ void k_gen_level_b(uint32_t W, uint32_t lr, uint32_t lm, uint32_t *__restrict__ ent) {
    const uint32_t R = 1u << lr, m = 1u << lm;
    const uint32_t per = lm * (m >> 1) + lr * (R >> 1) + m;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= W * per) return;
    const uint32_t w = e / per;
    uint32_t r = e % per;
    const uint32_t basew = w * (R + m);
    if (r < lm * (m >> 1)) {
        const uint32_t t = r / (m >> 1), q = r % (m >> 1);
        // q-th index with bit t set
        const uint32_t lo = ((q >> t) << (t + 1)) | (1u << t) | (q & ((1u << t) - 1));
        ent[e] = basew + R + lo;
        return;
    }
    r -= lm * (m >> 1);
    if (r < lr * (R >> 1)) {
        const uint32_t t = r / (R >> 1), q = r % (R >> 1);
        const uint32_t hi = ((q >> t) << (t + 1)) | (1u << t) | (q & ((1u << t) - 1));
        ent[e] = basew + hi;
        return;
    }
    r -= lr * (R >> 1);
    ent[e] = basew + R + r;
}
// segment tables: nper segments per group with lengths given by a small pattern
__global__ void k_gen_segs_a(uint32_t W, uint32_t nb, uint32_t lm, uint32_t *__restrict__ start,
                             uint32_t *__restrict__ len) {
    const uint32_t m = 1u << lm, R = nb >> lm, per = R + m;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > W * per) return;
    if (s == W * per) {
        start[s] = W * 2 * nb;
        return;
    }
    const uint32_t w = s / per, r = s % per;
    if (r < R) {
        start[s] = w * 2 * nb + r * m;
        len[s] = m;
    } else {
        start[s] = w * 2 * nb + nb + (r - R) * R;
        len[s] = R;
    }
}
__global__ void k_gen_segs_b(uint32_t W, uint32_t lr, uint32_t lm, uint32_t *__restrict__ start,
                             uint32_t *__restrict__ len) {
    const uint32_t R = 1u << lr, m = 1u << lm, c = lr + lm + 1;
    const uint32_t per = lm * (m >> 1) + lr * (R >> 1) + m;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > W * c) return;
    if (s == W * c) {
        start[s] = W * per;
        return;
    }
    const uint32_t w = s / c, q = s % c;
    if (q < lm) {
        start[s] = w * per + q * (m >> 1);
        len[s] = m >> 1;
    } else if (q < lm + lr) {
        start[s] = w * per + lm * (m >> 1) + (q - lm) * (R >> 1);
        len[s] = R >> 1;
    } else {
        start[s] = w * per + lm * (m >> 1) + lr * (R >> 1);
        len[s] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// reduction levels as inversion-free trees: one block sums one segment (len <= 2 * blockDim points named by an
// index list) in Lopez-Dahab coordinates through shared memory; only the last level converts back to affine
// (one table-driven inversion per segment).  Two launches replace ~14 latency-bound affine rounds.
// ------------------------------------------------------------------------------------------------
struct GfMulCall {
    __device__ __forceinline__ gf operator()(const gf &a, const gf &b) const { return gf_mul_call(a, b); }
};
__device__ __forceinline__ void ld_store(LdPt *p, const LdPt &v) {
    gf_store(&p->X, v.X);
    gf_store(&p->Y, v.Y);
    gf_store(&p->Z, v.Z);
}
__device__ __forceinline__ LdPt ld_load(const LdPt *p) {
    LdPt r;
    r.X = gf_load(&p->X);
    r.Y = gf_load(&p->Y);
    r.Z = gf_load(&p->Z);
    return r;
}
template <bool SRC_LD, bool OUT_AFFINE>
__global__ void __launch_bounds__(256)
    k_ld_tree(const void *__restrict__ src, const uint32_t *__restrict__ ent, const uint32_t *__restrict__ seg_start,
              const uint32_t *__restrict__ seg_len, void *__restrict__ dst, const gf *__restrict__ tabs) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    LdPt *sh = reinterpret_cast<LdPt *>(sh_raw);
    const uint32_t s = blockIdx.x, t = threadIdx.x;
    const uint32_t start = seg_start[s], len = seg_len[s];
    auto fetch = [&](uint32_t pos) -> LdPt {
        const uint32_t e = ent ? ent[start + pos] : start + pos;
        if (SRC_LD) return ld_load(reinterpret_cast<const LdPt *>(src) + e);
        return ld_from_affine(pt_load(reinterpret_cast<const AffPt *>(src) + e));
    };
    LdPt acc = ld_infinity();
    if (2 * t < len) {
        acc = fetch(2 * t);
        if (2 * t + 1 < len) acc = ld_add_t(acc, fetch(2 * t + 1), GfMulCall());
    }
    ld_store(&sh[t], acc);
    __syncthreads();
    for (uint32_t n = (len + 1) >> 1; n > 1; n = (n + 1) >> 1) {
        const bool act = 2 * t < n;
        if (act) {
            acc = ld_load(&sh[2 * t]);
            if (2 * t + 1 < n) acc = ld_add_t(acc, ld_load(&sh[2 * t + 1]), GfMulCall());
        }
        __syncthreads();
        if (act) ld_store(&sh[t], acc);
        __syncthreads();
    }
    if (t < 32) { // warp 0 (the block has at least 32 threads): the conversion to affine is shared by its lanes
        const LdPt r = len ? ld_load(&sh[0]) : ld_infinity();
        if (OUT_AFFINE) {
            AffPt o = pt_inf();
            if (!gf_is_zero(r.Z)) { // warp-uniform
                const WarpMulCtx c = warp_mul_ctx();
                const gf zi = gf_inv_tab_warp(r.Z, tabs, c);
                o.x = gf_mul_warp(r.X, zi, c);
                o.y = gf_mul_warp(r.Y, gf_sqr(zi), c);
            }
            if (t == 0) pt_store(reinterpret_cast<AffPt *>(dst) + s, o);
        } else if (t == 0) {
            ld_store(reinterpret_cast<LdPt *>(dst) + s, r);
        }
    }
}

// The same tree for a level with FEW segments (level B: 15 sums per virtual window): the additions of a tree level are
// dealt to the warps of the block and every addition is done by a whole warp with the cooperative multiplier
// (gf233_warp.cuh: 14 products of 0.37 us instead of 1.9), the final conversion to affine likewise.  A level of the
// tree costs ~6 us per addition and warp instead of ~29 us: 7 levels over 128 points in ~75 us instead of ~180.
struct GfMulWarp {
    const WarpMulCtx &c;
    __device__ __forceinline__ gf operator()(const gf &a, const gf &b) const { return gf_mul_warp(a, b, c); }
};
// SRC_LD: the inputs are Lopez-Dahab points named by the index list `ent` (level B); otherwise affine points at the
// positions of the segment (level A after its batched-affine rounds: short segments, thousands of them -- the thread-per-
// addition tree above spends 0.29 ms of pure latency on them at 2^20 points).  OUT_AFFINE as above.
template <bool SRC_LD, bool OUT_AFFINE>
__global__ void __launch_bounds__(512)
    k_ld_tree_warp(const void *__restrict__ src, const uint32_t *__restrict__ ent, const uint32_t *__restrict__ seg_start,
                   const uint32_t *__restrict__ seg_len, void *__restrict__ dst, const gf *__restrict__ tabs) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    LdPt *sh = reinterpret_cast<LdPt *>(sh_raw); // the points of the segment, halved level by level
    const uint32_t s = blockIdx.x, t = threadIdx.x, warp = t >> 5, nwarps = blockDim.x >> 5;
    const uint32_t start = seg_start[s], len = seg_len[s];
    for (uint32_t i = t; i < len; i += blockDim.x) {
        if (SRC_LD) ld_store(&sh[i], ld_load(reinterpret_cast<const LdPt *>(src) + ent[start + i]));
        else ld_store(&sh[i], ld_from_affine(pt_load(reinterpret_cast<const AffPt *>(src) + start + i)));
    }
    __syncthreads();
    const WarpMulCtx wc = warp_mul_ctx();
    const GfMulWarp mul{wc};
    for (uint32_t n = len; n > 1; n = (n + 1) >> 1) {
        // pairs (2i, 2i+1) -> i; every warp takes pairs i = warp, warp + nwarps, ...; the operands are read by all its
        // lanes (warp-uniform), the result is kept in registers until every warp has read its inputs
        LdPt acc[4]; // at most 4 pairs per warp and level (len <= 8 * nwarps)
        uint32_t cnt = 0;
        for (uint32_t i = warp; 2 * i < n; i += nwarps, cnt++) {
            LdPt a = ld_load(&sh[2 * i]);
            if (2 * i + 1 < n) a = ld_add_t(a, ld_load(&sh[2 * i + 1]), mul);
            acc[cnt & 3] = a;
        }
        __syncthreads();
        cnt = 0;
        for (uint32_t i = warp; 2 * i < n; i += nwarps, cnt++)
            if ((t & 31) == 0) ld_store(&sh[i], acc[cnt & 3]);
        __syncthreads();
    }
    if (t < 32) {
        const LdPt r = len ? ld_load(&sh[0]) : ld_infinity();
        if (OUT_AFFINE) {
            AffPt o = pt_inf();
            if (!gf_is_zero(r.Z)) {
                const gf zi = gf_inv_tab_warp(r.Z, tabs, wc);
                o.x = gf_mul_warp(r.X, zi, wc);
                o.y = gf_mul_warp(r.Y, gf_sqr(zi), wc);
            }
            if (t == 0) pt_store(reinterpret_cast<AffPt *>(dst) + s, o);
        } else if (t == 0) {
            ld_store(reinterpret_cast<LdPt *>(dst) + s, r);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
int MsmLane::init(int index) {
    // high priority: when an MSM overlaps throughput-bound work on another stream (dvp_prove runs the g_m MSM beside the
    // extends), its short latency-bound kernels must not queue behind that work's blocks
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    CK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
    // The large pass kernels of lane i outrank those of lane i+1: the lanes then run staggered instead of in lock
    // step -- lane 0's large rounds take the SMs first, lane 1 fills the gaps of lane 0's inversion chains and is still
    // in its large rounds when lane 0 reaches its latency-bound small rounds (lower number = higher priority)
    const int prio_big = std::min(prio_lo, std::max(prio_hi + 1, prio_lo - 3 + index));
    CK(cudaStreamCreateWithPriority(&stream_lo, cudaStreamNonBlocking, prio_big));
    CK(cudaEventCreateWithFlags(&ev_sw[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_sw[1], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    for (auto &e : ev_k) CK(cudaEventCreate(&e));
    for (auto &e : ev_s) CK(cudaEventCreate(&e));
    return 0;
}
void MsmLane::prof_begin(int cat) {
    if (prof_used == prof.size()) {
        ProfRec r{cat, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        prof.push_back(r);
    }
    prof[prof_used].cat = cat;
    cudaEventRecord(prof[prof_used].e0, stream);
}
void MsmLane::prof_end() { cudaEventRecord(prof[prof_used++].e1, stream); }

void MsmLane::destroy() {
    for (auto &r : prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    prof.clear();
    prof_used = 0;
    DevBuf *all[] = {&seg_len[0], &seg_len[1], &seg_start[0], &seg_start[1], &c_len, &c_start,
                     &blk, &blk_flag, &info, &info_r0, &pp[0], &pp[1], &prefix, &desc, &thr_total, &thr_inv, &lvl_pre[0],
                     &lvl_pre[1], &lvl_tot[0], &lvl_tot[1], &lvl_inv[0], &lvl_inv[1], &buckets, &rc, &ents2, &acc_ctl, &acc_times};
    plan_main[0].release();
    plan_main[1].release();
    plan_a.release();
    for (auto b : all) b->release();
    for (auto &e : ev_k)
        if (e) cudaEventDestroy(e), e = nullptr;
    for (auto &e : ev_s)
        if (e) cudaEventDestroy(e), e = nullptr;
    if (done) cudaEventDestroy(done), done = nullptr;
    for (auto &e : ev_sw)
        if (e) cudaEventDestroy(e), e = nullptr;
    if (stream_lo) cudaStreamDestroy(stream_lo), stream_lo = nullptr;
    if (stream) cudaStreamDestroy(stream), stream = nullptr;
}

int MsmEngine::init(cudaStream_t s) {
    stream = s;
    for (auto &e : ev) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&ev_recode, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreate(&ev_t0[k]));
        CK(cudaEventCreate(&ev_t1[k]));
    }
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        CK(cudaStreamCreateWithPriority(&sort_stream, cudaStreamNonBlocking, prio_lo));
    }
    int rc = msqr_tabs.reserve(MSQ_TABLES * MSQ_TABLE_ELEMS * sizeof(gf));
    if (rc) return rc;
    {
        // the persistent accumulation kernel needs its grid co-resident: blocks per SM x SMs
        int dev = 0, sms = 0, occ = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        // (register budgets of 80 and 64 per thread -- 3 and 4 resident blocks -- were 10-35 % slower: spills)
        if (cudaFuncSetAttribute(k_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_BYTES) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_accumulate, ACC_THREADS, STG_BYTES) != cudaSuccess)
            occ = 0;
        acc_capacity = coop ? (uint32_t)(sms * std::min(occ, 2)) : 0u;
    }
    k_build_msqr_tables<<<cdiv(MSQ_TABLES * MSQ_TABLE_ELEMS, 128), 128, 0, s>>>(msqr_tabs.as<gf>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return 0;
}
void MsmEngine::destroy() {
    for (auto &l : lanes) l.destroy();
    lanes.clear();
    keys.release();
    entries.release();
    entries_b.release();
    len_all_b.release();
    start_all_b.release();
    if (sort_stream) cudaStreamDestroy(sort_stream), sort_stream = nullptr;
    start_all.release();
    cursor_all.release();
    scan_blk.release();
    lane_info.release();
    if (h_lane) cudaFreeHost(h_lane);
    h_lane = nullptr;
    len_all.release();
    hb.release();
    msqr_tabs.release();
    mg_table.release();
    for (int k = 0; k < 2; k++) {
        if (h_pts[k]) cudaFreeHost(h_pts[k]);
        h_pts[k] = nullptr;
        h_pts_cap[k] = 0;
    }
    for (auto &e : ev)
        if (e) cudaEventDestroy(e), e = nullptr;
    if (ev_recode) cudaEventDestroy(ev_recode), ev_recode = nullptr;
    for (int k = 0; k < 2; k++) {
        if (ev_t0[k]) cudaEventDestroy(ev_t0[k]), ev_t0[k] = nullptr;
        if (ev_t1[k]) cudaEventDestroy(ev_t1[k]), ev_t1[k] = nullptr;
    }
}

namespace {

constexpr int ACC_CTL_WORDS = 128;  // control block of one k_accumulate launch: barrier, abort, longest segments, result
constexpr uint32_t PLAN_MAX_BLOCKS = 256; // k_plan's grid must be co-resident (blocks wait for their predecessors)
constexpr uint32_t BINV_G = 16;        // group size of one batched-inversion level (large batches)

struct Tree {
    MsmEngine &E;
    MsmLane &L;
    cudaStream_t st;
    Tree(MsmEngine &e, MsmLane &l) : E(e), L(l), st(l.stream) {}
    // development profiler: CUDA events around the launches of one category (E.profile, see MsmLane)
    void pb(int cat) {
        if (E.profile) L.prof_begin(cat);
    }
    void pe() {
        if (E.profile) L.prof_end();
    }

    // one launch: scans, next segment tables, descriptors, and the last point of every odd segment of `src` carried
    // over to `dst` (the round's output list); info (incl. the ticket counter) is cleared here
    int plan(const uint32_t *len, const uint32_t *in_start, const uint32_t *ent, uint32_t nseg, uint32_t *out_start,
             uint32_t *new_len, const AffPt *src, AffPt *dst) {
        uint32_t items = 1;
        while (cdiv(nseg, items * PLAN_THREADS) > PLAN_MAX_BLOCKS) items++;
        const uint32_t nblk = cdiv(nseg, items * PLAN_THREADS);
        pb(PC_PLAN);
        CK(cudaMemsetAsync(L.info.p, 0, 16, st));
        if (ent)
            k_plan<true><<<nblk, PLAN_THREADS, 0, st>>>(len, in_start, ent, nseg, items, L.blk.as<uint64_t>(),
                                                        L.blk_flag.as<uint32_t>(), ++L.epoch, out_start, new_len,
                                                        L.desc.as<uint4>(), L.info.as<uint32_t>(), src, dst);
        else
            k_plan<false><<<nblk, PLAN_THREADS, 0, st>>>(len, in_start, nullptr, nseg, items, L.blk.as<uint64_t>(),
                                                         L.blk_flag.as<uint32_t>(), ++L.epoch, out_start, new_len,
                                                         L.desc.as<uint4>(), L.info.as<uint32_t>(), src, dst);
        pe();
        L.launches++;
        CK(cudaGetLastError());
        return 0;
    }

    // inverses of the pass-1 thread totals; n_ub bounds their number, the live count is read on the device.
    // Up to binv_coop_max elements: one cooperative launch (k_binv_coop).  More: one or two thread-per-group levels
    // (throughput-bound at that size) bring the batch down to that size first.
    int batch_inv(const gf *in, gf *out, uint32_t n_ub, uint32_t unit, uint32_t div1, uint32_t div2, int depth) {
        const uint32_t *info = L.info.as<uint32_t>();
        const uint32_t COOP_MAX = E.binv_direct;
        if (n_ub <= COOP_MAX || depth >= 2) {
            // the smallest group that keeps the launch within one wave of resident warps
            uint32_t G = 1;
            while (G < 32 && cdiv(n_ub, G) > E.binv_coop_warps) G <<= 1;
            pb(PC_BINV_DIRECT);
            k_binv_coop<<<cdiv(cdiv(n_ub, G), 4), 128, 0, st>>>(in, out, info, unit, div1, div2, G, E.msqr_tabs.as<gf>());
            pe();
            L.launches++;
            CK(cudaGetLastError());
            return 0;
        }
        uint32_t G = 2;
        while (G < BINV_G && cdiv(n_ub, G) > COOP_MAX) G <<= 1;
        const uint32_t ng = cdiv(n_ub, G);
        gf *pre = L.lvl_pre[depth].as<gf>(), *tot = L.lvl_tot[depth].as<gf>(), *inv = L.lvl_inv[depth].as<gf>();
        pb(PC_BINV_UP);
        k_binv_up<<<cdiv(ng, 64), 64, 0, st>>>(in, info, unit, div1, div2, G, pre, tot);
        pe();
        L.launches++;
        int rc = depth == 0 ? batch_inv(tot, inv, ng, unit, G, 1, 1) : batch_inv(tot, inv, ng, unit, div1, G, 2);
        if (rc) return rc;
        pb(PC_BINV_DOWN);
        k_binv_down<<<cdiv(ng, 64), 64, 0, st>>>(in, info, unit, div1, div2, G, pre, inv, out);
        pe();
        L.launches++;
        CK(cudaGetLastError());
        return 0;
    }

    template <int B> int round_t(const AffPt *src, size_t task_ub, AffPt *dst) {
        const uint32_t nblk = cdiv(cdiv(task_ub, B), 256);
        const uint32_t nthr = nblk * 256;
        const uint32_t *info = L.info.as<uint32_t>();
        // Large pass kernels detour through the lane's low-priority stream: the short, latency-bound kernels of the
        // OTHER lanes (inversion chain, plans, small rounds) then take the next free block slot instead of queueing
        // behind every block of this launch.
        const bool detour = E.prio_split && task_ub >= (1u << 17);
        cudaStream_t sb = detour ? L.stream_lo : st;
        auto hop = [&](cudaStream_t from, cudaStream_t to, int e) {
            cudaEventRecord(L.ev_sw[e], from);
            cudaStreamWaitEvent(to, L.ev_sw[e], 0);
        };
        if (E.fused_rounds) {
            // pass 1, the warps' own inversions and pass 2 in one launch (no thread totals leave the warp)
            const bool mark = L.want_k;
            if (mark) cudaEventRecord(L.ev_k[0], st);
            pb(PC_PASS2);
            if (detour) hop(st, sb, 0);
            cudaFuncSetAttribute(k_round_fused<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_BYTES);
            k_round_fused<B><<<nblk, 256, STG_BYTES, sb>>>(src, info, L.desc.as<uint4>(), L.prefix.as<gf>(), dst, E.msqr_tabs.as<gf>());
            if (detour) hop(sb, st, 1);
            pe();
            if (mark) {
                cudaEventRecord(L.ev_k[1], st);
                L.want_k = false;
            }
            L.launches++;
            CK(cudaGetLastError());
            return 0;
        }
        pb(PC_PASS1);
        if (detour) hop(st, sb, 0);
        k_pass1<B><<<nblk, 256, 0, sb>>>(src, info, L.desc.as<uint4>(), L.prefix.as<gf>(), L.thr_total.as<gf>());
        if (detour) hop(sb, st, 1);
        pe();
        L.launches++;
        int rc = batch_inv(L.thr_total.as<gf>(), L.thr_inv.as<gf>(), nthr, 32u * B, 1, 1, 0);
        if (rc) return rc;
        const bool mark = L.want_k;
        if (mark) cudaEventRecord(L.ev_k[0], st);
        pb(PC_PASS2);
        if (detour) hop(st, sb, 0);
        // 88 KB of staging slots per block: above the 48 KB default, so the limit is raised (per device, cheap)
        cudaFuncSetAttribute(k_pass2<B, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_BYTES);
        cudaFuncSetAttribute(k_pass2<B, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_BYTES);
        cudaFuncSetAttribute(k_pass2<B, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_BYTES);
        if (E.pass2_minb == 2)
            k_pass2<B, 2><<<nblk, 256, STG_BYTES, sb>>>(src, info, L.desc.as<uint4>(), L.prefix.as<gf>(), L.thr_inv.as<gf>(), dst);
        else if (E.pass2_minb == 3)
            k_pass2<B, 3><<<nblk, 256, STG_BYTES, sb>>>(src, info, L.desc.as<uint4>(), L.prefix.as<gf>(), L.thr_inv.as<gf>(), dst);
        else
            k_pass2<B, 1><<<nblk, 256, STG_BYTES, sb>>>(src, info, L.desc.as<uint4>(), L.prefix.as<gf>(), L.thr_inv.as<gf>(), dst);
        if (detour) hop(sb, st, 1);
        pe();
        if (mark) {
            cudaEventRecord(L.ev_k[1], st);
            L.want_k = false;
        }
        L.launches++;
        CK(cudaGetLastError());
        return 0;
    }
    // a round of at most this many additions runs one warp per addition (k_round_warp)
    int round_warp(const AffPt *src, size_t task_ub, AffPt *dst) {
        pb(PC_PASS2);
        k_round_warp<<<cdiv(task_ub, 4), 128, 0, st>>>(src, L.info.as<uint32_t>(), L.desc.as<uint4>(), dst, E.msqr_tabs.as<gf>());
        pe();
        L.launches++;
        CK(cudaGetLastError());
        return 0;
    }
    int round(int B, const AffPt *src, size_t task_ub, AffPt *dst) {
        if (task_ub <= E.round_warp_max) return round_warp(src, task_ub, dst);
        if (B == 64) return round_t<64>(src, task_ub, dst);
        if (B == 16) return round_t<16>(src, task_ub, dst);
        if (B == 4) return round_t<4>(src, task_ub, dst);
        return round_t<1>(src, task_ub, dst);
    }

    // All rounds of one reduction in a single persistent launch (k_accumulate).  nr_fixed < 0: until every segment has
    // at most one point, then dst[s] = that point; nr_fixed >= 0: exactly that many rounds, the list is left in
    // pp[(nr-1) & 1] with the segment tables of set nr & 1 (dst may be null).  `slot` selects the control block.
    // The plan of all rounds of one k_accumulate launch into `ps`, on stream s.  With a key, a set that already holds
    // the plan of that key is left as it is.
    int plan_launch(cudaStream_t s, MsmLane::PlanSet &ps, const uint32_t *len0, const uint32_t *start0, const uint32_t *ent,
                    uint32_t nseg, size_t total_ub, int max_rounds, uint64_t key = 0) {
        if (key && ps.key == key) return 0;
        ps.key = 0;
        int rc;
        uint32_t items = 1;
        while (cdiv(nseg, items * PL_THREADS) > PL_MAX_BLOCKS) items++;
        const uint32_t nblk = cdiv(nseg, items * PL_THREADS);
        const uint32_t stride = (nseg + 32) & ~31u;
        if ((rc = ps.blk.reserve((size_t)nblk * PL_K * 8)) != 0) return rc;
        if ((rc = ps.start.reserve((size_t)(PL_K + 1) * stride * 4)) != 0) return rc;
        if ((rc = ps.len.reserve((size_t)(PL_K + 1) * stride * 4)) != 0) return rc;
        if ((rc = ps.ctl.reserve(PL_CTL_WORDS * 4)) != 0) return rc;
        if ((rc = ps.desc.reserve((total_ub + 64) * sizeof(uint4))) != 0) return rc; // all rounds: fewer additions than points
        ps.stride = stride;
        uint32_t *pc = ps.ctl.as<uint32_t>();
        unsigned long long *blk = ps.blk.as<unsigned long long>();
        CK(cudaMemsetAsync(pc, 0, PL_CTL_WORDS * 4, s));
        k_pa_aggr<<<nblk, PL_THREADS, 0, s>>>(len0, nseg, items, blk, pc);
        k_pa_scan<<<1, PL_THREADS, 0, s>>>(blk, nblk, max_rounds, pc);
        if (ent)
            k_pa_emit<true><<<nblk, PL_THREADS, 0, s>>>(len0, start0, ent, nseg, items, blk, pc, ps.start.as<uint32_t>(),
                                                        ps.len.as<uint32_t>(), stride, ps.desc.as<uint4>());
        else
            k_pa_emit<false><<<nblk, PL_THREADS, 0, s>>>(len0, start0, nullptr, nseg, items, blk, pc, ps.start.as<uint32_t>(),
                                                         ps.len.as<uint32_t>(), stride, ps.desc.as<uint4>());
        CK(cudaGetLastError());
        L.launches += 3;
        ps.key = key;
        return 0;
    }

    int accumulate(const AffPt *src, const uint32_t *ent, const uint32_t *start0, const uint32_t *len0, uint32_t nseg,
                   size_t total_ub, AffPt *dst, int nr_fixed, int slot, uint32_t grid_cap, MsmLane::PlanSet *ps = nullptr,
                   bool plan_here = true, uint64_t plan_key = 0) {
        AccArgs A;
        A.src0 = src;
        A.ent = ent;
        A.start0 = start0;
        A.len0 = len0;
        A.nseg = nseg;
        for (int i = 0; i < 2; i++) {
            A.pp[i] = L.pp[i].as<AffPt>();
            A.seg_start[i] = L.seg_start[i].as<uint32_t>();
            A.seg_len[i] = L.seg_len[i].as<uint32_t>();
        }
        A.desc = L.desc.as<uint4>();
        A.prefix = L.prefix.as<gf>();
        A.tot = L.thr_total.as<gf>();
        A.tot_inv = L.thr_inv.as<gf>();
        A.blk_sum = L.blk.as<unsigned long long>();
        A.ctl = L.acc_ctl.as<uint32_t>() + slot * ACC_CTL_WORDS;
        A.result = A.ctl + 64;
        A.times = E.profile ? L.acc_times.as<unsigned long long>() + slot * ACC_TIME_WORDS : nullptr;
        A.dst = dst;
        A.tabs = E.msqr_tabs.as<gf>();
        A.max_rounds = nr_fixed >= 0 ? nr_fixed : ACC_MAX_ROUNDS;
        A.tiny_max = (int)E.round_warp_max;
        A.pl_ctl = nullptr;
        A.pl_start = A.pl_len = nullptr;
        A.pl_stride = 0;
        if (ps) {
            // planned ahead (see k_pa_*): by the caller on the sort stream, here on the lane's stream, or kept from the
            // MSM before (level A)
            if (plan_here) {
                int rc = plan_launch(st, *ps, len0, start0, ent, nseg, total_ub, A.max_rounds, plan_key);
                if (rc) return rc;
            }
            A.pl_ctl = ps->ctl.as<uint32_t>();
            A.pl_start = ps->start.as<uint32_t>();
            A.pl_len = ps->len.as<uint32_t>();
            A.pl_stride = ps->stride;
            A.desc = ps->desc.as<uint4>(); // (also the round-by-round plan's, should the device fall back to it)
        }
        // enough threads for one addition each in round 0 (a small problem pays for its barriers by the block)
        const size_t want = std::max<size_t>(total_ub / 2 + 1, nseg);
        const uint32_t grid = (uint32_t)std::max<size_t>(1, std::min<size_t>(grid_cap, (want + ACC_THREADS - 1) / ACC_THREADS));
        CK(cudaMemsetAsync(A.ctl, 0, ACC_CTL_WORDS * 4, st));
        void *params[] = {(void *)&A};
        pb(PC_PASS2);
        CK(cudaLaunchCooperativeKernel((const void *)k_accumulate, dim3(grid), dim3(ACC_THREADS), params, STG_BYTES, st));
        pe();
        L.launches++;
        return 0;
    }

    // Plan of round 0: caller tables -> lane set 1 (+ descriptors); info[0] = longest segment,
    // info[1] = additions of round 0.
    int plan0(const uint32_t *start0, const uint32_t *len0, const uint32_t *ent, uint32_t nseg, const AffPt *src) {
        return plan(len0, start0, ent, nseg, L.seg_start[1].as<uint32_t>(), L.seg_len[1].as<uint32_t>(), src,
                    L.pp[0].as<AffPt>());
    }

    // Reduce every segment of the index list `ent` over `src` to one point: dst[s], s < nseg, given that
    // plan0 has run.  start0 (nseg+1) / len0 (nseg) describe the segments and are only read; they must not
    // be the lane's own seg_start[] / seg_len[] ping-pong arrays.  total_ub bounds the entry count.
    // With `stop` set the rounds end as soon as at most stop->max_left points can remain (a later pass finishes the
    // segments); *stop then describes the partially reduced list and dst is not written.
    struct Partial {
        size_t max_left;
        const AffPt *src;          // out: the list after the last round run
        const uint32_t *start, *len; // out: its segment tables (nseg entries; the caller's own when no round ran)
        uint32_t maxlen;           // out: bound on the remaining segment length
    };
    int rounds(const AffPt *src, const uint32_t *ent, const uint32_t *start0, const uint32_t *len0, uint32_t nseg,
               size_t total_ub, uint32_t maxlen, AffPt *dst, int *rounds_out, Partial *stop = nullptr) {
        uint32_t *elen[2] = {L.seg_len[0].as<uint32_t>(), L.seg_len[1].as<uint32_t>()};
        uint32_t *estart[2] = {L.seg_start[0].as<uint32_t>(), L.seg_start[1].as<uint32_t>()};
        int nr = 0;
        while ((1ull << nr) < maxlen) nr++;
        if (stop) {
            // after r rounds at most total/2^r + nseg points remain
            int r = 0;
            while (r < nr && (total_ub >> r) + nseg > stop->max_left) r++;
            nr = r;
            stop->src = src;
            stop->start = start0;
            stop->len = len0;
            stop->maxlen = maxlen;
        }
        if (rounds_out) *rounds_out = nr;
        if (nr == 0 && stop) return 0;
        if (nr == 0) {
            pb(PC_MISC);
            k_finalize<true><<<cdiv(nseg, 256), 256, 0, st>>>(src, ent, start0, len0, nseg, dst);
            pe();
            L.launches++;
            CK(cudaGetLastError());
            return 0;
        }
        const uint32_t *in_start = start0, *in_len = len0;
        const AffPt *cur_src = src;
        int rc;
        for (int r = 0; r < nr; r++) {
            const int o = (r + 1) & 1; // lane set written by this round's plan
            // tasks_r <= total/2^(r+1) + nseg/2
            const size_t task_ub = r == 0 ? total_ub / 2 + 1 : (total_ub >> (r + 1)) + nseg / 2 + 1;
            // additions chained per thread: long chains in the throughput-bound rounds; in the latency-bound ones just
            // enough that the thread totals fit one cooperative inversion launch
            int B = task_ub >= E.b64_min ? 64 : task_ub >= E.b16_min ? 16 : task_ub > E.binv_direct ? 4 : 1;
            B = std::min(B, E.pass_b_max);
            AffPt *out = L.pp[r & 1].as<AffPt>();
            if ((rc = round(B, cur_src, task_ub, out))) return rc; // (the odd leftovers were carried over by the plan)
            cur_src = out;
            in_start = estart[o];
            in_len = elen[o];
            if (r + 1 < nr &&
                (rc = plan(in_len, in_start, nullptr, nseg, estart[o ^ 1], elen[o ^ 1], cur_src, L.pp[(r + 1) & 1].as<AffPt>())))
                return rc;
        }
        if (stop) {
            stop->src = cur_src;
            stop->start = in_start;
            stop->len = in_len;
            stop->maxlen = (maxlen + (1u << nr) - 1) >> nr;
            return 0;
        }
        pb(PC_MISC);
        k_finalize<false><<<cdiv(nseg, 256), 256, 0, st>>>(cur_src, nullptr, in_start, in_len, nseg, dst);
        pe();
        L.launches++;
        CK(cudaGetLastError());
        return 0;
    }
};

} // namespace

// scratch of one tree round with up to task_ub additions
int MsmEngine::reserve_round(MsmLane &L, size_t task_ub) {
    int rc;
#define RS(buf, bytes) \
    if ((rc = (buf).reserve(bytes)) != 0) return rc
    RS(L.info, 64);
    RS(L.info_r0, 64);
    RS(L.prefix, task_ub * sizeof(gf));
    RS(L.desc, task_ub * sizeof(uint4));
    // pass-1 threads: rounds of >= 2^21 additions chain min(16, pass_b_max) per thread, smaller ones at least 4 or 1
    const size_t thr_ub = std::max<size_t>(task_ub / (size_t)std::min(16, pass_b_max), 1u << 19) + 1024;
    RS(L.thr_total, thr_ub * sizeof(gf));
    RS(L.thr_inv, thr_ub * sizeof(gf));
    RS(L.lvl_pre[0], thr_ub * sizeof(gf));
    RS(L.lvl_tot[0], (thr_ub / 2 + 2) * sizeof(gf));
    RS(L.lvl_inv[0], (thr_ub / 2 + 2) * sizeof(gf));
    RS(L.lvl_pre[1], (thr_ub / 2 + 2) * sizeof(gf));
    RS(L.lvl_tot[1], (thr_ub / 4 + 2) * sizeof(gf));
    RS(L.lvl_inv[1], (thr_ub / 4 + 2) * sizeof(gf));
#undef RS
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched fixed-base multiplication: out[i] = k_i G  (CurvePoint::generator().mul, curve.rs:84-91,129-137;
// compute_srs_matrices, srs.rs:126-160).  8-bit windows over a table T[j][d] = d 2^(8j) G; window j is one
// tree round in which every accumulator adds its table point, so all n additions share one inversion.
// ------------------------------------------------------------------------------------------------
constexpr int MG_WINDOWS = 30, MG_DIGITS = 255;
constexpr size_t MG_TABLE = (size_t)MG_WINDOWS * MG_DIGITS + 1; // + one point at infinity (zero digits)

__global__ void k_mulgen_desc(const uint32_t *__restrict__ scalars, uint32_t n, int j, uint32_t tab0,
                              uint4 *__restrict__ desc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr a;
    const uint4 *q = reinterpret_cast<const uint4 *>(scalars + (size_t)i * 8);
    const uint4 lo = q[0], hi = q[1];
    a.v[0] = lo.x; a.v[1] = lo.y; a.v[2] = lo.z; a.v[3] = lo.w;
    a.v[4] = hi.x; a.v[5] = hi.y; a.v[6] = hi.z; a.v[7] = hi.w;
    uint32_t k[8];
    fr_to_canonical(k, a);
    const uint32_t d = (k[j >> 2] >> (8 * (j & 3))) & 255u;
    desc[i] = make_uint4(i, d ? tab0 + (uint32_t)j * MG_DIGITS + d - 1 : tab0 + (uint32_t)(MG_TABLE - 1), i, 0);
}

__global__ void k_self_desc(uint32_t n, uint4 *__restrict__ desc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) desc[i] = make_uint4(i, i, i, 0);
}

// T[j] = 2^(width of window j-1) T[j-1]: rounds of n batched affine doublings (a pair (P, P) is the tangent case of
// the addition); the windows are those of run(): width base + (j < rem) with base = 233 / W, rem = 233 % W
int MsmEngine::build_table(const AffPt *d_points, size_t n, int W, AffPt *d_tab) {
    const int base = 233 / W, rem = 233 % W;
    if (n == 0) return 0;
    if (lanes.empty()) {
        lanes.emplace_back();
        int rc0 = lanes.back().init((int)lanes.size() - 1);
        if (rc0) return rc0;
    }
    MsmLane &L = lanes[0];
    cudaStream_t st = L.stream;
    int rc;
    const size_t CH = (size_t)1 << 22; // points per pass (bounds the scratch)
    const size_t chunk_max = std::min(n, CH);
    if ((rc = reserve_round(L, chunk_max + 1))) return rc;
    for (int i = 0; i < 2; i++)
        if ((rc = L.pp[i].reserve(chunk_max * sizeof(AffPt)))) return rc;
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemcpyAsync(d_tab, d_points, n * sizeof(AffPt), cudaMemcpyDeviceToDevice, st));
    Tree tree(*this, L);
    for (size_t off = 0; off < n; off += CH) {
        const uint32_t m = (uint32_t)std::min(CH, n - off);
        const uint32_t info_h[4] = {0, m, m, 0};
        CK(cudaMemcpyAsync(L.info.p, info_h, 16, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st)); // info_h is a stack temporary
        k_self_desc<<<cdiv(m, 256), 256, 0, st>>>(m, L.desc.as<uint4>());
        const int B = m >= (1u << 21) ? 16 : m >= (1u << 17) ? 4 : 1;
        for (int j = 1; j < W; j++) {
            const AffPt *src = d_tab + (size_t)(j - 1) * n + off;
            AffPt *dst = d_tab + (size_t)j * n + off;
            const int c = base + (j - 1 < rem ? 1 : 0);
            for (int r = 0; r < c; r++) {
                AffPt *out = r == c - 1 ? dst : L.pp[r & 1].as<AffPt>();
                if ((rc = tree.round(B, src, m, out))) return rc;
                src = out;
            }
        }
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int MsmEngine::mulgen(const uint32_t *d_scalars, size_t n, AffPt *d_out) {
    if (n == 0) return 0;
    if (lanes.empty()) {
        lanes.emplace_back();
        int rc0 = lanes.back().init((int)lanes.size() - 1);
        if (rc0) return rc0;
    }
    MsmLane &L = lanes[0];
    cudaStream_t st = L.stream;
    int rc;
    // table on the host (once): LD doublings / additions, one conversion per entry
    if (!mg_table.p) {
        std::vector<AffPt> tab(MG_TABLE);
        AffPt base = host::k233_generator();
        for (int j = 0; j < MG_WINDOWS; j++) {
            host::LdPt acc = host::ld_inf();
            for (int d = 1; d <= MG_DIGITS; d++) {
                acc = host::ld_add_affine(acc, base);
                tab[(size_t)j * MG_DIGITS + d - 1] = host::ld_to_affine(acc);
            }
            host::LdPt nb = host::ld_add_affine(acc, base); // 256 * base
            base = host::ld_to_affine(nb);
        }
        tab[MG_TABLE - 1] = pt_inf();
        if ((rc = mg_table.reserve(MG_TABLE * sizeof(AffPt)))) return rc;
        CK(cudaMemcpy(mg_table.p, tab.data(), MG_TABLE * sizeof(AffPt), cudaMemcpyHostToDevice));
    }
    const size_t CH = (size_t)1 << 22; // accumulators per pass (bounds the scratch)
    const size_t chunk_max = std::min(n, CH);
    if ((rc = reserve_round(L, chunk_max + 1))) return rc;
    for (int i = 0; i < 2; i++)
        if ((rc = L.pp[i].reserve((chunk_max + MG_TABLE) * sizeof(AffPt)))) return rc;
    CK(cudaStreamSynchronize(stream)); // the scalars were written on the context stream
    Tree tree(*this, L);
    for (size_t off = 0; off < n; off += CH) {
        const uint32_t m = (uint32_t)std::min(CH, n - off);
        AffPt *A[2] = {L.pp[0].as<AffPt>(), L.pp[1].as<AffPt>()};
        for (int i = 0; i < 2; i++)
            CK(cudaMemcpyAsync(A[i] + m, mg_table.p, MG_TABLE * sizeof(AffPt), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemsetAsync(A[0], 0, (size_t)m * sizeof(AffPt), st)); // accumulators start at infinity
        const uint32_t info_h[4] = {0, m, m, 0};
        CK(cudaMemcpyAsync(L.info.p, info_h, 16, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st)); // info_h is a stack temporary
        const int B = m >= (1u << 21) ? 16 : m >= (1u << 17) ? 4 : 1;
        for (int j = 0; j < MG_WINDOWS; j++) {
            k_mulgen_desc<<<cdiv(m, 256), 256, 0, st>>>(d_scalars + off * 8, m, j, m, L.desc.as<uint4>());
            if ((rc = tree.round(B, A[j & 1], m, A[(j + 1) & 1]))) return rc;
        }
        CK(cudaMemcpyAsync(d_out + off, A[MG_WINDOWS & 1], (size_t)m * sizeof(AffPt), cudaMemcpyDeviceToDevice, st));
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

// cost model of the shared-bucket-set layout: W n bucket additions + ~3 * 2^c for the reduction (measured optimum)
int choose_table_windows(size_t n) {
    int best = 16;
    double best_cost = 1e300;
    for (int W = 9; W <= 30; W++) {
        const int c = 233 / W + (233 % W ? 1 : 0);
        const double cost = (double)n * W + 3.0 * (double)(1ull << c);
        if (cost < best_cost) best_cost = cost, best = W;
    }
    return best;
}

int MsmEngine::run(const AffPt *d_points, const uint32_t *d_scalars, size_t n, AffPt *h_result, const MsmTable *tab) {
    *h_result = pt_inf();
    Pending P;
    int rc = enqueue(d_points, d_scalars, n, tab, 0, &P);
    if (rc) return rc;
    return finish(P, h_result);
}

int MsmEngine::enqueue(const AffPt *d_points, const uint32_t *d_scalars, size_t n, const MsmTable *tab, int k, Pending *P,
                       bool ahead, cudaEvent_t scalars_ready) {
    P->active = false;
    if (n == 0) return 0;
    if (n >= (1ull << 31) || k < 0 || k > 1) return DVP_ERR_BAD_ARG;
    // ---- window layout
    int W, base, rem, c;
    const bool uniform = tab != nullptr;
    if (uniform) {
        // shared bucket set over the tables T[j] = 2^(off_j) P
        W = tab->W;
        base = 233 / W;
        rem = 233 % W;
        c = base + (rem ? 1 : 0);
    } else {
        const int c_req = force_window_bits ? force_window_bits : choose_window_bits(n);
        if (c_req < 4 || c_req > 20) return DVP_ERR_BAD_ARG;
        // W windows of width base or base+1 covering exactly 233 bits (bit 232 of a scalar < p is zero, so the
        // top window never carries out); c = the widest window sizes the bucket tables
        W = (233 + c_req - 1) / c_req;
        base = 233 / W;
        rem = 233 % W;
        c = base + (rem ? 1 : 0);
    }
    const uint32_t nb = 1u << (c - 1);                    // buckets of one set
    const size_t NB = uniform ? nb : (size_t)W * nb;      // all buckets = segments of the sorted entry list
    const size_t total = (size_t)W * n;
    if (total >= (1ull << 32) || NB >= (1ull << 31) || c < 4 || c > 26) return DVP_ERR_BAD_ARG;
    // virtual windows: aligned ranges of nbv buckets, the unit of the reduction and of the lane split
    // (at most 32 virtual windows in the shared layout: the host tail handles V * cv points)
    const int cv = uniform ? std::min(c, std::max(15, c - 5)) : c; // log2(nbv) + 1
    const uint32_t nbv = 1u << (cv - 1);
    const uint32_t V = (uint32_t)(NB / nbv);
    // bucket matrix of a virtual window: m = 2^lm columns, R = 2^lr rows, nbv = R*m
    const uint32_t lm = (uint32_t)cv / 2, lr = (uint32_t)(cv - 1) - lm;
    const uint32_t R = 1u << lr, m = 1u << lm;
    const uint32_t per_b = lm * (m >> 1) + lr * (R >> 1) + m;
    // lanes: independent chains of rounds over disjoint ranges of virtual windows
    // The persistent kernel (one block of 255-register threads per SM) runs in ONE lane: two lanes would halve each
    // grid (2^20: 6.95 ms against 7.17); the separate-launch path keeps two lanes (2^22: 22.2 against 23.8 ms).
    const bool persistent_any =
        acc_capacity > 0 && (use_accumulate == 2 || (use_accumulate == 1 && n >= ((size_t)1 << 10) && n < ((size_t)3 << 20)));
    const bool nosync = persistent_any && !timing && total <= ((size_t)1 << 25);
    int NL = (profile && !force_lanes) ? 1 : force_lanes ? force_lanes : persistent_any ? 1 : (n >= (1u << 13) ? 2 : 1);
    NL = std::max(1, std::min<int>(NL, (int)V));
    while ((int)lanes.size() < NL) {
        lanes.emplace_back();
        int rc0 = lanes.back().init((int)lanes.size() - 1);
        if (rc0) return rc0;
    }

    int rc;
#define RS(buf, bytes) \
    if ((rc = (buf).reserve(bytes)) != 0) return rc
    // sort ahead: on its own stream, into buffer set k, ordered only after the scalars and after the MSM that used
    // set k before (two MSMs back in a pipelined batch)
    const bool side = ahead && sort_ahead && nosync && !profile && sort_stream;
    DevBuf &entries_k = (side && k) ? entries_b : entries, &len_k = (side && k) ? len_all_b : len_all,
           &start_k = (side && k) ? start_all_b : start_all;
    RS(keys, total * 4);
    RS(entries_k, total * 4);
    RS(len_k, (NB + 1) * 4);
    RS(start_k, (NB + 1) * 4);
    RS(cursor_all, (NB + 1) * 4);
    RS(scan_blk, (NB / SCAN_TILE + 8) * 8);
    RS(lane_info, 64 * 4);
    RS(hb, (size_t)V * cv * sizeof(AffPt));
    if (!h_lane) CK(cudaMallocHost(&h_lane, 256 * 4));
    uint32_t *const h_lane_k = (uint32_t *)h_lane + 128 * k;
    const size_t hb_bytes = (size_t)V * cv * sizeof(AffPt);
    if (h_pts_cap[k] < hb_bytes) {
        if (h_pts[k]) cudaFreeHost(h_pts[k]);
        h_pts[k] = nullptr;
        h_pts_cap[k] = 0;
        CK(cudaMallocHost(&h_pts[k], hb_bytes));
        h_pts_cap[k] = hb_bytes;
    }
    struct Part {
        uint32_t v0, vn;
        uint32_t nseg, nseg_a, nent_a, nseg_b, nent_b;
        size_t total;
        uint32_t maxlen;
    };
    std::vector<Part> part(NL);
    uint32_t bounds[17];
    for (int l = 0; l < NL; l++) {
        Part &p = part[l];
        p.v0 = (uint32_t)((unsigned long long)V * l / NL);
        p.vn = (uint32_t)((unsigned long long)V * (l + 1) / NL) - p.v0;
        p.nseg = p.vn * nbv;
        p.nseg_a = p.vn * (R + m);
        p.nent_a = p.vn * 2 * nbv;
        p.nseg_b = p.vn * (uint32_t)cv;
        p.nent_b = p.vn * per_b;
        bounds[l] = p.v0 * nbv;
    }
    bounds[NL] = (uint32_t)NB;

    // The persistent kernel wins from 2^10 to 2^21 points (with the rounds planned ahead: 2^10 0.77 against 0.81 ms, 2^12
    // 0.91 / 1.03, 2^13 1.03 / 1.23, 2^14 1.06 / 1.30, 2^15 1.00 / 1.19, 2^16 1.23 / 1.51 --
    // profiles/r2t_persistent_small.log, r2t_persistent_tiny.log; 2^18 and up were measured
    // with the round-by-round plan: 2^18 3.00 / 3.17, 2^19 4.67 / 4.95, 2^20 7.72 / 7.92, 2^21 13.60 / 13.72; 17
    // launches instead of 50-155); above, the separate large launches keep two blocks of one lane on every SM
    // (2^22: 24.4 / 23.4 ms then, 30.5 / 22.3 now: profiles/r2t_persistent_forced_large_sizes.log).
    // With the persistent kernel the host needs nothing from the sort: the rounds are counted on the device and the
    // scratch is sized from upper bounds (every entry in one lane), so the MSM is enqueued without a read-back in the
    // middle.  (Larger MSMs keep the exact sizes: twice the scratch would be gigabytes.)
    cudaStream_t st = stream;
    cudaEventRecord(ev_t0[k], st);
    if (timing) cudaEventRecord(ev[0], st);
    // ---- recode + histogram, counting sort of all (point, window) entries by bucket (context stream, or ahead of it)
    uint32_t *d_len_all = len_k.as<uint32_t>(), *d_start_all = start_k.as<uint32_t>();
    const cudaStream_t st_main = st;
    if (side) {
        st = sort_stream;
        if (scalars_ready) CK(cudaStreamWaitEvent(st, scalars_ready, 0));
        CK(cudaStreamWaitEvent(st, ev_t1[k], 0)); // the MSM that read set k before (never recorded: no wait)
    } else if (scalars_ready) {
        CK(cudaStreamWaitEvent(st, scalars_ready, 0));
    }
    CK(cudaMemsetAsync(d_len_all, 0, NB * 4, st));
    k_recode_count<<<cdiv(n, 128), 128, 0, st>>>(d_scalars, (uint32_t)n, base, rem, W, nb, uniform ? 1 : 0,
                                                 keys.as<uint32_t>(), d_len_all);
    {
        const uint32_t nblk = cdiv(NB, SCAN_TILE);
        k_scan1<<<nblk, SCAN_THREADS, 0, st>>>(d_len_all, (uint32_t)NB, scan_blk.as<uint64_t>());
        k_scan2<<<1, SCAN_THREADS, 0, st>>>(scan_blk.as<uint64_t>(), nblk);
        k_scan3<<<nblk, SCAN_THREADS, 0, st>>>(d_len_all, (uint32_t)NB, scan_blk.as<uint64_t>(), d_start_all,
                                               cursor_all.as<uint32_t>());
        k_scatter<<<cdiv(total, 256), 256, 0, st>>>(keys.as<uint32_t>(), (uint32_t)n, total,
                                                    uniform ? (uint32_t)tab->offset : 0u,
                                                    uniform ? (uint32_t)tab->stride : 0u, cursor_all.as<uint32_t>(),
                                                    entries_k.as<uint32_t>());
        if (!nosync) {
            CK(cudaMemcpyAsync(lane_info.as<uint32_t>() + 32, bounds, (NL + 1) * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(lane_info.p, 0, 32 * 4, st));
            k_lane_info<<<dim3(std::max(1u, std::min(148u, cdiv(NB / NL, 1024))), NL), 256, 0, st>>>(
                d_len_all, d_start_all, lane_info.as<uint32_t>() + 32, lane_info.as<uint32_t>());
            CK(cudaMemcpyAsync(h_lane_k, lane_info.p, 2 * NL * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaGetLastError());
    }
    // the stream the sort ran on also plans the rounds of the bucket accumulation (k_pa_*: they need the sorted list
    // only), so that in a pipelined batch the plan of MSM b+1 is ready before MSM b has finished
    const cudaStream_t ss = st;
    // Planning ahead pays in the table layout at every size measured (2^17 .. 2^21 points: 3-8 %) and in the plain layout
    // up to ~6 M entries (2^12 .. 2^18 points: 6-12 %); above that the plain layout LOSES 12-15 % (2^19: 6.03 against 5.24
    // ms, 2^20: 9.89 / 8.86 -- pass 2 of rounds 0 and 1 runs slower, profiles/r2u_timeline_plain_pre*.log), so it plans
    // round by round there.
    const bool pre_on = preplan && (uniform || total <= ((size_t)3 << 21));
    const bool plan_on_sort = pre_on && persistent_any && nosync;
    if (!plan_on_sort) CK(cudaEventRecord(ev_recode, ss));
    st = st_main;
    unsigned long long launches = nosync ? 6 : 7;
    if (!nosync) CK(cudaStreamSynchronize(st)); // the read-back: entries and longest bucket per lane size the rounds
    MsmStats stt;
    stt.window_bits = c;
    stt.windows = W;
    stt.lanes = NL;
    stt.tables = uniform ? 1 : 0;
    for (int l = 0; l < NL; l++) {
        Part &p = part[l];
        p.total = nosync ? total : h_lane_k[2 * l];
        p.maxlen = nosync ? (uint32_t)std::min<size_t>(total, 0xffffffffu) : h_lane_k[2 * l + 1];
        if (!nosync) stt.adds_total += p.total;
        MsmLane &L = lanes[l];
        const size_t nseg_max = std::max<size_t>(p.nseg, std::max(p.nseg_a, p.nseg_b)) + 1;
        const size_t ent_max = std::max<size_t>(p.total, std::max(p.nent_a, p.nent_b));
        for (int i = 0; i < 2; i++) {
            RS(L.seg_len[i], nseg_max * 4);
            RS(L.seg_start[i], nseg_max * 4);
        }
        RS(L.c_len, nseg_max * 4);
        RS(L.c_start, nseg_max * 4);
        RS(L.blk, (std::max<size_t>(PLAN_MAX_BLOCKS, acc_capacity) + 8) * 8);
        RS(L.acc_ctl, 2 * ACC_CTL_WORDS * 4);
        if (profile) RS(L.acc_times, 2 * ACC_TIME_WORDS * 8);
        if (!L.blk_flag.p) {
            RS(L.blk_flag, (PLAN_MAX_BLOCKS + 8) * 4);
            CK(cudaMemsetAsync(L.blk_flag.p, 0, (PLAN_MAX_BLOCKS + 8) * 4, L.stream));
            L.epoch = 0;
        }
        const size_t task_ub0 = ent_max / 2 + 1;           // round 0, the largest
        const size_t out_ub0 = ent_max / 2 + nseg_max + 1; // outputs of round 0 (ceil halves)
        RS(L.pp[0], out_ub0 * sizeof(AffPt));
        RS(L.pp[1], (out_ub0 / 2 + nseg_max + 1) * sizeof(AffPt));
        if ((rc = reserve_round(L, task_ub0)) != 0) return rc;
        RS(L.buckets, (size_t)p.nseg * sizeof(AffPt));
        RS(L.rc, (size_t)p.nseg_a * sizeof(LdPt));
        RS(L.ents2, (size_t)std::max(p.nent_a, p.nent_b) * 4);
    }
#undef RS
    for (int l = 0; l < NL; l++) {
        lanes[l].launches = 0;
        lanes[l].prof_used = 0;
    }
    if (plan_on_sort) {
        for (int l = 0; l < NL; l++) {
            Tree tree(*this, lanes[l]);
            if ((rc = tree.plan_launch(ss, lanes[l].plan_main[k], d_len_all + bounds[l], d_start_all + bounds[l],
                                       entries_k.as<uint32_t>(), part[l].nseg, part[l].total, ACC_MAX_ROUNDS)))
                return rc;
        }
        CK(cudaEventRecord(ev_recode, ss));
    }
    // a reduction level with more points than this starts with batched-affine rounds: large MSMs are throughput-bound
    // (5 instead of 15 multiplications per addition), small ones latency-bound (one launch instead of a round's six)
    // (persistent path: with the rounds planned ahead an affine round costs a barrier and an inversion, and five of them
    // before the tree are the optimum from 2^19 to 2^21 points -- profiles/r2s_ldmax_sweep.log)
    const size_t ld_max = ld_tree_max ? ld_tree_max
                          : (persistent_any && pre_on) ? (size_t)12000
                                                        : (n >= (1u << 21) ? (size_t)1 << 13 : (size_t)1 << 16);
    if (timing) cudaEventRecord(ev[3], st);
    // ---- per lane: accumulate buckets, then the two reduction levels into this lane's slice of hb
    for (int l = 0; l < NL; l++) {
        MsmLane &L = lanes[l];
        const Part &p = part[l];
        CK(cudaStreamWaitEvent(L.stream, ev_recode, 0));
        // sorted ahead: nothing on the context stream orders this MSM after the read-back of the previous one's
        // partial sums (hb is shared), so the lane waits for it itself
        if (side) CK(cudaStreamWaitEvent(L.stream, ev_t1[k ^ 1], 0));
        if (timing && l == 0) cudaEventRecord(L.ev_s[0], L.stream);
        Tree tree(*this, L);
        const uint32_t *len0 = d_len_all + bounds[l], *start0 = d_start_all + bounds[l];
        int r_main = 0, r_a = 0, r_b = 0;
        const bool persistent = persistent_any;
        const uint32_t acc_grid = std::max<uint32_t>(1, acc_capacity / (uint32_t)NL);
        if (persistent) CK(cudaMemsetAsync(L.acc_ctl.p, 0, 2 * ACC_CTL_WORDS * 4, L.stream)); // (a launch clears its own again)
        if (persistent && profile) CK(cudaMemsetAsync(L.acc_times.p, 0, 2 * ACC_TIME_WORDS * 8, L.stream));
        if (persistent) {
            // every round of the bucket accumulation in one persistent launch; the lanes' kernels share the SMs
            // (its plan first, unless the sort stream has made it: the timed bracket below is the kernel alone)
            if (pre_on && !plan_on_sort &&
                (rc = tree.plan_launch(L.stream, L.plan_main[k], len0, start0, entries_k.as<uint32_t>(), p.nseg, p.total,
                                       ACC_MAX_ROUNDS)))
                return rc;
            if (timing && l == 0) cudaEventRecord(L.ev_k[0], L.stream);
            if ((rc = tree.accumulate(d_points, entries_k.as<uint32_t>(), start0, len0, p.nseg, p.total, L.buckets.as<AffPt>(),
                                      -1, 0, acc_grid, pre_on ? &L.plan_main[k] : nullptr, /*plan_here=*/false)))
                return rc;
            if (timing && l == 0) cudaEventRecord(L.ev_k[1], L.stream);
            while ((1ull << r_main) < p.maxlen) r_main++;
        } else {
            if ((rc = tree.plan0(start0, len0, entries_k.as<uint32_t>(), p.nseg, d_points))) return rc;
            if (timing && l == 0) CK(cudaMemcpyAsync(L.info_r0.p, L.info.p, 16, cudaMemcpyDeviceToDevice, L.stream));
            L.want_k = timing && l == 0;
            rc = tree.rounds(d_points, entries_k.as<uint32_t>(), start0, len0, p.nseg, p.total, p.maxlen,
                             L.buckets.as<AffPt>(), &r_main);
            if (rc) return rc;
            L.want_k = false;
        }
        if (timing && l == 0) cudaEventRecord(L.ev_s[1], L.stream);
        uint32_t *d_start = L.c_start.as<uint32_t>(), *d_len = L.c_len.as<uint32_t>();
        const uint32_t tthr = std::max(32u, std::max(R, m) / 2); // a segment has at most max(R, m) points
        if (tthr <= 256) {
            // level A: row and column sums of each virtual window's bucket matrix (projective, one block per segment)
            k_gen_level_a<<<cdiv(p.nent_a, 256), 256, 0, L.stream>>>(p.vn, nbv, lm, L.ents2.as<uint32_t>());
            k_gen_segs_a<<<cdiv(p.nseg_a + 1, 256), 256, 0, L.stream>>>(p.vn, nbv, lm, d_start, d_len);
            // large levels start with batched-affine rounds (5 instead of 15 multiplications per addition) and
            // switch to the inversion-free tree once the work is latency-bound
            Tree::Partial part_a{ld_max, nullptr, nullptr, nullptr, 0};
            if (p.nent_a > ld_max && persistent) {
                // the segments of level A have fixed lengths, so the number of rounds is known here:
                // after r rounds at most nent_a / 2^r + nseg_a points remain
                const uint32_t mlen = std::max(R, m);
                int nr = 0, full = 0;
                while ((1u << full) < mlen) full++;
                while (nr < full && ((size_t)p.nent_a >> nr) + p.nseg_a > ld_max) nr++;
                if (nr > 0) {
                    // (fixed segment lengths: whether the rounds are planned ahead is known here, and with it where
                    // the tables of the list after the last round are)
                    const bool pre_a = pre_on && mlen <= (1u << PL_K);
                    // the index list and the segments of level A follow from the window layout alone, and so does its plan
                    const uint64_t key_a = ((uint64_t)p.vn << 48) ^ ((uint64_t)nbv << 16) ^ ((uint64_t)lm << 8) ^ (uint64_t)nr ^ (1ull << 63);
                    if ((rc = tree.accumulate(L.buckets.as<AffPt>(), L.ents2.as<uint32_t>(), d_start, d_len, p.nseg_a, p.nent_a,
                                              nullptr, nr, 1, acc_grid, pre_a ? &L.plan_a : nullptr, true, key_a)))
                        return rc;
                    part_a.src = L.pp[(nr - 1) & 1].as<AffPt>();
                    if (pre_a) {
                        part_a.start = L.plan_a.start.as<uint32_t>() + (size_t)nr * L.plan_a.stride;
                        part_a.len = L.plan_a.len.as<uint32_t>() + (size_t)nr * L.plan_a.stride;
                    } else {
                        part_a.start = L.seg_start[nr & 1].as<uint32_t>();
                        part_a.len = L.seg_len[nr & 1].as<uint32_t>();
                    }
                    part_a.maxlen = (mlen + (1u << nr) - 1) >> nr;
                }
                r_a = nr;
            } else if (p.nent_a > ld_max) {
                if ((rc = tree.plan0(d_start, d_len, L.ents2.as<uint32_t>(), p.nseg_a, L.buckets.as<AffPt>()))) return rc;
                rc = tree.rounds(L.buckets.as<AffPt>(), L.ents2.as<uint32_t>(), d_start, d_len, p.nseg_a, p.nent_a,
                                 std::max(R, m), nullptr, &r_a, &part_a);
                if (rc) return rc;
            }
            tree.pb(PC_MISC);
            if (r_a > 0 && part_a.maxlen <= 64 && ld_tree_warp_a) {
                // short segments: every addition by a whole warp, four pairs per warp at the first level (with a warp
                // per pair most warps of a block idle after that level and the block slots of an SM set the time)
                const uint32_t nw = std::max(1u, std::min(16u, part_a.maxlen / 8));
                k_ld_tree_warp<false, false><<<p.nseg_a, 32 * nw, std::max(1u, part_a.maxlen) * sizeof(LdPt), L.stream>>>(
                    part_a.src, nullptr, part_a.start, part_a.len, L.rc.p, msqr_tabs.as<gf>());
            } else if (r_a > 0) {
                const uint32_t th = std::max(32u, part_a.maxlen / 2);
                k_ld_tree<false, false><<<p.nseg_a, th, th * sizeof(LdPt), L.stream>>>(
                    part_a.src, nullptr, part_a.start, part_a.len, L.rc.p, msqr_tabs.as<gf>());
            } else {
                k_ld_tree<false, false><<<p.nseg_a, tthr, tthr * sizeof(LdPt), L.stream>>>(
                    L.buckets.p, L.ents2.as<uint32_t>(), d_start, d_len, L.rc.p, msqr_tabs.as<gf>());
            }
            tree.pe();
            // level B: per-bit subset sums of the row / column sums, converted to affine for the host tail
            k_gen_level_b<<<cdiv(p.nent_b, 256), 256, 0, L.stream>>>(p.vn, lr, lm, L.ents2.as<uint32_t>());
            k_gen_segs_b<<<cdiv(p.nseg_b + 1, 256), 256, 0, L.stream>>>(p.vn, lr, lm, d_start, d_len);
            tree.pb(PC_MISC);
            if (2 * tthr <= 8 * 16) // a segment has at most 2 tthr points: 4 pairs per warp and level with 16 warps
                k_ld_tree_warp<true, true><<<p.nseg_b, 512, 2 * tthr * sizeof(LdPt), L.stream>>>(
                    L.rc.p, L.ents2.as<uint32_t>(), d_start, d_len, hb.as<AffPt>() + (size_t)p.v0 * cv, msqr_tabs.as<gf>());
            else
                k_ld_tree<true, true><<<p.nseg_b, tthr, tthr * sizeof(LdPt), L.stream>>>(
                    L.rc.p, L.ents2.as<uint32_t>(), d_start, d_len, hb.as<AffPt>() + (size_t)p.v0 * cv, msqr_tabs.as<gf>());
            tree.pe();
            L.launches += 6;
            CK(cudaGetLastError());
            r_b = 1;
        } else {
            // very wide windows (forced): the same two levels as batched-affine tree rounds
            k_gen_level_a<<<cdiv(p.nent_a, 256), 256, 0, L.stream>>>(p.vn, nbv, lm, L.ents2.as<uint32_t>());
            k_gen_segs_a<<<cdiv(p.nseg_a + 1, 256), 256, 0, L.stream>>>(p.vn, nbv, lm, d_start, d_len);
            L.launches += 2;
            if ((rc = tree.plan0(d_start, d_len, L.ents2.as<uint32_t>(), p.nseg_a, L.buckets.as<AffPt>()))) return rc;
            rc = tree.rounds(L.buckets.as<AffPt>(), L.ents2.as<uint32_t>(), d_start, d_len, p.nseg_a, p.nent_a,
                             std::max(R, m), L.rc.as<AffPt>(), &r_a);
            if (rc) return rc;
            k_gen_level_b<<<cdiv(p.nent_b, 256), 256, 0, L.stream>>>(p.vn, lr, lm, L.ents2.as<uint32_t>());
            k_gen_segs_b<<<cdiv(p.nseg_b + 1, 256), 256, 0, L.stream>>>(p.vn, lr, lm, d_start, d_len);
            L.launches += 2;
            if ((rc = tree.plan0(d_start, d_len, L.ents2.as<uint32_t>(), p.nseg_b, L.rc.as<AffPt>()))) return rc;
            rc = tree.rounds(L.rc.as<AffPt>(), L.ents2.as<uint32_t>(), d_start, d_len, p.nseg_b, p.nent_b,
                             std::max(R >> 1, m), hb.as<AffPt>() + (size_t)p.v0 * cv, &r_b);
            if (rc) return rc;
        }
        if (timing && l == 0) cudaEventRecord(L.ev_s[2], L.stream);
        CK(cudaEventRecord(L.done, L.stream));
        stt.rounds_main = std::max(stt.rounds_main, r_main);
        stt.rounds_a = r_a;
        stt.rounds_b = r_b;
    }
    for (int l = 0; l < NL; l++) {
        CK(cudaStreamWaitEvent(st, lanes[l].done, 0));
        launches += lanes[l].launches;
    }
    CK(cudaMemcpyAsync(h_pts[k], hb.as<AffPt>(), hb_bytes, cudaMemcpyDeviceToHost, st));
    uint32_t *h_acc = h_lane_k + 64; // per lane: 2 control blocks x (abort flag, rounds, additions r0, additions)
    if (persistent_any)
        for (int l = 0; l < NL; l++)
            for (int q = 0; q < 2; q++) {
                const uint32_t *ctl = lanes[l].acc_ctl.as<uint32_t>() + q * ACC_CTL_WORDS;
                CK(cudaMemcpyAsync(h_acc + (2 * l + q) * 4, ctl + 1, 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(h_acc + (2 * l + q) * 4 + 1, ctl + 64, 12, cudaMemcpyDeviceToHost, st));
            }
    if (timing) {
        if (!persistent_any) CK(cudaMemcpyAsync(h_lane_k, lanes[0].info_r0.p, 16, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(ev[1], st);
    }
    CK(cudaEventRecord(ev_t1[k], st));
    P->active = true;
    P->uniform = uniform;
    P->persistent_any = persistent_any;
    P->nosync = nosync;
    P->k = k;
    P->c = c;
    P->cv = cv;
    P->base = base;
    P->rem = rem;
    P->NL = NL;
    P->V = V;
    P->nbv = nbv;
    P->launches = launches;
    P->stt = stt;
    return 0;
}

int MsmEngine::finish(Pending &P, AffPt *h_result) {
    *h_result = pt_inf();
    if (!P.active) return 0;
    P.active = false;
    const int k = P.k, NL = P.NL, c = P.c, cv = P.cv, base = P.base, rem = P.rem;
    const bool uniform = P.uniform, persistent_any = P.persistent_any, nosync = P.nosync;
    const uint32_t V = P.V, nbv = P.nbv;
    const unsigned long long launches = P.launches;
    MsmStats stt = P.stt;
    cudaStream_t st = stream;
    uint32_t *const h_lane_k = (uint32_t *)h_lane + 128 * k;
    uint32_t *h_acc = h_lane_k + 64;
    CK(cudaEventSynchronize(ev_t1[k]));
    if (persistent_any) {
        int rmax = 0;
        unsigned long long adds = 0;
        for (int l = 0; l < NL; l++) {
            rmax = std::max(rmax, (int)h_acc[(2 * l) * 4 + 1]);
            adds += h_acc[(2 * l) * 4 + 3];
        }
        stt.rounds_main = rmax;
        if (nosync) stt.adds_total = adds;
    }
    if (persistent_any)
        for (int l = 0; l < NL; l++)
            for (int q = 0; q < 2; q++)
                if (h_acc[(2 * l + q) * 4]) {
                    fprintf(stderr, "[dvpari] k_accumulate gave up at a grid barrier (lane %d, launch %d)\n", l, q);
                    return DVP_ERR_INTERNAL;
                }
    float ms_dev = 0;
    cudaEventElapsedTime(&ms_dev, ev_t0[k], ev_t1[k]);

    // ---- host tail.  Virtual window v holds buckets of digit values dbase_v + b + 1 (b local) at bit offset
    // off_v: sum_b (dbase_v + b + 1) B_b = sum_q 2^q H[v][q] + (dbase_v + 1) S[v]; one double-and-add pass
    // over all positions.  (Separate sets: off_v = the window's offset, dbase_v = 0.  Shared set: off_v = 0.)
    const auto t_host0 = std::chrono::steady_clock::now();
    {
        const AffPt *hp = (const AffPt *)h_pts[k];
        const int npos = uniform ? c + 1 : 233 + c + 1;
        std::vector<std::vector<uint32_t>> at(npos + 1);
        for (uint32_t v = 0; v < V; v++) {
            const int off = uniform ? 0 : (int)v * base + std::min((int)v, rem);
            const uint32_t dbase = uniform ? v * nbv : 0;
            for (int q = 0; q < cv - 1; q++) at[off + q].push_back(v * cv + q);
            const uint32_t ws = dbase + 1;
            for (int t = 0; t < 32; t++)
                if ((ws >> t) & 1) at[off + t].push_back(v * cv + cv - 1);
        }
        host::LdPt acc = host::ld_inf();
        for (int pos = npos; pos >= 0; pos--) {
            acc = host::ld_dbl(acc);
            for (uint32_t idx : at[pos]) acc = host::ld_add_affine(acc, hp[idx]);
        }
        *h_result = host::ld_to_affine(acc);
    }
    stt.launches = launches;
    stt.ms_device = ms_dev;
    stt.ms_tail_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    if (profile) {
        for (int i = 0; i < PC_COUNT; i++) prof_ms[i] = 0, prof_n[i] = 0;
        timeline.clear();
        for (int l = 0; l < NL; l++)
            for (size_t i = 0; i < lanes[l].prof_used; i++) {
                const auto &r = lanes[l].prof[i];
                float ms = 0, t0 = 0;
                cudaEventElapsedTime(&ms, r.e0, r.e1);
                cudaEventElapsedTime(&t0, ev_t0[k], r.e0);
                prof_ms[r.cat] += ms;
                prof_n[r.cat]++;
                timeline.push_back({(float)l, (float)r.cat, t0, t0 + ms}); // Gantt row: lane, category, start, end (ms)
            }
        // phases inside the persistent kernels: globaltimer of block 0 after every grid barrier, relative to the kernel's
        // own start (rows with lane = 100 + 10 lane + launch; categories as above, plan step 1 / 2 both as plan)
        if (persistent_any) {
            std::vector<unsigned long long> tm(2 * ACC_TIME_WORDS);
            for (int l = 0; l < NL; l++) {
                if (cudaMemcpy(tm.data(), lanes[l].acc_times.p, tm.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) break;
                for (int k = 0; k < 2; k++) {
                    const unsigned long long *t = tm.data() + k * ACC_TIME_WORDS, ts = t[ACC_TIME_WORDS - 1];
                    if (!ts) continue;
                    unsigned long long prev = ts;
                    const int cats[5] = {PC_PLAN, PC_PLAN, PC_PASS1, PC_BINV_DIRECT, PC_PASS2};
                    for (int r = 0; r <= ACC_MAX_ROUNDS; r++)
                        for (int ph = 0; ph < 6; ph++) {
                            const unsigned long long v = t[6 * r + ph];
                            if (!v) continue;
                            timeline.push_back({(float)(100 + 10 * l + k), (float)(ph < 5 ? cats[ph] : PC_MISC),
                                                (float)((prev - ts) * 1e-6), (float)((v - ts) * 1e-6)});
                            prev = v;
                        }
                }
            }
        }
    }
    if (timing) {
        cudaEventRecord(ev[2], st);
        cudaEventSynchronize(ev[2]);
        MsmLane &L0 = lanes[0];
        cudaEventElapsedTime(&stt.ms_recode_sort, ev[0], ev[3]);
        cudaEventElapsedTime(&stt.ms_accumulate, L0.ev_s[0], L0.ev_s[1]);
        cudaEventElapsedTime(&stt.ms_reduce, L0.ev_s[1], L0.ev_s[2]);
        cudaEventElapsedTime(&stt.ms_tail, ev[1], ev[2]);
        if (stt.rounds_main > 0) cudaEventElapsedTime(&stt.ms_pass2_round0, L0.ev_k[0], L0.ev_k[1]);
        if (persistent_any) {
            // the dominant kernel is lane 0's k_accumulate: every round of its buckets in one launch
            stt.adds_round0 = h_acc[3];
            stt.rounds_main = (int)h_acc[1];
        } else {
            stt.adds_round0 = h_lane_k[1]; // info of lane 0's first plan
        }
    }
    last = stt;
    return 0;
}

} // namespace dvp
