// Fr-side prover kernels and the device-resident Proof::prove for sm_100a.
//
// Replaces, on the prover path of /root/reference/src/proving.rs:426-688:
//   build_sect_ecfft_tree / get_both_domains      src/ec_fft.rs:93-239,179-189   -> dvp_domain_create
//   FFTree::extend(.., Moiety::S1) x4             src/proving.rs:410-422         -> k_extend_down/up
//   get_matrix_evaluations_from_witness           src/proving.rs:348-403         -> k_r1cs_eval
//   r / q / K-scalar pointwise passes             src/proving.rs:492-509,599-654 -> k_quotient, k_kscalars
//   evaluate_poly_at_alpha_using_barycentric_..   src/ec_fft.rs:455-491          -> k_bary_partial
//   prover precomputes (bar_wts, z_vals2inv)      src/proving.rs:225-325         -> chain-rule kernels
// Fr is 8 x u32 Montgomery limbs in ark's memory layout (fr.cuh).  Every vector stays in HBM between
// stages; the host sees only the two commitments, alpha and the two evaluations.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/dvpari.h"
#include "ctx.cuh"
#include "fr.cuh"
#include "fr29.cuh"
#include "host_gf.hpp"
#include "transcript_host.hpp"

using namespace dvp;

#define CKP(x)                                                                                            \
    do {                                                                                                  \
        cudaError_t e_ = (x);                                                                             \
        if (e_ != cudaSuccess) {                                                                          \
            fprintf(stderr, "[dvpari] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DVP_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

static inline uint32_t cdivp(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ fr fr_load(const fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_store(fr *p, const fr &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
__device__ __forceinline__ fr29 fr29_load(const fr *p) {
    const fr t = fr_load(p);
    fr29 r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.l[k] = t.v[k];
    return r;
}
__device__ __forceinline__ void fr29_store(fr *p, const fr29 &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// v^(2^e)
__host__ __device__ inline fr fr_pow2k(fr v, int e) {
    for (int i = 0; i < e; i++) v = fr_sqr(v);
    return v;
}

// ------------------------------------------------------------------------------------------------
// domain construction
// ------------------------------------------------------------------------------------------------
struct SwPt {
    fr x, y;
};
struct SwConsts {
    fr A;
    SwPt C;        // coset offset
    SwPt pow2[28]; // pow2[b] = 2^b * g  (g = generator of the order-N subgroup)
};

// leaf_i = x(C + i g): Jacobian accumulation over the set bits of i, one inversion per leaf.
__global__ void k_dom_leaves(const SwConsts *__restrict__ sc, uint32_t N, int logN, fr *__restrict__ leaves) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    // (X, Y, Z) with x = X/Z^2, y = Y/Z^3, start from the affine coset point
    fr X = sc->C.x, Y = sc->C.y, Z = fr_one();
    for (int b = 0; b < logN; b++) {
        if (!((i >> b) & 1)) continue;
        // mixed addition (X,Y,Z) + (x2,y2); the operands are never equal or opposite (C is off the subgroup)
        const fr x2 = sc->pow2[b].x, y2 = sc->pow2[b].y;
        const fr Z2 = fr_sqr(Z);
        const fr U2 = fr_mul(x2, Z2), S2 = fr_mul(y2, fr_mul(Z2, Z));
        const fr H = fr_sub(U2, X), Rr = fr_sub(S2, Y);
        const fr H2 = fr_sqr(H), H3 = fr_mul(H2, H), V = fr_mul(X, H2);
        fr X3 = fr_sub(fr_sub(fr_sqr(Rr), H3), fr_add(V, V));
        fr Y3 = fr_sub(fr_mul(Rr, fr_sub(V, X3)), fr_mul(Y, H3));
        Z = fr_mul(Z, H);
        X = X3;
        Y = Y3;
    }
    const fr zi = fr_inv(Z);
    fr_store(&leaves[i], fr_mul(X, fr_sqr(zi)));
}

// next layer: out[i] = psi(in[i]) = in[i] + t/(in[i] - x0), i < half
__global__ void k_dom_next_layer(const fr *__restrict__ in, uint32_t half, fr x0, fr t, fr *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const fr x = fr_load(&in[i]);
    fr_store(&out[i], fr_add(x, fr_mul(t, fr_inv(fr_sub(x, x0)))));
}

// v^(h-1) with h = 2^e:  prod_{i<e} v^(2^i)
__device__ inline fr pow_hm1(fr b, int e) {
    fr acc = fr_one();
    for (int i = 0; i < e; i++) {
        acc = fr_mul(acc, b);
        b = fr_sqr(b);
    }
    return acc;
}

// Level tables for sub-problems of size m = 2h on layer L (2m points): pair j uses the even leaves (2j, 2j+m) as sources
// and the odd leaves (2j+1, 2j+1+m) as targets,  P(s) = (P0(psi(s)) + s P1(psi(s))) v(s)^(h-1),  v(x) = x - x0.
// The plain matrices are  recombine = diag(vt0, vt1) (1 t0; 1 t1)  and  decompose = ((1 s0; 1 s1))^-1 diag(1/vs0, 1/vs1).
// Their diagonal factors depend on the position only, never on the sub-problem, so they are carried along as a scale
// per position instead of being multiplied in at every level: the values that flow through the levels are
// P~ = P / sigma on the way down and P~ = P / tau on the way up,
//   down:  sigma'(j) = -sigma(s0) / (vs0 (s1 - s0)),  beta = -sigma(s1) vs0 / (sigma(s0) vs1),
//          P0~ = -s1 x0~ - s0 beta x1~,  P1~ = x0~ + beta x1~                      (3 products instead of 4)
//   up:    tau(t_i) = vt_i tau'(psi(t_i)),  P~(t_i) = P0~ + t_i P1~                (2 products instead of 4)
// and the scale that is left, tau of level 0, is folded into the last recombine level (a plain 2x2 matrix).  The
// results are the same field elements (exact arithmetic, canonical encodings), 5 products per butterfly pair for 8.
// Table rows keep 4 entries per butterfly: down (-s1, -s0 beta, -, beta), up (-, t0, -, t1), top (tau0, tau0 t0, tau1, tau1 t1).
__global__ void k_dom_down(const fr *__restrict__ L, uint32_t h, int e, fr x0, const fr *__restrict__ sig,
                           fr *__restrict__ sig_next, fr *__restrict__ dec) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= h) return;
    const uint32_t m = 2 * h;
    const fr s0 = fr_load(&L[2 * j]), s1 = fr_load(&L[2 * j + m]);
    const fr vs0 = pow_hm1(fr_sub(s0, x0), e), vs1 = pow_hm1(fr_sub(s1, x0), e);
    const fr g0 = sig ? fr_load(&sig[j]) : fr_one(), g1 = sig ? fr_load(&sig[j + h]) : fr_one();
    // one inversion for 1/(vs0 (s1 - s0)) and 1/(g0 vs1)
    const fr A = fr_mul(vs0, fr_sub(s1, s0)), B = fr_mul(g0, vs1);
    const fr iab = fr_inv(fr_mul(A, B));
    const fr u0 = fr_mul(g0, fr_mul(iab, B));
    const fr beta = fr_neg(fr_mul(fr_mul(g1, vs0), fr_mul(iab, A)));
    fr_store(&sig_next[j], fr_neg(u0));
    fr_store(&dec[4 * j + 0], fr_neg(s1));
    fr_store(&dec[4 * j + 1], fr_neg(fr_mul(s0, beta)));
    fr_store(&dec[4 * j + 2], fr_zero());
    fr_store(&dec[4 * j + 3], beta);
}
__global__ void k_dom_up(const fr *__restrict__ L, uint32_t h, int e, fr x0, const fr *__restrict__ tau_next,
                         fr *__restrict__ tau, int top, fr *__restrict__ rec) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= h) return;
    const uint32_t m = 2 * h;
    const fr t0 = fr_load(&L[2 * j + 1]), t1 = fr_load(&L[2 * j + 1 + m]);
    const fr tn = fr_load(&tau_next[j]);
    const fr a0 = fr_mul(pow_hm1(fr_sub(t0, x0), e), tn), a1 = fr_mul(pow_hm1(fr_sub(t1, x0), e), tn);
    fr_store(&tau[j], a0);
    fr_store(&tau[j + h], a1);
    fr_store(&rec[4 * j + 0], top ? a0 : fr_zero());
    fr_store(&rec[4 * j + 1], top ? fr_mul(a0, t0) : t0);
    fr_store(&rec[4 * j + 2], top ? a1 : fr_zero());
    fr_store(&rec[4 * j + 3], top ? fr_mul(a1, t1) : t1);
}

// chain rule, one level: for the m points s_j = L[2j + shift] of S^k
//   deriv:  out[j] = v(s)^(m/2) psi'(s) prev[j mod h]      (Z'_S on its roots)
//   other:  out[j] = v(t)^(m/2) prev[j mod h], t = L[2j + 1 - shift]   (Z_S on the other half-domain)
__global__ void k_chain_level(const fr *__restrict__ L, uint32_t m, int e2, fr x0, fr t, int shift, int deriv,
                              const fr *__restrict__ prev, fr *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t h = m >> 1;
    const fr s = fr_load(&L[2 * j + (deriv ? shift : 1 - shift)]);
    const fr v = fr_sub(s, x0);
    fr f = fr_pow2k(v, e2);
    if (deriv) {
        const fr vi = fr_inv(v);
        f = fr_mul(f, fr_sub(fr_one(), fr_mul(t, fr_sqr(vi))));
    }
    fr_store(&out[j], fr_mul(f, fr_load(&prev[h ? j % h : 0])));
}
__global__ void k_fr_inv_each(fr *__restrict__ v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_store(&v[i], fr_inv(fr_load(&v[i])));
}

// ------------------------------------------------------------------------------------------------
// ECFFT extend: in-place 2x2 butterflies, level k has sub-problems of size m = n >> k
// ------------------------------------------------------------------------------------------------
// The matrices are stored as 29-bit limbs, pre-scaled by 2^-24 (fr29.cuh): a butterfly is two dot products with one
// Montgomery reduction each, every partial product an IMAD.WIDE with its 64-bit accumulate.
__global__ void k_mats_to29(fr *__restrict__ mats, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const fr29 r = fr29_prescale(fr_load(&mats[i]));
    fr o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = r.l[k];
    fr_store(&mats[i], o);
}
// butterfly forms (k_dom_down / k_dom_up): plain 2x2 (the top recombine level), decompose (one dot product and one
// multiply-add), recombine (two multiply-adds)
enum { BF_GEN = 0, BF_DOWN = 1, BF_UP = 2 };
template <int MODE>
__device__ __forceinline__ void butterfly29(const fr29 &m0, const fr29 &m1, const fr29 &m2, const fr29 &m3, const fr29 &a0,
                                            const fr29 &a1, fr29 &y0, fr29 &y1) {
    // the values between the levels are semi-reduced (< 2^232, fr29.cuh); the plain top level reduces fully
    if (MODE == BF_GEN) {
        const fr29 x[2] = {a0, a1}, r0[2] = {m0, m1}, r1[2] = {m2, m3};
        y0 = fr29_dotn<2>(r0, x);
        y1 = fr29_dotn<2>(r1, x);
    } else if (MODE == BF_DOWN) {
        y0 = fr29_dot2_semi(m0, a0, m1, a1);
        y1 = fr29_muladd_semi(m3, a1, a0);
    } else {
        y0 = fr29_muladd_semi(m1, a1, a0);
        y1 = fr29_muladd_semi(m3, a1, a0);
    }
}
// Between the first decompose level and the last recombine level the vectors stay on 29-bit limbs in global memory too
// (IN29 / OUT29): only the two ends of an extend convert from / to the 32-bit words of the ABI.
template <int NP, int MODE, bool IN29, bool OUT29>
__global__ void __launch_bounds__(256)
    k_extend_level(fr *__restrict__ data, uint32_t n, uint32_t h, const fr *__restrict__ mats, int npoly,
                   size_t stride) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; // butterfly index in [0, n/2)
    if (g >= (n >> 1)) return;
    const uint32_t j = g % h, s = g / h;
    const uint32_t i0 = s * 2 * h + j, i1 = i0 + h;
    // all loads of the thread are issued before the first product (NP polynomials share the matrix)
    fr x0[NP], x1[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < npoly) {
            x0[p] = fr_load(&data[(size_t)p * stride + i0]);
            x1[p] = fr_load(&data[(size_t)p * stride + i1]);
        }
    }
    fr29 m0, m1, m2, m3;
    if (MODE != BF_UP) m0 = fr29_load(&mats[4 * j]);
    m1 = fr29_load(&mats[4 * j + 1]);
    if (MODE == BF_GEN) m2 = fr29_load(&mats[4 * j + 2]);
    m3 = fr29_load(&mats[4 * j + 3]);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < npoly) {
            fr29 a0, a1;
            if (IN29) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    a0.l[q] = x0[p].v[q];
                    a1.l[q] = x1[p].v[q];
                    // fully reduced limbs; without the range the compiler widens the products (IMAD + IADD3 per term)
                    __builtin_assume(a0.l[q] <= DVP_M29);
                    __builtin_assume(a1.l[q] <= DVP_M29);
                }
            } else {
                a0 = fr29_from_fr(x0[p]);
                a1 = fr29_from_fr(x1[p]);
            }
            fr29 y0, y1;
            butterfly29<MODE>(m0, m1, m2, m3, a0, a1, y0, y1);
            if (OUT29) {
                fr29_store(&data[(size_t)p * stride + i0], y0);
                fr29_store(&data[(size_t)p * stride + i1], y1);
            } else {
                fr_store(&data[(size_t)p * stride + i0], fr_from_fr29(y0));
                fr_store(&data[(size_t)p * stride + i1], fr_from_fr29(y1));
            }
        }
    }
}
// one level of the extend of tree d over `len` points: decompose (down) or recombine level k; these are the levels above
// the fused ones, so level 0 going down reads 32-bit words and level 0 coming up (the plain matrices) writes them
static void extend_level_launch(fr *data, uint32_t len, uint32_t h, const fr *mats, int np, size_t stride, bool down, int k,
                                cudaStream_t st) {
    const dim3 grid(cdivp(len / 2, 256));
    if (down && k == 0) k_extend_level<3, BF_DOWN, false, true><<<grid, 256, 0, st>>>(data, len, h, mats, np, stride);
    else if (down) k_extend_level<3, BF_DOWN, true, true><<<grid, 256, 0, st>>>(data, len, h, mats, np, stride);
    else if (k == 0) k_extend_level<3, BF_GEN, true, false><<<grid, 256, 0, st>>>(data, len, h, mats, np, stride);
    else k_extend_level<3, BF_UP, true, true><<<grid, 256, 0, st>>>(data, len, h, mats, np, stride);
}

// ---- one output per butterfly: the levels ABOVE a rank's block in the sharded extend (extend_device_sharded).
// Going down, only the half of each sub-problem that contains the rank's block is kept (h outputs of a level instead of
// 2 h); coming back up, only the positions that lead to the rank's own range of the result are formed.
template <int NP, bool IN29>
__global__ void __launch_bounds__(256)
    k_extend_down_sel(fr *__restrict__ data, uint32_t h, const fr *__restrict__ mats, int npoly, size_t stride, int which) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; // butterfly of the one sub-problem at `data`
    if (j >= h) return;
    const uint32_t i0 = j, i1 = j + h;
    fr x0[NP], x1[NP];
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < npoly) {
            x0[p] = fr_load(&data[(size_t)p * stride + i0]);
            x1[p] = fr_load(&data[(size_t)p * stride + i1]);
        }
    const fr29 m0 = fr29_load(&mats[4 * j]), m1 = fr29_load(&mats[4 * j + 1]), m3 = fr29_load(&mats[4 * j + 3]);
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < npoly) {
            fr29 a0, a1;
            if (IN29) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    a0.l[q] = x0[p].v[q];
                    a1.l[q] = x1[p].v[q];
                    __builtin_assume(a0.l[q] <= DVP_M29);
                    __builtin_assume(a1.l[q] <= DVP_M29);
                }
            } else {
                a0 = fr29_from_fr(x0[p]);
                a1 = fr29_from_fr(x1[p]);
            }
            // the decompose butterfly (butterfly29<BF_DOWN>): y0 = m0 a0 + m1 a1 at i0, y1 = a0 + m3 a1 at i1
            if (which == 0) fr29_store(&data[(size_t)p * stride + i0], fr29_dot2_semi(m0, a0, m1, a1));
            else fr29_store(&data[(size_t)p * stride + i1], fr29_muladd_semi(m3, a1, a0));
        }
}
// level with sub-problems of 2 h points over n points; of every sub-problem only the `blk` positions starting at q0
// (inside its lower or its upper half) are formed.  TOP: the plain 2x2 matrices of level 0, 32-bit words out.
template <int NP, bool TOP>
__global__ void __launch_bounds__(256)
    k_extend_up_sel(fr *__restrict__ data, uint32_t n, uint32_t h, uint32_t blk, uint32_t q0, const fr *__restrict__ mats,
                    int npoly, size_t stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n / (2 * h)) * blk) return;
    const uint32_t c = t / blk, o = t % blk;
    const bool upper = q0 >= h;
    const uint32_t j = (upper ? q0 - h : q0) + o; // butterfly of its sub-problem
    const uint32_t i0 = c * 2 * h + j, i1 = i0 + h;
    fr x0[NP], x1[NP];
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < npoly) {
            x0[p] = fr_load(&data[(size_t)p * stride + i0]);
            x1[p] = fr_load(&data[(size_t)p * stride + i1]);
        }
    fr29 m0, m1;
    if (TOP) {
        m0 = fr29_load(&mats[4 * j + (upper ? 2 : 0)]);
        m1 = fr29_load(&mats[4 * j + (upper ? 3 : 1)]);
    } else {
        m1 = fr29_load(&mats[4 * j + (upper ? 3 : 1)]);
    }
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < npoly) {
            fr29 a0, a1;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                a0.l[q] = x0[p].v[q];
                a1.l[q] = x1[p].v[q];
                __builtin_assume(a0.l[q] <= DVP_M29);
                __builtin_assume(a1.l[q] <= DVP_M29);
            }
            fr *out = &data[(size_t)p * stride + (upper ? i1 : i0)];
            if (TOP) {
                const fr29 x[2] = {a0, a1}, r[2] = {m0, m1};
                fr_store(out, fr_from_fr29(fr29_dotn<2>(r, x)));
            } else {
                fr29_store(out, fr29_muladd_semi(m1, a1, a0)); // butterfly29<BF_UP>: y = a0 + t a1
            }
        }
}

// The deepest levels of an extend work on sub-problems of at most EXT_FUSE_M points: one block stages a whole
// sub-problem in shared memory (as 29-bit limbs, converted once on the way in and once on the way out), runs its
// decompose levels down and its recombine levels up with a block barrier between levels, and writes it back -- one
// launch and one round trip to HBM instead of two per level.
constexpr int EXT_FUSE_LOG = 11;
constexpr uint32_t EXT_FUSE_M = 1u << EXT_FUSE_LOG;
struct FusedMats {
    const fr *dec[EXT_FUSE_LOG], *rec[EXT_FUSE_LOG];
};
__global__ void __launch_bounds__(256)
    k_extend_fused(fr *__restrict__ data, int logm, int logc, FusedMats fm, uint32_t blocks_per_poly, size_t poly_stride,
                   int top) { // top: the first fused level is level 0 of the tree (its recombine matrices are plain)
    extern __shared__ __align__(16) unsigned char ext_sm_raw[];
    fr *sm = reinterpret_cast<fr *>(ext_sm_raw); // fr29 limbs stored in fr-sized slots
    // a block stages C = 2^logc points = C / 2^logm whole sub-problems of size 2^logm (they tile the chunk)
    const uint32_t M = 1u << logc;
    fr *base = data + (size_t)(blockIdx.x / blocks_per_poly) * poly_stride + (size_t)(blockIdx.x % blocks_per_poly) * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        // top: the vectors arrive as 32-bit words; otherwise the levels above left them on 29-bit limbs
        const fr29 v = top ? fr29_from_fr(fr_load(&base[i])) : fr29_load(&base[i]);
        fr o;
#pragma unroll
        for (int k = 0; k < 8; k++) o.v[k] = v.l[k];
        fr_store(&sm[i], o);
    }
    __syncthreads();
    for (int pass = 0; pass < 2 * logm; pass++) {
        const int k = pass < logm ? pass : 2 * logm - 1 - pass; // levels 0 .. logm-1 down, then back up
        const fr *mats = pass < logm ? fm.dec[k] : fm.rec[k];
        const uint32_t h = (1u << logm) >> (k + 1);
        const int mode = pass < logm ? BF_DOWN : (k == 0 && top) ? BF_GEN : BF_UP; // uniform over the block
        for (uint32_t g = threadIdx.x; g < (M >> 1); g += blockDim.x) {
            const uint32_t j = g % h, i0 = (g / h) * 2 * h + j, i1 = i0 + h;
            const fr29 m1 = fr29_load(&mats[4 * j + 1]), m3 = fr29_load(&mats[4 * j + 3]);
            const fr29 x0 = fr29_load(&sm[i0]), x1 = fr29_load(&sm[i1]);
            fr29 y0, y1;
            if (mode == BF_UP) {
                butterfly29<BF_UP>(m1, m1, m3, m3, x0, x1, y0, y1);
            } else if (mode == BF_DOWN) {
                const fr29 m0 = fr29_load(&mats[4 * j]);
                butterfly29<BF_DOWN>(m0, m1, m3, m3, x0, x1, y0, y1);
            } else {
                const fr29 m0 = fr29_load(&mats[4 * j]), m2 = fr29_load(&mats[4 * j + 2]);
                butterfly29<BF_GEN>(m0, m1, m2, m3, x0, x1, y0, y1);
            }
            fr o0, o1;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                o0.v[q] = y0.l[q];
                o1.v[q] = y1.l[q];
            }
            fr_store(&sm[i0], o0);
            fr_store(&sm[i1], o1);
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        if (top) fr_store(&base[i], fr_from_fr29(fr29_load(&sm[i])));
        else fr_store(&base[i], fr_load(&sm[i]));
    }
}

// ------------------------------------------------------------------------------------------------
// R1CS rows: a = A w, b = B w, c = C w - i, i = sum_j x_j d^j; first unsatisfied row recorded
// ------------------------------------------------------------------------------------------------
struct R1csDev {
    const uint32_t *rowptr[3];
    const uint32_t *wire[3];
    const uint32_t *coeff[3];
    const fr *coeffs;
    uint32_t nrows, n, k;
};
__device__ __forceinline__ fr row_dot(const R1csDev &r, int which, uint32_t row, const fr *__restrict__ w) {
    fr acc = fr_zero();
    for (uint32_t p = r.rowptr[which][row]; p < r.rowptr[which][row + 1]; p++)
        acc = fr_add(acc, fr_mul(fr_load(&r.coeffs[r.coeff[which][p]]), fr_load(&w[r.wire[which][p]])));
    return acc;
}
__global__ void __launch_bounds__(128)
    k_r1cs_eval(R1csDev r, uint32_t row_lo, uint32_t row_hi, const fr *__restrict__ w, const fr *__restrict__ leaves,
                fr *__restrict__ a, fr *__restrict__ b, fr *__restrict__ c, fr *__restrict__ iv,
                unsigned long long *__restrict__ first_bad) {
    const uint32_t row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    fr av = fr_zero(), bv = fr_zero(), cw = fr_zero();
    if (row < r.nrows) {
        av = row_dot(r, 0, row, w);
        bv = row_dot(r, 1, row, w);
        cw = row_dot(r, 2, row, w);
    }
    const fr d = fr_load(&leaves[2 * row]);
    fr pw = fr_one(), ival = fr_zero();
    for (uint32_t j = 0; j < r.k; j++) {
        ival = fr_add(ival, fr_mul(fr_load(&w[1 + j]), pw));
        pw = fr_mul(pw, d);
    }
    fr_store(&a[row], av);
    fr_store(&b[row], bv);
    fr_store(&c[row], fr_sub(cw, ival));
    fr_store(&iv[row], ival);
    if (!fr_eq(fr_mul(av, bv), cw)) atomicMin(first_bad, (unsigned long long)row);
}

// ---- row products on 29-bit limbs (large circuits) ------------------------------------------------------------
// One thread per (row, matrix) task; the tasks of every chunk of R1CS_CHUNK rows are ordered by term count, so the
// lanes of a warp run the same number of iterations (the geometric row lengths cost ~3x in divergence otherwise).
// Coefficients are stored pre-scaled on 29-bit limbs, the assignment is converted once; three terms share one
// Montgomery reduction (fr29_dotn) and the partial sums are added on the limbs.
constexpr uint32_t R1CS_CHUNK = 4096;
__global__ void k_fr_to29(const fr *__restrict__ in, uint32_t n, fr *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr29 r = fr29_from_fr(fr_load(&in[i]));
    fr o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = r.l[k];
    fr_store(&out[i], o);
}
__global__ void __launch_bounds__(128)
    k_r1cs_sides(R1csDev r, const uint32_t *__restrict__ tasks, uint32_t task_lo, uint32_t task_hi,
                 const fr *__restrict__ coeffs29, const fr *__restrict__ w29, fr *__restrict__ a, fr *__restrict__ b,
                 fr *__restrict__ c) {
    const uint32_t ti = task_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= task_hi) return;
    const uint32_t task = tasks[ti], which = task & 3u, row = task >> 2;
    const uint32_t *wire = r.wire[which], *cid = r.coeff[which];
    uint32_t p = r.rowptr[which][row];
    const uint32_t pe = r.rowptr[which][row + 1];
    fr29 acc;
#pragma unroll
    for (int k = 0; k < 8; k++) acc.l[k] = 0;
    for (; p + 3 <= pe; p += 3) {
        fr29 m[3], x[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            m[t] = fr29_load(&coeffs29[cid[p + t]]);
            x[t] = fr29_load(&w29[wire[p + t]]);
        }
        acc = fr29_add(acc, fr29_dotn<3>(m, x));
    }
    if (pe - p == 2) {
        fr29 m[2], x[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            m[t] = fr29_load(&coeffs29[cid[p + t]]);
            x[t] = fr29_load(&w29[wire[p + t]]);
        }
        acc = fr29_add(acc, fr29_dotn<2>(m, x));
    } else if (pe - p == 1) {
        const fr29 m = fr29_load(&coeffs29[cid[p]]), x = fr29_load(&w29[wire[p]]);
        acc = fr29_add(acc, fr29_dotn<1>(&m, &x));
    }
    fr *out = which == 0 ? a : which == 1 ? b : c;
    fr_store(&out[row], fr_from_fr29(acc));
}
// per row: i = sum_j x_j d^j, c = C w - i, the satisfiability check; rows beyond the dump are zero
__global__ void __launch_bounds__(128)
    k_r1cs_finish(R1csDev r, uint32_t row_lo, uint32_t row_hi, const fr *__restrict__ w, const fr *__restrict__ leaves,
                  fr *__restrict__ a, fr *__restrict__ b, fr *__restrict__ c, fr *__restrict__ iv,
                  unsigned long long *__restrict__ first_bad) {
    const uint32_t row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    fr av = fr_zero(), bv = fr_zero(), cw = fr_zero();
    if (row < r.nrows) {
        av = fr_load(&a[row]);
        bv = fr_load(&b[row]);
        cw = fr_load(&c[row]);
    } else {
        fr_store(&a[row], av);
        fr_store(&b[row], bv);
    }
    const fr d = fr_load(&leaves[2 * row]);
    fr pw = fr_one(), ival = fr_zero();
    for (uint32_t j = 0; j < r.k; j++) {
        ival = fr_add(ival, fr_mul(fr_load(&w[1 + j]), pw));
        pw = fr_mul(pw, d);
    }
    fr_store(&c[row], fr_sub(cw, ival));
    fr_store(&iv[row], ival);
    if (!fr_eq(fr_mul(av, bv), cw)) atomicMin(first_bad, (unsigned long long)row);
}

// Synthetic circuits (synth.py): every row's O side ends with the row's own fresh wire 1 + k + row with
// coefficient one and a row only reads fresh wires of lower levels (level = row mod nlevels), so one pass
// per level makes every row hold:  w[fresh] += a b - (C w).
__global__ void __launch_bounds__(128)
    k_r1cs_solve_level(R1csDev r, fr *__restrict__ w, uint32_t level, uint32_t nlevels) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t row64 = (uint64_t)idx * nlevels + level;
    if (row64 >= r.nrows) return;
    const uint32_t row = (uint32_t)row64;
    const fr av = row_dot(r, 0, row, w), bv = row_dot(r, 1, row, w), cw = row_dot(r, 2, row, w);
    fr *slot = &w[1 + r.k + row];
    fr_store(slot, fr_add(fr_load(slot), fr_sub(fr_mul(av, bv), cw)));
}

// i(X) has degree k-1 < n, so its extension to D' is its evaluation there
__global__ void k_ivals_ext(const fr *__restrict__ w, uint32_t k, const fr *__restrict__ leaves, uint32_t n,
                            fr *__restrict__ i2) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const fr d = fr_load(&leaves[2 * row + 1]);
    fr pw = fr_one(), ival = fr_zero();
    for (uint32_t j = 0; j < k; j++) {
        ival = fr_add(ival, fr_mul(fr_load(&w[1 + j]), pw));
        pw = fr_mul(pw, d);
    }
    fr_store(&i2[row], ival);
}

// r' = a' b' - i',  q = (r' - c') z_inv     (proving.rs:492-509)
// c2r holds c' on entry and r' on exit (r' is needed again for the K scalars, c' is not)
__global__ void k_quotient(const fr *__restrict__ a2, const fr *__restrict__ b2, fr *__restrict__ c2r,
                           const fr *__restrict__ i2, const fr *__restrict__ zinv, uint32_t n, fr *__restrict__ q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr cv = fr_load(&c2r[i]);
    const fr r = fr_sub(fr_mul(fr_load(&a2[i]), fr_load(&b2[i])), fr_load(&i2[i]));
    fr_store(&q[i], fr_mul(fr_sub(r, cv), fr_load(&zinv[i])));
    fr_store(&c2r[i], r);
}

// ------------------------------------------------------------------------------------------------
// batched Fr inversion (Montgomery trick, two levels of 16) of out[i] = 1/(leaves[i] - alpha)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t FRB = 16;
__global__ void k_frb_up(const fr *__restrict__ leaves, fr alpha, uint32_t n, fr *__restrict__ pre, fr *__restrict__ tot) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * FRB;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + FRB);
    fr acc = fr_one();
    for (uint32_t i = lo; i < hi; i++) {
        fr_store(&pre[i], acc);
        acc = fr_mul(acc, fr_sub(fr_load(&leaves[i]), alpha));
    }
    fr_store(&tot[g], acc);
}
__global__ void k_frb_down(const fr *__restrict__ leaves, fr alpha, uint32_t n, const fr *__restrict__ pre,
                           const fr *__restrict__ tot_inv, fr *__restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * FRB;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + FRB);
    fr inv = fr_load(&tot_inv[g]);
    for (uint32_t i = hi; i-- > lo;) {
        const fr v = fr_sub(fr_load(&leaves[i]), alpha);
        fr_store(&out[i], fr_mul(inv, fr_load(&pre[i])));
        if (i > lo) inv = fr_mul(inv, v);
    }
}

// partial sums of y_i w_i /(alpha - d_i) for y = a and b: one Fr pair per block  (ec_fft.rs:455-491)
// dinv[2i] = 1/(d_i - alpha), so the term is -y_i w_i dinv[2i]
__global__ void __launch_bounds__(256)
    k_bary_partial(const fr *__restrict__ a, const fr *__restrict__ b, const fr *__restrict__ wts,
                   const fr *__restrict__ dinv, uint32_t n, fr *__restrict__ part) {
    __shared__ fr sa[256], sb[256];
    fr accA = fr_zero(), accB = fr_zero();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const fr wi = fr_mul(fr_load(&wts[i]), fr_load(&dinv[2 * i]));
        accA = fr_add(accA, fr_mul(fr_load(&a[i]), wi));
        accB = fr_add(accB, fr_mul(fr_load(&b[i]), wi));
    }
    sa[threadIdx.x] = accA;
    sb[threadIdx.x] = accB;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sa[threadIdx.x] = fr_add(sa[threadIdx.x], sa[threadIdx.x + o]);
            sb[threadIdx.x] = fr_add(sb[threadIdx.x], sb[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = fr_neg(sa[0]);
        part[2 * blockIdx.x + 1] = fr_neg(sb[0]);
    }
}

// K scalars (proving.rs:599-654): k_a | k_b | k_r (interleaved D, D'), contiguous for the 4n-point MSM
__global__ void k_kscalars(const fr *__restrict__ a, const fr *__restrict__ b, const fr *__restrict__ iv,
                           const fr *__restrict__ r2, const fr *__restrict__ dinv, fr a0, fr b0, fr r0, uint32_t n,
                           fr *__restrict__ ks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr av = fr_load(&a[i]), bv = fr_load(&b[i]);
    const fr d1 = fr_load(&dinv[2 * i]), d2 = fr_load(&dinv[2 * i + 1]);
    fr_store(&ks[i], fr_mul(fr_sub(av, a0), d1));
    fr_store(&ks[n + i], fr_mul(fr_sub(bv, b0), d1));
    const fr rv = fr_sub(fr_mul(av, bv), fr_load(&iv[i]));
    fr_store(&ks[2 * n + 2 * i], fr_mul(fr_sub(rv, r0), d1));
    fr_store(&ks[2 * n + 2 * i + 1], fr_mul(fr_sub(fr_load(&r2[i]), r0), d2));
}

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct dvp_domain {
    dvp_ctx *ctx = nullptr;
    int log_n2 = 0;
    uint32_t n2 = 0, n = 0;
    int levels = 0;
    DevBuf leaves;               // 2n Fr
    std::vector<DevBuf> dec, rec; // per level
    DevBuf z_vals2inv, bar_wts;  // n Fr each
    std::vector<fr> x0, t;       // isogeny chain (host)
    fr last[2];
    DevBuf work;                 // extend workspace
};
struct dvp_r1cs {
    dvp_ctx *ctx = nullptr;
    R1csDev dev;
    DevBuf bufs[10];
    DevBuf tasks, coeffs29, w29; // 29-bit-limb row products: (row, matrix) tasks ordered by length per chunk
    bool fast = false;
    DevBuf tbufs[9]; // transposed matrices (colptr, row, coeff id per matrix), built on the first setup
    bool t_ready = false;
    size_t nwires = 0;
};
// One helper thread per prover handle, alive from dvp_prover_create to dvp_prover_destroy: dvp_prove hands it the
// witness commitment (the g_m MSM) so that its host-side launch sequence overlaps the main thread's.
struct ProverWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = true, quit = false;
    void start() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [this] { return has_job || quit; });
                if (quit) return;
                std::function<void()> f = std::move(job);
                has_job = false;
                lk.unlock();
                f();
                lk.lock();
                done = true;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> f) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(f);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return done; });
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv.notify_all();
        }
        th.join();
    }
};

struct dvp_prover {
    dvp_ctx *ctx = nullptr;
    ProverWorker worker;
    dvp_domain *dom = nullptr;
    dvp_r1cs *r1cs = nullptr;
    int slot_gm = 0, slot_gq = 0, slot_gk = 0;
    DevBuf vec;  // 13 n Fr: a b c i a' b' c' i' q | k_a k_b k_r(2n)   (r' reuses c' after q)
    DevBuf wit, dinv, pre, tot, tot2, pre2, part;
    // commit_p as one MSM (ctx->prove_joint): this rank's g_m | g_q shards side by side (with their own window tables,
    // built on first use) and the scalars w | q copied next to each other
    SrsSlot joint;
    uint64_t joint_ver[2] = {~0ull, ~0ull};
    DevBuf jscal;
    void *h_part = nullptr;
    float ms[8] = {0};
    cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// host-side Fr helpers on the isogeny chain
static void sw_dbl_host(SwPt &p, const fr &A) {
    const fr three = fr_from_u64(3);
    const fr lam = fr_mul(fr_add(fr_mul(three, fr_sqr(p.x)), A), fr_inv(fr_add(p.y, p.y)));
    const fr x3 = fr_sub(fr_sub(fr_sqr(lam), p.x), p.x);
    const fr y3 = fr_sub(fr_mul(lam, fr_sub(p.x, x3)), p.y);
    p.x = x3;
    p.y = y3;
}
static fr fr_from_dec_host(const char *s) {
    fr acc = fr_zero();
    const fr ten = fr_from_u64(10);
    for (; *s; s++) acc = fr_add(fr_mul(acc, ten), fr_from_u64((uint64_t)(*s - '0')));
    return acc;
}
// Z_S(x) for S = D (shift 0) or D' (shift 1):  Z_S(x) = v(x)^(|S|/2) Z_psi(S)(psi(x))
static fr vanish_at_host(const dvp_domain *d, int shift, fr x) {
    fr acc = fr_one();
    for (int k = 0; k < d->levels; k++) {
        const fr v = fr_sub(x, d->x0[k]);
        acc = fr_mul(acc, fr_pow2k(v, d->levels - k - 1));
        x = fr_add(x, fr_mul(d->t[k], fr_inv(v)));
    }
    return fr_mul(acc, fr_sub(x, d->last[shift]));
}

static int run_chain(dvp_domain *d, const std::vector<DevBuf *> &layers, int shift, int deriv, fr *out, fr *tmp) {
    // level `levels`: one root; walk up to level 0
    cudaStream_t st = d->ctx->stream;
    const fr base = deriv ? fr_one() : fr_sub(d->last[1 - shift], d->last[shift]);
    fr *cur = (d->levels & 1) ? tmp : out, *nxt = (d->levels & 1) ? out : tmp;
    CKP(cudaMemcpyAsync(cur, &base, sizeof(fr), cudaMemcpyHostToDevice, st));
    for (int k = d->levels - 1; k >= 0; k--) {
        const uint32_t m = d->n >> k;
        k_chain_level<<<cdivp(m, 128), 128, 0, st>>>(layers[k]->as<fr>(), m, d->levels - k - 1, d->x0[k], d->t[k], shift,
                                                    deriv, cur, nxt);
        fr *sw = cur;
        cur = nxt;
        nxt = sw;
    }
    CKP(cudaGetLastError());
    // after `levels` swaps the result sits in `out`
    k_fr_inv_each<<<cdivp(d->n, 128), 128, 0, st>>>(out, d->n);
    CKP(cudaGetLastError());
    return 0;
}

extern "C" {

// shifted = the tree on the coset C + g instead of C: its leaves are the leaves of the plain tree rotated by one, so
// its extend maps the ODD leaves of the plain tree to the even ones (rotated by one position) -- FFTree::extend(.., Moiety::S0)
static int domain_create_impl(dvp_ctx *ctx, unsigned log2_2n, bool shifted, dvp_domain **out);
int dvp_domain_create(dvp_ctx *ctx, unsigned log2_2n, dvp_domain **out) {
    return domain_create_impl(ctx, log2_2n, false, out);
}
} // extern "C"

static void sw_add_host(SwPt &p, const SwPt &q) { // distinct x
    const fr lam = fr_mul(fr_sub(q.y, p.y), fr_inv(fr_sub(q.x, p.x)));
    const fr x3 = fr_sub(fr_sub(fr_sqr(lam), p.x), q.x);
    const fr y3 = fr_sub(fr_mul(lam, fr_sub(p.x, x3)), p.y);
    p.x = x3;
    p.y = y3;
}

static int domain_create_impl(dvp_ctx *ctx, unsigned log2_2n, bool shifted, dvp_domain **out) {
    if (!ctx || !out || log2_2n < 2 || log2_2n > 28) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    CKP(cudaSetDevice(ctx->device));
    dvp_domain *d = new dvp_domain();
    d->ctx = ctx;
    d->log_n2 = (int)log2_2n;
    d->n2 = 1u << log2_2n;
    d->n = d->n2 >> 1;
    d->levels = d->log_n2 - 1;
    cudaStream_t st = ctx->stream;
    int rc = 0;
    // constants: /root/reference/src/ec_fft.rs:209-229
    SwConsts sc;
    sc.A = fr_from_dec_host("2125753088427212854352924174339172498722499297750753614229533284661082");
    SwPt G;
    G.x = fr_from_dec_host("1969398527398874941115360315313056361667745675958024267654083765592400");
    G.y = fr_from_dec_host("917696706299601920847965073366118878832337776859300472447868491055982");
    sc.C.x = fr_from_dec_host("1557215852494830750811239888869886110709986867282698163663807961412586");
    sc.C.y = fr_from_dec_host("2302954593454110051167704558708330032236229062988890422530712548754008");
    SwPt g = G;
    for (int i = 0; i < 28 - d->log_n2; i++) sw_dbl_host(g, sc.A); // ec_fft.rs:121-124
    if (shifted) sw_add_host(sc.C, g);
    sc.pow2[0] = g;
    for (int b = 1; b < 28; b++) {
        sc.pow2[b] = sc.pow2[b - 1];
        if (b < d->log_n2) sw_dbl_host(sc.pow2[b], sc.A);
    }
    // isogeny chain (x0_k, t_k): kernel = the order-2 point of <g_k>
    d->x0.resize(d->log_n2);
    d->t.resize(d->log_n2);
    {
        fr Ak = sc.A;
        SwPt gk = g;
        for (int k = 0; k < d->log_n2; k++) {
            SwPt K = gk;
            for (uint32_t o = d->n2 >> k; o > 2; o >>= 1) sw_dbl_host(K, Ak);
            if (!fr_is_zero(K.y)) {
                delete d;
                return DVP_ERR_INTERNAL;
            }
            const fr x0 = K.x, t = fr_add(fr_mul(fr_from_u64(3), fr_sqr(x0)), Ak);
            d->x0[k] = x0;
            d->t[k] = t;
            if ((d->n2 >> k) > 2) {
                const fr di = fr_inv(fr_sub(gk.x, x0)), q = fr_mul(t, di);
                SwPt ng;
                ng.x = fr_add(gk.x, q);
                ng.y = fr_sub(gk.y, fr_mul(fr_mul(q, di), gk.y));
                gk = ng;
                Ak = fr_sub(Ak, fr_mul(fr_from_u64(5), t));
            }
        }
    }
    // layers on the device
    std::vector<DevBuf> layer_store(d->log_n2);
    std::vector<DevBuf *> layers(d->log_n2);
    DevBuf dsc;
    auto fail = [&](int code) {
        for (auto &b : layer_store) b.release();
        dsc.release();
        dvp_domain_destroy(d);
        return code;
    };
    if ((rc = d->leaves.reserve((size_t)d->n2 * sizeof(fr))) || (rc = dsc.reserve(sizeof(SwConsts)))) return fail(rc);
    if (cudaMemcpyAsync(dsc.p, &sc, sizeof(sc), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DVP_ERR_CUDA);
    k_dom_leaves<<<cdivp(d->n2, 128), 128, 0, st>>>(dsc.as<SwConsts>(), d->n2, d->log_n2, d->leaves.as<fr>());
    layers[0] = &d->leaves;
    for (int k = 0; k + 1 < d->log_n2; k++) {
        const uint32_t half = d->n2 >> (k + 1);
        if ((rc = layer_store[k + 1].reserve((size_t)half * sizeof(fr)))) return fail(rc);
        layers[k + 1] = &layer_store[k + 1];
        k_dom_next_layer<<<cdivp(half, 128), 128, 0, st>>>(layers[k]->as<fr>(), half, d->x0[k], d->t[k],
                                                          layers[k + 1]->as<fr>());
    }
    if (cudaGetLastError() != cudaSuccess) return fail(DVP_ERR_CUDA);
    // the D / D' chains end on the two leaves of layer `levels`
    if (cudaMemcpyAsync(d->last, layers[d->levels]->p, 2 * sizeof(fr), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return fail(DVP_ERR_CUDA);
    // extend tables (k_dom_down / k_dom_up): the position scales sigma go down the levels, tau comes back up
    d->dec.resize(d->levels);
    d->rec.resize(d->levels);
    DevBuf scale[2];
    auto fail2 = [&](int code) {
        scale[0].release();
        scale[1].release();
        return fail(code);
    };
    if ((rc = scale[0].reserve((size_t)d->n * sizeof(fr))) || (rc = scale[1].reserve((size_t)d->n * sizeof(fr)))) return fail2(rc);
    const fr *sc_cur = nullptr; // sigma of level 0 is 1
    int flip = 0;
    for (int k = 0; k < d->levels; k++) {
        const uint32_t h = d->n >> (k + 1);
        if ((rc = d->dec[k].reserve((size_t)h * 4 * sizeof(fr))) || (rc = d->rec[k].reserve((size_t)h * 4 * sizeof(fr))))
            return fail2(rc);
        int e = 0;
        while ((1u << e) < h) e++;
        fr *nxt = scale[flip].as<fr>();
        k_dom_down<<<cdivp(h, 64), 64, 0, st>>>(layers[k]->as<fr>(), h, e, d->x0[k], sc_cur, nxt, d->dec[k].as<fr>());
        k_mats_to29<<<cdivp(4 * h, 128), 128, 0, st>>>(d->dec[k].as<fr>(), 4 * h);
        sc_cur = nxt;
        flip ^= 1;
    }
    for (int k = d->levels - 1; k >= 0; k--) { // sc_cur: tau of level k + 1 (the single sigma of the last level at first)
        const uint32_t h = d->n >> (k + 1);
        int e = 0;
        while ((1u << e) < h) e++;
        fr *nxt = scale[flip].as<fr>();
        k_dom_up<<<cdivp(h, 64), 64, 0, st>>>(layers[k]->as<fr>(), h, e, d->x0[k], sc_cur, nxt, k == 0 ? 1 : 0,
                                             d->rec[k].as<fr>());
        k_mats_to29<<<cdivp(4 * h, 128), 128, 0, st>>>(d->rec[k].as<fr>(), 4 * h);
        sc_cur = nxt;
        flip ^= 1;
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return fail2(DVP_ERR_CUDA);
    scale[0].release();
    scale[1].release();
    // prover precomputes: bar_wts = 1/Z'_D(d_i), z_vals2inv = 1/Z_D(d'_i)  (proving.rs:225-325)
    if ((rc = d->bar_wts.reserve((size_t)d->n * sizeof(fr))) || (rc = d->z_vals2inv.reserve((size_t)d->n * sizeof(fr))) ||
        (rc = d->work.reserve((size_t)d->n * sizeof(fr))))
        return fail(rc);
    if ((rc = run_chain(d, layers, 0, 1, d->bar_wts.as<fr>(), d->work.as<fr>()))) return fail(rc);
    if ((rc = run_chain(d, layers, 0, 0, d->z_vals2inv.as<fr>(), d->work.as<fr>()))) return fail(rc);
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DVP_ERR_CUDA);
    for (int k = 1; k < d->log_n2; k++) layer_store[k].release();
    dsc.release();
    *out = d;
    return DVP_OK;
}

extern "C" {

void dvp_domain_destroy(dvp_domain *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    d->leaves.release();
    for (auto &b : d->dec) b.release();
    for (auto &b : d->rec) b.release();
    d->z_vals2inv.release();
    d->bar_wts.release();
    d->work.release();
    delete d;
}

int dvp_domain_leaves(dvp_domain *d, uint64_t *out) {
    if (!d || !out) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(d->ctx->device));
    CKP(cudaMemcpy(out, d->leaves.p, (size_t)d->n2 * 32, cudaMemcpyDeviceToHost));
    return DVP_OK;
}
// read_minimal_fftree_from_file (tree_io.rs:419-433) for the prover's tree2n (proving.rs:436)
int dvp_domain_from_fftree(dvp_ctx *ctx, const uint8_t *file, size_t len, dvp_domain **out) {
    if (!ctx || !out) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    size_t n2 = 0;
    int rc = dvp_fftree_file_leaves(file, len, 0, &n2, nullptr);
    if (rc) return rc;
    unsigned lg = 0;
    while (((size_t)1 << lg) < n2) lg++;
    if (((size_t)1 << lg) != n2 || lg < 2 || lg > 28) return DVP_ERR_BAD_ARG;
    std::vector<uint64_t> want(4 * n2), got(4 * n2);
    if ((rc = dvp_fftree_file_leaves(file, len, 0, &n2, want.data()))) return rc;
    dvp_domain *d = nullptr;
    if ((rc = dvp_domain_create(ctx, lg, &d))) return rc;
    if ((rc = dvp_domain_leaves(d, got.data())) == DVP_OK && memcmp(want.data(), got.data(), 32 * n2) != 0)
        rc = DVP_ERR_DOMAIN_MISMATCH;
    if (rc) {
        dvp_domain_destroy(d);
        return rc;
    }
    *out = d;
    return DVP_OK;
}
int dvp_domain_precomputes(dvp_domain *d, uint64_t *z_vals2inv, uint64_t *bar_wts) {
    if (!d) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(d->ctx->device));
    if (z_vals2inv) CKP(cudaMemcpy(z_vals2inv, d->z_vals2inv.p, (size_t)d->n * 32, cudaMemcpyDeviceToHost));
    if (bar_wts) CKP(cudaMemcpy(bar_wts, d->bar_wts.p, (size_t)d->n * 32, cudaMemcpyDeviceToHost));
    return DVP_OK;
}
// Z_D(x) (shift 0) or Z_D'(x) (shift 1) for a Montgomery x; the value of z_poly.evaluate(x) (ec_fft.rs:475)
int dvp_domain_vanish_at(dvp_domain *d, int shift, const uint64_t x_mont[4], uint64_t out_mont[4]) {
    if (!d || !x_mont || !out_mont || (shift != 0 && shift != 1)) return DVP_ERR_BAD_ARG;
    fr x;
    memcpy(x.v, x_mont, 32);
    const fr r = vanish_at_host(d, shift, x);
    memcpy(out_mont, r.v, 32);
    return DVP_OK;
}

// in-place extend of npoly vectors of n Fr (device), D -> D'
// the levels of sub-problems of size <= EXT_FUSE_M in one launch: `count` contiguous groups ("polys") of
// per_poly points each, poly_stride elements apart; per_poly is a multiple of the sub-problem size
static int extend_fused_launch(dvp_domain *d, fr *data, uint32_t count, uint32_t per_poly, size_t poly_stride) {
    static bool attr_set = false;
    if (!attr_set) {
        CKP(cudaFuncSetAttribute(k_extend_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(EXT_FUSE_M * sizeof(fr))));
        attr_set = true;
    }
    const int logm = std::min(d->levels, EXT_FUSE_LOG), K = d->levels - logm;
    FusedMats fm;
    for (int k = 0; k < EXT_FUSE_LOG; k++) {
        fm.dec[k] = k < logm ? d->dec[K + k].as<fr>() : nullptr;
        fm.rec[k] = k < logm ? d->rec[K + k].as<fr>() : nullptr;
    }
    // chunk = whole sub-problems, at most EXT_FUSE_M points, but small enough that the grid still covers the SMs twice
    int logc = logm;
    while (logc < EXT_FUSE_LOG && (2u << logc) <= per_poly && ((size_t)count * per_poly >> (logc + 1)) >= 296) logc++;
    const uint32_t blocks_per_poly = per_poly >> logc;
    k_extend_fused<<<count * blocks_per_poly, 256, ((size_t)1 << logc) * sizeof(fr), d->ctx->stream>>>(
        data, logm, logc, fm, blocks_per_poly, poly_stride, K == 0 ? 1 : 0);
    CKP(cudaGetLastError());
    return 0;
}

// in-place extend of npoly vectors of n Fr (device), D -> D'
static int extend_device(dvp_domain *d, fr *data, int npoly, size_t stride) {
    cudaStream_t st = d->ctx->stream;
    const uint32_t n = d->n;
    const int K = std::max(0, d->levels - EXT_FUSE_LOG); // levels above the fused ones
    int rc;
    // polynomials in groups of at most 3 (the prover's a, b, c) share each matrix read
    for (int p0 = 0; p0 < npoly; p0 += 3) {
        const int np = std::min(3, npoly - p0);
        fr *base = data + (size_t)p0 * stride;
        for (int k = 0; k < K; k++) extend_level_launch(base, n, n >> (k + 1), d->dec[k].as<fr>(), np, stride, true, k, st);
        if ((rc = extend_fused_launch(d, base, (uint32_t)np, n, stride))) return rc;
        for (int k = K - 1; k >= 0; k--) extend_level_launch(base, n, n >> (k + 1), d->rec[k].as<fr>(), np, stride, false, k, st);
    }
    CKP(cudaGetLastError());
    return 0;
}

// The same extend with `world` ranks sharing every polynomial (all ranks hold the same inputs): after log2(world)
// decompose levels a vector is `world` independent sub-problems in contiguous blocks, so every rank runs the levels
// below on ITS block of each polynomial only, the blocks are all-gathered, and the top recombine levels finish the
// vectors.  The levels above the block compute one output per butterfly only (k_extend_down_sel / k_extend_up_sel:
// the half that contains the block going down, the positions that lead to the rank's own result range coming up), about
// one level pass of work for all of them; everything else is 1 / world of the work -- against one rank per polynomial
// (3 busy ranks of 8) before.  own_range_only: the result is valid on [rank n/world, (rank+1) n/world) of every
// polynomial only (all a prover rank reads); otherwise on the whole vectors.
// Falls back to owner-computes + broadcast when the shapes do not allow it.
static int extend_device_sharded(dvp_domain *d, fr *data, int npoly, size_t stride, bool own_range_only) {
    dvp_ctx *ctx = d->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t n = d->n;
    const int W = ctx->world, R = ctx->rank;
    const int K = std::max(0, d->levels - EXT_FUSE_LOG); // levels above the fused ones
    int s = 0;
    while ((1 << s) < W) s++;
    int rc;
    if ((1 << s) != W || s > K || npoly > 3) {
        for (int pl = 0; pl < npoly; pl++)
            if (pl % W == R && (rc = extend_device(d, data + (size_t)pl * stride, 1, stride))) return rc;
        if ((rc = comm_group(ctx, true))) return rc;
        for (int pl = 0; pl < npoly; pl++)
            if ((rc = comm_broadcast(ctx, data + (size_t)pl * stride, (size_t)n * sizeof(fr), pl % W))) {
                comm_group(ctx, false);
                return rc;
            }
        return comm_group(ctx, false);
    }
    const uint32_t blk = n >> s;
    fr *mine = data + (size_t)R * blk;
    // decompose levels above the block: of every sub-problem on the way only the half that contains the block
    {
        uint32_t off = 0;
        for (int k = 0; k < s; k++) {
            const uint32_t h = n >> (k + 1);
            const int which = (R >> (s - 1 - k)) & 1;
            if (k == 0)
                k_extend_down_sel<3, false><<<cdivp(h, 256), 256, 0, st>>>(data + off, h, d->dec[k].as<fr>(), npoly, stride, which);
            else
                k_extend_down_sel<3, true><<<cdivp(h, 256), 256, 0, st>>>(data + off, h, d->dec[k].as<fr>(), npoly, stride, which);
            off += which ? h : 0;
        }
    }
    for (int k = s; k < K; k++) extend_level_launch(mine, blk, n >> (k + 1), d->dec[k].as<fr>(), npoly, stride, true, k, st);
    if ((rc = extend_fused_launch(d, mine, (uint32_t)npoly, blk, stride))) return rc;
    for (int k = K - 1; k >= s; k--) extend_level_launch(mine, blk, n >> (k + 1), d->rec[k].as<fr>(), npoly, stride, false, k, st);
    CKP(cudaGetLastError());
    if ((rc = comm_group(ctx, true))) return rc;
    for (int pl = 0; pl < npoly; pl++) {
        fr *v = data + (size_t)pl * stride;
        if ((rc = comm_all_gather(ctx, v + (size_t)R * blk, v, (size_t)blk * sizeof(fr)))) {
            comm_group(ctx, false);
            return rc;
        }
    }
    if ((rc = comm_group(ctx, false))) return rc;
    if (own_range_only) {
        // recombine levels above the block: only what leads to the values on this rank's own range [R blk, (R+1) blk)
        for (int k = s - 1; k >= 0; k--) {
            const uint32_t h = n >> (k + 1), q0 = (uint32_t)(((size_t)R * blk) % (2 * (size_t)h));
            const uint32_t cnt = (n / (2 * h)) * blk;
            if (k == 0)
                k_extend_up_sel<3, true><<<cdivp(cnt, 256), 256, 0, st>>>(data, n, h, blk, q0, d->rec[k].as<fr>(), npoly, stride);
            else
                k_extend_up_sel<3, false><<<cdivp(cnt, 256), 256, 0, st>>>(data, n, h, blk, q0, d->rec[k].as<fr>(), npoly, stride);
        }
    } else {
        for (int k = s - 1; k >= 0; k--) extend_level_launch(data, n, n >> (k + 1), d->rec[k].as<fr>(), npoly, stride, false, k, st);
    }
    CKP(cudaGetLastError());
    return 0;
}

// FFTree::extend(evals, Moiety::S1) for npoly vectors (host buffers, npoly x n x 4 u64)
int dvp_ecfft_extend(dvp_domain *d, const uint64_t *in, uint64_t *out, int npoly) {
    if (!d || !in || !out || npoly <= 0) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(d->ctx->device));
    DevBuf tmp;
    int rc = tmp.reserve((size_t)npoly * d->n * sizeof(fr));
    if (rc) return rc;
    cudaStream_t st = d->ctx->stream;
    CKP(cudaMemcpyAsync(tmp.p, in, (size_t)npoly * d->n * 32, cudaMemcpyHostToDevice, st));
    rc = extend_device(d, tmp.as<fr>(), npoly, d->n);
    if (!rc) {
        CKP(cudaMemcpyAsync(out, tmp.p, (size_t)npoly * d->n * 32, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
    }
    tmp.release();
    return rc;
}
int dvp_ecfft_extend_device(dvp_domain *d, void *d_data, int npoly) {
    if (!d || !d_data || npoly <= 0) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(d->ctx->device));
    int rc = extend_device(d, (fr *)d_data, npoly, d->n);
    if (rc) return rc;
    CKP(cudaStreamSynchronize(d->ctx->stream));
    return DVP_OK;
}

int dvp_r1cs_load(dvp_ctx *ctx, size_t nrows, size_t k, size_t nwires, const uint32_t *const rowptr[3],
                  const uint32_t *const wire[3], const uint32_t *const coeff[3], const uint64_t *coeffs_mont,
                  size_t ncoeffs, dvp_r1cs **out) {
    if (!ctx || !out || !rowptr || !wire || !coeff || (!coeffs_mont && ncoeffs) || nwires < 1 + k) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    CKP(cudaSetDevice(ctx->device));
    // validate indices on the host: the reference would panic on an out-of-range index
    for (int w = 0; w < 3; w++) {
        if (!rowptr[w] || rowptr[w][0] != 0) return DVP_ERR_BAD_ARG;
        for (size_t r = 0; r < nrows; r++)
            if (rowptr[w][r + 1] < rowptr[w][r]) return DVP_ERR_BAD_ARG;
        const size_t nnz = rowptr[w][nrows];
        for (size_t p = 0; p < nnz; p++)
            if (wire[w][p] >= nwires || coeff[w][p] >= ncoeffs) return DVP_ERR_BAD_ARG;
    }
    dvp_r1cs *r = new dvp_r1cs();
    r->ctx = ctx;
    r->nwires = nwires;
    size_t n = 2;
    while (n < nrows) n <<= 1; // rows.next_power_of_two(), gnark_r1cs.rs:291 (at least one pair)
    r->dev.nrows = (uint32_t)nrows;
    r->dev.n = (uint32_t)n;
    r->dev.k = (uint32_t)k;
    int rc = 0;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> const void * {
        if (rc) return nullptr;
        if ((rc = b.reserve(bytes ? bytes : 4))) return nullptr;
        if (bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = DVP_ERR_CUDA;
        return b.p;
    };
    for (int w = 0; w < 3; w++) {
        const size_t nnz = rowptr[w][nrows];
        r->dev.rowptr[w] = (const uint32_t *)up(r->bufs[3 * w], rowptr[w], (nrows + 1) * 4);
        r->dev.wire[w] = (const uint32_t *)up(r->bufs[3 * w + 1], wire[w], nnz * 4);
        r->dev.coeff[w] = (const uint32_t *)up(r->bufs[3 * w + 2], coeff[w], nnz * 4);
    }
    r->dev.coeffs = (const fr *)up(r->bufs[9], coeffs_mont, ncoeffs * 32);
    if (!rc && nrows >= R1CS_CHUNK && ncoeffs) {
        // (row, matrix) tasks, counting-sorted by term count (longest first) inside every chunk of rows
        std::vector<uint32_t> tasks(3 * nrows), cnt;
        size_t pos = 0;
        for (size_t r0 = 0; r0 < nrows; r0 += R1CS_CHUNK) {
            const size_t r1 = std::min(nrows, r0 + R1CS_CHUNK);
            uint32_t mx = 0;
            for (int w = 0; w < 3; w++)
                for (size_t q = r0; q < r1; q++) mx = std::max(mx, rowptr[w][q + 1] - rowptr[w][q]);
            cnt.assign((size_t)mx + 2, 0u);
            for (int w = 0; w < 3; w++)
                for (size_t q = r0; q < r1; q++) cnt[mx - (rowptr[w][q + 1] - rowptr[w][q]) + 1]++;
            for (size_t i = 1; i < cnt.size(); i++) cnt[i] += cnt[i - 1];
            for (int w = 0; w < 3; w++)
                for (size_t q = r0; q < r1; q++)
                    tasks[pos + cnt[mx - (rowptr[w][q + 1] - rowptr[w][q])]++] = (uint32_t)(q << 2) | (uint32_t)w;
            pos += 3 * (r1 - r0);
        }
        up(r->tasks, tasks.data(), tasks.size() * 4);
        up(r->coeffs29, coeffs_mont, ncoeffs * 32);
        if (!rc) rc = r->w29.reserve(nwires * 32);
        if (!rc) {
            k_mats_to29<<<cdivp(ncoeffs, 128), 128, 0, ctx->stream>>>(r->coeffs29.as<fr>(), (uint32_t)ncoeffs);
            if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = DVP_ERR_CUDA;
        }
        r->fast = rc == 0;
    }
    if (rc) {
        for (auto &b : r->bufs) b.release();
        delete r;
        return rc;
    }
    *out = r;
    return DVP_OK;
}
void dvp_r1cs_destroy(dvp_r1cs *r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    for (auto &b : r->bufs) b.release();
    for (auto &b : r->tbufs) b.release();
    r->tasks.release();
    r->coeffs29.release();
    r->w29.release();
    delete r;
}

// rows [lo, hi) of this rank (all of them on one GPU), then the ranks exchange their ranges
static int r1cs_eval_device(dvp_r1cs *r, dvp_domain *d, const fr *d_w, fr *a, fr *b, fr *c, fr *iv, int64_t *first_bad) {
    dvp_ctx *ctx = r->ctx;
    cudaStream_t st = ctx->stream;
    int rc;
    const int W = ctx->world;
    const size_t n = r->dev.n;
    if (W > 1 && n % (size_t)W) return DVP_ERR_BAD_ARG;
    size_t lo, hi;
    dvp_shard_range(n, ctx->rank, W, &lo, &hi);
    if ((rc = ctx->small.reserve(64 + 8 * 64))) return rc;
    unsigned long long init = ~0ull, bad = 0;
    unsigned long long *d_bad = (unsigned long long *)ctx->small.p; // [0] mine, [8..8+W) gathered
    CKP(cudaMemcpyAsync(d_bad, &init, 8, cudaMemcpyHostToDevice, st));
    if (r->fast && lo % R1CS_CHUNK == 0 && (hi % R1CS_CHUNK == 0 || hi >= r->dev.nrows)) {
        // rows [lo, hi) own the tasks [3 lo, 3 min(hi, nrows)): chunks are whole
        const uint32_t t_lo = 3 * (uint32_t)std::min<size_t>(lo, r->dev.nrows), t_hi = 3 * (uint32_t)std::min<size_t>(hi, r->dev.nrows);
        k_fr_to29<<<cdivp(r->nwires, 128), 128, 0, st>>>(d_w, (uint32_t)r->nwires, r->w29.as<fr>());
        if (t_hi > t_lo)
            k_r1cs_sides<<<cdivp(t_hi - t_lo, 128), 128, 0, st>>>(r->dev, r->tasks.as<uint32_t>(), t_lo, t_hi, r->coeffs29.as<fr>(),
                                                                 r->w29.as<fr>(), a, b, c);
        k_r1cs_finish<<<cdivp(hi - lo, 128), 128, 0, st>>>(r->dev, (uint32_t)lo, (uint32_t)hi, d_w, d->leaves.as<fr>(), a, b, c,
                                                          iv, d_bad);
    } else {
        k_r1cs_eval<<<cdivp(hi - lo, 128), 128, 0, st>>>(r->dev, (uint32_t)lo, (uint32_t)hi, d_w, d->leaves.as<fr>(), a, b,
                                                        c, iv, d_bad);
    }
    CKP(cudaGetLastError());
    if (W > 1) {
        const size_t chunk = (n / W) * sizeof(fr);
        if ((rc = comm_group(ctx, true))) return rc;
        fr *vs[4] = {a, b, c, iv};
        for (fr *v : vs)
            if (!rc) rc = comm_all_gather(ctx, v + lo, v, chunk);
        if (!rc) rc = comm_all_gather(ctx, d_bad, d_bad + 8, 8);
        const int rc_end = comm_group(ctx, false); // the group is closed whatever happened inside it
        if (rc || (rc = rc_end)) return rc;
        unsigned long long all[64];
        CKP(cudaMemcpyAsync(all, d_bad + 8, 8 * W, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
        bad = ~0ull;
        for (int i = 0; i < W; i++) bad = all[i] < bad ? all[i] : bad;
    } else {
        CKP(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
    }
    if (first_bad) *first_bad = bad == ~0ull ? -1 : (int64_t)bad;
    return bad == ~0ull ? DVP_OK : DVP_ERR_UNSATISFIED;
}

// get_matrix_evaluations_from_witness (proving.rs:348-403), host buffers: assignment nwires x 4, outputs n x 4 each
int dvp_r1cs_eval(dvp_r1cs *r, dvp_domain *d, const uint64_t *assignment, uint64_t *a, uint64_t *b, uint64_t *c,
                  uint64_t *iv, int64_t *first_bad_row) {
    if (!r || !d || !assignment || !a || !b || !c || !iv) return DVP_ERR_BAD_ARG;
    if (d->n != r->dev.n) return DVP_ERR_LENGTH_MISMATCH;
    CKP(cudaSetDevice(r->ctx->device));
    DevBuf w, o;
    int rc;
    const size_t n = r->dev.n;
    if ((rc = w.reserve(r->nwires * 32)) || (rc = o.reserve(4 * n * 32))) {
        w.release();
        o.release();
        return rc;
    }
    CKP(cudaMemcpyAsync(w.p, assignment, r->nwires * 32, cudaMemcpyHostToDevice, r->ctx->stream));
    fr *ov = o.as<fr>();
    rc = r1cs_eval_device(r, d, w.as<fr>(), ov, ov + n, ov + 2 * n, ov + 3 * n, first_bad_row);
    if (rc == DVP_OK || rc == DVP_ERR_UNSATISFIED) {
        cudaMemcpy(a, ov, n * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(b, ov + n, n * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(c, ov + 2 * n, n * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(iv, ov + 3 * n, n * 32, cudaMemcpyDeviceToHost);
    }
    w.release();
    o.release();
    return rc;
}

// Fill the fresh wires of a synthetic circuit (see k_r1cs_solve_level); assignment: nwires x 4 u64, in place.
int dvp_r1cs_synth_solve(dvp_r1cs *r, uint64_t *assignment, unsigned nlevels) {
    if (!r || !assignment || nlevels == 0) return DVP_ERR_BAD_ARG;
    if (r->nwires < 1 + (size_t)r->dev.k + r->dev.nrows) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(r->ctx->device));
    cudaStream_t st = r->ctx->stream;
    DevBuf w;
    int rc = w.reserve(r->nwires * 32);
    if (rc) return rc;
    CKP(cudaMemcpyAsync(w.p, assignment, r->nwires * 32, cudaMemcpyHostToDevice, st));
    const uint32_t per = cdivp(r->dev.nrows, nlevels);
    for (unsigned l = 0; l < nlevels; l++)
        k_r1cs_solve_level<<<cdivp(per, 128), 128, 0, st>>>(r->dev, w.as<fr>(), l, nlevels);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(assignment, w.p, r->nwires * 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    w.release();
    return e == cudaSuccess ? DVP_OK : DVP_ERR_CUDA;
}

// Row evaluation alone, outputs left on the device: average milliseconds over `reps` runs (CUDA events), for bench.py
int dvp_r1cs_eval_time(dvp_r1cs *r, dvp_domain *d, const uint64_t *assignment, int reps, float *ms) {
    if (!r || !d || !assignment || !ms || reps < 1) return DVP_ERR_BAD_ARG;
    if (d->n != r->dev.n) return DVP_ERR_LENGTH_MISMATCH;
    CKP(cudaSetDevice(r->ctx->device));
    cudaStream_t st = r->ctx->stream;
    DevBuf w, o;
    int rc;
    const size_t n = r->dev.n;
    if ((rc = w.reserve(r->nwires * 32)) || (rc = o.reserve(4 * n * 32))) {
        w.release();
        o.release();
        return rc;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaMemcpyAsync(w.p, assignment, r->nwires * 32, cudaMemcpyHostToDevice, st);
    fr *ov = o.as<fr>();
    int64_t bad = -1;
    rc = r1cs_eval_device(r, d, w.as<fr>(), ov, ov + n, ov + 2 * n, ov + 3 * n, &bad); // warm-up
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps && !rc; i++) rc = r1cs_eval_device(r, d, w.as<fr>(), ov, ov + n, ov + 2 * n, ov + 3 * n, &bad);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    *ms = t / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    w.release();
    o.release();
    return rc;
}

static int prover_create_local(dvp_ctx *ctx, dvp_domain *dom, dvp_r1cs *r1cs, int slot_gm, int slot_gq, int slot_gk,
                               dvp_prover **out);
int dvp_prover_create(dvp_ctx *ctx, dvp_domain *dom, dvp_r1cs *r1cs, int slot_gm, int slot_gq, int slot_gk,
                      dvp_prover **out) {
    if (!ctx || !out) return DVP_ERR_BAD_ARG;
    int rc = prover_create_local(ctx, dom, r1cs, slot_gm, slot_gq, slot_gk, out);
    if (ctx->world > 1) {
        // every rank learns whether every rank has a prover: a rank that failed here (wrong slot sizes, out of memory)
        // must not leave the others waiting in the first collective of dvp_prove
        const int all = comm_agree(ctx, rc);
        if (all && !rc) {
            dvp_prover_destroy(*out);
            *out = nullptr;
            rc = all;
        }
    }
    return rc;
}
static int prover_create_local(dvp_ctx *ctx, dvp_domain *dom, dvp_r1cs *r1cs, int slot_gm, int slot_gq, int slot_gk,
                               dvp_prover **out) {
    if (!ctx || !dom || !r1cs || !out) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    if (dom->n != r1cs->dev.n) return DVP_ERR_LENGTH_MISMATCH;
    const size_t n = dom->n;
    // multi_scalar_mul asserts scalars.len() == points.len() (curve.rs:142)
    // (with a communicator: this rank's range of each vector)
    {
        const size_t tot[3] = {r1cs->nwires, n, 4 * n};
        const int sl[3] = {slot_gm, slot_gq, slot_gk};
        for (int i = 0; i < 3; i++) {
            size_t lo, hi;
            // g_k is sharded by the index range of D: g_k_0[ilo, ihi) | g_k_1[ilo, ihi) | g_k_2[2 ilo, 2 ihi)
            dvp_shard_range(i == 2 ? n : tot[i], ctx->rank, ctx->world, &lo, &hi);
            const size_t want = i == 2 ? 4 * (hi - lo) : hi - lo;
            if (sl[i] < 0 || sl[i] >= DVP_MAX_SRS_SLOTS || ctx->slots[sl[i]].n != want) return DVP_ERR_LENGTH_MISMATCH;
        }
    }
    CKP(cudaSetDevice(ctx->device));
    dvp_prover *p = new dvp_prover();
    p->ctx = ctx;
    p->dom = dom;
    p->r1cs = r1cs;
    p->slot_gm = slot_gm;
    p->slot_gq = slot_gq;
    p->slot_gk = slot_gk;
    int rc = 0;
    const size_t ng = (2 * n + FRB - 1) / FRB;
    if ((rc = p->vec.reserve(13 * n * sizeof(fr))) || (rc = p->wit.reserve((r1cs->nwires + 64) * sizeof(fr))) ||
        (rc = p->dinv.reserve(2 * n * sizeof(fr))) || (rc = p->pre.reserve(2 * n * sizeof(fr))) ||
        (rc = p->tot.reserve(ng * sizeof(fr))) || (rc = p->tot2.reserve((ng / FRB + 2) * sizeof(fr))) ||
        (rc = p->pre2.reserve(ng * sizeof(fr))) || (rc = p->part.reserve(2 * 1024 * sizeof(fr))) ||
        cudaMallocHost(&p->h_part, 2 * 1024 * sizeof(fr)) != cudaSuccess) {
        dvp_prover_destroy(p);
        return rc ? rc : DVP_ERR_OOM;
    }
    p->worker.start();
    *out = p;
    return DVP_OK;
}
void dvp_prover_destroy(dvp_prover *p) {
    if (!p) return;
    p->worker.stop();
    cudaSetDevice(p->ctx->device);
    DevBuf *all[] = {&p->vec, &p->wit, &p->dinv, &p->pre, &p->tot, &p->tot2, &p->pre2, &p->part, &p->jscal,
                     &p->joint.buf, &p->joint.table};
    for (auto b : all) b->release();
    if (p->h_part) cudaFreeHost(p->h_part);
    for (auto &e : p->ev)
        if (e) cudaEventDestroy(e);
    delete p;
}

} // extern "C"

// out[i] = 1/(leaves[i] - alpha) for the 2n leaves: two Montgomery-trick levels, then one inversion per group
__global__ void k_frb_up2(const fr *__restrict__ in, uint32_t n, fr *__restrict__ pre, fr *__restrict__ tot) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * FRB;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + FRB);
    fr acc = fr_one();
    for (uint32_t i = lo; i < hi; i++) {
        fr_store(&pre[i], acc);
        acc = fr_mul(acc, fr_load(&in[i]));
    }
    fr_store(&tot[g], fr_inv(acc)); // top level: invert the group product directly
}
__global__ void k_frb_down2(fr *__restrict__ inout, uint32_t n, const fr *__restrict__ pre, const fr *__restrict__ tot_inv) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = g * FRB;
    if (lo >= n) return;
    const uint32_t hi = min(n, lo + FRB);
    fr inv = fr_load(&tot_inv[g]);
    for (uint32_t i = hi; i-- > lo;) {
        const fr v = fr_load(&inout[i]);
        fr_store(&inout[i], fr_mul(inv, fr_load(&pre[i])));
        if (i > lo) inv = fr_mul(inv, v);
    }
}

// dinv[j] = 1/(leaves[leaf_lo + j] - alpha) for j < n2 (a rank's range of the leaves; all of them on one GPU)
static int denominators(dvp_prover *p, const fr &alpha, size_t leaf_lo, uint32_t n2) {
    cudaStream_t st = p->ctx->stream;
    const uint32_t ng = cdivp(n2, FRB), ng2 = cdivp(ng, FRB);
    const fr *leaves = p->dom->leaves.as<fr>() + leaf_lo;
    k_frb_up<<<cdivp(ng, 128), 128, 0, st>>>(leaves, alpha, n2, p->pre.as<fr>(), p->tot.as<fr>());
    k_frb_up2<<<cdivp(ng2, 64), 64, 0, st>>>(p->tot.as<fr>(), ng, p->pre2.as<fr>(), p->tot2.as<fr>());
    k_frb_down2<<<cdivp(ng2, 64), 64, 0, st>>>(p->tot.as<fr>(), ng, p->pre2.as<fr>(), p->tot2.as<fr>());
    k_frb_down<<<cdivp(ng, 128), 128, 0, st>>>(leaves, alpha, n2, p->pre.as<fr>(), p->tot.as<fr>(), p->dinv.as<fr>());
    CKP(cudaGetLastError());
    return 0;
}

static void fr_to_le29_host(uint8_t out[29], const fr &a) {
    uint32_t c[8];
    fr_to_canonical(c, a);
    for (int i = 0; i < 29; i++) out[i] = (uint8_t)(c[i >> 2] >> (8 * (i & 3)));
}

// Proof::prove (proving.rs:426-688) with everything resident on the device.
// stages (optional, host, 13 n x 4 u64): a b c i a' b' c' i' q k_a k_b k_r(2n)
static int prove_impl(dvp_prover *p, const uint64_t *pub, size_t k, const uint64_t *priv, size_t npriv,
                      uint8_t proof[118], uint64_t *stages) {
    dvp_ctx *ctx = p->ctx;
    dvp_domain *d = p->dom;
    dvp_r1cs *r = p->r1cs;
    cudaStream_t st = ctx->stream;
    const size_t n = d->n;
    if (k != r->dev.k) return DVP_ERR_BAD_ARG;                    // assert_eq!(inst.num_public_inputs, ..) proving.rs:361
    if (1 + k + npriv != r->nwires) return DVP_ERR_LENGTH_MISMATCH; // msm(assignment, g_m) length check, curve.rs:142
    CKP(cudaSetDevice(ctx->device));
    // stage-timing events live in the prover handle (created once), so error returns leak nothing
    cudaEvent_t *ev = p->ev, &ev_h2d = p->ev[8];
    if (!p->ev[0])
        for (auto &e : p->ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
    fr *V = p->vec.as<fr>();
    fr *a = V, *b = V + n, *c = V + 2 * n, *iv = V + 3 * n, *a2 = V + 4 * n, *b2 = V + 5 * n, *c2 = V + 6 * n,
       *i2 = V + 7 * n, *q = V + 8 * n, *ks = V + 9 * n;
    // assignment = [1, public.., private..]  (proving.rs:449-452)
    fr *w = p->wit.as<fr>();
    const fr one = fr_one();
    CKP(cudaMemcpyAsync(w, &one, 32, cudaMemcpyHostToDevice, st));
    if (k) CKP(cudaMemcpyAsync(w + 1, pub, k * 32, cudaMemcpyHostToDevice, st));
    if (ctx->world == 1) {
        if (npriv) CKP(cudaMemcpyAsync(w + 1 + k, priv, npriv * 32, cudaMemcpyHostToDevice, st));
    } else {
        // every rank uploads 1/world of the private witness over its own PCIe link, the rest arrives over NVLink
        const size_t Wn = (size_t)ctx->world, chunk = (r->nwires + Wn - 1) / Wn; // wires per rank, last one short
        const size_t lo = std::min(r->nwires, chunk * (size_t)ctx->rank), hi = std::min(r->nwires, lo + chunk);
        const size_t plo = std::max(lo, 1 + k), phi = std::max(hi, 1 + k); // the private part of [lo, hi)
        if (phi > plo)
            CKP(cudaMemcpyAsync(w + plo, priv + (plo - 1 - k) * 4, (phi - plo) * 32, cudaMemcpyHostToDevice, st));
        int rcg = comm_all_gather(ctx, w + chunk * (size_t)ctx->rank, w, chunk * sizeof(fr));
        if (rcg) return rcg;
    }
    cudaEventRecord(ev_h2d, st);
    const int W = ctx->world, R = ctx->rank;
    size_t wlo, whi, qlo, qhi, klo, khi;
    dvp_shard_range(r->nwires, R, W, &wlo, &whi);
    dvp_shard_range(n, R, W, &qlo, &qhi);
    dvp_shard_range(4 * n, R, W, &klo, &khi);
    // The commitment to the witness (proving.rs:462-463) depends on the assignment only: it runs on its own streams
    // from a helper thread while this thread evaluates the rows, extends and forms the quotient, so the latency-bound
    // phases of the MSM are filled with the Fr-side work.
    AffPt msm_gm, msm_q, kzg, part, part_gm;
    int rc_gm = 0;
    const size_t nm = whi - wlo, nq = qhi - qlo;
    // msm(w, g_m) + msm(q, g_q) (proving.rs:463, 512, 515) is ONE MSM over g_m | g_q -- one sort, one reduction, one
    // read-back and wider windows instead of two of each.  The other form (knob prove_joint = 0, or no memory for the
    // joint vector's window tables) runs the g_m MSM from the helper thread beside the Fr-side work.
    // (with several ranks the choice must be the same everywhere -- the forms differ in their collectives -- so it does
    // not depend on a rank's own memory there)
    const bool joint = ctx->prove_joint == 1 || (ctx->prove_joint < 0 && (ctx->world > 1 || !p->joint.table_failed));
    int rc_joint = 0; // a rank-local failure here travels with the partial sum: nobody is left waiting in a collective
    if (joint) {
        SrsSlot &gm = ctx->slots[p->slot_gm], &gq = ctx->slots[p->slot_gq];
        if (gm.n != nm || gq.n != nq) {
            rc_joint = DVP_ERR_LENGTH_MISMATCH; // a slot was reloaded with another length after dvp_prover_create
        } else if (p->joint.n != nm + nq || p->joint_ver[0] != gm.version || p->joint_ver[1] != gq.version) {
            p->joint.invalidate();
            p->joint.n = 0;
            if ((rc_joint = p->joint.buf.reserve((nm + nq) * sizeof(AffPt))) == 0 &&
                (rc_joint = p->jscal.reserve((nm + nq) * sizeof(fr))) == 0) {
                if ((nm && cudaMemcpyAsync(p->joint.buf.p, gm.buf.p, nm * sizeof(AffPt), cudaMemcpyDeviceToDevice, st) != cudaSuccess) ||
                    (nq && cudaMemcpyAsync(p->joint.buf.as<AffPt>() + nm, gq.buf.p, nq * sizeof(AffPt), cudaMemcpyDeviceToDevice,
                                           st) != cudaSuccess))
                    rc_joint = DVP_ERR_CUDA;
            }
            if (!rc_joint) {
                p->joint.n = nm + nq;
                p->joint_ver[0] = gm.version;
                p->joint_ver[1] = gq.version;
            }
        }
    }
    CKP(cudaEventRecord(ctx->ev_aux, st));
    part_gm = pt_inf();
    if (!joint)
        p->worker.submit([&, ctx, p, w, wlo, whi] {
            if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux, 0) != cudaSuccess) {
                rc_gm = DVP_ERR_CUDA;
                return;
            }
            rc_gm = slot_msm(ctx, p->slot_gm, 0, (const uint32_t *)(w + wlo), whi - wlo, &part_gm, ctx->aux_stream);
        });
    struct Joiner { // every return path waits for the helper: it writes into this frame
        ProverWorker &wk;
        ~Joiner() { wk.wait(); }
    } joiner{p->worker};
    int64_t bad = -1;
    int rc = r1cs_eval_device(r, d, w, a, b, c, iv, &bad);
    if (rc) return rc;
    cudaEventRecord(ev[1], st);
    // extend a, b, c to D' (i' in closed form), proving.rs:475-482
    CKP(cudaMemcpyAsync(a2, a, 3 * n * sizeof(fr), cudaMemcpyDeviceToDevice, st));
    if (W == 1) {
        if ((rc = extend_device(d, a2, 3, n))) return rc;
    } else {
        // every rank works on its block of all three polynomials (extend_device_sharded)
        if ((rc = extend_device_sharded(d, a2, 3, n, /*own_range_only=*/stages == nullptr))) return rc;
    }
    // i', r' (overwrites c') and q, proving.rs:492-509: a rank needs them on its own index range only (its g_q shard and
    // its K scalars), which is also where the sharded extend left valid values; the stage dump wants whole vectors
    const size_t elo = (W > 1 && !stages) ? qlo : 0, ecnt = (W > 1 && !stages) ? qhi - qlo : n;
    k_ivals_ext<<<cdivp(ecnt, 128), 128, 0, st>>>(w, (uint32_t)k, d->leaves.as<fr>() + 2 * elo, (uint32_t)ecnt, i2 + elo);
    if (stages) {
        CKP(cudaMemcpyAsync(stages, V, 8 * n * 32, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
    }
    k_quotient<<<cdivp(ecnt, 128), 128, 0, st>>>(a2 + elo, b2 + elo, c2 + elo, i2 + elo, d->z_vals2inv.as<fr>() + elo,
                                                (uint32_t)ecnt, q + elo);
    CKP(cudaGetLastError());
    cudaEventRecord(ev[3], st);
    AffPt commit;
    if (joint) {
        fr *js = p->jscal.as<fr>();
        if (!rc_joint && ((nm && cudaMemcpyAsync(js, w + wlo, nm * sizeof(fr), cudaMemcpyDeviceToDevice, st) != cudaSuccess) ||
                          (nq && cudaMemcpyAsync(js + nm, q + qlo, nq * sizeof(fr), cudaMemcpyDeviceToDevice, st) != cudaSuccess)))
            rc_joint = DVP_ERR_CUDA;
        cudaEventRecord(ev[2], st);
        part = pt_inf();
        rc = rc_joint ? rc_joint : slot_msm_at(ctx, p->joint, 0, (const uint32_t *)js, nm + nq, &part);
        // (a rank whose MSM failed still takes part in the exchange; every rank then returns that failure)
        if ((rc = comm_fold_points(ctx, part, rc, &commit))) return rc;
        cudaEventRecord(ev[4], st);
    } else {
        p->worker.wait();
        if ((rc = comm_fold_points(ctx, part_gm, rc_gm, &msm_gm))) return rc;
        cudaEventRecord(ev[2], st);
        part = pt_inf();
        rc = slot_msm(ctx, p->slot_gq, 0, (const uint32_t *)(q + qlo), qhi - qlo, &part);
        if ((rc = comm_fold_points(ctx, part, rc, &msm_q))) return rc;
        cudaEventRecord(ev[4], st);
        commit = host::aff_add(msm_q, msm_gm); // proving.rs:515
    }
    host::encode30(proof, commit);
    // Fiat-Shamir challenge, proving.rs:517-558
    std::vector<uint8_t> pub29(29 * k + 1);
    std::vector<fr> pubv(k);
    for (size_t j = 0; j < k; j++) {
        memcpy(pubv[j].v, pub + 4 * j, 32);
        fr_to_le29_host(&pub29[29 * j], pubv[j]);
    }
    uint8_t al[32];
    if (!host::transcript_alpha(proof, pub29.data(), k, al)) return DVP_ERR_BAD_ARG;
    uint32_t alc[8];
    memcpy(alc, al, 32);
    const fr alpha = fr_from_canonical(alc);
    // From here every rank works on its own index range [ilo, ihi) of D and D' (all of it on one GPU): the
    // denominators 1/(leaf - alpha), the partial barycentric sums and its part of the K scalars.  The g_k shard of a
    // rank is g_k_0[ilo, ihi) | g_k_1[ilo, ihi) | g_k_2[2 ilo, 2 ihi), which is exactly what its K scalars multiply.
    size_t ilo, ihi;
    dvp_shard_range(n, R, W, &ilo, &ihi);
    const uint32_t cnt = (uint32_t)(ihi - ilo);
    if ((rc = denominators(p, alpha, 2 * ilo, 2 * cnt))) return rc;
    // a0, b0 by barycentric evaluation, i0 by Horner (ec_fft.rs:455-491, srs.rs:412)
    const int nblk = 592;
    k_bary_partial<<<nblk, 256, 0, st>>>(a + ilo, b + ilo, d->bar_wts.as<fr>() + ilo, p->dinv.as<fr>(), cnt, p->part.as<fr>());
    CKP(cudaGetLastError());
    CKP(cudaMemcpyAsync(p->h_part, p->part.p, 2 * nblk * sizeof(fr), cudaMemcpyDeviceToHost, st));
    // Z_D(alpha), Z_D'(alpha) and i0 on the host (O(log^2 n) and O(k)) while the device forms the denominators and sums
    const fr za = vanish_at_host(d, 0, alpha);
    const bool alpha_in_domain = fr_is_zero(za) || fr_is_zero(vanish_at_host(d, 1, alpha));
    fr i0 = fr_zero(), pw = fr_one();
    for (size_t j = 0; j < k; j++) {
        i0 = fr_add(i0, fr_mul(pubv[j], pw));
        pw = fr_mul(pw, alpha);
    }
    CKP(cudaStreamSynchronize(st));
    if (alpha_in_domain && W == 1) return DVP_ERR_ALPHA_IN_DOMAIN;
    fr sa = fr_zero(), sb = fr_zero();
    const fr *hp = (const fr *)p->h_part;
    for (int i = 0; i < nblk; i++) {
        sa = fr_add(sa, hp[2 * i]);
        sb = fr_add(sb, hp[2 * i + 1]);
    }
    if (W > 1) {
        // sum of the ranks' partial sums: all-gather of 64 bytes per rank
        if ((rc = ctx->commbuf.reserve(16384))) return rc; // (sized when the communicator was made)
        fr *cb = reinterpret_cast<fr *>(ctx->commbuf.as<char>() + 8192);
        const fr mine[2] = {sa, sb};
        CKP(cudaMemcpyAsync(cb + 2 * W, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
        if ((rc = comm_all_gather(ctx, cb + 2 * W, cb, sizeof(mine)))) return rc;
        fr all[128];
        CKP(cudaMemcpyAsync(all, cb, (size_t)W * sizeof(mine), cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
        sa = fr_zero();
        sb = fr_zero();
        for (int i = 0; i < W; i++) {
            sa = fr_add(sa, all[2 * i]);
            sb = fr_add(sb, all[2 * i + 1]);
        }
    }
    if (alpha_in_domain) return DVP_ERR_ALPHA_IN_DOMAIN; // (the same alpha on every rank: all of them leave here)
    const fr a0 = fr_mul(sa, za), b0 = fr_mul(sb, za);
    const fr r0 = fr_sub(fr_mul(a0, b0), i0);
    // ks: k_a | k_b | k_r of this rank's range, contiguous (4 cnt scalars; with one rank the reference's own order)
    k_kscalars<<<cdivp(cnt, 128), 128, 0, st>>>(a + ilo, b + ilo, iv + ilo, c2 + ilo, p->dinv.as<fr>(), a0, b0, r0, cnt, ks);
    CKP(cudaGetLastError());
    cudaEventRecord(ev[5], st);
    if (stages) {
        CKP(cudaMemcpyAsync(stages + 8 * n * 4, q, 5 * n * 32, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
    }
    klo = 0;
    khi = 4 * (size_t)cnt;
    part = pt_inf();
    rc = slot_msm(ctx, p->slot_gk, 0, (const uint32_t *)(ks + klo), khi - klo, &part);
    if ((rc = comm_fold_points(ctx, part, rc, &kzg))) return rc;
    cudaEventRecord(ev[6], st);
    cudaEventSynchronize(ev[6]);
    host::encode30(proof + 30, kzg);
    fr_to_le29_host(proof + 60, a0);
    fr_to_le29_host(proof + 89, b0);
    // stream order of the events: 0 (h2d) r1cs 1 extend+quotient 3 [wait for the g_m MSM] 2 msm g_q 4 .. 5 msm g_k 6
    cudaEventElapsedTime(&p->ms[0], ev[0], ev[1]);
    cudaEventElapsedTime(&p->ms[2], ev[1], ev[3]);
    cudaEventElapsedTime(&p->ms[1], ev[3], ev[2]); // the part of the g_m MSM that was not hidden behind the Fr-side work
    cudaEventElapsedTime(&p->ms[3], ev[2], ev[4]);
    cudaEventElapsedTime(&p->ms[4], ev[4], ev[5]);
    cudaEventElapsedTime(&p->ms[5], ev[5], ev[6]);
    cudaEventElapsedTime(&p->ms[6], ev[0], ev_h2d); // witness upload, part of ms[0]
    p->ms[0] -= p->ms[6];
    return DVP_OK;
}

extern "C" {
int dvp_prove(dvp_prover *p, const uint64_t *public_mont, size_t k, const uint64_t *private_mont, size_t npriv,
              uint8_t proof118[118]) {
    if (!p || !proof118 || (!public_mont && k) || (!private_mont && npriv)) return DVP_ERR_BAD_ARG;
    return prove_impl(p, public_mont, k, private_mont, npriv, proof118, nullptr);
}
int dvp_prove_stages(dvp_prover *p, const uint64_t *public_mont, size_t k, const uint64_t *private_mont, size_t npriv,
                     uint8_t proof118[118], uint64_t *stages) {
    if (!p || !proof118 || !stages || (!public_mont && k) || (!private_mont && npriv)) return DVP_ERR_BAD_ARG;
    return prove_impl(p, public_mont, k, private_mont, npriv, proof118, stages);
}
// stage times of the last prove in ms: r1cs, msm g_m, extend+quotient, msm g_q, challenge+K scalars, msm g_k, witness upload
int dvp_prove_last_times(dvp_prover *p, float ms[7]) {
    if (!p || !ms) return DVP_ERR_BAD_ARG;
    for (int i = 0; i < 7; i++) ms[i] = p->ms[i];
    return DVP_OK;
}
}

// ------------------------------------------------------------------------------------------------
// Setup on the device: the discrete logs of the SRS from a trapdoor
// (SRS::verifier_runs_setup / compute_srs_matrices / accumulate_m_values, /root/reference/src/srs.rs:53-167,177-361),
// then the points by the batched fixed-base multiplication (MsmEngine::mulgen).  The reference reaches
// L_i(tau), Z_D(tau), the barycentric weights and Z on the other half-domain through vanish/exit/enter
// (O(n log^2 n), "2 hrs+"); here they come from the chain rule in O(n log n) like the prover precomputes.
// ------------------------------------------------------------------------------------------------
// per leaf j = 2i + sh:  L_i^S(tau) = Z_S(tau) / ((tau - s_i) Z'_S(s_i)) for S = D (sh = 0) or D' (sh = 1), and the
// unified-domain basis value ltl[j] = L_i^S(tau) Z_other(tau) / Z_other(s_i)   (ec_fft.rs:340-390,424-450)
__global__ void k_setup_lagrange(const fr *__restrict__ dinv /* 1/(leaf - tau) */, const fr *__restrict__ bw0,
                                 const fr *__restrict__ bw1, const fr *__restrict__ zo_inv0, const fr *__restrict__ zo_inv1,
                                 fr zt0, fr zt1, uint32_t n2, fr *__restrict__ lt0, fr *__restrict__ lt1, fr *__restrict__ ltl) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const uint32_t i = j >> 1, sh = j & 1;
    const fr ninv = fr_neg(fr_load(&dinv[j])); // 1/(tau - s)
    const fr l = fr_mul(fr_mul(sh ? zt1 : zt0, fr_load(sh ? &bw1[i] : &bw0[i])), ninv);
    fr_store(sh ? &lt1[i] : &lt0[i], l);
    fr_store(&ltl[j], fr_mul(fr_mul(l, sh ? zt0 : zt1), fr_load(sh ? &zo_inv0[i] : &zo_inv1[i])));
}

// accumulate_m_values (srs.rs:53-84) as a transposed product: one warp per wire over the three column lists
struct R1csT {
    const uint32_t *colptr[3];
    const uint32_t *row[3];
    const uint32_t *cid[3];
};
__global__ void __launch_bounds__(128)
    k_setup_mvals(R1csT t, const fr *__restrict__ coeffs, const fr *__restrict__ lt0, fr delta, fr delta2, fr eps,
                  uint32_t nwires, fr *__restrict__ sc_m) {
    const uint32_t wire = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wire >= nwires) return;
    fr tot = fr_zero();
    for (int which = 0; which < 3; which++) {
        fr acc = fr_zero();
        for (uint32_t p = t.colptr[which][wire] + lane; p < t.colptr[which][wire + 1]; p += 32)
            acc = fr_add(acc, fr_mul(fr_load(&coeffs[t.cid[which][p]]), fr_load(&lt0[t.row[which][p]])));
        if (which == 1) acc = fr_mul(acc, delta);
        if (which == 2) acc = fr_mul(acc, delta2);
        tot = fr_add(tot, acc);
    }
    for (int o = 16; o > 0; o >>= 1) {
        fr other;
#pragma unroll
        for (int k = 0; k < 8; k++) other.v[k] = __shfl_down_sync(0xffffffffu, tot.v[k], o);
        tot = fr_add(tot, other);
    }
    if (lane == 0) fr_store(&sc_m[wire], fr_mul(tot, eps));
}
// Vandermonde block (gnark_r1cs.rs:357-383): partial sums of d_i^j L_i(tau), j < k (k <= 8), one set per block
__global__ void __launch_bounds__(256)
    k_setup_vand(const fr *__restrict__ leaves, const fr *__restrict__ lt0, uint32_t n, uint32_t k, fr *__restrict__ part) {
    __shared__ fr sh[256];
    fr acc[8];
    for (uint32_t j = 0; j < 8; j++) acc[j] = fr_zero();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const fr d = fr_load(&leaves[2 * i]);
        fr pw = fr_load(&lt0[i]);
        for (uint32_t j = 0; j < k; j++) {
            acc[j] = fr_add(acc[j], pw);
            pw = fr_mul(pw, d);
        }
    }
    for (uint32_t j = 0; j < k; j++) {
        sh[threadIdx.x] = acc[j];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] = fr_add(sh[threadIdx.x], sh[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) part[blockIdx.x * 8 + j] = sh[0];
        __syncthreads();
    }
}
// sc_q[i] = f L'_i(tau);  sc_k = L(tau) | delta L(tau) | delta^2 ltl   (srs.rs:112-167)
__global__ void k_setup_qk(const fr *__restrict__ lt0, const fr *__restrict__ lt1, const fr *__restrict__ ltl, fr f, fr delta,
                           fr delta2, uint32_t n, fr *__restrict__ sc_q, fr *__restrict__ sc_k) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr l0 = fr_load(&lt0[i]);
    fr_store(&sc_q[i], fr_mul(f, fr_load(&lt1[i])));
    fr_store(&sc_k[i], l0);
    fr_store(&sc_k[n + i], fr_mul(l0, delta));
    fr_store(&sc_k[2 * n + 2 * i], fr_mul(fr_load(&ltl[2 * i]), delta2));
    fr_store(&sc_k[2 * n + 2 * i + 1], fr_mul(fr_load(&ltl[2 * i + 1]), delta2));
}

// the isogeny images of the leaves (layer k has n2 >> k points); layers[0] = the domain's own leaves
static int build_layers(dvp_domain *d, std::vector<DevBuf> &store, std::vector<DevBuf *> &layers) {
    cudaStream_t st = d->ctx->stream;
    store.resize(d->log_n2);
    layers.resize(d->log_n2);
    layers[0] = &d->leaves;
    int rc;
    for (int k = 0; k + 1 < d->log_n2; k++) {
        const uint32_t half = d->n2 >> (k + 1);
        if ((rc = store[k + 1].reserve((size_t)half * sizeof(fr)))) return rc;
        layers[k + 1] = &store[k + 1];
        k_dom_next_layer<<<cdivp(half, 128), 128, 0, st>>>(layers[k]->as<fr>(), half, d->x0[k], d->t[k], layers[k + 1]->as<fr>());
    }
    CKP(cudaGetLastError());
    return 0;
}

static int setup_scalars_device(dvp_r1cs *r, dvp_domain *d, const fr &tau, const fr &delta, const fr &eps, fr *sc_m,
                                fr *sc_q, fr *sc_k) {
    dvp_ctx *ctx = r->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t n = d->n, n2 = d->n2;
    const size_t nw = r->nwires;
    int rc = 0;
    const fr zt0 = vanish_at_host(d, 0, tau), zt1 = vanish_at_host(d, 1, tau);
    if (fr_is_zero(zt0) || fr_is_zero(zt1)) return DVP_ERR_ALPHA_IN_DOMAIN; // tau must not be a domain point
    const fr delta2 = fr_mul(delta, delta);
    std::vector<DevBuf> tmp(12);
    auto cleanup = [&](int code) {
        for (auto &b : tmp) b.release();
        return code;
    };
    DevBuf &bw1 = tmp[0], &zo1 = tmp[1], &work = tmp[2], &dinv = tmp[3], &pre = tmp[4], &tot = tmp[5], &pre2 = tmp[6],
           &tot2 = tmp[7], &lt0 = tmp[8], &lt1 = tmp[9], &ltl = tmp[10], &part = tmp[11];
    const uint32_t ng = cdivp(n2, FRB), ng2 = cdivp(ng, FRB);
    if ((rc = bw1.reserve((size_t)n * 32)) || (rc = zo1.reserve((size_t)n * 32)) || (rc = work.reserve((size_t)n * 32)) ||
        (rc = dinv.reserve((size_t)n2 * 32)) || (rc = pre.reserve((size_t)n2 * 32)) || (rc = tot.reserve((size_t)ng * 32)) ||
        (rc = pre2.reserve((size_t)ng * 32)) || (rc = tot2.reserve(((size_t)ng2 + 2) * 32)) || (rc = lt0.reserve((size_t)n * 32)) ||
        (rc = lt1.reserve((size_t)n * 32)) || (rc = ltl.reserve((size_t)n2 * 32)) || (rc = part.reserve(592 * 8 * 32)))
        return cleanup(rc);
    {
        // 1/Z'_{D'}(d'_i) and 1/Z_{D'}(d_i): the chain rule on the odd leaves (the even ones are domain precomputes)
        std::vector<DevBuf> store;
        std::vector<DevBuf *> layers;
        if ((rc = build_layers(d, store, layers)) || (rc = run_chain(d, layers, 1, 1, bw1.as<fr>(), work.as<fr>())) ||
            (rc = run_chain(d, layers, 1, 0, zo1.as<fr>(), work.as<fr>()))) {
            for (auto &b : store) b.release();
            return cleanup(rc);
        }
        cudaStreamSynchronize(st);
        for (auto &b : store) b.release();
    }
    // 1/(leaf - tau) for all 2n leaves
    const fr *leaves = d->leaves.as<fr>();
    k_frb_up<<<cdivp(ng, 128), 128, 0, st>>>(leaves, tau, n2, pre.as<fr>(), tot.as<fr>());
    k_frb_up2<<<cdivp(ng2, 64), 64, 0, st>>>(tot.as<fr>(), ng, pre2.as<fr>(), tot2.as<fr>());
    k_frb_down2<<<cdivp(ng2, 64), 64, 0, st>>>(tot.as<fr>(), ng, pre2.as<fr>(), tot2.as<fr>());
    k_frb_down<<<cdivp(ng, 128), 128, 0, st>>>(leaves, tau, n2, pre.as<fr>(), tot.as<fr>(), dinv.as<fr>());
    k_setup_lagrange<<<cdivp(n2, 128), 128, 0, st>>>(dinv.as<fr>(), d->bar_wts.as<fr>(), bw1.as<fr>(), d->z_vals2inv.as<fr>(),
                                                     zo1.as<fr>(), zt0, zt1, n2, lt0.as<fr>(), lt1.as<fr>(), ltl.as<fr>());
    if (cudaGetLastError() != cudaSuccess) return cleanup(DVP_ERR_CUDA);
    // transposed matrices (built once per circuit)
    if (!r->t_ready) {
        std::vector<uint32_t> rowptr, wire, cid, colptr(nw + 1), trow, tcid;
        for (int w = 0; w < 3; w++) {
            const size_t nrows = r->dev.nrows;
            rowptr.resize(nrows + 1);
            if (cudaMemcpy(rowptr.data(), r->dev.rowptr[w], (nrows + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
                return cleanup(DVP_ERR_CUDA);
            const size_t nnz = rowptr[nrows];
            wire.resize(nnz ? nnz : 1);
            cid.resize(nnz ? nnz : 1);
            if (nnz && (cudaMemcpy(wire.data(), r->dev.wire[w], nnz * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                        cudaMemcpy(cid.data(), r->dev.coeff[w], nnz * 4, cudaMemcpyDeviceToHost) != cudaSuccess))
                return cleanup(DVP_ERR_CUDA);
            std::fill(colptr.begin(), colptr.end(), 0u);
            for (size_t p = 0; p < nnz; p++) colptr[wire[p] + 1]++;
            for (size_t c = 0; c < nw; c++) colptr[c + 1] += colptr[c];
            trow.assign(nnz ? nnz : 1, 0u);
            tcid.assign(nnz ? nnz : 1, 0u);
            std::vector<uint32_t> cur(colptr.begin(), colptr.end() - 1);
            for (size_t row = 0; row < nrows; row++)
                for (uint32_t p = rowptr[row]; p < rowptr[row + 1]; p++) {
                    const uint32_t pos = cur[wire[p]]++;
                    trow[pos] = (uint32_t)row;
                    tcid[pos] = cid[p];
                }
            DevBuf *b = &r->tbufs[3 * w];
            if ((rc = b[0].reserve((nw + 1) * 4)) || (rc = b[1].reserve(trow.size() * 4)) || (rc = b[2].reserve(tcid.size() * 4)))
                return cleanup(rc);
            cudaMemcpy(b[0].p, colptr.data(), (nw + 1) * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(b[1].p, trow.data(), trow.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(b[2].p, tcid.data(), tcid.size() * 4, cudaMemcpyHostToDevice);
        }
        r->t_ready = true;
    }
    R1csT t;
    for (int w = 0; w < 3; w++) {
        t.colptr[w] = r->tbufs[3 * w].as<uint32_t>();
        t.row[w] = r->tbufs[3 * w + 1].as<uint32_t>();
        t.cid[w] = r->tbufs[3 * w + 2].as<uint32_t>();
    }
    k_setup_mvals<<<cdivp(nw * 32, 128), 128, 0, st>>>(t, r->dev.coeffs, lt0.as<fr>(), delta, delta2, eps, (uint32_t)nw, sc_m);
    // public wires: minus delta^2 sum_i d_i^j L_i(tau)  (the D block appended to the O rows)
    const uint32_t k = r->dev.k;
    if (k > 8) return cleanup(DVP_ERR_BAD_ARG);
    if (k) {
        const int nblk = 592;
        k_setup_vand<<<nblk, 256, 0, st>>>(leaves, lt0.as<fr>(), n, k, part.as<fr>());
        std::vector<fr> hp((size_t)nblk * 8), cur(k);
        if (cudaMemcpyAsync(hp.data(), part.p, hp.size() * 32, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(cur.data(), sc_m + 1, k * 32, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return cleanup(DVP_ERR_CUDA);
        for (uint32_t j = 0; j < k; j++) {
            fr s = fr_zero();
            for (int b = 0; b < nblk; b++) s = fr_add(s, hp[(size_t)b * 8 + j]);
            cur[j] = fr_sub(cur[j], fr_mul(fr_mul(s, delta2), eps));
        }
        if (cudaMemcpyAsync(sc_m + 1, cur.data(), k * 32, cudaMemcpyHostToDevice, st) != cudaSuccess) return cleanup(DVP_ERR_CUDA);
        cudaStreamSynchronize(st);
    }
    const fr f = fr_mul(fr_mul(zt0, delta2), eps);
    k_setup_qk<<<cdivp(n, 128), 128, 0, st>>>(lt0.as<fr>(), lt1.as<fr>(), ltl.as<fr>(), f, delta, delta2, n, sc_q, sc_k);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return cleanup(DVP_ERR_CUDA);
    return cleanup(DVP_OK);
}

extern "C" {

// The discrete logs of g_m / g_q / g_k for a trapdoor (tau, delta, epsilon), host buffers (nwires, n, 4n elements).
int dvp_setup_scalars(dvp_r1cs *r, dvp_domain *d, const uint64_t trapdoor_mont[12], uint64_t *sc_m, uint64_t *sc_q,
                      uint64_t *sc_k) {
    if (!r || !d || !trapdoor_mont || !sc_m || !sc_q || !sc_k) return DVP_ERR_BAD_ARG;
    if (d->n != r->dev.n) return DVP_ERR_LENGTH_MISMATCH;
    CKP(cudaSetDevice(r->ctx->device));
    fr td[3];
    memcpy(td, trapdoor_mont, 96);
    const size_t n = d->n, nw = r->nwires;
    DevBuf out;
    int rc = out.reserve((nw + 5 * n) * 32);
    if (rc) return rc;
    fr *o = out.as<fr>();
    rc = setup_scalars_device(r, d, td[0], td[1], td[2], o, o + nw, o + nw + n);
    if (!rc) {
        cudaMemcpy(sc_m, o, nw * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(sc_q, o + nw, n * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(sc_k, o + nw + n, 4 * n * 32, cudaMemcpyDeviceToHost);
    }
    out.release();
    return rc;
}

// SRS::verifier_runs_setup with the artifacts left resident: slot_gm / slot_gq / slot_gk receive this rank's range
// of g_m (nwires points), g_q (n) and g_k_0|g_k_1|g_k_2 (4n), ready for dvp_prover_create.
int dvp_setup(dvp_r1cs *r, dvp_domain *d, const uint64_t trapdoor_mont[12], int slot_gm, int slot_gq, int slot_gk) {
    if (!r || !d || !trapdoor_mont) return DVP_ERR_BAD_ARG;
    if (d->n != r->dev.n) return DVP_ERR_LENGTH_MISMATCH;
    dvp_ctx *ctx = r->ctx;
    const int slots[3] = {slot_gm, slot_gq, slot_gk};
    for (int s : slots)
        if (s < 0 || s >= DVP_MAX_SRS_SLOTS) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(ctx->device));
    fr td[3];
    memcpy(td, trapdoor_mont, 96);
    const size_t n = d->n, nw = r->nwires;
    DevBuf out;
    int rc = out.reserve((nw + 5 * n) * 32);
    if (rc) return rc;
    fr *o = out.as<fr>();
    rc = setup_scalars_device(r, d, td[0], td[1], td[2], o, o + nw, o + nw + n);
    // g_m and g_q: contiguous ranges; g_k: g_k_0[ilo, ihi) | g_k_1[ilo, ihi) | g_k_2[2 ilo, 2 ihi) for the rank's range of D
    size_t ilo, ihi;
    dvp_shard_range(n, ctx->rank, ctx->world, &ilo, &ihi);
    const size_t cnt = ihi - ilo;
    for (int i = 0; i < 3 && !rc; i++) {
        size_t lo, hi;
        dvp_shard_range(i == 0 ? nw : n, ctx->rank, ctx->world, &lo, &hi);
        const size_t m = i == 2 ? 4 * cnt : hi - lo;
        SrsSlot &s = ctx->slots[slots[i]];
        s.invalidate();
        if ((rc = s.buf.reserve((m ? m : 1) * sizeof(AffPt)))) break;
        s.n = m;
        if (!m) continue;
        if (i < 2) {
            rc = ctx->msm.mulgen((const uint32_t *)((i == 0 ? o : o + nw) + lo), m, s.buf.as<AffPt>());
        } else {
            const fr *sk = o + nw + n;
            AffPt *dst = s.buf.as<AffPt>();
            rc = ctx->msm.mulgen((const uint32_t *)(sk + ilo), cnt, dst);
            if (!rc) rc = ctx->msm.mulgen((const uint32_t *)(sk + n + ilo), cnt, dst + cnt);
            if (!rc) rc = ctx->msm.mulgen((const uint32_t *)(sk + 2 * n + 2 * ilo), 2 * cnt, dst + 2 * cnt);
        }
    }
    out.release();
    return rc;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// FFTree::enter (crate ecfft; reference call sites /root/reference/src/ec_fft.rs:317,411): coefficients of a
// polynomial of degree < n -> its values on the n leaves of the tree (natural order).  Restated from the ECFFT
// construction: P = U + x^(m) V with deg U, V < m; their values on the even leaves of the 2m-leaf tree are the
// values on the m-leaf tree (same points), the odd leaves come from EXTEND, and the two halves are combined with
// s^m.  Bottom-up over block sizes m = 1, 2, .., n/2; every level extends all n/m blocks in one pass because the
// blocks tile the array exactly like the sub-problems of a larger extend.  O(n log^2 n).
// ------------------------------------------------------------------------------------------------
struct ExitLevel { // constants of the tree with m = 2h leaves (S0 = even leaves, S1 = odd leaves)
    dvp_domain *rev = nullptr; // the tree on the shifted coset: extend S1 -> S0 (rotated by one)
    DevBuf xh0inv, xh1;        // s^-h on S0, s^h on S1
    DevBuf zz0, zz1;           // (Z0^2 mod x^h) on S0 and S1, Z0 = vanishing polynomial of S0
};
struct dvp_ecfft_plan {
    dvp_ctx *ctx = nullptr;
    int log_n = 0;
    std::vector<dvp_domain *> dom; // dom[l] = tree with 2^l leaves, l = 2 .. log_n
    std::vector<ExitLevel> lv;     // lv[l] for l = 2 .. log_n
    DevBuf a, b, c;                // n elements each
    DevBuf h[6];                   // n/2 elements each
};

// out block B (size 2m) from blocks 2B, 2B+1 (size m) of cur (values on the even leaves) and ext (odd leaves)
__global__ void k_enter_combine(const fr *__restrict__ cur, const fr *__restrict__ ext, const fr *__restrict__ leaves,
                                uint32_t leaf_stride, uint32_t n, uint32_t m, int log_m, fr *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; // index over n/2 (pair, i)
    if (t >= (n >> 1)) return;
    const uint32_t B = t / m, i = t % m;
    const fr s0 = fr_load(&leaves[(size_t)(2 * i) * leaf_stride]), s1 = fr_load(&leaves[(size_t)(2 * i + 1) * leaf_stride]);
    const fr p0 = fr_pow2k(s0, log_m), p1 = fr_pow2k(s1, log_m);
    const size_t u = (size_t)(2 * B) * m + i, v = u + m;
    fr_store(&out[(size_t)B * 2 * m + 2 * i], fr_add(fr_load(&cur[u]), fr_mul(p0, fr_load(&cur[v]))));
    fr_store(&out[(size_t)B * 2 * m + 2 * i + 1], fr_add(fr_load(&ext[u]), fr_mul(p1, fr_load(&ext[v]))));
}

// all blocks of size h of `v` (len elements in total) through the extend of tree `d` (whose half-domain has h points)
static int extend_blocks(dvp_domain *d, fr *v, uint32_t len) {
    cudaStream_t st = d->ctx->stream;
    const uint32_t h = d->n;
    const int K = std::max(0, d->levels - EXT_FUSE_LOG);
    for (int k = 0; k < K; k++) extend_level_launch(v, len, h >> (k + 1), d->dec[k].as<fr>(), 1, 0, true, k, st);
    int rc = extend_fused_launch(d, v, 1, len, 0);
    if (rc) return rc;
    for (int k = K - 1; k >= 0; k--) extend_level_launch(v, len, h >> (k + 1), d->rec[k].as<fr>(), 1, 0, false, k, st);
    CKP(cudaGetLastError());
    return 0;
}

// ENTER on the device: `io` holds 2^lg coefficients (low degree first) and receives the values on the 2^lg-leaf tree
static int enter_device(dvp_ecfft_plan *p, fr *io, int lg) {
    cudaStream_t st = p->ctx->stream;
    const uint32_t n = 1u << lg;
    fr *cur = io, *ext = p->b.as<fr>(), *nxt = p->c.as<fr>();
    for (int lm = 0; lm < lg; lm++) {
        const uint32_t m = 1u << lm;
        // the tree with 2m leaves; for m = 1 its two leaves are the even leaves of the 4-leaf tree
        dvp_domain *d = p->dom[std::max(2, lm + 1)];
        const uint32_t leaf_stride = lm + 1 < 2 ? 2 : 1;
        CKP(cudaMemcpyAsync(ext, cur, (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
        int rc;
        if (m >= 2 && (rc = extend_blocks(d, ext, n))) return rc;
        k_enter_combine<<<cdivp(n / 2, 128), 128, 0, st>>>(cur, ext, d->leaves.as<fr>(), leaf_stride, n, m, lm, nxt);
        CKP(cudaGetLastError());
        fr *t = cur;
        cur = nxt;
        nxt = t;
    }
    if (cur != io) CKP(cudaMemcpyAsync(io, cur, (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FFTree::exit (crate ecfft; reference call sites /root/reference/src/ec_fft.rs:266,897): values on the n leaves ->
// coefficients.  Restated from the ECFFT construction.  With h = n/2, S0 / S1 the even / odd leaves and Z0 the
// vanishing polynomial of S0 (monic, degree h, Z0(0) != 0), P = U + x^h V is split by a Montgomery reduction in
// which "division by Z0" is the cheap operation (pointwise on S1 after a subtraction that vanishes on S0):
//   REDC(Q) = (Q + x^h k) / Z0,  k = -Q x^-h on S0 extended to S1      (deg Q < 2h  ->  Q Z0^-1 mod x^h, deg < h)
//   U = P mod x^h = REDC(REDC(P) (Z0^2 mod x^h)),   V = (P - U) x^-h on S0
// then U and V are interpolated on the h-leaf tree.  Top-down over block sizes n, n/2, .., 4 and a closed form for 2
// leaves; each level runs two S0 -> S1 and two S1 -> S0 extends over all blocks at once.  O(n log^2 n).
// The per-level constants (Z0^2 mod x^h on S0 and S1) are built bottom-up with the smaller transforms:
// Z0 - x^h = EXIT_h(-s^h), its square mod x^h from two half-size products.
// ------------------------------------------------------------------------------------------------
__global__ void k_deinterleave(const fr *__restrict__ x, uint32_t half, fr *__restrict__ e, fr *__restrict__ o) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    fr_store(&e[t], fr_load(&x[2 * t]));
    fr_store(&o[t], fr_load(&x[2 * t + 1]));
}
// k = -q0 x^-h on S0
__global__ void k_redc_k(const fr *__restrict__ q0, const fr *__restrict__ xh0inv, uint32_t h, uint32_t len, fr *__restrict__ k) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    fr_store(&k[t], fr_neg(fr_mul(fr_load(&q0[t]), fr_load(&xh0inv[t % h]))));
}
// r1 = (q1 + x^h k1) / Z0 on S1, optionally times zz1
__global__ void k_redc_fin(const fr *__restrict__ q1, const fr *__restrict__ k1, const fr *__restrict__ xh1,
                           const fr *__restrict__ z01inv, const fr *__restrict__ mul1, uint32_t h, uint32_t len,
                           fr *__restrict__ r1, fr *__restrict__ r1m) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const uint32_t i = t % h;
    const fr r = fr_mul(fr_add(fr_load(&q1[t]), fr_mul(fr_load(&xh1[i]), fr_load(&k1[t]))), fr_load(&z01inv[i]));
    fr_store(&r1[t], r);
    if (mul1) fr_store(&r1m[t], fr_mul(r, fr_load(&mul1[i])));
}
// values on S0 out of the reverse extend (rotated by one inside every block), optionally times zz0
__global__ void k_unrotate(const fr *__restrict__ rot, const fr *__restrict__ mul0, uint32_t h, uint32_t len, fr *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const uint32_t b = t / h, i = t % h;
    fr v = fr_load(&rot[(size_t)b * h + (i + h - 1) % h]);
    if (mul0) v = fr_mul(v, fr_load(&mul0[i]));
    fr_store(&out[t], v);
}
// next level: block b of size 2h -> block 2b = U on S0, block 2b+1 = V = (P - U) x^-h on S0
__global__ void k_exit_split(const fr *__restrict__ p0, const fr *__restrict__ u0, const fr *__restrict__ xh0inv, uint32_t h,
                             uint32_t len, fr *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const uint32_t b = t / h, i = t % h;
    const fr u = fr_load(&u0[t]);
    fr_store(&out[(size_t)(2 * b) * h + i], u);
    fr_store(&out[(size_t)(2 * b + 1) * h + i], fr_mul(fr_sub(fr_load(&p0[t]), u), fr_load(&xh0inv[i])));
}
// two leaves: c1 = (p1 - p0)/(s1 - s0), c0 = p0 - c1 s0
__global__ void k_exit_base(fr *__restrict__ x, uint32_t pairs, fr s0, fr dinv) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pairs) return;
    const fr p0 = fr_load(&x[2 * t]), p1 = fr_load(&x[2 * t + 1]);
    const fr c1 = fr_mul(fr_sub(p1, p0), dinv);
    fr_store(&x[2 * t], fr_sub(p0, fr_mul(c1, s0)));
    fr_store(&x[2 * t + 1], c1);
}
// per-level constants: s^h on S1, s^-h on S0 (h = 2^log_h)
__global__ void k_exit_pows(const fr *__restrict__ leaves, uint32_t h, int log_h, fr *__restrict__ xh0inv, fr *__restrict__ xh1) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    fr_store(&xh0inv[i], fr_inv(fr_pow2k(fr_load(&leaves[2 * i]), log_h)));
    fr_store(&xh1[i], fr_pow2k(fr_load(&leaves[2 * i + 1]), log_h));
}
__global__ void k_neg_inv(const fr *__restrict__ in, uint32_t n, fr *__restrict__ out) { // out = -1/in
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store(&out[i], fr_neg(fr_inv(fr_load(&in[i]))));
}
__global__ void k_mul_pointwise(const fr *__restrict__ a, const fr *__restrict__ b, uint32_t n, fr *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store(&out[i], fr_mul(fr_load(&a[i]), fr_load(&b[i])));
}
// zzc[j] = cll[j] + 2 clh[j - h/2] (j >= h/2), padded with zeros to 2h
__global__ void k_zz_coeffs(const fr *__restrict__ cll, const fr *__restrict__ clh, uint32_t h, fr *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * h) return;
    fr v = fr_zero();
    if (j < h) {
        v = fr_load(&cll[j]);
        if (j >= h / 2) {
            const fr t = fr_load(&clh[j - h / 2]);
            v = fr_add(v, fr_add(t, t));
        }
    }
    fr_store(&out[j], v);
}

// EXIT on the device: `io` holds the values on the 2^lg-leaf tree and receives the 2^lg coefficients
static int exit_device(dvp_ecfft_plan *p, fr *io, int lg) {
    cudaStream_t st = p->ctx->stream;
    const uint32_t n = 1u << lg, half = n >> 1;
    fr *x = io, *y = p->c.as<fr>();
    fr *P0 = p->h[0].as<fr>(), *P1 = p->h[1].as<fr>(), *K = p->h[2].as<fr>(), *R1 = p->h[3].as<fr>(), *T1 = p->h[4].as<fr>(),
       *T0 = p->h[5].as<fr>();
    int rc;
    for (int l = lg; l >= 2; l--) {
        const uint32_t h = 1u << (l - 1);
        dvp_domain *fwd = p->dom[l];
        ExitLevel &L = p->lv[l];
        const fr *xh0inv = L.xh0inv.as<fr>(), *xh1 = L.xh1.as<fr>(), *z01inv = fwd->z_vals2inv.as<fr>();
        const uint32_t g = cdivp(half, 128);
        k_deinterleave<<<g, 128, 0, st>>>(x, half, P0, P1);
        // REDC #1: R = P / Z0 mod x^h on S1 (R1), and R (Z0^2 mod x^h) on S1 (T1)
        k_redc_k<<<g, 128, 0, st>>>(P0, xh0inv, h, half, K);
        if ((rc = extend_blocks(fwd, K, half))) return rc;
        k_redc_fin<<<g, 128, 0, st>>>(P1, K, xh1, z01inv, L.zz1.as<fr>(), h, half, R1, T1);
        // R on S0 (reverse extend, rotated), times Z0^2 mod x^h
        if ((rc = extend_blocks(L.rev, R1, half))) return rc;
        k_unrotate<<<g, 128, 0, st>>>(R1, L.zz0.as<fr>(), h, half, T0);
        // REDC #2: U = P mod x^h on S1, then on S0
        k_redc_k<<<g, 128, 0, st>>>(T0, xh0inv, h, half, K);
        if ((rc = extend_blocks(fwd, K, half))) return rc;
        k_redc_fin<<<g, 128, 0, st>>>(T1, K, xh1, z01inv, nullptr, h, half, R1, nullptr);
        if ((rc = extend_blocks(L.rev, R1, half))) return rc;
        k_unrotate<<<g, 128, 0, st>>>(R1, nullptr, h, half, T0);
        k_exit_split<<<g, 128, 0, st>>>(P0, T0, xh0inv, h, half, y);
        CKP(cudaGetLastError());
        fr *t = x;
        x = y;
        y = t;
    }
    {
        // two-leaf trees: the even leaves of the 4-leaf tree
        fr s[3];
        CKP(cudaMemcpyAsync(s, p->dom[2]->leaves.p, 3 * sizeof(fr), cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
        const fr dinv = fr_inv(fr_sub(s[2], s[0]));
        k_exit_base<<<cdivp(half, 128), 128, 0, st>>>(x, half, s[0], dinv);
        CKP(cudaGetLastError());
    }
    if (x != io) CKP(cudaMemcpyAsync(io, x, (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// constants of level l (tree with m = 2^l leaves, h = m/2), given that the levels below are ready
static int exit_level_build(dvp_ecfft_plan *p, int l) {
    dvp_ctx *ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t m = 1u << l, h = m >> 1;
    ExitLevel &L = p->lv[l];
    int rc;
    if ((rc = domain_create_impl(ctx, (unsigned)l, true, &L.rev))) return rc;
    if ((rc = L.xh0inv.reserve((size_t)h * 32)) || (rc = L.xh1.reserve((size_t)h * 32)) || (rc = L.zz0.reserve((size_t)h * 32)) ||
        (rc = L.zz1.reserve((size_t)h * 32)))
        return rc;
    k_exit_pows<<<cdivp(h, 64), 64, 0, st>>>(p->dom[l]->leaves.as<fr>(), h, l - 1, L.xh0inv.as<fr>(), L.xh1.as<fr>());
    CKP(cudaGetLastError());
    // R0 = Z0 - x^h has degree < h and equals -s^h on S0 (the leaves of the h-leaf tree): r0 = EXIT_h(-s^h)
    DevBuf t[4];
    auto done = [&](int code) {
        for (auto &b : t) b.release();
        return code;
    };
    for (auto &b : t)
        if ((rc = b.reserve((size_t)m * 32))) return done(rc);
    fr *r0 = t[0].as<fr>(), *eL = t[1].as<fr>(), *eH = t[2].as<fr>(), *w = t[3].as<fr>();
    k_neg_inv<<<cdivp(h, 64), 64, 0, st>>>(L.xh0inv.as<fr>(), h, r0); // -(s^h) = -1/(s^-h)
    if ((rc = exit_device(p, r0, l - 1))) return done(rc);
    // (Z0^2 mod x^h) = (R0^2 mod x^h) = L^2 + 2 x^(h/2) (L H mod x^(h/2)),  R0 = L + x^(h/2) H
    const uint32_t q = h >> 1;
    CKP(cudaMemsetAsync(eL, 0, (size_t)h * 32, st));
    CKP(cudaMemsetAsync(eH, 0, (size_t)h * 32, st));
    CKP(cudaMemcpyAsync(eL, r0, (size_t)q * 32, cudaMemcpyDeviceToDevice, st));
    CKP(cudaMemcpyAsync(eH, r0 + q, (size_t)q * 32, cudaMemcpyDeviceToDevice, st));
    if ((rc = enter_device(p, eL, l - 1)) || (rc = enter_device(p, eH, l - 1))) return done(rc);
    k_mul_pointwise<<<cdivp(h, 128), 128, 0, st>>>(eL, eH, h, eH); // L H on the h-leaf tree (degree < h - 1)
    k_mul_pointwise<<<cdivp(h, 128), 128, 0, st>>>(eL, eL, h, eL); // L^2
    if ((rc = exit_device(p, eL, l - 1)) || (rc = exit_device(p, eH, l - 1))) return done(rc);
    k_zz_coeffs<<<cdivp(m, 128), 128, 0, st>>>(eL, eH, h, w);
    if ((rc = enter_device(p, w, l))) return done(rc); // values on the m-leaf tree
    k_deinterleave<<<cdivp(h, 128), 128, 0, st>>>(w, h, L.zz0.as<fr>(), L.zz1.as<fr>());
    CKP(cudaGetLastError());
    CKP(cudaStreamSynchronize(st));
    return done(0);
}

extern "C" {

int dvp_ecfft_plan_create(dvp_ctx *ctx, unsigned log2_n, dvp_ecfft_plan **out) {
    if (!ctx || !out || log2_n < 1 || log2_n > 27) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    dvp_ecfft_plan *p = new dvp_ecfft_plan();
    p->ctx = ctx;
    p->log_n = (int)log2_n;
    const unsigned top = std::max(2u, log2_n);
    p->dom.assign(top + 1, nullptr);
    p->lv.resize(top + 1);
    int rc = 0;
    const size_t n = (size_t)1 << top;
    if ((rc = p->a.reserve(n * 32)) || (rc = p->b.reserve(n * 32)) || (rc = p->c.reserve(n * 32))) {
        dvp_ecfft_plan_destroy(p);
        return rc;
    }
    for (auto &b : p->h)
        if ((rc = b.reserve(n * 16))) {
            dvp_ecfft_plan_destroy(p);
            return rc;
        }
    for (unsigned l = 2; l <= top && !rc; l++) {
        rc = dvp_domain_create(ctx, l, &p->dom[l]);
        if (!rc) rc = exit_level_build(p, (int)l);
    }
    if (rc) {
        dvp_ecfft_plan_destroy(p);
        return rc;
    }
    *out = p;
    return DVP_OK;
}

void dvp_ecfft_plan_destroy(dvp_ecfft_plan *p) {
    if (!p) return;
    for (auto d : p->dom) dvp_domain_destroy(d);
    for (auto &L : p->lv) {
        dvp_domain_destroy(L.rev);
        L.xh0inv.release();
        L.xh1.release();
        L.zz0.release();
        L.zz1.release();
    }
    p->a.release();
    p->b.release();
    p->c.release();
    for (auto &b : p->h) b.release();
    delete p;
}

// coeffs (n x 4 u64 Montgomery, low degree first) -> evals on the n leaves x(C + i G_n), host buffers
int dvp_ecfft_enter(dvp_ecfft_plan *p, const uint64_t *coeffs, uint64_t *evals) {
    if (!p || !coeffs || !evals) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    const size_t n = (size_t)1 << p->log_n;
    fr *io = p->a.as<fr>();
    CKP(cudaMemcpyAsync(io, coeffs, n * 32, cudaMemcpyHostToDevice, st));
    int rc = enter_device(p, io, p->log_n);
    if (rc) return rc;
    CKP(cudaMemcpyAsync(evals, io, n * 32, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    return DVP_OK;
}

// evals on the n leaves -> coeffs (inverse of dvp_ecfft_enter), host buffers
int dvp_ecfft_exit(dvp_ecfft_plan *p, const uint64_t *evals, uint64_t *coeffs) {
    if (!p || !coeffs || !evals) return DVP_ERR_BAD_ARG;
    CKP(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    const size_t n = (size_t)1 << p->log_n;
    fr *io = p->a.as<fr>();
    CKP(cudaMemcpyAsync(io, evals, n * 32, cudaMemcpyHostToDevice, st));
    int rc = p->log_n >= 1 ? exit_device(p, io, p->log_n) : 0;
    if (rc) return rc;
    CKP(cudaMemcpyAsync(coeffs, io, n * 32, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    return DVP_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// SRS::verify (/root/reference/src/srs.rs:374-428): the designated verifier's check with the trapdoor.
//   alpha from the transcript of (public inputs, commit_p);  i0 = sum_j x_j alpha^j;  r0 = a0 b0 - i0
//   u0 = (a0 + delta b0 + delta^2 r0) epsilon,  v0 = (tau - alpha) epsilon
//   accept  <=>  multi_scalar_mul([v0, u0], [kzg_k, generator]) == commit_p  and all four fields decode
// The two-term MSM (src/srs.rs:422) runs on the device like every other multi_scalar_mul.
// ------------------------------------------------------------------------------------------------
extern "C" int dvp_verify(dvp_ctx *ctx, const uint64_t trapdoor_mont[12], const uint64_t *public_mont, size_t k,
                          const uint8_t proof118[118], int *accepted) {
    if (!ctx || !trapdoor_mont || (!public_mont && k) || !proof118 || !accepted) return DVP_ERR_BAD_ARG;
    *accepted = 0;
    fr td[3];
    memcpy(td, trapdoor_mont, 96);
    const fr &tau = td[0], &delta = td[1], &eps = td[2];
    // a0, b0: 29-byte canonical little-endian (FrBits::to_fr: must be below p)
    fr ab[2];
    bool fields_ok = true;
    for (int i = 0; i < 2; i++) {
        uint8_t buf[32] = {0};
        memcpy(buf, proof118 + 60 + 29 * i, 29);
        uint32_t c[8];
        memcpy(c, buf, 32);
        if (fr_geq_p(c)) fields_ok = false;
        else ab[i] = fr_from_canonical(c);
    }
    if (!fields_ok) return DVP_OK; // not accepted
    std::vector<uint8_t> pub29(29 * k + 1);
    std::vector<fr> pubv(k);
    for (size_t j = 0; j < k; j++) {
        memcpy(pubv[j].v, public_mont + 4 * j, 32);
        fr_to_le29_host(&pub29[29 * j], pubv[j]);
    }
    uint8_t al[32];
    if (!host::transcript_alpha(proof118, pub29.data(), k, al)) return DVP_ERR_BAD_ARG;
    uint32_t alc[8];
    memcpy(alc, al, 32);
    const fr alpha = fr_from_canonical(alc);
    fr i0 = fr_zero(), pw = fr_one();
    for (size_t j = 0; j < k; j++) {
        i0 = fr_add(i0, fr_mul(pubv[j], pw));
        pw = fr_mul(pw, alpha);
    }
    const fr r0 = fr_sub(fr_mul(ab[0], ab[1]), i0);
    const fr u0 = fr_mul(fr_add(fr_add(ab[0], fr_mul(delta, ab[1])), fr_mul(fr_mul(delta, delta), r0)), eps);
    const fr v0 = fr_mul(fr_sub(tau, alpha), eps);
    // [kzg_k, generator] x [v0, u0] on the device; an encoding that does not decode rejects the proof
    uint8_t pts[60], gen30[30], out30[30];
    memcpy(pts, proof118 + 30, 30);
    host::encode30(gen30, host::k233_generator());
    memcpy(pts + 30, gen30, 30);
    fr sc[2] = {v0, u0};
    // commit_p must decode as well: decode it together with kzg_k by a one-term product with scalar one
    uint8_t cp_out[30];
    const fr one = fr_one();
    int rc = dvp_msm_adhoc(ctx, proof118, (const uint64_t *)one.v, 1, cp_out);
    if (rc == DVP_ERR_INVALID_POINT) return DVP_OK;
    if (rc) return rc;
    rc = dvp_msm_adhoc(ctx, pts, (const uint64_t *)sc, 2, out30);
    if (rc == DVP_ERR_INVALID_POINT) return DVP_OK;
    if (rc) return rc;
    *accepted = memcmp(out30, proof118, 30) == 0 && memcmp(cp_out, proof118, 30) == 0;
    return DVP_OK;
}
