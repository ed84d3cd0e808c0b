/*
 * libdvpari -- C ABI of the B200 (sm_100a) accelerator for DV-Pari's prover hot path.
 *
 * Every entry point replaces one Rust call site of the reference (alpenlabs/dv-pari); the
 * file:line it replaces is cited beside it.  Conventions:
 *   - Fr values cross the boundary exactly as the reference holds them in memory:
 *     ark-ff `Fp256<MontBackend<FqConfig,4>>` = 4 x u64 little-endian limbs of v*2^256 mod p,
 *     fully reduced (src/curve.rs:16-22).  A `Vec<Fr>` can be passed zero-copy.
 *   - curve points cross as the 30-byte xsk233 encoding (`CompressedCurvePoint`, src/curve.rs:67).
 *   - the caller owns every host buffer; the library owns device memory behind opaque handles.
 *   - calls are synchronous, one thread per context; the functions never unwind: they return
 *     DVP_OK or a DVP_ERR_* code where the reference would panic.
 *   - there is no CPU fallback: without a CUDA device dvp_ctx_create fails with DVP_ERR_NO_DEVICE.
 */
#ifndef DVPARI_H
#define DVPARI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DVP_OK = 0,
    DVP_ERR_BAD_ARG = 1,
    DVP_ERR_CUDA = 2,
    DVP_ERR_OOM = 3,
    DVP_ERR_INVALID_POINT = 4,   /* a 30-byte encoding did not decode (reference: assert!, io_utils.rs:223) */
    DVP_ERR_LENGTH_MISMATCH = 5, /* reference: assert_eq!(scalars.len(), points.len()), curve.rs:142 */
    DVP_ERR_UNSATISFIED = 6,     /* reference: assert_eq!(a*b, c+i), proving.rs:389-395 */
    DVP_ERR_ALPHA_IN_DOMAIN = 7, /* reference: assert!(!dom.contains(alpha)), proving.rs:548-556 */
    DVP_ERR_INTERNAL = 8,
    DVP_ERR_NO_DEVICE = 9,
    DVP_ERR_NCCL = 10,
    DVP_ERR_DOMAIN_MISMATCH = 11 /* an FFTree file whose leaves are not the tree of the constants at ec_fft.rs:205-229 */
};

#define DVP_MAX_SRS_SLOTS 16

typedef struct dvp_ctx dvp_ctx;

typedef struct dvp_msm_stats {
    int window_bits, windows, rounds_main, rounds_a, rounds_b;
    unsigned long long launches; /* kernels launched by the last MSM on this context */
    float ms_recode_sort, ms_accumulate, ms_reduce, ms_tail; /* filled when timing is enabled */
    float ms_pass2_round0;            /* the dominant kernel: one launch, CUDA events on the launching stream */
    unsigned long long adds_round0;   /* affine additions that launch finished */
    float ms_device;                  /* scalars in HBM -> partial sums on the host, CUDA events on the context stream */
    int lanes;                        /* concurrent window groups used */
    int tables;                       /* 1 if the MSM ran on the slot's precomputed window multiples */
    float ms_tail_host;               /* host fold of the per-bit partial sums, wall clock */
} dvp_msm_stats;

const char *dvp_strerror(int code);
int dvp_abi_version(void);

/* One context = one CUDA device + one private stream + grow-only scratch. */
int dvp_ctx_create(int device, dvp_ctx **out);
void dvp_ctx_destroy(dvp_ctx *ctx);
/* knobs: "msm_window_bits" (0 = automatic), "msm_lanes" (0 = automatic: concurrent window groups),
 * "pass2_minb" (1..3), "timing" (0/1), "msm_profile" (0/1), "msm_tables" (0/1: precomputed window multiples of
 * large SRS slots, W x the slot's memory, built on first use), "msm_tables_min" (smallest such slot), "binv_direct",
 * "msm_preplan" (0/1: the rounds of the persistent accumulation kernel planned ahead of its launch), "msm_sort_ahead"
 * (0/1: in a batched call the next MSM is sorted and planned on a side stream), "prove_joint" (-1 automatic, 0, 1:
 * commit_p as ONE multi_scalar_mul over g_m | g_q), further tuning knobs listed in INTEGRATION.md section 5.
 * Unknown name -> DVP_ERR_BAD_ARG. */
int dvp_ctx_set(dvp_ctx *ctx, const char *name, long value);

/*
 * SRS slots: decode once, keep the affine points resident in HBM.
 * Replaces the per-prove `read_point_vec_from_file` + decode of g_m / g_q / g_k
 * (src/proving.rs:462,511,666-673; src/io_utils.rs:187-239).  `pts30` = n x 30 bytes, the payload
 * of a point-vector file after its u64 count.  On DVP_ERR_INVALID_POINT *first_invalid (if not
 * NULL) receives the index of the first encoding that failed to decode.
 */
int dvp_srs_load(dvp_ctx *ctx, int slot, const uint8_t *pts30, size_t n, int64_t *first_invalid);
/* Append to a slot (g_k_0 | g_k_1 | g_k_2 concatenation, src/proving.rs:666-673). */
int dvp_srs_append(dvp_ctx *ctx, int slot, const uint8_t *pts30, size_t n, int64_t *first_invalid);
/* Fill a slot with n uniformly random group elements, deterministic in (seed, index): synthetic SRS
 * for benchmarks and full-size tests (read them back with dvp_srs_read). */
int dvp_srs_random(dvp_ctx *ctx, int slot, size_t n, uint64_t seed);
/* slot[i] = scalars[i] * CurvePoint::generator(): the batched fixed-base multiplication of the SRS generation
 * (compute_srs_matrices, src/srs.rs:126-160; CurvePoint::generator / point_scalar_mul_gen, src/curve.rs:84-91,129-137).
 * scalars_mont: n x 4 u64 Montgomery limbs (host). */
int dvp_srs_mulgen(dvp_ctx *ctx, int slot, const uint64_t *scalars_mont, size_t n);
int dvp_srs_size(dvp_ctx *ctx, int slot, size_t *n);
int dvp_srs_free(dvp_ctx *ctx, int slot);
/* Read points back as 30-byte encodings (CurvePoint::to_bytes, src/curve.rs:93-100). */
int dvp_srs_read(dvp_ctx *ctx, int slot, size_t offset, size_t n, uint8_t *pts30);

/*
 * multi_scalar_mul(&[Fr], &[CurvePoint]) -> CurvePoint  (src/curve.rs:141-158).
 * scalars_mont: n x 4 u64 Montgomery limbs (host).  Points: slot[offset .. offset+n).
 * out30 = CurvePoint::to_bytes of the sum.  n = 0 gives the neutral (30 zero bytes).
 */
int dvp_msm(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *scalars_mont, size_t n, uint8_t out30[30]);
/* Same with the scalars already in device memory (n x 32 bytes). */
int dvp_msm_device(dvp_ctx *ctx, int slot, size_t offset, const void *d_scalars_mont, size_t n, uint8_t out30[30]);
/*
 * nb calls of multi_scalar_mul over the SAME points (src/curve.rs:141-158 called once per scalar vector, as a prover
 * that serves several witnesses against one SRS does): scalars_mont[b] is the b-th vector (n x 4 u64 Montgomery
 * limbs; host pointers, or device pointers if scalars_on_device), out30 + 30 b receives its sum.  The calls are
 * pipelined: the device side of MSM b+1 is enqueued, and its scalars are uploaded on a copy stream into a second
 * staging buffer, while MSM b runs; the host folds the partial sums of MSM b meanwhile.  Results are the same bytes
 * as nb separate dvp_msm calls.
 */
int dvp_msm_batch(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *const *scalars_mont, size_t n, size_t nb,
                  int scalars_on_device, uint8_t *out30);
/* One-shot form with encoded points from the host: decode + MSM (src/srs.rs:422). */
int dvp_msm_adhoc(dvp_ctx *ctx, const uint8_t *pts30, const uint64_t *scalars_mont, size_t n, uint8_t out30[30]);
int dvp_msm_last_stats(dvp_ctx *ctx, dvp_msm_stats *out);
/* Development aid: with the "msm_profile" knob set the MSM runs on one lane with CUDA events around every
 * launch; ms/count per category: 0 sort 1 plan 2 pass1 3 binv_up 4 binv_direct 5 binv_down 6 pass2 7 misc. */
int dvp_msm_last_profile(dvp_ctx *ctx, float ms[8], unsigned count[8]);
/* ... and the same brackets as a timeline: rows of (lane, category, start ms, end ms) since the start of the MSM */
int dvp_msm_last_timeline(dvp_ctx *ctx, float *rows4, size_t cap_rows, size_t *count);

/*
 * Multi-GPU: one process (context) per GPU.  The MSMs shard by contiguous point range, the row evaluation by row
 * range, the extends by polynomial; exchanges run over NCCL (NVLink / NVSwitch) on the context's stream.
 * Rank 0 makes a unique id, the host ships it to the other ranks by any channel, every rank calls dvp_comm_init.
 * With a communicator set, dvp_prover_create expects the SRS slots to hold THIS RANK'S part of the SRS and dvp_prove
 * returns the same proof on every rank:  g_m[lo, hi) with dvp_shard_range(nwires);  g_q[ilo, ihi) with
 * dvp_shard_range(n);  g_k as g_k_0[ilo, ihi) | g_k_1[ilo, ihi) | g_k_2[2 ilo, 2 ihi) (the three files of
 * src/proving.rs:666-673 cut by the same index range of D, so that a rank's K scalars need only its own range of the
 * denominators).  dvp_setup fills the slots in exactly this layout.
 */
int dvp_comm_unique_id(uint8_t id[128]);
int dvp_comm_init(dvp_ctx *ctx, const uint8_t id[128], int rank, int world);
/* The same sharding without NCCL: `world` contexts of ONE process (on one device or several) become ranks
 * 0 .. world-1; every rank must be driven by its own host thread, the exchanges are device-to-device copies behind a
 * host rendezvous.  This is how the partition logic is checked on a single GPU. */
int dvp_comm_init_local(dvp_ctx *const *ctxs, int world);
int dvp_comm_destroy(dvp_ctx *ctx);
int dvp_comm_info(dvp_ctx *ctx, int *rank, int *world);
/* [lo, hi) of `total` items owned by `rank`: lo = floor(total * rank / world) */
void dvp_shard_range(size_t total, int rank, int world, size_t *lo, size_t *hi);
/* multi_scalar_mul over a sharded point vector: the slot holds this rank's n points, scalars_mont its n scalars
 * (host memory, or device memory if scalars_on_device); every rank receives the encoding of the whole sum. */
int dvp_msm_sharded(dvp_ctx *ctx, int slot, const uint64_t *scalars_mont, size_t n, int scalars_on_device,
                    uint8_t out30[30]);

/* The batched form (see dvp_msm_batch): the nb local MSMs are pipelined and ONE all-gather carries the nb partial sums
 * of every rank (nb x 80 bytes per rank). */
int dvp_msm_sharded_batch(dvp_ctx *ctx, int slot, const uint64_t *const *scalars_mont, size_t n, size_t nb,
                          int scalars_on_device, uint8_t *out30);

/* CurvePoint::add on encodings (src/curve.rs:76-82): out = a (+) b.  Runs on the device of ctx. */
int dvp_point_add(dvp_ctx *ctx, const uint8_t a30[30], const uint8_t b30[30], uint8_t out30[30]);

/* Device-memory helpers so that callers (benchmarks, language bindings) can keep inputs resident. */
int dvp_dev_alloc(dvp_ctx *ctx, size_t bytes, void **dptr);
int dvp_dev_free(dvp_ctx *ctx, void *dptr);
int dvp_dev_upload(dvp_ctx *ctx, void *dptr, const void *host, size_t bytes);
int dvp_dev_download(dvp_ctx *ctx, void *host, const void *dptr, size_t bytes);

/* Self-test kernels: evaluate the device field/point primitives on caller data (parity tests).
 * op: 0 gf_mul(a,b)  1 gf_sqr(a)  2 gf_inv(a)  3 fr_mul(a,b)  4 fr_to_canonical(a)  5 affine add (a,b = 64-byte points)
 * a, b, out: n elements of 32 bytes (64 for op 5), host memory. */
int dvp_selftest_op(dvp_ctx *ctx, int op, const void *a, const void *b, void *out, size_t n);
/* Integer-pipe microbenchmark: op 0 gf_mul chain, 1 gf_sqr chain, 2 fr_mul chain; returns ops per second. */
int dvp_microbench(dvp_ctx *ctx, int op, int iters, double *ops_per_sec);

/*
 * Evaluation domain of the prover: D = even leaves, D' = odd leaves of the 2n-leaf ECFFT tree built from
 * the constants at src/ec_fft.rs:205-229 (build_sect_ecfft_tree / get_both_domains, ec_fft.rs:93-239).
 * Also holds the extend matrices (FFTree.decompose/recombine_matrices) and the domain-specific prover
 * precomputes bar_wts and z_vals2inv (prover_prepares_precomputes, src/proving.rs:225-325), all on device.
 */
typedef struct dvp_domain dvp_domain;
int dvp_domain_create(dvp_ctx *ctx, unsigned log2_2n, dvp_domain **out);
/* read_minimal_fftree_from_file (src/tree_io.rs:419-433, called at src/proving.rs:436) on the image of a "tree2n" file:
 * the size comes from the file's FLeaves section, the extend tables are rebuilt on the device (upstream's matrices are
 * an internal convention of crate ecfft; the extend result is unique given the leaves) and the device's leaves must
 * equal the file's, else DVP_ERR_DOMAIN_MISMATCH.  A malformed file gives DVP_ERR_BAD_ARG (reference: bail!/expect). */
int dvp_domain_from_fftree(dvp_ctx *ctx, const uint8_t *file, size_t len, dvp_domain **out);
void dvp_domain_destroy(dvp_domain *dom);
int dvp_domain_leaves(dvp_domain *dom, uint64_t *leaves_mont /* 2n x 4 */);
int dvp_domain_precomputes(dvp_domain *dom, uint64_t *z_vals2inv /* n x 4 or NULL */, uint64_t *bar_wts /* n x 4 or NULL */);
/* z_poly.evaluate(x) for the vanishing polynomial of D (shift 0) or D' (shift 1) (ec_fft.rs:475) */
int dvp_domain_vanish_at(dvp_domain *dom, int shift, const uint64_t x_mont[4], uint64_t out_mont[4]);

/* FFTree::extend(evals, Moiety::S1) (src/proving.rs:410-422): npoly vectors of n Fr on D -> values on D'. */
int dvp_ecfft_extend(dvp_domain *dom, const uint64_t *in, uint64_t *out, int npoly);
/* In place on device memory (npoly x n x 32 bytes, contiguous). */
int dvp_ecfft_extend_device(dvp_domain *dom, void *d_data, int npoly);

/*
 * FFTree::enter (crate ecfft; reference call sites src/ec_fft.rs:317,411): coefficients of a polynomial of degree < n
 * (n x 4 u64 Montgomery, low degree first) -> its values on the n leaves x(C + i G_n) of the n-leaf tree, natural order.
 * FFTree::exit (src/ec_fft.rs:266,897) is the inverse: values on the n leaves -> coefficients.
 * The plan holds the trees with 4 .. n leaves (plain and shifted coset) and the per-level constants of exit.
 * (Setup and prover do not need either transform: they obtain Z_D(tau), Z'_D(d_i), Z_D(d'_i), L_i(tau) from the
 * chain rule, see dvp_setup.)
 */
typedef struct dvp_ecfft_plan dvp_ecfft_plan;
int dvp_ecfft_plan_create(dvp_ctx *ctx, unsigned log2_n, dvp_ecfft_plan **out);
void dvp_ecfft_plan_destroy(dvp_ecfft_plan *plan);
int dvp_ecfft_enter(dvp_ecfft_plan *plan, const uint64_t *coeffs, uint64_t *evals);
int dvp_ecfft_exit(dvp_ecfft_plan *plan, const uint64_t *evals, uint64_t *coeffs);

/*
 * R1CS in the dump's own order (src/gnark_r1cs.rs:1-20): three CSR matrices L, R, O over one coefficient
 * table (Montgomery limbs).  nrows is padded to n = next_power_of_two (gnark_r1cs.rs:291); the
 * Vandermonde block of update_to_include_vandermode_matrix_d (gnark_r1cs.rs:333-386) is applied on the fly.
 */
typedef struct dvp_r1cs dvp_r1cs;
int dvp_r1cs_load(dvp_ctx *ctx, size_t nrows, size_t num_public, size_t nwires, const uint32_t *const rowptr[3],
                  const uint32_t *const wire[3], const uint32_t *const coeff[3], const uint64_t *coeffs_mont,
                  size_t ncoeffs, dvp_r1cs **out);
void dvp_r1cs_destroy(dvp_r1cs *r1cs);
/* get_matrix_evaluations_from_witness (src/proving.rs:348-403): assignment = [1, public.., private..] (nwires x 4);
 * a, b, c, i: n x 4 each.  DVP_ERR_UNSATISFIED and *first_bad_row on a row with a*b != c + i. */
int dvp_r1cs_eval(dvp_r1cs *r1cs, dvp_domain *dom, const uint64_t *assignment, uint64_t *a, uint64_t *b, uint64_t *c,
                  uint64_t *i, int64_t *first_bad_row);

/* Row evaluation alone with the outputs left on the device: average milliseconds over reps runs (CUDA events). */
int dvp_r1cs_eval_time(dvp_r1cs *r1cs, dvp_domain *dom, const uint64_t *assignment, int reps, float *ms);
/* Synthetic circuits only (benchmarks, full-size tests; dv-pari_b200/synth.py): every row's O side ends with the
 * row's own fresh wire 1 + num_public + row (coefficient one) and rows read fresh wires of lower levels
 * (level = row mod nlevels) only.  Fills those wires in place so that every row holds. */
int dvp_r1cs_synth_solve(dvp_r1cs *r1cs, uint64_t *assignment /* nwires x 4, in place */, unsigned nlevels);

/*
 * Setup on the device (SRS::verifier_runs_setup, src/srs.rs:177-361; compute_srs_matrices / accumulate_m_values,
 * src/srs.rs:53-167).  trapdoor_mont = tau | delta | epsilon, 3 x 4 u64 Montgomery limbs.  L_i(tau), Z_D(tau), the
 * barycentric weights and Z on the other half-domain come from the chain rule of the isogeny tower in O(n log n)
 * instead of vanish / exit / enter; the points from the batched fixed-base multiplication.
 */
int dvp_setup_scalars(dvp_r1cs *r1cs, dvp_domain *dom, const uint64_t trapdoor_mont[12], uint64_t *sc_m /* nwires x 4 */,
                      uint64_t *sc_q /* n x 4 */, uint64_t *sc_k /* 4n x 4 */);
/* the SRS itself, left resident: the slots receive this rank's range of g_m, g_q and g_k_0|g_k_1|g_k_2 */
int dvp_setup(dvp_r1cs *r1cs, dvp_domain *dom, const uint64_t trapdoor_mont[12], int slot_gm, int slot_gq, int slot_gk);

/*
 * Proof::prove(cache_dir, public_inputs, private_inputs) (src/proving.rs:426-688) with the artifacts resident:
 * SRS slots hold g_m (nwires points), g_q (n) and g_k_0|g_k_1|g_k_2 (4n).
 * proof118 = commit_p (30) | kzg_k (30) | a0 (29 bytes LE) | b0 (29 bytes LE)  -- the fields of `Proof`
 * (proving.rs:40-50; FrBits bit i = bit i of the little-endian bytes).
 */
typedef struct dvp_prover dvp_prover;
int dvp_prover_create(dvp_ctx *ctx, dvp_domain *dom, dvp_r1cs *r1cs, int slot_gm, int slot_gq, int slot_gk,
                      dvp_prover **out);
void dvp_prover_destroy(dvp_prover *p);
int dvp_prove(dvp_prover *p, const uint64_t *public_mont, size_t k, const uint64_t *private_mont, size_t npriv,
              uint8_t proof118[118]);
/* Same, also returning the intermediate vectors for parity tests:
 * stages = 13 n x 4 u64: a b c i a' b' c' i' q k_a k_b k_r(2n). */
int dvp_prove_stages(dvp_prover *p, const uint64_t *public_mont, size_t k, const uint64_t *private_mont, size_t npriv,
                     uint8_t proof118[118], uint64_t *stages);
/*
 * SRS::verify(trapdoor, public_inputs, proof) (src/srs.rs:374-428): the designated verifier's check; the two-term
 * multi_scalar_mul of src/srs.rs:422 runs on the device.  *accepted = 1 iff the proof verifies (a field or point of
 * the proof that does not decode gives 0, like the reference's `all_inputs_valid`).
 */
int dvp_verify(dvp_ctx *ctx, const uint64_t trapdoor_mont[12], const uint64_t *public_mont, size_t k,
               const uint8_t proof118[118], int *accepted);
/* ms per stage of the last prove: r1cs rows, msm g_m, extend+quotient, msm g_q, challenge+K scalars, msm g_k,
 * witness upload (host -> device) */
int dvp_prove_last_times(dvp_prover *p, float ms[7]);

/*
 * Artifact formats either side of the path (host-side conversions into the layouts above; no device work).
 *   Fr vector files   u64 LE count | 29-byte LE canonical elements          (src/io_utils.rs:27-66,113-165)
 *   SP1 / gnark dump  u32 nbCoeffs | 32-byte BE coefficients | u32 nbRows | (nL nR nO | (u32 wire, u32 coeff)...)
 *                                                                            (src/gnark_r1cs.rs:1-20,121-185)
 *   witness file      u32 BE count | 32-byte BE elements                    (src/gnark_r1cs.rs:58-77,188-199)
 *   FFTree file       "FFTR\0\0\0\0" | u64 LE byte count | node; node = u32 section count, u32 pad | section metas
 *                     (u8 id, 7 pad, u64 offset from the node's start, u64 length) | blobs; ids 0 f, 1 recombine_matrices,
 *                     2 decompose_matrices, 3..11 the enter/exit tables, 12 the child FFTree as a nested node; blobs are
 *                     ark-serialize compressed: u64 LE count | 29-byte Fr (x4 per Mat2x2)       (src/tree_io.rs:1-15,32-118)
 * A point-vector file is u64 LE count | 30-byte encodings: its payload goes to dvp_srs_load unchanged.
 */
int dvp_fr_from_le29(const uint8_t *in, size_t n, uint64_t *out_mont);           /* Fr::deserialize_uncompressed */
int dvp_fr_to_le29(const uint64_t *in_mont, size_t n, uint8_t *out);             /* Fr::serialize_uncompressed */
int dvp_fr_from_be32_mod_order(const uint8_t *in, size_t n, uint64_t *out_mont); /* Fr::from_be_bytes_mod_order */
/* sp1_generate_scalar_from_raw_public_input (src/gnark_r1cs.rs:218-236) */
int dvp_sp1_public_input(uint64_t raw, uint64_t out_mont[4]);
/* blake3::hash (any length) and Transcript::output (src/proving.rs:137-197): alpha from commit_p and the k public
 * inputs, exactly as dvp_prove / dvp_verify derive it.  Host code, no device needed. */
int dvp_blake3(const uint8_t *data, size_t len, uint8_t out32[32]);
int dvp_transcript_alpha(const uint8_t commit_p30[30], const uint64_t *public_mont, size_t k, uint64_t alpha_mont[4]);
/* load_sparse_r1cs_from_file in two passes over the file image: sizes, then CSR per matrix + coefficient table */
int dvp_r1cs_dump_sizes(const uint8_t *buf, size_t len, size_t *ncoeffs, size_t *nrows, size_t nnz[3], size_t *max_wire);
int dvp_r1cs_dump_parse(const uint8_t *buf, size_t len, size_t ncoeffs, size_t nrows, const size_t nnz[3],
                        uint32_t *const rowptr[3], uint32_t *const wire[3], uint32_t *const coeff[3],
                        uint64_t *coeffs_mont); /* sizes = what dvp_r1cs_dump_sizes reported; re-validated */
/* FFTree file image: the section table of the node `depth` subtrees below the root (offsets absolute in the image,
 * absent section = length 0) -- read_fftree_from_slice's header walk, src/tree_io.rs:243-261 */
int dvp_fftree_file_sections(const uint8_t *file, size_t len, size_t depth, uint64_t off[13], uint64_t slen[13]);
/* tree.f.leaves() of that node: *n_leaves, and the leaves as Montgomery limbs if leaves_mont != NULL */
int dvp_fftree_file_leaves(const uint8_t *file, size_t len, size_t depth, size_t *n_leaves, uint64_t *leaves_mont);
/* tree.recombine_matrices (which = 1) / decompose_matrices (which = 2): *count Mat2x2 of the heap array, entries as
 * Montgomery limbs (count x 16 u64) if out != NULL -- the other two sections read_minimal_fftree_from_slice loads */
int dvp_fftree_file_matrices(const uint8_t *file, size_t len, size_t depth, int which, size_t *count, uint64_t *out);

/* Single-warp latency of a dependent chain, microseconds per op: mode 0 gf inversion by squarings,
 * 1 table-driven gf inversion, 2 gf multiplication. */
int dvp_latency_probe(dvp_ctx *ctx, int mode, int iters, float *us_per_op);
/* Raw integer-pipe issue rates (thread-instructions per second over the whole GPU): the roofline
 * denominators for the field kernels.  mode: 0 IMAD.WIDE  1 LOP3  2 IMAD  3 IMAD.WIDE:LOP3 = 1:2  4 SHF  5 IADD */
int dvp_pipebench(dvp_ctx *ctx, int mode, int iters, int blocks_per_sm, double *instr_per_sec);
/* Same source as the kernels, evaluated on the CPU; used only by the CPU-side parity tests. */
int dvp_hostcheck_op(int op, const void *a, const void *b, void *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
