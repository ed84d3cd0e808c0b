//! FFI of libdvpari (include/dvpari.h) and the thin safe layer the reference's call sites use.
//!
//! Replaced call sites (alpenlabs/dv-pari):
//!   curve::multi_scalar_mul            src/curve.rs:141   -> Gpu::multi_scalar_mul
//!   FFTree::extend(.., Moiety::S1)     src/proving.rs:412 -> Domain::extend
//!   Proof::prove                       src/proving.rs:426 -> Prover::prove
//!   SRS::verifier_runs_setup           src/srs.rs:177     -> setup
//! `Fr` is four u64 Montgomery limbs in memory, so `&[Fr]` is passed as `*const u64` without a copy.
#![allow(non_camel_case_types)]
use std::os::raw::{c_int, c_uint};

pub const DVP_OK: c_int = 0;

#[repr(C)] pub struct dvp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct dvp_domain { _p: [u8; 0] }
#[repr(C)] pub struct dvp_r1cs { _p: [u8; 0] }
#[repr(C)] pub struct dvp_prover { _p: [u8; 0] }

extern "C" {
    pub fn dvp_ctx_create(device: c_int, out: *mut *mut dvp_ctx) -> c_int;
    pub fn dvp_ctx_destroy(ctx: *mut dvp_ctx);
    pub fn dvp_srs_load(ctx: *mut dvp_ctx, slot: c_int, pts30: *const u8, n: usize, first_invalid: *mut i64) -> c_int;
    pub fn dvp_srs_append(ctx: *mut dvp_ctx, slot: c_int, pts30: *const u8, n: usize, first_invalid: *mut i64) -> c_int;
    pub fn dvp_srs_read(ctx: *mut dvp_ctx, slot: c_int, offset: usize, n: usize, pts30: *mut u8) -> c_int;
    pub fn dvp_srs_mulgen(ctx: *mut dvp_ctx, slot: c_int, scalars_mont: *const u64, n: usize) -> c_int;
    pub fn dvp_msm(ctx: *mut dvp_ctx, slot: c_int, offset: usize, scalars_mont: *const u64, n: usize, out30: *mut u8) -> c_int;
    pub fn dvp_msm_adhoc(ctx: *mut dvp_ctx, pts30: *const u8, scalars_mont: *const u64, n: usize, out30: *mut u8) -> c_int;
    pub fn dvp_msm_batch(ctx: *mut dvp_ctx, slot: c_int, offset: usize, scalars_mont: *const *const u64, n: usize, nb: usize,
                         scalars_on_device: c_int, out30: *mut u8) -> c_int;
    pub fn dvp_domain_create(ctx: *mut dvp_ctx, log2_2n: c_uint, out: *mut *mut dvp_domain) -> c_int;
    pub fn dvp_domain_from_fftree(ctx: *mut dvp_ctx, file: *const u8, len: usize, out: *mut *mut dvp_domain) -> c_int;
    pub fn dvp_domain_destroy(dom: *mut dvp_domain);
    pub fn dvp_ecfft_extend(dom: *mut dvp_domain, input: *const u64, output: *mut u64, npoly: c_int) -> c_int;
    pub fn dvp_r1cs_load(ctx: *mut dvp_ctx, nrows: usize, num_public: usize, nwires: usize,
                         rowptr: *const *const u32, wire: *const *const u32, coeff: *const *const u32,
                         coeffs_mont: *const u64, ncoeffs: usize, out: *mut *mut dvp_r1cs) -> c_int;
    pub fn dvp_r1cs_destroy(r: *mut dvp_r1cs);
    pub fn dvp_setup(r: *mut dvp_r1cs, dom: *mut dvp_domain, trapdoor_mont: *const u64,
                     slot_gm: c_int, slot_gq: c_int, slot_gk: c_int) -> c_int;
    pub fn dvp_prover_create(ctx: *mut dvp_ctx, dom: *mut dvp_domain, r: *mut dvp_r1cs,
                             slot_gm: c_int, slot_gq: c_int, slot_gk: c_int, out: *mut *mut dvp_prover) -> c_int;
    pub fn dvp_prover_destroy(p: *mut dvp_prover);
    pub fn dvp_prove(p: *mut dvp_prover, public_mont: *const u64, k: usize, private_mont: *const u64, npriv: usize,
                     proof118: *mut u8) -> c_int;
    pub fn dvp_verify(ctx: *mut dvp_ctx, trapdoor_mont: *const u64, public_mont: *const u64, k: usize,
                      proof118: *const u8, accepted: *mut c_int) -> c_int;
    pub fn dvp_comm_unique_id(id: *mut u8) -> c_int;
    pub fn dvp_comm_init(ctx: *mut dvp_ctx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn dvp_shard_range(total: usize, rank: c_int, world: c_int, lo: *mut usize, hi: *mut usize);
    pub fn dvp_msm_sharded(ctx: *mut dvp_ctx, slot: c_int, scalars_mont: *const u64, n: usize, scalars_on_device: c_int,
                           out30: *mut u8) -> c_int;
    pub fn dvp_msm_sharded_batch(ctx: *mut dvp_ctx, slot: c_int, scalars_mont: *const *const u64, n: usize, nb: usize,
                                 scalars_on_device: c_int, out30: *mut u8) -> c_int;
}

// The reference's Fr must be exactly four u64 limbs for the zero-copy cast below.
const _: () = assert!(core::mem::size_of::<[u64; 4]>() == 32);

/// One GPU with its resident SRS slots.
pub struct Gpu { raw: *mut dvp_ctx }

impl Gpu {
    pub fn new(device: i32) -> Result<Self, i32> {
        let mut raw = core::ptr::null_mut();
        match unsafe { dvp_ctx_create(device, &mut raw) } { DVP_OK => Ok(Gpu { raw }), e => Err(e) }
    }
    /// `payload` = the bytes of a point-vector file after its u64 count (src/io_utils.rs:187-239).
    pub fn load_points(&self, slot: i32, payload: &[u8]) -> Result<(), (i32, i64)> {
        assert!(payload.len() % 30 == 0);
        let mut bad = -1i64;
        match unsafe { dvp_srs_load(self.raw, slot, payload.as_ptr(), payload.len() / 30, &mut bad) } {
            DVP_OK => Ok(()), e => Err((e, bad)),
        }
    }
    /// curve::multi_scalar_mul (src/curve.rs:141-158); `scalars` = `&[Fr]` reinterpreted as limbs (4 per element).
    pub fn multi_scalar_mul(&self, slot: i32, scalars: &[[u64; 4]]) -> [u8; 30] {
        let mut out = [0u8; 30];
        let rc = unsafe { dvp_msm(self.raw, slot, 0, scalars.as_ptr() as *const u64, scalars.len(), out.as_mut_ptr()) };
        assert_eq!(rc, DVP_OK, "dvp_msm failed: {rc}"); // the reference panics on a length mismatch (curve.rs:142)
        out
    }
    /// The same for several scalar vectors of one length against the same points, pipelined on the device.
    pub fn multi_scalar_mul_many(&self, slot: i32, vectors: &[&[[u64; 4]]]) -> Vec<[u8; 30]> {
        let n = vectors.first().map_or(0, |v| v.len());
        assert!(vectors.iter().all(|v| v.len() == n), "scalar vectors of different lengths");
        let ptrs: Vec<*const u64> = vectors.iter().map(|v| v.as_ptr() as *const u64).collect();
        let mut out = vec![0u8; 30 * vectors.len()];
        let rc = unsafe { dvp_msm_batch(self.raw, slot, 0, ptrs.as_ptr(), n, ptrs.len(), 0, out.as_mut_ptr()) };
        assert_eq!(rc, DVP_OK, "dvp_msm_batch failed: {rc}");
        out.chunks_exact(30).map(|c| { let mut a = [0u8; 30]; a.copy_from_slice(c); a }).collect()
    }
    pub fn raw(&self) -> *mut dvp_ctx { self.raw }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { dvp_ctx_destroy(self.raw) } } }

/// Proof { commit_p, kzg_k, a0, b0 } (src/proving.rs:40-50) as the 118 bytes dvp_prove returns.
pub struct ProofBytes(pub [u8; 118]);
impl ProofBytes {
    pub fn commit_p(&self) -> &[u8] { &self.0[0..30] }
    pub fn kzg_k(&self) -> &[u8] { &self.0[30..60] }
    pub fn a0_le29(&self) -> &[u8] { &self.0[60..89] }
    pub fn b0_le29(&self) -> &[u8] { &self.0[89..118] }
}

/// Proof::prove (src/proving.rs:426-688) with the artifacts resident on the device.
pub struct Prover { raw: *mut dvp_prover }
impl Prover {
    pub unsafe fn from_raw(raw: *mut dvp_prover) -> Self { Prover { raw } }
    pub fn prove(&self, public_inputs: &[[u64; 4]], private_inputs: &[[u64; 4]]) -> Result<ProofBytes, i32> {
        let mut out = [0u8; 118];
        let rc = unsafe {
            dvp_prove(self.raw, public_inputs.as_ptr() as *const u64, public_inputs.len(),
                      private_inputs.as_ptr() as *const u64, private_inputs.len(), out.as_mut_ptr())
        };
        if rc == DVP_OK { Ok(ProofBytes(out)) } else { Err(rc) }
    }
}
impl Drop for Prover { fn drop(&mut self) { unsafe { dvp_prover_destroy(self.raw) } } }
