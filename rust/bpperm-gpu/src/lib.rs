//! Drop-in host layer for ercembu/bulletproof-perm over libbpperm_cuda.so.  NOT compiled in the build image
//! (no Rust toolchain there); kept mechanical so that it can be checked by reading it against include/bpperm.h.
use bpperm_sys as sys;
use bulletproofs::ProofError;
use curve25519_dalek_ng::ristretto::{CompressedRistretto, RistrettoPoint};
use curve25519_dalek_ng::scalar::Scalar;
use curve25519_dalek_ng::traits::VartimeMultiscalarMul;
use once_cell::sync::Lazy;
use std::borrow::Borrow;
use std::sync::Mutex;

struct Ctx(*mut sys::bpp_ctx);
unsafe impl Send for Ctx {}
/// The trait is stateless (associated functions), so the GPU context is a lazily created process global.
static CTX: Lazy<Mutex<Ctx>> = Lazy::new(|| {
    let mut c = std::ptr::null_mut();
    let rc = unsafe { sys::bpp_init(0, &mut c) };
    assert_eq!(rc, 0, "bpp_init failed: no sm_100 GPU (this backend has no CPU fallback)");
    Mutex::new(Ctx(c))
});
fn with_ctx<T>(f: impl FnOnce(*mut sys::bpp_ctx) -> T) -> T { f(CTX.lock().unwrap().0) }

/// `RistrettoPoint::vartime_multiscalar_mul(..)` -> `GpuRistretto::vartime_multiscalar_mul(..)`
/// (circuit_lib.rs:187,202,216,363,374,385,396,407,498,504,509,525,535,552,568).
pub struct GpuRistretto;
impl VartimeMultiscalarMul for GpuRistretto {
    type Point = RistrettoPoint;
    fn optional_multiscalar_mul<I, J>(scalars: I, points: J) -> Option<RistrettoPoint>
    where I: IntoIterator, I::Item: Borrow<Scalar>, J: IntoIterator<Item = Option<RistrettoPoint>> {
        let s: Vec<u8> = scalars.into_iter().flat_map(|s| s.borrow().to_bytes().to_vec()).collect();
        let p: Vec<RistrettoPoint> = points.into_iter().collect::<Option<Vec<_>>>()?;
        assert_eq!(s.len() / 32, p.len()); // dalek asserts on the iterators' size hints
        // RistrettoPoint is a transparent wrapper of EdwardsPoint { X, Y, Z, T: FieldElement51([u64; 5]) }: 160 B
        let raw = unsafe { std::slice::from_raw_parts(p.as_ptr() as *const u8, p.len() * 160) };
        let mut out = [0u8; 32];
        let rc = with_ctx(|c| unsafe {
            sys::bpp_msm_vartime_host(c, s.as_ptr(), p.len(), sys::BPP_FMT_DALEK_XYZT, raw.as_ptr(), p.len(), out.as_mut_ptr())
        });
        assert_eq!(rc, 0);
        CompressedRistretto(out).decompress()
    }
}

fn bytes_of(v: &[Scalar]) -> &[u8] { unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, v.len() * 32) } }

/// util.rs:84-94 (same panic text on a length mismatch)
pub fn inner_product(a: &Vec<Scalar>, b: &Vec<Scalar>) -> Scalar {
    let mut out = [0u8; 32];
    let rc = with_ctx(|c| unsafe { sys::bpp_inner_product(c, bytes_of(a).as_ptr(), a.len(), bytes_of(b).as_ptr(), b.len(), out.as_mut_ptr()) });
    if rc == sys::BPP_ERR_LENGTH_MISMATCH { panic!("inner_product(a,b): lengths dont match, {}, {}", a.len(), b.len()); }
    assert_eq!(rc, 0);
    Scalar::from_canonical_bytes(out).unwrap()
}
/// util.rs:6-20
pub fn hadamard_V(a: &Vec<Scalar>, b: &Vec<Scalar>) -> Vec<Scalar> {
    let mut out = vec![Scalar::zero(); a.len()];
    let rc = with_ctx(|c| unsafe { sys::bpp_hadamard_V(c, bytes_of(a).as_ptr(), a.len(), bytes_of(b).as_ptr(), b.len(), out.as_mut_ptr() as *mut u8) });
    if rc == sys::BPP_ERR_LENGTH_MISMATCH { panic!("hadamard_V(a,b): lengths dont match, {}, {}", a.len(), b.len()); }
    assert_eq!(rc, 0);
    out
}
/// util.rs:63-65,139-157: exp_iter(x).take(n) with the reference iterator's (Fibonacci) exponents
pub fn exp_iter_take(x: Scalar, n: usize) -> Vec<Scalar> {
    let mut out = vec![Scalar::zero(); n];
    let rc = with_ctx(|c| unsafe { sys::bpp_exp_iter(c, x.as_bytes().as_ptr(), n, out.as_mut_ptr() as *mut u8) });
    assert_eq!(rc, 0);
    out
}

/// `count` ACProvers sharing one ACEssentials (circuit_lib.rs:58-88), proved and verified in lock-step.
pub struct ShuffleBatch { b: *mut sys::bpp_acp_batch, count: usize, proof_len: usize }
impl ShuffleBatch {
    pub fn new(cir: *const sys::bpp_circuit, gens: *const sys::bpp_gens, n: usize, mode: i32, count: usize, label: &[u8]) -> Self {
        let mut b = std::ptr::null_mut();
        let rc = with_ctx(|c| unsafe { sys::bpp_acp_batch_create(c, cir, gens, mode, count, label.as_ptr(), label.len(), &mut b) });
        assert_eq!(rc, 0);
        ShuffleBatch { b, count, proof_len: unsafe { sys::bpp_acproof_proof_len_mode(n, mode) } }
    }
    /// create .. blinding_values (lib.rs:219-228); seeds: ChaCha20Rng::from_seed per proof instead of thread_rng()
    pub fn prove(&mut self, a_l: &[Scalar], a_r: &[Scalar], a_o: &[Scalar], gamma: &[Scalar], seeds: &[[u8; 32]]) -> Vec<u8> {
        let mut proofs = vec![0u8; self.count * self.proof_len];
        unsafe {
            assert_eq!(sys::bpp_acp_batch_upload_witness(self.b, bytes_of(a_l).as_ptr(), bytes_of(a_r).as_ptr(), bytes_of(a_o).as_ptr(),
                                                         bytes_of(gamma).as_ptr(), seeds.as_ptr() as *const u8), 0);
            assert_eq!(sys::bpp_acp_batch_prove(self.b), 0);
            assert_eq!(sys::bpp_acp_batch_download_proofs(self.b, proofs.as_mut_ptr()), 0);
        }
        proofs
    }
    /// verify (lib.rs:230): one Result per proof
    pub fn verify(&mut self, proofs: &[u8], v: &[CompressedRistretto], verifier_seed: &[u8; 32]) -> Vec<Result<(), ProofError>> {
        let mut acc = vec![0u8; self.count];
        unsafe {
            assert_eq!(sys::bpp_acp_batch_upload_proofs(self.b, proofs.as_ptr(), v.as_ptr() as *const u8), 0);
            assert_eq!(sys::bpp_acp_batch_verify(self.b, verifier_seed.as_ptr()), 0);
            assert_eq!(sys::bpp_acp_batch_download_accept(self.b, acc.as_mut_ptr()), 0);
        }
        acc.into_iter().map(|a| if a == 1 { Ok(()) } else { Err(ProofError::VerificationError) }).collect()
    }
}
impl Drop for ShuffleBatch { fn drop(&mut self) { unsafe { sys::bpp_acp_batch_free(self.b) } } }

// ---- proof records on the wire (bpperm.h: bpp_acproof_to_wire / bpp_acproof_from_wire) ------------------------
// mode 2 records are bulletproofs 4.0.0 `R1CSProof::to_bytes` of a one-phase proof (version byte 0).
pub fn proofs_to_wire(n: usize, mode: i32, count: usize, proofs: &[u8]) -> Vec<u8> {
    let wlen = unsafe { sys::bpp_acproof_wire_len(n, mode) };
    let mut out = vec![0u8; count * wlen];
    assert_eq!(unsafe { sys::bpp_acproof_to_wire(n, mode, count, proofs.as_ptr(), out.as_mut_ptr()) }, 0);
    out
}
/// One `Result` per record: `Err(ProofError::FormatError)` for a wrong version byte or a non-canonical scalar
/// (what `R1CSProof::from_bytes` rejects); the proof bytes of a rejected record are zeroed.
pub fn proofs_from_wire(n: usize, mode: i32, count: usize, wire: &[u8]) -> Result<(Vec<u8>, Vec<Result<(), ProofError>>), ProofError> {
    let plen = unsafe { sys::bpp_acproof_proof_len_mode(n, mode) };
    let mut proofs = vec![0u8; count * plen];
    let mut status = vec![0u8; count];
    if count == 0 || wire.len() % count != 0 { return Err(ProofError::FormatError); }
    let rc = unsafe { sys::bpp_acproof_from_wire(n, mode, count, wire.as_ptr(), wire.len() / count, proofs.as_mut_ptr(), status.as_mut_ptr()) };
    if rc != 0 { return Err(ProofError::FormatError); }          // record length does not match the circuit
    Ok((proofs, status.into_iter().map(|s| if s == 0 { Ok(()) } else { Err(ProofError::FormatError) }).collect()))
}

// ---- many independent large MSMs: two in flight (bpp_msm_submit_dev / bpp_msm_wait_previous / bpp_msm_wait) ---
// `d_scalars[i]` / `d_out[i & 1]` are device pointers of the caller (e.g. from cust / cudarc); `consume(i, ptr)`
// enqueues whatever reads result i on the context's stream.  Same bytes as one bpp_msm_vartime_dev call per MSM.
pub unsafe fn msm_stream(points: *const sys::bpp_points, n: usize, d_scalars: &[*const core::ffi::c_void],
                         d_out: [*mut core::ffi::c_void; 2], mut consume: impl FnMut(usize, *mut core::ffi::c_void)) {
    with_ctx(|ctx| {
        for (i, sc) in d_scalars.iter().enumerate() {
            assert_eq!(sys::bpp_msm_submit_dev(ctx, *sc, points, 0, n, d_out[i & 1]), 0);
            assert_eq!(sys::bpp_msm_wait_previous(ctx), 0);      // results up to i - 1 are valid in stream order
            if i > 0 { consume(i - 1, d_out[(i - 1) & 1]); }
        }
        assert_eq!(sys::bpp_msm_wait(ctx), 0);
        if !d_scalars.is_empty() { consume(d_scalars.len() - 1, d_out[(d_scalars.len() - 1) & 1]); }
    })
}
