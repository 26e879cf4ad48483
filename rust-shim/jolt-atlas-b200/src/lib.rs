//! B200 back end for jolt-atlas behind the reference's two seams.  Not compiled in this repository (no Rust
//! toolchain in the build image); written against the reference at 434ab99.
pub mod hyperkzg_b200;
pub mod sumcheck_b200;

use ark_bn254::Fr;
use jolt_atlas_b200_sys as sys;
use std::ffi::CStr;

/// Every C entry returns a status; the reference prover panics on invariant violations, so the prover side panics too
/// (nothing unwinds across the ABI).  A verifier-side caller would map to `ProofVerifyError::InternalError`.
pub fn check(st: i32) {
    if st != sys::JA_OK {
        let mut buf = [0 as std::os::raw::c_char; 1024];
        unsafe { sys::ja_last_error(buf.as_mut_ptr(), buf.len()) };
        panic!("jolt_atlas_b200 error {st}: {}", unsafe { CStr::from_ptr(buf.as_ptr()) }.to_string_lossy());
    }
}

/// `ark_bn254::Fr` is 4 x u64 little-endian Montgomery limbs (`BigInt<4>`); the reference transmutes it the same way
/// (joltworks/src/field/ark.rs:21-29).
#[inline]
pub fn limbs(x: &Fr) -> [u64; 4] {
    x.0 .0
}
#[inline]
pub fn fr_from_limbs(l: &[u64]) -> Fr {
    ark_ff::Fp::new_unchecked(ark_ff::BigInt::new([l[0], l[1], l[2], l[3]]))
}

/// One library context per prover (`ja_init`); calls are serialised inside the library, so the handle may be shared by the rayon
/// workers that call `PCS::commit` (jolt-atlas-core/src/onnx_proof/prover.rs:243-248).
pub struct Ctx(pub *mut std::os::raw::c_void);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::ja_init(device, &mut h) });
        Ctx(h)
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::ja_shutdown(self.0) }
    }
}
