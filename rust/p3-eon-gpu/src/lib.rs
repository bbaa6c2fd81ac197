//! GPU-backed drop-ins for the two plugin traits of the BN254 KZG hot path of Plonky3-eon:
//!
//! * [`GpuDft`]    — `p3_dft::TwoAdicSubgroupDft<Fr>` (dft/src/traits.rs:27-249)
//! * [`GpuKzgPcs`] — `p3_commit::Pcs<Fr, Challenger>` with the associated types of `p3_kzg::KzgPcs`
//!                   (kzg/src/pcs.rs:209-216), so `StarkConfig`, `CanObserve<KzgCommitment>` and both provers
//!                   compile unchanged: `type MyPcs = GpuKzgPcs;`
//!
//! All compute happens in `libeon_kzg.so` (hand-written CUDA for sm_100a) through ONE process-wide
//! multi-device context (`eon_mctx_*`): the library shards the columns of every matrix over the GPUs named in
//! `EON_DEVICES` (default: device 0) and returns results that are byte-identical to the reference's.
//! There is no CPU fallback: without a usable sm_100 device every call panics.
//!
//! UNCOMPILED in the build image (no rustc): see rust/README.md for what is checked and how.
#![allow(clippy::missing_safety_doc)]

mod dft;
mod pcs;

use std::ffi::CStr;
use std::sync::OnceLock;

use eon_kzg_sys as sys;
use p3_bn254::Fr;

pub use dft::GpuDft;
pub use pcs::{GpuKzgPcs, GpuMatrixProverData};

/// `Fr { value: [u64; 4] }` (Montgomery, canonical) is passed through as its bytes (bn254/src/field.rs:96-105);
/// the reference itself transmutes it the same way (curve.rs:466-482).
#[inline]
pub(crate) fn fr_ptr(v: &[Fr]) -> *const u64 {
    v.as_ptr().cast()
}
#[inline]
pub(crate) fn fr_mut_ptr(v: &mut [Fr]) -> *mut u64 {
    v.as_mut_ptr().cast()
}
#[inline]
pub(crate) fn fr_limbs(x: &Fr) -> [u64; 4] {
    // SAFETY: Fr is a single-field struct around [u64; 4].
    unsafe { core::mem::transmute_copy(x) }
}

/// The process-wide multi-device context (TwoAdicSubgroupDft requires `Clone + Default`, so the state cannot
/// live in the Dft value; Radix2Dit keeps its twiddle cache behind shared state for the same reason,
/// dft/src/radix_2_dit.rs:33-58).
pub(crate) struct Context(pub *mut sys::eon_mctx);
// SAFETY: every eon_mctx_* entry point serialises on an internal mutex.
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

static CONTEXT: OnceLock<Context> = OnceLock::new();

pub(crate) fn ctx() -> *mut sys::eon_mctx {
    CONTEXT
        .get_or_init(|| {
            let devices: Vec<i32> = std::env::var("EON_DEVICES")
                .ok()
                .map(|s| s.split(',').filter_map(|t| t.trim().parse().ok()).collect())
                .filter(|v: &Vec<i32>| !v.is_empty())
                .unwrap_or_else(|| vec![0]);
            let mut m: *mut sys::eon_mctx = core::ptr::null_mut();
            let rc = unsafe { sys::eon_mctx_create(devices.as_ptr(), devices.len() as i32, &mut m) };
            assert!(
                rc == sys::EON_OK && !m.is_null(),
                "eon_mctx_create failed ({rc}): no usable sm_100 CUDA device; there is no CPU fallback"
            );
            Context(m)
        })
        .0
}

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::eon_mctx_last_error(ctx())) }.to_string_lossy().into_owned()
}

/// Prover-side failures panic, like the reference (`unwrap()` in kzg/src/pcs.rs:238-240,248).
#[track_caller]
pub(crate) fn check(rc: i32) {
    if rc != sys::EON_OK {
        panic!("libeon_kzg error {rc}: {}", last_error());
    }
}
