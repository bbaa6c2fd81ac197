// Addition to bn254/src/curve.rs (inside `impl G1`, after `to_bytes`, curve.rs:137-139).
//
// `G1` wraps `pub(crate) Halo2G1` (curve.rs:74), so a crate outside p3-bn254 cannot build one from coordinates.
// The GPU library returns commitments / witnesses as affine points whose coordinates are the Montgomery limbs of
// halo2curves' `Fq` (the same 4 x u64 little-endian layout `Fr` uses, field.rs:96-105); identity = all zero.
// halo2curves keeps `Fq(pub(crate) [u64; 4])`, hence the transmutes -- the same device the crate already uses for
// `Fr` (`fr_to_halo2`, curve.rs:466-469).

use halo2curves::bn256::{Fq as Halo2Fq, G1Affine as Halo2G1Affine};
use halo2curves::group::prime::PrimeCurveAffine;

impl G1 {
    /// Affine point from Montgomery-form coordinate limbs; (0, 0) is the identity.
    /// Panics if the point is not on the curve (the GPU library only returns curve points).
    pub fn from_affine_montgomery_limbs(x: [u64; 4], y: [u64; 4]) -> Self {
        if x == [0; 4] && y == [0; 4] {
            return Self::identity();
        }
        // SAFETY: halo2curves' Fq is a transparent wrapper of [u64; 4] holding the Montgomery form.
        let (fx, fy): (Halo2Fq, Halo2Fq) = unsafe { (core::mem::transmute(x), core::mem::transmute(y)) };
        let p = Halo2G1Affine::from_xy(fx, fy).expect("point returned by the GPU library is on the curve");
        Self(p.to_curve())
    }

    /// Affine Montgomery-form coordinate limbs (the inverse of `from_affine_montgomery_limbs`); one field
    /// inversion.  Used to hand an SRS produced on the CPU (`init_srs_unsafe`) to the GPU library.
    pub fn to_affine_montgomery_limbs(&self) -> ([u64; 4], [u64; 4]) {
        if self.is_identity() {
            return ([0; 4], [0; 4]);
        }
        let a = self.0.to_affine();
        // SAFETY: as above.
        unsafe { (core::mem::transmute(a.x), core::mem::transmute(a.y)) }
    }
}
