//! `GpuDft`: `TwoAdicSubgroupDft<Fr>` on the GPU (dft/src/traits.rs:27-249).
//!
//! The five batch methods are one library call each; everything else (`dft`, `coset_lde`, the `*_algebra_*`
//! forms, `lde_batch`) keeps the trait's default, which routes through them.  Results are the reference's
//! canonical limbs in natural row order (`Evaluations = RowMajorMatrix<Fr>`).
use eon_kzg_sys as sys;
use p3_bn254::Fr;
use p3_dft::TwoAdicSubgroupDft;
use p3_field::PrimeCharacteristicRing;
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;
use p3_util::log2_strict_usize;

use crate::{check, ctx, fr_limbs, fr_mut_ptr, fr_ptr};

#[derive(Clone, Copy, Debug, Default)]
pub struct GpuDft;

impl GpuDft {
    fn out_like(rows: usize, width: usize) -> Vec<Fr> {
        Fr::zero_vec(rows * width)
    }
}

impl TwoAdicSubgroupDft<Fr> for GpuDft {
    type Evaluations = RowMajorMatrix<Fr>;

    /// dft/src/traits.rs:61 (semantics: dft/src/naive.rs:15-31)
    fn dft_batch(&self, mat: RowMajorMatrix<Fr>) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let log_h = log2_strict_usize(h); // same panic as the reference on a non power of two
        let mut out = Self::out_like(h, w);
        check(unsafe { sys::eon_mctx_dft_batch(ctx(), fr_ptr(&mat.values), fr_mut_ptr(&mut out), log_h as u32, w) });
        RowMajorMatrix::new(out, w)
    }

    /// dft/src/traits.rs:83-91
    fn coset_dft_batch(&self, mat: RowMajorMatrix<Fr>, shift: Fr) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let log_h = log2_strict_usize(h);
        let mut out = Self::out_like(h, w);
        let s = fr_limbs(&shift);
        check(unsafe {
            sys::eon_mctx_coset_dft_batch(ctx(), fr_ptr(&mat.values), fr_mut_ptr(&mut out), log_h as u32, w, s.as_ptr())
        });
        RowMajorMatrix::new(out, w)
    }

    /// dft/src/traits.rs:111-122
    fn idft_batch(&self, mat: RowMajorMatrix<Fr>) -> RowMajorMatrix<Fr> {
        let (h, w) = (mat.height(), mat.width());
        let log_h = log2_strict_usize(h);
        let mut out = Self::out_like(h, w);
        check(unsafe { sys::eon_mctx_idft_batch(ctx(), fr_ptr(&mat.values), fr_mut_ptr(&mut out), log_h as u32, w) });
        RowMajorMatrix::new(out, w)
    }

    /// dft/src/traits.rs:144-153
    fn coset_idft_batch(&self, mat: RowMajorMatrix<Fr>, shift: Fr) -> RowMajorMatrix<Fr> {
        let (h, w) = (mat.height(), mat.width());
        let log_h = log2_strict_usize(h);
        let mut out = Self::out_like(h, w);
        let s = fr_limbs(&shift);
        check(unsafe {
            sys::eon_mctx_coset_idft_batch(ctx(), fr_ptr(&mat.values), fr_mut_ptr(&mut out), log_h as u32, w, s.as_ptr())
        });
        RowMajorMatrix::new(out, w)
    }

    /// dft/src/traits.rs:226-249: iDFT, zero-pad to `h << added_bits` rows, coset DFT — fused on the device
    /// (inverse with bit-reversed output feeding a forward transform with bit-reversed input: no permutation pass).
    fn coset_lde_batch(&self, mat: RowMajorMatrix<Fr>, added_bits: usize, shift: Fr) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let log_h = log2_strict_usize(h);
        let mut out = Self::out_like(h << added_bits, w);
        let s = fr_limbs(&shift);
        check(unsafe {
            sys::eon_mctx_coset_lde_batch(
                ctx(),
                fr_ptr(&mat.values),
                fr_mut_ptr(&mut out),
                log_h as u32,
                w,
                added_bits as u32,
                s.as_ptr(),
            )
        });
        RowMajorMatrix::new(out, w)
    }
}
