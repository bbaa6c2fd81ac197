//! `GpuKzgPcs`: the prover half of `KzgPcs as Pcs<Fr, Challenger>` (kzg/src/pcs.rs:209-335) on the GPU.
//!
//! Same associated types as `KzgPcs` except `ProverData` (opaque to callers, commit/src/pcs.rs:33): the
//! coefficient matrix stays in HBM behind a handle instead of a host `RowMajorMatrix`.  `verify` (pairings) is
//! delegated to the CPU `KzgPcs` unchanged.
use std::cell::RefCell;

use eon_kzg_sys as sys;
use p3_bn254::{Fr, G1};
use p3_commit::{OpenedValues, Pcs, PolynomialSpace};
use p3_field::coset::TwoAdicMultiplicativeCoset;
use p3_field::PrimeCharacteristicRing;
use p3_kzg::{
    KzgCommitment, KzgError, KzgPcs, KzgProof, MatrixCommitment, MatrixProof, PointProof, StructuredReferenceString,
};
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;
use p3_util::log2_strict_usize;

use crate::{check, ctx, fr_limbs, fr_mut_ptr, fr_ptr, last_error};

/// kzg/src/pcs.rs:52-61 with `coeffs` living on the device.  `evals` stays on the host for the same-domain fast
/// path of `get_evaluations_on_domain` (pcs.rs:274-276).
pub struct GpuMatrixProverData {
    pub domain: TwoAdicMultiplicativeCoset<Fr>,
    pub evals: RowMajorMatrix<Fr>,
    handle: sys::eon_handle,
    width: usize,
    /// evaluations on the hinted quotient coset, produced inside `commit` (see `GpuKzgPcs::with_lde_hint`)
    lde: RefCell<Option<(TwoAdicMultiplicativeCoset<Fr>, RowMajorMatrix<Fr>)>>,
}

impl GpuMatrixProverData {
    /// `MatrixProverData.coeffs` (pcs.rs:60): materialised on demand, nothing on the prover path reads it.
    pub fn coeffs(&self) -> RowMajorMatrix<Fr> {
        let mut out = Fr::zero_vec(self.evals.height() * self.width);
        check(unsafe { sys::eon_mctx_kzg_read_coeffs(ctx(), self.handle, fr_mut_ptr(&mut out)) });
        RowMajorMatrix::new(out, self.width)
    }
}

impl Drop for GpuMatrixProverData {
    fn drop(&mut self) {
        if self.handle != 0 {
            unsafe { sys::eon_mctx_handle_free(ctx(), self.handle) };
        }
    }
}

fn g1_from_wire(xy: &[u64]) -> G1 {
    // rust/patches/p3-bn254-g1-limbs.rs
    G1::from_affine_montgomery_limbs(xy[0..4].try_into().unwrap(), xy[4..8].try_into().unwrap())
}

#[derive(Clone)]
pub struct GpuKzgPcs {
    /// the CPU PCS: owns the SRS (`params`) and serves `verify`
    cpu: KzgPcs,
    /// (added_bits, shift) of the coset the prover evaluates every committed trace on right after committing it
    lde_hint: Option<(usize, Fr)>,
}

impl GpuKzgPcs {
    /// KzgPcs::from_srs (kzg/src/pcs.rs:170-176): the G1 powers are normalised to affine ONCE here and stay
    /// resident on every device (the reference re-normalises them in every multi_exp, bn254/src/curve.rs:170).
    pub fn from_srs(srs: StructuredReferenceString) -> Self {
        let mut xy = Vec::with_capacity(8 * srs.g1_powers.len());
        for p in &srs.g1_powers {
            let (x, y) = p.to_affine_montgomery_limbs();
            xy.extend_from_slice(&x);
            xy.extend_from_slice(&y);
        }
        check(unsafe { sys::eon_mctx_srs_load_affine(ctx(), xy.as_ptr(), srs.g1_powers.len()) });
        Self { cpu: KzgPcs::from_srs(srs), lde_hint: None }
    }

    /// KzgPcs::new (pcs.rs:198-203).  Testing only, like the reference.
    pub fn new(max_degree: usize, alpha: Fr) -> Self {
        Self::from_srs(p3_kzg::init_srs_unsafe(max_degree, alpha))
    }

    /// The uni-stark prover asks for the evaluations of every committed trace on the quotient coset
    /// (`trace_domain.create_disjoint_domain(n << added_bits)`: shift `Fr::GENERATOR`, commit/src/domain.rs:167)
    /// right after committing it (eon-uni-stark/src/prover.rs:186-187 then :307-322).  With the hint `commit`
    /// produces them in the same library call (the download hides under the MSM) and
    /// `get_evaluations_on_domain` hands them out.
    pub fn with_lde_hint(mut self, added_bits: usize, shift: Fr) -> Self {
        self.lde_hint = Some((added_bits, shift));
        self
    }

    fn degree_guard(&self, rc: i32, degree: usize) {
        if rc == sys::EON_ERR_SRS_TOO_SHORT {
            // `.ensure_supported(height - 1).unwrap()` (pcs.rs:238-240)
            panic!("{:?}", KzgError::DegreeTooLarge { degree, max: self.cpu.params.max_degree });
        }
        check(rc);
    }
}

impl<Challenger> Pcs<Fr, Challenger> for GpuKzgPcs {
    type Domain = TwoAdicMultiplicativeCoset<Fr>;
    type Commitment = KzgCommitment;
    type ProverData = Vec<GpuMatrixProverData>;
    type EvaluationsOnDomain<'a> = RowMajorMatrix<Fr>;
    type Proof = KzgProof;
    type Error = KzgError;

    const ZK: bool = false;

    /// pcs.rs:218-221
    fn natural_domain_for_degree(&self, degree: usize) -> Self::Domain {
        let log_n = log2_strict_usize(degree.next_power_of_two());
        TwoAdicMultiplicativeCoset::new(Fr::ONE, log_n).expect("valid domain")
    }

    /// pcs.rs:223-265: per matrix one library call = coset iDFT + one batched MSM over all columns, the columns
    /// sharded over the GPUs inside the library.
    fn commit(
        &self,
        evaluations: impl IntoIterator<Item = (Self::Domain, RowMajorMatrix<Fr>)>,
    ) -> (Self::Commitment, Self::ProverData) {
        let mut matrices = Vec::new();
        let mut prover = Vec::new();
        for (domain, evals) in evaluations {
            let (height, width) = (evals.height(), evals.width());
            assert_eq!(height, domain.size(), "evaluation height must match domain size");
            let shift = fr_limbs(&domain.shift());
            let mut xy = vec![0u64; 8 * width];
            let mut handle: sys::eon_handle = 0;
            let mut lde = None;
            let rc = match self.lde_hint {
                Some((added_bits, lde_shift)) if width > 0 => {
                    let log_size = domain.log_size() + added_bits;
                    let mut out = Fr::zero_vec((height << added_bits) * width);
                    let ls = fr_limbs(&lde_shift);
                    let rc = unsafe {
                        sys::eon_mctx_kzg_commit_lde(
                            ctx(),
                            fr_ptr(&evals.values),
                            domain.log_size() as u32,
                            width,
                            shift.as_ptr(),
                            xy.as_mut_ptr(),
                            &mut handle,
                            log_size as u32,
                            ls.as_ptr(),
                            fr_mut_ptr(&mut out),
                        )
                    };
                    let d = TwoAdicMultiplicativeCoset::new(lde_shift, log_size).expect("valid domain");
                    lde = Some((d, RowMajorMatrix::new(out, width)));
                    rc
                }
                _ => unsafe {
                    sys::eon_mctx_kzg_commit(
                        ctx(),
                        fr_ptr(&evals.values),
                        domain.log_size() as u32,
                        width,
                        shift.as_ptr(),
                        xy.as_mut_ptr(),
                        &mut handle,
                    )
                },
            };
            self.degree_guard(rc, height.saturating_sub(1));
            matrices.push(MatrixCommitment { columns: xy.chunks_exact(8).map(g1_from_wire).collect() });
            prover.push(GpuMatrixProverData { domain, evals, handle, width, lde: RefCell::new(lde) });
        }
        (KzgCommitment { matrices }, prover)
    }

    /// Override of the trait default (commit/src/pcs.rs:82-102): no `split_evals` copy on the host — chunk i is
    /// the pitched view "rows i, i + c, ..." of the uploaded matrix (commit/src/domain.rs:188-221) on the coset
    /// `shift * omega^i` (domain.rs:174-186) — and ONE batched MSM commits every column of every chunk.
    fn commit_quotient(
        &self,
        quotient_domain: Self::Domain,
        quotient_evaluations: RowMajorMatrix<Fr>,
        num_chunks: usize,
    ) -> (Self::Commitment, Self::ProverData) {
        let log_chunks = log2_strict_usize(num_chunks); // same panic as split_domains (domain.rs:175)
        let (height, width) = (quotient_evaluations.height(), quotient_evaluations.width());
        assert_eq!(height, quotient_domain.size(), "evaluation height must match domain size");
        let shift = fr_limbs(&quotient_domain.shift());
        let mut xy = vec![0u64; 8 * num_chunks * width];
        let mut handles = vec![0 as sys::eon_handle; num_chunks];
        let rc = unsafe {
            sys::eon_mctx_kzg_commit_quotient(
                ctx(),
                fr_ptr(&quotient_evaluations.values),
                quotient_domain.log_size() as u32,
                width,
                log_chunks as u32,
                shift.as_ptr(),
                xy.as_mut_ptr(),
                handles.as_mut_ptr(),
            )
        };
        self.degree_guard(rc, (height / num_chunks).saturating_sub(1));
        let sub_domains = quotient_domain.split_domains(num_chunks);
        // MatrixProverData.evals of chunk i (only read by the same-domain fast path): the strided rows
        let sub_evals = quotient_domain.split_evals(num_chunks, quotient_evaluations);
        let mut matrices = Vec::with_capacity(num_chunks);
        let mut prover = Vec::with_capacity(num_chunks);
        for (i, (domain, evals)) in sub_domains.into_iter().zip(sub_evals).enumerate() {
            let cols = &xy[8 * i * width..8 * (i + 1) * width];
            matrices.push(MatrixCommitment { columns: cols.chunks_exact(8).map(g1_from_wire).collect() });
            prover.push(GpuMatrixProverData { domain, evals, handle: handles[i], width, lde: RefCell::new(None) });
        }
        (KzgCommitment { matrices }, prover)
    }

    /// pcs.rs:267-287.  The reference evaluates by Horner, O(|domain| * h) per column; here: zero-pad (or fold,
    /// for a coset smaller than h) + one coset NTT from the device-resident coefficients — the same values.
    fn get_evaluations_on_domain<'a>(
        &self,
        prover_data: &'a Self::ProverData,
        idx: usize,
        domain: Self::Domain,
    ) -> Self::EvaluationsOnDomain<'a> {
        let m = &prover_data[idx];
        if m.domain.shift() == domain.shift() && m.domain.size() == domain.size() {
            return m.evals.clone();
        }
        if let Some((d, mat)) = m.lde.borrow().as_ref() {
            if d.shift() == domain.shift() && d.size() == domain.size() {
                return mat.clone();
            }
        }
        let mut out = Fr::zero_vec(domain.size() * m.width);
        let shift = fr_limbs(&domain.shift());
        check(unsafe {
            sys::eon_mctx_kzg_evals_on_coset(ctx(), m.handle, domain.log_size() as u32, shift.as_ptr(), fr_mut_ptr(&mut out))
        });
        RowMajorMatrix::new(out, m.width)
    }

    /// pcs.rs:289-335.  Every (round, matrix, point, column) of the call goes through ONE
    /// `eon_mctx_kzg_open_batch`: all quotients side by side, one batched MSM for all witnesses; the flat results
    /// are cut back into `[round][matrix][point][column]` (the layout the header documents).
    fn open(
        &self,
        commitment_data_with_opening_points: Vec<(&Self::ProverData, Vec<Vec<Fr>>)>,
        _fiat_shamir_challenger: &mut Challenger,
    ) -> (OpenedValues<Fr>, Self::Proof) {
        let mut handles: Vec<sys::eon_handle> = Vec::new();
        let mut npoints: Vec<usize> = Vec::new();
        let mut widths: Vec<usize> = Vec::new();
        let mut points: Vec<u64> = Vec::new();
        for (prover_data, points_per_matrix) in &commitment_data_with_opening_points {
            assert_eq!(prover_data.len(), points_per_matrix.len());
            for (m, pts) in prover_data.iter().zip(points_per_matrix) {
                handles.push(m.handle);
                npoints.push(pts.len());
                widths.push(m.width);
                for z in pts {
                    points.extend_from_slice(&fr_limbs(z));
                }
            }
        }
        let total: usize = npoints.iter().zip(&widths).map(|(n, w)| n * w).sum();
        let mut values = Fr::zero_vec(total.max(1));
        let mut wits = vec![0u64; 8 * total.max(1)];
        if !handles.is_empty() {
            let rc = unsafe {
                sys::eon_mctx_kzg_open_batch(
                    ctx(),
                    handles.len(),
                    handles.as_ptr(),
                    npoints.as_ptr(),
                    points.as_ptr(),
                    fr_mut_ptr(&mut values),
                    wits.as_mut_ptr(),
                )
            };
            if rc == sys::EON_ERR_SRS_TOO_SHORT {
                panic!("commit_column(&quotient).unwrap(): {}", last_error()); // pcs.rs:316
            }
            check(rc);
        }
        let mut opened_values = Vec::new();
        let mut rounds = Vec::new();
        let (mut k, mut i) = (0usize, 0usize);
        for (prover_data, _) in &commitment_data_with_opening_points {
            let mut matrix_values = Vec::new();
            let mut matrix_proofs = Vec::new();
            for _ in prover_data.iter() {
                let (w, c) = (widths[i], npoints[i]);
                let mut values_for_matrix = Vec::with_capacity(c);
                let mut proofs_for_matrix = Vec::with_capacity(c);
                for p in 0..c {
                    let lo = k + p * w;
                    values_for_matrix.push(values[lo..lo + w].to_vec());
                    proofs_for_matrix.push(PointProof {
                        witnesses: wits[8 * lo..8 * (lo + w)].chunks_exact(8).map(g1_from_wire).collect(),
                    });
                }
                matrix_values.push(values_for_matrix);
                matrix_proofs.push(MatrixProof { points: proofs_for_matrix });
                k += c * w;
                i += 1;
            }
            opened_values.push(matrix_values);
            rounds.push(matrix_proofs);
        }
        (opened_values, KzgProof { rounds })
    }

    /// pcs.rs:337-401: pairings, CPU side, unchanged.
    fn verify(
        &self,
        commitments_with_opening_points: Vec<(Self::Commitment, Vec<(Self::Domain, Vec<(Fr, Vec<Fr>)>)>)>,
        proof: &Self::Proof,
        fiat_shamir_challenger: &mut Challenger,
    ) -> Result<(), Self::Error> {
        <KzgPcs as Pcs<Fr, Challenger>>::verify(&self.cpu, commitments_with_opening_points, proof, fiat_shamir_challenger)
    }
}

/// `G1::multi_exp` over the resident SRS (what `commit_column` does, kzg/src/util.rs:37-40), one column.
pub fn commit_column(coeffs: &[Fr]) -> G1 {
    let mut xy = [0u64; 8];
    check(unsafe { sys::eon_mctx_msm_srs(ctx(), fr_ptr(coeffs), coeffs.len(), 1, 1, xy.as_mut_ptr()) });
    g1_from_wire(&xy)
}

/// `G1::multi_exp(points, scalars)` (bn254/src/curve.rs:158-180) over explicit bases.
pub fn multi_exp(points: &[G1], scalars: &[Fr]) -> G1 {
    assert_eq!(points.len(), scalars.len(), "points and scalars must have the same length");
    let mut pts = Vec::with_capacity(8 * points.len());
    for p in points {
        let (x, y) = p.to_affine_montgomery_limbs();
        pts.extend_from_slice(&x);
        pts.extend_from_slice(&y);
    }
    let mut xy = [0u64; 8];
    check(unsafe { sys::eon_mctx_msm_points(ctx(), pts.as_ptr(), fr_ptr(scalars), points.len(), xy.as_mut_ptr()) });
    g1_from_wire(&xy)
}
