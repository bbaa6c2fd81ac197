/*
 * eon_kzg.h — C ABI of the B200-native BN254 KZG hot path (libeon_kzg.so, sm_100a).
 *
 * Drop-in boundary for the reference's plugin surface (Lolazyx/plonky3-eon):
 *   p3-dft   TwoAdicSubgroupDft<Fr>      dft/src/traits.rs:27-507
 *   p3-bn254 G1::multi_exp               bn254/src/curve.rs:158-180
 *   p3-kzg   KzgPcs (Pcs<Fr, _>)          kzg/src/pcs.rs:143-335
 *            commit_column / quotient     kzg/src/util.rs:37-40,100-111
 *            init_srs_unsafe              kzg/src/params.rs:123-139
 *
 * Wire formats (identical to the reference's in-memory layout, so Rust can pass
 * `values.as_ptr() as *const u64` directly):
 *   Fr      4 x u64 little-endian, Montgomery form (a * 2^256 mod r), canonical (< r)
 *           — `Fr { value: [u64; 4] }`, bn254/src/field.rs:96-105.
 *   matrix  row-major, `height x width` Fr, row = width * 32 contiguous bytes
 *           — RowMajorMatrix<Fr>, matrix/src/dense.rs:24-37.
 *   G1      affine, 8 x u64: [x: 4 x u64][y: 4 x u64], Montgomery Fq (R = 2^256),
 *           identity = all zero (halo2curves G1Affine::identity()).
 *
 * Preconditions the reference's types guarantee and this ABI does NOT re-check on the hot path: every Fr handed
 * over (matrix entries, scalars) is canonical (< r; the reference's Fr is always reduced, field.rs:96-105) and
 * every affine point of eon_srs_load_affine / eon_msm_points lies on the curve (halo2curves' G1 cannot hold
 * anything else).  Shifts and opening points ARE checked (EON_ERR_BAD_ARG); compressed points are validated by
 * eon_g1_decompress / eon_srs_load_compressed (EON_ERR_BAD_POINT).
 *
 * All functions return EON_OK (0) or a negative error code; none aborts or unwinds.
 * `eon_last_error(ctx)` gives a human-readable message for the last failure on that ctx.
 * There is NO CPU fallback: every entry point fails with EON_ERR_CUDA if no sm_100 device
 * is usable.
 *
 * Pointer arguments named `h_*` are host pointers; `d_*` are device pointers on the
 * context's device (from eon_dev_alloc or any CUDA allocation, e.g. a torch tensor).
 */
#ifndef EON_KZG_H
#define EON_KZG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eon_ctx eon_ctx;
typedef uint64_t eon_handle; /* opaque id of device-resident prover data (0 = invalid) */

enum {
  EON_OK = 0,
  EON_ERR_BAD_ARG = -1,      /* null pointer, bad shape (e.g. height not a power of two:
                                the reference panics in log2_strict_usize, util/src/lib.rs:39) */
  EON_ERR_SRS_TOO_SHORT = -2,/* KzgError::DegreeTooLarge, kzg/src/params.rs:164-173 */
  EON_ERR_CUDA = -3,         /* CUDA runtime failure / no device */
  EON_ERR_OOM = -4,          /* device allocation failed */
  EON_ERR_BAD_HANDLE = -5,
  EON_ERR_TWO_ADICITY = -6,  /* transform size exceeds 2^28 (Fr::TWO_ADICITY, field.rs:564) */
  EON_ERR_BAD_POINT = -7     /* a compressed G1 encoding is not a curve point: serde's
                                "Invalid G1 point", bn254/src/curve.rs:94-96 */
};

/* ---- context ------------------------------------------------------------------------- */
/* One context per GPU (several GPUs behind one handle: eon_mctx below).  `stream` is a cudaStream_t (NULL = the legacy default stream); every
 * kernel of the context is launched on it, so callers can bracket calls with their own
 * events (e.g. torch.cuda.current_stream().cuda_stream).
 * Replaces the implicit "process-global Radix2Dit / KzgPcs state" of the reference
 * (twiddle caches dft/src/radix_2_dit.rs:33-58; SRS kzg/src/params.rs:57-77). */
int eon_ctx_create(int device, void* stream, eon_ctx** out);
void eon_ctx_destroy(eon_ctx* ctx);
const char* eon_last_error(const eon_ctx* ctx);
int eon_ctx_sync(eon_ctx* ctx);
/* number of CUDA kernels this context has launched so far (for bench.py "gpu_launches") */
uint64_t eon_ctx_launch_count(const eon_ctx* ctx);
/* library build tag, e.g. "eon_kzg sm_100a" */
const char* eon_version(void);

/* ---- device memory helpers (plain cudaMalloc / cudaMemcpyAsync on the ctx stream) -------- */
int eon_dev_alloc(eon_ctx* ctx, size_t bytes, void** d_out);
int eon_dev_free(eon_ctx* ctx, void* d_ptr);
int eon_h2d(eon_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
/* page-locked host memory (cudaHostAlloc) for matrices handed to / returned by the host-buffer entry
 * points: pageable buffers work too but copy at about half the rate.  A shim keeps a pool of these. */
int eon_host_alloc(size_t bytes, void** h_out);
int eon_host_free(void* h_ptr);
int eon_d2h(eon_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* ---- TwoAdicSubgroupDft<Fr> (dft/src/traits.rs) --------------------------------------------
 * All matrices natural row order in and out (what the trait's logical view is).
 * `shift` = 4 x u64 Montgomery Fr, must be non-zero.  height = 1 << log_h.                  */

/* dft_batch (traits.rs:61; semantics dft/src/naive.rs:15-31).  d_out may not alias d_in. */
int eon_dft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width);
/* coset_dft_batch (traits.rs:83-91) */
int eon_coset_dft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                            const uint64_t shift[4]);
/* idft_batch (traits.rs:111-122) */
int eon_idft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width);
/* coset_idft_batch (traits.rs:144-153) */
int eon_coset_idft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                             const uint64_t shift[4]);
/* coset_lde_batch (traits.rs:226-249): d_out has (1 << (log_h + added_bits)) rows. */
int eon_coset_lde_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                            unsigned added_bits, const uint64_t shift[4]);

/* Host-buffer forms (what a Rust `impl TwoAdicSubgroupDft<Fr>` binds): copy in, run, copy out. */
int eon_dft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width);
int eon_coset_dft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                        const uint64_t shift[4]);
int eon_idft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width);
int eon_coset_idft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                         const uint64_t shift[4]);
int eon_coset_lde_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                        unsigned added_bits, const uint64_t shift[4]);

/* coset_lde_batch on the columns [0, width) of a WIDER host matrix: ld_in / ld_out = row pitch of the host
 * matrices in Fr elements (0 = dense).  A caller that shards the columns of one RowMajorMatrix over several
 * contexts (one per GPU) passes (base + c0 * 4, ld = full width) to each: no host-side de-interleave. */
int eon_coset_lde_batch_ld(eon_ctx* ctx, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out,
                           unsigned log_h, size_t width, unsigned added_bits, const uint64_t shift[4]);

/* ---- SRS (kzg/src/params.rs) ------------------------------------------------------------- */
/* Upload g1_powers as n affine points (8 x u64 each).  Replaces the per-call `to_affine` of
 * bn254/src/curve.rs:170: points are normalised once by the caller and stay resident. */
int eon_srs_load_affine(eon_ctx* ctx, const uint64_t* h_xy, size_t n);
/* ---- compressed G1 points (G1::to_bytes bn254/src/curve.rs:136-139; Serialize / Deserialize for G1
 * :84-98, i.e. how StructuredReferenceString, KzgCommitment and KzgProof are (de)serialised and how
 * CanObserve<KzgCommitment> feeds commitments to the challenger, kzg/src/pcs.rs:409-438) ------------
 * 32 bytes per point = halo2curves' GroupEncoding of bn256::G1Affine.  halo2curves is not vendored in
 * the reference; the layout is restated from its published definition (csrc/codec.cu):
 *   EON_G1_ENC_HALO2   x little-endian canonical; byte 31 bit 6 = lowest bit of canonical y; byte 31
 *                      bit 7 = identity (every other bit zero).  halo2curves 0.4 and later, incl. "0.9".
 *   EON_G1_ENC_LEGACY  sign in byte 31 bit 7; identity = 32 zero bytes (halo2curves 0.3 and earlier). */
enum { EON_G1_ENC_HALO2 = 0, EON_G1_ENC_LEGACY = 1 };
/* n affine wire points (host) -> 32 n bytes (host), compressed on the device. */
int eon_g1_compress(eon_ctx* ctx, const uint64_t* h_xy, size_t n, uint8_t* h_out, int enc);
/* 32 n bytes (host) -> n affine wire points (host): batched square roots on the device.  Fails with
 * EON_ERR_BAD_POINT if an encoding is not a curve point; *bad_index (may be NULL) = index of the first
 * such encoding, or SIZE_MAX. */
int eon_g1_decompress(eon_ctx* ctx, const uint8_t* h_in, size_t n, uint64_t* h_xy, int enc, size_t* bad_index);
/* Deserialised-SRS ingest: g1_powers as n compressed points, decompressed straight into the resident
 * affine SRS (and its window tables); on EON_ERR_BAD_POINT the previous SRS is gone and none is loaded. */
int eon_srs_load_compressed(eon_ctx* ctx, const uint8_t* h_in, size_t n, int enc, size_t* bad_index);

/* init_srs_unsafe (params.rs:123-139), G1 part: g1_powers[i] = alpha^i * G for i < n, generated
 * on the device.  alpha: Montgomery Fr. */
int eon_srs_generate_unsafe(eon_ctx* ctx, const uint64_t alpha[4], size_t n);
/* number of G1 powers resident (max_degree + 1), 0 if none */
size_t eon_srs_size(const eon_ctx* ctx);
/* Window tables for MSMs over the resident SRS: tab[t][i] = 2^(c t) * g1_powers[i] for
 * t < ceil(255 / c), c = window_bits in [8, 20]; 0 drops the tables (plain per-window buckets).
 * Costs ceil(255/c) x the SRS memory; built automatically by the two loaders above for n >= 2^14
 * (c chosen by a cost model: 17 for 2^20 points, 20 for 2^24), as long as they fit in 64 GiB.  eon_srs_window_bits returns the c in use (0 = no tables). */
int eon_srs_set_window_tables(eon_ctx* ctx, unsigned window_bits);
unsigned eon_srs_window_bits(const eon_ctx* ctx);
/* MSM bucket accumulation strategy: `rounds` batched-affine pairwise rounds (one shared Fq inversion
 * per round, ~7.6 instead of 10 Fq products per addition) before the XYZZ finisher; 0 = XYZZ only,
 * -1 = automatic (3 rounds once buckets average >= 64 entries).  Results are identical either way. */
int eon_msm_set_rounds(eon_ctx* ctx, int rounds);
/* how entries are sorted by bucket: 0 = one pass of global atomics, 1 = coarse + fine coalesced
 * passes (csrc/msm_sort.cu; with the slice schedule the sort also writes round 0's pair records), 2 = the same
 * passes with entries and pair records built separately (the unfused flow), -1 = automatic.  Results are
 * identical either way. */
int eon_msm_set_sort_mode(eon_ctx* ctx, int mode);
/* order in which round 0 of the pairwise rounds walks its pairs: 1 = by 64 MiB slice of the base table
 * (the gathers of a multi-column commit then hit the L2 instead of DRAM; csrc/msm_tree.cu), 0 = in slot
 * order, -1 = automatic (slices when the table exceeds the L2 and every base is used by several columns).
 * Results are identical either way. */
int eon_msm_set_slice_schedule(eon_ctx* ctx, int mode);
/* An MSM over few columns (2..4: the shard of one GPU when 16 trace columns are split over 4-8 GPUs, the two chunk
 * columns of commit_quotient) spends a third of its time in phases that leave the GPU idle -- the single-warp
 * inversion trees of the batched-affine rounds, the bucket reduction.  1 / -1 (automatic): such an MSM runs as two
 * half-batches on two streams, so that the idle phases of one half overlap the wide kernels of the other;
 * 0 = one batch.  Results are identical either way. */
int eon_msm_set_split(eon_ctx* ctx, int mode);
/* rounds the most recent MSM on this context actually used */
unsigned eon_msm_rounds_used(const eon_ctx* ctx);
/* window bits c of the most recent MSM on this context (whole-SRS tables, range tables or the plain per-window c) */
unsigned eon_msm_window_bits_used(const eon_ctx* ctx);
/* copy SRS points [first, first + n) back to the host as affine wire points */
int eon_srs_read(eon_ctx* ctx, size_t first, size_t n, uint64_t* h_xy);

/* ---- G1::multi_exp (bn254/src/curve.rs:158-180) -------------------------------------------
 * out[c] = sum_{i<n} scalar[i][c] * bases[i]   for c in [0, ncols)
 * scalars: row-major matrix with `ld` Fr per row (ld >= ncols), Montgomery form;
 * bases: the resident SRS (points [0, n)) or an explicit affine array.
 * n == 0 -> identity (curve.rs:165-167).  Results: ncols affine wire points (host). */
int eon_msm_srs_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy);
int eon_msm_srs(eon_ctx* ctx, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy);
int eon_msm_points(eon_ctx* ctx, const uint64_t* h_points_xy, const uint64_t* h_scalars, size_t n, uint64_t* h_out_xy);
/* MSM over the SRS slice [first, first + n) — the index-range shard of a multi-GPU MSM. */
int eon_msm_srs_range_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t first, size_t n, size_t ncols, size_t ld,
                          uint64_t* h_out_xy);
/* The same shard with its ncols partial sums left ON THE DEVICE (d_out_xy: ncols affine wire points), queued on the
 * context's stream without a host synchronisation: the shards' sums are then gathered over NVLink (peer copy or
 * ncclAllGather on the same stream) and added by eon_g1_sum_cols_dev -- no host hop (EC addition is not an NCCL
 * reduction op, so the "partial-sum reduction" of BASELINE configs[4] is all_gather + one add kernel). */
int eon_msm_srs_range_partial_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t first, size_t n, size_t ncols,
                                  size_t ld, uint64_t* d_out_xy);
/* h_out_xy[c] = sum_p d_parts_xy[p * ncols + c]: nparts x ncols affine wire points on the device -> ncols on the host */
int eon_g1_sum_cols_dev(eon_ctx* ctx, const uint64_t* d_parts_xy, size_t nparts, size_t ncols, uint64_t* h_out_xy);
/* Window tables over the SRS index range [first, first + n) only, beside the whole-SRS tables: the shard one GPU
 * owns in an index-range sharded MSM gets a window sized for ITS length (2^21 points of a 2^24-point SRS: c = 17-18
 * instead of 20, an eighth of the buckets to reduce).  window_bits 0 = cost model; n = 0 drops them.  MSMs whose
 * bases lie inside the range use these tables. */
int eon_srs_set_range_tables(eon_ctx* ctx, size_t first, size_t n, unsigned window_bits);
/* out = sum of n affine points (combining per-GPU partial sums). */
int eon_g1_sum(eon_ctx* ctx, const uint64_t* h_points_xy, size_t n, uint64_t* h_out_xy);

/* ---- KzgPcs (kzg/src/pcs.rs) ------------------------------------------------------------- */
/* commit (pcs.rs:223-265) for ONE matrix: coefficients = coset_idft_batch(evals, shift);
 * commitment[c] = MSM(srs[..h], coefficients[:, c]).  The coefficient matrix stays on the
 * device behind `*out_handle` (MatrixProverData.coeffs, pcs.rs:252-256).
 * Fails with EON_ERR_SRS_TOO_SHORT if srs_size < h (pcs.rs:238-240). */
int eon_kzg_commit(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                   uint64_t* h_commit_xy, eon_handle* out_handle);
int eon_kzg_commit_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                       uint64_t* h_commit_xy, eon_handle* out_handle);
/* commit + the evaluations the prover asks for next, in one call: exactly eon_kzg_commit followed by
 * eon_kzg_evals_on_coset(handle, lde_log_size, lde_shift, h_lde_out) (pcs.rs:223-265 then :267-287, as
 * called back to back by eon-uni-stark/src/prover.rs:186-187,307-322), but the LDE of every column
 * group crosses PCIe while that group's MSM runs, so the 2^(lde_log_size) x width download is hidden,
 * and the LDE transform itself runs on a second stream beside the MSM (whose sort and base-gather
 * phases leave the integer pipe idle).  _dev: device buffers in and out.
 * A Pcs shim built with an "LDE hint" (quotient-domain size and shift, both known before the trace is
 * committed) calls this from commit() and hands the matrix out from get_evaluations_on_domain(). */
int eon_kzg_commit_lde_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_h, size_t width,
                           const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle,
                           unsigned lde_log_size, const uint64_t lde_shift[4], uint64_t* d_lde_out);
int eon_kzg_commit_lde(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                       uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                       const uint64_t lde_shift[4], uint64_t* h_lde_out);
/* The two entry points above on the columns [0, width) of a wider host matrix (row pitch ld_in; ld_out for the LDE
 * result), see eon_coset_lde_batch_ld. */
int eon_kzg_commit_ld(eon_ctx* ctx, const uint64_t* h_evals, size_t ld_in, unsigned log_h, size_t width,
                      const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle);
int eon_kzg_commit_lde_ld(eon_ctx* ctx, const uint64_t* h_evals, size_t ld_in, unsigned log_h, size_t width,
                          const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                          const uint64_t lde_shift[4], uint64_t* h_lde_out, size_t ld_out);
/* Pcs::commit_quotient (trait default, commit/src/pcs.rs:82-102) in one call.  `evals`: the quotient
 * evaluations on shift*<omega_{2^log_size}>, natural order, (1 << log_size) x width.  Chunk i (of
 * 2^log_chunks) is rows i, i + 2^log_chunks, ... (split_evals, commit/src/domain.rs:188-221) on the coset
 * shift*omega^i of size 2^(log_size - log_chunks) (split_domains, domain.rs:174-186); each chunk is committed
 * exactly as eon_kzg_commit would commit it, but nothing is de-interleaved on the host (a chunk is a pitched
 * view of the same device buffer) and ONE batched MSM covers every column of every chunk.
 *   h_commit_xy[(i*width + c)*8]  = commitment of column c of chunk i
 *   out_handles[i]               = prover data (coefficients) of chunk i, as from eon_kzg_commit
 * Fails with EON_ERR_SRS_TOO_SHORT if srs_size < 2^(log_size - log_chunks). */
int eon_kzg_commit_quotient(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_size, size_t width,
                            unsigned log_chunks, const uint64_t shift[4], uint64_t* h_commit_xy,
                            eon_handle* out_handles);
int eon_kzg_commit_quotient_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_size, size_t width,
                                unsigned log_chunks, const uint64_t shift[4], uint64_t* h_commit_xy,
                                eon_handle* out_handles);
/* KzgMmcs::commit (kzg/src/mmcs.rs:155-190) for ONE matrix: the columns are taken as polynomials
 * in COEFFICIENT form (no iDFT); `rows` may be any height (not only powers of two).
 * commitment[c] = commit_column(matrix[:, c]) (mmcs.rs:155-165).  The matrix is kept on the device
 * behind `*out_handle`, so eon_kzg_open on that handle is KzgMmcs::open_batch (mmcs.rs:192-237).
 * Fails with EON_ERR_SRS_TOO_SHORT if rows - 1 > max_degree (mmcs.rs:177-179). */
int eon_kzg_commit_coeffs(eon_ctx* ctx, const uint64_t* h_coeffs, size_t rows, size_t width, uint64_t* h_commit_xy,
                          eon_handle* out_handle);
int eon_kzg_commit_coeffs_dev(eon_ctx* ctx, const uint64_t* d_coeffs, size_t rows, size_t width,
                              uint64_t* h_commit_xy, eon_handle* out_handle);
/* copy the retained coefficient matrix (natural order, h x width) to the host */
int eon_kzg_read_coeffs(eon_ctx* ctx, eon_handle h, uint64_t* h_out);
/* get_evaluations_on_domain (pcs.rs:267-287) on the coset shift*<omega_{2^log_size}>,
 * natural order, (1 << log_size) x width.  Computed as zero-pad + coset NTT, which is
 * bit-identical to the reference's Horner evaluation; a coset smaller than h (the Horner loop accepts any)
 * is served by first reducing the coefficients mod X^(2^log_size) - shift^(2^log_size). */
int eon_kzg_evals_on_coset(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out);
int eon_kzg_evals_on_coset_dev(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* d_out);
/* ... into the columns [0, width) of a wider host matrix with row pitch ld_out (0 = dense) */
int eon_kzg_evals_on_coset_ld(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out,
                              size_t ld_out);
/* open (pcs.rs:289-335) of one matrix at `npoints` points: for point p and column c
 *   h_values[p*width + c]       = f_c(z_p)                         (Fr)
 *   h_witness_xy[(p*width+c)*8] = commit((f_c - f_c(z_p))/(X - z_p)) (G1 affine)
 * (quotient_and_eval, util.rs:100-111, then commit_column). */
int eon_kzg_open(eon_ctx* ctx, eon_handle h, const uint64_t* h_points, size_t npoints, uint64_t* h_values,
                 uint64_t* h_witness_xy);
/* open (pcs.rs:289-335) of `nmat` committed matrices in one call: matrix m (handles[m]) is opened at its
 * npoints[m] points, taken in order from h_points (4 limbs each, concatenated over the matrices).  Outputs are
 * laid out [matrix][point][column] like the nested OpenedValues / KzgProof of the reference:
 *   h_values[k*4], h_witness_xy[k*8]  with  k = sum_{m' < m} npoints[m']*width[m'] + p*width[m] + c
 * Every quotient is written beside the others and ONE batched MSM yields all witnesses (heights may differ;
 * results are identical to eon_kzg_open per matrix). */
int eon_kzg_open_batch(eon_ctx* ctx, size_t nmat, const eon_handle* handles, const size_t* npoints,
                       const uint64_t* h_points, uint64_t* h_values, uint64_t* h_witness_xy);
int eon_handle_dims(eon_ctx* ctx, eon_handle h, unsigned* log_h, size_t* width);
int eon_handle_free(eon_ctx* ctx, eon_handle h);

/* quotient_and_eval (util.rs:100-111) on a device coefficient matrix (h x width, natural):
 * d_quot gets an h x width matrix whose rows [0, h-1) are the quotient coefficients and
 * whose last row is zero; h_values gets the width evaluations f_c(z). */
int eon_quotient_and_eval_dev(eon_ctx* ctx, const uint64_t* d_coeffs, size_t h, size_t width, const uint64_t z[4],
                              uint64_t* d_quot, uint64_t* h_values);

/* ---- multi-device context: one process, several GPUs, the same host-matrix calls ---------------------------------
 * The reference prover is ONE process that hands whole matrices to pcs.commit / get_evaluations_on_domain / open
 * (eon-uni-stark/src/prover.rs:186-187,307-322,371-372,424-442); KzgPcs then walks the columns one by one
 * (kzg/src/pcs.rs:244-249,311-318).  An eon_mctx takes the same whole host matrices and splits the COLUMNS over its
 * devices (device g: a contiguous column range, copied straight out of / into the caller's row-major buffers with
 * strided copies; the SRS is replicated), so a Rust `Pcs` / `TwoAdicSubgroupDft` shim reaches every GPU of the box
 * without knowing about them.  Results are byte-identical to the single-device entry points of the same name.
 * A single MSM with fewer columns than devices is sharded by point index instead (per-device range tables, partial
 * sums pushed to the first device as NVLink peer copies, one add kernel there).
 * `devices`: CUDA ordinals; the same ordinal may appear more than once (several shard contexts on one GPU: how the
 * sharding logic is tested on a single-GPU box).  Handles returned by an eon_mctx are only valid with eon_mctx_*.
 * The per-device contexts are reachable through eon_mctx_ctx for tuning / timing calls (eon_msm_set_*,
 * eon_last_phase_ms, ...); do not run work on them while an eon_mctx_* call is in flight. */
typedef struct eon_mctx eon_mctx;
int eon_mctx_create(const int* devices, int n, eon_mctx** out);
void eon_mctx_destroy(eon_mctx* m);
const char* eon_mctx_last_error(const eon_mctx* m);
int eon_mctx_device_count(const eon_mctx* m);
eon_ctx* eon_mctx_ctx(eon_mctx* m, int i);
uint64_t eon_mctx_launch_count(const eon_mctx* m);
/* SRS, replicated on every device (init_srs_unsafe / deserialised g1_powers, kzg/src/params.rs:57-139) */
int eon_mctx_srs_generate_unsafe(eon_mctx* m, const uint64_t alpha[4], size_t n);
int eon_mctx_srs_load_affine(eon_mctx* m, const uint64_t* h_xy, size_t n);
int eon_mctx_srs_load_compressed(eon_mctx* m, const uint8_t* h_in, size_t n, int enc, size_t* bad_index);
size_t eon_mctx_srs_size(const eon_mctx* m);
int eon_mctx_srs_read(eon_mctx* m, size_t first, size_t n, uint64_t* h_xy);
/* TwoAdicSubgroupDft<Fr> (dft/src/traits.rs:61,83-91,111-122,144-153,226-249), whole host matrices */
int eon_mctx_dft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width);
int eon_mctx_coset_dft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                             const uint64_t shift[4]);
int eon_mctx_idft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width);
int eon_mctx_coset_idft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                              const uint64_t shift[4]);
int eon_mctx_coset_lde_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                             unsigned added_bits, const uint64_t shift[4]);
/* KzgPcs (kzg/src/pcs.rs:223-335) and KzgMmcs::commit (kzg/src/mmcs.rs:155-190), arguments as the eon_kzg_* forms */
int eon_mctx_kzg_commit(eon_mctx* m, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                        uint64_t* h_commit_xy, eon_handle* out_handle);
int eon_mctx_kzg_commit_lde(eon_mctx* m, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                            uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                            const uint64_t lde_shift[4], uint64_t* h_lde_out);
int eon_mctx_kzg_commit_coeffs(eon_mctx* m, const uint64_t* h_coeffs, size_t rows, size_t width, uint64_t* h_commit_xy,
                               eon_handle* out_handle);
int eon_mctx_kzg_commit_quotient(eon_mctx* m, const uint64_t* h_evals, unsigned log_size, size_t width,
                                 unsigned log_chunks, const uint64_t shift[4], uint64_t* h_commit_xy,
                                 eon_handle* out_handles);
int eon_mctx_kzg_evals_on_coset(eon_mctx* m, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out);
int eon_mctx_kzg_open_batch(eon_mctx* m, size_t nmat, const eon_handle* handles, const size_t* npoints,
                            const uint64_t* h_points, uint64_t* h_values, uint64_t* h_witness_xy);
int eon_mctx_kzg_read_coeffs(eon_mctx* m, eon_handle h, uint64_t* h_out);
int eon_mctx_handle_dims(eon_mctx* m, eon_handle h, unsigned* log_h, size_t* width);
int eon_mctx_handle_free(eon_mctx* m, eon_handle h);
/* G1::multi_exp (bn254/src/curve.rs:158-180): over the resident SRS (columns sharded when ncols >= devices, else the
 * points by index range) and over explicit bases (index range) */
int eon_mctx_msm_srs(eon_mctx* m, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy);
int eon_mctx_msm_points(eon_mctx* m, const uint64_t* h_points_xy, const uint64_t* h_scalars, size_t n,
                        uint64_t* h_out_xy);

/* ---- measurement helpers ------------------------------------------------------------------- */
/* Dependency-free integer-multiply microbenchmark: launches `iters` rounds on all SMs and
 * returns the achieved 32-bit multiply-add rate in 1e12 ops/s (the IMAD roofline
 * denominator; not in MEASURED_PEAKS.json).  kind: 0 = mad.lo.u32, 1 = mad.hi.u32,
 * 2 = mad.wide.u32 (counted as 2 ops). */
int eon_bench_imad_peak(eon_ctx* ctx, int kind, double* out_tops);
/* Montgomery-product throughput (independent chains), in 1e9 modmul/s.  field: 0 Fr, 1 Fq. */
int eon_bench_modmul(eon_ctx* ctx, int field, double* out_gmuls);
/* The same for one formulation of the product.  variant: 0 = the product the library uses, 1 = word-serial
 * (CIOS) Montgomery product, 2 = split form (one Karatsuba level + separate reduction), 3 = dedicated square
 * (each square followed by one modular add).  All variants return identical limbs.
 * 4 = fixed-operand (Shoup) product with a precomputed quotient operand, lazily reduced (csrc/fp_shoup.cuh: the
 * multiplier of the NTT butterflies, 214 instead of 272 IMAD), 5 = the word-serial product without its final
 * correction (what the NTT passes run today), for comparison with 4. */
int eon_bench_modmul_variant(eon_ctx* ctx, int field, int variant, double* out_gmuls);
/* Which multiplier the NTT butterflies use in this process: 1 = fixed-operand (Shoup) product on (plain,
 * quotient) twiddle pairs, 214 IMAD per butterfly (the default); 0 = word-serial Montgomery product on
 * Montgomery-form twiddles, 272 IMAD (EON_NTT_SHOUP=0).  Results are identical either way. */
int eon_ntt_twiddle_form(void);
/* per-phase device time (ms, CUDA events on the ctx stream) summed over every call since the last
 * eon_phase_reset; names via eon_phase_name (phases 0 .. eon_phase_count() - 1).  The msm_tree_* and
 * msm_finish phases are sub-intervals of msm_accumulate. */
int eon_last_phase_ms(eon_ctx* ctx, int phase, float* out_ms);
int eon_phase_reset(eon_ctx* ctx);
int eon_phase_count(void);
/* strided host<->device copy rate (rows x width_bytes out of a pinned host matrix with row pitch
 * pitch_bytes), ms per copy: sizes the column-group pipelining of the host-buffer entry points */
int eon_bench_copy2d(eon_ctx* ctx, void* h_ptr, size_t rows, size_t width_bytes, size_t pitch_bytes, int to_device,
                     float* out_ms);
const char* eon_phase_name(int phase);

#ifdef __cplusplus
}
#endif
#endif /* EON_KZG_H */
