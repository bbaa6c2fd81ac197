// eon_kzg.hpp — C++ host-side mirror of the reference's plugin surface over the C ABI (eon_kzg.h).
//
// The reference's host language is Rust (no cargo/rustc in this image), so this header is the compiled-language
// binding that can actually be built and run here: the same types, method names, argument nesting and failure
// behaviour as the Rust traits it stands in for, written the way the Rust shim of INTEGRATION.md would be.
//
//   p3_matrix::dense::RowMajorMatrix<Fr>          matrix/src/dense.rs:24-37      -> RowMajorMatrix
//   p3_field::coset::TwoAdicMultiplicativeCoset   field/src/coset.rs:55-90,
//                                                  commit/src/domain.rs:144-221   -> TwoAdicMultiplicativeCoset
//   p3_dft::TwoAdicSubgroupDft<Fr>                dft/src/traits.rs:27-507       -> GpuDft
//   p3_bn254::G1 (multi_exp, to_bytes)            bn254/src/curve.rs:74-180      -> G1, multi_exp
//   p3_kzg::{KzgCommitment, KzgProof, KzgPcs}     kzg/src/pcs.rs:22-335          -> KzgCommitment, KzgProof, GpuKzgPcs
//   p3_kzg::KzgError::DegreeTooLarge              kzg/src/params.rs:164-211      -> DegreeTooLarge
//
// Rust panics on the prover path (height != domain size, SRS too short after unwrap(), non power-of-two heights,
// length mismatch in multi_exp) become C++ exceptions: std::logic_error for the assertions, DegreeTooLarge for
// the degree guard.  All compute happens in libeon_kzg.so; the only host arithmetic here is the handful of Fr
// products needed for coset shifts, done with the library's own fp.cuh compiled for the host.
#pragma once
#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../plonky3_eon_b200/csrc/fp.cuh"
#include "eon_kzg.h"

namespace p3eon {

using Fr = eon::Fr;  // 8 x u32 little-endian Montgomery limbs == the reference's [u64; 4]
static_assert(sizeof(Fr) == 32, "Fr wire size");

inline Fr fr_from_u64(uint64_t x) { return eon::fp_from_u64<eon::FrParams>(x); }       // Fr::from_u64
inline Fr fr_generator() { return fr_from_u64(5); }                                   // Fr::GENERATOR, field.rs:372
inline Fr fr_mul(const Fr& a, const Fr& b) { return eon::fp_mul(a, b); }
inline Fr fr_pow(Fr a, uint64_t e) { return eon::fp_pow_u64(a, e); }
// two_adic_generator(bits), field.rs:567-573
inline Fr fr_two_adic_generator(unsigned bits) {
  if (bits > 28) throw std::logic_error("two_adic_generator: bits > TWO_ADICITY");
  const uint32_t w28[8] = EON_FR_OMEGA28;
  Fr o;
  std::memcpy(o.v, w28, 32);
  for (unsigned i = bits; i < 28; i++) o = eon::fp_sqr(o);
  return o;
}
// p3_util::log2_strict_usize (util/src/lib.rs:39): panics on non powers of two
inline unsigned log2_strict(size_t n) {
  if (n == 0 || (n & (n - 1))) throw std::logic_error("Not a power of two: " + std::to_string(n));
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

// ---- errors -----------------------------------------------------------------------------------------------
struct EonError : std::runtime_error {
  int code;
  EonError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct DegreeTooLarge : EonError { using EonError::EonError; };   // KzgError::DegreeTooLarge
struct InvalidG1Point : EonError {                                 // serde "Invalid G1 point", curve.rs:95
  size_t index;
  InvalidG1Point(int c, const std::string& m, size_t i) : EonError(c, m), index(i) {}
};

// ---- context: one per GPU, shared by every Dft / Pcs clone (twiddle caches, SRS, scratch live in it) -------
class Context {
 public:
  explicit Context(int device = 0, void* stream = nullptr) {
    eon_ctx* c = nullptr;
    int rc = eon_ctx_create(device, stream, &c);
    if (rc != EON_OK) throw EonError(rc, "eon_ctx_create failed (no sm_100 device? there is no CPU fallback)");
    ctx_ = std::shared_ptr<eon_ctx>(c, [](eon_ctx* p) { eon_ctx_destroy(p); });
  }
  eon_ctx* raw() const { return ctx_.get(); }
  void check(int rc) const {
    if (rc == EON_OK) return;
    std::string msg = eon_last_error(ctx_.get());
    if (rc == EON_ERR_SRS_TOO_SHORT) throw DegreeTooLarge(rc, msg);
    throw EonError(rc, msg);
  }

 private:
  std::shared_ptr<eon_ctx> ctx_;
};

// ---- RowMajorMatrix<Fr> -------------------------------------------------------------------------------------
struct RowMajorMatrix {
  std::vector<Fr> values;
  size_t width_ = 0;
  RowMajorMatrix() = default;
  RowMajorMatrix(std::vector<Fr> v, size_t w) : values(std::move(v)), width_(w) {
    if (w && values.size() % w) throw std::logic_error("RowMajorMatrix: length not a multiple of the width");
  }
  size_t width() const { return width_; }
  size_t height() const { return width_ ? values.size() / width_ : 0; }
  const uint64_t* wire() const { return reinterpret_cast<const uint64_t*>(values.data()); }
  uint64_t* wire() { return reinterpret_cast<uint64_t*>(values.data()); }
  const Fr& at(size_t r, size_t c) const { return values[r * width_ + c]; }
};

// ---- TwoAdicMultiplicativeCoset ---------------------------------------------------------------------------------
struct TwoAdicMultiplicativeCoset {
  Fr shift;
  unsigned log_size;
  TwoAdicMultiplicativeCoset(const Fr& s, unsigned l) : shift(s), log_size(l) {
    if (s.is_zero() || l > 28) throw std::logic_error("invalid coset");  // ::new returns None
  }
  size_t size() const { return (size_t)1 << log_size; }
  Fr subgroup_generator() const { return fr_two_adic_generator(log_size); }
  Fr first_point() const { return shift; }
  Fr next_point(const Fr& x) const { return fr_mul(x, subgroup_generator()); }          // domain.rs:144-146
  TwoAdicMultiplicativeCoset create_disjoint_domain(size_t min_size) const {             // domain.rs:155-168
    unsigned l = 0;
    while (((size_t)1 << l) < min_size) l++;
    return TwoAdicMultiplicativeCoset(fr_mul(shift, fr_generator()), l);
  }
  std::vector<TwoAdicMultiplicativeCoset> split_domains(size_t num_chunks) const {       // domain.rs:174-186
    const unsigned lc = log2_strict(num_chunks);
    const Fr g = subgroup_generator();
    std::vector<TwoAdicMultiplicativeCoset> out;
    Fr s = shift;
    for (size_t i = 0; i < num_chunks; i++) {
      out.emplace_back(s, log_size - lc);
      s = fr_mul(s, g);
    }
    return out;
  }
  std::vector<RowMajorMatrix> split_evals(size_t num_chunks, const RowMajorMatrix& evals) const {  // domain.rs:188-221
    if (evals.height() != size()) throw std::logic_error("split_evals: height must match the domain size");
    const size_t w = evals.width(), hc = evals.height() / num_chunks;
    std::vector<RowMajorMatrix> out;
    for (size_t i = 0; i < num_chunks; i++) {
      std::vector<Fr> v(hc * w);
      for (size_t r = 0; r < hc; r++)
        std::memcpy(&v[r * w], &evals.values[(r * num_chunks + i) * w], w * sizeof(Fr));
      out.emplace_back(std::move(v), w);
    }
    return out;
  }
};
using Domain = TwoAdicMultiplicativeCoset;

// ---- TwoAdicSubgroupDft<Fr> -------------------------------------------------------------------------------------
// Clone + Default like the trait requires: copies share the context.
class GpuDft {
 public:
  explicit GpuDft(Context ctx) : ctx_(std::move(ctx)) {}
  RowMajorMatrix dft_batch(RowMajorMatrix m) const {                                       // traits.rs:61
    return run(m, 0, [&](const uint64_t* in, uint64_t* out, unsigned lh, size_t w) {
      return eon_dft_batch(ctx_.raw(), in, out, lh, w);
    });
  }
  RowMajorMatrix coset_dft_batch(RowMajorMatrix m, const Fr& shift) const {                // traits.rs:83-91
    return run(m, 0, [&](const uint64_t* in, uint64_t* out, unsigned lh, size_t w) {
      return eon_coset_dft_batch(ctx_.raw(), in, out, lh, w, wire(shift));
    });
  }
  RowMajorMatrix idft_batch(RowMajorMatrix m) const {                                      // traits.rs:111-122
    return run(m, 0, [&](const uint64_t* in, uint64_t* out, unsigned lh, size_t w) {
      return eon_idft_batch(ctx_.raw(), in, out, lh, w);
    });
  }
  RowMajorMatrix coset_idft_batch(RowMajorMatrix m, const Fr& shift) const {               // traits.rs:144-153
    return run(m, 0, [&](const uint64_t* in, uint64_t* out, unsigned lh, size_t w) {
      return eon_coset_idft_batch(ctx_.raw(), in, out, lh, w, wire(shift));
    });
  }
  RowMajorMatrix lde_batch(RowMajorMatrix m, unsigned added_bits) const {                  // traits.rs:187-192
    return coset_lde_batch(std::move(m), added_bits, Fr::one());
  }
  RowMajorMatrix coset_lde_batch(RowMajorMatrix m, unsigned added_bits, const Fr& shift) const {  // traits.rs:226-249
    return run(m, added_bits, [&](const uint64_t* in, uint64_t* out, unsigned lh, size_t w) {
      return eon_coset_lde_batch(ctx_.raw(), in, out, lh, w, added_bits, wire(shift));
    });
  }
  static const uint64_t* wire(const Fr& x) { return reinterpret_cast<const uint64_t*>(x.v); }

 private:
  template <class F>
  RowMajorMatrix run(const RowMajorMatrix& m, unsigned added_bits, F f) const {
    const unsigned lh = log2_strict(m.height());
    RowMajorMatrix out(std::vector<Fr>((m.height() << added_bits) * m.width()), m.width());
    ctx_.check(f(m.wire(), out.wire(), lh, m.width()));
    return out;
  }
  Context ctx_;
};

// ---- G1 ---------------------------------------------------------------------------------------------------------
struct G1 {
  std::array<uint64_t, 8> xy{};  // affine Montgomery wire point; all zero = identity
  static G1 identity() { return G1(); }
  bool is_identity() const {
    for (uint64_t v : xy) if (v) return false;
    return true;
  }
  bool operator==(const G1& o) const { return xy == o.xy; }
  bool operator!=(const G1& o) const { return !(*this == o); }
};
static_assert(sizeof(G1) == 64, "G1 wire size");

// G1::multi_exp, bn254/src/curve.rs:158-180
inline G1 multi_exp(const Context& ctx, const std::vector<G1>& points, const std::vector<Fr>& scalars) {
  if (points.size() != scalars.size()) throw std::logic_error("points and scalars must have the same length");
  G1 out;
  ctx.check(eon_msm_points(ctx.raw(), reinterpret_cast<const uint64_t*>(points.data()),
                           reinterpret_cast<const uint64_t*>(scalars.data()), points.size(), out.xy.data()));
  return out;
}
// G1::to_bytes (curve.rs:136-139) for a batch
inline std::vector<std::array<uint8_t, 32>> to_bytes(const Context& ctx, const std::vector<G1>& points,
                                                     int enc = EON_G1_ENC_HALO2) {
  std::vector<std::array<uint8_t, 32>> out(points.size());
  ctx.check(eon_g1_compress(ctx.raw(), reinterpret_cast<const uint64_t*>(points.data()), points.size(),
                            reinterpret_cast<uint8_t*>(out.data()), enc));
  return out;
}

// ---- KzgPcs -----------------------------------------------------------------------------------------------------
struct MatrixCommitment { std::vector<G1> columns; };                // pcs.rs:22-29
struct KzgCommitment { std::vector<MatrixCommitment> matrices; };     // pcs.rs:31-40
// proof.rounds[round][matrix][point] = one witness per column     (pcs.rs:42-50)
struct KzgProof { std::vector<std::vector<std::vector<std::vector<G1>>>> rounds; };
// opened_values[round][matrix][point] = one value per column       (commit/src/pcs.rs OpenedValues)
using OpenedValues = std::vector<std::vector<std::vector<std::vector<Fr>>>>;

// MatrixProverData (pcs.rs:52-61): evals stay on the host, coefficients on the device behind a handle that is
// released when the last copy goes away (Drop).
struct MatrixProverData {
  Domain domain;
  RowMajorMatrix evals;
  std::shared_ptr<eon_handle> handle;
};
using ProverData = std::vector<MatrixProverData>;

class GpuKzgPcs {
 public:
  static constexpr bool ZK = false;  // pcs.rs:216
  // KzgPcs::new -> init_srs_unsafe(max_degree, alpha), params.rs:123-139
  static GpuKzgPcs new_unsafe(Context ctx, size_t max_degree, const Fr& alpha) {
    ctx.check(eon_srs_generate_unsafe(ctx.raw(), GpuDft::wire(alpha), max_degree + 1));
    return GpuKzgPcs(std::move(ctx));
  }
  // from an existing SRS: g1_powers as affine points (normalised once, not per MSM as curve.rs:170)
  static GpuKzgPcs from_srs(Context ctx, const std::vector<G1>& g1_powers) {
    ctx.check(eon_srs_load_affine(ctx.raw(), reinterpret_cast<const uint64_t*>(g1_powers.data()), g1_powers.size()));
    return GpuKzgPcs(std::move(ctx));
  }
  // from a serialised SRS: g1_powers as 32-byte compressed points (Deserialize for G1, curve.rs:91-98)
  static GpuKzgPcs from_srs_bytes(Context ctx, const std::vector<std::array<uint8_t, 32>>& bytes,
                                  int enc = EON_G1_ENC_HALO2) {
    size_t bad = 0;
    int rc = eon_srs_load_compressed(ctx.raw(), reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), enc, &bad);
    if (rc == EON_ERR_BAD_POINT) throw InvalidG1Point(rc, eon_last_error(ctx.raw()), bad);
    ctx.check(rc);
    return GpuKzgPcs(std::move(ctx));
  }
  size_t max_degree() const { return eon_srs_size(ctx_.raw()) - 1; }
  const Context& context() const { return ctx_; }

  Domain natural_domain_for_degree(size_t degree) const {                                  // pcs.rs:218-221
    size_t n = 1;
    while (n < degree) n <<= 1;
    return Domain(Fr::one(), log2_strict(n));
  }

  // pcs.rs:223-265
  std::pair<KzgCommitment, ProverData> commit(std::vector<std::pair<Domain, RowMajorMatrix>> evaluations) const {
    KzgCommitment commitment;
    ProverData prover;
    for (auto& de : evaluations) {
      const Domain& domain = de.first;
      RowMajorMatrix& evals = de.second;
      if (evals.height() != domain.size()) throw std::logic_error("evaluation height must match domain size");
      MatrixCommitment mc;
      mc.columns.resize(evals.width());
      eon_handle h = 0;
      ctx_.check(eon_kzg_commit(ctx_.raw(), evals.wire(), domain.log_size, evals.width(), GpuDft::wire(domain.shift),
                                reinterpret_cast<uint64_t*>(mc.columns.data()), &h));
      Context keep = ctx_;
      std::shared_ptr<eon_handle> hp(new eon_handle(h), [keep](eon_handle* p) {
        eon_handle_free(keep.raw(), *p);
        delete p;
      });
      commitment.matrices.push_back(std::move(mc));
      prover.push_back(MatrixProverData{domain, std::move(evals), hp});
    }
    return {std::move(commitment), std::move(prover)};
  }

  // Pcs::commit_quotient, trait default commit/src/pcs.rs:82-102
  std::pair<KzgCommitment, ProverData> commit_quotient(const Domain& quotient_domain, const RowMajorMatrix& evals,
                                                       size_t num_chunks) const {
    auto subs = quotient_domain.split_evals(num_chunks, evals);
    auto doms = quotient_domain.split_domains(num_chunks);
    std::vector<std::pair<Domain, RowMajorMatrix>> in;
    for (size_t i = 0; i < num_chunks; i++) in.emplace_back(doms[i], std::move(subs[i]));
    return commit(std::move(in));
  }

  // pcs.rs:267-287 (zero-pad + coset NTT from the resident coefficients instead of the quadratic Horner loop)
  RowMajorMatrix get_evaluations_on_domain(const ProverData& pd, size_t idx, const Domain& domain) const {
    const MatrixProverData& m = pd.at(idx);
    if (m.domain.shift == domain.shift && m.domain.log_size == domain.log_size) return m.evals;
    RowMajorMatrix out(std::vector<Fr>(domain.size() * m.evals.width()), m.evals.width());
    ctx_.check(eon_kzg_evals_on_coset(ctx_.raw(), *m.handle, domain.log_size, GpuDft::wire(domain.shift), out.wire()));
    return out;
  }

  // pcs.rs:289-335: rounds of (prover data, opening points per matrix)
  std::pair<OpenedValues, KzgProof> open(
      const std::vector<std::pair<const ProverData*, std::vector<std::vector<Fr>>>>& rounds) const {
    OpenedValues values;
    KzgProof proof;
    for (const auto& rd : rounds) {
      const ProverData& pd = *rd.first;
      if (pd.size() != rd.second.size()) throw std::logic_error("one list of points per matrix");
      std::vector<std::vector<std::vector<Fr>>> mv;
      std::vector<std::vector<std::vector<G1>>> mp;
      for (size_t mi = 0; mi < pd.size(); mi++) {
        const auto& pts = rd.second[mi];
        const size_t w = pd[mi].evals.width(), np = pts.size();
        std::vector<Fr> vals(np * w);
        std::vector<G1> wits(np * w);
        ctx_.check(eon_kzg_open(ctx_.raw(), *pd[mi].handle, reinterpret_cast<const uint64_t*>(pts.data()), np,
                                reinterpret_cast<uint64_t*>(vals.data()), reinterpret_cast<uint64_t*>(wits.data())));
        std::vector<std::vector<Fr>> pv;
        std::vector<std::vector<G1>> pw;
        for (size_t p = 0; p < np; p++) {
          pv.emplace_back(vals.begin() + p * w, vals.begin() + (p + 1) * w);
          pw.emplace_back(wits.begin() + p * w, wits.begin() + (p + 1) * w);
        }
        mv.push_back(std::move(pv));
        mp.push_back(std::move(pw));
      }
      values.push_back(std::move(mv));
      proof.rounds.push_back(std::move(mp));
    }
    return {std::move(values), std::move(proof)};
  }

  // what CanObserve<KzgCommitment> absorbs (pcs.rs:409-438): compressed bytes -> four LE u64 -> Fr::from_u64
  std::vector<Fr> observe(const KzgCommitment& c, int enc = EON_G1_ENC_HALO2) const {
    std::vector<Fr> out;
    for (const auto& m : c.matrices)
      for (const auto& b : to_bytes(ctx_, m.columns, enc))
        for (int k = 0; k < 4; k++) {
          uint64_t v = 0;
          for (int i = 0; i < 8; i++) v |= (uint64_t)b[8 * k + i] << (8 * i);
          out.push_back(fr_from_u64(v));
        }
    return out;
  }

 private:
  explicit GpuKzgPcs(Context ctx) : ctx_(std::move(ctx)) {}
  Context ctx_;
};

// ---- the same Pcs over ONE multi-device context (eon_mctx_*): the compiled twin of rust/p3-eon-gpu/src/pcs.rs -------
// Whole host matrices in, the library shards their columns over its GPUs.  `commit` can carry the LDE hint,
// `commit_quotient` is the single batched call (no split_evals upload), `open` is ONE eon_mctx_kzg_open_batch whose
// flat [matrix][point][column] results are cut back into the nested OpenedValues / KzgProof.
class MultiContext {
 public:
  explicit MultiContext(const std::vector<int>& devices) {
    eon_mctx* m = nullptr;
    int rc = eon_mctx_create(devices.data(), (int)devices.size(), &m);
    if (rc != EON_OK) throw EonError(rc, "eon_mctx_create failed (no sm_100 device? there is no CPU fallback)");
    m_ = std::shared_ptr<eon_mctx>(m, [](eon_mctx* p) { eon_mctx_destroy(p); });
  }
  eon_mctx* raw() const { return m_.get(); }
  void check(int rc) const {
    if (rc == EON_OK) return;
    std::string msg = eon_mctx_last_error(m_.get());
    if (rc == EON_ERR_SRS_TOO_SHORT) throw DegreeTooLarge(rc, msg);
    throw EonError(rc, msg);
  }

 private:
  std::shared_ptr<eon_mctx> m_;
};

struct MultiMatrixProverData {
  Domain domain;
  RowMajorMatrix evals;
  std::shared_ptr<eon_handle> handle;
  std::shared_ptr<std::pair<Domain, RowMajorMatrix>> lde;  // produced inside commit() under an LDE hint
};
using MultiProverData = std::vector<MultiMatrixProverData>;

class MultiGpuKzgPcs {
 public:
  static MultiGpuKzgPcs new_unsafe(MultiContext ctx, size_t max_degree, const Fr& alpha) {
    ctx.check(eon_mctx_srs_generate_unsafe(ctx.raw(), GpuDft::wire(alpha), max_degree + 1));
    return MultiGpuKzgPcs(std::move(ctx));
  }
  // eon-uni-stark/src/prover.rs:186-187 then :307-322: the prover evaluates every committed trace on the quotient
  // coset right away; with the hint commit() produces that matrix in the same call
  MultiGpuKzgPcs with_lde_hint(unsigned added_bits, const Fr& shift) const {
    MultiGpuKzgPcs r = *this;
    r.hint_ = true;
    r.hint_bits_ = added_bits;
    r.hint_shift_ = shift;
    return r;
  }

  std::pair<KzgCommitment, MultiProverData> commit(std::vector<std::pair<Domain, RowMajorMatrix>> evaluations) const {
    KzgCommitment commitment;
    MultiProverData prover;
    for (auto& de : evaluations) {
      const Domain& domain = de.first;
      RowMajorMatrix& evals = de.second;
      if (evals.height() != domain.size()) throw std::logic_error("evaluation height must match domain size");
      MatrixCommitment mc;
      mc.columns.resize(evals.width());
      eon_handle h = 0;
      std::shared_ptr<std::pair<Domain, RowMajorMatrix>> lde;
      if (hint_ && evals.width() > 0) {
        Domain ld(hint_shift_, domain.log_size + hint_bits_);
        RowMajorMatrix out(std::vector<Fr>(ld.size() * evals.width()), evals.width());
        ctx_.check(eon_mctx_kzg_commit_lde(ctx_.raw(), evals.wire(), domain.log_size, evals.width(),
                                           GpuDft::wire(domain.shift), reinterpret_cast<uint64_t*>(mc.columns.data()), &h,
                                           ld.log_size, GpuDft::wire(ld.shift), out.wire()));
        lde = std::make_shared<std::pair<Domain, RowMajorMatrix>>(ld, std::move(out));
      } else {
        ctx_.check(eon_mctx_kzg_commit(ctx_.raw(), evals.wire(), domain.log_size, evals.width(),
                                       GpuDft::wire(domain.shift), reinterpret_cast<uint64_t*>(mc.columns.data()), &h));
      }
      commitment.matrices.push_back(std::move(mc));
      prover.push_back(MultiMatrixProverData{domain, std::move(evals), own(h), lde});
    }
    return {std::move(commitment), std::move(prover)};
  }

  // override of the trait default (commit/src/pcs.rs:82-102): one call, chunk i = rows i, i + c, ... of the matrix
  std::pair<KzgCommitment, MultiProverData> commit_quotient(const Domain& quotient_domain, const RowMajorMatrix& evals,
                                                            size_t num_chunks) const {
    const unsigned log_chunks = log2_strict(num_chunks);
    if (evals.height() != quotient_domain.size()) throw std::logic_error("evaluation height must match domain size");
    const size_t w = evals.width();
    std::vector<G1> cols(num_chunks * w);
    std::vector<eon_handle> hs(num_chunks, 0);
    ctx_.check(eon_mctx_kzg_commit_quotient(ctx_.raw(), evals.wire(), quotient_domain.log_size, w, log_chunks,
                                            GpuDft::wire(quotient_domain.shift), reinterpret_cast<uint64_t*>(cols.data()),
                                            hs.data()));
    auto doms = quotient_domain.split_domains(num_chunks);
    auto subs = quotient_domain.split_evals(num_chunks, evals);
    KzgCommitment commitment;
    MultiProverData prover;
    for (size_t i = 0; i < num_chunks; i++) {
      MatrixCommitment mc;
      mc.columns.assign(cols.begin() + i * w, cols.begin() + (i + 1) * w);
      commitment.matrices.push_back(std::move(mc));
      prover.push_back(MultiMatrixProverData{doms[i], std::move(subs[i]), own(hs[i]), nullptr});
    }
    return {std::move(commitment), std::move(prover)};
  }

  RowMajorMatrix get_evaluations_on_domain(const MultiProverData& pd, size_t idx, const Domain& domain) const {
    const MultiMatrixProverData& m = pd.at(idx);
    if (m.domain.shift == domain.shift && m.domain.log_size == domain.log_size) return m.evals;
    if (m.lde && m.lde->first.shift == domain.shift && m.lde->first.log_size == domain.log_size) return m.lde->second;
    RowMajorMatrix out(std::vector<Fr>(domain.size() * m.evals.width()), m.evals.width());
    ctx_.check(eon_mctx_kzg_evals_on_coset(ctx_.raw(), *m.handle, domain.log_size, GpuDft::wire(domain.shift), out.wire()));
    return out;
  }

  std::pair<OpenedValues, KzgProof> open(
      const std::vector<std::pair<const MultiProverData*, std::vector<std::vector<Fr>>>>& rounds) const {
    std::vector<eon_handle> handles;
    std::vector<size_t> npoints, widths;
    std::vector<Fr> points;
    for (const auto& rd : rounds) {
      if (rd.first->size() != rd.second.size()) throw std::logic_error("one list of points per matrix");
      for (size_t mi = 0; mi < rd.first->size(); mi++) {
        handles.push_back(*(*rd.first)[mi].handle);
        npoints.push_back(rd.second[mi].size());
        widths.push_back((*rd.first)[mi].evals.width());
        points.insert(points.end(), rd.second[mi].begin(), rd.second[mi].end());
      }
    }
    size_t total = 0;
    for (size_t i = 0; i < handles.size(); i++) total += npoints[i] * widths[i];
    std::vector<Fr> vals(total ? total : 1);
    std::vector<G1> wits(total ? total : 1);
    if (!handles.empty())
      ctx_.check(eon_mctx_kzg_open_batch(ctx_.raw(), handles.size(), handles.data(), npoints.data(),
                                         reinterpret_cast<const uint64_t*>(points.data()),
                                         reinterpret_cast<uint64_t*>(vals.data()), reinterpret_cast<uint64_t*>(wits.data())));
    OpenedValues values;
    KzgProof proof;
    size_t k = 0, i = 0;
    for (const auto& rd : rounds) {
      std::vector<std::vector<std::vector<Fr>>> mv;
      std::vector<std::vector<std::vector<G1>>> mp;
      for (size_t mi = 0; mi < rd.first->size(); mi++, i++) {
        const size_t w = widths[i], np = npoints[i];
        std::vector<std::vector<Fr>> pv;
        std::vector<std::vector<G1>> pw;
        for (size_t p = 0; p < np; p++) {
          pv.emplace_back(vals.begin() + k + p * w, vals.begin() + k + (p + 1) * w);
          pw.emplace_back(wits.begin() + k + p * w, wits.begin() + k + (p + 1) * w);
        }
        k += np * w;
        mv.push_back(std::move(pv));
        mp.push_back(std::move(pw));
      }
      values.push_back(std::move(mv));
      proof.rounds.push_back(std::move(mp));
    }
    return {std::move(values), std::move(proof)};
  }

  RowMajorMatrix coset_lde_batch(const RowMajorMatrix& m, unsigned added_bits, const Fr& shift) const {
    const unsigned log_h = log2_strict(m.height());
    RowMajorMatrix out(std::vector<Fr>((m.height() << added_bits) * m.width()), m.width());
    ctx_.check(eon_mctx_coset_lde_batch(ctx_.raw(), m.wire(), out.wire(), log_h, m.width(), added_bits, GpuDft::wire(shift)));
    return out;
  }

 private:
  explicit MultiGpuKzgPcs(MultiContext ctx) : ctx_(std::move(ctx)) {}
  std::shared_ptr<eon_handle> own(eon_handle h) const {
    MultiContext keep = ctx_;
    return std::shared_ptr<eon_handle>(new eon_handle(h), [keep](eon_handle* p) {
      eon_mctx_handle_free(keep.raw(), *p);
      delete p;
    });
  }
  MultiContext ctx_;
  bool hint_ = false;
  unsigned hint_bits_ = 0;
  Fr hint_shift_ = Fr::one();
};

}  // namespace p3eon
