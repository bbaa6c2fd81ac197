// C++ twin of the reference's own tests, driven through include/eon_kzg.hpp (the compiled-language host mirror
// of the Rust plugin surface) on a real GPU.  Expected values are computed on the host with the library's own
// field / curve arithmetic compiled for the CPU (fp.cuh, ec.cuh with a software carry flag), which
// tests/test_host_arith.py pins to the big-int oracle.
//
//   test_g1_multi_exp          bn254/src/curve.rs:597-628
//   NaiveDft::basic            dft/src/naive.rs:49-85
//   dft round trips            dft/src/naive.rs:87-104, field-testing/src/dft_testing.rs
//   pcs_roundtrip              kzg/src/tests.rs:19-48   (alpha = 7, evals x + 1 on the size-8 subgroup)
//   degree guard / height      kzg/src/pcs.rs:233-240
//   commit_quotient            commit/src/pcs.rs:82-102
//
// Build (tests/test_cxx_mirror.py does this): g++ -std=c++17 -O2 cxx_mirror_test.cpp -L<pkg> -leon_kzg
#include <cstdio>
#include <random>

#include "../../include/eon_kzg.hpp"
#include "../../plonky3_eon_b200/csrc/ec.cuh"

using namespace p3eon;

static int g_fail = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      g_fail++;                                                          \
    }                                                                    \
  } while (0)

static Fr fr(uint64_t x) { return fr_from_u64(x); }
static Fr fr_neg_u64(uint64_t x) { return eon::fp_neg(fr(x)); }

// k * G on the host
static G1 kG(uint64_t k) {
  uint32_t kk[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
  eon::G1Affine a = eon::g1_to_affine(eon::g1_mul_canonical(eon::G1Affine::generator(), kk));
  G1 r;
  std::memcpy(r.xy.data(), a.x.v, 32);
  std::memcpy(r.xy.data() + 4, a.y.v, 32);
  return r;
}

static Fr random_fr(std::mt19937_64& rng) {  // the reference's sampler, field.rs:534-551
  for (;;) {
    Fr x;
    for (int i = 0; i < 8; i += 2) {
      uint64_t v = rng();
      x.v[i] = (uint32_t)v;
      x.v[i + 1] = (uint32_t)(v >> 32);
    }
    x.v[7] &= 0x3fffffffu;
    bool lt = false;
    for (int i = 7; i >= 0; i--) {
      if (x.v[i] != eon::FrParams::mod(i)) { lt = x.v[i] < eon::FrParams::mod(i); break; }
    }
    if (lt) return x;
  }
}

static void test_g1_multi_exp(const Context& ctx) {
  CHECK(multi_exp(ctx, {}, {}).is_identity());
  CHECK(multi_exp(ctx, {kG(1)}, {fr(5)}) == kG(5));
  CHECK(multi_exp(ctx, {kG(1), kG(1)}, {fr(2), fr(3)}) == kG(5));
  CHECK(multi_exp(ctx, {kG(7), kG(11)}, {fr(3), fr(5)}) == kG(76));
  bool threw = false;
  try { multi_exp(ctx, {kG(1)}, {fr(1), fr(2)}); } catch (const std::logic_error&) { threw = true; }
  CHECK(threw);  // length mismatch panics, curve.rs:159-163
}

static void test_naive_dft_basic(const Context& ctx) {
  GpuDft dft(ctx);
  // columns 5 + 4x, 2 + 3x, 0  ->  rows [9, 5, 0], [1, -1, 0]
  RowMajorMatrix m({fr(5), fr(2), fr(0), fr(4), fr(3), fr(0)}, 3);
  RowMajorMatrix out = dft.dft_batch(m);
  const Fr want[6] = {fr(9), fr(5), fr(0), fr(1), fr_neg_u64(1), fr(0)};
  for (int i = 0; i < 6; i++) CHECK(out.values[i] == want[i]);
  bool threw = false;
  try { dft.dft_batch(RowMajorMatrix(std::vector<Fr>(9, fr(1)), 3)); } catch (const std::logic_error&) { threw = true; }
  CHECK(threw);  // height 3: log2_strict_usize panics
}

static void test_dft_roundtrips(const Context& ctx) {
  GpuDft dft(ctx);
  std::mt19937_64 rng(1);
  const size_t h = 64, w = 3;
  std::vector<Fr> v(h * w);
  for (auto& x : v) x = random_fr(rng);
  RowMajorMatrix m(v, w);
  const Fr g = fr_generator();
  CHECK(dft.idft_batch(dft.dft_batch(m)).values == m.values);
  CHECK(dft.coset_idft_batch(dft.coset_dft_batch(m, g), g).values == m.values);
  // coset_lde_batch(m, k, s) == coset_dft(zero-pad(idft(m)), s)   (traits.rs:226-249)
  RowMajorMatrix coeffs = dft.idft_batch(m);
  std::vector<Fr> padded(4 * h * w, Fr::zero());
  std::copy(coeffs.values.begin(), coeffs.values.end(), padded.begin());
  RowMajorMatrix want = dft.coset_dft_batch(RowMajorMatrix(padded, w), g);
  CHECK(dft.coset_lde_batch(m, 2, g).values == want.values);
  CHECK(dft.lde_batch(m, 1).height() == 2 * h);
}

static void test_pcs_roundtrip(const Context& ctx) {
  GpuKzgPcs pcs = GpuKzgPcs::new_unsafe(ctx, 8, fr(7));
  CHECK(pcs.max_degree() == 8);
  Domain domain(Fr::one(), 3);
  std::vector<Fr> evals;
  Fr x = domain.first_point();
  for (size_t i = 0; i < domain.size(); i++) {
    evals.push_back(eon::fp_add(x, Fr::one()));
    x = domain.next_point(x);
  }
  auto [commit, prover_data] = pcs.commit({{domain, RowMajorMatrix(evals, 1)}});
  // coefficients are [1, 1, 0, ...]: commit = (1 + alpha) G = 8 G
  CHECK(commit.matrices.size() == 1 && commit.matrices[0].columns.size() == 1);
  CHECK(commit.matrices[0].columns[0] == kG(8));
  auto [opened, proof] = pcs.open({{&prover_data, {{fr(2)}}}});
  CHECK(opened[0][0][0][0] == fr(3));          // f(2) = 3
  CHECK(proof.rounds[0][0][0][0] == kG(1));    // quotient (x + 1 - 3) / (x - 2) = 1
  // get_evaluations_on_domain: same domain returns the evaluations; the disjoint domain is the LDE
  CHECK(pcs.get_evaluations_on_domain(prover_data, 0, domain).values == evals);
  Domain q = domain.create_disjoint_domain(16);
  RowMajorMatrix lde = pcs.get_evaluations_on_domain(prover_data, 0, q);
  CHECK(lde.height() == 16);
  Fr y = q.first_point();
  for (size_t i = 0; i < 16; i++) {
    CHECK(lde.values[i] == eon::fp_add(y, Fr::one()));
    y = q.next_point(y);
  }
  // what the challenger would absorb: 4 field elements per column commitment
  CHECK(pcs.observe(commit).size() == 4);
  // x of 8 G, little-endian, with the flag bits of byte 31 clear for the low three words
  auto bytes = to_bytes(ctx, commit.matrices[0].columns);
  uint32_t xc[8];
  eon::Fq gx;
  std::memcpy(gx.v, commit.matrices[0].columns[0].xy.data(), 32);
  eon::fp_from_mont<eon::FqParams>(xc, gx);
  CHECK(std::memcmp(bytes[0].data(), xc, 28) == 0);
}

static void test_guards(const Context& ctx) {
  GpuKzgPcs pcs = GpuKzgPcs::new_unsafe(ctx, 3, fr(12345));  // SRS of 4 powers
  Domain d3(Fr::one(), 3), d2(Fr::one(), 2);
  bool degree = false, height = false;
  try { pcs.commit({{d3, RowMajorMatrix(std::vector<Fr>(8, fr(1)), 1)}}); } catch (const DegreeTooLarge&) { degree = true; }
  CHECK(degree);   // ensure_supported(...).unwrap(), pcs.rs:238-240
  try { pcs.commit({{d3, RowMajorMatrix(std::vector<Fr>(4, fr(1)), 1)}}); } catch (const std::logic_error&) { height = true; }
  CHECK(height);   // pcs.rs:233-237
  auto [c, pd] = pcs.commit({{d2, RowMajorMatrix(std::vector<Fr>(4, fr(9)), 1)}});
  CHECK(c.matrices[0].columns[0] == kG(9));  // constant polynomial 9
}

static void test_commit_quotient(const Context& ctx) {
  GpuKzgPcs pcs = GpuKzgPcs::new_unsafe(ctx, 15, fr(12345));
  std::mt19937_64 rng(7);
  Domain qd = Domain(Fr::one(), 3).create_disjoint_domain(16);  // 5 * <omega_16>
  std::vector<Fr> v(16);
  for (auto& x : v) x = random_fr(rng);
  RowMajorMatrix q(v, 1);
  auto [c, pd] = pcs.commit_quotient(qd, q, 2);
  CHECK(c.matrices.size() == 2 && pd.size() == 2);
  // chunk i holds rows i, i + 2, ... on the coset (5 omega_16^i) <omega_8>: committing it directly agrees
  auto doms = qd.split_domains(2);
  auto subs = qd.split_evals(2, q);
  for (int i = 0; i < 2; i++) {
    CHECK(subs[i].values[1] == v[2 + i]);
    auto [ci, pdi] = pcs.commit({{doms[i], subs[i]}});
    CHECK(ci.matrices[0].columns[0] == c.matrices[i].columns[0]);
  }
}

// The PCS call sequence of eon_uni_stark::prove (prover.rs:186-187, 307-322, 371-372, 416-442) through the
// single-device mirror and through the multi-device one (device 0 listed twice: two column shards): identical
// commitments, evaluations, opened values and witnesses.
static void test_multi_device_sequence(const Context& ctx) {
  const unsigned log_h = 10;
  const size_t h = 1u << log_h, w = 5;
  std::mt19937_64 rng(11);
  std::vector<Fr> v(h * w), qv(2 * h);
  for (auto& x : v) x = random_fr(rng);
  for (auto& x : qv) x = random_fr(rng);
  Domain dom(Fr::one(), log_h);
  Domain qdom = dom.create_disjoint_domain(2 * h);
  const Fr zeta = random_fr(rng);
  const Fr zeta_next = dom.next_point(zeta);

  GpuKzgPcs one = GpuKzgPcs::new_unsafe(ctx, 2 * h - 1, fr(12345));
  auto [c1, pd1] = one.commit({{dom, RowMajorMatrix(v, w)}});
  RowMajorMatrix lde1 = one.get_evaluations_on_domain(pd1, 0, qdom);
  auto [qc1, qpd1] = one.commit_quotient(qdom, RowMajorMatrix(qv, 1), 2);
  auto [ov1, pr1] = one.open({{&pd1, {{zeta, zeta_next}}}, {&qpd1, {{zeta}, {zeta}}}});

  MultiGpuKzgPcs many = MultiGpuKzgPcs::new_unsafe(MultiContext({0, 0}), 2 * h - 1, fr(12345)).with_lde_hint(1, fr_generator());
  auto [c2, pd2] = many.commit({{dom, RowMajorMatrix(v, w)}});
  RowMajorMatrix lde2 = many.get_evaluations_on_domain(pd2, 0, qdom);
  auto [qc2, qpd2] = many.commit_quotient(qdom, RowMajorMatrix(qv, 1), 2);
  auto [ov2, pr2] = many.open({{&pd2, {{zeta, zeta_next}}}, {&qpd2, {{zeta}, {zeta}}}});

  CHECK(c1.matrices[0].columns == c2.matrices[0].columns);
  CHECK(lde1.values == lde2.values);
  CHECK(many.coset_lde_batch(RowMajorMatrix(v, w), 1, fr_generator()).values == lde1.values);
  CHECK(qc1.matrices.size() == 2 && qc2.matrices.size() == 2);
  for (int i = 0; i < 2; i++) CHECK(qc1.matrices[i].columns == qc2.matrices[i].columns);
  CHECK(ov1 == ov2);
  CHECK(pr1.rounds.size() == 2 && pr2.rounds.size() == 2);
  for (size_t r = 0; r < pr1.rounds.size(); r++) CHECK(pr1.rounds[r] == pr2.rounds[r]);
  // a domain the hint does not cover still goes through eon_mctx_kzg_evals_on_coset
  Domain other(fr(7), log_h + 1);
  CHECK(one.get_evaluations_on_domain(pd1, 0, other).values == many.get_evaluations_on_domain(pd2, 0, other).values);
}

int main() {
  try {
    Context ctx(0);
    test_g1_multi_exp(ctx);
    test_naive_dft_basic(ctx);
    test_dft_roundtrips(ctx);
    test_pcs_roundtrip(ctx);
    test_guards(ctx);
    test_commit_quotient(ctx);
    test_multi_device_sequence(ctx);
  } catch (const std::exception& e) {
    std::printf("FAIL: exception %s\n", e.what());
    return 2;
  }
  if (g_fail) {
    std::printf("%d checks failed\n", g_fail);
    return 1;
  }
  std::printf("ALL OK\n");
  return 0;
}
