// Fixed-operand (Shoup / Barrett-style) modular product for 8 x 32-bit limbs: the multiplier of the NTT
// butterflies (csrc/ntt.cu), whose second operand is always a precomputed twiddle.  Measured by
// eon_bench_modmul_variant (variant 4: 82.9 G products/s on B200 against 69.6 G/s for the word-serial Montgomery
// product without its final correction) and checked on the host by tests/test_host_arith.py.
//
//   given   w < p  and  wq = floor(w * 2^256 / p)   (both precomputed per twiddle),  a = any 256-bit value
//   q~ = floor( (sum_{i+j >= 6} a_i wq_j 2^(32(i+j))) / 2^256 )        43 limb products (not 64)
//   r  = (a*w - q~*p) mod 2^256                                        2 x 36 low-half limb products
// With Q = floor(a w / p):  q~ in {Q-2, Q-1, Q}  (wq is a floor: -1; the dropped columns 0..5 of a*wq sum to
// < 6*2^224 < 2^256: -1), so r = a*w mod p + {0, p, 2p}: r in [0, 3p) and 3p < 2^256, hence exact.
// Cost: 86 + 64 + 64 = 214 IMAD (lo/hi counted separately) against 272 for the word-serial Montgomery
// product (fp.cuh).  No Montgomery factor is involved: (a R) * w = (a w) R, so Montgomery-form data times a
// plain-form twiddle stays in Montgomery form.
//
// Products are accumulated like in fp.cuh: limb pairs at even columns in one array, at odd columns in another,
// so that every (lo, hi) pair sits on an aligned register pair; one carry chain per (row, parity).
#pragma once
#include "fp.cuh"

namespace eon {
namespace shoup {

// X[BASE + (i - I0)], X[BASE + (i - I0) + 1] += lo, hi of a[i] * b for i = I0, I0 + 2, ... <= 7; the carry out of
// the chain is deposited in the next limb (which holds only such deposits so far: see the column argument in
// hi_row), so nothing is lost.
template <int I0, int BASE>
EON_HD void chain_full(u32* X, const u32 a[8], u32 b) {
  static_assert(I0 >= 0 && I0 <= 7 && BASE >= 0, "chain out of range");
  X[BASE] = cc::mad_lo_cc(a[I0], b, X[BASE]);
  X[BASE + 1] = cc::madc_hi_cc(a[I0], b, X[BASE + 1]);
#pragma unroll
  for (int i = I0 + 2; i <= 7; i += 2) {
    X[BASE + i - I0] = cc::madc_lo_cc(a[i], b, X[BASE + i - I0]);
    X[BASE + i - I0 + 1] = cc::madc_hi_cc(a[i], b, X[BASE + i - I0 + 1]);
  }
  constexpr int TOP = BASE + ((7 - I0) / 2) * 2 + 2;
  X[TOP] = cc::addc(X[TOP], 0);
}

// Row J of the truncated high product: all a_i * wq_J with i + J >= 6.  E[k] is column k + 6 (pairs at even
// columns), O[k] is column k + 7 (pairs at odd columns).  A row's chains end at column <= J + 8 and deposit
// their carry one column above every product column written so far (<= J + 7 before row J + 1 starts).
template <int J>
EON_HD void hi_row(u32* E, u32* O, const u32 a[8], u32 wj) {
  constexpr int IMIN = (6 - J) > 0 ? (6 - J) : 0;
  constexpr int I0E = ((IMIN + J) % 2 == 0) ? IMIN : IMIN + 1;
  constexpr int I0O = ((IMIN + J) % 2 == 1) ? IMIN : IMIN + 1;
  chain_full<I0E, I0E + J - 6>(E, a, wj);
  chain_full<I0O, I0O + J - 7>(O, a, wj);
}

// Row J of a low-half product accumulation: L + 2^32 M += x * y_J * 2^(32 J)  (mod 2^256).
// L[c], L[c+1]: pair at even column c;  M[c-1], M[c]: pair at odd column c (M[k] is column k + 1); the product at
// column 7 contributes its low word only.  Every chain reaches column 7, so no carry needs a deposit.
template <int J>
EON_HD void lo_row(u32* L, u32* M, const u32 x[8], u32 y) {
  constexpr int I0E = J % 2;  // i + J even
  if constexpr (I0E + J <= 6) {
    L[I0E + J] = cc::mad_lo_cc(x[I0E], y, L[I0E + J]);
    L[I0E + J + 1] = cc::madc_hi_cc(x[I0E], y, L[I0E + J + 1]);
#pragma unroll
    for (int i = I0E + 2; i + J <= 6; i += 2) {
      L[i + J] = cc::madc_lo_cc(x[i], y, L[i + J]);
      L[i + J + 1] = cc::madc_hi_cc(x[i], y, L[i + J + 1]);
    }
  }
  constexpr int I0O = (J + 1) % 2;  // i + J odd
  if constexpr (I0O + J == 7) {
    M[6] += cc::mul_lo(x[I0O], y);
  } else {
    M[I0O + J - 1] = cc::mad_lo_cc(x[I0O], y, M[I0O + J - 1]);
    M[I0O + J] = cc::madc_hi_cc(x[I0O], y, M[I0O + J]);
#pragma unroll
    for (int i = I0O + 2; i + J <= 5; i += 2) {
      M[i + J - 1] = cc::madc_lo_cc(x[i], y, M[i + J - 1]);
      M[i + J] = cc::madc_hi_cc(x[i], y, M[i + J]);
    }
    M[6] = cc::madc_lo_cc(x[7 - J], y, M[6]);  // the term at column 7: low word (carry out is beyond 2^256)
  }
}

// limb i of 2^256 - p (p odd, so the +1 of the two's complement never carries out of limb 0)
template <class PP>
EON_HD constexpr u32 neg_mod(int i) { return i == 0 ? (0u - PP::mod(0)) : ~PP::mod(i); }

// r = a*w - q~*p in [0, 3p),  r == a*w (mod p).   a: any 256-bit value;  w < p;  wq = floor(w 2^256 / p).
template <class PP>
EON_HD void mul_lazy(u32 r[8], const u32 a[8], const u32 w[8], const u32 wq[8]) {
  u32 E[12], O[12];
#pragma unroll
  for (int i = 0; i < 12; i++) E[i] = O[i] = 0;
  hi_row<0>(E, O, a, wq[0]);
  hi_row<1>(E, O, a, wq[1]);
  hi_row<2>(E, O, a, wq[2]);
  hi_row<3>(E, O, a, wq[3]);
  hi_row<4>(E, O, a, wq[4]);
  hi_row<5>(E, O, a, wq[5]);
  hi_row<6>(E, O, a, wq[6]);
  hi_row<7>(E, O, a, wq[7]);
  // q~ = columns 8..15 of E + 2^32 O (column 7 only feeds its carry)
  u32 q[8];
  (void)cc::add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 0; k < 7; k++) q[k] = cc::addc_cc(E[k + 2], O[k + 1]);
  q[7] = cc::addc(E[9], O[8]);

  u32 L[8], M[8];
#pragma unroll
  for (int i = 0; i < 8; i++) L[i] = M[i] = 0;
  lo_row<0>(L, M, a, w[0]);
  lo_row<1>(L, M, a, w[1]);
  lo_row<2>(L, M, a, w[2]);
  lo_row<3>(L, M, a, w[3]);
  lo_row<4>(L, M, a, w[4]);
  lo_row<5>(L, M, a, w[5]);
  lo_row<6>(L, M, a, w[6]);
  lo_row<7>(L, M, a, w[7]);
  // - q~ * p  ==  + q~ * (2^256 - p)   (mod 2^256)
  lo_row<0>(L, M, q, neg_mod<PP>(0));
  lo_row<1>(L, M, q, neg_mod<PP>(1));
  lo_row<2>(L, M, q, neg_mod<PP>(2));
  lo_row<3>(L, M, q, neg_mod<PP>(3));
  lo_row<4>(L, M, q, neg_mod<PP>(4));
  lo_row<5>(L, M, q, neg_mod<PP>(5));
  lo_row<6>(L, M, q, neg_mod<PP>(6));
  lo_row<7>(L, M, q, neg_mod<PP>(7));
  r[0] = L[0];
  r[1] = cc::add_cc(L[1], M[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) r[k] = cc::addc_cc(L[k], M[k - 1]);
  r[7] = cc::addc(L[7], M[6]);
}

// limb i of p^-1 mod 2^256
template <class PP>
EON_HD constexpr u32 pinv256(int i);
template <>
EON_HD constexpr u32 pinv256<FrParams>(int i) { constexpr u32 m[8] = EON_FR_PINV256; return m[i]; }
template <>
EON_HD constexpr u32 pinv256<FqParams>(int i) { constexpr u32 m[8] = EON_FQ_PINV256; return m[i]; }

// The operand pair of a twiddle from its Montgomery form rho = w * 2^256 mod p:
//   w  = rho / 2^256 mod p                      (plain form)
//   wq = floor(w * 2^256 / p) = (w 2^256 - rho) / p, an exact division, hence = -rho * p^-1 (mod 2^256)
template <class PP>
EON_HD void precompute(u32 w[8], u32 wq[8], const Fp<PP>& rho) {
  fp_from_mont<PP>(w, rho);
  u32 n[8];
  n[0] = cc::sub_cc(0u, rho.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) n[i] = cc::subc_cc(0u, rho.v[i]);
  n[7] = cc::subc(0u, rho.v[7]);
  u32 L[8], M[8];
#pragma unroll
  for (int i = 0; i < 8; i++) L[i] = M[i] = 0;
  lo_row<0>(L, M, n, pinv256<PP>(0));
  lo_row<1>(L, M, n, pinv256<PP>(1));
  lo_row<2>(L, M, n, pinv256<PP>(2));
  lo_row<3>(L, M, n, pinv256<PP>(3));
  lo_row<4>(L, M, n, pinv256<PP>(4));
  lo_row<5>(L, M, n, pinv256<PP>(5));
  lo_row<6>(L, M, n, pinv256<PP>(6));
  lo_row<7>(L, M, n, pinv256<PP>(7));
  wq[0] = L[0];
  wq[1] = cc::add_cc(L[1], M[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) wq[k] = cc::addc_cc(L[k], M[k - 1]);
  wq[7] = cc::addc(L[7], M[6]);
}

// r in [0, 3p) -> canonical
template <class PP>
EON_HD Fp<PP> canon_3p(const u32 a[8]) {
  u32 t[8];
  fp_final_sub<PP>(t, a);  // a < 3p: after one subtraction < 2p ... only if a >= p; fp_final_sub keeps a when a < p
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

}  // namespace shoup
}  // namespace eon
