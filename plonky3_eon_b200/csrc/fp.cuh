// 256-bit Montgomery prime-field arithmetic for BN254 Fr and Fq on sm_100a.
//
// Replaces (bit-exactly) the reference's 4x64-bit CPU arithmetic:
//   monty_mul                  bn254/src/helpers.rs:188-205
//   Add / Sub for Fr           bn254/src/field.rs:464-508
//   Fq arithmetic              halo2curves 0.9 bn256::Fq (not vendored in the reference)
// Elements are 8 x u32 little-endian limbs in Montgomery form (a*2^256 mod p), always
// fully reduced (< p) on entry and exit, so the bytes equal the reference's [u64;4].
//
// The multiplier is a word-serial (CIOS) Montgomery product over 32-bit digits written as
// mad.lo.cc / madc.hi.cc carry chains.  Products a[j]*b_i with even j and odd j are
// accumulated in two separate 8-limb arrays ("even" aligned at limb 0, "odd" at limb 1) so
// that every (lo,hi) pair lands on an aligned register pair and ptxas can fuse the pair
// into one IMAD.WIDE.U32 with carry-in/out.  After each digit the two arrays swap roles,
// which performs the divide-by-2^32 without moving registers.
//
// The same source compiles for the host (g++/nvcc host pass) with the carry flag emulated
// in software; tests/host/ uses that to check the arithmetic on CPU (no GPU needed).
#pragma once
#include <stdint.h>
#include "consts.cuh"

#if defined(__CUDACC__)
#define EON_HD __host__ __device__ __forceinline__
#define EON_D __device__ __forceinline__
#else
#define EON_HD inline
#define EON_D inline
#endif

namespace eon {

typedef uint32_t u32;
typedef uint64_t u64;

// ------------------------------------------------------------------------------------
// carry-chain primitives: PTX on device, software carry flag on host
// ------------------------------------------------------------------------------------
namespace cc {
#if defined(__CUDA_ARCH__)
EON_D u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
EON_D u32 mul_lo(u32 a, u32 b) { return a * b; }
EON_D u32 mul_hi(u32 a, u32 b) { return __umulhi(a, b); }
EON_D u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
EON_D u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
EON_D u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
EON_D u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
EON_D u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
static thread_local u32 g_cf = 0;  // carry (add chains) / borrow (sub chains), like PTX CC.CF
inline u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 addc(u32 a, u32 b) { return a + b + g_cf; }
inline u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; g_cf = (u32)((t >> 32) & 1); return (u32)t; }
inline u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - g_cf; g_cf = (u32)((t >> 32) & 1); return (u32)t; }
inline u32 subc(u32 a, u32 b) { return a - b - g_cf; }
inline u32 mul_lo(u32 a, u32 b) { return a * b; }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)(((u64)a * b) >> 32) + c + g_cf; }
#endif
}  // namespace cc

// ------------------------------------------------------------------------------------
// field parameter packs
// ------------------------------------------------------------------------------------
struct FrParams {
  static EON_HD constexpr u32 mod(int i) { constexpr u32 m[8] = EON_FR_MOD; return m[i]; }
  static EON_HD constexpr u32 one(int i) { constexpr u32 m[8] = EON_FR_ONE; return m[i]; }
  static EON_HD constexpr u32 r2(int i) { constexpr u32 m[8] = EON_FR_R2; return m[i]; }
  static constexpr u32 INV = EON_FR_INV32;
};
struct FqParams {
  static EON_HD constexpr u32 mod(int i) { constexpr u32 m[8] = EON_FQ_MOD; return m[i]; }
  static EON_HD constexpr u32 one(int i) { constexpr u32 m[8] = EON_FQ_ONE; return m[i]; }
  static EON_HD constexpr u32 r2(int i) { constexpr u32 m[8] = EON_FQ_R2; return m[i]; }
  static constexpr u32 INV = EON_FQ_INV32;
};

// ------------------------------------------------------------------------------------
// Fp<Params>: value type, 8 x u32, Montgomery form, canonical (< p)
// ------------------------------------------------------------------------------------
template <class PP>
struct Fp {
  u32 v[8];

  static EON_HD Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static EON_HD Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = PP::one(i);
    return r;
  }
  static EON_HD Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = PP::r2(i);
    return r;
  }
  EON_HD bool is_zero() const {
    u32 o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  EON_HD bool operator==(const Fp& b) const {
    u32 o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  EON_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a   (a < 2p)
template <class PP>
EON_HD void fp_final_sub(u32 r[8], const u32 a[8]) {
  u32 s[8];
  s[0] = cc::sub_cc(a[0], PP::mod(0));
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = cc::subc_cc(a[i], PP::mod(i));
  u32 borrow = cc::subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? a[i] : s[i];
}

// Canonical modular add: (a + b) mod p, a,b < p.   Reference: field.rs:464-485.
template <class PP>
EON_HD Fp<PP> fp_add(const Fp<PP>& a, const Fp<PP>& b) {
  u32 t[8];
  t[0] = cc::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i] = cc::addc_cc(a.v[i], b.v[i]);
  t[7] = cc::addc(a.v[7], b.v[7]);  // < 2^255, no carry out
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

// Canonical modular sub: (a - b) mod p.   Reference: field.rs:487-508.
template <class PP>
EON_HD Fp<PP> fp_sub(const Fp<PP>& a, const Fp<PP>& b) {
  u32 t[8];
  t[0] = cc::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = cc::subc_cc(a.v[i], b.v[i]);
  u32 mask = cc::subc(0, 0);  // all-ones if borrow
  Fp<PP> r;
  r.v[0] = cc::add_cc(t[0], PP::mod(0) & mask);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = cc::addc_cc(t[i], PP::mod(i) & mask);
  r.v[7] = cc::addc(t[7], PP::mod(7) & mask);
  return r;
}

// ---- lazily reduced arithmetic (NTT butterflies) ---------------------------------------------------
// Both moduli are < 2^254, so 4p < 2^256: values may float in [0, 2p) or [0, 4p) between butterfly
// layers and are canonicalised once at the end, saving two of the three conditional corrections per
// butterfly (Harvey-style).  All of these take and return raw limbs, not canonical Fp values.

// r = a - 2p if a >= 2p else a      (a < 4p  ->  r < 2p)
template <class PP>
EON_HD void fp_reduce_2p(u32 r[8], const u32 a[8]) {
  constexpr u32 C = 0;
  (void)C;
  u32 p2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) p2[i] = (PP::mod(i) << 1) | (i ? (PP::mod(i - 1) >> 31) : 0u);
  u32 s[8];
  s[0] = cc::sub_cc(a[0], p2[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = cc::subc_cc(a[i], p2[i]);
  u32 borrow = cc::subc(0, 0);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? a[i] : s[i];
}

// r = a + b  (no reduction; caller guarantees a + b < 2^256)
EON_HD void fp_add_raw(u32 r[8], const u32 a[8], const u32 b[8]) {
  r[0] = cc::add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = cc::addc_cc(a[i], b[i]);
  r[7] = cc::addc(a[7], b[7]);
}

// r = a - b + 2p  (a, b < 2p  ->  0 < r < 4p)
template <class PP>
EON_HD void fp_sub_plus_2p(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 p2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) p2[i] = (PP::mod(i) << 1) | (i ? (PP::mod(i - 1) >> 31) : 0u);
  u32 t[8];
  t[0] = cc::add_cc(a[0], p2[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i] = cc::addc_cc(a[i], p2[i]);
  t[7] = cc::addc(a[7], p2[7]);
  r[0] = cc::sub_cc(t[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = cc::subc_cc(t[i], b[i]);
  r[7] = cc::subc(t[7], b[7]);
}

// canonical value of a < 4p
template <class PP>
EON_HD Fp<PP> fp_canon_4p(const u32 a[8]) {
  u32 t[8];
  fp_reduce_2p<PP>(t, a);
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

template <class PP>
EON_HD Fp<PP> fp_neg(const Fp<PP>& a) {
  // p - a, with 0 -> 0
  u32 t[8];
  t[0] = cc::sub_cc(PP::mod(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i] = cc::subc_cc(PP::mod(i), a.v[i]);
  t[7] = cc::subc(PP::mod(7), a.v[7]);
  bool z = a.is_zero();
  Fp<PP> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : t[i];
  return r;
}

template <class PP>
EON_HD Fp<PP> fp_dbl(const Fp<PP>& a) { return fp_add(a, a); }

// ---- Montgomery product ------------------------------------------------------------
namespace detail {
// One Montgomery reduction digit.  On entry the running total is T = E + 2^32 * O.
// m = E[0] * (-p^-1); E += sum_{j even} p[j]*m*2^(32j); O += sum_{j odd} p[j]*m*2^(32(j-1)).
// Afterwards E[0] == 0.  The carry out of E (weight 2^256) is folded into O[7] (same weight).
template <class PP>
EON_HD void redc_digit(u32 E[8], u32 O[8]) {
  const u32 m = E[0] * PP::INV;
  O[0] = cc::mad_lo_cc(PP::mod(1), m, O[0]);
  O[1] = cc::madc_hi_cc(PP::mod(1), m, O[1]);
  O[2] = cc::madc_lo_cc(PP::mod(3), m, O[2]);
  O[3] = cc::madc_hi_cc(PP::mod(3), m, O[3]);
  O[4] = cc::madc_lo_cc(PP::mod(5), m, O[4]);
  O[5] = cc::madc_hi_cc(PP::mod(5), m, O[5]);
  O[6] = cc::madc_lo_cc(PP::mod(7), m, O[6]);
  O[7] = cc::madc_hi(PP::mod(7), m, O[7]);  // total < 2^288: no carry out
  E[0] = cc::mad_lo_cc(PP::mod(0), m, E[0]);
  E[1] = cc::madc_hi_cc(PP::mod(0), m, E[1]);
  E[2] = cc::madc_lo_cc(PP::mod(2), m, E[2]);
  E[3] = cc::madc_hi_cc(PP::mod(2), m, E[3]);
  E[4] = cc::madc_lo_cc(PP::mod(4), m, E[4]);
  E[5] = cc::madc_hi_cc(PP::mod(4), m, E[5]);
  E[6] = cc::madc_lo_cc(PP::mod(6), m, E[6]);
  E[7] = cc::madc_hi_cc(PP::mod(6), m, E[7]);
  O[7] = cc::addc(O[7], 0);
}

// First digit: E = even-limb products, O = odd-limb products (no accumulation yet).
template <class PP>
EON_HD void mul_first(u32 E[8], u32 O[8], const u32 a[8], u32 bi) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    E[j] = cc::mul_lo(a[j], bi);
    E[j + 1] = cc::mul_hi(a[j], bi);
    O[j] = cc::mul_lo(a[j + 1], bi);
    O[j + 1] = cc::mul_hi(a[j + 1], bi);
  }
  redc_digit<PP>(E, O);
}

// Next digit.  X = previous odd array (becomes the even array after the 2^-32 shift),
// Y = previous even array with Y[0] == 0 (its limbs 2..7 become the new odd array,
// limb 1 is added at the bottom of X).  Then X += a_even*bi, Y += a_odd*bi, reduce.
template <class PP>
EON_HD void mul_next(u32 X[8], u32 Y[8], const u32 a[8], u32 bi) {
  X[0] = cc::add_cc(X[0], Y[1]);
  Y[0] = cc::madc_lo_cc(a[1], bi, Y[2]);
  Y[1] = cc::madc_hi_cc(a[1], bi, Y[3]);
  Y[2] = cc::madc_lo_cc(a[3], bi, Y[4]);
  Y[3] = cc::madc_hi_cc(a[3], bi, Y[5]);
  Y[4] = cc::madc_lo_cc(a[5], bi, Y[6]);
  Y[5] = cc::madc_hi_cc(a[5], bi, Y[7]);
  Y[6] = cc::madc_lo_cc(a[7], bi, 0);
  Y[7] = cc::madc_hi(a[7], bi, 0);
  X[0] = cc::mad_lo_cc(a[0], bi, X[0]);
  X[1] = cc::madc_hi_cc(a[0], bi, X[1]);
  X[2] = cc::madc_lo_cc(a[2], bi, X[2]);
  X[3] = cc::madc_hi_cc(a[2], bi, X[3]);
  X[4] = cc::madc_lo_cc(a[4], bi, X[4]);
  X[5] = cc::madc_hi_cc(a[4], bi, X[5]);
  X[6] = cc::madc_lo_cc(a[6], bi, X[6]);
  X[7] = cc::madc_hi_cc(a[6], bi, X[7]);
  Y[7] = cc::addc(Y[7], 0);
  redc_digit<PP>(X, Y);
}
}  // namespace detail

// ---- split form: full 512-bit product (Karatsuba) or square, then a separate Montgomery reduction -------
// The word-serial product above spends 64 limb MACs on a*b and 72 on the reduction.  One Karatsuba level
// brings a*b down to 3 x 16 = 48 MACs (a dedicated square to 36) at the price of ~70 more IADD3/LOP3.
// MEASURED ON B200 (tools/modmul_variants.py, profiles/r01j_modmul_variants.json): word-serial 67.4 G
// products/s, split 65.4 G/s, square+add 71.5 G/s; whole step 51.6 ms (word-serial) vs 56.3 ms (split, whose
// larger live set also costs 10-16 registers per kernel).  ptxas already fuses the word-serial form into 123
// IMAD.WIDE per product and the extra ALU work is not free next to them, so the library uses the word-serial
// form; the split form stays as a tested alternative behind -DEON_FP_SPLIT.
namespace detail {
// r[0..8) = a[0..4) * b[0..4).  Products at even limb positions accumulate in E, at odd positions in O
// (O[k] has weight 2^(32(k+1))), one carry chain per (row, parity); r = E + 2^32 O.
EON_HD void mul4(u32 r[8], const u32 a[4], const u32 b[4]) {
  u32 E[8], O[7];
  E[0] = cc::mul_lo(a[0], b[0]);  E[1] = cc::mul_hi(a[0], b[0]);
  E[2] = cc::mul_lo(a[2], b[0]);  E[3] = cc::mul_hi(a[2], b[0]);
  O[0] = cc::mul_lo(a[1], b[0]);  O[1] = cc::mul_hi(a[1], b[0]);
  O[2] = cc::mul_lo(a[3], b[0]);  O[3] = cc::mul_hi(a[3], b[0]);
  // row 1: a0 b1 @1, a2 b1 @3 (odd); a1 b1 @2, a3 b1 @4 (even)
  O[0] = cc::mad_lo_cc(a[0], b[1], O[0]);   O[1] = cc::madc_hi_cc(a[0], b[1], O[1]);
  O[2] = cc::madc_lo_cc(a[2], b[1], O[2]);  O[3] = cc::madc_hi_cc(a[2], b[1], O[3]);
  O[4] = cc::addc(0, 0);
  E[2] = cc::mad_lo_cc(a[1], b[1], E[2]);   E[3] = cc::madc_hi_cc(a[1], b[1], E[3]);
  E[4] = cc::madc_lo_cc(a[3], b[1], 0);     E[5] = cc::madc_hi(a[3], b[1], 0);
  // row 2: a0 b2 @2, a2 b2 @4 (even); a1 b2 @3, a3 b2 @5 (odd)
  E[2] = cc::mad_lo_cc(a[0], b[2], E[2]);   E[3] = cc::madc_hi_cc(a[0], b[2], E[3]);
  E[4] = cc::madc_lo_cc(a[2], b[2], E[4]);  E[5] = cc::madc_hi_cc(a[2], b[2], E[5]);
  E[6] = cc::addc(0, 0);
  O[2] = cc::mad_lo_cc(a[1], b[2], O[2]);   O[3] = cc::madc_hi_cc(a[1], b[2], O[3]);
  O[4] = cc::madc_lo_cc(a[3], b[2], O[4]);  O[5] = cc::madc_hi(a[3], b[2], 0);
  // row 3: a0 b3 @3, a2 b3 @5 (odd); a1 b3 @4, a3 b3 @6 (even)
  O[2] = cc::mad_lo_cc(a[0], b[3], O[2]);   O[3] = cc::madc_hi_cc(a[0], b[3], O[3]);
  O[4] = cc::madc_lo_cc(a[2], b[3], O[4]);  O[5] = cc::madc_hi_cc(a[2], b[3], O[5]);
  O[6] = cc::addc(0, 0);
  E[4] = cc::mad_lo_cc(a[1], b[3], E[4]);   E[5] = cc::madc_hi_cc(a[1], b[3], E[5]);
  E[6] = cc::madc_lo_cc(a[3], b[3], E[6]);  E[7] = cc::madc_hi(a[3], b[3], 0);
  r[0] = E[0];
  r[1] = cc::add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) r[k] = cc::addc_cc(E[k], O[k - 1]);
  r[7] = cc::addc(E[7], O[6]);
}

// r[0..8) = a[0..4)^2: the six off-diagonal products once, doubled, plus the four squares.
EON_HD void sqr4(u32 r[8], const u32 a[4]) {
  u32 O[6], E2, E3, E4, E5, f[8];
  O[0] = cc::mul_lo(a[0], a[1]);  O[1] = cc::mul_hi(a[0], a[1]);   // @1
  O[2] = cc::mul_lo(a[0], a[3]);  O[3] = cc::mul_hi(a[0], a[3]);   // @3
  O[4] = cc::mul_lo(a[2], a[3]);  O[5] = cc::mul_hi(a[2], a[3]);   // @5
  O[2] = cc::mad_lo_cc(a[1], a[2], O[2]);  O[3] = cc::madc_hi_cc(a[1], a[2], O[3]);  // @3
  O[4] = cc::addc_cc(O[4], 0);  O[5] = cc::addc(O[5], 0);
  E2 = cc::mul_lo(a[0], a[2]);  E3 = cc::mul_hi(a[0], a[2]);       // @2
  E4 = cc::mul_lo(a[1], a[3]);  E5 = cc::mul_hi(a[1], a[3]);       // @4
  // f = off-diagonal sum (limb 0 is empty), < 2^225
  f[1] = O[0];
  f[2] = cc::add_cc(E2, O[1]);
  f[3] = cc::addc_cc(E3, O[2]);
  f[4] = cc::addc_cc(E4, O[3]);
  f[5] = cc::addc_cc(E5, O[4]);
  f[6] = cc::addc_cc(0, O[5]);
  f[7] = cc::addc(0, 0);
  // 2 f + squares: the squares sit on disjoint limb pairs, so one mad chain adds them all
  f[7] = (f[7] << 1) | (f[6] >> 31);
#pragma unroll
  for (int k = 6; k >= 2; k--) f[k] = (f[k] << 1) | (f[k - 1] >> 31);
  f[1] <<= 1;
  r[0] = cc::mul_lo(a[0], a[0]);
  r[1] = cc::mad_hi_cc(a[0], a[0], f[1]);
  r[2] = cc::madc_lo_cc(a[1], a[1], f[2]);  r[3] = cc::madc_hi_cc(a[1], a[1], f[3]);
  r[4] = cc::madc_lo_cc(a[2], a[2], f[4]);  r[5] = cc::madc_hi_cc(a[2], a[2], f[5]);
  r[6] = cc::madc_lo_cc(a[3], a[3], f[6]);  r[7] = cc::madc_hi(a[3], a[3], f[7]);
}

// T[0..16) = a * b.  a = a0 + 2^128 a1, b = b0 + 2^128 b1:
//   a b = a0 b0 + 2^256 a1 b1 + 2^128 (a0 b0 + a1 b1 + (a0 - a1)(b1 - b0))
EON_HD void mul8_wide(u32 T[16], const u32 a[8], const u32 b[8]) {
  mul4(T, a, b);
  mul4(T + 8, a + 4, b + 4);
  u32 da[4], db[4], m[8], z[9];
  da[0] = cc::sub_cc(a[0], a[4]);
#pragma unroll
  for (int i = 1; i < 4; i++) da[i] = cc::subc_cc(a[i], a[4 + i]);
  const u32 sa = cc::subc(0, 0);  // all-ones: a0 < a1
  db[0] = cc::sub_cc(b[4], b[0]);
#pragma unroll
  for (int i = 1; i < 4; i++) db[i] = cc::subc_cc(b[4 + i], b[i]);
  const u32 sb = cc::subc(0, 0);
  // |x| = (x ^ s) - s
  da[0] = cc::sub_cc(da[0] ^ sa, sa);
  da[1] = cc::subc_cc(da[1] ^ sa, sa);
  da[2] = cc::subc_cc(da[2] ^ sa, sa);
  da[3] = cc::subc(da[3] ^ sa, sa);
  db[0] = cc::sub_cc(db[0] ^ sb, sb);
  db[1] = cc::subc_cc(db[1] ^ sb, sb);
  db[2] = cc::subc_cc(db[2] ^ sb, sb);
  db[3] = cc::subc(db[3] ^ sb, sb);
  mul4(m, da, db);
  const u32 neg = sa ^ sb;  // all-ones: the cross term enters with a minus sign
  z[0] = cc::add_cc(T[0], T[8]);
#pragma unroll
  for (int i = 1; i < 8; i++) z[i] = cc::addc_cc(T[i], T[8 + i]);
  z[8] = cc::addc(0, 0);
  // z += neg ? -m : m   (two's complement over 9 limbs; the true value is non-negative)
  (void)cc::add_cc(neg, neg);  // carry flag = neg & 1
#pragma unroll
  for (int i = 0; i < 8; i++) z[i] = cc::addc_cc(z[i], m[i] ^ neg);
  z[8] = cc::addc(z[8], neg);
  T[4] = cc::add_cc(T[4], z[0]);
#pragma unroll
  for (int i = 1; i < 9; i++) T[4 + i] = cc::addc_cc(T[4 + i], z[i]);
  T[13] = cc::addc_cc(T[13], 0);
  T[14] = cc::addc_cc(T[14], 0);
  T[15] = cc::addc(T[15], 0);
}

// T[0..16) = a^2 = a0^2 + 2^256 a1^2 + 2^129 a0 a1
EON_HD void sqr8_wide(u32 T[16], const u32 a[8]) {
  sqr4(T, a);
  sqr4(T + 8, a + 4);
  u32 m[8];
  mul4(m, a, a + 4);
  T[4] = cc::add_cc(T[4], m[0] << 1);
#pragma unroll
  for (int i = 1; i < 8; i++) T[4 + i] = cc::addc_cc(T[4 + i], (m[i] << 1) | (m[i - 1] >> 31));
  T[12] = cc::addc_cc(T[12], m[7] >> 31);
  T[13] = cc::addc_cc(T[13], 0);
  T[14] = cc::addc_cc(T[14], 0);
  T[15] = cc::addc(T[15], 0);
}

// One reduction digit of the split form.  X = previous odd array (the even array after the 2^-32 shift),
// Y = previous even array with Y[0] == 0; tin = the limb of T that enters the 9-limb window at this digit
// (weight 2^(32*7) in the new frame).  The window never overflows: after digit i it holds
// (T mod 2^(32(i+8)) + sum m_k p 2^(32k)) / 2^(32 i) < 2^(32*9).
template <class PP>
EON_HD void redc_next(u32 X[8], u32 Y[8], u32 tin) {
  X[0] = cc::add_cc(X[0], Y[1]);
  const u32 m = X[0] * PP::INV;  // mul.lo leaves the carry flag alone
  Y[0] = cc::madc_lo_cc(PP::mod(1), m, Y[2]);
  Y[1] = cc::madc_hi_cc(PP::mod(1), m, Y[3]);
  Y[2] = cc::madc_lo_cc(PP::mod(3), m, Y[4]);
  Y[3] = cc::madc_hi_cc(PP::mod(3), m, Y[5]);
  Y[4] = cc::madc_lo_cc(PP::mod(5), m, Y[6]);
  Y[5] = cc::madc_hi_cc(PP::mod(5), m, Y[7]);
  Y[6] = cc::madc_lo_cc(PP::mod(7), m, tin);
  Y[7] = cc::madc_hi(PP::mod(7), m, 0);
  X[0] = cc::mad_lo_cc(PP::mod(0), m, X[0]);
  X[1] = cc::madc_hi_cc(PP::mod(0), m, X[1]);
  X[2] = cc::madc_lo_cc(PP::mod(2), m, X[2]);
  X[3] = cc::madc_hi_cc(PP::mod(2), m, X[3]);
  X[4] = cc::madc_lo_cc(PP::mod(4), m, X[4]);
  X[5] = cc::madc_hi_cc(PP::mod(4), m, X[5]);
  X[6] = cc::madc_lo_cc(PP::mod(6), m, X[6]);
  X[7] = cc::madc_hi_cc(PP::mod(6), m, X[7]);
  Y[7] = cc::addc(Y[7], 0);
}

// r = T * 2^-256 mod p in [0, 2p) for T < p * 2^256
template <class PP>
EON_HD void redc_wide(u32 r[8], const u32 T[16]) {
  u32 ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { ev[i] = T[i]; od[i] = 0; }
  redc_digit<PP>(ev, od);
  redc_next<PP>(od, ev, T[8]);
  redc_next<PP>(ev, od, T[9]);
  redc_next<PP>(od, ev, T[10]);
  redc_next<PP>(ev, od, T[11]);
  redc_next<PP>(od, ev, T[12]);
  redc_next<PP>(ev, od, T[13]);
  redc_next<PP>(od, ev, T[14]);
  r[0] = cc::add_cc(ev[0], od[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = cc::addc_cc(ev[i], od[i + 1]);
  r[7] = cc::addc(ev[7], T[15]);
}
}  // namespace detail

// r = a*b*2^-256 mod p in [0, 2p), word-serial form  (a < p; b any 256-bit value)
template <class PP>
EON_HD void fp_mul_lazy_cios(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 ev[8], od[8];
  detail::mul_first<PP>(ev, od, a, b[0]);
  detail::mul_next<PP>(od, ev, a, b[1]);
  detail::mul_next<PP>(ev, od, a, b[2]);
  detail::mul_next<PP>(od, ev, a, b[3]);
  detail::mul_next<PP>(ev, od, a, b[4]);
  detail::mul_next<PP>(od, ev, a, b[5]);
  detail::mul_next<PP>(ev, od, a, b[6]);
  detail::mul_next<PP>(od, ev, a, b[7]);
  // now: even array = od (od[0] == 0), odd array = ev.  result = ev + od[1] + 2^32*od[2..7]
  r[0] = cc::add_cc(ev[0], od[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = cc::addc_cc(ev[i], od[i + 1]);
  r[7] = cc::addc(ev[7], 0);
}

// r = a*b*2^-256 mod p in [0, 2p), split form (same value, same range as the word-serial form)
template <class PP>
EON_HD void fp_mul_lazy_split(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 T[16];
  detail::mul8_wide(T, a, b);
  detail::redc_wide<PP>(r, T);
}

// r = a*a*2^-256 mod p in [0, 2p), dedicated square + separate reduction
template <class PP>
EON_HD void fp_sqr_lazy_split(u32 r[8], const u32 a[8]) {
  u32 T[16];
  detail::sqr8_wide(T, a);
  detail::redc_wide<PP>(r, T);
}

// r = a*b*2^-256 mod p in [0, 2p)  (a < p; b any 256-bit value)
template <class PP>
EON_HD void fp_mul_lazy(u32 r[8], const u32 a[8], const u32 b[8]) {
#if defined(EON_FP_SPLIT)
  fp_mul_lazy_split<PP>(r, a, b);
#else
  fp_mul_lazy_cios<PP>(r, a, b);
#endif
}

// r = a*a*2^-256 mod p in [0, 2p)  (a < p)
template <class PP>
EON_HD void fp_sqr_lazy(u32 r[8], const u32 a[8]) {
#if defined(EON_FP_SPLIT) || defined(EON_FP_SQR_SPLIT)
  fp_sqr_lazy_split<PP>(r, a);
#else
  fp_mul_lazy_cios<PP>(r, a, a);
#endif
}

// Canonical Montgomery product.  Reference: monty_mul, bn254/src/helpers.rs:188-205.
template <class PP>
EON_HD Fp<PP> fp_mul(const Fp<PP>& a, const Fp<PP>& b) {
  u32 t[8];
  fp_mul_lazy<PP>(t, a.v, b.v);
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

template <class PP>
EON_HD Fp<PP> fp_sqr(const Fp<PP>& a) {
  u32 t[8];
  fp_sqr_lazy<PP>(t, a.v);
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

// The same square through the dedicated squaring (36 limb products + separate reduction, fp_sqr_lazy_split) whatever
// the build's default: 7 % faster than a product in isolation (profiles/r01m_modmul_variants.json), at the price of
// ~10 more live registers.  Used where that is free (k_tree_bwd: 18.19 -> 18.04 ms, still 92 registers); in the
// XYZZ finisher it spills (4.88 -> 5.17 ms), so fp_sqr itself stays the word-serial product
// (profiles/r03h_dedicated_square.txt).
template <class PP>
EON_HD Fp<PP> fp_sqr_dedicated(const Fp<PP>& a) {
  u32 t[8];
  fp_sqr_lazy_split<PP>(t, a.v);
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return r;
}

// Montgomery -> canonical integer limbs (a * 1 * R^-1).  Reference: as_canonical_biguint, field.rs:455-461.
template <class PP>
EON_HD void fp_from_mont(u32 r[8], const Fp<PP>& a) {
  u32 one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  u32 t[8];
  fp_mul_lazy<PP>(t, a.v, one);
  fp_final_sub<PP>(r, t);
}

// canonical integer limbs (< p) -> Montgomery
template <class PP>
EON_HD Fp<PP> fp_to_mont(const u32 a[8]) {
  Fp<PP> x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = a[i];
  return fp_mul(x, Fp<PP>::r2());
}

template <class PP>
EON_HD Fp<PP> fp_from_u64(u64 x) {
  u32 a[8] = {(u32)x, (u32)(x >> 32), 0, 0, 0, 0, 0, 0};
  return fp_to_mont<PP>(a);
}

// a^e for a 64-bit exponent
template <class PP>
EON_HD Fp<PP> fp_pow_u64(Fp<PP> a, u64 e) {
  Fp<PP> r = Fp<PP>::one();
  while (e) {
    if (e & 1) r = fp_mul(r, a);
    a = fp_sqr(a);
    e >>= 1;
  }
  return r;
}

template <class PP>
EON_HD Fp<PP> fp_inv(const Fp<PP>& a);

// a^(p-2) (Fermat): 256 squarings + ~128 products in one dependent chain.  Kept as the cross-check of fp_inv.
template <class PP>
EON_HD Fp<PP> fp_inv_fermat(const Fp<PP>& a) {
  // exponent p - 2, scanned from the top bit
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = PP::mod(i);
  e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
  Fp<PP> r = Fp<PP>::one();
  for (int i = 255; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
  }
  return r;
}

// Inverse by the binary extended Euclid, fp_inv_bgcd (the reference inverts with a binary GCD as well: try_inverse ->
// gcd_inversion, bn254/src/field.rs:385-392; the result is the unique field element either way).  Inverse of 0 is 0.
//
// Why not Fermat here: the inversions of this library sit on latency-critical single-warp paths (the top of the
// batched-affine inversion tree of every MSM round, the final to-affine of every column), where a lone warp runs
// one Montgomery product in ~1200 cycles (272 carry-chained IMADs): a^(p-2) costs ~0.23 ms per inversion, four of
// them in a row per MSM.  One iteration below is ~120 shift / add / select instructions on the 8 limbs, and
// <= 508 iterations (each removes at least one bit from u or v) replace the 384 products: ~10x shorter.
//
// Invariants (x = the input limbs read as an integer, i.e. a*R):  x1 * x = u,  x2 * x = v  (mod p), gcd(u, v) = 1.
// Every iteration picks a "target" pair and leaves it in (u, x1) -- the two pairs are symmetric, so they are simply
// swapped by masks, no branch depends on the data (lanes of a warp do not diverge inside an iteration):
//   u even            -> u /= 2
//   u odd, v even     -> swap, then as above
//   both odd          -> larger -= smaller (swap if u < v), which makes it even, then /= 2
// with x1 following along mod p.  When u reaches 1, x1 = x^-1 = a^-1 R^-1, and one product by R^3 returns the
// Montgomery form a^-1 R.
template <class PP>
EON_HD Fp<PP> fp_inv_bgcd(const Fp<PP>& a) {
  if (a.is_zero()) return a;
  u32 u[8], v[8], x1[8], x2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.v[i];
    v[i] = PP::mod(i);
    x1[i] = (i == 0) ? 1u : 0u;
    x2[i] = 0u;
  }
  for (int it = 0; it < 1024; it++) {  // bounded: <= 508 iterations for any input below p
    if (u[0] == 1u && (u[1] | u[2] | u[3] | u[4] | u[5] | u[6] | u[7]) == 0u) break;
    // borrow of u - v  ->  lt = all ones iff u < v
    u32 t = cc::sub_cc(u[0], v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t = cc::subc_cc(u[i], v[i]);
    const u32 lt = cc::subc(0u, 0u);
    (void)t;
    const u32 uo = 0u - (u[0] & 1u), vo = 0u - (v[0] & 1u);  // all ones iff odd
    const u32 both = uo & vo;
    const u32 sw = (uo & ~vo) | (both & lt);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const u32 d = (u[i] ^ v[i]) & sw, e = (x1[i] ^ x2[i]) & sw;
      u[i] ^= d;
      v[i] ^= d;
      x1[i] ^= e;
      x2[i] ^= e;
    }
    // u = (u - (both ? v : 0)) / 2
    u[0] = cc::sub_cc(u[0], v[0] & both);
#pragma unroll
    for (int i = 1; i < 8; i++) u[i] = cc::subc_cc(u[i], v[i] & both);
#pragma unroll
    for (int i = 0; i < 7; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
    u[7] >>= 1;
    // x1 = (x1 - (both ? x2 : 0)) / 2  (mod p)
    x1[0] = cc::sub_cc(x1[0], x2[0] & both);
#pragma unroll
    for (int i = 1; i < 8; i++) x1[i] = cc::subc_cc(x1[i], x2[i] & both);
    const u32 neg = cc::subc(0u, 0u);  // all ones iff the difference went below zero: add p back
    x1[0] = cc::add_cc(x1[0], PP::mod(0) & neg);
#pragma unroll
    for (int i = 1; i < 8; i++) x1[i] = cc::addc_cc(x1[i], PP::mod(i) & neg);
    const u32 odd = 0u - (x1[0] & 1u);  // odd: add p (odd) first; x1 + p < 2p < 2^255, no carry out
    x1[0] = cc::add_cc(x1[0], PP::mod(0) & odd);
#pragma unroll
    for (int i = 1; i < 8; i++) x1[i] = cc::addc_cc(x1[i], PP::mod(i) & odd);
#pragma unroll
    for (int i = 0; i < 7; i++) x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31);
    x1[7] >>= 1;
  }
  Fp<PP> y;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = x1[i];
  const Fp<PP> r3 = fp_mul(Fp<PP>::r2(), Fp<PP>::r2());  // R^2 * R^2 / R = R^3
  return fp_mul(y, r3);
}

// ---- inverse by "safegcd" divsteps (Bernstein-Yang 2019), the form every latency-critical inversion now uses -------
// The binary GCD above moves one bit per iteration and touches all 256 bits of four numbers each time (~120
// instructions, <= 508 iterations: ~65 us for a lone warp).  Divsteps decide 30 steps at a time from the LOW 30 bits
// of (f, g) alone -- a chain of single-word operations -- and collect them in a 2x2 matrix of 31-bit entries that
// is applied to the full numbers once per batch: 20 batches x (30 word-steps + two 9-limb matrix applications whose
// products are independent) ~ 4x fewer instructions and far shorter dependent chains.  The formulation is the
// constant-time one (no data-dependent branch: the lanes of a warp stay together) with zeta = -(delta + 1/2):
// 590 divsteps suffice for any 256-bit modulus, 20 x 30 = 600 are done.
// Numbers are signed, 9 limbs of 30 bits (value = sum v[i] 2^(30 i)); d, e stay in (-2p, p).
// Invariants: d * x = f, e * x = g (mod p); at the end g = 0, f = +-1, so x^-1 = +-d.  Inverse of 0 is 0.
namespace sgcd {
struct S30 {
  int v[9];
};
template <class PP>
EON_HD void modulus30(S30& m) {
  u32 w[9];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = PP::mod(i);
  w[8] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, j = bit >> 5, sft = bit & 31;
    const u64 two = (u64)w[j] | ((j + 1 < 9) ? ((u64)w[j + 1] << 32) : 0ull);
    m.v[i] = (int)((two >> sft) & 0x3fffffffu);
  }
}
EON_HD void to30(S30& r, const u32 (&a)[8]) {
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, j = bit >> 5, sft = bit & 31;
    const u64 lo = (j < 8) ? a[j] : 0u;
    const u64 hi = (j + 1 < 8) ? a[j + 1] : 0u;
    r.v[i] = (int)(((lo | (hi << 32)) >> sft) & 0x3fffffffu);
  }
}
// 30 divsteps on the low words; t = (u, v, q, r) with t * (f, g) = 2^30 * (f', g')
EON_HD int divsteps30(int zeta, u32 f0, u32 g0, int (&t)[4]) {
  u32 u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    u32 c1 = (u32)(zeta >> 31);  // all ones iff zeta < 0
    const u32 c2 = 0u - (g & 1u);
    const u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // conditionally negated f, u, v
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;
    zeta = (int)(((u32)zeta ^ c1) - 1u);
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t[0] = (int)u;
  t[1] = (int)v;
  t[2] = (int)q;
  t[3] = (int)r;
  return zeta;
}
// (f, g) <- t * (f, g) / 2^30 (exact)
EON_HD void update_fg(S30& f, S30& g, const int (&t)[4]) {
  const long long u = t[0], v = t[1], q = t[2], r = t[3];
  long long cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cf += u * f.v[i] + v * g.v[i];
    cg += q * f.v[i] + r * g.v[i];
    f.v[i - 1] = (int)((u32)cf & 0x3fffffffu);
    g.v[i - 1] = (int)((u32)cg & 0x3fffffffu);
    cf >>= 30;
    cg >>= 30;
  }
  f.v[8] = (int)cf;
  g.v[8] = (int)cg;
}
// (d, e) <- t * (d, e) / 2^30 mod p: a multiple of p is added so that the low 30 bits vanish
EON_HD void update_de(S30& d, S30& e, const int (&t)[4], const S30& m, u32 pinv30) {
  const int u = t[0], v = t[1], q = t[2], r = t[3];
  const int sd = d.v[8] >> 31, se = e.v[8] >> 31;  // sign masks of d, e
  int md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  long long cd = (long long)u * d.v[0] + (long long)v * e.v[0];
  long long ce = (long long)q * d.v[0] + (long long)r * e.v[0];
  md -= (int)((pinv30 * (u32)cd + (u32)md) & 0x3fffffffu);
  me -= (int)((pinv30 * (u32)ce + (u32)me) & 0x3fffffffu);
  cd += (long long)m.v[0] * md;
  ce += (long long)m.v[0] * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cd += (long long)u * d.v[i] + (long long)v * e.v[i] + (long long)m.v[i] * md;
    ce += (long long)q * d.v[i] + (long long)r * e.v[i] + (long long)m.v[i] * me;
    d.v[i - 1] = (int)((u32)cd & 0x3fffffffu);
    e.v[i - 1] = (int)((u32)ce & 0x3fffffffu);
    cd >>= 30;
    ce >>= 30;
  }
  d.v[8] = (int)cd;
  e.v[8] = (int)ce;
}
// d in (-2p, p) -> sign * d mod p in [0, p), sign = -1 iff fsign < 0; then back to 8 x 32 bits
EON_HD void normalize(u32 (&out)[8], S30 r, int fsign, const S30& m) {
  int add = r.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & add;
  const int neg = fsign >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ neg) - neg;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= 0x3fffffff;
  }
  add = r.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & add;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= 0x3fffffff;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int bit = 32 * j, i = bit / 30, sft = bit % 30;  // bits [32 j, 32 j + 32) of the 30-bit limb string
    u64 acc = (u64)(u32)r.v[i] >> sft;
    acc |= (u64)(u32)r.v[i + 1] << (30 - sft);
    if (i + 2 < 9) acc |= (u64)(u32)r.v[i + 2] << (60 - sft);
    out[j] = (u32)acc;
  }
}
}  // namespace sgcd

template <class PP>
EON_HD Fp<PP> fp_inv_safegcd(const Fp<PP>& a) {
  sgcd::S30 m, d, e, f, g;
  sgcd::modulus30<PP>(m);
  const u32 pinv30 = (0u - PP::INV) & 0x3fffffffu;  // p^-1 mod 2^30 (INV = -p^-1 mod 2^32)
#pragma unroll
  for (int i = 0; i < 9; i++) {
    d.v[i] = 0;
    e.v[i] = (i == 0) ? 1 : 0;
  }
  f = m;
  sgcd::to30(g, a.v);
  int zeta = -1;
  for (int it = 0; it < 20; it++) {
    int t[4];
    zeta = sgcd::divsteps30(zeta, (u32)f.v[0], (u32)g.v[0], t);
    sgcd::update_de(d, e, t, m, pinv30);
    sgcd::update_fg(f, g, t);
  }
  Fp<PP> y;
  sgcd::normalize(y.v, d, f.v[8], m);
  const Fp<PP> r3 = fp_mul(Fp<PP>::r2(), Fp<PP>::r2());  // R^2 * R^2 / R = R^3
  return fp_mul(y, r3);
}

// The inversion of the library: safegcd.  MEASURED on B200 (profiles/r03c_inversion.txt): k_tree_top (<= 2048
// independent inversions, one warp per SM) against the binary GCD above, kept as fp_inv_bgcd for the cross-check
// (tests/test_host_arith.py: both against pow(x, -1, p) and against each other, bit for bit).
template <class PP>
EON_HD Fp<PP> fp_inv(const Fp<PP>& a) {
#if defined(EON_INV_BGCD)
  return fp_inv_bgcd(a);
#else
  return fp_inv_safegcd(a);
#endif
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace eon
