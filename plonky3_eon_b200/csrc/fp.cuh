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

// r = a*b*2^-256 mod p in [0, 2p)  (a < p; b any 256-bit value)
template <class PP>
EON_HD void fp_mul_lazy(u32 r[8], const u32 a[8], const u32 b[8]) {
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
EON_HD Fp<PP> fp_sqr(const Fp<PP>& a) { return fp_mul(a, a); }

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

// a^(p-2) (Fermat).  Inverse of 0 is 0.  Rare on this path (shift^-1, to-affine).
template <class PP>
EON_HD Fp<PP> fp_inv(const Fp<PP>& a) {
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

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace eon
