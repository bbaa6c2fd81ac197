/*
 * oracle.c — CPU restatement (C port) of the reference's BN254 KZG hot path.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing under plonky3_eon_b200/ links or calls this.
 * It is (a) the fast checker for sizes the Python big-int oracle cannot reach and (b) the
 * `cpu_baseline` / `--impl reference` arm of bench.py ("kind": "port" — the reference itself is
 * Rust + the un-vendored halo2curves 0.9 crate and cannot be built in this image).
 * Parity of this file is pinned by tests/test_oracle_c.py against the Python oracle, which is
 * pinned against the reference's known-answer tests.
 *
 * Restated reference code:
 *   Fr / Fq Montgomery product   bn254/src/helpers.rs:75-205 (mul_small, mul_small_and_acc,
 *                                interleaved_monty_reduction, monty_mul) — 4 x u64, mu = +P^-1
 *   add / sub                    bn254/src/field.rs:464-508
 *   DFT family                   dft/src/traits.rs:83-249 defaults over a radix-2 DIT network
 *                                (dft/src/radix_2_dit.rs:64-122), cache-blocked and rayon-style
 *                                parallel like Radix2DitParallel (radix_2_dit_parallel.rs:148-228)
 *   coset_shift_cols / divide    dft/src/util.rs:15-36
 *   MSM                          G1::multi_exp, bn254/src/curve.rs:158-180 -> window Pippenger in
 *                                the style of halo2curves msm (signed digits, per-thread chunks)
 *   KZG                          kzg/src/util.rs:37-40,100-111; kzg/src/pcs.rs:223-265
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct { u64 v[4]; } fe;

typedef struct {
  u64 mod[4];
  u64 mu;     /* +mod^-1 mod 2^64 (the reduction subtracts) */
  u64 one[4]; /* R mod p */
  u64 r2[4];  /* R^2 mod p */
} field_t;

static const field_t FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x3d1e0a6c10000001ULL,
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

static const field_t FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x78d8d1eeb5a0c4a9ULL, /* placeholder, fixed up in oc_init() */
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};

static field_t FQm; /* FQ with mu computed at init */
static int g_init = 0;

static u64 inv64(u64 a) { /* a^-1 mod 2^64, a odd (Newton) */
  u64 x = a;
  for (int i = 0; i < 6; i++) x *= 2 - a * x;
  return x;
}

void oc_init(void) {
  if (g_init) return;
  FQm = FQ;
  FQm.mu = inv64(FQ.mod[0]);
  g_init = 1;
}

/* ---- helpers.rs:75-127 ---------------------------------------------------------------- */
/* (out0, out[4]) = lhs * rhs_limb (+ add) as 5 limbs */
static inline u64 mul_small_acc(const u64 lhs[4], u64 rhs, const u64 add[4], u64 out[4]) {
  u128 acc = (u128)lhs[0] * rhs + (add ? add[0] : 0);
  u64 out0 = (u64)acc;
  acc >>= 64;
  for (int i = 1; i < 4; i++) {
    acc += (u128)lhs[i] * rhs + (add ? add[i] : 0);
    out[i - 1] = (u64)acc;
    acc >>= 64;
  }
  out[3] = (u64)acc;
  return out0;
}

/* helpers.rs:168-179 */
static inline void imr(const field_t* F, u64 acc0, const u64 acc[4], u64 res[4]) {
  u64 t = acc0 * F->mu;
  u64 u[4];
  (void)mul_small_acc(F->mod, t, NULL, u);
  u64 borrow = 0;
  u64 sub[4];
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)acc[i] - u[i] - borrow;
    sub[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  if (borrow) {
    u64 c = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)sub[i] + F->mod[i] + c;
      res[i] = (u64)s;
      c = (u64)(s >> 64);
    }
  } else {
    memcpy(res, sub, 32);
  }
}

/* helpers.rs:188-205 — literal restatement (4 rounds of multiply-accumulate + subtracting IMR) */
static inline void f_mul_ref(const field_t* F, fe* r, const fe* a, const fe* b) {
  u64 acc[4], res[4];
  u64 a0 = mul_small_acc(a->v, b->v[0], NULL, acc);
  imr(F, a0, acc, res);
  for (int i = 1; i < 4; i++) {
    a0 = mul_small_acc(a->v, b->v[i], res, acc);
    imr(F, a0, acc, res);
  }
  memcpy(r->v, res, 32);
}

/* Same product (the canonical value a*b*2^-256 mod p is unique), computed with the adding form of
 * the interleaved reduction (m = t0 * (-p^-1), t += m*p) fully unrolled so gcc emits mulx/adc
 * chains: this is what keeps the CPU baseline honest (~2.5x faster than f_mul_ref here).
 * Valid because both moduli are < 2^254 (no carry out of the top limb). */
static inline void f_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
  const u64 p0 = F->mod[0], p1 = F->mod[1], p2 = F->mod[2], p3 = F->mod[3];
  const u64 ninv = (u64)0 - F->mu;
  u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define EON_ROUND(bi)                                           \
  do {                                                          \
    u128 x = (u128)a->v[0] * (bi) + t0;                         \
    u64 lo = (u64)x, A = (u64)(x >> 64);                        \
    u64 m = lo * ninv;                                          \
    u128 y = (u128)m * p0 + lo;                                 \
    u64 Cc = (u64)(y >> 64);                                    \
    x = (u128)a->v[1] * (bi) + t1 + A;                          \
    A = (u64)(x >> 64);                                         \
    y = (u128)m * p1 + (u64)x + Cc;                             \
    t0 = (u64)y; Cc = (u64)(y >> 64);                           \
    x = (u128)a->v[2] * (bi) + t2 + A;                          \
    A = (u64)(x >> 64);                                         \
    y = (u128)m * p2 + (u64)x + Cc;                             \
    t1 = (u64)y; Cc = (u64)(y >> 64);                           \
    x = (u128)a->v[3] * (bi) + t3 + A;                          \
    A = (u64)(x >> 64);                                         \
    y = (u128)m * p3 + (u64)x + Cc;                             \
    t2 = (u64)y; Cc = (u64)(y >> 64);                           \
    t3 = Cc + A;                                                \
  } while (0)
  EON_ROUND(b->v[0]);
  EON_ROUND(b->v[1]);
  EON_ROUND(b->v[2]);
  EON_ROUND(b->v[3]);
#undef EON_ROUND
  /* t < 2p: one conditional subtraction */
  u64 s0, s1, s2, s3, bw;
  u128 d = (u128)t0 - p0;
  s0 = (u64)d; bw = (u64)(d >> 64) & 1;
  d = (u128)t1 - p1 - bw;
  s1 = (u64)d; bw = (u64)(d >> 64) & 1;
  d = (u128)t2 - p2 - bw;
  s2 = (u64)d; bw = (u64)(d >> 64) & 1;
  d = (u128)t3 - p3 - bw;
  s3 = (u64)d; bw = (u64)(d >> 64) & 1;
  r->v[0] = bw ? t0 : s0;
  r->v[1] = bw ? t1 : s1;
  r->v[2] = bw ? t2 : s2;
  r->v[3] = bw ? t3 : s3;
}

static inline int ge_mod(const field_t* F, const u64 a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > F->mod[i]) return 1;
    if (a[i] < F->mod[i]) return 0;
  }
  return 1;
}

/* field.rs:464-485 */
static inline void f_add(const field_t* F, fe* r, const fe* a, const fe* b) {
  u64 s[4], c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a->v[i] + b->v[i] + c;
    s[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  if (ge_mod(F, s)) {
    u64 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)s[i] - F->mod[i] - bw;
      s[i] = (u64)d;
      bw = (u64)(d >> 64) & 1;
    }
  }
  memcpy(r->v, s, 32);
}

/* field.rs:487-508 */
static inline void f_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
  u64 s[4], bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->v[i] - b->v[i] - bw;
    s[i] = (u64)d;
    bw = (u64)(d >> 64) & 1;
  }
  if (bw) {
    u64 c = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)s[i] + F->mod[i] + c;
      s[i] = (u64)t;
      c = (u64)(t >> 64);
    }
  }
  memcpy(r->v, s, 32);
}

static inline int f_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int f_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }
static inline void f_one(const field_t* F, fe* r) { memcpy(r->v, F->one, 32); }
static inline void f_zero(fe* r) { memset(r, 0, 32); }
static inline void f_neg(const field_t* F, fe* r, const fe* a) {
  fe z;
  f_zero(&z);
  f_sub(F, r, &z, a);
}
static inline void f_sqr(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static void f_pow_u64(const field_t* F, fe* r, const fe* a, u64 e) {
  fe acc, base = *a;
  f_one(F, &acc);
  while (e) {
    if (e & 1) f_mul(F, &acc, &acc, &base);
    f_sqr(F, &base, &base);
    e >>= 1;
  }
  *r = acc;
}
static void f_inv(const field_t* F, fe* r, const fe* a) { /* a^(p-2) */
  u64 e[4];
  memcpy(e, F->mod, 32);
  e[0] -= 2;
  fe acc;
  f_one(F, &acc);
  for (int i = 255; i >= 0; i--) {
    f_sqr(F, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
  }
  *r = acc;
}
static void f_from_mont(const field_t* F, u64 out[4], const fe* a) {
  fe one = {{1, 0, 0, 0}}, r;
  f_mul(F, &r, a, &one);
  memcpy(out, r.v, 32);
}
static void f_from_u64(const field_t* F, fe* r, u64 x) {
  fe t = {{x, 0, 0, 0}}, r2;
  memcpy(r2.v, F->r2, 32);
  f_mul(F, r, &r2, &t); /* Fr::new: monty_mul(R^2, [v,0,0,0]), field.rs:110-116 */
}

/* exported element-wise ops for tests: which 0 = Fr, 1 = Fq; op 0 mul 1 add 2 sub 3 inv 4 mul (literal) */
void oc_field_op(int which, int op, const u64* a, const u64* b, u64* r, size_t n) {
  oc_init();
  const field_t* F = which ? &FQm : &FR;
  for (size_t i = 0; i < n; i++) {
    const fe* x = (const fe*)(a + 4 * i);
    const fe* y = (const fe*)(b + 4 * i);
    fe* z = (fe*)(r + 4 * i);
    if (op == 0) f_mul(F, z, x, y);
    else if (op == 4) f_mul_ref(F, z, x, y);
    else if (op == 1) f_add(F, z, x, y);
    else if (op == 2) f_sub(F, z, x, y);
    else f_inv(F, z, x);
  }
}

/* ======================================================================================== */
/* DFT                                                                                       */
/* ======================================================================================== */
static const u64 OMEGA28[4] = {0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL,
                               0x1860ef942963f9e7ULL}; /* field.rs:556-561 */

static void two_adic_generator(fe* r, unsigned bits) { /* field.rs:567-573 */
  memcpy(r->v, OMEGA28, 32);
  for (unsigned i = bits; i < 28; i++) f_sqr(&FR, r, r);
}

static inline size_t bitrev(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

/* reverse_matrix_index_bits, matrix/src/util.rs:36-56 */
static void reverse_rows(fe* m, unsigned log_h, size_t w) {
  size_t h = (size_t)1 << log_h;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < h; i++) {
    size_t j = bitrev(i, log_h);
    if (i < j) {
      fe* a = m + i * w;
      fe* b = m + j * w;
      for (size_t c = 0; c < w; c++) {
        fe t = a[c];
        a[c] = b[c];
        b[c] = t;
      }
    }
  }
}

/* in-place DFT of every column, natural in/out: bit-reverse then DIT layers (radix_2_dit.rs:64-122),
 * with the layers grouped into cache-sized blocks of rows processed by independent threads
 * (the blocking idea of Radix2DitParallel, radix_2_dit_parallel.rs:296-421). */
static void dft_inplace(fe* m, unsigned log_h, size_t w) {
  if (log_h == 0 || w == 0) return;
  size_t h = (size_t)1 << log_h;
  /* twiddles: omega^j, j < h/2 */
  fe* tw = (fe*)malloc((h / 2) * sizeof(fe));
  fe g;
  two_adic_generator(&g, log_h);
  {
    /* parallel fill: each thread starts from g^start */
#pragma omp parallel
    {
#ifdef _OPENMP
      int nt = omp_get_num_threads(), id = omp_get_thread_num();
#else
      int nt = 1, id = 0;
#endif
      size_t per = (h / 2 + nt - 1) / nt, s = per * id, e = s + per;
      if (e > h / 2) e = h / 2;
      if (s < e) {
        fe cur;
        f_pow_u64(&FR, &cur, &g, s);
        for (size_t j = s; j < e; j++) {
          tw[j] = cur;
          f_mul(&FR, &cur, &cur, &g);
        }
      }
    }
  }
  reverse_rows(m, log_h, w);
  /* row-block size so that 2^r rows * w * 32 B ~ 256 KiB */
  unsigned rmax = 1;
  while (rmax < 12 && ((size_t)1 << (rmax + 1)) * w * 32 <= (256u << 10)) rmax++;
  for (unsigned l0 = 0; l0 < log_h;) {
    unsigned r = log_h - l0 < rmax ? log_h - l0 : rmax;
    size_t ntiles = h >> r; /* (hi, lo) pairs */
    size_t lo_cnt = (size_t)1 << l0;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t tile = 0; tile < ntiles; tile++) {
      size_t lo = tile & (lo_cnt - 1), hi = tile >> l0;
      size_t base = (hi << (l0 + r)) + lo;
      for (unsigned t = 0; t < r; t++) {
        unsigned l = l0 + t;
        size_t half = (size_t)1 << t;
        size_t tw_stride = h >> (l + 1);
        for (size_t b = 0; b < ((size_t)1 << (r - 1)); b++) {
          size_t i0 = ((b >> t) << (t + 1)) | (b & (half - 1));
          size_t row0 = base + (i0 << l0), row1 = row0 + (half << l0);
          size_t j = row0 & (((size_t)1 << l) - 1);
          const fe* twj = &tw[j * tw_stride];
          fe* a = m + row0 * w;
          fe* bb = m + row1 * w;
          if (j == 0) { /* twiddle 1: radix_2_dit.rs:113-116 */
            for (size_t c = 0; c < w; c++) {
              fe x = bb[c], y = a[c];
              f_add(&FR, &a[c], &y, &x);
              f_sub(&FR, &bb[c], &y, &x);
            }
          } else {
            for (size_t c = 0; c < w; c++) {
              fe x, y = a[c];
              f_mul(&FR, &x, &bb[c], twj);
              f_add(&FR, &a[c], &y, &x);
              f_sub(&FR, &bb[c], &y, &x);
            }
          }
        }
      }
    }
    l0 += r;
  }
  free(tw);
}

/* coset_shift_cols, dft/src/util.rs:28-36 */
static void coset_shift_rows(fe* m, size_t h, size_t w, const fe* shift) {
  fe one;
  f_one(&FR, &one);
  if (f_eq(shift, &one)) return; /* (the reference multiplies anyway; result identical) */
#pragma omp parallel
  {
#ifdef _OPENMP
    int nt = omp_get_num_threads(), id = omp_get_thread_num();
#else
    int nt = 1, id = 0;
#endif
    size_t per = (h + nt - 1) / nt, s = per * id, e = s + per;
    if (e > h) e = h;
    if (s < e) {
      fe wgt;
      f_pow_u64(&FR, &wgt, shift, s);
      for (size_t i = s; i < e; i++) {
        for (size_t c = 0; c < w; c++) f_mul(&FR, &m[i * w + c], &m[i * w + c], &wgt);
        f_mul(&FR, &wgt, &wgt, shift);
      }
    }
  }
}

/* idft_batch, traits.rs:111-122: dft, divide_by_height (util.rs:15-25), swap rows i <-> h-i */
static void idft_inplace(fe* m, unsigned log_h, size_t w) {
  size_t h = (size_t)1 << log_h;
  dft_inplace(m, log_h, w);
  fe hv, hinv;
  f_from_u64(&FR, &hv, (u64)h);
  f_inv(&FR, &hinv, &hv);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < h * w; i++) f_mul(&FR, &m[i], &m[i], &hinv);
#pragma omp parallel for schedule(static)
  for (size_t row = 1; row < h / 2; row++) {
    fe* a = m + row * w;
    fe* b = m + (h - row) * w;
    for (size_t c = 0; c < w; c++) {
      fe t = a[c];
      a[c] = b[c];
      b[c] = t;
    }
  }
}

void oc_dft_batch(u64* mat, unsigned log_h, size_t w) {
  oc_init();
  dft_inplace((fe*)mat, log_h, w);
}
void oc_coset_dft_batch(u64* mat, unsigned log_h, size_t w, const u64 shift[4]) { /* traits.rs:83-91 */
  oc_init();
  coset_shift_rows((fe*)mat, (size_t)1 << log_h, w, (const fe*)shift);
  dft_inplace((fe*)mat, log_h, w);
}
void oc_idft_batch(u64* mat, unsigned log_h, size_t w) {
  oc_init();
  idft_inplace((fe*)mat, log_h, w);
}
void oc_coset_idft_batch(u64* mat, unsigned log_h, size_t w, const u64 shift[4]) { /* traits.rs:144-153 */
  oc_init();
  idft_inplace((fe*)mat, log_h, w);
  fe sinv;
  f_inv(&FR, &sinv, (const fe*)shift);
  coset_shift_rows((fe*)mat, (size_t)1 << log_h, w, &sinv);
}
/* coset_lde_batch, traits.rs:226-249.  out has h << added_bits rows. */
void oc_coset_lde_batch(const u64* in, u64* out, unsigned log_h, size_t w, unsigned added_bits, const u64 shift[4]) {
  oc_init();
  size_t h = (size_t)1 << log_h;
  memcpy(out, in, h * w * 32);
  idft_inplace((fe*)out, log_h, w);
  memset(out + h * w * 4, 0, ((h << added_bits) - h) * w * 32);
  oc_coset_dft_batch(out, log_h + added_bits, w, shift);
}

/* ======================================================================================== */
/* G1: y^2 = x^3 + 3 over Fq (halo2curves bn256::G1), XYZZ accumulators                       */
/* ======================================================================================== */
typedef struct { fe x, y; } aff;            /* identity = (0,0) */
typedef struct { fe X, Y, ZZ, ZZZ; } xyzz;  /* identity <=> ZZ == 0 */

#define Q (&FQm)
static inline int aff_is_id(const aff* p) { return f_is_zero(&p->x) && f_is_zero(&p->y); }
static inline void xyzz_set_id(xyzz* p) { memset(p, 0, sizeof(*p)); }

static void xyzz_dbl_affine(xyzz* r, const aff* p) {
  if (aff_is_id(p)) { xyzz_set_id(r); return; }
  fe U, V, W, S, M, t, X2;
  f_add(Q, &U, &p->y, &p->y);
  f_sqr(Q, &V, &U);
  f_mul(Q, &W, &U, &V);
  f_mul(Q, &S, &p->x, &V);
  f_sqr(Q, &X2, &p->x);
  f_add(Q, &M, &X2, &X2);
  f_add(Q, &M, &M, &X2);
  f_sqr(Q, &r->X, &M);
  f_sub(Q, &r->X, &r->X, &S);
  f_sub(Q, &r->X, &r->X, &S);
  f_sub(Q, &t, &S, &r->X);
  f_mul(Q, &t, &M, &t);
  fe wy;
  f_mul(Q, &wy, &W, &p->y);
  f_sub(Q, &r->Y, &t, &wy);
  r->ZZ = V;
  r->ZZZ = W;
}

static void xyzz_dbl(xyzz* r, const xyzz* p) {
  if (f_is_zero(&p->ZZ)) { xyzz_set_id(r); return; }
  fe U, V, W, S, M, t, X2, wy;
  xyzz o;
  f_add(Q, &U, &p->Y, &p->Y);
  f_sqr(Q, &V, &U);
  f_mul(Q, &W, &U, &V);
  f_mul(Q, &S, &p->X, &V);
  f_sqr(Q, &X2, &p->X);
  f_add(Q, &M, &X2, &X2);
  f_add(Q, &M, &M, &X2);
  f_sqr(Q, &o.X, &M);
  f_sub(Q, &o.X, &o.X, &S);
  f_sub(Q, &o.X, &o.X, &S);
  f_sub(Q, &t, &S, &o.X);
  f_mul(Q, &t, &M, &t);
  f_mul(Q, &wy, &W, &p->Y);
  f_sub(Q, &o.Y, &t, &wy);
  f_mul(Q, &o.ZZ, &V, &p->ZZ);
  f_mul(Q, &o.ZZZ, &W, &p->ZZZ);
  *r = o;
}

static void xyzz_add_mixed(xyzz* acc, const aff* q, int negate) {
  if (aff_is_id(q)) return;
  aff qq = *q;
  if (negate) f_neg(Q, &qq.y, &q->y);
  if (f_is_zero(&acc->ZZ)) {
    acc->X = qq.x;
    acc->Y = qq.y;
    f_one(Q, &acc->ZZ);
    f_one(Q, &acc->ZZZ);
    return;
  }
  fe U2, S2, P_, R_, PP, PPP, Qv, t, X3, Y3;
  f_mul(Q, &U2, &qq.x, &acc->ZZ);
  f_mul(Q, &S2, &qq.y, &acc->ZZZ);
  f_sub(Q, &P_, &U2, &acc->X);
  f_sub(Q, &R_, &S2, &acc->Y);
  if (f_is_zero(&P_)) {
    if (f_is_zero(&R_)) xyzz_dbl_affine(acc, &qq);
    else xyzz_set_id(acc);
    return;
  }
  f_sqr(Q, &PP, &P_);
  f_mul(Q, &PPP, &P_, &PP);
  f_mul(Q, &Qv, &acc->X, &PP);
  f_sqr(Q, &X3, &R_);
  f_sub(Q, &X3, &X3, &PPP);
  f_sub(Q, &X3, &X3, &Qv);
  f_sub(Q, &X3, &X3, &Qv);
  f_sub(Q, &t, &Qv, &X3);
  f_mul(Q, &Y3, &R_, &t);
  f_mul(Q, &t, &acc->Y, &PPP);
  f_sub(Q, &Y3, &Y3, &t);
  acc->X = X3;
  acc->Y = Y3;
  f_mul(Q, &acc->ZZ, &acc->ZZ, &PP);
  f_mul(Q, &acc->ZZZ, &acc->ZZZ, &PPP);
}

static void xyzz_add(xyzz* acc, const xyzz* q) {
  if (f_is_zero(&q->ZZ)) return;
  if (f_is_zero(&acc->ZZ)) { *acc = *q; return; }
  fe U1, U2, S1, S2, P_, R_, PP, PPP, Qv, t, X3, Y3;
  f_mul(Q, &U1, &acc->X, &q->ZZ);
  f_mul(Q, &U2, &q->X, &acc->ZZ);
  f_mul(Q, &S1, &acc->Y, &q->ZZZ);
  f_mul(Q, &S2, &q->Y, &acc->ZZZ);
  f_sub(Q, &P_, &U2, &U1);
  f_sub(Q, &R_, &S2, &S1);
  if (f_is_zero(&P_)) {
    if (f_is_zero(&R_)) xyzz_dbl(acc, acc);
    else xyzz_set_id(acc);
    return;
  }
  f_sqr(Q, &PP, &P_);
  f_mul(Q, &PPP, &P_, &PP);
  f_mul(Q, &Qv, &U1, &PP);
  f_sqr(Q, &X3, &R_);
  f_sub(Q, &X3, &X3, &PPP);
  f_sub(Q, &X3, &X3, &Qv);
  f_sub(Q, &X3, &X3, &Qv);
  f_sub(Q, &t, &Qv, &X3);
  f_mul(Q, &Y3, &R_, &t);
  f_mul(Q, &t, &S1, &PPP);
  f_sub(Q, &Y3, &Y3, &t);
  acc->X = X3;
  acc->Y = Y3;
  f_mul(Q, &acc->ZZ, &acc->ZZ, &q->ZZ);
  f_mul(Q, &acc->ZZ, &acc->ZZ, &PP);
  f_mul(Q, &acc->ZZZ, &acc->ZZZ, &q->ZZZ);
  f_mul(Q, &acc->ZZZ, &acc->ZZZ, &PPP);
}

static void xyzz_to_affine(aff* r, const xyzz* p) {
  if (f_is_zero(&p->ZZ)) { memset(r, 0, sizeof(*r)); return; }
  fe z5, i, a, b;
  f_mul(Q, &z5, &p->ZZ, &p->ZZZ);
  f_inv(Q, &i, &z5);
  f_mul(Q, &a, &i, &p->ZZZ);
  f_mul(Q, &b, &i, &p->ZZ);
  f_mul(Q, &r->x, &p->X, &a);
  f_mul(Q, &r->y, &p->Y, &b);
}

/* serial signed-window Pippenger over points [0, n) with scalars at stride `ld` Fr */
static void msm_serial(const aff* pts, const u64* scalars, size_t ld, size_t n, xyzz* out) {
  xyzz_set_id(out);
  if (n == 0) return;
  unsigned c = n < 4 ? 1 : n < 32 ? 3 : (unsigned)ceil(log((double)n)); /* halo2curves window choice */
  if (c > 20) c = 20;
  unsigned W = (256 + c - 1) / c;
  if (c == 1) W = 256;
  size_t nb = (size_t)1 << (c - 1);
  /* canonical scalars */
  u64* k = (u64*)malloc(n * 5 * sizeof(u64));
  for (size_t i = 0; i < n; i++) {
    f_from_mont(&FR, k + 5 * i, (const fe*)(scalars + 4 * i * ld));
    k[5 * i + 4] = 0;
  }
  /* signed digits, window-major */
  int32_t* dig = (int32_t*)malloc(n * W * sizeof(int32_t));
  for (size_t i = 0; i < n; i++) {
    unsigned carry = 0;
    for (unsigned w = 0; w < W; w++) {
      size_t bit = (size_t)w * c;
      unsigned limb = (unsigned)(bit >> 6), off = (unsigned)(bit & 63);
      u64 raw = 0;
      if (limb < 4) {
        u128 two = (u128)k[5 * i + limb] | ((u128)k[5 * i + limb + 1] << 64);
        raw = (u64)(two >> off) & (((u64)1 << c) - 1);
      }
      raw += carry;
      if (raw >= ((u64)1 << (c - 1)) && c > 1) {
        dig[w * n + i] = (int32_t)((int64_t)raw - ((int64_t)1 << c));
        carry = 1;
      } else if (c == 1) {
        dig[w * n + i] = (int32_t)(raw & 1);
        carry = (unsigned)(raw >> 1);
      } else {
        dig[w * n + i] = (int32_t)raw;
        carry = 0;
      }
    }
  }
  xyzz* buckets = (xyzz*)malloc(nb * sizeof(xyzz));
  xyzz total;
  xyzz_set_id(&total);
  for (int w = (int)W - 1; w >= 0; w--) {
    for (unsigned d = 0; d < c; d++) xyzz_dbl(&total, &total);
    memset(buckets, 0, nb * sizeof(xyzz));
    const int32_t* dw = dig + (size_t)w * n;
    for (size_t i = 0; i < n; i++) {
      int32_t d = dw[i];
      if (d > 0) xyzz_add_mixed(&buckets[d - 1], &pts[i], 0);
      else if (d < 0) xyzz_add_mixed(&buckets[-d - 1], &pts[i], 1);
    }
    xyzz run, sum;
    xyzz_set_id(&run);
    xyzz_set_id(&sum);
    for (size_t b = nb; b-- > 0;) {
      xyzz_add(&run, &buckets[b]);
      xyzz_add(&sum, &run);
    }
    xyzz_add(&total, &sum);
  }
  *out = total;
  free(buckets);
  free(dig);
  free(k);
}

/* G1::multi_exp for `ncols` scalar columns sharing the bases: out[c] = sum_i s[i*ld + c] * P_i.
 * Threads split the (column, point-chunk) space like halo2curves' per-thread chunking. */
void oc_msm(const u64* points_xy, const u64* scalars, size_t n, size_t ncols, size_t ld, u64* out_xy) {
  oc_init();
  if (ncols == 0) return;
  const aff* pts = (const aff*)points_xy;
#ifdef _OPENMP
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  size_t chunks = 1;
  if (ncols < (size_t)nt) chunks = ((size_t)nt + ncols - 1) / ncols;
  if (chunks > n && n > 0) chunks = n;
  if (n < 256) chunks = 1;
  xyzz* part = (xyzz*)malloc(ncols * chunks * sizeof(xyzz));
  size_t per = chunks ? (n + chunks - 1) / chunks : 0;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (size_t c = 0; c < ncols; c++) {
    for (size_t ch = 0; ch < chunks; ch++) {
      size_t s = ch * per, e = s + per;
      if (e > n) e = n;
      if (s > e) s = e;
      msm_serial(pts + s, scalars + 4 * (s * ld + c), ld, e - s, &part[c * chunks + ch]);
    }
  }
  for (size_t c = 0; c < ncols; c++) {
    xyzz acc;
    xyzz_set_id(&acc);
    for (size_t ch = 0; ch < chunks; ch++) xyzz_add(&acc, &part[c * chunks + ch]);
    xyzz_to_affine((aff*)(out_xy + 8 * c), &acc);
  }
  free(part);
}

/* init_srs_unsafe (kzg/src/params.rs:123-139), G1 part, normalised to affine once.
 * Fixed-base windows of G make it O(32 additions) per point; the points are identical. */
void oc_srs_generate(const u64 alpha[4], size_t n, u64* out_xy) {
  oc_init();
  if (n == 0) return;
  /* table[w][d] = d * 256^w * G, d in [0,256), affine */
  aff* table = (aff*)malloc(32 * 256 * sizeof(aff));
  aff g;
  f_one(Q, &g.x);
  f_add(Q, &g.y, &g.x, &g.x);
  xyzz base;
  xyzz_set_id(&base);
  xyzz_add_mixed(&base, &g, 0);
  for (int w = 0; w < 32; w++) {
    aff base_aff;
    xyzz_to_affine(&base_aff, &base);
    xyzz acc;
    xyzz_set_id(&acc);
    memset(&table[w * 256], 0, sizeof(aff));
    for (int d = 1; d < 256; d++) {
      xyzz_add_mixed(&acc, &base_aff, 0);
      xyzz_to_affine(&table[w * 256 + d], &acc);
    }
    for (int i = 0; i < 8; i++) xyzz_dbl(&base, &base);
  }
  aff* out = (aff*)out_xy;
#pragma omp parallel
  {
#ifdef _OPENMP
    int nt = omp_get_num_threads(), id = omp_get_thread_num();
#else
    int nt = 1, id = 0;
#endif
    size_t per = (n + nt - 1) / nt, s = per * id, e = s + per;
    if (e > n) e = n;
    if (s < e) {
      fe pw;
      f_pow_u64(&FR, &pw, (const fe*)alpha, s);
      for (size_t i = s; i < e; i++) {
        u64 k[4];
        f_from_mont(&FR, k, &pw);
        xyzz acc;
        xyzz_set_id(&acc);
        for (int w = 0; w < 32; w++) {
          unsigned d = (unsigned)((k[w >> 3] >> ((w & 7) * 8)) & 0xff);
          if (d) xyzz_add_mixed(&acc, &table[w * 256 + d], 0);
        }
        xyzz_to_affine(&out[i], &acc);
        f_mul(&FR, &pw, &pw, (const fe*)alpha);
      }
    }
  }
  free(table);
}

/* quotient_and_eval, kzg/src/util.rs:100-111, column `col` of a row-major h x w matrix.
 * quot: h-1 Fr (contiguous). */
void oc_quotient_and_eval(const u64* coeffs, size_t h, size_t w, size_t col, const u64 z[4], u64* quot, u64 value[4]) {
  oc_init();
  const fe* c = (const fe*)coeffs;
  fe* q = (fe*)quot;
  if (h == 0) { memset(value, 0, 32); return; }
  fe carry = c[(h - 1) * w + col];
  for (size_t i = h - 1; i-- > 0;) {
    q[i] = carry;
    fe t;
    f_mul(&FR, &t, &carry, (const fe*)z);
    f_add(&FR, &carry, &c[i * w + col], &t);
  }
  memcpy(value, carry.v, 32);
}

/* KzgPcs::commit for one matrix (kzg/src/pcs.rs:223-265): coeffs = coset_idft_batch(evals, shift)
 * (written to coeffs_out), commits[c] = multi_exp(srs[..h], coeffs[:, c]).
 * `ncols_msm` <= w lets the baseline time a bounded sample of the columns. */
void oc_kzg_commit(const u64* evals, unsigned log_h, size_t w, const u64 shift[4], const u64* srs_xy, size_t ncols_msm,
                   u64* coeffs_out, u64* commits_xy) {
  oc_init();
  size_t h = (size_t)1 << log_h;
  memcpy(coeffs_out, evals, h * w * 32);
  oc_coset_idft_batch(coeffs_out, log_h, w, shift);
  oc_msm(srs_xy, coeffs_out, h, ncols_msm, w, commits_xy);
}

int oc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
