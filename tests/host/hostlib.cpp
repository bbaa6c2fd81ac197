// Host-compiled view of csrc/fp.cuh + ec.cuh (software carry flag) so that the field and
// curve arithmetic shipped in the CUDA library can be checked against the oracle without a GPU.
#include "../../plonky3_eon_b200/csrc/ec.cuh"
#include "../../plonky3_eon_b200/csrc/fp_shoup.cuh"
#include <string.h>
using namespace eon;

template <class F> static F ld(const uint32_t* p) { F x; memcpy(x.v, p, 32); return x; }
template <class F> static void st(uint32_t* p, const F& x) { memcpy(p, x.v, 32); }

extern "C" {
// which: 0 = Fr, 1 = Fq.  op: 0 mul, 1 add, 2 sub, 3 neg, 4 inv (binary GCD), 5 from_mont, 6 to_mont, 7 sqr,
// 8 inv by Fermat (cross-check), 9 inv by safegcd divsteps, 10 inv by the binary GCD (4 = the library's fp_inv)
void host_fp_op(int which, int op, const uint32_t* a, const uint32_t* b, uint32_t* r, int n) {
  for (int i = 0; i < n; i++, a += 8, b += 8, r += 8) {
    if (which == 0) {
      Fr x = ld<Fr>(a), y = ld<Fr>(b), z;
      switch (op) {
        case 0: z = fp_mul(x, y); break;
        case 1: z = fp_add(x, y); break;
        case 2: z = fp_sub(x, y); break;
        case 3: z = fp_neg(x); break;
        case 4: z = fp_inv(x); break;
        case 5: fp_from_mont(z.v, x); break;
        case 6: z = fp_to_mont<FrParams>(x.v); break;
        case 8: z = fp_inv_fermat(x); break;
        case 9: z = fp_inv_safegcd(x); break;
        case 10: z = fp_inv_bgcd(x); break;
        default: z = fp_sqr(x);
      }
      st(r, z);
    } else {
      Fq x = ld<Fq>(a), y = ld<Fq>(b), z;
      switch (op) {
        case 0: z = fp_mul(x, y); break;
        case 1: z = fp_add(x, y); break;
        case 2: z = fp_sub(x, y); break;
        case 3: z = fp_neg(x); break;
        case 4: z = fp_inv(x); break;
        case 5: fp_from_mont(z.v, x); break;
        case 6: z = fp_to_mont<FqParams>(x.v); break;
        case 8: z = fp_inv_fermat(x); break;
        case 9: z = fp_inv_safegcd(x); break;
        case 10: z = fp_inv_bgcd(x); break;
        default: z = fp_sqr(x);
      }
      st(r, z);
    }
  }
}

// Raw lazy products (result in [0, 2p), not canonicalised).  variant: 0 = word-serial (CIOS),
// 1 = split (Karatsuba product + separate reduction), 2 = split square of a (b ignored).
// a < p; b is any 256-bit value.
void host_fp_mul_lazy(int which, int variant, const uint32_t* a, const uint32_t* b, uint32_t* r, int n) {
  for (int i = 0; i < n; i++, a += 8, b += 8, r += 8) {
    if (which == 0) {
      if (variant == 0) fp_mul_lazy_cios<FrParams>(r, a, b);
      else if (variant == 1) fp_mul_lazy_split<FrParams>(r, a, b);
      else fp_sqr_lazy_split<FrParams>(r, a);
    } else {
      if (variant == 0) fp_mul_lazy_cios<FqParams>(r, a, b);
      else if (variant == 1) fp_mul_lazy_split<FqParams>(r, a, b);
      else fp_sqr_lazy_split<FqParams>(r, a);
    }
  }
}
// Fixed-operand product (fp_shoup.cuh): r = a*w - q~*p in [0, 3p) for a = any 256-bit value, w < p and
// wq = floor(w 2^256 / p); canon != 0 additionally reduces r to [0, p).
void host_shoup_mul(int which, int canon, const uint32_t* a, const uint32_t* w, const uint32_t* wq, uint32_t* r, int n) {
  for (int i = 0; i < n; i++, a += 8, w += 8, wq += 8, r += 8) {
    uint32_t t[8];
    if (which == 0) {
      shoup::mul_lazy<FrParams>(t, a, w, wq);
      if (canon) { Fr c = shoup::canon_3p<FrParams>(t); memcpy(t, c.v, 32); }
    } else {
      shoup::mul_lazy<FqParams>(t, a, w, wq);
      if (canon) { Fq c = shoup::canon_3p<FqParams>(t); memcpy(t, c.v, 32); }
    }
    memcpy(r, t, 32);
  }
}
// (w, wq) of a twiddle given in Montgomery form (fp_shoup.cuh precompute)
void host_shoup_precompute(int which, const uint32_t* rho, uint32_t* w, uint32_t* wq, int n) {
  for (int i = 0; i < n; i++, rho += 8, w += 8, wq += 8) {
    if (which == 0) shoup::precompute<FrParams>(w, wq, ld<Fr>(rho));
    else shoup::precompute<FqParams>(w, wq, ld<Fq>(rho));
  }
}
// T[0..16) = a * b (variant 0) or a^2 (variant 1) over full 256-bit operands
void host_wide_product(int variant, const uint32_t* a, const uint32_t* b, uint32_t* T, int n) {
  for (int i = 0; i < n; i++, a += 8, b += 8, T += 16) {
    if (variant == 0) detail::mul8_wide(T, a, b);
    else detail::sqr8_wide(T, a);
  }
}

// Lazily reduced NTT butterflies (fp.cuh): one forward (DIT) and one inverse (DIF) radix-2 butterfly on
// raw 256-bit limbs, then canonicalised.  a, b may be anywhere in [0, 4r) (dit) / [0, 2r) (dif).
void host_lazy_butterfly(int dif, const uint32_t* a, const uint32_t* b, const uint32_t* tw, uint32_t* o0, uint32_t* o1) {
  uint32_t x[8], t[8], s[8], d[8];
  if (!dif) {
    fp_reduce_2p<FrParams>(x, a);
    fp_mul_lazy<FrParams>(t, tw, b);
    fp_add_raw(s, x, t);
    fp_sub_plus_2p<FrParams>(d, x, t);
  } else {
    uint32_t u[8];
    fp_add_raw(u, a, b);
    fp_reduce_2p<FrParams>(s, u);
    fp_sub_plus_2p<FrParams>(x, a, b);
    fp_mul_lazy<FrParams>(d, tw, x);
  }
  st(o0, fp_canon_4p<FrParams>(s));
  st(o1, fp_canon_4p<FrParams>(d));
}

// The same butterflies with the twiddle as a (plain, quotient) pair and the fixed-operand product, exactly as
// k_ntt_pass<.., SH = true> runs them: t = reduce_2p(shoup(b)), everything else unchanged.  tw: Montgomery form
// (what k_gen_twiddles starts from); raw != 0 additionally returns the un-canonicalised outputs in r0 / r1.
void host_lazy_butterfly_shoup(int dif, const uint32_t* a, const uint32_t* b, const uint32_t* tw, uint32_t* o0,
                               uint32_t* o1, uint32_t* r0, uint32_t* r1) {
  uint32_t w[8], wq[8], x[8], t3[8], t[8], s[8], d[8];
  shoup::precompute<FrParams>(w, wq, ld<Fr>(tw));
  if (!dif) {
    fp_reduce_2p<FrParams>(x, a);
    shoup::mul_lazy<FrParams>(t3, b, w, wq);
    fp_reduce_2p<FrParams>(t, t3);
    fp_add_raw(s, x, t);
    fp_sub_plus_2p<FrParams>(d, x, t);
  } else {
    uint32_t u[8];
    fp_add_raw(u, a, b);
    fp_reduce_2p<FrParams>(s, u);
    fp_sub_plus_2p<FrParams>(x, a, b);
    shoup::mul_lazy<FrParams>(t3, x, w, wq);
    fp_reduce_2p<FrParams>(d, t3);
  }
  memcpy(r0, s, 32);
  memcpy(r1, d, 32);
  st(o0, fp_canon_4p<FrParams>(s));
  st(o1, fp_canon_4p<FrParams>(d));
}

static G1Affine lda(const uint32_t* p) { G1Affine a; memcpy(a.x.v, p, 32); memcpy(a.y.v, p + 8, 32); return a; }
static void sta(uint32_t* p, const G1Affine& a) { memcpy(p, a.x.v, 32); memcpy(p + 8, a.y.v, 32); }

// r = sum of n affine points, via mixed adds (exercises identity / doubling / inverse branches)
void host_g1_sum(const uint32_t* pts, int n, uint32_t* r) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = 0; i < n; i++) g1_add_mixed(acc, lda(pts + 16 * i));
  sta(r, g1_to_affine(acc));
}
// r = (a0 + a1) + (b0 + b1) using the full XYZZ+XYZZ addition
void host_g1_add_full(const uint32_t* a0, const uint32_t* a1, const uint32_t* b0, const uint32_t* b1, uint32_t* r) {
  G1Xyzz A = G1Xyzz::identity(), B = G1Xyzz::identity();
  g1_add_mixed(A, lda(a0)); g1_add_mixed(A, lda(a1));
  g1_add_mixed(B, lda(b0)); g1_add_mixed(B, lda(b1));
  g1_add(A, B);
  sta(r, g1_to_affine(A));
}
// r = k * p, k canonical 256-bit
void host_g1_mul(const uint32_t* p, const uint32_t* k, uint32_t* r) {
  sta(r, g1_to_affine(g1_mul_canonical(lda(p), k)));
}
void host_g1_mul_u32(const uint32_t* p, uint32_t k, uint32_t* r) {
  G1Xyzz P = G1Xyzz::from_affine(lda(p));
  P = g1_dbl(P);  // make it non-trivially projective: computes k * (2p)
  sta(r, g1_to_affine(g1_mul_u32(P, k)));
}
}
