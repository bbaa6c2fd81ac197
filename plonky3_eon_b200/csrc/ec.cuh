// BN254 G1 (y^2 = x^3 + 3 over Fq) group law in extended-Jacobian "XYZZ" coordinates.
//
// Replaces the halo2curves 0.9 `bn256::G1` arithmetic the reference reaches through
// bn254/src/curve.rs:74,142-180 (G1, mul_scalar, multi_exp).  Group elements are unique, so
// any complete addition law gives the same affine result; XYZZ is used because a mixed
// addition with an affine SRS point costs 8M+2S and needs no inversion.
//
//   affine  : (x, y) Montgomery Fq; identity = (0, 0)   [wire format, include/eon_kzg.h]
//   XYZZ    : (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ == 0
#pragma once
#include "fp.cuh"

namespace eon {

struct G1Affine {
  Fq x, y;
  EON_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  static EON_HD G1Affine identity() { G1Affine p; p.x = Fq::zero(); p.y = Fq::zero(); return p; }
  static EON_HD G1Affine generator() {
    G1Affine p;
    p.x = Fq::one();
    p.y = fp_add(Fq::one(), Fq::one());
    return p;
  }
};

struct G1Xyzz {
  Fq X, Y, ZZ, ZZZ;
  EON_HD bool is_identity() const { return ZZ.is_zero(); }
  static EON_HD G1Xyzz identity() {
    G1Xyzz p;
    p.X = Fq::zero(); p.Y = Fq::zero(); p.ZZ = Fq::zero(); p.ZZZ = Fq::zero();
    return p;
  }
  static EON_HD G1Xyzz from_affine(const G1Affine& a) {
    G1Xyzz p;
    if (a.is_identity()) return identity();
    p.X = a.x; p.Y = a.y; p.ZZ = Fq::one(); p.ZZZ = Fq::one();
    return p;
  }
};

// 2 * (affine point), result XYZZ.  (mdbl-2008-s-1, a = 0)
EON_HD G1Xyzz g1_dbl_affine(const G1Affine& p) {
  if (p.is_identity() || p.y.is_zero()) return G1Xyzz::identity();
  G1Xyzz r;
  Fq U = fp_dbl(p.y);
  Fq V = fp_sqr(U);
  Fq W = fp_mul(U, V);
  Fq S = fp_mul(p.x, V);
  Fq X2 = fp_sqr(p.x);
  Fq M = fp_add(fp_dbl(X2), X2);
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.y));
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}

// 2 * (XYZZ point).  (dbl-2008-s-1, a = 0)
EON_HD G1Xyzz g1_dbl(const G1Xyzz& p) {
  if (p.is_identity() || p.Y.is_zero()) return G1Xyzz::identity();
  G1Xyzz r;
  Fq U = fp_dbl(p.Y);
  Fq V = fp_sqr(U);
  Fq W = fp_mul(U, V);
  Fq S = fp_mul(p.X, V);
  Fq X2 = fp_sqr(p.X);
  Fq M = fp_add(fp_dbl(X2), X2);
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.Y));
  r.ZZ = fp_mul(V, p.ZZ);
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}

// acc += (affine q), complete.  (madd-2008-s; 8M + 2S on the generic path)
EON_HD void g1_add_mixed(G1Xyzz& acc, const G1Affine& q) {
  if (q.is_identity()) return;
  if (acc.is_identity()) {
    acc.X = q.x; acc.Y = q.y; acc.ZZ = Fq::one(); acc.ZZZ = Fq::one();
    return;
  }
  Fq U2 = fp_mul(q.x, acc.ZZ);
  Fq S2 = fp_mul(q.y, acc.ZZZ);
  Fq Pp = fp_sub(U2, acc.X);
  Fq Rr = fp_sub(S2, acc.Y);
  if (Pp.is_zero()) {
    if (Rr.is_zero()) acc = g1_dbl_affine(q);  // same point
    else acc = G1Xyzz::identity();             // opposite points
    return;
  }
  Fq PP = fp_sqr(Pp);
  Fq PPP = fp_mul(Pp, PP);
  Fq Qq = fp_mul(acc.X, PP);
  Fq X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  Fq Y3 = fp_sub(fp_mul(Rr, fp_sub(Qq, X3)), fp_mul(acc.Y, PPP));
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = fp_mul(acc.ZZ, PP);
  acc.ZZZ = fp_mul(acc.ZZZ, PPP);
}

// acc += q (both XYZZ), complete.  (add-2008-s; 12M + 2S)
EON_HD void g1_add(G1Xyzz& acc, const G1Xyzz& q) {
  if (q.is_identity()) return;
  if (acc.is_identity()) { acc = q; return; }
  Fq U1 = fp_mul(acc.X, q.ZZ);
  Fq U2 = fp_mul(q.X, acc.ZZ);
  Fq S1 = fp_mul(acc.Y, q.ZZZ);
  Fq S2 = fp_mul(q.Y, acc.ZZZ);
  Fq Pp = fp_sub(U2, U1);
  Fq Rr = fp_sub(S2, S1);
  if (Pp.is_zero()) {
    if (Rr.is_zero()) acc = g1_dbl(acc);
    else acc = G1Xyzz::identity();
    return;
  }
  Fq PP = fp_sqr(Pp);
  Fq PPP = fp_mul(Pp, PP);
  Fq Qq = fp_mul(U1, PP);
  Fq X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  Fq Y3 = fp_sub(fp_mul(Rr, fp_sub(Qq, X3)), fp_mul(S1, PPP));
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = fp_mul(fp_mul(acc.ZZ, q.ZZ), PP);
  acc.ZZZ = fp_mul(fp_mul(acc.ZZZ, q.ZZZ), PPP);
}

EON_HD G1Affine g1_neg(const G1Affine& p) {
  G1Affine r;
  r.x = p.x;
  r.y = fp_neg(p.y);
  return r;
}

// XYZZ -> affine (one Fq inversion).  identity -> (0, 0).
EON_HD G1Affine g1_to_affine(const G1Xyzz& p) {
  if (p.is_identity()) return G1Affine::identity();
  Fq i = fp_inv(fp_mul(p.ZZ, p.ZZZ));  // 1 / Z^5
  Fq zz_inv = fp_mul(i, p.ZZZ);        // 1 / ZZ
  Fq zzz_inv = fp_mul(i, p.ZZ);        // 1 / ZZZ
  G1Affine r;
  r.x = fp_mul(p.X, zz_inv);
  r.y = fp_mul(p.Y, zzz_inv);
  return r;
}

// k * p for a canonical (non-Montgomery) 256-bit scalar k given as 8 x u32.
EON_HD G1Xyzz g1_mul_canonical(const G1Affine& p, const u32 k[8]) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = 255; i >= 0; i--) {
    acc = g1_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) g1_add_mixed(acc, p);
  }
  return acc;
}

// k * p for a small integer k (bucket-chunk offsets in the MSM reduction)
EON_HD G1Xyzz g1_mul_u32(const G1Xyzz& p, u32 k) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = 31; i >= 0; i--) {
    acc = g1_dbl(acc);
    if ((k >> i) & 1) g1_add(acc, p);
  }
  return acc;
}

}  // namespace eon
