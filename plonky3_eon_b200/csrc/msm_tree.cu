// Batched-affine pairwise rounds of the MSM bucket accumulation (sm_100a).
//
// The bucket phase of G1::multi_exp (bn254/src/curve.rs:158-180 -> halo2curves msm_best) adds every
// (point, window) entry into its bucket.  With an inversion-free XYZZ accumulator that is 10 Fq
// products per addition (msm.cu, k_msm_accumulate).  An AFFINE addition costs 2M + 1S plus one field
// inversion; sharing ONE inversion among all additions of a launch (Montgomery's trick, 3 products
// per element) brings the total to ~7.6 products per addition.
//
// The entry array is sorted by bucket with every bucket starting on a multiple of 2^R slots and
// unused slots holding ENTRY_NONE (= the identity).  A round is then a flat, bucket-agnostic
//      out[j] = in[2j] + in[2j+1]            for all j
// (identity + P = P, so padding needs no special handling), and after R rounds bucket g owns the
// slots [start_g >> R, ceil(end_g / 2^R)), which the serial XYZZ finisher sums.  R = 3 moves 7/8 of
// all additions to the affine form.
//
// One round =
//   k_tree_fwd   thread t: denominators d_j of its B = 8 pairs, pre[j] = d_0 .. d_(j-1), T0[t] = prod d_j
//   k_tree_up    T(l+1)[u] = prod of 8 consecutive T(l)        (until <= 256 values remain)
//   k_tree_top   Fermat inversion of the top level              (the only inversions of the round)
//   k_tree_down  T(l)[i] <- 1 / T(l)[i] from 1 / T(l+1)[u]
//   k_tree_bwd   thread t: peel 1/d_j = pre[j] / (d_0 .. d_j) off 1/T0[t], last pair first, and finish
//                lambda = (y2 - y1)/d, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1
// Degenerate pairs (identity operand, P + P, P + (-P)) contribute d = 1 and are resolved in k_tree_bwd,
// so the group law stays complete and the final affine sums are the same unique group elements.
#include <stdlib.h>

#include "msm.cuh"

namespace eon {

constexpr int TREE_THREADS = 128;
constexpr u64 TREE_TOP = 256;      // level size at which the values are inverted directly

enum { PAIR_TRIVIAL = 0, PAIR_ADD = 1, PAIR_DBL = 2 };

struct TreeSrc {
  const G1Affine* pts;  // rounds >= 1
  const u32* entries;   // round 0: entry = base index | sign << 31, or ENTRY_NONE
  const G1Affine* bases;
};

// 32-byte field elements move as single 256-bit accesses (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100):
// a random gather then costs one request per coordinate instead of two that miss separately.
__device__ __forceinline__ Fq ldg_fq(const Fq* p) {
  Fq r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_fq(Fq* p, const Fq& a) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]),
               "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
               : "memory");
}

// One operand of a pair, loaded lazily: x first (the common path needs nothing else).
template <bool R0>
struct Operand {
  const G1Affine* p;  // null: ENTRY_NONE
  bool neg;
  Fq x;
  __device__ __forceinline__ void open(const TreeSrc& s, u64 slot) {
    if (R0) {
      u32 v = __ldg(s.entries + slot);
      neg = (v & SIGN_BIT) != 0;
      p = (v == ENTRY_NONE) ? nullptr : s.bases + (v & ~SIGN_BIT);
    } else {
      neg = false;
      p = s.pts + slot;
    }
    x = p ? ldg_fq(&p->x) : Fq::zero();
  }
  __device__ __forceinline__ Fq y() const {
    if (!p) return Fq::zero();
    Fq v = ldg_fq(&p->y);
    return neg ? fp_neg(v) : v;
  }
  __device__ __forceinline__ bool is_identity() const { return !p || (x.is_zero() && ldg_fq(&p->y).is_zero()); }
};

// Denominator of pair j and what kind of pair it is.  Used identically by k_tree_fwd and by both
// passes of k_tree_bwd, so the three always agree.
template <bool R0>
__device__ __forceinline__ int pair_denominator(const Operand<R0>& P, const Operand<R0>& Q, Fq& d) {
  if (P.is_identity() || Q.is_identity()) return PAIR_TRIVIAL;
  d = fp_sub(Q.x, P.x);
  if (!d.is_zero()) return PAIR_ADD;
  Fq py = P.y(), qy = Q.y();
  if (py == qy && !py.is_zero()) {  // P == Q: tangent slope 3 x^2 / (2 y)
    d = fp_dbl(py);
    return PAIR_DBL;
  }
  return PAIR_TRIVIAL;  // P == -Q (or a 2-torsion input): the sum is the identity
}

template <bool R0, int TREE_B>
__global__ void __launch_bounds__(TREE_THREADS)
k_tree_fwd(TreeSrc src, u64 npairs, Fq* __restrict__ T0, Fq* __restrict__ pre) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u64 j0 = t * TREE_B;
  if (j0 >= npairs) return;
  Fq run = Fq::one();
  bool any = false;
#pragma unroll 1
  for (int i = 0; i < TREE_B; i++) {
    u64 j = j0 + i;
    if (j >= npairs) break;
    st_fq(pre + j, run);  // product of the thread's non-trivial denominators before pair j
    Operand<R0> P, Q;
    P.open(src, 2 * j);
    Q.open(src, 2 * j + 1);
    Fq d;
    if (pair_denominator<R0>(P, Q, d) != PAIR_TRIVIAL) {
      run = any ? fp_mul(run, d) : d;
      any = true;
    }
  }
  st_fq(T0 + t, run);
}

// T_out[u] = prod T_in[8u .. 8u+8)
__global__ void __launch_bounds__(TREE_THREADS) k_tree_up(const Fq* __restrict__ Tin, u64 n, Fq* __restrict__ Tout) {
  const u64 u = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u64 i0 = u * 8;
  if (i0 >= n) return;
  Fq run = ldg_fq(Tin + i0);
  for (u64 i = i0 + 1; i < min(n, i0 + 8); i++) run = fp_mul(run, ldg_fq(Tin + i));
  st_fq(Tout + u, run);
}

__global__ void __launch_bounds__(32) k_tree_top(Fq* __restrict__ T, u64 n) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < n) T[i] = fp_inv(T[i]);
}

// T[8u + i] <- 1 / T[8u + i], given Tinv_up[u] = 1 / prod_i T[8u + i]
__global__ void __launch_bounds__(TREE_THREADS) k_tree_down(Fq* __restrict__ T, u64 n, const Fq* __restrict__ Tinv_up) {
  const u64 u = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u64 i0 = u * 8;
  if (i0 >= n) return;
  const int cnt = (int)min((u64)8, n - i0);
  Fq e[8], pre[8];
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (i < cnt) e[i] = ldg_fq(T + i0 + i);
  pre[0] = e[0];
#pragma unroll
  for (int i = 1; i < 8; i++)
    if (i < cnt) pre[i] = fp_mul(pre[i - 1], e[i]);
  Fq inv = ldg_fq(Tinv_up + u);
#pragma unroll
  for (int i = 7; i >= 1; i--) {
    if (i < cnt) {
      st_fq(T + i0 + i, fp_mul(inv, pre[i - 1]));
      inv = fp_mul(inv, e[i]);
    }
  }
  st_fq(T + i0, inv);
}

template <bool R0, int TREE_B>
__global__ void __launch_bounds__(TREE_THREADS, 4)
k_tree_bwd(TreeSrc src, u64 npairs, const Fq* __restrict__ T0inv, const Fq* __restrict__ pre_all,
           G1Affine* __restrict__ out) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u64 j0 = t * TREE_B;
  if (j0 >= npairs) return;
  const int cnt = (int)min((u64)TREE_B, npairs - j0);
  // backward: inv = 1 / (d_0 ... d_i) at the top of iteration i
  Fq inv = ldg_fq(T0inv + t);
#pragma unroll 1
  for (int i = cnt - 1; i >= 0; i--) {
    Operand<R0> P, Q;
    P.open(src, 2 * (j0 + i));
    Q.open(src, 2 * (j0 + i) + 1);
    Fq d;
    const int kind = pair_denominator<R0>(P, Q, d);
    G1Affine r;
    if (kind == PAIR_TRIVIAL) {
      const bool pid = P.is_identity(), qid = Q.is_identity();
      if (pid && !qid) { r.x = Q.x; r.y = Q.y(); }
      else if (qid && !pid) { r.x = P.x; r.y = P.y(); }
      else r = G1Affine::identity();  // both identity, or P == -Q
    } else {
      const Fq pre = ldg_fq(pre_all + j0 + i);
      const Fq dinv = fp_mul(inv, pre);  // pre == 1 (Montgomery one) for the first non-trivial pair
      inv = fp_mul(inv, d);
      const Fq py = P.y();
      Fq lam;
      Fq xsum;
      if (kind == PAIR_ADD) {
        lam = fp_mul(fp_sub(Q.y(), py), dinv);
        xsum = fp_add(P.x, Q.x);
      } else {
        Fq xx = fp_sqr(P.x);
        lam = fp_mul(fp_add(fp_dbl(xx), xx), dinv);
        xsum = fp_dbl(P.x);
      }
      r.x = fp_sub(fp_sqr(lam), xsum);
      r.y = fp_sub(fp_mul(lam, fp_sub(P.x, r.x)), py);
    }
    st_fq(&out[j0 + i].x, r.x);
    st_fq(&out[j0 + i].y, r.y);
  }
}

static unsigned grid_for(u64 threads) { return (unsigned)((threads + TREE_THREADS - 1) / TREE_THREADS); }

template <int B>
static void launch_fwd(bool r0, unsigned grid, cudaStream_t st, const TreeSrc& src, u64 npairs, Fq* T0, Fq* pre) {
  if (r0) k_tree_fwd<true, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0, pre);
  else k_tree_fwd<false, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0, pre);
}
template <int B>
static void launch_bwd(bool r0, unsigned grid, cudaStream_t st, const TreeSrc& src, u64 npairs, const Fq* T0inv,
                       const Fq* pre, G1Affine* out) {
  if (r0) k_tree_bwd<true, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0inv, pre, out);
  else k_tree_bwd<false, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0inv, pre, out);
}

// pairs per thread (one shared denominator product per thread): 8, 16 or 32
static int tree_b() {
  static int b = 0;
  if (!b) {
    const char* e = getenv("EON_TREE_B");
    b = e ? atoi(e) : 32;
    if (b != 8 && b != 16 && b != 32) b = 32;
  }
  return b;
}

int msm_tree_rounds(eon_ctx* ctx, const G1Affine* d_bases, const u32* d_entries, u64 total_slots, u32 rounds,
                    const G1Affine** out_pts) {
  const u64 TREE_B = (u64)tree_b();
  if (rounds == 0 || (total_slots & ((1ull << rounds) - 1)))
    return fail(ctx, EON_ERR_BAD_ARG, "msm_tree_rounds: slot count not aligned to 2^rounds");
  cudaStream_t st = ctx->stream;
  const u64 m0 = total_slots / 2;  // pairs of round 0
  void *bufA, *bufB, *bufT;
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_A, m0 * sizeof(G1Affine), &bufA));
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_B, (m0 / 2 + 1) * sizeof(G1Affine), &bufB));
  // level sizes of round 0 (the largest round): n0 = ceil(m0 / B), n(l+1) = ceil(n(l) / 8)
  u64 tcap = 0;
  for (u64 n = (m0 + TREE_B - 1) / TREE_B;; n = (n + 7) / 8) {
    tcap += n;
    if (n <= TREE_TOP) break;
  }
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_T, tcap * sizeof(Fq), &bufT));
  void* bufP;
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_P, m0 * sizeof(Fq), &bufP));

  TreeSrc src;
  src.pts = nullptr;
  src.entries = d_entries;
  src.bases = d_bases;
  u64 npairs = m0;
  G1Affine* outs[2] = {(G1Affine*)bufA, (G1Affine*)bufB};
  for (u32 r = 0; r < rounds; r++) {
    G1Affine* out = outs[r & 1];
    std::vector<std::pair<Fq*, u64>> lv;  // (array, size) per level
    Fq* T = (Fq*)bufT;
    for (u64 n = (npairs + TREE_B - 1) / TREE_B;; n = (n + 7) / 8) {
      lv.push_back(std::make_pair(T, n));
      T += n;
      if (n <= TREE_TOP) break;
    }
    const unsigned g0 = grid_for(lv[0].second);
    phase_begin(ctx, PH_MSM_TREE_FWD);
    if (TREE_B == 8) launch_fwd<8>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    else if (TREE_B == 16) launch_fwd<16>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    else launch_fwd<32>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_TREE_FWD);
    phase_begin(ctx, PH_MSM_TREE_INV);
    for (size_t l = 0; l + 1 < lv.size(); l++) {
      k_tree_up<<<grid_for(lv[l + 1].second), TREE_THREADS, 0, st>>>(lv[l].first, lv[l].second, lv[l + 1].first);
      EON_LAUNCHED(ctx);
    }
    k_tree_top<<<(unsigned)((lv.back().second + 31) / 32), 32, 0, st>>>(lv.back().first, lv.back().second);
    EON_LAUNCHED(ctx);
    for (size_t l = lv.size() - 1; l-- > 0;) {
      k_tree_down<<<grid_for(lv[l + 1].second), TREE_THREADS, 0, st>>>(lv[l].first, lv[l].second, lv[l + 1].first);
      EON_LAUNCHED(ctx);
    }
    phase_end(ctx, PH_MSM_TREE_INV);
    phase_begin(ctx, PH_MSM_TREE_BWD);
    if (TREE_B == 8) launch_bwd<8>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out);
    else if (TREE_B == 16) launch_bwd<16>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out);
    else launch_bwd<32>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_TREE_BWD);
    src.pts = out;
    npairs /= 2;
  }
  *out_pts = src.pts;
  return EON_OK;
}

}  // namespace eon
