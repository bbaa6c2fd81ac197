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
//   k_tree_up    T(l+1)[u] = prod of 8 consecutive T(l)        (until <= TREE_TOP values remain)
//   k_tree_top   inversion of every value of the top level       (the only inversions of the round)
//   k_tree_down  T(l)[i] <- 1 / T(l)[i] from 1 / T(l+1)[u]
//   k_tree_bwd   thread t: peel 1/d_j = pre[j] / (d_0 .. d_j) off 1/T0[t], last pair first, and finish
//                lambda = (y2 - y1)/d, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1
// Degenerate pairs (identity operand, P + P, P + (-P)) contribute d = 1 and are resolved in k_tree_bwd,
// so the group law stays complete and the final affine sums are the same unique group elements.
#include <stdlib.h>

#include <algorithm>

#include "msm.cuh"

namespace eon {

constexpr int TREE_THREADS = 128;
// level size at which the values are inverted directly: with the divstep inversion (fp.cuh, ~15 k instructions)
// 16 k inversions are 512 warps = one per SMSP, i.e. the latency of a single inversion (~20 us), and every level of
// the product tree saved is two launches and ~22 dependent products less per round (2048 with the binary GCD)
constexpr u64 TREE_TOP = 16384;

enum { PAIR_TRIVIAL = 0, PAIR_ADD = 1, PAIR_DBL = 2 };

struct TreeSrc {
  // rounds >= 1: the previous round's sums as two arrays (x | y).  k_tree_fwd needs the x coordinates only (the
  // denominators x2 - x1), so with the coordinates apart it streams half the bytes of an array of points
  // (round 1 + 2 of a 2^20 x 16 commit: 12 -> 6 GB of reads).
  const Fq* pts_x;
  const Fq* pts_y;
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

// L2 residency hints for the slice schedule: base points are kept (evict-last), the streams that pass through
// once -- pair records in, prefix products and sums out -- are marked evict-first so they do not push the
// slice out of the L2.  -DEON_NO_L2_HINTS drops them (plain accesses).
__device__ __forceinline__ Fq ldg_fq_keep(const Fq* p) {
#if defined(EON_NO_L2_HINTS)
  return ldg_fq(p);
#else
  Fq r;
  asm volatile("ld.global.nc.L1::evict_last.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
#endif
}
__device__ __forceinline__ void st_fq_stream(Fq* p, const Fq& a) {
#if defined(EON_NO_L2_HINTS)
  st_fq(p, a);
#else
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
               "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
               : "memory");
#endif
}
__device__ __forceinline__ Fq ldg_fq_stream(const Fq* p) {
#if defined(EON_NO_L2_HINTS)
  return ldg_fq(p);
#else
  Fq r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
#endif
}
__device__ __forceinline__ uint2 ldg_rec(const uint2* p) {
#if defined(EON_NO_L2_HINTS)
  return __ldg(p);
#else
  uint2 r;
  // (the L2 eviction modifiers exist for 256-bit accesses only)
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
#endif
}

// One operand of a pair, loaded lazily: x first (the common path needs nothing else).
template <bool R0>
struct Operand {
  const Fq* py;       // where the y coordinate lives; null: ENTRY_NONE
  bool neg;
  bool keep;          // slice schedule: the point comes from an L2-resident table slice
  Fq x;
  __device__ __forceinline__ Fq load(const Fq* q) const { return keep ? ldg_fq_keep(q) : ldg_fq(q); }
  __device__ __forceinline__ void open_entry(const G1Affine* bases, u32 v) {
    keep = true;
    neg = (v & SIGN_BIT) != 0;
    const G1Affine* p = (v == ENTRY_NONE) ? nullptr : bases + (v & ~SIGN_BIT);
    py = p ? &p->y : nullptr;
    x = p ? load(&p->x) : Fq::zero();
  }
  __device__ __forceinline__ void open(const TreeSrc& s, u64 slot) {
    keep = false;
    if (R0) {
      u32 v = __ldg(s.entries + slot);
      neg = (v & SIGN_BIT) != 0;
      const G1Affine* p = (v == ENTRY_NONE) ? nullptr : s.bases + (v & ~SIGN_BIT);
      py = p ? &p->y : nullptr;
      x = p ? ldg_fq(&p->x) : Fq::zero();
    } else {
      neg = false;
      py = s.pts_y + slot;
      x = ldg_fq(s.pts_x + slot);
    }
  }
  __device__ __forceinline__ Fq y() const {
    if (!py) return Fq::zero();
    Fq v = load(py);
    return neg ? fp_neg(v) : v;
  }
  __device__ __forceinline__ bool is_identity() const { return !py || (x.is_zero() && load(py).is_zero()); }
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

// Sum of one pair given the shared inverse chain: inv = 1 / (d_0 ... d_i) on entry, 1 / (d_0 ... d_(i-1)) on
// exit (unchanged for a trivial pair); pre = d_0 ... d_(i-1) (read only for a non-trivial pair).
template <bool R0>
__device__ __forceinline__ G1Affine pair_sum(const Operand<R0>& P, const Operand<R0>& Q, int kind, const Fq& d,
                                             Fq& inv, const Fq* pre_ptr) {
  G1Affine r;
  if (kind == PAIR_TRIVIAL) {
    const bool pid = P.is_identity(), qid = Q.is_identity();
    if (pid && !qid) { r.x = Q.x; r.y = Q.y(); }
    else if (qid && !pid) { r.x = P.x; r.y = P.y(); }
    else r = G1Affine::identity();  // both identity, or P == -Q
  } else {
    const Fq pre = P.keep ? ldg_fq_stream(pre_ptr) : ldg_fq(pre_ptr);
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
    r.x = fp_sub(fp_sqr_dedicated(lam), xsum);
    r.y = fp_sub(fp_mul(lam, fp_sub(P.x, r.x)), py);
  }
  return r;
}

template <bool R0, int TREE_B>
__global__ void __launch_bounds__(TREE_THREADS, 4)
k_tree_bwd(TreeSrc src, u64 npairs, const Fq* __restrict__ T0inv, const Fq* __restrict__ pre_all,
           Fq* __restrict__ out_x, Fq* __restrict__ out_y, u32 ostride) {
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
    const G1Affine r = pair_sum<R0>(P, Q, kind, d, inv, pre_all + j0 + i);
    st_fq(out_x + (j0 + i) * ostride, r.x);
    st_fq(out_y + (j0 + i) * ostride, r.y);
  }
}


// ---- round 0 scheduled by table slice ("slice schedule") ------------------------------------------------
// With window tables every base is used once per column, but the entries are sorted by bucket, so the round-0
// gathers of a commit walk the whole table (15 x 64 MiB at 2^20 points) at random: 47-50 G accesses/s from
// DRAM against 227 G/s (x only; 127 G/s for x and y) when the working set fits the L2, which holds up to
// ~80 MiB at full rate (tools/probes/l2_gather_probe.cu, profiles/r01j_l2_gather_probe.log).  The bucket sums
// do not depend on which entries of a bucket are paired nor on the order in which pairs are processed, so:
//   1. the sort orders the entries of every bucket by table slice (2^19 points = 32 MiB; k_sort_fine<true>),
//      so that most pairs take both operands from the same slice;
//   2. a pair record (entry0, entry1, destination slot) is appended to the list of the slice of its first
//      operand (tile-local counting sort, k_pair_hist / k_pair_scatter -- or, in the fused form of the sort,
//      written by k_sort_place2 itself: msm_sort.cu);
//   3. k_tree_fwd_sliced / k_tree_bwd_sliced walk the records in that order with coalesced record reads: at
//      any time the in-flight pairs gather from one or two slices, which stay L2-resident.
// Each result is written to the slot the slot-order schedule would have used, so rounds >= 1 and the finisher
// are unchanged.  Measured at 2^20 x 16 (B200): tree fwd 9.9 -> 6.4 ms, tree bwd 18.9 -> 18.2 ms, step
// 51.6 -> 47.2 ms; without step 1 only 51.6 -> 50.8 ms (the second operand's random gathers evict the slice).
constexpr u32 SLICE_MAX_BINS = 1024;
constexpr int PAIR_TILE = 2048;         // pairs per CTA of the record sort (256 threads x 8)

__device__ __forceinline__ u32 pair_bin(u32 e0, u32 e1, u32 shift) {
  const u32 v = (e0 != ENTRY_NONE) ? e0 : e1;
  return (v == ENTRY_NONE) ? 0u : ((v & ~SIGN_BIT) >> shift);
}

__global__ void __launch_bounds__(256) k_pair_hist(const uint2* __restrict__ pairs, u64 npairs, u32 nbins, u32 shift,
                                                   unsigned long long* __restrict__ counts) {
  __shared__ u32 h[SLICE_MAX_BINS];
  for (u32 b = threadIdx.x; b < nbins; b += blockDim.x) h[b] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * PAIR_TILE;
#pragma unroll
  for (int i = 0; i < PAIR_TILE / 256; i++) {
    const u64 j = base + i * 256 + threadIdx.x;
    if (j < npairs) {
      const uint2 e = __ldg(pairs + j);
      atomicAdd(&h[pair_bin(e.x, e.y, shift)], 1u);
    }
  }
  __syncthreads();
  for (u32 b = threadIdx.x; b < nbins; b += blockDim.x)
    if (h[b]) atomicAdd(&counts[b], (unsigned long long)h[b]);
}

// counts[b] -> cursor[b] = exclusive prefix (one warp; nbins <= 1024)
__global__ void k_pair_scan(const unsigned long long* __restrict__ counts, u32 nbins,
                            unsigned long long* __restrict__ cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long acc = 0;
    for (u32 b = 0; b < nbins; b++) {
      cursor[b] = acc;
      acc += counts[b];
    }
  }
}

// Tile-local counting sort of the pair records by slice: the tile's records are grouped by bin in shared
// memory, one global reservation per (tile, bin), then copied out so that consecutive threads write
// consecutive records of a run.
__global__ void __launch_bounds__(256) k_pair_scatter(const uint2* __restrict__ pairs, u64 npairs, u32 nbins, u32 shift,
                                                      unsigned long long* __restrict__ cursor,
                                                      uint2* __restrict__ rec_e, u32* __restrict__ rec_dest) {
  __shared__ u32 h[SLICE_MAX_BINS];          // count, then the bin's first slot in the tile
  __shared__ unsigned long long start[SLICE_MAX_BINS];
  __shared__ uint2 s_e[PAIR_TILE];
  __shared__ unsigned short s_bin[PAIR_TILE];
  __shared__ unsigned short s_src[PAIR_TILE];  // index of the pair inside the tile
  __shared__ u32 s_warp[8];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (u32 b = tid; b < nbins; b += 256) h[b] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * PAIR_TILE;
  uint2 e[PAIR_TILE / 256];
  u32 rank[PAIR_TILE / 256], bin[PAIR_TILE / 256];
#pragma unroll
  for (int i = 0; i < PAIR_TILE / 256; i++) {
    const u64 j = base + i * 256 + tid;
    if (j < npairs) {
      e[i] = __ldg(pairs + j);
      bin[i] = pair_bin(e[i].x, e[i].y, shift);
      rank[i] = atomicAdd(&h[bin[i]], 1u);
    }
  }
  __syncthreads();
  // exclusive scan of the bin counts (each thread owns a contiguous strip) + one global reservation per bin
  {
    const u32 per = (nbins + 255) / 256;
    const u32 b0 = tid * per, b1 = min(nbins, b0 + per);
    u32 sum = 0;
    for (u32 b = b0; b < b1; b++) sum += h[b];
    u32 x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (u32)o) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    u32 run = x - sum;
    for (u32 w = 0; w < wid; w++) run += s_warp[w];
    for (u32 b = b0; b < b1; b++) {
      const u32 c = h[b];
      if (c) start[b] = atomicAdd(&cursor[b], (unsigned long long)c) - run;  // global slot of tile slot 0 of this bin
      h[b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < PAIR_TILE / 256; i++) {
    const u64 j = base + i * 256 + tid;
    if (j < npairs) {
      const u32 slot = h[bin[i]] + rank[i];
      // the operand that selected the slice goes first (identity + P = P + identity)
      s_e[slot] = (e[i].x != ENTRY_NONE) ? e[i] : make_uint2(e[i].y, e[i].x);
      s_bin[slot] = (unsigned short)bin[i];
      s_src[slot] = (unsigned short)(i * 256 + tid);
    }
  }
  __syncthreads();
  const u32 cnt = (u32)min((u64)PAIR_TILE, npairs - base);
  for (u32 slot = tid; slot < cnt; slot += 256) {
    const u64 k = start[s_bin[slot]] + slot;
    rec_e[k] = s_e[slot];
    rec_dest[k] = (u32)(base + s_src[slot]);
  }
}

// Thread `tid` of CTA `b` owns the records b * 128 * B + i * 128 + tid, i < B (coalesced record reads);
// its denominators are chained in that order by both kernels.
template <int TREE_B>
__global__ void __launch_bounds__(TREE_THREADS)
k_tree_fwd_sliced(const uint2* __restrict__ rec_e, u64 npairs, const G1Affine* __restrict__ bases,
                  Fq* __restrict__ T0, Fq* __restrict__ pre) {
  const u64 k0 = (u64)blockIdx.x * (TREE_THREADS * TREE_B) + threadIdx.x;
  Fq run = Fq::one();
  bool any = false;
#pragma unroll 1
  for (int i = 0; i < TREE_B; i++) {
    const u64 k = k0 + (u64)i * TREE_THREADS;
    if (k >= npairs) break;
    st_fq_stream(pre + k, run);
    const uint2 e = ldg_rec(rec_e + k);
    Operand<true> P, Q;
    P.open_entry(bases, e.x);
    Q.open_entry(bases, e.y);
    Fq d;
    if (pair_denominator<true>(P, Q, d) != PAIR_TRIVIAL) {
      run = any ? fp_mul(run, d) : d;
      any = true;
    }
  }
  st_fq(T0 + (u64)blockIdx.x * TREE_THREADS + threadIdx.x, run);
}

template <int TREE_B>
__global__ void __launch_bounds__(TREE_THREADS, 4)
k_tree_bwd_sliced(const uint2* __restrict__ rec_e, const u32* __restrict__ rec_dest, u64 npairs,
                  const G1Affine* __restrict__ bases, const Fq* __restrict__ T0inv, const Fq* __restrict__ pre_all,
                  Fq* __restrict__ out_x, Fq* __restrict__ out_y, u32 ostride) {
  const u64 k0 = (u64)blockIdx.x * (TREE_THREADS * TREE_B) + threadIdx.x;
  if (k0 >= npairs) return;
  const int cnt = (int)min((u64)TREE_B, (npairs - k0 + TREE_THREADS - 1) / TREE_THREADS);
  Fq inv = ldg_fq(T0inv + (u64)blockIdx.x * TREE_THREADS + threadIdx.x);
#pragma unroll 1
  for (int i = cnt - 1; i >= 0; i--) {
    const u64 k = k0 + (u64)i * TREE_THREADS;
    const uint2 e = ldg_rec(rec_e + k);
    const u32 dest = __ldg(rec_dest + k);
    Operand<true> P, Q;
    P.open_entry(bases, e.x);
    Q.open_entry(bases, e.y);
    Fq d;
    const int kind = pair_denominator<true>(P, Q, d);
    const G1Affine r = pair_sum<true>(P, Q, kind, d, inv, pre_all + k);
    st_fq_stream(out_x + (u64)dest * ostride, r.x);
    st_fq_stream(out_y + (u64)dest * ostride, r.y);
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
                       const Fq* pre, Fq* out_x, Fq* out_y, u32 ostride) {
  if (r0) k_tree_bwd<true, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0inv, pre, out_x, out_y, ostride);
  else k_tree_bwd<false, B><<<grid, TREE_THREADS, 0, st>>>(src, npairs, T0inv, pre, out_x, out_y, ostride);
}

// Pairs per thread (one shared denominator product per thread): 8, 16 or 32.  More pairs per thread = fewer
// products for the shared inversion (3 / B per pair), but also B x more pairs in flight per resident thread, and
// the in-flight pairs of round 0 gather from the window tables: with few columns (few pairs in total) a launch
// with B = 32 has a quarter of the table in flight at once and the gathers miss the L2.  MEASURED (2^20 rows,
// profiles/r02g_pairs_per_thread.txt): 16 columns 43.40 / 43.49 / 44.28 ms for B = 32 / 16 / 8, 8 columns 22.61 /
// 22.42 / 22.68, 4 columns 12.24 / 12.16 / 12.11, 2 columns 7.31 / 7.16 / 7.08.  EON_TREE_B / EON_TREE_SLICED_B force.
static int tree_b_policy(u64 pairs_round0) { return pairs_round0 < (48ull << 20) ? 8 : (pairs_round0 < (100ull << 20) ? 16 : 32); }
static int tree_b(u64 pairs_round0) {
  static int b = -1;
  if (b < 0) {
    const char* e = getenv("EON_TREE_B");
    b = e ? atoi(e) : 0;
    if (b != 8 && b != 16 && b != 32) b = 0;
  }
  return b ? b : tree_b_policy(pairs_round0);
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Slice schedule (see k_tree_fwd_sliced): worth it when the bases do not fit the L2 and every base is gathered
// several times per launch (one use per column with window tables).  eon_msm_set_slice_schedule /
// EON_TREE_SLICED = 0 / 1 force it off / on; EON_SLICE_SHIFT = log2 of the points per slice.
SlicePlan msm_slice_plan(const eon_ctx* ctx, u64 nbases, u64 total_slots, u32 rounds) {
  static const int sliced_env = env_int("EON_TREE_SLICED", -1);
  static const int shift_env = env_int("EON_SLICE_SHIFT", 19);
  SlicePlan p;
  p.shift = (shift_env >= 12 && shift_env <= 24) ? (u32)shift_env : 19u;
  const u64 nbins = ((nbases ? nbases - 1 : 0) >> p.shift) + 1;
  p.nbins = (u32)std::min<u64>(nbins, 0xffffffffull);
  const int mode = ctx->msm_slice_mode >= 0 ? ctx->msm_slice_mode : sliced_env;
  p.on = rounds > 0 && nbins <= SLICE_MAX_BINS && total_slots / 2 < 0xffffffffull;
  if (mode == 0) p.on = false;
  else if (mode != 1)
    p.on = p.on && nbins <= SLICE_ORDER_MAX && nbases * sizeof(G1Affine) > (96ull << 20) && total_slots >= 2 * nbases;
  return p;
}

template <int B>
static void launch_sliced(bool fwd, unsigned grid, cudaStream_t st, const uint2* rec_e, const u32* rec_dest, u64 npairs,
                          const G1Affine* bases, Fq* T0, Fq* pre, Fq* out_x, Fq* out_y, u32 ostride) {
  if (fwd) k_tree_fwd_sliced<B><<<grid, TREE_THREADS, 0, st>>>(rec_e, npairs, bases, T0, pre);
  else k_tree_bwd_sliced<B><<<grid, TREE_THREADS, 0, st>>>(rec_e, rec_dest, npairs, bases, T0, pre, out_x, out_y, ostride);
}

// Where the pair records of the slice schedule live: round 1's output buffer, which is idle during round 0 (12 bytes
// per pair against the 32 bytes per pair of that buffer).  Also called by the fused sort, which writes them.
int msm_tree_records(eon_ctx* ctx, u64 total_slots, uint2** rec_e, u32** rec_dest) {
  const u64 m0 = total_slots / 2;
  void* bufB;
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_B, (m0 / 2 + 1) * sizeof(G1Affine), &bufB));
  *rec_e = (uint2*)bufB;
  *rec_dest = (u32*)(*rec_e + m0);
  return EON_OK;
}

int msm_tree_rounds(eon_ctx* ctx, const G1Affine* d_bases, const SlicePlan& plan, const u32* d_entries,
                    u64 total_slots, u32 rounds, const G1Affine** out_pts, bool records_ready) {
  const u64 TREE_B = (u64)tree_b(total_slots / 2);
  if (rounds == 0 || (total_slots & ((1ull << rounds) - 1)))
    return fail(ctx, EON_ERR_BAD_ARG, "msm_tree_rounds: slot count not aligned to 2^rounds");
  cudaStream_t st = ctx->stream;
  const u64 m0 = total_slots / 2;  // pairs of round 0
  void *bufA, *bufB, *bufT;
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_A, m0 * sizeof(G1Affine), &bufA));
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_B, (m0 / 2 + 1) * sizeof(G1Affine), &bufB));
  // level sizes of round 0 (the largest round): n0 = ceil(m0 / B), n(l+1) = ceil(n(l) / 8)
  static const int sliced_b_env = env_int("EON_TREE_SLICED_B", 0);
  const u64 SLICED_B = (sliced_b_env == 8 || sliced_b_env == 16 || sliced_b_env == 32) ? (u64)sliced_b_env
                                                                                         : (u64)tree_b_policy(m0);
  const bool sliced = plan.on;
  const u64 nbins = plan.nbins;
  const u64 n0_sliced = ((m0 + TREE_THREADS * SLICED_B - 1) / (TREE_THREADS * SLICED_B)) * TREE_THREADS;
  u64 tcap = 0;
  for (u64 n = std::max((m0 + TREE_B - 1) / TREE_B, sliced ? n0_sliced : (u64)0);; n = (n + 7) / 8) {
    tcap += n;
    if (n <= TREE_TOP) break;
  }
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_T, tcap * sizeof(Fq), &bufT));
  void* bufP;
  EON_TRY(scratch_get(ctx, SC_MSM_TREE_P, m0 * sizeof(Fq), &bufP));

  TreeSrc src;
  src.pts_x = src.pts_y = nullptr;
  src.entries = d_entries;
  src.bases = d_bases;
  u64 npairs = m0;
  Fq* outs[2] = {(Fq*)bufA, (Fq*)bufB};  // a round's sums: npairs x coordinates, then npairs y coordinates
  // pair records of the slice schedule live in round 1's output buffer, which is idle during round 0
  // (12 bytes per pair against the 32 bytes per pair of that buffer)
  uint2* rec_e = (uint2*)bufB;
  u32* rec_dest = (u32*)(rec_e + m0);
  for (u32 r = 0; r < rounds; r++) {
    // intermediate rounds write the coordinates apart (the next k_tree_fwd reads x only); the last round writes
    // points (x, y interleaved): the finisher reads both coordinates of consecutive sums (measured: with the
    // coordinates apart there too it lost 0.6 ms of the 0.9 ms the forward passes gained)
    const bool last = r + 1 == rounds;
    const u32 ostride = last ? 2 : 1;
    Fq* out_x = outs[r & 1];
    Fq* out_y = last ? out_x + 1 : out_x + npairs;
    const bool sl = sliced && r == 0;
    std::vector<std::pair<Fq*, u64>> lv;  // (array, size) per level
    Fq* T = (Fq*)bufT;
    for (u64 n = sl ? n0_sliced : (npairs + TREE_B - 1) / TREE_B;; n = (n + 7) / 8) {
      lv.push_back(std::make_pair(T, n));
      T += n;
      if (n <= TREE_TOP) break;
    }
    const unsigned g0 = grid_for(lv[0].second);
    phase_begin(ctx, PH_MSM_TREE_FWD);
    if (sl && records_ready) {  // the sort wrote the records (msm_sort_place)
      if (SLICED_B == 8) launch_sliced<8>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else if (SLICED_B == 16) launch_sliced<16>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else launch_sliced<32>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
    } else if (sl) {
      void* bufC;
      EON_TRY(scratch_get(ctx, SC_MSM_SLICE, 2 * SLICE_MAX_BINS * sizeof(unsigned long long), &bufC));
      unsigned long long* counts = (unsigned long long*)bufC;
      unsigned long long* cursor = counts + SLICE_MAX_BINS;
      const unsigned gp = (unsigned)((npairs + PAIR_TILE - 1) / PAIR_TILE);
      const uint2* pairs = reinterpret_cast<const uint2*>(d_entries);
      EON_CUDA(ctx, cudaMemsetAsync(counts, 0, SLICE_MAX_BINS * sizeof(unsigned long long), st));
      k_pair_hist<<<gp, 256, 0, st>>>(pairs, npairs, (u32)nbins, plan.shift, counts);
      EON_LAUNCHED(ctx);
      k_pair_scan<<<1, 32, 0, st>>>(counts, (u32)nbins, cursor);
      EON_LAUNCHED(ctx);
      k_pair_scatter<<<gp, 256, 0, st>>>(pairs, npairs, (u32)nbins, plan.shift, cursor, rec_e, rec_dest);
      EON_LAUNCHED(ctx);
      if (SLICED_B == 8) launch_sliced<8>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else if (SLICED_B == 16) launch_sliced<16>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else launch_sliced<32>(true, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
    } else if (TREE_B == 8) launch_fwd<8>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    else if (TREE_B == 16) launch_fwd<16>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    else launch_fwd<32>(r == 0, g0, st, src, npairs, lv[0].first, (Fq*)bufP);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_TREE_FWD);
    phase_begin(ctx, PH_MSM_TREE_INV);
    {
      cudaStream_t ts = tiny_begin(ctx);  // the product tree: a dependent chain of small launches
      for (size_t l = 0; l + 1 < lv.size(); l++) {
        k_tree_up<<<grid_for(lv[l + 1].second), TREE_THREADS, 0, ts>>>(lv[l].first, lv[l].second, lv[l + 1].first);
        EON_LAUNCHED(ctx);
      }
      k_tree_top<<<(unsigned)((lv.back().second + 31) / 32), 32, 0, ts>>>(lv.back().first, lv.back().second);
      EON_LAUNCHED(ctx);
      for (size_t l = lv.size() - 1; l-- > 0;) {
        k_tree_down<<<grid_for(lv[l + 1].second), TREE_THREADS, 0, ts>>>(lv[l].first, lv[l].second, lv[l + 1].first);
        EON_LAUNCHED(ctx);
      }
      tiny_end(ctx);
    }
    phase_end(ctx, PH_MSM_TREE_INV);
    phase_begin(ctx, PH_MSM_TREE_BWD);
    if (sl) {
      if (SLICED_B == 8) launch_sliced<8>(false, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else if (SLICED_B == 16) launch_sliced<16>(false, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
      else launch_sliced<32>(false, g0, st, rec_e, rec_dest, npairs, d_bases, lv[0].first, (Fq*)bufP, out_x, out_y, ostride);
    } else if (TREE_B == 8) launch_bwd<8>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out_x, out_y, ostride);
    else if (TREE_B == 16) launch_bwd<16>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out_x, out_y, ostride);
    else launch_bwd<32>(r == 0, g0, st, src, npairs, lv[0].first, (const Fq*)bufP, out_x, out_y, ostride);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_TREE_BWD);
    if (r == 0 && ctx->after_round0) {  // work the caller wants queued behind round 0 (its gathers live in the L2)
      std::function<int()> f = std::move(ctx->after_round0);
      ctx->after_round0 = nullptr;
      EON_TRY(f());
    }
    src.pts_x = out_x;
    src.pts_y = out_y;
    npairs /= 2;
  }
  *out_pts = reinterpret_cast<const G1Affine*>(src.pts_x);  // the last round wrote points
  return EON_OK;
}

}  // namespace eon
