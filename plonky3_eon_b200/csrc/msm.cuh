// Shared declarations of the MSM translation units (msm.cu: Pippenger pipeline; msm_tree.cu: batched-
// affine pairwise rounds).
#pragma once
#include "common.cuh"

namespace eon {

constexpr int MSM_THREADS = 256;
constexpr u32 SIGN_BIT = 0x80000000u;

struct MsmShape {
  u32 c;        // window bits (2..20)
  u32 W;        // windows = ceil(256 / c)
  u32 NB;       // buckets per bucket set = 2^(c-1)
  u32 nsets;    // bucket sets per column: W (plain) or 1 (merged: windows share one set because
                // window t reads the precomputed table 2^(c t) * P_i instead of P_i)
  u32 merged;
  u32 chunk;    // buckets per reduction chunk
  u32 nchunks;  // NB / chunk
  u64 tab_stride;  // merged: points per table level
  u64 base_first;  // merged: index of this MSM's point 0 inside a table level
  u64 seg_cap;     // entry slots per segment: n (plain) or n * W (merged), plus alignment padding
  u32 mont_digits; // merged: the window tables hold R^-1 * 2^(c t) * P_i, so the digits are taken from the Montgomery
                   // limbs of a scalar as they are (sum (s R) (R^-1 P) = sum s P): no from-Montgomery product per scalar
  u32 rounds;      // batched-affine pairwise rounds before the serial XYZZ finisher (0 = none);
                   // bucket starts are aligned to 2^rounds entry slots
};

constexpr u32 ENTRY_NONE = 0xffffffffu;  // unused entry slot (reads as the identity point)

// R rounds of out[j] = in[2j] + in[2j+1] over the flat slot array (all segments), affine with one
// shared inversion per round.  Round 0 reads entries/bases, later rounds read points.  On return
// *out_pts points to the final array of (total_slots >> rounds) affine points.
// Slice schedule of round 0 (msm_tree.cu): the base table is cut into slices of 2^shift points and the pairs
// are processed slice by slice, so their gathers hit the L2.  on = false: slot order.
struct SlicePlan {
  bool on;
  u32 shift;
  u32 nbins;
};
// the sort can order the entries of a bucket by slice only up to this many slices ((bucket, slice) counters of
// a bin in shared memory); without that order the schedule gains next to nothing (one operand of every pair
// still comes from a random slice and evicts the resident one), so the automatic policy stops here
constexpr u32 SLICE_ORDER_MAX = 32;
// nbases = number of points addressable through d_bases (all window-table levels).
SlicePlan msm_slice_plan(const eon_ctx* ctx, u64 nbases, u64 total_slots, u32 rounds);
// records_ready: the pair records of the slice schedule have been written by the sort (msm_sort_place) into the
// arrays msm_tree_records names; d_entries is then not read.
int msm_tree_rounds(eon_ctx* ctx, const G1Affine* d_bases, const SlicePlan& plan, const u32* d_entries,
                    u64 total_slots, u32 rounds, const G1Affine** out_pts, bool records_ready = false);
int msm_tree_records(eon_ctx* ctx, u64 total_slots, uint2** rec_e, u32** rec_dest);


// Counting sort of the (point, window) entries by bucket in coalesced passes (msm_sort.cu): bin histogram,
// coarse scatter by the high bucket bits into a temporary array, bucket histogram of every bin, aligned scan,
// per-bin fine placement.  Does the whole digits / scan / scatter part of an MSM batch on its own.
// Out: hist[] = bucket starts (aligned exclusive scan), seg_total[] = aligned slots used per segment,
// cur[g] = starts[g] + count(g), entries[] sorted by bucket (unused slots = ENTRY_NONE when sh.rounds > 0).
// Returns EON_OK, or a positive value WITHOUT having launched anything if the shape is not supported (the
// caller falls back to the global histogram + one-pass atomic scatter).
// plan.on: the entries of every bucket are additionally ordered by table slice (so that round 0 pairs operands
// of the same slice); any order inside a bucket gives the same sums.
// The array arguments are those of the whole batch (nseg_total segments); the call handles the segments
// [seg0, seg0 + ncols * nsets).  With the slice schedule the sort takes a fused form: *deferred = true means the group
// has been sorted coarsely and counted, and msm_sort_place (once per batch, after the last group) places everything
// and writes the pair records round 0 walks, instead of the entry array.  msm_sort_begin: once per batch, before
// the first group.
int msm_sort_begin(eon_ctx* ctx);
int msm_sort_entries(eon_ctx* ctx, const Fr* d_scalars, size_t n, size_t ncols, size_t ld, const MsmShape& sh,
                     const SlicePlan& plan, size_t seg0, size_t nseg_total, u32* d_hist, u32* d_seg_total, u32* d_cur,
                     u32* d_entries, bool* deferred);
int msm_sort_place(eon_ctx* ctx, const MsmShape& sh, const SlicePlan& plan, size_t nseg_total, u32* d_hist,
                   u32* d_seg_total, u32* d_cur, u32* d_entries, uint2* rec_e, u32* rec_dest);
// aligned exclusive scan of the bucket histogram, one block per segment (k_msm_scan, msm.cu)
int msm_scan_run(eon_ctx* ctx, u32* d_hist, u32* d_cur, u32 NB, u32 align, u32* d_seg_total, size_t nseg);

// One batch of an MSM (msm.cu): workspace, shape and slice plan, so that the sort can be queued per column group
// before the pairwise rounds / finisher / reduction run over all columns of the batch.
struct MsmBatch {
  const G1Affine* bases = nullptr;
  size_t n = 0, ncols = 0;
  MsmShape sh;
  SlicePlan plan;
  u32 chunk_min = 0, max_tasks = 0;
  bool place_deferred = false;  // set by msm_batch_sort: msm_batch_finish runs msm_sort_place first
  void *p_hist = nullptr, *p_cur = nullptr, *p_ent = nullptr, *p_bkt = nullptr, *p_tasks = nullptr, *p_tpart = nullptr,
       *p_part = nullptr, *p_seg = nullptr, *p_misc = nullptr, *p_ord = nullptr, *p_segtot = nullptr;
};
int msm_batch_setup(eon_ctx* ctx, const G1Affine* d_bases, size_t n, size_t ncols, const MsmShape& sh, MsmBatch* B);
// columns [c0, c0 + nc) of the batch: d_scalars points at column c0 (row pitch ld)
int msm_batch_sort(eon_ctx* ctx, MsmBatch& B, const Fr* d_scalars, size_t ld, size_t c0, size_t nc);
int msm_batch_finish(eon_ctx* ctx, const MsmBatch& B, G1Affine* d_out);
// An MSM over the resident SRS whose sort is queued group by group (host-buffer commit: group g is sorted while
// group g + 1 crosses PCIe).  begin returns 1 without queueing anything when the MSM would not run as one unsplit
// batch; the caller then falls back to msm_run.  Then msm_batch_sort per group and msm_batch_finish.
int msm_stream_begin(eon_ctx* ctx, const G1Affine* d_bases, size_t n, size_t ncols, MsmBatch* B);

#if defined(__CUDACC__)
// ---- 1. scalar -> signed window digits --------------------------------------------------------
// s * 1 * R^-1 gives the canonical integer (the reference hands halo2curves Montgomery scalars,
// bn254/src/curve.rs:173-174).  Digits d_w in [-2^(c-1), 2^(c-1)) with sum d_w 2^(c w) = scalar;
// The top window never wraps: W*c >= 255 leaves it at most c - 1 scalar bits, so raw + carry <=
// 2^(c-1), which still has a bucket (index 2^(c-1) - 1).  f(w, d) for d != 0.
// k: canonical scalar limbs.  The limbs are consumed through a 64-bit bit buffer in a fully unrolled
// loop, so k[] stays in registers (indexing it by a runtime window position would spill it to local
// memory).
template <class F>
__device__ __forceinline__ void for_each_digit_canonical(const u32 (&k)[8], const MsmShape& sh, F f) {
  const u32 c = sh.c, W = sh.W;
  const u32 mask = (1u << c) - 1;
  const u32 half = 1u << (c - 1);
  u64 buf = 0;
  u32 nbits = 0, w = 0, carry = 0;
  auto emit = [&](u32 raw) {
    raw += carry;
    int d;
    if (raw >= half && w + 1 < W) {
      d = (int)raw - (int)(1u << c);
      carry = 1;
    } else {
      d = (int)raw;
      carry = 0;
    }
    if (d != 0) f(w, d);
    w++;
  };
#pragma unroll
  for (int j = 0; j < 8; j++) {
    buf |= (u64)k[j] << nbits;
    nbits += 32;
    while (nbits >= c && w < W) {
      emit((u32)buf & mask);
      buf >>= c;
      nbits -= c;
    }
  }
  while (w < W) {  // top window(s): remaining bits, zero-extended
    emit((u32)buf & mask);
    buf >>= c;
  }
}

// The same digits with the window size known at compile time (C = sh.c, W = ceil(255 / C) = sh.W): every digit is
// one funnel shift and a mask at a fixed limb / bit position, where the generic walk above is a data-dependent
// loop over a 64-bit bit buffer (~33 instructions per digit in the SASS of the sort kernels, which visit every
// digit three times).  C = 0 selects the generic walk.
// ALL (C != 0 only): f is also called for zero digits (d == 0), so a warp stays converged through the whole walk
// and f may use full-mask warp collectives.
template <int C, bool ALL = false, class F>
__device__ __forceinline__ void for_each_digit_c(const u32 (&k)[8], const MsmShape& sh, F f) {
  if constexpr (C == 0) {
    for_each_digit_canonical(k, sh, f);
  } else {
    constexpr u32 W = (255 + C - 1) / C;
    constexpr u32 mask = (1u << C) - 1;
    constexpr u32 half = 1u << (C - 1);
    u32 carry = 0;
#pragma unroll
    for (u32 w = 0; w < W; w++) {
      const u32 o = C * w, j = o >> 5, sft = o & 31;  // compile-time after unrolling; j <= 7 because o < 255
      const u32 lo = k[j];
      const u32 hi = (j + 1 < 8) ? k[(j + 1) & 7] : 0u;
      u32 raw = (sft ? __funnelshift_r(lo, hi, sft) : lo) & mask;
      raw += carry;
      int d;
      if (raw >= half && w + 1 < W) {
        d = (int)raw - (int)(1u << C);
        carry = 1;
      } else {
        d = (int)raw;
        carry = 0;
      }
      if (ALL || d != 0) f(w, d);
    }
  }
}

// the integer whose digits index the buckets: the canonical scalar, or its Montgomery limbs with pre-scaled tables
__device__ __forceinline__ void scalar_digits_source(u32 (&k)[8], const Fr& s, const MsmShape& sh) {
  if (sh.mont_digits) {
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = s.v[j];
  } else {
    fp_from_mont(k, s);
  }
}

template <class F>
__device__ __forceinline__ void for_each_digit(const Fr& s, const MsmShape& sh, F f) {
  u32 k[8];
  scalar_digits_source(k, s, sh);
  for_each_digit_canonical(k, sh, f);
}

__device__ __forceinline__ Fr load_scalar(const Fr* __restrict__ scalars, size_t i, size_t ld, u32 col) {
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + i * ld + col);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  Fr s;
  s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
  s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
  return s;
}

#endif  // __CUDACC__

}  // namespace eon
