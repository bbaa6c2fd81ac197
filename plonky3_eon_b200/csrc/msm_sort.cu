// Two-pass counting sort of the MSM entries by bucket (sm_100a).
//
// The Pippenger bucket phase behind G1::multi_exp (bn254/src/curve.rs:158-180) needs, for every bucket
// of every (column, bucket set) segment, the list of bases that fall into it.  The one-pass scatter
// (k_msm_scatter in msm.cu) does one returning global atomic and one isolated 4-byte store per entry
// — 2.5e8 of each at 2^20 x 16 — and is bound by L2 atomic/partial-sector traffic (6.1 ms, 8 GB of
// DRAM traffic for a 1 GB result).  Here the same permutation is done MSD-radix style:
//
//   k_sort_coarse  a CTA takes a tile of 1024 scalars of one column, histograms their digits by the
//                  HIGH bucket bits (bin = bucket >> 7) in shared memory, reserves one contiguous run
//                  per bin with a single global atomic, groups the entries by bin in shared memory and
//                  copies (payload, low bucket bits) out coalesced: 512 global atomics per tile and
//                  full-sector stores instead of one atomic + one isolated store per entry.
//   k_sort_fine    one CTA per (segment, bin): 128 bucket cursors in shared memory (from the aligned
//                  bucket starts); the bin's entries stream in coalesced, are placed in a shared-memory
//                  image of the bin's ~120 KB window of the final array, and the window (padding =
//                  ENTRY_NONE) is written out coalesced.
// (An SM retires scattered 4-byte stores at about one sector per clock whatever L2 merges afterwards;
// that, not DRAM, bounded the one-pass scatter and a first version of this file without staging.)
//
// The bucket histogram comes out of the sort itself: k_msm_hist's 2.5e8 global atomics (1.2 ms at 2^20 x 16)
// are replaced by
//   k_bin_hist     the same tiles counted by BIN in shared memory (512 global atomics per tile), then a
//                  per-segment scan gives every bin an exact, unpadded run of the temporary array;
//   k_sort_count   after the coarse pass, one CTA per bin counts its entries by bucket (shared memory,
//                  coalesced 2-byte key reads) and writes the bucket histogram without atomics;
// the aligned scan (k_msm_scan) then runs on that histogram as before and k_sort_fine places.
#include <stdlib.h>

#include "msm.cuh"

namespace eon {

constexpr u32 SORT_THREADS = 256;
constexpr u32 SORT_FINE_THREADS = 1024;
constexpr u32 SORT_MAX_TILE_BINS = 4096;   // nsets * nbins: shared histogram / offsets / run bases of a tile
constexpr u32 SORT_WIN_CAP = 40960;        // entry slots of a bin window staged in shared memory (160 KiB) at 2^7 buckets per bin
#ifndef EON_SORT_FB_DEFAULT
#define EON_SORT_FB_DEFAULT 7
#endif
constexpr size_t SORT_COARSE_SMEM = 160 * 1024;

// counter[key] += 1 for every active lane, returning each lane's rank.  Plain shared-memory atomics
// (MATCH.ANY-style aggregation costs more than the few conflicts it saves on uniform scalars), except
// when the whole warp hits ONE counter — all-equal / tiny scalars — where one atomic serves the warp.
__device__ __forceinline__ u32 smem_rank(u32* counter, u32 key) {
  const u32 mask = __activemask();
  const u32 leader = __ffs(mask) - 1;
  const u32 k0 = __shfl_sync(mask, key, leader);
  if (__all_sync(mask, key == k0)) {
    const u32 lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(counter + key, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
  }
  return atomicAdd(counter + key, 1u);
}

// The same for a CONVERGED warp (every lane calls, `on` says whether the lane has an entry): full-mask collectives,
// no divergence bookkeeping (the divergent form above compiles to ~100 instructions per call site with its
// BRA.DIV / WARPSYNC paths, and the coarse pass has 60 call sites per thread).
__device__ __forceinline__ u32 smem_rank_conv(u32* counter, u32 key, bool on) {
  const u32 act = __ballot_sync(0xffffffffu, on);
  if (act == 0) return 0;
  const u32 leader = __ffs(act) - 1;
  const u32 k0 = __shfl_sync(0xffffffffu, key, leader);
  if (__all_sync(0xffffffffu, !on || key == k0)) {
    const u32 lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(counter + k0, __popc(act));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(act & ((1u << lane) - 1));
  }
  return on ? atomicAdd(counter + key, 1u) : 0u;
}

// Bin histogram.  grid: ncols * ceil(n / tile) blocks, column index fastest (as the coarse pass).
// bin_count[(col * nsets + set) * nbins + k] += entries of the tile whose bucket >> fb == k
template <bool MERGED, int C>
__global__ void __launch_bounds__(SORT_THREADS)
k_bin_hist(const Fr* __restrict__ scalars, size_t n, size_t ld, u32 ncols, MsmShape sh, u32 nbins, u32 fb, u32 tile,
           u32* __restrict__ bin_count) {
  extern __shared__ u32 smem[];
  const u32 tile_bins = sh.nsets * nbins;
  const u32 tid = threadIdx.x;
  const u32 col = blockIdx.x % ncols;
  const size_t i0 = (size_t)(blockIdx.x / ncols) * tile;
  for (u32 k = tid; k < tile_bins; k += SORT_THREADS) smem[k] = 0;
  __syncthreads();
  // the tile's scalars of this thread are loaded together (tile <= 4 * SORT_THREADS), then digitised
  Fr sc[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const u32 t = tid + q * SORT_THREADS;
    const size_t i = i0 + t;
    sc[q] = (t < tile && i < n) ? load_scalar(scalars, i, ld, col) : Fr::zero();  // zero scalar: no digits
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    u32 kk[8];
    scalar_digits_source(kk, sc[q], sh);
    for_each_digit_c<C>(kk, sh, [&](u32 w, int d) {
      u32 b = (u32)(d < 0 ? -d : d) - 1;
      atomicAdd(&smem[(MERGED ? 0 : w * nbins) + (b >> fb)], 1u);
    });
  }
  __syncthreads();
  u32* dst = bin_count + (size_t)col * sh.nsets * nbins;
  for (u32 k = tid; k < tile_bins; k += SORT_THREADS)
    if (smem[k]) atomicAdd(dst + k, smem[k]);
}

// Per segment (one block): tmp_start[k] = region_cursor[k] = exclusive scan of bin_count over the segment's
// bins (nbins <= 1024).
__global__ void __launch_bounds__(1024) k_bin_scan(const u32* __restrict__ bin_count, u32 nbins,
                                                   u32* __restrict__ tmp_start, u32* __restrict__ region_cursor) {
  __shared__ u32 warp_sums[32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const size_t base = (size_t)blockIdx.x * nbins;
  const u32 v = tid < nbins ? bin_count[base + tid] : 0;
  u32 x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (u32)o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  u32 wbase = 0;
  for (u32 j = 0; j < wid; j++) wbase += warp_sums[j];
  if (tid < nbins) {
    tmp_start[base + tid] = wbase + x - v;
    region_cursor[base + tid] = wbase + x - v;
  }
}

// Bucket histogram of one bin from the coarse pass's output.  grid: nseg * nbins blocks.
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_count(const unsigned short* __restrict__ tmp_key, const u32* __restrict__ tmp_start,
             const u32* __restrict__ region_cursor, u32 NB, u32 nbins, u32 fb, u64 seg_cap, u32* __restrict__ hist) {
  __shared__ u32 s_cnt[1u << 7];
  const u32 fine = 1u << fb;
  const size_t seg = blockIdx.x / nbins;
  const u32 k = blockIdx.x % nbins;
  const u32 tid = threadIdx.x;
  for (u32 f = tid; f < fine; f += SORT_THREADS) s_cnt[f] = 0;
  __syncthreads();
  const u32 begin = tmp_start[seg * nbins + k], end = region_cursor[seg * nbins + k];
  // 8 keys per 16-byte load over the aligned interior of the run, single keys at its two ends
  const u64 g0 = seg * seg_cap + begin, g1 = seg * seg_cap + end;
  u64 a0 = (g0 + 7) & ~7ull, a1 = g1 & ~7ull;
  if (a0 > g1) a0 = g1;
  if (a1 < a0) a1 = a0;
  for (u64 i = g0 + tid; i < a0; i += SORT_THREADS) atomicAdd(&s_cnt[tmp_key[i]], 1u);
  for (u64 i = a1 + tid; i < g1; i += SORT_THREADS) atomicAdd(&s_cnt[tmp_key[i]], 1u);
  const uint4* kv = reinterpret_cast<const uint4*>(tmp_key + a0);
  const u64 nv = (a1 - a0) >> 3;
  for (u64 v = tid; v < nv; v += SORT_THREADS) {
    const uint4 q = __ldg(kv + v);
    const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      atomicAdd(&s_cnt[w[j] & 0xffffu], 1u);
      atomicAdd(&s_cnt[w[j] >> 16], 1u);
    }
  }
  __syncthreads();
  u32* h = hist + seg * NB + ((size_t)k << fb);
  for (u32 f = tid; f < fine; f += SORT_THREADS) h[f] = s_cnt[f];
}

// Coarse pass.  grid: ncols * ceil(n / tile) blocks, column index fastest (cf. k_msm_hist).
// Shared memory: cnt[tile_bins] | off[tile_bins + 1] | gbase[tile_bins] | stage_pay[tile * W] u32 |
// stage_key[tile * W] u16.  The tile's entries are grouped by bin in shared memory first, so the copy
// to the temporary array is coalesced (consecutive threads -> consecutive slots of a run).
// MERGED (window tables, the commit path) with NB <= 2^16: the staged key is the bucket index itself, so the
// copy-out is flat -- one thread per staged slot, the bin read back from the key -- instead of one warp per bin
// (~30 entries per run at 2^20 x 16: 70 instructions per run, 36 % of the kernel's instructions in ncu's source
// view, profiles/r02q_sort_kernels.txt).
template <bool MERGED, int C>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_coarse(const Fr* __restrict__ scalars, size_t n, size_t ld, u32 ncols, MsmShape sh, u32 nbins, u32 fb, u32 tile,
              u32* __restrict__ region_cursor, u32* __restrict__ tmp_pay, unsigned short* __restrict__ tmp_key) {
  extern __shared__ u32 smem[];
  const u32 tile_bins = sh.nsets * nbins;
  u32* s_cnt = smem;
  u32* s_off = s_cnt + tile_bins;        // tile_bins + 1
  u32* s_gbase = s_off + tile_bins + 1;
  u32* s_pay = s_gbase + tile_bins;
  unsigned short* s_key = reinterpret_cast<unsigned short*>(s_pay + (size_t)tile * sh.W);
  __shared__ u32 s_warp[SORT_THREADS / 32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 col = blockIdx.x % ncols;
  const size_t i0 = (size_t)(blockIdx.x / ncols) * tile;
  const size_t seg0 = (size_t)col * sh.nsets;
  const u32 fmask = (1u << fb) - 1;

  for (u32 k = tid; k < tile_bins; k += SORT_THREADS) s_cnt[k] = 0;
  __syncthreads();
  // canonical scalars of this thread (tile <= 4 * SORT_THREADS), kept in registers for both sweeps
  u32 kk[4][8];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const u32 t = tid + q * SORT_THREADS;
    const size_t i = i0 + t;
    if (t < tile && i < n) {
      scalar_digits_source(kk[q], load_scalar(scalars, i, ld, col), sh);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) kk[q][j] = 0;  // zero scalar: no digits
    }
  }
  // histogram of the tile by bin
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if ((u32)q * SORT_THREADS >= tile) continue;  // (uniform) a tile of 3 scalars per thread has no fourth sweep
    for_each_digit_c<C>(kk[q], sh, [&](u32 w, int d) {
      u32 b = (u32)(d < 0 ? -d : d) - 1;
      atomicAdd(&s_cnt[(MERGED ? 0 : w * nbins) + (b >> fb)], 1u);
    });
  }
  __syncthreads();
  // exclusive scan over the bins (each thread owns a contiguous strip), one global reservation per bin
  {
    const u32 per = (tile_bins + SORT_THREADS - 1) / SORT_THREADS;
    const u32 k0 = tid * per, k1 = min(tile_bins, k0 + per);
    u32 sum = 0;
    for (u32 k = k0; k < k1; k++) sum += s_cnt[k];
    u32 x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (u32)o) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    u32 wbase = 0;
    for (u32 j = 0; j < wid; j++) wbase += s_warp[j];
    u32 run = wbase + x - sum;
    for (u32 k = k0; k < k1; k++) {
      u32 c = s_cnt[k];
      s_off[k] = run;
      run += c;
      // first slot of the bin's run in the temporary array, minus the bin's first staged slot
      s_gbase[k] = (c ? atomicAdd(&region_cursor[seg0 * nbins + k], c) : 0) - run + c;
      s_cnt[k] = 0;
    }
    if (tid == SORT_THREADS - 1) s_off[tile_bins] = run;
  }
  __syncthreads();
  // group the entries by bin in shared memory
  const bool flat = MERGED && sh.NB <= 65536u;  // the bucket index fits the 16-bit staged key
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if ((u32)q * SORT_THREADS >= tile) continue;  // (uniform) a tile of 3 scalars per thread has no fourth sweep
    const u32 i = (u32)(i0 + tid + q * SORT_THREADS);
    const u32 base0 = MERGED ? (u32)sh.base_first + i : i;  // entry = base index: + w * tab_stride with tables
    const u32 stride = MERGED ? (u32)sh.tab_stride : 0u;
    for_each_digit_c<C, C != 0>(kk[q], sh, [&](u32 w, int d) {
      const bool on = d != 0;  // (always true on the generic walk, which skips zero digits)
      u32 b = on ? (u32)(d < 0 ? -d : d) - 1 : 0u;
      u32 bin = (MERGED ? 0 : w * nbins) + (b >> fb);
      u32 rank;
      if constexpr (C != 0) rank = smem_rank_conv(s_cnt, bin, on);
      else rank = smem_rank(s_cnt, bin);
      if (on) {
        u32 slot = s_off[bin] + rank;
        s_pay[slot] = (base0 + w * stride) | ((u32)d & SIGN_BIT);
        s_key[slot] = (unsigned short)(flat ? b : (b & fmask));  // flat: the bucket index itself
      }
    });
  }
  __syncthreads();
  if (flat) {
    // coalesced copy-out, flat over the staged slots (grouped by bin, so consecutive threads write consecutive
    // slots of a run); the bin is read back from the staged bucket index
    const u32 total = s_off[tile_bins];
    const size_t dst0 = seg0 * sh.seg_cap;
    for (u32 idx = tid; idx < total; idx += SORT_THREADS) {
      const u32 b = s_key[idx];
      const size_t dst = dst0 + (u32)(s_gbase[b >> fb] + idx);
      tmp_pay[dst] = s_pay[idx];
      tmp_key[dst] = (unsigned short)(b & fmask);
    }
  } else {
    // one warp per bin: lanes -> consecutive slots of the bin's run (bin and low bucket bits together need 17 bits)
    for (u32 bin = wid; bin < tile_bins; bin += SORT_THREADS / 32) {
      const u32 o0 = s_off[bin], o1 = s_off[bin + 1];
      if (o0 == o1) continue;
      const size_t dst = (seg0 + (MERGED ? 0 : bin / nbins)) * sh.seg_cap;
      for (u32 idx = o0 + lane; idx < o1; idx += 32) {
        tmp_pay[dst + (u32)(s_gbase[bin] + idx)] = s_pay[idx];
        tmp_key[dst + (u32)(s_gbase[bin] + idx)] = s_key[idx];
      }
    }
  }
}

// Fine pass.  grid: nseg * nbins blocks.  Bin k of segment seg holds tmp slots [starts[k << fb],
// region_cursor) and owns the window [starts[k << fb], starts[(k + 1) << fb]) of the final array.
// Windows up to SORT_WIN_CAP slots are built in shared memory (padding = ENTRY_NONE) and written out
// coalesced; larger ones (skewed scalars, very large inputs) are scattered directly.
// ORDERED: inside every bucket the entries are additionally ordered by table slice (base index >> slice_shift,
// ns slices), so that round 0 of the pairwise rounds pairs operands of the same slice (msm_tree.cu): a first
// sweep counts per (bucket, slice), the second places.  Any order inside a bucket gives the same bucket sum.
template <bool ORDERED>
__global__ void __launch_bounds__(SORT_FINE_THREADS)
k_sort_fine(const u32* __restrict__ tmp_pay, const unsigned short* __restrict__ tmp_key,
            const u32* __restrict__ tmp_start, const u32* __restrict__ region_cursor, const u32* __restrict__ starts,
            const u32* __restrict__ seg_total,
            u32 NB, u32 nbins, u32 fb, u64 seg_cap, u32 slice_shift, u32 ns, u32 win_cap, u32* __restrict__ ends,
            u32* __restrict__ entries) {
  extern __shared__ u32 smem[];
  const u32 fine = 1u << fb;
  u32* s_cur = smem;          // fine
  u32* s_win = smem + fine;   // win_cap
  u32* s_cnt2 = s_win + win_cap;  // ORDERED: fine * ns (bucket, slice) counters, then cursors
  const size_t seg = blockIdx.x / nbins;
  const u32 k = blockIdx.x % nbins;
  const size_t g0 = seg * NB + ((size_t)k << fb);
  const size_t off = seg * seg_cap;
  const u32 tid = threadIdx.x;
  const u32 begin = starts[g0];  // first slot of the bin's window of the final array
  const u32 tbegin = tmp_start[seg * nbins + k], tend = region_cursor[seg * nbins + k];  // its run of the temporary array
  for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) s_cur[f] = starts[g0 + f];
  // window end: the next bin's first start, or (last bin) the aligned end of the segment
  const u32 wend = (k + 1 < nbins) ? starts[g0 + fine] : seg_total[seg];
  const u32 wlen = wend - begin;
  const bool staged = wlen <= win_cap;
  if (staged)
    for (u32 j = tid; j < wlen; j += SORT_FINE_THREADS) s_win[j] = ENTRY_NONE;
  // The run is read FINE_U entries per thread at a time: all loads of a batch are in flight before the first
  // shared-memory atomic needs its value (one load -> one atomic per iteration left the kernel waiting on DRAM
  // latency: long-scoreboard stall 10 per issue in ncu, profiles/r02q_sort_kernels.txt).
  constexpr int FINE_U = 8;
  if (ORDERED && staged) {
    for (u32 j = tid; j < fine * ns; j += SORT_FINE_THREADS) s_cnt2[j] = 0;
    __syncthreads();
    for (u32 i0 = tbegin; i0 < tend; i0 += SORT_FINE_THREADS * FINE_U) {
      u32 key[FINE_U], pay[FINE_U];
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          key[u] = tmp_key[off + i];
          pay[u] = tmp_pay[off + i];
        }
      }
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          const u32 sl = min((pay[u] & ~SIGN_BIT) >> slice_shift, ns - 1);
          atomicAdd(&s_cnt2[key[u] * ns + sl], 1u);
        }
      }
    }
    __syncthreads();
    for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) {
      u32 run = s_cur[f];
      for (u32 q = 0; q < ns; q++) {
        const u32 c = s_cnt2[f * ns + q];
        s_cnt2[f * ns + q] = run;
        run += c;
      }
      s_cur[f] = run;  // the bucket's end
    }
    __syncthreads();
    for (u32 i0 = tbegin; i0 < tend; i0 += SORT_FINE_THREADS * FINE_U) {
      u32 key[FINE_U], pay[FINE_U];
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          key[u] = tmp_key[off + i];
          pay[u] = tmp_pay[off + i];
        }
      }
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          const u32 sl = min((pay[u] & ~SIGN_BIT) >> slice_shift, ns - 1);
          s_win[smem_rank(s_cnt2, key[u] * ns + sl) - begin] = pay[u];
        }
      }
    }
  } else {
    __syncthreads();
    for (u32 i0 = tbegin; i0 < tend; i0 += SORT_FINE_THREADS * FINE_U) {
      u32 key[FINE_U], pay[FINE_U];
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          key[u] = tmp_key[off + i];
          pay[u] = tmp_pay[off + i];
        }
      }
#pragma unroll
      for (int u = 0; u < FINE_U; u++) {
        const u32 i = i0 + u * SORT_FINE_THREADS + tid;
        if (i < tend) {
          const u32 pos = smem_rank(s_cur, key[u]);
          if (staged) s_win[pos - begin] = pay[u];
          else entries[off + pos] = pay[u];
        }
      }
    }
  }
  __syncthreads();
  if (staged)
    for (u32 j = tid; j < wlen; j += SORT_FINE_THREADS) entries[off + begin + j] = s_win[j];
  for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) ends[g0 + f] = s_cur[f];
}


// ---- fused form for the slice schedule: count by (bucket, slice) first, then place AND emit the pair records ------
// With the slice schedule (msm_tree.cu) round 0 does not read the sorted entries at all: it walks PAIR RECORDS
// (entry0, entry1, destination) grouped by the table slice of the first operand.  The unfused flow wrote the entries
// (k_sort_fine<true>: a counting sweep per (bucket, slice) and a placement sweep, 1 GB out), then k_pair_hist /
// k_pair_scatter read them back to count and group the pairs (2 GB in, 1.5 GB out, a memset of 1 GB before): 2.7 ms
// of a 2^20 x 16 commit.  Here
//   k_sort_count2   one CTA per bin counts its run by (bucket, slice) ONCE (small shared memory, full occupancy),
//                   writes the bucket histogram, the count matrix, and -- positions inside a bucket are ordered by
//                   slice and buckets start on even slots, so the slice of every pair's first operand follows from
//                   the counts alone -- the number of pairs per slice (one global add per slice and CTA);
//   (aligned scan of the histogram, exclusive scan of the slice totals)
//   k_sort_place2   one CTA per bin: cursors from the matrix, ONE sweep places the run in the shared-memory window,
//                   then every pair of the window goes straight to its place in the slice's record list (rank =
//                   prefix over the bin's buckets + position inside the bucket: no atomics per record);
//   k_pair_tail     the unused slots behind every segment's last bucket, as (NONE, NONE) records.
// All-padding pairs form one extra list behind the last slice (they only write the identity).  The entry array is
// not written (nothing reads it: rounds >= 1 and the finisher read round outputs), except by a bin whose window does
// not fit the shared-memory image (skewed scalars), which builds it in global memory and reads it back.
constexpr int SORT_U = 8;

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_count2(const u32* __restrict__ tmp_pay, const unsigned short* __restrict__ tmp_key,
              const u32* __restrict__ tmp_start, const u32* __restrict__ region_cursor, u32 NB, u32 nbins, u32 fb,
              u64 seg_cap, u32 slice_shift, u32 ns, u32* __restrict__ hist, u32* __restrict__ mat,
              unsigned long long* __restrict__ counts) {
  extern __shared__ u32 s_cnt2[];  // fine * ns
  __shared__ u32 s_pairs[SLICE_ORDER_MAX];
  const u32 fine = 1u << fb;
  const size_t seg = blockIdx.x / nbins;
  const u32 k = blockIdx.x % nbins;
  const u32 tid = threadIdx.x;
  for (u32 j = tid; j < fine * ns; j += SORT_THREADS) s_cnt2[j] = 0;
  if (tid < SLICE_ORDER_MAX) s_pairs[tid] = 0;
  __syncthreads();
  const u32 tbegin = tmp_start[seg * nbins + k], tend = region_cursor[seg * nbins + k];
  const size_t off = seg * seg_cap;
  for (u32 i0 = tbegin; i0 < tend; i0 += SORT_THREADS * SORT_U) {
    u32 key[SORT_U], pay[SORT_U];
#pragma unroll
    for (int u = 0; u < SORT_U; u++) {
      const u32 i = i0 + u * SORT_THREADS + tid;
      if (i < tend) {
        key[u] = tmp_key[off + i];
        pay[u] = tmp_pay[off + i];
      }
    }
#pragma unroll
    for (int u = 0; u < SORT_U; u++) {
      const u32 i = i0 + u * SORT_THREADS + tid;
      if (i < tend) atomicAdd(&s_cnt2[key[u] * ns + min((pay[u] & ~SIGN_BIT) >> slice_shift, ns - 1)], 1u);
    }
  }
  __syncthreads();
  u32* h = hist + seg * NB + ((size_t)k << fb);
  for (u32 f = tid; f < fine; f += SORT_THREADS) {
    u32 run = 0;  // position inside the bucket (its first slot is even)
    for (u32 q = 0; q < ns; q++) {
      const u32 c = s_cnt2[f * ns + q];
      const u32 pq = ((run + c + 1) >> 1) - ((run + 1) >> 1);  // even positions in [run, run + c)
      if (pq) atomicAdd(&s_pairs[q], pq);
      run += c;
    }
    h[f] = run;
  }
  u32* m = mat + ((seg * nbins + k) << fb) * ns;
  for (u32 j = tid; j < fine * ns; j += SORT_THREADS) m[j] = s_cnt2[j];
  __syncthreads();
  if (tid < ns && s_pairs[tid]) atomicAdd(&counts[tid], (unsigned long long)s_pairs[tid]);
}

// cursor[q] = sum of counts[< q] for q <= ns: the lists of the ns slices, then the all-padding list
__global__ void k_pair_scan2(const unsigned long long* __restrict__ counts, u32 ns, unsigned long long* __restrict__ cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long acc = 0;
    for (u32 q = 0; q < ns; q++) {
      cursor[q] = acc;
      acc += counts[q];
    }
    cursor[ns] = acc;
  }
}

__global__ void __launch_bounds__(SORT_FINE_THREADS)
k_sort_place2(const u32* __restrict__ tmp_pay, const unsigned short* __restrict__ tmp_key,
              const u32* __restrict__ tmp_start, const u32* __restrict__ region_cursor, const u32* __restrict__ starts,
              const u32* __restrict__ seg_total, const u32* __restrict__ mat, u32 NB, u32 nbins, u32 fb, u64 seg_cap,
              u32 slice_shift, u32 ns, u32 win_cap, unsigned long long* __restrict__ cursor, u32* __restrict__ ends,
              u32* entries, uint2* __restrict__ rec_e, u32* __restrict__ rec_dest) {
  extern __shared__ u32 smem[];
  const u32 fine = 1u << fb;
  u32* s_start = smem;                    // fine: first slot of every bucket
  u32* s_slots = s_start + fine;          // fine: slots the bucket owns (its entries rounded up)
  u32* s_cur = s_slots + fine;            // fine: end of the bucket's entries
  u32* s_c2 = s_cur + fine;               // fine * ns: (bucket, slice) counts, then cursors
  u32* s_rb = s_c2 + fine * ns;           // fine * (ns + 1): pairs per (bucket, list), then their exclusive prefix
  u32* s_gb = s_rb + fine * (ns + 1);     // 2 * (ns + 1): first record of this bin in every list (64 bits)
  u32* s_win = s_gb + 2 * (SLICE_ORDER_MAX + 1);
  const size_t seg = blockIdx.x / nbins;
  const u32 k = blockIdx.x % nbins;
  const size_t g0 = seg * NB + ((size_t)k << fb);
  const size_t off = seg * seg_cap;
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 begin = starts[g0];
  const u32 wend = (k + 1 < nbins) ? starts[g0 + fine] : seg_total[seg];
  const u32 wlen = wend - begin;
  const bool staged = wlen <= win_cap;
  const u32 tbegin = tmp_start[seg * nbins + k], tend = region_cursor[seg * nbins + k];
  const u32* m = mat + ((seg * nbins + k) << fb) * ns;
  for (u32 j = tid; j < fine * ns; j += SORT_FINE_THREADS) s_c2[j] = m[j];
  for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) {
    const u32 st = starts[g0 + f];
    s_start[f] = st;
    s_slots[f] = ((f + 1 < fine) ? starts[g0 + f + 1] : wend) - st;
  }
  if (staged)
    for (u32 j = tid; j < wlen; j += SORT_FINE_THREADS) s_win[j] = ENTRY_NONE;
  __syncthreads();
  for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) {
    const u32 st = s_start[f];
    u32 run = 0;
    for (u32 q = 0; q < ns; q++) {
      const u32 c = s_c2[f * ns + q];
      s_rb[f * (ns + 1) + q] = ((run + c + 1) >> 1) - ((run + 1) >> 1);
      s_c2[f * ns + q] = st + run;
      run += c;
    }
    s_cur[f] = st + run;
    s_rb[f * (ns + 1) + ns] = (s_slots[f] >> 1) - ((run + 1) >> 1);  // all-padding pairs of the bucket
    if (!staged)  // no shared-memory image: the bucket's padding slots are written here
      for (u32 j = st + run; j < st + s_slots[f]; j++) entries[off + j] = ENTRY_NONE;
  }
  __syncthreads();
  // exclusive prefix over the bin's buckets, one warp per list; lane l owns buckets [l * per, (l + 1) * per)
  for (u32 q = wid; q <= ns; q += SORT_FINE_THREADS / 32) {
    const u32 per = (fine + 31) / 32;
    u32 sum = 0;
    for (u32 i = 0; i < per; i++) {
      const u32 f = lane * per + i;
      if (f < fine) sum += s_rb[f * (ns + 1) + q];
    }
    u32 x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (u32)o) x += y;
    }
    const u32 total = __shfl_sync(0xffffffffu, x, 31);
    u32 run = x - sum;
    for (u32 i = 0; i < per; i++) {
      const u32 f = lane * per + i;
      if (f < fine) {
        const u32 c = s_rb[f * (ns + 1) + q];
        s_rb[f * (ns + 1) + q] = run;
        run += c;
      }
    }
    if (lane == 0) {
      const unsigned long long gb = total ? atomicAdd(&cursor[q], (unsigned long long)total) : 0ull;
      s_gb[2 * q] = (u32)gb;
      s_gb[2 * q + 1] = (u32)(gb >> 32);
    }
  }
  __syncthreads();
  // placement: one sweep over the bin's run of the temporary array (warps stay converged: lanes past the end of the
  // run take part in the ranking collectives with on = false)
  constexpr int PLACE_U = 4;  // (8 spills at the 64 registers 1024 threads leave)
  for (u32 i0 = tbegin; i0 < tend; i0 += SORT_FINE_THREADS * PLACE_U) {
    u32 key[PLACE_U], pay[PLACE_U];
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) {
      const u32 i = i0 + u * SORT_FINE_THREADS + tid;
      key[u] = 0;
      pay[u] = 0;
      if (i < tend) {
        key[u] = tmp_key[off + i];
        pay[u] = tmp_pay[off + i];
      }
    }
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) {
      const bool on = i0 + u * SORT_FINE_THREADS + tid < tend;
      const u32 sl = min((pay[u] & ~SIGN_BIT) >> slice_shift, ns - 1);
      const u32 pos = smem_rank_conv(s_c2, key[u] * ns + sl, on);
      if (on) {
        if (staged) s_win[pos - begin] = pay[u];
        else entries[off + pos] = pay[u];
      }
    }
  }
  __syncthreads();  // (also orders this CTA's global writes before its reads below)
  // pair records of the window, one warp per bucket: lanes take the bucket's pairs in turn (no search for the
  // bucket of a slot; the per-slot form with its 7-step search was half of the kernel's instructions in ncu)
  const volatile u32* gwin = entries + off + begin;
  for (u32 f = wid; f < fine; f += SORT_FINE_THREADS / 32) {
    const u32 st = s_start[f];
    const u32 npair = s_slots[f] >> 1;
    const u32 nent = s_cur[f] - st;
    const u32 w0 = st - begin;  // the bucket's first slot inside the window
    for (u32 pp = lane; pp < npair; pp += 32) {
      const u32 e0 = staged ? s_win[w0 + 2 * pp] : gwin[w0 + 2 * pp];
      const u32 e1 = staged ? s_win[w0 + 2 * pp + 1] : gwin[w0 + 2 * pp + 1];
      u32 q, first;
      if (e0 != ENTRY_NONE) {
        q = min((e0 & ~SIGN_BIT) >> slice_shift, ns - 1);
        first = (q ? s_c2[f * ns + q - 1] : st) - st;  // cursors have advanced to the end of their runs
      } else {
        q = ns;
        first = nent;
      }
      const u32 r = s_rb[f * (ns + 1) + q] + pp - ((first + 1) >> 1);
      const unsigned long long kk = ((unsigned long long)s_gb[2 * q + 1] << 32 | s_gb[2 * q]) + r;
      rec_e[kk] = make_uint2(e0, e1);
      rec_dest[kk] = (u32)((off + st + 2 * pp) >> 1);
    }
  }
  for (u32 f = tid; f < fine; f += SORT_FINE_THREADS) ends[g0 + f] = s_cur[f];
}

// (NONE, NONE) records for the slots behind the last bucket of every segment.  grid: (chunks, nseg)
__global__ void __launch_bounds__(256)
k_pair_tail(const u32* __restrict__ seg_total, u64 seg_cap, u32 ns, unsigned long long* __restrict__ cursor,
            uint2* __restrict__ rec_e, u32* __restrict__ rec_dest) {
  __shared__ unsigned long long s_base;
  const size_t seg = blockIdx.y;
  const u64 t0 = seg_total[seg] >> 1, t1 = seg_cap >> 1;
  for (u64 c0 = t0 + (u64)blockIdx.x * 1024; c0 < t1; c0 += (u64)gridDim.x * 1024) {
    const u32 cnt = (u32)min((u64)1024, t1 - c0);
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(&cursor[ns], (unsigned long long)cnt);
    __syncthreads();
    for (u32 i = threadIdx.x; i < cnt; i += 256) {
      rec_e[s_base + i] = make_uint2(ENTRY_NONE, ENTRY_NONE);
      rec_dest[s_base + i] = (u32)(((seg * seg_cap) >> 1) + c0 + i);
    }
  }
}

static bool sort_fused_env() {
  static const int v = getenv("EON_SORT_FUSED") ? atoi(getenv("EON_SORT_FUSED")) : 1;
  return v != 0;
}
// the entries of a bucket are ordered by slice (and the pair records come out of the sort) when the slice schedule
// is on and its slices fit the (bucket, slice) counters of a bin
static bool sort_ordered(const SlicePlan& plan) {
  static const int order_env = getenv("EON_SORT_ORDERED") ? atoi(getenv("EON_SORT_ORDERED")) : 1;
  return plan.on && plan.nbins > 1 && plan.nbins <= SLICE_ORDER_MAX && order_env;
}

// Geometry shared by the two halves of the sort.
struct SortGeom {
  u32 fb, nbins, tile_bins, tile, win_cap;
  size_t smem_coarse, smem_fine;
};
static bool sort_geom(const MsmShape& sh, SortGeom* g) {
  // Worth it only while a tile still fills runs of tens of entries per bin (<= 1024 bins per tile) and a
  // bin's window fits the shared-memory image; otherwise (2^22+ points per column at c = 20, or many
  // bucket sets per column) the one-pass scatter is faster (measured: 2^24 x 1, 4.2 vs 14.6 ms).
  if (sh.NB < 256) return false;
  // buckets per bin: 2^7 with a 160 KiB window image (one fine / placement CTA per SM) or 2^6 with 80 KiB (two per
  // SM: the phases of one -- load, count, place, write -- overlap with the other's).  EON_SORT_FB selects.
  static const int fb_env = getenv("EON_SORT_FB") ? atoi(getenv("EON_SORT_FB")) : EON_SORT_FB_DEFAULT;
  g->fb = (fb_env == 6 && (sh.nsets * (sh.NB >> 6)) <= 1024) ? 6 : 7;
  g->win_cap = SORT_WIN_CAP >> (7 - g->fb);
  g->nbins = sh.NB >> g->fb;
  g->tile_bins = sh.nsets * g->nbins;
  if (g->tile_bins > 1024) return false;
  if (sh.seg_cap / g->nbins > ((size_t)g->win_cap * 9) / 10) return false;
  // scalars per coarse tile: stage (6 bytes per entry, up to W entries per scalar) within the smem budget
  // k_sort_coarse keeps up to 4 scalars per thread in registers; 3 per thread (75 KiB of staging at 15 windows)
  // lets three CTAs share an SM: measured 1.86 -> 1.59 ms at 2^20 x 16 (profiles/r02y_sort_tile.txt)
  u32 tile = 3 * SORT_THREADS;
  if (const char* e = getenv("EON_SORT_TILE")) tile = (u32)atoi(e);
  if (tile > 4 * SORT_THREADS || tile < 64) tile = 4 * SORT_THREADS;
  const size_t fixed = ((size_t)3 * g->tile_bins + 1) * sizeof(u32);
  while (tile > 64 && fixed + (size_t)tile * sh.W * 6 > SORT_COARSE_SMEM) tile >>= 1;
  if (fixed + (size_t)tile * sh.W * 6 > SORT_COARSE_SMEM) return false;
  g->tile = tile;
  g->smem_coarse = fixed + (size_t)tile * sh.W * 6 + 16;
  g->smem_fine = ((size_t)(1u << g->fb) + g->win_cap) * sizeof(u32);
  return true;
}
static size_t place2_smem(u32 fb, u32 ns, u32 win_cap) {
  const size_t fine = (size_t)1 << fb;
  return (3 * fine + fine * ns + fine * (ns + 1) + 2 * (SLICE_ORDER_MAX + 1) + win_cap) * sizeof(u32);
}

// Segments [seg0, seg0 + ncols * nsets) of a batch of nseg_total segments; the array arguments are those of the WHOLE
// batch.  *deferred = true: the group has been counted by (bucket, slice) only, msm_sort_place (once, after the last
// group) places all segments and emits the pair records.
int msm_sort_entries(eon_ctx* ctx, const Fr* d_scalars, size_t n, size_t ncols, size_t ld, const MsmShape& sh,
                     const SlicePlan& plan, size_t seg0, size_t nseg_total, u32* d_hist_all, u32* d_seg_total_all,
                     u32* d_cur_all, u32* d_entries_all, bool* deferred) {
  *deferred = false;
  SortGeom G;
  if (!sort_geom(sh, &G)) return 1;
  const u32 fb = G.fb, nbins = G.nbins, tile_bins = G.tile_bins, tile = G.tile;
  const size_t smem_coarse = G.smem_coarse, smem_fine = G.smem_fine;
  const size_t nseg = ncols * sh.nsets;
  const size_t total_bins = nseg * nbins;
  const size_t tiles = (n + tile - 1) / tile;
  const u32 tile_hist = 4 * SORT_THREADS;  // the bin histogram has no staging: always 4 scalars per thread
  const size_t tiles_hist = (n + tile_hist - 1) / tile_hist;
  if (tiles * ncols > 0x7fffffffull || nseg_total * nbins > 0x7fffffffull) return 1;
  const bool ordered = sort_ordered(plan);
  const bool fused = ordered && sort_fused_env() && ctx->msm_sort_mode != 2 && ((u64)nseg_total * sh.seg_cap) / 2 < 0xffffffffull;
  if (!ctx->sort_attr_set) {  // per context: the attribute belongs to the context's device
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_coarse<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SORT_COARSE_SMEM + 16)));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_coarse<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SORT_COARSE_SMEM + 16)));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_coarse<true, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SORT_COARSE_SMEM + 16)));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_coarse<false, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SORT_COARSE_SMEM + 16)));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_fine<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(((size_t)(1u << 12) + SORT_WIN_CAP) * sizeof(u32))));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_fine<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(((size_t)(1u << 7) * (1 + SLICE_ORDER_MAX) + SORT_WIN_CAP) * sizeof(u32))));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_sort_place2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)place2_smem(7, SLICE_ORDER_MAX, SORT_WIN_CAP)));
    ctx->sort_attr_set = true;
  }
  void *p_reg, *p_pay, *p_key;
  // (sized for the whole batch: with the placement deferred, every group's temporary runs are still needed then)
  // per bin: count | start of its run in the temporary array | cursor (ends as the end of that run)
  const size_t all_bins = nseg_total * nbins;
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_REGION, 3 * all_bins * sizeof(u32), &p_reg));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_PAY, nseg_total * sh.seg_cap * sizeof(u32), &p_pay));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_KEY, nseg_total * sh.seg_cap * sizeof(unsigned short), &p_key));
  u32* bin_count = (u32*)p_reg + seg0 * nbins;
  u32* tmp_start = (u32*)p_reg + all_bins + seg0 * nbins;
  u32* region_cursor = (u32*)p_reg + 2 * all_bins + seg0 * nbins;
  u32* tmp_pay = (u32*)p_pay + seg0 * sh.seg_cap;
  unsigned short* tmp_key = (unsigned short*)p_key + seg0 * sh.seg_cap;
  u32* d_hist = d_hist_all + seg0 * sh.NB;
  u32* d_cur = d_cur_all + seg0 * sh.NB;
  u32* d_seg_total = d_seg_total_all + seg0;
  u32* d_entries = d_entries_all + seg0 * sh.seg_cap;
  void *p_mat = nullptr, *p_slice = nullptr;
  if (fused) {
    EON_TRY(scratch_get(ctx, SC_MSM_SORT_MAT, ((all_bins << fb) * plan.nbins) * sizeof(u32), &p_mat));
    EON_TRY(scratch_get(ctx, SC_MSM_SLICE, 2 * (1024 + 8) * sizeof(unsigned long long), &p_slice));
  }
  cudaStream_t st = ctx->stream;
  const unsigned grid_tiles = (unsigned)(tiles * ncols);

  // window size known at compile time for the shapes the cost model picks (c = 17 up to 2^22 points)
  const bool c17 = sh.c == 17 && sh.W == 15;
  phase_begin(ctx, PH_MSM_DIGITS);
  EON_CUDA(ctx, cudaMemsetAsync(bin_count, 0, total_bins * sizeof(u32), st));
#define EON_SORT_LAUNCH(K, GRID, SMEM, ...)                                                                    \
  do {                                                                                                         \
    if (sh.merged && c17) K<true, 17><<<GRID, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                          \
    else if (sh.merged) K<true, 0><<<GRID, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                             \
    else if (c17) K<false, 17><<<GRID, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                                 \
    else K<false, 0><<<GRID, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                                           \
  } while (0)
  EON_SORT_LAUNCH(k_bin_hist, (unsigned)(tiles_hist * ncols), tile_bins * sizeof(u32), d_scalars, n, ld, (u32)ncols, sh,
                  nbins, fb, tile_hist, bin_count);
  EON_LAUNCHED(ctx);
  k_bin_scan<<<(unsigned)nseg, 1024, 0, st>>>(bin_count, nbins, tmp_start, region_cursor);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_DIGITS);

  phase_begin(ctx, PH_MSM_SCATTER);
  if (sh.rounds && !fused)  // unused slots (bucket padding, segment tails) must read as ENTRY_NONE
    EON_CUDA(ctx, cudaMemsetAsync(d_entries, 0xff, nseg * sh.seg_cap * sizeof(u32), st));
  EON_SORT_LAUNCH(k_sort_coarse, grid_tiles, smem_coarse, d_scalars, n, ld, (u32)ncols, sh, nbins, fb, tile,
                  region_cursor, tmp_pay, tmp_key);
#undef EON_SORT_LAUNCH
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCATTER);

  if (fused) {
    phase_begin(ctx, PH_MSM_SCAN);
    u32* mat = (u32*)p_mat + ((seg0 * nbins) << fb) * plan.nbins;
    k_sort_count2<<<(unsigned)total_bins, SORT_THREADS, ((size_t)plan.nbins << fb) * sizeof(u32), st>>>(
        tmp_pay, tmp_key, tmp_start, region_cursor, sh.NB, nbins, fb, sh.seg_cap, plan.shift, plan.nbins, d_hist, mat,
        (unsigned long long*)p_slice);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_SCAN);
    *deferred = true;
    return EON_OK;
  }

  phase_begin(ctx, PH_MSM_SCAN);
  k_sort_count<<<(unsigned)total_bins, SORT_THREADS, 0, st>>>(tmp_key, tmp_start, region_cursor, sh.NB, nbins, fb,
                                                              sh.seg_cap, d_hist);
  EON_LAUNCHED(ctx);
  EON_TRY(msm_scan_run(ctx, d_hist, d_cur, sh.NB, 1u << sh.rounds, d_seg_total, nseg));
  phase_end(ctx, PH_MSM_SCAN);

  phase_begin(ctx, PH_MSM_SCATTER);
  if (ordered)
    k_sort_fine<true><<<(unsigned)total_bins, SORT_FINE_THREADS, smem_fine + ((size_t)plan.nbins << fb) * sizeof(u32), st>>>(
        tmp_pay, tmp_key, tmp_start, region_cursor, d_hist, d_seg_total, sh.NB, nbins, fb, sh.seg_cap, plan.shift,
        plan.nbins, G.win_cap, d_cur, d_entries);
  else
    k_sort_fine<false><<<(unsigned)total_bins, SORT_FINE_THREADS, smem_fine, st>>>(
        tmp_pay, tmp_key, tmp_start, region_cursor, d_hist, d_seg_total, sh.NB, nbins, fb, sh.seg_cap, 0u, 1u,
        G.win_cap, d_cur, d_entries);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCATTER);
  return EON_OK;
}

// Zeroes the per-slice pair counters of a batch whose sort may take the fused form (before its first group).
int msm_sort_begin(eon_ctx* ctx) {
  void* p_slice = nullptr;
  EON_TRY(scratch_get(ctx, SC_MSM_SLICE, 2 * (1024 + 8) * sizeof(unsigned long long), &p_slice));
  EON_CUDA(ctx, cudaMemsetAsync(p_slice, 0, (1024 + 8) * sizeof(unsigned long long), ctx->stream));
  return EON_OK;
}

// Second half of the fused form, once per batch after every group has been counted: aligned scan of the bucket
// histogram, scan of the slice totals, placement + pair records of all segments, tail records.  rec_e / rec_dest:
// the record arrays msm_tree_rounds walks (total_slots / 2 records).
int msm_sort_place(eon_ctx* ctx, const MsmShape& sh, const SlicePlan& plan, size_t nseg_total, u32* d_hist,
                   u32* d_seg_total, u32* d_cur, u32* d_entries, uint2* rec_e, u32* rec_dest) {
  SortGeom G;
  if (!sort_geom(sh, &G)) return fail(ctx, EON_ERR_BAD_ARG, "msm_sort_place: shape without a sort geometry");
  const u32 fb = G.fb, nbins = G.nbins;
  const size_t all_bins = nseg_total * nbins;
  void *p_reg, *p_pay, *p_key, *p_mat, *p_slice;
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_REGION, 3 * all_bins * sizeof(u32), &p_reg));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_PAY, nseg_total * sh.seg_cap * sizeof(u32), &p_pay));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_KEY, nseg_total * sh.seg_cap * sizeof(unsigned short), &p_key));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_MAT, ((all_bins << fb) * plan.nbins) * sizeof(u32), &p_mat));
  EON_TRY(scratch_get(ctx, SC_MSM_SLICE, 2 * (1024 + 8) * sizeof(unsigned long long), &p_slice));
  const u32* tmp_start = (const u32*)p_reg + all_bins;
  const u32* region_cursor = (const u32*)p_reg + 2 * all_bins;
  unsigned long long* counts = (unsigned long long*)p_slice;
  unsigned long long* cursor = counts + 1024 + 8;
  cudaStream_t st = ctx->stream;
  phase_begin(ctx, PH_MSM_SCAN);
  EON_TRY(msm_scan_run(ctx, d_hist, d_cur, sh.NB, 1u << sh.rounds, d_seg_total, nseg_total));
  k_pair_scan2<<<1, 32, 0, st>>>(counts, plan.nbins, cursor);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCAN);
  phase_begin(ctx, PH_MSM_SCATTER);
  k_sort_place2<<<(unsigned)all_bins, SORT_FINE_THREADS, place2_smem(fb, plan.nbins, G.win_cap), st>>>(
      (const u32*)p_pay, (const unsigned short*)p_key, tmp_start, region_cursor, d_hist, d_seg_total, (const u32*)p_mat,
      sh.NB, nbins, fb, sh.seg_cap, plan.shift, plan.nbins, G.win_cap, cursor, d_cur, d_entries, rec_e, rec_dest);
  EON_LAUNCHED(ctx);
  k_pair_tail<<<dim3(64, (unsigned)nseg_total), 256, 0, st>>>(d_seg_total, sh.seg_cap, plan.nbins, cursor, rec_e, rec_dest);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCATTER);
  return EON_OK;
}

}  // namespace eon
