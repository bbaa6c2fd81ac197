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
constexpr u32 SORT_WIN_CAP = 40960;        // entry slots of a bin window staged in shared memory (160 KiB)
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
    fp_from_mont(kk, sc[q]);
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
      fp_from_mont(kk[q], load_scalar(scalars, i, ld, col));
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) kk[q][j] = 0;  // zero scalar: no digits
    }
  }
  // histogram of the tile by bin
#pragma unroll
  for (int q = 0; q < 4; q++) {
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
            u32 NB, u32 nbins, u32 fb, u64 seg_cap, u32 slice_shift, u32 ns, u32* __restrict__ ends,
            u32* __restrict__ entries) {
  extern __shared__ u32 smem[];
  const u32 fine = 1u << fb;
  u32* s_cur = smem;          // fine
  u32* s_win = smem + fine;   // SORT_WIN_CAP
  u32* s_cnt2 = s_win + SORT_WIN_CAP;  // ORDERED: fine * ns (bucket, slice) counters, then cursors
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
  const bool staged = wlen <= SORT_WIN_CAP;
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


int msm_sort_entries(eon_ctx* ctx, const Fr* d_scalars, size_t n, size_t ncols, size_t ld, const MsmShape& sh,
                     const SlicePlan& plan, u32* d_hist, u32* d_seg_total, u32* d_cur, u32* d_entries) {
  // Worth it only while a tile still fills runs of tens of entries per bin (<= 1024 bins per tile) and a
  // bin's window fits the shared-memory image; otherwise (2^22+ points per column at c = 20, or many
  // bucket sets per column) the one-pass scatter is faster (measured: 2^24 x 1, 4.2 vs 14.6 ms).
  if (sh.NB < 256) return 1;
  const u32 fb = 7;
  const u32 nbins = sh.NB >> fb;
  const u32 tile_bins = sh.nsets * nbins;
  if (tile_bins > 1024) return 1;
  if (sh.seg_cap / nbins > (SORT_WIN_CAP * 9) / 10) return 1;
  // scalars per coarse tile: stage (6 bytes per entry, up to W entries per scalar) within the smem budget
  u32 tile = 4 * SORT_THREADS;  // k_sort_coarse keeps 4 scalars per thread in registers
  if (const char* e = getenv("EON_SORT_TILE")) tile = (u32)atoi(e);
  if (tile > 4 * SORT_THREADS || tile < 64) tile = 4 * SORT_THREADS;
  const size_t fixed = ((size_t)3 * tile_bins + 1) * sizeof(u32);
  while (tile > 64 && fixed + (size_t)tile * sh.W * 6 > SORT_COARSE_SMEM) tile >>= 1;
  if (fixed + (size_t)tile * sh.W * 6 > SORT_COARSE_SMEM) return 1;
  const size_t smem_coarse = fixed + (size_t)tile * sh.W * 6 + 16;
  const size_t smem_fine = ((size_t)(1u << fb) + SORT_WIN_CAP) * sizeof(u32);
  const size_t nseg = ncols * sh.nsets;
  const size_t total_bins = nseg * nbins;
  const size_t tiles = (n + tile - 1) / tile;
  if (tiles * ncols > 0x7fffffffull || total_bins > 0x7fffffffull) return 1;
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
    ctx->sort_attr_set = true;
  }
  void *p_reg, *p_pay, *p_key;
  // per bin: count | start of its run in the temporary array | cursor (ends as the end of that run)
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_REGION, 3 * total_bins * sizeof(u32), &p_reg));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_PAY, nseg * sh.seg_cap * sizeof(u32), &p_pay));
  EON_TRY(scratch_get(ctx, SC_MSM_SORT_KEY, nseg * sh.seg_cap * sizeof(unsigned short), &p_key));
  u32* bin_count = (u32*)p_reg;
  u32* tmp_start = bin_count + total_bins;
  u32* region_cursor = tmp_start + total_bins;
  cudaStream_t st = ctx->stream;
  const unsigned grid_tiles = (unsigned)(tiles * ncols);

  // window size known at compile time for the shapes the cost model picks (c = 17 up to 2^22 points)
  const bool c17 = sh.c == 17 && sh.W == 15;
  phase_begin(ctx, PH_MSM_DIGITS);
  EON_CUDA(ctx, cudaMemsetAsync(bin_count, 0, total_bins * sizeof(u32), st));
#define EON_SORT_LAUNCH(K, SMEM, ...)                                                                          \
  do {                                                                                                         \
    if (sh.merged && c17) K<true, 17><<<grid_tiles, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                    \
    else if (sh.merged) K<true, 0><<<grid_tiles, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                       \
    else if (c17) K<false, 17><<<grid_tiles, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                           \
    else K<false, 0><<<grid_tiles, SORT_THREADS, SMEM, st>>>(__VA_ARGS__);                                     \
  } while (0)
  EON_SORT_LAUNCH(k_bin_hist, tile_bins * sizeof(u32), d_scalars, n, ld, (u32)ncols, sh, nbins, fb, tile, bin_count);
  EON_LAUNCHED(ctx);
  k_bin_scan<<<(unsigned)nseg, 1024, 0, st>>>(bin_count, nbins, tmp_start, region_cursor);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_DIGITS);

  phase_begin(ctx, PH_MSM_SCATTER);
  if (sh.rounds)  // unused slots (bucket padding, segment tails) must read as ENTRY_NONE
    EON_CUDA(ctx, cudaMemsetAsync(d_entries, 0xff, nseg * sh.seg_cap * sizeof(u32), st));
  EON_SORT_LAUNCH(k_sort_coarse, smem_coarse, d_scalars, n, ld, (u32)ncols, sh, nbins, fb, tile, region_cursor,
                  (u32*)p_pay, (unsigned short*)p_key);
#undef EON_SORT_LAUNCH
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCATTER);

  phase_begin(ctx, PH_MSM_SCAN);
  k_sort_count<<<(unsigned)total_bins, SORT_THREADS, 0, st>>>((const unsigned short*)p_key, tmp_start, region_cursor,
                                                              sh.NB, nbins, fb, sh.seg_cap, d_hist);
  EON_LAUNCHED(ctx);
  EON_TRY(msm_scan_run(ctx, d_hist, d_cur, sh.NB, 1u << sh.rounds, d_seg_total, nseg));
  phase_end(ctx, PH_MSM_SCAN);

  phase_begin(ctx, PH_MSM_SCATTER);
  static const int order_env = getenv("EON_SORT_ORDERED") ? atoi(getenv("EON_SORT_ORDERED")) : 1;
  if (plan.on && plan.nbins > 1 && plan.nbins <= SLICE_ORDER_MAX && order_env)
    k_sort_fine<true><<<(unsigned)total_bins, SORT_FINE_THREADS, smem_fine + ((size_t)plan.nbins << fb) * sizeof(u32), st>>>(
        (const u32*)p_pay, (const unsigned short*)p_key, tmp_start, region_cursor, d_hist, d_seg_total, sh.NB, nbins, fb,
        sh.seg_cap, plan.shift, plan.nbins, d_cur, d_entries);
  else
    k_sort_fine<false><<<(unsigned)total_bins, SORT_FINE_THREADS, smem_fine, st>>>(
        (const u32*)p_pay, (const unsigned short*)p_key, tmp_start, region_cursor, d_hist, d_seg_total, sh.NB, nbins, fb,
        sh.seg_cap, 0u, 1u, d_cur, d_entries);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_MSM_SCATTER);
  return EON_OK;
}

}  // namespace eon
