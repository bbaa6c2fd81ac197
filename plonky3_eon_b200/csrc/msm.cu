// Batched G1 multi-scalar multiplication (signed-window Pippenger) on sm_100a.
//
// Replaces G1::multi_exp (bn254/src/curve.rs:158-180), i.e. the per-call `to_affine` of every
// SRS point (curve.rs:170) followed by halo2curves::msm::msm_best (curve.rs:177), and the
// per-column loops around it (kzg/src/pcs.rs:244-249,311-318).  All `ncols` columns of a
// coefficient matrix share the resident affine bases and are processed by the same launches.
//
// Pipeline (segment = one (column, window) pair, NB = 2^(c-1) buckets per segment):
//   1 digits     scalar: Montgomery -> canonical (x * 1 * R^-1), signed c-bit digits as int16,
//                bucket histogram with global REDs
//   2 scan       exclusive scan of each segment's histogram -> bucket start offsets
//   3 scatter    counting sort: entries[seg][start[b]++] = point index | sign << 31
//   4 accumulate one thread per bucket: XYZZ += +-base (mixed add, 8M+2S); buckets larger than
//                a chunk are split into extra tasks so skewed scalars stay load-balanced
//   5 reduce     per segment sum_b (b+1) * B_b by chunked running sums, tree-combined per block
//   6 combine    per column Horner over windows (c doublings + 1 add each), to affine
// The result is the canonical affine point, hence bit-identical to the reference whatever the
// order of additions.
#include <stdlib.h>

#include <algorithm>

#include "msm.cuh"

namespace eon {

// Windows needed for a canonical scalar (< r < 2^254) in signed c-bit digits: the top window must
// absorb the last carry without wrapping, i.e. hold at most c - 1 scalar bits: W*c >= 255.
static u32 msm_windows(u32 c) { return (255 + c - 1) / c; }

static u32 ceil_log2(size_t n) {
  u32 lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  return lg;
}

static MsmShape msm_shape_plain(size_t n) {
  int c = (int)ceil_log2(n) - 3;
  if (c < 2) c = 2;
  if (c > 16) c = 16;
  MsmShape s;
  memset(&s, 0, sizeof(s));
  s.c = (u32)c;
  s.W = msm_windows(s.c);
  s.NB = 1u << (s.c - 1);
  s.nsets = s.W;
  s.merged = 0;
  s.chunk = s.NB < 32 ? s.NB : 32;
  s.nchunks = s.NB / s.chunk;
  s.seg_cap = n;
  return s;
}

static MsmShape msm_shape_merged(size_t n, u32 c, u64 tab_stride, u64 first) {
  static int chunk_env = 0;
  if (!chunk_env) {
    const char* e = getenv("EON_MSM_CHUNK");
    chunk_env = e ? atoi(e) : 32;
    if (chunk_env != 8 && chunk_env != 16 && chunk_env != 64 && chunk_env != 128) chunk_env = 32;
  }
  MsmShape s;
  memset(&s, 0, sizeof(s));
  s.c = c;
  s.W = msm_windows(c);
  s.NB = 1u << (c - 1);
  s.nsets = 1;
  s.merged = 1;
  s.mont_digits = 1;  // k_srs_tables pre-scales every base by R^-1
  s.chunk = (u32)chunk_env;
  s.nchunks = s.NB / s.chunk;
  s.tab_stride = tab_stride;
  s.base_first = first;
  s.seg_cap = (u64)n * s.W;
  return s;
}

// bucket histogram.  1-D grid of ncols * ceil(n / threads) blocks, column index fastest: blocks
// that run at the same time then update the bucket sets of ALL columns, which spreads the L2
// atomics over ncols x more cache lines (they serialise per line: measured 3.5x on the hist pass).
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_hist(const Fr* __restrict__ scalars, size_t n, size_t ld, u32 ncols, MsmShape sh, u32* __restrict__ hist) {
  u32 col = blockIdx.x % ncols;
  size_t i = (size_t)(blockIdx.x / ncols) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = load_scalar(scalars, i, ld, col);
  for_each_digit(s, sh, [&](u32 w, int d) {
    size_t seg = (size_t)col * sh.nsets + (sh.merged ? 0 : w);
    u32 b = (u32)(d < 0 ? -d : d) - 1;
    atomicAdd(&hist[seg * sh.NB + b], 1u);
  });
}

// ---- 2. exclusive scan per segment ----------------------------------------------------------
// one block per segment; hist -> starts (in place), cursor = copy of starts.  Every bucket's slot
// range is rounded up to a multiple of `align` (a power of two; 2^rounds for the batched-affine
// pairwise rounds of msm_tree.cu, else 1), so starts are multiples of `align`.
__global__ void __launch_bounds__(1024)
k_msm_scan(u32* __restrict__ hist, u32* __restrict__ cursor, u32 NB, u32 align, u32* __restrict__ seg_total) {
  __shared__ u32 warp_sums[32];
  __shared__ u32 s_carry;
  u32* h = hist + (size_t)blockIdx.x * NB;
  u32* cur = cursor + (size_t)blockIdx.x * NB;
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  // four buckets per thread and sweep (16-byte accesses; NB is a power of two >= 4 on this path, and smaller bucket
  // sets take the scalar tail below): a quarter of the block-wide scans of the one-bucket-per-thread form, which ran
  // 64 dependent sweeps for the 2^16 buckets of a commit column with the whole GPU waiting (62 us -> ~20)
  const bool vec = (NB & 3u) == 0;
  const u32 per = vec ? 4u : 1u;
  for (u32 base = 0; base < NB; base += 1024 * per) {
    const u32 idx = base + tid * per;
    u32 v[4] = {0, 0, 0, 0};
    if (vec) {
      if (idx < NB) {
        const uint4 q = *reinterpret_cast<const uint4*>(h + idx);
        v[0] = (q.x + align - 1) & ~(align - 1);
        v[1] = (q.y + align - 1) & ~(align - 1);
        v[2] = (q.z + align - 1) & ~(align - 1);
        v[3] = (q.w + align - 1) & ~(align - 1);
      }
    } else if (idx < NB) {
      v[0] = (h[idx] + align - 1) & ~(align - 1);
    }
    const u32 mine = v[0] + v[1] + v[2] + v[3];
    u32 x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (u32)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      u32 ws = warp_sums[lane];
      u32 z = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, z, o);
        if (lane >= (u32)o) z += y;
      }
      warp_sums[lane] = z - ws;  // exclusive
    }
    __syncthreads();
    const u32 carry = s_carry;
    const u32 excl = carry + warp_sums[wid] + x - mine;
    if (idx < NB) {
      if (vec) {
        const uint4 o4 = make_uint4(excl, excl + v[0], excl + v[0] + v[1], excl + v[0] + v[1] + v[2]);
        *reinterpret_cast<uint4*>(h + idx) = o4;
        *reinterpret_cast<uint4*>(cur + idx) = o4;
      } else {
        h[idx] = excl;
        cur[idx] = excl;
      }
    }
    __syncthreads();
    if (tid == 1023) s_carry = excl + mine;
    __syncthreads();
  }
  if (tid == 0) seg_total[blockIdx.x] = s_carry;  // aligned slots in use by the segment
}

int msm_scan_run(eon_ctx* ctx, u32* d_hist, u32* d_cur, u32 NB, u32 align, u32* d_seg_total, size_t nseg) {
  k_msm_scan<<<(unsigned)nseg, 1024, 0, ctx->stream>>>(d_hist, d_cur, NB, align, d_seg_total);
  EON_LAUNCHED(ctx);
  return EON_OK;
}

// ---- 3. scatter (counting sort by bucket) ----------------------------------------------------
// The digits are recomputed from the scalar (one modmul) instead of being stored by pass 1: that is
// cheaper than writing and re-reading 2-4 bytes per (point, window).
// entry = index of the base to add | sign << 31; merged mode indexes the table level of window w.
// grid: as k_msm_hist
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_scatter(const Fr* __restrict__ scalars, size_t n, size_t ld, u32 ncols, MsmShape sh, u32* __restrict__ cursor,
              u32* __restrict__ entries) {
  u32 col = blockIdx.x % ncols;
  size_t i = (size_t)(blockIdx.x / ncols) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = load_scalar(scalars, i, ld, col);
  for_each_digit(s, sh, [&](u32 w, int d) {
    size_t seg = (size_t)col * sh.nsets + (sh.merged ? 0 : w);
    u32 b = (u32)(d < 0 ? -d : d) - 1;
    u32 pos = atomicAdd(&cursor[seg * sh.NB + b], 1u);
    u32 base = sh.merged ? (u32)(w * sh.tab_stride + sh.base_first + i) : (u32)i;
    entries[seg * sh.seg_cap + pos] = base | (d < 0 ? SIGN_BIT : 0u);
  });
}

// ---- 3b. bucket order: largest buckets first, equal sizes adjacent --------------------------------
// One thread accumulates one bucket, so a warp runs as long as its fullest bucket: with ~32 +- 6
// entries per bucket (uniform scalars) 28 % of the lanes idle (ncu: 22.97 active threads/inst).
// Sorting the bucket ids of every segment by entry count (counting sort, 256 bins, descending)
// gives every warp buckets of (nearly) one size.  One block per segment.
constexpr u32 ORDER_BINS = 256;
__global__ void __launch_bounds__(1024)
k_msm_order(const u32* __restrict__ starts, const u32* __restrict__ ends, u32 NB, u32 rshift, u32* __restrict__ order) {
  __shared__ u32 bin_count[ORDER_BINS];
  __shared__ u32 bin_pos[ORDER_BINS];
  const size_t seg = blockIdx.x;
  const u32* st = starts + seg * NB;
  const u32* en = ends + seg * NB;
  u32* ord = order + seg * NB;
  for (u32 i = threadIdx.x; i < ORDER_BINS; i += blockDim.x) bin_count[i] = 0;
  __syncthreads();
  const u32 rnd = (1u << rshift) - 1;  // after `rshift` pairwise rounds a bucket holds ceil(cnt / 2^rshift) points
  // (warp-aggregated increments through MATCH.ANY were tried: 56 -> 93 us per launch, profiles/r03s_ncu_launches_agg.txt)
  for (u32 b = threadIdx.x; b < NB; b += blockDim.x) {
    u32 cnt = (en[b] - st[b] + rnd) >> rshift;
    atomicAdd(&bin_count[min(cnt, ORDER_BINS - 1)], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 acc = 0;
    for (int k = ORDER_BINS - 1; k >= 0; k--) {  // descending size
      bin_pos[k] = acc;
      acc += bin_count[k];
    }
  }
  __syncthreads();
  for (u32 b = threadIdx.x; b < NB; b += blockDim.x) {
    u32 cnt = (en[b] - st[b] + rnd) >> rshift;
    u32 pos = atomicAdd(&bin_pos[min(cnt, ORDER_BINS - 1)], 1u);
    ord[pos] = b;
  }
}

// ---- 4. bucket accumulation --------------------------------------------------------------------
struct MsmTask {
  u32 bucket;  // global bucket id = seg * NB + b
  u32 seg;
  u32 begin, end;  // entry range inside the segment
};

__device__ __forceinline__ G1Affine load_base(const G1Affine* __restrict__ bases, u64 idx) {
  // two 256-bit loads (LDG.E.ENL2.256): one request per coordinate of the random 64-byte gather
  const Fq* p = reinterpret_cast<const Fq*>(bases + idx);
  G1Affine r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.x.v[0]), "=r"(r.x.v[1]), "=r"(r.x.v[2]), "=r"(r.x.v[3]), "=r"(r.x.v[4]), "=r"(r.x.v[5]),
                 "=r"(r.x.v[6]), "=r"(r.x.v[7])
               : "l"(p));
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.y.v[0]), "=r"(r.y.v[1]), "=r"(r.y.v[2]), "=r"(r.y.v[3]), "=r"(r.y.v[4]), "=r"(r.y.v[5]),
                 "=r"(r.y.v[6]), "=r"(r.y.v[7])
               : "l"(p + 1));
  return r;
}

// Where the finisher reads its addends: the sorted entries (index | sign into `bases`), or — after
// the batched-affine pairwise rounds of msm_tree.cu — the flat array of partial sums `pts`, in which
// bucket [begin, end) of a segment occupies slots [(seg*seg_cap + begin) >> rshift,
// ceil((seg*seg_cap + end) / 2^rshift)).
struct AccSrc {
  const G1Affine* bases;
  const u32* entries;
  u64 seg_cap;
  const G1Affine* pts;  // partial sums after the pairwise rounds (msm_tree.cu)
  u32 rshift;
};

template <bool PTS>
__device__ __forceinline__ G1Xyzz accumulate_range(const AccSrc& s, u32 seg, u32 begin, u32 end) {
  G1Xyzz acc = G1Xyzz::identity();
  if (begin >= end) return acc;
  if (PTS) {
    const u64 base = (u64)seg * s.seg_cap;
    const u64 b = (base + begin) >> s.rshift, e = (base + end + ((1ull << s.rshift) - 1)) >> s.rshift;
    G1Affine p_next = load_base(s.pts, b);
    for (u64 i = b; i < e; i++) {
      G1Affine p = p_next;
      if (i + 1 < e) p_next = load_base(s.pts, i + 1);
      g1_add_mixed(acc, p);
    }
    return acc;
  }
  // software pipeline: the next base (a random 64-byte gather, HBM-resident with window tables) is
  // in flight while the current mixed addition (~11 k cycles per warp) runs
  const u32* ent = s.entries + (u64)seg * s.seg_cap;
  u32 v_next = __ldg(ent + begin);
  G1Affine p_next = load_base(s.bases, v_next & ~SIGN_BIT);
  for (u32 e = begin; e < end; e++) {
    u32 v = v_next;
    G1Affine p = p_next;
    if (e + 1 < end) {
      v_next = __ldg(ent + e + 1);
      p_next = load_base(s.bases, v_next & ~SIGN_BIT);
    }
    if (v & SIGN_BIT) p.y = fp_neg(p.y);
    g1_add_mixed(acc, p);
  }
  return acc;
}

__device__ __forceinline__ void store_xyzz(G1Xyzz* dst, const G1Xyzz& p) { *dst = p; }

// one thread per bucket; starts[] = bucket begin, cursor[] = bucket end (after the scatter).
// chunk_min is in entry slots (a PTS chunk of chunk_min slots holds chunk_min >> rshift points).
template <bool PTS>
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_accumulate(AccSrc src, const u32* __restrict__ starts, const u32* __restrict__ ends,
                 const u32* __restrict__ order, u32 NB, size_t total_buckets, u32 chunk_min,
                 G1Xyzz* __restrict__ buckets, MsmTask* __restrict__ tasks, u32* __restrict__ ntasks, u32 max_tasks) {
  size_t slot = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (slot >= total_buckets) return;
  u32 seg = (u32)(slot / NB);
  size_t g = (size_t)seg * NB + order[slot];  // bucket handled by this thread
  u32 begin = starts[g], end = ends[g];
  u32 cnt = end - begin;
  u32 my_end = end;
  if (cnt > chunk_min) {
    // split: at most 512 chunks per bucket, each at least chunk_min entries (a multiple of 2^rshift)
    u32 ch = (cnt + 511) / 512;
    if (ch < chunk_min) ch = chunk_min;
    ch = (ch + ((1u << src.rshift) - 1)) & ~((1u << src.rshift) - 1);
    u32 extra = (cnt + ch - 1) / ch - 1;
    u32 slot = atomicAdd(ntasks, extra);
    if (slot + extra <= max_tasks) {
      for (u32 t = 0; t < extra; t++) {
        MsmTask tk;
        tk.bucket = (u32)g;
        tk.seg = seg;
        tk.begin = begin + (t + 1) * ch;
        tk.end = min(end, begin + (t + 2) * ch);
        tasks[slot + t] = tk;
      }
      my_end = begin + ch;
    }
    // else: task buffer exhausted (cannot happen with the sizing in msm_run); fall back to serial
  }
  G1Xyzz acc = accumulate_range<PTS>(src, seg, begin, my_end);
  store_xyzz(buckets + g, acc);
}

// extra chunks of oversized buckets -> partial sums
template <bool PTS>
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_accumulate_tasks(AccSrc src, const MsmTask* __restrict__ tasks, const u32* __restrict__ ntasks, u32 max_tasks,
                       G1Xyzz* __restrict__ partial) {
  u32 nt = min(*ntasks, max_tasks);
  for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    MsmTask tk = tasks[t];
    G1Xyzz acc = accumulate_range<PTS>(src, tk.seg, tk.begin, tk.end);
    partial[t] = acc;
  }
}

// fold the partial sums of each split bucket back into the bucket (tasks of one bucket are contiguous)
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_fold_tasks(const MsmTask* __restrict__ tasks, const u32* __restrict__ ntasks, u32 max_tasks,
                 const G1Xyzz* __restrict__ partial, G1Xyzz* __restrict__ buckets) {
  u32 nt = min(*ntasks, max_tasks);
  for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    u32 b = tasks[t].bucket;
    if (t > 0 && tasks[t - 1].bucket == b) continue;  // not the first task of its bucket
    G1Xyzz acc = buckets[b];
    for (u32 u = t; u < nt && tasks[u].bucket == b; u++) g1_add(acc, partial[u]);
    buckets[b] = acc;
  }
}

// ---- 5. bucket reduction ----------------------------------------------------------------------
// Per segment: sum_b (b + 1) B_b over its NB = 2^(c-1) buckets.  The usual running sums are one long dependent
// chain per thread, and this phase runs when the GPU is otherwise empty (2.3 ms^-1 ... a lone warp takes ~1200 cycles
// per Fq product), so the chain length IS the time: 64 + 18 + 23 XYZZ additions + a Fermat inversion measured
// 1.16 ms per MSM whatever the column count.  Here the bucket index is split b = hi L + lo (L = 2^ceil(log NB / 2)):
//     sum_b (b+1) B_b  =  S  +  sum_lo lo C_lo  +  L sum_hi hi R_hi,     R_hi = sum_lo B[hi][lo]   (row sums)
//                                                                          C_lo = sum_hi B[hi][lo]   (column sums)
//                                                                          S    = sum_hi R_hi
//   k_bucket_rowcol     all row and column sums: P lanes per sum (serial strip + shuffle tree), same 2 additions
//                       per bucket as the running sums but in chains of NB^(1/2) / P + log P
//   k_bucket_weighted   one CTA per vector (R or C of a segment): sum_i i X_i = sum_(j>=1) suffix_j by a block-wide
//                       suffix scan and a block-wide sum, 21 additions deep
//   k_bucket_finish     S + W_C + L W_R per segment
// ~45 additions deep instead of ~130, and no scalar multiplications.
__device__ __forceinline__ G1Xyzz ld_xyzz(const G1Xyzz* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = __ldg(q + i);
  G1Xyzz r;
  u32* d = reinterpret_cast<u32*>(&r);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[4 * i] = w[i].x; d[4 * i + 1] = w[i].y; d[4 * i + 2] = w[i].z; d[4 * i + 3] = w[i].w;
  }
  return r;
}
__device__ __forceinline__ void st_xyzz(G1Xyzz* p, const G1Xyzz& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  const u32* d = reinterpret_cast<const u32*>(&v);
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = make_uint4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
}
__device__ __forceinline__ G1Xyzz shfl_down_xyzz(const G1Xyzz& v, u32 delta, int width = 32) {
  G1Xyzz r;
  const u32* s = reinterpret_cast<const u32*>(&v);
  u32* d = reinterpret_cast<u32*>(&r);
#pragma unroll
  for (int i = 0; i < 32; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta, width);
  return r;
}

// grid: ceil(nseg * (H + L) * P / 128) CTAs of 128 threads; group g (P lanes) sums vector g % (H + L) of segment
// g / (H + L): vectors [0, H) are rows, [H, H + L) columns.
template <int P>
__global__ void __launch_bounds__(128)
k_bucket_rowcol(const G1Xyzz* __restrict__ buckets, u32 logL, u32 logH, size_t nseg, G1Xyzz* __restrict__ rows,
                G1Xyzz* __restrict__ cols) {
  const u32 L = 1u << logL, H = 1u << logH;
  const size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / P;
  const u32 sub = threadIdx.x & (P - 1);
  const bool live = g < nseg * (size_t)(H + L);  // whole groups are live or not: shuffles stay converged per group
  const size_t seg = live ? g / (H + L) : 0;
  const u32 w = live ? (u32)(g % (H + L)) : 0;
  const G1Xyzz* B = buckets + (seg << (logL + logH));
  G1Xyzz acc = G1Xyzz::identity();
  if (live) {
    if (w < H) {
      for (u32 i = sub; i < L; i += P) g1_add(acc, ld_xyzz(B + ((size_t)w << logL) + i));
    } else {
      const u32 lo = w - H;
      for (u32 i = sub; i < H; i += P) g1_add(acc, ld_xyzz(B + ((size_t)i << logL) + lo));
    }
  }
#pragma unroll
  for (int d = P / 2; d >= 1; d >>= 1) {
    const G1Xyzz t = shfl_down_xyzz(acc, d, P);
    if (sub < (u32)d) g1_add(acc, t);
  }
  if (live && sub == 0) st_xyzz(w < H ? rows + seg * H + w : cols + seg * L + (w - H), acc);
}

// One CTA per vector: blockIdx.x = 2 seg + which (0: the H row sums -> W_R and S, 1: the L column sums -> W_C).
// out[3 seg + 0] = W_R, [3 seg + 1] = W_C, [3 seg + 2] = S.  The vector is walked from its top in strips of
// blockDim elements (one strip up to c = 17, i.e. 256 buckets per side); `above` carries the sum of the strips
// already done, so that every thread ends up with the suffix sum of the whole vector from its element on.
constexpr int BW_THREADS = 256;  // 255 registers per thread: the XYZZ additions below run without spills
__global__ void __launch_bounds__(BW_THREADS)
k_bucket_weighted(const G1Xyzz* __restrict__ rows, const G1Xyzz* __restrict__ cols, u32 logL, u32 logH,
                  G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz s_tot[32];
  __shared__ G1Xyzz s_carry[32];
  __shared__ G1Xyzz s_above;
  const size_t seg = blockIdx.x >> 1;
  const u32 which = blockIdx.x & 1;
  const u32 n = 1u << (which ? logL : logH);
  const G1Xyzz* X = which ? cols + (seg << logL) : rows + (seg << logH);
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  G1Xyzz wsum = G1Xyzz::identity();  // thread 0: sum_i i X_i so far
  if (tid == 0) s_above = G1Xyzz::identity();
  const u32 strips = (n + blockDim.x - 1) / blockDim.x;
  for (u32 sidx = strips; sidx-- > 0;) {
    const u32 e = sidx * blockDim.x + tid;  // this thread's element
    G1Xyzz x = e < n ? ld_xyzz(X + e) : G1Xyzz::identity();
    // inclusive suffix sums inside the warp: x = sum of the warp's elements at lanes >= lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const G1Xyzz t = shfl_down_xyzz(x, d);
      if (lane + d < 32) g1_add(x, t);
    }
    __syncthreads();  // s_tot / s_carry of the previous strip are no longer read; s_above is written
    if (lane == 0) s_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      G1Xyzz y = lane < nw ? s_tot[lane] : G1Xyzz::identity();
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const G1Xyzz t = shfl_down_xyzz(y, d);
        if (lane + d < 32) g1_add(y, t);
      }
      // carry into warp a = the totals of the warps after it + everything above this strip
      G1Xyzz nxt = shfl_down_xyzz(y, 1);
      if (lane + 1 >= 32) nxt = G1Xyzz::identity();
      g1_add(nxt, s_above);
      s_carry[lane] = nxt;
    }
    __syncthreads();
    g1_add(x, s_carry[wid]);  // x = suffix sum over the whole vector from element e
    // sum_i i X_i = sum of the suffix sums from element 1 on
    G1Xyzz c = (e >= 1 && e < n) ? x : G1Xyzz::identity();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const G1Xyzz t = shfl_down_xyzz(c, d);
      if (lane < (u32)d) g1_add(c, t);
    }
    __syncthreads();  // every warp has read s_carry
    if (lane == 0) s_tot[wid] = c;
    if (tid == 0) s_above = x;  // suffix from the first element of this strip = everything from here up
    __syncthreads();
    if (wid == 0) {
      G1Xyzz y = lane < nw ? s_tot[lane] : G1Xyzz::identity();
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const G1Xyzz t = shfl_down_xyzz(y, d);
        if (lane < (u32)d) g1_add(y, t);
      }
      if (lane == 0) g1_add(wsum, y);
    }
  }
  __syncthreads();
  if (tid == 0) {
    st_xyzz(out + 3 * seg + which, wsum);
    if (which == 0) st_xyzz(out + 3 * seg + 2, s_above);  // S = the suffix sum from element 0
  }
}

// The same weighted sum by the BITS of the index: sum_i i X_i = sum_b 2^b S_b with S_b = the sum of the elements
// whose index has bit b set.  One warp per bit (n / 2 elements: a strip per lane + a shuffle tree), one more warp for
// the plain total S; then b doublings of S_b (all bits in parallel) and a tree over the bits: ~16 dependent
// additions for n = 256 against ~23 of the suffix-scan form above -- this launch runs alone on the GPU, so the depth
// of the chain is its duration.  blockDim = 32 * (max(logL, logH) + 1).  EON_BUCKET_W2=0: the suffix-scan form.
constexpr int BW2_MAX_THREADS = 384;  // up to 12 warps (n <= 2^11); 170 registers per thread
__global__ void __launch_bounds__(BW2_MAX_THREADS)
k_bucket_weighted2(const G1Xyzz* __restrict__ rows, const G1Xyzz* __restrict__ cols, u32 logL, u32 logH,
                   G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz s_grp[16];
  const size_t seg = blockIdx.x >> 1;
  const u32 which = blockIdx.x & 1;
  const u32 logn = which ? logL : logH;
  const u32 n = 1u << logn;
  const G1Xyzz* X = which ? cols + (seg << logL) : rows + (seg << logH);
  const u32 tid = threadIdx.x, lane = tid & 31, g = tid >> 5;  // warp g: bit g, or (g == logn) the total
  if (g <= logn) {
    G1Xyzz acc = G1Xyzz::identity();
    const u32 count = (g == logn) ? n : (n >> 1);
    for (u32 k = lane; k < count; k += 32) {
      // k-th index with bit g set: bit g inserted into k
      const u32 i = (g == logn) ? k : ((((k >> g) << 1) | 1u) << g) | (k & ((1u << g) - 1));
      g1_add(acc, ld_xyzz(X + i));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const G1Xyzz t = shfl_down_xyzz(acc, d);
      if (lane < (u32)d) g1_add(acc, t);
    }
    if (lane == 0) {
      if (g < logn)
        for (u32 j = 0; j < g; j++) acc = g1_dbl(acc);  // 2^g S_g
      s_grp[g] = acc;
    }
  }
  __syncthreads();
  if (g == 0) {
    G1Xyzz v = lane < logn ? s_grp[lane] : G1Xyzz::identity();
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {  // logn <= 15
      const G1Xyzz t = shfl_down_xyzz(v, d);
      if (lane < (u32)d) g1_add(v, t);
    }
    if (lane == 0) {
      st_xyzz(out + 3 * seg + which, v);
      if (which == 0) st_xyzz(out + 3 * seg + 2, s_grp[logn]);  // S = the plain sum of the row sums
    }
  }
}

// segsum[seg] = S + W_C + 2^logL W_R
__global__ void __launch_bounds__(32) k_bucket_finish(const G1Xyzz* __restrict__ parts, u32 logL, size_t nseg,
                                                      G1Xyzz* __restrict__ segsum) {
  const size_t seg = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (seg >= nseg) return;
  G1Xyzz t = ld_xyzz(parts + 3 * seg);
  for (u32 i = 0; i < logL; i++) t = g1_dbl(t);
  g1_add(t, ld_xyzz(parts + 3 * seg + 1));
  g1_add(t, ld_xyzz(parts + 3 * seg + 2));
  st_xyzz(segsum + seg, t);
}

// ---- 6. window combination + to-affine ----------------------------------------------------------
__global__ void k_msm_combine(const G1Xyzz* __restrict__ segsum, MsmShape sh, size_t ncols, G1Affine* __restrict__ out) {
  size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const G1Xyzz* S = segsum + col * sh.nsets;
  G1Xyzz acc = S[sh.nsets - 1];
  for (int w = (int)sh.nsets - 2; w >= 0; w--) {
    for (u32 i = 0; i < sh.c; i++) acc = g1_dbl(acc);
    g1_add(acc, S[w]);
  }
  out[col] = g1_to_affine(acc);
}

__global__ void k_fill_identity(G1Affine* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = G1Affine::identity();
}

// sum of n affine points -> 1 affine point (n is small: per-GPU partial sums)
__global__ void k_g1_sum(const G1Affine* __restrict__ pts, size_t n, G1Affine* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    G1Xyzz acc = G1Xyzz::identity();
    for (size_t i = 0; i < n; i++) g1_add_mixed(acc, pts[i]);
    *out = g1_to_affine(acc);
  }
}

// out[c] = sum_p pts[p * ncols + c]: the per-GPU partial sums of an index-range sharded MSM, all columns in one launch
__global__ void k_g1_sum_cols(const G1Affine* __restrict__ pts, size_t nparts, size_t ncols, G1Affine* __restrict__ out) {
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  G1Xyzz acc = G1Xyzz::identity();
  for (size_t p = 0; p < nparts; p++) g1_add_mixed(acc, pts[p * ncols + c]);
  out[c] = g1_to_affine(acc);
}

int g1_sum_cols_run(eon_ctx* ctx, const G1Affine* d_parts, size_t nparts, size_t ncols, G1Affine* d_out) {
  if (ncols == 0) return EON_OK;
  k_g1_sum_cols<<<(unsigned)((ncols + 31) / 32), 32, 0, ctx->stream>>>(d_parts, nparts, ncols, d_out);
  EON_LAUNCHED(ctx);
  return EON_OK;
}

int g1_sum_run(eon_ctx* ctx, const G1Affine* d_points, size_t n, G1Affine* d_out) {
  k_g1_sum<<<1, 32, 0, ctx->stream>>>(d_points, n, d_out);
  EON_LAUNCHED(ctx);
  return EON_OK;
}

// ---- orchestration -----------------------------------------------------------------------------
// Batched-affine pairwise rounds (msm_tree.cu) before the serial finisher: worth it once buckets hold
// many entries (each round halves them at ~7.6 instead of 10 products per addition).
// eon_msm_set_rounds / EON_MSM_ROUNDS override (0 = classic XYZZ accumulation only).
static u32 msm_pick_rounds(const eon_ctx* ctx, size_t n, const MsmShape& sh) {
  static int env_forced = -2;
  if (env_forced == -2) {
    const char* e = getenv("EON_MSM_ROUNDS");
    env_forced = e ? atoi(e) : -1;
  }
  int forced = ctx->msm_rounds >= 0 ? ctx->msm_rounds : env_forced;
  if (forced > 6) forced = 6;
  if (forced >= 0) return (u32)forced;
  const double per_bucket = (double)(sh.merged ? n * sh.W : n) / sh.NB;
  if (per_bucket >= 64) return 3;
  if (per_bucket >= 32) return 2;
  return 0;
}

static void msm_set_rounds(MsmShape& sh, size_t n, u32 rounds) {
  sh.rounds = rounds;
  const u64 raw = sh.merged ? (u64)n * sh.W : (u64)n;
  const u64 al = 1ull << rounds;
  // every bucket may waste up to 2^rounds - 1 padding slots
  sh.seg_cap = rounds ? ((raw + (u64)sh.NB * (al - 1) + al - 1) & ~(al - 1)) : raw;
}

// One batch of an MSM in three steps, so that the sort of a column group can be queued as soon as that group's
// scalars exist (the host-buffer commit sorts group g while group g + 1 is still crossing PCIe, msm_stream_*):
// msm_batch_setup (workspace, slice plan), msm_batch_sort (columns [c0, c0 + nc): digits, histogram, scan, sort --
// every column is its own segment, so the groups write disjoint parts of the same arrays), msm_batch_finish
// (pairwise rounds, finisher, bucket reduction over ALL columns of the batch at once: the slice schedule of round 0
// needs every column in the same launch).
int msm_batch_setup(eon_ctx* ctx, const G1Affine* d_bases, size_t n, size_t ncols, const MsmShape& sh, MsmBatch* B) {
  B->bases = d_bases;
  B->n = n;
  B->ncols = ncols;
  B->sh = sh;
  const size_t nseg = ncols * sh.nsets;
  const size_t total_buckets = nseg * sh.NB;
  B->chunk_min = 256u << sh.rounds;  // entry slots per split-off task: 256 addends in the finisher
  size_t max_tasks_sz = (nseg * sh.seg_cap) / B->chunk_min + 1024;
  if (max_tasks_sz > 0x7fffffffull) max_tasks_sz = 0x7fffffffull;
  B->max_tasks = (u32)max_tasks_sz;
  if (sh.seg_cap >= 0xffffffffull) return fail(ctx, EON_ERR_BAD_ARG, "msm: segment too large");
  const size_t grid_pts_sz = ((n + MSM_THREADS - 1) / MSM_THREADS) * ncols;
  if (grid_pts_sz > 0x7fffffffull) return fail(ctx, EON_ERR_BAD_ARG, "msm: grid too large");

  EON_TRY(scratch_get(ctx, SC_MSM_ORDER, total_buckets * sizeof(u32), &B->p_ord));
  EON_TRY(scratch_get(ctx, SC_MSM_HIST, total_buckets * sizeof(u32), &B->p_hist));
  EON_TRY(scratch_get(ctx, SC_MSM_CURSOR, total_buckets * sizeof(u32), &B->p_cur));
  EON_TRY(scratch_get(ctx, SC_MSM_ENTRIES, nseg * sh.seg_cap * sizeof(u32), &B->p_ent));
  EON_TRY(scratch_get(ctx, SC_MSM_BUCKETS, total_buckets * sizeof(G1Xyzz), &B->p_bkt));
  EON_TRY(scratch_get(ctx, SC_MSM_TASKS, (size_t)B->max_tasks * sizeof(MsmTask), &B->p_tasks));
  EON_TRY(scratch_get(ctx, SC_MSM_TASKPART, (size_t)B->max_tasks * sizeof(G1Xyzz), &B->p_tpart));
  // row sums, column sums (<= 2^ceil((c-1)/2) each) and 3 partial results per segment of the bucket reduction
  EON_TRY(scratch_get(ctx, SC_MSM_PARTIALS, nseg * ((size_t)2 << ((sh.c - 1 + 1) / 2)) * sizeof(G1Xyzz) + nseg * 3 * sizeof(G1Xyzz),
                      &B->p_part));
  EON_TRY(scratch_get(ctx, SC_MSM_SEGSUM, nseg * sizeof(G1Xyzz), &B->p_seg));
  EON_TRY(scratch_get(ctx, SC_MSM_MISC, 256, &B->p_misc));
  EON_TRY(scratch_get(ctx, SC_MSM_SEGTOTAL, nseg * sizeof(u32), &B->p_segtot));
  B->plan = msm_slice_plan(ctx, sh.merged ? (u64)sh.W * sh.tab_stride : (u64)n, (u64)nseg * sh.seg_cap, sh.rounds);
  EON_CUDA(ctx, cudaMemsetAsync(B->p_misc, 0, sizeof(u32), ctx->stream));
  B->place_deferred = false;
  if (B->plan.on) EON_TRY(msm_sort_begin(ctx));
  return EON_OK;
}

int msm_batch_sort(eon_ctx* ctx, MsmBatch& S, const Fr* d_scalars, size_t ld, size_t c0, size_t nc) {
  const MsmShape& sh = S.sh;
  const size_t n = S.n;
  const size_t seg0 = c0 * sh.nsets, nseg = nc * sh.nsets;
  u32* d_hist = (u32*)S.p_hist + seg0 * sh.NB;
  u32* d_cur = (u32*)S.p_cur + seg0 * sh.NB;
  u32* d_ent = (u32*)S.p_ent + seg0 * sh.seg_cap;
  u32* d_segtot = (u32*)S.p_segtot + seg0;
  const size_t total_buckets = nseg * sh.NB;
  cudaStream_t st = ctx->stream;
  const unsigned grid_pts = (unsigned)(((n + MSM_THREADS - 1) / MSM_THREADS) * nc);
  static int env_mode = -2;
  if (env_mode == -2) {
    const char* e = getenv("EON_MSM_SORT");
    env_mode = e ? atoi(e) : -1;
  }
  const int mode = ctx->msm_sort_mode >= 0 ? ctx->msm_sort_mode : env_mode;
  int rc = 1;
  // coalesced multi-pass sort (msm_sort.cu), which also produces the bucket histogram; small inputs and
  // unsupported shapes: global histogram + one-pass atomic scatter
  if (mode != 0 && (mode >= 1 || n >= 4096)) {
    bool deferred = false;
    rc = msm_sort_entries(ctx, d_scalars, n, nc, ld, sh, S.plan, seg0, S.ncols * sh.nsets, (u32*)S.p_hist,
                          (u32*)S.p_segtot, (u32*)S.p_cur, (u32*)S.p_ent, &deferred);
    if (rc == EON_OK && deferred) S.place_deferred = true;
  }
  if (rc < 0) return rc;
  if (rc > 0) {
    phase_begin(ctx, PH_MSM_DIGITS);
    EON_CUDA(ctx, cudaMemsetAsync(d_hist, 0, total_buckets * sizeof(u32), st));
    k_msm_hist<<<grid_pts, MSM_THREADS, 0, st>>>(d_scalars, n, ld, (u32)nc, sh, d_hist);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_DIGITS);
    phase_begin(ctx, PH_MSM_SCAN);
    EON_TRY(msm_scan_run(ctx, d_hist, d_cur, sh.NB, 1u << sh.rounds, d_segtot, nseg));
    phase_end(ctx, PH_MSM_SCAN);
    phase_begin(ctx, PH_MSM_SCATTER);
    if (sh.rounds)  // unused slots (bucket padding, segment tails) must read as ENTRY_NONE
      EON_CUDA(ctx, cudaMemsetAsync(d_ent, 0xff, nseg * sh.seg_cap * sizeof(u32), st));
    k_msm_scatter<<<grid_pts, MSM_THREADS, 0, st>>>(d_scalars, n, ld, (u32)nc, sh, d_cur, d_ent);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_SCATTER);
  }
  return EON_OK;
}

int msm_batch_finish(eon_ctx* ctx, const MsmBatch& S, G1Affine* d_out) {
  const MsmShape& sh = S.sh;
  const size_t ncols = S.ncols;
  const size_t nseg = ncols * sh.nsets;
  const size_t total_buckets = nseg * sh.NB;
  const u32 chunk_min = S.chunk_min, max_tasks = S.max_tasks;
  void *p_hist = S.p_hist, *p_cur = S.p_cur, *p_ent = S.p_ent, *p_bkt = S.p_bkt, *p_tasks = S.p_tasks,
       *p_tpart = S.p_tpart, *p_part = S.p_part, *p_seg = S.p_seg, *p_ord = S.p_ord;
  u32* d_ntasks = (u32*)S.p_misc;
  const G1Affine* d_bases = S.bases;
  const SlicePlan& plan = S.plan;
  cudaStream_t st = ctx->stream;

  if (S.place_deferred) {  // fused sort: placement of all segments + the pair records of round 0
    uint2* rec_e = nullptr;
    u32* rec_dest = nullptr;
    EON_TRY(msm_tree_records(ctx, (u64)nseg * sh.seg_cap, &rec_e, &rec_dest));
    EON_TRY(msm_sort_place(ctx, sh, plan, nseg, (u32*)p_hist, (u32*)S.p_segtot, (u32*)p_cur, (u32*)p_ent, rec_e, rec_dest));
  }
  if (ctx->ev_stagger) {  // two-stream split: the second half-batch starts here (see msm_run)
    EON_CUDA(ctx, cudaEventRecord(ctx->ev_stagger, st));
    ctx->ev_stagger = nullptr;
  }
  phase_begin(ctx, PH_MSM_ACCUM);
  {
    AccSrc src;
    src.bases = d_bases;
    src.entries = (const u32*)p_ent;
    src.seg_cap = sh.seg_cap;
    src.pts = nullptr;
    src.rshift = sh.rounds;
    if (sh.rounds)
      EON_TRY(msm_tree_rounds(ctx, d_bases, plan, (const u32*)p_ent, (u64)nseg * sh.seg_cap, sh.rounds, &src.pts,
                              S.place_deferred));
    phase_begin(ctx, PH_MSM_FINISH);
    k_msm_order<<<(unsigned)nseg, 1024, 0, st>>>((const u32*)p_hist, (const u32*)p_cur, sh.NB, sh.rounds, (u32*)p_ord);
    EON_LAUNCHED(ctx);
    unsigned blocks = (unsigned)((total_buckets + MSM_THREADS - 1) / MSM_THREADS);
    unsigned tb = (unsigned)ctx->num_sms * 8;
    if (sh.rounds) {
      k_msm_accumulate<true><<<blocks, MSM_THREADS, 0, st>>>(src, (const u32*)p_hist, (const u32*)p_cur,
                                                             (const u32*)p_ord, sh.NB, total_buckets, chunk_min,
                                                             (G1Xyzz*)p_bkt, (MsmTask*)p_tasks, d_ntasks, max_tasks);
      EON_LAUNCHED(ctx);
      k_msm_accumulate_tasks<true><<<tb, MSM_THREADS, 0, st>>>(src, (const MsmTask*)p_tasks, d_ntasks, max_tasks,
                                                               (G1Xyzz*)p_tpart);
    } else {
      k_msm_accumulate<false><<<blocks, MSM_THREADS, 0, st>>>(src, (const u32*)p_hist, (const u32*)p_cur,
                                                              (const u32*)p_ord, sh.NB, total_buckets, chunk_min,
                                                              (G1Xyzz*)p_bkt, (MsmTask*)p_tasks, d_ntasks, max_tasks);
      EON_LAUNCHED(ctx);
      k_msm_accumulate_tasks<false><<<tb, MSM_THREADS, 0, st>>>(src, (const MsmTask*)p_tasks, d_ntasks, max_tasks,
                                                                (G1Xyzz*)p_tpart);
    }
    EON_LAUNCHED(ctx);
    k_msm_fold_tasks<<<tb, MSM_THREADS, 0, st>>>((const MsmTask*)p_tasks, d_ntasks, max_tasks,
                                                 (const G1Xyzz*)p_tpart, (G1Xyzz*)p_bkt);
    EON_LAUNCHED(ctx);
    phase_end(ctx, PH_MSM_FINISH);
  }
  phase_end(ctx, PH_MSM_ACCUM);

  phase_begin(ctx, PH_MSM_REDUCE);
  {
    const u32 logNB = sh.c - 1, logL = (logNB + 1) / 2, logH = logNB - logL;
    const size_t nvec = nseg << logL;  // >= the row and the column count of a segment
    G1Xyzz* rows = (G1Xyzz*)p_part;            // [nseg][H]
    G1Xyzz* cols = rows + nvec;                // [nseg][L]
    G1Xyzz* parts = cols + nvec;               // [nseg][3]
    const size_t groups = nseg * (((size_t)1 << logH) + ((size_t)1 << logL));
    // lanes per row / column sum: the whole warp while there are few sums (short chains matter), narrower groups
    // once there is enough work to fill the GPU without paying for idle tree lanes
    const int P = groups * 32 <= (size_t)ctx->num_sms * 1024 ? 32 : (groups * 16 <= (size_t)ctx->num_sms * 2048 ? 16 : 8);
    const unsigned grid = (unsigned)((groups * P + 127) / 128);
    if (P == 32) k_bucket_rowcol<32><<<grid, 128, 0, st>>>((const G1Xyzz*)p_bkt, logL, logH, nseg, rows, cols);
    else if (P == 16) k_bucket_rowcol<16><<<grid, 128, 0, st>>>((const G1Xyzz*)p_bkt, logL, logH, nseg, rows, cols);
    else k_bucket_rowcol<8><<<grid, 128, 0, st>>>((const G1Xyzz*)p_bkt, logL, logH, nseg, rows, cols);
    EON_LAUNCHED(ctx);
    const unsigned bt = std::min<unsigned>(BW_THREADS, std::max(32u, 1u << logL));
    cudaStream_t ts = tiny_begin(ctx);  // three single-CTA-per-segment launches in a dependent chain
    static const int w2_env = getenv("EON_BUCKET_W2") ? atoi(getenv("EON_BUCKET_W2")) : 1;
    if (w2_env && 32u * (std::max(logL, logH) + 1) <= (u32)BW2_MAX_THREADS)
      k_bucket_weighted2<<<(unsigned)(2 * nseg), 32u * (std::max(logL, logH) + 1), 0, ts>>>(rows, cols, logL, logH, parts);
    else
      k_bucket_weighted<<<(unsigned)(2 * nseg), bt, 0, ts>>>(rows, cols, logL, logH, parts);
    EON_LAUNCHED(ctx);
    k_bucket_finish<<<(unsigned)((nseg + 31) / 32), 32, 0, ts>>>(parts, logL, nseg, (G1Xyzz*)p_seg);
    EON_LAUNCHED(ctx);
    k_msm_combine<<<(unsigned)((ncols + 31) / 32), 32, 0, ts>>>((const G1Xyzz*)p_seg, sh, ncols, d_out);
    EON_LAUNCHED(ctx);
    tiny_end(ctx);
  }
  phase_end(ctx, PH_MSM_REDUCE);
  return EON_OK;
}

static int msm_batch(eon_ctx* ctx, const G1Affine* d_bases, const Fr* d_scalars, size_t n, size_t ncols, size_t ld,
                     const MsmShape& sh, G1Affine* d_out) {
  MsmBatch B;
  EON_TRY(msm_batch_setup(ctx, d_bases, n, ncols, sh, &B));
  EON_TRY(msm_batch_sort(ctx, B, d_scalars, ld, 0, ncols));
  return msm_batch_finish(ctx, B, d_out);
}

// Shape of an MSM over d_bases (window tables when the bases lie inside the resident SRS), pairwise rounds, and the
// number of columns one batch may hold.
static void msm_select(eon_ctx* ctx, const G1Affine* d_bases, size_t n, MsmShape* out_sh, const G1Affine** out_bases,
                       size_t* out_batch) {
  MsmShape sh = msm_shape_plain(n);
  const G1Affine* bases = d_bases;
  // bases inside the resident SRS and window tables available: all windows share one bucket set
  bool have_tables = false;
  if (ctx->d_rng_tab && d_bases >= ctx->d_srs + ctx->rng_first && d_bases + n <= ctx->d_srs + ctx->rng_first + ctx->rng_n) {
    // tables built for exactly this index range (the shard of a multi-GPU MSM): their window size fits the shard
    MsmShape m = msm_shape_merged(n, ctx->rng_c, ctx->rng_n, (u64)(d_bases - (ctx->d_srs + ctx->rng_first)));
    if ((u64)n * m.W >= 4ull * m.NB) {
      sh = m;
      bases = ctx->d_rng_tab;
      have_tables = true;
    }
  }
  if (!have_tables && ctx->d_srs_tab && d_bases >= ctx->d_srs && d_bases + n <= ctx->d_srs + ctx->srs_n) {
    MsmShape m = msm_shape_merged(n, ctx->srs_tab_c, ctx->srs_n, (u64)(d_bases - ctx->d_srs));
    if ((u64)n * m.W >= 4ull * m.NB) {  // enough entries per bucket for the larger window to pay off
      sh = m;
      bases = ctx->d_srs_tab;
    }
  }
  msm_set_rounds(sh, n, msm_pick_rounds(ctx, n, sh));
  ctx->msm_rounds_used = sh.rounds;
  ctx->msm_c_used = sh.c;
  // column batches: bound the sort + pairwise-round workspace (entries 4 B, round outputs 32 + 16 B,
  // level products ~5 B, prefix products 16 B per entry slot)
  size_t per_slot = 4 + (sh.rounds ? 32 + 16 + 5 + 16 : 0);
  size_t per_col = sh.seg_cap * sh.nsets * per_slot + (size_t)sh.nsets * sh.NB * (sizeof(G1Xyzz) + 16);
  // 64 GiB of the 180 while the tables are small enough for the slice schedule (<= 32 slices: up to 2^20 points at
  // c = 17): the 32 quotient columns of an opening at two points stay ONE batch -- every base gathered by all of them
  // per launch, one set of single-warp phases: open 72.6 -> 70.6 ms.  Larger tables are gathered at random whatever
  // the batch, and there wider batches measured SLOWER (2^24 x 8 columns, c = 20: 460.9 ms at 3 columns per batch
  // against 432.7 ms at 1; profiles/r02p_msm_batch_budget.txt): 24 GiB.  EON_MSM_BUDGET_GB overrides.
  static const size_t budget_gb = getenv("EON_MSM_BUDGET_GB") ? (size_t)atoll(getenv("EON_MSM_BUDGET_GB")) : 0;
  const bool sliceable = sh.merged && (((u64)sh.W * sh.tab_stride) >> 19) + 1 <= SLICE_ORDER_MAX;
  size_t budget = (budget_gb ? budget_gb : (sliceable ? 64 : 24)) << 30;
  size_t batch = budget / per_col;
  if (batch < 1) batch = 1;
  if (batch > 64) batch = 64;
  *out_sh = sh;
  *out_bases = bases;
  *out_batch = batch;
}

// ---- an MSM whose sort is queued per column group (see msm_batch_setup) -----------------------------------------
// msm_stream_begin returns 1 (nothing queued) when the MSM would not run as ONE unsplit batch: the caller then uses
// msm_run after the last group.
int msm_stream_begin(eon_ctx* ctx, const G1Affine* d_bases, size_t n, size_t ncols, MsmBatch* B) {
  if (n == 0 || ncols == 0 || n >= 0x7fffffffull) return 1;
  MsmShape sh;
  const G1Affine* bases;
  size_t batch;
  msm_select(ctx, d_bases, n, &sh, &bases, &batch);
  static const int split_env = getenv("EON_MSM_SPLIT") ? atoi(getenv("EON_MSM_SPLIT")) : -1;
  const int split_mode = ctx->msm_split_mode >= 0 ? ctx->msm_split_mode : split_env;
  // several batches, or the two-stream split of few columns (msm_run)
  if (ncols > batch || split_mode == 1 || (split_mode != 0 && ncols <= 4)) return 1;
  return msm_batch_setup(ctx, bases, n, ncols, sh, B);
}

int msm_run(eon_ctx* ctx, const G1Affine* d_bases, const Fr* d_scalars, size_t n, size_t ncols, size_t ld,
            G1Affine* d_out) {
  if (ncols == 0) return EON_OK;
  if (n == 0) {  // empty MSM -> identity (curve.rs:165-167)
    k_fill_identity<<<(unsigned)((ncols + 255) / 256), 256, 0, ctx->stream>>>(d_out, ncols);
    EON_LAUNCHED(ctx);
    return EON_OK;
  }
  if (n >= 0x7fffffffull) return fail(ctx, EON_ERR_BAD_ARG, "msm: more than 2^31 - 1 points");
  MsmShape sh;
  const G1Affine* bases;
  size_t batch;
  msm_select(ctx, d_bases, n, &sh, &bases, &batch);
  // Few columns: two half-batches on two streams.  A batch of 2 columns of 2^20 points spends ~1.8 of its 6.8 ms in
  // phases that leave the GPU idle (three inversion trees, the bucket reduction: single-warp dependent chains);
  // with two independent halves in flight those phases of one half run under the wide kernels of the other.
  // (At 8+ columns it loses: every base is gathered by half as many columns per launch, which is what the slice
  // schedule of round 0 lives on -- measured 47.1 vs 45.1 ms at 16 columns.)
  static const int split_env = getenv("EON_MSM_SPLIT") ? atoi(getenv("EON_MSM_SPLIT")) : -1;
  const int split_mode = ctx->msm_split_mode >= 0 ? ctx->msm_split_mode : split_env;
  const bool split = ncols >= 2 && ncols <= batch && ctx->bank == 0 &&
                     (split_mode == 1 || (split_mode != 0 && ncols <= 4 && n >= ((size_t)1 << 16)));
  if (split) {
    if (!ctx->split_stream) {
      int least = 0, greatest = 0;
      EON_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
      // wide kernels of the second half one level below the top; the top level is for the single-warp phases
      const int wide = greatest < least ? greatest + 1 : greatest;
      EON_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->split_stream, cudaStreamNonBlocking, wide));
      for (auto& e : ctx->ev_split) EON_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      for (int b = 0; b < 2; b++) {
        EON_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->tiny_stream[b], cudaStreamNonBlocking, greatest));
        for (auto& e : ctx->ev_tiny[b]) EON_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
    }
    static const int tiny_env = getenv("EON_MSM_TINY_PRIO") ? atoi(getenv("EON_MSM_TINY_PRIO")) : 1;
    ctx->tiny_on = tiny_env != 0;
    const size_t h0 = (ncols + 1) / 2;
    cudaStream_t main_stream = ctx->stream;
    // The second half is ordered after everything queued so far (the scalars are produced on the main stream).
    // EON_MSM_STAGGER=1 starts it after the SORT of the first half instead, so that the halves do not run in
    // lockstep; measured (profiles/r02h_stagger.txt) within noise of starting together: 7.23 vs 7.15 ms at 2
    // columns, 12.32 vs 12.28 at 4 -- off by default.
    static const int stagger_env = getenv("EON_MSM_STAGGER") ? atoi(getenv("EON_MSM_STAGGER")) : 0;
    if (stagger_env) ctx->ev_stagger = ctx->ev_split[0];
    else EON_CUDA(ctx, cudaEventRecord(ctx->ev_split[0], main_stream));
    int rc = msm_batch(ctx, bases, d_scalars, n, h0, ld, sh, d_out);
    if (ctx->ev_stagger) {  // the first half failed before its sort was queued
      ctx->ev_stagger = nullptr;
      cudaEventRecord(ctx->ev_split[0], main_stream);
    }
    EON_CUDA(ctx, cudaStreamWaitEvent(ctx->split_stream, ctx->ev_split[0], 0));
    if (rc == EON_OK) {
      ctx->stream = ctx->split_stream;
      ctx->bank = 1;
      rc = msm_batch(ctx, bases, d_scalars + h0, n, ncols - h0, ld, sh, d_out + h0);
      ctx->bank = 0;
      ctx->stream = main_stream;
    }
    ctx->tiny_on = false;
    // join in any case: nothing may outlive the call on the second stream
    cudaError_t e = cudaEventRecord(ctx->ev_split[1], ctx->split_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, ctx->ev_split[1], 0);
    if (rc == EON_OK && e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("msm split join failed: ") + cudaGetErrorString(e));
    return rc;
  }
  for (size_t c0 = 0; c0 < ncols; c0 += batch) {
    size_t nc = std::min(batch, ncols - c0);
    EON_TRY(msm_batch(ctx, bases, d_scalars + c0, n, nc, ld, sh, d_out + c0));
  }
  return EON_OK;
}

// ---- window tables: tab[t][i] = 2^(c t) * P_i, affine -------------------------------------------
// Static bases (the SRS never changes between calls, kzg/src/params.rs:57-77) let every window of a
// scalar use its own pre-shifted copy of the point, so all W windows of a column fall into ONE set
// of 2^(c-1) buckets: the bucket reduction shrinks W-fold and c can grow (fewer windows).
// Every table point carries the factor R^-1 mod r (R = 2^256, r = the group order): an MSM through the tables then
// takes its digits from the Montgomery limbs s R mod r of a scalar as they arrive -- sum (s_i R) (R^-1 P_i) =
// sum s_i P_i -- instead of paying a from-Montgomery product per scalar in every pass of the sort (two per scalar and
// MSM: 0.5 ms of a 2^20 x 16 commit).  The price is one 254-bit scalar multiplication per SRS point when the tables
// are built (once per SRS).
struct RinvLimbs {
  u32 k[8];
};
__global__ void __launch_bounds__(128) k_srs_tables(const G1Affine* __restrict__ srs, size_t n, u32 c, u32 W,
                                                    RinvLimbs rinv, G1Affine* __restrict__ tab) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Xyzz acc = g1_mul_canonical(srs[i], rinv.k);
  tab[i] = g1_to_affine(acc);
  for (u32 t = 1; t < W; t++) {
    for (u32 k = 0; k < c; k++) acc = g1_dbl(acc);
    tab[(size_t)t * n + i] = g1_to_affine(acc);
  }
}

static RinvLimbs rinv_limbs() {
  Fr one_int = Fr::zero();
  one_int.v[0] = 1;            // the INTEGER 1 read as limbs: from-Montgomery gives 1 * R^-1 mod r
  RinvLimbs r;
  fp_from_mont(r.k, one_int);
  return r;
}

int srs_build_tables(eon_ctx* ctx, unsigned window_bits) {
  if (ctx->d_srs_tab) {
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EON_CUDA(ctx, cudaFree(ctx->d_srs_tab));
    ctx->d_srs_tab = nullptr;
    ctx->srs_tab_c = 0;
  }
  const size_t n = ctx->srs_n;
  if (n == 0 || window_bits == 0) return EON_OK;
  if (window_bits < 8 || window_bits > 20) return fail(ctx, EON_ERR_BAD_ARG, "window bits must be in [8, 20]");
  const u32 c = window_bits, W = msm_windows(c);
  if ((u64)W * n >= 0x7fffffffull) return fail(ctx, EON_ERR_BAD_ARG, "SRS too large for window tables");
  EON_CUDA(ctx, cudaMalloc(&ctx->d_srs_tab, (size_t)W * n * sizeof(G1Affine)));
  k_srs_tables<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_srs, n, c, W, rinv_limbs(), ctx->d_srs_tab);
  EON_LAUNCHED(ctx);
  ctx->srs_tab_c = c;
  return EON_OK;
}

// cost model shared by the table policies: n * W(c) mixed additions (10 modmul) + 2^(c-1) bucket-reduction steps
// MEASURED (B200, one column, profiles/r02d_msm_window_sweep.txt): 2^21 points: c = 17 6.74 ms, c = 20 7.31 ms,
// c = 16 7.41 ms; 2^24 points: c = 20 39.2 ms, c = 19 45.2 ms, c = 18 61.9 ms.  Every bucket costs far more than
// its two reduction additions (the finisher's serial chain, the reduction's dependent steps): 200 products per
// bucket reproduces the measured optimum 17 up to 2^22 points and 20 from 2^23 on.  c = 18, 19 are excluded: their
// shapes run the two-pass sort with one bucket set per tile and measured far off the model (finisher 3-5x slower).
static u32 msm_best_window(size_t n) {
  u32 best_c = 0;
  double best = 0;
  for (u32 c = 10; c <= 20; c++) {
    if (c == 18 || c == 19) continue;
    double cost = (double)n * msm_windows(c) * 10.0 + (double)(1u << (c - 1)) * 200.0;
    if (!best_c || cost < best) {
      best_c = c;
      best = cost;
    }
  }
  return best_c;
}

// Window tables for the SRS index range [first, first + n) only: the shard one GPU owns in an index-range sharded
// MSM.  window_bits 0 = chosen for n points by the cost model; n = 0 drops them.
int srs_build_range_tables(eon_ctx* ctx, size_t first, size_t n, unsigned window_bits) {
  if (ctx->d_rng_tab) {
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EON_CUDA(ctx, cudaFree(ctx->d_rng_tab));
    ctx->d_rng_tab = nullptr;
    ctx->rng_first = ctx->rng_n = 0;
    ctx->rng_c = 0;
  }
  if (n == 0) return EON_OK;
  if (first > ctx->srs_n || n > ctx->srs_n - first) return fail(ctx, EON_ERR_BAD_ARG, "SRS range out of bounds");
  if (window_bits == 0) window_bits = msm_best_window(n);
  if (window_bits < 8 || window_bits > 20) return fail(ctx, EON_ERR_BAD_ARG, "window bits must be in [8, 20]");
  const u32 c = window_bits, W = msm_windows(c);
  if ((u64)W * n >= 0x7fffffffull) return fail(ctx, EON_ERR_BAD_ARG, "range too large for window tables");
  EON_CUDA(ctx, cudaMalloc(&ctx->d_rng_tab, (size_t)W * n * sizeof(G1Affine)));
  k_srs_tables<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_srs + first, n, c, W, rinv_limbs(), ctx->d_rng_tab);
  EON_LAUNCHED(ctx);
  ctx->rng_first = first;
  ctx->rng_n = n;
  ctx->rng_c = c;
  return EON_OK;
}

// default table policy after an SRS load (n >= 2^14 points, tables within 64 GiB): the c in [10, 20]
// that minimises  n * W(c) mixed additions (10 modmul)  +  2^(c-1) bucket-reduction steps (~60 modmul
// with the chunk offsets); 2^20 points -> c = 17 (15 windows).  Otherwise plain per-window buckets.
int srs_build_default_tables(eon_ctx* ctx) {
  const size_t n = ctx->srs_n;
  if (n < ((size_t)1 << 14)) return srs_build_tables(ctx, 0);
  const u32 best_c = msm_best_window(n);
  u32 W = msm_windows(best_c);
  if ((size_t)W * n * sizeof(G1Affine) > ((size_t)64 << 30)) return srs_build_tables(ctx, 0);
  return srs_build_tables(ctx, best_c);
}

// ---- synthetic SRS: g1_powers[i] = alpha^i * G (init_srs_unsafe, kzg/src/params.rs:123-139) ----
__global__ void __launch_bounds__(128) k_srs_generate(G1Affine* out, size_t n, Fr alpha) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = fp_pow_u64(alpha, (u64)i);
  u32 k[8];
  fp_from_mont(k, s);
  G1Xyzz p = g1_mul_canonical(G1Affine::generator(), k);
  out[i] = g1_to_affine(p);
}

int srs_generate(eon_ctx* ctx, const Fr& alpha, size_t n) {
  if (ctx->d_srs) {
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EON_CUDA(ctx, cudaFree(ctx->d_srs));
    ctx->d_srs = nullptr;
    ctx->srs_n = 0;
    EON_TRY(srs_build_tables(ctx, 0));
    EON_TRY(srs_build_range_tables(ctx, 0, 0, 0));
  }
  if (n == 0) return EON_OK;
  EON_CUDA(ctx, cudaMalloc(&ctx->d_srs, n * sizeof(G1Affine)));
  k_srs_generate<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_srs, n, alpha);
  EON_LAUNCHED(ctx);
  ctx->srs_n = n;
  return srs_build_default_tables(ctx);
}

}  // namespace eon
