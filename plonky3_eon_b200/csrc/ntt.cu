// Batched coset NTT over BN254 Fr on row-major matrices (sm_100a).
//
// Replaces the reference's CPU transforms behind TwoAdicSubgroupDft<Fr>:
//   Radix2Dit::dft_batch            dft/src/radix_2_dit.rs:64-122   (log2(h) full-matrix passes)
//   trait defaults                  dft/src/traits.rs:83-91,111-122,144-153,226-249
//   divide_by_height / coset_shift  dft/src/util.rs:15-36
//   reverse_matrix_index_bits       matrix/src/util.rs:36-56
// with 2-3 HBM passes per transform: every pass stages a tile of 2^r rows x cv columns in
// shared memory, runs r butterfly layers there and writes the tile back.  Bit-reversal,
// zero-padding (as replication through the first added_bits layers), the coset shift (baked
// into per-layer twiddles) and the 1/h scale are all fused into those passes.
//
// Forward transform  = DIT network, layers ascending, input in bit-reversed position order:
//     (a, b) -> (a + t*b, a - t*b),  t = TW_l[j] = shift^(n/2^(l+1)) * omega_(2^(l+1))^j
// Inverse transform  = the exact inverse network, layers descending, output bit-reversed:
//     (u, v) -> (u + v, (u - v) * TW_l[j]^-1),  then one multiply by 1/n on the final store.
// Both are exact field arithmetic, so results equal the reference's canonical limbs.
#include <stdlib.h>

#include "common.cuh"
#include "fp_shoup.cuh"

// MEASURED ON B200 (profiles/r01n_bench_*.json, 2^20 x 16 commit + LDE): NTT passes 10.37 ms with Montgomery
// twiddles (3 CTAs/SM), 9.69 ms with fixed-operand twiddles at 3 CTAs/SM (80 registers, some spills), 9.05 ms at
// 2 CTAs/SM (128 registers, none) -> fixed-operand twiddles at 2 CTAs/SM are the default.
#ifndef EON_NTT_SHOUP_DEFAULT
#define EON_NTT_SHOUP_DEFAULT 1
#endif
#ifndef EON_NTT_DB_DEFAULT
#define EON_NTT_DB_DEFAULT 0
#endif

namespace eon {

// ---- 16-byte unit helpers ---------------------------------------------------------------
__device__ __forceinline__ Fr fr_from_units(const uint4& lo, const uint4& hi) {
  Fr r;
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
__device__ __forceinline__ void fr_to_units(const Fr& a, uint4& lo, uint4& hi) {
  lo = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  hi = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ Fr fr_ldg(const uint4* p) {
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  return fr_from_units(lo, hi);
}

// ---- lazily reduced butterflies (fp.cuh): Fr used as a plain 256-bit container ---------------------
// forward (DIT) values live in [0, 4r) between layers, inverse (DIF) values in [0, 2r); the last pass
// of a transform canonicalises (or multiplies by 1/n, which canonicalises too).
__device__ __forceinline__ Fr lz_red(const Fr& a) { Fr r; fp_reduce_2p<FrParams>(r.v, a.v); return r; }
__device__ __forceinline__ Fr lz_add(const Fr& a, const Fr& b) { Fr r; fp_add_raw(r.v, a.v, b.v); return r; }
__device__ __forceinline__ Fr lz_sub(const Fr& a, const Fr& b) { Fr r; fp_sub_plus_2p<FrParams>(r.v, a.v, b.v); return r; }
// A twiddle as the butterflies consume it.  TwM: Montgomery form, multiplied with the word-serial Montgomery
// product (272 IMAD).  TwS: plain form w plus wq = floor(w 2^256 / r), multiplied with the fixed-operand product
// of fp_shoup.cuh (214 IMAD; Montgomery-form data times a plain-form twiddle stays in Montgomery form).
struct TwM { Fr w; };
struct TwS { Fr w, wq; };
template <bool SH> struct TwOf { typedef TwM T; };
template <> struct TwOf<true> { typedef TwS T; };
// tw canonical (< r), x any 256-bit value -> [0, 2r)
__device__ __forceinline__ Fr lz_mul(const TwM& tw, const Fr& x) { Fr r; fp_mul_lazy<FrParams>(r.v, tw.w.v, x.v); return r; }
__device__ __forceinline__ Fr lz_mul(const TwS& tw, const Fr& x) {
  u32 t[8];
  shoup::mul_lazy<FrParams>(t, x.v, tw.w.v, tw.wq.v);  // [0, 3r)
  Fr r;
  fp_reduce_2p<FrParams>(r.v, t);                       // -> [0, 2r)
  return r;
}
// forward butterfly: a in [0, 4r), b in [0, 4r) -> both outputs in [0, 4r)
template <class TW>
__device__ __forceinline__ void bf_dit(Fr& a, Fr& b, const TW& tw) {
  Fr x = lz_red(a), t = lz_mul(tw, b);
  a = lz_add(x, t);
  b = lz_sub(x, t);
}
// inverse butterfly: u, v in [0, 2r) -> both outputs in [0, 2r)
template <class TW>
__device__ __forceinline__ void bf_dif(Fr& u, Fr& v, const TW& tw) {
  Fr s = lz_red(lz_add(u, v));
  v = lz_mul(tw, lz_sub(u, v));
  u = s;
}
// entry idx of a twiddle table: 32 bytes (TwM) or 64 bytes (TwS: w then wq)
template <bool SH>
__device__ __forceinline__ typename TwOf<SH>::T tw_ldg(const uint4* tab, u64 idx) {
  typename TwOf<SH>::T t;
  if constexpr (SH) {
    t.w = fr_ldg(tab + idx * 4);
    t.wq = fr_ldg(tab + idx * 4 + 2);
  } else {
    t.w = fr_ldg(tab + idx * 2);
  }
  return t;
}

// ---- twiddle tables ---------------------------------------------------------------------
// tw[(1<<l) - 1 + j] = base[l] * gen[l]^j   for l < log_n, j < 2^l
// SH: every entry is the pair (plain form, floor(plain * 2^256 / r)) for the fixed-operand product
template <bool SH>
__global__ void k_gen_twiddles(Fr* tw, u32 log_n, const Fr* __restrict__ base, const Fr* __restrict__ gen) {
  u64 total = (1ull << log_n) - 1;
  for (u64 idx = blockIdx.x * (u64)blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
    u32 l = 63 - __clzll((long long)(idx + 1));
    u64 j = idx + 1 - (1ull << l);
    Fr g = gen[l];
    Fr r = fp_mul(base[l], fp_pow_u64(g, j));
    if constexpr (SH) {
      Fr w, wq;
      shoup::precompute<FrParams>(w.v, wq.v, r);
      tw[2 * idx] = w;
      tw[2 * idx + 1] = wq;
    } else {
      tw[idx] = r;
    }
  }
}

// butterflies multiply by (plain, quotient) twiddle pairs with the fixed-operand product (EON_NTT_SHOUP=0: by
// Montgomery-form twiddles with the word-serial product, the form every other kernel uses)
static int g_ntt_shoup = -1;
static bool ntt_use_shoup() {
  if (g_ntt_shoup < 0) {
    const char* e = getenv("EON_NTT_SHOUP");
    g_ntt_shoup = e ? (atoi(e) != 0) : EON_NTT_SHOUP_DEFAULT;
  }
  return g_ntt_shoup != 0;
}

static int get_twiddles(eon_ctx* ctx, unsigned log_n, const Fr& shift, int inverse, bool sh, const Fr** out) {
  TwiddleKey key;
  key.log_n = log_n;
  key.inverse = inverse | (sh ? 2 : 0);  // the two table forms are cached side by side
  memcpy(key.shift, shift.v, 32);
  auto it = ctx->twiddles.find(key);
  if (it != ctx->twiddles.end()) {
    *out = it->second;
    return EON_OK;
  }
  if (log_n == 0) {
    *out = nullptr;
    return EON_OK;
  }
  phase_begin(ctx, PH_NTT_TWIDDLE);
  // host: per-layer base and generator
  std::vector<Fr> hb(2 * log_n);
  Fr s = inverse ? fp_inv(shift) : shift;
  std::vector<Fr> sq(log_n);
  sq[0] = s;
  for (unsigned i = 1; i < log_n; i++) sq[i] = fp_sqr(sq[i - 1]);
  for (unsigned l = 0; l < log_n; l++) {
    hb[l] = sq[log_n - 1 - l];
    Fr g = fr_two_adic_generator(l + 1);
    hb[log_n + l] = inverse ? fp_inv(g) : g;
  }
  Fr* d_tab = nullptr;
  size_t entries = ((size_t)1 << log_n) - 1;
  // the cache is keyed by (size, shift, direction); callers with ever-changing shifts must not grow it without
  // bound: past the budget every cached table is dropped (stream-ordered work on them finishes first)
  static size_t budget = 0;
  if (!budget) {
    const char* e = getenv("EON_TWIDDLE_CACHE_MB");
    budget = (e && atoll(e) > 0 ? (size_t)atoll(e) : (size_t)8192) << 20;
  }
  const size_t tab_bytes = (entries + 1) * sizeof(Fr) * (sh ? 2 : 1);
  if (ctx->twiddle_bytes + tab_bytes > budget && !ctx->twiddles.empty()) {
    // (a transform may be running on either compute stream of the context: wait for the whole device)
    EON_CUDA(ctx, cudaDeviceSynchronize());
    for (auto& kv : ctx->twiddles) cudaFree(kv.second);
    ctx->twiddles.clear();
    ctx->twiddle_bytes = 0;
  }
  EON_CUDA(ctx, cudaMalloc(&d_tab, tab_bytes));
  ctx->twiddle_bytes += tab_bytes;
  void* d_small = nullptr;
  EON_TRY(scratch_get(ctx, SC_SMALL, 2 * 32 * sizeof(Fr), &d_small));
  EON_CUDA(ctx, cudaMemcpyAsync(d_small, hb.data(), hb.size() * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  // the host vector must outlive the async copy from pageable memory: cudaMemcpyAsync from
  // pageable memory is staged before returning, so this is safe.
  unsigned blocks = (unsigned)std::min<size_t>((entries + 255) / 256, (size_t)ctx->num_sms * 16);
  if (blocks == 0) blocks = 1;
  if (sh) k_gen_twiddles<true><<<blocks, 256, 0, ctx->stream>>>(d_tab, log_n, (const Fr*)d_small, (const Fr*)d_small + log_n);
  else k_gen_twiddles<false><<<blocks, 256, 0, ctx->stream>>>(d_tab, log_n, (const Fr*)d_small, (const Fr*)d_small + log_n);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_NTT_TWIDDLE);
  ctx->twiddles[key] = d_tab;
  *out = d_tab;
  return EON_OK;
}

// 16-byte asynchronous global -> shared copy (LDGSTS): the tile goes from HBM to shared memory without a stop in
// registers, and all of a thread's units are in flight at once
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ Fr fr_lds(const uint4* p) { return fr_from_units(p[0], p[1]); }

// ---- one HBM pass: r butterfly layers on a shared-memory tile -----------------------------
struct PassParams {
  const uint4* src;
  uint4* dst;
  const uint4* tw;
  u64 ld_src, ld_dst;  // physical row pitch (elements) of src / dst: w for a dense matrix, more when the
                       // transform runs on a column group of a wider matrix
  u64 V;        // virtual width = 2^l0 * w elements between consecutive tile rows
  u64 tiles_v;  // ceil(V / cv)
  u32 log_n, w;
  int w_shift;  // log2(w) if w is a power of two, else -1
  u32 l0, r;
  u32 log_cv;
  u32 k;        // source has 2^(log_n - k) rows; position p reads source row (p >> k)
  int in_rev;   // ... bit-reversed over (log_n - k) bits
  int out_rev;  // destination row = bitrev_{log_n}(position)   (only with l0 == 0)
  int dif;      // 0: forward DIT butterflies, layers ascending; 1: inverse butterflies, descending
  int scale;    // last pass of a transform: 1 = multiply every output by scale_c, 2 = canonicalise
  int tw_smem;  // the pass's 2^r - 1 twiddles are the same for every element of a tile: staged in shared memory
  Fr scale_c;
};

constexpr int NTT_THREADS = 256;
constexpr u32 NTT_TILE_MAX = 2048;  // 64 KiB of shared memory per CTA -> 3 CTAs per SM
static u32 g_ntt_tile = 0;            // elements per tile in use (EON_NTT_TILE = 1024 | 2048)
static u32 ntt_tile_elems() {
  if (!g_ntt_tile) {
    const char* e = getenv("EON_NTT_TILE");
    g_ntt_tile = e ? (u32)atoi(e) : NTT_TILE_MAX;
    if (g_ntt_tile != 1024 && g_ntt_tile != 2048) g_ntt_tile = NTT_TILE_MAX;
  }
  return g_ntt_tile;
}

__device__ __forceinline__ void split_vidx(const PassParams& p, u64 vidx, u64& lo, u32& col) {
  if (p.w_shift >= 0) {
    lo = vidx >> p.w_shift;
    col = (u32)(vidx & (p.w - 1));
  } else {
    lo = vidx / p.w;
    col = (u32)(vidx - lo * p.w);
  }
}

// TWS (fixed-operand twiddles only): the twiddles of the pass are staged in shared memory.  Inside a tile the
// twiddle of layer t depends on the tile row only -- index ((row & (2^t - 1)) << l0) + lo, and lo (the position
// inside the 2^l0 block) is the same for all cv virtual columns of a tile when w is a multiple of cv (or l0 = 0) --
// so a tile uses 2^r - 1 pairs (8 KiB at r = 7) instead of fetching 48 bytes per butterfly from the L2 in the
// middle of the dependent IMAD chains.
template <int MINB, bool R4, bool SH, bool TWS = false>
__global__ void __launch_bounds__(NTT_THREADS, MINB) k_ntt_pass(const PassParams p) {
  typedef typename TwOf<SH>::T Tw;
  extern __shared__ uint4 smem[];
  const u32 R = 1u << p.r;
  const u32 cv = 1u << p.log_cv;
  const u32 tile_elems = R << p.log_cv;
  uint4* s_lo = smem;
  uint4* s_hi = smem + tile_elems + 4;  // +64 B: the two 16-byte planes of one element hit different banks
  uint4* s_tw = s_hi + tile_elems + 4;  // TWS: (2^r - 1) x 64 bytes, layer t at entry 2^t - 1
  const u32 tid = threadIdx.x;

  const u64 tile = blockIdx.x;
  const u64 vt = tile % p.tiles_v;
  const u64 hi = tile / p.tiles_v;
  const u64 v0 = vt << p.log_cv;
  const u32 ncv = (u32)min((u64)cv, p.V - v0);
  const u64 row_base = hi << p.r;  // tile row m is matrix "row group" row_base + m

  // ---- load tile (16-byte units, consecutive threads -> consecutive units) ----
  const bool plain_in = (p.k == 0 && !p.in_rev && p.ld_src == p.w);
  const bool plain_out = (p.ld_dst == p.w);
  // every 16-byte unit of the tile as one asynchronous copy straight into its shared-memory slot
  for (u32 u = tid; u < tile_elems * 2; u += NTT_THREADS) {
    const u32 e = u >> 1, half = u & 1;
    const u32 m = e >> p.log_cv, vc = e & (cv - 1);
    if (vc >= ncv) continue;
    const u64 vidx = v0 + vc;
    u64 src_elem;
    if (plain_in) {
      src_elem = (row_base + m) * p.V + vidx;
    } else {
      u64 lo;
      u32 col;
      split_vidx(p, vidx, lo, col);
      u64 pos = ((row_base + m) << p.l0) + lo;
      u64 srow = pos >> p.k;
      if (p.in_rev) {
        u32 bits = p.log_n - p.k;
        srow = bits ? (u64)(__brev((u32)srow) >> (32 - bits)) : 0;
      }
      src_elem = srow * p.ld_src + col;
    }
    cp_async16((half ? s_hi : s_lo) + e, p.src + src_elem * 2 + half);
  }
  if constexpr (TWS) {
    // the tile's twiddles: entry i = layer t = floor(log2(i + 1)), row bits jr = i + 1 - 2^t
    u64 lo_tile = 0;
    if (p.l0) {
      u32 col;
      split_vidx(p, v0, lo_tile, col);
    }
    const u32 ntw = R - 1;
    for (u32 u = tid; u < ntw * 4; u += NTT_THREADS) {
      const u32 i = u >> 2, q = u & 3;
      const u32 t = 31 - __clz(i + 1);
      const u32 jr = i + 1 - (1u << t);
      const u64 g = ((1ull << (p.l0 + t)) - 1) + ((u64)jr << p.l0) + lo_tile;
      cp_async16(s_tw + i * 4 + q, p.tw + g * 4 + q);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- butterfly layers ----
  // Two layers at a time on quartets held in registers (half the shared-memory traffic and barriers
  // of a layer-by-layer sweep), then a single radix-2 layer if r is odd.
  u32 step = 0;
  if (R4) {
    const u32 nq = tile_elems >> 2;
    for (; step + 1 < p.r; step += 2) {
      // forward: layers (t, t+1) ascending; inverse: layers (t+1, t) descending
      const u32 t = p.dif ? (p.r - 2 - step) : step;
      const u32 tmask = (1u << t) - 1;
      const u64 twA_base = (1ull << (p.l0 + t)) - 1;
      const u64 twB_base = (1ull << (p.l0 + t + 1)) - 1;
      for (u32 idx = tid; idx < nq; idx += NTT_THREADS) {
        u32 vc = idx & (cv - 1);
        if (vc >= ncv) continue;
        u32 b = idx >> p.log_cv;
        u32 i0 = ((b >> t) << (t + 2)) | (b & tmask);
        u32 e0 = (i0 << p.log_cv) + vc;
        u32 st1 = (1u << t) << p.log_cv;
        u32 e1 = e0 + st1, e2 = e0 + 2 * st1, e3 = e0 + 3 * st1;
        u64 j = (u64)(b & tmask) << p.l0;
        if (p.l0) {
          u64 lo;
          u32 col;
          split_vidx(p, v0 + vc, lo, col);
          j += lo;
        }
        Fr x0 = fr_from_units(s_lo[e0], s_hi[e0]);
        Fr x1 = fr_from_units(s_lo[e1], s_hi[e1]);
        Fr x2 = fr_from_units(s_lo[e2], s_hi[e2]);
        Fr x3 = fr_from_units(s_lo[e3], s_hi[e3]);
        if constexpr (SH && TWS) {
          const u32 jr = b & tmask;
          const uint4* tA = s_tw + (((1u << t) - 1) + jr) * 4;
          const uint4* tB0 = s_tw + (((2u << t) - 1) + jr) * 4;
          const uint4* tB1 = tB0 + (4u << t);
          if (!p.dif) {
            {
              Tw twA;
              twA.w = fr_lds(tA);
              twA.wq = fr_lds(tA + 2);
              bf_dit(x0, x1, twA);
              bf_dit(x2, x3, twA);
            }
            {
              Tw twB;
              twB.w = fr_lds(tB0);
              twB.wq = fr_lds(tB0 + 2);
              bf_dit(x0, x2, twB);
            }
            {
              Tw twB;
              twB.w = fr_lds(tB1);
              twB.wq = fr_lds(tB1 + 2);
              bf_dit(x1, x3, twB);
            }
          } else {
            {
              Tw twB;
              twB.w = fr_lds(tB0);
              twB.wq = fr_lds(tB0 + 2);
              bf_dif(x0, x2, twB);
            }
            {
              Tw twB;
              twB.w = fr_lds(tB1);
              twB.wq = fr_lds(tB1 + 2);
              bf_dif(x1, x3, twB);
            }
            {
              Tw twA;
              twA.w = fr_lds(tA);
              twA.wq = fr_lds(tA + 2);
              bf_dif(x0, x1, twA);
              bf_dif(x2, x3, twA);
            }
          }
        } else if constexpr (SH) {
          // 64-byte twiddle pairs: fetched right before their use, so that at most one is live beside the quartet
          if (!p.dif) {
            {
              const Tw twA = tw_ldg<SH>(p.tw, twA_base + j);
              bf_dit(x0, x1, twA);
              bf_dit(x2, x3, twA);
            }
            {
              const Tw twB0 = tw_ldg<SH>(p.tw, twB_base + j);
              bf_dit(x0, x2, twB0);
            }
            {
              const Tw twB1 = tw_ldg<SH>(p.tw, twB_base + j + (1ull << (p.l0 + t)));
              bf_dit(x1, x3, twB1);
            }
          } else {
            {
              const Tw twB0 = tw_ldg<SH>(p.tw, twB_base + j);
              bf_dif(x0, x2, twB0);
            }
            {
              const Tw twB1 = tw_ldg<SH>(p.tw, twB_base + j + (1ull << (p.l0 + t)));
              bf_dif(x1, x3, twB1);
            }
            {
              const Tw twA = tw_ldg<SH>(p.tw, twA_base + j);
              bf_dif(x0, x1, twA);
              bf_dif(x2, x3, twA);
            }
          }
        } else {
          const Tw twA = tw_ldg<SH>(p.tw, twA_base + j);
          const Tw twB0 = tw_ldg<SH>(p.tw, twB_base + j);
          const Tw twB1 = tw_ldg<SH>(p.tw, twB_base + j + (1ull << (p.l0 + t)));
          if (!p.dif) {
            bf_dit(x0, x1, twA);
            bf_dit(x2, x3, twA);
            bf_dit(x0, x2, twB0);
            bf_dit(x1, x3, twB1);
          } else {
            bf_dif(x0, x2, twB0);
            bf_dif(x1, x3, twB1);
            bf_dif(x0, x1, twA);
            bf_dif(x2, x3, twA);
          }
        }
        fr_to_units(x0, s_lo[e0], s_hi[e0]);
        fr_to_units(x1, s_lo[e1], s_hi[e1]);
        fr_to_units(x2, s_lo[e2], s_hi[e2]);
        fr_to_units(x3, s_lo[e3], s_hi[e3]);
      }
      __syncthreads();
    }
  }
  const u32 nbf = tile_elems >> 1;
  for (; step < p.r; step++) {
    const u32 t = p.dif ? (p.r - 1 - step) : step;
    const u32 tmask = (1u << t) - 1;
    const u64 tw_base = (1ull << (p.l0 + t)) - 1;
    for (u32 idx = tid; idx < nbf; idx += NTT_THREADS) {
      u32 vc = idx & (cv - 1);
      if (vc >= ncv) continue;
      u32 b = idx >> p.log_cv;
      u32 i0 = ((b >> t) << (t + 1)) | (b & tmask);
      u32 e0 = (i0 << p.log_cv) + vc;
      u32 e1 = e0 + ((1u << t) << p.log_cv);
      u64 j = (u64)(b & tmask) << p.l0;
      if (p.l0) {
        u64 lo;
        u32 col;
        split_vidx(p, v0 + vc, lo, col);
        j += lo;
      }
      Tw tw;
      if constexpr (SH && TWS) {
        const uint4* tp = s_tw + (((1u << t) - 1) + (b & tmask)) * 4;
        tw.w = fr_lds(tp);
        tw.wq = fr_lds(tp + 2);
      } else {
        tw = tw_ldg<SH>(p.tw, tw_base + j);
      }
      Fr a = fr_from_units(s_lo[e0], s_hi[e0]);
      Fr bb = fr_from_units(s_lo[e1], s_hi[e1]);
      Fr o0 = a, o1 = bb;
      if (!p.dif) bf_dit(o0, o1, tw);
      else bf_dif(o0, o1, tw);
      fr_to_units(o0, s_lo[e0], s_hi[e0]);
      fr_to_units(o1, s_lo[e1], s_hi[e1]);
    }
    __syncthreads();
  }

  // ---- store tile ----
  if (!p.scale) {
    for (u32 u = tid; u < tile_elems * 2; u += NTT_THREADS) {
      u32 e = u >> 1, half = u & 1;
      u32 m = e >> p.log_cv, vc = e & (cv - 1);
      if (vc >= ncv) continue;
      u64 vidx = v0 + vc;
      u64 dst_elem;
      if (p.out_rev) {
        u64 pos = row_base + m;  // l0 == 0, V == w
        u64 drow = p.log_n ? (u64)(__brev((u32)pos) >> (32 - p.log_n)) : 0;
        dst_elem = drow * p.ld_dst + vidx;
      } else if (plain_out) {
        dst_elem = (row_base + m) * p.V + vidx;
      } else {
        u64 lo;
        u32 col;
        split_vidx(p, vidx, lo, col);
        dst_elem = (((row_base + m) << p.l0) + lo) * p.ld_dst + col;
      }
      p.dst[dst_elem * 2 + half] = (half ? s_hi : s_lo)[e];
    }
  } else {
    for (u32 e = tid; e < tile_elems; e += NTT_THREADS) {
      u32 m = e >> p.log_cv, vc = e & (cv - 1);
      if (vc >= ncv) continue;
      u64 vidx = v0 + vc;
      u64 dst_elem;
      if (p.out_rev) {
        u64 pos = row_base + m;
        u64 drow = p.log_n ? (u64)(__brev((u32)pos) >> (32 - p.log_n)) : 0;
        dst_elem = drow * p.ld_dst + vidx;
      } else if (plain_out) {
        dst_elem = (row_base + m) * p.V + vidx;
      } else {
        u64 lo;
        u32 col;
        split_vidx(p, vidx, lo, col);
        dst_elem = (((row_base + m) << p.l0) + lo) * p.ld_dst + col;
      }
      Fr x = fr_from_units(s_lo[e], s_hi[e]);
      x = (p.scale == 1) ? fp_mul(p.scale_c, x) : fp_canon_4p<FrParams>(x.v);
      uint4 lo, hi4;
      fr_to_units(x, lo, hi4);
      p.dst[dst_elem * 2] = lo;
      p.dst[dst_elem * 2 + 1] = hi4;
    }
  }
}

// ---- persistent, double-buffered form of the pass (fixed-operand twiddles in shared memory only) ----------
// k_ntt_pass gives a CTA ONE tile: load, wait, butterflies, store -- the integer pipe of an SM idles whenever both of
// its resident CTAs are in a memory phase at once (ncu: sm__pipe_fmaheavy_cycles_active 75-81 %).  Here a CTA walks
// tiles blockIdx.x, blockIdx.x + gridDim.x, ... with TWO shared-memory buffers: the asynchronous copies of tile
// i + 1 (and of its twiddles) are issued before the butterflies of tile i start and are only waited for when tile i
// has been stored, so every CTA always has arithmetic to issue.  Same butterflies, same order, same results.
struct TileGeom {
  u64 v0, row_base;
  u32 ncv;
};
__device__ __forceinline__ TileGeom tile_geom(const PassParams& p, u64 tile) {
  TileGeom g;
  const u64 vt = tile % p.tiles_v;
  const u64 hi = tile / p.tiles_v;
  g.v0 = vt << p.log_cv;
  g.ncv = (u32)min((u64)(1u << p.log_cv), p.V - g.v0);
  g.row_base = hi << p.r;
  return g;
}

template <int THREADS>
__device__ __forceinline__ void db_issue_tile(const PassParams& p, u64 tile, uint4* s_lo, uint4* s_hi, uint4* s_tw,
                                              u32 tid) {
  const TileGeom g = tile_geom(p, tile);
  const u32 cv = 1u << p.log_cv;
  const u32 tile_elems = (1u << p.r) << p.log_cv;
  const bool plain_in = (p.k == 0 && !p.in_rev && p.ld_src == p.w);
  for (u32 u = tid; u < tile_elems * 2; u += THREADS) {
    const u32 e = u >> 1, half = u & 1;
    const u32 m = e >> p.log_cv, vc = e & (cv - 1);
    if (vc >= g.ncv) continue;
    const u64 vidx = g.v0 + vc;
    u64 src_elem;
    if (plain_in) {
      src_elem = (g.row_base + m) * p.V + vidx;
    } else {
      u64 lo;
      u32 col;
      split_vidx(p, vidx, lo, col);
      u64 pos = ((g.row_base + m) << p.l0) + lo;
      u64 srow = pos >> p.k;
      if (p.in_rev) {
        u32 bits = p.log_n - p.k;
        srow = bits ? (u64)(__brev((u32)srow) >> (32 - bits)) : 0;
      }
      src_elem = srow * p.ld_src + col;
    }
    cp_async16((half ? s_hi : s_lo) + e, p.src + src_elem * 2 + half);
  }
  u64 lo_tile = 0;
  if (p.l0) {
    u32 col;
    split_vidx(p, g.v0, lo_tile, col);
  }
  const u32 ntw = (1u << p.r) - 1;
  for (u32 u = tid; u < ntw * 4; u += THREADS) {
    const u32 i = u >> 2, q = u & 3;
    const u32 t = 31 - __clz(i + 1);
    const u32 jr = i + 1 - (1u << t);
    const u64 gi = ((1ull << (p.l0 + t)) - 1) + ((u64)jr << p.l0) + lo_tile;
    cp_async16(s_tw + i * 4 + q, p.tw + gi * 4 + q);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int THREADS>
__device__ __forceinline__ void db_butterflies(const PassParams& p, u32 ncv, uint4* s_lo, uint4* s_hi,
                                               const uint4* s_tw, u32 tid) {
  const u32 cv = 1u << p.log_cv;
  const u32 tile_elems = (1u << p.r) << p.log_cv;
  u32 step = 0;
  const u32 nq = tile_elems >> 2;
  for (; step + 1 < p.r; step += 2) {
    const u32 t = p.dif ? (p.r - 2 - step) : step;
    const u32 tmask = (1u << t) - 1;
    for (u32 idx = tid; idx < nq; idx += THREADS) {
      const u32 vc = idx & (cv - 1);
      if (vc >= ncv) continue;
      const u32 b = idx >> p.log_cv;
      const u32 i0 = ((b >> t) << (t + 2)) | (b & tmask);
      const u32 e0 = (i0 << p.log_cv) + vc;
      const u32 st1 = (1u << t) << p.log_cv;
      const u32 e1 = e0 + st1, e2 = e0 + 2 * st1, e3 = e0 + 3 * st1;
      Fr x0 = fr_from_units(s_lo[e0], s_hi[e0]);
      Fr x1 = fr_from_units(s_lo[e1], s_hi[e1]);
      Fr x2 = fr_from_units(s_lo[e2], s_hi[e2]);
      Fr x3 = fr_from_units(s_lo[e3], s_hi[e3]);
      const u32 jr = b & tmask;
      const uint4* tA = s_tw + (((1u << t) - 1) + jr) * 4;
      const uint4* tB0 = s_tw + (((2u << t) - 1) + jr) * 4;
      const uint4* tB1 = tB0 + (4u << t);
      if (!p.dif) {
        {
          TwS twA;
          twA.w = fr_lds(tA);
          twA.wq = fr_lds(tA + 2);
          bf_dit(x0, x1, twA);
          bf_dit(x2, x3, twA);
        }
        {
          TwS twB;
          twB.w = fr_lds(tB0);
          twB.wq = fr_lds(tB0 + 2);
          bf_dit(x0, x2, twB);
        }
        {
          TwS twB;
          twB.w = fr_lds(tB1);
          twB.wq = fr_lds(tB1 + 2);
          bf_dit(x1, x3, twB);
        }
      } else {
        {
          TwS twB;
          twB.w = fr_lds(tB0);
          twB.wq = fr_lds(tB0 + 2);
          bf_dif(x0, x2, twB);
        }
        {
          TwS twB;
          twB.w = fr_lds(tB1);
          twB.wq = fr_lds(tB1 + 2);
          bf_dif(x1, x3, twB);
        }
        {
          TwS twA;
          twA.w = fr_lds(tA);
          twA.wq = fr_lds(tA + 2);
          bf_dif(x0, x1, twA);
          bf_dif(x2, x3, twA);
        }
      }
      fr_to_units(x0, s_lo[e0], s_hi[e0]);
      fr_to_units(x1, s_lo[e1], s_hi[e1]);
      fr_to_units(x2, s_lo[e2], s_hi[e2]);
      fr_to_units(x3, s_lo[e3], s_hi[e3]);
    }
    __syncthreads();
  }
  const u32 nbf = tile_elems >> 1;
  for (; step < p.r; step++) {
    const u32 t = p.dif ? (p.r - 1 - step) : step;
    const u32 tmask = (1u << t) - 1;
    for (u32 idx = tid; idx < nbf; idx += THREADS) {
      const u32 vc = idx & (cv - 1);
      if (vc >= ncv) continue;
      const u32 b = idx >> p.log_cv;
      const u32 i0 = ((b >> t) << (t + 1)) | (b & tmask);
      const u32 e0 = (i0 << p.log_cv) + vc;
      const u32 e1 = e0 + ((1u << t) << p.log_cv);
      const uint4* tp = s_tw + (((1u << t) - 1) + (b & tmask)) * 4;
      TwS tw;
      tw.w = fr_lds(tp);
      tw.wq = fr_lds(tp + 2);
      Fr o0 = fr_from_units(s_lo[e0], s_hi[e0]);
      Fr o1 = fr_from_units(s_lo[e1], s_hi[e1]);
      if (!p.dif) bf_dit(o0, o1, tw);
      else bf_dif(o0, o1, tw);
      fr_to_units(o0, s_lo[e0], s_hi[e0]);
      fr_to_units(o1, s_lo[e1], s_hi[e1]);
    }
    __syncthreads();
  }
}

template <int THREADS>
__device__ __forceinline__ void db_store_tile(const PassParams& p, u64 tile, const uint4* s_lo, const uint4* s_hi,
                                              u32 tid) {
  const TileGeom g = tile_geom(p, tile);
  const u32 cv = 1u << p.log_cv;
  const u32 tile_elems = (1u << p.r) << p.log_cv;
  const bool plain_out = (p.ld_dst == p.w);
  auto dst_of = [&](u32 m, u64 vidx) -> u64 {
    if (p.out_rev) {
      const u64 pos = g.row_base + m;  // l0 == 0, V == w
      const u64 drow = p.log_n ? (u64)(__brev((u32)pos) >> (32 - p.log_n)) : 0;
      return drow * p.ld_dst + vidx;
    }
    if (plain_out) return (g.row_base + m) * p.V + vidx;
    u64 lo;
    u32 col;
    split_vidx(p, vidx, lo, col);
    return (((g.row_base + m) << p.l0) + lo) * p.ld_dst + col;
  };
  if (!p.scale) {
    for (u32 u = tid; u < tile_elems * 2; u += THREADS) {
      const u32 e = u >> 1, half = u & 1;
      const u32 m = e >> p.log_cv, vc = e & (cv - 1);
      if (vc >= g.ncv) continue;
      p.dst[dst_of(m, g.v0 + vc) * 2 + half] = (half ? s_hi : s_lo)[e];
    }
  } else {
    for (u32 e = tid; e < tile_elems; e += THREADS) {
      const u32 m = e >> p.log_cv, vc = e & (cv - 1);
      if (vc >= g.ncv) continue;
      const u64 d = dst_of(m, g.v0 + vc);
      Fr x = fr_from_units(s_lo[e], s_hi[e]);
      x = (p.scale == 1) ? fp_mul(p.scale_c, x) : fp_canon_4p<FrParams>(x.v);
      uint4 lo, hi4;
      fr_to_units(x, lo, hi4);
      p.dst[d * 2] = lo;
      p.dst[d * 2 + 1] = hi4;
    }
  }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_ntt_pass_db(const PassParams p, const u64 ntiles) {
  extern __shared__ uint4 smem[];
  const u32 tile_elems = (1u << p.r) << p.log_cv;
  const u32 buf_units = 2 * (tile_elems + 4) + (4u << p.r);  // two 16-byte planes (+64 B skew each) | twiddle pairs
  const u32 tid = threadIdx.x;
  u64 tile = blockIdx.x;
  if (tile >= ntiles) return;
  u32 b = 0;
  db_issue_tile<THREADS>(p, tile, smem, smem + tile_elems + 4, smem + 2 * (tile_elems + 4), tid);
  for (; tile < ntiles; tile += gridDim.x, b ^= 1) {
    uint4* s_lo = smem + b * buf_units;
    uint4* s_hi = s_lo + tile_elems + 4;
    uint4* s_tw = s_hi + tile_elems + 4;
    const u64 next = tile + gridDim.x;
    if (next < ntiles) {
      // the other buffer was last read by the store of the previous iteration, which ended with a barrier
      uint4* n_lo = smem + (b ^ 1) * buf_units;
      db_issue_tile<THREADS>(p, next, n_lo, n_lo + tile_elems + 4, n_lo + 2 * (tile_elems + 4), tid);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    db_butterflies<THREADS>(p, tile_geom(p, tile).ncv, s_lo, s_hi, s_tw, tid);
    db_store_tile<THREADS>(p, tile, s_lo, s_hi, tid);
    __syncthreads();
  }
}

// ---- pass planning ------------------------------------------------------------------------
struct PassPlan {
  u32 l0, r, log_cv;
};

static u32 log_cv_for(u64 V, u32 cap = 4) {
  u32 lc = 0;
  while (lc < cap && (1ull << lc) < V) lc++;
  return lc;  // cv = min(2^cap, next_pow2(V))
}
static u32 ilog2_u32(u32 x) {
  u32 l = 0;
  while ((1u << (l + 1)) <= x) l++;
  return l;
}

// split layers [first, last) into passes (ascending l0)
static std::vector<PassPlan> plan_passes(u32 first, u32 last, size_t w, u32 tile_log = 0, u32 cv_cap = 4) {
  std::vector<PassPlan> out;
  if (!tile_log) tile_log = ilog2_u32(ntt_tile_elems());
  auto log_cv_for = [&](u64 V) { return eon::log_cv_for(V, cv_cap); };
  auto rmax_at = [&](u32 l0) {
    u64 V = ((u64)w) << l0;
    return tile_log - log_cv_for(V);
  };
  if (last <= first) {
    out.push_back({first, 0, log_cv_for(((u64)w) << first)});
    return out;
  }
  // Narrow matrices (w < 16: the column shard of one GPU when 16 trace columns are split over 4-8 GPUs): at l0 =
  // first a tile row holds only w << first < 16 elements, so an evenly sized first pass would stage a tile of a
  // few hundred elements for 256 threads.  Give that pass all the layers a full tile can hold instead (r = 10 at
  // w = 2) and spread the remaining layers evenly.
  if (log_cv_for(((u64)w) << first) < cv_cap && last - first > rmax_at(first)) {
    const u32 r0 = rmax_at(first);
    out.push_back({first, r0, log_cv_for(((u64)w) << first)});
    std::vector<PassPlan> rest = plan_passes(first + r0, last, w, tile_log, cv_cap);
    out.insert(out.end(), rest.begin(), rest.end());
    return out;
  }
  // greedy count
  u32 n = 0;
  for (u32 l = first; l < last; n++) l += std::min(rmax_at(l), last - l);
  // even distribution subject to per-pass maxima
  u32 l = first;
  for (u32 i = 0; i < n; i++) {
    u32 remaining = last - l;
    u32 target = (remaining + (n - i) - 1) / (n - i);
    u32 r = std::min(std::min(rmax_at(l), target), remaining);
    out.push_back({l, r, log_cv_for(((u64)w) << l)});
    l += r;
  }
  if (l < last) {  // distribution fell short because of a small early maximum: finish greedily
    while (l < last) {
      u32 r = std::min(rmax_at(l), last - l);
      out.push_back({l, r, log_cv_for(((u64)w) << l)});
      l += r;
    }
  }
  return out;
}


// MEASURED ON B200 (profiles/r02r_ntt_db.txt, 2^20 x 16 commit + LDE): passes 8.79 ms with k_ntt_pass, 9.73 ms
// double-buffered at 256 threads / 1024-element tiles, 10.41 ms at 512 threads / 2048-element tiles -> OFF by
// default.  The one-tile-per-CTA kernel at 2 CTAs per SM already hides its loads behind the other CTA; what the
// persistent form adds is barriers (half the work per barrier interval, or 16 warps per barrier instead of 8).
// Double-buffered persistent passes (k_ntt_pass_db).  EON_NTT_DB: 0 = off, 1 = 256 threads / 1024-element tiles /
// 2 CTAs per SM, 2 = 512 threads / 2048-element tiles / 1 CTA per SM.  Used for fixed-operand twiddles on matrices
// whose width is a power of two >= 8 (every tile then has ONE twiddle set, see k_ntt_pass) and passes with at least
// two tiles per resident CTA; everything else takes k_ntt_pass.
static int ntt_db_mode() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("EON_NTT_DB");
    m = e ? atoi(e) : EON_NTT_DB_DEFAULT;
    if (m < 0 || m > 2) m = 0;
  }
  return m;
}
static bool ntt_db_shape(size_t w, bool sh) { return sh && ntt_db_mode() && w >= 8 && (w & (w - 1)) == 0; }
static std::vector<PassPlan> plan_passes_for(u32 first, u32 last, size_t w, bool sh) {
  if (!ntt_db_shape(w, sh)) {
    std::vector<PassPlan> plan = plan_passes(first, last, w);
    // Narrow matrices (w < 16: a column group of the host-buffer pipeline, the shard of one GPU at N = 4, 8): a
    // pass with fewer layers than its tile could hold (the 6 + 5 layers behind a 9-layer first pass at w = 4) would
    // stage 1024 or 512 elements for 256 threads.  Such a pass takes more virtual columns instead (up to 64: rows of
    // the 2^l0-blocks are contiguous in memory, so wider tile rows are longer contiguous runs), keeping the tile at
    // its full size.  (With w >= 16 the tile row stays at 16 columns = one twiddle set per tile, see k_ntt_pass.)
    static const int widen = getenv("EON_NTT_WIDEN") ? atoi(getenv("EON_NTT_WIDEN")) : 1;
    if (w < 16 && widen) {
      const u32 tile_log = ilog2_u32(ntt_tile_elems());
      for (PassPlan& pl : plan) {
        const u64 V = ((u64)w) << pl.l0;
        while (pl.log_cv < 6 && pl.log_cv + pl.r < tile_log && (2ull << pl.log_cv) <= V) pl.log_cv++;
      }
    }
    return plan;
  }
  const u32 tile_log = ntt_db_mode() == 1 ? 10 : 11;
  // layers per pass as if every tile row were 8 elements wide (r <= tile_log - 3); a pass with fewer layers than
  // that widens its rows (cv = 16) so that the tile keeps its size
  std::vector<PassPlan> plan = plan_passes(first, last, w, tile_log, 3);
  for (PassPlan& pl : plan) {
    const u64 V = ((u64)w) << pl.l0;
    u32 lc = pl.log_cv;
    while (lc < 4 && lc + pl.r < tile_log && (2ull << lc) <= V && (2ull << lc) <= w) lc++;
    pl.log_cv = lc;
  }
  return plan;
}

static int launch_pass(eon_ctx* ctx, PassParams& p, const PassPlan& pl, unsigned log_n, size_t w, bool p_shoup) {
  p.log_n = log_n;
  p.w = (u32)w;
  p.w_shift = (w & (w - 1)) == 0 ? (int)ilog2_u32((u32)w) : -1;
  p.l0 = pl.l0;
  p.r = pl.r;
  p.log_cv = pl.log_cv;
  p.V = ((u64)w) << pl.l0;
  u64 cv = 1ull << pl.log_cv;
  p.tiles_v = (p.V + cv - 1) / cv;
  u64 tiles_hi = 1ull << (log_n - pl.l0 - pl.r);
  u64 grid = tiles_hi * p.tiles_v;
  if (grid == 0 || grid > 0x7fffffffull) return fail(ctx, EON_ERR_BAD_ARG, "ntt: grid too large");
  size_t smem = ((size_t)1 << (pl.r + pl.log_cv)) * 32 + 128;
  // twiddles in shared memory: fixed-operand form, and one twiddle set per tile (see k_ntt_pass)
  static const int tws_env = getenv("EON_NTT_TWS") ? atoi(getenv("EON_NTT_TWS")) : 1;
  const bool tws = p_shoup && tws_env && pl.r >= 1 &&
                   (pl.l0 == 0 || (p.w_shift >= 0 && w >= ((size_t)1 << pl.log_cv)));
  p.tw_smem = tws ? 1 : 0;
  if (tws) smem += (((size_t)1 << pl.r) - 1) * 64;
  static int radix4 = -1;
  if (radix4 < 0) {
    const char* e = getenv("EON_NTT_RADIX4");
    radix4 = e ? atoi(e) : 1;
  }
  static int minb = -1;
  if (minb < 0) {
    const char* e = getenv("EON_NTT_MINB");
    minb = e ? atoi(e) : (ntt_tile_elems() == 1024 ? 4 : 3);
  }
  if (!ctx->ntt_attr_set) {  // per context: the attribute belongs to the context's device
    const int mx = (int)(NTT_TILE_MAX * 32 + 128 + 2048 * 64);  // tile + (TWS) up to 2^11 - 1 twiddle pairs
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<3, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ctx->ntt_attr_set = true;
  }
  if (tws && ntt_db_shape(w, p_shoup)) {
    const int mode = ntt_db_mode();
    const u64 resident = (u64)ctx->num_sms * (mode == 1 ? 2 : 1);
    const size_t buf_units = 2 * (((size_t)1 << (pl.r + pl.log_cv)) + 4) + ((size_t)4 << pl.r);
    const size_t smem_db = 2 * buf_units * 16;
    if (grid >= 2 * resident && smem_db <= (mode == 1 ? 100u : 200u) * 1024) {
      if (!ctx->ntt_db_attr_set) {
        EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass_db<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        EON_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass_db<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->ntt_db_attr_set = true;
      }
      if (mode == 1) k_ntt_pass_db<256, 2><<<(unsigned)resident, 256, smem_db, ctx->stream>>>(p, grid);
      else k_ntt_pass_db<512, 1><<<(unsigned)resident, 512, smem_db, ctx->stream>>>(p, grid);
      EON_LAUNCHED(ctx);
      return EON_OK;
    }
  }
  if (p_shoup) {  // fixed-operand twiddles: radix-4 quartets only; 2 CTAs per SM unless EON_NTT_MINB asks for 3
    static int minb_sh = -1;
    if (minb_sh < 0) {
      const char* e = getenv("EON_NTT_MINB");
      minb_sh = e ? atoi(e) : 2;
    }
    if (tws) {
      if (minb_sh >= 3) k_ntt_pass<3, true, true, true><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
      else k_ntt_pass<2, true, true, true><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
    } else if (minb_sh >= 3) k_ntt_pass<3, true, true><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
    else k_ntt_pass<2, true, true><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
  } else if (radix4) {
    if (minb >= 4) k_ntt_pass<4, true, false><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
    else if (minb == 3) k_ntt_pass<3, true, false><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
    else k_ntt_pass<2, true, false><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
  } else {
    if (minb >= 4) k_ntt_pass<4, false, false><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
    else k_ntt_pass<3, false, false><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(p);
  }
  EON_LAUNCHED(ctx);
  return EON_OK;
}

bool ntt_twiddles_are_fixed_operand() { return ntt_use_shoup(); }

int ntt_forward(eon_ctx* ctx, const Fr* d_src, Fr* d_dst, unsigned log_n, unsigned k, size_t width, const Fr& shift,
                Layout src_layout, size_t ld_src, size_t ld_dst) {
  if (ld_src == 0) ld_src = width;
  if (ld_dst == 0) ld_dst = width;
  if (width == 0) return EON_OK;
  if (log_n > 28) return fail(ctx, EON_ERR_TWO_ADICITY, "transform size exceeds 2^28 (Fr::TWO_ADICITY)");
  if (k > log_n) return fail(ctx, EON_ERR_BAD_ARG, "ntt_forward: added bits exceed transform size");
  if (width > 0xffffffffull) return fail(ctx, EON_ERR_BAD_ARG, "ntt: width too large");
  const Fr* tw = nullptr;
  const bool sh = ntt_use_shoup();
  EON_TRY(get_twiddles(ctx, log_n, shift, 0, sh, &tw));
  std::vector<PassPlan> plan = plan_passes_for(k, log_n, width, sh);
  phase_begin(ctx, PH_NTT_PASSES);
  for (size_t i = 0; i < plan.size(); i++) {
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.tw = (const uint4*)tw;
    p.dif = 0;
    if (i == 0) {
      p.src = (const uint4*)d_src;
      p.ld_src = ld_src;
      p.k = k;
      p.in_rev = (src_layout == LAYOUT_NATURAL) ? 1 : 0;
    } else {
      p.src = (const uint4*)d_dst;
      p.ld_src = ld_dst;
    }
    p.dst = (uint4*)d_dst;
    p.ld_dst = ld_dst;
    if (i + 1 == plan.size()) p.scale = 2;  // lazily reduced values leave the transform canonical
    EON_TRY(launch_pass(ctx, p, plan[i], log_n, width, sh));
  }
  phase_end(ctx, PH_NTT_PASSES);
  return EON_OK;
}

int ntt_inverse(eon_ctx* ctx, const Fr* d_src, Fr* d_dst, unsigned log_n, size_t width, const Fr& shift,
                Layout dst_layout, size_t ld_src, size_t ld_dst) {
  if (ld_src == 0) ld_src = width;
  if (ld_dst == 0) ld_dst = width;
  if (width == 0) return EON_OK;
  if (log_n > 28) return fail(ctx, EON_ERR_TWO_ADICITY, "transform size exceeds 2^28 (Fr::TWO_ADICITY)");
  if (width > 0xffffffffull) return fail(ctx, EON_ERR_BAD_ARG, "ntt: width too large");
  const Fr* tw = nullptr;
  const bool sh = ntt_use_shoup();
  EON_TRY(get_twiddles(ctx, log_n, shift, 1, sh, &tw));
  std::vector<PassPlan> plan = plan_passes_for(0, log_n, width, sh);
  const size_t np = plan.size();
  const bool out_rev = (dst_layout == LAYOUT_NATURAL);
  Fr* work = d_dst;
  size_t ld_work = ld_dst;
  if (np > 1 && out_rev) {
    void* tmp = nullptr;
    EON_TRY(scratch_get(ctx, SC_NTT_TMP, ((size_t)width << log_n) * sizeof(Fr), &tmp));
    work = (Fr*)tmp;
    ld_work = width;
  }
  Fr n_inv = Fr::one();
  if (log_n) n_inv = fp_inv(fp_from_u64<FrParams>(1ull << log_n));
  phase_begin(ctx, PH_NTT_PASSES);
  for (size_t s = 0; s < np; s++) {
    const PassPlan& pl = plan[np - 1 - s];  // descending layers
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.tw = (const uint4*)tw;
    p.dif = 1;
    const bool last = (s == np - 1);
    p.src = (const uint4*)(s == 0 ? d_src : work);
    p.ld_src = (s == 0) ? ld_src : ld_work;
    p.dst = (uint4*)(last ? d_dst : work);
    p.ld_dst = last ? ld_dst : ld_work;
    if (last) {
      p.out_rev = out_rev ? 1 : 0;
      if (log_n) {
        p.scale = 1;
        p.scale_c = n_inv;
      } else {
        p.scale = 2;
      }
    }
    EON_TRY(launch_pass(ctx, p, pl, log_n, width, sh));
  }
  phase_end(ctx, PH_NTT_PASSES);
  return EON_OK;
}

}  // namespace eon
