// Shared host-side plumbing of libeon_kzg: context, error handling, workspace, caches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/eon_kzg.h"
#include "ec.cuh"

namespace eon {

enum Phase {
  PH_NTT_TWIDDLE = 0,
  PH_NTT_PASSES,
  PH_MSM_DIGITS,
  PH_MSM_SCAN,
  PH_MSM_SCATTER,
  PH_MSM_ACCUM,
  PH_MSM_REDUCE,
  PH_QUOTIENT,
  PH_MSM_TREE_FWD,   // the three below are sub-intervals of PH_MSM_ACCUM (msm_tree.cu)
  PH_MSM_TREE_INV,
  PH_MSM_TREE_BWD,
  PH_MSM_FINISH,
  PH_COUNT
};

struct TwiddleKey {
  unsigned log_n;
  int inverse;
  u32 shift[8];
  bool operator<(const TwiddleKey& o) const {
    if (log_n != o.log_n) return log_n < o.log_n;
    if (inverse != o.inverse) return inverse < o.inverse;
    return memcmp(shift, o.shift, sizeof(shift)) < 0;
  }
};

struct ProverMatrix {
  Fr* d_coeffs;  // natural order, rows x width
  size_t rows;   // any height for KzgMmcs matrices (kzg/src/mmcs.rs:168-190); 2^log_h for KzgPcs
  unsigned log_h;  // log2(rows), or NOT_POW2 when rows is not a power of two (no NTT on it)
  size_t width;
  size_t cap;    // bytes allocated behind d_coeffs
};

constexpr unsigned NOT_POW2 = 0xffffffffu;

// grow-only device scratch buffers, one per role, reused across calls (no allocation in steady state)
struct Scratch {
  void* ptr = nullptr;
  size_t cap = 0;
};

enum ScratchId {
  SC_NTT_TMP = 0,
  SC_MSM_HIST,
  SC_MSM_CURSOR,
  SC_MSM_ENTRIES,
  SC_MSM_BUCKETS,
  SC_MSM_TASKS,
  SC_MSM_TASKPART,
  SC_MSM_PARTIALS,
  SC_MSM_SEGSUM,
  SC_MSM_RESULT,
  SC_MSM_MISC,
  SC_MSM_ORDER,
  SC_MSM_TREE_A,
  SC_MSM_TREE_B,
  SC_MSM_TREE_T,
  SC_MSM_TREE_P,
  SC_MSM_SORT_REGION,
  SC_MSM_SORT_PAY,
  SC_MSM_SORT_KEY,
  SC_MSM_SORT_MAT,
  SC_MSM_SEGTOTAL,
  SC_MSM_SLICE,
  SC_IO_A,
  SC_IO_B,
  SC_QUOT,
  SC_QUOT_AUX,
  SC_SMALL,
  SC_COUNT
};

}  // namespace eon

struct eon_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  std::string last_error;
  uint64_t launches = 0;
  std::mutex mu;  // callers may share a ctx across threads: one call at a time

  // two banks: bank 1 is the workspace of the second half-batch when an MSM over few columns runs as two concurrent
  // halves on two streams (msm_run); everything else lives in bank 0
  eon::Scratch scratch[2][eon::SC_COUNT];
  int bank = 0;
  std::map<eon::TwiddleKey, eon::Fr*> twiddles;
  size_t twiddle_bytes = 0;  // device bytes behind `twiddles` (bounded: see get_twiddles)
  // opt-in shared-memory sizes (cudaFuncSetAttribute) are per device: set once per context, not per process
  // run once by msm_tree_rounds right after round 0 of the next MSM has been queued (then cleared): the device-
  // resident commit + LDE queues its LDE transform there, behind round 0 (see kzg_commit_locked)
  std::function<int()> after_round0;
  bool ntt_attr_set = false;
  bool ntt_db_attr_set = false;
  bool sort_attr_set = false;

  eon::G1Affine* d_srs = nullptr;
  size_t srs_n = 0;
  // window tables tab[t][i] = 2^(c t) * srs[i] (t < ceil(256/c)), built once per SRS; null = none
  eon::G1Affine* d_srs_tab = nullptr;
  unsigned srs_tab_c = 0;
  // a second table set over the SRS index range [rng_first, rng_first + rng_n) only, with a window sized for that
  // range: the shard this GPU owns in an index-range sharded MSM (eon_srs_set_range_tables)
  eon::G1Affine* d_rng_tab = nullptr;
  size_t rng_first = 0, rng_n = 0;
  unsigned rng_c = 0;
  // batched-affine pairwise rounds before the XYZZ finisher: -1 = automatic (msm_pick_rounds)
  int msm_rounds = -1;
  int msm_sort_mode = -1;        // -1 automatic, 0 one-pass atomic scatter, 1 two-pass coalesced sort
  int msm_slice_mode = -1;       // round 0 of the pairwise rounds by table slice: -1 automatic, 0 off, 1 on
  unsigned msm_rounds_used = 0;  // rounds of the most recent MSM (reporting)
  unsigned msm_c_used = 0;       // window bits of the most recent MSM (reporting)

  // second stream + events: the host-buffer entry points move column groups over PCIe while the
  // previous group computes (created on first use)
  cudaStream_t copy_stream = nullptr;
  cudaStream_t copy_stream2 = nullptr;
  cudaStream_t aux_stream = nullptr;    // second compute stream: the hinted LDE runs beside the MSM, whose
                                        // sort / gather phases leave the integer pipe idle  // opposite PCIe direction (downloads while uploads are in flight)
  cudaStream_t prio_stream = nullptr;   // high-priority compute stream: the MSM while an LDE transform runs beside it
  cudaStream_t split_stream = nullptr;  // second half-batch of an MSM over few columns (msm_run), high priority
  cudaEvent_t ev_split[2] = {nullptr, nullptr};
  cudaEvent_t ev_stagger = nullptr;     // set by msm_run for ONE msm_batch: recorded after that batch's sort
  // While two half-batches share the GPU, the single-warp phases of one half (inversion trees, bucket reduction)
  // would queue behind the wide grids of the other; they run on a stream of the HIGHEST priority (one per half),
  // the wide kernels one level below (tiny_on: set by msm_run around the two msm_batch calls).
  cudaStream_t tiny_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_tiny[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  bool tiny_on = false;
  int msm_split_mode = -1;              // -1 automatic (2..4 columns, >= 2^16 points), 0 never, 1 whenever >= 2 columns
  cudaEvent_t ev_pipe[20] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  std::map<eon_handle, eon::ProverMatrix> handles;
  eon_handle next_handle = 1;
  // freed coefficient buffers kept for the next commit of a similar size (a prover commits and
  // frees matrices of the same shape over and over; cudaMalloc/cudaFree would synchronise)
  std::vector<std::pair<size_t, void*>> coeff_pool;

  // per-phase device timing: every phase_begin/phase_end pair since the last eon_phase_reset
  // is kept; eon_last_phase_ms sums them (events are pooled and reused)
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs[eon::PH_COUNT];
  cudaEvent_t ev_open[eon::PH_COUNT];
};

namespace eon {

inline int fail(eon_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

#define EON_CUDA(ctx, expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      char _b[512];                                                                                  \
      snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return eon::fail(ctx, _e == cudaErrorMemoryAllocation ? EON_ERR_OOM : EON_ERR_CUDA, _b);       \
    }                                                                                                \
  } while (0)

#define EON_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != EON_OK) return _rc; \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define EON_LAUNCHED(ctx)                         \
  do {                                            \
    (ctx)->launches++;                            \
    EON_CUDA(ctx, cudaPeekAtLastError());         \
  } while (0)

inline int scratch_get(eon_ctx* ctx, int id, size_t bytes, void** out) {
  Scratch& s = ctx->scratch[ctx->bank][id];
  if (bytes > s.cap) {
    if (s.ptr) {
      // outstanding work on ANY stream of the context (caller's, auxiliary, high-priority, copy streams) may still
      // use the old buffer: growing a scratch buffer is rare, so wait for the whole device
      EON_CUDA(ctx, cudaDeviceSynchronize());
      EON_CUDA(ctx, cudaFree(s.ptr));
      s.ptr = nullptr;
      s.cap = 0;
    }
    size_t cap = bytes + (bytes >> 3) + 256;
    EON_CUDA(ctx, cudaMalloc(&s.ptr, cap));
    s.cap = cap;
  }
  *out = s.ptr;
  return EON_OK;
}

// Latency-critical launches of an MSM (see eon_ctx::tiny_stream): tiny_begin() returns the stream to launch them on
// (ordered after what ctx->stream has queued so far), tiny_end() orders ctx->stream after them.  Without a split
// in flight both are no-ops on ctx->stream.
inline cudaStream_t tiny_begin(eon_ctx* ctx) {
  if (!ctx->tiny_on) return ctx->stream;
  cudaStream_t ts = ctx->tiny_stream[ctx->bank];
  cudaEventRecord(ctx->ev_tiny[ctx->bank][0], ctx->stream);
  cudaStreamWaitEvent(ts, ctx->ev_tiny[ctx->bank][0], 0);
  return ts;
}
inline void tiny_end(eon_ctx* ctx) {
  if (!ctx->tiny_on) return;
  cudaEventRecord(ctx->ev_tiny[ctx->bank][1], ctx->tiny_stream[ctx->bank]);
  cudaStreamWaitEvent(ctx->stream, ctx->ev_tiny[ctx->bank][1], 0);
}

inline cudaEvent_t phase_event(eon_ctx* ctx) {
  if (ctx->ev_next == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[ctx->ev_next++];
}
inline void phase_reset(eon_ctx* ctx) {
  ctx->ev_next = 0;
  for (int p = 0; p < PH_COUNT; p++) ctx->ev_pairs[p].clear();
}
inline void phase_begin(eon_ctx* ctx, int ph) {
  if (ctx->ev_next > 4096) phase_reset(ctx);  // nobody is reading: do not grow without bound
  cudaEvent_t e = phase_event(ctx);
  cudaEventRecord(e, ctx->stream);
  ctx->ev_open[ph] = e;
}
inline void phase_end(eon_ctx* ctx, int ph) {
  cudaEvent_t e = phase_event(ctx);
  cudaEventRecord(e, ctx->stream);
  ctx->ev_pairs[ph].push_back(std::make_pair(ctx->ev_open[ph], e));
}

// host-side Fr helpers (same arithmetic as the device, software carry flag)
inline Fr fr_from_wire(const uint64_t w[4]) {
  Fr r;
  memcpy(r.v, w, 32);
  return r;
}
inline bool fr_wire_is_canonical(const uint64_t w[4]) {
  // lexicographic compare with the modulus, top limb first
  const u32* v = reinterpret_cast<const u32*>(w);
  for (int i = 7; i >= 0; i--) {
    u32 m = FrParams::mod(i);
    if (v[i] < m) return true;
    if (v[i] > m) return false;
  }
  return false;
}
// two_adic_generator(bits): omega_28 squared (28 - bits) times.  Reference: field.rs:567-573.
inline Fr fr_two_adic_generator(unsigned bits) {
  const u32 w28[8] = EON_FR_OMEGA28;
  Fr o;
  memcpy(o.v, w28, 32);
  for (unsigned i = bits; i < 28; i++) o = fp_sqr(o);
  return o;
}

// ---- internal cross-file API (device pointers, ctx->mu held by the caller) ----------------
enum Layout { LAYOUT_NATURAL = 0, LAYOUT_BITREV = 1 };

// forward coset NTT of size 2^log_n from 2^(log_n-k) coefficient rows (zero-padded).
// ld_src / ld_dst: row pitch in elements (0 = dense, i.e. width): a column group of a wider matrix.
int ntt_forward(eon_ctx* ctx, const Fr* d_src, Fr* d_dst, unsigned log_n, unsigned k, size_t width, const Fr& shift,
                Layout src_layout, size_t ld_src = 0, size_t ld_dst = 0);
// inverse coset NTT of size 2^log_n: evaluations on shift*H (natural) -> coefficients.
int ntt_inverse(eon_ctx* ctx, const Fr* d_src, Fr* d_dst, unsigned log_n, size_t width, const Fr& shift,
                Layout dst_layout, size_t ld_src = 0, size_t ld_dst = 0);

bool ntt_twiddles_are_fixed_operand();

// out[c] = sum_i scalars[i*ld + c] * bases[i], c < ncols.  d_out: ncols affine points (device).
int msm_run(eon_ctx* ctx, const G1Affine* d_bases, const Fr* d_scalars, size_t n, size_t ncols, size_t ld,
            G1Affine* d_out);
int g1_sum_run(eon_ctx* ctx, const G1Affine* d_points, size_t n, G1Affine* d_out);
// d_out[c] = sum_p d_parts[p * ncols + c]
int g1_sum_cols_run(eon_ctx* ctx, const G1Affine* d_parts, size_t nparts, size_t ncols, G1Affine* d_out);
int srs_build_range_tables(eon_ctx* ctx, size_t first, size_t n, unsigned window_bits);
int srs_generate(eon_ctx* ctx, const Fr& alpha, size_t n);
int srs_build_tables(eon_ctx* ctx, unsigned window_bits);
int srs_build_default_tables(eon_ctx* ctx);

int g1_compress_run(eon_ctx* ctx, const G1Affine* d_pts, size_t n, uint8_t* h_out, int enc);
int g1_decompress_run(eon_ctx* ctx, const uint8_t* h_in, size_t n, G1Affine* d_out, int enc, size_t* bad_index);

int quotient_run(eon_ctx* ctx, const Fr* d_coeffs, size_t h, size_t width, size_t ld_out, const Fr& z, Fr* d_quot,
                 Fr* d_values);

// d_out[i][c] = sum_j d_coeffs[i + j 2^log_n][c] * (shift^(2^log_n))^j, i < 2^log_n (2^log_n <= h)
int fold_coeffs_run(eon_ctx* ctx, const Fr* d_coeffs, size_t h, size_t width, unsigned log_n, const Fr& shift,
                    Fr* d_out);

// ---- host-buffer entry points shared with the multi-device context (multi.cu); every one takes ctx->mu ----------
enum DftKind { DFT_PLAIN = 0, DFT_COSET, DFT_INV, DFT_COSET_INV, DFT_COSET_LDE };
// transform of the columns [0, width) of a host matrix with row pitch ld_in (ld_out for the result), 0 = dense
int dft_host_locked(eon_ctx* ctx, int kind, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out,
                    unsigned log_h, size_t width, unsigned added_bits, const uint64_t shift[4]);
int dft_host(eon_ctx* ctx, int kind, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out, unsigned log_h,
             size_t width, unsigned added_bits, const uint64_t shift[4]);
// Pcs::commit_quotient of the chunk range [chunk_first, chunk_first + chunk_count) from the WHOLE host quotient matrix
int kzg_commit_quotient_range_host(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_size, size_t width,
                                   unsigned log_chunks, size_t chunk_first, size_t chunk_count, const uint64_t shift[4],
                                   uint64_t* h_commit_xy, eon_handle* out_handles);
// KzgMmcs::commit of the columns [0, width) of a host coefficient matrix with row pitch ld (0 = dense)
int kzg_commit_coeffs_host_ld(eon_ctx* ctx, const uint64_t* h_coeffs, size_t ld, size_t rows, size_t width,
                              uint64_t* h_commit_xy, eon_handle* out_handle);
// rows [first, first + n) of a host scalar matrix (row pitch ld, ncols used) against SRS[first, first + n): the
// partial sums stay on the device (*d_partial: ncols affine points, valid until the next MSM on this context);
// queued on ctx->stream, not synchronised
int msm_srs_range_host_partial(eon_ctx* ctx, const uint64_t* h_scalars_rows, size_t first, size_t n, size_t ncols,
                               size_t ld, const G1Affine** d_partial);
// columns [0, ncols) of a host scalar matrix with row pitch ld over SRS[0, n) -> ncols affine points on the host
int msm_srs_host_ld(eon_ctx* ctx, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy);

int bench_imad(eon_ctx* ctx, int kind, double* out_tops);
int bench_modmul(eon_ctx* ctx, int field, int variant, double* out_gmuls);

}  // namespace eon
