// extern "C" entry points of libeon_kzg (see include/eon_kzg.h for the contract and the
// reference file:line each function replaces).
#include <stdlib.h>

#include <algorithm>
#include <functional>

#include "common.cuh"
#include "msm.cuh"

using namespace eon;

namespace {

struct Lock {
  std::unique_lock<std::mutex> l;
  explicit Lock(eon_ctx* c) : l(c->mu) {}
};

int check_shift(eon_ctx* ctx, const uint64_t shift[4], Fr* out) {
  if (!shift) return fail(ctx, EON_ERR_BAD_ARG, "shift is null");
  if (!fr_wire_is_canonical(shift)) return fail(ctx, EON_ERR_BAD_ARG, "shift is not a canonical Fr (>= modulus)");
  *out = fr_from_wire(shift);
  // TwoAdicMultiplicativeCoset::new rejects a zero shift (field/src/coset.rs:77-86)
  if (out->is_zero()) return fail(ctx, EON_ERR_BAD_ARG, "shift must be non-zero");
  return EON_OK;
}

int set_device(eon_ctx* ctx) {
  EON_CUDA(ctx, cudaSetDevice(ctx->device));
  return EON_OK;
}

size_t mat_bytes(unsigned log_h, size_t width) { return (((size_t)1) << log_h) * width * sizeof(Fr); }

// host-buffer wrapper: H2D into scratch A, run f(d_in, d_out), D2H from scratch B.  The device copies are dense
// (row pitch = width); ld_in / ld_out are the host row pitches in elements (0 = dense): a column range of a wider
// host matrix goes through strided copies (the multi-device context hands every GPU its columns that way).
template <class F>
int host_io(eon_ctx* ctx, const uint64_t* h_in, size_t rows_in, size_t ld_in, uint64_t* h_out, size_t rows_out,
            size_t ld_out, size_t width, F f) {
  const size_t in_bytes = rows_in * width * sizeof(Fr), out_bytes = rows_out * width * sizeof(Fr);
  if ((in_bytes && !h_in) || (out_bytes && !h_out)) return fail(ctx, EON_ERR_BAD_ARG, "null host buffer");
  if ((ld_in && ld_in < width) || (ld_out && ld_out < width)) return fail(ctx, EON_ERR_BAD_ARG, "row pitch < width");
  void *d_in = nullptr, *d_out = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_A, in_bytes + 32, &d_in));
  EON_TRY(scratch_get(ctx, SC_IO_B, out_bytes + 32, &d_out));
  const size_t wb = width * sizeof(Fr);
  if (in_bytes) {
    if (ld_in == 0 || ld_in == width)
      EON_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    else
      EON_CUDA(ctx, cudaMemcpy2DAsync(d_in, wb, h_in, ld_in * sizeof(Fr), wb, rows_in, cudaMemcpyHostToDevice, ctx->stream));
  }
  EON_TRY(f((const Fr*)d_in, (Fr*)d_out));
  if (out_bytes) {
    if (ld_out == 0 || ld_out == width)
      EON_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    else
      EON_CUDA(ctx, cudaMemcpy2DAsync(h_out, ld_out * sizeof(Fr), d_out, wb, wb, rows_out, cudaMemcpyDeviceToHost, ctx->stream));
  }
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int check_dims(eon_ctx* ctx, unsigned log_h, size_t width, unsigned added = 0) {
  if (log_h + added > 28) return fail(ctx, EON_ERR_TWO_ADICITY, "transform size exceeds 2^28 (Fr::TWO_ADICITY)");
  if (width > 0xffffffffull) return fail(ctx, EON_ERR_BAD_ARG, "width too large");
  return EON_OK;
}

}  // namespace

extern "C" {

const char* eon_version(void) { return "eon_kzg 0.1 sm_100a"; }

int eon_ctx_create(int device, void* stream, eon_ctx** out) {
  if (!out) return EON_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return EON_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return EON_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return EON_ERR_CUDA;
  if (prop.major < 10) return EON_ERR_CUDA;  // sm_100a cubins only; no fallback path
  eon_ctx* ctx = new eon_ctx();
  ctx->device = device;
  ctx->stream = (cudaStream_t)stream;
  ctx->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("EON_L2_FETCH")) {  // experiment: L2 fetch granularity for the random base gathers
    size_t before = 0, after = 0;
    cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
    cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "eon: L2 fetch granularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(rc));
  }
  *out = ctx;
  return EON_OK;
}

void eon_ctx_destroy(eon_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& bank : ctx->scratch)
    for (auto& s : bank)
      if (s.ptr) cudaFree(s.ptr);
  if (ctx->split_stream) cudaStreamDestroy(ctx->split_stream);
  for (int b = 0; b < 2; b++) {
    if (ctx->tiny_stream[b]) cudaStreamDestroy(ctx->tiny_stream[b]);
    for (cudaEvent_t e : ctx->ev_tiny[b])
      if (e) cudaEventDestroy(e);
  }
  for (cudaEvent_t e : ctx->ev_split)
    if (e) cudaEventDestroy(e);
  for (auto& kv : ctx->twiddles) cudaFree(kv.second);
  for (auto& kv : ctx->handles) cudaFree(kv.second.d_coeffs);
  for (auto& kv : ctx->coeff_pool) cudaFree(kv.second);
  if (ctx->d_srs) cudaFree(ctx->d_srs);
  if (ctx->d_srs_tab) cudaFree(ctx->d_srs_tab);
  if (ctx->d_rng_tab) cudaFree(ctx->d_rng_tab);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_pipe)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->prio_stream) cudaStreamDestroy(ctx->prio_stream);
  delete ctx;
}

const char* eon_last_error(const eon_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int eon_ctx_sync(eon_ctx* ctx) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

uint64_t eon_ctx_launch_count(const eon_ctx* ctx) { return ctx ? ctx->launches : 0; }

int eon_dev_alloc(eon_ctx* ctx, size_t bytes, void** d_out) {
  if (!ctx || !d_out) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaMalloc(d_out, bytes ? bytes : 1));
  return EON_OK;
}
int eon_dev_free(eon_ctx* ctx, void* d_ptr) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  EON_CUDA(ctx, cudaFree(d_ptr));
  return EON_OK;
}
// page-locked host buffers for the matrices that cross PCIe (pageable memory halves the copy rate and a
// freshly allocated pageable output pays a page fault per 4 KiB on top)
int eon_host_alloc(size_t bytes, void** h_out) {
  if (!h_out) return EON_ERR_BAD_ARG;
  *h_out = nullptr;
  cudaError_t e = cudaHostAlloc(h_out, bytes ? bytes : 1, cudaHostAllocDefault);
  return e == cudaSuccess ? EON_OK : (e == cudaErrorMemoryAllocation ? EON_ERR_OOM : EON_ERR_CUDA);
}
int eon_host_free(void* h_ptr) { return cudaFreeHost(h_ptr) == cudaSuccess ? EON_OK : EON_ERR_CUDA; }

int eon_h2d(eon_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}
int eon_d2h(eon_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

// ---- DFT --------------------------------------------------------------------------------------
static int dft_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width, const Fr& shift) {
  if (!d_in || !d_out) return fail(ctx, EON_ERR_BAD_ARG, "null device buffer");
  if (d_in == d_out) return fail(ctx, EON_ERR_BAD_ARG, "in-place transform not supported: d_out aliases d_in");
  EON_TRY(check_dims(ctx, log_h, width));
  return ntt_forward(ctx, (const Fr*)d_in, (Fr*)d_out, log_h, 0, width, shift, LAYOUT_NATURAL);
}
static int idft_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width, const Fr& shift) {
  if (!d_in || !d_out) return fail(ctx, EON_ERR_BAD_ARG, "null device buffer");
  if (d_in == d_out) return fail(ctx, EON_ERR_BAD_ARG, "in-place transform not supported: d_out aliases d_in");
  EON_TRY(check_dims(ctx, log_h, width));
  return ntt_inverse(ctx, (const Fr*)d_in, (Fr*)d_out, log_h, width, shift, LAYOUT_NATURAL);
}
static int lde_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width, unsigned added,
                   const Fr& shift) {
  if (!d_in || !d_out) return fail(ctx, EON_ERR_BAD_ARG, "null device buffer");
  if (d_in == d_out) return fail(ctx, EON_ERR_BAD_ARG, "in-place transform not supported: d_out aliases d_in");
  EON_TRY(check_dims(ctx, log_h, width, added));
  if (width == 0) return EON_OK;
  // idft (no shift) -> coefficients in bit-reversed order -> zero-pad (replicate) + coset DFT:
  // no permutation pass in between (cf. Radix2DFTSmallBatch, dft/src/radix_2_small_batch.rs:246-351)
  void* tmp = nullptr;
  EON_TRY(scratch_get(ctx, SC_NTT_TMP, mat_bytes(log_h, width), &tmp));
  EON_TRY(ntt_inverse(ctx, (const Fr*)d_in, (Fr*)tmp, log_h, width, Fr::one(), LAYOUT_BITREV));
  return ntt_forward(ctx, (const Fr*)tmp, (Fr*)d_out, log_h + added, added, width, shift, LAYOUT_BITREV);
}

int eon_dft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return dft_dev(ctx, d_in, d_out, log_h, width, Fr::one());
}
int eon_coset_dft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                            const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  return dft_dev(ctx, d_in, d_out, log_h, width, s);
}
int eon_idft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return idft_dev(ctx, d_in, d_out, log_h, width, Fr::one());
}
int eon_coset_idft_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                             const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  return idft_dev(ctx, d_in, d_out, log_h, width, s);
}
int eon_coset_lde_batch_dev(eon_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, unsigned log_h, size_t width,
                            unsigned added_bits, const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  return lde_dev(ctx, d_in, d_out, log_h, width, added_bits, s);
}

}  // extern "C"

namespace eon {
// Host-buffer transform of one matrix (or of a column range of a wider host matrix: ld_in / ld_out = host row
// pitch in elements, 0 = dense), ctx->mu held.  kind: DFT_* below; shift ignored for the plain kinds.
int dft_host_locked(eon_ctx* ctx, int kind, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out,
                    unsigned log_h, size_t width, unsigned added_bits, const uint64_t shift[4]) {
  EON_TRY(set_device(ctx));
  if (kind != DFT_COSET_LDE) added_bits = 0;
  EON_TRY(check_dims(ctx, log_h, width, added_bits));
  Fr s = Fr::one();
  if (kind == DFT_COSET || kind == DFT_COSET_INV || kind == DFT_COSET_LDE) EON_TRY(check_shift(ctx, shift, &s));
  const size_t rows = (size_t)1 << log_h;
  return host_io(ctx, h_in, rows, ld_in, h_out, rows << added_bits, ld_out, width, [&](const Fr* i, Fr* o) {
    const uint64_t* di = (const uint64_t*)i;
    uint64_t* dout = (uint64_t*)o;
    switch (kind) {
      case DFT_PLAIN:
      case DFT_COSET: return dft_dev(ctx, di, dout, log_h, width, s);
      case DFT_INV:
      case DFT_COSET_INV: return idft_dev(ctx, di, dout, log_h, width, s);
      case DFT_COSET_LDE: return lde_dev(ctx, di, dout, log_h, width, added_bits, s);
    }
    return fail(ctx, EON_ERR_BAD_ARG, "unknown transform kind");
  });
}
}  // namespace eon

extern "C" {

int eon_dft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_PLAIN, h_in, 0, h_out, 0, log_h, width, 0, nullptr);
}
int eon_coset_dft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                        const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_COSET, h_in, 0, h_out, 0, log_h, width, 0, shift);
}
int eon_idft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_INV, h_in, 0, h_out, 0, log_h, width, 0, nullptr);
}
int eon_coset_idft_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                         const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_COSET_INV, h_in, 0, h_out, 0, log_h, width, 0, shift);
}
int eon_coset_lde_batch(eon_ctx* ctx, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                        unsigned added_bits, const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_COSET_LDE, h_in, 0, h_out, 0, log_h, width, added_bits, shift);
}
int eon_coset_lde_batch_ld(eon_ctx* ctx, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out,
                           unsigned log_h, size_t width, unsigned added_bits, const uint64_t shift[4]) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  return dft_host_locked(ctx, DFT_COSET_LDE, h_in, ld_in, h_out, ld_out, log_h, width, added_bits, shift);
}

// ---- SRS ---------------------------------------------------------------------------------------
static int srs_drop(eon_ctx* ctx) {
  if (ctx->d_srs) {
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EON_CUDA(ctx, cudaFree(ctx->d_srs));
    ctx->d_srs = nullptr;
    ctx->srs_n = 0;
    EON_TRY(srs_build_tables(ctx, 0));
    EON_TRY(srs_build_range_tables(ctx, 0, 0, 0));
  }
  return EON_OK;
}

int eon_srs_load_affine(eon_ctx* ctx, const uint64_t* h_xy, size_t n) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n && !h_xy) return fail(ctx, EON_ERR_BAD_ARG, "null SRS pointer");
  EON_TRY(srs_drop(ctx));
  if (n == 0) return EON_OK;
  EON_CUDA(ctx, cudaMalloc(&ctx->d_srs, n * sizeof(G1Affine)));
  EON_CUDA(ctx, cudaMemcpyAsync(ctx->d_srs, h_xy, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
  ctx->srs_n = n;
  EON_TRY(srs_build_default_tables(ctx));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_srs_load_compressed(eon_ctx* ctx, const uint8_t* h_in, size_t n, int enc, size_t* bad_index) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n && !h_in) return fail(ctx, EON_ERR_BAD_ARG, "null SRS pointer");
  if (bad_index) *bad_index = (size_t)-1;
  EON_TRY(srs_drop(ctx));
  if (n == 0) return EON_OK;
  EON_CUDA(ctx, cudaMalloc(&ctx->d_srs, n * sizeof(G1Affine)));
  int rc = g1_decompress_run(ctx, h_in, n, ctx->d_srs, enc, bad_index);
  if (rc != EON_OK) {
    cudaFree(ctx->d_srs);
    ctx->d_srs = nullptr;
    return rc;
  }
  ctx->srs_n = n;
  EON_TRY(srs_build_default_tables(ctx));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_g1_compress(eon_ctx* ctx, const uint64_t* h_xy, size_t n, uint8_t* h_out, int enc) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n && (!h_xy || !h_out)) return fail(ctx, EON_ERR_BAD_ARG, "null pointer");
  if (n == 0) return EON_OK;
  void* d_pts;
  EON_TRY(scratch_get(ctx, SC_IO_A, n * sizeof(G1Affine), &d_pts));
  EON_CUDA(ctx, cudaMemcpyAsync(d_pts, h_xy, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
  return g1_compress_run(ctx, (const G1Affine*)d_pts, n, h_out, enc);
}

int eon_g1_decompress(eon_ctx* ctx, const uint8_t* h_in, size_t n, uint64_t* h_xy, int enc, size_t* bad_index) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n && (!h_in || !h_xy)) return fail(ctx, EON_ERR_BAD_ARG, "null pointer");
  if (bad_index) *bad_index = (size_t)-1;
  if (n == 0) return EON_OK;
  void* d_pts;
  EON_TRY(scratch_get(ctx, SC_IO_A, n * sizeof(G1Affine), &d_pts));
  EON_TRY(g1_decompress_run(ctx, h_in, n, (G1Affine*)d_pts, enc, bad_index));
  EON_CUDA(ctx, cudaMemcpyAsync(h_xy, d_pts, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_srs_generate_unsafe(eon_ctx* ctx, const uint64_t alpha[4], size_t n) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (!alpha || !fr_wire_is_canonical(alpha)) return fail(ctx, EON_ERR_BAD_ARG, "alpha is not a canonical Fr");
  EON_TRY(srs_generate(ctx, fr_from_wire(alpha), n));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

size_t eon_srs_size(const eon_ctx* ctx) { return ctx ? ctx->srs_n : 0; }

int eon_srs_set_window_tables(eon_ctx* ctx, unsigned window_bits) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_TRY(srs_build_tables(ctx, window_bits));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}
unsigned eon_srs_window_bits(const eon_ctx* ctx) { return ctx ? ctx->srs_tab_c : 0; }

int eon_msm_set_sort_mode(eon_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 2) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  ctx->msm_sort_mode = mode;
  return EON_OK;
}

int eon_msm_set_split(eon_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  ctx->msm_split_mode = mode;
  return EON_OK;
}

int eon_msm_set_slice_schedule(eon_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  ctx->msm_slice_mode = mode;
  return EON_OK;
}

unsigned eon_msm_rounds_used(const eon_ctx* ctx) { return ctx ? ctx->msm_rounds_used : 0; }
unsigned eon_msm_window_bits_used(const eon_ctx* ctx) { return ctx ? ctx->msm_c_used : 0; }

int eon_msm_set_rounds(eon_ctx* ctx, int rounds) {
  if (!ctx || rounds < -1 || rounds > 6) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  ctx->msm_rounds = rounds;
  return EON_OK;
}

int eon_srs_read(eon_ctx* ctx, size_t first, size_t n, uint64_t* h_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (first > ctx->srs_n || n > ctx->srs_n - first) return fail(ctx, EON_ERR_BAD_ARG, "SRS range out of bounds");
  if (n == 0) return EON_OK;
  EON_CUDA(ctx, cudaMemcpyAsync(h_xy, ctx->d_srs + first, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

// ---- MSM ---------------------------------------------------------------------------------------
static int msm_to_host(eon_ctx* ctx, const G1Affine* d_bases, const Fr* d_scalars, size_t n, size_t ncols, size_t ld,
                       uint64_t* h_out_xy) {
  if (ncols == 0) return EON_OK;
  if (!h_out_xy) return fail(ctx, EON_ERR_BAD_ARG, "null output pointer");
  if (ld < ncols) return fail(ctx, EON_ERR_BAD_ARG, "ld < ncols");
  void* d_res = nullptr;
  EON_TRY(scratch_get(ctx, SC_MSM_RESULT, ncols * sizeof(G1Affine), &d_res));
  EON_TRY(msm_run(ctx, d_bases, d_scalars, n, ncols, ld, (G1Affine*)d_res));
  EON_CUDA(ctx, cudaMemcpyAsync(h_out_xy, d_res, ncols * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_msm_srs_range_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t first, size_t n, size_t ncols, size_t ld,
                          uint64_t* h_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (first > ctx->srs_n || n > ctx->srs_n - first) {
    // commit_column's degree guard (kzg/src/util.rs:38, params.rs:164-173)
    char b[128];
    snprintf(b, sizeof(b), "DegreeTooLarge: need %zu SRS points, have %zu", first + n, ctx->srs_n);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, b);
  }
  if (n && ncols && !d_scalars) return fail(ctx, EON_ERR_BAD_ARG, "null scalars");
  return msm_to_host(ctx, ctx->d_srs + first, (const Fr*)d_scalars, n, ncols, ld, h_out_xy);
}

// index-range shard with the partial sums left on the device (ncols affine points at d_out_xy), queued on the
// context's stream without a host synchronisation: the caller gathers the shards' partial sums over NVLink
// (peer copy or ncclAllGather) and adds them with eon_g1_sum_cols_dev -- no host hop in between.
int eon_msm_srs_range_partial_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t first, size_t n, size_t ncols,
                                  size_t ld, uint64_t* d_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (first > ctx->srs_n || n > ctx->srs_n - first) {
    char b[128];
    snprintf(b, sizeof(b), "DegreeTooLarge: need %zu SRS points, have %zu", first + n, ctx->srs_n);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, b);
  }
  if (ncols == 0) return EON_OK;
  if ((n && !d_scalars) || !d_out_xy) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  if (ld < ncols) return fail(ctx, EON_ERR_BAD_ARG, "ld < ncols");
  return msm_run(ctx, ctx->d_srs + first, (const Fr*)d_scalars, n, ncols, ld, (G1Affine*)d_out_xy);
}

int eon_g1_sum_cols_dev(eon_ctx* ctx, const uint64_t* d_parts_xy, size_t nparts, size_t ncols, uint64_t* h_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (ncols == 0) return EON_OK;
  if ((nparts && !d_parts_xy) || !h_out_xy) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  void* d_res = nullptr;
  EON_TRY(scratch_get(ctx, SC_MSM_RESULT, ncols * sizeof(G1Affine) + 64, &d_res));
  EON_TRY(g1_sum_cols_run(ctx, (const G1Affine*)d_parts_xy, nparts, ncols, (G1Affine*)d_res));
  EON_CUDA(ctx, cudaMemcpyAsync(h_out_xy, d_res, ncols * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_srs_set_range_tables(eon_ctx* ctx, size_t first, size_t n, unsigned window_bits) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_TRY(srs_build_range_tables(ctx, first, n, window_bits));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_msm_srs_dev(eon_ctx* ctx, const uint64_t* d_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy) {
  return eon_msm_srs_range_dev(ctx, d_scalars, 0, n, ncols, ld, h_out_xy);
}

int eon_msm_srs(eon_ctx* ctx, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n > ctx->srs_n) return fail(ctx, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: polynomial longer than the SRS");
  if (n && ncols && !h_scalars) return fail(ctx, EON_ERR_BAD_ARG, "null scalars");
  if (ld < ncols) return fail(ctx, EON_ERR_BAD_ARG, "ld < ncols");
  void* d_sc = nullptr;
  size_t bytes = n * ld * sizeof(Fr);
  EON_TRY(scratch_get(ctx, SC_IO_A, bytes + 32, &d_sc));
  if (bytes) EON_CUDA(ctx, cudaMemcpyAsync(d_sc, h_scalars, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return msm_to_host(ctx, ctx->d_srs, (const Fr*)d_sc, n, ncols, ld, h_out_xy);
}

int eon_msm_points(eon_ctx* ctx, const uint64_t* h_points_xy, const uint64_t* h_scalars, size_t n,
                   uint64_t* h_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n && (!h_points_xy || !h_scalars)) return fail(ctx, EON_ERR_BAD_ARG, "null input");
  void *d_sc = nullptr, *d_pt = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_A, n * sizeof(Fr) + 32, &d_sc));
  EON_TRY(scratch_get(ctx, SC_IO_B, n * sizeof(G1Affine) + 64, &d_pt));
  if (n) {
    EON_CUDA(ctx, cudaMemcpyAsync(d_sc, h_scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    EON_CUDA(ctx, cudaMemcpyAsync(d_pt, h_points_xy, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
  }
  return msm_to_host(ctx, (const G1Affine*)d_pt, (const Fr*)d_sc, n, 1, 1, h_out_xy);
}

int eon_g1_sum(eon_ctx* ctx, const uint64_t* h_points_xy, size_t n, uint64_t* h_out_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if ((n && !h_points_xy) || !h_out_xy) return fail(ctx, EON_ERR_BAD_ARG, "null input");
  void* d_pt = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_B, (n + 1) * sizeof(G1Affine), &d_pt));
  G1Affine* pts = (G1Affine*)d_pt;
  if (n) EON_CUDA(ctx, cudaMemcpyAsync(pts + 1, h_points_xy, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
  EON_TRY(g1_sum_run(ctx, pts + 1, n, pts));
  EON_CUDA(ctx, cudaMemcpyAsync(h_out_xy, pts, sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

// ---- KZG ---------------------------------------------------------------------------------------
// coefficient buffers come from a small pool of freed ones (see eon_handle_free)
static int coeff_buffer_get(eon_ctx* ctx, size_t need, Fr** out, size_t* out_cap) {
  for (size_t i = 0; i < ctx->coeff_pool.size(); i++) {
    if (ctx->coeff_pool[i].first >= need && ctx->coeff_pool[i].first <= 2 * need) {
      *out_cap = ctx->coeff_pool[i].first;
      *out = (Fr*)ctx->coeff_pool[i].second;
      ctx->coeff_pool.erase(ctx->coeff_pool.begin() + i);
      return EON_OK;
    }
  }
  EON_CUDA(ctx, cudaMalloc(out, need));
  *out_cap = need;
  return EON_OK;
}

static eon_handle handle_new(eon_ctx* ctx, Fr* d_coeffs, size_t rows, unsigned log_h, size_t width, size_t cap) {
  ProverMatrix pm;
  pm.d_coeffs = d_coeffs;
  pm.rows = rows;
  pm.log_h = log_h;
  pm.width = width;
  pm.cap = cap;
  eon_handle id = ctx->next_handle++;
  ctx->handles[id] = pm;
  return id;
}

static int pipe_init(eon_ctx* ctx);

// Coset NTT of a (column group of a) coefficient matrix on the auxiliary compute stream: queued behind
// what the main stream has done so far (the iDFT that produced the coefficients), it then runs BESIDE
// whatever the main stream does next — the MSM, whose histogram / sort / base-gather phases leave the
// integer pipe mostly idle.  ev_pipe[ev_slot] is recorded on the auxiliary stream when it is done.
static int lde_on_aux(eon_ctx* ctx, const Fr* d_coeffs, Fr* d_lde, unsigned lde_log_size, unsigned added, size_t gw,
                      const Fr& ls, size_t ld, int ev_slot) {
  EON_TRY(pipe_init(ctx));
  EON_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[ev_slot], ctx->stream));
  EON_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_pipe[ev_slot], 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->aux_stream;
  int rc = ntt_forward(ctx, d_coeffs, d_lde, lde_log_size, added, gw, ls, LAYOUT_NATURAL, ld, ld);
  ctx->stream = main_stream;
  EON_TRY(rc);
  EON_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[ev_slot], ctx->aux_stream));
  return EON_OK;
}

// While an LDE transform shares the GPU with an MSM, the MSM runs on a HIGH-PRIORITY stream: its CTAs are scheduled
// first, so the transform (normal priority, auxiliary stream) fills exactly the cycles the MSM leaves idle -- its
// single-warp inversion trees and bucket reduction, its sort passes -- instead of delaying the MSM's wide kernels.
// (A stream cannot be made lower than the caller's default-priority stream, hence the MSM moves up.)
// EON_MSM_PRIO=0: everything on the caller's stream as before.
static bool msm_prio_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EON_MSM_PRIO");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
// f() issues MSM work on ctx->stream; here that is the high-priority stream, ordered after what the caller's
// stream has queued so far; ev_pipe[ev_slot] is recorded when the MSM work is done (the caller's stream is NOT made
// to wait: see msm_prio_join).
static int msm_on_prio(eon_ctx* ctx, int ev_slot, const std::function<int()>& f) {
  EON_TRY(pipe_init(ctx));
  EON_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[ev_slot], ctx->stream));
  EON_CUDA(ctx, cudaStreamWaitEvent(ctx->prio_stream, ctx->ev_pipe[ev_slot], 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->prio_stream;
  int rc = f();
  ctx->stream = main_stream;
  EON_TRY(rc);
  EON_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[ev_slot], ctx->prio_stream));
  return EON_OK;
}
static int msm_prio_join(eon_ctx* ctx, int ev_slot) {
  EON_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[ev_slot], 0));
  return EON_OK;
}

static int kzg_commit_locked(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_h, size_t width,
                             const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle,
                             unsigned lde_log_size = 0, const uint64_t* lde_shift = nullptr, Fr* d_lde = nullptr) {
  if (!out_handle) return fail(ctx, EON_ERR_BAD_ARG, "null handle pointer");
  *out_handle = 0;
  EON_TRY(check_dims(ctx, log_h, width));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  const bool want_lde = lde_log_size != 0 && width != 0;
  Fr ls = Fr::one();
  if (want_lde) {
    if (lde_log_size < log_h) return fail(ctx, EON_ERR_BAD_ARG, "LDE domain smaller than the trace domain");
    EON_TRY(check_dims(ctx, lde_log_size, width));
    EON_TRY(check_shift(ctx, lde_shift, &ls));
    if (!d_lde) return fail(ctx, EON_ERR_BAD_ARG, "null LDE output");
  }
  const size_t h = (size_t)1 << log_h;
  if (h > ctx->srs_n) {  // ensure_supported(height - 1), kzg/src/pcs.rs:238-240
    char b[128];
    snprintf(b, sizeof(b), "DegreeTooLarge: degree %zu > max %zu", h - 1, ctx->srs_n ? ctx->srs_n - 1 : 0);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, b);
  }
  if (width && (!d_evals || !h_commit_xy)) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  Fr* d_coeffs = nullptr;
  size_t cap = 0;
  EON_TRY(coeff_buffer_get(ctx, mat_bytes(log_h, width) + 32, &d_coeffs, &cap));
  int rc = ntt_inverse(ctx, (const Fr*)d_evals, d_coeffs, log_h, width, s, LAYOUT_NATURAL);
  // The LDE transform is queued BEHIND round 0 of the MSM (hook in msm_tree_rounds): the sort passes leave no
  // shared memory for a transform CTA beside them, round 0 lives on its table slice staying in the L2 (a transform
  // streaming 1.5 GB through it at the same time cost 3-5 ms when the two happened to meet), and what the transform
  // can really fill are the single-warp phases after it -- the inversion trees of rounds 1-2, the bucket reduction.
  // EON_LDE_AFTER_R0=0: queued before the MSM, as in round 1.
  static const int lde_after_r0 = getenv("EON_LDE_AFTER_R0") ? atoi(getenv("EON_LDE_AFTER_R0")) : 1;
  const bool hook = want_lde && msm_prio_enabled() && lde_after_r0;
  auto queue_lde = [&] { return lde_on_aux(ctx, d_coeffs, d_lde, lde_log_size, lde_log_size - log_h, width, ls, width, 19); };
  if (rc == EON_OK && want_lde && !hook) rc = queue_lde();
  if (rc == EON_OK) {
    if (want_lde && msm_prio_enabled()) {
      if (hook) ctx->after_round0 = queue_lde;
      rc = msm_on_prio(ctx, 18, [&] { return msm_to_host(ctx, ctx->d_srs, d_coeffs, h, width, width, h_commit_xy); });
      if (ctx->after_round0) {  // the MSM had no pairwise rounds (small input) or failed before them
        ctx->after_round0 = nullptr;
        if (rc == EON_OK) rc = queue_lde();
      }
      if (rc == EON_OK) rc = msm_prio_join(ctx, 18);
    } else {
      rc = msm_to_host(ctx, ctx->d_srs, d_coeffs, h, width, width, h_commit_xy);
    }
  }
  if (rc == EON_OK && want_lde) {  // the call returns with the LDE complete as well, and later work on the caller's
    // stream is ordered after it
    if (cudaStreamSynchronize(ctx->aux_stream) != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, "LDE stream failed");
  }
  if (rc != EON_OK) {
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->prio_stream) cudaStreamSynchronize(ctx->prio_stream);
    cudaFree(d_coeffs);
    return rc;
  }
  *out_handle = handle_new(ctx, d_coeffs, h, log_h, width, cap);
  return EON_OK;
}

// KzgMmcs::commit (kzg/src/mmcs.rs:155-190): the matrix columns ARE the coefficient vectors (no
// iDFT), any height.  d_coeffs_in is copied into a pooled buffer that the handle owns.
static int kzg_commit_coeffs_locked(eon_ctx* ctx, const uint64_t* src, bool src_is_host, size_t rows, size_t width,
                                    uint64_t* h_commit_xy, eon_handle* out_handle, size_t src_ld = 0) {
  if (!out_handle) return fail(ctx, EON_ERR_BAD_ARG, "null handle pointer");
  *out_handle = 0;
  if (width > 0xffffffffull || (width && rows > (~(size_t)0) / (width * sizeof(Fr))))
    return fail(ctx, EON_ERR_BAD_ARG, "matrix too large");
  const size_t degree = rows ? rows - 1 : 0;  // ensure_supported(height.saturating_sub(1)), mmcs.rs:177-179
  if (ctx->srs_n == 0 || degree > ctx->srs_n - 1) {
    char b[128];
    snprintf(b, sizeof(b), "DegreeTooLarge: degree %zu > max %zu", degree, ctx->srs_n ? ctx->srs_n - 1 : 0);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, b);
  }
  const size_t bytes = rows * width * sizeof(Fr);
  if (bytes && (!src || !h_commit_xy)) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  Fr* d_coeffs = nullptr;
  size_t cap = 0;
  EON_TRY(coeff_buffer_get(ctx, bytes + 32, &d_coeffs, &cap));
  int rc = EON_OK;
  if (src_ld && src_ld < width) {
    cudaFree(d_coeffs);
    return fail(ctx, EON_ERR_BAD_ARG, "row pitch < width");
  }
  if (bytes) {
    const cudaMemcpyKind kind = src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaError_t e = (src_ld == 0 || src_ld == width)
                        ? cudaMemcpyAsync(d_coeffs, src, bytes, kind, ctx->stream)
                        : cudaMemcpy2DAsync(d_coeffs, width * sizeof(Fr), src, src_ld * sizeof(Fr), width * sizeof(Fr), rows,
                                            kind, ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("copy of the coefficient matrix failed: ") + cudaGetErrorString(e));
  }
  if (rc == EON_OK) rc = msm_to_host(ctx, ctx->d_srs, d_coeffs, rows, width, width, h_commit_xy);
  if (rc != EON_OK) {
    cudaFree(d_coeffs);
    return rc;
  }
  unsigned log_h = NOT_POW2;
  if (rows && (rows & (rows - 1)) == 0) {
    log_h = 0;
    while (((size_t)1 << log_h) < rows) log_h++;
  }
  *out_handle = handle_new(ctx, d_coeffs, rows, log_h, width, cap);
  return EON_OK;
}

int eon_kzg_commit_coeffs(eon_ctx* ctx, const uint64_t* h_coeffs, size_t rows, size_t width, uint64_t* h_commit_xy,
                          eon_handle* out_handle) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_coeffs_locked(ctx, h_coeffs, true, rows, width, h_commit_xy, out_handle);
}

int eon_kzg_commit_coeffs_dev(eon_ctx* ctx, const uint64_t* d_coeffs, size_t rows, size_t width,
                              uint64_t* h_commit_xy, eon_handle* out_handle) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_coeffs_locked(ctx, d_coeffs, false, rows, width, h_commit_xy, out_handle);
}

int eon_kzg_commit_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                       uint64_t* h_commit_xy, eon_handle* out_handle) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_locked(ctx, d_evals, log_h, width, shift, h_commit_xy, out_handle);
}

int eon_kzg_commit_lde_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                           uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                           const uint64_t lde_shift[4], uint64_t* d_lde_out) {
  if (!ctx || lde_log_size == 0) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_locked(ctx, d_evals, log_h, width, shift, h_commit_xy, out_handle, lde_log_size, lde_shift,
                           (Fr*)d_lde_out);
}

// Column groups for the PCIe pipelining of the host-buffer entry points.  Every column's transform
// and MSM is independent (kzg/src/pcs.rs:244-249), so group g+1 can cross PCIe while group g computes.
// Strided copies of >= 128-byte row pieces run at >= 85 % of the contiguous rate (tools/copy2d_probe.py).
static std::vector<std::pair<size_t, size_t>> column_groups(size_t width, size_t bytes, bool small_first) {
  std::vector<std::pair<size_t, size_t>> g;
  if (width < 8 || bytes < ((size_t)32 << 20) || getenv("EON_NO_PIPELINE")) {
    g.push_back(std::make_pair((size_t)0, width));
    return g;
  }
  // upload: a small first group (its copy is the only exposed one); download: two halves (256-byte rows)
  size_t g0 = small_first ? std::max<size_t>(4, (width / 4) & ~(size_t)3) : ((width / 2 + 3) & ~(size_t)3);
  if (g0 >= width) g0 = width / 2;
  g.push_back(std::make_pair((size_t)0, g0));
  g.push_back(std::make_pair(g0, width - g0));
  return g;
}

// Column groups of the commit upload: groups of 4 columns (128-byte row pieces, >= 85 % of the contiguous PCIe rate),
// at most 8 groups.  The transforms of a group (coset iDFT, and the hinted LDE on the auxiliary stream) run while
// the next groups cross PCIe; ONE MSM over all columns follows.  Measured at 2^20 x 16 against the former scheme
// (a small first group, then an MSM per group so that the MSM of group g hides the upload of group g + 1): the
// upload (10 ms) now hides under the 9 ms of transforms instead, the MSM runs once at its 16-column efficiency and
// without the LDE beside it.  EON_PIPE_MSM_PER_GROUP=1 restores the former scheme.
static std::vector<std::pair<size_t, size_t>> upload_groups(size_t width, size_t bytes) {
  std::vector<std::pair<size_t, size_t>> g;
  if (width < 8 || bytes < ((size_t)32 << 20) || getenv("EON_NO_PIPELINE")) {
    g.push_back(std::make_pair((size_t)0, width));
    return g;
  }
  size_t gw = 4;
  while ((width + gw - 1) / gw > 8) gw += 4;
  for (size_t c0 = 0; c0 < width; c0 += gw) g.push_back(std::make_pair(c0, std::min(gw, width - c0)));
  return g;
}

static int pipe_init(eon_ctx* ctx) {
  if (ctx->copy_stream && ctx->copy_stream2 && ctx->aux_stream && ctx->prio_stream && ctx->ev_pipe[19]) return EON_OK;
  if (!ctx->copy_stream) EON_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if (!ctx->copy_stream2) EON_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
  if (!ctx->aux_stream) EON_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
  if (!ctx->prio_stream) {
    int least = 0, greatest = 0;
    EON_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    // one level below the top: the top is kept for the single-warp phases of a split MSM (msm_run)
    EON_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->prio_stream, cudaStreamNonBlocking, greatest < least ? greatest + 1 : greatest));
  }
  for (auto& e : ctx->ev_pipe)
    if (!e) EON_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return EON_OK;
}

// commit from host buffers, optionally with the evaluations on a second coset (the LDE the prover asks
// for next, eon-uni-stark/src/prover.rs:307-322) produced in the same call: lde_log_size == 0 -> none.
// h_ld_in / h_ld_out: row pitch of the host matrices in elements (0 = dense, i.e. width): the columns
// [c0, c0 + width) of a wider host matrix are passed as (base + c0, ld = full width).
static int kzg_commit_host(eon_ctx* ctx, const uint64_t* h_evals, size_t h_ld_in, unsigned log_h, size_t width,
                           const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                           const uint64_t lde_shift[4], uint64_t* h_lde_out, size_t h_ld_out) {
  EON_TRY(check_dims(ctx, log_h, width));
  if (h_ld_in == 0) h_ld_in = width;
  if (h_ld_out == 0) h_ld_out = width;
  if (h_ld_in < width || h_ld_out < width) return fail(ctx, EON_ERR_BAD_ARG, "row pitch < width");
  const size_t hpitch_in = h_ld_in * sizeof(Fr), hpitch_out = h_ld_out * sizeof(Fr);
  const bool want_lde = lde_log_size != 0;
  Fr ls = Fr::one();
  if (want_lde) {
    if (lde_log_size < log_h) return fail(ctx, EON_ERR_BAD_ARG, "LDE domain smaller than the trace domain");
    EON_TRY(check_dims(ctx, lde_log_size, width));
    EON_TRY(check_shift(ctx, lde_shift, &ls));
    if (width && !h_lde_out) return fail(ctx, EON_ERR_BAD_ARG, "null LDE output");
  }
  size_t b = mat_bytes(log_h, width);
  if (b && !h_evals) return fail(ctx, EON_ERR_BAD_ARG, "null evals");
  void* d_in = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_A, b + 32, &d_in));
  static const int per_group_env = getenv("EON_PIPE_MSM_PER_GROUP") ? atoi(getenv("EON_PIPE_MSM_PER_GROUP")) : 0;
  const bool msm_per_group = per_group_env != 0;
  auto groups = msm_per_group ? column_groups(width, b, true) : upload_groups(width, b);
  if (groups.size() == 1 && !want_lde) {
    if (b && h_ld_in == width) EON_CUDA(ctx, cudaMemcpyAsync(d_in, h_evals, b, cudaMemcpyHostToDevice, ctx->stream));
    else if (b)
      EON_CUDA(ctx, cudaMemcpy2DAsync(d_in, width * sizeof(Fr), h_evals, hpitch_in, width * sizeof(Fr), (size_t)1 << log_h,
                                      cudaMemcpyHostToDevice, ctx->stream));
    return kzg_commit_locked(ctx, (const uint64_t*)d_in, log_h, width, shift, h_commit_xy, out_handle);
  }
  if (width == 0) {
    if (!out_handle) return fail(ctx, EON_ERR_BAD_ARG, "null handle pointer");
    return kzg_commit_locked(ctx, (const uint64_t*)d_in, log_h, width, shift, h_commit_xy, out_handle);
  }
  // pipelined: H2D of group g+1 (copy stream) under the iDFT + MSM of group g (compute stream)
  if (!out_handle) return fail(ctx, EON_ERR_BAD_ARG, "null handle pointer");
  *out_handle = 0;
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  const size_t h = (size_t)1 << log_h;
  if (h > ctx->srs_n) {  // ensure_supported(height - 1), kzg/src/pcs.rs:238-240 — before any copy
    char msg[128];
    snprintf(msg, sizeof(msg), "DegreeTooLarge: degree %zu > max %zu", h - 1, ctx->srs_n ? ctx->srs_n - 1 : 0);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, msg);
  }
  if (!h_commit_xy) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  EON_TRY(pipe_init(ctx));
  Fr* d_coeffs = nullptr;
  size_t cap = 0;
  EON_TRY(coeff_buffer_get(ctx, b + 32, &d_coeffs, &cap));
  void* d_commit = nullptr;
  int rc = scratch_get(ctx, SC_MSM_RESULT, width * sizeof(G1Affine) + 64, &d_commit);
  void* d_lde = nullptr;
  const size_t lde_rows = want_lde ? (size_t)1 << lde_log_size : 0;
  if (rc == EON_OK && want_lde) rc = scratch_get(ctx, SC_IO_B, mat_bytes(lde_log_size, width) + 32, &d_lde);
  const size_t pitch = width * sizeof(Fr);
  // earlier work on the compute stream may still read d_in (same scratch): copies wait for it
  if (rc == EON_OK && cudaEventRecord(ctx->ev_pipe[17], ctx->stream) != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, "event");
  if (rc == EON_OK) cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[17], 0);
  for (size_t g = 0; rc == EON_OK && g < groups.size(); g++) {
    const size_t c0 = groups[g].first, gw = groups[g].second;
    cudaError_t e = cudaMemcpy2DAsync((Fr*)d_in + c0, pitch, (const Fr*)h_evals + c0, hpitch_in, gw * sizeof(Fr), h,
                                      cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_pipe[g], ctx->copy_stream);
    if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("group upload failed: ") + cudaGetErrorString(e));
  }
  // The MSM's digit / sort passes of a group are queued right behind its iDFT: they run while the next groups are
  // still crossing PCIe (2.6 ms per group of 4 columns at 2^20 rows against 0.7 ms of iDFT), and only the pairwise
  // rounds onwards -- which need every column in the same launch -- wait for the last group.  EON_PIPE_EARLY_SORT=0:
  // the whole MSM after the last group.
  static const int early_sort_env = getenv("EON_PIPE_EARLY_SORT") ? atoi(getenv("EON_PIPE_EARLY_SORT")) : 1;
  MsmBatch msm_batch_state;
  bool early_sort = false;
  if (rc == EON_OK && !msm_per_group && early_sort_env && groups.size() > 1) {
    const int brc = msm_stream_begin(ctx, ctx->d_srs, h, width, &msm_batch_state);
    if (brc < 0) rc = brc;
    early_sort = brc == EON_OK;
  }
  for (size_t g = 0; rc == EON_OK && g < groups.size(); g++) {
    const size_t c0 = groups[g].first, gw = groups[g].second;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[g], 0);
    rc = ntt_inverse(ctx, (const Fr*)d_in + c0, d_coeffs + c0, log_h, gw, s, LAYOUT_NATURAL, width, width);
    if (rc == EON_OK && want_lde) {
      // the group's LDE goes out over PCIe (second copy stream: downloads run beside the uploads)
      // while its MSM, the long part, runs
      rc = lde_on_aux(ctx, d_coeffs + c0, (Fr*)d_lde + c0, lde_log_size, lde_log_size - log_h, gw, ls, width, 8 + (int)g);
      if (rc == EON_OK) {
        cudaError_t e = cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_pipe[8 + g], 0);
        // Downloads start when the LAST upload is through: the MSM's rounds cannot start before the last group is
        // on the device, so the uploads are on the critical path, while the downloads have the whole of the rounds
        // (30 ms at 2^20 x 16) to hide under -- and PCIe traffic in both directions at once slows the uploads
        // (measured: e2e 47.98 ms with the two directions overlapped).  EON_PIPE_D2H_EARLY=1: as soon as a group's
        // LDE is done.
        static const int d2h_early = getenv("EON_PIPE_D2H_EARLY") ? atoi(getenv("EON_PIPE_D2H_EARLY")) : 0;
        if (e == cudaSuccess && !d2h_early && g == 0 && groups.size() > 1)
          e = cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_pipe[groups.size() - 1], 0);
        if (e == cudaSuccess)
          e = cudaMemcpy2DAsync((Fr*)h_lde_out + c0, hpitch_out, (const Fr*)d_lde + c0, pitch, gw * sizeof(Fr), lde_rows,
                                cudaMemcpyDeviceToHost, ctx->copy_stream2);
        if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("LDE download failed: ") + cudaGetErrorString(e));
      }
    }
    // (no priority inversion here, unlike the device-resident entry point: with host buffers the LDE has to finish
    // EARLY, its 2^lde_log_size x width download is the long pole and hides under the MSM; measured at 2^20 x 16:
    // 52.9 ms this way, 64.2 ms with the MSM on the high-priority stream)
    if (rc == EON_OK && msm_per_group)
      rc = msm_run(ctx, ctx->d_srs, d_coeffs + c0, h, gw, width, (G1Affine*)d_commit + c0);
    if (rc == EON_OK && early_sort) rc = msm_batch_sort(ctx, msm_batch_state, d_coeffs + c0, width, c0, gw);
  }
  // one MSM over all columns, after the last group's iDFT (the uploads and the LDE downloads are long under way)
  if (rc == EON_OK && early_sort) rc = msm_batch_finish(ctx, msm_batch_state, (G1Affine*)d_commit);
  else if (rc == EON_OK && !msm_per_group) rc = msm_run(ctx, ctx->d_srs, d_coeffs, h, width, width, (G1Affine*)d_commit);
  if (rc == EON_OK) {
    cudaError_t e = cudaMemcpyAsync(h_commit_xy, d_commit, width * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && want_lde) e = cudaStreamSynchronize(ctx->copy_stream2);
    if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("commit readback failed: ") + cudaGetErrorString(e));
  }
  if (rc != EON_OK) {
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamSynchronize(ctx->copy_stream2);
    if (ctx->prio_stream) cudaStreamSynchronize(ctx->prio_stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_coeffs);
    return rc;
  }
  *out_handle = handle_new(ctx, d_coeffs, h, log_h, width, cap);
  return EON_OK;
}

int eon_kzg_commit(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                   uint64_t* h_commit_xy, eon_handle* out_handle) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_host(ctx, h_evals, 0, log_h, width, shift, h_commit_xy, out_handle, 0, nullptr, nullptr, 0);
}

int eon_kzg_commit_ld(eon_ctx* ctx, const uint64_t* h_evals, size_t ld_in, unsigned log_h, size_t width,
                      const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_host(ctx, h_evals, ld_in, log_h, width, shift, h_commit_xy, out_handle, 0, nullptr, nullptr, 0);
}

int eon_kzg_commit_lde(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                       uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                       const uint64_t lde_shift[4], uint64_t* h_lde_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  if (lde_log_size == 0) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_host(ctx, h_evals, 0, log_h, width, shift, h_commit_xy, out_handle, lde_log_size, lde_shift,
                         h_lde_out, 0);
}

int eon_kzg_commit_lde_ld(eon_ctx* ctx, const uint64_t* h_evals, size_t ld_in, unsigned log_h, size_t width,
                          const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                          const uint64_t lde_shift[4], uint64_t* h_lde_out, size_t ld_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  if (lde_log_size == 0) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_host(ctx, h_evals, ld_in, log_h, width, shift, h_commit_xy, out_handle, lde_log_size, lde_shift,
                         h_lde_out, ld_out);
}

// ---- Pcs::commit_quotient (commit/src/pcs.rs:82-102) ---------------------------------------------
// The trait default splits the quotient evaluations on shift*<omega_{2^log_size}> into 2^log_chunks matrices
// (row r -> chunk r mod 2^log_chunks, commit/src/domain.rs:188-221), pairs chunk i with the coset
// shift*omega^i of size 2^(log_size - log_chunks) (domain.rs:174-186) and commits them one after the other.
// Here nothing is de-interleaved: chunk i IS the column group [i*width, (i+1)*width) of the same buffer
// read with a row pitch of 2^log_chunks * width, the coset iDFTs write side by side into one
// h x (chunks*width) coefficient matrix, and ONE batched MSM commits every column of every chunk.
// chunk_first / chunk_count: the contiguous chunk range [chunk_first, chunk_first + chunk_count) this call commits
// (everything for the single-device entry points; the multi-device context gives every GPU a range).  d_evals
// holds ONLY those chunks: 2^(log_size - log_chunks) rows of chunk_count * width elements.
static int kzg_commit_quotient_locked(eon_ctx* ctx, const Fr* d_evals, unsigned log_size, size_t width,
                                      unsigned log_chunks, size_t chunk_first, size_t chunk_count,
                                      const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handles) {
  if (log_chunks > log_size) return fail(ctx, EON_ERR_BAD_ARG, "more chunks than quotient rows");
  if (log_chunks > 16) return fail(ctx, EON_ERR_BAD_ARG, "too many quotient chunks");
  EON_TRY(check_dims(ctx, log_size, width));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  if (!out_handles) return fail(ctx, EON_ERR_BAD_ARG, "null handle pointer");
  const size_t nchunks = chunk_count;
  if (chunk_first + chunk_count > ((size_t)1 << log_chunks)) return fail(ctx, EON_ERR_BAD_ARG, "chunk range out of bounds");
  for (size_t i = 0; i < nchunks; i++) out_handles[i] = 0;
  const unsigned log_h = log_size - log_chunks;
  const size_t h = (size_t)1 << log_h;
  const size_t cw = nchunks * width;
  if (cw > 0xffffffffull) return fail(ctx, EON_ERR_BAD_ARG, "width too large");
  if (h > ctx->srs_n) {  // ensure_supported(height - 1) of every chunk, kzg/src/pcs.rs:238-240
    char b[128];
    snprintf(b, sizeof(b), "DegreeTooLarge: degree %zu > max %zu", h - 1, ctx->srs_n ? ctx->srs_n - 1 : 0);
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, b);
  }
  if (cw && (!d_evals || !h_commit_xy)) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  void *d_comb = nullptr, *d_res = nullptr;
  EON_TRY(scratch_get(ctx, SC_QUOT, h * cw * sizeof(Fr) + 32, &d_comb));
  EON_TRY(scratch_get(ctx, SC_MSM_RESULT, cw * sizeof(G1Affine) + 64, &d_res));
  const Fr g = fr_two_adic_generator(log_size);
  Fr si = s;
  for (size_t i = 0; i < chunk_first; i++) si = fp_mul(si, g);
  for (size_t i = 0; i < nchunks; i++) {
    EON_TRY(ntt_inverse(ctx, d_evals + i * width, (Fr*)d_comb + i * width, log_h, width, si, LAYOUT_NATURAL, cw, cw));
    si = fp_mul(si, g);
  }
  EON_TRY(msm_run(ctx, ctx->d_srs, (const Fr*)d_comb, h, cw, cw, (G1Affine*)d_res));
  if (cw)
    EON_CUDA(ctx, cudaMemcpyAsync(h_commit_xy, d_res, cw * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  // MatrixProverData.coeffs of every chunk (pcs.rs:252-256): dense h x width matrices behind their own handles
  std::vector<std::pair<Fr*, size_t>> bufs;
  int rc = EON_OK;
  for (size_t i = 0; rc == EON_OK && i < nchunks; i++) {
    Fr* d_coeffs = nullptr;
    size_t cap = 0;
    rc = coeff_buffer_get(ctx, mat_bytes(log_h, width) + 32, &d_coeffs, &cap);
    if (rc != EON_OK) break;
    bufs.push_back(std::make_pair(d_coeffs, cap));
    if (width) {
      cudaError_t e = cudaMemcpy2DAsync(d_coeffs, width * sizeof(Fr), (const Fr*)d_comb + i * width, cw * sizeof(Fr),
                                        width * sizeof(Fr), h, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("chunk coefficient copy failed: ") + cudaGetErrorString(e));
    }
  }
  if (rc == EON_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, "commit_quotient failed");
  if (rc != EON_OK) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : bufs) cudaFree(b.first);
    return rc;
  }
  for (size_t i = 0; i < nchunks; i++) out_handles[i] = handle_new(ctx, bufs[i].first, h, log_h, width, bufs[i].second);
  return EON_OK;
}

// host evaluations (the WHOLE quotient matrix, 2^log_size x width) -> the chunk range of this device
static int kzg_commit_quotient_host(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_size, size_t width,
                                    unsigned log_chunks, size_t chunk_first, size_t chunk_count, const uint64_t shift[4],
                                    uint64_t* h_commit_xy, eon_handle* out_handles) {
  EON_TRY(check_dims(ctx, log_size, width));
  if (log_chunks > log_size || log_chunks > 16) return fail(ctx, EON_ERR_BAD_ARG, "bad chunk count");
  const size_t nchunks = (size_t)1 << log_chunks;
  if (chunk_first + chunk_count > nchunks) return fail(ctx, EON_ERR_BAD_ARG, "chunk range out of bounds");
  const size_t h = (size_t)1 << (log_size - log_chunks);
  const size_t b = h * chunk_count * width * sizeof(Fr);
  if (b && !h_evals) return fail(ctx, EON_ERR_BAD_ARG, "null evals");
  void* d_in = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_A, b + 32, &d_in));
  if (b) {
    if (chunk_count == nchunks)
      EON_CUDA(ctx, cudaMemcpyAsync(d_in, h_evals, b, cudaMemcpyHostToDevice, ctx->stream));
    else
      EON_CUDA(ctx, cudaMemcpy2DAsync(d_in, chunk_count * width * sizeof(Fr), (const Fr*)h_evals + chunk_first * width,
                                      nchunks * width * sizeof(Fr), chunk_count * width * sizeof(Fr), h,
                                      cudaMemcpyHostToDevice, ctx->stream));
  }
  return kzg_commit_quotient_locked(ctx, (const Fr*)d_in, log_size, width, log_chunks, chunk_first, chunk_count, shift,
                                    h_commit_xy, out_handles);
}

int eon_kzg_commit_quotient_dev(eon_ctx* ctx, const uint64_t* d_evals, unsigned log_size, size_t width,
                                unsigned log_chunks, const uint64_t shift[4], uint64_t* h_commit_xy,
                                eon_handle* out_handles) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (log_chunks > 16) return fail(ctx, EON_ERR_BAD_ARG, "too many quotient chunks");
  return kzg_commit_quotient_locked(ctx, (const Fr*)d_evals, log_size, width, log_chunks, 0, (size_t)1 << log_chunks,
                                    shift, h_commit_xy, out_handles);
}

int eon_kzg_commit_quotient(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_size, size_t width, unsigned log_chunks,
                            const uint64_t shift[4], uint64_t* h_commit_xy, eon_handle* out_handles) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (log_chunks > 16) return fail(ctx, EON_ERR_BAD_ARG, "too many quotient chunks");
  return kzg_commit_quotient_host(ctx, h_evals, log_size, width, log_chunks, 0, (size_t)1 << log_chunks, shift,
                                  h_commit_xy, out_handles);
}

static int find_handle(eon_ctx* ctx, eon_handle h, ProverMatrix* pm) {
  auto it = ctx->handles.find(h);
  if (it == ctx->handles.end()) return fail(ctx, EON_ERR_BAD_HANDLE, "unknown prover-data handle");
  *pm = it->second;
  return EON_OK;
}

int eon_handle_dims(eon_ctx* ctx, eon_handle h, unsigned* log_h, size_t* width) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  if (log_h) *log_h = pm.log_h;
  if (width) *width = pm.width;
  return EON_OK;
}

int eon_handle_free(eon_ctx* ctx, eon_handle h) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  ctx->handles.erase(h);
  if (ctx->coeff_pool.size() < 4) {
    // stream order makes reuse safe: later work on this buffer is queued behind earlier readers
    ctx->coeff_pool.push_back(std::make_pair(pm.cap, (void*)pm.d_coeffs));
    return EON_OK;
  }
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  EON_CUDA(ctx, cudaFree(pm.d_coeffs));
  return EON_OK;
}

int eon_kzg_read_coeffs(eon_ctx* ctx, eon_handle h, uint64_t* h_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  size_t b = pm.rows * pm.width * sizeof(Fr);
  if (b == 0) return EON_OK;
  if (!h_out) return fail(ctx, EON_ERR_BAD_ARG, "null output");
  EON_CUDA(ctx, cudaMemcpyAsync(h_out, pm.d_coeffs, b, cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

static int evals_on_coset_locked(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], Fr* d_out) {
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  if (pm.log_h == NOT_POW2)
    return fail(ctx, EON_ERR_BAD_ARG, "prover data has a non-power-of-two height (KzgMmcs matrix): no coset evaluation");
  EON_TRY(check_dims(ctx, log_size, pm.width));
  if (pm.width == 0) return EON_OK;
  if (!d_out) return fail(ctx, EON_ERR_BAD_ARG, "null output");
  if (log_size < pm.log_h) {
    // a coset smaller than the polynomial (allowed by the Horner loop of kzg/src/pcs.rs:278-286): reduce the
    // coefficients mod X^n - shift^n first, then the size-n coset NTT
    void* folded = nullptr;
    EON_TRY(scratch_get(ctx, SC_QUOT_AUX, mat_bytes(log_size, pm.width) + 32, &folded));
    EON_TRY(fold_coeffs_run(ctx, pm.d_coeffs, pm.rows, pm.width, log_size, s, (Fr*)folded));
    return ntt_forward(ctx, (const Fr*)folded, d_out, log_size, 0, pm.width, s, LAYOUT_NATURAL);
  }
  return ntt_forward(ctx, pm.d_coeffs, d_out, log_size, log_size - pm.log_h, pm.width, s, LAYOUT_NATURAL);
}

int eon_kzg_evals_on_coset_dev(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4],
                               uint64_t* d_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return evals_on_coset_locked(ctx, h, log_size, shift, (Fr*)d_out);
}

static int evals_on_coset_host(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out,
                               size_t h_ld_out) {
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  if (log_size > 28) return fail(ctx, EON_ERR_TWO_ADICITY, "domain exceeds 2^28");
  if (h_ld_out == 0) h_ld_out = pm.width;
  if (h_ld_out < pm.width) return fail(ctx, EON_ERR_BAD_ARG, "row pitch < width");
  const size_t hpitch = h_ld_out * sizeof(Fr);
  size_t b = mat_bytes(log_size, pm.width);
  void* d_out = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_B, b + 32, &d_out));
  auto groups = column_groups(pm.width, b, false);
  const size_t w = pm.width, rows = (size_t)1 << log_size, pitch = w * sizeof(Fr);
  if (groups.size() == 1 || pm.log_h == NOT_POW2 || log_size < pm.log_h) {
    EON_TRY(evals_on_coset_locked(ctx, h, log_size, shift, (Fr*)d_out));
    if (b) {
      if (!h_out) return fail(ctx, EON_ERR_BAD_ARG, "null output");
      if (h_ld_out == w) EON_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, b, cudaMemcpyDeviceToHost, ctx->stream));
      else EON_CUDA(ctx, cudaMemcpy2DAsync(h_out, hpitch, d_out, pitch, pitch, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EON_OK;
  }
  // pipelined: D2H of group g (copy stream) under the coset NTT of group g+1 (compute stream)
  Fr s;
  EON_TRY(check_shift(ctx, shift, &s));
  EON_TRY(check_dims(ctx, log_size, pm.width));
  if (!h_out) return fail(ctx, EON_ERR_BAD_ARG, "null output");
  EON_TRY(pipe_init(ctx));
  int rc = EON_OK;
  for (size_t g = 0; rc == EON_OK && g < groups.size(); g++) {
    const size_t c0 = groups[g].first, gw = groups[g].second;
    rc = ntt_forward(ctx, pm.d_coeffs + c0, (Fr*)d_out + c0, log_size, log_size - pm.log_h, gw, s, LAYOUT_NATURAL, w, w);
    if (rc != EON_OK) break;
    cudaError_t e = cudaEventRecord(ctx->ev_pipe[g], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[g], 0);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync((Fr*)h_out + c0, hpitch, (const Fr*)d_out + c0, pitch, gw * sizeof(Fr), rows,
                            cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e != cudaSuccess) rc = fail(ctx, EON_ERR_CUDA, std::string("group download failed: ") + cudaGetErrorString(e));
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
  if (rc == EON_OK && (e1 != cudaSuccess || e2 != cudaSuccess))
    rc = fail(ctx, EON_ERR_CUDA, std::string("download failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return rc;
}

int eon_kzg_evals_on_coset(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return evals_on_coset_host(ctx, h, log_size, shift, h_out, 0);
}

int eon_kzg_evals_on_coset_ld(eon_ctx* ctx, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out,
                              size_t ld_out) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return evals_on_coset_host(ctx, h, log_size, shift, h_out, ld_out);
}

int eon_quotient_and_eval_dev(eon_ctx* ctx, const uint64_t* d_coeffs, size_t h, size_t width, const uint64_t z[4],
                              uint64_t* d_quot, uint64_t* h_values) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (!z || !fr_wire_is_canonical(z)) return fail(ctx, EON_ERR_BAD_ARG, "z is not a canonical Fr");
  if (width == 0) return EON_OK;
  if ((h && (!d_coeffs || !d_quot)) || !h_values) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  void* d_vals = nullptr;
  EON_TRY(scratch_get(ctx, SC_MSM_RESULT, width * sizeof(Fr) + 64, &d_vals));
  EON_TRY(quotient_run(ctx, (const Fr*)d_coeffs, h, width, width, fr_from_wire(z), (Fr*)d_quot, (Fr*)d_vals));
  EON_CUDA(ctx, cudaMemcpyAsync(h_values, d_vals, width * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

int eon_kzg_open(eon_ctx* ctx, eon_handle h, const uint64_t* h_points, size_t npoints, uint64_t* h_values,
                 uint64_t* h_witness_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  ProverMatrix pm;
  EON_TRY(find_handle(ctx, h, &pm));
  const size_t w = pm.width, rows = pm.rows;
  if (npoints == 0 || w == 0) return EON_OK;
  if (!h_points || !h_values || !h_witness_xy) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  for (size_t p = 0; p < npoints; p++)
    if (!fr_wire_is_canonical(h_points + 4 * p)) return fail(ctx, EON_ERR_BAD_ARG, "opening point is not canonical");
  // the quotient of a length-h polynomial has h-1 coefficients: witness = MSM over srs[..h-1]
  // (commit_column(&quotient), kzg/src/pcs.rs:316); degree guard as in util.rs:38
  if (rows && rows - 1 > ctx->srs_n) return fail(ctx, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: quotient longer than the SRS");
  const size_t ncols = npoints * w;
  void *d_quot = nullptr, *d_vals = nullptr, *d_wit = nullptr;
  EON_TRY(scratch_get(ctx, SC_QUOT, rows * ncols * sizeof(Fr) + 32, &d_quot));
  EON_TRY(scratch_get(ctx, SC_IO_B, ncols * (sizeof(Fr) + sizeof(G1Affine)) + 64, &d_vals));
  d_wit = (char*)d_vals + ncols * sizeof(Fr);
  for (size_t p = 0; p < npoints; p++) {
    Fr z = fr_from_wire(h_points + 4 * p);
    EON_TRY(quotient_run(ctx, pm.d_coeffs, rows, w, ncols, z, (Fr*)d_quot + p * w, (Fr*)d_vals + p * w));
  }
  EON_TRY(msm_run(ctx, ctx->d_srs, (const Fr*)d_quot, rows ? rows - 1 : 0, ncols, ncols, (G1Affine*)d_wit));
  EON_CUDA(ctx, cudaMemcpyAsync(h_values, d_vals, ncols * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaMemcpyAsync(h_witness_xy, d_wit, ncols * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

// KzgPcs::open over every (matrix, point) of a call at once (kzg/src/pcs.rs:289-335 walks rounds, matrices, points
// and columns one commit_column at a time): all quotients land side by side in one matrix and ONE batched MSM
// produces every witness.  Columns of shorter matrices are zero below their own length.
int eon_kzg_open_batch(eon_ctx* ctx, size_t nmat, const eon_handle* handles, const size_t* npoints,
                       const uint64_t* h_points, uint64_t* h_values, uint64_t* h_witness_xy) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (nmat == 0) return EON_OK;
  if (!handles || !npoints) return fail(ctx, EON_ERR_BAD_ARG, "null handle / point-count array");
  std::vector<ProverMatrix> pms(nmat);
  size_t total_cols = 0, total_points = 0, max_rows = 0;
  bool ragged = false;
  for (size_t m = 0; m < nmat; m++) {
    EON_TRY(find_handle(ctx, handles[m], &pms[m]));
    if (npoints[m] == 0 || pms[m].width == 0) {
      total_points += npoints[m];
      continue;
    }
    if (npoints[m] > (~(size_t)0 - total_cols) / pms[m].width) return fail(ctx, EON_ERR_BAD_ARG, "too many openings");
    total_cols += npoints[m] * pms[m].width;
    total_points += npoints[m];
    max_rows = std::max(max_rows, pms[m].rows);
  }
  if (total_cols == 0) return EON_OK;
  if (!h_points || !h_values || !h_witness_xy) return fail(ctx, EON_ERR_BAD_ARG, "null buffer");
  for (size_t p = 0; p < total_points; p++)
    if (!fr_wire_is_canonical(h_points + 4 * p)) return fail(ctx, EON_ERR_BAD_ARG, "opening point is not canonical");
  for (size_t m = 0; m < nmat; m++)
    if (npoints[m] && pms[m].width && pms[m].rows != max_rows) ragged = true;
  // the quotient of a length-h polynomial has h-1 coefficients (commit_column(&quotient), pcs.rs:316; util.rs:38)
  if (max_rows && max_rows - 1 > ctx->srs_n)
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: quotient longer than the SRS");
  if (total_cols > 0xffffffffull || (max_rows && total_cols > (~(size_t)0) / (max_rows * sizeof(Fr))))
    return fail(ctx, EON_ERR_BAD_ARG, "opening too large");
  void *d_quot = nullptr, *d_vals = nullptr;
  EON_TRY(scratch_get(ctx, SC_QUOT, max_rows * total_cols * sizeof(Fr) + 32, &d_quot));
  EON_TRY(scratch_get(ctx, SC_IO_B, total_cols * (sizeof(Fr) + sizeof(G1Affine)) + 64, &d_vals));
  void* d_wit = (char*)d_vals + total_cols * sizeof(Fr);
  if (ragged) EON_CUDA(ctx, cudaMemsetAsync(d_quot, 0, max_rows * total_cols * sizeof(Fr), ctx->stream));
  size_t col = 0, pt = 0;
  for (size_t m = 0; m < nmat; m++) {
    const ProverMatrix& pm = pms[m];
    for (size_t p = 0; p < npoints[m]; p++, pt++) {
      if (pm.width == 0) continue;
      Fr z = fr_from_wire(h_points + 4 * pt);
      EON_TRY(quotient_run(ctx, pm.d_coeffs, pm.rows, pm.width, total_cols, z, (Fr*)d_quot + col, (Fr*)d_vals + col));
      col += pm.width;
    }
  }
  EON_TRY(msm_run(ctx, ctx->d_srs, (const Fr*)d_quot, max_rows ? max_rows - 1 : 0, total_cols, total_cols,
                  (G1Affine*)d_wit));
  EON_CUDA(ctx, cudaMemcpyAsync(h_values, d_vals, total_cols * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaMemcpyAsync(h_witness_xy, d_wit, total_cols * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

}  // extern "C"

namespace eon {
int dft_host(eon_ctx* ctx, int kind, const uint64_t* h_in, size_t ld_in, uint64_t* h_out, size_t ld_out, unsigned log_h,
             size_t width, unsigned added_bits, const uint64_t shift[4]) {
  Lock lk(ctx);
  return dft_host_locked(ctx, kind, h_in, ld_in, h_out, ld_out, log_h, width, added_bits, shift);
}
int kzg_commit_quotient_range_host(eon_ctx* ctx, const uint64_t* h_evals, unsigned log_size, size_t width,
                                   unsigned log_chunks, size_t chunk_first, size_t chunk_count, const uint64_t shift[4],
                                   uint64_t* h_commit_xy, eon_handle* out_handles) {
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (log_chunks > 16) return fail(ctx, EON_ERR_BAD_ARG, "too many quotient chunks");
  return kzg_commit_quotient_host(ctx, h_evals, log_size, width, log_chunks, chunk_first, chunk_count, shift, h_commit_xy,
                                  out_handles);
}
int kzg_commit_coeffs_host_ld(eon_ctx* ctx, const uint64_t* h_coeffs, size_t ld, size_t rows, size_t width,
                              uint64_t* h_commit_xy, eon_handle* out_handle) {
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return kzg_commit_coeffs_locked(ctx, h_coeffs, true, rows, width, h_commit_xy, out_handle, ld);
}
int msm_srs_range_host_partial(eon_ctx* ctx, const uint64_t* h_scalars_rows, size_t first, size_t n, size_t ncols,
                               size_t ld, const G1Affine** d_partial) {
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (first > ctx->srs_n || n > ctx->srs_n - first)
    return fail(ctx, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: polynomial longer than the SRS");
  if (ld < ncols) return fail(ctx, EON_ERR_BAD_ARG, "ld < ncols");
  if (n && ncols && !h_scalars_rows) return fail(ctx, EON_ERR_BAD_ARG, "null scalars");
  void *d_sc = nullptr, *d_res = nullptr;
  const size_t wb = ncols * sizeof(Fr);
  EON_TRY(scratch_get(ctx, SC_IO_A, n * wb + 32, &d_sc));
  EON_TRY(scratch_get(ctx, SC_MSM_RESULT, ncols * sizeof(G1Affine) + 64, &d_res));
  if (n && ncols) {
    if (ld == ncols) EON_CUDA(ctx, cudaMemcpyAsync(d_sc, h_scalars_rows, n * wb, cudaMemcpyHostToDevice, ctx->stream));
    else EON_CUDA(ctx, cudaMemcpy2DAsync(d_sc, wb, h_scalars_rows, ld * sizeof(Fr), wb, n, cudaMemcpyHostToDevice, ctx->stream));
  }
  EON_TRY(msm_run(ctx, ctx->d_srs + first, (const Fr*)d_sc, n, ncols, ncols, (G1Affine*)d_res));
  *d_partial = (const G1Affine*)d_res;
  return EON_OK;
}
int msm_srs_host_ld(eon_ctx* ctx, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy) {
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  if (n > ctx->srs_n) return fail(ctx, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: polynomial longer than the SRS");
  if (n && ncols && !h_scalars) return fail(ctx, EON_ERR_BAD_ARG, "null scalars");
  if (ld < ncols) return fail(ctx, EON_ERR_BAD_ARG, "ld < ncols");
  void* d_sc = nullptr;
  const size_t wb = ncols * sizeof(Fr);
  EON_TRY(scratch_get(ctx, SC_IO_A, n * wb + 32, &d_sc));
  if (n && ncols) {
    if (ld == ncols) EON_CUDA(ctx, cudaMemcpyAsync(d_sc, h_scalars, n * wb, cudaMemcpyHostToDevice, ctx->stream));
    else EON_CUDA(ctx, cudaMemcpy2DAsync(d_sc, wb, h_scalars, ld * sizeof(Fr), wb, n, cudaMemcpyHostToDevice, ctx->stream));
  }
  return msm_to_host(ctx, ctx->d_srs, (const Fr*)d_sc, n, ncols, ncols, h_out_xy);
}
}  // namespace eon

extern "C" {

// ---- measurement ---------------------------------------------------------------------------------
int eon_bench_imad_peak(eon_ctx* ctx, int kind, double* out_tops) {
  if (!ctx || !out_tops) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return bench_imad(ctx, kind, out_tops);
}
int eon_bench_modmul(eon_ctx* ctx, int field, double* out_gmuls) {
  if (!ctx || !out_gmuls) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return bench_modmul(ctx, field, 0, out_gmuls);
}
int eon_bench_modmul_variant(eon_ctx* ctx, int field, int variant, double* out_gmuls) {
  if (!ctx || !out_gmuls || variant < 0 || variant > 5) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  return bench_modmul(ctx, field, variant, out_gmuls);
}

int eon_ntt_twiddle_form(void) { return ntt_twiddles_are_fixed_operand() ? 1 : 0; }

int eon_last_phase_ms(eon_ctx* ctx, int phase, float* out_ms) {
  if (!ctx || !out_ms || phase < 0 || phase >= PH_COUNT) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  *out_ms = 0.f;
  for (auto& pr : ctx->ev_pairs[phase]) {
    float ms = 0.f;
    EON_CUDA(ctx, cudaEventSynchronize(pr.second));
    EON_CUDA(ctx, cudaEventElapsedTime(&ms, pr.first, pr.second));
    *out_ms += ms;
  }
  return EON_OK;
}

int eon_phase_reset(eon_ctx* ctx) {
  if (!ctx) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  phase_reset(ctx);
  return EON_OK;
}

// measurement helper: strided (2-D) copy rate between a pinned host matrix and the device, ms per copy
int eon_bench_copy2d(eon_ctx* ctx, void* h_ptr, size_t rows, size_t width_bytes, size_t pitch_bytes, int to_device,
                     float* out_ms) {
  if (!ctx || !h_ptr || !out_ms) return EON_ERR_BAD_ARG;
  Lock lk(ctx);
  EON_TRY(set_device(ctx));
  void* d = nullptr;
  EON_TRY(scratch_get(ctx, SC_IO_A, rows * width_bytes + 32, &d));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 2; it++) {
    cudaEventRecord(e0, ctx->stream);
    cudaError_t rc = to_device ? cudaMemcpy2DAsync(d, width_bytes, h_ptr, pitch_bytes, width_bytes, rows,
                                                   cudaMemcpyHostToDevice, ctx->stream)
                               : cudaMemcpy2DAsync(h_ptr, pitch_bytes, d, width_bytes, width_bytes, rows,
                                                   cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(e1, ctx->stream);
    if (rc != cudaSuccess) return fail(ctx, EON_ERR_CUDA, cudaGetErrorString(rc));
    EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  cudaEventElapsedTime(out_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return EON_OK;
}

int eon_phase_count(void) { return PH_COUNT; }
const char* eon_phase_name(int phase) {
  static const char* names[PH_COUNT] = {"ntt_twiddle", "ntt_passes", "msm_digits", "msm_scan",
                                         "msm_scatter", "msm_accumulate", "msm_reduce", "quotient",
                                         "msm_tree_fwd", "msm_tree_inv", "msm_tree_bwd", "msm_finish"};
  return (phase >= 0 && phase < PH_COUNT) ? names[phase] : "?";
}

}  // extern "C"
