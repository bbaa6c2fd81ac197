// KZG opening quotient as a parallel suffix scan, plus integer-pipe microbenchmarks.
//
// Replaces quotient_and_eval (kzg/src/util.rs:100-111):
//     carry = c[h-1]; for i = h-2..0 { q[i] = carry; carry = c[i] + carry*z }   -> (q, carry = f(z))
// i.e. q[i] = sum_{j>i} c[j] z^(j-i-1).  The serial recurrence is cut into chunks of B rows:
//   A  local Horner value of every chunk                      (thread per chunk x column)
//   B  suffix values at chunk boundaries, two-level            (super-chunks of B2 chunks)
//   C  replay every chunk from its incoming carry, emit q      (thread per chunk x column)
// ~2 modmul per coefficient; exact field arithmetic => identical quotient and evaluation.
#include "common.cuh"
#include "fp_shoup.cuh"

namespace eon {

constexpr u32 QB = 256;   // rows per chunk
constexpr u32 QB2 = 128;  // chunks per super-chunk

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(q), hi = __ldg(q + 1);
  Fr r;
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// A: local[chunk][col] = sum_{j in chunk} c[j][col] * z^(j - chunk_start)
__global__ void __launch_bounds__(256)
k_quot_local(const Fr* __restrict__ coeffs, size_t h, size_t width, Fr z, size_t nchunks, Fr* __restrict__ local) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= nchunks * width) return;
  size_t ch = g / width, col = g % width;
  size_t s = ch * QB, e = min(h, s + QB);
  Fr acc = Fr::zero();
  for (size_t j = e; j-- > s;) acc = fp_add(fp_mul(acc, z), ld_fr(coeffs + j * width + col));
  st_fr(local + g, acc);
}

// generic level: out[k][col] = sum_{i in group k} in[i][col] * m^(i - group_start)
__global__ void __launch_bounds__(256)
k_quot_group(const Fr* __restrict__ in, size_t nin, size_t width, Fr m, u32 group, size_t ngroups, Fr* __restrict__ out) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= ngroups * width) return;
  size_t k = g / width, col = g % width;
  size_t s = k * group, e = min(nin, s + (size_t)group);
  Fr acc = Fr::zero();
  for (size_t i = e; i-- > s;) acc = fp_add(fp_mul(acc, m), ld_fr(in + i * width + col));
  st_fr(out + g, acc);
}

// serial suffix over the (few) groups of one column: carry[k] = sum_{i>k} v[i] * m^(i-k-1)
__global__ void __launch_bounds__(64)
k_quot_suffix_serial(const Fr* __restrict__ v, size_t n, size_t width, Fr m, Fr* __restrict__ carry) {
  size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (col >= width) return;
  Fr c = Fr::zero();
  for (size_t k = n; k-- > 0;) {
    st_fr(carry + k * width + col, c);
    c = fp_add(fp_mul(c, m), ld_fr(v + k * width + col));
  }
}

// expand group carries to member carries: for members i of group k (descending)
//   carry_out[i] = c;  c = in[i] + c * m      starting from c = carry_in[k]
__global__ void __launch_bounds__(256)
k_quot_expand(const Fr* __restrict__ in, size_t nin, size_t width, Fr m, u32 group, size_t ngroups,
              const Fr* __restrict__ carry_in, Fr* __restrict__ carry_out) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= ngroups * width) return;
  size_t k = g / width, col = g % width;
  size_t s = k * group, e = min(nin, s + (size_t)group);
  Fr c = ld_fr(carry_in + g);
  for (size_t i = e; i-- > s;) {
    st_fr(carry_out + i * width + col, c);
    c = fp_add(fp_mul(c, m), ld_fr(in + i * width + col));
  }
}

// C: replay each chunk from its carry; write q rows; chunk 0 also writes f(z)
__global__ void __launch_bounds__(256)
k_quot_final(const Fr* __restrict__ coeffs, size_t h, size_t width, Fr z, size_t nchunks,
             const Fr* __restrict__ carry, Fr* __restrict__ quot, size_t ld_out, Fr* __restrict__ values) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= nchunks * width) return;
  size_t ch = g / width, col = g % width;
  size_t s = ch * QB, e = min(h, s + QB);
  Fr c = ld_fr(carry + g);
  for (size_t j = e; j-- > s;) {
    st_fr(quot + j * ld_out + col, c);
    c = fp_add(fp_mul(c, z), ld_fr(coeffs + j * width + col));
  }
  if (ch == 0) st_fr(values + col, c);
}

__global__ void k_zero_fr(Fr* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) st_fr(p + i, Fr::zero());
}

int quotient_run(eon_ctx* ctx, const Fr* d_coeffs, size_t h, size_t width, size_t ld_out, const Fr& z, Fr* d_quot,
                 Fr* d_values) {
  if (width == 0) return EON_OK;
  cudaStream_t st = ctx->stream;
  if (h == 0) {  // empty polynomial: value 0 (util.rs:101-103)
    k_zero_fr<<<(unsigned)((width + 255) / 256), 256, 0, st>>>(d_values, width);
    EON_LAUNCHED(ctx);
    return EON_OK;
  }
  phase_begin(ctx, PH_QUOTIENT);
  const size_t nchunks = (h + QB - 1) / QB;
  const size_t nsuper = (nchunks + QB2 - 1) / QB2;
  void* aux = nullptr;
  // local[nchunks], carry[nchunks], local2[nsuper], carry2[nsuper]   (each x width)
  EON_TRY(scratch_get(ctx, SC_QUOT_AUX, (2 * nchunks + 2 * nsuper) * width * sizeof(Fr), &aux));
  Fr* local = (Fr*)aux;
  Fr* carry = local + nchunks * width;
  Fr* local2 = carry + nchunks * width;
  Fr* carry2 = local2 + nsuper * width;
  const Fr zB = fp_pow_u64(z, QB);
  const Fr zBB = fp_pow_u64(zB, QB2);
  auto blocks = [](size_t t) { return (unsigned)((t + 255) / 256); };
  k_quot_local<<<blocks(nchunks * width), 256, 0, st>>>(d_coeffs, h, width, z, nchunks, local);
  EON_LAUNCHED(ctx);
  k_quot_group<<<blocks(nsuper * width), 256, 0, st>>>(local, nchunks, width, zB, QB2, nsuper, local2);
  EON_LAUNCHED(ctx);
  k_quot_suffix_serial<<<(unsigned)((width + 63) / 64), 64, 0, st>>>(local2, nsuper, width, zBB, carry2);
  EON_LAUNCHED(ctx);
  k_quot_expand<<<blocks(nsuper * width), 256, 0, st>>>(local, nchunks, width, zB, QB2, nsuper, carry2, carry);
  EON_LAUNCHED(ctx);
  k_quot_final<<<blocks(nchunks * width), 256, 0, st>>>(d_coeffs, h, width, z, nchunks, carry, d_quot, ld_out,
                                                        d_values);
  EON_LAUNCHED(ctx);
  phase_end(ctx, PH_QUOTIENT);
  return EON_OK;
}

// ---- coefficient folding: p mod (X^n - m) ---------------------------------------------------------
// get_evaluations_on_domain (kzg/src/pcs.rs:278-286) evaluates the committed polynomial by Horner on ANY coset,
// including one smaller than the polynomial: on x = s*omega_n^k every x^n equals m = s^n, so
//     p(x) = sum_{i<n} ( sum_j c[i + j n] m^j ) x^i
// and the size-n coset NTT of the folded coefficients gives exactly the values the Horner loop gives.
__global__ void __launch_bounds__(256)
k_fold_coeffs(const Fr* __restrict__ coeffs, size_t h, size_t width, size_t n, Fr m, Fr* __restrict__ out) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= n * width) return;
  size_t i = g / width, col = g % width;
  size_t top = i + ((h - 1 - i) / n) * n;  // last row congruent to i (h > i because n <= h)
  Fr acc = ld_fr(coeffs + top * width + col);
  for (size_t j = top; j >= n + i; ) {
    j -= n;
    acc = fp_add(fp_mul(acc, m), ld_fr(coeffs + j * width + col));
  }
  st_fr(out + g, acc);
}

int fold_coeffs_run(eon_ctx* ctx, const Fr* d_coeffs, size_t h, size_t width, unsigned log_n, const Fr& shift,
                    Fr* d_out) {
  if (width == 0 || h == 0) return EON_OK;
  const size_t n = (size_t)1 << log_n;
  Fr m = shift;
  for (unsigned i = 0; i < log_n; i++) m = fp_sqr(m);  // shift^n
  k_fold_coeffs<<<(unsigned)((n * width + 255) / 256), 256, 0, ctx->stream>>>(d_coeffs, h, width, n, m, d_out);
  EON_LAUNCHED(ctx);
  return EON_OK;
}

// ---- integer-pipe microbenchmarks (roofline denominators) ----------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) k_imad_peak(u32* out, u32 iters, u32 seed) {
  // 16 independent accumulator chains per thread; multiplier and addend change every step
  u32 a[16];
  u64 wacc[8];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = seed + threadIdx.x * 16 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) wacc[i] = ((u64)a[2 * i] << 32) | a[2 * i + 1];
  u32 m = seed | 1;
  for (u32 it = 0; it < iters; it++) {
    if (KIND == 0) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = a[i] * m + a[(i + 1) & 15];
    } else if (KIND == 1) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = __umulhi(a[i], m) + a[(i + 1) & 15];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) wacc[i] = (u64)(u32)wacc[i] * (u64)m + wacc[(i + 1) & 7];
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r ^= a[i];
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= (u32)wacc[i] ^ (u32)(wacc[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int bench_imad(eon_ctx* ctx, int kind, double* out_tops) {
  const unsigned blocks = (unsigned)ctx->num_sms * 8;
  const u32 iters = 4096;
  void* buf = nullptr;
  EON_TRY(scratch_get(ctx, SC_SMALL, (size_t)blocks * 256 * sizeof(u32), &buf));
  cudaEvent_t e0, e1;
  EON_CUDA(ctx, cudaEventCreate(&e0));
  EON_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    EON_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (kind == 0) k_imad_peak<0><<<blocks, 256, 0, ctx->stream>>>((u32*)buf, iters, 12345u + rep);
    else if (kind == 1) k_imad_peak<1><<<blocks, 256, 0, ctx->stream>>>((u32*)buf, iters, 12345u + rep);
    else k_imad_peak<2><<<blocks, 256, 0, ctx->stream>>>((u32*)buf, iters, 12345u + rep);
    EON_LAUNCHED(ctx);
    EON_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    EON_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    EON_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double per_thread = (kind == 2) ? 8.0 * 2.0 : 16.0;  // a wide mad counts as lo + hi
  double ops = (double)blocks * 256.0 * iters * per_thread;
  *out_tops = ops / (best * 1e-3) / 1e12;
  return EON_OK;
}

// VARIANT: 0 = fp_mul as the library uses it, 1 = word-serial (CIOS) product, 2 = split product
// (Karatsuba + separate reduction), 3 = dedicated square + one modular add
template <class PP, int VARIANT>
__device__ __forceinline__ Fp<PP> bench_product(const Fp<PP>& a, const Fp<PP>& b) {
  if (VARIANT == 0) return fp_mul(a, b);
  u32 t[8];
  if (VARIANT == 1) fp_mul_lazy_cios<PP>(t, a.v, b.v);
  else if (VARIANT == 2) fp_mul_lazy_split<PP>(t, a.v, b.v);
  else fp_sqr_lazy_split<PP>(t, a.v);
  Fp<PP> r;
  fp_final_sub<PP>(r.v, t);
  return VARIANT == 3 ? fp_add(r, b) : r;  // the add keeps the two chains dependent on each other
}

template <class PP, int VARIANT>
__global__ void __launch_bounds__(256) k_modmul_peak(Fp<PP>* out, u32 iters, u32 seed) {
  Fp<PP> a = fp_from_u64<PP>(seed + threadIdx.x + 1);
  Fp<PP> b = fp_from_u64<PP>(seed * 3 + blockIdx.x + 7);
  Fp<PP> c = fp_from_u64<PP>(seed * 5 + threadIdx.x * 11 + 13);
  if constexpr (VARIANT >= 4) {
    // fixed-operand products as the NTT butterflies use them: the multiplier (b, b2) never changes, the
    // running values stay lazily reduced.  4 = Shoup form (fp_shoup.cuh; bq / b2q stand in for the precomputed
    // floor(w 2^256 / p): the instruction stream does not depend on their values), 5 = word-serial Montgomery
    // product without the final correction (fp_mul_lazy, what k_ntt_pass runs today).
    Fp<PP> b2 = fp_from_u64<PP>(seed * 7 + blockIdx.x + 3);
    Fp<PP> bq = fp_from_u64<PP>(seed * 9 + threadIdx.x + 5), b2q = fp_from_u64<PP>(seed * 11 + threadIdx.x + 9);
    for (u32 it = 0; it < iters; it++) {
      u32 t[8], u[8];
      if (VARIANT == 4) {
        shoup::mul_lazy<PP>(t, a.v, b.v, bq.v);
        shoup::mul_lazy<PP>(u, c.v, b2.v, b2q.v);
      } else {
        fp_mul_lazy_cios<PP>(t, b.v, a.v);
        fp_mul_lazy_cios<PP>(u, b2.v, c.v);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        a.v[i] = t[i];
        c.v[i] = u[i];
      }
      c.v[0] ^= a.v[7];  // the two chains stay in step, like a = f(a, b); c = f(c, a) below
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = fp_add(shoup::canon_3p<PP>(a.v), shoup::canon_3p<PP>(c.v));
  } else {
    for (u32 it = 0; it < iters; it++) {
      a = bench_product<PP, VARIANT>(a, b);
      c = bench_product<PP, VARIANT>(c, a);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = fp_add(a, c);
  }
}

template <class PP>
static void launch_modmul_peak(int variant, unsigned blocks, cudaStream_t st, Fp<PP>* out, u32 iters, u32 seed) {
  if (variant == 1) k_modmul_peak<PP, 1><<<blocks, 256, 0, st>>>(out, iters, seed);
  else if (variant == 2) k_modmul_peak<PP, 2><<<blocks, 256, 0, st>>>(out, iters, seed);
  else if (variant == 3) k_modmul_peak<PP, 3><<<blocks, 256, 0, st>>>(out, iters, seed);
  else if (variant == 4) k_modmul_peak<PP, 4><<<blocks, 256, 0, st>>>(out, iters, seed);
  else if (variant == 5) k_modmul_peak<PP, 5><<<blocks, 256, 0, st>>>(out, iters, seed);
  else k_modmul_peak<PP, 0><<<blocks, 256, 0, st>>>(out, iters, seed);
}

int bench_modmul(eon_ctx* ctx, int field, int variant, double* out_gmuls) {
  const unsigned blocks = (unsigned)ctx->num_sms * 8;
  const u32 iters = 512;
  void* buf = nullptr;
  EON_TRY(scratch_get(ctx, SC_SMALL, (size_t)blocks * 256 * 32, &buf));
  cudaEvent_t e0, e1;
  EON_CUDA(ctx, cudaEventCreate(&e0));
  EON_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    EON_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (field == 0) launch_modmul_peak<FrParams>(variant, blocks, ctx->stream, (Fr*)buf, iters, 99u + rep);
    else launch_modmul_peak<FqParams>(variant, blocks, ctx->stream, (Fq*)buf, iters, 99u + rep);
    EON_LAUNCHED(ctx);
    EON_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    EON_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    EON_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double muls = (double)blocks * 256.0 * iters * 2.0;
  *out_gmuls = muls / (best * 1e-3) / 1e9;
  return EON_OK;
}

}  // namespace eon
