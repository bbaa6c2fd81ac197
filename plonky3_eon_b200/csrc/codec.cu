// Compressed G1 encoding on the device: batched compression of commitments / proofs and batched
// decompression of a serialised SRS (sm_100a).
//
// Replaces, for whole arrays of points at once,
//   G1::to_bytes                           bn254/src/curve.rs:136-139   (-> halo2curves G1Affine::to_bytes)
//   Serialize / Deserialize for G1         bn254/src/curve.rs:84-98     (StructuredReferenceString serde,
//                                                                        kzg/src/params.rs:56-77)
// The 32-byte form is halo2curves' GroupEncoding of bn256::G1Affine.  halo2curves ("0.9", bn254/Cargo.toml:22)
// is not vendored in the reference, so the layout is restated from its published definition:
//   EON_G1_ENC_HALO2  (0.4 and later): x canonical little-endian in bytes 0..31 (x < q < 2^254), byte 31 bit 6 =
//                     sign = lowest bit of canonical y, byte 31 bit 7 = identity (all other bits zero)
//   EON_G1_ENC_LEGACY (0.3 and earlier): sign in byte 31 bit 7, identity = 32 zero bytes
// Decompression: y = (x^3 + 3)^((q+1)/4) (q = 3 mod 4), rejected unless y^2 = x^3 + 3 and x < q; the root
// whose parity matches the sign bit is taken.  G1 has cofactor 1, so on-curve means in the group.
#include "common.cuh"

namespace eon {

constexpr int CODEC_THREADS = 128;
constexpr unsigned long long NO_BAD = ~0ull;

__device__ __forceinline__ Fq fq_three() {
  const Fq one = Fq::one();
  return fp_add(fp_add(one, one), one);
}

// x^((q+1)/4), the candidate square root for q = 3 mod 4
__device__ Fq fq_sqrt_candidate(const Fq& a) {
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = FqParams::mod(i);
  e[0] += 1;  // q ends in ...47: no carry
#pragma unroll
  for (int i = 0; i < 7; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
  e[7] >>= 2;
  Fq r = Fq::one();
#pragma unroll 1
  for (int i = 253; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
  }
  return r;
}

__global__ void __launch_bounds__(CODEC_THREADS)
k_g1_compress(const G1Affine* __restrict__ pts, u64 n, u32* __restrict__ out, int enc) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G1Affine p = pts[i];
  u32 w[8];
  if (p.is_identity()) {
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = 0;
    if (enc == EON_G1_ENC_HALO2) w[7] = 0x80000000u;
  } else {
    u32 y[8];
    fp_from_mont<FqParams>(w, p.x);
    fp_from_mont<FqParams>(y, p.y);
    w[7] |= (y[0] & 1u) << (enc == EON_G1_ENC_HALO2 ? 30 : 31);
  }
  uint4* o = reinterpret_cast<uint4*>(out + 8 * i);
  o[0] = make_uint4(w[0], w[1], w[2], w[3]);
  o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__global__ void __launch_bounds__(CODEC_THREADS)
k_g1_decompress(const u32* __restrict__ in, u64 n, G1Affine* __restrict__ out, int enc,
                unsigned long long* __restrict__ first_bad) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(in + 8 * i);
  const uint4 lo = src[0], hi = src[1];
  u32 w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  u32 sign, inf;
  if (enc == EON_G1_ENC_HALO2) {
    inf = w[7] >> 31;
    sign = (w[7] >> 30) & 1u;
    w[7] &= 0x3fffffffu;
  } else {
    inf = 0;
    sign = w[7] >> 31;
    w[7] &= 0x7fffffffu;
  }
  u32 any = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) any |= w[k];
  G1Affine r = G1Affine::identity();
  bool bad = false;
  if (inf) {
    bad = (any | sign) != 0;  // the identity carries no other bit
  } else if (enc == EON_G1_ENC_LEGACY && any == 0 && sign == 0) {
    // 32 zero bytes: the identity
  } else {
    // x must be a canonical field element
    bool lt = false;
#pragma unroll
    for (int k = 7; k >= 0; k--) {
      const u32 m = FqParams::mod(k);
      if (w[k] != m) { lt = w[k] < m; break; }
    }
    if (!lt) {
      bad = true;
    } else {
      const Fq x = fp_to_mont<FqParams>(w);
      const Fq rhs = fp_add(fp_mul(fp_sqr(x), x), fq_three());
      Fq y = fq_sqrt_candidate(rhs);
      if (fp_sqr(y) != rhs) {
        bad = true;
      } else {
        u32 yc[8];
        fp_from_mont<FqParams>(yc, y);
        if ((yc[0] & 1u) != sign) y = fp_neg(y);
        // y = 0 has one root only; a set sign bit cannot be honoured (no such point on this curve: 3 | order
        // would be needed, and the group order is prime) -- treat it like halo2curves: accept the root found
        r.x = x;
        r.y = y;
      }
    }
  }
  if (bad) atomicMin(first_bad, (unsigned long long)i);
  out[i] = r;
}

static int check_enc(eon_ctx* ctx, int enc) {
  if (enc != EON_G1_ENC_HALO2 && enc != EON_G1_ENC_LEGACY) return fail(ctx, EON_ERR_BAD_ARG, "unknown G1 encoding");
  return EON_OK;
}

// d_pts (device) -> h_out (host, 32 n bytes)
int g1_compress_run(eon_ctx* ctx, const G1Affine* d_pts, size_t n, uint8_t* h_out, int enc) {
  EON_TRY(check_enc(ctx, enc));
  if (n == 0) return EON_OK;
  void* d_out;
  EON_TRY(scratch_get(ctx, SC_IO_B, n * 32, &d_out));
  k_g1_compress<<<(unsigned)((n + CODEC_THREADS - 1) / CODEC_THREADS), CODEC_THREADS, 0, ctx->stream>>>(
      d_pts, n, (u32*)d_out, enc);
  EON_LAUNCHED(ctx);
  EON_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return EON_OK;
}

// h_in (host, 32 n bytes) -> d_out (device, n affine points).  *bad_index = first invalid encoding or SIZE_MAX.
int g1_decompress_run(eon_ctx* ctx, const uint8_t* h_in, size_t n, G1Affine* d_out, int enc, size_t* bad_index) {
  EON_TRY(check_enc(ctx, enc));
  if (bad_index) *bad_index = (size_t)-1;
  if (n == 0) return EON_OK;
  void *d_in, *d_flag;
  EON_TRY(scratch_get(ctx, SC_IO_B, n * 32, &d_in));
  EON_TRY(scratch_get(ctx, SC_SMALL, sizeof(unsigned long long), &d_flag));
  EON_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  EON_CUDA(ctx, cudaMemsetAsync(d_flag, 0xff, sizeof(unsigned long long), ctx->stream));
  k_g1_decompress<<<(unsigned)((n + CODEC_THREADS - 1) / CODEC_THREADS), CODEC_THREADS, 0, ctx->stream>>>(
      (const u32*)d_in, n, d_out, enc, (unsigned long long*)d_flag);
  EON_LAUNCHED(ctx);
  unsigned long long bad = NO_BAD;
  EON_CUDA(ctx, cudaMemcpyAsync(&bad, d_flag, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  EON_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad != NO_BAD) {
    if (bad_index) *bad_index = (size_t)bad;
    char b[96];
    snprintf(b, sizeof(b), "Invalid G1 point at index %llu", bad);  // bn254/src/curve.rs:95
    return fail(ctx, EON_ERR_BAD_POINT, b);
  }
  return EON_OK;
}

}  // namespace eon
