// Probe: could the FP64 pipe carry part of the 256-bit modular products?  (DESIGN.md §7, round-2 candidates.)
// A 52 x 52 -> 104-bit partial product can be had from two DFMA.RZ and one DADD (Emmart et al.'s split):
//     hi = fma_rz(a, b, 2^104)                  mantissa = floor(a b / 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)    = 2^52 + (a b mod 2^52), exact
// and both halves are accumulated as 64-bit integers on their raw bit patterns.  One such term replaces
// (52/32)^2 = 2.64 32-bit limb products.  This probe measures, on all SMs:
//   0  DFMA peak (independent chains)
//   1  split terms alone            (2 DFMA + 1 DADD + 2 x 64-bit integer add per term)
//   2  IMAD alone                   (the library's integer roofline kernel, 16 chains of mad.lo)
//   3  split terms and IMAD chains interleaved in the same thread (do the pipes overlap?)
// Output: one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_split_probe fp64_split_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define C104 0x1p104
#define C104_52 (0x1p104 + 0x1p52)

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(unsigned long long* out, int iters, unsigned seed) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  // 52-bit integers held in doubles
  double a[8], b[8];
  unsigned long long acc_hi[8], acc_lo[8];
  unsigned ia[16];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = (double)((((unsigned long long)(seed + t) * 0x9E3779B97F4A7C15ull) >> 12) + i);
    b[i] = (double)((((unsigned long long)(seed ^ t) * 0xBF58476D1CE4E5B9ull) >> 12) + 3 * i);
    acc_hi[i] = acc_lo[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) ia[i] = seed + t * 16 + i;
  const unsigned m = seed | 1;
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = a[i];
  for (int it = 0; it < iters; it += 8) {
#pragma unroll
    for (int r = 0; r < 8; r++) {  // 8 rounds per trip: operand rotation by compile-time index, no register moves
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = __fma_rn(d[i], 1.0000001, b[i]);
      }
      if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const double x = a[(i + r) & 7];
          const double hi = __fma_rz(x, b[i], C104);
          const double lo = __fma_rz(x, b[i], C104_52 - hi);
          acc_hi[i] += (unsigned long long)__double_as_longlong(hi);
          acc_lo[i] += (unsigned long long)__double_as_longlong(lo);
        }
        // a data dependency on the accumulators, so that nothing is loop-invariant (1 LOP3 per 8 terms)
        a[r] = __hiloint2double(__double2hiint(a[r]), __double2loint(a[r]) ^ (int)((unsigned)acc_lo[r] & 1u));
      }
      if (MODE == 2 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 16; i++) ia[i] = ia[i] * m + ia[(i + 1) & 15];
      }
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= acc_hi[i] ^ (acc_lo[i] << 1) ^ (unsigned long long)__double_as_longlong(d[i]);
#pragma unroll
  for (int i = 0; i < 16; i++) r ^= ia[i];
  out[t] = r;
}

template <int MODE>
static float run(unsigned long long* buf, int blocks, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k_probe<MODE><<<blocks, 256>>>(buf, iters, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
    fprintf(stderr, "no CUDA device\n");
    return 1;
  }
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  unsigned long long* buf = nullptr;
  cudaMalloc(&buf, (size_t)blocks * 256 * sizeof(unsigned long long));
  const double threads = (double)blocks * 256.0;
  const float ms0 = run<0>(buf, blocks, iters), ms1 = run<1>(buf, blocks, iters), ms2 = run<2>(buf, blocks, iters),
              ms3 = run<3>(buf, blocks, iters);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  const double dfma_t = threads * iters * 8.0 / (ms0 * 1e-3) / 1e12;
  const double terms_g = threads * iters * 8.0 / (ms1 * 1e-3) / 1e9;
  const double imad_t = threads * iters * 16.0 / (ms2 * 1e-3) / 1e12;
  // a 256-bit Montgomery product in 5 x 52-bit limbs = 2 x 25 split terms (plus per-limb overhead not measured here)
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tops\": %.3f, \"split_terms_g_per_s\": %.2f, "
         "\"split_modmul_equiv_g_per_s\": %.2f, \"imad_tops\": %.3f, \"ms_dfma\": %.3f, \"ms_split\": %.3f, "
         "\"ms_imad\": %.3f, \"ms_split_and_imad_interleaved\": %.3f, \"overlap\": %.3f}\n",
         prop.name, prop.multiProcessorCount, dfma_t, terms_g, terms_g / 50.0, imad_t, ms0, ms1, ms2, ms3,
         (ms1 + ms2) / ms3);
  cudaFree(buf);
  return 0;
}
