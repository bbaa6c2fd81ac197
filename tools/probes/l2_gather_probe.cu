// Probe: random 32-byte and 64-byte gather rate as a function of the table size (L2-resident vs DRAM).
// Sizes the "one window table per phase" layout of the MSM round-0 gathers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_probe l2_gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) V8 { uint32_t v[8]; };

__device__ __forceinline__ V8 load32(const V8* p) {
  V8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}

// BOTH: gather the 64-byte point (two adjacent 32-byte halves), else only its first half
template <bool BOTH>
__global__ void k_gather(const V8* __restrict__ tab, uint64_t npoints, int per, uint32_t* out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t s = t * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int i = 0; i < per; i++) {
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    const V8* p = tab + 2 * (s % npoints);
    V8 v = load32(p);
    acc += v.v[0] ^ v.v[7];
    if (BOTH) { V8 w = load32(p + 1); acc += w.v[3]; }
  }
  out[t] = acc;
}

int main() {
  const int threads = 256, blocks = 148 * 32, per = 64;
  uint32_t* out;
  cudaMalloc(&out, (size_t)threads * blocks * 4);
  const uint64_t maxb = 2048ull << 20;
  V8* tab; cudaMalloc(&tab, maxb); cudaMemset(tab, 1, maxb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gathers = (double)threads * blocks * per;
  const int mibs[] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 128, 192, 256, 512, 1024, 2048};
  printf("table_MiB  x_only_Ggather/s  xy_Ggather/s\n");
  for (int mi : mibs) {
    uint64_t npoints = ((uint64_t)mi << 20) / 64;
    float ms[2];
    for (int both = 0; both < 2; both++) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        if (both) k_gather<true><<<blocks, threads>>>(tab, npoints, per, out);
        else k_gather<false><<<blocks, threads>>>(tab, npoints, per, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < best) best = t;
      }
      ms[both] = best;
    }
    printf("%8d  %16.1f  %12.1f\n", mi, gathers / ms[0] / 1e6, gathers / ms[1] / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
