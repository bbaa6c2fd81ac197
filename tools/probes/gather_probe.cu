// Probe: DRAM bytes fetched per random 32-byte / 64-byte gather from a table much larger than L2,
// for different load flavours.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
// Run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) V8 { uint32_t v[8]; };

template <int MODE>
__device__ __forceinline__ V8 load32(const V8* p) {
  V8 r;
  if (MODE == 0) {  // ld.global.nc 256-bit
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  } else if (MODE == 1) {  // plain ld.global 256-bit
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  } else if (MODE == 2) {  // two 128-bit .cg loads
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldcg(q), b = __ldcg(q + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else if (MODE == 3) {  // two 128-bit __ldg
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else if (MODE == 4) {  // 256-bit with evict-first / no-allocate style hint
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  } else {  // 256-bit, L2 64-byte prefetch hint
    asm volatile("ld.global.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  }
  return r;
}

// every thread gathers `per` random 32-byte elements (element stride `stride32` x 32 bytes apart in index space)
template <int MODE>
__global__ void k_gather(const V8* __restrict__ tab, uint64_t nelem, int per, uint32_t* out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t s = t * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int i = 0; i < per; i++) {
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    V8 v = load32<MODE>(tab + (s % nelem));
    acc += v.v[0] ^ v.v[7];
  }
  out[t] = acc;
}

int main() {
  const uint64_t bytes = 4ull << 30;           // 4 GiB table
  const uint64_t nelem = bytes / 32;
  V8* tab; uint32_t* out;
  cudaMalloc(&tab, bytes); cudaMemset(tab, 1, bytes);
  const int threads = 256, blocks = 148 * 64, per = 64;
  cudaMalloc(&out, (size_t)threads * blocks * 4);
  printf("loads per launch: %llu (x 32 B = %.2f GB useful)\n", (unsigned long long)threads * blocks * per,
         (double)threads * blocks * per * 32 / 1e9);
  k_gather<0><<<blocks, threads>>>(tab, nelem, per, out);
  k_gather<1><<<blocks, threads>>>(tab, nelem, per, out);
  k_gather<2><<<blocks, threads>>>(tab, nelem, per, out);
  k_gather<3><<<blocks, threads>>>(tab, nelem, per, out);
  k_gather<4><<<blocks, threads>>>(tab, nelem, per, out);
  k_gather<5><<<blocks, threads>>>(tab, nelem, per, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
