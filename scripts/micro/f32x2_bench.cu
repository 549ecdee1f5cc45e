// Microbenchmark: throughput of packed f32x2 adds / fmas against scalar ones on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float addv(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmav(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void k(float2* out, int iters) {
  float2 acc[8];
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x + j, j);
  float2 inc = make_float2(1.0f + threadIdx.x * 1e-6f, 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) { acc[j].x = addv(acc[j].x, inc.x); acc[j].y = addv(acc[j].y, inc.y); }
      if (MODE == 1) { u64 r = add2(*(u64*)&acc[j], *(u64*)&inc); acc[j] = *(float2*)&r; }
      if (MODE == 2) { acc[j].x = fmav(acc[j].x, inc.x, inc.y); acc[j].y = fmav(acc[j].y, inc.y, inc.x); }
      if (MODE == 3) { u64 r = fma2(*(u64*)&acc[j], *(u64*)&inc, *(u64*)&inc); acc[j] = *(float2*)&r; }
    }
  }
  float2 s = make_float2(0, 0);
  for (int j = 0; j < 8; ++j) { s.x += acc[j].x; s.y += acc[j].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
  float2* out; cudaMalloc(&out, 148 * 4 * 512 * sizeof(float2));
  const int iters = 1 << 14;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 512>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 4, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cplx_ops = 148.0 * 4 * 512 * iters * 8;
  printf("%-12s %.3f ms  %.2f T complex-ops/s\n", name, ms, cplx_ops / ms * 1e-9);
  cudaFree(out);
}
int main() { run<0>("FADD x2"); run<1>("FADD2"); run<2>("FFMA x2"); run<3>("FFMA2"); return 0; }
