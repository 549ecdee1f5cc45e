// Shared device/host helpers for libtikeb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <type_traits>

#include "../../include/tike_b200.h"

namespace tb {

// ---------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative/cuda code and
// leaves a thread-local message for tb_last_error().
// ---------------------------------------------------------------------------
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

#define TB_REQUIRE(cond, code, ...)                       \
  do {                                                    \
    if (!(cond)) return ::tb::set_error((code), __VA_ARGS__); \
  } while (0)


// ---------------------------------------------------------------------------
// complex arithmetic on float2
// ---------------------------------------------------------------------------
// On the device a complex add / subtract is ONE packed instruction
// (add.f32x2 -> FADD2 on sm_100a).  FADD2 has the FLOP rate of two FADDs but
// takes a single issue slot, and the radix butterflies are mostly add/sub.
#if defined(__CUDA_ARCH__) && !defined(TB_NO_PACKED_F32X2)
__device__ __forceinline__ unsigned long long tb_pack(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 tb_unpack(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(tb_pack(a)), "l"(tb_pack(b)));
  return tb_unpack(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(tb_pack(a)), "l"(tb_pack(b)));
  return tb_unpack(r);
}
#else
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  return make_float2(a.x + b.x, a.y + b.y);
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
  return make_float2(a.x - b.x, a.y - b.y);
}
#endif
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) {
  return make_float2(a.x * s, a.y * s);
}
__host__ __device__ __forceinline__ float cabs2(float2 a) { return a.x * a.x + a.y * a.y; }

// fire-and-forget vector reduction into global memory (complex64 accumulate)
__device__ __forceinline__ void red_add_f32x2(float2* addr, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x),
               "f"(v.y)
               : "memory");
}
// 16-byte variant: two adjacent complex64 values in one reduction
__device__ __forceinline__ void red_add_f32x4(float2* addr, float2 a, float2 b) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a.x),
               "f"(a.y), "f"(b.x), "f"(b.y)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// L2 residency hints.  The spilled far-field waves are written once and read
// back within the same position: keep them (evict_last) ahead of the streaming
// diffraction data (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_f32x4_hint(float4* addr, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 ld_f32x4_hint(const float4* addr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(addr), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_f32_hint(const float* addr, uint64_t pol) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(addr), "l"(pol));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* addr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of up to K floats per thread; result valid in ALL threads.
// scratch must hold K * 32 floats.  Contains __syncthreads().
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float x = (lane < nwarp) ? scratch[k * 32 + lane] : 0.f;
    v[k] = warp_sum(x);
  }
}

// Bilinear geometry of one scan position (reference: convolution.cu:102-133).
struct Corner {
  int iy, ix;         // floor(scan)
  float w00, w01, w10, w11;  // (1-fx)(1-fy), fx(1-fy), (1-fx)fy, fx*fy
};
__device__ __forceinline__ Corner make_corner(const float* scan, long s) {
  const float sy = scan[2 * s], sx = scan[2 * s + 1];
  const float fy0 = floorf(sy), fx0 = floorf(sx);
  const float fy = sy - fy0, fx = sx - fx0;
  Corner c;
  c.iy = (int)fy0;
  c.ix = (int)fx0;
  c.w00 = (1.0f - fx) * (1.0f - fy);
  c.w01 = fx * (1.0f - fy);
  c.w10 = (1.0f - fx) * fy;
  c.w11 = fx * fy;
  return c;
}

// Interpolated object value at patch pixel (py, px); out-of-range neighbours
// contribute zero (the reference reads them with zero weight).
__device__ __forceinline__ float2 patch_value(const float2* __restrict__ img,
                                              int H, int W, const Corner& c,
                                              int py, int px) {
  const int y = c.iy + py, x = c.ix + px;
  float2 r = make_float2(0.f, 0.f);
  const bool y0 = (y >= 0) & (y < H), y1 = (y + 1 >= 0) & (y + 1 < H);
  const bool x0 = (x >= 0) & (x < W), x1 = (x + 1 >= 0) & (x + 1 < W);
  const float2* p = img + (long)y * W + x;
  if (y0 & x0) { float2 v = __ldg(p);         r.x = v.x * c.w00;  r.y = v.y * c.w00; }
  if (y0 & x1) { float2 v = __ldg(p + 1);     r.x += v.x * c.w01; r.y += v.y * c.w01; }
  if (y1 & x0) { float2 v = __ldg(p + W);     r.x += v.x * c.w10; r.y += v.y * c.w10; }
  if (y1 & x1) { float2 v = __ldg(p + W + 1); r.x += v.x * c.w11; r.y += v.y * c.w11; }
  return r;
}

}  // namespace tb
