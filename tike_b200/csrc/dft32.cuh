// Radix-32 butterflies in registers for the 512-point transforms of the
// large-detector pipeline (large_k2r.cu, large_k13r.cu): radix-2 with the constant
// twiddles w32^n, then two radix-16.  Slot p of the output holds frequency 2 p
// (p < 16) or 2 (p - 16) + 1.
#pragma once

#include "fft.cuh"

namespace tb {

__host__ __device__ constexpr float c32(int n) {
  constexpr float t[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f,
                           0.19509032201612833f, 0.0f, -0.19509032201612819f,
                           -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f,
                           -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
  return t[n];
}
__host__ __device__ constexpr float s32(int n) {
  constexpr float t[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                           0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f,
                           0.98078528040323043f, 1.0f, 0.98078528040323043f, 0.92387953251128674f,
                           0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f,
                           0.38268343236508989f, 0.19509032201612861f};
  return t[n];
}
// forward radix-32 (decimation in frequency), outputs in slot order
__host__ __device__ __forceinline__ void dft32(float2 (&x)[32]) {
  float2 a[16], b[16];
  auto tw = [&](auto N_) {
    constexpr int n = decltype(N_)::value;
    const float2 d = csub(x[n], x[n + 16]);
    a[n] = cadd(x[n], x[n + 16]);
    // d * (c - i s)
    b[n] = make_float2(d.x * c32(n) + d.y * s32(n), d.y * c32(n) - d.x * s32(n));
  };
  tw(std::integral_constant<int, 0>{}); tw(std::integral_constant<int, 1>{});
  tw(std::integral_constant<int, 2>{}); tw(std::integral_constant<int, 3>{});
  tw(std::integral_constant<int, 4>{}); tw(std::integral_constant<int, 5>{});
  tw(std::integral_constant<int, 6>{}); tw(std::integral_constant<int, 7>{});
  tw(std::integral_constant<int, 8>{}); tw(std::integral_constant<int, 9>{});
  tw(std::integral_constant<int, 10>{}); tw(std::integral_constant<int, 11>{});
  tw(std::integral_constant<int, 12>{}); tw(std::integral_constant<int, 13>{});
  tw(std::integral_constant<int, 14>{}); tw(std::integral_constant<int, 15>{});
  dft<16>(a);
  dft<16>(b);
#pragma unroll
  for (int n = 0; n < 16; ++n) { x[n] = a[n]; x[n + 16] = b[n]; }
}
// unscaled inverse of dft32: slot order in, natural order out
__host__ __device__ __forceinline__ void idft32(float2 (&x)[32]) {
  float2 a[16], b[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) { a[n] = x[n]; b[n] = x[n + 16]; }
  idft<16>(a);
  idft<16>(b);
  auto tw = [&](auto N_) {
    constexpr int n = decltype(N_)::value;
    // b * (c + i s)
    const float2 d = make_float2(b[n].x * c32(n) - b[n].y * s32(n), b[n].y * c32(n) + b[n].x * s32(n));
    x[n] = cadd(a[n], d);
    x[n + 16] = csub(a[n], d);
  };
  tw(std::integral_constant<int, 0>{}); tw(std::integral_constant<int, 1>{});
  tw(std::integral_constant<int, 2>{}); tw(std::integral_constant<int, 3>{});
  tw(std::integral_constant<int, 4>{}); tw(std::integral_constant<int, 5>{});
  tw(std::integral_constant<int, 6>{}); tw(std::integral_constant<int, 7>{});
  tw(std::integral_constant<int, 8>{}); tw(std::integral_constant<int, 9>{});
  tw(std::integral_constant<int, 10>{}); tw(std::integral_constant<int, 11>{});
  tw(std::integral_constant<int, 12>{}); tw(std::integral_constant<int, 13>{});
  tw(std::integral_constant<int, 14>{}); tw(std::integral_constant<int, 15>{});
}
// frequency held by slot p of dft32's output
__host__ __device__ constexpr int dft32_freq(int p) { return p < 16 ? 2 * p : 2 * (p - 16) + 1; }

}  // namespace tb
