// In-shared-memory mixed-radix complex FFT building blocks (sm_100a).
//
// Replaces the cuFFT C2C calls behind the reference's Propagation operator
// (src/tike/operators/cupy/propagation.py:43-73, cache.py:32-82).
//
// Design:
//  * N = R_1 * R_2 * ... * R_s with R_i in {2,4,8,16}.  Forward transforms are
//    decimation-in-frequency: every stage is IN PLACE (each thread loads R
//    elements, does a radix-R DFT in registers, multiplies the stage twiddles
//    and stores to the same R locations), so no stage needs a read-all /
//    write-all barrier pair, only one __syncthreads() between stages.
//  * The forward output is left in digit-reversed order: location
//    l = k_1*S_1 + k_2*S_2 + ... + k_s holds frequency
//    k = k_1 + R_1*k_2 + R_1*R_2*k_3 ...  (S_i = R_{i+1}*...*R_s).
//    Fused kernels index the measured data through loc2freq[] instead of
//    reordering; the inverse transform (decimation-in-time, stages in reverse
//    order, conjugate twiddles first) consumes digit-reversed input and
//    produces natural order.
//  * A 2-D transform runs the 1-D passes over rows (vectors = rows) and then
//    over columns (vectors = columns).  Lanes of a warp always walk over
//    *vectors*: for rows the bank step is the row pitch (ND+1 complex, odd =>
//    conflict free for 64-bit accesses), for columns it is one element.
//  * Inverse radix butterflies are the forward ones with conjugated constants
//    (dft<R, true>).
#pragma once

#include "common.cuh"

namespace tb {

// Thread coordinates: real ones on the device; (0, 1) when the header is
// compiled for the host-side unit test (tests/csrc/fft_host_test.cu).
__host__ __device__ __forceinline__ int tb_tid() {
#ifdef __CUDA_ARCH__
  return threadIdx.x;
#else
  return 0;
#endif
}
__host__ __device__ __forceinline__ int tb_nthreads() {
#ifdef __CUDA_ARCH__
  return blockDim.x;
#else
  return 1;
#endif
}
__host__ __device__ __forceinline__ void tb_sync() {
#ifdef __CUDA_ARCH__
  __syncthreads();
#endif
}

// cos / sin of 2*pi*j/16 (forward twiddle w16^j = cos - i sin)
__host__ __device__ constexpr float cos16(int j) {
  constexpr float t[16] = {1.0f,
                           0.92387953251128674f,
                           0.70710678118654752f,
                           0.38268343236508977f,
                           0.0f,
                           -0.38268343236508977f,
                           -0.70710678118654752f,
                           -0.92387953251128674f,
                           -1.0f,
                           -0.92387953251128674f,
                           -0.70710678118654752f,
                           -0.38268343236508977f,
                           0.0f,
                           0.38268343236508977f,
                           0.70710678118654752f,
                           0.92387953251128674f};
  return t[j & 15];
}
__host__ __device__ constexpr float sin16(int j) { return cos16(j - 4); }

// Inverse butterflies: native conjugate-twiddle form (default) or the
// swap(re,im) . DFT . swap(re,im) identity (-DTB_EXP_SWAP_INVERSE=1).  The
// swaps break the (re, im) register pairing that packed FADD2 needs, which
// costs one MOV per element in the inverse stages (4 % of the fused rPIE
// kernel's instructions, profiles/r01s_*).
#ifndef TB_EXP_SWAP_INVERSE
#define TB_EXP_SWAP_INVERSE 0
#endif

// x *= w_R^J (forward sign; INV: conjugate), J and R compile-time; trivial
// cases are free.
template <int J, int R, bool INV = false>
__host__ __device__ __forceinline__ float2 mul_w(float2 v) {
  constexpr int j16 = (J * (16 / R)) & 15;
  if constexpr (j16 == 0) {
    return v;
  } else if constexpr (j16 == 4) {  // -i (INV: +i)
    return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
  } else if constexpr (j16 == 8) {  // -1
    return make_float2(-v.x, -v.y);
  } else if constexpr (j16 == 12) {  // +i (INV: -i)
    return INV ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
  } else {
    constexpr float c = cos16(j16), s = INV ? -sin16(j16) : sin16(j16);
    // (x + iy)(c - is)
    return make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
  }
}

template <int R, bool INV = false>
__host__ __device__ __forceinline__ void dft(float2 (&x)[R]);

// R = A*B Cooley-Tukey in registers: n = B*na + nb, k = ka + A*kb.
template <int R, int A, int B, bool INV>
__host__ __device__ __forceinline__ void dft_composite(float2 (&x)[R]) {
  float2 y[B][A];
#pragma unroll
  for (int nb = 0; nb < B; ++nb) {
    float2 t[A];
#pragma unroll
    for (int na = 0; na < A; ++na) t[na] = x[B * na + nb];
    dft<A, INV>(t);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) y[nb][ka] = t[ka];
  }
  // twiddles w_R^(nb*ka): fully unrolled with compile-time exponents
  auto tw = [&](auto NB, auto KA) {
    constexpr int nb = decltype(NB)::value, ka = decltype(KA)::value;
    y[nb][ka] = mul_w<nb * ka, R, INV>(y[nb][ka]);
  };
  auto for_ka = [&](auto NB) {
    if constexpr (A > 1) tw(NB, std::integral_constant<int, 1>{});
    if constexpr (A > 2) tw(NB, std::integral_constant<int, 2>{});
    if constexpr (A > 3) tw(NB, std::integral_constant<int, 3>{});
  };
  if constexpr (B > 1) for_ka(std::integral_constant<int, 1>{});
  if constexpr (B > 2) for_ka(std::integral_constant<int, 2>{});
  if constexpr (B > 3) for_ka(std::integral_constant<int, 3>{});
#pragma unroll
  for (int ka = 0; ka < A; ++ka) {
    float2 t[B];
#pragma unroll
    for (int nb = 0; nb < B; ++nb) t[nb] = y[nb][ka];
    dft<B, INV>(t);
#pragma unroll
    for (int kb = 0; kb < B; ++kb) x[ka + A * kb] = t[kb];
  }
}

template <int R, bool INV>
__host__ __device__ __forceinline__ void dft(float2 (&x)[R]) {
  static_assert(R == 2 || R == 4 || R == 8 || R == 16, "radix");
  if constexpr (R == 2) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  } else if constexpr (R == 4) {
    const float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    const float2 t2 = cadd(x[1], x[3]);
    const float2 t3 = make_float2(x[1].x - x[3].x, x[1].y - x[3].y);
    x[0] = cadd(t0, t2);
    x[2] = csub(t0, t2);
    // forward: X1 = t1 - i t3, X3 = t1 + i t3 (inverse: the other way round),
    // with -i t3 formed by two scalar ops so that both outputs are packed
    // add / sub again
    const float2 r3 = make_float2(t3.y, -t3.x);
    x[1] = INV ? csub(t1, r3) : cadd(t1, r3);
    x[3] = INV ? cadd(t1, r3) : csub(t1, r3);
  } else if constexpr (R == 8) {
    dft_composite<8, 4, 2, INV>(x);
  } else {
    dft_composite<16, 4, 4, INV>(x);
  }
}

// unscaled inverse radix-R DFT
template <int R>
__host__ __device__ __forceinline__ void idft(float2 (&x)[R]) {
#if TB_EXP_SWAP_INVERSE
#pragma unroll
  for (int k = 0; k < R; ++k) x[k] = make_float2(x[k].y, x[k].x);
  dft<R>(x);
#pragma unroll
  for (int k = 0; k < R; ++k) x[k] = make_float2(x[k].y, x[k].x);
#else
  dft<R, true>(x);
#endif
}

// ---------------------------------------------------------------------------
// Plans: radices per size.
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int plan_radix(int n, int i) {
  switch (n) {
    case 16:   return i == 0 ? 16 : 1;
    case 32:   return i == 0 ? 8 : (i == 1 ? 4 : 1);
    case 64:   return i < 2 ? 8 : 1;
    case 128:  return i == 0 ? 8 : (i == 1 ? 16 : 1);
    case 256:  return i < 2 ? 16 : 1;
    case 512:  return 8;
    case 1024: return i == 0 ? 16 : 8;
    case 2048: return i < 2 ? 16 : 8;
    default:   return 1;
  }
}
template <int N> struct Plan {
  static constexpr int NS = (N == 16) ? 1 : (N <= 256 ? 2 : 3);
  __host__ __device__ static constexpr int r(int i) { return plan_radix(N, i); }
};

// frequency index held at digit-reversed location l after the forward pass
template <int N>
__host__ __device__ inline int loc2freq(int l) {
  int k = 0, mult = 1, L = N;
#pragma unroll
  for (int i = 0; i < Plan<N>::NS; ++i) {
    const int R = Plan<N>::r(i);
    const int S = L / R;
    const int ki = l / S;
    l -= ki * S;
    k += ki * mult;
    mult *= R;
    L = S;
  }
  return k;
}
template <int N>
__host__ __device__ inline int freq2loc(int k) {
  int l = 0, L = N;
#pragma unroll
  for (int i = 0; i < Plan<N>::NS; ++i) {
    const int R = Plan<N>::r(i);
    const int S = L / R;
    l += (k % R) * S;
    k /= R;
    L = S;
  }
  return l;
}

// forward twiddle table tw[n] = exp(-2 pi i n / N), n in [0, N)
template <int N>
__host__ __device__ __forceinline__ void fill_twiddles(float2* tw) {
  for (int n = tb_tid(); n < N; n += tb_nthreads()) {
    float s, c;
    sincospif(2.0f * (float)n / (float)N, &s, &c);
    tw[n] = make_float2(c, -s);
  }
}

// One radix-R stage (sub-FFT length L) over 2^LOGNVEC vectors.  Element i of
// vector v lives at s[v * VSTRIDE + i * ESTRIDE] (all compile-time, so the
// address arithmetic folds into immediates).  No barrier inside.
template <int N, int R, int L, bool INV, int LOGNVEC, int VSTRIDE, int ESTRIDE>
__host__ __device__ __forceinline__ void fft_stage(float2* __restrict__ s,
                                                   const float2* __restrict__ tw) {
  constexpr int S = L / R;        // element stride inside a sub-FFT
  constexpr int BF = N / R;       // butterflies per vector
  constexpr int TWS = N / L;      // twiddle table stride for w_L
  constexpr int total = BF << LOGNVEC;
  constexpr int vmask = (1 << LOGNVEC) - 1;
  for (int b = tb_tid(); b < total; b += tb_nthreads()) {
    const int v = b & vmask;
    const int j = b >> LOGNVEC;
    const int blk = j / S, n2 = j - blk * S;
    float2* p = s + v * VSTRIDE + (blk * L + n2) * ESTRIDE;
    float2 x[R];
#pragma unroll
    for (int k = 0; k < R; ++k) x[k] = p[k * S * ESTRIDE];
    if constexpr (!INV) {
      dft<R>(x);
      if constexpr (S > 1) {
#pragma unroll
        for (int k = 1; k < R; ++k) x[k] = cmul(x[k], tw[n2 * k * TWS]);
      }
    } else {
      if constexpr (S > 1) {
#pragma unroll
        for (int k = 1; k < R; ++k) x[k] = cmulc(tw[n2 * k * TWS], x[k]);
      }
      idft<R>(x);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) p[k * S * ESTRIDE] = x[k];
  }
}

// All stages of a length-N 1-D transform over 2^LOGNVEC vectors.  Forward:
// natural in, digit-reversed out.  Inverse: digit-reversed in, natural out.
// Unscaled.  A barrier follows every stage (including the last).
template <int N, bool INV, int LOGNVEC, int VSTRIDE, int ESTRIDE>
__host__ __device__ __forceinline__ void fft_pass(float2* s, const float2* tw) {
  using P = Plan<N>;
  constexpr int R0 = P::r(0), R1 = P::r(1), R2 = P::r(2);
  constexpr int L0 = N, L1 = N / R0, L2 = N / (R0 * R1);
  if constexpr (!INV) {
    fft_stage<N, R0, L0, false, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
    tb_sync();
    if constexpr (P::NS > 1) {
      fft_stage<N, R1, L1, false, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
      tb_sync();
    }
    if constexpr (P::NS > 2) {
      fft_stage<N, R2, L2, false, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
      tb_sync();
    }
  } else {
    if constexpr (P::NS > 2) {
      fft_stage<N, R2, L2, true, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
      tb_sync();
    }
    if constexpr (P::NS > 1) {
      fft_stage<N, R1, L1, true, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
      tb_sync();
    }
    fft_stage<N, R0, L0, true, LOGNVEC, VSTRIDE, ESTRIDE>(s, tw);
    tb_sync();
  }
}

template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// 2-D transform of an N x N tile with row pitch N+1 held in shared memory.
// Forward: natural -> digit-reversed in both axes.  Inverse: the opposite.
// Caller must have synchronised the tile before the call; tile is
// synchronised on return.
template <int N, bool INV>
__host__ __device__ __forceinline__ void fft2_tile(float2* s, const float2* tw) {
  constexpr int PITCH = N + 1;
  constexpr int LG = Log2<N>::v;
  if constexpr (!INV) {
    fft_pass<N, false, LG, PITCH, 1>(s, tw);   // rows
    fft_pass<N, false, LG, 1, PITCH>(s, tw);   // columns
  } else {
    fft_pass<N, true, LG, 1, PITCH>(s, tw);    // columns
    fft_pass<N, true, LG, PITCH, 1>(s, tw);    // rows
  }
}

}  // namespace tb
