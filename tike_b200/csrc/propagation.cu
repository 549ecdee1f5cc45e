// Propagation operator: batched in-place 2-D complex FFT, natural order in
// and out.  Replaces cupyx.scipy.fft.fftn/ifftn through CachedFFT
// (src/tike/operators/cupy/propagation.py:43-73, cache.py:32-82).
//
//  n <= 128 : one CTA per image, whole image resident in shared memory.
//  n >= 256 : two passes through HBM — column transforms on (n x C) tiles,
//             then row transforms on (T x n) tiles (the "two-pass row/column
//             FFT" for detectors that exceed shared memory).
#include "../../include/tike_b200.h"
#include "fft.cuh"

namespace tb {

template <int N>
__global__ void __launch_bounds__((N >= 128) ? 512 : (N >= 64 ? 256 : 128))
fft2_smem_kernel(float2* __restrict__ x, long batch, int inverse, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + N * (N + 1);
  unsigned short* f2l = reinterpret_cast<unsigned short*>(tw + N);
  fill_twiddles<N>(tw);
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    f2l[i] = (unsigned short)freq2loc<N>(i);
  __syncthreads();
  for (long b = blockIdx.x; b < batch; b += gridDim.x) {
    float2* img = x + b * (long)N * N;
    if (!inverse) {
      for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int r = idx / N, c = idx - r * N;
        tile[r * (N + 1) + c] = img[idx];
      }
      __syncthreads();
      fft2_tile<N, false>(tile, tw);
      for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int ky = idx / N, kx = idx - ky * N;
        img[idx] = cscale(tile[f2l[ky] * (N + 1) + f2l[kx]], scale);
      }
    } else {
      for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int ky = idx / N, kx = idx - ky * N;
        tile[f2l[ky] * (N + 1) + f2l[kx]] = img[idx];
      }
      __syncthreads();
      fft2_tile<N, true>(tile, tw);
      for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int r = idx / N, c = idx - r * N;
        img[idx] = cscale(tile[r * (N + 1) + c], scale);
      }
    }
    __syncthreads();
  }
}

// One axis of a large 2-D transform.  AXIS 0: columns (tile N rows x C cols),
// AXIS 1: rows (tile T rows x N cols).  V = vectors per tile (C or T).
// Forward writes natural frequency order along the transformed axis; inverse
// reads natural order.
template <int N, int V, int AXIS>
__global__ void __launch_bounds__(512)
fft_axis_kernel(float2* __restrict__ x, long batch, int inverse, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  constexpr int TILE = (AXIS == 0) ? N * (V + 1) : V * (N + 1);
  float2* tw = tile + TILE;
  unsigned short* f2l = reinterpret_cast<unsigned short*>(tw + N);
  fill_twiddles<N>(tw);
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    f2l[i] = (unsigned short)freq2loc<N>(i);
  __syncthreads();
  constexpr int TILES_PER_IMAGE = N / V;
  const long ntiles = batch * TILES_PER_IMAGE;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long b = t / TILES_PER_IMAGE;
    const int v0 = (int)(t - b * TILES_PER_IMAGE) * V;
    float2* img = x + b * (long)N * N;
    if (AXIS == 0) {
      // tile[row][c], pitch V+1; vectors = columns (vstride 1, estride V+1)
      for (int idx = threadIdx.x; idx < N * V; idx += blockDim.x) {
        const int r = idx / V, c = idx - r * V;
        const int rs = inverse ? (int)f2l[r] : r;  // inverse: natural -> digit-reversed slot
        tile[rs * (V + 1) + c] = img[(long)r * N + v0 + c];
      }
      __syncthreads();
      if (inverse) fft_pass<N, true, Log2<V>::v, 1, V + 1>(tile, tw);
      else         fft_pass<N, false, Log2<V>::v, 1, V + 1>(tile, tw);
      for (int idx = threadIdx.x; idx < N * V; idx += blockDim.x) {
        const int r = idx / V, c = idx - r * V;
        const int rs = inverse ? r : (int)f2l[r];  // forward: frequency r sits at slot f2l[r]
        img[(long)r * N + v0 + c] = cscale(tile[rs * (V + 1) + c], scale);
      }
    } else {
      // tile[v][col], pitch N+1; vectors = rows (vstride N+1, estride 1)
      for (int idx = threadIdx.x; idx < V * N; idx += blockDim.x) {
        const int r = idx / N, c = idx - r * N;
        const int cs = inverse ? (int)f2l[c] : c;
        tile[r * (N + 1) + cs] = img[(long)(v0 + r) * N + c];
      }
      __syncthreads();
      if (inverse) fft_pass<N, true, Log2<V>::v, N + 1, 1>(tile, tw);
      else         fft_pass<N, false, Log2<V>::v, N + 1, 1>(tile, tw);
      for (int idx = threadIdx.x; idx < V * N; idx += blockDim.x) {
        const int r = idx / N, c = idx - r * N;
        const int cs = inverse ? c : (int)f2l[c];
        img[(long)(v0 + r) * N + c] = cscale(tile[r * (N + 1) + cs], scale);
      }
    }
    __syncthreads();
  }
}

template <int N>
int launch_smem(float2* x, long batch, int inverse, float scale, cudaStream_t st) {
  constexpr int NT = (N >= 128) ? 512 : (N >= 64 ? 256 : 128);
  const size_t smem = (size_t)N * (N + 1) * 8 + N * 8 + N * 2;
  auto k = fft2_smem_kernel<N>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error((int)e, "tb_fft2: %s", cudaGetErrorString(e));
  int sms = 148;
  tb_sm_count(&sms);
  const long per_sm = (N >= 128) ? 1 : (N >= 64 ? 4 : 8);
  long grid = batch < sms * per_sm ? batch : sms * per_sm;
  k<<<(unsigned)grid, NT, smem, st>>>(x, batch, inverse, scale);
  return check_launch("tb_fft2");
}

template <int N, int V>
int launch_two_pass(float2* x, long batch, int inverse, float scale, cudaStream_t st) {
  const size_t smem0 = (size_t)N * (V + 1) * 8 + N * 8 + N * 2;
  const size_t smem1 = (size_t)V * (N + 1) * 8 + N * 8 + N * 2;
  auto k0 = fft_axis_kernel<N, V, 0>;
  auto k1 = fft_axis_kernel<N, V, 1>;
  cudaError_t e = cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return set_error((int)e, "tb_fft2: %s", cudaGetErrorString(e));
  int sms = 148;
  tb_sm_count(&sms);
  const long ntiles = batch * (N / V);
  const long grid = ntiles < sms ? ntiles : sms;
  // forward: columns then rows; inverse: rows then columns (any order works,
  // the scale is applied once)
  if (!inverse) {
    k0<<<(unsigned)grid, 512, smem0, st>>>(x, batch, 0, 1.0f);
    k1<<<(unsigned)grid, 512, smem1, st>>>(x, batch, 0, scale);
  } else {
    k1<<<(unsigned)grid, 512, smem1, st>>>(x, batch, 1, 1.0f);
    k0<<<(unsigned)grid, 512, smem0, st>>>(x, batch, 1, scale);
  }
  return check_launch("tb_fft2");
}

}  // namespace tb

extern "C" int tb_fft2(void* x, int64_t batch, int n, int inverse, float scale,
                       tb_stream_t stream) {
  TB_REQUIRE(x != nullptr || batch == 0, TB_ERR_INVALID, "tb_fft2: null pointer");
  TB_REQUIRE(batch >= 0, TB_ERR_INVALID, "tb_fft2: negative batch");
  if (batch == 0) return TB_OK;
  float2* p = (float2*)x;
  cudaStream_t st = (cudaStream_t)stream;
  switch (n) {
    case 16:   return tb::launch_smem<16>(p, batch, inverse, scale, st);
    case 32:   return tb::launch_smem<32>(p, batch, inverse, scale, st);
    case 64:   return tb::launch_smem<64>(p, batch, inverse, scale, st);
    case 128:  return tb::launch_smem<128>(p, batch, inverse, scale, st);
    case 256:  return tb::launch_two_pass<256, 64>(p, batch, inverse, scale, st);
    case 512:  return tb::launch_two_pass<512, 32>(p, batch, inverse, scale, st);
    case 1024: return tb::launch_two_pass<1024, 16>(p, batch, inverse, scale, st);
    case 2048: return tb::launch_two_pass<2048, 8>(p, batch, inverse, scale, st);
    default:
      return tb::set_error(TB_ERR_UNSUPPORTED,
                           "tb_fft2: detector width %d is not a power of two in [16, 2048]", n);
  }
}
