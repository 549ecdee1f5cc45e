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

// ---- arbitrary sizes: Bluestein (chirp-z) on top of the power-of-two path ----
// The reference's cuFFT accepts any detector width.  For n that is not a
// power of two, n k = (n^2 + k^2 - (k - n)^2) / 2 turns the DFT into a circular
// convolution of size P >= 2n - 1 (a power of two):
//   X[k1,k2] = a[k1] a[k2] * sum_n (x[n1,n2] a[n1] a[n2]) b[k1-n1] b[k2-n2],
//   a[n] = exp(-i pi n^2 / n_), b = conj(a),
// i.e. chirp multiply + zero pad -> P x P transform -> x spectrum of the chirp
// filter (separable: fb[k1] fb[k2]) -> inverse P x P transform -> chirp
// multiply.  Phases use n^2 mod 2n in integers, so they are exact to float32.
// The inverse transform is conj . forward . conj.

__global__ void bluestein_tables_kernel(int n, int P, float2* __restrict__ a, float2* __restrict__ fb) {
  // a[j] = exp(-i pi j^2 / n); fb = length-P DFT of the wrapped conj chirp (double accumulation)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const long q = ((long)t * t) % (2L * n);
    float sn, cs;
    sincospif((float)q / (float)n, &sn, &cs);
    a[t] = make_float2(cs, -sn);
  }
  if (t < P) {
    double re = 0.0, im = 0.0;
    for (int m = -(n - 1); m <= n - 1; ++m) {
      const long q = ((long)m * m) % (2L * n);       // b[m] = exp(+i pi m^2 / n)
      const int mp = m < 0 ? m + P : m;              // wrapped position
      const long r = ((long)mp * t) % P;             // DFT phase -2 pi mp t / P
      double sb, cb, sw, cw;
      sincospi((double)q / (double)n, &sb, &cb);
      sincospi(2.0 * (double)r / (double)P, &sw, &cw);
      // (cb + i sb) * (cw - i sw)
      re += cb * cw + sb * sw;
      im += sb * cw - cb * sw;
    }
    fb[t] = make_float2((float)re, (float)im);
  }
}

__global__ void __launch_bounds__(256)
bluestein_pad_kernel(const float2* __restrict__ x, float2* __restrict__ buf, long count, int n,
                     int P, const float2* __restrict__ a, int inverse) {
  const long total = count * P * (long)P;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long img = i / ((long)P * P);
    const int r = (int)((i / P) % P), c = (int)(i % P);
    float2 v = make_float2(0.f, 0.f);
    if (r < n && c < n) {
      v = x[(img * n + r) * n + c];
      if (inverse) v.y = -v.y;
      v = cmul(v, cmul(a[r], a[c]));
    }
    buf[i] = v;
  }
}

__global__ void __launch_bounds__(256)
bluestein_filter_kernel(float2* __restrict__ buf, long count, int P, const float2* __restrict__ fb) {
  const long total = count * P * (long)P;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int r = (int)((i / P) % P), c = (int)(i % P);
    buf[i] = cmul(buf[i], cmul(fb[r], fb[c]));
  }
}

__global__ void __launch_bounds__(256)
bluestein_unpad_kernel(const float2* __restrict__ buf, float2* __restrict__ x, long count, int n,
                       int P, const float2* __restrict__ a, int inverse, float scale) {
  const long total = count * n * (long)n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long img = i / ((long)n * n);
    const int r = (int)((i / n) % n), c = (int)(i % n);
    float2 v = cmul(buf[(img * P + r) * P + c], cmul(a[r], a[c]));
    if (inverse) v.y = -v.y;
    x[i] = cscale(v, scale);
  }
}

static int fft2_pow2(float2* p, long batch, int n, int inverse, float scale, cudaStream_t st);

static int fft2_bluestein(float2* x, long batch, int n, int inverse, float scale, cudaStream_t st) {
  int P = 16;
  while (P < 2 * n - 1) P *= 2;
  if (P > 2048)
    return set_error(TB_ERR_UNSUPPORTED,
                     "tb_fft2: width %d is not a power of two and exceeds 1024", n);
  int sms = 148;
  tb_sm_count(&sms);
  long chunk = (256L << 20) / ((long)P * P * 8);
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  // stream-ordered scratch from the driver's pool: tables + one chunk of padded images
  const size_t bytes = ((size_t)n + P + (size_t)chunk * P * P) * sizeof(float2);
  float2* scratch = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&scratch, bytes, st);
  if (e != cudaSuccess) return set_error((int)e, "tb_fft2: scratch: %s", cudaGetErrorString(e));
  float2* a = scratch;
  float2* fb = a + n;
  float2* buf = fb + P;
  bluestein_tables_kernel<<<(P + 127) / 128, 128, 0, st>>>(n, P, a, fb);
  int rc = check_launch("tb_fft2(bluestein tables)");
  const unsigned grid = (unsigned)(sms * 16);
  for (long i0 = 0; rc == TB_OK && i0 < batch; i0 += chunk) {
    const long count = batch - i0 < chunk ? batch - i0 : chunk;
    float2* xi = x + i0 * (long)n * n;
    bluestein_pad_kernel<<<grid, 256, 0, st>>>(xi, buf, count, n, P, a, inverse);
    rc = check_launch("tb_fft2(bluestein pad)");
    if (rc == TB_OK) rc = fft2_pow2(buf, count, P, 0, 1.0f, st);
    if (rc == TB_OK) {
      bluestein_filter_kernel<<<grid, 256, 0, st>>>(buf, count, P, fb);
      rc = check_launch("tb_fft2(bluestein filter)");
    }
    if (rc == TB_OK) rc = fft2_pow2(buf, count, P, 1, 1.0f / ((float)P * (float)P), st);
    if (rc == TB_OK) {
      bluestein_unpad_kernel<<<grid, 256, 0, st>>>(buf, xi, count, n, P, a, inverse, scale);
      rc = check_launch("tb_fft2(bluestein unpad)");
    }
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

static int fft2_pow2(float2* p, long batch, int n, int inverse, float scale, cudaStream_t st) {
  switch (n) {
    case 16:   return launch_smem<16>(p, batch, inverse, scale, st);
    case 32:   return launch_smem<32>(p, batch, inverse, scale, st);
    case 64:   return launch_smem<64>(p, batch, inverse, scale, st);
    case 128:  return launch_smem<128>(p, batch, inverse, scale, st);
    case 256:  return launch_two_pass<256, 64>(p, batch, inverse, scale, st);
    case 512:  return launch_two_pass<512, 32>(p, batch, inverse, scale, st);
    case 1024: return launch_two_pass<1024, 16>(p, batch, inverse, scale, st);
    case 2048: return launch_two_pass<2048, 8>(p, batch, inverse, scale, st);
    default:   return set_error(TB_ERR_UNSUPPORTED, "tb_fft2: width %d", n);
  }
}

}  // namespace tb

extern "C" int tb_fft2(void* x, int64_t batch, int n, int inverse, float scale,
                       tb_stream_t stream) {
  TB_REQUIRE(x != nullptr || batch == 0, TB_ERR_INVALID, "tb_fft2: null pointer");
  TB_REQUIRE(batch >= 0, TB_ERR_INVALID, "tb_fft2: negative batch");
  if (batch == 0) return TB_OK;
  float2* p = (float2*)x;
  cudaStream_t st = (cudaStream_t)stream;
  TB_REQUIRE(n >= 2 && n <= 2048, TB_ERR_UNSUPPORTED,
             "tb_fft2: detector width %d is outside [2, 2048]", n);
  const bool pow2 = (n & (n - 1)) == 0;
  if (pow2 && n >= 16) return tb::fft2_pow2(p, batch, n, inverse, scale, st);
  // any other width (the reference's cuFFT takes them all): chirp-z transform
  return tb::fft2_bluestein(p, batch, n, inverse, scale, st);
}
