// Fused forward model: object patch x probe -> zero pad -> 2-D FFT ->
// far-field wave and/or detector intensity, one CTA per scan position with the
// wavefront resident in shared memory.
// Replaces Ptycho.fwd / _compute_intensity
// (src/tike/operators/cupy/ptycho.py:114-204; convolution.py:58-101;
//  propagation.py:43-57; ptycho/ptycho.py:95-124).
#include "../../include/tike_b200.h"
#include "wave.cuh"

namespace tb {

template <int ND> struct FwdCfg {
  static constexpr int NT = (ND >= 128) ? 512 : (ND >= 64 ? 256 : 128);
  static constexpr int PER_SM = (ND >= 128) ? 1 : (ND >= 64 ? 4 : 8);
  static constexpr size_t smem = (size_t)ND * (ND + 1) * 8 + ND * ND * 4 + ND * 8 + ND * 4;
};

template <int ND>
__global__ void __launch_bounds__(FwdCfg<ND>::NT)
ptycho_fwd_kernel(tb_batch b, float2* __restrict__ farplane,
                  float* __restrict__ intensity) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float* acc = reinterpret_cast<float*>(tile + ND * (ND + 1));
  float2* tw = reinterpret_cast<float2*>(acc + ND * ND);
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  unsigned short* f2l = l2f + ND;
  fill_twiddles<ND>(tw);
  fill_perm<ND>(l2f, f2l);
  __syncthreads();

  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const int pad = (ND - b.probe_width) / 2;
  const float2* psi = (const float2*)b.psi;
  const float s2 = b.fwd_scale * b.fwd_scale;

  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner c = make_corner(b.scan, s);
    if (intensity)
      for (int i = threadIdx.x; i < ND * ND; i += blockDim.x) acc[i] = 0.f;
    for (int m = 0; m < b.nmodes; ++m) {
      build_exitwave<ND>(tile, psi, b.height, b.width, c, ps, s, m, pad);
      __syncthreads();
      fft2_tile<ND, false>(tile, tw);
      if (farplane) {
        float2* out = farplane + ((long)s * b.nmodes + m) * ND * ND;
        for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x) {
          const int ky = idx / ND, kx = idx - ky * ND;
          out[idx] = cscale(tile[f2l[ky] * (ND + 1) + f2l[kx]], b.fwd_scale);
        }
      }
      if (intensity) {
        for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x) {
          const int ly = idx / ND, lx = idx - ly * ND;
          acc[idx] += cabs2(tile[ly * (ND + 1) + lx]) * s2;
        }
      }
      __syncthreads();
    }
    if (intensity) {
      float* out = intensity + (long)s * ND * ND;
      for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x) {
        const int ky = idx / ND, kx = idx - ky * ND;
        out[idx] = acc[f2l[ky] * ND + f2l[kx]];
      }
      __syncthreads();
    }
  }
}

// Large detectors: write the zero-padded exit wave to HBM, transform it with
// the two-pass tb_fft2, reduce the intensity with a third kernel.
__global__ void exitwave_kernel(tb_batch b, float2* __restrict__ nearplane) {
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const int ND = b.detector_width, N = b.probe_width;
  const int pad = (ND - N) / 2;
  const float2* psi = (const float2*)b.psi;
  const long per_pos = (long)b.nmodes * ND * ND;
  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner c = make_corner(b.scan, s);
    // gridDim.y CTAs share the pixels of one position
    for (long i = (long)blockIdx.y * blockDim.x + threadIdx.x; i < (long)ND * ND;
         i += (long)blockDim.x * gridDim.y) {
      const int ly = (int)(i / ND), lx = (int)(i - (long)ly * ND);
      const int py = ly - pad, px = lx - pad;
      const bool inside = py >= 0 && py < N && px >= 0 && px < N;
      float2 o = make_float2(0.f, 0.f);
      if (inside) o = patch_value(psi, b.height, b.width, c, py, px);
      for (int m = 0; m < b.nmodes; ++m) {
        float2 v = make_float2(0.f, 0.f);
        if (inside) v = cmul(probe_value(ps, s, m, py, px), o);
        nearplane[s * per_pos + (long)m * ND * ND + i] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
intensity_kernel(const float2* __restrict__ farplane, float* __restrict__ intensity,
                 long npos, int M, long npix) {
  const long total = npos * npix;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long s = i / npix, p = i - s * npix;
    float a = 0.f;
    for (int m = 0; m < M; ++m) a += cabs2(farplane[(s * M + m) * npix + p]);
    intensity[i] = a;
  }
}

template <int ND>
int launch_fwd(const tb_batch& b, float2* farplane, float* intensity, cudaStream_t st) {
  auto k = ptycho_fwd_kernel<ND>;
  const size_t smem = FwdCfg<ND>::smem;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error((int)e, "tb_ptycho_fwd: %s", cudaGetErrorString(e));
  int sms = 148;
  tb_sm_count(&sms);
  long grid = (long)sms * FwdCfg<ND>::PER_SM;
  if (b.npos < grid) grid = b.npos;
  k<<<(unsigned)grid, FwdCfg<ND>::NT, smem, st>>>(b, farplane, intensity);
  return check_launch("tb_ptycho_fwd");
}

int check_batch(const tb_batch* b, const char* who) {
  TB_REQUIRE(b != nullptr, TB_ERR_INVALID, "%s: null batch", who);
  // an empty batch (npos == 0) may come with null scan / data pointers
  TB_REQUIRE(b->psi && b->probe && (b->scan || b->npos == 0), TB_ERR_INVALID,
             "%s: null array", who);
  TB_REQUIRE(b->height > 0 && b->width > 0 && b->npos >= 0 && b->nmodes > 0 &&
                 b->probe_width > 0, TB_ERR_INVALID, "%s: bad shape", who);
  TB_REQUIRE(b->detector_width >= b->probe_width, TB_ERR_INVALID,
             "%s: probe width %d exceeds detector width %d", who,
             b->probe_width, b->detector_width);
  TB_REQUIRE(b->eigen_probe == nullptr || b->eigen_weights != nullptr,
             TB_ERR_INVALID, "%s: eigen_probe without eigen_weights", who);
  TB_REQUIRE(b->eigen_probe == nullptr ||
                 (b->neigen > 0 && b->eigen_modes > 0 && b->eigen_modes <= b->nmodes),
             TB_ERR_INVALID, "%s: bad eigen probe shape", who);
  return TB_OK;
}

}  // namespace tb

extern "C" int tb_ptycho_fwd(const tb_batch* b, void* farplane, float* intensity,
                             tb_stream_t stream) {
  int rc = tb::check_batch(b, "tb_ptycho_fwd");
  if (rc != TB_OK) return rc;
  if (b->npos == 0) return TB_OK;
  TB_REQUIRE(farplane || intensity, TB_ERR_INVALID, "tb_ptycho_fwd: no output requested");
  tb_batch bb = *b;
  if (bb.eigen_probe == nullptr) bb.neigen = 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (b->detector_width) {
    case 16:  return tb::launch_fwd<16>(bb, (float2*)farplane, intensity, st);
    case 32:  return tb::launch_fwd<32>(bb, (float2*)farplane, intensity, st);
    case 64:  return tb::launch_fwd<64>(bb, (float2*)farplane, intensity, st);
    case 128: return tb::launch_fwd<128>(bb, (float2*)farplane, intensity, st);
    default: break;
  }
  // two-pass path for detectors that exceed shared memory
  TB_REQUIRE(farplane != nullptr, TB_ERR_INVALID,
             "tb_ptycho_fwd: detector width %d needs a farplane buffer (two-pass FFT)",
             b->detector_width);
  int sms = 148;
  tb_sm_count(&sms);
  long grid = (long)sms * 8;
  if (bb.npos < grid) grid = bb.npos;
  tb::exitwave_kernel<<<(unsigned)grid, 256, 0, st>>>(bb, (float2*)farplane);
  rc = tb::check_launch("tb_ptycho_fwd(exitwave)");
  if (rc != TB_OK) return rc;
  rc = tb_fft2(farplane, bb.npos * bb.nmodes, bb.detector_width, 0, bb.fwd_scale, stream);
  if (rc != TB_OK) return rc;
  if (intensity) {
    const long npix = (long)bb.detector_width * bb.detector_width;
    tb::intensity_kernel<<<(unsigned)(sms * 8), 256, 0, st>>>(
        (const float2*)farplane, intensity, bb.npos, bb.nmodes, npix);
    rc = tb::check_launch("tb_ptycho_fwd(intensity)");
  }
  return rc;
}
