// Multislice objects (D > 1 slices), rPIE only -- the slice loop of the
// reference fork: Multislice.fwd_return_intermediate_probes
// (operators/cupy/multislice.py:97-139), FresnelSpectProp.fwd/adj
// (fresnelspectprop.py:52-111), the gradient loop of
// rpie._get_nearplane_gradients (ptycho/solvers/rpie.py:374, 441-474) and the
// per-slice object preconditioner (_preconditioner.py:48-100).
//
// Chunked through HBM: the probe incident on every slice is kept per position
// (`probes`, D x C x M x N x N), because the gradient of slice t needs it
// again.  Reference behaviour reproduced on purpose: the residual is carried
// to the previous slice by the adjoint Fresnel propagator ALONE (rpie.py:474,
// no multiplication by the conjugate transmission), the object gradient keeps
// its 1/M, and only probe_update_numerator[0] is consumed by _update.
// Requires probe width == detector width (multislice.py:127 stores the
// propagated exit wave in an array of the probe's shape).
#include <cstdlib>

#include "solver_dev.cuh"

namespace tb {

// forward.cu / large.cu
__global__ void exitwave_kernel(tb_batch b, float2* __restrict__ nearplane);
__global__ void modulus_kernel(RpieDev a, float2* __restrict__ far, float* __restrict__ iplane,
                               long s0, long count);
__global__ void gradient_kernel(RpieDev a, const float2* __restrict__ near,
                                float2* __restrict__ gobj, long s0, long count);
__global__ void intensity_kernel(const float2* __restrict__ farplane, float* __restrict__ intensity,
                                 long npos, int M, long npix);

// x[b, i] *= p[i]  (or conj(p[i])), times a real scale
__global__ void __launch_bounds__(256)
cmul_bcast_kernel(float2* __restrict__ x, const float2* __restrict__ p, long batch, long n,
                  int conj, float scale) {
  const long total = batch * n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    float2 w = __ldg(p + i % n);
    if (conj) w.y = -w.y;
    x[i] = cscale(cmul(x[i], w), scale);
  }
}

// psi_precond slice += scatter_s( sum_m |probe[s, m]|^2 ) with the four
// bilinear weights (_preconditioner.py:78-94 -> patch.adj)
__global__ void __launch_bounds__(256)
precond_scatter_var_kernel(const float2* __restrict__ probes, int M, int N,
                           const float* __restrict__ scan, long npos, float* __restrict__ amp_all,
                           float2* __restrict__ out, int H, int W) {
  // amp_all: (npos, N, N) float scratch in global memory (any N fits)
  for (long s = blockIdx.x; s < npos; s += gridDim.x) {
    const Corner c = make_corner(scan, s);
    const float2* ps = probes + s * (long)M * N * N;
    float* amp = amp_all + s * (long)N * N;
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
      float a = 0.f;
      for (int m = 0; m < M; ++m) a += cabs2(ps[(long)m * N * N + i]);
      amp[i] = a;
    }
    __syncthreads();  // this block's amp[] is complete and visible to the block
    const int T = N + 1;
    for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
      const int ty = t / T, tx = t - ty * T;
      const int y = c.iy + ty, x = c.ix + tx;
      if (y < 0 || y >= H || x < 0 || x >= W) continue;
      float v = 0.f;
      const bool a0 = ty < N, a1 = ty > 0 && (y - 1) >= 0, b0 = tx < N, b1 = tx > 0 && (x - 1) >= 0;
      if (a0 & b0) v += c.w00 * amp[ty * N + tx];
      if (a0 & b1) v += c.w01 * amp[ty * N + tx - 1];
      if (a1 & b0) v += c.w10 * amp[(ty - 1) * N + tx];
      if (a1 & b1) v += c.w11 * amp[(ty - 1) * N + tx - 1];
      red_add_f32(reinterpret_cast<float*>(out + (long)y * W + x), v);
    }
  }
}

// One Fresnel-spectrum step with the whole image resident in shared memory
// (ND in 16..128, powers of two): [exit wave of slice t built in place |or| image loaded] ->
// forward 2-D transform -> x H (or conj H) -> inverse transform -> store.
// One launch and one HBM write (plus one read when not BUILD) instead of
// exit-wave kernel + tb_fft2 + multiply + tb_fft2.
template <int ND, bool BUILD>
__global__ void __launch_bounds__((ND >= 128) ? 512 : (ND >= 64 ? 256 : 128))
fresnel_tile_kernel(tb_batch b, float2* __restrict__ x, const float2* __restrict__ prop, int conj,
                    long nimg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  constexpr int P = ND + 1, LG = Log2<ND>::v;
  float2* tw = tile + ND * P;
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  fill_twiddles<ND>(tw);
  for (int i = threadIdx.x; i < ND; i += blockDim.x) l2f[i] = (unsigned short)loc2freq<ND>(i);
  __syncthreads();
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const float inv_n2 = 1.0f / ((float)ND * (float)ND);
  for (long img = blockIdx.x; img < nimg; img += gridDim.x) {
    float2* g = x + img * (long)ND * ND;
    if constexpr (BUILD) {
      const long s = img / b.nmodes;
      const int m = (int)(img - s * b.nmodes);
      const Corner c = make_corner(b.scan, s);
      build_exitwave<ND>(tile, (const float2*)b.psi, b.height, b.width, c, ps, s, m, 0);
    } else {
      for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x)
        tile[(idx >> LG) * P + (idx & (ND - 1))] = g[idx];
    }
    __syncthreads();
    fft2_tile<ND, false>(tile, tw);
    // slot (r, c) holds frequency (l2f[r], l2f[c])
    for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x) {
      const int r = idx >> LG, c = idx & (ND - 1);
      float2 h = __ldg(prop + (int)l2f[r] * ND + (int)l2f[c]);
      if (conj) h.y = -h.y;
      tile[r * P + c] = cscale(cmul(tile[r * P + c], h), inv_n2);
    }
    __syncthreads();
    fft2_tile<ND, true>(tile, tw);
    for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x)
      g[idx] = tile[(idx >> LG) * P + (idx & (ND - 1))];
    __syncthreads();
  }
}

template <int ND, bool BUILD>
static int launch_fresnel_tile(const tb_batch& b, float2* x, const float2* prop, int conj,
                               long nimg, int sms, cudaStream_t st) {
  constexpr int NT = (ND >= 128) ? 512 : (ND >= 64 ? 256 : 128);
  const size_t smem = (size_t)ND * (ND + 1) * 8 + ND * 8 + ND * 2;
  auto k = fresnel_tile_kernel<ND, BUILD>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error((int)e, "multislice: %s", cudaGetErrorString(e));
  const long per_sm = (ND >= 128) ? 1 : (ND >= 64 ? 4 : 8);
  const long grid = nimg < sms * per_sm ? nimg : sms * per_sm;
  k<<<(unsigned)grid, NT, smem, st>>>(b, x, prop, conj, nimg);
  return check_launch("multislice: Fresnel step");
}

// dispatch; returns TB_ERR_UNSUPPORTED for widths that do not fit shared memory
template <bool BUILD>
static int fresnel_tile(const tb_batch& b, float2* x, const float2* prop, int conj, long nimg,
                        int sms, cudaStream_t st) {
  switch (b.probe_width) {
    case 16:  return launch_fresnel_tile<16, BUILD>(b, x, prop, conj, nimg, sms, st);
    case 32:  return launch_fresnel_tile<32, BUILD>(b, x, prop, conj, nimg, sms, st);
    case 64:  return launch_fresnel_tile<64, BUILD>(b, x, prop, conj, nimg, sms, st);
    case 128: return launch_fresnel_tile<128, BUILD>(b, x, prop, conj, nimg, sms, st);
    default:  return TB_ERR_UNSUPPORTED;
  }
}

static inline long ms_chunk(const tb_batch& b) {
  const long per_pos = (long)b.nmodes * b.probe_width * b.probe_width * 8;
  long c = (512L << 20) / per_pos;  // ~512 MiB of wavefronts per slice per chunk: enough
                                    // positions per launch to fill the GPU
  if (c < 1) c = 1;
  if (c > b.npos) c = b.npos;
  return c;
}

struct MsLayout {
  long chunk, wave_elems;
  float2 *probes, *wave, *gobj, *replicas;
};

static int64_t ms_bytes(const tb_batch& b, int D) {
  const long c = ms_chunk(b);
  const long wave = c * (long)b.nmodes * b.probe_width * b.probe_width;
  const long gobj = c * (long)b.probe_width * b.probe_width;
  const long rep = (long)D * kMaxReplicas * b.nmodes * b.probe_width * b.probe_width;
  return ((long)(D + 1) * wave + gobj + rep) * 8;
}

static MsLayout ms_layout(const tb_batch& b, int D, void* workspace) {
  MsLayout L;
  L.chunk = ms_chunk(b);
  L.wave_elems = L.chunk * (long)b.nmodes * b.probe_width * b.probe_width;
  L.probes = (float2*)workspace;             // D planes; plane 0 unused
  L.wave = L.probes + (long)D * L.wave_elems;
  L.gobj = L.wave + L.wave_elems;
  L.replicas = L.gobj + L.chunk * (long)b.probe_width * b.probe_width;
  return L;
}

static int grid1d(long total, int sms) {
  const long blocks = (total + 255) / 256;
  const long cap = (long)sms * 16;
  return (int)(blocks < cap ? blocks : cap);
}

// probes[t+1] = Fresnel( probes[t] x patch(psi[t]) ), in place in `dst`
static int fresnel(float2* x, const float2* prop, long batch, int n, int conj, int sms,
                   cudaStream_t st) {
  if (fused_width(n)) {
    tb_batch dummy{};
    dummy.probe_width = n;
    dummy.nmodes = 1;
    return fresnel_tile<false>(dummy, x, prop, conj, batch, sms, st);
  }
  int rc = tb_fft2(x, batch, n, 0, 1.0f, st);
  if (rc != TB_OK) return rc;
  cmul_bcast_kernel<<<grid1d(batch * n * n, sms), 256, 0, st>>>(
      x, prop, batch, (long)n * n, conj, 1.0f / ((float)n * (float)n));
  rc = check_launch("multislice: propagator");
  if (rc != TB_OK) return rc;
  return tb_fft2(x, batch, n, 1, 1.0f, st);
}

// Forward model of one chunk: fills L.probes[1..D-1] and leaves the exit wave
// of the last slice in L.wave.
static int ms_forward_chunk(const tb_batch& b, int D, const float2* prop, const MsLayout& L,
                            long s0, long count, int sms, cudaStream_t st) {
  const long hw = (long)b.height * b.width;
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  for (int t = 0; t < D; ++t) {
    tb_batch sub = b;
    sub.psi = (const float2*)b.psi + (long)t * hw;
    sub.scan = b.scan + 2 * s0;
    sub.npos = count;
    if (t == 0) {
      if (b.probe_per_position) sub.probe = (const float2*)b.probe + s0 * n;
      if (b.eigen_weights)
        sub.eigen_weights = b.eigen_weights + s0 * (long)(b.neigen + 1) * b.nmodes;
    } else {
      sub.probe = L.probes + (long)t * L.wave_elems;
      sub.probe_per_position = 1;
      sub.eigen_probe = nullptr; sub.eigen_weights = nullptr; sub.neigen = 0;
    }
    float2* dst = (t == D - 1) ? L.wave : L.probes + (long)(t + 1) * L.wave_elems;
    if (t < D - 1 && fused_width(b.probe_width)) {
      // exit wave and the Fresnel step to the next slice in one launch
      const int rc = fresnel_tile<true>(sub, dst, prop, 0, count * b.nmodes, sms, st);
      if (rc != TB_OK) return rc;
      continue;
    }
    long grid = (long)sms * 8 < count ? (long)sms * 8 : count;
    exitwave_kernel<<<(unsigned)grid, 256, 0, st>>>(sub, dst);
    int rc = check_launch("multislice: exit wave");
    if (rc != TB_OK) return rc;
    if (t < D - 1) {
      rc = fresnel(dst, prop, count * b.nmodes, b.probe_width, 0, sms, st);
      if (rc != TB_OK) return rc;
    }
  }
  return TB_OK;
}

static int ms_check(const tb_batch* b, int D, const void* prop, const char* who) {
  int rc = check_batch(b, who);
  if (rc != TB_OK) return rc;
  TB_REQUIRE(D >= 1, TB_ERR_INVALID, "%s: nslices must be >= 1", who);
  TB_REQUIRE(prop != nullptr || D == 1, TB_ERR_INVALID, "%s: null propagator", who);
  TB_REQUIRE(b->probe_width == b->detector_width, TB_ERR_UNSUPPORTED,
             "%s: multislice needs probe width == detector width (%d != %d)", who,
             b->probe_width, b->detector_width);
  return TB_OK;
}

}  // namespace tb

extern "C" {

int64_t tb_multislice_workspace_size(const tb_batch* b, int nslices) {
  if (!b || nslices < 1) return 0;
  int64_t need = tb::ms_bytes(*b, nslices);
  const int64_t fused = tb::multislice_fused_workspace_bytes(*b, nslices);
  const int64_t precond = tb::multislice_precond_fused_workspace_bytes(*b, nslices);
  if (fused > need) need = fused;
  if (precond > need) need = precond;
  return need;
}

int tb_multislice_fwd(const tb_batch* b, int nslices, const void* propagator, void* farplane,
                      float* intensity, void* workspace, int64_t workspace_bytes,
                      tb_stream_t stream) {
  int rc = tb::ms_check(b, nslices, propagator, "tb_multislice_fwd");
  if (rc != TB_OK) return rc;
  if (b->npos == 0) return TB_OK;
  TB_REQUIRE(workspace && workspace_bytes >= tb::ms_bytes(*b, nslices), TB_ERR_INVALID,
             "tb_multislice_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  tb_sm_count(&sms);
  const tb::MsLayout L = tb::ms_layout(*b, nslices, workspace);
  const int nd = b->detector_width;
  const long npix = (long)nd * nd;
  for (long s0 = 0; s0 < b->npos; s0 += L.chunk) {
    const long count = (b->npos - s0 < L.chunk) ? b->npos - s0 : L.chunk;
    rc = tb::ms_forward_chunk(*b, nslices, (const float2*)propagator, L, s0, count, sms, st);
    if (rc != TB_OK) return rc;
    rc = tb_fft2(L.wave, count * b->nmodes, nd, 0, b->fwd_scale, st);
    if (rc != TB_OK) return rc;
    if (farplane) {
      cudaError_t e = cudaMemcpyAsync((float2*)farplane + s0 * b->nmodes * npix, L.wave,
                                      (size_t)count * b->nmodes * npix * 8,
                                      cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess)
        return tb::set_error((int)e, "tb_multislice_fwd: %s", cudaGetErrorString(e));
    }
    if (intensity) {
      tb::intensity_kernel<<<(unsigned)(sms * 8), 256, 0, st>>>(L.wave, intensity + s0 * npix,
                                                               count, b->nmodes, npix);
      rc = tb::check_launch("tb_multislice_fwd(intensity)");
      if (rc != TB_OK) return rc;
    }
  }
  return TB_OK;
}

int tb_multislice_rpie_batch(const tb_rpie_args* a, int nslices, const void* propagator,
                             tb_stream_t stream) {
  TB_REQUIRE(a != nullptr, TB_ERR_INVALID, "tb_multislice_rpie_batch: null args");
  int rc = tb::ms_check(&a->batch, nslices, propagator, "tb_multislice_rpie_batch");
  if (rc != TB_OK) return rc;
  if (a->batch.npos == 0) return TB_OK;
  TB_REQUIRE(a->data && a->costs, TB_ERR_INVALID, "tb_multislice_rpie_batch: null data/costs");
  TB_REQUIRE(!a->accumulate_object || (a->psi_numerator && a->probe_numerator), TB_ERR_INVALID,
             "tb_multislice_rpie_batch: numerators required");
  TB_REQUIRE(a->noise_model == TB_NOISE_GAUSSIAN || a->noise_model == TB_NOISE_POISSON,
             TB_ERR_INVALID, "tb_multislice_rpie_batch: unknown noise model %d", a->noise_model);
  TB_REQUIRE(a->num_measured > 0, TB_ERR_INVALID, "tb_multislice_rpie_batch: num_measured");
  TB_REQUIRE(a->batch.nmodes <= 64, TB_ERR_UNSUPPORTED, "tb_multislice_rpie_batch: > 64 modes");
  const tb_batch& b = a->batch;
  if (b.npos == 0) return TB_OK;
  TB_REQUIRE(a->workspace && a->workspace_bytes >= tb_multislice_workspace_size(&b, nslices),
             TB_ERR_INVALID, "tb_multislice_rpie_batch: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // probe as wide as the detector (32 / 64 / 128), shared probe, Gaussian model:
  // the whole slice loop of a position in one persistent CTA (multislice_fused.cu);
  // TB_MULTISLICE_UNFUSED=1 forces the chunked chain below (A/B and tests)
  if (tb::multislice_fused_applies(*a, nslices)) {
    const char* e = getenv("TB_MULTISLICE_UNFUSED");
    if (!(e && atoi(e) != 0)) return tb::run_multislice_fused(*a, nslices, propagator, st);
  }
  int sms = 148;
  tb_sm_count(&sms);
  const int D = nslices, nd = b.detector_width;
  const tb::MsLayout L = tb::ms_layout(b, D, a->workspace);
  const long hw = (long)b.height * b.width;
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  const bool back = a->accumulate_object || a->eigen_weight_step;

  tb::RpieDev d{};
  d.b = b;
  if (d.b.eigen_probe == nullptr) d.b.neigen = 0;
  d.data = a->data;
  d.data_u16 = (a->data_dtype == TB_DATA_U16);
  d.mask = a->mask;
  d.noise_model = a->noise_model;
  d.step_mode = a->step_mode;
  d.step_start = a->step_length_start;
  d.step_weight = a->step_length_weight;
  d.unmeasured_factor = a->unmeasured_scaling - 1.0f;
  d.inv_nmeasured = 1.0f / (float)a->num_measured;
  d.costs = a->costs;
  d.divide_by_modes = 1;
  d.nrep = tb::kMaxReplicas;
  if (a->accumulate_object) {
    cudaError_t e = cudaMemsetAsync(L.replicas, 0, (size_t)D * d.nrep * n * 8, st);
    if (e != cudaSuccess)
      return tb::set_error((int)e, "tb_multislice_rpie_batch: %s", cudaGetErrorString(e));
  }
  for (long s0 = 0; s0 < b.npos; s0 += L.chunk) {
    const long count = (b.npos - s0 < L.chunk) ? b.npos - s0 : L.chunk;
    rc = tb::ms_forward_chunk(b, D, (const float2*)propagator, L, s0, count, sms, st);
    if (rc != TB_OK) return rc;
    rc = tb_fft2(L.wave, count * b.nmodes, nd, 0, b.fwd_scale, st);
    if (rc != TB_OK) return rc;
    long grid = (long)sms * 2 < count ? (long)sms * 2 : count;
    float* iplane = 2L * b.probe_width * b.probe_width >= (long)nd * nd ? (float*)L.gobj : nullptr;
    tb::modulus_kernel<<<(unsigned)grid, 512, 0, st>>>(d, L.wave, iplane, s0, count);
    rc = tb::check_launch("tb_multislice_rpie_batch(modulus)");
    if (rc != TB_OK || !back) { if (rc != TB_OK) return rc; continue; }
    rc = tb_fft2(L.wave, count * b.nmodes, nd, 1, b.inv_scale, st);
    if (rc != TB_OK) return rc;
    for (int t = D - 1; t >= 0; --t) {
      tb::RpieDev g = d;
      g.b.psi = (const float2*)b.psi + (long)t * hw;
      g.b.scan = b.scan + 2 * s0;
      g.b.npos = count;
      if (t == 0) {
        if (b.probe_per_position) g.b.probe = (const float2*)b.probe + s0 * n;
        if (b.eigen_weights)
          g.b.eigen_weights = b.eigen_weights + s0 * (long)(b.neigen + 1) * b.nmodes;
        g.eig_step = a->eigen_weight_step ? a->eigen_weight_step + s0 : nullptr;
      } else {
        g.b.probe = L.probes + (long)t * L.wave_elems;
        g.b.probe_per_position = 1;
        g.b.eigen_probe = nullptr; g.b.eigen_weights = nullptr; g.b.neigen = 0;
        g.eig_step = nullptr;
      }
      g.accumulate_object = a->accumulate_object;
      g.probe_sums = a->accumulate_object;
      g.psi_num = a->accumulate_object ? (float2*)a->psi_numerator + (long)t * hw : nullptr;
      g.replicas = L.replicas + (long)t * d.nrep * n;
      // gradient_kernel indexes positions from 0 inside the chunk (s0 = 0)
      tb::gradient_kernel<<<(unsigned)grid, 512, 0, st>>>(g, L.wave, L.gobj, 0, count);
      rc = tb::check_launch("tb_multislice_rpie_batch(gradient)");
      if (rc != TB_OK) return rc;
      if (t == 0) break;
      rc = tb::fresnel(L.wave, (const float2*)propagator, count * b.nmodes, b.probe_width, 1,
                       sms, st);  // rpie.py:474
      if (rc != TB_OK) return rc;
    }
  }
  if (a->accumulate_object) {
    const long blocks = (n + 255) / 256;
    for (int t = 0; t < D; ++t) {
      tb::reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
          L.replicas + (long)t * d.nrep * n, d.nrep, n, n, (float2*)a->probe_numerator + (long)t * n);
      rc = tb::check_launch("tb_multislice_rpie_batch(reduce)");
      if (rc != TB_OK) return rc;
    }
  }
  return TB_OK;
}

// lstsq_grad on a multislice object, exactly as far as the fork takes it
// (lstsq.py:422-530): the far field comes from the multislice forward model
// (op.fwd), everything after the back-propagation to the near plane treats chi
// as the exit wave of slice 0 -- object gradient through the UNPROPAGATED unique
// probe into object_upd_sum[0] (lstsq.py:512-520), probe gradient through the
// patches of psi[0] (lstsq.py:522-539), position sums on those patches too.
int tb_multislice_lstsq_phase1(const tb_lstsq_args* a, int nslices, const void* propagator,
                               tb_stream_t stream) {
  TB_REQUIRE(a != nullptr, TB_ERR_INVALID, "tb_multislice_lstsq_phase1: null args");
  int rc = tb::ms_check(&a->batch, nslices, propagator, "tb_multislice_lstsq_phase1");
  if (rc != TB_OK) return rc;
  if (a->batch.npos == 0) return TB_OK;
  TB_REQUIRE(a->data && a->costs && a->chi, TB_ERR_INVALID,
             "tb_multislice_lstsq_phase1: null data/costs/chi");
  TB_REQUIRE(!a->recover_psi || a->object_upd_sum, TB_ERR_INVALID,
             "tb_multislice_lstsq_phase1: object_upd_sum required");
  TB_REQUIRE(!a->recover_probe || a->probe_upd_sum, TB_ERR_INVALID,
             "tb_multislice_lstsq_phase1: probe_upd_sum required");
  TB_REQUIRE(!a->recover_positions || (a->position_num && a->position_den), TB_ERR_INVALID,
             "tb_multislice_lstsq_phase1: position buffers required");
  TB_REQUIRE(a->noise_model == TB_NOISE_GAUSSIAN || a->noise_model == TB_NOISE_POISSON,
             TB_ERR_INVALID, "tb_multislice_lstsq_phase1: unknown noise model %d",
             a->noise_model);
  TB_REQUIRE(a->num_measured > 0, TB_ERR_INVALID, "tb_multislice_lstsq_phase1: num_measured");
  TB_REQUIRE(a->batch.nmodes <= 64, TB_ERR_UNSUPPORTED,
             "tb_multislice_lstsq_phase1: > 64 modes");
  const tb_batch& b = a->batch;
  TB_REQUIRE(a->workspace && a->workspace_bytes >= tb::ms_bytes(b, nslices), TB_ERR_INVALID,
             "tb_multislice_lstsq_phase1: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  tb_sm_count(&sms);
  const int D = nslices, nd = b.detector_width;
  const tb::MsLayout L = tb::ms_layout(b, D, a->workspace);
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;

  tb::RpieDev d{};
  d.b = b;
  if (d.b.eigen_probe == nullptr) d.b.neigen = 0;
  d.data = a->data;
  d.data_u16 = (a->data_dtype == TB_DATA_U16);
  d.mask = a->mask;
  d.noise_model = a->noise_model;
  d.step_mode = a->step_mode;
  d.step_start = a->step_length_start;
  d.step_weight = a->step_length_weight;
  d.unmeasured_factor = a->unmeasured_scaling - 1.0f;
  d.inv_nmeasured = 1.0f / (float)a->num_measured;
  d.costs = a->costs;
  d.divide_by_modes = 0;
  d.poisson_eps = 1;
  d.nrep = tb::kMaxReplicas;
  if (a->recover_probe) {
    cudaError_t e = cudaMemsetAsync(L.replicas, 0, (size_t)d.nrep * n * 8, st);
    if (e != cudaSuccess)
      return tb::set_error((int)e, "tb_multislice_lstsq_phase1: %s", cudaGetErrorString(e));
  }
  for (long s0 = 0; s0 < b.npos; s0 += L.chunk) {
    const long count = (b.npos - s0 < L.chunk) ? b.npos - s0 : L.chunk;
    rc = tb::ms_forward_chunk(b, D, (const float2*)propagator, L, s0, count, sms, st);
    if (rc != TB_OK) return rc;
    rc = tb_fft2(L.wave, count * b.nmodes, nd, 0, b.fwd_scale, st);
    if (rc != TB_OK) return rc;
    long grid = (long)sms * 2 < count ? (long)sms * 2 : count;
    float* iplane = 2L * b.probe_width * b.probe_width >= (long)nd * nd ? (float*)L.gobj : nullptr;
    tb::modulus_kernel<<<(unsigned)grid, 512, 0, st>>>(d, L.wave, iplane, s0, count);
    rc = tb::check_launch("tb_multislice_lstsq_phase1(modulus)");
    if (rc != TB_OK) return rc;
    rc = tb_fft2(L.wave, count * b.nmodes, nd, 1, b.inv_scale, st);
    if (rc != TB_OK) return rc;
    // slice 0, the caller's own probe; positions indexed from 0 inside the chunk
    tb::RpieDev g = d;
    g.b.scan = b.scan + 2 * s0;
    g.b.npos = count;
    if (b.probe_per_position) g.b.probe = (const float2*)b.probe + s0 * n;
    if (b.eigen_weights)
      g.b.eigen_weights = b.eigen_weights + s0 * (long)(b.neigen + 1) * b.nmodes;
    g.accumulate_object = a->recover_psi ? 1 : 0;
    g.psi_num = (float2*)a->object_upd_sum;
    g.probe_sums = a->recover_probe ? 1 : 0;
    g.replicas = L.replicas;
    g.chi_out = (float2*)a->chi + s0 * n;
    if (a->recover_positions) {
      g.pos_num = a->position_num + 2 * s0;
      g.pos_den = a->position_den + 2 * s0;
      for (int i = 0; i < 5; ++i) g.taps[i] = a->gradient_taps[i];
    }
    tb::gradient_kernel<<<(unsigned)grid, 512, 0, st>>>(g, L.wave, L.gobj, 0, count);
    rc = tb::check_launch("tb_multislice_lstsq_phase1(gradient)");
    if (rc != TB_OK) return rc;
  }
  if (a->recover_probe) {
    const long blocks = (n + 255) / 256;
    tb::reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
        L.replicas, d.nrep, n, n, (float2*)a->probe_upd_sum);
    rc = tb::check_launch("tb_multislice_lstsq_phase1(reduce)");
    if (rc != TB_OK) return rc;
  }
  return TB_OK;
}

int tb_multislice_precond_psi(const tb_batch* b, int nslices, const void* propagator,
                              void* psi_precond, void* workspace, int64_t workspace_bytes,
                              tb_stream_t stream) {
  int rc = tb::ms_check(b, nslices, propagator, "tb_multislice_precond_psi");
  if (rc != TB_OK) return rc;
  TB_REQUIRE(psi_precond != nullptr, TB_ERR_INVALID, "tb_multislice_precond_psi: null output");
  TB_REQUIRE(workspace && workspace_bytes >= tb_multislice_workspace_size(b, nslices),
             TB_ERR_INVALID, "tb_multislice_precond_psi: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  tb_sm_count(&sms);
  const long hw = (long)b->height * b->width;
  const int N = b->probe_width;
  // slice 0: the plain preconditioner of the (shared) probe; slices >= 1: the
  // probe that reaches the slice, per position
  rc = tb_precond_psi(b->probe, b->nmodes, N, b->scan, nullptr, b->npos, psi_precond, b->height, b->width,
                      (float*)workspace, st);
  if (rc != TB_OK || nslices == 1) return rc;
  cudaError_t e = cudaMemsetAsync((float2*)psi_precond + hw, 0, (size_t)(nslices - 1) * hw * 8, st);
  if (e != cudaSuccess)
    return tb::set_error((int)e, "tb_multislice_precond_psi: %s", cudaGetErrorString(e));
  if (b->npos == 0) return TB_OK;
  const tb::MsLayout L = tb::ms_layout(*b, nslices, workspace);
  tb_batch plain = *b;  // _preconditioner.py:76 starts from parameters.probe, no weights
  plain.eigen_probe = nullptr; plain.eigen_weights = nullptr; plain.neigen = 0;
  plain.probe_per_position = 0;
  if (tb::multislice_precond_fused_applies(plain, nslices)) {
    const char* e2 = getenv("TB_MULTISLICE_UNFUSED");
    if (!(e2 && atoi(e2) != 0))
      return tb::run_multislice_precond_fused(plain, nslices, propagator, psi_precond, workspace,
                                              st);
  }
  for (long s0 = 0; s0 < b->npos; s0 += L.chunk) {
    const long count = (b->npos - s0 < L.chunk) ? b->npos - s0 : L.chunk;
    rc = tb::ms_forward_chunk(plain, nslices, (const float2*)propagator, L, s0, count, sms, st);
    if (rc != TB_OK) return rc;
    for (int t = 1; t < nslices; ++t) {
      long grid = (long)sms * 8 < count ? (long)sms * 8 : count;
      tb::precond_scatter_var_kernel<<<(unsigned)grid, 256, 0, st>>>(
          L.probes + (long)t * L.wave_elems, b->nmodes, N, b->scan + 2 * s0, count,
          (float*)L.gobj, (float2*)psi_precond + (long)t * hw, b->height, b->width);
      rc = tb::check_launch("tb_multislice_precond_psi");
      if (rc != TB_OK) return rc;
    }
  }
  return TB_OK;
}

}  // extern "C"
