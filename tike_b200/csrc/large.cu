// Solver batch for detectors that exceed shared memory (ND >= 256).
//
// Same contract as the fused kernel (RpieDev), but the wavefronts of a chunk
// of positions live in HBM and every 2-D transform is the two-pass row/column
// FFT of propagation.cu (tb_fft2):
//
//   exit wave (probe x patch, zero padded)  -> tb_fft2 forward
//   -> intensity / cost / modulus or Poisson factor (one CTA per position)
//   -> tb_fft2 inverse -> gradients (object scatter, probe sums, chi, ...)
//
// Replaces the same reference code as rpie.cu (rpie.py:355-505,
// lstsq.py:422-579) for BASELINE configs 3 (256^2) and 5 (512^2).
#include <cstdlib>

#include "solver_dev.cuh"

namespace tb {

// forward.cu
__global__ void exitwave_kernel(tb_batch b, float2* __restrict__ nearplane);
// large_fused.cu
int run_large_fused_chunk(const RpieDev& a, float2* wave, float* sums, long s0, long count,
                          bool need_back, int sms, cudaStream_t st, const char* who);

static inline bool pow2_large_width(int nd) {
  return nd == 256 || nd == 512 || nd == 1024 || nd == 2048;
}

// fused three-kernel pipeline (large_fused.cu) or the unfused chain below
static inline bool large_uses_fused(const tb_batch& b, int noise_model) {
  static const bool unfused = getenv("TB_LARGE_UNFUSED") != nullptr;  // development switch
  return noise_model == TB_NOISE_GAUSSIAN && !unfused && pow2_large_width(b.detector_width);
}

static inline long large_chunk(const tb_batch& b, int noise_model) {
  const long per_pos = (long)b.nmodes * b.detector_width * b.detector_width * 8;
  static const long chunk_mb = [] {
    const char* e = getenv("TB_LARGE_CHUNK_MB");  // development switch
    const long v = e ? atol(e) : 0;
    return v > 0 ? v : 256L;
  }();
  // ~256 MiB of wavefronts per chunk; the unfused chain runs one CTA per
  // position, so it takes 1 GiB to keep every SM busy
  long c = ((large_uses_fused(b, noise_model) ? chunk_mb : 4 * chunk_mb) << 20) / per_pos;
  if (c < 1) c = 1;
  if (c > b.npos) c = b.npos;
  return c;
}

int64_t large_workspace_bytes(const tb_batch& b, bool replica, int noise_model) {
  const long c = large_chunk(b, noise_model);
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  const long wave = c * (long)b.nmodes * b.detector_width * b.detector_width;
  const long gobj = c * (long)b.probe_width * b.probe_width;
  return (wave + gobj + (replica ? (long)kMaxReplicas * n : 0)) * 8;
}

// One CTA per position: intensity over modes, cost, factor applied in place.
// far: (C, M, ND, ND) natural frequency order, already scaled by fwd_scale.
// iplane: optional (C, ND, ND) float scratch; the Poisson passes then read the
// intensity of pass 1 back instead of re-summing all M planes every pass.
__global__ void __launch_bounds__(512)
modulus_kernel(RpieDev a, float2* __restrict__ far, float* __restrict__ iplane, long s0,
               long count) {
  __shared__ float red[3 * 32];
  __shared__ float steps[64];  // Poisson step length per mode (M <= 64)
  const int ND = a.b.detector_width, M = a.b.nmodes;
  const long npix = (long)ND * ND;
  const bool gaussian = a.noise_model == TB_NOISE_GAUSSIAN;
  for (long i = blockIdx.x; i < count; i += gridDim.x) {
    const long s = s0 + i;
    float2* w = far + i * M * npix;
    float* Ic = (iplane && !gaussian) ? iplane + i * npix : nullptr;
    const long dbase = s * npix;
    float sums[3] = {0.f, 0.f, 0.f};
    float step_dom = a.step_start;
    // pass 1: cost (+ gaussian factor applied right away)
    for (long p = threadIdx.x; p < npix; p += blockDim.x) {
      float I = 0.f;
      for (int m = 0; m < M; ++m) I += cabs2(w[m * npix + p]);
      const bool meas = a.mask ? (a.mask[p] != 0) : true;
      if (meas) {
        const float d = load_data(a.data, a.data_u16, dbase + p);
        if (gaussian) {
          const float sd = sqrtf(d), sI = sqrtf(I);
          const float t = sI - sd;
          sums[0] += t * t;
          const float f = -(1.0f - sd / (sI + 1e-9f));
          for (int m = 0; m < M; ++m) w[m * npix + p] = cscale(w[m * npix + p], f);
        } else {
          if (Ic) Ic[p] = I;
          sums[0] += I - d * logf(I + 1e-9f);
          if (a.step_mode == TB_STEP_DOMINANT_MODE) {
            const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
            sums[1] += xi * xi * I;
            sums[2] += xi * (I - d / (1.0f - step_dom * xi));
          }
        }
      } else {
        for (int m = 0; m < M; ++m)
          w[m * npix + p] = cscale(w[m * npix + p], a.unmeasured_factor);
      }
    }
    block_sum<3>(sums, red);
    if (threadIdx.x == 0) a.costs[s] = sums[0] * a.inv_nmeasured;
    if (gaussian) continue;

    // Poisson step lengths (exitwave.py:122-234), then scale measured pixels
    if (a.step_mode == TB_STEP_DOMINANT_MODE) {
      step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (sums[2] / sums[1]);
      float s1[1] = {0.f};
      for (long p = threadIdx.x; p < npix; p += blockDim.x) {
        const bool meas = a.mask ? (a.mask[p] != 0) : true;
        if (!meas) continue;
        float I = 0.f;
        if (Ic) I = Ic[p];
        else for (int m = 0; m < M; ++m) I += cabs2(w[m * npix + p]);
        const float d = load_data(a.data, a.data_u16, dbase + p);
        const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
        s1[0] += xi * (I - d / (1.0f - step_dom * xi));
      }
      block_sum<1>(s1, red);
      step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (s1[0] / sums[1]);
    }
    // Per-mode step lengths.  Nothing is scaled until the final pass, so the
    // unscaled |Psi_m|^2 and the intensity can be recomputed from `w`.
    for (int m = 0; m < M; ++m) {
      float step = step_dom;
      if (a.step_mode == TB_STEP_ALL_MODES) {
        step = a.step_start;
        float q0 = 0.f;
        for (int it = 0; it < 2; ++it) {
          float q[2] = {0.f, 0.f};
          for (long p = threadIdx.x; p < npix; p += blockDim.x) {
            const bool meas = a.mask ? (a.mask[p] != 0) : true;
            if (!meas) continue;
            float I = 0.f;
            if (Ic) I = Ic[p];
            else for (int mm = 0; mm < M; ++mm) I += cabs2(w[mm * npix + p]);
            const float d = load_data(a.data, a.data_u16, dbase + p);
            const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
            const float ab = cabs2(w[m * npix + p]);
            const float t = xi * step - 1.0f;
            const float den = ab * t * t + I - ab;
            q[0] += xi * xi * ab;
            q[1] += xi * ab * (1.0f + (d * t) / den);
          }
          block_sum<2>(q, red);
          if (it == 0) q0 = q[0];
          step = step * (1.0f - a.step_weight) + (q[1] / q0) * a.step_weight;
        }
      }
      if (threadIdx.x == 0) steps[m] = step;
    }
    __syncthreads();
    for (long p = threadIdx.x; p < npix; p += blockDim.x) {
      const bool meas = a.mask ? (a.mask[p] != 0) : true;
      if (!meas) continue;
      float I = 0.f;
      if (Ic) I = Ic[p];
      else for (int mm = 0; mm < M; ++mm) I += cabs2(w[mm * npix + p]);
      const float d = load_data(a.data, a.data_u16, dbase + p);
      const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
      for (int mm = 0; mm < M; ++mm)
        w[mm * npix + p] = cscale(w[mm * npix + p], -steps[mm] * xi);
    }
    __syncthreads();
  }
}

// One CTA per position: gradients from the back-propagated waves.
// near: (C, M, ND, ND) = chi (padded); gobj: (C, N, N) scratch.
template <bool POS>
__device__ __forceinline__ void
gradient_body(const RpieDev& a, const float2* __restrict__ near, float2* __restrict__ gobj,
              long s0, long count, float* red) {
  const tb_batch& b = a.b;
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const int ND = b.detector_width, N = b.probe_width, M = b.nmodes;
  const int pad = (ND - N) / 2, H = b.height, W = b.width;
  const long npix = (long)ND * ND;
  const float2* psi = (const float2*)b.psi;
  const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * N * N : nullptr;
  for (long i = blockIdx.x; i < count; i += gridDim.x) {
    const long s = s0 + i;
    const Corner c = make_corner(b.scan, s);
    const float2* chi_all = near + i * M * npix;
    float2* G = gobj + i * (long)N * N;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // eig num/den, pos sums
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      const int py = idx / N, px = idx - py * N;
      const float2 o = patch_value(psi, H, W, c, py, px);
      float2 acc = make_float2(0.f, 0.f);
      for (int m = 0; m < M; ++m) {
        const float2 chi = chi_all[m * npix + (long)(pad + py) * ND + pad + px];
        if (a.chi_out) a.chi_out[((long)s * M + m) * N * N + idx] = chi;
        if (a.accumulate_object) {
          const float2 g = cmulc(probe_value(ps, s, m, py, px), chi);
          acc.x += g.x;
          acc.y += g.y;
        }
        if (replica) red_add_f32x2(replica + (long)m * N * N + idx, cmulc(o, chi));
        if (m == 0 && a.eig_step) {
          const float2 p0 = __ldg(ps.probe + (ps.per_position ? s * (long)M * N * N : 0) + idx);
          const float2 op = cmul(o, p0);
          v[0] += op.x * chi.x + op.y * chi.y;
          v[1] += cabs2(op);
        }
        if (POS && m == 0 && a.pos_num) {
          const int crop = N / 4;
          if (py >= crop && py < N - crop && px >= crop && px < N - crop) {
            float2 gy = make_float2(0.f, 0.f), gx = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = -2; t <= 2; ++t) {
              const float wt = a.taps[t + 2];
              const int qy = min(max(py + t, 0), N - 1), qx = min(max(px + t, 0), N - 1);
              const float2 oy = patch_value(psi, H, W, c, qy, px);
              const float2 ox = patch_value(psi, H, W, c, py, qx);
              gy.x -= wt * oy.x; gy.y -= wt * oy.y;
              gx.x -= wt * ox.x; gx.y -= wt * ox.y;
            }
            const float2 p0u = probe_value(ps, s, 0, py, px);
            const float2 ay = cmul(gy, p0u), ax = cmul(gx, p0u);
            v[2] += ay.x * chi.x + ay.y * chi.y;
            v[3] += cabs2(ay);
            v[4] += ax.x * chi.x + ax.y * chi.y;
            v[5] += cabs2(ax);
          }
        }
      }
      if (a.accumulate_object) {
        const int y = c.iy + py, x = c.ix + px;
        const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
        G[idx] = lead_ok ? cscale(acc, inv_m) : make_float2(0.f, 0.f);
      }
    }
    if (a.eig_step || (POS && a.pos_num)) {
      block_sum<6>(v, red);
      if (threadIdx.x == 0) {
        if (a.eig_step) a.eig_step[s] = 0.1f * (v[0] / v[1]);
        if (POS && a.pos_num) {
          a.pos_num[2 * s] = v[2];
          a.pos_den[2 * s] = v[3];
          a.pos_num[2 * s + 1] = v[4];
          a.pos_den[2 * s + 1] = v[5];
        }
      }
    }
    if (a.accumulate_object) {
      __syncthreads();  // G (global scratch) complete for this position
      const int T = N + 1;
      for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
        const int ty = t / T, tx = t - ty * T;
        const int y = c.iy + ty, x = c.ix + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float2 r = make_float2(0.f, 0.f);
        const bool a0 = ty < N, a1 = ty > 0, b0 = tx < N, b1 = tx > 0;
        if (a0 & b0) { const float2 g = __ldcg(G + ty * N + tx);           r.x += c.w00 * g.x; r.y += c.w00 * g.y; }
        if (a0 & b1) { const float2 g = __ldcg(G + ty * N + tx - 1);       r.x += c.w01 * g.x; r.y += c.w01 * g.y; }
        if (a1 & b0) { const float2 g = __ldcg(G + (ty - 1) * N + tx);     r.x += c.w10 * g.x; r.y += c.w10 * g.y; }
        if (a1 & b1) { const float2 g = __ldcg(G + (ty - 1) * N + tx - 1); r.x += c.w11 * g.x; r.y += c.w11 * g.y; }
        red_add_f32x2(a.psi_num + (long)y * W + x, r);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(512)
gradient_kernel(RpieDev a, const float2* __restrict__ near, float2* __restrict__ gobj,
                long s0, long count) {
  __shared__ float red[6 * 32];
  gradient_body<true>(a, near, gobj, s0, count, red);
}

// without the position-gradient sums (the register-hungry part): 2 CTAs per SM
__global__ void __launch_bounds__(512, 2)
gradient_nopos_kernel(RpieDev a, const float2* __restrict__ near, float2* __restrict__ gobj,
                      long s0, long count) {
  __shared__ float red[6 * 32];
  gradient_body<false>(a, near, gobj, s0, count, red);
}

int run_large(RpieDev a, int64_t workspace_bytes, void* workspace, float2* probe_out,
              cudaStream_t st, const char* who) {
  const tb_batch& b = a.b;
  const int nd = b.detector_width;
  // powers of two from 256 up: fused three-kernel pipeline; any other width
  // (the reference's cuFFT takes them all): unfused chain on the chirp-z tb_fft2
  const bool pow2_large = pow2_large_width(nd);
  TB_REQUIRE(pow2_large || (nd >= 2 && nd <= 1024 && !fused_width(nd)), TB_ERR_UNSUPPORTED,
             "%s: detector width %d is not supported (powers of two up to 2048, "
             "other widths up to 1024)", who, nd);
  TB_REQUIRE(b.nmodes <= 64, TB_ERR_UNSUPPORTED, "%s: more than 64 probe modes", who);
  if (!probe_out) a.probe_sums = 0;
  const bool replica = a.probe_sums != 0;
  const int64_t need = large_workspace_bytes(b, replica, a.noise_model);
  TB_REQUIRE(workspace && workspace_bytes >= need, TB_ERR_INVALID,
             "%s: workspace too small (%lld < %lld bytes)", who,
             (long long)workspace_bytes, (long long)need);
  const long chunk = large_chunk(b, a.noise_model);
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  const long npix = (long)nd * nd;
  float2* wave = (float2*)workspace;
  float2* gobj = wave + chunk * b.nmodes * npix;
  a.replicas = gobj + chunk * (long)b.probe_width * b.probe_width;
  a.nrep = kMaxReplicas;
  if (replica) {
    cudaError_t e = cudaMemsetAsync(a.replicas, 0, (size_t)a.nrep * n * 8, st);
    if (e != cudaSuccess) return set_error((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
  }
  const bool need_back = a.accumulate_object || a.probe_sums || a.eig_step || a.chi_out || a.pos_num;
  int sms = 148;
  tb_sm_count(&sms);
  if (large_uses_fused(b, a.noise_model)) {
    // fused three-kernel pipeline (large_fused.cu); costs are accumulated
    cudaError_t e = cudaMemsetAsync(a.costs, 0, (size_t)b.npos * sizeof(float), st);
    if (e != cudaSuccess) return set_error((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
    for (long s0 = 0; s0 < b.npos; s0 += chunk) {
      const long count = (b.npos - s0 < chunk) ? b.npos - s0 : chunk;
      const int rc = run_large_fused_chunk(a, wave, (float*)gobj, s0, count, need_back, sms, st, who);
      if (rc != TB_OK) return rc;
    }
    if (replica) {
      const long blocks = (n + 255) / 256;
      reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
          a.replicas, a.nrep, n, n, probe_out);
      return check_launch("reduce_replicas");
    }
    return TB_OK;
  }
  for (long s0 = 0; s0 < b.npos; s0 += chunk) {
    const long count = (b.npos - s0 < chunk) ? b.npos - s0 : chunk;
    tb_batch sub = b;
    sub.scan = b.scan + 2 * s0;
    sub.npos = count;
    if (b.probe_per_position) sub.probe = (const float2*)b.probe + s0 * n;
    if (b.eigen_weights) sub.eigen_weights = b.eigen_weights + s0 * (long)(b.neigen + 1) * b.nmodes;
    long grid = (long)sms * 8 < count ? (long)sms * 8 : count;
    long gy = ((long)sms * 8 + grid - 1) / grid;  // short chunks: split positions over CTAs
    if (gy > npix / 256) gy = npix / 256;
    if (gy < 1) gy = 1;
    exitwave_kernel<<<dim3((unsigned)grid, (unsigned)gy), 256, 0, st>>>(sub, wave);
    int rc = check_launch(who);
    if (rc != TB_OK) return rc;
    rc = tb_fft2(wave, count * b.nmodes, nd, 0, b.fwd_scale, st);
    if (rc != TB_OK) return rc;
    grid = (long)sms * 3 < count ? (long)sms * 3 : count;  // 40 registers: 3 CTAs per SM
    // the object-gradient scratch is idle until gradient_kernel: park I there
    float* iplane = 2L * b.probe_width * b.probe_width >= npix ? (float*)gobj : nullptr;
    modulus_kernel<<<(unsigned)grid, 512, 0, st>>>(a, wave, iplane, s0, count);
    rc = check_launch(who);
    if (rc != TB_OK) return rc;
    if (!need_back) continue;
    rc = tb_fft2(wave, count * b.nmodes, nd, 1, b.inv_scale, st);
    if (rc != TB_OK) return rc;
    grid = (long)sms * 2 < count ? (long)sms * 2 : count;
    if (a.pos_num)
      gradient_kernel<<<(unsigned)grid, 512, 0, st>>>(a, wave, gobj, s0, count);
    else
      gradient_nopos_kernel<<<(unsigned)grid, 512, 0, st>>>(a, wave, gobj, s0, count);
    rc = check_launch(who);
    if (rc != TB_OK) return rc;
  }
  if (replica) {
    const long blocks = (n + 255) / 256;
    reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
        a.replicas, a.nrep, n, n, probe_out);
    return check_launch("reduce_replicas");
  }
  return TB_OK;
}

}  // namespace tb
