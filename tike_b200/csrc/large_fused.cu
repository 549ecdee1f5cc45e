// Large-detector (ND >= 256) solver batch, Gaussian noise model: the two-pass
// row/column FFT with the elementwise work fused into its three trips
// through HBM.  The wavefronts of a chunk of positions live in one buffer
// `wave` (count, M, ND, ND) that every kernel updates in place, and stay in
// digit-reversed frequency order between the kernels (the inverse DIT passes
// want exactly that), so no permutation pass exists:
//
//   K1  exit wave (probe x patch, zero padded) built in shared memory
//       + forward column transforms                         -> wave
//   K2  forward row transforms + |Psi|^2 over modes + cost + modulus factor
//       + inverse row transforms (one CTA: all modes of V rows)  wave -> wave
//   K3  inverse column transforms + conj(probe).chi / conj(patch).chi /
//       chi_out / eigen-weight and position sums + object scatter   wave ->
//
// That is 1 write + 1 read/write + 1 read of M*ND^2*8 bytes per position
// instead of 12 such trips in the unfused chain (exit wave, 4 FFT passes,
// modulus, gradient), which remains in large.cu for the Poisson model.
// (Measured and dropped in round 2: a K2 that keeps all modes of a row block in
// one tile, so that nothing is written back and re-read between its forward
// and inverse row passes -- 172 k vs 175 k patterns/s per epoch at 256^2 x 4
// modes: the write-back / re-read of the mode-by-mode kernel stays in L2.)
// Replaces: rpie.py:355-505, lstsq.py:422-579 at BASELINE configs 3 and 5.
#include "solver_dev.cuh"

namespace tb {

template <int ND> struct LargeCfg {
  // vectors per tile: K1/K3 work on ND x V column blocks, K2 on V x ND row blocks
  static constexpr int V = (ND <= 256) ? 64 : (ND <= 512 ? 32 : (ND <= 1024 ? 16 : 8));
  // K1 / K3 column block and CTA size: at 256^2 a 16-column block (one 128-byte
  // line per row) lets two K3 CTAs and four K1 CTAs share an SM
  static constexpr int VC = (ND <= 256) ? V / 4 : V / 2;
  static constexpr int NTC = (ND <= 256) ? 256 : 512;
  static constexpr int CTAS_K1 = (ND <= 256) ? 4 : 2;
  static constexpr int CTAS_K3 = (ND <= 256) ? 2 : 1;
  static constexpr int VR = V / 4;  // K2 row block: four small CTAs of NTR threads per SM, so
  static constexpr int NTR = 128;   // that the loads of some overlap the transforms of others
  static constexpr size_t smem_col = (size_t)ND * (VC + 1) * 8 + ND * 8;
  static constexpr size_t smem_grad = smem_col + 2 * (size_t)ND * VC * 8;
  static constexpr size_t smem_row = (size_t)VR * (ND + 1) * 8 + (size_t)VR * ND * 4 + ND * 8 +
                                     2 * ND * 2 + 32 * 4;
};

__device__ __forceinline__ ProbeSet probe_set(const tb_batch& b) {
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  return ps;
}

// ---- K1: exit wave + forward column transforms ------------------------------
template <int ND>
__global__ void __launch_bounds__(LargeCfg<ND>::NTC, LargeCfg<ND>::CTAS_K1)
large_exit_cols_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count) {
  constexpr int VC = LargeCfg<ND>::VC, P = VC + 1, NCB = ND / VC, LV = Log2<VC>::v;
  constexpr int NTC = LargeCfg<ND>::NTC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + ND * P;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const ProbeSet ps = probe_set(b);
  const int N = b.probe_width, M = b.nmodes, pad = (ND - N) / 2;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const bool simple = (N == ND) && ps.weights == nullptr && !ps.per_position;
  const long total = count * M * NCB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const int m = (int)((t / NCB) % M);
    const long i = t / ((long)NCB * M);
    const long s = s0 + i;
    const Corner c = make_corner(b.scan, s);
    const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < b.height) & (c.ix + N < b.width);
    if (simple && interior) {
      // probe width == detector width, shared probe, patch inside the object:
      // plain batched loads, no bounds logic
      const float2* __restrict__ pm = ps.probe + (long)m * ND * ND + cb * VC;
      const float2* __restrict__ o0 = psi + (long)c.iy * b.width + c.ix + cb * VC;
      const int W = b.width;
      constexpr int KPT = ND * VC / NTC;
#pragma unroll
      for (int k0 = 0; k0 < KPT; k0 += 4) {
        float2 pv[4], q[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NTC;
          const int r = idx >> LV, cc = idx & (VC - 1);
          pv[j] = __ldg(pm + (long)r * ND + cc);
          const float2* r0 = o0 + (long)r * W + cc;
          q[j][0] = __ldg(r0); q[j][1] = __ldg(r0 + 1);
          q[j][2] = __ldg(r0 + W); q[j][3] = __ldg(r0 + W + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NTC;
          const int r = idx >> LV, cc = idx & (VC - 1);
          float2 o;
          o.x = q[j][0].x * c.w00; o.y = q[j][0].y * c.w00;
          o.x += q[j][1].x * c.w01; o.y += q[j][1].y * c.w01;
          o.x += q[j][2].x * c.w10; o.y += q[j][2].y * c.w10;
          o.x += q[j][3].x * c.w11; o.y += q[j][3].y * c.w11;
          tile[r * P + cc] = cmul(pv[j], o);
        }
      }
    } else {
      for (int idx = threadIdx.x; idx < ND * VC; idx += NTC) {
        const int r = idx >> LV, cc = idx & (VC - 1);
        const int py = r - pad, px = cb * VC + cc - pad;
        float2 v = make_float2(0.f, 0.f);
        if (py >= 0 && py < N && px >= 0 && px < N)
          v = cmul(probe_value(ps, s, m, py, px), patch_value(psi, b.height, b.width, c, py, px));
        tile[r * P + cc] = v;
      }
    }
    __syncthreads();
    fft_pass<ND, false, LV, 1, P>(tile, tw);  // columns; rows end up in slot order
    float2* img = wave + (i * M + m) * (long)ND * ND + cb * VC;
    for (int idx = threadIdx.x; idx < ND * VC; idx += NTC) {
      const int r = idx >> LV, cc = idx & (VC - 1);
      img[(long)r * ND + cc] = tile[r * P + cc];
    }
    __syncthreads();
  }
}

// ---- K2: forward rows + intensity + cost + modulus + inverse rows ------------
template <int ND>
__global__ void __launch_bounds__(LargeCfg<ND>::NTR, 4)
large_rows_modulus_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count,
                          int need_back) {
  constexpr int V = LargeCfg<ND>::VR, P = ND + 1, NRB = ND / V, LV = Log2<V>::v;
  constexpr int NT = LargeCfg<ND>::NTR;
  constexpr int LG = Log2<ND>::v;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float* F = reinterpret_cast<float*>(tile + V * P);
  float2* tw = reinterpret_cast<float2*>(F + V * ND);
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  unsigned short* f2l = l2f + ND;
  float* red = reinterpret_cast<float*>(f2l + ND);
  fill_twiddles<ND>(tw);
  fill_perm<ND>(l2f, f2l);
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const long total = count * NRB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int rb = (int)(t % NRB);
    const long i = t / NRB;
    const long s = s0 + i;
    float2* base = wave + i * M * (long)ND * ND + (long)rb * V * ND;
    for (int idx = threadIdx.x; idx < V * ND; idx += NT) F[idx] = 0.f;
    for (int m = 0; m < M; ++m) {
      float2* img = base + (long)m * ND * ND;
      // pull the rows this CTA reads next (next mode, else the first mode of
      // its next task) into L2 while this mode is transformed
      {
        const float2* nxt = nullptr;
        if (m + 1 < M) {
          nxt = img + (long)ND * ND;
        } else if (t + gridDim.x < total) {
          const long tn = t + gridDim.x;
          nxt = wave + (tn / NRB) * M * (long)ND * ND + (long)(tn % NRB) * V * ND;
        }
        if (nxt)
          for (int ln = threadIdx.x; ln < V * ND * 8 / 128; ln += NT)
            prefetch_l2((const char*)nxt + ln * 128);
      }
      // batches of independent loads keep the memory pipe full
      constexpr int KPT = V * ND / NT;
#pragma unroll 1
      for (int k0 = 0; k0 < KPT; k0 += 8) {
        float2 w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __ldcs(img + threadIdx.x + (k0 + j) * NT);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NT;
          tile[(idx >> LG) * P + (idx & (ND - 1))] = w[j];
        }
      }
      __syncthreads();
      fft_pass<ND, false, LV, P, 1>(tile, tw);  // rows; columns end up in slot order
      const bool keep = (m == M - 1);           // the last mode stays in the tile
      for (int idx = threadIdx.x; idx < V * ND; idx += NT) {
        const int r = idx >> LG, cc = idx & (ND - 1);
        const float2 w = tile[r * P + cc];
        F[idx] += cabs2(w) * s2;
        if (need_back && !keep) img[idx] = w;
      }
      __syncthreads();
    }
    // cost and modulus factor (objective.py:11-66); data is visited in its
    // natural order, the matching slot comes from the digit-reversal tables
    {
      float sums[1] = {0.f};
      constexpr int KPT = V * ND / NT;
#pragma unroll 1
      for (int k0 = 0; k0 < KPT; k0 += 8) {
        float d[8];
        bool meas[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NT;
          const long pix = (long)l2f[rb * V + (idx >> LG)] * ND + (idx & (ND - 1));
          meas[j] = a.mask ? (a.mask[pix] != 0) : true;
          d[j] = meas[j] ? load_data(a.data, a.data_u16, s * (long)ND * ND + pix) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NT;
          const int l = (idx >> LG) * ND + (int)f2l[idx & (ND - 1)];
          if (meas[j]) {
            const float sd = sqrtf(d[j]), sI = sqrtf(F[l]);
            const float dv = sI - sd;
            sums[0] += dv * dv;
            F[l] = -(1.0f - sd / (sI + 1e-9f)) * rt;
          } else {
            F[l] = a.unmeasured_factor * rt;
          }
        }
      }
      block_sum<1>(sums, red);
      if (threadIdx.x == 0) atomicAdd(a.costs + s, sums[0] * a.inv_nmeasured);
    }
    __syncthreads();
    if (!need_back) continue;
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi == 0) ? M - 1 : mi - 1;
      float2* img = base + (long)m * ND * ND;
      constexpr int KPT = V * ND / NT;
      if (mi == 0) {
        for (int idx = threadIdx.x; idx < V * ND; idx += NT) {
          float2& w = tile[(idx >> LG) * P + (idx & (ND - 1))];
          w = cscale(w, F[idx]);
        }
      } else {
#pragma unroll 1
        for (int k0 = 0; k0 < KPT; k0 += 8) {
          float2 w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = img[threadIdx.x + (k0 + j) * NT];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int idx = threadIdx.x + (k0 + j) * NT;
            tile[(idx >> LG) * P + (idx & (ND - 1))] = cscale(w[j], F[idx]);
          }
        }
      }
      __syncthreads();
      fft_pass<ND, true, LV, P, 1>(tile, tw);
      for (int idx = threadIdx.x; idx < V * ND; idx += NT) {
        const int r = idx >> LG, cc = idx & (ND - 1);
        img[idx] = tile[r * P + cc];
      }
      __syncthreads();
    }
  }
}

// ---- K3: inverse columns + gradients -----------------------------------------
// sums: (count, 6) partial sums per position: eigen num/den, position y num/den, x num/den
// The interpolated patch of the column block is computed once per position and
// parked in shared memory next to the tile, and so is the object-gradient
// accumulator over the modes (each thread only touches its own entries).
template <int ND>
__global__ void __launch_bounds__(LargeCfg<ND>::NTC, LargeCfg<ND>::CTAS_K3)
large_cols_gradient_kernel(RpieDev a, const float2* __restrict__ wave, float* __restrict__ sums,
                           long s0, long count) {
  constexpr int VC = LargeCfg<ND>::VC, P = VC + 1, NCB = ND / VC, LV = Log2<VC>::v;
  constexpr int NTC = LargeCfg<ND>::NTC;
  constexpr int KPT = ND * VC / NTC;  // pixels per thread
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[6 * 32];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + ND * P;
  float2* O = tw + ND;      // ND * VC patch values, linear in idx
  float2* A = O + ND * VC;  // ND * VC object-gradient accumulator, linear in idx
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const ProbeSet ps = probe_set(b);
  const int N = b.probe_width, M = b.nmodes, pad = (ND - N) / 2, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
  const bool simple = (N == ND) && ps.weights == nullptr && !ps.per_position;
  const bool extras = a.eig_step != nullptr || a.pos_num != nullptr;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * N * N : nullptr;
  const long total = count * NCB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const long i = t / NCB;
    const long s = s0 + i;
    const Corner c = make_corner(b.scan, s);
    const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < H) & (c.ix + N < W);
    const bool fastpath = simple && interior && !extras;
    // patch of this column block
    if (simple && interior) {
      const float2* __restrict__ o0 = psi + (long)c.iy * W + c.ix + cb * VC;
#pragma unroll
      for (int k0 = 0; k0 < KPT; k0 += 4) {
        float2 q[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NTC;
          const float2* r0 = o0 + (long)(idx >> LV) * W + (idx & (VC - 1));
          q[j][0] = __ldg(r0); q[j][1] = __ldg(r0 + 1);
          q[j][2] = __ldg(r0 + W); q[j][3] = __ldg(r0 + W + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 o;
          o.x = q[j][0].x * c.w00; o.y = q[j][0].y * c.w00;
          o.x += q[j][1].x * c.w01; o.y += q[j][1].y * c.w01;
          o.x += q[j][2].x * c.w10; o.y += q[j][2].y * c.w10;
          o.x += q[j][3].x * c.w11; o.y += q[j][3].y * c.w11;
          O[threadIdx.x + (k0 + j) * NTC] = o;
        }
      }
    } else {
      for (int idx = threadIdx.x; idx < ND * VC; idx += NTC) {
        const int py = (idx >> LV) - pad, px = cb * VC + (idx & (VC - 1)) - pad;
        O[idx] = (py >= 0 && py < N && px >= 0 && px < N) ? patch_value(psi, H, W, c, py, px)
                                                          : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < KPT; ++k) A[threadIdx.x + k * NTC] = make_float2(0.f, 0.f);
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int m = 0; m < M; ++m) {
      const float2* img = wave + (i * M + m) * (long)ND * ND + cb * VC;
      {
        // next column block this CTA will read -> L2 (VC * 8 bytes per row)
        const float2* nxt = nullptr;
        if (m + 1 < M) {
          nxt = img + (long)ND * ND;
        } else if (t + gridDim.x < total) {
          const long tn = t + gridDim.x;
          nxt = wave + (tn / NCB) * M * (long)ND * ND + (tn % NCB) * VC;
        }
        if (nxt) {
          constexpr int LPR = (VC * 8 + 127) / 128;  // lines per row
          for (int ln = threadIdx.x; ln < ND * LPR; ln += NTC)
            prefetch_l2((const char*)(nxt + (long)(ln / LPR) * ND) + (ln % LPR) * 128);
        }
      }
#pragma unroll
      for (int k0 = 0; k0 < KPT; k0 += 8) {
        float2 w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NTC;
          w[j] = __ldcs(img + (long)(idx >> LV) * ND + (idx & (VC - 1)));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = threadIdx.x + (k0 + j) * NTC;
          tile[(idx >> LV) * P + (idx & (VC - 1))] = w[j];
        }
      }
      __syncthreads();
      fft_pass<ND, true, LV, 1, P>(tile, tw);  // slot order in, natural rows out
      if (fastpath) {
        const float2* __restrict__ pm = ps.probe + (long)m * ND * ND + cb * VC;
        float2* __restrict__ cout =
            a.chi_out ? a.chi_out + ((long)s * M + m) * ND * ND + cb * VC : nullptr;
        float2* __restrict__ rep = replica ? replica + (long)m * ND * ND + cb * VC : nullptr;
#pragma unroll
        for (int k0 = 0; k0 < KPT; k0 += 8) {
          float2 pv[8];
          if (a.accumulate_object) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int idx = threadIdx.x + (k0 + j) * NTC;
              pv[j] = __ldg(pm + (long)(idx >> LV) * ND + (idx & (VC - 1)));
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int idx = threadIdx.x + (k0 + j) * NTC;
            const int r = idx >> LV, cc = idx & (VC - 1);
            const float2 chi = tile[r * P + cc];
            if (cout) __stcs(cout + (long)r * ND + cc, chi);
            if (a.accumulate_object) {
              const float2 g = cmulc(pv[j], chi);
              float2 t = A[idx];
              t.x += g.x;
              t.y += g.y;
              A[idx] = t;
            }
            if (rep) red_add_f32x2(rep + (long)r * ND + cc, cmulc(O[idx], chi));
          }
        }
      } else {
#pragma unroll 1
        for (int k = 0; k < KPT; ++k) {
          const int idx = threadIdx.x + k * NTC;
          const int r = idx >> LV, cc = idx & (VC - 1);
          const int py = r - pad, px = cb * VC + cc - pad;
          if (py < 0 || py >= N || px < 0 || px >= N) continue;
          const float2 chi = tile[r * P + cc];
          const long pidx = (long)py * N + px;
          if (a.chi_out) a.chi_out[((long)s * M + m) * N * N + pidx] = chi;
          if (a.accumulate_object) {
            const float2 g = cmulc(probe_value(ps, s, m, py, px), chi);
            float2 t = A[idx];
            t.x += g.x;
            t.y += g.y;
            A[idx] = t;
          }
          const float2 o = O[idx];
          if (replica) red_add_f32x2(replica + (long)m * N * N + pidx, cmulc(o, chi));
          if (m == 0 && a.eig_step) {
            // rpie.py:493-506 / lstsq.py:721-736: shared probe mode 0
            const float2 p0 = __ldg(ps.probe + (ps.per_position ? s * (long)M * N * N : 0) + pidx);
            const float2 op = cmul(o, p0);
            v[0] += op.x * chi.x + op.y * chi.y;
            v[1] += cabs2(op);
          }
          if (m == 0 && a.pos_num) {
            // lstsq.py:545-579 on the centre crop [N/4, N - N/4)
            const int crop = N / 4;
            if (py >= crop && py < N - crop && px >= crop && px < N - crop) {
              float2 gy = make_float2(0.f, 0.f), gx = make_float2(0.f, 0.f);
#pragma unroll
              for (int q = -2; q <= 2; ++q) {
                const float wt = a.taps[q + 2];
                const int qy = min(max(py + q, 0), N - 1), qx = min(max(px + q, 0), N - 1);
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
      }
      __syncthreads();
    }
    if (extras) {
      block_sum<6>(v, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) atomicAdd(sums + i * 6 + q, v[q]);
      }
    }
    if (a.accumulate_object) {
      // the block's share of the gradient goes through the tile so that the
      // four bilinear taps of an object pixel leave as one reduction
      // (convolution.cu:57-64); taps owned by the neighbouring column block
      // arrive with that block's reductions
#pragma unroll
      for (int k = 0; k < KPT; ++k) {
        const int idx = threadIdx.x + k * NTC;
        const int r = idx >> LV, cc = idx & (VC - 1);
        const int py = r - pad, px = cb * VC + cc - pad;
        const int y = c.iy + py, x = c.ix + px;
        const bool ok = (py >= 0) & (py < N) & (px >= 0) & (px < N) & (y >= 0) & (y < H) &
                        (x >= 0) & (x < W);
        tile[r * P + cc] = ok ? cscale(A[idx], inv_m) : make_float2(0.f, 0.f);
      }
      __syncthreads();
      // output pixels (ty, tx): rows 0..ND, columns 0..VC of this block
      for (int idx = threadIdx.x; idx < (ND + 1) * (VC + 1); idx += NTC) {
        const int ty = idx / (VC + 1), tx = idx - ty * (VC + 1);
        const int y = c.iy + ty - pad, x = c.ix + cb * VC + tx - pad;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float2 r = make_float2(0.f, 0.f);
        const bool a0 = ty < ND, a1 = ty > 0, b0 = tx < VC, b1 = tx > 0;
        if (a0 & b0) { const float2 g = tile[ty * P + tx];           r.x += c.w00 * g.x; r.y += c.w00 * g.y; }
        if (a0 & b1) { const float2 g = tile[ty * P + tx - 1];       r.x += c.w01 * g.x; r.y += c.w01 * g.y; }
        if (a1 & b0) { const float2 g = tile[(ty - 1) * P + tx];     r.x += c.w10 * g.x; r.y += c.w10 * g.y; }
        if (a1 & b1) { const float2 g = tile[(ty - 1) * P + tx - 1]; r.x += c.w11 * g.x; r.y += c.w11 * g.y; }
        if (r.x != 0.f || r.y != 0.f) red_add_f32x2(a.psi_num + (long)y * W + x, r);
      }
    }
    __syncthreads();
  }
}

__global__ void large_finalize_sums_kernel(RpieDev a, const float* __restrict__ sums, long s0,
                                           long count) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* v = sums + i * 6;
  const long s = s0 + i;
  if (a.eig_step) a.eig_step[s] = 0.1f * (v[0] / v[1]);
  if (a.pos_num) {
    a.pos_num[2 * s] = v[2];
    a.pos_den[2 * s] = v[3];
    a.pos_num[2 * s + 1] = v[4];
    a.pos_den[2 * s + 1] = v[5];
  }
}

template <int ND>
static int run_chunk(const RpieDev& a, float2* wave, float* sums, long s0, long count,
                     bool need_back, int sms, cudaStream_t st, const char* who) {
  using Cfg = LargeCfg<ND>;
  auto k1 = large_exit_cols_kernel<ND>;
  auto k2 = large_rows_modulus_kernel<ND>;
  auto k3 = large_cols_gradient_kernel<ND>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_col);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_row);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_grad);
    if (e != cudaSuccess) return set_error((int)e, "%s: kernel attributes: %s", who, cudaGetErrorString(e));
    configured = true;
  }
  const int M = a.b.nmodes;
  const long t1 = count * M * (ND / Cfg::VC), t2 = count * (ND / Cfg::VR), t3 = count * (ND / Cfg::VC);
  long g1 = t1 < (long)sms * Cfg::CTAS_K1 ? t1 : (long)sms * Cfg::CTAS_K1;
  long g2 = t2 < (long)sms * 4 ? t2 : (long)sms * 4;
  long g3 = t3 < (long)sms * Cfg::CTAS_K3 ? t3 : (long)sms * Cfg::CTAS_K3;
  const bool reg13 = k13_reg_applies(a);  // register-resident K1 / K3 (large_k13r.cu)
  int rc;
  if (reg13) {
    rc = launch_k1_reg(a, wave, s0, count, sms, st, who);
  } else {
    k1<<<(unsigned)g1, Cfg::NTC, Cfg::smem_col, st>>>(a, wave, s0, count);
    rc = check_launch(who);
  }
  if (rc != TB_OK) return rc;
  if (k2_reg_applies(a)) {
    rc = launch_k2_reg(a, wave, s0, count, need_back, sms, st, who);
  } else {
    k2<<<(unsigned)g2, Cfg::NTR, Cfg::smem_row, st>>>(a, wave, s0, count, need_back ? 1 : 0);
    rc = check_launch(who);
  }
  if (rc != TB_OK || !need_back) return rc;
  const bool want_sums = a.eig_step || a.pos_num;
  if (want_sums) {
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)count * 6 * sizeof(float), st);
    if (e != cudaSuccess) return set_error((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
  }
  if (reg13) {
    rc = launch_k3_reg(a, wave, s0, count, sms, st, who);
  } else {
    k3<<<(unsigned)g3, Cfg::NTC, Cfg::smem_grad, st>>>(a, wave, sums, s0, count);
    rc = check_launch(who);
  }
  if (rc != TB_OK) return rc;
  if (want_sums) {
    large_finalize_sums_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(a, sums, s0, count);
    rc = check_launch(who);
  }
  return rc;
}

// Gaussian-model chunk through the fused pipeline; `sums` holds count * 6 floats.
int run_large_fused_chunk(const RpieDev& a, float2* wave, float* sums, long s0, long count,
                          bool need_back, int sms, cudaStream_t st, const char* who) {
  switch (a.b.detector_width) {
    case 256:  return run_chunk<256>(a, wave, sums, s0, count, need_back, sms, st, who);
    case 512:  return run_chunk<512>(a, wave, sums, s0, count, need_back, sms, st, who);
    case 1024: return run_chunk<1024>(a, wave, sums, s0, count, need_back, sms, st, who);
    case 2048: return run_chunk<2048>(a, wave, sums, s0, count, need_back, sms, st, who);
    default:
      return set_error(TB_ERR_UNSUPPORTED, "%s: detector width %d", who, a.b.detector_width);
  }
}

}  // namespace tb
