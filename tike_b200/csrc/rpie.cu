// Fused rPIE batch kernel and its per-epoch companions.
//
// One CTA owns one scan position at a time (persistent, grid-strided over the
// batch) and keeps the ND x ND complex64 wavefront in shared memory for the
// whole chain  patch -> x probe -> FFT -> intensity -> cost -> modulus /
// Poisson step -> inverse FFT -> conj(probe).chi and conj(patch).chi, so the
// exit wave never visits HBM.  Replaces the ~60 CuPy launches per 64-pattern
// chunk of rpie._get_nearplane_gradients
// (src/tike/ptycho/solvers/rpie.py:315-567) and objective.py:11-124,
// exitwave.py:122-234.
//
// Modes are processed sequentially in two sweeps (intensity first, then the
// gradient with the forward transform recomputed) because M wavefronts do not
// fit on one SM; see DESIGN.md for the smem / register budget.
#include <cstdlib>

#include "solver_dev.cuh"

namespace tb {

template <int ND> struct RpieCfg {
  static constexpr int NT = (ND >= 128) ? 512 : (ND >= 64 ? 256 : (ND >= 32 ? 128 : 64));
  static constexpr int PER_SM = (ND >= 128) ? 1 : (ND >= 64 ? 4 : 8);
  static constexpr int KMAX = ND * ND / NT;  // owned pixels per thread
  static constexpr size_t smem = (size_t)ND * (ND + 1) * 8 + ND * ND * 4 +
                                 ND * 8 + ND * 4 + 4 * 32 * 4;
};

// Per-CTA scratch in global memory (L2 resident): the interpolated patch of
// the current position, the far-field waves of all modes (so the forward FFT
// is not recomputed in the gradient sweep) and the private probe numerator.
struct CtaScratch {
  float2* patch;    // N * N
  float2* waves;    // M * ND * ND, digit-reversed tile order, unscaled
  float2* replica;  // M * N * N or nullptr
};

__host__ __device__ inline long scratch_elems(int M, int N, int ND) {
  return (long)N * N + (long)M * ND * ND;
}

// FAST = the headline configuration, resolved at compile time: probe width ==
// detector width (no padding), shared probe (no per-position weights),
// Gaussian noise model, no eigen-weight / position-gradient outputs.  Every
// index is then a shift/mask of (tid + k*NT) and every inner loop is a few
// instructions per pixel.  Everything else runs the GENERIC variant.
template <int ND, bool FAST>
__global__ void __launch_bounds__(RpieCfg<ND>::NT, (ND >= 128) ? 1 : 2)
rpie_batch_kernel(RpieDev a) {
  using Cfg = RpieCfg<ND>;
  constexpr int NT = Cfg::NT, KMAX = Cfg::KMAX, P = ND + 1, LG = Log2<ND>::v;
  constexpr int NWARP = NT / 32;
  constexpr int KP = KMAX / 2;  // pixel pairs per thread (FAST variant)
  constexpr int LB = KP >= 4 ? 4 : KP;  // global loads issued back to back
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float* F = reinterpret_cast<float*>(tile + ND * P);
  float2* tw = reinterpret_cast<float2*>(F + ND * ND);
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  unsigned short* f2l = l2f + ND;
  float* red = reinterpret_cast<float*>(f2l + ND);
  fill_twiddles<ND>(tw);
  fill_perm<ND>(l2f, f2l);
  __syncthreads();

  const tb_batch& b = a.b;
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const int N = FAST ? ND : b.probe_width;
  const int M = b.nmodes;
  const int pad = FAST ? 0 : (ND - N) / 2;
  const int H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;  // round-trip normalisation
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool gaussian = FAST || a.noise_model == TB_NOISE_GAUSSIAN;
  const bool need_back = a.accumulate_object || a.probe_sums || a.eig_step || a.chi_out || a.pos_num;
  [[maybe_unused]] const uint64_t pol_keep = l2_policy_evict_last();
  [[maybe_unused]] const uint64_t pol_stream = l2_policy_evict_first();

  CtaScratch sc;
  {
    const long per_cta = scratch_elems(M, N, ND);
    float2* base = a.scratch + (long)blockIdx.x * per_cta;
    sc.patch = base;
    sc.waves = base + (long)N * N;
    sc.replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * N * N : nullptr;
  }

  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner c = make_corner(b.scan, s);
    const long dbase = s * (long)ND * ND;

    // ---------------- patch, once per position -----------------------------
    // FAST: every thread owns KP pixel pairs (2q, 2q+1), q = tid + k*NT; the
    // pairs stay in registers for sweep 1 and go to scratch for sweep 2.
    [[maybe_unused]] float4 o2[FAST ? KP : 1];
    {
      const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < H) & (c.ix + N < W);
      if constexpr (FAST) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const int q = tid + k * NT, l0 = 2 * q;
          const int py = l0 >> LG, px = l0 & (ND - 1);
          float2 oa, ob;
          if (interior) {
            const float2* r0 = psi + (long)(c.iy + py) * W + c.ix + px;
            const float2 a0 = __ldg(r0), a1 = __ldg(r0 + 1), a2 = __ldg(r0 + 2);
            const float2 b0 = __ldg(r0 + W), b1 = __ldg(r0 + W + 1), b2 = __ldg(r0 + W + 2);
            oa.x = a0.x * c.w00; oa.y = a0.y * c.w00;
            oa.x += a1.x * c.w01; oa.y += a1.y * c.w01;
            oa.x += b0.x * c.w10; oa.y += b0.y * c.w10;
            oa.x += b1.x * c.w11; oa.y += b1.y * c.w11;
            ob.x = a1.x * c.w00; ob.y = a1.y * c.w00;
            ob.x += a2.x * c.w01; ob.y += a2.y * c.w01;
            ob.x += b1.x * c.w10; ob.y += b1.y * c.w10;
            ob.x += b2.x * c.w11; ob.y += b2.y * c.w11;
          } else {
            oa = patch_value(psi, H, W, c, py, px);
            ob = patch_value(psi, H, W, c, py, px + 1);
          }
          o2[k] = make_float4(oa.x, oa.y, ob.x, ob.y);
          __stcg(reinterpret_cast<float4*>(sc.patch) + q, o2[k]);
        }
      } else {
        for (int py = warp; py < N; py += NWARP) {
          const float2* r0 = psi + (long)(c.iy + py) * W + c.ix;
          for (int px = lane; px < N; px += 32) {
            float2 o;
            if (interior) {
              const float2 v00 = __ldg(r0 + px), v01 = __ldg(r0 + px + 1);
              const float2 v10 = __ldg(r0 + W + px), v11 = __ldg(r0 + W + px + 1);
              o.x = v00.x * c.w00; o.y = v00.y * c.w00;
              o.x += v01.x * c.w01; o.y += v01.y * c.w01;
              o.x += v10.x * c.w10; o.y += v10.y * c.w10;
              o.x += v11.x * c.w11; o.y += v11.y * c.w11;
            } else {
              o = patch_value(psi, H, W, c, py, px);
            }
            __stcg(sc.patch + py * N + px, o);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) F[tid + k * NT] = 0.f;
    __syncthreads();  // patch visible to the whole CTA

    // ---------------- sweep 1: far field of every mode, intensity ----------
    for (int m = 0; m < M; ++m) {
      if constexpr (FAST) {
        const float4* __restrict__ pm2 = reinterpret_cast<const float4*>(ps.probe + (long)m * ND * ND);
#pragma unroll
        for (int k0 = 0; k0 < KP; k0 += LB) {
          float4 p[LB];  // issue the whole batch of loads before using any
#pragma unroll
          for (int j = 0; j < LB; ++j) p[j] = __ldg(pm2 + tid + (k0 + j) * NT);
#pragma unroll
          for (int j = 0; j < LB; ++j) {
            const int l0 = 2 * (tid + (k0 + j) * NT);
            float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
            t[0] = cmul(make_float2(p[j].x, p[j].y), make_float2(o2[k0 + j].x, o2[k0 + j].y));
            t[1] = cmul(make_float2(p[j].z, p[j].w), make_float2(o2[k0 + j].z, o2[k0 + j].w));
          }
        }
      } else {
        for (int idx = tid; idx < ND * ND; idx += NT) {
          const int ly = idx >> LG, lx = idx & (ND - 1);
          const int py = ly - pad, px = lx - pad;
          float2 v = make_float2(0.f, 0.f);
          if (py >= 0 && py < N && px >= 0 && px < N)
            v = cmul(probe_value(ps, s, m, py, px), __ldcg(sc.patch + py * N + px));
          tile[ly * P + lx] = v;
        }
      }
      __syncthreads();
      fft2_tile<ND, false>(tile, tw);
      float2* wave = sc.waves + (long)m * ND * ND;
      if constexpr (FAST) {
        float2* F2 = reinterpret_cast<float2*>(F);
        if (need_back) {
#pragma unroll 8
          for (int k = 0; k < KP; ++k) {
            const int q = tid + k * NT, l0 = 2 * q;
            const float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
            const float2 w0 = t[0], w1 = t[1];
            float2 f = F2[q];
            f.x += cabs2(w0) * s2;
            f.y += cabs2(w1) * s2;
            F2[q] = f;
            st_f32x4_hint(reinterpret_cast<float4*>(wave) + q,
                          make_float4(w0.x, w0.y, w1.x, w1.y), pol_keep);
          }
        } else {
#pragma unroll 8
          for (int k = 0; k < KP; ++k) {
            const int q = tid + k * NT, l0 = 2 * q;
            const float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
            float2 f = F2[q];
            f.x += cabs2(t[0]) * s2;
            f.y += cabs2(t[1]) * s2;
            F2[q] = f;
          }
        }
      } else {
#pragma unroll 8
        for (int k = 0; k < KMAX; ++k) {
          const int l = tid + k * NT;
          const float2 w = tile[(l >> LG) * P + (l & (ND - 1))];
          F[l] += cabs2(w) * s2;
          if (need_back) __stcg(wave + l, w);
        }
      }
      __syncthreads();
    }

    // ---------------- cost, modulus factor / Poisson step ----------------
    float step_dom = a.step_start;
    {
      float sums[3] = {0.f, 0.f, 0.f};  // cost, (poisson dominant) denom, numer
#pragma unroll 4
      for (int k = 0; k < KMAX; ++k) {
        const int l = tid + k * NT;
        const int pix = (int)l2f[l >> LG] * ND + (int)l2f[l & (ND - 1)];
        const bool meas = a.mask ? (a.mask[pix] != 0) : true;
        const float I = F[l];
        if (meas) {
          const float d = load_data_stream(a.data, a.data_u16, dbase + pix, pol_stream);
          if (gaussian) {
            const float sd = sqrtf(d), sI = sqrtf(I);
            const float t = sI - sd;
            sums[0] += t * t;
            F[l] = -(1.0f - sd / (sI + 1e-9f));
          } else {
            sums[0] += I - d * logf(I + 1e-9f);
            if (a.step_mode == TB_STEP_DOMINANT_MODE) {
              const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
              sums[1] += xi * xi * I;
              sums[2] += xi * (I - d / (1.0f - step_dom * xi));
            }
          }
        } else if (gaussian) {
          F[l] = a.unmeasured_factor;
        }
      }
      block_sum<3>(sums, red);
      if (tid == 0) a.costs[s] = sums[0] * a.inv_nmeasured;
      if constexpr (!FAST) {
        if (!gaussian && a.step_mode == TB_STEP_DOMINANT_MODE) {
          // exitwave.py:183-234, two fixed-point iterations
          step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (sums[2] / sums[1]);
          float s1[1] = {0.f};
          for (int l = tid; l < ND * ND; l += NT) {
            const int pix = (int)l2f[l >> LG] * ND + (int)l2f[l & (ND - 1)];
            const bool meas = a.mask ? (a.mask[pix] != 0) : true;
            if (meas) {
              const float d = load_data(a.data, a.data_u16, dbase + pix);
              const float I = F[l];
              const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
              s1[0] += xi * (I - d / (1.0f - step_dom * xi));
            }
          }
          block_sum<1>(s1, red);
          step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (s1[0] / sums[1]);
        }
      }
    }
    if (!need_back) { __syncthreads(); continue; }

    // ---------------- sweep 2: gradients ----------------------------------
    float2 acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] = make_float2(0.f, 0.f);
    float eig[2] = {0.f, 0.f};
    float pg[4] = {0.f, 0.f, 0.f, 0.f};  // position gradient sums (lstsq)
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      const float2* wave = sc.waves + (long)m * ND * ND;
      if constexpr (FAST) {
        // reload the far field and apply the modulus factor in one pass
        const float2* F2 = reinterpret_cast<const float2*>(F);
#pragma unroll
        for (int k0 = 0; k0 < KP; k0 += LB) {
          float4 w[LB];
#pragma unroll
          for (int j = 0; j < LB; ++j)
            w[j] = ld_f32x4_hint(reinterpret_cast<const float4*>(wave) + tid + (k0 + j) * NT, pol_keep);
#pragma unroll
          for (int j = 0; j < LB; ++j) {
            const int q = tid + (k0 + j) * NT, l0 = 2 * q;
            const float2 f = F2[q];
            float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
            t[0] = cscale(make_float2(w[j].x, w[j].y), f.x * rt);
            t[1] = cscale(make_float2(w[j].z, w[j].w), f.y * rt);
          }
        }
      } else if (gaussian) {
#pragma unroll 8
        for (int k = 0; k < KMAX; ++k) {
          const int l = tid + k * NT;
          tile[(l >> LG) * P + (l & (ND - 1))] = cscale(__ldcg(wave + l), F[l] * rt);
        }
      } else {
        for (int l = tid; l < ND * ND; l += NT)
          tile[(l >> LG) * P + (l & (ND - 1))] = __ldcg(wave + l);
        float step = step_dom;
        if (a.step_mode == TB_STEP_ALL_MODES) {
          // exitwave.py:122-180 for this mode (each thread re-reads only the
          // tile entries it wrote, no barrier needed)
          step = a.step_start;
          float q0 = 0.f;
          for (int it = 0; it < 2; ++it) {
            float q[2] = {0.f, 0.f};  // denom_final, numer
            for (int l = tid; l < ND * ND; l += NT) {
              const int pix = (int)l2f[l >> LG] * ND + (int)l2f[l & (ND - 1)];
              const bool meas = a.mask ? (a.mask[pix] != 0) : true;
              if (meas) {
                const float d = load_data(a.data, a.data_u16, dbase + pix);
                const float I = F[l];
                const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
                const float ab = cabs2(tile[(l >> LG) * P + (l & (ND - 1))]) * s2;
                const float t = xi * step - 1.0f;
                const float den = ab * t * t + I - ab;
                q[0] += xi * xi * ab;
                q[1] += xi * ab * (1.0f + (d * t) / den);
              }
            }
            block_sum<2>(q, red);
            if (it == 0) q0 = q[0];
            step = step * (1.0f - a.step_weight) + (q[1] / q0) * a.step_weight;
          }
        }
        for (int l = tid; l < ND * ND; l += NT) {
          const int pix = (int)l2f[l >> LG] * ND + (int)l2f[l & (ND - 1)];
          const bool meas = a.mask ? (a.mask[pix] != 0) : true;
          float f = a.unmeasured_factor;
          if (meas) {
            const float d = load_data(a.data, a.data_u16, dbase + pix);
            const float I = F[l];
            const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
            f = -step * xi;
          }
          float2& w = tile[(l >> LG) * P + (l & (ND - 1))];
          w = cscale(w, f * rt);
        }
      }
      __syncthreads();
      fft2_tile<ND, true>(tile, tw);
      // chi = tile[pad:pad+N, pad:pad+N]
      if constexpr (FAST) {
        const float4* __restrict__ pm2 = reinterpret_cast<const float4*>(ps.probe + (long)m * ND * ND);
        const float4* patch2 = reinterpret_cast<const float4*>(sc.patch);
        float2* rep = sc.replica ? sc.replica + (long)m * ND * ND : nullptr;
        float4* cout = a.chi_out ? reinterpret_cast<float4*>(a.chi_out + ((long)s * M + m) * ND * ND) : nullptr;
        // three passes over chi (cheap LDS) so that each pass can issue its
        // global loads in batches without per-iteration branches
        if (a.accumulate_object) {
#pragma unroll
          for (int k0 = 0; k0 < KP; k0 += LB) {
            float4 p[LB];
#pragma unroll
            for (int j = 0; j < LB; ++j) p[j] = __ldg(pm2 + tid + (k0 + j) * NT);
#pragma unroll
            for (int j = 0; j < LB; ++j) {
              const int k = k0 + j, l0 = 2 * (tid + k * NT);
              const float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
              const float2 g0 = cmulc(make_float2(p[j].x, p[j].y), t[0]);
              const float2 g1 = cmulc(make_float2(p[j].z, p[j].w), t[1]);
              acc[2 * k].x += g0.x; acc[2 * k].y += g0.y;
              acc[2 * k + 1].x += g1.x; acc[2 * k + 1].y += g1.y;
            }
          }
        }
        if (rep) {
#pragma unroll
          for (int k0 = 0; k0 < KP; k0 += LB) {
            float4 o[LB];
#pragma unroll
            for (int j = 0; j < LB; ++j) o[j] = __ldcg(patch2 + tid + (k0 + j) * NT);
#pragma unroll
            for (int j = 0; j < LB; ++j) {
              const int l0 = 2 * (tid + (k0 + j) * NT);
              const float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
              red_add_f32x4(rep + l0, cmulc(make_float2(o[j].x, o[j].y), t[0]),
                            cmulc(make_float2(o[j].z, o[j].w), t[1]));
            }
          }
        }
        if (cout) {
#pragma unroll 4
          for (int k = 0; k < KP; ++k) {
            const int q = tid + k * NT, l0 = 2 * q;
            const float2* t = tile + (l0 >> LG) * P + (l0 & (ND - 1));
            const float2 chi0 = t[0], chi1 = t[1];
            cout[q] = make_float4(chi0.x, chi0.y, chi1.x, chi1.y);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const int idx = tid + k * NT;
          if (idx < N * N) {
            const int py = idx / N, px = idx - py * N;
            const float2 chi = tile[(pad + py) * P + pad + px];
            if (a.chi_out) a.chi_out[((long)s * M + m) * N * N + idx] = chi;
            const float2 o = __ldcg(sc.patch + idx);
            if (a.accumulate_object) {
              const float2 p = probe_value(ps, s, m, py, px);
              const float2 g = cmulc(p, chi);
              acc[k].x += g.x;
              acc[k].y += g.y;
            }
            if (sc.replica)
              red_add_f32x2(sc.replica + (long)m * N * N + idx, cmulc(o, chi));
            if (m == 0 && a.pos_num) {
              // lstsq.py:545-579 on the centre crop [N/4, N - N/4)
              const int crop = N / 4;
              if (py >= crop && py < N - crop && px >= crop && px < N - crop) {
                float2 gy = make_float2(0.f, 0.f), gx = make_float2(0.f, 0.f);
#pragma unroll
                for (int t = -2; t <= 2; ++t) {
                  const float wt = a.taps[t + 2];
                  const int qy = min(max(py + t, 0), N - 1), qx = min(max(px + t, 0), N - 1);
                  const float2 oy = __ldcg(sc.patch + qy * N + px);
                  const float2 ox = __ldcg(sc.patch + py * N + qx);
                  gy.x -= wt * oy.x; gy.y -= wt * oy.y;
                  gx.x -= wt * ox.x; gx.y -= wt * ox.y;
                }
                const float2 p0u = probe_value(ps, s, 0, py, px);
                const float2 ay = cmul(gy, p0u), ax = cmul(gx, p0u);
                pg[0] += ay.x * chi.x + ay.y * chi.y;
                pg[1] += cabs2(ay);
                pg[2] += ax.x * chi.x + ax.y * chi.y;
                pg[3] += cabs2(ax);
              }
            }
            if (m == 0 && a.eig_step) {
              // rpie.py:493-506 / lstsq.py:721-736: shared probe mode 0
              const float2 p0 = __ldg(ps.probe + (ps.per_position ? s * (long)M * N * N : 0) + idx);
              const float2 op = cmul(o, p0);
              eig[0] += op.x * chi.x + op.y * chi.y;
              eig[1] += cabs2(op);
            }
          }
        }
      }
      __syncthreads();
    }
    if constexpr (!FAST) {
      if (a.eig_step) {
        block_sum<2>(eig, red);
        if (tid == 0) a.eig_step[s] = 0.1f * (eig[0] / eig[1]);
      }
      if (a.pos_num) {
        block_sum<4>(pg, red);
        if (tid == 0) {
          a.pos_num[2 * s] = pg[0];
          a.pos_den[2 * s] = pg[1];
          a.pos_num[2 * s + 1] = pg[2];
          a.pos_den[2 * s + 1] = pg[3];
        }
      }
    }

    // ---------------- scatter-add of the object gradient ------------------
    if (a.accumulate_object) {
      float2* G = tile;  // N x N, pitch N
      const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        // FAST: acc[2k], acc[2k+1] belong to pixels 2q, 2q+1 with q = tid + k*NT
        const int idx = FAST ? 2 * (tid + (k >> 1) * NT) + (k & 1) : tid + k * NT;
        if (FAST || idx < N * N) {
          const int py = FAST ? (idx >> LG) : idx / N;
          const int px = idx - py * N;
          const int y = c.iy + py, x = c.ix + px;
          const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
          G[idx] = lead_ok ? cscale(acc[k], inv_m) : make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      for (int ty = warp; ty <= N; ty += NWARP) {
        const int y = c.iy + ty;
        if (y < 0 || y >= H) continue;
        const bool a0 = ty < N, a1 = ty > 0;
        for (int tx = lane; tx <= N; tx += 32) {
          const int x = c.ix + tx;
          if (x < 0 || x >= W) continue;
          float2 v = make_float2(0.f, 0.f);
          const bool b0 = tx < N, b1 = tx > 0;
          if (a0 & b0) { const float2 g = G[ty * N + tx];           v.x += c.w00 * g.x; v.y += c.w00 * g.y; }
          if (a0 & b1) { const float2 g = G[ty * N + tx - 1];       v.x += c.w01 * g.x; v.y += c.w01 * g.y; }
          if (a1 & b0) { const float2 g = G[(ty - 1) * N + tx];     v.x += c.w10 * g.x; v.y += c.w10 * g.y; }
          if (a1 & b1) { const float2 g = G[(ty - 1) * N + tx - 1]; v.x += c.w11 * g.x; v.y += c.w11 * g.y; }
          red_add_f32x2(a.psi_num + (long)y * W + x, v);
        }
      }
    }
    __syncthreads();
  }
}

// Sum the private probe numerators of all CTAs (stride = per-CTA scratch).
__global__ void __launch_bounds__(256)
reduce_replicas_kernel(const float2* __restrict__ rep, int R, long stride, long n,
                       float2* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    float2 a = make_float2(0.f, 0.f);
    for (int r = 0; r < R; ++r) {
      const float2 v = rep[(long)r * stride + i];
      a.x += v.x;
      a.y += v.y;
    }
    out[i] = a;
  }
}

__global__ void __launch_bounds__(256)
zero_replicas_kernel(float2* __restrict__ rep, int R, long stride, long n) {
  const long total = (long)R * n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / n;
    rep[r * stride + (i - r * n)] = make_float2(0.f, 0.f);
  }
}

// max over the real parts of n complex values (preconditioners are >= 0 with
// zero imaginary part, so this equals cupy's complex .max()).
__global__ void __launch_bounds__(256)
max_real_kernel(const float2* __restrict__ x, long n, float* __restrict__ out) {
  float m = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));  // m >= 0
}

__global__ void __launch_bounds__(256)
rpie_update_psi_kernel(float2* __restrict__ psi, const float2* __restrict__ num,
                       const float2* __restrict__ precond, long n, float alpha,
                       const float* __restrict__ maxv) {
  const float mx = *maxv;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    // complex division by the complex-typed denominator (imag = 0)
    const float2 pc = precond[i];
    const float dr = (1.0f - alpha) * pc.x + alpha * mx;
    const float di = (1.0f - alpha) * pc.y;
    const float2 g = num[i];
    float2 q;
    if (di == 0.f) {
      q = make_float2(g.x / dr, g.y / dr);
    } else {
      const float dd = dr * dr + di * di;
      q = make_float2((g.x * dr + g.y * di) / dd, (g.y * dr - g.x * di) / dd);
    }
    float2 p = psi[i];
    p.x += q.x;
    p.y += q.y;
    psi[i] = p;
  }
}

__global__ void __launch_bounds__(256)
rpie_update_probe_kernel(float2* __restrict__ probe, const float2* __restrict__ num,
                         long n, float alpha, const float* __restrict__ maxv) {
  const float deno = alpha * (*maxv);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    float2 p = probe[i];
    const float2 g = num[i];
    p.x += g.x / deno;
    p.y += g.y / deno;
    probe[i] = p;
  }
}

template <int ND, bool FAST>
int launch_rpie_variant(const RpieDev& a, int grid, cudaStream_t st) {
  auto k = rpie_batch_kernel<ND, FAST>;
  const size_t smem = RpieCfg<ND>::smem;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error((int)e, "rpie kernel attr: %s", cudaGetErrorString(e));
  k<<<(unsigned)grid, RpieCfg<ND>::NT, smem, st>>>(a);
  return check_launch("tb_rpie_batch");
}

template <int ND>
int launch_rpie(const RpieDev& a, int grid, cudaStream_t st) {
  const tb_batch& b = a.b;
  const bool fast = b.probe_width == ND && b.eigen_weights == nullptr &&
                    !b.probe_per_position && a.noise_model == TB_NOISE_GAUSSIAN &&
                    a.eig_step == nullptr && a.pos_num == nullptr;
  return fast ? launch_rpie_variant<ND, true>(a, grid, st)
              : launch_rpie_variant<ND, false>(a, grid, st);
}

static int fused_grid(int nd, long npos) {
  int sms = 148;
  tb_sm_count(&sms);
  const int per_sm = (nd >= 128) ? 1 : (nd >= 64 ? 4 : 8);
  long g = (long)sms * per_sm;
  if (npos < g) g = npos;
  return (int)(g < 1 ? 1 : g);
}

int64_t fused_workspace_bytes(const tb_batch& b, bool replica) {
  const int grid = fused_grid(b.detector_width, b.npos);
  const int nrep = grid < kMaxReplicas ? grid : kMaxReplicas;
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  return ((int64_t)grid * scratch_elems(b.nmodes, b.probe_width, b.detector_width) +
          (replica ? (int64_t)nrep * n : 0)) * 8 + 64;  // + the position ticket
}

int run_fused(RpieDev a, int64_t workspace_bytes, void* workspace,
              float2* probe_out, cudaStream_t st, const char* who) {
  const tb_batch& b = a.b;
  const int nd = b.detector_width;
  TB_REQUIRE(nd == 16 || nd == 32 || nd == 64 || nd == 128, TB_ERR_UNSUPPORTED,
             "%s: fused kernel supports detector widths 16..128, got %d", who, nd);
  const int grid = fused_grid(nd, b.npos);
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  if (!probe_out) a.probe_sums = 0;
  const bool replica = a.probe_sums != 0;
  const int64_t need = fused_workspace_bytes(b, replica);
  TB_REQUIRE(workspace && workspace_bytes >= need, TB_ERR_INVALID,
             "%s: workspace too small (%lld < %lld bytes)", who,
             (long long)workspace_bytes, (long long)need);
  a.scratch = (float2*)workspace;
  {
    const char* e = getenv("TB_PREFETCH_NEXT");
    a.prefetch_next = e ? atoi(e) : 3;  // bit 0: next position, bit 1: next spilled wave
  }
#ifdef TB_PHASE_TIMING
  {
    // development only: drop one of the gradient outputs to see what it costs
    const char* e = getenv("TB_DEBUG_DROP");
    const int drop = e ? atoi(e) : 0;
    if (drop & 1) a.probe_sums = 0;
    if (drop & 2) a.accumulate_object = 0;
  }
#endif
  a.nrep = grid < kMaxReplicas ? grid : kMaxReplicas;
  a.replicas = a.scratch + (long)grid * scratch_elems(b.nmodes, b.probe_width, nd);
  if (replica) {
    cudaError_t e = cudaMemsetAsync(a.replicas, 0, (size_t)a.nrep * n * 8, st);
    if (e != cudaSuccess) return set_error((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
  }
  {
    // The persistent CTAs of the stage-fused kernel draw their positions from
    // a counter: a CTA that gets its SM late (e.g. behind a communication
    // kernel of the multi-GPU exchange) then simply takes fewer positions.
    a.ticket = (unsigned int*)(a.replicas + (replica ? (long)a.nrep * n : 0));
    cudaError_t e = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return set_error((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
  }
  int rc;
  // TB_RPIE_P3=0 keeps the plain 128 x 128 batch on rpie_fast_kernel (A/B timing)
  static const bool use_p3 = [] {
    const char* e = getenv("TB_RPIE_P3");
    return e ? atoi(e) != 0 : true;
  }();
  if (use_p3 && p3_kernel_applies(a)) {
    rc = launch_p3(a, grid, st);
  } else if (fast_kernel_applies(a)) {
    rc = launch_fast(a, grid, st);
  } else
  switch (nd) {
    case 16:  rc = launch_rpie<16>(a, grid, st); break;
    case 32:  rc = launch_rpie<32>(a, grid, st); break;
    case 64:  rc = launch_rpie<64>(a, grid, st); break;
    default:  rc = launch_rpie<128>(a, grid, st); break;
  }
  if (rc != TB_OK) return rc;
  if (replica) {
    const long blocks = (n + 255) / 256;
    reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
        a.replicas, a.nrep, n, n, probe_out);
    rc = check_launch("reduce_replicas");
  }
  return rc;
}

}  // namespace tb

extern "C" {

int64_t tb_rpie_workspace_size(const tb_rpie_args* a) {
  if (!a) return 0;
  if (!tb::fused_width(a->batch.detector_width))
    return tb::large_workspace_bytes(a->batch, a->accumulate_object != 0, a->noise_model);
  return tb::fused_workspace_bytes(a->batch, a->accumulate_object != 0);
}

int tb_rpie_batch(const tb_rpie_args* a, tb_stream_t stream) {
  TB_REQUIRE(a != nullptr, TB_ERR_INVALID, "tb_rpie_batch: null args");
  int rc = tb::check_batch(&a->batch, "tb_rpie_batch");
  if (rc != TB_OK) return rc;
  if (a->batch.npos == 0) return TB_OK;
  TB_REQUIRE(a->data && a->costs, TB_ERR_INVALID, "tb_rpie_batch: null data/costs");
  TB_REQUIRE(!a->accumulate_object || (a->psi_numerator && a->probe_numerator),
             TB_ERR_INVALID, "tb_rpie_batch: numerators required");
  TB_REQUIRE(a->noise_model == TB_NOISE_GAUSSIAN || a->noise_model == TB_NOISE_POISSON,
             TB_ERR_INVALID, "tb_rpie_batch: unknown noise model %d", a->noise_model);
  TB_REQUIRE(a->num_measured > 0, TB_ERR_INVALID, "tb_rpie_batch: num_measured must be > 0");
  if (a->batch.npos == 0) return TB_OK;
  tb::RpieDev d{};
  d.b = a->batch;
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
  d.accumulate_object = a->accumulate_object;
  d.divide_by_modes = 1;
  d.psi_num = (float2*)a->psi_numerator;
  d.costs = a->costs;
  d.eig_step = a->eigen_weight_step;
  d.probe_sums = a->accumulate_object;
  if (!tb::fused_width(d.b.detector_width))
    return tb::run_large(d, a->workspace_bytes, a->workspace, (float2*)a->probe_numerator,
                         (cudaStream_t)stream, "tb_rpie_batch");
  return tb::run_fused(d, a->workspace_bytes, a->workspace, (float2*)a->probe_numerator,
                       (cudaStream_t)stream, "tb_rpie_batch");
}

int64_t tb_lstsq_workspace_size(const tb_lstsq_args* a) {
  if (!a) return 0;
  if (!tb::fused_width(a->batch.detector_width))
    return tb::large_workspace_bytes(a->batch, a->recover_probe != 0, a->noise_model);
  return tb::fused_workspace_bytes(a->batch, a->recover_probe != 0);
}

int tb_lstsq_phase1(const tb_lstsq_args* a, tb_stream_t stream) {
  TB_REQUIRE(a != nullptr, TB_ERR_INVALID, "tb_lstsq_phase1: null args");
  int rc = tb::check_batch(&a->batch, "tb_lstsq_phase1");
  if (rc != TB_OK) return rc;
  if (a->batch.npos == 0) return TB_OK;
  TB_REQUIRE(a->data && a->costs && a->chi, TB_ERR_INVALID,
             "tb_lstsq_phase1: null data/costs/chi");
  TB_REQUIRE(!a->recover_psi || a->object_upd_sum, TB_ERR_INVALID,
             "tb_lstsq_phase1: object_upd_sum required");
  TB_REQUIRE(!a->recover_probe || a->probe_upd_sum, TB_ERR_INVALID,
             "tb_lstsq_phase1: probe_upd_sum required");
  TB_REQUIRE(!a->recover_positions || (a->position_num && a->position_den),
             TB_ERR_INVALID, "tb_lstsq_phase1: position buffers required");
  TB_REQUIRE(a->noise_model == TB_NOISE_GAUSSIAN || a->noise_model == TB_NOISE_POISSON,
             TB_ERR_INVALID, "tb_lstsq_phase1: unknown noise model %d", a->noise_model);
  TB_REQUIRE(a->num_measured > 0, TB_ERR_INVALID, "tb_lstsq_phase1: num_measured must be > 0");
  if (a->batch.npos == 0) return TB_OK;
  tb::RpieDev d{};
  d.b = a->batch;
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
  d.accumulate_object = a->recover_psi ? 1 : 0;
  d.divide_by_modes = 0;
  d.psi_num = (float2*)a->object_upd_sum;
  d.costs = a->costs;
  d.chi_out = (float2*)a->chi;
  d.poisson_eps = 1;
  if (a->recover_positions) {
    d.pos_num = a->position_num;
    d.pos_den = a->position_den;
    for (int i = 0; i < 5; ++i) d.taps[i] = a->gradient_taps[i];
  }
  d.probe_sums = a->recover_probe ? 1 : 0;
  if (!tb::fused_width(d.b.detector_width))
    return tb::run_large(d, a->workspace_bytes, a->workspace,
                         a->recover_probe ? (float2*)a->probe_upd_sum : nullptr,
                         (cudaStream_t)stream, "tb_lstsq_phase1");
  return tb::run_fused(d, a->workspace_bytes, a->workspace,
                       a->recover_probe ? (float2*)a->probe_upd_sum : nullptr,
                       (cudaStream_t)stream, "tb_lstsq_phase1");
}

int tb_rpie_update_psi(void* psi, const void* numerator, const void* precond,
                       int64_t n, float alpha, float* scratch, tb_stream_t stream) {
  TB_REQUIRE(psi && numerator && precond && scratch, TB_ERR_INVALID,
             "tb_rpie_update_psi: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, sizeof(float), st);
  const long blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 2368 ? blocks : 2368);
  tb::max_real_kernel<<<grid, 256, 0, st>>>((const float2*)precond, n, scratch);
  tb::rpie_update_psi_kernel<<<grid, 256, 0, st>>>((float2*)psi, (const float2*)numerator,
                                                  (const float2*)precond, n, alpha, scratch);
  return tb::check_launch("tb_rpie_update_psi");
}

int tb_rpie_update_probe(void* probe, const void* numerator, const void* probe_precond,
                         int nmodes, int64_t n2, float alpha, float* scratch,
                         tb_stream_t stream) {
  TB_REQUIRE(probe && numerator && probe_precond && scratch, TB_ERR_INVALID,
             "tb_rpie_update_probe: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, sizeof(float), st);
  const long b2 = (n2 + 255) / 256;
  tb::max_real_kernel<<<(unsigned)(b2 < 1024 ? b2 : 1024), 256, 0, st>>>(
      (const float2*)probe_precond, n2, scratch);
  const long n = n2 * nmodes;
  const long blocks = (n + 255) / 256;
  tb::rpie_update_probe_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, st>>>(
      (float2*)probe, (const float2*)numerator, n, alpha, scratch);
  return tb::check_launch("tb_rpie_update_probe");
}

}  // extern "C"
