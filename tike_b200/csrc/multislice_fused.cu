// Per-position fused rPIE batch for multislice objects (D slices): the whole
// slice loop of one scan position runs in one persistent CTA, with the
// wavefront in shared memory, instead of one launch per slice and pass through
// HBM (multislice.cu keeps that chunked chain for everything this kernel does
// not take).
//
// Replaces, for probe width == detector width in {32, 64, 128}, Gaussian noise,
// shared probe: Multislice.fwd_return_intermediate_probes
// (operators/cupy/multislice.py:97-139), FresnelSpectProp.fwd/adj
// (fresnelspectprop.py:52-111) and the slice loop of
// rpie._get_nearplane_gradients (ptycho/solvers/rpie.py:374-474).
//
// Per position and mode m:
//   forward:  psi_0 = P_m o_0 ;  probe_{t+1} = F^-1[ H F[psi_t] ] ; psi_{t+1} = probe_{t+1} o_{t+1}
//             Psi_m = F[psi_{D-1}] ;  I += |Psi_m|^2
//   backward: chi_{D-1} = F^-1[ factor Psi_m ] ; for t = D-1 .. 0:
//             G_t += conj(probe_t) chi_t / M ;  Q_{t,m} += conj(o_t) chi_t ;
//             chi_{t-1} = F^-1[ conj(H) F[chi_t] ]          (rpie.py:474: no conj(o_t))
// The patches o_t, the probes incident on slices >= 1 and the far fields of all
// but the last mode live in a per-CTA scratch that stays in L2; the D
// object-gradient accumulators live in Tensor Memory; the probe numerators go
// to L2-resident replicas by vector reductions like in rpie_fast.cu.
#include "solver_dev.cuh"
#include "tmem.cuh"

namespace tb {

__device__ __forceinline__ void st_f32x2_keep(float2* addr, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(v.x),
               "f"(v.y), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float2 ld_f32x2_stream(const float2* addr, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y)
               : "l"(addr), "l"(pol));
  return v;
}

template <int ND> struct MsfCfg {
  static constexpr int NT = (ND >= 128) ? 512 : (ND >= 64 ? 256 : 128);
  static constexpr int KMAX = ND * ND / NT;          // pixels per thread
  static constexpr int NWARP = NT / 32;
  static constexpr int COLS_PER_SLICE = 2 * KMAX * ((NWARP + 3) / 4);  // TMEM columns
  static constexpr int MAX_SLICES = 512 / COLS_PER_SLICE;
  static constexpr size_t smem = (size_t)ND * (ND + 1) * 8 + ND * ND * 4 + ND * 8 + ND * 4 +
                                 4 * 32 * 4;
};

// per-CTA scratch (complex values): D patches, (D - 1) * M incident probes, M - 1 far fields
__host__ __device__ inline long msf_scratch_elems(int D, int M, int N) {
  return (long)D * N * N + (long)(D - 1) * M * N * N + (long)(M > 1 ? M - 1 : 1) * N * N;
}

// Hperm[r * N + c] = H[l2f(r) * N + l2f(c)] / N^2: the Fresnel spectrum in the
// digit-reversed slot order the in-place transforms leave a tile in
template <int ND>
__global__ void msf_permute_propagator_kernel(const float2* __restrict__ h,
                                              float2* __restrict__ out) {
  const float inv = 1.0f / ((float)ND * (float)ND);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ND * ND;
       idx += gridDim.x * blockDim.x) {
    const int r = idx / ND, c = idx - r * ND;
    out[idx] = cscale(__ldg(h + loc2freq<ND>(r) * ND + loc2freq<ND>(c)), inv);
  }
}

template <int ND>
__global__ void __launch_bounds__(MsfCfg<ND>::NT, 1)
rpie_ms_fused_kernel(RpieDev a, int D, const float2* __restrict__ hperm,
                     float2* __restrict__ psi_num_all, float2* __restrict__ replicas_all) {
  using Cfg = MsfCfg<ND>;
  constexpr int NT = Cfg::NT, KMAX = Cfg::KMAX, P = ND + 1, LG = Log2<ND>::v;
  constexpr int NWARP = Cfg::NWARP, NCH = KMAX / 8;  // x16 TMEM chunks per slice and thread
  static_assert(KMAX % 8 == 0, "eight complex values per Tensor Memory access");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float* F = reinterpret_cast<float*>(tile + ND * P);
  float2* tw = reinterpret_cast<float2*>(F + ND * ND);
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  unsigned short* f2l = l2f + ND;
  float* red = reinterpret_cast<float*>(f2l + ND);
  __shared__ uint32_t tmem_slot;
  __shared__ long sh_next;
  fill_twiddles<ND>(tw);
  fill_perm<ND>(l2f, f2l);
  constexpr uint32_t TCOLS_WARP = 2 * KMAX;
  const uint32_t need_cols = (uint32_t)D * Cfg::COLS_PER_SLICE;
  const uint32_t tcols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128
                         : (need_cols <= 256 ? 256 : 512)));
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, tcols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // accumulator of slice t: this warp's lane quadrant, D * TCOLS_WARP private columns
  const uint32_t tbase = tmem_slot + ((((uint32_t)warp & 3u) * 32u) << 16) +
                         ((uint32_t)warp >> 2) * (uint32_t)D * TCOLS_WARP;

  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const long nn = (long)ND * ND, hw = (long)H * W;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const float inv_m = 1.0f / (float)M;
  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_stream = l2_policy_evict_first();

  float2* patches = a.scratch + (long)blockIdx.x * msf_scratch_elems(D, M, ND);
  float2* pnext = patches + (long)D * nn;           // [(t - 1) * M + m] for slices t >= 1
  float2* waves = pnext + (long)(D - 1) * M * nn;   // far fields of modes 0 .. M - 2
  const long nrep_stride = (long)a.nrep * M * nn;   // replicas of one slice
  float2* rep0 = replicas_all + (long)(blockIdx.x % a.nrep) * M * nn;

  auto slot = [&](int idx) { return (idx >> LG) * P + (idx & (ND - 1)); };
  // tile *= (conj) Hperm, elementwise in slot order
  // (global loads are issued in batches of eight before their first use: the
  // volatile stores / reductions next to them would otherwise serialise every
  // load behind its consumer)
  auto times_propagator = [&](bool conj) {
#pragma unroll 1
    for (int k0 = 0; k0 < KMAX; k0 += 8) {
      float2 h[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) h[q] = __ldg(hperm + tid + (k0 + q) * NT);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = tid + (k0 + q) * NT;
        if (conj) h[q].y = -h[q].y;
        tile[slot(idx)] = cmul(tile[slot(idx)], h[q]);
      }
    }
    __syncthreads();
  };

  long s_next = 0;
  for (long s = blockIdx.x; s < b.npos; s = s_next) {
    unsigned int tk = 0;
    if (tid == 0 && a.ticket) tk = atomicAdd(a.ticket, 1u);
    const Corner c = make_corner(b.scan, s);
    const long dbase = s * nn;

    // ---- patches of all slices (each thread keeps the pixels it owns) ----------
    {
      const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + ND < H) & (c.ix + ND < W);
      for (int t = 0; t < D; ++t) {
        const float2* img = psi + (long)t * hw;
#pragma unroll 8
        for (int k = 0; k < KMAX; ++k) {
          const int idx = tid + k * NT, py = idx >> LG, px = idx & (ND - 1);
          float2 o;
          if (interior) {
            const float2* r0 = img + (long)(c.iy + py) * W + c.ix + px;
            const float2 v00 = __ldg(r0), v01 = __ldg(r0 + 1);
            const float2 v10 = __ldg(r0 + W), v11 = __ldg(r0 + W + 1);
            o.x = v00.x * c.w00; o.y = v00.y * c.w00;
            o.x += v01.x * c.w01; o.y += v01.y * c.w01;
            o.x += v10.x * c.w10; o.y += v10.y * c.w10;
            o.x += v11.x * c.w11; o.y += v11.y * c.w11;
          } else {
            o = patch_value(img, H, W, c, py, px);
          }
          __stcg(patches + (long)t * nn + idx, o);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) F[tid + k * NT] = 0.f;
    {
      float z[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = 0.f;
      for (int t = 0; t < D; ++t)
#pragma unroll
        for (int j = 0; j < NCH; ++j) tmem_st16(tbase + t * TCOLS_WARP + j * 16, z);
      tmem_wait_st();
    }
    if (tid == 0) sh_next = a.ticket ? (long)tk + gridDim.x : s + gridDim.x;

    // ---- forward: far field of every mode, intensity -----------------------------
    for (int m = 0; m < M; ++m) {
      const float2* pm = probe + (long)m * nn;
#pragma unroll 1
      for (int k0 = 0; k0 < KMAX; k0 += 8) {
        float2 pv[8], ov[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          pv[q] = __ldg(pm + tid + (k0 + q) * NT);
          ov[q] = __ldcg(patches + tid + (k0 + q) * NT);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) tile[slot(tid + (k0 + q) * NT)] = cmul(pv[q], ov[q]);
      }
      __syncthreads();
      for (int t = 0; t + 1 < D; ++t) {
        fft2_tile<ND, false>(tile, tw);
        times_propagator(false);
        fft2_tile<ND, true>(tile, tw);  // the probe incident on slice t + 1
        float2* keep = pnext + ((long)t * M + m) * nn;
        const float2* onext = patches + (long)(t + 1) * nn;
#pragma unroll 1
        for (int k0 = 0; k0 < KMAX; k0 += 8) {
          float2 ov[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) ov[q] = __ldcg(onext + tid + (k0 + q) * NT);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (k0 + q) * NT;
            const float2 p = tile[slot(idx)];
            st_f32x2_keep(keep + idx, p, pol_keep);
            tile[slot(idx)] = cmul(p, ov[q]);
          }
        }
        __syncthreads();
      }
      fft2_tile<ND, false>(tile, tw);
      const bool last = (m == M - 1);  // stays in the tile for the backward sweep
      float2* wave = waves + (long)m * nn;
#pragma unroll 4
      for (int k = 0; k < KMAX; ++k) {
        const int idx = tid + k * NT;
        const float2 w = tile[slot(idx)];
        F[idx] += cabs2(w) * s2;
        if (!last) st_f32x2_keep(wave + idx, w, pol_keep);
      }
      if (!last) __syncthreads();
    }
    __syncthreads();

    // ---- cost and modulus factor (objective.py:11-66), data in natural order ----
    {
      float sums[1] = {0.f};
      constexpr int CB = 8;
#pragma unroll 1
      for (int k0 = 0; k0 < KMAX; k0 += CB) {
        float d[CB];
        bool meas[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          const int pix = tid + (k0 + j) * NT;
          meas[j] = a.mask ? (a.mask[pix] != 0) : true;
          d[j] = 0.f;
          if (meas[j]) d[j] = load_data_stream(a.data, a.data_u16, dbase + pix, pol_stream);
        }
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          const int pix = tid + (k0 + j) * NT;
          const int l = (int)f2l[pix >> LG] * ND + (int)f2l[pix & (ND - 1)];
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
      if (tid == 0) a.costs[s] = sums[0] * a.inv_nmeasured;
    }
    s_next = sh_next;
    __syncthreads();
    if (!a.accumulate_object) continue;

    // ---- backward: gradients of every slice ------------------------------------
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi == 0) ? M - 1 : mi - 1;  // the last mode is still in the tile
      if (mi == 0) {
#pragma unroll 4
        for (int k = 0; k < KMAX; ++k) {
          const int idx = tid + k * NT;
          tile[slot(idx)] = cscale(tile[slot(idx)], F[idx]);
        }
      } else {
        const float2* wave = waves + (long)m * nn;
#pragma unroll 1
        for (int k0 = 0; k0 < KMAX; k0 += 8) {
          float2 w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) w[q] = ld_f32x2_stream(wave + tid + (k0 + q) * NT, pol_stream);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (k0 + q) * NT;
            tile[slot(idx)] = cscale(w[q], F[idx]);
          }
        }
      }
      __syncthreads();
      fft2_tile<ND, true>(tile, tw);  // chi of the last slice
      for (int t = D - 1; t >= 0; --t) {
        const float2* pt = (t == 0) ? probe + (long)m * nn : pnext + ((long)(t - 1) * M + m) * nn;
        const float2* ot = patches + (long)t * nn;
        float2* rep = rep0 + (long)t * nrep_stride + (long)m * nn;
#pragma unroll 1
        for (int j = 0; j < NCH; ++j) {
          float2 pv[8], ov[8], xv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (j * 8 + q) * NT;
            pv[q] = __ldcg(pt + idx);
            ov[q] = __ldcg(ot + idx);
          }
          float v[16];
          tmem_ld16_issue(tbase + t * TCOLS_WARP + j * 16, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) xv[q] = tile[slot(tid + (j * 8 + q) * NT)];
          tmem_wait_ld(v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 g = cmulc(pv[q], xv[q]);
            v[2 * q] += g.x;
            v[2 * q + 1] += g.y;
          }
          tmem_st16(tbase + t * TCOLS_WARP + j * 16, v);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_f32x2(rep + tid + (j * 8 + q) * NT, cmulc(ov[q], xv[q]));
        }
        tmem_wait_st();
        if (t == 0) break;
        __syncthreads();
        fft2_tile<ND, false>(tile, tw);
        times_propagator(true);  // adjoint Fresnel step, rpie.py:474
        fft2_tile<ND, true>(tile, tw);
      }
      __syncthreads();
    }

    // ---- scatter-add of the object gradients, slice by slice -------------------
    for (int t = 0; t < D; ++t) {
      float2* G = tile;  // ND x ND, pitch ND
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float v[16];
        tmem_ld16(tbase + t * TCOLS_WARP + j * 16, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = tid + (j * 8 + q) * NT, py = idx >> LG, px = idx & (ND - 1);
          const int y = c.iy + py, x = c.ix + px;
          const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
          G[idx] = lead_ok ? make_float2(v[2 * q] * inv_m, v[2 * q + 1] * inv_m)
                           : make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      float2* out = psi_num_all + (long)t * hw;
      for (int ty = warp; ty <= ND; ty += NWARP) {
        const int y = c.iy + ty;
        if (y < 0 || y >= H) continue;
        const bool a0 = ty < ND, a1 = ty > 0;
        for (int tx = lane; tx <= ND; tx += 32) {
          const int x = c.ix + tx;
          if (x < 0 || x >= W) continue;
          float2 v = make_float2(0.f, 0.f);
          const bool b0 = tx < ND, b1 = tx > 0;
          if (a0 & b0) { const float2 g = G[ty * ND + tx];           v.x += c.w00 * g.x; v.y += c.w00 * g.y; }
          if (a0 & b1) { const float2 g = G[ty * ND + tx - 1];       v.x += c.w01 * g.x; v.y += c.w01 * g.y; }
          if (a1 & b0) { const float2 g = G[(ty - 1) * ND + tx];     v.x += c.w10 * g.x; v.y += c.w10 * g.y; }
          if (a1 & b1) { const float2 g = G[(ty - 1) * ND + tx - 1]; v.x += c.w11 * g.x; v.y += c.w11 * g.y; }
          red_add_f32x2(out + (long)y * W + x, v);
        }
      }
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_slot, tcols);
}

// ---- object preconditioner of the slices >= 1 (_preconditioner.py:76-94) ----
// psi_precond[t] += scatter_s( sum_m |probe_{t,s,m}|^2 ), where probe_t is the
// (unweighted, shared) probe carried through the slices before t.  Same forward
// chain as above, one position per CTA pass; the per-slice amplitude planes live
// in the per-CTA scratch (each thread re-reads only what it wrote).
template <int ND>
__global__ void __launch_bounds__(MsfCfg<ND>::NT, 1)
ms_precond_fused_kernel(tb_batch b, int D, const float2* __restrict__ hperm,
                        float2* __restrict__ scratch, float2* __restrict__ out,
                        unsigned int* ticket) {
  using Cfg = MsfCfg<ND>;
  constexpr int NT = Cfg::NT, KMAX = Cfg::KMAX, P = ND + 1, LG = Log2<ND>::v;
  constexpr int NWARP = Cfg::NWARP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = reinterpret_cast<float2*>(reinterpret_cast<float*>(tile + ND * P) + ND * ND);
  __shared__ long sh_next;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = b.nmodes, H = b.height, W = b.width;
  const long nn = (long)ND * ND, hw = (long)H * W;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  // per CTA: D - 1 patches (slices 0 .. D - 2), then D - 1 amplitude planes (floats)
  float2* patches = scratch + (long)blockIdx.x * ((long)(D - 1) * nn + ((long)(D - 1) * nn + 1) / 2);
  float* amp = reinterpret_cast<float*>(patches + (long)(D - 1) * nn);
  auto slot = [&](int idx) { return (idx >> LG) * P + (idx & (ND - 1)); };

  long s_next = 0;
  for (long s = blockIdx.x; s < b.npos; s = s_next) {
    if (tid == 0) sh_next = ticket ? (long)atomicAdd(ticket, 1u) + gridDim.x : s + gridDim.x;
    const Corner c = make_corner(b.scan, s);
    for (int t = 0; t + 1 < D; ++t) {
      const float2* img = psi + (long)t * hw;
#pragma unroll 8
      for (int k = 0; k < KMAX; ++k) {
        const int idx = tid + k * NT;
        __stcg(patches + (long)t * nn + idx, patch_value(img, H, W, c, idx >> LG, idx & (ND - 1)));
      }
    }
    for (int m = 0; m < M; ++m) {
      const float2* pm = probe + (long)m * nn;
#pragma unroll 1
      for (int k0 = 0; k0 < KMAX; k0 += 8) {
        float2 pv[8], ov[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          pv[q] = __ldg(pm + tid + (k0 + q) * NT);
          ov[q] = __ldcg(patches + tid + (k0 + q) * NT);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) tile[slot(tid + (k0 + q) * NT)] = cmul(pv[q], ov[q]);
      }
      __syncthreads();
      for (int t = 0; t + 1 < D; ++t) {
        fft2_tile<ND, false>(tile, tw);
#pragma unroll 1
        for (int k0 = 0; k0 < KMAX; k0 += 8) {
          float2 h[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) h[q] = __ldg(hperm + tid + (k0 + q) * NT);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (k0 + q) * NT;
            tile[slot(idx)] = cmul(tile[slot(idx)], h[q]);
          }
        }
        __syncthreads();
        fft2_tile<ND, true>(tile, tw);  // the probe incident on slice t + 1
        float* at = amp + (long)t * nn;
        const bool more = t + 2 < D;
        const float2* onext = patches + (long)(t + 1) * nn;
#pragma unroll 1
        for (int k0 = 0; k0 < KMAX; k0 += 8) {
          float acc[8];
          float2 ov[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (k0 + q) * NT;
            acc[q] = m == 0 ? 0.f : __ldcg(at + idx);
            if (more) ov[q] = __ldcg(onext + idx);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int idx = tid + (k0 + q) * NT;
            const float2 p = tile[slot(idx)];
            __stcg(at + idx, acc[q] + cabs2(p));
            if (more) tile[slot(idx)] = cmul(p, ov[q]);
          }
        }
        __syncthreads();
      }
    }
    s_next = sh_next;
    // scatter of every plane with the four bilinear weights
    float* G = reinterpret_cast<float*>(tile);  // ND x ND floats
    for (int t = 1; t < D; ++t) {
      const float* at = amp + (long)(t - 1) * nn;
#pragma unroll 8
      for (int k = 0; k < KMAX; ++k) {
        const int idx = tid + k * NT, py = idx >> LG, px = idx & (ND - 1);
        const int y = c.iy + py, x = c.ix + px;
        const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
        G[idx] = lead_ok ? __ldcg(at + idx) : 0.f;
      }
      __syncthreads();
      float2* o = out + (long)t * hw;
      for (int ty = warp; ty <= ND; ty += NWARP) {
        const int y = c.iy + ty;
        if (y < 0 || y >= H) continue;
        const bool a0 = ty < ND, a1 = ty > 0;
        for (int tx = lane; tx <= ND; tx += 32) {
          const int x = c.ix + tx;
          if (x < 0 || x >= W) continue;
          float v = 0.f;
          const bool b0 = tx < ND, b1 = tx > 0;
          if (a0 & b0) v += c.w00 * G[ty * ND + tx];
          if (a0 & b1) v += c.w01 * G[ty * ND + tx - 1];
          if (a1 & b0) v += c.w10 * G[(ty - 1) * ND + tx];
          if (a1 & b1) v += c.w11 * G[(ty - 1) * ND + tx - 1];
          red_add_f32(reinterpret_cast<float*>(o + (long)y * W + x), v);
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
static int msf_grid(long npos) {
  int sms = 148;
  tb_sm_count(&sms);
  return (int)(npos < sms ? npos : sms);
}

template <int ND>
static int msf_max_slices() { return MsfCfg<ND>::MAX_SLICES; }

bool multislice_fused_applies(const tb_rpie_args& a, int D) {
  const tb_batch& b = a.batch;
  if (D < 2 || b.probe_width != b.detector_width || b.probe_per_position ||
      b.eigen_weights != nullptr || a.eigen_weight_step != nullptr ||
      a.noise_model != TB_NOISE_GAUSSIAN)
    return false;
  switch (b.detector_width) {
    case 32:  return D <= msf_max_slices<32>();
    case 64:  return D <= msf_max_slices<64>();
    case 128: return D <= msf_max_slices<128>();
    default:  return false;
  }
}

int64_t multislice_fused_workspace_bytes(const tb_batch& b, int D) {
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  const int grid = msf_grid(b.npos);
  return ((int64_t)grid * msf_scratch_elems(D, b.nmodes, b.probe_width) +
          (int64_t)D * kMaxReplicas * n + (int64_t)b.probe_width * b.probe_width) * 8 + 64;
}

template <int ND>
static int launch_msf(RpieDev d, int D, const float2* prop, float2* hperm, float2* psi_num,
                      float2* replicas, int grid, cudaStream_t st) {
  msf_permute_propagator_kernel<ND><<<(ND * ND + 255) / 256, 256, 0, st>>>(prop, hperm);
  int rc = check_launch("tb_multislice_rpie_batch(fused: propagator)");
  if (rc != TB_OK) return rc;
  auto k = rpie_ms_fused_kernel<ND>;
  const size_t smem = MsfCfg<ND>::smem;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return set_error((int)e, "tb_multislice_rpie_batch(fused): %s", cudaGetErrorString(e));
  k<<<(unsigned)grid, MsfCfg<ND>::NT, smem, st>>>(d, D, hperm, psi_num, replicas);
  return check_launch("tb_multislice_rpie_batch(fused)");
}

bool multislice_precond_fused_applies(const tb_batch& b, int D) {
  if (D < 2 || b.probe_width != b.detector_width) return false;
  return b.detector_width == 32 || b.detector_width == 64 || b.detector_width == 128;
}

static int64_t msp_elems_per_cta(const tb_batch& b, int D) {
  const long nn = (long)b.probe_width * b.probe_width;
  return (long)(D - 1) * nn + ((long)(D - 1) * nn + 1) / 2;
}

int64_t multislice_precond_fused_workspace_bytes(const tb_batch& b, int D) {
  return ((int64_t)msf_grid(b.npos) * msp_elems_per_cta(b, D) +
          (int64_t)b.probe_width * b.probe_width) * 8 + 64;
}

template <int ND>
static int launch_msp(const tb_batch& b, int D, const float2* prop, float2* hperm,
                      float2* scratch, float2* out, unsigned int* ticket, int grid,
                      cudaStream_t st) {
  msf_permute_propagator_kernel<ND><<<(ND * ND + 255) / 256, 256, 0, st>>>(prop, hperm);
  int rc = check_launch("tb_multislice_precond_psi(fused: propagator)");
  if (rc != TB_OK) return rc;
  auto k = ms_precond_fused_kernel<ND>;
  const size_t smem = MsfCfg<ND>::smem;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return set_error((int)e, "tb_multislice_precond_psi(fused): %s", cudaGetErrorString(e));
  k<<<(unsigned)grid, MsfCfg<ND>::NT, smem, st>>>(b, D, hperm, scratch, out, ticket);
  return check_launch("tb_multislice_precond_psi(fused)");
}

// slices >= 1 of the object preconditioner (slice planes of `out` already zeroed)
int run_multislice_precond_fused(const tb_batch& plain, int D, const void* propagator,
                                 void* out, void* workspace, cudaStream_t st) {
  const int grid = msf_grid(plain.npos);
  float2* scratch = (float2*)workspace;
  float2* hperm = scratch + (long)grid * msp_elems_per_cta(plain, D);
  unsigned int* ticket = (unsigned int*)(hperm + (long)plain.probe_width * plain.probe_width);
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess)
    return set_error((int)e, "tb_multislice_precond_psi(fused): %s", cudaGetErrorString(e));
  switch (plain.detector_width) {
    case 32:  return launch_msp<32>(plain, D, (const float2*)propagator, hperm, scratch, (float2*)out, ticket, grid, st);
    case 64:  return launch_msp<64>(plain, D, (const float2*)propagator, hperm, scratch, (float2*)out, ticket, grid, st);
    default:  return launch_msp<128>(plain, D, (const float2*)propagator, hperm, scratch, (float2*)out, ticket, grid, st);
  }
}

// One fused launch for the whole batch; probe_numerator (D, M, N, N) overwritten.
int run_multislice_fused(const tb_rpie_args& a, int D, const void* propagator, cudaStream_t st) {
  const tb_batch& b = a.batch;
  const long n = (long)b.nmodes * b.probe_width * b.probe_width;
  const int grid = msf_grid(b.npos);
  RpieDev d{};
  d.b = b;
  d.b.neigen = 0;
  d.data = a.data;
  d.data_u16 = (a.data_dtype == TB_DATA_U16);
  d.mask = a.mask;
  d.noise_model = a.noise_model;
  d.unmeasured_factor = a.unmeasured_scaling - 1.0f;
  d.inv_nmeasured = 1.0f / (float)a.num_measured;
  d.costs = a.costs;
  d.accumulate_object = a.accumulate_object;
  d.divide_by_modes = 1;
  d.scratch = (float2*)a.workspace;
  d.nrep = grid < kMaxReplicas ? grid : kMaxReplicas;
  float2* replicas = d.scratch + (long)grid * msf_scratch_elems(D, b.nmodes, b.probe_width);
  float2* hperm = replicas + (long)D * d.nrep * n;
  d.ticket = (unsigned int*)(hperm + (long)b.probe_width * b.probe_width);
  cudaError_t e = cudaMemsetAsync(d.ticket, 0, sizeof(unsigned int), st);
  if (e == cudaSuccess && a.accumulate_object)
    e = cudaMemsetAsync(replicas, 0, (size_t)D * d.nrep * n * 8, st);
  if (e != cudaSuccess)
    return set_error((int)e, "tb_multislice_rpie_batch(fused): %s", cudaGetErrorString(e));
  int rc;
  switch (b.detector_width) {
    case 32:  rc = launch_msf<32>(d, D, (const float2*)propagator, hperm, (float2*)a.psi_numerator, replicas, grid, st); break;
    case 64:  rc = launch_msf<64>(d, D, (const float2*)propagator, hperm, (float2*)a.psi_numerator, replicas, grid, st); break;
    default:  rc = launch_msf<128>(d, D, (const float2*)propagator, hperm, (float2*)a.psi_numerator, replicas, grid, st); break;
  }
  if (rc != TB_OK || !a.accumulate_object) return rc;
  const long blocks = (n + 255) / 256;
  for (int t = 0; t < D; ++t) {
    reduce_replicas_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(
        replicas + (long)t * d.nrep * n, d.nrep, n, n, (float2*)a.probe_numerator + (long)t * n);
    rc = check_launch("tb_multislice_rpie_batch(fused: reduce)");
    if (rc != TB_OK) return rc;
  }
  return TB_OK;
}

}  // namespace tb
