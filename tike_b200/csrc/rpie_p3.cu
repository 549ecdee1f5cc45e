// Fused rPIE / DM / lstsq-phase-1 batch kernel, three-pass variant for the
// headline tile: detector width = probe width = 128, Gaussian model, shared or
// varying probe (BASELINE configs 2 and 4).
//
// rpie_fast.cu runs a 2-D transform as four shared-memory stages (column 8,
// row 8, row 16, column 16: the tile is read 3x and written 3x).  Here every
// thread keeps 32 complex values and each pass advances BOTH axes:
//
//   pass 1   column radix-8  (x) row radix-4    rows n2 + 16 k, columns m + 32 a
//   pass 2   column radix-16 (x) row radix-2    rows 16 k1 + n2, columns 32 a1 + 16 b + p
//   pass 3   row radix-16, two blocks           row r, columns 16 B + p
//
// (column plan 8 x 16 as in fft.cuh, row plan 4 x 2 x 16; decimation in
// frequency, outputs digit-reversed, inverse = the mirror image).  The tile is
// written twice and read twice per transform, and there are three block
// barriers per transform instead of four.  Everything that touches global
// memory sits in pass 1 (probe x patch, gradients: lanes walk over columns) or
// in pass 3, whose far-field values stay with the thread that produced them:
//   * the spilled far fields use a private thread-major layout (coalesced
//     whatever the digit order), the intensity plane is a private float4 layout;
//   * the last mode is parked in the accumulator columns of Tensor Memory (free
//     until the gradient sweep) between the forward and the inverse transform;
//   * the measured pattern is read in natural order (coalesced), parked in the
//     idle tile under an XOR swizzle and picked up conflict-free by the owners,
//     so cost and modulus factor need no pass of their own over the plane.
// Column twiddles are warp-uniform here (n2 = warp): they come from constant
// memory, not from shared memory.  Global operands are fetched into registers
// one pass ahead, across the block barrier (TB_P3_PREFETCH).  Measurements and
// the variants that did not pay: DESIGN.md section 4, item 0.
// Replaces: rpie.py:355-505, objective.py:11-66 (same scope as rpie_fast.cu).
#include "solver_dev.cuh"
#include "tmem.cuh"

#ifdef TB_PHASE_TIMING
__device__ unsigned long long tb_p3_phase_cycles[16];
#define P3_PHASE_DECL __shared__ unsigned int ph[13]; if (threadIdx.x == 0) { for (int i_ = 0; i_ < 12; ++i_) ph[i_] = 0; ph[12] = (unsigned int)clock64(); }
#define P3_PHASE(i) do { if (threadIdx.x == 0) { const unsigned int t_ = (unsigned int)clock64(); ph[i] += t_ - ph[12]; ph[12] = t_; } } while (0)
#define P3_PHASE_FLUSH do { if (threadIdx.x == 0) { for (int i_ = 0; i_ < 12; ++i_) atomicAdd(&tb_p3_phase_cycles[i_], (unsigned long long)ph[i_]); } } while (0)
extern "C" int tb_debug_phases_p3(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, tb_p3_phase_cycles, sizeof(tb_p3_phase_cycles));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tb_p3_phase_cycles, z, sizeof(z)); }
  return 0;
}
#else
#define P3_PHASE_DECL
#define P3_PHASE(i)
#define P3_PHASE_FLUSH
#endif

// sqrt.approx / div.approx instead of the IEEE sequences in the cost / modulus
// phase (1-2 ulp, far inside the 1e-4 parity tolerance; 20.34 -> 20.01 ms per
// 20k positions, profiles/r02q_*); -DTB_P3_APPROX_MODULUS=0 restores IEEE
#ifndef TB_P3_APPROX_MODULUS
#define TB_P3_APPROX_MODULUS 1
#endif

// how the reloaded far field leaves L2 (A/B): 0 = no discard (20.3 vs 20.0 ms per
// 20k positions: the dead waves are written back to HBM), 1 = one
// discard.global.L2 per 128-byte line issued by the lane that starts it (8
// instructions, four lanes each, per thread and block of 16 slots).  Measured
// and removed: one instruction with a line per lane (24.7 ms: a 32-line discard
// stalls the memory pipe), the same discards issued after the barrier under the
// shared-memory-only pass 2 (19.8 vs 19.2 ms).
// global operands fetched into registers one pass ahead, across the barrier.
// Measured and removed: the spilled far field of the next mode fetched at the
// end of inverse pass 1 (20.4 vs 20.0 ms).
#ifndef TB_P3_PREFETCH
#define TB_P3_PREFETCH 13  // bit 0: probe values (forward), 2: probe values (inverse), 3: measured pattern
#endif
#ifndef TB_P3_DISCARD
#define TB_P3_DISCARD 1
#endif

namespace tb {

namespace p3 {

constexpr int ND = 128, NT = 512, P = ND + 1, NW = NT / 32, KMAX = ND * ND / NT;
constexpr int WP = ND + 4;  // object window row pitch (16-byte multiples)
// tile | intensity / factor plane | per-lane twiddles | block-sum scratch
constexpr size_t kSmem = (size_t)ND * P * 8 + (size_t)ND * ND * 4 + 4 * 32 * 8 + 32 * 4;
static_assert((size_t)(ND + 1) * WP * 8 <= (size_t)ND * P * 8 + (size_t)ND * ND * 4,
              "object window exceeds tile + factor plane");

// w128^j = exp(-2 pi i j / 128); read with warp-uniform indices
__constant__ float2 c_tw[128] = {
#include "tw128.inc"
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "P3_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra P3_DONE_%=;\n"
      "bra P3_WAIT_%=;\n"
      "P3_DONE_%=:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy by the TMA unit (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes,
                                              unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"((unsigned)__cvta_generic_to_shared(dst)),
      "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}
// two slots per 16-byte access: half the global-memory instructions of the
// spill / reload (those passes wait on the load-store queue, ncu: lg_throttle)
__device__ __forceinline__ void st_wave2(float2* addr, float2 u, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr),
               "f"(u.x), "f"(u.y), "f"(v.x), "f"(v.y), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void ld_wave2(const float2* addr, float2& u, float2& v, uint64_t pol) {
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(u.x), "=f"(u.y), "=f"(v.x), "=f"(v.y)
               : "l"(addr), "l"(pol));
}
// read-only load that keeps its place in the instruction stream (the compiler
// sinks plain __ldg loads down to their first use)
__device__ __forceinline__ float2 ld_probe(const float2* addr) {
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(addr));
  return v;
}
__device__ __forceinline__ void discard_line(const void* addr) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(addr) : "memory");
}

}  // namespace p3

// VP = per-position varying probe (probe.py:272-303: mode m of position s is
// w[s,0,m] * P_m + sum_c w[s,c+1,m] * E_c,m) and the rPIE eigen-weight step
// (rpie.py:493-506) when a.eig_step is set: BASELINE config 4
template <bool VP>
__global__ void __launch_bounds__(p3::NT, 1) rpie_p3_kernel(RpieDev a) {
  using namespace p3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float4* F4 = reinterpret_cast<float4*>(tile + ND * P);  // [8][NT], slot j = 4 * (j / 4) + comp
  float2* twl = reinterpret_cast<float2*>(F4 + 8 * NT);    // [4][32] per-lane twiddles
  float* red = reinterpret_cast<float*>(twl + 4 * 32);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long win_bar;
  __shared__ long sh_next;

  const int tid = threadIdx.x, lane = tid & 31;
  // warp index as a value the compiler knows to be uniform (constant-bank
  // twiddles, uniform address arithmetic)
  const int wu = __shfl_sync(0xffffffffu, tid >> 5, 0);

  // per-lane row twiddles: w128^(lane * a1), a1 = 1..3, and w32^(lane & 15)
  if (tid < 128) {
    const int a1 = tid >> 5, l = tid & 31;
    twl[tid] = (a1 == 0) ? c_tw[(4 * (l & 15)) & 127] : c_tw[(l * a1) & 127];
  }
  unsigned win_phase = 0;
  if (tid == 0) {
    mbar_init(&win_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) tmem_alloc(&tmem_slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // Tensor Memory: 64 columns of accumulator and 64 of patch per thread; the
  // four warps of a lane quadrant take different column ranges
  const uint32_t tacc = tmem_slot + (((uint32_t)(wu & 3) * 32u) << 16) + (uint32_t)(wu >> 2) * 64u;
  const uint32_t tpat = tacc + 256u;
  const float2* const twl_lane = twl + lane;  // [0]: w32^(lane & 15), [32 a1]: w128^(lane a1)

  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const bool need_back = a.accumulate_object || a.probe_sums || a.chi_out;

  // per-CTA scratch, laid out like rpie_fast.cu's: patch (unused here), waves
  float2* waves = a.scratch + (long)blockIdx.x * ((long)ND * ND + (long)M * ND * ND) + ND * ND;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * ND * ND : nullptr;

  // pass-1 ownership: rows wu + 16 k, columns lane + 32 a
  const int o1 = wu * ND + lane;          // into (ND, ND) arrays, + 16 k * ND + 32 a
  float2* const t1 = tile + wu * P + lane;  // + 16 k * P + 32 a
  // pass-2 ownership: rows 16 k1 + n2, columns 32 a1 + 16 b + p
  float2* const t2 = tile + (16 * (tid >> 6)) * P + 32 * ((tid >> 4) & 3) + (tid & 15);
  // pass-3 ownership: row 32 (wu & 3) + lane, columns 32 (wu >> 2) + 16 q + p
  float2* const t3 = tile + (32 * (wu & 3) + lane) * P + 32 * (wu >> 2);
  // frequency of pass-3 slot (q, p1): row 8 (lane & 15) + 2 (wu & 3) + (lane >> 4),
  // column (wu >> 2) + 4 q + 8 p1; the measured pattern is parked as
  // D[fr * ND + (fc ^ sw(fr))], sw(fr) = ((fr >> 3) & 15) | ((fr & 1) << 4) = lane here
  const float* const Dmine = reinterpret_cast<const float*>(tile) +
                             (8 * (lane & 15) + 2 * (wu & 3) + (lane >> 4)) * ND;

  P3_PHASE_DECL
  long s_next = 0;
  for (long s = blockIdx.x; s < b.npos; s = s_next) {
    unsigned int tk = 0;
    if (tid == 0 && a.ticket) tk = atomicAdd(a.ticket, 1u);
    const long dbase = s * (long)ND * ND;
    P3_PHASE(11);
    // unique probe of this position in the pass-1 ownership (column block aa):
    // scale the shared mode, add the eigen probes
    [[maybe_unused]] const float* wpos =
        VP ? b.eigen_weights + s * (long)(b.neigen + 1) * M : nullptr;
    [[maybe_unused]] auto vary = [&](float2 (&x)[8], int m, int aa) {
      const float w0 = __ldg(wpos + m);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = cscale(x[k], w0);
      const float2* __restrict__ eigen = (const float2*)b.eigen_probe;
      if (eigen != nullptr && m < b.eigen_modes) {
        for (int e = 0; e < b.neigen; ++e) {
          const float we = __ldg(wpos + (e + 1) * M + m);
          const float2* em = eigen + ((long)e * b.eigen_modes + m) * ND * ND + o1 + 32 * aa;
          float2 ev[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) ev[k] = __ldg(em + 16 * k * ND);
#pragma unroll
          for (int k = 0; k < 8; ++k) { x[k].x += we * ev[k].x; x[k].y += we * ev[k].y; }
        }
      }
    };
    // Global operands are fetched into registers one pass ahead of their use,
    // across the block barrier: when a pass starts its data has landed and the
    // 16 warps do not all wait on L2 together.  pr: probe values of the next
    // forward pass 1 (this position's mode 0 flies under the patch phase).
    float2 pr[4][8];
#if TB_P3_PREFETCH & 1
#pragma unroll
    for (int aa = 0; aa < 2; ++aa)
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[aa][k] = __ldg(probe + o1 + 16 * k * ND + 32 * aa);
#endif

    // ------------- patch, pass-1 ownership, parked in Tensor Memory ----------
    {
      const Corner c = make_corner(b.scan, s);
      const int ixa = c.ix & ~1;
      const bool ok = (c.iy >= 0) & (c.iy + ND + 1 <= H) & (ixa >= 0) & (ixa + WP <= W) &
                      ((W & 1) == 0) & ((reinterpret_cast<uintptr_t>(psi) & 15) == 0);
      if (ok) {  // uniform over the CTA: window by TMA bulk copies, one per row
        float2* win = tile;
        if (tid == 0) mbar_expect_tx(&win_bar, (unsigned)((ND + 1) * WP * 8));
        if (tid <= ND)
          bulk_copy_g2s(win + tid * WP, psi + (long)(c.iy + tid) * W + ixa, WP * 8, &win_bar);
        mbar_wait(&win_bar, win_phase);
        win_phase ^= 1;
        const float2* q0 = win + wu * WP + lane + (c.ix - ixa);
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          float v[16];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2* q = q0 + 16 * k * WP + 32 * aa;
            const float2 a00 = q[0], a01 = q[1], a10 = q[WP], a11 = q[WP + 1];
            float2 r;
            r.x = a00.x * c.w00; r.y = a00.y * c.w00;
            r.x += a01.x * c.w01; r.y += a01.y * c.w01;
            r.x += a10.x * c.w10; r.y += a10.y * c.w10;
            r.x += a11.x * c.w11; r.y += a11.y * c.w11;
            v[2 * k] = r.x;
            v[2 * k + 1] = r.y;
          }
          tmem_st16(tpat + aa * 16, v);
        }
      } else {
#pragma unroll 1
        for (int aa = 0; aa < 4; ++aa) {
          float v[16];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 r = patch_value(psi, H, W, c, wu + 16 * k, lane + 32 * aa);
            v[2 * k] = r.x;
            v[2 * k + 1] = r.y;
          }
          tmem_st16(tpat + aa * 16, v);
        }
      }
      tmem_wait_st();
      __syncthreads();  // window consumed before the plane is cleared / the tile written
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) F4[j4 * NT + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    P3_PHASE(0);

    // the measured pattern of this position in natural order (coalesced);
    // -1 marks pixels that were not measured
    // (value j rides in pr[j / 16][(j / 2) % 8].x / .y: in the last mode's pass 3
    // the pattern takes the registers of the probe prefetch)
    auto dslot = [&](int j) -> float& {
      float2& c = pr[j >> 4][(j >> 1) & 7];
      return (j & 1) ? c.y : c.x;
    };
    auto load_pattern = [&]() {
      const uint64_t pol_stream = l2_policy_evict_first();
      auto go = [&](auto U16) {
        constexpr bool u16 = decltype(U16)::value;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          const int pix = tid + j * NT;
          const bool meas = a.mask ? (a.mask[pix] != 0) : true;
          float d = -1.0f;
          if (meas) d = load_data_stream(a.data, u16 ? 1 : 0, dbase + pix, pol_stream);
          dslot(j) = d;
        }
      };
      if (a.data_u16) go(std::true_type{});
      else go(std::false_type{});
    };
    // ------------- sweep 1: far field of every mode, intensity ---------------
    for (int m = 0; m < M; ++m) {
      const float2* __restrict__ pm = probe + (long)m * ND * ND + o1;
      // pass 1: exit wave, column radix-8, row radix-4
      {
        float2 v[4][8];
        const float2 tw_r1 = twl_lane[32], tw_r2 = twl_lane[64], tw_r3 = twl_lane[96];
#if TB_P3_PREFETCH & 1
        // column blocks 0 and 1 came in under the previous pass; 2 and 3 fly
        // under their butterflies
#pragma unroll
        for (int aa = 2; aa < 4; ++aa)
#pragma unroll
          for (int k = 0; k < 8; ++k) pr[aa][k] = __ldg(pm + 16 * k * ND + 32 * aa);
#endif
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
#if !(TB_P3_PREFETCH & 1)
#pragma unroll
          for (int k = 0; k < 8; ++k) pr[aa][k] = __ldg(pm + 16 * k * ND + 32 * aa);
#endif
          if constexpr (VP) vary(pr[aa], m, aa);
          float pt[16];
          tmem_ld16(tpat + aa * 16, pt);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v[aa][k] = cmul(pr[aa][k], make_float2(pt[2 * k], pt[2 * k + 1]));
          dft<8>(v[aa]);
#pragma unroll
          for (int k = 1; k < 8; ++k) v[aa][k] = cmul(v[aa][k], c_tw[wu * k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float2 t[4] = {v[0][k], v[1][k], v[2][k], v[3][k]};
          dft<4>(t);
          t1[16 * k * P] = t[0];
          t1[16 * k * P + 32] = cmul(t[1], tw_r1);
          t1[16 * k * P + 64] = cmul(t[2], tw_r2);
          t1[16 * k * P + 96] = cmul(t[3], tw_r3);
        }
      }
      __syncthreads();
      P3_PHASE(1);
      // pass 2: column radix-16, row radix-2
      {
        float2 u[2][16];
        const float2 tw_p = twl_lane[0];
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int n = 0; n < 16; ++n) u[bb][n] = t2[n * P + 16 * bb];
        dft<16>(u[0]);
        dft<16>(u[1]);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          t2[n * P] = cadd(u[0][n], u[1][n]);
          t2[n * P + 16] = cmul(csub(u[0][n], u[1][n]), tw_p);
        }
      }
      __syncthreads();
      P3_PHASE(2);
      // pass 3: row radix-16; intensity; the far field goes to the per-CTA
      // scratch, the last mode's to the (still unused) accumulator columns of
      // Tensor Memory
      {
        const bool last = (m == M - 1);
        float2* wave = waves + (long)m * ND * ND + 2 * tid;  // slot pair (p, p + 1) at [(p / 2) * 2 NT + 2 tid]
#if TB_P3_PREFETCH & 1
        if (!last) {
#pragma unroll
          for (int aa = 0; aa < 2; ++aa)
#pragma unroll
            for (int k = 0; k < 8; ++k) pr[aa][k] = __ldg(pm + ND * ND + 16 * k * ND + 32 * aa);
        } else {
#if TB_P3_PREFETCH & 8
          load_pattern();  // last mode: the measured pattern of the cost phase
#else  // (a definition on every path keeps the old values from staying live)
#pragma unroll
          for (int aa = 0; aa < 2; ++aa)
#pragma unroll
            for (int k = 0; k < 8; ++k) pr[aa][k] = make_float2(0.f, 0.f);
#endif
        }
#endif
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float2 z[16];
#pragma unroll
          for (int p = 0; p < 16; ++p) z[p] = t3[16 * q + p];
          dft<16>(z);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 f = F4[(4 * q + j4) * NT + tid];
            f.x += cabs2(z[4 * j4 + 0]) * s2;
            f.y += cabs2(z[4 * j4 + 1]) * s2;
            f.z += cabs2(z[4 * j4 + 2]) * s2;
            f.w += cabs2(z[4 * j4 + 3]) * s2;
            F4[(4 * q + j4) * NT + tid] = f;
          }
          if (need_back) {
            if (last) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float v[16];
#pragma unroll
                for (int p = 0; p < 8; ++p) { v[2 * p] = z[8 * h + p].x; v[2 * p + 1] = z[8 * h + p].y; }
                tmem_st16(tacc + (2 * q + h) * 16, v);
              }
            } else {
              const uint64_t pol_keep = l2_policy_evict_last();
#pragma unroll
              for (int p = 0; p < 16; p += 2)
                st_wave2(wave + (16 * q + p) * NT, z[p], z[p + 1], pol_keep);
            }
          }
        }
        if (need_back && last) tmem_wait_st();
      }
      __syncthreads();
      P3_PHASE(3);
    }

    if (tid == 0) sh_next = a.ticket ? (long)tk + gridDim.x : s + gridDim.x;
    // ------------- cost and modulus factor (objective.py:11-66) --------------
    // the pattern in natural order (coalesced) -> swizzled floats in the idle tile
    {
      float* D = reinterpret_cast<float*>(tile);
#if !(TB_P3_PREFETCH & 8)
      load_pattern();
#endif
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int fr = (tid >> 7) + 4 * j, fc = tid & (ND - 1);
        const int sw = ((fr >> 3) & 15) | ((fr & 1) << 4);
        D[fr * ND + (fc ^ sw)] = dslot(j);
      }
    }
    __syncthreads();
    float cost[1] = {0.f};
    {
      const int fc0 = wu >> 2;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float4 f = F4[(4 * q + j4) * NT + tid];
          float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int p1 = 4 * j4 + e;
            const float d = Dmine[(fc0 + 4 * q + 8 * p1) ^ lane];
            float fac = a.unmeasured_factor * rt;
            if (d >= 0.f) {
#if TB_P3_APPROX_MODULUS
              float sd, sI;
              asm("sqrt.approx.f32 %0, %1;" : "=f"(sd) : "f"(d));
              asm("sqrt.approx.f32 %0, %1;" : "=f"(sI) : "f"(fv[e]));
              const float t = sI - sd;
              cost[0] += t * t;
              fac = -(1.0f - __fdividef(sd, sI + 1e-9f)) * rt;
#else
              const float sd = sqrtf(d), sI = sqrtf(fv[e]);
              const float t = sI - sd;
              cost[0] += t * t;
              fac = -(1.0f - sd / (sI + 1e-9f)) * rt;
#endif
            }
            fv[e] = fac;
          }
          F4[(4 * q + j4) * NT + tid] = make_float4(fv[0], fv[1], fv[2], fv[3]);
        }
      }
    }
    block_sum<1>(cost, red);  // two block barriers: every owner has read the pattern
    if (tid == 0) a.costs[s] = cost[0] * a.inv_nmeasured;
    // While this position runs its gradient sweep, pull what the NEXT position
    // of this CTA will read first (its pattern and its object tile) into L2.
    s_next = sh_next;
    {
      const long sn = s_next;
      if ((a.prefetch_next & 1) && sn < b.npos) {
        const char* dn = (const char*)a.data + sn * (long)ND * ND * (a.data_u16 ? 2 : 4);
        const int dbytes = ND * ND * (a.data_u16 ? 2 : 4);
        for (int off = tid * 128; off < dbytes; off += NT * 128) prefetch_l2(dn + off);
        const Corner cn = make_corner(b.scan, sn);
        if (cn.iy >= 0 && cn.ix >= 0 && cn.iy + ND < H && cn.ix + ND < W) {
          constexpr int LINES = ((ND + 1) * 8 + 127) / 128 + 1;
          for (int t = tid; t < (ND + 1) * LINES; t += NT) {
            const int row = t / LINES, ln = t - row * LINES;
            prefetch_l2((const char*)(psi + (long)(cn.iy + row) * W + cn.ix) + ln * 128);
          }
        }
      }
    }
    P3_PHASE(4);
    if (!need_back) {
      fence_proxy_async();
      __syncthreads();
      continue;
    }

    // ------------- sweep 2: gradients ----------------------------------------
    // finish of inverse pass 3 for one block of 16 slots: x modulus factor, row
    // radix-16, rows of the tile
    auto p3inv_block = [&](float2 (&z)[16], int q) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 f = F4[(4 * q + j4) * NT + tid];
        z[4 * j4 + 0] = cscale(z[4 * j4 + 0], f.x);
        z[4 * j4 + 1] = cscale(z[4 * j4 + 1], f.y);
        z[4 * j4 + 2] = cscale(z[4 * j4 + 2], f.z);
        z[4 * j4 + 3] = cscale(z[4 * j4 + 3], f.w);
      }
      idft<16>(z);
#pragma unroll
      for (int p = 0; p < 16; ++p) t3[16 * q + p] = z[p];
    };
    // the spilled wave is dead once reloaded: drop its lines from L2 without
    // write-back (the warp read 16 slots x 256 bytes per block)
    auto discard_block = [&](const float2* wave, int q) {
#if TB_P3_DISCARD == 1
      if ((lane & 7) == 0) {  // 8 lanes x 16 bytes = one line
#pragma unroll
        for (int p = 0; p < 16; p += 2) discard_line(wave + (16 * q + p) * NT);
      }
#endif
    };
    // inverse pass 3 of the last mode, parked in Tensor Memory
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float2 z[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[16];
        tmem_ld16(tacc + (2 * q + h) * 16, v);
#pragma unroll
        for (int p = 0; p < 8; ++p) z[8 * h + p] = make_float2(v[2 * p], v[2 * p + 1]);
      }
      p3inv_block(z, q);
    }
    if (a.accumulate_object) {  // the accumulator columns are free now
      float zz[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) zz[j] = 0.f;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) tmem_st16(tacc + aa * 16, zz);
      tmem_wait_st();
    }
    __syncthreads();
    P3_PHASE(5);
    [[maybe_unused]] float eig[2] = {0.f, 0.f};
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi == 0) ? M - 1 : mi - 1;  // last mode first
      // pull the next mode's spilled wave towards L2 while passes 2 and 1 run
      if (mi + 1 < M && (a.prefetch_next & 2) && wu == NW - 1) {
        const char* nxt = (const char*)(waves + (long)mi * ND * ND);  // next m = mi
        for (int ln = lane; ln < ND * ND * 8 / 128; ln += 32) prefetch_l2(nxt + ln * 128);
      }
      // probe values of inverse pass 1, two column blocks ahead of their use:
      // blocks 0 and 1 are fetched here, under inverse pass 2, blocks 2 and 3
      // when block 0 / 1 has been consumed
      const float2* __restrict__ pmi = probe + (long)m * ND * ND + o1;
      float2 pvA[8], pvB[8];
      const bool need_pv = a.accumulate_object || (VP && m == 0 && a.eig_step);
#if TB_P3_PREFETCH & 4
      if (need_pv) {
#pragma unroll
        for (int k = 0; k < 8; ++k) pvA[k] = ld_probe(pmi + 16 * k * ND);
#pragma unroll
        for (int k = 0; k < 8; ++k) pvB[k] = ld_probe(pmi + 16 * k * ND + 32);
      } else {  // (a definition on every path keeps old values from staying live)
#pragma unroll
        for (int k = 0; k < 8; ++k) pvA[k] = pvB[k] = make_float2(0.f, 0.f);
      }
#endif
      // pass 2 inverse: row radix-2, column radix-16
      {
        float2 u[2][16];
        const float2 tw_p = twl_lane[0];
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const float2 x0 = t2[n * P], x1 = cmulc(tw_p, t2[n * P + 16]);
          u[0][n] = cadd(x0, x1);
          u[1][n] = csub(x0, x1);
        }
        idft<16>(u[0]);
        idft<16>(u[1]);
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int n = 0; n < 16; ++n) t2[n * P + 16 * bb] = u[bb][n];
      }
      __syncthreads();
      P3_PHASE(6);
      // pass 1 inverse: row radix-4, column radix-8, gradients from registers
      {
        const float2* __restrict__ pm = probe + (long)m * ND * ND + o1;
        float2* rep = replica ? replica + (long)m * ND * ND + o1 : nullptr;
        float2 v[4][8];
        const float2 tw_r1 = twl_lane[32], tw_r2 = twl_lane[64], tw_r3 = twl_lane[96];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float2 t[4];
          t[0] = t1[16 * k * P];
          t[1] = cmulc(tw_r1, t1[16 * k * P + 32]);
          t[2] = cmulc(tw_r2, t1[16 * k * P + 64]);
          t[3] = cmulc(tw_r3, t1[16 * k * P + 96]);
          idft<4>(t);
          v[0][k] = t[0]; v[1][k] = t[1]; v[2][k] = t[2]; v[3][k] = t[3];
        }
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          float2 (&pv)[8] = (aa & 1) ? pvB : pvA;
#if !(TB_P3_PREFETCH & 4)
          if (need_pv) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pv[k] = __ldg(pm + 16 * k * ND + 32 * aa);
          }
#endif
#pragma unroll
          for (int k = 1; k < 8; ++k) v[aa][k] = cmulc(c_tw[wu * k], v[aa][k]);
          idft<8>(v[aa]);
          if (a.chi_out) {  // lstsq_grad keeps chi for its second phase
            float2* cout = a.chi_out + ((long)s * M + m) * ND * ND + o1 + 32 * aa;
#pragma unroll
            for (int k = 0; k < 8; ++k) cout[16 * k * ND] = v[aa][k];
          }
          if constexpr (VP) {
            if (m == 0 && a.eig_step) {
              // rpie.py:493-506: projection on the SHARED main mode times the patch
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float ov[8];
                tmem_ld8(tpat + aa * 16 + h * 8, ov);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 op = cmul(make_float2(ov[2 * k], ov[2 * k + 1]), pv[4 * h + k]);
                  eig[0] += op.x * v[aa][4 * h + k].x + op.y * v[aa][4 * h + k].y;
                  eig[1] += cabs2(op);
                }
              }
            }
            if (a.accumulate_object) vary(pv, m, aa);
          }
          // Tensor Memory in groups of four values: this pass is at the
          // register limit (64 for v, the probe values of this and the next
          // column block)
          if (a.accumulate_object) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float acc[8];
              tmem_ld8(tacc + aa * 16 + h * 8, acc);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 g = cmulc(pv[4 * h + k], v[aa][4 * h + k]);
                acc[2 * k] += g.x;
                acc[2 * k + 1] += g.y;
              }
              tmem_st8(tacc + aa * 16 + h * 8, acc);
            }
          }
#if TB_P3_PREFETCH & 4
          if (aa < 2 && need_pv) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pv[k] = ld_probe(pm + 16 * k * ND + 32 * (aa + 2));
          }
#endif
          if (rep) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float ov[8];
              tmem_ld8(tpat + aa * 16 + h * 8, ov);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                red_add_f32x2(rep + 16 * (4 * h + k) * ND + 32 * aa,
                              cmulc(make_float2(ov[2 * k], ov[2 * k + 1]), v[aa][4 * h + k]));
            }
          }
        }
        tmem_wait_st();
      }
      __syncthreads();
      P3_PHASE(7);
      // pass 3 inverse of the next mode: reload x modulus factor, row radix-16
      if (mi + 1 < M) {
        const float2* wave = waves + (long)mi * ND * ND + 2 * tid;
        const uint64_t pol_stream = l2_policy_evict_first();
        float2 zr0[16], zr1[16];
#pragma unroll
        for (int p = 0; p < 16; p += 2) ld_wave2(wave + p * NT, zr0[p], zr0[p + 1], pol_stream);
#pragma unroll
        for (int p = 0; p < 16; p += 2)
          ld_wave2(wave + (16 + p) * NT, zr1[p], zr1[p + 1], pol_stream);
        p3inv_block(zr0, 0);
        discard_block(wave, 0);
        p3inv_block(zr1, 1);
        discard_block(wave, 1);
        __syncthreads();
        P3_PHASE(5);
      }
    }

    if constexpr (VP) {
      if (a.eig_step) {
        block_sum<2>(eig, red);
        if (tid == 0) a.eig_step[s] = 0.1f * (eig[0] / eig[1]);
      }
    }
    // ------------- scatter-add of the object gradient ------------------------
    if (a.accumulate_object) {
      const Corner c = make_corner(b.scan, s);
      float2* G = tile;  // ND x ND, pitch ND
      const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        float v[16];
        tmem_ld16(tacc + aa * 16, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int py = wu + 16 * k, px = lane + 32 * aa;
          const int y = c.iy + py, x = c.ix + px;
          const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
          G[py * ND + px] = lead_ok ? make_float2(v[2 * k] * inv_m, v[2 * k + 1] * inv_m)
                                    : make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      // footprint pixels (ND + 1)^2, spread evenly over the threads
      for (int e = tid; e < (ND + 1) * (ND + 1); e += NT) {
        const int ty = e / (ND + 1), tx = e - ty * (ND + 1);
        const int y = c.iy + ty, x = c.ix + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        const bool a0 = ty < ND, a1 = ty > 0, b0 = tx < ND, b1 = tx > 0;
        float2 v = make_float2(0.f, 0.f);
        if (a0 & b0) { const float2 g = G[ty * ND + tx];           v.x += c.w00 * g.x; v.y += c.w00 * g.y; }
        if (a0 & b1) { const float2 g = G[ty * ND + tx - 1];       v.x += c.w01 * g.x; v.y += c.w01 * g.y; }
        if (a1 & b0) { const float2 g = G[(ty - 1) * ND + tx];     v.x += c.w10 * g.x; v.y += c.w10 * g.y; }
        if (a1 & b1) { const float2 g = G[(ty - 1) * ND + tx - 1]; v.x += c.w11 * g.x; v.y += c.w11 * g.y; }
        red_add_f32x2(a.psi_num + (long)y * W + x, v);
      }
    }
    // the next position's window arrives through the async proxy: order this
    // position's generic accesses to the tile before it
    fence_proxy_async();
    __syncthreads();
    P3_PHASE(8);
  }
  P3_PHASE_FLUSH;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_slot, 512);
}

// the three-pass kernel takes the 128 x 128 Gaussian batch with a shared or a
// varying probe (rPIE, DM, phase 1 of lstsq_grad); position-gradient sums,
// Poisson, padding and every other width stay with rpie_fast_kernel
bool p3_kernel_applies(const RpieDev& a) {
  const tb_batch& b = a.b;
  const bool varying = b.eigen_weights != nullptr;
  if (a.eig_step != nullptr && !varying) return false;
  return b.detector_width == 128 && b.probe_width == 128 && !b.probe_per_position &&
         a.pos_num == nullptr && a.noise_model == TB_NOISE_GAUSSIAN;
}

template <bool VP>
static int launch_p3_variant(const RpieDev& a, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rpie_p3_kernel<VP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p3::kSmem);
  if (e != cudaSuccess) return set_error((int)e, "rpie p3 kernel attr: %s", cudaGetErrorString(e));
  rpie_p3_kernel<VP><<<(unsigned)grid, p3::NT, p3::kSmem, st>>>(a);
  return check_launch("tb_rpie_batch(p3)");
}

int launch_p3(const RpieDev& a, int grid, cudaStream_t st) {
  return a.b.eigen_weights != nullptr ? launch_p3_variant<true>(a, grid, st)
                                      : launch_p3_variant<false>(a, grid, st);
}

}  // namespace tb
