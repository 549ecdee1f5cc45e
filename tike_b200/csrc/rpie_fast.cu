// Fused rPIE / lstsq batch kernel, stage-fused variant for probe width ==
// detector width in {32, 64, 128}.  The plain instantiation is the headline
// configuration (shared probe, Gaussian model); compile-time variants add the
// varying probe / eigen-weight step (VP), the lstsq position-gradient sums
// (PG) and the Poisson model with its step lengths (PO).
//
// Same per-position pipeline as rpie.cu, but every pass that touches global
// memory is merged into the first or last *column* stage of a 2-D transform,
// where the lanes of a warp walk over columns (coalesced 8-byte accesses):
//
//   forward :  colA [probe x patch built in registers] -> rowA -> rowB
//              -> colB [ |Psi|^2 accumulated, Psi spilled ]
//   inverse :  colB^-1 [Psi reloaded x modulus factor] -> rowB^-1 -> rowA^-1
//              -> colA^-1 [chi consumed from registers: conj(P) chi, conj(O) chi]
//
// Row and column stages act on different indices, so interleaving them this
// way is still the separable 2-D DFT.  Per transform the tile is read 3x and
// written 3x (5x / 5x in rpie.cu) and two block barriers disappear.
// Replaces: rpie.py:355-505, objective.py:11-66, lstsq.py:422-543 (phase 1).
#include "solver_dev.cuh"
#include "tmem.cuh"

#ifdef TB_PHASE_TIMING
// Development aid (python -m tike_b200.build with TB_NVCC_EXTRA=-DTB_PHASE_TIMING):
// thread 0 of every CTA accumulates the SM cycles spent between block barriers.
__device__ unsigned long long tb_phase_cycles[16];
#define TB_PHASE_DECL __shared__ unsigned int ph[13]; if (threadIdx.x == 0) { for (int i_ = 0; i_ < 12; ++i_) ph[i_] = 0; ph[12] = (unsigned int)clock64(); }
#define TB_PHASE(i) do { if (threadIdx.x == 0) { const unsigned int t_ = (unsigned int)clock64(); ph[i] += t_ - ph[12]; ph[12] = t_; } } while (0)
#define TB_PHASE_FLUSH do { if (threadIdx.x == 0) { for (int i_ = 0; i_ < 12; ++i_) atomicAdd(&tb_phase_cycles[i_], (unsigned long long)ph[i_]); } } while (0)
extern "C" int tb_debug_phases(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, tb_phase_cycles, sizeof(tb_phase_cycles));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tb_phase_cycles, z, sizeof(z)); }
  return 0;
}
#else
#define TB_PHASE_DECL
#define TB_PHASE(i)
#define TB_PHASE_FLUSH
#endif

// How the reloaded far field leaves L2: RELOAD_STREAM marks the lines
// evict_first when they are read back, DISCARD drops them without write-back to
// HBM (the wave is dead after the reload).  Both on: 23.3 -> 22.7 ms per 20k
// positions (profiles/r01o_variants.log); -DTB_EXP_...=0 through
// scripts/build_variant.py builds the library without them for A/B timing.
#ifndef TB_EXP_RELOAD_STREAM
#define TB_EXP_RELOAD_STREAM 1
#endif
#ifndef TB_EXP_DISCARD
#define TB_EXP_DISCARD 1
#endif
// Not measured yet (default off; build with scripts/build_variant.py and time
// with scripts/batch_order_experiment.py): issue the first probe loads of a
// column stage one stage early, so that their L2 latency sits under the
// previous stage instead of at the start of the next one.
//   HOIST_PROBE: colA of mode m + 1 gets its first butterfly's probe values
//                at the start of colB of mode m
//   HOIST_PV:    colA^-1 gets its first butterfly's probe values before the
//                last inverse row stage
//   APPROX_MODULUS: sqrt.approx / div.approx (1-2 ulp) instead of the IEEE
//                sequences in the Gaussian cost / modulus phase (6.7 % of the
//                kernel, issue bound on those sequences); far inside the 1e-4
//                parity tolerance but not bit-identical to the IEEE build
#ifndef TB_EXP_APPROX_MODULUS
#define TB_EXP_APPROX_MODULUS 0
#endif
//   GROUP_PIPE:  (ND = 128) inverse sweep with group-pipelined stages: the four
//                warps that wrote the 16 rows of a colB^-1 block run both
//                inverse row stages of those rows behind a 128-thread named
//                barrier, so the four groups drift apart and one group's L2
//                wait sits under another group's row butterflies
#ifndef TB_EXP_GROUP_PIPE
#define TB_EXP_GROUP_PIPE 0
#endif
#ifndef TB_EXP_HOIST_PROBE
#define TB_EXP_HOIST_PROBE 0
#endif
#ifndef TB_EXP_HOIST_PV
#define TB_EXP_HOIST_PV 0
#endif

//   TMA_WINDOW:  (plain variant) the object window under the footprint, (ND + 1)
//                rows, is brought into the (idle) tile by the TMA unit -- one
//                cp.async.bulk per row, completion on an mbarrier -- and the
//                bilinear patch is interpolated from shared memory instead of
//                four global loads per pixel
#ifndef TB_EXP_TMA_WINDOW
#define TB_EXP_TMA_WINDOW 1
#endif

namespace tb {

// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) and their mbarrier -------
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
      "TB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TB_DONE_%=;\n"
      "bra TB_WAIT_%=;\n"
      "TB_DONE_%=:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// generic-proxy accesses to shared memory before, async-proxy (TMA) accesses after
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes,
                                              unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"((unsigned)__cvta_generic_to_shared(dst)),
      "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void discard_l2_line(const void* addr) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(addr) : "memory");
}

// threads per CTA at ND = 128 (A/B: -DTB_EXP_NT128=1024 runs 32 warps on the tile
// at 64 registers per thread)
#ifndef TB_EXP_NT128
#define TB_EXP_NT128 512
#endif
template <int ND> struct FastCfg {
  static constexpr int NT = (ND >= 128) ? TB_EXP_NT128 : (ND >= 64 ? 256 : 128);
  static constexpr int R0 = plan_radix(ND, 0), R1 = plan_radix(ND, 1);
  static constexpr int NBA = ND * R1 / NT;  // radix-R0 column butterflies per thread
  static constexpr int NBB = ND * R0 / NT;  // radix-R1 column butterflies per thread
  static constexpr int KMAX = ND * ND / NT;
  static constexpr size_t smem = (size_t)ND * (ND + 1) * 8 + ND * ND * 4 + ND * 8 +
                                 ND * 4 + 4 * 32 * 4;
};

//   NO_CLOBBER:  the far-field spill stores and the probe-numerator reductions
//                carry no "memory" clobber (they stay volatile, i.e. ordered among
//                themselves and against the barriers): the compiler may then batch
//                the shared-memory accesses of the factor plane and the next
//                butterfly's loads around them.  Their targets are only read after a
//                block barrier (the spill) or by a later kernel (the reductions).
#ifndef TB_EXP_NO_CLOBBER
#define TB_EXP_NO_CLOBBER 0
#endif
__device__ __forceinline__ void st_f32x2_hint(float2* addr, float2 v, uint64_t pol) {
#if TB_EXP_NO_CLOBBER
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(v.x),
               "f"(v.y), "l"(pol));
#else
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(v.x),
               "f"(v.y), "l"(pol)
               : "memory");
#endif
}
__device__ __forceinline__ void red_add_f32x2_fast(float2* addr, float2 v) {
#if TB_EXP_NO_CLOBBER
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y));
#else
  red_add_f32x2(addr, v);
#endif
}
__device__ __forceinline__ float2 ld_f32x2_hint(const float2* addr, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y)
               : "l"(addr), "l"(pol));
  return v;
}

// fft_stage of fft.cuh for a subset of vectors handled by NTH threads (thread
// index t in [0, NTH)); used by the group-pipelined inverse sweep.
template <int N, int R, int L, bool INV, int LOGNVEC, int VSTRIDE, int ESTRIDE, int NTH>
__device__ __forceinline__ void fft_stage_sub(float2* __restrict__ s,
                                              const float2* __restrict__ tw, int t) {
  constexpr int S = L / R, BF = N / R, TWS = N / L;
  constexpr int total = BF << LOGNVEC;
  constexpr int vmask = (1 << LOGNVEC) - 1;
#pragma unroll
  for (int b0 = 0; b0 < total; b0 += NTH) {
    const int b = b0 + t;
    if (total % NTH != 0 && b >= total) break;
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
// barrier among the 128 threads of group g (named barriers 1..4; 0 is __syncthreads)
__device__ __forceinline__ void group_barrier(int g) {
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

// VP = the variant with the per-position extras: varying probe
// (probe.py:272-303: mode m of position s is w[s,0,m] * P_m + sum_c w[s,c+1,m] *
// E_c,m), the rPIE eigen-weight step (rpie.py:493-506) when a.eig_step is set,
// PG = the VP variant that also emits the lstsq position-gradient sums
// (lstsq.py:545-579) into a.pos_num / a.pos_den.
// PO = Poisson noise model (objective.py:72-124) with the fixed-point step
// lengths of exitwave.py:122-234 (per mode, or dominant mode only).
// PAD = probe narrower than the detector: the N x N exit wave sits zero-padded
// in the middle of the ND x ND tile (convolution.py:58-101, pad = (ND - N) / 2).
template <int ND, bool TM, bool VP, bool PG, bool PO, bool PAD>
__global__ void __launch_bounds__(FastCfg<ND>::NT, (ND >= 128) ? 1 : 2)
rpie_fast_kernel(RpieDev a) {
  using Cfg = FastCfg<ND>;
  constexpr int NT = Cfg::NT, R0 = Cfg::R0, R1 = Cfg::R1, NBA = Cfg::NBA, NBB = Cfg::NBB;
  constexpr int KMAX = Cfg::KMAX, P = ND + 1, LG = Log2<ND>::v, NWARP = NT / 32;
  static_assert(R0 * R1 == ND && NBA >= 1 && NBB >= 1, "two-stage plans only");
  static_assert(!VP || TM, "the varying-probe variant is written for the TMEM build");
  static_assert(!PG || VP, "position gradients live in the VP variant");
  static_assert(!PO || (TM && !PG), "the Poisson variant reuses the patch scratch for the data");
  static_assert(!PAD || (TM && !VP && !PG && !PO), "padding is offered for the plain variant");
  constexpr int NA2 = (NBA % 2 == 0 && R0 <= 8) ? 2 : 1;  // colA butterflies loaded together
  constexpr int GB = R0 < 8 ? R0 : 8;                     // gradient load batch
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float* F = reinterpret_cast<float*>(tile + ND * P);
  float2* tw = reinterpret_cast<float2*>(F + ND * ND);
  unsigned short* l2f = reinterpret_cast<unsigned short*>(tw + ND);
  float* red = reinterpret_cast<float*>(l2f + 2 * ND);
  fill_twiddles<ND>(tw);
  unsigned short* f2l = l2f + ND;
  for (int i = threadIdx.x; i < ND; i += NT) {
    l2f[i] = (unsigned short)loc2freq<ND>(i);
    f2l[i] = (unsigned short)freq2loc<ND>(i);
  }
  // TMEM columns: 2*KMAX floats per thread, one column range per warp of a quadrant
  static_assert(!TM || R0 == 8, "TMEM accumulator path moves 16 floats per butterfly");
  constexpr uint32_t TCOLS_WARP = 2 * KMAX;
  constexpr uint32_t TCOLS_RAW = TCOLS_WARP * ((NWARP + 3) / 4);
  // second half of the allocation: the interpolated patch of the current position
  constexpr uint32_t TCOLS_BOTH = 2 * TCOLS_RAW;
  static_assert(!TM || TCOLS_BOTH <= 512, "accumulator + patch exceed Tensor Memory");
  constexpr uint32_t TCOLS = TCOLS_BOTH <= 32 ? 32 : (TCOLS_BOTH <= 64 ? 64 : (TCOLS_BOTH <= 128 ? 128 : (TCOLS_BOTH <= 256 ? 256 : 512)));
  __shared__ uint32_t tmem_slot;
  // object-window loads by the TMA unit: plain variant, window rows of WP complex
  // values (16-byte multiples) laid over the tile (the last rows reach into F,
  // which is only initialised after the patch has been taken)
  constexpr bool TMAWIN = TB_EXP_TMA_WINDOW && TM && !VP && !PG && !PO && !PAD;
  constexpr int WP = ND + 4;
  static_assert(!TMAWIN || (size_t)(ND + 1) * WP * 8 <= (size_t)ND * P * 8 + (size_t)ND * ND * 4,
                "object window exceeds tile + factor plane");
  __shared__ __align__(8) unsigned long long win_bar;
  [[maybe_unused]] unsigned win_phase = 0;
  if constexpr (TMAWIN) {
    if (threadIdx.x == 0) {
      mbar_init(&win_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  uint32_t tacc = 0;
  [[maybe_unused]] uint32_t tpat = 0;
  if constexpr (TM) {
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, TCOLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if constexpr (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t wq = (threadIdx.x >> 5) & 3, wc = threadIdx.x >> 7;
    tacc = tmem_slot + ((wq * 32u) << 16) + wc * TCOLS_WARP;
    tpat = tacc + TCOLS_RAW;
  }

  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool need_back = a.accumulate_object || a.probe_sums || a.chi_out;
  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_stream = l2_policy_evict_first();

  // probe / patch width and its offset inside the tile
  const int N = PAD ? b.probe_width : ND;
  const int pad = PAD ? (ND - N) / 2 : 0;
  // tile pixel (row, col) -> inside the probe support? / index into (N, N) arrays
  auto inside = [&](int row, int col) {
    return !PAD || ((row >= pad) & (row < pad + N) & (col >= pad) & (col < pad + N));
  };
  auto pidx = [&](int row, int col) { return PAD ? (row - pad) * N + (col - pad) : row * ND + col; };

  // per-CTA scratch: patch (N*N) then waves (M*ND*ND)
  float2* patch = a.scratch + (long)blockIdx.x * ((long)N * N + (long)M * ND * ND);
  float2* waves = patch + (long)N * N;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * N * N : nullptr;

  // column-stage coordinates of this thread (fixed for the whole kernel)
  int colA[NBA], n2A[NBA], colB[NBB], k1B[NBB];
#pragma unroll
  for (int i = 0; i < NBA; ++i) { const int q = tid + i * NT; colA[i] = q & (ND - 1); n2A[i] = q >> LG; }
#pragma unroll
  for (int i = 0; i < NBB; ++i) { const int q = tid + i * NT; colB[i] = q & (ND - 1); k1B[i] = q >> LG; }

  TB_PHASE_DECL
  // Positions are handed out dynamically: the first one is blockIdx.x, every
  // further one comes from a global counter.  A CTA that reaches its SM late
  // takes fewer positions instead of leaving a tail (matters when a
  // communication kernel of the multi-GPU exchange borrows SMs mid-launch).
  __shared__ long sh_next;
  long s_next = 0;
  for (long s = blockIdx.x; s < b.npos; s = s_next) {
    [[maybe_unused]] unsigned int tk = 0;
    if (tid == 0 && a.ticket) tk = atomicAdd(a.ticket, 1u);
    const Corner c = make_corner(b.scan, s);
    const long dbase = s * (long)ND * ND;
    TB_PHASE(11);
    [[maybe_unused]] const float* wpos =
        (VP && b.eigen_weights) ? b.eigen_weights + s * (long)(b.neigen + 1) * M : nullptr;
    [[maybe_unused]] const float2* __restrict__ eigen = (const float2*)b.eigen_probe;
    // unique probe of this position: scale the shared mode, add the eigen probes
    [[maybe_unused]] auto vary = [&](float2 (&x)[R0], int m, int i) {
      const float w0 = __ldg(wpos + m);
#pragma unroll
      for (int k = 0; k < R0; ++k) x[k] = cscale(x[k], w0);
      if (eigen != nullptr && m < b.eigen_modes) {
        for (int e = 0; e < b.neigen; ++e) {
          const float we = __ldg(wpos + (e + 1) * M + m);
          const float2* em = eigen + ((long)e * b.eigen_modes + m) * ND * ND;
          float2 ev[R0];
#pragma unroll
          for (int k = 0; k < R0; ++k) ev[k] = __ldg(em + (n2A[i] + R1 * k) * ND + colA[i]);
#pragma unroll
          for (int k = 0; k < R0; ++k) { x[k].x += we * ev[k].x; x[k].y += we * ev[k].y; }
        }
      }
    };

    // ------------- patch in the colA ownership: rows n2 + R1*k, column c ----
    // (TMEM build: the patch is parked in Tensor Memory, one x16 row per butterfly)
    float2 o[TM ? 1 : NBA][R0];
    [[maybe_unused]] bool win_done = false;
    if constexpr (TMAWIN) {
      // rows c.iy .. c.iy + ND, columns ixa .. ixa + WP - 1 with ixa = c.ix rounded
      // down to an even column (16-byte aligned source): needs an even row
      // pitch, an aligned base and the window inside the image
      const int ixa = c.ix & ~1;
      const bool ok = (c.iy >= 0) & (c.iy + ND + 1 <= H) & (ixa >= 0) & (ixa + WP <= W) &
                      ((W & 1) == 0) & ((reinterpret_cast<uintptr_t>(psi) & 15) == 0);
      if (ok) {  // uniform over the CTA
        float2* win = tile;
        if (tid == 0) mbar_expect_tx(&win_bar, (unsigned)((ND + 1) * WP * 8));
        if (tid <= ND)
          bulk_copy_g2s(win + tid * WP, psi + (long)(c.iy + tid) * W + ixa, WP * 8, &win_bar);
        mbar_wait(&win_bar, win_phase);
        win_phase ^= 1;
        const int dx = c.ix - ixa;
#pragma unroll
        for (int i = 0; i < NBA; ++i) {
          float v[16];
#pragma unroll
          for (int k = 0; k < R0; ++k) {
            const float2* q = win + (n2A[i] + R1 * k) * WP + colA[i] + dx;
            const float2 a00 = q[0], a01 = q[1], a10 = q[WP], a11 = q[WP + 1];
            float2 r;
            r.x = a00.x * c.w00; r.y = a00.y * c.w00;
            r.x += a01.x * c.w01; r.y += a01.y * c.w01;
            r.x += a10.x * c.w10; r.y += a10.y * c.w10;
            r.x += a11.x * c.w11; r.y += a11.y * c.w11;
            v[2 * k] = r.x;
            v[2 * k + 1] = r.y;
          }
          tmem_st16(tpat + i * 16, v);
        }
        tmem_wait_st();
        __syncthreads();  // window consumed before colA writes the tile
        win_done = true;
      }
    }
    if (!win_done) {
      const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < H) & (c.ix + N < W);
#pragma unroll
      for (int i = 0; i < NBA; ++i) {
#pragma unroll
        for (int k0 = 0; k0 < R0; k0 += 4) {
          float2 v[4][4];
          if (interior) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int row = n2A[i] + R1 * (k0 + j);
              if (inside(row, colA[i])) {
                const float2* r0 = psi + (long)(c.iy + row - pad) * W + c.ix + colA[i] - pad;
                v[j][0] = __ldg(r0); v[j][1] = __ldg(r0 + 1);
                v[j][2] = __ldg(r0 + W); v[j][3] = __ldg(r0 + W + 1);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row = n2A[i] + R1 * (k0 + j);
            float2 r;
            if (!inside(row, colA[i])) {
              r = make_float2(0.f, 0.f);  // zero padding around the exit wave
            } else if (interior) {
              r.x = v[j][0].x * c.w00; r.y = v[j][0].y * c.w00;
              r.x += v[j][1].x * c.w01; r.y += v[j][1].y * c.w01;
              r.x += v[j][2].x * c.w10; r.y += v[j][2].y * c.w10;
              r.x += v[j][3].x * c.w11; r.y += v[j][3].y * c.w11;
            } else {
              r = patch_value(psi, H, W, c, row - pad, colA[i] - pad);
            }
            o[TM ? 0 : i][k0 + j] = r;
            if ((!TM && need_back) || PG) __stcg(patch + row * ND + colA[i], r);
          }
        }
        if constexpr (TM) {
          float v[16];
#pragma unroll
          for (int k = 0; k < R0; ++k) { v[2 * k] = o[0][k].x; v[2 * k + 1] = o[0][k].y; }
          tmem_st16(tpat + i * 16, v);
        }
      }
      if constexpr (TM) tmem_wait_st();
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) F[tid + k * NT] = 0.f;
    TB_PHASE(0);

    // probe value of mode mm at tile pixel (row, col); zero outside the support
    [[maybe_unused]] auto probe_of = [&](int mm, int row, int col) {
      const float2* q = probe + (long)mm * N * N;
      if constexpr (PAD) {
        return inside(row, col) ? __ldg(q + pidx(row, col)) : make_float2(0.f, 0.f);
      } else {
        return __ldg(q + row * ND + col);
      }
    };
    // the hoists only pay where registers are left: the plain variant
    constexpr bool PLAIN = !VP && !PG && !PO && !PAD;
    constexpr bool HOIST_PROBE = TB_EXP_HOIST_PROBE && PLAIN;
    constexpr bool HOIST_PV = TB_EXP_HOIST_PV && PLAIN && TM;
    // group-pipelined inverse sweep: one colB^-1 butterfly per thread and round,
    // groups of 128 threads owning the R1 = 16 rows of a block
    constexpr bool GROUP_PIPE = TB_EXP_GROUP_PIPE && ND == 128 && NT == 512 && R1 == 16;
    // probe values of the first colA butterfly of the next mode
    [[maybe_unused]] float2 pre[HOIST_PROBE ? R0 : 1];
    if constexpr (HOIST_PROBE) {
#pragma unroll
      for (int k = 0; k < R0; ++k) pre[k] = probe_of(0, n2A[0] + R1 * k, colA[0]);
    }
    // ------------- sweep 1: far field of every mode, intensity -------------
    for (int m = 0; m < M; ++m) {
      const float2* __restrict__ pm = probe + (long)m * N * N;
      // probe value at tile pixel (row, col); zero outside the support
      auto probe_at = [&](int row, int col) {
        if constexpr (PAD) {
          return inside(row, col) ? __ldg(pm + pidx(row, col)) : make_float2(0.f, 0.f);
        } else {
          return __ldg(pm + row * ND + col);
        }
      };
      // colA fused with the exit-wave build; the probe loads of butterfly
      // i + 1 are in flight while butterfly i is computed
      {
        float2 nxt[R0];
        if constexpr (HOIST_PROBE) {
#pragma unroll
          for (int k = 0; k < R0; ++k) nxt[k] = pre[k];
        } else {
#pragma unroll
          for (int k = 0; k < R0; ++k) nxt[k] = probe_at(n2A[0] + R1 * k, colA[0]);
        }
#pragma unroll
        for (int i = 0; i < NBA; ++i) {
          float2 x[R0];
#pragma unroll
          for (int k = 0; k < R0; ++k) x[k] = nxt[k];
          if (i + 1 < NBA) {
#pragma unroll
            for (int k = 0; k < R0; ++k) nxt[k] = probe_at(n2A[i + 1] + R1 * k, colA[i + 1]);
          }
          if constexpr (VP) {
            if (wpos) vary(x, m, i);
          }
          if constexpr (TM) {
            float v[16];
            tmem_ld16(tpat + i * 16, v);
#pragma unroll
            for (int k = 0; k < R0; ++k) x[k] = cmul(x[k], make_float2(v[2 * k], v[2 * k + 1]));
          } else {
#pragma unroll
            for (int k = 0; k < R0; ++k) x[k] = cmul(x[k], o[i][k]);
          }
          dft<R0>(x);
#pragma unroll
          for (int k = 1; k < R0; ++k) x[k] = cmul(x[k], tw[n2A[i] * k]);
#pragma unroll
          for (int k = 0; k < R0; ++k) tile[(n2A[i] + R1 * k) * P + colA[i]] = x[k];
        }
      }
      __syncthreads();
      TB_PHASE(1);
      fft_stage<ND, R0, ND, false, LG, P, 1>(tile, tw);  // rows, stage A
      __syncthreads();
      TB_PHASE(2);
      fft_stage<ND, R1, R1, false, LG, P, 1>(tile, tw);  // rows, stage B
      __syncthreads();
      TB_PHASE(3);
      // colB fused with the intensity accumulation and the spill
      float2* wave = waves + (long)m * ND * ND;
      if constexpr (HOIST_PROBE) {
        if (m + 1 < M) {
#pragma unroll
          for (int k = 0; k < R0; ++k) pre[k] = probe_of(m + 1, n2A[0] + R1 * k, colA[0]);
        }
      }
#pragma unroll
      for (int i = 0; i < NBB; ++i) {
        float2 x[R1];
#pragma unroll
        for (int n = 0; n < R1; ++n) x[n] = tile[(k1B[i] * R1 + n) * P + colB[i]];
        dft<R1>(x);
        // the last mode stays in the tile: sweep 2 starts with it, so its
        // far field never goes to global memory (all of it when M == 1)
        const bool keep_in_tile = (m == M - 1);
#pragma unroll
        for (int n = 0; n < R1; ++n) {
          const int l = (k1B[i] * R1 + n) * ND + colB[i];
          F[l] += cabs2(x[n]) * s2;
          if (need_back) {
            if (keep_in_tile) tile[(k1B[i] * R1 + n) * P + colB[i]] = x[n];
            else st_f32x2_hint(wave + l, x[n], pol_keep);
          }
        }
      }
      __syncthreads();
      TB_PHASE(4);
    }

    // next position of this CTA (visible to all threads after the block_sum below)
    if (tid == 0) sh_next = a.ticket ? (long)tk + gridDim.x : s + gridDim.x;
    // ------------- cost and modulus factor (objective.py:11-66) -------------
    // Threads walk the detector in natural pixel order (coalesced, batched
    // loads of the measured pattern); the matching tile location comes from
    // the digit-reversal table.
    // Poisson: the measured pattern re-ordered to tile slots (negative = not
    // measured) lives in the per-CTA patch scratch, F keeps the intensity
    [[maybe_unused]] float* dperm = reinterpret_cast<float*>(patch);
    [[maybe_unused]] const bool per_mode_steps = PO && a.step_mode == TB_STEP_ALL_MODES;
    if constexpr (PO) {
      float sums[3] = {0.f, 0.f, 0.f};  // cost, dominant-mode denominator, numerator
      float step_dom = a.step_start;
      const bool dominant = a.step_mode == TB_STEP_DOMINANT_MODE;
      constexpr int CB = KMAX >= 8 ? 8 : KMAX;
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
            const float I = F[l];
            sums[0] += I - d[j] * logf(I + 1e-9f);
            if (dominant) {
              const float xi = a.poisson_eps ? 1.0f - d[j] / (I + 1e-9f) : 1.0f - d[j] / I;
              sums[1] += xi * xi * I;
              sums[2] += xi * (I - d[j] / (1.0f - step_dom * xi));
            }
            __stcg(dperm + l, d[j]);
          } else {
            __stcg(dperm + l, -1.0f);
          }
        }
      }
      block_sum<3>(sums, red);
      if (tid == 0) a.costs[s] = sums[0] * a.inv_nmeasured;
      if (dominant) {
        // exitwave.py:183-234: second fixed-point iteration, then the factor
        // plane is the same for every mode
        step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (sums[2] / sums[1]);
        float s1[1] = {0.f};
#pragma unroll 4
        for (int k = 0; k < KMAX; ++k) {
          const int l = tid + k * NT;
          const float d = __ldcg(dperm + l);
          if (d >= 0.f) {
            const float I = F[l];
            const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
            s1[0] += xi * (I - d / (1.0f - step_dom * xi));
          }
        }
        block_sum<1>(s1, red);
        step_dom = (1.0f - a.step_weight) * step_dom + a.step_weight * (s1[0] / sums[1]);
#pragma unroll 4
        for (int k = 0; k < KMAX; ++k) {
          const int l = tid + k * NT;
          const float d = __ldcg(dperm + l);
          float f = a.unmeasured_factor;
          if (d >= 0.f) {
            const float I = F[l];
            const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
            f = -step_dom * xi;
          }
          F[l] = f * rt;
        }
      }
    } else {
      float sums[1] = {0.f};
      constexpr int CB = KMAX >= 8 ? 8 : KMAX;
      // the data type is resolved outside the batched loads (a per-load branch
      // keeps the compiler from issuing the CB loads back to back)
      auto cost_pass = [&](auto U16) {
        constexpr bool u16 = decltype(U16)::value;
#pragma unroll 1
        for (int k0 = 0; k0 < KMAX; k0 += CB) {
          float d[CB];
          bool meas[CB];
#pragma unroll
          for (int j = 0; j < CB; ++j) {
            const int pix = tid + (k0 + j) * NT;
            meas[j] = a.mask ? (a.mask[pix] != 0) : true;
            d[j] = 0.f;
            if (meas[j]) d[j] = load_data_stream(a.data, u16 ? 1 : 0, dbase + pix, pol_stream);
          }
#pragma unroll
          for (int j = 0; j < CB; ++j) {
            const int pix = tid + (k0 + j) * NT;
            const int l = (int)f2l[pix >> LG] * ND + (int)f2l[pix & (ND - 1)];
            if (meas[j]) {
#if TB_EXP_APPROX_MODULUS
              const float sd = sqrt_approx(d[j]), sI = sqrt_approx(F[l]);
              const float t = sI - sd;
              sums[0] += t * t;
              F[l] = -(1.0f - __fdividef(sd, sI + 1e-9f)) * rt;
#else
              const float sd = sqrtf(d[j]), sI = sqrtf(F[l]);
              const float t = sI - sd;
              sums[0] += t * t;
              F[l] = -(1.0f - sd / (sI + 1e-9f)) * rt;
#endif
            } else {
              F[l] = a.unmeasured_factor * rt;
            }
          }
        }
      };
      if (a.data_u16) cost_pass(std::true_type{});
      else cost_pass(std::false_type{});
      block_sum<1>(sums, red);
      if (tid == 0) a.costs[s] = sums[0] * a.inv_nmeasured;
    }
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
        if (cn.iy >= 0 && cn.ix >= 0 && cn.iy + N < H && cn.ix + N < W) {
          // (N + 1) rows of (N + 1) complex values; one 128-byte line per lane
          constexpr int LINES = ((ND + 1) * 8 + 127) / 128 + 1;
          for (int t = tid; t < (N + 1) * LINES; t += NT) {
            const int row = t / LINES, ln = t - row * LINES;
            prefetch_l2((const char*)(psi + (long)(cn.iy + row) * W + cn.ix) + ln * 128);
          }
        }
      }
    }
    if (!need_back) {
      if constexpr (TMAWIN) fence_proxy_async();
      __syncthreads();
      continue;
    }
    __syncthreads();  // factors visible to the colB^-1 ownership
    TB_PHASE(5);

    // ------------- sweep 2: gradients ---------------------------------------
    [[maybe_unused]] float2 acc[TM ? 1 : NBA][TM ? 1 : R0];
    if constexpr (TM) {
      if (a.accumulate_object) {
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0.f;
#pragma unroll
        for (int i = 0; i < NBA; ++i) tmem_st16(tacc + i * 16, z);
        tmem_wait_st();
      }
    } else {
#pragma unroll
      for (int i = 0; i < NBA; ++i)
#pragma unroll
        for (int k = 0; k < R0; ++k) acc[i][k] = make_float2(0.f, 0.f);
    }

    [[maybe_unused]] float eig[2] = {0.f, 0.f};
    [[maybe_unused]] float pg[4] = {0.f, 0.f, 0.f, 0.f};  // position gradient sums
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi == 0) ? M - 1 : mi - 1;  // last mode first: it is still in the tile
      const bool from_tile = (mi == 0);
      const float2* wave = waves + (long)m * ND * ND;
      // Poisson, one step length per mode (exitwave.py:122-180): two
      // fixed-point iterations over |Psi_m|^2 of the whole pattern first
      [[maybe_unused]] float step_m = a.step_start;
      if constexpr (PO) {
        if (per_mode_steps) {
          float ab[NBB * R1];
#pragma unroll
          for (int i = 0; i < NBB; ++i)
#pragma unroll
            for (int n = 0; n < R1; ++n) {
              const int l = (k1B[i] * R1 + n) * ND + colB[i];
              const float2 xv = from_tile ? tile[(k1B[i] * R1 + n) * P + colB[i]]
                                          : ld_f32x2_hint(wave + l, pol_keep);
              ab[i * R1 + n] = cabs2(xv) * s2;
            }
          float q0 = 0.f;
#pragma unroll 1
          for (int it = 0; it < 2; ++it) {
            float q[2] = {0.f, 0.f};
#pragma unroll
            for (int i = 0; i < NBB; ++i)
#pragma unroll
              for (int n = 0; n < R1; ++n) {
                const int l = (k1B[i] * R1 + n) * ND + colB[i];
                const float d = __ldcg(dperm + l);
                if (d >= 0.f) {
                  const float I = F[l], av = ab[i * R1 + n];
                  const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
                  const float t = xi * step_m - 1.0f;
                  const float den = av * t * t + I - av;
                  q[0] += xi * xi * av;
                  q[1] += xi * av * (1.0f + (d * t) / den);
                }
              }
            block_sum<2>(q, red);
            if (it == 0) q0 = q[0];
            step_m = step_m * (1.0f - a.step_weight) + (q[1] / q0) * a.step_weight;
          }
        }
      }
      // colB^-1 fused with the reload and the modulus factor
#pragma unroll
      for (int i0 = 0; i0 < NBB; i0 += ((NBB >= 2 && R1 <= 8) ? 2 : 1)) {
        constexpr int NB2 = (NBB >= 2 && R1 <= 8) ? 2 : 1;
        float2 x[NB2][R1];
        if (from_tile) {
#pragma unroll
          for (int j = 0; j < NB2; ++j)
#pragma unroll
            for (int n = 0; n < R1; ++n)
              x[j][n] = tile[(k1B[i0 + j] * R1 + n) * P + colB[i0 + j]];
        } else {
#pragma unroll
          for (int j = 0; j < NB2; ++j)
#pragma unroll
            for (int n = 0; n < R1; ++n)
              x[j][n] = ld_f32x2_hint(wave + (k1B[i0 + j] * R1 + n) * ND + colB[i0 + j],
                                      TB_EXP_RELOAD_STREAM ? pol_stream : pol_keep);
        }
#pragma unroll
        for (int j = 0; j < NB2; ++j) {
#pragma unroll
          for (int n = 0; n < R1; ++n) {
            const int l = (k1B[i0 + j] * R1 + n) * ND + colB[i0 + j];
            float f = F[l];
            if constexpr (PO) {
              if (per_mode_steps) {  // F still holds the intensity
                const float d = __ldcg(dperm + l);
                const float I = f;
                f = a.unmeasured_factor * rt;
                if (d >= 0.f) {
                  const float xi = a.poisson_eps ? 1.0f - d / (I + 1e-9f) : 1.0f - d / I;
                  f = -step_m * xi * rt;
                }
              }
            }
            x[j][n] = cscale(x[j][n], f);
          }
          idft<R1>(x[j]);
#pragma unroll
          for (int n = 0; n < R1; ++n)
            tile[(k1B[i0 + j] * R1 + n) * P + colB[i0 + j]] = x[j][n];
#if TB_EXP_DISCARD
          // the spilled wave is dead once it has been reloaded: tell L2 so that
          // the dirty lines are dropped instead of written back to HBM (each
          // half warp read one 128-byte line per row)
          if (!from_tile && (lane & 15) == 0) {
#pragma unroll
            for (int n = 0; n < R1; ++n)
              discard_l2_line(wave + (k1B[i0 + j] * R1 + n) * ND + colB[i0 + j]);
          }
#endif
        }
        if constexpr (GROUP_PIPE) {
          // threads tid >> 7 == g wrote all ND columns of the R1 rows of block
          // k1B[i0] = g + 4 * i0; they transform those rows themselves
          const int g = tid >> 7, t = tid & 127;
          float2* rows = tile + k1B[i0] * R1 * P;
          group_barrier(g);
          fft_stage_sub<ND, R1, R1, true, 4, P, 1, 128>(rows, tw, t);  // stage B inverse
          group_barrier(g);
          fft_stage_sub<ND, R0, ND, true, 4, P, 1, 128>(rows, tw, t);  // stage A inverse
        }
      }
      __syncthreads();
      TB_PHASE(6);
      // pull the next mode's spilled wave towards L2 while the row stages run
      if (mi + 1 < M && (a.prefetch_next & 2)) {
        const char* nxt = (const char*)(waves + (long)mi * ND * ND);  // next m = (mi + 1) - 1
        // issued by one warp only: the stage ends at a barrier, and 32 extra
        // instructions in one warp cost less than 2 in each of the 16
        if (warp == NWARP - 1)
          for (int ln = lane; ln < ND * ND * 8 / 128; ln += 32) prefetch_l2(nxt + ln * 128);
      }
      if constexpr (!GROUP_PIPE) {
        fft_stage<ND, R1, R1, true, LG, P, 1>(tile, tw);  // rows, stage B inverse
        __syncthreads();
      }
      TB_PHASE(7);
      [[maybe_unused]] float2 pv0[HOIST_PV ? R0 : 1];
      if constexpr (HOIST_PV) {
        if (a.accumulate_object) {
#pragma unroll
          for (int k = 0; k < R0; ++k) pv0[k] = probe_of(m, n2A[0] + R1 * k, colA[0]);
        }
      }
      if constexpr (!GROUP_PIPE) {
        fft_stage<ND, R0, ND, true, LG, P, 1>(tile, tw);  // rows, stage A inverse
        __syncthreads();
      }
      TB_PHASE(8);
      // colA^-1 fused with the gradient accumulation
      const float2* __restrict__ pm = probe + (long)m * N * N;
      float2* rep = replica ? replica + (long)m * N * N : nullptr;
      float2* cout = a.chi_out ? a.chi_out + ((long)s * M + m) * N * N : nullptr;
#pragma unroll
      for (int i = 0; i < NBA; ++i) {
        // probe values first: their L2 latency hides behind the butterfly
        [[maybe_unused]] float2 pv[TM ? R0 : 1];
        if constexpr (TM) {
          if (a.accumulate_object || (VP && m == 0 && (a.eig_step || PG))) {
            if (HOIST_PV && i == 0) {
#pragma unroll
              for (int k = 0; k < R0; ++k) pv[k] = pv0[HOIST_PV ? k : 0];
            } else {
#pragma unroll
              for (int k = 0; k < R0; ++k) {
                const int row = n2A[i] + R1 * k;
                if constexpr (PAD) {
                  pv[k] = inside(row, colA[i]) ? __ldg(pm + pidx(row, colA[i]))
                                               : make_float2(0.f, 0.f);
                } else {
                  pv[k] = __ldg(pm + row * ND + colA[i]);
                }
              }
            }
          }
        }
        float2 x[R0];
#pragma unroll
        for (int k = 0; k < R0; ++k) x[k] = tile[(n2A[i] + R1 * k) * P + colA[i]];
#pragma unroll
        for (int k = 1; k < R0; ++k) x[k] = cmulc(tw[n2A[i] * k], x[k]);
        idft<R0>(x);
        if (cout) {
#pragma unroll
          for (int k = 0; k < R0; ++k) {
            const int row = n2A[i] + R1 * k;
            if (inside(row, colA[i])) cout[pidx(row, colA[i])] = x[k];
          }
        }
        if constexpr (TM) {
          // accumulator and patch both live in TMEM
          [[maybe_unused]] float ov[16];
          if (rep || (VP && m == 0 && a.eig_step)) tmem_ld16(tpat + i * 16, ov);
          if constexpr (VP) {
            if (m == 0 && a.eig_step) {
              // rpie.py:493-506: projection on the SHARED main mode times the patch
#pragma unroll
              for (int k = 0; k < R0; ++k) {
                const float2 op = cmul(make_float2(ov[2 * k], ov[2 * k + 1]), pv[k]);
                eig[0] += op.x * x[k].x + op.y * x[k].y;
                eig[1] += cabs2(op);
              }
            }
            if (wpos && (a.accumulate_object || (m == 0 && PG))) vary(pv, m, i);
            if (PG && m == 0) {
              // lstsq.py:545-579 on the centre crop [N/4, N - N/4): Gaussian
              // derivative of the patch (neighbours from the L2 copy) times the
              // unique main probe, projected on chi
              constexpr int crop = ND / 4;
              const int px = colA[i];
              if (px >= crop && px < ND - crop) {
#pragma unroll
                for (int k = 0; k < R0; ++k) {
                  const int py = n2A[i] + R1 * k;
                  if (py >= crop && py < ND - crop) {
                    float2 gy = make_float2(0.f, 0.f), gx = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t = -2; t <= 2; ++t) {
                      const float wt = a.taps[t + 2];
                      const float2 oy = __ldcg(patch + (py + t) * ND + px);
                      const float2 ox = __ldcg(patch + py * ND + px + t);
                      gy.x -= wt * oy.x; gy.y -= wt * oy.y;
                      gx.x -= wt * ox.x; gx.y -= wt * ox.y;
                    }
                    const float2 ay = cmul(gy, pv[k]), ax = cmul(gx, pv[k]);
                    pg[0] += ay.x * x[k].x + ay.y * x[k].y;
                    pg[1] += cabs2(ay);
                    pg[2] += ax.x * x[k].x + ax.y * x[k].y;
                    pg[3] += cabs2(ax);
                  }
                }
              }
            }
          }
          if (a.accumulate_object) {
            float v[16];
            tmem_ld16(tacc + i * 16, v);
#pragma unroll
            for (int k = 0; k < R0; ++k) {
              const float2 g = cmulc(pv[k], x[k]);
              v[2 * k] += g.x;
              v[2 * k + 1] += g.y;
            }
            tmem_st16(tacc + i * 16, v);
          }
          if (rep) {
#pragma unroll
            for (int k = 0; k < R0; ++k) {
              const int row = n2A[i] + R1 * k;
              if (inside(row, colA[i]))
                red_add_f32x2_fast(rep + pidx(row, colA[i]),
                              cmulc(make_float2(ov[2 * k], ov[2 * k + 1]), x[k]));
            }
          }
        } else {
  if (a.accumulate_object && rep) {
#pragma unroll
            for (int k0 = 0; k0 < R0; k0 += GB) {
              float2 p[GB], q[GB];  // probe and patch loads issued together
#pragma unroll
              for (int j = 0; j < GB; ++j) {
                const int l = (n2A[i] + R1 * (k0 + j)) * ND + colA[i];
                p[j] = __ldg(pm + l);
                q[j] = __ldcg(patch + l);
              }
#pragma unroll
              for (int j = 0; j < GB; ++j) {
                const int l = (n2A[i] + R1 * (k0 + j)) * ND + colA[i];
                const float2 g = cmulc(p[j], x[k0 + j]);
                acc[i][k0 + j].x += g.x;
                acc[i][k0 + j].y += g.y;
                red_add_f32x2(rep + l, cmulc(q[j], x[k0 + j]));
              }
            }
          } else if (a.accumulate_object) {
#pragma unroll
            for (int k0 = 0; k0 < R0; k0 += GB) {
              float2 p[GB];
#pragma unroll
              for (int j = 0; j < GB; ++j) p[j] = __ldg(pm + (n2A[i] + R1 * (k0 + j)) * ND + colA[i]);
#pragma unroll
              for (int j = 0; j < GB; ++j) {
                const float2 g = cmulc(p[j], x[k0 + j]);
                acc[i][k0 + j].x += g.x;
                acc[i][k0 + j].y += g.y;
              }
            }
          } else if (rep) {
#pragma unroll
            for (int k0 = 0; k0 < R0; k0 += GB) {
              float2 q[GB];
#pragma unroll
              for (int j = 0; j < GB; ++j) q[j] = __ldcg(patch + (n2A[i] + R1 * (k0 + j)) * ND + colA[i]);
#pragma unroll
              for (int j = 0; j < GB; ++j)
                red_add_f32x2(rep + (n2A[i] + R1 * (k0 + j)) * ND + colA[i], cmulc(q[j], x[k0 + j]));
            }
          }
        }
      }
      if constexpr (TM) tmem_wait_st();
      __syncthreads();
      TB_PHASE(9);
    }

    if constexpr (VP) {
      if (a.eig_step) {
        block_sum<2>(eig, red);
        if (tid == 0) a.eig_step[s] = 0.1f * (eig[0] / eig[1]);
      }
      if constexpr (PG) {
        block_sum<4>(pg, red);
        if (tid == 0) {
          a.pos_num[2 * s] = pg[0];
          a.pos_den[2 * s] = pg[1];
          a.pos_num[2 * s + 1] = pg[2];
          a.pos_den[2 * s + 1] = pg[3];
        }
      }
    }
    // ------------- scatter-add of the object gradient -----------------------
    if (a.accumulate_object) {
      float2* G = tile;  // ND x ND, pitch ND
      const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
#pragma unroll
      for (int i = 0; i < NBA; ++i) {
        [[maybe_unused]] float v[16];
        if constexpr (TM) tmem_ld16(tacc + i * 16, v);
#pragma unroll
        for (int k = 0; k < R0; ++k) {
          const int py = n2A[i] + R1 * k, px = colA[i];  // tile coordinates
          const int y = c.iy + py - pad, x = c.ix + px - pad;
          const bool lead_ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
          float2 g;
          if constexpr (TM) g = make_float2(v[2 * k], v[2 * k + 1]);
          else g = acc[i][k];
          G[py * ND + px] = lead_ok ? cscale(g, inv_m) : make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      // footprint pixels in tile coordinates: [pad, pad + N] in both axes (the
      // gradient is zero outside the probe support, G holds zeros there)
      for (int ty = pad + warp; ty <= pad + N; ty += NWARP) {
        const int y = c.iy + ty - pad;
        if (y < 0 || y >= H) continue;
        const bool a0 = ty < pad + N, a1 = ty > pad;
        for (int tx = pad + lane; tx <= pad + N; tx += 32) {
          const int x = c.ix + tx - pad;
          if (x < 0 || x >= W) continue;
          float2 v = make_float2(0.f, 0.f);
          const bool b0 = tx < pad + N, b1 = tx > pad;
          if (a0 & b0) { const float2 g = G[ty * ND + tx];           v.x += c.w00 * g.x; v.y += c.w00 * g.y; }
          if (a0 & b1) { const float2 g = G[ty * ND + tx - 1];       v.x += c.w01 * g.x; v.y += c.w01 * g.y; }
          if (a1 & b0) { const float2 g = G[(ty - 1) * ND + tx];     v.x += c.w10 * g.x; v.y += c.w10 * g.y; }
          if (a1 & b1) { const float2 g = G[(ty - 1) * ND + tx - 1]; v.x += c.w11 * g.x; v.y += c.w11 * g.y; }
          red_add_f32x2(a.psi_num + (long)y * W + x, v);
        }
      }
    }
    // the next position's window arrives through the async proxy: order this
    // position's generic accesses to the tile before it
    if constexpr (TMAWIN) fence_proxy_async();
    __syncthreads();
    TB_PHASE(10);
  }
  TB_PHASE_FLUSH;
  if constexpr (TM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_slot, TCOLS);
  }
}

template <int ND, bool VP, bool PG, bool PO, bool PAD = false>
static int launch_fast_nd(const RpieDev& a, int grid, cudaStream_t st) {
  auto k = rpie_fast_kernel<ND, FastCfg<ND>::R0 == 8, VP, PG, PO, PAD>;
  const size_t smem = FastCfg<ND>::smem;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error((int)e, "rpie fast kernel attr: %s", cudaGetErrorString(e));
  k<<<(unsigned)grid, FastCfg<ND>::NT, smem, st>>>(a);
  return check_launch("tb_rpie_batch(fast)");
}

bool fast_kernel_applies(const RpieDev& a) {
  const tb_batch& b = a.b;
  const int nd = b.detector_width;
  const bool varying = b.eigen_weights != nullptr;
  if (a.eig_step != nullptr && !varying) return false;
  if (a.noise_model != TB_NOISE_GAUSSIAN && a.pos_num != nullptr) return false;
  if (!(nd == 32 || nd == 64 || nd == 128) || b.probe_per_position) return false;
  if (b.probe_width == nd) return true;
  // probe narrower than the detector: plain variant only (even padding)
  return ((nd - b.probe_width) % 2 == 0) && !varying && a.eig_step == nullptr &&
         a.pos_num == nullptr && a.noise_model == TB_NOISE_GAUSSIAN;
}

template <int ND>
static int launch_fast_variant(const RpieDev& a, int grid, cudaStream_t st) {
  const bool vp = a.b.eigen_weights != nullptr;
  if (a.b.probe_width != ND) return launch_fast_nd<ND, false, false, false, true>(a, grid, st);
  if (a.noise_model != TB_NOISE_GAUSSIAN)
    return vp ? launch_fast_nd<ND, true, false, true>(a, grid, st)
              : launch_fast_nd<ND, false, false, true>(a, grid, st);
  if (a.pos_num != nullptr) return launch_fast_nd<ND, true, true, false>(a, grid, st);
  if (vp) return launch_fast_nd<ND, true, false, false>(a, grid, st);
  return launch_fast_nd<ND, false, false, false>(a, grid, st);
}

int launch_fast(const RpieDev& a, int grid, cudaStream_t st) {
  switch (a.b.detector_width) {
    case 32:  return launch_fast_variant<32>(a, grid, st);
    case 64:  return launch_fast_variant<64>(a, grid, st);
    default:  return launch_fast_variant<128>(a, grid, st);
  }
}

}  // namespace tb
