// K2 of the large-detector pipeline at ND = 256 (and 512, below), register-resident variant:
// forward row transforms + |Psi|^2 over modes + cost + modulus factor + inverse
// row transforms of all modes of an 8-row block (same contract as
// large_rows_modulus_kernel<256> in large_fused.cu: `wave` in, `wave` out, rows
// in K1's slot order, columns natural).
//
// A 256-point row transform is two radix-16 stages.  Each thread owns 16 values
// per stage and takes them straight from / to global memory, so the tile is
// written once and read once per transform (3 + 3 in the generic kernel):
//   forward   global -> radix-16 over k (elements n2 + 16 k) -> twiddle -> tile
//             tile -> radix-16 over n2 (elements 16 k1 + n2) -> registers
//   inverse   registers -> radix-16^-1 -> tile -> conj twiddle -> radix-16^-1 -> global
// The far field only exists in the second ownership (row r, block k1), so what
// is keyed by frequency is private to a thread: the intensity / factor values
// stay in 16 registers, the last mode stays in registers between the two
// transforms, the other modes are spilled IN PLACE in a thread-major layout
// (coalesced) and the measured pattern is read directly (16 consecutive
// frequencies per half warp).  The input of the next mode (or of the next row
// block) is fetched into registers before the current one is transformed.
// The tile pads one element per 16 (offset c + c / 16): both ownerships are
// conflict free for 64-bit accesses with lanes = n2 resp. k1.
// (Measured and dropped: parking the far fields of modes 0 .. M - 2 in shared
// memory instead of the in-place spill -- 17.3 vs 17.4 ms per lstsq_grad epoch
// of 4000 positions: the 16 KB a CTA spills and reloads stay in L2 anyway.)
// Replaces (with K1, K3): rpie.py:355-505, lstsq.py:422-579, objective.py:11-66.
#include "solver_dev.cuh"
#include "dft32.cuh"

namespace tb {

namespace k2r {
constexpr int ND = 256, VR = 8, NT = 128, PR = 272, NRB = ND / VR;
constexpr size_t kSmem = (size_t)VR * PR * 8 + ND * 8 + 32 * 4;
__device__ __forceinline__ int sidx(int r, int c) { return r * PR + c + (c >> 4); }
}  // namespace k2r

#ifndef TB_K2R_CTAS
#define TB_K2R_CTAS 4  // CTAs per SM the register allocation is bounded for (A/B)
#endif
namespace k2r {
// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) and their mbarrier ---------
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
      "K2R_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra K2R_DONE_%=;\n"
      "bra K2R_WAIT_%=;\n"
      "K2R_DONE_%=:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes,
                                              unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"((unsigned)__cvta_generic_to_shared(dst)),
      "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}
constexpr int DEPTH = 2;                       // input tiles in flight per CTA
constexpr int TILE_BYTES = VR * ND * 8;        // 8 rows of one mode: contiguous in `wave`
constexpr size_t kSmemRing = kSmem + (size_t)DEPTH * TILE_BYTES + 64;
}  // namespace k2r

// RING = the input tiles (8 contiguous rows of one mode, 16 KB) arrive through a
// ring of DEPTH shared-memory buffers filled by the TMA unit (one cp.async.bulk
// per tile, completion on an mbarrier) DEPTH tiles ahead of their transform,
// instead of through registers one tile ahead.  Same speed as the register
// look-ahead (the kernel's long_scoreboard stalls, profiles/r02ag_*, are not on
// these loads); kept selectable (TB_LARGE_K2R_TMA=1)
template <bool RING>
__global__ void __launch_bounds__(k2r::NT, TB_K2R_CTAS)
large_rows_modulus_reg_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count,
                              int need_back) {
  using namespace k2r;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + VR * PR;
  float* red = reinterpret_cast<float*>(tw + ND);
  [[maybe_unused]] unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + 32);
  [[maybe_unused]] float2* ring = reinterpret_cast<float2*>(bars + 8);  // 16-byte aligned
  // tw[k * 16 + h] = w256^(h k): lanes (h) read consecutive entries (a plain
  // w256^n table indexed h * k costs up to 8-way bank conflicts, profiles/r02ag_*)
  for (int n = threadIdx.x; n < ND; n += NT) {
    float sn, cs;
    sincospif(2.0f * (float)((n & 15) * (n >> 4)) / (float)ND, &sn, &cs);
    tw[n] = make_float2(cs, -sn);
  }
  if constexpr (RING) {
    if (threadIdx.x == 0) {
      for (int d = 0; d < DEPTH; ++d) mbar_init(bars + d, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const int tid = threadIdx.x, h = tid & 15, r = tid >> 4;
  float2* const tA = tile + sidx(r, h);        // + 17 k: element h + 16 k
  float2* const tB = tile + sidx(r, 16 * h);   // + n: element 16 h + n
  const long total = count * NRB;
  const long img_off = (long)r * ND + h;       // natural layout: + 16 k

  // tile q of this CTA = mode q % M of its task number q / M
  [[maybe_unused]] auto issue_tile = [&](long q) {
    const long tq = blockIdx.x + (q / M) * (long)gridDim.x;
    if (tq >= total) return;
    const float2* src = wave + (tq / NRB) * M * (long)ND * ND + (q % M) * (long)ND * ND +
                        (tq % NRB) * (long)VR * ND;
    const int slot = (int)(q % DEPTH);
    mbar_expect_tx(bars + slot, TILE_BYTES);
    bulk_copy_g2s(ring + (long)slot * VR * ND, src, TILE_BYTES, bars + slot);
  };
  [[maybe_unused]] long q_tile = 0;
  float2 nx[16];  // input of the next forward transform, fetched ahead
  if constexpr (RING) {
    if (tid == 0)
      for (int d = 0; d < DEPTH; ++d) issue_tile(d);
  } else {
    if ((long)blockIdx.x < total) {
      const long t = blockIdx.x;
      const float2* img = wave + (t / NRB) * M * (long)ND * ND + (t % NRB) * (long)VR * ND + img_off;
#pragma unroll
      for (int k = 0; k < 16; ++k) nx[k] = __ldcs(img + 16 * k);
    }
  }
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int rb = (int)(t % NRB);
    const long i = t / NRB;
    const long s = s0 + i;
    float2* base = wave + i * M * (long)ND * ND + (long)rb * VR * ND;
    float F[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) F[p] = 0.f;
    float2 last[16];
    for (int m = 0; m < M; ++m) {
      float2* img = base + (long)m * ND * ND;
      float2 x[16];
      if constexpr (RING) {
        const int slot = (int)(q_tile % DEPTH);
        mbar_wait(bars + slot, (unsigned)((q_tile / DEPTH) & 1));
        const float2* in = ring + (long)slot * VR * ND + img_off;
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = in[16 * k];
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = nx[k];
      }
      if constexpr (!RING) {  // next mode of this block, else the first mode of this CTA's next block
        const float2* nimg = nullptr;
        if (m + 1 < M) {
          nimg = img + (long)ND * ND + img_off;
        } else if (t + gridDim.x < total) {
          const long tn = t + gridDim.x;
          nimg = wave + (tn / NRB) * M * (long)ND * ND + (tn % NRB) * (long)VR * ND + img_off;
        }
        if (nimg) {
#pragma unroll
          for (int k = 0; k < 16; ++k) nx[k] = __ldcs(nimg + 16 * k);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) nx[k] = make_float2(0.f, 0.f);
        }
      }
      dft<16>(x);
      tA[0] = x[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) tA[17 * k] = cmul(x[k], tw[16 * k + h]);
      __syncthreads();
      if constexpr (RING) {  // every thread has read the slot: refill it
        if (tid == 0) issue_tile(q_tile + DEPTH);
        ++q_tile;
      }
      float2 y[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) y[n] = tB[n];
      dft<16>(y);
#pragma unroll
      for (int p = 0; p < 16; ++p) F[p] += cabs2(y[p]) * s2;
      if (m == M - 1) {
#pragma unroll
        for (int p = 0; p < 16; ++p) last[p] = y[p];
      } else {
        if (need_back) {  // in place, thread-major: every thread has consumed its input
#pragma unroll
          for (int p = 0; p < 16; ++p) img[p * NT + tid] = y[p];
        }
#pragma unroll
        for (int p = 0; p < 16; ++p) last[p] = make_float2(0.f, 0.f);
      }
      __syncthreads();
    }
    // cost and modulus factor (objective.py:11-66): this thread's frequencies
    // are row l2f(rb * VR + r), columns h + 16 p
    {
      const long rowpix = (long)loc2freq<ND>(rb * VR + r) * ND + h;
      float d[16];
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const long pix = rowpix + 16 * p;
        const bool meas = a.mask ? (a.mask[pix] != 0) : true;
        d[p] = meas ? load_data(a.data, a.data_u16, s * (long)ND * ND + pix) : -1.0f;
      }
      float sums[1] = {0.f};
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        if (d[p] >= 0.f) {
          const float sd = sqrtf(d[p]), sI = sqrtf(F[p]);
          const float dv = sI - sd;
          sums[0] += dv * dv;
          F[p] = -(1.0f - sd / (sI + 1e-9f)) * rt;
        } else {
          F[p] = a.unmeasured_factor * rt;
        }
      }
      block_sum<1>(sums, red);
      if (tid == 0) atomicAdd(a.costs + s, sums[0] * a.inv_nmeasured);
    }
    if (!need_back) continue;
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi == 0) ? M - 1 : mi - 1;  // the last mode is in registers
      float2* img = base + (long)m * ND * ND;
      float2 y[16];
      if (mi == 0) {
#pragma unroll
        for (int p = 0; p < 16; ++p) y[p] = last[p];
      } else {
#pragma unroll
        for (int p = 0; p < 16; ++p) y[p] = img[p * NT + tid];
      }
#pragma unroll
      for (int p = 0; p < 16; ++p) y[p] = cscale(y[p], F[p]);
      idft<16>(y);
#pragma unroll
      for (int n = 0; n < 16; ++n) tB[n] = y[n];
      __syncthreads();
      float2 x[16];
      x[0] = tA[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) x[k] = cmulc(tw[16 * k + h], tA[17 * k]);
      idft<16>(x);
#pragma unroll
      for (int k = 0; k < 16; ++k) img[img_off + 16 * k] = x[k];
      __syncthreads();
    }
  }
}

// ---- ND = 512 ----------------------------------------------------------------
// A 512-point row is a radix-16 stage (elements n2 + 32 k; two butterflies per
// thread, one after the other) and a radix-32 stage in registers (elements
// 32 k1 + n): radix-2 with the constant twiddles w32^n, then two radix-16.  Slot
// p of the radix-32 output holds frequency 2 p (p < 16) or 2 (p - 16) + 1 within
// the block.  The tile pads one element per 32 (offset c + c / 32).  All modes
// are spilled in place (thread-major) unless there is only one, which stays in
// registers (BASELINE config 5).
namespace k2r512 {
constexpr int ND = 512, VR = 8, NT = 128, PR = 528, NRB = ND / VR;
constexpr size_t kSmem = (size_t)VR * PR * 8 + ND * 8 + 32 * 4;
__device__ __forceinline__ int sidx(int r, int c) { return r * PR + c + (c >> 5); }
}  // namespace k2r512

#ifndef TB_K2R512_CTAS
#define TB_K2R512_CTAS 4
#endif
__global__ void __launch_bounds__(k2r512::NT, TB_K2R512_CTAS)
large_rows_modulus_reg512_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count,
                                 int need_back, int rows_16x32) {
  using namespace k2r512;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + VR * PR;
  float* red = reinterpret_cast<float*>(tw + ND);
  // tw[(q * 16 + k) * 16 + h] = w512^((h + 16 q) k): conflict free for lanes = h
  for (int n = threadIdx.x; n < ND; n += NT) {
    const int hh = n & 15, kk = (n >> 4) & 15, qq = n >> 8;
    float sn, cs;
    sincospif(2.0f * (float)((hh + 16 * qq) * kk) / (float)ND, &sn, &cs);
    tw[n] = make_float2(cs, -sn);
  }
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes;
  const float s2 = b.fwd_scale * b.fwd_scale;
  const float rt = b.fwd_scale * b.inv_scale;
  const int tid = threadIdx.x, h = tid & 15, r = tid >> 4;
  float2* const tA = tile + sidx(r, h);        // + 16 q + 33 k: element h + 16 q + 32 k
  float2* const tB = tile + sidx(r, 32 * h);   // + n: element 32 h + n
  const long total = count * NRB;
  const long img_off = (long)r * ND + h;       // natural layout: + 16 q + 32 k
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int rb = (int)(t % NRB);
    const long i = t / NRB;
    const long s = s0 + i;
    float2* base = wave + i * M * (long)ND * ND + (long)rb * VR * ND;
    float F[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) F[p] = 0.f;
    float2 y[32];  // far field of the mode in flight (the only one when M == 1)
    for (int m = 0; m < M; ++m) {
      float2* img = base + (long)m * ND * ND;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float2 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = __ldcs(img + img_off + 16 * q + 32 * k);
        dft<16>(x);
        tA[16 * q] = x[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) tA[16 * q + 33 * k] = cmul(x[k], tw[(q * 16 + k) * 16 + h]);
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < 32; ++n) y[n] = tB[n];
      dft32(y);
#pragma unroll
      for (int p = 0; p < 32; ++p) F[p] += cabs2(y[p]) * s2;
      if (M > 1 && need_back) {  // in place, thread-major: every thread has consumed its input
#pragma unroll
        for (int p = 0; p < 32; ++p) img[p * NT + tid] = y[p];
      }
      __syncthreads();
    }
    // cost and modulus factor (objective.py:11-66): this thread's frequencies
    // are row l2f(rb * VR + r), columns h + 16 f(p)
    {
      // row slot -> row frequency: the plan of the kernel that made the column
      // transforms (generic K1: 8 x 8 x 8; large_k13r.cu: 16 x 32 with dft32 slots)
      const int slot = rb * VR + r;
      const long rowpix = (long)(rows_16x32 ? (slot >> 5) + 16 * dft32_freq(slot & 31)
                                            : loc2freq<ND>(slot)) * ND + h;
      float sums[1] = {0.f};
#pragma unroll
      for (int p0 = 0; p0 < 32; p0 += 8) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int p = p0 + j;
          const long pix = rowpix + 16 * (p < 16 ? 2 * p : 2 * (p - 16) + 1);
          const bool meas = a.mask ? (a.mask[pix] != 0) : true;
          d[j] = meas ? load_data(a.data, a.data_u16, s * (long)ND * ND + pix) : -1.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int p = p0 + j;
          if (d[j] >= 0.f) {
            const float sd = sqrtf(d[j]), sI = sqrtf(F[p]);
            const float dv = sI - sd;
            sums[0] += dv * dv;
            F[p] = -(1.0f - sd / (sI + 1e-9f)) * rt;
          } else {
            F[p] = a.unmeasured_factor * rt;
          }
        }
      }
      block_sum<1>(sums, red);
      if (tid == 0) atomicAdd(a.costs + s, sums[0] * a.inv_nmeasured);
    }
    if (!need_back) continue;
    for (int mi = 0; mi < M; ++mi) {
      const int m = M - 1 - mi;  // the last mode first: it is still in registers
      float2* img = base + (long)m * ND * ND;
      if (mi > 0) {
#pragma unroll
        for (int p = 0; p < 32; ++p) y[p] = img[p * NT + tid];
      }
#pragma unroll
      for (int p = 0; p < 32; ++p) y[p] = cscale(y[p], F[p]);
      idft32(y);
#pragma unroll
      for (int n = 0; n < 32; ++n) tB[n] = y[n];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float2 x[16];
        x[0] = tA[16 * q];
#pragma unroll
        for (int k = 1; k < 16; ++k) x[k] = cmulc(tw[(q * 16 + k) * 16 + h], tA[16 * q + 33 * k]);
        idft<16>(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) img[img_off + 16 * q + 32 * k] = x[k];
      }
      __syncthreads();
    }
  }
}

bool k2_reg_applies(const RpieDev& a) {
  static const bool on = [] {
    const char* e = getenv("TB_LARGE_K2R");  // 0: keep the generic K2 (A/B timing)
    return e ? atoi(e) != 0 : true;
  }();
  return on && (a.b.detector_width == 256 || a.b.detector_width == 512);
}

int launch_k2_reg(const RpieDev& a, float2* wave, long s0, long count, bool need_back, int sms,
                  cudaStream_t st, const char* who) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(large_rows_modulus_reg_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)k2r::kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(large_rows_modulus_reg_kernel<true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)k2r::kSmemRing);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(large_rows_modulus_reg512_kernel,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2r512::kSmem);
    if (e != cudaSuccess) return set_error((int)e, "%s: kernel attributes: %s", who, cudaGetErrorString(e));
    configured = true;
  }
  if (a.b.detector_width == 512) {
    const long tasks = count * k2r512::NRB;
    const long g = tasks < (long)sms * TB_K2R512_CTAS ? tasks : (long)sms * TB_K2R512_CTAS;
    large_rows_modulus_reg512_kernel<<<(unsigned)g, k2r512::NT, k2r512::kSmem, st>>>(
        a, wave, s0, count, need_back ? 1 : 0, k13_reg_applies(a) ? 1 : 0);
    return check_launch(who);
  }
  const long tasks = count * k2r::NRB;
  const long g = tasks < (long)sms * TB_K2R_CTAS ? tasks : (long)sms * TB_K2R_CTAS;
  // 1: input tiles through the TMA ring instead of registers (measured equal:
  // 17.2 vs 17.1 ms per lstsq_grad epoch of 4000 positions; read per launch so
  // that the tests can compare the two)
  const char* ring_env = getenv("TB_LARGE_K2R_TMA");
  const bool use_ring = ring_env ? atoi(ring_env) != 0 : false;
  if (use_ring)
    large_rows_modulus_reg_kernel<true><<<(unsigned)g, k2r::NT, k2r::kSmemRing, st>>>(
        a, wave, s0, count, need_back ? 1 : 0);
  else
    large_rows_modulus_reg_kernel<false><<<(unsigned)g, k2r::NT, k2r::kSmem, st>>>(
        a, wave, s0, count, need_back ? 1 : 0);
  return check_launch(who);
}

}  // namespace tb
