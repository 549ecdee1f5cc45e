// Device-side argument block shared by the fused (rpie.cu) and the
// large-detector (large.cu) implementations of one solver batch.
#pragma once

#include "../../include/tike_b200.h"
#include "wave.cuh"

namespace tb {

struct RpieDev {
  tb_batch b;
  const void* data;
  int data_u16;
  const unsigned char* mask;
  int noise_model, step_mode;
  float step_start, step_weight;
  float unmeasured_factor;  // unmeasured_pixels_scaling - 1
  float inv_nmeasured;
  int accumulate_object;
  int divide_by_modes;      // rPIE: object gradient / M (rpie.py:450)
  float2* psi_num;
  float2* scratch;          // per-CTA scratch (patch, waves, probe numerator)
  float* costs;
  float* eig_step;
  float2* chi_out;          // lstsq: (npos, M, N, N) or nullptr
  int poisson_eps;          // lstsq.py:456 adds 1e-9 to the intensity in xi
  float* pos_num;           // lstsq position gradient sums (npos, 2) or nullptr
  float* pos_den;
  float taps[5];            // Gaussian first-derivative taps (position.py:779-810)
  int probe_sums;           // accumulate sum_s conj(o) chi into the replicas
  float2* replicas;         // (nrep, M, N, N) shared probe numerators (RED targets)
  int nrep;
  int prefetch_next;        // L2-prefetch the next position's pattern / object tile
  unsigned int* ticket;     // position counter of the persistent CTAs (zeroed per launch)
};

__device__ __forceinline__ float load_data(const void* data, int u16, long i) {
  return u16 ? (float)__ldg((const unsigned short*)data + i)
             : __ldg((const float*)data + i);
}
// same, marking the line evict_first in L2 (each pattern is read once per epoch)
__device__ __forceinline__ float load_data_stream(const void* data, int u16, long i,
                                                  uint64_t pol) {
  if (u16) {
    unsigned short v;
    asm volatile("ld.global.L2::cache_hint.u16 %0, [%1], %2;"
                 : "=h"(v)
                 : "l"((const unsigned short*)data + i), "l"(pol));
    return (float)v;
  }
  return ld_f32_hint((const float*)data + i, pol);
}


#ifndef TB_MAX_REPLICAS
#define TB_MAX_REPLICAS 16  // A/B: scripts/build_variant.py rep32 rpie.cu -DTB_MAX_REPLICAS=32
#endif
constexpr int kMaxReplicas = TB_MAX_REPLICAS;  // probe-numerator copies that take the REDs

// detector widths whose wavefront lives in shared memory (rpie.cu, rpie_fast.cu);
// every other width goes through large.cu
inline bool fused_width(int nd) { return nd == 16 || nd == 32 || nd == 64 || nd == 128; }

int check_batch(const tb_batch* b, const char* who);

// rpie.cu: detector widths 16..128, wavefront resident in shared memory
int64_t fused_workspace_bytes(const tb_batch& b, bool replica);
int run_fused(RpieDev a, int64_t workspace_bytes, void* workspace, float2* probe_out,
              cudaStream_t st, const char* who);
// rpie_fast.cu: stage-fused variant for the headline configuration
bool fast_kernel_applies(const RpieDev& a);
int launch_fast(const RpieDev& a, int grid, cudaStream_t st);
// rpie_p3.cu: three-pass variant (32 values per thread) for the plain 128 x 128 batch
bool p3_kernel_applies(const RpieDev& a);
int launch_p3(const RpieDev& a, int grid, cudaStream_t st);
// large.cu: detector widths >= 256, chunked pipeline through HBM with the
// two-pass row/column FFT
int64_t large_workspace_bytes(const tb_batch& b, bool replica, int noise_model);
int run_large(RpieDev a, int64_t workspace_bytes, void* workspace, float2* probe_out,
              cudaStream_t st, const char* who);

// large_k2r.cu: register-resident K2 (row transforms + modulus) at ND = 256
bool k2_reg_applies(const RpieDev& a);
int launch_k2_reg(const RpieDev& a, float2* wave, long s0, long count, bool need_back, int sms,
                  cudaStream_t st, const char* who);

// large_k13r.cu: register-resident K1 / K3 (column transforms) at ND = 256, plain case
bool k13_reg_applies(const RpieDev& a);
int launch_k1_reg(const RpieDev& a, float2* wave, long s0, long count, int sms, cudaStream_t st,
                  const char* who);
int launch_k3_reg(const RpieDev& a, const float2* wave, long s0, long count, int sms,
                  cudaStream_t st, const char* who);

// multislice_fused.cu: per-position fused slice loop (rPIE, D >= 2)
bool multislice_fused_applies(const tb_rpie_args& a, int nslices);
int64_t multislice_fused_workspace_bytes(const tb_batch& b, int nslices);
int run_multislice_fused(const tb_rpie_args& a, int nslices, const void* propagator,
                         cudaStream_t st);
bool multislice_precond_fused_applies(const tb_batch& b, int nslices);
int64_t multislice_precond_fused_workspace_bytes(const tb_batch& b, int nslices);
int run_multislice_precond_fused(const tb_batch& plain, int nslices, const void* propagator,
                                 void* out, void* workspace, cudaStream_t st);

__global__ void reduce_replicas_kernel(const float2* __restrict__ rep, int R, long stride,
                                       long n, float2* __restrict__ out);

}  // namespace tb
