// lstsq_grad (Odstrcil 2018) kernels.
//
// Phase 1 reuses the fused rPIE pipeline (rpie.cu) with three differences:
// the back-propagated residual chi is spilled to HBM because phase 2 needs it
// after a batch-wide reduction (lstsq.py:394-410, 504-507), the common object
// gradient has no 1/M (lstsq.py:512-520) and the position-gradient sums of
// lstsq.py:545-579 are produced on request.  Patches and unique probes are
// recomputed in phase 2 instead of being stored.
//
// Phase 2 = the per-position sums of _precondition_nearplane_gradients
// (lstsq.py:619-718); the 2x2 solve and the batch means are O(npos) and stay
// on the host side of the C-ABI.
#include "../../include/tike_b200.h"
#include "wave.cuh"

namespace tb {

struct RpieDev;
int check_batch(const tb_batch* b, const char* who);

// out[s] = {A1, A4, b1, b2, Re A2, Im A2}
__global__ void __launch_bounds__(256)
lstsq_phase2_kernel(tb_batch b, const float2* __restrict__ chi,
                    const float2* __restrict__ object_update,
                    const float2* __restrict__ m_probe_update, int mode,
                    float eps, float* __restrict__ out) {
  __shared__ float red[6 * 32];
  ProbeSet ps;
  ps.probe = (const float2*)b.probe;
  ps.eigen = (const float2*)b.eigen_probe;
  ps.weights = b.eigen_weights;
  ps.M = b.nmodes; ps.N = b.probe_width; ps.E = b.neigen; ps.Me = b.eigen_modes;
  ps.per_position = b.probe_per_position;
  const int N = b.probe_width, M = b.nmodes;
  const float2* psi = (const float2*)b.psi;
  const bool simple = ps.weights == nullptr && !ps.per_position;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner c = make_corner(b.scan, s);
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float2* chi0 = chi + ((long)s * M + mode) * N * N;
    const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < b.height) & (c.ix + N < b.width);
    if (simple && interior && object_update && m_probe_update) {
      // common case: shared probe, patch inside the object, both updates.
      // A warp walks a row (coalesced), two rows per iteration for more
      // independent loads in flight; no divisions, no bounds logic.
      const int W = b.width;
      const float2* __restrict__ pm = ps.probe + (long)mode * N * N;
      const float2* __restrict__ ou = object_update + (long)c.iy * W + c.ix;
      const float2* __restrict__ ob = psi + (long)c.iy * W + c.ix;
      for (int py = 2 * warp; py < N; py += 2 * nwarp) {
        for (int px = lane; px < N; px += 32) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int y = py + r;
            if (y >= N) break;
            const float2 x = __ldcs(chi0 + y * N + px);
            const float2 p = __ldg(pm + y * N + px), dp = __ldg(m_probe_update + y * N + px);
            const float2* u = ou + (long)y * W + px;
            const float2* o = ob + (long)y * W + px;
            const float2 u00 = __ldg(u), u01 = __ldg(u + 1), u10 = __ldg(u + W), u11 = __ldg(u + W + 1);
            const float2 o00 = __ldg(o), o01 = __ldg(o + 1), o10 = __ldg(o + W), o11 = __ldg(o + W + 1);
            float2 du, ov;
            du.x = u00.x * c.w00; du.y = u00.y * c.w00;
            du.x += u01.x * c.w01; du.y += u01.y * c.w01;
            du.x += u10.x * c.w10; du.y += u10.y * c.w10;
            du.x += u11.x * c.w11; du.y += u11.y * c.w11;
            ov.x = o00.x * c.w00; ov.y = o00.y * c.w00;
            ov.x += o01.x * c.w01; ov.y += o01.y * c.w01;
            ov.x += o10.x * c.w10; ov.y += o10.y * c.w10;
            ov.x += o11.x * c.w11; ov.y += o11.y * c.w11;
            const float2 dop = cmul(du, p), dpo = cmul(dp, ov);
            v[0] += cabs2(dop) + eps;
            v[2] += dop.x * x.x + dop.y * x.y;
            v[4] += dop.x * dpo.x + dop.y * dpo.y;
            v[5] += dop.y * dpo.x - dop.x * dpo.y;
            v[1] += cabs2(dpo) + eps;
            v[3] += dpo.x * x.x + dpo.y * x.y;
          }
        }
      }
    } else
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      const int py = idx / N, px = idx - py * N;
      const float2 x = chi0[idx];
      if (object_update) {
        const float2 dop = cmul(patch_value(object_update, b.height, b.width, c, py, px),
                                probe_value(ps, s, mode, py, px));
        v[0] += cabs2(dop) + eps;
        v[2] += dop.x * x.x + dop.y * x.y;
        if (m_probe_update) {
          const float2 dpo = cmul(__ldg(m_probe_update + idx),
                                  patch_value(psi, b.height, b.width, c, py, px));
          // dOP * conj(dPO)
          v[4] += dop.x * dpo.x + dop.y * dpo.y;
          v[5] += dop.y * dpo.x - dop.x * dpo.y;
        }
      }
      if (m_probe_update) {
        const float2 dpo = cmul(__ldg(m_probe_update + idx),
                                patch_value(psi, b.height, b.width, c, py, px));
        v[1] += cabs2(dpo) + eps;
        v[3] += dpo.x * x.x + dpo.y * x.y;
      }
    }
    block_sum<6>(v, red);
    if (threadIdx.x < 6) {
      float r = v[0];
      if (threadIdx.x == 1) r = v[1];
      if (threadIdx.x == 2) r = v[2];
      if (threadIdx.x == 3) r = v[3];
      if (threadIdx.x == 4) r = v[4];
      if (threadIdx.x == 5) r = v[5];
      out[s * 6 + threadIdx.x] = r;
    }
  }
}

__global__ void __launch_bounds__(256)
max_real_kernel2(const float2* __restrict__ x, long n, float* __restrict__ out) {
  float m = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));
}

__global__ void __launch_bounds__(256)
precondition_object_kernel(float2* __restrict__ out, const float2* __restrict__ upd,
                           const float2* __restrict__ precond, long n, float alpha,
                           const float* __restrict__ maxv) {
  const float am = alpha * (*maxv);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    // sqrt(((1-alpha) d)^2 + (alpha max d)^2) with d real-valued complex
    const float d = (1.0f - alpha) * precond[i].x;
    const float den = sqrtf(d * d + am * am);
    const float2 g = upd[i];
    out[i] = make_float2(g.x / den, g.y / den);
  }
}

__global__ void __launch_bounds__(256)
caxpy_kernel(float2* __restrict__ y, const float2* __restrict__ x, long n, float a,
             const float* __restrict__ a_dev) {
  const float s = a_dev ? a * (*a_dev) : a;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    float2 v = y[i];
    const float2 u = x[i];
    v.x += s * u.x;
    v.y += s * u.y;
    y[i] = v;
  }
}

}  // namespace tb

extern "C" {

int tb_lstsq_phase2(const tb_batch* b, const void* chi, const void* object_update,
                    const void* m_probe_update, int mode, float eps, float* out,
                    tb_stream_t stream) {
  int rc = tb::check_batch(b, "tb_lstsq_phase2");
  if (rc != TB_OK) return rc;
  TB_REQUIRE(chi && out, TB_ERR_INVALID, "tb_lstsq_phase2: null pointer");
  TB_REQUIRE(mode >= 0 && mode < b->nmodes, TB_ERR_INVALID, "tb_lstsq_phase2: bad mode");
  if (b->npos == 0) return TB_OK;
  tb_batch bb = *b;
  if (bb.eigen_probe == nullptr) bb.neigen = 0;
  int sms = 148;
  tb_sm_count(&sms);
  long grid = (long)sms * 8;
  if (bb.npos < grid) grid = bb.npos;
  tb::lstsq_phase2_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      bb, (const float2*)chi, (const float2*)object_update,
      (const float2*)m_probe_update, mode, eps, out);
  return tb::check_launch("tb_lstsq_phase2");
}

int tb_lstsq_precondition_object(void* out, const void* object_upd, const void* precond,
                                 int64_t n, float alpha, float* scratch,
                                 tb_stream_t stream) {
  TB_REQUIRE(out && object_upd && precond && scratch, TB_ERR_INVALID,
             "tb_lstsq_precondition_object: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, sizeof(float), st);
  const long blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 2368 ? blocks : 2368);
  tb::max_real_kernel2<<<grid, 256, 0, st>>>((const float2*)precond, n, scratch);
  tb::precondition_object_kernel<<<grid, 256, 0, st>>>(
      (float2*)out, (const float2*)object_upd, (const float2*)precond, n, alpha, scratch);
  return tb::check_launch("tb_lstsq_precondition_object");
}

int tb_caxpy(void* y, const void* x, int64_t n, float a, const float* a_dev,
             tb_stream_t stream) {
  TB_REQUIRE(y && x, TB_ERR_INVALID, "tb_caxpy: null pointer");
  if (n == 0) return TB_OK;
  const long blocks = (n + 255) / 256;
  tb::caxpy_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(
      (float2*)y, (const float2*)x, n, a, a_dev);
  return tb::check_launch("tb_caxpy");
}

}  // extern "C"
