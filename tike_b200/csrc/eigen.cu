// Variable-probe (orthogonal probe relaxation) updates of lstsq_grad for one
// batch, without the batch-sized temporaries of the reference.
//
// Replaces lstsq._update_nearplane (ptycho/solvers/lstsq.py:297-364) with
// _get_coefs_intensity (:721-736), _get_residuals (:739-746), _update_residuals
// (:749-761) and probe.update_eigen_probe (ptycho/probe.py:362-476).  There the
// patches, the per-position probe updates, the residuals R, the projections and
// phi are all (B, 1, 1, N, N) arrays; here one CTA walks one position and
// rebuilds what it needs from the object, chi and the (N, N) probes:
//
//   o_s          bilinear patch of psi at position s
//   R_s^(c)      conj(o_s) chi_s - mean probe update - sum_{c' < c} k_{s,c'} E_c'
//                (k = projection coefficients of earlier eigen probes, kept per
//                position in `coefs`)
//
// pass 1 (per eigen probe c): t_s = sum Re(conj(R_s) E_c),
//          update += R_s (t_s / N^2 + w_sc) / sum_s w_sc^2     [-> new E_c on the host side,
//                                                                 an (N, N) computation]
// pass 2: n_s = mean Re(chi conj(o E_c)), d_s = mean |o E_c|^2, k_sc = <R_s, E_c> / <E_c, E_c>
// Both passes also serve the main-probe intensity coefficients (:721-736).
#include "common.cuh"

namespace tb {

constexpr int kEigThreads = 512;

struct EigArgs {
  tb_batch b;
  const float2* chi;      // (B, M, N, N)
  int mode;
  const float2* mpu;      // (N, N) mean probe update of `mode`
  const float2* eigen;    // eigen probe 0 of `mode`; probe c at eigen + c * eigen_stride
  long eigen_stride;
  int c;                  // 1-based index of the eigen probe being updated (0: none)
  float2* coefs;          // (B, ncoef) projection coefficients
  int ncoef;
  const float* w;         // weights[s, c, mode], element stride w_stride
  long w_stride;
  const float* inv_normw;  // device scalar: 1 / sum_s w_sc^2 over the union batch
  float2* update;         // (N, N), accumulated
  float* numden;          // (B, 2) main-probe intensity sums or nullptr
  float* n_out;           // (B,)
  float* d_out;           // (B,)
};

// residual of position s at pixel idx (without the projections of probes >= c)
__device__ __forceinline__ float2 eig_residual(const EigArgs& a, long s, int idx, float2 o,
                                               float2 chi, int upto) {
  float2 r = cmulc(o, chi);
  const float2 mp = __ldg(a.mpu + idx);
  r.x -= mp.x;
  r.y -= mp.y;
  for (int k = 0; k < upto; ++k) {
    const float2 kc = a.coefs[s * a.ncoef + k];
    const float2 e = __ldg(a.eigen + k * a.eigen_stride + idx);
    const float2 p = cmul(kc, e);
    r.x -= p.x;
    r.y -= p.y;
  }
  return r;
}

// pass 1: one CTA per position, strided over the batch.  NPT = pixels per
// thread whose share of `update` is kept in registers over all positions of the
// CTA (N*N <= NPT * 512); NPT = 0: any width, one reduction per pixel and position.
template <int NPT>
__global__ void __launch_bounds__(kEigThreads)
eigen_pass1_kernel(EigArgs a) {
  __shared__ float red[3 * 32];
  const tb_batch& b = a.b;
  const int N = b.probe_width, M = b.nmodes, H = b.height, W = b.width;
  const int nn = N * N;
  const float2* psi = (const float2*)b.psi;
  const float2* p0 = (const float2*)b.probe + (long)a.mode * nn;
  const float2* ec = a.c > 0 ? a.eigen + (long)(a.c - 1) * a.eigen_stride : nullptr;
  float2 acc[NPT > 0 ? NPT : 1];
#pragma unroll
  for (int k = 0; k < (NPT > 0 ? NPT : 1); ++k) acc[k] = make_float2(0.f, 0.f);
  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner cn = make_corner(b.scan, s);
    const float2* chi = a.chi + ((long)s * M + a.mode) * nn;
    float v[3] = {0.f, 0.f, 0.f};  // t_s, intensity numerator, denominator
    for (int idx = threadIdx.x; idx < nn; idx += kEigThreads) {
      const int py = idx / N, px = idx - py * N;
      const float2 o = patch_value(psi, H, W, cn, py, px);
      const float2 x = chi[idx];
      if (a.numden) {
        const float2 op = cmul(o, __ldg(p0 + idx));
        v[1] += op.x * x.x + op.y * x.y;
        v[2] += cabs2(op);
      }
      if (ec) {
        const float2 r = eig_residual(a, s, idx, o, x, a.c - 1);
        const float2 e = __ldg(ec + idx);
        v[0] += r.x * e.x + r.y * e.y;  // Re(conj(R) E)
      }
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0 && a.numden) {
      a.numden[2 * s] = v[1];
      a.numden[2 * s + 1] = v[2];
    }
    if (ec) {
      const float ps = (v[0] / (float)nn + a.w[s * a.w_stride]) * (*a.inv_normw);
      if constexpr (NPT > 0) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
          const int idx = threadIdx.x + k * kEigThreads;
          if (idx < nn) {
            const int py = idx / N, px = idx - py * N;
            const float2 o = patch_value(psi, H, W, cn, py, px);
            const float2 r = eig_residual(a, s, idx, o, chi[idx], a.c - 1);
            acc[k].x += r.x * ps;
            acc[k].y += r.y * ps;
          }
        }
      } else {
        for (int idx = threadIdx.x; idx < nn; idx += kEigThreads) {
          const int py = idx / N, px = idx - py * N;
          const float2 o = patch_value(psi, H, W, cn, py, px);
          const float2 r = eig_residual(a, s, idx, o, chi[idx], a.c - 1);
          red_add_f32x2(a.update + idx, make_float2(r.x * ps, r.y * ps));
        }
      }
    }
  }
  if constexpr (NPT > 0) {
    if (ec) {
#pragma unroll
      for (int k = 0; k < NPT; ++k) {
        const int idx = threadIdx.x + k * kEigThreads;
        if (idx < nn) red_add_f32x2(a.update + idx, acc[k]);
      }
    }
  }
}

// pass 2: new weights of eigen probe c and its projection coefficient
__global__ void __launch_bounds__(kEigThreads)
eigen_pass2_kernel(EigArgs a) {
  __shared__ float red[6 * 32];
  const tb_batch& b = a.b;
  const int N = b.probe_width, M = b.nmodes, H = b.height, W = b.width;
  const int nn = N * N;
  const float2* psi = (const float2*)b.psi;
  const float2* ec = a.eigen + (long)(a.c - 1) * a.eigen_stride;
  for (long s = blockIdx.x; s < b.npos; s += gridDim.x) {
    const Corner cn = make_corner(b.scan, s);
    const float2* chi = a.chi + ((long)s * M + a.mode) * nn;
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // n, d, Re <R,E>, Im <R,E>, <E,E>
    for (int idx = threadIdx.x; idx < nn; idx += kEigThreads) {
      const int py = idx / N, px = idx - py * N;
      const float2 o = patch_value(psi, H, W, cn, py, px);
      const float2 x = chi[idx];
      const float2 e = __ldg(ec + idx);
      const float2 phi = cmul(o, e);
      v[0] += x.x * phi.x + x.y * phi.y;  // Re(chi conj(phi))
      v[1] += cabs2(phi);
      if (a.coefs) {
        const float2 r = eig_residual(a, s, idx, o, x, a.c - 1);
        const float2 q = cmulc(e, r);  // conj(E) R
        v[2] += q.x;
        v[3] += q.y;
        v[4] += cabs2(e);
      }
    }
    block_sum<5>(v, red);
    if (threadIdx.x == 0) {
      a.n_out[s] = v[0] / (float)nn;
      a.d_out[s] = v[1] / (float)nn;
      if (a.coefs) a.coefs[s * a.ncoef + (a.c - 1)] = make_float2(v[2] / v[4], v[3] / v[4]);
    }
  }
}

}  // namespace tb

extern "C" {

static int eig_check(const tb_batch* b, const void* chi, int mode, const char* who) {
  TB_REQUIRE(b && b->psi && b->scan && b->probe && chi, TB_ERR_INVALID, "%s: null pointer", who);
  TB_REQUIRE(b->probe_width >= 1, TB_ERR_INVALID, "%s: probe width %d", who, b->probe_width);
  TB_REQUIRE(mode >= 0 && mode < b->nmodes, TB_ERR_INVALID, "%s: mode %d of %d", who, mode,
             b->nmodes);
  TB_REQUIRE(!b->probe_per_position, TB_ERR_INVALID, "%s: needs the shared probe", who);
  return TB_OK;
}

int tb_lstsq_eigen_pass1(const tb_batch* b, const void* chi, int mode,
                         const void* m_probe_update, const void* eigen_probe,
                         int64_t eigen_stride, int c, const void* coefs, int ncoef,
                         const float* weights, int64_t weight_stride, const float* inv_norm_weights,
                         void* update, float* intensity_sums, tb_stream_t stream) {
  int rc = eig_check(b, chi, mode, "tb_lstsq_eigen_pass1");
  if (rc != TB_OK) return rc;
  TB_REQUIRE(c >= 0 && (c == 0 || (m_probe_update && eigen_probe && weights && update &&
                                   inv_norm_weights)),
             TB_ERR_INVALID, "tb_lstsq_eigen_pass1: eigen probe %d needs its arrays", c);
  TB_REQUIRE(c <= 1 || (coefs && ncoef >= c - 1), TB_ERR_INVALID,
             "tb_lstsq_eigen_pass1: projection coefficients of the earlier probes required");
  TB_REQUIRE(c > 0 || intensity_sums, TB_ERR_INVALID, "tb_lstsq_eigen_pass1: nothing to do");
  if (b->npos == 0) return TB_OK;
  tb::EigArgs a{};
  a.b = *b;
  a.chi = (const float2*)chi;
  a.mode = mode;
  a.mpu = (const float2*)m_probe_update;
  a.eigen = (const float2*)eigen_probe;
  a.eigen_stride = eigen_stride;
  a.c = c;
  a.coefs = (float2*)coefs;
  a.ncoef = ncoef;
  a.w = weights;
  a.w_stride = weight_stride;
  a.inv_normw = inv_norm_weights;
  a.update = (float2*)update;
  a.numden = intensity_sums;
  int sms = 148;
  tb_sm_count(&sms);
  const long grid = b->npos < 2L * sms ? b->npos : 2L * sms;
  const long nn = (long)b->probe_width * b->probe_width;
  cudaStream_t st = (cudaStream_t)stream;
  if (nn <= 2L * tb::kEigThreads)
    tb::eigen_pass1_kernel<2><<<(unsigned)grid, tb::kEigThreads, 0, st>>>(a);
  else if (nn <= 8L * tb::kEigThreads)
    tb::eigen_pass1_kernel<8><<<(unsigned)grid, tb::kEigThreads, 0, st>>>(a);
  else if (nn <= 32L * tb::kEigThreads)
    tb::eigen_pass1_kernel<32><<<(unsigned)grid, tb::kEigThreads, 0, st>>>(a);
  else
    tb::eigen_pass1_kernel<0><<<(unsigned)grid, tb::kEigThreads, 0, st>>>(a);
  return tb::check_launch("tb_lstsq_eigen_pass1");
}

int tb_lstsq_eigen_pass2(const tb_batch* b, const void* chi, int mode,
                         const void* m_probe_update, const void* eigen_probe,
                         int64_t eigen_stride, int c, void* coefs, int ncoef, float* n_out,
                         float* d_out, tb_stream_t stream) {
  int rc = eig_check(b, chi, mode, "tb_lstsq_eigen_pass2");
  if (rc != TB_OK) return rc;
  TB_REQUIRE(c >= 1 && eigen_probe && n_out && d_out, TB_ERR_INVALID,
             "tb_lstsq_eigen_pass2: null pointer");
  TB_REQUIRE(!coefs || (m_probe_update && ncoef >= c), TB_ERR_INVALID,
             "tb_lstsq_eigen_pass2: coefficient table too small");
  if (b->npos == 0) return TB_OK;
  tb::EigArgs a{};
  a.b = *b;
  a.chi = (const float2*)chi;
  a.mode = mode;
  a.mpu = (const float2*)m_probe_update;
  a.eigen = (const float2*)eigen_probe;
  a.eigen_stride = eigen_stride;
  a.c = c;
  a.coefs = (float2*)coefs;
  a.ncoef = ncoef;
  a.n_out = n_out;
  a.d_out = d_out;
  int sms = 148;
  tb_sm_count(&sms);
  const long grid = b->npos < 4L * sms ? b->npos : 4L * sms;
  tb::eigen_pass2_kernel<<<(unsigned)grid, tb::kEigThreads, 0, (cudaStream_t)stream>>>(a);
  return tb::check_launch("tb_lstsq_eigen_pass2");
}

}  // extern "C"
