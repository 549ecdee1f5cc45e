// Object-sized update and constraint kernels: everything that runs over the
// (D, H, W) object once per batch or once per epoch outside the fused batch
// pipeline, as single passes instead of chains of library elementwise ops.
//
// Replaces: rpie._update with adaptive moments (rpie.py:233-267, opt.py:165-213),
// lstsq_grad's momentum object step (lstsq.py:176-193, opt.py:67-82), the
// per-epoch object constraints (ptycho.py:811-851, object.py:208-253, 324-335)
// and the "accumulate, one update per epoch" DM step (solvers/dm.py).  The
// *_given_max entry points take max(preconditioner) from the caller: with the
// object rows split over ranks (communicators/comm.py, RowPlan) the maximum is
// a cross-rank quantity.
#include "common.cuh"

namespace tb {

constexpr int kUpdThreads = 256;
constexpr unsigned kUpdMaxGrid = 2368;  // 16 CTAs of 256 threads per SM

static inline unsigned upd_grid(long n) {
  const long blocks = (n + kUpdThreads - 1) / kUpdThreads;
  return (unsigned)(blocks < 1 ? 1 : (blocks < (long)kUpdMaxGrid ? blocks : (long)kUpdMaxGrid));
}

#define TB_GRID_STRIDE(i, n)                                              \
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (n);   \
       i += (long)gridDim.x * blockDim.x)

// out = max(out, max_i Re x_i); values are >= 0 (sums of squared moduli), so
// the float ordering equals the ordering of the bit patterns
__global__ void __launch_bounds__(kUpdThreads)
upd_max_real_kernel(const float2* __restrict__ x, long n, float* __restrict__ out) {
  float m = 0.f;
  TB_GRID_STRIDE(i, n) m = fmaxf(m, x[i].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));
}

// complex quotient g / (dr + i di)
__device__ __forceinline__ float2 cdiv(float2 g, float dr, float di) {
  if (di == 0.f) return make_float2(g.x / dr, g.y / dr);
  const float dd = dr * dr + di * di;
  return make_float2((g.x * dr + g.y * di) / dd, (g.y * dr - g.x * di) / dd);
}

// rpie.py:233-238 with the maximum supplied
__global__ void __launch_bounds__(kUpdThreads)
upd_rpie_psi_kernel(float2* __restrict__ psi, const float2* __restrict__ num,
                    const float2* __restrict__ precond, long n, float alpha,
                    const float* __restrict__ maxv) {
  const float mx = *maxv;
  TB_GRID_STRIDE(i, n) {
    const float2 pc = precond[i];
    const float2 q = cdiv(num[i], (1.0f - alpha) * pc.x + alpha * mx, (1.0f - alpha) * pc.y);
    float2 p = psi[i];
    p.x += q.x;
    p.y += q.y;
    psi[i] = p;
  }
}

// rpie.py:233-267 without `errors`: plain step, ADAM moments of the numerator
// (opt.py:207-213: no bias-correction power), second step through the same
// denominator.  v is real, m complex.
__global__ void __launch_bounds__(kUpdThreads)
upd_rpie_adam_kernel(float2* __restrict__ psi, const float2* __restrict__ num,
                     const float2* __restrict__ precond, float* __restrict__ v,
                     float2* __restrict__ m, long n, float alpha, float vdecay, float omv,
                     float mdecay, float omm, const float* __restrict__ maxv) {
  // omv = 1 - vdecay, omm = 1 - mdecay rounded from double like the scalar
  // operands of the reference's array expressions
  const float mx = *maxv;
  const float eps = 1e-8f;
  TB_GRID_STRIDE(i, n) {
    const float2 pc = precond[i];
    const float dr = (1.0f - alpha) * pc.x + alpha * mx, di = (1.0f - alpha) * pc.y;
    const float2 g = num[i];
    float2 mm = m[i];
    mm.x = mdecay * mm.x + omm * g.x;
    mm.y = mdecay * mm.y + omm * g.y;
    const float vv = vdecay * v[i] + omv * cabs2(g);
    m[i] = mm;
    v[i] = vv;
    const float den = sqrtf(vv / omv) + eps;
    const float2 d = make_float2(mm.x / omm / den, mm.y / omm / den);
    const float2 q0 = cdiv(g, dr, di), q1 = cdiv(d, dr, di);
    float2 p = psi[i];
    p.x = (p.x + q0.x) + q1.x;
    p.y = (p.y + q0.y) + q1.y;
    psi[i] = p;
  }
}

// lstsq.py:176-193 with opt.momentum (opt.py:67-82): m = mdecay m + (1 - mdecay)
// beta x ; psi += m.  beta is a device scalar (batch mean of the step lengths).
__global__ void __launch_bounds__(kUpdThreads)
upd_momentum_kernel(float2* __restrict__ psi, const float2* __restrict__ x,
                    float2* __restrict__ m, long n, float mdecay, float omm,
                    const float* __restrict__ beta) {
  const float b = *beta;
  TB_GRID_STRIDE(i, n) {
    const float2 g = x[i];
    float2 mm = m[i];
    mm.x = mdecay * mm.x + omm * (b * g.x);
    mm.y = mdecay * mm.y + omm * (b * g.y);
    m[i] = mm;
    float2 p = psi[i];
    p.x += mm.x;
    p.y += mm.y;
    psi[i] = p;
  }
}

__global__ void __launch_bounds__(kUpdThreads)
upd_lstsq_precondition_kernel(float2* __restrict__ out, const float2* __restrict__ upd,
                              const float2* __restrict__ precond, long n, float alpha,
                              const float* __restrict__ maxv) {
  const float am = alpha * (*maxv);
  TB_GRID_STRIDE(i, n) {
    const float d = (1.0f - alpha) * precond[i].x;
    const float den = sqrtf(d * d + am * am);
    const float2 g = upd[i];
    out[i] = make_float2(g.x / den, g.y / den);
  }
}

// y += num / (Re precond + eps)   (solvers/dm.py)
__global__ void __launch_bounds__(kUpdThreads)
upd_add_quotient_kernel(float2* __restrict__ y, const float2* __restrict__ num,
                        const float2* __restrict__ precond, long n, long period, float eps) {
  TB_GRID_STRIDE(i, n) {
    const float den = precond[period > 0 ? i % period : i].x + eps;
    const float2 g = num[i];
    float2 p = y[i];
    p.x += g.x / den;
    p.y += g.y / den;
    y[i] = p;
  }
}

// positivity (object.py:208-224: r |x| + (1 - r) x) and clip_magnitude
// (ptycho.py:257-262) in one pass, in this order like ptycho.py:811-851
__global__ void __launch_bounds__(kUpdThreads)
upd_object_pointwise_kernel(float2* __restrict__ psi, long n, float positivity,
                            int clip, float a_max) {
  TB_GRID_STRIDE(i, n) {
    float2 p = psi[i];
    if (positivity > 0.f) {
      const float mag = sqrtf(cabs2(p));
      p.x = positivity * mag + (1.0f - positivity) * p.x;
      p.y = (1.0f - positivity) * p.y;
    }
    if (clip) {
      const float mag = sqrtf(cabs2(p));
      if (mag > a_max) {
        const float s = a_max / mag;
        p.x *= s;
        p.y *= s;
      }
    }
    psi[i] = p;
  }
}

// object.py:227-253: a on the eight neighbours, 1 - 8a in the centre, edges
// replicated; one thread per output pixel, rows streamed through L1/L2
__global__ void __launch_bounds__(kUpdThreads)
upd_smooth_kernel(float2* __restrict__ out, const float2* __restrict__ in, int D, int H,
                  int W, float a) {
  const long n = (long)D * H * W;
  const float c = 1.0f - 8.0f * a;
  TB_GRID_STRIDE(i, n) {
    const int x = (int)(i % W);
    const long r = i / W;
    const int y = (int)(r % H);
    const float2* img = in + (r / H) * (long)H * W;
    const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
    const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
    float sx = 0.f, sy = 0.f;
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (j == 1 && k == 1) continue;
        const float2 v = __ldg(img + (long)ys[j] * W + xs[k]);
        sx += v.x;
        sy += v.y;
      }
    const float2 v0 = __ldg(img + (long)y * W + x);
    out[i] = make_float2(a * sx + c * v0.x, a * sy + c * v0.y);
  }
}

// remove_object_ambiguity (object.py:324-335): out[0] += sum |psi|^2 Re W,
// out[1] += sum (Re W)^2 in double
__global__ void __launch_bounds__(kUpdThreads)
upd_weighted_norm_kernel(const float2* __restrict__ psi, const float2* __restrict__ w,
                         long n, double* __restrict__ out) {
  double a = 0.0, b = 0.0;
  TB_GRID_STRIDE(i, n) {
    const float ww = w[i].x;
    a += (double)(cabs2(psi[i]) * ww);
    b += (double)(ww * ww);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, a);
    atomicAdd(out + 1, b);
  }
}

// y *= s or y /= s with s read from the device
__global__ void __launch_bounds__(kUpdThreads)
upd_scale_kernel(float2* __restrict__ y, long n, const float* __restrict__ s, int divide) {
  const float f = divide ? 1.0f / (*s) : *s;
  TB_GRID_STRIDE(i, n) {
    float2 v = y[i];
    v.x *= f;
    v.y *= f;
    y[i] = v;
  }
}

}  // namespace tb

extern "C" {

int tb_max_real(const void* x, int64_t n, float* out, tb_stream_t stream) {
  TB_REQUIRE(x && out && n >= 0, TB_ERR_INVALID, "tb_max_real: null pointer");
  if (n == 0) return TB_OK;
  tb::upd_max_real_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (const float2*)x, n, out);
  return tb::check_launch("tb_max_real");
}

int tb_rpie_update_psi_given_max(void* psi, const void* numerator, const void* precond,
                                 int64_t n, float alpha, const float* precond_max,
                                 tb_stream_t stream) {
  TB_REQUIRE(psi && numerator && precond && precond_max, TB_ERR_INVALID,
             "tb_rpie_update_psi_given_max: null pointer");
  if (n == 0) return TB_OK;
  tb::upd_rpie_psi_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)psi, (const float2*)numerator, (const float2*)precond, n, alpha, precond_max);
  return tb::check_launch("tb_rpie_update_psi_given_max");
}

int tb_rpie_update_psi_adam(void* psi, const void* numerator, const void* precond, float* v,
                            void* m, int64_t n, float alpha, double vdecay, double mdecay,
                            const float* precond_max, tb_stream_t stream) {
  TB_REQUIRE(psi && numerator && precond && v && m && precond_max, TB_ERR_INVALID,
             "tb_rpie_update_psi_adam: null pointer");
  TB_REQUIRE(vdecay < 1.0 && mdecay < 1.0, TB_ERR_INVALID,
             "tb_rpie_update_psi_adam: decays must be below 1");
  if (n == 0) return TB_OK;
  tb::upd_rpie_adam_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)psi, (const float2*)numerator, (const float2*)precond, v, (float2*)m, n, alpha,
      (float)vdecay, (float)(1.0 - vdecay), (float)mdecay, (float)(1.0 - mdecay), precond_max);
  return tb::check_launch("tb_rpie_update_psi_adam");
}

int tb_momentum_update(void* psi, const void* direction, void* m, int64_t n, double mdecay,
                       const float* beta, tb_stream_t stream) {
  TB_REQUIRE(psi && direction && m && beta, TB_ERR_INVALID, "tb_momentum_update: null pointer");
  if (n == 0) return TB_OK;
  tb::upd_momentum_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)psi, (const float2*)direction, (float2*)m, n, (float)mdecay,
      (float)(1.0 - mdecay), beta);
  return tb::check_launch("tb_momentum_update");
}

int tb_lstsq_precondition_object_given_max(void* out, const void* object_upd,
                                           const void* precond, int64_t n, float alpha,
                                           const float* precond_max, tb_stream_t stream) {
  TB_REQUIRE(out && object_upd && precond && precond_max, TB_ERR_INVALID,
             "tb_lstsq_precondition_object_given_max: null pointer");
  if (n == 0) return TB_OK;
  tb::upd_lstsq_precondition_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0,
                                      (cudaStream_t)stream>>>(
      (float2*)out, (const float2*)object_upd, (const float2*)precond, n, alpha, precond_max);
  return tb::check_launch("tb_lstsq_precondition_object_given_max");
}

int tb_add_quotient(void* y, const void* numerator, const void* precond, int64_t n,
                    int64_t precond_period, float eps, tb_stream_t stream) {
  TB_REQUIRE(y && numerator && precond, TB_ERR_INVALID, "tb_add_quotient: null pointer");
  TB_REQUIRE(precond_period >= 0, TB_ERR_INVALID, "tb_add_quotient: negative period");
  if (n == 0) return TB_OK;
  tb::upd_add_quotient_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)y, (const float2*)numerator, (const float2*)precond, n, precond_period, eps);
  return tb::check_launch("tb_add_quotient");
}

int tb_object_pointwise_constraints(void* psi, int64_t n, float positivity, int clip,
                                    float a_max, tb_stream_t stream) {
  TB_REQUIRE(psi, TB_ERR_INVALID, "tb_object_pointwise_constraints: null pointer");
  TB_REQUIRE(positivity >= 0.f && positivity <= 1.f, TB_ERR_INVALID,
             "Positivity constraint must be in the range [0, 1] not %g.", (double)positivity);
  if (n == 0) return TB_OK;
  tb::upd_object_pointwise_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0,
                                    (cudaStream_t)stream>>>((float2*)psi, n, positivity, clip,
                                                            a_max);
  return tb::check_launch("tb_object_pointwise_constraints");
}

int tb_object_smoothness(void* out, const void* psi, int nslices, int height, int width,
                         float a, tb_stream_t stream) {
  TB_REQUIRE(out && psi && out != psi, TB_ERR_INVALID,
             "tb_object_smoothness: needs distinct in / out arrays");
  TB_REQUIRE(a >= 0.f && a < 0.125f, TB_ERR_INVALID,
             "Smoothness constraint must be in range [0, 1/8) not %g.", (double)a);
  TB_REQUIRE(nslices > 0 && height > 0 && width > 0, TB_ERR_INVALID,
             "tb_object_smoothness: bad shape");
  const long n = (long)nslices * height * width;
  tb::upd_smooth_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)out, (const float2*)psi, nslices, height, width, a);
  return tb::check_launch("tb_object_smoothness");
}

int tb_weighted_norm_sums(const void* psi, const void* weight, int64_t n, double* out,
                          tb_stream_t stream) {
  TB_REQUIRE(psi && weight && out, TB_ERR_INVALID, "tb_weighted_norm_sums: null pointer");
  cudaMemsetAsync(out, 0, 2 * sizeof(double), (cudaStream_t)stream);
  if (n == 0) return TB_OK;
  tb::upd_weighted_norm_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (const float2*)psi, (const float2*)weight, n, out);
  return tb::check_launch("tb_weighted_norm_sums");
}

int tb_scale_by_device_scalar(void* y, int64_t n, const float* s, int divide,
                              tb_stream_t stream) {
  TB_REQUIRE(y && s, TB_ERR_INVALID, "tb_scale_by_device_scalar: null pointer");
  if (n == 0) return TB_OK;
  tb::upd_scale_kernel<<<tb::upd_grid(n), tb::kUpdThreads, 0, (cudaStream_t)stream>>>(
      (float2*)y, n, s, divide);
  return tb::check_launch("tb_scale_by_device_scalar");
}

}  // extern "C"
