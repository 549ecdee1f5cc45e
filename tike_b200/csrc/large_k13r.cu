// K1 and K3 of the large-detector pipeline at ND = 256, register-resident
// variants for the plain case (probe width = detector width, shared probe, no
// eigen-weight / position sums); same contract as large_exit_cols_kernel<256> and
// large_cols_gradient_kernel<256> in large_fused.cu.
//
// A CTA owns a block of 16 columns of one position and loops over its modes.
// The 256-point column transform is two radix-16 stages with 16 values per
// thread; lanes always walk over the 16 columns (one 128-byte line per row), so
// both ownerships are conflict free without padding:
//   stage-1 ownership   rows g + 16 k, column c      (k = 0..15)
//   stage-2 ownership   rows 16 g + n, column c      (n = 0..15)
//   K1   probe x patch built in registers (the interpolated patch of the block is
//        computed once per position and stays in registers over the modes)
//        -> radix-16 -> twiddle -> tile -> radix-16 -> wave
//   K3   wave -> radix-16^-1 -> tile -> conj twiddle -> radix-16^-1 -> chi in
//        registers: conj(probe) chi accumulated over the modes in registers,
//        conj(patch) chi reduced into the probe-numerator replicas, chi_out
// The tile is written once and read once per transform (3 + 3 in the generic
// kernels) and the operand of the next mode (probe values in K1, wave in K3) is
// fetched into registers before the current one is transformed.
// Replaces (with K2): rpie.py:355-505, lstsq.py:422-543 at BASELINE config 3.
#include "solver_dev.cuh"
#include "dft32.cuh"

namespace tb {

namespace k13r {
constexpr int ND = 256, VC = 16, NT = 256, NCB = ND / VC, PS = VC + 1;
// K1: ND x VC tile + twiddles; K3: (ND + 1) x (VC + 1) scatter tile (the
// transform uses its first ND * VC entries), patch block, twiddles
constexpr size_t kSmem1 = (size_t)ND * VC * 8 + ND * 8;
constexpr size_t kSmem3 = (size_t)(ND + 1) * PS * 8 + (size_t)ND * VC * 8 + ND * 8;

// interpolated patch values of this thread's stage-1 ownership
__device__ __forceinline__ void load_patch(float2 (&o)[16], const float2* __restrict__ psi, int H,
                                           int W, const Corner& c, int g, int col) {
  const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + ND < H) & (c.ix + ND < W);
  if (interior) {
    const float2* __restrict__ o0 = psi + (long)(c.iy + g) * W + c.ix + col;
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
      float2 q[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2* r0 = o0 + (long)(16 * (k0 + j)) * W;
        q[j][0] = __ldg(r0); q[j][1] = __ldg(r0 + 1);
        q[j][2] = __ldg(r0 + W); q[j][3] = __ldg(r0 + W + 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 v;
        v.x = q[j][0].x * c.w00; v.y = q[j][0].y * c.w00;
        v.x += q[j][1].x * c.w01; v.y += q[j][1].y * c.w01;
        v.x += q[j][2].x * c.w10; v.y += q[j][2].y * c.w10;
        v.x += q[j][3].x * c.w11; v.y += q[j][3].y * c.w11;
        o[k0 + j] = v;
      }
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < 16; ++k) o[k] = patch_value(psi, H, W, c, g + 16 * k, col);
  }
}
// read-only load that keeps its place in the instruction stream
__device__ __forceinline__ float2 ld_nc(const float2* addr) {
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(addr));
  return v;
}
}  // namespace k13r

// ---- K1: exit wave + forward column transforms ------------------------------
__global__ void __launch_bounds__(k13r::NT, 2)
large_exit_cols_reg_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count) {
  using namespace k13r;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + ND * VC;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const int tid = threadIdx.x, c = tid & 15, g = tid >> 4;
  float2* const tA = tile + g * VC + c;        // + 16 k * VC: row g + 16 k
  float2* const tB = tile + 16 * g * VC + c;   // + n * VC: row 16 g + n
  const long total = count * NCB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const long i = t / NCB;
    const Corner cn = make_corner(b.scan, s0 + i);
    const int col = cb * VC + c;
    const float2* __restrict__ pm = probe + (long)g * ND + col;  // + 16 k * ND, + m * ND^2
    float2 nx[16];  // probe values of the next mode, fetched ahead
#pragma unroll
    for (int k = 0; k < 16; ++k) nx[k] = __ldg(pm + 16 * k * ND);
    float2 o[16];
    load_patch(o, psi, H, W, cn, g, col);
    for (int m = 0; m < M; ++m) {
      float2 x[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = cmul(nx[k], o[k]);
      if (m + 1 < M) {
#pragma unroll
        for (int k = 0; k < 16; ++k) nx[k] = __ldg(pm + (long)(m + 1) * ND * ND + 16 * k * ND);
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) nx[k] = make_float2(0.f, 0.f);
      }
      dft<16>(x);
      tA[0] = x[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) tA[16 * k * VC] = cmul(x[k], tw[g * k]);
      __syncthreads();
      float2 y[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) y[n] = tB[n * VC];
      dft<16>(y);  // rows end up in slot order: slot 16 g + p holds frequency g + 16 p
      float2* img = wave + (i * M + m) * (long)ND * ND + (long)(16 * g) * ND + col;
#pragma unroll
      for (int p = 0; p < 16; ++p) img[(long)p * ND] = y[p];
      __syncthreads();
    }
  }
}

// ---- K3: inverse column transforms + gradients --------------------------------
__global__ void __launch_bounds__(k13r::NT, 2)
large_cols_gradient_reg_kernel(RpieDev a, const float2* __restrict__ wave, long s0, long count) {
  using namespace k13r;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* O = tile + (ND + 1) * PS;  // patch of the block, [k][tid]
  float2* tw = O + ND * VC;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const float inv_m = a.divide_by_modes ? 1.0f / (float)M : 1.0f;
  const int tid = threadIdx.x, c = tid & 15, g = tid >> 4;
  float2* const tA = tile + g * VC + c;
  float2* const tB = tile + 16 * g * VC + c;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * M * ND * ND : nullptr;
  const long total = count * NCB;
  const long wave_off = (long)(16 * g) * ND + c;  // stage-2 ownership: + p * ND

  float2 nx[16];  // wave values of the next inverse transform, fetched ahead
  if ((long)blockIdx.x < total) {
    const long t = blockIdx.x;
    const float2* img = wave + (t / NCB) * M * (long)ND * ND + (t % NCB) * VC + wave_off;
#pragma unroll
    for (int p = 0; p < 16; ++p) nx[p] = __ldcs(img + (long)p * ND);
  }
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const long i = t / NCB;
    const long s = s0 + i;
    const Corner cn = make_corner(b.scan, s);
    const int col = cb * VC + c;
    if (replica) {  // the patch is only needed for the probe numerator
      float2 o[16];
      load_patch(o, psi, H, W, cn, g, col);
#pragma unroll
      for (int k = 0; k < 16; ++k) O[k * NT + tid] = o[k];
    }
    float2 acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int m = 0; m < M; ++m) {
      float2 y[16];
#pragma unroll
      for (int p = 0; p < 16; ++p) y[p] = nx[p];
      {
        const float2* nimg = nullptr;
        if (m + 1 < M) {
          nimg = wave + (i * M + m + 1) * (long)ND * ND + cb * VC + wave_off;
        } else if (t + gridDim.x < total) {
          const long tn = t + gridDim.x;
          nimg = wave + (tn / NCB) * M * (long)ND * ND + (tn % NCB) * VC + wave_off;
        }
        if (nimg) {
#pragma unroll
          for (int p = 0; p < 16; ++p) nx[p] = __ldcs(nimg + (long)p * ND);
        } else {
#pragma unroll
          for (int p = 0; p < 16; ++p) nx[p] = make_float2(0.f, 0.f);
        }
      }
      idft<16>(y);  // slot order in
#pragma unroll
      for (int n = 0; n < 16; ++n) tB[n * VC] = y[n];
      __syncthreads();
      const float2* __restrict__ pm = probe + (long)m * ND * ND + (long)g * ND + col;
      // probe values in two halves (this kernel is at the register limit: 32
      // values of accumulator, next wave and current transform each)
      float2 pv[8];
      if (a.accumulate_object) {
#pragma unroll
        for (int k = 0; k < 8; ++k) pv[k] = ld_nc(pm + 16 * k * ND);
      }
      float2 x[16];
      x[0] = tA[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) x[k] = cmulc(tw[g * k], tA[16 * k * VC]);
      idft<16>(x);  // chi at rows g + 16 k, column col
      if (a.chi_out) {
        float2* cout = a.chi_out + ((long)s * M + m) * ND * ND + (long)g * ND + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) __stcs(cout + 16 * k * ND, x[k]);
      }
      if (a.accumulate_object) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 gr = cmulc(pv[k], x[8 * hh + k]);
            acc[8 * hh + k].x += gr.x;
            acc[8 * hh + k].y += gr.y;
          }
          if (hh == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pv[k] = ld_nc(pm + 16 * (8 + k) * ND);
          }
        }
      }
      if (replica) {
        float2* rep = replica + (long)m * ND * ND + (long)g * ND + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) red_add_f32x2(rep + 16 * k * ND, cmulc(O[k * NT + tid], x[k]));
      }
      __syncthreads();
    }
    if (a.accumulate_object) {
      // the block's share of the gradient goes through the tile so that the
      // four bilinear taps of an object pixel leave as one reduction
      // (convolution.cu:57-64); taps owned by the neighbouring column block
      // arrive with that block's reductions
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int py = g + 16 * k;
        const int y = cn.iy + py, x = cn.ix + col;
        const bool ok = (y >= 0) & (y < H) & (x >= 0) & (x < W);
        tile[py * PS + c] = ok ? cscale(acc[k], inv_m) : make_float2(0.f, 0.f);
      }
      __syncthreads();
      for (int idx = tid; idx < (ND + 1) * (VC + 1); idx += NT) {
        const int ty = idx / (VC + 1), tx = idx - ty * (VC + 1);
        const int y = cn.iy + ty, x = cn.ix + cb * VC + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float2 r = make_float2(0.f, 0.f);
        const bool a0 = ty < ND, a1 = ty > 0, b0 = tx < VC, b1 = tx > 0;
        if (a0 & b0) { const float2 q = tile[ty * PS + tx];           r.x += cn.w00 * q.x; r.y += cn.w00 * q.y; }
        if (a0 & b1) { const float2 q = tile[ty * PS + tx - 1];       r.x += cn.w01 * q.x; r.y += cn.w01 * q.y; }
        if (a1 & b0) { const float2 q = tile[(ty - 1) * PS + tx];     r.x += cn.w10 * q.x; r.y += cn.w10 * q.y; }
        if (a1 & b1) { const float2 q = tile[(ty - 1) * PS + tx - 1]; r.x += cn.w11 * q.x; r.y += cn.w11 * q.y; }
        if (r.x != 0.f || r.y != 0.f) red_add_f32x2(a.psi_num + (long)y * W + x, r);
      }
      __syncthreads();
    }
  }
}

// ---- ND = 512, one mode (BASELINE config 5) ------------------------------------
// Column transform = a radix-16 stage (rows n2 + 32 k; two butterflies per
// thread, n2 = g and g + 16) and a radix-32 stage in registers (rows 32 g + n,
// dft32.cuh).  Row slot 32 g + p holds row frequency g + 16 * dft32_freq(p): K2
// (large_rows_modulus_reg512_kernel) is told so.  With a single mode the object
// gradient conj(probe) chi is written IN PLACE into the (pitch 17) tile by the
// thread that just read those entries, which is then the scatter tile; no
// accumulator exists.
namespace k13r512 {
constexpr int ND = 512, VC = 16, NT = 256, NCB = ND / VC, PS = VC + 1;
constexpr size_t kSmem1 = (size_t)ND * VC * 8 + ND * 8;
constexpr size_t kSmem3 = (size_t)(ND + 1) * PS * 8 + ND * 8;

// interpolated patch values at rows n2 + 32 k, k = 0..15, column col
__device__ __forceinline__ void load_patch(float2 (&o)[16], const float2* __restrict__ psi, int H,
                                           int W, const Corner& c, int n2, int col) {
  const bool interior = (c.iy >= 0) & (c.ix >= 0) & (c.iy + ND < H) & (c.ix + ND < W);
  if (interior) {
    const float2* __restrict__ o0 = psi + (long)(c.iy + n2) * W + c.ix + col;
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
      float2 q[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2* r0 = o0 + (long)(32 * (k0 + j)) * W;
        q[j][0] = __ldg(r0); q[j][1] = __ldg(r0 + 1);
        q[j][2] = __ldg(r0 + W); q[j][3] = __ldg(r0 + W + 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 v;
        v.x = q[j][0].x * c.w00; v.y = q[j][0].y * c.w00;
        v.x += q[j][1].x * c.w01; v.y += q[j][1].y * c.w01;
        v.x += q[j][2].x * c.w10; v.y += q[j][2].y * c.w10;
        v.x += q[j][3].x * c.w11; v.y += q[j][3].y * c.w11;
        o[k0 + j] = v;
      }
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < 16; ++k) o[k] = patch_value(psi, H, W, c, n2 + 32 * k, col);
  }
}
}  // namespace k13r512

__global__ void __launch_bounds__(k13r512::NT, 2)
large_exit_cols_reg512_kernel(RpieDev a, float2* __restrict__ wave, long s0, long count) {
  using namespace k13r512;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* tw = tile + ND * VC;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const int M = b.nmodes, H = b.height, W = b.width;
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const int tid = threadIdx.x, c = tid & 15, g = tid >> 4;
  const long total = count * NCB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const long i = t / NCB;
    const Corner cn = make_corner(b.scan, s0 + i);
    const int col = cb * VC + c;
    for (int m = 0; m < M; ++m) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n2 = g + 16 * q;
        const float2* __restrict__ pm = probe + (long)m * ND * ND + (long)n2 * ND + col;
        float2 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = __ldg(pm + 32 * k * ND);
        float2 o[16];
        load_patch(o, psi, H, W, cn, n2, col);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = cmul(x[k], o[k]);
        dft<16>(x);
        float2* tA = tile + n2 * VC + c;  // + 32 k * VC: row n2 + 32 k
        tA[0] = x[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) tA[32 * k * VC] = cmul(x[k], tw[n2 * k]);
      }
      __syncthreads();
      float2 y[32];
      const float2* tB = tile + 32 * g * VC + c;  // + n * VC: row 32 g + n
#pragma unroll
      for (int n = 0; n < 32; ++n) y[n] = tB[n * VC];
      dft32(y);
      float2* img = wave + (i * M + m) * (long)ND * ND + (long)(32 * g) * ND + col;
#pragma unroll
      for (int p = 0; p < 32; ++p) img[(long)p * ND] = y[p];
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(k13r512::NT, 2)
large_cols_gradient_reg512_kernel(RpieDev a, const float2* __restrict__ wave, long s0,
                                  long count) {
  using namespace k13r512;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);  // (ND + 1) x PS, transform and scatter
  float2* tw = tile + (ND + 1) * PS;
  fill_twiddles<ND>(tw);
  __syncthreads();
  const tb_batch& b = a.b;
  const int H = b.height, W = b.width;  // one mode
  const float2* __restrict__ psi = (const float2*)b.psi;
  const float2* __restrict__ probe = (const float2*)b.probe;
  const int tid = threadIdx.x, c = tid & 15, g = tid >> 4;
  float2* replica = a.probe_sums ? a.replicas + (long)(blockIdx.x % a.nrep) * ND * ND : nullptr;
  const long total = count * NCB;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int cb = (int)(t % NCB);
    const long i = t / NCB;
    const long s = s0 + i;
    const Corner cn = make_corner(b.scan, s);
    const int col = cb * VC + c;
    {
      const float2* img = wave + i * (long)ND * ND + (long)(32 * g) * ND + col;
      float2 y[32];
#pragma unroll
      for (int p = 0; p < 32; ++p) y[p] = __ldcs(img + (long)p * ND);
      idft32(y);  // slot order in, rows 32 g + n out
      float2* tB = tile + 32 * g * PS + c;
#pragma unroll
      for (int n = 0; n < 32; ++n) tB[n * PS] = y[n];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n2 = g + 16 * q;
      float2* tA = tile + n2 * PS + c;  // + 32 k * PS: row n2 + 32 k
      const float2* __restrict__ pm = probe + (long)n2 * ND + col;
      float2 pv[16];
      if (a.accumulate_object) {
#pragma unroll
        for (int k = 0; k < 16; ++k) pv[k] = __ldg(pm + 32 * k * ND);
      }
      float2 x[16];
      x[0] = tA[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) x[k] = cmulc(tw[n2 * k], tA[32 * k * PS]);
      idft<16>(x);  // chi at rows n2 + 32 k, column col
      if (a.chi_out) {
        float2* cout = a.chi_out + (long)s * ND * ND + (long)n2 * ND + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) __stcs(cout + 32 * k * ND, x[k]);
      }
      if (replica) {
        float2 o[16];
        load_patch(o, psi, H, W, cn, n2, col);
        float2* rep = replica + (long)n2 * ND + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) red_add_f32x2(rep + 32 * k * ND, cmulc(o[k], x[k]));
      }
      if (a.accumulate_object) {
        // in place: these are the entries this thread has just read
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int y = cn.iy + n2 + 32 * k, xx = cn.ix + col;
          const bool ok = (y >= 0) & (y < H) & (xx >= 0) & (xx < W);
          tA[32 * k * PS] = ok ? cmulc(pv[k], x[k]) : make_float2(0.f, 0.f);
        }
      }
    }
    __syncthreads();
    if (a.accumulate_object) {
      // four bilinear taps of an object pixel leave as one reduction
      // (convolution.cu:57-64); one mode, so divide_by_modes is a no-op
      for (int idx = tid; idx < (ND + 1) * (VC + 1); idx += NT) {
        const int ty = idx / (VC + 1), tx = idx - ty * (VC + 1);
        const int y = cn.iy + ty, x = cn.ix + cb * VC + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float2 r = make_float2(0.f, 0.f);
        const bool a0 = ty < ND, a1 = ty > 0, b0 = tx < VC, b1 = tx > 0;
        if (a0 & b0) { const float2 q = tile[ty * PS + tx];           r.x += cn.w00 * q.x; r.y += cn.w00 * q.y; }
        if (a0 & b1) { const float2 q = tile[ty * PS + tx - 1];       r.x += cn.w01 * q.x; r.y += cn.w01 * q.y; }
        if (a1 & b0) { const float2 q = tile[(ty - 1) * PS + tx];     r.x += cn.w10 * q.x; r.y += cn.w10 * q.y; }
        if (a1 & b1) { const float2 q = tile[(ty - 1) * PS + tx - 1]; r.x += cn.w11 * q.x; r.y += cn.w11 * q.y; }
        if (r.x != 0.f || r.y != 0.f) red_add_f32x2(a.psi_num + (long)y * W + x, r);
      }
      __syncthreads();
    }
  }
}

bool k13_reg_applies(const RpieDev& a) {
  static const bool on = [] {
    const char* e = getenv("TB_LARGE_K13R");  // 0: keep the generic K1 / K3 (A/B timing)
    return e ? atoi(e) != 0 : true;
  }();
  const tb_batch& b = a.b;
  const bool plain = b.probe_width == b.detector_width && b.eigen_weights == nullptr &&
                     !b.probe_per_position && a.eig_step == nullptr && a.pos_num == nullptr;
  if (!on || !plain) return false;
  if (b.detector_width == 256) return true;
  // 512: one mode, and K2 must be the kernel that knows the 16 x 32 row slots
  return b.detector_width == 512 && b.nmodes == 1 && k2_reg_applies(a);
}

static int k13_configure(const char* who) {
  static bool configured = false;
  if (configured) return TB_OK;
  cudaError_t e = cudaFuncSetAttribute(large_exit_cols_reg_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)k13r::kSmem1);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(large_cols_gradient_reg_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k13r::kSmem3);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(large_exit_cols_reg512_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k13r512::kSmem1);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(large_cols_gradient_reg512_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k13r512::kSmem3);
  if (e != cudaSuccess) return set_error((int)e, "%s: kernel attributes: %s", who, cudaGetErrorString(e));
  configured = true;
  return TB_OK;
}

int launch_k1_reg(const RpieDev& a, float2* wave, long s0, long count, int sms, cudaStream_t st,
                  const char* who) {
  int rc = k13_configure(who);
  if (rc != TB_OK) return rc;
  if (a.b.detector_width == 512) {
    const long tasks = count * k13r512::NCB;
    const long g = tasks < (long)sms * 2 ? tasks : (long)sms * 2;
    large_exit_cols_reg512_kernel<<<(unsigned)g, k13r512::NT, k13r512::kSmem1, st>>>(a, wave, s0,
                                                                                  count);
    return check_launch(who);
  }
  const long tasks = count * k13r::NCB;
  const long g = tasks < (long)sms * 2 ? tasks : (long)sms * 2;
  large_exit_cols_reg_kernel<<<(unsigned)g, k13r::NT, k13r::kSmem1, st>>>(a, wave, s0, count);
  return check_launch(who);
}

int launch_k3_reg(const RpieDev& a, const float2* wave, long s0, long count, int sms,
                  cudaStream_t st, const char* who) {
  int rc = k13_configure(who);
  if (rc != TB_OK) return rc;
  if (a.b.detector_width == 512) {
    const long tasks = count * k13r512::NCB;
    const long g = tasks < (long)sms * 2 ? tasks : (long)sms * 2;
    large_cols_gradient_reg512_kernel<<<(unsigned)g, k13r512::NT, k13r512::kSmem3, st>>>(
        a, wave, s0, count);
    return check_launch(who);
  }
  const long tasks = count * k13r::NCB;
  const long g = tasks < (long)sms * 2 ? tasks : (long)sms * 2;
  large_cols_gradient_reg_kernel<<<(unsigned)g, k13r::NT, k13r::kSmem3, st>>>(a, wave, s0, count);
  return check_launch(who);
}

}  // namespace tb
