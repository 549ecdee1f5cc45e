// Patch operator: bilinear gather of object patches and its scatter-add
// adjoint.  Replaces the reference's NVRTC kernels fwd_patch / adj_patch
// (src/tike/operators/cupy/convolution.cu:35-165, launched from
// patch.py:79-188).  Differences by design: one CTA covers a whole patch
// (not one row), neighbours outside the image are predicated off instead of
// being dereferenced with zero weight, and the adjoint issues 64-bit vector
// reductions (re,im together) instead of two scalar atomics.
#include "../../include/tike_b200.h"
#include "common.cuh"

namespace tb {

__global__ void __launch_bounds__(256)
patch_fwd_kernel(const float2* __restrict__ images, float2* __restrict__ patches,
                 const float* __restrict__ scan, int H, int W, int nscan,
                 int nrepeat, int pw, int padded) {
  const int ti = blockIdx.y;
  const int pad = (padded - pw) / 2;
  const float2* img = images + (long)ti * H * W;
  const long image_offset = (long)padded * padded * nrepeat * nscan * ti;
  for (int ts = blockIdx.x; ts < nscan; ts += gridDim.x) {
    const Corner c = make_corner(scan, (long)ti * nscan + ts);
    for (int idx = threadIdx.x; idx < pw * pw; idx += blockDim.x) {
      const int py = idx / pw, px = idx - py * pw;
      const int y = c.iy + py, x = c.ix + px;
      if (y < 0 || y >= H || x < 0 || x >= W) continue;  // convolution.cu:110,118
      const float2 v = patch_value(img, H, W, c, py, px);
      const long pi = image_offset + (long)(pad + py) * padded + pad + px;
      for (int r = 0; r < nrepeat; ++r) {
        patches[pi + (long)padded * padded * ((long)ts * nrepeat + r)] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
patch_adj_kernel(float2* __restrict__ images, const float2* __restrict__ patches,
                 const float* __restrict__ scan, int H, int W, int nscan,
                 int nrepeat, int pw, int padded, int npatch) {
  const int ti = blockIdx.y;
  const int pad = (padded - pw) / 2;
  float2* img = images + (long)ti * H * W;
  const long image_offset = (long)padded * padded * npatch * ti;
  for (int ts = blockIdx.x; ts < nscan; ts += gridDim.x) {
    const Corner c = make_corner(scan, (long)ti * nscan + ts);
    const long first = ((long)nrepeat * ts) % npatch;  // convolution.cu:136-137
    for (int idx = threadIdx.x; idx < pw * pw; idx += blockDim.x) {
      const int py = idx / pw, px = idx - py * pw;
      const int y = c.iy + py, x = c.ix + px;
      if (y < 0 || y >= H || x < 0 || x >= W) continue;
      float2 acc = make_float2(0.f, 0.f);
      const long pi = image_offset + (long)(pad + py) * padded + pad + px;
      for (int r = 0; r < nrepeat; ++r) {
        const float2 v = patches[pi + (long)padded * padded * (first + r)];
        acc.x += v.x;
        acc.y += v.y;
      }
      float2* p = img + (long)y * W + x;
      const bool y1 = (y + 1 < H), x1 = (x + 1 < W);
      red_add_f32x2(p, cscale(acc, c.w00));
      if (x1) red_add_f32x2(p + 1, cscale(acc, c.w01));
      if (y1) red_add_f32x2(p + W, cscale(acc, c.w10));
      if (y1 && x1) red_add_f32x2(p + W + 1, cscale(acc, c.w11));
    }
  }
}

}  // namespace tb

extern "C" {

int tb_patch_fwd(const void* images, void* patches, const float* positions,
                 int nimage, int height, int width, int nscan, int nrepeat,
                 int patch_width, int padded_width, tb_stream_t stream) {
  TB_REQUIRE(images && patches && positions, TB_ERR_INVALID,
             "tb_patch_fwd: null pointer");
  TB_REQUIRE(nimage > 0 && height > 0 && width > 0 && nrepeat > 0 &&
                 patch_width > 0 && padded_width >= patch_width && nscan >= 0,
             TB_ERR_INVALID, "tb_patch_fwd: bad shape");
  if (nscan == 0) return TB_OK;
  dim3 grid((unsigned)(nscan < 65535 * 16 ? nscan : 65535 * 16), (unsigned)nimage);
  tb::patch_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const float2*)images, (float2*)patches, positions, height, width, nscan,
      nrepeat, patch_width, padded_width);
  return tb::check_launch("tb_patch_fwd");
}

int tb_patch_adj(void* images, const void* patches, const float* positions,
                 int nimage, int height, int width, int nscan, int nrepeat,
                 int patch_width, int padded_width, int npatch,
                 tb_stream_t stream) {
  TB_REQUIRE(images && patches && positions, TB_ERR_INVALID,
             "tb_patch_adj: null pointer");
  TB_REQUIRE(nimage > 0 && height > 0 && width > 0 && nrepeat > 0 &&
                 patch_width > 0 && padded_width >= patch_width && nscan >= 0,
             TB_ERR_INVALID, "tb_patch_adj: bad shape");
  TB_REQUIRE(npatch >= nrepeat && ((long)nscan * nrepeat) % npatch == 0,
             TB_ERR_INVALID,
             "tb_patch_adj: (nscan * nrepeat) %% npatch != 0 or npatch < nrepeat");
  if (nscan == 0) return TB_OK;
  dim3 grid((unsigned)(nscan < 65535 * 16 ? nscan : 65535 * 16), (unsigned)nimage);
  tb::patch_adj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (float2*)images, (const float2*)patches, positions, height, width, nscan,
      nrepeat, patch_width, padded_width, npatch);
  return tb::check_launch("tb_patch_adj");
}

}  // extern "C"
