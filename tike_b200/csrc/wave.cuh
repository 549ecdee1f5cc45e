// Device helpers shared by the fused forward / gradient kernels: building the
// zero-padded exit wave  psi_m = probe_m(s) * patch(s)  directly in shared
// memory (reference: operators/cupy/convolution.py:58-101 and
// ptycho/probe.py:272-303 for the per-position varying probe).
#pragma once

#include "common.cuh"
#include "fft.cuh"

namespace tb {

struct ProbeSet {
  const float2* probe;    // (M, N, N) shared probe, or (B, M, N, N) if per_position
  const float2* eigen;    // (E, Me, N, N) or nullptr
  const float* weights;   // (B, E + 1, M) or nullptr
  int M, N, E, Me;
  int per_position;       // probe has a leading position axis
};

// probe value of mode m at (py, px) for scan position s
__device__ __forceinline__ float2 probe_value(const ProbeSet& ps, long s, int m,
                                              int py, int px) {
  const long off = ((long)m * ps.N + py) * ps.N + px;
  const long pos_off = ps.per_position ? s * (long)ps.M * ps.N * ps.N : 0;
  float2 v = __ldg(ps.probe + pos_off + off);
  if (ps.weights != nullptr) {
    const float* w = ps.weights + s * (long)(ps.E + 1) * ps.M;
    v = cscale(v, __ldg(w + m));
    if (ps.eigen != nullptr && m < ps.Me) {
      for (int c = 0; c < ps.E; ++c) {
        const float wc = __ldg(w + (c + 1) * ps.M + m);
        const float2 e = __ldg(ps.eigen + (((long)c * ps.Me + m) * ps.N + py) * ps.N + px);
        v.x += wc * e.x;
        v.y += wc * e.y;
      }
    }
  }
  return v;
}

// Fill the ND x ND (pitch ND+1) tile with the zero-padded exit wave of mode m.
template <int ND>
__device__ __forceinline__ void build_exitwave(float2* tile, const float2* psi,
                                               int H, int W, const Corner& c,
                                               const ProbeSet& ps, long s,
                                               int m, int pad) {
  const int N = ps.N;
  for (int idx = threadIdx.x; idx < ND * ND; idx += blockDim.x) {
    const int ly = idx / ND, lx = idx - ly * ND;
    const int py = ly - pad, px = lx - pad;
    float2 v = make_float2(0.f, 0.f);
    if (py >= 0 && py < N && px >= 0 && px < N) {
      v = cmul(probe_value(ps, s, m, py, px), patch_value(psi, H, W, c, py, px));
    }
    tile[ly * (ND + 1) + lx] = v;
  }
}

template <int ND>
__device__ __forceinline__ void fill_perm(unsigned short* l2f, unsigned short* f2l) {
  for (int i = threadIdx.x; i < ND; i += blockDim.x) {
    l2f[i] = (unsigned short)loc2freq<ND>(i);
    f2l[i] = (unsigned short)freq2loc<ND>(i);
  }
}

}  // namespace tb
