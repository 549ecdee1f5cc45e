// Object and probe preconditioners, recomputed once per epoch over all
// positions of a worker (solvers/_preconditioner.py:48-167):
//
//   psi_precond   (H, W) = scatter_s( sum_m |P_m|^2 )     (Patch.adj of a constant plane)
//   probe_precond (N, N) = sum_s |patch_s(psi)|^2          (Patch.fwd, then |.|^2)
//
// Two implementations:
//
// * window kernels (probe width <= 128): the caller passes the positions in a
//   band-sorted order (rows in bands of kBand pixels, ascending column inside a
//   band).  One persistent CTA per SM walks a contiguous run of that order and
//   keeps a (kBand + N + 1) x RW window of the object (probe sum), or of the
//   output (object sum), in shared memory.  The window is a ring in the column
//   direction: when a position needs columns beyond it, only the new columns
//   are loaded (or the retiring ones flushed).  Consecutive positions overlap
//   by ~90 %, so the object is read from L2 about once instead of once per
//   position, and the object sum issues one global reduction per window pixel
//   instead of one per footprint pixel per position.
// * direct kernels (any width, any order): one footprint at a time straight
//   from / to global memory; used for N > 128 and as the path for positions
//   whose footprint leaves the object.
#include <climits>

#include "solver_dev.cuh"

namespace tb {

constexpr int kBand = 16;               // band height of the sorted order (rows)
constexpr int kWinThreads = 1024;
constexpr int kWinRows = 16;            // patch rows per thread, probe sum
constexpr int kFootRows = 20;           // footprint rows per thread, object sum
constexpr size_t kWinSmem = 200 * 1024;  // shared memory budget of a window kernel

// ---- A = sum_m |P_m|^2 ---------------------------------------------------
__global__ void __launch_bounds__(256)
probe_amp_kernel(const float2* __restrict__ probe, int M, long n2, float* __restrict__ A) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2;
       i += (long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int m = 0; m < M; ++m) a += cabs2(__ldg(probe + (long)m * n2 + i));
    A[i] = a;
  }
}

// ---- direct kernels --------------------------------------------------------
// One footprint pixel per thread and one scalar reduction per footprint pixel.
__device__ __forceinline__ void scatter_amp_direct(const float* __restrict__ A, int N,
                                                   const Corner& c, float2* __restrict__ out,
                                                   int H, int W) {
  const int T = N + 1;
  for (int t = threadIdx.x; t < T * T; t += blockDim.x) {
    const int ty = t / T, tx = t - ty * T;
    const int y = c.iy + ty, x = c.ix + tx;
    if (y < 0 || y >= H || x < 0 || x >= W) continue;
    float v = 0.f;
    // a patch pixel whose leading image pixel is outside contributes nothing
    const bool a0 = ty < N, a1 = ty > 0 && (y - 1) >= 0, b0 = tx < N, b1 = tx > 0 && (x - 1) >= 0;
    if (a0 & b0) v += c.w00 * __ldg(A + ty * N + tx);
    if (a0 & b1) v += c.w01 * __ldg(A + ty * N + tx - 1);
    if (a1 & b0) v += c.w10 * __ldg(A + (ty - 1) * N + tx);
    if (a1 & b1) v += c.w11 * __ldg(A + (ty - 1) * N + tx - 1);
    red_add_f32(reinterpret_cast<float*>(out + (long)y * W + x), v);
  }
}

__global__ void __launch_bounds__(256)
precond_psi_kernel(const float* __restrict__ A, int N, const float* __restrict__ scan,
                   const int* __restrict__ order, long npos, float2* __restrict__ out,
                   int H, int W) {
  for (long i = blockIdx.x; i < npos; i += gridDim.x) {
    const Corner c = make_corner(scan, order ? order[i] : i);
    scatter_amp_direct(A, N, c, out, H, W);
  }
}

// probe sum: accumulated in registers per CTA, one reduction per pixel per CTA
constexpr int PP_K = 16;  // pixels per thread -> 4096 pixels per blockIdx.y
__global__ void __launch_bounds__(256)
precond_probe_kernel(const float2* __restrict__ psi, int H, int W,
                     const float* __restrict__ scan, const int* __restrict__ order,
                     long npos, int N, float2* __restrict__ out) {
  float acc[PP_K];
  const int base = blockIdx.y * 256 * PP_K;
#pragma unroll
  for (int k = 0; k < PP_K; ++k) acc[k] = 0.f;
  // a contiguous run of the (sorted) order per CTA keeps its reads in L2
  const long per = (npos + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = (lo + per < npos) ? lo + per : npos;
  for (long i = lo; i < hi; ++i) {
    const Corner c = make_corner(scan, order ? order[i] : i);
#pragma unroll
    for (int k = 0; k < PP_K; ++k) {
      const int idx = base + threadIdx.x + k * 256;
      if (idx < N * N) {
        const int py = idx / N, px = idx - py * N;
        const int y = c.iy + py, x = c.ix + px;
        if (y >= 0 && y < H && x >= 0 && x < W)
          acc[k] += cabs2(patch_value(psi, H, W, c, py, px));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < PP_K; ++k) {
    const int idx = base + threadIdx.x + k * 256;
    if (idx < N * N) red_add_f32(reinterpret_cast<float*>(out + idx), acc[k]);
  }
}

// window width for an element size, or 0 when the window kernel does not apply
__host__ __device__ constexpr int window_width(int N, size_t elem, size_t fixed_bytes,
                                                int slack_min) {
  if (N > 128 || N < 1 || fixed_bytes >= kWinSmem) return 0;
  long rw = (long)((kWinSmem - fixed_bytes) / ((size_t)(kBand + N + 1) * elem));
  if (rw > 4 * (N + 1)) rw = 4 * (N + 1);
  return rw >= N + 1 + slack_min ? (int)rw : 0;
}
__host__ __device__ constexpr size_t amp_plane_bytes(int N) {
  return (size_t)(N + 2) * (N + 2) * 4;
}
__host__ __device__ constexpr int probe_window_width(int N) { return window_width(N, 8, 0, 8); }
__host__ __device__ constexpr int psi_window_width(int N) {
  return window_width(N, 4, amp_plane_bytes(N), 8);
}

// ---- window kernels -----------------------------------------------------------
// Window geometry shared by both: rows [wy0, wy0 + RH) with wy0 a multiple of
// kBand and RH = kBand + N + 1 (every footprint whose top row lies in the band
// fits), columns [wx0, wx0 + RW) stored at ring slot (column mod RW).
// The positions of a CTA's run, read two steps ahead (order) and one step
// ahead (scan): the dependent loads order[i] -> scan[order[i]] would otherwise
// sit exposed at the start of every position.
struct CornerStream {
  const float2* scan;
  const int* order;
  long hi;
  long next_index;   // order[i + 1]
  float2 next_scan;  // scan[order[i]]
  __device__ __forceinline__ long index_at(long i) const {
    return i < hi ? (order ? (long)__ldg(order + i) : i) : 0;
  }
  __device__ __forceinline__ void start(const float* scan_, const int* order_, long lo, long hi_) {
    scan = reinterpret_cast<const float2*>(scan_);
    order = order_;
    hi = hi_;
    next_scan = make_float2(0.f, 0.f);
    if (lo < hi) next_scan = __ldg(scan + index_at(lo));
    next_index = index_at(lo + 1);
  }
  // corner of position i; issues the loads for i + 1 and i + 2
  __device__ __forceinline__ Corner take(long i) {
    const float sy = next_scan.x, sx = next_scan.y;
    if (i + 1 < hi) next_scan = __ldg(scan + next_index);
    next_index = index_at(i + 2);
    const float fy0 = floorf(sy), fx0 = floorf(sx);
    const float fy = sy - fy0, fx = sx - fx0;
    Corner c;
    c.iy = (int)fy0;
    c.ix = (int)fx0;
    c.w00 = (1.0f - fx) * (1.0f - fy);
    c.w01 = fx * (1.0f - fy);
    c.w10 = (1.0f - fx) * fy;
    c.w11 = fx * fy;
    return c;
  }
};

__device__ __forceinline__ bool footprint_inside(const Corner& c, int N, int H, int W) {
  return (c.iy >= 0) & (c.ix >= 0) & (c.iy + N < H) & (c.ix + N < W);
}

// probe sum.  Thread = one patch column x a vertical run of RPT rows, so the
// lower pair of taps of one pixel is the upper pair of the next.
// CN = compile-time probe width (0: run-time width): with it the row loops
// unroll without guards and the window offsets fold into the instructions.
template <int CN>
__global__ void __launch_bounds__(kWinThreads, 1)
precond_probe_win_kernel(const float2* __restrict__ psi, int H, int W,
                         const float* __restrict__ scan, const int* __restrict__ order,
                         long npos, int N_, int RH_, int RW_, float2* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* win = reinterpret_cast<float2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = CN ? CN : N_;
  const int RH = CN ? kBand + CN + 1 : RH_;
  const int RW = CN ? probe_window_width(CN ? CN : 1) : RW_;
  const int G = kWinThreads / N;             // row groups
  const int RPT = (N + G - 1) / G;           // rows per thread (<= kWinRows)
  const int col = tid % N, g = tid / N;
  const int r0 = g * RPT;
  const bool active = (g < G) & (r0 < N);
  float acc[kWinRows];
#pragma unroll
  for (int k = 0; k < kWinRows; ++k) acc[k] = 0.f;

  const long per = (npos + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = (lo + per < npos) ? lo + per : npos;
  int wy0 = INT_MIN, wx0 = 0;  // window anchor; identical in every thread

  // columns [c0, c1) of all window rows -> ring slots
  auto load_cols = [&](int c0, int c1) {
    for (int r = warp; r < RH; r += kWinThreads / 32) {
      const int y = wy0 + r;
      for (int x = c0 + lane; x < c1; x += 32) {
        float2 v = make_float2(0.f, 0.f);
        if (y < H && x < W) v = __ldg(psi + (long)y * W + x);
        win[r * RW + x % RW] = v;
      }
    }
  };

  CornerStream corners;
  corners.start(scan, order, lo, hi);
  for (long i = lo; i < hi; ++i) {
    const Corner c = corners.take(i);
    if (!footprint_inside(c, N, H, W)) {
      // footprint leaves the object: straight from global memory
      if (active) {
#pragma unroll
        for (int k = 0; k < kWinRows; ++k) {
          const int py = r0 + k;
          if (k < RPT && py < N) {
            const int y = c.iy + py, x = c.ix + col;
            if (y >= 0 && y < H && x >= 0 && x < W)
              acc[k] += cabs2(patch_value(psi, H, W, c, py, col));
          }
        }
      }
      continue;
    }
    const int band = (c.iy / kBand) * kBand;
    if (band != wy0 || c.ix < wx0 || c.ix >= wx0 + RW) {
      __syncthreads();  // readers of the old window are done
      wy0 = band;
      wx0 = c.ix;
      load_cols(wx0, wx0 + RW);
      __syncthreads();
    } else if (c.ix + N >= wx0 + RW) {
      __syncthreads();
      load_cols(wx0 + RW, c.ix + RW);  // only the columns that are new
      wx0 = c.ix;
      __syncthreads();
    }
    if (active) {
      const int x = c.ix + col;
      const int s0 = x % RW, s1 = (s0 + 1 == RW) ? 0 : s0 + 1;
      const float2* row = win + (c.iy - wy0 + r0) * RW;
      float2 a0 = row[s0], a1 = row[s1];
      auto one_row = [&](int k) {
        row += RW;
        const float2 b0 = row[s0], b1 = row[s1];
        float2 r;
        r.x = a0.x * c.w00; r.y = a0.y * c.w00;
        r.x += a1.x * c.w01; r.y += a1.y * c.w01;
        r.x += b0.x * c.w10; r.y += b0.y * c.w10;
        r.x += b1.x * c.w11; r.y += b1.y * c.w11;
        acc[k] += cabs2(r);
        a0 = b0; a1 = b1;
      };
      if (CN != 0 && r0 + RPT <= N) {
#pragma unroll
        for (int k = 0; k < kWinRows; ++k)
          if (k < RPT) one_row(k);  // RPT is a constant here
      } else {
#pragma unroll
        for (int k = 0; k < kWinRows; ++k)
          if (k < RPT && r0 + k < N) one_row(k);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < kWinRows; ++k) {
      const int py = r0 + k;
      if (k < RPT && py < N) red_add_f32(reinterpret_cast<float*>(out + py * N + col), acc[k]);
    }
  }
}

// object sum.  The window accumulates in shared memory (plain read-modify-
// write, one barrier per position) and reaches global memory once per pixel
// when its columns retire.  Ap is A with a border of
// zeros: Ap[ty + 1][tx + 1] = A[ty][tx], so the four taps need no guards.
template <int CN>
__global__ void __launch_bounds__(kWinThreads, 1)
precond_psi_win_kernel(const float* __restrict__ A, int N_, const float* __restrict__ scan,
                       const int* __restrict__ order, long npos, int RH_, int RW_,
                       float2* __restrict__ out, int H, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = CN ? CN : N_;
  const int RH = CN ? kBand + CN + 1 : RH_;
  const int RW = CN ? psi_window_width(CN ? CN : 1) : RW_;
  const int T = N + 1, AP = N + 2;
  float* Ap = reinterpret_cast<float*>(smem_raw);
  float* win = Ap + AP * AP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < AP * AP; t += kWinThreads) {
    const int r = t / AP - 1, q = t % AP - 1;
    Ap[t] = (r >= 0 && r < N && q >= 0 && q < N) ? __ldg(A + r * N + q) : 0.f;
  }
  for (int t = tid; t < RH * RW; t += kWinThreads) win[t] = 0.f;
  const int G = kWinThreads / T;
  const int RPT = (T + G - 1) / G;  // footprint rows per thread (<= kFootRows)
  const int col = tid % T, g = tid / T;
  const int r0 = g * RPT;
  const bool active = (g < G) & (r0 < T);

  const long per = (npos + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = (lo + per < npos) ? lo + per : npos;
  int wy0 = INT_MIN, wx0 = 0;

  // columns [c0, c1) of all window rows: add to the output, clear the slots
  auto flush_cols = [&](int c0, int c1) {
    for (int r = warp; r < RH; r += kWinThreads / 32) {
      const int y = wy0 + r;
      for (int x = c0 + lane; x < c1; x += 32) {
        float* slot = win + r * RW + x % RW;
        const float v = *slot;
        if (v != 0.f) {
          *slot = 0.f;
          if (y < H && x < W) red_add_f32(reinterpret_cast<float*>(out + (long)y * W + x), v);
        }
      }
    }
  };
  __syncthreads();

  CornerStream corners;
  corners.start(scan, order, lo, hi);
  for (long i = lo; i < hi; ++i) {
    const Corner c = corners.take(i);
    if (!footprint_inside(c, N, H, W)) {
      scatter_amp_direct(A, N, c, out, H, W);
      continue;
    }
    const int band = (c.iy / kBand) * kBand;
    // (the barrier that ends every position also covers the flushes)
    if (band != wy0 || c.ix < wx0 || c.ix >= wx0 + RW) {
      if (wy0 != INT_MIN) flush_cols(wx0, wx0 + RW);
      wy0 = band;
      wx0 = c.ix;
      __syncthreads();
    } else if (c.ix + N >= wx0 + RW) {
      flush_cols(wx0, c.ix);  // the columns that retire
      wx0 = c.ix;
      __syncthreads();
    }
    if (active) {
      const int x = c.ix + col;
      const int s0 = x % RW;
      float* wrow = win + (c.iy - wy0 + r0) * RW + s0;
      const float* arow = Ap + r0 * AP + col;  // Ap[ty][tx], Ap[ty][tx + 1]
      float u0 = arow[0], u1 = arow[1];        // taps of patch row ty - 1
      auto one_row = [&](int) {
        arow += AP;
        const float l0 = arow[0], l1 = arow[1];  // taps of patch row ty
        float v = c.w00 * l1;
        v += c.w01 * l0;
        v += c.w10 * u1;
        v += c.w11 * u0;
        *wrow += v;  // this thread is the only writer of the pixel for this position
        wrow += RW;
        u0 = l0; u1 = l1;
      };
      if (CN != 0 && r0 + RPT <= T) {
#pragma unroll
        for (int k = 0; k < kFootRows; ++k)
          if (k < RPT) one_row(k);  // RPT is a constant here
      } else {
#pragma unroll
        for (int k = 0; k < kFootRows; ++k)
          if (k < RPT && r0 + k < T) one_row(k);
      }
    }
    // the next position's footprint overlaps this one with another thread
    // assignment (a shared-memory float reduction would be a CAS loop)
    __syncthreads();
  }
  if (wy0 != INT_MIN) flush_cols(wx0, wx0 + RW);
}

static int window_grid(long npos) {
  int sms = 148;
  tb_sm_count(&sms);
  // at least 8 consecutive positions per CTA, else the window is not reused
  long g = (npos + 7) / 8;
  if (g > sms) g = sms;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace tb

extern "C" {

int tb_precond_psi(const void* probe, int nmodes, int probe_width, const float* scan,
                   const int32_t* order, int64_t npos, void* psi_precond, int height,
                   int width, float* scratch, tb_stream_t stream) {
  TB_REQUIRE(probe && (scan || npos == 0) && psi_precond && scratch, TB_ERR_INVALID,
             "tb_precond_psi: null pointer");
  TB_REQUIRE(probe_width > 0 && nmodes > 0, TB_ERR_INVALID, "tb_precond_psi: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(psi_precond, 0, (size_t)height * width * 8, st);
  if (e != cudaSuccess) return tb::set_error((int)e, "tb_precond_psi: %s", cudaGetErrorString(e));
  if (npos == 0) return TB_OK;
  const int N = probe_width;
  const long n2 = (long)N * N;
  tb::probe_amp_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(
      (const float2*)probe, nmodes, n2, scratch);
  const size_t ap_bytes = tb::amp_plane_bytes(N);
  const int RW = order ? tb::psi_window_width(N) : 0;
  if (RW > 0 && tb::kWinThreads / (N + 1) >= 1 &&
      (N + 1 + tb::kWinThreads / (N + 1) - 1) / (tb::kWinThreads / (N + 1)) <= tb::kFootRows) {
    const int RH = tb::kBand + N + 1;
    const size_t smem = ap_bytes + (size_t)RH * RW * 4;
    auto kernel = N == 128 ? tb::precond_psi_win_kernel<128>
                  : N == 64 ? tb::precond_psi_win_kernel<64> : tb::precond_psi_win_kernel<0>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tb::set_error((int)e, "tb_precond_psi: %s", cudaGetErrorString(e));
    kernel<<<tb::window_grid(npos), tb::kWinThreads, smem, st>>>(
        scratch, N, scan, order, npos, RH, RW, (float2*)psi_precond, height, width);
    return tb::check_launch("tb_precond_psi(window)");
  }
  int sms = 148;
  tb_sm_count(&sms);
  long grid = (long)sms * 8;
  if (npos < grid) grid = npos;
  tb::precond_psi_kernel<<<(unsigned)grid, 256, 0, st>>>(
      scratch, N, scan, order, npos, (float2*)psi_precond, height, width);
  return tb::check_launch("tb_precond_psi");
}

int tb_precond_probe(const void* psi, int height, int width, const float* scan,
                     const int32_t* order, int64_t npos, int probe_width,
                     void* probe_precond, tb_stream_t stream) {
  TB_REQUIRE(psi && (scan || npos == 0) && probe_precond, TB_ERR_INVALID,
             "tb_precond_probe: null pointer");
  TB_REQUIRE(probe_width > 0, TB_ERR_INVALID, "tb_precond_probe: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = probe_width;
  const long n2 = (long)N * N;
  cudaError_t e = cudaMemsetAsync(probe_precond, 0, (size_t)n2 * 8, st);
  if (e != cudaSuccess) return tb::set_error((int)e, "tb_precond_probe: %s", cudaGetErrorString(e));
  if (npos == 0) return TB_OK;
  const int RW = order ? tb::probe_window_width(N) : 0;
  if (RW > 0) {
    const int G = tb::kWinThreads / N;
    if ((N + G - 1) / G <= tb::kWinRows) {
      const int RH = tb::kBand + N + 1;
      const size_t smem = (size_t)RH * RW * 8;
      auto kernel = N == 128 ? tb::precond_probe_win_kernel<128>
                    : N == 64 ? tb::precond_probe_win_kernel<64>
                              : tb::precond_probe_win_kernel<0>;
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess)
        return tb::set_error((int)e, "tb_precond_probe: %s", cudaGetErrorString(e));
      kernel<<<tb::window_grid(npos), tb::kWinThreads, smem, st>>>(
          (const float2*)psi, height, width, scan, order, npos, N, RH, RW,
          (float2*)probe_precond);
      return tb::check_launch("tb_precond_probe(window)");
    }
  }
  int sms = 148;
  tb_sm_count(&sms);
  const unsigned gy = (unsigned)((n2 + 256 * tb::PP_K - 1) / (256 * tb::PP_K));
  long gx = ((long)sms * 8 + gy - 1) / gy;
  if (npos < gx) gx = npos;
  dim3 grid((unsigned)gx, gy);
  tb::precond_probe_kernel<<<grid, 256, 0, st>>>((const float2*)psi, height, width, scan, order,
                                                npos, N, (float2*)probe_precond);
  return tb::check_launch("tb_precond_probe");
}

}  // extern "C"
