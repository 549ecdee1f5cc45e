// Host-side helper for tike_b200.cluster.wobbly_center: the O(P^2) growth
// loop of the reference's heterogeneous clustering
// (src/tike/cluster.py:360-376) in C, bit-exact with the NumPy expressions it
// replaces:
//   centre = np.mean(population[labels == c], axis=0)   -> sequential float32
//            accumulation in index order, then / float32(count)
//   dist   = np.linalg.norm(population[free] - centre, axis=1)
//            -> sqrtf(dx*dx + dy*dy), each operation rounded to float32
//   far    = np.argmax(dist)                             -> first maximum
// (no FMA contraction: the products must be rounded before the add).
// NumPy needs ~6 ms per step at 100k positions (10 minutes per call); this
// loop takes seconds.  The seeding (argpartition) stays in NumPy.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/tike_b200.h"

namespace {

// d[k] = sqrt((x[k]-cx)^2 + (y[k]-cy)^2) + bias[k]  (bias: 0 alive, -inf dead;
// v + 0 is exact), returns the position of the FIRST maximum.  Every float
// operation is rounded separately (no FMA), so the AVX2 path gives the same
// bits as the scalar one and as NumPy.
__attribute__((optimize("fp-contract=off")))
size_t farthest_scalar(const float* x, const float* y, const float* bias, size_t n, float cx,
                       float cy, float* d) {
  float best = -std::numeric_limits<float>::infinity();
  for (size_t k = 0; k < n; ++k) {
    const float dx = x[k] - cx, dy = y[k] - cy;
    const float v = sqrtf(dx * dx + dy * dy) + bias[k];
    d[k] = v;
    best = v > best ? v : best;
  }
  size_t pos = 0;
  while (d[pos] != best) ++pos;
  return pos;
}

#if defined(__x86_64__)
__attribute__((target("avx2"), optimize("fp-contract=off")))
size_t farthest_avx2(const float* x, const float* y, const float* bias, size_t n, float cx,
                     float cy, float* d) {
  const __m256 vcx = _mm256_set1_ps(cx), vcy = _mm256_set1_ps(cy);
  __m256 vbest = _mm256_set1_ps(-std::numeric_limits<float>::infinity());
  size_t k = 0;
  for (; k + 8 <= n; k += 8) {
    const __m256 dx = _mm256_sub_ps(_mm256_loadu_ps(x + k), vcx);
    const __m256 dy = _mm256_sub_ps(_mm256_loadu_ps(y + k), vcy);
    const __m256 s = _mm256_add_ps(_mm256_mul_ps(dx, dx), _mm256_mul_ps(dy, dy));
    const __m256 v = _mm256_add_ps(_mm256_sqrt_ps(s), _mm256_loadu_ps(bias + k));
    _mm256_storeu_ps(d + k, v);
    vbest = _mm256_max_ps(vbest, v);
  }
  float lanes[8];
  _mm256_storeu_ps(lanes, vbest);
  float best = lanes[0];
  for (int j = 1; j < 8; ++j) best = lanes[j] > best ? lanes[j] : best;
  for (; k < n; ++k) {
    const float dx = x[k] - cx, dy = y[k] - cy;
    const float v = sqrtf(dx * dx + dy * dy) + bias[k];
    d[k] = v;
    best = v > best ? v : best;
  }
  const __m256 vb = _mm256_set1_ps(best);
  size_t pos = 0;
  for (; pos + 8 <= n; pos += 8) {
    const int m = _mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(d + pos), vb, _CMP_EQ_OQ));
    if (m) return pos + (size_t)__builtin_ctz((unsigned)m);
  }
  while (d[pos] != best) ++pos;
  return pos;
}
#endif

size_t farthest(const float* x, const float* y, const float* bias, size_t n, float cx, float cy,
                float* d) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return farthest_avx2(x, y, bias, n, cx, cy, d);
#endif
  return farthest_scalar(x, y, bias, n, cx, cy, d);
}

}  // namespace

extern "C" __attribute__((optimize("fp-contract=off")))
int tb_cluster_grow(const float* population, int64_t npoints, int ndim,
                    uint16_t* labels, int num_cluster, int64_t steps) {
  if (!population || !labels || ndim != 2 || num_cluster < 1 || npoints < 1) return TB_ERR_INVALID;
  const uint16_t UNASSIGNED = 0xFFFF;
  // Free points in index order, coordinates contiguous so the distance loop
  // vectorises (SSE/AVX float arithmetic rounds every operation to float32
  // exactly like NumPy).  Assigned points are tombstoned (bias = -inf) and the
  // arrays are compacted once a quarter of them is dead, instead of erasing
  // one element per step.
  std::vector<int64_t> free_idx;
  std::vector<float> fx, fy, bias, dist;
  // Members of every cluster in index order with the running float32 sums
  // np.mean accumulates in that order: prefix[k] = x_0 + ... + x_(k-1).
  // Inserting a member at position p only invalidates the sums after p.
  struct Cluster { std::vector<int64_t> idx; std::vector<float> sx, sy; };
  std::vector<Cluster> clusters((size_t)num_cluster);
  for (int64_t i = 0; i < npoints; ++i) {
    if (labels[i] == UNASSIGNED) {
      free_idx.push_back(i);
      fx.push_back(population[2 * i]);
      fy.push_back(population[2 * i + 1]);
    } else if (labels[i] < num_cluster) {
      clusters[labels[i]].idx.push_back(i);
    }
  }
  bias.assign(free_idx.size(), 0.0f);
  dist.resize(free_idx.size());
  for (auto& c : clusters) {
    c.sx.resize(c.idx.size() + 1);
    c.sy.resize(c.idx.size() + 1);
    float ax = 0.f, ay = 0.f;
    c.sx[0] = 0.f; c.sy[0] = 0.f;
    for (size_t k = 0; k < c.idx.size(); ++k) {
      ax = ax + population[2 * c.idx[k]];
      ay = ay + population[2 * c.idx[k] + 1];
      c.sx[k + 1] = ax; c.sy[k + 1] = ay;
    }
  }
  size_t nfree = free_idx.size(), ndead = 0;
  for (int64_t step = 0; step < steps; ++step) {
    if (nfree == ndead) break;
    Cluster& cl = clusters[(size_t)(step % num_cluster)];
    if (cl.idx.empty()) return TB_ERR_INVALID;  // np.mean of an empty set is NaN
    const float cnt = (float)cl.idx.size();
    const float cx = cl.sx.back() / cnt, cy = cl.sy.back() / cnt;
    const size_t pos = farthest(fx.data(), fy.data(), bias.data(), nfree, cx, cy, dist.data());
    const int64_t chosen = free_idx[pos];
    labels[chosen] = (uint16_t)(step % num_cluster);
    bias[pos] = -std::numeric_limits<float>::infinity();
    ++ndead;
    // insert into the sorted member list and redo the running sums after it
    const size_t at = (size_t)(std::upper_bound(cl.idx.begin(), cl.idx.end(), chosen) - cl.idx.begin());
    cl.idx.insert(cl.idx.begin() + (long)at, chosen);
    cl.sx.push_back(0.f);
    cl.sy.push_back(0.f);
    float ax = cl.sx[at], ay = cl.sy[at];
    for (size_t k = at; k < cl.idx.size(); ++k) {
      ax = ax + population[2 * cl.idx[k]];
      ay = ay + population[2 * cl.idx[k] + 1];
      cl.sx[k + 1] = ax; cl.sy[k + 1] = ay;
    }
    if (ndead * 4 > nfree && nfree > 1024) {  // compact, keeping index order
      size_t w = 0;
      for (size_t k = 0; k < nfree; ++k) {
        if (bias[k] != 0.0f) continue;
        free_idx[w] = free_idx[k]; fx[w] = fx[k]; fy[w] = fy[k]; ++w;
      }
      nfree = w; ndead = 0;
      free_idx.resize(w); fx.resize(w); fy.resize(w);
      bias.assign(w, 0.0f);
    }
  }
  return TB_OK;
}

// One sweep of the pairwise-swap refinement of cluster.compact
// (src/tike/cluster.py:587-626) in C, same float64 expressions and the same
// visiting order as the NumPy loop it replaces:
//   for p in order:                      # order = argsort(happiness) at sweep start
//     if happiness[p] < 0:
//       gain = ((dist[p, labels[p]] + dist[all, labels]) - dist[p, labels]) - dist[all, labels[p]]
//       o = first argmax of gain over {gain > 0 and labels != labels[p]}
//       swap labels[o], labels[p]; refresh happiness[o], happiness[p]
// dist is (n, num_cluster) float64 row-major.  Returns 1 if anything was
// swapped, 0 if not, negative on bad arguments.
extern "C" __attribute__((optimize("fp-contract=off")))
int tb_cluster_compact_sweep(const double* dist, uint16_t* labels, const int64_t* wanted,
                             double* happiness, const int64_t* order, int64_t n,
                             int num_cluster) {
  if (!dist || !labels || !wanted || !happiness || !order || n < 1 || num_cluster < 1)
    return TB_ERR_INVALID;
  int swapped = 0;
  const int64_t C = num_cluster;
  for (int64_t t = 0; t < n; ++t) {
    const int64_t p = order[t];
    if (!(happiness[p] < 0)) continue;
    const uint16_t lp = labels[p];
    const double dpp = dist[p * C + lp];
    double best = 0.0;  // only gains > 0 qualify
    int64_t o = -1;
    for (int64_t q = 0; q < n; ++q) {
      const uint16_t lq = labels[q];
      if (lq == lp) continue;
      const double gain = ((dpp + dist[q * C + lq]) - dist[p * C + lq]) - dist[q * C + lp];
      if (gain > best) { best = gain; o = q; }  // strict: keeps the first maximum
    }
    if (o >= 0) {
      swapped = 1;
      labels[p] = labels[o];
      labels[o] = lp;
      happiness[o] = dist[o * C + wanted[o]] - dist[o * C + labels[o]];
      happiness[p] = dist[p * C + wanted[p]] - dist[p * C + labels[p]];
    }
  }
  return swapped;
}


// Inlier test of the RANSAC affine fit (position.py:277-327), the per-iteration
// pass over all positions: residual of the candidate transform, squared error
// against the threshold.  Same float64 expressions, evaluated left to right
// and rounded one by one (no FMA), as the NumPy lines they replace in
// tike_b200/ptycho/position.py:
//   rx = x0 * m00 + y0 * m10 + t0 - x1;  ry = x0 * m01 + y0 * m11 + t1 - y1
//   inlier = rx * rx + ry * ry <= max_error^2
extern "C" __attribute__((optimize("fp-contract=off")))
int tb_affine_inliers(const double* x0, const double* y0, const double* x1, const double* y1,
                      int64_t n, const double* m /* m00 m01 m10 m11 */, double t0, double t1,
                      double max_error_sq, uint8_t* inlier, int64_t* count) {
  if (n < 0 || (n > 0 && (!x0 || !y0 || !x1 || !y1)) || !m || !count) return TB_ERR_INVALID;
  const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
  int64_t c = 0;
  for (int64_t k = 0; k < n; ++k) {
    const double rx = ((x0[k] * m00 + y0[k] * m10) + t0) - x1[k];
    const double ry = ((x0[k] * m01 + y0[k] * m11) + t1) - y1[k];
    const bool in = (rx * rx + ry * ry) <= max_error_sq;
    if (inlier) inlier[k] = in ? 1 : 0;
    c += in ? 1 : 0;
  }
  *count = c;
  return TB_OK;
}
