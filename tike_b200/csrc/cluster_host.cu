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
#include <vector>

#include "../../include/tike_b200.h"

extern "C" __attribute__((optimize("fp-contract=off")))
int tb_cluster_grow(const float* population, int64_t npoints, int ndim,
                    uint16_t* labels, int num_cluster, int64_t steps) {
  if (!population || !labels || ndim != 2 || num_cluster < 1 || npoints < 1) return TB_ERR_INVALID;
  const uint16_t UNASSIGNED = 0xFFFF;
  // free points, compacted in index order (coordinates kept contiguous so the
  // distance loop vectorises; SSE float arithmetic rounds every operation to
  // float32 exactly like NumPy)
  std::vector<int64_t> free_idx;
  std::vector<float> fx, fy, dist;
  // members of every cluster in index order (np.mean sums them in that order)
  std::vector<std::vector<int64_t>> members((size_t)num_cluster);
  for (int64_t i = 0; i < npoints; ++i) {
    if (labels[i] == UNASSIGNED) {
      free_idx.push_back(i);
      fx.push_back(population[2 * i]);
      fy.push_back(population[2 * i + 1]);
    } else if (labels[i] < num_cluster) {
      members[labels[i]].push_back(i);
    }
  }
  dist.resize(free_idx.size());
  for (int64_t step = 0; step < steps; ++step) {
    const size_t nfree = free_idx.size();
    if (nfree == 0) break;
    const int c = (int)(step % num_cluster);
    float sx = 0.f, sy = 0.f;
    const std::vector<int64_t>& mem = members[(size_t)c];
    if (mem.empty()) return TB_ERR_INVALID;  // np.mean of an empty set is NaN
    for (size_t k = 0; k < mem.size(); ++k) {
      sx = sx + population[2 * mem[k]];
      sy = sy + population[2 * mem[k] + 1];
    }
    const float cx = sx / (float)mem.size(), cy = sy / (float)mem.size();
    const float* px = fx.data();
    const float* py = fy.data();
    float* d = dist.data();
    float best = -1.0f;
    for (size_t k = 0; k < nfree; ++k) {
      const float dx = px[k] - cx, dy = py[k] - cy;
      const float v = sqrtf(dx * dx + dy * dy);
      d[k] = v;
      best = v > best ? v : best;
    }
    size_t pos = 0;
    while (d[pos] != best) ++pos;  // first maximum, like np.argmax
    const int64_t chosen = free_idx[pos];
    labels[chosen] = (uint16_t)c;
    // keep the member list sorted by index
    std::vector<int64_t>& m2 = members[(size_t)c];
    m2.insert(std::upper_bound(m2.begin(), m2.end(), chosen), chosen);
    free_idx.erase(free_idx.begin() + (long)pos);
    fx.erase(fx.begin() + (long)pos);
    fy.erase(fy.begin() + (long)pos);
  }
  return TB_OK;
}
