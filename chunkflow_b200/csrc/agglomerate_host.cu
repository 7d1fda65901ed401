// Hierarchical agglomeration of a region graph on the HOST: the merge loop of waterz's
// IterativeRegionMerging with the scoring function the reference's plugin uses, OneMinus<MeanAffinity<...>>
// (reference plugins/agglomerate.py:13,38-41 -> waterz backend/IterativeRegionMerging.hpp; waterz runs this loop on the CPU in
// C++ too: it is a priority-queue walk over a graph of fragments, thousands to millions of nodes, not voxel work).  The
// voxel passes either side of it (watershed, region-graph statistics, relabel) are CUDA kernels (segmentation.cu).
//
// Rule set (restated in oracle/agglomeration_oracle.py: agglomerate_edges, the two are compared edge list by edge list):
//   score(edge) = 1 - sum / (count * 2^30)                     (double arithmetic; sums are 2^-30 fixed point)
//   repeat: take the edge with the smallest (score, smaller id, larger id); stop when score >= threshold;
//           merge the larger id into the smaller one; edges of both to a common neighbour pool sum and count.
#include <cstdint>
#include <queue>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "chunkflow_b200.h"
#include "common.cuh"

namespace {

struct Stat {
  uint64_t sum;
  uint32_t count;
};

struct Entry {
  double score;
  uint32_t a, b;
  uint64_t sum;
  uint32_t count;
};

struct Later {  // std::priority_queue keeps the LARGEST element on top: order by "comes later"
  bool operator()(const Entry& x, const Entry& y) const {
    if (x.score != y.score) return x.score > y.score;
    if (x.a != y.a) return x.a > y.a;
    return x.b > y.b;
  }
};

inline uint64_t edge_key(uint32_t p, uint32_t q) { return p < q ? ((uint64_t)p << 32) | q : ((uint64_t)q << 32) | p; }

inline double edge_score(uint64_t sum, uint32_t count) { return 1.0 - (double)sum / ((double)count * 1073741824.0); }

}  // namespace

extern "C" int cfb_agglomerate_edges_host(int64_t num_nodes, int64_t num_edges, const uint32_t* u, const uint32_t* v,
                                          const uint64_t* sum_fixed, const uint32_t* count, float threshold, uint32_t* root_of) {
  try {
    if (num_nodes < 1 || num_nodes > (int64_t)UINT32_MAX || num_edges < 0 || !root_of) throw std::invalid_argument("agglomerate: bad sizes");
    if (num_edges && (!u || !v || !sum_fixed || !count)) throw std::invalid_argument("agglomerate: null edge arrays");
    std::unordered_map<uint64_t, Stat> edges;
    edges.reserve((size_t)num_edges * 2);
    std::vector<std::vector<uint32_t>> nbrs((size_t)num_nodes);
    std::vector<Entry> initial;
    initial.reserve((size_t)num_edges);
    for (int64_t i = 0; i < num_edges; ++i) {
      if (u[i] >= num_nodes || v[i] >= num_nodes || u[i] == v[i] || !count[i])
        throw std::invalid_argument("agglomerate: edge with an id outside [0, num_nodes), a self loop or a zero count");
      const uint32_t a = u[i] < v[i] ? u[i] : v[i], b = u[i] < v[i] ? v[i] : u[i];
      if (!edges.emplace(edge_key(a, b), Stat{sum_fixed[i], count[i]}).second) throw std::invalid_argument("agglomerate: duplicate edge");
      nbrs[a].push_back(b);
      nbrs[b].push_back(a);
      initial.push_back(Entry{edge_score(sum_fixed[i], count[i]), a, b, sum_fixed[i], count[i]});
    }
    std::priority_queue<Entry, std::vector<Entry>, Later> heap(Later(), std::move(initial));
    std::vector<uint8_t> alive((size_t)num_nodes, 1);
    std::vector<uint32_t> parent((size_t)num_nodes);
    for (int64_t i = 0; i < num_nodes; ++i) parent[i] = (uint32_t)i;
    const double thr = (double)threshold;
    while (!heap.empty()) {
      const Entry e = heap.top();
      heap.pop();
      if (!alive[e.a] || !alive[e.b]) continue;
      const auto it = edges.find(edge_key(e.a, e.b));
      if (it == edges.end() || it->second.sum != e.sum || it->second.count != e.count) continue;  // superseded entry
      if (!(e.score < thr)) break;
      const uint32_t a = e.a, b = e.b;  // a < b: b is merged into a
      alive[b] = 0;
      parent[b] = a;
      edges.erase(it);
      for (const uint32_t n : nbrs[b]) {
        if (n == a || !alive[n]) continue;
        const auto eb = edges.find(edge_key(b, n));
        if (eb == edges.end()) continue;  // an entry left behind by an earlier merge
        Stat st = eb->second;
        edges.erase(eb);
        const auto ea = edges.find(edge_key(a, n));
        if (ea != edges.end()) {
          st.sum += ea->second.sum;
          st.count += ea->second.count;
          ea->second = st;
        } else {
          edges.emplace(edge_key(a, n), st);
          nbrs[a].push_back(n);
          nbrs[n].push_back(a);
        }
        heap.push(Entry{edge_score(st.sum, st.count), a < n ? a : n, a < n ? n : a, st.sum, st.count});
      }
      std::vector<uint32_t>().swap(nbrs[b]);
    }
    for (int64_t i = 0; i < num_nodes; ++i) {
      uint32_t r = (uint32_t)i;
      while (parent[r] != r) r = parent[r];
      root_of[i] = r;
    }
    return CFB_OK;
  } catch (const std::invalid_argument& ex) {
    cfb::set_last_error(ex.what());
    return CFB_ERR_INVALID_ARGUMENT;
  } catch (const std::exception& ex) {
    cfb::set_last_error(ex.what());
    return CFB_ERR_UNSUPPORTED;
  }
}
