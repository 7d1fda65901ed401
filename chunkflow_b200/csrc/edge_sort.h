// Host helper of the region graph (segmentation.cu: cfb_region_graph_read; also compiled into tests/host_emulation): the
// permutation that sorts edge keys (smaller id << 32 | larger id).  The device gathers the table in arbitrary order; a
// comparison sort of millions of keys through an index array took 1 s of the 1.15 s the region graph of 7.5 M edges cost, so:
// counting sort on the smaller id (ids are dense, 1..N), then each node's few edges by their larger id.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

inline void sorted_edge_order(const unsigned long long* keys, size_t n, std::vector<uint32_t>& order) {
  order.resize(n);
  if (!n) return;
  uint32_t max_u = 0;
  for (size_t i = 0; i < n; ++i) max_u = std::max(max_u, (uint32_t)(keys[i] >> 32));
  std::vector<uint32_t> start((size_t)max_u + 2, 0);
  for (size_t i = 0; i < n; ++i) ++start[(size_t)(keys[i] >> 32) + 1];
  for (size_t u = 1; u < start.size(); ++u) start[u] += start[u - 1];
  std::vector<uint32_t> cursor(start.begin(), start.end() - 1);
  for (size_t i = 0; i < n; ++i) order[cursor[(size_t)(keys[i] >> 32)]++] = (uint32_t)i;
  for (size_t u = 0; u + 1 < start.size(); ++u)
    if (start[u + 1] - start[u] > 1)
      std::sort(order.begin() + start[u], order.begin() + start[u + 1], [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
}
