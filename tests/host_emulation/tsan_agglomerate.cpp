// TEST INFRASTRUCTURE: ThreadSanitizer run of the threaded rounds of csrc/agglomerate_host.cu (tests/test_segmentation_agglomerate.py
// builds this file together with that source, -fsanitize=thread): random graphs, one worker against eight workers on small
// slices, the results must be identical and the sanitizer silent.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>
#include <string>
extern "C" int cfb_agglomerate_edges_host(int64_t, int64_t, const uint32_t*, const uint32_t*, const uint64_t*, const uint32_t*, float, uint32_t*);
namespace cfb { void set_last_error(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); } }
int main() {
  std::mt19937_64 rng(5);
  int bad = 0;
  for (int trial = 0; trial < 24; ++trial) {
    const int n = 50 + (int)(rng() % 3000);
    std::set<std::pair<uint32_t, uint32_t>> pairs;
    const int m = n * (2 + (int)(rng() % 4));
    for (int i = 0; i < m; ++i) {
      uint32_t a = 1 + (uint32_t)(rng() % (n - 1)), b = 1 + (uint32_t)(rng() % (n - 1));
      if (a != b) pairs.insert({std::min(a, b), std::max(a, b)});
    }
    std::vector<uint32_t> u, v, c; std::vector<uint64_t> s;
    const int levels = (trial % 3 == 0) ? 4 : 100000;
    for (auto& p : pairs) { u.push_back(p.first); v.push_back(p.second); uint32_t cc = 1 + (uint32_t)(rng() % 3); c.push_back(cc);
      s.push_back((uint64_t)((double)(rng() % (levels + 1)) / levels * cc * 1073741824.0)); }
    std::vector<uint32_t> r1(n), r2(n);
    const float thr = 0.2f + 0.6f * (float)(rng() % 100) / 100.f;
    setenv("CFB_AGGLOMERATE_MODE", "2", 1); setenv("CFB_AGGLOMERATE_THREADS", "1", 1); setenv("CFB_AGGLOMERATE_GRAIN", "2048", 1);
    cfb_agglomerate_edges_host(n, (int64_t)u.size(), u.data(), v.data(), s.data(), c.data(), thr, r1.data());
    setenv("CFB_AGGLOMERATE_THREADS", "8", 1); setenv("CFB_AGGLOMERATE_GRAIN", trial % 2 ? "16" : "300", 1);
    setenv("CFB_AGGLOMERATE_MODE", trial % 2 ? "2" : "1", 1);
    cfb_agglomerate_edges_host(n, (int64_t)u.size(), u.data(), v.data(), s.data(), c.data(), thr, r2.data());
    if (r1 != r2) { ++bad; printf("MISMATCH trial %d\n", trial); }
  }
  printf("bad %d\n", bad);
  return bad != 0;
}
