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
//
// Data structures (the loop is memory-latency bound -- about a hundred cache misses per merge; one host core, 2.4 M fragments /
// 11.2 M edges of a noisy 64x512x512 map, threshold 0.5, 1.8 M merges: 42 s with std::unordered_map and one std::vector per
// node, 19 s like this, of which 1.4 s set-up):
//   * edges: ONE open-addressing table, key = (smaller id << 32 | larger id), linear probing, tombstones; pooled edges
//     never outnumber the initial ones, so the table is sized once and only rebuilt when tombstones pile up;
//   * adjacency: singly linked lists in one pool (head per node); merged lists are walked lazily, entries whose edge is gone
//     are skipped;
//   * heap: only edges with score < threshold ever enter it -- the loop stops at the first score >= threshold, so entries
//     at or above it can never be popped before the end, and a pooled edge is pushed when ITS score is below.
#include <algorithm>
#include <cstdint>
#include <queue>
#include <stdexcept>
#include <vector>

#include "chunkflow_b200.h"
#include "common.cuh"

namespace {

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

// key 0 = never used (ids >= 1 would make key 0 impossible anyway: node 0 may not appear with itself), ~0 = tombstone
class EdgeTable {
 public:
  struct Slot {
    uint64_t key, sum;
    uint32_t count;
  };
  static constexpr uint64_t kEmpty = 0, kDead = ~0ULL;

  explicit EdgeTable(size_t expected) {
    size_t cap = 16;
    while (cap < expected * 2 + 2) cap <<= 1;
    slots_.assign(cap, Slot{kEmpty, 0, 0});
    mask_ = cap - 1;
  }
  Slot* find(uint64_t key) {
    for (size_t h = hash(key) & mask_;; h = (h + 1) & mask_) {
      Slot& s = slots_[h];
      if (s.key == key) return &s;
      if (s.key == kEmpty) return nullptr;
    }
  }
  // the key must not be present
  void insert(uint64_t key, uint64_t sum, uint32_t count) {
    if ((used_ + 1) * 10 > slots_.size() * 7) rebuild();
    for (size_t h = hash(key) & mask_;; h = (h + 1) & mask_) {
      Slot& s = slots_[h];
      if (s.key == kEmpty || s.key == kDead) {
        if (s.key == kEmpty) ++used_;
        s = Slot{key, sum, count};
        ++live_;
        return;
      }
    }
  }
  void erase(Slot* s) {
    s->key = kDead;
    --live_;
  }

 private:
  static size_t hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (size_t)k;
  }
  void rebuild() {  // drop the tombstones (same capacity unless the live entries alone fill half of it)
    std::vector<Slot> old;
    old.swap(slots_);
    size_t cap = old.size();
    while (cap < live_ * 2 + 2) cap <<= 1;
    slots_.assign(cap, Slot{kEmpty, 0, 0});
    mask_ = cap - 1;
    used_ = live_ = 0;
    for (const Slot& s : old)
      if (s.key != kEmpty && s.key != kDead) {
        for (size_t h = hash(s.key) & mask_;; h = (h + 1) & mask_)
          if (slots_[h].key == kEmpty) { slots_[h] = s; break; }
        ++used_; ++live_;
      }
  }
  std::vector<Slot> slots_;
  size_t mask_ = 0, used_ = 0 /* slots that are not kEmpty */, live_ = 0;
};

struct Link {
  uint32_t node, next;
};
constexpr uint32_t kNil = 0xFFFFFFFFu;

}  // namespace

extern "C" int cfb_agglomerate_edges_host(int64_t num_nodes, int64_t num_edges, const uint32_t* u, const uint32_t* v,
                                          const uint64_t* sum_fixed, const uint32_t* count, float threshold, uint32_t* root_of) {
  try {
    if (num_nodes < 1 || num_nodes > (int64_t)UINT32_MAX || num_edges < 0 || num_edges >= (int64_t)1 << 31 || !root_of)
      throw std::invalid_argument("agglomerate: bad sizes");
    if (num_edges && (!u || !v || !sum_fixed || !count)) throw std::invalid_argument("agglomerate: null edge arrays");
    const double thr = (double)threshold;
    EdgeTable edges((size_t)num_edges);
    std::vector<uint32_t> head((size_t)num_nodes, kNil);
    std::vector<Link> pool;
    pool.reserve((size_t)num_edges * 2 + 16);
    auto link = [&](uint32_t from, uint32_t to) {
      pool.push_back(Link{to, head[from]});
      head[from] = (uint32_t)(pool.size() - 1);
    };
    std::vector<Entry> initial;
    for (int64_t i = 0; i < num_edges; ++i) {
      if (u[i] >= num_nodes || v[i] >= num_nodes || u[i] == v[i] || !count[i])
        throw std::invalid_argument("agglomerate: edge with an id outside [0, num_nodes), a self loop or a zero count");
      const uint32_t a = u[i] < v[i] ? u[i] : v[i], b = u[i] < v[i] ? v[i] : u[i];
      const uint64_t key = edge_key(a, b);
      if (edges.find(key)) throw std::invalid_argument("agglomerate: duplicate edge");
      edges.insert(key, sum_fixed[i], count[i]);
      link(a, b);
      link(b, a);
      const double sc = edge_score(sum_fixed[i], count[i]);
      if (sc < thr) initial.push_back(Entry{sc, a, b, sum_fixed[i], count[i]});
    }
    std::priority_queue<Entry, std::vector<Entry>, Later> heap(Later(), std::move(initial));
    std::vector<uint8_t> alive((size_t)num_nodes, 1);
    std::vector<uint32_t> parent((size_t)num_nodes);
    for (int64_t i = 0; i < num_nodes; ++i) parent[i] = (uint32_t)i;
    while (!heap.empty()) {
      const Entry e = heap.top();
      heap.pop();
      if (!alive[e.a] || !alive[e.b]) continue;
      EdgeTable::Slot* it = edges.find(edge_key(e.a, e.b));
      if (!it || it->sum != e.sum || it->count != e.count) continue;  // superseded entry
      // (every entry in the heap has score < threshold: nothing to test here; the loop ends when the heap runs dry)
      const uint32_t a = e.a, b = e.b;  // a < b: b is merged into a
      alive[b] = 0;
      parent[b] = a;
      edges.erase(it);
      for (uint32_t l = head[b]; l != kNil;) {
        const Link lk = pool[l];   // (pool may grow below: copy, do not hold a reference)
        l = lk.next;
        const uint32_t n = lk.node;
        if (n == a || !alive[n]) continue;
        EdgeTable::Slot* eb = edges.find(edge_key(b, n));
        if (!eb) continue;  // an entry left behind by an earlier merge
        uint64_t sum = eb->sum;
        uint32_t cnt = eb->count;
        edges.erase(eb);
        if (EdgeTable::Slot* ea = edges.find(edge_key(a, n))) {
          sum += ea->sum;
          cnt += ea->count;
          ea->sum = sum;
          ea->count = cnt;
        } else {
          edges.insert(edge_key(a, n), sum, cnt);
          link(a, n);
          link(n, a);
        }
        const double sc = edge_score(sum, cnt);
        if (sc < thr) heap.push(Entry{sc, a < n ? a : n, a < n ? n : a, sum, cnt});
      }
      head[b] = kNil;
    }
    for (int64_t i = 0; i < num_nodes; ++i) {
      uint32_t r = (uint32_t)i;
      while (parent[r] != r) r = parent[r];
      for (uint32_t c = (uint32_t)i; parent[c] != r;) { const uint32_t nx = parent[c]; parent[c] = r; c = nx; }  // compress
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
