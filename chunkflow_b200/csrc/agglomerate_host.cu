// Hierarchical agglomeration of a region graph on the HOST: the merge loop of waterz's
// IterativeRegionMerging with the scoring function the reference's plugin uses, OneMinus<MeanAffinity<...>>
// (reference plugins/agglomerate.py:13,38-41 -> waterz backend/IterativeRegionMerging.hpp; waterz runs this loop on the CPU in
// C++ too: it is a priority-queue walk over a graph of fragments, thousands to millions of nodes, not voxel work).  The
// voxel passes either side of it (watershed, region-graph statistics, relabel) are CUDA kernels (segmentation.cu).
//
// Rule set (restated in oracle/agglomeration_oracle.py: agglomerate_edges, the two are compared edge list by edge list):
//   score(edge) = 1 - sum / (count * 2^30)                     (double arithmetic; sums are 2^-30 fixed point)
//   anchor(edge) = the smallest (u, v) pair of ORIGINAL fragments among the faces pooled into the edge
//   repeat: take the edge with the smallest (score, anchor); stop when score >= threshold; merge its two clusters -- the
//           merged cluster is known by the smaller of the two ids; edges of both to a common neighbour pool sum and count
//           and keep the smaller anchor.
// Neither the order nor the result depends on which cluster's edges are MOVED, so the loop moves the shorter adjacency list
// (every edge is moved O(log n) times instead of once per merge of its cluster).
//
// Data structures (the loop is memory-latency bound; one host core, 2.4 M fragments / 11.2 M edges of a noisy 64x512x512 map,
// threshold 0.5, 1.8 M merges: 42 s with std::unordered_map, one std::vector per node and one binary heap; see
// profiles/r02c_host_merge_loop.txt for the steps down from there):
//   * edges: ONE open-addressing table, key = (smaller id << 32 | larger id), linear probing, tombstones; pooled edges
//     never outnumber the initial ones, so the table is sized once and only rebuilt when tombstones pile up;
//   * clusters: a structural root (the node whose lists are alive) and a label (the smallest fragment id in it); the edge
//     table and the queue use structural ids, the caller sees labels;
//   * adjacency: the initial neighbours of a node are one contiguous run (CSR) -- when the node is merged away, the table
//     slots of all its edges (and of the edges they pool with) are prefetched before the first is touched --; neighbours a
//     node gains through merges go to a linked overflow list; entries whose edge is gone are skipped lazily; the cluster
//     with FEWER list entries is the one merged away;
//   * priority queue: only edges with score < threshold ever enter it (the loop stops at the first score >= threshold, so
//     entries at or above it can never be popped).  The mean of pooled edges lies between the means pooled, i.e. a new
//     entry never scores below the edge being merged: the queue is MONOTONE, so 65 536 score buckets replace the global heap
//     -- append-only vectors, and a small heap (a few hundred entries, cache resident) for the bucket being drained, which
//     keeps the exact (score, anchor) order inside it.
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
  uint64_t anchor;  // smallest original (u << 32 | v) pooled into the edge: the tie rule
  uint64_t sum;
  uint32_t a, b;    // structural roots of the two clusters when the entry was pushed
  uint32_t count;
};

struct Later {  // std::priority_queue keeps the LARGEST element on top: order by "comes later"
  bool operator()(const Entry& x, const Entry& y) const {
    if (x.score != y.score) return x.score > y.score;
    return x.anchor > y.anchor;
  }
};

inline uint64_t edge_key(uint32_t p, uint32_t q) { return p < q ? ((uint64_t)p << 32) | q : ((uint64_t)q << 32) | p; }

inline double edge_score(uint64_t sum, uint32_t count) { return 1.0 - (double)sum / ((double)count * 1073741824.0); }

// key 0 = never used (ids >= 1 would make key 0 impossible anyway: node 0 may not appear with itself), ~0 = tombstone
class EdgeTable {
 public:
  struct Slot {
    uint64_t key, sum, anchor;
    uint32_t count;
  };
  static constexpr uint64_t kEmpty = 0, kDead = ~0ULL;

  explicit EdgeTable(size_t expected) {
    size_t cap = 16;
    while (cap < expected * 2 + 2) cap <<= 1;
    slots_.assign(cap, Slot{kEmpty, 0, 0, 0});
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
  void insert(uint64_t key, uint64_t sum, uint64_t anchor, uint32_t count) {
    if ((used_ + 1) * 10 > slots_.size() * 7) rebuild();
    for (size_t h = hash(key) & mask_;; h = (h + 1) & mask_) {
      Slot& s = slots_[h];
      if (s.key == kEmpty || s.key == kDead) {
        if (s.key == kEmpty) ++used_;
        s = Slot{key, sum, anchor, count};
        ++live_;
        return;
      }
    }
  }
  void erase(Slot* s) {
    s->key = kDead;
    --live_;
  }
  void prefetch(uint64_t key) const { __builtin_prefetch(&slots_[hash(key) & mask_]); }

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
    slots_.assign(cap, Slot{kEmpty, 0, 0, 0});
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

// Monotone priority queue over scores in [lo, hi): see the header comment.  pop() returns entries in exactly the order one
// global heap ordered by (score, a, b) would, provided no entry is pushed with a key below the last one popped's bucket --
// and even then it is only popped next (it joins the bucket being drained), which is what the global heap would do too.
class BucketQueue {
 public:
  BucketQueue(double lo, double hi, std::vector<Entry>&& initial) : lo_(lo), scale_(hi > lo ? (double)kBuckets / (hi - lo) : 0.0) {
    buckets_.resize(kBuckets);
    std::vector<uint32_t> n(kBuckets, 0);
    for (const Entry& e : initial) ++n[index(e.score)];
    for (int i = 0; i < kBuckets; ++i) buckets_[i].reserve(n[i]);
    for (const Entry& e : initial) buckets_[index(e.score)].push_back(e);
    std::vector<Entry>().swap(initial);
  }
  void push(const Entry& e) {
    const int i = index(e.score);
    if (i <= cur_) heap_.push(e); else buckets_[i].push_back(e);
  }
  bool pop(Entry& e) {
    while (heap_.empty()) {
      if (++cur_ >= kBuckets) return false;
      if (!buckets_[cur_].empty()) {
        heap_ = std::priority_queue<Entry, std::vector<Entry>, Later>(Later(), std::move(buckets_[cur_]));
        std::vector<Entry>().swap(buckets_[cur_]);
      }
    }
    e = heap_.top();
    heap_.pop();
    return true;
  }

 private:
  static constexpr int kBuckets = 1 << 16;
  int index(double score) const {
    const double x = (score - lo_) * scale_;
    if (!(x > 0.0)) return 0;
    return x < (double)(kBuckets - 1) ? (int)x : kBuckets - 1;
  }
  double lo_, scale_;
  int cur_ = -1;
  std::vector<std::vector<Entry>> buckets_;
  std::priority_queue<Entry, std::vector<Entry>, Later> heap_;
};

}  // namespace

extern "C" int cfb_agglomerate_edges_host(int64_t num_nodes, int64_t num_edges, const uint32_t* u, const uint32_t* v,
                                          const uint64_t* sum_fixed, const uint32_t* count, float threshold, uint32_t* root_of) {
  try {
    if (num_nodes < 1 || num_nodes > (int64_t)UINT32_MAX || num_edges < 0 || num_edges >= (int64_t)1 << 31 || !root_of)
      throw std::invalid_argument("agglomerate: bad sizes");
    if (num_edges && (!u || !v || !sum_fixed || !count)) throw std::invalid_argument("agglomerate: null edge arrays");
    const double thr = (double)threshold;
    EdgeTable edges((size_t)num_edges);
    // initial adjacency as CSR; neighbours gained later in an overflow list per node
    std::vector<uint32_t> adj_start((size_t)num_nodes + 1, 0), adj((size_t)num_edges * 2);
    for (int64_t i = 0; i < num_edges; ++i) {
      if (u[i] >= num_nodes || v[i] >= num_nodes || u[i] == v[i] || !count[i])
        throw std::invalid_argument("agglomerate: edge with an id outside [0, num_nodes), a self loop or a zero count");
      ++adj_start[(size_t)u[i] + 1];
      ++adj_start[(size_t)v[i] + 1];
    }
    for (int64_t i = 0; i < num_nodes; ++i) adj_start[i + 1] += adj_start[i];
    {
      std::vector<uint32_t> cursor(adj_start.begin(), adj_start.end() - 1);
      for (int64_t i = 0; i < num_edges; ++i) { adj[cursor[u[i]]++] = v[i]; adj[cursor[v[i]]++] = u[i]; }
    }
    std::vector<uint32_t> head((size_t)num_nodes, kNil);
    std::vector<Link> pool;
    auto link = [&](uint32_t from, uint32_t to) {
      pool.push_back(Link{to, head[from]});
      head[from] = (uint32_t)(pool.size() - 1);
    };
    std::vector<Entry> initial;
    double lowest = thr;
    for (int64_t i = 0; i < num_edges; ++i) {
      const uint32_t a = u[i] < v[i] ? u[i] : v[i], b = u[i] < v[i] ? v[i] : u[i];
      const uint64_t key = edge_key(a, b);
      if (edges.find(key)) throw std::invalid_argument("agglomerate: duplicate edge");
      edges.insert(key, sum_fixed[i], key, count[i]);
      const double sc = edge_score(sum_fixed[i], count[i]);
      if (sc < thr) {
        initial.push_back(Entry{sc, key, sum_fixed[i], a, b, count[i]});
        if (sc < lowest) lowest = sc;
      }
    }
    BucketQueue queue(lowest, thr, std::move(initial));
    std::vector<uint8_t> alive((size_t)num_nodes, 1);
    std::vector<uint32_t> parent((size_t)num_nodes), label((size_t)num_nodes), entries((size_t)num_nodes);
    for (int64_t i = 0; i < num_nodes; ++i) {
      parent[i] = label[i] = (uint32_t)i;
      entries[i] = adj_start[i + 1] - adj_start[i];   // list entries (CSR run + overflow), dead ones included
    }
    Entry e;
    while (queue.pop(e)) {
      if (!alive[e.a] || !alive[e.b]) continue;
      EdgeTable::Slot* it = edges.find(edge_key(e.a, e.b));
      if (!it || it->sum != e.sum || it->count != e.count) continue;  // superseded entry
      // (every entry in the queue has score < threshold: nothing to test here; the loop ends when the queue runs dry)
      // the cluster with fewer list entries is merged away (`gone`) into the other (`keep`); equal: the larger root goes
      const bool a_goes = entries[e.a] < entries[e.b] || (entries[e.a] == entries[e.b] && e.a > e.b);
      const uint32_t keep = a_goes ? e.b : e.a, gone = a_goes ? e.a : e.b;
      alive[gone] = 0;
      parent[gone] = keep;
      if (label[gone] < label[keep]) label[keep] = label[gone];
      edges.erase(it);
      auto visit = [&](uint32_t n) {
        if (n == keep || !alive[n]) return;
        EdgeTable::Slot* eg = edges.find(edge_key(gone, n));
        if (!eg) return;  // an entry left behind by an earlier merge
        uint64_t sum = eg->sum, anchor = eg->anchor;
        uint32_t cnt = eg->count;
        edges.erase(eg);
        if (EdgeTable::Slot* ek = edges.find(edge_key(keep, n))) {
          sum += ek->sum;
          cnt += ek->count;
          if (ek->anchor < anchor) anchor = ek->anchor;
          ek->sum = sum;
          ek->count = cnt;
          ek->anchor = anchor;
        } else {
          edges.insert(edge_key(keep, n), sum, anchor, cnt);
          link(keep, n);
          link(n, keep);
          ++entries[keep];
          ++entries[n];
        }
        const double sc = edge_score(sum, cnt);
        if (sc < thr) queue.push(Entry{sc, anchor, sum, keep, n, cnt});
      };
      const uint32_t lo = adj_start[gone], hi = adj_start[gone + 1];
      for (uint32_t i = lo; i < hi; ++i) {   // all the cache misses of this merge in flight at once
        const uint32_t n = adj[i];
        edges.prefetch(edge_key(gone, n));
        edges.prefetch(edge_key(keep, n));
        __builtin_prefetch(&head[n]);
      }
      for (uint32_t i = lo; i < hi; ++i) visit(adj[i]);
      for (uint32_t l = head[gone]; l != kNil;) {
        const Link lk = pool[l];   // (pool may grow in visit(): copy, do not hold a reference)
        l = lk.next;
        visit(lk.node);
      }
      head[gone] = kNil;
    }
    for (int64_t i = 0; i < num_nodes; ++i) {
      uint32_t r = (uint32_t)i;
      while (parent[r] != r) r = parent[r];
      for (uint32_t c = (uint32_t)i; parent[c] != r;) { const uint32_t nx = parent[c]; parent[c] = r; c = nx; }  // compress
      root_of[i] = label[r];
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
