// Device <-> PAGEABLE host memory at PCIe speed.
//
// The reference's contract hands the caller a fresh numpy array (Inferencer.__call__,
// chunkflow/flow/divid_conquer/inferencer.py:360,479): 12.9 GB of pageable, not yet touched memory for a
// 1024^3 chunk.  cudaMemcpyAsync into such memory is staged by the driver through one small pinned buffer
// on the calling thread (a few GB/s, and the first touch of every page is serialised behind it).  Here the
// engine stages the copy itself: finished output planes are copied into a ring of pinned slots on the copy
// stream, and a pool of host threads moves every landed slot into the caller's array (first touch included)
// while later patch rows are still being computed.  The input chunk travels the other way through the same ring (`upload`): host
// threads stage 16 MB pieces into pinned slots in parallel slices, each piece is sent with cudaMemcpyAsync while the next is staged.
#pragma once
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace cfb {

class HostStager {
 public:
  HostStager(int device, size_t slot_bytes = (size_t)16 << 20, int nslots = 24, int nworkers = 0) : device_(device), slot_bytes_(slot_bytes) {
    if (nworkers <= 0) nworkers = (int)std::max(2u, std::min(12u, std::thread::hardware_concurrency() / 4));
    CFB_CUDA(cudaMallocHost(&ring_, slot_bytes * nslots));
    slots_.resize(nslots);
    for (auto& s : slots_) CFB_CUDA(cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming));
    for (int i = 0; i < nworkers; ++i) workers_.emplace_back([this] { worker(); });
    pump_ = std::thread([this] { pump(); });
  }
  ~HostStager() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    if (pump_.joinable()) pump_.join();
    for (auto& t : workers_) t.join();
    for (auto& s : slots_) cudaEventDestroy(s.landed);
    for (cudaEvent_t e : event_pool_) cudaEventDestroy(e);
    for (cudaEvent_t e : up_done_) cudaEventDestroy(e);
    cudaFreeHost(ring_);
  }

  // Called by the thread that enqueues the kernels: `bytes` at d_src are final once everything enqueued on
  // `producer` so far has run; they go to h_dst (pageable).  Returns immediately.
  void push(const void* d_src, void* h_dst, size_t bytes, cudaStream_t producer, cudaStream_t copy_stream) {
    cudaEvent_t ready = nullptr;
    {
      std::lock_guard<std::mutex> g(m_);
      if (!event_pool_.empty()) { ready = event_pool_.back(); event_pool_.pop_back(); }
    }
    if (!ready) CFB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CFB_CUDA(cudaEventRecord(ready, producer));
    {
      std::lock_guard<std::mutex> g(m_);
      requests_.push_back(Request{(const uint8_t*)d_src, (uint8_t*)h_dst, bytes, ready, copy_stream});
      ++open_requests_;
    }
    cv_.notify_all();
  }

  // Pageable host memory -> device at PCIe speed: the worker threads copy piece after piece into pinned ring slots (in parallel
  // slices), each piece goes to the device with cudaMemcpyAsync on `stream` while the next one is being staged.  Blocks until the
  // whole range is ENQUEUED from pinned memory (the caller's buffer is no longer read); the copies complete in stream order.
  void upload(const void* h_src, void* d_dst, size_t bytes, cudaStream_t stream) {
    const int nup = (int)std::min<size_t>(slots_.size(), 8);
    if (up_done_.empty()) {
      up_done_.resize(nup);
      for (auto& e : up_done_) CFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint8_t* src = static_cast<const uint8_t*>(h_src);
    uint8_t* dst = static_cast<uint8_t*>(d_dst);
    const size_t nslices = std::max<size_t>(1, std::min<size_t>(workers_.size(), 8));
    size_t piece = 0;
    for (size_t off = 0; off < bytes; off += slot_bytes_, ++piece) {
      const size_t n = std::min(slot_bytes_, bytes - off);
      const int slot = (int)(piece % nup);
      if (piece >= (size_t)nup) CFB_CUDA(cudaEventSynchronize(up_done_[slot]));  // the previous H2D out of this slot has finished
      uint8_t* stage = ring_ + (size_t)slot * slot_bytes_;
      const size_t slice = ((n + nslices - 1) / nslices + 4095) & ~(size_t)4095;
      {
        std::lock_guard<std::mutex> g(m_);
        for (size_t o = 0; o < n; o += slice) {
          tasks_.push_back(Task{-1, stage + o, std::min(slice, n - o), src + off + o});
          ++open_copies_;
        }
      }
      cv_.notify_all();
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return open_copies_ == 0; });
      }
      CFB_CUDA(cudaMemcpyAsync(dst + off, stage, n, cudaMemcpyHostToDevice, stream));
      CFB_CUDA(cudaEventRecord(up_done_[slot], stream));
    }
    for (size_t i = 0; i < std::min<size_t>(piece, (size_t)nup); ++i) CFB_CUDA(cudaEventSynchronize(up_done_[i]));  // ring slots are free again
  }

  // Blocks until every pushed byte has reached its destination; rethrows a CUDA failure of the helper threads.
  void drain() {
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return (open_requests_ == 0 && open_tasks_ == 0) || failed_; });
    if (failed_) { failed_ = false; open_requests_ = 0; throw CudaError("host staging thread: " + error_); }
  }

 private:
  struct Request { const uint8_t* src; uint8_t* dst; size_t bytes; cudaEvent_t ready; cudaStream_t copy_stream; };
  struct Slot { cudaEvent_t landed = nullptr; bool busy = false; };
  struct Task { int slot; uint8_t* dst; size_t bytes; const uint8_t* src = nullptr; };  // slot < 0: plain host copy src -> dst (upload staging)

  void fail(const std::string& what) {
    std::lock_guard<std::mutex> g(m_);
    failed_ = true; error_ = what;
    requests_.clear(); tasks_.clear(); open_tasks_ = 0;
    cv_.notify_all();
  }

  // one thread turns requests into (D2H into a free pinned slot) + a move task for the workers
  void pump() {
    cudaSetDevice(device_);
    int next = 0;
    for (;;) {
      Request r;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !requests_.empty(); });
        if (stop_) return;
        r = requests_.front();
        requests_.pop_front();
      }
      cudaError_t err = cudaStreamWaitEvent(r.copy_stream, r.ready, 0);
      for (size_t off = 0; off < r.bytes && err == cudaSuccess; off += slot_bytes_) {
        const size_t n = std::min(slot_bytes_, r.bytes - off);
        const int slot = next;
        next = (next + 1) % (int)slots_.size();
        {
          std::unique_lock<std::mutex> g(m_);
          cv_.wait(g, [&] { return stop_ || !slots_[slot].busy; });
          if (stop_) return;
          slots_[slot].busy = true;
        }
        err = cudaMemcpyAsync(ring_ + (size_t)slot * slot_bytes_, r.src + off, n, cudaMemcpyDeviceToHost, r.copy_stream);
        if (err == cudaSuccess) err = cudaEventRecord(slots_[slot].landed, r.copy_stream);
        {
          std::lock_guard<std::mutex> g(m_);
          tasks_.push_back(Task{slot, r.dst + off, n, nullptr});
          ++open_tasks_;
        }
        cv_.notify_all();
      }
      if (err != cudaSuccess) { fail(cudaGetErrorString(err)); continue; }
      {
        std::lock_guard<std::mutex> g(m_);
        event_pool_.push_back(r.ready);  // waited on by the copy stream already (stream order keeps it valid)
        --open_requests_;
      }
      cv_.notify_all();
    }
  }

  void worker() {
    cudaSetDevice(device_);
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !tasks_.empty(); });
        if (stop_) return;
        t = tasks_.front();
        tasks_.pop_front();
      }
      if (t.slot < 0) {   // upload staging: pageable -> pinned slice
        std::memcpy(t.dst, t.src, t.bytes);
        {
          std::lock_guard<std::mutex> g(m_);
          --open_copies_;
        }
        cv_.notify_all();
        continue;
      }
      const cudaError_t err = cudaEventSynchronize(slots_[t.slot].landed);
      if (err != cudaSuccess) { fail(cudaGetErrorString(err)); continue; }
      std::memcpy(t.dst, ring_ + (size_t)t.slot * slot_bytes_, t.bytes);
      {
        std::lock_guard<std::mutex> g(m_);
        slots_[t.slot].busy = false;
        --open_tasks_;
      }
      cv_.notify_all();
    }
  }

  int device_;
  size_t slot_bytes_;
  uint8_t* ring_ = nullptr;
  std::vector<Slot> slots_;
  std::vector<std::thread> workers_;
  std::thread pump_;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<Request> requests_;
  std::deque<Task> tasks_;
  std::vector<cudaEvent_t> event_pool_;
  int open_requests_ = 0, open_tasks_ = 0, open_copies_ = 0;
  std::vector<cudaEvent_t> up_done_;
  bool stop_ = false, failed_ = false;
  std::string error_;
};

}  // namespace cfb
