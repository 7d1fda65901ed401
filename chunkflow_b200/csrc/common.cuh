// Shared declarations of the chunkflow_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace cfb {

struct Int3 {
  int z, y, x;
};

inline int64_t vol(const Int3& s) { return (int64_t)s.z * s.y * s.x; }

void set_last_error(const std::string& msg);

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define CFB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      throw ::cfb::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +      \
                             " (" __FILE__ ":" + std::to_string(__LINE__) + ")");             \
    }                                                                                         \
  } while (0)

#define CFB_LAUNCH_CHECK() CFB_CUDA(cudaGetLastError())

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Patch table entry in device memory: chunk-local input start and output-buffer start.
struct PatchPos {
  int iz, iy, ix;  // input patch start inside the chunk
  int oz, oy, ox;  // start of the (cropped) output patch inside the output buffer (may be <0 / clipped)
};

// Per-axis coverage table for the weight-volume gather: for every output coordinate the
// (ascending) list of patch axis-indices that cover it; -1 terminated.
constexpr int kMaxCover = 4;

}  // namespace cfb
