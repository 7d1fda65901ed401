// tcgen05 / TMA implicit-GEMM 3x3x3 convolution for sm_100a.  See kernels_umma.cuh for the
// activation layout.  One CTA owns a (batch, y-tile, x-tile) column of the patch and marches
// along z with a ring of three z-planes in shared memory:
//
//   warp 0   : A producer  -- one TMA box per z-plane (halo rows/cols, zero fill outside the patch)
//   warp 1   : B producer  -- packed weight blocks per (tap, K-group) via cp.async.bulk
//   warp 2   : MMA issuer  -- per z-plane job: 27 taps x K-steps x G tiles of tcgen05.mma
//                             (M=128 voxels, N=Cout or 2*Cout, K=16) accumulating in TMEM
//   warps 3-6: epilogue    -- tcgen05.ld, bias+ReLU, fp16 (hi/lo) pack, coalesced 16-byte stores
//
// Precision: fp16 operands, fp32 accumulation.  In split mode every value is hi+lo
// (two fp16, ~22 bits): MMA1 = a_hi x [w_hi | w_lo] (N = 2*Cout, one pass over A) and
// MMA2 = a_lo x w_hi (N = Cout); the epilogue adds the two accumulator halves.
#include "kernels_umma.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "act_format.cuh"
#include "chunkflow_b200.h"

namespace cfb {

namespace {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a launch failure (trap) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      printf("chunkflow_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// One elected lane of a converged warp (PTX elect.sync).  ptxas knows the guarded region is
// single-threaded and keeps descriptors in uniform registers; a `lane == 0` test instead makes
// it wrap every tcgen05.mma in an ELECT/branch loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x), "r"(a.y),
               "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
// rows of a TMEM matrix move one lane down (lane i <- lane i+1) inside every 32-lane group; 32-byte elements
__device__ __forceinline__ void tc_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_f16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp8 (e4m3 x e4m3, K = 32) products into the same fp32 accumulators: the correction terms of the f16f8 mode (act_format.cuh)
__device__ __forceinline__ void tc_mma_f8_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared memory matrix descriptor (cute::UMMA::SmemDescriptor), 64 bits:
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4 (between the two 8-element K chunks)
//   [32,46) stride byte offset >> 4 (between 8-row groups)           [46,48) version = 1 (Blackwell)
//   [61,64) layout type = 0 (no swizzle).  Built inline in the MMA issuer as a (hi, lo) pair.
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, M=128.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ReLU that lets NaN through like torch.relu (fmaxf(NaN, 0) would return 0 and hide a broken weight from the range check)
__device__ __forceinline__ float relu_nan(float v) { return v < 0.f ? 0.f : v; }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Fused network tail for the last 3x3x3 layer (Cout = 16): 1x1x1 head + sigmoid + crop + bump mask +
// red.global.add into the output chunk, straight from the fp32 accumulator values.
struct FusedTail {
  const float* head_w;   // (channels, 16) first rows of the head weight
  const float* head_b;   // (channels)
  const PatchPos* patches;
  const float* mask;     // (op.z, op.y, op.x)
  float* out;            // (channels, os.z, os.y, os.x)
  int channels;
  Int3 op, crop, os;
  float scale;
};

template <int COUT, bool SPLIT>
__device__ __forceinline__ void store_cp8_16(const float (&v)[16], int cb, int b, size_t vox, size_t plane_vox,
                                             uint4* __restrict__ out16) {
  constexpr int P = SPLIT ? 2 : 1;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int chunk = cb * 2 + h;
    float hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hi[i] = __half2float(__float2half_rn(v[h * 8 + i]));
    const size_t plane = ((size_t)b * (COUT / 8) + chunk) * P;
    out16[plane * plane_vox + vox] = make_uint4(pack_half2(hi[0], hi[1]), pack_half2(hi[2], hi[3]),
                                                pack_half2(hi[4], hi[5]), pack_half2(hi[6], hi[7]));
    if (SPLIT) {
      float lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) lo[i] = v[h * 8 + i] - hi[i];
      out16[(plane + 1) * plane_vox + vox] = make_uint4(pack_half2(lo[0], lo[1]), pack_half2(lo[2], lo[3]),
                                                        pack_half2(lo[4], lo[5]), pack_half2(lo[6], lo[7]));
    }
  }
}

// f16f8 format: 16 channels (one K step) of a voxel -> H 0..7, H 8..15, A8 0..15, L8 0..15 records
template <int COUT>
__device__ __forceinline__ void store_cp8_16_f8(const float (&v)[16], int cb, int b, size_t vox, size_t plane_vox,
                                                uint4* __restrict__ out16) {
  uint4 h0, h1, a8, l8;
  af_encode16(v, h0, h1, a8, l8);
  const size_t plane = ((size_t)b * (COUT / 8) + cb * 2) * 2;  // (chunk 2cb, part 0)
  out16[plane * plane_vox + vox] = h0;
  out16[(plane + 1) * plane_vox + vox] = a8;
  out16[(plane + 2) * plane_vox + vox] = h1;
  out16[(plane + 3) * plane_vox + vox] = l8;
}

__device__ __forceinline__ void head_blend_16(const float (&v)[16], const FusedTail& t, const float* __restrict__ s_head,
                                              const PatchPos& pp, int z, int y, int x) {
  const int oz = z - t.crop.z, oy = y - t.crop.y, ox = x - t.crop.x;  // coordinates in the cropped output patch
  if (oz < 0 || oz >= t.op.z || oy < 0 || oy >= t.op.y || ox < 0 || ox >= t.op.x) return;
  int sy = oy, sx = ox;
  if (pp.flags) tta_map(pp.flags, t.op.y, t.op.x, oy, ox, sy, sx);  // augmented variant: write back un-transformed
  const int gz = pp.oz + oz, gy = pp.oy + sy, gx = pp.ox + sx;
  if (gz < 0 || gz >= t.os.z || gy < 0 || gy >= t.os.y || gx < 0 || gx >= t.os.x) return;  // clipped by the chunk
  const float m = __ldg(t.mask + ((size_t)oz * t.op.y + oy) * t.op.x + ox) * t.scale;
  float* dst = t.out + ((size_t)gz * t.os.y + gy) * t.os.x + gx;
  const size_t out_vol = (size_t)t.os.z * t.os.y * t.os.x;
  if (t.channels == 3) {
    // the affinity head: three independent FMA chains; sigmoid with ex2.approx / rcp.approx (~1e-7 of the exact value,
    // the tolerance of this path is 1e-3) -- the fused tail is bound by this epilogue, not by the tensor core
    float a0 = s_head[48], a1 = s_head[49], a2 = s_head[50];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = fmaf(v[k], s_head[k], a0);
      a1 = fmaf(v[k], s_head[16 + k], a1);
      a2 = fmaf(v[k], s_head[32 + k], a2);
    }
    float s0 = __fdividef(m, 1.0f + __expf(-a0)), s1 = __fdividef(m, 1.0f + __expf(-a1)), s2 = __fdividef(m, 1.0f + __expf(-a2));
    if (pp.flags & kTtaChannelSym) {  // reference-literal --augment: the variant plus its channel-reversed copy
      const float e = s0 + s2;
      s0 = e; s2 = e; s1 += s1;
    }
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(s0) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + out_vol), "f"(s1) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + 2 * out_vol), "f"(s2) : "memory");
    return;
  }
  float sig[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    if (co < t.channels) {
      float acc = s_head[t.channels * 16 + co];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(v[k], s_head[co * 16 + k], acc);
      sig[co] = __fdiv_rn(1.0f, 1.0f + expf(-acc));
    }
  }
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    if (co < t.channels) {
      float o = sig[co];
      if (pp.flags & kTtaChannelSym) {
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) if (c2 == t.channels - 1 - co) o += sig[c2];
      }
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (size_t)co * out_vol), "f"(o * m) : "memory");
    }
  }
}

#ifdef CFB_TS_TRACE
bool g_trace_print = false;   // set around the post-tuning launch
int g_trace_left = 64;        // print at most this many launches
#endif
// Development instrumentation of the TMEM-shift kernel (build with CFB_NVCC_DEFINES=-DCFB_TS_TRACE): ablation switches
// (-DCFB_TS_ABLATE + env CFB_ABLATE, results become wrong) and per-role mbarrier wait-cycle counters.  Compiled out of the product build:
// with one MMA-issuing thread the kernel was issue bound and even a predicate test per MMA cost ~20 %.
#ifdef CFB_TS_ABLATE
#define CFB_ABL(p, bit) (((p).ablate & (bit)) != 0)
#else
#define CFB_ABL(p, bit) false
#endif
#ifdef CFB_TS_TRACE
#define CFB_TRACE_WAIT(ctr, stmt) do { const long long t_ = clock64(); stmt; (ctr) += clock64() - t_; } while (0)
#else
#define CFB_TRACE_WAIT(ctr, stmt) do { stmt; } while (0)
#endif

struct UmmaConvParams {
  int Z, Y, X;
  int XT, TY, pitch, tile_stride, G;
  int tiles_x, tiles_y;
  int planes_a, planes_b;  // 8-channel chunks of source A / source B
  uint32_t plane_stride;   // bytes of one (TY+2) x pitch x 16 B plane in smem
  uint32_t slot_stride;    // bytes of one z-plane slot (all chunk/part planes)
  const __half* wpacked;
  const float* bias;
  __half* out;
  int relu;
  float acc_scale;  // f16f8 mode: 1 / (alpha * beta), the scale the operands carry (act_format.cuh); 1 otherwise
  int bstages;    // weight block stages in shared memory
  int bresident;  // 1: all 27*KG blocks stay resident (loaded once per CTA), 0: streamed through a ring
  int wide_map;   // 1: 5-D tensor map with the 16-byte record as inner dimension, 0: 4-D map over 8-byte elements
  int T;          // z-stacked kernel: output planes per job
  const __half* wpacked_zs;  // z-stacked weight blocks
  FusedTail tail;  // used by the TAIL = true instantiations only
  int total_items; // z-stacked kernel: work items = batch x tiles x z blocks (persistent CTAs)
  const __half* wpacked_ts;  // TMEM-shift kernel: (dy, kg) triples of z-stacked blocks
  int niss;        // TMEM-shift kernel: MMA-issuing threads (1..4; M tiles are dealt round-robin so every accumulator has ONE issuer)
  int ring;        // TMEM-shift kernel: z-plane slots in shared memory (2 or 3)
  int ngroups;     // TMEM-shift kernel: depth of the ring of A-tile groups in tensor memory (2..4)
  __half* out_pool;  // TMEM-shift kernel, POOL instantiations: (1,2,2) max-pooled copy of the output (CP8, half the y / x extent)
  int ablate;      // CFB_TS_ABLATE builds only: 1 no global stores, 2 no epilogue TMEM reads, 4 no loader copies,
                   // 8 no shifts, 16 no MMAs, 32 no TMA plane loads
  long long* trace;  // CFB_TS_TRACE builds only: 16 cycle counters per CTA
};

constexpr int kRing = 3;       // z-plane ring slots
constexpr int kMaxBStages = 64; // weight block stages: resident (27 * KG <= 54 blocks) or a ring of up to 64
constexpr int kThreads = 224;  // 7 warps
constexpr int kThreadsConvT = 352;  // transposed convolution: 4 more epilogue warps
constexpr int kTailPad = 2304; // dense M tiles may read up to 129 voxel records past the last plane
constexpr int kBufCols = 256;  // TMEM columns per accumulator buffer (2 buffers)
constexpr int kBarBytes = (10 + 2 * 64) * 8 + 16 + 640;  // mbarriers + TMEM base slot + head weights (fused tail)

template <int CIN, int COUT, bool SPLIT>
struct ConvCfg {
  static constexpr int P = SPLIT ? 2 : 1;
  static constexpr int NPL = P * CIN / 8;            // planes per slot
  static constexpr int NB = P * COUT;                // rows of a weight block = N of MMA1
  static constexpr int KB = CIN >= 32 ? 32 : 16;     // channels per weight block
  static constexpr int KG = CIN / KB;                // weight blocks per tap
  static constexpr int KS = KB / 16;                 // K=16 steps per weight block
  static constexpr int BSTAGE = NB * KB * 2;         // bytes
  static constexpr int MAXG = kBufCols / NB;
};

template <int CIN, int COUT, bool SPLIT, bool TAIL>
__global__ void __launch_bounds__(kThreads, 1)
conv3_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const UmmaConvParams p) {
  using Cfg = ConvCfg<CIN, COUT, SPLIT>;
  constexpr int P = Cfg::P;
  extern __shared__ uint8_t smem_raw[];
  // TMA destinations need 128-byte alignment; the launch reserves 128 spare bytes for this
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x % p.tiles_x;
  const int ty = (blockIdx.x / p.tiles_x) % p.tiles_y;
  const int b = blockIdx.x / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.XT, y0 = ty * p.TY;

  uint8_t* sA = smem;
  uint8_t* sB = smem + kRing * p.slot_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.bstages * Cfg::BSTAGE);
  // barrier map: [0..2] a_full, [3..5] a_empty, [6,7] acc_full, [8,9] acc_empty,
  //              [10, 10+64) b_full, [74, 74+64) b_empty
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int kBF = 10, kBE = 10 + kMaxBStages, kAccF = 6, kAccE = 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBE + kMaxBStages);
  float* s_head = reinterpret_cast<float*>(bars + kBE + kMaxBStages + 2);
  PatchPos pp{};
  if constexpr (TAIL) {
    for (int i = threadIdx.x; i < p.tail.channels * 16; i += kThreads) s_head[i] = p.tail.head_w[i];
    for (int i = threadIdx.x; i < p.tail.channels; i += kThreads) s_head[p.tail.channels * 16 + i] = p.tail.head_b[i];
    pp = p.tail.patches[b];
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(BAR(i), 1); mbar_init(BAR(3 + i), 1); }
    for (int i = 0; i < p.bstages; ++i) { mbar_init(BAR(kBF + i), 1); mbar_init(BAR(kBE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(kAccF + i), 1); mbar_init(BAR(kAccE + i), 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Z = p.Z;
  if (warp == 0) {
    // ---------------- A producer: z-planes -1 .. Z into the ring ----------------
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)Cfg::NPL * p.plane_stride;
      const int plane_a0 = b * p.planes_a * P, plane_b0 = b * p.planes_b * P;
      int slot = 0;
      uint32_t use_parity = 1;  // parity of the PREVIOUS use of the slot (no wait during the first round)
      for (int q = 0; q < Z + 2; ++q, ++slot) {
        if (slot == kRing) { slot = 0; use_parity ^= 1; }
        if (q >= kRing) mbar_wait(BAR(3 + slot), use_parity);
        mbar_expect_tx(BAR(slot), tx_bytes);
        const uint32_t dst = smem_u32(sA + (size_t)slot * p.slot_stride);
        const uint32_t dst_b = dst + (uint32_t)(p.planes_a * P) * p.plane_stride;
        if (p.wide_map) {  // 5-D map, 16-byte record as inner dimension (box rows of up to 256 records)
          tma_load_5d(dst, &mapA, BAR(slot), 0, x0 - 1, y0 - 1, q - 1, plane_a0);
          if (p.planes_b > 0) tma_load_5d(dst_b, &mapB, BAR(slot), 0, x0 - 1, y0 - 1, q - 1, plane_b0);
        } else {           // 4-D map over 8-byte elements: a record is two elements (see make_map)
          tma_load_4d(dst, &mapA, BAR(slot), 2 * (x0 - 1), y0 - 1, q - 1, plane_a0);
          if (p.planes_b > 0) tma_load_4d(dst_b, &mapB, BAR(slot), 2 * (x0 - 1), y0 - 1, q - 1, plane_b0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- B producer: weight blocks, 27 * KG per z-plane job ----------------
    if (elect_one()) {
      const uint32_t per_job = 27u * Cfg::KG;
      const uint32_t nbs = (uint32_t)p.bstages;
      // resident: every block is loaded exactly once; ring: blocks stream in job order
      const uint32_t total = p.bresident ? per_job : (uint32_t)Z * per_job;
      uint32_t st = 0, blk = 0, prev_parity = 1;
      for (uint32_t i = 0; i < total; ++i) {
        if (i >= nbs) mbar_wait(BAR(kBE + st), prev_parity);
        mbar_expect_tx(BAR(kBF + st), Cfg::BSTAGE);
        bulk_load(smem_u32(sB + st * Cfg::BSTAGE),
                  reinterpret_cast<const uint8_t*>(p.wpacked) + (size_t)blk * Cfg::BSTAGE, Cfg::BSTAGE, BAR(kBF + st));
        if (++st == nbs) { st = 0; prev_parity ^= 1; }
        if (++blk == per_job) blk = 0;
      }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t IDESC1 = make_idesc(Cfg::NB);
      constexpr uint32_t IDESC2 = make_idesc(COUT);
      // Descriptors are (hi, lo) 32-bit pairs; hi is constant (SBO = 128 B, version 1), lo holds the
      // start address and the leading byte offset, so stepping a tap / tile / K step is one integer add.
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t plane16 = p.plane_stride >> 4;
      const uint32_t a_lbo = (P * plane16) << 16;
      constexpr uint32_t b_lbo = (uint32_t)Cfg::NB << 16;
      const uint32_t sA16 = smem_u32(sA) >> 4, slot16 = p.slot_stride >> 4;
      const uint32_t sB16 = smem_u32(sB) >> 4;
      const uint32_t pitch = (uint32_t)p.pitch, tstride = (uint32_t)p.tile_stride, G = (uint32_t)p.G;
      const uint32_t kstep16 = 2u * P * plane16;  // two 8-channel chunks per K = 16 step
      const bool resident = p.bresident != 0;
      const uint32_t nbs = (uint32_t)p.bstages;
      auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      uint32_t ring_st = 0, ring_parity = 0;
      uint32_t slot_lo = 0, slot_parity = 0;  // ring slot / parity of plane q = z (first plane of the job)
      for (int z = 0; z < Z; ++z) {
        const uint32_t buf = (uint32_t)z & 1u;
        if (z >= 2) mbar_wait(BAR(kAccE + buf), ((z >> 1) - 1) & 1);
        tc_fence_after();
        uint32_t acc = 0;  // the first K step of the job overwrites the accumulators
        uint32_t blk = 0;
        uint32_t slot = slot_lo, sparity = slot_parity;
        const uint32_t d0 = tmem_base + buf * kBufCols;
        for (int dz = 0; dz < 3; ++dz) {
          mbar_wait(BAR(slot), sparity);
          tc_fence_after();
          uint32_t a_row = a_lbo | (sA16 + slot * slot16);
          for (int dy = 0; dy < 3; ++dy, a_row += pitch) {
            for (uint32_t dx = 0; dx < 3; ++dx) {
              const uint32_t a_tap = a_row + dx;
              for (int kg = 0; kg < Cfg::KG; ++kg) {
                uint32_t b16;
                if (resident) {
                  if (z == 0) { mbar_wait(BAR(kBF + blk), 0); tc_fence_after(); }  // blocks arrive once
                  b16 = sB16 + blk * (Cfg::BSTAGE >> 4);
                  ++blk;
                } else {
                  mbar_wait(BAR(kBF + ring_st), ring_parity);
                  tc_fence_after();
                  b16 = sB16 + ring_st * (Cfg::BSTAGE >> 4);
                }
#pragma unroll
                for (int ks = 0; ks < Cfg::KS; ++ks) {
                  const uint64_t bdesc = desc(b_lbo | (b16 + (uint32_t)ks * 2u * Cfg::NB));
                  uint32_t a_lo = a_tap + (uint32_t)(kg * Cfg::KS + ks) * kstep16;
                  uint32_t d = d0;
                  for (uint32_t g = 0; g < G; ++g, a_lo += tstride, d += Cfg::NB) {
                    tc_mma_f16(d, desc(a_lo), bdesc, IDESC1, acc);
                    if (SPLIT) tc_mma_f16(d, desc(a_lo + plane16), bdesc, IDESC2, 1u);
                  }
                  acc = 1;
                }
                if (!resident) {
                  tc_commit(BAR(kBE + ring_st));  // weight block consumed
                  if (++ring_st == nbs) { ring_st = 0; ring_parity ^= 1; }
                }
              }
            }
          }
          if (dz == 0) tc_commit(BAR(3 + slot));  // plane z-1 is dead: the producer may refill its slot
          if (++slot == kRing) { slot = 0; sparity ^= 1; }
        }
        tc_commit(BAR(kAccF + buf));  // accumulators of this z-plane are complete
        if (++slot_lo == kRing) { slot_lo = 0; slot_parity ^= 1; }
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> bias/ReLU -> fp16 (hi/lo) -> HBM ----------------
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int ty_valid = min(p.TY, p.Y - y0), xt_valid = min(p.XT, p.X - x0);
    const size_t plane_vox = (size_t)p.Z * p.Y * p.X;
    const float inv_pitch = 1.0f / (float)p.pitch;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    for (int z = 0; z < Z; ++z) {
      const int buf = z & 1;
      mbar_wait(BAR(kAccF + buf), (z >> 1) & 1);
      tc_fence_after();
      for (int g = 0; g < p.G; ++g) {
        const int m = wq * 32 + lane;
        const int qpos = g * p.tile_stride + m;
        // exact for qpos < 2^20: (qpos + 0.5) / pitch never lands within rounding error of an integer
        const int row = __float2int_rd(((float)qpos + 0.5f) * inv_pitch), col = qpos - row * p.pitch;
        const bool valid = row < ty_valid && col < xt_valid;
        const size_t vox = ((size_t)z * p.Y + (y0 + row)) * p.X + (x0 + col);
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * kBufCols + g * Cfg::NB);
#pragma unroll
        for (int cb = 0; cb < COUT / 16; ++cb) {
          uint32_t r[16];
          tc_ld16(taddr + cb * 16, r);
          float v[16];
          if (SPLIT) {
            uint32_t r2[16];
            tc_ld16(taddr + COUT + cb * 16, r2);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]);
          } else {
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] += __ldg(p.bias + cb * 16 + i);
            if (p.relu) v[i] = relu_nan(v[i]);
          }
          if (valid) {
            if constexpr (TAIL) head_blend_16(v, p.tail, s_head, pp, z, y0 + row, x0 + col);
            else store_cp8_16<COUT, SPLIT>(v, cb, b, vox, plane_vox, out16);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(kAccE + buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}


// ------------------------------------------------------------------------------------------
// z-stacked variant: one activation-tile read serves the three z-taps.
//
// The plain kernel re-reads the 4 KB activation tile of a position run for each of the 27 taps.
// Here a job covers T consecutive OUTPUT planes of the column; INPUT plane q is multiplied, for each
// (dy, dx), by the weight rows of dz = 2, 1, 0 stacked along N (one tcgen05.mma, N = 3 * NB), whose
// column groups land in the accumulators of output planes q-1, q, q+1 -- the same TMEM lanes
// (positions), adjacent column ranges.  A tile is therefore read 9 * (T+2) / T times per output
// plane instead of 27, and N per instruction triples (tensor pipe duty cycle up).  Accumulators are
// zeroed by the epilogue after draining (tcgen05.st), so every MMA accumulates.
// Weight block (dy, dx, kg): [KB/8][3 * NB rows: dz=2 | dz=1 | dz=0, each (hi COUT | lo COUT)][8].
// In split mode the a_lo pass uses the same window (it adds the exact a_lo * w_lo term as well).
// ------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool SPLIT, bool TAIL>
__global__ void __launch_bounds__(kThreads, 1)
conv3_zs_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const UmmaConvParams p) {
  using Cfg = ConvCfg<CIN, COUT, SPLIT>;
  constexpr int P = Cfg::P;
  constexpr int NB = Cfg::NB;
  constexpr int BSTAGE = 3 * Cfg::BSTAGE;  // three dz row groups per block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent CTA: work item = (batch, y tile, x tile, z block); item i is handled by CTA i mod gridDim.x, every
  // warp role walks the same item sequence.  Items are independent (each loads its own halo planes).
  const int njobs = (p.Z + p.T - 1) / p.T;
  const int ncols = p.tiles_x * p.tiles_y;
  struct Item { int b, x0, y0, z0; };
  auto item_of = [&](int item) {
    const int j = item % njobs, col = item / njobs;
    return Item{col / ncols, (col % p.tiles_x) * p.XT, ((col / p.tiles_x) % p.tiles_y) * p.TY, j * p.T};
  };

  uint8_t* sA = smem;
  uint8_t* sB = smem + kRing * p.slot_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.bstages * BSTAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int kBF = 10, kBE = 10 + kMaxBStages, kAccF = 6, kAccE = 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBE + kMaxBStages);
  float* s_head = reinterpret_cast<float*>(bars + kBE + kMaxBStages + 2);
  PatchPos pp{};
  if constexpr (TAIL) {
    for (int i = threadIdx.x; i < p.tail.channels * 16; i += kThreads) s_head[i] = p.tail.head_w[i];
    for (int i = threadIdx.x; i < p.tail.channels; i += kThreads) s_head[p.tail.channels * 16 + i] = p.tail.head_b[i];
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(BAR(i), 1); mbar_init(BAR(3 + i), 1); }
    for (int i = 0; i < p.bstages; ++i) { mbar_init(BAR(kBF + i), 1); mbar_init(BAR(kBE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(kAccF + i), 1); mbar_init(BAR(kAccE + i), 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Z = p.Z, T = p.T;

  if (warp == 0) {
    // ---------------- A producer: for every job its input planes max(z0-1,0) .. min(z0+T, Z-1) ----------------
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)Cfg::NPL * p.plane_stride;
      int slot = 0, ld = 0;
      uint32_t prev_parity = 1;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item it = item_of(item);
        const int x0 = it.x0, y0 = it.y0, z0 = it.z0;
        const int plane_a0 = it.b * p.planes_a * P, plane_b0 = it.b * p.planes_b * P;
        const int qlo = max(z0 - 1, 0), qhi = min(z0 + T, Z - 1);
        for (int q = qlo; q <= qhi; ++q, ++ld) {
          if (ld >= kRing) mbar_wait(BAR(3 + slot), prev_parity);
          mbar_expect_tx(BAR(slot), tx_bytes);
          const uint32_t dst = smem_u32(sA + (size_t)slot * p.slot_stride);
          const uint32_t dst_b = dst + (uint32_t)(p.planes_a * P) * p.plane_stride;
          if (p.wide_map) {
            tma_load_5d(dst, &mapA, BAR(slot), 0, x0 - 1, y0 - 1, q, plane_a0);
            if (p.planes_b > 0) tma_load_5d(dst_b, &mapB, BAR(slot), 0, x0 - 1, y0 - 1, q, plane_b0);
          } else {
            tma_load_4d(dst, &mapA, BAR(slot), 2 * (x0 - 1), y0 - 1, q, plane_a0);
            if (p.planes_b > 0) tma_load_4d(dst_b, &mapB, BAR(slot), 2 * (x0 - 1), y0 - 1, q, plane_b0);
          }
          if (++slot == kRing) { slot = 0; prev_parity ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- B producer: 9 * KG blocks per input plane ----------------
    if (elect_one()) {
      const uint32_t per_plane = 9u * Cfg::KG;
      const uint32_t nbs = (uint32_t)p.bstages;
      uint32_t planes = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int z0 = item_of(item).z0;
        planes += (uint32_t)(min(z0 + T, Z - 1) - max(z0 - 1, 0) + 1);
      }
      const uint32_t total = p.bresident ? per_plane : planes * per_plane;
      uint32_t st = 0, blk = 0, prev_parity = 1;
      for (uint32_t i = 0; i < total; ++i) {
        if (i >= nbs) mbar_wait(BAR(kBE + st), prev_parity);
        mbar_expect_tx(BAR(kBF + st), BSTAGE);
        bulk_load(smem_u32(sB + st * BSTAGE), reinterpret_cast<const uint8_t*>(p.wpacked_zs) + (size_t)blk * BSTAGE,
                  BSTAGE, BAR(kBF + st));
        if (++st == nbs) { st = 0; prev_parity ^= 1; }
        if (++blk == per_plane) blk = 0;
      }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t plane16 = p.plane_stride >> 4;
      const uint32_t a_lbo = (P * plane16) << 16;
      constexpr uint32_t b_lbo = (uint32_t)(3 * NB) << 16;
      const uint32_t sA16 = smem_u32(sA) >> 4, slot16 = p.slot_stride >> 4;
      const uint32_t sB16 = smem_u32(sB) >> 4;
      const uint32_t pitch = (uint32_t)p.pitch, tstride = (uint32_t)p.tile_stride, G = (uint32_t)p.G;
      const uint32_t kstep16 = 2u * P * plane16;
      const bool resident = p.bresident != 0;
      const uint32_t nbs = (uint32_t)p.bstages;
      auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      uint32_t ring_st = 0, ring_parity = 0;
      uint32_t slot = 0, sparity = 0;
      bool first_plane = true;
      int jj = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++jj) {
        const uint32_t buf = (uint32_t)jj & 1u;
        mbar_wait(BAR(kAccE + buf), (uint32_t)(jj >> 1) & 1u);  // zeroed by the epilogue (initially and after draining)
        tc_fence_after();
        const int z0 = item_of(item).z0, z1 = min(z0 + T, Z), qlo = max(z0 - 1, 0), qhi = min(z0 + T, Z - 1);
        const uint32_t dbuf = tmem_base + buf * kBufCols;
        for (int q = qlo; q <= qhi; ++q) {
          mbar_wait(BAR(slot), sparity);
          tc_fence_after();
          const int plo = max(q - 1, z0), phi = min(q + 1, z1 - 1);
          const uint32_t ng = (uint32_t)(phi - plo + 1);
          const uint32_t row0 = (uint32_t)(2 - (q - plo + 1)) * NB;  // first weight row group: dz = q - plo + 1
          const uint32_t idesc = make_idesc((int)(ng * NB));
          const uint32_t dcol0 = dbuf + (uint32_t)(plo - z0) * NB;
          uint32_t a_row = a_lbo | (sA16 + slot * slot16);
          uint32_t blk = 0;
          for (int dy = 0; dy < 3; ++dy, a_row += pitch) {
            for (uint32_t dx = 0; dx < 3; ++dx) {
              const uint32_t a_tap = a_row + dx;
              for (int kg = 0; kg < Cfg::KG; ++kg) {
                uint32_t b16;
                if (resident) {
                  if (first_plane) { mbar_wait(BAR(kBF + blk), 0); tc_fence_after(); }
                  b16 = sB16 + blk * (BSTAGE >> 4);
                  ++blk;
                } else {
                  mbar_wait(BAR(kBF + ring_st), ring_parity);
                  tc_fence_after();
                  b16 = sB16 + ring_st * (BSTAGE >> 4);
                }
#pragma unroll
                for (int ks = 0; ks < Cfg::KS; ++ks) {
                  const uint64_t bdesc = desc(b_lbo | (b16 + (uint32_t)ks * 2u * (3 * NB) + row0));
                  uint32_t a_lo = a_tap + (uint32_t)(kg * Cfg::KS + ks) * kstep16;
                  uint32_t d = dcol0;
                  for (uint32_t g = 0; g < G; ++g, a_lo += tstride, d += (uint32_t)T * NB) {
                    tc_mma_f16(d, desc(a_lo), bdesc, idesc, 1u);
                    if (SPLIT) tc_mma_f16(d, desc(a_lo + plane16), bdesc, idesc, 1u);
                  }
                }
                if (!resident) {
                  tc_commit(BAR(kBE + ring_st));
                  if (++ring_st == nbs) { ring_st = 0; ring_parity ^= 1; }
                }
              }
            }
          }
          first_plane = false;
          tc_commit(BAR(3 + slot));  // this input plane is consumed
          if (++slot == kRing) { slot = 0; sparity ^= 1; }
        }
        tc_commit(BAR(kAccF + buf));
      }
    }
  } else {
    // ---------------- epilogue ----------------
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    // zero both accumulator buffers once (TMEM is not cleared by allocation)
    {
      const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < 512; c += 16) tc_st16(tmem_base + lane_base + c, zero);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(kAccE + 0));
      mbar_arrive(BAR(kAccE + 1));
    }
    const size_t plane_vox = (size_t)p.Z * p.Y * p.X;
    const float inv_pitch = 1.0f / (float)p.pitch;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    int jj = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++jj) {
      const int buf = jj & 1;
      const Item it = item_of(item);
      const int b = it.b, x0 = it.x0, y0 = it.y0, z0 = it.z0, z1 = min(z0 + T, Z);
      const int ty_valid = min(p.TY, p.Y - y0), xt_valid = min(p.XT, p.X - x0);
      if constexpr (TAIL) pp = p.tail.patches[b];
      mbar_wait(BAR(kAccF + buf), (uint32_t)(jj >> 1) & 1u);
      tc_fence_after();
      for (int g = 0; g < p.G; ++g) {
        const int m = wq * 32 + lane;
        const int qpos = g * p.tile_stride + m;
        const int row = __float2int_rd(((float)qpos + 0.5f) * inv_pitch), col = qpos - row * p.pitch;
        const bool valid = row < ty_valid && col < xt_valid;
        for (int pz = z0; pz < z1; ++pz) {
          const size_t vox = ((size_t)pz * p.Y + (y0 + row)) * p.X + (x0 + col);
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * kBufCols + (g * T + (pz - z0)) * NB);
#pragma unroll
          for (int cb = 0; cb < COUT / 16; ++cb) {
            uint32_t r[16];
            tc_ld16(taddr + cb * 16, r);
            float v[16];
            if (SPLIT) {
              uint32_t r2[16];
              tc_ld16(taddr + COUT + cb * 16, r2);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]);
            } else {
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] += __ldg(p.bias + cb * 16 + i);
              if (p.relu) v[i] = relu_nan(v[i]);
            }
            if (valid) {
              if constexpr (TAIL) head_blend_16(v, p.tail, s_head, pp, pz, y0 + row, x0 + col);
              else store_cp8_16<COUT, SPLIT>(v, cb, b, vox, plane_vox, out16);
            }
          }
          // clear the accumulator for its next use
          {
            const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < NB; c += 16) tc_st16(taddr + c, zero);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(kAccE + buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------
// z-stacked + TMEM-shift variant ("TS"): the dx taps re-use one TMEM-resident activation tile.
//
// On top of the z-stacking above, the A operand of the MMAs comes from TENSOR MEMORY: four loader warps
// copy the 128-position tile of one (input plane, dy, K step, hi/lo part) from the shared-memory plane
// into an 8-column TMEM buffer (row = lane, K = 16 fp16 = 8 packed columns; tcgen05.st), and the single
// MMA thread issues   MMA(dx=0) ; tcgen05.shift.down ; MMA(dx=1) ; tcgen05.shift.down ; MMA(dx=2)
// back to back -- the shift moves row i+1 into lane i inside every 32-lane group (measured with
// tools/probes/probe_tmem.cu: executes in issue order with the MMAs; the last lane of a group keeps its
// value), i.e. the dx = +1 tap.  A tile is thus read from shared memory ONCE per (dy, K step, part) by
// ordinary ld.shared instead of three times by the tensor core's operand fetch, which was the limiter.
// Because two lanes per 32-lane group go stale, an M tile carries 4 x 30 = 120 valid positions:
// lane (k, i) <-> position 120 g + 30 k + i, i < 30 valid.
// Weight blocks: (dy, kg) triples of the z-stacked blocks of dx = 0, 1, 2.
// ------------------------------------------------------------------------------------------
#ifndef CFB_TS_LOADER_SETS
#define CFB_TS_LOADER_SETS 1
#endif
constexpr int kTsLoaderSets = CFB_TS_LOADER_SETS;  // sets of 4 TMEM-loader warps; set s fills the A-tile groups n with n % sets == s
constexpr int kThreadsTS = 384 + 128 * (kTsLoaderSets - 1);   // warps: A producer, B producer, MMA issuer 0, 4 epilogue, 4 TMEM loaders, MMA issuer 1 (, 4 more loaders)
constexpr int kThreadsTSTail = 512;  // fused tail: 4 more epilogue warps (12..15) and ONE loader set (a 640-thread CTA would cap at 96 registers and spill)
constexpr int kTsGroups = 4;      // at most: ring of A-tile GROUPS in TMEM (16 tiles / (P * G) of them); a group = all (part, tile) A tiles of one (plane, dy, K step)
constexpr int kTsMaxIssuers = 2;   // measured: 3 or 4 issuers are no faster than 2 (the kernel is no longer issue bound)
constexpr int kTsMaxTiles = 8;    // P * G <= 8 tiles of 8 columns per group
constexpr int kTsACol0 = 384;     // groups live at columns [384, 512)
constexpr int kTsAccCols = 192;   // accumulator columns per buffer (2 buffers)
constexpr int kTsBarBytes = (10 + 2 * 64 + 2 * kTsGroups) * 8 + 16 + 640;

template <int CIN, int COUT, bool SPLIT, bool TAIL, bool F8 = false, bool POOL = false>
__global__ void __launch_bounds__(TAIL ? kThreadsTSTail : kThreadsTS, 1)
conv3_ts_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const UmmaConvParams p) {
  using Cfg = ConvCfg<CIN, COUT, SPLIT>;
  constexpr int P = Cfg::P;
  constexpr int NB = Cfg::NB;
  constexpr int BSTAGE = 9 * Cfg::BSTAGE;  // (dy, kg) triple: dx = 0,1,2; block rows = [hi: dz 2,1,0 | lo: dz 2,1,0] x COUT
  // With the A operand in tensor memory an extra pass over A is free, so hi/lo products are NOT concatenated
  // along N here: a_hi x w_hi, a_hi x w_lo and a_lo x w_hi all accumulate into the same COUT columns of a plane
  // (half the TMEM per plane -> twice the z-stacking depth T, fewer tensor cycles and fewer weight bytes per tap).
  constexpr int KSTEPS = CIN / 16;
  constexpr int kEpiSets = TAIL ? 2 : 1;  // the fused tail is epilogue bound: 8 epilogue warps (launched with 512 threads), else 4
  constexpr int LS = TAIL ? 1 : kTsLoaderSets;  // sets of four TMEM-loader warps
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int njobs = (p.Z + p.T - 1) / p.T;
  const int ncols = p.tiles_x * p.tiles_y;
  struct Item { int b, x0, y0, z0; };
  auto item_of = [&](int item) {
    const int j = item % njobs, col = item / njobs;
    return Item{col / ncols, (col % p.tiles_x) * p.XT, ((col / p.tiles_x) % p.tiles_y) * p.TY, j * p.T};
  };

  const int ring = p.ring;  // plane slots (2 or 3)
  uint8_t* sA = smem;
  uint8_t* sB = smem + ring * p.slot_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.bstages * BSTAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int kBF = 10, kBE = 10 + kMaxBStages, kAccF = 6, kAccE = 8;
  constexpr int kTF = 10 + 2 * kMaxBStages, kTE = kTF + kTsGroups;  // TMEM A-tile group full / empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTE + kTsGroups);
  float* s_head = reinterpret_cast<float*>(bars + kTE + kTsGroups + 2);
  // POOL: staging buffer [TY][XT / 2][16] fp32 of the x-pooled values of one (plane, 16-channel block), behind the barrier area
  [[maybe_unused]] float4* s_pool = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + kTsBarBytes);
  PatchPos pp{};
  if constexpr (TAIL) {
    for (int i = threadIdx.x; i < p.tail.channels * 16; i += blockDim.x) s_head[i] = p.tail.head_w[i];
    for (int i = threadIdx.x; i < p.tail.channels; i += blockDim.x) s_head[p.tail.channels * 16 + i] = p.tail.head_b[i];
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(BAR(i), 1); mbar_init(BAR(3 + i), 4 * LS); }  // plane slots are released by the loader warps
    for (int i = 0; i < p.bstages; ++i) { mbar_init(BAR(kBF + i), 1); mbar_init(BAR(kBE + i), p.niss); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(kAccF + i), p.niss); mbar_init(BAR(kAccE + i), 128 * kEpiSets); }
    for (int i = 0; i < kTsGroups; ++i) { mbar_init(BAR(kTF + i), 128); mbar_init(BAR(kTE + i), p.niss); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Z = p.Z, T = p.T;

  if (warp == 0) {
    // ---------------- A producer (TMA z-plane ring), as in the z-stacked kernel ----------------
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)Cfg::NPL * p.plane_stride;
      [[maybe_unused]] long long tr_wait0 = 0, tr_t0 = clock64();
      int slot = 0, ld = 0;
      uint32_t prev_parity = 1;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item it = item_of(item);
        const int x0 = it.x0, y0 = it.y0, z0 = it.z0;
        const int plane_a0 = it.b * p.planes_a * P, plane_b0 = it.b * p.planes_b * P;
        const int qlo = max(z0 - 1, 0), qhi = min(z0 + T, Z - 1);
        for (int q = qlo; q <= qhi; ++q, ++ld) {
          if (ld >= ring) CFB_TRACE_WAIT(tr_wait0, mbar_wait(BAR(3 + slot), prev_parity));
          if (CFB_ABL(p, 32)) {
            mbar_arrive(BAR(slot));
          } else {
          mbar_expect_tx(BAR(slot), tx_bytes);
          const uint32_t dst = smem_u32(sA + (size_t)slot * p.slot_stride);
          const uint32_t dst_b = dst + (uint32_t)(p.planes_a * P) * p.plane_stride;
          if (p.wide_map) {
            tma_load_5d(dst, &mapA, BAR(slot), 0, x0 - 1, y0 - 1, q, plane_a0);
            if (p.planes_b > 0) tma_load_5d(dst_b, &mapB, BAR(slot), 0, x0 - 1, y0 - 1, q, plane_b0);
          } else {
            tma_load_4d(dst, &mapA, BAR(slot), 2 * (x0 - 1), y0 - 1, q, plane_a0);
            if (p.planes_b > 0) tma_load_4d(dst_b, &mapB, BAR(slot), 2 * (x0 - 1), y0 - 1, q, plane_b0);
          }
          }
          if (++slot == ring) { slot = 0; prev_parity ^= 1; }
        }
      }
#ifdef CFB_TS_TRACE
      if (p.trace) { p.trace[blockIdx.x * 16 + 9] = clock64() - tr_t0; p.trace[blockIdx.x * 16 + 10] = tr_wait0; }
#endif
    }
  } else if (warp == 1) {
    // ---------------- B producer: 3 * KG (dy, kg) triples per input plane ----------------
    if (elect_one()) {
      const uint32_t per_plane = 3u * Cfg::KG;
      const uint32_t nbs = (uint32_t)p.bstages;
      uint32_t planes = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int z0 = item_of(item).z0;
        planes += (uint32_t)(min(z0 + T, Z - 1) - max(z0 - 1, 0) + 1);
      }
      const uint32_t total = p.bresident ? per_plane : planes * per_plane;
      uint32_t st = 0, blk = 0, prev_parity = 1;
      for (uint32_t i = 0; i < total; ++i) {
        if (i >= nbs) mbar_wait(BAR(kBE + st), prev_parity);
        mbar_expect_tx(BAR(kBF + st), BSTAGE);
        for (int part = 0; part < 9; ++part)  // bulk copies of at most one per-tap block each
          bulk_load(smem_u32(sB + st * BSTAGE + part * Cfg::BSTAGE),
                    reinterpret_cast<const uint8_t*>(p.wpacked_ts) + (size_t)blk * BSTAGE + (size_t)part * Cfg::BSTAGE,
                    Cfg::BSTAGE, BAR(kBF + st));
        if (++st == nbs) { st = 0; prev_parity ^= 1; }
        if (++blk == per_plane) blk = 0;
      }
    }
  } else if (warp == 2 || warp == 11) {
    // ---------------- MMA issuers: issuer i owns the M tiles g = i, i + niss, ... (and their accumulators) ----------------
    const uint32_t iss = warp == 2 ? 0u : (uint32_t)(warp - 10);
    if (iss < (uint32_t)p.niss && elect_one()) {
      const uint32_t niss = (uint32_t)p.niss;
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      constexpr uint32_t b_lbo = (uint32_t)(3 * NB) << 16;
      const uint32_t sB16 = smem_u32(sB) >> 4;
      const uint32_t G = (uint32_t)p.G;
      const uint32_t dstep1 = (uint32_t)p.T * COUT, dstep = niss * dstep1, astep = niss * 8;
      const uint32_t ngroups = (uint32_t)p.ngroups, gstride = (uint32_t)(P * p.G * 8);
      const bool resident = p.bresident != 0;
      const uint32_t nbs = (uint32_t)p.bstages;
      auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      uint32_t ring_st = 0, ring_parity = 0;
      uint32_t tgrp = 0, tparity = 0;  // TMEM A-tile group ring position
      bool first_plane = true;
      [[maybe_unused]] long long tr_acc = 0, tr_tf = 0, tr_b = 0, tr_issue = 0, tr_t0 = clock64();
      int jj = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++jj) {
        const uint32_t buf = (uint32_t)jj & 1u;
        CFB_TRACE_WAIT(tr_acc, mbar_wait(BAR(kAccE + buf), (uint32_t)(jj >> 1) & 1u));
        tc_fence_after();
        const int z0 = item_of(item).z0, z1 = min(z0 + T, Z), qlo = max(z0 - 1, 0), qhi = min(z0 + T, Z - 1);
        const uint32_t dbuf = tmem_base + buf * kTsAccCols;
        for (int q = qlo; q <= qhi; ++q) {
          const int plo = max(q - 1, z0), phi = min(q + 1, z1 - 1);
          const uint32_t ng = (uint32_t)(phi - plo + 1);
          const uint32_t row0 = (uint32_t)(2 - (q - plo + 1)) * COUT;  // first dz row group inside the hi (or lo) rows
#ifdef CFB_TS_SMALLN
          const uint32_t idesc = make_idesc(16);  // experiment: quarter-size MMAs (wrong results) to tell issue-bound from tensor-bound
#else
          const uint32_t idesc = make_idesc((int)(ng * COUT));
#endif
          const uint32_t dcol0 = dbuf + (uint32_t)(plo - z0) * COUT;
          uint32_t blk = 0;
          for (int dy = 0; dy < 3; ++dy) {
            for (int kg = 0; kg < Cfg::KG; ++kg) {
              uint32_t b16;
              if (resident) {
                if (first_plane) { CFB_TRACE_WAIT(tr_b, mbar_wait(BAR(kBF + blk), 0)); tc_fence_after(); }
                b16 = sB16 + blk * (BSTAGE >> 4);
                ++blk;
              } else {
                CFB_TRACE_WAIT(tr_b, mbar_wait(BAR(kBF + ring_st), ring_parity));
                tc_fence_after();
                b16 = sB16 + ring_st * (BSTAGE >> 4);
              }
#pragma unroll
              for (int ks = 0; ks < Cfg::KS; ++ks) {
                const uint32_t bk = b16 + (uint32_t)ks * 2u * (3 * NB) + row0;
                CFB_TRACE_WAIT(tr_tf, mbar_wait(BAR(kTF + tgrp), tparity));  // the loader warps filled this group of TMEM A tiles
                tc_fence_after();
                const uint32_t a_grp = tmem_base + kTsACol0 + tgrp * gstride + iss * 8;
#ifdef CFB_TS_TRACE
                const long long tr_i0 = clock64();
#endif
                // (issuing the shifts and the MMAs of all tiles of a group interleaved -- all shifts of a dx step, then all its MMAs --
                //  to space dependent instructions apart measured 15 % SLOWER on every layer: profiles/r02_ab_issue_order.txt)
                for (int part = 0; part < P; ++part) {  // A tile part: 0 = hi, 1 = lo
                  uint32_t d = dcol0 + iss * dstep1;
                  uint32_t a_tm = a_grp + (uint32_t)part * G * 8;
                  for (uint32_t g = iss; g < G; g += niss, d += dstep, a_tm += astep) {
#pragma unroll
                    for (uint32_t dx = 0; dx < 3; ++dx) {
                      if (dx && !CFB_ABL(p, 8)) tc_shift_down(a_tm);
                      const uint32_t bt = bk + dx * (3 * Cfg::BSTAGE >> 4);
                      if (CFB_ABL(p, 16)) continue;
                      if constexpr (F8) {
                        if (part == 0) tc_mma_f16_ta(d, a_tm, desc(b_lbo | bt), idesc, 1u);
                        else tc_mma_f8_ta(d, a_tm, desc(b_lbo | (bt + 3 * COUT)), idesc, 1u);
                      } else {
                        tc_mma_f16_ta(d, a_tm, desc(b_lbo | bt), idesc, 1u);                                  // x w_hi
                        if (SPLIT && part == 0) tc_mma_f16_ta(d, a_tm, desc(b_lbo | (bt + 3 * COUT)), idesc, 1u);  // a_hi x w_lo
                      }
                    }
                  }
                }
#ifdef CFB_TS_TRACE
                tr_issue += clock64() - tr_i0;
#endif
                tc_commit(BAR(kTE + tgrp));
                if (++tgrp == ngroups) { tgrp = 0; tparity ^= 1; }
              }
              if (!resident) {
                tc_commit(BAR(kBE + ring_st));
                if (++ring_st == nbs) { ring_st = 0; ring_parity ^= 1; }
              }
            }
          }
          first_plane = false;
        }
        tc_commit(BAR(kAccF + buf));
      }
#ifdef CFB_TS_TRACE
      if (p.trace) {
        long long* t = p.trace + blockIdx.x * 16 + (iss ? 12 : 0);
        if (iss > 1) {
        } else if (iss) { t[0] = clock64() - tr_t0; t[1] = tr_issue; }
        else { t[0] = clock64() - tr_t0; t[1] = tr_acc; t[2] = tr_tf; t[3] = tr_b; t[11] = tr_issue; }
      }
#endif
    }
  } else if ((warp >= 7 && warp <= 10) || (LS == 2 && warp >= 12)) {
    // ---------------- TMEM loaders: shared-memory plane -> A tiles in tensor memory ----------------
    // (the copy of a group is a latency chain -- ld.shared, tcgen05.st, wait::st, arrive --, so two sets of four
    //  warps take alternate groups)
    const uint32_t lset = warp >= 11 ? 1u : 0u;
    const int wq = warp & 3;                    // lane quarter this warp may access
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t pos_in_tile = (uint32_t)(30 * wq + lane);  // lane (k, i) <-> position 30 k + i
    const uint32_t plane16 = p.plane_stride >> 4;
    const uint4* sA16 = reinterpret_cast<const uint4*>(sA);
    uint32_t slot = 0, sparity = 0;
    uint32_t tgrp = 0, tparity = 1;  // parity of the previous release (none during the first round)
    uint32_t gcount = 0;
    const uint32_t ngroups = (uint32_t)p.ngroups, gstride = (uint32_t)(P * p.G * 8);
    auto next_group = [&]() { ++gcount; if (++tgrp == ngroups) { tgrp = 0; tparity ^= 1; } };
    [[maybe_unused]] long long tr_slot = 0, tr_te = 0, tr_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int z0 = item_of(item).z0, qlo = max(z0 - 1, 0), qhi = min(z0 + T, Z - 1);
      for (int q = qlo; q <= qhi; ++q) {
        CFB_TRACE_WAIT(tr_slot, mbar_wait(BAR(slot), sparity));  // TMA has landed this z-plane
        const uint4* plane = sA16 + (size_t)slot * (p.slot_stride >> 4);
        for (int dy = 0; dy < 3; ++dy) {
          for (int kstep = 0; kstep < KSTEPS; ++kstep) {
            if (LS == 2 && (gcount & 1u) != lset) { next_group(); continue; }  // the other set's group
            // one group = the P * G A tiles of this (plane, dy, K step): loads first, then the TMEM stores, one hand-off
            uint4 c0[kTsMaxTiles], c1[kTsMaxTiles];
            const uint4* src = plane + (size_t)(kstep * 2) * P * plane16 + dy * p.pitch + pos_in_tile;
#pragma unroll
            for (int i = 0; i < kTsMaxTiles; ++i) {
              const int part = i / p.G, g = i - part * p.G;  // i = part * G + g
              if (i < P * p.G && !CFB_ABL(p, 4)) {
                const uint4* sp = src + (size_t)part * plane16 + g * 120;
                c0[i] = sp[0];
                c1[i] = sp[(size_t)P * plane16];
              }
            }
            if (gcount >= ngroups) CFB_TRACE_WAIT(tr_te, mbar_wait(BAR(kTE + tgrp), tparity));
            tc_fence_after();
            const uint32_t t0 = tmem_base + lane_base + kTsACol0 + tgrp * gstride;
#pragma unroll
            for (int i = 0; i < kTsMaxTiles; ++i)
              if (i < P * p.G && !CFB_ABL(p, 4)) tc_st8(t0 + i * 8, c0[i], c1[i]);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(BAR(kTF + tgrp));
            next_group();
          }
        }
        // the tensor core never reads A from shared memory in this kernel: once this warp's copies of the plane are
        // in tensor memory (tcgen05.wait::st above) the slot can be refilled by TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(3 + slot));
        if (++slot == ring) { slot = 0; sparity ^= 1; }
      }
    }
#ifdef CFB_TS_TRACE
    if (p.trace && warp == 7 && lane == 0) {
      long long* t = p.trace + blockIdx.x * 16;
      t[4] = clock64() - tr_t0; t[5] = tr_slot; t[6] = tr_te;
    }
#endif
  } else {
    // ---------------- epilogue (warps 3..6 = set 0, warps 12..15 = set 1): the two warps of a lane quarter take
    // alternate (tile, plane) steps -- a step is a latency chain (tcgen05.ld, convert, store / head + sigmoid + red) ----------------
    const int wq = warp & 3;
    const int eset = (kEpiSets == 2 && warp >= 12) ? 1 : 0;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    {
      const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = eset * kTsAccCols; c < (kEpiSets == 2 ? eset + 1 : 2) * kTsAccCols; c += 16) tc_st16(tmem_base + lane_base + c, zero);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(kAccE + 0));
      mbar_arrive(BAR(kAccE + 1));
    }
    const size_t plane_vox = (size_t)p.Z * p.Y * p.X;
    const float inv_pitch = 1.0f / (float)p.pitch;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    int jj = 0;
    [[maybe_unused]] long long tr_accf = 0, tr_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++jj) {
      const int buf = jj & 1;
      const Item it = item_of(item);
      const int b = it.b, x0 = it.x0, y0 = it.y0, z0 = it.z0, z1 = min(z0 + T, Z);
      const int ty_valid = min(p.TY, p.Y - y0), xt_valid = min(p.XT, p.X - x0);
      if constexpr (TAIL) pp = p.tail.patches[b];
      CFB_TRACE_WAIT(tr_accf, mbar_wait(BAR(kAccF + buf), (uint32_t)(jj >> 1) & 1u));
      tc_fence_after();
      if constexpr (POOL) {
        // Fused (1,2,2) max pooling (the layer's output is both a skip connection and, pooled, the next level's input): per
        // (plane, 16-channel block) every thread stores its full-resolution voxel, takes the max with its x neighbour (the
        // adjacent lane: tile bases, pitch and x0 are even) and the even lane parks the pair in shared memory; the y partner of a
        // row lives in another warp's lane quarter or in the next M tile, so after a barrier of the four epilogue warps thread t
        // finishes pooled voxel t from two staged rows, encodes and stores it.
        const int hx = p.XT >> 1;
        const int npool = (p.TY >> 1) * hx;
        const size_t pool_plane_vox = (size_t)p.Z * (p.Y >> 1) * (p.X >> 1);
        uint4* pool16 = reinterpret_cast<uint4*>(p.out_pool);
        const int tid = (warp - 3) * 32 + lane;   // 0..127 over the epilogue warps 3..6
        for (int pz = z0; pz < z1; ++pz) {
#pragma unroll
          for (int cb = 0; cb < COUT / 16; ++cb) {
            for (int g = 0; g < p.G; ++g) {
              const int qpos = g * 120 + 30 * wq + lane;
              const int row = __float2int_rd(((float)qpos + 0.5f) * inv_pitch), col = qpos - row * p.pitch;
              const bool valid = lane < 30 && row < ty_valid && col < xt_valid;
              const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * kTsAccCols + (g * T + (pz - z0)) * COUT) + cb * 16;
              uint32_t r[16];
              tc_ld16(taddr, r);
              float v[16];
              tc_wait_ld();
              {
                const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                tc_st16(taddr, zero);   // clear the accumulator columns for their next use
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]);
                if constexpr (F8) v[i] *= p.acc_scale;
                v[i] += __ldg(p.bias + cb * 16 + i);
                if (p.relu) v[i] = relu_nan(v[i]);
              }
              if (valid) {
                const size_t vox = ((size_t)pz * p.Y + (y0 + row)) * p.X + (x0 + col);
                if constexpr (F8) store_cp8_16_f8<COUT>(v, cb, b, vox, plane_vox, out16);
                else store_cp8_16<COUT, SPLIT>(v, cb, b, vox, plane_vox, out16);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
              if (valid && !(lane & 1)) {
                float4* dst = s_pool + ((size_t)row * hx + (col >> 1)) * 4;
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                dst[2] = make_float4(v[8], v[9], v[10], v[11]);
                dst[3] = make_float4(v[12], v[13], v[14], v[15]);
              }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps: every row of the plane is staged
            for (int t = tid; t < npool; t += 128) {
              const int pr = t / hx, pc = t - pr * hx;
              if (2 * pr < ty_valid && 2 * pc < xt_valid) {
                const float4* a = s_pool + ((size_t)(2 * pr) * hx + pc) * 4;
                const float4* c = s_pool + ((size_t)(2 * pr + 1) * hx + pc) * 4;
                float m[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 u = a[q], w = c[q];
                  m[4 * q] = fmaxf(u.x, w.x); m[4 * q + 1] = fmaxf(u.y, w.y); m[4 * q + 2] = fmaxf(u.z, w.z); m[4 * q + 3] = fmaxf(u.w, w.w);
                }
                const size_t pvox = ((size_t)pz * (p.Y >> 1) + ((y0 >> 1) + pr)) * (p.X >> 1) + ((x0 >> 1) + pc);
                if constexpr (F8) store_cp8_16_f8<COUT>(m, cb, b, pvox, pool_plane_vox, pool16);
                else store_cp8_16<COUT, SPLIT>(m, cb, b, pvox, pool_plane_vox, pool16);
              }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");   // staged rows are consumed: the buffer may be rewritten
          }
        }
      } else
      for (int g = 0; g < p.G; ++g) {
        const int qpos = g * 120 + 30 * wq + lane;
        const int row = __float2int_rd(((float)qpos + 0.5f) * inv_pitch), col = qpos - row * p.pitch;
        const bool valid = lane < 30 && row < ty_valid && col < xt_valid;
        for (int pz = z0; pz < z1; ++pz) {
          if (kEpiSets == 2 && ((g * T + (pz - z0)) & 1) != eset) continue;
          const size_t vox = ((size_t)pz * p.Y + (y0 + row)) * p.X + (x0 + col);
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * kTsAccCols + (g * T + (pz - z0)) * COUT);
#pragma unroll
          for (int cb = 0; cb < COUT / 16; ++cb) {
            uint32_t r[16];
            if (!CFB_ABL(p, 2)) tc_ld16(taddr + cb * 16, r);
            float v[16];
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            if constexpr (F8) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= p.acc_scale;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] += __ldg(p.bias + cb * 16 + i);
              if (p.relu) v[i] = relu_nan(v[i]);
            }
            if (valid && !CFB_ABL(p, 1)) {
              if constexpr (TAIL) head_blend_16(v, p.tail, s_head, pp, pz, y0 + row, x0 + col);
              else if constexpr (F8) store_cp8_16_f8<COUT>(v, cb, b, vox, plane_vox, out16);
              else store_cp8_16<COUT, SPLIT>(v, cb, b, vox, plane_vox, out16);
            }
          }
          {
            const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < COUT; c += 16) tc_st16(taddr + c, zero);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(kAccE + buf));
    }
#ifdef CFB_TS_TRACE
    if (p.trace && warp == 3 && lane == 0) {
      long long* t = p.trace + blockIdx.x * 16;
      t[7] = clock64() - tr_t0; t[8] = tr_accf;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------
// Transposed convolution kernel = stride = (1,2,2) on tcgen05: a plain GEMM per input voxel,
//   D[voxel, (tap, cout)] = sum_ci A[voxel, ci] * W[ci, cout, tap],   tap = (a, b) in {0,1}^2,
// followed by a scatter epilogue: tap (a, b) of input voxel (z, y, x) is output voxel
// (z, 2y+a, 2x+b).  Same pipeline as the 3x3x3 kernel but no halo (box = tile) and one "tap";
// the weight block rows are ordered [hi: 4 taps x COUT | lo: 4 taps x COUT] so that in split mode
// MMA1 = a_hi x [w_hi | w_lo] (N = 8*COUT) and MMA2 = a_lo x w_hi (N = 4*COUT, first half).
// ------------------------------------------------------------------------------------------
struct UmmaConvTParams {
  int Z, Y, X;       // INPUT size; output is (Z, 2Y, 2X)
  int XT, TY, G;     // tile of TY x XT input voxels = G M-tiles of 128 positions (dense, pitch = XT)
  int tiles_x, tiles_y;
  int planes;        // 8-channel chunks of the input
  uint32_t plane_stride, slot_stride;
  const __half* wpacked;
  const float* bias;
  __half* out;
  float acc_scale;   // f16f8 mode: 1 / (alpha * beta)
  int ring;          // z-plane slots in shared memory (3; 2 where two CTAs per SM would not fit otherwise)
};

template <int CIN, int COUT, bool SPLIT, bool F8 = false>
__global__ void __launch_bounds__(kThreadsConvT, 1)
convT_umma_kernel(const __grid_constant__ CUtensorMap mapA, const UmmaConvTParams p) {
  constexpr int P = SPLIT ? 2 : 1;
  constexpr int NPL = P * CIN / 8;
  constexpr int N1 = 4 * P * COUT, N2 = 4 * COUT;
  constexpr int KS = CIN / 16;
  constexpr int WBYTES = N1 * CIN * 2;
  // (f16f8 needs only N2 accumulator columns per tile; running TWO CTAs per SM on 256 columns each was measured SLOWER --
  //  up1 2.5 -> 3.0 ms, up0 4.8 -> 5.0 ms per 99 patches -- so every mode keeps one CTA per SM and the N1 column stride)
  constexpr int ACC = N1;                    // accumulator columns per M tile
  constexpr int TCOLS = 512;                 // tensor-memory columns of this CTA (2 buffers)
  constexpr int BUF = TCOLS / 2;
  const int ring = p.ring;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x % p.tiles_x;
  const int ty = (blockIdx.x / p.tiles_x) % p.tiles_y;
  const int b = blockIdx.x / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.XT, y0 = ty * p.TY;

  uint8_t* sA = smem;
  uint8_t* sB = smem + ring * p.slot_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + WBYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };  // [0..2] a_full [3..5] a_empty [6,7] acc_full [8,9] acc_empty [10] w_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(BAR(i), 1); mbar_init(BAR(3 + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(6 + i), 1); mbar_init(BAR(8 + i), 256); }  // two sets of four epilogue warps
    mbar_init(BAR(10), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Z = p.Z;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)NPL * p.plane_stride;
      int slot = 0;
      uint32_t prev_parity = 1;
      for (int z = 0; z < Z; ++z, ++slot) {
        if (slot == ring) { slot = 0; prev_parity ^= 1; }
        if (z >= ring) mbar_wait(BAR(3 + slot), prev_parity);
        mbar_expect_tx(BAR(slot), tx_bytes);
        tma_load_4d(smem_u32(sA + (size_t)slot * p.slot_stride), &mapA, BAR(slot), 2 * x0, y0, z, b * p.planes * P);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_expect_tx(BAR(10), WBYTES);
      for (int off = 0; off < WBYTES; off += 8192)
        bulk_load(smem_u32(sB + off), reinterpret_cast<const uint8_t*>(p.wpacked) + off, min(8192, WBYTES - off), BAR(10));
    }
  } else if (warp == 2) {
    if (elect_one()) {
      constexpr uint32_t IDESC1 = make_idesc(N1);
      constexpr uint32_t IDESC2 = make_idesc(N2);
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t plane16 = p.plane_stride >> 4;
      const uint32_t a_lbo = (P * plane16) << 16;
      constexpr uint32_t b_lbo = (uint32_t)N1 << 16;
      const uint32_t sA16 = smem_u32(sA) >> 4, slot16 = p.slot_stride >> 4, sB16 = smem_u32(sB) >> 4;
      auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      mbar_wait(BAR(10), 0);
      tc_fence_after();
      uint32_t slot = 0, sparity = 0;
      for (int z = 0; z < Z; ++z) {
        const uint32_t buf = (uint32_t)z & 1u;
        if (z >= 2) mbar_wait(BAR(8 + buf), ((z >> 1) - 1) & 1);
        mbar_wait(BAR(slot), sparity);
        tc_fence_after();
        const uint32_t a0 = a_lbo | (sA16 + slot * slot16);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          // weight block: [CIN/8 chunks][N1 rows][8]; K step ks = chunks 2ks, 2ks+1
          const uint64_t bdesc = desc(b_lbo | (sB16 + (uint32_t)ks * 2u * N1));
          uint32_t a_lo = a0 + (uint32_t)ks * 2u * P * plane16;
          uint32_t d = tmem_base + buf * BUF;
          for (int g = 0; g < p.G; ++g, a_lo += 128, d += ACC) {
            if constexpr (F8) {
              // H x WH (rows 0..N2) in fp16, [A8 | L8] x [WL8 ; W8] (rows N2..2 N2, K = 32) in e4m3, same accumulator columns
              tc_mma_f16(d, desc(a_lo), bdesc, IDESC2, ks == 0 ? 0u : 1u);
              tc_mma_f8(d, desc(a_lo + plane16), bdesc + (uint64_t)N2, IDESC2, 1u);
            } else {
              tc_mma_f16(d, desc(a_lo), bdesc, IDESC1, ks == 0 ? 0u : 1u);
              if (SPLIT) tc_mma_f16(d, desc(a_lo + plane16), bdesc, IDESC2, 1u);
            }
          }
        }
        tc_commit(BAR(3 + slot));
        tc_commit(BAR(6 + buf));
        if (++slot == ring) { slot = 0; sparity ^= 1; }
      }
    }
  } else {
    // epilogue (warps 3..6 = set 0, 7..10 = set 1): the scatter of the four taps is a chain of tcgen05.ld, encode and
    // stores -- the two warps of a lane quarter take one output row (y tap) each and write its two x taps as 32-byte stores
    const int wq = warp & 3, eset = warp >= 7 ? 1 : 0;
    const int ty_valid = min(p.TY, p.Y - y0), xt_valid = min(p.XT, p.X - x0);
    const int OY = 2 * p.Y, OX = 2 * p.X;
    const size_t oplane_vox = (size_t)p.Z * OY * OX;
    const float inv_xt = 1.0f / (float)p.XT;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    for (int z = 0; z < Z; ++z) {
      const int buf = z & 1;
      mbar_wait(BAR(6 + buf), (z >> 1) & 1);
      tc_fence_after();
      for (int g = 0; g < p.G; ++g) {
        const int m = wq * 32 + lane;
        const int qpos = g * 128 + m;
        const int row = __float2int_rd(((float)qpos + 0.5f) * inv_xt), col = qpos - row * p.XT;
        const bool valid = row < ty_valid && col < xt_valid;
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * BUF + g * ACC);
        {
          // The two x taps of this set's output row are adjacent 16-byte records of every plane: encode both and write each plane
          // with ONE 32-byte store (st.global.v8.b32) -- whole sectors, half the store instructions (two half-sector stores from
          // different instructions ran at 65 % of the HBM peak, this runs at 92 %).
          const size_t ovox = ((size_t)z * OY + (2 * (y0 + row) + eset)) * OX + 2 * (x0 + col);
          constexpr int NREC = F8 ? 4 : 2 * P;   // records (planes) per 16 channels
#pragma unroll
          for (int cb = 0; cb < COUT / 16; ++cb) {
            uint4 rec[2][4];
#pragma unroll
            for (int bx = 0; bx < 2; ++bx) {
              const int t = eset * 2 + bx;
              uint32_t r[16];
              tc_ld16(taddr + t * COUT + cb * 16, r);
              float v[16];
              if constexpr (SPLIT && !F8) {
                uint32_t r2[16];
                tc_ld16(taddr + N2 + t * COUT + cb * 16, r2);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]) + __ldg(p.bias + cb * 16 + i);
              } else {
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(r[i]), F8 ? p.acc_scale : 1.0f, __ldg(p.bias + cb * 16 + i));
              }
              if constexpr (F8) {
                af_encode16(v, rec[bx][0], rec[bx][2], rec[bx][1], rec[bx][3]);  // plane order: H 0..7, A8, H 8..15, L8
              } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  float hi[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) hi[i] = __half2float(__float2half_rn(v[h * 8 + i]));
                  rec[bx][h * P] = make_uint4(pack_half2(hi[0], hi[1]), pack_half2(hi[2], hi[3]), pack_half2(hi[4], hi[5]), pack_half2(hi[6], hi[7]));
                  if constexpr (SPLIT)
                    rec[bx][h * P + 1] = make_uint4(pack_half2(v[h * 8] - hi[0], v[h * 8 + 1] - hi[1]), pack_half2(v[h * 8 + 2] - hi[2], v[h * 8 + 3] - hi[3]),
                                                    pack_half2(v[h * 8 + 4] - hi[4], v[h * 8 + 5] - hi[5]), pack_half2(v[h * 8 + 6] - hi[6], v[h * 8 + 7] - hi[7]));
                }
              }
            }
            if (valid) {
              const size_t plane = ((size_t)b * (COUT / 8) + cb * 2) * P;
#pragma unroll
              for (int k = 0; k < NREC; ++k) {
                uint4* dst = out16 + (plane + k) * oplane_vox + ovox;
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(rec[0][k].x), "r"(rec[0][k].y),
                             "r"(rec[0][k].z), "r"(rec[0][k].w), "r"(rec[1][k].x), "r"(rec[1][k].y), "r"(rec[1][k].z), "r"(rec[1][k].w)
                             : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(8 + buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
  }
}

// ------------------------------------------------------------------------------------------
// First layer (Cin = 1) on tcgen05, fused with patch extraction: the 27 taps of a voxel are the K
// dimension (padded to 32).  Producer warps gather the uint8 neighbourhood of 128 consecutive x
// positions straight from the chunk (through the test-time-augmentation coordinate map) and write
// the im2col tile in the canonical K-major layout; uint8 values are exact in fp16, so one MMA
// pass x [w_hi | w_lo] gives the fp32-accurate sum of w * x, and the epilogue divides by 255
// (the reference normalises x / 255 first: inferencer.py:395-399), adds bias, ReLU, writes CP8.
// ------------------------------------------------------------------------------------------
struct FirstConvParams {
  const uint8_t* chunk;
  Int3 cs;
  const PatchPos* patches;
  Int3 ps;
  const __half* wpacked;  // [4 chunks of 8 taps][NBR rows][8]
  const float* bias;
  __half* out;
  int tiles_x, total_tiles;
};

constexpr int kFcThreads = 288;  // 4 producer warps, 1 MMA warp, 4 epilogue warps
constexpr int kFcAcc = 4;        // accumulator slots in TMEM

template <bool SPLIT>
__global__ void __launch_bounds__(kFcThreads, 1) first_conv_umma_kernel(const FirstConvParams p) {
  constexpr int P = SPLIT ? 2 : 1;
  constexpr int NBR = 16 * P;           // weight rows: hi (16) | lo (16)
  constexpr int ABYTES = 4 * 128 * 16;  // one im2col tile: 4 K-chunks x 128 rows x 16 B
  constexpr int WBYTES = 4 * NBR * 16;
  __shared__ __align__(128) uint8_t sA[2][ABYTES];
  __shared__ __align__(128) uint8_t sW[WBYTES];
  __shared__ __align__(8) uint64_t bars[2 + 2 + 2 * kFcAcc + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // [0,1] a_full (128 producer arrivals)  [2,3] a_empty  [4..7] acc_full  [8..11] acc_empty (128)  [12] w_full
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(i), 128); mbar_init(BAR(2 + i), 1); }
    for (int i = 0; i < kFcAcc; ++i) { mbar_init(BAR(4 + i), 1); mbar_init(BAR(8 + i), 128); }
    mbar_init(BAR(12), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int X = p.ps.x, Y = p.ps.y, Z = p.ps.z;

  auto tile_of = [&](int t, int& b, int& z, int& y, int& x0) {
    x0 = (t % p.tiles_x) * 128;
    int r = t / p.tiles_x;
    y = r % Y; r /= Y;
    z = r % Z;
    b = r / Z;
  };

  if (warp < 4) {
    // ---------------- producers: im2col tile of 128 x-positions ----------------
    const int m = threadIdx.x;
    int k = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++k) {
      const int buf = k & 1;
      if (k >= 2) mbar_wait(BAR(2 + buf), ((k >> 1) - 1) & 1);
      int b, z, y, x0;
      tile_of(t, b, z, y, x0);
      const PatchPos pp = p.patches[b];
      const int x = x0 + m;
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = 0u;
      if (x < X) {
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
          const int zz = z + tap / 9 - 1, yy = y + (tap / 3) % 3 - 1, xx = x + tap % 3 - 1;
          float v = 0.f;
          if (zz >= 0 && zz < Z && yy >= 0 && yy < Y && xx >= 0 && xx < X) {
            int sy = yy, sx = xx;
            if (pp.flags) tta_map(pp.flags, Y, X, yy, xx, sy, sx);
            v = (float)__ldg(p.chunk + ((size_t)(pp.iz + zz) * p.cs.y + (pp.iy + sy)) * p.cs.x + pp.ix + sx);
          }
          const uint32_t hv = (uint32_t)__half_as_ushort(__float2half_rn(v));
          h[tap >> 1] |= (tap & 1) ? (hv << 16) : hv;
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(sA[buf]);
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c * 128 + m] = make_uint4(h[c * 4], h[c * 4 + 1], h[c * 4 + 2], h[c * 4 + 3]);
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(BAR(buf));
    }
  } else if (warp == 4) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      mbar_expect_tx(BAR(12), WBYTES);
      bulk_load(smem_u32(sW), p.wpacked, WBYTES, BAR(12));
      mbar_wait(BAR(12), 0);
      tc_fence_after();
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      constexpr uint32_t IDESC = make_idesc(NBR);
      auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      const uint32_t w16 = smem_u32(sW) >> 4;
      int k = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++k) {
        const int buf = k & 1, a = k % kFcAcc;
        if (k >= kFcAcc) mbar_wait(BAR(8 + a), ((k / kFcAcc) - 1) & 1);
        mbar_wait(BAR(buf), (k >> 1) & 1);
        tc_fence_after();
        const uint32_t a16 = smem_u32(sA[buf]) >> 4;
        const uint32_t d = tmem_base + (uint32_t)a * NBR;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)  // K = 32 taps = two K=16 steps = chunk pairs (0,1), (2,3)
          tc_mma_f16(d, desc((128u << 16) | (a16 + ks * 256)), desc(((uint32_t)NBR << 16) | (w16 + ks * 2 * NBR)), IDESC,
                     ks ? 1u : 0u);
        tc_commit(BAR(2 + buf));
        tc_commit(BAR(4 + a));
      }
    }
  } else {
    // ---------------- epilogue ----------------
    const int wq = warp & 3;
    const size_t plane_vox = (size_t)Z * Y * X;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    int k = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++k) {
      const int a = k % kFcAcc;
      int b, z, y, x0;
      tile_of(t, b, z, y, x0);
      mbar_wait(BAR(4 + a), (k / kFcAcc) & 1);
      tc_fence_after();
      const int x = x0 + wq * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)a * NBR;
      uint32_t r[16];
      tc_ld16(taddr, r);
      float v[16];
      if (SPLIT) {
        uint32_t r2[16];
        tc_ld16(taddr + 16, r2);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]);
      } else {
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
      mbar_arrive(BAR(8 + a));
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(__fdiv_rn(v[i], 255.0f) + __ldg(p.bias + i), 0.f);
      if (x < X) store_cp8_16<16, SPLIT>(v, 0, b, ((size_t)z * Y + y) * X + x, plane_vox, out16);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

// ------------------------------------------------------------------------------------------
// Host side: tensor maps, tile selection, launch
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !ptr) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// CP8 tensor (planes, Z, Y, X, 8 x fp16) as a TMA tensor map.  Two flavours:
//  * narrow (4-D over 8-byte elements): the 16-byte voxel records of one x row are contiguous in
//    memory, so (x, channel) is ONE inner dimension of 2*X uint64 elements and a box row is one
//    contiguous run (full 32-byte sectors; box rows limited to 256 elements = 128 records);
//  * wide (5-D, the 16-byte record is the inner dimension): box rows of up to 256 records, but
//    the TMA engine fetches one 32-byte sector per record (2x the L2->SM traffic).
// Out-of-bounds elements are zero-filled either way: SAME padding at the patch border.
CUtensorMap make_map(const __half* base, int planes, Int3 sz, int bx, int by, int bplanes, bool wide) {
  CUtensorMap m;
  CUresult r;
  if (wide) {
    cuuint64_t gdim[5] = {8, (cuuint64_t)sz.x, (cuuint64_t)sz.y, (cuuint64_t)sz.z, (cuuint64_t)planes};
    cuuint64_t gstr[4] = {16, (cuuint64_t)sz.x * 16, (cuuint64_t)sz.x * sz.y * 16, (cuuint64_t)sz.x * sz.y * sz.z * 16};
    cuuint32_t box[5] = {8, (cuuint32_t)bx, (cuuint32_t)by, 1, (cuuint32_t)bplanes};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    if (bx > 128) throw std::runtime_error("narrow TMA box row limited to 128 voxel records");
    cuuint64_t gdim[4] = {(cuuint64_t)sz.x * 2, (cuuint64_t)sz.y, (cuuint64_t)sz.z, (cuuint64_t)planes};
    cuuint64_t gstr[3] = {(cuuint64_t)sz.x * 16, (cuuint64_t)sz.x * sz.y * 16, (cuuint64_t)sz.x * sz.y * sz.z * 16};
    cuuint32_t box[4] = {(cuuint32_t)bx * 2, (cuuint32_t)by, 1, (cuuint32_t)bplanes};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<__half*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

constexpr int kMaxSmem = 232448;  // 227 KB

// All feasible tilings of one layer, cheapest first by a static cost model:
//   (halo amplification of the z-plane loads)/2 + (M-tile positions per useful output), scaled by
//   the wave quantisation of the grid, + a penalty for shallow weight rings.
template <int CIN, int COUT, bool SPLIT>
std::vector<ConvTile> enumerate_tiles(int nb, Int3 sz, int sm_count, bool ts_only = false, bool pool = false) {
  using Cfg = ConvCfg<CIN, COUT, SPLIT>;
  std::vector<ConvTile> out;
  const int ty_cap = std::min(16, (sz.y + 1) & ~1);
  std::vector<std::pair<int, bool>> xts;  // (XT, wide map)
  if (sz.x >= 128) xts.push_back({128, true});
  for (int k = 1; k <= 16; ++k) {
    int xt = ceil_div(sz.x, k);
    xt += xt & 1;
    if (xt + 2 > 128 || (xt < 16 && k > 1)) continue;
    bool dup = false;
    for (auto& e : xts) dup |= (e.first == xt && !e.second);
    if (!dup) xts.push_back({xt, false});
  }
  for (auto& e : xts) {
    const int XT = e.first, pitch = XT + 2;
    const bool aligned = XT == 128;  // one 128-voxel M tile per row; otherwise dense 128-position runs
    for (int tyc = ty_cap; tyc >= 2; tyc -= 2) {
      const size_t plane = (size_t)(tyc + 2) * pitch * 16;
      const size_t slot = (Cfg::NPL * plane + 127) / 128 * 128;
      const int G = aligned ? tyc : ceil_div(tyc * pitch, 128);
      if (G > Cfg::MAXG || slot >= (1u << 18)) continue;
      const size_t fixed = (size_t)kBarBytes + kTailPad + 128;
      if ((size_t)kRing * slot + fixed >= (size_t)kMaxSmem) continue;
      const size_t room = (size_t)kMaxSmem - (size_t)kRing * slot - fixed;
      const int all_blocks = 27 * Cfg::KG;
      int bs = (int)std::min<size_t>(room / Cfg::BSTAGE, kMaxBStages);
      if (bs >= all_blocks) bs = all_blocks;
      if (bs < 3) continue;
      ConvTile t;
      t.T = 0;
      t.XT = XT; t.TY = tyc; t.bstages = bs; t.resident = bs == all_blocks; t.wide = e.second;
      const double useful = (double)std::min(tyc, sz.y) * std::min(XT, sz.x);
      const double lookahead = t.resident ? 1e9 : (double)(bs - 1) * G * Cfg::KS * Cfg::P;
      const int ctas = nb * ceil_div(sz.x, XT) * ceil_div(sz.y, tyc);
      const double waves = (double)ctas / sm_count;
      const double quant = std::ceil(waves) / waves;
      t.cost = (((double)(tyc + 2) * pitch / useful) * 0.5 + ((double)G * 128 / useful)) * quant +
               (lookahead >= 96 ? 0.0 : 0.4 * (96 - lookahead) / 96);
      out.push_back(t);
    }
    // z-stacked variants: T output planes per job, accumulators (G * T * NB columns) per TMEM buffer
    if (3 * Cfg::NB <= 256 && !getenv("CFB_NO_ZSTACK")) {
      for (int T : {2, 3, 4, 6, 8}) {
        if (T > sz.z && T != 2) continue;
        for (int tyc = ty_cap; tyc >= 2; tyc -= 2) {
          const size_t plane = (size_t)(tyc + 2) * pitch * 16;
          const size_t slot = (Cfg::NPL * plane + 127) / 128 * 128;
          const int G = aligned ? tyc : ceil_div(tyc * pitch, 128);
          if (G * T * Cfg::NB > kBufCols || slot >= (1u << 18)) continue;
          const size_t fixed = (size_t)kBarBytes + kTailPad + 128;
          if ((size_t)kRing * slot + fixed >= (size_t)kMaxSmem) continue;
          const size_t room = (size_t)kMaxSmem - (size_t)kRing * slot - fixed;
          const int all_blocks = 9 * Cfg::KG;
          const int bstage = 3 * Cfg::BSTAGE;
          int bs = (int)std::min<size_t>(room / bstage, kMaxBStages);
          if (bs >= all_blocks) bs = all_blocks;
          if (bs < 2) continue;
          ConvTile t;
          t.T = T; t.XT = XT; t.TY = tyc; t.bstages = bs; t.resident = bs == all_blocks; t.wide = e.second;
          const double useful = (double)std::min(tyc, sz.y) * std::min(XT, sz.x);
          const double lookahead = t.resident ? 1e9 : (double)(bs - 1) * G * Cfg::KS * Cfg::P * 3;
          const double quant = 1.0;  // persistent CTAs over (column, z block) items: no wave quantisation
          // one tile read serves three z-taps: (T + 2) / (3 T) of the plain kernel's operand traffic
          const double wbytes = t.resident ? 0.0 : 27.0 * Cfg::KG * Cfg::BSTAGE * (double)(T + 2) / ((double)T * useful);
          t.cost = (((double)(tyc + 2) * pitch / useful) * 0.5 + ((double)G * 128 / useful) * ((double)(T + 2) / (3.0 * T) + 0.25)) * quant +
                   (lookahead >= 96 ? 0.0 : 0.4 * (96 - lookahead) / 96) + 0.3 * wbytes / 1024.0;
          out.push_back(t);
        }
      }
    }
  }
  // z-stacked + TMEM-shift variants: M tiles of 120 positions, accumulators G * T * NB <= 224 columns per buffer
  if (ts_only || (!getenv("CFB_NO_TSHIFT") && !getenv("CFB_NO_ZSTACK"))) {
    for (auto& e : xts) {
      const int XT = e.first, pitch = XT + 2;
      for (int T : {2, 3, 4, 6, 8}) {
        if (T > sz.z && T != 2) continue;
        for (int tyc = ty_cap; tyc >= 2; tyc -= 2) {
          const size_t plane = (size_t)(tyc + 2) * pitch * 16;
          const size_t slot = (Cfg::NPL * plane + 127) / 128 * 128;
          const int G = ceil_div(tyc * pitch, 120);
          if (G * T * COUT > kTsAccCols || Cfg::P * G > kTsMaxTiles || slot >= (1u << 18)) continue;
          const size_t fixed = (size_t)kTsBarBytes + kTailPad + 128 + (pool ? (size_t)tyc * (XT / 2) * 64 : 0);  // + pooling staging buffer
          for (int ring = 3; ring >= 2; --ring) {
          // two plane slots are enough in steady state (the loaders free a slot as soon as its copies are in tensor memory)
          // and leave room for larger tiles / resident weights on the wide layers
          if (ring == 2 && CIN < 32) continue;
          if ((size_t)ring * slot + fixed >= (size_t)kMaxSmem) continue;
          const size_t room = (size_t)kMaxSmem - (size_t)ring * slot - fixed;
          const int all_blocks = 3 * Cfg::KG;
          const int bstage = 9 * Cfg::BSTAGE;
          int bs = (int)std::min<size_t>(room / bstage, kMaxBStages);
          if (bs >= all_blocks) bs = all_blocks;
          if (bs < 2) continue;
          ConvTile t;
          t.ring = ring;
          t.T = T; t.shift = true; t.XT = XT; t.TY = tyc; t.bstages = bs; t.resident = bs == all_blocks; t.wide = e.second;
          const double useful = (double)std::min(tyc, sz.y) * std::min(XT, sz.x);
          // tensor/B-bound rather than tile-read bound: about half the z-stacked kernel's cost per position
          // streamed weights are re-fetched from L2 for every (input plane, tile group): bytes per useful output
          const double wbytes = t.resident ? 0.0 : 27.0 * Cfg::KG * Cfg::BSTAGE * (double)(T + 2) / ((double)T * useful);
          t.cost = (((double)(tyc + 2) * pitch / useful) * 0.5 + ((double)G * 128 / useful) * ((double)(T + 2) / (3.0 * T)) * 0.6) +
                   0.3 * wbytes / 1024.0 + (ring == 2 ? 0.02 : 0.0);
          out.push_back(t);
          }
        }
      }
    }
  }
  if (ts_only) {  // f16f8 mode: only the TMEM-shift kernel implements the mixed fp16 / e4m3 products
    std::vector<ConvTile> only;
    for (const ConvTile& t : out) if (t.shift) only.push_back(t);
    out.swap(only);
  }
  if (const char* force = getenv("CFB_FORCE_ZSTACK")) {  // tests: exercise one kernel variant only
    const int T = atoi(force);
    const bool want_shift = ts_only || getenv("CFB_FORCE_TSHIFT") != nullptr;
    std::vector<ConvTile> only;
    for (const ConvTile& t : out) if (t.T == T && t.shift == want_shift) only.push_back(t);
    if (!only.empty()) out.swap(only);
  }
  std::sort(out.begin(), out.end(), [](const ConvTile& a, const ConvTile& b) { return a.cost < b.cost; });
  return out;
}

int sm_count();

template <int CIN, int COUT, bool SPLIT, bool F8 = false>
void launch_tile(const ConvTile& t, const __half* srcA, int ca, const __half* srcB, int cb, const PackedConv& w,
                 __half* out, int nb, Int3 sz, bool relu, cudaStream_t s, const FusedTail* tail = nullptr, __half* pool_out = nullptr) {
  using Cfg = ConvCfg<CIN, COUT, SPLIT>;
  UmmaConvParams p{};
  p.Z = sz.z; p.Y = sz.y; p.X = sz.x;
  p.XT = t.XT;
  p.pitch = p.XT + 2;
  const bool row_aligned = p.XT == 128;
  p.tile_stride = row_aligned ? p.pitch : 128;
  p.tiles_x = ceil_div(sz.x, p.XT);
  p.TY = t.TY;
  p.bstages = t.bstages;
  p.bresident = t.resident ? 1 : 0;
  p.wide_map = t.wide ? 1 : 0;
  p.G = t.shift ? ceil_div(p.TY * p.pitch, 120) : (row_aligned ? p.TY : ceil_div(p.TY * p.pitch, 128));
  p.tiles_y = ceil_div(sz.y, p.TY);
  p.plane_stride = (uint32_t)((p.TY + 2) * p.pitch * 16);
  p.slot_stride = (uint32_t)((Cfg::NPL * (size_t)p.plane_stride + 127) / 128 * 128);
  p.planes_a = ca / 8; p.planes_b = cb / 8;
  p.wpacked = w.w; p.wpacked_zs = w.w_zs; p.wpacked_ts = w.w_ts; p.bias = w.bias; p.out = out; p.relu = relu ? 1 : 0;
  p.acc_scale = w.acc_scale;
  if (F8 && !t.shift) throw std::runtime_error("the f16f8 mode runs on the TMEM-shift kernel only");
  p.T = t.T;
  {
    static const int max_iss = [] { const char* e = std::getenv("CFB_TS_NISS"); return e ? atoi(e) : kTsMaxIssuers; }();
    p.niss = t.shift ? std::max(1, std::min<int>(std::min<int>(p.G, max_iss), kTsMaxIssuers)) : 1;
    p.ngroups = t.shift ? std::max(2, std::min<int>(kTsGroups, 16 / (Cfg::P * p.G))) : 0;
#ifdef CFB_TS_ABLATE
    static const int ablate = [] { const char* e = std::getenv("CFB_ABLATE"); return e ? atoi(e) : 0; }();
    p.ablate = ablate;
#endif
#ifdef CFB_TS_TRACE
    static long long* trace = [] { long long* t = nullptr; cudaMalloc(&t, 256 * 16 * sizeof(long long)); return t; }();
    p.trace = g_trace_print ? trace : nullptr;
    if (p.trace) cudaMemsetAsync(trace, 0, 256 * 16 * sizeof(long long), s);
#endif
  }
  const size_t bstage = t.shift ? 9 * (size_t)Cfg::BSTAGE : (t.T ? 3 * (size_t)Cfg::BSTAGE : (size_t)Cfg::BSTAGE);
  p.ring = t.shift ? t.ring : kRing;
  p.out_pool = pool_out;
  const size_t pool_bytes = (pool_out && t.shift) ? (size_t)p.TY * (p.XT / 2) * 64 : 0;
  const size_t smem = (size_t)p.ring * p.slot_stride + (size_t)p.bstages * bstage + (t.shift ? kTsBarBytes : kBarBytes) + kTailPad + 128 + pool_bytes;
  const CUtensorMap mapA = make_map(srcA, nb * p.planes_a * Cfg::P, sz, p.pitch, p.TY + 2, p.planes_a * Cfg::P, t.wide);
  const CUtensorMap mapB = cb > 0 ? make_map(srcB, nb * p.planes_b * Cfg::P, sz, p.pitch, p.TY + 2, p.planes_b * Cfg::P, t.wide) : mapA;
  int grid = nb * p.tiles_x * p.tiles_y;
  if (t.T) {  // persistent CTAs striding over (column, z block) work items
    p.total_items = grid * ceil_div(sz.z, t.T);
    grid = std::min<int>(p.total_items, sm_count());
  }
  auto run = [&](auto kern, int threads = kThreads) {
    CFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, s>>>(mapA, mapB, p);
    CFB_LAUNCH_CHECK();
#ifdef CFB_TS_TRACE
    if (p.trace && t.shift && g_trace_left > 0) {
      --g_trace_left;
      std::vector<long long> h(256 * 16);
      cudaStreamSynchronize(s);
      cudaMemcpy(h.data(), p.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double a[16] = {0};
      for (int c = 0; c < grid; ++c)
        for (int i = 0; i < 16; ++i) a[i] += (double)h[c * 16 + i] / grid;
      fprintf(stderr, "[cfb-trace] %d->%d %dx%dx%d nb=%d T=%d XT=%d TY=%d G=%d items/cta=%.1f | mma: total %.0f wait accE %.0f TF %.0f B %.0f issue %.0f (issuer 1: total %.0f issue %.0f) | loader: total %.0f "
              "wait plane %.0f TE %.0f | epi: total %.0f wait accF %.0f | prodA: total %.0f wait empty %.0f  (cycles, CTA average)\n",
              CIN, COUT, sz.z, sz.y, sz.x, nb, t.T, t.XT, t.TY, p.G, (double)p.total_items / grid, a[0], a[1], a[2], a[3], a[11], a[12], a[13], a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
    }
#endif
  };
  if (tail) {
    if constexpr (CIN == 16 && COUT == 16) {
      p.tail = *tail;
      if (t.shift) run(conv3_ts_umma_kernel<CIN, COUT, SPLIT, true, F8>, kThreadsTSTail);
      else if (t.T) run(conv3_zs_umma_kernel<CIN, COUT, SPLIT, true>);
      else run(conv3_umma_kernel<CIN, COUT, SPLIT, true>);
      return;
    } else {
      throw std::runtime_error("the fused head+blend tail needs a 16->16 layer");
    }
  }
  if (t.shift) {
    if (pool_out) {  // (1,2,2) max pooling fused into the epilogue (the two encoder layers that feed a pool)
      if constexpr ((CIN == 16 && COUT == 16) || (CIN == 32 && COUT == 32)) run(conv3_ts_umma_kernel<CIN, COUT, SPLIT, false, F8, true>, kThreadsTS);
      else throw std::runtime_error("fused pooling exists for the 16->16 and 32->32 layers");
      return;
    }
    run(conv3_ts_umma_kernel<CIN, COUT, SPLIT, false, F8>, kThreadsTS);
    return;
  }
  struct PoolAfter {  // the other kernel variants pool with the stand-alone kernel
    __half* out; __half* pool; int cout, fmt, nb; Int3 sz; cudaStream_t s;
    ~PoolAfter() noexcept(false) { if (pool && !std::uncaught_exceptions()) launch_maxpool_cp8(out, pool, cout, fmt, nb, sz, s); }
  } pool_after{out, pool_out, COUT, w.fmt ? w.fmt : (SPLIT ? kFmtF16x2 : kFmtF16), nb, sz, s};
  if (t.T) {
    if constexpr (3 * Cfg::NB <= 256) {
      run(conv3_zs_umma_kernel<CIN, COUT, SPLIT, false>);
      return;
    } else {
      throw std::runtime_error("z-stacked kernel needs 3 * NB <= 256");
    }
  }
  run(conv3_umma_kernel<CIN, COUT, SPLIT, false>);
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    CFB_CUDA(cudaGetDevice(&dev));
    CFB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// First launch of a (layer, size, batch): time the most promising tilings on the real buffers (every
// tiling computes bit-identical results) and cache the winner in the layer's PackedConv.
template <int CIN, int COUT, bool SPLIT, bool F8 = false>
void launch_cfg(const __half* srcA, int ca, const __half* srcB, int cb, const PackedConv& w, __half* out, int nb,
                Int3 sz, bool relu, cudaStream_t s, const FusedTail* tail, __half* pool_out = nullptr) {
  const uint64_t key = ((uint64_t)sz.z << 48) ^ ((uint64_t)sz.y << 32) ^ ((uint64_t)sz.x << 16) ^ (uint64_t)nb;
  auto it = w.tuned->find(key);
  if (it == w.tuned->end()) {
    std::vector<ConvTile> cands = enumerate_tiles<CIN, COUT, SPLIT>(nb, sz, sm_count(), F8, pool_out != nullptr);
    if (cands.empty()) throw std::runtime_error("conv3_umma: no tile configuration fits shared memory / TMEM");
    ConvTile best = cands[0];
    const int64_t work = (int64_t)nb * vol(sz);
    const bool tune = !getenv("CFB_NO_AUTOTUNE") && work >= (1 << 18) && cands.size() > 1;
    if (tune) {
      cudaEvent_t e0, e1;
      CFB_CUDA(cudaEventCreate(&e0));
      CFB_CUDA(cudaEventCreate(&e1));
      float best_ms = 1e30f;
      // the most promising tilings of each kernel variant: per-tap, z-stacked, and z-stacked + TMEM shift with
      // 1, 2 and >= 3 M tiles per group (= MMA-issuing threads; the static model ranks these poorly against each other)
      std::vector<ConvTile> pick;
      const int quota[5] = {8, 12, 10, 12, 12};
      for (int cls = 0; cls < 5; ++cls) {
        int taken = 0;
        for (const ConvTile& c : cands) {
          const int g = ceil_div(c.TY * (c.XT + 2), 120);
          const int k = c.shift ? (g == 1 ? 2 : (g == 2 ? 3 : 4)) : (c.T ? 1 : 0);
          if (k == cls && taken < quota[cls]) { pick.push_back(c); ++taken; }
        }
      }
      cands.swap(pick);
      const size_t n = cands.size();
      for (size_t i = 0; i < n; ++i) {
        launch_tile<CIN, COUT, SPLIT, F8>(cands[i], srcA, ca, srcB, cb, w, out, nb, sz, relu, s, nullptr, pool_out);  // warm
        float ms = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {  // best of two: one sample per candidate picked a 5-10 % slower tiling now and then
          CFB_CUDA(cudaEventRecord(e0, s));
          launch_tile<CIN, COUT, SPLIT, F8>(cands[i], srcA, ca, srcB, cb, w, out, nb, sz, relu, s, nullptr, pool_out);
          CFB_CUDA(cudaEventRecord(e1, s));
          CFB_CUDA(cudaEventSynchronize(e1));
          float t = 0.f;
          CFB_CUDA(cudaEventElapsedTime(&t, e0, e1));
          ms = std::min(ms, t);
        }
        if (ms < best_ms) { best_ms = ms; best = cands[i]; }
        if (getenv("CFB_DEBUG_TUNE"))
          fprintf(stderr, "[cfb-tune] %d->%d %dx%dx%d nb=%d: T=%d%s%s XT=%d TY=%d G=%d bstages=%d resident=%d  %.3f ms\n", CIN, COUT, sz.z, sz.y, sz.x, nb,
                  cands[i].T, cands[i].shift ? "+shift" : "", cands[i].ring == 2 ? " ring2" : "", cands[i].XT, cands[i].TY, ceil_div(cands[i].TY * (cands[i].XT + 2), cands[i].shift ? 120 : 128),
                  cands[i].bstages, (int)cands[i].resident, ms);
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      best.cost = best_ms;
    }
    if (getenv("CFB_DEBUG_CFG"))
      fprintf(stderr, "[cfb] conv3 %d->%d split=%d size=%dx%dx%d nb=%d: zstack T=%d%s XT=%d%s TY=%d bstages=%d resident=%d (%s %.3f)\n",
              CIN, COUT, (int)SPLIT, sz.z, sz.y, sz.x, nb, best.T, best.shift ? (best.ring == 2 ? "+shift(ring2)" : "+shift") : "", best.XT, best.wide ? "w" : "", best.TY, best.bstages,
              (int)best.resident, tune ? "tuned ms" : "model cost", best.cost);
    it = w.tuned->emplace(key, best).first;
  }
#ifdef CFB_TS_TRACE
  g_trace_print = getenv("CFB_TS_TRACE_PRINT") != nullptr;
#endif
  launch_tile<CIN, COUT, SPLIT, F8>(it->second, srcA, ca, srcB, cb, w, out, nb, sz, relu, s, tail, pool_out);
#ifdef CFB_TS_TRACE
  g_trace_print = false;
#endif
}


template <int CIN, int COUT, bool SPLIT, bool F8 = false>
void launch_convT_cfg(const __half* in, const PackedConv& w, __half* out, int nb, Int3 sz, cudaStream_t s) {
  constexpr int P = SPLIT ? 2 : 1;
  constexpr int NPL = P * CIN / 8, N1 = 4 * P * COUT, WBYTES = N1 * CIN * 2;
  UmmaConvTParams p{};
  p.Z = sz.z; p.Y = sz.y; p.X = sz.x;
  const int maxg = kBufCols / N1;
  // tile: XT = X (<= 128, one contiguous TMA row), TY rows so that TY * XT = G * 128 positions
  p.XT = std::min(sz.x + (sz.x & 1), 128);
  p.TY = std::max(1, (maxg * 128) / p.XT);
  p.TY = std::min(p.TY, sz.y);
  p.G = ceil_div(p.TY * p.XT, 128);
  while (p.G > maxg && p.TY > 1) { --p.TY; p.G = ceil_div(p.TY * p.XT, 128); }
  if (p.G > maxg) throw std::runtime_error("convT_umma: tile does not fit TMEM");
  p.tiles_x = ceil_div(sz.x, p.XT);
  p.tiles_y = ceil_div(sz.y, p.TY);
  p.planes = CIN / 8;
  p.plane_stride = (uint32_t)(p.TY * p.XT * 16);
  p.slot_stride = (uint32_t)((NPL * (size_t)p.plane_stride + 127) / 128 * 128);
  p.wpacked = w.w; p.bias = w.bias; p.out = out; p.acc_scale = w.acc_scale;
  p.ring = kRing;
  auto smem_for = [&](int ring) { return (size_t)ring * p.slot_stride + WBYTES + 128 + kTailPad + 128; };
  const size_t smem = smem_for(p.ring);
  if (smem > (size_t)kMaxSmem) throw std::runtime_error("convT_umma: shared memory exceeded");
  const CUtensorMap mapA = make_map(in, nb * p.planes * P, sz, p.XT, p.TY, p.planes * P, /*wide=*/false);
  auto kern = convT_umma_kernel<CIN, COUT, SPLIT, F8>;
  CFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nb * p.tiles_x * p.tiles_y, kThreadsConvT, smem, s>>>(mapA, p);
  CFB_LAUNCH_CHECK();
}

template <bool SPLIT, bool F8 = false>
void dispatch(const __half* srcA, int ca, const __half* srcB, int cb, const PackedConv& w, __half* out, int nb, Int3 sz,
              bool relu, cudaStream_t s, const FusedTail* tail, __half* pool_out) {
  const int cin = ca + cb, cout = w.cout;
#define CFB_CASE(CI, CO) \
  if (cin == CI && cout == CO) return launch_cfg<CI, CO, SPLIT, F8>(srcA, ca, srcB, cb, w, out, nb, sz, relu, s, tail, pool_out);
  CFB_CASE(16, 16) CFB_CASE(16, 32) CFB_CASE(32, 32) CFB_CASE(32, 64) CFB_CASE(64, 64) CFB_CASE(64, 32) CFB_CASE(32, 16)
#undef CFB_CASE
  throw std::runtime_error("conv3_umma: unsupported channel configuration " + std::to_string(cin) + "->" + std::to_string(cout));
}

}  // namespace

void launch_conv3_umma(const __half* srcA, int ca, const __half* srcB, int cb, const PackedConv& w, __half* out, int nb,
                       Int3 sz, bool relu, cudaStream_t s, const ConvTail* tail, __half* pool_out) {
  if (pool_out && (tail || (sz.y & 1) || (sz.x & 1))) throw std::runtime_error("conv3_umma: fused pooling needs even y, x and no fused tail");
  if (ca % 16 || (cb % 16) || w.cin != ca + cb) throw std::runtime_error("conv3_umma: channel mismatch");
  FusedTail ft{};
  if (tail) {
    if (tail->channels > 8) throw std::runtime_error("fused tail: at most 8 channels");
    ft.head_w = tail->head_w; ft.head_b = tail->head_b; ft.patches = tail->patches; ft.mask = tail->mask; ft.out = tail->out;
    ft.channels = tail->channels; ft.op = tail->out_patch; ft.crop = tail->crop; ft.os = tail->out_size;
    ft.scale = tail->scale;
  }
  if (w.fmt == kFmtF16F8) dispatch<true, true>(srcA, ca, srcB, cb, w, out, nb, sz, relu, s, tail ? &ft : nullptr, pool_out);
  else if (w.parts == 2) dispatch<true>(srcA, ca, srcB, cb, w, out, nb, sz, relu, s, tail ? &ft : nullptr, pool_out);
  else dispatch<false>(srcA, ca, srcB, cb, w, out, nb, sz, relu, s, tail ? &ft : nullptr, pool_out);
}

// ------------------------------------------------------------------------------------------
// Weight packing (host)
// ------------------------------------------------------------------------------------------
// power of two >= the largest |w| of a layer (f16f8 mode: beta = 2^14 / wmax', act_format.cuh)
static float weight_beta(const float* w, size_t n) {
  float wmax = 0.f;
  for (size_t i = 0; i < n; ++i) wmax = std::max(wmax, std::fabs(w[i]));
  if (!(wmax > 0.f) || !std::isfinite(wmax)) return 16384.0f;
  return 16384.0f / std::exp2(std::ceil(std::log2(wmax)));
}
static uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

void pack_conv3_weights(const float* h_w, const float* h_bias, int cin, int cout, int fmt, PackedConv& out) {
  free_packed(out);
  const int parts = fmt_planes(fmt);
  const int KB = cin >= 32 ? 32 : 16, KG = cin / KB, NB = parts * cout;
  const size_t block = (size_t)NB * KB;  // halves
  std::vector<__half> buf((size_t)27 * KG * block);
  for (int t = 0; t < 27; ++t)
    for (int g = 0; g < KG; ++g)
      for (int kc = 0; kc < KB / 8; ++kc)
        for (int n = 0; n < NB; ++n)
          for (int e = 0; e < 8; ++e) {
            const int ci = g * KB + kc * 8 + e;
            const int co = n % cout;
            const float wv = h_w[((size_t)co * cin + ci) * 27 + t];
            const __half hi = __float2half_rn(wv);
            const __half val = n < cout ? hi : __float2half_rn(wv - __half2float(hi));
            buf[((size_t)t * KG + g) * block + ((size_t)kc * NB + n) * 8 + e] = val;
          }
  out.cin = cin; out.cout = cout; out.parts = parts; out.fmt = fmt;
  out.tuned = std::make_shared<std::map<uint64_t, ConvTile>>();
  out.bytes = buf.size() * sizeof(__half);
  CFB_CUDA(cudaMalloc(&out.w, out.bytes));
  CFB_CUDA(cudaMemcpy(out.w, buf.data(), out.bytes, cudaMemcpyHostToDevice));
  CFB_CUDA(cudaMalloc(&out.bias, cout * sizeof(float)));
  CFB_CUDA(cudaMemcpy(out.bias, h_bias, cout * sizeof(float), cudaMemcpyHostToDevice));
  // z-stacked blocks: (dy, dx, kg): [KB/8][3*NB rows: dz = 2 | 1 | 0][8]
  std::vector<__half> zs((size_t)27 * KG * block);
  for (int t9 = 0; t9 < 9; ++t9)
    for (int g = 0; g < KG; ++g)
      for (int kc = 0; kc < KB / 8; ++kc)
        for (int zi = 0; zi < 3; ++zi)
          for (int n = 0; n < NB; ++n)
            for (int e = 0; e < 8; ++e) {
              const int dz = 2 - zi, t = dz * 9 + t9;
              const int ci = g * KB + kc * 8 + e;
              const int co = n % cout;
              const float wv = h_w[((size_t)co * cin + ci) * 27 + t];
              const __half hi = __float2half_rn(wv);
              const __half val = n < cout ? hi : __float2half_rn(wv - __half2float(hi));
              zs[((size_t)t9 * KG + g) * (3 * block) + ((size_t)kc * 3 * NB + (size_t)zi * NB + n) * 8 + e] = val;
            }
  CFB_CUDA(cudaMalloc(&out.w_zs, zs.size() * sizeof(__half)));
  CFB_CUDA(cudaMemcpy(out.w_zs, zs.data(), zs.size() * sizeof(__half), cudaMemcpyHostToDevice));
  // TMEM-shift kernel: blocks ordered (dy, kg, dx); rows of a block = [hi: dz 2,1,0 | lo: dz 2,1,0] x cout
  std::vector<__half> ts(zs.size());
  for (int dy = 0; dy < 3; ++dy)
    for (int g = 0; g < KG; ++g)
      for (int dx = 0; dx < 3; ++dx)
        for (int kc = 0; kc < KB / 8; ++kc)
          for (int part = 0; part < parts; ++part)
            for (int zi = 0; zi < 3; ++zi)
              for (int co = 0; co < cout; ++co)
                for (int e = 0; e < 8; ++e) {
                  const int t = (2 - zi) * 9 + dy * 3 + dx;
                  const int ci = g * KB + kc * 8 + e;
                  const float wv = h_w[((size_t)co * cin + ci) * 27 + t];
                  const __half hi = __float2half_rn(wv);
                  const size_t row = (size_t)part * 3 * cout + (size_t)zi * cout + co;
                  ts[(((size_t)dy * KG + g) * 3 + dx) * (3 * block) + ((size_t)kc * 3 * NB + row) * 8 + e] =
                      part == 0 ? hi : __float2half_rn(wv - __half2float(hi));
                }
  if (fmt == kFmtF16F8) {
    // f16f8 (act_format.cuh): part-0 rows = WH = fp16(w beta); part-1 rows hold e4m3 bytes, K step ks = 16-byte chunks
    // (2 ks) = WL8 = e4m3((w beta - WH) mu) and (2 ks + 1) = W8 = e4m3(w delta) of the SAME 16 channels
    const float beta = weight_beta(h_w, (size_t)cout * cin * 27), delta = beta / kActLambda;
    out.acc_scale = 1.0f / (kActAlpha * beta);
    uint8_t* tb = reinterpret_cast<uint8_t*>(ts.data());
    for (int dy = 0; dy < 3; ++dy)
      for (int g = 0; g < KG; ++g)
        for (int dx = 0; dx < 3; ++dx)
          for (int kc = 0; kc < KB / 8; ++kc)
            for (int zi = 0; zi < 3; ++zi)
              for (int co = 0; co < cout; ++co) {
                const int t = (2 - zi) * 9 + dy * 3 + dx;
                const size_t blk0 = (((size_t)dy * KG + g) * 3 + dx) * (3 * block);
                const size_t row_hi = (size_t)zi * cout + co, row_p1 = (size_t)3 * cout + row_hi;
                for (int e = 0; e < 8; ++e) {
                  const int ci = g * KB + kc * 8 + e;
                  ts[blk0 + ((size_t)kc * 3 * NB + row_hi) * 8 + e] = __float2half_rn(h_w[((size_t)co * cin + ci) * 27 + t] * beta);
                }
                uint8_t* dst = tb + 2 * (blk0 + ((size_t)kc * 3 * NB + row_p1) * 8);
                for (int j = 0; j < 16; ++j) {
                  const int ci = g * KB + (kc / 2) * 16 + j;
                  const float wb = h_w[((size_t)co * cin + ci) * 27 + t] * beta;
                  const float wl = wb - __half2float(__float2half_rn(wb));
                  dst[j] = (kc & 1) ? to_e4m3(h_w[((size_t)co * cin + ci) * 27 + t] * delta) : to_e4m3(wl * kWgtMu);
                }
              }
  }
  CFB_CUDA(cudaMalloc(&out.w_ts, ts.size() * sizeof(__half)));
  CFB_CUDA(cudaMemcpy(out.w_ts, ts.data(), ts.size() * sizeof(__half), cudaMemcpyHostToDevice));
}

void launch_convT_umma(const __half* in, const PackedConv& w, __half* out, int nb, Int3 in_size, cudaStream_t s) {
  const bool split = w.parts == 2;
#define CFB_CASE(CI, CO)                                                                     \
  if (w.cin == CI && w.cout == CO) {                                                         \
    if (w.fmt == kFmtF16F8) return launch_convT_cfg<CI, CO, true, true>(in, w, out, nb, in_size, s); \
    if (split) return launch_convT_cfg<CI, CO, true>(in, w, out, nb, in_size, s);            \
    return launch_convT_cfg<CI, CO, false>(in, w, out, nb, in_size, s);                      \
  }
  CFB_CASE(64, 32) CFB_CASE(32, 16)
#undef CFB_CASE
  throw std::runtime_error("convT_umma: unsupported channel configuration");
}

void pack_convT_weights(const float* h_w, const float* h_bias, int cin, int cout, int fmt, PackedConv& out) {
  free_packed(out);
  const int parts = fmt_planes(fmt);
  const int N1 = 4 * parts * cout, N2 = 4 * cout;
  std::vector<__half> buf((size_t)N1 * cin);
  for (int kc = 0; kc < cin / 8; ++kc)
    for (int n = 0; n < N1; ++n)
      for (int e = 0; e < 8; ++e) {
        const int ci = kc * 8 + e;
        const int nn = n % N2, t = nn / cout, co = nn % cout;
        const float wv = h_w[((size_t)ci * cout + co) * 4 + t];  // (cin, cout, 1, 2, 2)
        const __half hi = __float2half_rn(wv);
        buf[((size_t)kc * N1 + n) * 8 + e] = n < N2 ? hi : __float2half_rn(wv - __half2float(hi));
      }
  out.fmt = fmt;
  if (fmt == kFmtF16F8) {  // rows 0..N2: WH; rows N2..2 N2: e4m3 bytes, chunk 2 ks = WL8, chunk 2 ks + 1 = W8 (16 channels of K step ks)
    const float beta = weight_beta(h_w, (size_t)cin * cout * 4), delta = beta / kActLambda;
    out.acc_scale = 1.0f / (kActAlpha * beta);
    uint8_t* bb = reinterpret_cast<uint8_t*>(buf.data());
    for (int kc = 0; kc < cin / 8; ++kc)
      for (int n = 0; n < N2; ++n) {
        const int t = n / cout, co = n % cout;
        for (int e = 0; e < 8; ++e)
          buf[((size_t)kc * N1 + n) * 8 + e] = __float2half_rn(h_w[((size_t)(kc * 8 + e) * cout + co) * 4 + t] * beta);
        uint8_t* dst = bb + 2 * (((size_t)kc * N1 + N2 + n) * 8);
        for (int j = 0; j < 16; ++j) {
          const int ci = (kc / 2) * 16 + j;
          const float w0 = h_w[((size_t)ci * cout + co) * 4 + t], wb = w0 * beta;
          dst[j] = (kc & 1) ? to_e4m3(w0 * delta) : to_e4m3((wb - __half2float(__float2half_rn(wb))) * kWgtMu);
        }
      }
  }
  out.cin = cin; out.cout = cout; out.parts = parts;
  out.tuned = std::make_shared<std::map<uint64_t, ConvTile>>();
  out.bytes = buf.size() * sizeof(__half);
  CFB_CUDA(cudaMalloc(&out.w, out.bytes));
  CFB_CUDA(cudaMemcpy(out.w, buf.data(), out.bytes, cudaMemcpyHostToDevice));
  CFB_CUDA(cudaMalloc(&out.bias, cout * sizeof(float)));
  CFB_CUDA(cudaMemcpy(out.bias, h_bias, cout * sizeof(float), cudaMemcpyHostToDevice));
}

void pack_first_conv_weights(const float* h_w, const float* h_bias, int fmt, PackedConv& out) {
  free_packed(out);
  const int parts = fmt_planes(fmt);
  const int NBR = 16 * parts;
  std::vector<__half> buf((size_t)4 * NBR * 8, __float2half_rn(0.f));
  for (int c = 0; c < 4; ++c)
    for (int n = 0; n < NBR; ++n)
      for (int e = 0; e < 8; ++e) {
        const int tap = c * 8 + e;
        if (tap >= 27) continue;
        const float wv = h_w[(size_t)(n % 16) * 27 + tap];  // (16, 1, 3, 3, 3)
        const __half hi = __float2half_rn(wv);
        buf[((size_t)c * NBR + n) * 8 + e] = n < 16 ? hi : __float2half_rn(wv - __half2float(hi));
      }
  out.cin = 1; out.cout = 16; out.parts = parts;
  out.tuned = std::make_shared<std::map<uint64_t, ConvTile>>();
  out.bytes = buf.size() * sizeof(__half);
  CFB_CUDA(cudaMalloc(&out.w, out.bytes));
  CFB_CUDA(cudaMemcpy(out.w, buf.data(), out.bytes, cudaMemcpyHostToDevice));
  CFB_CUDA(cudaMalloc(&out.bias, 16 * sizeof(float)));
  CFB_CUDA(cudaMemcpy(out.bias, h_bias, 16 * sizeof(float), cudaMemcpyHostToDevice));
}

void launch_first_conv_umma(const void* chunk_u8, Int3 cs, const PatchPos* patches, int nb, Int3 ps, const PackedConv& w,
                            __half* out, cudaStream_t s) {
  FirstConvParams p{};
  p.chunk = static_cast<const uint8_t*>(chunk_u8); p.cs = cs; p.patches = patches; p.ps = ps;
  p.wpacked = w.w; p.bias = w.bias; p.out = out;
  p.tiles_x = ceil_div(ps.x, 128);
  p.total_tiles = nb * ps.z * ps.y * p.tiles_x;
  const int grid = std::min<int>(p.total_tiles, 2 * sm_count());
  if (w.parts == 2) first_conv_umma_kernel<true><<<grid, kFcThreads, 0, s>>>(p);
  else first_conv_umma_kernel<false><<<grid, kFcThreads, 0, s>>>(p);
  CFB_LAUNCH_CHECK();
}

void free_packed(PackedConv& p) {
  if (p.w) cudaFree(p.w);
  if (p.w_zs) cudaFree(p.w_zs);
  if (p.w_ts) cudaFree(p.w_ts);
  if (p.bias) cudaFree(p.bias);
  p = PackedConv{};
}

}  // namespace cfb
