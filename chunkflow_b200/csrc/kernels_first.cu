// First layer of the U-Net (1 -> 16 channels, 3x3x3) on tcgen05, fused with patch extraction from the uint8 chunk
// (reference: Chunk.cutout chunk/base.py:761-781, the `/255` of inferencer.py:395-399 and the model's first Conv3d).
//
// The CUDA-core version (kernels_cp8.cu) issues 432 FFMA per voxel and is FFMA-issue bound (93 ms per 1024^3 chunk);
// here the 27 taps are the K dimension of a tensor-core product and the kernel is bound by writing the 16-channel
// activation (64 B per voxel) instead.  Same scheme as the TMEM-shift kernel of the other layers (kernels_umma.cu):
//
//   * a persistent CTA walks (patch, y tile, x tile) columns; 3 producer warps stream the uint8 halo planes
//     q = -1 .. Z of the column into a ring of shared-memory slots (zero outside the PATCH = SAME padding; the
//     test-time-augmentation coordinate map is applied here),
//   * 4 loader warps build, per output plane and M tile of 120 positions, the A tile in TENSOR MEMORY: row = position,
//     K = 16 fp16 = the 9 (dz, dy) taps of one dx column (uint8 values are exact in fp16) + 7 zeros (tcgen05.st),
//   * one thread issues  MMA(dx=0) ; tcgen05.shift.down ; MMA(dx=1) ; tcgen05.shift.down ; MMA(dx=2)  per tile, each MMA
//     twice (weights hi and lo: fp16(w) and fp16(w - fp16(w)), so the products carry the fp32 weight), N = 16, fp32
//     accumulation in TMEM,
//   * 8 epilogue warps: tcgen05.ld, x 1/255, + bias, ReLU, encode to the activation format of the precision mode
//     (act_format.cuh), 16-byte stores.
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "act_format.cuh"
#include "chunkflow_b200.h"
#include "kernels_umma.cuh"

namespace cfb {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
    if (clock64() - t0 > 4000000000ll) {
      printf("chunkflow_b200 first_conv: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
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
__device__ __forceinline__ void tc_st8(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %6, %6};" ::"r"(taddr), "r"(a), "r"(b), "r"(c),
               "r"(d), "r"(e), "r"(0u)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kXT = 32, kTY = 14, kPitch = kXT + 2;  // tile of 14 x 32 outputs; raster pitch incl. the two halo columns
constexpr int kG = 4;                                // M tiles of 120 positions (4 x 30) per plane: 14 * 34 = 476 <= 480
constexpr int kRows = kTY + 3;                       // rows of a plane slot: 2 halo rows + 1 for the positions past TY * pitch
constexpr int kSlotBytes = 592;                      // >= kRows * kPitch = 578
constexpr int kRing = 6;                             // uint8 plane slots
constexpr int kGroups = 4;                           // A-tile groups (kG tiles x 8 columns) in tensor memory
constexpr int kAccs = 4;                             // accumulator buffers (kG tiles x 16 columns)
constexpr int kACol0 = 256;                          // accumulators at [0, 256), A groups at [256, 384)
constexpr int kThreads = 512;                        // warps 0-3 loaders, 4-11 epilogue, 12 MMA issuer, 13-15 producers
constexpr int kProducers = 3;
constexpr int kWBytes = 6 * 512;                     // weight blocks (dx, hi | lo): [2 K chunks][16 rows][8 halves]

struct FirstTsParams {
  const uint8_t* chunk;
  Int3 cs;
  const PatchPos* patches;
  Int3 ps;
  const __half* wblocks;  // 6 blocks of 512 bytes
  const float* bias;
  __half* out;
  int fmt;                // ActFmt of the output
  int tiles_x, tiles_y, total_items;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 16 channels of one voxel -> the records of the output format FMT; pl[k] = base of plane k of this patch's two 8-channel
// chunks (hoisted out of the voxel loop: the 64-bit plane products were a third of the epilogue's instructions)
template <int FMT>
__device__ __forceinline__ void store_voxel16(const float (&v)[16], uint4* const (&pl)[4], size_t vox) {
  if constexpr (FMT == kFmtF16F8) {
    uint4 h0, h1, a8, l8;
    af_encode16(v, h0, h1, a8, l8);
    pl[0][vox] = h0;
    pl[1][vox] = a8;
    pl[2][vox] = h1;
    pl[3][vox] = l8;
  } else {
    constexpr int P = FMT == kFmtF16x2 ? 2 : 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float hi[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hi[i] = __half2float(__float2half_rn(v[h * 8 + i]));
      pl[h * P][vox] = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
      if constexpr (P == 2)
        pl[h * P + 1][vox] = make_uint4(pack_h2(v[h * 8] - hi[0], v[h * 8 + 1] - hi[1]), pack_h2(v[h * 8 + 2] - hi[2], v[h * 8 + 3] - hi[3]),
                                        pack_h2(v[h * 8 + 4] - hi[4], v[h * 8 + 5] - hi[5]), pack_h2(v[h * 8 + 6] - hi[6], v[h * 8 + 7] - hi[7]));
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) first_conv_ts_kernel(const FirstTsParams p) {
  __shared__ __align__(128) uint8_t s_planes[kRing * kSlotBytes];
  __shared__ __align__(128) uint8_t s_w[kWBytes];
  __shared__ __align__(8) uint64_t s_bars[2 * kRing + 2 * kGroups + 2 * kAccs];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[16];
  const uint32_t bar0 = smem_u32(s_bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int kFull = 0, kEmpty = kRing, kTF = 2 * kRing, kTE = kTF + kGroups, kAccF = kTE + kGroups, kAccE = kAccF + kAccs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wblocks)[i];
  if (threadIdx.x < 16) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(BAR(kFull + i), 32); mbar_init(BAR(kEmpty + i), 4); }
    for (int i = 0; i < kGroups; ++i) { mbar_init(BAR(kTF + i), 128); mbar_init(BAR(kTE + i), 1); }
    for (int i = 0; i < kAccs; ++i) { mbar_init(BAR(kAccF + i), 1); mbar_init(BAR(kAccE + i), 256); }
    fence_barrier_init();
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();  // the weight blocks were written by generic-proxy stores, the tensor core reads them through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int Z = p.ps.z, Y = p.ps.y, X = p.ps.x;
  const int ncols = p.tiles_x * p.tiles_y;

  if (warp >= 13) {
    // ---------------- producers: plane n of the CTA's plane sequence is filled by producer n % 3 ----------------
    const int me = warp - 13;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int b = item / ncols, col = item % ncols;
      const int x0 = (col % p.tiles_x) * kXT, y0 = (col / p.tiles_x) * kTY;
      const PatchPos pp = p.patches[b];
      for (int q = -1; q <= Z; ++q, ++n) {
        if ((int)(n % kProducers) != me) continue;
        const uint32_t slot = n % kRing, use = n / kRing;
        uint8_t vals[19];
#pragma unroll
        for (int i = 0; i < 19; ++i) {
          const int idx = lane + 32 * i;
          const int r = idx / kPitch, c = idx - r * kPitch;
          const int gy = y0 + r - 1, gx = x0 + c - 1;
          uint8_t v = 0;  // zero outside the PATCH: SAME padding at the patch border
          if (idx < kRows * kPitch && q >= 0 && q < Z && gy >= 0 && gy < Y && gx >= 0 && gx < X) {
            int sy = gy, sx = gx;
            if (pp.flags & 7) tta_map(pp.flags, Y, X, gy, gx, sy, sx);  // augmented variant reads the original patch
            v = __ldg(p.chunk + ((size_t)(pp.iz + q) * p.cs.y + (pp.iy + sy)) * p.cs.x + pp.ix + sx);
          }
          vals[i] = v;
        }
        if (use > 0) mbar_wait(BAR(kEmpty + slot), (use - 1) & 1u);
        uint8_t* dst = s_planes + slot * kSlotBytes;
#pragma unroll
        for (int i = 0; i < 19; ++i) {
          const int idx = lane + 32 * i;
          if (idx < kRows * kPitch) dst[idx] = vals[i];
        }
        mbar_arrive(BAR(kFull + slot));  // release: every lane's stores are visible to the loaders that acquire the phase
      }
    }
  } else if (warp < 4) {
    // ---------------- loaders: uint8 planes -> fp16 A tiles (row = position, K = 9 (dz, dy) taps + 7 zeros) in TMEM ----------------
    const int wq = warp;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    int base[kG];  // offset of this lane's position inside a plane slot, per M tile
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int pos = g * 120 + 30 * wq + lane;
      base[g] = pos;  // (row * pitch + col) == pos: the slot IS the raster
    }
    uint32_t n0 = 0, gi = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, n0 += (uint32_t)Z + 2) {
      for (int pl = 0; pl < Z; ++pl, ++gi) {
        // planes n0 + pl, + 1, + 2 (q = pl - 1, pl, pl + 1); earlier ones were awaited by the previous iteration
        for (int j = pl == 0 ? 0 : 2; j < 3; ++j) {
          const uint32_t n = n0 + (uint32_t)(pl + j);
          mbar_wait(BAR(kFull + n % kRing), (n / kRing) & 1u);
        }
        const uint8_t* s0 = s_planes + ((n0 + pl) % kRing) * kSlotBytes;
        const uint8_t* s1 = s_planes + ((n0 + pl + 1) % kRing) * kSlotBytes;
        const uint8_t* s2 = s_planes + ((n0 + pl + 2) % kRing) * kSlotBytes;
        uint32_t w[kG][5];
#pragma unroll
        for (int g = 0; g < kG; ++g) {
          const int o = base[g];
          float t[9];
          t[0] = (float)s0[o]; t[1] = (float)s0[o + kPitch]; t[2] = (float)s0[o + 2 * kPitch];
          t[3] = (float)s1[o]; t[4] = (float)s1[o + kPitch]; t[5] = (float)s1[o + 2 * kPitch];
          t[6] = (float)s2[o]; t[7] = (float)s2[o + kPitch]; t[8] = (float)s2[o + 2 * kPitch];
          w[g][0] = pack_h2(t[0], t[1]); w[g][1] = pack_h2(t[2], t[3]); w[g][2] = pack_h2(t[4], t[5]); w[g][3] = pack_h2(t[6], t[7]);
          w[g][4] = pack_h2(t[8], 0.f);
        }
        const uint32_t grp = gi % kGroups, use = gi / kGroups;
        if (use > 0) mbar_wait(BAR(kTE + grp), (use - 1) & 1u);
        tc_fence_after();
        const uint32_t t0 = tmem_base + lane_base + kACol0 + grp * (kG * 8);
#pragma unroll
        for (int g = 0; g < kG; ++g) tc_st8(t0 + g * 8, w[g][0], w[g][1], w[g][2], w[g][3], w[g][4]);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(BAR(kTF + grp));
        // plane q = pl - 1 is dead now; after the last output plane the two remaining planes of the column are, too
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(kEmpty + (n0 + pl) % kRing));
          if (pl == Z - 1) { mbar_arrive(BAR(kEmpty + (n0 + pl + 1) % kRing)); mbar_arrive(BAR(kEmpty + (n0 + pl + 2) % kRing)); }
        }
      }
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t DESC_HI = 8u | (1u << 14);  // SBO = 128 B, descriptor version 1
      constexpr uint32_t IDESC = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);  // D = f32, A = B = f16, N = 16, M = 128
      const uint32_t w16 = smem_u32(s_w) >> 4;
      auto bdesc = [&](int blk) { return ((uint64_t)DESC_HI << 32) | ((16u << 16) | (w16 + (uint32_t)blk * 32u)); };  // LBO = 256 B
      uint32_t gi = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        for (int pl = 0; pl < Z; ++pl, ++gi) {
          const uint32_t grp = gi % kGroups, ab = gi % kAccs;
          if (gi >= kAccs) mbar_wait(BAR(kAccE + ab), ((gi / kAccs) - 1) & 1u);
          mbar_wait(BAR(kTF + grp), (gi / kGroups) & 1u);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < kG; ++g) {
            const uint32_t a = tmem_base + kACol0 + grp * (kG * 8) + g * 8, d = tmem_base + ab * (kG * 16) + g * 16;
            tc_mma_f16_ta(d, a, bdesc(0), IDESC, 0u);
            tc_mma_f16_ta(d, a, bdesc(1), IDESC, 1u);
            tc_shift_down(a);
            tc_mma_f16_ta(d, a, bdesc(2), IDESC, 1u);
            tc_mma_f16_ta(d, a, bdesc(3), IDESC, 1u);
            tc_shift_down(a);
            tc_mma_f16_ta(d, a, bdesc(4), IDESC, 1u);
            tc_mma_f16_ta(d, a, bdesc(5), IDESC, 1u);
          }
          tc_commit(BAR(kTE + grp));
          tc_commit(BAR(kAccF + ab));
        }
      }
    }
  } else {
    // ---------------- epilogue: two sets of four warps take alternate M tiles ----------------
    const int wq = warp & 3, eset = warp >= 8 ? 1 : 0;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const size_t plane_vox = (size_t)Z * Y * X;
    uint4* out16 = reinterpret_cast<uint4*>(p.out);
    float bias[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bias[i] = s_bias[i];
    uint32_t gi = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int b = item / ncols, col = item % ncols;
      const int x0 = (col % p.tiles_x) * kXT, y0 = (col / p.tiles_x) * kTY;
      constexpr int NPL = FMT == kFmtF16 ? 2 : 4;   // planes of the patch's 16 channels
      uint4* planes[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) planes[k] = out16 + ((size_t)b * NPL + (k < NPL ? k : 0)) * plane_vox;
      for (int pl = 0; pl < Z; ++pl, ++gi) {
        const uint32_t ab = gi % kAccs;
        mbar_wait(BAR(kAccF + ab), (gi / kAccs) & 1u);
        tc_fence_after();
#pragma unroll
        for (int gg = 0; gg < kG / 2; ++gg) {
          const int g = gg * 2 + eset;
          const int pos = g * 120 + 30 * wq + lane;
          const int row = pos / kPitch, c = pos - row * kPitch;
          const int y = y0 + row, x = x0 + c;
          const bool valid = lane < 30 && row < kTY && c < kXT && y < Y && x < X;
          uint32_t r[16];
          tc_ld16(tmem_base + lane_base + ab * (kG * 16) + g * 16, r);
          tc_wait_ld();
          if (valid) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = fmaf(__uint_as_float(r[i]), 1.0f / 255.0f, bias[i]);  // the reference divides the INPUT by 255 (inferencer.py:395-399)
              v[i] = t < 0.f ? 0.f : t;                                             // ReLU; NaN passes, like torch.relu
            }
            store_voxel16<FMT>(v, planes, ((size_t)pl * Y + y) * X + x);
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(kAccE + ab));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

// 6 weight blocks (dx = 0, 1, 2) x (hi, lo): [2 K chunks][16 rows = output channels][8 halves], K index = dz * 3 + dy.
void pack_first_conv_ts_weights(const float* h_w, PackedConv& out) {
  std::vector<__half> buf((size_t)6 * 256, __float2half_rn(0.f));
  for (int dx = 0; dx < 3; ++dx)
    for (int part = 0; part < 2; ++part)
      for (int n = 0; n < 16; ++n)
        for (int k = 0; k < 9; ++k) {
          const int dz = k / 3, dy = k % 3;
          const float wv = h_w[(size_t)n * 27 + dz * 9 + dy * 3 + dx];  // (16, 1, 3, 3, 3)
          const __half hi = __float2half_rn(wv);
          buf[(size_t)(dx * 2 + part) * 256 + ((size_t)(k / 8) * 16 + n) * 8 + (k % 8)] = part == 0 ? hi : __float2half_rn(wv - __half2float(hi));
        }
  if (out.w_ts) cudaFree(out.w_ts);
  out.w_ts = nullptr;
  CFB_CUDA(cudaMalloc(&out.w_ts, buf.size() * sizeof(__half)));
  CFB_CUDA(cudaMemcpy(out.w_ts, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice));
}

void launch_first_conv_ts(const void* chunk_u8, Int3 cs, const PatchPos* patches, int nb, Int3 ps, const PackedConv& w, __half* out,
                          int fmt, cudaStream_t s) {
  if (!w.w_ts) throw std::runtime_error("first_conv_ts: weights not packed");
  FirstTsParams p{};
  p.chunk = static_cast<const uint8_t*>(chunk_u8); p.cs = cs; p.patches = patches; p.ps = ps;
  p.wblocks = w.w_ts; p.bias = w.bias; p.out = out; p.fmt = fmt;
  p.tiles_x = ceil_div(ps.x, kXT); p.tiles_y = ceil_div(ps.y, kTY);
  p.total_items = nb * p.tiles_x * p.tiles_y;
  int dev = 0, sms = 148;
  CFB_CUDA(cudaGetDevice(&dev));
  CFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_items, sms);
  if (fmt == kFmtF16F8) first_conv_ts_kernel<kFmtF16F8><<<grid, kThreads, 0, s>>>(p);
  else if (fmt == kFmtF16x2) first_conv_ts_kernel<kFmtF16x2><<<grid, kThreads, 0, s>>>(p);
  else first_conv_ts_kernel<kFmtF16><<<grid, kThreads, 0, s>>>(p);
  CFB_LAUNCH_CHECK();
}

}  // namespace cfb
