// Throughput probe for the tensor-core instruction stream used by conv3_ts_umma_kernel:
// how many SM cycles do  {MMA(A in TMEM), tcgen05.shift.down, MMA(A in shared memory)}  cost back to back?
// One elected thread issues R repetitions of a pattern, commits, waits; cycles = clock64 delta / R.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tc_rate probe_tc_rate.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ta(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(bdesc), "r"(idesc), "r"(1));
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(1));
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void shift(uint32_t a) { asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(a) : "memory"); }

// pattern: 0 = 3 MMA(TA)           1 = MMA,shift,MMA,shift,MMA (TA)      2 = 2 shifts only
//          3 = 3 MMA(SS)           4 = 3 MMA(TA) on 3 different A tiles  5 = 6 MMA(TA)+2 shifts (hi part of the split kernel)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// NOTE: the issuing thread must be chosen with elect.sync -- with `threadIdx.x == 0` ptxas wraps every UTCHMMA in an
// ELECT/branch loop and the probe measures ~70 cycles of issue overhead per MMA instead of the tensor pipe.
template <int pattern>
__global__ void __launch_bounds__(128, 1) probe(int N, int R, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar, bar2;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;  // 64 KB of small fp16 values
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t sb = smem_u32(smem);
    // B: N rows x K=16, K-major no swizzle: LBO (K chunk) = N*16 B, SBO (8 rows) = 128 B
    const uint64_t bdesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)(N * 16 >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    // A (SS mode): 128 rows x K=16 at smem + 16 KB, LBO = 2048 B
    const uint64_t adesc = (uint64_t)(((sb + 16384) >> 4) & 0x3FFF) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = tmem + 384, d0 = tmem;
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
      switch (pattern) {
        case 0: mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); break;
        case 1: mma_ta(d0, a0, bdesc, idesc); shift(a0); mma_ta(d0, a0, bdesc, idesc); shift(a0); mma_ta(d0, a0, bdesc, idesc); break;
        case 2: shift(a0); shift(a0); break;
        case 3: mma_ss(d0, adesc, bdesc, idesc); mma_ss(d0, adesc, bdesc, idesc); mma_ss(d0, adesc, bdesc, idesc); break;
        case 4: mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0 + 8, bdesc, idesc); mma_ta(d0, a0 + 16, bdesc, idesc); break;
        case 5:
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); shift(a0);
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); shift(a0);
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); break;
        case 7: mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); commit(smem_u32(&bar2)); break;
        case 8:
#pragma unroll
          for (int i = 0; i < 12; ++i) mma_ta(d0, a0, bdesc, idesc);
          commit(smem_u32(&bar2)); break;
        case 9:
#pragma unroll
          for (int i = 0; i < 12; ++i) mma_ta(d0, a0, bdesc, idesc);
          break;
        case 10: commit(smem_u32(&bar2)); break;
        case 11:  // 4 MMAs, 4 different B blocks (4 KB apart), same A / D
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc + 256, idesc); mma_ta(d0, a0, bdesc + 512, idesc); mma_ta(d0, a0, bdesc + 768, idesc); break;
        case 12:  // 4 MMAs, same B, 4 different (A tile, D columns)
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0 + 96, a0 + 8, bdesc, idesc); mma_ta(d0 + 192, a0 + 16, bdesc, idesc); mma_ta(d0 + 288, a0 + 24, bdesc, idesc); break;
        case 13:  // 4 MMAs, different B, different (A tile, D columns)
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0 + 96, a0 + 8, bdesc + 256, idesc); mma_ta(d0 + 192, a0 + 16, bdesc + 512, idesc); mma_ta(d0 + 288, a0 + 24, bdesc + 768, idesc); break;
        case 14:  // the kernel's order for one tile: (hi, lo B) x dx with shifts, different B each
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc + 256, idesc); shift(a0);
          mma_ta(d0, a0, bdesc + 512, idesc); mma_ta(d0, a0, bdesc + 768, idesc); shift(a0);
          mma_ta(d0, a0, bdesc + 1024, idesc); mma_ta(d0, a0, bdesc + 1280, idesc); break;
        case 15: case 16: case 17: case 18: {  // one full group of the 16->16 TMEM-shift kernel: G = 4 tiles, T = 3, hi + lo parts
          const uint32_t dofs = pattern == 16 ? 16u : (pattern == 17 ? 32u : 0u);
          uint32_t a = a0;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            uint32_t d = d0 + dofs;
#pragma unroll
            for (int g = 0; g < 4; ++g, d += 48, a += 8) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                if (dx) shift(a);
                mma_ta(d, a, bdesc + dx * 288, idesc);
                if (part == 0) mma_ta(d, a, bdesc + dx * 288 + 48, idesc);
              }
            }
          }
          commit(smem_u32(&bar2));
          break;
        }
        case 6:  // alternate accumulators (no D dependency between consecutive MMAs)
          mma_ta(d0, a0, bdesc, idesc); mma_ta(d0 + 128, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); break;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  const char* names[] = {"3 MMA(TA)", "MMA,shift,MMA,shift,MMA", "2 shifts", "3 MMA(SS)", "3 MMA(TA) 3 tiles", "6 MMA + 2 shifts", "3 MMA(TA) alt D", "3 MMA + commit", "12 MMA + commit", "12 MMA", "commit only", "4 MMA 4 B", "4 MMA 4 (A,D)", "4 MMA 4 (A,D,B)", "6 MMA 6 B + 2 shifts", "group16 D+0", "group16 D+16", "group16 D+32", "group16 D+0 again"};
  const int R = 400;
  using K = void (*)(int, int, long long*);
  K kernels[] = {probe<0>, probe<1>, probe<2>, probe<3>, probe<4>, probe<5>, probe<6>, probe<7>, probe<8>, probe<9>, probe<10>, probe<11>, probe<12>, probe<13>, probe<14>, probe<15>, probe<16>, probe<17>, probe<18>};
  for (K k : kernels) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
  for (int grid : {148})
    for (int N : {16, 32, 48})
      for (int pat : {5, 15, 16, 17}) {
        kernels[pat]<<<grid, 128, 65536 + 256>>>(N, R, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pattern %d N %d: %s\n", pat, N, cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d N %3d  %-26s %8.1f cyc/iter   (floor 128*N/256 = %d per MMA)\n", grid, N, names[pat], (double)c / R, N / 2);
      }
  return 0;
}
