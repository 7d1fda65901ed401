// Probe: can tcgen05.cp (shared memory -> tensor memory, issued by ONE thread, asynchronous) replace the four loader warps
// of conv3_ts_umma_kernel (ld.shared + tcgen05.st, bound by the LSU pipe)?  Measures
//   * correctness of tcgen05.cp.128x256b with the canonical no-swizzle K-major descriptor at an arbitrary 16-byte start,
//   * cycles per copy, alone and in the kernel's instruction stream (copy, MMA, shift, MMA, shift, MMA), with and without
//     issuing the next tile's copy ahead of the current tile's MMAs,
//   * kind::f8f6f4 (e4m3, K = 32) MMA rate next to kind::f16.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tc_cp probe_tc_cp.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ta(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(bdesc), "r"(idesc), "r"(1));
}
__device__ __forceinline__ void mma8_ta(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(bdesc), "r"(idesc), "r"(1));
}
__device__ __forceinline__ void cp128(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void shift(uint32_t a) { asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// ---- correctness: records of 16 bytes, record j of K-chunk c holds the words (c << 16 | j) x 4 ----------------------------
__global__ void __launch_bounds__(128, 1) check_cp(int p0, int lbo_records, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  uint4* rec = reinterpret_cast<uint4*>(smem);
  for (int j = threadIdx.x; j < 2048; j += 128) {   // chunk 0: records 0..1023, chunk 1 starts lbo_records later (filled separately)
    const uint32_t v0 = (uint32_t)j;
    rec[j] = make_uint4(v0, v0 + 0x10000u, v0 + 0x20000u, v0 + 0x30000u);
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t sb = smem_u32(smem) + 16u * (uint32_t)p0;
    const uint64_t sdesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)(uint32_t)lbo_records << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    cp128(tmem, sdesc);
    commit(smem_u32(&bar));
  }
  wait_bar(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[8];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 8 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// ---- rates ------------------------------------------------------------------------------------------------------------
// 0: cp only (8 different source tiles / TMEM slots)      1: MMA,shift,MMA,shift,MMA (no copy)
// 2: per tile: cp ; MMA,shift,MMA,shift,MMA               3: same, copy of tile i+1 issued before the MMAs of tile i
// 4: 3 x f16 MMA                                          5: 3 x e4m3 MMA (K = 32)
// 6: like 3 with f16 / e4m3 MMAs alternating per tile (the f16f8 group)
template <int pattern>
__global__ void __launch_bounds__(128, 1) rate(int N, int R, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 32768; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;  // 128 KB
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t bdesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)(N * 16 >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    // source tiles: 128 records from smem + 16 KB + i * 4800 B, second K chunk 32 KB later
    auto sdesc = [&](int i) {
      const uint32_t a = sb + 16384u + (uint32_t)i * 4800u;
      return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(32768u >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    };
    const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = tmem + 384, d0 = tmem;
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
      if (pattern == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cp128(a0 + 8 * i, sdesc(i));
      } else if (pattern == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t a = a0 + 8 * i, d = d0 + 48 * (i & 3);
          mma_ta(d, a, bdesc, idesc); shift(a); mma_ta(d, a, bdesc + 256, idesc); shift(a); mma_ta(d, a, bdesc + 512, idesc);
        }
      } else if (pattern == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t a = a0 + 8 * i, d = d0 + 48 * (i & 3);
          cp128(a, sdesc(i));
          mma_ta(d, a, bdesc, idesc); shift(a); mma_ta(d, a, bdesc + 256, idesc); shift(a); mma_ta(d, a, bdesc + 512, idesc);
        }
      } else if (pattern == 3 || pattern == 6) {
        cp128(a0, sdesc(0));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t a = a0 + 8 * i, d = d0 + 48 * (i & 3);
          if (i < 7) cp128(a + 8, sdesc(i + 1));
          if (pattern == 6 && (i & 4)) {
            mma8_ta(d, a, bdesc, idesc); shift(a); mma8_ta(d, a, bdesc + 256, idesc); shift(a); mma8_ta(d, a, bdesc + 512, idesc);
          } else {
            mma_ta(d, a, bdesc, idesc); shift(a); mma_ta(d, a, bdesc + 256, idesc); shift(a); mma_ta(d, a, bdesc + 512, idesc);
          }
        }
      } else if (pattern == 4) {
        mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc); mma_ta(d0, a0, bdesc, idesc);
      } else if (pattern == 5) {
        mma8_ta(d0, a0, bdesc, idesc); mma8_ta(d0, a0, bdesc, idesc); mma8_ta(d0, a0, bdesc, idesc);
      }
    }
    commit(smem_u32(&bar));
    wait_bar(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  // ---- correctness
  uint32_t* d_chk; cudaMalloc(&d_chk, 128 * 8 * 4);
  cudaFuncSetAttribute(check_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int p0 : {0, 8, 3, 35}) {
    const int lbo = 1024;   // second K chunk = records 1024.. (16 KB later)
    check_cp<<<1, 128, 32768 + 256>>>(p0, lbo, d_chk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check_cp p0=%d: %s\n", p0, cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> h(128 * 8);
    cudaMemcpy(h.data(), d_chk, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int l = 0; l < 128; ++l)
      for (int i = 0; i < 8; ++i) {
        const uint32_t want = (uint32_t)(p0 + l + (i >= 4 ? lbo : 0)) + 0x10000u * (uint32_t)(i & 3);
        if (h[l * 8 + i] != want) { if (bad < 4) printf("  lane %d col %d: got %08x want %08x\n", l, i, h[l * 8 + i], want); ++bad; }
      }
    printf("tcgen05.cp.128x256b start record %2d: %s (%d mismatches)\n", p0, bad ? "MISMATCH" : "lane l = record p0 + l, cols 0-3 chunk 0, cols 4-7 chunk 1: OK", bad);
  }
  // ---- rates
  long long* d_out; cudaMalloc(&d_out, 8);
  const char* names[] = {"8 copies", "8 x (MMA,shift,MMA,shift,MMA)", "8 x (copy; MMA,shift,MMA,shift,MMA)", "same, copy issued one tile ahead",
                         "3 MMA f16", "3 MMA e4m3 K=32", "copy ahead, f16 / e4m3 tiles (f16f8 group)"};
  using K = void (*)(int, int, long long*);
  K kernels[] = {rate<0>, rate<1>, rate<2>, rate<3>, rate<4>, rate<5>, rate<6>};
  for (K k : kernels) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int R = 300;
  for (int N : {16, 32, 48, 96})
    for (int pat = 0; pat < 7; ++pat) {
      kernels[pat]<<<148, 128, 131072 + 256>>>(N, R, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pattern %d N %d: %s\n", pat, N, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
      printf("N %3d  %-44s %8.1f cyc/iter\n", N, names[pat], (double)c / R);
    }
  return 0;
}
