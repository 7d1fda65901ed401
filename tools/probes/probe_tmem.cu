// Probe of TMEM semantics needed for a TMEM-resident activation tile (round-2 idea):
//  (1) tcgen05.shift.down: which rows move where, what fills the first row, behaviour at 32-lane boundaries
//  (2) tcgen05.mma with the A operand in TMEM: row m = lane m, K=16 fp16 = 8 packed 32-bit columns?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tmem probe_tmem.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out_shift, float* out_mma, int* status) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(128) __half sB[2 * 16 * 8];  // B: N=16 rows x K=16, K-major no-swizzle: [2 chunks][16 rows][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // B[n][k] = (n == k) ? 1 : 0  -> D = A
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += 128) {
    const int kc = i / 128, n = (i / 8) % 16, e = i % 8;
    sB[i] = __float2half((n == kc * 8 + e) ? 1.0f : 0.0f);
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  // ---- A tile in TMEM columns [0, 8): row m, packed fp16 pairs: A[m][k] = m + k/100
  {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      __half2 h = __floats2half2_rn((float)m + (2 * j) * 0.01f, (float)m + (2 * j + 1) * 0.01f);
      r[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  // ---- (2) MMA with A from TMEM: D[128 x 16] (cols 32..47) = A(tmem cols 0..7) * B^T
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    const uint64_t bdesc = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)16 << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + 32), "r"(tmem), "l"(bdesc), "r"(idesc), "r"(0));
    // back-to-back, no waits: shift, MMA into cols 64.., shift, MMA into cols 96.. (tests issue-order execution)
    asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem) : "memory");
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + 64), "r"(tmem), "l"(bdesc), "r"(idesc), "r"(0));
    asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem) : "memory");
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + 96), "r"(tmem), "l"(bdesc), "r"(idesc), "r"(0));
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + lane_base + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out_mma[m * 16 + j] = __uint_as_float(r[j]);
    for (int t = 1; t <= 2; ++t) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tmem + lane_base + 32 + 32 * t));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out_mma[2048 * t + m * 16] = __uint_as_float(r[0]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  // ---- (1) shift the A tile down by one row, read it back
  if (threadIdx.x == 0) {
    asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + lane_base));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out_shift[m * 8 + j] = r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
  if (threadIdx.x == 0) *status = 1;
}

int main() {
  uint32_t* d_shift; float* d_mma; int* d_status;
  cudaMalloc(&d_shift, 128 * 8 * 4); cudaMalloc(&d_mma, 3 * 2048 * 4); cudaMalloc(&d_status, 4);
  cudaMemset(d_status, 0, 4);
  probe<<<1, 128>>>(d_shift, d_mma, d_status);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint32_t> sh(128 * 8); std::vector<float> mm(3 * 2048);
  cudaMemcpy(sh.data(), d_shift, sh.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(mm.data(), d_mma, mm.size() * 4, cudaMemcpyDeviceToHost);
  // MMA check: D[m][n] should equal A[m][n] = m + n/100 (fp16 rounded)
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    float ref = __half2float(__float2half((float)m + n * 0.01f));
    double err = fabs(mm[m * 16 + n] - ref); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
  }
  printf("A-from-TMEM MMA: max err %.4g, bad %d / 2048;  D[0][0..3] = %.3f %.3f %.3f %.3f  D[5][0..3] = %.3f %.3f %.3f %.3f\n", maxerr, bad,
         mm[0], mm[1], mm[2], mm[3], mm[80], mm[81], mm[82], mm[83]);
  for (int t = 1; t <= 2; ++t) {
    int ok = 0, tot = 0;
    for (int m = 0; m < 128; ++m) if (m % 32 < 32 - t) { ++tot; ok += (mm[2048 * t + m * 16] == (float)(m + t)); }
    printf("pipelined shift x%d then MMA (no waits in between): %d / %d lanes see row m+%d;  lanes 0,1,30,31,32: %g %g %g %g %g\n", t, ok, tot, t,
           mm[2048 * t], mm[2048 * t + 16], mm[2048 * t + 30 * 16], mm[2048 * t + 31 * 16], mm[2048 * t + 32 * 16]);
  }
  // shift: print the row id (integer part of first element) now held by selected lanes
  printf("after shift.down, lane -> source row (first half of column 0): ");
  for (int m : {0, 1, 2, 3, 30, 31, 32, 33, 34, 62, 63, 64, 65, 95, 96, 97, 126, 127}) {
    __half2 h = *reinterpret_cast<__half2*>(&sh[m * 8]);
    printf("%d:%g ", m, __half2float(__low2half(h)));
  }
  printf("\ncolumn 7 check: ");
  for (int m : {0, 1, 2, 33, 127}) { __half2 h = *reinterpret_cast<__half2*>(&sh[m * 8 + 7]); printf("%d:%g ", m, __half2float(__high2half(h))); }
  printf("\n");
  return 0;
}
