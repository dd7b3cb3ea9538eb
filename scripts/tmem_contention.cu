// Does tcgen05.ld compete with tcgen05.mma for TMEM?  One CTA per SM: warp 0 issues a chain of i8 MMAs
// (M128 N256 K32, accumulator columns 0-255) while warps 4.. read accumulator columns 256-511 with
// tcgen05.ld.32x32b.x16 in a loop.  Prints MMA cycles/instr and load bytes/clk, with and without the other.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
#define LD16(addr, v)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                \
               : "r"(addr))
__global__ void __launch_bounds__(640, 1) k(int mma_iters, int ld_iters, int ld_warps, unsigned long long *out, uint32_t *sink) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 0) {
    if (lane == 0 && mma_iters > 0) {
      const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 16384);
      const uint32_t idesc = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      const long long t0 = clock64();
      for (int it = 0; it < mma_iters; ++it)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(1));
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
      out[2 * blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    uint32_t acc = 0, v[16];
    const long long t0 = clock64();
    for (int it = 0; it < ld_iters; ++it) {
      LD16(base + ((it * 16 + (warp >> 2) * 64) & 255), v);
      LD16(base + ((it * 16 + 128 + (warp >> 2) * 64) & 255), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] ^ v[15];
    }
    if (warp == 4 && lane == 0) out[2 * blockIdx.x + 1] = (unsigned long long)(clock64() - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
int main() {
  unsigned long long *out; uint32_t *sink;
  cudaMalloc(&out, 16 * 148 * 2); cudaMalloc(&sink, 4 * 148 * 640);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int cfgs[][3] = {{4000, 0, 0}, {0, 4000, 16}, {4000, 40000, 16}, {4000, 40000, 8}, {4000, 40000, 4}};
  for (auto &c : cfgs) {
    cudaMemset(out, 0, 16 * 148 * 2);
    k<<<148, 640, 64 * 1024>>>(c[0], c[1], c[2], out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[2];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("mma_iters=%5d ld_iters=%6d ld_warps=%2d :", c[0], c[1], c[2]);
    if (c[0]) printf("  MMA %.1f cycles/instr", (double)h[0] / (c[0] * 4));
    if (c[1]) printf("  LD  %.1f cycles/iter/warp -> %.0f B/clk/SM", (double)h[1] / c[1], (double)c[2] * 32 * 32 * 4 / ((double)h[1] / c[1]));
    printf("\n");
  }
  return 0;
}
