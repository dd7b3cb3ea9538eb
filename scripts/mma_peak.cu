// MMA-loop-only peak probe for tcgen05.mma on sm_100a: issues long chains of M x 256 x K MMAs from shared
// memory (operands never reloaded), one CTA (or CTA pair) per SM, and reports the sustained rate per kind.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_peak scripts/mma_peak.cu && ./mma_peak
// Output: one line per (kind, cta_group): TMAC/s over the whole chip.  This is the denominator the int8
// tensor roofline of this repository uses (BASELINE.md asked for a measured int8 figure).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// KIND: 0 = i8 (K=32), 1 = f16/bf16 (K=16), 2 = f8f6f4 e4m3 (K=32).  GROUP: 1 or 2 CTAs.
template <int KIND, int GROUP>
__global__ void __launch_bounds__(128, 1) k_peak(int iters, unsigned long long *cycles, int ncols) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  uint32_t rank = 0;
  if (GROUP == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    if (GROUP == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (GROUP == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 16384);
  // instruction descriptor: N = 256, M = 128 * GROUP, K-major operands
  const uint32_t mdim = (128 * GROUP) >> 4;
  uint32_t idesc = (((uint32_t)ncols >> 3) << 17) | (mdim << 24);
  if (KIND == 0) idesc |= (2u << 4);                               // S32 accumulate, u8 x u8
  if (KIND == 1) idesc |= (1u << 4) | (1u << 7) | (1u << 10);      // F32 accumulate, bf16 x bf16
  if (KIND == 2) idesc |= (1u << 4);                               // F32 accumulate, e4m3 x e4m3
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t a = da + 2 * k, b = db + 2 * k;
        if (KIND == 0 && GROUP == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
        if (KIND == 0 && GROUP == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
        if (KIND == 1 && GROUP == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
        if (KIND == 1 && GROUP == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
        if (KIND == 2 && GROUP == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
        if (KIND == 2 && GROUP == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1));
      }
    }
    if (GROUP == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (GROUP == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x < 32) {
    if (GROUP == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

template <int KIND, int GROUP>
void run(const char *name, int sms, int ncols = 256) {
  const int iters = getenv("MMA_ITERS") ? atoi(getenv("MMA_ITERS")) : 4000;
  unsigned long long *cyc;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMemset(cyc, 0, sizeof(unsigned long long) * sms);
  const size_t smem = 64 * 1024;
  cudaFuncSetAttribute(k_peak<KIND, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (GROUP == 1) {
      k_peak<KIND, GROUP><<<sms, 128, smem>>>(iters, cyc, ncols);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_peak<KIND, GROUP>, iters, cyc, ncols);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); return; }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[256];
  cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
  const int K = KIND == 1 ? 16 : 32;
  const double macs_per_instr = 128.0 * GROUP * ncols * K;
  const double instrs = (double)iters * 4 * (GROUP == 1 ? sms : sms / 2);
  printf("%-28s %8.3f ms  %8.1f TMAC/s (%.1f T-op/s)  cycles/instr (CTA 0) = %.1f\n", name, ms,
         instrs * macs_per_instr / (ms * 1e-3) / 1e12, 2 * instrs * macs_per_instr / (ms * 1e-3) / 1e12,
         (double)h[0] / (iters * 4));
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount & ~1;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0, 1>("i8   cta_group::1 M128 N256", sms);
  run<0, 2>("i8   cta_group::2 M256 N256", sms);
  run<1, 1>("bf16 cta_group::1 M128 N256", sms);
  run<1, 2>("bf16 cta_group::2 M256 N256", sms);
  run<2, 1>("e4m3 cta_group::1 M128 N256", sms);
  run<2, 2>("e4m3 cta_group::2 M256 N256", sms);
  // narrower accumulator tiles (the lo / hi split of the same-key schedule issues N = 128 MMAs)
  run<0, 2>("i8   cta_group::2 M256 N128", sms, 128);
  run<0, 2>("i8   cta_group::2 M256 N64", sms, 64);
  run<0, 1>("i8   cta_group::1 M128 N128", sms, 128);
  return 0;
}
