// What does one trip of an MMA issue loop cost on sm_100a, piece by piece?  One warp of a CTA pair's leader runs the loop
// warp-uniformly (one elected lane issues, as in csrc/umma_pair.cuh); the MMAs are tiny (M = 256, N = 16: no tensor time
// to hide behind), every barrier it polls is already complete, so cycles per trip = the instruction / latency cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_loop_probe scripts/issue_loop_probe.cu && ./issue_loop_probe
// Variants (bits of V):  1 = four MMAs per trip            2 = tcgen05.commit (multicast to both CTAs) per trip
//                        4 = mbarrier.try_wait per trip     8 = tcgen05.fence::after_thread_sync per trip
//                       16 = __syncwarp before and after the elected block
//                       32 = ring counters + descriptor arithmetic as in the production loop (stage wrap, parity, two descriptors)
//                       64 = a second commit per trip (A slot release)
//                      128 = plain lane-0 branch instead of elect.sync
//                      256 = the fifteen other warps of the CTA are busy (dependent integer math + shared-memory loads), as epilogue warps are
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\t.reg .b32 r;\n\telect.sync r|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int V>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) k_probe(int iters, unsigned long long *cycles, int nB) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[32];      // [0..15] "full" (complete from the start: waited with the parity of the phase before), [16..31] "empty" (commit targets)
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bars[i]), 1u << 20);   // never completes: commits only add arrivals
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  const uint32_t idesc = (2u << 4) | ((16u >> 3) << 17) | ((256u >> 4) << 24);   // i8, N = 16, M = 256
  const uint32_t a_lo0 = smem_u32(smem) >> 4, b_lo0 = (smem_u32(smem) + 4 * 16384) >> 4;   // 4 A slots, then nB <= 6 B stages: 160 KB
  const uint64_t desc_hi = make_desc(0) & ~0x3FFFull;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[16]);
  const int lane = threadIdx.x & 31;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if ((V & 256) && threadIdx.x >= 32 && rank == 0) {
    // busy neighbours: a dependent chain of integer ops and shared-memory loads until the issuer is done
    uint32_t x = threadIdx.x, acc = 0;
    const uint32_t *sm = (const uint32_t *)smem;
    while (!stop) {
#pragma unroll 8
      for (int i = 0; i < 64; ++i) { x = x * 1664525u + 1013904223u; acc += sm[(x >> 8) & 8191u]; }
    }
    if (acc == 0x12345678u) cycles[1] = acc;
  }
  if (threadIdx.x < 32 && rank == 0) {
    uint32_t sb = 0, b_par = 0, sa = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // a barrier in its first phase: waiting for parity 1 (the phase "before" it) returns at once
      if (V & 4) mbar_wait(full0 + 8u * sb, (V & 32) ? (b_par | 1u) : 1u);
      if (V & 16) __syncwarp();
      if (V & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint64_t da = desc_hi | (uint64_t)a_lo0, db = desc_hi | (uint64_t)b_lo0;
      if (V & 32) {
        da = desc_hi | (uint64_t)(a_lo0 + sa * 1024u);
        db = desc_hi | (uint64_t)(b_lo0 + sb * 1024u);
      }
      const bool me = (V & 128) ? lane == 0 : elect_one();
      if (me) {
        mma(tmem, da, db, idesc, 1u);
        if (V & 1) {
          mma(tmem, da + 2, db + 2, idesc, 1u);
          mma(tmem, da + 4, db + 4, idesc, 1u);
          mma(tmem, da + 6, db + 6, idesc, 1u);
        }
        if (V & 2) commit(empty0 + 8u * sb);
        if (V & 64) commit(empty0 + 64u + 8u * sa);
      }
      if (V & 16) __syncwarp();
      if (V & 32) {
        if (++sb == (uint32_t)nB) { sb = 0; b_par ^= 1; }
        if (++sa == 4u) sa = 0;
      }
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int V>
void run(const char *what) {
  const int iters = 20000, sms = 2;
  unsigned long long *cyc;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMemset(cyc, 0, sizeof(unsigned long long) * sms);
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k_probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k_probe<V><<<sms, 512, smem>>>(iters, cyc, 6);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("V=%3d: CUDA error %s\n", V, cudaGetErrorString(err)); exit(1); }
  unsigned long long h[2];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("V=%3d  %7.1f cycles per trip   %s\n", V, (double)h[0] / iters, what);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s; MMAs are M = 256, N = 16, K = 32 (about 8 cycles of tensor work each): a trip costs what its instructions cost\n", p.name);
  run<0>("one MMA, elect");
  run<1>("four MMAs");
  run<1 | 2>("four MMAs + commit");
  run<1 | 2 | 64>("four MMAs + two commits");
  run<1 | 2 | 4>("try_wait (complete barrier) + four MMAs + commit");
  run<1 | 2 | 4 | 8>("... + tcgen05.fence::after_thread_sync");
  run<1 | 2 | 4 | 8 | 16>("... + __syncwarp before and after the elected block");
  run<1 | 2 | 4 | 8 | 16 | 32>("... + ring counters and descriptor arithmetic (the production trip without A-slot logic)");
  run<1 | 2 | 4 | 8 | 16 | 32 | 64>("... + second commit");
  run<1 | 2 | 4 | 8 | 16 | 32 | 128>("same, lane 0 instead of elect.sync");
  run<1 | 2 | 4 | 8 | 16 | 32 | 64 | 256>("the production trip, fifteen busy warps beside the issuer");
  run<1 | 2 | 256>("four MMAs + commit, fifteen busy warps");
  run<4>("try_wait + one MMA");
  run<4 | 16>("try_wait + syncwarps + one MMA");
  return 0;
}
