// TMEM -> register read throughput probe (tcgen05.ld) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_peak scripts/tmem_ld_peak.cu && ./tmem_ld_peak
// One CTA per SM, W warps (W = 4, 8, 16), each warp issues back-to-back loads of its own 32 TMEM lanes.
// Reports accumulator bytes per clock per SM for: 32x32b.x16, .x32, .x64, and .x16.pack::16b / .x32.pack::16b
// (32 / 64 columns of which only the low 16 bits are returned).  The epilogue of the NTRU kernels is bound by
// this number: every accumulator column has to come back through tcgen05.ld.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD16(addr, v)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                \
               : "r"(addr))
#define LD16P(addr, v)                                                                                             \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                \
               : "r"(addr))
#define LD32(addr, v)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"   \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                          \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),       \
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),      \
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])               \
               : "r"(addr))

// SHAPE: 0 = x16, 1 = x32, 2 = x16.pack::16b (covers 32 columns), 3 = 16x256b.x4 (not used)
template <int SHAPE>
__global__ void __launch_bounds__(512, 1) k_ld(int iters, unsigned long long *cycles, uint32_t *sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tslot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[32];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 64 + (warp >> 2) * 32) & 255);
    if (SHAPE == 0) { LD16(base + col, v); LD16(base + col + 16, v); }
    if (SHAPE == 1) { LD32(base + col, v); }
    if (SHAPE == 2) { LD16P(base + col, v); }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += v[0] ^ v[15];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0 && warp == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "n"(512));
}

template <int SHAPE>
void run(const char *name, int warps) {
  const int iters = 2000, sms = 148;
  unsigned long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMalloc(&sink, sizeof(uint32_t) * sms * 512);
  k_ld<SHAPE><<<sms, warps * 32>>>(iters, cyc, sink);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  // every iteration each warp reads 32 lanes x 32 columns (x 4 bytes of accumulator)
  const double bytes = (double)iters * warps * 32 * 32 * 4;
  printf("%-26s warps=%2d  %7.1f cycles/iter  %6.1f accumulator B/clk/SM\n", name, warps, (double)h[0] / iters,
         bytes / (double)h[0]);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("32x32b.x16 (two per iter)", w);
    run<1>("32x32b.x32", w);
    run<2>("32x32b.x16.pack::16b", w);
  }
  return 0;
}
