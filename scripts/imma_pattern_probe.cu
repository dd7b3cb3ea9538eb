// What bounds a stream of IMMA.16832 whose operands change from one instruction to the next (the distinct-key schedule,
// imma_kernels.cu: six A fragments x seven B fragments per group of steps, accumulator sg + d)?
//   mode 0: one A and one B fragment for every MMA (the peak probe, scripts/imma_peak.cu)
//   mode 1: SG x D distinct fragments held in registers, accumulator sg + d
//   mode 2: mode 1, the fragments re-read from shared memory every pass (LDS.32 x 3 + funnel shifts per A, LDS.64 per B)
//   mode 3: mode 2 with two MMAs per (sg, d) into the same accumulator (the merged two-limb form: B and 64 B)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_pattern_probe scripts/imma_pattern_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int SG = 3, D = 6, NJ = SG + D - 1;

template <int MODE>
__global__ void k_probe(int iters, int *out) {
  __shared__ __align__(16) uint8_t sm[32 * 1024];
  for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = i * 2654435761u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t *base = sm + (warp & 7) * 4096 + 64;
  const int g = lane >> 2, tq = lane & 3;
  const uint8_t *pa = base + 1024 + 4 * ((16 * (tq >> 1) + 2 * g + 8 * (tq & 1)) >> 2);
  const uint8_t *pb = base + 16 * (g + 1 - (tq >> 1)) + 8 * (tq & 1);
  const uint32_t sh0 = 8u * (g & 1) + 8u, sh1 = sh0 + 8u;
  int c[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
  uint32_t a[SG][4];
  uint2 b[D];
#pragma unroll
  for (int i = 0; i < SG; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = threadIdx.x * (i + 3u) + j;
#pragma unroll
  for (int d = 0; d < D; ++d) b[d] = make_uint2(threadIdx.x ^ (d * 77u), d + 11u);
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < SG; ++i) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(pa + 128 * i + 32 * (it & 7));
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        a[i][0] = __funnelshift_rc(w0, w1, sh1);
        a[i][1] = __funnelshift_r(w0, w1, sh0);
        a[i][2] = __funnelshift_rc(w1, w2, sh1);
        a[i][3] = __funnelshift_r(w1, w2, sh0);
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      uint2 bb = b[d];
      if (MODE >= 2) bb = *reinterpret_cast<const uint2 *>(pb + 128 * d + 32 * (it & 3));
#pragma unroll
      for (int i = 0; i < SG; ++i) {
        if (MODE == 0) imma(c[i + d], a[0], b[0].x, b[0].y);
        else imma(c[i + d], a[i], bb.x, bb.y);
        if (MODE == 3) imma(c[i + d], a[(i + 1) % SG], (bb.x << 6) & 0xC0C0C0C0u, (bb.y << 6) & 0xC0C0C0C0u);
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 0x7fffffff) out[0] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

template <int MODE>
void run(int sms, int *out, double clock_hz) {
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_probe<MODE><<<sms, warps * 32>>>(iters, out); });
    const double n = (double)sms * warps * iters * SG * D * (MODE == 3 ? 2 : 1);
    printf("mode %d  %2d warps/SM  %.3f ms  %.1f TMAC/s  %.2f cycles per IMMA and scheduler\n", MODE, warps, ms, n * 4096 / ms / 1e9,
           ms * 1e-3 * clock_hz / (n / sms / 4));
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int *out;
  cudaMalloc(&out, 4);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz, SG %d x D %d\n", prop.name, sms, prop.clockRate, SG, D);
  run<0>(sms, out, prop.clockRate * 1e3);
  run<1>(sms, out, prop.clockRate * 1e3);
  run<2>(sms, out, prop.clockRate * 1e3);
  run<3>(sms, out, prop.clockRate * 1e3);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
