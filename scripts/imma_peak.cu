// Register-fragment tensor path probe on sm_100a: sustained rate of mma.sync.m16n8k32 (u8 x s8 -> s32, SASS IMMA.16832)
// and of dp4a (IDP.4A) from registers only, for the distinct-key schedule (one small product per ciphertext, no
// operand shared between ciphertexts, so tcgen05's 128-row tiles do not apply).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak scripts/imma_peak.cu && ./imma_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ACC>
__global__ void k_imma(int iters, int *out) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
  int c[ACC][4];
#pragma unroll
  for (int j = 0; j < ACC; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) imma(c[j], a, b);
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < ACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (s == 0x7fffffff) out[0] = s;
}

template <int ACC>
__global__ void k_dp4a(int iters, int *out) {
  uint32_t a = threadIdx.x * 0x01010101u, b = threadIdx.x ^ 0x02030405u;
  int c[ACC];
#pragma unroll
  for (int j = 0; j < ACC; ++j) c[j] = j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) c[j] = __dp4a((int)a, (int)(b + j), c[j]);
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < ACC; ++j) s += c[j];
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

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int *out;
  cudaMalloc(&out, 4);
  const int sms = prop.multiProcessorCount, iters = 20000;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { k_imma<8><<<sms, warps * 32>>>(iters, out); });
    double macs = (double)sms * warps * iters * 8 * (16.0 * 8 * 32);
    printf("IMMA.16832 u8xs8  %2d warps/SM x 8 accumulators  %.3f ms  %.1f TMAC/s  (%.0f MAC/clk/SM at max clock)\n", warps, ms,
           macs / ms / 1e9, macs / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
  }
  for (int warps : {8, 16, 32}) {
    float ms = time_ms([&] { k_dp4a<8><<<sms, warps * 32>>>(iters, out); });
    double macs = (double)sms * warps * 32 * iters * 8 * 4.0;
    printf("IDP.4A            %2d warps/SM x 8 accumulators  %.3f ms  %.1f TMAC/s  (%.0f MAC/clk/SM at max clock)\n", warps, ms,
           macs / ms / 1e9, macs / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
