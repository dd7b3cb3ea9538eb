// Semantics probe for the next step of the same-key schedule: a TMA load that is BOTH .cta_group::2 (completion may be
// credited to the pair leader's mbarrier) AND .multicast::cluster (one L2 read lands in two CTAs of a 4-CTA cluster).
// Question: when CTA r issues
//     cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster
//         [dst], [map, {0, 0}], [mbar], mask
// with mask = {r, r ^ 2} and mbar = the barrier at offset X in the LEADER of r's pair (mapa(X, r & ~1)), which
// mbarrier(s) receive the complete_tx: offset X in (a) the leader of EACH destination's pair, (b) each destination
// CTA itself, or (c) only the literal mbar?  Every CTA initialises barrier X with one pending arrival, posts
// arrive.expect_tx(bytes expected if it were credited), polls with a bounded number of try_waits and reports.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o multicast_probe scripts/multicast_probe.cu -lcuda && ./multicast_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: loaders = ranks 0 and 1 (pair 0), each multicasts its 4 KB tile to {r, r+2}; mbar = leader of OWN pair (rank 0).
// Every leader (0 and 2) expects 8 KB; non-leaders expect nothing and only report what landed in their shared memory.
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
k_probe(const __grid_constant__ CUtensorMap map, int mode, uint32_t *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t b = smem_u32(&bar);
  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const bool leader = (rank & 1) == 0;
  if (threadIdx.x == 0) {
    // leaders expect both halves of their pair (8 KB), the others only their own 4 KB: who completes tells where the
    // complete_tx of a multicast load is credited
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(leader ? 8192 : 4096) : "memory");
    if (rank < 2) {   // loaders
      const uint32_t dst = smem_u32(smem);
      const uint32_t mbar = mode == 0 ? mapa(b, rank & ~1u) : b;
      const uint16_t mask = (uint16_t)((1u << rank) | (1u << (rank ^ 2u)));
      asm volatile(
          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
          " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst), "l"(&map), "r"(0), "r"((int)rank * 32), "r"(mbar), "h"(mask)
          : "memory");
    }
    // bounded wait: ~2 ms
    bool done = false;
    const long long t0 = clock64();
    while (clock64() - t0 < 4000000ll) {
      if (try_wait(b, 0)) { done = true; break; }
    }
    out[rank * 4 + 0] = done ? 1u : 0u;
    out[rank * 4 + 3] = (uint32_t)(clock64() - t0);
  }
  __syncthreads();
  // give stragglers time, then report what is in shared memory (first word, and a word from row 31)
  if (threadIdx.x == 0) {
    const long long t1 = clock64();
    while (clock64() - t1 < 200000ll) {}
    out[rank * 4 + 1] = reinterpret_cast<volatile uint32_t *>(smem)[0];
    out[rank * 4 + 2] = reinterpret_cast<volatile uint32_t *>(smem)[31 * 32];
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || !fp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  // global matrix: 64 rows x 128 bytes; rows 0-31 hold 0x11111111 (tile of rank 0), rows 32-63 hold 0x22222222 (rank 1)
  uint32_t *g, h[64 * 32];
  for (int r = 0; r < 64; ++r)
    for (int c = 0; c < 32; ++c) h[r * 32 + c] = r < 32 ? 0x11111111u : 0x22222222u;
  cudaMalloc(&g, sizeof h);
  cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t gdim[2] = {128, 64}, gstride[1] = {128};
  cuuint32_t box[2] = {128, 32}, estr[2] = {1, 1};
  CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
  uint32_t *out, ho[16];
  cudaMalloc(&out, sizeof ho);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(out, 0, sizeof ho);
    k_probe<<<4, 128, 4096>>>(map, mode, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
    printf("mode %d (mbar = %s): each loader r in {0,1} multicasts its 4 KB tile to CTAs {r, r+2}; leaders 0 and 2 expect 8 KB, CTAs 1 and 3 expect 4 KB\n", mode,
           mode == 0 ? "leader of the loader's pair, mapa(X, r & ~1)" : "the loader's own barrier X");
    for (int r = 0; r < 4; ++r)
      printf("  CTA %d: barrier %s after %u cycles; smem[0] = %08x, smem[row 31] = %08x\n", r, ho[r * 4] ? "COMPLETED" : "not completed",
             ho[r * 4 + 3], ho[r * 4 + 1], ho[r * 4 + 2]);
  }
  // how many clusters of 2 / 4 / 8 CTAs with the pair kernel's footprint (one CTA per SM: 225 KB dynamic shared memory,
  // 608 threads) can be resident at once
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 225 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_probe, &cfg);
    printf("cluster size %d: max active clusters = %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
