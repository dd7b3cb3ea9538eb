// tcgen05 schedule: many ciphertexts under ONE key = a dense int8 contraction against the key's
// Toeplitz matrices, int32 accumulators in TMEM, reduction / fold / witness fused in the epilogue.
//
// One launch computes, for a batch of rows x (the per-ciphertext polynomial: r, e or b) and a fixed
// key polynomial y (h, f or fp), the two halves of the linear product c = lin(x, y):
//     cyc[k] = sum_i x[i] * y[(k-i) mod N]          (= c[k] + c[k+N], the remainder of c / (1 - x^N))
//     hi [k] = sum_{i>k} x[i] * y[k+N-i]            (= c[k+N];  the quotient is -hi)
// which is everything multiplyPolynomials + dividePolynomials(., I, .) (index.js:319-401) produce on
// the hot path (closed form: SURVEY.md section 8a).  Three modes share the kernel:
//     ENC  : x = r (bytes 0..2),   y = h (mod q)   -> value/remainderE = (cyc + m) mod q, quotientE = -hi mod q
//     DEC1 : x = e (uint16 mod q), y = f (ternary) -> remainder1 = cyc mod q, quotient1 = -hi mod q,
//                                                     b = (remainder1 + [remainder1 > q/2]) mod 3   (index.js:117)
//     DEC2 : x = b (bytes 0..2),   y = fp (0..2)   -> value/remainder2 = cyc mod 3, quotient2 = -hi mod 3
//
// GEMM view per 128-row tile: D[128 x NC] += A[128 x K] * B[NC x K]^T with
//   A = the batch operand in the UMMA K-major SWIZZLE_128B layout.  r and b are bytes already: TMA
//       loads them straight from the pitched global rows (out-of-range rows / columns read as zero).
//       e is uint16: "transform" warps split it into byte limbs on the way from global to shared.
//   B = rows of the key matrix Mat (one row per accumulator column, K-major), precomputed once per
//       key by k_build_keymat and streamed by TMA (at most 3 MB, L2 resident).
// Operands wider than 8 bits:
//   ENC : h = h0 + 256*h1 -> two accumulator column groups per chunk ("N limbs": B rows [h0 | h1]),
//         recombined in the epilogue as acc0 + (acc1 << 8); A is the same r slice for both.
//   DEC1: e = e0 + 256*e1 -> two K ranges ("K limbs") A = [e0 | e1<<2], B = [f | f<<6], so that ONE
//         int32 accumulator receives the exact product (u8 x s8).
// All-zero K ranges of the triangular hi matrix are skipped at 128-byte granularity.
//
// Two kernels implement it: k_umma_pair (umma_pair.cuh, the default: one CTA pair per 256 rows, cta_group::2,
// resident A operand, B ring shared by the pair, TMA-store epilogue in two warp groups) and k_umma_product below
// (single CTA per 128 rows, kept as a cross-check).  Warp roles of k_umma_product (576 threads, 1 CTA per SM,
// persistent over row tiles):
//   warp 0      TMA producer (A and B slices)      warp 1      tcgen05.mma issuer, owns TMEM
//   warps 2-9   DEC1: e -> byte-limb A slices; ENC / DEC2: epilogue   warps 10-17 epilogue (TMEM -> regs -> global)
// Pipelines: a 4-stage shared-memory ring (full/empty mbarriers) and two 256-column TMEM accumulators
// (tmem_full/tmem_empty mbarriers) so that the epilogue of chunk j overlaps the MMAs of chunk j+1.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ntru_internal.cuh"

namespace ntru {

namespace {

enum Mode { ENC = 0, DEC1 = 1, DEC2 = 2 };

constexpr int kStages = 4;
constexpr int kTileRows = 128;
constexpr int kAtomK = 128;                      // bytes of K per pipeline slice (one 128B swizzle atom)
constexpr int kABytes = kTileRows * kAtomK;      // 16 KB
constexpr int kBBytesMax = 256 * kAtomK;         // 32 KB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kThreads = 576;
constexpr int kBuilderWarp0 = 2, kEpilogueWarp0 = 10;
constexpr int kAccCols = 256;                    // TMEM columns per accumulator buffer
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct UmmaArgs {
  int N, P, Kp, atoms;
  int kl;                 // K limbs (DEC1 with q > 256: 2)
  int nl;                 // N limbs (ENC with q > 256: 2)
  int NC;                 // accumulator columns per chunk = MMA N = nl * NCo
  int NCo;                // output coefficients per chunk
  int nchunks, with_hi, q;
  uint32_t qmask;
  size_t B;
  int ntiles;
  // 2-CTA variant (umma_pair.cuh)
  int npairs;             // 256-row pair tiles
  int nA, nB;             // shared-memory slots for A and stages for B
  int a_resident;         // A slots hold the whole tile (loaded once per tile)
  int a_tma;              // DEC1, streaming A: raw e atoms arrive by TMA in an A slot pair and are transformed in place
  int nM, nS;             // message slots (ENC), store-staging slots
  int a_rel;              // commits that free an A slot: 2 (both MMA issuers) for resident A, else 1
  int two_issuers;        // second MMA issuer warp enabled (needs slices per chunk < B ring stages)
  int debug_flags;        // timing experiments only (NTRU_DEBUG_NOB: bit 0 = skip the B operand loads)
  int mat_rows;           // rows of the key matrix (2 * nchunks * NC); K block kb starts at row kb * mat_rows
  int out_mask;           // which outputs exist: bit0 = cyc #1, bit1 = cyc #2, bit2 = hi
  const void *a_src;      // DEC1: e rows (uint16), pitch P elements
  const uint8_t *m;       // ENC: message rows
  uint16_t *o16_cyc;      // ENC: value, DEC1: remainder1
  uint16_t *o16_cyc2;     // ENC: remainderE (same data, second destination)
  uint16_t *o16_hi;       // ENC: quotientE, DEC1: quotient1
  uint8_t *o8_cyc;        // DEC1: b, DEC2: value
  uint8_t *o8_cyc2;       // DEC2: remainder2
  uint8_t *o8_hi;         // DEC2: quotient2
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4, LBO = 1 (ignored for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// cute::UMMA::InstrDescriptor for kind::i8: c_format=S32 (2) at [4,6), a_format at [7,10), b_format at [10,13)
// (0 = u8, 1 = s8), K-major A and B, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int a_signed, int b_signed, int n) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileRows >> 4) << 24);
}

// first 128-byte K atom that can hold a non-zero entry of the hi matrix for chunk c: i >= c*NCo + 1
__device__ __forceinline__ int first_atom(const UmmaArgs &a, int part_hi, int c) {
  if (!part_hi) return 0;
  const int a0 = (c * a.NCo + 1) / kAtomK;
  return a0 < a.atoms ? a0 : a.atoms - 1;
}

// Enumerates the (tile, part, chunk, atom) work items of one CTA in pipeline order.
struct AtomIter {
  int tile, part, c, at, parts;
  bool valid;
  __device__ __forceinline__ void start(const UmmaArgs &a) {
    parts = a.with_hi ? 2 : 1;
    tile = blockIdx.x; part = 0; c = 0;
    valid = tile < a.ntiles;
    at = valid ? first_atom(a, a.with_hi, 0) : 0;
  }
  __device__ __forceinline__ void next(const UmmaArgs &a) {
    if (++at < a.atoms) return;
    if (++c == a.nchunks) {
      c = 0;
      if (++part == parts) {
        part = 0;
        tile += gridDim.x;
        if (tile >= a.ntiles) { valid = false; return; }
      }
    }
    at = first_atom(a, a.with_hi && part == 0, c);
  }
};

// One accumulator chunk of one epilogue warp: prefetch what the chunk needs from global memory, wait for the
// MMAs, then TMEM -> registers -> reduction / fold / witness -> global.  `sub` = this warp's index among the kSub
// warps of its TMEM lane quadrant; it owns the 16-coefficient units sub, sub + kSub, ...
struct NoMark {
  __device__ __forceinline__ void operator()(int) const {}
};

template <int MODE, int kSub, class WaitFn, class MarkFn = NoMark>
__device__ __forceinline__ void epilogue_chunk(const UmmaArgs &a, int hi, int c, int sub, bool row_ok, size_t rbase,
                                               uint32_t t_addr, WaitFn wait_acc, MarkFn mark = MarkFn()) {
  constexpr int kUnitsPerWarp = 16 / kSub;
  const int units = a.NCo >> 4;
  const uint32_t Q2 = a.qmask | (a.qmask << 16);             // the modulus mask in both 16-bit lanes
  const uint32_t lift_add = ((uint32_t)a.q >> 1) - 1;        // x > q/2  <=>  (x + q/2 - 1) >> log2(q)
  const int logq = 31 - __clz(a.q);
  // ENC: fetch this thread's message bytes for the whole chunk before waiting on the accumulator;
  // bytes of coefficients >= N are cleared so that the pad of every output row is written as zero
  uint4 mm[kUnitsPerWarp];
  if (MODE == ENC && !hi) {
#pragma unroll
    for (int ui = 0; ui < kUnitsPerWarp; ++ui) {
      const int u = sub + kSub * ui;
      const int kk = c * a.NCo + u * 16;
      const bool ok = row_ok && u < units && kk < a.N;
      uint4 x = ok ? __ldg(reinterpret_cast<const uint4 *>(a.m + rbase + kk)) : make_uint4(0, 0, 0, 0);
      const int nvalid = a.N - kk;
      if (ok && nvalid < 16) {
        uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
          const int nb = nvalid - 4 * wd;               // valid bytes in this word
          xs[wd] = nb >= 4 ? xs[wd] : (nb <= 0 ? 0u : (xs[wd] & (0xffffffffu >> (8 * (4 - nb)))));
        }
        x = make_uint4(xs[0], xs[1], xs[2], xs[3]);
      }
      mm[ui] = x;
    }
  }
  wait_acc();
  tc_fence_after();
#pragma unroll
  for (int ui = 0; ui < kUnitsPerWarp; ++ui) {
    const int u = sub + kSub * ui;
    if (u >= units) break;
    uint32_t w[32];
    tmem_ld16(t_addr + u * 16, w);
    if (MODE == ENC && a.nl == 2) {
      uint32_t w1[32];
      tmem_ld16(t_addr + a.NCo + u * 16, w1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] += w1[j] << 8;
    } else {
      tmem_ld_wait();
    }
    mark(3 + 3 * ui);   // accumulators of this unit are in registers
    const int kk = c * a.NCo + u * 16;
    if (!row_ok || kk >= a.P) continue;
    if (MODE == ENC || MODE == DEC1) {
      // two coefficients per 32-bit word, reduced mod q in both lanes at once
      uint32_t pk[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) pk[jj] = __byte_perm(w[2 * jj], w[2 * jj + 1], 0x5410);
      if (hi) {             // -hi mod q = ((q-1-x) + 1) mod q, no carry between lanes
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) pk[jj] = ((~pk[jj] & Q2) + 0x00010001u) & Q2;
      } else if (MODE == ENC) {
        const uint32_t mw[4] = {mm[ui].x, mm[ui].y, mm[ui].z, mm[ui].w};
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t mp = __byte_perm(mw[jj >> 1], 0u, (jj & 1) ? 0x4342 : 0x4140);   // bytes -> lanes
          pk[jj] = ((pk[jj] & Q2) + mp) & Q2;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) pk[jj] &= Q2;
      }
      const uint4 p0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      const uint4 p1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      mark(4 + 3 * ui);   // math done
      uint16_t *d0 = hi ? a.o16_hi : a.o16_cyc;
      if (d0) {
        reinterpret_cast<uint4 *>(d0 + rbase + kk)[0] = p0;
        reinterpret_cast<uint4 *>(d0 + rbase + kk)[1] = p1;
      }
      if (!hi && a.o16_cyc2) {
        reinterpret_cast<uint4 *>(a.o16_cyc2 + rbase + kk)[0] = p0;
        reinterpret_cast<uint4 *>(a.o16_cyc2 + rbase + kk)[1] = p1;
      }
      if (MODE == DEC1 && !hi && a.o8_cyc) {
        uint32_t bq[4];
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
          uint32_t bb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t x = w[4 * wd + j] & a.qmask;
            const uint32_t y = x + ((x + lift_add) >> logq);          // index.js:117
            bb[j] = y - 3u * __umulhi(y, 0x55555556u);
          }
          bq[wd] = __byte_perm(__byte_perm(bb[0], bb[1], 0x0040), __byte_perm(bb[2], bb[3], 0x0040), 0x5410);
        }
        *reinterpret_cast<uint4 *>(a.o8_cyc + rbase + kk) = make_uint4(bq[0], bq[1], bq[2], bq[3]);
      }
    } else {   // DEC2: mod 3 (hi: -x mod 3 = 2x mod 3)
      uint32_t bq[4];
#pragma unroll
      for (int wd = 0; wd < 4; ++wd) {
        uint32_t bb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t y = hi ? 2u * w[4 * wd + j] : w[4 * wd + j];
          bb[j] = y - 3u * __umulhi(y, 0x55555556u);
        }
        bq[wd] = __byte_perm(__byte_perm(bb[0], bb[1], 0x0040), __byte_perm(bb[2], bb[3], 0x0040), 0x5410);
      }
      const uint4 pk4 = make_uint4(bq[0], bq[1], bq[2], bq[3]);
      uint8_t *d0 = hi ? a.o8_hi : a.o8_cyc;
      if (d0) *reinterpret_cast<uint4 *>(d0 + rbase + kk) = pk4;
      if (!hi && a.o8_cyc2) *reinterpret_cast<uint4 *>(a.o8_cyc2 + rbase + kk) = pk4;
    }
  }
}

// ---- the kernel --------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
k_umma_product(const UmmaArgs a, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kStages * kStageBytes);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, then the TMEM base address
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kStages + 2 + b); };
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), MODE == DEC1 ? 1 + 8 : 1);   // TMA expect_tx arrival (+ 8 transform warps)
      mbar_init(empty_bar(s), 1);                         // tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);        // tcgen05.commit
      mbar_init(tempty_bar(b), MODE == DEC1 ? 8 : 16);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int parts = a.with_hi ? 2 : 1;          // part 0 = hi (when present), last part = cyc

  if (warp == 0) {
    // ===================== TMA producer: B slices (and A slices unless DEC1) =====================
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t bytes = (uint32_t)a.NC * kAtomK + (MODE == DEC1 ? 0u : (uint32_t)kABytes);
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int part = 0; part < parts; ++part) {
          const int hi = a.with_hi && part == 0;
          for (int c = 0; c < a.nchunks; ++c) {
            const int a0 = first_atom(a, hi, c);
            const int row0 = (hi * a.nchunks + c) * a.NC;
            for (int at = a0; at < a.atoms; ++at) {
              for (int lk = 0; lk < a.kl; ++lk, ++it) {
                const int s = it % kStages;
                mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
                mbar_arrive_expect_tx(full_bar(s), bytes);
                if (MODE != DEC1)
                  tma_load_2d(smem_base + s * kStageBytes, &tmapA, at * kAtomK, tile * kTileRows, full_bar(s));
                tma_load_2d(smem_base + s * kStageBytes + kABytes, &tmapB, 0, (lk * a.atoms + at) * a.mat_rows + row0, full_bar(s));
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(0, MODE == DEC1 ? 1 : 0, a.NC);
      uint32_t it = 0, cc = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int part = 0; part < parts; ++part) {
          const int hi = a.with_hi && part == 0;
          for (int c = 0; c < a.nchunks; ++c, ++cc) {
            const int a0 = first_atom(a, hi, c);
            const int buf = cc & 1;
            mbar_wait(tempty_bar(buf), ((cc >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * kAccCols;
            uint32_t first = 1;
            for (int at = a0; at < a.atoms; ++at) {
              for (int lk = 0; lk < a.kl; ++lk, ++it) {
                const int s = it % kStages;
                mbar_wait(full_bar(s), (it / kStages) & 1);
                tc_fence_after();
                const uint64_t da = make_smem_desc(smem_base + s * kStageBytes);
                const uint64_t db = make_smem_desc(smem_base + s * kStageBytes + kABytes);
#pragma unroll
                for (int k = 0; k < kAtomK / 32; ++k) {
                  umma_i8(d_tmem, da + 2 * k, db + 2 * k, idesc, first ? 0u : 1u);   // +32 bytes of K per step
                  first = 0;
                }
                umma_commit(empty_bar(s));
              }
            }
            umma_commit(tfull_bar(buf));
          }
        }
      }
    }
  } else if (MODE == DEC1 && warp < kEpilogueWarp0) {
    // ===================== DEC1 transform: e (uint16, global) -> byte-limb A slices (swizzled smem) =====
    {
      const int t = threadIdx.x - kBuilderWarp0 * 32;      // 0..255
      const int chunk = t & 7;                             // 16-byte chunk of the 128-byte A row
      const int r0 = t >> 3;                               // rows r0, r0+32, r0+64, r0+96
      const uint16_t *src = reinterpret_cast<const uint16_t *>(a.a_src);
      auto load_atom = [&](const AtomIter &w, uint4 (&raw)[8]) {
        const int col = w.at * kAtomK + chunk * 16;        // first coefficient of this thread's chunk
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const size_t row = (size_t)w.tile * kTileRows + r0 + 32 * j;
          const bool ok = w.valid && row < a.B && col < a.P;
          const uint4 *ptr = reinterpret_cast<const uint4 *>(src + (ok ? row * (size_t)a.P + col : 0));
          uint4 x0 = __ldg(ptr), x1 = __ldg(ptr + 1);
          if (!ok) x0 = x1 = make_uint4(0, 0, 0, 0);
          raw[2 * j] = x0;
          raw[2 * j + 1] = x1;
        }
      };
      AtomIter cur, nxt;
      cur.start(a);
      uint4 raw[8], raw_next[8];
      if (cur.valid) load_atom(cur, raw);
      uint32_t it = 0;
      while (cur.valid) {
        nxt = cur;
        nxt.next(a);
        load_atom(nxt, raw_next);                           // prefetch one atom ahead (zeros when !valid)
        const int s0 = it % kStages, s1 = (it + 1) % kStages;
        mbar_wait(empty_bar(s0), ((it / kStages) & 1) ^ 1);
        if (a.kl == 2) mbar_wait(empty_bar(s1), (((it + 1) / kStages) & 1) ^ 1);
        uint8_t *dst0 = smem + (size_t)s0 * kStageBytes;
        uint8_t *dst1 = smem + (size_t)s1 * kStageBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r_in = r0 + 32 * j;
          const int off = (r_in >> 3) * 1024 + (r_in & 7) * 128 + ((chunk ^ (r_in & 7)) << 4);
          const uint4 w0 = raw[2 * j], w1 = raw[2 * j + 1];
          uint4 lo;
          lo.x = __byte_perm(w0.x, w0.y, 0x6420);
          lo.y = __byte_perm(w0.z, w0.w, 0x6420);
          lo.z = __byte_perm(w1.x, w1.y, 0x6420);
          lo.w = __byte_perm(w1.z, w1.w, 0x6420);
          *reinterpret_cast<uint4 *>(dst0 + off) = lo;
          if (a.kl == 2) {                                  // (e >> 8) << 2 in the low byte of each 16-bit field
            uint4 hi4;
            hi4.x = __byte_perm((w0.x >> 6) & 0x00FC00FCu, (w0.y >> 6) & 0x00FC00FCu, 0x6420);
            hi4.y = __byte_perm((w0.z >> 6) & 0x00FC00FCu, (w0.w >> 6) & 0x00FC00FCu, 0x6420);
            hi4.z = __byte_perm((w1.x >> 6) & 0x00FC00FCu, (w1.y >> 6) & 0x00FC00FCu, 0x6420);
            hi4.w = __byte_perm((w1.z >> 6) & 0x00FC00FCu, (w1.w >> 6) & 0x00FC00FCu, 0x6420);
            *reinterpret_cast<uint4 *>(dst1 + off) = hi4;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(full_bar(s0));
          if (a.kl == 2) mbar_arrive(full_bar(s1));
        }
        it += a.kl;
        cur = nxt;
#pragma unroll
        for (int j = 0; j < 8; ++j) raw[j] = raw_next[j];
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    // DEC1: warps 10-17 (2 per TMEM lane quadrant); ENC / DEC2: warps 2-17 (4 per quadrant).
    constexpr int kSub = MODE == DEC1 ? 2 : 4;                 // epilogue warps per quadrant
    const int ew = warp - (MODE == DEC1 ? kEpilogueWarp0 : kBuilderWarp0);
    const int quad = warp & 3;                     // TMEM lanes [32*quad, 32*quad+32) are this warp's
    const int sub = ew >> 2;                       // this warp's units: sub, sub + kSub, ...
    uint32_t cc = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const size_t row = (size_t)tile * kTileRows + quad * 32 + lane;
      const bool row_ok = row < a.B;
      const size_t rbase = row * (size_t)a.P;
      for (int part = 0; part < parts; ++part) {
        const int hi = a.with_hi && part == 0;
        for (int c = 0; c < a.nchunks; ++c, ++cc) {
          const int buf = cc & 1;
          const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kAccCols;
          epilogue_chunk<MODE, kSub>(a, hi, c, sub, row_ok, rbase, t_addr,
                                     [&] { mbar_wait(tfull_bar(buf), (cc >> 1) & 1); });
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

#include "umma_pair.cuh"

// ---- key matrix ---------------------------------------------------------------------------------
// Mat[row][kb]: row = ((part * nchunks + c) * nl + ln) * NCo + j  (part 0 = cyc, 1 = hi; output k = c*NCo + j),
//               kb  = lk * Kp + i.
__global__ void k_build_keymat(int mode, int N, int Kp, int kl, int nl, int NCo, int nchunks, const void *poly,
                               uint8_t *mat) {
  const int klen = kl * Kp;
  const size_t total = (size_t)2 * nchunks * nl * NCo * klen;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int kb = (int)(idx % klen);
    int row = (int)(idx / klen);
    const int j = row % NCo; row /= NCo;
    const int ln = row % nl; row /= nl;
    const int c = row % nchunks;
    const int part = row / nchunks;
    const int k = c * NCo + j;
    const int lk = kb / Kp, i = kb % Kp;
    int coef = 0;
    bool nz = k < N && i < N;
    int src = 0;
    if (nz) {
      if (part == 0) {
        src = k - i;
        if (src < 0) src += N;
      } else {
        nz = i > k;
        src = k + N - i;
      }
    }
    if (nz) {
      if (mode == ENC) coef = reinterpret_cast<const uint16_t *>(poly)[src];
      else if (mode == DEC1) coef = reinterpret_cast<const int8_t *>(poly)[src];
      else coef = reinterpret_cast<const uint8_t *>(poly)[src];
    }
    uint8_t out;
    if (mode == ENC) out = ln == 0 ? (uint8_t)(coef & 0xff) : (uint8_t)(coef >> 8);
    else if (mode == DEC1) out = (uint8_t)(int8_t)(lk == 0 ? coef : coef * 64);
    else out = (uint8_t)coef;
    // tile-major storage: for each 128-byte K block all rows are contiguous (128-byte pitch), so that the box
    // of one pipeline slice (NC or NC/2 rows x 128 B) is one contiguous run of global memory
    const size_t rows_total = (size_t)2 * nchunks * nl * NCo;
    mat[((size_t)(kb / kAtomK) * rows_total + (size_t)(idx / klen)) * kAtomK + (kb % kAtomK)] = out;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

void geometry(const ntru_ctx *ctx, int mode, int kl, int nl, KeyMatrix &km) {
  const int N = ctx->N;
  km.limbs = kl;
  km.nlimbs = nl;
  // output coefficients per chunk: a 256-column accumulator holds 256 / nl; ENC is capped at 128 so that a
  // chunk's message bytes are exactly one 128-byte TMA atom per row
  const int max_out = mode == ENC ? 128 : 256 / nl;
  km.nchunks = (N + max_out - 1) / max_out;
  const int per = (N + km.nchunks - 1) / km.nchunks;
  // the CTA-pair kernel's epilogue stages whole passes of 32 (ENC, DEC1) or 64 (DEC2) coefficients per warp:
  // every warp of a TMEM lane quadrant must own a whole number of passes of each chunk
  const int round = mode == ENC ? 128 : (mode == DEC1 ? 64 : 256);
  km.out_cols = ((per + round - 1) / round) * round;
  if (km.out_cols > max_out) km.out_cols = max_out;
  km.chunk_cols = km.out_cols * nl;
  const int Kp = ((N + kAtomK - 1) / kAtomK) * kAtomK;
  km.klen = kl * Kp;
}

int encode_2d_ex(ntru_ctx *ctx, void *out, void *base, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t stride_bytes,
                 uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ctx, NTRU_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15) != 0) return fail(ctx, NTRU_E_PARAM, "device arrays must be 16-byte aligned");
  const ntru_ctx::TmapKey key = {base, inner, rows, stride_bytes, box_inner, box_rows, elem_bytes, (int)swz};
  for (const auto &e : ctx->tmap_cache)
    if (e.key == key) {
      memcpy(out, e.map, 128);
      return NTRU_OK;
    }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap *>(out), elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                   2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NTRU_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  if (ctx->tmap_cache.size() < ntru_ctx::kTmapCacheMax) {
    ctx->tmap_cache.emplace_back();
    ctx->tmap_cache.back().key = key;
    memcpy(ctx->tmap_cache.back().map, out, 128);
  } else {
    auto &e = ctx->tmap_cache[ctx->tmap_next++ % ntru_ctx::kTmapCacheMax];
    e.key = key;
    memcpy(e.map, out, 128);
  }
  return NTRU_OK;
}

int encode_2d(ntru_ctx *ctx, void *out, void *base, uint64_t inner, uint64_t rows, uint64_t stride_bytes, uint32_t box_inner,
              uint32_t box_rows) {
  return encode_2d_ex(ctx, out, base, 1, inner, rows, stride_bytes, box_inner, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

int build_keymat(ntru_ctx *ctx, int mode, int kl, int nl, const void *poly, KeyMatrix &km) {
  geometry(ctx, mode, kl, nl, km);
  const int rows = 2 * km.nchunks * km.chunk_cols;
  const int Kp = km.klen / kl;
  const size_t bytes = (size_t)rows * km.klen;
  NTRU_CUDA(ctx, km.mat.reserve(bytes));
  {
    LaunchTimer timer(ctx, NTRU_K_OTHER);
    k_build_keymat<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(mode, ctx->N, Kp, kl, nl, km.out_cols, km.nchunks, poly,
                                                               (uint8_t *)km.mat.ptr);
  }
  NTRU_CUDA(ctx, cudaGetLastError());
  const uint64_t tiled_rows = (uint64_t)rows * (km.klen / kAtomK);
  int rc = encode_2d(ctx, km.tmap, km.mat.ptr, kAtomK, tiled_rows, kAtomK, kAtomK, (uint32_t)km.chunk_cols);
  if (rc) return rc;
  rc = encode_2d(ctx, km.tmap_half, km.mat.ptr, kAtomK, tiled_rows, kAtomK, kAtomK, (uint32_t)km.chunk_cols / 2);
  if (rc) return rc;
  NTRU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  km.ready = true;
  return NTRU_OK;
}

template <int MODE>
int launch_product(ntru_ctx *ctx, const KeyMatrix &km, UmmaArgs &a, const void *a_bytes) {
  if (!(ctx->umma_attr_set & (1 << MODE))) {      // per context: function attributes are per device
    NTRU_CUDA(ctx, cudaFuncSetAttribute(k_umma_product<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NTRU_CUDA(ctx, cudaFuncSetAttribute(k_umma_pair<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
    ctx->umma_attr_set |= 1 << MODE;
  }
  a.N = ctx->N; a.P = ctx->P; a.kl = km.limbs; a.nl = km.nlimbs; a.Kp = km.klen / km.limbs; a.atoms = a.Kp / kAtomK;
  a.NC = km.chunk_cols; a.NCo = km.out_cols; a.nchunks = km.nchunks; a.q = ctx->q; a.qmask = (uint32_t)ctx->q - 1;
  a.mat_rows = 2 * a.nchunks * a.NC;
  a.ntiles = (int)((a.B + kTileRows - 1) / kTileRows);
  a.npairs = (int)((a.B + 2 * kTileRows - 1) / (2 * kTileRows));
  a.nS = 2;
  a.nM = MODE == ENC ? 2 : 0;
  const int a_slots = a.atoms * a.kl;
  // ENC above N = 512: k_umma_pair<ENC, 0, 1> stages one unit per TMA store (one staging slot) and reads the message
  // bytes from global memory (no message slots).  The three slots go to the B ring, and the A operand stays resident up
  // to N = 1024: with 4 stages the ring covered ~2000 cycles of TMA latency at 900 cycles per slice (clock trace,
  // NTRU_DEBUG_NOB timing), and streaming A doubled the L2 -> SM traffic at N = 821.
  bool pu1 = MODE == ENC && ctx->tensor_variant == 0 && a.atoms >= 5;
#ifdef NTRU_TRACE   // timing experiments exist in trace builds only: the shipped library reads no environment variable
  if (getenv("NTRU_DEBUG_NO_PU1")) pu1 = false;
#endif
  if (pu1) { a.nS = 1; a.nM = 0; }
  const int avail = kPairSlots - a.nS - a.nM;
  if (a_slots + 4 <= avail) {
    a.a_resident = 1; a.nA = a_slots; a.nB = avail - a_slots;
  } else {
    a.a_resident = 0; a.nA = 6; a.nB = avail - 6;
  }
  // streaming DEC1: the raw e atoms (two A slots each) are prefetched by TMA, four atoms deep
  a.a_tma = MODE == DEC1 && !a.a_resident && a.kl == 2 && ctx->tensor_variant == 0 && avail >= 11;
  if (a.a_tma) { a.nA = 8; a.nB = avail - 8; }
  // resident A slots are read by both MMA issuer warps (alternating chunks) whenever a tile has >= 2 chunks
  a.two_issuers = 0;   // a second issuer warp was tried: the B ring is too short for two chunks in flight
  a.a_rel = (a.a_resident && a.two_issuers) ? 2 : 1;
  CUtensorMap tmB, tmA, tmM, tmO[3];
  const bool pair = ctx->tensor_variant == 0;
  memcpy(&tmB, pair ? km.tmap_half : km.tmap, sizeof tmB);
  memset(&tmA, 0, sizeof tmA);
  memset(&tmM, 0, sizeof tmM);
  memset(tmO, 0, sizeof tmO);
  const uint64_t N = (uint64_t)ctx->N, P = (uint64_t)ctx->P;
  int rc;
  if (MODE != DEC1) {
    // byte rows straight into the UMMA layout: inner extent N (columns beyond read as zero), row pitch P
    rc = encode_2d(ctx, &tmA, const_cast<void *>(a_bytes), N, (uint64_t)a.B, P, kAtomK, kTileRows);
    if (rc) return rc;
  } else if (a.a_tma) {
    // raw uint16 rows, one atom = 128 coefficients x 128 rows = 32 KB, row-major (256-byte rows) in shared memory
    rc = encode_2d_ex(ctx, &tmA, const_cast<void *>(a.a_src), 2, P, (uint64_t)a.B, P * 2, kAtomK, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  if (pair) {
    if (MODE == ENC) {
      rc = encode_2d(ctx, &tmM, const_cast<uint8_t *>(a.m), N, (uint64_t)a.B, P, kAtomK, kTileRows);
      if (rc) return rc;
    }
    // outputs: per-warp tiles of 32 rows x 64 bytes (SWIZZLE_64B); DEC1's b is 32 rows x 32 bytes, unswizzled
    void *optr[3];
    int oelem[3], obox[3];
    CUtensorMapSwizzle oswz[3];
    if (MODE == ENC || MODE == DEC1) {
      const int ob = pu1 ? 16 : 32;
      const CUtensorMapSwizzle osw = pu1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
      optr[0] = a.o16_cyc; oelem[0] = 2; obox[0] = ob; oswz[0] = osw;
      optr[2] = a.o16_hi; oelem[2] = 2; obox[2] = ob; oswz[2] = osw;
      if (MODE == ENC) { optr[1] = a.o16_cyc2; oelem[1] = 2; obox[1] = ob; oswz[1] = osw; }
      else { optr[1] = a.o8_cyc; oelem[1] = 1; obox[1] = 32; oswz[1] = CU_TENSOR_MAP_SWIZZLE_NONE; }
    } else {
      optr[0] = a.o8_cyc; optr[1] = a.o8_cyc2; optr[2] = a.o8_hi;
      for (int i = 0; i < 3; ++i) { oelem[i] = 1; obox[i] = 64; oswz[i] = CU_TENSOR_MAP_SWIZZLE_64B; }
    }
    a.out_mask = 0;
    for (int i = 0; i < 3; ++i) {
      if (!optr[i]) continue;
      a.out_mask |= 1 << i;
      rc = encode_2d_ex(ctx, &tmO[i], optr[i], oelem[i], P, (uint64_t)a.B, P * oelem[i], (uint32_t)obox[i], 32, oswz[i]);
      if (rc) return rc;
    }
#ifdef NTRU_TRACE
    if (getenv("NTRU_DEBUG_NOSTORE")) a.out_mask = 0;   // timing experiment only: results are not written
    if (getenv("NTRU_DEBUG_NOB")) a.debug_flags |= 1;
#endif
  }
  // The accumulator chunks cover nchunks * NCo >= N output columns; when N is a multiple of the chunk width (N = 512,
  // 640, 768, 1024) that stops short of the row pitch, and the pad columns the kernel never sees are zeroed here so that
  // every output row keeps the contract of include/ntru_b200.h.
  {
    const size_t cov = (size_t)a.nchunks * a.NCo;
    if (cov < P) {
      void *o16[3] = {a.o16_cyc, a.o16_cyc2, a.o16_hi};
      void *o8[3] = {MODE == DEC2 ? a.o8_cyc : nullptr, a.o8_cyc2, a.o8_hi};
      for (int i = 0; i < 3; ++i) {
        if (o16[i]) NTRU_CUDA(ctx, cudaMemset2DAsync((uint16_t *)o16[i] + cov, P * 2, 0, (P - cov) * 2, a.B, ctx->stream));
        if (o8[i]) NTRU_CUDA(ctx, cudaMemset2DAsync((uint8_t *)o8[i] + cov, P, 0, P - cov, a.B, ctx->stream));
      }
    }
  }
  {
    LaunchTimer timer(ctx, MODE == ENC ? NTRU_K_ENC_TENSOR : (MODE == DEC1 ? NTRU_K_DEC1_TENSOR : NTRU_K_DEC2_TENSOR));
    if (pair) {
      const int clusters = a.npairs < ctx->sm_count / 2 ? a.npairs : ctx->sm_count / 2;
#ifdef NTRU_TRACE
      const int dbg = getenv("NTRU_DEBUG_EPI") ? atoi(getenv("NTRU_DEBUG_EPI")) : 0;
#define NTRU_DBG_LAUNCH(D)                                                                                                  \
  case D:                                                                                                                  \
    cudaFuncSetAttribute(k_umma_pair<MODE, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);          \
    k_umma_pair<MODE, D><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmA, tmM, tmO[0], tmO[1], tmO[2]); \
    break;
      switch (dbg) {
        NTRU_DBG_LAUNCH(1) NTRU_DBG_LAUNCH(2) NTRU_DBG_LAUNCH(4) NTRU_DBG_LAUNCH(8) NTRU_DBG_LAUNCH(12) NTRU_DBG_LAUNCH(16)
        default:
          if (pu1) {
            cudaFuncSetAttribute(k_umma_pair<ENC, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);
            k_umma_pair<ENC, 0, 1><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmA, tmM, tmO[0], tmO[1], tmO[2]);
          } else {
            k_umma_pair<MODE><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmA, tmM, tmO[0], tmO[1], tmO[2]);
          }
      }
#else
      if (pu1) {
        if (!(ctx->umma_attr_set & 8)) {
          NTRU_CUDA(ctx, cudaFuncSetAttribute(k_umma_pair<ENC, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
          ctx->umma_attr_set |= 8;
        }
        k_umma_pair<ENC, 0, 1><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmA, tmM, tmO[0], tmO[1], tmO[2]);
      } else {
        k_umma_pair<MODE><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmA, tmM, tmO[0], tmO[1], tmO[2]);
      }
#endif
    } else {
      const int grid = a.ntiles < ctx->sm_count ? a.ntiles : ctx->sm_count;
      k_umma_product<MODE><<<grid, kThreads, kSmemBytes, ctx->stream>>>(a, tmB, tmA);
    }
  }
  NTRU_CUDA(ctx, cudaGetLastError());
  return NTRU_OK;
}

}  // namespace

#ifdef NTRU_TRACE
extern "C" int ntru_debug_trace_dump(unsigned long long *out, unsigned int cap) {
  cudaDeviceSynchronize();
  const unsigned int n = kTraceLanes * kTraceCap;
  if (cap < n) return -1;
  cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * n);
  return (int)n;
}
#endif

int umma_init(ntru_ctx *ctx) {
  ctx->tensor_ok = false;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return NTRU_OK;
  if (prop.major != 10) return NTRU_OK;                       // tcgen05 needs sm_100a
  if ((size_t)prop.sharedMemPerBlockOptin < kSmemBytes || (size_t)prop.sharedMemPerBlockOptin < kPairSmemBytes) return NTRU_OK;
  if (!get_encode_fn()) return NTRU_OK;
  ctx->tensor_ok = true;
  return NTRU_OK;
}

int umma_prepare_public(ntru_ctx *ctx) {
  return build_keymat(ctx, ENC, 1, ctx->q > 256 ? 2 : 1, ctx->d_h.ptr, ctx->km_h);
}

int umma_prepare_private(ntru_ctx *ctx) {
  int rc = build_keymat(ctx, DEC1, ctx->q > 256 ? 2 : 1, 1, ctx->d_f.ptr, ctx->km_f);
  if (rc) return rc;
  return build_keymat(ctx, DEC2, 1, 1, ctx->d_fp.ptr, ctx->km_fp);
}

int umma_encrypt(ntru_ctx *ctx, size_t B, const uint8_t *r, const uint8_t *m, uint16_t *value, uint16_t *quo,
                 uint16_t *rem) {
  if (B == 0) return NTRU_OK;
  UmmaArgs a = {};
  a.B = B; a.m = m;
  a.with_hi = quo != nullptr;
  a.o16_cyc = value ? value : rem;
  a.o16_cyc2 = value ? rem : nullptr;
  a.o16_hi = quo;
  return launch_product<ENC>(ctx, ctx->km_h, a, r);
}

int umma_decrypt(ntru_ctx *ctx, size_t B, const uint16_t *e, uint8_t *value, uint16_t *q1, uint16_t *r1, uint8_t *q2,
                 uint8_t *r2) {
  if (B == 0) return NTRU_OK;
  NTRU_CUDA(ctx, ctx->d_b.reserve(B * (size_t)ctx->P));
  UmmaArgs a = {};
  a.B = B; a.a_src = e;
  a.with_hi = q1 != nullptr;
  a.o16_cyc = r1; a.o16_hi = q1; a.o8_cyc = (uint8_t *)ctx->d_b.ptr;
  int rc = launch_product<DEC1>(ctx, ctx->km_f, a, nullptr);
  if (rc) return rc;
#ifdef NTRU_TRACE
  if (getenv("NTRU_TRACE_DEC1_ONLY")) return NTRU_OK;   // leaves the DEC1 timeline in g_trace
#endif
  UmmaArgs b = {};
  b.B = B;
  b.with_hi = q2 != nullptr;
  b.o8_cyc = value ? value : r2;
  b.o8_cyc2 = value ? r2 : nullptr;
  b.o8_hi = q2;
  return launch_product<DEC2>(ctx, ctx->km_fp, b, ctx->d_b.ptr);
}

}  // namespace ntru
