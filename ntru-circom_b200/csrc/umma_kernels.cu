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
// All-zero K ranges of the triangular hi matrix are skipped at 128-byte granularity, and the last K atom issues only the
// 32-byte MMA steps that hold coefficients below N.
//
// The N output columns are cut into accumulator CHUNKS (at most 256 TMEM columns each, two chunks in flight).  Chunk
// widths come from a table (UmmaArgs::col0): the widths are as equal as the epilogue's granularity allows and add up to
// N rounded up to that granularity, not to a multiple of 256 -- at N = 821 DEC1 computes 832 columns (224 + 224 + 192 +
// 192) where uniform 256-wide chunks computed 1024.
//
// The kernel is k_umma_pair (umma_pair.cuh): one CTA pair per 256 rows, cta_group::2, resident A operand, B ring shared
// by the pair, TMA-store epilogue in two warp groups.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ntru_internal.cuh"

namespace ntru {

namespace {

// DEC1F: the first decrypt product for 256 < q <= 2048 on kind::f16 tiles.  A uint16 coefficient e < 2048 read as an fp16
// bit pattern IS the number e * 2^-24 (the subnormals and the first normal binade of fp16 form one linear ramp), so the
// caller's rows are the A operand as they lie in memory: no byte-limb transform, no transform warps, all sixteen warps
// drain accumulators.  f is exact in fp16, every partial sum is a multiple of 2^-24 below 2^-2 and therefore exact in
// the fp32 accumulators, and the integer comes back in the epilogue as the low bits of (acc + 0.75f).  The MMA work
// and the operand bytes equal those of the two byte limbs (K = 16 per instruction instead of 32, two bytes per entry).
enum Mode { ENC = 0, DEC1 = 1, DEC2 = 2, DEC1F = 3 };

constexpr int kTileRows = 128;
constexpr int kAtomK = 128;                      // bytes of K per pipeline slice (one 128B swizzle atom)
constexpr int kABytes = kTileRows * kAtomK;      // 16 KB
constexpr int kAccCols = 256;                    // TMEM columns per accumulator buffer

// One tile (256 rows of a CTA pair) is a list of PHASES, each a run of MMAs into one TMEM accumulator buffer followed by
// one epilogue read of it; phase j of a tile uses buffer (running phase count) & 1.
//   PH_CYC: the cyclic matrix over every K atom, fresh accumulator           -> remainder epilogue
//   PH_HI : the hi matrix over the atoms that hold i > k, fresh accumulator   -> quotient epilogue
//   PH_LO : the lo matrix (i <= k) over the atoms up to the chunk's last column, ACCUMULATED ON TOP of the hi product the
//           same buffer still holds (cyc = lo + hi; the quotient epilogue only read it)  -> remainder epilogue
// With the quotient witness, chunks are taken in pairs [HI a, HI b, LO a, LO b] (two buffers in flight; a LO phase finds
// its HI product two phases back in the same buffer): lo + hi is N^2 + the diagonal blocks instead of the 1.5 N^2 of
// cyc + hi -- 12 instead of 14 slices per tile and K limb at N = 509, 35 instead of 44 at N = 821.  An odd last chunk
// runs [CYC, HI] (its lo product spans every atom anyway).  Without the witness: [CYC ...] only.
constexpr int kMaxPhases = 2 * kMaxChunks;
enum PhaseKind { PH_CYC = 0, PH_HI = 1, PH_LO = 2 };
struct Phase {
  int8_t c, kind, a0, a1;   // chunk, kind, K atoms [a0, a1)
  uint16_t first, rel;      // resident A operand: atoms whose first / last read of the tile happens in this phase
};

struct UmmaArgs {
  int nph;                // phases per tile
  Phase ph[kMaxPhases];
  int8_t a_order[16];     // resident A operand: the K atoms in the order of their first use in a tile
  int N, P, Kp, atoms;
  int ea, ea_shift;       // coefficients per 128-byte K atom and its log2: 128 (byte operands), 64 (DEC1F: 16-bit operands)
  int k_last;             // 32-byte MMA steps of the last K atom that hold coefficients below N (1..4)
  int kl;                 // K limbs (DEC1 with q > 256: 2)
  int nl;                 // N limbs (ENC with q > 256: 2)
  int nchunks;            // accumulator chunks per part; chunk c computes output columns [col0[c], col0[c+1])
  int col0[kMaxChunks + 1];
  int w0;                 // width of chunk 0: chunks of this width load B through tmapB, the others through tmapB2
  int with_hi, q;
  uint32_t qmask;
  size_t B;
  int npairs;             // 256-row pair tiles
  int nA, nB;             // shared-memory slots for A and stages for B
  int a_resident;         // A slots hold the whole tile (loaded once per tile)
  int a_tma;              // DEC1, streaming A: raw e atoms arrive by TMA in an A slot pair and are transformed in place
  int nM, nS;             // message slots (ENC), store-staging slots
  int a_rel;              // commits that free an A slot: 2 (both MMA issuers) for resident A, else 1
  int two_issuers;        // second MMA issuer warp enabled (needs slices per chunk < B ring stages)
  int one_group;          // every epilogue warp drains every phase (instead of two groups, one per TMEM buffer)
  int debug_flags;        // NTRU_TRACE builds only (NTRU_DEBUG_NOB: bit 0 = skip the B operand loads)
  int mat_rows;           // rows of the key matrix (3 parts * nl * covered columns); K block kb starts at row kb * mat_rows
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

#include "umma_pair.cuh"

// ---- key matrix ---------------------------------------------------------------------------------
// Mat[row][kb]: row = part * (nl * T) + nl * col0[c] + ln * w_c + j   (part 0 = cyc, 1 = hi, 2 = lo; output k = col0[c] + j,
//               w_c = width of chunk c, T = col0[nchunks] = covered columns),   kb = lk * Kp + i.
struct ChunkTable {
  int nchunks;
  int col0[kMaxChunks + 1];
};

__global__ void k_build_keymat(int mode, int N, int Kp, int kl, int nl, const ChunkTable ct, const void *poly, uint8_t *mat) {
  const int klen = kl * Kp;
  const int T = ct.col0[ct.nchunks];
  const size_t rows_total = (size_t)3 * nl * T;
  const size_t total = rows_total * klen;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int kb = (int)(idx % klen);
    const int row = (int)(idx / klen);
    const int part = row / (nl * T);
    const int rin = row % (nl * T);               // nl * col0[c] + ln * w_c + j
    int c = 0;
    while (c + 1 < ct.nchunks && rin >= nl * ct.col0[c + 1]) ++c;
    const int w = ct.col0[c + 1] - ct.col0[c];
    const int off = rin - nl * ct.col0[c];
    const int ln = off / w, j = off % w;
    const int k = ct.col0[c] + j;
    const int lk = kb / Kp, i = mode == DEC1F ? (kb % Kp) >> 1 : kb % Kp;   // DEC1F: two bytes per coefficient
    int coef = 0;
    bool nz = k < N && i < N;
    int src = 0;
    if (nz) {
      if (part == PH_HI) {
        nz = i > k;
        src = k + N - i;
      } else {
        nz = part == PH_CYC || i <= k;
        src = k - i;
        if (src < 0) src += N;
      }
    }
    if (nz) {
      if (mode == ENC) coef = reinterpret_cast<const uint16_t *>(poly)[src];
      else if (mode == DEC1 || mode == DEC1F) coef = reinterpret_cast<const int8_t *>(poly)[src];
      else coef = reinterpret_cast<const uint8_t *>(poly)[src];
    }
    uint8_t out;
    if (mode == ENC) out = ln == 0 ? (uint8_t)(coef & 0xff) : (uint8_t)(coef >> 8);
    else if (mode == DEC1) out = (uint8_t)(int8_t)(lk == 0 ? coef : coef * 64);
    else if (mode == DEC1F) out = (kb & 1) ? (coef > 0 ? 0x3C : (coef < 0 ? 0xBC : 0)) : 0;   // fp16 +-1.0 = 0x3C00 / 0xBC00
    else out = (uint8_t)coef;
    // tile-major storage: for each 128-byte K block all rows are contiguous (128-byte pitch), so that the box
    // of one pipeline slice (half a chunk's rows x 128 B) is one contiguous run of global memory
    mat[((size_t)(kb / kAtomK) * rows_total + (size_t)row) * kAtomK + (kb % kAtomK)] = out;
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

// Chunk table of one product.  g = the epilogue's column granularity (every warp of a TMEM lane quadrant owns a whole
// number of store passes per chunk): ENC 64 (two-unit passes, two warps per quadrant and group), 32 with one-unit
// passes (N > 512); DEC1 32; DEC2 128.  The covered columns T = N rounded up to g are split into
// ceil(T / max_out) chunks whose widths differ by at most g (wider chunks first): at most two distinct widths.
// max_out: a 256-column accumulator holds 256 / nl outputs; ENC is capped at 128 so that a chunk's message bytes are
// one 128-byte TMA atom per row.
// Which kernels run their epilogue as ONE group (every epilogue warp drains every phase) instead of two groups, one per
// TMEM buffer.  One group hands a buffer back after half the time but pays the fixed cost of a phase in every warp; it
// doubles the column granularity of the chunk table.  Measured (profiles/r2_one_group.txt): it pays for the byte-limb
// first decrypt product with its eight epilogue warps (DEC1 at N = 509 0.841 -> 0.814 ms, N = 167 0.311 -> 0.297) and,
// marginally, for DEC2 up to N = 512 (0.347 -> 0.342); ENC loses 3-15 % at every N.  NTRU_OPT_EPILOGUE overrides.
bool one_group_mode(const ntru_ctx *ctx, int mode) {
  if (ctx->opt_epilogue == 1) return false;
  if (ctx->opt_epilogue == 2) return true;
  const int N = ctx->N;
  if (mode == DEC1) return N <= 768;                                            // (N = 701: 1.535 -> 1.49 ms; N = 821: 1.99 -> 2.04)
  if (mode == DEC1F) return N <= 512;                                           // 0.836 -> 0.796 ms at N = 509; no change at N = 677
  if (mode == DEC2) return N <= 512 && (N + 255) / 256 * 256 == (N + 127) / 128 * 128;   // no extra accumulator columns
  return false;
}

void geometry(const ntru_ctx *ctx, int mode, int kl, int nl, KeyMatrix &km) {
  const int N = ctx->N;
  km.limbs = kl;
  km.nlimbs = nl;
  const int Kp = (((mode == DEC1F ? 2 * N : N) + kAtomK - 1) / kAtomK) * kAtomK;   // bytes of K per limb
  km.klen = kl * Kp;
  const int max_out = mode == ENC ? 128 : 256 / nl;
  const bool pu1 = mode == ENC && Kp / kAtomK >= 5;          // launch_product's choice of the one-unit ENC instantiation
  const int g = (mode == ENC ? (pu1 ? 32 : 64) : (mode == DEC1 ? 32 : (mode == DEC1F ? 64 : 128))) * (one_group_mode(ctx, mode) ? 2 : 1);
  const int T = ((N + g - 1) / g) * g;
  km.nchunks = (T + max_out - 1) / max_out;
  const int base = (T / km.nchunks / g) * g;
  const int wide = (T - base * km.nchunks) / g;              // this many chunks are g wider
  km.col0[0] = 0;
  for (int c = 0; c < km.nchunks; ++c) km.col0[c + 1] = km.col0[c] + base + (c < wide ? g : 0);
  km.w[0] = km.col0[1] - km.col0[0];
  km.w[1] = km.col0[km.nchunks] - km.col0[km.nchunks - 1];
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
  const int T = km.col0[km.nchunks];
  const int rows = 3 * nl * T;
  const int Kp = km.klen / kl;
  const size_t bytes = (size_t)rows * km.klen;
  NTRU_CUDA(ctx, km.mat.reserve(bytes));
  ChunkTable ct;
  ct.nchunks = km.nchunks;
  for (int c = 0; c <= kMaxChunks; ++c) ct.col0[c] = c <= km.nchunks ? km.col0[c] : 0;
  {
    LaunchTimer timer(ctx, NTRU_K_OTHER);
    k_build_keymat<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(mode, ctx->N, Kp, kl, nl, ct, poly, (uint8_t *)km.mat.ptr);
  }
  NTRU_CUDA(ctx, cudaGetLastError());
  // one tensor map per distinct chunk width: the box is the half of a chunk's B rows that one CTA of the pair loads
  const uint64_t tiled_rows = (uint64_t)rows * (km.klen / kAtomK);
  for (int i = 0; i < 2; ++i) {
    int rc = encode_2d(ctx, km.tmap_half[i], km.mat.ptr, kAtomK, tiled_rows, kAtomK, kAtomK, (uint32_t)(nl * km.w[i]) / 2);
    if (rc) return rc;
  }
  NTRU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  km.ready = true;
  return NTRU_OK;
}

// The phase list of one tile (see struct Phase).  lohi = false: the round-1 order [CYC of every chunk, HI of every chunk].
void build_schedule(UmmaArgs &a, bool lohi) {
  auto hi_a0 = [&](int c) {   // first K atom that can hold a non-zero entry of the hi matrix for chunk c: i >= col0[c] + 1
    const int v = (a.col0[c] + 1) >> a.ea_shift;
    return v < a.atoms ? v : a.atoms - 1;
  };
  auto lo_a1 = [&](int c) {   // one past the last K atom of the lo matrix: i <= k <= min(col0[c + 1], N) - 1
    const int last = (a.col0[c + 1] < a.N ? a.col0[c + 1] : a.N) - 1;
    return (last >> a.ea_shift) + 1;
  };
  a.nph = 0;
  auto push = [&](int c, int kind) {
    Phase &p = a.ph[a.nph++];
    p.c = (int8_t)c; p.kind = (int8_t)kind;
    p.a0 = (int8_t)(kind == PH_HI ? hi_a0(c) : 0);
    p.a1 = (int8_t)(kind == PH_LO ? lo_a1(c) : a.atoms);
    p.first = p.rel = 0;
  };
  if (!a.with_hi) {
    for (int c = 0; c < a.nchunks; ++c) push(c, PH_CYC);
  } else if (!lohi) {
    for (int c = 0; c < a.nchunks; ++c) push(c, PH_CYC);
    for (int c = 0; c < a.nchunks; ++c) push(c, PH_HI);
  } else {
    // from the last chunk down: the phases at the end of a tile read the low atoms only, so the high atoms of a resident
    // A operand are released early -- and the next tile starts with the chunks that need exactly those
    int c = a.nchunks - 1;
    if (a.nchunks & 1) { push(c, PH_CYC); push(c, PH_HI); --c; }
    for (; c >= 1; c -= 2) { push(c, PH_HI); push(c - 1, PH_HI); push(c, PH_LO); push(c - 1, PH_LO); }
  }
  uint32_t seen = 0;
  int no = 0;
  for (int j = 0; j < a.nph; ++j)
    for (int at = a.ph[j].a0; at < a.ph[j].a1; ++at)
      if (!(seen >> at & 1u)) { seen |= 1u << at; a.ph[j].first |= (uint16_t)(1u << at); if (no < 16) a.a_order[no++] = (int8_t)at; }
  seen = 0;
  for (int j = a.nph - 1; j >= 0; --j)
    for (int at = a.ph[j].a0; at < a.ph[j].a1; ++at)
      if (!(seen >> at & 1u)) { seen |= 1u << at; a.ph[j].rel |= (uint16_t)(1u << at); }
}

template <int MODE>
int launch_product(ntru_ctx *ctx, const KeyMatrix &km, UmmaArgs &a, const void *a_bytes) {
  if (!(ctx->umma_attr_set & (1 << MODE))) {      // per context: function attributes are per device
    NTRU_CUDA(ctx, cudaFuncSetAttribute(k_umma_pair<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
    ctx->umma_attr_set |= 1 << MODE;
  }
  a.N = ctx->N; a.P = ctx->P; a.kl = km.limbs; a.nl = km.nlimbs; a.Kp = km.klen / km.limbs; a.atoms = a.Kp / kAtomK;
  a.ea = MODE == DEC1F ? kAtomK / 2 : kAtomK;
  a.ea_shift = MODE == DEC1F ? 6 : 7;
  a.k_last = (ctx->N - (a.atoms - 1) * a.ea + a.ea / 4 - 1) / (a.ea / 4);
  a.nchunks = km.nchunks; a.q = ctx->q; a.qmask = (uint32_t)ctx->q - 1;
  for (int c = 0; c <= kMaxChunks; ++c) a.col0[c] = c <= km.nchunks ? km.col0[c] : km.col0[km.nchunks];
  a.w0 = km.w[0];
  a.mat_rows = 3 * a.nl * a.col0[a.nchunks];
  build_schedule(a, ctx->opt_lohi == 0);
  a.npairs = (int)((a.B + 2 * kTileRows - 1) / (2 * kTileRows));
  a.nS = 2;
  a.nM = MODE == ENC ? 2 : 0;
  const int a_slots = a.atoms * a.kl;
  // ENC above N = 512: k_umma_pair<ENC, 0, 1> stages one unit per TMA store (one staging slot) and reads the message
  // bytes from global memory (no message slots).  The three slots go to the B ring, and the A operand stays resident up
  // to N = 1024: with 4 stages the ring covered ~2000 cycles of TMA latency at 900 cycles per slice (clock trace,
  // NTRU_DEBUG_NOB timing), and streaming A doubled the L2 -> SM traffic at N = 821.  (geometry() makes the same choice:
  // the chunk widths of ENC are multiples of 32 only with one-unit passes.)
  const bool pu1 = MODE == ENC && a.atoms >= 5;
  if (pu1) { a.nS = 1; a.nM = 0; }
  const int avail = kPairSlots - a.nS - a.nM;
  if (a_slots + 4 <= avail) {
    a.a_resident = 1; a.nA = a_slots; a.nB = avail - a_slots;
  } else {
    a.a_resident = 0; a.nA = 6; a.nB = avail - 6;
  }
  // streaming DEC1: the raw e atoms (two A slots each) are prefetched by TMA, four atoms deep
  a.a_tma = MODE == DEC1 && !a.a_resident && a.kl == 2 && avail >= 11;
  if (a.a_tma) { a.nA = 8; a.nB = avail - 8; }
  // resident A slots are read by both MMA issuer warps (alternating chunks) whenever a tile has >= 2 chunks
  a.two_issuers = 0;   // a second issuer warp was tried: the B ring is too short for two chunks in flight
  a.a_rel = (a.a_resident && a.two_issuers) ? 2 : 1;
  a.one_group = one_group_mode(ctx, MODE) ? 1 : 0;
  CUtensorMap tmB, tmB2, tmA, tmM, tmO[3];
  memcpy(&tmB, km.tmap_half[0], sizeof tmB);
  memcpy(&tmB2, km.tmap_half[1], sizeof tmB2);
  memset(&tmA, 0, sizeof tmA);
  memset(&tmM, 0, sizeof tmM);
  memset(tmO, 0, sizeof tmO);
  const uint64_t N = (uint64_t)ctx->N, P = (uint64_t)ctx->P;
  int rc;
  if (MODE == DEC1F) {
    // the caller's uint16 rows ARE the fp16 operand: 64 coefficients (128 bytes) x 128 rows per slot, SWIZZLE_128B
    rc = encode_2d_ex(ctx, &tmA, const_cast<void *>(a.a_src), 2, N, (uint64_t)a.B, P * 2, kAtomK / 2, kTileRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else if (MODE != DEC1) {
    // byte rows straight into the UMMA layout: inner extent N (columns beyond read as zero), row pitch P
    rc = encode_2d(ctx, &tmA, const_cast<void *>(a_bytes), N, (uint64_t)a.B, P, kAtomK, kTileRows);
    if (rc) return rc;
  } else if (a.a_tma) {
    // raw uint16 rows, one atom = 128 coefficients x 128 rows = 32 KB, row-major (256-byte rows) in shared memory
    rc = encode_2d_ex(ctx, &tmA, const_cast<void *>(a.a_src), 2, P, (uint64_t)a.B, P * 2, kAtomK, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  if (MODE == ENC && !pu1) {
    rc = encode_2d(ctx, &tmM, const_cast<uint8_t *>(a.m), N, (uint64_t)a.B, P, kAtomK, kTileRows);
    if (rc) return rc;
  }
  {
    // outputs: per-warp tiles of 32 rows x 64 bytes (SWIZZLE_64B); DEC1's b is 32 rows x 32 bytes, unswizzled
    void *optr[3];
    int oelem[3], obox[3];
    CUtensorMapSwizzle oswz[3];
    if (MODE == ENC || MODE == DEC1 || MODE == DEC1F) {
      const int ob = pu1 ? 16 : 32;
      const CUtensorMapSwizzle osw = pu1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
      optr[0] = a.o16_cyc; oelem[0] = 2; obox[0] = ob; oswz[0] = osw;
      optr[2] = a.o16_hi; oelem[2] = 2; obox[2] = ob; oswz[2] = osw;
      if (MODE == ENC) { optr[1] = a.o16_cyc2; oelem[1] = 2; obox[1] = ob; oswz[1] = osw; }
      else if (MODE == DEC1) { optr[1] = a.o8_cyc; oelem[1] = 1; obox[1] = 32; oswz[1] = CU_TENSOR_MAP_SWIZZLE_NONE; }
      else optr[1] = nullptr;                       // DEC1F stores b with plain 16-byte global stores (no staging slot for it)
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
#ifdef NTRU_TRACE   // timing experiments exist in trace builds only: the shipped library reads no environment variable
    if (getenv("NTRU_DEBUG_NOSTORE")) a.out_mask = 0;   // results are not written
    if (getenv("NTRU_DEBUG_NOB")) a.debug_flags |= 1;
    if (getenv("NTRU_DEBUG_ONE_MMA")) a.debug_flags |= 2;
    if (getenv("NTRU_TRACE_SLICES")) a.debug_flags |= 4;
    if (getenv("NTRU_TRACE_LIGHT")) a.debug_flags |= 16;    // cycle sums of the issuer in registers instead of per-event stores
    if (getenv("NTRU_DEBUG_DOUBLE_MMA")) a.debug_flags |= 32;   // eight MMAs per slice: does the trip cost hide behind more tensor work?    // per-slice events of the issuer and the producer, tagged with the running slice number   // one of the four 32-byte MMA steps per slice: what does an issued slice cost without tensor work?
#endif
  }
  // The accumulator chunks cover col0[nchunks] >= N output columns (N rounded up to the epilogue's granularity); where
  // that stops short of the row pitch (N a multiple of the granularity: N = 512, 640, 768, 1024, ...) the pad columns
  // the kernel never sees are zeroed here so that every output row keeps the contract of include/ntru_b200.h.
  {
    const size_t cov = (size_t)a.col0[a.nchunks];
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
    LaunchTimer timer(ctx, MODE == ENC ? NTRU_K_ENC_TENSOR : (MODE == DEC2 ? NTRU_K_DEC2_TENSOR : NTRU_K_DEC1_TENSOR));
    const int clusters = a.npairs < ctx->sm_count / 2 ? a.npairs : ctx->sm_count / 2;
#ifdef NTRU_TRACE
    const int dbg = getenv("NTRU_DEBUG_EPI") ? atoi(getenv("NTRU_DEBUG_EPI")) : 0;
#define NTRU_DBG_LAUNCH(D)                                                                                                  \
  case D:                                                                                                                  \
    cudaFuncSetAttribute(k_umma_pair<MODE, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);          \
    k_umma_pair<MODE, D><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmB2, tmA, tmM, tmO[0], tmO[1], tmO[2]); \
    break;
    switch (pu1 ? 0 : dbg) {
      NTRU_DBG_LAUNCH(1) NTRU_DBG_LAUNCH(2) NTRU_DBG_LAUNCH(4) NTRU_DBG_LAUNCH(8) NTRU_DBG_LAUNCH(12) NTRU_DBG_LAUNCH(16)
      default:
        if (pu1) {
          cudaFuncSetAttribute(k_umma_pair<ENC, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);
          k_umma_pair<ENC, 0, 1><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmB2, tmA, tmM, tmO[0], tmO[1], tmO[2]);
        } else {
          k_umma_pair<MODE><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmB2, tmA, tmM, tmO[0], tmO[1], tmO[2]);
        }
    }
#else
    if (pu1) {
      if (!(ctx->umma_attr_set & 16)) {
        NTRU_CUDA(ctx, cudaFuncSetAttribute(k_umma_pair<ENC, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
        ctx->umma_attr_set |= 16;
      }
      k_umma_pair<ENC, 0, 1><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmB2, tmA, tmM, tmO[0], tmO[1], tmO[2]);
    } else {
      k_umma_pair<MODE><<<2 * clusters, kPairThreads, kPairSmemBytes, ctx->stream>>>(a, tmB, tmB2, tmA, tmM, tmO[0], tmO[1], tmO[2]);
    }
#endif
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
  if ((size_t)prop.sharedMemPerBlockOptin < kPairSmemBytes) return NTRU_OK;
  if (!get_encode_fn()) return NTRU_OK;
  ctx->tensor_ok = true;
  return NTRU_OK;
}

int umma_prepare_public(ntru_ctx *ctx) {
  return build_keymat(ctx, ENC, 1, ctx->q > 256 ? 2 : 1, ctx->d_h.ptr, ctx->km_h);
}

int umma_prepare_private(ntru_ctx *ctx) {
  // 256 < q <= 2048: the fp16 form of the first product (a uint16 below 2048 is its own fp16 encoding, scaled by 2^-24)
  // measured on B200: faster above N = 512 (streamed A operand: 1.43 against 1.61 ms at N = 677, profiles/r2_dec1_fp16_form.jsonl)
  // and, with the one-group epilogue, up to it (0.796 against 0.819 ms at N = 509, profiles/r2_one_group.txt)
  const bool f16 = ctx->q > 256 && ctx->q <= 2048 && ctx->opt_dec1_form != 1;
  int rc = f16 ? build_keymat(ctx, DEC1F, 1, 1, ctx->d_f.ptr, ctx->km_f)
               : build_keymat(ctx, DEC1, ctx->q > 256 ? 2 : 1, 1, ctx->d_f.ptr, ctx->km_f);
  ctx->km_f.f16 = f16;
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
  int rc = ctx->km_f.f16 ? launch_product<DEC1F>(ctx, ctx->km_f, a, nullptr) : launch_product<DEC1>(ctx, ctx->km_f, a, nullptr);
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
