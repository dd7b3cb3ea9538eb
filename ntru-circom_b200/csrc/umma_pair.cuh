// 2-CTA variant of the tcgen05 schedule (included by umma_kernels.cu, inside its anonymous namespace).
//
// Two CTAs of a cluster (one SM pair) share every B slice: tcgen05.mma.cta_group::2 multiplies a 256-row
// batch tile (128 rows per CTA, each in its own shared memory and TMEM) against NC accumulator columns whose
// B rows are split between the two CTAs (NC/2 rows each).  Per ciphertext this halves both the L2 -> SM
// traffic and the shared-memory reads of the key matrix -- the v2 single-CTA kernel was bound by exactly that
// (ncu: lts throughput 65-75 % of peak, tensor pipe 34 %).
//
// The A operand (128 rows x K bytes per CTA) is kept RESIDENT in shared memory for the whole tile when it
// fits beside a B ring of at least 4 stages (ENC and DEC2 always, DEC1 up to N = 512): it is loaded (TMA) or built
// (DEC1 transform warps, from registers prefetched one atom ahead) once per tile instead of once per accumulator
// chunk.  DEC1 above N = 512 streams A: the raw uint16 atom lands by TMA in the slot pair of its two byte limbs
// (four atoms in flight) and the transform warps split it in place (a.a_tma).  ENC above N = 512 runs the PU = 1
// instantiation (one staging slot, message bytes from global memory): the B ring gets the three slots.
//
// Shared memory: 14 slots of 16 KB = nA A-slots + nB B-stages + nM message slots + nS staging slots.  Barriers (same offsets in both CTAs):
//   a_full[i], b_full[j]   : used in the LEADER CTA only; both CTAs' TMA loads / transform warps signal them
//   a_empty[i], b_empty[j] : in both CTAs, signalled by tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[b]           : in both CTAs (commit multicast);  tmem_empty[b]: leader only, all epilogue warps
// Only the leader's MMA warp issues MMAs; producer, transform and epilogue warps run in both CTAs.

#ifdef NTRU_TRACE
// debug timeline of cluster 0 / CTA 0: each traced thread appends (tag << 40 | clock) words to its own lane of a
// global buffer (one fire-and-forget store; the shared-memory layout stays the production one).
constexpr int kTraceLanes = 4, kTraceCap = 2048;   // MMA issuer 0, epilogue warp 0, producer, MMA issuer 1
__device__ unsigned long long g_trace[kTraceLanes * kTraceCap];
#define TRACE(role, ev, idx)                                                                      \
  do {                                                                                            \
    if (blockIdx.x == 0 && !(a.debug_flags & 16) && trace_n[role - 1] < kTraceCap) {              \
      g_trace[(role - 1) * kTraceCap + trace_n[role - 1]++] =                                     \
          ((unsigned long long)(((ev) << 12) | ((idx) & 0xfff)) << 40) | (clock64() & 0xffffffffffull); \
    }                                                                                             \
  } while (0)
#define TRACE_NS(role, ev, idx)                                                                   \
  do {                                                                                            \
    if (blockIdx.x == 0 && !(a.debug_flags & 16) && trace_n[role - 1] < kTraceCap) {              \
      unsigned long long _ns;                                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_ns));                                     \
      g_trace[(role - 1) * kTraceCap + trace_n[role - 1]++] =                                     \
          ((unsigned long long)(((ev) << 12) | ((idx) & 0xfff)) << 40) | (_ns & 0xffffffffffull); \
    }                                                                                             \
  } while (0)
#else
#define TRACE(role, ev, idx) do {} while (0)
#define TRACE_NS(role, ev, idx) do {} while (0)
#endif

constexpr int kSlotBytes = 16384;
constexpr int kPairSlots = 14;
constexpr int kPairBars = 4 * kPairSlots + 8;   // a_full/a_empty/b_full/b_empty[kPairSlots], tmem full/empty[2], m full/empty[2]
// Warp roles: 16 = B-ring TMA producer, 17 = MMA issuer, 18 = A / message producer.  Warps 0-15 are epilogue
// warps in two groups of 8, one per TMEM buffer (DEC1: 0-7 transform, 8-15 epilogue in two groups of 4); an
// epilogue warp reads TMEM lanes 32*(warp%4)...
constexpr int kPairProducerWarp = 16, kPairMmaWarp = 17, kPairAuxWarp = 18, kPairEpiWarp0Dec1 = 8;
constexpr int kPairThreads = 19 * 32;
constexpr size_t kPairSmemBytes = (size_t)kPairSlots * kSlotBytes + 1024 /*align*/ + 8 * kPairBars + 64;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA).  Default semantics (release at CTA
// scope) on purpose: a cluster-scope release makes ptxas emit MEMBAR.ALL.GPU + ERRBAR, which stalls the epilogue
// warps until their global stores drain (ncu, profiles/r1_ncu_pair_v1: 40 % of all stall samples); the data these
// barriers guard is ordered by fence.proxy.async / tcgen05.fence, not by the arrive itself.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of the pair; completion bytes are credited to `bar_cluster` (the leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster)
      : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  if (F16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrives on the barrier at this offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc_pair(int a_signed, int b_signed, int n) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}
// kind::f16: fp16 x fp16 (formats 0) into fp32 accumulators (format 1)
__host__ __device__ constexpr uint32_t make_idesc_pair_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// DBG (timing experiments, instantiated in NTRU_TRACE builds only; results are then wrong on purpose):
//   1 epilogue hands the buffers back without reading them     2 TMEM reads only (no arithmetic, no stores)
//   4 no TMEM reads (arithmetic and stores on register garbage) 8 no staging stores and no TMA stores
//   16 staging stores and fences, but no TMA stores
//
// PU = units (16 coefficients) a warp stages per TMA store.  ENC with PU = 1 halves the staging area (one slot instead of
// two): at 768 < N <= 896 that slot is what lets the A operand stay resident beside a 4-stage B ring (launch_product).
template <int MODE, int DBG = 0, int PU = (MODE == 2 ? 4 : 2)>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
k_umma_pair(const UmmaArgs a, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapB2,
            const __grid_constant__ CUtensorMap tmapA,
            const __grid_constant__ CUtensorMap tmapM, const __grid_constant__ CUtensorMap tmapO0,
            const __grid_constant__ CUtensorMap tmapO1, const __grid_constant__ CUtensorMap tmapO2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kPairSlots * kSlotBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kPairBars);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](uint32_t i) { return bar0 + 8u * i; };
  auto a_empty = [&](uint32_t i) { return bar0 + 8u * (kPairSlots + i); };
  auto b_full = [&](uint32_t i) { return bar0 + 8u * (2 * kPairSlots + i); };
  auto b_empty = [&](uint32_t i) { return bar0 + 8u * (3 * kPairSlots + i); };
  auto tfull_bar = [&](uint32_t b) { return bar0 + 8u * (4 * kPairSlots + b); };
  auto tempty_bar = [&](uint32_t b) { return bar0 + 8u * (4 * kPairSlots + 2 + b); };
  auto m_full = [&](uint32_t b) { return bar0 + 8u * (4 * kPairSlots + 4 + b); };
  auto m_empty = [&](uint32_t b) { return bar0 + 8u * (4 * kPairSlots + 6 + b); };
  const uint32_t smem_base = smem_u32(smem);
  // slot order: [A slots nA][B stages nB][message slots nM][store staging nS]
  auto a_slot = [&](uint32_t i) { return smem_base + i * kSlotBytes; };
  auto b_slot = [&](uint32_t j) { return smem_base + (a.nA + j) * kSlotBytes; };
  auto m_slot = [&](uint32_t j) { return smem_base + (a.nA + a.nB + j) * kSlotBytes; };
  const uint32_t stage_base = smem_base + (a.nA + a.nB + a.nM) * kSlotBytes;

#ifdef NTRU_TRACE
  int trace_n[kTraceLanes] = {0, 0, 0, 0};
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int kEpiWarps = MODE == DEC1 ? 8 : 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nA; ++i) {
      mbar_init(a_full(i), MODE == DEC1 ? 16 : 2);   // DEC1: 8 transform warps per CTA; else one producer per CTA
      mbar_init(a_empty(i), a.a_rel);                // tcgen05.commit (multicast) from each issuer that reads it
    }
    if (MODE == DEC1 && a.a_tma) {
      for (int j = 0; j < (a.nA >> 1); ++j) mbar_init(a_full(a.nA + j), 1);   // raw_full[j]: this CTA's TMA (+ bytes)
    }
    for (int j = 0; j < a.nB; ++j) {
      mbar_init(b_full(j), 2);                       // one producer arrival per CTA (+ transaction bytes)
      mbar_init(b_empty(j), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), a.one_group ? 2 * kEpiWarps : kEpiWarps);   // the epilogue warps that drain this buffer, both CTAs
      mbar_init(m_full(b), 1);                       // this CTA's producer (+ transaction bytes)
      mbar_init(m_empty(b), a.one_group ? kEpiWarps : kEpiWarps / 2);   // the epilogue warps that drain the chunk, this CTA
    }
    fence_barrier_init();
  }
  if (warp == kPairMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barrier addresses in the LEADER CTA, usable from either CTA
  auto lead = [&](uint32_t local_bar) { return mapa_u32(local_bar, 0); };

  if (warp == kPairProducerWarp) {
    // ===================== B producer (both CTAs): key-matrix slices through the ring =====================
    // Lean on purpose (uniform control flow for the whole warp, incremental ring counters, one elected lane
    // issues): a slice holds 512 cycles of MMA work, and a producer written over a generic slice iterator cost
    // ~900 cycles per slice, so the ring never filled.  It waits on nothing but the ring itself: A-operand and
    // message loads depend on epilogue progress and live in their own warp (below), otherwise a late epilogue
    // delays the B loads of the next chunk and MMA, epilogue and loads run one after the other (clock trace).
    const uint32_t lead_b_full = lead(b_full(0));
    const int part_rows = a.mat_rows / 3;                  // key-matrix rows of one part (cyc / hi / lo)
    uint32_t sb = 0, b_par = 0;
#ifdef NTRU_TRACE
    int gs = 0;   // running slice number (trace tag)
#endif
    if (lane == 0) { TRACE(3, 10, 0); TRACE_NS(3, 11, 0); }
    for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
      for (int j = 0; j < a.nph; ++j) {
        const Phase ph = a.ph[j];
        const int c = ph.c;
        const int wc = a.col0[c + 1] - a.col0[c];
        const int half_rows = (a.nl * wc) >> 1;            // B rows of this chunk that this CTA loads
        const uint32_t b_bytes = 2u * (uint32_t)half_rows * kAtomK;
        const CUtensorMap *bmap = wc == a.w0 ? &tmapB : &tmapB2;   // box rows = half_rows
        const int row0 = ph.kind * part_rows + a.nl * a.col0[c] + (int)rank * half_rows;
        for (int at = ph.a0; at < ph.a1; ++at) {
          for (int lk = 0; lk < a.kl; ++lk) {
#ifdef NTRU_TRACE
            if (lane == 0 && (a.debug_flags & 4)) TRACE(3, 0, gs);
#endif
            mbar_wait(b_empty(sb), b_par ^ 1);
#ifdef NTRU_TRACE
            if (lane == 0 && (a.debug_flags & 4)) TRACE(3, 1, gs);
            ++gs;
#endif
            if (elect_one()) {
#ifdef NTRU_TRACE
              if (a.debug_flags & 1) {   // timing experiment: no B traffic at all (operands are stale shared memory)
                if (leader) mbar_arrive(b_full(sb)); else mbar_arrive_cluster(lead_b_full + 8u * sb);
              } else
#endif
              {
                if (leader) mbar_arrive_expect_tx(b_full(sb), b_bytes);
                else mbar_arrive_cluster(lead_b_full + 8u * sb);
                tma_load_2d_pair(b_slot(sb), bmap, 0, (lk * a.atoms + at) * a.mat_rows + row0, lead_b_full + 8u * sb);
              }
            }
            __syncwarp();
            if (++sb == (uint32_t)a.nB) { sb = 0; b_par ^= 1; }
          }
        }
      }
    }
    if (lane == 0) { TRACE(3, 12, 0); TRACE_NS(3, 13, 0); }
  } else if (warp == kPairAuxWarp) {
    // ===================== A / message producer (both CTAs) =====================
    if (MODE == DEC1) {
      if (a.a_tma) {
        // streaming DEC1: the raw uint16 atom (128 rows x 256 bytes) lands in the A slot pair (2 pj, 2 pj + 1) that
        // will hold its two byte limbs; the transform warps rewrite it in place.  TMA keeps nA / 2 atoms in flight,
        // which the register prefetch of the transform warps (one atom) could not: a clock trace showed 2000-2500
        // cycles per atom, load latency, against the 1024 cycles of MMA work an atom feeds.
        const uint32_t np = (uint32_t)a.nA >> 1;
        uint32_t pj = 0, par = 0;
        for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
          const int a_row = T * 256 + (int)rank * kTileRows;
          for (int j = 0; j < a.nph; ++j) {
            const int at1 = a.ph[j].a1;
            for (int at = a.ph[j].a0; at < at1; ++at) {
              mbar_wait(a_empty(2 * pj), par ^ 1);
              mbar_wait(a_empty(2 * pj + 1), par ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(a_full(a.nA + pj), 2u * kABytes);
                tma_load_2d(a_slot(2 * pj), &tmapA, at * kAtomK, a_row, a_full(a.nA + pj));
              }
              __syncwarp();
              if (++pj == np) { pj = 0; par ^= 1; }
            }
          }
        }
      }
    } else {
      const uint32_t a_bytes = 2u * kABytes;
      const uint32_t lead_a_full = lead(a_full(0));
      const bool resident = a.a_resident != 0;
      uint32_t sas = 0, a_par_s = 0, t_par = 0, mc = 0;
      auto load_a = [&](uint32_t sa, uint32_t par, int at, int a_row) {
        mbar_wait(a_empty(sa), par ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(a_full(sa), a_bytes);
          else mbar_arrive_cluster(lead_a_full + 8u * sa);
          tma_load_2d_pair(a_slot(sa), &tmapA, at * a.ea, a_row, lead_a_full + 8u * sa);
        }
        __syncwarp();
      };
      for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1, t_par ^= 1) {
        const int a_row = T * 256 + (int)rank * kTileRows;
        if (resident) {                                       // the whole tile once, atoms in the order of their first use (kl == 1 here)
          for (int k = 0; k < a.atoms; ++k) load_a((uint32_t)a.a_order[k], t_par, a.a_order[k], a_row);
        }
        for (int j = 0; j < a.nph; ++j) {
          const Phase ph = a.ph[j];
          if (!resident) {
            for (int at = ph.a0; at < ph.a1; ++at) {
              load_a(sas, a_par_s, at, a_row);
              if (++sas == (uint32_t)a.nA) { sas = 0; a_par_s ^= 1; }
            }
          }
          if (MODE == ENC && PU != 1 && ph.kind != PH_HI) {   // the chunk's 128 message bytes per row, for this CTA's epilogue
            const uint32_t ms = mc & 1;
            mbar_wait(m_empty(ms), ((mc >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(m_full(ms), kABytes);
              tma_load_2d(m_slot(ms), &tmapM, a.col0[ph.c], a_row, m_full(ms));
            }
            __syncwarp();
            ++mc;
          }
        }
      }
    }
  } else if (warp == kPairMmaWarp) {
    // ===================== MMA issuers (leader CTA only): two warps, alternating chunks =====================
    // The whole warp runs the (uniform) control flow; one elected lane issues tcgen05.mma / commit.  A slice
    // holds 512 cycles of tensor work but costs one issuer ~600 cycles (mbarrier try_wait latency, ~85 SASS
    // instructions of uniform-datapath descriptor arithmetic, commits), so ONE issuer left the tensor pipe 58 %
    // busy.  Issuer w owns chunks cc == w (mod 2), i.e. TMEM buffer w; the B ring is consumed in order, each
    // stage by exactly one issuer; a resident A slot is freed when BOTH issuers have committed their last read
    // of it in the tile (a_empty count 2).  Phase parity only tells adjacent uses of a ring stage apart, so the
    // second issuer is enabled only when a chunk has fewer slices than the B ring has stages (a.two_issuers):
    // then no issuer can wait on a stage two phases ahead of its oldest unconsumed use.
    if (leader) {
      // One flat loop per chunk with running slot / stage / descriptor counters: the first version walked
      // (atom, K limb) in nested loops with run-time bounds and cost ~95 SASS instructions per slice on a single
      // warp (~600 cycles against the 512 cycles of tensor work a slice holds; clock trace, ncu source counters).
      // Slices of a chunk use consecutive resident A slots: sa = (first atom) * kl + i.
      const uint64_t desc_hi = make_smem_desc(0) & ~0x3FFFull;             // everything but the start address
      const uint32_t a_lo0 = smem_base >> 4, b_lo0 = (smem_base + a.nA * kSlotBytes) >> 4;   // 16-byte units
      const bool resident = a.a_resident != 0;
      const uint32_t nB = (uint32_t)a.nB, nA = (uint32_t)a.nA;
      const uint32_t bfull0 = b_full(0), bempty0 = b_empty(0), afull0 = a_full(0), aempty0 = a_empty(0);
      const uint32_t klast = (uint32_t)a.k_last, kl_u = (uint32_t)a.kl;
      uint32_t sb = 0, b_par = 0;           // B ring position / phase parity
      uint32_t sas = 0, a_par_s = 0;        // streaming A ring position / phase parity
      uint32_t cc = 0, t_par = 0;           // phase counter, resident-A phase parity (per tile)
#ifdef NTRU_TRACE
      int gs = 0;   // running slice number (trace tag of the per-slice events when debug flag 4 is set)
      // debug flag 16: light profile -- no per-event stores, only cycle sums in registers (three 32-bit clock reads per slice)
      const bool light = (a.debug_flags & 16) != 0;
      unsigned long long lp_tmem = 0, lp_ops = 0, lp_issue = 0, lp_other = 0, lp_slices = 0, lp_phases = 0;
      uint32_t lp_t = light ? (uint32_t)clock() : 0u;
#define LP_ADD(acc) do { if (light) { const uint32_t now_ = (uint32_t)clock(); acc += now_ - lp_t; lp_t = now_; } } while (0)
      const long long lp_start = clock64();
#else
#define LP_ADD(acc) do {} while (0)
#endif
      for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1, t_par ^= 1) {
        for (int j = 0; j < a.nph; ++j, ++cc) {
          const Phase ph = a.ph[j];
          const uint32_t nsl = (uint32_t)(ph.a1 - ph.a0) * kl_u;             // slices of this phase
          const int wc = a.col0[ph.c + 1] - a.col0[ph.c];
          const uint32_t idesc = MODE == DEC1F ? make_idesc_pair_f16(wc) : make_idesc_pair(0, MODE == DEC1 ? 1 : 0, a.nl * wc);
          // resident A: atoms read for the first time in the tile are waited for, atoms read for the last time are released
          const uint32_t first = resident ? ph.first : 0xffffu, rel = resident ? ph.rel : 0xffffu;
          const uint32_t buf = cc & 1;
          if (lane == 0) TRACE(1, 0, cc);
          LP_ADD(lp_other);
          mbar_wait(tempty_bar(buf), ((cc >> 1) & 1) ^ 1);
          LP_ADD(lp_tmem);
#ifdef NTRU_TRACE
          ++lp_phases;
#endif
          if (lane == 0) TRACE(1, 1, cc);
          const uint32_t d_tmem = tmem_base + buf * kAccCols;
          const uint32_t tfull = tfull_bar(buf);
          uint32_t sa = resident ? (uint32_t)ph.a0 * kl_u : sas;
          uint32_t at = (uint32_t)ph.a0, lk = 0;                             // K atom / K limb of the current slice
          uint32_t accumulate = ph.kind == PH_LO ? 1u : 0u;                  // a LO phase continues on the HI product in this buffer
          for (uint32_t i = 0; i < nsl; ++i) {
            // even lanes poll the B stage, odd lanes the A slot, in ONE try_wait round trip (the trace showed ~180
            // cycles per poll of an already complete barrier, paid once per barrier and slice in the streaming modes)
            LP_ADD(lp_other);
            if ((first >> at) & 1u) mbar_wait((lane & 1) ? afull0 + 8u * sa : bfull0 + 8u * sb, (lane & 1) ? (resident ? t_par : a_par_s) : b_par);
            else mbar_wait(bfull0 + 8u * sb, b_par);
            __syncwarp();
            LP_ADD(lp_ops);
#ifdef NTRU_TRACE
            if (lane == 0) TRACE(1, 3, (a.debug_flags & 4) ? gs : (int)cc);
#endif
            tc_fence_after();
            const uint64_t da = desc_hi | (uint64_t)(a_lo0 + sa * (kSlotBytes >> 4));
            const uint64_t db = desc_hi | (uint64_t)(b_lo0 + sb * (kSlotBytes >> 4));
            // the slices of the last K atom (one per K limb) hold coefficients below N only in their first k_last 32-byte steps
#ifdef NTRU_TRACE
            const uint32_t nk = (a.debug_flags & 2) ? 1u : (at == (uint32_t)(a.atoms - 1) ? klast : 4u);   // timing experiment: one MMA per slice
#else
            const uint32_t nk = at == (uint32_t)(a.atoms - 1) ? klast : 4u;
#endif
            if (elect_one()) {
              umma_pair<MODE == DEC1F>(d_tmem, da, db, idesc, accumulate);
              if (nk > 1) umma_pair<MODE == DEC1F>(d_tmem, da + 2, db + 2, idesc, 1u);
              if (nk > 2) umma_pair<MODE == DEC1F>(d_tmem, da + 4, db + 4, idesc, 1u);
              if (nk > 3) umma_pair<MODE == DEC1F>(d_tmem, da + 6, db + 6, idesc, 1u);
#ifdef NTRU_TRACE
              if (a.debug_flags & 32) {   // timing experiment: the four MMAs of the slice once more (twice the tensor work per trip, results wrong)
                umma_pair<MODE == DEC1F>(d_tmem, da, db, idesc, 1u);
                umma_pair<MODE == DEC1F>(d_tmem, da + 2, db + 2, idesc, 1u);
                umma_pair<MODE == DEC1F>(d_tmem, da + 4, db + 4, idesc, 1u);
                umma_pair<MODE == DEC1F>(d_tmem, da + 6, db + 6, idesc, 1u);
              }
#endif
              umma_commit_pair(bempty0 + 8u * sb);
              if ((rel >> at) & 1u) umma_commit_pair(aempty0 + 8u * sa);
              if (i == nsl - 1) umma_commit_pair(tfull);
            }
            __syncwarp();
            LP_ADD(lp_issue);
#ifdef NTRU_TRACE
            if (lane == 0) TRACE(1, 5, (a.debug_flags & 4) ? gs : (int)cc);
            ++gs;
            ++lp_slices;
#endif
            accumulate = 1;
            if (++sb == nB) { sb = 0; b_par ^= 1; }
            if (++lk == kl_u) { lk = 0; ++at; }
            if (resident) {
              ++sa;
            } else {
              if (++sas == nA) { sas = 0; a_par_s ^= 1; }
              sa = sas;
            }
          }
        }
      }
#ifdef NTRU_TRACE
      if (light && blockIdx.x == 0 && lane == 0) {     // lane 0 of the trace buffer: the light profile of this launch
        g_trace[0] = 0x4c50ull << 48;                   // "LP"
        g_trace[1] = (unsigned long long)(clock64() - lp_start);
        g_trace[2] = lp_tmem; g_trace[3] = lp_ops; g_trace[4] = lp_issue; g_trace[5] = lp_other; g_trace[6] = lp_slices; g_trace[7] = lp_phases;
      }
#endif
  }
  } else if (MODE == DEC1 && warp < kPairEpiWarp0Dec1) {
    // ===================== DEC1 transform (both CTAs): e (uint16, global) -> byte-limb A slots =====
    const int t = threadIdx.x;                           // 0..255 (warps 0-7)
    const int chunk = t & 7;                             // 16-byte chunk of the 128-byte A row
    const int r0 = t >> 3;                               // rows r0, r0+32, r0+64, r0+96
    if (a.a_tma) {
      // in place: raw atom (row r at byte 256 r of the slot pair) -> limb 0 in slot 2 pj, limb 1 in slot 2 pj + 1.
      // Lanes 0-3 / 4-7 of a row read the halves of their 32 bytes in opposite order (no bank conflict).
      const uint32_t np = (uint32_t)a.nA >> 1;
      const uint32_t lead_a_full = lead(a_full(0));
      const uint32_t sel = (uint32_t)(chunk >> 2) & 1u;
      uint32_t pj = 0, par = 0;
      int xn = 0;   // atoms built so far (trace tag)
      int per_tile = 0;                                   // atom visits per tile: every phase streams its own atoms
      for (int j = 0; j < a.nph; ++j) per_tile += a.ph[j].a1 - a.ph[j].a0;
      for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
            for (int v = 0; v < per_tile; ++v, ++xn) {
              if (t == 0) TRACE(4, 0, xn);
              mbar_wait(a_full(a.nA + pj), par);
              if (t == 0) TRACE(4, 1, xn);
              const uint32_t raw0 = a_slot(2 * pj);
              uint4 w[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t base = raw0 + (uint32_t)(r0 + 32 * j) * 256u + (uint32_t)chunk * 32u;
                const uint4 x = lds128(base + sel * 16u), y = lds128(base + (sel ^ 1u) * 16u);
                w[2 * j] = sel ? y : x;
                w[2 * j + 1] = sel ? x : y;
              }
              asm volatile("bar.sync 1, 256;" ::: "memory");   // every transform thread has read its part of the raw atom
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int r_in = r0 + 32 * j;
                const uint32_t off = (uint32_t)((r_in >> 3) * 1024 + (r_in & 7) * 128 + ((chunk ^ (r_in & 7)) << 4));
                const uint4 w0 = w[2 * j], w1 = w[2 * j + 1];
                sts128(raw0 + off, make_uint4(__byte_perm(w0.x, w0.y, 0x6420), __byte_perm(w0.z, w0.w, 0x6420),
                                              __byte_perm(w1.x, w1.y, 0x6420), __byte_perm(w1.z, w1.w, 0x6420)));
                sts128(raw0 + kSlotBytes + off,
                       make_uint4(__byte_perm((w0.x >> 6) & 0x00FC00FCu, (w0.y >> 6) & 0x00FC00FCu, 0x6420),
                                  __byte_perm((w0.z >> 6) & 0x00FC00FCu, (w0.w >> 6) & 0x00FC00FCu, 0x6420),
                                  __byte_perm((w1.x >> 6) & 0x00FC00FCu, (w1.y >> 6) & 0x00FC00FCu, 0x6420),
                                  __byte_perm((w1.z >> 6) & 0x00FC00FCu, (w1.w >> 6) & 0x00FC00FCu, 0x6420)));
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive_cluster(lead_a_full + 16u * pj);
                mbar_arrive_cluster(lead_a_full + 16u * pj + 8u);
              }
              if (t == 0) TRACE(4, 5, xn);
              if (++pj == np) { pj = 0; par ^= 1; }
            }
      }
    } else {
    const uint16_t *src = reinterpret_cast<const uint16_t *>(a.a_src);
    // work list: the atoms whose A slots must be (re)built, in pipeline order
    // (resident: the atoms of a tile once, in the order of their first use; streaming: the atoms of every phase)
    struct Item { int T, j, k, at; uint32_t ia; int titer; bool valid; };
    auto advance = [&](Item &w) {
      if (a.a_resident) {
        if (++w.k < a.atoms) { w.at = a.a_order[w.k]; return; }
      } else {
        w.ia += a.kl;
        if (++w.at < a.ph[w.j].a1) return;
        if (++w.j < a.nph) { w.at = a.ph[w.j].a0; return; }
      }
      w.T += gridDim.x >> 1;
      ++w.titer;
      w.valid = w.T < a.npairs;
      w.j = 0; w.k = 0;
      w.at = a.a_resident ? a.a_order[0] : a.ph[0].a0;
    };
    auto load_atom = [&](const Item &w, uint4 (&raw)[8]) {
      const int col = w.at * kAtomK + chunk * 16;        // first coefficient of this thread's chunk
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t row = (size_t)w.T * 256 + rank * kTileRows + r0 + 32 * j;
        const bool ok = w.valid && row < a.B && col < a.P;
        const uint4 *ptr = reinterpret_cast<const uint4 *>(src + (ok ? row * (size_t)a.P + col : 0));
        uint4 x0 = __ldg(ptr), x1 = __ldg(ptr + 1);
        if (!ok) x0 = x1 = make_uint4(0, 0, 0, 0);
        raw[2 * j] = x0;
        raw[2 * j + 1] = x1;
      }
    };
    Item cur;
    cur.T = blockIdx.x >> 1; cur.j = 0; cur.k = 0; cur.ia = 0; cur.titer = 0;
    cur.valid = cur.T < a.npairs;
    cur.at = a.a_resident ? a.a_order[0] : a.ph[0].a0;
    uint4 raw[8], raw_next[8];
    int xn = 0;   // atoms built so far (trace tag)
    if (cur.valid) load_atom(cur, raw);
    while (cur.valid) {
      Item nxt = cur;
      advance(nxt);
      load_atom(nxt, raw_next);                           // prefetch one atom ahead (zeros when !valid)
      uint32_t sa0, sa1, par0, par1;
      if (a.a_resident) {
        sa0 = cur.at * a.kl; sa1 = sa0 + 1; par0 = par1 = cur.titer & 1;
      } else {
        sa0 = cur.ia % a.nA; par0 = (cur.ia / a.nA) & 1;
        sa1 = (cur.ia + 1) % a.nA; par1 = ((cur.ia + 1) / a.nA) & 1;
      }
      if (t == 0) TRACE(4, 0, xn);
      mbar_wait(a_empty(sa0), par0 ^ 1);
      if (a.kl == 2) mbar_wait(a_empty(sa1), par1 ^ 1);
      if (t == 0) TRACE(4, 1, xn);
      uint8_t *dst0 = smem + (size_t)sa0 * kSlotBytes;
      uint8_t *dst1 = smem + (size_t)sa1 * kSlotBytes;
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
        mbar_arrive_cluster(lead(a_full(sa0)));
        if (a.kl == 2) mbar_arrive_cluster(lead(a_full(sa1)));
      }
      if (t == 0) TRACE(4, 5, xn);
      ++xn;
      cur = nxt;
#pragma unroll
      for (int j = 0; j < 8; ++j) raw[j] = raw_next[j];
    }
    }
  } else {
    // ===================== epilogue (both CTAs): TMEM -> registers -> shared staging -> TMA store ==========
    // Two GROUPS of epilogue warps, one per TMEM accumulator buffer: group g drains the chunks with cc & 1 == g, so
    // the two buffers are drained concurrently and every warp pays the fixed cost of a chunk (barrier waits, fences,
    // store issue) once per TWO chunks.  The first version put all warps on every chunk: ~275 instructions per
    // warp and chunk, 170 of them fixed cost, and the kernel was bound by the epilogue's instruction issue
    // (ncu: IPC 0.74 per scheduler; clock trace: 2450 cycles per chunk against 2048 cycles of MMA work or less).
    // Inside a group a warp owns 32 rows (its TMEM lane quadrant) and a contiguous run of 16-coefficient units;
    // results are staged in a per-warp shared-memory tile and written with cp.async.bulk.tensor stores (per-lane
    // row stores cost 32 cache lines per warp instruction).  ENC reads the message bytes from a TMA-loaded tile.
    // a.one_group: ALL epilogue warps drain every phase (half the columns per warp, the buffer is handed back after half
    // the time; every warp pays the fixed cost of every phase) -- for kernels whose epilogue is much longer than the
    // MMAs of a phase, where the hand-over chain of the two buffers, (M + E) / 2 per phase, is what a tile takes.
    const int kSub = (a.one_group ? kEpiWarps : kEpiWarps / 2) / 4;   // warps per TMEM lane quadrant that share a phase
    // ENC with PU = 1 (large N) also reads its message bytes from global memory instead of a TMA-loaded tile: no message
    // slots, so the B ring gets them (launch_product).
    constexpr bool kMsgGlobal = MODE == ENC && PU == 1;
    constexpr int kPassUnits = PU;                                 // units staged per TMA store (64-byte rows; 32-byte rows for PU = 1)
    const int ew = warp - (MODE == DEC1 ? kPairEpiWarp0Dec1 : 0);
    const int quad = warp & 3;
    const uint32_t grp = (uint32_t)(ew >> 2) & 1u;
    const int sub = a.one_group ? ew >> 2 : ew >> 3;
    const uint32_t Q2 = a.qmask | (a.qmask << 16);
    const uint32_t LA2 = (((uint32_t)a.q >> 1) - 1u) * 0x00010001u;   // x > q/2  <=>  bit log2(q) of x + q/2 - 1
    const int logq = 31 - __clz(a.q);
    constexpr uint32_t kRowBytes = (MODE == DEC2 ? 16u : 32u) * PU;   // staging row: 64 bytes (SWIZZLE_64B) or 32 (SWIZZLE_32B)
    static_assert(kRowBytes == 64 || kRowBytes == 32, "staging rows are one 64-byte or 32-byte swizzle span");
    const uint32_t stage = stage_base + (uint32_t)ew * (MODE == DEC1 ? 4096u : 32u * kRowBytes);   // this warp's staging tile
    // swizzled position of 16-byte chunk `ch` of this lane's staging row: ch ^ st_x
    const uint32_t st_row = stage + (uint32_t)lane * kRowBytes;
    const uint32_t st_x = kRowBytes == 64 ? (uint32_t)(lane >> 1) & 3u : (uint32_t)(lane >> 2) & 1u;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t m_row = (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128), m_x = (uint32_t)row_in_tile & 7u;
    const uint32_t t_addr0 = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tfull0 = tfull_bar(0), lead_tempty0 = lead(tempty_bar(0));
    uint32_t cc = 0, mc = 0;
    bool store_pending = false;
    for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
      const int out_row = T * 256 + (int)rank * kTileRows + quad * 32;
      {
        for (int jp = 0; jp < a.nph; ++jp, ++cc) {
          const int c = a.ph[jp].c;
          const int hi = a.ph[jp].kind == PH_HI;                    // quotient epilogue; CYC and LO phases end in the remainder epilogue
          const uint32_t ms = mc & 1, m_par = (mc >> 1) & 1;
          if (MODE == ENC && !hi) ++mc;
          if (!a.one_group && (cc & 1u) != grp) continue;
          const uint32_t buf = cc & 1u;
          const uint32_t t_addr = t_addr0 + buf * kAccCols, my_tfull = tfull0 + 8u * buf, lead_tempty = lead_tempty0 + 8u * buf;
          const int col0c = a.col0[c], wc = a.col0[c + 1] - col0c;
          const int upw = (wc >> 4) / kSub;                        // units per warp in this chunk
          const int npass = upw / kPassUnits;
          if (lane == 0 && ew == 0) TRACE(2, 0, cc);
          uint4 mg[4];                                            // kMsgGlobal: this lane's message bytes of the chunk
          if (kMsgGlobal && !hi) {                                // (in flight while the accumulators are still being computed)
            const size_t grow = (size_t)(out_row + lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = col0c + (sub * upw + j) * 16;
              mg[j] = (j < upw && grow < a.B && col < a.N) ? __ldg(reinterpret_cast<const uint4 *>(a.m + grow * (size_t)a.P + col))
                                                          : make_uint4(0, 0, 0, 0);
            }   // (four independent loads, nothing waits for them here: the unit that straddles N is masked at its use)
          }
          mbar_wait(my_tfull, (cc >> 1) & 1);
          tc_fence_after();
          if (lane == 0 && ew == 0) TRACE(2, 1, cc);
          if (MODE == ENC && !kMsgGlobal && !hi) mbar_wait(m_full(ms), m_par);
          if (lane == 0 && ew == 0) TRACE(2, 3, cc);
          if (DBG & 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_cluster(lead_tempty);
              if (MODE == ENC && !kMsgGlobal && !hi) mbar_arrive(m_empty(ms));
            }
            continue;
          }
          // passes of this warp that hold at least one column below the pitch (the last chunk of a tile is padding
          // beyond it: 5 of 8 units at N = 167, 677, 4 of 8 at N = 821); the others would be clipped by the TMA store
          int npe = (a.P - (col0c + sub * upw * 16) + kPassUnits * 16 - 1) / (kPassUnits * 16);
          npe = npe < 0 ? 0 : (npe > npass ? npass : npe);
          if (npe == 0) {                                           // nothing to drain: hand the buffers back (in phase)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_cluster(lead_tempty);
              if (MODE == ENC && !kMsgGlobal && !hi) mbar_arrive(m_empty(ms));
            }
            continue;
          }
          for (int ps = 0; ps < npe; ++ps) {
            const int u0 = sub * upw + ps * kPassUnits;
            const bool last_pass = ps == npe - 1;
            uint32_t res[kPassUnits * (MODE == DEC2 ? 4 : 8)];      // packed results of this pass
            uint32_t bres[kPassUnits * 4];                          // DEC1: lifted polynomial b (bytes)
            uint4 mm[kPassUnits];                                   // ENC: message bytes of the pass
            if (kMsgGlobal) {
              if (!hi) {
                mm[0] = ps == 0 ? mg[0] : (ps == 1 ? mg[1] : (ps == 2 ? mg[2] : mg[3]));   // PU == 1: pass ps = unit ps
                const int keep = a.N - (col0c + u0 * 16);   // < 16 for the unit that straddles N: the pad bytes of the
                if (keep < 16) {                                // caller's row do not count (a TMA tile reads them as zero)
                  uint32_t w[4] = {mm[0].x, mm[0].y, mm[0].z, mm[0].w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    const int kb = keep - 4 * t;
                    w[t] = kb >= 4 ? w[t] : (kb <= 0 ? 0u : (w[t] & ((1u << (8 * kb)) - 1u)));
                  }
                  mm[0] = make_uint4(w[0], w[1], w[2], w[3]);
                }
              }
            } else if (MODE == ENC && !hi) {
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) mm[j] = lds128(m_slot(ms) + m_row + ((((uint32_t)(u0 + j)) ^ m_x) << 4));
            }
            // all accumulator reads of the pass first (one round trip); after the last pass's reads the TMEM buffer
            // goes back to the MMA issuer before any arithmetic
            uint32_t acc[kPassUnits][32];
            {
              uint32_t acc1[kPassUnits][32];
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                if (DBG & 4) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) { acc[j][i] = (uint32_t)(i + lane + ps); acc1[j][i] = (uint32_t)(i ^ lane); }
                  continue;
                }
                tmem_ld16(t_addr + (u0 + j) * 16, acc[j]);
                if (MODE == ENC && a.nl == 2) tmem_ld16(t_addr + wc + (u0 + j) * 16, acc1[j]);
              }
              if (!(DBG & 4)) tmem_ld_wait();
              if (lane == 0 && ew == 0) TRACE(2, 4, cc);
              if (MODE == ENC && a.nl == 2) {
#pragma unroll
                for (int j = 0; j < kPassUnits; ++j)
#pragma unroll
                  for (int i = 0; i < 16; ++i) acc[j][i] += acc1[j][i] << 8;
              }
              if (MODE == DEC1F) {
                // fp32 accumulator = S * 2^-24 with the integer |S| < 2^22: 0.75f + S ulps of the binade [0.5, 1) is the
                // bit pattern 0x3F400000 + S, whose low 16 bits are S mod 2^16 (two's complement for negative S)
#pragma unroll
                for (int j = 0; j < kPassUnits; ++j)
#pragma unroll
                  for (int i = 0; i < 16; ++i) acc[j][i] = __float_as_uint(__uint_as_float(acc[j][i]) + 0.75f);
              }
            }
            if (last_pass) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive_cluster(lead_tempty);
                if (MODE == ENC && !kMsgGlobal && !hi) mbar_arrive(m_empty(ms));   // message tile is in registers
              }
            }
            if (DBG & 2) {
              uint32_t x = 0;
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) x ^= acc[j][0] ^ acc[j][15];
              if (x == 0x12345678u) sts128(st_row, make_uint4(x, x, x, x));
              continue;
            }
#pragma unroll
            for (int j = 0; j < kPassUnits; ++j) {
              uint32_t (&w)[32] = acc[j];
              if (MODE == ENC || MODE == DEC1 || MODE == DEC1F) {
                uint32_t *pk = res + 8 * j;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) pk[jj] = __byte_perm(w[2 * jj], w[2 * jj + 1], 0x5410);
                if (hi) {
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) pk[jj] = ((~pk[jj] & Q2) + 0x00010001u) & Q2;
                } else if (MODE == ENC) {
                  const uint32_t mw[4] = {mm[j].x, mm[j].y, mm[j].z, mm[j].w};
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    const uint32_t mp = __byte_perm(mw[jj >> 1], 0u, (jj & 1) ? 0x4342 : 0x4140);
                    pk[jj] = __vadd2(pk[jj], mp) & Q2;     // VIADD.16x2: the lanes wrap mod 2^16, q divides it
                  }
                } else {
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) pk[jj] &= Q2;
                  if (a.o8_cyc) {
                    // multiplier of the second product: any byte congruent to b = (x + [x > q/2]) mod 3 (index.js:117)
                    // will do, the reduction mod 3 happens on the accumulators of DEC2.  64 = 1 (mod 3), so a value
                    // y is folded to (y & 63) + (y >> 6) = y - 63 (y >> 6) <= 63 + 128, two coefficients per word.
                    if (logq & 1) {
                      // q = 2 (mod 3) and q/2 = 1:  (x + q/2 - 1) mod q = x + q/2 - 1 - q [x > q/2] = x + [x > q/2] (mod 3):
                      // the comparison is the carry the mask drops.  Five instructions per word instead of eight.
#pragma unroll
                      for (int wd = 0; wd < 4; ++wd) {
                        uint32_t y2[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                          const uint32_t y = (pk[2 * wd + i] + LA2) & Q2;
                          y2[i] = y - 63u * ((y >> 6) & 0x00FF00FFu);
                        }
                        bres[4 * j + wd] = __byte_perm(y2[0], y2[1], 0x6420);
                      }
                    } else {
#pragma unroll
                      for (int wd = 0; wd < 4; ++wd) {
                        uint32_t y2[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                          const uint32_t x = pk[2 * wd + i];
                          const uint32_t y = x + (((x + LA2) >> logq) & 0x00010001u);      // <= q
                          y2[i] = y - 63u * ((y >> 6) & 0x00FF00FFu);
                        }
                        bres[4 * j + wd] = __byte_perm(y2[0], y2[1], 0x6420);
                      }
                    }
                  }
                }
              } else {   // DEC2: mod 3 (hi: -x mod 3 = 2x mod 3)
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                  uint32_t bb[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint32_t y = hi ? 2u * w[4 * wd + i] : w[4 * wd + i];
                    bb[i] = y - 3u * __umulhi(y, 0x55555556u);
                  }
                  res[4 * j + wd] = __byte_perm(__byte_perm(bb[0], bb[1], 0x0040), __byte_perm(bb[2], bb[3], 0x0040), 0x5410);
                }
              }
            }
            if (lane == 0 && ew == 0) TRACE(2, 5, cc);
            if (DBG & 8) {
              uint32_t x = 0;
#pragma unroll
              for (int j = 0; j < kPassUnits * (MODE == DEC2 ? 4 : 8); ++j) x ^= res[j];
              if (x == 0x12345678u) sts128(st_row, make_uint4(x, x, x, x));
              continue;
            }
            // the previous store must have finished reading the staging tile before it is overwritten
            if (store_pending) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
            if (MODE == DEC2) {
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j)
                sts128(st_row + ((((uint32_t)j) ^ st_x) << 4), make_uint4(res[4 * j], res[4 * j + 1], res[4 * j + 2], res[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                sts128(st_row + ((((uint32_t)(2 * j)) ^ st_x) << 4), make_uint4(res[8 * j], res[8 * j + 1], res[8 * j + 2], res[8 * j + 3]));
                sts128(st_row + ((((uint32_t)(2 * j + 1)) ^ st_x) << 4), make_uint4(res[8 * j + 4], res[8 * j + 5], res[8 * j + 6], res[8 * j + 7]));
              }
              if (MODE == DEC1 && !hi && a.o8_cyc) {   // b: 32-byte rows, no swizzle needed
#pragma unroll
                for (int j = 0; j < kPassUnits; ++j)
                  sts128(stage + 2048u + (uint32_t)lane * 32u + (uint32_t)j * 16u,
                         make_uint4(bres[4 * j], bres[4 * j + 1], bres[4 * j + 2], bres[4 * j + 3]));
              }
              if (MODE == DEC1F && !hi && a.o8_cyc) {
                // b goes straight to global memory, 16 bytes per unit and lane (whole 32-byte sectors per pass): a
                // staged tile for it would cost the third staging slot, i.e. a B-ring stage
                const size_t grow = (size_t)(out_row + lane);
                const int col = col0c + u0 * 16;
                if (grow < a.B) {
#pragma unroll
                  for (int j = 0; j < kPassUnits; ++j)
                    if (col + 16 * j < a.P)
                      *reinterpret_cast<uint4 *>(a.o8_cyc + grow * (size_t)a.P + col + 16 * j) =
                          make_uint4(bres[4 * j], bres[4 * j + 1], bres[4 * j + 2], bres[4 * j + 3]);
                }
              }
            }
            if (lane == 0 && ew == 0) TRACE(2, 6, cc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && ew == 0) TRACE(2, 7, cc);
            if (DBG & 16) continue;
            if (lane == 0) {
              const int col = col0c + u0 * 16;
              if (hi) {
                if (a.out_mask & 4) tma_store_2d(&tmapO2, stage, col, out_row);
              } else {
                if (a.out_mask & 1) tma_store_2d(&tmapO0, stage, col, out_row);
                if (MODE == DEC1) {
                  if (a.out_mask & 2) tma_store_2d(&tmapO1, stage + 2048, col, out_row);
                } else {
                  if (a.out_mask & 2) tma_store_2d(&tmapO1, stage, col, out_row);
                }
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
          }
          if (lane == 0 && ew == 0) TRACE(2, 2, cc);
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
  }

#ifdef NTRU_TRACE
  if (blockIdx.x == 0 && lane == 0 && !(a.debug_flags & 16)) {   // terminate the lanes this launch wrote
    const int role = warp == kPairProducerWarp ? 3 : (warp == kPairMmaWarp ? 1 : (warp == (MODE == DEC1 ? kPairEpiWarp0Dec1 : 0) ? 2 : (MODE == DEC1 && warp == 0 ? 4 : 0)));
    if (role && trace_n[role - 1] < kTraceCap) g_trace[(role - 1) * kTraceCap + trace_n[role - 1]] = 0ull;
  }
#endif
  tc_fence_before();
  cluster_sync_all();
  if (warp == kPairMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}
