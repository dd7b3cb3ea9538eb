// lo / hi split of the CTA-pair tcgen05 schedule (included by umma_kernels.cu after umma_pair.cuh).
//
// k_umma_pair accumulates, per output chunk, the full cyclic product (all K atoms) and then once more the strictly
// upper "hi" part (quotient): 1.5 N^2 MACs and as many key-matrix bytes out of L2, which is what bounds the kernel
// (DESIGN.md section 4).  Here one chunk of outputs keeps BOTH triangular halves of the linear product in one TMEM
// buffer -- columns [0,128) = lo[k] = sum_{i<=k} x[i] y[k-i], columns [128,256) = hi[k] = sum_{i>k} x[i] y[k+N-i] --
// and the epilogue forms remainder = lo + hi and quotient = -hi itself.  A K atom (128 values of i) strictly below
// the chunk's outputs only feeds lo, one strictly above only hi, the diagonal atom both:
//     diagonal atom : one slice, MMA N = 256 (B rows: CTA 0 the lo rows, CTA 1 the hi rows), 16 KB per CTA
//     other atoms   : half slices, MMA N = 128 into the lo or the hi columns, 8 KB per CTA, two per ring stage
// N^2 MACs and two thirds of the key-matrix traffic (N = 509 ENC: 20 instead of 26 slice-equivalents per tile).
// A chunk is 128 accumulator columns per half: 64 outputs with two N limbs (ENC, q > 256), else 128 outputs.
// The A operand must be resident (every chunk reads every atom); callers fall back to k_umma_pair otherwise.
// Ring, barriers, warp roles and the two epilogue groups are those of umma_pair.cuh.

// Key matrix of the split schedule: row = c * 256 + part * 128 + ln * NCo + j  (part 0 = lo, 1 = hi; output k = c NCo + j),
// K index kb = lk * Kp + i, stored tile-major per 128-byte K block like the pair kernel's matrix.
__global__ void k_build_keymat_split(int mode, int N, int Kp, int kl, int nl, int NCo, int nchunks, const void *poly, uint8_t *mat) {
  const int klen = kl * Kp;
  const size_t rows_total = (size_t)nchunks * 256;
  const size_t total = rows_total * klen;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int kb = (int)(idx % klen);
    const int row = (int)(idx / klen);
    const int c = row >> 8, part = (row >> 7) & 1, within = row & 127;
    const int ln = within / NCo, j = within % NCo;
    const int k = c * NCo + j;
    const int lk = kb / Kp, i = kb % Kp;
    int coef = 0;
    if (ln < nl && k < N && i < N) {
      const bool lo_term = i <= k;
      if (part == 0 ? lo_term : !lo_term) {
        const int src = lo_term ? k - i : k + N - i;
        if (mode == ENC) coef = reinterpret_cast<const uint16_t *>(poly)[src];
        else if (mode == DEC1) coef = reinterpret_cast<const int8_t *>(poly)[src];
        else coef = reinterpret_cast<const uint8_t *>(poly)[src];
      }
    }
    uint8_t out;
    if (mode == ENC) out = ln == 0 ? (uint8_t)(coef & 0xff) : (uint8_t)(coef >> 8);
    else if (mode == DEC1) out = (uint8_t)(int8_t)(lk == 0 ? coef : coef * 64);
    else out = (uint8_t)coef;
    mat[((size_t)(kb / kAtomK) * rows_total + (size_t)row) * kAtomK + (kb % kAtomK)] = out;
  }
}

constexpr int kSplitW = 128;      // accumulator columns of one half (lo or hi) of a chunk

// ring stage t of a chunk: t < kl is the diagonal atom (K limb t); later stages hold up to two half slices, entry e
// of the chunk's list = (atom index into the non-diagonal atoms, K limb)
struct SplitEntry {
  int at, lk, hi;     // K atom, K limb, 1 = feeds the hi columns
};
__device__ __forceinline__ SplitEntry split_entry(int e, int kl, int d) {
  SplitEntry s;
  s.lk = kl == 2 ? (e & 1) : 0;
  const int ai = kl == 2 ? (e >> 1) : e;
  s.at = ai < d ? ai : ai + 1;
  s.hi = s.at > d;
  return s;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
k_umma_split(const UmmaArgs a, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapA,
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int kEpiWarps = MODE == DEC1 ? 8 : 16;
  const int kl = a.kl, atoms = a.atoms, nchunks = a.nchunks, NCo = a.NCo;
  const int nhalf = (atoms - 1) * kl;                   // half slices per chunk
  const int nst = kl + ((nhalf + 1) >> 1);              // ring stages per chunk

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nA; ++i) {
      mbar_init(a_full(i), MODE == DEC1 ? 16 : 2);
      mbar_init(a_empty(i), 1);
    }
    for (int j = 0; j < a.nB; ++j) {
      mbar_init(b_full(j), 2);
      mbar_init(b_empty(j), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kEpiWarps);
      mbar_init(m_full(b), 1);
      mbar_init(m_empty(b), kEpiWarps / 2);
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
  auto lead = [&](uint32_t local_bar) { return mapa_u32(local_bar, 0); };

  if (warp == kPairProducerWarp) {
    // ===================== B producer (both CTAs) =====================
    const uint32_t lead_b_full = lead(b_full(0));
    const uint32_t box_bytes = 64u * kAtomK;                                  // one TMA box: 64 rows
    uint32_t sb = 0, b_par = 0;
    for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
      for (int c = 0; c < nchunks; ++c) {
        const int d = (c * NCo) >> 7;
        const int crow = c * 256;
        for (int t = 0; t < nst; ++t) {
          mbar_wait(b_empty(sb), b_par ^ 1);
          if (elect_one()) {
            const uint32_t dst = b_slot(sb), bar = lead_b_full + 8u * sb;
            if (t < kl) {
              // diagonal atom: this CTA's 128 rows (CTA 0 the lo part, CTA 1 the hi part) as two boxes
              if (leader) mbar_arrive_expect_tx(b_full(sb), 4u * box_bytes); else mbar_arrive_cluster(bar);
              const int row = (t * atoms + d) * a.mat_rows + crow + (int)rank * 128;
              tma_load_2d_pair(dst, &tmapB, 0, row, bar);
              tma_load_2d_pair(dst + box_bytes, &tmapB, 0, row + 64, bar);
            } else {
              const int e0 = 2 * (t - kl);
              const int g = nhalf - e0 >= 2 ? 2 : 1;
              if (leader) mbar_arrive_expect_tx(b_full(sb), 2u * (uint32_t)g * box_bytes); else mbar_arrive_cluster(bar);
              for (int h = 0; h < g; ++h) {
                const SplitEntry s = split_entry(e0 + h, kl, d);
                const int row = (s.lk * atoms + s.at) * a.mat_rows + crow + s.hi * 128 + (int)rank * 64;
                tma_load_2d_pair(dst + (uint32_t)h * box_bytes, &tmapB, 0, row, bar);
              }
            }
          }
          __syncwarp();
          if (++sb == (uint32_t)a.nB) { sb = 0; b_par ^= 1; }
        }
      }
    }
  } else if (warp == kPairAuxWarp) {
    // ===================== A / message producer (both CTAs; ENC and DEC2 only) =====================
    if (MODE != DEC1) {
      const uint32_t a_bytes = 2u * kABytes;
      const uint32_t lead_a_full = lead(a_full(0));
      const uint32_t m_bytes = (uint32_t)kTileRows * (uint32_t)NCo;          // message tile: 128 rows x NCo bytes
      uint32_t t_par = 0, mc = 0;
      for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1, t_par ^= 1) {
        const int a_row = T * 256 + (int)rank * kTileRows;
        for (int at = 0; at < atoms; ++at) {                                  // resident A: once per tile
          mbar_wait(a_empty(at), t_par ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(a_full(at), a_bytes); else mbar_arrive_cluster(lead_a_full + 8u * at);
            tma_load_2d_pair(a_slot(at), &tmapA, at * kAtomK, a_row, lead_a_full + 8u * at);
          }
          __syncwarp();
        }
        if (MODE == ENC) {
          for (int c = 0; c < nchunks; ++c, ++mc) {
            const uint32_t ms = mc & 1;
            mbar_wait(m_empty(ms), ((mc >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(m_full(ms), m_bytes);
              tma_load_2d(m_slot(ms), &tmapM, c * NCo, a_row, m_full(ms));
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kPairMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc_full = make_idesc_pair(0, MODE == DEC1 ? 1 : 0, 256);
      const uint32_t idesc_half = make_idesc_pair(0, MODE == DEC1 ? 1 : 0, 128);
      const uint64_t desc_hi = make_smem_desc(0) & ~0x3FFFull;
      const uint32_t a_lo0 = smem_base >> 4, b_lo0 = (smem_base + a.nA * kSlotBytes) >> 4;   // 16-byte units
      uint32_t sb = 0, b_par = 0, cc = 0, t_par = 0;
      for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1, t_par ^= 1) {
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const int d = (c * NCo) >> 7;
          const bool first = c == 0, last_chunk = c == nchunks - 1;
          const uint32_t buf = cc & 1;
          mbar_wait(tempty_bar(buf), ((cc >> 1) & 1) ^ 1);
          const uint32_t d_tmem = tmem_base + buf * kAccCols;
          for (int t = 0; t < nst; ++t) {
            mbar_wait(b_full(sb), b_par);
            const bool last = t == nst - 1;
            const uint64_t db = desc_hi | (uint64_t)(b_lo0 + sb * (kSlotBytes >> 4));
            if (t < kl) {
              const uint32_t sa = (uint32_t)(d * kl + t);
              if (first) mbar_wait(a_full(sa), t_par);
              tc_fence_after();
              const uint64_t da = desc_hi | (uint64_t)(a_lo0 + sa * (kSlotBytes >> 4));
              if (elect_one()) {
                umma_i8_pair(d_tmem, da, db, idesc_full, t == 0 ? 0u : 1u);
                umma_i8_pair(d_tmem, da + 2, db + 2, idesc_full, 1u);
                umma_i8_pair(d_tmem, da + 4, db + 4, idesc_full, 1u);
                umma_i8_pair(d_tmem, da + 6, db + 6, idesc_full, 1u);
                umma_commit_pair(b_empty(sb));
                if (last_chunk) umma_commit_pair(a_empty(sa));
                if (last) umma_commit_pair(tfull_bar(buf));
              }
              __syncwarp();
            } else {
              const int e0 = 2 * (t - kl);
              const int g = nhalf - e0 >= 2 ? 2 : 1;
              const SplitEntry s0 = split_entry(e0, kl, d), s1 = split_entry(e0 + (g - 1), kl, d);
              const uint32_t sa0 = (uint32_t)(s0.at * kl + s0.lk), sa1 = (uint32_t)(s1.at * kl + s1.lk);
              if (first) {
                mbar_wait(a_full(sa0), t_par);
                if (g == 2) mbar_wait(a_full(sa1), t_par);
              }
              tc_fence_after();
              const uint64_t da0 = desc_hi | (uint64_t)(a_lo0 + sa0 * (kSlotBytes >> 4));
              const uint64_t da1 = desc_hi | (uint64_t)(a_lo0 + sa1 * (kSlotBytes >> 4));
              const uint64_t db1 = db + ((64 * kAtomK) >> 4);
              const uint32_t d0 = d_tmem + (uint32_t)s0.hi * kSplitW, d1 = d_tmem + (uint32_t)s1.hi * kSplitW;
              if (elect_one()) {
                umma_i8_pair(d0, da0, db, idesc_half, 1u);
                umma_i8_pair(d0, da0 + 2, db + 2, idesc_half, 1u);
                umma_i8_pair(d0, da0 + 4, db + 4, idesc_half, 1u);
                umma_i8_pair(d0, da0 + 6, db + 6, idesc_half, 1u);
                if (g == 2) {
                  umma_i8_pair(d1, da1, db1, idesc_half, 1u);
                  umma_i8_pair(d1, da1 + 2, db1 + 2, idesc_half, 1u);
                  umma_i8_pair(d1, da1 + 4, db1 + 4, idesc_half, 1u);
                  umma_i8_pair(d1, da1 + 6, db1 + 6, idesc_half, 1u);
                }
                umma_commit_pair(b_empty(sb));
                if (last_chunk) {
                  umma_commit_pair(a_empty(sa0));
                  if (g == 2) umma_commit_pair(a_empty(sa1));
                }
                if (last) umma_commit_pair(tfull_bar(buf));
              }
              __syncwarp();
            }
            if (++sb == (uint32_t)a.nB) { sb = 0; b_par ^= 1; }
          }
        }
      }
    }
  } else if (MODE == DEC1 && warp < kPairEpiWarp0Dec1) {
    // ===================== DEC1 transform (both CTAs): e (uint16, global) -> byte-limb A slots, once per tile =====
    const int t = threadIdx.x;
    const int chunk = t & 7;
    const int r0 = t >> 3;
    const uint16_t *src = reinterpret_cast<const uint16_t *>(a.a_src);
    auto load_atom = [&](int T, int at, bool valid, uint4 (&raw)[8]) {
      const int col = at * kAtomK + chunk * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t row = (size_t)T * 256 + rank * kTileRows + r0 + 32 * j;
        const bool ok = valid && row < a.B && col < a.P;
        const uint4 *ptr = reinterpret_cast<const uint4 *>(src + (ok ? row * (size_t)a.P + col : 0));
        uint4 x0 = __ldg(ptr), x1 = __ldg(ptr + 1);
        if (!ok) x0 = x1 = make_uint4(0, 0, 0, 0);
        raw[2 * j] = x0;
        raw[2 * j + 1] = x1;
      }
    };
    int T = blockIdx.x >> 1, at = 0;
    uint32_t t_par = 0;
    bool valid = T < a.npairs;
    uint4 raw[8], raw_next[8];
    if (valid) load_atom(T, at, true, raw);
    while (valid) {
      int Tn = T, atn = at + 1;
      uint32_t parn = t_par;
      if (atn == atoms) { atn = 0; Tn += gridDim.x >> 1; parn ^= 1; }
      const bool validn = Tn < a.npairs;
      load_atom(Tn, atn, validn, raw_next);                 // prefetch one atom ahead (zeros when !validn)
      const uint32_t sa0 = (uint32_t)(at * kl), sa1 = sa0 + 1;
      mbar_wait(a_empty(sa0), t_par ^ 1);
      if (kl == 2) mbar_wait(a_empty(sa1), t_par ^ 1);
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
        if (kl == 2) {
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
        if (kl == 2) mbar_arrive_cluster(lead(a_full(sa1)));
      }
      T = Tn; at = atn; t_par = parn; valid = validn;
#pragma unroll
      for (int j = 0; j < 8; ++j) raw[j] = raw_next[j];
    }
  } else {
    // ===================== epilogue (both CTAs): two groups, one per TMEM buffer =====================
    // Per pass a warp reads the hi columns of its units first (quotient = -hi goes out at once), then the lo columns,
    // and forms remainder = lo + hi (+ m).  The 2 KB staging tile is used twice per pass (quotient, then remainder).
    constexpr int kGroupWarps = kEpiWarps / 2;
    constexpr int kSub = kGroupWarps / 4;
    constexpr int kPassUnits = 2;
    const int ew = warp - (MODE == DEC1 ? kPairEpiWarp0Dec1 : 0);
    const int quad = warp & 3;
    const uint32_t grp = (uint32_t)(ew >> 2) & 1u;
    const int sub = ew >> 3;
    const int upw = (NCo >> 4) / kSub;
    const int npass = upw / kPassUnits;
    const int nl = a.nl;
    const uint32_t Q2 = a.qmask | (a.qmask << 16);
    const uint32_t LA2 = (((uint32_t)a.q >> 1) - 1u) * 0x00010001u;
    const int logq = 31 - __clz(a.q);
    const uint32_t stage = stage_base + (uint32_t)ew * (MODE == DEC1 ? 4096u : 2048u);
    const uint32_t st_row = stage + (uint32_t)lane * 64u, st_x = (uint32_t)(lane >> 1) & 3u;   // SWIZZLE_64B rows
    const int row_in_tile = quad * 32 + lane;
    // message tile: 128-byte rows (SWIZZLE_128B) when NCo = 128, 64-byte rows (SWIZZLE_64B) when NCo = 64
    const uint32_t m_row = NCo == 128 ? (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128) : (uint32_t)row_in_tile * 64u;
    const uint32_t m_x = NCo == 128 ? ((uint32_t)row_in_tile & 7u) : (((uint32_t)row_in_tile >> 1) & 3u);
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * kAccCols;
    const uint32_t my_tfull = tfull_bar(grp), lead_tempty = lead(tempty_bar(grp));
    uint32_t cc = 0, mc = 0;
    bool store_pending = false;
    auto wait_stage = [&]() {          // the previous store has finished reading the staging tile
      if (store_pending) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
    };
    for (int T = blockIdx.x >> 1; T < a.npairs; T += gridDim.x >> 1) {
      const int out_row = T * 256 + (int)rank * kTileRows + quad * 32;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const uint32_t ms = mc & 1, m_par = (mc >> 1) & 1;
        if (MODE == ENC) ++mc;
        if ((cc & 1u) != grp) continue;
        mbar_wait(my_tfull, (cc >> 1) & 1);
        tc_fence_after();
        if (MODE == ENC) mbar_wait(m_full(ms), m_par);
        for (int ps = 0; ps < npass; ++ps) {
          const int u0 = sub * upw + ps * kPassUnits;
          const bool last_pass = ps == npass - 1;
          const int col = c * NCo + u0 * 16;
          uint4 mm[kPassUnits];
          if (MODE == ENC) {
#pragma unroll
            for (int j = 0; j < kPassUnits; ++j) mm[j] = lds128(m_slot(ms) + m_row + ((((uint32_t)(u0 + j)) ^ m_x) << 4));
          }
          if (MODE == ENC || MODE == DEC1) {
            uint32_t hq[kPassUnits][8], rq[kPassUnits][8];
            {   // ---- hi columns ----
              uint32_t acc[kPassUnits][32], acc1[kPassUnits][32];
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                tmem_ld16(t_addr + kSplitW + (u0 + j) * 16, acc[j]);
                if (MODE == ENC && nl == 2) tmem_ld16(t_addr + kSplitW + NCo + (u0 + j) * 16, acc1[j]);
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  uint32_t x0 = acc[j][2 * jj], x1 = acc[j][2 * jj + 1];
                  if (MODE == ENC && nl == 2) { x0 += acc1[j][2 * jj] << 8; x1 += acc1[j][2 * jj + 1] << 8; }
                  hq[j][jj] = __byte_perm(x0, x1, 0x5410) & Q2;
                }
            }
            {   // ---- lo columns ----
              uint32_t acc[kPassUnits][32], acc1[kPassUnits][32];
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                tmem_ld16(t_addr + (u0 + j) * 16, acc[j]);
                if (MODE == ENC && nl == 2) tmem_ld16(t_addr + NCo + (u0 + j) * 16, acc1[j]);
              }
              tmem_ld_wait();
              if (last_pass) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  mbar_arrive_cluster(lead_tempty);
                  if (MODE == ENC) mbar_arrive(m_empty(ms));
                }
              }
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                const uint32_t mw[4] = {mm[j].x, mm[j].y, mm[j].z, mm[j].w};
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  uint32_t x0 = acc[j][2 * jj], x1 = acc[j][2 * jj + 1];
                  if (MODE == ENC && nl == 2) { x0 += acc1[j][2 * jj] << 8; x1 += acc1[j][2 * jj + 1] << 8; }
                  uint32_t s = (__byte_perm(x0, x1, 0x5410) & Q2) + hq[j][jj];
                  if (MODE == ENC) s += __byte_perm(mw[jj >> 1], 0u, (jj & 1) ? 0x4342 : 0x4140);
                  rq[j][jj] = s & Q2;
                }
              }
            }
            // quotient = -hi mod q
            if (a.out_mask & 4) {
              wait_stage();
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                uint32_t v[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) v[jj] = ((~hq[j][jj] & Q2) + 0x00010001u) & Q2;
                sts128(st_row + ((((uint32_t)(2 * j)) ^ st_x) << 4), make_uint4(v[0], v[1], v[2], v[3]));
                sts128(st_row + ((((uint32_t)(2 * j + 1)) ^ st_x) << 4), make_uint4(v[4], v[5], v[6], v[7]));
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmapO2, stage, col, out_row);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              store_pending = true;
            }
            // remainder (and, DEC1, the multiplier of the second product)
            wait_stage();
#pragma unroll
            for (int j = 0; j < kPassUnits; ++j) {
              sts128(st_row + ((((uint32_t)(2 * j)) ^ st_x) << 4), make_uint4(rq[j][0], rq[j][1], rq[j][2], rq[j][3]));
              sts128(st_row + ((((uint32_t)(2 * j + 1)) ^ st_x) << 4), make_uint4(rq[j][4], rq[j][5], rq[j][6], rq[j][7]));
            }
            if (MODE == DEC1 && a.o8_cyc) {
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) {
                uint32_t bw[4];
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                  uint32_t y2[2];
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    const uint32_t x = rq[j][2 * wd + i];
                    const uint32_t y = x + (((x + LA2) >> logq) & 0x00010001u);        // index.js:117
                    y2[i] = (y & 0x003F003Fu) + ((y >> 6) & 0x00FF00FFu);              // = y (mod 3), one byte
                  }
                  bw[wd] = __byte_perm(y2[0], y2[1], 0x6420);
                }
                sts128(stage + 2048u + (uint32_t)lane * 32u + (uint32_t)j * 16u, make_uint4(bw[0], bw[1], bw[2], bw[3]));
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (a.out_mask & 1) tma_store_2d(&tmapO0, stage, col, out_row);
              if (MODE == DEC1) {
                if (a.out_mask & 2) tma_store_2d(&tmapO1, stage + 2048, col, out_row);
              } else {
                if (a.out_mask & 2) tma_store_2d(&tmapO1, stage, col, out_row);
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
          } else {
            // ---- DEC2: remainder2 = (lo + hi) mod 3, quotient2 = -hi = 2 hi (mod 3); 32-byte rows ----
            uint32_t hi4[kPassUnits][16];
            uint32_t rem[kPassUnits][4], quo[kPassUnits][4];
            {
              uint32_t acc[kPassUnits][32];
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) tmem_ld16(t_addr + kSplitW + (u0 + j) * 16, acc[j]);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) hi4[j][i] = acc[j][i];
            }
            {
              uint32_t acc[kPassUnits][32];
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j) tmem_ld16(t_addr + (u0 + j) * 16, acc[j]);
              tmem_ld_wait();
              if (last_pass) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(lead_tempty);
              }
#pragma unroll
              for (int j = 0; j < kPassUnits; ++j)
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                  uint32_t rb[4], qb[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint32_t s = acc[j][4 * wd + i] + hi4[j][4 * wd + i], d2 = 2u * hi4[j][4 * wd + i];
                    rb[i] = s - 3u * __umulhi(s, 0x55555556u);
                    qb[i] = d2 - 3u * __umulhi(d2, 0x55555556u);
                  }
                  rem[j][wd] = __byte_perm(__byte_perm(rb[0], rb[1], 0x0040), __byte_perm(rb[2], rb[3], 0x0040), 0x5410);
                  quo[j][wd] = __byte_perm(__byte_perm(qb[0], qb[1], 0x0040), __byte_perm(qb[2], qb[3], 0x0040), 0x5410);
                }
            }
            wait_stage();
#pragma unroll
            for (int j = 0; j < kPassUnits; ++j) {
              sts128(stage + (uint32_t)lane * 32u + (uint32_t)j * 16u, make_uint4(rem[j][0], rem[j][1], rem[j][2], rem[j][3]));
              sts128(stage + 1024u + (uint32_t)lane * 32u + (uint32_t)j * 16u, make_uint4(quo[j][0], quo[j][1], quo[j][2], quo[j][3]));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (a.out_mask & 1) tma_store_2d(&tmapO0, stage, col, out_row);
              if (a.out_mask & 2) tma_store_2d(&tmapO1, stage, col, out_row);
              if (a.out_mask & 4) tma_store_2d(&tmapO2, stage + 1024, col, out_row);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kPairMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}
