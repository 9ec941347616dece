// hb_serial.cuh -- the scalar side of the sweep as ONE serial CTA + helper CTAs (mixture models B / C / R).
//
// The ring of workers (scalar_role, hb_sweep.cuh) hands the corrections for the next tile from worker to worker
// through L2; that hand-over (a store, a trip through a loaded L2, the spread of 256 polls) and the correction for
// the tile after next were what paced the round-1 kernel: 4.9 us per tile against 2.5-3 us of work.  Here the serial
// dependence of the Gibbs chain (Bayes.cpp:751-802: SNP j+1 sees the residual after SNP j) never leaves one SM:
//
//   helper CTA (tile t -> helper t mod NH)
//     P  as soon as the tile's dots are complete: speculated classes, candidate list, the candidates' Gram rows
//        (diagonal block and block towards t+1), the solved chain matrix -- everything phase S needs that does not
//        depend on the tiles still in flight -- written as one contiguous *package* to global memory, then a flag.
//     C  after the serial CTA has published the tile's changes: the corrections the tile owes to the tiles
//        t+3 .. t+D-1 (Gram band rows from L2), posted to those tiles' correction slots.
//   serial CTA
//     loader warp: per tile one cp.async.bulk of the package into a double buffer (mbarrier full/empty), two tiles
//        ahead of the chain.
//     phase S of every tile, back to back: exact right-hand sides from the package's dots, the far corrections
//        (slots written by the helpers, read one tile ahead), the correction from t-2 (Gram entries fetched into
//        registers one tile ahead, summed here) and the correction from t-1, which never leaves the thread that
//        computed it; candidate chain (one matrix-vector product); every SNP's exact right-hand side; each SNP's class
//        is checked against the interval of rhs^2 on which the speculated class holds (two compares).  A class that
//        differs starts the round-based repair of scalar_role (re-compaction from the rows in shared memory, missing
//        rows fetched from the Gram band).  Publishes the changes for the AXPY warps and the helpers.
//
// Arithmetic (order of every sum) is the same as in scalar_role, so both modes give bit-identical effects.
#pragma once

namespace hbk {

// ---- package layout (bytes; B = SNPs per tile, KC = row slots = largest candidate list shipped)
struct PkgLayout {
  uint32_t o_hdr, o_base0, o_addback, o_lo, o_hi, o_info, o_slotof;
  uint32_t o_rhs0, o_iv, o_sdz, o_gold, o_delta, o_gnew, o_idx, o_cls, o_slot, o_M, o_rows;
  uint32_t fixed_bytes, stride, row_bytes;
};
__host__ __device__ constexpr PkgLayout pkg_layout(int B, int KC) {
  PkgLayout L{};
  uint32_t o = 0;
  L.o_hdr = o; o += 64;
  L.o_base0 = o; o += 8 * B;
  L.o_addback = o; o += 8 * B;
  L.o_lo = o; o += 8 * B;
  L.o_hi = o; o += 8 * B;
  L.o_info = o; o += 4 * B;
  L.o_slotof = o; o += 4 * B;
  const uint32_t kc8 = 8 * ((KC + 1) & ~1u), kc4 = 4 * ((KC + 3) & ~3u);
  L.o_rhs0 = o; o += kc8;
  L.o_iv = o; o += kc8;
  L.o_sdz = o; o += kc8;
  L.o_gold = o; o += kc8;
  L.o_delta = o; o += kc8;
  L.o_gnew = o; o += kc8;
  L.o_idx = o; o += kc4;
  L.o_cls = o; o += kc4;
  L.o_slot = o; o += kc4;
  L.o_M = o; o += 32 * 33 * 8;
  o = (o + 127) & ~127u;
  L.o_rows = o;
  L.fixed_bytes = o;
  L.row_bytes = 2u * B * 4u;   // a row slot: the SNP's Gram row in the diagonal block, then in the block towards t+1
  L.stride = (o + (uint32_t)KC * L.row_bytes + 127) & ~127u;
  return L;
}
// header words
enum { PK_K = 0, PK_NS = 1, PK_FAST = 2, PK_MOK = 3, PK_TILE = 4, PK_HAS1 = 5 };
// info word of a SNP: class | candidate << 4 | active << 5 | exact-check << 6 | rank << 8
constexpr int kPkgNoRows = 1 << 20;
constexpr int kMaxDC = 11;   // serial mode: lag of at most 12 tiles (far-correction words a thread asks for one tile ahead)
constexpr int kSerialB = 256;   // serial mode is compiled for tiles of 256 SNPs (every offset below is a constant)

// private shared memory of the serial CTA behind the two package buffers (byte offsets from its start)
struct PrivLayout {
  uint32_t o_part_rhs, o_part_corr, o_c2buf, o_coef, o_dl2, o_ix2, o_new_snp, o_rank, o_g2land, o_wcnt, o_gctl, o_pc, o_full, o_empty, bytes;
};
__host__ __device__ constexpr PrivLayout priv_layout(int B) {
  PrivLayout P{};
  uint32_t o = 0;
  P.o_part_rhs = o; o += 8 * B;
  P.o_part_corr = o; o += 8 * B;
  P.o_c2buf = o; o += 2 * 8 * B;      // [2][B] correction owed by tile t-2, by parity of the receiving tile
  P.o_coef = o; o += 32 * 32 * 8;     // chain coefficients of the first 32 candidates (no solved matrix / repair rounds)
  P.o_dl2 = o; o += 2 * 32 * 8;       // first 32 changes of the last two tiles ...
  P.o_ix2 = o; o += 2 * 32 * 4;       // ... and their SNPs
  P.o_new_snp = o; o += 4 * B;
  P.o_rank = o; o += 4 * B;           // candidates before SNP i (repair rounds: the secondary threads read it)
  P.o_g2land = o; o += 32 * 32 * 4;   // landing zone of the last primary warp's Gram entries towards t+2
  P.o_wcnt = o; o += 64 * 4;
  P.o_gctl = o; o += 16 * 4;          // [1] k, [2] row slots, [8+b] k of the tile of parity b
  P.o_pc = o; o += 18 * 8;
  P.o_full = o; o += 2 * 8;
  P.o_empty = o; o += 2 * 8;
  P.bytes = (o + 127) & ~127u;
  return P;
}
__host__ __device__ constexpr int serial_krow(int B) {
  const uint32_t cap = 226 * 1024 - 2048;
  int kc = B;
  for (; kc > 12; --kc)
    if (2 * pkg_layout(B, kc).stride + priv_layout(B).bytes <= cap) break;
  return kc;
}
__host__ inline size_t serial_smem_bytes(int B) {
  return 2 * (size_t)pkg_layout(B, serial_krow(B)).stride + priv_layout(B).bytes;
}

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// serial CTA
// ------------------------------------------------------------------------------------------
// the serial CTA's copy of the kernel parameters (shared memory: read all along the chain)
__shared__ SweepParams g_serial_ps;

// What a thread knows about its SNP while a tile is being decided (primary threads; the secondary threads use k, ns
// and the flags only).
struct SerialTile {
  double rhs0, lo, hi, gold, myiv, mysdz, c1next, my_delta, my_gnew;
  int cls, cls2, myrank, slot, k, ns, nrounds;
  bool act, cand, chk_exact, generic, m_ok, dead, has1;
};

// compile-time geometry of the serial CTA
template <int B>
struct SerialGeo {
  static constexpr int KC = serial_krow(B);
  static constexpr PkgLayout L = pkg_layout(B, KC);
  static constexpr PrivLayout P = priv_layout(B);
  static constexpr int RS = 2 * B;          // ints between row slots
  static constexpr int NTS = 2 * B - 32;    // threads of the chain (the last warp of the block is the loader)
  static constexpr uint32_t priv0 = 2 * L.stride;
};

// The rare paths of a tile, kept out of the chain's instruction stream and registers: a package without rows (more
// candidates than a package holds: the classes are decided here, chain and sums straight from the Gram band), and the
// repair rounds after a class differed from its speculation (re-compaction from the rows in shared memory, missing rows
// from the Gram band), round after round as in scalar_role, until every class agrees with its exact right-hand side.
template <int NF, bool PRIM, int B>
__device__ __noinline__ void serial_slow(SerialTile& st, int t, bool from_start) {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PkgLayout L = G::L;
  constexpr PrivLayout P = G::P;
  constexpr int KC = G::KC, RS = G::RS, NTS = G::NTS;
  constexpr bool prim = PRIM;
  constexpr int h = PRIM ? 0 : 1;
  const SweepParams& p = g_serial_ps;
  const int tid = threadIdx.x, i = tid - h * B, warp = i >> 5, lane = i & 31;
  constexpr int nwarp = B / 32;
  const bool tailw = prim && warp == nwarp - 1;
  const int D = p.D, model = p.model;
  const int nf = (model == HB_MODEL_R) ? p.F : 2;
  const size_t mp = p.m_pad;
  const int j = t * B + i;
  uint8_t* pk = smem + (size_t)(t & 1) * L.stride;
  uint8_t* priv = smem + G::priv0;
  double* part_rhs = (double*)(priv + P.o_part_rhs);
  double* part_corr = (double*)(priv + P.o_part_corr);
  double* coef = (double*)(priv + P.o_coef);
  int* new_snp = (int*)(priv + P.o_new_snp);
  int* rank_sh = (int*)(priv + P.o_rank);
  int* wcnt = (int*)(priv + P.o_wcnt);
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  int32_t* rows = (int32_t*)(pk + L.o_rows);
  const int32_t* G0 = p.gram + ((size_t)t * D) * B * B;
  const bool has1 = st.has1;
  CandSet cs;
  cs.rhs0 = (double*)(pk + L.o_rhs0); cs.iv = (double*)(pk + L.o_iv); cs.sdz = (double*)(pk + L.o_sdz);
  cs.gold = (double*)(pk + L.o_gold); cs.delta = (double*)(pk + L.o_delta); cs.gnew = (double*)(pk + L.o_gnew);
  cs.idx = (int*)(pk + L.o_idx); cs.cls = (int*)(pk + L.o_cls); cs.slot = (int*)(pk + L.o_slot);

  double TL[NF - 1], TH[NF - 1];
#pragma unroll
  for (int q = 0; q < NF - 1; ++q) { TL[q] = -1.0; TH[q] = -1.0; }
  if (prim && st.act) {
#pragma unroll
    for (int q = 0; q < NF - 1; ++q)
      if (q < nf - 1) {
        TL[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q, mp, j));
        TH[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q + 1, mp, j));
      }
  }
  auto classify_full = [&](double rhs) -> int {
    const double rr = rhs * rhs;
    const int c0 = thr_class<NF>(nf, rr, TL, TH);
    if (c0 >= 0) return c0;
    return classify_exact<NF>(p.prm, mp, j, nf, rr, p.logpi0);
  };
  auto load_draw = [&]() {   // 1/v and sd*z of the SNP's current class
    st.myiv = 0.0; st.mysdz = 0.0;
    if (st.cls > 0) {
      st.myiv = __ldcg(p.prm + prm_idx(4 + 4 * (st.cls - 1), mp, j));
      st.mysdz = __ldcg(p.prm + prm_idx(5 + 4 * (st.cls - 1), mp, j));
    }
  };
  int k = st.k, ns = st.ns;
  bool generic = st.generic;
  // candidate list of the current classes, from the rows in shared memory
  auto compact_rows = [&]() {
    st.cand = st.act && (st.cls > 0 || st.gold != 0.0);
    const bool need_row = st.cand && st.slot < 0;
    if (prim) {
      const unsigned bal = __ballot_sync(0xffffffffu, st.cand), bal2 = __ballot_sync(0xffffffffu, need_row);
      if (lane == 0) { wcnt[warp] = __popc(bal); wcnt[32 + warp] = __popc(bal2); }
      hb::named_bar_sync(4, B);   // primary threads only
      int pre = 0, pre2 = 0, tot2 = 0;
      k = 0;
      for (int w = 0; w < nwarp; ++w) {
        const int c = wcnt[w], c2 = wcnt[32 + w];
        if (w < warp) { pre += c; pre2 += c2; }
        k += c; tot2 += c2;
      }
      st.myrank = pre + __popc(bal & ((1u << lane) - 1u));
      rank_sh[i] = st.myrank;
      if (need_row) {
        const int nsl = pre2 + __popc(bal2 & ((1u << lane) - 1u));
        st.slot = ns + nsl;
        if (ns + nsl < B) new_snp[nsl] = i;
      }
      if (i == 0) { gctl[1] = k; gctl[2] = ns + tot2; }
    }
    hb::named_bar_sync(1, NTS);
    k = gctl[1];
    if (!prim) st.myrank = rank_sh[i];
    const int ns_new = gctl[2];
    if (k > KC || ns_new > KC) { generic = true; return; }
    if (prim && st.cand) {
      cs.idx[st.myrank] = i; cs.gold[st.myrank] = st.gold; cs.cls[st.myrank] = st.cls; cs.slot[st.myrank] = st.slot;
      cs.iv[st.myrank] = st.myiv; cs.sdz[st.myrank] = st.mysdz;
    }
    if (ns_new > ns) {
      // rows of the new candidates: both blocks straight from the Gram band (one trip to L2)
      const int nnew = ns_new - ns;
      for (int e = tid; e < nnew * RS; e += NTS) {
        const int sl = e / RS, c = e - sl * RS;
        const int blk = c / B, col = c - blk * B;
        if (blk == 0 || has1)
          cp_async4((uint32_t)__cvta_generic_to_shared(rows + (size_t)(ns + sl) * RS + c),
                    G0 + (size_t)blk * B * B + (size_t)new_snp[sl] * B + col);
      }
      cp_async_wait_all();
      ns = ns_new;
    }
    hb::named_bar_sync(1, NTS);
    const int kk = min(k, 32);
    for (int e = tid; e < 32 * 32; e += NTS) {
      const int lp = e >> 5, sc = e & 31;
      double v = 0.0;
      if (lp < sc && sc < kk) v = gram_as_double(rows[(size_t)cs.slot[lp] * RS + cs.idx[sc]]) * (-cs.iv[sc]);
      coef[e] = v;
    }
  };
  // the same without rows: candidate arrays of B entries laid over the (then unused) row area of the buffer
  auto compact_generic = [&]() {
    cs = make_candset((uint8_t*)rows, B);
    st.cand = st.act && (st.cls > 0 || st.gold != 0.0);
    if (prim) {
      const unsigned bal = __ballot_sync(0xffffffffu, st.cand);
      if (lane == 0) wcnt[warp] = __popc(bal);
      hb::named_bar_sync(4, B);
      int pre = 0;
      k = 0;
      for (int w = 0; w < nwarp; ++w) {
        const int c = wcnt[w];
        if (w < warp) pre += c;
        k += c;
      }
      st.myrank = pre + __popc(bal & ((1u << lane) - 1u));
      rank_sh[i] = st.myrank;
      if (i == 0) gctl[1] = k;
      if (st.cand) {
        cs.idx[st.myrank] = i; cs.gold[st.myrank] = st.gold; cs.cls[st.myrank] = st.cls; cs.slot[st.myrank] = -1;
        cs.iv[st.myrank] = st.myiv; cs.sdz[st.myrank] = st.mysdz;
      }
    }
    hb::named_bar_sync(1, NTS);
    k = gctl[1];
    if (!prim) st.myrank = rank_sh[i];
  };

  if (from_start) {
    // no rows in the package: decide the classes from the complete right-hand side
    if (prim) {
      st.gold = __ldcg(p.g + j);
      if (st.act) { st.cls = classify_full(st.rhs0); load_draw(); }
    }
    compact_generic();
  } else {
    // a class differed from its speculation in the first round
    if (prim && st.cls2 != st.cls) { st.cls = st.cls2; load_draw(); }
    if (!generic) compact_rows();
    if (generic) { hb::named_bar_sync(1, NTS); compact_generic(); }
  }
  for (;;) {
    ++st.nrounds;
    if (st.cand && prim) cs.rhs0[st.myrank] = st.rhs0;
    hb::named_bar_sync(1, NTS);
    double prhs = 0.0, pcorr = 0.0, prhs2 = 0.0, pcorr2 = 0.0;
    if (!generic) {
      if (tid < 32 && k > 0) chain_candidates<true>(cs, k, G0, rows, B, lane, coef, RS);
      hb::named_bar_sync(1, NTS);
#pragma unroll 4
      for (int sidx = h; sidx < k; sidx += 2) {
        const double d = cs.delta[sidx];
        const int sl = cs.slot[sidx];
        const double g0 = gram_as_double(rows[(size_t)sl * RS + i]), g1 = gram_as_double(rows[(size_t)sl * RS + B + i]);
        prhs = fma(sidx < st.myrank ? g0 : 0.0, d, prhs);
        pcorr = fma(has1 ? g1 : 0.0, d, pcorr);
      }
      if (tailw) {
#pragma unroll 4
        for (int sidx = 1; sidx < k; sidx += 2) {
          const double d = cs.delta[sidx];
          const int sl = cs.slot[sidx];
          const double g0 = gram_as_double(rows[(size_t)sl * RS + i]), g1 = gram_as_double(rows[(size_t)sl * RS + B + i]);
          prhs2 = fma(sidx < st.myrank ? g0 : 0.0, d, prhs2);
          pcorr2 = fma(has1 ? g1 : 0.0, d, pcorr2);
        }
      }
    } else {
      const double sv = slow_chain_and_sums_cs(cs, k, st.myrank, G0, B, i, h, has1, false, model, NTS);
      if (prim) prhs = sv; else pcorr = sv;
      if (tailw) pcorr2 = has1 ? band_correction(cs, k, G0 + (size_t)B * B, B, i) : 0.0;
    }
    if (!prim) { part_rhs[i] = prhs; part_corr[i] = pcorr; }
    hb::named_bar_sync(1, NTS);
    int cls2 = st.cls;
    if (prim) {
      double rhs;
      if (tailw) { rhs = st.rhs0 - (prhs + prhs2); st.c1next = pcorr + pcorr2; }
      else { rhs = st.rhs0 - (prhs + part_rhs[i]); st.c1next = pcorr + part_corr[i]; }
      if (st.act) cls2 = classify_full(rhs);
    }
    const bool bad = prim && st.act && (cls2 != st.cls);
    const bool redo = hb::named_bar_or(1, NTS, bad);
    if (!redo) break;
    if (prim && cls2 != st.cls) { st.cls = cls2; load_draw(); }
    if (!generic) compact_rows();
    if (generic) { hb::named_bar_sync(1, NTS); compact_generic(); }
  }
  // the tile is final: this thread's change, and the first 32 changes for the correction towards t+2
  if (prim && st.cand) { st.my_delta = cs.delta[st.myrank]; st.my_gnew = cs.gnew[st.myrank]; }
  if (tid < 32) {
    ((double*)(priv + P.o_dl2))[(t & 1) * 32 + tid] = (tid < k) ? cs.delta[tid] : 0.0;
    ((int*)(priv + P.o_ix2))[(t & 1) * 32 + tid] = (tid < k) ? cs.idx[tid] : 0;
  }
  st.k = k; st.ns = ns; st.generic = generic;
}

// The primary threads (one per SNP: inputs, classes, verification) and the secondary threads (odd candidates of the
// sums, Gram entries towards t+2 in registers) run the same sequence of barriers from two instantiations of this
// function, so that neither carries the other's registers.
template <int NF, bool PRIM, int B>
__device__ __noinline__ void serial_threads() {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PkgLayout L = G::L;
  constexpr PrivLayout P = G::P;
  constexpr int RS = G::RS, NTS = G::NTS;
  constexpr bool prim = PRIM;
  constexpr int h = PRIM ? 0 : 1;
  constexpr int nwarp = B / 32;
  const SweepParams& p = g_serial_ps;
  const int D = p.D, T = p.T;
  const int tid = threadIdx.x;
  const int i = tid - h * B, warp = i >> 5, lane = i & 31;
  const bool tailw = prim && warp == nwarp - 1;   // primary threads without a secondary partner
  int* ctrl = p.ctrl;
  uint8_t* priv = smem + G::priv0;
  double* part_rhs = (double*)(priv + P.o_part_rhs);
  double* part_corr = (double*)(priv + P.o_part_corr);
  double* c2buf = (double*)(priv + P.o_c2buf);
  double* coef = (double*)(priv + P.o_coef);
  double* dl2 = (double*)(priv + P.o_dl2);
  int* ix2 = (int*)(priv + P.o_ix2);
  int32_t* g2land = (int32_t*)(priv + P.o_g2land);
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  long long* pc = (long long*)(priv + P.o_pc);
  uint64_t* full = (uint64_t*)(priv + P.o_full);
  uint64_t* empty = (uint64_t*)(priv + P.o_empty);

  const int DC = D - 1;
  bool dead = false;
  int rounds_total = 0, changed_total = 0, repaired = 0, generic_tiles = 0;
  double corr1 = 0.0;               // correction owed by the previous tile to this thread's SNP (never leaves the thread)
  unsigned long long cwn[PRIM ? kMaxDC : 1];   // far corrections of the next tile, asked for one tile ahead
  int g2[PRIM ? 1 : 32];            // Gram entries towards the tile after next (secondary threads), asked for one tile ahead
#pragma unroll
  for (int q = 0; q < (PRIM ? kMaxDC : 1); ++q) cwn[q] = kCorrEmpty;
#pragma unroll
  for (int e = 0; e < (PRIM ? 1 : 32); ++e) g2[e] = 0;

  const bool timing = PRIM && tid == 0 && (p.dbg & 64);
#define HB_SPHASE(n) do { if (timing) { const long long _now = clock64(); pc[n] += _now - pc[16]; pc[16] = _now; } } while (0)
  const int abl = p.dbg;   // timing experiments (results are wrong with them): 256 no t+2 correction, 512 no far corrections, 1024 no effect/class stores
  for (int t = 0; t < T; ++t) {
    const int b = t & 1;
    const int j = t * B + i;
    uint8_t* pk = smem + (size_t)b * L.stride;
    if (!mbar_wait(full + b, (uint32_t)((t >> 1) & 1), ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
    HB_SPHASE(0);
    if (tid == 0) HB_TRACE(t, 2);
    const int* hdr = (const int*)(pk + L.o_hdr);
    // (plain locals: everything the chain touches stays in registers; a SerialTile is filled only for the rare paths)
    int k = dead ? 0 : hdr[PK_K];
    const int ns0 = dead ? 0 : hdr[PK_NS];
    const bool fast = !dead && hdr[PK_FAST] != 0;
    const bool m_ok = !dead && hdr[PK_MOK] != 0;
    const bool has1 = (D > 1 && t + 1 < T);
    double rhs0 = 0.0, lo = -1.0, hi = HUGE_VAL, gold = 0.0, myiv = 0.0, mysdz = 0.0, c1next = 0.0, my_delta = 0.0, my_gnew = 0.0;
    int cls = 0, cls2 = 0, myrank = 0, slot = -1, nrounds = 0;
    bool act = false, cand = false, chk_exact = false;
    double* c_rhs0 = (double*)(pk + L.o_rhs0);
    double* c_delta = (double*)(pk + L.o_delta);
    const int* c_slot = (const int*)(pk + L.o_slot);
    const int32_t* rows = (const int32_t*)(pk + L.o_rows);

    // ---- this SNP's inputs
    if constexpr (PRIM) if (!dead) {
      const double base0 = ((const double*)(pk + L.o_base0))[i];
      const double addback = ((const double*)(pk + L.o_addback))[i];
      lo = ((const double*)(pk + L.o_lo))[i];
      hi = ((const double*)(pk + L.o_hi))[i];
      const int info = ((const int*)(pk + L.o_info))[i];
      slot = ((const int*)(pk + L.o_slotof))[i];
      cls = info & 15; cand = (info >> 4) & 1; act = (info >> 5) & 1; chk_exact = (info >> 6) & 1;
      myrank = info >> 8;
      if (cand && fast) {
        gold = ((const double*)(pk + L.o_gold))[myrank];
        myiv = ((const double*)(pk + L.o_iv))[myrank];
        mysdz = ((const double*)(pk + L.o_sdz))[myrank];
      }
      const int dmax = min(DC, t);
      double cold = 0.0;
#pragma unroll
      for (int q = kMaxDC - 1; q >= 2; --q)
        if (q + 1 <= dmax && !(abl & 512)) {
          double v = __longlong_as_double((long long)cwn[q]);
          if (cwn[q] == kCorrEmpty) {
            if (!poll_corr_slow(p.corr + ((size_t)t * DC + q) * B + i, v, ctrl)) { dead = true; v = 0.0; }
          }
          cold += v;
        }
      if (dmax >= 2 && !(abl & 256)) cold += c2buf[b * B + i];
      const double c1 = (dmax >= 1) ? corr1 : 0.0;
      rhs0 = ((base0 - cold) - c1) + addback;
    }
    if constexpr (!PRIM) if (!dead) myrank = ((const int*)(pk + L.o_info))[i] >> 8;   // candidates before SNP i
    HB_SPHASE(1);

    if (fast) {
      if (!m_ok && k > 0) {
        // more than 32 candidates: no solved chain matrix in the package; the coefficients of the first 32 for the
        // step-by-step chain
        const int* c_idx = (const int*)(pk + L.o_idx);
        const double* c_iv = (const double*)(pk + L.o_iv);
        for (int e = tid; e < 32 * 32; e += NTS) {
          const int lp = e >> 5, sc = e & 31;
          double v = 0.0;
          if (lp < sc && sc < min(k, 32)) v = gram_as_double(rows[(size_t)c_slot[lp] * RS + c_idx[sc]]) * (-c_iv[sc]);
          coef[e] = v;
        }
      }
      if (prim && cand && !dead) c_rhs0[myrank] = rhs0;
    }
    if (hb::named_bar_or(1, NTS, dead)) { dead = true; break; }
    bool slow = !fast;
    if (fast) {
      // ---- first round: the speculated candidate list
      nrounds = 1;
      HB_SPHASE(3);
      if (tid < 32 && k > 0) {
        CandSet cs;
        cs.rhs0 = c_rhs0; cs.iv = (double*)(pk + L.o_iv); cs.sdz = (double*)(pk + L.o_sdz);
        cs.gold = (double*)(pk + L.o_gold); cs.delta = c_delta; cs.gnew = (double*)(pk + L.o_gnew);
        cs.idx = (int*)(pk + L.o_idx); cs.cls = (int*)(pk + L.o_cls); cs.slot = (int*)(pk + L.o_slot);
        if (m_ok) chain_matvec(cs, k, (const double*)(pk + L.o_M), lane);
        else chain_candidates<true>(cs, k, nullptr, rows, B, lane, coef, RS);
      }
      HB_SPHASE(4);
      hb::named_bar_sync(1, NTS);
      HB_SPHASE(8);
      // exact right-hand side of every SNP, and the corrections owed to the next tile: the primary thread takes the
      // even candidates, the secondary the odd ones (the last primary warp has no partner and takes both)
      double prhs = 0.0, pcorr = 0.0, prhs2 = 0.0, pcorr2 = 0.0;
#pragma unroll 4
      for (int sidx = h; sidx < k; sidx += 2) {
        const double d = c_delta[sidx];
        const int sl = c_slot[sidx];
        const double g0 = gram_as_double(rows[sl * RS + i]), g1 = gram_as_double(rows[sl * RS + B + i]);
        prhs = fma(sidx < myrank ? g0 : 0.0, d, prhs);
        pcorr = fma(has1 ? g1 : 0.0, d, pcorr);
      }
      if (tailw) {
#pragma unroll 4
        for (int sidx = 1; sidx < k; sidx += 2) {
          const double d = c_delta[sidx];
          const int sl = c_slot[sidx];
          const double g0 = gram_as_double(rows[sl * RS + i]), g1 = gram_as_double(rows[sl * RS + B + i]);
          prhs2 = fma(sidx < myrank ? g0 : 0.0, d, prhs2);
          pcorr2 = fma(has1 ? g1 : 0.0, d, pcorr2);
        }
      }
      if (!prim) { part_rhs[i] = prhs; part_corr[i] = pcorr; }
      HB_SPHASE(9);
      hb::named_bar_sync(1, NTS);
      HB_SPHASE(10);
      cls2 = cls;
      if constexpr (PRIM) {
        double rhs;
        if (tailw) { rhs = rhs0 - (prhs + prhs2); c1next = pcorr + pcorr2; }
        else { rhs = rhs0 - (prhs + part_rhs[i]); c1next = pcorr + part_corr[i]; }
        if (act) {
          const double rr = rhs * rhs;
          if (chk_exact || !(rr >= lo && rr <= hi)) {
            // outside the interval of the speculated class (or no certified interval): the class itself
            const int nf = (p.model == HB_MODEL_R) ? p.F : 2;
            double TL[NF - 1], TH[NF - 1];
#pragma unroll
            for (int q = 0; q < NF - 1; ++q) {
              TL[q] = -1.0; TH[q] = -1.0;
              if (q < nf - 1) {
                TL[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q, p.m_pad, j));
                TH[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q + 1, p.m_pad, j));
              }
            }
            int c0 = thr_class<NF>(nf, rr, TL, TH);
            if (c0 < 0) c0 = classify_exact<NF>(p.prm, p.m_pad, j, nf, rr, p.logpi0);
            cls2 = c0;
          }
        }
      }
      const bool bad = prim && act && (cls2 != cls);
      HB_SPHASE(11);
      slow = hb::named_bar_or(1, NTS, bad);
      HB_SPHASE(12);
      if (!slow) {
        if (prim && cand) { my_delta = c_delta[myrank]; my_gnew = ((const double*)(pk + L.o_gnew))[myrank]; }
        if (tid < 32) {
          dl2[b * 32 + tid] = (tid < k) ? c_delta[tid] : 0.0;
          ix2[b * 32 + tid] = (tid < k) ? ((const int*)(pk + L.o_idx))[tid] : 0;
        }
      }
    }
    if (slow) {
      // the rare paths (no rows in the package / a class differed): out of line, with their own registers
      if (!fast) ++generic_tiles;
      SerialTile st;
      st.rhs0 = rhs0; st.lo = lo; st.hi = hi; st.gold = gold; st.myiv = myiv; st.mysdz = mysdz; st.c1next = c1next;
      st.my_delta = 0.0; st.my_gnew = 0.0;
      st.cls = cls; st.cls2 = cls2; st.myrank = myrank; st.slot = slot; st.k = k; st.ns = ns0; st.nrounds = nrounds;
      st.act = act; st.cand = cand; st.chk_exact = chk_exact; st.generic = !fast; st.m_ok = m_ok; st.dead = false; st.has1 = has1;
      serial_slow<NF, PRIM, B>(st, t, !fast);
      c1next = st.c1next; my_delta = st.my_delta; my_gnew = st.my_gnew;
      cls = st.cls; myrank = st.myrank; k = st.k; nrounds = st.nrounds; cand = st.cand;
    }
    corr1 = c1next;
    rounds_total += nrounds;
    changed_total += k;
    if (nrounds > 1) { ++repaired; if (tid == 0) st_relaxed_s32(p.miss_tile, t); }
    HB_SPHASE(5);
    if (tid == 0) HB_TRACE(t, 3);
    // ---- the tile is final: effects, classes, and the tile's changes for the AXPY warps and the helpers
    if constexpr (PRIM) {
      if (cand) {
        st_relaxed_u64(p.q_delta + (size_t)t * B + myrank, (unsigned long long)__double_as_longlong(my_delta));
        st_relaxed_s32(p.q_snp + (size_t)t * B + myrank, j);
        if (!(abl & 1024)) p.g[j] = my_gnew;
      }
      if (i == 0) { st_relaxed_s32(p.tile_cnt + t, k); HB_TRACE(t, 4); gctl[8 + b] = k; }
      if (act && !(abl & 1024)) p.tracker[j] = cls;
    }
    HB_SPHASE(6);
    // ---- correction owed by tile t-1 to tile t+1: the Gram entries were asked for one tile ago
    if ((!prim || tailw) && t >= 1 && t + 1 < T && D > 2 && !(abl & 256)) {
      const int kp = gctl[8 + (b ^ 1)];
      const double* dprev = dl2 + (b ^ 1) * 32;
      double c = 0.0;
      if constexpr (PRIM) {
        // (last primary warp: its entries landed in shared memory)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int e = 0; e < min(kp, 32); ++e) c = fma(gram_as_double(g2land[e * 32 + lane]), dprev[e], c);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < kp) c = fma(gram_as_double(g2[e]), dprev[e], c);
      }
      if (kp > 32) {
        // rare: the rest of the list from the published queue of tile t-1 (same order as band_correction_raw)
        const int32_t* gb = p.gram + ((size_t)(t - 1) * D + 2) * B * B;
        for (int e = 32; e < kp; ++e) {
          const int sn = hb::ld_relaxed(p.q_snp + (size_t)(t - 1) * B + e) - (t - 1) * B;
          const double d = __longlong_as_double((long long)ld_relaxed_u64(p.q_delta + (size_t)(t - 1) * B + e));
          c = fma(gram_as_double(__ldcg(gb + (size_t)sn * B + i)), d, c);
        }
      }
      c2buf[(b ^ 1) * B + i] = c;
    }
    hb::named_bar_sync(1, NTS);   // dl2 / ix2 / gctl[8+b] of this tile are visible; c2buf for t+1 is complete
    if ((!prim || tailw) && t + 2 < T && D > 2 && !(abl & 256)) {
      const int32_t* gb = p.gram + ((size_t)t * D + 2) * B * B;
      const int* ixp = ix2 + b * 32;
      const int kk = min(k, 32);
      if constexpr (PRIM) {
        for (int e = 0; e < kk; ++e)
          cp_async4((uint32_t)__cvta_generic_to_shared(g2land + e * 32 + lane), gb + (size_t)ixp[e] * B + i);
        asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) g2[e] = (e < kk) ? __ldcg(gb + (size_t)ixp[e] * B + i) : 0;
      }
    }
    // far corrections of the next tile (posted by the helpers): asked for now, looked at when the tile starts
    if constexpr (PRIM) if (t + 1 < T && !(abl & 512)) {
      const int dmax1 = min(DC, t + 1);
#pragma unroll
      for (int q = 2; q < kMaxDC; ++q)
        cwn[q] = (q + 1 <= dmax1) ? ld_relaxed_u64(p.corr + ((size_t)(t + 1) * DC + q) * B + i) : kCorrEmpty;
    }
    if (tid == 0) { hb::mbar_arrive(empty + b); HB_TRACE(t, 5); }
    HB_SPHASE(7);
  }
#undef HB_SPHASE
  if (dead) atomicCAS(ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
  if (tid == 0) {
    for (int k = 0; k < 16; ++k) p.out->phase_clk[0][k] = pc[k];
    p.out->phase_clk[1][0] = repaired;
    p.out->phase_clk[1][1] = generic_tiles;
    atomicAdd(&p.out->rounds, rounds_total);
    atomicAdd(&p.out->pad, repaired);
    atomicAdd(&p.out->n_changed, changed_total);
  }
}

template <int NF>
__device__ void serial_role(const SweepParams& pin, uint8_t* smem) {
  constexpr int B = kSerialB;
  using G = SerialGeo<B>;
  for (int w = threadIdx.x; w < (int)(sizeof(SweepParams) / 4); w += blockDim.x) ((int*)&g_serial_ps)[w] = ((const int*)&pin)[w];
  __syncthreads();
  const SweepParams& p = g_serial_ps;
  const int T = p.T;
  const int tid = threadIdx.x;
  if (tid >= 2 * B) return;
  if (p.dbg & 48) return;   // timing experiments of the streaming side: the helpers publish empty tiles
  // threads: [0, B) primary (one per SNP), [B, 2B-32) secondary (odd candidates of SNP i < B-32), last warp = loader
  constexpr int NTS = G::NTS;
  int* ctrl = p.ctrl;
  constexpr PkgLayout L = G::L;
  constexpr PrivLayout P = G::P;
  uint8_t* priv = smem + G::priv0;
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  long long* pc = (long long*)(priv + P.o_pc);
  uint64_t* full = (uint64_t*)(priv + P.o_full);
  uint64_t* empty = (uint64_t*)(priv + P.o_empty);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { hb::mbar_init(full + b, 1); hb::mbar_init(empty + b, 1); }
    hb::mbar_fence_init();
    for (int k = 0; k < 16; ++k) gctl[k] = 0;
    for (int k = 0; k < 16; ++k) pc[k] = 0;
    pc[16] = clock64();
  }
  hb::named_bar_sync(5, 2 * B);
  if (tid >= NTS) {
    // ---------------- loader: package of tile t -> buffer t & 1, as soon as the buffer is free and the flag is up
    if (tid == NTS) {
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t >= 2 && !mbar_wait(empty + b, (uint32_t)(((t >> 1) - 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
        int f = ld_acquire_s32(p.pkg_flag + t);
        if (f == 0) {
          Waiter w;
          do {
            __nanosleep(40);
            if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) return;
            f = ld_acquire_s32(p.pkg_flag + t);
          } while (f == 0);
        }
        HB_TRACE(t, 11);
        asm volatile("fence.proxy.async;" ::: "memory");   // the helpers' (generic) stores before the bulk (async) read
        const int nrow = (f & kPkgNoRows) ? 0 : ((f & 0xffff) - 1);
        const uint32_t bytes = L.fixed_bytes + (uint32_t)nrow * L.row_bytes;
        hb::mbar_arrive_expect_tx(full + b, bytes);
        hb::tma_load_1d(smem + (size_t)b * L.stride, p.pkg + (size_t)t * p.pkg_stride, bytes, full + b);
      }
    }
    return;
  }
  if (tid < B) serial_threads<NF, true, B>();
  else serial_threads<NF, false, B>();
}

// ------------------------------------------------------------------------------------------
// helper CTAs
// ------------------------------------------------------------------------------------------
template <int NF>
__device__ __noinline__ void helper_role(const SweepParams& pin, uint8_t* smem) {
  __shared__ SweepParams ps;
  for (int w = threadIdx.x; w < (int)(sizeof(SweepParams) / 4); w += blockDim.x) ((int*)&ps)[w] = ((const int*)&pin)[w];
  __syncthreads();
  const SweepParams& p = ps;
  const int B = p.B, D = p.D, T = p.T, F = p.F, model = p.model;
  const int tid = threadIdx.x;
  if (tid >= 2 * B) return;
  const int h = tid / B, i = tid - h * B, warp = i >> 5, lane = i & 31, nwarp = B / 32;
  const bool prim = (h == 0);
  const int helper = (int)blockIdx.x - p.scalar0 - 1, NH = p.NH;
  int* ctrl = p.ctrl;
  // ---- shared memory: the carve-up of scalar_role (candidate arrays, partial sums, row buffers, chain matrices)
  CandSet cs;
  int *wcnt, *slot_of, *slot_snp;
  volatile int* gctl;
  int32_t *rows0, *rows1;
  double *coef, *cmat;
  {
    cs = make_candset(smem, B);
    double* d = (double*)smem;
    d += 8 * (size_t)B;
    int* ip = (int*)d;
    slot_of = ip + 4 * B; slot_snp = ip + 5 * B; ip += 6 * (size_t)B;
    wcnt = ip; ip += 64;
    gctl = ip; ip += 16;
    uint8_t* rb = smem + scalar_fixed_bytes(B);
    coef = (double*)(rb - kChainCoefBytes);
    cmat = coef + 32 * 32;
    rows0 = (int32_t*)rb;
    rows1 = rows0 + (size_t)p.KROW * B;
  }
  const int KROW = p.KROW, KC = p.KROW_S;
  const PkgLayout L = pkg_layout(B, KC);
  const int NT2 = 2 * B;
  if (tid < 8) gctl[tid] = 0;
  if (tid == 8) gctl[4] = hb::ld_relaxed(p.miss_tile);
  hb::named_bar_sync(1, NT2);

  const size_t mp = p.m_pad;
  const int nf = (model == HB_MODEL_R) ? F : 2;
  const int DC = D - 1;
  bool dead = false;

  for (int t = helper; t < T; t += NH) {
    const int j = t * B + i;
    if (p.dbg & 48) {
      if (!(p.dbg & 32) && prim) {
        Waiter w;
        while ((ld_relaxed_u64(p.dacc + j) & 0xffull) != (unsigned long long)p.S)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) break;
      }
      hb::named_bar_sync(1, NT2);
      if (tid == 0) st_relaxed_s32(p.tile_cnt + t, 0);
      continue;
    }
    // ================= phase P
    bool act = false;
    double xx = 0.0, gold = 0.0;
    double TL[NF - 1], TH[NF - 1], civ[NF - 1], csdz[NF - 1];
#pragma unroll
    for (int q = 0; q < NF - 1; ++q) { TL[q] = -1.0; TH[q] = -1.0; civ[q] = 0.0; csdz[q] = 0.0; }
    if (prim) {
      act = (j < p.m) && __ldcg(p.active + j);
      xx = __ldcg(p.xpx + j);
      gold = __ldcg(p.g + j);
#pragma unroll
      for (int q = 0; q < NF - 1; ++q)
        if (q < nf - 1) {
          TL[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q, mp, j));
          TH[q] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * q + 1, mp, j));
          civ[q] = __ldcg(p.prm + prm_idx(4 + 4 * q, mp, j));
          csdz[q] = __ldcg(p.prm + prm_idx(5 + 4 * q, mp, j));
        }
    }
    bool spec_exact = false;
    auto classify = [&](double rhs) -> int {
      const double rr = rhs * rhs;
      const int c0 = thr_class<NF>(nf, rr, TL, TH);
      if (c0 >= 0) return c0;
      spec_exact = true;
      return classify_exact<NF>(p.prm, mp, j, nf, rr, p.logpi0);
    };
    // the dots: complete when the arrival count in the low byte equals the number of slabs
    double base0 = 0.0;
    if (prim) {
      unsigned long long w = ld_relaxed_u64(p.dacc + j);
      if ((w & 0xffull) != (unsigned long long)p.S) {
        Waiter wt;
        do {
          __nanosleep(200);
          if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { dead = true; break; }
          w = ld_relaxed_u64(p.dacc + j);
        } while ((w & 0xffull) != (unsigned long long)p.S);
      }
      long long fx = (long long)w >> 8;
      if (p.world > 1) {
        // row-sharded runs: this rank's exact integer part of the dot goes to every rank's accumulator (NVLink peer
        // atomics); every rank ends up with the same sum and takes the same decisions
        const unsigned long long part = ((unsigned long long)fx << 4) + 1ull;
        for (int g = 0; g < p.world; ++g)
          asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p.peer_acc[g] + j), "l"(part) : "memory");
        unsigned long long w2;
        Waiter wt;
        for (;;) {
          asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w2) : "l"(p.peer_acc[p.rank] + j) : "memory");
          if ((w2 & 0xfull) == (unsigned long long)p.world) break;
          __nanosleep(200);
          if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { dead = true; break; }
        }
        fx = (long long)w2 >> 4;
      }
      base0 = (double)fx * p.inv_dscale;
    }
    // far corrections that have been posted by now go into the speculation (those of the nearest tiles cannot be
    // there yet: the chain has not reached them)
    double cspec = 0.0;
    if (prim) {
      const int dmax = min(DC, t);
      unsigned long long cw[kMaxDC];
#pragma unroll
      for (int q = 0; q < kMaxDC; ++q) cw[q] = (q >= 2 && q + 1 <= dmax) ? ld_relaxed_u64(p.corr + ((size_t)t * DC + q) * B + i) : kCorrEmpty;
#pragma unroll
      for (int q = kMaxDC - 1; q >= 2; --q)
        if (cw[q] != kCorrEmpty) cspec += __longlong_as_double((long long)cw[q]);
    }
    if (tid == 0) HB_TRACE(t, 0);
    const double addback = (act && gold != 0.0) ? xx * gold : 0.0;
    int cls = 0;
    const double rhs_spec = (base0 - cspec) + addback;
    if (act) cls = classify(rhs_spec);
    int k = 0, myrank = 0, ns = 0;
    bool cand = false, fast = false;
    const bool has1 = (D > 1 && t + 1 < T);
    const int32_t* G0 = p.gram + ((size_t)t * D) * B * B;
    // row set: every candidate, plus -- while the chain has recently needed repairs -- the SNPs within 30 % of their
    // first class boundary; trimmed to the candidates when the set would not fit a package
    const bool widen = (t - gctl[4]) < 16 * NH;   // (read by one thread before the last barrier: the same for all)
    auto select_rows = [&](bool with_near) {
      if (prim) {
        const double rr_spec = rhs_spec * rhs_spec;
        const bool near = with_near && TH[0] > 0.0 && TH[0] < 1e300 && rr_spec >= 0.49 * TH[0];
        const bool want = act && (gold != 0.0 || cls > 0 || near);
        const unsigned bal = __ballot_sync(0xffffffffu, want);
        if (lane == 0) wcnt[warp] = __popc(bal);
        hb::named_bar_sync(4, B);
        int pre = 0, tot = 0;
        for (int w = 0; w < nwarp; ++w) {
          const int c = wcnt[w];
          if (w < warp) pre += c;
          tot += c;
        }
        const int sl = pre + __popc(bal & ((1u << lane) - 1u));
        slot_of[i] = want ? sl : -1;
        if (want) slot_snp[sl] = i;
        if (i == 0) gctl[2] = tot;
      }
      hb::named_bar_sync(1, NT2);
      ns = gctl[2];
    };
    select_rows(widen);
    if (widen && ns > min(KC, KROW)) { hb::named_bar_sync(1, NT2); select_rows(false); }
    fast = (ns <= KROW) && (ns <= KC);
    if (fast) {
      {
        // far blocks of the same rows (the tile after next for the serial CTA, t+3.. for phase C): ask for them in L2
        const int lpr = B / 32;
        const int nfar = min(D, T - t) - 2;
        for (int l = tid; l < ns * lpr * nfar; l += NT2) {
          const int line = l % lpr, sl = (l / lpr) % ns, dt = 2 + l / (lpr * ns);
          const int32_t* a = G0 + (size_t)dt * B * B + (size_t)slot_snp[sl] * B + line * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
      if (prim) gather_rows(rows0, G0, slot_snp, ns, B, i);
      else if (has1) gather_rows(rows1, G0 + (size_t)B * B, slot_snp, ns, B, i);
    }
    // candidate list of the speculated classes
    cand = act && (cls > 0 || gold != 0.0);
    if (prim) {
      const unsigned bal = __ballot_sync(0xffffffffu, cand);
      if (lane == 0) wcnt[warp] = __popc(bal);
      hb::named_bar_sync(4, B);
      int pre = 0;
      k = 0;
      for (int w = 0; w < nwarp; ++w) {
        const int c = wcnt[w];
        if (w < warp) pre += c;
        k += c;
      }
      myrank = pre + __popc(bal & ((1u << lane) - 1u));
      if (i == 0) gctl[1] = k;
      if (cand && fast) {
        cs.idx[myrank] = i;
        cs.gold[myrank] = gold;
        cs.cls[myrank] = cls;
        cs.slot[myrank] = slot_of[i];
        double iv = 0.0, sdz = 0.0;
#pragma unroll
        for (int kk = 1; kk < NF; ++kk)
          if (kk == cls) { iv = civ[kk - 1]; sdz = csdz[kk - 1]; }
        cs.iv[myrank] = iv;
        cs.sdz[myrank] = sdz;
      }
    }
    hb::named_bar_sync(1, NT2);
    k = gctl[1];
    bool m_ok = false;
    if (fast) {
      const int kk = min(k, 32);
      for (int e = tid; e < 32 * 32; e += NT2) {
        const int lp = e >> 5, sc = e & 31;
        double v = 0.0;
        if (lp < sc && sc < kk) v = gram_as_double(rows0[(size_t)cs.slot[lp] * B + cs.idx[sc]]) * (-cs.iv[sc]);
        coef[e] = v;
      }
      hb::named_bar_sync(1, NT2);
      if (k > 0 && k <= 32 && !(p.dbg & 128)) {
        if (!prim && warp == 0) chain_build_matrix(coef, cmat, k, lane);
        m_ok = true;
      }
      hb::named_bar_sync(1, NT2);
    }
    // ---- the package
    {
      uint8_t* pk = p.pkg + (size_t)t * p.pkg_stride;
      if (prim) {
        double lo = -1.0, hi = HUGE_VAL;
        if (act) {
#pragma unroll
          for (int q = 0; q < NF - 1; ++q) {
            if (q < cls) lo = fmax(lo, TH[q]);
            if (q == cls && q < nf - 1) hi = TL[q];
          }
        }
        ((double*)(pk + L.o_base0))[i] = base0;
        ((double*)(pk + L.o_addback))[i] = addback;
        ((double*)(pk + L.o_lo))[i] = lo;
        ((double*)(pk + L.o_hi))[i] = hi;
        ((int*)(pk + L.o_info))[i] = cls | ((int)cand << 4) | ((int)act << 5) | ((int)(spec_exact && act) << 6) | (myrank << 8);
        ((int*)(pk + L.o_slotof))[i] = fast ? slot_of[i] : -1;
        if (i == 0) {
          int* hdr = (int*)(pk + L.o_hdr);
          hdr[PK_K] = k; hdr[PK_NS] = fast ? ns : 0; hdr[PK_FAST] = fast ? 1 : 0; hdr[PK_MOK] = m_ok ? 1 : 0;
          hdr[PK_TILE] = t; hdr[PK_HAS1] = has1 ? 1 : 0;
        }
      }
      if (fast) {
        for (int e = tid; e < k; e += NT2) {
          ((double*)(pk + L.o_iv))[e] = cs.iv[e];
          ((double*)(pk + L.o_sdz))[e] = cs.sdz[e];
          ((double*)(pk + L.o_gold))[e] = cs.gold[e];
          ((int*)(pk + L.o_idx))[e] = cs.idx[e];
          ((int*)(pk + L.o_cls))[e] = cs.cls[e];
          ((int*)(pk + L.o_slot))[e] = cs.slot[e];
        }
        if (m_ok)
          for (int e = tid; e < 32 * 33; e += NT2) ((double*)(pk + L.o_M))[e] = cmat[e];
        // rows: slot s = [diagonal block row | next block row], 16 bytes per store
        const int q4 = B / 4;   // int4 per block row
        int4* dst = (int4*)(pk + L.o_rows);
        for (int e = tid; e < ns * 2 * q4; e += NT2) {
          const int sl = e / (2 * q4), c = e - sl * 2 * q4;
          const int blk = c / q4, cc = c - blk * q4;
          int4 v = make_int4(0, 0, 0, 0);
          if (blk == 0) v = ((const int4*)(rows0 + (size_t)sl * B))[cc];
          else if (has1) v = ((const int4*)(rows1 + (size_t)sl * B))[cc];
          dst[e] = v;
        }
      }
      __threadfence();
      hb::named_bar_sync(1, NT2);
      if (tid == 0) {
        st_release_s32(p.pkg_flag + t, fast ? (1 + ns) : (1 | kPkgNoRows));
        HB_TRACE(t, 1);
      }
    }
    if (hb::named_bar_or(1, NT2, dead)) { dead = true; break; }
    // ================= phase C: the corrections this tile owes to the tiles t+3 .. t+D-1, once it is final
    {
      int* fidx = cs.idx;        // the final list (the candidate arrays are free now)
      double* fdel = cs.delta;
      if (tid < 32) {
        int cnt = -1;
        Waiter w;
        for (;;) {
          int c = -1;
          if (lane == 0) c = hb::ld_relaxed(p.tile_cnt + t);
          cnt = __shfl_sync(0xffffffffu, c, 0);
          if (cnt >= 0) break;
          __nanosleep(100);
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { cnt = -2; break; }
        }
        if (lane == 0) gctl[3] = cnt;
      }
      hb::named_bar_sync(1, NT2);
      const int kf = gctl[3];
      if (kf < 0) { dead = true; break; }
      if (tid == 0) HB_TRACE(t, 8);
      for (int e = tid; e < kf; e += NT2) {
        int jl = -1;
        unsigned long long dw = kCorrEmpty;
        Waiter w;
        for (;;) {
          if (jl < 0) jl = hb::ld_relaxed(p.q_snp + (size_t)t * B + e);
          if (dw == kCorrEmpty) dw = ld_relaxed_u64(p.q_delta + (size_t)t * B + e);
          if (jl >= 0 && dw != kCorrEmpty) break;
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { dead = true; break; }
        }
        fidx[e] = jl - t * B;
        fdel[e] = __longlong_as_double((long long)dw);
      }
      if (hb::named_bar_or(1, NT2, dead)) { dead = true; break; }
      for (int dt = 3 + h; dt < D; dt += 2) {
        if (t + dt >= T) break;
        const double cv = band_correction_raw(fidx, fdel, kf, G0 + (size_t)dt * B * B, B, i);
        post_corr(p.corr + ((size_t)(t + dt) * DC + (dt - 1)) * B + i, cv);
        if (tid == 0 && dt == 3) HB_TRACE(t, 9);
      }
      if (tid == 0) { gctl[4] = hb::ld_relaxed(p.miss_tile); HB_TRACE(t, 10); }
      hb::named_bar_sync(1, NT2);   // the arrays are free again
    }
  }
  if (dead) atomicCAS(ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
}

}  // namespace hbk
