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
  L.o_base0 = o; o += 8 * B;          // x'r of the stale residual + xpx * g (the corrections are still to be subtracted)
  L.o_addback = L.o_base0;
  L.o_lo = o; o += 4 * B;             // interval of rhs^2 on which the speculated class holds, as floats rounded inwards
  L.o_hi = o; o += 4 * B;
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
constexpr int kMaxDC = 7;    // serial mode: lag of at most 8 tiles (landing zone of a tile's far-correction slots)
constexpr int kSerialB = 256;   // serial mode is compiled for tiles of 256 SNPs (every offset below is a constant)

// private shared memory of the serial CTA behind the two package buffers (byte offsets from its start)
constexpr int kNear2 = 33;   // changes of a tile whose Gram entries towards t+2 travel in registers (3 subsets of 11)
struct PrivLayout {
  uint32_t o_part_rhs, o_part_corr, o_coldbuf, o_c2part, o_coef, o_dl2, o_ix2, o_new_snp, o_rank, o_stage, o_wcnt, o_gctl, o_pc,
      o_full, o_empty, o_farfull, o_stagefull, bytes;
};
__host__ __device__ constexpr PrivLayout priv_layout(int B) {
  PrivLayout P{};
  uint32_t o = 0;
  P.o_part_rhs = o;                   // (unused: the primary threads keep their sums)
  P.o_part_corr = o; o += 8 * B;      // corrections owed to the next tile, from the threads that summed them
  P.o_c2part = o; o += 3 * 8 * B;     // [3][B] partial sums of the correction towards t+2 (three subsets of the changes)
  P.o_coldbuf = o; o += 2 * 8 * B;    // [2][B] sum of the far corrections (dt >= 3) of a tile, by tile parity (input warp)
  P.o_coef = 0;                       // (the chain coefficients live in the package's M area: never both in use)
  P.o_dl2 = o; o += 2 * 36 * 8;       // first kNear2 changes of the last two tiles ...
  P.o_ix2 = o; o += 2 * 36 * 4;       // ... and their SNPs
  P.o_new_snp = o; o += 4 * B;
  P.o_rank = o; o += 4 * B;           // candidates before SNP i (repair rounds: the secondary threads read it)
  P.o_stage = o; o += (kMaxDC - 2) * 8 * B;   // landing zone of a tile's far-correction slots (input warp)
  P.o_wcnt = o; o += 64 * 4;
  P.o_gctl = o; o += 16 * 4;          // [1] k, [2] row slots, [8+b] k of the tile of parity b
  P.o_pc = o; o += 18 * 8;
  P.o_full = o; o += 2 * 8;
  P.o_empty = o; o += 2 * 8;
  P.o_farfull = o; o += 2 * 8;
  P.o_stagefull = o; o += 8;
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
  double rhs0, rhs1, lo, hi, gold, myiv, mysdz, c1next, my_delta, my_gnew;   // rhs1: exact right-hand side of the first round
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

// Exact right-hand sides and the corrections owed to the next tile, from the row slots in shared memory:
//   primary thread i     sum over the changes before SNP i of G0[c][i] * delta_c        (returned)
//   secondary thread i   sum over all changes of G1[c][i] * delta_c -> part_corr[i]     (i < B - 32)
//   primary warp 0       the same for the last 32 SNPs, which have no secondary thread (its own sums are the shortest:
//                        hardly any change precedes the first 32 SNPs of a tile)
// Four partial sums per thread (changes e, e+4, ... each), added at the end: the dependent chain of fmas is a quarter of
// the list.
template <bool PRIM, int B>
__device__ __forceinline__ double tile_sums(const int32_t* rows, const double* c_delta, const int* c_slot, int k, int myrank,
                                            bool has1, int i, int warp, int lane, double* part_corr) {
  constexpr int RS = 2 * B;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if constexpr (PRIM) {
    const int kw = __shfl_sync(0xffffffffu, myrank, 31);   // changes before the warp's last SNP: nobody needs more
    int e = 0;
    for (; e + 3 < kw; e += 4) {
      const double d0 = c_delta[e], d1 = c_delta[e + 1], d2 = c_delta[e + 2], d3 = c_delta[e + 3];
      const int s0 = c_slot[e], s1 = c_slot[e + 1], s2 = c_slot[e + 2], s3 = c_slot[e + 3];
      a0 = fma(e < myrank ? gram_as_double(rows[s0 * RS + i]) : 0.0, d0, a0);
      a1 = fma(e + 1 < myrank ? gram_as_double(rows[s1 * RS + i]) : 0.0, d1, a1);
      a2 = fma(e + 2 < myrank ? gram_as_double(rows[s2 * RS + i]) : 0.0, d2, a2);
      a3 = fma(e + 3 < myrank ? gram_as_double(rows[s3 * RS + i]) : 0.0, d3, a3);
    }
    for (; e < kw; ++e) a0 = fma(e < myrank ? gram_as_double(rows[c_slot[e] * RS + i]) : 0.0, c_delta[e], a0);
    const double prhs = (a0 + a1) + (a2 + a3);
    if (warp == 0 && has1) {
      const int i2 = B - 32 + lane;
      double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
      int f = 0;
      for (; f + 3 < k; f += 4) {
        b0 = fma(gram_as_double(rows[c_slot[f] * RS + B + i2]), c_delta[f], b0);
        b1 = fma(gram_as_double(rows[c_slot[f + 1] * RS + B + i2]), c_delta[f + 1], b1);
        b2 = fma(gram_as_double(rows[c_slot[f + 2] * RS + B + i2]), c_delta[f + 2], b2);
        b3 = fma(gram_as_double(rows[c_slot[f + 3] * RS + B + i2]), c_delta[f + 3], b3);
      }
      for (; f < k; ++f) b0 = fma(gram_as_double(rows[c_slot[f] * RS + B + i2]), c_delta[f], b0);
      part_corr[i2] = (b0 + b1) + (b2 + b3);
    } else if (warp == 0) {
      part_corr[B - 32 + lane] = 0.0;
    }
    return prhs;
  } else {
    if (has1) {
      int f = 0;
#pragma unroll 1
      for (; f + 3 < k; f += 4) {
        a0 = fma(gram_as_double(rows[c_slot[f] * RS + B + i]), c_delta[f], a0);
        a1 = fma(gram_as_double(rows[c_slot[f + 1] * RS + B + i]), c_delta[f + 1], a1);
        a2 = fma(gram_as_double(rows[c_slot[f + 2] * RS + B + i]), c_delta[f + 2], a2);
        a3 = fma(gram_as_double(rows[c_slot[f + 3] * RS + B + i]), c_delta[f + 3], a3);
      }
#pragma unroll 1
      for (; f < k; ++f) a0 = fma(gram_as_double(rows[c_slot[f] * RS + B + i]), c_delta[f], a0);
    }
    part_corr[i] = (a0 + a1) + (a2 + a3);
    return 0.0;
  }
}

// The same for a package as the helper ships it: the candidates' rows are slots 0 .. k-1 in order, so the slots need not be
// looked up, and the changes are read two at a time (half as many shared-memory instructions: the chain's threads all run
// this at once and the load/store unit is what bounds it).
template <bool PRIM, int B>
__device__ __forceinline__ double tile_sums_seq(const int32_t* rows, const double* c_delta, int k, int myrank, bool has1, int i,
                                                int warp, int lane, double* part_corr) {
  constexpr int RS = 2 * B;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if constexpr (PRIM) {
    const int kw = __shfl_sync(0xffffffffu, myrank, 31);   // changes before the warp's last SNP: nobody needs more
    const int32_t* r0 = rows + i;
    int e = 0;
#pragma unroll 1
    for (; e + 3 < kw; e += 4) {
      const double2 da = *(const double2*)(c_delta + e), db = *(const double2*)(c_delta + e + 2);
      const int g0 = r0[e * RS], g1 = r0[(e + 1) * RS], g2 = r0[(e + 2) * RS], g3 = r0[(e + 3) * RS];
      a0 = fma(e < myrank ? gram_as_double(g0) : 0.0, da.x, a0);
      a1 = fma(e + 1 < myrank ? gram_as_double(g1) : 0.0, da.y, a1);
      a2 = fma(e + 2 < myrank ? gram_as_double(g2) : 0.0, db.x, a2);
      a3 = fma(e + 3 < myrank ? gram_as_double(g3) : 0.0, db.y, a3);
    }
#pragma unroll 1
    for (; e < kw; ++e) a0 = fma(e < myrank ? gram_as_double(r0[e * RS]) : 0.0, c_delta[e], a0);
    const double prhs = (a0 + a1) + (a2 + a3);
    if (warp == 0) {
      double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
      if (has1) {
        const int32_t* r1 = rows + B + (B - 32 + lane);
        int f = 0;
#pragma unroll 1
  #pragma unroll 1
      for (; f + 3 < k; f += 4) {
          const double2 da = *(const double2*)(c_delta + f), db = *(const double2*)(c_delta + f + 2);
          b0 = fma(gram_as_double(r1[f * RS]), da.x, b0);
          b1 = fma(gram_as_double(r1[(f + 1) * RS]), da.y, b1);
          b2 = fma(gram_as_double(r1[(f + 2) * RS]), db.x, b2);
          b3 = fma(gram_as_double(r1[(f + 3) * RS]), db.y, b3);
        }
#pragma unroll 1
        for (; f < k; ++f) b0 = fma(gram_as_double(r1[f * RS]), c_delta[f], b0);
      }
      part_corr[B - 32 + lane] = (b0 + b1) + (b2 + b3);
    }
    return prhs;
  } else {
    if (has1) {
      const int32_t* r1 = rows + B + i;
      int f = 0;
#pragma unroll 1
      for (; f + 3 < k; f += 4) {
        const double2 da = *(const double2*)(c_delta + f), db = *(const double2*)(c_delta + f + 2);
        a0 = fma(gram_as_double(r1[f * RS]), da.x, a0);
        a1 = fma(gram_as_double(r1[(f + 1) * RS]), da.y, a1);
        a2 = fma(gram_as_double(r1[(f + 2) * RS]), db.x, a2);
        a3 = fma(gram_as_double(r1[(f + 3) * RS]), db.y, a3);
      }
#pragma unroll 1
      for (; f < k; ++f) a0 = fma(gram_as_double(r1[f * RS]), c_delta[f], a0);
    }
    part_corr[i] = (a0 + a1) + (a2 + a3);
    return 0.0;
  }
}

// Candidates beyond the first 32 of a package (the solved chain matrix covers 32): one after the other, each from the
// changes of all candidates before it -- lane l multiplies the changes of candidates l and l + 32 with their Gram entries
// towards candidate s, a shuffle tree adds the products.  A rolled loop: a few dozen instructions that ~7 % of the
// tiles run for a handful of candidates (the step-by-step chain of scalar_role unrolls to 2000 instructions, which
// would push the chain's loop out of the instruction cache).
template <int B>
__device__ __forceinline__ void chain_tail(const int32_t* rows, const int* c_idx, const int* c_cls, const double* c_rhs0,
                                           const double* c_iv, const double* c_sdz, const double* c_gold, double* c_delta,
                                           double* c_gnew, int k, int lane) {
  constexpr int RS = 2 * B;
  for (int s = 32; s < k; ++s) {
    const int col = c_idx[s];
    double part = 0.0;
    if (lane < s) part = gram_as_double(rows[lane * RS + col]) * c_delta[lane];
    if (lane + 32 < s) part = fma(gram_as_double(rows[(lane + 32) * RS + col]), c_delta[lane + 32], part);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) {
      const double gold = c_gold[s];
      const double e = fma(c_rhs0[s] - part, c_iv[s], c_sdz[s]) - gold;
      c_delta[s] = e;
      c_gnew[s] = (c_cls[s] > 0) ? gold + e : 0.0;
    }
    __syncwarp();
  }
}

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
  double* coef = (double*)(pk + L.o_M);   // chain coefficients of the first 32 candidates: in the (then unused) M area
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
    // a class may differ from its speculation after the first round: every class from the round's exact right-hand side
    if (prim && st.act) {
      st.cls2 = classify_full(st.rhs1);
      if (st.cls2 != st.cls) { st.cls = st.cls2; load_draw(); }
    }
    if (!generic) compact_rows();
    if (generic) { hb::named_bar_sync(1, NTS); compact_generic(); }
  }
  for (;;) {
    ++st.nrounds;
    if (st.cand && prim) cs.rhs0[st.myrank] = st.rhs0;
    hb::named_bar_sync(1, NTS);
    double prhs = 0.0;
    if (!generic) {
      if (tid < 32 && k > 0) chain_candidates<true>(cs, k, G0, rows, B, lane, coef, RS);
      hb::named_bar_sync(1, NTS);
      prhs = tile_sums<PRIM, B>(rows, cs.delta, cs.slot, k, st.myrank, has1, i, warp, lane, part_corr);
    } else {
      const double sv = slow_chain_and_sums_cs(cs, k, st.myrank, G0, B, i, h, has1, false, model, NTS);
      if (prim) prhs = sv; else part_corr[i] = sv;
      if (tailw) part_corr[i] = has1 ? band_correction(cs, k, G0 + (size_t)B * B, B, i) : 0.0;
    }
    hb::named_bar_sync(1, NTS);
    int cls2 = st.cls;
    if (prim) {
      const double rhs = st.rhs0 - prhs;
      st.c1next = part_corr[i];
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
  if (tid < 36) {
    ((double*)(priv + P.o_dl2))[(t & 1) * 36 + tid] = (tid < k) ? cs.delta[tid] : 0.0;
    ((int*)(priv + P.o_ix2))[(t & 1) * 36 + tid] = (tid < k) ? cs.idx[tid] : 0;
  }
  st.k = k; st.ns = ns; st.generic = generic;
}

// Bounded wait on an mbarrier without a function call: a call inside the chain's loop would make the compiler park the
// registers that live across it (the Gram words towards t+2, ...) in local memory, and with 225 KB of shared memory in use
// local memory is an L2 round trip.
__device__ __forceinline__ bool mbar_wait_nocall(uint64_t* bar, uint32_t parity, int* ctrl, int code) {
  if (hb::mbar_try_wait(bar, parity)) return true;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  while (!hb::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3f) == 0) {
      if (*((volatile int*)(ctrl + 1)) != 0) return false;
      const unsigned long long now = gtimer();
      if (t0 == 0) t0 = now;
      else if ((long long)(now - t0) > kTimeoutNs) { atomicCAS(ctrl + 1, 0, code); return false; }
    }
  }
  return true;
}

// What a chain thread carries from tile to tile.
constexpr int kNGA = 6, kNGB = 5;   // Gram words towards t+2 per thread: batch A (changes 0..17 of a tile), batch B (18..32)
static_assert(3 * (kNGA + kNGB) == kNear2, "batches cover the near list");
struct SerialCarry {
  int t;              // next tile
  bool dead;
  double corr1;       // correction owed by the previous tile to this thread's SNP (never leaves the thread)
  double corr2;       // correction owed by the tile before that (summed from c2part at the end of the previous tile)
  int4 gA[kNGA], gB[kNGB];   // (secondary warps 0..5) Gram words towards t+2 on their way
};

// End of a tile, once it is final: effects, classes and the tile's changes for the AXPY warps and the helpers; the
// correction towards t+2 (batch B of the previous tile's list summed, batch A of this tile's list asked for); the package
// buffer handed back to the input warp.
template <bool PRIM, int B>
__device__ __forceinline__ void serial_tile_end(int t, int k, int nrounds, bool cand, bool act, int myrank, int cls, double my_delta,
                                                double my_gnew, double& corr2, int4 (&gA)[kNGA], int4 (&gB)[kNGB], bool timing) {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PrivLayout P = G::P;
  constexpr int NTS = G::NTS;
  constexpr int h = PRIM ? 0 : 1;
  const SweepParams& p = g_serial_ps;
  const int D = p.D, T = p.T;
  const int tid = threadIdx.x, i = tid - h * B, warp = i >> 5, lane = i & 31;
  const int b = t & 1, j = t * B + i;
  const int abl = p.dbg;
  uint8_t* priv = smem + G::priv0;
  double* c2part = (double*)(priv + P.o_c2part);
  const double* dl2 = (const double*)(priv + P.o_dl2);
  const int* ix2 = (const int*)(priv + P.o_ix2);
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  long long* pc = (long long*)(priv + P.o_pc);
  uint64_t* empty = (uint64_t*)(priv + P.o_empty);
  const bool g2on = !PRIM && warp < 6;
  const int g2s = warp >> 1, g2j = (warp & 1) * 32 + lane;
#define HB_SPHASE(n) do { if (timing) { const long long _now = clock64(); pc[n] += _now - pc[16]; pc[16] = _now; } } while (0)
  if (tid == 0) {
    gctl[10] += nrounds; gctl[11] += k;
    if (nrounds > 1) { gctl[12]++; st_relaxed_s32(p.miss_tile, t); }
  }
  HB_SPHASE(5);
  if (tid == 0) HB_TRACE(t, 3);
  if constexpr (PRIM) {
    if (cand) {
      st_relaxed_u64(p.q_delta + (size_t)t * B + myrank, (unsigned long long)__double_as_longlong(my_delta));
      st_relaxed_s32(p.q_snp + (size_t)t * B + myrank, j);
      p.g[j] = my_gnew;
    }
    if (i == 0) { st_relaxed_s32(p.tile_cnt + t, k); HB_TRACE(t, 4); gctl[8 + b] = k; }
    if (act) p.tracker[j] = cls;
  }
  HB_SPHASE(6);
  const bool c2on = t >= 1 && t + 1 < T && D > 2 && !(abl & 256);
  if constexpr (!PRIM) {
    if (g2on && c2on && gctl[8 + (b ^ 1)] > 3 * kNGA) {
      // batch B of tile t-1's list: added to the partial sums of batch A (same thread, same entries)
      const int kp = min(gctl[8 + (b ^ 1)], kNear2);
      const double* dprev = dl2 + (b ^ 1) * 36;
      double* cp = c2part + g2s * B + 4 * g2j;
      double a0 = cp[0], a1 = cp[1], a2 = cp[2], a3 = cp[3];
#pragma unroll
      for (int r = 0; r < kNGB; ++r) {
        const int e = 3 * kNGA + g2s + 3 * r;
        if (e < kp) {
          const double d = dprev[e];
          a0 = fma(gram_as_double(gB[r].x), d, a0); a1 = fma(gram_as_double(gB[r].y), d, a1);
          a2 = fma(gram_as_double(gB[r].z), d, a2); a3 = fma(gram_as_double(gB[r].w), d, a3);
        }
      }
      cp[0] = a0; cp[1] = a1; cp[2] = a2; cp[3] = a3;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's accesses to the package buffer before its refill
  hb::named_bar_sync(1, NTS);   // dl2 / ix2 / gctl[8+b] of this tile and the partial sums towards t+1 are visible
  if constexpr (PRIM) {
    if (c2on) {
      corr2 = (c2part[i] + c2part[B + i]) + c2part[2 * B + i];
      const int kp = gctl[8 + (b ^ 1)];
      if (kp > kNear2) {
        // rare: the rest of the list from the published queue of tile t-1
        const int32_t* gb = p.gram + ((size_t)(t - 1) * D + 2) * B * B;
        for (int e = kNear2; e < kp; ++e) {
          const int sn = hb::ld_relaxed(p.q_snp + (size_t)(t - 1) * B + e) - (t - 1) * B;
          const double d = __longlong_as_double((long long)ld_relaxed_u64(p.q_delta + (size_t)(t - 1) * B + e));
          corr2 = fma(gram_as_double(__ldcg(gb + (size_t)sn * B + i)), d, corr2);
        }
      }
    } else {
      corr2 = 0.0;
    }
  } else {
    // (every path assigns the words: a conditional assignment would keep the old values alive around the whole loop
    // body and the compiler would park them in local memory)
    const bool on = g2on && t + 2 < T && D > 2 && !(abl & 256);
    const int32_t* gb = p.gram + ((size_t)t * D + 2) * B * B + 4 * g2j;
    const int* ixp = ix2 + b * 36;
    const int kk = on ? min(k, kNear2) : 0;
#pragma unroll
    for (int r = 0; r < kNGA; ++r) {
      const int e = g2s + 3 * r;
      gA[r] = (e < kk) ? __ldcg((const int4*)(gb + (size_t)ixp[e] * B)) : make_int4(0, 0, 0, 0);
    }
  }
  if (tid == 0) { hb::mbar_arrive(empty + b); HB_TRACE(t, 5); }
  HB_SPHASE(7);
#undef HB_SPHASE
}

// The chain itself: tiles from c.t on, back to back, until the last tile or until a tile needs one of the rare paths
// (returns 1 with the tile's state in st; the caller runs serial_slow and serial_tile_end for it and comes back).  No
// function is called from here: a call inside the loop makes the compiler park whatever lives across it in local memory,
// and with 225 KB of shared memory in use local memory is an L2 round trip.
template <int NF, bool PRIM, int B>
__device__ __noinline__ int serial_fast(SerialCarry& c, SerialTile& st) {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PkgLayout L = G::L;
  constexpr PrivLayout P = G::P;
  constexpr int RS = G::RS, NTS = G::NTS;
  constexpr bool prim = PRIM;
  constexpr int h = PRIM ? 0 : 1;
  const SweepParams& p = g_serial_ps;
  const int D = p.D, T = p.T;
  const int tid = threadIdx.x;
  const int i = tid - h * B, warp = i >> 5, lane = i & 31;
  // secondary warps 0..5 carry the correction towards t+2: subset g2s of the tile's changes (e = g2s, g2s+3, ...), four
  // SNPs 4*g2j .. 4*g2j+3 of the tile after next per thread (16-byte loads of the Gram rows)
  const bool g2on = !prim && warp < 6;
  const int g2s = warp >> 1, g2j = (warp & 1) * 32 + lane;
  int* ctrl = p.ctrl;
  uint8_t* priv = smem + G::priv0;
  double* part_corr = (double*)(priv + P.o_part_corr);
  const double* coldbuf = (const double*)(priv + P.o_coldbuf);
  double* c2part = (double*)(priv + P.o_c2part);
  double* dl2 = (double*)(priv + P.o_dl2);
  int* ix2 = (int*)(priv + P.o_ix2);
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  long long* pc = (long long*)(priv + P.o_pc);
  uint64_t* full = (uint64_t*)(priv + P.o_full);
  uint64_t* farfull = (uint64_t*)(priv + P.o_farfull);

  const int DC = D - 1;
  bool dead = c.dead;
  double corr1 = c.corr1, corr2 = c.corr2;
  int4 gA[kNGA], gB[kNGB];
#pragma unroll
  for (int e = 0; e < kNGA; ++e) gA[e] = c.gA[e];
#pragma unroll
  for (int e = 0; e < kNGB; ++e) gB[e] = c.gB[e];

  const bool timing = PRIM && tid == 0 && (p.dbg & 64);
#define HB_SPHASE(n) do { if (timing) { const long long _now = clock64(); pc[n] += _now - pc[16]; pc[16] = _now; } } while (0)
  const int abl = p.dbg;   // timing experiments (results are wrong with them): 256 no t+2 correction, 512 no far corrections
  int ret = 0;
  int t = c.t;
  for (; t < T; ++t) {
    const int b = t & 1;
    const int j = t * B + i;
    uint8_t* pk = smem + (size_t)b * L.stride;
    {
      const uint32_t par = (uint32_t)((t >> 1) & 1);
      if (!mbar_wait_nocall(full + b, par, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
      if (tid == 0) HB_TRACE(t, 15);
      if (PRIM && !dead && !mbar_wait_nocall(farfull + b, par, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
    }
    HB_SPHASE(0);
    if (tid == 0) HB_TRACE(t, 2);
    const int* hdr = (const int*)(pk + L.o_hdr);
    const int k = dead ? 0 : hdr[PK_K];
    const int ns0 = dead ? 0 : hdr[PK_NS];
    const bool fast = !dead && hdr[PK_FAST] != 0;
    const bool m_ok = !dead && hdr[PK_MOK] != 0;
    const bool has1 = (D > 1 && t + 1 < T);
    double rhs0 = 0.0, rhs1 = 0.0, c1next = 0.0;
    int cls = 0, myrank = 0;
    bool act = false, cand = false, chk_exact = false;
    double* c_rhs0 = (double*)(pk + L.o_rhs0);
    double* c_delta = (double*)(pk + L.o_delta);
    const int* c_slot = (const int*)(pk + L.o_slot);
    const int32_t* rows = (const int32_t*)(pk + L.o_rows);
    double* coef = (double*)(pk + L.o_M);   // chain coefficients of the first 32 candidates: in the (then unused) M area

    // ---- this SNP's inputs
    if (!dead) {
      const int info = ((const int*)(pk + L.o_info))[i];
      myrank = info >> 8;   // candidates before SNP i
      if constexpr (PRIM) {
        const double base = ((const double*)(pk + L.o_base0))[i];
        cls = info & 15; cand = (info >> 4) & 1; act = (info >> 5) & 1; chk_exact = (info >> 6) & 1;
        const int dmax = min(DC, t);
        double cold = 0.0;
        if (dmax >= 3 && !(abl & 512)) cold = coldbuf[b * B + i];   // far corrections, summed oldest first by the input warp
        if (dmax >= 2 && !(abl & 256)) cold += corr2;
        const double c1 = (dmax >= 1) ? corr1 : 0.0;
        rhs0 = (base - cold) - c1;
      }
    }
    HB_SPHASE(1);

    if (fast) {
      if (prim && cand) c_rhs0[myrank] = rhs0;
    }
    if (hb::named_bar_or(1, NTS, dead)) { dead = true; break; }
    // ---- (secondary warps, while the chain warp works) correction owed by tile t-1 to tile t+1: batch A of the Gram words
    // was asked for at the end of the previous tile; partial sums of the three subsets (each in ascending order of the
    // changes), added up by the receiving SNP's thread at the end of this tile; then batch B of the same list is asked for
    if constexpr (!PRIM) {
      const bool on = g2on && t >= 1 && t + 1 < T && D > 2 && !(abl & 256);
      if (!on) {
#pragma unroll
        for (int r = 0; r < kNGB; ++r) gB[r] = make_int4(0, 0, 0, 0);
      } else {
        const int kp = min(gctl[8 + (b ^ 1)], kNear2);
        const double* dprev = dl2 + (b ^ 1) * 36;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int r = 0; r < kNGA; ++r) {
          const int e = g2s + 3 * r;
          if (e < kp) {
            const double d = dprev[e];
            a0 = fma(gram_as_double(gA[r].x), d, a0); a1 = fma(gram_as_double(gA[r].y), d, a1);
            a2 = fma(gram_as_double(gA[r].z), d, a2); a3 = fma(gram_as_double(gA[r].w), d, a3);
          }
        }
        double* cp = c2part + g2s * B + 4 * g2j;
        cp[0] = a0; cp[1] = a1; cp[2] = a2; cp[3] = a3;
        const int32_t* gb = p.gram + ((size_t)(t - 1) * D + 2) * B * B + 4 * g2j;
        const int* ixp = ix2 + (b ^ 1) * 36;
#pragma unroll
        for (int r = 0; r < kNGB; ++r) {
          const int e = 3 * kNGA + g2s + 3 * r;
          gB[r] = (e < kp) ? __ldcg((const int4*)(gb + (size_t)ixp[e] * B)) : make_int4(0, 0, 0, 0);
        }
      }
    }
    bool slow = !fast;
    if (fast) {
      // ---- first round: the speculated candidate list
      HB_SPHASE(3);
      if (tid < 32 && k > 0) {
        CandSet cs;
        cs.rhs0 = c_rhs0; cs.iv = (double*)(pk + L.o_iv); cs.sdz = (double*)(pk + L.o_sdz);
        cs.gold = (double*)(pk + L.o_gold); cs.delta = c_delta; cs.gnew = (double*)(pk + L.o_gnew);
        cs.idx = (int*)(pk + L.o_idx); cs.cls = (int*)(pk + L.o_cls); cs.slot = (int*)(pk + L.o_slot);
        chain_matvec(cs, min(k, 32), (const double*)(pk + L.o_M), lane);   // (the package always has the matrix of the first 32)
        if (k > 32) chain_tail<B>(rows, cs.idx, cs.cls, c_rhs0, cs.iv, cs.sdz, cs.gold, c_delta, cs.gnew, k, lane);
        if (abl & 2048) chain_matvec(cs, min(k, 32), (const double*)(pk + L.o_M), lane);   // (timing: the chain twice, same result)
      }
      HB_SPHASE(4);
      hb::named_bar_sync(1, NTS);
      HB_SPHASE(8);
      // exact right-hand side of every SNP, and the corrections owed to the next tile
      double prhs = tile_sums_seq<PRIM, B>(rows, c_delta, k, myrank, has1, i, warp, lane, part_corr);
      if (abl & 4096) prhs = tile_sums_seq<PRIM, B>(rows, c_delta, k, myrank, has1, i, warp, lane, part_corr);   // (timing: twice)
      HB_SPHASE(9);
      hb::named_bar_sync(1, NTS);
      HB_SPHASE(10);
      bool bad = false;
      if constexpr (PRIM) {
        rhs1 = rhs0 - prhs;
        c1next = part_corr[i];
        if (act) {
          // the class holds as long as rhs^2 stays inside the interval of the speculated class; outside it (or without a
          // certified interval) the classes are re-decided in serial_slow
          const double rr = rhs1 * rhs1;
          const double lo = (double)((const float*)(pk + L.o_lo))[i], hi = (double)((const float*)(pk + L.o_hi))[i];
          bad = chk_exact || !(rr >= lo && rr <= hi);
        }
      }
      HB_SPHASE(11);
      slow = hb::named_bar_or(1, NTS, bad);
      HB_SPHASE(12);
    }
    if (slow) {
      // the rare paths (no rows in the package / a class may differ): out of line, from the caller
      st.rhs0 = rhs0; st.rhs1 = rhs1; st.lo = 0.0; st.hi = 0.0; st.gold = 0.0; st.myiv = 0.0; st.mysdz = 0.0; st.c1next = c1next;
      st.slot = -1;
      if (PRIM && !dead) {
        // (what the chain did not need: read from the package only now)
        st.slot = ((const int*)(pk + L.o_slotof))[i];
        if (cand && fast) {
          st.gold = ((const double*)(pk + L.o_gold))[myrank];
          st.myiv = ((const double*)(pk + L.o_iv))[myrank];
          st.mysdz = ((const double*)(pk + L.o_sdz))[myrank];
        }
      }
      st.my_delta = 0.0; st.my_gnew = 0.0;
      st.cls = cls; st.cls2 = -1; st.myrank = myrank; st.k = k; st.ns = ns0; st.nrounds = fast ? 1 : 0;
      st.act = act; st.cand = cand; st.chk_exact = chk_exact; st.generic = !fast; st.m_ok = m_ok; st.dead = false; st.has1 = has1;
      ret = 1;
      break;
    }
    double my_delta = 0.0, my_gnew = 0.0;
    if (prim && cand) { my_delta = c_delta[myrank]; my_gnew = ((const double*)(pk + L.o_gnew))[myrank]; }
    if (tid < 36) {
      dl2[b * 36 + tid] = (tid < k) ? c_delta[tid] : 0.0;
      ix2[b * 36 + tid] = (tid < k) ? ((const int*)(pk + L.o_idx))[tid] : 0;
    }
    corr1 = c1next;
    serial_tile_end<PRIM, B>(t, k, 1, cand, act, myrank, cls, my_delta, my_gnew, corr2, gA, gB, timing);
  }
#undef HB_SPHASE
  c.t = t; c.dead = dead; c.corr1 = corr1; c.corr2 = corr2;
#pragma unroll
  for (int e = 0; e < kNGA; ++e) c.gA[e] = gA[e];
#pragma unroll
  for (int e = 0; e < kNGB; ++e) c.gB[e] = gB[e];
  return ret;
}

// The primary threads (one per SNP: inputs, classes, verification) and the secondary threads (the corrections owed to
// the next tile, Gram words towards t+2 in registers) run the same sequence of barriers from two instantiations of these
// functions, so that neither carries the other's registers.
template <int NF, bool PRIM, int B>
__device__ __noinline__ void serial_threads() {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PrivLayout P = G::P;
  const SweepParams& p = g_serial_ps;
  const int tid = threadIdx.x;
  volatile int* gctl = (volatile int*)(smem + G::priv0 + P.o_gctl);
  long long* pc = (long long*)(smem + G::priv0 + P.o_pc);
  SerialCarry c;
  c.t = 0; c.dead = false; c.corr1 = 0.0; c.corr2 = 0.0;
  for (int e = 0; e < kNGA; ++e) c.gA[e] = make_int4(0, 0, 0, 0);
  for (int e = 0; e < kNGB; ++e) c.gB[e] = make_int4(0, 0, 0, 0);
  for (;;) {
    SerialTile st;
    if (serial_fast<NF, PRIM, B>(c, st) == 0) break;
    const int t = c.t;
    if (st.generic && tid == 0) gctl[13]++;
    serial_slow<NF, PRIM, B>(st, t, st.generic);
    c.corr1 = st.c1next;
    serial_tile_end<PRIM, B>(t, st.k, st.nrounds, st.cand, st.act, st.myrank, st.cls, st.my_delta, st.my_gnew, c.corr2, c.gA, c.gB,
                             PRIM && tid == 0 && (p.dbg & 64));
    c.t = t + 1;
  }
  if (c.dead) atomicCAS(p.ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
  if (tid == 0) {
    for (int k = 0; k < 16; ++k) p.out->phase_clk[0][k] = pc[k];
    p.out->phase_clk[1][0] = gctl[12];
    p.out->phase_clk[1][1] = gctl[13];
    atomicAdd(&p.out->rounds, gctl[10]);
    atomicAdd(&p.out->pad, gctl[12]);
    atomicAdd(&p.out->n_changed, gctl[11]);
  }
}

// The input warp of the serial CTA (its last warp): per tile, in order, (1) the package -> buffer t & 1 with one bulk
// copy as soon as the buffer is free and the helper's flag is up, (2) the tile's far corrections (dt >= 3, one slot of B
// doubles per source tile, contiguous in global memory) -> landing zone with one bulk copy, summed oldest first into
// coldbuf[t & 1]; a slot that had not been posted yet when the copy ran is polled word by word.
template <int B>
__device__ __noinline__ void serial_input_warp() {
  extern __shared__ __align__(128) uint8_t smem[];
  using G = SerialGeo<B>;
  constexpr PkgLayout L = G::L;
  constexpr PrivLayout P = G::P;
  const SweepParams& p = g_serial_ps;
  const int T = p.T, DC = p.D - 1;
  const int lane = threadIdx.x & 31;
  int* ctrl = p.ctrl;
  uint8_t* priv = smem + G::priv0;
  double* coldbuf = (double*)(priv + P.o_coldbuf);
  const unsigned long long* stage = (const unsigned long long*)(priv + P.o_stage);
  uint64_t* full = (uint64_t*)(priv + P.o_full);
  uint64_t* empty = (uint64_t*)(priv + P.o_empty);
  uint64_t* farfull = (uint64_t*)(priv + P.o_farfull);
  uint64_t* stagefull = (uint64_t*)(priv + P.o_stagefull);
  // lane 0 keeps the flag of the next tile: read (acquire) and fenced towards the async proxy while this tile's far
  // corrections are on their way, so that neither trip is on the path of the next package
  int fnext = 0;
  if (lane == 0) {
    fnext = ld_acquire_s32(p.pkg_flag);
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
  for (int t = 0; t < T; ++t) {
    const int b = t & 1;
    if (t >= 2 && !mbar_wait(empty + b, (uint32_t)(((t >> 1) - 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
    const int dmax = min(DC, t);
    const int nfar = max(0, dmax - 2);   // slots q = 2 .. dmax-1
    int f = 0;
    if (lane == 0) {
      f = fnext;
      if (f == 0) {
        Waiter w;
        do {
          __nanosleep(40);
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) break;
          f = ld_acquire_s32(p.pkg_flag + t);
        } while (f == 0);
        asm volatile("fence.proxy.async.global;" ::: "memory");   // the helper's (generic) stores before the bulk (async) reads
      }
      HB_TRACE(t, 11);
    }
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f == 0) return;   // abandoned
    const int nrow = (f & kPkgNoRows) ? 0 : ((f & 0xffff) - 1);
    const uint32_t bytes = L.fixed_bytes + (uint32_t)nrow * L.row_bytes;
    // the package in pieces of 8 KB, one lane each: a bulk copy keeps only so many requests in flight (a single 75 KB copy
    // took ~4 us here), several copies run side by side -- and so does their issue
    if (lane == 0) {
      hb::mbar_arrive_expect_tx(full + b, bytes);
      if (nfar > 0) hb::mbar_arrive_expect_tx(stagefull, (uint32_t)nfar * B * 8);
    }
    __syncwarp();
    for (uint32_t off = 8192u * lane; off < bytes; off += 8192u * 32)
      hb::tma_load_1d(smem + (size_t)b * L.stride + off, p.pkg + (size_t)t * p.pkg_stride + off, min(8192u, bytes - off), full + b);
    if (lane == 31 && nfar > 0) hb::tma_load_1d((void*)stage, p.corr + ((size_t)t * DC + 2) * B, (uint32_t)nfar * B * 8, stagefull);
    if (lane == 0) {
      HB_TRACE(t, 12);
      fnext = 0;
      if (t + 1 < T) {
        fnext = ld_acquire_s32(p.pkg_flag + t + 1);   // consumed in the next iteration
        if (fnext) asm volatile("fence.proxy.async.global;" ::: "memory");
      }
    }
    __syncwarp();
    if (nfar > 0) {
      if (!mbar_wait(stagefull, (uint32_t)((t - 3) & 1), ctrl, HB_ABORT_TIMEOUT_SCALAR)) return;
      if (lane == 0) HB_TRACE(t, 13);
      // far corrections of SNP s, summed oldest first; a slot that had not been posted yet when the copy ran is polled.
      // Four SNPs at a time, all their slot words asked for before the first is used (the loads are independent).
      for (int s0 = lane; s0 < B; s0 += 128) {
        unsigned long long w[4][kMaxDC - 2];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < kMaxDC - 2; ++q)
            w[r][q] = (q < nfar && s0 + 32 * r < B) ? stage[(size_t)q * B + s0 + 32 * r] : 0ull;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int s = s0 + 32 * r;
          if (s < B) {
            double cold = 0.0;
#pragma unroll
            for (int q = kMaxDC - 3; q >= 0; --q)
              if (q < nfar) {
                unsigned long long v = w[r][q];
                if (v == kCorrEmpty) {
                  unsigned spins = 0;
                  const unsigned long long* slot = (const unsigned long long*)(p.corr + ((size_t)t * DC + q + 2) * B + s);
                  do {
                    v = ld_relaxed_u64(slot);
                    if ((++spins & 0xffff) == 0 && *((volatile int*)(ctrl + 1)) != 0) return;
                  } while (v == kCorrEmpty);
                }
                cold += __longlong_as_double((long long)v);
              }
            coldbuf[b * B + s] = cold;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) { hb::mbar_arrive(farfull + b); HB_TRACE(t, 14); }   // (release at CTA scope: the sums are visible to the waiters)
  }
}

template <int NF>
__device__ void serial_role(const SweepParams& pin, uint8_t* smem) {
  constexpr int B = kSerialB;
  using G = SerialGeo<B>;
  for (int w = threadIdx.x; w < (int)(sizeof(SweepParams) / 4); w += blockDim.x) ((int*)&g_serial_ps)[w] = ((const int*)&pin)[w];
  __syncthreads();
  const SweepParams& p = g_serial_ps;
  const int tid = threadIdx.x;
  if (tid >= 2 * B) return;
  if (p.dbg & 48) return;   // timing experiments of the streaming side: the helpers publish empty tiles
  // threads: [0, B) primary (one per SNP), [B, 2B-32) secondary (odd candidates of SNP i < B-32), last warp = input warp
  constexpr int NTS = G::NTS;
  constexpr PrivLayout P = G::P;
  uint8_t* priv = smem + G::priv0;
  volatile int* gctl = (volatile int*)(priv + P.o_gctl);
  long long* pc = (long long*)(priv + P.o_pc);
  if (tid == 0) {
    uint64_t* full = (uint64_t*)(priv + P.o_full);
    uint64_t* empty = (uint64_t*)(priv + P.o_empty);
    uint64_t* farfull = (uint64_t*)(priv + P.o_farfull);
    for (int b = 0; b < 2; ++b) { hb::mbar_init(full + b, 1); hb::mbar_init(empty + b, 1); hb::mbar_init(farfull + b, 1); }
    hb::mbar_init((uint64_t*)(priv + P.o_stagefull), 1);
    hb::mbar_fence_init();
    for (int k = 0; k < 16; ++k) gctl[k] = 0;
    for (int k = 0; k < 16; ++k) pc[k] = 0;
    pc[16] = clock64();
  }
  hb::named_bar_sync(5, 2 * B);
  if (tid >= NTS) serial_input_warp<B>();
  else if (tid < B) serial_threads<NF, true, B>();
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
    // Slots are numbered with the candidates first, in SNP order (slot of candidate e = e: the serial CTA then walks the
    // rows of a tile's changes without looking the slots up), the other wanted rows after them.
    auto select_rows = [&](bool with_near) {
      if (prim) {
        const double rr_spec = rhs_spec * rhs_spec;
        const bool near = with_near && TH[0] > 0.0 && TH[0] < 1e300 && rr_spec >= 0.49 * TH[0];
        const bool isc = act && (gold != 0.0 || cls > 0);
        const bool extra = act && !isc && near;
        const unsigned bal = __ballot_sync(0xffffffffu, isc), bal2 = __ballot_sync(0xffffffffu, extra);
        if (lane == 0) { wcnt[warp] = __popc(bal); wcnt[32 + warp] = __popc(bal2); }
        hb::named_bar_sync(4, B);
        int pre = 0, tot = 0, pre2 = 0, tot2 = 0;
        for (int w = 0; w < nwarp; ++w) {
          const int c = wcnt[w], c2 = wcnt[32 + w];
          if (w < warp) { pre += c; pre2 += c2; }
          tot += c; tot2 += c2;
        }
        const int sl = isc ? pre + __popc(bal & ((1u << lane) - 1u)) : tot + pre2 + __popc(bal2 & ((1u << lane) - 1u));
        slot_of[i] = (isc || extra) ? sl : -1;
        if (isc || extra) slot_snp[sl] = i;
        if (i == 0) gctl[2] = tot + tot2;
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
      if (k > 0) {
        if (!prim && warp == 0) chain_build_matrix(coef, cmat, min(k, 32), lane);
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
        ((double*)(pk + L.o_base0))[i] = base0 + addback;
        ((float*)(pk + L.o_lo))[i] = __double2float_ru(lo);   // (inwards: a class is never accepted on a rounded bound)
        ((float*)(pk + L.o_hi))[i] = __double2float_rd(hi);
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
    // ================= phase C: the corrections this tile owes to the tiles t+3 .. t+D-1, once it is final.
    // While the tile waits for its turn in the chain, the far Gram rows of its speculated candidates (slot e = candidate e)
    // are brought into shared memory (the row buffers are free once the package is written): when the changes arrive the
    // corrections cost no trip to memory, and the first of them is posted ~1.5 us after the tile was published -- the
    // serial CTA's input warp reads it one tile later.
    {
      int* fidx = cs.idx;        // the final list (the candidate arrays are free now)
      double* fdel = cs.delta;
      int* fslot = cs.cls;       // row slot of a final change in the far buffers, -1: not there (read from the Gram band)
      const int nblk = max(0, min(D, T - t) - 3);          // far blocks dt = 3 .. 3 + nblk - 1
      const int kpre = fast ? min(k, KROW) : 0;            // candidates whose far rows are fetched ahead
      const int cap = (2 * KROW) / max(1, kpre);           // blocks that fit the two row buffers
      const int npre = min(nblk, cap);
      int32_t* far = rows0;                                // [npre][kpre][B]
      if (kpre > 0 && npre > 0) {
        const int q4 = B / 4;   // 16-byte pieces per row
        for (int e = tid; e < npre * kpre * q4; e += NT2) {
          const int c = e % q4, r = (e / q4) % kpre, d = e / (q4 * kpre);
          const int32_t* src = G0 + (size_t)(3 + d) * B * B + (size_t)slot_snp[r] * B + 4 * c;
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(far + ((size_t)d * kpre + r) * B + 4 * c);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        cp_async_wait_all();
      }
      // the tile's changes: the count and the first 32 entries are asked for in the same trip to L2
      if (tid < 32) {
        int cnt = -1, jl0 = -1;
        unsigned long long dw0 = kCorrEmpty;
        Waiter w;
        for (;;) {
          int c = -1;
          if (lane == 0) c = hb::ld_relaxed(p.tile_cnt + t);
          if (jl0 < 0) jl0 = hb::ld_relaxed(p.q_snp + (size_t)t * B + lane);
          if (dw0 == kCorrEmpty) dw0 = ld_relaxed_u64(p.q_delta + (size_t)t * B + lane);
          cnt = __shfl_sync(0xffffffffu, c, 0);
          if (cnt >= 0 && __all_sync(0xffffffffu, lane >= cnt || (jl0 >= 0 && dw0 != kCorrEmpty))) break;
          if (cnt < 0) __nanosleep(100);
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { cnt = -2; break; }
        }
        if (lane < cnt) { fidx[lane] = jl0 - t * B; fdel[lane] = __longlong_as_double((long long)dw0); }
        if (lane == 0) gctl[3] = cnt;
      }
      hb::named_bar_sync(1, NT2);
      const int kf = gctl[3];
      if (kf < 0) { dead = true; break; }
      if (tid == 0) HB_TRACE(t, 8);
      for (int e = 32 + tid; e < kf; e += NT2) {
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
      if (kf > 32 && hb::named_bar_or(1, NT2, dead)) { dead = true; break; }
      // row slot of every final change (slot_of: candidates are slots 0 .. k-1)
      bool all_pre = true;
      for (int e = tid; e < kf; e += NT2) {
        const int sl = slot_of[fidx[e]];
        fslot[e] = (sl >= 0 && sl < kpre) ? sl : -1;
      }
      hb::named_bar_sync(1, NT2);
      for (int e = 0; e < kf; ++e) all_pre = all_pre && (fslot[e] >= 0);
      for (int dt = 3 + h; dt < D; dt += 2) {
        if (t + dt >= T) break;
        double cv;
        if (dt - 3 < npre && all_pre) {
          // (ascending order of the changes, one fma each: the sum band_correction_raw forms)
          const int32_t* fr = far + (size_t)(dt - 3) * kpre * B + i;
          cv = 0.0;
          for (int e = 0; e < kf; ++e) cv = fma(gram_as_double(fr[(size_t)fslot[e] * B]), fdel[e], cv);
        } else {
          cv = band_correction_raw(fidx, fdel, kf, G0 + (size_t)dt * B * B, B, i);
        }
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
