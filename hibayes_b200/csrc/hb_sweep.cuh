// hb_sweep.cuh -- the fused Gibbs sweep kernel (sm_100a).
//
// One cooperative launch per MCMC iteration replaces the switch(model_index) block of Bayes()
// (/root/reference/src/Bayes.cpp:586-816): grid = S streaming CTAs (one row slab each, one per SM)
// + NG scalar CTAs.  X is read from HBM exactly once per sweep.
//
// Streaming CTA (warp-specialised):
//   TMA warp      cp.async.bulk ring: sub-stages of B/4 SNP columns x R rows (raw int8)
//   compute warps x_j'r over the slab: a lane keeps RL residual rows in registers for a whole tile and one
//                 running dot per column (PRMT + DFMA per genotype); a transposed shuffle reduction over
//                 the 16 lanes of a half-warp completes the slab dots, which are added as fixed-point
//                 integers (with an arrival count in the low byte) to the per-SNP accumulators in L2:
//                 order-independent, hence deterministic, and complete exactly when the count says so
//   AXPY warps    own the master copy of the slab's residual/u rows in registers; D tiles later they apply
//                 the effect changes the scalar CTAs published (r -= x*delta, u += x*delta,
//                 Bayes.cpp:787-789, in SNP order) and republish the slab for the compute warps
//
// Scalar CTAs (one worker each, tiles round-robin, one thread pair per SNP of a tile of B SNPs):
//   turn the reduced dots of tile t into the conditional draws of Bayes.cpp:756-801.  The chain x_j'r
//   depends on every earlier change; inside a tile that dependence is the exact integer Gram block
//   G = X_t'X_t, across the D-1 tiles still in flight it is the Gram band.  Classes are speculated for the
//   whole tile, the changed SNPs ("candidates") are chained by one warp as a small triangular recurrence
//   in SNP order, every SNP is then re-evaluated with its exact right-hand side and the first SNP whose
//   class differs from the speculation restarts the round.  Random draws are position-addressed
//   (hb_rng.h), so re-evaluation reuses the same uniform/normal and the result equals the
//   one-SNP-at-a-time sweep.  Workers hand over through global memory; every handed-over value is its own
//   flag (a NaN payload / -1 means "not yet written"), so no hand-over needs a fence or a second round trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hibayes_b200.h"
#include "hb_device.cuh"
#include "hb_limbs.h"

struct SweepOutDev {
  double count[HB_MAX_FOLD];
  double varg_acc;
  double sum_vargL;
  double sum_r, sum_r2, sum_u, var_u;
  int n_changed;
  int status;
  int rounds;        // speculation rounds summed over tiles (diagnostic)
  int pad;
  long long phase_clk[2][16];  // scalar workers 0 and 1: SM cycles per phase (diagnostic)
};

struct SweepParams {
  const uint8_t* Xp;
  double *r, *u;
  const double* xpx;
  const uint8_t* active;
  double* g;
  int32_t* tracker;
  const int32_t* gram;
  unsigned long long* dacc;   // per SNP: sum over slabs of (fixed-point dot << 8) + 1
  unsigned long long* peer_acc[8];   // row-sharded runs: every rank's second-level accumulators (this sweep's buffer):
                                     // sum over ranks of (local fixed-point dot << 4) + 1; [rank] is the local one
  int world, rank;
  int* q_snp;                 // [T][B] changed SNPs of tile t (global index), -1 = not yet written
  double* q_delta;            // [T][B] their effect changes, kCorrEmpty = not yet written
  int* tile_cnt;              // per tile: number of changes, -1 until known
  double* corr;               // [T][D-1][B] corrections owed to tile t by tile t-dt, kCorrEmpty = not yet written
  int* ctrl;                  // [1] abort code
  unsigned long long* trace;  // diagnostics (HB_TRACE): [T][16] globaltimer stamps of a tile's events, or null
  const double* prm;
  SweepOutDev* out;
  size_t slab_stride, m_pad;
  int n, m, S, R, T, B, D, NS;
  int NCW, NAW, SUBB, NG, KROW;
  uint32_t stage_bytes, off_rbuf, off_bar, off_qbuf;
  int model, F;
  double fold[HB_MAX_FOLD];
  double logpi0;
  double dscale, inv_dscale, mu_shift;
  int use_thr;   // class decisions of the mixture models by certified thresholds on rhs^2 (k_prep), no exp in the chain
  int dbg;       // timing experiments only (HB_DEBUG env): 1 skip AXPY, 2 skip dot FMAs, 16/32 streaming side alone
  int scalar0;   // block index of the first scalar CTA (>= S; blocks S .. scalar0-1 are idle padding)
  int cluster2;  // launched as clusters of 2 CTAs: worker pairs (0,1), (2,3), ... hand over through distributed shared memory
  // LIMBS variant only (HB_LIMBS=1, experimental): the residual slab is handed to the compute warps as 48-bit fixed
  // point, q = rint(r * rscale), in six 8-bit limbs (hb_limbs.h); rscale = dscale * 2^rshift
  double rscale;
  int rshift;
  // serial mode (hb_serial.cuh): one serial CTA (block scalar0) runs phase S of every tile, NH helper CTAs (blocks
  // scalar0+1 ..) prepare the tiles (phase P -> a package in global memory) and post the far corrections (phase C)
  int serial;                 // 1: serial CTA + helpers; 0: ring of NG workers (scalar_role)
  int NH;                     // helper CTAs
  int KROW_S;                 // row slots (and candidates) of a package
  uint32_t pkg_stride;        // bytes between the packages of consecutive tiles
  uint8_t* pkg;               // [T] packages
  int* pkg_flag;              // [T] 0 = not written yet, else 1 + number of row slots (| 1 << 20: no rows, generic path)
  int* miss_tile;             // [1] last tile that needed a second round (the helpers widen their row sets after it)
  int xevict;                 // genotype tiles are streamed with an L2 evict_first hint (HB_XEVICT)
  double near_frac;           // rows are also gathered for SNPs with rhs^2 >= near_frac * (first class boundary) while speculation misses
};

enum { HB_ABORT_TIMEOUT_STREAM = 1, HB_ABORT_TIMEOUT_SCALAR = 2, HB_ABORT_TIMEOUT_TMA = 3, HB_ABORT_OVERFLOW = 4,
       HB_ABORT_TIMEOUT_PIPE = 5 };

// prm layout (SoA over m_pad): [0] u, [1] z, then for k = 1..F-1: a_k, c_k, 1/v_k, sd_k*z; then for every class
// boundary b = 0..F-2 the certified thresholds TL_b, TH_b on rhs^2 (solve_thresholds)
__host__ __device__ __forceinline__ size_t prm_idx(int field, size_t m_pad, int j) { return (size_t)field * m_pad + j; }
constexpr int kThrField0 = 2 + 4 * (HB_MAX_FOLD - 1);
constexpr int kPrmFields = kThrField0 + 2 * (HB_MAX_FOLD - 1);

namespace hbk {

constexpr double kTwo513 = 2.6815615859885194e154;     // 2^513
constexpr double kTwoM513 = 3.7291703656001034e-155;   // 2^-513
constexpr long long kTimeoutNs = 4000000000ll;          // every wait is bounded: a lost signal aborts, it never hangs
constexpr double kFixLimit = 1.8e16;                    // |fixed-point partial dot| < 2^54: 8 bits left for the count
// A slot that has not been written yet holds this NaN payload (k_prep fills the arrays).
constexpr unsigned long long kCorrEmpty = 0x7ff8dead0badf00dull;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// distributed shared memory of a 2-CTA cluster (hand-over between the two scalar workers of a cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_peer_shared_u64(uint32_t local_addr, uint32_t peer_rank, unsigned long long v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(peer_rank));
  asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_own_shared_cluster_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_own_shared_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.relaxed.cluster.shared::cta.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

struct Waiter {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  // returns false when the wait must be abandoned (abort flag raised somewhere, or timeout)
  __device__ __noinline__ bool keep_waiting(int* ctrl, int code) {
    if ((++spins & 0x3f) == 0) {
      __nanosleep(32);
      if (*((volatile int*)(ctrl + 1)) != 0) return false;
      const unsigned long long now = gtimer();
      if (t0 == 0) t0 = now;
      else if ((long long)(now - t0) > kTimeoutNs) {
        atomicCAS(ctrl + 1, 0, code);
        return false;
      }
    }
    return true;
  }
};

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* ctrl, int code) {
  if (hb::mbar_try_wait(bar, parity)) return true;
  Waiter w;
  while (!hb::mbar_try_wait(bar, parity))
    if (!w.keep_waiting(ctrl, code)) return false;
  return true;
}

#define HB_TRACE(t, ev) do { if (p.trace) p.trace[(size_t)(t) * 16 + (ev)] = gtimer(); } while (0)

__device__ __forceinline__ double byte_as_scaled(uint32_t w, uint32_t sel) {
  // genotype byte -> mantissa bits 48..55 of a double: value = byte * 2^-1026 (exact, denormal)
  return __hiloint2double((int)__byte_perm(w, 0u, sel), 0);
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void st_relaxed_u64(void* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_s32(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread-private hand-over: spin on one 8-byte word until its producer has stored a value
__device__ __noinline__ bool poll_corr_slow(const double* slot, double& v, int* ctrl) {
  unsigned long long w;
  Waiter wt;
  do {
    if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) return false;
    w = ld_relaxed_u64(slot);
  } while (w == kCorrEmpty);
  v = __longlong_as_double((long long)w);
  return true;
}
// The wait on the serial path (corrections of the previous tile): one load at a time would look at the slot once per
// L2 round trip (~1 us under the streaming load).  Four loads are kept in flight, issued a quarter of a round trip
// apart, so that the value is seen at most ~0.25 us after it could be (eight in flight measured no better).
__device__ __noinline__ bool poll_corr_staggered(const double* slot, double& v, int* ctrl) {
  unsigned long long w[4];
  const long long t0 = clock64();
  w[0] = ld_relaxed_u64(slot);
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    while (clock64() - t0 < 450ll * q) {}
    w[q] = ld_relaxed_u64(slot);
  }
  Waiter wt;
  for (;;) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (w[q] != kCorrEmpty) {
        v = __longlong_as_double((long long)w[q]);
        return true;
      }
      w[q] = ld_relaxed_u64(slot);
    }
    if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) return false;
  }
}
__device__ __forceinline__ bool poll_corr(const double* slot, double& v, int* ctrl) {
  const unsigned long long w = ld_relaxed_u64(slot);
  if (w == kCorrEmpty) return poll_corr_slow(slot, v, ctrl);
  v = __longlong_as_double((long long)w);
  return true;
}
__device__ __forceinline__ void post_corr(double* slot, double v) { st_relaxed_u64(slot, (unsigned long long)__double_as_longlong(v)); }

// ------------------------------------------------------------------------------------------
// streaming CTA
// ------------------------------------------------------------------------------------------
// Layout of one CTA (slab of R = 16*RL rows, tile of B = 32*NCW SNP columns, sub-stage = B/4 columns):
//   compute warp w, half-warp h, lane l: rows [RL*l, RL*l + RL) of the slab; in sub-stage q it takes the four
//   columns 8w + 4h + {0..3} and keeps one running dot per column (16 per tile).  After the tile's four
//   sub-stages the 16 lanes of a half-warp hold 16 x 16 partial dots; a transposed shuffle reduction (8+4+2+1
//   exchanges, fixed tree) leaves lane l with the complete slab dot of accumulator l.
// LIMBS variant of the compute warps (experimental, HB_LIMBS=1; written at the end of round 1 and NOT yet run on
// hardware): integer-only dots.  The AXPY warps publish the residual slab as six 8-bit limbs of q = rint(r * rscale)
// (hb_limbs.h; limb-major, lbuf[buffer][limb][row]); a lane keeps the limb words of its RL rows in registers for a whole
// tile; four genotypes of a column -- one 32-bit word as it lies in shared memory, no PRMT -- cost six dp4a, the six
// int32 sums of a lane's rows merge into one int64, and the half-warp reduction is exact.  The only rounding is the
// quantisation of r (<= sum(x) / (2 rscale) per slab dot) and the shift to the accumulators' scale.
template <int RL>
__device__ void stream_compute_limbs(const SweepParams& p, uint8_t* smem, uint64_t* full, uint64_t* empty, uint64_t* rfull,
                                     uint64_t* rempty) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = p.NS, B = p.B, T = p.T, R = p.R, SUBB = p.SUBB;
  constexpr int Q = 4;
  constexpr int NW = RL / 4;   // limb words per lane and limb
  uint8_t* stage0 = smem;
  const uint8_t* lbuf = smem + p.off_rbuf;   // 2 x HB_NLIMB x R bytes (inside the 2 x R doubles of the fp64 variant)
  int* ctrl = p.ctrl;
  const int h = lane >> 4, l = lane & 15;
  const uint32_t lane_off = (uint32_t)((8 * warp + 4 * h) * R + RL * l);
  const int my_col = (l >> 2) * SUBB + 8 * warp + 4 * h + (l & 3);
  int st = 0;
  uint32_t st_par = 0;
  for (int t = 0; t < T; ++t) {
    if (!mbar_wait(rfull + (t & 1), (uint32_t)((t >> 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
    uint32_t lw[HB_NLIMB][NW];
    {
      const uint8_t* lb = lbuf + (size_t)(t & 1) * HB_NLIMB * R + RL * l;
#pragma unroll
      for (int k = 0; k < HB_NLIMB; ++k)
#pragma unroll
        for (int w = 0; w < NW; ++w) lw[k][w] = *(const uint32_t*)(lb + (size_t)k * R + 4 * w);
    }
    __syncwarp();
    if (lane == 0) hb::mbar_arrive(rempty + (t & 1));
    long long acc[16];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      if (!mbar_wait(full + st, st_par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
      const uint8_t* sp = stage0 + (size_t)st * p.stage_bytes + lane_off;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {   // two columns at a time: twelve independent dp4a chains
        int a0[HB_NLIMB], a1[HB_NLIMB];
#pragma unroll
        for (int k = 0; k < HB_NLIMB; ++k) { a0[k] = 0; a1[k] = 0; }
        const uint8_t* c0p = sp + (size_t)(2 * c2) * R;
        const uint8_t* c1p = c0p + R;
#pragma unroll
        for (int wd = 0; wd < RL / 8; ++wd) {
          const uint2 v0 = *(const uint2*)(c0p + 8 * wd), v1 = *(const uint2*)(c1p + 8 * wd);
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < HB_NLIMB; ++k) {
              a0[k] = hb_limb_dp4a(v0.x, lw[k][2 * wd], k, a0[k]);
              a1[k] = hb_limb_dp4a(v1.x, lw[k][2 * wd], k, a1[k]);
              a0[k] = hb_limb_dp4a(v0.y, lw[k][2 * wd + 1], k, a0[k]);
              a1[k] = hb_limb_dp4a(v1.y, lw[k][2 * wd + 1], k, a1[k]);
            }
          }
        }
        acc[4 * q + 2 * c2] = hb_limb_merge(a0);
        acc[4 * q + 2 * c2 + 1] = hb_limb_merge(a1);
      }
      __syncwarp();
      if (lane == 0) hb::mbar_arrive(empty + st);
      if (++st == NS) { st = 0; st_par ^= 1u; }
    }
    // transposed reduction over the 16 lanes of the half-warp (the tree of the fp64 variant; integer sums are exact)
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
      const bool up = (l & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const long long keep = up ? acc[i + o] : acc[i];
        const long long send = up ? acc[i] : acc[i + o];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    {
      // x'q is in units of 1/rscale; the accumulators count units of 1/dscale, rscale = dscale * 2^rshift
      long long fx = acc[0];
      if (p.rshift > 0) fx = (fx + (1ll << (p.rshift - 1))) >> p.rshift;
      else if (p.rshift < 0) fx = fx * (1ll << (-p.rshift));
      if (!(fx > -(long long)kFixLimit && fx < (long long)kFixLimit)) atomicCAS(ctrl + 1, 0, HB_ABORT_OVERFLOW);
      atomicAdd(p.dacc + (size_t)t * B + my_col, ((unsigned long long)fx << 8) + 1ull);
      if (blockIdx.x == 0 && tid == 0) HB_TRACE(t, 7);
    }
  }
}

template <int RL, bool LIMBS = false>
__device__ void stream_role(const SweepParams& p, uint8_t* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x;
  const int NS = p.NS, B = p.B, D = p.D, T = p.T, R = p.R, NCW = p.NCW, NAW = p.NAW;
  const int SUBB = p.SUBB;
  constexpr int Q = 4;
  uint8_t* stage0 = smem;
  double* rbuf = (double*)(smem + p.off_rbuf);   // 2 x R   (residual slab * 2^513)
  uint64_t* full = (uint64_t*)(smem + p.off_bar);
  uint64_t* empty = full + NS;
  uint64_t* rfull = empty + NS;
  uint64_t* rempty = rfull + 2;
  double* qd = (double*)(smem + p.off_qbuf);     // the current tile's residual updates: B deltas, B SNPs, count
  int* qj = (int*)(qd + B);
  int* qcnt = qj + B;
  int* ctrl = p.ctrl;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { hb::mbar_init(full + i, 1); hb::mbar_init(empty + i, NCW); }
    for (int i = 0; i < 2; ++i) { hb::mbar_init(rfull + i, NAW); hb::mbar_init(rempty + i, NCW); }
    hb::mbar_fence_init();
  }
  __syncthreads();
  const uint8_t* Xs = p.Xp + (size_t)s * p.slab_stride;
  const int nsub = T * Q;

  if (warp < NCW) {
    // ---------------- compute warps
    if constexpr (LIMBS) {
      stream_compute_limbs<RL>(p, smem, full, empty, rfull, rempty);
      return;
    }
    const int h = lane >> 4, l = lane & 15;
    const uint32_t lane_off = (uint32_t)((8 * warp + 4 * h) * R + RL * l);
    // accumulator l = 4*q + k  <->  column q*SUBB + 8w + 4h + k of the tile
    const int my_col = (l >> 2) * SUBB + 8 * warp + 4 * h + (l & 3);
    const double to_fix = p.dscale;
    double rs[RL];
    int st = 0;
    uint32_t st_par = 0;
    for (int t = 0; t < T; ++t) {
      if (!mbar_wait(rfull + (t & 1), (uint32_t)((t >> 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
      {
        const double2* rb = (const double2*)(rbuf + (size_t)(t & 1) * R + RL * l);
#pragma unroll
        for (int i = 0; i < RL / 2; ++i) { const double2 v = rb[i]; rs[2 * i] = v.x; rs[2 * i + 1] = v.y; }
      }
      __syncwarp();
      if (lane == 0) hb::mbar_arrive(rempty + (t & 1));
      double acc[16];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        if (!mbar_wait(full + st, st_par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        {
          const uint8_t* sp = stage0 + (size_t)st * p.stage_bytes + lane_off;
          uint2 v0 = *(const uint2*)(sp), v1 = *(const uint2*)(sp + R), v2 = *(const uint2*)(sp + 2 * R),
                v3 = *(const uint2*)(sp + 3 * R);
#pragma unroll
          for (int wd = 0; wd < RL / 8; ++wd) {
            uint2 n0 = v0, n1 = v1, n2 = v2, n3 = v3;
            if (wd + 1 < RL / 8) {
              n0 = *(const uint2*)(sp + 8 * (wd + 1)); n1 = *(const uint2*)(sp + R + 8 * (wd + 1));
              n2 = *(const uint2*)(sp + 2 * R + 8 * (wd + 1)); n3 = *(const uint2*)(sp + 3 * R + 8 * (wd + 1));
            }
            if (!(p.dbg & 2)) {
#define HB_ROW4(i, W, SEL)                                               \
  a0 = fma(rs[8 * wd + (i)], byte_as_scaled(v0.W, SEL), a0);            \
  a1 = fma(rs[8 * wd + (i)], byte_as_scaled(v1.W, SEL), a1);            \
  a2 = fma(rs[8 * wd + (i)], byte_as_scaled(v2.W, SEL), a2);            \
  a3 = fma(rs[8 * wd + (i)], byte_as_scaled(v3.W, SEL), a3);
              HB_ROW4(0, x, 0x4044) HB_ROW4(1, x, 0x4144) HB_ROW4(2, x, 0x4244) HB_ROW4(3, x, 0x4344)
              HB_ROW4(4, y, 0x4044) HB_ROW4(5, y, 0x4144) HB_ROW4(6, y, 0x4244) HB_ROW4(7, y, 0x4344)
#undef HB_ROW4
            }
            v0 = n0; v1 = n1; v2 = n2; v3 = n3;
          }
        }
        __syncwarp();
        if (lane == 0) hb::mbar_arrive(empty + st);
        if (++st == NS) { st = 0; st_par ^= 1u; }
        acc[4 * q + 0] = a0; acc[4 * q + 1] = a1; acc[4 * q + 2] = a2; acc[4 * q + 3] = a3;
      }
      // transposed reduction over the 16 lanes of the half-warp: afterwards lane l holds accumulator l
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) {
        const bool up = (l & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
          const double keep = up ? acc[i + o] : acc[i];
          const double send = up ? acc[i] : acc[i + o];
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      {
        const double scaled = (acc[0] * kTwo513) * to_fix;
        if (!(fabs(scaled) < kFixLimit)) atomicCAS(ctrl + 1, 0, HB_ABORT_OVERFLOW);
        const unsigned long long fx = (unsigned long long)__double2ll_rn(scaled);
        atomicAdd(p.dacc + (size_t)t * B + my_col, (fx << 8) + 1ull);   // fire and forget: the count validates the sum
        if (s == 0 && tid == 0) HB_TRACE(t, 7);
      }
    }
    return;
  }
  if (warp == NCW) {
    // ---------------- TMA producer: streams this slab's sub-stages through the ring
    if (lane == 0) {
      int st = 0;
      uint32_t par = 1;   // parity of the previous phase of empty[st]
      const uint64_t pol = hb::l2_policy_evict_first();
      for (int gsub = 0; gsub < nsub; ++gsub) {
        if (gsub >= NS && !mbar_wait(empty + st, par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
        hb::mbar_arrive_expect_tx(full + st, p.stage_bytes);
        if (p.xevict) hb::tma_load_1d_hint(stage0 + (size_t)st * p.stage_bytes, Xs + (size_t)gsub * p.stage_bytes, p.stage_bytes, full + st, pol);
        else hb::tma_load_1d(stage0 + (size_t)st * p.stage_bytes, Xs + (size_t)gsub * p.stage_bytes, p.stage_bytes, full + st);
        if (++st == NS) { st = 0; par ^= 1u; }
      }
    }
    return;
  }
  if (warp < NCW + 1 + NAW) {
    // ---------------- AXPY warps: master copy of the slab's residual and u rows (4 rows / thread,
    // held as value * 2^-513 so that the denormal genotype factors stay exact)
    const int a = tid - 32 * (NCW + 1);
    const int row0 = 4 * a;
    const bool has = row0 < R;
    double rm[4], um[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t grow = (size_t)s * R + row0 + i;
      double rv = 0.0, uv = 0.0;
      if (has && grow < (size_t)p.n) { rv = p.r[grow] + p.mu_shift; uv = p.u[grow]; }
      rm[i] = rv * kTwoM513;
      um[i] = uv * kTwoM513;
    }
    for (int t = 0; t < T + D; ++t) {
      if (t >= D) {
        // residual updates published by the scalar workers for tile t-D.  The first AXPY warp fetches them into
        // shared memory (one polite poller per CTA); an entry is valid once it differs from its "not yet
        // written" pattern, so no fence is needed on either side.
        const int tt = t - D;
        const size_t qb = (size_t)tt * B;
        if (a < 32) {
          // the count and the first 32 entries are asked for in the same trip to L2 (entries beyond the count stay
          // "not yet written" and are ignored)
          int cnt = -1, jl0 = -1;
          unsigned long long dw0 = kCorrEmpty;
          {
            Waiter w;
            for (;;) {
              int c = -1;
              if (lane == 0) c = hb::ld_relaxed(p.tile_cnt + tt);
              if (jl0 < 0) jl0 = hb::ld_relaxed(p.q_snp + qb + lane);
              if (dw0 == kCorrEmpty) dw0 = ld_relaxed_u64(p.q_delta + qb + lane);
              cnt = __shfl_sync(0xffffffffu, c, 0);
              if (cnt >= 0 && __all_sync(0xffffffffu, lane >= cnt || (jl0 >= 0 && dw0 != kCorrEmpty))) break;
              if (cnt < 0) __nanosleep(100);
              if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_STREAM)) { cnt = -2; break; }
            }
          }
          if (p.dbg & 1) cnt = min(cnt, 0);
          if (lane < cnt) {
            qj[lane] = jl0;
            qd[lane] = __longlong_as_double((long long)dw0);
          }
          for (int q0 = 32; q0 < cnt; q0 += 32) {
            int jl = -1;
            unsigned long long dw = kCorrEmpty;
            if (q0 + lane < cnt) {
              Waiter w;
              for (;;) {
                if (jl < 0) jl = hb::ld_relaxed(p.q_snp + qb + q0 + lane);
                if (dw == kCorrEmpty) dw = ld_relaxed_u64(p.q_delta + qb + q0 + lane);
                if (jl >= 0 && dw != kCorrEmpty) break;
                if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_STREAM)) { cnt = -2; break; }
              }
              qj[q0 + lane] = jl;
              qd[q0 + lane] = __longlong_as_double((long long)dw);
            }
          }
          cnt = __reduce_min_sync(0xffffffffu, cnt);
          if (lane == 0) *qcnt = cnt;
          if (s == 0 && lane == 0) HB_TRACE(tt, 6);
        }
        hb::named_bar_sync(2, 32 * NAW);
        const int cnt = *(volatile int*)qcnt;
        if (cnt < 0) return;
        // the genotype words of sixteen changed SNPs in flight at a time (from L2: streamed D tiles ago)
        for (int e0 = 0; e0 < cnt; e0 += 16) {
          uint32_t xw[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const bool v = e0 + e < cnt;
            const int js = v ? qj[e0 + e] : 0;
            xw[e] = (v && has) ? __ldg((const uint32_t*)(Xs + (size_t)js * R + row0)) : 0u;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e0 + e < cnt) {
              const double ds = qd[e0 + e] * kTwo513;
              const double x0 = byte_as_scaled(xw[e], 0x4044), x1 = byte_as_scaled(xw[e], 0x4144);
              const double x2 = byte_as_scaled(xw[e], 0x4244), x3 = byte_as_scaled(xw[e], 0x4344);
              rm[0] = fma(-x0, ds, rm[0]); um[0] = fma(x0, ds, um[0]);   // yadj -= x*delta (Bayes.cpp:787), u += x*delta (:789)
              rm[1] = fma(-x1, ds, rm[1]); um[1] = fma(x1, ds, um[1]);
              rm[2] = fma(-x2, ds, rm[2]); um[2] = fma(x2, ds, um[2]);
              rm[3] = fma(-x3, ds, rm[3]); um[3] = fma(x3, ds, um[3]);
            }
        }
        hb::named_bar_sync(3, 32 * NAW);   // everybody has read the entries before they are overwritten
      }
      if (t < T) {
        if (t >= 2 && !mbar_wait(rempty + (t & 1), (uint32_t)(((t >> 1) - 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
        if constexpr (LIMBS) {
          if (has) {
            // the slab as fixed-point limbs: limb k of this thread's four rows is one 32-bit word
            uint32_t w[HB_NLIMB];
#pragma unroll
            for (int k = 0; k < HB_NLIMB; ++k) w[k] = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint8_t l6[HB_NLIMB];
              if (!hb_limb_split(rm[i] * kTwo513, p.rscale, l6)) {
                atomicCAS(ctrl + 1, 0, HB_ABORT_OVERFLOW);
#pragma unroll
                for (int k = 0; k < HB_NLIMB; ++k) l6[k] = 0;
              }
#pragma unroll
              for (int k = 0; k < HB_NLIMB; ++k) w[k] |= (uint32_t)l6[k] << (8 * i);
            }
            uint8_t* lb = (uint8_t*)rbuf + (size_t)(t & 1) * HB_NLIMB * R + row0;
#pragma unroll
            for (int k = 0; k < HB_NLIMB; ++k) *(uint32_t*)(lb + (size_t)k * R) = w[k];
          }
        } else if (has) {
          double* rb = rbuf + (size_t)(t & 1) * R + row0;
#pragma unroll
          for (int i = 0; i < 4; ++i) rb[i] = (rm[i] * kTwo513) * kTwo513;
        }
        __syncwarp();
        if (lane == 0) hb::mbar_arrive(rfull + (t & 1));
      }
    }
    if (has) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t grow = (size_t)s * R + row0 + i;
        p.r[grow] = rm[i] * kTwo513;
        p.u[grow] = um[i] * kTwo513;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// class decisions
// ------------------------------------------------------------------------------------------
// Cumulative class probabilities of one SNP given rr = rhs^2 (Bayes.cpp:759-770 in soft-max form).
template <int NF>
__host__ __device__ __forceinline__ void class_cum(int nf, double rr, const double* a, const double* c, double logpi0, double* cum) {
  double sv[NF];
  sv[0] = logpi0;
  double smax = logpi0;
#pragma unroll
  for (int k = 1; k < NF; ++k)
    if (k < nf) {
      sv[k] = fma(rr, c[k - 1], a[k - 1]);
      smax = fmax(smax, sv[k]);
    }
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < NF; ++k)
    if (k < nf) {
      sv[k] = exp(sv[k] - smax);
      tot += sv[k];
    }
  const double inv = 1.0 / tot;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < NF; ++k)
    if (k < nf) {
      acc = fma(sv[k], inv, acc);
      cum[k] = acc;
    }
}
// inverse-CDF class draw with one uniform (Bayes.cpp:773-781): first k with u < cum_k, 0 if none
template <int NF>
__host__ __device__ __forceinline__ int class_from_cum(int nf, double u, const double* cum) {
  int cls = 0;
  bool found = false;
#pragma unroll
  for (int k = 0; k < NF; ++k)
    if (k < nf && !found && u < cum[k]) { cls = k; found = true; }
  return cls;
}

// Class as a step function of rr = rhs^2.  With class-ordered slopes c_1 <= c_2 <= ... every cumulative
// probability cum_b(rr) decreases in rr, so "first b with u < cum_b" equals "first b with rr below the root
// theta_b of cum_b = u".  k_prep brackets each root as [TL_b, TH_b] and certifies both ends with class_cum
// itself; inside a bracket (or where certification failed: TL = -1, TH = inf) the caller evaluates exactly.
// Returns the class, or -1 when rr falls inside a bracket.
template <int NF>
__host__ __device__ __forceinline__ int thr_class(int nf, double rr, const double* TL, const double* TH) {
  // Level b is looked at only while all earlier levels were passed (rr >= TH): below TL[b] the class is b, between
  // TL[b] and TH[b] (or NaN) it is undecided.  Written without branches: the compares of all levels are independent.
  int c = 0;
  bool live = true, und = false;
#pragma unroll
  for (int b = 0; b < NF - 1; ++b)
    if (b < nf - 1) {
      const bool le = rr <= TL[b], ge = rr >= TH[b];
      und = und || (live && !le && !ge);
      live = live && ge && !le;
      c += live ? 1 : 0;
    }
  return und ? -1 : c;
}

// Brackets of the class boundaries of one SNP (device, k_prep).  phi_b(rr) = log S_hi - log S_lo - log((1-u)/u)
// with S_lo = sum_{l<=b} e^{s_l}, S_hi = sum_{l>b} e^{s_l}, s_0 = log pi_0, s_l = a_l + c_l rr, is increasing;
// safeguarded Newton finds its root; a small bracket around it is certified at both ends with a margin far
// above the rounding error of class_cum.
template <int NF>
__host__ __device__ inline void solve_thresholds(int nf, double u, const double* a, const double* c, double logpi0, double* TL, double* TH) {
  const double INF = HUGE_VAL;
  const double lam = log1p(-u) - log(u);
  const bool u_ok = (u < 1.0 - 1e-12) && (u > 1e-300);
  for (int b = 0; b < nf - 1; ++b) {
    TL[b] = -1.0;
    TH[b] = INF;
    if (!u_ok) continue;
    auto phi = [&](double rr, double& dphi) {
      double mlo = logpi0, mhi = -INF;
      for (int l = 1; l < nf; ++l) {
        const double sl = fma(rr, c[l - 1], a[l - 1]);
        if (l <= b) mlo = fmax(mlo, sl); else mhi = fmax(mhi, sl);
      }
      double Slo = exp(logpi0 - mlo), Clo = 0.0, Shi = 0.0, Chi = 0.0;
      for (int l = 1; l < nf; ++l) {
        const double sl = fma(rr, c[l - 1], a[l - 1]);
        if (l <= b) { const double e = exp(sl - mlo); Slo += e; Clo += c[l - 1] * e; }
        else { const double e = exp(sl - mhi); Shi += e; Chi += c[l - 1] * e; }
      }
      dphi = Chi / Shi - Clo / Slo;
      return (mhi + log(Shi)) - (mlo + log(Slo)) - lam;
    };
    double df, f = phi(0.0, df);
    double theta = 0.0;
    bool have = false;
    if (f >= 0.0) {
      // never at or below class b: certify at rr = 0 (cum_b decreases from there)
      double cum[NF];
      class_cum<NF>(nf, 0.0, a, c, logpi0, cum);
      if (cum[b] < u - 1e-13) { TL[b] = -1.0; TH[b] = -1.0; }
      continue;
    }
    {
      double lo = 0.0, hi = INF, rr = 0.0;
      for (int it = 0; it < 60; ++it) {
        if (!(df > 0.0)) break;
        double rn = rr - f / df;
        if (!(rn > lo) || !(rn < hi)) rn = (hi < INF) ? 0.5 * (lo + hi) : fmax(2.0 * rr, rr + 1.0);
        rr = rn;
        f = phi(rr, df);
        if (f < 0.0) lo = rr; else hi = rr;
        if (fabs(f) < 1e-12) { have = true; theta = rr; break; }
        if (hi < INF && (hi - lo) <= 1e-13 * hi) { have = true; theta = 0.5 * (lo + hi); f = phi(theta, df); break; }
      }
    }
    if (!have || !(df > 0.0)) continue;
    // bracket: 2e-8 in log-odds on either side of the root (the chance that rhs^2 falls inside is ~1e-8 per
    // decision); the ends are certified with an absolute margin of 1e-13 on the cumulative probability,
    // two orders above the rounding error of class_cum
    const double w = 1e-9 * theta + 2e-8 / df;
    const double tl = theta - w, th = theta + w;
    double cum[NF];
    class_cum<NF>(nf, th, a, c, logpi0, cum);
    if (!(cum[b] < u - 1e-13)) continue;
    if (tl >= 0.0) {
      class_cum<NF>(nf, tl, a, c, logpi0, cum);
      if (!(cum[b] > u + 1e-13)) continue;
      TL[b] = tl;
    }
    TH[b] = th;
  }
}

// ------------------------------------------------------------------------------------------
// scalar CTAs
// ------------------------------------------------------------------------------------------
// candidate arrays of a worker (B entries each)
struct CandSet {
  double *rhs0, *iv, *sdz, *gold, *delta, *gnew;
  int *idx, *cls, *slot;   // slot = row of the candidate in the row buffers
};

__device__ __forceinline__ CandSet make_candset(uint8_t* smem, int B) {
  CandSet cs;
  double* d = (double*)smem;
  cs.rhs0 = d; cs.iv = d + B; cs.sdz = d + 2 * B; cs.gold = d + 3 * B; cs.delta = d + 4 * B; cs.gnew = d + 5 * B;
  int* ip = (int*)(d + 8 * (size_t)B);
  cs.idx = ip; cs.cls = ip + B; cs.slot = ip + 2 * B;
  return cs;
}

// Shared memory of a scalar CTA: candidate arrays, two partial-sum arrays, and three buffers with the Gram rows of
// the tile's candidates (exact int32, as stored): rows0 = diagonal block, rows1 = block towards the next tile, rows2 =
// block towards the tile after that (its corrections are the next thing on the serial path after the hand-over: read
// from L2/HBM in phase C they arrived after the hand-over of the tile in between and paced every second tile).
// coefficients H (32x32) + solved chain matrix M (32x33, padded rows) + hand-over buffer of the cluster mode (2 x 256)
constexpr size_t kChainCoefBytes = 32 * 32 * 8 + 32 * 33 * 8 + 2 * 256 * 8;
__host__ __device__ inline size_t scalar_fixed_bytes(int B) {
  size_t b = (8 * (size_t)B) * 8 + (6 * (size_t)B + 64 + 16) * 4 + 18 * 8 + 64;   // candidates + partials, ints, timers
  b = (b + 127) / 128 * 128;
  return b + kChainCoefBytes;   // + the chain's coefficient matrix (last kChainCoefBytes of the fixed part)
}
__host__ inline int scalar_krow(int B) {
  const size_t cap = 226 * 1024;
  const size_t fixed = scalar_fixed_bytes(B);
  size_t rows = (cap - fixed) / (3 * (size_t)B * 4);
  if (rows > (size_t)B) rows = (size_t)B;
  return (int)rows;
}
__host__ inline size_t scalar_smem_bytes(int B) { return scalar_fixed_bytes(B) + 3 * (size_t)scalar_krow(B) * B * 4; }

// exact int32 -> double for 0 <= g < 2^31 on the full-rate pipe (one DADD instead of a quarter-rate I2F)
__device__ __forceinline__ double gram_as_double(int g) {
  return __hiloint2double(0x43300000, g) - 4503599627370496.0;
}

// Gathers the Gram rows of the k candidates from one band block into a row buffer as doubles: thread i takes
// column i of every row (coalesced), sixteen loads in flight.
__device__ __noinline__ void gather_rows(int32_t* dst, const int32_t* __restrict__ blk, const int* idx, int k, int B, int i) {
  // all rows in flight at once: asynchronous 4-byte copies global -> shared (a warp's copies of one row coalesce into
  // one 128-byte request), so the gather costs one trip to L2 instead of one per group of registers.  The Gram band
  // never changes, so the copy may go through L1.
  for (int sb = 0; sb < k; ++sb) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (size_t)sb * B + i);
    const int32_t* src = blk + (size_t)idx[sb] * B + i;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Candidate chain of the mixture models (one warp).  Lane s tracks
//   e_s = rhs_s/v_s + sd_s z_s - gold_s        (its effect change if nothing before it in the tile changed)
// and every earlier candidate s' lowers it by (G[c_s'][c_s]/v_s) delta_s'; when the chain reaches lane s its
// e_s is final (= delta_s).  One shuffle and one fma per candidate on the dependent path.
// ROWS: the candidates' Gram rows are in shared memory (doubles); otherwise they are read from global.
template <bool ROWS>
__device__ __forceinline__ void chain_candidates(const CandSet& cs, int k, const int32_t* __restrict__ G, const int32_t* rows, int B, int lane,
                                                 const double* coef = nullptr, int rs = 0) {
  if (rs == 0) rs = B;   // stride between the rows of the row buffer (the serial CTA interleaves two blocks: 2 B)
  for (int sb = 0; sb < k; sb += 32) {
    const int sidx = sb + lane;
    const bool valid = sidx < k;
    const int ci = valid ? cs.idx[sidx] : 0;
    const double iv = valid ? cs.iv[sidx] : 0.0, gold = valid ? cs.gold[sidx] : 0.0;
    const double niv = -iv;
    double e = valid ? fma(cs.rhs0[sidx], iv, cs.sdz[sidx]) - gold : 0.0;
    auto gval = [&](int sp) -> double {
      return gram_as_double(ROWS ? rows[(size_t)cs.slot[sp] * rs + ci] : __ldcg(G + (size_t)cs.idx[sp] * B + ci));
    };
    // candidates of earlier chunks: their changes are final
    for (int sp = 0; sp < sb; sp += 8) {
      double gv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) gv[q] = valid ? gval(sp + q) * niv : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) e = fma(gv[q], cs.delta[sp + q], e);
    }
    const int nl = min(32, k - sb);
    if (ROWS && sb == 0) {
      // first 32 candidates: the coefficients -G[c_lp][c_s]/v_s were laid out in shared memory when the candidate
      // list was built (phase P, off the serial path): coef[32 lp + s], zero for lp >= s.  Two halves of 16 steps,
      // one shuffle and one fma each; steps beyond the last candidate are skipped.  (A rolled loop with the next
      // coefficient fetched a step ahead was measured: 135 instead of 105 cycles per step.)
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        if (16 * hc < nl) {
          double hreg[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) hreg[q] = coef[(16 * hc + q) * 32 + lane];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            if (16 * hc + q >= nl) break;
            const double d = __shfl_sync(0xffffffffu, e, 16 * hc + q);
            e = fma(hreg[q], d, e);   // hreg = 0 for the lanes at or before the step: their e is final
          }
        }
      }
    } else if (ROWS) {
      // two halves of 16 steps: the coefficients -G[c_(sb+lp)][c_s]/v_s of a half go to registers first, then
      // the steps run fully unrolled: one shuffle and one fma each
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        if (16 * hc < nl) {
          double hreg[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int lp = 16 * hc + q;
            const int row = min(sb + lp, k - 1);
            const double gv = gram_as_double(rows[(size_t)cs.slot[row] * rs + ci]) * niv;
            hreg[q] = (valid && lp < lane && lp < nl) ? gv : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const double d = __shfl_sync(0xffffffffu, e, 16 * hc + q);
            e = fma(hreg[q], d, e);   // hreg = 0 for the lanes at or before the step: their e is final
          }
        }
      }
    } else {
      // -G[c_(sb+lp)][c_s]/v_s for the next four steps (software pipeline: loads stay off the chain)
      auto hload = [&](int lp) -> double {
        double h = 0.0;
        if (valid && lp < lane && lp < nl) h = gval(sb + lp) * niv;
        return h;
      };
      double h0 = hload(0), h1 = hload(1), h2 = hload(2), h3 = hload(3);
#pragma unroll 4
      for (int lp = 0; lp < nl; ++lp) {
        const double hcur = h0;
        h0 = h1; h1 = h2; h2 = h3; h3 = hload(lp + 4);
        const double d = __shfl_sync(0xffffffffu, e, lp);
        e = fma(hcur, d, e);   // hcur = 0 for the lanes at or before lp: their e is final
      }
    }
    if (valid) { cs.delta[sidx] = e; cs.gnew[sidx] = (cs.cls[sidx] > 0) ? gold + e : 0.0; }
    __syncwarp();
  }
}

// Candidate chain of the dense models (RR/A/L; BayesL clamps the effect, Bayes.cpp:728, so the right-hand side
// itself is chained):  rhs_s = rhs0_s - sum_{s' < s} G[c_s'][c_s] * delta_s', gnew_s = rhs_s/v + sd*z.
// The chain of the first 32 candidates is the triangular system (I - H) e = e0 with H[s][lp] = coef[32 lp + s]
// (strictly lower triangular).  Its inverse M depends only on the candidate list, so one idle warp solves it column
// by column in phase P (lane c owns column c, forward substitution) ...
// (M is stored column by column, M[c * 33 + s] = entry (row s, column c): lane c builds its own column, and in
// chain_matvec lane s reads entry (s, lp) of consecutive lanes from consecutive words)
__device__ __forceinline__ void chain_build_matrix(const double* coef, double* M, int kk, int c) {
  for (int sr = 0; sr < 32; ++sr) {
    double acc0 = (sr == c) ? 1.0 : 0.0, acc1 = 0.0;
    if (sr < kk) {
      int lp = 0;
      for (; lp + 1 < sr; lp += 2) {
        acc0 = fma(coef[lp * 32 + sr], M[c * 33 + lp], acc0);
        acc1 = fma(coef[(lp + 1) * 32 + sr], M[c * 33 + lp + 1], acc1);
      }
      if (lp < sr) acc0 = fma(coef[lp * 32 + sr], M[c * 33 + lp], acc0);
    }
    M[c * 33 + sr] = (sr < kk) ? acc0 + acc1 : 0.0;
  }
}
// ... and on the serial path the chain is one 32x32 matrix-vector product by one warp: lane s takes row s, the
// right-hand sides e0 come by shuffle, four independent partial sums (no step waits for the one before).
// The right-hand sides e0 go through shared memory (the candidates' delta array serves as the scratch: it is written
// only at the end), so a step is two independent loads and one fma -- no shuffle on the way.
__device__ __forceinline__ void chain_matvec(const CandSet& cs, int k, const double* M, int lane) {
  const bool valid = lane < k;
  const double iv = valid ? cs.iv[lane] : 0.0, gold = valid ? cs.gold[lane] : 0.0;
  const double e0 = valid ? fma(cs.rhs0[lane], iv, cs.sdz[lane]) - gold : 0.0;
  const int cls = valid ? cs.cls[lane] : 0;
  double* e0buf = cs.delta;
  e0buf[lane] = e0;
  __syncwarp();
  const double* col = M + lane;   // entry (row lane, column lp) at col[33 * lp]
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int lp = 0; lp < 32; lp += 4) {
    if (lp >= k) break;
    const double2 ea = *(const double2*)(e0buf + lp), eb = *(const double2*)(e0buf + lp + 2);
    a0 = fma(col[33 * lp], ea.x, a0);
    a1 = fma(col[33 * (lp + 1)], ea.y, a1);
    a2 = fma(col[33 * (lp + 2)], eb.x, a2);
    a3 = fma(col[33 * (lp + 3)], eb.y, a3);
  }
  const double e = (a0 + a1) + (a2 + a3);
  __syncwarp();
  if (valid) { cs.delta[lane] = e; cs.gnew[lane] = (cs.cls[lane] > 0) ? gold + e : 0.0; }
  (void)cls;
  __syncwarp();
}
template <bool ROWS>
__device__ void solve_candidates(const CandSet& cs, int k, const int32_t* __restrict__ G, const int32_t* rows, int B, int model,
                                 int lane) {
  for (int sb = 0; sb < k; sb += 32) {
    const int sidx = sb + lane;
    const bool valid = sidx < k;
    const int ci = valid ? cs.idx[sidx] : 0;
    double rhs = valid ? cs.rhs0[sidx] : 0.0;
    const double iv = valid ? cs.iv[sidx] : 0.0, sdz = valid ? cs.sdz[sidx] : 0.0, gold = valid ? cs.gold[sidx] : 0.0;
    const int cls = valid ? cs.cls[sidx] : 0;
    auto gval = [&](int sp) -> double {
      return gram_as_double(ROWS ? rows[(size_t)cs.slot[sp] * B + ci] : __ldcg(G + (size_t)cs.idx[sp] * B + ci));
    };
    for (int sp = 0; sp < sb; sp += 8) {
      double gv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[e] = valid ? gval(sp + e) : 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) rhs = fma(-gv[e], cs.delta[sp + e], rhs);
    }
    const int nl = min(32, k - sb);
    double mydelta = 0.0, mygnew = gold;
    auto gload = [&](int lp) -> double { return (valid && lp < lane && lp < nl) ? gval(sb + lp) : 0.0; };
    double g0 = gload(0), g1 = gload(1), g2 = gload(2), g3 = gload(3);
#pragma unroll 4
    for (int lp = 0; lp < nl; ++lp) {
      const double gcur = g0;
      g0 = g1; g1 = g2; g2 = g3; g3 = gload(lp + 4);
      double gn = (cls > 0) ? fma(rhs, iv, sdz) : 0.0;
      if (model == HB_MODEL_L && fabs(gn) < 1e-6) gn = 1e-6;
      const double dl = gn - gold;
      const double d = __shfl_sync(0xffffffffu, dl, lp);
      if (lane == lp) { mydelta = dl; mygnew = gn; }
      rhs = fma(-gcur, d, rhs);   // gcur = 0 for the lanes at or before lp
    }
    if (valid) { cs.delta[sidx] = mydelta; cs.gnew[sidx] = mygnew; }
    __syncwarp();
  }
}

// the same for two band blocks at once (one wait for both)
__device__ __noinline__ void gather_rows_2(int32_t* dst1, const int32_t* __restrict__ blk1, int32_t* dst2, const int32_t* __restrict__ blk2,
                                          const int* idx, int k, int B, int i) {
  for (int sb = 0; sb < k; ++sb) {
    const size_t o = (size_t)idx[sb] * B + i;
    if (blk1) {
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst1 + (size_t)sb * B + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(blk1 + o) : "memory");
    }
    if (blk2) {
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst2 + (size_t)sb * B + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(blk2 + o) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// corr_i = sum_s G[c_s][i] * delta_s over the k candidates (ascending s), read from global memory
__device__ __noinline__ double band_correction_raw(const int* idx, const double* delta, int k, const int32_t* __restrict__ gb, int B, int i) {
  // 32 Gram entries in flight: one trip to L2 for most tiles.  The correction for tile t+2 is the next thing that
  // tile's phase S waits for after the hand-over, so its latency is on the serial path of every second tile.
  double corr = 0.0;
#pragma unroll 1
  for (int sb = 0; sb < k; sb += 32) {
    int gv[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) gv[e] = (sb + e < k) ? __ldcg(gb + (size_t)idx[sb + e] * B + i) : 0;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (sb + e < k) corr = fma(gram_as_double(gv[e]), delta[sb + e], corr);
  }
  return corr;
}
__device__ __forceinline__ double band_correction(const CandSet& cs, int k, const int32_t* __restrict__ gb, int B, int i) {
  return band_correction_raw(cs.idx, cs.delta, k, gb, B, i);
}

// Tiles with more candidates than a row buffer holds (k > KROW): chain and sums straight from the Gram band in
// global memory.  Rare and slow; kept out of line so that the common path stays compact in the instruction cache.
__device__ __noinline__ double slow_chain_and_sums_cs(const CandSet& cs, int k, int myrank, const int32_t* __restrict__ G0, int B,
                                                    int i, int h, bool has1, bool dense, int model, int nthreads) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32 && k > 0) {
    if (dense) solve_candidates<false>(cs, k, G0, nullptr, B, model, lane);
    else chain_candidates<false>(cs, k, G0, nullptr, B, lane);
  }
  hb::named_bar_sync(1, nthreads);
  // the primary half returns its right-hand-side sum, the secondary half the corrections for the next tile
  if (h == 0) return band_correction(cs, myrank, G0, B, i);
  return has1 ? band_correction(cs, k, G0 + (size_t)B * B, B, i) : 0.0;
}
__device__ __noinline__ void slow_chain_only_cs(const CandSet& cs, int k, const int32_t* __restrict__ G0, int B, bool dense, int model,
                                               int nthreads) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32 && k > 0) {
    if (dense) solve_candidates<false>(cs, k, G0, nullptr, B, model, lane);
    else chain_candidates<false>(cs, k, G0, nullptr, B, lane);
  }
  hb::named_bar_sync(1, nthreads);
}
__device__ __forceinline__ void slow_chain_only(int k, const int32_t* __restrict__ G0, int B, bool dense, int model, int nthreads) {
  extern __shared__ __align__(128) uint8_t smem_slow2[];
  const CandSet cs = make_candset(smem_slow2, B);
  slow_chain_only_cs(cs, k, G0, B, dense, model, nthreads);
}
__device__ __forceinline__ double slow_chain_and_sums(int k, int myrank, const int32_t* __restrict__ G0, int B,
                                                    int i, int h, bool has1, bool dense, int model, int nthreads) {
  extern __shared__ __align__(128) uint8_t smem_slow[];
  const CandSet cs = make_candset(smem_slow, B);
  return slow_chain_and_sums_cs(cs, k, myrank, G0, B, i, h, has1, dense, model, nthreads);
}

// exact class of SNP j given rr = rhs^2 (reads its a_k, c_k and uniform from the parameter table)
template <int NF>
__device__ __noinline__ int classify_exact(const double* __restrict__ prm, size_t mp, int j, int nf, double rr, double logpi0) {
  double a[NF - 1], c[NF - 1], cum[NF];
#pragma unroll
  for (int kk = 0; kk < NF - 1; ++kk) {
    a[kk] = 0.0; c[kk] = 0.0;
    if (kk < nf - 1) { a[kk] = __ldcg(prm + prm_idx(2 + 4 * kk, mp, j)); c[kk] = __ldcg(prm + prm_idx(3 + 4 * kk, mp, j)); }
  }
  class_cum<NF>(nf, rr, a, c, logpi0, cum);
  return class_from_cum<NF>(nf, __ldcg(prm + prm_idx(0, mp, j)), cum);
}

// LEAD (experiment, HB_LEAD=1|2|4): the speculation of phase P ignores the corrections owed by the LEAD nearest tiles, as
// if it had run that many tiles earlier; the sweep's `respec`/`rounds` counters then say how often a candidate list
// built that early is incomplete (one serial CTA with packages).  LEAD = 0 is the product.
template <int NF, bool DENSE, int LEAD = 0>
__device__ void scalar_role(const SweepParams& pin, uint8_t* smem) {
  // the parameters are read all along the serial phase: keep a copy in shared memory instead of going through the
  // constant cache, which the long code of a tile keeps evicting (a miss there costs a trip to L2)
  __shared__ SweepParams ps;
  for (int w = threadIdx.x; w < (int)(sizeof(SweepParams) / 4); w += blockDim.x) ((int*)&ps)[w] = ((const int*)&pin)[w];
  __syncthreads();
  const SweepParams& p = ps;
  // Worker w (= scalar CTA w) owns tiles w, w + NG, w + 2 NG, ...  Two threads per SNP of the tile: the
  // primary half (h = 0) and the secondary half (h = 1) split the gathers and the correction sums.
  // A tile goes through three phases:
  //   P  (any time after its dots arrived)  inputs, corrections owed by the tiles t-D+1 .. t-2, speculated
  //      classes, candidate list, gather of the candidates' Gram rows
  //   S  (serial: starts when tile t-1 has handed over its corrections for tile t)  exact right-hand sides,
  //      candidate chain, verification; ends by handing the corrections for tile t+1 to the next worker
  //   C  commit: effects, classes, the tile's residual updates for the streaming CTAs, corrections for
  //      t+2 .. t+D-1
  const int B = p.B, D = p.D, T = p.T, F = p.F, model = p.model;
  const int tid = threadIdx.x;
  if (tid >= 2 * B) return;
  const int h = tid / B, i = tid - h * B, warp = i >> 5, lane = i & 31, nwarp = B / 32;
  const bool prim = (h == 0);
  const int worker = (int)blockIdx.x - p.scalar0, nworker = p.NG;
  const bool cl2 = p.cluster2 != 0;   // (only with B = 256: every thread of the block is a worker thread)
  int* ctrl = p.ctrl;
  // ---- shared memory carve-up
  CandSet cs;
  double *part_rhs, *part_corr;   // secondary half's partial sums
  int *wcnt, *rank_sh;            // rank_sh[i] = number of candidates before SNP i
  int *slot_of, *slot_snp;        // row buffers: slot of SNP i (-1: none), SNP of a slot
  volatile int* gctl;             // [0] abort flag, [1] number of candidates
  long long* phase;
  int32_t *rows0, *rows1, *rows2;
  double *coef, *cmat;            // [32][32] chain coefficients of the first 32 candidates, [32][33] solved chain matrix
  double* hbuf;                   // [2][256] cluster mode: corrections handed over by the other worker of the cluster
  {
    cs = make_candset(smem, B);
    double* d = (double*)smem;
    part_rhs = d + 6 * B; part_corr = d + 7 * B;
    d += 8 * (size_t)B;
    int* ip = (int*)d;
    rank_sh = ip + 3 * B; slot_of = ip + 4 * B; slot_snp = ip + 5 * B; ip += 6 * (size_t)B;
    wcnt = ip; ip += 64;
    gctl = ip; ip += 16;
    phase = (long long*)ip;
    uint8_t* rb = smem + scalar_fixed_bytes(B);
    coef = (double*)(rb - kChainCoefBytes);
    cmat = coef + 32 * 32;
    hbuf = cmat + 32 * 33;
    rows0 = (int32_t*)rb;
    rows1 = rows0 + (size_t)p.KROW * B;
    rows2 = rows1 + (size_t)p.KROW * B;
  }
  const int KROW = p.KROW;
  const int NT2 = 2 * B;   // threads of the worker
  if (tid < 8) gctl[tid] = 0;
  if (cl2) {
    for (int w = tid; w < 2 * 256; w += NT2) ((unsigned long long*)hbuf)[w] = kCorrEmpty;
    cluster_sync_all();   // both workers' buffers are "empty" before either hands anything over
  }
  const uint32_t hbuf_addr = (uint32_t)__cvta_generic_to_shared(hbuf);
  hb::named_bar_sync(1, NT2);

  const size_t mp = p.m_pad;
  constexpr bool dense = DENSE;   // RR / A / L: every SNP changes in every sweep
  const int nf = (model == HB_MODEL_R) ? F : 2;
  const bool use_thr = p.use_thr && !dense;
  const int DC = D - 1;   // correction slots per tile
  bool dead = false;
  int rounds_total = 0, changed_total = 0, respec = 0, widen = 0;

  long long* pc = phase;   // [16] = last time stamp
  if (tid == 0) { for (int k = 0; k < 16; ++k) pc[k] = 0; pc[16] = clock64(); }
#define HB_PHASE(n) do { if (tid == 0) { const long long _now = clock64(); pc[n] += _now - pc[16]; pc[16] = _now; } } while (0)
  for (int t = worker; t < T; t += nworker) {
    const int j = t * B + i;
    if (p.dbg & 48) {
      // timing experiments, the streaming side alone: 16 publishes every tile (without changes) as soon as
      // its dots have arrived, 32 publishes it at once
      if (!(p.dbg & 32) && prim) {
        Waiter w;
        while ((ld_relaxed_u64(p.dacc + j) & 0xffull) != (unsigned long long)p.S)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) break;
      }
      hb::named_bar_sync(1, NT2);
      if (tid == 0) st_relaxed_s32(p.tile_cnt + t, 0);
      continue;
    }
    // ---- phase P: inputs, speculation, gather of the Gram rows.  The primary half decides, the secondary half
    // only helps with gathers and sums.
    bool act = false;
    double xx = 0.0, gold = 0.0;
    double TL[NF - 1], TH[NF - 1];
    double civ[NF - 1], csdz[NF - 1];   // 1/v_k and sd_k z of every class (dense models: [0])
#pragma unroll
    for (int b = 0; b < NF - 1; ++b) { TL[b] = -1.0; TH[b] = -1.0; civ[b] = 0.0; csdz[b] = 0.0; }
    if (prim) {
      act = (j < p.m) && __ldcg(p.active + j);
      xx = __ldcg(p.xpx + j);
      gold = __ldcg(p.g + j);
#pragma unroll
      for (int b = 0; b < NF - 1; ++b)
        if (b < nf - 1) {
          if (use_thr) {
            TL[b] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * b, mp, j));
            TH[b] = __ldcg(p.prm + prm_idx(kThrField0 + 2 * b + 1, mp, j));
          }
          civ[b] = __ldcg(p.prm + prm_idx(4 + 4 * b, mp, j));
          csdz[b] = __ldcg(p.prm + prm_idx(5 + 4 * b, mp, j));
        }
    }
    // class of this SNP given its right-hand side: thresholds, or the exact evaluation (rare with thresholds)
    auto classify = [&](double rhs) -> int {
      if (dense) return 1;
      const double rr = rhs * rhs;
      if (use_thr) {
        const int c0 = thr_class<NF>(nf, rr, TL, TH);
        if (c0 >= 0) return c0;
      }
      return classify_exact<NF>(p.prm, mp, j, nf, rr, p.logpi0);
    };
    // corrections owed by the tiles t-D+1 .. t-1: whatever has been posted by now goes into the speculation
    // (no waiting here); phase S takes them all, summed oldest first
    double cspec = 0.0;
    const int dmax = min(DC, t);
    if (prim) {
      unsigned long long cw[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) cw[q] = (q + 1 <= dmax) ? ld_relaxed_u64(p.corr + ((size_t)t * DC + q) * B + i) : kCorrEmpty;
#pragma unroll
      for (int q = 7; q >= 0; --q)
        if (cw[q] != kCorrEmpty && q >= LEAD) cspec += __longlong_as_double((long long)cw[q]);
    }
    // the dots: complete when the arrival count in the low byte equals the number of slabs.  Every primary thread
    // waits politely on its own accumulator (16 cache lines per tile, 8 workers: no hot spot), so that the tile
    // starts one trip to L2 after its last dot arrived.
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
        // rows are sharded over the ranks: add this rank's exact integer part of the dot to every rank's
        // accumulator over NVLink (peer atomics), then wait until all parts have arrived here.  Integer sums do
        // not depend on the order, so every rank ends up with the same dot and takes the same decisions.
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
    HB_PHASE(0);
    if (tid == 0) HB_TRACE(t, 0);
    // rhs = x_j' yadj (+ xpx_j g_j)    (Bayes.cpp:593-594, 756-757)
    const double addback = (act && (dense || gold != 0.0)) ? xx * gold : 0.0;
    int cls = 0;
    double gnew = gold;
    // the corrections of the previous tile are still missing here: this is only the speculation
    if (act) cls = classify((base0 - cspec) + addback);
    int k = 0, myrank = 0;
    bool cand = false, fast = false;
    const bool has1 = (D > 1 && t + 1 < T);
    const bool has2 = (D > 2 && t + 2 < T);
    const int32_t* G0 = p.gram + ((size_t)t * D) * B * B;
    // Row buffers: the Gram rows of every SNP that is, or may soon become, a candidate (effect not zero, class not
    // zero, or right-hand side within 30 % of its first class boundary) are gathered once, as soon as the dots are
    // there; candidate lists are then rebuilt from these rows without touching memory again.
    bool extra = false;   // candidate that had no row yet
    int ns = 0;
    // `widen`: the wider row set is only worth its gather while speculation keeps missing (chains far from
    // equilibrium, very large n); every tile that needed a second look switches it on for this worker's next 16 tiles
    auto select_rows = [&](double rr_spec) {
      if (prim) {
        const bool near = widen > 0 && use_thr && TH[0] > 0.0 && TH[0] < 1e300 && rr_spec >= p.near_frac * TH[0];
        const bool want = act && (dense || gold != 0.0 || cls > 0 || near || extra);
        const unsigned bal = __ballot_sync(0xffffffffu, want);
        if (lane == 0) wcnt[warp] = __popc(bal);
        hb::named_bar_sync(4, B);   // primary half only
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
      fast = (ns <= KROW);
      if (fast) {
        // The far blocks of the same rows (corrections for tiles t+2 .. t+D-1, computed in phase C) come from HBM:
        // ask for them in L2 now.  The correction for tile t+2 is the last thing that tile's phase S waits for, and a
        // miss to HBM under the streaming load costs several microseconds.
        {
          const int lpr = B / 32;   // 128-byte lines per row of a block
          const int nfar = min(D, T - t) - 3;
          for (int l = tid; l < ns * lpr * nfar; l += NT2) {
            const int line = l % lpr, sl = (l / lpr) % ns, dt = 3 + l / (lpr * ns);
            const int32_t* a = G0 + (size_t)dt * B * B + (size_t)slot_snp[sl] * B + line * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
        if (prim) gather_rows(rows0, G0, slot_snp, ns, B, i);
        else if (has1) gather_rows_2(rows1, G0 + (size_t)B * B, rows2, has2 ? G0 + (size_t)2 * B * B : nullptr, slot_snp, ns, B, i);
      }
    };
    bool m_ok = false;   // the solved chain matrix matches the current candidate list
    // candidate list of the current classes
    auto compact = [&]() {
      cand = act && (cls > 0 || gold != 0.0);
      bool missing = false;
      if (prim) {
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) wcnt[warp] = __popc(bal);
        hb::named_bar_sync(4, B);   // primary half only
        int pre = 0;
        k = 0;
        for (int w = 0; w < nwarp; ++w) {
          const int c = wcnt[w];
          if (w < warp) pre += c;
          k += c;
        }
        myrank = pre + __popc(bal & ((1u << lane) - 1u));   // = number of candidates before SNP i
        rank_sh[i] = myrank;
        if (i == 0) gctl[1] = k;
        if (cand) {
          cs.idx[myrank] = i;
          cs.gold[myrank] = gold;
          cs.cls[myrank] = cls;
          cs.slot[myrank] = slot_of[i];
          missing = slot_of[i] < 0;
          double iv = 0.0, sdz = 0.0;
#pragma unroll
          for (int kk = 1; kk < NF; ++kk)
            if (kk == cls) { iv = civ[kk - 1]; sdz = csdz[kk - 1]; }
          cs.iv[myrank] = iv;
          cs.sdz[myrank] = sdz;
        }
      }
      if (hb::named_bar_or(1, NT2, missing && fast)) {
        // a candidate without a row (rare): gather again with it included
        extra = extra || missing;
        select_rows(0.0);
        if (prim && cand) cs.slot[myrank] = slot_of[i];
        hb::named_bar_sync(1, NT2);
      }
      if (!prim) { k = gctl[1]; myrank = rank_sh[i]; }
      // coefficients of the chain's first 32 candidates (read by the chain warp after the next barrier)
      m_ok = false;
      if (fast && !dense) {
        const int kk = min(k, 32);
        for (int e = tid; e < 32 * 32; e += NT2) {
          const int lp = e >> 5, sc = e & 31;
          double v = 0.0;
          if (lp < sc && sc < kk) v = gram_as_double(rows0[(size_t)cs.slot[lp] * B + cs.idx[sc]]) * (-cs.iv[sc]);
          coef[e] = v;
        }
      }
    };
    // Experiment (HB_DEBUG bit 256): the chain as a matrix-vector product delta = M e0 with the chain matrix of the current
    // candidate list solved by the first warp of the secondary half -- in phase P while the tile waits for its turn, and
    // again for a list rebuilt on the serial path, so that a tile's changes come out of the same arithmetic whether or
    // not its first speculation held (which depends on timing).  It shortens phase S by 0.4 us without shortening the
    // period, and a rebuild on the serial path costs ~7 us, which multiplies in the flip-heavy regime (many ranks, large
    // n): the default is the step-by-step chain (chain_candidates) in every round -- same arithmetic, bit for bit, too.
    auto build_matrix = [&]() {
      if (fast && !dense && k > 0 && k <= 32 && (p.dbg & 256)) {
        hb::named_bar_sync(1, NT2);
        if (!prim && warp == 0) chain_build_matrix(coef, cmat, k, lane);
        m_ok = true;
      }
    };
    select_rows(act ? ((base0 - cspec) + addback) * ((base0 - cspec) + addback) : 0.0);
    compact();
    build_matrix();
    HB_PHASE(1);
    if (tid == 0) HB_TRACE(t, 1);
    // ---- phase S: the previous tile is final once its corrections for this tile are here
    double cold = 0.0, c1 = 0.0;
    // cluster mode: the odd worker of a cluster gets the previous tile's corrections in its own shared memory
    const bool c1_local = cl2 && (worker & 1) && t >= 1 && DC >= 1;
    if (prim) {
      unsigned long long cw[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        cw[q] = (q + 1 <= dmax && !(q == 0 && c1_local)) ? ld_relaxed_u64(p.corr + ((size_t)t * DC + q) * B + i) : 0ull;
      if (c1_local) {
        // the older corrections first (they arrive earlier; asking again only after the hand-over would add a trip to L2)
#pragma unroll
        for (int q = 7; q >= 1; --q)
          if (q + 1 <= dmax && cw[q] == kCorrEmpty) {
            double v;
            if (!poll_corr_slow(p.corr + ((size_t)t * DC + q) * B + i, v, ctrl)) { dead = true; v = 0.0; }
            cw[q] = (unsigned long long)__double_as_longlong(v);
          }
        const uint32_t a = hbuf_addr + (uint32_t)((((t / nworker) & 1) * 256 + i) * 8);
        unsigned long long w = ld_own_shared_cluster_u64(a);
        if (w == kCorrEmpty) {
          // polite: 8 warps spinning on shared memory would take the issue slots of the warp that is still solving
          // the chain matrix
          Waiter wt;
          do {
            __nanosleep(64);
            if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { dead = true; w = 0ull; break; }
            w = ld_own_shared_cluster_u64(a);
          } while (w == kCorrEmpty);
        }
        st_own_shared_cluster_u64(a, kCorrEmpty);   // free for the hand-over two rounds of this worker later
        cw[0] = w;
      }
#pragma unroll
      for (int q = 7; q >= 0; --q)
        if (q + 1 <= dmax) {
          double v = __longlong_as_double((long long)cw[q]);
          if (cw[q] == kCorrEmpty) {
            const double* slot = p.corr + ((size_t)t * DC + q) * B + i;
            if (!(q == 0 ? poll_corr_staggered(slot, v, ctrl) : poll_corr_slow(slot, v, ctrl))) { dead = true; v = 0.0; }
          }
          if (q == 0) c1 = v; else cold += v;
        }
    }
    HB_PHASE(2);
    if (tid == 0) HB_TRACE(t, 2);
    const double rhs0 = ((base0 - cold) - c1) + addback;
    int nrounds = 0;
    bool respec_tile = false;
    double corr1 = 0.0;
    for (;;) {
      ++nrounds;
      if (cand && prim) cs.rhs0[myrank] = rhs0;
      // The speculation of phase P may have missed corrections that arrived later: with the complete right-hand
      // side (still without the changes inside this tile) the classes are re-decided first, and if any differs
      // the candidate list and its rows are rebuilt before the chain runs (instead of after a wasted round).
      // The barrier that publishes the candidates' right-hand sides carries the verdict.
      int cls0 = cls;
      // (only while speculation has recently missed on this worker: `widen`; otherwise the final check below is
      // enough and the serial path saves one classification)
      if (nrounds == 1 && widen > 0 && prim && act) cls0 = classify(rhs0);
      if (hb::named_bar_or(1, NT2, cls0 != cls)) {
        cls = cls0;
        compact();
        build_matrix();
        if (cand && prim) cs.rhs0[myrank] = rhs0;
        hb::named_bar_sync(1, NT2);
        ++respec;
        respec_tile = true;
      }
      HB_PHASE(3);
      double prhs = 0.0, pcorr = 0.0;
      if (fast) {
        if (tid < 32 && k > 0) {
          if (dense) solve_candidates<true>(cs, k, G0, rows0, B, model, lane);
          else if (m_ok) chain_matvec(cs, k, cmat, lane);
          else chain_candidates<true>(cs, k, G0, rows0, B, lane, coef);
        }
        HB_PHASE(4);
        hb::named_bar_sync(1, NT2);
        HB_PHASE(8);
        // exact right-hand side of every SNP of the tile:  x_i'(r - sum_{c<i} x_c delta_c), and in the same pass
        // the corrections this tile owes to the next one (all k changes).  The primary thread takes the even
        // candidates, the secondary the odd ones.
        // (a warp's 32 SNPs need the diagonal block only for the candidates before its last SNP: beyond that the loop
        // reads the block towards the next tile alone -- the pass is bound by shared-memory bandwidth)
        const int rank_hi = __shfl_sync(0xffffffffu, myrank, 31);
        int sidx = h;
#pragma unroll 4
        for (; sidx < rank_hi; sidx += 2) {
          const double d = cs.delta[sidx];
          const int sl = cs.slot[sidx];
          const double g0 = gram_as_double(rows0[(size_t)sl * B + i]), g1 = gram_as_double(rows1[(size_t)sl * B + i]);
          prhs = fma(sidx < myrank ? g0 : 0.0, d, prhs);
          pcorr = fma(has1 ? g1 : 0.0, d, pcorr);
        }
        if (has1) {
#pragma unroll 4
          for (; sidx < k; sidx += 2)
            pcorr = fma(gram_as_double(rows1[(size_t)cs.slot[sidx] * B + i]), cs.delta[sidx], pcorr);
        }
      } else {
        // more candidates than row slots: chain and sums straight from the band in global memory -- the SAME order of
        // additions as above (even / odd candidates per half), so that a tile's result does not depend on which of the
        // two paths it took: that depends on the row set, i.e. on `widen`, i.e. on timing, and ranks of a row-sharded
        // run must not drift apart in the last bit (their decisions would follow)
        slow_chain_only(k, G0, B, dense, model, NT2);
        const int32_t* G1 = G0 + (size_t)B * B;
        for (int sidx = h; sidx < k; sidx += 2) {
          const double d = cs.delta[sidx];
          const size_t o = (size_t)cs.idx[sidx] * B + i;
          const double g0 = (sidx < myrank) ? gram_as_double(__ldcg(G0 + o)) : 0.0;
          const double g1 = has1 ? gram_as_double(__ldcg(G1 + o)) : 0.0;
          prhs = fma(g0, d, prhs);
          pcorr = fma(g1, d, pcorr);
        }
      }
      if (!prim) { part_rhs[i] = prhs; part_corr[i] = pcorr; }
      HB_PHASE(9);
      hb::named_bar_sync(1, NT2);
      HB_PHASE(10);
      double rhs = rhs0;
      if (prim) {
        rhs = rhs0 - (prhs + part_rhs[i]);
        corr1 = pcorr + part_corr[i];
      }
      int cls2 = 0;
      double gnew2 = gold;
      if (act && prim) {
        cls2 = classify(rhs);
        if (dense) {
          gnew2 = fma(rhs, civ[0], csdz[0]);
          if (model == HB_MODEL_L && fabs(gnew2) < 1e-6) gnew2 = 1e-6;  // :728
        }
      }
      const bool bad = prim && act && (cls2 != cls);
      HB_PHASE(11);
      // one barrier tells everybody whether any class differs from its speculation (or a wait was abandoned)
      const bool redo = hb::named_bar_or(1, NT2, bad || dead);
      HB_PHASE(12);
      if (prim) { cls = cls2; gnew = gnew2; }
      HB_PHASE(5);
      if (!redo) break;   // the tile is final
      if (hb::named_bar_or(1, NT2, dead)) { dead = true; break; }
      // a class differed: speculate again with the corrected classes
      compact();
      build_matrix();
    }
    if (dead) break;
    rounds_total += nrounds;
    changed_total += k;
    widen = (nrounds > 1 || respec_tile) ? 16 : max(widen - 1, 0);
    // ---- the tile is final.  First what the next tile waits for: its corrections (dt = 1)
    if (has1 && prim) {
      if (cl2 && !(worker & 1))   // the next tile belongs to the other worker of this cluster
        st_peer_shared_u64(hbuf_addr + (uint32_t)(((((t + 1) / nworker) & 1) * 256 + i) * 8), cluster_ctarank() ^ 1u,
                           (unsigned long long)__double_as_longlong(corr1));
      else
        post_corr(p.corr + ((size_t)(t + 1) * DC) * B + i, corr1);
    }
    HB_PHASE(6);
    if (tid == 0) HB_TRACE(t, 3);
    // ---- then what the tile after next waits for: its corrections (dt = 2), from the rows already in shared memory.
    // Both halves sum their share of the candidates (even / odd, ascending) in the same order whether the rows are in
    // shared memory or -- more candidates than row slots -- come from the band in global memory.
    if (has2) {
      double pc2 = 0.0;
      if (fast) {
#pragma unroll 4
        for (int sidx = h; sidx < k; sidx += 2)
          pc2 = fma(gram_as_double(rows2[(size_t)cs.slot[sidx] * B + i]), cs.delta[sidx], pc2);
      } else {
        const int32_t* gb = G0 + (size_t)2 * B * B;
        for (int sidx = h; sidx < k; sidx += 2)
          pc2 = fma(gram_as_double(__ldcg(gb + (size_t)cs.idx[sidx] * B + i)), cs.delta[sidx], pc2);
      }
      if (!prim) part_corr[i] = pc2;
      hb::named_bar_sync(1, NT2);
      if (prim) post_corr(p.corr + ((size_t)(t + 2) * DC + 1) * B + i, pc2 + part_corr[i]);
    }
    // ---- phase C: commit.  The tile's residual updates go to the streaming CTAs first
    if (prim) {
      if (cand) {
        gnew = cs.gnew[myrank];
        st_relaxed_u64(p.q_delta + (size_t)t * B + myrank, (unsigned long long)__double_as_longlong(cs.delta[myrank]));
        st_relaxed_s32(p.q_snp + (size_t)t * B + myrank, j);
        p.g[j] = gnew;
      }
      if (i == 0) { st_relaxed_s32(p.tile_cnt + t, k); HB_TRACE(t, 4); }
      if (act) p.tracker[j] = cls;
    }
    // corrections owed to the tiles further ahead, whose dots were (or will be) taken before these updates
    // land: block dt goes to the half with the parity of dt
    for (int dt = 3 + h; dt < D; dt += 2) {
      if (t + dt >= T) break;
      const double cv = band_correction(cs, k, G0 + (size_t)dt * B * B, B, i);
      post_corr(p.corr + ((size_t)(t + dt) * DC + (dt - 1)) * B + i, cv);
    }
    HB_PHASE(7);
    if (tid == 0) HB_TRACE(t, 5);
    hb::named_bar_sync(1, NT2);   // the candidate arrays are free again
  }
  if (dead) atomicCAS(ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
  if (cl2) cluster_sync_all();   // nobody leaves while the other worker may still write into this CTA
  if (tid == 0) {
    if (worker < 2)
      for (int k = 0; k < 16; ++k) p.out->phase_clk[worker][k] = pc[k];
    atomicAdd(&p.out->rounds, rounds_total);
    atomicAdd(&p.out->pad, respec);   // diagnostic: tiles whose candidate list was rebuilt before the chain
    atomicAdd(&p.out->n_changed, changed_total);
  }
}

}  // namespace hbk

#include "hb_serial.cuh"

// SERIAL: the mixture models' scalar side as one serial CTA + helper CTAs (hb_serial.cuh) instead of the ring of
// workers (scalar_role); the streaming CTAs are the same.
template <int MAXT, int NF, int RL, bool DENSE, bool LIMBS = false, int LEAD = 0, bool SERIAL = false>
__global__ void __launch_bounds__(MAXT, 1) k_sweep(const __grid_constant__ SweepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  if ((int)blockIdx.x < p.S) { hbk::stream_role<RL, LIMBS>(p, smem); return; }
  if ((int)blockIdx.x < p.scalar0) return;
  if constexpr (SERIAL) {
    if ((int)blockIdx.x == p.scalar0) hbk::serial_role<NF>(p, smem);
    else hbk::helper_role<NF>(p, smem);
  } else {
    hbk::scalar_role<NF, DENSE, LEAD>(p, smem);
  }
}
