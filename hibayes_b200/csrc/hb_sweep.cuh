// hb_sweep.cuh -- the fused Gibbs sweep kernel (sm_100a).
//
// One cooperative launch per MCMC iteration replaces the switch(model_index) block of Bayes()
// (/root/reference/src/Bayes.cpp:586-816): grid = S streaming CTAs (one row slab each, one per SM)
// + 1 scalar CTA.  X is read from HBM exactly once per sweep.
//
// Streaming CTA (warp-specialised):
//   TMA warp      cp.async.bulk ring: sub-stages of SUBB SNP columns x R rows (raw int8)
//   compute warps x_j'r over the slab: the residual slab lives in registers (16 rows / thread),
//                 one PRMT + one DFMA per genotype, per-thread partials -> shared memory
//   reducer warp  fixed-order sum of the partials over row groups, fixed-point int64 atomicAdd to
//                 the per-SNP accumulators in L2 (order-independent => deterministic), tile arrival
//   AXPY warps    own the master copy of the slab's residual/u rows in registers; D tiles later
//                 they apply the effect changes the scalar CTA published (r -= x*delta, u += x*delta,
//                 Bayes.cpp:787-789, in SNP order) and republish the slab for the compute warps
//
// Scalar CTA (two thread groups ping-pong over the tiles, one thread per SNP of a tile of B SNPs):
//   turns the reduced dots of tile t into the conditional draws of Bayes.cpp:756-801.  The chain
//   x_j'r depends on every earlier change; inside a tile that dependence is the exact integer Gram
//   block G = X_t'X_t, across the D-1 tiles still in flight it is the Gram band.  Classes are
//   speculated for the whole tile, the changed SNPs ("candidates") are chained by one warp as a
//   small triangular recurrence in SNP order, every SNP is then re-evaluated with its exact
//   right-hand side and the first SNP whose class differs from the speculation restarts the round
//   (everything before it is final).  Random draws are position-addressed (hb_rng.h), so
//   re-evaluation reuses the same uniform/normal and the result equals the one-SNP-at-a-time sweep.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hibayes_b200.h"
#include "hb_device.cuh"

struct SweepOutDev {
  double count[HB_MAX_FOLD];
  double varg_acc;
  double sum_vargL;
  double sum_r, sum_r2, sum_u, var_u;
  int n_changed;
  int status;
  int rounds;        // speculation rounds summed over tiles (diagnostic)
  int pad;
  long long phase_clk[2][8];   // scalar CTA: SM cycles per phase, per thread group (diagnostic)
};

struct SweepParams {
  const uint8_t* Xp;
  double *r, *u;
  const double* xpx;
  const uint8_t* active;
  double* g;
  int32_t* tracker;
  const int32_t* gram;
  unsigned long long* dacc;
  unsigned int* arrive;
  int* q_snp;
  double* q_delta;
  int* tile_qend;   // per tile: end of its entries in the queue, -1 until published (release)
  double* corr;     // [T][D-1][B] corrections owed to tile t by tile t-dt, each value its own flag
  int* hq;          // per tile: queue position after the tile, -1 until known
  int* ctrl;  // [0] tiles committed by the scalar CTA, [1] abort code
  const double* prm;
  SweepOutDev* out;
  size_t slab_stride, m_pad;
  int n, m, S, R, T, B, D, NS;
  int NCW, NAW, SUBB, NG;
  uint32_t stage_bytes, off_rbuf, off_bar;
  int model, F;
  double fold[HB_MAX_FOLD];
  double logpi0;
  double dscale, inv_dscale, mu_shift;
  unsigned arrive_target;
  uint32_t rowbuf;   // bytes of one Gram-row buffer of the scalar CTA
  int use_thr;   // class decisions of the mixture models by certified thresholds on rhs^2 (k_prep), no exp in the chain
  int dbg;   // timing experiments only (HB_DEBUG env): 1 skip AXPY, 2 skip dot FMAs, 4 skip chain+verify
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
constexpr long long kTimeoutNs = 4000000000ll;
constexpr int kDotBars = 16;          // every wait is bounded: a lost signal aborts, it never hangs

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Waiter {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  // returns false when the wait must be abandoned (abort flag raised somewhere, or timeout)
  __device__ __forceinline__ bool keep_waiting(int* ctrl, int code) {
    if ((++spins & 0xff) == 0) {
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

__device__ __forceinline__ double byte_as_scaled(uint32_t w, uint32_t sel) {
  // genotype byte -> mantissa bits 48..55 of a double: value = byte * 2^-1026 (exact, denormal)
  return __hiloint2double((int)__byte_perm(w, 0u, sel), 0);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------
// streaming CTA
// ------------------------------------------------------------------------------------------
// Layout of one CTA (slab of R = 16*RL rows, tile of B = 32*NCW SNP columns, sub-stage = B/4 columns):
//   compute warp w, half-warp h, lane l: rows [RL*l, RL*l + RL) of the slab, held in registers for the
//   whole tile; in every sub-stage q it takes the four columns 8w + 4h + {0..3} and keeps one running
//   dot per column (16 per tile).  After the tile's four sub-stages the 16 lanes of a half-warp hold
//   16 x 16 partial dots; a transposed shuffle reduction (8+4+2+1 exchanges, fixed tree) leaves lane l
//   with the complete slab dot of accumulator l, i.e. of column 64*(l>>2)... (see col_of_acc) -- no
//   shared-memory partials, no cross-warp reduction.  The lane adds it, as fixed-point int64, to the
//   per-SNP accumulator in L2 (order-independent => deterministic).
template <int RL>
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
  uint64_t* dfull = rempty + 2;                  // kDotBars barriers: dots of tile t added to L2 by every compute warp
  int* ctrl = p.ctrl;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { hb::mbar_init(full + i, 1); hb::mbar_init(empty + i, NCW); }
    for (int i = 0; i < 2; ++i) { hb::mbar_init(rfull + i, NAW); hb::mbar_init(rempty + i, NCW); }
    for (int i = 0; i < kDotBars; ++i) hb::mbar_init(dfull + i, NCW);
    hb::mbar_fence_init();
  }
  __syncthreads();
  const uint8_t* Xs = p.Xp + (size_t)s * p.slab_stride;
  const int nsub = T * Q;

  if (warp < NCW) {
    // ---------------- compute warps
    const int h = lane >> 4, l = lane & 15;
    const uint32_t lane_off = (uint32_t)((8 * warp + 4 * h) * R + RL * l);
    double rs[RL];
    int st = 0;
    uint32_t st_par = 0;
    // dots of the previous sub-stage, reduced; their addition to L2 is issued after the next loads
    double pend = 0.0;
    size_t pend_idx = 0;
    bool have_pend = false;
    const bool red_lane = (l & 3) == 0;
    auto flush = [&]() {
      if (have_pend && red_lane) {
        const double scaled = (pend * kTwo513) * p.dscale;
        if (!(fabs(scaled) < 4.0e18)) atomicCAS(ctrl + 1, 0, HB_ABORT_OVERFLOW);
        atomicAdd(p.dacc + pend_idx, (unsigned long long)__double2ll_rn(scaled));
      }
      have_pend = false;
    };
    for (int t = 0; t < T; ++t) {
      if (!mbar_wait(rfull + (t & 1), (uint32_t)((t >> 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
      {
        const double2* rb = (const double2*)(rbuf + (size_t)(t & 1) * R + RL * l);
#pragma unroll
        for (int i = 0; i < RL / 2; ++i) { const double2 v = rb[i]; rs[2 * i] = v.x; rs[2 * i + 1] = v.y; }
      }
      __syncwarp();
      if (lane == 0) hb::mbar_arrive(rempty + (t & 1));
#pragma unroll 1
      for (int q = 0; q < Q; ++q) {
        if (!mbar_wait(full + st, st_par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        {
          const uint8_t* sp = stage0 + (size_t)st * p.stage_bytes + lane_off;
          uint2 v0 = *(const uint2*)(sp), v1 = *(const uint2*)(sp + R), v2 = *(const uint2*)(sp + 2 * R),
                v3 = *(const uint2*)(sp + 3 * R);
          flush();   // the previous sub-stage's dots go out while these loads are in flight
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
        // transposed reduction of the four dots over the 16 lanes of the half-warp (fixed tree):
        // lanes 4c .. 4c+3 end up with the slab dot of column c of this warp's four
        {
          const bool up8 = (l & 8) != 0, up4 = (l & 4) != 0;
          const double k0 = up8 ? a2 : a0, s0 = up8 ? a0 : a2;
          const double k1 = up8 ? a3 : a1, s1 = up8 ? a1 : a3;
          const double b0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
          const double b1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
          const double k2 = up4 ? b1 : b0, s2 = up4 ? b0 : b1;
          double c0 = k2 + __shfl_xor_sync(0xffffffffu, s2, 4);
          c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
          c0 += __shfl_xor_sync(0xffffffffu, c0, 1);
          pend = c0;
          // lanes with bit 3 set hold columns 2,3; bit 2 selects the odd one
          const int kcol = ((l >> 3) & 1) * 2 + ((l >> 2) & 1);
          pend_idx = (size_t)t * B + q * SUBB + 8 * warp + 4 * h + kcol;
          have_pend = true;
        }
      }
      flush();
      __syncwarp();
      if (lane == 0) hb::mbar_arrive(dfull + (t % kDotBars));
    }
    return;
  }
  if (warp == NCW) {
    // ---------------- TMA producer: streams this slab's sub-stages through the ring
    if (lane == 0) {
      int st = 0;
      uint32_t par = 1;   // parity of the previous phase of empty[st]
      for (int gsub = 0; gsub < nsub; ++gsub) {
        if (gsub >= NS && !mbar_wait(empty + st, par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
        hb::mbar_arrive_expect_tx(full + st, p.stage_bytes);
        hb::tma_load_1d(stage0 + (size_t)st * p.stage_bytes, Xs + (size_t)gsub * p.stage_bytes, p.stage_bytes, full + st);
        if (++st == NS) { st = 0; par ^= 1u; }
      }
    }
    return;
  }
  if (warp == NCW + 1) {
    // ---------------- committer: once every compute warp has added its dots of tile t, make them visible
    // device-wide and count the slab in (the fence stalls only this warp)
    if (lane == 0) {
      for (int t = 0; t < T; ++t) {
        if (!mbar_wait(dfull + (t % kDotBars), (uint32_t)((t / kDotBars) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
        __threadfence();
        atomicAdd(p.arrive + t, 1u);
      }
    }
    return;
  }
  if (warp < NCW + 2 + NAW) {
    // ---------------- AXPY warps: master copy of the slab's residual and u rows (4 rows / thread,
    // held as value * 2^-513 so that the denormal genotype factors stay exact)
    const int a = tid - 32 * (NCW + 2);
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
    int applied = 0;
    for (int t = 0; t < T + D; ++t) {
      if (t >= D) {
        // residual updates published by the scalar workers for tile t-D
        int qend = 0;
        if (lane == 0) {
          qend = hb::ld_acquire(p.tile_qend + (t - D));
          if (qend < 0) {
            Waiter w;
            while ((qend = hb::ld_acquire(p.tile_qend + (t - D))) < 0)
              if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_STREAM)) break;
          }
        }
        qend = __shfl_sync(0xffffffffu, qend, 0);
        if (qend < 0) return;
        // the tile's changes: 32 queue entries per coalesced read, then the genotype words of eight
        // changed SNPs in flight at a time (all from L2: the columns were streamed D tiles ago)
        for (int q0 = applied; q0 < ((p.dbg & 1) ? applied : qend); q0 += 32) {
          const int nq = min(32, qend - q0);
          int jl = 0;
          double dll = 0.0;
          if (lane < nq) { jl = __ldcg(p.q_snp + q0 + lane); dll = __ldcg(p.q_delta + q0 + lane); }
          for (int e0 = 0; e0 < nq; e0 += 8) {
            uint32_t xw[8];
            double dl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int js = __shfl_sync(0xffffffffu, jl, (e0 + e) & 31);
              dl[e] = __shfl_sync(0xffffffffu, dll, (e0 + e) & 31);
              xw[e] = (e0 + e < nq && has) ? __ldg((const uint32_t*)(Xs + (size_t)js * R + row0)) : 0u;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e0 + e < nq) {
                const double ds = dl[e] * kTwo513;
                const double x0 = byte_as_scaled(xw[e], 0x4044), x1 = byte_as_scaled(xw[e], 0x4144);
                const double x2 = byte_as_scaled(xw[e], 0x4244), x3 = byte_as_scaled(xw[e], 0x4344);
                rm[0] = fma(-x0, ds, rm[0]); um[0] = fma(x0, ds, um[0]);   // yadj -= x*delta (Bayes.cpp:787), u += x*delta (:789)
                rm[1] = fma(-x1, ds, rm[1]); um[1] = fma(x1, ds, um[1]);
                rm[2] = fma(-x2, ds, rm[2]); um[2] = fma(x2, ds, um[2]);
                rm[3] = fma(-x3, ds, rm[3]); um[3] = fma(x3, ds, um[3]);
              }
          }
        }
        applied = qend;
      }
      if (t < T) {
        if (t >= 2 && !mbar_wait(rempty + (t & 1), (uint32_t)(((t >> 1) - 1) & 1), ctrl, HB_ABORT_TIMEOUT_PIPE)) return;
        if (has) {
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
// scalar CTA
// ------------------------------------------------------------------------------------------
template <int NF>
struct SnpPrm {
  double u;
  double a[NF - 1], c[NF - 1], iv[NF - 1], sdz[NF - 1];
};

// Cumulative class probabilities of one SNP given rr = rhs^2 (Bayes.cpp:759-770 in soft-max form).
template <int NF>
__device__ __forceinline__ void class_cum(int nf, double rr, const double* a, const double* c, double logpi0, double* cum) {
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
__device__ __forceinline__ int class_from_cum(int nf, double u, const double* cum) {
  int cls = 0;
  bool found = false;
#pragma unroll
  for (int k = 0; k < NF; ++k)
    if (k < nf && !found && u < cum[k]) { cls = k; found = true; }
  return cls;
}

// Conditional draw of one SNP given its right-hand side (Bayes.cpp:592-601, 639-664, 756-801).
// `nf` = number of mixture classes in play (2 for B/C, F for R); dense models always return class 1.
template <int NF>
__device__ __forceinline__ void eval_snp(int model, int nf, double rhs, const SnpPrm<NF>& q, double logpi0, int& cls, double& gnew) {
  if (model == HB_MODEL_RR || model == HB_MODEL_A || model == HB_MODEL_L) {
    cls = 1;
    gnew = fma(rhs, q.iv[0], q.sdz[0]);
    if (model == HB_MODEL_L && fabs(gnew) < 1e-6) gnew = 1e-6;  // :728
    return;
  }
  double cum[NF];
  class_cum<NF>(nf, rhs * rhs, q.a, q.c, logpi0, cum);
  cls = class_from_cum<NF>(nf, q.u, cum);
  gnew = 0.0;
#pragma unroll
  for (int k = 1; k < NF; ++k)
    if (k == cls) gnew = fma(rhs, q.iv[k - 1], q.sdz[k - 1]);
}

// Class as a step function of rr = rhs^2.  With class-ordered slopes c_1 <= c_2 <= ... every cumulative
// probability cum_b(rr) decreases in rr, so "first b with u < cum_b" equals "first b with rr below the root
// theta_b of cum_b = u".  k_prep brackets each root as [TL_b, TH_b] and certifies both ends with class_cum
// itself; inside a bracket (or where certification failed: TL = -1, TH = inf) the caller evaluates exactly.
// Returns the class, or -1 when rr falls inside a bracket.
template <int NF>
__device__ __forceinline__ int thr_class(int nf, double rr, const double* TL, const double* TH) {
#pragma unroll
  for (int b = 0; b < NF - 1; ++b)
    if (b < nf - 1) {
      if (rr <= TL[b]) return b;
      if (!(rr >= TH[b])) return -1;
    }
  return nf - 1;
}

// Brackets of the class boundaries of one SNP (device, k_prep).  phi_b(rr) = log S_hi - log S_lo - log((1-u)/u)
// with S_lo = sum_{l<=b} e^{s_l}, S_hi = sum_{l>b} e^{s_l}, s_0 = log pi_0, s_l = a_l + c_l rr, is increasing;
// safeguarded Newton finds its root, the bracket is widened by 1e-9 relative and both ends are certified with
// a margin far above the rounding error of class_cum.
template <int NF>
__device__ void solve_thresholds(int nf, double u, const double* a, const double* c, double logpi0, double* TL, double* TH) {
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
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
    if (f > 1e-9) {
      // never at or below class b: certify at rr = 0 (cum_b decreases from there)
      double cum[NF];
      class_cum<NF>(nf, 0.0, a, c, logpi0, cum);
      if (cum[b] < u * (1.0 - 1e-11)) { TL[b] = -1.0; TH[b] = -1.0; }
      continue;
    }
    if (f < -1e-9) {
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
    const double w = 1e-9 * theta + 2e-10 / df;
    const double tl = theta - w, th = theta + w;
    double cum[NF];
    class_cum<NF>(nf, th, a, c, logpi0, cum);
    if (!(cum[b] < u * (1.0 - 1e-11))) continue;
    if (tl >= 0.0) {
      class_cum<NF>(nf, tl, a, c, logpi0, cum);
      if (!(cum[b] > u * (1.0 + 1e-11))) continue;
      TL[b] = tl;
    }
    TH[b] = th;
  }
}

// candidate arrays of one thread group (B entries each)
struct CandSet {
  double *rhs0, *iv, *sdz, *gold, *delta, *gnew;
  int *idx, *cls;
};

// Shared memory of a scalar CTA: two workers (thread groups of B threads), each with its candidate arrays
// and two row buffers into which the Gram rows of its tile's candidates are gathered by TMA bulk copies
// (one 4B-byte row per candidate and band block).
__host__ __device__ inline size_t scalar_fixed_bytes(int B, int D) {
  (void)D;
  size_t b = (12 * (size_t)B) * 8 + (4 * (size_t)B + 64 + 16) * 4 + 32 + 18 * 8;   // candidates, ints, barriers, timers
  return (b + 127) / 128 * 128;
}
// largest row buffer (multiple of 1 KB, at most 44 KB) that still fits the 227 KB of an SM
__host__ inline size_t scalar_rowbuf_bytes(int B, int D) {
  const size_t cap = 227 * 1024 - 1024;
  const size_t fixed = scalar_fixed_bytes(B, D);
  size_t rb = fixed < cap ? (cap - fixed) / 4 : 0;
  rb = rb / 1024 * 1024;
  return rb > 44 * 1024 ? 44 * 1024 : rb;
}
__host__ inline size_t scalar_smem_bytes(int B, int D) {
  return scalar_fixed_bytes(B, D) + 4 * scalar_rowbuf_bytes(B, D);
}

// A correction slot that has not been written yet holds this NaN payload (k_prep fills the array).
constexpr unsigned long long kCorrEmpty = 0x7ff8dead0badf00dull;
// thread-private hand-off: spin on one 8-byte word until its producer has stored a value
__device__ __forceinline__ bool poll_corr(const double* slot, double& v, int* ctrl) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(slot) : "memory");
  if (w == kCorrEmpty) {
    Waiter wt;
    do {
      if (!wt.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) return false;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(slot) : "memory");
    } while (w == kCorrEmpty);
  }
  v = __longlong_as_double((long long)w);
  return true;
}
__device__ __forceinline__ void post_corr(double* slot, double v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(__double_as_longlong(v)) : "memory");
}

// TMA gather of the Gram rows of the k candidates from one band block into a row buffer (one warp)
__device__ __forceinline__ void issue_gather(int32_t* dst, const int32_t* __restrict__ blk, const CandSet& cs, int k, int B,
                                             uint64_t* bar, int lane) {
  hb::fence_proxy_async();
  if (lane == 0) hb::mbar_arrive_expect_tx(bar, (uint32_t)(k * B * 4));
  __syncwarp();
  for (int sidx = lane; sidx < k; sidx += 32)
    hb::tma_load_1d(dst + (size_t)sidx * B, blk + (size_t)cs.idx[sidx] * B, (uint32_t)(B * 4), bar);
}

// Chains the k candidates of a tile in SNP order (one warp).  Candidate s has right-hand side
//   rhs_s = rhs0_s - sum_{s' < s} G[c_s'][c_s] * delta_s'      (ascending s', one fma each)
// and effect  gnew_s = class > 0 ? rhs_s/v + sd*z : 0,  delta_s = gnew_s - gold_s.
// ROWS: G rows of the candidates are in shared memory (rows[s'][.]); otherwise they are read from global.
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
    // candidates of earlier chunks: their deltas are final
    for (int sp = 0; sp < sb; sp += 8) {
      int gv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[e] = !valid ? 0 : ROWS ? rows[(size_t)(sp + e) * B + ci] : __ldg(G + (size_t)cs.idx[sp + e] * B + ci);
#pragma unroll
      for (int e = 0; e < 8; ++e) rhs = fma(-(double)gv[e], cs.delta[sp + e], rhs);
    }
    const int nl = min(32, k - sb);
    double mydelta = 0.0, mygnew = gold;
    // G[c_(sb+lp)][c_s] for the next four steps (software pipeline: the load latency stays off the chain)
    auto gload = [&](int lp) -> double {
      int gv = 0;
      if (valid && lp < lane && lp < nl) gv = ROWS ? rows[(size_t)(sb + lp) * B + ci] : __ldg(G + (size_t)cs.idx[sb + lp] * B + ci);
      return (double)gv;
    };
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

// corr_i = sum_s G[c_s][i] * delta_s over the k candidates (ascending s), rows in shared memory
__device__ __forceinline__ double band_correction_rows(const CandSet& cs, int k, const int32_t* rows, int B, int i) {
  double corr = 0.0;
#pragma unroll 4
  for (int sidx = 0; sidx < k; ++sidx) corr = fma((double)rows[(size_t)sidx * B + i], cs.delta[sidx], corr);
  return corr;
}
// same from global memory (tiles with more candidates than a row buffer holds)
__device__ __forceinline__ double band_correction(const CandSet& cs, int k, const int32_t* __restrict__ gb, int B, int i) {
  double corr = 0.0;
  for (int sb = 0; sb < k; sb += 16) {
    int gv[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) gv[e] = (sb + e < k) ? __ldg(gb + (size_t)cs.idx[sb + e] * B + i) : 0;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (sb + e < k) corr = fma((double)gv[e], cs.delta[sb + e], corr);
  }
  return corr;
}

// exact int32 -> double for 0 <= g < 2^31 on the full-rate pipe (one DADD instead of a quarter-rate I2F)
__device__ __forceinline__ double gram_as_double(int g) {
  return __hiloint2double(0x43300000, g) - 4503599627370496.0;
}

// Candidate chain of the mixture models (one warp, rows in shared memory).  Lane s tracks
//   e_s = rhs_s/v_s + sd_s z_s - gold_s        (its effect change if nothing before it in the tile changed)
// and every earlier candidate s' lowers it by (G[c_s'][c_s]/v_s) delta_s'; when the chain reaches lane s its
// e_s is final (= delta_s).  One shuffle and one fma per candidate on the dependent path.
__device__ void chain_candidates(const CandSet& cs, int k, const int32_t* rows, int B, int lane) {
  for (int sb = 0; sb < k; sb += 32) {
    const int sidx = sb + lane;
    const bool valid = sidx < k;
    const int ci = valid ? cs.idx[sidx] : 0;
    const double iv = valid ? cs.iv[sidx] : 0.0, gold = valid ? cs.gold[sidx] : 0.0;
    const double niv = -iv;
    double e = valid ? fma(cs.rhs0[sidx], iv, cs.sdz[sidx]) - gold : 0.0;
    // candidates of earlier chunks: their changes are final
    for (int sp = 0; sp < sb; sp += 8) {
      double gv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) gv[q] = valid ? gram_as_double(rows[(size_t)(sp + q) * B + ci]) * niv : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) e = fma(gv[q], cs.delta[sp + q], e);
    }
    const int nl = min(32, k - sb);
    // -G[c_(sb+lp)][c_s]/v_s for the next four steps (software pipeline: loads and conversions stay off the chain)
    auto hload = [&](int lp) -> double {
      double h = 0.0;
      if (valid && lp < lane && lp < nl) h = gram_as_double(rows[(size_t)(sb + lp) * B + ci]) * niv;
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
    if (valid) { cs.delta[sidx] = e; cs.gnew[sidx] = (cs.cls[sidx] > 0) ? gold + e : 0.0; }
    __syncwarp();
  }
}

template <int NF>
__device__ void scalar_role(const SweepParams& p, uint8_t* smem) {
  // The scalar CTAs hold 2 workers each (thread groups of B threads, one thread per SNP of a tile); worker w
  // owns tiles w, w + W, w + 2W, ...  A tile goes through three phases:
  //   P  (any time after its dots arrived)  inputs, corrections owed by the tiles t-D+1 .. t-2, speculated
  //      classes, candidate list, TMA gather of the candidates' Gram rows
  //   S  (serial: starts when tile t-1 has handed over its corrections for tile t)  exact right-hand sides,
  //      candidate chain, verification; ends by handing the corrections for tile t+1 to the next worker
  //   C  commit: effects, classes, residual-update queue for the streaming CTAs, corrections for t+2 .. t+D-1
  // Workers exchange data through global memory only: every correction value is its own flag (poll_corr).
  const int B = p.B, D = p.D, T = p.T, F = p.F, model = p.model;
  const int tid = threadIdx.x;
  const int ngrp = 2;
  if (tid >= ngrp * B) return;
  const int grp = tid / B, i = tid - grp * B, warp = i >> 5, lane = i & 31, nwarp = B / 32;
  const int worker = ((int)blockIdx.x - p.S) * ngrp + grp, nworker = p.NG * ngrp;
  int* ctrl = p.ctrl;
  // ---- shared memory carve-up
  CandSet cs;        // this worker's candidate arrays
  int *wcnt, *wbad;  // 16 per worker
  volatile int* gctl; // per worker: [0] abort flag, [1] qbase
  uint64_t* rbar;    // this worker's two row-buffer barriers
  int32_t *rows0, *rows1;
  long long* phase;
  {
    double* d = (double*)smem;
    double* cbase = d + (size_t)grp * 6 * B; d += 12 * (size_t)B;
    cs.rhs0 = cbase; cs.iv = cbase + B; cs.sdz = cbase + 2 * B; cs.gold = cbase + 3 * B;
    cs.delta = cbase + 4 * B; cs.gnew = cbase + 5 * B;
    int* ip = (int*)d;
    cs.idx = ip + (size_t)grp * 2 * B; cs.cls = cs.idx + B; ip += 4 * (size_t)B;
    wcnt = ip + grp * 16; wbad = ip + 32 + grp * 16; ip += 64;
    gctl = ip + grp * 8; ip += 16;
    rbar = (uint64_t*)ip + 2 * grp;
    phase = (long long*)((uint64_t*)ip + 4);
    uint8_t* rb = smem + scalar_fixed_bytes(B, D);
    rows0 = (int32_t*)(rb + (size_t)(2 * grp) * p.rowbuf);
    rows1 = (int32_t*)(rb + (size_t)(2 * grp + 1) * p.rowbuf);
  }
  const int KROW = (int)(p.rowbuf / ((size_t)B * 4));
  const int gbar = 2 + grp;
  if (i < 8) gctl[i] = 0;
  if (i == 0) { hb::mbar_init(rbar, 1); hb::mbar_init(rbar + 1, 1); hb::mbar_fence_init(); }
  hb::named_bar_sync(1, ngrp * B);

  const size_t mp = p.m_pad;
  const bool dense = (model == HB_MODEL_RR || model == HB_MODEL_A || model == HB_MODEL_L);
  const int nf = (model == HB_MODEL_R) ? F : 2;
  const bool use_thr = p.use_thr && !dense;
  const int NONE = 1 << 30;
  const int DC = D - 1;   // correction slots per tile
  unsigned n0 = 0, n1 = 0;   // gathers issued so far into rows0 / rows1 (parity of the phase to wait for)
  bool dead = false;
  int rounds_total = 0;

  long long* pc = phase + grp * 9;   // [8] = last time stamp
  if (i == 0) { for (int k = 0; k < 8; ++k) pc[k] = 0; pc[8] = clock64(); }
#define HB_PHASE(n) do { if (i == 0) { const long long _now = clock64(); pc[n] += _now - pc[8]; pc[8] = _now; } } while (0)
  for (int t = worker; t < T; t += nworker) {
    const int j = t * B + i;
    if (p.dbg & 48) {
      // timing experiments, the streaming side alone: 16 publishes every tile (without changes) as soon as
      // its dots have arrived, 32 publishes it at once
      if (i == 0) {
        bool ok = true;
        if (!(p.dbg & 32)) {
          Waiter w;
          while (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target)
            if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
        }
        if (ok) { __threadfence(); hb::st_release(p.tile_qend + t, 0); }
      }
      continue;
    }
    // ---- phase P: inputs, speculation, Gram-row gather
    const bool act = (j < p.m) && p.active[j];
    const double xx = p.xpx[j];
    const double gold = p.g[j];
    double TL[NF - 1], TH[NF - 1];
    double div0 = 0.0, dsdz0 = 0.0;   // dense models: 1/v and sd*z
    if (use_thr) {
#pragma unroll
      for (int b = 0; b < NF - 1; ++b) {
        TL[b] = -1.0; TH[b] = -1.0;
        if (b < nf - 1) {
          TL[b] = p.prm[prm_idx(kThrField0 + 2 * b, mp, j)];
          TH[b] = p.prm[prm_idx(kThrField0 + 2 * b + 1, mp, j)];
        }
      }
    } else if (dense) {
      div0 = p.prm[prm_idx(4, mp, j)];
      dsdz0 = p.prm[prm_idx(5, mp, j)];
    }
    // class of this SNP given its right-hand side: thresholds, or the exact evaluation (rare with thresholds)
    auto classify = [&](double rhs) -> int {
      if (dense) return 1;
      const double rr = rhs * rhs;
      if (use_thr) {
        const int c0 = thr_class<NF>(nf, rr, TL, TH);
        if (c0 >= 0) return c0;
      }
      double a[NF - 1], c[NF - 1], cum[NF];
#pragma unroll
      for (int kk = 0; kk < NF - 1; ++kk) {
        a[kk] = 0.0; c[kk] = 0.0;
        if (kk < nf - 1) { a[kk] = p.prm[prm_idx(2 + 4 * kk, mp, j)]; c[kk] = p.prm[prm_idx(3 + 4 * kk, mp, j)]; }
      }
      class_cum<NF>(nf, rr, a, c, p.logpi0, cum);
      return class_from_cum<NF>(nf, p.prm[prm_idx(0, mp, j)], cum);
    };
    if (i == 0) {
      bool ok = true;
      if (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target) {
        Waiter w;
        while (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
      }
      hb::fence_acq_rel_gpu();
      gctl[0] = (!ok || *((volatile int*)(ctrl + 1)) != 0) ? 1 : 0;
    }
    hb::named_bar_sync(gbar, B);
    if (gctl[0]) { dead = true; break; }
    HB_PHASE(0);
    const double base0 = (double)(long long)__ldcg(p.dacc + j) * p.inv_dscale;
    // rhs = x_j' yadj (+ xpx_j g_j)    (Bayes.cpp:593-594, 756-757)
    const double addback = (act && (dense || gold != 0.0)) ? xx * gold : 0.0;
    // corrections owed by the tiles t-D+1 .. t-2 (oldest first); the one of tile t-1 comes in phase S
    double cold = 0.0;
    for (int dt = min(DC, t); dt >= 2; --dt) {
      double v;
      if (!poll_corr(p.corr + ((size_t)t * DC + (dt - 1)) * B + i, v, ctrl)) { dead = true; v = 0.0; }
      cold += v;
    }
    int cls = 0;
    double gnew = gold;
    // the corrections of the previous tile are still missing here: this is only the speculation
    if (act) cls = classify((base0 - cold) + addback);
    int k = 0, myrank = 0;
    bool cand = false, fast = false;
    const bool has1 = (D > 1 && t + 1 < T);
    const int32_t* G0 = p.gram + ((size_t)t * D) * B * B;
    // candidate list of the speculated classes; used as is by the first round of the serial phase
    auto compact = [&]() {
      cand = act && (cls > 0 || gold != 0.0);
      const unsigned bal = __ballot_sync(0xffffffffu, cand);
      if (lane == 0) wcnt[warp] = __popc(bal);
      hb::named_bar_sync(gbar, B);
      int pre = 0;
      k = 0;
      for (int w = 0; w < nwarp; ++w) {
        const int c = wcnt[w];
        if (w < warp) pre += c;
        k += c;
      }
      myrank = pre + __popc(bal & ((1u << lane) - 1u));   // = number of candidates before SNP i
      if (cand) {
        cs.idx[myrank] = i;
        cs.gold[myrank] = gold;
        cs.cls[myrank] = cls;
        double iv = 0.0, sdz = 0.0;
        if (dense) { iv = div0; sdz = dsdz0; }
        else if (cls > 0) { iv = p.prm[prm_idx(4 * cls, mp, j)]; sdz = p.prm[prm_idx(4 * cls + 1, mp, j)]; }   // 1/v_k, sd_k z of class k
        cs.iv[myrank] = iv;
        cs.sdz[myrank] = sdz;
      }
      hb::named_bar_sync(gbar, B);
      fast = (k <= KROW);
      if (fast) {
        if (warp == 0) {
          issue_gather(rows0, G0, cs, k, B, rbar, lane);
          if (has1) issue_gather(rows1, G0 + (size_t)B * B, cs, k, B, rbar + 1, lane);
        }
        ++n0;
        if (has1) ++n1;
      }
    };
    compact();
    HB_PHASE(1);
    // ---- phase S: the previous tile is final once its corrections for this tile are here
    double c1 = 0.0;
    if (t >= 1 && D > 1 && !poll_corr(p.corr + ((size_t)t * DC) * B + i, c1, ctrl)) dead = true;
    HB_PHASE(2);
    const double rhs0 = ((base0 - cold) - c1) + addback;
    int nrounds = 0;
    bool rows1_pending = false;
    double corr1 = 0.0;
    for (;;) {
      ++nrounds;
      if (cand) cs.rhs0[myrank] = rhs0;
      hb::named_bar_sync(gbar, B);
      if (fast && !hbk::mbar_wait(rbar, (n0 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
      HB_PHASE(3);
      if (warp == 0 && k > 0 && !dead) {
        if (fast && !dense) chain_candidates(cs, k, rows0, B, lane);
        else if (fast) solve_candidates<true>(cs, k, G0, rows0, B, model, lane);
        else solve_candidates<false>(cs, k, G0, rows0, B, model, lane);
      }
      if (fast && has1 && !hbk::mbar_wait(rbar + 1, (n1 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
      rows1_pending = false;
      hb::named_bar_sync(gbar, B);
      HB_PHASE(4);
      // exact right-hand side of every SNP of the tile:  x_i'(r - sum_{c<i} x_c delta_c), ascending c; in the
      // same pass the corrections this tile owes to the next one (all k changes)
      double rhs = rhs0;
      corr1 = 0.0;
      if (fast) {
        if (!dead) {
          if (has1) {
#pragma unroll 4
            for (int sidx = 0; sidx < k; ++sidx) {
              const double d = cs.delta[sidx];
              if (sidx < myrank) rhs = fma(-gram_as_double(rows0[(size_t)sidx * B + i]), d, rhs);
              corr1 = fma(gram_as_double(rows1[(size_t)sidx * B + i]), d, corr1);
            }
          } else {
#pragma unroll 4
            for (int sidx = 0; sidx < myrank; ++sidx) rhs = fma(-gram_as_double(rows0[(size_t)sidx * B + i]), cs.delta[sidx], rhs);
          }
        }
      } else {
        for (int sb = 0; sb < myrank; sb += 8) {
          int gv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) gv[e] = (sb + e < myrank) ? __ldg(G0 + (size_t)cs.idx[sb + e] * B + i) : 0;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (sb + e < myrank) rhs = fma(-(double)gv[e], cs.delta[sb + e], rhs);
        }
        if (has1) corr1 = band_correction(cs, k, G0 + (size_t)B * B, B, i);
      }
      int cls2 = 0;
      double gnew2 = gold;
      if (act) {
        cls2 = classify(rhs);
        if (dense) {
          gnew2 = fma(rhs, div0, dsdz0);
          if (model == HB_MODEL_L && fabs(gnew2) < 1e-6) gnew2 = 1e-6;  // :728
        }
      }
      const bool bad = act && (cls2 != cls);
      const unsigned bal2 = __ballot_sync(0xffffffffu, bad);
      if (lane == 0) wbad[warp] = bal2 ? (warp * 32 + __ffs(bal2) - 1) : NONE;
      if (__any_sync(0xffffffffu, dead) && lane == 0) wbad[warp] = -1;
      hb::named_bar_sync(gbar, B);
      int first = NONE;
      for (int w = 0; w < nwarp; ++w) first = min(first, wbad[w]);
      cls = cls2;
      gnew = gnew2;
      HB_PHASE(5);
      if (first == NONE) break;   // every class equals its speculation: the tile is final
      if (first < 0) { dead = true; break; }
      // a class differed: everything before it is final; speculate again with the corrected classes
      if (rows1_pending && !hbk::mbar_wait(rbar + 1, (n1 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
      compact();
    }
    if (dead) break;
    rounds_total += nrounds;
    // ---- the tile is final.  First what the next tile waits for: its corrections (dt = 1)
    if (has1) post_corr(p.corr + ((size_t)(t + 1) * DC) * B + i, corr1);
    HB_PHASE(6);
    // ---- phase C: commit.  Position of this tile's changes in the residual-update queue
    if (i == 0) {
      int qb = 0;
      bool ok = true;
      if (t > 0) {
        qb = hb::ld_relaxed(p.hq + (t - 1));
        if (qb < 0) {
          Waiter w;
          while ((qb = hb::ld_relaxed(p.hq + (t - 1))) < 0)
            if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
        }
      }
      if (ok) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.hq + t), "r"(qb + k) : "memory");
      gctl[1] = qb;
      gctl[0] = ok ? 0 : 1;
    }
    hb::named_bar_sync(gbar, B);
    if (gctl[0]) { dead = true; break; }
    const int qbase = gctl[1];
    if (cand) gnew = cs.gnew[myrank];
    if (act) {
      p.g[j] = gnew;
      p.tracker[j] = cls;
    }
    if (cand) {
      p.q_snp[qbase + myrank] = j;
      p.q_delta[qbase + myrank] = cs.delta[myrank];
    }
    // corrections owed to the tiles further ahead, whose dots were (or will be) taken before these updates land
    for (int dt = 2; dt < D; dt += 2) {
      if (t + dt >= T) break;
      const bool two = (dt + 1 < D) && (t + dt + 1 < T);
      double ca, cb = 0.0;
      if (fast) {
        if (warp == 0) {
          issue_gather(rows0, G0 + (size_t)dt * B * B, cs, k, B, rbar, lane);
          if (two) issue_gather(rows1, G0 + (size_t)(dt + 1) * B * B, cs, k, B, rbar + 1, lane);
        }
        ++n0;
        if (two) ++n1;
        if (!hbk::mbar_wait(rbar, (n0 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
        ca = dead ? 0.0 : band_correction_rows(cs, k, rows0, B, i);
        if (two) {
          if (!hbk::mbar_wait(rbar + 1, (n1 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
          cb = dead ? 0.0 : band_correction_rows(cs, k, rows1, B, i);
        }
        if (dt + 2 < D) hb::named_bar_sync(gbar, B);   // every thread is done with the row buffers before they are refilled
      } else {
        ca = band_correction(cs, k, G0 + (size_t)dt * B * B, B, i);
        if (two) cb = band_correction(cs, k, G0 + (size_t)(dt + 1) * B * B, B, i);
      }
      post_corr(p.corr + ((size_t)(t + dt) * DC + (dt - 1)) * B + i, ca);
      if (two) post_corr(p.corr + ((size_t)(t + dt + 1) * DC + dt) * B + i, cb);
    }
    // publish the tile's residual updates to the streaming CTAs
    __threadfence();
    hb::named_bar_sync(gbar, B);
    if (i == 0) hb::st_release(p.tile_qend + t, qbase + k);
    if (i == 0 && t == T - 1) p.out->n_changed = qbase + k;
    HB_PHASE(7);
  }
  if (dead) atomicCAS(ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
  if (i == 0) {
    if (worker < 2)
      for (int k = 0; k < 8; ++k) p.out->phase_clk[worker][k] = pc[k];
    atomicAdd(&p.out->rounds, rounds_total);
  }
}

}  // namespace hbk

template <int MAXT, int NF, int RL>
__global__ void __launch_bounds__(MAXT, 1) k_sweep(const __grid_constant__ SweepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  if ((int)blockIdx.x >= p.S) hbk::scalar_role<NF>(p, smem);
  else hbk::stream_role<RL>(p, smem);
}
