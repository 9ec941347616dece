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
  int* tile_qend;
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
  int dbg;   // timing experiments only (HB_DEBUG env): 1 skip AXPY, 2 skip dot FMAs, 4 skip chain+verify
};

enum { HB_ABORT_TIMEOUT_STREAM = 1, HB_ABORT_TIMEOUT_SCALAR = 2, HB_ABORT_TIMEOUT_TMA = 3, HB_ABORT_OVERFLOW = 4,
       HB_ABORT_TIMEOUT_PIPE = 5 };

// prm layout (SoA over m_pad): [0] u, [1] z, then for k = 1..F-1: a_k, c_k, 1/v_k, sd_k*z
__host__ __device__ __forceinline__ size_t prm_idx(int field, size_t m_pad, int j) { return (size_t)field * m_pad + j; }
constexpr int kPrmFields = 2 + 4 * (HB_MAX_FOLD - 1);

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
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        if (!mbar_wait(full + st, st_par, ctrl, HB_ABORT_TIMEOUT_TMA)) return;
        if (!(p.dbg & 2)) {
          const uint8_t* sp = stage0 + (size_t)st * p.stage_bytes + lane_off;
#pragma unroll
          for (int wd = 0; wd < RL / 8; ++wd) {
            uint2 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = *(const uint2*)(sp + k * R + 8 * wd);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              double a = acc[4 * q + k];
              a = fma(rs[8 * wd + 0], byte_as_scaled(v[k].x, 0x4044), a);
              a = fma(rs[8 * wd + 1], byte_as_scaled(v[k].x, 0x4144), a);
              a = fma(rs[8 * wd + 2], byte_as_scaled(v[k].x, 0x4244), a);
              a = fma(rs[8 * wd + 3], byte_as_scaled(v[k].x, 0x4344), a);
              a = fma(rs[8 * wd + 4], byte_as_scaled(v[k].y, 0x4044), a);
              a = fma(rs[8 * wd + 5], byte_as_scaled(v[k].y, 0x4144), a);
              a = fma(rs[8 * wd + 6], byte_as_scaled(v[k].y, 0x4244), a);
              a = fma(rs[8 * wd + 7], byte_as_scaled(v[k].y, 0x4344), a);
              acc[4 * q + k] = a;
            }
          }
        }
        __syncwarp();
        if (lane == 0) hb::mbar_arrive(empty + st);
        if (++st == NS) { st = 0; st_par ^= 1u; }
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
        // accumulator l = 4*q + k  <->  column 64q' ... of the tile: sub-stage q, column 8w + 4h + k of it
        const int col = (l >> 2) * SUBB + 8 * warp + 4 * h + (l & 3);
        const double scaled = (acc[0] * kTwo513) * p.dscale;
        if (!(fabs(scaled) < 4.0e18)) atomicCAS(ctrl + 1, 0, HB_ABORT_OVERFLOW);
        const long long fx = __double2ll_rn(scaled);
        atomicAdd(p.dacc + (size_t)t * B + col, (unsigned long long)fx);
      }
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
        // residual updates published by the scalar CTA for tile t-D
        const int need = t - D + 1;
        int ok = 1;
        if (lane == 0) {
          if (hb::ld_acquire(ctrl) < need) {
            Waiter w;
            while (hb::ld_acquire(ctrl) < need)
              if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_STREAM)) { ok = 0; break; }
          }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return;
        const int qend = __ldcg(p.tile_qend + (t - D));
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
  double sv[NF];
  const double rr = rhs * rhs;
  sv[0] = logpi0;
  double smax = logpi0;
#pragma unroll
  for (int k = 1; k < NF; ++k)
    if (k < nf) {
      sv[k] = fma(rr, q.c[k - 1], q.a[k - 1]);
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
  cls = 0;
  bool found = false;
#pragma unroll
  for (int k = 0; k < NF; ++k)
    if (k < nf && !found) {
      acc += sv[k] * inv;
      if (q.u < acc) { cls = k; found = true; }
    }
  gnew = 0.0;
#pragma unroll
  for (int k = 1; k < NF; ++k)
    if (k == cls) gnew = fma(rhs, q.iv[k - 1], q.sdz[k - 1]);
}

// candidate arrays of one thread group (B entries each)
struct CandSet {
  double *rhs0, *iv, *sdz, *gold, *delta, *gnew;
  int *idx, *cls;
};

// Shared memory of the scalar CTA.  Each thread group owns two row buffers into which the Gram rows of
// its tile's candidates are gathered by TMA bulk copies (one 4B-byte row per candidate and band block).
__host__ __device__ inline size_t scalar_fixed_bytes(int B, int D) {
  size_t b = (2 * (size_t)D * B + 12 * (size_t)B) * 8 + (4 * (size_t)B + 64 + 16) * 4 + 32 + 18 * 8;   // ring, candidates, ints, barriers, timers
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

template <int NF>
__device__ void scalar_role(const SweepParams& p, uint8_t* smem) {
  const int B = p.B, D = p.D, T = p.T, F = p.F, model = p.model;
  const int tid = threadIdx.x;
  const int ngrp = (p.dbg & 8) ? 1 : 2;   // timing experiment: one thread group does every tile
  if (p.dbg & 48) {
    // timing experiments: the streaming side alone.  16: every tile is published (without changes) as soon
    // as its dots have arrived; 32: all tiles are published up front.
    if (tid == 0) {
      if (p.dbg & 32) { for (int t = 0; t < T; ++t) p.tile_qend[t] = 0; __threadfence(); hb::st_release(p.ctrl, T); }
      else for (int t = 0; t < T; ++t) {
        Waiter w;
        while (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target)
          if (!w.keep_waiting(p.ctrl, HB_ABORT_TIMEOUT_SCALAR)) return;
        p.tile_qend[t] = 0;
        __threadfence();
        hb::st_release(p.ctrl, t + 1);
      }
      p.out->n_changed = 0; p.out->rounds = 0;
    }
    return;
  }
  if (tid >= ngrp * B) return;
  const int grp = tid / B, i = tid - grp * B, warp = i >> 5, lane = i & 31, nwarp = B / 32;
  // ---- shared memory carve-up
  double* ring;      // [2 groups][D][B]  corrections owed to the tiles in flight, one array per writing group
  CandSet cs;        // this group's candidate arrays
  int *wcnt, *wbad;  // 16 per group
  volatile int* ctl; // [0] tiles final (+ next-tile corrections in place), [1] qbase, [2],[3] per-group abort, [4] rounds, [5] tiles published
  uint64_t* rbar;    // this group's two row-buffer barriers
  int32_t *rows0, *rows1;
  long long* phase;
  {
    double* d = (double*)smem;
    ring = d; d += 2 * (size_t)D * B;
    double* cbase = d + (size_t)grp * 6 * B; d += 12 * (size_t)B;
    cs.rhs0 = cbase; cs.iv = cbase + B; cs.sdz = cbase + 2 * B; cs.gold = cbase + 3 * B;
    cs.delta = cbase + 4 * B; cs.gnew = cbase + 5 * B;
    int* ip = (int*)d;
    cs.idx = ip + (size_t)grp * 2 * B; cs.cls = cs.idx + B; ip += 4 * (size_t)B;
    wcnt = ip + grp * 16; wbad = ip + 32 + grp * 16; ip += 64;
    ctl = ip; ip += 16;
    rbar = (uint64_t*)ip + 2 * grp;
    phase = (long long*)((uint64_t*)ip + 4);
    uint8_t* rb = smem + scalar_fixed_bytes(B, D);
    rows0 = (int32_t*)(rb + (size_t)(2 * grp) * p.rowbuf);
    rows1 = (int32_t*)(rb + (size_t)(2 * grp + 1) * p.rowbuf);
  }
  const int KROW = (int)(p.rowbuf / ((size_t)B * 4));
  double* ring_mine = ring + (size_t)grp * D * B;
  int* ctrl = p.ctrl;
  const int gbar = 2 + grp;
  for (int d = 0; d < D; ++d) ring_mine[(size_t)d * B + i] = 0.0;
  if (tid < 8) ctl[tid] = 0;
  if (i == 0) { hb::mbar_init(rbar, 1); hb::mbar_init(rbar + 1, 1); hb::mbar_fence_init(); }
  hb::named_bar_sync(1, ngrp * B);

  const size_t mp = p.m_pad;
  const bool dense = (model == HB_MODEL_RR || model == HB_MODEL_A || model == HB_MODEL_L);
  const int nf = (model == HB_MODEL_R) ? F : 2;
  const int NONE = 1 << 30;
  unsigned n0 = 0, n1 = 0;   // gathers issued so far into rows0 / rows1 (parity of the phase to wait for)
  bool dead = false;

  long long* pc = phase + grp * 9;   // [8] = last time stamp
  if (i == 0) { for (int k = 0; k < 8; ++k) pc[k] = 0; pc[8] = clock64(); }
#define HB_PHASE(n) do { if (i == 0) { const long long _now = clock64(); pc[n] += _now - pc[8]; pc[8] = _now; } } while (0)
  for (int t = grp; t < T; t += ngrp) {
    const int j = t * B + i;
    const int slot = t % D;
    // ---- phase 1 (overlaps the other group's serial phase): inputs, speculation, Gram-row gather
    const bool act = (j < p.m) && p.active[j];
    const double xx = p.xpx[j];
    const double gold = p.g[j];
    SnpPrm<NF> q;
    q.u = p.prm[prm_idx(0, mp, j)];
#pragma unroll
    for (int k = 0; k < NF - 1; ++k) {
      q.a[k] = 0; q.c[k] = 0; q.iv[k] = 0; q.sdz[k] = 0;
      if (k < nf - 1) {
        q.a[k] = p.prm[prm_idx(2 + 4 * k, mp, j)];
        q.c[k] = p.prm[prm_idx(3 + 4 * k, mp, j)];
        q.iv[k] = p.prm[prm_idx(4 + 4 * k, mp, j)];
        q.sdz[k] = p.prm[prm_idx(5 + 4 * k, mp, j)];
      }
    }
    if (i == 0) {
      bool ok = true;
      if (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target) {
        Waiter w;
        while (hb::ld_relaxed_u(p.arrive + t) < p.arrive_target)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
      }
      hb::fence_acq_rel_gpu();
      ctl[2 + grp] = (!ok || *((volatile int*)(ctrl + 1)) != 0) ? 1 : 0;
    }
    hb::named_bar_sync(gbar, B);
    if (ctl[2 + grp]) break;
    HB_PHASE(0);
    const double base0 = (double)(long long)__ldcg(p.dacc + j) * p.inv_dscale;
    // rhs = x_j' yadj (+ xpx_j g_j)    (Bayes.cpp:593-594, 756-757)
    const double addback = (act && (dense || gold != 0.0)) ? xx * gold : 0.0;
    int cls = 0;
    double gnew = gold;
    if (act) {
      // the corrections of the previous tile may still be missing here: this is only the speculation
      const double rguess = ((base0 - ring[(size_t)slot * B + i]) - ring[(size_t)(D + slot) * B + i]) + addback;
      eval_snp<NF>(model, nf, rguess, q, p.logpi0, cls, gnew);
    }
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
#pragma unroll
        for (int kk = 1; kk < NF; ++kk)
          if (kk == cls) { iv = q.iv[kk - 1]; sdz = q.sdz[kk - 1]; }
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
    // ---- wait until the previous tile is final and its corrections for this tile are in the ring
    if (i == 0) {
      bool ok = true;
      if (ctl[0] < t) {
        Waiter w;
        while (ctl[0] < t)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
      }
      __threadfence_block();
      ctl[2 + grp] = (!ok || *((volatile int*)(ctrl + 1)) != 0) ? 1 : 0;
    }
    hb::named_bar_sync(gbar, B);
    if (ctl[2 + grp]) break;
    HB_PHASE(2);
    // ---- serial phase
    const double rhs0 = ((base0 - ring[(size_t)slot * B + i]) - ring[(size_t)(D + slot) * B + i]) + addback;
    ring[(size_t)slot * B + i] = 0.0;
    ring[(size_t)(D + slot) * B + i] = 0.0;
    int nrounds = 0;
    bool rows1_pending = false;
    for (;;) {
      ++nrounds;
      if (cand) cs.rhs0[myrank] = rhs0;
      hb::named_bar_sync(gbar, B);
      if (fast && !hbk::mbar_wait(rbar, (n0 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
      rows1_pending = fast && has1;
      HB_PHASE(3);
      if (warp == 0 && k > 0 && !dead) {
        if (fast) solve_candidates<true>(cs, k, G0, rows0, B, model, lane);
        else solve_candidates<false>(cs, k, G0, rows0, B, model, lane);
      }
      hb::named_bar_sync(gbar, B);
      HB_PHASE(4);
      // exact right-hand side of every SNP of the tile:  x_i'(r - sum_{c<i} x_c delta_c), ascending c
      double rhs = rhs0;
      if (fast) {
        if (!dead) {
#pragma unroll 4
          for (int sidx = 0; sidx < myrank; ++sidx) rhs = fma(-(double)rows0[(size_t)sidx * B + i], cs.delta[sidx], rhs);
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
      }
      int cls2 = 0;
      double gnew2 = gold;
      if (act) eval_snp<NF>(model, nf, rhs, q, p.logpi0, cls2, gnew2);
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
    // ---- the tile is final.  First what the next tile waits for: its corrections (dt = 1)
    if (has1) {
      double corr;
      if (fast) {
        if (!hbk::mbar_wait(rbar + 1, (n1 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
        corr = dead ? 0.0 : band_correction_rows(cs, k, rows1, B, i);
      } else {
        corr = band_correction(cs, k, G0 + (size_t)B * B, B, i);
      }
      ring_mine[(size_t)((t + 1) % D) * B + i] += corr;
    }
    const int qbase = ctl[1];
    hb::named_bar_sync(gbar, B);
    if (i == 0) {
      ctl[1] = qbase + k;
      ctl[4] = ctl[4] + nrounds;
      __threadfence_block();
      ctl[0] = t + 1;          // hand over to the other group
    }
    HB_PHASE(6);
    // ---- commit (overlaps the other group's serial phase)
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
      if (fast) {
        if (warp == 0) {
          issue_gather(rows0, G0 + (size_t)dt * B * B, cs, k, B, rbar, lane);
          if (two) issue_gather(rows1, G0 + (size_t)(dt + 1) * B * B, cs, k, B, rbar + 1, lane);
        }
        ++n0;
        if (two) ++n1;
        if (!hbk::mbar_wait(rbar, (n0 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
        if (!dead) ring_mine[(size_t)((t + dt) % D) * B + i] += band_correction_rows(cs, k, rows0, B, i);
        if (two) {
          if (!hbk::mbar_wait(rbar + 1, (n1 - 1) & 1, ctrl, HB_ABORT_TIMEOUT_SCALAR)) dead = true;
          if (!dead) ring_mine[(size_t)((t + dt + 1) % D) * B + i] += band_correction_rows(cs, k, rows1, B, i);
        }
        if (dt + 2 < D) hb::named_bar_sync(gbar, B);   // every thread is done with the row buffers before they are refilled
      } else {
        ring_mine[(size_t)((t + dt) % D) * B + i] += band_correction(cs, k, G0 + (size_t)dt * B * B, B, i);
        if (two) ring_mine[(size_t)((t + dt + 1) % D) * B + i] += band_correction(cs, k, G0 + (size_t)(dt + 1) * B * B, B, i);
      }
    }
    hb::named_bar_sync(gbar, B);
    if (i == 0) {
      // publish to the streaming CTAs, in tile order
      bool ok = true;
      if (ctl[5] < t) {
        Waiter w;
        while (ctl[5] < t)
          if (!w.keep_waiting(ctrl, HB_ABORT_TIMEOUT_SCALAR)) { ok = false; break; }
      }
      if (ok) {
        p.tile_qend[t] = qbase + k;
        __threadfence();
        hb::st_release(ctrl, t + 1);
        ctl[5] = t + 1;
      }
    }
    HB_PHASE(7);
  }
  if (dead) atomicCAS(ctrl + 1, 0, HB_ABORT_TIMEOUT_SCALAR);
  if (i == 0)
    for (int k = 0; k < 8; ++k) p.out->phase_clk[grp][k] = pc[k];
  hb::named_bar_sync(1, ngrp * B);
  if (tid == 0) {
    p.out->n_changed = ctl[1];
    p.out->rounds = ctl[4];
  }
}

}  // namespace hbk

template <int MAXT, int NF, int RL>
__global__ void __launch_bounds__(MAXT, 1) k_sweep(const __grid_constant__ SweepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  if ((int)blockIdx.x >= p.S) hbk::scalar_role<NF>(p, smem);
  else hbk::stream_role<RL>(p, smem);
}
