// sbayes.cu -- device engine for the summary-statistics sweeps of SBayesD() (dense LD,
// /root/reference/src/SBayesD.cpp:253-456) and SBayesS() (sparse LD, src/SBayesS.cpp:279-525).
//
// The reference walks the SNPs in order: rhs = r_hat[i] (+ xpx_i g_i), the same conditional draw as Bayes(), and,
// only if the effect changed, r_hat += (g_old - g_new) * n * LD[:, i] -- a length-m daxpy on column i of the dense
// matrix (SBayesD.cpp:351-356) or a walk over the stored entries of column i of the sp_mat (SBayesS.cpp:292-296,
// 403-407).  Here one cooperative launch does a whole sweep, tile by tile (B = 256 SNPs), with ONE grid barrier per
// tile:
//   CTA 0, step t       first brings the rows of tile t up to date with the changes of tile t-1 (the 256 x 256 LD block
//                       between the two tiles), then takes the tile's decisions: the dependence of r_hat[i] on the
//                       earlier changes of the same tile is n * LD[i, c] -- the diagonal LD block plays the role the
//                       Gram block plays in the genotype sweep (hb_sweep.cuh): classes are speculated, the changed
//                       SNPs are chained by one warp in SNP order, every SNP is re-evaluated with its exact right-hand
//                       side, a mismatch restarts the round.  The changed SNPs go to a double-buffered list.
//   other CTAs, step t  the column updates of tile t-1's changed SNPs on every other row of r_hat, in SNP order
//                       (deterministic, no atomics) -- overlapped with CTA 0's decisions for tile t.
// Storage of LD on the device:
//   dense (SBayesD)     the m x m fp64 matrix as given; a column update streams m * 8 bytes, coalesced over rows.
//   CSC (SBayesS)       colptr / rowidx / val as given (8 + 4 bytes per stored entry, nothing for the zeros), plus, built
//                       on the device at load, the dense 256 x 256 blocks CTA 0 needs (diagonal and first sub-diagonal
//                       block of every tile: 4 KB per SNP) and, for every column, where each worker's row range starts
//                       in it.  A worker owns a contiguous row range and walks the stored entries of a changed column
//                       that fall into it -- the reference's `for (it = ldm.begin_col(i); ...) r_hat[it.row()] += ...`.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/hibayes_b200.h"
#include "hb_rng.h"

namespace cg = cooperative_groups;
int hb_set_error(const char* fmt, ...);
#define CU(call)                                                                             \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                          cudaGetErrorString(_e));                                           \
  } while (0)

constexpr int LB = 256;   // SNPs per tile = threads per CTA

struct LdOutDev {
  double count[HB_MAX_FOLD];
  double varg_acc, sum_vargL, d_minus, d_plus;
  int n_changed, rounds;
  unsigned long long entries;   // LD entries streamed by the column updates (CSC: stored entries walked)
};

struct LdParams {
  const double* ldm;            // dense storage: m x m column-major (nullptr with CSC storage)
  // CSC storage (SBayesS)
  const int *colptr, *rowidx;
  const double* val;
  const double *blk0, *blk1;    // [T][LB][LB]: blk0[t][c][r] = LD[LB t + r, LB t + c], blk1[t][c][r] = LD[LB (t+1) + r, LB t + c]
  const int* rstart;            // [m][W + 1]: first stored entry of column c whose row is >= w * chunk
  int chunk, W;                 // rows per worker range, number of workers
  int m;
  double nscale;
  double *r_hat, *g, *vargL;
  const double *xpx, *xy;
  const uint8_t* ifest;
  int32_t* tracker;
  int iter, model, F;
  double logpi[HB_MAX_FOLD], vara_fold[HB_MAX_FOLD], fold[HB_MAX_FOLD];
  double vare, dfvara, s2varg, lambda, lambda2;
  hb_key_t key;
  // SBayesS: per-SNP residual variance varei = varediff_j * vara + vare, re-draws of too large effects for BayesC/Cpi and
  // BayesR (SBayesS.cpp:131-141, 285, 388-398)
  int sparse;
  const double *varediff, *vx;
  double vara, vary;
  uint8_t* looped;   // [m] the re-draw loop ran for this SNP in this sweep
  double* last2;     // [m] square of its last re-draw
  int* q_idx;        // [2][LB] global SNP index of a tile's changed SNPs (lists of consecutive tiles alternate)
  double* q_dn;      // [2][LB] (g_old - g_new) * n
  int* q_cnt;        // [2]
  LdOutDev* out;
};

// cumulative class probabilities (soft-max form of SBayesD.cpp:403-415) and the inverse-CDF class (:419-425)
__device__ __forceinline__ int ld_class(int nf, double rr, const double* a, const double* c, double logpi0, double u) {
  double sv[HB_MAX_FOLD];
  sv[0] = logpi0;
  double smax = logpi0;
  for (int k = 1; k < nf; ++k) { sv[k] = fma(rr, c[k], a[k]); smax = fmax(smax, sv[k]); }
  double tot = 0.0;
  for (int k = 0; k < nf; ++k) { sv[k] = exp(sv[k] - smax); tot += sv[k]; }
  const double inv = 1.0 / tot;
  double acc = 0.0;
  for (int k = 0; k < nf; ++k) { acc = fma(sv[k], inv, acc); if (u < acc) return k; }
  return 0;
}

__device__ double block_sum_256(double v, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// CSC: the kernel is compiled for both storages; `CSC` selects where an LD entry comes from
template <bool CSC>
__global__ void __launch_bounds__(LB) k_ld_sweep(const __grid_constant__ LdParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double c_rhs0[LB], c_iv[LB], c_sdz[LB], c_gold[LB], c_delta[LB], c_gnew[LB], red[LB], c_sd[LB], c_vx[LB], c_l2[LB];
  __shared__ int c_idx[LB], c_cls[LB], c_loop[LB], wcnt[LB / 32], s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = p.m, model = p.model, F = p.F;
  const bool dense = (model == HB_MODEL_RR || model == HB_MODEL_A || model == HB_MODEL_L);
  const int nf = (model == HB_MODEL_R) ? F : 2;
  const int T = (m + LB - 1) / LB;
  const double nscale = p.nscale;
  const bool redraw = p.sparse && (model == HB_MODEL_C || model == HB_MODEL_R);
  const bool solo = gridDim.x == 1;
  int rounds = 0, changed = 0;
  unsigned long long entries = 0;
  // entry (row LB t + rl, column LB t + cl) of the diagonal block of tile t
  auto diag = [&](int t, int cl, int rl) -> double {
    if (CSC) return p.blk0[((size_t)t * LB + cl) * LB + rl];
    return p.ldm[((size_t)t * LB + cl) * m + (size_t)t * LB + rl];
  };
  // the column updates of list `lt` (the changed SNPs of tile lt) on the rows of worker w, except the rows of the tiles
  // lt and lt + 1, which CTA 0 updates itself (right after the decisions of lt / before those of lt + 1)
  auto apply_columns = [&](int lt, int w) {
    const int buf = lt & 1;
    const int ku = *(volatile int*)(p.q_cnt + buf);
    const int* qi = p.q_idx + buf * LB;
    const double* qd = p.q_dn + buf * LB;
    const int s0 = lt * LB, s1 = s0 + 2 * LB;
    if (!CSC) {
      for (int row = w * LB + tid; row < m; row += p.W * LB) {
        if (row >= s0 && row < s1) continue;
        double acc = __ldcg(p.r_hat + row);
        // sixteen LD entries in flight per row (a tile changes up to 256 SNPs: four at a time was 64 trips to memory)
        for (int sq = 0; sq < ku; sq += 16) {
          double lv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) lv[q] = (sq + q < ku) ? p.ldm[(size_t)__ldcg(qi + sq + q) * m + row] : 0.0;
#pragma unroll
          for (int q = 0; q < 16; ++q) if (sq + q < ku) acc = fma(__ldcg(qd + sq + q), lv[q], acc);
        }
        p.r_hat[row] = acc;
      }
      if (tid == 0) entries += (unsigned long long)ku * (unsigned long long)((m - w * LB + p.W * LB - 1) / (p.W * LB)) * LB;
    } else {
      // stored entries of every changed column inside this worker's row range, one column after the other (SNP order)
      for (int s = 0; s < ku; ++s) {
        const int c = __ldcg(qi + s);
        const double dn = __ldcg(qd + s);
        const int p0 = p.rstart[(size_t)c * (p.W + 1) + w], p1 = p.rstart[(size_t)c * (p.W + 1) + w + 1];
        for (int q = p0 + tid; q < p1; q += LB) {
          const int row = p.rowidx[q];
          if (row >= s0 && row < s1) continue;
          p.r_hat[row] = fma(dn, p.val[q], __ldcg(p.r_hat + row));
        }
        if (tid == 0) entries += (unsigned long long)(p1 - p0);
        __syncthreads();   // the next column may hold the same rows
      }
    }
  };
  for (int t = 0; t <= T; ++t) {
    if (blockIdx.x == 0 && t < T) {
      const int j = t * LB + tid;
      // ---------------- the rows of this tile first: the changes of the previous tile (block between the two tiles)
      if (t > 0 && j < m) {
        const int buf = (t - 1) & 1;
        const int ku = *(volatile int*)(p.q_cnt + buf);
        double acc = __ldcg(p.r_hat + j);
        for (int sq = 0; sq < ku; sq += 16) {   // (on the serial path: sixteen entries in flight)
          double lv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int c = (sq + q < ku) ? __ldcg(p.q_idx + buf * LB + sq + q) : (t - 1) * LB;
            lv[q] = CSC ? p.blk1[((size_t)(t - 1) * LB + (c - (t - 1) * LB)) * LB + tid] : p.ldm[(size_t)c * m + j];
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) if (sq + q < ku) acc = fma(__ldcg(p.q_dn + buf * LB + sq + q), lv[q], acc);
        }
        p.r_hat[j] = acc;
        if (tid == 0) entries += (unsigned long long)ku * LB;
      }
      // ---------------- the tile's decisions
      const bool act = j < m && p.ifest[j];
      double xx = 0, gold = 0, rbase = 0, uu = 0.5, zz = 0, vare = p.vare, vxj = 0;
      double sd[HB_MAX_FOLD];
      for (int k = 0; k < HB_MAX_FOLD; ++k) sd[k] = 0;
      double a[HB_MAX_FOLD], c[HB_MAX_FOLD], iv[HB_MAX_FOLD], sdz[HB_MAX_FOLD];
      for (int k = 0; k < HB_MAX_FOLD; ++k) { a[k] = 0; c[k] = 0; iv[k] = 0; sdz[k] = 0; }
      if (act) {
        xx = p.xpx[j];
        gold = p.g[j];
        if (p.sparse) { vare = p.varediff[j] * p.vara + p.vare; vxj = p.vx[j]; }   // varei
        rbase = __ldcg(p.r_hat + j);
        if (gold != 0.0) rbase += xx * gold;   // :334-335 (every model)
        hb_draw_uz(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)j, HB_SL_MAIN, 0, &uu, &zz);
        if (model == HB_MODEL_R) {
          for (int k = 1; k < F; ++k) {
            const double vf = p.vara_fold[k], v = xx + vare / vf;
            a[k] = -0.5 * log(vf * (xx / vare) + 1.0) + p.logpi[k];
            c[k] = 0.5 / (vare * v);
            iv[k] = 1.0 / v;
            sd[k] = sqrt(vare / v);
            sdz[k] = sd[k] * zz;
          }
        } else {
          double varg = p.vara_fold[1];
          if (model == HB_MODEL_A || model == HB_MODEL_B)   // :277, :298
            varg = (gold * gold + p.s2varg * p.dfvara) /
                   hb_draw_chisq(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)j, HB_SL_CHI, p.dfvara + 1.0);
          const double v = (model == HB_MODEL_L) ? (xx + 1.0 / p.vargL[j]) : (xx + vare / varg);
          if (model == HB_MODEL_B || model == HB_MODEL_C) a[1] = -0.5 * log(varg * (xx / vare) + 1.0) + p.logpi[1];
          c[1] = 0.5 / (vare * v);
          iv[1] = 1.0 / v;
          sd[1] = sqrt(vare / v);
          sdz[1] = sd[1] * zz;
        }
      }
      auto classify = [&](double rhs) -> int { return dense ? 1 : ld_class(nf, rhs * rhs, a, c, p.logpi[0], uu); };
      int cls = act ? classify(rbase) : 0;
      int k = 0, myrank = 0;
      bool cand = false;
      double gnew = gold;
      for (;;) {
        ++rounds;
        // candidate list of the current classes
        cand = act && (dense || cls > 0 || gold != 0.0);
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int pre = 0;
        k = 0;
        for (int w = 0; w < LB / 32; ++w) { if (w < warp) pre += wcnt[w]; k += wcnt[w]; }
        myrank = pre + __popc(bal & ((1u << lane) - 1u));
        if (cand) {
          c_idx[myrank] = tid; c_gold[myrank] = gold; c_cls[myrank] = cls;
          c_iv[myrank] = iv[cls]; c_sdz[myrank] = sdz[cls]; c_rhs0[myrank] = rbase; c_sd[myrank] = sd[cls]; c_vx[myrank] = vxj;
        }
        __syncthreads();
        // chain of the candidates in SNP order (one warp): rhs_s = rhs0_s - sum_{s' < s} n LD[c_s, c_s'] delta_s'
        // blocked forward substitution: warp 0 chains 32 candidates, then all threads take those 32 changes out of the
        // right-hand sides of the later candidates (32 LD entries in flight per thread, same order of additions as a
        // candidate-by-candidate sweep) -- the one warp used to fetch them itself, eight at a time, for every block again
        for (int sb = 0; sb < k; sb += 32) {
          if (warp == 0) {
            const int sidx = sb + lane;
            const bool valid = sidx < k;
            const int li = valid ? c_idx[sidx] : 0;   // position inside the tile
            double rhs = valid ? c_rhs0[sidx] : 0.0;
            const double siv = valid ? c_iv[sidx] : 0.0, ssdz = valid ? c_sdz[sidx] : 0.0, sgold = valid ? c_gold[sidx] : 0.0;
            const int scls = valid ? c_cls[sidx] : 0;
            const double ssd = valid ? c_sd[sidx] : 0.0, svx = valid ? c_vx[sidx] : 0.0;
            const int nl = min(32, k - sb);
            // the LD entries towards the earlier candidates of this block, all loads in flight at once: the chain below
            // then waits for a shuffle and a fused multiply-add per step, not for a trip to L2
            double ldv[32];
#pragma unroll
            for (int lp = 0; lp < 32; ++lp) ldv[lp] = (valid && lp < nl && lane > lp) ? nscale * diag(t, c_idx[sb + lp], li) : 0.0;
            double mydelta = 0.0, mygnew = sgold, myl2 = 0.0;
            int myloop = 0;
#pragma unroll
            for (int lp = 0; lp < 32; ++lp) {
              if (lp >= nl) break;
              double gn = (scls > 0) ? fma(rhs, siv, ssdz) : 0.0;
              if (model == HB_MODEL_L && fabs(gn) < 1e-6) gn = 1e-6;   // :373
              if (redraw && lane == lp && scls > 0 && gn * gn * svx > p.vary) {   // SBayesS.cpp:388-398, 489-499
                int ii = 0;
                double l2 = 0.0;
                while (gn * gn * svx > p.vary) {
                  gn = fma(rhs, siv, ssd * hb_draw_z(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)(t * LB + li), HB_SL_RETRY, (uint32_t)(ii + 1)));
                  l2 = gn * gn;
                  ++ii;
                  if (ii > 100) gn = 0.0;
                }
                myloop = 1;
                myl2 = l2;
              }
              const double dl = gn - sgold;
              const double d = __shfl_sync(0xffffffffu, dl, lp);
              if (lane == lp) { mydelta = dl; mygnew = gn; }
              rhs = fma(-ldv[lp], d, rhs);
            }
            // (the re-draw flags belong to this round only: a round that is redone must not leave its flag behind)
            if (valid) { c_delta[sidx] = mydelta; c_gnew[sidx] = mygnew; c_loop[sidx] = myloop; c_l2[sidx] = myl2; }
            __syncwarp();
          }
          __syncthreads();
          if (sb + 32 < k) {
            const int sidx = sb + 32 + tid;   // (k <= LB: at most one later candidate per thread)
            if (sidx < k) {
              const int li = c_idx[sidx];
              double r = c_rhs0[sidx];
              double lq[32];
#pragma unroll
              for (int lp = 0; lp < 32; ++lp) lq[lp] = nscale * diag(t, c_idx[sb + lp], li);
#pragma unroll
              for (int lp = 0; lp < 32; ++lp) r = fma(-lq[lp], c_delta[sb + lp], r);
              c_rhs0[sidx] = r;
            }
            __syncthreads();
          }
        }
        // exact right-hand side of every SNP of the tile and its class (32 LD entries in flight)
        double rhs = rbase;
        if (act) {
          for (int sq = 0; sq < myrank; sq += 32) {
            double lq[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) lq[q] = (sq + q < myrank) ? nscale * diag(t, c_idx[sq + q], tid) : 0.0;
#pragma unroll
            for (int q = 0; q < 32; ++q) if (sq + q < myrank) rhs = fma(-lq[q], c_delta[sq + q], rhs);
          }
        }
        const int cls2 = act ? classify(rhs) : 0;
        if (tid == 0) s_flag = 0;
        __syncthreads();
        if (act && cls2 != cls) s_flag = 1;
        __syncthreads();
        const bool redo = s_flag != 0;
        __syncthreads();
        cls = cls2;
        if (!redo) break;
      }
      // commit the tile
      if (cand) gnew = c_gnew[myrank];
      if (act) {
        p.g[j] = gnew;
        p.tracker[j] = cls;
        if (redraw && cand && c_loop[myrank]) { p.looped[j] = 1; p.last2[j] = c_l2[myrank]; }
        if (model == HB_MODEL_L) {   // :374-375
          double u2, z2;
          hb_draw_uz(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)j, HB_SL_IG, 0, &u2, &z2);
          const double vargi = 1.0 / hb_invgauss_from_uz(sqrt(vare) * p.lambda / fabs(gnew), p.lambda2, u2, z2);   // varei in SBayesS
          if (vargi > 0) p.vargL[j] = vargi;
        }
      }
      // the reference updates r_hat only when the effect changed (:351; models 1 and 2 always, :263, :284).  The
      // changes inside the tile are applied to its own rows here, by their owners (diagonal block, SNP order).
      const bool upd = cand && (model == HB_MODEL_RR || model == HB_MODEL_A || gnew != gold);
      const unsigned balu = __ballot_sync(0xffffffffu, upd);
      if (lane == 0) wcnt[warp] = __popc(balu);
      __syncthreads();
      int pre = 0, ku = 0;
      for (int w = 0; w < LB / 32; ++w) { if (w < warp) pre += wcnt[w]; ku += wcnt[w]; }
      const int buf = t & 1;
      if (upd) {
        const int r = pre + __popc(balu & ((1u << lane) - 1u));
        p.q_idx[buf * LB + r] = j;
        p.q_dn[buf * LB + r] = (gold - gnew) * nscale;   // gi_ = (g[i] - gi) * n
        c_idx[r] = tid; c_delta[r] = (gold - gnew) * nscale;
      }
      if (tid == 0) p.q_cnt[buf] = ku;
      changed += ku;
      __syncthreads();
      if (j < m) {
        double acc = __ldcg(p.r_hat + j);
        for (int sq = 0; sq < ku; sq += 32) {
          double lq[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) lq[q] = (sq + q < ku) ? diag(t, c_idx[sq + q], tid) : 0.0;
#pragma unroll
          for (int q = 0; q < 32; ++q) if (sq + q < ku) acc = fma(c_delta[sq + q], lq[q], acc);
        }
        p.r_hat[j] = acc;
      }
      if (tid == 0) entries += (unsigned long long)ku * LB;
      __syncthreads();
      __threadfence();
    }
    // ---------------- the column updates of the previous tile's changes on all other rows, overlapped with the decisions
    if (t > 0) {
      // rows of tile t-1 itself were updated by CTA 0 right after its decisions, rows of tile t at the start of this step
      if (blockIdx.x > 0) apply_columns(t - 1, (int)blockIdx.x - 1);
      else if (solo) apply_columns(t - 1, 0);
    }
    grid.sync();
  }
  // ---------------- end of the sweep: the sums the host needs (:269, :353, :445-447, :460-467)
  if (tid == 0) atomicAdd(&p.out->entries, entries);
  if (blockIdx.x == 0) {
    double cnt[HB_MAX_FOLD], vacc = 0, sl = 0, dm = 0, dp = 0;
    for (int k = 0; k < HB_MAX_FOLD; ++k) cnt[k] = 0;
    for (int j = tid; j < m; j += LB) {
      const double gj = p.g[j], xyj = p.xy[j], rh = __ldcg(p.r_hat + j);
      dm += gj * (xyj - rh);
      dp += gj * (xyj + rh);
      if (model == HB_MODEL_L) sl += p.vargL[j];
      if (model == HB_MODEL_RR) vacc += gj * gj;
      if (p.ifest[j] && !dense) {
        const int cl = p.tracker[j];
        for (int k = 0; k < HB_MAX_FOLD; ++k) if (k == cl) cnt[k] += 1.0;
        if (cl > 0) vacc += (model == HB_MODEL_R) ? gj * gj / p.fold[cl] : gj * gj;
      }
    }
    if (p.sparse && model == HB_MODEL_C) {
      // SBayesS.cpp:392 overwrites the running sum of squares inside the re-draw loop: the sum restarts at the last
      // SNP whose loop ran, from the square of its last re-draw
      int lastj = -1;
      for (int j = tid; j < m; j += LB) if (p.looped[j]) lastj = j;
      __shared__ int s_last;
      if (tid == 0) s_last = -1;
      __syncthreads();
      atomicMax(&s_last, lastj);
      __syncthreads();
      lastj = s_last;
      if (lastj >= 0) {
        vacc = 0.0;
        for (int j = tid; j < m; j += LB)
          if (j >= lastj && p.ifest[j] && p.tracker[j] > 0) vacc += p.g[j] * p.g[j];
        if (tid == 0) vacc += p.last2[lastj];
      }
      __syncthreads();
    }
    for (int k = 0; k < HB_MAX_FOLD; ++k) { const double v = block_sum_256(cnt[k], red); if (tid == 0) p.out->count[k] = v; }
    vacc = block_sum_256(vacc, red); sl = block_sum_256(sl, red); dm = block_sum_256(dm, red); dp = block_sum_256(dp, red);
    if (tid == 0) {
      p.out->varg_acc = vacc; p.out->sum_vargL = sl; p.out->d_minus = dm; p.out->d_plus = dp;
      p.out->n_changed = changed; p.out->rounds = rounds;
    }
  }
}

// ------------------------------------------------------------------------------------------
// CSC helpers (device, at load)
// one warp per column: stored entries that fall into the tile's diagonal block or the block below it
__global__ void k_csc_blocks(const int* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ val, int m,
                             double* __restrict__ blk0, double* __restrict__ blk1) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= m) return;
  const int tc = c / LB, cl = c % LB;
  for (int q = colptr[c] + lane; q < colptr[c + 1]; q += 32) {
    const int r = rowidx[q], tr = r / LB;
    if (tr == tc) blk0[((size_t)tc * LB + cl) * LB + (r % LB)] = val[q];
    else if (tr == tc + 1) blk1[((size_t)tc * LB + cl) * LB + (r % LB)] = val[q];
  }
}
// rstart[c][w] = first stored entry of column c with row >= w * chunk (w = 0 .. W)
__global__ void k_csc_rstart(const int* __restrict__ colptr, const int* __restrict__ rowidx, int m, int W, int chunk, int* __restrict__ rstart) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)m * (W + 1)) return;
  const int c = (int)(idx / (W + 1)), w = (int)(idx % (W + 1));
  const long long target = (long long)w * chunk;
  int lo = colptr[c], hi = colptr[c + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowidx[mid] < target) lo = mid + 1; else hi = mid;
  }
  rstart[idx] = lo;
}

struct hb_ld_engine {
  int device = 0, m = 0, grid = 1, W = 1, chunk = 0;
  uint64_t seed = 0;
  double *ldm = nullptr, *r_hat = nullptr, *g = nullptr, *vargL = nullptr, *xpx = nullptr, *xy = nullptr, *q_dn = nullptr;
  int *colptr = nullptr, *rowidx = nullptr, *rstart = nullptr;
  double *val = nullptr, *blk0 = nullptr, *blk1 = nullptr;
  bool csc = false;
  uint64_t ld_bytes = 0;
  uint8_t *ifest = nullptr, *looped = nullptr;
  double *varediff = nullptr, *vx = nullptr, *last2 = nullptr;
  bool sparse_ready = false;
  int32_t* tracker = nullptr;
  int *q_idx = nullptr, *q_cnt = nullptr;
  LdOutDev* out = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float ms_sweep = 0;
  bool ld_ready = false, state_ready = false;
};

extern "C" int hb_ld_engine_create(int device, int m, uint64_t seed, hb_ld_engine** out) {
  if (!out || m <= 0) return hb_set_error("hb_ld_engine_create: bad argument");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return hb_set_error("hb_ld_engine_create: CUDA device %d not available (%d visible) -- no CPU fallback", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  hb_ld_engine* e = new hb_ld_engine();
  e->device = device; e->m = m; e->seed = seed;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ld_sweep<false>, LB, 0));
  int per_sm2 = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_ld_sweep<true>, LB, 0));
  per_sm = std::max(1, std::min(per_sm, per_sm2));
  // CTA 0 decides, the others update: one CTA per 256 rows is enough, and never more than are co-resident
  e->grid = std::max(1, std::min(prop.multiProcessorCount * per_sm, 1 + (m + LB - 1) / LB));
  e->W = std::max(1, e->grid - 1);
  e->chunk = (m + e->W - 1) / e->W;
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&e->ev[0])); CU(cudaEventCreate(&e->ev[1]));
  const size_t mm = (size_t)m;
  CU(cudaMalloc(&e->r_hat, mm * 8)); CU(cudaMalloc(&e->g, mm * 8)); CU(cudaMalloc(&e->vargL, mm * 8));
  CU(cudaMalloc(&e->xpx, mm * 8)); CU(cudaMalloc(&e->xy, mm * 8)); CU(cudaMalloc(&e->ifest, mm));
  CU(cudaMalloc(&e->tracker, mm * 4)); CU(cudaMalloc(&e->q_idx, 2 * LB * 4)); CU(cudaMalloc(&e->q_dn, 2 * LB * 8));
  CU(cudaMalloc(&e->q_cnt, 2 * 4)); CU(cudaMalloc(&e->out, sizeof(LdOutDev)));
  CU(cudaMemsetAsync(e->g, 0, mm * 8, e->stream)); CU(cudaMemsetAsync(e->tracker, 0, mm * 4, e->stream));
  CU(cudaMemsetAsync(e->vargL, 0, mm * 8, e->stream));
  CU(cudaMemsetAsync(e->q_cnt, 0, 2 * 4, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *out = e;
  return 0;
}
static void free_ld(hb_ld_engine* e) {
  cudaFree(e->ldm); cudaFree(e->colptr); cudaFree(e->rowidx); cudaFree(e->val); cudaFree(e->blk0); cudaFree(e->blk1); cudaFree(e->rstart);
  e->ldm = nullptr; e->colptr = nullptr; e->rowidx = nullptr; e->val = nullptr; e->blk0 = nullptr; e->blk1 = nullptr; e->rstart = nullptr;
  e->ld_ready = false;
}
extern "C" void hb_ld_engine_destroy(hb_ld_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  free_ld(e);
  cudaFree(e->r_hat); cudaFree(e->g); cudaFree(e->vargL); cudaFree(e->xpx); cudaFree(e->xy);
  cudaFree(e->looped); cudaFree(e->varediff); cudaFree(e->vx); cudaFree(e->last2);
  cudaFree(e->ifest); cudaFree(e->tracker); cudaFree(e->q_idx); cudaFree(e->q_dn); cudaFree(e->q_cnt); cudaFree(e->out);
  if (e->ev[0]) cudaEventDestroy(e->ev[0]);
  if (e->ev[1]) cudaEventDestroy(e->ev[1]);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}
extern "C" int hb_ld_engine_load_dense(hb_ld_engine* e, const double* ldm) {
  if (!e || !ldm) return hb_set_error("hb_ld_engine_load_dense: null argument");
  CU(cudaSetDevice(e->device));
  free_ld(e);
  const size_t bytes = (size_t)e->m * e->m * 8;
  if (cudaMalloc(&e->ldm, bytes) != cudaSuccess) {
    cudaGetLastError();
    return hb_set_error("hb_ld_engine_load_dense: the %d x %d fp64 LD matrix (%.1f GB) does not fit on the device", e->m, e->m, bytes / 1e9);
  }
  CU(cudaMemcpy(e->ldm, ldm, bytes, cudaMemcpyHostToDevice));
  e->csc = false; e->ld_bytes = bytes;
  e->ld_ready = true;
  return 0;
}
/* arma::sp_mat ldm of SBayesS() (SBayesS.cpp:21-40): compressed sparse columns, row indices ascending inside a column */
extern "C" int hb_ld_engine_load_csc(hb_ld_engine* e, const int32_t* colptr, const int32_t* rowidx, const double* val) {
  if (!e || !colptr || !rowidx || !val) return hb_set_error("hb_ld_engine_load_csc: null argument");
  const int m = e->m;
  if (colptr[0] != 0) return hb_set_error("hb_ld_engine_load_csc: colptr[0] must be 0");
  for (int c = 0; c < m; ++c) {
    if (colptr[c + 1] < colptr[c]) return hb_set_error("hb_ld_engine_load_csc: column pointers must not decrease");
    for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
      if (rowidx[q] < 0 || rowidx[q] >= m) return hb_set_error("hb_ld_engine_load_csc: row index %d out of range in column %d", rowidx[q], c);
      if (q > colptr[c] && rowidx[q] <= rowidx[q - 1]) return hb_set_error("hb_ld_engine_load_csc: row indices of column %d are not strictly ascending", c);
    }
  }
  CU(cudaSetDevice(e->device));
  free_ld(e);
  const size_t nnz = (size_t)colptr[m], T = (size_t)(m + LB - 1) / LB;
  CU(cudaMalloc(&e->colptr, ((size_t)m + 1) * 4)); CU(cudaMalloc(&e->rowidx, std::max<size_t>(nnz, 1) * 4));
  CU(cudaMalloc(&e->val, std::max<size_t>(nnz, 1) * 8));
  CU(cudaMalloc(&e->blk0, T * LB * LB * 8)); CU(cudaMalloc(&e->blk1, T * LB * LB * 8));
  CU(cudaMalloc(&e->rstart, (size_t)m * (e->W + 1) * 4));
  CU(cudaMemcpyAsync(e->colptr, colptr, ((size_t)m + 1) * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->rowidx, rowidx, nnz * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->val, val, nnz * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemsetAsync(e->blk0, 0, T * LB * LB * 8, e->stream)); CU(cudaMemsetAsync(e->blk1, 0, T * LB * LB * 8, e->stream));
  k_csc_blocks<<<(unsigned)(((size_t)m * 32 + 255) / 256), 256, 0, e->stream>>>(e->colptr, e->rowidx, e->val, m, e->blk0, e->blk1);
  CU(cudaGetLastError());
  const size_t nrs = (size_t)m * (e->W + 1);
  k_csc_rstart<<<(unsigned)((nrs + 255) / 256), 256, 0, e->stream>>>(e->colptr, e->rowidx, m, e->W, e->chunk, e->rstart);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  e->csc = true; e->ld_bytes = nnz * 12 + ((size_t)m + 1) * 4;
  e->ld_ready = true;
  return 0;
}
extern "C" int hb_ld_engine_describe(hb_ld_engine* e, int* csc, uint64_t* ld_bytes, int* grid) {
  if (!e) return hb_set_error("null engine");
  if (csc) *csc = e->csc ? 1 : 0;
  if (ld_bytes) *ld_bytes = e->ld_bytes;
  if (grid) *grid = e->grid;
  return 0;
}
extern "C" int hb_ld_engine_set_state(hb_ld_engine* e, const double* xpx, const uint8_t* ifest, const double* xy, const double* r_hat) {
  if (!e || !xpx || !ifest || !xy || !r_hat) return hb_set_error("hb_ld_engine_set_state: null argument");
  CU(cudaSetDevice(e->device));
  const size_t mm = (size_t)e->m;
  CU(cudaMemcpy(e->xpx, xpx, mm * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->ifest, ifest, mm, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->xy, xy, mm * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->r_hat, r_hat, mm * 8, cudaMemcpyHostToDevice));
  e->state_ready = true;
  return 0;
}
extern "C" int hb_ld_engine_set_sparse_info(hb_ld_engine* e, const double* varediff, const double* vx) {
  if (!e || !varediff || !vx) return hb_set_error("hb_ld_engine_set_sparse_info: null argument");
  CU(cudaSetDevice(e->device));
  const size_t mm = (size_t)e->m;
  if (!e->varediff) { CU(cudaMalloc(&e->varediff, mm * 8)); CU(cudaMalloc(&e->vx, mm * 8)); CU(cudaMalloc(&e->last2, mm * 8)); CU(cudaMalloc(&e->looped, mm)); }
  CU(cudaMemcpy(e->varediff, varediff, mm * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->vx, vx, mm * 8, cudaMemcpyHostToDevice));
  e->sparse_ready = true;
  return 0;
}
extern "C" int hb_ld_engine_set_vargL(hb_ld_engine* e, const double* v) {
  if (!e || !v) return hb_set_error("null argument");
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpy(e->vargL, v, (size_t)e->m * 8, cudaMemcpyHostToDevice));
  return 0;
}
extern "C" int hb_ld_engine_get(hb_ld_engine* e, double* g, int32_t* tracker, double* r_hat) {
  if (!e) return hb_set_error("null engine");
  CU(cudaSetDevice(e->device));
  if (g) CU(cudaMemcpy(g, e->g, (size_t)e->m * 8, cudaMemcpyDeviceToHost));
  if (tracker) CU(cudaMemcpy(tracker, e->tracker, (size_t)e->m * 4, cudaMemcpyDeviceToHost));
  if (r_hat) CU(cudaMemcpy(r_hat, e->r_hat, (size_t)e->m * 8, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int hb_ld_engine_sweep(hb_ld_engine* e, const hb_ld_sweep_in* in, hb_ld_sweep_out* out) {
  if (!e || !in || !out) return hb_set_error("hb_ld_engine_sweep: null argument");
  if (!e->ld_ready || !e->state_ready) return hb_set_error("hb_ld_engine_sweep: LD matrix or state not loaded");
  if (in->model_index < 1 || in->model_index > 6 || in->n_fold < 2 || in->n_fold > HB_MAX_FOLD) return hb_set_error("hb_ld_engine_sweep: bad model or n_fold");
  CU(cudaSetDevice(e->device));
  LdParams p;
  memset(&p, 0, sizeof p);
  p.ldm = e->ldm; p.colptr = e->colptr; p.rowidx = e->rowidx; p.val = e->val; p.blk0 = e->blk0; p.blk1 = e->blk1; p.rstart = e->rstart;
  p.chunk = e->chunk; p.W = e->W;
  p.m = e->m; p.nscale = in->nscale; p.r_hat = e->r_hat; p.g = e->g; p.vargL = e->vargL; p.xpx = e->xpx;
  p.xy = e->xy; p.ifest = e->ifest; p.tracker = e->tracker; p.iter = in->iter; p.model = in->model_index; p.F = in->n_fold;
  for (int k = 0; k < HB_MAX_FOLD; ++k) { p.logpi[k] = in->logpi[k]; p.vara_fold[k] = in->vara_fold[k]; p.fold[k] = in->fold[k]; }
  p.vare = in->vare; p.dfvara = in->dfvara; p.s2varg = in->s2varg; p.lambda = in->lambda; p.lambda2 = in->lambda2;
  p.sparse = in->sparse_mode ? 1 : 0; p.vara = in->vara; p.vary = in->vary;
  if (p.sparse) {
    if (!e->sparse_ready) return hb_set_error("hb_ld_engine_sweep: sparse_mode without hb_ld_engine_set_sparse_info");
    p.varediff = e->varediff; p.vx = e->vx; p.looped = e->looped; p.last2 = e->last2;
    CU(cudaMemsetAsync(e->looped, 0, (size_t)e->m, e->stream));
  }
  p.key = hb_make_key(e->seed); p.q_idx = e->q_idx; p.q_dn = e->q_dn; p.q_cnt = e->q_cnt; p.out = e->out;
  CU(cudaMemsetAsync(e->out, 0, sizeof(LdOutDev), e->stream));
  void* args[] = {(void*)&p};
  CU(cudaEventRecord(e->ev[0], e->stream));
  if (e->csc) CU(cudaLaunchCooperativeKernel((const void*)k_ld_sweep<true>, dim3(e->grid), dim3(LB), args, 0, e->stream));
  else CU(cudaLaunchCooperativeKernel((const void*)k_ld_sweep<false>, dim3(e->grid), dim3(LB), args, 0, e->stream));
  CU(cudaEventRecord(e->ev[1], e->stream));
  LdOutDev h;
  CU(cudaMemcpyAsync(&h, e->out, sizeof h, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->ms_sweep, e->ev[0], e->ev[1]));
  for (int k = 0; k < HB_MAX_FOLD; ++k) out->count[k] = h.count[k];
  out->varg_acc = h.varg_acc; out->sum_vargL = h.sum_vargL; out->g_xy_minus_rhat = h.d_minus; out->g_xy_plus_rhat = h.d_plus;
  out->n_changed = h.n_changed; out->status = 0; out->rounds = h.rounds; out->sweep_ms = e->ms_sweep;
  out->ld_entries = h.entries;
  return 0;
}
