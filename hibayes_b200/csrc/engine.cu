// engine.cu -- B200 (sm_100a) device engine for the hibayes single-site Gibbs sweep.
//
// Replaces the switch(model_index) block of Bayes() (/root/reference/src/Bayes.cpp:586-816)
// and the per-iteration reductions (:480,:819,:823) behind the C ABI of
// include/hibayes_b200.h.  See DESIGN.md for the algorithm; in short:
//
//   * X is held once in HBM as raw int8, tile-major  Xp[slab][tile][snp-in-tile][row-in-slab];
//     every streaming CTA owns one row slab and reads its part of X exactly once per sweep
//     through a TMA (cp.async.bulk) ring in shared memory.
//   * per tile of B SNPs each streaming CTA computes the B partial dots x_j'r over its slab
//     (PRMT + DFMA per genotype, the residual slab lives in registers) and adds them, as
//     fixed-point int64, to per-SNP accumulators in L2 -> the sum is order-independent.
//   * one scalar CTA runs the sequential Gibbs chain: it turns the dots of tile t into
//     conditional draws, resolving the dependence between SNPs with the precomputed exact
//     Gram band  G = X_t'[X_t .. X_{t+D-1}]  (speculate all, commit up to the first changed
//     SNP, patch the later right-hand sides, repeat) and publishes the changed effects.
//   * streaming CTAs apply those residual updates (r -= x_j*delta, u += x_j*delta) D tiles
//     later, so D tiles are in flight and the chain latency is hidden behind streaming.
//
// All random draws are position-addressed (hb_rng.h), so the result does not depend on the
// decomposition.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>
#include <thread>
#include <atomic>

#include "../../include/hibayes_b200.h"
#include "hb_bed.cuh"
#include "hb_device.cuh"
#include "hb_rng.h"
#include "hb_sweep.cuh"
#include "hb_synth.h"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
extern "C" const char* hb_last_error(void) { return g_err; }
int hb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return 1;
}
#define CU(call)                                                                             \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                          cudaGetErrorString(_e));                                           \
  } while (0)

extern "C" int hb_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return -1;
  return c;
}

// ------------------------------------------------------------------------------------------
// engine state
// ------------------------------------------------------------------------------------------
struct hb_engine {
  hb_engine_config cfg;
  int n, m, B, D, S, R, NRG, CL, T, m_pad, NS, nsm;
  int NTC, NTCp, NCW, NAW, SUBB, NG, RL, block_threads;
  size_t Npad, slab_stride, stage_bytes, smem_bytes;
  size_t off_rbuf, off_bar, off_qbuf;
  uint8_t* Xp = nullptr;
  double *r = nullptr, *u = nullptr, *xpx = nullptr, *g = nullptr, *vargL = nullptr, *gsum = nullptr;
  double *nzrate = nullptr, *wppa = nullptr;
  uint8_t* active = nullptr;
  int32_t* tracker = nullptr;
  int32_t* gram = nullptr;
  bool gram_ready = false, info_ready = false, geno_ready = false;
  unsigned long long* dacc = nullptr;
  int* q_snp = nullptr;
  double* q_delta = nullptr;
  int* tile_cnt = nullptr;
  double* corr = nullptr;
  double xpx_max = 0.0;
  unsigned long long* acc2 = nullptr;          // [2][m_pad] second-level dot accumulators (row-sharded runs), by sweep parity
  unsigned long long* peer_acc2[8] = {nullptr};  // every rank's acc2 (IPC mappings; [rank] = acc2)
  bool peers_ready = false;
  unsigned sweep_no = 0;
  unsigned long long* trace = nullptr;
  int KROW = 0;
  int lead = 0;       // HB_LEAD=1|2|4 (experiment): phase P speculates as if it ran that many tiles earlier
  int limbs = 0;      // HB_LIMBS=1 (experimental): integer-only dots in the streaming CTAs (k_sweep<..., LIMBS>)
  double* absmax_dev = nullptr;
  int cluster2 = 0;   // HB_CLUSTER=1: scalar workers in clusters of 2 (hand-over through distributed shared memory)
  // serial mode (hb_serial.cuh; mixture models): one serial CTA + NH helper CTAs instead of the ring of NG workers
  int serial = 0, NH = 0, KROW_S = 0;
  uint32_t pkg_stride = 0;
  uint8_t* pkg = nullptr;
  int* pkg_flag = nullptr;
  int* miss_tile = nullptr;
  int scalar0 = 0;    // block index of the first scalar CTA
  int* ctrl = nullptr;  // [0] progress, [1] abort
  double* prm = nullptr;  // per-SNP sweep parameters, SoA
  int prm_fold = 0;
  SweepOutDev* out_dev = nullptr;
  double* post_partial = nullptr;  // kPostBlocks x (HB_MAX_FOLD+1)
  double* fold_dev = nullptr;      // HB_MAX_FOLD
  // windows (CSR)
  int nw = 0;
  int *wstart = nullptr, *wmem = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms_prep = 0, ms_sweep = 0, ms_tail = 0, ms_predict = 0;
};


// ------------------------------------------------------------------------------------------
// packing, synthetic data, column statistics
// ------------------------------------------------------------------------------------------
// One thread per 16-row chunk of one column: gathers 16 int8 genotypes (0 beyond n) and stores
// them as one 16-byte vector at Xp[slab][tile][col][16*rg ..].
// bad (may be NULL): [0] = 1 + the first column seen with a value outside {0,1,2}, [1] = such a value
__global__ void k_pack_i8(const int8_t* __restrict__ src, size_t ld, int n, int col0, int ncols, uint8_t* __restrict__ Xp,
                          int S, int R, int NRG, int T, int B, int* __restrict__ bad = nullptr) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per_col = (size_t)S * NRG;
  if (idx >= per_col * ncols) return;
  int c = (int)(idx / per_col);
  int ch = (int)(idx % per_col);
  int s = ch / NRG, rg = ch % NRG;
  size_t row0 = (size_t)s * R + 16 * rg;
  const int8_t* col = src + (size_t)c * ld;
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    size_t row = row0 + i;
    uint32_t x = (row < (size_t)n) ? (uint32_t)(uint8_t)col[row] : 0u;
    w[i >> 2] |= x << (8 * (i & 3));
  }
  int j = col0 + c, t = j / B, cj = j % B;
  if (bad) {
    // a byte above 2 in any of the sixteen: (w & 0xfc...) catches 4 .. 255, the pair of low bits both set catches 3
    uint32_t o = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) o |= (w[q] & 0xfcfcfcfcu) | (w[q] & (w[q] >> 1) & 0x01010101u);
    if (o && atomicCAS(bad, 0, j + 1) == 0) {
      int v = 0;
      for (int i = 0; i < 16; ++i) { const int x = (int)(int8_t)((w[i >> 2] >> (8 * (i & 3))) & 0xffu); if ((uint8_t)x > 2 && !v) v = x; }
      bad[1] = v;
    }
  }
  uint4* dst = (uint4*)(Xp + ((((size_t)s * T + t) * B + cj) * R + 16 * rg));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

// The same layout straight from a PLINK .bed column (2 bits per genotype, hb_bed.cuh): one thread per
// 16-row chunk of one SNP; output row i is file individual rows[i] (or i), missing genotypes become the
// SNP's major genotype (read_bed.cpp:186-229).
__global__ void k_pack_bed(const uint8_t* __restrict__ bed, size_t bps, const int32_t* __restrict__ rows, int n, int col0,
                           int ncols, int d, const uint8_t* __restrict__ info, uint8_t* __restrict__ Xp, int S, int R, int NRG,
                           int T, int B) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per_col = (size_t)S * NRG;
  if (idx >= per_col * ncols) return;
  int c = (int)(idx / per_col);
  int ch = (int)(idx % per_col);
  int s = ch / NRG, rg = ch % NRG;
  size_t row0 = (size_t)s * R + 16 * rg;
  const uint8_t* src = bed + (size_t)c * bps;
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    size_t row = row0 + i;
    if (row < (size_t)n) w[i >> 2] |= (uint32_t)hb::bed_value(src, rows, row, d, 1, 0, info[c]) << (8 * (i & 3));
  }
  int j = col0 + c, t = j / B, cj = j % B;
  uint4* dst = (uint4*)(Xp + ((((size_t)s * T + t) * B + cj) * R + 16 * rg));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void k_synth(uint8_t* __restrict__ Xp, int n, int m, int S, int R, int NRG, int T, int B, hb_key_t key,
                        long long row_offset) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per_col = (size_t)S * NRG;
  if (idx >= per_col * (size_t)m) return;
  int j = (int)(idx / per_col);
  int ch = (int)(idx % per_col);
  int s = ch / NRG, rg = ch % NRG;
  size_t row0 = (size_t)s * R + 16 * rg;
  double t0, t1;
  hb_synth_thresholds(key, (uint32_t)j, &t0, &t1);
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    size_t lrow = row0 + 4 * k;
    uint32_t word = 0;
    if (lrow < (size_t)n) {
      word = hb_synth_word(key, (uint32_t)j, (uint64_t)(row_offset + (long long)lrow) >> 2, t0, t1);
      // mask rows beyond n inside the last word
      for (int b = 0; b < 4; ++b)
        if (lrow + b >= (size_t)n) word &= ~(0xffu << (8 * b));
    }
    w[k] = word;
  }
  int t = j / B, cj = j % B;
  uint4* dst = (uint4*)(Xp + ((((size_t)s * T + t) * B + cj) * R + 16 * rg));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

// Column sums: one warp per column; exact integer sums of x and x^2 (Bayes.cpp:310-315).
__global__ void k_col_stats(const uint8_t* __restrict__ Xp, int m, int S, int R, int NRG, int T, int B,
                            double* __restrict__ xpx, double* __restrict__ sumx) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= m) return;
  int j = warp, t = j / B, cj = j % B;
  long long s1 = 0, s2 = 0;
  int chunks = S * NRG;
  for (int ch = lane; ch < chunks; ch += 32) {
    int s = ch / NRG, rg = ch % NRG;
    uint4 v = *(const uint4*)(Xp + ((((size_t)s * T + t) * B + cj) * R + 16 * rg));
    unsigned a = 0, b = 0;
    a = __dp4a(v.x, 0x01010101u, a); b = __dp4a(v.x, v.x, b);
    a = __dp4a(v.y, 0x01010101u, a); b = __dp4a(v.y, v.y, b);
    a = __dp4a(v.z, 0x01010101u, a); b = __dp4a(v.z, v.z, b);
    a = __dp4a(v.w, 0x01010101u, a); b = __dp4a(v.w, v.w, b);
    s1 += a; s2 += b;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { sumx[j] = (double)s1; xpx[j] = (double)s2; }
}

// ------------------------------------------------------------------------------------------
// band Gram:  G[t][dt][a][b] = x_{tB+a}' x_{(t+dt)B+b}   (exact int32)
// v1: dp4a, 64x64 output block per CTA, rows staged 128 at a time.
// ------------------------------------------------------------------------------------------
constexpr int GK = 128;  // rows per staging step
__global__ void __launch_bounds__(256) k_gram_dp4a(const uint8_t* __restrict__ Xp, int32_t* __restrict__ gram, int S, int R,
                                                   int T, int B, int D, int t_base) {
  __shared__ uint32_t As[64][GK / 4 + 1];
  __shared__ uint32_t Bs[64][GK / 4 + 1];
  const int t = t_base + blockIdx.z;
  const int nb = B / 64;
  const int a0 = (blockIdx.y % nb) * 64;
  const int dt = blockIdx.x / nb, b0 = (blockIdx.x % nb) * 64;
  const int t2 = t + dt;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  unsigned acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][k] = 0u;
  if (t2 < T) {
    for (int s = 0; s < S; ++s) {
      const uint8_t* Abase = Xp + (((size_t)s * T + t) * B + a0) * R;
      const uint8_t* Bbase = Xp + (((size_t)s * T + t2) * B + b0) * R;
      for (int r0 = 0; r0 < R; r0 += GK) {
        const int nwords = min(GK, R - r0) / 4;  // R is a multiple of 16
        for (int e = tid; e < 64 * (GK / 4); e += 256) {
          int c = e / (GK / 4), w = e % (GK / 4);
          uint32_t va = 0, vb = 0;
          if (w < nwords) {
            va = *(const uint32_t*)(Abase + (size_t)c * R + r0 + 4 * w);
            vb = *(const uint32_t*)(Bbase + (size_t)c * R + r0 + 4 * w);
          }
          As[c][w] = va;
          Bs[c][w] = vb;
        }
        __syncthreads();
#pragma unroll 4
        for (int w = 0; w < GK / 4; ++w) {
          uint32_t av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][w];
#pragma unroll
          for (int k = 0; k < 4; ++k) bv[k] = Bs[tx + 16 * k][w];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][k] = __dp4a(av[i], bv[k], acc[i][k]);
        }
        __syncthreads();
      }
    }
  }
  int32_t* out = gram + (((size_t)t * D + dt) * B) * B;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) out[(size_t)(a0 + ty * 4 + i) * B + (b0 + tx + 16 * k)] = (int32_t)acc[i][k];
}

// v2: the same band Gram on the integer tensor-core path (mma.sync m16n8k32 u8 -> s32, exact).
// This is the one dense contraction of the code base (SURVEY.md K6 / tXXmat family); it runs once
// per data set.  CTA = 64x64 outputs of block (t, dt), 4 warps of 32x32; rows staged 128 at a
// time through a double-buffered cp.async pipeline; fragments come from ldmatrix.
constexpr int GI_RK = 128;            // rows (bytes) per stage
constexpr int GI_STRIDE = GI_RK + 16; // padded row stride: conflict-free ldmatrix
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = hb::smem_u32(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(hb::smem_u32(p)));
}
__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128) k_gram_imma(const uint8_t* __restrict__ Xp, int32_t* __restrict__ gram, int S, int R,
                                                   int T, int B, int D, int t_base) {
  __shared__ __align__(16) uint8_t As[2][64 * GI_STRIDE];
  __shared__ __align__(16) uint8_t Bs[2][64 * GI_STRIDE];
  const int t = t_base + blockIdx.z;
  const int nb = B / 64;
  const int a0 = (blockIdx.y % nb) * 64;
  const int dt = blockIdx.x / nb, b0 = (blockIdx.x % nb) * 64;
  const int t2 = t + dt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  int acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0;
  int32_t* out = gram + (((size_t)t * D + dt) * B) * B;
  if (t2 < T) {
    const int cps = (R + GI_RK - 1) / GI_RK;  // chunks per slab
    const int nchunks = S * cps;
    auto issue = [&](int ch, int buf) {
      const int s = ch / cps, r0 = (ch % cps) * GI_RK;
      const uint8_t* Abase = Xp + (((size_t)s * T + t) * B + a0) * R + r0;
      const uint8_t* Bbase = Xp + (((size_t)s * T + t2) * B + b0) * R + r0;
      // 64 columns x 8 vectors of 16 B per operand = 512 vectors; 128 threads -> 4 + 4 each
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = tid + 128 * q;
        const int c = v >> 3, w = v & 7;
        const bool ok = r0 + 16 * w < R;
        cp_async16(&As[buf][c * GI_STRIDE + 16 * w], Abase + (size_t)c * R + (ok ? 16 * w : 0), ok);
        cp_async16(&Bs[buf][c * GI_STRIDE + 16 * w], Bbase + (size_t)c * R + (ok ? 16 * w : 0), ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int buf = ch & 1;
      if (ch + 1 < nchunks) {
        issue(ch + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < GI_RK / 32; ++ks) {
        uint32_t af[2][4], bf[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = wm + 16 * i + (lane & 7) + 8 * ((lane >> 3) & 1);
          ldmatrix_x4(af[i], &As[buf][row * GI_STRIDE + 32 * ks + 16 * (lane >> 4)]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = wn + 16 * j + (lane & 7) + 8 * (lane >> 4);
          ldmatrix_x4(bf[j], &Bs[buf][row * GI_STRIDE + 32 * ks + 16 * ((lane >> 3) & 1)]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) imma_16832(acc[i][j], af[i], bf[j >> 1][2 * (j & 1)], bf[j >> 1][2 * (j & 1) + 1]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = a0 + wm + 16 * i + (lane >> 2);
      const int col = b0 + wn + 8 * j + 2 * (lane & 3);
      *(int2*)&out[(size_t)row * B + col] = make_int2(acc[i][j][0], acc[i][j][1]);
      *(int2*)&out[(size_t)(row + 8) * B + col] = make_int2(acc[i][j][2], acc[i][j][3]);
    }
}

// ------------------------------------------------------------------------------------------
// per-sweep preparation: everything about SNP j that does not depend on the residual
// ------------------------------------------------------------------------------------------
struct PrepParams {
  int m, m_pad, T, iter, model, F, use_thr;
  double fold[HB_MAX_FOLD], logpi[HB_MAX_FOLD], vara_fold[HB_MAX_FOLD];
  double vare, dfvara, s2varg;
  hb_key_t key;
};

__global__ void k_prep(PrepParams p, const double* __restrict__ xpx, const uint8_t* __restrict__ active,
                       const double* __restrict__ g, const double* __restrict__ vargL, double* __restrict__ prm,
                       unsigned long long* __restrict__ dacc, int* __restrict__ ctrl, SweepOutDev* __restrict__ out,
                       int* __restrict__ tile_cnt, int* __restrict__ q_snp, unsigned long long* __restrict__ q_delta,
                       unsigned long long* __restrict__ corr, int B, int DC, unsigned long long* __restrict__ acc2_next,
                       int* __restrict__ pkg_flag, int* __restrict__ miss_tile) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    ctrl[0] = 0; ctrl[1] = 0;
    if (miss_tile) *miss_tile = -(1 << 28);
    out->n_changed = 0; out->status = 0; out->rounds = 0; out->pad = 0; out->varg_acc = 0; out->sum_vargL = 0;
    for (int k = 0; k < HB_MAX_FOLD; ++k) out->count[k] = 0;
  }
  if (j < p.T) { tile_cnt[j] = -1; if (pkg_flag) pkg_flag[j] = 0; }
  if (j >= p.m_pad) return;
  dacc[j] = 0ull;
  if (acc2_next) acc2_next[j] = 0ull;   // the buffer of the sweep after this one (see hb_engine_sweep)
  q_snp[j] = -1;
  q_delta[j] = hbk::kCorrEmpty;
  {
    const int t = j / B, i = j - t * B;
    for (int d = 0; d < DC; ++d) corr[((size_t)t * DC + d) * B + i] = hbk::kCorrEmpty;
  }
  if (j >= p.m || !active[j]) return;
  const double xx = xpx[j];
  double u, z;
  hb_draw_uz(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)j, HB_SL_MAIN, 0, &u, &z);
  prm[prm_idx(0, p.m_pad, j)] = u;
  prm[prm_idx(1, p.m_pad, j)] = z;
  const double vare = p.vare;
  if (p.model == HB_MODEL_R) {
    for (int k = 1; k < p.F; ++k) {
      double vf = p.vara_fold[k];
      double v = xx + vare / vf;
      int f = 2 + 4 * (k - 1);
      prm[prm_idx(f + 0, p.m_pad, j)] = -0.5 * log(vf * (xx / vare) + 1.0) + p.logpi[k];
      prm[prm_idx(f + 1, p.m_pad, j)] = 0.5 / (vare * v);
      prm[prm_idx(f + 2, p.m_pad, j)] = 1.0 / v;
      prm[prm_idx(f + 3, p.m_pad, j)] = sqrt(vare / v) * z;
    }
    if (p.use_thr) {
      double a[HB_MAX_FOLD - 1], c[HB_MAX_FOLD - 1], TL[HB_MAX_FOLD - 1], TH[HB_MAX_FOLD - 1];
      for (int k = 1; k < p.F; ++k) { a[k - 1] = prm[prm_idx(2 + 4 * (k - 1), p.m_pad, j)]; c[k - 1] = prm[prm_idx(3 + 4 * (k - 1), p.m_pad, j)]; }
      hbk::solve_thresholds<HB_MAX_FOLD>(p.F, u, a, c, p.logpi[0], TL, TH);
      for (int b = 0; b < p.F - 1; ++b) {
        prm[prm_idx(kThrField0 + 2 * b, p.m_pad, j)] = TL[b];
        prm[prm_idx(kThrField0 + 2 * b + 1, p.m_pad, j)] = TH[b];
      }
    }
  } else {
    double varg = p.vara_fold[1];
    if (p.model == HB_MODEL_A || p.model == HB_MODEL_B) {
      double og = g[j];
      varg = (og * og + p.s2varg * p.dfvara) /
             hb_draw_chisq(p.key, HB_DOM_SNP, (uint32_t)p.iter, (uint32_t)j, HB_SL_CHI, p.dfvara + 1.0);
    }
    double v = (p.model == HB_MODEL_L) ? (xx + 1.0 / vargL[j]) : (xx + vare / varg);
    double a = 0.0;
    if (p.model == HB_MODEL_B || p.model == HB_MODEL_C) a = -0.5 * log(varg * (xx / vare) + 1.0) + p.logpi[1];
    prm[prm_idx(2, p.m_pad, j)] = a;
    prm[prm_idx(3, p.m_pad, j)] = 0.5 / (vare * v);
    prm[prm_idx(4, p.m_pad, j)] = 1.0 / v;
    prm[prm_idx(5, p.m_pad, j)] = sqrt(vare / v) * z;
    if (p.use_thr && (p.model == HB_MODEL_B || p.model == HB_MODEL_C)) {
      double a1[1] = {a}, c1[1] = {0.5 / (vare * v)}, TL[1], TH[1];
      hbk::solve_thresholds<2>(2, u, a1, c1, p.logpi[0], TL, TH);
      prm[prm_idx(kThrField0, p.m_pad, j)] = TL[0];
      prm[prm_idx(kThrField0 + 1, p.m_pad, j)] = TH[0];
    }
  }
}

// kernel variants: CTA size (register budget) x number of mixture classes held in registers
template <int RL>
static const void* sweep_kernel_rl(int nf, bool dense) {
  if (dense) return (const void*)k_sweep<512, 2, RL, true>;
  return nf <= 2 ? (const void*)k_sweep<512, 2, RL, false> : nf <= 4 ? (const void*)k_sweep<512, 4, RL, false> : (const void*)k_sweep<512, HB_MAX_FOLD, RL, false>;
}
// kernel variants: rows per lane x number of mixture classes held in registers x dense/mixture chain
static const void* sweep_kernel_for(int nf, int rl, bool dense) {
  return rl == 8 ? sweep_kernel_rl<8>(nf, dense) : rl == 16 ? sweep_kernel_rl<16>(nf, dense) : sweep_kernel_rl<24>(nf, dense);
}
// the mixture models with the serial CTA + helpers (hb_serial.cuh)
template <int RL>
static const void* serial_kernel_rl(int nf) {
  return nf <= 2 ? (const void*)k_sweep<512, 2, RL, false, false, 0, true>
       : nf <= 4 ? (const void*)k_sweep<512, 4, RL, false, false, 0, true> : (const void*)k_sweep<512, HB_MAX_FOLD, RL, false, false, 0, true>;
}
static const void* serial_kernel_for(int nf, int rl) {
  return rl == 8 ? serial_kernel_rl<8>(nf) : rl == 16 ? serial_kernel_rl<16>(nf) : serial_kernel_rl<24>(nf);
}

// ------------------------------------------------------------------------------------------
// per-iteration reductions (Bayes.cpp:480, 819, 823) -- one CTA, fixed reduction tree
// ------------------------------------------------------------------------------------------
__device__ double block_sum_1024(double v, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  double r = sh[0];
  __syncthreads();
  return r;
}
// Class counts and the variance accumulator of the sweep (Bayes.cpp:603, 698, 791, 803-805), taken
// from the committed effects in a fixed order (deterministic): stage 1 = per-block partials, stage 2 =
// one warp adds the partials in block order.
constexpr int kPostBlocks = 148, kPostThreads = 256;
__global__ void __launch_bounds__(kPostThreads) k_post1(int m, int model, const int32_t* __restrict__ tracker,
                                                        const double* __restrict__ g, const uint8_t* __restrict__ active,
                                                        const double* __restrict__ fold8, double* __restrict__ partial) {
  __shared__ double sh[kPostThreads][HB_MAX_FOLD + 1];
  double cnt[HB_MAX_FOLD + 1];
#pragma unroll
  for (int k = 0; k <= HB_MAX_FOLD; ++k) cnt[k] = 0.0;
  double fold[HB_MAX_FOLD];
#pragma unroll
  for (int k = 0; k < HB_MAX_FOLD; ++k) fold[k] = fold8[k];
  const bool dense = (model == HB_MODEL_RR || model == HB_MODEL_A || model == HB_MODEL_L);
  for (int j = blockIdx.x * kPostThreads + threadIdx.x; j < m; j += kPostBlocks * kPostThreads) {
    if (!active[j]) continue;
    const int cls = dense ? 1 : tracker[j];
    const double gj = g[j];
#pragma unroll
    for (int k = 0; k < HB_MAX_FOLD; ++k)
      if (k == cls) {
        cnt[k] += 1.0;
        if (k > 0) cnt[HB_MAX_FOLD] += (model == HB_MODEL_R) ? (gj * gj / fold[k]) : (gj * gj);
      }
  }
#pragma unroll
  for (int k = 0; k <= HB_MAX_FOLD; ++k) sh[threadIdx.x][k] = cnt[k];
  __syncthreads();
  if (threadIdx.x <= HB_MAX_FOLD) {
    double a = 0.0;
    for (int q = 0; q < kPostThreads; ++q) a += sh[q][threadIdx.x];
    partial[blockIdx.x * (HB_MAX_FOLD + 1) + threadIdx.x] = a;
  }
}
__global__ void k_post2(const double* __restrict__ partial, SweepOutDev* out) {
  const int k = threadIdx.x;
  if (k > HB_MAX_FOLD) return;
  double a = 0.0;
  for (int b = 0; b < kPostBlocks; ++b) a += partial[b * (HB_MAX_FOLD + 1) + k];
  if (k < HB_MAX_FOLD) out->count[k] = a; else out->varg_acc = a;
}

__global__ void __launch_bounds__(1024) k_tail(const double* __restrict__ r, const double* __restrict__ u, int n,
                                               SweepOutDev* out, const int* ctrl) {
  __shared__ double sh[1024];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < n; i += 1024) { double x = r[i]; a += x; b += x * x; c += u[i]; }
  a = block_sum_1024(a, sh); b = block_sum_1024(b, sh); c = block_sum_1024(c, sh);
  const double mean = c / n;
  double e = 0, f = 0;
  for (int i = threadIdx.x; i < n; i += 1024) { double x = mean - u[i]; e += x * x; f += x; }
  e = block_sum_1024(e, sh); f = block_sum_1024(f, sh);
  if (threadIdx.x == 0) {
    out->sum_r = a; out->sum_r2 = b; out->sum_u = c;
    out->var_u = (n > 1) ? (e - f * f / n) / (n - 1) : 0.0;  // op_var::direct_var
    out->status = ctrl[1];
  }
}

// BayesL: per-SNP variance draw after the sweep (Bayes.cpp:729-730) and sum(vargL) (:739)
__global__ void k_bayesl_post(int m, int iter, hb_key_t key, const uint8_t* __restrict__ active, const double* __restrict__ g,
                              double* __restrict__ vargL, double vare, double lambda, double lambda2) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m || !active[j]) return;
  double uu, zz;
  hb_draw_uz(key, HB_DOM_SNP, (uint32_t)iter, (uint32_t)j, HB_SL_IG, 0, &uu, &zz);
  double vargi = 1.0 / hb_invgauss_from_uz(sqrt(vare) * lambda / fabs(g[j]), lambda2, uu, zz);
  if (vargi >= 0) vargL[j] = vargi;
}
// max_i |r_i + shift| (one block): the bound behind the fixed-point scale of the LIMBS variant
__global__ void __launch_bounds__(1024) k_absmax(const double* __restrict__ r, int n, double shift, double* out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) v = fmax(v, fabs(r[i] + shift));
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) *out = v;
  }
}
__global__ void __launch_bounds__(1024) k_sum(const double* __restrict__ x, int n, double* out) {
  __shared__ double sh[1024];
  double a = 0;
  for (int i = threadIdx.x; i < n; i += 1024) a += x[i];
  a = block_sum_1024(a, sh);
  if (threadIdx.x == 0) *out = a;
}

// PIP / WPPA counters and effect sums
__global__ void k_pip(int m, const int32_t* __restrict__ tracker, double* __restrict__ nzrate) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m && tracker[j]) nzrate[j] += 1.0;
}
__global__ void k_wppa(int nw, const int* __restrict__ wstart, const int* __restrict__ wmem,
                       const int32_t* __restrict__ tracker, double* __restrict__ wppa) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  for (int q = wstart[w]; q < wstart[w + 1]; ++q)
    if (tracker[wmem[q]]) { wppa[w] += 1.0; return; }
}
__global__ void k_axpy1(int m, const double* __restrict__ g, double* __restrict__ gsum) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) gsum[j] += g[j];
}

// out[row] = sum_j x[row][j] * alpha[j]  (the X*g of Bayes.cpp:971).  grid (S, nchunk): block
// (s, c) covers slab s and a contiguous range of tiles and writes partial[c][row]; k_gemv_sum
// adds the chunks in order, so the result is deterministic.
__global__ void __launch_bounds__(1024) k_gemv_part(const uint8_t* __restrict__ Xp, const double* __restrict__ alpha, int m,
                                                    int R, int NRG, int CL, int T, int B, size_t slab_stride, size_t Npad,
                                                    double* __restrict__ partial) {
  extern __shared__ double sh[];  // CL * R
  const int s = blockIdx.x, c = blockIdx.y, nchunk = gridDim.y;
  const int tid = threadIdx.x, NTC = NRG * CL;
  const int t_lo = (int)((long long)T * c / nchunk), t_hi = (int)((long long)T * (c + 1) / nchunk);
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  if (tid < NTC) {
    const int rg = tid % NRG, cl = tid / NRG;
    const uint8_t* Xs = Xp + (size_t)s * slab_stride;
    for (int t = t_lo; t < t_hi; ++t)
      for (int cj = cl; cj < B; cj += CL) {
        const int j = t * B + cj;
        if (j >= m) continue;
        const double a = alpha[j];
        if (a == 0.0) continue;
        const uint4 v = *(const uint4*)(Xs + ((size_t)t * B + cj) * R + 16 * rg);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma((double)((w[i >> 2] >> (8 * (i & 3))) & 0xffu), a, acc[i]);
      }
#pragma unroll
    for (int i = 0; i < 16; ++i) sh[(size_t)cl * R + 16 * rg + i] = acc[i];
  }
  __syncthreads();
  for (int row = tid; row < R; row += blockDim.x) {
    double v = 0.0;
    for (int cl = 0; cl < CL; ++cl) v += sh[(size_t)cl * R + row];
    partial[(size_t)c * Npad + (size_t)s * R + row] = v;
  }
}
__global__ void k_gemv_sum(const double* __restrict__ partial, int nchunk, size_t Npad, int n, double* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n) return;
  double v = 0.0;
  for (int c = 0; c < nchunk; ++c) v += partial[(size_t)c * Npad + i];
  out[i] = v;
}

// Genetic values of a block of MCMC samples, `M %*% MCMCsamples$alpha` (R/bayes.r:303-304): part[c][row][rec] = sum over
// the SNPs of chunk c of x[row][j] * alpha[j][rec] for up to kGsRec records at once, so that X is read once per block of
// records instead of once per record.  CTA (slab s, chunk c); warp w: row group w % RW (lane l owns rows 4 l .. 4 l + 3 of
// it, the four genotype bytes of one 32-bit word) and records 16 (w / RW) .. + 15: 64 fp64 accumulators per thread, one
// int8 -> fp64 conversion per 16 fused multiply-adds.  Sub-tiles of kGsSub SNPs (genotypes of the slab + their alpha rows)
// are double-buffered in shared memory with cp.async.  fp64 FMA-bound: n * m * records / (64 per SM and clock).
constexpr int kGsRec = 64, kGsSub = 32;   // (64 accumulators per thread: 12 warps of ~170 registers fill an SM's register file)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__global__ void __launch_bounds__(384, 1) k_gemm_samples(const uint8_t* __restrict__ Xp, const double* __restrict__ at /* [m_pad][kGsRec] */,
                                                         int m_pad, int R, int RW, int NG, size_t slab_stride, size_t Npad,
                                                         double* __restrict__ part /* [nchunk][Npad][kGsRec] */) {
  extern __shared__ __align__(16) uint8_t gs_smem[];
  const int s = blockIdx.x, c = blockIdx.y, nchunk = gridDim.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  const int rw = warp % RW, rg = warp / RW;      // row group of 128 rows, record group of 16 records
  const int nsub = m_pad / kGsSub;
  const int sub_lo = (int)((long long)nsub * c / nchunk), sub_hi = (int)((long long)nsub * (c + 1) / nchunk);
  const size_t xbytes = (size_t)kGsSub * R, abytes = (size_t)kGsSub * kGsRec * 8;
  uint8_t* xb[2] = {gs_smem, gs_smem + xbytes + abytes};
  const uint8_t* Xs = Xp + (size_t)s * slab_stride;
  auto issue = [&](int sub, int b) {
    const uint8_t* xg = Xs + (size_t)sub * kGsSub * R;                   // SNP columns of a slab are contiguous blocks of R bytes
    const uint8_t* ag = (const uint8_t*)(at + (size_t)sub * kGsSub * kGsRec);
    for (size_t o = (size_t)tid * 16; o < xbytes; o += (size_t)nthr * 16) cp_async16(xb[b] + o, xg + o);
    for (size_t o = (size_t)tid * 16; o < abytes; o += (size_t)nthr * 16) cp_async16(xb[b] + xbytes + o, ag + o);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[i][k] = 0.0;
  if (sub_lo < sub_hi) issue(sub_lo, 0);
  for (int sub = sub_lo; sub < sub_hi; ++sub) {
    const int b = (sub - sub_lo) & 1;
    if (sub + 1 < sub_hi) { issue(sub + 1, b ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (rg < NG) {
      const uint8_t* xs = xb[b] + 128 * rw + 4 * lane;
      const double* as = (const double*)(xb[b] + xbytes) + 16 * rg;
#pragma unroll 2
      for (int j = 0; j < kGsSub; ++j) {
        const uint32_t w = *(const uint32_t*)(xs + (size_t)j * R);
        const double x0 = (double)(w & 0xffu), x1 = (double)((w >> 8) & 0xffu), x2 = (double)((w >> 16) & 0xffu), x3 = (double)(w >> 24);
        const double2* ap = (const double2*)(as + (size_t)j * kGsRec);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double2 a = ap[k];
          acc[0][2 * k] = fma(x0, a.x, acc[0][2 * k]); acc[0][2 * k + 1] = fma(x0, a.y, acc[0][2 * k + 1]);
          acc[1][2 * k] = fma(x1, a.x, acc[1][2 * k]); acc[1][2 * k + 1] = fma(x1, a.y, acc[1][2 * k + 1]);
          acc[2][2 * k] = fma(x2, a.x, acc[2][2 * k]); acc[2][2 * k + 1] = fma(x2, a.y, acc[2][2 * k + 1]);
          acc[3][2 * k] = fma(x3, a.x, acc[3][2 * k]); acc[3][2 * k + 1] = fma(x3, a.y, acc[3][2 * k + 1]);
        }
      }
    }
    __syncthreads();
  }
  if (rg < NG) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* o = part + ((size_t)c * Npad + (size_t)s * R + 128 * rw + 4 * lane + i) * kGsRec + 16 * rg;
#pragma unroll
      for (int k = 0; k < 16; ++k) o[k] = acc[i][k];
    }
  }
}
// at[j][r] = alpha[r][j] (records r0 .. r0 + nrec - 1 of the caller's m x records matrix), zero beyond m / nrec
__global__ void k_gs_transpose(const double* __restrict__ alpha, size_t ld, int m, int m_pad, int nrec, double* __restrict__ at) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)m_pad * kGsRec) return;
  const int j = (int)(idx / kGsRec), r = (int)(idx % kGsRec);
  at[idx] = (j < m && r < nrec) ? alpha[(size_t)r * ld + j] : 0.0;
}
// out[r][row] = sum over the chunks, in order (deterministic)
__global__ void k_gs_sum(const double* __restrict__ part, int nchunk, size_t Npad, int n, int nrec, double* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * nrec) return;
  const int r = (int)(idx / n), row = (int)(idx % n);
  double v = 0.0;
  for (int c = 0; c < nchunk; ++c) v += part[((size_t)c * Npad + row) * kGsRec + r];
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------
// host side of the engine
// ------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int hb_engine_create(const hb_engine_config* cfg, hb_engine** out) {
  if (!cfg || !out) return hb_set_error("hb_engine_create: null argument");
  if (cfg->n <= 0 || cfg->m <= 0) return hb_set_error("hb_engine_create: n and m must be positive");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
    return hb_set_error("hb_engine_create: CUDA device %d not available (%d visible) -- this engine has no CPU fallback",
                        cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return hb_set_error("hb_engine_create: device sm_%d%d is not a Blackwell (sm_100a) GPU", prop.major, prop.minor);
  hb_engine* e = new hb_engine();
  e->cfg = *cfg;
  e->n = cfg->n; e->m = cfg->m;
  e->nsm = prop.multiProcessorCount;
  e->B = cfg->tile_snps > 0 ? cfg->tile_snps : 256;
  // Scalar side of the mixture models: the ring of NG workers (default; 18.8 ms per sweep at the metric shape) or, with
  // HB_SERIAL=1, one serial CTA + helper CTAs (hb_serial.cuh; 20.8 ms: its chain never leaves one SM, but the package
  // traffic into that SM under the streaming load paces it -- DESIGN.md)
  e->serial = getenv("HB_SERIAL") && atoi(getenv("HB_SERIAL")) ? 1 : 0;
  if (e->B != hbk::kSerialB) e->serial = 0;   // (compiled for tiles of 256 SNPs)
  // tiles in flight between a tile's dots and its residual update (measured: 5 -> 20.0, 6 -> 19.0, 7 -> 18.9, 8 -> 18.8 ms)
  e->D = cfg->lag_tiles > 0 ? cfg->lag_tiles : 8;
  if (e->B != 64 && e->B != 128 && e->B != 256) { delete e; return hb_set_error("tile_snps must be 64, 128 or 256"); }
  if (e->D > 8) { delete e; return hb_set_error("lag_tiles must be <= 8"); }
  e->NG = 8;   // scalar CTAs (one worker each)
  if (const char* ng = getenv("HB_NG")) e->NG = std::max(1, std::min(16, atoi(ng)));
  e->NG = std::min(e->NG, std::max(1, e->nsm / 8));
  // Row slabs: one streaming CTA per slab, R = 16 lanes x RL rows (RL = 8, 16 or 24); the smallest R whose
  // slab count fits the SMs left beside the scalar CTAs keeps the most SMs streaming.
  int s_max = e->nsm - e->NG;
  if (cfg->n_slabs > 0) s_max = std::min(s_max, cfg->n_slabs);
  s_max = std::max(s_max, 1);
  e->RL = 0;
  for (int rl : {8, 16, 24})
    if (((size_t)e->n + 16 * rl - 1) / (16 * rl) <= (size_t)s_max) { e->RL = rl; break; }
  if (!e->RL && cfg->n_slabs > 0 && ((size_t)e->n + 383) / 384 <= (size_t)(e->nsm - e->NG)) e->RL = 24;   // n_slabs is a hint
  if (!e->RL) {
    delete e;
    return hb_set_error("n = %d rows per GPU exceeds this build's %d slabs x 384 rows; shard the individuals over more GPUs", cfg->n, s_max);
  }
  e->R = 16 * e->RL;
  e->S = (e->n + e->R - 1) / e->R;
  e->NRG = e->R / 16;
  e->Npad = (size_t)e->S * e->R;
  e->CL = 16;                          // column lanes of the prediction kernel (k_gemv_part)
  e->NTC = e->NRG * e->CL;
  e->NTCp = (e->NTC + 31) & ~31;
  e->NCW = e->B / 32;                  // compute warps: 32 columns of the tile each
  e->NAW = e->R / 128;                 // AXPY threads own 4 rows each
  e->block_threads = std::max(32 * (e->NCW + 1 + e->NAW), 2 * e->B);
  e->T = (e->m + e->B - 1) / e->B;
  e->m_pad = e->T * e->B;
  e->slab_stride = (size_t)e->T * e->B * e->R;
  // shared memory of a streaming CTA: NS sub-stages of B/4 columns + 2 residual-slab buffers + barriers
  e->SUBB = e->B / 4;
  e->stage_bytes = (size_t)e->SUBB * e->R;
  const size_t rbuf_bytes = 2 * (size_t)e->R * sizeof(double);
  const size_t bar_bytes = (2 * 16 + 4) * 8 + 64;
  const size_t budget = 226 * 1024;
  const size_t qbuf_bytes = (size_t)e->B * 12 + 16;   // the current tile's residual updates
  // sub-stages in flight: enough to cover the HBM latency at full bandwidth and no more -- every byte queued
  // beyond that only delays the latency-critical reads of the scalar CTAs (4 x 24 KB x 131 SMs = 12 MB in flight)
  int ns_cap = 4;
  if (const char* nsv = getenv("HB_NS")) ns_cap = std::max(2, std::min(16, atoi(nsv)));
  e->NS = (int)std::min<size_t>(ns_cap, (budget - rbuf_bytes - bar_bytes - qbuf_bytes - 256) / e->stage_bytes);
  if (e->NS < 2) { delete e; return hb_set_error("sub-stage of %zu bytes does not fit shared memory", e->stage_bytes); }
  e->off_rbuf = align_up((size_t)e->NS * e->stage_bytes, 128);
  e->off_bar = align_up(e->off_rbuf + rbuf_bytes, 16);
  e->off_qbuf = align_up(e->off_bar + bar_bytes, 16);
  const size_t stream_smem = e->off_qbuf + qbuf_bytes;
  const size_t scalar_smem = hbk::scalar_smem_bytes(e->B);
  e->KROW = hbk::scalar_krow(e->B);
  // test hook: fewer row slots, so that tiles take the path that reads the Gram band from global memory (a tile's result
  // must not depend on the path: tests/test_scalar_modes_gpu.py)
  if (const char* kr = getenv("HB_KROW")) e->KROW = std::max(1, std::min(e->KROW, atoi(kr)));
  e->smem_bytes = std::max(stream_smem, scalar_smem);
  if (e->serial) {
    e->KROW_S = hbk::serial_krow(e->B);
    e->pkg_stride = hbk::pkg_layout(e->B, e->KROW_S).stride;
    e->smem_bytes = std::max(e->smem_bytes, hbk::serial_smem_bytes(e->B));
    e->NH = std::max(1, std::min(16, e->nsm - e->S - 1));
    if (const char* nh = getenv("HB_NH")) e->NH = std::max(1, std::min(e->NH, atoi(nh)));
  }
  for (int v = 0; v < 4; ++v) {
    const void* fn = sweep_kernel_for(v == 0 ? 2 : v == 1 ? 4 : HB_MAX_FOLD, e->RL, v == 3);
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, e->block_threads, e->smem_bytes));
    if (occ < 1) { delete e; return hb_set_error("sweep kernel does not fit an SM (threads %d, smem %zu)", e->block_threads, e->smem_bytes); }
  }
  if (e->serial)
    for (int v = 0; v < 3; ++v) {
      const void* fn = serial_kernel_for(v == 0 ? 2 : v == 1 ? 4 : HB_MAX_FOLD, e->RL);
      CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
      int occ = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, e->block_threads, e->smem_bytes));
      if (occ < 1) { delete e; return hb_set_error("serial sweep kernel does not fit an SM (threads %d, smem %zu)", e->block_threads, e->smem_bytes); }
    }
  if (const char* ld = getenv("HB_LEAD")) {
    const int v = atoi(ld);
    if ((v == 1 || v == 2 || v == 4) && e->RL == 24) {
      const void* fn = v == 1 ? (const void*)k_sweep<512, 4, 24, false, false, 1>
                     : v == 2 ? (const void*)k_sweep<512, 4, 24, false, false, 2> : (const void*)k_sweep<512, 4, 24, false, false, 4>;
      CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
      e->lead = v;
    } else if (v) {
      fprintf(stderr, "[hb] HB_LEAD takes 1, 2 or 4 and needs 384-row slabs (RL = %d): ignored\n", e->RL);
    }
  }
  {
    // Integer dots in the streaming CTAs (dp4a on a 48-bit fixed-point residual, hb_limbs.h): the default for the mixture
    // models with at most 4 classes on 384-row slabs (the shapes with enough rows for the streaming side to matter; the
    // kernel is instantiated for those); HB_LIMBS=0 keeps the fp64 PRMT + DFMA loop.  Measured at the metric shape with
    // the scalar side no longer pacing the sweep: 16.3 ms against 18.6 ms per sweep.
    const char* lb = getenv("HB_LIMBS");
    const bool want = lb ? atoi(lb) != 0 : true;
    if (want && e->RL == 24) {
      CU(cudaFuncSetAttribute((const void*)k_sweep<512, 4, 24, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
      CU(cudaFuncSetAttribute((const void*)k_sweep<512, 2, 24, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
      CU(cudaMalloc(&e->absmax_dev, 8));
      e->limbs = 1;
    } else if (lb && want) {
      fprintf(stderr, "[hb] HB_LIMBS needs 384-row slabs (RL = 24); this engine has RL = %d: fp64 dots\n", e->RL);
    }
  }
  e->scalar0 = e->S;
  if (const char* cl = getenv("HB_CLUSTER")) {
    // experimental: the whole grid in clusters of 2 so that the scalar workers 2c, 2c+1 share distributed shared memory
    const int s0 = (e->S + 1) & ~1;
    if (atoi(cl) && e->B == 256 && e->block_threads == 512 && e->NG % 2 == 0 && s0 + e->NG <= e->nsm) {
      e->cluster2 = 1;
      e->scalar0 = s0;
    }
  }
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&e->ev[i]));
  const size_t xbytes = (size_t)e->S * e->slab_stride;
  CU(cudaMalloc(&e->Xp, xbytes));
  CU(cudaMemsetAsync(e->Xp, 0, xbytes, e->stream));
  CU(cudaMalloc(&e->r, e->Npad * 8)); CU(cudaMemsetAsync(e->r, 0, e->Npad * 8, e->stream));
  CU(cudaMalloc(&e->u, e->Npad * 8)); CU(cudaMemsetAsync(e->u, 0, e->Npad * 8, e->stream));
  const size_t mp = e->m_pad;
  CU(cudaMalloc(&e->xpx, mp * 8)); CU(cudaMemsetAsync(e->xpx, 0, mp * 8, e->stream));
  CU(cudaMalloc(&e->g, mp * 8)); CU(cudaMemsetAsync(e->g, 0, mp * 8, e->stream));
  CU(cudaMalloc(&e->gsum, mp * 8)); CU(cudaMemsetAsync(e->gsum, 0, mp * 8, e->stream));
  CU(cudaMalloc(&e->nzrate, mp * 8)); CU(cudaMemsetAsync(e->nzrate, 0, mp * 8, e->stream));
  CU(cudaMalloc(&e->vargL, mp * 8)); CU(cudaMemsetAsync(e->vargL, 0, mp * 8, e->stream));
  CU(cudaMalloc(&e->active, mp)); CU(cudaMemsetAsync(e->active, 0, mp, e->stream));
  CU(cudaMalloc(&e->tracker, mp * 4)); CU(cudaMemsetAsync(e->tracker, 0, mp * 4, e->stream));
  CU(cudaMalloc(&e->dacc, mp * 8));
  CU(cudaMalloc(&e->q_snp, mp * 4));
  CU(cudaMalloc(&e->q_delta, mp * 8));
  CU(cudaMalloc(&e->tile_cnt, (size_t)e->T * 4));
  if (cfg->world > 1) {
    if (cfg->world > 8 || cfg->rank < 0 || cfg->rank >= cfg->world) { hb_engine_destroy(e); return hb_set_error("hb_engine_create: bad rank/world (at most 8 ranks)"); }
    CU(cudaMalloc(&e->acc2, 2 * mp * 8));
    CU(cudaMemsetAsync(e->acc2, 0, 2 * mp * 8, e->stream));
    e->peer_acc2[cfg->rank] = e->acc2;
  }
  CU(cudaMalloc(&e->corr, std::max<size_t>(1, (size_t)e->m_pad * (e->D - 1)) * 8));
  CU(cudaMalloc(&e->ctrl, 64)); CU(cudaMemsetAsync(e->ctrl, 0, 64, e->stream));
  if (e->serial) {
    CU(cudaMalloc(&e->pkg, (size_t)e->T * e->pkg_stride));
    CU(cudaMalloc(&e->pkg_flag, (size_t)e->T * 4));
    CU(cudaMalloc(&e->miss_tile, 64));
  }
  CU(cudaMalloc(&e->out_dev, sizeof(SweepOutDev)));
  CU(cudaMalloc(&e->post_partial, (size_t)kPostBlocks * (HB_MAX_FOLD + 1) * 8));
  CU(cudaMalloc(&e->fold_dev, HB_MAX_FOLD * 8));
  CU(cudaStreamSynchronize(e->stream));
  *out = e;
  return 0;
}

extern "C" void hb_engine_destroy(hb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaFree(e->Xp); cudaFree(e->r); cudaFree(e->u); cudaFree(e->xpx); cudaFree(e->g); cudaFree(e->gsum);
  cudaFree(e->nzrate); cudaFree(e->wppa); cudaFree(e->vargL); cudaFree(e->active); cudaFree(e->tracker);
  cudaFree(e->gram); cudaFree(e->dacc); cudaFree(e->q_snp); cudaFree(e->q_delta);
  cudaFree(e->tile_cnt); cudaFree(e->corr); cudaFree(e->trace); cudaFree(e->absmax_dev);
  cudaFree(e->pkg); cudaFree(e->pkg_flag); cudaFree(e->miss_tile);
  for (int g = 0; g < 8; ++g) if (e->peer_acc2[g] && e->peer_acc2[g] != e->acc2) cudaIpcCloseMemHandle(e->peer_acc2[g]);
  cudaFree(e->acc2); cudaFree(e->ctrl); cudaFree(e->prm); cudaFree(e->out_dev); cudaFree(e->post_partial); cudaFree(e->fold_dev); cudaFree(e->wstart); cudaFree(e->wmem);
  for (int i = 0; i < 4; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int hb_engine_describe(hb_engine* e, int* n_slabs, int* rows_per_slab, int* tile_snps, int* lag_tiles,
                                  uint64_t* geno_bytes, uint64_t* gram_bytes) {
  if (!e) return hb_set_error("null engine");
  if (n_slabs) *n_slabs = e->S;
  if (rows_per_slab) *rows_per_slab = e->R;
  if (tile_snps) *tile_snps = e->B;
  if (lag_tiles) *lag_tiles = e->D;
  if (geno_bytes) *geno_bytes = (uint64_t)e->S * e->slab_stride;
  if (gram_bytes) *gram_bytes = (uint64_t)e->T * e->D * e->B * e->B * 4;
  return 0;
}

// Host matrix -> device tiles.  The matrix goes through two pinned bounce buffers in chunks of <= 256 MB: host threads
// fill one (plain copies of int8 columns; fp64 columns converted and checked on the way) while the DMA engine drains the
// other, and the range check of int8 input runs on the device inside the pack kernel -- the single-threaded host scan and the
// pageable copies of the first version took 19 s for the 50 GB of the metric shape.
static int load_chunked(hb_engine* e, const int8_t* X8, const double* X64, size_t ld) {
  CU(cudaSetDevice(e->cfg.device));
  const int n = e->n, m = e->m;
  size_t budget = (size_t)256u << 20;
  if (const char* cb = getenv("HB_LOAD_CHUNK")) budget = std::max<size_t>(1, (size_t)atoll(cb));   // (test hook: many small chunks)
  size_t cols_per_chunk = std::max<size_t>(1, budget / (size_t)n);
  cols_per_chunk = std::min<size_t>(cols_per_chunk, (size_t)m);
  const size_t chunk_bytes = cols_per_chunk * (size_t)n;
  int8_t* stage[2] = {nullptr, nullptr};
  int8_t* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int* d_bad = nullptr;
  int rc = 0;
  std::atomic<long long> bad64{-1};   // first fp64 entry (linear index inside its chunk) that is not 0, 1 or 2
  const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
#define TRYL(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); goto done; } } while (0)
  for (int b = 0; b < 2; ++b) {
    TRYL(cudaMalloc(&stage[b], chunk_bytes));
    TRYL(cudaHostAlloc(&pin[b], chunk_bytes, cudaHostAllocDefault));
    TRYL(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
  }
  TRYL(cudaMalloc(&d_bad, 8));
  TRYL(cudaMemsetAsync(d_bad, 0, 8, e->stream));
  {
    int b = 0;
    for (size_t c0 = 0; c0 < (size_t)m; c0 += cols_per_chunk, b ^= 1) {
      const size_t nc = std::min(cols_per_chunk, (size_t)m - c0);
      TRYL(cudaEventSynchronize(ev[b]));   // the copy that last read this bounce buffer is through (no-op the first time)
      // fill the bounce buffer: columns split over the host threads
      const unsigned nt = (unsigned)std::min<size_t>(hw, std::max<size_t>(1, nc * (size_t)n >> 20));
      auto fill = [&](unsigned w) {
        const size_t ca = nc * w / nt, cb = nc * (w + 1) / nt;
        if (X64) {
          for (size_t c = ca; c < cb; ++c) {
            const double* col = X64 + (c0 + c) * ld;
            int8_t* dst = pin[b] + c * (size_t)n;
            bool ok = true;
            for (int i = 0; i < n; ++i) { const double v = col[i]; ok &= (v == 0.0) | (v == 1.0) | (v == 2.0); dst[i] = (int8_t)(int)v; }
            if (!ok)
              for (int i = 0; i < n; ++i) {
                const double v = col[i];
                if (!(v == 0.0 || v == 1.0 || v == 2.0)) {
                  long long want = -1;
                  bad64.compare_exchange_strong(want, (long long)(c * (size_t)n + i));
                  break;
                }
              }
          }
        } else if (ld == (size_t)n) {
          memcpy(pin[b] + ca * (size_t)n, X8 + (c0 + ca) * ld, (cb - ca) * (size_t)n);
        } else {
          for (size_t c = ca; c < cb; ++c) memcpy(pin[b] + c * (size_t)n, X8 + (c0 + c) * ld, (size_t)n);
        }
      };
      if (nt <= 1) fill(0);
      else {
        std::vector<std::thread> th;
        for (unsigned w = 1; w < nt; ++w) th.emplace_back(fill, w);
        fill(0);
        for (auto& t : th) t.join();
      }
      if (bad64.load() >= 0) {
        const long long q = bad64.load();
        const size_t c = (size_t)q / (size_t)n;
        const int i = (int)((size_t)q % (size_t)n);
        rc = hb_set_error("genotype (%d,%zu) = %g: this engine holds genotypes as int8 in {0,1,2}", i, c0 + c, X64[(c0 + c) * ld + i]);
        goto done;
      }
      TRYL(cudaMemcpyAsync(stage[b], pin[b], nc * (size_t)n, cudaMemcpyHostToDevice, e->stream));
      TRYL(cudaEventRecord(ev[b], e->stream));
      const size_t work = (size_t)e->S * e->NRG * nc;
      k_pack_i8<<<(unsigned)((work + 255) / 256), 256, 0, e->stream>>>(stage[b], (size_t)n, n, (int)c0, (int)nc, e->Xp, e->S, e->R,
                                                                      e->NRG, e->T, e->B, X64 ? nullptr : d_bad);
      TRYL(cudaGetLastError());
      // (stage[b] is next written two chunks later, behind this kernel on the same stream)
    }
  }
  {
    int hbad[2] = {0, 0};
    TRYL(cudaMemcpyAsync(hbad, d_bad, 8, cudaMemcpyDeviceToHost, e->stream));
    TRYL(cudaStreamSynchronize(e->stream));
    if (hbad[0]) { rc = hb_set_error("genotype value %d outside {0,1,2} (column %d)", hbad[1], hbad[0] - 1); goto done; }
  }
  e->geno_ready = true;
  e->gram_ready = false;
done:
#undef TRYL
  cudaStreamSynchronize(e->stream);
  for (int b = 0; b < 2; ++b) { cudaFree(stage[b]); cudaFreeHost(pin[b]); if (ev[b]) cudaEventDestroy(ev[b]); }
  cudaFree(d_bad);
  return rc;
}

extern "C" int hb_engine_load_geno_i8(hb_engine* e, const int8_t* X, size_t ld) {
  if (!e || !X) return hb_set_error("hb_engine_load_geno_i8: null argument");
  if (ld < (size_t)e->n) return hb_set_error("hb_engine_load_geno_i8: ld < n");
  return load_chunked(e, X, nullptr, ld);
}
extern "C" int hb_engine_load_geno_f64(hb_engine* e, const double* X, size_t ld) {
  if (!e || !X) return hb_set_error("hb_engine_load_geno_f64: null argument");
  if (ld < (size_t)e->n) return hb_set_error("hb_engine_load_geno_f64: ld < n");
  return load_chunked(e, nullptr, X, ld);
}

// PLINK .bed image -> device tiles without the fp64 / big.matrix detour of R/bayes.r:284.
extern "C" int hb_engine_load_bed(hb_engine* e, const uint8_t* file, size_t len, int nid, const int32_t* rows, int impt,
                                  int dominance) {
  if (!e || !file) return hb_set_error("hb_engine_load_bed: null argument");
  if (nid <= 0) return hb_set_error("hb_engine_load_bed: nid must be positive");
  const size_t bps = ((size_t)nid + 3) / 4;
  if (len < 3 || file[0] != 0x6c || file[1] != 0x1b) return hb_set_error("hb_engine_load_bed: not a PLINK .bed image (magic bytes)");
  if (file[2] != 0x01) return hb_set_error("hb_engine_load_bed: individual-major .bed files are not supported");
  if (len < 3 + bps * (size_t)e->m) return hb_set_error("hb_engine_load_bed: image has %zu bytes, %zu needed for %d individuals x %d SNPs", len, 3 + bps * (size_t)e->m, nid, e->m);
  if (!rows && nid != e->n) return hb_set_error("hb_engine_load_bed: the file has %d individuals, the engine %d rows; pass the row selection", nid, e->n);
  if (rows)
    for (int i = 0; i < e->n; ++i)
      if (rows[i] < 0 || rows[i] >= nid) return hb_set_error("hb_engine_load_bed: rows[%d] = %d outside the file's %d individuals", i, rows[i], nid);
  CU(cudaSetDevice(e->cfg.device));
  const size_t cols_per_chunk = std::min<size_t>((size_t)e->m, std::max<size_t>(1, (size_t)(256u << 20) / bps));
  uint8_t *stage = nullptr, *info = nullptr;
  int32_t* drows = nullptr;
  std::vector<uint8_t> hinfo(cols_per_chunk);
  int rc = 0;
#define TRYB(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); goto done; } } while (0)
  TRYB(cudaMalloc(&stage, cols_per_chunk * bps));
  TRYB(cudaMalloc(&info, cols_per_chunk));
  if (rows) {
    TRYB(cudaMalloc(&drows, (size_t)e->n * 4));
    TRYB(cudaMemcpyAsync(drows, rows, (size_t)e->n * 4, cudaMemcpyHostToDevice, e->stream));
  }
  for (size_t c0 = 0; c0 < (size_t)e->m; c0 += cols_per_chunk) {
    const size_t nc = std::min(cols_per_chunk, (size_t)e->m - c0);
    TRYB(cudaMemcpyAsync(stage, file + 3 + c0 * bps, nc * bps, cudaMemcpyHostToDevice, e->stream));
    hb::k_bed_info<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, e->stream>>>(stage, bps, nid, (int)nc, dominance, info);
    TRYB(cudaGetLastError());
    if (!impt) {
      TRYB(cudaMemcpyAsync(hinfo.data(), info, nc, cudaMemcpyDeviceToHost, e->stream));
      TRYB(cudaStreamSynchronize(e->stream));
      for (size_t c = 0; c < nc; ++c)
        if (hinfo[c] & 0x80) {
          rc = hb_set_error("hb_engine_load_bed: SNP %zu has missing genotypes and impute is off; this engine holds genotypes in {0,1,2}", c0 + c);
          goto done;
        }
    }
    const size_t work = (size_t)e->S * e->NRG * nc;
    k_pack_bed<<<(unsigned)((work + 255) / 256), 256, 0, e->stream>>>(stage, bps, drows, e->n, (int)c0, (int)nc, dominance, info,
                                                                     e->Xp, e->S, e->R, e->NRG, e->T, e->B);
    TRYB(cudaGetLastError());
    TRYB(cudaStreamSynchronize(e->stream));
  }
  e->geno_ready = true;
  e->gram_ready = false;
done:
#undef TRYB
  cudaFree(stage); cudaFree(info); cudaFree(drows);
  return rc;
}

extern "C" int hb_engine_synth_geno(hb_engine* e, uint64_t seed, int64_t row_offset) {
  if (!e) return hb_set_error("null engine");
  if (row_offset % 4 != 0) return hb_set_error("row_offset must be a multiple of 4");
  CU(cudaSetDevice(e->cfg.device));
  const size_t work = (size_t)e->S * e->NRG * (size_t)e->m;
  const size_t blocks = (work + 255) / 256;
  if (blocks > 0x7fffffffull) return hb_set_error("synthetic matrix too large for one launch");
  k_synth<<<(unsigned)blocks, 256, 0, e->stream>>>(e->Xp, e->n, e->m, e->S, e->R, e->NRG, e->T, e->B, hb_make_key(seed),
                                                  (long long)row_offset);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  e->geno_ready = true;
  e->gram_ready = false;
  return 0;
}

extern "C" int hb_synth_geno_host_cols(int8_t* X, int n, int col0, int ncols, uint64_t seed, int64_t row_offset);
extern "C" int hb_synth_geno_host(int8_t* X, int n, int m, uint64_t seed, int64_t row_offset) {
  return hb_synth_geno_host_cols(X, n, 0, m, seed, row_offset);
}
extern "C" int hb_synth_geno_host_cols(int8_t* X, int n, int col0, int ncols, uint64_t seed, int64_t row_offset) {
  if (!X || row_offset % 4 != 0 || col0 < 0 || ncols < 0) return hb_set_error("hb_synth_geno_host: bad argument");
  hb_key_t key = hb_make_key(seed);
  for (int j = col0; j < col0 + ncols; ++j) {
    double t0, t1;
    hb_synth_thresholds(key, (uint32_t)j, &t0, &t1);
    int8_t* col = X + (size_t)(j - col0) * n;
    for (int i = 0; i < n; i += 4) {
      uint32_t w = hb_synth_word(key, (uint32_t)j, (uint64_t)(row_offset + i) >> 2, t0, t1);
      for (int b = 0; b < 4 && i + b < n; ++b) col[i + b] = (int8_t)((w >> (8 * b)) & 0xff);
    }
  }
  return 0;
}

extern "C" int hb_engine_col_stats(hb_engine* e, double* xpx, double* sumx) {
  if (!e || !xpx || !sumx) return hb_set_error("hb_engine_col_stats: null argument");
  if (!e->geno_ready) return hb_set_error("hb_engine_col_stats: genotypes not loaded");
  CU(cudaSetDevice(e->cfg.device));
  double *dx = nullptr, *ds = nullptr;
  CU(cudaMalloc(&dx, (size_t)e->m * 8));
  CU(cudaMalloc(&ds, (size_t)e->m * 8));
  const size_t threads = (size_t)e->m * 32;
  k_col_stats<<<(unsigned)((threads + 255) / 256), 256, 0, e->stream>>>(e->Xp, e->m, e->S, e->R, e->NRG, e->T, e->B, dx, ds);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(xpx, dx, (size_t)e->m * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(sumx, ds, (size_t)e->m * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(dx); cudaFree(ds);
  return 0;
}

extern "C" int hb_engine_set_snp_info(hb_engine* e, const double* xpx_global, const uint8_t* active) {
  if (!e || !xpx_global || !active) return hb_set_error("hb_engine_set_snp_info: null argument");
  CU(cudaSetDevice(e->cfg.device));
  e->xpx_max = 0.0;
  for (int q = 0; q < e->m; ++q) e->xpx_max = std::max(e->xpx_max, xpx_global[q]);
  CU(cudaMemcpyAsync(e->xpx, xpx_global, (size_t)e->m * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->active, active, (size_t)e->m, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->info_ready = true;
  return 0;
}

extern "C" int hb_engine_build_gram(hb_engine* e) {
  if (!e) return hb_set_error("null engine");
  if (!e->geno_ready) return hb_set_error("hb_engine_build_gram: genotypes not loaded");
  CU(cudaSetDevice(e->cfg.device));
  const size_t gbytes = (size_t)e->T * e->D * e->B * e->B * 4;
  if (!e->gram) CU(cudaMalloc(&e->gram, gbytes));
  const int nb = e->B / 64;
  // grid.z is limited to 65535 tiles per launch
  for (int t0 = 0; t0 < e->T; t0 += 65535) {
    const int nt = std::min(65535, e->T - t0);
    dim3 grid(e->D * nb, nb, nt);
    if (getenv("HB_GRAM_DP4A")) k_gram_dp4a<<<grid, 256, 0, e->stream>>>(e->Xp, e->gram, e->S, e->R, e->T, e->B, e->D, t0);
    else k_gram_imma<<<grid, 128, 0, e->stream>>>(e->Xp, e->gram, e->S, e->R, e->T, e->B, e->D, t0);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(e->stream));
  e->gram_ready = true;
  return 0;
}

static int copy_vec(hb_engine* e, double* dst_dev, const double* src_host, size_t cnt) {
  CU(cudaSetDevice(e->cfg.device));
  CU(cudaMemcpyAsync(dst_dev, src_host, cnt * 8, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
static int fetch_vec(hb_engine* e, double* dst_host, const double* src_dev, size_t cnt) {
  CU(cudaSetDevice(e->cfg.device));
  CU(cudaMemcpyAsync(dst_host, src_dev, cnt * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
extern "C" int hb_engine_set_residual(hb_engine* e, const double* y) { if (!e || !y) return hb_set_error("null argument"); return copy_vec(e, e->r, y, e->n); }
extern "C" int hb_engine_get_residual(hb_engine* e, double* y) { if (!e || !y) return hb_set_error("null argument"); return fetch_vec(e, y, e->r, e->n); }
extern "C" int hb_engine_set_u(hb_engine* e, const double* u) { if (!e || !u) return hb_set_error("null argument"); return copy_vec(e, e->u, u, e->n); }
extern "C" int hb_engine_get_u(hb_engine* e, double* u) { if (!e || !u) return hb_set_error("null argument"); return fetch_vec(e, u, e->u, e->n); }
extern "C" int hb_engine_device_state(hb_engine* e, double** r_dev, double** u_dev, void** cuda_stream, int* device, int* n) {
  if (!e) return hb_set_error("hb_engine_device_state: null argument");
  if (r_dev) *r_dev = e->r;
  if (u_dev) *u_dev = e->u;
  if (cuda_stream) *cuda_stream = (void*)e->stream;
  if (device) *device = e->cfg.device;
  if (n) *n = e->n;
  return 0;
}
extern "C" int hb_engine_set_effects(hb_engine* e, const double* g) { if (!e || !g) return hb_set_error("null argument"); return copy_vec(e, e->g, g, e->m); }
extern "C" int hb_engine_get_effects(hb_engine* e, double* g) { if (!e || !g) return hb_set_error("null argument"); return fetch_vec(e, g, e->g, e->m); }
extern "C" int hb_engine_set_vargL(hb_engine* e, const double* v) { if (!e || !v) return hb_set_error("null argument"); return copy_vec(e, e->vargL, v, e->m); }
extern "C" int hb_engine_get_effect_sums(hb_engine* e, double* g) { if (!e || !g) return hb_set_error("null argument"); return fetch_vec(e, g, e->gsum, e->m); }
extern "C" int hb_engine_get_tracker(hb_engine* e, int32_t* t) {
  if (!e || !t) return hb_set_error("null argument");
  CU(cudaSetDevice(e->cfg.device));
  CU(cudaMemcpyAsync(t, e->tracker, (size_t)e->m * 4, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

extern "C" int hb_engine_sweep(hb_engine* e, const hb_sweep_in* in, hb_sweep_out* out) {
  if (!e || !in || !out) return hb_set_error("hb_engine_sweep: null argument");
  if (!e->geno_ready || !e->info_ready || !e->gram_ready)
    return hb_set_error("hb_engine_sweep: engine not ready (genotypes %d, snp info %d, gram %d)", (int)e->geno_ready,
                        (int)e->info_ready, (int)e->gram_ready);
  if (in->model_index < 1 || in->model_index > 6) return hb_set_error("hb_engine_sweep: bad model_index");
  const int F = in->n_fold;
  if (F < 2 || F > HB_MAX_FOLD) return hb_set_error("hb_engine_sweep: n_fold must be in [2, %d]", HB_MAX_FOLD);
  CU(cudaSetDevice(e->cfg.device));
  if (!e->prm) CU(cudaMalloc(&e->prm, (size_t)kPrmFields * e->m_pad * 8));
  PrepParams pp;
  memset(&pp, 0, sizeof pp);
  pp.m = e->m; pp.m_pad = e->m_pad; pp.T = e->T; pp.iter = in->iter; pp.model = in->model_index; pp.F = F;
  for (int k = 0; k < HB_MAX_FOLD; ++k) { pp.fold[k] = in->fold[k]; pp.logpi[k] = in->logpi[k]; pp.vara_fold[k] = in->vara_fold[k]; }
  pp.vare = in->vare; pp.dfvara = in->dfvara; pp.s2varg = in->s2varg;
  pp.key = hb_make_key(e->cfg.seed);
  // class decisions by thresholds need class-ordered variances (every cumulative probability then decreases in rhs^2)
  {
    bool ordered = true;
    if (in->model_index == HB_MODEL_R)
      for (int k = 2; k < F; ++k) ordered = ordered && (in->vara_fold[k] >= in->vara_fold[k - 1]) && (in->vara_fold[k - 1] > 0);
    const bool mixture = in->model_index == HB_MODEL_B || in->model_index == HB_MODEL_C || in->model_index == HB_MODEL_R;
    pp.use_thr = (mixture && ordered && !getenv("HB_NO_THR")) ? 1 : 0;
  }

  SweepParams sp;
  memset(&sp, 0, sizeof sp);
  sp.Xp = e->Xp; sp.r = e->r; sp.u = e->u; sp.xpx = e->xpx; sp.active = e->active; sp.g = e->g; sp.tracker = e->tracker;
  sp.gram = e->gram; sp.dacc = e->dacc; sp.q_snp = e->q_snp; sp.q_delta = e->q_delta;
  if (getenv("HB_TRACE") && !e->trace) { CU(cudaMalloc(&e->trace, (size_t)e->T * 128)); CU(cudaMemset(e->trace, 0, (size_t)e->T * 128)); }
  sp.trace = e->trace;
  sp.world = std::max(1, e->cfg.world); sp.rank = e->cfg.rank;
  if (sp.world > 1) {
    if (!e->peers_ready) return hb_set_error("hb_engine_sweep: world = %d but hb_engine_set_peers has not been called", sp.world);
    // second-level accumulators alternate between two buffers: a rank that is one sweep ahead adds into the buffer
    // this rank has already cleared (k_prep of the previous sweep), never into the one still in use
    for (int g = 0; g < sp.world; ++g) sp.peer_acc[g] = e->peer_acc2[g] + (size_t)(e->sweep_no & 1) * e->m_pad;
  }
  sp.tile_cnt = e->tile_cnt; sp.corr = e->corr; sp.ctrl = e->ctrl; sp.prm = e->prm; sp.out = e->out_dev;
  sp.slab_stride = e->slab_stride; sp.m_pad = e->m_pad;
  sp.n = e->n; sp.m = e->m; sp.S = e->S; sp.R = e->R; sp.T = e->T; sp.B = e->B; sp.D = e->D;
  sp.NS = e->NS; sp.NCW = e->NCW; sp.NAW = e->NAW; sp.SUBB = e->SUBB; sp.NG = e->NG; sp.KROW = e->KROW;
  sp.stage_bytes = (uint32_t)e->stage_bytes; sp.off_rbuf = (uint32_t)e->off_rbuf; sp.off_qbuf = (uint32_t)e->off_qbuf;
  sp.off_bar = (uint32_t)e->off_bar;
  sp.model = in->model_index; sp.F = F; sp.use_thr = pp.use_thr;
  for (int k = 0; k < HB_MAX_FOLD; ++k) sp.fold[k] = in->fold[k];
  sp.logpi0 = in->logpi[0];
  sp.mu_shift = in->mu_shift;
  // fixed-point scale of the dot accumulators: |x_j'r| <= ||x_j|| ||r|| <= sqrt(max_j xpx_j) ||r|| for the whole
  // dot and for every slab's part of it, with a factor 8 of head-room for the residual changing during the
  // sweep; the sums must stay below 2^54 (the low byte of an accumulator counts the slabs that have arrived)
  {
    const double xmax = e->xpx_max > 0 ? e->xpx_max : 4.0 * (double)e->n * std::max(1, e->cfg.world);
    double bound = sqrt(xmax * std::max(in->rnorm2_bound, 1e-300)) * 8.0;
    int ex = (int)floor(log2(hbk::kFixLimit / bound));
    ex = std::max(-900, std::min(ex, 900));
    sp.dscale = ldexp(1.0, ex);
    sp.inv_dscale = ldexp(1.0, -ex);
  }
  { const char* dbg = getenv("HB_DEBUG"); sp.dbg = dbg ? atoi(dbg) : 0; }
  // the genotype tiles are streamed with an L2 evict_first hint, so that they do not push the scalar side's small working
  // set (Gram rows, correction slots, parameters) out of L2: -3 % per sweep; HB_XEVICT=0 switches it off
  sp.xevict = 1;
  if (const char* xe = getenv("HB_XEVICT")) sp.xevict = atoi(xe);
  sp.near_frac = 0.49;   // |rhs| within 30 % of the boundary
  if (const char* nf_ = getenv("HB_NEAR")) sp.near_frac = atof(nf_);
  if (getenv("HB_PHASES")) sp.dbg |= 64;   // per-phase cycle counters of the serial CTA (they cost ~1 us per tile)

  const bool dense_model = in->model_index == HB_MODEL_RR || in->model_index == HB_MODEL_A || in->model_index == HB_MODEL_L;
  CU(cudaEventRecord(e->ev[0], e->stream));
  k_prep<<<(e->m_pad + 255) / 256, 256, 0, e->stream>>>(pp, e->xpx, e->active, e->g, e->vargL, e->prm, e->dacc, e->ctrl, e->out_dev,
                                                       e->tile_cnt, e->q_snp, (unsigned long long*)e->q_delta, (unsigned long long*)e->corr, e->B, e->D - 1,
                                                       e->acc2 ? e->acc2 + (size_t)((e->sweep_no + 1) & 1) * e->m_pad : nullptr,
                                                       e->pkg_flag, e->miss_tile);
  CU(cudaGetLastError());
  CU(cudaEventRecord(e->ev[1], e->stream));
  {
    const void* fn = sweep_kernel_for(in->model_index == HB_MODEL_R ? F : 2, e->RL, dense_model);
    const int nf_kernel = in->model_index == HB_MODEL_R ? F : 2;
    // serial CTA + helpers: mixture models whose classes are decided by thresholds
    const bool serial = e->serial && !dense_model && pp.use_thr && !e->lead && !e->cluster2;
    int nscalar = e->NG;
    if (serial) {
      fn = serial_kernel_for(nf_kernel, e->RL);
      sp.serial = 1; sp.NH = e->NH; sp.KROW_S = e->KROW_S; sp.pkg_stride = e->pkg_stride;
      sp.pkg = e->pkg; sp.pkg_flag = e->pkg_flag; sp.miss_tile = e->miss_tile;
      nscalar = 1 + e->NH;
    }
    if (e->lead && !dense_model && nf_kernel > 2 && nf_kernel <= 4)
      fn = e->lead == 1 ? (const void*)k_sweep<512, 4, 24, false, false, 1>
         : e->lead == 2 ? (const void*)k_sweep<512, 4, 24, false, false, 2> : (const void*)k_sweep<512, 4, 24, false, false, 4>;
    else if (e->limbs && !dense_model && nf_kernel <= 4 && !serial) {
      // fixed-point scale of the residual limbs: |q| = |r| rscale must stay below 2^47 while the residual moves during
      // the sweep (factor 4 of head-room over today's largest element; an overflow aborts the sweep with a message)
      double amax = 0.0;
      k_absmax<<<1, 1024, 0, e->stream>>>(e->r, e->n, in->mu_shift, e->absmax_dev);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(&amax, e->absmax_dev, 8, cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      const double bound = 4.0 * std::max(amax, 1e-300);
      int ex_r = (int)floor(log2(140737488355328.0 / bound));   // 2^47 / bound
      const int ex_d = (int)lrint(log2(sp.dscale));
      ex_r = std::max(ex_d - 40, std::min(ex_r, ex_d + 40));
      sp.rscale = ldexp(1.0, ex_r);
      sp.rshift = ex_r - ex_d;
      fn = nf_kernel <= 2 ? (const void*)k_sweep<512, 2, 24, false, true> : (const void*)k_sweep<512, 4, 24, false, true>;
    }
    void* args[] = {(void*)&sp};
    bool launched = false;
    if (e->cluster2) {
      sp.cluster2 = 1;
      sp.scalar0 = e->scalar0;
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof lc);
      lc.gridDim = dim3(e->scalar0 + e->NG);
      lc.blockDim = dim3(e->block_threads);
      lc.dynamicSmemBytes = e->smem_bytes;
      lc.stream = e->stream;
      cudaLaunchAttribute at[2];
      memset(at, 0, sizeof at);
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeCooperative;
      at[1].val.cooperative = 1;
      lc.attrs = at;
      lc.numAttrs = 2;
      int ncl = 0;
      cudaError_t err = cudaOccupancyMaxActiveClusters(&ncl, fn, &lc);
      if (err == cudaSuccess && 2 * ncl >= (int)lc.gridDim.x) err = cudaLaunchKernelExC(&lc, fn, args);
      else if (err == cudaSuccess) err = cudaErrorCooperativeLaunchTooLarge;
      if (err == cudaSuccess) launched = true;
      else {
        fprintf(stderr, "[hb] cluster launch not possible (%s, %d clusters of 2 fit, %u blocks wanted): plain launch\n",
                cudaGetErrorString(err), ncl, lc.gridDim.x);
        (void)cudaGetLastError();
        e->cluster2 = 0;
        e->scalar0 = e->S;
      }
    }
    if (!launched) {
      sp.cluster2 = 0;
      sp.scalar0 = e->S;
      CU(cudaLaunchCooperativeKernel(fn, dim3(e->S + nscalar), dim3(e->block_threads), args, e->smem_bytes, e->stream));
    }
  }
  CU(cudaEventRecord(e->ev[2], e->stream));
  CU(cudaMemcpyAsync(e->fold_dev, in->fold, HB_MAX_FOLD * 8, cudaMemcpyHostToDevice, e->stream));
  k_post1<<<kPostBlocks, kPostThreads, 0, e->stream>>>(e->m, in->model_index, e->tracker, e->g, e->active, e->fold_dev, e->post_partial);
  CU(cudaGetLastError());
  k_post2<<<1, 32, 0, e->stream>>>(e->post_partial, e->out_dev);
  CU(cudaGetLastError());
  if (in->model_index == HB_MODEL_L) {
    k_bayesl_post<<<(e->m + 255) / 256, 256, 0, e->stream>>>(e->m, in->iter, pp.key, e->active, e->g, e->vargL, in->vare,
                                                             in->lambda, in->lambda2);
    CU(cudaGetLastError());
    k_sum<<<1, 1024, 0, e->stream>>>(e->vargL, e->m, &e->out_dev->sum_vargL);
    CU(cudaGetLastError());
  }
  k_tail<<<1, 1024, 0, e->stream>>>(e->r, e->u, e->n, e->out_dev, e->ctrl);
  CU(cudaGetLastError());
  CU(cudaEventRecord(e->ev[3], e->stream));
  e->sweep_no++;
  SweepOutDev h;
  CU(cudaMemcpyAsync(&h, e->out_dev, sizeof h, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->ms_prep, e->ev[0], e->ev[1]));
  CU(cudaEventElapsedTime(&e->ms_sweep, e->ev[1], e->ev[2]));
  CU(cudaEventElapsedTime(&e->ms_tail, e->ev[2], e->ev[3]));
  for (int k = 0; k < HB_MAX_FOLD; ++k) out->count[k] = h.count[k];
  out->varg_acc = h.varg_acc; out->sum_vargL = h.sum_vargL;
  out->sum_r = h.sum_r; out->sum_r2 = h.sum_r2; out->sum_u = h.sum_u; out->var_u = h.var_u;
  out->n_changed = h.n_changed; out->status = h.status; out->rounds = h.rounds; out->reserved = 0;
  if (e->trace && getenv("HB_TRACE")) {
    // event stamps of the tiles of this sweep -> file (ns): see tools/trace_report.py
    std::vector<unsigned long long> tr((size_t)e->T * 16);
    CU(cudaMemcpy(tr.data(), e->trace, tr.size() * 8, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("HB_TRACE"), "wb")) { fwrite(tr.data(), 8, tr.size(), f); fclose(f); }
  }
  if (getenv("HB_PHASES") && sp.serial) {
    static const char* nm[13] = {"wait_pkg", "rhs0", "-", "bar_open", "chain", "final", "publish", "tail", "bar_chain", "loop", "bar_part",
                                 "verify", "bar_bad"};
    fprintf(stderr, "[hb] serial CTA: tiles repaired %d, generic %lld, rounds %d (tiles %d)\n[hb phases serial]", h.pad,
            h.phase_clk[1][1], h.rounds, e->T);
    for (int k = 0; k < 13; ++k) fprintf(stderr, " %s=%.0f", nm[k], (double)h.phase_clk[0][k] / std::max(1, e->T));
    fprintf(stderr, " (cycles per tile)\n");
  } else if (getenv("HB_PHASES")) {
    fprintf(stderr, "[hb] tiles re-speculated before the chain: %d, rounds %d (tiles %d)\n", h.pad, h.rounds, e->T);
    static const char* nm[16] = {"wait_dots", "guess", "wait_prev", "bar_rhs0", "chain", "first", "post", "commit",
                                 "bar_chain", "loop", "bar_part", "classify", "bar_bad", "-", "-", "-"};
    for (int g = 0; g < 2; ++g) {
      fprintf(stderr, "[hb phases worker %d]", g);
      for (int k = 0; k < 13; ++k) fprintf(stderr, " %s=%.0f", nm[k], (double)h.phase_clk[g][k] / std::max(1, e->T / e->NG));
      fprintf(stderr, " (cycles per tile)\n");
    }
  }
  if (h.status != 0)
    return hb_set_error("sweep kernel aborted with device status %d (%s)", h.status,
                        h.status == HB_ABORT_OVERFLOW ? "fixed-point dot overflow" : "timeout waiting on a tile signal");
  return 0;
}

extern "C" int hb_engine_last_sweep_ms(hb_engine* e, float* a, float* b, float* c) {
  if (!e) return hb_set_error("null engine");
  if (a) *a = e->ms_prep;
  if (b) *b = e->ms_sweep;
  if (c) *c = e->ms_tail;
  return 0;
}

extern "C" int hb_engine_set_windows(hb_engine* e, const int32_t* windindx) {
  if (!e) return hb_set_error("null engine");
  CU(cudaSetDevice(e->cfg.device));
  cudaFree(e->wstart); cudaFree(e->wmem); cudaFree(e->wppa);
  e->wstart = e->wmem = nullptr; e->wppa = nullptr; e->nw = 0;
  if (!windindx) return 0;
  int nw = 0;
  for (int i = 0; i < e->m; ++i) nw = std::max(nw, (int)windindx[i]);
  if (nw <= 0) return 0;
  std::vector<int> ws(nw + 1, 0), wm(e->m), fill(nw, 0);
  for (int i = 0; i < e->m; ++i) if (windindx[i] >= 1) ws[windindx[i]]++;
  for (int w = 0; w < nw; ++w) ws[w + 1] += ws[w];
  for (int i = 0; i < e->m; ++i) if (windindx[i] >= 1) { int w = windindx[i] - 1; wm[ws[w] + fill[w]++] = i; }
  CU(cudaMalloc(&e->wstart, (nw + 1) * 4));
  CU(cudaMalloc(&e->wmem, (size_t)e->m * 4));
  CU(cudaMalloc(&e->wppa, (size_t)nw * 8));
  CU(cudaMemcpy(e->wstart, ws.data(), (nw + 1) * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->wmem, wm.data(), (size_t)e->m * 4, cudaMemcpyHostToDevice));
  CU(cudaMemset(e->wppa, 0, (size_t)nw * 8));
  e->nw = nw;
  return 0;
}
extern "C" int hb_engine_accumulate_pip(hb_engine* e) {
  if (!e) return hb_set_error("null engine");
  CU(cudaSetDevice(e->cfg.device));
  k_pip<<<(e->m + 255) / 256, 256, 0, e->stream>>>(e->m, e->tracker, e->nzrate);
  CU(cudaGetLastError());
  if (e->nw) {
    k_wppa<<<(e->nw + 127) / 128, 128, 0, e->stream>>>(e->nw, e->wstart, e->wmem, e->tracker, e->wppa);
    CU(cudaGetLastError());
  }
  return 0;
}
extern "C" int hb_engine_get_pip_counts(hb_engine* e, double* nzrate, double* wppa, int nw) {
  if (!e) return hb_set_error("null engine");
  CU(cudaSetDevice(e->cfg.device));
  if (nzrate) CU(cudaMemcpyAsync(nzrate, e->nzrate, (size_t)e->m * 8, cudaMemcpyDeviceToHost, e->stream));
  if (wppa && e->nw) {
    if (nw != e->nw) return hb_set_error("window count mismatch");
    CU(cudaMemcpyAsync(wppa, e->wppa, (size_t)nw * 8, cudaMemcpyDeviceToHost, e->stream));
  }
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}
extern "C" int hb_engine_accumulate_effects(hb_engine* e) {
  if (!e) return hb_set_error("null engine");
  CU(cudaSetDevice(e->cfg.device));
  k_axpy1<<<(e->m + 255) / 256, 256, 0, e->stream>>>(e->m, e->g, e->gsum);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int hb_engine_get_gram(hb_engine* e, int32_t* out) {
  if (!e || !out) return hb_set_error("hb_engine_get_gram: null argument");
  if (!e->gram_ready) return hb_set_error("hb_engine_get_gram: gram not built");
  CU(cudaSetDevice(e->cfg.device));
  CU(cudaMemcpy(out, e->gram, (size_t)e->T * e->D * e->B * e->B * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int hb_engine_predict(hb_engine* e, const double* alpha, double* out) {
  if (!e || !alpha || !out) return hb_set_error("hb_engine_predict: null argument");
  if (!e->geno_ready) return hb_set_error("hb_engine_predict: genotypes not loaded");
  CU(cudaSetDevice(e->cfg.device));
  const int nchunk = 32;
  double *da = nullptr, *dp = nullptr, *dout = nullptr;
  CU(cudaMalloc(&da, (size_t)e->m * 8));
  CU(cudaMalloc(&dp, (size_t)nchunk * e->Npad * 8));
  CU(cudaMalloc(&dout, (size_t)e->n * 8));
  CU(cudaMemcpyAsync(da, alpha, (size_t)e->m * 8, cudaMemcpyHostToDevice, e->stream));
  const size_t sh = (size_t)e->CL * e->R * 8;
  CU(cudaFuncSetAttribute(k_gemv_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  k_gemv_part<<<dim3(e->S, nchunk), e->NTCp, sh, e->stream>>>(e->Xp, da, e->m, e->R, e->NRG, e->CL, e->T, e->B, e->slab_stride,
                                                             e->Npad, dp);
  CU(cudaGetLastError());
  k_gemv_sum<<<(e->n + 255) / 256, 256, 0, e->stream>>>(dp, nchunk, e->Npad, e->n, dout);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout, (size_t)e->n * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(da); cudaFree(dp); cudaFree(dout);
  return 0;
}

// out[j] = sum_i x[i][j] v[i]  (X.t() * v, Bayes.cpp:961: the BSLMM polygenic values as SNP effects).  One warp per
// SNP walks the slabs in order (a slab's rows of a column are contiguous: R bytes), lane-strided partial sums, fixed
// shuffle tree: deterministic.  One pass over X, used once at the end of a run.
__global__ void k_xt_vec(const uint8_t* __restrict__ Xp, const double* __restrict__ v, int m, int S, int R, size_t slab_stride,
                         double* __restrict__ out) {
  const int j = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (j >= m) return;
  double s = 0.0;
  for (int sl = 0; sl < S; ++sl) {
    const uint8_t* col = Xp + (size_t)sl * slab_stride + (size_t)j * R;
    const double* vs = v + (size_t)sl * R;
    for (int r = 4 * lane; r < R; r += 128) {
      const uint32_t w = *(const uint32_t*)(col + r);
      s = fma((double)(w & 0xffu), vs[r], s);
      s = fma((double)((w >> 8) & 0xffu), vs[r + 1], s);
      s = fma((double)((w >> 16) & 0xffu), vs[r + 2], s);
      s = fma((double)(w >> 24), vs[r + 3], s);
    }
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
}
extern "C" int hb_engine_xt_vec(hb_engine* e, const double* v, double* out) {
  if (!e || !v || !out) return hb_set_error("hb_engine_xt_vec: null argument");
  if (!e->geno_ready) return hb_set_error("hb_engine_xt_vec: genotypes not loaded");
  CU(cudaSetDevice(e->cfg.device));
  struct Bufs { double *v = nullptr, *o = nullptr; ~Bufs() { cudaFree(v); cudaFree(o); } } b;
  CU(cudaMalloc(&b.v, e->Npad * 8));
  CU(cudaMalloc(&b.o, (size_t)e->m * 8));
  CU(cudaMemsetAsync(b.v, 0, e->Npad * 8, e->stream));   // padding rows: genotype 0 anyway
  CU(cudaMemcpyAsync(b.v, v, (size_t)e->n * 8, cudaMemcpyHostToDevice, e->stream));
  k_xt_vec<<<(unsigned)(((size_t)e->m * 32 + 255) / 256), 256, 0, e->stream>>>(e->Xp, b.v, e->m, e->S, e->R, e->slab_stride, b.o);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, b.o, (size_t)e->m * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return 0;
}

// Genetic values of every stored MCMC sample, `M %*% res$MCMCsamples$alpha` of R/bayes.r:303-304:
// out[:, c] = X alpha[:, c].  Blocks of up to 64 records go through k_gemm_samples, which reads X once per block.
extern "C" int hb_engine_predict_samples(hb_engine* e, const double* alpha, size_t ld_alpha, int n_records, double* out,
                                         size_t ld_out) {
  if (!e || !alpha || !out) return hb_set_error("hb_engine_predict_samples: null argument");
  if (n_records < 0 || ld_alpha < (size_t)e->m || ld_out < (size_t)e->n)
    return hb_set_error("hb_engine_predict_samples: bad leading dimension or record count");
  if (!e->geno_ready) return hb_set_error("hb_engine_predict_samples: genotypes not loaded");
  if (n_records == 0) return 0;
  CU(cudaSetDevice(e->cfg.device));
  const int RW = e->R / 128;   // R is 128, 256 or 384
  const int nchunk = std::max(1, std::min(e->m_pad / kGsSub, (e->nsm + e->S - 1) / e->S));
  struct Bufs { double *a = nullptr, *at = nullptr, *part = nullptr, *o = nullptr; ~Bufs() { cudaFree(a); cudaFree(at); cudaFree(part); cudaFree(o); } } b;
  const int blk = std::min(n_records, kGsRec);
  CU(cudaMalloc(&b.a, (size_t)e->m * blk * 8));
  CU(cudaMalloc(&b.at, (size_t)e->m_pad * kGsRec * 8));
  CU(cudaMalloc(&b.part, (size_t)nchunk * e->Npad * kGsRec * 8));
  CU(cudaMalloc(&b.o, (size_t)e->n * blk * 8));
  const size_t sh = 2 * ((size_t)kGsSub * e->R + (size_t)kGsSub * kGsRec * 8);
  CU(cudaFuncSetAttribute(k_gemm_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  float ms_total = 0.f;
  for (int r0 = 0; r0 < n_records; r0 += kGsRec) {
    const int nrec = std::min(kGsRec, n_records - r0), NG = (nrec + 15) / 16;
    CU(cudaMemcpy2DAsync(b.a, (size_t)e->m * 8, alpha + (size_t)r0 * ld_alpha, ld_alpha * 8, (size_t)e->m * 8, nrec, cudaMemcpyHostToDevice, e->stream));
    k_gs_transpose<<<(unsigned)(((size_t)e->m_pad * kGsRec + 255) / 256), 256, 0, e->stream>>>(b.a, (size_t)e->m, e->m, e->m_pad, nrec, b.at);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev[0], e->stream));
    k_gemm_samples<<<dim3(e->S, nchunk), 32 * RW * NG, sh, e->stream>>>(e->Xp, b.at, e->m_pad, e->R, RW, NG, e->slab_stride, e->Npad, b.part);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev[1], e->stream));
    k_gs_sum<<<(unsigned)(((size_t)e->n * nrec + 255) / 256), 256, 0, e->stream>>>(b.part, nchunk, e->Npad, e->n, nrec, b.o);
    CU(cudaGetLastError());
    CU(cudaMemcpy2DAsync(out + (size_t)r0 * ld_out, ld_out * 8, b.o, (size_t)e->n * 8, (size_t)e->n * 8, nrec, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
    ms_total += ms;
  }
  e->ms_predict = ms_total;
  return 0;
}
extern "C" int hb_engine_last_predict_ms(hb_engine* e, float* ms) {
  if (!e || !ms) return hb_set_error("hb_engine_last_predict_ms: null argument");
  *ms = e->ms_predict;
  return 0;
}

// ------------------------------------------------------------------------------------------
// row sharding over several GPUs (one process and one engine per GPU)
// ------------------------------------------------------------------------------------------
extern "C" int hb_engine_ipc_handle(hb_engine* e, void* handle64) {
  if (!e || !handle64) return hb_set_error("hb_engine_ipc_handle: null argument");
  if (!e->acc2) return hb_set_error("hb_engine_ipc_handle: engine was created with world = 1");
  CU(cudaSetDevice(e->cfg.device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, e->acc2));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int hb_engine_set_peers(hb_engine* e, const void* handles) {
  if (!e || !handles) return hb_set_error("hb_engine_set_peers: null argument");
  if (!e->acc2) return hb_set_error("hb_engine_set_peers: engine was created with world = 1");
  CU(cudaSetDevice(e->cfg.device));
  for (int g = 0; g < e->cfg.world; ++g) {
    if (g == e->cfg.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)g, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    e->peer_acc2[g] = (unsigned long long*)ptr;
  }
  e->peers_ready = true;
  return 0;
}
extern "C" int hb_engine_gram_device(hb_engine* e, void** ptr, uint64_t* count) {
  if (!e || !ptr || !count) return hb_set_error("hb_engine_gram_device: null argument");
  if (!e->gram_ready) return hb_set_error("hb_engine_gram_device: gram not built");
  *ptr = e->gram;
  *count = (uint64_t)e->T * e->D * e->B * e->B;
  return 0;
}
// sum over the local rows of (mean - u)^2 and (mean - u): the two accumulators of Armadillo's var(), Bayes.cpp:819
__global__ void __launch_bounds__(1024) k_centered(const double* __restrict__ u, int n, double mean, double* out2) {
  __shared__ double sh[1024];
  double e = 0, f = 0;
  for (int i = threadIdx.x; i < n; i += 1024) { double x = mean - u[i]; e += x * x; f += x; }
  e = block_sum_1024(e, sh); f = block_sum_1024(f, sh);
  if (threadIdx.x == 0) { out2[0] = e; out2[1] = f; }
}
extern "C" int hb_engine_u_centered_sums(hb_engine* e, double mean, double* ss, double* s1) {
  if (!e || !ss || !s1) return hb_set_error("hb_engine_u_centered_sums: null argument");
  CU(cudaSetDevice(e->cfg.device));
  k_centered<<<1, 1024, 0, e->stream>>>(e->u, e->n, mean, e->post_partial);
  CU(cudaGetLastError());
  double h[2];
  CU(cudaMemcpyAsync(h, e->post_partial, 16, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *ss = h[0]; *s1 = h[1];
  return 0;
}

// ------------------------------------------------------------------------------------------
// host-side access to the class-decision code of the sweep (the same source compiled for the host): lets the CPU
// tests check the certified thresholds against the exact evaluation without a GPU
// ------------------------------------------------------------------------------------------
extern "C" int hb_test_class_thresholds(int n_fold, double u, const double* a, const double* c, double logpi0, double* TL, double* TH) {
  if (n_fold < 2 || n_fold > HB_MAX_FOLD || !a || !c || !TL || !TH) return hb_set_error("hb_test_class_thresholds: bad argument");
  hbk::solve_thresholds<HB_MAX_FOLD>(n_fold, u, a, c, logpi0, TL, TH);
  return 0;
}
// The same two decisions ON THE DEVICE for `count` (rr, u) pairs of one SNP's parameters (test hook for the disagreement
// count against the reference's literal form, tests/test_class_decisions_gpu.py): by_threshold[i] = thr_class after
// solve_thresholds(u[i]) as k_prep / the sweep do it (-1 = inside a bracket), exact[i] = the soft-max evaluation they fall
// back to.
__global__ void k_test_class_batch(int nf, long long count, const double* __restrict__ rr, const double* __restrict__ u,
                                   const double* __restrict__ a, const double* __restrict__ c, double logpi0,
                                   int8_t* __restrict__ by_threshold, int8_t* __restrict__ exact) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double av[HB_MAX_FOLD], cv[HB_MAX_FOLD], TL[HB_MAX_FOLD], TH[HB_MAX_FOLD], cum[HB_MAX_FOLD];
  for (int k = 0; k < nf - 1; ++k) { av[k] = a[k]; cv[k] = c[k]; }
  hbk::solve_thresholds<HB_MAX_FOLD>(nf, u[i], av, cv, logpi0, TL, TH);
  by_threshold[i] = (int8_t)hbk::thr_class<HB_MAX_FOLD>(nf, rr[i], TL, TH);
  hbk::class_cum<HB_MAX_FOLD>(nf, rr[i], av, cv, logpi0, cum);
  exact[i] = (int8_t)hbk::class_from_cum<HB_MAX_FOLD>(nf, u[i], cum);
}
extern "C" int hb_test_class_batch_device(int device, int n_fold, long long count, const double* rr, const double* u, const double* a,
                                          const double* c, double logpi0, int8_t* by_threshold, int8_t* exact) {
  if (n_fold < 2 || n_fold > HB_MAX_FOLD || count <= 0 || !rr || !u || !a || !c || !by_threshold || !exact)
    return hb_set_error("hb_test_class_batch_device: bad argument");
  CU(cudaSetDevice(device));
  double *d_rr = nullptr, *d_u = nullptr, *d_a = nullptr, *d_c = nullptr;
  int8_t *d_t = nullptr, *d_e = nullptr;
  int rc = 0;
#define TRYT(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); goto done; } } while (0)
  TRYT(cudaMalloc(&d_rr, count * 8)); TRYT(cudaMalloc(&d_u, count * 8));
  TRYT(cudaMalloc(&d_a, HB_MAX_FOLD * 8)); TRYT(cudaMalloc(&d_c, HB_MAX_FOLD * 8));
  TRYT(cudaMalloc(&d_t, count)); TRYT(cudaMalloc(&d_e, count));
  TRYT(cudaMemcpy(d_rr, rr, count * 8, cudaMemcpyHostToDevice)); TRYT(cudaMemcpy(d_u, u, count * 8, cudaMemcpyHostToDevice));
  TRYT(cudaMemcpy(d_a, a, (n_fold - 1) * 8, cudaMemcpyHostToDevice)); TRYT(cudaMemcpy(d_c, c, (n_fold - 1) * 8, cudaMemcpyHostToDevice));
  k_test_class_batch<<<(unsigned)((count + 255) / 256), 256>>>(n_fold, count, d_rr, d_u, d_a, d_c, logpi0, d_t, d_e);
  TRYT(cudaGetLastError());
  TRYT(cudaMemcpy(by_threshold, d_t, count, cudaMemcpyDeviceToHost)); TRYT(cudaMemcpy(exact, d_e, count, cudaMemcpyDeviceToHost));
done:
#undef TRYT
  cudaFree(d_rr); cudaFree(d_u); cudaFree(d_a); cudaFree(d_c); cudaFree(d_t); cudaFree(d_e);
  return rc;
}
// class from the thresholds (-1 = inside a bracket) and from the exact cumulative probabilities
extern "C" int hb_test_class_of(int n_fold, double rr, double u, const double* a, const double* c, double logpi0, const double* TL,
                                const double* TH, int* by_threshold, int* exact) {
  if (n_fold < 2 || n_fold > HB_MAX_FOLD || !by_threshold || !exact) return hb_set_error("hb_test_class_of: bad argument");
  double cum[HB_MAX_FOLD];
  hbk::class_cum<HB_MAX_FOLD>(n_fold, rr, a, c, logpi0, cum);
  *exact = hbk::class_from_cum<HB_MAX_FOLD>(n_fold, u, cum);
  *by_threshold = hbk::thr_class<HB_MAX_FOLD>(n_fold, rr, TL, TH);
  return 0;
}
