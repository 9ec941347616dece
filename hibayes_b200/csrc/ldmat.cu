// ldmat.cu -- the two stages in front of the Gibbs sweeps on the device:
//
//   * hb_ldmat_*      the LD / X'X builder: tXXmat_Geno() and tXXmat_Chr() of
//                     /root/reference/src/tXXmat.cpp:100-185, 504-605 with BigStat() (:43-77).  The
//                     reference's O(m^2 n) scalar triple loop becomes an exact int8 x int8 -> int32
//                     tensor-core Gram (mma.sync m16n8k32; products of genotypes are small integers,
//                     so the inner product is exact) with the centring, the r^2 n <= chisq filter
//                     and the fp64 conversion fused into the epilogue, evaluated in the reference's
//                     operation order (so every off-diagonal entry is bit-identical to the oracle).
//   * hb_bed_decode   read_bed<char>() of /root/reference/src/read_bed.cpp:97-232 (hb_bed.cuh).
//
// Data layout: Xc[Mpad][Kpad] int8, one row per SNP, individuals contiguous (K-major for both mma
// operands), zero padded to Kpad = 128-multiple of n and Mpad = 64-multiple of m.  The m x m result is
// produced in column panels of W columns (fp64, column-major, <= 1 GiB) that are either copied to the
// caller's dense matrix or compacted on the device into CSC (dgCMatrix) pieces.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/hibayes_b200.h"
#include "hb_bed.cuh"
#include "hb_device.cuh"

int hb_set_error(const char* fmt, ...);  // engine.cu
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                          cudaGetErrorString(_e));                                                \
  } while (0)

struct hb_ldmat {
  int device = 0, n = 0, m = 0, Kpad = 0, Mpad = 0, panel_cols = 0;
  int8_t* Xc = nullptr;
  double *sum = nullptr, *mean = nullptr, *xx = nullptr;
  int32_t* chr = nullptr;
  double* pan = nullptr;
  size_t pan_elems = 0;
  bool loaded = false, stats_ready = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float ms_gram = 0.f;
  CUtensorMap tmap;          // Xc as a 2-D tensor (individuals x SNP rows) for the TMA loads of k_ld_panel_tc
  bool tmap_ready = false;
  int use_tc = 1;            // tcgen05 panel kernel (default); HB_LD_MMA_SYNC=1 keeps the mma.sync kernel
  bool panel_user = false;   // the caller fixed the panel width (hb_ldmat_set_panel_cols)
  int* status = nullptr;     // device word: a bounded wait of k_ld_panel_tc gave up
  std::vector<long long> colptr;
  std::vector<int32_t> rowidx;
  std::vector<double> val;
  bool sparse_ready = false;
};

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// fp64 operations that must round exactly where the reference's expressions round: explicit round-to-nearest
// intrinsics on the device (no FMA contraction), plain operators on the host (built with -ffp-contract=off)
#ifdef __CUDA_ARCH__
#define LD_MUL(a, b) __dmul_rn((a), (b))
#define LD_ADD(a, b) __dadd_rn((a), (b))
#define LD_SUB(a, b) __dsub_rn((a), (b))
#define LD_DIV(a, b) __ddiv_rn((a), (b))
#else
#define LD_MUL(a, b) ((a) * (b))
#define LD_ADD(a, b) ((a) + (b))
#define LD_SUB(a, b) ((a) - (b))
#define LD_DIV(a, b) ((a) / (b))
#endif

#ifdef __CUDA_ARCH__
#define LD_SQRT(a) __dsqrt_rn(a)
#else
#define LD_SQRT(a) sqrt(a)
#endif

// BigStat (tXXmat.cpp:43-77) of one SNP row of Xc, individuals in file order: the second pass is a sequential fp64
// sum of (x - mean)^2 whose rounding depends on the order, and the sparse branch thresholds on it (:142-143), so the
// order is kept.  The first pass sums integers (exact).  __host__ __device__: the CPU tests run this function too.
__host__ __device__ __forceinline__ void ld_stats_row(const int8_t* row, int Kpad, int n, double* sum, double* mean, double* xx) {
  long long s = 0;
  for (int k0 = 0; k0 < Kpad; k0 += 16) {  // padding bytes are 0
    const int4 v = *(const int4*)(row + k0);
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) s += (int8_t)(w[q]) + (int8_t)(w[q] >> 8) + (int8_t)(w[q] >> 16) + (int8_t)(w[q] >> 24);
  }
  const double sm = (double)s;
  const double mu = LD_DIV(sm, (double)n);
  double p1 = 0.0;
  for (int k0 = 0; k0 < n; k0 += 16) {
    const int4 v = *(const int4*)(row + k0);
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (k0 + q < n) {
        const double dv = LD_SUB((double)(int8_t)(w[q >> 2] >> (8 * (q & 3))), mu);
        p1 = LD_ADD(p1, LD_MUL(dv, dv));
      }
    }
  }
  *sum = sm;
  *mean = mu;
  *xx = LD_SQRT(p1);
}

// one thread per SNP
__global__ void k_ld_stats(const int8_t* __restrict__ Xc, int Kpad, int n, int m, double* __restrict__ sum,
                           double* __restrict__ mean, double* __restrict__ xx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  ld_stats_row(Xc + (size_t)j * Kpad, Kpad, n, sum + j, mean + j, xx + j);
}

struct LdEpi {
  const double *sum, *mean, *xx;
  const int32_t* chr;  // may be null
  int n, m, has_chisq;
  double chisq;
};

// One LD entry from the exact inner product (tXXmat.cpp:141-148 / :157): the statistics of the SNP
// with the smaller index play the role of sum1/m1/p1 (outer loop variable j of the reference).
// __host__ __device__: tests/ run this very function on the CPU (hb_test_ld_entries below).
__host__ __device__ __forceinline__ double ld_entry(const LdEpi& e, int i, int j, int G) {
  if (e.chr && e.chr[i] != e.chr[j]) return 0.0;
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  const double ind = (double)e.n;
  const double p1 = e.xx[lo];
  if (!e.has_chisq && lo == hi) return LD_DIV(LD_MUL(p1, p1), ind);
  const double m1 = e.mean[lo], sum1 = e.sum[lo];
  const double p2 = e.xx[hi], m2 = e.mean[hi], sum2 = e.sum[hi];
  const double t = LD_SUB(LD_ADD(LD_MUL(sum1, m2), LD_MUL(sum2, m1)), LD_MUL(LD_MUL(ind, m1), m2));
  const double p12 = LD_SUB((double)G, t);
  if (e.has_chisq) {
    const double r = LD_DIV(p12, LD_MUL(p1, p2));
    if (LD_MUL(LD_MUL(r, r), ind) <= e.chisq) return 0.0;
  }
  return LD_DIV(p12, ind);
}

constexpr int LD_RK = 128;              // individuals (bytes) per stage
constexpr int LD_STRIDE = LD_RK + 16;   // padded shared-memory row: conflict-free ldmatrix
__device__ __forceinline__ void ld_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hb::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ld_ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(hb::smem_u32(p)));
}
__device__ __forceinline__ void ld_imma_s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CTA = 64 x 64 entries: rows [64 bx, +64) x panel columns [j0 + 64 by, +64); 4 warps of 32 x 32;
// 128 individuals per stage through a double-buffered cp.async pipeline, fragments by ldmatrix (the
// structure of k_gram_imma in engine.cu, which builds the sweep's band Gram).  pan is column-major
// with leading dimension m.
__global__ void __launch_bounds__(128) k_ld_panel(const int8_t* __restrict__ Xc, int Kpad, int j0, LdEpi epi,
                                                  double* __restrict__ pan) {
  __shared__ __align__(16) uint8_t As[2][64 * LD_STRIDE];
  __shared__ __align__(16) uint8_t Bs[2][64 * LD_STRIDE];
  const int a0 = blockIdx.x * 64;
  const int b0 = j0 + blockIdx.y * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  int acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0;
  const int nchunks = Kpad / LD_RK;
  auto issue = [&](int ch, int buf) {
    const int8_t* Abase = Xc + (size_t)a0 * Kpad + (size_t)ch * LD_RK;
    const int8_t* Bbase = Xc + (size_t)b0 * Kpad + (size_t)ch * LD_RK;
    // 64 SNP rows x 8 vectors of 16 B per operand = 512 vectors; 128 threads -> 4 + 4 each
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = tid + 128 * q;
      const int c = v >> 3, w = v & 7;
      ld_cp_async16(&As[buf][c * LD_STRIDE + 16 * w], Abase + (size_t)c * Kpad + 16 * w);
      ld_cp_async16(&Bs[buf][c * LD_STRIDE + 16 * w], Bbase + (size_t)c * Kpad + 16 * w);
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
    for (int ks = 0; ks < LD_RK / 32; ++ks) {
      uint32_t af[2][4], bf[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = wm + 16 * i + (lane & 7) + 8 * ((lane >> 3) & 1);
        ld_ldmatrix_x4(af[i], &As[buf][row * LD_STRIDE + 32 * ks + 16 * (lane >> 4)]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = wn + 16 * j + (lane & 7) + 8 * (lane >> 4);
        ld_ldmatrix_x4(bf[j], &Bs[buf][row * LD_STRIDE + 32 * ks + 16 * ((lane >> 3) & 1)]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ld_imma_s8(acc[i][j], af[i], bf[j >> 1][2 * (j & 1)], bf[j >> 1][2 * (j & 1) + 1]);
    }
    __syncthreads();
  }
  // epilogue: accumulator element (row, col) per the m16n8 C layout
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = a0 + wm + 16 * i + (lane >> 2) + 8 * (k >> 1);
        const int col = b0 + wn + 8 * j + 2 * (lane & 3) + (k & 1);
        if (row < epi.m && col < epi.m) pan[(size_t)(col - j0) * epi.m + row] = ld_entry(epi, row, col, acc[i][j][k]);
      }
}

// ---- the same panel on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulator in tensor memory) ----
// CTA = 128 x 128 entries: SNP rows [128 bx, +128) x panel columns [j0 + 128 by, +128).  Both operands are K-major
// tiles of Xc (128 SNP rows x 128 individuals = 16 KB) that the TMA unit (cp.async.bulk.tensor.2d, 128-byte swizzle)
// drops into a 4-stage shared-memory ring; one elected thread issues four tcgen05.mma (M = 128, N = 128, K = 32, s8 x s8 ->
// s32, exact) per stage into a 128-lane x 128-column accumulator in tensor memory and commits the stage back to the
// producer; four epilogue warps read their 32 lanes with tcgen05.ld (32 columns at a time) and turn every exact inner
// product into the reference's fp64 value with ld_entry() -- the same function, the same bits, as the mma.sync kernel.
// Warp roles: 0 TMA producer, 1 tensor-memory allocation + MMA issue, 2-9 epilogue (lane quadrant = warp % 4, two warps
// per quadrant with 64 columns each; the fp64 epilogue -- two IEEE divisions per entry -- is what a tile spends its time on).
// Every wait is bounded (2 s): a lost signal sets *status instead of hanging the device.
constexpr int TC_TILE = 128, TC_KB = 128, TC_STAGES = 3;   // 3 x 32 KB: two CTAs per SM, one's MMAs under the other's epilogue
constexpr uint32_t TC_STAGE_BYTES = 2u * TC_TILE * TC_KB;   // A + B
struct TcShared {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full;
  uint32_t tmem_base;
  int dead;
};
__device__ __forceinline__ unsigned long long tc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity, volatile int* dead) {
  if (hb::mbar_try_wait(bar, parity)) return true;
  const unsigned long long t0 = tc_now();
  for (;;) {
    for (int i = 0; i < 64; ++i)
      if (hb::mbar_try_wait(bar, parity)) return true;
    if (*dead || tc_now() - t0 > 2000000000ull) { *dead = 1; return false; }
  }
}
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  // K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (stride byte offset),
  // descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B; the leading byte offset is not used by this layout
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(320, 2) k_ld_panel_tc(const __grid_constant__ CUtensorMap tmap, int Kpad, int j0, LdEpi epi,
                                                        double* __restrict__ pan, int* __restrict__ status, int sym) {
  // sym: the panel is the whole matrix -- only the tiles on and below the diagonal are computed (tXXmat.cpp:161-182 loops
  // over i <= j and assigns both entries); a tile below the diagonal also writes its mirror image
  if (sym && blockIdx.x < blockIdx.y) return;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)tc_smem + 1023) & ~(uintptr_t)1023);
  TcShared* sh = (TcShared*)(tiles + (size_t)TC_STAGES * TC_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = blockIdx.x * TC_TILE, b0 = j0 + blockIdx.y * TC_TILE;
  const int nchunks = Kpad / TC_KB;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { hb::mbar_init(&sh->full[i], 1); hb::mbar_init(&sh->empty[i], 1); }
    hb::mbar_init(&sh->acc_full, 1);
    sh->dead = 0;
    hb::mbar_fence_init();
  }
  if (warp == 1) {   // 128 columns of tensor memory: the 128 x 128 s32 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hb::smem_u32(&sh->tmem_base)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sh->tmem_base;
  volatile int* dead = &sh->dead;
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      for (int ch = 0; ch < nchunks; ++ch) {
        const int st = ch % TC_STAGES;
        if (ch >= TC_STAGES && !tc_wait(&sh->empty[st], (uint32_t)((ch / TC_STAGES - 1) & 1), dead)) break;
        hb::mbar_arrive_expect_tx(&sh->full[st], TC_STAGE_BYTES);
        uint8_t* A = tiles + (size_t)st * TC_STAGE_BYTES;
        uint8_t* B = A + (size_t)TC_TILE * TC_KB;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(hb::smem_u32(A)), "l"(&tmap), "r"(ch * TC_KB), "r"(a0), "r"(hb::smem_u32(&sh->full[st])) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(hb::smem_u32(B)), "l"(&tmap), "r"(ch * TC_KB), "r"(b0), "r"(hb::smem_u32(&sh->full[st])) : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = s32 (2 at bit 4), A and B signed 8-bit (1 at bits 7 and 10), both K-major, N / 8 at
      // bit 17, M / 16 at bit 24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_TILE >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
      bool ok = true;
      for (int ch = 0; ch < nchunks && ok; ++ch) {
        const int st = ch % TC_STAGES;
        ok = tc_wait(&sh->full[st], (uint32_t)((ch / TC_STAGES) & 1), dead);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t A = hb::smem_u32(tiles + (size_t)st * TC_STAGE_BYTES), B = A + TC_TILE * TC_KB;
#pragma unroll
        for (int k = 0; k < TC_KB / 32; ++k) {
          const uint64_t da = tc_smem_desc(A + 32 * k), db = tc_smem_desc(B + 32 * k);
          const uint32_t accum = (ch | k) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
        // frees the stage for the producer once the four MMAs above have read it
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hb::smem_u32(&sh->empty[st])) : "memory");
      }
      if (ok) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hb::smem_u32(&sh->acc_full)) : "memory");
    }
  } else {
    // ---- epilogue: thread = accumulator lane = SNP row; 32 panel columns per tcgen05.ld
    const bool ok = tc_wait(&sh->acc_full, 0u, dead);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
      const int quad = warp & 3, half = (warp - 2) >> 2;
      const int row = a0 + 32 * quad + lane;
#pragma unroll 1
      for (int cb = 2 * half; cb < 2 * half + 2; ++cb) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(32 * cb);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
            "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < epi.m) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = b0 + 32 * cb + i;
            if (col < epi.m) {
              const double val = ld_entry(epi, row, col, (int)v[i]);
              pan[(size_t)(col - j0) * epi.m + row] = val;
              if (sym && blockIdx.x > blockIdx.y) pan[(size_t)row * epi.m + col] = val;   // (j0 = 0)
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  if (threadIdx.x == 0 && sh->dead) *status = 1;
}

// Stored entries per panel column: what an arma::sp_mat keeps is every assigned value != 0.
__global__ void k_ld_count(const double* __restrict__ pan, int m, int wcols, int* __restrict__ counts) {
  const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= wcols) return;
  const double* col = pan + (size_t)w * m;
  int c = 0;
  for (int i = lane; i < m; i += 32) c += (col[i] != 0.0) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[w] = c;
}
// Ordered compaction of one panel column per warp (row indices ascending, as in a dgCMatrix).
__global__ void k_ld_compact(const double* __restrict__ pan, int m, int wcols, const long long* __restrict__ off,
                             int32_t* __restrict__ rowidx, double* __restrict__ val) {
  const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= wcols) return;
  const double* col = pan + (size_t)w * m;
  long long base = off[w];
  for (int i0 = 0; i0 < m; i0 += 32) {
    const int i = i0 + lane;
    const double v = (i < m) ? col[i] : 0.0;
    const bool keep = v != 0.0;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const long long pos = base + __popc(mask & ((1u << lane) - 1u));
      rowidx[pos] = i;
      val[pos] = v;
    }
    base += __popc(mask);
  }
}

// .bed -> Xc rows: one thread per 16 individuals of one SNP (bytes beyond n stay 0).
__global__ void k_bed_to_rows(const uint8_t* __restrict__ bed, size_t bps, const int32_t* __restrict__ rows, int n, int col0,
                              int ncols, int d, int impt, int na, const uint8_t* __restrict__ info, int8_t* __restrict__ Xc,
                              int Kpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per_col = (size_t)Kpad / 16;
  if (idx >= per_col * (size_t)ncols) return;
  const int c = (int)(idx / per_col);
  const size_t row0 = (idx % per_col) * 16;
  const uint8_t* src = bed + (size_t)c * bps;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const size_t row = row0 + i;
    if (row < (size_t)n) w[i >> 2] |= (uint32_t)(uint8_t)hb::bed_value(src, rows, row, d, impt, na, info[c]) << (8 * (i & 3));
  }
  *(uint4*)(Xc + (size_t)(col0 + c) * Kpad + row0) = make_uint4(w[0], w[1], w[2], w[3]);
}

// .bed -> plain column-major int8 (the "char" big.matrix read_bed() fills), one thread per genotype.
__global__ void k_bed_to_colmajor(const uint8_t* __restrict__ bed, size_t bps, int nid, int ncols, int d, int impt, int na,
                                  const uint8_t* __restrict__ info, int8_t* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nid * (size_t)ncols) return;
  const int c = (int)(idx / (size_t)nid);
  const size_t row = idx % (size_t)nid;
  out[idx] = (int8_t)hb::bed_value(bed + (size_t)c * bps, nullptr, row, d, impt, na, info[c]);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static int check_bed_image(const uint8_t* file, size_t len, long long nid, long long m, size_t* bps_out) {
  if (!file) return hb_set_error("bed: null file image");
  if (nid <= 0 || m <= 0) return hb_set_error("bed: need at least one individual and one SNP");
  const size_t bps = (size_t)((nid + 3) / 4);
  if (len < 3 || file[0] != 0x6c || file[1] != 0x1b) return hb_set_error("bed: not a PLINK .bed image (magic bytes)");
  if (file[2] != 0x01) return hb_set_error("bed: individual-major .bed files are not supported (third byte must be 0x01)");
  if (len < 3 + bps * (size_t)m) return hb_set_error("bed: image has %zu bytes, %zu needed for %lld individuals x %lld SNPs", len, 3 + bps * (size_t)m, nid, m);
  *bps_out = bps;
  return 0;
}

extern "C" int hb_ldmat_create(int device, int n, int m, hb_ldmat** out) {
  if (!out) return hb_set_error("hb_ldmat_create: null out");
  *out = nullptr;
  if (n <= 0 || m <= 0) return hb_set_error("hb_ldmat_create: n and m must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return hb_set_error("hb_ldmat_create: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return hb_set_error("hb_ldmat_create: device %d out of range", device);
  CU(cudaSetDevice(device));
  hb_ldmat* h = new hb_ldmat();
  h->device = device;
  h->n = n;
  h->m = m;
  h->Kpad = (n + LD_RK - 1) / LD_RK * LD_RK;
  h->Mpad = (m + 127) / 128 * 128;
  // panel of at most 1 GiB of fp64
  long long W = ((1ll << 27) / m) / 128 * 128;
  h->panel_cols = (int)std::min<long long>(h->Mpad, std::max<long long>(128, W));
#define LDCK(call)                 \
  do {                             \
    if ((call) != cudaSuccess) {   \
      cudaError_t _e = cudaGetLastError(); \
      hb_ldmat_destroy(h);         \
      return hb_set_error("hb_ldmat_create: %s", cudaGetErrorString(_e)); \
    }                              \
  } while (0)
  LDCK(cudaStreamCreate(&h->stream));
  LDCK(cudaEventCreate(&h->ev0));
  LDCK(cudaEventCreate(&h->ev1));
  LDCK(cudaMalloc(&h->Xc, (size_t)h->Mpad * h->Kpad));
  LDCK(cudaMalloc(&h->sum, (size_t)m * 8));
  LDCK(cudaMalloc(&h->mean, (size_t)m * 8));
  LDCK(cudaMalloc(&h->xx, (size_t)m * 8));
  LDCK(cudaMalloc(&h->chr, (size_t)m * 4));
  LDCK(cudaMalloc(&h->status, 4));
  LDCK(cudaMemset(h->status, 0, 4));
#undef LDCK
  if (const char* ev = getenv("HB_LD_MMA_SYNC")) h->use_tc = atoi(ev) ? 0 : 1;
  if (h->use_tc) {
    // Xc[Mpad][Kpad] as a 2-D tensor: dimension 0 = individuals (bytes, contiguous), dimension 1 = SNP rows
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      hb_ldmat_destroy(h);
      return hb_set_error("hb_ldmat_create: cuTensorMapEncodeTiled not available from the driver");
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)h->Kpad, (cuuint64_t)h->Mpad};
    const cuuint64_t gstr[1] = {(cuuint64_t)h->Kpad};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_TILE};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = ((PFN_cuTensorMapEncodeTiled_v12000)fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, h->Xc, gdim, gstr, box, estr,
                                                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      hb_ldmat_destroy(h);
      return hb_set_error("hb_ldmat_create: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    }
    h->tmap_ready = true;
    const size_t shb = (size_t)TC_STAGES * TC_STAGE_BYTES + sizeof(TcShared) + 1024;
    if (cudaFuncSetAttribute(k_ld_panel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb) != cudaSuccess) {
      cudaError_t _e = cudaGetLastError();
      hb_ldmat_destroy(h);
      return hb_set_error("hb_ldmat_create: %s", cudaGetErrorString(_e));
    }
  }
  *out = h;
  return 0;
}

extern "C" void hb_ldmat_destroy(hb_ldmat* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->Xc); cudaFree(h->sum); cudaFree(h->mean); cudaFree(h->xx); cudaFree(h->chr); cudaFree(h->pan); cudaFree(h->status);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int hb_ldmat_set_panel_cols(hb_ldmat* h, int cols) {
  if (!h) return hb_set_error("hb_ldmat_set_panel_cols: null handle");
  if (cols < 64 || cols % 64) return hb_set_error("hb_ldmat_set_panel_cols: need a positive multiple of 64");
  cols = (cols + 127) / 128 * 128;   // (tiles of the tcgen05 kernel)
  h->panel_cols = std::min(cols, h->Mpad);
  h->panel_user = true;
  return 0;
}

extern "C" int hb_ldmat_load_i8(hb_ldmat* h, const int8_t* X, size_t ld) {
  if (!h || !X) return hb_set_error("hb_ldmat_load_i8: null argument");
  if (ld < (size_t)h->n) return hb_set_error("hb_ldmat_load_i8: ld < n");
  // the int32 accumulators hold sum |x_i x_j| only while max|x|^2 n < 2^31
  int amax = 0;
  for (int j = 0; j < h->m; ++j) {
    const int8_t* col = X + (size_t)j * ld;
    for (int i = 0; i < h->n; ++i) amax = std::max(amax, std::abs((int)col[i]));
  }
  if ((double)amax * amax * h->n >= 2147483648.0)
    return hb_set_error("hb_ldmat_load_i8: |x| up to %d over %d individuals overflows the exact int32 inner product", amax, h->n);
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->Xc, 0, (size_t)h->Mpad * h->Kpad, h->stream));
  CU(cudaMemcpy2DAsync(h->Xc, (size_t)h->Kpad, X, ld, (size_t)h->n, (size_t)h->m, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->loaded = true;
  h->stats_ready = false;
  h->sparse_ready = false;
  return 0;
}

extern "C" int hb_ldmat_load_bed(hb_ldmat* h, const uint8_t* file, size_t len, int nid, const int32_t* rows, int impt,
                                 int dominance) {
  if (!h) return hb_set_error("hb_ldmat_load_bed: null handle");
  size_t bps = 0;
  if (check_bed_image(file, len, nid, h->m, &bps)) return 1;
  if (!rows && nid != h->n) return hb_set_error("hb_ldmat_load_bed: the file has %d individuals, the handle %d; pass the row selection", nid, h->n);
  if (!impt && (double)h->n * 16384.0 >= 2147483648.0)
    return hb_set_error("hb_ldmat_load_bed: without imputation a missing genotype is -128, which overflows the exact int32 inner product over %d individuals", h->n);
  if (rows)
    for (int i = 0; i < h->n; ++i)
      if (rows[i] < 0 || rows[i] >= nid) return hb_set_error("hb_ldmat_load_bed: rows[%d] = %d outside the file's %d individuals", i, rows[i], nid);
  CU(cudaSetDevice(h->device));
  const int na = -128;  // NA_CHAR of bigmemory (read_bed.cpp:241)
  const size_t cols_per_chunk = std::min<size_t>((size_t)h->m, std::max<size_t>(1, (size_t)(256u << 20) / bps));
  uint8_t *stage = nullptr, *info = nullptr;
  int32_t* drows = nullptr;
  int rc = 0;
  auto fail = [&](cudaError_t e, int line) {
    rc = hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e), __FILE__, line, cudaGetErrorString(e));
  };
#define TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { fail(_e, __LINE__); goto done; } } while (0)
  TRY(cudaMalloc(&stage, cols_per_chunk * bps));
  TRY(cudaMalloc(&info, cols_per_chunk));
  if (rows) {
    TRY(cudaMalloc(&drows, (size_t)h->n * 4));
    TRY(cudaMemcpyAsync(drows, rows, (size_t)h->n * 4, cudaMemcpyHostToDevice, h->stream));
  }
  TRY(cudaMemsetAsync(h->Xc, 0, (size_t)h->Mpad * h->Kpad, h->stream));
  for (size_t c0 = 0; c0 < (size_t)h->m; c0 += cols_per_chunk) {
    const size_t nc = std::min(cols_per_chunk, (size_t)h->m - c0);
    TRY(cudaMemcpyAsync(stage, file + 3 + c0 * bps, nc * bps, cudaMemcpyHostToDevice, h->stream));
    hb::k_bed_info<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, h->stream>>>(stage, bps, nid, (int)nc, dominance, info);
    const size_t work = (size_t)(h->Kpad / 16) * nc;
    k_bed_to_rows<<<(unsigned)((work + 255) / 256), 256, 0, h->stream>>>(stage, bps, drows, h->n, (int)c0, (int)nc, dominance, impt,
                                                                        na, info, h->Xc, h->Kpad);
    TRY(cudaGetLastError());
    TRY(cudaStreamSynchronize(h->stream));
  }
  h->loaded = true;
  h->stats_ready = false;
  h->sparse_ready = false;
done:
#undef TRY
  cudaFree(stage); cudaFree(info); cudaFree(drows);
  return rc;
}

static int ensure_stats(hb_ldmat* h) {
  if (!h->loaded) return hb_set_error("hb_ldmat: genotypes not loaded");
  if (h->stats_ready) return 0;
  CU(cudaSetDevice(h->device));
  k_ld_stats<<<(h->m + 127) / 128, 128, 0, h->stream>>>(h->Xc, h->Kpad, h->n, h->m, h->sum, h->mean, h->xx);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->stats_ready = true;
  return 0;
}

extern "C" int hb_ldmat_stats(hb_ldmat* h, double* mean, double* sum, double* xx) {
  if (!h) return hb_set_error("hb_ldmat_stats: null handle");
  if (ensure_stats(h)) return 1;
  if (mean) CU(cudaMemcpy(mean, h->mean, (size_t)h->m * 8, cudaMemcpyDeviceToHost));
  if (sum) CU(cudaMemcpy(sum, h->sum, (size_t)h->m * 8, cudaMemcpyDeviceToHost));
  if (xx) CU(cudaMemcpy(xx, h->xx, (size_t)h->m * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// Runs the panels; per panel calls sink(j0, wcols) with the panel (column-major, ld = m) complete on the stream.
template <class Sink>
static int run_panels(hb_ldmat* h, const int32_t* chr, int has_chisq, double chisq, Sink sink) {
  if (ensure_stats(h)) return 1;
  CU(cudaSetDevice(h->device));
  if (chr) CU(cudaMemcpyAsync(h->chr, chr, (size_t)h->m * 4, cudaMemcpyHostToDevice, h->stream));
  // With the tcgen05 kernel the whole matrix is one panel when it fits (<= 64 GB and half of the free memory): only the
  // lower triangle is then computed, each tile writing its mirror image too (HB_LD_FULL=0 keeps the column panels).
  int W = h->panel_cols;
  int sym = 0;
  if (h->use_tc && !h->panel_user && !(getenv("HB_LD_FULL") && !atoi(getenv("HB_LD_FULL")))) {
    size_t fr = 0, tot = 0;
    const size_t full = (size_t)h->Mpad * h->m * 8;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && full <= ((size_t)64 << 30) && full <= (fr + h->pan_elems * 8) / 2) { W = h->Mpad; sym = 1; }
  }
  const size_t need = (size_t)W * h->m;
  if (h->pan_elems < need) {
    cudaFree(h->pan);
    h->pan = nullptr;
    h->pan_elems = 0;
    CU(cudaMalloc(&h->pan, need * 8));
    h->pan_elems = need;
  }
  LdEpi epi{h->sum, h->mean, h->xx, chr ? h->chr : nullptr, h->n, h->m, has_chisq, chisq};
  h->ms_gram = 0.f;
  for (int j0 = 0; j0 < h->m; j0 += W) {
    const int wpad = std::min(W, h->Mpad - j0);
    const int wcols = std::min(W, h->m - j0);
    CU(cudaEventRecord(h->ev0, h->stream));
    if (h->use_tc) {
      const size_t shb = (size_t)TC_STAGES * TC_STAGE_BYTES + sizeof(TcShared) + 1024;
      k_ld_panel_tc<<<dim3(h->Mpad / TC_TILE, wpad / TC_TILE), 320, shb, h->stream>>>(h->tmap, h->Kpad, j0, epi, h->pan, h->status, sym);
    } else {
      k_ld_panel<<<dim3(h->Mpad / 64, wpad / 64), 128, 0, h->stream>>>(h->Xc, h->Kpad, j0, epi, h->pan);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, h->stream));
    if (sink(j0, wcols)) return 1;
    if (h->use_tc) {
      int st = 0;
      CU(cudaMemcpyAsync(&st, h->status, 4, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      if (st) return hb_set_error("hb_ldmat: the tcgen05 panel kernel gave up waiting on a pipeline barrier");
    }
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->ms_gram += ms;
  }
  return 0;
}

extern "C" int hb_ldmat_dense(hb_ldmat* h, const int32_t* chr, int has_chisq, double chisq, double* out, size_t ldo) {
  if (!h || !out) return hb_set_error("hb_ldmat_dense: null argument");
  if (ldo < (size_t)h->m) return hb_set_error("hb_ldmat_dense: ldo < m");
  if (has_chisq && !(chisq >= 0)) return hb_set_error("hb_ldmat_dense: chisq must be >= 0");
  return run_panels(h, chr, has_chisq, chisq, [&](int j0, int wcols) -> int {
    CU(cudaMemcpy2DAsync(out + (size_t)j0 * ldo, ldo * 8, h->pan, (size_t)h->m * 8, (size_t)h->m * 8, (size_t)wcols,
                         cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  });
}

extern "C" int hb_ldmat_sparse(hb_ldmat* h, const int32_t* chr, int has_chisq, double chisq, long long* nnz) {
  if (!h) return hb_set_error("hb_ldmat_sparse: null handle");
  if (has_chisq && !(chisq >= 0)) return hb_set_error("hb_ldmat_sparse: chisq must be >= 0");
  h->sparse_ready = false;
  h->colptr.assign((size_t)h->m + 1, 0);
  h->rowidx.clear();
  h->val.clear();
  int* dcounts = nullptr;
  long long* doff = nullptr;
  int32_t* drow = nullptr;
  double* dval = nullptr;
  size_t cap = 0;
  CU(cudaSetDevice(h->device));
  CU(cudaMalloc(&dcounts, (size_t)h->panel_cols * 4));
  if (cudaMalloc(&doff, (size_t)h->panel_cols * 8) != cudaSuccess) { cudaFree(dcounts); return hb_set_error("hb_ldmat_sparse: out of device memory"); }
  std::vector<int> counts((size_t)h->panel_cols);
  std::vector<long long> off((size_t)h->panel_cols);
  int rc = run_panels(h, chr, has_chisq, chisq, [&](int j0, int wcols) -> int {
    const unsigned blocks = (unsigned)(((size_t)wcols * 32 + 255) / 256);
    k_ld_count<<<blocks, 256, 0, h->stream>>>(h->pan, h->m, wcols, dcounts);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts.data(), dcounts, (size_t)wcols * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    long long tot = 0;
    for (int w = 0; w < wcols; ++w) {
      off[w] = tot;
      tot += counts[w];
      h->colptr[(size_t)j0 + w + 1] = counts[w];
    }
    if (tot == 0) return 0;
    if ((size_t)tot > cap) {
      cudaFree(drow); cudaFree(dval);
      drow = nullptr; dval = nullptr; cap = 0;
      CU(cudaMalloc(&drow, (size_t)tot * 4));
      CU(cudaMalloc(&dval, (size_t)tot * 8));
      cap = (size_t)tot;
    }
    CU(cudaMemcpyAsync(doff, off.data(), (size_t)wcols * 8, cudaMemcpyHostToDevice, h->stream));
    k_ld_compact<<<blocks, 256, 0, h->stream>>>(h->pan, h->m, wcols, doff, drow, dval);
    CU(cudaGetLastError());
    const size_t old = h->rowidx.size();
    h->rowidx.resize(old + (size_t)tot);
    h->val.resize(old + (size_t)tot);
    CU(cudaMemcpyAsync(h->rowidx.data() + old, drow, (size_t)tot * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->val.data() + old, dval, (size_t)tot * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  });
  cudaFree(dcounts); cudaFree(doff); cudaFree(drow); cudaFree(dval);
  if (rc) return rc;
  for (int j = 0; j < h->m; ++j) h->colptr[(size_t)j + 1] += h->colptr[(size_t)j];
  h->sparse_ready = true;
  if (nnz) *nnz = h->colptr[(size_t)h->m];
  return 0;
}

extern "C" int hb_ldmat_sparse_get(hb_ldmat* h, long long* colptr, int32_t* rowidx, double* val) {
  if (!h || !colptr || !rowidx || !val) return hb_set_error("hb_ldmat_sparse_get: null argument");
  if (!h->sparse_ready) return hb_set_error("hb_ldmat_sparse_get: call hb_ldmat_sparse first");
  memcpy(colptr, h->colptr.data(), ((size_t)h->m + 1) * sizeof(long long));
  if (!h->rowidx.empty()) {
    memcpy(rowidx, h->rowidx.data(), h->rowidx.size() * 4);
    memcpy(val, h->val.data(), h->val.size() * 8);
  }
  return 0;
}

extern "C" int hb_ldmat_last_ms(hb_ldmat* h, float* gram_ms) {
  if (!h || !gram_ms) return hb_set_error("hb_ldmat_last_ms: null argument");
  *gram_ms = h->ms_gram;
  return 0;
}

// read_bed() into the caller's "char" big.matrix memory (nid x m column-major), miss = per-SNP flag.
extern "C" int hb_bed_decode(int device, const uint8_t* file, size_t len, int nid, int m, int impt, int dominance,
                             int8_t* out, uint8_t* miss) {
  if (!out) return hb_set_error("hb_bed_decode: null out");
  size_t bps = 0;
  if (check_bed_image(file, len, nid, m, &bps)) return 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return hb_set_error("hb_bed_decode: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return hb_set_error("hb_bed_decode: device %d out of range", device);
  CU(cudaSetDevice(device));
  const int na = -128;
  const size_t cols_per_chunk = std::min<size_t>((size_t)m, std::max<size_t>(1, (size_t)(64u << 20) / bps));
  uint8_t *stage = nullptr, *info = nullptr;
  int8_t* dout = nullptr;
  std::vector<uint8_t> hinfo(cols_per_chunk);
  int rc = 0;
#define TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); goto done; } } while (0)
  TRY(cudaMalloc(&stage, cols_per_chunk * bps));
  TRY(cudaMalloc(&info, cols_per_chunk));
  TRY(cudaMalloc(&dout, cols_per_chunk * (size_t)nid));
  for (size_t c0 = 0; c0 < (size_t)m; c0 += cols_per_chunk) {
    const size_t nc = std::min(cols_per_chunk, (size_t)m - c0);
    TRY(cudaMemcpy(stage, file + 3 + c0 * bps, nc * bps, cudaMemcpyHostToDevice));
    hb::k_bed_info<<<(unsigned)((nc * 32 + 255) / 256), 256>>>(stage, bps, nid, (int)nc, dominance, info);
    const size_t work = nc * (size_t)nid;
    k_bed_to_colmajor<<<(unsigned)((work + 255) / 256), 256>>>(stage, bps, nid, (int)nc, dominance, impt, na, info, dout);
    TRY(cudaGetLastError());
    TRY(cudaMemcpy(out + c0 * (size_t)nid, dout, work, cudaMemcpyDeviceToHost));
    if (miss) {
      TRY(cudaMemcpy(hinfo.data(), info, nc, cudaMemcpyDeviceToHost));
      for (size_t c = 0; c < nc; ++c) miss[c0 + c] = hinfo[c] >> 7;
    }
  }
done:
#undef TRY
  cudaFree(stage); cudaFree(info); cudaFree(dout);
  return rc;
}

// ------------------------------------------------------------------------------------------
// host builds of the byte-level decode code, for the CPU tests (no device needed)
// ------------------------------------------------------------------------------------------
// Decodes one SNP's byte row with the very functions the kernels use (bed_count_byte, bed_major,
// bed_field, bed_code): out[row] for `n` output rows, *info = major | missing << 7.
extern "C" int hb_test_bed_decode_snp(const uint8_t* snp_bytes, int nid, const int32_t* rows, int n, int impt, int dominance,
                                      int8_t* out, uint8_t* info_out) {
  if (!snp_bytes || !out || nid <= 0 || n < 0) return hb_set_error("hb_test_bed_decode_snp: bad argument");
  unsigned c[4] = {0u, 0u, 0u, 0u};
  const size_t full = (size_t)nid >> 2;
  for (size_t b = 0; b < full; ++b) hb::bed_count_byte(snp_bytes[b], 4, c);
  if (nid & 3) hb::bed_count_byte(snp_bytes[full], nid & 3, c);
  const unsigned long long t[4] = {c[0], c[1], c[2], c[3]};
  const uint8_t info = (uint8_t)(hb::bed_major(t, dominance) | (t[1] ? 0x80 : 0));
  for (int row = 0; row < n; ++row) {
    const size_t id = rows ? (size_t)rows[row] : (size_t)row;
    const unsigned f = hb::bed_field(snp_bytes, id);
    out[row] = (int8_t)(f == 1u ? (impt ? (int)(info & 0x7f) : -128) : hb::bed_code(f, dominance, -128));
  }
  if (info_out) *info_out = info;
  return 0;
}

// The epilogue's arithmetic on the host: out (m x m column-major) from an exact Gram matrix (int32, m x m) and the
// column statistics, through ld_entry() -- the function k_ld_panel calls per accumulator element.
extern "C" int hb_test_ld_entries(int n, int m, const int32_t* gram, const double* sum, const double* mean, const double* xx,
                                  const int32_t* chr, int has_chisq, double chisq, double* out) {
  if (!gram || !sum || !mean || !xx || !out || n <= 0 || m <= 0) return hb_set_error("hb_test_ld_entries: bad argument");
  LdEpi e{sum, mean, xx, chr, n, m, has_chisq, chisq};
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) out[(size_t)j * m + i] = ld_entry(e, i, j, gram[(size_t)j * m + i]);
  return 0;
}

// BigStat on the host through ld_stats_row(): Xc is m rows of Kpad bytes (16-byte aligned, zero padded beyond n).
extern "C" int hb_test_ld_stats(const int8_t* Xc, int Kpad, int n, int m, double* sum, double* mean, double* xx) {
  if (!Xc || !sum || !mean || !xx || Kpad % 16 || Kpad < n || ((uintptr_t)Xc & 15)) return hb_set_error("hb_test_ld_stats: bad argument");
  for (int j = 0; j < m; ++j) ld_stats_row(Xc + (size_t)j * Kpad, Kpad, n, sum + j, mean + j, xx + j);
  return 0;
}

// ------------------------------------------------------------------------------------------
// host build of hb_limbs.h (a round-2 building block of the sweep's inner loop; CPU tests only)
// ------------------------------------------------------------------------------------------
#include "hb_limbs.h"
// dot of one genotype column (n bytes in {0,1,2}, n a multiple of 4) with r through the limb scheme: returns x'q (exact
// int64) with q_i = rint(r_i scale); *ok = 0 if some |q_i| >= 2^47.
extern "C" int hb_test_limb_dot(const uint8_t* x, const double* r, int n, double scale, long long* dot_q, int* ok) {
  if (!x || !r || !dot_q || !ok || n % 4) return hb_set_error("hb_test_limb_dot: bad argument");
  std::vector<uint8_t> lim((size_t)HB_NLIMB * n);   // limb-major: lim[k][row]
  *ok = 1;
  for (int i = 0; i < n; ++i) {
    uint8_t l6[HB_NLIMB];
    if (!hb_limb_split(r[i], scale, l6)) { *ok = 0; l6[0] = l6[1] = l6[2] = l6[3] = l6[4] = l6[5] = 0; }
    for (int k = 0; k < HB_NLIMB; ++k) lim[(size_t)k * n + i] = l6[k];
  }
  int acc[HB_NLIMB] = {0, 0, 0, 0, 0, 0};
  long long total = 0;
  for (int i = 0; i < n; i += 4) {
    uint32_t xw;
    memcpy(&xw, x + i, 4);
    for (int k = 0; k < HB_NLIMB; ++k) {
      uint32_t lw;
      memcpy(&lw, &lim[(size_t)k * n + i], 4);
      acc[k] = hb_limb_dp4a(xw, lw, k, acc[k]);
    }
    if ((i / 4) % 6 == 5 || i + 4 >= n) {   // merge every 24 rows, as a lane of the sweep kernel would
      total += hb_limb_merge(acc);
      for (int k = 0; k < HB_NLIMB; ++k) acc[k] = 0;
    }
  }
  *dot_q = total;
  return 0;
}
