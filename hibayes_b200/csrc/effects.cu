// effects.cu -- the non-SNP steps of an MCMC iteration of Bayes() on the device (sm_100a), so that the residual
// yadj and the genetic values u never leave HBM inside the loop.
//
// Reference (/root/reference/src):
//   covariates               Bayes.cpp:484-494   rhs = C_i'yadj + C_i'C_i b_i, draw, yadj += (b_i - b_i') C_i
//   env. random effects      Bayes.cpp:496-516   Z'yadj per level, one draw per level, yadj += Z (r - r')
//   single-step J            Bayes.cpp:555-562   as a covariate, u moves the other way
//   single-step epsilon      Bayes.cpp:563-582   RHS = Z'yadj.tail(ne) + Z'Z eps, sparse Gauss-Seidel sampler
//                            solver.cpp:131-140  Gibbs(sp_mat): x_i = N(x_i + (b_i - A_i.x)/a_ii, ve/a_ii) in order i = 0, 1, ...
//
// The vectors are the engine's own r and u (hb_engine_device_state).  Reductions return their value to the host
// driver (one small copy), which owns the draws of the scalar effects and, when individuals are sharded over ranks,
// all-reduces it; level sums and the epsilon right-hand side are exact per-level sums in the reference's order.
// The Gauss-Seidel sampler keeps the reference's sequential semantics: unknown i uses the new values of its neighbours
// j < i and the old ones of j > i.  Unknowns are grouped into wavefront levels (level(i) = 1 + max level of the
// neighbours below i in the symmetrised pattern), a level's unknowns are independent of each other, and one CTA runs
// the levels in order -- every unknown computed by one thread in the stored order of its column, with the reference's
// operation order, so the result is the sequential one bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/hibayes_b200.h"
#include "hb_rng.h"

int hb_set_error(const char* fmt, ...);
#define CU(call)                                                                             \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return hb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                          cudaGetErrorString(_e));                                           \
  } while (0)

namespace {

constexpr int kDotBlocks = 128, kDotThreads = 256;

// partial[b] = sum over a fixed strided pattern of a[i] * b[i]; the host adds the kDotBlocks partials in order
__global__ void k_fx_dot(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ partial) {
  __shared__ double sh[kDotThreads];
  double s = 0.0;
  for (int i = blockIdx.x * kDotThreads + threadIdx.x; i < n; i += kDotBlocks * kDotThreads) s = fma(a[i], b[i], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// r += ar * x, u += au * x  (x == nullptr: a vector of ones)
__global__ void k_fx_axpy(double* __restrict__ r, double* __restrict__ u, const double* __restrict__ x, double ar, double au, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xv = x ? x[i] : 1.0;
  if (ar != 0.0) r[i] = fma(ar, xv, r[i]);
  if (au != 0.0) u[i] = fma(au, xv, u[i]);
}

// sums[l] = sum of r over the rows of level l (rows listed in ascending order): one warp per level, lane-strided
// partial sums, fixed shuffle tree
__global__ void k_fx_level_sums(const double* __restrict__ r, const int* __restrict__ start, const int* __restrict__ rows, int nlev,
                                double* __restrict__ sums) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nlev) return;
  double s = 0.0;
  for (int p = start[w] + lane; p < start[w + 1]; p += 32) s += r[rows[p]];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) sums[w] = s;
}

// r[k] += diff[lev[k]]
__global__ void k_fx_level_apply(double* __restrict__ r, const int* __restrict__ lev, const double* __restrict__ diff, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) r[k] += diff[lev[k]];
}

// epsilon right-hand side, Bayes.cpp:567-568 first half: rhs[q] = sum of yadj.tail(ne) over the records of entry q, in
// record order (one thread per entry: the reference's order of additions)
__global__ void k_eps_rhs(const double* __restrict__ rtail, const int* __restrict__ start, const int* __restrict__ recs, int qe,
                          double* __restrict__ rhs) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qe) return;
  double s = 0.0;
  for (int p = start[q]; p < start[q + 1]; ++p) s += rtail[recs[p]];
  rhs[q] = s;
}

struct EpsDev {
  int qe, nlevels;
  const int *colptr, *rowidx;
  const double* val;
  const double* cnt;       // Z'Z diagonal: records per entry (all ranks)
  const int *order, *lstart;  // unknowns sorted by wavefront level; first unknown of every level
  double *x, *rhs;         // epsl_estR_tmp, RHS
};

// Gibbs(sp_mat A, x, b, ve), solver.cpp:131-140, with A = Z'Z + Gi * ratio (Bayes.cpp:565-566) never formed; one CTA
__global__ void __launch_bounds__(1024, 1) k_eps_gibbs(EpsDev d, double ratio, double vare, hb_key_t key, uint32_t it) {
  for (int q = threadIdx.x; q < d.qe; q += blockDim.x) d.rhs[q] += d.cnt[q] * d.x[q];   // + Z'Z eps (:568)
  __syncthreads();
  for (int l = 0; l < d.nlevels; ++l) {
    for (int s = d.lstart[l] + threadIdx.x; s < d.lstart[l + 1]; s += blockDim.x) {
      const int i = d.order[s];
      double aii = d.cnt[i], Ax = 0.0;
      bool have_diag = false;
      for (int p = d.colptr[i]; p < d.colptr[i + 1]; ++p) {
        const int rix = d.rowidx[p];
        const double aval = __dadd_rn(__dmul_rn(d.val[p], ratio), (rix == i ? d.cnt[i] : 0.0));
        if (rix == i) { aii = aval; have_diag = true; }
        Ax = __dadd_rn(Ax, __dmul_rn(aval, d.x[rix]));
      }
      if (!have_diag) Ax = __dadd_rn(Ax, __dmul_rn(d.cnt[i], d.x[i]));
      const double invlhs = 1.0 / aii;
      const double uu = __dadd_rn(__dmul_rn(invlhs, __dsub_rn(d.rhs[i], Ax)), d.x[i]);
      d.x[i] = __dadd_rn(uu, __dmul_rn(sqrt(__dmul_rn(invlhs, vare)), hb_draw_z(key, HB_DOM_EPS, it, (uint32_t)i, 0, 0)));
    }
    __syncthreads();
  }
}

// eps_old - eps_new goes to the records' rows (Bayes.cpp:572-576), then eps_old = eps_new
__global__ void k_eps_apply(double* __restrict__ rtail, double* __restrict__ utail, const int* __restrict__ index0, int ne,
                            const double* __restrict__ est, const double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const int q = index0[i];
  const double dlt = est[q] - x[q];
  rtail[i] += dlt;
  utail[i] -= dlt;
}

// colsum[c] * x[c] with colsum = Gi[:, c]' x in the stored order (Bayes.cpp:577); block partials in a fixed tree
__global__ void k_eps_quad(EpsDev d, double* __restrict__ partial) {
  __shared__ double sh[kDotThreads];
  double s = 0.0;
  for (int c = blockIdx.x * kDotThreads + threadIdx.x; c < d.qe; c += kDotBlocks * kDotThreads) {
    double colsum = 0.0;
    for (int p = d.colptr[c]; p < d.colptr[c + 1]; ++p) colsum += d.val[p] * d.x[d.rowidx[p]];
    s += colsum * d.x[c];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---- BSLMM (Bayes.cpp:518-552): dense products with the eigenvectors K (n x nk column-major)
// out[j] = K[:, j] . v : one warp per column, lane-strided partial sums, fixed shuffle tree
__global__ void k_gemv_t(const double* __restrict__ K, int n, int nk, const double* __restrict__ v, const double* __restrict__ v2,
                         double* __restrict__ out) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= nk) return;
  const double* col = K + (size_t)j * n;
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = fma(col[i], v2 ? v[i] + v2[i] : v[i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
}
// out[i] = sum_j K[i, j] w[j] : one thread per row, columns in order
__global__ void k_gemv_n(const double* __restrict__ K, int n, int nk, const double* __restrict__ w, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < nk; ++j) s = fma(K[(size_t)j * n + i], w[j], s);
  out[i] = s;
}
// w[j] = c1[j] * t[j] + c2[j] * z(iter, j)   (:532, :535)
__global__ void k_k_weights(const double* __restrict__ c1, const double* __restrict__ c2, const double* __restrict__ t, int nk, hb_key_t key,
                            uint32_t it, double* __restrict__ w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nk) w[j] = c1[j] * t[j] + c2[j] * hb_draw_z(key, HB_DOM_K, it, (uint32_t)j, 0, 0);
}
// yadj += k_old - k_new, u -= k_old - k_new, k_old = k_new   (:537-540, :551)
__global__ void k_k_apply(double* __restrict__ r, double* __restrict__ u, double* __restrict__ kcur, const double* __restrict__ knew, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = kcur[i] - knew[i];
  r[i] += d;
  u[i] -= d;
  kcur[i] = knew[i];
}
// partial sums of t[j]^2 / kval[j]
__global__ void k_k_quad(const double* __restrict__ t, const double* __restrict__ kval, int nk, double* __restrict__ partial) {
  __shared__ double sh[kDotThreads];
  double s = 0.0;
  for (int j = blockIdx.x * kDotThreads + threadIdx.x; j < nk; j += kDotBlocks * kDotThreads) s += t[j] * ((1.0 / kval[j]) * t[j]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_k_scale(double* __restrict__ t, const double* __restrict__ kval, double a, int nk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nk) t[j] = (t[j] / kval[j]) / a;
}
__global__ void k_vec_scale_to(const double* __restrict__ x, double a, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] / a;
}

__global__ void k_vec_add(double* __restrict__ acc, const double* __restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}

template <typename T>
int upload(T** dst, const T* src, size_t cnt, cudaStream_t st) {
  CU(cudaMalloc((void**)dst, std::max<size_t>(cnt, 1) * sizeof(T)));
  if (cnt) CU(cudaMemcpyAsync(*dst, src, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

}  // namespace

struct hb_fx {
  int device = 0, n = 0, nc = 0, nr = 0, ne = 0, qe = 0, n_levels = 0, eps_levels = 0;
  bool haveJ = false;
  cudaStream_t stream = nullptr;
  double *r = nullptr, *u = nullptr;          // the engine's vectors (borrowed)
  double *C = nullptr, *J = nullptr;
  int *lev = nullptr;                          // [nr][n] global level id (offset of the term + level) of every row
  int *lstart = nullptr, *lrows = nullptr;     // CSR over the global level ids: rows of every level, ascending
  std::vector<int> R_off;
  double *lsum = nullptr, *ldiff = nullptr;    // [n_levels]
  // epsilon
  int *e_index0 = nullptr, *e_start = nullptr, *e_recs = nullptr, *g_colptr = nullptr, *g_rowidx = nullptr, *e_order = nullptr,
      *e_lstart = nullptr;
  double *g_val = nullptr, *e_cnt = nullptr, *e_x = nullptr, *e_rhs = nullptr, *e_est = nullptr, *e_sum = nullptr;
  // BSLMM
  int nk = 0;
  double *K = nullptr, *k_cur = nullptr, *k_new = nullptr, *k_sum = nullptr, *k_t = nullptr, *k_w = nullptr, *k_c1 = nullptr, *k_c2 = nullptr,
         *k_val = nullptr;
  double* partial = nullptr;                   // kDotBlocks
  double* h_partial = nullptr;                 // pinned
  hb_key_t key;
};

extern "C" void hb_fx_destroy(hb_fx* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  void* p[] = {f->C, f->J, f->lev, f->lstart, f->lrows, f->lsum, f->ldiff, f->e_index0, f->e_start, f->e_recs, f->g_colptr,
               f->g_rowidx, f->e_order, f->e_lstart, f->g_val, f->e_cnt, f->e_x, f->e_rhs, f->e_est, f->e_sum, f->partial,
               f->K, f->k_cur, f->k_new, f->k_sum, f->k_t, f->k_w, f->k_c1, f->k_c2, f->k_val};
  for (void* q : p) if (q) cudaFree(q);
  if (f->h_partial) cudaFreeHost(f->h_partial);
  delete f;
}

extern "C" int hb_fx_create(hb_engine* e, const hb_fx_desc* d, hb_fx** out) {
  if (!e || !d || !out) return hb_set_error("hb_fx_create: null argument");
  hb_fx* f = new hb_fx;
  void* st = nullptr;
  int n_eng = 0;
  if (hb_engine_device_state(e, &f->r, &f->u, &st, &f->device, &n_eng) != 0) { delete f; return 1; }
  f->stream = (cudaStream_t)st;
  if (d->n != n_eng) { delete f; return hb_set_error("hb_fx_create: n = %d does not match the engine's %d rows", d->n, n_eng); }
  f->n = d->n; f->nc = d->nc; f->nr = d->nr; f->ne = d->ne; f->qe = d->Gi_colptr ? d->qe : 0; f->haveJ = d->J != nullptr;
  f->key = hb_make_key(d->seed);
  const int n = f->n;
  CU(cudaSetDevice(f->device));
  struct Guard { hb_fx* f; ~Guard() { if (f) hb_fx_destroy(f); } } guard{f};
  CU(cudaMalloc((void**)&f->partial, kDotBlocks * sizeof(double)));
  CU(cudaMallocHost((void**)&f->h_partial, kDotBlocks * sizeof(double)));
  if (f->nc) { if (upload(&f->C, d->C, (size_t)n * f->nc, f->stream)) return 1; }
  if (f->haveJ) { if (upload(&f->J, d->J, (size_t)n, f->stream)) return 1; }
  std::vector<int> lev, lstart, lrows;   // (alive until the synchronize below)
  if (f->nr) {
    f->R_off.assign(f->nr + 1, 0);
    for (int i = 0; i < f->nr; ++i) {
      if (d->nlev[i] <= 0) return hb_set_error("hb_fx_create: random term %d has no levels", i);
      f->R_off[i + 1] = f->R_off[i] + d->nlev[i];
    }
    f->n_levels = f->R_off[f->nr];
    lev.resize((size_t)f->nr * n);
    lstart.assign(f->n_levels + 1, 0);
    for (int i = 0; i < f->nr; ++i)
      for (int k = 0; k < n; ++k) {
        const int c = d->Rlev[(size_t)i * n + k];
        if (c < 0 || c >= d->nlev[i]) return hb_set_error("hb_fx_create: level code %d of random term %d out of range", c, i);
        lev[(size_t)i * n + k] = f->R_off[i] + c;
        ++lstart[f->R_off[i] + c + 1];
      }
    for (int q = 0; q < f->n_levels; ++q) lstart[q + 1] += lstart[q];
    lrows.resize((size_t)f->nr * n);
    std::vector<int> fill(lstart.begin(), lstart.end() - 1);
    for (int i = 0; i < f->nr; ++i)
      for (int k = 0; k < n; ++k) lrows[fill[lev[(size_t)i * n + k]]++] = k;
    if (upload(&f->lev, lev.data(), lev.size(), f->stream) || upload(&f->lstart, lstart.data(), lstart.size(), f->stream) ||
        upload(&f->lrows, lrows.data(), lrows.size(), f->stream)) return 1;
    CU(cudaMalloc((void**)&f->lsum, f->n_levels * sizeof(double)));
    CU(cudaMalloc((void**)&f->ldiff, f->n_levels * sizeof(double)));
  }
  std::vector<int> index0, estart, erecs, order, elstart;
  if (f->qe) {
    const int qe = f->qe, ne = f->ne;
    if (!d->Gi_rowidx || !d->Gi_val || (ne && !d->epsl_index) || qe <= 0 || ne > n || ne < 0)
      return hb_set_error("hb_fx_create: incomplete single-step term");
    index0.resize(ne);
    estart.assign(qe + 1, 0);
    for (int i = 0; i < ne; ++i) {
      const int q = d->epsl_index[i] - 1;   // Bayes.cpp:255-256
      if (q < 0 || q >= qe) return hb_set_error("hb_fx_create: epsl_index[%d] = %d outside 1..%d", i, d->epsl_index[i], qe);
      index0[i] = q;
      ++estart[q + 1];
    }
    for (int q = 0; q < qe; ++q) estart[q + 1] += estart[q];
    erecs.resize(ne);
    { std::vector<int> fill(estart.begin(), estart.end() - 1); for (int i = 0; i < ne; ++i) erecs[fill[index0[i]]++] = i; }
    // wavefront levels over the symmetrised pattern of Gi
    const int nnz = d->Gi_colptr[qe];
    std::vector<int> lowcnt(qe + 1, 0);
    for (int c = 0; c < qe; ++c)
      for (int p = d->Gi_colptr[c]; p < d->Gi_colptr[c + 1]; ++p) {
        const int r = d->Gi_rowidx[p];
        if (r < 0 || r >= qe) return hb_set_error("hb_fx_create: row index %d of epsl_Gi out of range", r);
        if (r != c) ++lowcnt[std::max(r, c) + 1];
      }
    for (int q = 0; q < qe; ++q) lowcnt[q + 1] += lowcnt[q];
    std::vector<int> low(std::max(nnz, 1)), fill(lowcnt.begin(), lowcnt.end() - 1);
    for (int c = 0; c < qe; ++c)
      for (int p = d->Gi_colptr[c]; p < d->Gi_colptr[c + 1]; ++p) {
        const int r = d->Gi_rowidx[p];
        if (r != c) low[fill[std::max(r, c)]++] = std::min(r, c);
      }
    std::vector<int> lvl(qe, 0);
    int nl = 0;
    for (int i = 0; i < qe; ++i) {
      int l = 0;
      for (int p = lowcnt[i]; p < lowcnt[i + 1]; ++p) l = std::max(l, lvl[low[p]] + 1);
      lvl[i] = l;
      nl = std::max(nl, l + 1);
    }
    f->eps_levels = nl;
    elstart.assign(nl + 1, 0);
    for (int i = 0; i < qe; ++i) ++elstart[lvl[i] + 1];
    for (int l = 0; l < nl; ++l) elstart[l + 1] += elstart[l];
    order.resize(qe);
    { std::vector<int> fl(elstart.begin(), elstart.end() - 1); for (int i = 0; i < qe; ++i) order[fl[lvl[i]]++] = i; }
    if (upload(&f->e_index0, index0.data(), index0.size(), f->stream) || upload(&f->e_start, estart.data(), estart.size(), f->stream) ||
        upload(&f->e_recs, erecs.data(), erecs.size(), f->stream) || upload(&f->g_colptr, d->Gi_colptr, (size_t)qe + 1, f->stream) ||
        upload(&f->g_rowidx, d->Gi_rowidx, (size_t)nnz, f->stream) || upload(&f->g_val, d->Gi_val, (size_t)nnz, f->stream) ||
        upload(&f->e_order, order.data(), order.size(), f->stream) || upload(&f->e_lstart, elstart.data(), elstart.size(), f->stream))
      return 1;
    double** z[] = {&f->e_cnt, &f->e_x, &f->e_rhs, &f->e_est, &f->e_sum};
    for (double** q : z) {
      CU(cudaMalloc((void**)q, qe * sizeof(double)));
      CU(cudaMemsetAsync(*q, 0, qe * sizeof(double), f->stream));
    }
  }
  if (d->nk > 0) {
    if (!d->Ki || d->nk != n) return hb_set_error("hb_fx_create: Ki must be n x n (nk = %d, n = %d)", d->nk, n);
    f->nk = d->nk;
    if (cudaMalloc((void**)&f->K, (size_t)n * f->nk * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      return hb_set_error("hb_fx_create: the %d x %d eigenvector matrix (%.1f GB) does not fit on the device", n, f->nk, (double)n * f->nk * 8 / 1e9);
    }
    CU(cudaMemcpyAsync(f->K, d->Ki, (size_t)n * f->nk * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    double** z[] = {&f->k_cur, &f->k_new, &f->k_sum, &f->k_t, &f->k_w, &f->k_c1, &f->k_c2, &f->k_val};
    for (double** q : z) {
      CU(cudaMalloc((void**)q, (size_t)n * sizeof(double)));
      CU(cudaMemsetAsync(*q, 0, (size_t)n * sizeof(double), f->stream));
    }
  }
  CU(cudaStreamSynchronize(f->stream));
  guard.f = nullptr;
  *out = f;
  return 0;
}

static int finish_partials(hb_fx* f, double* out) {
  CU(cudaMemcpyAsync(f->h_partial, f->partial, kDotBlocks * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  double s = 0.0;
  for (int b = 0; b < kDotBlocks; ++b) s += f->h_partial[b];
  *out = s;
  return 0;
}

static const double* fx_vector(hb_fx* f, int kind, int idx) {
  if (kind == HB_FX_COV) return (idx >= 0 && idx < f->nc) ? f->C + (size_t)idx * f->n : nullptr;
  if (kind == HB_FX_J) return f->J;
  if (kind == HB_FX_RESID) return f->r;
  return nullptr;
}

extern "C" int hb_fx_dot(hb_fx* f, int kind, int idx, double* out) {
  if (!f || !out) return hb_set_error("hb_fx_dot: null argument");
  const double* x = fx_vector(f, kind, idx);
  if (!x) return hb_set_error("hb_fx_dot: no such vector (kind %d, index %d)", kind, idx);
  CU(cudaSetDevice(f->device));
  k_fx_dot<<<kDotBlocks, kDotThreads, 0, f->stream>>>(x, f->r, f->n, f->partial);
  CU(cudaGetLastError());
  return finish_partials(f, out);
}

extern "C" int hb_fx_self_dot(hb_fx* f, int kind, int idx, double* out) {
  if (!f || !out) return hb_set_error("hb_fx_self_dot: null argument");
  const double* x = fx_vector(f, kind, idx);
  if (!x) return hb_set_error("hb_fx_self_dot: no such vector (kind %d, index %d)", kind, idx);
  CU(cudaSetDevice(f->device));
  k_fx_dot<<<kDotBlocks, kDotThreads, 0, f->stream>>>(x, x, f->n, f->partial);
  CU(cudaGetLastError());
  return finish_partials(f, out);
}

extern "C" int hb_fx_axpy(hb_fx* f, int kind, int idx, double a_r, double a_u) {
  if (!f) return hb_set_error("hb_fx_axpy: null argument");
  const double* x = nullptr;
  if (kind != HB_FX_ONES) {
    x = fx_vector(f, kind, idx);
    if (!x || kind == HB_FX_RESID) return hb_set_error("hb_fx_axpy: no such vector (kind %d, index %d)", kind, idx);
  }
  CU(cudaSetDevice(f->device));
  k_fx_axpy<<<(f->n + 255) / 256, 256, 0, f->stream>>>(f->r, f->u, x, a_r, a_u, f->n);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int hb_fx_level_sums(hb_fx* f, int term, double* sums) {
  if (!f || !sums || term < 0 || term >= f->nr) return hb_set_error("hb_fx_level_sums: bad argument");
  const int off = f->R_off[term], q = f->R_off[term + 1] - off;
  CU(cudaSetDevice(f->device));
  k_fx_level_sums<<<(q * 32 + 255) / 256, 256, 0, f->stream>>>(f->r, f->lstart + off, f->lrows, q, f->lsum + off);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(sums, f->lsum + off, q * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return 0;
}

extern "C" int hb_fx_level_apply(hb_fx* f, int term, const double* diff) {
  if (!f || !diff || term < 0 || term >= f->nr) return hb_set_error("hb_fx_level_apply: bad argument");
  const int off = f->R_off[term], q = f->R_off[term + 1] - off;
  CU(cudaSetDevice(f->device));
  CU(cudaMemcpyAsync(f->ldiff + off, diff, q * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  k_fx_level_apply<<<(f->n + 255) / 256, 256, 0, f->stream>>>(f->r, f->lev + (size_t)term * f->n, f->ldiff, f->n);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(f->stream));   // `diff` may be reused by the caller
  return 0;
}

extern "C" int hb_fx_eps_set_counts(hb_fx* f, const double* cnt) {
  if (!f || !cnt || !f->qe) return hb_set_error("hb_fx_eps_set_counts: bad argument");
  CU(cudaSetDevice(f->device));
  CU(cudaMemcpyAsync(f->e_cnt, cnt, f->qe * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return 0;
}

extern "C" int hb_fx_eps_rhs(hb_fx* f, double* rhs_host) {
  if (!f || !f->qe) return hb_set_error("hb_fx_eps_rhs: no single-step term");
  CU(cudaSetDevice(f->device));
  k_eps_rhs<<<(f->qe + 255) / 256, 256, 0, f->stream>>>(f->r + (f->n - f->ne), f->e_start, f->e_recs, f->qe, f->e_rhs);
  CU(cudaGetLastError());
  if (rhs_host) {
    CU(cudaMemcpyAsync(rhs_host, f->e_rhs, f->qe * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
  }
  return 0;
}

extern "C" int hb_fx_eps_set_rhs(hb_fx* f, const double* rhs_host) {
  if (!f || !f->qe || !rhs_host) return hb_set_error("hb_fx_eps_set_rhs: bad argument");
  CU(cudaSetDevice(f->device));
  CU(cudaMemcpyAsync(f->e_rhs, rhs_host, f->qe * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return 0;
}

extern "C" int hb_fx_eps_sample(hb_fx* f, int iter, double vare, double ratio, double* quad) {
  if (!f || !f->qe || !quad) return hb_set_error("hb_fx_eps_sample: bad argument");
  CU(cudaSetDevice(f->device));
  EpsDev d;
  d.qe = f->qe; d.nlevels = f->eps_levels; d.colptr = f->g_colptr; d.rowidx = f->g_rowidx; d.val = f->g_val; d.cnt = f->e_cnt;
  d.order = f->e_order; d.lstart = f->e_lstart; d.x = f->e_x; d.rhs = f->e_rhs;
  k_eps_gibbs<<<1, 1024, 0, f->stream>>>(d, ratio, vare, f->key, (uint32_t)iter);
  CU(cudaGetLastError());
  if (f->ne) {
    k_eps_apply<<<(f->ne + 255) / 256, 256, 0, f->stream>>>(f->r + (f->n - f->ne), f->u + (f->n - f->ne), f->e_index0, f->ne, f->e_est, f->e_x);
    CU(cudaGetLastError());
  }
  k_eps_quad<<<kDotBlocks, kDotThreads, 0, f->stream>>>(d, f->partial);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(f->e_est, f->e_x, f->qe * sizeof(double), cudaMemcpyDeviceToDevice, f->stream));   // :580
  return finish_partials(f, quad);
}

extern "C" int hb_fx_eps_accumulate(hb_fx* f) {
  if (!f || !f->qe) return hb_set_error("hb_fx_eps_accumulate: no single-step term");
  CU(cudaSetDevice(f->device));
  k_vec_add<<<(f->qe + 255) / 256, 256, 0, f->stream>>>(f->e_sum, f->e_est, f->qe);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int hb_fx_eps_get(hb_fx* f, double* est, double* sum) {
  if (!f || !f->qe) return hb_set_error("hb_fx_eps_get: no single-step term");
  CU(cudaSetDevice(f->device));
  if (est) CU(cudaMemcpyAsync(est, f->e_est, f->qe * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  if (sum) CU(cudaMemcpyAsync(sum, f->e_sum, f->qe * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return 0;
}

extern "C" int hb_fx_describe(hb_fx* f, int* eps_levels) {
  if (!f) return hb_set_error("hb_fx_describe: null argument");
  if (eps_levels) *eps_levels = f->eps_levels;
  return 0;
}

extern "C" int hb_fx_k_step(hb_fx* f, int iter, const double* c1, const double* c2, const double* kival, double* quad) {
  if (!f || !f->nk || !c1 || !c2 || !kival || !quad) return hb_set_error("hb_fx_k_step: bad argument");
  CU(cudaSetDevice(f->device));
  const int n = f->n, nk = f->nk;
  CU(cudaMemcpyAsync(f->k_c1, c1, nk * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  CU(cudaMemcpyAsync(f->k_c2, c2, nk * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  CU(cudaMemcpyAsync(f->k_val, kival, nk * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  const unsigned gw = (unsigned)(((size_t)nk * 32 + 255) / 256), gn = (unsigned)((n + 255) / 256), gk = (unsigned)((nk + 255) / 256);
  k_gemv_t<<<gw, 256, 0, f->stream>>>(f->K, n, nk, f->r, f->k_cur, f->k_t);                 // K'(yadj + k_old)   :519, :532
  k_k_weights<<<gk, 256, 0, f->stream>>>(f->k_c1, f->k_c2, f->k_t, nk, f->key, (uint32_t)iter, f->k_w);
  k_gemv_n<<<gn, 256, 0, f->stream>>>(f->K, n, nk, f->k_w, f->k_new);                       // K (...)            :532, :535
  k_k_apply<<<gn, 256, 0, f->stream>>>(f->r, f->u, f->k_cur, f->k_new, n);                  // :537-540, :551
  k_gemv_t<<<gw, 256, 0, f->stream>>>(f->K, n, nk, f->k_cur, nullptr, f->k_t);              // Kg = K' k_new      :543
  k_k_quad<<<kDotBlocks, kDotThreads, 0, f->stream>>>(f->k_t, f->k_val, nk, f->partial);    // :544
  CU(cudaGetLastError());
  return finish_partials(f, quad);
}

extern "C" int hb_fx_k_accumulate(hb_fx* f) {
  if (!f || !f->nk) return hb_set_error("hb_fx_k_accumulate: no polygenic term");
  CU(cudaSetDevice(f->device));
  k_vec_add<<<(f->nk + 255) / 256, 256, 0, f->stream>>>(f->k_sum, f->k_cur, f->nk);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int hb_fx_k_ghat_vec(hb_fx* f, const double* kival, double sumvx, double count, double* v_host) {
  if (!f || !f->nk || !kival || !v_host) return hb_set_error("hb_fx_k_ghat_vec: bad argument");
  CU(cudaSetDevice(f->device));
  const int n = f->n, nk = f->nk;
  const unsigned gw = (unsigned)(((size_t)nk * 32 + 255) / 256), gn = (unsigned)((n + 255) / 256), gk = (unsigned)((nk + 255) / 256);
  CU(cudaMemcpyAsync(f->k_val, kival, nk * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  k_vec_scale_to<<<gk, 256, 0, f->stream>>>(f->k_sum, count, f->k_new, nk);           // k_estR_store /= count  :956
  k_gemv_t<<<gw, 256, 0, f->stream>>>(f->K, n, nk, f->k_new, nullptr, f->k_t);              // K' k_mean              :957
  k_k_scale<<<gk, 256, 0, f->stream>>>(f->k_t, f->k_val, sumvx, nk);                  // / Kval / sumvx         :958-959
  k_gemv_n<<<gn, 256, 0, f->stream>>>(f->K, n, nk, f->k_t, f->k_w);                         // K * Kg                 :961
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(v_host, f->k_w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return 0;
}
