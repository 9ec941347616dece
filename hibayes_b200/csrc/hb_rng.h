/*
 * hb_rng.h -- the position-addressed random-number contract shared by the
 * device kernels, the C++ host driver and the CPU oracle.
 *
 * The reference consumes ONE sequential libR stream (unif_rand / norm_rand /
 * R::rgamma / R::rchisq, /root/reference/src/stats.cpp:3-28,55-76), which makes
 * the order of draws part of its result and forces a strictly serial scalar
 * phase.  This engine replaces "order" by "position": every draw the reference
 * makes is given a fixed address (domain, iteration, index, slot, attempt) and
 * is produced by Philox4x32-10 keyed with a 64-bit run key.  Any thread on any
 * GPU (and the CPU oracle) can therefore evaluate any draw independently and
 * gets the same value.  SURVEY.md Appendix A lists the reference draw sites the
 * addresses below correspond to.
 *
 * Plain C99, also compiled by nvcc as __host__ __device__ code.
 */
#ifndef HB_RNG_H
#define HB_RNG_H

#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD static inline
#endif

/* ---- draw addresses ------------------------------------------------------ */
/* domains */
#define HB_DOM_ITER 0u  /* per-iteration scalars, index = HB_IT_*            */
#define HB_DOM_SNP  1u  /* per-SNP draws in the sweep, index = SNP j          */
#define HB_DOM_COV  2u  /* covariate effects, index = covariate c             */
#define HB_DOM_RAND 3u  /* env. random-effect levels, index = global level id */
#define HB_DOM_EPS  4u  /* single-step epsilon entries, index = entry         */
#define HB_DOM_K    5u  /* BSLMM polygenic randn, index = eigen direction     */
/* HB_DOM_ITER indices */
#define HB_IT_MU      0u  /* intercept normal            Bayes.cpp:480 */
#define HB_IT_VARG    1u  /* marker variance chi^2       Bayes.cpp:603,713,807 */
#define HB_IT_VARE    2u  /* residual variance chi^2     Bayes.cpp:823 */
#define HB_IT_LAMBDA  3u  /* BayesL lambda^2 gamma       Bayes.cpp:740 */
#define HB_IT_J       4u  /* single-step J normal        Bayes.cpp:559 */
#define HB_IT_VEPS    5u  /* single-step Veps chi^2      Bayes.cpp:579 */
#define HB_IT_VB      6u  /* BSLMM vb chi^2              Bayes.cpp:547 */
#define HB_IT_VARA    7u  /* SBayes genetic var chi^2    SBayesD.cpp:461 */
#define HB_IT_PI0    16u  /* + k: Dirichlet gamma k      stats.cpp:69-76 */
#define HB_IT_VR0   256u  /* + i: env. random variance i Bayes.cpp:511 */
/* HB_DOM_SNP slots */
#define HB_SL_CHI   0u  /* BayesA/B per-SNP chi^2(df+1)  Bayes.cpp:613,636 */
#define HB_SL_MAIN  1u  /* (U,Z): inclusion uniform + effect normal :645/649 ... */
#define HB_SL_IG    2u  /* (U,Z): BayesL inverse-Gaussian stats.cpp:57,60 */
#define HB_SL_RETRY 3u  /* SBayesS rejection loop normals SBayesS.cpp:390-397 */

typedef struct { uint32_t k0, k1; } hb_key_t;

HB_HD hb_key_t hb_make_key(uint64_t seed) {
  hb_key_t k;
  k.k0 = (uint32_t)(seed & 0xffffffffu);
  k.k1 = (uint32_t)(seed >> 32);
  return k;
}

/* ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------- */
HB_HD void hb_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#ifdef __CUDA_ARCH__
  *lo = a * b;
  *hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *lo = (uint32_t)p;
  *hi = (uint32_t)(p >> 32);
#endif
}

HB_HD void hb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    hb_mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
    hb_mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 52-bit uniform strictly inside (0,1): (x + 1/2) * 2^-52, x in [0, 2^52).
 * Exactly representable for every x, so host and device agree bit for bit. */
HB_HD double hb_u01(uint32_t a, uint32_t b) {
  uint64_t x = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
  return ((double)x + 0.5) * 2.220446049250313e-16; /* 2^-52 */
}

/* ---- standard normal by inversion: Wichura's AS241 PPND16 ----------------- */
/* (R's default normal.kind is also inversion through this same routine.)
 * Evaluated with explicit fma-free Horner steps so that host (gcc
 * -ffp-contract=off) and device produce the same central-region values; the
 * tail regions call log/sqrt and may differ from libm in the last ulp. */
#ifdef __CUDA_ARCH__
#define HB_MUL(a, b) __dmul_rn((a), (b))
#define HB_ADD(a, b) __dadd_rn((a), (b))
#else
#define HB_MUL(a, b) ((a) * (b))
#define HB_ADD(a, b) ((a) + (b))
#endif
#define HB_H(acc, r, c) HB_ADD(HB_MUL((acc), (r)), (c))

HB_HD double hb_qnorm(double p) {
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = HB_ADD(0.180625, -HB_MUL(q, q));
    double num = 2509.0809287301226727;
    num = HB_H(num, r, 33430.575583588128105);
    num = HB_H(num, r, 67265.770927008700853);
    num = HB_H(num, r, 45921.953931549871457);
    num = HB_H(num, r, 13731.693765509461125);
    num = HB_H(num, r, 1971.5909503065514427);
    num = HB_H(num, r, 133.14166789178437745);
    num = HB_H(num, r, 3.387132872796366608);
    double den = 5226.495278852545925;
    den = HB_H(den, r, 28729.085735721942674);
    den = HB_H(den, r, 39307.89580009271061);
    den = HB_H(den, r, 21213.794301586595867);
    den = HB_H(den, r, 5394.1960214247511077);
    den = HB_H(den, r, 687.1870074920579083);
    den = HB_H(den, r, 42.313330701600911252);
    den = HB_H(den, r, 1.0);
    return HB_MUL(q, num) / den;
  }
  r = (q < 0.0) ? p : (1.0 - p);
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    double num = 7.7454501427834140764e-4;
    num = HB_H(num, r, 0.0227238449892691845833);
    num = HB_H(num, r, 0.24178072517745061177);
    num = HB_H(num, r, 1.27045825245236838258);
    num = HB_H(num, r, 3.64784832476320460504);
    num = HB_H(num, r, 5.7694972214606914055);
    num = HB_H(num, r, 4.6303378461565452959);
    num = HB_H(num, r, 1.42343711074968357734);
    double den = 1.05075007164441684324e-9;
    den = HB_H(den, r, 5.475938084995344946e-4);
    den = HB_H(den, r, 0.0151986665636164571966);
    den = HB_H(den, r, 0.14810397642748007459);
    den = HB_H(den, r, 0.68976733498510000455);
    den = HB_H(den, r, 1.6763848301838038494);
    den = HB_H(den, r, 2.05319162663775882187);
    den = HB_H(den, r, 1.0);
    val = num / den;
  } else {
    r -= 5.0;
    double num = 2.01033439929228813265e-7;
    num = HB_H(num, r, 2.71155556874348757815e-5);
    num = HB_H(num, r, 0.0012426609473880784386);
    num = HB_H(num, r, 0.026532189526576123093);
    num = HB_H(num, r, 0.29656057182850489123);
    num = HB_H(num, r, 1.7848265399172913358);
    num = HB_H(num, r, 5.4637849111641143699);
    num = HB_H(num, r, 6.6579046435011037772);
    double den = 2.04426310338993978564e-15;
    den = HB_H(den, r, 1.4215117583164458887e-7);
    den = HB_H(den, r, 1.8463183175100546818e-5);
    den = HB_H(den, r, 7.868691311456132591e-4);
    den = HB_H(den, r, 0.0148753612908506148525);
    den = HB_H(den, r, 0.13692988092273580531);
    den = HB_H(den, r, 0.59983220655588793769);
    den = HB_H(den, r, 1.0);
    val = num / den;
  }
  return (q < 0.0) ? -val : val;
}

/* ---- addressed draws ------------------------------------------------------ */
/* One Philox block per address yields an independent (uniform, normal) pair. */
HB_HD void hb_draw_uz(hb_key_t key, uint32_t domain, uint32_t iter, uint32_t index,
                      uint32_t slot, uint32_t attempt, double* u, double* z) {
  uint32_t w[4];
  hb_philox4x32_10(index, (slot << 16) | (attempt & 0xffffu), iter, domain, key.k0, key.k1, w);
  *u = hb_u01(w[0], w[1]);
  *z = hb_qnorm(hb_u01(w[2], w[3]));
}

HB_HD double hb_draw_u(hb_key_t key, uint32_t domain, uint32_t iter, uint32_t index,
                       uint32_t slot, uint32_t attempt) {
  uint32_t w[4];
  hb_philox4x32_10(index, (slot << 16) | (attempt & 0xffffu), iter, domain, key.k0, key.k1, w);
  return hb_u01(w[0], w[1]);
}

HB_HD double hb_draw_z(hb_key_t key, uint32_t domain, uint32_t iter, uint32_t index,
                       uint32_t slot, uint32_t attempt) {
  uint32_t w[4];
  hb_philox4x32_10(index, (slot << 16) | (attempt & 0xffffu), iter, domain, key.k0, key.k1, w);
  return hb_qnorm(hb_u01(w[2], w[3]));
}

/* Gamma(shape, scale=1) by Marsaglia & Tsang (2000); every rejection round
 * takes the next `attempt` at the same address, so the draw is a pure function
 * of its address.  (R uses Ahrens-Dieter on its sequential stream; only the
 * distribution is shared -- see tests/test_samplers.py.) */
HB_HD double hb_draw_gamma(hb_key_t key, uint32_t domain, uint32_t iter, uint32_t index,
                           uint32_t slot, double shape) {
  double boost = 1.0;
  uint32_t attempt = 0;
  if (shape < 1.0) {
    double u0 = hb_draw_u(key, domain, iter, index, slot, 0xffffu);
    boost = pow(u0, 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double u, z;
    hb_draw_uz(key, domain, iter, index, slot, attempt, &u, &z);
    ++attempt;
    double t = 1.0 + c * z;
    if (t <= 0.0) {
      if (attempt >= 0xfff0u) return d * boost; /* unreachable in practice */
      continue;
    }
    double v = t * t * t;
    if (log(u) < 0.5 * z * z + d - d * v + d * log(v) || attempt >= 0xfff0u)
      return d * v * boost;
  }
}

/* chi^2(df) = 2 * Gamma(df/2, 1)   (R::rchisq, stats.cpp:22-24) */
HB_HD double hb_draw_chisq(hb_key_t key, uint32_t domain, uint32_t iter, uint32_t index,
                           uint32_t slot, double df) {
  return 2.0 * hb_draw_gamma(key, domain, iter, index, slot, 0.5 * df);
}

/* inverse-Gaussian(mu, lambda) from one (U,Z) pair  (stats.cpp:55-67).
 * The reference writes the smaller root of the Michael-Schucany-Haas quadratic as
 *     x = mu + 0.5*mu*mu*y/lambda - (0.5*mu/lambda)*sqrt(4*mu*lambda*y + mu*mu*y*y),   y = z*z,
 * which cancels catastrophically once w = mu*y/(2*lambda) >> 1 (BayesL reaches w ~ 1e5 through the
 * |g| >= 1e-6 clamp, Bayes.cpp:728): one ulp on mu then moves x by ~1e-6 and no two BLAS builds of
 * the reference agree with each other.  With  (1+w)^2 - (w^2+2w) = 1  the same root is
 *     x = mu * (1 + w - sqrt(w*w + 2*w)) = mu / (1 + w + sqrt(w*(w+2))),
 * algebraically identical and well conditioned; oracle and device both use this form (the
 * accept/flip step `u <= mu/(mu+x) ? x : mu*mu/x` is the reference's, stats.cpp:60-66).
 * Explicit unfused operations keep host and device rounding identical. */
HB_HD double hb_invgauss_from_uz(double mu, double lambda, double u, double z) {
  double y = HB_MUL(z, z);
  double w = HB_MUL(mu, y) / HB_MUL(2.0, lambda);
  double root = sqrt(HB_MUL(w, HB_ADD(w, 2.0)));
  double x = mu / HB_ADD(HB_ADD(1.0, w), root);
  return (u <= mu / HB_ADD(mu, x)) ? x : (HB_MUL(mu, mu) / x);
}

/* the reference's literal expression (stats.cpp:57-59), kept for the tests that show the two forms
 * agree wherever the literal one is well conditioned */
HB_HD double hb_invgauss_literal_root(double mu, double lambda, double z) {
  double y = z * z;
  return mu + 0.5 * mu * mu * y / lambda - (0.5 * mu / lambda) * sqrt(4.0 * mu * lambda * y + mu * mu * y * y);
}

#endif /* HB_RNG_H */
