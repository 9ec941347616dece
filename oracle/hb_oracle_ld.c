/*
 * hb_oracle_ld.c -- CPU oracle of the .bed decoder and the LD builder.  TEST INFRASTRUCTURE ONLY
 * (see hb_oracle_ld.h).  Literal restatements of /root/reference/src/read_bed.cpp:97-232 and
 * /root/reference/src/tXXmat.cpp:43-77,100-185,504-605; compiled with -ffp-contract=off so that the
 * fp64 expressions round where the reference's do.
 */
#include "hb_oracle_ld.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* read_bed.cpp:97-232 */
int hbo_read_bed(const uint8_t* file, size_t len, int nid, int m, int impt, int d, int na_code, int8_t* out,
                 uint8_t* miss_out) {
  long n = nid / 4; /* :109 "4 individual = 1 bit" (meaning: one byte) */
  if (nid % 4 != 0) n++; /* :112-113 */
  if (len < 3 + (size_t)n * (size_t)m) return 1;
  /* code map :118-122 */
  int code[4];
  code[3] = 0;
  code[2] = 1;
  code[1] = na_code;
  code[0] = d ? 0 : 2;
  size_t ggvec[3] = {0, 1, d ? 0u : 2u}; /* :124-129 */
  const uint8_t* buffer = file + 3;       /* :146-147: three bytes read and dropped */
  uint8_t* miss = (uint8_t*)calloc((size_t)m > 0 ? (size_t)m : 1, 1);
  const size_t total = (size_t)n * (size_t)m;
  for (size_t j = 0; j < total; j++) { /* :157-170 with one block */
    size_t r = j / (size_t)n;
    size_t c = j % (size_t)n * 4;
    uint8_t p = buffer[j];
    for (size_t x = 0; x < 4 && (c + x) < (size_t)nid; x++) {
      int8_t gg = (int8_t)code[(p >> (2 * x)) & 0x03];
      out[r * (size_t)nid + c + x] = gg;
      if (gg == (int8_t)na_code) miss[r] = 1;
    }
  }
  int NMISS = 0; /* :176-184 */
  for (int i = 0; i < m; i++) {
    if (miss[i]) break;
    NMISS++;
  }
  if (impt && NMISS != m) { /* :186-229 */
    size_t* na_index = (size_t*)malloc(sizeof(size_t) * (size_t)(nid > 0 ? nid : 1));
    for (int i = 0; i < m; i++) {
      if (!miss[i]) continue;
      size_t n_na = 0, counts[3] = {0, 0, 0};
      int max = 0;
      int8_t major = 0;
      int8_t* col = out + (size_t)i * (size_t)nid;
      for (int j = 0; j < nid; j++) {
        int v = (int)col[j];
        if (v == 0) counts[0]++;
        else if (v == 1) counts[1]++;
        else if (!d && v == 2) counts[2]++; /* the dominance switch has no case 2 (:203-209) */
        else na_index[n_na++] = (size_t)j;
      }
      for (size_t j = 0; j < 3; j++) { /* :221-226: first strict maximum */
        if ((long)counts[j] > (long)max) {
          max = (int)counts[j];
          major = (int8_t)ggvec[j];
        }
      }
      for (size_t q = 0; q < n_na; q++) col[na_index[q]] = major; /* :229-231 */
    }
    free(na_index);
  }
  if (miss_out) memcpy(miss_out, miss, (size_t)m);
  free(miss);
  return 0;
}

/* tXXmat.cpp:43-77 */
void hbo_bigstat(const int8_t* X, size_t ld, int n, int m, double* mean, double* sum, double* xx) {
#pragma omp parallel for
  for (int j = 0; j < m; j++) {
    const int8_t* col = X + (size_t)j * ld;
    double p1 = 0.0;
    for (int k = 0; k < n; k++) p1 += col[k]; /* :58-60 */
    sum[j] = p1;
    mean[j] = p1 / n; /* :62 */
  }
#pragma omp parallel for
  for (int j = 0; j < m; j++) {
    const int8_t* col = X + (size_t)j * ld;
    double p1 = 0.0;
    for (int k = 0; k < n; k++) { /* :68-71 */
      double scale_mean = (col[k] - mean[j]);
      p1 += scale_mean * scale_mean;
    }
    xx[j] = sqrt(p1); /* :72 */
  }
}

/* tXXmat.cpp:100-185 (chr == NULL) and :504-605 (chr != NULL) */
void hbo_txxmat(const int8_t* X, size_t ld, int n, int m, const int32_t* chr, int has_chisq, double chisq,
                double* out) {
  const int ind = n;
  double* mean_all = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  double* sum_all = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  double* xx_all = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  hbo_bigstat(X, ld, n, m, mean_all, sum_all, xx_all);
  memset(out, 0, sizeof(double) * (size_t)m * (size_t)m);
#pragma omp parallel for schedule(dynamic)
  for (int j = 0; j < m; j++) {
    const int8_t* xj = X + (size_t)j * ld;
    const double p1 = xx_all[j], m1 = mean_all[j], sum1 = sum_all[j];
    if (!has_chisq) out[(size_t)j * m + j] = p1 * p1 / ind; /* :157 / :584 */
    for (int i = has_chisq ? j : j + 1; i < m; i++) {          /* :133 / :158 */
      if (chr && chr[i] != chr[j]) continue;                    /* per-chromosome index sets, :525-527 */
      const int8_t* xi = X + (size_t)i * ld;
      double p12 = 0;
      const double p2 = xx_all[i], m2 = mean_all[i], sum2 = sum_all[i];
      for (int k = 0; k < ind; k++) p12 += (xi[k]) * (xj[k]); /* :138-140: int product, exact sum */
      p12 -= sum1 * m2 + sum2 * m1 - ind * m1 * m2;           /* :141 */
      if (has_chisq) {
        double r = p12 / (p1 * p2); /* :142 */
        if (r * r * ind <= chisq) continue; /* :143-145 */
      }
      out[(size_t)j * m + i] = out[(size_t)i * m + j] = p12 / ind; /* :148 / :168 */
    }
  }
  free(mean_all);
  free(sum_all);
  free(xx_all);
}
