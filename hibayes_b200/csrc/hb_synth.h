/*
 * hb_synth.h -- synthetic genotype generator shared by host and device:
 * allele frequency p_j ~ U(0.05, 0.5), x_ij ~ Binomial(2, p_j), addressed by
 * (seed, global row, column) so that the matrix does not depend on how rows are sharded.
 * Four genotypes (rows 4q..4q+3 of column j) come from one Philox block.
 */
#ifndef HB_SYNTH_H
#define HB_SYNTH_H
#include "hb_rng.h"

HB_HD void hb_synth_thresholds(hb_key_t key, uint32_t col, double* t0, double* t1) {
  uint32_t w[4];
  hb_philox4x32_10(col, 0u, 0u, 0x78u, key.k0, key.k1, w);
  double p = HB_ADD(0.05, HB_MUL(0.45, hb_u01(w[0], w[1])));
  double q = HB_ADD(1.0, -p);
  *t0 = HB_MUL(q, q);
  *t1 = HB_ADD(*t0, HB_MUL(HB_MUL(2.0, p), q));
}

/* genotypes of rows 4*rowblock .. 4*rowblock+3 packed little-endian into one word */
HB_HD uint32_t hb_synth_word(hb_key_t key, uint32_t col, uint64_t rowblock, double t0, double t1) {
  uint32_t w[4];
  hb_philox4x32_10(col, (uint32_t)rowblock, (uint32_t)(rowblock >> 32), 0x77u, key.k0, key.k1, w);
  uint32_t out = 0;
  for (int k = 0; k < 4; ++k) {
    double u = HB_MUL(HB_ADD((double)w[k], 0.5), 2.3283064365386963e-10); /* 2^-32 */
    uint32_t x = (u < t0) ? 0u : ((u < t1) ? 1u : 2u);
    out |= x << (8 * k);
  }
  return out;
}
#endif
