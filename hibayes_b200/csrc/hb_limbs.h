// hb_limbs.h -- fixed-point limb arithmetic for an integer-only dot x'r (streaming side).
//
// NOT YET USED BY THE SWEEP KERNEL: a building block, verified on the host (tests/test_limbs.py), for the round-2
// inner loop that replaces PRMT + DFMA per genotype by IDP4A.  A residual value r is held as the 48-bit fixed-point
// integer q = rint(r * scale), |q| < 2^47, split into six 8-bit limbs
//     q = u0 + u1 2^8 + u2 2^16 + u3 2^24 + u4 2^32 + s5 2^40,   u_k in [0, 255] (unsigned), s5 in [-128, 127] (signed),
// stored limb-major so that the limbs k of four consecutive rows form one 32-bit word.  Four genotypes of a column are
// one 32-bit word as they lie in memory (bytes in {0,1,2}); the dot over those four rows is six dp4a (five u8 x u8, one
// s8 x s8), each exact in int32, and  x'q = sum_k acc_k 2^(8k)  is exact in int64.  The only rounding is the
// quantisation of r: |x'r - x'q / scale| <= sum_i x_i / (2 scale).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define HB_LIMB_FN __host__ __device__ __forceinline__
#else
#define HB_LIMB_FN static inline
#endif

#define HB_NLIMB 6

// q = rint(r * scale) must satisfy |q| < 2^47; returns 0 when it does not (the caller's scale is too large).
HB_LIMB_FN int hb_limb_split(double r, double scale, uint8_t limb[HB_NLIMB]) {
  const double v = r * scale;
  if (!(v > -140737488355328.0 && v < 140737488355328.0)) return 0;   // 2^47
#ifdef __CUDA_ARCH__
  long long q = __double2ll_rn(v);
#else
  long long q = (long long)__builtin_rint(v);
#endif
  for (int k = 0; k < HB_NLIMB - 1; ++k) {
    limb[k] = (uint8_t)(q & 0xff);
    q >>= 8;   // arithmetic shift: floor division, so the remaining top part carries the sign
  }
  limb[HB_NLIMB - 1] = (uint8_t)(int8_t)q;   // in [-128, 127] because |q_original| < 2^47
  return 1;
}

// one dp4a step: four genotype bytes (one word) against the limb-k bytes of the same four rows
HB_LIMB_FN int hb_limb_dp4a(uint32_t xword, uint32_t limbword, int k, int acc) {
#ifdef __CUDA_ARCH__
  return k == HB_NLIMB - 1 ? __dp4a((int)xword, (int)limbword, acc) : (int)__dp4a(xword, limbword, (unsigned)acc);
#else
  for (int b = 0; b < 4; ++b) {
    const int x = (int)((xword >> (8 * b)) & 0xff);   // genotypes are 0, 1, 2: the same value signed or unsigned
    const int l = (int)((limbword >> (8 * b)) & 0xff);
    acc += x * (k == HB_NLIMB - 1 ? (int)(int8_t)l : l);
  }
  return acc;
#endif
}

// x'q from the six int32 limb sums.  Valid while 0 <= acc[0..4] < 2^15 and |acc[5]| < 2^15, i.e. for sums over at
// most 32 rows of genotypes in {0,1,2} (a lane of the sweep kernel merges after its 24 rows): the two halves then fit
// 32-bit integers (lo < 2^31, |hi| < 2^31) and the merge costs a handful of integer multiply-adds.
HB_LIMB_FN long long hb_limb_merge(const int acc[HB_NLIMB]) {
  const int lo = acc[0] + acc[1] * 256 + acc[2] * 65536;
  const int hi = acc[3] + acc[4] * 256 + acc[5] * 65536;
  return (long long)hi * 16777216ll + (long long)lo;
}
