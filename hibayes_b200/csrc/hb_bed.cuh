// hb_bed.cuh -- PLINK .bed (SNP-major, 2 bits per genotype) decoding on the device.
//
// Replaces read_bed<char>() of /root/reference/src/read_bed.cpp:97-232: the code map (:118-122),
// the per-SNP missing flag (:149,166-168) and the imputation of missing genotypes by the major
// genotype (:186-229).  The byte-level functions are __host__ __device__ so that tests/ can run the
// very same code on the CPU (hb_test_bed_* in ldmat.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// 2-bit field -> genotype (read_bed.cpp:118-122): 0 -> 2 (0 with dominance coding), 1 -> missing,
// 2 -> 1, 3 -> 0.  `na` is what a missing genotype becomes.
__host__ __device__ __forceinline__ int bed_code(unsigned field, int d, int na) {
  return field == 3u ? 0 : field == 2u ? 1 : field == 1u ? na : (d ? 0 : 2);
}

// Genotype of individual `id` in a SNP's byte row.
__host__ __device__ __forceinline__ unsigned bed_field(const uint8_t* snp_bytes, size_t id) {
  return ((unsigned)snp_bytes[id >> 2] >> (2u * (unsigned)(id & 3u))) & 3u;
}

// Major genotype from the four field counts c[f] of one SNP (read_bed.cpp:199-226): counts over the
// genotype values 0, 1, 2 (dominance: 0 and 1 only, both homozygotes count as 0), the first strict
// maximum wins, 0 when every genotype is missing.
__host__ __device__ __forceinline__ int bed_major(const unsigned long long c[4], int d) {
  const unsigned long long counts[3] = {d ? c[3] + c[0] : c[3], c[2], d ? 0ull : c[0]};
  const int ggvec[3] = {0, 1, d ? 0 : 2};
  unsigned long long best = 0;
  int major = 0;
  for (int j = 0; j < 3; ++j)
    if (counts[j] > best) {
      best = counts[j];
      major = ggvec[j];
    }
  return major;
}

// Adds the field counts of one byte of which the first `valid` (1..4) fields belong to individuals.
__host__ __device__ __forceinline__ void bed_count_byte(unsigned byte, int valid, unsigned c[4]) {
  for (int x = 0; x < valid; ++x) c[(byte >> (2 * x)) & 3u]++;
}

#ifdef __CUDACC__
// One warp per SNP: field counts over all `nid` individuals of the file -> major genotype and the
// missing flag.  info[j] = major | (has_missing << 7).
static __global__ void k_bed_info(const uint8_t* __restrict__ bed, size_t bps, int nid, int ncols, int d,
                                  uint8_t* __restrict__ info) {
  const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= ncols) return;
  const uint8_t* row = bed + (size_t)warp * bps;
  unsigned c[4] = {0u, 0u, 0u, 0u};
  const size_t full = (size_t)nid >> 2;  // bytes whose four fields are all individuals
  for (size_t b = lane; b < full; b += 32) bed_count_byte(row[b], 4, c);
  if (lane == 0 && (nid & 3)) bed_count_byte(row[full], nid & 3, c);
  unsigned long long t[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    unsigned v = c[f];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    t[f] = v;
  }
  if (lane == 0) info[warp] = (uint8_t)(bed_major(t, d) | (t[1] ? 0x80 : 0));
}

// Genotype of output row `row` of SNP column c as the consumers want it: file individual
// rows[row] (or `row` itself), missing -> major genotype when imputing, else `na`.
__device__ __forceinline__ int bed_value(const uint8_t* snp_bytes, const int32_t* __restrict__ rows, size_t row, int d,
                                         int impt, int na, uint8_t info) {
  const size_t id = rows ? (size_t)rows[row] : row;
  const unsigned f = bed_field(snp_bytes, id);
  if (f == 1u) return impt ? (int)(info & 0x7f) : na;
  return bed_code(f, d, na);
}
#endif

}  // namespace hb
