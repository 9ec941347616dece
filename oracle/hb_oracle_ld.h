/*
 * hb_oracle_ld.h -- CPU oracle for the two stages in front of the Gibbs sweeps:
 * the PLINK .bed decoder and the LD / X'X builder.  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU arms may load this; hibayes_b200/ never links it.
 *
 * PARITY PINNED against the reference itself: tXXmat.cpp and read_bed.cpp compile unmodified against the stand-in
 * Rcpp / Armadillo / bigmemory / RcppProgress headers of oracle/ref_shim/ into oracle/_ref/libhibayes_ref.so, and
 * BigStat(), tXXmat_Geno(), tXXmat_Chr() and read_bed<char>() return the same bits as the functions below
 * (tests/test_reference_pin.py: the bundled demo.bed, ragged files with missing genotypes, dense / sparse / per-chromosome
 * LD with a monomorphic SNP).  Both functions are literal restatements of the reference loops -- same operation order.
 * The decoder is additionally pinned against an independent numpy decode of the reference's bundled
 * inst/extdata/demo.bed (tests/golden/demo_bed.npz, tests/test_ldmat_bed.py).
 */
#ifndef HB_ORACLE_LD_H
#define HB_ORACLE_LD_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* read_bed<char>() of /root/reference/src/read_bed.cpp:97-232.
 *   file/len   image of the .bed file INCLUDING its three leading bytes (the reference skips them
 *              unread, :146-147; this restatement also does not look at them)
 *   nid, m     individuals (rows of the big.matrix) and SNPs (columns)
 *   impt, d    the reference's arguments: impute missing by the major genotype; dominance coding
 *   na_code    NA_CHAR of bigmemory (-128) for a "char" big.matrix (:241)
 *   out        nid x m column-major int8 (the big.matrix)
 *   miss       m flags (miss[r] of :149), may be NULL
 * Returns 0, or 1 when the file is shorter than 3 + m*ceil(nid/4) bytes. */
int hbo_read_bed(const uint8_t* file, size_t len, int nid, int m, int impt, int d, int na_code, int8_t* out,
                 uint8_t* miss);

/* BigStat<char>() of /root/reference/src/tXXmat.cpp:43-77: mean, sum, xx = sqrt(sum (x-mean)^2),
 * each accumulated sequentially over the individuals in fp64. */
void hbo_bigstat(const int8_t* X, size_t ld, int n, int m, double* mean, double* sum, double* xx);

/* tXXmat_Geno<char>() (/root/reference/src/tXXmat.cpp:100-185) and tXXmat_Chr<char>() (:504-605) on an
 * n x m column-major int8 matrix.
 *   chr        NULL -> tXXmat_Geno; else m chromosome codes -> tXXmat_Chr (pairs on different
 *              chromosomes are never touched and stay 0)
 *   has_chisq  0: `chisq = R_NilValue` (dense branch: diagonal = xx^2/ind, :157 and :584);
 *              1: sparse branch with threshold chisq (every pair incl. the diagonal goes through
 *                 `r*r*ind <= chisq -> dropped`, :137-144 and :548-556).
 *              tXXmat_Geno takes the sparse branch only for chisq > 0 (:118-121), tXXmat_Chr for any
 *              non-NULL chisq (:520-523); the caller passes has_chisq accordingly.
 *   out        m x m column-major, fully written (0 where the reference stores nothing).  For the
 *              branches that return an arma::sp_mat the stored entries are exactly the non-zero
 *              entries of `out` (assigning 0 to an sp_mat element stores nothing). */
void hbo_txxmat(const int8_t* X, size_t ld, int n, int m, const int32_t* chr, int has_chisq, double chisq,
                double* out);

#ifdef __cplusplus
}
#endif
#endif
