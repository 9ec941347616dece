// Drop-in body for the char big.matrix path of hibayes' src/read_bed.cpp (:97-247): the .bed image is decoded, and its
// missing genotypes imputed by the major genotype, on the GPU through hb_bed_decode() (include/hibayes_b200.h).  The exported
// signature stays; rMap_c() and the other big.matrix types keep the reference's CPU code.
#include "hb_dropin.h"
#include <bigmemory/BigMatrix.h>
#include <bigmemory/MatrixAccessor.hpp>
#include <cstdio>

using namespace Rcpp;

// [[Rcpp::export]]
void read_bed(std::string bfile, const SEXP pBigMat, const long maxLine, const bool impt = true, const bool d=false, const int threads=0){
    XPtr<BigMatrix> xpMat(pBigMat);
    if(xpMat->matrix_type() != 1)  throw Rcpp::exception("the GPU path takes a big.matrix of type char (what read_plink() builds)");
    std::string ending = ".bed";                                                     // :99-102
    if (bfile.length() <= ending.length() || 0 != bfile.compare(bfile.length() - ending.length(), ending.length(), ending))
        bfile += ending;
    FILE *fin = fopen(bfile.c_str(), "rb");
    if(!fin)  throw Rcpp::exception(("Error: can not open the file [" + bfile + "].").c_str());
    fseek(fin, 0, SEEK_END);
    const long length = ftell(fin);
    rewind(fin);
    std::vector<uint8_t> img(length > 0 ? length : 0);
    const size_t got = fread(img.data(), 1, img.size(), fin);
    fclose(fin);
    if(got != img.size())  throw Rcpp::exception("Error: short read on the .bed file.");
    const int nid = xpMat->nrow(), m = xpMat->ncol();
    std::vector<uint8_t> miss(m);
    if(hb_bed_decode(0, img.data(), img.size(), nid, m, impt ? 1 : 0, d ? 1 : 0, reinterpret_cast<int8_t*>(xpMat->matrix()), miss.data()) != 0)
        throw Rcpp::exception(hb_last_error());                                      // column-major char matrix, NA_CHAR = -128
    bool any = false;
    for(int j = 0; j < m; j++)  if(miss[j]){ any = true; break; }
    if(any && impt)  Rcout << "Imputing missing values by major genotype..." << std::endl;   // :188
}
