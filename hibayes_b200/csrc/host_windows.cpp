// host_windows.cpp -- the window cutters that produce `windindx` for the WPPA counters of the sweeps
// (/root/reference/src/cutwind.cpp:13-65; called from R/bayes.r, R/sbayes.r:176-182).  Host-only index bookkeeping;
// it lives behind the same C ABI because the sweep drivers consume its output (hb_bayes_args.windindx).
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/hibayes_b200.h"

int hb_set_error(const char* fmt, ...);  // engine.cu

// sorted unique values, as arma::unique() returns them
static std::vector<double> unique_sorted(const double* v, int m) {
  std::vector<double> u(v, v + m);
  std::sort(u.begin(), u.end());
  u.erase(std::unique(u.begin(), u.end()), u.end());
  return u;
}

// cutwind.cpp:13-36: per chromosome (ascending code), windows [bp0, bp0 + bp) starting at bp0 = 1; empty windows get no
// number.  SNPs at positions < 1 belong to no window (the reference leaves their entry uninitialised): 0 here.
extern "C" int hb_cutwind_by_bp(const double* chr, const double* pos, int m, double bp, int32_t* windindx) {
  if (!chr || !pos || !windindx || m <= 0) return hb_set_error("hb_cutwind_by_bp: bad argument");
  if (!(bp > 0)) return hb_set_error("hb_cutwind_by_bp: window size must be positive");
  std::fill(windindx, windindx + m, 0);
  int count = 1;
  for (double c : unique_sorted(chr, m)) {
    std::vector<int> idx;
    double maxbp = 0;
    bool first = true;
    for (int i = 0; i < m; ++i)
      if (chr[i] == c) {
        idx.push_back(i);
        if (first || pos[i] > maxbp) maxbp = pos[i];
        first = false;
      }
    for (double bp0 = 1; bp0 <= maxbp; bp0 = bp0 + bp) {   // :24-31, the same floating-point recurrence
      bool any = false;
      for (int i : idx)
        if (pos[i] >= bp0 && pos[i] < (bp0 + bp)) { windindx[i] = count; any = true; }
      if (any) count++;
    }
  }
  return 0;
}

// cutwind.cpp:39-65: per chromosome, consecutive groups of fixN SNPs in order of position (a chromosome with at most
// fixN SNPs is one window).  Ties in position keep file order (the reference's sort_index does not define it).
extern "C" int hb_cutwind_by_num(const double* chr, const double* pos, int m, int fixN, int32_t* windindx) {
  if (!chr || !pos || !windindx || m <= 0) return hb_set_error("hb_cutwind_by_num: bad argument");
  if (fixN <= 0) return hb_set_error("hb_cutwind_by_num: window length must be positive");
  std::fill(windindx, windindx + m, 0);
  int count = 1;
  for (double c : unique_sorted(chr, m)) {
    std::vector<int> idx;
    for (int i = 0; i < m; ++i)
      if (chr[i] == c) idx.push_back(i);
    const int chrlen = (int)idx.size();
    if (chrlen <= fixN) {
      for (int i : idx) windindx[i] = count;
      count++;
      continue;
    }
    std::vector<int> order(chrlen);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pos[idx[a]] < pos[idx[b]]; });
    int st = 0, end = 0;
    while (end < chrlen - 1) {   // :53-60
      end = st + fixN - 1;
      if (end > chrlen - 1) end = chrlen - 1;
      for (int k = st; k <= end; ++k) windindx[idx[order[k]]] = count;
      st += fixN;
      count++;
    }
  }
  return 0;
}
