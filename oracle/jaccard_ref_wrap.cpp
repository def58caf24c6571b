// C entry point around the REFERENCE's own jaccard_coeff (src/clustering.cpp:13-58), which is
// #included from the reference tree (never copied) and compiled against oracle/rcpp_shim/Rcpp.h.
// TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref_native.py uses the resulting
// oracle/_ref/libm3ref.so to pin oracle/monocle3_ref.py::jaccard_coeff (and through it the CUDA
// kernel m3b_jaccard_coeff) to the reference's native code.
#include "clustering.cpp"  // found via -I <reference>/src

extern "C" int m3ref_jaccard_coeff(const double* idx /* n x k column-major, 1-based ids */, int n, int k, int weight,
                                   double* out /* (n*k) x 3 column-major */) {
  Rcpp::NumericMatrix m(n, k);
  for (long long q = 0; q < (long long)n * k; ++q) m.data()[q] = idx[q];
  bool w = weight != 0;
  Rcpp::NumericMatrix r = jaccard_coeff(static_cast<SEXP>(&m), static_cast<SEXP>(&w));
  for (long long q = 0; q < (long long)n * k * 3; ++q) out[q] = r.data()[q];
  return 0;
}
