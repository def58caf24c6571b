// Minimal stand-in for the few Rcpp containers the reference's only native source file
// (src/clustering.cpp) uses, so that file can be compiled UNMODIFIED, from where it lies in the
// reference tree, into oracle/_ref/libm3ref.so (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
// nothing under monocle3_b200/ includes or links this.
//
// What is emulated (semantics per the Rcpp documentation):
//   NumericMatrix(nrow, ncol) zero-filled, column-major; NumericMatrix(SEXP); m(i, j);
//   m(i, _) -> row as NumericVector;  m(_, j) -> column proxy (assignable, "/ double", max());
//   intersect(a, b) -> the UNIQUE values present in both (Rcpp sugar set semantics);
//   as<bool>(SEXP);  R::pnorm(x, mu, sigma, lower_tail, log_p).
#pragma once
#include <algorithm>
#include <cmath>
#include <set>
#include <vector>

typedef void* SEXP;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

namespace Rcpp {

struct Underscore {};
static const Underscore _ = Underscore();

class NumericVector {
 public:
  NumericVector() {}
  explicit NumericVector(size_t n) : v_(n, 0.0) {}
  size_t size() const { return v_.size(); }
  double& operator[](size_t k) { return v_[k]; }
  double operator[](size_t k) const { return v_[k]; }
  std::vector<double>::const_iterator begin() const { return v_.begin(); }
  std::vector<double>::const_iterator end() const { return v_.end(); }

 private:
  std::vector<double> v_;
};

class NumericMatrix;

class ColumnProxy {
 public:
  ColumnProxy(NumericMatrix& m, int col) : m_(m), col_(col) {}
  ColumnProxy& operator=(const NumericVector& v);
  ColumnProxy& operator=(const ColumnProxy& o);
  NumericVector get() const;

 private:
  NumericMatrix& m_;
  int col_;
};

class NumericMatrix {
 public:
  NumericMatrix(int nrow, int ncol) : nr_(nrow), nc_(ncol), d_((size_t)nrow * ncol, 0.0) {}
  NumericMatrix(SEXP s) { *this = *static_cast<const NumericMatrix*>(s); }
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double& operator()(int i, int j) { return d_[(size_t)j * nr_ + i]; }
  double operator()(int i, int j) const { return d_[(size_t)j * nr_ + i]; }
  NumericVector operator()(int i, Underscore) const {
    NumericVector r((size_t)nc_);
    for (int j = 0; j < nc_; ++j) r[j] = (*this)(i, j);
    return r;
  }
  ColumnProxy operator()(Underscore, int j) { return ColumnProxy(*this, j); }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }

 private:
  int nr_ = 0, nc_ = 0;
  std::vector<double> d_;
};

inline NumericVector ColumnProxy::get() const {
  NumericVector r((size_t)m_.nrow());
  for (int i = 0; i < m_.nrow(); ++i) r[i] = m_(i, col_);
  return r;
}
inline ColumnProxy& ColumnProxy::operator=(const NumericVector& v) {
  for (int i = 0; i < m_.nrow(); ++i) m_(i, col_) = v[i];
  return *this;
}
inline ColumnProxy& ColumnProxy::operator=(const ColumnProxy& o) { return *this = o.get(); }

inline NumericVector operator/(const ColumnProxy& c, double d) {
  NumericVector r = c.get();
  for (size_t k = 0; k < r.size(); ++k) r[k] = r[k] / d;
  return r;
}
inline double max(const ColumnProxy& c) {
  NumericVector r = c.get();
  double m = r.size() ? r[0] : -INFINITY;
  for (size_t k = 1; k < r.size(); ++k) m = r[k] > m ? r[k] : m;
  return m;
}

// set intersection: every value reported once
inline NumericVector intersect(const NumericVector& a, const NumericVector& b) {
  std::set<double> sb(b.begin(), b.end()), seen;
  std::vector<double> out;
  for (double x : a)
    if (sb.count(x) && seen.insert(x).second) out.push_back(x);
  NumericVector r(out.size());
  for (size_t k = 0; k < out.size(); ++k) r[k] = out[k];
  return r;
}

template <typename T>
T as(SEXP s) {
  return *static_cast<const T*>(s);
}

}  // namespace Rcpp

namespace R {
inline double pnorm(double x, double mu, double sigma, int lower_tail, int log_p) {
  const double z = (x - mu) / sigma;
  double p = lower_tail ? 0.5 * std::erfc(-z / std::sqrt(2.0)) : 0.5 * std::erfc(z / std::sqrt(2.0));
  return log_p ? std::log(p) : p;
}
}  // namespace R
