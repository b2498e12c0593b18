// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Stand-in for <Rcpp.h>, just wide enough that the reference translation unit
// /root/reference/src/mixdir_impl.cpp compiles UNMODIFIED with plain g++ (no R,
// no Rcpp in this image).  It provides the surface that file touches
// (SURVEY.md Appendix A.1):
//   Rcpp::NumericVector  - length-constructed (zero-filled) or a view over a
//                          caller buffer; operator[] taking an arithmetic index
//                          (the reference indexes with a double expression,
//                          mixdir_impl.cpp:32,80,145); static is_na(); iterable
//                          for Rcpp::sum (mixdir_impl.cpp:59).
//   Rcpp::NumericMatrix  - column-major view, operator()(i,j), nrow(), ncol(),
//                          static is_na(); copies alias the caller's buffer,
//                          like an Rcpp matrix wrapping a REALSXP
//                          (mixdir_impl.cpp:83,86).
//   R::digamma           - forwarded to the oracle's own digamma
//                          (oracle_digamma, oracle/mixdir_oracle.c).
// R's NA_real_ is a NaN payload; here NA == any NaN (is_na = isnan), which is
// how the test harness encodes missing cells.
#ifndef MIXDIR_ORACLE_RCPP_SHIM_H
#define MIXDIR_ORACLE_RCPP_SHIM_H

#include <cmath>
#include <cstddef>
#include <memory>
#include <vector>

extern "C" double oracle_digamma(double x);

namespace R {
inline double digamma(double x) { return oracle_digamma(x); }
}  // namespace R

namespace Rcpp {

class NumericVector {
 public:
  // zero-filled, owning (mixdir_impl.cpp:61,118,131)
  explicit NumericVector(int n)
      : own_(std::make_shared<std::vector<double>>(static_cast<size_t>(n), 0.0)),
        p_(own_->data()), n_(n) {}
  // non-owning view over caller memory
  NumericVector(double* p, std::ptrdiff_t n) : p_(p), n_(n) {}

  template <typename I>
  double& operator[](I i) { return p_[static_cast<std::ptrdiff_t>(i)]; }
  template <typename I>
  const double& operator[](I i) const { return p_[static_cast<std::ptrdiff_t>(i)]; }

  std::ptrdiff_t size() const { return n_; }
  const double* begin() const { return p_; }
  const double* end() const { return p_ + n_; }
  static bool is_na(double v) { return std::isnan(v); }

 private:
  std::shared_ptr<std::vector<double>> own_;
  double* p_;
  std::ptrdiff_t n_;
};

class NumericMatrix {
 public:
  NumericMatrix(double* p, int nrow, int ncol) : p_(p), nr_(nrow), nc_(ncol) {}
  double& operator()(int i, int j) { return p_[i + static_cast<std::ptrdiff_t>(j) * nr_]; }
  const double& operator()(int i, int j) const { return p_[i + static_cast<std::ptrdiff_t>(j) * nr_]; }
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double* data() { return p_; }
  static bool is_na(double v) { return std::isnan(v); }

 private:
  double* p_;
  int nr_, nc_;
};

// Rcpp::sum over a NumericVector; Rcpp sugar accumulates in double.
inline double sum(const NumericVector& v) {
  double s = 0;
  for (const double* it = v.begin(); it != v.end(); ++it) s += *it;
  return s;
}

}  // namespace Rcpp

#endif
