// Stand-in for <Rcpp.h>, TEST INFRASTRUCTURE ONLY: the image has no R, so r_adapter/*.cpp cannot be
// built against the real headers.  This file declares -- with the signatures of Rcpp 1.0.x -- exactly
// the part of the API the adapter uses, so that `g++ -fsyntax-only` parses and type-checks it
// (tests/test_abi.py::test_r_adapter_parses).  Nothing here is ever linked or run.
#ifndef MIXDIR_B200_RCPP_STUB_H
#define MIXDIR_B200_RCPP_STUB_H

#include <cstddef>
#include <cstdint>
#include <string>

struct SEXPREC;
typedef SEXPREC* SEXP;
extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define REALSXP 14
#define INTSXP 13
#define LGLSXP 10
#define VECSXP 19

namespace Rcpp {

template <class T> class PreserveStorage {};

namespace internal {
struct NamedPlaceHolder;
}

template <int RTYPE> struct storage_of { typedef double type; };
template <> struct storage_of<INTSXP> { typedef int type; };
template <> struct storage_of<LGLSXP> { typedef int type; };

template <int RTYPE>
class Vector {
 public:
  typedef typename storage_of<RTYPE>::type stored_type;
  typedef stored_type* iterator;
  typedef const stored_type* const_iterator;
  Vector();
  Vector(SEXP x);
  explicit Vector(int n);
  explicit Vector(std::size_t n);
  Vector(int n, const stored_type& fill);
  iterator begin();
  iterator end();
  const_iterator begin() const;
  const_iterator end() const;
  long size() const;
  stored_type& operator[](long i);
  const stored_type& operator[](long i) const;
  operator SEXP() const;
};

typedef Vector<REALSXP> NumericVector;
typedef Vector<INTSXP> IntegerVector;
typedef Vector<LGLSXP> LogicalVector;

template <int RTYPE>
class MatrixColumn {
 public:
  operator Vector<RTYPE>() const;
};

template <int RTYPE>
class Matrix : public Vector<RTYPE> {
 public:
  typedef typename Vector<RTYPE>::stored_type stored_type;
  Matrix();
  Matrix(SEXP x);
  Matrix(int nrow, int ncol);
  int nrow() const;
  int ncol() const;
  stored_type& operator()(int i, int j);
  MatrixColumn<RTYPE> operator()(const internal::NamedPlaceHolder&, int j);
};

typedef Matrix<REALSXP> NumericMatrix;

template <int RTYPE>
Vector<RTYPE> head(const Vector<RTYPE>& v, int n);

namespace traits {
template <typename T>
struct named_object {
  named_object(const std::string& name, const T& o);
};
}  // namespace traits

class Argument {
 public:
  explicit Argument(const std::string& name);
  template <typename T>
  traits::named_object<T> operator=(const T& t) const;
};

namespace internal {
struct NamedPlaceHolder {
  Argument operator[](const std::string& name) const;
};
}  // namespace internal
extern internal::NamedPlaceHolder _;

class List {
 public:
  template <typename... Args>
  static List create(const traits::named_object<Args>&... args);
  operator SEXP() const;
};

template <typename T>
void standard_delete_finalizer(T* obj);

template <typename T, template <class> class StoragePolicy = PreserveStorage,
          void Finalizer(T*) = standard_delete_finalizer<T>, bool finalizeOnExit = false>
class XPtr {
 public:
  explicit XPtr(SEXP x);
  explicit XPtr(T* p, bool set_delete_finalizer = true, SEXP tag = R_NilValue, SEXP prot = R_NilValue);
  T* get() const;
  T* checked_get() const;
  void release();
  operator SEXP() const;
};

template <typename... Args>
[[noreturn]] void stop(const char* fmt, Args&&... args);
template <typename... Args>
void warning(const char* fmt, Args&&... args);

}  // namespace Rcpp

#endif
