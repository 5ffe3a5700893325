// Stand-in for <RcppArmadillo.h> -- TEST INFRASTRUCTURE ONLY (part of the oracle).
//
// The reference's native routines (src/utils.cpp: row_sds, row_sds_perm, ProbSampleReplace,
// bg_sample_helper, euc_dist) include RcppArmadillo, which this image does not have (no R, no
// Rcpp, no Armadillo).  This header implements ONLY the handful of Armadillo / Rcpp calls that file
// makes, so that the UNMODIFIED reference source compiles where it lies
// (/root/reference/src/utils.cpp) into oracle/_ref/ -- see oracle/Makefile.  Nothing here is taken
// from Armadillo or Rcpp sources; the numerical routines follow Armadillo's published algorithms
// (version unpinned, DESCRIPTION gives none):
//   * accumulate / sum: two interleaved accumulators (even, odd positions), added at the end;
//   * mean: accumulate / n, running-mean form when that is not finite;
//   * var (norm_type 0): two-pass with the mean, corrected by the residual sum
//     (acc2 - acc3^2/n)/(n-1), Welford form when that is not finite; stddev = sqrt(var);
//   * linspace: start + i*delta computed in double, converted to the element type.
// R's unif_rand() is supplied by oracle/refbuild/ref_capi.cpp (R's Mersenne-Twister + set.seed
// scrambling, pinned there to known runif() values).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

extern "C" double unif_rand(void);   // R API (R_ext/Random.h); defined in ref_capi.cpp

namespace arma {

typedef unsigned long long uword;

namespace fill { struct fill_zeros {}; static const fill_zeros zeros = fill_zeros(); }
struct datum { static constexpr double nan = std::numeric_limits<double>::quiet_NaN(); };
struct span { uword a, b; span(uword a_, uword b_) : a(a_), b(b_) {} };

template <typename T> struct Col;
template <typename T> struct Row;
template <typename T> struct subview_row;

// column-major dense matrix
template <typename T> struct Mat {
  typedef T* iterator;
  typedef T elem_type;
  uword n_rows, n_cols, n_elem;
  std::vector<T> mem;
  Mat() : n_rows(0), n_cols(0), n_elem(0) {}
  Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem((size_t)(r * c), T(0)) {}
  Mat(uword r, uword c, const fill::fill_zeros&) : Mat(r, c) {}
  T* begin() { return mem.data(); }
  T* end() { return mem.data() + n_elem; }
  const T* begin() const { return mem.data(); }
  const T* end() const { return mem.data() + n_elem; }
  T* memptr() { return mem.data(); }
  const T* memptr() const { return mem.data(); }
  uword size() const { return n_elem; }
  T& operator[](uword i) { return mem[(size_t)i]; }
  const T& operator[](uword i) const { return mem[(size_t)i]; }
  T& operator()(uword i) { return mem[(size_t)i]; }
  const T& operator()(uword i) const { return mem[(size_t)i]; }
  T& operator()(uword r, uword c) { return mem[(size_t)(r + c * n_rows)]; }
  const T& operator()(uword r, uword c) const { return mem[(size_t)(r + c * n_rows)]; }
  subview_row<T> row(uword r) { return subview_row<T>(*this, r); }
  Col<T> col(uword c) const;
  Mat<T> cols(const Col<uword>& ix) const;
};

template <typename T> struct Col : public Mat<T> {
  Col() : Mat<T>() {}
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(const Mat<T>& m) : Mat<T>(m) {}             // n x 1 result of a matrix expression
  using Mat<T>::operator();
  Col<T> operator()(const Col<uword>& ix) const {  // element gather
    Col<T> out(ix.n_elem);
    for (uword i = 0; i < ix.n_elem; ++i) out[i] = (*this)[ix[i]];
    return out;
  }
  Col<T>& operator/=(T s) { for (uword i = 0; i < this->n_elem; ++i) (*this)[i] /= s; return *this; }
};

template <typename T> struct Row : public Mat<T> {
  Row() : Mat<T>() {}
  explicit Row(uword n) : Mat<T>(1, n) {}
  using Mat<T>::operator();
  Row<T> operator()(const Col<uword>& ix) const {
    Row<T> out(ix.n_elem);
    for (uword i = 0; i < ix.n_elem; ++i) out[i] = (*this)[ix[i]];
    return out;
  }
  Row<T> operator()(const span& s) const {
    Row<T> out(s.b - s.a + 1);
    for (uword i = s.a; i <= s.b; ++i) out[i - s.a] = (*this)[i];
    return out;
  }
};

template <typename T> struct subview_row {
  Mat<T>& m; uword r;
  subview_row(Mat<T>& m_, uword r_) : m(m_), r(r_) {}
  operator Row<T>() const {
    Row<T> out(m.n_cols);
    for (uword c = 0; c < m.n_cols; ++c) out[c] = m(r, c);
    return out;
  }
  subview_row& operator=(const Row<T>& v) {
    for (uword c = 0; c < m.n_cols; ++c) m(r, c) = v[c];
    return *this;
  }
};

template <typename T> Col<T> Mat<T>::col(uword c) const {
  Col<T> out(n_rows);
  for (uword r = 0; r < n_rows; ++r) out[r] = (*this)(r, c);
  return out;
}
template <typename T> Mat<T> Mat<T>::cols(const Col<uword>& ix) const {
  Mat<T> out(n_rows, ix.n_elem);
  for (uword c = 0; c < ix.n_elem; ++c)
    for (uword r = 0; r < n_rows; ++r) out(r, c) = (*this)(r, ix[c]);
  return out;
}

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Row<double> rowvec;
typedef Mat<uword> umat;
typedef Col<uword> uvec;
typedef Row<uword> urowvec;

template <typename T> inline Row<T> operator-(const subview_row<T>& a, const subview_row<T>& b) {
  Row<T> out(a.m.n_cols);
  for (uword c = 0; c < a.m.n_cols; ++c) out[c] = a.m(a.r, c) - b.m(b.r, c);
  return out;
}
template <typename T> inline Row<T> operator+(const Row<T>& a, int s) {
  Row<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] + (T)s;
  return out;
}
template <typename T> inline Col<T> operator/(const Col<T>& a, const Col<T>& b) {
  Col<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] / b[i];
  return out;
}
// relational expression `v == k` -> 0/1 vector; find() lists the non-zero positions
inline uvec operator==(const uvec& a, uword k) {
  uvec out(a.n_elem);
  for (uword i = 0; i < a.n_elem; ++i) out[i] = (a[i] == k) ? 1 : 0;
  return out;
}
inline uvec find(const uvec& mask) {
  uword n = 0;
  for (uword i = 0; i < mask.n_elem; ++i) n += mask[i] != 0;
  uvec out(n);
  n = 0;
  for (uword i = 0; i < mask.n_elem; ++i) if (mask[i] != 0) out[n++] = i;
  return out;
}
template <typename T> inline uvec find_finite(const Mat<T>& x) {
  uword n = 0;
  for (uword i = 0; i < x.n_elem; ++i) n += std::isfinite(x[i]) ? 1 : 0;
  uvec out(n);
  n = 0;
  for (uword i = 0; i < x.n_elem; ++i) if (std::isfinite(x[i])) out[n++] = i;
  return out;
}
template <typename T> inline Row<T> square(const Row<T>& a) {
  Row<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] * a[i];
  return out;
}

namespace detail {
inline double accumulate(const double* X, uword n) {
  double acc1 = 0.0, acc2 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { acc1 += X[i]; acc2 += X[j]; }
  if (i < n) acc1 += X[i];
  return acc1 + acc2;
}
inline double mean_robust(const double* X, uword n) {
  double r = 0.0;
  for (uword i = 0; i < n; ++i) r += (X[i] - r) / double(i + 1);
  return r;
}
inline double mean(const double* X, uword n) {
  const double r = accumulate(X, n) / double(n);
  return std::isfinite(r) ? r : mean_robust(X, n);
}
inline double var_robust(const double* X, uword n) {
  double r_mean = X[0], r_var = 0.0;
  for (uword i = 1; i < n; ++i) {
    const double tmp = X[i] - r_mean;
    const double ip1 = double(i + 1);
    r_var = double(i - 1) / double(i) * r_var + (tmp * tmp) / ip1;
    r_mean = r_mean + tmp / ip1;
  }
  return r_var;
}
inline double var(const double* X, uword n) {   // norm_type 0: divide by n-1
  if (n < 2) return 0.0;
  const double m = mean(X, n);
  double acc2 = 0.0, acc3 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    const double ti = m - X[i], tj = m - X[j];
    acc2 += ti * ti + tj * tj;
    acc3 += ti + tj;
  }
  if (i < n) { const double ti = m - X[i]; acc2 += ti * ti; acc3 += ti; }
  const double v = (acc2 - acc3 * acc3 / double(n)) / double(n - 1);
  return std::isfinite(v) ? v : var_robust(X, n);
}
}  // namespace detail

template <typename T> inline T sum(const Mat<T>& x) { return (T)detail::accumulate(x.memptr(), x.n_elem); }
inline double stddev(const Mat<double>& v) { return std::sqrt(detail::var(v.memptr(), v.n_elem)); }
// stddev(X, norm_type, dim): dim = 1 -> one value per row (n_rows x 1)
inline Mat<double> stddev(const Mat<double>& X, uword norm_type, uword dim) {
  (void)norm_type;   // the reference only calls it with 0 (n-1)
  if (dim == 1) {
    Mat<double> out(X.n_rows, 1);
    std::vector<double> tmp((size_t)X.n_cols);
    for (uword r = 0; r < X.n_rows; ++r) {
      for (uword c = 0; c < X.n_cols; ++c) tmp[(size_t)c] = X(r, c);
      out[r] = std::sqrt(detail::var(tmp.data(), X.n_cols));
    }
    return out;
  }
  Mat<double> out(1, X.n_cols);
  for (uword c = 0; c < X.n_cols; ++c) out[c] = std::sqrt(detail::var(X.memptr() + c * X.n_rows, X.n_rows));
  return out;
}

template <typename V> inline V linspace(double start, double end, uword N) {
  V x(N);
  if (N == 1) { x[0] = (typename V::elem_type)end; return x; }
  const double delta = (end - start) / double(N - 1);
  for (uword i = 0; i < N; ++i) x[i] = (typename V::elem_type)(start + double(i) * delta);
  return x;
}

}  // namespace arma

namespace Rcpp {
// NumericVector(first, last): the only constructor the reference uses
struct NumericVector {
  std::vector<double> v;
  NumericVector() {}
  explicit NumericVector(int n) : v((size_t)n, 0.0) {}
  template <typename It> NumericVector(It first, It last) : v(first, last) {}
  size_t size() const { return v.size(); }
  double operator[](size_t i) const { return v[i]; }
};
}  // namespace Rcpp
