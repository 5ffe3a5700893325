// Stand-in for <RcppArmadilloExtensions/sample.h> -- TEST INFRASTRUCTURE ONLY (see ../RcppArmadillo.h).
// Only the call the reference makes (src/utils.cpp:30): sample(x, size, replace = true) without
// probabilities, whose published algorithm is index[i] = floor(n * unif_rand()), result x[index[i]].
#pragma once
#include <stdexcept>
#include "../RcppArmadillo.h"

namespace Rcpp { namespace RcppArmadillo {
template <class T> T sample(const T& x, const int size, const bool replace,
                            NumericVector prob_ = NumericVector(0)) {
  if (!replace || prob_.size() != 0) throw std::runtime_error("sample stand-in: only replace = TRUE without prob");
  const int nOrig = (int)x.size();
  T ret(size);
  for (int ii = 0; ii < size; ++ii) {
    const arma::uword jj = (arma::uword)(nOrig * unif_rand());
    ret[ii] = x[jj];
  }
  return ret;
}
}}  // namespace Rcpp::RcppArmadillo
