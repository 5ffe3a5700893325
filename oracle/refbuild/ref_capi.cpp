// C entry points around the reference's own native routines -- TEST INFRASTRUCTURE ONLY.
//
// Linked with the object compiled from the UNMODIFIED /root/reference/src/utils.cpp (against the
// stand-in headers in oracle/shim/) into oracle/_ref/libchromvar_ref.so.  The wrappers mirror what
// the reference's generated glue does (src/RcppExports.cpp:59-119): copy the arguments into
// Armadillo objects, call the routine, copy the result out.  Also here: R's default uniform
// generator (Mersenne-Twister, R src/main/RNG.c: MT_genrand, fixup, and set.seed's
// linear-congruential scrambling) because the reference draws through R's unif_rand()
// (src/utils.cpp:72); pinned in tests/test_ref_native.py to well-known set.seed()/runif() values.
#include <RcppArmadillo.h>
#include <cstring>

// ---- prototypes of the reference routines (src/utils.cpp:9,28,38,84,107) -----------------------
Rcpp::NumericVector row_sds(arma::mat& X, bool na_rm);
Rcpp::NumericVector row_sds_perm(arma::mat& X, bool na_rm);
arma::urowvec ProbSampleReplace(int size, arma::vec prob);
arma::umat bg_sample_helper(arma::uvec bin_membership, arma::mat bin_p, arma::vec bin_density,
                            arma::uword niterations);
arma::mat euc_dist(arma::mat x);

// ---- R's Mersenne-Twister ------------------------------------------------------------------------
namespace {
const int MT_N = 624, MT_M = 397;
uint32_t mt[MT_N];
int mti = MT_N + 1;
uint64_t draws = 0;

void mt_sgenrand(uint32_t seed) {
  for (int i = 0; i < MT_N; i++) {
    mt[i] = seed & 0xffff0000u;
    seed = 69069u * seed + 1;
    mt[i] |= (seed & 0xffff0000u) >> 16;
    seed = 69069u * seed + 1;
  }
  mti = MT_N;
}

double mt_genrand() {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t y;
  if (mti >= MT_N) {
    if (mti == MT_N + 1) mt_sgenrand(4357);
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 0x1];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 0x1];
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 0x1];
    mti = 0;
  }
  y = mt[mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return (double)y * 2.3283064365386963e-10;   // [0,1)
}
}  // namespace

extern "C" double unif_rand(void) {
  const double i2_32m1 = 2.328306437080797e-10;
  double x = mt_genrand();
  ++draws;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

extern "C" {

// set.seed(seed) for kind "Mersenne-Twister": 50 LCG warm-up steps, then 625 state words, the
// first of which (the position) is reset to 624 so the next draw regenerates the block.
void ref_set_seed(uint32_t seed) {
  for (int j = 0; j < 50; j++) seed = 69069u * seed + 1;
  uint32_t dummy0 = 0;
  for (int j = 0; j < 625; j++) {
    seed = 69069u * seed + 1;
    if (j == 0) dummy0 = seed; else mt[j - 1] = seed;
  }
  (void)dummy0;
  mti = MT_N;
  draws = 0;
}

void ref_unif_rand(int64_t n, double* out) { for (int64_t i = 0; i < n; i++) out[i] = unif_rand(); }
uint64_t ref_draws(void) { return draws; }

static arma::mat to_mat(const double* x, int64_t r, int64_t c) {
  arma::mat m((arma::uword)r, (arma::uword)c);
  std::memcpy(m.memptr(), x, sizeof(double) * (size_t)(r * c));
  return m;
}

// X: K x N column-major; out: K
int ref_row_sds(const double* X, int64_t K, int64_t N, int na_rm, double* out) {
  try {
    arma::mat m = to_mat(X, K, N);
    Rcpp::NumericVector r = row_sds(m, na_rm != 0);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
    return 0;
  } catch (...) { return 1; }
}

int ref_row_sds_perm(const double* X, int64_t K, int64_t N, int na_rm, double* out) {
  try {
    arma::mat m = to_mat(X, K, N);
    Rcpp::NumericVector r = row_sds_perm(m, na_rm != 0);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
    return 0;
  } catch (...) { return 1; }
}

// out: size indices, 1-based (src/utils.cpp:76)
int ref_prob_sample_replace(int size, const double* prob, int64_t n, int64_t* out) {
  try {
    arma::vec p((arma::uword)n);
    std::memcpy(p.memptr(), prob, sizeof(double) * (size_t)n);
    arma::urowvec r = ProbSampleReplace(size, p);
    for (arma::uword i = 0; i < r.n_elem; i++) out[i] = (int64_t)r[i];
    return 0;
  } catch (...) { return 1; }
}

// bin_membership: 0-based (the R caller subtracts 1, R/background_peaks.R:151-152); bin_p: nb x nb
// column-major; out: P x niterations column-major, 1-based
int ref_bg_sample_helper(const int64_t* bin_membership, int64_t P, const double* bin_p, int64_t nb,
                         const double* bin_density, int64_t niterations, int64_t* out) {
  try {
    arma::uvec bm((arma::uword)P);
    for (int64_t i = 0; i < P; i++) bm[i] = (arma::uword)bin_membership[i];
    arma::mat bp = to_mat(bin_p, nb, nb);
    arma::vec bd((arma::uword)nb);
    std::memcpy(bd.memptr(), bin_density, sizeof(double) * (size_t)nb);
    arma::umat r = bg_sample_helper(bm, bp, bd, (arma::uword)niterations);
    for (arma::uword i = 0; i < r.n_elem; i++) out[i] = (int64_t)r[i];
    return 0;
  } catch (...) { return 1; }
}

// x: n x d column-major; out: n x n
int ref_euc_dist(const double* x, int64_t n, int64_t d, double* out) {
  try {
    arma::mat r = euc_dist(to_mat(x, n, d));
    std::memcpy(out, r.memptr(), sizeof(double) * (size_t)(n * n));
    return 0;
  } catch (...) { return 1; }
}

}  // extern "C"
