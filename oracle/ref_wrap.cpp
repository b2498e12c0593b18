// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// extern "C" doors into the reference's own native functions.  This file is
// linked with /root/reference/src/mixdir_impl.cpp compiled verbatim (where it
// lies, never copied) against oracle/shim/Rcpp.h; the result goes to
// oracle/_ref/libmixdir_ref.so (git-ignored).  See oracle/Makefile.
//
// The three functions are the reference's .Call targets
// (src/RcppExports.cpp:9-60, R/RcppExports.R:4-14).
#include <Rcpp.h>

using Rcpp::NumericMatrix;
using Rcpp::NumericVector;

// prototypes exactly as in /root/reference/src/mixdir_impl.cpp:15,58,114
double expec_log_x_cpp(NumericMatrix X, NumericVector phia, NumericMatrix zeta, int n_quest,
                       int n_latent, int n_cat);
NumericMatrix update_zeta_cpp(NumericMatrix zeta, NumericMatrix X, NumericVector phia,
                              NumericVector omega, int n_ind, int n_quest, int n_latent,
                              int n_cat);
NumericMatrix update_zeta_dp_cpp(NumericMatrix zeta, NumericMatrix X, NumericVector phia,
                                 NumericVector kappa1, NumericVector kappa2, int n_ind,
                                 int n_quest, int n_latent, int n_cat);

extern "C" {

// X: n_ind x n_quest column-major doubles (1..C_j, NaN = NA)
// phia: [n_quest, n_latent, n_cat_dim] column-major, NaN-padded (R/mixdir_vi.R:196-207)
// zeta: n_ind x n_latent column-major
double ref_expec_log_x(double* X, int n_ind, int n_quest, double* phia, int n_cat_dim,
                       double* zeta, int n_latent, int n_cat) {
  NumericMatrix Xm(X, n_ind, n_quest);
  NumericVector ph(phia, static_cast<std::ptrdiff_t>(n_quest) * n_latent * n_cat_dim);
  NumericMatrix zm(zeta, n_ind, n_latent);
  return expec_log_x_cpp(Xm, ph, zm, n_quest, n_latent, n_cat);
}

// zeta is overwritten in place, like the Rcpp original (mixdir_impl.cpp:83,86)
void ref_update_zeta(double* zeta, double* X, double* phia, int n_cat_dim, double* omega,
                     int n_ind, int n_quest, int n_latent, int n_cat) {
  NumericMatrix zm(zeta, n_ind, n_latent);
  NumericMatrix Xm(X, n_ind, n_quest);
  NumericVector ph(phia, static_cast<std::ptrdiff_t>(n_quest) * n_latent * n_cat_dim);
  NumericVector om(omega, n_latent);
  update_zeta_cpp(zm, Xm, ph, om, n_ind, n_quest, n_latent, n_cat);
}

void ref_update_zeta_dp(double* zeta, double* X, double* phia, int n_cat_dim, double* kappa1,
                        double* kappa2, int n_ind, int n_quest, int n_latent, int n_cat) {
  NumericMatrix zm(zeta, n_ind, n_latent);
  NumericMatrix Xm(X, n_ind, n_quest);
  NumericVector ph(phia, static_cast<std::ptrdiff_t>(n_quest) * n_latent * n_cat_dim);
  NumericVector k1(kappa1, n_latent);
  NumericVector k2(kappa2, n_latent);
  update_zeta_dp_cpp(zm, Xm, ph, k1, k2, n_ind, n_quest, n_latent, n_cat);
}

// Route oracle_fit's three native steps through the reference's own functions
// (see the hook declarations in mixdir_oracle.c).  Only meaningful inside
// oracle/_ref/libmixdir_ref.so, which links both.
typedef void (*hook_update_zeta_t)(double*, const double*, const double*, int, const double*, int,
                                   int, int, int);
typedef void (*hook_update_zeta_dp_t)(double*, const double*, const double*, int, const double*,
                                      const double*, int, int, int, int);
typedef double (*hook_expec_log_x_t)(const double*, int, const double*, int, const double*, int,
                                     int, int);
extern hook_update_zeta_t oracle_hook_update_zeta;
extern hook_update_zeta_dp_t oracle_hook_update_zeta_dp;
extern hook_expec_log_x_t oracle_hook_expec_log_x;

static void hk_update_zeta(double* zeta, const double* X, const double* phia, int cdim,
                           const double* omega, int n_ind, int n_quest, int n_latent, int n_cat) {
  ref_update_zeta(zeta, const_cast<double*>(X), const_cast<double*>(phia), cdim,
                  const_cast<double*>(omega), n_ind, n_quest, n_latent, n_cat);
}
static void hk_update_zeta_dp(double* zeta, const double* X, const double* phia, int cdim,
                              const double* k1, const double* k2, int n_ind, int n_quest,
                              int n_latent, int n_cat) {
  ref_update_zeta_dp(zeta, const_cast<double*>(X), const_cast<double*>(phia), cdim,
                     const_cast<double*>(k1), const_cast<double*>(k2), n_ind, n_quest, n_latent,
                     n_cat);
}
static double hk_expec_log_x(const double* X, int n_ind, const double* phia, int cdim,
                             const double* zeta, int n_quest, int n_latent, int n_cat) {
  return ref_expec_log_x(const_cast<double*>(X), n_ind, n_quest, const_cast<double*>(phia), cdim,
                         const_cast<double*>(zeta), n_latent, n_cat);
}
void ref_install_hooks(int on) {
  oracle_hook_update_zeta = on ? hk_update_zeta : nullptr;
  oracle_hook_update_zeta_dp = on ? hk_update_zeta_dp : nullptr;
  oracle_hook_expec_log_x = on ? hk_expec_log_x : nullptr;
}

}  // extern "C"
