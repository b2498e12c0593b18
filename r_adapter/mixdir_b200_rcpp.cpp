// Reference-side binding of the mixdir_b200 engine.  A maintainer of const-ae/mixdir adds this
// file as src/mixdir_b200_rcpp.cpp, runs Rcpp::compileAttributes() (which regenerates
// src/RcppExports.cpp and R/RcppExports.R with the `_mixdir_*_b200_cpp` entries next to the three
// existing ones, src/RcppExports.cpp:62-72) and links with -lmixdir_b200.
// The image has no R: here the file is only parsed and type-checked against a stand-in Rcpp.h
// (tests/rcpp_stub, tests/test_abi.py::test_r_adapter_parses).
//
// It follows the conventions of the existing glue: arguments arrive as Rcpp vectors that wrap
// the caller's SEXPs without copying, results are built with Rcpp::List, C++ exceptions become R
// errors through BEGIN_RCPP/END_RCPP (generated).  All outputs are allocated here (R owns them);
// the CUDA library only sees raw pointers.
//
// The uploaded code matrix lives in an EXTERNAL POINTER (Rcpp::XPtr with a finalizer that calls
// mixdir_b200_destroy): mixdir() uploads X once and reuses the handle for every repetition
// (R/mixdir.R:120-134), find_defining_features() uploads its subsample once for all its rounds
// (R/predict.R:299-331), predict() uploads the new rows once.
#include <Rcpp.h>

#include <algorithm>
#include <vector>

#include "mixdir_b200.h"

using namespace Rcpp;

namespace {

void destroy_handle(mixdir_b200_handle* h) { mixdir_b200_destroy(h); }
typedef XPtr<mixdir_b200_handle, PreserveStorage, destroy_handle, true> HandlePtr;

mixdir_b200_handle* handle_of(SEXP xp) {
  HandlePtr p(xp);
  mixdir_b200_handle* h = p.get();
  if (!h) stop("mixdir_b200: the handle was released");
  return h;
}

void raise(int st, int n_features) {
  if (st == MIXDIR_B200_OK) return;
  if (st == MIXDIR_B200_UNDERFLOW)  // the stop() of R/mixdir_vi.R:77-81, same text
    stop("There was an underflow in the calculation of zeta. Cannot continue.\n"
         "The problem probably came from the large number of columns in the input data (%d). "
         "Is it possible that you want to work on t(X)?", n_features);
  stop("mixdir_b200: %s", mixdir_b200_last_error());
}

void warn_from_flags(int warn) {
  if (warn & MIXDIR_B200_WARN_ELBO_DECREASED)  // R/mixdir_vi.R:105-108
    warning("The ELBO decreased. This should not happen, it might be due to numerical "
            "instabilities or a bug in the code. Please contact the maintainer to report this.\n");
  if (warn & MIXDIR_B200_WARN_DP_LAST_COMPONENT)  // R/mixdir_vi_dp.R:159-162
    warning("The model put considerable weight on the last component.\n"
            "            Consider re-running the model with an increased number of latent categories.");
}

mixdir_b200_fit_params fit_params(int K, bool dp, double alpha1, double alpha2, double beta, int max_iter,
                                  double epsilon, int n_cat) {
  mixdir_b200_fit_params p;
  p.n_latent = K;  p.dp = dp ? 1 : 0;  p.alpha1 = alpha1;  p.alpha2 = alpha2;  p.beta = beta;
  p.max_iter = max_iter;  p.epsilon = epsilon;  p.n_cat_bound = n_cat;
  return p;
}

}  // namespace

// X: the numeric matrix mixdir() builds (codes 1..C_j, NA; R/mixdir.R:117-118), or new data coded
// against the training categories (R/predict.R:64-84).  ncat = sapply(categories, length).
// devices: integer(0) = the current device; 0:7 = row shards over eight GPUs of this process.
// [[Rcpp::export]]
SEXP upload_b200_cpp(NumericMatrix X, IntegerVector ncat, IntegerVector devices) {
  mixdir_b200_handle* h = nullptr;
  const int st = mixdir_b200_create_from_r_matrix(X.begin(), X.nrow(), X.ncol(), ncat.begin(),
                                                  (int)devices.size(),
                                                  devices.size() ? devices.begin() : nullptr, &h);
  raise(st, X.ncol());
  return HandlePtr(h, true);
}

// release the device memory now instead of at garbage collection
// [[Rcpp::export]]
void release_b200_cpp(SEXP handle) {
  HandlePtr p(handle);
  p.release();
}

// The body of mixdir_vi() / mixdir_vi_dp() (R/mixdir_vi.R:50-155, R/mixdir_vi_dp.R:56-185) for
// `repetitions` restarts at once (R/mixdir.R:120-134).  colsum_init: K x R matrix, column r =
// colSums(zeta_init of repetition r) (R/mixdir_vi.R:61); phi_init: ragged x R matrix, column r =
// unlist(phi_init of repetition r) in (j,k,r) order.  Returns the per-repetition scalars and traces,
// and the full result of the repetition mixdir() keeps (first largest ELBO).
// [[Rcpp::export]]
List mixdir_vi_b200_cpp(SEXP handle, IntegerVector ncat, int n_latent, bool dp, double alpha1,
                        double alpha2, double beta, int max_iter, double epsilon,
                        NumericMatrix colsum_init, NumericMatrix phi_init,
                        int n_cat) {  // max(X, na.rm=TRUE)    (R/mixdir_vi.R:8)
  mixdir_b200_handle* h = handle_of(handle);
  const int J = ncat.size(), K = n_latent, R = colsum_init.ncol();
  const int64_t N = mixdir_b200_n_rows(h);
  const size_t nrag = mixdir_b200_ragged_size(J, ncat.begin(), K);
  if (colsum_init.nrow() != K || phi_init.ncol() != R || (size_t)phi_init.nrow() != nrag)
    stop("mixdir_b200: colsum_init must be K x repetitions and phi_init ragged x repetitions");
  const mixdir_b200_fit_params p = fit_params(K, dp, alpha1, alpha2, beta, max_iter, epsilon, n_cat);
  const int mi = max_iter > 1 ? max_iter : 1;
  NumericMatrix trace(mi, R), lambda(K, R), prior(dp ? 2 * K : K, R), phi(nrag, R), U(nrag, R);
  std::fill(trace.begin(), trace.end(), NA_REAL);
  NumericVector elbo(R, NA_REAL);
  IntegerVector n_iter(R), converged(R), warn(R), status(R);
  NumericMatrix class_prob((int)N, K);
  IntegerVector pred_class((int)N);
  int best = -1;
  const int st = mixdir_b200_fit_batch(h, &p, R, colsum_init.begin(), phi_init.begin(), trace.begin(),
                                       n_iter.begin(), converged.begin(), elbo.begin(), lambda.begin(),
                                       prior.begin(), phi.begin(), U.begin(), warn.begin(),
                                       status.begin(), &best, class_prob.begin(), pred_class.begin());
  raise(st, J);
  for (int r = 0; r < R; r++) warn_from_flags(warn[r]);
  const NumericVector tr = trace(_, best), lam = lambda(_, best), pr = prior(_, best),
                      ub = U(_, best), pb = phi(_, best);
  const int nb = n_iter[best];
  const bool cb = converged[best] != 0;
  const double eb = elbo[best];
  return List::create(_["best"] = best + 1, _["ELBO_all"] = elbo, _["n_iter_all"] = n_iter,
                      _["converged"] = cb, _["convergence"] = head(tr, nb), _["ELBO"] = eb,
                      _["lambda"] = lam, _["pred_class"] = pred_class, _["class_prob"] = class_prob,
                      _["U_flat"] = ub, _["phi_flat"] = pb, _["prior"] = pr);
}

// The arithmetic of predict_class() (R/predict.R:88-103) on the rows of the handle; the name
// matching, the "(Missing)" level and the unseen-level warning of R/predict.R:39-86 stay in R.
// [[Rcpp::export]]
NumericMatrix predict_class_b200_cpp(SEXP handle, NumericVector lambda,
                                     NumericVector category_prob_flat) {  // unlist(category_prob)
  mixdir_b200_handle* h = handle_of(handle);
  const int K = lambda.size();
  const int64_t N = mixdir_b200_n_rows(h);
  NumericMatrix class_prob((int)N, K);
  const int st = mixdir_b200_predict(h, K, lambda.begin(), category_prob_flat.begin(),
                                     class_prob.begin(), nullptr);
  raise(st, 0);
  return class_prob;
}

// One round of find_defining_features() (R/predict.R:299-310): the quality loss for every
// remaining variable if it were removed as well, in ONE device pass instead of one
// predict.mixdir() call per variable.  The handle holds the coded subsample X[subsample, ] (same
// coding as the training data: the levels of mixdir_obj$category_prob) and is reused by every
// round; `active` marks the remaining variables.  Element n_features + 1 of the result is the value
// for the active set itself (the final quality of R/predict.R:326-331).  ari: FALSE = "JS", TRUE =
// "ARI" (returned as 1 - ARI like the R code).
// [[Rcpp::export]]
NumericVector feature_divergence_b200_cpp(SEXP handle, NumericVector lambda,
                                          NumericVector category_prob_flat, LogicalVector active,
                                          bool ari) {
  mixdir_b200_handle* h = handle_of(handle);
  const int J = active.size(), K = lambda.size();
  std::vector<unsigned char> act(J);
  for (int j = 0; j < J; j++) act[j] = active[j] ? 1 : 0;
  NumericVector out(J + 1);
  const int st = mixdir_b200_feature_divergence(h, K, lambda.begin(), category_prob_flat.begin(),
                                                act.data(), ari ? 1 : 0, nullptr, 0, out.begin());
  raise(st, J);
  return out;
}
