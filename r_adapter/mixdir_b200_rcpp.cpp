// Reference-side binding of the mixdir_b200 engine (NOT compiled in this repository: the image
// has no R / Rcpp).  A maintainer of const-ae/mixdir adds this file as
// src/mixdir_b200_rcpp.cpp, runs Rcpp::compileAttributes() (which regenerates
// src/RcppExports.cpp and R/RcppExports.R with the `_mixdir_mixdir_vi_b200_cpp` entry next to
// the three existing ones, src/RcppExports.cpp:62-72) and links with -lmixdir_b200.
//
// It follows the conventions of the existing glue: arguments arrive as Rcpp vectors that wrap
// the caller's SEXPs without copying, results are built with Rcpp::List / Rcpp::wrap, C++
// exceptions become R errors through BEGIN_RCPP/END_RCPP (generated).  All outputs are
// allocated here (R owns them); the CUDA library only sees raw pointers.
#include <Rcpp.h>

#include "mixdir_b200.h"

using namespace Rcpp;

// [[Rcpp::export]]
List mixdir_vi_b200_cpp(NumericMatrix X,            // codes 1..C_j, NA  (R/mixdir.R:117-118)
                        IntegerVector ncat,         // sapply(categories, length)
                        int n_latent, bool dp, double alpha1, double alpha2, double beta,
                        int max_iter, double epsilon,
                        NumericVector colsum_init,  // colSums(zeta_init)     (R/mixdir_vi.R:61)
                        NumericVector phi_init,     // unlist(phi_init): (j,k,r) order
                        int n_cat,                  // max(X, na.rm=TRUE)    (R/mixdir_vi.R:8)
                        IntegerVector devices) {    // integer(0): current device
  const int N = X.nrow(), J = X.ncol(), K = n_latent;
  mixdir_b200_handle* h = nullptr;
  int st = mixdir_b200_create_from_r_matrix(X.begin(), N, J, ncat.begin(), devices.size(),
                                            devices.size() ? devices.begin() : nullptr, &h);
  if (st != MIXDIR_B200_OK) stop("mixdir_b200: %s", mixdir_b200_last_error());

  mixdir_b200_fit_params p;
  p.n_latent = K;  p.dp = dp ? 1 : 0;  p.alpha1 = alpha1;  p.alpha2 = alpha2;  p.beta = beta;
  p.max_iter = max_iter;  p.epsilon = epsilon;  p.n_cat_bound = n_cat;

  const size_t nrag = mixdir_b200_ragged_size(J, ncat.begin(), K);
  NumericVector trace(max_iter, NA_REAL), lambda(K), prior(dp ? 2 * K : K), phi(nrag), U(nrag);
  NumericMatrix class_prob(N, K);
  IntegerVector pred_class(N);
  int n_iter = 0, converged = 0, warn = 0;
  double elbo = NA_REAL;
  st = mixdir_b200_fit(h, &p, colsum_init.begin(), phi_init.begin(), trace.begin(), &n_iter,
                       &converged, &elbo, lambda.begin(), prior.begin(), phi.begin(), U.begin(),
                       class_prob.begin(), pred_class.begin(), &warn);
  mixdir_b200_destroy(h);
  if (st == MIXDIR_B200_UNDERFLOW)  // the stop() of R/mixdir_vi.R:77-81, same text
    stop("There was an underflow in the calculation of zeta. Cannot continue.\n"
         "The problem probably came from the large number of columns in the input data (%d). "
         "Is it possible that you want to work on t(X)?", J);
  if (st != MIXDIR_B200_OK) stop("mixdir_b200: %s", mixdir_b200_last_error());
  if (warn & MIXDIR_B200_WARN_ELBO_DECREASED)  // R/mixdir_vi.R:105-108
    warning("The ELBO decreased. This should not happen, it might be due to numerical "
            "instabilities or a bug in the code. Please contact the maintainer to report this.\n");
  if (warn & MIXDIR_B200_WARN_DP_LAST_COMPONENT)  // R/mixdir_vi_dp.R:159-162
    warning("The model put considerable weight on the last component.\n"
            "            Consider re-running the model with an increased number of latent categories.");

  return List::create(_["converged"] = (bool)converged,
                      _["convergence"] = head(trace, n_iter),
                      _["ELBO"] = elbo, _["lambda"] = lambda, _["pred_class"] = pred_class,
                      _["class_prob"] = class_prob, _["U_flat"] = U, _["phi_flat"] = phi,
                      _["prior"] = prior);
}

// One round of find_defining_features() (R/predict.R:299-310): the quality loss for every
// remaining variable if it were removed as well, in ONE device pass instead of one
// predict.mixdir() call per variable.  X is the coded subsample X[subsample, ] (same coding as the
// training data: the levels of mixdir_obj$category_prob), `active` marks the remaining variables.
// Element ncol(X) + 1 of the result is the value for the active set itself (the final quality of
// R/predict.R:326-331).  measure: FALSE = "JS", TRUE = "ARI" (returned as 1 - ARI like the R code).
// [[Rcpp::export]]
NumericVector feature_divergence_b200_cpp(NumericMatrix X, IntegerVector ncat, NumericVector lambda,
                                          NumericVector category_prob_flat,  // unlist(category_prob)
                                          LogicalVector active, bool ari) {
  const int N = X.nrow(), J = X.ncol(), K = lambda.size();
  mixdir_b200_handle* h = nullptr;
  int st = mixdir_b200_create_from_r_matrix(X.begin(), N, J, ncat.begin(), 0, nullptr, &h);
  if (st != MIXDIR_B200_OK) stop("mixdir_b200: %s", mixdir_b200_last_error());
  std::vector<unsigned char> act(J);
  for (int j = 0; j < J; j++) act[j] = active[j] ? 1 : 0;
  NumericVector out(J + 1);
  st = mixdir_b200_feature_divergence(h, K, lambda.begin(), category_prob_flat.begin(), act.data(),
                                      ari ? 1 : 0, nullptr, 0, out.begin());
  mixdir_b200_destroy(h);
  if (st != MIXDIR_B200_OK) stop("mixdir_b200: %s", mixdir_b200_last_error());
  return out;  // a caller that loops (find_defining_features) keeps the handle instead: see api.py
}
