/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the shipped product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product (mixdir_b200/csrc) never links it.
 *
 * CPU restatement, in plain C / fp64, of the reference's variational-inference
 * hot path.  Every function cites the reference lines it follows
 * (paths under /root/reference).
 *
 * Pinning status: the three native functions below (oracle_update_zeta,
 * oracle_update_zeta_dp, oracle_expec_log_x) are checked in tests against the
 * reference's OWN src/mixdir_impl.cpp compiled verbatim (oracle/_ref, see
 * oracle/Makefile).  The R-level loop (R/mixdir_vi.R, R/mixdir_vi_dp.R) has no
 * runnable form here (no R in this image) and the reference ships no numeric
 * golden files, so the loop restatement is pinned only by the reference's test
 * invariants and README known answers (soft, R-RNG dependent).  digamma /
 * lgamma VALUES come from R's nmath in the reference, which is not vendored:
 * "parity unpinned" for those values beyond the mpmath check in tests/.
 *
 * Conventions shared with the reference:
 *   X      n_ind x n_quest, column-major doubles, values 1..C_j, NaN = NA
 *          (R/mixdir.R:117-118)
 *   zeta   n_ind x n_latent, column-major
 *   phia   [n_quest, n_latent, cdim] column-major, index j + k*J + r*J*K,
 *          NaN-padded for r >= C_j (R/mixdir_vi.R:196-207)
 *   phi    ragged R list phi[[j]][[k]][r], flattened here in (j, k, r) order:
 *          element (j,k,r) at  K*sum_{j'<j} C_j'  +  k*C_j  +  r
 *   R's sum()/colSums()/rowSums()/cumsum()/prod() accumulate in long double;
 *   this file does the same wherever the reference calls them.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* special functions                                                          */
/* ------------------------------------------------------------------------- */

/* digamma for x > 0 (all call sites have x >= beta, alpha > 0): upward
 * recurrence psi(x) = psi(x+1) - 1/x until x >= 10, then the asymptotic series
 * ln x - 1/(2x) - sum_n B_2n / (2n x^2n).  Stands in
 * for R::digamma (src/mixdir_impl.cpp:26,32,59,70,80,83,127,135,145,148) and
 * base-R digamma (R/mixdir_vi.R:165-191, R/mixdir_vi_dp.R:198-219). */
double oracle_digamma(double x) {
  if (!(x > 0.0)) return NAN;
  double acc = 0.0;
  while (x < 10.0) {
    acc -= 1.0 / x;
    x += 1.0;
  }
  const double inv = 1.0 / x;
  const double inv2 = inv * inv;
  /* B_2n/(2n): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12 */
  double series = 1.0 / 12.0;
  series = series * inv2 - 691.0 / 32760.0;
  series = series * inv2 + 1.0 / 132.0;
  series = series * inv2 - 1.0 / 240.0;
  series = series * inv2 + 1.0 / 252.0;
  series = series * inv2 - 1.0 / 120.0;
  series = series * inv2 + 1.0 / 12.0;
  return acc + (log(x) - 0.5 * inv - series * inv2);
}

double oracle_lgamma(double x) { return lgamma(x); }

/* ------------------------------------------------------------------------- */
/* the three native functions (src/mixdir_impl.cpp)                           */
/* ------------------------------------------------------------------------- */

static inline int is_na(double v) { return isnan(v); }

/* psi(sum_{r<n_cat, non-NA} phia[j,k,r]) for all (j,k): the pre-pass of
 * src/mixdir_impl.cpp:61-72 (and :118-129, :20-26).  Note the loop bound is
 * n_cat = max(X), NOT the third dimension of phia. */
static void dg_phi_sums(const double* phia, int J, int K, int n_cat, double* out) {
  for (int k = 0; k < K; k++) {
    for (int j = 0; j < J; j++) {
      double s = 0;
      for (int r = 0; r < n_cat; r++) {
        double v = phia[j + (size_t)k * J + (size_t)r * J * K];
        if (!is_na(v)) s += v;
      }
      out[j + (size_t)k * J] = oracle_digamma(s);
    }
  }
}

/* psi(phia[j,k,r]) for every non-NA entry: the reference evaluates this per
 * (i,j,k) cell (mixdir_impl.cpp:80); the value depends only on (j,k,r), so the
 * restatement tabulates it once.  Same numbers, same order of additions. */
static double* dg_phi_table(const double* phia, int J, int K, int cdim) {
  size_t n = (size_t)J * K * cdim;
  double* t = (double*)malloc(n * sizeof(double));
  for (size_t i = 0; i < n; i++) t[i] = is_na(phia[i]) ? NAN : oracle_digamma(phia[i]);
  return t;
}

/* src/mixdir_impl.cpp:57-87 update_zeta_cpp.  zeta is overwritten in place
 * (un-normalised). */
void oracle_update_zeta(double* zeta, const double* X, const double* phia, int cdim,
                        const double* omega, int n_ind, int n_quest, int n_latent, int n_cat) {
  const int J = n_quest, K = n_latent;
  double so = 0; /* Rcpp::sum accumulates in double (mixdir_impl.cpp:59) */
  for (int k = 0; k < K; k++) so += omega[k];
  const double dg_sum_omega = oracle_digamma(so);
  double* sums = (double*)malloc((size_t)J * K * sizeof(double));
  dg_phi_sums(phia, J, K, n_cat, sums);
  double* tab = dg_phi_table(phia, J, K, cdim);
  for (int k = 0; k < K; k++) {
    const double dgo = oracle_digamma(omega[k]);
    for (int i = 0; i < n_ind; i++) {
      double acc = 0; /* mixdir_impl.cpp:77-82 */
      for (int j = 0; j < J; j++) {
        double x = X[i + (size_t)j * n_ind];
        if (!is_na(x)) {
          size_t idx = (size_t)j + (size_t)k * J + (size_t)(x - 1) * J * K;
          acc += tab[idx] - sums[j + (size_t)k * J];
        }
      }
      zeta[i + (size_t)k * n_ind] = exp(dgo - dg_sum_omega + acc - 1); /* :83 */
    }
  }
  free(tab);
  free(sums);
}

/* src/mixdir_impl.cpp:113-152 update_zeta_dp_cpp. */
void oracle_update_zeta_dp(double* zeta, const double* X, const double* phia, int cdim,
                           const double* kappa1, const double* kappa2, int n_ind, int n_quest,
                           int n_latent, int n_cat) {
  const int J = n_quest, K = n_latent;
  double* sums = (double*)malloc((size_t)J * K * sizeof(double));
  dg_phi_sums(phia, J, K, n_cat, sums);
  double* tab = dg_phi_table(phia, J, K, cdim);
  double* skp = (double*)calloc(K, sizeof(double)); /* :131-138 */
  for (int k = 0; k < K; k++) {
    double s = 0;
    for (int kp = 0; kp < k; kp++)
      s += oracle_digamma(kappa2[kp]) - oracle_digamma(kappa1[kp] + kappa2[kp]);
    skp[k] = s;
  }
  for (int k = 0; k < K; k++) {
    const double a = oracle_digamma(kappa1[k]);
    const double b = oracle_digamma(kappa1[k] + kappa2[k]);
    for (int i = 0; i < n_ind; i++) {
      double acc = 0;
      for (int j = 0; j < J; j++) {
        double x = X[i + (size_t)j * n_ind];
        if (!is_na(x)) {
          size_t idx = (size_t)j + (size_t)k * J + (size_t)(x - 1) * J * K;
          acc += tab[idx] - sums[j + (size_t)k * J];
        }
      }
      zeta[i + (size_t)k * n_ind] = exp(a - b + skp[k] + acc - 1); /* :148 */
    }
  }
  free(skp);
  free(tab);
  free(sums);
}

/* src/mixdir_impl.cpp:14-38 expec_log_x_cpp (NA cell contributes +1 per
 * class, pinned by tests/testthat/test_cpp.R:5-11). */
double oracle_expec_log_x(const double* X, int n_ind, const double* phia, int cdim,
                          const double* zeta, int n_quest, int n_latent, int n_cat) {
  const int J = n_quest, K = n_latent;
  double* tab = dg_phi_table(phia, J, K, cdim);
  double result = 0;
  for (int j = 0; j < J; j++) {
    for (int k = 0; k < K; k++) {
      double s = 0;
      for (int r = 0; r < n_cat; r++) {
        double v = phia[j + (size_t)k * J + (size_t)r * J * K];
        if (!is_na(v)) s += v;
      }
      const double dg_sum = oracle_digamma(s);
      for (int i = 0; i < n_ind; i++) {
        double x = X[i + (size_t)j * n_ind];
        if (is_na(x)) {
          result += 1;
        } else {
          size_t idx = (size_t)j + (size_t)k * J + (size_t)(x - 1) * J * K;
          result += zeta[i + (size_t)k * n_ind] * (tab[idx] - dg_sum);
        }
      }
    }
  }
  free(tab);
  return result;
}

/* Optional hooks: when set (oracle/_ref installs them, see ref_wrap.cpp
 * ref_install_hooks), oracle_fit calls the REFERENCE'S OWN native functions
 * (src/mixdir_impl.cpp compiled verbatim) instead of the restatements above.
 * That is the "reference" CPU baseline of bench.py. */
typedef void (*hook_update_zeta_t)(double*, const double*, const double*, int, const double*, int,
                                   int, int, int);
typedef void (*hook_update_zeta_dp_t)(double*, const double*, const double*, int, const double*,
                                      const double*, int, int, int, int);
typedef double (*hook_expec_log_x_t)(const double*, int, const double*, int, const double*, int,
                                     int, int);
hook_update_zeta_t oracle_hook_update_zeta = 0;
hook_update_zeta_dp_t oracle_hook_update_zeta_dp = 0;
hook_expec_log_x_t oracle_hook_expec_log_x = 0;

/* ------------------------------------------------------------------------- */
/* R-level helpers                                                            */
/* ------------------------------------------------------------------------- */

static size_t phi_off(const int* ncat, int K, int j) {
  size_t o = 0;
  for (int jj = 0; jj < j; jj++) o += (size_t)K * ncat[jj];
  return o;
}

/* R/mixdir_vi.R:196-207 conv_phi_to_array */
static void conv_phi_to_array(const double* phi, const int* ncat, int J, int K, int cdim,
                              double* phia) {
  size_t n = (size_t)J * K * cdim;
  for (size_t i = 0; i < n; i++) phia[i] = NAN;
  size_t o = 0;
  for (int j = 0; j < J; j++) {
    for (int k = 0; k < K; k++)
      for (int r = 0; r < ncat[j]; r++)
        phia[j + (size_t)k * J + (size_t)r * J * K] = phi[o + (size_t)k * ncat[j] + r];
    o += (size_t)K * ncat[j];
  }
}

/* R colSums(): long double accumulation per column */
static void col_sums(const double* m, int nrow, int ncol, double* out) {
  for (int c = 0; c < ncol; c++) {
    long double s = 0;
    for (int i = 0; i < nrow; i++) s += m[i + (size_t)c * nrow];
    out[c] = (double)s;
  }
}

static double sum_ld(const double* v, int n) {
  long double s = 0;
  for (int i = 0; i < n; i++) s += v[i];
  return (double)s;
}

/* R/mixdir_vi.R:164-166 */
static double expec_log_lambda(const double* omega, double alpha, int K) {
  const double dso = oracle_digamma(sum_ld(omega, K));
  long double sl = 0, st = 0, sa = 0;
  for (int k = 0; k < K; k++) {
    sa += alpha;
    sl += lgamma(alpha);
    st += (alpha - 1) * (oracle_digamma(omega[k]) - dso);
  }
  return lgamma((double)sa) - (double)sl + (double)st;
}

/* R/mixdir_vi.R:172-174 */
static double expec_log_ujk(const double* phi_jk, double beta, int C) {
  const double dsp = oracle_digamma(sum_ld(phi_jk, C));
  long double sb = 0, sl = 0, st = 0;
  for (int r = 0; r < C; r++) {
    sb += beta;
    sl += lgamma(beta);
    st += (beta - 1) * (oracle_digamma(phi_jk[r]) - dsp);
  }
  return lgamma((double)sb) - (double)sl + (double)st;
}

/* R/mixdir_vi.R:177-179 entrop_omega and :190-192 entrop_phi (same formula) */
static double entrop_dirichlet(const double* w, int n) {
  const double sw = sum_ld(w, n);
  const double dsw = oracle_digamma(sw);
  long double sl = 0, st = 0;
  for (int k = 0; k < n; k++) {
    sl += lgamma(w[k]);
    st += (w[k] - 1) * (oracle_digamma(w[k]) - dsw);
  }
  return -lgamma(sw) + (double)sl - (double)st;
}

/* R/mixdir_vi.R:181-184 entrop_zeta, for row i of a column-major matrix */
static double entrop_zeta_row(const double* zeta, int n_ind, int K, int i) {
  long double s = 0;
  for (int k = 0; k < K; k++) {
    double z = zeta[i + (size_t)k * n_ind];
    if (z <= DBL_MIN) z = DBL_MIN;
    s += z * log(z);
  }
  return -(double)s;
}

/* R/mixdir_vi_dp.R:196-200 */
static double dp_expec_log_v(const double* ka, const double* kb, double a1, double a2, int K) {
  long double s = 0;
  for (int k = 0; k < K; k++) {
    const double dab = oracle_digamma(ka[k] + kb[k]);
    s += (a1 - 1) * (oracle_digamma(ka[k]) - dab) + (a2 - 1) * (oracle_digamma(kb[k]) - dab);
  }
  return (double)s;
}

/* R/mixdir_vi_dp.R:202-211, row i */
static double dp_expec_log_zi(const double* zeta, int n_ind, int K, int i, const double* e1,
                              const double* e2) {
  long double s = 0;
  for (int k = 0; k < K; k++) {
    const double zk = zeta[i + (size_t)k * n_ind];
    if (k != K - 1) {
      long double tail = 0;
      for (int l = k + 1; l < K; l++) tail += zeta[i + (size_t)l * n_ind];
      s += (double)tail * e2[k] + zk * e1[k];
    } else {
      s += zk * e1[k];
    }
  }
  return (double)s;
}

/* R/mixdir_vi_dp.R:214-221 */
static double dp_entrop_v(const double* ka, const double* kb, int K) {
  long double s = 0;
  for (int k = 0; k < K; k++) {
    s += lgamma(ka[k]) + lgamma(kb[k]) - lgamma(ka[k] + kb[k]) -
         (ka[k] - 1) * oracle_digamma(ka[k]) - (kb[k] - 1) * oracle_digamma(kb[k]) +
         (ka[k] + kb[k] - 2) * oracle_digamma(ka[k] + kb[k]);
  }
  return (double)s;
}

/* kappa update, R/mixdir_vi_dp.R:68-71 (and :148-151) */
static void dp_kappa(const double* zeta, int n_ind, int K, double a1, double a2, double* k1,
                     double* k2) {
  col_sums(zeta, n_ind, K, k1);
  for (int k = 0; k < K; k++) k1[k] += a1;
  /* summed_up_phi[i,k] = sum_{l>k} zeta[i,l]: rev(cumsum(rev(row))) shifted */
  double* su = (double*)malloc((size_t)n_ind * K * sizeof(double));
  for (int i = 0; i < n_ind; i++) {
    long double c = 0;
    su[i + (size_t)(K - 1) * n_ind] = 0;
    for (int k = K - 1; k >= 1; k--) {
      c += zeta[i + (size_t)k * n_ind];
      su[i + (size_t)(k - 1) * n_ind] = (double)c;
    }
  }
  col_sums(su, n_ind, K, k2);
  for (int k = 0; k < K; k++) k2[k] += a2;
  free(su);
}

/* final pass shared by R/mixdir_vi.R:136-148, R/mixdir_vi_dp.R:165-177 and
 * R/predict.R:88-103: prob_z[i,k] = lambda_k * exp(sum_j log(NA ? 1 : U[j,k,x_ij])),
 * row-normalised, no max shift; pred_class = which.max (first maximum, 1-based;
 * 0 if the row is all-NaN, where R would return integer(0)). */
void oracle_predict(const double* X, int n_ind, int n_quest, const int* ncat, int n_latent,
                    const double* lambda, const double* U, double* class_prob,
                    int* pred_class) {
  const int J = n_quest, K = n_latent;
  for (int k = 0; k < K; k++) {
    for (int i = 0; i < n_ind; i++) {
      long double s = 0; /* rowSums */
      size_t o = 0;
      for (int j = 0; j < J; j++) {
        double x = X[i + (size_t)j * n_ind];
        double u = is_na(x) ? 1.0 : U[o + (size_t)k * ncat[j] + (size_t)(x - 1)];
        s += log(u);
        o += (size_t)K * ncat[j];
      }
      class_prob[i + (size_t)k * n_ind] = lambda[k] * exp((double)s);
    }
  }
  for (int i = 0; i < n_ind; i++) {
    long double rs = 0;
    for (int k = 0; k < K; k++) rs += class_prob[i + (size_t)k * n_ind];
    int best = 0;
    double bestv = -INFINITY;
    for (int k = 0; k < K; k++) {
      double p = class_prob[i + (size_t)k * n_ind] / (double)rs;
      class_prob[i + (size_t)k * n_ind] = p;
      if (!isnan(p) && p > bestv) {
        bestv = p;
        best = k + 1;
      }
    }
    if (pred_class) pred_class[i] = best;
  }
}

/* ------------------------------------------------------------------------- */
/* the VI drivers                                                             */
/* ------------------------------------------------------------------------- */

/* Literal restatement of mixdir_vi (R/mixdir_vi.R:2-156) for dp == 0 and of
 * mixdir_vi_dp (R/mixdir_vi_dp.R:2-187) for dp != 0, with explicit
 * zeta_init / phi_init ("same initialisation"; the R defaults draw them from
 * R's RNG, R/mixdir_vi.R:31-39).
 *
 * outputs: elbo_hist[max_iter] (first *n_iter entries valid), lambda[K],
 *   class_prob[N*K col-major], pred_class[N], U and phi_out ragged like phi,
 *   prior_out: omega[K] (dp==0) or kappa1[K],kappa2[K] (dp!=0),
 *   warn_flags bit0 = "ELBO decreased" warning fired, bit1 = DP last component
 *   > 0.01 warning (R/mixdir_vi_dp.R:159-162).
 * returns 0, or 1 on the zeta underflow stop() (R/mixdir_vi.R:77-81). */
int oracle_fit(const double* X, int n_ind, int n_quest, const int* ncat, int n_latent, int dp,
               double alpha1, double alpha2, double beta, int max_iter, double epsilon,
               const double* zeta_init, const double* phi_init, double* elbo_hist, int* n_iter,
               int* converged_out, double* lambda, double* class_prob, int* pred_class,
               double* U, double* prior_out, double* phi_out, int* warn_flags) {
  const int N = n_ind, J = n_quest, K = n_latent;
  int cdim = 0;
  size_t nphi = 0;
  for (int j = 0; j < J; j++) {
    if (ncat[j] > cdim) cdim = ncat[j];
    nphi += (size_t)K * ncat[j];
  }
  double mx = -INFINITY; /* n_cat <- max(X, na.rm=TRUE)  (R/mixdir_vi.R:8) */
  for (size_t i = 0; i < (size_t)N * J; i++)
    if (!is_na(X[i]) && X[i] > mx) mx = X[i];
  const int n_cat = (int)mx;

  double* zeta = (double*)malloc((size_t)N * K * sizeof(double));
  memcpy(zeta, zeta_init, (size_t)N * K * sizeof(double));
  double* phi = (double*)malloc(nphi * sizeof(double));
  memcpy(phi, phi_init, nphi * sizeof(double));
  double* phia = (double*)malloc((size_t)J * K * cdim * sizeof(double));
  conv_phi_to_array(phi, ncat, J, K, cdim, phia);
  double* omega = (double*)calloc(K, sizeof(double));
  double* kappa2 = (double*)calloc(K, sizeof(double));
  double* cs = (double*)calloc(K, sizeof(double));
  double* tmpcol = (double*)malloc((size_t)N * sizeof(double));
  double* e1 = (double*)calloc(K, sizeof(double));
  double* e2 = (double*)calloc(K, sizeof(double));
  int* ord = (int*)calloc(K, sizeof(int));

  int iter = 1, converged = 0, status = 0, warn = 0;
  double elbo = NAN;
  while (iter <= max_iter && !converged) {
    if (!dp) {
      col_sums(zeta, N, K, omega); /* R/mixdir_vi.R:61 */
      for (int k = 0; k < K; k++) omega[k] += alpha1;
      (oracle_hook_update_zeta ? oracle_hook_update_zeta : oracle_update_zeta)(
          zeta, X, phia, cdim, omega, N, J, K, n_cat); /* :76 */
    } else {
      dp_kappa(zeta, N, K, alpha1, alpha2, omega, kappa2); /* R/mixdir_vi_dp.R:68-71 */
      (oracle_hook_update_zeta_dp ? oracle_hook_update_zeta_dp : oracle_update_zeta_dp)(
          zeta, X, phia, cdim, omega, kappa2, N, J, K, n_cat); /* :90 */
    }
    /* underflow guard + normalise (R/mixdir_vi.R:77-82, R/mixdir_vi_dp.R:91-96) */
    int under = 0;
    for (int i = 0; i < N; i++) {
      long double rs = 0;
      for (int k = 0; k < K; k++) rs += zeta[i + (size_t)k * N];
      tmpcol[i] = (double)rs;
      if (tmpcol[i] == 0) under = 1;
    }
    if (under) {
      status = 1;
      break;
    }
    for (int k = 0; k < K; k++)
      for (int i = 0; i < N; i++) zeta[i + (size_t)k * N] /= tmpcol[i];

    if (dp) { /* zeta <- zeta[, order(-colSums(zeta))]  (R/mixdir_vi_dp.R:97), stable */
      col_sums(zeta, N, K, cs);
      for (int k = 0; k < K; k++) ord[k] = k;
      for (int a = 1; a < K; a++) { /* insertion sort: stable, descending by cs */
        int v = ord[a], b = a - 1;
        while (b >= 0 && cs[ord[b]] < cs[v]) {
          ord[b + 1] = ord[b];
          b--;
        }
        ord[b + 1] = v;
      }
      double* z2 = (double*)malloc((size_t)N * K * sizeof(double));
      for (int k = 0; k < K; k++)
        memcpy(z2 + (size_t)k * N, zeta + (size_t)ord[k] * N, (size_t)N * sizeof(double));
      free(zeta);
      zeta = z2;
    }

    /* phi M-step (R/mixdir_vi.R:85-91; dp :101-107) */
    {
      size_t o = 0;
      for (int j = 0; j < J; j++) {
        for (int k = 0; k < K; k++) {
          for (int r = 0; r < ncat[j]; r++) {
            long double s = 0; /* sum(zeta[,k] * (X[,j]==r), na.rm=TRUE) */
            for (int i = 0; i < N; i++) {
              double x = X[i + (size_t)j * N];
              if (is_na(x)) continue;
              s += zeta[i + (size_t)k * N] * (x == (double)(r + 1) ? 1.0 : 0.0);
            }
            phi[o + (size_t)k * ncat[j] + r] = (double)s + beta;
          }
        }
        o += (size_t)K * ncat[j];
      }
    }
    conv_phi_to_array(phi, ncat, J, K, cdim, phia);

    /* ELBO (R/mixdir_vi.R:94-102; dp R/mixdir_vi_dp.R:110-118) */
    double tA, tB, tC, tD, tE, tF, tG;
    if (!dp) {
      tA = expec_log_lambda(omega, alpha1, K);
      const double dso = oracle_digamma(sum_ld(omega, K));
      for (int k = 0; k < K; k++) e1[k] = oracle_digamma(omega[k]) - dso;
      long double sB = 0;
      for (int i = 0; i < N; i++) { /* expec_log_zi, R/mixdir_vi.R:168-170 */
        long double s = 0;
        for (int k = 0; k < K; k++) s += zeta[i + (size_t)k * N] * e1[k];
        sB += (double)s;
      }
      tB = (double)sB;
      tE = entrop_dirichlet(omega, K);
    } else {
      tA = dp_expec_log_v(omega, kappa2, alpha1, alpha2, K);
      for (int k = 0; k < K; k++) {
        const double dab = oracle_digamma(omega[k] + kappa2[k]);
        e1[k] = oracle_digamma(omega[k]) - dab;
        e2[k] = oracle_digamma(kappa2[k]) - dab;
      }
      long double sB = 0;
      for (int i = 0; i < N; i++) sB += dp_expec_log_zi(zeta, N, K, i, e1, e2);
      tB = (double)sB;
      tE = dp_entrop_v(omega, kappa2, K);
    }
    {
      long double sC = 0, sG = 0;
      size_t o = 0;
      for (int j = 0; j < J; j++) {
        long double sCj = 0, sGj = 0;
        for (int k = 0; k < K; k++) {
          sCj += expec_log_ujk(phi + o + (size_t)k * ncat[j], beta, ncat[j]);
          sGj += entrop_dirichlet(phi + o + (size_t)k * ncat[j], ncat[j]);
        }
        sC += (double)sCj;
        sG += (double)sGj;
        o += (size_t)K * ncat[j];
      }
      tC = (double)sC;
      tG = (double)sG;
    }
    tD = (oracle_hook_expec_log_x ? oracle_hook_expec_log_x : oracle_expec_log_x)(
        X, N, phia, cdim, zeta, J, K, n_cat);
    {
      long double sF = 0;
      for (int i = 0; i < N; i++) sF += entrop_zeta_row(zeta, N, K, i);
      tF = (double)sF;
    }
    elbo = tA + tB + tC + tD + tE + tF + tG;

    /* R/mixdir_vi.R:105-109 */
    if (iter != 1 && !isinf(elbo) && elbo - elbo_hist[iter - 2] < -epsilon) warn |= 1;
    if (iter != 1 && !isinf(elbo) && elbo - elbo_hist[iter - 2] < epsilon) converged = 1;
    elbo_hist[iter - 1] = elbo;
    iter++;
  }
  *n_iter = iter - 1;
  *converged_out = converged;

  if (status == 0) {
    /* U (R/mixdir_vi.R:119-130) */
    size_t o = 0;
    for (int j = 0; j < J; j++) {
      for (int k = 0; k < K; k++) {
        const double* p = phi + o + (size_t)k * ncat[j];
        const double sp = sum_ld(p, ncat[j]);
        for (int r = 0; r < ncat[j]; r++) U[o + (size_t)k * ncat[j] + r] = p[r] / sp;
      }
      o += (size_t)K * ncat[j];
    }
    if (!dp) { /* R/mixdir_vi.R:132-133 */
      col_sums(zeta, N, K, omega);
      for (int k = 0; k < K; k++) omega[k] += alpha1;
      const double so = sum_ld(omega, K);
      for (int k = 0; k < K; k++) lambda[k] = omega[k] / so;
      memcpy(prior_out, omega, K * sizeof(double));
    } else { /* R/mixdir_vi_dp.R:148-162 */
      dp_kappa(zeta, N, K, alpha1, alpha2, omega, kappa2);
      long double prod = 1;
      for (int k = 0; k < K; k++) {
        const double nu = omega[k] / (omega[k] + kappa2[k]);
        lambda[k] = nu * (double)prod;
        prod *= (1 - nu);
      }
      if (lambda[K - 1] > 0.01) warn |= 2;
      memcpy(prior_out, omega, K * sizeof(double));
      memcpy(prior_out + K, kappa2, K * sizeof(double));
    }
    if (class_prob) /* NULL: skip the final pass (bench.py times iterations only) */
      oracle_predict(X, N, J, ncat, K, lambda, U, class_prob, pred_class);
    memcpy(phi_out, phi, nphi * sizeof(double));
  }
  *warn_flags = warn;

  free(ord);
  free(e2);
  free(e1);
  free(tmpcol);
  free(cs);
  free(kappa2);
  free(omega);
  free(phia);
  free(phi);
  free(zeta);
  return status;
}

/* ------------------------------------------------------------------------- */
/* fused-state formulation (what one GPU pass must produce, SURVEY.md 8a      */
/* "derived engine contract")                                                 */
/* ------------------------------------------------------------------------- */

/* One E+M pass over rows [row0,row1) of a compact code matrix.
 *   codes    row-major, code_bytes 1 or 2 per cell, 0 = NA, else 1..C_j
 *   logprior[K]  per-class prior term INCLUDING the reference's constants:
 *                psi(omega_k) - psi(sum omega) - 1  (mixdir_impl.cpp:83) or the
 *                DP form (mixdir_impl.cpp:148)
 *   T        ragged like phi: T[j][k][r] = psi(phi_jkr) - psi(sum_r phi_jkr)
 * accumulates (+=) into
 *   S        ragged like phi: sum_i zeta_ik [x_ij == r+1]   (R/mixdir_vi.R:88)
 *   cs[K]    sum_i zeta_ik                                   (R/mixdir_vi.R:61)
 *   *H       sum_i entrop_zeta(zeta_i)                       (R/mixdir_vi.R:181-184)
 *   *underflow  set to 1 if some row has sum_k exp(logit_ik) == 0 (R/mixdir_vi.R:77)
 * zeta_out (nullable): normalised responsibilities, (row1-row0) x K column-major.
 * Plain double accumulation in row order. */
void oracle_estep_stats(const void* codes, int code_bytes, int64_t row0, int64_t row1, int n_quest,
                        const int* ncat, int n_latent, const double* logprior, const double* T,
                        double* S, double* cs, double* H, int* underflow, double* zeta_out) {
  const int J = n_quest, K = n_latent;
  const uint8_t* c8 = (const uint8_t*)codes;
  const uint16_t* c16 = (const uint16_t*)codes;
  size_t* off = (size_t*)malloc((size_t)J * sizeof(size_t));
  for (int j = 0; j < J; j++) off[j] = phi_off(ncat, K, j);
  double* logit = (double*)malloc((size_t)K * sizeof(double));
  double* z = (double*)malloc((size_t)K * sizeof(double));
  const int64_t nr = row1 - row0;
  for (int64_t i = row0; i < row1; i++) {
    for (int k = 0; k < K; k++) {
      double acc = 0;
      for (int j = 0; j < J; j++) {
        unsigned x = code_bytes == 1 ? c8[(size_t)i * J + j] : c16[(size_t)i * J + j];
        if (x) acc += T[off[j] + (size_t)k * ncat[j] + (x - 1)];
      }
      logit[k] = logprior[k] + acc;
    }
    double rs = 0;
    for (int k = 0; k < K; k++) {
      z[k] = exp(logit[k]);
      rs += z[k];
    }
    if (rs == 0) {
      *underflow = 1;
      continue;
    }
    double h = 0;
    for (int k = 0; k < K; k++) {
      z[k] /= rs;
      cs[k] += z[k];
      double zz = z[k] <= DBL_MIN ? DBL_MIN : z[k];
      h += zz * log(zz);
      if (zeta_out) zeta_out[(i - row0) + (size_t)k * nr] = z[k];
    }
    *H -= h;
    for (int j = 0; j < J; j++) {
      unsigned x = code_bytes == 1 ? c8[(size_t)i * J + j] : c16[(size_t)i * J + j];
      if (!x) continue;
      for (int k = 0; k < K; k++) S[off[j] + (size_t)k * ncat[j] + (x - 1)] += z[k];
    }
  }
  free(z);
  free(logit);
  free(off);
}
