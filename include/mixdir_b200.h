/* mixdir_b200 -- C ABI of the B200 engine for mixdir's variational-inference hot path.
 *
 * Plain pointers and sizes only; no R, Rcpp or torch types.  Everything an R
 * (Rcpp), Python (ctypes) or C host needs to run the coordinate-ascent loop of
 *   mixdir_vi()      /root/reference/R/mixdir_vi.R:2-156      (fixed K, Dirichlet prior)
 *   mixdir_vi_dp()   /root/reference/R/mixdir_vi_dp.R:2-187   (Dirichlet-process prior)
 * on one or more B200s.  The boundary sits ONE LEVEL ABOVE the reference's
 * three .Call entry points (`_mixdir_update_zeta_cpp`, `_mixdir_update_zeta_dp_cpp`,
 * `_mixdir_expec_log_x_cpp`, /root/reference/src/RcppExports.cpp:9-60), because
 * those move the full N x K zeta and N x J X through host memory twice per
 * iteration.  The reference-side binding (one new
 * `// [[Rcpp::export]]` function + an `engine=` selector in mixdir()) is shown
 * in INTEGRATION.md.
 *
 * Ownership: the caller allocates every output array; the library owns only
 * device memory, streams and NCCL communicators inside the handle.
 * Errors: every function returns a status code and never throws or aborts;
 * mixdir_b200_last_error() returns the message of the last failure on the
 * calling thread.
 * Threading: calls on one handle must be serialised by the caller (R is single
 * threaded); different handles may be used from different threads.
 */
#ifndef MIXDIR_B200_H
#define MIXDIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIXDIR_B200_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------- */
#define MIXDIR_B200_OK 0
/* some row had sum_k zeta_ik == 0: the adapter raises the reference's stop()
 * text of R/mixdir_vi.R:77-81 (R/mixdir_vi_dp.R:91-95) */
#define MIXDIR_B200_UNDERFLOW 1
#define MIXDIR_B200_CUDA_ERROR 2 /* CUDA / NCCL failure, see last_error() */
#define MIXDIR_B200_BAD_ARGUMENT 3

/* ---- warn_flags bits (adapter turns them into R warning()s) -------------- */
/* "The ELBO decreased ..." R/mixdir_vi.R:105-108, R/mixdir_vi_dp.R:120-122 */
#define MIXDIR_B200_WARN_ELBO_DECREASED 1
/* "considerable weight on the last component" R/mixdir_vi_dp.R:159-162 */
#define MIXDIR_B200_WARN_DP_LAST_COMPONENT 2

/* ---- data layouts -------------------------------------------------------- *
 * codes      category codes as produced by mixdir() (R/mixdir.R:97-118):
 *            1..C_j, with 0 standing for NA.  Row-major [n_rows][n_features],
 *            code_bytes = 1 (all C_j <= 255) or 2.
 * ncat[j]    C_j = length(categories[[j]])   (R/mixdir.R:109)
 * "ragged"   the R list phi[[j]][[k]][r] (R/mixdir_vi.R:34-39) flattened in
 *            (j, k, r) order: element (j,k,r) at  K*sum_{j'<j} C_j' + k*C_j + r.
 *            Used for phi_init, phi, category_prob (U) alike.
 * class_prob n_rows x K column-major doubles, exactly R's matrix layout.
 * pred_class 1-based, first maximum (R which.max, R/mixdir_vi.R:148); 0 marks a
 *            row whose class_prob is all-NaN.
 */

typedef struct mixdir_b200_handle mixdir_b200_handle;

typedef struct mixdir_b200_fit_params {
  int n_latent;     /* K  (mixdir(n_latent=), R/mixdir.R:80) */
  int dp;           /* 0: mixdir_vi, 1: mixdir_vi_dp (select_latent, R/mixdir.R:122-128) */
  double alpha1;    /* alpha (dp=0) or alpha[1] (dp=1); defaults 1 (R/mixdir_vi.R:10-17) */
  double alpha2;    /* alpha[2] (dp=1 only); default 1 (R/mixdir_vi_dp.R:10-20) */
  double beta;      /* default 0.1 (R/mixdir_vi.R:19-26) */
  int max_iter;     /* R/mixdir.R:84 */
  double epsilon;   /* absolute ELBO tolerance, R/mixdir.R:85; -INFINITY = run max_iter */
  int n_cat_bound;  /* n_cat = max(X, na.rm=TRUE) (R/mixdir_vi.R:8): the upper bound of the
                       digamma-sum loop in src/mixdir_impl.cpp:65.  <= 0: use the largest
                       code present in the data (what the reference computes). */
} mixdir_b200_fit_params;

/* Number of doubles in a ragged (j,k,r) array: K * sum_j ncat[j]. */
size_t mixdir_b200_ragged_size(int n_features, const int* ncat, int n_latent);

/* ---- handle life cycle ---------------------------------------------------
 * Replaces: the `X` argument of mixdir_vi()/mixdir_vi_dp() (R/mixdir_vi.R:2,
 * R/mixdir.R:117-118).  Uploads the code matrix ONCE; the handle is then reused
 * across `repetitions` (R/mixdir.R:120-134) and predict calls.
 *
 * n_devices/device_ids: single-process multi-GPU (what an R session needs: R is
 * one process).  Rows are split into contiguous blocks, one per device; each
 * iteration ends with one NCCL all-reduce of the statistics buffer.
 * n_devices = 0 or 1 with device_ids == NULL uses the current CUDA device.
 *
 * The upload is ASYNCHRONOUS: create returns once the chunked host-to-device copies are queued,
 * and the first mixdir_b200_fit iteration consumes the chunks as they arrive (copy and compute
 * overlap).  With pageable host memory the runtime stages the data before returning; with
 * page-locked memory the caller must leave `codes` untouched until the first fit / predict /
 * getter call on the handle has returned.  Code validity (code <= ncat[j]) is therefore
 * reported by that first call (MIXDIR_B200_BAD_ARGUMENT), not by create. */
int mixdir_b200_create(const void* codes, int code_bytes, int64_t n_rows, int n_features,
                       const int* ncat, int n_devices, const int* device_ids,
                       mixdir_b200_handle** out);

/* Same from the reference's own representation of X: an n_rows x n_features
 * column-major double matrix, 1..C_j with NaN/NA for missing
 * (R/mixdir.R:117-118).  Packing to the compact code matrix happens inside: a few host threads
 * (MIXDIR_B200_HOST_THREADS, default min(16, cores / LOCAL_WORLD_SIZE); AVX2 when the CPU has it)
 * write 1- or 2-byte codes into page-locked
 * staging buffers, chunk by chunk, and every chunk is copied to the device while the next one is
 * packed.  X may be pageable and is no longer needed when the call returns. */
int mixdir_b200_create_from_r_matrix(const double* X, int64_t n_rows, int n_features,
                                     const int* ncat, int n_devices, const int* device_ids,
                                     mixdir_b200_handle** out);

/* Multi-process flavour (one process per GPU, e.g. torchrun): every rank passes its own
 * contiguous row shard.  The NCCL communicator lives in a separate object so that it is set up
 * once and shared by any number of handles (repetitions, predict, re-uploads):
 * unique_id is the 128-byte NCCL id produced by mixdir_b200_nccl_unique_id() on rank 0 and
 * broadcast by the host program; comm_init is collective over all ranks, and so are
 * create_dist / create_dist_from_r_matrix (the ranks exchange their row counts).  The communicator
 * must outlive the handles created on it. */
typedef struct mixdir_b200_comm mixdir_b200_comm;
int mixdir_b200_nccl_unique_id(void* out128);
int mixdir_b200_comm_init(int device, int rank, int world, const void* unique_id128,
                          mixdir_b200_comm** out);
int mixdir_b200_comm_destroy(mixdir_b200_comm* comm);
int mixdir_b200_create_dist(const void* codes_local, int code_bytes, int64_t n_rows_local,
                            int n_features, const int* ncat, mixdir_b200_comm* comm,
                            mixdir_b200_handle** out);
/* The same from this rank's rows of the reference's fp64 matrix (column-major, leading dimension
 * n_rows_local; 1..C_j, NaN/NA missing), as mixdir_b200_create_from_r_matrix. */
int mixdir_b200_create_dist_from_r_matrix(const double* X_local, int64_t n_rows_local, int n_features,
                                          const int* ncat, mixdir_b200_comm* comm,
                                          mixdir_b200_handle** out);

int mixdir_b200_destroy(mixdir_b200_handle* h);

/* Device and page-locked buffers of destroyed handles are kept in a process-wide cache and reused
 * by later handles of the same shapes (allocation and page-locking cost milliseconds).  This
 * call returns all cached blocks to the driver. */
int mixdir_b200_release_cache(void);

/* rows held by this handle (local shard rows in the multi-process flavour) */
int64_t mixdir_b200_n_rows(const mixdir_b200_handle* h);
/* largest code present = the reference's n_cat (R/mixdir_vi.R:8); global over ranks */
int mixdir_b200_max_code(const mixdir_b200_handle* h);
/* number of NA cells, global over ranks */
int64_t mixdir_b200_n_missing(const mixdir_b200_handle* h);

/* ---- the fit --------------------------------------------------------------
 * Replaces: the body of mixdir_vi() (R/mixdir_vi.R:50-155) / mixdir_vi_dp()
 * (R/mixdir_vi_dp.R:56-185) given explicit initial values.
 *
 * colsum_init[K]  colSums(zeta_init): the only way zeta_init enters the fit
 *                 (R/mixdir_vi.R:61 runs before zeta is overwritten at :76).
 *                 For dp=1 kappa2 is derived from the same column sums
 *                 (R/mixdir_vi_dp.R:68-71).
 * phi_init        ragged, values as in R/mixdir_vi.R:34-39.
 * outputs:
 *   elbo_trace[max_iter]  `convergence` (first *n_iter entries are valid)
 *   n_iter, converged, elbo   (`converged`, `ELBO`)
 *   lambda[K]             `lambda`  (dp=1: stick-breaking weights, not normalised)
 *   prior_param           dp=0: omega[K];  dp=1: kappa1[K] then kappa2[K]
 *   phi, category_prob    ragged `specific_params$phi`, `category_prob`
 *   class_prob (nullable) n_rows x K column-major `class_prob`
 *   pred_class (nullable) n_rows `pred_class`
 *   warn_flags            MIXDIR_B200_WARN_* bits
 * In the multi-process flavour class_prob / pred_class cover the local rows;
 * everything else is identical on all ranks.
 * Returns MIXDIR_B200_OK, or MIXDIR_B200_UNDERFLOW (outputs other than
 * elbo_trace/n_iter are then unspecified). */
int mixdir_b200_fit(mixdir_b200_handle* h, const mixdir_b200_fit_params* params,
                    const double* colsum_init, const double* phi_init, double* elbo_trace,
                    int* n_iter, int* converged, double* elbo, double* lambda,
                    double* prior_param, double* phi, double* category_prob,
                    double* class_prob, int32_t* pred_class, int* warn_flags);

/* ---- repetitions: several fits of one batch ----------------------------------
 * Replaces: the `repetitions` loop of mixdir() (R/mixdir.R:120-134): n_fits independent restarts
 * on the same data and hyper-parameters, each from its own initial values, the result with the
 * largest ELBO kept (`if(result$ELBO < result_tmp$ELBO) result <- result_tmp`: the FIRST maximum).
 * Small fits are bound by launch latency, not by the GPU, so the library runs up to 8 of them
 * concurrently (one CUDA stream, workspace and host thread each, all reading the ONE uploaded code
 * matrix; MIXDIR_B200_BATCH_STREAMS overrides the number); a fit whose row pass fills the GPU
 * (> 2M rows), and any multi-GPU handle, runs its restarts one after the other.  Every fit is
 * bit-identical to the same mixdir_b200_fit call made alone.
 *   colsum_init [n_fits][K], phi_init [n_fits][ragged]            inputs, fit after fit
 *   elbo_trace  [n_fits][max(max_iter,1)] (nullable), n_iter / converged / warn_flags / status
 *   [n_fits] (nullable), elbo [n_fits], lambda [n_fits][K], prior_param [n_fits][K or 2K] (nullable),
 *   phi (nullable) / category_prob [n_fits][ragged]               outputs, fit after fit
 *   best        index of the kept fit (-1 on error)
 *   class_prob, pred_class (nullable)  of the kept fit only, computed from its lambda and
 *               category_prob as predict_class does (R/predict.R:88-103 == R/mixdir_vi.R:136-148)
 * Returns the status of the first failed fit (an error in any repetition ends mixdir()), else OK. */
int mixdir_b200_fit_batch(mixdir_b200_handle* h, const mixdir_b200_fit_params* params, int n_fits,
                          const double* colsum_init, const double* phi_init, double* elbo_trace,
                          int* n_iter, int* converged, double* elbo, double* lambda,
                          double* prior_param, double* phi, double* category_prob, int* warn_flags,
                          int* status, int* best, double* class_prob, int32_t* pred_class);

/* ---- predict --------------------------------------------------------------
 * Replaces: the arithmetic of predict_class() (R/predict.R:88-103), identical to
 * the final pass of the drivers (R/mixdir_vi.R:136-148).  The handle holds the
 * (new) data, coded against the TRAINING categories; unseen levels / missing
 * columns are NA = 0 (the adapter keeps the name matching and the warning of
 * R/predict.R:64-84). */
int mixdir_b200_predict(mixdir_b200_handle* h, int n_latent, const double* lambda,
                        const double* category_prob, double* class_prob,
                        int32_t* pred_class);

/* ---- unit-test door at the level of the reference's native functions -------
 * One E+M pass with caller-supplied tables; mirrors what update_zeta_cpp /
 * update_zeta_dp_cpp (src/mixdir_impl.cpp:57-87,113-152) followed by the
 * normalisation and the M-step (R/mixdir_vi.R:77-91) produce.
 *   log_prior[K]   psi(omega_k)-psi(sum omega)   (dp: the kappa form, impl.cpp:148);
 *                  the reference's "-1" is applied inside.
 *   table          ragged T[j][k][r] = psi(phi_jkr) - psi(sum_r phi_jkr)
 *   S (ragged), colsum[K], entropy, underflow: the statistics buffer
 *   zeta (nullable) n_rows x K column-major normalised responsibilities
 *   engine         0 = automatic, 1 = generic kernel, 2 = fused kernel, 4 = unit kernel (1-byte
 *                  codes), 5 = sorted-segment kernel (1-byte codes; S is then an exact sum of 32-bit
 *                  fixed-point responsibilities, units of 1 / (2^32 - 1)) */
int mixdir_b200_estep(mixdir_b200_handle* h, int n_latent, const double* log_prior,
                      const double* table, int engine, double* S, double* colsum,
                      double* entropy, int* underflow, double* zeta);

/* ---- introspection for benchmarks ------------------------------------------ */
typedef struct mixdir_b200_timing {
  double fit_loop_ms;   /* CUDA-event time of the iteration loop of the last fit (device 0) */
  double em_kernel_ms;  /* summed CUDA-event time of the E+M kernel launches in that loop (0 for handles
                           of fewer than 65,536 rows unless MIXDIR_B200_TIME_EM is set: the two event
                           records per iteration cost a small fit several per cent) */
  int iterations;       /* iterations executed */
  int em_launches;      /* E+M kernel launches in the loop */
  int total_launches;   /* all kernel launches of the library inside the loop (per device) */
  int engine;           /* 1 generic, 2 fused (k_em_fused), 4 unit kernel (k_em_units),
                           5 sorted-segment kernel (k_em_seg) */
  int classes_per_cta;  /* fused: KC */
  int cluster_size;     /* fused: P */
  int grid_ctas;        /* fused: CTAs launched */
  int smem_bytes;       /* fused: dynamic shared memory per CTA */
  int unit_elements;    /* fused: gather/scatter elements per row (features, pairs, padding) */
  int unit_pairs;       /* fused: features combined two-by-two into one element */
  int unit_table_rows;  /* fused: rows of the unit tables (sum over elements of their codes) */
  int final_pass_engine;        /* last final pass / predict: 6 unit kernel (k_predict_units), 7 feature
                                   table in shared memory (k_predict_smem), 8 general kernel; 0 none */
  double final_pass_kernel_ms;  /* its CUDA-event time on device 0: table build + the pass, no copies */
} mixdir_b200_timing;

/* Inner loop of find_defining_features() (/root/reference/R/predict.R:266-339; the `vapply` over the
 * remaining variables at :299-310).  With p = predict(all features of the handle's rows) and, for every
 * feature i with active[i] != 0, q_i = predict(active features without i) -- both in the arithmetic of
 * predict_class (R/predict.R:88-103; dropped features and NA cells contribute 1) --
 *   measure 0 ("JS"):  out[i] = sum_{rows,k} (q (log q - log p) + p (log p - log q)) / 2    (:304)
 *   measure 1 ("ARI"): out[i] = 1 - adjusted Rand index(which.max q, which.max p)           (:306;
 *                      the reference calls mcclust::arandi, not vendored: standard Hubert-Arabie form)
 * out[n_features] is the same quantity for the active set itself (nothing else removed; the final
 * quality of :326-331); out[i] = NaN for inactive i.  `rows` (optional): 0-based row subsample of the
 * handle's rows (R: X[subsample, ]), NULL = every row.  One pass over the codes serves all candidates:
 * the reference makes |active| predict() calls per round.
 *   lambda         [K]    category_prob  ragged (j,k,r) as returned by fit      out  [n_features + 1] */
int mixdir_b200_feature_divergence(mixdir_b200_handle* h, int n_latent, const double* lambda,
                                   const double* category_prob, const unsigned char* active,
                                   int measure, const int64_t* rows, int64_t n_rows_sub, double* out);

int mixdir_b200_last_timing(const mixdir_b200_handle* h, mixdir_b200_timing* out);
/* 0 = automatic (default: 1-byte codes: the sorted-segment kernel from 16384 rows per shard on, else
 * the unit kernel; 2-byte codes: the fused kernel; the generic kernel where none of them fits).
 * The sorted-segment kernel rounds the responsibilities that feed phi to 32 bits (unbiased; sums exact):
 * ELBO traces typically within 1e-10 of fp64 statistics, transiently more while classes merge
 * (DESIGN.md section 5); engine 4 keeps fp64 statistics at ~20 % more time per iteration.
 * 1 = generic kernel only, 2 = fused kernel or fail, 4 = unit kernel or fail,
 * 5 = sorted-segment kernel or fail */
int mixdir_b200_set_engine(mixdir_b200_handle* h, int engine);
/* Feature pairing of the fused and unit kernels (default on): two features share one table gather and one
 * histogram update where shared memory allows.  Results change only by floating-point
 * re-association inside the per-row logit sum (mixdir_impl.cpp:77-82).  0 switches it off. */
int mixdir_b200_set_pairing(mixdir_b200_handle* h, int on);
/* Diagnostic, host only (no device): the unit plan with n_pairs feature pairs for 1-byte codes.
 * has_na (nullable; the unit kernel passes the upload scan's result): has_na[j] == 0 = feature j holds
 * no NA anywhere and gets no NA row (its unit-local code is x - 1); NULL = every feature keeps its NA row.
 * desc: n_elements x {byte offset of feature a (<0: padding), of feature b (<0: single), unit-local
 * codes of b, bit 0/1 = a/b without NA row};
 * gbase: first unit-table row of each element; t2src: n_table_rows x {feature-table row a, row b or
 * -1} (only written when n_pairs > 0); marg: sum_j(C_j+1) x {start, stride, count, 0}: the unit rows a
 * feature-table row collects.  Output arrays may be NULL. */
int mixdir_b200_unit_plan(int n_features, const int* ncat, int n_pairs, const unsigned char* has_na,
                          int* n_elements, int* n_table_rows, int* desc, int* gbase, int* t2src,
                          int* marg);
/* Diagnostic, host only: the unit kernel's shared-memory placement of that plan for
 * classes_per_cta (KC) classes per CTA.  Q = max(1, 16/KC) table rows of different elements share a
 * 128-byte shared-memory row; element e sits at position e mod Q.  slot: n_table_rows entries, shared-
 * memory slot (row * Q + position) of every unit-table row (row 0 of each position is the zero /
 * scratch row of the padding elements); eoff_rot: 4 x n_elements byte offsets of element
 * (e & ~3) | ((e + r) & 3) for rotation r.  desc / gbase are in the (re-ordered) element order. */
int mixdir_b200_unit_layout(int n_features, const int* ncat, int n_pairs, const unsigned char* has_na,
                            int classes_per_cta, int* n_elements, int* n_table_rows, int* smem_rows,
                            int* desc, int* gbase, int* slot, int* eoff_rot);

const char* mixdir_b200_last_error(void);
int mixdir_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MIXDIR_B200_H */
