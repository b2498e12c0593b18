// k_iter -- everything of one coordinate-ascent iteration that is NOT the pass over the rows, in ONE
// launch (round 1 used k_prior + k_param + k_build_T2 + k_finalize + a host round trip):
//   [DP] class re-ordering by mass          /root/reference/R/mixdir_vi_dp.R:97
//   phi M-step  phi = S + beta              /root/reference/R/mixdir_vi.R:85-91
//   digamma table T (n_cat quirk)           /root/reference/src/mixdir_impl.cpp:61-72,80
//   expec_log_x_cpp collapsed to sum S*T    /root/reference/src/mixdir_impl.cpp:14-38
//   ELBO terms C, D, G per (j,k)            /root/reference/R/mixdir_vi.R:172-174,190-192
//   unit table T2 (feature pairs)           kernels.cuh "Units"
//   ELBO assembly + convergence + warnings  /root/reference/R/mixdir_vi.R:94-109
//   omega / kappa of the NEXT iteration     /root/reference/R/mixdir_vi.R:61, R/mixdir_vi_dp.R:68-71
//
// One block per ELEMENT (a feature, or the two features of a pair: the block that computes T of
// both can build the element's T2 rows without a grid-wide dependency); threads work on (class,
// category) cells, so the serial special-function chains are C_j times shorter than with one
// thread per (feature, class).  The last block to finish (atomic ticket) reduces the per-block
// ELBO pieces in block order, assembles the ELBO, applies the convergence test and computes the
// prior of the next iteration; everything is ordered => bit-reproducible run to run.
//
// The host does not wait for an iteration: it queues several ahead; once `status->stop` is set
// (converged, underflow, invalid codes) every later kernel of the fit returns at once.
#pragma once

#include "kernels.cuh"

namespace mxd {

struct IterArgs {
  int J, K, KP, dp;
  int mode;           // 0: table from the phi already in `phi` + prior from cs_cur (fit start)
                      // 1: full iteration step from the statistics buffer
  int n_cat_bound;    // loop bound of the digamma-sum (mixdir_impl.cpp:65 quirk)
  int NR, NR2;
  int stats_fixed;    // 1: S is int64 fixed point (scale 2^-31, k_em_seg); 0: doubles
  double beta, a1, a2, epsilon;
  double termC_const; // K * sum_j [lgamma(C_j beta) - C_j lgamma(beta)]: expec_log_ujk, R/mixdir_vi.R:172-174
  double termA_const; // dp=0: lgamma(K alpha) - K lgamma(alpha), R/mixdir_vi.R:164-166
  const int* ncat;    // [J]
  const int* rowoff;  // [J]
  const int2* elem;   // [gridDim.x] (feature a, feature b or -1)
  const int2* erange; // [gridDim.x] (first unit-table row, rows) of the element; rows 0: no T2
  const int2* t2src;  // [NR2] feature-table rows a unit row is the sum of (null: identity plan)
  double* stats;      // [S : NR*KP | cs : KP | H | underflow | nNA | bad]   (mode 1; S zeroed after
                      //  use when stats_fixed: the E+M kernel adds into it with atomics)
  double* phi;        // [NR][KP]
  double* T;          // [NR][KP]
  double* T2;         // [NR2][KP]
  double* part;       // [gridDim.x][3] per-block ELBO pieces (C, D, G terms)
  unsigned int* ticket;
  int* underflow_flag;  // the E+M kernels' device flag, cleared for the next iteration
  PriorState* st;
  double* logprior;   // [KP]
  double* cs_cur;     // [K] class masses the NEXT prior is built from (in: colSums(zeta_init))
  double* trace;      // [max_iter]
  IterStatus* status;
  IterStatus* status_host;  // mapped page-locked copy the host polls (nullable)
};

__device__ __forceinline__ double stat_value(const IterArgs& a, size_t i) {
  if (a.stats_fixed && i < (size_t)a.NR * a.KP) return (double)reinterpret_cast<const long long*>(a.stats)[i] * 4.656612873077392578125e-10;
  return a.stats[i];
}

// omega / kappa, prior logits and the ELBO terms A and E from the class masses cs[K].
// Called by all threads of one block; red: >= 6*kMaxClasses doubles of shared memory.
__device__ void prior_step(const IterArgs& a, const double* cs, double* red) {
  const int K = a.K;
  PriorState* st = a.st;
  double *lg1 = red, *lg2 = red + kMaxClasses, *lg12 = red + 2 * kMaxClasses,
         *d1 = red + 3 * kMaxClasses, *d2 = red + 4 * kMaxClasses, *d12 = red + 5 * kMaxClasses;
  if (!a.dp) {
    // omega <- alpha + colSums(zeta)   (R/mixdir_vi.R:61)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const double om = a.a1 + cs[k];
      st->omega[k] = om;
      d1[k] = dev_digamma(om);
      lg1[k] = lgamma(om);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double so = 0;
      for (int k = 0; k < K; k++) so += st->omega[k];
      const double dso = dev_digamma(so);  // mixdir_impl.cpp:59
      double sA = 0, sl = 0, sE = 0;
      for (int k = 0; k < K; k++) {
        const double e = d1[k] - dso;
        st->e1[k] = e;
        a.logprior[k] = e;
        sA += (a.a1 - 1.0) * e;
        sl += lg1[k];
        sE += (st->omega[k] - 1.0) * e;
      }
      for (int k = K; k < a.KP; k++) a.logprior[k] = -INFINITY;
      st->termA = a.termA_const + sA;                         // R/mixdir_vi.R:164-166
      st->termE = -lgamma(so) + sl - sE;                      // R/mixdir_vi.R:177-179
    }
  } else {
    // kappa1 <- a1 + colSums(zeta); kappa2_k <- a2 + sum_{l>k} colSums(zeta)_l  (dp.R:68-71)
    if (threadIdx.x == 0) {
      double tail = 0;
      for (int k = K - 1; k >= 0; k--) {
        st->kappa2[k] = a.a2 + tail;
        tail += cs[k];
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const double k1 = a.a1 + cs[k];
      const double k2 = st->kappa2[k];
      st->omega[k] = k1;
      d1[k] = dev_digamma(k1);
      d2[k] = dev_digamma(k2);
      d12[k] = dev_digamma(k1 + k2);
      lg1[k] = lgamma(k1);
      lg2[k] = lgamma(k2);
      lg12[k] = lgamma(k1 + k2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double skp = 0, sA = 0, sE = 0;
      for (int k = 0; k < K; k++) {
        const double k1 = st->omega[k], k2 = st->kappa2[k];
        const double e1 = d1[k] - d12[k];
        const double e2 = d2[k] - d12[k];
        st->e1[k] = e1;
        st->e2[k] = e2;
        a.logprior[k] = e1 + skp;  // mixdir_impl.cpp:131-138,148
        skp += e2;
        sA += (a.a1 - 1.0) * e1 + (a.a2 - 1.0) * e2;  // dp.R:196-200
        sE += lg1[k] + lg2[k] - lg12[k] - (k1 - 1.0) * d1[k] - (k2 - 1.0) * d2[k] +
              (k1 + k2 - 2.0) * d12[k];  // dp.R:214-221
      }
      for (int k = K; k < a.KP; k++) a.logprior[k] = -INFINITY;
      st->termA = sA;
      st->termE = sE;
    }
  }
  __syncthreads();
}

constexpr int kIterThreads = 256;

__global__ void __launch_bounds__(kIterThreads) k_iter(IterArgs a) {
  __shared__ int perm[kMaxClasses];
  __shared__ double red[6 * kMaxClasses];  // [3][256] reductions; digamma sums; prior scratch
  __shared__ int is_last;
  if (a.mode == 1 && a.status->stop) return;  // uniform over the grid (set by an earlier launch)
  const int K = a.K, KP = a.KP, nt = blockDim.x, tid = threadIdx.x;
  const size_t tab = (size_t)a.NR * KP;

  // zeta <- zeta[, order(-colSums(zeta))]  (R/mixdir_vi_dp.R:97): stable, descending; applied
  // to the REDUCED statistics
  // (identity first: NaN class masses -- an underflowing pass -- have no order and must not leave
  // perm[] entries unset)
  for (int k = tid; k < K; k += nt) perm[k] = k;
  __syncthreads();
  for (int k = tid; k < K; k += nt) {
    if (a.dp && a.mode == 1) {
      const double c = stat_value(a, tab + k);
      int rank = 0;
      for (int l = 0; l < K; l++) {
        const double cl = stat_value(a, tab + l);
        rank += (cl > c) || (cl == c && l < k);
      }
      perm[rank] = k;
    }
  }
  __syncthreads();

  const int2 el = a.elem[blockIdx.x];
  double p0 = 0, p1 = 0, p2 = 0;  // this thread's share of the C, D, G terms
  double* ds_bound = red;                   // [K]
  double* ds_all = red + kMaxClasses;       // [K]
  for (int f = 0; f < 2; f++) {
    const int j = f == 0 ? el.x : el.y;
    if (j < 0) break;
    const int C = a.ncat[j];
    const size_t g0 = (size_t)a.rowoff[j];
    const int cells = K * C;
    // phi[[j]][[k]][r] <- sum(zeta[,k] * (X[,j]==r), na.rm=TRUE) + beta  (R/mixdir_vi.R:88)
    if (a.mode == 1)
      for (int i = tid; i < cells; i += nt) {
        const int r = i / K, k = i - r * K;
        a.phi[(g0 + 1 + r) * KP + k] = stat_value(a, (g0 + 1 + r) * KP + perm[k]) + a.beta;
      }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
      double sum_all = 0, sum_bound = 0;
      for (int r = 0; r < C; r++) {
        const double ph = a.phi[(g0 + 1 + r) * KP + k];
        sum_all += ph;
        if (r < a.n_cat_bound) sum_bound += ph;
      }
      ds_bound[k] = dev_digamma(sum_bound);  // mixdir_impl.cpp:61-72
      ds_all[k] = dev_digamma(sum_all);      // R-level digamma(sum(phi_jk))
      a.T[g0 * KP + k] = 0.0;                // NA row
      if (a.mode == 1) p2 -= lgamma(sum_all);  // entrop_phi, R/mixdir_vi.R:190-192
    }
    __syncthreads();
    for (int i = tid; i < cells; i += nt) {
      const int r = i / K, k = i - r * K;
      const size_t g = g0 + 1 + r;
      const double ph = a.phi[g * KP + k];
      const double dg = dev_digamma(ph);
      const double t = dg - ds_bound[k];  // mixdir_impl.cpp:80
      a.T[g * KP + k] = t;
      if (a.mode == 1) {
        const double da = dg - ds_all[k];
        p0 += (a.beta - 1.0) * da;
        p1 += stat_value(a, g * KP + perm[k]) * t;  // S * T: expec_log_x_cpp collapsed
        p2 += lgamma(ph) - (ph - 1.0) * da;
      }
    }
    __syncthreads();
  }
  // unit table rows of this element: T2[g2][k] = T[src.x][k] + T[src.y][k]
  const int2 er = a.erange[blockIdx.x];
  if (a.t2src && er.y > 0) {
    for (int i = tid; i < er.y * KP; i += nt) {
      const int g2 = er.x + i / KP, k = i % KP;
      const int2 s = __ldg(&a.t2src[g2]);
      double v = k < K ? a.T[(size_t)s.x * KP + k] : 0.0;
      if (s.y >= 0 && k < K) v += a.T[(size_t)s.y * KP + k];
      a.T2[(size_t)g2 * KP + k] = v;
    }
  }
  // block reduction of the pieces, fixed order
  if (a.mode == 1) {
    __syncthreads();
    red[tid] = p0;
    red[kIterThreads + tid] = p1;
    red[2 * kIterThreads + tid] = p2;
    __syncthreads();
    for (int o = nt >> 1; o > 0; o >>= 1) {
      if (tid < o) {
        red[tid] += red[tid + o];
        red[kIterThreads + tid] += red[kIterThreads + tid + o];
        red[2 * kIterThreads + tid] += red[2 * kIterThreads + tid + o];
      }
      __syncthreads();
    }
    if (tid == 0) {
      a.part[(size_t)blockIdx.x * 3 + 0] = red[0];
      a.part[(size_t)blockIdx.x * 3 + 1] = red[kIterThreads];
      a.part[(size_t)blockIdx.x * 3 + 2] = red[2 * kIterThreads];
    }
  }
  // ---- last block: ELBO, convergence, next prior ----
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (a.mode == 1) {
    const int nb = gridDim.x;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int b = tid; b < nb; b += nt) {
      s0 += __ldcg(&a.part[(size_t)b * 3 + 0]);
      s1 += __ldcg(&a.part[(size_t)b * 3 + 1]);
      s2 += __ldcg(&a.part[(size_t)b * 3 + 2]);
    }
    __syncthreads();
    red[tid] = s0;
    red[kIterThreads + tid] = s1;
    red[2 * kIterThreads + tid] = s2;
    __syncthreads();
    for (int o = nt >> 1; o > 0; o >>= 1) {
      if (tid < o) {
        red[tid] += red[tid + o];
        red[kIterThreads + tid] += red[kIterThreads + tid + o];
        red[2 * kIterThreads + tid] += red[2 * kIterThreads + tid + o];
      }
      __syncthreads();
    }
    const double tC = red[0] + a.termC_const, tG = red[2 * kIterThreads];
    const double tD = red[kIterThreads] + (double)K * a.stats[tab + KP + 2];  // +1 per NA cell per class
    __syncthreads();
    // class masses of the new zeta, permuted: what the next prior is built from
    for (int k = tid; k < K; k += nt) red[k] = stat_value(a, tab + perm[k]);
    __syncthreads();
    if (tid == 0) {
      const PriorState* st = a.st;  // still the prior this iteration's E-step used (old omega/kappa)
      double tB = 0;
      if (!a.dp) {
        for (int k = 0; k < K; k++) tB += red[k] * st->e1[k];  // expec_log_zi, R/mixdir_vi.R:168-170
      } else {
        double tail = 0;  // sum_{l>k} of the permuted class masses
        for (int k = K - 1; k >= 0; k--) {
          tB += tail * st->e2[k] + red[k] * st->e1[k];  // dp_expec_log_zi, dp.R:202-211
          tail += red[k];
        }
      }
      const double H = a.stats[tab + KP];
      const double elbo = st->termA + tB + tC + tD + st->termE + H + tG;
      IterStatus* s = a.status;
      const int it = s->iter;
      int conv = 0, warn = s->warn;
      if (it != 0 && !isinf(elbo)) {
        const double diff = elbo - a.trace[it - 1];
        if (diff < -a.epsilon) warn |= 1;  // R/mixdir_vi.R:105-108
        if (diff < a.epsilon) conv = 1;    // R/mixdir_vi.R:109
      }
      a.trace[it] = elbo;
      s->elbo = elbo;
      s->converged = conv;
      s->warn = warn;
      s->underflow = (a.stats[tab + KP + 1] > 0.0) ? 1 : 0;
      s->bad = (a.stats[tab + KP + 3] > 0.0) ? 1 : 0;
      s->iter = it + 1;
      s->stop = (conv || s->underflow || s->bad) ? 1 : 0;
      *a.underflow_flag = 0;
    }
    for (int k = tid; k < K; k += nt) a.cs_cur[k] = red[k];
    __syncthreads();
    if (a.stats_fixed) {  // the next E+M pass adds into S with atomics
      long long* z = reinterpret_cast<long long*>(a.stats);
      for (size_t i = tid; i < tab; i += nt) z[i] = 0;
    }
  }
  prior_step(a, a.cs_cur, red);
  if (tid == 0) {
    *a.ticket = 0u;
    if (a.status_host && a.mode == 1) {
      *a.status_host = *a.status;
      __threadfence_system();
    }
  }
}

}  // namespace mxd
