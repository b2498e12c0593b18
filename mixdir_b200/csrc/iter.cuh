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
  int stats_fixed;    // 1 (k_em_seg): S is int64 fixed point (units of 1 / (2^32 - 1)), the class masses
                      // two int64 words each (seg.cuh: cs_value), the tail follows them; 0: doubles
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
  if (a.stats_fixed && i < (size_t)a.NR * a.KP) return (double)reinterpret_cast<const long long*>(a.stats)[i] * kFixedInv;
  return a.stats[i];
}
// class mass of class k: doubles behind the table, or k_em_seg's two integer words (seg.cuh: cs_value)
__device__ __forceinline__ double mass_value(const IterArgs& a, int k) {
  const size_t tab = (size_t)a.NR * a.KP;
  if (!a.stats_fixed) return a.stats[tab + k];
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a.stats) + tab;
  return ((double)w[k] * 4294967296.0 + (double)w[a.KP + k]) * (1.0 / 4611686018427387904.0);
}

// omega / kappa, prior logits and the ELBO terms A and E from the class masses cs[K].
// Called by all threads of one block; red: >= 6*kMaxClasses doubles of shared memory.
// The special functions of a class (digamma / lgamma of omega, or of kappa1, kappa2 and their sum)
// are independent: they are dealt to different threads, so their serial latencies overlap.
__device__ void prior_step(const IterArgs& a, const double* cs, double* red) {
  const int K = a.K, nt = blockDim.x, tid = threadIdx.x;
  PriorState* st = a.st;
  double *lg1 = red, *lg2 = red + kMaxClasses, *lg12 = red + 2 * kMaxClasses,
         *d1 = red + 3 * kMaxClasses, *d2 = red + 4 * kMaxClasses, *d12 = red + 5 * kMaxClasses;
  __shared__ double om_s[kMaxClasses], k2_s[kMaxClasses], tot_s[2];
  if (!a.dp) {
    // omega <- alpha + colSums(zeta)   (R/mixdir_vi.R:61)
    for (int k = tid; k < K; k += nt) om_s[k] = a.a1 + cs[k];
    __syncthreads();
    for (int i = tid; i < 2 * K + 2; i += nt) {
      if (i < K) {
        d1[i] = dev_digamma(om_s[i]);
      } else if (i < 2 * K) {
        lg1[i - K] = lgamma(om_s[i - K]);
      } else {  // digamma / lgamma of sum(omega): mixdir_impl.cpp:59, R/mixdir_vi.R:177-179
        double so = 0;
        for (int k = 0; k < K; k++) so += om_s[k];
        tot_s[i - 2 * K] = (i == 2 * K) ? dev_digamma(so) : lgamma(so);
      }
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
      const double e = d1[k] - tot_s[0];
      st->omega[k] = om_s[k];
      st->e1[k] = e;
      a.logprior[k] = e;
      d2[k] = e;
    }
    for (int k = K + tid; k < a.KP; k += nt) a.logprior[k] = -INFINITY;
    __syncthreads();
    if (tid == 0) {
      double sA = 0, sl = 0, sE = 0;
      for (int k = 0; k < K; k++) {
        sA += (a.a1 - 1.0) * d2[k];
        sl += lg1[k];
        sE += (om_s[k] - 1.0) * d2[k];
      }
      st->termA = a.termA_const + sA;   // R/mixdir_vi.R:164-166
      st->termE = -tot_s[1] + sl - sE;  // R/mixdir_vi.R:177-179
    }
  } else {
    // kappa1 <- a1 + colSums(zeta); kappa2_k <- a2 + sum_{l>k} colSums(zeta)_l  (dp.R:68-71)
    if (tid == 0) {
      double tail = 0;
      for (int k = K - 1; k >= 0; k--) {
        k2_s[k] = a.a2 + tail;
        tail += cs[k];
      }
    }
    for (int k = tid; k < K; k += nt) om_s[k] = a.a1 + cs[k];
    __syncthreads();
    for (int i = tid; i < 6 * K; i += nt) {
      const int k = i % K, what = i / K;
      const double k1 = om_s[k], k2 = k2_s[k];
      if (what == 0) d1[k] = dev_digamma(k1);
      else if (what == 1) d2[k] = dev_digamma(k2);
      else if (what == 2) d12[k] = dev_digamma(k1 + k2);
      else if (what == 3) lg1[k] = lgamma(k1);
      else if (what == 4) lg2[k] = lgamma(k2);
      else lg12[k] = lgamma(k1 + k2);
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
      st->omega[k] = om_s[k];
      st->kappa2[k] = k2_s[k];
    }
    if (tid == 0) {
      double skp = 0, sA = 0, sE = 0;
      for (int k = 0; k < K; k++) {
        const double k1 = om_s[k], k2 = k2_s[k];
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
constexpr int kIterCells = 1792;  // (class, category) cells of an element kept in shared memory

__global__ void __launch_bounds__(kIterThreads) k_iter(IterArgs a) {
  __shared__ int perm[kMaxClasses];
  __shared__ double red[6 * kMaxClasses];  // [3][256] reductions; digamma sums; prior scratch
  __shared__ double ph_s[kIterCells];      // phi of the element's cells, then their table values T
  __shared__ double dg_s[kIterCells];      // digamma(phi)
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
  if (a.dp && a.mode == 1) {
    for (int k = tid; k < K; k += nt) red[k] = mass_value(a, k);
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
      const double c = red[k];
      int rank = 0;
      for (int l = 0; l < K; l++) {
        const double cl = red[l];
        rank += (cl > c) || (cl == c && l < k);
      }
      perm[rank] = k;
    }
    __syncthreads();
  }

  // The element of this block: one feature, or the two features of a pair.  Its (class, category)
  // cells are numbered i = r * K + k, feature a first.  Everything that needs a special function
  // is an independent work item dealt to the threads (their latencies, ~1 us each at this
  // occupancy, were paid one after the other before):
  //   item [0, cells)          digamma(phi_cell)
  //   item [cells, 2 cells)    lgamma(phi_cell)                 (entrop_phi, mode 1)
  //   then 3 per (feature, k): digamma(sum_{r<n_cat} phi), digamma(sum_r phi), lgamma(sum_r phi)
  const int2 el = a.elem[blockIdx.x];
  const int ja = el.x, jb = el.y;
  const int Ca = a.ncat[ja], Cb = jb >= 0 ? a.ncat[jb] : 0;
  const size_t g0a = (size_t)a.rowoff[ja], g0b = jb >= 0 ? (size_t)a.rowoff[jb] : 0;
  const int cellsA = K * Ca, cells = cellsA + K * Cb;
  const bool in_smem = cells <= kIterCells;
  const int nf = jb >= 0 ? 2 : 1;
  double* dsb = red;                    // [2][K] digamma of the bounded sum (mixdir_impl.cpp:61-72)
  double* dsa = red + 2 * kMaxClasses;  // [2][K] digamma of the full sum
  double p0 = 0, p1 = 0, p2 = 0;        // this thread's share of the C, D, G terms

  // phi[[j]][[k]][r] <- sum(zeta[,k] * (X[,j]==r), na.rm=TRUE) + beta  (R/mixdir_vi.R:88)
  for (int i = tid; i < cells; i += nt) {
    const bool fa = i < cellsA;
    const int ii = fa ? i : i - cellsA;
    const int r = ii / K, k = ii - r * K;
    const size_t g = (fa ? g0a : g0b) + 1 + r;
    double ph;
    if (a.mode == 1) {
      ph = stat_value(a, g * KP + perm[k]) + a.beta;
      a.phi[g * KP + k] = ph;
    } else {
      ph = a.phi[g * KP + k];
    }
    if (in_smem) ph_s[i] = ph;
  }
  __syncthreads();
  const int items = 2 * cells + 3 * nf * K;
  for (int v = tid; v < items; v += nt) {
    if (v < 2 * cells) {
      const int i = v < cells ? v : v - cells;
      double ph;
      if (in_smem) {
        ph = ph_s[i];
      } else {
        const bool fa = i < cellsA;
        const int ii = fa ? i : i - cellsA;
        const int r = ii / K, k = ii - r * K;
        ph = a.phi[((fa ? g0a : g0b) + 1 + r) * KP + k];
      }
      if (v < cells) {
        if (in_smem) dg_s[i] = dev_digamma(ph);  // (elements too large for shared memory compute it below)
      } else if (a.mode == 1) {
        p2 += lgamma(ph);  // entrop_phi, R/mixdir_vi.R:190-192
      }
    } else {
      const int w = v - 2 * cells;
      const int what = w / (nf * K), fk = w - what * (nf * K);
      const int f = fk / K, k = fk - f * K;
      const int C = f == 0 ? Ca : Cb;
      const size_t g0 = f == 0 ? g0a : g0b;
      const int base = f == 0 ? 0 : cellsA;
      const int bound = what == 0 ? min(C, a.n_cat_bound) : C;
      double sum = 0;
      for (int r = 0; r < bound; r++) sum += in_smem ? ph_s[base + r * K + k] : a.phi[(g0 + 1 + r) * KP + k];
      if (what == 0) {
        dsb[f * kMaxClasses + k] = dev_digamma(sum);  // mixdir_impl.cpp:61-72
        a.T[g0 * KP + k] = 0.0;                       // NA row
      } else if (what == 1) {
        dsa[f * kMaxClasses + k] = dev_digamma(sum);  // R-level digamma(sum(phi_jk))
      } else if (a.mode == 1) {
        p2 -= lgamma(sum);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < cells; i += nt) {
    const bool fa = i < cellsA;
    const int f = fa ? 0 : 1;
    const int ii = fa ? i : i - cellsA;
    const int r = ii / K, k = ii - r * K;
    const size_t g = (fa ? g0a : g0b) + 1 + r;
    const double ph = in_smem ? ph_s[i] : a.phi[g * KP + k];
    const double dg = in_smem ? dg_s[i] : dev_digamma(ph);
    const double t = dg - dsb[f * kMaxClasses + k];  // mixdir_impl.cpp:80
    a.T[g * KP + k] = t;
    if (in_smem) ph_s[i] = t;
    if (a.mode == 1) {
      const double da = dg - dsa[f * kMaxClasses + k];
      p0 += (a.beta - 1.0) * da;
      p1 += stat_value(a, g * KP + perm[k]) * t;  // S * T: expec_log_x_cpp collapsed
      p2 -= (ph - 1.0) * da;
    }
  }
  __syncthreads();
  // unit table rows of this element: T2[g2][k] = T[src.x][k] + T[src.y][k]
  const int2 er = a.erange[blockIdx.x];
  if (a.t2src && er.y > 0) {
    for (int i = tid; i < er.y * KP; i += nt) {
      const int g2 = er.x + i / KP, k = i % KP;
      const int2 s = __ldg(&a.t2src[g2]);
      double v = 0.0;
      if (k < K) {
        if (in_smem) {  // source rows are rows of this element's features (row g0: NA, value 0)
          const int ra = s.x - (int)g0a - 1;
          if (ra >= 0) v = ph_s[ra * K + k];
          if (s.y >= 0) {
            const int rb = s.y - (int)g0b - 1;
            if (rb >= 0) v += ph_s[cellsA + rb * K + k];
          }
        } else {
          v = a.T[(size_t)s.x * KP + k];
          if (s.y >= 0) v += a.T[(size_t)s.y * KP + k];
        }
      }
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
    // the old prior's terms and the new class masses (permuted) come in while the sums settle
    double* e1_s = ph_s;                    // [K]
    double* e2_s = ph_s + kMaxClasses;      // [K]
    double* cs_s = ph_s + 2 * kMaxClasses;  // [K] permuted class masses of the new zeta
    const PriorState* st = a.st;  // still the prior this iteration's E-step used (old omega/kappa)
    for (int k = tid; k < K; k += nt) {
      e1_s[k] = st->e1[k];
      e2_s[k] = a.dp ? st->e2[k] : 0.0;
      cs_s[k] = mass_value(a, perm[k]);
    }
    const double termA = st->termA, termE = st->termE;
    const size_t tail = tab + (size_t)(a.stats_fixed ? 2 : 1) * KP;  // [H | underflow | nNA | bad]
    const double H = a.stats[tail], nna = a.stats[tail + 2];
    const double uf = a.stats[tail + 1], bad = a.stats[tail + 3];
    IterStatus* s = a.status;
    const int it = s->iter;
    const int warn0 = s->warn;
    const double prev = it > 0 ? a.trace[it - 1] : 0.0;
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
    if (tid == 0) {
      const double tC = red[0] + a.termC_const, tG = red[2 * kIterThreads];
      const double tD = red[kIterThreads] + (double)K * nna;  // +1 per NA cell per class
      double tB = 0;
      if (!a.dp) {
        for (int k = 0; k < K; k++) tB += cs_s[k] * e1_s[k];  // expec_log_zi, R/mixdir_vi.R:168-170
      } else {
        double tail = 0;  // sum_{l>k} of the permuted class masses
        for (int k = K - 1; k >= 0; k--) {
          tB += tail * e2_s[k] + cs_s[k] * e1_s[k];  // dp_expec_log_zi, dp.R:202-211
          tail += cs_s[k];
        }
      }
      const double elbo = termA + tB + tC + tD + termE + H + tG;
      int conv = 0, warn = warn0;
      if (it != 0 && !isinf(elbo)) {
        const double diff = elbo - prev;
        if (diff < -a.epsilon) warn |= 1;  // R/mixdir_vi.R:105-108
        if (diff < a.epsilon) conv = 1;    // R/mixdir_vi.R:109
      }
      a.trace[it] = elbo;
      s->elbo = elbo;
      s->converged = conv;
      s->warn = warn;
      s->underflow = (uf > 0.0) ? 1 : 0;
      s->bad = (bad > 0.0) ? 1 : 0;
      s->iter = it + 1;
      s->stop = (conv || s->underflow || s->bad) ? 1 : 0;
      *a.underflow_flag = 0;
    }
    for (int k = tid; k < K; k += nt) a.cs_cur[k] = cs_s[k];
    __syncthreads();
    if (a.stats_fixed) {  // the next E+M pass adds into S with atomics
      long long* z = reinterpret_cast<long long*>(a.stats);
      for (size_t i = tid; i < tab + 2 * (size_t)KP; i += nt) z[i] = 0;
    }
    prior_step(a, cs_s, red);
  } else {
    prior_step(a, a.cs_cur, red);
  }
  if (tid == 0) {
    *a.ticket = 0u;
    // mirror for the host's early-exit poll (posted write; the values the fit returns are read after
    // the stream has drained, so no system-wide fence is paid here)
    if (a.status_host && a.mode == 1) *a.status_host = *a.status;
  }
}

}  // namespace mxd
