// mixdir_b200 device code (sm_100a).  Hand-written CUDA for the variational-
// inference hot path of const-ae/mixdir:
//   zeta E-step      /root/reference/src/mixdir_impl.cpp:57-87 (fixed K), :113-152 (DP)
//   normalise+guard  /root/reference/R/mixdir_vi.R:77-82
//   phi M-step       /root/reference/R/mixdir_vi.R:85-91
//   ELBO             /root/reference/R/mixdir_vi.R:94-102,164-192; R/mixdir_vi_dp.R:110-118,196-221
//   convergence      /root/reference/R/mixdir_vi.R:105-109
//   final pass       /root/reference/R/mixdir_vi.R:119-148; R/predict.R:88-103
//
// Device data layout (DESIGN.md "Data layout in HBM"):
//   codes   row-major, each row padded to a whole number of 32-bit words (WPR
//           words); a cell is 1 or 2 bytes, 0 = NA.  The allocation carries 512
//           extra all-zero rows so that a row tile may overrun n_rows.
//   "category rows": feature j owns rows rowoff[j] .. rowoff[j]+C_j; row
//           rowoff[j]+0 is the NA row (table value 0 => an NA cell adds nothing,
//           mixdir_impl.cpp:79), row rowoff[j]+x belongs to category x.
//           NR = sum_j (C_j + 1).
//   tables  T[g][k], phi[g][k], S[g][k]: NR x KP doubles, class index fastest,
//           KP >= K (padded classes carry log-prior -inf and never contribute).
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "fastexp.cuh"

namespace mxd {

namespace cg = cooperative_groups;

// exp(x) == 0 in fp64 for x below ln(2^-1075); a row whose largest logit is
// below this has rowSums(zeta) == 0 in the reference (R/mixdir_vi.R:77).
__device__ constexpr double kExpUnderflow = -745.13321910194122211;
// k_em_seg's fixed point for the responsibilities that feed phi: zeta in [0, 1] -> [0, 2^32 - 1]
// (a full 32-bit word, no clamp at zeta = 1); sums of them are exact 64-bit integers
constexpr double kFixedScale = 4294967295.0;
constexpr double kFixedInv = 1.0 / 4294967295.0;
constexpr int kMaxClasses = 256;  // K limit of the engine (8 classes per lane in the generic kernel)
constexpr int kRowPad = 512;      // zero rows appended to the code matrix

// digamma for x > 0: same recipe as the oracle (oracle/mixdir_oracle.c
// oracle_digamma) so that host and device agree to a few ulp.
__device__ __forceinline__ double dev_digamma(double x) {
  if (!(x > 0.0)) return nan("");
  // psi(x) = psi(x + n) - sum_{i<n} 1/(x+i), n = smallest integer with x + n >= 10.  The
  // reciprocals are independent (the oracle's loop divides one after the other: same terms, same
  // order of the additions).
  double acc = 0.0;
  if (x < 10.0) {
    const int n = (int)ceil(10.0 - x);
    double r[10];
#pragma unroll
    for (int i = 0; i < 10; i++) r[i] = (i < n) ? 1.0 / (x + (double)i) : 0.0;
#pragma unroll
    for (int i = 0; i < 10; i++) acc -= r[i];
    x += (double)n;
  }
  const double inv = 1.0 / x;
  const double inv2 = inv * inv;
  double series = 1.0 / 12.0;
  series = series * inv2 - 691.0 / 32760.0;
  series = series * inv2 + 1.0 / 132.0;
  series = series * inv2 - 1.0 / 240.0;
  series = series * inv2 + 1.0 / 252.0;
  series = series * inv2 - 1.0 / 120.0;
  series = series * inv2 + 1.0 / 12.0;
  return acc + (log(x) - 0.5 * inv - series * inv2);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// code scan at upload: NA count, largest code (the reference's n_cat,
// R/mixdir_vi.R:8) and a validity check (code <= C_j).
// ---------------------------------------------------------------------------
struct ScanOut {
  unsigned long long n_missing;
  int max_code;
  int bad;  // some code exceeded its feature's C_j
};

// na_feat (nullable): per-feature NA counts, collected in shared memory (J * 4 bytes of dynamic
// shared memory) -- a feature that never holds an NA needs no NA row in the unit tables.
template <int CB>
__global__ void k_scan_codes(const uint8_t* __restrict__ codes, int row_bytes, int64_t n_rows, int J,
                             const int* __restrict__ ncat, ScanOut* out, unsigned int* na_feat) {
  extern __shared__ unsigned int na_s[];
  if (na_feat) {
    for (int j = threadIdx.x; j < J; j += blockDim.x) na_s[j] = 0u;
    __syncthreads();
  }
  unsigned long long miss = 0;
  int mx = 0, bad = 0;
  // one 32-bit code word per thread and step; (row, word-in-row) advance incrementally so the
  // loop has no 64-bit division
  constexpr int FPW = 4 / CB;
  const int WPR = row_bytes / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t srow = stride / WPR;
  const int scol = (int)(stride - srow * WPR);
  const int64_t w0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = w0 / WPR;
  int col = (int)(w0 - row * WPR);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(codes);
  while (row < n_rows) {
    const uint32_t cw = words[row * WPR + col];
#pragma unroll
    for (int f = 0; f < FPW; f++) {
      const int j = col * FPW + f;
      if (j < J) {
        const int x = CB == 1 ? (int)((cw >> (8 * f)) & 0xffu) : (int)((cw >> (16 * f)) & 0xffffu);
        miss += (x == 0);
        if (x == 0 && na_feat) atomicAdd(&na_s[j], 1u);
        mx = max(mx, x);
        bad |= (x > __ldg(ncat + j));
      }
    }
    row += srow;
    col += scol;
    if (col >= WPR) {
      col -= WPR;
      row++;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    miss += __shfl_xor_sync(0xffffffffu, miss, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out->n_missing, miss);
    atomicMax(&out->max_code, mx);
    if (bad) atomicOr(&out->bad, 1);
  }
  if (na_feat) {
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += blockDim.x)
      if (na_s[j]) atomicAdd(&na_feat[j], na_s[j]);
  }
}

// ---------------------------------------------------------------------------
// prior step: omega / kappa from the class masses of the previous zeta, the
// per-class prior logit, and the ELBO terms that depend on omega/kappa only.
// One block.  K <= kMaxClasses.
// ---------------------------------------------------------------------------
struct PriorState {
  double omega[kMaxClasses];   // omega (dp=0) or kappa1 (dp=1)
  double kappa2[kMaxClasses];  // dp=1 only
  double e1[kMaxClasses];      // psi(omega_k)-psi(sum omega)        | psi(k1)-psi(k1+k2)
  double e2[kMaxClasses];      //                                     | psi(k2)-psi(k1+k2)
  double termA;                // expec_log_lambda (R/mixdir_vi.R:164) | dp_expec_log_v (dp.R:196)
  double termE;                // entrop_omega (R/mixdir_vi.R:177)     | dp_entrop_v (dp.R:214)
};

struct IterStatus {
  double elbo;
  int converged;
  int warn;
  int underflow;
  int iter;  // iterations finished so far
  int bad;   // some code exceeded its feature's number of categories (any rank)
  int stop;  // converged, underflow or invalid codes: every later kernel of the fit returns at once
};

// ---------------------------------------------------------------------------
// results: U = phi / sum(phi) (R/mixdir_vi.R:119-130), final omega/kappa and
// lambda (R/mixdir_vi.R:132-133; R/mixdir_vi_dp.R:148-157), log-U table for the
// final pass, and phi / U in the caller's ragged (j,k,r) layout.
// ---------------------------------------------------------------------------
struct ResultArgs {
  int J, K, KP, dp;
  double a1, a2;
  const int* ncat;
  const int* rowoff;
  const int64_t* ragoff;  // [J] offset of feature j in the ragged layout (= K * sum_{j'<j} C_j')
  const double* phi;      // [NR][KP]
  const double* cs;       // [K] class masses of the final zeta (already permuted)
  double* logU;           // [NR][KP] (NA row 0)
  double* phi_rag;        // ragged out
  double* U_rag;          // ragged out
  double* lambda;         // [K]
  double* loglambda_pad;  // [KP] lambda for the predict kernel (0 for padded classes)
  double* prior_out;      // [K] or [2K]
  int* warn;              // bit1 set if dp and lambda[K-1] > 0.01
};

__global__ void k_results(ResultArgs a) {
  const int K = a.K, KP = a.KP;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < a.J * K) {
    const int j = id / K, k = id - j * K;
    const int C = a.ncat[j];
    const size_t g0 = (size_t)a.rowoff[j];
    double s = 0;
    for (int r = 0; r < C; r++) s += a.phi[(g0 + 1 + r) * KP + k];
    for (int r = 0; r < C; r++) {
      const double ph = a.phi[(g0 + 1 + r) * KP + k];
      const double u = ph / s;
      a.logU[(g0 + 1 + r) * KP + k] = log(u);
      a.phi_rag[a.ragoff[j] + (int64_t)k * C + r] = ph;
      a.U_rag[a.ragoff[j] + (int64_t)k * C + r] = u;
    }
    a.logU[g0 * KP + k] = 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!a.dp) {
      double so = 0;
      for (int k = 0; k < K; k++) so += a.a1 + a.cs[k];
      for (int k = 0; k < K; k++) {
        const double om = a.a1 + a.cs[k];
        a.prior_out[k] = om;
        a.lambda[k] = om / so;
      }
    } else {
      double tail = 0;
      for (int k = K - 1; k >= 0; k--) {
        a.prior_out[K + k] = a.a2 + tail;
        tail += a.cs[k];
      }
      double prod = 1.0;
      for (int k = 0; k < K; k++) {
        const double k1 = a.a1 + a.cs[k];
        a.prior_out[k] = k1;
        const double nu = k1 / (k1 + a.prior_out[K + k]);
        a.lambda[k] = nu * prod;
        prod *= (1.0 - nu);
      }
      if (a.lambda[K - 1] > 0.01) atomicOr(a.warn, 2);
    }
    for (int k = 0; k < KP; k++) a.loglambda_pad[k] = k < K ? a.lambda[k] : 0.0;
  }
}

// caller-supplied ragged table -> device table layout (NA row = fill)
struct ScatterRagArgs {
  int J, K, KP;
  const int* ncat;
  const int* rowoff;
  const int64_t* ragoff;
  const double* rag;
  double* tab;
  int take_log;  // 1: tab = log(rag)
  double na_fill;
};
__global__ void k_ragged_to_table(ScatterRagArgs a) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.J * a.K) return;
  const int j = id / a.K, k = id - j * a.K;
  const int C = a.ncat[j];
  const size_t g0 = (size_t)a.rowoff[j];
  for (int r = 0; r < C; r++) {
    const double v = a.rag[a.ragoff[j] + (int64_t)k * C + r];
    a.tab[(g0 + 1 + r) * a.KP + k] = a.take_log ? log(v) : v;
  }
  a.tab[g0 * a.KP + k] = a.na_fill;
}
// device table layout -> ragged (j,k,r)
__global__ void k_table_to_ragged(ScatterRagArgs a, double* out) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.J * a.K) return;
  const int j = id / a.K, k = id - j * a.K;
  const int C = a.ncat[j];
  const size_t g0 = (size_t)a.rowoff[j];
  for (int r = 0; r < C; r++) out[a.ragoff[j] + (int64_t)k * C + r] = a.tab[(g0 + 1 + r) * a.KP + k];
}

// ---------------------------------------------------------------------------
// statistics buffer:  [ S : NR*KP | cs : KP | H | underflow | nNA | bad ]   (doubles)
// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// Units: what the fused kernels gather and scatter is a UNIT -- one feature, or a PAIR of
// features (ja, jb) whose codes are combined at upload into one byte
//     u = x_a * (C_b + 1) + x_b            (0 = both NA),
// so that one table gather / one histogram read-modify-write serves two features:
//     T2[u] = T[ja, x_a] + T[jb, x_b],     S[ja, x] = sum_{x_b} S2[x * (C_b + 1) + x_b],  etc.
// This halves the shared-memory traffic per code; the price is table rows
// ((C_a+1)(C_b+1) instead of C_a + C_b + 2), so the host planner pairs as many features as the
// shared-memory budget allows (engine.cu: make_units).  Results change only by the
// re-association (T_a + T_b) inside the per-row logit sum.
// ---------------------------------------------------------------------------
struct UnitDesc {
  int off_a;  // byte offset of feature ja inside a raw code row; < 0: padding element (code 0)
  int off_b;  // byte offset of feature jb; < 0: single feature
  int mul;    // number of unit-local codes of feature b: C_b + 1, or C_b if b never holds an NA
  int pad;    // bit 0 / bit 1: feature a / b has no NA row (unit-local code = x - 1)
};

// raw row-major codes (1 byte per cell) -> unit codes, WPRu 32-bit words per row.
// R == 0: row-major output [row][WPRu]; R > 0: tile-transposed output [row / R][WPRu][row % R]
// (k_em_units); `out` and the row numbering start at a tile boundary.
__global__ void k_pack_units(const uint8_t* __restrict__ raw, int row_bytes, int64_t n_rows,
                             const UnitDesc* __restrict__ units, int WPRu, int R,
                             uint32_t* __restrict__ out) {
  const int64_t total = (R > 0 ? (n_rows + R - 1) / R * R : n_rows) * WPRu;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (int64_t)gridDim.x * blockDim.x) {
    int64_t row;
    int col;
    if (R > 0) {  // consecutive threads -> consecutive rows of one word column (coalesced stores)
      const int64_t tile = w / ((int64_t)R * WPRu);
      const int rem = (int)(w - tile * (int64_t)R * WPRu);
      col = rem / R;
      row = tile * R + (rem - col * R);
      if (row >= n_rows) continue;  // tail of the last tile: stays zero
    } else {
      row = w / WPRu;
      col = (int)(w - row * WPRu);
    }
    const uint8_t* r = raw + row * row_bytes;
    uint32_t cw = 0;
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int4 uu = __ldg(reinterpret_cast<const int4*>(units) + col * 4 + f);
      const UnitDesc u = {uu.x, uu.y, uu.z, uu.w};
      uint32_t c = 0;
      if (u.off_a >= 0) {  // pad bit 0 / 1: feature a / b never holds an NA, its codes start at 0
        c = r[u.off_a] - (uint32_t)(u.pad & 1);
        if (u.off_b >= 0) c = c * (uint32_t)u.mul + (r[u.off_b] - (uint32_t)((u.pad >> 1) & 1));
      }
      cw |= (c & 0xffu) << (8 * f);
    }
    out[w] = cw;
  }
}
static_assert(sizeof(UnitDesc) == 16, "UnitDesc is read as int4");

// feature-table row g = (j, x) collects the unit rows start + i*stride, i < count
// (count 0: NA row, its statistics are never used)
struct MargDesc {
  int start, stride, count, pad;
};

// unit table from the feature table: T2[g2][k] = T[src.x][k] + T[src.y][k]  (src.y < 0: single).
// Only the unit-test door (mixdir_b200_estep) needs it as a kernel of its own; inside a fit the
// iteration kernel (iter.cuh) builds T2 element by element.
__global__ void k_build_T2(const int2* __restrict__ src, int NR2, int KP, const double* __restrict__ T,
                           double* __restrict__ T2) {
  const size_t n = (size_t)NR2 * KP;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    const int g2 = (int)(e / KP), k = (int)(e - (size_t)g2 * KP);
    const int2 s = __ldg(&src[g2]);
    double v = T[(size_t)s.x * KP + k];
    if (s.y >= 0) v += T[(size_t)s.y * KP + k];
    T2[e] = v;
  }
}

// deterministic reduction of the per-cluster partials over clusters AND over the partner
// feature's codes: unit statistics S2 -> feature statistics S.
// One block per (feature-table row, 32 classes); the count * n_clusters terms of a row are dealt
// round-robin to the 8 warps (lanes = classes: 256-byte coalesced reads), every warp adds its
// terms in index order and warp 0 adds the 8 partial sums in warp order: a fixed summation tree,
// ~150 dependent loads per thread at S4 instead of the 1184 a thread-per-output loop needs.
struct ReduceUnitsArgs {
  int NR, NR2, KP;
  int n_clusters, n_ctas;
  const double* Spart;   // [n_clusters][NR2*KP]
  const double* cspart;  // [n_clusters][KP]
  const double* Hpart;   // [n_ctas]
  const int* underflow;
  const ScanOut* scan;
  const MargDesc* marg;  // [NR]
  double* stats;
  const int* stop;       // fit already over (converged / error): nothing to do
};
__global__ void __launch_bounds__(256) k_reduce_units(ReduceUnitsArgs a) {
  __shared__ double part[8][32];
  if (a.stop && *a.stop) return;
  const size_t n = (size_t)a.NR * a.KP;
  const size_t n2 = (size_t)a.NR2 * a.KP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, k = blockIdx.y * 32 + lane;
  const int4 mm = __ldg(reinterpret_cast<const int4*>(a.marg) + g);
  const MargDesc m = {mm.x, mm.y, mm.z, mm.w};
  const int terms = m.count * a.n_clusters;
  double s = 0;
  if (k < a.KP) {
    // eight independent loads in flight per thread, added in the same (index) order as before
    for (int t0 = warp; t0 < terms; t0 += 64) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int t = t0 + 8 * u;
        v[u] = 0.0;
        if (t < terms) {
          const int i = t / a.n_clusters, c = t - i * a.n_clusters;
          v[u] = __ldcg(&a.Spart[(size_t)c * n2 + (size_t)(m.start + i * m.stride) * a.KP + k]);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) s += v[u];
    }
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && k < a.KP) {
    double t = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; w++) t += part[w][lane];
    a.stats[(size_t)g * a.KP + k] = t;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    // class masses: warp w takes classes w, w+8, ...; lanes take the clusters round-robin, then a
    // fixed shuffle tree.  Entropy: 256 threads round-robin over the CTAs, then a fixed tree.
    __syncthreads();
    for (int kk = warp; kk < a.KP; kk += 8) {
      double t = 0;
      for (int c = lane; c < a.n_clusters; c += 32) t += a.cspart[(size_t)c * a.KP + kk];
      t = warp_sum(t);
      if (lane == 0) a.stats[n + kk] = t;
    }
    double h = 0;
    for (int c = threadIdx.x; c < a.n_ctas; c += 256) h += a.Hpart[c];
    h = warp_sum(h);
    if (lane == 0) part[0][warp] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int w = 0; w < 8; w++) t += part[0][w];
      a.stats[n + a.KP] = t;
      a.stats[n + a.KP + 1] = (*a.underflow) ? 1.0 : 0.0;
      a.stats[n + a.KP + 2] = (double)a.scan->n_missing;
      a.stats[n + a.KP + 3] = a.scan->bad ? 1.0 : 0.0;
    }
  }
}

// tail of the statistics buffer for the generic kernel: [underflow | nNA | bad]
__global__ void k_tail_to_stats(const int* underflow, const ScanOut* scan, double* slot, const int* stop) {
  if (*stop) return;
  slot[0] = (*underflow) ? 1.0 : 0.0;
  slot[1] = (double)scan->n_missing;
  slot[2] = scan->bad ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------
// generic E+M pass: any K <= 256, any table size; tables stay in global/L2,
// statistics accumulate with fp64 global atomics.  Correct everywhere; used
// when the fused kernel's shared-memory plan does not fit, for the zeta-output
// unit-test door, and as an on-device cross-check of the fused kernel.
// ---------------------------------------------------------------------------
struct GenericArgs {
  const uint8_t* codes;
  int row_bytes;
  int64_t n_rows;
  int J, K, KP;
  const int* rowoff;
  const double* T;         // [NR][KP]
  const double* logprior;  // [KP]
  double* stats;           // S | cs | H | (flag slot unused here); nullable S when only zeta wanted
  int NR;
  int* underflow;
  double* zeta_out;        // nullable; n_rows x K column-major
  int64_t zeta_ld;         // leading dimension (rows) of zeta_out
  int64_t zeta_row0;       // row offset of this launch inside zeta_out
  int do_scatter;
  const int* bad;          // invalid codes seen by the upload scan: do nothing
  const int* stop;         // fit already over (converged / error): do nothing
};

template <int CB>
__global__ void __launch_bounds__(256) k_em_generic(GenericArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int K = a.K, KP = a.KP;
  const int cpl = (K + 31) >> 5;
  if (*a.bad || *a.stop) return;  // a code above C_j would index outside the tables
  double csacc[8];
#pragma unroll
  for (int c = 0; c < 8; c++) csacc[c] = 0;
  double hacc = 0;
  const double kTiny = DBL_MIN * log(DBL_MIN);

  for (int64_t row = gw; row < a.n_rows; row += nw) {
    const uint8_t* rp = a.codes + row * a.row_bytes;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = 0;
    for (int j = 0; j < a.J; j++) {
      const int x = CB == 1 ? rp[j] : reinterpret_cast<const uint16_t*>(rp)[j];
      if (x) {
        const double* t = a.T + (size_t)(a.rowoff[j] + x) * KP;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int k = lane + 32 * c;
          if (c < cpl && k < K) acc[c] += __ldg(t + k);
        }
      }
    }
    double l[8], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      l[c] = (c < cpl && k < K) ? (a.logprior[k] + acc[c]) - 1.0 : -INFINITY;  // impl.cpp:83
      m = fmax(m, l[c]);
    }
    m = warp_max(m);
    if (m < kExpUnderflow) {  // rowSums(zeta) == 0  (R/mixdir_vi.R:77)
      if (lane == 0) atomicExch(a.underflow, 1);
      continue;
    }
    double e[8], s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      e[c] = (l[c] == -INFINITY) ? 0.0 : exp(l[c] - m);
      s += e[c];
    }
    s = warp_sum(s);
    const double logs = log(s);
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      if (c < cpl && k < K) {
        const double z = e[c] / s;  // zeta / rowSums(zeta)  (R/mixdir_vi.R:82)
        e[c] = z;
        csacc[c] += z;
        hacc -= (z > DBL_MIN) ? z * ((l[c] - m) - logs) : kTiny;  // entrop_zeta, R/mixdir_vi.R:181
        if (a.zeta_out) a.zeta_out[a.zeta_row0 + row + (int64_t)k * a.zeta_ld] = z;
      }
    }
    if (a.do_scatter) {
      for (int j = 0; j < a.J; j++) {
        const int x = CB == 1 ? rp[j] : reinterpret_cast<const uint16_t*>(rp)[j];
        if (x) {
          double* sp = a.stats + (size_t)(a.rowoff[j] + x) * KP;
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const int k = lane + 32 * c;
            if (c < cpl && k < K) atomicAdd(sp + k, e[c]);
          }
        }
      }
    }
  }
  if (a.do_scatter) {
    double* csg = a.stats + (size_t)a.NR * KP;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      if (c < cpl && k < K) atomicAdd(csg + k, csacc[c]);
    }
    hacc = warp_sum(hacc);
    if (lane == 0) atomicAdd(csg + KP, hacc);
  }
}

// ---------------------------------------------------------------------------
// final pass / predict:  prob_z[i,k] = lambda_k * exp(sum_j log U[j,k,x_ij]),
// row-normalised without max shift (R/mixdir_vi.R:136-141, R/predict.R:88-103),
// pred_class = which.max (first maximum).
// ---------------------------------------------------------------------------
struct PredictArgs {
  const uint8_t* codes;
  int row_bytes;
  int64_t n_rows;
  int J, K, KP;
  const int* rowoff;
  const double* logU;    // [NR][KP], NA row 0
  const double* lambda;  // [KP]
  double* class_prob;    // nullable; n_rows x K column-major
  int32_t* pred_class;   // nullable
};

template <int CB>
__global__ void __launch_bounds__(256) k_predict(PredictArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int K = a.K, KP = a.KP;
  const int cpl = (K + 31) >> 5;
  for (int64_t row = gw; row < a.n_rows; row += nw) {
    const uint8_t* rp = a.codes + row * a.row_bytes;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = 0;
    for (int j = 0; j < a.J; j++) {
      const int x = CB == 1 ? rp[j] : reinterpret_cast<const uint16_t*>(rp)[j];
      if (x) {
        const double* t = a.logU + (size_t)(a.rowoff[j] + x) * KP;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int k = lane + 32 * c;
          if (c < cpl && k < K) acc[c] += __ldg(t + k);
        }
      }
    }
    double p[8], s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      p[c] = (c < cpl && k < K) ? a.lambda[k] * exp(acc[c]) : 0.0;
      s += p[c];
    }
    s = warp_sum(s);
    double bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      if (c < cpl && k < K) {
        const double q = p[c] / s;
        if (a.class_prob) a.class_prob[row + (int64_t)k * a.n_rows] = q;
        if (q > bv || (q == bv && k < bi)) {
          bv = q;
          bi = k;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (a.pred_class && lane == 0) a.pred_class[row] = (bi == 0x7fffffff) ? 0 : bi + 1;
  }
}
}  // namespace mxd
