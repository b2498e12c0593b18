// Inner loop of find_defining_features() (/root/reference/R/predict.R:266-339) on the device.
//
// The reference removes features greedily: with p = predict(all features) and, for every feature i
// still in the model, q_i = predict(remaining features without i) -- both with the arithmetic of
// predict_class (R/predict.R:88-103: prob_z ~ lambda_k * exp(sum_j log U[j,k,x_ij]), plain row
// normalisation, NA and dropped columns contribute 1) -- it computes
//     JS : sum over rows and classes of (q (log q - log p) + p (log p - log q)) / 2   (R/predict.R:304)
//     ARI: 1 - adjusted Rand index of (which.max q, which.max p)                      (R/predict.R:306)
// That is |remaining| predict() calls over all rows per round in R; here ONE pass over the codes gives
// every candidate: per row the logits of the active set are summed once and candidate i only
// subtracts its own term.  Candidate index J stands for "remove nothing more" (R/predict.R:326-331).
#pragma once

#include "kernels.cuh"

namespace mxd {

struct FeatureDivArgs {
  const uint8_t* codes;
  int row_bytes;
  int64_t n_rows;           // rows of this shard
  int64_t row0;             // global index of the shard's first row
  const int64_t* rows;      // optional subsample: global row indices (nullptr: every row)
  int64_t n_sub;
  int J, K, KP;
  const int* rowoff;
  const double* logU;       // [NR][KP], NA rows 0
  const double* lambda;     // [KP]
  const unsigned char* active;  // [J] 1: feature still in the model
  int measure;              // 0 JS, 1 ARI
  double* partial;          // [gridDim.x][J + 1] JS sums of this block
  unsigned long long* cont; // ARI: [J + 1][K][K] counts, index (cand * K + which.max q) * K + which.max p
};

// first maximum over the classes of a warp (8 per lane at most); returns K when every value is NaN
__device__ __forceinline__ int warp_first_argmax(const double* v, int cpl, int K, int lane) {
  double bv = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int k = lane + 32 * c;
    if (c < cpl && k < K && (v[c] > bv || (v[c] == bv && k < bi))) {
      bv = v[c];
      bi = k;
    }
  }
  // warp maximum of the values, then the smallest class index among the lanes that hold it
  const double vmax = warp_max(bv);
  bi = __reduce_min_sync(0xffffffffu, bv == vmax ? bi : 0x7fffffff);
  return bi == 0x7fffffff ? 0 : bi;
}

template <int CB>
__global__ void __launch_bounds__(256) k_feature_div(FeatureDivArgs a) {
  extern __shared__ double acc_s[];  // [warps per block][J + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int J = a.J, K = a.K, KP = a.KP;
  const int cpl = (K + 31) >> 5;
  for (int i = threadIdx.x; i < nwarp * (J + 1); i += blockDim.x) acc_s[i] = 0.0;
  __syncthreads();
  double* acc = acc_s + (size_t)warp * (J + 1);
  const int64_t gw = (int64_t)blockIdx.x * nwarp + warp;
  const int64_t nw = (int64_t)gridDim.x * nwarp;
  const int64_t n = a.rows ? a.n_sub : a.n_rows;
  for (int64_t idx = gw; idx < n; idx += nw) {
    int64_t row = idx;
    if (a.rows) {
      row = a.rows[idx] - a.row0;
      if (row < 0 || row >= a.n_rows) continue;  // another shard's row
    }
    const uint8_t* rp = a.codes + row * a.row_bytes;
    double full[8], act[8];
#pragma unroll
    for (int c = 0; c < 8; c++) full[c] = act[c] = 0;
    for (int j = 0; j < J; j++) {
      const int x = CB == 1 ? rp[j] : reinterpret_cast<const uint16_t*>(rp)[j];
      if (x) {
        const double* t = a.logU + (size_t)(a.rowoff[j] + x) * KP;
        const bool on = a.active[j] != 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int k = lane + 32 * c;
          if (c < cpl && k < K) {
            const double v = __ldg(t + k);
            full[c] += v;
            if (on) act[c] += v;
          }
        }
      }
    }
    // p = predict(all features): lambda_k * exp(rowSums(log U)), row-normalised (R/predict.R:88-103)
    double p[8], lp[8], s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int k = lane + 32 * c;
      p[c] = (c < cpl && k < K) ? a.lambda[k] * exp(full[c]) : 0.0;
      s += p[c];
    }
    s = warp_sum(s);
#pragma unroll
    for (int c = 0; c < 8; c++) {
      p[c] /= s;
      lp[c] = log(p[c]);
    }
    const int ap = a.measure == 1 ? warp_first_argmax(p, cpl, K, lane) : 0;
    for (int i = 0; i <= J; i++) {
      if (i < J && !a.active[i]) continue;
      const int x = i == J ? 0 : (CB == 1 ? rp[i] : reinterpret_cast<const uint16_t*>(rp)[i]);
      const double* t = a.logU + (size_t)((i < J ? a.rowoff[i] : 0) + x) * KP;
      double q[8], sq = 0;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int k = lane + 32 * c;
        q[c] = 0.0;
        if (c < cpl && k < K) q[c] = a.lambda[k] * exp(x ? act[c] - __ldg(t + k) : act[c]);
        sq += q[c];
      }
      sq = warp_sum(sq);
      if (a.measure == 0) {
        double js = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int k = lane + 32 * c;
          if (c < cpl && k < K) {
            const double qq = q[c] / sq, lq = log(qq);
            js += qq * (lq - lp[c]) + p[c] * (lp[c] - lq);  // R/predict.R:304 (0 * Inf = NaN as in R)
          }
        }
        js = warp_sum(js);
        if (lane == 0) acc[i] += js;
      } else {
        // which.max(q / sum(q)) == which.max(q) for a positive, finite sum; keep R's result (the
        // first class) when the row's probabilities are all NaN
        const int aq = (sq > 0.0 && sq < INFINITY) ? warp_first_argmax(q, cpl, K, lane) : 0;
        if (lane == 0) atomicAdd(&a.cont[((size_t)i * K + aq) * K + ap], 1ull);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= J; i += blockDim.x) {
    double t = 0;
    for (int w = 0; w < nwarp; w++) t += acc_s[(size_t)w * (J + 1) + i];
    a.partial[(size_t)blockIdx.x * (J + 1) + i] = t;
  }
}

// fixed-order sum of the per-block partials, halved (the "/ 2" of R/predict.R:304)
__global__ void k_feature_div_reduce(const double* __restrict__ partial, int n_blocks, int n,
                                     double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = 0;
  for (int b = 0; b < n_blocks; b++) t += partial[(size_t)b * n + i];
  out[i] = 0.5 * t;
}

}  // namespace mxd
