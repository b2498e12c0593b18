// Final pass / predict with the log-U table in shared memory (K <= 32 and NR*KC*8 bytes fit).
//   prob_z[i,k] = lambda_k * exp(sum_j log U[j,k,x_ij]), row-normalised WITHOUT max shift, and
//   pred_class = which.max (first maximum): /root/reference/R/mixdir_vi.R:136-148,
//   /root/reference/R/predict.R:88-103.
// Same streaming skeleton as the E-step of k_em_fused (TMA-staged row tiles, lanes = classes,
// conflict-free 8-byte gathers), one CTA per SM, no cluster.  k_predict (kernels.cuh) remains
// the general fallback.
#pragma once

#include "fused.cuh"

namespace mxd {

struct PredictSmemArgs {
  const uint32_t* codes;  // [n_rows + kRowPad][WPR]
  int64_t n_rows;
  int n_tiles;
  int WPR, NR, K, KP;
  const int* rowoff_pad;
  const double* logU;     // [NR][KP], NA row 0
  const double* lambda;   // [KP]
  double* class_prob;     // nullable; n_rows x K column-major
  int32_t* pred_class;    // nullable
};

struct PredictSmem {
  size_t off_T, off_tile, off_ro, off_bar, total;
  __host__ __device__ PredictSmem(int NR, int KC, int W, int RG, int WPR, int FPW) {
    const size_t R = (size_t)W * RG * (32 / KC);
    size_t o = 0;
    off_T = o;     o += (size_t)NR * KC * 8;
    off_tile = o;  o += 2 * R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_ro = o;    o += (size_t)WPR * FPW * 4;
    o = (o + 15) & ~(size_t)15;
    off_bar = o;   o += 16;
    total = o;
  }
};

template <int KC, int CB>
__global__ void __launch_bounds__(512, 1) k_predict_smem(PredictSmemArgs a) {
  constexpr int RG = 8;
  constexpr int RPW = 32 / KC;
  constexpr int FPW = 4 / CB;
  constexpr uint32_t ROWB = KC * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane / KC, kc = lane % KC;
  const int WPR = a.WPR, NR = a.NR, KP = a.KP;
  const int R = W * RG * RPW;
  const PredictSmem L(NR, KC, W, RG, WPR, FPW);
  double* T_s = reinterpret_cast<double*>(smem_raw + L.off_T);
  uint32_t* tile_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_tile);
  int* ro_s = reinterpret_cast<int*>(smem_raw + L.off_ro);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && (int)blockIdx.x < a.n_tiles) {
    mbar_expect_tx(&bar[0], tile_bytes);
    tma_load_1d(tile_s, a.codes + (size_t)blockIdx.x * R * WPR, tile_bytes, &bar[0]);
  }
  for (int i = threadIdx.x; i < NR * KC; i += blockDim.x) {
    const int g = i / KC, c = i - g * KC;
    T_s[i] = c < KP ? a.logU[(size_t)g * KP + c] : 0.0;
  }
  for (int i = threadIdx.x; i < WPR * FPW; i += blockDim.x) ro_s[i] = a.rowoff_pad[i] * (int)ROWB;
  const bool kreal = kc < a.K;
  const double lam = kreal ? a.lambda[kc] : 0.0;
  __syncthreads();

  const uint32_t Tbase = smem_u32(T_s) + kc * 8;
  const int row_in_tile0 = warp * RG * RPW + slot;
  int it = 0;
  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t* tile = tile_s + (size_t)buf * R * WPR;
    if (threadIdx.x == 0 && t + (int)gridDim.x < a.n_tiles) {
      mbar_expect_tx(&bar[buf ^ 1], tile_bytes);
      tma_load_1d(tile_s + (size_t)(buf ^ 1) * R * WPR,
                  a.codes + (size_t)(t + gridDim.x) * R * WPR, tile_bytes, &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    const uint32_t* trow = tile + (size_t)row_in_tile0 * WPR;
    double acc[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) acc[rg] = 0.0;
    for (int w = 0; w < WPR; w++) {
      int ro[FPW];
      load_ro<FPW>(ro_s, w, ro);
      uint32_t tb[FPW];
#pragma unroll
      for (int f = 0; f < FPW; f++) tb[f] = Tbase + (uint32_t)ro[f];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        const uint32_t cw = trow[(size_t)rg * RPW * WPR + w];
#pragma unroll
        for (int f = 0; f < FPW; f++) acc[rg] += lds_f64(tb[f] + code_of<CB>(cw, f) * ROWB);
      }
    }
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const int64_t row = (int64_t)t * R + row_in_tile0 + rg * RPW;
      const double p = kreal ? lam * exp(acc[rg]) : 0.0;
      double s = p;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const double q = p / s;  // prob_z / rowSums(prob_z)
      double bv = kreal ? q : -INFINITY;
      int bi = (kreal && !(q != q)) ? kc : 0x7fffffff;  // NaN never wins (R which.max skips NA)
      if (bi == 0x7fffffff) bv = -INFINITY;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (row < a.n_rows) {
        if (a.class_prob && kreal) a.class_prob[row + (int64_t)kc * a.n_rows] = q;
        if (a.pred_class && kc == 0) a.pred_class[row] = (bi == 0x7fffffff) ? 0 : bi + 1;
      }
    }
    __syncthreads();  // the tile buffer may be refilled
  }
}

}  // namespace mxd
