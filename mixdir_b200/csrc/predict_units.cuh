// Final pass / predict over UNITS: the arithmetic of predict_class()
//   prob_z[i,k] = lambda_k * exp(sum_j log U[j,k,x_ij]);  class_prob = prob_z / rowSums(prob_z)
//   pred_class  = which.max(class_prob[i,])                 (first maximum, NaN rows -> 0)
// /root/reference/R/predict.R:88-103 == the end of the drivers, /root/reference/R/mixdir_vi.R:136-148
// (no max shift: a row whose terms all underflow is NaN, exactly as in R).
//
// Streaming skeleton of the E-step of k_em_units (units.cuh): packed unit codes (a feature PAIR costs
// one gather), tile-transposed and fetched by bulk copies into a double buffer, the pair table of
// log U in shared memory, lanes = classes, KC classes per CTA, a cluster of P CTAs splits the classes
// when K > KC (K = 64: two CTAs of 32) and exchanges one partial row sum -- and, for pred_class, its
// best class -- per row through distributed shared memory.  There is no statistics table and no zeta
// tile here, so 32 classes of the S4 shape (800 pair rows x 256 bytes) fit ONE CTA.
//
// Output: the RG = 8 rows a lane owns are consecutive, so a lane writes 64 contiguous bytes of its
// class's column of the column-major class_prob matrix (16-byte stores, every 32-byte sector is
// written whole); pred_class goes out as one 32-byte run per row slot.
// Bytes per row: J codes + 8 K (class_prob) + 4 (pred_class); shared-memory gathers per
// row: elements x KP / 16 wavefronts -- like the E-step this pass is bound by them, not by HBM.
#pragma once

#include "units.cuh"

namespace mxd {

struct PredictUnitArgs {
  const uint32_t* codes;   // packed unit codes, tile-transposed [tile][WPR][R]
  int64_t n_rows;
  int n_tiles;
  int R;                   // rows per tile of the packed layout (multiple of RG * 32 / KC)
  int WPR, NR2, srows, K, KP, P;
  const int* eoff_rot;     // [4][4*WPR] (units.cuh)
  const int* slot_of_row;  // [NR2]
  const double* T2;        // [NR2][KP] pair table of log U (NA rows 0)
  const double* lambda;    // [KP], padded classes 0
  double* class_prob;      // nullable; n_rows x K column-major (ld = n_rows)
  int32_t* pred_class;     // nullable
};

struct PredictUnitSmem {
  size_t off_T, off_tile, off_sum, off_best, off_ro, off_bar, total;
  __host__ __device__ PredictUnitSmem(int srows, int KC, int R, int WPR, int P) {
    const size_t RS = KC * 8 > 128 ? KC * 8 : 128;
    size_t o = 0;
    off_T = o;     o += (size_t)srows * RS;
    off_tile = o;  o += 2 * (size_t)R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_sum = o;   o += P > 1 ? (size_t)R * 8 : 0;    // this CTA's partial row sums (peers read them)
    off_best = o;  o += P > 1 ? (size_t)R * 16 : 0;   // (best value, best class) per row
    off_ro = o;    o += (size_t)4 * 4 * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_bar = o;   o += 16;
    total = o;
  }
};

template <int KC>
__global__ void __launch_bounds__(512, 1) k_predict_units(PredictUnitArgs a) {
  constexpr int RG = 8;
  constexpr int RPW = 32 / KC;
  constexpr int RW = RG * RPW;  // rows per warp and slice
  constexpr uint32_t RS = KC * 8 > 128 ? KC * 8 : 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  cg::cluster_group cluster = cg::this_cluster();
  const int P = a.P;
  const int rank = (P > 1) ? (int)cluster.block_rank() : 0;
  const int cluster_id = blockIdx.x / P;
  const int n_clusters = gridDim.x / P;
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane / KC, kc = lane % KC;
  const int kg = rank * KC + kc;
  const int WPR = a.WPR, KP = a.KP, R = a.R;
  const int U = 4 * WPR;
  const int n_slices = R / RW;

  const PredictUnitSmem L(a.srows, KC, R, WPR, P);
  double* T_s = reinterpret_cast<double*>(smem_raw + L.off_T);
  uint32_t* tile_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_tile);
  double* sum_s = reinterpret_cast<double*>(smem_raw + L.off_sum);
  double2* best_s = reinterpret_cast<double2*>(smem_raw + L.off_best);
  int* ro_s = reinterpret_cast<int*>(smem_raw + L.off_ro);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && cluster_id < a.n_tiles) {
    mbar_expect_tx(&bar[0], tile_bytes);
    tma_load_1d(tile_s, a.codes + (size_t)cluster_id * R * WPR, tile_bytes, &bar[0]);
  }
  const int tab_doubles = a.srows * (int)(RS / 8);
  for (int i = threadIdx.x; i < tab_doubles; i += blockDim.x) T_s[i] = 0.0;
  for (int i = threadIdx.x; i < 4 * U; i += blockDim.x) ro_s[i] = a.eoff_rot[i];
  __syncthreads();
  {  // this CTA's class slice of the pair table: eight independent global loads in flight per thread
    const int n = a.NR2 * KC, step = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * step) {
      double v[8];
      int sl[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + u * step;
        if (i < n) {
          const int g2 = i / KC, c = i - g2 * KC;
          sl[u] = __ldg(a.slot_of_row + g2) * KC + c;
          v[u] = rank * KC + c < KP ? __ldg(a.T2 + (size_t)g2 * KP + rank * KC + c) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (i0 + u * step < n) T_s[sl[u]] = v[u];
    }
  }
  const bool kreal = kg < a.K;
  const double lam = kreal ? a.lambda[kg] : 0.0;
  __syncthreads();
  if (P > 1) cluster.sync();

  const uint32_t Tbase = smem_u32(T_s) + kc * 8;
  const int rot = slot & 3;
  const int4* ro_mine = reinterpret_cast<const int4*>(ro_s + rot * U);
  const bool cp_vec = (a.n_rows & 1) == 0;  // columns start 16-byte aligned

  int it = 0;
  for (int t = cluster_id; t < a.n_tiles; t += n_clusters, ++it) {
    const int buf = it & 1;
    const uint32_t* tile = tile_s + (size_t)buf * R * WPR;
    if (threadIdx.x == 0 && t + n_clusters < a.n_tiles) {
      mbar_expect_tx(&bar[buf ^ 1], tile_bytes);
      tma_load_1d(tile_s + (size_t)(buf ^ 1) * R * WPR,
                  a.codes + (size_t)(t + n_clusters) * R * WPR, tile_bytes, &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    double* sums = sum_s;  // (rewritten only after the cluster barrier that ends the tile)
    double2* bests = best_s;

    // a warp takes the slices warp, warp + W, ... of the tile; with a cluster every warp must pass
    // the same number of cluster barriers, so the loop bound is the same for all warps
    const int rounds = (n_slices + W - 1) / W;
    for (int rd = 0; rd < rounds; rd++) {
      const int sl = rd * W + warp;
      const bool work = sl < n_slices;
      const int row_in_tile0 = (work ? sl : 0) * RW + slot * RG;
      const int64_t row0 = (int64_t)t * R + row_in_tile0;
      double p[RG];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) p[rg] = 0.0;
      if (work) {
        const uint32_t* trow = tile + row_in_tile0;
        double acc[RG];
        uint32_t cwn[RG];
#pragma unroll
        for (int rg = 0; rg < RG; rg++) acc[rg] = 0.0;
        load_code_words<RG>(trow, cwn);
        for (int w = 0; w < WPR; w++) {
          const int4 r4 = ro_mine[w];
          uint32_t cw[RG];
          const int wn = (w + 1 < WPR) ? w + 1 : w;
#pragma unroll
          for (int rg = 0; rg < RG; rg++) cw[rg] = __funnelshift_r(cwn[rg], cwn[rg], 8 * rot);
          load_code_words<RG>(trow + (size_t)wn * R, cwn);
          const uint32_t tb0 = Tbase + (uint32_t)r4.x, tb1 = Tbase + (uint32_t)r4.y,
                         tb2 = Tbase + (uint32_t)r4.z, tb3 = Tbase + (uint32_t)r4.w;
#pragma unroll
          for (int rg = 0; rg < RG; rg++) {
            acc[rg] += lds_f64(tb0 + code_of<1>(cw[rg], 0) * RS);
            acc[rg] += lds_f64(tb1 + code_of<1>(cw[rg], 1) * RS);
            acc[rg] += lds_f64(tb2 + code_of<1>(cw[rg], 2) * RS);
            acc[rg] += lds_f64(tb3 + code_of<1>(cw[rg], 3) * RS);
          }
        }
#pragma unroll
        for (int rg = 0; rg < RG; rg++) p[rg] = kreal ? lam * exp(acc[rg]) : 0.0;  // R/predict.R:88-99
      }
      // rowSums(prob_z): over this CTA's classes by shuffles, over the cluster in rank order
      double s[RG];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        double v = p[rg];
#pragma unroll
        for (int o = KC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        s[rg] = v;
      }
      if (P > 1) {
        double mine = 0.0;
#pragma unroll
        for (int rg = 0; rg < RG; rg++)
          if (kc == rg) mine = s[rg];
        if (work && kc < RG) sums[row_in_tile0 + kc] = mine;
        cluster.sync();
        if (work) {
          // lane kc < RG of a slot adds up row kc of the slot over the CTAs, then hands it round
          double tot = 0.0;
          if (kc < RG)
            for (int q = 0; q < P; q++)
              tot += (q == rank) ? sums[row_in_tile0 + kc] : *cluster.map_shared_rank(sums + row_in_tile0 + kc, q);
#pragma unroll
          for (int rg = 0; rg < RG; rg++) s[rg] = __shfl_sync(0xffffffffu, tot, slot * KC + rg);
        }
      }
      // class_prob and this CTA's best class per row
      double bv[RG];
      int bi[RG];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        const double q = p[rg] / s[rg];  // prob_z / rowSums(prob_z)   (R/predict.R:103)
        p[rg] = q;
        double v = kreal ? q : -INFINITY;
        int i = (kreal && !(q != q)) ? kg : 0x7fffffff;  // NaN never wins (R which.max skips NA)
        if (i == 0x7fffffff) v = -INFINITY;
        if (!a.pred_class) {
          // (class_prob only: no arg-max)
        } else if constexpr (KC == 32) {
          // whole warp = one row: the bit pattern of q >= 0 orders like q; maximum of the high words,
          // then of the low words among their holders (REDUX), first holder by ballot
          const unsigned long long key = (i == 0x7fffffff) ? 0ull : (unsigned long long)__double_as_longlong(q) + 1ull;
          const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
          const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
          const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
          const unsigned holders = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
          const int win = __ffs(holders) - 1;
          i = (mh | ml) ? rank * KC + win : 0x7fffffff;
          if (P > 1) v = __shfl_sync(0xffffffffu, v, win);
        } else {
#pragma unroll
          for (int o = KC / 2; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) {
              v = ov;
              i = oi;
            }
          }
        }
        bv[rg] = v;
        bi[rg] = i;
      }
      if (work && a.class_prob && kreal) {
        double* col = a.class_prob + (size_t)kg * a.n_rows + row0;
        if (cp_vec && row0 + RG <= a.n_rows) {
#pragma unroll
          for (int rg = 0; rg < RG; rg += 2) *reinterpret_cast<double2*>(col + rg) = make_double2(p[rg], p[rg + 1]);
        } else {
#pragma unroll
          for (int rg = 0; rg < RG; rg++)
            if (row0 + rg < a.n_rows) col[rg] = p[rg];
        }
      }
      if (a.pred_class) {
        // lane kc < RG of a slot keeps row kc
        double mv = -INFINITY;
        int mi = 0x7fffffff;
#pragma unroll
        for (int rg = 0; rg < RG; rg++)
          if (kc == rg) {
            mv = bv[rg];
            mi = bi[rg];
          }
        if (P > 1) {
          if (work && kc < RG) bests[row_in_tile0 + kc] = make_double2(mv, __hiloint2double(0, mi));
          cluster.sync();
          if (work && rank == 0 && kc < RG) {
            for (int q = 1; q < P; q++) {  // rank order = class order: a later CTA wins only if larger
              const double2 o = *cluster.map_shared_rank(bests + row_in_tile0 + kc, q);
              const int oi = __double2loint(o.y);
              if (o.x > mv || (o.x == mv && oi < mi)) {
                mv = o.x;
                mi = oi;
              }
            }
          }
        }
        if (work && rank == 0 && kc < RG && row0 + kc < a.n_rows)
          a.pred_class[row0 + kc] = (mi == 0x7fffffff) ? 0 : mi + 1;
      }
    }
    if (P > 1)
      cluster.sync();  // peers have read this tile's sums / bests before the buffers are reused
    else
      __syncthreads();  // the tile buffer may be refilled
  }
  if (P > 1) cluster.sync();
}

}  // namespace mxd
