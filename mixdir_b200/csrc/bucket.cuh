// FUSED E+M pass, bucketed M-step variant ("engine 3").  Same contract, same E-step and the
// same cluster/TMA structure as k_em_fused (fused.cuh); only the M-step differs:
//
//   k_em_fused   warps own ROWS in the M-step: zeta stays in registers, every (row, feature)
//                costs one shared-memory read-modify-write of the histogram row
//                (2 wavefronts per row and feature and 16-class slice);
//   k_em_bucket  warps own FEATURES in the M-step: zeta of the tile goes to shared memory once,
//                the owner of feature j counting-sorts the tile's rows by their code x_ij
//                (warp match/ballot, no atomics) and then streams the sorted rows, adding
//                zeta rows in a REGISTER accumulator that is flushed to the histogram only when
//                the code changes (C_j flushes per tile instead of R): one zeta-row read
//                (1 wavefront) per row and feature.  NA cells are skipped altogether.
//
// Shared-memory bytes per (row, feature, class): 8 (table gather) + 8 (zeta read) instead of
// 8 + 16.  Summation order inside a histogram entry becomes "sorted by tile, then by row" instead
// of row order; everything stays fp64 (reference: R/mixdir_vi.R:88 sums in row order in long double).
#pragma once

#include "fused.cuh"

namespace mxd {

struct BucketArgs {
  FusedArgs f;
  int J;               // real features
  const int* ncat;     // [J]
  const int* rowoff;   // [J] category-row offset of feature j
};

struct BucketSmem {
  size_t off_T, off_S, off_tile, off_z, off_exch, off_ro, off_meta, off_perm, off_hist, off_red,
      off_bar, total;
  __host__ __device__ BucketSmem(int NR, int KC, int W, int RG, int WPR, int FPW, int P, int J,
                                 int cmax, int entry_bytes) {
    const int RPW = 32 / KC;
    const size_t R = (size_t)W * RG * RPW;
    size_t o = 0;
    off_T = o;     o += (size_t)NR * KC * 8;
    off_S = o;     o += (size_t)NR * KC * 8;
    off_tile = o;  o += 2 * R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_z = o;     o += R * KC * 8;
    off_exch = o;  o += 2 * R * P * 16;
    off_ro = o;    o += (size_t)WPR * FPW * 4;
    o = (o + 15) & ~(size_t)15;
    off_meta = o;  o += (size_t)J * 8;                          // per feature: {rowoff bytes, C_j}
    o = (o + 15) & ~(size_t)15;
    off_perm = o;  o += (size_t)W * (R + 8) * entry_bytes;      // sorted (row, code) entries, per warp
    o = (o + 15) & ~(size_t)15;
    off_hist = o;  o += (size_t)W * (cmax + 2) * 4;             // per warp bucket counters
    o = (o + 15) & ~(size_t)15;
    off_red = o;   o += (size_t)(W * KC + W) * 8;
    off_bar = o;   o += 16;
    total = o;
  }
};

template <int KC, int CB, int RG>
__global__ void __launch_bounds__(512, 1) k_em_bucket(BucketArgs ba) {
  const FusedArgs& a = ba.f;
  constexpr int RPW = 32 / KC;
  constexpr int FPW = 4 / CB;
  constexpr int RW = RG * RPW;
  constexpr uint32_t ROWB = KC * 8;
  static_assert(RW <= 32, "one owner lane per row");
  using entry_t = typename std::conditional<CB == 1, uint16_t, uint32_t>::type;
  constexpr int ESH = CB == 1 ? 8 : 16;  // entry = row | code << ESH
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
  const int WPR = a.WPR, NR = a.NR, KP = a.KP, J = ba.J;
  const int R = W * RW;
  const int row_bytes = WPR * 4;

  int cmax = 0;
  for (int j = 0; j < J; j++) cmax = max(cmax, ba.ncat[j]);  // uniform, tiny
  const BucketSmem L(NR, KC, W, RG, WPR, FPW, P, J, cmax, (int)sizeof(entry_t));
  double* T_s = reinterpret_cast<double*>(smem_raw + L.off_T);
  double* S_s = reinterpret_cast<double*>(smem_raw + L.off_S);
  uint32_t* tile_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_tile);
  double* z_s = reinterpret_cast<double*>(smem_raw + L.off_z);
  double2* exch_s = reinterpret_cast<double2*>(smem_raw + L.off_exch);
  int* ro_s = reinterpret_cast<int*>(smem_raw + L.off_ro);
  int2* meta_s = reinterpret_cast<int2*>(smem_raw + L.off_meta);
  entry_t* perm_s = reinterpret_cast<entry_t*>(smem_raw + L.off_perm) + (size_t)warp * (R + 8);
  int* hist_s = reinterpret_cast<int*>(smem_raw + L.off_hist) + (size_t)warp * (cmax + 2);
  double* red_s = reinterpret_cast<double*>(smem_raw + L.off_red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;
  if (*a.bad) return;  // uniform over the grid: a code above C_j would index outside the tables

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
  for (int i = threadIdx.x; i < NR * KC; i += blockDim.x) {
    const int g = i / KC, c = i - g * KC;
    T_s[i] = a.T[(size_t)g * KP + rank * KC + c];
    S_s[i] = 0.0;
  }
  for (int i = threadIdx.x; i < WPR * FPW; i += blockDim.x) ro_s[i] = a.rowoff_pad[i] * (int)ROWB;
  for (int j = threadIdx.x; j < J; j += blockDim.x)
    meta_s[j] = make_int2(ba.rowoff[j] * (int)ROWB, ba.ncat[j]);
  const double lp = a.logprior[kg];
  const bool kreal = kg < a.K;
  __syncthreads();
  if (P > 1) cluster.sync();

  const uint32_t Tbase = smem_u32(T_s) + kc * 8;
  const uint32_t Sbase = smem_u32(S_s) + kc * 8;
  const uint32_t Zbase = smem_u32(z_s) + kc * 8;
  const uint32_t Pbase = smem_u32(perm_s);
  double csacc = 0, hacc = 0;
  int uflag = 0;
  const double kTiny = DBL_MIN * log(DBL_MIN);
  const int wrow0 = warp * RW;
  const int row_in_tile0 = wrow0 + slot;
  const unsigned slot_mask = (KC == 32) ? 0xffffffffu : (((1u << KC) - 1u) << (slot * KC));
  const unsigned lanemask_lt = (1u << lane) - 1u;

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
    const uint32_t* trow = tile + (size_t)row_in_tile0 * WPR;

    // ---- E-step gather (identical to k_em_fused) ----
    double acc[RG];
    uint32_t cwn[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      acc[rg] = 0.0;
      cwn[rg] = trow[(size_t)rg * RPW * WPR];
    }
    for (int w = 0; w < WPR; w++) {
      int ro[FPW];
      load_ro<FPW>(ro_s, w, ro);
      uint32_t cw[RG];
      const int wn = (w + 1 < WPR) ? w + 1 : w;
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        cw[rg] = cwn[rg];
        cwn[rg] = trow[(size_t)rg * RPW * WPR + wn];
      }
      uint32_t tb[FPW];
#pragma unroll
      for (int f = 0; f < FPW; f++) tb[f] = Tbase + (uint32_t)ro[f];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
#pragma unroll
        for (int f = 0; f < FPW; f++) acc[rg] += lds_f64(tb[f] + code_of<CB>(cw[rg], f) * ROWB);
      }
    }

    // ---- soft-max over classes, cluster-wide (identical to k_em_fused) ----
    double2* ex = exch_s + (size_t)buf * R * P;
    double lg[RG], ez[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const double l = (lp + acc[rg]) - 1.0;
      float mf = __double2float_rn(l);
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
      const double m = (double)mf;
      const double e = (mf == -INFINITY) ? 0.0 : exp(l - m);
      double s = e;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const bool live = (__ballot_sync(0xffffffffu, l >= kExpUnderflow) & slot_mask) != 0u;
      lg[rg] = l;
      ez[rg] = e;
      if (kc < P) {
        double2* dst = &ex[(size_t)(row_in_tile0 + rg * RPW) * P + rank];
        if (P > 1) dst = cluster.map_shared_rank(dst, kc);
        *dst = make_double2(m, live ? s : -s);
      }
    }
    if (P > 1)
      cluster.sync();
    else
      __syncwarp();
    double M_o = -INFINITY, f_o = 0.0, logs_o = 0.0;
    if (lane < RW) {
      const double2* e = ex + (size_t)(wrow0 + lane) * P;
      double M = -INFINITY;
      bool live = false;
      for (int p = 0; p < P; p++) {
        M = fmax(M, e[p].x);
        live = live || !signbit(e[p].y);
      }
      double stot = 0.0;
      for (int p = 0; p < P; p++)
        if (e[p].x != -INFINITY) stot += fabs(e[p].y) * exp(e[p].x - M);
      const double mine = e[rank].x;
      f_o = (mine == -INFINITY) ? 0.0 : exp(mine - M) / stot;
      logs_o = log(stot);
      M_o = M;
      const bool valid = (int64_t)t * R + wrow0 + lane < a.n_rows;
      if (valid && !live) uflag = 1;
    }
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const int src = rg * RPW + slot;
      const double M = __shfl_sync(0xffffffffu, M_o, src);
      const double f = __shfl_sync(0xffffffffu, f_o, src);
      const double logs = __shfl_sync(0xffffffffu, logs_o, src);
      const bool valid = (int64_t)t * R + row_in_tile0 + rg * RPW < a.n_rows;
      double zz = 0.0;
      if (valid && kreal) {
        zz = ez[rg] * f;
        hacc -= (zz > DBL_MIN) ? zz * ((lg[rg] - M) - logs) : kTiny;
        csacc += zz;
      }
      // zeta of the tile -> shared memory (row-major, this CTA's KC classes)
      sts_f64_ordered(Zbase + (uint32_t)(row_in_tile0 + rg * RPW) * ROWB, zz);
    }
    __syncthreads();  // zeta tile complete

    // ---- M-step: feature-owned, counting sort by code, register accumulation ----
    const uint8_t* tile8 = reinterpret_cast<const uint8_t*>(tile);
    for (int j = warp; j < J; j += W) {
      const int2 meta = meta_s[j];
      const int C = meta.y;
      const uint32_t srow = Sbase + (uint32_t)meta.x;
      // pass 1: bucket sizes (code 0 = NA is not counted; rows past n_rows are all-NA padding)
      for (int c = lane; c <= C; c += 32) hist_s[c] = 0;
      __syncwarp();
      for (int r0 = 0; r0 < R; r0 += 32) {
        const uint8_t* cp = tile8 + (size_t)(r0 + lane) * row_bytes + j * CB;
        const int x = CB == 1 ? (int)*cp : (int)*reinterpret_cast<const uint16_t*>(cp);
        const unsigned m = __match_any_sync(0xffffffffu, x);
        if (x != 0 && (m & lanemask_lt) == 0) hist_s[x] += __popc(m);  // one leader per value
        __syncwarp();
      }
      // exclusive prefix sum over the buckets 1..C -> start positions
      int carry = 0;
      for (int c0 = 1; c0 <= C; c0 += 32) {
        const int c = c0 + lane;
        const int v = c <= C ? hist_s[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        __syncwarp();
        if (c <= C) hist_s[c] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      const int n = carry;  // non-NA rows of this feature in the tile
      __syncwarp();
      // pass 2: fill (row, code) entries in bucket order
      for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        const uint8_t* cp = tile8 + (size_t)r * row_bytes + j * CB;
        const int x = CB == 1 ? (int)*cp : (int)*reinterpret_cast<const uint16_t*>(cp);
        const unsigned m = __match_any_sync(0xffffffffu, x);
        if (x != 0) perm_s[hist_s[x] + __popc(m & lanemask_lt)] = (entry_t)(r | (x << ESH));
        __syncwarp();
        if (x != 0 && (m & lanemask_lt) == 0) hist_s[x] += __popc(m);
        __syncwarp();
      }
      if (lane < 8) perm_s[n + lane] = 0;  // sentinel entries (code 0 never occurs in the list)
      __syncwarp();
      // stream the sorted list: slot s takes the chunk range [s*nch/RPW, (s+1)*nch/RPW) of
      // 4-entry chunks; accumulate in a register, flush when the code changes
      const int nch = (n + 3) >> 2;
      const int ch_lo = slot * nch / RPW, ch_hi = (slot + 1) * nch / RPW;
      const int nit = (nch + RPW - 1) / RPW;  // uniform trip count
      int cur = 0;                            // current code (0 = none yet)
      double sacc = 0.0;
      for (int i = 0; i < nit; i++) {
        const int ch = ch_lo + i;
        const bool act = ch < ch_hi;
        uint32_t e4[4];
        if (CB == 1) {
          const uint2 v = *reinterpret_cast<const uint2*>(perm_s + (act ? ch : 0) * 4);
          e4[0] = v.x & 0xffffu; e4[1] = v.x >> 16; e4[2] = v.y & 0xffffu; e4[3] = v.y >> 16;
        } else {
          const uint4 v = *reinterpret_cast<const uint4*>(perm_s + (act ? ch : 0) * 4);
          e4[0] = v.x; e4[1] = v.y; e4[2] = v.z; e4[3] = v.w;
        }
        double zv[4];  // the four zeta rows first (sentinel entries read row 0: harmless)
#pragma unroll
        for (int q = 0; q < 4; q++)
          zv[q] = lds_f64_ordered(Zbase + (e4[q] & ((1u << ESH) - 1u)) * ROWB);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t e = act ? e4[q] : 0u;
          const int c = (int)(e >> ESH);
          if (c != 0) {
            if (c != cur) {
              if (cur != 0) {
                const uint32_t ad = srow + (uint32_t)cur * ROWB;
                sts_f64_ordered(ad, lds_f64_ordered(ad) + sacc);
              }
              cur = c;
              sacc = 0.0;
            }
            sacc += zv[q];
          }
        }
      }
      // last bucket of every slot, one slot at a time (neighbouring slots may end/start in the
      // same bucket)
#pragma unroll
      for (int s = 0; s < RPW; s++) {
        __syncwarp();
        if (slot == s && cur != 0) {
          const uint32_t ad = srow + (uint32_t)cur * ROWB;
          sts_f64_ordered(ad, lds_f64_ordered(ad) + sacc);
        }
      }
      __syncwarp();
    }
    __syncthreads();  // zeta tile and code tile may be overwritten
  }

  // ---- flush (identical to k_em_fused) ----
  {
    double* Sp = a.Spart + (size_t)cluster_id * NR * KP;
    for (int i = threadIdx.x; i < NR * KC; i += blockDim.x) {
      const int g = i / KC, c = i - g * KC;
      double* dst = &Sp[(size_t)g * KP + rank * KC + c];
      *dst = a.accumulate ? *dst + S_s[i] : S_s[i];
    }
#pragma unroll
    for (int o = KC; o < 32; o <<= 1) csacc += __shfl_xor_sync(0xffffffffu, csacc, o);
    hacc = warp_sum(hacc);
    if (slot == 0) red_s[warp * KC + kc] = csacc;
    if (lane == 0) red_s[W * KC + warp] = hacc;
    __syncthreads();
    if (threadIdx.x < KC) {
      double s = 0;
      for (int w = 0; w < W; w++) s += red_s[w * KC + threadIdx.x];
      double* dst = &a.cspart[(size_t)cluster_id * KP + rank * KC + threadIdx.x];
      *dst = a.accumulate ? *dst + s : s;
    }
    if (threadIdx.x == 0) {
      double h = 0;
      for (int w = 0; w < W; w++) h += red_s[W * KC + w];
      a.Hpart[blockIdx.x] = a.accumulate ? a.Hpart[blockIdx.x] + h : h;
    }
    if (uflag) atomicExch(a.underflow, 1);
  }
  if (P > 1) cluster.sync();
}

}  // namespace mxd
