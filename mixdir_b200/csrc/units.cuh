// FUSED E+M pass over UNITS (engine 4) -- the hot kernel of mixdir_b200 for 1-byte codes (sm_100a).
// Same job as k_em_fused (fused.cuh), i.e. in one launch per iteration
//   update_zeta_cpp / update_zeta_dp_cpp     /root/reference/src/mixdir_impl.cpp:57-87,113-152
//   underflow guard + row normalisation      /root/reference/R/mixdir_vi.R:77-82
//   phi M-step statistics                    /root/reference/R/mixdir_vi.R:85-91
//   colSums(zeta), sum_i entrop_zeta(zeta_i) /root/reference/R/mixdir_vi.R:61,101,181-184
// but built around what the measurements say bounds it: shared-memory wavefronts per code
// (profiles/README.md: 2 per table gather + 4 per histogram read-modify-write of a 256-byte row).
// What differs from k_em_fused:
//
//  * ELEMENTS, not features: the row tiles hold packed unit codes (kernels.cuh "Units"), an
//    element is one feature or a PAIR of features, the tables are the unit tables T2 / S2; a
//    feature that holds no NA anywhere has no NA row (smaller pair tables, more pairs fit).
//  * TILE-TRANSPOSED CODES: [tile][word][row]; the RG rows a lane owns are consecutive, one
//    128-bit load brings a code word of 4 rows.
//  * BANK POSITION BY ELEMENT: with KC = 8 classes per CTA a table row is 64 bytes -- half a
//    128-byte wavefront -- and 4 row slots share a warp instruction.  Which half a slot hits must
//    not depend on the (random) code, or slots collide in the banks (measured x1.5).  Here a
//    128-byte shared-memory row holds Q = 16/KC table rows of DIFFERENT elements side by side:
//    element e owns bank position e mod Q of its rows.  A slot s works on element
//    (e & ~3) | ((e + s) & 3), so the slots of one instruction always cover every position
//    equally often: exactly 2 wavefronts per instruction, whatever the codes.
//  * FINER ROTATION GROUPS + TWO ROWS IN FLIGHT: the M-step rotates warps over groups of FPG
//    (2 or 4) elements instead of code words, so 16-20 warps stay busy with few elements per
//    row; each lane keeps the read-modify-writes of two rows in flight and resolves the case
//    "both rows carry the same code" in registers (the second store then carries both adds).
//  * SOFT-MAX EXCHANGE BY DSMEM READS: every CTA keeps its own (max, sum) per row; after the
//    cluster barrier the owner lanes read the peers' values.  The buffer is P times smaller than
//    an all-to-all of writes, which is what lets 8 row groups per lane fit at P = 8 and 20 warps
//    at P = 2.
//
// Everything else (thread-block clusters splitting the classes, lanes = classes, TMA row tiles,
// owner lanes, fixed-order flush) follows fused.cuh.
#pragma once

#include "fused.cuh"

namespace mxd {

struct UnitArgs {
  const uint32_t* codes;   // packed unit codes, TILE-TRANSPOSED: [tile][WPR][R] 32-bit words, so the
                           // RG rows a lane owns (consecutive inside a tile) lie side by side for
                           // every code word: one 128-bit load brings 4 rows x 4 elements
  int64_t n_rows;
  int n_tiles;
  int WPR;                 // words per packed row; 4 elements per word
  int NR2;                 // rows of the global unit tables
  int srows;               // rows of the shared-memory tables (row 0: zeros / scratch for padding)
  int K, KP;
  int P;                   // cluster size
  const int* eoff_rot;     // [4][4*WPR] byte offset of element (e & ~3) | ((e + r) & 3) in the
                           //            shared-memory tables, r = rotation
  const int* slot_of_row;  // [NR2] global unit-table row -> shared-memory slot (row * Q + position)
  const double* T2;        // [NR2][KP]
  const double* logprior;
  double* Spart;           // [n_clusters][NR2][KP]
  double* cspart;
  double* Hpart;
  int* underflow;
  int accumulate;
  const int* bad;
  const int* stop;         // fit already over (converged / error): do nothing
};

struct UnitSmem {
  size_t off_T, off_S, off_tile, off_exch, off_ro, off_red, off_bar, total;
  __host__ __device__ UnitSmem(int srows, int KC, int W, int RG, int WPR, int P) {
    const int RPW = 32 / KC;
    const size_t R = (size_t)W * RG * RPW;
    const size_t RS = KC * 8 > 128 ? KC * 8 : 128;
    size_t o = 0;
    off_T = o;     o += (size_t)srows * RS;
    off_S = o;     o += (size_t)srows * RS;
    off_tile = o;  o += 2 * R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_exch = o;  o += 2 * R * 16;  // own (max, sum) per row, double buffered; peers READ it (DSMEM)
    off_ro = o;    o += (size_t)4 * 4 * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_red = o;   o += (size_t)(W * KC + W) * 8;
    off_bar = o;   o += 16;
    total = o;
  }
};

// code word w of the RG consecutive rows starting at p (tile-transposed layout): vector loads
template <int RG>
__device__ __forceinline__ void load_code_words(const uint32_t* p, uint32_t* out) {
  if constexpr (RG % 4 == 0) {
#pragma unroll
    for (int i = 0; i < RG / 4; i++) {
      const uint4 v = reinterpret_cast<const uint4*>(p)[i];
      out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    out[0] = v.x; out[1] = v.y;
  }
}

// Row-level soft-max scalars from the (max, sum) pairs of the P CTAs of a cluster (owner lanes, once per
// row): M = largest maximum, stot = sum_p |s_p| exp(m_p - M) in rank order, e_mine = exp(m_own - M).
// This section runs on every warp of the CTA at the same moment with nothing to overlap it, so its
// dependent fp64 chains count: one CTA needs no exp at all (exp(0) = 1 exactly), two need one (the
// larger maximum has weight 1), and for more the exps are evaluated side by side, not one after the
// other.  Same bits as the plain loop  stot += |s_p| * exp(m_p - M).
__device__ __forceinline__ void row_scalars(const double2 (&v)[8], int P, int rank, double& M, bool& live,
                                            double& stot, double& e_mine) {
  M = -INFINITY;
  live = false;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    if (p < P) {
      M = fmax(M, v[p].x);
      live = live || !signbit(v[p].y);
    }
  }
  if (P == 1) {
    const bool on = v[0].x != -INFINITY;
    stot = on ? fabs(v[0].y) : 0.0;
    e_mine = on ? 1.0 : 0.0;
  } else if (P == 2) {
    const bool first_max = v[0].x >= v[1].x;
    const double eo = exp_nonpos((first_max ? v[1].x : v[0].x) - M);
    const double e0 = first_max ? 1.0 : eo, e1 = first_max ? eo : 1.0;
    const double t0 = (v[0].x != -INFINITY) ? fabs(v[0].y) * e0 : 0.0;
    const double t1 = (v[1].x != -INFINITY) ? fabs(v[1].y) * e1 : 0.0;
    stot = t0 + t1;
    e_mine = rank == 0 ? ((v[0].x == -INFINITY) ? 0.0 : e0) : ((v[1].x == -INFINITY) ? 0.0 : e1);
  } else {
    double e[8];
#pragma unroll
    for (int p = 0; p < 8; p++) e[p] = exp_nonpos(v[p].x - M);  // (p >= P: -inf -> 0)
    stot = 0.0;
    e_mine = 0.0;
#pragma unroll
    for (int p = 0; p < 8; p++) {
      if (p < P && v[p].x != -INFINITY) {
        stot += fabs(v[p].y) * e[p];
        if (p == rank) e_mine = e[p];
      }
    }
  }
}

// MAXT: launch bound in threads.  512 (16 warps) leaves the compiler 128 registers; 640 (20 warps,
// <= 102 registers) is used when the tiles of 20 warps fit next to the tables.
template <int KC, int RG, int FPG, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_em_units(UnitArgs a) {
  constexpr int RPW = 32 / KC;                       // rows per warp instruction ("slots")
  // (Q = 16/KC table rows sit side by side in a 128-byte shared-memory row; the placement is
  //  done on the host -- engine.cu: layout_units -- and arrives through eoff_rot / slot_of_row)
  constexpr int RW = RG * RPW;                       // rows per warp per tile
  constexpr uint32_t RS = KC * 8 > 128 ? KC * 8 : 128;  // bytes per shared-memory table row
  static_assert(RPW <= FPG, "slots of a warp must work on distinct elements of a group");
  static_assert(FPG == 2 || FPG == 4, "group size");
  static_assert(RW <= 32, "one owner lane per row");
  static_assert(RG % 2 == 0, "rows are scattered two at a time");
  extern __shared__ __align__(128) unsigned char smem_raw[];

  cg::cluster_group cluster = cg::this_cluster();
  const int P = a.P;
  const int rank = (P > 1) ? (int)cluster.block_rank() : 0;
  const int cluster_id = blockIdx.x / P;
  const int n_clusters = gridDim.x / P;
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane / KC, kc = lane % KC;
  const int kg = rank * KC + kc;  // global (padded) class index of this lane
  const int WPR = a.WPR, KP = a.KP;
  const int U = 4 * WPR;          // elements per row
  const int G = U / FPG;          // M-step rotation groups
  const int R = W * RW;           // rows per tile

  const UnitSmem L(a.srows, KC, W, RG, WPR, P);
  double* T_s = reinterpret_cast<double*>(smem_raw + L.off_T);
  double* S_s = reinterpret_cast<double*>(smem_raw + L.off_S);
  uint32_t* tile_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_tile);
  double2* exch_s = reinterpret_cast<double2*>(smem_raw + L.off_exch);
  int* ro_s = reinterpret_cast<int*>(smem_raw + L.off_ro);
  double* red_s = reinterpret_cast<double*>(smem_raw + L.off_red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;
  // uniform over the grid: a code above C_j would index outside the tables / the fit is over
  if (*a.bad || *a.stop) return;

  // ---- prologue: barriers, first tile in flight, tables into shared memory ----
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
  for (int i = threadIdx.x; i < tab_doubles; i += blockDim.x) {
    T_s[i] = 0.0;
    S_s[i] = 0.0;
  }
  for (int i = threadIdx.x; i < 4 * U; i += blockDim.x) ro_s[i] = a.eoff_rot[i];
  __syncthreads();
  {  // table slice of this CTA: eight independent global loads in flight per thread
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
          v[u] = __ldg(a.T2 + (size_t)g2 * KP + rank * KC + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (i0 + u * step < n) T_s[sl[u]] = v[u];
    }
  }
  const double lp = a.logprior[kg];
  const bool kreal = kg < a.K;
  __syncthreads();
  if (P > 1) cluster.sync();  // every CTA of the cluster is resident before DSMEM traffic

  const uint32_t Tbase = smem_u32(T_s) + kc * 8;
  const uint32_t Sbase = smem_u32(S_s) + kc * 8;
  double csacc = 0, hacc = 0;
  int uflag = 0;
  const double kTiny = DBL_MIN * log(DBL_MIN);
  // A slot (row x this CTA's classes) whose responsibilities are all below 2^-100 is skipped by
  // the M-step: over <= 2^31 rows such terms add up to < 2^-69, far below half an ulp of
  // phi = S + beta >= beta (R/mixdir_vi.R:88), so phi is unchanged bit for bit.
  const double kDeadZ = 7.888609052210118e-31;
  const int wrow0 = warp * RW;                 // first row of this warp inside the tile
  const int row_in_tile0 = wrow0 + slot * RG;  // + rg: a slot owns RG consecutive rows
  const unsigned slot_mask = (KC == 32) ? 0xffffffffu : (((1u << KC) - 1u) << (slot * KC));
  const int rot = slot & 3;               // this slot's element rotation inside a code word
  const int4* ro_mine = reinterpret_cast<const int4*>(ro_s + rot * U);

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
    const uint32_t* trow = tile + row_in_tile0;  // + w*R: word w of this lane's RG rows

    // ---- E-step gather: acc[rg] = sum over the elements of T2[e, u_ie, k]; slot s walks the
    //      elements of a word in the order s, s+1, s+2, s+3 (mod 4) ----
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
      load_code_words<RG>(trow + (size_t)wn * R, cwn);  // software pipeline: next word's codes
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

    // ---- soft-max over classes (R/mixdir_vi.R:82), cluster-wide ----
    double2* ex = exch_s + (size_t)buf * R;
    double lg[RG], ez[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const double l = (lp + acc[rg]) - 1.0;  // mixdir_impl.cpp:83 / :148
      float mf = __double2float_rn(l);        // any shift near the row maximum is exact algebra
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
      const double m = (double)mf;
      const double e = exp_nonpos(l - m);  // (-inf - -inf = NaN -> 0)
      double s = e;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      // exact underflow test of the reference: rowSums(exp(logit)) == 0  (R/mixdir_vi.R:77)
      const bool live = (__ballot_sync(0xffffffffu, l >= kExpUnderflow) & slot_mask) != 0u;
      lg[rg] = l;
      ez[rg] = e;
      // this CTA's (max, sum) of the row stays in its own shared memory; after the cluster barrier
      // the owner lanes of every CTA read the P values through distributed shared memory (an
      // all-to-all of writes would need P times the buffer: 64 KB at P = 8)
      if (kc == 0) ex[row_in_tile0 + rg] = make_double2(m, live ? s : -s);  // sign bit: "not live"
    }
    if (P > 1)
      cluster.sync();
    else
      __syncwarp();
    // owner lanes: one lane per row finishes the row-level scalars once
    uint32_t alive = 0;  // bit rg: some class of this lane's slot has a non-negligible zeta
    double M_o = -INFINITY, f_o = 0.0, logs_o = 0.0;
    if (lane < RW) {
      double2* mine_p = ex + wrow0 + lane;
      double2 v[8];  // P <= 8; the loads are independent, their latencies overlap
#pragma unroll
      for (int p = 0; p < 8; p++) {
        v[p] = make_double2(-INFINITY, -0.0);
        if (p < P) v[p] = (P > 1 && p != rank) ? *cluster.map_shared_rank(mine_p, p) : *mine_p;
      }
      double M, stot, e_mine;
      bool live;
      row_scalars(v, P, rank, M, live, stot, e_mine);
      f_o = e_mine / stot;
      logs_o = log(stot);
      M_o = M;
      const bool valid = (int64_t)t * R + wrow0 + lane < a.n_rows;
      if (valid && !live) uflag = 1;
    }
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const int src = slot * RG + rg;
      const double M = __shfl_sync(0xffffffffu, M_o, src);
      const double f = __shfl_sync(0xffffffffu, f_o, src);
      const double logs = __shfl_sync(0xffffffffu, logs_o, src);
      const bool valid = (int64_t)t * R + row_in_tile0 + rg < a.n_rows;
      double zz = 0.0;
      if (valid && kreal) {
        zz = ez[rg] * f;  // zeta / rowSums(zeta)
        hacc -= (zz > DBL_MIN) ? zz * ((lg[rg] - M) - logs) : kTiny;  // R/mixdir_vi.R:181-184
        csacc += zz;
      }
      ez[rg] = zz;
      if (__ballot_sync(0xffffffffu, zz > kDeadZ) & slot_mask) alive |= 1u << rg;
    }

    // ---- M-step scatter: warps rotate over groups of FPG elements (block barrier between
    //      steps: no two warps on the same histogram rows); slot s of a warp takes element
    //      (s + b) mod FPG of the group at sub-step b; two rows in flight per lane ----
    for (int step = 0; step < G; step++) {
      int g = warp + step;
      if (g >= G) g -= G;
      const int e0 = g * FPG;
      const int wg = e0 >> 2, b0 = e0 & 3;
      uint32_t sb[FPG], sel[FPG];
#pragma unroll
      for (int b = 0; b < FPG; b++) {
        const int i = (slot + b) & (FPG - 1);
        sb[b] = Sbase + (uint32_t)ro_s[e0 + i];
        sel[b] = 0x4440u | (uint32_t)(b0 + i);
      }
      uint32_t cws[RG];
      load_code_words<RG>(trow + (size_t)wg * R, cws);
#pragma unroll
      for (int rg = 0; rg < RG; rg += 2) {
        const uint32_t c0 = cws[rg], c1 = cws[rg + 1];
        const uint32_t on0 = (alive >> rg) & 1u, on1 = (alive >> (rg + 1)) & 1u;
#pragma unroll
        for (int b = 0; b < FPG; b++) {
          const uint32_t ad0 = sb[b] + __byte_perm(c0, 0u, sel[b]) * RS;
          const uint32_t ad1 = sb[b] + __byte_perm(c1, 0u, sel[b]) * RS;
          const double v0 = lds_f64_ordered_if(ad0, on0);
          const double v1 = lds_f64_ordered_if(ad1, on1);
          const double n0 = v0 + ez[rg];
          // same histogram cell twice: the second store must carry both adds
          const double n1 = ((on0 && ad0 == ad1) ? n0 : v1) + ez[rg + 1];
          sts_f64_ordered_if(ad0, n0, on0);
          sts_f64_ordered_if(ad1, n1, on1);
          __syncwarp();
        }
      }
      __syncthreads();
    }
  }

  // ---- flush: this CTA's class slice of S2, class masses, entropy ----
  {
    double* Sp = a.Spart + (size_t)cluster_id * a.NR2 * KP;
    for (int i = threadIdx.x; i < a.NR2 * KC; i += blockDim.x) {
      const int g2 = i / KC, c = i - g2 * KC;
      const double v = S_s[(size_t)__ldg(a.slot_of_row + g2) * KC + c];
      double* dst = &Sp[(size_t)g2 * KP + rank * KC + c];
      *dst = a.accumulate ? *dst + v : v;
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
  if (P > 1) cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

}  // namespace mxd
