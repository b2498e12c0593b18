// FUSED E+M pass with a SORTED-SEGMENT M-step (engine 5, "k_em_seg") -- sm_100a.
// Same job as k_em_units (units.cuh), i.e. in one launch per iteration
//   update_zeta_cpp / update_zeta_dp_cpp     /root/reference/src/mixdir_impl.cpp:57-87,113-152
//   underflow guard + row normalisation      /root/reference/R/mixdir_vi.R:77-82
//   phi M-step statistics                    /root/reference/R/mixdir_vi.R:85-91
//   colSums(zeta), sum_i entrop_zeta(zeta_i) /root/reference/R/mixdir_vi.R:61,101,181-184
// and the same E-step (packed unit codes, pair tables T2 in shared memory, lanes = classes,
// clusters splitting the classes, fp64 logits and soft-max).  What differs is the M-step.
//
// Measured (profiles/r02_microbench_lds_patterns.txt): shared memory delivers 128 bytes per clock and
// SM to the register file whatever the access pattern, so the pass costs what it moves.  A histogram
// read-modify-write per (row, element) moves 16 bytes per class (8 in, 8 out); k_em_units spends two
// thirds of its shared-memory time there.  The codes never change between iterations, so the M-step
// of a feature j over a tile of rows,   S[j, c, k] += sum over the rows with x_ij = c of zeta[i, k],
// can walk the tile's rows SORTED BY CODE: k_pack_seg stores, once per handle, for every (tile,
// feature) the row numbers grouped by code (one byte each) and the group sizes.  A warp then owns a
// feature, reads zeta rows in list order from a shared-memory zeta tile (lanes = classes,
// conflict-free), adds them up in registers and touches the histogram once per (tile, feature,
// code) instead of once per row.  Per (row, feature, class) that moves 4 bytes:
//
//   * zeta enters the M-step as 32-bit FIXED POINT, floor(zeta * (2^32 - 1) + u) with u in [0, 1) a
//     32-bit mix of (global row, class) -- stochastic, i.e. UNBIASED rounding -- and all sums are
//     64-bit INTEGERS: exact and order-independent (SURVEY 8e recommends exactly this).  The class
//     masses colSums(zeta) (they order the classes of the DP variant and enter omega / kappa / lambda)
//     keep full precision as two-word integers (cs_value below).  Per row, logits, soft-max and entropy
//     are fp64, and every row adds its logit terms in the same order wherever it sits in a tile.
//     Together: everything that feeds the next iteration is the same whatever the tile order, the grid
//     size, the number of GPUs or the place where a shard starts -- fits are bit-identical on 1 and N
//     GPUs (only the entropy term of the ELBO is an fp64 sum).  Stated precision: the responsibilities
//     that feed phi carry zero-mean noise of 2^-32 per entry.  (Round-to-nearest was measured first:
//     it sends every responsibility below half a unit to 0, a one-sided error that enters the ELBO's
//     data term weighted with logits of -25 .. -100: 5e-9 .. 9e-9 relative at K = 32 whatever the
//     number of rows.)  What remains is noise of ~1e-11 .. 1e-10 of the ELBO that the coordinate
//     ascent itself amplifies while classes merge (profiles/r02_seg_precision.txt, DESIGN.md section
//     5); the automatic engine choice takes this kernel only from 16384 rows on, and
//     mixdir_b200_set_engine(h, 4) selects fp64 statistics.
//   * no partial buffers and no reduction kernel: every CTA adds its histogram slice to the global
//     int64 statistics with RED.ADD.U64; the last CTA to finish (atomic ticket) sums the per-CTA
//     entropy partials in index order and writes the tail of the statistics buffer.
//   * M-step of the feature split: a lane takes two classes (one 64-bit load), the half-warps take
//     different rows of a list word -- 2.6 instructions per (row, feature).
//
//   * FEATURE SPLIT (FS, clusters of two CTAs with 16 classes each, i.e. K = 17..32): the E-step
//     splits the CLASSES over the cluster as before, but the M-step splits the FEATURES: every CTA
//     writes its 16 responsibilities of a row into BOTH CTAs' zeta tiles (distributed shared
//     memory), so each CTA sees whole 32-class rows (128 bytes = one conflict-free warp load) and
//     owns half of the features for all classes.  Without it (any other cluster shape) a CTA keeps
//     all features for its own classes and the two half-warps of a warp share a feature.
//
// Segment lists (per tile, TB bytes): one block per M-step owner (FS: two, else one), each
// [ group counts: one byte per (feature, code 1..C_j) | row lists ].  Every group is padded to a
// multiple of G entries (G = 4 with FS or KC = 32, else 8) with row number 255, whose zeta-tile
// row is always zero; a count is the group's size in units of G.  Rows with an NA in feature j are
// simply not listed (R/mixdir_vi.R:88 na.rm=TRUE).
#pragma once

#include "units.cuh"

namespace mxd {

// Rounding dither: a 32-bit mix x of (global row, class), u = x / 2^32 in [0, 1).  The same (row,
// class) always gets the same x, so a fit is reproducible run to run and independent of tile order,
// grid size and the number of GPUs.  The lattice point row * C1 + class * C2 is computed once per tile
// and lane (dither_base) and stepped by C1 per row; two xorshift-multiply rounds mix it (uniformity and
// row / class correlations checked on the host, tools/dither_hash_check.py).
constexpr uint32_t kDitherC1 = 0x9E3779B1u, kDitherC2 = 0xC2B2AE3Du;
__device__ __forceinline__ uint32_t dither_base(int64_t row, int k) {
  return (uint32_t)row * kDitherC1 + (uint32_t)(row >> 32) * 0x85EBCA77u + (uint32_t)k * kDitherC2;
}
__device__ __forceinline__ uint32_t dither_mix(uint32_t h) {
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  h *= 0x297A2D39u;
  return h;
}
// floor(zeta * (2^32 - 1) + x / 2^32) = floor(zeta * (2^32 - 1) * 2^32 + x) >> 32: one DFMA, one F2I
constexpr double kFixedScale32 = 4294967295.0 * 4294967296.0;

// Class masses colSums(zeta) order the classes of the DP variant (R/mixdir_vi_dp.R:97) and enter
// omega / kappa / lambda: they keep full precision.  A responsibility zeta in [0, 1] becomes the 63-bit
// integer q = round(zeta * 2^62), split into its two 32-bit halves; both are summed as 64-bit integers
// (no overflow below 2^32 rows, all GPUs together) -- exact and order-independent like the
// statistics, resolution 2^-62 per entry (fp64 sums of a few hundred terms carry more rounding error
// than that).  value = (hi * 2^32 + lo) * 2^-62.
constexpr double kCsScale = 4611686018427387904.0;  // 2^62
constexpr double kCsInv = 1.0 / 4611686018427387904.0;
__host__ __device__ __forceinline__ double cs_value(unsigned long long hi, unsigned long long lo) {
  return ((double)hi * 4294967296.0 + (double)lo) * kCsInv;
}

constexpr int kSegDummyRow = 255;

struct SegArgs {
  const uint32_t* codes;      // packed unit codes, tile-transposed [tile][WPR][R] (as k_em_units)
  const unsigned char* seg;   // [n_tiles][TB] segment lists
  int64_t n_rows;
  int64_t row_base;           // global number of the first row of this launch (seeds the rounding hash)
  int n_tiles;
  int WPR;
  int NR2;                    // rows of the global unit table
  int srows;                  // rows of the shared-memory unit table
  int K, KP, P;
  int TB;                     // bytes of one tile's segment lists (multiple of 16)
  // per M-step owner (FS: cluster rank; else index 0): its block inside a tile's lists, its jobs
  // (features) and its rows of the shared-memory statistics table
  int blk_off[2], blk_size[2];
  int job0[2], njobs[2];
  int sc0[2], SC[2];
  int maxSC, maxJobs, maxBlk;  // shared-memory sizing (maxima over the owners)
  int ws_we;                  // k_em_seg_ws (seg_ws.cuh): number of E-step warps = rows per tile / 16
  int ws_debug;               // timing experiments only (results invalid): 1 = M warps skip their jobs,
                              // 2 = E warps skip gather and soft-max
  int gs;                     // 1: no shared-memory statistics table (SC = 0); jobs[].w is the feature's
                              // first row of the GLOBAL table and every group sum is added there
  const int* eoff_rot;        // as k_em_units
  const int* slot_of_row;
  const double* T2;
  const double* logprior;
  const int4* jobs;           // {idx_off, cnt_off (both inside the owner's block), C_j | feature << 8,
                              //  first statistics row (inside the owner's table)}
  const int* srow_to_stat;    // shared-memory statistics row -> feature-table row
  unsigned long long* stats;  // [NR*KP] fixed-point sums, added with atomics
  unsigned long long* cs_fixed;  // [2][KP] class masses colSums(zeta) as TWO-WORD fixed point (see
                              // cs_split), added with atomics: rows NR and NR + 1 of the table
  double* stats_tail;         // [H | underflow | nNA | bad] doubles, written by the last CTA
  double* Hpart;              // [>= gridDim.x] per-CTA entropy partials
  int n_hsum;                 // CTA partials the last CTA adds up (in index order)
  unsigned int* ticket;
  int* underflow;
  const ScanOut* scan;
  int accumulate;             // chunked first iteration: Hpart accumulates over the launches
  const int* bad;
  const int* stop;
};

struct SegSmem {
  size_t off_T, off_S, off_z, off_tile, off_seg, off_exch, off_ro, off_jobs, off_red, off_bar, off_xbar, total;
  // SW: classes per statistics row (FS: 32, else KC); SC, J, TB: maxima over the M-step owners
  __host__ __device__ SegSmem(int srows, int SC, int SW, int KC, int W, int RG, int WPR, int J, int TB) {
    const int RPW = 32 / KC;
    const size_t R = (size_t)W * RG * RPW;
    const size_t RS = KC * 8 > 128 ? KC * 8 : 128;
    size_t o = 0;
    off_T = o;     o += (size_t)srows * RS;
    off_S = o;     o += (size_t)SC * SW * 8;
    off_z = o;     o += (size_t)256 * 128;
    off_tile = o;  o += R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_seg = o;   o += (size_t)TB + 16;  // + slack: the M-step prefetches one list word ahead
    o = (o + 15) & ~(size_t)15;
    off_exch = o;  o += 2 * R * 16;
    off_ro = o;    o += (size_t)4 * 4 * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_jobs = o;  o += (size_t)J * 16;
    off_red = o;   o += (size_t)(W * KC + W) * 8;
    off_bar = o;   o += 16;
    off_xbar = o;  o += 32 * 8;  // per-warp exchange barriers (FS)
    total = o;
  }
};

// mbarrier across the CTAs of a cluster: arrive on the barrier at the same offset in CTA `cta`
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// read-only during the M-step (written before the barrier that opens it): the compiler may schedule freely
__device__ __forceinline__ uint2 lds_u32x2(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int KC, int MAXT, bool FS>
__global__ void __launch_bounds__(MAXT, 1) k_em_seg(SegArgs a) {
  constexpr int RG = 8;
  constexpr int RPW = 32 / KC;
  constexpr int RW = RG * RPW;
  constexpr uint32_t RS = KC * 8 > 128 ? KC * 8 : 128;
  // zeta tile: 128 bytes per row.  KC = 32: the 32 classes.  KC = 16: TWO copies of the 16 classes,
  // slot 0 of a warp reads the copy in banks 0-15, slot 1 the one in banks 16-31 -- the two rows a
  // warp instruction of the M-step reads are arbitrary, with one copy they would collide in the
  // banks every other time.
  constexpr uint32_t ZS = 128;
  static_assert(KC == 16 || KC == 32, "classes per CTA");
  static_assert(!FS || KC == 16, "feature split: two CTAs of 16 classes");
  constexpr int SW = FS ? 32 : KC;          // classes per statistics row of this CTA
  constexpr int MSLOTS = FS ? 1 : RPW;      // rows a warp instruction of the M-step reads
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
  const int WPR = a.WPR, KP = a.KP;
  const int U = 4 * WPR;
  const int R = W * RW;

  const int own = FS ? rank : 0;  // which M-step owner this CTA is
  const SegSmem L(a.srows, a.maxSC, SW, KC, W, RG, WPR, a.maxJobs, a.maxBlk);
  double* T_s = reinterpret_cast<double*>(smem_raw + L.off_T);
  unsigned long long* S_s = reinterpret_cast<unsigned long long*>(smem_raw + L.off_S);
  uint32_t* z_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_z);
  uint32_t* tile_s = reinterpret_cast<uint32_t*>(smem_raw + L.off_tile);
  unsigned char* seg_s = smem_raw + L.off_seg;
  double2* exch_s = reinterpret_cast<double2*>(smem_raw + L.off_exch);
  int* ro_s = reinterpret_cast<int*>(smem_raw + L.off_ro);
  int4* jobs_s = reinterpret_cast<int4*>(smem_raw + L.off_jobs);
  double* red_s = reinterpret_cast<double*>(smem_raw + L.off_red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(smem_raw + L.off_xbar);

  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;
  if (*a.bad || *a.stop) return;  // uniform over the grid

  // ---- prologue ----
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);  // code tile
    mbar_init(&bar[1], 1);  // segment lists
    if (FS)
      for (int w = 0; w < W; w++) mbar_init(&xbar[w], 1);  // one arrival per tile: the peer's warp w
    mbar_init(&xbar[24], W);                               // "M-step of the tile done": one arrival per warp
    *reinterpret_cast<int*>(&xbar[25]) = 0;                // ... and a count: the last warp refills the lists
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && cluster_id < a.n_tiles) {
    mbar_expect_tx(&bar[0], tile_bytes);
    tma_load_1d(tile_s, a.codes + (size_t)cluster_id * R * WPR, tile_bytes, &bar[0]);
    mbar_expect_tx(&bar[1], (uint32_t)a.blk_size[own]);
    tma_load_1d(seg_s, a.seg + (size_t)cluster_id * a.TB + a.blk_off[own], (uint32_t)a.blk_size[own], &bar[1]);
  }
  const int tab_doubles = a.srows * (int)(RS / 8);
  for (int i = threadIdx.x; i < tab_doubles; i += blockDim.x) T_s[i] = 0.0;
  for (int i = threadIdx.x; i < a.SC[own] * SW; i += blockDim.x) S_s[i] = 0ull;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) z_s[i] = 0u;  // row 255 stays zero
  for (int i = threadIdx.x; i < 4 * U; i += blockDim.x) ro_s[i] = a.eoff_rot[i];
  for (int i = threadIdx.x; i < a.njobs[own]; i += blockDim.x) jobs_s[i] = a.jobs[a.job0[own] + i];
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
  if (P > 1) cluster.sync();

  const uint32_t Tbase = smem_u32(T_s) + kc * 8;
  const uint32_t zbase = smem_u32(z_s) + lane * 4;  // KC = 16 without FS: slot 1 reads the second copy
  uint32_t* z_peer = FS ? cluster.map_shared_rank(z_s, rank ^ 1) : nullptr;
  const uint32_t seg_addr = smem_u32(seg_s);
  double hacc = 0;
  unsigned long long cs_hi = 0ull, cs_lo = 0ull;  // class mass of (this lane's rows, class kg): cs_split
  int uflag = 0;
  const double kTiny = DBL_MIN * log(DBL_MIN);
  const int wrow0 = warp * RW;
  const int row_in_tile0 = wrow0 + slot * RG;
  const unsigned slot_mask = (KC == 32) ? 0xffffffffu : (((1u << KC) - 1u) << (slot * KC));
  // No element rotation here (units.cuh rotates the elements of a code word per slot to spread 64-byte
  // table rows over the banks; with 16 or 32 classes per CTA a table row is a whole 128-byte line): every
  // row adds its logit terms in the SAME order wherever it sits in a tile, so the responsibilities -- and
  // with the integer statistics the whole fit -- do not depend on where a shard starts.
  constexpr int rot = 0;
  const int4* ro_mine = reinterpret_cast<const int4*>(ro_s + rot * U);

  int it = 0;
  for (int t = cluster_id; t < a.n_tiles; t += n_clusters, ++it) {
    const uint32_t par = (uint32_t)(it & 1);
    const uint32_t dbase = dither_base(a.row_base + (int64_t)t * R + row_in_tile0, kg);
    mbar_wait(&bar[0], par);
    const uint32_t* trow = tile_s + row_in_tile0;

    // ---- E-step gather (as k_em_units) ----
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

    // ---- soft-max over classes (R/mixdir_vi.R:82), cluster-wide (as k_em_units) ----
    double2* ex = exch_s + (size_t)par * R;
    double lg[RG], ez[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const double l = (lp + acc[rg]) - 1.0;  // mixdir_impl.cpp:83 / :148
      float mf = __double2float_rn(l);
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
      const double m = (double)mf;
      const double e = exp_nonpos(l - m);  // (-inf - -inf = NaN -> 0)
      double s = e;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const bool live = (__ballot_sync(0xffffffffu, l >= kExpUnderflow) & slot_mask) != 0u;
      lg[rg] = l;
      ez[rg] = e;
      if (kc == 0) ex[row_in_tile0 + rg] = make_double2(m, live ? s : -s);
    }
    if (FS) {
      // Warp w only needs the (max, sum) pairs of the PEER's warp w (same rows): signal and wait
      // warp to warp through an mbarrier instead of stopping both CTAs at a cluster barrier, so the
      // warps drift apart and gather, soft-max arithmetic and exchange latency overlap.  (The
      // cluster barrier below, once per tile, keeps the two CTAs within one phase of each other.)
      __syncwarp();
      // No block barrier closes the M-step of the previous tile (below): before this warp tells the peer
      // "my pairs are written", every warp of THIS CTA must have left the previous zeta tile -- the peer
      // takes the arrival as leave to write into it.  By now that wait is over almost always.
      if (it > 0) mbar_wait(&xbar[24], (uint32_t)((it - 1) & 1));
      if (lane == 0) mbar_arrive_remote(&xbar[warp], (uint32_t)(rank ^ 1));
      mbar_wait_cluster(&xbar[warp], par);
    } else if (P > 1) {
      cluster.sync();
    } else {
      __syncwarp();
    }
    // per-row scalar, once per row by an owner lane: f = exp(m_own - M) / sum, so that zeta = e * f.
    // Entropy of a row (R/mixdir_vi.R:181-184): -sum_k zeta_k log zeta_k with log zeta_k = l_k - c,
    // c = M + log(sum), i.e. c * sum_k zeta_k - sum_k zeta_k l_k = c - sum_k zeta_k l_k: the owner lane adds
    // c once per row (the log stays off the path the other lanes wait on), every lane its zeta_k l_k.
    double f_o = 0.0;
    if (lane < RW) {
      double2* mine_p = ex + wrow0 + lane;
      bool live = false;
      double M = -INFINITY, stot = 0.0, e_mine = 0.0;
      if constexpr (FS) {
        // two CTAs: one of the two weights exp(m - M) is exp(0) = 1 exactly, so ONE exp is enough (the
        // general path below evaluates three, one after the other, while the whole warp waits)
        const double2 v0 = *mine_p, v1 = *cluster.map_shared_rank(mine_p, rank ^ 1);
        M = fmax(v0.x, v1.x);
        live = !signbit(v0.y) || !signbit(v1.y);
        const bool own_max = v0.x >= v1.x;
        const double eo = exp_nonpos((own_max ? v1.x : v0.x) - M);
        const double e0 = own_max ? 1.0 : eo, e1 = own_max ? eo : 1.0;
        const double t0 = (v0.x != -INFINITY) ? fabs(v0.y) * e0 : 0.0;
        const double t1 = (v1.x != -INFINITY) ? fabs(v1.y) * e1 : 0.0;
        stot = rank == 0 ? t0 + t1 : t1 + t0;  // rank order, as the general path
        e_mine = (v0.x == -INFINITY) ? 0.0 : e0;
      } else {
        double2 v[8];
#pragma unroll
        for (int p = 0; p < 8; p++) {
          v[p] = make_double2(-INFINITY, -0.0);
          if (p < P) v[p] = (P > 1 && p != rank) ? *cluster.map_shared_rank(mine_p, p) : *mine_p;
        }
        row_scalars(v, P, rank, M, live, stot, e_mine);  // (units.cuh)
      }
      f_o = e_mine / stot;
      const bool valid = (int64_t)t * R + wrow0 + lane < a.n_rows;
      if (valid && !live) uflag = 1;
      if (valid && rank == 0) hacc += M + log(stot);  // (one CTA of the cluster counts the row's c)
    }
    // (without the feature split the zeta tile is this CTA's own: before the first store below every
    // warp of the CTA must have left the previous tile's M-step -- see the end of the M-step)
    if (!FS && it > 0) mbar_wait(&xbar[24], (uint32_t)((it - 1) & 1));
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const int src = slot * RG + rg;
      const double f = __shfl_sync(0xffffffffu, f_o, src);
      const bool valid = (int64_t)t * R + row_in_tile0 + rg < a.n_rows;
      double zz = 0.0;
      if (valid && kreal) {
        zz = ez[rg] * f;  // zeta / rowSums(zeta)
        hacc -= (zz > DBL_MIN) ? zz * lg[rg] : kTiny;  // (responsibilities below xmin count as xmin, as in R)
      }
      zz = (zz > 0.0) ? zz : 0.0;  // NaN (an underflowing row: the pass is discarded) -> 0
      {  // class mass: two-word integer (cs_value)
        const unsigned long long q = __double2ull_rn(zz * kCsScale);
        cs_hi += q >> 32;
        cs_lo += q & 0xffffffffull;
      }
      // the responsibility as the M-step sees it: 32-bit fixed point, rounded stochastically (see the
      // header): floor(zeta (2^32 - 1) + u), u = mix(row, class) / 2^32 in [0, 1)
      const uint32_t zf = (uint32_t)(__double2ull_rd(
          fma(zz, kFixedScale32, (double)dither_mix(dbase + (uint32_t)rg * kDitherC1))) >> 32);
      if (FS) {  // both CTAs get the whole row: column = global class
        z_s[(row_in_tile0 + rg) * 32 + kg] = zf;
        z_peer[(row_in_tile0 + rg) * 32 + kg] = zf;
      } else {
        z_s[(row_in_tile0 + rg) * 32 + kc] = zf;
        if (KC == 16) z_s[(row_in_tile0 + rg) * 32 + 16 + kc] = zf;
      }
    }
    // zeta tile complete (FS: in both CTAs); nobody reads the code tile any more
    if (FS)
      cluster.sync();
    else
      __syncthreads();

    if (threadIdx.x == 0 && t + n_clusters < a.n_tiles) {
      mbar_expect_tx(&bar[0], tile_bytes);
      tma_load_1d(tile_s, a.codes + (size_t)(t + n_clusters) * R * WPR, tile_bytes, &bar[0]);
    }
    mbar_wait(&bar[1], par);

    // ---- M-step: a warp owns a feature and walks the tile's rows in code order ----
    if (FS && !a.gs) {
      // Feature split: a zeta row is 32 classes x 4 bytes.  A lane takes TWO classes (one 64-bit load)
      // and the two half-warps take different rows of a list word (bytes 0, 2 / 1, 3): a warp
      // instruction covers two rows (two conflict-free 128-byte lines), 2.6 instead of 5 instructions
      // per (row, feature); the halves meet in one shuffle round per code group.
      const int half = lane >> 4, c2 = lane & 15;
      const uint32_t zb2 = smem_u32(z_s) + c2 * 8;
      const uint32_t selA = half ? 0x4441u : 0x4440u, selB = half ? 0x4443u : 0x4442u;
      for (int job = warp; job < a.njobs[own]; job += W) {
        const int4 jd = jobs_s[job];
        const int C = jd.z & 0xff;
        uint32_t p = seg_addr + (uint32_t)jd.x;
        const uint32_t* cnt = reinterpret_cast<const uint32_t*>(seg_s + jd.y);  // 4-byte aligned
        ulonglong2* srow = reinterpret_cast<ulonglong2*>(S_s + (size_t)jd.w * SW) + c2;
        uint32_t w = lds_u32(p);
        uint32_t cw = 0;
        for (int c = 0; c < C; c++, srow += SW / 2) {
          if ((c & 3) == 0) cw = cnt[c >> 2];  // group sizes of four codes at a time
          const int n = (int)((cw >> (8 * (c & 3))) & 0xffu);
          if (n == 0) continue;  // warp-uniform
          const ulonglong2 old = *srow;  // early: its latency hides behind the loop
          unsigned long long sa = 0ull, sb = 0ull;
          // two list words (8 rows) per round; an odd last word is paired with four dummy rows (zero
          // responsibilities), so the loop has no tail block
          for (int g = 0; g < n; g += 2) {
            const bool two = g + 1 < n;
            const uint32_t w1r = lds_u32(p + 4);  // (one word of slack behind the lists)
            const uint32_t w1 = two ? w1r : 0xffffffffu;
            p += two ? 8 : 4;
            const uint32_t wn = two ? lds_u32(p) : w1r;  // next list word
            const uint2 z0 = lds_u32x2(zb2 + __byte_perm(w, 0u, selA) * ZS);
            const uint2 z1 = lds_u32x2(zb2 + __byte_perm(w, 0u, selB) * ZS);
            const uint2 z2 = lds_u32x2(zb2 + __byte_perm(w1, 0u, selA) * ZS);
            const uint2 z3 = lds_u32x2(zb2 + __byte_perm(w1, 0u, selB) * ZS);
            sa += z0.x; sb += z0.y; sa += z1.x; sb += z1.y;
            sa += z2.x; sb += z2.y; sa += z3.x; sb += z3.y;
            w = wn;
          }
          sa += __shfl_xor_sync(0xffffffffu, sa, 16);
          sb += __shfl_xor_sync(0xffffffffu, sb, 16);
          if (half == 0) *srow = make_ulonglong2(old.x + sa, old.y + sb);
        }
      }
    } else
    for (int job = warp; job < a.njobs[own]; job += W) {
      const int4 jd = jobs_s[job];
      const int C = jd.z & 0xff;
      uint32_t p = seg_addr + (uint32_t)jd.x + (FS ? 0 : slot * 4);
      const uint32_t* cnt = reinterpret_cast<const uint32_t*>(seg_s + jd.y);  // 4-byte aligned
      unsigned long long* srow = S_s + (size_t)(a.gs ? 0 : jd.w) * SW + (FS ? lane : kc);
      unsigned long long* grow = a.stats + (size_t)jd.w * KP + (FS ? lane : rank * KC + kc);  // (gs)
      uint32_t w = lds_u32(p);
      uint32_t cw = 0;
      for (int c = 0; c < C; c++, srow += SW) {
        if ((c & 3) == 0) cw = cnt[c >> 2];  // group sizes of four codes at a time
        const int n = (int)((cw >> (8 * (c & 3))) & 0xffu);
        if (n == 0) continue;  // warp-uniform
        const unsigned long long old = a.gs ? 0ull : *srow;  // early: its latency hides behind the loop
        unsigned long long s0 = 0ull, s1 = 0ull;
        int g = 0;
        for (; g + 2 <= n; g += 2) {  // two list words (8 rows per slot) per round
          const uint32_t w1 = lds_u32(p + 4 * MSLOTS);
          p += 8 * MSLOTS;
          const uint32_t wn = lds_u32(p);  // next list word (one word of slack behind the lists)
          const uint32_t z0 = lds_u32(zbase + __byte_perm(w, 0u, 0x4440u) * ZS);
          const uint32_t z1 = lds_u32(zbase + __byte_perm(w, 0u, 0x4441u) * ZS);
          const uint32_t z2 = lds_u32(zbase + __byte_perm(w, 0u, 0x4442u) * ZS);
          const uint32_t z3 = lds_u32(zbase + __byte_perm(w, 0u, 0x4443u) * ZS);
          const uint32_t z4 = lds_u32(zbase + __byte_perm(w1, 0u, 0x4440u) * ZS);
          const uint32_t z5 = lds_u32(zbase + __byte_perm(w1, 0u, 0x4441u) * ZS);
          const uint32_t z6 = lds_u32(zbase + __byte_perm(w1, 0u, 0x4442u) * ZS);
          const uint32_t z7 = lds_u32(zbase + __byte_perm(w1, 0u, 0x4443u) * ZS);
          s0 += z0; s1 += z1; s0 += z2; s1 += z3;
          s0 += z4; s1 += z5; s0 += z6; s1 += z7;
          w = wn;
        }
        if (g < n) {
          p += 4 * MSLOTS;
          const uint32_t wn = lds_u32(p);
          const uint32_t z0 = lds_u32(zbase + __byte_perm(w, 0u, 0x4440u) * ZS);
          const uint32_t z1 = lds_u32(zbase + __byte_perm(w, 0u, 0x4441u) * ZS);
          const uint32_t z2 = lds_u32(zbase + __byte_perm(w, 0u, 0x4442u) * ZS);
          const uint32_t z3 = lds_u32(zbase + __byte_perm(w, 0u, 0x4443u) * ZS);
          s0 += z0; s1 += z1; s0 += z2; s1 += z3;
          w = wn;
        }
        s0 += s1;
        if (MSLOTS == 2) s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        if (FS || slot == 0) {
          if (!a.gs)
            *srow = old + s0;
          else if (s0)
            // 128 (256) contiguous bytes per warp instruction.  The address is formed per group (not
            // stepped in place): stepping would wait for the atomic to have read its address registers
            atomicAdd(grow + (size_t)c * KP, s0);
        }
      }
    }
    {
      // split-phase end of the M-step: a warp arrives and goes on to the next tile's gather; whoever
      // arrives last refills the segment lists (the waits are in the E-step, see above)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cta(&xbar[24]);
        int* cnt_done = reinterpret_cast<int*>(&xbar[25]);
        if (atomicAdd(cnt_done, 1) == W - 1) {
          *cnt_done = 0;
          if (t + n_clusters < a.n_tiles) {
            mbar_expect_tx(&bar[1], (uint32_t)a.blk_size[own]);
            tma_load_1d(seg_s, a.seg + (size_t)(t + n_clusters) * a.TB + a.blk_off[own],
                        (uint32_t)a.blk_size[own], &bar[1]);
          }
        }
      }
      __syncwarp();
    }
  }

  // ---- flush: integer adds into the global statistics; last CTA writes the tail ----
  __syncthreads();  // (the feature-split loop ends without a block barrier)
  {
    {  // (the row map is read in batches of eight: one global-load latency per batch, not per entry)
      const int n = a.SC[own] * SW, step = blockDim.x;
      const int* map = a.srow_to_stat + a.sc0[own];
      for (int i0 = threadIdx.x; i0 < n; i0 += 8 * step) {
        int gr[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * step;
          gr[u] = i < n ? __ldg(map + i / SW) : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * step;
          if (i < n) {
            const unsigned long long v = S_s[i];
            if (v) atomicAdd(&a.stats[(size_t)gr[u] * KP + (FS ? 0 : rank * KC) + i % SW], v);
          }
        }
      }
    }
#pragma unroll
    for (int o = KC; o < 32; o <<= 1) {
      cs_hi += __shfl_xor_sync(0xffffffffu, cs_hi, o);
      cs_lo += __shfl_xor_sync(0xffffffffu, cs_lo, o);
    }
    hacc = warp_sum(hacc);
    unsigned long long* redi_s = reinterpret_cast<unsigned long long*>(red_s);
    for (int word = 0; word < 2; word++) {  // the two words of the class masses, one after the other
      if (slot == 0) redi_s[warp * KC + kc] = word ? cs_lo : cs_hi;
      __syncthreads();
      if (threadIdx.x < KC) {
        unsigned long long c = 0ull;
        for (int w = 0; w < W; w++) c += redi_s[w * KC + threadIdx.x];
        if (c) atomicAdd(&a.cs_fixed[word * KP + rank * KC + threadIdx.x], c);
      }
      __syncthreads();
    }
    if (lane == 0) red_s[W * KC + warp] = hacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double h = 0;
      for (int w = 0; w < W; w++) h += red_s[W * KC + w];
      a.Hpart[blockIdx.x] = a.accumulate ? a.Hpart[blockIdx.x] + h : h;
    }
    if (uflag) atomicExch(a.underflow, 1);
    __threadfence();
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {  // fixed-order sum of the per-CTA entropy partials; flags
      __threadfence();
      if (threadIdx.x == 0) {
        double h = 0;
        for (int c = 0; c < a.n_hsum; c++) h += __ldcg(&a.Hpart[c]);
        a.stats_tail[0] = h;
        a.stats_tail[1] = __ldcg(a.underflow) ? 1.0 : 0.0;
        a.stats_tail[2] = (double)a.scan->n_missing;
        a.stats_tail[3] = a.scan->bad ? 1.0 : 0.0;
        *a.ticket = 0u;
      }
    }
  }
  if (P > 1) cluster.sync();
}

// ---------------------------------------------------------------------------
// segment lists: for every (tile of R rows, feature) the tile's rows grouped by code (stable
// counting sort by one warp), groups padded to multiples of G entries with kSegDummyRow.
// ---------------------------------------------------------------------------
struct PackSegArgs {
  const uint8_t* raw;   // raw codes of the first row of tile 0 of this launch, row-major, 1 byte per cell
  int row_bytes;
  int64_t n_rows;       // rows from `raw` on (rows beyond are NA)
  int n_tiles;
  int R, J, G, TB;
  int split;            // jobs [0, split) belong to owner 0, the rest to owner 1
  int blk_off[2];
  const int4* jobs;     // [J]
  unsigned char* out;   // [n_tiles][TB]
};

__global__ void __launch_bounds__(256) k_pack_seg(PackSegArgs a) {
  __shared__ int cnt_s[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    unsigned char* blk = a.out + (size_t)t * a.TB;
    const int64_t row0 = (int64_t)t * a.R;
    for (int job = warp; job < a.J; job += 8) {
      int4 jd = __ldg(&a.jobs[job]);
      const int C = jd.z & 0xff, feat = jd.z >> 8;
      const int boff = a.blk_off[job < a.split ? 0 : 1];
      jd.x += boff;
      jd.y += boff;
      int* cnt = cnt_s[warp];
      for (int c = lane; c <= C; c += 32) cnt[c] = 0;
      __syncwarp();
      // pass 1: group sizes
      for (int r0 = 0; r0 < a.R; r0 += 32) {
        const int r = r0 + lane;
        int c = 0;
        if (r < a.R && row0 + r < a.n_rows) c = a.raw[(row0 + r) * a.row_bytes + feat];
        if (c) atomicAdd(&cnt[c], 1);
      }
      __syncwarp();
      // padded group starts (exclusive scan over the codes), group counts out
      int base = 0;
      for (int c0 = 1; c0 <= C; c0 += 32) {
        const int c = c0 + lane;
        const int v = (c <= C) ? (cnt[c] + a.G - 1) / a.G * a.G : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        if (c <= C) {
          cnt[c] = base + inc - v;  // start of the group
          blk[jd.y + c - 1] = (unsigned char)(v / a.G);
        }
        base += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
      // fill the used part of the list with the dummy row, then place the rows (stable)
      uint32_t* lw = reinterpret_cast<uint32_t*>(blk + jd.x);
      for (int i = lane; i < base / 4; i += 32) lw[i] = 0xffffffffu;
      __syncwarp();
      for (int r0 = 0; r0 < a.R; r0 += 32) {
        const int r = r0 + lane;
        int c = 0;
        if (r < a.R && row0 + r < a.n_rows) c = a.raw[(row0 + r) * a.row_bytes + feat];
        const unsigned m = __match_any_sync(0xffffffffu, c);
        const int rank = __popc(m & lt);
        if (c) blk[jd.x + cnt[c] + rank] = (unsigned char)r;
        __syncwarp();
        if (c && rank == 0) cnt[c] += __popc(m);
        __syncwarp();
      }
    }
  }
}

// unit-test door: fixed-point statistics -> doubles, in place
__global__ void k_fixed_to_double(double* stats, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    stats[i] = (double)reinterpret_cast<const long long*>(stats)[i] * kFixedInv;
}

}  // namespace mxd
