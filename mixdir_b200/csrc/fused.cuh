// FUSED E+M pass -- the hot kernel of mixdir_b200 (sm_100a).  See DESIGN.md "Fused kernel".
//
// Replaces, in one launch per iteration:
//   update_zeta_cpp / update_zeta_dp_cpp     /root/reference/src/mixdir_impl.cpp:57-87,113-152
//   underflow guard + row normalisation      /root/reference/R/mixdir_vi.R:77-82
//   phi M-step statistics                    /root/reference/R/mixdir_vi.R:85-91
//   colSums(zeta), sum_i entrop_zeta(zeta_i) /root/reference/R/mixdir_vi.R:61,101,181-184
//
//  * persistent thread-block clusters of P CTAs; CTA `rank` owns the KC classes
//    [rank*KC, rank*KC+KC): its slice of the table T and of the statistics S
//    lives in shared memory for the whole launch (no atomics anywhere).
//  * lanes = classes: a warp instruction covers RPW = 32/KC rows ("slots").
//  * row tiles arrive through the TMA engine (cp.async.bulk + mbarrier),
//    double buffered; every CTA of the cluster pulls the same tile (the second
//    read is an L2 hit, DRAM traffic stays at the algorithmic N*J bytes).
//  * E-step: per row, J table gathers from shared memory, added in feature
//    order exactly like mixdir_impl.cpp:77-82; soft-max over the classes with
//    warp shuffles inside the CTA and one (max, sum) exchange over distributed
//    shared memory across the cluster; the per-row scalars (1/sum, log sum,
//    cluster rescale) are computed once per row by an "owner" lane and shuffled
//    back, not once per lane.
//  * M-step: warps own rows; the features are cut into WPR word-groups and warp w
//    works on group (w + t) mod WPR at step t, with a block barrier between
//    steps, so no two warps ever touch the same histogram rows; the RPW slots of
//    a warp start SB features apart and advance in batches of SB features, so
//    they never share a feature inside a batch either.  Plain shared-memory
//    read-modify-write, fp64.
//  * flush: per-cluster partials to global, reduced in fixed order by
//    k_reduce_units (bit-reproducible run to run).
#pragma once

#include "kernels.cuh"

namespace mxd {

struct FusedArgs {
  const uint32_t* codes;  // [n_rows + kRowPad][WPR]
  int64_t n_rows;
  int n_tiles;
  int WPR;
  int NR;
  int K, KP;
  int P;                  // cluster size
  const int* rowoff_pad;  // [WPR*FPW]: category-row offset per (padded) feature
  const double* T;
  const double* logprior;
  double* Spart;
  double* cspart;
  double* Hpart;
  int* underflow;
  int accumulate;         // 1: add to the partial buffers (chunked first iteration), 0: overwrite
  const int* bad;         // invalid codes seen by the upload scan: do nothing
  const int* stop;        // fit already over (converged / error): do nothing
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// table read: the table does not change during the main loop, so the compiler may schedule freely
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// histogram read-modify-write: volatile keeps program order among these two
__device__ __forceinline__ double lds_f64_ordered(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64_ordered(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// predicated forms: lanes with on == 0 issue no shared-memory traffic
__device__ __forceinline__ double lds_f64_ordered_if(uint32_t addr, uint32_t on) {
  double v = 0.0;
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.shared.f64 %0, [%1];\n}\n"
      : "+d"(v)
      : "r"(addr), "r"(on)
      : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64_ordered_if(uint32_t addr, double v, uint32_t on) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.f64 [%0], %1;\n}\n" ::"r"(addr),
               "d"(v), "r"(on)
               : "memory");
}

// code f of a 32-bit code word (CB bytes per code), zero extended: one PRMT
template <int CB>
__device__ __forceinline__ uint32_t code_of(uint32_t cw, int f) {
  if constexpr (CB == 1) return __byte_perm(cw, 0u, 0x4440u | (uint32_t)f);
  return __byte_perm(cw, 0u, f == 0 ? 0x4410u : 0x4432u);
}

template <int FPW>
__device__ __forceinline__ void load_ro(const int* ro_s, int w, int* ro) {
  if constexpr (FPW == 4) {
    const int4 r = reinterpret_cast<const int4*>(ro_s)[w];
    ro[0] = r.x; ro[1] = r.y; ro[2] = r.z; ro[3] = r.w;
  } else {
    const int2 r = reinterpret_cast<const int2*>(ro_s)[w];
    ro[0] = r.x; ro[1] = r.y;
  }
}

// dynamic shared memory carve-up, shared by host (sizing) and device
struct FusedSmem {
  size_t off_T, off_S, off_tile, off_exch, off_ro, off_red, off_bar, total;
  __host__ __device__ FusedSmem(int NR, int KC, int W, int RG, int WPR, int FPW, int P) {
    const int RPW = 32 / KC;
    const size_t R = (size_t)W * RG * RPW;
    size_t o = 0;
    off_T = o;     o += (size_t)NR * KC * 8;
    off_S = o;     o += (size_t)NR * KC * 8;
    off_tile = o;  o += 2 * R * WPR * 4;
    o = (o + 15) & ~(size_t)15;
    off_exch = o;  o += 2 * R * P * 16;
    off_ro = o;    o += (size_t)WPR * FPW * 4;
    o = (o + 15) & ~(size_t)15;
    off_red = o;   o += (size_t)(W * KC + W) * 8;
    off_bar = o;   o += 16;
    total = o;
  }
};

template <int KC, int CB, int RG>
__global__ void __launch_bounds__(512, 1) k_em_fused(FusedArgs a) {
  constexpr int RPW = 32 / KC;   // rows per warp instruction ("slots")
  constexpr int FPW = 4 / CB;    // features per 32-bit code word
  constexpr int SB = FPW / RPW;  // features a slot handles per scatter batch
  constexpr int RW = RG * RPW;   // rows per warp per tile
  constexpr uint32_t ROWB = KC * 8;  // bytes per category row of the class slice
  static_assert(RPW <= FPW, "slots of a warp must be able to take distinct features of a word");
  static_assert(RW <= 32, "one owner lane per row");
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
  const int WPR = a.WPR, NR = a.NR, KP = a.KP;
  const int R = W * RW;  // rows per tile

  const FusedSmem L(NR, KC, W, RG, WPR, FPW, P);
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
  for (int i = threadIdx.x; i < NR * KC; i += blockDim.x) {
    const int g = i / KC, c = i - g * KC;
    T_s[i] = a.T[(size_t)g * KP + rank * KC + c];
    S_s[i] = 0.0;
  }
  for (int i = threadIdx.x; i < WPR * FPW; i += blockDim.x) ro_s[i] = a.rowoff_pad[i] * (int)ROWB;
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
  const int wrow0 = warp * RW;              // first row of this warp inside the tile
  const int row_in_tile0 = wrow0 + slot;    // + rg*RPW
  const unsigned slot_mask = (KC == 32) ? 0xffffffffu : (((1u << KC) - 1u) << (slot * KC));

  int it = 0;
  for (int t = cluster_id; t < a.n_tiles; t += n_clusters, ++it) {
    const int buf = it & 1;
    const uint32_t* tile = tile_s + (size_t)buf * R * WPR;
    // prefetch the next tile of this cluster into the other buffer (its last
    // readers finished at the closing barrier of the previous iteration)
    if (threadIdx.x == 0 && t + n_clusters < a.n_tiles) {
      mbar_expect_tx(&bar[buf ^ 1], tile_bytes);
      tma_load_1d(tile_s + (size_t)(buf ^ 1) * R * WPR,
                  a.codes + (size_t)(t + n_clusters) * R * WPR, tile_bytes, &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    const uint32_t* trow = tile + (size_t)row_in_tile0 * WPR;

    // ---- E-step gather: acc[rg] = sum_j T[j, x_ij, k] in feature order ----
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
        cwn[rg] = trow[(size_t)rg * RPW * WPR + wn];  // software pipeline: next word's codes
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

    // ---- soft-max over classes (R/mixdir_vi.R:82), cluster-wide ----
    double2* ex = exch_s + (size_t)buf * R * P;
    double lg[RG], ez[RG];
#pragma unroll
    for (int rg = 0; rg < RG; rg++) {
      const double l = (lp + acc[rg]) - 1.0;  // mixdir_impl.cpp:83 / :148
      // shift: any value near the row maximum works (soft-max is shift invariant);
      // take the maximum of the fp32-rounded logits: half the shuffle traffic
      float mf = __double2float_rn(l);
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
      const double m = (double)mf;
      const double e = (mf == -INFINITY) ? 0.0 : exp(l - m);
      double s = e;
#pragma unroll
      for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      // exact underflow test of the reference: rowSums(exp(logit)) == 0  (R/mixdir_vi.R:77)
      const bool live = (__ballot_sync(0xffffffffu, l >= kExpUnderflow) & slot_mask) != 0u;
      lg[rg] = l;
      ez[rg] = e;
      if (kc < P) {
        double2* dst = &ex[(size_t)(row_in_tile0 + rg * RPW) * P + rank];
        if (P > 1) dst = cluster.map_shared_rank(dst, kc);
        *dst = make_double2(m, live ? s : -s);  // sign bit of the sum carries "not live"
      }
    }
    if (P > 1)
      cluster.sync();
    else
      __syncwarp();
    // owner lanes: one lane per row finishes the row-level scalars once
    uint32_t alive = 0;  // bit rg: some class of this lane's slot has a non-negligible zeta
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
        zz = ez[rg] * f;  // zeta / rowSums(zeta)
        hacc -= (zz > DBL_MIN) ? zz * ((lg[rg] - M) - logs) : kTiny;  // R/mixdir_vi.R:181-184
        csacc += zz;
      }
      ez[rg] = zz;
      if (__ballot_sync(0xffffffffu, zz > kDeadZ) & slot_mask) alive |= 1u << rg;
    }

    // ---- M-step scatter: rotating feature-group schedule, no atomics ----
    for (int step = 0; step < WPR; step++) {
      int g = warp + step;
      if (g >= WPR) g -= WPR;
      int ro0[FPW];
      load_ro<FPW>(ro_s, g, ro0);
      uint32_t sb[FPW];
#pragma unroll
      for (int f = 0; f < FPW; f++) {
        int r = ro0[f];
#pragma unroll
        for (int sl = 1; sl < RPW; sl++)
          if (slot == sl) r = ro0[(f + SB * sl) % FPW];
        sb[f] = Sbase + (uint32_t)r;
      }
      uint32_t cw[RG];
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        const uint32_t c = trow[(size_t)rg * RPW * WPR + g];
        cw[rg] = __funnelshift_r(c, c, (8 * CB * SB * slot) & 31);  // slot s starts SB*s features in
      }
#pragma unroll
      for (int rg = 0; rg < RG; rg++) {
        const uint32_t on = (alive >> rg) & 1u;
#pragma unroll
        for (int b = 0; b < RPW; b++) {
          uint32_t ad[SB];
          double v[SB];
#pragma unroll
          for (int f = 0; f < SB; f++) {
            ad[f] = sb[b * SB + f] + code_of<CB>(cw[rg], b * SB + f) * ROWB;
            v[f] = lds_f64_ordered_if(ad[f], on);
          }
#pragma unroll
          for (int f = 0; f < SB; f++) sts_f64_ordered_if(ad[f], v[f] + ez[rg], on);
          __syncwarp();
        }
      }
      __syncthreads();
    }
  }

  // ---- flush: this CTA's class slice of S, class masses, entropy ----
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
  if (P > 1) cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

}  // namespace mxd
