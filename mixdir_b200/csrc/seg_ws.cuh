// k_em_seg_ws: the feature-split instance of k_em_seg (seg.cuh: K = 17..32, clusters of two CTAs with
// 16 classes each) with WARP SPECIALISATION -- sm_100a.  Same arithmetic, same data structures, same
// results bit for bit; what changes is who does what when.
//
// k_em_seg runs the E-step and the M-step of a tile one after the other on the same 15 warps, and the
// time of a tile is the latency of one warp's chain through both (profiles/README.md: tile time
// independent of the warp count, issue slots 62 % and shared-memory pipe 62 % busy).  More warps would
// overlap more chains, but 15 warps x 120 registers is the register file.  Here a CTA has 32 warps:
//
//   warps 0..WE-1 (WE = rows per tile / 16, at most 15)   E-step warps, `setmaxnreg.inc` to RE registers:
//       gather + soft-max + exchange of tile t+1 while the M-step of tile t is still running;
//       the finished responsibilities wait in registers until the zeta tile is free
//   warp 15                                                copy warp: refills the code tile and the
//       segment lists (bulk copies) as soon as their readers have arrived on the `free` barriers
//   warps 16..31                                           M-step warps, `setmaxnreg.dec` to RM registers:
//       walk the code-sorted row lists of their features over the zeta tile
//
// Hand-over through mbarriers, all cluster-wide because every CTA writes its 16 classes of a row into
// BOTH CTAs' zeta tiles:   zfull (2 WE arrivals: the E warps of both CTAs have stored tile t)
//                          zfree (32 arrivals: the M warps of both CTAs have finished tile t)
// plus tile_free (WE) and seg_free (16) inside the CTA.  One zeta tile is enough: the E warps of tile
// t+1 compute for longer than the M warps of tile t need.
#pragma once

#include "seg.cuh"

namespace mxd {

// arrive on a barrier of THIS CTA, releasing at cluster scope (the waiters may be fed by the peer too)
__device__ __forceinline__ void mbar_arrive_local_cluster(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int RE, int RM>
__global__ void __launch_bounds__(1024, 1) k_em_seg_ws(SegArgs a) {
  constexpr int KC = 16, RG = 8, RPW = 2, RW = RG * RPW;
  constexpr uint32_t RS = 128, ZS = 128;
  constexpr int SW = 32;
  constexpr int MW = 16;  // M-step warps
  extern __shared__ __align__(128) unsigned char smem_raw[];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();  // P = 2
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int WE = a.ws_we;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane / KC, kc = lane % KC;
  const int kg = rank * KC + kc;
  const int WPR = a.WPR, KP = a.KP;
  const int U = 4 * WPR;
  const int R = WE * RW;
  const int own = rank;

  const SegSmem L(a.srows, a.maxSC, SW, KC, WE, RG, WPR, a.maxJobs, a.maxBlk);
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
  uint64_t* zfull = xbar + 16;
  uint64_t* zfree = xbar + 17;
  uint64_t* tile_free = xbar + 18;
  uint64_t* seg_free = xbar + 19;

  const uint32_t tile_bytes = (uint32_t)R * WPR * 4;
  if (*a.bad || *a.stop) return;  // uniform over the grid

  // ---- prologue (all warps) ----
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);  // code tile landed
    mbar_init(&bar[1], 1);  // segment lists landed
    for (int w = 0; w < WE; w++) mbar_init(&xbar[w], 1);  // soft-max exchange, warp to warp
    mbar_init(zfull, 2 * WE);
    mbar_init(zfree, 2 * MW);
    mbar_init(tile_free, WE);
    mbar_init(seg_free, MW);
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
  {
    const int n = a.NR2 * KC, step = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * step) {
      double v[4];
      int sl[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * step;
        if (i < n) {
          const int g2 = i / KC, c = i - g2 * KC;
          sl[u] = __ldg(a.slot_of_row + g2) * KC + c;
          v[u] = __ldg(a.T2 + (size_t)g2 * KP + rank * KC + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i0 + u * step < n) T_s[sl[u]] = v[u];
    }
  }
  __syncthreads();
  cluster.sync();

  double hacc = 0;
  unsigned long long cs_hi = 0ull, cs_lo = 0ull;
  int uflag = 0;

  if (warp < MW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RE));
    if (warp < WE) {
      // =========================== E-step warps ===========================
      const double lp = a.logprior[kg];
      const bool kreal = kg < a.K;
      const uint32_t Tbase = smem_u32(T_s) + kc * 8;
      uint32_t* z_peer = cluster.map_shared_rank(z_s, rank ^ 1);
      const double kTiny = DBL_MIN * log(DBL_MIN);
      const int wrow0 = warp * RW;
      const int row_in_tile0 = wrow0 + slot * RG;
      const unsigned slot_mask = ((1u << KC) - 1u) << (slot * KC);
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

        // ---- gather (as k_em_seg) ----
        double acc[RG];
        uint32_t cwn[RG];
#pragma unroll
        for (int rg = 0; rg < RG; rg++) acc[rg] = 0.0;
        load_code_words<RG>(trow, cwn);
        for (int w = 0; w < ((a.ws_debug & 2) ? 0 : WPR); w++) {
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
        // this warp has read its rows of the code tile: the copy warp may refill it
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(tile_free);

        // ---- soft-max over classes (R/mixdir_vi.R:82), cluster-wide ----
        double2* ex = exch_s + (size_t)par * R;
        double lg[RG], ez[RG];
#pragma unroll
        for (int rg = 0; rg < RG; rg++) {
          const double l = (lp + acc[rg]) - 1.0;  // mixdir_impl.cpp:83 / :148
          float mf = __double2float_rn(l);
#pragma unroll
          for (int o = KC / 2; o > 0; o >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
          const double m = (double)mf;
          const double e = exp_nonpos(l - m);
          double s = e;
#pragma unroll
          for (int o = KC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const bool live = (__ballot_sync(0xffffffffu, l >= kExpUnderflow) & slot_mask) != 0u;
          lg[rg] = l;
          ez[rg] = e;
          if (kc == 0) ex[row_in_tile0 + rg] = make_double2(m, live ? s : -s);
        }
        __syncwarp();
        if (!(a.ws_debug & 16)) {  // (16: timing experiment without the exchange across the cluster)
          if (lane == 0) mbar_arrive_remote(&xbar[warp], (uint32_t)(rank ^ 1));
          mbar_wait_cluster(&xbar[warp], par);
        }

        double f_o = 0.0;  // per-row scalar (seg.cuh): exp(m_own - M) / sum; the owner lane adds the row's c
        if (lane < RW) {
          double2* mine_p = ex + wrow0 + lane;
          const double2 v0 = *mine_p;
          const double2 v1 = (a.ws_debug & 16) ? v0 : *cluster.map_shared_rank(mine_p, rank ^ 1);
          const double M = fmax(v0.x, v1.x);
          const bool live = !signbit(v0.y) || !signbit(v1.y);
          const bool own_max = v0.x >= v1.x;
          const double eo = exp_nonpos((own_max ? v1.x : v0.x) - M);
          const double e0 = own_max ? 1.0 : eo, e1 = own_max ? eo : 1.0;
          const double t0 = (v0.x != -INFINITY) ? fabs(v0.y) * e0 : 0.0;
          const double t1 = (v1.x != -INFINITY) ? fabs(v1.y) * e1 : 0.0;
          const double stot = rank == 0 ? t0 + t1 : t1 + t0;
          f_o = ((v0.x == -INFINITY) ? 0.0 : e0) / stot;
          const bool valid = (int64_t)t * R + wrow0 + lane < a.n_rows;
          if (valid && !live) uflag = 1;
          if (valid && rank == 0) hacc += M + log(stot);
        }
        uint32_t zf[RG];
#pragma unroll
        for (int rg = 0; rg < RG; rg++) {
          const int src = slot * RG + rg;
          const double f = __shfl_sync(0xffffffffu, f_o, src);
          const bool valid = (int64_t)t * R + row_in_tile0 + rg < a.n_rows;
          double zz = 0.0;
          if (valid && kreal) {
            zz = ez[rg] * f;  // zeta / rowSums(zeta)
            hacc -= (zz > DBL_MIN) ? zz * lg[rg] : kTiny;  // R/mixdir_vi.R:181-184
          }
          zz = (zz > 0.0) ? zz : 0.0;
          {
            const unsigned long long q = __double2ull_rn(zz * kCsScale);
            cs_hi += q >> 32;
            cs_lo += q & 0xffffffffull;
          }
          zf[rg] = (uint32_t)(__double2ull_rd(
              fma(zz, kFixedScale32, (double)dither_mix(dbase + (uint32_t)rg * kDitherC1))) >> 32);
        }
        // the M-step warps of BOTH CTAs must have left the previous tile's responsibilities
        if (it > 0) mbar_wait_cluster(zfree, (uint32_t)((it - 1) & 1));
#pragma unroll
        for (int rg = 0; rg < RG; rg++) {
          z_s[(row_in_tile0 + rg) * 32 + kg] = zf[rg];
          if (!(a.ws_debug & 32)) z_peer[(row_in_tile0 + rg) * 32 + kg] = zf[rg];
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_local_cluster(zfull);
          if (a.ws_debug & 32)  // (32: timing experiment without the remote half of the zeta tile)
            mbar_arrive_local_cluster(zfull);
          else
            mbar_arrive_remote(zfull, (uint32_t)(rank ^ 1));
        }
      }
    } else if (warp == MW - 1) {
      // =========================== copy warp ===========================
      if (lane == 0) {
        int it = 0;
        for (int t = cluster_id; t < a.n_tiles; t += n_clusters, ++it) {
          const uint32_t par = (uint32_t)(it & 1);
          const int tn = t + n_clusters;
          if (tn >= a.n_tiles) break;
          mbar_wait(tile_free, par);
          mbar_expect_tx(&bar[0], tile_bytes);
          tma_load_1d(tile_s, a.codes + (size_t)tn * R * WPR, tile_bytes, &bar[0]);
          mbar_wait(seg_free, par);
          mbar_expect_tx(&bar[1], (uint32_t)a.blk_size[own]);
          tma_load_1d(seg_s, a.seg + (size_t)tn * a.TB + a.blk_off[own], (uint32_t)a.blk_size[own], &bar[1]);
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RM));
    // =========================== M-step warps ===========================
    const int mw = warp - MW;
    const int half = lane >> 4, c2 = lane & 15;  // two classes per lane, the half-warps on different rows (seg.cuh)
    const uint32_t zb2 = smem_u32(z_s) + c2 * 8;
    const uint32_t selA = half ? 0x4441u : 0x4440u, selB = half ? 0x4443u : 0x4442u;
    const uint32_t seg_addr = smem_u32(seg_s);
    int it = 0;
    for (int t = cluster_id; t < a.n_tiles; t += n_clusters, ++it) {
      const uint32_t par = (uint32_t)(it & 1);
      mbar_wait(&bar[1], par);
      mbar_wait_cluster(zfull, par);
      if (a.ws_debug & 8) {  // timing experiment: an M-step that only takes time (no instructions, no loads)
        const long long t0 = clock64();
        while (clock64() - t0 < 14000) __nanosleep(500);
      }
      for (int job = mw; job < ((a.ws_debug & 1) ? 0 : a.njobs[own]); job += MW) {
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
          const ulonglong2 old = *srow;
          unsigned long long sa = 0ull, sb = 0ull;
          int g = 0;
          for (; g + 2 <= n; g += 2) {
            const uint32_t w1 = lds_u32(p + 4);
            p += 8;
            const uint32_t wn = lds_u32(p);
            const uint2 z0 = lds_u32x2(zb2 + __byte_perm(w, 0u, selA) * ZS);
            const uint2 z1 = lds_u32x2(zb2 + __byte_perm(w, 0u, selB) * ZS);
            const uint2 z2 = lds_u32x2(zb2 + __byte_perm(w1, 0u, selA) * ZS);
            const uint2 z3 = lds_u32x2(zb2 + __byte_perm(w1, 0u, selB) * ZS);
            sa += z0.x; sb += z0.y; sa += z1.x; sb += z1.y;
            sa += z2.x; sb += z2.y; sa += z3.x; sb += z3.y;
            w = wn;
          }
          if (g < n) {
            p += 4;
            const uint32_t wn = lds_u32(p);
            const uint2 z0 = lds_u32x2(zb2 + __byte_perm(w, 0u, selA) * ZS);
            const uint2 z1 = lds_u32x2(zb2 + __byte_perm(w, 0u, selB) * ZS);
            sa += z0.x; sb += z0.y; sa += z1.x; sb += z1.y;
            w = wn;
          }
          sa += __shfl_xor_sync(0xffffffffu, sa, 16);
          sb += __shfl_xor_sync(0xffffffffu, sb, 16);
          if (half == 0) *srow = make_ulonglong2(old.x + sa, old.y + sb);
        }
      }
      // this warp has left the zeta tile (both CTAs' E warps wait for it) and the segment lists
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_local_cluster(zfree);
        if (a.ws_debug & 32)
          mbar_arrive_local_cluster(zfree);
        else
          mbar_arrive_remote(zfree, (uint32_t)(rank ^ 1));
        mbar_arrive_cta(seg_free);
      }
    }
  }
  __syncthreads();

  // ---- flush: integer adds into the global statistics; last CTA writes the tail ----
  {
    const int n = a.SC[own] * SW;
    const int* map = a.srow_to_stat + a.sc0[own];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long v = S_s[i];
      if (v) atomicAdd(&a.stats[(size_t)__ldg(map + i / SW) * KP + i % SW], v);
    }
#pragma unroll
    for (int o = KC; o < 32; o <<= 1) {
      cs_hi += __shfl_xor_sync(0xffffffffu, cs_hi, o);
      cs_lo += __shfl_xor_sync(0xffffffffu, cs_lo, o);
    }
    hacc = warp_sum(hacc);
    unsigned long long* redi_s = reinterpret_cast<unsigned long long*>(red_s);
    for (int word = 0; word < 2; word++) {
      if (warp < WE && slot == 0) redi_s[warp * KC + kc] = word ? cs_lo : cs_hi;
      __syncthreads();
      if (threadIdx.x < KC) {
        unsigned long long c = 0ull;
        for (int w = 0; w < WE; w++) c += redi_s[w * KC + threadIdx.x];
        if (c) atomicAdd(&a.cs_fixed[word * KP + rank * KC + threadIdx.x], c);
      }
      __syncthreads();
    }
    if (warp < WE && lane == 0) red_s[WE * KC + warp] = hacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double h = 0;
      for (int w = 0; w < WE; w++) h += red_s[WE * KC + w];
      a.Hpart[blockIdx.x] = a.accumulate ? a.Hpart[blockIdx.x] + h : h;
    }
    if (uflag) atomicExch(a.underflow, 1);
    __threadfence();
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
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
  cluster.sync();
}

}  // namespace mxd
