// Microbenchmark 2: cost per (table gather + histogram read-modify-write) pair, 256 B rows
// (32 lanes x fp64), warp-uniform pseudo-random row indices generated in registers.
//   mode 0: gather from smem + RMW in smem, same warp (what k_em_fused does)        -> model 6 cyc
//   mode 1: gather from TMEM + RMW in smem, same warp
//   mode 2: warp-specialised: even warps gather from TMEM, odd warps RMW in smem
//   mode 3: gather from TMEM only     mode 4: gather from smem only     mode 5: RMW smem only
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return (s >> 20) & 255u; }

template <int U>
__global__ void __launch_bounds__(1024, 1) k(int mode, int iters, double* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint32_t tbase_s;
  double* tab = reinterpret_cast<double*>(sm);          // T: [256][32] doubles = 64 KB
  double* his = reinterpret_cast<double*>(sm + 65536);  // S: [256][32] doubles = 64 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) { tab[i] = (double)(i % 97) * 1e-3; his[i] = 0; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s + ((uint32_t)((warp & 3) * 32) << 16);
  if (warp < 4) {
    for (int r = 0; r < 256; r++) {
      const double v = (double)((r * 32 + lane) % 89) * 1e-3;
      tm_st2(tb + 2 * r, __double2loint(v), __double2hiint(v));
    }
    tm_wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  double acc = 0;
  uint32_t s1 = 1234567u + warp * 7919u + blockIdx.x, s2 = 7654321u + warp * 104729u + blockIdx.x;
  const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tab) + lane * 8;
  const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(his) + lane * 8;
  const bool do_tm = mode == 1 || mode == 3 || (mode == 2 && !(warp & 1));
  const bool do_lds = mode == 0 || mode == 4;
  const bool do_rmw = mode == 0 || mode == 1 || mode == 5 || (mode == 2 && (warp & 1));
  const double z = 1e-9 * lane;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (do_tm) {
      uint32_t lo[U], hi[U];
#pragma unroll
      for (int u = 0; u < U; u++) tm_ld2(tb + 2 * lcg(s1), lo[u], hi[u]);
      tm_wait_ld();
#pragma unroll
      for (int u = 0; u < U; u++) acc += __hiloint2double(hi[u], lo[u]);
    }
    if (do_lds) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(tbase + lcg(s1) * 256));
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];
    }
    if (do_rmw) {
      uint32_t ad[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        ad[u] = hbase + lcg(s2) * 256;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(ad[u]) : "memory");
      }
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad[u]), "d"(v[u] + z) : "memory");
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (lane == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + his[threadIdx.x];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

int main() {
  const int NB = 148;
  double* out;
  long long* cyc;
  cudaMalloc(&out, NB * 1024 * 8);
  cudaMalloc(&cyc, NB * 32 * 8);
  cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  long long hc[NB * 32];
  const int iters = 4096;
  const char* names[] = {"smem gather + smem RMW (same warp)", "TMEM gather + smem RMW (same warp)",
                         "TMEM gather | smem RMW (specialised)", "TMEM gather only", "smem gather only",
                         "smem RMW only"};
  for (int U : {8, 4})
    for (int nw : {8, 16, 32})
      for (int mode = 0; mode < 6; mode++) {
        for (int rep = 0; rep < 2; rep++) {
          if (U == 8) k<8><<<NB, nw * 32, 131072>>>(mode, iters, out, cyc);
          else k<4><<<NB, nw * 32, 131072>>>(mode, iters, out, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s: %s\n", names[mode], cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nw; w++) mx = hc[w] > mx ? hc[w] : mx;
        // "pairs" processed by the CTA: mode 2 splits the warps between the two halves of a pair
        const double pairs = (double)iters * U * (mode == 2 ? nw / 2 : nw);
        printf("U=%d warps=%2d %-40s %7.3f cyc per op(pair) per SM\n", U, nw, names[mode], mx / pairs);
      }
  return 0;
}
