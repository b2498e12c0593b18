// Microbenchmark: can tensor memory (TMEM) serve as a table cache for warp-uniform row gathers?
// Measures cycles per 256-byte row read (32 lanes x fp64) from TMEM (tcgen05.ld.32x32b.x2) and
// from shared memory (LDS.64), alone and concurrently, with data-dependent row indices.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_gather tmem_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode bit0: TMEM gathers by warps [0, ntm); bit1: smem RMW by warps [ntm, ntm+nsm)
// smem_kind 0: gather only (LDS.64), 1: read-modify-write (LDS.64 + DADD + STS.64)
template <int U>
__global__ void __launch_bounds__(1024, 1) k(int ntm, int nsm, int smem_kind, int iters, const uint32_t* idx,
                                              double* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint32_t tbase_s;
  double* tab = reinterpret_cast<double*>(sm);  // [256 rows][32] doubles = 64 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = (double)(i % 97) * 1e-3;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s + ((uint32_t)((warp & 3) * 32) << 16);
  // fill this warp's quadrant (warps 0..3 do it): 256 rows x 2 columns
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
  const uint32_t* my = idx + (size_t)warp * 4096;
  long long t0 = clock64();
  if (warp < ntm) {
    for (int it = 0; it < iters; it++) {
      uint32_t lo[U], hi[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t r = __ldg(&my[(it * U + u) & 4095]) & 255u;  // warp-uniform, data dependent
        tm_ld2(tb + 2 * r, lo[u], hi[u]);
      }
      tm_wait_ld();
#pragma unroll
      for (int u = 0; u < U; u++) acc += __hiloint2double(hi[u], lo[u]);
    }
  } else if (warp < ntm + nsm) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + lane * 8;
    for (int it = 0; it < iters; it++) {
      uint32_t ad[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t r = __ldg(&my[(it * U + u) & 4095]) & 255u;
        ad[u] = base + r * 256;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(ad[u]) : "memory");
      }
      if (smem_kind == 1) {
#pragma unroll
        for (int u = 0; u < U; u++) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad[u]), "d"(v[u] + 1e-9) : "memory");
      } else {
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u];
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (lane == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

int main() {
  const int NB = 148, NT = 1024;
  uint32_t* idx;
  double* out;
  long long* cyc;
  cudaMalloc(&idx, 32 * 4096 * 4);
  cudaMalloc(&out, NB * NT * 8);
  cudaMalloc(&cyc, NB * 32 * 8);
  uint32_t* h = new uint32_t[32 * 4096];
  uint64_t s = 12345;
  for (int i = 0; i < 32 * 4096; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; h[i] = (uint32_t)(s >> 33); }
  cudaMemcpy(idx, h, 32 * 4096 * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long hc[NB * 32];
  const int iters = 4096, U = 8;
  struct Cfg { int ntm, nsm, kind; const char* name; };
  Cfg cfgs[] = {{4, 0, 0, "tmem 4w"}, {8, 0, 0, "tmem 8w"}, {16, 0, 0, "tmem 16w"}, {1, 0, 0, "tmem 1w"},
                {0, 4, 0, "lds 4w"}, {0, 8, 0, "lds 8w"}, {0, 16, 0, "lds 16w"},
                {0, 8, 1, "rmw 8w"}, {0, 16, 1, "rmw 16w"},
                {8, 8, 1, "tmem 8w + rmw 8w"}, {16, 16, 1, "tmem 16w + rmw 16w"}, {8, 8, 0, "tmem 8w + lds 8w"},
                {4, 12, 1, "tmem 4w + rmw 12w"}, {16, 8, 1, "tmem 16w + rmw 8w"}};
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; rep++) {
      k<8><<<NB, NT, 65536>>>(c.ntm, c.nsm, c.kind, iters, idx, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mt = 0, ms = 0;
    for (int w = 0; w < c.ntm; w++) mt = hc[w] > mt ? hc[w] : mt;
    for (int w = c.ntm; w < c.ntm + c.nsm; w++) ms = hc[w] > ms ? hc[w] : ms;
    const double nt = (double)iters * U * c.ntm, ns = (double)iters * U * c.nsm;
    printf("%-22s tmem: %8.3f cyc per 256B row/SM (%6.1f B/cyc)   smem: %8.3f cyc per op/SM\n", c.name,
           c.ntm ? mt / nt : 0.0, c.ntm ? 256.0 * nt / mt : 0.0, c.nsm ? ms / ns : 0.0);
  }
  return 0;
}
