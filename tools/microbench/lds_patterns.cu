// Microbenchmark 3: shared-memory wavefronts of the access patterns a "lanes = rows" E-step and a
// "sorted segment" M-step would use, next to the patterns k_em_units uses today.
// Question: when the 32 lanes of one LDS.64 / LDS.128 pick (with repeats) among the 16 entries of ONE
// 128-byte line, does the instruction cost one wavefront (bank array read once, broadcast) or is it
// bound by the 256/512 bytes it returns to the register file?
//
//   0  LDS.64  lanes contiguous, 256 B row, warp-uniform random row       (k_em_units gather, KC=32)
//   1  LDS.64  lane-random entry among 16 x 8 B of one 128 B line         (lanes = rows, C <= 16)
//   2  LDS.64  lane-random entry among  8 x 8 B of one  64 B half line    (lanes = rows, C <= 8)
//   3  LDS.128 lane-random entry among 16 x 16 B over 256 B (2 lines)     (2 classes per load, C <= 16)
//   4  LDS.128 lane-random entry among  8 x 16 B of one 128 B line        (2 classes per load, C <= 8)
//   5  LDS.32  lanes contiguous, 128 B row                                (fp32 / fixed-point zeta row)
//   6  LDS.64  half-warps read two different contiguous 128 B lines       (k_em_units gather, KC=16)
//   7  LDS.64  all lanes the same address
//   8  LDS.32  lane-random entry among 16 x 4 B of one 64 B half line
//   9  LDS.64  lane-random entry among 32 x 8 B over 256 B (2 lines)
//  10  LDS.128 lanes contiguous, 512 B                                    (zeta row, K=64 as fp64 pairs)
//  11  LDS.64  lanes contiguous 256 B + DADD into one accumulator chain   (segment sum, dependent adds)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 16; }

template <int MODE, int U>
__global__ void __launch_bounds__(1024, 1) k(int iters, double* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* tab = reinterpret_cast<double*>(sm);  // 64 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (double)(i % 97) * 1e-3;
  __syncthreads();
  uint32_t sl = 99991u * (threadIdx.x + 1) + blockIdx.x;   // per-lane stream
  uint32_t sw = 1234567u + warp * 7919u + blockIdx.x;      // warp-uniform stream
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
  uint32_t off[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t r = lcg(sl);
    switch (MODE) {
      case 0: case 11: off[u] = lane * 8; break;
      case 1: off[u] = (r & 15) * 8; break;
      case 2: off[u] = (r & 7) * 8; break;
      case 3: off[u] = (r & 15) * 16; break;
      case 4: off[u] = (r & 7) * 16; break;
      case 5: off[u] = lane * 4; break;
      case 6: off[u] = (lane & 15) * 8 + (lane >> 4) * 128 * 37; break;
      case 7: off[u] = 0; break;
      case 8: off[u] = (r & 15) * 4; break;
      case 9: off[u] = (r & 31) * 8; break;
      case 10: off[u] = lane * 16; break;
    }
  }
  double acc = 0, acc2 = 0;
  float facc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    // warp-uniform line / row: 512-byte granules inside the first 32 KB
    const uint32_t line = base + ((lcg(sw) & 63u) << 9);
    if constexpr (MODE == 3 || MODE == 4 || MODE == 10) {
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "r"(line + off[u]));
#pragma unroll
      for (int u = 0; u < U; u++) { acc += v[u].x; acc2 += v[u].y; }
    } else if constexpr (MODE == 5 || MODE == 8) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[u]) : "r"(line + off[u]));
#pragma unroll
      for (int u = 0; u < U; u++) facc += v[u];
    } else if constexpr (MODE == 11) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(line + off[u] + u * 256));
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];  // one dependent chain
    } else {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(line + off[u]));
#pragma unroll
      for (int u = 0; u < U; u += 2) { acc += v[u]; acc2 += v[u + 1]; }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (lane == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2 + facc;
}

template <int MODE>
void run(const char* name, double* out, long long* cyc) {
  const int NB = 148, iters = 4096, U = 8;
  static long long hc[NB * 32];
  cudaFuncSetAttribute(k<MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int nw : {8, 16, 32}) {
    for (int rep = 0; rep < 2; rep++) {
      k<MODE, U><<<NB, nw * 32, 65536>>>(iters, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < nw; w++) mx = hc[w] > mx ? hc[w] : mx;
    printf("mode %2d warps=%2d %-62s %6.3f cyc per warp instruction per SM\n", MODE, nw, name,
           (double)mx / ((double)iters * U * nw));
  }
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 32 * 8);
  run<0>("LDS.64 contiguous 256 B row", out, cyc);
  run<1>("LDS.64 lane-random among 16 x 8 B (one 128 B line)", out, cyc);
  run<2>("LDS.64 lane-random among 8 x 8 B (64 B)", out, cyc);
  run<3>("LDS.128 lane-random among 16 x 16 B (two lines)", out, cyc);
  run<4>("LDS.128 lane-random among 8 x 16 B (one line)", out, cyc);
  run<5>("LDS.32 contiguous 128 B row", out, cyc);
  run<6>("LDS.64 half-warps on two contiguous 128 B lines", out, cyc);
  run<7>("LDS.64 all lanes one address", out, cyc);
  run<8>("LDS.32 lane-random among 16 x 4 B (64 B)", out, cyc);
  run<9>("LDS.64 lane-random among 32 x 8 B (two lines)", out, cyc);
  run<10>("LDS.128 contiguous 512 B", out, cyc);
  run<11>("LDS.64 contiguous 256 B rows + one dependent DADD chain", out, cyc);
  return 0;
}
