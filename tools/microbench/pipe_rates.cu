// Microbenchmark 4: issue rates of the instruction classes the E+M kernels are made of, per SM, and
// whether two classes overlap when different warps of one SM run them (the question behind k_em_seg_ws:
// why do the fp64 E-step and the integer M-step not overlap on separate warps?).
//   mode 0  DADD   8 independent chains per thread
//   mode 1  DFMA   8 independent chains per thread
//   mode 2  IADD3  32-bit, 8 chains
//   mode 3  64-bit integer add (IADD3 + IADD3.X or IADD.64), 8 chains
//   mode 4  IMAD   32-bit, 8 chains
//   mode 5  even warps DADD, odd warps 32-bit IADD3            (overlap?)
//   mode 6  even warps DADD, odd warps LDS.32 contiguous rows  (overlap?)
//   mode 7  LDS.32 contiguous 128 B rows
//   mode 8  FADD   fp32, 8 chains
// Reported: SM cycles per warp instruction of the class (per SM), for 8 / 16 / 32 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, double* out, long long* cyc) {
  __shared__ __align__(128) uint32_t tab[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = i * 2654435761u;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double d[8];
  unsigned u[8];
  unsigned long long q[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i] = threadIdx.x * 1e-3 + i; u[i] = threadIdx.x + i; q[i] = threadIdx.x * 77ull + i; f[i] = threadIdx.x + i; }
  const double da = 1.0000001, db = 1e-9;
  const unsigned ua = blockIdx.x * 3 + 1;
  const unsigned long long qa = 0x100000001ull * (blockIdx.x + 1);
  const bool fp = (MODE == 5 || MODE == 6) ? ((warp >> 2) & 1) == 0 : true;  // every scheduler gets both kinds
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4;
  long long t0 = clock64();
  if (MODE == 5 || MODE == 6) {  // two different loops, chosen per warp (no predicated-off instructions)
    if (fp) {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int i = 0; i < 8; i++) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
    } else if (MODE == 5) {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(ua));
    } else {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int i = 0; i < 8; i++) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + ((u[i] & 63u) << 7)));
            u[i] ^= v;
          }
    }
  } else
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0 || ((MODE == 5 || MODE == 6) && fp)) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
        if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        if (MODE == 2 || (MODE == 5 && !fp)) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(ua));
        if (MODE == 3) asm volatile("add.u64 %0, %0, %1;" : "+l"(q[i]) : "l"(qa));
        if (MODE == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ua));
        if (MODE == 7 || (MODE == 6 && !fp)) {
          uint32_t v;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + ((u[i] & 63u) << 7)));
          u[i] ^= v;
        }
        if (MODE == 8) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0001f));
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i] + u[i] + (double)q[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  // the slowest warp of the block (in the mixed modes the two kinds of warps may finish apart)
  atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
  if (lane == 0 && (MODE == 5 || MODE == 6)) atomicMax((unsigned long long*)&cyc[148 + blockIdx.x * 2 + (fp ? 0 : 1)], (unsigned long long)(t1 - t0));
}

template <int MODE>
void run(const char* name, double per_iter_class_instr /* warp instructions of the class per warp and iteration */,
         double frac_warps) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 3 * 8);
  for (int warps : {8, 16, 32}) {
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(10, out, cyc);
    cudaMemset(cyc, 0, 148 * 3 * 8);
    k<MODE><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 3];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, a0 = 0, a1 = 0;
    for (int i = 0; i < 148; i++) { avg += h[i]; a0 += h[148 + 2 * i]; a1 += h[148 + 2 * i + 1]; }
    avg /= 148; a0 /= 148; a1 /= 148;
    const double n = (double)iters * per_iter_class_instr * warps * frac_warps;
    if (MODE == 5 || MODE == 6)
      printf("%-58s warps=%2d  fp64 warps done after %.3f, the other warps after %.3f cycles per instruction of their kind per SM\n", name, warps, a0 / n, a1 / n);
    else
      printf("%-58s warps=%2d  %.3f cycles per warp instruction per SM\n", name, warps, avg / n);
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("0 DADD", 32, 1);
  run<1>("1 DFMA", 32, 1);
  run<2>("2 IADD3 32-bit (+ LOP)", 32, 1);
  run<3>("3 64-bit integer add (+ LOP)", 32, 1);
  run<4>("4 IMAD 32-bit", 32, 1);
  run<8>("8 FADD", 32, 1);
  run<7>("7 LDS.32 contiguous 128 B (+ LOP, SHL, IADD)", 32, 1);
  run<5>("5 half the warps DADD [cycles per DADD], other half IADD3", 32, 0.5);
  run<6>("6 half the warps DADD [cycles per DADD], other half LDS.32", 32, 0.5);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
