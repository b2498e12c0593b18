// exp(x) for the soft-max weights exp(logit - shift), x <= ~0: branch-free, so the compiler can
// interleave the eight rows a lane holds (libdevice's exp carries a branch for the out-of-range
// scaling, which serialises the unrolled rows: profiles/README.md, round 2).
//   n = rint(x / ln 2) by the 1.5 * 2^52 trick, r = x - n ln2 (Cody-Waite, two-part ln 2, exact
//   n * ln2_hi for |n| < 2^11), e^r by a degree-13 Taylor polynomial (|r| <= 0.3466: truncation
//   4e-18), scaling by an exponent-field add.  Maximum error against libm over [-708, 0.01]:
//   < 1 ulp (tests/test_host_logic.py::test_fast_exp_accuracy compiles this file for the host).
//   Results below 2^-1022 (x < -708) are flushed to 0: a weight that small relative to the row
//   maximum cannot change any sum (the reference's `rowSums(zeta) == 0` underflow test is made on
//   the unshifted logits, separately).  -inf and NaN give 0, and so does x >= 64 (not a soft-max
//   weight: the shift was -inf, i.e. every logit of the row is below the fp32 range -- such a row
//   raises the underflow error).
#pragma once

#if defined(__CUDACC__)
#define MXD_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define MXD_HD inline
#endif

namespace mxd {

MXD_HD double exp_nonpos(double x) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 1.4426950408889634074, kMagic);
  const double nf = t - kMagic;
  double r = fma(nf, -6.93147180369123816490e-01, x);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;           // 1/13!
  p = fma(p, r, 2.0876756987868100e-09);       // 1/12!
  p = fma(p, r, 2.5052108385441720e-08);       // 1/11!
  p = fma(p, r, 2.7557319223985890e-07);       // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
  p = fma(p, r, 2.4801587301587302e-05);       // 1/8!
  p = fma(p, r, 1.9841269841269841e-04);       // 1/7!
  p = fma(p, r, 1.3888888888888889e-03);       // 1/6!
  p = fma(p, r, 8.3333333333333332e-03);       // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
#if defined(__CUDA_ARCH__)
  const int n = __double2loint(t);
  const double v = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#else
  uint64_t tb, pb;
  std::memcpy(&tb, &t, 8);
  std::memcpy(&pb, &p, 8);
  const int32_t n = (int32_t)(uint32_t)tb;
  pb += (uint64_t)((int64_t)n << 52);
  double v;
  std::memcpy(&v, &pb, 8);
#endif
  return (x > -708.0 && x < 64.0) ? v : 0.0;
}

}  // namespace mxd
