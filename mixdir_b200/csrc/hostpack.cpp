// Host side of mixdir_b200_create_from_r_matrix: the reference's X -- an fp64 column-major matrix
// with values 1..C_j and NA (/root/reference/R/mixdir.R:117-118) -- packed to 1-byte row-major codes
// (0 = NA) by AVX2 code.  The scalar packer in engine.cu spends ~8 instructions per cell on 600 M
// cells at the S4 shape; here 8 rows x 4 columns are converted, checked and transposed in registers
// (one 64-bit store per row and group of 8 columns).  The columns of X lie n_rows * 8 bytes apart: a
// thread walks a block of `rows_per_block` rows column group by column group, so every column is read in
// runs long enough for the hardware prefetchers (16 KB at 2048 rows) while the block's output
// (2048 x J bytes) stays in L2.  Measured on a 8-core host with 12.5 GB/s of read bandwidth, 2M x 60:
// 256-row blocks and 4-column groups 4.8 GB/s, 2048-row blocks 8.9 GB/s, with 8-column groups 10.6 GB/s.
// Built by g++ (see Makefile); engine.cu calls it only when the CPU reports AVX2, and falls back to
// its scalar loop otherwise.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern "C" int mxd_have_avx2() { return __builtin_cpu_supports("avx2") ? 1 : 0; }

// 8 doubles of one column -> 8 x int32 codes (0 for NA and for invalid cells); *bad |= invalid seen
__attribute__((target("avx2"))) static inline __m256i codes8(const double* col, __m128i lim, __m128i* bad) {
  const __m256d v0 = _mm256_loadu_pd(col), v1 = _mm256_loadu_pd(col + 4);
  const __m128i x0 = _mm256_cvttpd_epi32(v0), x1 = _mm256_cvttpd_epi32(v1);
  // exact positive integer not above the limit?  (NaN compares false; cvtt of NaN / huge = INT_MIN)
  const __m256d e0 = _mm256_cmp_pd(_mm256_cvtepi32_pd(x0), v0, _CMP_EQ_OQ);
  const __m256d e1 = _mm256_cmp_pd(_mm256_cvtepi32_pd(x1), v1, _CMP_EQ_OQ);
  const __m256d n0 = _mm256_cmp_pd(v0, v0, _CMP_UNORD_Q), n1 = _mm256_cmp_pd(v1, v1, _CMP_UNORD_Q);
  // 64-bit masks -> 32-bit masks (take the low half of every lane)
  const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
  const __m128i eq0 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_castpd_si256(e0), idx));
  const __m128i eq1 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_castpd_si256(e1), idx));
  const __m128i na0 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_castpd_si256(n0), idx));
  const __m128i na1 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_castpd_si256(n1), idx));
  const __m128i zero = _mm_setzero_si128();
  const __m128i ok0 = _mm_and_si128(eq0, _mm_andnot_si128(_mm_cmpgt_epi32(x0, lim), _mm_cmpgt_epi32(x0, zero)));
  const __m128i ok1 = _mm_and_si128(eq1, _mm_andnot_si128(_mm_cmpgt_epi32(x1, lim), _mm_cmpgt_epi32(x1, zero)));
  *bad = _mm_or_si128(*bad, _mm_or_si128(_mm_andnot_si128(_mm_or_si128(ok0, na0), _mm_set1_epi32(-1)),
                                         _mm_andnot_si128(_mm_or_si128(ok1, na1), _mm_set1_epi32(-1))));
  return _mm256_set_m128i(_mm_and_si128(x1, ok1), _mm_and_si128(x0, ok0));
}

// rows [r0, r1) of the column-major X (leading dimension ld) -> dst (row r0 first), 1 byte per cell,
// row_bytes per row.  Returns 1 if a cell was neither NA nor an integer in 1..limit.
extern "C" __attribute__((target("avx2"))) int mxd_pack_u8_avx2(const double* X, int64_t ld, int64_t r0,
                                                                 int64_t r1, int J, int row_bytes,
                                                                 unsigned char* dst, int limit) {
  const __m128i lim = _mm_set1_epi32(limit);
  __m128i bad = _mm_setzero_si128();
  int sbad = 0;
  const int J4 = J & ~3;
  // rows per block (MIXDIR_B200_PACK_ROWS overrides, multiple of 8)
  static const int64_t RB = [] {
    const char* e = getenv("MIXDIR_B200_PACK_ROWS");
    const int64_t v = e ? atoll(e) : 0;
    return v >= 8 ? (v & ~(int64_t)7) : (int64_t)2048;
  }();
  for (int64_t b0 = r0; b0 < r1; b0 += RB) {
    const int64_t b1 = b0 + RB < r1 ? b0 + RB : r1;
    const int64_t n8 = b0 + ((b1 - b0) & ~(int64_t)7);
    int j = 0;
    for (; j + 8 <= J4; j += 8) {  // 8 columns x 8 rows: one 64-bit store per row
      const double *c0 = X + (int64_t)j * ld, *c1 = c0 + ld, *c2 = c1 + ld, *c3 = c2 + ld;
      const double *c4 = c3 + ld, *c5 = c4 + ld, *c6 = c5 + ld, *c7 = c6 + ld;
      for (int64_t i = b0; i < n8; i += 8) {
        const __m256i a = codes8(c0 + i, lim, &bad), b = codes8(c1 + i, lim, &bad),
                      c = codes8(c2 + i, lim, &bad), d = codes8(c3 + i, lim, &bad);
        const __m256i w0 = _mm256_or_si256(_mm256_or_si256(a, _mm256_slli_epi32(b, 8)),
                                           _mm256_or_si256(_mm256_slli_epi32(c, 16), _mm256_slli_epi32(d, 24)));
        const __m256i e = codes8(c4 + i, lim, &bad), f = codes8(c5 + i, lim, &bad),
                      g = codes8(c6 + i, lim, &bad), h = codes8(c7 + i, lim, &bad);
        const __m256i w1 = _mm256_or_si256(_mm256_or_si256(e, _mm256_slli_epi32(f, 8)),
                                           _mm256_or_si256(_mm256_slli_epi32(g, 16), _mm256_slli_epi32(h, 24)));
        // (columns 0-3, columns 4-7) of a row side by side: rows 0,1,4,5 | rows 2,3,6,7
        alignas(32) uint64_t t[8];
        _mm256_store_si256((__m256i*)t, _mm256_unpacklo_epi32(w0, w1));
        _mm256_store_si256((__m256i*)(t + 4), _mm256_unpackhi_epi32(w0, w1));
        unsigned char* o = dst + (i - r0) * row_bytes + j;
        static const int rowmap[8] = {0, 1, 4, 5, 2, 3, 6, 7};
        for (int r = 0; r < 8; r++) memcpy(o + (int64_t)rowmap[r] * row_bytes, &t[r], 8);
      }
    }
    for (; j < J4; j += 4) {
      const double *c0 = X + (int64_t)j * ld, *c1 = c0 + ld, *c2 = c1 + ld, *c3 = c2 + ld;
      for (int64_t i = b0; i < n8; i += 8) {
        const __m256i a = codes8(c0 + i, lim, &bad), b = codes8(c1 + i, lim, &bad),
                      c = codes8(c2 + i, lim, &bad), d = codes8(c3 + i, lim, &bad);
        const __m256i w = _mm256_or_si256(_mm256_or_si256(a, _mm256_slli_epi32(b, 8)),
                                          _mm256_or_si256(_mm256_slli_epi32(c, 16), _mm256_slli_epi32(d, 24)));
        alignas(32) uint32_t t[8];
        _mm256_store_si256((__m256i*)t, w);
        unsigned char* o = dst + (i - r0) * row_bytes + j;
        for (int r = 0; r < 8; r++) *(uint32_t*)(o + (int64_t)r * row_bytes) = t[r];  // (j % 4 == 0)
      }
    }
    // the rest: last columns, last rows of the block
    for (int j = 0; j < J; j++) {
      const double* col = X + (int64_t)j * ld;
      for (int64_t i = (j < J4 ? n8 : b0); i < b1; i++) {
        const double v = col[i];
        int x = 0;
        if (v == v) {
          x = (int)v;
          if (v < 1.0 || v != (double)x || x > limit) {
            sbad = 1;
            x = 0;
          }
        }
        dst[(i - r0) * row_bytes + j] = (unsigned char)x;
      }
    }
  }
  return (sbad || !_mm_testz_si128(bad, bad)) ? 1 : 0;
}
