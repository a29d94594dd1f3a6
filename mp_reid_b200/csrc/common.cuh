// Shared helpers for the mpreid_b200 kernels (host+device scalar code lives here so that the
// CPU unit tests can exercise exactly the arithmetic the kernels run).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/mpreid_b200.h"

#define HD __host__ __device__ __forceinline__

namespace mpreid {

void set_error(const char* fmt, ...);

#define MPREID_CUDA_CHECK(expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::mpreid::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MPREID_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define MPREID_REQUIRE(cond, ...)                                                            \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::mpreid::set_error(__VA_ARGS__);                                                      \
      return MPREID_ERR_INVALID;                                                             \
    }                                                                                        \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count_of_current_device();

// ---------------------------------------------------------------------------------------------
// Total order on fp32 that reproduces numpy's sort order: ascending value, -0.0 == +0.0, every NaN
// last (all NaNs tie).  Ties are then broken by the gallery index (== kind='stable').
// ---------------------------------------------------------------------------------------------
HD uint32_t f32_bits(float v) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(v);
#else
  union { float f; uint32_t u; } c; c.f = v; return c.u;
#endif
}
HD float bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
HD uint32_t order_key(float v) {
  uint32_t u = f32_bits(v + 0.0f);                // -0.0 -> +0.0
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
HD float order_key_inv(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return bits_f32(u);
}
HD uint64_t make_key(float v, uint32_t idx) { return ((uint64_t)order_key(v) << 32) | idx; }

// ---------------------------------------------------------------------------------------------
// Average precision exactly as numpy evaluates utils/metrics.py:73-79:
//   tmp = cumsum(match) / arange(1..n) * match ; AP = tmp.sum() / num_rel      (float64)
// tmp is zero except at the positions of the correct matches, so the sum is a function of the
// sorted 0-based positions pos[0..m) only -- but numpy adds with its pairwise scheme
// (blocks <= 128 with 8 strided accumulators, recursive halving rounded to a multiple of 8), and
// float64 addition is not associative, so the same tree is walked here over the sparse terms.
// term(i) = double(i+1) / double(pos[i]+1).
// ---------------------------------------------------------------------------------------------
// `rank` holds the 1-based ranks of the correct matches, ascending; position = rank - 1.
HD double ap_term(const int32_t* rank, int i) { return (double)(i + 1) / (double)rank[i]; }

HD int lower_bound_pos(const int32_t* rank, int a, int b, int64_t x) {  // first i in [a,b) with rank[i]-1 >= x
  while (a < b) {
    int mid = (a + b) >> 1;
    if ((int64_t)rank[mid] - 1 < x) a = mid + 1; else b = mid;
  }
  return a;
}

HD double pairwise_leaf(const int32_t* rank, int64_t lo, int64_t n, int a, int b) {
  if (n < 8) {
    double res = 0.0;
    for (int i = a; i < b; ++i) res = res + ap_term(rank, i);
    return res;
  }
  double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
  const int64_t body_end = lo + (n - (n % 8));
  int i = a;
  for (; i < b && (int64_t)rank[i] - 1 < body_end; ++i) {
    const double t = ap_term(rank, i);
    switch (((int64_t)rank[i] - 1 - lo) & 7) {
      case 0: r0 += t; break; case 1: r1 += t; break; case 2: r2 += t; break; case 3: r3 += t; break;
      case 4: r4 += t; break; case 5: r5 += t; break; case 6: r6 += t; break; default: r7 += t; break;
    }
  }
  double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (; i < b; ++i) res = res + ap_term(rank, i);
  return res;
}

// sum of the length-n vector whose non-zeros sit at positions rank[0..m)-1 (ascending), numpy pairwise order.
// Empty subtrees contribute +0.0 and x + 0.0 == x, so the value is the non-empty leaves (blocks <= 128)
// combined in the bracketing the halving tree induces: two neighbouring leaves join at their lowest common
// tree node, deeper joins first.  One descent per non-empty leaf finds both the leaf and the depth at
// which it parts from its predecessor; a small stack holds the partial sums still waiting for their sibling.
HD double pairwise_sparse_sum(const int32_t* rank, int m, int64_t n) {
  if (m <= 0) return 0.0;
  double val[64];
  int junc[64];          // junc[j] = depth of the tree node that joins stack entries j-1 and j
  int sp = 0, a = 0;
  int64_t prev = -1;     // a position inside the previous non-empty leaf
  while (a < m) {
    const int64_t p = (int64_t)rank[a] - 1;
    int64_t lo = 0, len = n;
    int depth = 0, d = -1;
    bool together = prev >= 0;
    while (len > 128) {
      int64_t n2 = len / 2; n2 -= n2 % 8;
      const bool right = p >= lo + n2;
      if (together && (prev >= lo + n2) != right) { d = depth; together = false; }
      if (right) { lo += n2; len -= n2; } else { len = n2; }
      ++depth;
    }
    int b = a + 1;
    while (b < m && (int64_t)rank[b] - 1 < lo + len) ++b;
    const double v = pairwise_leaf(rank, lo, len, a, b);
    while (sp >= 2 && junc[sp - 1] > d) { val[sp - 2] = val[sp - 2] + val[sp - 1]; --sp; }
    val[sp] = v; junc[sp] = d; ++sp;
    prev = p; a = b;
  }
  while (sp >= 2) { val[sp - 2] = val[sp - 2] + val[sp - 1]; --sp; }
  return val[0];
}

// Fused top-k mode of the symmetric all-pairs launch (re-ranking, utils/reranking.py:46-48 without the N x N matrix):
// instead of storing a tile and its transpose, the epilogue appends every element that is not above the per-sample
// threshold to the candidate list of its row -- and, for mirrored tiles, of its column -- and stores only the
// [Q, G] query-to-gallery block that the lambda blend needs (:95).
struct TopkFuse {
  const float* thr;              // [N] raw-domain threshold of sample i (as a row and, by symmetry, as a column)
  unsigned long long* cand;      // [N, cap] entries (column << 32 | fp32 bits), unordered
  int* cand_cnt;                 // [N] entries appended (keeps counting past cap: overflow is detected downstream)
  int cap;
  int keep_rows;                 // Q: rows < Q ...
  int keep_col0;                 // ... and columns >= keep_col0 (= Q rounded down to 32) are stored at out[row, col - keep_col0]
  int own_mod, own_rank;         // multi-GPU: this launch contracts only the tiles of the 256-row blocks p with p % own_mod == own_rank
};

// np.around(k1 / 2) -- round half to even (utils/reranking.py:60)
HD int round_half_even_div2(int k1) {
  int h = k1 / 2;
  if (k1 & 1) return (h & 1) ? h + 1 : h;  // x.5 -> nearest even
  return h;
}

HD uint32_t next_pow2_u32(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

HD uint64_t mix64(uint64_t x) {  // splitmix64 finaliser, hashes pids into the label table
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

}  // namespace mpreid
