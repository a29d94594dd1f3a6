// Feature preparation: one pass that replaces F.normalize (utils/metrics.py:114), the squared-norm
// terms of utils/metrics.py:10-11 / utils/reranking.py:38-39 and produces the tensor-core operand
// planes (TF32 hi/lo split for the fp32-accurate 3xTF32 GEMM, bf16 for the bf16 GEMM).
// HBM-bound: reads x once (the second sweep of a row hits L1/L2), writes each requested plane once.
#include "common.cuh"

namespace mpreid {

static constexpr int kPrepThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (sh[0] + sh[1]) + (sh[2] + sh[3]);
  __syncthreads();
  return t;
}

__device__ __forceinline__ float block_max_128(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
  __syncthreads();
  return t;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(kPrepThreads)
k_prep_rows(const float* __restrict__ x, int64_t rows, int D, int64_t ld_x, int normalize,
            float* __restrict__ xn, int64_t ld_xn, float* __restrict__ sqnorm, float* __restrict__ norm,
            float* __restrict__ hi, float* __restrict__ lo, __nv_bfloat16* __restrict__ bf,
            __half* __restrict__ h_hi, __half* __restrict__ h_lo, float* __restrict__ h_scale_inv, int Dp) {
  __shared__ float sh[4];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * ld_x;
    float den = 1.0f, hscale = 1.0f;
    if (normalize || h_hi) {
      float s = 0.f, mx = 0.f;
      for (int c = threadIdx.x; c < D; c += kPrepThreads) { float v = xr[c]; s = fmaf(v, v, s); mx = fmaxf(mx, fabsf(v)); }
      if (normalize) {
        s = block_sum_128(s, sh);
        den = fmaxf(sqrtf(s), 1e-12f);  // F.normalize: x / max(||x||_2, eps)
      }
      if (h_hi) {
        mx = block_max_128(mx, sh);
        if (normalize) mx = mx / den;
        // power-of-two row scale: max|xn| * 2^s in [512, 1024) keeps hi AND lo in fp16's normal range
        int e = 0;
        if (mx > 0.f && mx < INFINITY) { (void)frexpf(mx, &e); hscale = ldexpf(1.0f, 10 - e); }
        if (threadIdx.x == 0) h_scale_inv[r] = 1.0f / hscale;
      }
    }
    float s2 = 0.f;
    for (int c = threadIdx.x; c < Dp; c += kPrepThreads) {
      float v = 0.f;
      if (c < D) {
        v = xr[c];
        if (normalize) v = v / den;
        s2 = fmaf(v, v, s2);
        if (xn) xn[r * ld_xn + c] = v;
      }
      if (hi) {
        const float h = to_tf32(v);
        hi[r * (int64_t)Dp + c] = h;
        lo[r * (int64_t)Dp + c] = to_tf32(v - h);
      }
      if (bf) bf[r * (int64_t)Dp + c] = __float2bfloat16_rn(v);
      if (h_hi) {
        const float xs = v * hscale;               // exact (power of two)
        const __half h = __float2half_rn(xs);
        h_hi[r * (int64_t)Dp + c] = h;
        h_lo[r * (int64_t)Dp + c] = __float2half_rn(xs - __half2float(h));
      }
    }
    s2 = block_sum_128(s2, sh);
    if (threadIdx.x == 0) {
      if (sqnorm) sqnorm[r] = s2;
      if (norm) norm[r] = sqrtf(s2);
    }
  }
}

}  // namespace mpreid

using namespace mpreid;

extern "C" int mpreid_prep_rows(const float* x, int64_t rows, int64_t D, int64_t ld_x, int normalize,
                                float* xn, int64_t ld_xn, float* sqnorm, float* norm,
                                float* hi, float* lo, uint16_t* bf,
                                uint16_t* h_hi, uint16_t* h_lo, float* h_scale_inv, int64_t Dp, void* stream) {
  MPREID_REQUIRE(x && rows > 0 && D > 0 && ld_x >= D, "prep_rows: bad input (rows=%lld D=%lld ld=%lld)",
                 (long long)rows, (long long)D, (long long)ld_x);
  MPREID_REQUIRE((hi == nullptr) == (lo == nullptr), "prep_rows: hi and lo planes go together");
  MPREID_REQUIRE(!xn || ld_xn >= D, "prep_rows: ld_xn < D");
  MPREID_REQUIRE((h_hi == nullptr) == (h_lo == nullptr) && (h_hi == nullptr) == (h_scale_inv == nullptr),
                 "prep_rows: h_hi, h_lo and h_scale_inv go together");
  const bool planes = hi || bf || h_hi;
  MPREID_REQUIRE(!planes || (Dp >= D && Dp % 32 == 0), "prep_rows: Dp must be a multiple of 32 and >= D");
  MPREID_REQUIRE(!(bf || h_hi) || Dp % 64 == 0, "prep_rows: 16-bit planes need Dp to be a multiple of 64");
  MPREID_REQUIRE(D < (1 << 30), "prep_rows: D too large");
  const int dp = planes ? (int)Dp : (int)D;
  int64_t grid = rows < 148 * 16 ? rows : 148 * 16;
  k_prep_rows<<<(unsigned)grid, kPrepThreads, 0, (cudaStream_t)stream>>>(
      x, rows, (int)D, ld_x, normalize, xn, ld_xn, sqnorm, norm, hi, lo, (__nv_bfloat16*)bf,
      (__half*)h_hi, (__half*)h_lo, h_scale_inv, dp);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
