// Feature preparation: one pass that replaces F.normalize (utils/metrics.py:114), the squared-norm
// terms of utils/metrics.py:10-11 / utils/reranking.py:38-39 and produces the tensor-core operand
// planes (TF32 hi/lo split for the fp32-accurate 3xTF32 GEMM, bf16 for the bf16 GEMM).
// HBM-bound: reads x once (the second sweep of a row hits L1/L2), writes each requested plane once.
#include <cstdlib>
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

// One CTA per row (grid-stride).  Sweep 1: sum of squares + max |x| (only if normalising / scaling),
// sweep 2 (the row is in L1/L2 by then): normalise, accumulate the squared norm of the OUTPUT row (the
// reference takes the norms of the normalised features, utils/metrics.py:10-11) and emit the planes.
// VEC: four consecutive columns per thread with 128-bit loads / stores (needs 16-byte aligned rows).
template <bool VEC>
__global__ void __launch_bounds__(kPrepThreads)
k_prep_rows(const float* __restrict__ x, int64_t rows, int D, int64_t ld_x, int normalize,
            float* __restrict__ xn, int64_t ld_xn, float* __restrict__ sqnorm, float* __restrict__ norm,
            float* __restrict__ hi, float* __restrict__ lo, __nv_bfloat16* __restrict__ bf,
            __half* __restrict__ h_hi, __half* __restrict__ h_lo, float* __restrict__ h_scale_inv, int Dp) {
  __shared__ float sh[4];
  constexpr int W = VEC ? 4 : 1;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * ld_x;
    float den = 1.0f, hscale = 1.0f;
    if (normalize || h_hi) {
      float s = 0.f, mx = 0.f;
      for (int c = threadIdx.x * W; c < D; c += kPrepThreads * W) {
        float v[W];
        if (VEC) { const float4 t = *reinterpret_cast<const float4*>(xr + c); v[0] = t.x; v[W > 1 ? 1 : 0] = t.y; v[W > 2 ? 2 : 0] = t.z; v[W > 3 ? 3 : 0] = t.w; }
        else v[0] = xr[c];
#pragma unroll
        for (int j = 0; j < W; ++j) { s = fmaf(v[j], v[j], s); mx = fmaxf(mx, fabsf(v[j])); }
      }
      if (normalize) {
        s = block_sum_128(s, sh);
        den = fmaxf(sqrtf(s), 1e-12f);  // F.normalize: x / max(||x||_2, eps)
      }
      if (h_hi) {
        mx = block_max_128(mx, sh);
        if (normalize) mx = mx / den;
        // power-of-two row scale: max|xn| * 2^s in [512, 1024) keeps hi AND lo in fp16's normal range
        // (the exponent is clamped so that both 2^s and 2^-s stay normal fp32 numbers: rows below 2^-116 keep 2^126)
        int e = 0;
        if (mx > 0.f && mx < INFINITY) { (void)frexpf(mx, &e); hscale = ldexpf(1.0f, min(10 - e, 126)); }
        if (threadIdx.x == 0) h_scale_inv[r] = 1.0f / hscale;
      }
    }
    float s2 = 0.f;
    for (int c = threadIdx.x * W; c < Dp; c += kPrepThreads * W) {
      float v[W];
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = 0.f;
      if (c < D) {   // D % 4 == 0 in the vector path, so a group is entirely inside or entirely padding
        if (VEC) { const float4 t = *reinterpret_cast<const float4*>(xr + c); v[0] = t.x; v[W > 1 ? 1 : 0] = t.y; v[W > 2 ? 2 : 0] = t.z; v[W > 3 ? 3 : 0] = t.w; }
        else v[0] = xr[c];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          if (normalize) v[j] = v[j] / den;
          s2 = fmaf(v[j], v[j], s2);
        }
        if (xn) {
          if (VEC) *reinterpret_cast<float4*>(xn + r * ld_xn + c) = make_float4(v[0], v[W > 1 ? 1 : 0], v[W > 2 ? 2 : 0], v[W > 3 ? 3 : 0]);
          else xn[r * ld_xn + c] = v[0];
        }
      }
      const int64_t o = r * (int64_t)Dp + c;
      if (hi) {
        float h[W], l[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { h[j] = to_tf32(v[j]); l[j] = to_tf32(v[j] - h[j]); }
        if (VEC) {
          *reinterpret_cast<float4*>(hi + o) = make_float4(h[0], h[W > 1 ? 1 : 0], h[W > 2 ? 2 : 0], h[W > 3 ? 3 : 0]);
          *reinterpret_cast<float4*>(lo + o) = make_float4(l[0], l[W > 1 ? 1 : 0], l[W > 2 ? 2 : 0], l[W > 3 ? 3 : 0]);
        } else { hi[o] = h[0]; lo[o] = l[0]; }
      }
      if (bf) {
#pragma unroll
        for (int j = 0; j < W; ++j) bf[o + j] = __float2bfloat16_rn(v[j]);
      }
      if (h_hi) {
        __half hh[W], hl[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float xs = v[j] * hscale;               // exact (power of two)
          hh[j] = __float2half_rn(xs);
          hl[j] = __float2half_rn(xs - __half2float(hh[j]));
        }
        if (VEC) {
          uint2 ph, pl;
          ph.x = (uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[W > 1 ? 1 : 0]) << 16);
          ph.y = (uint32_t)__half_as_ushort(hh[W > 2 ? 2 : 0]) | ((uint32_t)__half_as_ushort(hh[W > 3 ? 3 : 0]) << 16);
          pl.x = (uint32_t)__half_as_ushort(hl[0]) | ((uint32_t)__half_as_ushort(hl[W > 1 ? 1 : 0]) << 16);
          pl.y = (uint32_t)__half_as_ushort(hl[W > 2 ? 2 : 0]) | ((uint32_t)__half_as_ushort(hl[W > 3 ? 3 : 0]) << 16);
          *reinterpret_cast<uint2*>(h_hi + o) = ph;
          *reinterpret_cast<uint2*>(h_lo + o) = pl;
        } else { h_hi[o] = hh[0]; h_lo[o] = hl[0]; }
      }
    }
    s2 = block_sum_128(s2, sh);
    if (threadIdx.x == 0) {
      if (sqnorm) sqnorm[r] = s2;
      if (norm) norm[r] = sqrtf(s2);
    }
  }
}

// Warp-per-row variant for rows of up to 128 * VPL floats (D = 1280 -> VPL = 10): the row is read ONCE into registers
// (VPL float4 per lane, every load a 512-byte warp transaction), the reductions are shuffles, no block barrier and no
// second sweep.  Same arithmetic per element as k_prep_rows; only the order of the sum of squares differs.
static constexpr int kPrepWarps = 8;

template <int VPL>
__global__ void __launch_bounds__(kPrepWarps * 32)
k_prep_rows_warp(const float* __restrict__ x, int64_t rows, int D, int64_t ld_x, int normalize,
                 float* __restrict__ xn, int64_t ld_xn, float* __restrict__ sqnorm, float* __restrict__ norm,
                 float* __restrict__ hi, float* __restrict__ lo, __nv_bfloat16* __restrict__ bf,
                 __half* __restrict__ h_hi, __half* __restrict__ h_lo, float* __restrict__ h_scale_inv, int Dp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * kPrepWarps + warp; r < rows; r += (int64_t)gridDim.x * kPrepWarps) {
    const float* xr = x + r * ld_x;
    float4 v[VPL];
    float s = 0.f, mx = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = c < D ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      s = fmaf(v[j].x, v[j].x, s); s = fmaf(v[j].y, v[j].y, s); s = fmaf(v[j].z, v[j].z, s); s = fmaf(v[j].w, v[j].w, s);
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
    }
    float den = 1.0f, hscale = 1.0f;
    if (normalize) {
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      den = fmaxf(sqrtf(s), 1e-12f);  // F.normalize: x / max(||x||_2, eps)
    }
    if (h_hi) {
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (normalize) mx = mx / den;
      int e = 0;
      if (mx > 0.f && mx < INFINITY) { (void)frexpf(mx, &e); hscale = ldexpf(1.0f, min(10 - e, 126)); }
      if (lane == 0) h_scale_inv[r] = 1.0f / hscale;
    }
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c >= Dp) continue;
      float t[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      if (c < D) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { if (normalize) t[u] = t[u] / den; s2 = fmaf(t[u], t[u], s2); }
        if (xn) *reinterpret_cast<float4*>(xn + r * ld_xn + c) = make_float4(t[0], t[1], t[2], t[3]);
      }
      const int64_t o = r * (int64_t)Dp + c;
      if (hi) {
        float h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { h[u] = to_tf32(t[u]); l[u] = to_tf32(t[u] - h[u]); }
        *reinterpret_cast<float4*>(hi + o) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(lo + o) = make_float4(l[0], l[1], l[2], l[3]);
      }
      if (bf) {
        uint2 p;
        p.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[1])) << 16);
        p.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[3])) << 16);
        *reinterpret_cast<uint2*>(bf + o) = p;
      }
      if (h_hi) {
        __half hh[4], hl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float xs = t[u] * hscale;               // exact (power of two)
          hh[u] = __float2half_rn(xs);
          hl[u] = __float2half_rn(xs - __half2float(hh[u]));
        }
        uint2 ph, pl;
        ph.x = (uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[1]) << 16);
        ph.y = (uint32_t)__half_as_ushort(hh[2]) | ((uint32_t)__half_as_ushort(hh[3]) << 16);
        pl.x = (uint32_t)__half_as_ushort(hl[0]) | ((uint32_t)__half_as_ushort(hl[1]) << 16);
        pl.y = (uint32_t)__half_as_ushort(hl[2]) | ((uint32_t)__half_as_ushort(hl[3]) << 16);
        *reinterpret_cast<uint2*>(h_hi + o) = ph;
        *reinterpret_cast<uint2*>(h_lo + o) = pl;
      }
    }
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) {
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
  int64_t grid = rows < 148 * 32 ? rows : 148 * 32;
  const bool vec = D % 4 == 0 && ld_x % 4 == 0 && ((uintptr_t)x & 15) == 0 && dp % 4 == 0 &&
                   (!xn || (ld_xn % 4 == 0 && ((uintptr_t)xn & 15) == 0));
  if (vec && dp <= 2048 && !getenv("MPREID_PREP_CTA")) {
    // rows that fit the registers of one warp: single sweep, eight rows per CTA
    const int64_t wgrid = ceil_div(rows, kPrepWarps) < 148 * 8 ? ceil_div(rows, kPrepWarps) : 148 * 8;
#define MPREID_PREP_WARP(V) k_prep_rows_warp<V><<<(unsigned)wgrid, kPrepWarps * 32, 0, (cudaStream_t)stream>>>( \
      x, rows, (int)D, ld_x, normalize, xn, ld_xn, sqnorm, norm, hi, lo, (__nv_bfloat16*)bf, (__half*)h_hi, (__half*)h_lo, h_scale_inv, dp)
    if (dp <= 512) MPREID_PREP_WARP(4); else if (dp <= 1024) MPREID_PREP_WARP(8); else if (dp <= 1280) MPREID_PREP_WARP(10); else if (dp <= 1536) MPREID_PREP_WARP(12);
    else MPREID_PREP_WARP(16);
#undef MPREID_PREP_WARP
    MPREID_CUDA_CHECK(cudaGetLastError());
    return MPREID_OK;
  }
  auto kern = vec ? k_prep_rows<true> : k_prep_rows<false>;
  kern<<<(unsigned)grid, kPrepThreads, 0, (cudaStream_t)stream>>>(
      x, rows, (int)D, ld_x, normalize, xn, ld_xn, sqnorm, norm, hi, lo, (__nv_bfloat16*)bf,
      (__half*)h_hi, (__half*)h_lo, h_scale_inv, dp);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
