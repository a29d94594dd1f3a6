// Plain fp32 FFMA distance kernel (MPREID_FP32_SIMT).  Not the fast path: it is the on-device
// validation twin of the tcgen05 kernel (dist_tc.cu) and serves shapes too small to tile.
// 64x64 output tile per CTA, 16x16 threads, 4x4 micro-tile, K staged 16 at a time through smem.
#include "epilogue.cuh"

namespace mpreid {

static constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
k_dist_simt(const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ q_aux,
            const float* __restrict__ g_aux, int Q, int G, int K, int64_t ldk, int metric,
            float* __restrict__ out, int64_t ld_out, float* __restrict__ row_max) {
  __shared__ float sq[TK][TM + 4];
  __shared__ float sg[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r, gk = k0 + c;
      sq[c][r] = (gm < Q && gk < K) ? q[(int64_t)gm * ldk + gk] : 0.f;
      const int gn = n0 + r;
      sg[c][r] = (gn < G && gk < K) ? g[(int64_t)gn * ldk + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sq[kk][ty * 4 + i]; b[i] = sg[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= Q) continue;
    const float qa = q_aux ? q_aux[gm] : 0.f;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= G) continue;
      const float d = finish_distance_rt(metric, acc[i][j], qa, g_aux ? g_aux[gn] : 0.f);
      out[(int64_t)gm * ld_out + gn] = d;
      mx = fmaxf(mx, d);
    }
    if (row_max && mx > -INFINITY) atomic_max_f32(&row_max[gm], mx);
  }
}

int launch_dist_simt(const float* q, const float* g, const float* q_aux, const float* g_aux, int64_t Q, int64_t G,
                     int64_t K, int64_t ldk, int metric, float* out, int64_t ld_out, float* row_max, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(G, TN), (unsigned)ceil_div(Q, TM));
  MPREID_REQUIRE(grid.y <= 65535, "dist_simt: Q too large for the validation kernel (%lld)", (long long)Q);
  k_dist_simt<<<grid, 256, 0, st>>>(q, g, q_aux, g_aux, (int)Q, (int)G, (int)K, ldk, metric, out, ld_out, row_max);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

}  // namespace mpreid
