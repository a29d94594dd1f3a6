// Stage-1 cached-feature contrastive step: processor/processor_uniprompt_stage1.py:88-93 evaluates
//     loss = SupConLoss(image, text, y, y) + SupConLoss(text, image, y, y)          (loss/supcontrast.py:17-31)
// on the cached image features of a batch and the text features the prompt learner produced for its labels.  Both terms
// read the same similarity matrix S = A.B^T / T, once along its rows and once along its columns, so one pass produces
// both losses and dS; the feature gradients are two small products with dS.
//   row term   L_r = -(1/Ba) sum_i (1/|P_i|) sum_{j in P_i} (z_ij - logsumexp_j z_ij),     P_i = {j : la_i == lb_j}
//   col term   L_c = the same along the columns of S (== SupConLoss(B, A))
//   dL/dz_ij   = (softmax_row_ij - m_ij/|P_i|) / Ba + (softmax_col_ij - m_ij/|P_j|) / Bb
// S itself comes from the distance kernels (metric MPREID_DOT: tcgen05 for large batches, SIMT for the usual 64..128).
// A row (column) without positives gives 0/0 = NaN in the reference; the same happens here.
#include "common.cuh"

namespace mpreid {

// stats[0][r] = max, [1][r] = log sum exp(z - max), [2][r] = number of positives, [3][r] = sum of positive z
// One warp per row (rows = 1) or per column (rows = 0) of S [Ba, Bb].
__global__ void k_supcon_stats(const float* __restrict__ S, int64_t ld, int Ba, int Bb, float inv_t,
                               const int64_t* __restrict__ la, const int64_t* __restrict__ lb, int by_rows, float* __restrict__ stats, int n_pad) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n_lines = by_rows ? Ba : Bb, n_el = by_rows ? Bb : Ba;
  if (r >= n_lines) return;
  const int64_t me = by_rows ? la[r] : lb[r];
  const int64_t* other = by_rows ? lb : la;
  float mx = -INFINITY;
  for (int e = lane; e < n_el; e += 32) mx = fmaxf(mx, (by_rows ? S[(int64_t)r * ld + e] : S[(int64_t)e * ld + r]) * inv_t);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f, cnt = 0.f, ps = 0.f;
  for (int e = lane; e < n_el; e += 32) {
    const float z = (by_rows ? S[(int64_t)r * ld + e] : S[(int64_t)e * ld + r]) * inv_t;
    se += expf(z - mx);
    if (other[e] == me) { cnt += 1.f; ps += z; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); ps += __shfl_xor_sync(0xffffffffu, ps, o);
  }
  if (lane == 0) { stats[r] = mx; stats[n_pad + r] = logf(se); stats[2 * n_pad + r] = cnt; stats[3 * n_pad + r] = ps; }
}

// loss[0] = row term, loss[1] = column term, loss[2] = their sum.  One CTA; fixed summation order.
__global__ void k_supcon_loss(const float* __restrict__ rs, const float* __restrict__ cs, int Ba, int Bb, int pa, int pb, int do_cols, float* loss) {
  __shared__ float sh[32];
  float terms[2] = {0.f, 0.f};
  for (int side = 0; side < (do_cols ? 2 : 1); ++side) {
    const float* st = side ? cs : rs;
    const int n = side ? Bb : Ba, p = side ? pb : pa;
    float acc = 0.f;
    for (int r = threadIdx.x; r < n; r += blockDim.x) acc += st[3 * p + r] / st[2 * p + r] - st[r] - st[p + r];   // mean_log_prob_pos (:28)
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
      terms[side] = -t / (float)n;                                                                                  // (:29)
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { loss[0] = terms[0]; loss[1] = terms[1]; loss[2] = terms[0] + terms[1]; }
}

// dS[i][j] = dL/dS_ij (the 1/T of z = S/T included)
__global__ void k_supcon_ds(const float* __restrict__ S, int64_t ld, int Ba, int Bb, float inv_t,
                            const int64_t* __restrict__ la, const int64_t* __restrict__ lb,
                            const float* __restrict__ rs, const float* __restrict__ cs, int pa, int pb, int do_cols,
                            float gscale_rows, float gscale_cols, float* __restrict__ dS, int64_t ld_ds) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= Bb || i >= Ba) return;
  const float z = S[(int64_t)i * ld + j] * inv_t;
  const float m = la[i] == lb[j] ? 1.f : 0.f;
  float g = gscale_rows * (expf(z - rs[i] - rs[pa + i]) - m / rs[2 * pa + i]) / (float)Ba;
  if (do_cols) g += gscale_cols * (expf(z - cs[j] - cs[pb + j]) - m / cs[2 * pb + j]) / (float)Bb;
  dS[(int64_t)i * ld_ds + j] = g * inv_t;
}

// C[M, N] = op(A) . B with B [K, N] row-major; op(A) = A [M, K] (trans = 0) or A^T with A stored [K, M] (trans = 1).
// 32 x 32 output tile per CTA, fp32, fixed k order: small (batch-sized) products only.
__global__ void __launch_bounds__(256)
k_small_gemm(const float* __restrict__ A, int64_t lda, int trans, const float* __restrict__ Bm, int64_t ldb, int M, int N, int K,
             float* __restrict__ C, int64_t ldc) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      // sa[r][tx] = op(A)[m0 + r][k0 + tx]
      float a = 0.f;
      if (!trans) { if (m0 + r < M && k0 + tx < K) a = A[(int64_t)(m0 + r) * lda + k0 + tx]; }
      else { if (m0 + tx < M && k0 + r < K) a = A[(int64_t)(k0 + r) * lda + m0 + tx]; }
      if (!trans) sa[r][tx] = a; else sa[tx][r] = a;
      sb[r][tx] = (k0 + r < K && n0 + tx < N) ? Bm[(int64_t)(k0 + r) * ldb + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float b = sb[kk][tx];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(sa[ty + 8 * u][kk], b, acc[u]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int m = m0 + ty + 8 * u, n = n0 + tx;
    if (m < M && n < N) C[(int64_t)m * ldc + n] = acc[u];
  }
}

}  // namespace mpreid

using namespace mpreid;

extern "C" size_t mpreid_supcon_workspace_bytes(int64_t Ba, int64_t Bb) {
  if (Ba <= 0 || Bb <= 0) return 0;
  const size_t pa = (size_t)((Ba + 63) & ~63), pb = (size_t)((Bb + 63) & ~63);
  return align_up((size_t)Ba * (size_t)((Bb + 31) & ~31) * 4, 256) + align_up(4 * pa * 4, 256) + align_up(4 * pb * 4, 256);
}

extern "C" int mpreid_supcon_step(const float* S, int64_t ld_s, int64_t Ba, int64_t Bb, const int64_t* labels_a, const int64_t* labels_b,
                                  float temperature, int both_directions, float grad_scale_rows, float grad_scale_cols,
                                  const float* a, int64_t ld_a, const float* b, int64_t ld_b, int64_t D,
                                  float* loss, float* grad_a, int64_t ld_ga, float* grad_b, int64_t ld_gb,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MPREID_REQUIRE(S && labels_a && labels_b && loss && workspace, "supcon_step: null pointer");
  MPREID_REQUIRE(Ba > 0 && Bb > 0 && ld_s >= Bb && Ba < (1 << 24) && Bb < (1 << 24) && temperature > 0.f, "supcon_step: bad shape / temperature");
  MPREID_REQUIRE((!grad_a && !grad_b) || (a && b && D > 0 && ld_a >= D && ld_b >= D), "supcon_step: gradients need both feature matrices");
  MPREID_REQUIRE(workspace_bytes >= mpreid_supcon_workspace_bytes(Ba, Bb) && ((uintptr_t)workspace & 255) == 0, "supcon_step: workspace too small or misaligned");
  const int pa = (int)((Ba + 63) & ~63), pb = (int)((Bb + 63) & ~63);
  const int64_t ld_ds = (Bb + 31) & ~31;
  char* base = (char*)workspace;
  float* dS = (float*)base; base += align_up((size_t)Ba * ld_ds * 4, 256);
  float* rs = (float*)base; base += align_up(4 * (size_t)pa * 4, 256);
  float* cs = (float*)base;
  const float inv_t = 1.0f / temperature;
  k_supcon_stats<<<(unsigned)ceil_div(Ba, 8), 256, 0, st>>>(S, ld_s, (int)Ba, (int)Bb, inv_t, labels_a, labels_b, 1, rs, pa);
  if (both_directions) k_supcon_stats<<<(unsigned)ceil_div(Bb, 8), 256, 0, st>>>(S, ld_s, (int)Ba, (int)Bb, inv_t, labels_a, labels_b, 0, cs, pb);
  k_supcon_loss<<<1, 256, 0, st>>>(rs, cs, (int)Ba, (int)Bb, pa, pb, both_directions, loss);
  if (grad_a || grad_b) {
    dim3 grid((unsigned)ceil_div(Bb, 128), (unsigned)Ba);
    k_supcon_ds<<<grid, 128, 0, st>>>(S, ld_s, (int)Ba, (int)Bb, inv_t, labels_a, labels_b, rs, cs, pa, pb, both_directions,
                                      grad_scale_rows, grad_scale_cols, dS, ld_ds);
    if (grad_a) {   // dL/dA = dS . B
      dim3 g((unsigned)ceil_div(D, 32), (unsigned)ceil_div(Ba, 32));
      k_small_gemm<<<g, 256, 0, st>>>(dS, ld_ds, 0, b, ld_b, (int)Ba, (int)D, (int)Bb, grad_a, ld_ga);
    }
    if (grad_b) {   // dL/dB = dS^T . A
      dim3 g((unsigned)ceil_div(D, 32), (unsigned)ceil_div(Bb, 32));
      k_small_gemm<<<g, 256, 0, st>>>(dS, ld_ds, 1, a, ld_a, (int)Bb, (int)D, (int)Ba, grad_b, ld_gb);
    }
  }
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
