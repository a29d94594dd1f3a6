// Batch-hard triplet distances with a hand-written backward: loss/triplet_loss.py:16-31 (euclidean_dist:
// sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 1e-12))) and :50-103 (hard_example_mining) fused into one forward kernel, and the
// gradient of (dist_ap, dist_an) with respect to the features as a second, atomics-free kernel.  The reference calls
// this under autograd every training step (loss/make_loss.py:46-50) on a batch of 64 x {768, 512, 1280}: far too
// small for the tensor-core path, so both kernels are fp32 SIMT, one CTA per anchor row.
#include "common.cuh"

namespace mpreid {

static constexpr int kTriThreads = 128;

// One CTA per anchor i.  Warps walk the batch rows j; lanes stride the feature dimension and accumulate x_i.x_j and
// |x_j|^2 together, so every row is read once per anchor (the batch is L2 resident).  Then warp 0 mines:
// dist_ap = max over same-label j (the diagonal included, as in the reference), dist_an = min over other-label j,
// lowest index on ties.
__global__ void __launch_bounds__(kTriThreads)
k_triplet_forward(const float* __restrict__ x, int64_t ld, int B, int D, const int64_t* __restrict__ labels,
                  float* __restrict__ dist_ap, float* __restrict__ dist_an, int64_t* __restrict__ p_inds, int64_t* __restrict__ n_inds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xi = reinterpret_cast<float*>(smem_raw);          // [D]
  float* dist = xi + ((D + 3) & ~3);                       // [B]
  const int i = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xr = x + (int64_t)i * ld;
  float sq = 0.f;
  for (int c = tid; c < D; c += kTriThreads) { const float v = xr[c]; xi[c] = v; }
  __syncthreads();
  // |x_i|^2 by every warp on its own (same order in all of them)
  for (int c = lane; c < D; c += 32) sq = fmaf(xi[c], xi[c], sq);
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  for (int j = warp; j < B; j += kTriThreads / 32) {
    const float* xj = x + (int64_t)j * ld;
    float dot = 0.f, sj = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = __ldg(xj + c); dot = fmaf(xi[c], v, dot); sj = fmaf(v, v, sj); }
    for (int o = 16; o > 0; o >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); sj += __shfl_xor_sync(0xffffffffu, sj, o); }
    if (lane == 0) dist[j] = sqrtf(fmaxf((sq + sj) - 2.0f * dot, 1e-12f));   // loss/triplet_loss.py:26-30
  }
  __syncthreads();
  if (warp == 0) {
    const int64_t me = labels[i];
    float bp = -INFINITY, bn = INFINITY;
    int ip = -1, in = -1;
    for (int j = lane; j < B; j += 32) {
      const float v = dist[j];
      if (labels[j] == me) { if (ip < 0 || v > bp) { bp = v; ip = j; } }
      else { if (in < 0 || v < bn) { bn = v; in = j; } }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float obp = __shfl_xor_sync(0xffffffffu, bp, o); const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
      const float obn = __shfl_xor_sync(0xffffffffu, bn, o); const int oin = __shfl_xor_sync(0xffffffffu, in, o);
      if (oip >= 0 && (ip < 0 || obp > bp || (obp == bp && oip < ip))) { bp = obp; ip = oip; }
      if (oin >= 0 && (in < 0 || obn < bn || (obn == bn && oin < in))) { bn = obn; in = oin; }
    }
    if (lane == 0) {
      dist_ap[i] = bp; dist_an[i] = bn;
      p_inds[i] = ip; n_inds[i] = in;
    }
  }
}

// d dist(a, b) / d x_a = (x_a - x_b) / dist(a, b)  (zero where the clamp was active).  One CTA per feature row j gathers
// every term that touches x_j -- as the anchor of its own pair, and as the selected positive / negative of any other
// anchor -- in a fixed order: no atomics, bit-reproducible.
__global__ void __launch_bounds__(kTriThreads)
k_triplet_backward(const float* __restrict__ x, int64_t ld, int B, int D, const int64_t* __restrict__ p_inds, const int64_t* __restrict__ n_inds,
                   const float* __restrict__ dist_ap, const float* __restrict__ dist_an,
                   const float* __restrict__ g_ap, const float* __restrict__ g_an, float* __restrict__ grad, int64_t ld_g) {
  __shared__ int s_other[2 * 1024];
  __shared__ float s_coef[2 * 1024];
  __shared__ int s_n;
  const int j = blockIdx.x, tid = threadIdx.x;
  const float clamp_d = sqrtf(1e-12f);
  // the (other row, coefficient) pairs of row j, in anchor order; at most 2B of them, processed in chunks of 2048
  for (int i0 = 0; i0 < B; i0 += 1024) {
    if (tid == 0) {
      int n = 0;
      const int i1 = min(B, i0 + 1024);
      for (int i = i0; i < i1; ++i) {
        const int p = (int)p_inds[i], q = (int)n_inds[i];
        const float dp = dist_ap[i], dn = dist_an[i];
        if (p >= 0 && dp > clamp_d && p != i) {
          if (i == j) { s_other[n] = p; s_coef[n] = g_ap[i] / dp; ++n; }
          else if (p == j) { s_other[n] = i; s_coef[n] = g_ap[i] / dp; ++n; }
        }
        if (q >= 0 && dn > clamp_d && dn < INFINITY) {
          if (i == j) { s_other[n] = q; s_coef[n] = g_an[i] / dn; ++n; }
          else if (q == j) { s_other[n] = i; s_coef[n] = g_an[i] / dn; ++n; }
        }
      }
      s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const float* xj = x + (int64_t)j * ld;
    for (int c = tid; c < D; c += kTriThreads) {
      float acc = i0 == 0 ? 0.f : grad[(int64_t)j * ld_g + c];
      const float v = xj[c];
      for (int t = 0; t < n; ++t) acc = fmaf(s_coef[t], v - __ldg(x + (int64_t)s_other[t] * ld + c), acc);
      grad[(int64_t)j * ld_g + c] = acc;
    }
    __syncthreads();
  }
}

}  // namespace mpreid

using namespace mpreid;

extern "C" int mpreid_triplet_forward(const float* x, int64_t ld_x, int64_t B, int64_t D, const int64_t* labels,
                                      float* dist_ap, float* dist_an, int64_t* p_inds, int64_t* n_inds, void* stream) {
  MPREID_REQUIRE(x && labels && dist_ap && dist_an && p_inds && n_inds, "triplet_forward: null pointer");
  MPREID_REQUIRE(B > 0 && D > 0 && ld_x >= D && B < (1 << 20), "triplet_forward: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  const size_t smem = (size_t)(((D + 3) & ~3) + B) * sizeof(float);
  MPREID_REQUIRE(smem <= 200 * 1024, "triplet_forward: B + D too large for one CTA's shared memory (%zu bytes)", smem);
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_triplet_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_triplet_forward<<<(unsigned)B, kTriThreads, smem, (cudaStream_t)stream>>>(x, ld_x, (int)B, (int)D, labels, dist_ap, dist_an, p_inds, n_inds);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_triplet_backward(const float* x, int64_t ld_x, int64_t B, int64_t D, const int64_t* p_inds, const int64_t* n_inds,
                                       const float* dist_ap, const float* dist_an, const float* g_ap, const float* g_an,
                                       float* grad_x, int64_t ld_g, void* stream) {
  MPREID_REQUIRE(x && p_inds && n_inds && dist_ap && dist_an && g_ap && g_an && grad_x, "triplet_backward: null pointer");
  MPREID_REQUIRE(B > 0 && D > 0 && ld_x >= D && ld_g >= D && B < (1 << 20), "triplet_backward: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  k_triplet_backward<<<(unsigned)B, kTriThreads, 0, (cudaStream_t)stream>>>(x, ld_x, (int)B, (int)D, p_inds, n_inds, dist_ap, dist_an, g_ap, g_an,
                                                                            grad_x, ld_g);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
