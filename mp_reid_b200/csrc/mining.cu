// Batch-hard mining on a distance matrix (forward pass of loss/triplet_loss.py:50-103):
// per anchor row the farthest same-label entry (the diagonal counts, as in the reference) and the
// closest other-label entry, first index on ties (torch.max / torch.min semantics).  One warp per row.
#include "common.cuh"

namespace mpreid {

__global__ void k_hard_mining(const float* __restrict__ dist, int64_t ld, int N, const int64_t* __restrict__ labels,
                              float* __restrict__ dist_ap, float* __restrict__ dist_an,
                              int64_t* __restrict__ p_inds, int64_t* __restrict__ n_inds) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int64_t me = labels[row];
  const float* d = dist + (int64_t)row * ld;
  float bp = -INFINITY, bn = INFINITY;
  int ip = -1, in = -1;
  for (int j = lane; j < N; j += 32) {
    const float v = d[j];
    if (labels[j] == me) { if (v > bp || ip < 0) { bp = v; ip = j; } }
    else { if (v < bn || in < 0) { bn = v; in = j; } }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float obp = __shfl_xor_sync(0xffffffffu, bp, o); const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
    const float obn = __shfl_xor_sync(0xffffffffu, bn, o); const int oin = __shfl_xor_sync(0xffffffffu, in, o);
    if (oip >= 0 && (ip < 0 || obp > bp || (obp == bp && oip < ip))) { bp = obp; ip = oip; }
    if (oin >= 0 && (in < 0 || obn < bn || (obn == bn && oin < in))) { bn = obn; in = oin; }
  }
  if (lane == 0) {
    dist_ap[row] = bp; dist_an[row] = bn;
    if (p_inds) p_inds[row] = ip;
    if (n_inds) n_inds[row] = in;
  }
}

}  // namespace mpreid

using namespace mpreid;

extern "C" int mpreid_hard_example_mining(const float* dist, int64_t ld_dist, int64_t N, const int64_t* labels,
                                          float* dist_ap, float* dist_an, int64_t* p_inds, int64_t* n_inds, void* stream) {
  MPREID_REQUIRE(dist && labels && dist_ap && dist_an && N > 0 && N < INT32_MAX && ld_dist >= N, "hard_example_mining: bad arguments");
  const int rows_per_cta = 8;
  k_hard_mining<<<(unsigned)ceil_div(N, rows_per_cta), rows_per_cta * 32, 0, (cudaStream_t)stream>>>(dist, ld_dist, (int)N, labels, dist_ap,
                                                                                                   dist_an, p_inds, n_inds);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
