// C-ABI plumbing: error reporting, device queries, the distance-matrix dispatcher.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mpreid {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count_of_current_device() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int launch_dist_simt(const float* q, const float* g, const float* q_aux, const float* g_aux, int64_t Q, int64_t G,
                     int64_t K, int64_t ldk, int metric, float* out, int64_t ld_out, float* row_max, cudaStream_t st);
int launch_dist_tc(const void* qa, const void* qb, const void* ga, const void* gb, const float* q_aux, const float* g_aux,
                   const float* q_scale, const float* g_scale, int64_t Q, int64_t G, int64_t ldk, int metric, int precision, float* out, int64_t ld_out,
                   float* row_max, int symmetric, cudaStream_t st, const TopkFuse* fuse);

}  // namespace mpreid

using namespace mpreid;

extern "C" const char* mpreid_last_error(void) { return g_err; }
extern "C" int mpreid_abi_version(void) { return MPREID_ABI_VERSION; }

extern "C" int mpreid_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05) {
  cudaDeviceProp p;
  MPREID_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (has_tcgen05) *has_tcgen05 = (p.major == 10) ? 1 : 0;
  return MPREID_OK;
}

extern "C" int mpreid_dist_matrix(const void* qa, const void* qb, const void* ga, const void* gb,
                                  const float* q_aux, const float* g_aux, const float* q_scale, const float* g_scale,
                                  int64_t Q, int64_t G, int64_t K, int64_t ldk,
                                  int metric, int precision,
                                  float* out, int64_t ld_out, float* row_max, void* stream) {
  MPREID_REQUIRE(qa && ga && out, "dist_matrix: null operand");
  MPREID_REQUIRE(Q > 0 && G > 0 && K > 0 && ldk >= K && ld_out >= G && Q < INT32_MAX && G < INT32_MAX,
                 "dist_matrix: bad shape Q=%lld G=%lld K=%lld ldk=%lld ld_out=%lld", (long long)Q, (long long)G,
                 (long long)K, (long long)ldk, (long long)ld_out);
  MPREID_REQUIRE(metric >= MPREID_SQEUCLID && metric <= MPREID_DOT, "dist_matrix: unknown metric %d", metric);
  MPREID_REQUIRE(metric == MPREID_ONE_MINUS_DOT || metric == MPREID_DOT || (q_aux && g_aux), "dist_matrix: metric %d needs q_aux/g_aux", metric);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MPREID_FP32_SIMT)
    return launch_dist_simt((const float*)qa, (const float*)ga, q_aux, g_aux, Q, G, K, ldk, metric, out, ld_out, row_max, st);
  MPREID_REQUIRE(precision == MPREID_3XTF32 || precision == MPREID_BF16 || precision == MPREID_3XFP16 || precision == MPREID_2XFP16,
                 "dist_matrix: unknown precision %d", precision);
  MPREID_REQUIRE(precision == MPREID_BF16 || (qb && (gb || precision == MPREID_2XFP16)), "dist_matrix: the split modes need the lo planes");
  MPREID_REQUIRE((precision != MPREID_3XFP16 && precision != MPREID_2XFP16) || (q_scale && g_scale),
                 "dist_matrix: the FP16 split modes need the per-row scales");
  return launch_dist_tc(qa, qb, ga, gb, q_aux, g_aux, q_scale, g_scale, Q, G, ldk, metric, precision, out, ld_out, row_max, 0, st, nullptr);
}

extern "C" int mpreid_dist_matrix_symmetric(const void* xa, const void* xb, const float* x_aux, const float* x_scale,
                                            int64_t N, int64_t K, int64_t ldk, int metric, int precision,
                                            float* out, int64_t ld_out, float* row_max, void* stream) {
  MPREID_REQUIRE(xa && out, "dist_matrix_symmetric: null operand");
  MPREID_REQUIRE(N > 0 && K > 0 && ldk >= K && ld_out >= N && N < INT32_MAX, "dist_matrix_symmetric: bad shape N=%lld K=%lld", (long long)N,
                 (long long)K);
  MPREID_REQUIRE(metric >= MPREID_SQEUCLID && metric <= MPREID_DOT, "dist_matrix_symmetric: unknown metric %d", metric);
  MPREID_REQUIRE(metric == MPREID_ONE_MINUS_DOT || metric == MPREID_DOT || x_aux, "dist_matrix_symmetric: metric %d needs x_aux", metric);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MPREID_FP32_SIMT)   // the validation kernel has no mirrored mode: plain all-pairs launch
    return launch_dist_simt((const float*)xa, (const float*)xa, x_aux, x_aux, N, N, K, ldk, metric, out, ld_out, row_max, st);
  MPREID_REQUIRE(precision == MPREID_3XTF32 || precision == MPREID_BF16 || precision == MPREID_3XFP16 || precision == MPREID_2XFP16,
                 "dist_matrix_symmetric: unknown precision %d", precision);
  MPREID_REQUIRE(precision == MPREID_BF16 || xb, "dist_matrix_symmetric: the split modes need the lo plane");
  MPREID_REQUIRE((precision != MPREID_3XFP16 && precision != MPREID_2XFP16) || x_scale,
                 "dist_matrix_symmetric: the FP16 split modes need the per-row scales");
  return launch_dist_tc(xa, xb, xa, xb, x_aux, x_aux, x_scale, x_scale, N, N, ldk, metric, precision, out, ld_out, row_max, 1, st, nullptr);
}

extern "C" double mpreid_host_average_precision(const int32_t* ranks_host, int m, int64_t n) {
  if (!ranks_host || m <= 0) return 0.0;
  return pairwise_sparse_sum(ranks_host, m, n) / (double)m;
}

extern "C" void mpreid_host_order_keys(const float* values_host, int64_t n, uint32_t* keys_host) {
  for (int64_t i = 0; i < n; ++i) keys_host[i] = order_key(values_host[i]);
}

// Fused flavour of the all-pairs launch for re-ranking (utils/reranking.py:36-48 without the N x N matrix): see the header.
extern "C" int mpreid_dist_symmetric_topk(const void* xa, const void* xb, const float* x_sqnorm, const float* x_scale,
                                          int64_t N, int64_t K, int64_t ldk, int precision,
                                          const float* thr, uint64_t* cand, int32_t* cand_cnt, int64_t cand_cap,
                                          int64_t Q, float* out_qg, int64_t ld_out, float* row_max, int own_mod, int own_rank, void* stream) {
  MPREID_REQUIRE(xa && x_sqnorm && thr && cand && cand_cnt && out_qg && row_max, "dist_symmetric_topk: null pointer");
  MPREID_REQUIRE(own_mod >= 1 && own_rank >= 0 && own_rank < own_mod, "dist_symmetric_topk: bad tile ownership %d of %d", own_rank, own_mod);
  MPREID_REQUIRE(N > 1 && K > 0 && ldk >= K && N < INT32_MAX && Q > 0 && Q < N, "dist_symmetric_topk: bad shape N=%lld Q=%lld", (long long)N, (long long)Q);
  MPREID_REQUIRE(cand_cap >= 1 && cand_cap < (1 << 24) && ld_out >= (N - Q) + (Q & 31), "dist_symmetric_topk: bad candidate capacity / ld_out");
  MPREID_REQUIRE(precision == MPREID_3XTF32 || precision == MPREID_BF16 || precision == MPREID_3XFP16 || precision == MPREID_2XFP16,
                 "dist_symmetric_topk: needs a tensor-core precision mode");
  MPREID_REQUIRE(precision == MPREID_BF16 || xb, "dist_symmetric_topk: the split modes need the lo plane");
  MPREID_REQUIRE((precision != MPREID_3XFP16 && precision != MPREID_2XFP16) || x_scale, "dist_symmetric_topk: the FP16 split modes need the per-row scales");
  TopkFuse f;
  f.thr = thr; f.cand = (unsigned long long*)cand; f.cand_cnt = cand_cnt; f.cap = (int)cand_cap;
  f.keep_rows = (int)Q; f.keep_col0 = (int)(Q & ~(int64_t)31);
  f.own_mod = own_mod; f.own_rank = own_rank;
  return launch_dist_tc(xa, xb, xa, xb, x_sqnorm, x_sqnorm, x_scale, x_scale, N, N, ldk, MPREID_SQEUCLID, precision, out_qg, ld_out, row_max, 1,
                        (cudaStream_t)stream, &f);
}
