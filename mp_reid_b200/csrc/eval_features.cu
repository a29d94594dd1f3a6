// The fused "features -> rank" form of the evaluation (mpreid_rank_eval(dist or NULL + features, ...,
// optional idx)): what R1_mAP_eval.compute() does between torch.cat and eval_func (utils/metrics.py:111-132), as ONE C
// call for consumers without the Python layer: normalise + operand planes (both sides), distance matrix in gallery
// chunks, ranking + CMC / AP, and optionally the first k columns of the stable argsort of every row.  Everything lives in
// the caller's workspace; the distance matrix is written to `dist_out` if given, else to the workspace.
#include "common.cuh"

using namespace mpreid;

namespace {

struct Layout {
  size_t q_hi, q_lo, q_scale, q_sq, q_norm, g_hi, g_lo, g_scale, g_sq, g_norm, q_xn, g_xn, dist, rank_ws, total;
  int64_t Dp, ld_dist;
  size_t rank_bytes;
};

static int plane_elem(int precision) { return precision == MPREID_3XTF32 ? 4 : 2; }

static Layout layout(int64_t Q, int64_t G, int64_t D, int precision, int64_t pos_capacity, bool own_dist) {
  Layout L;
  memset(&L, 0, sizeof(L));
  const int64_t pad = precision == MPREID_3XTF32 ? 32 : 64;
  L.Dp = (D + pad - 1) / pad * pad;
  L.ld_dist = (G + 31) / 32 * 32;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t pe = (size_t)plane_elem(precision);
  const bool two = precision == MPREID_3XTF32 || precision == MPREID_3XFP16 || precision == MPREID_2XFP16;
  const bool simt = precision == MPREID_FP32_SIMT;
  L.q_xn = take(simt ? (size_t)Q * D * 4 : 0); L.g_xn = take(simt ? (size_t)G * D * 4 : 0);
  L.q_hi = take(simt ? 0 : (size_t)Q * L.Dp * pe); L.q_lo = take(two ? (size_t)Q * L.Dp * pe : 0);
  L.g_hi = take(simt ? 0 : (size_t)G * L.Dp * pe); L.g_lo = take(two ? (size_t)G * L.Dp * pe : 0);
  L.q_scale = take((size_t)Q * 4); L.g_scale = take((size_t)G * 4);
  L.q_sq = take((size_t)Q * 4); L.g_sq = take((size_t)G * 4);
  L.q_norm = take((size_t)Q * 4); L.g_norm = take((size_t)G * 4);
  L.dist = take(own_dist ? (size_t)Q * L.ld_dist * 4 : 0);
  L.rank_bytes = mpreid_rank_eval_workspace_bytes(Q, G, pos_capacity);
  L.rank_ws = take(L.rank_bytes);
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t mpreid_eval_features_workspace_bytes(int64_t Q, int64_t G, int64_t D, int precision, int64_t pos_capacity, int own_dist) {
  if (Q <= 0 || G <= 0 || D <= 0 || precision < MPREID_FP32_SIMT || precision > MPREID_2XFP16) return 0;
  return layout(Q, G, D, precision, pos_capacity, own_dist != 0).total;
}

extern "C" int mpreid_eval_features(const float* qf, int64_t ld_q, const float* gf, int64_t ld_g, int64_t Q, int64_t G, int64_t D,
                                    int normalize, int metric, int precision,
                                    const int64_t* q_pid, const int64_t* g_pid, const int64_t* q_cam, const int64_t* g_cam, int junk_mode,
                                    int32_t* first_hit, double* ap, int32_t* num_rel,
                                    int32_t* topk_idx, int topk, float* dist_out, int64_t ld_dist_out,
                                    void* workspace, size_t workspace_bytes, int64_t pos_capacity, int32_t* status, void* stream) {
  MPREID_REQUIRE(qf && gf && q_pid && g_pid && first_hit && ap && num_rel && workspace && status, "eval_features: null pointer");
  MPREID_REQUIRE(Q > 0 && G > 0 && D > 0 && ld_q >= D && ld_g >= D, "eval_features: bad shape Q=%lld G=%lld D=%lld", (long long)Q, (long long)G, (long long)D);
  MPREID_REQUIRE(precision >= MPREID_FP32_SIMT && precision <= MPREID_2XFP16, "eval_features: unknown precision %d", precision);
  MPREID_REQUIRE(!topk_idx || topk >= 1, "eval_features: topk must be >= 1 when topk_idx is given");
  MPREID_REQUIRE(!dist_out || ld_dist_out >= G, "eval_features: ld_dist_out < G");
  MPREID_REQUIRE(((uintptr_t)workspace & 255) == 0, "eval_features: workspace must be 256-byte aligned");
  const Layout L = layout(Q, G, D, precision, pos_capacity, dist_out == nullptr);
  if (workspace_bytes < L.total) { set_error("eval_features: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total); return MPREID_ERR_WORKSPACE; }
  char* w = (char*)workspace;
  const bool simt = precision == MPREID_FP32_SIMT, tf = precision == MPREID_3XTF32, bf = precision == MPREID_BF16;
  const bool h16 = precision == MPREID_3XFP16 || precision == MPREID_2XFP16;
  float* dist = dist_out ? dist_out : (float*)(w + L.dist);
  const int64_t ld_dist = dist_out ? ld_dist_out : L.ld_dist;
  int rc;
  // utils/metrics.py:111-114 (+ the norms of :10-11) for both sides
  for (int side = 0; side < 2; ++side) {
    const float* x = side ? gf : qf;
    const int64_t rows = side ? G : Q, ldx = side ? ld_g : ld_q;
    char* hi = w + (side ? L.g_hi : L.q_hi); char* lo = w + (side ? L.g_lo : L.q_lo);
    float* xn = simt ? (float*)(w + (side ? L.g_xn : L.q_xn)) : nullptr;
    rc = mpreid_prep_rows(x, rows, D, ldx, normalize, xn, D, (float*)(w + (side ? L.g_sq : L.q_sq)), (float*)(w + (side ? L.g_norm : L.q_norm)),
                          tf ? (float*)hi : nullptr, tf ? (float*)lo : nullptr, bf ? (uint16_t*)hi : nullptr,
                          h16 ? (uint16_t*)hi : nullptr, h16 ? (uint16_t*)lo : nullptr, h16 ? (float*)(w + (side ? L.g_scale : L.q_scale)) : nullptr,
                          simt ? D : L.Dp, stream);
    if (rc != MPREID_OK) return rc;
  }
  const float* q_aux = metric == MPREID_ARCCOS ? (float*)(w + L.q_norm) : ((metric == MPREID_ONE_MINUS_DOT || metric == MPREID_DOT) ? nullptr : (float*)(w + L.q_sq));
  const float* g_aux = metric == MPREID_ARCCOS ? (float*)(w + L.g_norm) : ((metric == MPREID_ONE_MINUS_DOT || metric == MPREID_DOT) ? nullptr : (float*)(w + L.g_sq));
  // :124-131
  if (simt)
    rc = mpreid_dist_matrix(w + L.q_xn, nullptr, w + L.g_xn, nullptr, q_aux, g_aux, nullptr, nullptr, Q, G, D, D, metric, precision, dist, ld_dist, nullptr, stream);
  else
    rc = mpreid_dist_matrix(w + L.q_hi, (tf || h16) ? w + L.q_lo : nullptr, w + L.g_hi, (tf || h16) ? w + L.g_lo : nullptr, q_aux, g_aux,
                            h16 ? (float*)(w + L.q_scale) : nullptr, h16 ? (float*)(w + L.g_scale) : nullptr, Q, G, L.Dp, L.Dp, metric, precision,
                            dist, ld_dist, nullptr, stream);
  if (rc != MPREID_OK) return rc;
  // :132 (per-query part of eval_func)
  rc = mpreid_rank_eval(dist, ld_dist, Q, G, q_pid, g_pid, q_cam, g_cam, junk_mode, first_hit, ap, num_rel, w + L.rank_ws, L.rank_bytes, pos_capacity,
                        status, stream);
  if (rc != MPREID_OK) return rc;
  // optional: the first `topk` entries of np.argsort(distmat, axis=1, kind='stable') (:39) of every row
  if (topk_idx) rc = mpreid_row_topk(dist, ld_dist, Q, G, topk, nullptr, topk_idx, nullptr, stream);
  return rc;
}
