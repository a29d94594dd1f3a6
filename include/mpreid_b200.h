/*
 * mpreid_b200 — C ABI of the B200-native MP-ReID evaluation / retrieval hot path.
 *
 * The reference (MP-ReID/mp-reid) is pure Python: its "FFI" for this path is the set of functions
 * in utils/metrics.py and utils/reranking.py.  Each entry point below names the reference lines it
 * replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C, no exceptions; every call returns 0 on success or a negative MPREID_ERR_* code,
 *     the message is available from mpreid_last_error() (thread-local);
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all buffers;
 *   - every call takes the CUDA stream to launch on (cudaStream_t passed as void*), never
 *     synchronises it and never allocates; workspaces are sized by the *_workspace_bytes queries;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - labels (pid / camid) are int64, as numpy gives them; INT64_MIN is reserved.
 */
#ifndef MPREID_B200_H
#define MPREID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPREID_ABI_VERSION 1

#if defined(__GNUC__)
#define MPREID_API __attribute__((visibility("default")))
#else
#define MPREID_API
#endif

enum mpreid_error {
  MPREID_OK = 0,
  MPREID_ERR_INVALID = -1,    /* bad argument                                   */
  MPREID_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed            */
  MPREID_ERR_UNSUPPORTED = -3,/* e.g. tcgen05 path requested on a non-sm_100 GPU */
  MPREID_ERR_WORKSPACE = -4   /* workspace too small                            */
};

/* distance epilogues */
enum mpreid_metric {
  MPREID_SQEUCLID = 0,      /* ||q||^2+||g||^2-2q.g, no sqrt, no clamp   utils/metrics.py:7-13          */
  MPREID_ARCCOS = 1,        /* arccos(clip(q.g/(|q||g|), +-(1-1e-5)))    utils/metrics.py:15-25         */
  MPREID_ONE_MINUS_DOT = 2, /* 1 - q.g                      processor/processor_uniprompt_stage2.py:466 */
  MPREID_SQRT_EUCLID = 3,   /* sqrt(clamp(sqeuclid, 1e-12))              loss/triplet_loss.py:16-31     */
  MPREID_DOT = 4            /* q.g (similarity logits)                   loss/supcontrast.py:23         */
};

/* how the q.g contraction is computed */
enum mpreid_precision {
  MPREID_FP32_SIMT = 0, /* plain fp32 FFMA tiles (validation / tiny shapes)                        */
  MPREID_3XTF32 = 1,    /* tcgen05 kind::tf32, error-compensated hi/lo split: fp32-accurate        */
  MPREID_BF16 = 2,      /* tcgen05 kind::f16 on bf16-rounded operands, fp32 accumulate             */
  MPREID_3XFP16 = 3,    /* tcgen05 kind::f16 on a per-row power-of-two scaled fp16 hi/lo split:
                           fp32-accurate like 3xTF32 at twice the MMA rate                          */
  MPREID_2XFP16 = 4     /* same planes, but the gallery side contributes only its hi plane (drops hi*lo):
                           two MMAs per k-step, error ~2^-12 per product; a stated fast mode, same
                           operands as MPREID_3XFP16 (gb may be NULL)                                */
};

enum mpreid_junk {
  MPREID_JUNK_NONE = 0,   /* HEAD behaviour: utils/metrics.py:55 (remove = False)                  */
  MPREID_JUNK_PID_CAM = 1 /* same pid AND same camid removed: utils/metrics.py:54 (commented),
                             processor/processor_uniprompt_stage2.py:483-488 (live)                */
};

MPREID_API const char* mpreid_last_error(void);
MPREID_API int mpreid_abi_version(void);
/* sm count, compute capability and whether the tcgen05 kernels can run on `device` */
MPREID_API int mpreid_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05);

/* ---- feature preparation -------------------------------------------------------------------
 * Replaces torch.nn.functional.normalize(feats, dim=1, p=2) (utils/metrics.py:114) and the
 * torch.pow(x,2).sum(1) terms of utils/metrics.py:10-11 / utils/reranking.py:38-39 in ONE pass
 * over x, and emits the tensor-core operand planes:
 *   xn      [rows, D]   fp32, normalised (or copied) features             (may be NULL)
 *   sqnorm  [rows]      fp32, sum of squares of the xn row                (may be NULL)
 *   norm    [rows]      fp32, sqrt of the above                           (may be NULL)
 *   hi, lo  [rows, Dp]  fp32 holding TF32-representable values, xn = hi + lo (+2^-22 rel.), zero
 *                       padded from D to Dp                                (both or neither)
 *   bf      [rows, Dp]  bf16 (round-to-nearest-even) copy of xn, zero padded (may be NULL)
 *   h_hi, h_lo [rows, Dp] fp16 split of the row scaled by 2^s (s per row such that max|xn|*2^s is in
 *                       [512, 1024)):  xn * 2^s = h_hi + h_lo (+2^-22 rel.);  h_scale_inv[rows] = 2^-s
 *                       (all three or none)
 * Dp must be a multiple of 32 (fp32 planes) / 64 (16-bit planes): TMA boxes are 128 B wide.  x is fp32. */
MPREID_API int mpreid_prep_rows(const float* x, int64_t rows, int64_t D, int64_t ld_x, int normalize,
                     float* xn, int64_t ld_xn, float* sqnorm, float* norm,
                     float* hi, float* lo, uint16_t* bf,
                     uint16_t* h_hi, uint16_t* h_lo, float* h_scale_inv, int64_t Dp, void* stream);

/* ---- distance matrix --------------------------------------------------------------------------
 * Replaces euclidean_distance (utils/metrics.py:7-13), cosine_similarity (utils/metrics.py:15-25),
 * the inline 1-q.g (processor_uniprompt_stage2.py:466-468) and the all-pairs matrix of
 * utils/reranking.py:36-41.
 *   precision MPREID_FP32_SIMT : qa/ga = fp32 features [*, ldk]; qb/gb ignored
 *   precision MPREID_3XTF32    : qa/ga = hi planes, qb/gb = lo planes, [*, ldk] fp32, ldk % 32 == 0
 *   precision MPREID_BF16      : qa/ga = bf16 planes [*, ldk], ldk % 64 == 0;    qb/gb ignored
 *   precision MPREID_3XFP16    : qa/ga = fp16 hi planes, qb/gb = fp16 lo planes, [*, ldk], ldk % 64 == 0,
 *                                q_scale/g_scale = the per-row 2^-s of mpreid_prep_rows (NULL otherwise)
 *   q_aux/g_aux: squared norms for the euclidean metrics, norms for MPREID_ARCCOS, unused (may be
 *   NULL) for MPREID_ONE_MINUS_DOT.
 *   row_max (optional, [Q], must be pre-filled with -inf): per-row maximum of the written values
 *   (the column max of utils/reranking.py:46, by symmetry).                                    */
MPREID_API int mpreid_dist_matrix(const void* qa, const void* qb, const void* ga, const void* gb,
                       const float* q_aux, const float* g_aux, const float* q_scale, const float* g_scale,
                       int64_t Q, int64_t G, int64_t K, int64_t ldk,
                       int metric, int precision,
                       float* out, int64_t ld_out, float* row_max, void* stream);

/* All-pairs flavour (utils/reranking.py:36-41: the stacked features against themselves).  Only the
 * tiles on or right of the diagonal are contracted; each of them also stores its transpose, so the
 * matrix is exactly symmetric and the tensor-core work is halved.  row_max as above.             */
MPREID_API int mpreid_dist_matrix_symmetric(const void* xa, const void* xb, const float* x_aux, const float* x_scale,
                                 int64_t N, int64_t K, int64_t ldk, int metric, int precision,
                                 float* out, int64_t ld_out, float* row_max, void* stream);

/* Fused flavour of the all-pairs launch for re-ranking: utils/reranking.py:36-48 WITHOUT the (Q+G)^2 matrix in HBM.
 * Squared-euclidean, symmetric tiles as above, but nothing is stored except
 *   out_qg [Q, ld_out]  the query-to-gallery block the lambda blend needs (:95); element (q, g) sits at column
 *                       (Q & 31) + g, so that 32-column groups keep their 128-byte alignment; ld_out >= (Q & 31) + N-Q;
 *   row_max [N]         (pre-filled with -inf) the column maxima of :46;
 *   cand [N, cand_cap], cand_cnt [N] (zeroed by the caller): every element d(i, j) <= thr[i] is appended to the
 *                       candidate list of row i as (j << 32 | fp32 bits), in no particular order (both orientations of a
 *                       mirrored tile append).  cand_cnt keeps counting past cand_cap.
 * thr [N]: per-sample raw-domain thresholds, e.g. the (k+2)-th smallest distance to a sample of the columns plus a few
 * ulps; mpreid_cand_topk then selects the first k of np.argsort(row / row_max, kind='stable') from the lists and
 * reports in status[0] how many rows could NOT be decided from their list (overflow, fewer than k entries, or a
 * possible tie with an element outside the list) -- the caller falls back to the materialising calls if it is not 0.
 * status (device, int32[4]): [0] undecided rows, [1] longest list.                                              */
MPREID_API int mpreid_dist_symmetric_topk(const void* xa, const void* xb, const float* x_sqnorm, const float* x_scale,
                               int64_t N, int64_t K, int64_t ldk, int precision,
                               const float* thr, uint64_t* cand, int32_t* cand_cnt, int64_t cand_cap,
                               int64_t Q, float* out_qg, int64_t ld_out, float* row_max, int own_mod, int own_rank, void* stream);
/* keys (optional, [N, k] uint64): PARTIAL mode for one rank of a multi-GPU run -- the k smallest sort keys
 * (ordered fp32 bits of value / row_scale, then the column) of what THIS rank's tiles contributed to each row, ~0 = none;
 * idx / val may then be NULL and only list overflow is reported.  mpreid_merge_topk merges the all-gathered keys_all
 * [P, N, k] of all ranks into idx / val and validates as above.
 * own_mod / own_rank of mpreid_dist_symmetric_topk: the launch contracts only the tiles whose 256-row block p satisfies
 * p % own_mod == own_rank (1 / 0 = all): the row-sharded multi-GPU form; out_qg then holds the rows of those blocks only. */
MPREID_API int mpreid_cand_topk(const uint64_t* cand, const int32_t* cand_cnt, int64_t cand_cap, int64_t N, int k,
                     const float* row_scale, const float* thr, int32_t* idx, float* val, uint64_t* keys, int32_t* status, void* stream);
MPREID_API int mpreid_merge_topk(const uint64_t* keys_all, int P, int64_t N, int k, const float* row_scale, const float* thr,
                      int32_t* idx, float* val, int32_t* status, void* stream);

/* ---- ranking + CMC / AP ------------------------------------------------------------------------
 * Replaces eval_func (utils/metrics.py:28-88).  The Q x G argsort is never formed: for every query
 * the kernel takes the gallery entries with the query's pid (a hash-grouped label index), sorts
 * those few (distance, index) keys and counts, in ONE streaming pass over the distance row, how
 * many entries precede each of them under the stable order (distance, then gallery index) ==
 * np.argsort(kind='stable').  Outputs per query:
 *   first_hit [Q] int32  1-based rank (after junk removal) of the first correct match, 0 = the
 *                        query's pid is absent from the (kept) gallery            (:61-63)
 *   ap        [Q] fp64   average precision, bit-equal to the float64 numpy expression of :73-79
 *                        (numpy's pairwise summation order is emulated)
 *   num_rel   [Q] int32  number of correct matches                                 (:73)
 * The final cmc = float32 sum / num_valid and mAP = np.mean (:84-86) are left to the host so that
 * they are computed by numpy itself on the gathered per-query values (also across GPUs).
 * status_host semantics: the call is asynchronous; `status` (device, int32[4]) receives
 *   [0] overflow flag (workspace `pos_capacity` too small), [1] entries needed, [2] max positives
 *   of any query, [3] reserved.                                                                  */
MPREID_API size_t mpreid_rank_eval_workspace_bytes(int64_t Q, int64_t G, int64_t pos_capacity);
MPREID_API int mpreid_rank_eval(const float* dist, int64_t ld_dist, int64_t Q, int64_t G,
                     const int64_t* q_pid, const int64_t* g_pid,
                     const int64_t* q_cam, const int64_t* g_cam, int junk_mode,
                     int32_t* first_hit, double* ap, int32_t* num_rel,
                     void* workspace, size_t workspace_bytes, int64_t pos_capacity,
                     int32_t* status, void* stream);

/* The same from FEATURES in one call (what R1_mAP_eval.compute() runs between torch.cat and eval_func,
 * utils/metrics.py:111-132): normalise (optional) + operand planes of both sides, distance matrix, ranking + CMC / AP,
 * and optionally topk_idx [Q, topk] = the first topk columns of np.argsort(distmat, axis=1, kind='stable') (:39).
 * qf [Q, D], gf [G, D] fp32.  The matrix goes to dist_out [Q, ld_dist_out] if given, else it lives in the workspace
 * (size query: own_dist = 1).  status as for mpreid_rank_eval.                                                     */
MPREID_API size_t mpreid_eval_features_workspace_bytes(int64_t Q, int64_t G, int64_t D, int precision, int64_t pos_capacity, int own_dist);
MPREID_API int mpreid_eval_features(const float* qf, int64_t ld_q, const float* gf, int64_t ld_g, int64_t Q, int64_t G, int64_t D,
                         int normalize, int metric, int precision,
                         const int64_t* q_pid, const int64_t* g_pid, const int64_t* q_cam, const int64_t* g_cam, int junk_mode,
                         int32_t* first_hit, double* ap, int32_t* num_rel,
                         int32_t* topk_idx, int topk, float* dist_out, int64_t ld_dist_out,
                         void* workspace, size_t workspace_bytes, int64_t pos_capacity, int32_t* status, void* stream);

/* ---- per-row top-k ---------------------------------------------------------------------------
 * The first k entries of np.argsort(row / row_scale, kind='stable') (utils/reranking.py:46-48 with
 * k = k1+1; retrieval top-100).  row_scale may be NULL (no division).  idx [Q,k] int32, val [Q,k]
 * fp32 (the divided values; may be NULL).  Entries beyond min(k,G) are -1 / +inf.              */
MPREID_API int mpreid_row_topk(const float* dist, int64_t ld_dist, int64_t Q, int64_t G, int k,
                    const float* row_scale, int32_t* idx, float* val, void* stream);
MPREID_API int mpreid_row_max(const float* dist, int64_t ld_dist, int64_t Q, int64_t G, float* row_max, void* stream);
/* out[r] = the t-th smallest value (1-based, numpy sort order) of row r of a short-row matrix [R, S], S <= 4096: the
 * per-row thresholds of the fused all-pairs pass (np.partition(row, t-1)[t-1]).                                      */
MPREID_API int mpreid_row_kth(const float* dist, int64_t ld_dist, int64_t R, int64_t S, int t, float* out, void* stream);
/* A cheap upper bound of the same (t <= 128): the t-th smallest of a subset of the row (every lane of a warp keeps the 4 to
 * 16 smallest of its strided share), i.e. >= the exact value and usually equal to it.  What the thresholds need: at least
 * t elements of the row are <= out[r].                                                                               */
MPREID_API int mpreid_row_kth_bound(const float* dist, int64_t ld_dist, int64_t R, int64_t S, int t, float* out, void* stream);

/* ---- k-reciprocal re-ranking (utils/reranking.py:29-100) --------------------------------------
 * `dist` is the all-pairs matrix in the orientation dist[i][j] = reference distmat[j][i] (:46
 * transposes), N = Q + G rows.  One call runs: row max + top-(k1+1) (:46-48), k-reciprocal sets
 * with the 2/3 expansion rule and the Gaussian-kernel V rows in fp16 (:51-71), k2 query expansion
 * (:73-78), inverted index (:80-82), Jaccard distance with the fp16 accumulator (:84-93) and the
 * lambda blend (:95), writing final[Q, G] fp32 (:99).  fp16 rounding points are the reference's.
 * row_max_in (optional, [N]): the per-row maxima if the caller already has them (mpreid_dist_matrix
 * produces them in its epilogue); NULL = computed here with one more pass over `dist`.            */
MPREID_API size_t mpreid_rerank_workspace_bytes(int64_t N, int64_t Q, int k1, int k2);
/* The same pipeline in stages, for row-sharded multi-GPU runs (the one exchange step of this path:
 * neighbour lists and V0 rows are all-gathered between the stages):
 *   1. per rank, on its block of rows of the all-pairs matrix: mpreid_row_max / mpreid_row_topk
 *      with k = mpreid_rerank_neighbor_count(k1, k2)                                    (:46-48)
 *   2. mpreid_rerank_build_v0: V0 rows (ELL, capacity mpreid_rerank_v0_capacity per row) of the
 *      rank's rows; row_ids[R] = global sample index of every local row (NULL = 0..R-1); needs the
 *      neighbour lists of ALL samples                                                    (:51-71)
 *   3. mpreid_rerank_finish: query expansion + inverted index over all samples (cheap, done
 *      redundantly on every rank), Jaccard + blend for the rank's Qs query rows: dist_qrows[Qs, N]
 *      are those rows of the all-pairs matrix, q_ids[Qs] their global indices (NULL = 0..Qs-1),
 *      row_max_q[Qs] their maxima; writes final[Qs, N-Q]                                (:73-99)   */
MPREID_API int mpreid_rerank_neighbor_count(int k1, int k2);
MPREID_API int mpreid_rerank_v0_capacity(int k1, int64_t N);
MPREID_API int mpreid_rerank_build_v0(const float* dist_rows, int64_t ld_dist, const int32_t* row_ids, int64_t R, int64_t N,
                           int k1, const int32_t* nbr_all, int K, const float* row_max_rows,
                           int32_t* v0_col, uint16_t* v0_val, int32_t* v0_len, void* stream);
/* mpreid_rerank_build_v0 without matrix rows (fused all-pairs pass): original_dist[i, idx] (:70) is the stored neighbour
 * value nbr_val_all[i, m] (already divided by the row maximum; the same accumulator the matrix would hold) when idx is
 * the m-th neighbour of i, and (|x_i|^2 + |x_idx|^2 - 2 x_i.x_idx) / row_max in fp32 from the feature rows xn [N, ld_xn]
 * for the few expansion members outside the neighbour list.                                                      */
MPREID_API int mpreid_rerank_build_v0_sparse(const int32_t* row_ids, int64_t R, int64_t N, int k1, const int32_t* nbr_all,
                                  const float* nbr_val_all, int K, const float* row_max_rows,
                                  const float* xn, int64_t ld_xn, int64_t D, const float* sqnorm,
                                  int32_t* v0_col, uint16_t* v0_val, int32_t* v0_len, void* stream);
MPREID_API size_t mpreid_rerank_finish_workspace_bytes(int64_t N, int64_t Q, int k1, int k2);
MPREID_API int mpreid_rerank_finish(const int32_t* nbr_all, int K, const int32_t* v0_col, const uint16_t* v0_val, const int32_t* v0_len,
                         const float* dist_qrows, int64_t ld_dist, const int32_t* q_ids, const float* row_max_q,
                         int64_t N, int64_t Q, int64_t Qs, int k1, int k2, float lambda_value,
                         float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes, void* stream);
/* General form of mpreid_rerank_finish: gallery sample 0 sits at column col0 of dist_q (col0 = Q for rows of the
 * all-pairs matrix; the pad (Q & 31) for the [Q, G] block mpreid_dist_symmetric_topk keeps; ld_dist >= col0 + N-Q),
 * and the parts may run as separate calls on the same workspace: stages is a mask of 1 = query expansion + inverted index
 * (:73-82), 2 = sparse Jaccard accumulation and blend of the touched entries (:84-95), 4 = the dense default blend (:95
 * with temp_min = 0; needs only the distance block and the maxima, so it may run early on another stream; 2 must come
 * after 4); 7 = everything.  Sharded query expansion: 8 = expand only the rows [qe_lo, qe_hi) into the workspace (the caller
 * all-gathers the other rows into the arrays mpreid_rerank_finish_layout locates), 16 = build the inverted index from a
 * complete V.
 * v0_stride: row stride of v0_col / v0_val in entries (0 = mpreid_rerank_v0_capacity; a sharded run all-gathers the V0
 * rows trimmed to their longest length).  rows_global != 0: dist_q and row_max_q are addressed by the GLOBAL query index
 * q_ids[il] (the [Q, .] block and the [N] maxima a rank holds in the row-sharded form) instead of the local row il.   */
MPREID_API int mpreid_rerank_finish_ex(const int32_t* nbr_all, int K, const int32_t* v0_col, const uint16_t* v0_val, const int32_t* v0_len,
                            const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* q_ids, const float* row_max_q,
                            int64_t N, int64_t Q, int64_t Qs, int k1, int k2, float lambda_value,
                            float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes, int stages,
                            int64_t v0_stride, int rows_global, int64_t qe_lo, int64_t qe_hi, void* stream);
/* byte offsets of v_col int32 [N, C1], v_val fp16 [N, C1], v_len int32 [N] inside the finish workspace, and C1 (k2 > 1) */
MPREID_API int mpreid_rerank_finish_layout(int64_t N, int64_t Q, int k1, int k2, int64_t* out4);
/* Stage 4 alone: final[il, c] = fp16(1 - lambda) + lambda * dist_q[row(il), col0 + c] / row_max_q[row(il)], row(il) =
 * src_rows[il] if given (global addressing) else il.  Needs nothing from the sparse stages.  ctas_per_sm > 0 caps the
 * launch to that many 256-thread CTAs per SM (a background launch next to latency-bound kernels); 0 = full occupancy. */
MPREID_API int mpreid_rerank_blend_default(const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* src_rows, const float* row_max_q,
                                int64_t Qs, int64_t G, float lambda_value, float* final_dist, int64_t ld_final, int ctas_per_sm, void* stream);
MPREID_API int mpreid_rerank(const float* dist, int64_t ld_dist, const float* row_max_in, int64_t N, int64_t Q, int k1, int k2,
                  float lambda_value, float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes,
                  int32_t* status, void* stream);

/* ---- batch-hard mining (forward only) --------------------------------------------------------------
 * loss/triplet_loss.py:50-103 on an [N, N] distance matrix (e.g. MPREID_SQRT_EUCLID of a batch against
 * itself): dist_ap[i] = max over same-label j (the diagonal included), dist_an[i] = min over other-label
 * j; p_inds / n_inds (optional) are the absolute column indices, first index on ties.  A row without
 * other-label entries gets dist_an = +inf, n_inds = -1.                                              */
MPREID_API int mpreid_hard_example_mining(const float* dist, int64_t ld_dist, int64_t N, const int64_t* labels,
                               float* dist_ap, float* dist_an, int64_t* p_inds, int64_t* n_inds, void* stream);

/* ---- batch-hard triplet distances with a backward (loss/triplet_loss.py:16-31,50-103) ------------------
 * forward: x [B, D] fp32 -> the pairwise sqrt/clamp euclidean distances of the batch against itself, mined in the same
 * kernel: dist_ap[i] = max over same-label j (diagonal included), dist_an[i] = min over other-label j (+inf, n_inds = -1
 * if there is none), p_inds / n_inds = the selected columns (lowest index on ties).
 * backward: grad_x [B, D] = d(sum_i g_ap[i] dist_ap[i] + g_an[i] dist_an[i]) / dx, zero through an active clamp;
 * gathered per row in a fixed order (no atomics: bit-reproducible).                                                */
MPREID_API int mpreid_triplet_forward(const float* x, int64_t ld_x, int64_t B, int64_t D, const int64_t* labels,
                           float* dist_ap, float* dist_an, int64_t* p_inds, int64_t* n_inds, void* stream);
MPREID_API int mpreid_triplet_backward(const float* x, int64_t ld_x, int64_t B, int64_t D, const int64_t* p_inds, const int64_t* n_inds,
                            const float* dist_ap, const float* dist_an, const float* g_ap, const float* g_an,
                            float* grad_x, int64_t ld_g, void* stream);

/* ---- stage-1 contrastive step on cached features (loss/supcontrast.py:17-31, called twice by
 * processor/processor_uniprompt_stage1.py:88-93) ----------------------------------------------------
 * S [Ba, Bb] = A.B^T (mpreid_dist_matrix with MPREID_DOT).  loss[0] = SupConLoss(A, B, labels_a, labels_b) (rows of S),
 * and with both_directions: loss[1] = SupConLoss(B, A, ...) (columns of S), loss[2] = their sum.  grad_a / grad_b
 * (optional, [Ba, D] / [Bb, D]) = d(grad_scale_rows * loss[0] + grad_scale_cols * loss[1]) / dA, dB.                */
MPREID_API size_t mpreid_supcon_workspace_bytes(int64_t Ba, int64_t Bb);
MPREID_API int mpreid_supcon_step(const float* S, int64_t ld_s, int64_t Ba, int64_t Bb, const int64_t* labels_a, const int64_t* labels_b,
                       float temperature, int both_directions, float grad_scale_rows, float grad_scale_cols,
                       const float* a, int64_t ld_a, const float* b, int64_t ld_b, int64_t D,
                       float* loss, float* grad_a, int64_t ld_ga, float* grad_b, int64_t ld_gb,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU surface for non-PyTorch consumers ---------------------------------------------
 * Thin wrappers over NCCL (resolved at run time: dlopen("libnccl.so.2"), or the path in MPREID_NCCL_LIB), one
 * communicator per process / GPU.  The path needs three collectives: the gallery broadcast, all-gathers (per-query
 * results, neighbour lists, partial top-k keys, V0 rows) and a max all-reduce (row maxima of the sharded all-pairs pass).
 * All buffers are device pointers; the calls are asynchronous on `stream`.
 *   rank 0: mpreid_comm_unique_id(id) -> ship the 128 bytes to the other ranks out of band -> every rank, with ITS GPU
 *   current: mpreid_comm_init(&comm, world, rank, id).  mpreid_comm_from_nccl adopts an existing ncclComm_t instead
 *   (not destroyed by mpreid_comm_destroy).                                                                          */
typedef struct mpreid_comm mpreid_comm;
MPREID_API int mpreid_comm_unique_id(void* id_out_128_bytes);
MPREID_API int mpreid_comm_init(mpreid_comm** comm, int world, int rank, const void* unique_id_128_bytes);
MPREID_API int mpreid_comm_from_nccl(mpreid_comm** comm, void* nccl_comm, int world, int rank);
MPREID_API int mpreid_comm_size(const mpreid_comm* comm, int* world, int* rank);
MPREID_API int mpreid_comm_broadcast(mpreid_comm* comm, void* buf, size_t bytes, int root, void* stream);
MPREID_API int mpreid_comm_allgather(mpreid_comm* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream);
MPREID_API int mpreid_comm_allreduce_max_f32(mpreid_comm* comm, float* buf, size_t count, void* stream);
MPREID_API int mpreid_comm_destroy(mpreid_comm* comm);

/* ---- host-side hooks (no GPU needed) -----------------------------------------------------------
 * The scalar arithmetic the kernels run is __host__ __device__ code; these two entry points run it
 * on the CPU so the CPU-only test-suite can check it bit-for-bit against numpy.
 *   mpreid_host_average_precision: AP of one query from the ascending 1-based ranks of its m
 *     correct matches in a kept list of n entries (utils/metrics.py:73-79, numpy pairwise order).
 *   mpreid_host_order_keys: the 32-bit sort key of each fp32 value (numpy sort order).             */
MPREID_API double mpreid_host_average_precision(const int32_t* ranks_host, int m, int64_t n);
MPREID_API void mpreid_host_order_keys(const float* values_host, int64_t n, uint32_t* keys_host);

#ifdef __cplusplus
}
#endif
#endif /* MPREID_B200_H */
