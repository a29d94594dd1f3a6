// Per-row top-k == the first k entries of np.argsort(row / scale, kind='stable')
// (utils/reranking.py:46-48 with k = k1+1; the retrieval top-100 of BASELINE config 5).
//
// One CTA streams one row ONCE.  Candidates (64-bit keys = ordered fp32 bits of the divided value,
// then the column index, so ties resolve to the lower index exactly like a stable sort) are appended
// to a shared-memory buffer if they beat the current k-th best; when the buffer fills it is sorted
// (bitonic) and cut back to k, which tightens the threshold.  After the first cut almost nothing
// passes the filter, so the kernel is a pure HBM stream: 4 B per element, read once.
//
// The division by the row scale (the reference normalises BEFORE sorting, which can create ties) is
// exact but lazy: v/r is monotone in v for r > 0, so a raw-domain bound B = max{x : fl(x/r) <= t}
// rejects with one compare and only survivors pay the IEEE division.
#include "common.cuh"

namespace mpreid {

static constexpr int kTopkThreads = 256;
static constexpr int kTopkWarps = kTopkThreads / 32;
static constexpr int kTopkChunk = 32 * 16;  // elements one warp classifies per trip (4 x float4 per lane)
static constexpr int kTopkMaxK = 2048;

__device__ __forceinline__ float4 ldg_stream4_topk(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// largest raw x with fl(x / r) <= t  (r > 0, t finite); +inf ("no filtering") when no safe bound is found.
// x = t*r is within an ulp or two of the answer, so a few nextafter steps settle it.
__device__ float raw_bound(float t, float r, bool has_scale) {
  if (!has_scale) return t;
  if (!(r > 0.f) || !(fabsf(t) <= 3.0e38f)) return INFINITY;
  float x = t * r;
  if (!(fabsf(x) <= 3.0e38f)) return INFINITY;
  int it = 0;
  for (; it < 4 && x / r > t; ++it) x = nextafterf(x, -INFINITY);
  if (x / r > t) return INFINITY;
  for (it = 0; it < 4; ++it) {
    const float xn = nextafterf(x, INFINITY);
    if (xn / r <= t) x = xn; else return x;
  }
  return INFINITY;  // did not converge: a bound that is too small would drop candidates
}

// ---- warp-level selection over a shared-memory array of unique 64-bit keys (no block barriers) ----
struct WarpSel {
  uint64_t* q;      // [cap] candidate buffer of this warp
  uint32_t* hist;   // [256]
  int lane;
};

// k-th smallest (1-based) of q[0..count): MSB-first radix select, 8 bits per pass, early exit once the
// selected bin holds a single key
__device__ uint64_t warp_select_kth(const WarpSel& w, int count, int k) {
  uint64_t prefix = 0, mask = 0;
  uint32_t want = (uint32_t)k;
  for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
    for (int b = 0; b < 8; ++b) w.hist[w.lane * 8 + b] = 0;
    __syncwarp();
    for (int e = w.lane; e < count; e += 32) {
      const uint64_t key = w.q[e];
      if ((key & mask) == prefix) atomicAdd(&w.hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c[b] = w.hist[w.lane * 8 + b]; sum += c[b]; }
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (w.lane >= o) incl += y; }
    const uint32_t excl = incl - sum;
    const bool mine = want > excl && want <= incl;
    uint32_t bin = 0, rest = 0, bcount = 0;
    if (mine) {
      uint32_t run = excl;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (want > run && want <= run + c[b]) { bin = w.lane * 8 + b; rest = want - run; bcount = c[b]; }
        run += c[b];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    bin = __shfl_sync(0xffffffffu, bin, src);
    want = __shfl_sync(0xffffffffu, rest, src);
    bcount = __shfl_sync(0xffffffffu, bcount, src);
    prefix |= (uint64_t)bin << shift;
    mask |= (uint64_t)255u << shift;
    __syncwarp();
    if (bcount == 1 && shift > 0) {   // the only key with this prefix is the answer
      uint64_t found = 0;
      for (int e0 = 0; e0 < count; e0 += 32) {
        const int e = e0 + w.lane;
        const uint64_t key = e < count ? w.q[e] : ~0ull;
        const bool hit = e < count && (key & mask) == prefix;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) { found = __shfl_sync(0xffffffffu, key, __ffs(bal) - 1); break; }
      }
      return found;
    }
  }
  return prefix;
}

// keep the entries of q[from..count) for which pred(key) holds, compacted (in order) to q[to..]; returns the new count
template <class Pred>
__device__ int warp_compact(const WarpSel& w, int from, int count, int to, Pred pred) {
  int out = to;
  for (int e0 = from; e0 < count; e0 += 32) {
    const int e = e0 + w.lane;
    uint64_t key = e < count ? w.q[e] : 0ull;
    bool keep = e < count;
    if (keep) keep = pred(key);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) w.q[out + __popc(bal & ((1u << w.lane) - 1u))] = key;
    out += __popc(bal);
    __syncwarp();
  }
  return out;
}

__device__ void warp_bitonic(const WarpSel& w, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = w.lane; i < n_pow2; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t x = w.q[i], y = w.q[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { w.q[i] = y; w.q[ixj] = x; }
        }
      }
      __syncwarp();
    }
  }
}

// One WARP per row, no block-level barrier anywhere.  q[0..nkeys) holds survivors as sort keys
// (ordered value bits << 32 | column); q[nkeys..count) holds raw entries (column << 32 | float bits)
// appended by the streaming loop.  cut(): raw -> key (exact IEEE division by the row scale), drop what
// no longer beats the threshold, and if more than k remain select the k smallest.
struct RowState {
  int count, nkeys;
  uint64_t thr;   // current k-th best key, ~0 = none yet
  float bound;    // raw-domain rejection bound
};

__device__ void warp_cut(const WarpSel& w, RowState& st, int k, float r, bool has_scale) {
  const uint64_t thr = st.thr;
  // convert + filter the raw tail in place
  int out = st.nkeys;
  for (int e0 = st.nkeys; e0 < st.count; e0 += 32) {
    const int e = e0 + w.lane;
    uint64_t key = 0;
    bool keep = false;
    if (e < st.count) {
      const uint64_t raw = w.q[e];
      const float v = __uint_as_float((uint32_t)(raw & 0xffffffffu));
      key = make_key(has_scale ? v / r : v, (uint32_t)(raw >> 32));
      keep = key < thr;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) w.q[out + __popc(bal & ((1u << w.lane) - 1u))] = key;
    out += __popc(bal);
    __syncwarp();
  }
  st.count = out;
  if (st.count > k) {
    const uint64_t kth = warp_select_kth(w, st.count, k);
    st.count = warp_compact(w, 0, st.count, 0, [kth](uint64_t key) { return key <= kth; });
    st.thr = kth;
    const uint32_t hi = (uint32_t)(kth >> 32);
    float b = INFINITY;
    if (w.lane == 0 && hi != 0xffffffffu) b = raw_bound(order_key_inv(hi), r, has_scale);
    st.bound = __shfl_sync(0xffffffffu, b, 0);
  }
  st.nkeys = st.count;
}

__global__ void __launch_bounds__(kTopkThreads)
k_row_topk(const float* __restrict__ dist, int64_t ld, int Q, int G, int k, const float* __restrict__ row_scale,
           int32_t* __restrict__ idx_out, float* __restrict__ val_out, int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;   // 8 for the usual small k, fewer when the per-warp buffers grow with k
  WarpSel w;
  w.q = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * cap;
  w.hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)warps * cap * 8) + warp * 256;
  w.lane = lane;
  const int keff = min(k, G);
  const int drain_at = cap - kTopkChunk;   // cut before a full chunk could overflow the buffer
  const bool has_scale = row_scale != nullptr;
  for (int row_i = blockIdx.x * warps + warp; row_i < Q; row_i += gridDim.x * warps) {
    const float* row = dist + (int64_t)row_i * ld;
    const float r = has_scale ? row_scale[row_i] : 1.0f;
    RowState st;
    st.count = 0; st.nkeys = 0; st.thr = ~0ull; st.bound = INFINITY;

    int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
    head = min(head, G);
    const int nvec = (G - head) >> 2;
    const float4* rv = reinterpret_cast<const float4*>(row + head);
    {  // ragged ends (< 4 elements each)
      uint32_t mask = 0; float v[2] = {0.f, 0.f}; uint32_t jj[2] = {0u, 0u};
      if (lane < head) { v[0] = row[lane]; jj[0] = lane; mask |= 1u; }
      const int t = head + 4 * nvec + lane;
      if (t < G) { v[1] = row[t]; jj[1] = t; mask |= 2u; }
      const int n = __popc(mask);
      int incl = n;
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      int pos = st.count + incl - n;
      if (mask & 1u) w.q[pos++] = ((uint64_t)jj[0] << 32) | __float_as_uint(v[0]);
      if (mask & 2u) w.q[pos++] = ((uint64_t)jj[1] << 32) | __float_as_uint(v[1]);
      st.count += __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();
    }
    constexpr int U = 4;
    for (int v0 = 0; v0 < nvec; v0 += 32 * U) {
      float4 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * 32 + lane;
        x[u] = v < nvec ? ldg_stream4_topk(rv + v) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
      }
      const float bound = st.bound;
      uint32_t mask = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        mask |= (!(x[u].x > bound) ? 1u : 0u) << (4 * u);
        mask |= (!(x[u].y > bound) ? 2u : 0u) << (4 * u);
        mask |= (!(x[u].z > bound) ? 4u : 0u) << (4 * u);
        mask |= (!(x[u].w > bound) ? 8u : 0u) << (4 * u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v0 + u * 32 + lane >= nvec) mask &= ~(0xfu << (4 * u));
      const int n = __popc(mask);
      int incl = n;
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total) {
        int pos = st.count + incl - n;
        if (mask) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint64_t j = (uint64_t)(head + 4 * (v0 + u * 32 + lane));
            if (mask & (1u << (4 * u))) w.q[pos++] = (j << 32) | __float_as_uint(x[u].x);
            if (mask & (2u << (4 * u))) w.q[pos++] = ((j + 1) << 32) | __float_as_uint(x[u].y);
            if (mask & (4u << (4 * u))) w.q[pos++] = ((j + 2) << 32) | __float_as_uint(x[u].z);
            if (mask & (8u << (4 * u))) w.q[pos++] = ((j + 3) << 32) | __float_as_uint(x[u].w);
          }
        }
        st.count += total;
        __syncwarp();
        // cut when another chunk could overflow the buffer (the first cut comes after max(2k, 256) candidates)
        if (st.count > drain_at) warp_cut(w, st, keff, r, has_scale);
      }
    }
    warp_cut(w, st, keff, r, has_scale);
    const int P = (int)next_pow2_u32((uint32_t)max(st.count, 1));
    for (int i = st.count + lane; i < P; i += 32) w.q[i] = ~0ull;
    __syncwarp();
    warp_bitonic(w, P);
    for (int i = lane; i < k; i += 32) {
      int32_t id = -1;
      float val = INFINITY;
      if (i < keff) {
        id = (int32_t)(w.q[i] & 0xffffffffu);
        val = has_scale ? row[id] / r : row[id];
      }
      idx_out[(int64_t)row_i * k + i] = id;
      if (val_out) val_out[(int64_t)row_i * k + i] = val;
    }
    __syncwarp();
  }
}

// The same selection over the CANDIDATE LISTS the fused all-pairs pass leaves behind (dist_tc.cu, TopkFuse): row i holds
// min(cnt[i], cap) unordered entries (column << 32 | fp32 bits) -- every element of the row that is not above thr[i].
// One warp per row.  status[0] counts rows whose result cannot be trusted (the caller then falls back to the
// materialising path): the list overflowed, it holds fewer than k entries, or the k-th selected value is not strictly
// below fl(thr / scale), in which case an element outside the list could tie with it.
__global__ void __launch_bounds__(kTopkThreads)
k_cand_topk(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int64_t cand_cap, int N, int k,
            const float* __restrict__ row_scale, const float* __restrict__ thr,
            int32_t* __restrict__ idx_out, float* __restrict__ val_out, unsigned long long* __restrict__ keys_out, int32_t* status, int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  WarpSel w;
  w.q = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * cap;
  w.hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)warps * cap * 8) + warp * 256;
  w.lane = lane;
  const int keff = min(k, N);
  const int drain_at = cap - kTopkChunk;
  const bool has_scale = row_scale != nullptr;
  const bool partial = keys_out != nullptr;   // one rank's share of the row: fewer than k entries is normal, the merge validates
  for (int row_i = blockIdx.x * warps + warp; row_i < N; row_i += gridDim.x * warps) {
    const float r = has_scale ? row_scale[row_i] : 1.0f;
    const int cnt_all = cand_cnt[row_i];
    const int cnt = (int)min((int64_t)cnt_all, cand_cap);
    const unsigned long long* src = cand + (int64_t)row_i * cand_cap;
    RowState st;
    st.count = 0; st.nkeys = 0; st.thr = ~0ull; st.bound = INFINITY;
    for (int e0 = 0; e0 < cnt; e0 += kTopkChunk) {
      const int n = min(kTopkChunk, cnt - e0);
      for (int e = lane; e < n; e += 32) w.q[st.count + e] = src[e0 + e];
      st.count += n;
      __syncwarp();
      if (st.count > drain_at) warp_cut(w, st, keff, r, has_scale);
    }
    warp_cut(w, st, keff, r, has_scale);
    const int P = (int)next_pow2_u32((uint32_t)max(st.count, 1));
    for (int i = st.count + lane; i < P; i += 32) w.q[i] = ~0ull;
    __syncwarp();
    warp_bitonic(w, P);
    bool bad = cnt_all > cand_cap || (!partial && st.count < keff);
    if (!bad && !partial && keff > 0) {
      const float t = has_scale ? thr[row_i] / r : thr[row_i];
      bad = !((uint32_t)(w.q[keff - 1] >> 32) < order_key(t));
    }
    if (bad && lane == 0) atomicAdd(&status[0], 1);
    if (lane == 0) atomicMax(&status[1], cnt_all);
    if (partial) {
      for (int i = lane; i < k; i += 32) keys_out[(int64_t)row_i * k + i] = (i < keff && i < st.count) ? w.q[i] : ~0ull;
      __syncwarp();
      continue;
    }
    for (int i = lane; i < k; i += 32) {
      int32_t id = -1;
      float val = INFINITY;
      if (i < keff && i < st.count) {
        id = (int32_t)(w.q[i] & 0xffffffffu);
        val = order_key_inv((uint32_t)(w.q[i] >> 32));   // the divided value itself (the key is a bijection on non-NaN floats)
      }
      idx_out[(int64_t)row_i * k + i] = id;
      if (val_out) val_out[(int64_t)row_i * k + i] = val;
    }
    __syncwarp();
  }
}

// Merge of per-rank partial selections (row-sharded multi-GPU re-ranking): keys_all [P, N, k] holds, per rank, the k smallest
// sort keys of the row among the tiles that rank contracted (~0 = none).  The union contains the k smallest of the whole
// row; one warp per row sorts the P * k keys and validates like k_cand_topk (k valid keys, k-th strictly below thr).
static constexpr int kMergeMax = 1024;
__global__ void __launch_bounds__(128)
k_merge_topk(const unsigned long long* __restrict__ keys_all, int P, int N, int k, const float* __restrict__ row_scale,
             const float* __restrict__ thr, int32_t* __restrict__ idx_out, float* __restrict__ val_out, int32_t* status) {
  __shared__ uint64_t sbuf[4][kMergeMax];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSel w;
  w.q = sbuf[warp]; w.hist = nullptr; w.lane = lane;
  const int keff = min(k, N);
  const int T = P * k;
  const int Pw = (int)next_pow2_u32((uint32_t)T);
  for (int row_i = blockIdx.x * 4 + warp; row_i < N; row_i += gridDim.x * 4) {
    for (int t = lane; t < Pw; t += 32) {
      uint64_t key = ~0ull;
      if (t < T) { const int p = t / k, e = t - p * k; key = keys_all[((int64_t)p * N + row_i) * k + e]; }
      w.q[t] = key;
    }
    __syncwarp();
    warp_bitonic(w, Pw);
    const float r = row_scale ? row_scale[row_i] : 1.0f;
    bool bad = false;
    if (keff > 0) {
      const uint64_t kth = w.q[keff - 1];
      const float t = row_scale ? thr[row_i] / r : thr[row_i];
      bad = kth == ~0ull || !((uint32_t)(kth >> 32) < order_key(t));
    }
    if (bad && lane == 0) atomicAdd(&status[0], 1);
    for (int i = lane; i < k; i += 32) {
      int32_t id = -1; float val = INFINITY;
      if (i < keff && w.q[i] != ~0ull) { id = (int32_t)(w.q[i] & 0xffffffffu); val = order_key_inv((uint32_t)(w.q[i] >> 32)); }
      idx_out[(int64_t)row_i * k + i] = id;
      val_out[(int64_t)row_i * k + i] = val;
    }
    __syncwarp();
  }
}

// t-th smallest value (1-based) of every row of a SHORT-row matrix [R, S], S <= 32 * PER: the per-row thresholds of the
// fused all-pairs pass (distances to 2,048 sampled columns).  One warp per row, the row's sort keys in registers, the
// answer found bit by bit (the largest prefix with fewer than t keys below it): 32 rounds of PER compares and one warp sum.
template <int PER>
__global__ void __launch_bounds__(256)
k_row_kth(const float* __restrict__ m, int64_t ld, int R, int S, int t, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < R; row += gridDim.x * 8) {
    const float* r = m + (int64_t)row * ld;
    uint32_t key[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      key[j] = c < S ? order_key(__ldg(r + c)) : 0xffffffffu;
    }
    uint32_t ans = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t trial = ans | (1u << bit);
      int c = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) c += key[j] < trial ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c < t) ans = trial;
    }
    if (lane == 0) out[row] = order_key_inv(ans);
  }
}

// A cheap UPPER BOUND of the t-th smallest value of every row (what the fused all-pairs pass needs from the sampled
// columns: any threshold that at least t elements of the row stay below is valid; a looser one only lengthens the
// candidate lists).  Every lane keeps the M smallest of its strided share of the row (three min/max per element), then
// the warp pops the t smallest of those 32 * M values: the result is the t-th smallest of a SUBSET of the row, hence
// >= the row's t-th smallest, and equal to it unless some lane held more than M of the row's t smallest values.
template <int M>
__global__ void __launch_bounds__(256)
k_row_kth_bound(const float* __restrict__ m, int64_t ld, int R, int S, int t, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < R; row += gridDim.x * 8) {
    const float* r = m + (int64_t)row * ld;
    float a[M];
#pragma unroll
    for (int i = 0; i < M; ++i) a[i] = INFINITY;
    for (int c0 = 0; c0 < S; c0 += 128) {           // four independent loads per lane and trip
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int c = c0 + u * 32 + lane; v[u] = c < S ? __ldg(r + c) : INFINITY; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float x = v[u] == v[u] ? v[u] : INFINITY;   // NaN sorts last
#pragma unroll
        for (int i = 0; i < M; ++i) { const float hi = fmaxf(a[i], x); a[i] = fminf(a[i], x); x = hi; }   // insertion into the sorted M
      }
    }
    float kth = INFINITY;
    for (int round = 0; round < t; ++round) {       // pop the minimum of the 32 lane heads t times
      float mn = a[0];
      for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      kth = mn;
      const unsigned who = __ballot_sync(0xffffffffu, a[0] == mn);
      if (who == 0u) break;                         // (only NaN / empty left)
      if (lane == __ffs(who) - 1) {
#pragma unroll
        for (int i = 0; i + 1 < M; ++i) a[i] = a[i + 1];
        a[M - 1] = INFINITY;
      }
    }
    if (lane == 0) out[row] = kth;
  }
}

__global__ void __launch_bounds__(256)
k_row_max(const float* __restrict__ dist, int64_t ld, int Q, int G, float* __restrict__ row_max) {
  __shared__ float sh[8];
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    const float* row = dist + (int64_t)q * ld;
    float m = -INFINITY;
    int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
    head = min(head, G);
    if ((int)threadIdx.x < head) m = fmaxf(m, row[threadIdx.x]);
    const int nvec = (G - head) >> 2;
    const float4* rv = reinterpret_cast<const float4*>(row + head);
    for (int v = threadIdx.x; v < nvec; v += 256) {
      const float4 x = ldg_stream4_topk(rv + v);
      m = fmaxf(fmaxf(m, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
    }
    for (int j = head + 4 * nvec + threadIdx.x; j < G; j += 256) m = fmaxf(m, row[j]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = sh[0];
      for (int w = 1; w < 8; ++w) t = fmaxf(t, sh[w]);
      row_max[q] = t;
    }
    __syncthreads();
  }
}

}  // namespace mpreid

using namespace mpreid;

extern "C" int mpreid_row_topk(const float* dist, int64_t ld_dist, int64_t Q, int64_t G, int k,
                               const float* row_scale, int32_t* idx, float* val, void* stream) {
  MPREID_REQUIRE(dist && idx && Q > 0 && G > 0 && ld_dist >= G && Q < INT32_MAX && G < INT32_MAX, "row_topk: bad arguments");
  MPREID_REQUIRE(k >= 1 && k <= kTopkMaxK, "row_topk: k must be in [1, %d], got %d", kTopkMaxK, k);
  // per-warp candidate buffer: max(2k, 256) entries between cuts plus one full chunk
  int cap = (2 * k > 256 ? 2 * k : 256) + kTopkChunk;
  cap = (cap + 31) / 32 * 32;
  const int per_warp = cap * 8 + 256 * 4;
  int warps = (220 * 1024) / per_warp;            // as many warps per CTA as shared memory allows, at most 8
  warps = warps > kTopkWarps ? kTopkWarps : warps;
  MPREID_REQUIRE(warps >= 1, "row_topk: k=%d needs more shared memory than one SM has", k);
  const int smem = warps * per_warp;
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_row_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int sms = sm_count_of_current_device();
  int ctas_per_sm = (220 * 1024) / (smem + 1024);
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t want = ceil_div(Q, warps);
  const int64_t grid = want < (int64_t)sms * ctas_per_sm ? want : (int64_t)sms * ctas_per_sm;
  k_row_topk<<<(unsigned)grid, warps * 32, smem, (cudaStream_t)stream>>>(dist, ld_dist, (int)Q, (int)G, k, row_scale, idx, val, cap);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_row_max(const float* dist, int64_t ld_dist, int64_t Q, int64_t G, float* row_max, void* stream) {
  MPREID_REQUIRE(dist && row_max && Q > 0 && G > 0 && ld_dist >= G && Q < INT32_MAX && G < INT32_MAX, "row_max: bad arguments");
  const int sms = sm_count_of_current_device();
  const int64_t grid = Q < (int64_t)sms * 8 ? Q : (int64_t)sms * 8;
  k_row_max<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(dist, ld_dist, (int)Q, (int)G, row_max);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_cand_topk(const uint64_t* cand, const int32_t* cand_cnt, int64_t cand_cap, int64_t N, int k,
                                const float* row_scale, const float* thr, int32_t* idx, float* val, uint64_t* keys, int32_t* status, void* stream) {
  MPREID_REQUIRE(cand && cand_cnt && thr && (idx || keys) && status && N > 0 && N < INT32_MAX && cand_cap >= 1, "cand_topk: bad arguments");
  MPREID_REQUIRE(k >= 1 && k <= kTopkMaxK, "cand_topk: k must be in [1, %d], got %d", kTopkMaxK, k);
  int cap = (2 * k > 256 ? 2 * k : 256) + kTopkChunk;
  cap = (cap + 31) / 32 * 32;
  const int per_warp = cap * 8 + 256 * 4;
  int warps = (220 * 1024) / per_warp;
  warps = warps > kTopkWarps ? kTopkWarps : warps;
  MPREID_REQUIRE(warps >= 1, "cand_topk: k=%d needs more shared memory than one SM has", k);
  const int smem = warps * per_warp;
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_cand_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  MPREID_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), (cudaStream_t)stream));
  const int sms = sm_count_of_current_device();
  int ctas_per_sm = (220 * 1024) / (smem + 1024);
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t want = ceil_div(N, warps);
  const int64_t grid = want < (int64_t)sms * ctas_per_sm ? want : (int64_t)sms * ctas_per_sm;
  k_cand_topk<<<(unsigned)grid, warps * 32, smem, (cudaStream_t)stream>>>((const unsigned long long*)cand, cand_cnt, cand_cap, (int)N, k, row_scale, thr,
                                                                          idx, val, (unsigned long long*)keys, status, cap);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_merge_topk(const uint64_t* keys_all, int P, int64_t N, int k, const float* row_scale, const float* thr,
                                 int32_t* idx, float* val, int32_t* status, void* stream) {
  MPREID_REQUIRE(keys_all && thr && idx && val && status && P >= 1 && N > 0 && N < INT32_MAX && k >= 1, "merge_topk: bad arguments");
  MPREID_REQUIRE((int64_t)P * k <= kMergeMax, "merge_topk: ranks * k = %lld exceeds %d", (long long)P * k, kMergeMax);
  MPREID_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), (cudaStream_t)stream));
  const int sms = sm_count_of_current_device();
  const int64_t want = ceil_div(N, 4);
  const int64_t grid = want < (int64_t)sms * 6 ? want : (int64_t)sms * 6;
  k_merge_topk<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>((const unsigned long long*)keys_all, P, (int)N, k, row_scale, thr, idx, val, status);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_row_kth(const float* dist, int64_t ld_dist, int64_t R, int64_t S, int t, float* out, void* stream) {
  MPREID_REQUIRE(dist && out && R > 0 && S > 0 && ld_dist >= S && R < INT32_MAX, "row_kth: bad arguments");
  MPREID_REQUIRE(S <= 4096 && t >= 1 && t <= S, "row_kth: needs S <= 4096 and 1 <= t <= S (got S=%lld t=%d)", (long long)S, t);
  const int sms = sm_count_of_current_device();
  const int64_t want = ceil_div(R, 8);
  const int64_t grid = want < (int64_t)sms * 8 ? want : (int64_t)sms * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 512) k_row_kth<16><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  else if (S <= 2048) k_row_kth<64><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  else k_row_kth<128><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

extern "C" int mpreid_row_kth_bound(const float* dist, int64_t ld_dist, int64_t R, int64_t S, int t, float* out, void* stream) {
  MPREID_REQUIRE(dist && out && R > 0 && S > 0 && ld_dist >= S && R < INT32_MAX, "row_kth_bound: bad arguments");
  MPREID_REQUIRE(t >= 1 && t <= 128 && t <= S, "row_kth_bound: needs 1 <= t <= min(128, S) (got S=%lld t=%d)", (long long)S, t);
  const int sms = sm_count_of_current_device();
  const int64_t want = ceil_div(R, 8);
  const int64_t grid = want < (int64_t)sms * 8 ? want : (int64_t)sms * 8;
  cudaStream_t st = (cudaStream_t)stream;
  // M per lane: the row's t smallest land ~t/32 per lane (Poisson); M is chosen so that a lane holds more than M of them
  // in well under 1 % of the rows (then the bound is exact)
  if (t <= 32) k_row_kth_bound<4><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  else if (t <= 64) k_row_kth_bound<8><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  else k_row_kth_bound<16><<<(unsigned)grid, 256, 0, st>>>(dist, ld_dist, (int)R, (int)S, t, out);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
