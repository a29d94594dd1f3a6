// Ranking + CMC / AP without an argsort  (replaces eval_func, utils/metrics.py:28-88).
//
// For a query q only the gallery entries carrying q's pid matter: they are the correct matches and,
// with the junk rule on, the removed entries (same pid AND same camera).  Their 1-based rank under
// np.argsort(kind='stable') is  1 + #{g : (d[q,g], g) < (d[q,p], p)}  lexicographically, which one
// streaming pass over the distance row delivers for all of them at once: sort the few (d, index)
// keys, binary-search every row element among them, histogram.  HBM traffic = the row, once.
//
// Kernels (all launched on the caller's stream):
//   label_*        group gallery indices by pid with an open-addressing hash table (arbitrary int64)
//   rank_count     per query: keys -> sort -> stream row -> ranks of the same-pid entries, junk fix-up
//   ap_finalize    per query: float64 AP in numpy's pairwise summation order (common.cuh)
#include <cstddef>
#include "common.cuh"

namespace mpreid {

static constexpr int64_t kEmptyKey = INT64_MIN;
static constexpr int kRankThreads = 256;
static constexpr int kRankCap = 1024;  // same-pid keys sorted per pass in shared memory

struct RankWs {
  int64_t* keys;    // [T]
  int32_t* cnt;     // [T]
  int32_t* start;   // [T]
  int32_t* fill;    // [T]
  int32_t* g_slot;  // [G]
  int32_t* list;    // [G]   gallery indices grouped by pid
  int32_t* q_start; // [Q]
  int32_t* q_cnt;   // [Q]
  int32_t* q_off;   // [Q+1] exclusive scan of q_cnt
  int32_t* row_len; // [Q]   gallery entries kept after junk removal
  int32_t* cursor;  // [1]
  int32_t* pos_tmp; // [cap]
  int32_t* pos_rank;// [cap]
  int64_t T;
};

static int64_t table_size(int64_t G) {
  int64_t t = 64;
  while (t < 2 * G) t <<= 1;
  return t;
}

static size_t carve(RankWs* w, char* base, int64_t Q, int64_t G, int64_t cap) {
  size_t off = 0;
  const int64_t T = table_size(G);
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return base ? base + o : nullptr; };
  char* p;
  p = take(T * 8); if (w) w->keys = (int64_t*)p;
  p = take(T * 4); if (w) w->cnt = (int32_t*)p;
  p = take(T * 4); if (w) w->start = (int32_t*)p;
  p = take(T * 4); if (w) w->fill = (int32_t*)p;
  p = take(G * 4); if (w) w->g_slot = (int32_t*)p;
  p = take(G * 4); if (w) w->list = (int32_t*)p;
  p = take(Q * 4); if (w) w->q_start = (int32_t*)p;
  p = take(Q * 4); if (w) w->q_cnt = (int32_t*)p;
  p = take((Q + 1) * 4); if (w) w->q_off = (int32_t*)p;
  p = take(Q * 4); if (w) w->row_len = (int32_t*)p;
  p = take(256); if (w) w->cursor = (int32_t*)p;
  p = take(cap * 4); if (w) w->pos_tmp = (int32_t*)p;
  p = take(cap * 4); if (w) w->pos_rank = (int32_t*)p;
  if (w) w->T = T;
  return off;
}

// ------------------------------------------------------------------------------- label index
__global__ void k_table_init(int64_t* keys, int32_t* cnt, int64_t T, int32_t* cursor) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < T) { keys[i] = kEmptyKey; cnt[i] = 0; }
  if (i == 0) *cursor = 0;
}

__device__ __forceinline__ int32_t table_insert(int64_t* keys, int64_t T, int64_t pid) {
  uint64_t s = mix64((uint64_t)pid) & (uint64_t)(T - 1);
  while (true) {
    unsigned long long prev = atomicCAS((unsigned long long*)&keys[s], (unsigned long long)kEmptyKey, (unsigned long long)pid);
    if ((int64_t)prev == kEmptyKey || (int64_t)prev == pid) return (int32_t)s;
    s = (s + 1) & (uint64_t)(T - 1);
  }
}

__device__ __forceinline__ int32_t table_find(const int64_t* keys, int64_t T, int64_t pid) {
  uint64_t s = mix64((uint64_t)pid) & (uint64_t)(T - 1);
  while (true) {
    int64_t k = keys[s];
    if (k == pid) return (int32_t)s;
    if (k == kEmptyKey) return -1;
    s = (s + 1) & (uint64_t)(T - 1);
  }
}

__global__ void k_gallery_insert(const int64_t* __restrict__ g_pid, int64_t G, int64_t* keys, int32_t* cnt,
                                 int32_t* g_slot, int64_t T) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= G) return;
  int32_t s = table_insert(keys, T, g_pid[j]);
  g_slot[j] = s;
  atomicAdd(&cnt[s], 1);
}

__global__ void k_slot_alloc(const int32_t* cnt, int32_t* start, int32_t* fill, int32_t* cursor, int64_t T) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= T) return;
  int32_t c = cnt[i];
  fill[i] = 0;
  start[i] = c > 0 ? atomicAdd(cursor, c) : 0;
}

__global__ void k_gallery_fill(const int32_t* __restrict__ g_slot, const int32_t* __restrict__ start, int32_t* fill,
                               int32_t* list, int64_t G) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= G) return;
  int32_t s = g_slot[j];
  int32_t p = atomicAdd(&fill[s], 1);
  list[start[s] + p] = (int32_t)j;
}

__global__ void k_query_lookup(const int64_t* __restrict__ q_pid, int64_t Q, const int64_t* keys, const int32_t* cnt,
                               const int32_t* start, int32_t* q_start, int32_t* q_cnt, int64_t T) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int32_t s = table_find(keys, T, q_pid[q]);
  q_start[q] = s < 0 ? 0 : start[s];
  q_cnt[q] = s < 0 ? 0 : cnt[s];
}

// single-CTA exclusive scan of q_cnt -> q_off[0..Q]; also reports total and max to status
__global__ void k_scan_counts(const int32_t* __restrict__ in, int32_t* out, int64_t n, int64_t capacity, int32_t* status) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  __shared__ int32_t max_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry_s = 0; max_s = 0; }
  __syncthreads();
  int32_t local_max = 0;
  for (int64_t base = 0; base < n; base += blockDim.x) {
    int64_t i = base + tid;
    int32_t v = i < n ? in[i] : 0;
    local_max = max(local_max, v);
    int64_t x = v;
    for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    int64_t prefix = carry_s + (wid > 0 ? warp_sums[wid - 1] : 0) + x - v;
    if (i < n) out[i] = (int32_t)min(prefix, (int64_t)INT32_MAX);
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s += warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  atomicMax(&max_s, local_max);
  __syncthreads();
  if (tid == 0) {
    out[n] = (int32_t)min(carry_s, (int64_t)INT32_MAX);
    status[0] = carry_s > capacity ? 1 : 0;
    status[1] = (int32_t)min(carry_s, (int64_t)INT32_MAX);
    status[2] = max_s;
    status[3] = 0;
  }
}

// ------------------------------------------------------------------------------- block helpers
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_u64(uint64_t* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += THREADS) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t x = a[i], y = a[ixj];
          bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// in-place inclusive scan of a[0..n) in shared memory (n <= a few thousand)
template <int THREADS>
__device__ __forceinline__ void block_inclusive_scan(uint32_t* a, int n, uint32_t* scratch /*[THREADS]*/) {
  const int per = (n + THREADS - 1) / THREADS;
  const int lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += a[i];
  scratch[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < THREADS; o <<= 1) {
    uint32_t v = threadIdx.x >= o ? scratch[threadIdx.x - o] : 0;
    __syncthreads();
    scratch[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = scratch[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) { run += a[i]; a[i] = run; }
  __syncthreads();
}

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------- rank_count
// Row streaming is split in two so that the expensive part never runs divergently:
//   classify  every lane tests its 16 elements against the farthest same-pid distance with ONE float
//             compare each (branch free) and appends the few survivors to its warp's queue;
//   drain     when a warp's queue holds >= 128 entries the warp processes them densely, one entry per
//             lane: 256-bin lookup table -> short binary search among the sorted keys -> histogram
//             bump, aggregated across the lanes that hit the same bucket (match.any) so the popular
//             last bucket costs one shared-memory atomic per warp, not 32 serialised ones.
// (With ~7 % survivors a per-element branch would be taken by ~90 % of the warps.)
static constexpr int kWarps = kRankThreads / 32;
static constexpr int kQueueDrain = 128;                 // drain threshold
static constexpr int kQueueCap = kQueueDrain + 32 * 16; // + one full warp chunk

struct RankSmem {
  uint64_t keys[kRankCap];
  uint32_t hist[kRankCap + 1];
  uint32_t scratch[kRankThreads];
  uint16_t lut[257];
  float qval[kWarps][kQueueCap];      // per-warp candidate queue: row values ...
  uint32_t qidx[kWarps][kQueueCap];   // ... and their gallery indices
  int32_t misc[4];
};

struct RowCtx {
  const uint64_t* keys; const uint16_t* lut; uint32_t* hist;
  int m; uint32_t omin, tmax; int shift; float tmax_f;
};

// One warp's candidate queue, structure-of-arrays so that an insertion is two plain 32-bit stores
struct WarpQueue { float* val; uint32_t* idx; };

__device__ __forceinline__ void drain_queue(const RowCtx& c, const WarpQueue& q, int count, int lane) {
  for (int e0 = 0; e0 < count; e0 += 32) {   // warp-uniform trip count: every lane reaches match.any
    const int e = e0 + lane;
    int b = -1;
    if (e < count) {
      const uint32_t o = order_key(q.val[e]);
      if (o <= c.tmax) {   // re-check in key space (the float pre-test lets -0.0 / NaN through)
        const uint64_t key = ((uint64_t)o << 32) | q.idx[e];
        int lo = 0, hi = 0;
        if (o >= c.omin) { const uint32_t bin = (o - c.omin) >> c.shift; lo = c.lut[bin]; hi = c.lut[bin + 1]; }
        // keys before lo sit in earlier bins (< key), keys from hi on in later bins (> key): #keys < key is in [lo, hi]
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (c.keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        b = lo;
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (b >= 0 && (int)(__ffs(peers) - 1) == lane) atomicAdd(&c.hist[b], (uint32_t)__popc(peers));
  }
}

// queue append of one row value, hand-scheduled: 5 instructions per element (setp, index add, two stores,
// cursor add), stores and cursor predicated on "value not above tmax" (unordered passes: NaN is sorted out by the drain)
static constexpr int kQueueIdxOffset = kWarps * kQueueCap * 4;   // byte distance qval[w][e] -> qidx[w][e]
static_assert(offsetof(RankSmem, qidx) - offsetof(RankSmem, qval) == kQueueIdxOffset, "queue planes must be adjacent");
__device__ __forceinline__ void queue_push(uint32_t& cursor, float v, float tmax_f, uint32_t idx) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.leu.f32 p, %1, %2;\n\t"
      "@p st.shared.f32 [%0], %1;\n\t"
      "@p st.shared.u32 [%0+%4], %3;\n\t"
      "@p add.u32 %0, %0, 4;\n\t}"
      : "+r"(cursor) : "f"(v), "f"(tmax_f), "r"(idx), "n"(kQueueIdxOffset) : "memory");
}
// -1 if v is a candidate (not above tmax), else 0
__device__ __forceinline__ int cand_mask(float v, float tmax_f) {
  int r;
  asm("set.leu.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(v), "f"(tmax_f));
  return r;
}

// One pass of kRankThreads * U float4 of the row (already in registers): count the candidates (row value
// not above the largest same-pid value), one warp scan for the queue positions, then predicated (value,
// index) stores.  FULL = every thread has all U vectors in range (no bounds tests in the steady state).
template <bool FULL, int U>
__device__ __forceinline__ void stream_block(const float4 (&x)[U], int v0, int nvec, int head, float tmax_f,
                                             const RowCtx& c, const WarpQueue& q, int& qcount) {
  const int tid = threadIdx.x, lane = tid & 31;
  int n = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (FULL || v0 + u * kRankThreads + tid < nvec)
      n -= cand_mask(x[u].x, tmax_f) + cand_mask(x[u].y, tmax_f) + cand_mask(x[u].z, tmax_f) + cand_mask(x[u].w, tmax_f);
  }
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total) {
    uint32_t cursor = (uint32_t)__cvta_generic_to_shared(q.val + qcount + incl - n);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (FULL || v0 + u * kRankThreads + tid < nvec) {
        const uint32_t j = (uint32_t)(head + 4 * (v0 + u * kRankThreads + tid));
        queue_push(cursor, x[u].x, tmax_f, j);
        queue_push(cursor, x[u].y, tmax_f, j + 1);
        queue_push(cursor, x[u].z, tmax_f, j + 2);
        queue_push(cursor, x[u].w, tmax_f, j + 3);
      }
    }
    qcount += total;
    __syncwarp();
    if (qcount >= kQueueDrain) { drain_queue(c, q, qcount, lane); qcount = 0; __syncwarp(); }
  }
}

__device__ __forceinline__ void stream_row(const float* __restrict__ row, int G, const RowCtx& c, const WarpQueue& q) {
  const int tid = threadIdx.x, lane = tid & 31;
  int qcount = 0;  // warp-uniform
  int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
  head = min(head, G);
  const int nvec = (G - head) >> 2;
  const float4* rv = reinterpret_cast<const float4*>(row + head);
  constexpr int U = 4;
  const float tmax_f = c.tmax_f;
  // ragged ends (< 4 elements each) go through the same queue, handled by warp 0
  if (tid < 32) {
    uint32_t mask = 0; float v[2] = {0.f, 0.f}; uint32_t jj[2] = {0u, 0u};
    if (lane < head) { v[0] = row[lane]; jj[0] = lane; mask |= !(v[0] > tmax_f) ? 1u : 0u; }
    const int t = head + 4 * nvec + lane;
    if (t < G) { v[1] = row[t]; jj[1] = t; mask |= !(v[1] > tmax_f) ? 2u : 0u; }
    const int n = __popc(mask);
    int incl = n;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    int pos = qcount + incl - n;
    if (mask & 1u) { q.val[pos] = v[0]; q.idx[pos] = jj[0]; ++pos; }
    if (mask & 2u) { q.val[pos] = v[1]; q.idx[pos] = jj[1]; ++pos; }
    qcount += __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
  }
  constexpr int kStep = kRankThreads * U;
  // software pipeline: the next block's 64 bytes per thread are in flight while this block is classified
  float4 cur[U], nxt[U];
  int v0 = 0;
  if (kStep <= nvec) {
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = ldg_stream4(rv + u * kRankThreads + tid);
  }
  for (; v0 + kStep <= nvec; v0 += kStep) {
    const int v1 = v0 + kStep;
    if (v1 + kStep <= nvec) {
#pragma unroll
      for (int u = 0; u < U; ++u) nxt[u] = ldg_stream4(rv + v1 + u * kRankThreads + tid);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v1 + u * kRankThreads + tid;
        nxt[u] = v < nvec ? ldg_stream4(rv + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    stream_block<true, U>(cur, v0, nvec, head, tmax_f, c, q, qcount);
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
  if (v0 < nvec) {
    if (v0 == 0) {   // row shorter than one full block: nothing was prefetched
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = u * kRankThreads + tid;
        cur[u] = v < nvec ? ldg_stream4(rv + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    stream_block<false, U>(cur, v0, nvec, head, tmax_f, c, q, qcount);
  }
  if (qcount) drain_queue(c, q, qcount, lane);
}

// ---- rows with at most 32 same-pid entries (the common case): warp 0 sorts the keys and finishes the row
//      with shuffles / ballots, so the CTA pays 4 barriers per row instead of ~65
__device__ __forceinline__ uint64_t warp_sort_u64(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      const bool take_min = lower == up;
      key = (take_min == (other < key)) ? other : key;
    }
  }
  return key;
}

__global__ void __launch_bounds__(kRankThreads)
k_rank_count(const float* __restrict__ dist, int64_t ld, int Q, int G,
             const int64_t* __restrict__ q_cam, const int64_t* __restrict__ g_cam, int junk_mode,
             const int32_t* __restrict__ list, const int32_t* __restrict__ q_start, const int32_t* __restrict__ q_cnt,
             const int32_t* __restrict__ q_off, int32_t* status,
             int32_t* pos_tmp, int32_t* pos_rank, int32_t* first_hit, int32_t* num_rel, int32_t* row_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RankSmem& s = *reinterpret_cast<RankSmem*>(smem_raw);
  if (status[0] != 0) return;  // workspace overflow: the host re-runs with a larger capacity
  const int tid = threadIdx.x;
  // rows cost anything between "no candidates" and "every element is one": CTAs take rows from a ticket
  // counter (status[3], zeroed by k_scan_counts) instead of a fixed stride
  for (;;) {
    if (tid == 0) s.misc[1] = atomicAdd(&status[3], 1);
    __syncthreads();
    const int q = s.misc[1];
    __syncthreads();
    if (q >= Q) break;
    const int cnt = q_cnt[q];
    if (cnt == 0) {
      if (tid == 0) { first_hit[q] = 0; num_rel[q] = 0; row_len[q] = G; }
      continue;
    }
    const int start = q_start[q], off = q_off[q];
    const float* row = dist + (int64_t)q * ld;
    const int64_t qc = junk_mode ? q_cam[q] : 0;

    const bool small = cnt <= 32;   // block-uniform
    for (int c0 = 0; c0 < cnt; c0 += kRankCap) {
      const int m = min(kRankCap, cnt - c0);
      const int P = (int)next_pow2_u32((uint32_t)m);
      if (small) {
        if (tid < 32) {
          uint64_t key = ~0ull;
          if (tid < m) { const int32_t j = list[start + tid]; key = make_key(row[j], (uint32_t)j); }
          s.keys[tid] = warp_sort_u64(key, tid);
        }
        for (int i = tid; i <= m; i += kRankThreads) s.hist[i] = 0;
        __syncthreads();
      } else {
        for (int i = tid; i < P; i += kRankThreads) {
          uint64_t key = ~0ull;
          if (i < m) { const int32_t j = list[start + c0 + i]; key = make_key(row[j], (uint32_t)j); }
          s.keys[i] = key;
        }
        for (int i = tid; i <= m; i += kRankThreads) s.hist[i] = 0;
        __syncthreads();
        bitonic_sort_u64<kRankThreads>(s.keys, P);
      }
      RowCtx c;
      c.keys = s.keys; c.lut = s.lut; c.hist = s.hist; c.m = m;
      c.tmax = (uint32_t)(s.keys[m - 1] >> 32);
      c.omin = (uint32_t)(s.keys[0] >> 32);
      const uint32_t range = c.tmax - c.omin;
      c.shift = range < 256u ? 0 : (32 - __clz(range)) - 8;
      // float twin of tmax for the branch-free pre-test; NaN keys (0xffffffff) mean "everything may precede"
      c.tmax_f = c.tmax == 0xffffffffu ? INFINITY : order_key_inv(c.tmax);
      {
        // lut[b] = #keys whose 32-bit order key is below the first value of bin b; lut[256] = m
        for (int b = tid; b < 256; b += kRankThreads) {
          const uint64_t lower = (uint64_t)c.omin + ((uint64_t)b << c.shift);
          int lo = 0, hi = m;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((uint64_t)(s.keys[mid] >> 32) < lower) lo = mid + 1; else hi = mid;
          }
          s.lut[b] = (uint16_t)lo;
        }
        if (tid == 0) s.lut[256] = (uint16_t)m;
      }
      __syncthreads();
      stream_row(row, G, c, WarpQueue{s.qval[tid >> 5], s.qidx[tid >> 5]});
      __syncthreads();

      if (small) {
        // warp 0: ranks = inclusive scan of the interval counts; junk entries drop out and shift the kept ranks
        if (tid < 32) {
          uint32_t r = tid < m ? s.hist[tid] : 0u;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, r, o); if (tid >= o) r += y; }
          bool junk = false;
          if (tid < m && junk_mode) junk = g_cam[(int32_t)(s.keys[tid] & 0xffffffffu)] == qc;
          const unsigned jm = __ballot_sync(0xffffffffu, junk);
          const int jb = __popc(jm & (0xffffffffu >> (31 - tid)));   // junk among sorted positions <= tid
          if (tid < m && !junk) pos_rank[off + tid - jb] = (int32_t)r - jb;
          const unsigned kept = __ballot_sync(0xffffffffu, tid < m && !junk);
          const int first = kept ? __ffs(kept) - 1 : 0;
          const int32_t r_first = (int32_t)__shfl_sync(0xffffffffu, r, first) - __popc(jm & (0xffffffffu >> (31 - first)));
          if (tid == 0) {
            const int nj = __popc(jm);
            num_rel[q] = cnt - nj;
            row_len[q] = G - nj;
            first_hit[q] = kept ? r_first : 0;
          }
        }
        __syncthreads();
        continue;
      }
      // ---- rank of sorted key i = #row entries <= key i
      block_inclusive_scan<kRankThreads>(s.hist, m, s.scratch);
      for (int i = tid; i < m; i += kRankThreads) {
        const int32_t j = (int32_t)(s.keys[i] & 0xffffffffu);
        const bool junk = junk_mode && (g_cam[j] == qc);
        const int32_t r = (int32_t)s.hist[i];
        pos_tmp[off + c0 + i] = junk ? -r : r;
      }
      __syncthreads();
    }
    if (small) continue;

    // ---- junk fix-up: kept rank = rank - #junk ranked before; compact the kept ranks (ascending)
    if (cnt <= kRankCap) {
      for (int i = tid; i < cnt; i += kRankThreads) s.hist[i] = pos_tmp[off + i] < 0 ? 1u : 0u;
      __syncthreads();
      block_inclusive_scan<kRankThreads>(s.hist, cnt, s.scratch);
      for (int i = tid; i < cnt; i += kRankThreads) {
        const int32_t r = pos_tmp[off + i];
        if (r > 0) {
          const int32_t jb = (int32_t)s.hist[i];  // junk among sorted positions <= i (i itself is kept)
          pos_rank[off + (i - jb)] = r - jb;
        }
      }
      __syncthreads();
      if (tid == 0) {
        const int32_t nj = (int32_t)s.hist[cnt - 1];
        num_rel[q] = cnt - nj;
        row_len[q] = G - nj;
      }
    } else {
      // rare: more same-pid entries than one shared-memory pass holds -> O(cnt^2) fix-up in global
      if (tid == 0) s.misc[0] = 0;
      __syncthreads();
      int32_t my_junk = 0;
      for (int i = tid; i < cnt; i += kRankThreads) {
        const int32_t r = pos_tmp[off + i];
        if (r < 0) { ++my_junk; continue; }
        int32_t jb = 0, kb = 0;
        for (int f = 0; f < cnt; ++f) {
          const int32_t rf = pos_tmp[off + f];
          if (rf < 0) jb += (-rf < r); else kb += (rf < r);
        }
        pos_rank[off + kb] = r - jb;
      }
      atomicAdd(&s.misc[0], my_junk);
      __syncthreads();
      if (tid == 0) { num_rel[q] = cnt - s.misc[0]; row_len[q] = G - s.misc[0]; }
    }
    __syncthreads();
    if (tid == 0) first_hit[q] = num_rel[q] > 0 ? pos_rank[off] : 0;
    __syncthreads();
  }
}

__global__ void k_ap_finalize(const int32_t* __restrict__ pos_rank, const int32_t* __restrict__ q_off,
                              const int32_t* __restrict__ num_rel, const int32_t* __restrict__ row_len,
                              const int32_t* __restrict__ status, double* ap, int Q) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q || status[0] != 0) return;
  const int m = num_rel[q];
  if (m == 0) { ap[q] = 0.0; return; }
  ap[q] = pairwise_sparse_sum(pos_rank + q_off[q], m, (int64_t)row_len[q]) / (double)m;
}

}  // namespace mpreid

using namespace mpreid;

extern "C" size_t mpreid_rank_eval_workspace_bytes(int64_t Q, int64_t G, int64_t pos_capacity) {
  if (Q < 0 || G < 0 || pos_capacity < 0) return 0;
  return carve(nullptr, nullptr, Q, G, pos_capacity);
}

extern "C" int mpreid_rank_eval(const float* dist, int64_t ld_dist, int64_t Q, int64_t G,
                                const int64_t* q_pid, const int64_t* g_pid,
                                const int64_t* q_cam, const int64_t* g_cam, int junk_mode,
                                int32_t* first_hit, double* ap, int32_t* num_rel,
                                void* workspace, size_t workspace_bytes, int64_t pos_capacity,
                                int32_t* status, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MPREID_REQUIRE(dist && q_pid && g_pid && first_hit && ap && num_rel && workspace && status, "rank_eval: null pointer");
  MPREID_REQUIRE(Q > 0 && G > 0 && Q < INT32_MAX && G < INT32_MAX && ld_dist >= G, "rank_eval: bad shape Q=%lld G=%lld ld=%lld",
                 (long long)Q, (long long)G, (long long)ld_dist);
  MPREID_REQUIRE(junk_mode == MPREID_JUNK_NONE || (q_cam && g_cam), "rank_eval: junk mode needs camids");
  MPREID_REQUIRE(((uintptr_t)workspace & 255) == 0, "rank_eval: workspace must be 256-byte aligned");
  if (workspace_bytes < carve(nullptr, nullptr, Q, G, pos_capacity)) {
    set_error("rank_eval: workspace too small (%zu bytes)", workspace_bytes);
    return MPREID_ERR_WORKSPACE;
  }
  RankWs w;
  carve(&w, (char*)workspace, Q, G, pos_capacity);
  const int B = 256;
  k_table_init<<<(unsigned)ceil_div(w.T, B), B, 0, st>>>(w.keys, w.cnt, w.T, w.cursor);
  k_gallery_insert<<<(unsigned)ceil_div(G, B), B, 0, st>>>(g_pid, G, w.keys, w.cnt, w.g_slot, w.T);
  k_slot_alloc<<<(unsigned)ceil_div(w.T, B), B, 0, st>>>(w.cnt, w.start, w.fill, w.cursor, w.T);
  k_gallery_fill<<<(unsigned)ceil_div(G, B), B, 0, st>>>(w.g_slot, w.start, w.fill, w.list, G);
  k_query_lookup<<<(unsigned)ceil_div(Q, B), B, 0, st>>>(q_pid, Q, w.keys, w.cnt, w.start, w.q_start, w.q_cnt, w.T);
  k_scan_counts<<<1, 1024, 0, st>>>(w.q_cnt, w.q_off, Q, pos_capacity, status);
  int sms = sm_count_of_current_device();
  int ctas_per_sm = 4;
  int grid = (int)((Q < (int64_t)sms * ctas_per_sm) ? Q : (int64_t)sms * ctas_per_sm);
  const int rank_smem = (int)sizeof(RankSmem);
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_rank_count, cudaFuncAttributeMaxDynamicSharedMemorySize, rank_smem));  // per device, cheap
  k_rank_count<<<grid, kRankThreads, rank_smem, st>>>(dist, ld_dist, (int)Q, (int)G, q_cam, g_cam, junk_mode != MPREID_JUNK_NONE,
                                              w.list, w.q_start, w.q_cnt, w.q_off, status, w.pos_tmp, w.pos_rank,
                                              first_hit, num_rel, w.row_len);
  k_ap_finalize<<<(unsigned)ceil_div(Q, 128), 128, 0, st>>>(w.pos_rank, w.q_off, num_rel, w.row_len, status, ap, (int)Q);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
