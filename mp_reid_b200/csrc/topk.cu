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
static constexpr int kTopkCap = 6144;       // candidate buffer entries (48 KB)
static constexpr int kTopkChunkVec = 4;     // float4 loads per thread per chunk
static constexpr int kTopkMaxK = 2048;

struct TopkSmem {
  uint64_t buf[kTopkCap];
  uint64_t keep[kTopkMaxK];   // survivors of a cut, staged before they move to the buffer front
  uint32_t hist[256];
  int cnt, keep_cnt, sel_rank, sel_bin;
  float bound;       // raw-domain rejection bound
  uint64_t thr;      // current k-th best key (divided domain); ~0 = none yet
  uint64_t sel_prefix;
};

__device__ __forceinline__ float4 ldg_stream4_topk(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += kTopkThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
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

__device__ __forceinline__ void offer(TopkSmem& s, float bound, uint64_t thr, float v, uint32_t j, float r, bool has_scale) {
  if (v > bound) return;                         // NaN falls through and sorts last
  const float dv = has_scale ? v / r : v;
  const uint64_t key = make_key(dv, j);
  if (key < thr) {
    const int slot = atomicAdd(&s.cnt, 1);
    s.buf[slot] = key;                           // capacity is guaranteed by the chunk protocol
  }
}

// Cut the candidate buffer back to its k smallest keys WITHOUT sorting it: MSB-first radix select
// (8 bits per pass, shared-memory histogram) finds the k-th smallest 64-bit key, then one sweep
// keeps the keys <= it (keys are unique, so exactly k survive).  All threads call; cnt >= k >= 1.
__device__ void cut_to_k(TopkSmem& s, int k, float r, bool has_scale) {
  const int tid = threadIdx.x;
  __syncthreads();
  const int cnt = s.cnt;
  if (tid == 0) { s.sel_prefix = 0; s.sel_rank = k; s.keep_cnt = 0; }
  uint64_t mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    s.hist[tid] = 0;
    __syncthreads();
    const uint64_t prefix = s.sel_prefix;
    for (int i = tid; i < cnt; i += kTopkThreads) {
      const uint64_t key = s.buf[i];
      if ((key & mask) == prefix) atomicAdd(&s.hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins 8l..8l+7; find the bin holding the sel_rank-th smallest
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) { c[b] = s.hist[tid * 8 + b]; sum += c[b]; }
      uint32_t incl = sum;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
      const uint32_t excl = incl - sum;
      const uint32_t want = (uint32_t)s.sel_rank;
      __syncwarp();
      if (want > excl && want <= incl) {
        uint32_t run = excl;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (want > run && want <= run + c[b]) {
            s.sel_bin = (c[b] == 1u) ? 1 : 0;          // 1 = the chosen bin holds a single key: it IS the k-th
            s.sel_rank = (int)(want - run);
            s.sel_prefix = prefix | ((uint64_t)(tid * 8 + b) << shift);
          }
          run += c[b];
        }
      }
    }
    mask |= (uint64_t)255u << shift;
    __syncthreads();
    if (s.sel_bin && shift > 0) {
      // early exit: fetch the unique key that carries the selected prefix
      const uint64_t pfx = s.sel_prefix;
      __syncthreads();
      for (int i = tid; i < cnt; i += kTopkThreads) {
        const uint64_t key = s.buf[i];
        if ((key & mask) == pfx) s.sel_prefix = key;
      }
      __syncthreads();
      break;
    }
  }
  const uint64_t kth = s.sel_prefix;
  for (int i = tid; i < cnt; i += kTopkThreads) {
    const uint64_t key = s.buf[i];
    if (key <= kth) s.keep[atomicAdd(&s.keep_cnt, 1)] = key;
  }
  __syncthreads();
  for (int i = tid; i < k; i += kTopkThreads) s.buf[i] = s.keep[i];
  if (tid == 0) {
    s.cnt = k;
    s.thr = kth;
    const uint32_t hi = (uint32_t)(kth >> 32);
    s.bound = hi == 0xffffffffu ? INFINITY : raw_bound(order_key_inv(hi), r, has_scale);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kTopkThreads)
k_row_topk(const float* __restrict__ dist, int64_t ld, int Q, int G, int k, const float* __restrict__ row_scale,
           int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TopkSmem& s = *reinterpret_cast<TopkSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int keff = min(k, G);
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    const float* row = dist + (int64_t)q * ld;
    const bool has_scale = row_scale != nullptr;
    const float r = has_scale ? row_scale[q] : 1.0f;
    if (tid == 0) { s.cnt = 0; s.thr = ~0ull; s.bound = INFINITY; }
    __syncthreads();

    int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
    head = min(head, G);
    if (tid < head) offer(s, INFINITY, ~0ull, row[tid], tid, r, has_scale);
    const int nvec = (G - head) >> 2;
    const float4* rv = reinterpret_cast<const float4*>(row + head);
    const int tail0 = head + 4 * nvec;
    if (tail0 + tid < G) offer(s, INFINITY, ~0ull, row[tail0 + tid], tail0 + tid, r, has_scale);  // < 4 tail elements
    __syncthreads();

    // chunk 0 is one float4 per thread so that a threshold exists early; later chunks are 4 float4 per
    // thread, and the loads of chunk c+1 are issued before chunk c is processed (they fly across the
    // barriers of the cut protocol)
    float4 cur[kTopkChunkVec], nxt[kTopkChunkVec];
    int v0 = 0, U = 1;
#pragma unroll
    for (int u = 0; u < kTopkChunkVec; ++u) {
      const int v = v0 + u * kTopkThreads + tid;
      if (u < U && v < nvec) cur[u] = ldg_stream4_topk(rv + v);
    }
    while (v0 < nvec) {
      const int v1 = v0 + kTopkThreads * U;
      const int Un = kTopkChunkVec;
#pragma unroll
      for (int u = 0; u < kTopkChunkVec; ++u) {
        const int v = v1 + u * kTopkThreads + tid;
        if (v < nvec) nxt[u] = ldg_stream4_topk(rv + v);
      }
      // everyone reads the decision inputs, then a barrier, so that no thread appends before all have read
      const int chunk_elems = kTopkThreads * U * 4;
      const bool need = s.cnt >= keff && (s.cnt > kTopkCap - chunk_elems || (s.thr == ~0ull && s.cnt >= max(2 * keff, 1024)));
      __syncthreads();
      if (need) cut_to_k(s, keff, r, has_scale);
      const float bound = s.bound;
      const uint64_t thr = s.thr;
#pragma unroll
      for (int u = 0; u < kTopkChunkVec; ++u) {
        const int v = v0 + u * kTopkThreads + tid;
        if (u < U && v < nvec) {
          const uint32_t j = head + 4 * v;
          offer(s, bound, thr, cur[u].x, j, r, has_scale);
          offer(s, bound, thr, cur[u].y, j + 1, r, has_scale);
          offer(s, bound, thr, cur[u].z, j + 2, r, has_scale);
          offer(s, bound, thr, cur[u].w, j + 3, r, has_scale);
        }
      }
#pragma unroll
      for (int u = 0; u < kTopkChunkVec; ++u) cur[u] = nxt[u];
      v0 = v1;
      U = Un;
      __syncthreads();
    }
    // final: cut to keff, sort those few, emit
    {
      __syncthreads();
      if (s.cnt > keff) cut_to_k(s, keff, r, has_scale);
      const int cnt = s.cnt;
      const int P = (int)next_pow2_u32((uint32_t)max(cnt, 1));
      for (int i = cnt + tid; i < P; i += kTopkThreads) s.buf[i] = ~0ull;
      __syncthreads();
      bitonic_sort_smem(s.buf, P);
      for (int i = tid; i < k; i += kTopkThreads) {
        int32_t id = -1;
        float val = INFINITY;
        if (i < keff) {
          const uint64_t key = s.buf[i];
          id = (int32_t)(key & 0xffffffffu);
          val = has_scale ? row[id] / r : row[id];
        }
        idx_out[(int64_t)q * k + i] = id;
        if (val_out) val_out[(int64_t)q * k + i] = val;
      }
      __syncthreads();
    }
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
  static bool attr_set = false;
  const int smem = (int)sizeof(TopkSmem);
  if (!attr_set) {
    MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_row_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int sms = sm_count_of_current_device();
  const int64_t grid = Q < (int64_t)sms * 3 ? Q : (int64_t)sms * 3;
  k_row_topk<<<(unsigned)grid, kTopkThreads, smem, (cudaStream_t)stream>>>(dist, ld_dist, (int)Q, (int)G, k, row_scale, idx, val);
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
