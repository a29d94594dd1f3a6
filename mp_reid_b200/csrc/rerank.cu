// k-reciprocal re-ranking (utils/reranking.py:29-100) as sparse gather kernels.
//
// The reference materialises dense N x N fp16 matrices V / V_qe and walks them with three Python
// loops.  V has at most (k1+1)(round(k1/2)+2) non-zeros per row, so everything here is sparse rows
// in ELL storage (sorted by column), an inverted index (CSC) over the gallery rows, and one dense
// pass only where the output itself is dense (the Q x G blend).
//
// Orientation: dist[i][j] == reference distmat[j][i]; the reference normalises by the column max and
// transposes (:46), so row i of `dist` divided by its own max is row i of the reference's
// original_dist.  Rounding points are the reference's: V and V_qe are fp16, kernel weights are fp32
// exp / fp32 numpy-pairwise sum, the Jaccard accumulator is fp16, (1-lambda) multiplies in fp16.
//
//   k_build_v0      :51-71  k-reciprocal set, 2/3 expansion rule, np.unique, Gaussian kernel row
//   k_query_expand  :73-78  V_qe[i] = fp16(mean of the V rows of the first k2 neighbours)
//   k_csc_*         :80-82  inverted index over gallery rows
//   k_jaccard       :84-99  fp16 min-accumulation in ascending column order, Jaccard, lambda blend
#include <cstdlib>
#include "common.cuh"

namespace mpreid {

static constexpr int kMaxK1 = 100;

HD int v0_capacity(int k1, int64_t N) {
  const int K1 = k1 + 1, half = round_half_even_div2(k1) + 1;
  int64_t c = (int64_t)K1 * (1 + half);
  return (int)(c < N ? c : N);
}
HD int64_t v_capacity(int k1, int k2, int64_t N) {
  if (k2 == 1) return v0_capacity(k1, N);
  int64_t c = (int64_t)k2 * v0_capacity(k1, N);
  return c < N ? c : N;
}

static constexpr int kQeSmemEntries = 4096;

// numpy float32 pairwise sum (np.sum(weight), utils/reranking.py:71), dense version
__device__ float pairwise_sum_f32(const float* a, int n) {
  struct Frame { int lo, n, stage; float left; };
  Frame st[24];
  int sp = 1;
  st[0].lo = 0; st[0].n = n; st[0].stage = 0; st[0].left = 0.f;
  float ret = 0.f;
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.stage == 0) {
      if (f.n < 8) {
        float res = 0.f;
        for (int i = 0; i < f.n; ++i) res += a[f.lo + i];
        ret = res; --sp; continue;
      }
      if (f.n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[f.lo + j];
        int i = 8;
        for (; i < f.n - (f.n % 8); i += 8)
          for (int j = 0; j < 8; ++j) r[j] += a[f.lo + i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < f.n; ++i) res += a[f.lo + i];
        ret = res; --sp; continue;
      }
      int n2 = f.n / 2; n2 -= n2 % 8;
      f.stage = 1;
      Frame& c = st[sp++];
      c.lo = f.lo; c.n = n2; c.stage = 0;
    } else if (f.stage == 1) {
      f.left = ret; f.stage = 2;
      int n2 = f.n / 2; n2 -= n2 % 8;
      Frame& c = st[sp++];
      c.lo = f.lo + n2; c.n = f.n - n2; c.stage = 0;
    } else {
      ret = f.left + ret; --sp;
    }
  }
  return ret;
}

template <typename T>
__device__ __forceinline__ void block_bitonic(T* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const T x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive block scan of one int per thread (blockDim.x <= 1024); returns the prefix, total in *total
__device__ __forceinline__ int block_exclusive_scan(int v, int* sh /*[33]*/, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = lane < (int)((blockDim.x + 31) >> 5) ? sh[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
    sh[lane] = w;
  }
  __syncthreads();
  const int prefix = (wid > 0 ? sh[wid - 1] : 0) + x - v;
  *total = sh[((blockDim.x + 31) >> 5) - 1];
  __syncthreads();
  return prefix;
}

// ----------------------------------------------------------------------------------- V0 rows
static constexpr int kV0Threads = 128;

// Shared-memory view of one V0 CTA, carved from dynamic shared memory with sizes that follow k1 (252 list
// entries for k1=20 but 5,353 for k1=100): small k1 -> a few KB per CTA -> many resident CTAs to hide the
// dependent neighbour-list loads.
struct V0Smem {
  int32_t* fwd; int32_t* recip; int32_t* cand_cnt; int32_t* cand_common;   // [K1]
  int32_t* list;   // [list_cap] (power of two >= K1 * (half + 1))
  float* w;        // [K1 * (half + 1)]
  uint8_t* flag;   // [K1 * half]
  int32_t* ctrl;   // n_recip, n_list, n_out, wsum bits
  int* sh;         // [33]
};
struct V0Sizes { int K1, half, list_cap, bytes; };
__host__ __device__ __forceinline__ V0Sizes v0_sizes(int k1) {
  V0Sizes z;
  z.K1 = k1 + 1;
  z.half = round_half_even_div2(k1) + 1;
  z.list_cap = (int)next_pow2_u32((uint32_t)(z.K1 * (z.half + 1)));
  const int words = 4 * z.K1 + z.list_cap + z.K1 * (z.half + 1) + 4 + 33;
  z.bytes = words * 4 + ((z.K1 * z.half + 15) & ~15);
  return z;
}
__device__ __forceinline__ V0Smem v0_carve(unsigned char* base, const V0Sizes& z) {
  V0Smem s;
  int32_t* p = reinterpret_cast<int32_t*>(base);
  s.fwd = p; p += z.K1; s.recip = p; p += z.K1; s.cand_cnt = p; p += z.K1; s.cand_common = p; p += z.K1;
  s.list = p; p += z.list_cap;
  s.w = reinterpret_cast<float*>(p); p += z.K1 * (z.half + 1);
  s.ctrl = p; p += 4;
  s.sh = reinterpret_cast<int*>(p); p += 33;
  s.flag = reinterpret_cast<uint8_t*>(p);
  return s;
}

// Where the distances original_dist[i, idx] (:70) come from: the rows of the all-pairs matrix (dist != nullptr), or --
// when the matrix was never written (fused all-pairs pass) -- the neighbour values of row i (the same accumulators, already
// divided by the row maximum) for idx inside the top-K list, and an fp32 dot product of the feature rows for the few
// expansion members outside it (typically < 20 per row: the 2/3 rule only admits sets that mostly overlap R(i)).
struct V0Source {
  const float* dist; int64_t ld;                    // matrix rows [R, >= N]
  const float* nbr_val;                              // [N, K] divided neighbour values
  const float* xn; int64_t ld_xn; int D;             // feature rows [N, D] fp32
  const float* sqnorm;                               // [N]
};

__global__ void __launch_bounds__(kV0Threads)
k_build_v0(const V0Source src, const int32_t* __restrict__ row_ids, int R, int k1, int K, int Keff_in,
           const int32_t* __restrict__ nbr, const float* __restrict__ rowmax,
           int32_t* __restrict__ v0_col, uint16_t* __restrict__ v0_val, int32_t* __restrict__ v0_len, int C0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const V0Smem s = v0_carve(smem_raw, v0_sizes(k1));
  int32_t& n_recip = s.ctrl[0]; int32_t& n_list = s.ctrl[1]; int32_t& n_out = s.ctrl[2];
  float& wsum_s = *reinterpret_cast<float*>(&s.ctrl[3]);
  const int tid = threadIdx.x;
  const int K1 = min(k1 + 1, Keff_in);                       // forward list length (:53)
  const int half = min(round_half_even_div2(k1) + 1, Keff_in);  // candidate list length (:60)
  // il = row of this block of the all-pairs matrix, i = the sample it belongs to (global index)
  for (int il = blockIdx.x; il < R; il += gridDim.x) {
    const int i = row_ids ? row_ids[il] : il;
    if (tid == 0) { n_recip = 0; n_list = 0; }
    for (int m = tid; m < K1; m += kV0Threads) s.fwd[m] = nbr[(int64_t)i * K + m];
    __syncthreads();
    // reciprocity: i in the first K1 neighbours of fwd[m]   (:54-56)
    for (int m = tid; m < K1; m += kV0Threads) {
      const int32_t* row = nbr + (int64_t)s.fwd[m] * K;
      bool hit = false;
      for (int t = 0; t < K1; ++t) hit |= (row[t] == i);
      if (hit) { const int p = atomicAdd(&n_recip, 1); s.recip[p] = s.fwd[m]; }
    }
    __syncthreads();
    const int nR = n_recip;
    // candidate k-reciprocal sets with the half-size lists  (:58-64)
    for (int p = tid; p < nR * half; p += kV0Threads) {
      const int j = p / half, m = p - j * half;
      const int32_t c = s.recip[j];
      const int32_t x = nbr[(int64_t)c * K + m];
      const int32_t* row = nbr + (int64_t)x * K;
      bool hit = false;
      for (int t = 0; t < half; ++t) hit |= (row[t] == c);
      s.flag[j * half + m] = hit ? 1 : 0;
    }
    for (int j = tid; j < nR; j += kV0Threads) { s.cand_cnt[j] = 0; s.cand_common[j] = 0; }
    __syncthreads();
    for (int p = tid; p < nR * half; p += kV0Threads) {
      if (!s.flag[p]) continue;
      const int j = p / half, m = p - j * half;
      const int32_t x = nbr[(int64_t)s.recip[j] * K + m];
      bool common = false;
      for (int t = 0; t < nR; ++t) common |= (s.recip[t] == x);
      atomicAdd(&s.cand_cnt[j], 1);
      if (common) atomicAdd(&s.cand_common[j], 1);
    }
    __syncthreads();
    // expansion list = R(i) plus every accepted candidate set  (:65-67), then np.unique (:69)
    for (int t = tid; t < nR; t += kV0Threads) { const int p = atomicAdd(&n_list, 1); s.list[p] = s.recip[t]; }
    for (int p = tid; p < nR * half; p += kV0Threads) {
      if (!s.flag[p]) continue;
      const int j = p / half, m = p - j * half;
      if ((double)s.cand_common[j] > (2.0 / 3.0) * (double)s.cand_cnt[j]) {
        const int q = atomicAdd(&n_list, 1);
        s.list[q] = nbr[(int64_t)s.recip[j] * K + m];
      }
    }
    __syncthreads();
    const int nL = n_list;
    const int P = (int)next_pow2_u32((uint32_t)max(nL, 1));
    for (int t = nL + tid; t < P; t += kV0Threads) s.list[t] = INT32_MAX;
    __syncthreads();
    block_bitonic(s.list, P);
    // unique -> compact in place (sorted, so compaction keeps order); n_out <= C0
    if (tid == 0) {
      int n = 0;
      for (int t = 0; t < nL; ++t)
        if (t == 0 || s.list[t] != s.list[t - 1]) s.list[n++] = s.list[t];
      n_out = n;
    }
    __syncthreads();
    const int nU = n_out;
    const float rmax = rowmax[il];
    if (src.dist) {
      const float* drow = src.dist + (int64_t)il * src.ld;
      for (int t = tid; t < nU; t += kV0Threads) s.w[t] = drow[s.list[t]] / rmax;   // original_dist[i, idx]  (:46)
    } else {
      for (int t = tid; t < nU; t += kV0Threads) {
        const int32_t x = s.list[t];
        float dn = __int_as_float(0x7fc00000);          // NaN = "not a neighbour of i": computed below
        for (int m = 0; m < K1; ++m) if (s.fwd[m] == x) { dn = src.nbr_val[(int64_t)i * K + m]; break; }
        s.w[t] = dn;
      }
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      const float* xi = src.xn + (int64_t)i * src.ld_xn;
      const bool vec = (src.D & 3) == 0 && (src.ld_xn & 3) == 0 && ((uintptr_t)src.xn & 15) == 0;
      for (int t = warp; t < nU; t += kV0Threads / 32) {
        if (s.w[t] == s.w[t]) continue;                 // warp-uniform
        const int32_t x = s.list[t];
        const float* xx = src.xn + (int64_t)x * src.ld_xn;
        float acc = 0.f;
        if (vec) {
          for (int c = lane * 4; c < src.D; c += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xi + c)), b = __ldg(reinterpret_cast<const float4*>(xx + c));
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
          }
        } else {
          for (int c = lane; c < src.D; c += 32) acc = fmaf(__ldg(xi + c), __ldg(xx + c), acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        __syncwarp();
        if (lane == 0) s.w[t] = fmaf(-2.0f, acc, src.sqnorm[i] + src.sqnorm[x]) / rmax;   // utils/reranking.py:38-40,46
      }
    }
    __syncthreads();
    for (int t = tid; t < nU; t += kV0Threads) s.w[t] = (float)exp((double)(-s.w[t]));   // np.exp on float32 (:70)
    __syncthreads();
    if (tid == 0) wsum_s = pairwise_sum_f32(s.w, nU);   // np.sum(weight)         (:71)
    __syncthreads();
    const float wsum = wsum_s;
    // V[i, idx] = fp16(weight / sum); entries that underflow to 0 are not stored (V != 0 tests, :82,88)
    int keep = 0;
    uint16_t hv = 0;
    // nU can exceed blockDim: process in rounds, keeping column order
    int written = 0;
    for (int t0 = 0; t0 < nU; t0 += kV0Threads) {
      const int t = t0 + tid;
      keep = 0;
      if (t < nU) {
        const __half h = __float2half_rn(s.w[t] / wsum);
        hv = __half_as_ushort(h);
        keep = (hv & 0x7fff) != 0;
      }
      int total;
      const int pos = block_exclusive_scan(keep, s.sh, &total);
      if (keep) {
        v0_col[(int64_t)il * C0 + written + pos] = s.list[t];
        v0_val[(int64_t)il * C0 + written + pos] = hv;
      }
      written += total;
    }
    if (tid == 0) v0_len[il] = written;
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------- query expansion
static constexpr int kQeThreads = 256;

__global__ void __launch_bounds__(kQeThreads)
k_query_expand(int N, int K, int k2, const int32_t* __restrict__ nbr,
               const int32_t* __restrict__ v0_col, const uint16_t* __restrict__ v0_val, const int32_t* __restrict__ v0_len, int C0,
               int32_t* __restrict__ v_col, uint16_t* __restrict__ v_val, int32_t* __restrict__ v_len, int64_t C1,
               uint64_t* __restrict__ scratch, int64_t scratch_P, int skip_upto, int row_lo) {
  __shared__ uint64_t sbuf[kQeSmemEntries];
  __shared__ int sh[33];
  __shared__ int s_total;
  const int tid = threadIdx.x;
  const float inv_cnt = (float)k2;
  for (int i = row_lo + blockIdx.x; i < N; i += gridDim.x) {   // rows [row_lo, N)
    // gather (col, m, val) of the k2 neighbour rows; key = col<<32 | m<<16 | fp16 bits
    if (tid == 0) {
      int t = 0;
      for (int m = 0; m < k2; ++m) {
        const int32_t r = nbr[(int64_t)i * K + m];
        t += r >= 0 ? v0_len[r] : 0;
      }
      s_total = t;
    }
    __syncthreads();
    const int T = s_total;
    if (T <= skip_upto) { __syncthreads(); continue; }   // block-uniform: this row belongs to the warp-per-row kernel
    const int P = (int)next_pow2_u32((uint32_t)max(T, 1));
    uint64_t* buf = (P <= kQeSmemEntries) ? sbuf : scratch + (int64_t)blockIdx.x * scratch_P;
    {
      int base = 0;
      for (int m = 0; m < k2; ++m) {
        const int32_t r = nbr[(int64_t)i * K + m];
        const int len = r >= 0 ? v0_len[r] : 0;
        for (int e = tid; e < len; e += kQeThreads) {
          const uint64_t col = (uint32_t)v0_col[(int64_t)r * C0 + e];
          buf[base + e] = (col << 32) | ((uint64_t)m << 16) | v0_val[(int64_t)r * C0 + e];
        }
        base += len;
      }
      for (int t = T + tid; t < P; t += kQeThreads) buf[t] = ~0ull;
    }
    __syncthreads();
    block_bitonic(buf, P);
    // segment heads sum their (<= k2) members in m order, fp32, then / k2 -> fp16   (:76)
    int written = 0;
    for (int t0 = 0; t0 < T; t0 += kQeThreads) {
      const int t = t0 + tid;
      int keep = 0;
      uint16_t hv = 0;
      int32_t col = 0;
      if (t < T) {
        const uint64_t key = buf[t];
        col = (int32_t)(key >> 32);
        const bool head = (t == 0) || ((int32_t)(buf[t - 1] >> 32) != col);
        if (head) {
          float sum = __half2float(__ushort_as_half((uint16_t)(key & 0xffff)));
          for (int u = t + 1; u < T && (int32_t)(buf[u] >> 32) == col; ++u)
            sum += __half2float(__ushort_as_half((uint16_t)(buf[u] & 0xffff)));
          hv = __half_as_ushort(__float2half_rn(sum / inv_cnt));
          keep = (hv & 0x7fff) != 0;
        }
      }
      int total;
      const int pos = block_exclusive_scan(keep, sh, &total);
      if (keep) {
        v_col[(int64_t)i * C1 + written + pos] = col;
        v_val[(int64_t)i * C1 + written + pos] = hv;
      }
      written += total;
    }
    if (tid == 0) v_len[i] = written;
    __syncthreads();
  }
}

// Warp-per-row variant for the common case (the k2 gathered V0 rows hold at most kQeWarpEntries entries together, e.g.
// 6 x ~25 for k1 = 20, k2 = 6): gather, warp-level bitonic sort by (column, m), segment sums.  No CTA barrier; eight
// rows in flight per CTA.  Rows with more entries are left to k_query_expand (skip_upto).
static constexpr int kQeWarpEntries = 512;
static constexpr int kQeWarps = 8;

__global__ void __launch_bounds__(kQeWarps * 32)
k_query_expand_warp(int N, int K, int k2, const int32_t* __restrict__ nbr,
                    const int32_t* __restrict__ v0_col, const uint16_t* __restrict__ v0_val, const int32_t* __restrict__ v0_len, int C0,
                    int32_t* __restrict__ v_col, uint16_t* __restrict__ v_val, int32_t* __restrict__ v_len, int64_t C1, int row_lo) {
  __shared__ uint64_t sbuf[kQeWarps][kQeWarpEntries];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* buf = sbuf[warp];
  const float cnt = (float)k2;
  for (int i = row_lo + blockIdx.x * kQeWarps + warp; i < N; i += gridDim.x * kQeWarps) {   // rows [row_lo, N)
    // lane m (< k2 <= 64: two rows per lane) holds neighbour m and the length of its V0 row
    int32_t r[2]; int len[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      r[h] = m < k2 ? nbr[(int64_t)i * K + m] : -1;
      len[h] = r[h] >= 0 ? v0_len[r[h]] : 0;
    }
    int incl0 = len[0];
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl0, o); if (lane >= o) incl0 += y; }
    const int tot0 = __shfl_sync(0xffffffffu, incl0, 31);
    int incl1 = len[1];
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl1, o); if (lane >= o) incl1 += y; }
    const int T = tot0 + __shfl_sync(0xffffffffu, incl1, 31);
    if (T > kQeWarpEntries) continue;                    // warp-uniform: the CTA kernel takes this row
    // gather (col, m, val): key = col << 32 | m << 16 | fp16 bits
    for (int m = 0; m < k2; ++m) {
      const int h = m >> 5, src = m & 31;
      const int32_t rm = __shfl_sync(0xffffffffu, r[h], src);
      const int lm = __shfl_sync(0xffffffffu, len[h], src);
      const int base = (h ? tot0 : 0) + __shfl_sync(0xffffffffu, (h ? incl1 : incl0) - len[h], src);
      for (int e = lane; e < lm; e += 32) {
        const uint64_t col = (uint32_t)v0_col[(int64_t)rm * C0 + e];
        buf[base + e] = (col << 32) | ((uint64_t)m << 16) | v0_val[(int64_t)rm * C0 + e];
      }
    }
    const int P = (int)next_pow2_u32((uint32_t)max(T, 1));
    for (int t = T + lane; t < P; t += 32) buf[t] = ~0ull;
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (P >> 1); t += 32) {      // compare-exchange pair t of this stage: every lane works
          const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
          const uint64_t x = buf[lo], y = buf[hi];
          const bool up = (lo & kk) == 0;
          if ((x > y) == up) { buf[lo] = y; buf[hi] = x; }
        }
        __syncwarp();
      }
    }
    // segment heads sum their (<= k2) members in m order, fp32, then / k2 -> fp16   (:76)
    int written = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      uint16_t hv = 0;
      int32_t col = 0;
      if (t < T) {
        const uint64_t key = buf[t];
        col = (int32_t)(key >> 32);
        const bool head = (t == 0) || ((int32_t)(buf[t - 1] >> 32) != col);
        if (head) {
          float sum = __half2float(__ushort_as_half((uint16_t)(key & 0xffff)));
          for (int u = t + 1; u < T && (int32_t)(buf[u] >> 32) == col; ++u)
            sum += __half2float(__ushort_as_half((uint16_t)(buf[u] & 0xffff)));
          hv = __half_as_ushort(__float2half_rn(sum / cnt));
          keep = (hv & 0x7fff) != 0;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = written + __popc(bal & ((1u << lane) - 1u));
        v_col[(int64_t)i * C1 + pos] = col;
        v_val[(int64_t)i * C1 + pos] = hv;
      }
      written += __popc(bal);
    }
    if (lane == 0) v_len[i] = written;
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------- inverted index
__global__ void k_zero_i32(int32_t* a, int32_t* b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { a[i] = 0; b[i] = 0; }
}

__global__ void k_csc_count(int N, int Q, const int32_t* __restrict__ v_col, const int32_t* __restrict__ v_len, int64_t C1,
                            int32_t* col_cnt) {
  // one warp per gallery row
  const int g = Q + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (g >= N) return;
  const int len = v_len[g];
  for (int e = threadIdx.x & 31; e < len; e += 32) atomicAdd(&col_cnt[v_col[(int64_t)g * C1 + e]], 1);
}

__global__ void k_scan_i64(const int32_t* __restrict__ in, int64_t* out, int64_t n) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    const int64_t v = i < n ? in[i] : 0;
    int64_t x = v;
    for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    if (i < n) out[i] = carry_s + (wid > 0 ? warp_sums[wid - 1] : 0) + x - v;
    __syncthreads();
    if (tid == 0) carry_s += warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

__global__ void k_csc_fill(int N, int Q, const int32_t* __restrict__ v_col, const uint16_t* __restrict__ v_val,
                           const int32_t* __restrict__ v_len, int64_t C1, const int64_t* __restrict__ col_off,
                           int32_t* col_fill, int32_t* __restrict__ csc_row, uint16_t* __restrict__ csc_val) {
  const int g = Q + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (g >= N) return;
  const int len = v_len[g];
  for (int e = threadIdx.x & 31; e < len; e += 32) {
    const int32_t c = v_col[(int64_t)g * C1 + e];
    const int64_t p = col_off[c] + atomicAdd(&col_fill[c], 1);
    csc_row[p] = g;
    csc_val[p] = v_val[(int64_t)g * C1 + e];
  }
}

// ----------------------------------------------------------------------------------- Jaccard + blend
// Two kernels.  (1) k_blend_default: the dense part of :93-95 for gallery entries the query shares no V column with
// (temp_min = 0 -> jaccard = 1 -> fp16(1 * fp16(1 - lambda))): a pure stream, 8 bytes per (query, gallery) pair, run
// at full occupancy.  (2) k_jaccard_sparse: the fp16 min-accumulation (:86-92) over the inverted lists and the
// overwrite of the few entries it touches.  Probe (oracle, k1=20, k2=6): 570 .. 3,700 touched gallery entries and
// 1,700 .. 7,800 accumulator updates per query, i.e. ~1 % of the dense work.
static constexpr int kBlendThreads = 256;
static constexpr int kBlendPer = 8;        // elements per thread per trip
static constexpr int kJacSteps = 64;       // V entries of the query row staged per round
static constexpr int kJacMaxTile = 41600;  // fp16 accumulator entries per CTA (81 KB): two CTAs per SM, the MSMT17 gallery is two tiles

__device__ __forceinline__ float jaccard_blend(float a /*temp_min != 0*/, float dn, float lambda_value, __half one_minus_lambda) {
  const __half den = __float2half_rn(2.0f - a);                                   // 2 - temp_min              (:93)
  const __half quo = __float2half_rn(a / __half2float(den));                      // temp_min / (2 - temp_min)
  const __half jac = __float2half_rn(1.0f - __half2float(quo));                   // 1 - ...
  const __half jl = __float2half_rn(__half2float(jac) * __half2float(one_minus_lambda));   // fp16 * fp16(1 - lambda)  (:95)
  return __fadd_rn(__half2float(jl), __fmul_rn(dn, lambda_value));                // two roundings, no FMA contraction
}

// final[il, c] = float(fp16(1 - lambda)) + (dist[il, col0 + c] / rowmax[il]) * lambda      (:46,72,95 with temp_min = 0)
template <bool VEC>
__global__ void __launch_bounds__(kBlendThreads)
k_blend_default(const float* __restrict__ dist, int64_t ld, int64_t col0, int Qs, int G, float lambda_value,
                const float* __restrict__ rowmax, float* __restrict__ final_dist, int64_t ld_final, const int32_t* __restrict__ src_rows) {
  const float base = __half2float(__float2half_rn((float)(1.0 - (double)lambda_value)));
  constexpr int kChunk = kBlendThreads * kBlendPer;
  const int chunks = (G + kChunk - 1) / kChunk;
  const int64_t total = (int64_t)Qs * chunks;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int il = (int)(w / chunks);
    const int c0 = (int)(w - (int64_t)il * chunks) * kChunk;
    const int sr = src_rows ? src_rows[il] : il;   // row of `dist` / `rowmax` that holds query il (its global index in the sharded form)
    const float rmax = __ldg(rowmax + sr);
    const float* drow = dist + (int64_t)sr * ld + col0;
    float* orow = final_dist + (int64_t)il * ld_final;
    if (VEC) {
      // both rows 16-byte aligned: two float4 per thread, 512 B per warp instruction
#pragma unroll
      for (int h = 0; h < kBlendPer / 4; ++h) {
        const int c = c0 + (h * kBlendThreads + threadIdx.x) * 4;
        if (c + 4 <= G) {
          float4 v;
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(drow + c));
          v.x = __fadd_rn(base, __fmul_rn(v.x / rmax, lambda_value));
          v.y = __fadd_rn(base, __fmul_rn(v.y / rmax, lambda_value));
          v.z = __fadd_rn(base, __fmul_rn(v.z / rmax, lambda_value));
          v.w = __fadd_rn(base, __fmul_rn(v.w / rmax, lambda_value));
          __stcs(reinterpret_cast<float4*>(orow + c), v);
        } else {
          for (int j = c; j < G; ++j) orow[j] = __fadd_rn(base, __fmul_rn(drow[j] / rmax, lambda_value));
        }
      }
    } else {
      float v[kBlendPer];
#pragma unroll
      for (int j = 0; j < kBlendPer; ++j) {
        const int c = c0 + j * kBlendThreads + threadIdx.x;
        v[j] = c < G ? __ldg(drow + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kBlendPer; ++j) {
        const int c = c0 + j * kBlendThreads + threadIdx.x;
        if (c < G) orow[c] = __fadd_rn(base, __fmul_rn(v[j] / rmax, lambda_value));
      }
    }
  }
}

// ---- k_jaccard_bucket: the production kernel.  Work item = (query row, gallery tile) as in k_jaccard_sparse below,
// but without a barrier per step (measured at the MSMT17 shape: ~150 steps and ~26,000 accumulator updates per query;
// with one or two resident CTAs per SM the per-step barriers and a binary search per entry cost 4.2 ms).
// The flat sequence of list entries (step-major == the reference's accumulation order) is cut into rounds of
// <= kJacRound entries.  Per round:
//   A. every warp takes a contiguous 1/16 of the round (16 entries per lane, kept in registers), forms
//      min(V[i,k], V[g,k]) and counts its entries per OWNER warp (the tile is split into 16 contiguous ranges);
//   B. exact offsets: bucket of owner o = the entries of producer 0, then producer 1, ...: global step order is kept;
//   C. the producers scatter from registers into the buckets;
//   D. every owner warp applies its bucket 32 entries at a time; entries of the same gallery sample inside one group
//      (different steps) go in lane order, everything else in parallel.
// Four barriers per round instead of one per step; every entry is read from global memory once per tile.
static constexpr int kJacRound = 8192;
static constexpr int kJacBWarps = 16;
static constexpr int kJacPerLane = kJacRound / (kJacBWarps * 32);   // 16

struct JacBStage {
  int64_t b[kJacSteps];
  int32_t pre[kJacSteps + 1];
  uint16_t v[kJacSteps];
  int32_t cnt[kJacBWarps][kJacBWarps];    // [producer][owner]
  int32_t off[kJacBWarps][kJacBWarps];    // [producer][owner] start inside the entry buffer
  int32_t obase[kJacBWarps + 1];          // bucket boundaries
};

__global__ void __launch_bounds__(kJacBWarps * 32)
k_jaccard_bucket(const float* __restrict__ dist, int64_t ld, int64_t col0, const int32_t* __restrict__ q_ids, int Qs, int N, int Q,
                 float lambda_value, const float* __restrict__ rowmax,
                 const int32_t* __restrict__ v_col, const uint16_t* __restrict__ v_val, const int32_t* __restrict__ v_len, int64_t C1,
                 const int64_t* __restrict__ col_off, const int32_t* __restrict__ csc_row, const uint16_t* __restrict__ csc_val,
                 float* __restrict__ final_dist, int64_t ld_final, int tile_cols, int n_tiles, int rows_global) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  JacBStage& st = *reinterpret_cast<JacBStage*>(smem_raw);
  int32_t* ent_c = reinterpret_cast<int32_t*>(smem_raw + ((sizeof(JacBStage) + 15) & ~size_t(15)));   // [kJacRound] column inside the tile
  uint16_t* ent_m = reinterpret_cast<uint16_t*>(ent_c + kJacRound);                                     // [kJacRound] min(V[i,k], V[g,k])
  __half* acc = reinterpret_cast<__half*>(ent_m + kJacRound);                                           // [tile_cols] temp_min (:87)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = N - Q;
  const __half one_minus_lambda = __float2half_rn((float)(1.0 - (double)lambda_value));
  const int64_t items = (int64_t)Qs * n_tiles;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int il = (int)(item / n_tiles);
    const int t0 = (int)(item - (int64_t)il * n_tiles) * tile_cols;
    const int i = q_ids ? q_ids[il] : il;
    const int len = v_len[i];
    const int tn = min(tile_cols, G - t0);
    const int n8 = (tn + 7) >> 3;
    const float own_scale = (float)kJacBWarps / (float)tn;
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    for (int c = tid; c < n8; c += kJacBWarps * 32) a4[c] = make_uint4(0u, 0u, 0u, 0u);
    const int gbase = Q + t0;
    for (int e0 = 0; e0 < len; e0 += kJacSteps) {
      const int nb = min(kJacSteps, len - e0);
      __syncthreads();   // previous stage fully consumed (and the zero fill visible)
      int n_l = 0;
      if (tid < kJacSteps) {
        if (tid < nb) {
          const int32_t k = v_col[(int64_t)i * C1 + e0 + tid];
          const int64_t b = col_off[k];
          n_l = (int)(col_off[k + 1] - b);
          st.b[tid] = b;
          st.v[tid] = v_val[(int64_t)i * C1 + e0 + tid];
        }
        // inclusive prefix over the 64 steps (two warps)
        int incl = n_l;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        st.pre[tid + 1] = incl;                                  // second warp: corrected below by the total of the first
        if (tid == 0) st.pre[0] = 0;
      }
      __syncthreads();
      const int first32 = st.pre[32];
      if (tid >= 32 && tid < kJacSteps) st.pre[tid + 1] += first32;
      __syncthreads();
      const int total = st.pre[nb];
      for (int r0 = 0; r0 < total; r0 += kJacRound) {
        const int rn = min(kJacRound, total - r0);
        // ---- A: this warp's contiguous share of the round, in registers
        const int per_warp = (rn + kJacBWarps - 1) / kJacBWarps;
        const int w0 = warp * per_warp, w1 = min(rn, w0 + per_warp);
        int32_t ec[kJacPerLane]; uint16_t em[kJacPerLane]; int8_t eo[kJacPerLane];
        for (int o = lane; o < kJacBWarps; o += 32) st.cnt[warp][o] = 0;
        __syncwarp();
        int step = 0;
#pragma unroll
        for (int h = 0; h < kJacPerLane / 8; ++h) {
          // addresses first, then all sixteen loads of the half in flight, then the arithmetic: a load whose value is
          // consumed right away would cost one full memory latency per entry
          int64_t src[8]; uint16_t vk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int x = w0 + (h * 8 + u) * 32 + lane;
            src[u] = -1; vk[u] = 0;
            if (x < w1) {
              const int xf = r0 + x;
              while (st.pre[step + 1] <= xf) ++step;             // monotone in x: amortised O(1)
              src[u] = st.b[step] + (xf - st.pre[step]);
              vk[u] = st.v[step];
            }
          }
          int32_t gg[8]; uint16_t ww[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            gg[u] = 0; ww[u] = 0;
            if (src[u] >= 0) { gg[u] = __ldg(csc_row + src[u]); ww[u] = __ldg(csc_val + src[u]); }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int it = h * 8 + u;
            ec[it] = -1; em[it] = 0; eo[it] = -1;
            const unsigned c = (unsigned)(gg[u] - gbase);
            if (src[u] >= 0 && c < (unsigned)tn) {
              const __half vg = __ushort_as_half(ww[u]), vik = __ushort_as_half(vk[u]);
              em[it] = __half_as_ushort(__hlt(vg, vik) ? vg : vik);                        // np.minimum on fp16  (:90-91)
              ec[it] = (int32_t)c;
              eo[it] = (int8_t)min(kJacBWarps - 1, (int)((float)c * own_scale));
            }
          }
        }
#pragma unroll
        for (int it = 0; it < kJacPerLane; ++it) {
          const int o = eo[it];
          const unsigned peers = __match_any_sync(0xffffffffu, o);
          if (o >= 0 && (peers & ((1u << lane) - 1u)) == 0) st.cnt[warp][o] += __popc(peers);   // group leader; one leader per owner
          __syncwarp();
        }
        __syncthreads();
        // ---- B: offsets  off[p][o] = sum_{o' < o} total[o'] + sum_{p' < p} cnt[p'][o]
        if (tid < kJacBWarps) {
          int tot = 0;
          for (int p = 0; p < kJacBWarps; ++p) { st.off[p][tid] = tot; tot += st.cnt[p][tid]; }
          int incl = tot;
          for (int o = 1; o < kJacBWarps; o <<= 1) { const int y = __shfl_up_sync(0x0000ffffu, incl, o); if (lane >= o) incl += y; }
          st.obase[tid + 1] = incl;
          if (tid == 0) st.obase[0] = 0;
          const int before = incl - tot;
          for (int p = 0; p < kJacBWarps; ++p) st.off[p][tid] += before;
        }
        __syncthreads();
        // ---- C: scatter (same grouping as in A, so ranks inside a group are the lane order)
#pragma unroll
        for (int it = 0; it < kJacPerLane; ++it) {
          const int o = eo[it];
          const unsigned peers = __match_any_sync(0xffffffffu, o);
          if (o >= 0) {
            const int pos = st.off[warp][o] + __popc(peers & ((1u << lane) - 1u));
            ent_c[pos] = ec[it]; ent_m[pos] = em[it];
          }
          __syncwarp();
          if (o >= 0 && (peers & ((1u << lane) - 1u)) == 0) st.off[warp][o] += __popc(peers);
          __syncwarp();
        }
        __syncthreads();
        // ---- D: owner warps apply their buckets in order
        {
          const int x1 = st.obase[warp + 1];
          for (int x0 = st.obase[warp]; x0 < x1; x0 += 32) {
            const int x = x0 + lane;
            const bool act = x < x1;
            const int32_t c = act ? ent_c[x] : -1 - lane;       // inactive lanes: distinct dummies
            const uint16_t m = act ? ent_m[x] : 0;
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            const int my = __popc(peers & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(0xffffffffu, act ? __popc(peers) : 0);
            for (int r = 0; r < rounds; ++r) {
              if (act && my == r)
                acc[c] = __float2half_rn(__half2float(acc[c]) + __half2float(__ushort_as_half(m)));   // fp16 accumulator (:87-91)
              __syncwarp();
            }
          }
        }
        __syncthreads();   // buckets consumed before the next round rewrites them
      }
    }
    __syncthreads();       // (also covers len == 0: the zero fill is complete)
    // touched entries only: Jaccard + blend, overwriting the default.  The tile is swept in slices; the touched entries
    // of a slice are first collected in the entry buffers, then blended with four independent distance loads per thread
    // in flight (the reads are scattered: one dependent load per entry would serialise the memory latency)
    const int sr = rows_global ? i : il;   // sharded form: distance rows and maxima are addressed by the global query index
    const float rmax = rowmax[sr];
    const float* drow = dist + (int64_t)sr * ld + col0;
    float* orow = final_dist + (int64_t)il * ld_final;
    int* fix_n = &st.cnt[0][0];
    constexpr int kSlice8 = kJacRound / 8;                       // uint4 chunks per slice: at most kJacRound entries
    for (int s8 = 0; s8 < n8; s8 += kSlice8) {
      if (tid == 0) *fix_n = 0;
      __syncthreads();
      const int e8 = min(n8, s8 + kSlice8);
      for (int c8 = s8 + tid; c8 < e8; c8 += kJacBWarps * 32) {
        const uint4 w = a4[c8];
        if ((w.x | w.y | w.z | w.w) == 0u) continue;
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint16_t hb = (uint16_t)(ww[j >> 1] >> ((j & 1) * 16));
          const int c = c8 * 8 + j;
          if ((hb & 0x7fffu) != 0 && c < tn) {
            const int pos = atomicAdd(fix_n, 1);
            ent_c[pos] = c; ent_m[pos] = hb;
          }
        }
      }
      __syncthreads();
      const int nf = *fix_n;
      for (int x0 = tid; x0 < nf; x0 += 4 * kJacBWarps * 32) {
        float dv[4]; int cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int x = x0 + u * kJacBWarps * 32;
          cc[u] = x < nf ? ent_c[x] : -1;
          dv[u] = cc[u] >= 0 ? __ldg(drow + t0 + cc[u]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (cc[u] >= 0) {
            const float a = __half2float(__ushort_as_half(ent_m[x0 + u * kJacBWarps * 32]));
            orow[t0 + cc[u]] = jaccard_blend(a, dv[u] / rmax, lambda_value, one_minus_lambda);   // original_dist[i, Q+g] (:46,72)
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- k_jaccard_sparse: the straightforward variant (a barrier per step), kept as the in-tree cross-check of the bucket
// kernel (MPREID_JACCARD=tile; tests compare the two bit for bit).
struct JacStage {
  int64_t b[kJacSteps];        // start of the inverted list of column k_e in the CSC arrays
  int32_t n[kJacSteps];        // its length
  int32_t pre[kJacSteps + 1];  // prefix of the lengths inside the round
  uint16_t v[kJacSteps];       // V[i, k_e]
  int32_t fit;                 // steps of this round whose lists fit the entry buffer (0: the first list alone is too long)
  int32_t pad;
};

__device__ __forceinline__ void jac_apply(__half* acc, unsigned c, uint16_t vg_bits, __half vik) {
  const __half vg = __ushort_as_half(vg_bits);
  const __half mn = __hlt(vg, vik) ? vg : vik;                                  // np.minimum on fp16  (:90-91)
  acc[c] = __float2half_rn(__half2float(acc[c]) + __half2float(mn));           // fp16 accumulator: add in fp32, round once (:87-91)
}

// Work item = (query row, gallery tile): the CTA keeps an fp16 accumulator for the tile in shared memory.  The non-zero
// columns k of V[i] are taken in ascending order (the reference's accumulation order, :88-92): step k adds
// min(V[i,k], V[g,k]) to acc[g] for every g of the inverted list of k.  Rounds of up to 64 steps: the list entries of
// a round (<= kJacEntries) are first copied to shared memory with every load in flight at once -- walking the lists
// step by step straight from global memory costs one full memory latency per step -- and then applied step by step;
// within a step every g occurs once, so the threads share the list freely; steps are separated by a barrier.
// Afterwards only the touched entries are blended and written over the defaults of k_blend_default.
static constexpr int kJacEntries = 4096;

__global__ void __launch_bounds__(512)
k_jaccard_sparse(const float* __restrict__ dist, int64_t ld, int64_t col0, const int32_t* __restrict__ q_ids, int Qs, int N, int Q,
                 float lambda_value, const float* __restrict__ rowmax,
                 const int32_t* __restrict__ v_col, const uint16_t* __restrict__ v_val, const int32_t* __restrict__ v_len, int64_t C1,
                 const int64_t* __restrict__ col_off, const int32_t* __restrict__ csc_row, const uint16_t* __restrict__ csc_val,
                 float* __restrict__ final_dist, int64_t ld_final, int tile_cols, int n_tiles, int rows_global) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  JacStage& st = *reinterpret_cast<JacStage*>(smem_raw);
  int32_t* ent_g = reinterpret_cast<int32_t*>(smem_raw + ((sizeof(JacStage) + 15) & ~size_t(15)));   // [kJacEntries]
  uint16_t* ent_v = reinterpret_cast<uint16_t*>(ent_g + kJacEntries);                                  // [kJacEntries]
  __half* acc = reinterpret_cast<__half*>(ent_v + kJacEntries);                                         // [tile_cols] temp_min (:87)
  const int tid = threadIdx.x, T = blockDim.x;
  const int G = N - Q;
  const __half one_minus_lambda = __float2half_rn((float)(1.0 - (double)lambda_value));  // fp16(1 - lambda)  (:95)
  const int64_t items = (int64_t)Qs * n_tiles;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int il = (int)(item / n_tiles);
    const int t0 = (int)(item - (int64_t)il * n_tiles) * tile_cols;
    const int i = q_ids ? q_ids[il] : il;     // global query index: selects the V row; il selects the distance / output row
    const int len = v_len[i];
    const int tn = min(tile_cols, G - t0);
    const int n8 = (tn + 7) >> 3;
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    for (int c = tid; c < n8; c += T) a4[c] = make_uint4(0u, 0u, 0u, 0u);
    const int gbase = Q + t0;
    int e0 = 0;
    while (e0 < len) {
      const int nb = min(kJacSteps, len - e0);
      __syncthreads();   // previous round applied (and the zero fill visible) before the stage is rewritten
      if (tid < nb) {
        const int32_t k = v_col[(int64_t)i * C1 + e0 + tid];
        const int64_t b = col_off[k];
        st.b[tid] = b;
        st.n[tid] = (int32_t)(col_off[k + 1] - b);
        st.v[tid] = v_val[(int64_t)i * C1 + e0 + tid];
      }
      __syncthreads();
      if (tid < 32) {
        // inclusive prefix of the list lengths (two steps per lane), then the number of steps whose lists fit the buffer
        const int a = 2 * tid < nb ? st.n[2 * tid] : 0, b2 = 2 * tid + 1 < nb ? st.n[2 * tid + 1] : 0;
        int incl = a + b2;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
        const int before = incl - a - b2;
        if (tid == 0) st.pre[0] = 0;
        st.pre[2 * tid + 1] = before + a;
        if (2 * tid + 2 <= kJacSteps) st.pre[2 * tid + 2] = incl;
        const unsigned ok0 = __ballot_sync(0xffffffffu, 2 * tid < nb && before + a <= kJacEntries);
        const unsigned ok1 = __ballot_sync(0xffffffffu, 2 * tid + 1 < nb && incl <= kJacEntries);
        if (tid == 0) st.fit = __popc(ok0) + __popc(ok1);   // the prefix is monotone: the fitting steps form a leading run
      }
      __syncthreads();
      const int fit = st.fit;
      if (fit == 0) {
        // a single inverted list longer than the entry buffer: applied straight from global memory
        const int n = st.n[0];
        const int64_t b = st.b[0];
        const __half vik = __ushort_as_half(st.v[0]);
        for (int u = tid; u < n; u += T) {
          const unsigned c = (unsigned)(csc_row[b + u] - gbase);
          if (c < (unsigned)tn) jac_apply(acc, c, csc_val[b + u], vik);
        }
        e0 += 1;
        continue;
      }
      // copy the entries of these steps: four independent (row, value) loads per thread and trip
      const int total = st.pre[fit];
      for (int x0 = tid; x0 < total; x0 += 4 * T) {
        int32_t gg[4]; uint16_t vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = x0 + j * T;
          gg[j] = -1; vv[j] = 0;
          if (x < total) {
            int lo = 0, hi = fit;                       // step s with pre[s] <= x < pre[s+1]
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (st.pre[mid] <= x) lo = mid; else hi = mid; }
            const int64_t src = st.b[lo] + (x - st.pre[lo]);
            gg[j] = csc_row[src]; vv[j] = csc_val[src];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = x0 + j * T;
          if (x < total) { ent_g[x] = gg[j]; ent_v[x] = vv[j]; }
        }
      }
      __syncthreads();
      for (int sidx = 0; sidx < fit; ++sidx) {
        const __half vik = __ushort_as_half(st.v[sidx]);
        const int x1 = st.pre[sidx + 1];
        for (int x = st.pre[sidx] + tid; x < x1; x += T) {
          const unsigned c = (unsigned)(ent_g[x] - gbase);
          if (c < (unsigned)tn) jac_apply(acc, c, ent_v[x], vik);
        }
        __syncthreads();   // the next step may hit the same accumulator entries
      }
      e0 += fit;
    }
    __syncthreads();       // (also covers len == 0: the zero fill is complete)
    // touched entries only: Jaccard + blend, overwriting the default
    const int sr = rows_global ? i : il;
    const float rmax = rowmax[sr];
    const float* drow = dist + (int64_t)sr * ld + col0;
    float* orow = final_dist + (int64_t)il * ld_final;
    for (int c8 = tid; c8 < n8; c8 += T) {
      const uint4 w = a4[c8];
      if ((w.x | w.y | w.z | w.w) == 0u) continue;
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint16_t hb = (uint16_t)(ww[j >> 1] >> ((j & 1) * 16));
        const int c = c8 * 8 + j;
        if ((hb & 0x7fffu) != 0 && c < tn) {
          const float a = __half2float(__ushort_as_half(hb));
          const float dn = drow[t0 + c] / rmax;                                          // original_dist[i, Q+g]   (:46,72)
          orow[t0 + c] = jaccard_blend(a, dn, lambda_value, one_minus_lambda);
        }
      }
    }
    __syncthreads();       // the tile is zeroed again by the next item
  }
}

}  // namespace mpreid

using namespace mpreid;

namespace mpreid {

struct FinishWs {
  int32_t* v_col; uint16_t* v_val; int32_t* v_len;       // ELL [N, C1] (unused when k2 == 1)
  int32_t* col_cnt; int32_t* col_fill; int64_t* col_off; // [N], [N], [N+1]
  int32_t* csc_row; uint16_t* csc_val;                   // [(N-Q) * C1]
  uint64_t* qe_scratch;                                  // [qe_grid * qe_P]
  int C0; int64_t C1; int qe_grid; int64_t qe_P;
};

static size_t carve_finish(FinishWs* w, char* base, int64_t N, int64_t Q, int k1, int k2, int sms) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return base ? base + o : nullptr; };
  const int C0 = v0_capacity(k1, N);
  const int64_t C1 = v_capacity(k1, k2, N);
  const int qe_grid = sms * 6;   // 6 x 256 threads x 32 KB static smem per SM
  int64_t qe_P = 1;
  while (qe_P < (int64_t)k2 * C0) qe_P <<= 1;
  if (qe_P <= kQeSmemEntries || k2 == 1) qe_P = 0;  // fits shared memory: no global scratch
  char* p;
  if (k2 != 1) {
    p = take(N * (size_t)C1 * 4); if (w) w->v_col = (int32_t*)p;
    p = take(N * (size_t)C1 * 2); if (w) w->v_val = (uint16_t*)p;
    p = take(N * 4); if (w) w->v_len = (int32_t*)p;
  }
  p = take(N * 4); if (w) w->col_cnt = (int32_t*)p;
  p = take(N * 4); if (w) w->col_fill = (int32_t*)p;
  p = take((N + 1) * 8); if (w) w->col_off = (int64_t*)p;
  p = take((size_t)(N - Q) * C1 * 4); if (w) w->csc_row = (int32_t*)p;
  p = take((size_t)(N - Q) * C1 * 2); if (w) w->csc_val = (uint16_t*)p;
  p = take((size_t)qe_grid * qe_P * 8); if (w) w->qe_scratch = (uint64_t*)p;
  if (w) { w->C0 = C0; w->C1 = C1; w->qe_grid = qe_grid; w->qe_P = qe_P; }
  return off;
}

static int neighbor_count(int k1, int k2) { return (k1 + 1) > k2 ? (k1 + 1) : k2; }

}  // namespace mpreid

extern "C" int mpreid_rerank_neighbor_count(int k1, int k2) { return neighbor_count(k1, k2); }
extern "C" int mpreid_rerank_v0_capacity(int k1, int64_t N) { return (k1 < 1 || k1 > kMaxK1 || N < 1) ? 0 : v0_capacity(k1, N); }

namespace mpreid {
static int launch_build_v0(const V0Source& src, const int32_t* row_ids, int64_t R, int64_t N, int k1, const int32_t* nbr_all, int K,
                           const float* row_max_rows, int32_t* v0_col, uint16_t* v0_val, int32_t* v0_len, cudaStream_t st) {
  MPREID_REQUIRE(nbr_all && row_max_rows && v0_col && v0_val && v0_len, "rerank_build_v0: null pointer");
  MPREID_REQUIRE(R > 0 && N > 1 && R <= N && N < INT32_MAX, "rerank_build_v0: bad shape R=%lld N=%lld", (long long)R, (long long)N);
  MPREID_REQUIRE(k1 >= 1 && k1 <= kMaxK1 && K >= k1 + 1, "rerank_build_v0: k1 must be in [1, %d] and K >= k1+1", kMaxK1);
  const int sms = sm_count_of_current_device();
  const int v0_smem = v0_sizes(k1).bytes;
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_build_v0, cudaFuncAttributeMaxDynamicSharedMemorySize, v0_smem));
  const int Keff = (int)(K < N ? K : N);
  int v0_ctas = (200 * 1024) / (v0_smem + 1024);
  v0_ctas = v0_ctas > 16 ? 16 : (v0_ctas < 1 ? 1 : v0_ctas);   // 16 x 128 threads fill an SM
  const int64_t grid = R < (int64_t)sms * v0_ctas ? R : (int64_t)sms * v0_ctas;
  k_build_v0<<<(unsigned)grid, kV0Threads, v0_smem, st>>>(src, row_ids, (int)R, k1, K, Keff, nbr_all, row_max_rows, v0_col, v0_val, v0_len,
                                                          v0_capacity(k1, N));
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}
}  // namespace mpreid

extern "C" int mpreid_rerank_build_v0(const float* dist_rows, int64_t ld_dist, const int32_t* row_ids, int64_t R, int64_t N,
                                      int k1, const int32_t* nbr_all, int K, const float* row_max_rows,
                                      int32_t* v0_col, uint16_t* v0_val, int32_t* v0_len, void* stream) {
  MPREID_REQUIRE(dist_rows && ld_dist >= N, "rerank_build_v0: bad distance rows");
  V0Source src;
  memset(&src, 0, sizeof(src));
  src.dist = dist_rows; src.ld = ld_dist;
  return launch_build_v0(src, row_ids, R, N, k1, nbr_all, K, row_max_rows, v0_col, v0_val, v0_len, (cudaStream_t)stream);
}

extern "C" int mpreid_rerank_build_v0_sparse(const int32_t* row_ids, int64_t R, int64_t N, int k1, const int32_t* nbr_all,
                                             const float* nbr_val_all, int K, const float* row_max_rows,
                                             const float* xn, int64_t ld_xn, int64_t D, const float* sqnorm,
                                             int32_t* v0_col, uint16_t* v0_val, int32_t* v0_len, void* stream) {
  MPREID_REQUIRE(nbr_val_all && xn && sqnorm && D > 0 && ld_xn >= D && D < INT32_MAX, "rerank_build_v0_sparse: bad feature rows");
  V0Source src;
  memset(&src, 0, sizeof(src));
  src.nbr_val = nbr_val_all; src.xn = xn; src.ld_xn = ld_xn; src.D = (int)D; src.sqnorm = sqnorm;
  return launch_build_v0(src, row_ids, R, N, k1, nbr_all, K, row_max_rows, v0_col, v0_val, v0_len, (cudaStream_t)stream);
}

extern "C" size_t mpreid_rerank_finish_workspace_bytes(int64_t N, int64_t Q, int k1, int k2) {
  if (N <= 1 || Q <= 0 || Q >= N || k1 < 1 || k1 > kMaxK1 || k2 < 1 || k2 > 64) return 0;
  return carve_finish(nullptr, nullptr, N, Q, k1, k2, sm_count_of_current_device());
}

// Where the expanded V rows live inside the finish workspace (k2 != 1): byte offsets of v_col int32 [N, C1], v_val fp16
// [N, C1], v_len int32 [N], and C1 -- a sharded run expands its own rows (stage 8) and all-gathers the rest into place.
extern "C" int mpreid_rerank_finish_layout(int64_t N, int64_t Q, int k1, int k2, int64_t* out4) {
  MPREID_REQUIRE(out4 && N > 1 && Q > 0 && Q < N && k1 >= 1 && k1 <= kMaxK1 && k2 >= 2 && k2 <= 64, "rerank_finish_layout: bad arguments (k2 must be > 1)");
  FinishWs w;
  memset(&w, 0, sizeof(w));
  char* base = (char*)256;   // any non-null base: only the differences matter
  carve_finish(&w, base, N, Q, k1, k2, sm_count_of_current_device());
  out4[0] = (char*)w.v_col - base; out4[1] = (char*)w.v_val - base; out4[2] = (char*)w.v_len - base; out4[3] = w.C1;
  return MPREID_OK;
}

namespace mpreid {

static int launch_blend_default(const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* src_rows, const float* row_max_q,
                                int64_t Qs, int64_t G, float lambda_value, float* final_dist, int64_t ld_final, int ctas_per_sm, cudaStream_t st) {
  const int sms = sm_count_of_current_device();
  const bool vec = (((uintptr_t)(dist_q + col0) | (uintptr_t)final_dist) & 15) == 0 && ld_dist % 4 == 0 && ld_final % 4 == 0;
  const int64_t work = Qs * ceil_div(G, kBlendThreads * kBlendPer);
  // persistent grid; ctas_per_sm > 0 caps the footprint (a background launch that shares the SMs with latency-bound kernels)
  const int64_t per_sm = ctas_per_sm > 0 ? ctas_per_sm : 16;
  const int64_t grid = work < (int64_t)sms * per_sm ? work : (int64_t)sms * per_sm;
  if (vec) k_blend_default<true><<<(unsigned)grid, kBlendThreads, 0, st>>>(dist_q, ld_dist, col0, (int)Qs, (int)G, lambda_value, row_max_q, final_dist, ld_final, src_rows);
  else k_blend_default<false><<<(unsigned)grid, kBlendThreads, 0, st>>>(dist_q, ld_dist, col0, (int)Qs, (int)G, lambda_value, row_max_q, final_dist, ld_final, src_rows);
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

// :73-99 for the query rows dist_q[Qs, ...]: dist_q[il, col0 + c] is the distance of query row il to gallery sample c
static int rerank_finish_impl(const int32_t* nbr_all, int K, const int32_t* v0_col, const uint16_t* v0_val, const int32_t* v0_len,
                              const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* q_ids, const float* row_max_q,
                              int64_t N, int64_t Q, int64_t Qs, int k1, int k2, float lambda_value,
                              float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes, int stages, int64_t v0_stride,
                              int rows_global, int64_t qe_lo, int64_t qe_hi, cudaStream_t st) {
  MPREID_REQUIRE(nbr_all && v0_col && v0_val && v0_len && dist_q && row_max_q && final_dist && workspace, "rerank_finish: null pointer");
  MPREID_REQUIRE(stages >= 1 && stages <= 31, "rerank_finish: stages is a mask of 1 (expand + index), 2 (sparse Jaccard), 4 (default blend), 8 (expand rows [qe_lo, qe_hi) only), 16 (index only)");
  MPREID_REQUIRE(!(stages & 8) || (qe_lo >= 0 && qe_lo <= qe_hi && qe_hi <= N), "rerank_finish: bad query-expansion row range");
  MPREID_REQUIRE(!rows_global || q_ids, "rerank_finish: rows_global needs q_ids");
  MPREID_REQUIRE(N > 1 && Q > 0 && Q < N && Qs > 0 && Qs <= Q && N < INT32_MAX && ld_dist >= col0 + (N - Q) && ld_final >= N - Q,
                 "rerank_finish: bad shape N=%lld Q=%lld Qs=%lld", (long long)N, (long long)Q, (long long)Qs);
  MPREID_REQUIRE(k1 >= 1 && k1 <= kMaxK1 && k2 >= 1 && k2 <= 64 && K >= neighbor_count(k1, k2), "rerank_finish: bad k1/k2/K");
  MPREID_REQUIRE(((uintptr_t)workspace & 255) == 0, "rerank_finish: workspace must be 256-byte aligned");
  const int sms = sm_count_of_current_device();
  if (workspace_bytes < carve_finish(nullptr, nullptr, N, Q, k1, k2, sms)) {
    set_error("rerank_finish: workspace too small (%zu bytes)", workspace_bytes);
    return MPREID_ERR_WORKSPACE;
  }
  FinishWs w;
  memset(&w, 0, sizeof(w));
  carve_finish(&w, (char*)workspace, N, Q, k1, k2, sms);
  const int Keff = (int)(K < N ? K : N);
  const int32_t* v_col = v0_col; const uint16_t* v_val = v0_val; const int32_t* v_len = v0_len;
  if (k2 != 1) { v_col = w.v_col; v_val = w.v_val; v_len = w.v_len; }
  const int v0s = (int)(v0_stride > 0 ? v0_stride : w.C0);   // row stride of the V0 arrays (the capacity, or a trimmed width)
  const int64_t vstride = k2 != 1 ? w.C1 : (int64_t)v0s;      // row stride of V (== V0 when there is no query expansion)
  if (stages & (1 | 8)) {
    // :73-78  all N rows (stage 1: every rank expands everything), or the rows [qe_lo, qe_hi) only (stage 8: the sharded
    // form, the expanded rows are then all-gathered by the caller straight into the workspace arrays)
    if (k2 != 1) {
      const int64_t r_lo = (stages & 1) ? 0 : qe_lo, r_hi = (stages & 1) ? N : qe_hi;
      const int64_t n_rows = r_hi - r_lo;
      if (n_rows > 0) {
        // rows whose k2 gathered V0 rows hold <= 512 entries: one warp each; the rest (large k1 / k2): one CTA each
        const int k2e = k2 < Keff ? k2 : Keff;
        const int64_t wgrid = ceil_div(n_rows, kQeWarps) < (int64_t)sms * 6 ? ceil_div(n_rows, kQeWarps) : (int64_t)sms * 6;
        k_query_expand_warp<<<(unsigned)wgrid, kQeWarps * 32, 0, st>>>((int)r_hi, K, k2e, nbr_all, v0_col, v0_val, v0_len, v0s,
                                                                     w.v_col, w.v_val, w.v_len, w.C1, (int)r_lo);
        const int64_t qe_grid = n_rows < w.qe_grid ? n_rows : w.qe_grid;
        k_query_expand<<<(unsigned)qe_grid, kQeThreads, 0, st>>>((int)r_hi, K, k2e, nbr_all, v0_col, v0_val, v0_len, v0s,
                                                                 w.v_col, w.v_val, w.v_len, w.C1, w.qe_scratch, w.qe_P, kQeWarpEntries, (int)r_lo);
      }
    }
    MPREID_CUDA_CHECK(cudaGetLastError());
  }
  if (stages & (1 | 16)) {
    // :80-82 (gallery rows only: the output keeps columns Q.. only, :99)
    k_zero_i32<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(w.col_cnt, w.col_fill, N);
    const int rows_per_cta = 8;
    const unsigned csc_grid = (unsigned)ceil_div(N - Q, rows_per_cta);
    k_csc_count<<<csc_grid, rows_per_cta * 32, 0, st>>>((int)N, (int)Q, v_col, v_len, vstride, w.col_cnt);
    k_scan_i64<<<1, 1024, 0, st>>>(w.col_cnt, w.col_off, N);
    k_csc_fill<<<csc_grid, rows_per_cta * 32, 0, st>>>((int)N, (int)Q, v_col, v_val, v_len, vstride, w.col_off, w.col_fill, w.csc_row, w.csc_val);
    MPREID_CUDA_CHECK(cudaGetLastError());
  }
  if (!(stages & 6)) return MPREID_OK;
  // :84-99  dense default, then the sparse accumulation over the touched entries
  const int64_t G = N - Q;
  if (stages & 4) {
    const int rc = launch_blend_default(dist_q, ld_dist, col0, rows_global ? q_ids : nullptr, row_max_q, Qs, G, lambda_value, final_dist, ld_final, 0, st);
    if (rc != MPREID_OK) return rc;
  }
  if (!(stages & 2)) return MPREID_OK;
  const char* jac_env = getenv("MPREID_JACCARD");          // "tile": the per-step-barrier kernel (tests / comparison)
  if (jac_env && jac_env[0] == 't') {
    // tile: at most 41,600 gallery entries (81 KB) so that two 512-thread CTAs share an SM at MSMT17 size; a gallery
    // that needs several tiles is covered by several work items per query (each walks the lists once)
    const int64_t n_tiles = ceil_div(G, kJacMaxTile);
    int tile_cols = (int)ceil_div(G, n_tiles);
    tile_cols = (tile_cols + 7) & ~7;
    const int jac_fixed = (int)((sizeof(JacStage) + 15) & ~size_t(15)) + kJacEntries * 6;
    const int jac_smem = jac_fixed + tile_cols * 2;
    MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_jaccard_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, jac_fixed + kJacMaxTile * 2 + 16));
    int ctas_per_sm = (225 * 1024) / (jac_smem + 1024);
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
    const int jac_threads = ctas_per_sm >= 4 ? 256 : 512;
    const int64_t items = Qs * n_tiles;
    const int64_t jac_grid = items < (int64_t)sms * ctas_per_sm ? items : (int64_t)sms * ctas_per_sm;
    k_jaccard_sparse<<<(unsigned)jac_grid, jac_threads, jac_smem, st>>>(dist_q, ld_dist, col0, q_ids, (int)Qs, (int)N, (int)Q, lambda_value, row_max_q,
                                                                        v_col, v_val, v_len, vstride, w.col_off, w.csc_row, w.csc_val, final_dist,
                                                                        ld_final, tile_cols, (int)n_tiles, rows_global);
  } else {
    // bucket kernel: the whole gallery in one tile when it fits next to the 48 KB entry buffers (up to ~88,000 gallery
    // samples: one CTA per SM), else equal tiles; small galleries leave room for several CTAs per SM
    const int fixed = (int)((sizeof(JacBStage) + 15) & ~size_t(15)) + kJacRound * 6;
    const int max_tile = ((227 * 1024 - 1024 - fixed) / 2) & ~7;
    const int64_t n_tiles = ceil_div(G, max_tile);
    int tile_cols = (int)ceil_div(G, n_tiles);
    tile_cols = (tile_cols + 7) & ~7;
    const int smem = fixed + tile_cols * 2;
    MPREID_CUDA_CHECK(cudaFuncSetAttribute(k_jaccard_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, fixed + max_tile * 2));
    int ctas_per_sm = (227 * 1024) / (smem + 1024);
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
    const int64_t items = Qs * n_tiles;
    const int64_t grid = items < (int64_t)sms * ctas_per_sm ? items : (int64_t)sms * ctas_per_sm;
    k_jaccard_bucket<<<(unsigned)grid, kJacBWarps * 32, smem, st>>>(dist_q, ld_dist, col0, q_ids, (int)Qs, (int)N, (int)Q, lambda_value, row_max_q,
                                                                    v_col, v_val, v_len, vstride, w.col_off, w.csc_row, w.csc_val, final_dist,
                                                                    ld_final, tile_cols, (int)n_tiles, rows_global);
  }
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

}  // namespace mpreid

extern "C" int mpreid_rerank_finish(const int32_t* nbr_all, int K, const int32_t* v0_col, const uint16_t* v0_val, const int32_t* v0_len,
                                    const float* dist_qrows, int64_t ld_dist, const int32_t* q_ids, const float* row_max_q,
                                    int64_t N, int64_t Q, int64_t Qs, int k1, int k2, float lambda_value,
                                    float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes, void* stream) {
  return rerank_finish_impl(nbr_all, K, v0_col, v0_val, v0_len, dist_qrows, ld_dist, Q, q_ids, row_max_q, N, Q, Qs, k1, k2, lambda_value,
                            final_dist, ld_final, workspace, workspace_bytes, 7, 0, 0, 0, 0, (cudaStream_t)stream);
}

// The dense default blend alone (stage 4 of mpreid_rerank_finish_ex without any of the sparse-stage arguments).
extern "C" int mpreid_rerank_blend_default(const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* src_rows, const float* row_max_q,
                                           int64_t Qs, int64_t G, float lambda_value, float* final_dist, int64_t ld_final, int ctas_per_sm,
                                           void* stream) {
  MPREID_REQUIRE(dist_q && row_max_q && final_dist && Qs > 0 && G > 0 && col0 >= 0 && ld_dist >= col0 + G && ld_final >= G && Qs < INT32_MAX && G < INT32_MAX,
                 "rerank_blend_default: bad arguments");
  return launch_blend_default(dist_q, ld_dist, col0, src_rows, row_max_q, Qs, G, lambda_value, final_dist, ld_final, ctas_per_sm, (cudaStream_t)stream);
}

// General form: gallery sample 0 sits at column col0 of dist_q (col0 = Q for rows of the all-pairs matrix, 0 or the
// alignment pad for the [Qs, G] block the fused all-pairs pass keeps), and the two halves can run as separate calls
// on the same workspace (stages: mask of 1 = query expansion + inverted index, 2 = sparse Jaccard accumulation over the
// touched entries, 4 = the dense default blend; 4 depends on nothing but the distance block and may run early, on another stream).
extern "C" int mpreid_rerank_finish_ex(const int32_t* nbr_all, int K, const int32_t* v0_col, const uint16_t* v0_val, const int32_t* v0_len,
                                       const float* dist_q, int64_t ld_dist, int64_t col0, const int32_t* q_ids, const float* row_max_q,
                                       int64_t N, int64_t Q, int64_t Qs, int k1, int k2, float lambda_value,
                                       float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes, int stages,
                                       int64_t v0_stride, int rows_global, int64_t qe_lo, int64_t qe_hi, void* stream) {
  MPREID_REQUIRE(col0 >= 0 && v0_stride >= 0, "rerank_finish_ex: col0 / v0_stride < 0");
  return rerank_finish_impl(nbr_all, K, v0_col, v0_val, v0_len, dist_q, ld_dist, col0, q_ids, row_max_q, N, Q, Qs, k1, k2, lambda_value,
                            final_dist, ld_final, workspace, workspace_bytes, stages, v0_stride, rows_global, qe_lo, qe_hi, (cudaStream_t)stream);
}

// single-GPU convenience: neighbours + V0 rows + finish on the whole matrix
extern "C" size_t mpreid_rerank_workspace_bytes(int64_t N, int64_t Q, int k1, int k2) {
  if (N <= 1 || Q <= 0 || Q >= N || k1 < 1 || k1 > kMaxK1 || k2 < 1 || k2 > 64) return 0;
  const int K = neighbor_count(k1, k2), C0 = v0_capacity(k1, N);
  size_t off = 0;
  auto take = [&](size_t bytes) { off = align_up(off + bytes, 256); };
  take(N * 4); take(N * (size_t)K * 4); take(N * (size_t)C0 * 4); take(N * (size_t)C0 * 2); take(N * 4);
  return off + mpreid_rerank_finish_workspace_bytes(N, Q, k1, k2);
}

extern "C" int mpreid_rerank(const float* dist, int64_t ld_dist, const float* row_max_in, int64_t N, int64_t Q, int k1, int k2,
                             float lambda_value, float* final_dist, int64_t ld_final, void* workspace, size_t workspace_bytes,
                             int32_t* status, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MPREID_REQUIRE(dist && final_dist && workspace, "rerank: null pointer");
  MPREID_REQUIRE(N > 1 && Q > 0 && Q < N && N < INT32_MAX && ld_dist >= N && ld_final >= N - Q, "rerank: bad shape N=%lld Q=%lld",
                 (long long)N, (long long)Q);
  MPREID_REQUIRE(k1 >= 1 && k1 <= kMaxK1, "rerank: k1 must be in [1, %d]", kMaxK1);
  MPREID_REQUIRE(k2 >= 1 && k2 <= 64, "rerank: k2 must be in [1, 64]");
  MPREID_REQUIRE(((uintptr_t)workspace & 255) == 0, "rerank: workspace must be 256-byte aligned");
  const size_t need = mpreid_rerank_workspace_bytes(N, Q, k1, k2);
  if (need == 0 || workspace_bytes < need) {
    set_error("rerank: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return MPREID_ERR_WORKSPACE;
  }
  const int K = neighbor_count(k1, k2), C0 = v0_capacity(k1, N);
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base + off; off = align_up(off + bytes, 256); return p; };
  float* rowmax = (float*)take(N * 4);
  int32_t* nbr = (int32_t*)take(N * (size_t)K * 4);
  int32_t* v0_col = (int32_t*)take(N * (size_t)C0 * 4);
  uint16_t* v0_val = (uint16_t*)take(N * (size_t)C0 * 2);
  int32_t* v0_len = (int32_t*)take(N * 4);
  int rc;
  // :46-48  row max (== the reference's column max in this orientation) and the first K neighbours
  if (row_max_in) {
    MPREID_CUDA_CHECK(cudaMemcpyAsync(rowmax, row_max_in, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if ((rc = mpreid_row_max(dist, ld_dist, N, N, rowmax, stream)) != MPREID_OK) {
    return rc;
  }
  if ((rc = mpreid_row_topk(dist, ld_dist, N, N, K, rowmax, nbr, nullptr, stream)) != MPREID_OK) return rc;
  // :51-71
  if ((rc = mpreid_rerank_build_v0(dist, ld_dist, nullptr, N, N, k1, nbr, K, rowmax, v0_col, v0_val, v0_len, stream)) != MPREID_OK) return rc;
  // :73-99
  if ((rc = mpreid_rerank_finish(nbr, K, v0_col, v0_val, v0_len, dist, ld_dist, nullptr, rowmax, N, Q, Q, k1, k2, lambda_value,
                                 final_dist, ld_final, base + off, workspace_bytes - off, stream)) != MPREID_OK) return rc;
  if (status) MPREID_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  return MPREID_OK;
}
