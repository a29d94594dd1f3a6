// Distance matrix on the 5th-generation tensor cores (sm_100a): TMA -> shared memory -> tcgen05.mma
// -> TMEM -> fused distance epilogue.  Replaces the torch-CPU sgemm of utils/metrics.py:12,17 and
// utils/reranking.py:40.
//
//   out[i, j] = epilogue( sum_k q[i,k] * g[j,k] )         both operands K-major ("NT" GEMM)
//
// Precision modes
//   3xTF32  the fp32 operands are pre-split (prep.cu) into TF32-exact planes  x = hi + lo;  the
//           accumulator receives  lo*hi + hi*lo + hi*hi  (error ~2^-21 relative per product, i.e.
//           fp32-level; the lo*lo term is below fp32 rounding).  Three kind::tf32 MMAs per k-step.
//   3xFP16  (default) the same split on per-row power-of-two scaled fp16 planes: twice the MMA rate.
//   2xFP16  stated fast mode: the gallery side contributes only its hi plane (two MMAs per k-step).
//   BF16    bf16-rounded operands, one kind::f16 MMA per k-step, fp32 accumulate.
//
// Kernel shape: persistent, one CTA per SM, 192 threads = 6 warps
//   warp 0      TMA producer (one elected lane): 128B-swizzled boxes of 128 B x {128|256} rows
//   warp 1      TMEM allocator + MMA issuer (one elected lane), 128x256 fp32 accumulator in TMEM,
//               double buffered (2 x 256 of the 512 columns) so the epilogue of tile i overlaps the
//               MMAs of tile i+1
//   warps 2..5  epilogue: tcgen05.ld 32 columns at a time, norm add / arccos / 1-dot, row max,
//               store.  Warp w owns TMEM lanes 32*(w%4)..+31 (hardware restriction), i.e. one output
//               row per thread.
// Tiles are visited band-major (m fastest inside a band of query blocks sized to ~24 MB of operand planes)
// so the query band stays in L2 and every gallery tile streams past once per band.
//
// CTA-pair variant (template parameter CTA2, rectangular launches of the split-fp16 modes): clusters of two
// CTAs on one TPC share a 256x256 tile through tcgen05.mma.cta_group::2.  Each CTA stages its own 128
// query rows and half of the gallery rows (64 KB per k-block -> three stages), the leader CTA issues the
// MMAs and multicast-commits to the barriers of both CTAs; each CTA's epilogue warps read their own 128
// accumulator lanes.  Per output element the accumulation order is the single-CTA one: results are
// bit-identical (tests/test_gpu_parity.py::test_cta_pair_gemm_bit_identical_to_single_cta).
#include <cuda.h>
#include <cstdlib>

#include "epilogue.cuh"

namespace mpreid {
namespace tc {

static constexpr int BM = 128;       // UMMA_M
static constexpr int BN = 256;       // UMMA_N
static constexpr int THREADS = 192;
static constexpr int TMEM_COLS = 512;
static constexpr int BAND = 16;      // m-blocks per L2 band

template <int PREC> struct Cfg;
template <> struct Cfg<MPREID_3XTF32> {
  static constexpr int PLANES = 2, ELEM = 4, UMMA_K = 8, STAGES = 2, FMT = 2 /*TF32*/;
};
template <> struct Cfg<MPREID_BF16> {
  static constexpr int PLANES = 1, ELEM = 2, UMMA_K = 16, STAGES = 4, FMT = 1 /*BF16*/;
};
template <> struct Cfg<MPREID_3XFP16> {
  static constexpr int PLANES = 2, ELEM = 2, UMMA_K = 16, STAGES = 2, FMT = 0 /*F16*/;
};
// 2xFP16: the query side keeps the hi/lo split, the gallery side only its hi plane (11 bits): drops the
// hi*lo term -> two MMAs per k-step, 64 KB stages x 3.  Error ~2^-12 per product: a stated fast mode.
template <> struct Cfg<MPREID_2XFP16> {
  static constexpr int PLANES = 2, ELEM = 2, UMMA_K = 16, STAGES = 3, FMT = 0 /*F16*/;
};
template <int PREC> struct PlanesB { static constexpr int value = Cfg<PREC>::PLANES; };
template <> struct PlanesB<MPREID_2XFP16> { static constexpr int value = 1; };

// per-row-width pipeline shape: 64-byte rows halve the stage and double the depth
template <int PREC, int ROWB> struct Pipe {
  static constexpr int PLANES = Cfg<PREC>::PLANES, PLANES_B = PlanesB<PREC>::value;   // operand planes: query side, gallery side
  static constexpr int ELEM = Cfg<PREC>::ELEM, UMMA_K = Cfg<PREC>::UMMA_K, FMT = Cfg<PREC>::FMT;
  static constexpr int STAGES = Cfg<PREC>::STAGES * (128 / ROWB);
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) flavours: two CTAs of a cluster on one TPC share one 256x256 tile; each loads
//      its own 128 query rows and HALF of the gallery rows, the leader (cluster rank 0) issues the MMAs
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address) in the CTA with cluster rank `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// data lands in this CTA's shared memory, the transaction bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int PREC>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == MPREID_3XTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}

template <int PREC>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == MPREID_3XTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major operand, 128B swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
// ROWB = 64 is the same with 64-byte rows / 64B swizzle (8-row groups 512 B apart, layout = 4).
template <int ROWB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(((8 * ROWB) >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;
  return d;
}

// instruction descriptor: c_format F32 (bit 4), a/b format (bits 7-9 / 10-12), K-major both,
// N>>3 at bits 17-22, M>>4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TileCoord { int m_blk, n_blk; };
__device__ __forceinline__ TileCoord decode_tile(int tile, int m_blocks, int n_blocks, int band = BAND) {
  const int band_tiles = band * n_blocks;
  const int b = tile / band_tiles;
  const int rem = tile - b * band_tiles;
  const int h = min(band, m_blocks - b * band);
  TileCoord t;
  t.n_blk = rem / h;
  t.m_blk = b * band + (rem - t.n_blk * h);
  return t;
}

// Symmetric (all-pairs) mode visits only the tiles with n_blk >= m_blk/2 (on or right of the diagonal block
// column), enumerated band by band in the same order as above so that the round-robin over CTAs stays balanced.
__host__ __device__ __forceinline__ int sym_band_tiles(int b, int m_blocks, int n_blocks, int band, int* h_out, int* jt_out) {
  const int m0 = b * band;
  const int h = (band < m_blocks - m0) ? band : (m_blocks - m0);
  const int avail = n_blocks - (m0 >> 1);          // gallery blocks from the band's first diagonal block on
  const int jf = (h - 1) / 2;                      // leading columns that hold fewer than h valid tiles: 2, 4, ...
  const int jt = avail < jf ? (avail < 0 ? 0 : avail) : jf;
  if (h_out) { *h_out = h; *jt_out = jt; }
  const int flat = avail - jf;
  return jt * (jt + 1) + (flat > 0 ? flat * h : 0);
}
__host__ __device__ __forceinline__ int sym_total_tiles(int m_blocks, int n_blocks, int band) {
  int t = 0;
  for (int b = 0; b * band < m_blocks; ++b) t += sym_band_tiles(b, m_blocks, n_blocks, band, nullptr, nullptr);
  return t;
}
__device__ __forceinline__ TileCoord sym_decode_tile(int tile, int m_blocks, int n_blocks, int band) {
  int b = 0, h = 0, jt = 0;
  for (;; ++b) {
    const int tb = sym_band_tiles(b, m_blocks, n_blocks, band, &h, &jt);
    if (tile < tb) break;
    tile -= tb;
  }
  const int m0 = b * band, n0 = m0 >> 1;
  TileCoord t;
  if (tile < jt * (jt + 1)) {
    int j = 0;
    while (tile >= 2 * j + 2) { tile -= 2 * j + 2; ++j; }
    t.n_blk = n0 + j; t.m_blk = m0 + tile;
  } else {
    tile -= jt * (jt + 1);
    t.n_blk = n0 + jt + tile / h; t.m_blk = m0 + tile % h;
  }
  return t;
}

// CTA-pair symmetric mode: square 256 x 256 pair tiles (mp, n) with n >= mp, bands of BAND/2 pair rows;
// inside a band the first h columns form a triangle (column j holds rows 0..j), the rest are full.
__host__ __device__ __forceinline__ int sym_pair_band_tiles(int b, int Mp, int band, int* h_out) {
  const int PB = band / 2;
  const int p0 = b * PB;
  const int h = (PB < Mp - p0) ? PB : (Mp - p0);
  if (h_out) *h_out = h;
  return h * (h + 1) / 2 + (Mp - p0 - h) * h;
}
__host__ __device__ __forceinline__ int sym_pair_total(int Mp, int band) {
  int t = 0;
  for (int b = 0; b * (band / 2) < Mp; ++b) t += sym_pair_band_tiles(b, Mp, band, nullptr);
  return t;
}
// -> m_blk = pair row (units of 256 rows), n_blk = column block
__device__ __forceinline__ TileCoord sym_pair_decode(int pt, int Mp, int band) {
  int b = 0, h = 0;
  for (;; ++b) {
    const int tb = sym_pair_band_tiles(b, Mp, band, &h);
    if (pt < tb) break;
    pt -= tb;
  }
  const int p0 = b * (band / 2);
  TileCoord t;
  const int tri = h * (h + 1) / 2;
  if (pt < tri) {
    int j = 0;
    while (pt >= j + 1) { pt -= j + 1; ++j; }
    t.n_blk = p0 + j; t.m_blk = p0 + pt;
  } else {
    pt -= tri;
    t.n_blk = p0 + h + pt / h; t.m_blk = p0 + pt % h;
  }
  return t;
}

// tile -> (128-row query block, 256-column gallery block) of THIS CTA
template <bool CTA2>
__device__ __forceinline__ TileCoord tile_coord(int tile, int cta_rank, int m_blocks, int n_blocks, int symmetric, int band) {
  if (!CTA2) return symmetric ? sym_decode_tile(tile, m_blocks, n_blocks, band) : decode_tile(tile, m_blocks, n_blocks, band);
  if (!symmetric) return decode_tile(tile + cta_rank, m_blocks, n_blocks, band);
  TileCoord t = sym_pair_decode(tile >> 1, m_blocks >> 1, band);
  t.m_blk = 2 * t.m_blk + cta_rank;
  return t;
}

// The same for a role that visits tile indices in INCREASING order (every loop of the kernel does): the symmetric
// enumerations are decoded band by band, and restarting at band 0 for every tile costs O(bands) each time -- with the
// multi-GPU ownership filter a CTA decodes eight times more tiles than it contracts, and that walk showed up as ~0.9 ms
// of an 8-GPU all-pairs pass.  The walker remembers the band it is in.
struct TileWalk {
  int b = 0, base = 0;   // current band and the index of its first (pair) tile
  template <bool CTA2>
  __device__ __forceinline__ TileCoord at(int tile, int cta_rank, int m_blocks, int n_blocks, int symmetric, int band) {
    if (!symmetric) return tile_coord<CTA2>(tile, cta_rank, m_blocks, n_blocks, 0, band);
    TileCoord t;
    if (!CTA2) {
      int h = 0, jt = 0;
      for (;;) {
        const int tb = sym_band_tiles(b, m_blocks, n_blocks, band, &h, &jt);
        if (tile - base < tb) break;
        base += tb; ++b;
      }
      int local = tile - base;
      const int m0 = b * band, n0 = m0 >> 1;
      if (local < jt * (jt + 1)) {
        int j = 0;
        while (local >= 2 * j + 2) { local -= 2 * j + 2; ++j; }
        t.n_blk = n0 + j; t.m_blk = m0 + local;
      } else {
        local -= jt * (jt + 1);
        t.n_blk = n0 + jt + local / h; t.m_blk = m0 + local % h;
      }
    } else {
      const int Mp = m_blocks >> 1;
      int h = 0;
      int pt = tile >> 1;
      for (;;) {
        const int tb = sym_pair_band_tiles(b, Mp, band, &h);
        if (pt - base < tb) break;
        base += tb; ++b;
      }
      pt -= base;
      const int p0 = b * (band / 2);
      const int tri = h * (h + 1) / 2;
      if (pt < tri) {
        int j = 0;
        while (pt >= j + 1) { pt -= j + 1; ++j; }
        t.n_blk = p0 + j; t.m_blk = p0 + pt;
      } else {
        pt -= tri;
        t.n_blk = p0 + h + pt / h; t.m_blk = p0 + pt % h;
      }
      t.m_blk = 2 * t.m_blk + cta_rank;
    }
    return t;
  }
};

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

static constexpr int kFuseQueue = 160;   // per-warp staging queue (entries); flushed with one atomic round per 32 entries

// CTA-pair pipeline depth: as many stages as fit 192 KB (3 x 64 KB for the split-fp16 / TF32 modes)
__host__ __device__ constexpr int pair_stages(int stage_bytes) { return (196608 / stage_bytes) < 8 ? (196608 / stage_bytes) : 8; }

template <int PREC, int ROW_BYTES, int METRIC, bool VEC, bool CTA2, bool FUSE>
__global__ void __launch_bounds__(THREADS, 1)
k_dist_tc(const __grid_constant__ Maps maps, const float* __restrict__ q_aux, const float* __restrict__ g_aux,
          const float* __restrict__ q_scale, const float* __restrict__ g_scale, int Q, int G, int num_k_blocks,
          float* __restrict__ out, int64_t ld_out,
          float* __restrict__ row_max, int m_blocks, int n_blocks, int symmetric, int band, const TopkFuse fuse) {
  using C = Pipe<PREC, ROW_BYTES>;
  constexpr int BN_LOCAL = CTA2 ? BN / 2 : BN;            // gallery rows THIS CTA stages per k-block
  constexpr int A_PLANE = BM * ROW_BYTES, B_PLANE = BN_LOCAL * ROW_BYTES;
  constexpr int STAGE_BYTES = C::PLANES * A_PLANE + C::PLANES_B * B_PLANE;   // per CTA
  constexpr int NSTAGES = CTA2 ? pair_stages(STAGE_BYTES) : C::STAGES;
  constexpr int K_PER_BLOCK = ROW_BYTES / C::ELEM;       // elements of K per stage
  constexpr int K_STEPS = K_PER_BLOCK / C::UMMA_K;       // 4
  constexpr uint32_t IDESC = make_idesc(C::FMT, CTA2 ? 2 * BM : BM, BN);
  constexpr bool kScaled = PREC == MPREID_3XFP16 || PREC == MPREID_2XFP16;   // operands carry per-row 2^s scales

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024-B aligned tiles
  const uint32_t bar_base = base + NSTAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * NSTAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * NSTAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGES + 4);
  const uint32_t gvec_base = bar_base + 256u;              // 2 x 256 float2: per-column (aux, scale) of the tile
  const uint32_t stage_base = gvec_base + 2u * BN * 8u;    // 4 warps x 4 KB output staging
  const uint32_t gthr_base = stage_base + 4u * 4096u;      // FUSE: 2 x 256 column thresholds, then 4 per-warp candidate queues
  const uint32_t fq_base = gthr_base + 2u * BN * 4u;       //       (kFuseQueue x 8 B entries + kFuseQueue x 4 B rows each)
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / tmem-empty barriers); the pair takes the
  // tiles (2p, 2p+1) of the band-major order, which are vertically adjacent (band heights are even)
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  const int tile_first = CTA2 ? (int)((blockIdx.x >> 1) << 1) : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)(gridDim.x & ~1u) : (int)gridDim.x;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // CTA pairs count tiles in units of one CTA (two per pair tile) so that the loops below are shared
  const int total_tiles = symmetric ? (CTA2 ? 2 * sym_pair_total(m_blocks >> 1, band) : sym_total_tiles(m_blocks, n_blocks, band)) : m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a_hi); prefetch_tmap(&maps.b_hi);
    if (C::PLANES == 2) prefetch_tmap(&maps.a_lo);
    if (C::PLANES_B == 2) prefetch_tmap(&maps.b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CTA2 ? 8 : 4); }
      fence_barrier_init();
    }
    __syncwarp();
    if (CTA2) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();   // the peer's barriers must exist before anything targets them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      TileWalk walk;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        // symmetric mode: tiles below the diagonal block column are never visited, the mirrors fill them
        const TileCoord t = walk.at<CTA2>(tile, (int)cta_rank, m_blocks, n_blocks, symmetric, band);
        if (FUSE && fuse.own_mod > 1 && ((t.m_blk >> 1) % fuse.own_mod) != fuse.own_rank) continue;   // another rank's row block
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint32_t sb = sa + C::PLANES * A_PLANE;
          const int kc = kb * K_PER_BLOCK;
          if (CTA2) {
            // both CTAs' bytes are counted on the leader's barrier (armed by the leader alone)
            if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            const uint32_t lbar = map_to_cta(full_bar(stage), 0u);
            const int brow = t.n_blk * BN + (int)cta_rank * BN_LOCAL;
            tma_load_2d_pair(sa, &maps.a_hi, lbar, kc, t.m_blk * BM);
            if (C::PLANES == 2) tma_load_2d_pair(sa + A_PLANE, &maps.a_lo, lbar, kc, t.m_blk * BM);
            tma_load_2d_pair(sb, &maps.b_hi, lbar, kc, brow);
            if (C::PLANES_B == 2) tma_load_2d_pair(sb + B_PLANE, &maps.b_lo, lbar, kc, brow);
          } else {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(sa, &maps.a_hi, full_bar(stage), kc, t.m_blk * BM);
            if (C::PLANES == 2) tma_load_2d(sa + A_PLANE, &maps.a_lo, full_bar(stage), kc, t.m_blk * BM);
            tma_load_2d(sb, &maps.b_hi, full_bar(stage), kc, t.n_blk * BN);
            if (C::PLANES_B == 2) tma_load_2d(sb + B_PLANE, &maps.b_lo, full_bar(stage), kc, t.n_blk * BN);
          }
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (CTA pair: the leader only) ================================
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    TileWalk walk;
    for (int tile = tile_first; leader && tile < total_tiles; tile += tile_step) {
      if (FUSE && fuse.own_mod > 1) {
        const TileCoord t = walk.at<CTA2>(tile, 0, m_blocks, n_blocks, symmetric, band);
        if (((t.m_blk >> 1) % fuse.own_mod) != fuse.own_rank) continue;
      }
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      ++it;
      mbar_wait(tempty_bar(as), aph ^ 1u);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint32_t sb = sa + C::PLANES * A_PLANE;
          const uint64_t a_hi = make_smem_desc<ROW_BYTES>(sa), b_hi = make_smem_desc<ROW_BYTES>(sb);
          auto mma = [&](uint64_t ad, uint64_t bd, uint32_t acc) {
            if (CTA2) umma_pair<PREC>(tmem_d, ad, bd, IDESC, acc); else umma<PREC>(tmem_d, ad, bd, IDESC, acc);
          };
#pragma unroll
          for (int k = 0; k < K_STEPS; ++k) {
            const uint64_t koff = (uint64_t)((k * C::UMMA_K * C::ELEM) >> 4);  // +32 B per k-step inside the atom
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (C::PLANES == 2 && C::PLANES_B == 2) {
              const uint64_t a_lo = make_smem_desc<ROW_BYTES>(sa + A_PLANE), b_lo = make_smem_desc<ROW_BYTES>(sb + B_PLANE);
              mma(a_lo + koff, b_hi + koff, acc);
              mma(a_hi + koff, b_lo + koff, 1u);
              mma(a_hi + koff, b_hi + koff, 1u);
            } else if (C::PLANES == 2) {
              const uint64_t a_lo = make_smem_desc<ROW_BYTES>(sa + A_PLANE);
              mma(a_lo + koff, b_hi + koff, acc);
              mma(a_hi + koff, b_hi + koff, 1u);
            } else {
              mma(a_hi + koff, b_hi + koff, acc);
            }
          }
          // smem slot reusable (in both CTAs) once these MMAs retire; accumulator complete after the last k-block
          if (CTA2) {
            umma_commit_pair(empty_bar(stage));
            if (kb == num_k_blocks - 1) umma_commit_pair(tfull_bar(as));
          } else {
            umma_commit(empty_bar(stage));
            if (kb == num_k_blocks - 1) umma_commit(tfull_bar(as));
          }
        }
        __syncwarp();
        if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue ================================
    // Per tile: (1) the 4 epilogue warps stage the tile's 256 gallery terms (norm, 2^-s scale) in
    // shared memory once; (2) TMEM is read 32 columns at a time with the NEXT chunk's tcgen05.ld in
    // flight while the current one is finished; (3) results go through a per-warp 32x32 swizzled
    // staging tile so that every global store instruction writes four full 128-byte lines.
    const int quad = warp & 3;
    const int et = (warp - 2) * 32 + lane;            // 0..127 among the epilogue threads
    float* stage = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) + (warp - 2) * 1024;
    float2* gvec_all = reinterpret_cast<float2*>(smem_raw + (gvec_base - smem_u32(smem_raw)));
    // FUSE: column thresholds of the tile and this warp's candidate queue
    float* gthr_all = reinterpret_cast<float*>(smem_raw + (gthr_base - smem_u32(smem_raw)));
    unsigned long long* fq_ent = reinterpret_cast<unsigned long long*>(smem_raw + (fq_base - smem_u32(smem_raw))) + (warp - 2) * kFuseQueue;
    int* fq_row = reinterpret_cast<int*>(smem_raw + (fq_base - smem_u32(smem_raw)) + 4 * kFuseQueue * 8) + (warp - 2) * kFuseQueue;
    int fq_count = 0;                                  // warp-uniform
    // flush: all atomics of the queue are issued before the first dependent store (one memory round trip per flush)
    auto fq_flush = [&]() {
      constexpr int kRounds = kFuseQueue / 32;
      int slot[kRounds];
#pragma unroll
      for (int i = 0; i < kRounds; ++i) {
        const int e = lane + 32 * i;
        slot[i] = e < fq_count ? atomicAdd(fuse.cand_cnt + fq_row[e], 1) : 0x7fffffff;
      }
#pragma unroll
      for (int i = 0; i < kRounds; ++i) {
        const int e = lane + 32 * i;
        if (slot[i] < fuse.cap) fuse.cand[(int64_t)fq_row[e] * fuse.cap + slot[i]] = fq_ent[e];
      }
      __syncwarp();
      fq_count = 0;
    };
    // value j of this lane's row in the swizzled staging tile (stage: lane == row, 16-byte chunk c at c ^ (row & 7))
    auto staged = [&](int j) { return stage[lane * 32 + ((((j >> 2) ^ (lane & 7)) << 2) | (j & 3))]; };
    // append this lane's hits: bit j of mrow = element (row_a, col0 + j) qualifies for row row_a; bit j of mcol = it
    // qualifies for the mirrored row col0 + j (column row_a).  The values are read back from the staging tile.
    auto fq_push = [&](uint32_t mrow, uint32_t mcol, int row_a, int col0) {
      const int n = __popc(mrow) + __popc(mcol);
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) return;
      if (total > kFuseQueue) {
        // thresholds that pass (almost) everything: no staging, every lane appends on its own (correct, slow; the
        // candidate lists overflow in this regime anyway and the caller falls back to the materialising path)
        while (mrow) {
          const int j = __ffs(mrow) - 1; mrow &= mrow - 1;
          const int slot = atomicAdd(fuse.cand_cnt + row_a, 1);
          if (slot < fuse.cap) fuse.cand[(int64_t)row_a * fuse.cap + slot] = ((unsigned long long)(uint32_t)(col0 + j) << 32) | __float_as_uint(staged(j));
        }
        while (mcol) {
          const int j = __ffs(mcol) - 1; mcol &= mcol - 1;
          const int slot = atomicAdd(fuse.cand_cnt + col0 + j, 1);
          if (slot < fuse.cap) fuse.cand[(int64_t)(col0 + j) * fuse.cap + slot] = ((unsigned long long)(uint32_t)row_a << 32) | __float_as_uint(staged(j));
        }
        return;
      }
      if (fq_count + total > kFuseQueue) fq_flush();
      int pos = fq_count + incl - n;
      while (mrow) {
        const int j = __ffs(mrow) - 1; mrow &= mrow - 1;
        fq_ent[pos] = ((unsigned long long)(uint32_t)(col0 + j) << 32) | __float_as_uint(staged(j)); fq_row[pos] = row_a; ++pos;
      }
      while (mcol) {
        const int j = __ffs(mcol) - 1; mcol &= mcol - 1;
        fq_ent[pos] = ((unsigned long long)(uint32_t)row_a << 32) | __float_as_uint(staged(j)); fq_row[pos] = col0 + j; ++pos;
      }
      fq_count += total;
    };
    int it = 0;
    TileWalk walk;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const TileCoord t = walk.at<CTA2>(tile, (int)cta_rank, m_blocks, n_blocks, symmetric, band);
      if (FUSE && fuse.own_mod > 1 && ((t.m_blk >> 1) % fuse.own_mod) != fuse.own_rank) continue;
      // symmetric (all-pairs) mode: a tile strictly right of the diagonal block column also writes its
      // transpose, which is exactly the set of tiles skipped above; diagonal tiles (n == m/2) do not
      const bool mirror = symmetric && t.n_blk > (t.m_blk >> 1);
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      ++it;
      const int gm0 = t.m_blk * BM + quad * 32;
      const int gm = gm0 + lane;
      const bool row_ok = gm < Q;
      const float qa = (row_ok && q_aux) ? q_aux[gm] : 0.f;
      const float qs = (kScaled && row_ok) ? q_scale[gm] : 1.f;
      const float thr_r = (FUSE && row_ok) ? __ldg(fuse.thr + gm) : -INFINITY;
      float rmax = -INFINITY;
      float2* gvec = gvec_all + as * BN;
      float* gthr = gthr_all + as * BN;
#pragma unroll
      for (int c = et; c < BN; c += 128) {
        const int gn = t.n_blk * BN + c;
        float2 v = make_float2(1.f, 1.f);
        float th = -INFINITY;
        if (gn < G) {
          if (g_aux) v.x = __ldg(g_aux + gn);
          if (kScaled) v.y = __ldg(g_scale + gn);
          if (FUSE) th = __ldg(fuse.thr + gn);
        }
        gvec[c] = v;
        if (FUSE) gthr[c] = th;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // gvec visible to the 4 epilogue warps
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        // 64 accumulator columns per TMEM load; the wait follows the load directly (an asynchronous
        // tcgen05.ld must not stay in flight across compiler-scheduled code: its destination
        // registers may have been re-assigned by then)
        uint32_t acc[64];
        tmem_ld64(taddr0 + (uint32_t)c0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int cc = c0 + half * 32;
          const int gn0 = t.n_blk * BN + cc;
          if (gm0 < Q && gn0 < G) {   // warp-uniform
            float d[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 gv = gvec[cc + j];                       // broadcast read
              float dot = __uint_as_float(acc[half * 32 + j]);
              if (kScaled) dot = dot * qs * gv.y;     // undo the 2^s row scales (exact)
              d[j] = finish_distance<METRIC>(dot, qa, gv.x);
              if (gn0 + j < G) rmax = fmaxf(rmax, d[j]);
            }
            // FUSE: only the [Q, G] block is stored (warp-uniform test), shifted so that 32-column groups stay aligned
            const bool store_direct = !FUSE || (gm0 < fuse.keep_rows && gn0 >= fuse.keep_col0);
            if (FUSE || store_direct) {
              // stage: lane == row, 16-byte chunk c of the row goes to position c ^ (row & 7)
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(stage + lane * 32 + ((c ^ (lane & 7)) << 2)) =
                    make_float4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
              __syncwarp();
            }
            if (store_direct) {
              const int row_lim = FUSE ? min(Q, fuse.keep_rows) : Q;
              const int col_shift = FUSE ? fuse.keep_col0 : 0;
              // drain: 8 lanes cover one 128-byte row, 4 rows per instruction
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + (lane >> 3), c = lane & 7;
                const float4 v = *reinterpret_cast<const float4*>(stage + r * 32 + ((c ^ (r & 7)) << 2));
                const int grow = gm0 + r, gcol = gn0 + 4 * c;
                if (grow < row_lim) {
                  float* dst = out + (int64_t)grow * ld_out + (gcol - col_shift);
                  if (VEC && gcol + 4 <= G) {
                    __stcs(reinterpret_cast<float4*>(dst), v);   // streaming: the matrix must not evict the operand band from L2
                  } else {
                    if (gcol < G) dst[0] = v.x;
                    if (gcol + 1 < G) dst[1] = v.y;
                    if (gcol + 2 < G) dst[2] = v.z;
                    if (gcol + 3 < G) dst[3] = v.w;
                  }
                }
              }
              if (!FUSE) __syncwarp();
            }
            if (FUSE) {
              // candidates: two instructions per element and side (set + lop3); the ~1 % hits go through the warp queue
              const uint32_t colmask = gn0 + 32 <= G ? 0xffffffffu : ((1u << (G - gn0)) - 1u);
              uint32_t mrow = 0, mcol = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                uint32_t a, b;
                asm("set.le.u32.f32 %0, %1, %2;" : "=r"(a) : "f"(d[j]), "f"(thr_r));
                asm("set.le.u32.f32 %0, %1, %2;" : "=r"(b) : "f"(d[j]), "f"(gthr[cc + j]));
                mrow |= a & (1u << j);
                mcol |= b & (1u << j);
              }
              mrow &= colmask;
              mcol = (mirror && row_ok) ? (mcol & colmask) : 0u;
              fq_push(mrow, mcol, gm, gn0);
              __syncwarp();   // the staging tile is rewritten by the next chunk
            }
            if (mirror) {
              // transpose: for a fixed column the 32 lanes hold 32 consecutive rows -> one 128-byte store
              if (!FUSE) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (row_ok && gn0 + j < G) __stcs(out + (int64_t)(gn0 + j) * ld_out + gm, d[j]);
              }
              if (row_max) {
                // column maxima (the row maxima of the mirrored block): butterfly transpose-reduce, 31 shuffles
#pragma unroll
                for (int j = 0; j < 32; ++j) if (!row_ok) d[j] = -INFINITY;
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                  const bool upper = (lane & sft) != 0;
#pragma unroll
                  for (int tt = 0; tt < sft; ++tt) {
                    const float send = upper ? d[tt] : d[tt + sft];
                    const float keep = upper ? d[tt + sft] : d[tt];
                    d[tt] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, sft));
                  }
                }
                if (gn0 + lane < G && d[0] > -INFINITY) atomic_max_f32(&row_max[gn0 + lane], d[0]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(map_to_cta(tempty_bar(as), 0u));   // the leader's MMA warp waits for all 8 epilogue warps
        else mbar_arrive(tempty_bar(as));
      }
      if (row_max && row_ok && rmax > -INFINITY) atomic_max_f32(&row_max[gm], rmax);
    }
    if (FUSE && fq_count) fq_flush();
  }
  tc_fence_before();
  if (CTA2) {
    cluster_sync_all();   // neither CTA may leave while the other can still touch its barriers / shared memory
    if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

static int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t ldk, int elem, int box_rows, int rowb) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return MPREID_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldk * elem};
  cuuint32_t box[2] = {(cuuint32_t)(rowb / elem), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return MPREID_ERR_CUDA; }
  return MPREID_OK;
}

template <int PREC, int ROW_BYTES, bool CTA2>
static int launch(const void* qa, const void* qb, const void* ga, const void* gb, const float* q_aux, const float* g_aux,
                  const float* q_scale, const float* g_scale, int64_t Q, int64_t G, int64_t ldk, int metric, float* out, int64_t ld_out, float* row_max,
                  int symmetric, cudaStream_t st, const TopkFuse* fuse_in) {
  using C = Cfg<PREC>;
  constexpr int kpb = ROW_BYTES / C::ELEM;
  constexpr int BN_LOCAL = CTA2 ? BN / 2 : BN;
  MPREID_REQUIRE(ldk % kpb == 0, "dist_tc: operand planes must be padded to a multiple of %d elements (got %lld)", kpb, (long long)ldk);
  MPREID_REQUIRE(((uintptr_t)qa & 15) == 0 && ((uintptr_t)ga & 15) == 0, "dist_tc: operands must be 16-byte aligned");
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = make_map(&maps.a_hi, qa, Q, ldk, C::ELEM, BM, ROW_BYTES)) != MPREID_OK) return rc;
  if ((rc = make_map(&maps.b_hi, ga, G, ldk, C::ELEM, BN_LOCAL, ROW_BYTES)) != MPREID_OK) return rc;
  if (C::PLANES == 2 && (rc = make_map(&maps.a_lo, qb, Q, ldk, C::ELEM, BM, ROW_BYTES)) != MPREID_OK) return rc;
  if (PlanesB<PREC>::value == 2 && (rc = make_map(&maps.b_lo, gb, G, ldk, C::ELEM, BN_LOCAL, ROW_BYTES)) != MPREID_OK) return rc;
  int m_blocks = (int)ceil_div(Q, BM);
  if (CTA2) m_blocks += m_blocks & 1;   // pairs take two vertically adjacent 128-row blocks; a padding block is all out of range
  const int n_blocks = (int)ceil_div(G, BN);
  MPREID_REQUIRE((int64_t)m_blocks * n_blocks < INT32_MAX, "dist_tc: too many tiles");
  constexpr int STAGE_BYTES = (C::PLANES * BM + PlanesB<PREC>::value * BN_LOCAL) * ROW_BYTES;
  constexpr int NSTAGES = CTA2 ? pair_stages(STAGE_BYTES) : C::STAGES * (128 / ROW_BYTES);
  const bool fused = fuse_in != nullptr;
  MPREID_REQUIRE(!fused || (symmetric && metric == MPREID_SQEUCLID), "dist_tc: the fused top-k mode is the symmetric squared-euclidean launch");
  TopkFuse fuse;
  memset(&fuse, 0, sizeof(fuse));
  if (fused) fuse = *fuse_in;
  const int smem = NSTAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + 2 * BN * 8 /*gvec*/ + 4 * 4096 /*staging*/ +
                   (fused ? 2 * BN * 4 /*column thresholds*/ + 4 * kFuseQueue * 12 /*candidate queues*/ : 0);
  const bool vec = (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0);
  using KernT = decltype(&k_dist_tc<PREC, ROW_BYTES, MPREID_SQEUCLID, true, CTA2, false>);   // no casts: a signature mismatch must not compile
  KernT kern = nullptr;
#define MPREID_PICK(M) kern = vec ? &k_dist_tc<PREC, ROW_BYTES, M, true, CTA2, false> : &k_dist_tc<PREC, ROW_BYTES, M, false, CTA2, false>
  switch (metric) {
    case MPREID_SQEUCLID: MPREID_PICK(MPREID_SQEUCLID); break;
    case MPREID_ARCCOS: MPREID_PICK(MPREID_ARCCOS); break;
    case MPREID_ONE_MINUS_DOT: MPREID_PICK(MPREID_ONE_MINUS_DOT); break;
    case MPREID_DOT: MPREID_PICK(MPREID_DOT); break;
    default: MPREID_PICK(MPREID_SQRT_EUCLID); break;
  }
#undef MPREID_PICK
  if (fused) kern = vec ? &k_dist_tc<PREC, ROW_BYTES, MPREID_SQEUCLID, true, CTA2, true> : &k_dist_tc<PREC, ROW_BYTES, MPREID_SQEUCLID, false, CTA2, true>;
  MPREID_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int nkb = (int)(ldk / kpb);
  // Query-band height of the tile order (m fastest inside a band): the band's operand planes must stay L2-resident
  // while every gallery tile streams past once per band, so fewer / taller bands mean fewer gallery re-reads.  Budget
  // 24 MB for the band (measured at MSMT17 shape, CTA pairs: DRAM reads 6.8 GB at 16 blocks, 3.3 GB at 32, same time),
  // balanced over the bands, even (CTA pairs take two vertically adjacent blocks).
  const int64_t block_bytes = (int64_t)BM * ldk * C::ELEM * C::PLANES;
  int band = (int)((24ll << 20) / (block_bytes > 0 ? block_bytes : 1));
  band = band < BAND ? BAND : band;
  const int n_bands = (m_blocks + band - 1) / band;
  band = (m_blocks + n_bands - 1) / n_bands;
  band += band & 1;
  if (const char* band_env = getenv("MPREID_GEMM_BAND")) { const int b = atoi(band_env); if (b >= 2 && !(b & 1)) band = b; }
  const int Qi = (int)Q, Gi = (int)G;
  const int64_t total = symmetric ? (CTA2 ? 2 * (int64_t)sym_pair_total(m_blocks >> 1, band) : sym_total_tiles(m_blocks, n_blocks, band)) : (int64_t)m_blocks * n_blocks;
  const int sms = sm_count_of_current_device();
  int grid = (int)(total < sms ? total : sms);
  if (CTA2) grid &= ~1;
  if (CTA2) {
    MPREID_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MPREID_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, maps, q_aux, g_aux, q_scale, g_scale, Qi, Gi, nkb, out, ld_out, row_max,
                                         m_blocks, n_blocks, symmetric, band, fuse));
  } else {
    kern<<<grid, THREADS, smem, st>>>(maps, q_aux, g_aux, q_scale, g_scale, Qi, Gi, nkb, out, ld_out, row_max, m_blocks, n_blocks, symmetric, band, fuse);
  }
  MPREID_CUDA_CHECK(cudaGetLastError());
  return MPREID_OK;
}

}  // namespace tc

int launch_dist_tc(const void* qa, const void* qb, const void* ga, const void* gb, const float* q_aux, const float* g_aux,
                   const float* q_scale, const float* g_scale, int64_t Q, int64_t G, int64_t ldk, int metric, int precision, float* out, int64_t ld_out,
                   float* row_max, int symmetric, cudaStream_t st, const TopkFuse* fuse) {
  int dev = 0, major = 0;
  MPREID_CUDA_CHECK(cudaGetDevice(&dev));
  MPREID_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("dist_tc: the tcgen05 kernels need an sm_100 GPU (found compute capability %d.x)", major);
    return MPREID_ERR_UNSUPPORTED;
  }
  // pipeline shape: 128-byte smem rows (128B swizzle).  A 64-byte-row / twice-as-deep variant (template parameter
  // ROW_BYTES = 64) was measured 7 % slower at MSMT17 shape (6.18 vs 5.78 ms) and is not instantiated.
#define MPREID_ARGS qa, qb, ga, gb, q_aux, g_aux, q_scale, g_scale, Q, G, ldk, metric, out, ld_out, row_max, symmetric, st, fuse
  // CTA pairs (cta_group::2, 256x256 tile per pair, 3 x 64 KB stages per CTA) for the rectangular GEMM of the
  // split-fp16 modes (rectangular and symmetric all-pairs); MPREID_GEMM_PAIR=0 selects the single-CTA kernel.
  const char* pair_env = getenv("MPREID_GEMM_PAIR");   // read per call: tests flip it inside one process
  const bool pair_ok = !(pair_env && pair_env[0] == '0');
  // symmetric all-pairs launches are long enough to sit at the power cap either way; the pair kernel measured 3 %
  // slower there (46.7 vs 45.0 ms for the MSMT17 re-ranking pass), so it is used only on request (MPREID_GEMM_PAIR=2)
  const bool pair = pair_ok && Q > tc::BM && (!symmetric || (pair_env && pair_env[0] == '2'));
  if (precision == MPREID_3XTF32) return tc::launch<MPREID_3XTF32, 128, false>(MPREID_ARGS);
  if (precision == MPREID_3XFP16) return pair ? tc::launch<MPREID_3XFP16, 128, true>(MPREID_ARGS) : tc::launch<MPREID_3XFP16, 128, false>(MPREID_ARGS);
  if (precision == MPREID_2XFP16) return pair ? tc::launch<MPREID_2XFP16, 128, true>(MPREID_ARGS) : tc::launch<MPREID_2XFP16, 128, false>(MPREID_ARGS);
  return tc::launch<MPREID_BF16, 128, false>(MPREID_ARGS);
#undef MPREID_ARGS
}

}  // namespace mpreid
