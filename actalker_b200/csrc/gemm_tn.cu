// Dense projections of the layer on the tensor cores (C-ABI entry actk_gemm_tn_fwd): C = epi(A @ W^T).
//
// Replaces, on the 16-bit route, every nn.Linear / einsum of SS2D_cond_v10.forward (reference
// src/models/base/mamba_layer.py):
//   :1960, :1966, :1977   id_proj / audio_proj / exp_proj + SiLU          (epilogue = SiLU of the rounded product)
//   :1961, :1972          in_proj1 / in_proj2                             (one launch, x read once: stacked W, 2 output planes)
//   :1521                 x_proj  einsum("b k d l, k c d -> b k c l")     (both directions as one N = 4*16 + 2*Rp product)
//   :1523                 dt_proj einsum("b k r l, k d r -> b k d l")     (block-diagonal W over both directions)
//   :1985                 out_proj
// Same contraction and rounding points as the reference's GEMMs: fp32 accumulation, ONE rounding to the activation
// dtype (and for the condition tokens SiLU evaluated on that rounded tensor and rounded again).
//
// All of these are skinny (K <= 1024, N <= 5120, M = B'*L = 129 600 at BASELINE config 2) and therefore HBM-bound: the
// kernel is a streaming pipeline, not a flop machine.
//   persistent CTAs, one per SM, 320 threads:
//     warp 0     TMA producer: A (128 x 64) and W (BN x 64) slabs, SWIZZLE_128B, into a 4-6 deep shared-memory ring
//     warp 1     one thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 16; SASS UTCHMMA) into one of TWO
//                accumulator buffers in tensor memory (2 x 256 columns) and commits to the ring's / buffer's mbarriers
//     warps 2-9  epilogue, 64 output columns at a time: every warp reads 32 columns of its 32 rows from tensor memory
//                (tcgen05.ld / LDTM, lane = row), rounds (+ SiLU) and writes 16-byte pieces into a 128 x 64 staging tile
//                in the 128-byte-swizzled layout; one thread sends the tile with ONE TMA store (UTMASTG, 16 KB, full
//                128-byte lines); three staging tiles rotate so stores stay in flight (first version: one 32 x 32 box
//                per warp and chunk = 8x as many 2 KB stores of half lines: dt_proj 229 us against 125 us for cuBLAS).
//                The accumulator buffer is released as soon as it has been read, so tile i+1's MMAs run under tile
//                i's stores
// Up to 4 independent problems per launch (both branches, latent + tail tokens) share one grid through a tile table.
// Tails of M / N / K need no special code: TMA zero-fills loads and clips stores at the tensor bounds.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace actk {

constexpr int kGemmThreads = 320;
constexpr int kBM = 128;               // rows per tile == MMA M
constexpr int kBK = 64;                // K elements per slab: 128-byte rows, one SWIZZLE_128B atom wide
constexpr int kEpiWarps = 8;
constexpr uint32_t kABytes = kBM * kBK * 2;
constexpr uint32_t kEpiBufBytes = kBM * 64 * 2;        // one staging tile: 128 rows x 64 columns of 16-bit elements
constexpr int kMaxGemmProblems = ACTK_GEMM_MAX_PROBLEMS;

struct GemmProblemDev {
  CUtensorMap a, w, c;   // c: 64-column store boxes (SWIZZLE_128B)
  CUtensorMap c32;       // 32-column store boxes (SWIZZLE_64B) for the last chunk of a tile whose width is 32 mod 64
  CUtensorMap cf32;      // fp32 side output: boxes of 32 fp32 columns x 128 rows (SWIZZLE_128B)
  int f32_cols;          // leading output columns that are ALSO written, widened from the rounded result, to c_f32 (0 / 32 / 64)
  int n_tiles;        // column tiles per row tile
  int k_slabs;
  int bn;             // columns per tile (multiple of 32, <= 256)
  int tile_begin;     // index of this problem's first tile in the launch's tile table
  int plane_cols;     // output columns per plane (N when the output is one tensor)
  int M, N;
  int epilogue;       // ACTK_GEMM_EPI_* of this problem
};
struct alignas(64) GemmParams {
  GemmProblemDev p[kMaxGemmProblems];
  int n_problems, total_tiles, stages, epilogue;
  uint32_t stage_bytes;
  // Fused all-gather (multi-GPU, batch-first split of one call): problem 0's output tiles are stored to n_peers gather
  // buffers — every rank's, own included, mapped into this process over NVLink peer memory — instead of one tensor.
  CUtensorMap peer_c[ACTK_GEMM_MAX_PEERS];
  int n_peers;
  int epi_bufs;        // staging tiles of the epilogue (3, or 2 when the ring needs the room)
  int acc_bufs;        // accumulator buffers in tensor memory: 2 (tile i+1's MMAs under tile i's epilogue) or 1
  int acc_hstride;     // MH == 2: tensor-memory columns between the accumulators of a tile's two row halves
};

__device__ __forceinline__ uint64_t gemm_sw128_desc(uint32_t smem_addr) {   // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void gemm_tma_load_2d(uint32_t dst_smem, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void gemm_tma_store_3d(const void *tmap, int c0, int c1, int c2, uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(c2), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void gemm_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gemm_bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gemm_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gemm_bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
}
__device__ __forceinline__ void gemm_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gemm_commit(uint32_t bar) {   // arrives once every MMA issued so far by this thread is done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2): the two CTAs of a cluster work on one 256-row tile.  Each loads its own 128 rows of A and
// HALF of the W slab; the leader (cluster rank 0) issues M = 256 MMAs that read both CTAs' shared memory and write both
// CTAs' tensor memory, so a W byte is fetched from L2 once per 256 output rows.
__device__ __forceinline__ uint32_t gemm_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t gemm_leader_addr(uint32_t smem_addr) {   // the same offset in the leader CTA's shared memory
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_addr));
  return r;
}
__device__ __forceinline__ void gemm_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// load into THIS CTA's shared memory, bytes counted on the LEADER's mbarrier (bar: a shared::cluster address)
__device__ __forceinline__ void gemm_tma_load_2d_pair(uint32_t dst_smem, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void gemm_umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gemm_commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void gemm_bar_arrive_cluster(uint32_t bar) {   // bar: a shared::cluster address (gemm_leader_addr)
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gemm_tmem_ld16(uint32_t taddr, uint32_t *r) {   // no wait: see gemm_tmem_wait32
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void gemm_tmem_wait16(uint32_t *r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void gemm_tmem_wait32(uint32_t *r) {   // ties all 32 registers to the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// which problem owns tile t (at most 4 problems: a compare chain)
__device__ __forceinline__ int gemm_problem_of(const GemmParams &P, int t) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < kMaxGemmProblems; ++i)
    if (i < P.n_problems && t >= P.p[i].tile_begin) g = i;
  return g;
}

// MH: row halves per tile.  1 (default): 128-row tiles, two accumulator buffers.  2 (ACTK_GEMM_MH=2): 256-row tiles —
// every W slab feeds two M = 128 MMAs, which cuts the TMA fill traffic per multiply-add by a third (see launch_gemm for
// what that measured).
// F32SIDE: some problem of the launch has the fp32 side output (compile-time, like EPI: the write-bound dt_proj launch lost
// 20 us to the mere presence of the extra chunk loop in its epilogue).
// PAIR: two CTAs per 256-row tile (cta_group::2, above); launched as clusters of two, MH == 1.
template <typename T, int EPI, int MH, bool F32SIDE, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tn_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t gemm_smem[];
  const uint32_t base = (smem_u32(gemm_smem) + 1023u) & ~1023u;     // SWIZZLE_128B atoms want 1024-byte alignment
  const int S = P.stages;
  const uint32_t epi_base = base + (uint32_t)S * P.stage_bytes;
  const uint32_t bar_base = epi_base + (uint32_t)P.epi_bufs * kEpiBufBytes;
  constexpr uint32_t kATile = MH * kABytes;
  const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * S, tfull_bar = bar_base + 16 * S,
                 tempty_bar = tfull_bar + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gemm_smem + (tempty_bar + 16 - smem_u32(gemm_smem)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? gemm_cluster_rank() : 0;                   // 0: the CTA that issues the pair's MMAs

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { gemm_bar_init(full_bar + 8 * s, 1); gemm_bar_init(empty_bar + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { gemm_bar_init(tfull_bar + 8 * b, 1); gemm_bar_init(tempty_bar + 8 * b, PAIR ? 2 * kEpiWarps : kEpiWarps); }
    mbar_fence_init();
    for (int g = 0; g < P.n_problems; ++g) {
      tmap_prefetch(&P.p[g].a); tmap_prefetch(&P.p[g].w); tmap_prefetch(&P.p[g].c); tmap_prefetch(&P.p[g].c32);
      if (P.p[g].f32_cols) tmap_prefetch(&P.p[g].cf32);
    }
    for (int pe = 0; pe < P.n_peers; ++pe) tmap_prefetch(&P.peer_c[pe]);
  }
  __syncwarp();
  if (warp == 1) {   // 2 accumulator buffers x 256 columns (PAIR: warp 1 of both CTAs, the same columns in both)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) gemm_cluster_sync(); else __syncthreads();   // PAIR: the peer's barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------------- TMA producer
      uint32_t it = 0;
      for (int tile = (PAIR ? blockIdx.x >> 1 : blockIdx.x); tile < P.total_tiles; tile += (PAIR ? gridDim.x >> 1 : gridDim.x)) {
        const GemmProblemDev &pr = P.p[gemm_problem_of(P, tile)];
        const int local = tile - pr.tile_begin;
        const int m = local / pr.n_tiles, n = local - m * pr.n_tiles;
        const uint32_t tx = kATile + (PAIR ? (uint32_t)pr.bn >> 1 : (uint32_t)pr.bn) * (kBK * 2);   // PAIR: half of the W slab
        for (int ks = 0; ks < pr.k_slabs; ++ks, ++it) {
          const uint32_t s = it % S, use = it / S;
          if (use > 0) gemm_bar_wait(empty_bar + 8 * s, (use - 1) & 1);
          const uint32_t sa = base + s * P.stage_bytes;
          if (PAIR) {    // both CTAs' bytes are counted on the leader's barrier, which its own producer arms for both
            if (rank == 0) gemm_bar_expect_tx(full_bar + 8 * s, 2 * tx);
            const uint32_t lb = gemm_leader_addr(full_bar + 8 * s);
            gemm_tma_load_2d_pair(sa, &pr.a, ks * kBK, (2 * m + (int)rank) * kBM, lb);
            gemm_tma_load_2d_pair(sa + kATile, &pr.w, ks * kBK, n * pr.bn + (int)rank * (pr.bn >> 1), lb);
          } else {
            gemm_bar_expect_tx(full_bar + 8 * s, tx);
            gemm_tma_load_2d(sa, &pr.a, ks * kBK, m * (MH * kBM), full_bar + 8 * s);       // box of MH * 128 rows
            gemm_tma_load_2d(sa + kATile, &pr.w, ks * kBK, n * pr.bn, full_bar + 8 * s);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // ------------------------------------------------------- MMA issuer
      uint32_t it = 0, tc = 0;
      constexpr uint32_t fmt = IO<T>::is_bf16 ? 1u : 0u;
      for (int tile = (PAIR ? blockIdx.x >> 1 : blockIdx.x); tile < P.total_tiles; tile += (PAIR ? gridDim.x >> 1 : gridDim.x), ++tc) {
        const GemmProblemDev &pr = P.p[gemm_problem_of(P, tile)];
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(pr.bn >> 3) << 17) | ((uint32_t)((PAIR ? 2 * kBM : kBM) >> 4) << 24);
        const uint32_t b = P.acc_bufs == 2 ? (tc & 1) : 0, ub = P.acc_bufs == 2 ? (tc >> 1) : tc;
        if (ub > 0) gemm_bar_wait(tempty_bar + 8 * b, (ub - 1) & 1);   // the epilogue has read this buffer's previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem + b * 256;
        for (int ks = 0; ks < pr.k_slabs; ++ks, ++it) {
          const uint32_t s = it % S, use = it / S;
          gemm_bar_wait(full_bar + 8 * s, use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * P.stage_bytes, sw = sa + kATile;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
#pragma unroll
            for (int hh = 0; hh < MH; ++hh) {
              // PAIR, M = 256: rows 0-127 from this CTA's A slab, 128-255 from the peer's; the W halves likewise
              if (PAIR) gemm_umma_pair(acc, gemm_sw128_desc(sa + k * 32), gemm_sw128_desc(sw + k * 32), idesc, (uint32_t)((ks | k) != 0));
              else gemm_umma(acc + hh * P.acc_hstride, gemm_sw128_desc(sa + hh * kABytes + k * 32), gemm_sw128_desc(sw + k * 32), idesc,
                             (uint32_t)((ks | k) != 0));
            }
          // the slab may be overwritten once these MMAs have read it (PAIR: in both CTAs)
          if (PAIR) gemm_commit_pair(empty_bar + 8 * s); else gemm_commit(empty_bar + 8 * s);
        }
        // accumulators of this tile complete
        if (PAIR) gemm_commit_pair(tfull_bar + 8 * b); else gemm_commit(tfull_bar + 8 * b);
      }
    }
    __syncwarp();
  } else {             // ---------------------------------------------------------------- epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                  // tensor-memory lane quadrant this warp may read: warp id % 4
    const int h = ew >> 2;                   // which 32 of a chunk's 64 columns
    const int r = q * 32 + lane;             // row of the tile this thread converts
    const bool leader = ew == 0 && lane == 0;
    uint32_t tc = 0, chunk_count = 0;
    for (int tile = (PAIR ? blockIdx.x >> 1 : blockIdx.x); tile < P.total_tiles; tile += (PAIR ? gridDim.x >> 1 : gridDim.x), ++tc) {
      const GemmProblemDev &pr = P.p[gemm_problem_of(P, tile)];
      const int local = tile - pr.tile_begin;
      const int m = local / pr.n_tiles, n = local - m * pr.n_tiles;
      // column tiles run over the stacked product (all planes side by side)
      const int gcol0 = n * pr.bn;                                   // first column of the tile in the stacked product
      const int plane0 = gcol0 / pr.plane_cols;
      const int col0 = gcol0 - plane0 * pr.plane_cols;               // ... and within its plane
      const uint32_t b = P.acc_bufs == 2 ? (tc & 1) : 0, ub = P.acc_bufs == 2 ? (tc >> 1) : tc;
      gemm_bar_wait(tfull_bar + 8 * b, ub & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nchunks = (pr.bn + 63) >> 6;
      // fp32 side output (x_proj: the scan reads B|C as fp32 without widening them per tile): the leading f32_cols
      // columns of the first column tile go out a second time, as the ROUNDED result widened to fp32, in extra chunks of
      // 32 fp32 columns (128-byte staging rows, the same swizzle and the same staging-tile rotation as the 16-bit chunks)
      const int nf32 = (F32SIDE && MH == 1 && n == 0) ? (pr.f32_cols >> 5) : 0;
      for (int hc = 0; hc < MH * nchunks + nf32; ++hc, ++chunk_count) {
        const uint32_t buf = epi_base + (chunk_count % (uint32_t)P.epi_bufs) * kEpiBufBytes;
        if (F32SIDE && hc >= MH * nchunks) {
          const int v = hc - MH * nchunks;      // fp32 columns [32 v, 32 v + 32): every warp converts 16 of them for its rows
          const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + b * 256;
          uint32_t f16r[16];
          gemm_tmem_ld16(trow + v * 32 + h * 16, f16r);
          gemm_tmem_wait16(f16r);
#pragma unroll
          for (int i = 0; i < 16; ++i) f16r[i] = __float_as_uint(IO<T>::rnd(__uint_as_float(f16r[i])));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = buf + (uint32_t)r * 128 + (uint32_t)(((4 * h + j) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(f16r[4 * j]), "r"(f16r[4 * j + 1]),
                         "r"(f16r[4 * j + 2]), "r"(f16r[4 * j + 3])
                         : "memory");
          }
          fence_proxy_async();
          if (leader) {
            if (P.epi_bufs >= 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (leader) {
            if ((PAIR ? 2 * m + (int)rank : m) * kBM < pr.M) gemm_tma_store_3d(&pr.cf32, v * 32, (PAIR ? 2 * m + (int)rank : m) * kBM, 0, buf);
            bulk_commit();
          }
          continue;
        }
        const int hh = MH == 2 ? (hc >= nchunks ? 1 : 0) : 0, cc = hc - hh * nchunks;   // row half, 64-column chunk
        const int row0 = PAIR ? (2 * m + (int)rank) * kBM : (m * MH + hh) * kBM;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + b * 256 + hh * P.acc_hstride;
        const int width = pr.bn - cc * 64 >= 64 ? 64 : 32;          // bn is a multiple of 32
        if (h * 32 < width) {
          uint32_t v32[32];
          gemm_tmem_ld16(trow + cc * 64 + h * 32, v32);
          gemm_tmem_ld16(trow + cc * 64 + h * 32 + 16, v32 + 16);
          gemm_tmem_wait32(v32);
          // act(Linear(x)): the product is a `dtype` tensor first.  EPI says whether ANY problem of the launch wants it
          // (compile-time: launches without SiLU carry no trace of it), the problem's own flag decides per tile
          if (EPI == ACTK_GEMM_EPI_SILU && pr.epilogue == ACTK_GEMM_EPI_SILU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v32[i] = __float_as_uint(silu(IO<T>::rnd(__uint_as_float(v32[i]))));
          }
          uint4 v[4];
          T *e = reinterpret_cast<T *>(v);
#pragma unroll
          for (int i = 0; i < 32; ++i) IO<T>::st(e + i, __uint_as_float(v32[i]));
          // staging tile, row r: 64 columns = 128 bytes in the SWIZZLE_128B pattern (16-byte piece j at j ^ (r & 7)),
          // or, for a 32-column last chunk, 64 bytes in the SWIZZLE_64B pattern (piece j at j ^ ((r >> 1) & 3))
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = width == 64 ? buf + (uint32_t)r * 128 + (uint32_t)(((4 * h + j) ^ (r & 7)) << 4)
                                              : buf + (uint32_t)r * 64 + (uint32_t)((j ^ ((r >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w)
                         : "memory");
          }
          fence_proxy_async();
        }
        // the leader makes sure the staging tile of the NEXT chunk is free (its store, epi_bufs chunks back, has read it)
        // before anyone passes the barrier, then sends this chunk: one barrier per chunk
        if (leader) {
          if (P.epi_bufs >= 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        __syncwarp();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (leader) {
          if (gcol0 + cc * 64 < pr.N && row0 < pr.M) {
            // a tile may straddle two planes; each of its 64-column chunks lies in ONE of them (the host picks bn so)
            int plane = plane0, pcol = col0 + cc * 64;
            if (pcol >= pr.plane_cols) { pcol -= pr.plane_cols; ++plane; }
            if (P.n_peers > 0) {    // the same staging tile goes to every rank's buffer: GEMM and all-gather in one kernel
              for (int pe = 0; pe < P.n_peers; ++pe) gemm_tma_store_3d(&P.peer_c[pe], pcol, row0, plane, buf);
            } else {
              gemm_tma_store_3d(width == 64 ? &pr.c : &pr.c32, pcol, row0, plane, buf);
            }
          }
          // ONE group per chunk, also for a chunk that lies outside the output (empty group): the wait above counts
          // groups, and a chunk without one would let the tile two chunks on overwrite a staging tile whose stores are
          // still reading it (seen at 8 GPUs: the last-issued, slowest peer stores of the fused all-gather picked up the
          // next tile's data)
          bulk_commit();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {    // this warp's reads of the buffer are done (PAIR: the leader's MMA thread counts both CTAs' warps)
        if (PAIR) gemm_bar_arrive_cluster(gemm_leader_addr(tempty_bar + 8 * b)); else gemm_bar_arrive(tempty_bar + 8 * b);
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) gemm_cluster_sync(); else __syncthreads();   // PAIR: neither CTA leaves while the other may still signal it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*GemmEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static GemmEncodeFn gemm_encode_fn() {
  static GemmEncodeFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GemmEncodeFn>(p);
  }
  return fn;
}

// column-tile width: the fewest tiles of at most 256 columns, then the least padding; multiples of 32 (of 64 for the fused
// all-gather, whose per-rank store maps are 64 columns wide); when the output is split into planes a tile must not
// straddle two of them
static int gemm_bn_cap() {   // tuning: ACTK_GEMM_BN_MAX=128|192|256 caps the column-tile width
  const char *e = getenv("ACTK_GEMM_BN_MAX");
  const int v = e ? atoi(e) : 0;
  return (v >= 64 && v <= 256) ? v / 32 * 32 : 256;
}

static int gemm_pick_bn(int N, int plane_cols, int k_slabs, int step = 32) {
  // Column tiles run over the stacked product.  Every 64-column chunk of a tile must lie in one plane: either tiles do not
  // straddle planes (bn divides plane_cols), or planes are multiples of 64 columns at least one tile wide and tiles begin
  // at multiples of 64 (a tile then straddles at most two planes).  The second form only for products with several K slabs
  // per tile: in_proj at d_model 320 (two planes of 640 columns) takes 256- instead of 160-column tiles, 140 -> 116 us.
  const int span = N;
  const bool free_planes = plane_cols != N && plane_cols % 64 == 0 && plane_cols >= 256 && k_slabs >= 2;
  if (free_planes && step < 64) step = 64;
  int best = step, best_tiles = 1 << 30, best_pad = 1 << 30;
  for (int bn = gemm_bn_cap() / step * step; bn >= step; bn -= step) {
    if (plane_cols != N && !free_planes && plane_cols % bn != 0) continue;
    const int tiles = (span + bn - 1) / bn, pad = tiles * bn - span;
    if (tiles < best_tiles || (tiles == best_tiles && pad < best_pad)) { best = bn; best_tiles = tiles; best_pad = pad; }
  }
  return best;
}

static const char *gemm_check(const actk_gemm_problem &p, int es) {
  if (p.n_peers < 0 || p.n_peers > ACTK_GEMM_MAX_PEERS) return "n_peers out of range";
  if (!p.a || !p.w || (!p.c && p.n_peers == 0)) return "NULL pointer";
  for (int i = 0; i < p.n_peers; ++i) {
    if (!p.peer_c[i]) return "NULL pointer (peer buffer)";
    if (reinterpret_cast<uintptr_t>(p.peer_c[i]) & 15) return "peer pointer not aligned to 16 bytes";
  }
  if (p.n_peers > 0 && (p.planes != 1 || p.N % 64 != 0)) return "the fused all-gather needs one plane and N a multiple of 64";
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return "non-positive size";
  if (p.lda < p.K || p.ldw < p.K) return "row pitch smaller than K";
  if (p.planes < 1 || p.N % p.planes != 0) return "N is not a multiple of planes";
  const int pc = p.N / p.planes;
  if (p.ldc < pc) return "ldc smaller than the columns of a plane";
  if (p.planes > 1 && pc % 32 != 0) return "columns per plane must be a multiple of 32 when planes > 1";
  if ((p.lda * es) % 16 || (p.ldw * es) % 16 || (p.ldc * es) % 16 || (p.plane_stride * es) % 16) return "row pitch not a multiple of 16 bytes";
  if ((reinterpret_cast<uintptr_t>(p.a) | reinterpret_cast<uintptr_t>(p.w) | (p.n_peers ? 0 : reinterpret_cast<uintptr_t>(p.c))) & 15)
    return "pointer not aligned to 16 bytes";
  if (p.f32_cols != 0) {
    if (!p.c_f32) return "NULL pointer (fp32 side output)";
    if ((p.f32_cols != 32 && p.f32_cols != 64) || p.f32_cols > p.N || p.planes != 1 || p.n_peers != 0)
      return "the fp32 side output covers the first 32 or 64 columns of a one-plane product";
    if (p.ldc_f32 < p.f32_cols || (p.ldc_f32 * 4) % 16 || (reinterpret_cast<uintptr_t>(p.c_f32) & 15))
      return "fp32 side output: row pitch / pointer not a multiple of 16 bytes";
  }
  return nullptr;
}

// MH from the environment for tuning (ACTK_GEMM_MH=1|2), else by problem size
static int gemm_forced_mh() {
  const char *e = getenv("ACTK_GEMM_MH");
  return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
}

template <typename T, int MH, bool PAIR>
static int launch_gemm_mh(const actk_gemm_problem *pr, int n, int dtype, int sms, int smem_max, int dev, cudaStream_t stream) {
  GemmEncodeFn fn = gemm_encode_fn();
  if (!fn) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available from this driver");
  const CUtensorMapDataType dt = dtype == ACTK_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.n_problems = n;
  int epilogue = ACTK_GEMM_EPI_NONE;
  for (int g = 0; g < n; ++g)
    if (pr[g].epilogue == ACTK_GEMM_EPI_SILU) epilogue = ACTK_GEMM_EPI_SILU;
  P.epilogue = epilogue;
  static_assert(!PAIR || MH == 1, "a CTA pair works on 2 x 128 rows");
  constexpr int kRows = (PAIR ? 2 : MH) * kBM;          // rows per tile (PAIR: 128 per CTA)
  int tiles = 0, bn_max = 32;
  for (int g = 0; g < n; ++g) {
    const actk_gemm_problem &p = pr[g];
    GemmProblemDev &d = P.p[g];
    const int pc = p.N / p.planes;
    d.bn = gemm_pick_bn(p.N, pc, (p.K + kBK - 1) / kBK, p.n_peers > 0 ? 64 : 32);
    d.plane_cols = pc;
    d.M = p.M; d.N = p.N;
    d.epilogue = p.epilogue;
    d.k_slabs = (p.K + kBK - 1) / kBK;
    d.n_tiles = (p.N + d.bn - 1) / d.bn;
    d.tile_begin = tiles;
    tiles += ((p.M + kRows - 1) / kRows) * d.n_tiles;
    bn_max = d.bn > bn_max ? d.bn : bn_max;
    const cuuint32_t estr[3] = {1, 1, 1};
    {   // A (K, M): slabs of 64 columns x (MH * 128) rows
      cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
      cuuint64_t strides[1] = {(cuuint64_t)p.lda * sizeof(T)};
      cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(MH * kBM)};
      CUresult r = fn(&d.a, dt, 2, const_cast<void *>(p.a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled (A of problem %d) failed with CUresult %d", g, (int)r);
    }
    {   // W (K, N): slabs of 64 x bn; with planes the tile (plane, j) covers rows plane*pc + j*bn of the stacked weight
      cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
      cuuint64_t strides[1] = {(cuuint64_t)p.ldw * sizeof(T)};
      cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(PAIR ? d.bn / 2 : d.bn)};      // PAIR: each CTA fetches half the rows
      CUresult r = fn(&d.w, dt, 2, const_cast<void *>(p.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled (W of problem %d) failed with CUresult %d", g, (int)r);
    }
    if (p.n_peers > 0) {   // one 64-column store map per rank's gather buffer (same shape and pitch as the local output)
      if (g != 0 || n != 1) ACTK_FAIL(ACTK_ERR_BAD_ARG, "gemm_tn: the fused all-gather takes one problem per launch");
      P.n_peers = p.n_peers;
      for (int pe = 0; pe < p.n_peers; ++pe) {
        cuuint64_t dims[3] = {(cuuint64_t)pc, (cuuint64_t)p.M, 1};
        cuuint64_t strides[2] = {(cuuint64_t)p.ldc * sizeof(T), (cuuint64_t)p.ldc * p.M * sizeof(T)};
        cuuint32_t box[3] = {64u, (cuuint32_t)kBM, 1};
        CUresult r = fn(&P.peer_c[pe], dt, 3, p.peer_c[pe], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled (peer %d) failed with CUresult %d", pe, (int)r);
      }
    }
    d.f32_cols = 0;
    if (p.f32_cols != 0) {   // fp32 side output (M, f32_cols): boxes of 32 columns x 128 rows
      if (MH != 1 || d.bn < p.f32_cols) ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "gemm_tn: the fp32 side output needs 128-row tiles of at least %d columns", p.f32_cols);
      d.f32_cols = p.f32_cols;
      cuuint64_t dims[3] = {(cuuint64_t)p.f32_cols, (cuuint64_t)p.M, 1};
      cuuint64_t strides[2] = {(cuuint64_t)p.ldc_f32 * 4, (cuuint64_t)p.ldc_f32 * 4 * p.M};
      cuuint32_t box[3] = {32u, (cuuint32_t)kBM, 1};
      CUresult r = fn(&d.cf32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.c_f32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled (fp32 side output of problem %d) failed with CUresult %d", g, (int)r);
    }
    for (int narrow = 0; narrow < 2 && p.n_peers == 0; ++narrow) {   // C (plane_cols, M, planes): 128-row store boxes of 64 / 32 columns
      cuuint64_t dims[3] = {(cuuint64_t)pc, (cuuint64_t)p.M, (cuuint64_t)p.planes};
      cuuint64_t strides[2] = {(cuuint64_t)p.ldc * sizeof(T),
                               (cuuint64_t)(p.planes > 1 ? p.plane_stride : (long long)p.ldc * p.M) * sizeof(T)};
      cuuint32_t box[3] = {narrow ? 32u : 64u, (cuuint32_t)kBM, 1};
      CUresult r = fn(narrow ? &d.c32 : &d.c, dt, 3, p.c, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled (C of problem %d) failed with CUresult %d", g, (int)r);
    }
  }
  P.total_tiles = tiles;
  P.stage_bytes = MH * kABytes + (uint32_t)(PAIR ? bn_max / 2 : bn_max) * (kBK * 2);
  // tensor memory: 512 columns.  Two accumulator buffers when a tile's accumulators fit in 256 columns, else one.
  if (MH == 1) { P.acc_bufs = 2; P.acc_hstride = 0; }
  else if (bn_max <= 128) { P.acc_bufs = 2; P.acc_hstride = 128; }
  else { P.acc_bufs = 1; P.acc_hstride = 256; }
  // shared memory: ring stages + epilogue staging tiles; three staging tiles unless that leaves fewer than three stages
  const size_t bars = 16 * 6 + 64;
  auto stages_for = [&](int bufs) { return (int)(((size_t)smem_max - 1024 - (size_t)bufs * kEpiBufBytes - bars) / P.stage_bytes); };
  // Products with several K slabs per tile (in_proj, x_proj, out_proj: K >= 256) gain from a deeper ring — measured on B200,
  // in_proj with 2 / 3 / 4 / 5 stages: 212 / 168 / 143 / 135 us — so they trade the third staging tile for a fifth stage; the
  // write-bound dt_proj (one K slab) keeps three staging tiles (149 vs 156 us).  ACTK_GEMM_EPI_BUFS / ACTK_GEMM_STAGES: tuning.
  int k_slabs_max = 1;
  for (int g = 0; g < n; ++g) k_slabs_max = P.p[g].k_slabs > k_slabs_max ? P.p[g].k_slabs : k_slabs_max;
  P.epi_bufs = stages_for(3) >= 3 ? 3 : 2;
  if (k_slabs_max >= 4 && P.n_peers == 0 && stages_for(2) > stages_for(3) && stages_for(3) < 6) P.epi_bufs = 2;
  if (const char *e = getenv("ACTK_GEMM_EPI_BUFS")) {
    if (e[0] == '2' || e[0] == '3') P.epi_bufs = e[0] - '0';
  }
  int stages = stages_for(P.epi_bufs);
  stages = stages > 6 ? 6 : stages;
  if (const char *e = getenv("ACTK_GEMM_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v < stages) stages = v;
  }
  if (stages < 2) ACTK_FAIL(ACTK_ERR_CUDA, "gemm_tn: %d bytes of shared memory per block do not hold two pipeline stages", smem_max);
  P.stages = stages;
  const size_t smem = 1024 + (size_t)stages * P.stage_bytes + (size_t)P.epi_bufs * kEpiBufBytes + 16 * (size_t)stages + 64;
  bool f32side = false;
  for (int g = 0; g < n; ++g) f32side = f32side || P.p[g].f32_cols != 0;
  auto kern = epilogue == ACTK_GEMM_EPI_SILU ? gemm_tn_kernel<T, ACTK_GEMM_EPI_SILU, MH, false, PAIR> : gemm_tn_kernel<T, ACTK_GEMM_EPI_NONE, MH, false, PAIR>;
  if (f32side) kern = epilogue == ACTK_GEMM_EPI_SILU ? gemm_tn_kernel<T, ACTK_GEMM_EPI_SILU, MH, true, PAIR> : gemm_tn_kernel<T, ACTK_GEMM_EPI_NONE, MH, true, PAIR>;
  static int configured[64][4] = {};   // per device and kernel variant: dynamic shared memory limit raised
  const int ei = (epilogue == ACTK_GEMM_EPI_SILU ? 1 : 0) + (f32side ? 2 : 0);
  if (dev < 64 && !configured[dev][ei]) {
    ACTK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    configured[dev][ei] = 1;
  }
  if (PAIR) {   // clusters of two CTAs (one TPC): one pair per 256-row tile
    const int pairs = sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ACTK_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, P));
  } else {
    kern<<<tiles < sms ? tiles : sms, kGemmThreads, smem, stream>>>(P);
  }
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

template <typename T>
static int launch_gemm(const actk_gemm_problem *pr, int n, int dtype, cudaStream_t stream) {
  int dev = 0, sms = 0, smem_max = 0;
  ACTK_CUDA_OK(cudaGetDevice(&dev));
  ACTK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ACTK_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // 128-row tiles by default.  256-row tiles (ACTK_GEMM_MH=2) were built to halve the W fill traffic and measured on B200
  // (profiles/r02_gemm_tn_mh.txt): out_proj 71.9 -> 68.4 us, x_proj 76.6 -> 74.8, but in_proj 142.7 -> 148.8 and dt_proj
  // 128 -> 171 us (one accumulator buffer: the epilogue no longer hides under the next tile's MMAs).  What bounds the
  // K >= 320 products is the MMAs' own operand reads from shared memory (12 KB per M128 N256 K16 instruction), which a taller
  // tile does not change.
  // CTA pairs (cta_group::2: clusters of two, every W byte fetched from L2 once per 256 rows) are opt-in (ACTK_GEMM_PAIR=1):
  // built to cut the L2 -> SM traffic of the K >= 320 products by 30-45 %, parity-green and bit-identical, and measured SLOWER
  // on B200 at config 2 (tools/bench_gemm_tn.py, us, single / pair): in_proj 115-126 / 150-172, dt_proj 144 / 202, out_proj
  // 70.5 / 72.7, x_proj 76.7 / 80.0; only x_proj at d_model 1280 gains (54.7 -> 50.2).  A pair advances at the pace of its
  // slower CTA at every ring slot and every accumulator hand-over, and these launches are not short of L2 bandwidth.
  bool pair = false;
  if (const char *e = getenv("ACTK_GEMM_PAIR")) {
    if (e[0] == '1') {
      pair = true;
      for (int g = 0; g < n; ++g) if (pr[g].n_peers > 0) pair = false;     // the fused all-gather keeps single CTAs
    }
  }
  if (pair) return launch_gemm_mh<T, 1, true>(pr, n, dtype, sms, smem_max, dev, stream);
  const int mh = gemm_forced_mh() == 2 ? 2 : 1;
  return mh == 2 ? launch_gemm_mh<T, 2, false>(pr, n, dtype, sms, smem_max, dev, stream)
                 : launch_gemm_mh<T, 1, false>(pr, n, dtype, sms, smem_max, dev, stream);
}

}  // namespace actk

using namespace actk;

extern "C" int actk_gemm_tn_supported(const actk_gemm_problem *p, int dtype) {
  if (!p || (dtype != ACTK_F16 && dtype != ACTK_BF16)) return 0;
  return gemm_check(*p, 2) == nullptr;
}

extern "C" int actk_gemm_tn_fwd(const actk_gemm_problem *problems, int n_problems, int dtype, void *stream) {
  if (!problems) ACTK_FAIL(ACTK_ERR_BAD_ARG, "gemm_tn: problems is NULL");
  if (n_problems < 1 || n_problems > kMaxGemmProblems)
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "gemm_tn: n_problems=%d (1..%d per launch)", n_problems, kMaxGemmProblems);
  if (dtype != ACTK_F16 && dtype != ACTK_BF16)
    ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "gemm_tn: dtype=%d (the tensor-core route is f16 / bf16; fp32 projections stay with the caller)", dtype);
  for (int g = 0; g < n_problems; ++g) {
    const actk_gemm_problem &p = problems[g];
    if (p.epilogue != ACTK_GEMM_EPI_NONE && p.epilogue != ACTK_GEMM_EPI_SILU)
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "gemm_tn: problem %d: epilogue=%d", g, p.epilogue);
    const char *why = gemm_check(p, 2);
    if (why) {
      const bool align = strstr(why, "16 bytes") != nullptr;
      ACTK_FAIL(align ? ACTK_ERR_BAD_ALIGN : (strstr(why, "NULL") ? ACTK_ERR_BAD_ARG : ACTK_ERR_BAD_SHAPE),
                "gemm_tn: problem %d (M=%d N=%d K=%d lda=%lld ldw=%lld ldc=%lld planes=%d): %s", g, p.M, p.N, p.K, p.lda, p.ldw,
                p.ldc, p.planes, why);
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ACTK_F16) return launch_gemm<__half>(problems, n_problems, dtype, st);
  return launch_gemm<__nv_bfloat16>(problems, n_problems, dtype, st);
}
