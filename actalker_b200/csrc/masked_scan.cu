// Fused masked bidirectional selective scan (C-ABI entry actk_masked_scan_fwd).
//
// Replaces, for one call of SS2D_cond_v10.forward (reference src/models/base/mamba_layer.py):
//   :1963/:1974  gather of the mask-selected tokens           -> tiles are fetched through idx[] (TMA boxes / cp.async rows)
//   :1965-1967   cat([selected, id, cond])                    -> tail rows come from a second base pointer
//   :1508-1519   HSCANS_dynamic identity encode + flip + cat  -> direction 1 walks the same rows downwards
//   :1532-1538   selective_scan_fn (bias, softplus, scan, D)  -> ChannelScan::step, fp32 state in registers
//   :1969-1970   slice [:n_sel] + index_put_ scatter          -> y of position p is stored to latent row idx[p]
// The direction sum (:1542-1547) and branch sum (:1983) need the reference's rounding points and are done by
// actk_merge_layernorm_fwd, which reads the two per-direction outputs written here.
//
// Work decomposition (B200, 148 SMs): one CTA = 64 channels x one (batch, branch, direction), one thread per
// channel with its 16 states in registers; config 2 (B'=25, D=640) gives 1000 CTAs = 6.8 per SM, all resident.
// Time is cut into 16-step tiles staged through a shared-memory ring:
//   * FAST tiles (16 selected tokens whose latent rows are consecutive — every tile under the all-ones masks
//     the shipped pipeline feeds, Inference.py:545-546): one elected thread issues three 3-D tensor-map TMA
//     loads (u, delta, B|C boxes; SASS UTMALDG) completing on the stage's mbarrier, and the y tile goes back
//     with one TMA store (UTMASTG);
//   * RAGGED tiles (mask edges, the id/cond tail, partial last tile, D % 64 != 0): all threads gather 16-byte
//     pieces with cp.async (LDGSTS) arriving on the same mbarrier, and store y rows with 128-bit STG.
// One __syncthreads per tile publishes the fp32-widened B|C rows and releases the oldest stage for refill; there
// is no producer warp and nothing spins.
//
// Three launch shapes share the kernel body (template MODE):
//   plain      (nseg <= 1, chain_chunks <= 1)  grid (D/64, B', 2*branches), one CTA per sequence;
//   chain      (chain_chunks > 1, the default for launches that fill the GPU)  every sequence is cut into
//              sequentially dependent chunks drawn from an atomic work counter — balances the 3-vs-4-warp schedulers
//              and the last wave; bit-identical results;
//   two-level  (nseg > 1, small batches / one long sequence)  chunk summaries (MODE 1) + scan_carry_kernel + rescan.
#include <cuda.h>
#include <string.h>

#include "scan_core.cuh"

namespace actk {

constexpr int kCh = 64;  // channels per CTA == threads per CTA
constexpr int kT = 16;   // time steps per tile
constexpr int kGroup = 4;  // steps software-pipelined together (ChannelScan::run)
#ifndef ACTK_STAGES16
#define ACTK_STAGES16 4
#endif
#ifndef ACTK_STAGES_FUSED
#define ACTK_STAGES_FUSED 3
#endif
// KS > 0: dt_proj is computed in the kernel (16-bit I/O only) from KS 16-wide slabs of the x_dbl dt columns.
template <typename T, int KS>
constexpr int ring_stages() { return sizeof(T) == 4 ? 3 : (KS > 0 ? ACTK_STAGES_FUSED : ACTK_STAGES16); }

template <typename T>
struct BranchDev {
  const T *xz, *tail, *xdbl, *xdbl_tail, *delta, *delta_tail, *w_dt;
  const int *idx;
  const float *A, *Dskip, *dt_bias;
  T *ydir;
  int n_sel, n_tail;
  int idx_iota;  // idx[p] == p for all p (n_sel == L): no index loads needed
};
template <typename T>
struct MaskedParams {
  BranchDev<T> br[2];
  int first_branch;
  int Bp, L, D, xw;
  int tma_ok;  // D >= 64: boxes are 64 channels wide, a partial last block relies on TMA out-of-bounds handling
  // two-level scan (nseg > 1): the sequence is cut into nseg chunks of whole tiles, scanned by different CTAs
  int nseg;
  float *ws_hend;   // (Bp, 2 branches, 2 dirs, nseg, D, 16) chunk-local end state (zero initial state)
  float *ws_sumdt;  // (Bp, 2, 2, nseg, D)                   sum of dt over the chunk
  float *ws_h0;     // (Bp, 2, 2, nseg, D, 16)               carried-in state of every chunk (written by scan_carry)
  // chain mode (MODE 2): nseg chunks of a sequence run one after another on whichever CTA slot frees up first
  int nq, nblk;       // sequences (= CTAs of a single-level launch) and channel blocks per (batch, item)
  int *chain_ctr;     // [1]   work counter, zeroed before launch
  int *chain_flag;    // [nq]  number of finished chunks of sequence q, zeroed before launch
  float *chain_state; // [nq][64][16] state handed from chunk c to chunk c+1
};
struct alignas(64) BranchMaps {
  CUtensorMap xz, xdbl, delta, ydir, xdbl_dt;
};
struct alignas(64) MaskedMaps {
  BranchMaps m[2];
};

// One ring slot.  Unfused (KS == 0): the dt_proj output tile arrives from HBM.  Fused: the tile of dt_proj INPUT
// columns (16*KS wide, zero-padded rank) arrives instead, already in the tensor cores' K-major core-matrix order
// [k-chunk of 8][token row][8 elements] (a 4-D tensor map writes that order directly), and the CTA produces the
// 64 x 16 delta tile itself with tcgen05.mma into tensor memory (see the kernel).
template <typename T, int KS>
struct alignas(128) Stage {
  T u[kT][kCh];
  T bc[kT][2 * kN];
  T dtin[2 * KS][kT][8];
};
template <typename T>
struct alignas(128) Stage<T, 0> {
  T u[kT][kCh];
  T dt[kT][kCh];
  T bc[kT][2 * kN];
};

// ---- tcgen05 / tensor-memory helpers for the fused dt_proj (SASS: UTCHMMA, LDTM) --------------------------------
// delta[ch][tok] = sum_r W[ch][r] * dtin[tok][r]  as D(128 x 16, fp32 in TMEM) = A(128 x K, smem) * B(16 x K, smem)^T,
// (rows 64-127 of A are whatever follows the 64 weight rows: their results land in TMEM lanes this CTA never reads)
// both operands K-major without swizzle: 8-row x 16-byte core matrices, contiguous 128 B each; LBO = byte distance
// between the two k-chunks of one K=16 instruction, SBO = distance between 8-row groups.
// The descriptor's low word holds (address >> 4) and LBO, the high word SBO and the version: moving the operand by
// `bytes` adds bytes >> 4 to the low word.
__device__ __forceinline__ uint32_t umma_desc_lo(const void *smem, uint32_t lbo_bytes) {
  return ((smem_u32(smem) & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo, uint32_t sbo_bytes) {
  return (uint64_t)lo | ((uint64_t)((sbo_bytes >> 4) | (1u << 14)) << 32);   // version 1 (sm_100), no swizzle
}
// instruction descriptor: fp32 accumulate, A/B both `fmt` (0 = f16, 1 = bf16), both K-major, M = 128, N = 16
__device__ __forceinline__ constexpr uint32_t umma_idesc_m128n16(uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
}
// Both are PREDICATED inside the asm instead of sitting in an `if (tid == 0)`: a thread-dependent branch (or a call)
// in the tile loop made the compiler move the scan loop's address arithmetic from the uniform datapath into vector
// registers (+6 instructions per step in an issue-bound loop: 1.58 -> 1.72 ms at config 2).  `issue` is non-zero in
// exactly one thread of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t issue, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t issue, uint64_t *bar) {   // arrives when all prior MMAs of the thread are done
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc32(uint32_t *slot_smem) {   // one full warp; 32 columns
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot_smem)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc32(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr) : "memory");
}
// thread i of the warp reads 4 (or 1) consecutive 32-bit columns of TMEM lane (lane field of taddr) + i.
// The load is asynchronous: tmem_ld4_issue starts it, tmem_ld4_wait makes the registers valid (and ties them to the
// wait through in/out operands so no use can be scheduled ahead of it).
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4_wait(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r0;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0) :: "memory");
  return __uint_as_float(r0);
}

struct TileGeo {
  int nrows;    // valid rows (16 except for the last tile)
  int l_lo;     // lowest sequence position of the tile
  int l_first;  // sequence position held by smem row 0 (can be negative for direction 1's last tile)
  int row0;     // FAST only: latent row of smem row 0
  bool fast;
};

// MODE 0: scan with outputs (chunk c > 0 starts from ws_h0).  MODE 1: chunk summary — state only, no C, no y;
// writes the chunk-local end state and sum(dt) for scan_carry_kernel.
// MODE 2: chained chunks.  2000 warp-sequences on 592 warp schedulers cannot be balanced statically (3 or 4 warps
// per scheduler; the 4-warp ones set the pace of a single-level launch).  Here every sequence is cut into nseg
// sequentially dependent chunks; a 1-D grid of nq*nseg CTAs draws (chunk, sequence) work items from an atomic
// counter in chunk-major order, waits (acquire) until the previous chunk of its sequence has published its state,
// scans, and publishes (release).  Chunks of one sequence land on different SMs, so every sequence advances at the
// average rate and fast SMs simply take more items.  A waiting CTA only ever waits for a CTA that drew a smaller
// ticket, i.e. one that is already running: no deadlock whatever the hardware's dispatch order.
template <typename T, bool POWER_A, int MODE, int KS>
__global__ void __launch_bounds__(kCh) masked_scan_kernel(const __grid_constant__ MaskedParams<T> P,
                                                          const __grid_constant__ MaskedMaps M) {
  constexpr int S = ring_stages<T, KS>();
  constexpr bool k16 = sizeof(T) == 2;
  constexpr bool kFused = KS > 0;
  // the single-thread TMA work (tile loads, y stores) runs in warp 1 when warp 0 issues the MMAs: both are serial
  // instruction chains on the tile's critical path, so they go side by side
  constexpr int kTmaTid = kFused ? 32 : 0;
  constexpr int RP = 16 * KS;                     // padded dt rank
  static_assert(!kFused || k16, "the fused dt_proj uses 16-bit tensor-core operands");
  __shared__ Stage<T, KS> st[S];
  __shared__ alignas(128) T ybuf[2][kT][kCh];
  // fused: dt_projs_weight rows of this CTA's channels as the MMA's A operand, [k-chunk][channel][8] — the image
  // actk_pack_dt_proj_weight prepares, fetched with one bulk copy.  64 rows of slack: the M=128 instruction's unused
  // rows 64-127 of the last k-chunk read them (results never looked at).
  constexpr uint32_t kWBytes = 2 * KS * kCh * 16;
  __shared__ alignas(128) T wsm[kFused ? 2 * KS * kCh * 8 + kCh * 8 : 8];
  __shared__ alignas(8) uint64_t mma_bar[2];
  __shared__ alignas(8) uint64_t w_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ alignas(16) float bcf[k16 ? 2 : 1][k16 ? kT : 1][2 * kN];  // fp32 view of B|C for 16-bit I/O
  __shared__ alignas(8) uint64_t full_bar[S];

  const int tid = threadIdx.x;
  const int nseg = P.nseg;
  int bx, b, zi, seg, q = 0;
  if (MODE == 2) {
    __shared__ int ticket;
    if (tid == 0) ticket = atomicAdd(P.chain_ctr, 1);
    __syncthreads();
    seg = ticket / P.nq;
    q = ticket - seg * P.nq;
    bx = q % P.nblk;
    b = (q / P.nblk) % P.Bp;
    zi = q / (P.nblk * P.Bp);
  } else {
    bx = blockIdx.x; b = blockIdx.y;
    seg = blockIdx.z % nseg;
    zi = blockIdx.z / nseg;
  }
  const int d0 = bx * kCh;
  const int bi = P.first_branch + (zi >> 1);
  const int k = zi & 1;
  const BranchDev<T> br = P.br[bi];
  const BranchMaps &maps = M.m[bi];
  const int n_sel = br.n_sel, n_tail = br.n_tail;
  const int Lp = n_sel + n_tail;
  const int D = P.D, L = P.L;
  const int nch = min(kCh, D - d0);
  const int ntiles = (Lp + kT - 1) / kT;
  // chunk `seg` owns tiles [t_begin, t_end) in processing order; an empty chunk still reports a zero summary
  const int seg_tiles = (ntiles + nseg - 1) / nseg;
  const int t_begin = min(seg * seg_tiles, ntiles), t_end = min(t_begin + seg_tiles, ntiles);
  const size_t ws_row = ((((size_t)b * 2 + bi) * 2 + k) * nseg + seg) * D + d0;   // + channel

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full_bar[s], kCh);
    mbar_fence_init();
    if (P.tma_ok) {
      tmap_prefetch(&maps.xz); tmap_prefetch(&maps.xdbl); tmap_prefetch(&maps.ydir);
      tmap_prefetch(kFused ? &maps.xdbl_dt : &maps.delta);
    }
  }
  if constexpr (kFused) {
    if (tid == 0) {
      mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1); mbar_init(&w_bar, 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(&w_bar, kWBytes);   // only the MMA-issuing thread ever waits for the weights
      bulk_g2s(wsm, br.w_dt + ((size_t)k * P.nblk + bx) * (kWBytes / sizeof(T)), kWBytes, &w_bar);
    }
    __syncwarp();
    if (tid < 32) tmem_alloc32(&tmem_slot);     // warp 0 owns the allocation (2 x 16 accumulator columns)
    tc_fence_before();
  }
  __syncthreads();
  uint32_t tmem = 0;
  if constexpr (kFused) { tc_fence_after(); tmem = tmem_slot; }

  auto geo = [&](int t) {
    TileGeo g;
    const int p0 = t * kT;
    g.nrows = min(kT, Lp - p0);
    int l_hi;
    if (k == 0) { g.l_lo = p0; l_hi = p0 + g.nrows - 1; g.l_first = p0; }
    else { l_hi = Lp - 1 - p0; g.l_lo = l_hi - g.nrows + 1; g.l_first = l_hi - (kT - 1); }
    g.fast = false; g.row0 = 0;
    if (P.tma_ok && g.nrows == kT && l_hi < n_sel) {
      if (br.idx_iota) { g.fast = true; g.row0 = g.l_lo; }
      else {
        const int r_lo = __ldg(br.idx + g.l_lo), r_hi = __ldg(br.idx + l_hi);
        g.fast = (r_hi - r_lo) == kT - 1;
        g.row0 = r_lo;
      }
    }
    return g;
  };

  // source pointers of sequence position l (ragged path)
  auto src_rows = [&](int l, const T *&usrc, const T *&bsrc) {
    if (l < n_sel) {
      const int row = br.idx_iota ? l : __ldg(br.idx + l);
      const size_t tok = (size_t)b * L + row;
      usrc = br.xz + tok * D + d0;
      bsrc = br.xdbl + ((size_t)b * n_sel + l) * P.xw;     // x_dbl of the selected tokens is in sequence order
    } else {
      const size_t tok = (size_t)b * n_tail + (l - n_sel);
      usrc = br.tail + tok * D + d0;
      bsrc = br.xdbl_tail + tok * P.xw;
    }
  };

  uint32_t fastmask = 0;   // fused: bit (tile - t_begin) & 31 = that tile was written by TMA (no proxy fence needed)
  auto issue_load = [&](int t, const TileGeo &g) {
    Stage<T, KS> &sg = st[(t - t_begin) % S];
    uint64_t *bar = &full_bar[(t - t_begin) % S];
    if constexpr (kFused) {
      const uint32_t bit = 1u << ((t - t_begin) & 31);
      fastmask = g.fast ? (fastmask | bit) : (fastmask & ~bit);
    }
    if (g.fast) {
      if (tid == kTmaTid) {
        mbar_expect_tx(bar, (uint32_t)sizeof(Stage<T, KS>));
        if constexpr (kFused) tma_load_4d(&sg.dtin[0][0][0], &maps.xdbl_dt, 0, g.l_first, (4 * kN + k * RP) / 8, b, bar);
        else tma_load_3d(&sg.dt[0][0], &maps.delta, k * D + d0, g.l_first, b, bar);
        tma_load_3d(&sg.u[0][0], &maps.xz, d0, g.row0, b, bar);
        tma_load_3d(&sg.bc[0][0], &maps.xdbl, k * 2 * kN, g.l_first, b, bar);
      }
      mbar_arrive(bar);
    } else {
      constexpr int kPer = 16 / sizeof(T);                 // elements per 16-byte piece
      const int cu = nch / kPer, cb = 2 * kN / kPer;       // pieces per u row, per B|C row
      const int cd = kFused ? RP / kPer : cu;              // pieces per dt-input row (fused) / delta row
      const int per_row = cu + cd + cb;
      if (nch == kCh) {
        // full channel block: four threads per tile row, fixed pieces per thread — one index lookup per thread and
        // no divisions (mask edges make a third to two thirds of the tiles ragged under rectangle masks)
        constexpr int CU = kCh / kPer, CB = 2 * kN / kPer, CD = kFused ? RP / kPer : CU;
        const int jj = tid >> 2, q = tid & 3;
        if (jj < g.nrows) {
          const int l = g.l_lo + jj, j = l - g.l_first;
          const T *usrc, *bsrc;
          src_rows(l, usrc, bsrc);
#pragma unroll
          for (int i = 0; i < CU / 4; ++i) cp_async16(&sg.u[j][(q + 4 * i) * kPer], usrc + (q + 4 * i) * kPer);
          if constexpr (kFused) {
#pragma unroll
            for (int i = 0; i < (CD + 3) / 4; ++i)
              if (q + 4 * i < CD) cp_async16(&sg.dtin[q + 4 * i][j][0], bsrc + 4 * kN + k * RP + (q + 4 * i) * kPer);
          } else {
            const T *dsrc = (l < n_sel ? br.delta + (((size_t)b * n_sel + l) * 2 + k) * D
                                       : br.delta_tail + (((size_t)b * n_tail + (l - n_sel)) * 2 + k) * D) + d0;
#pragma unroll
            for (int i = 0; i < CU / 4; ++i) cp_async16(&sg.dt[j][(q + 4 * i) * kPer], dsrc + (q + 4 * i) * kPer);
          }
#pragma unroll
          for (int i = 0; i < CB / 4; ++i) cp_async16(&sg.bc[j][(q + 4 * i) * kPer], bsrc + k * 2 * kN + (q + 4 * i) * kPer);
        }
      } else
      for (int id = tid; id < g.nrows * per_row; id += kCh) {
        const int jj = id / per_row, w = id - jj * per_row;
        const int l = g.l_lo + jj, j = l - g.l_first;
        if (w < cu) {
          const T *usrc, *bsrc;
          src_rows(l, usrc, bsrc);
          cp_async16(&sg.u[j][w * kPer], usrc + w * kPer);
        } else if (w < cu + cd) {
          if constexpr (kFused) {
            const T *usrc, *bsrc;
            src_rows(l, usrc, bsrc);
            cp_async16(&sg.dtin[w - cu][j][0], bsrc + 4 * kN + k * RP + (w - cu) * kPer);
          } else {
            const T *dsrc = (l < n_sel ? br.delta + (((size_t)b * n_sel + l) * 2 + k) * D
                                       : br.delta_tail + (((size_t)b * n_tail + (l - n_sel)) * 2 + k) * D) + d0;
            cp_async16(&sg.dt[j][(w - cu) * kPer], dsrc + (w - cu) * kPer);
          }
        } else {
          const T *usrc, *bsrc;
          src_rows(l, usrc, bsrc);
          cp_async16(&sg.bc[j][(w - cu - cd) * kPer], bsrc + k * 2 * kN + (w - cu - cd) * kPer);
        }
      }
      cp_async_arrive_noinc(bar);
    }
  };

  // Fused dt_proj of tile t on the tensor cores (SASS UTCHMMA): D (128 x 16 tile rows, fp32, TMEM columns
  // [16*(tr&1), +16)) = W (rows 0-63 = this CTA's channels) * dtin(16 x RP)^T, one M=128 instruction per 16 ranks.
  // Accumulator row r lives in TMEM lane r, so thread tid later reads its own channel with tcgen05.ld (warp w can
  // reach lanes 32w .. 32w+31); lanes 64-127 hold the unused rows.  Called by ALL lanes of warp 0 (converged): the
  // address arithmetic stays on the uniform datapath and one elected lane issues — the MMA warp is on the tile's
  // critical path (the scan is latency-bound per sequence), so every instruction here counts.
  // Executed by EVERY thread (no thread-dependent branch, see umma_f16); thread 0 is the one that issues.
  const uint32_t mma_issuer = tid == 0 ? 1u : 0u;
  auto issue_mma = [&](int t) {
    if constexpr (kFused) {
      const int tr1 = t - t_begin, s1 = tr1 % S;
      if (tr1 == 0) mbar_wait(&w_bar, 0);        // weights landed (first tile of this CTA only)
      mbar_wait(&full_bar[s1], (tr1 / S) & 1);   // the tile's dt columns have landed
      if (!((fastmask >> (tr1 & 31)) & 1)) fence_proxy_async();   // ragged tiles are written by cp.async (generic proxy)
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_m128n16(IO<T>::is_bf16 ? 1u : 0u);
      const uint32_t d = tmem + (uint32_t)(tr1 & 1) * kT;
      const uint32_t alo = umma_desc_lo(wsm, kCh * 16), blo = umma_desc_lo(&st[s1].dtin[0][0][0], kT * 16);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        umma_f16(mma_issuer, d, umma_desc(alo + ks * (2 * kCh * 16 >> 4), 128), umma_desc(blo + ks * (2 * kT * 16 >> 4), 128),
                 idesc, ks > 0);
      umma_commit(mma_issuer, &mma_bar[tr1 & 1]);
    }
  };

  T *ydst = br.ydir + ((size_t)k * P.Bp + b) * L * D + d0;
  auto store_y = [&](int t, const TileGeo &g) {
    if (g.fast) {
      if (tid == kTmaTid) {
        tma_store_3d(&maps.ydir, d0, g.row0, k * P.Bp + b, &ybuf[(t - t_begin) & 1][0][0]);
        bulk_commit();
      }
    } else {
      constexpr int kPer = 16 / sizeof(T);
      const int cu = nch / kPer;
      if (nch == kCh) {   // four threads per tile row, fixed pieces (see issue_load)
        constexpr int CU = kCh / kPer;
        const int jj = tid >> 2, q = tid & 3, l = g.l_lo + jj;
        if (jj < g.nrows && l < n_sel) {
          const int row = br.idx_iota ? l : __ldg(br.idx + l);
#pragma unroll
          for (int i = 0; i < CU / 4; ++i) {
            const uint4 v = *reinterpret_cast<const uint4 *>(&ybuf[(t - t_begin) & 1][l - g.l_first][(q + 4 * i) * kPer]);
            *reinterpret_cast<uint4 *>(ydst + (size_t)row * D + (q + 4 * i) * kPer) = v;
          }
        }
      } else
      for (int id = tid; id < g.nrows * cu; id += kCh) {
        const int jj = id / cu, w = id - jj * cu;
        const int l = g.l_lo + jj;
        if (l < n_sel) {
          const int row = br.idx_iota ? l : __ldg(br.idx + l);
          const uint4 v = *reinterpret_cast<const uint4 *>(&ybuf[(t - t_begin) & 1][l - g.l_first][w * kPer]);
          *reinterpret_cast<uint4 *>(ydst + (size_t)row * D + w * kPer) = v;
        }
      }
    }
  };

  const bool live = tid < nch;
  const int ch = k * D + d0 + (live ? tid : 0);
  ChannelScan<POWER_A> cs;
  cs.init(br.A + (size_t)ch * kN, br.Dskip[ch], br.dt_bias[ch]);
  // ring slots and y double-buffer are indexed by the tile number relative to the chunk start.  The first tiles are
  // requested BEFORE a chained chunk waits for its predecessor's state: their HBM latency overlaps the wait.
  for (int t = t_begin; t < min(t_begin + S - 1, t_end); ++t) issue_load(t, geo(t));

  if (MODE == 0 && seg > 0 && live) {
    const float *h0 = P.ws_h0 + (ws_row + tid) * kN;
#pragma unroll
    for (int j = 0; j < kN / 2; ++j) cs.h[j] = pk(h0[2 * j], h0[2 * j + 1]);
  }
  if (MODE == 2 && seg > 0) {
    if (tid == 0) {   // acquire: the previous chunk of this sequence has published its state
      int done;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(P.chain_flag + q) : "memory");
        if (done < seg) __nanosleep(200);
      } while (done < seg);
    }
    __syncthreads();
    const float4 *h0 = reinterpret_cast<const float4 *>(P.chain_state + ((size_t)q * kCh + tid) * kN);
#pragma unroll
    for (int j = 0; j < kN / 4; ++j) {
      const float4 v = __ldcg(h0 + j);
      cs.h[2 * j] = pk(v.x, v.y);
      cs.h[2 * j + 1] = pk(v.z, v.w);
    }
  }
  float sumdt = 0.f;


  if constexpr (kFused) {
    if (t_begin < t_end) issue_mma(t_begin);
  }

  TileGeo prev = {};
  for (int t = t_begin; t < t_end; ++t) {
    const int tr = t - t_begin;
    const int s = tr % S;
    const TileGeo g = geo(t);
    mbar_wait(&full_bar[s], (tr / S) & 1);
    if (k16) {  // widen this tile's B|C rows to fp32 once per CTA: thread -> (row tid/4, 8 values)
      const int j = tid >> 2, q = (tid & 3) * 8;
      uint4 w = *reinterpret_cast<const uint4 *>(&st[s].bc[j][q]);
      const T *e = reinterpret_cast<const T *>(&w);
      float4 lo = make_float4(IO<T>::ld(e + 0), IO<T>::ld(e + 1), IO<T>::ld(e + 2), IO<T>::ld(e + 3));
      float4 hi = make_float4(IO<T>::ld(e + 4), IO<T>::ld(e + 5), IO<T>::ld(e + 6), IO<T>::ld(e + 7));
      float4 *dst = reinterpret_cast<float4 *>(&bcf[tr & 1][j][q]);
      dst[0] = lo;
      dst[1] = hi;
    }
    if (tid == kTmaTid) bulk_wait_read<0>();  // the y tile stored two iterations ago has left ybuf[t & 1]
    if constexpr (kFused) tc_fence_before();   // this thread's TMEM reads of tile t-1 precede the barrier
    __syncthreads();                    // B|C published; everyone is done with tile t-1 (its stage, y tile, TMEM buffer)
    if (MODE != 1 && tr > 0) store_y(t - 1, prev);
    if (t + S - 1 < t_end) issue_load(t + S - 1, geo(t + S - 1));
    uint32_t tacc = 0;                  // TMEM address of this thread's delta row: lane = channel, column = tile row
    if constexpr (kFused) {
      if (t + 1 < t_end) issue_mma(t + 1);                      // tile t+1's delta forms while tile t is scanned
      mbar_wait(&mma_bar[tr & 1], (tr >> 1) & 1);               // tile t's delta is in TMEM
      tc_fence_after();
      tacc = tmem + ((uint32_t)(tid & 32) << 16) + (uint32_t)(tr & 1) * kT;
    }

    // Fused: tcgen05.ld is warp-collective, so lanes of a partial channel block run along (their u columns are
    // zero-filled / never stored; their y columns are clipped by the stores).
    if (live || kFused) {
      const T *us = &st[s].u[0][tid];
      const T *ds = nullptr;
      if constexpr (!kFused) ds = &st[s].dt[0][tid];
      const float *bcs = k16 ? &bcf[tr & 1][0][0] : reinterpret_cast<const float *>(&st[s].bc[0][0]);
      T *ys = &ybuf[tr & 1][0][tid];
      // raw delta of tile row j for this thread's channel, rounded to T where the reference holds the dts tensor
      auto delta_row = [&](int j) -> float {
        if constexpr (kFused) return IO<T>::rnd(tmem_ld1(tacc + j));
        else return IO<T>::ld(ds + j * kCh);
      };
      if (MODE == 1) {
        int r = 0;
        if (g.nrows == kT) {
          uint32_t dv[kGroup] = {}, dn[kGroup] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc + (k ? kT - kGroup : 0), dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (; r < kT; r += kGroup) {
            const int j0 = k ? kT - 1 - r : r, dj = k ? -1 : 1;
            if constexpr (kFused) {   // next group's delta values travel from TMEM while this group is scanned
              if (r + kGroup < kT) tmem_ld4_issue(tacc + (k ? kT - 2 * kGroup - r : r + kGroup), dn);
            }
            sumdt += cs.template run_state<kGroup, true>(
                [&](int i) { return IO<T>::ld(us + (j0 + dj * i) * kCh); },
                [&](int i) {
                  if constexpr (kFused) return IO<T>::rnd(__uint_as_float(k ? dv[kGroup - 1 - i] : dv[i]));
                  else return IO<T>::ld(ds + (j0 + dj * i) * kCh);
                },
                [&](int i) { return bcs + (j0 + dj * i) * 2 * kN; });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kGroup; ++i) dv[i] = dn[i];
            }
          }
        }
        for (; r < g.nrows; ++r) {
          const int j = k ? kT - 1 - r : r;
          const StepIn si = cs.template prologue<true>(IO<T>::ld(us + j * kCh), delta_row(j));
          uint64_t p[kN / 2];
          cs.decay(si.dt, p);
          cs.apply_state(p, si, bcs + j * 2 * kN);
          sumdt += si.dt;
        }
      } else if (g.nrows == kT) {
        // smem row of step r: r (direction 0) or 15 - r (direction 1); kGroup steps are software-pipelined
        if (k == 0) {
          uint32_t dv[kGroup] = {}, dn[kGroup] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc, dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (int r0 = 0; r0 < kT; r0 += kGroup) {
            const T *u0 = us + r0 * kCh, *dl0 = ds + r0 * kCh;
            const float *b0 = bcs + r0 * 2 * kN;
            T *y0 = ys + r0 * kCh;
            if constexpr (kFused) {   // next group's delta values travel from TMEM while this group is scanned
              if (r0 + kGroup < kT) tmem_ld4_issue(tacc + r0 + kGroup, dn);
            }
            cs.template run<kGroup, true>([&](int i) { return IO<T>::ld(u0 + i * kCh); },
                                          [&](int i) {
                                            if constexpr (kFused) return IO<T>::rnd(__uint_as_float(dv[i]));
                                            else return IO<T>::ld(dl0 + i * kCh);
                                          },
                                          [&](int i) { return b0 + i * 2 * kN; },
                                          [&](int i, float y) { IO<T>::st(y0 + i * kCh, y); });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kGroup; ++i) dv[i] = dn[i];
            }
          }
        } else {
          uint32_t dv[kGroup] = {}, dn[kGroup] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc + kT - kGroup, dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (int r0 = 0; r0 < kT; r0 += kGroup) {
            const int j0 = kT - 1 - r0;
            const T *u0 = us + j0 * kCh, *dl0 = ds + j0 * kCh;
            const float *b0 = bcs + j0 * 2 * kN;
            T *y0 = ys + j0 * kCh;
            if constexpr (kFused) {
              if (r0 + kGroup < kT) tmem_ld4_issue(tacc + kT - 2 * kGroup - r0, dn);
            }
            cs.template run<kGroup, true>([&](int i) { return IO<T>::ld(u0 - i * kCh); },
                                          [&](int i) {
                                            if constexpr (kFused) return IO<T>::rnd(__uint_as_float(dv[kGroup - 1 - i]));
                                            else return IO<T>::ld(dl0 - i * kCh);
                                          },
                                          [&](int i) { return b0 - i * 2 * kN; },
                                          [&](int i, float y) { IO<T>::st(y0 - i * kCh, y); });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kGroup; ++i) dv[i] = dn[i];
            }
          }
        }
      } else {
        for (int r = 0; r < g.nrows; ++r) {
          const int j = k ? kT - 1 - r : r;
          float y = cs.template step<true>(IO<T>::ld(us + j * kCh), delta_row(j), bcs + j * 2 * kN);
          IO<T>::st(ys + j * kCh, y);
        }
      }
    }
    fence_proxy_async();  // make this thread's ybuf writes visible to the TMA store issued after the next barrier
    prev = g;
  }
  if (MODE == 1) {
    if (live) {
      float *he = P.ws_hend + (ws_row + tid) * kN;
#pragma unroll
      for (int j = 0; j < kN / 2; ++j) upk(cs.h[j], he[2 * j], he[2 * j + 1]);
      P.ws_sumdt[ws_row + tid] = sumdt;
    }
    if constexpr (kFused) {
      tc_fence_before();
      __syncthreads();
      if (tid < 32) tmem_dealloc32(tmem);
    }
    return;
  }
  if (MODE == 2 && seg + 1 < nseg) {   // publish the state for the next chunk of this sequence (release)
    float4 *hs = reinterpret_cast<float4 *>(P.chain_state + ((size_t)q * kCh + tid) * kN);
#pragma unroll
    for (int j = 0; j < kN / 4; ++j) {
      float4 v;
      upk(cs.h[2 * j], v.x, v.y);
      upk(cs.h[2 * j + 1], v.z, v.w);
      __stcg(hs + j, v);
    }
    __threadfence();
  }
  if (tid == kTmaTid) bulk_wait_read<0>();
  if constexpr (kFused) tc_fence_before();
  __syncthreads();
  if constexpr (kFused) {
    if (tid < 32) tmem_dealloc32(tmem);   // every thread's TMEM reads are behind the barrier
  }
  if (MODE == 2 && seg + 1 < nseg && tid == 0)
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(P.chain_flag + q), "r"(seg + 1) : "memory");
  if (t_end > t_begin) store_y(t_end - 1, prev);
  if (tid == kTmaTid) bulk_wait_read<0>();  // shared memory must outlive the last TMA store's reads
}

// Inter-chunk carry of the two-level scan: one thread per (batch, branch, direction, channel) walks the nseg
// chunk summaries in processing order:  h0[c+1] = exp(A * sumdt[c]) * h0[c] + hend[c],  h0[0] = 0.
__global__ void __launch_bounds__(128) scan_carry_kernel(const float *__restrict__ A0, const float *__restrict__ A1,
                                                         const float *__restrict__ hend, const float *__restrict__ sumdt,
                                                         float *__restrict__ h0, int Bp, int D, int nseg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // ((b*2 + bi)*2 + k)*D + d
  if (i >= (long long)Bp * 4 * D) return;
  const int d = (int)(i % D), k = (int)((i / D) & 1), bi = (int)((i / (2LL * D)) & 1);
  const float *A = (bi ? A1 : A0);
  if (!A) return;                                   // branch not live
  A += ((size_t)k * D + d) * kN;
  float a[kN], h[kN];
#pragma unroll
  for (int n = 0; n < kN; ++n) { a[n] = A[n] * kLog2e; h[n] = 0.f; }
  const long long base = (i / D) * (long long)nseg * D + d;              // chunk c sits at base + c*D
  for (int c = 0; c < nseg; ++c) {
    const size_t row = (size_t)(base + (long long)c * D);
#pragma unroll
    for (int n = 0; n < kN; ++n) h0[row * kN + n] = h[n];
    const float sd = sumdt[row];
#pragma unroll
    for (int n = 0; n < kN; ++n) h[n] = fmaf(ex2(a[n] * sd), h[n], hend[row * kN + n]);
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (inner, rows, outer) tensor with dense rows of `inner` elements; box = (box_inner, kT, 1)
static int make_map(CUtensorMap *m, CUtensorMapDataType dt, int es, const void *base, uint64_t inner, uint64_t rows,
                    uint64_t outer, uint32_t box_inner) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available from this driver");
  cuuint64_t dims[3] = {inner, rows, outer};
  cuuint64_t strides[2] = {inner * es, inner * rows * es};
  cuuint32_t box[3] = {box_inner, (cuuint32_t)kT, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ACTK_OK;
}

// One kernel launch for a given (T, KS); MODE and POWER_A are runtime here.
template <typename T, int KS>
static void launch_ks(bool pw, int mode, dim3 grid, cudaStream_t stream, const MaskedParams<T> &P, const MaskedMaps &M) {
  if (mode == 2) {
    if (pw) masked_scan_kernel<T, true, 2, KS><<<grid, kCh, 0, stream>>>(P, M);
    else masked_scan_kernel<T, false, 2, KS><<<grid, kCh, 0, stream>>>(P, M);
  } else if (mode == 1) {
    if (pw) masked_scan_kernel<T, true, 1, KS><<<grid, kCh, 0, stream>>>(P, M);
    else masked_scan_kernel<T, false, 1, KS><<<grid, kCh, 0, stream>>>(P, M);
  } else {
    if (pw) masked_scan_kernel<T, true, 0, KS><<<grid, kCh, 0, stream>>>(P, M);
    else masked_scan_kernel<T, false, 0, KS><<<grid, kCh, 0, stream>>>(P, M);
  }
}
template <typename T>
static void launch_any(int ks, bool pw, int mode, dim3 grid, cudaStream_t stream, const MaskedParams<T> &P,
                       const MaskedMaps &M) {
  if constexpr (sizeof(T) == 2) {
    switch (ks) {
      case 2: return launch_ks<T, 2>(pw, mode, grid, stream, P, M);
      case 3: return launch_ks<T, 3>(pw, mode, grid, stream, P, M);
      case 5: return launch_ks<T, 5>(pw, mode, grid, stream, P, M);
      default: break;
    }
  }
  launch_ks<T, 0>(pw, mode, grid, stream, P, M);
}

// dt-input columns of x_dbl as (8 elements, rows, 16-byte chunks of a row, outer): a box of (8, kT, rp/8, 1) lands
// in shared memory as [chunk][row][8] — the K-major core-matrix order tcgen05.mma reads without swizzle.
static int make_map_dtin(CUtensorMap *m, CUtensorMapDataType dt, int es, const void *base, uint64_t xw, uint64_t rows,
                         uint64_t outer, uint32_t rp) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available from this driver");
  cuuint64_t dims[4] = {8, rows, xw / 8, outer};
  cuuint64_t strides[3] = {xw * es, 16, xw * rows * es};
  cuuint32_t box[4] = {8, (cuuint32_t)kT, rp / 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled (dt columns) failed with CUresult %d", (int)r);
  return ACTK_OK;
}

template <typename T>
static int launch_masked(const actk_masked_scan_args *a, cudaStream_t stream) {
  MaskedParams<T> P;
  MaskedMaps M;
  memset(&M, 0, sizeof(M));
  P.Bp = a->Bp; P.L = a->L; P.D = a->D; P.xw = a->xw;
  const int nseg = a->nseg > 1 ? a->nseg : 1;
  P.nseg = nseg;
  P.ws_hend = P.ws_sumdt = P.ws_h0 = nullptr;
  P.chain_ctr = P.chain_flag = nullptr; P.chain_state = nullptr; P.nq = 0;
  P.nblk = (a->D + kCh - 1) / kCh;
  const bool chain = a->chain_chunks > 1;
  if (chain && a->nseg > 1) ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: nseg and chain_chunks are mutually exclusive");
  if (nseg > 1) {
    const size_t rows = (size_t)a->Bp * 4 * nseg * a->D;
    if (!a->workspace || a->workspace_bytes < (long long)(rows * (2 * kN + 1) * sizeof(float)))
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: nseg=%d needs a workspace of %lld bytes (actk_masked_scan_workspace_bytes)",
                nseg, (long long)(rows * (2 * kN + 1) * sizeof(float)));
    P.ws_hend = static_cast<float *>(a->workspace);
    P.ws_h0 = P.ws_hend + rows * kN;
    P.ws_sumdt = P.ws_h0 + rows * kN;
  }
  // boxes are always 64 channels wide; a last partial channel block relies on TMA's out-of-bounds handling
  // (zero fill on load, clipping on store), so any D >= 64 qualifies (channel-sharded slices such as D/8 = 80)
  P.tma_ok = (a->D >= kCh) ? 1 : 0;
  const int es = sizeof(T);
  const int ks = a->dt_rank_pad / 16;   // 0: delta tensors given; 2/3/5: dt_proj fused into the scan
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                         : (a->dtype == ACTK_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  for (int i = 0; i < 2; ++i) {
    const actk_branch_args &s = a->br[i < a->n_branches ? i : 0];
    BranchDev<T> &d = P.br[i];
    d.xz = (const T *)s.xz; d.tail = (const T *)s.tail; d.xdbl = (const T *)s.xdbl;
    d.xdbl_tail = (const T *)s.xdbl_tail; d.delta = (const T *)s.delta; d.delta_tail = (const T *)s.delta_tail;
    d.w_dt = (const T *)s.w_dt;
    d.idx = s.idx;
    d.A = s.A; d.Dskip = s.Dskip; d.dt_bias = s.dt_bias; d.ydir = (T *)s.ydir;
    d.n_sel = s.n_sel; d.n_tail = s.n_tail;
    d.idx_iota = s.n_sel == a->L;   // ascending distinct rows in [0, L): all L of them means idx[p] == p
    if (P.tma_ok && i < a->n_branches && s.n_sel > 0) {
      const uint64_t Lp = (uint64_t)s.n_sel + s.n_tail;
      int rc;
      if ((rc = make_map(&M.m[i].xz, dt, es, s.xz, a->D, a->L, a->Bp, kCh))) return rc;
      if ((rc = make_map(&M.m[i].xdbl, dt, es, s.xdbl, a->xw, (uint64_t)s.n_sel, a->Bp, 2 * kN))) return rc;
      if (ks) {
        if ((rc = make_map_dtin(&M.m[i].xdbl_dt, dt, es, s.xdbl, a->xw, (uint64_t)s.n_sel, a->Bp, 16 * ks))) return rc;
      } else if ((rc = make_map(&M.m[i].delta, dt, es, s.delta, 2ull * a->D, (uint64_t)s.n_sel, a->Bp, kCh))) return rc;
      if ((rc = make_map(&M.m[i].ydir, dt, es, s.ydir, a->D, a->L, 2ull * a->Bp, kCh))) return rc;
    }
  }
  // group consecutive live branches that share an A kind into one launch (better tail balance)
  int i = 0;
  while (i < a->n_branches) {
    if (a->br[i].n_sel == 0) { ++i; continue; }
    int j = i + 1;
    while (j < a->n_branches && a->br[j].n_sel > 0 && a->br[j].a_kind == a->br[i].a_kind) ++j;
    P.first_branch = i;
    dim3 grid((a->D + kCh - 1) / kCh, a->Bp, 2 * (j - i) * nseg);
    const bool pw = a->br[i].a_kind == ACTK_A_POWER;
    if (chain) {
      P.nblk = (a->D + kCh - 1) / kCh;
      P.nq = P.nblk * a->Bp * 2 * (j - i);
      P.nseg = a->chain_chunks;
      const size_t state_bytes = (size_t)P.nq * kCh * kN * sizeof(float);
      const size_t need = state_bytes + ((size_t)P.nq + 1) * sizeof(int);
      if (!a->workspace || (size_t)a->workspace_bytes < need)
        ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: chain_chunks=%d needs a workspace of %zu bytes", a->chain_chunks, need);
      P.chain_state = static_cast<float *>(a->workspace);
      P.chain_flag = reinterpret_cast<int *>(static_cast<char *>(a->workspace) + state_bytes);
      P.chain_ctr = P.chain_flag + P.nq;
      ACTK_CUDA_OK(cudaMemsetAsync(P.chain_flag, 0, ((size_t)P.nq + 1) * sizeof(int), stream));
      const unsigned nblocks = (unsigned)P.nq * a->chain_chunks;
      launch_any<T>(ks, pw, 2, dim3(nblocks), stream, P, M);
      ACTK_CUDA_OK(cudaGetLastError());
      i = j;
      continue;
    }
    if (nseg > 1) {   // level 1: chunk summaries; level 2: carries; then the scan proper starts every chunk from its carry
      launch_any<T>(ks, pw, 1, grid, stream, P, M);
      ACTK_CUDA_OK(cudaGetLastError());
      const float *A0 = (i == 0) ? a->br[0].A : nullptr;
      const float *A1 = (i <= 1 && j >= 2) ? a->br[1].A : nullptr;
      const long long n = (long long)a->Bp * 4 * a->D;
      scan_carry_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(A0, A1, P.ws_hend, P.ws_sumdt, P.ws_h0, a->Bp,
                                                                          a->D, nseg);
      ACTK_CUDA_OK(cudaGetLastError());
    }
    launch_any<T>(ks, pw, 0, grid, stream, P, M);
    ACTK_CUDA_OK(cudaGetLastError());
    i = j;
  }
  return ACTK_OK;
}

static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

}  // namespace actk

using namespace actk;

namespace actk {
// dt_projs_weight (2, D, R) -> per (direction, 64-channel block) shared-memory image of the fused dt_proj's A operand:
// [k-chunk of 8 ranks][channel of the block][8] (the K-major core-matrix order of masked_scan_kernel's issue_mma),
// rank zero-padded to rp, channels >= D zero.  One thread per 16-byte piece.
template <typename T>
__global__ void pack_dt_weight_kernel(const T *__restrict__ w, T *__restrict__ img, int D, int R, int rp, int nblk) {
  const int pieces = rp / 8;
  const long long n = 2LL * nblk * pieces * kCh;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = (int)(i % kCh), c = (int)((i / kCh) % pieces), bx = (int)((i / ((long long)kCh * pieces)) % nblk);
  const int k = (int)(i / ((long long)kCh * pieces * nblk));
  const int ch = bx * kCh + slot;
  T *dst = img + i * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = c * 8 + e;
    dst[e] = (ch < D && r < R) ? w[((size_t)k * D + ch) * R + r] : IO<T>::zero();
  }
}
}  // namespace actk

extern "C" long long actk_dt_proj_image_bytes(int D, int dt_rank_pad, int elsize) {
  return 2LL * ((D + kCh - 1) / kCh) * kCh * dt_rank_pad * elsize;
}

extern "C" int actk_pack_dt_proj_weight(const void *w, int D, int R, int dt_rank_pad, int dtype, void *img, void *stream) {
  if (!w || !img) ACTK_FAIL(ACTK_ERR_BAD_ARG, "pack_dt_proj_weight: NULL pointer");
  if (dtype != ACTK_F16 && dtype != ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "pack_dt_proj_weight: dtype=%d (f16 / bf16)", dtype);
  if (D <= 0 || R <= 0 || R > dt_rank_pad || (dt_rank_pad != 32 && dt_rank_pad != 48 && dt_rank_pad != 80))
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "pack_dt_proj_weight: D=%d R=%d dt_rank_pad=%d", D, R, dt_rank_pad);
  if (misaligned(img)) ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "pack_dt_proj_weight: image not aligned to 16 bytes");
  const int nblk = (D + kCh - 1) / kCh;
  const long long n = 2LL * nblk * (dt_rank_pad / 8) * kCh;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (dtype == ACTK_F16)
    pack_dt_weight_kernel<__half><<<blocks, 128, 0, st>>>((const __half *)w, (__half *)img, D, R, dt_rank_pad, nblk);
  else
    pack_dt_weight_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>((const __nv_bfloat16 *)w, (__nv_bfloat16 *)img, D, R,
                                                                   dt_rank_pad, nblk);
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

extern "C" long long actk_masked_scan_workspace_bytes(const actk_masked_scan_args *a) {
  if (!a) return 0;
  if (a->chain_chunks > 1) {   // state of every sequence-CTA + flags + counter (sized for both branches in one launch)
    const long long nq = (long long)((a->D + kCh - 1) / kCh) * a->Bp * 4;
    return nq * kCh * kN * (long long)sizeof(float) + (nq + 1) * (long long)sizeof(int);
  }
  if (a->nseg <= 1) return 0;
  return (long long)a->Bp * 4 * a->nseg * a->D * (2 * kN + 1) * (long long)sizeof(float);
}

extern "C" int actk_masked_scan_fwd(const actk_masked_scan_args *a, void *stream) {
  if (!a) ACTK_FAIL(ACTK_ERR_BAD_ARG, "actk_masked_scan_fwd: args is NULL");
  if (a->dtype < ACTK_F32 || a->dtype > ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "masked_scan: dtype=%d", a->dtype);
  if (a->n_branches < 1 || a->n_branches > 2) ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: n_branches=%d", a->n_branches);
  if (a->N != kN) ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "masked_scan: d_state=%d, this build has %d", a->N, kN);
  if (a->Bp <= 0 || a->L <= 0 || a->D <= 0 || a->xw < 4 * kN)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: Bp=%d L=%d D=%d xw=%d", a->Bp, a->L, a->D, a->xw);
  if (a->Bp > 32767) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: Bp=%d exceeds grid.y / 2", a->Bp);
  if (a->chain_chunks < 0 || a->chain_chunks > 1024)
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: chain_chunks=%d (0/1 = off, <= 1024)", a->chain_chunks);
  if (a->nseg < 0 || a->nseg > 4096) ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: nseg=%d (0/1 = single level, <= 4096)", a->nseg);
  const int es = a->dtype == ACTK_F32 ? 4 : 2;
  const int rp = a->dt_rank_pad;
  if (rp != 0) {
    if (es != 2) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "masked_scan: the fused dt_proj (dt_rank_pad=%d) needs f16 / bf16 I/O", rp);
    if (rp != 32 && rp != 48 && rp != 80)
      ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "masked_scan: dt_rank_pad=%d, this build has 32 / 48 / 80 (0 = delta tensors given)", rp);
    if (a->xw % 8 != 0 || a->xw < 4 * kN + 2 * rp)
      ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: xw=%d < 4N + 2*dt_rank_pad = %d", a->xw, 4 * kN + 2 * rp);
  }
  if ((a->D * es) % 16 != 0)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: D=%d must be a multiple of %d (16-byte channel rows)", a->D, 16 / es);
  if ((a->xw * es) % 16 != 0) ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "masked_scan: xdbl row pitch %d B not a multiple of 16", a->xw * es);
  for (int i = 0; i < a->n_branches; ++i) {
    const actk_branch_args &s = a->br[i];
    if (s.n_sel < 0 || s.n_sel > a->L || s.n_tail < 0)
      ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: branch %d n_sel=%d n_tail=%d L=%d", i, s.n_sel, s.n_tail, a->L);
    if (s.a_kind != ACTK_A_GENERAL && s.a_kind != ACTK_A_POWER)
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: branch %d a_kind=%d", i, s.a_kind);
    if (s.n_sel == 0) continue;
    if (!s.xz || !s.xdbl || !s.idx || !s.A || !s.Dskip || !s.dt_bias || !s.ydir ||
        (s.n_tail > 0 && (!s.tail || !s.xdbl_tail)) ||
        (rp == 0 && (!s.delta || (s.n_tail > 0 && !s.delta_tail))) || (rp != 0 && !s.w_dt))
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: branch %d has a NULL pointer", i);
    if (misaligned(s.xz) || misaligned(s.tail) || misaligned(s.xdbl) || misaligned(s.xdbl_tail) ||
        (rp == 0 && (misaligned(s.delta) || misaligned(s.delta_tail))) || (rp != 0 && misaligned(s.w_dt)) ||
        misaligned(s.ydir))
      ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "masked_scan: branch %d has a pointer not aligned to 16 bytes", i);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case ACTK_F32: return launch_masked<float>(a, st);
    case ACTK_F16: return launch_masked<__half>(a, st);
    default: return launch_masked<__nv_bfloat16>(a, st);
  }
}
