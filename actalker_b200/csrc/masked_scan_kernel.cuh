// Kernel of the fused masked bidirectional selective scan (C-ABI entry actk_masked_scan_fwd, masked_scan.cu).
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
#pragma once
#include "masked_scan_types.cuh"

namespace actk {

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
template <int NR>
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[NR]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <int NR>
__device__ __forceinline__ void tmem_ld4_wait(uint32_t (&r)[NR]) {
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
// SHORT_RING: a 3-slot ring for 16-bit I/O (23.6 KB of shared memory, 8 CTAs per SM instead of 7) — the host picks it
// when a launch has more sequences than 7 CTAs per SM can hold (CFG x2 / x4, the UNet's d_model 640 / 1280 layers):
// measured on B200, config-2 shape at B' = 50 / 100: 2.96 -> 2.90 ms, 5.88 -> 5.76 ms; at B' = 25 (1000 sequences, all
// resident either way) the 4-slot ring is 3.5 % faster and stays.
template <typename T, bool POWER_A, int MODE, int KS, bool SHORT_RING = false>
__global__ void __launch_bounds__(kCh) masked_scan_kernel(const __grid_constant__ MaskedParams<T> P,
                                                          const __grid_constant__ MaskedMaps M) {
  constexpr int S = SHORT_RING ? 3 : ring_stages<T, KS>();
  constexpr bool k16 = sizeof(T) == 2;
  constexpr bool kFused = KS > 0;
  // Steps software-pipelined together.  8 (two groups per tile) for 16-bit launches that leave a free register budget
  // at 7 CTAs per SM: all prologues (softplus chains) of 8 steps overlap and decay(i+1) runs ahead of apply(i) across
  // the whole group — measured on B200: config 2 1.508 -> 1.482 ms, a rank's d_inner slice 1.03 -> 0.98 ms (N=2),
  // 0.70 -> 0.66 ms (N=8), rectangle masks at B'=25 0.418 -> 0.400 ms.  With the 3-slot ring (8 CTAs per SM, CFG x4) 4
  // is 1 % faster and stays; fp32 I/O and the fused dt_proj (4 TMEM columns per load) keep 4.
#ifndef ACTK_CHAIN_POLL_NS
#define ACTK_CHAIN_POLL_NS 200
#endif
#ifndef ACTK_GROUP_WIDE
#define ACTK_GROUP_WIDE 8
#endif
  constexpr int kG = (sizeof(T) == 2 && !kFused && !SHORT_RING) ? ACTK_GROUP_WIDE : kGroup;
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
  const int bi = P.first_branch + ((zi >> 1) ^ P.long_first);
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
  ChannelScan<POWER_A, k16> cs;
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
        if (done < seg) __nanosleep(ACTK_CHAIN_POLL_NS);
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
          uint32_t dv[kG] = {}, dn[kG] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc + (k ? kT - kG : 0), dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (; r < kT; r += kG) {
            const int j0 = k ? kT - 1 - r : r, dj = k ? -1 : 1;
            if constexpr (kFused) {   // next group's delta values travel from TMEM while this group is scanned
              if (r + kG < kT) tmem_ld4_issue(tacc + (k ? kT - 2 * kG - r : r + kG), dn);
            }
            sumdt += cs.template run_state<kG, true>(
                [&](int i) { return IO<T>::ld(us + (j0 + dj * i) * kCh); },
                [&](int i) {
                  if constexpr (kFused) return IO<T>::rnd(__uint_as_float(k ? dv[kG - 1 - i] : dv[i]));
                  else return IO<T>::ld(ds + (j0 + dj * i) * kCh);
                },
                [&](int i) { return bcs + (j0 + dj * i) * 2 * kN; });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kG; ++i) dv[i] = dn[i];
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
        // smem row of step r: r (direction 0) or 15 - r (direction 1); kG steps are software-pipelined
        if (k == 0) {
          uint32_t dv[kG] = {}, dn[kG] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc, dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (int r0 = 0; r0 < kT; r0 += kG) {
            const T *u0 = us + r0 * kCh, *dl0 = ds + r0 * kCh;
            const float *b0 = bcs + r0 * 2 * kN;
            T *y0 = ys + r0 * kCh;
            if constexpr (kFused) {   // next group's delta values travel from TMEM while this group is scanned
              if (r0 + kG < kT) tmem_ld4_issue(tacc + r0 + kG, dn);
            }
            cs.template run<kG, true>([&](int i) { return IO<T>::ld(u0 + i * kCh); },
                                          [&](int i) {
                                            if constexpr (kFused) return IO<T>::rnd(__uint_as_float(dv[i]));
                                            else return IO<T>::ld(dl0 + i * kCh);
                                          },
                                          [&](int i) { return b0 + i * 2 * kN; },
                                          [&](int i, float y) { IO<T>::st(y0 + i * kCh, y); });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kG; ++i) dv[i] = dn[i];
            }
          }
        } else {
          uint32_t dv[kG] = {}, dn[kG] = {};
          if constexpr (kFused) { tmem_ld4_issue(tacc + kT - kG, dv); tmem_ld4_wait(dv); }
#pragma unroll 1
          for (int r0 = 0; r0 < kT; r0 += kG) {
            const int j0 = kT - 1 - r0;
            const T *u0 = us + j0 * kCh, *dl0 = ds + j0 * kCh;
            const float *b0 = bcs + j0 * 2 * kN;
            T *y0 = ys + j0 * kCh;
            if constexpr (kFused) {
              if (r0 + kG < kT) tmem_ld4_issue(tacc + kT - 2 * kG - r0, dn);
            }
            cs.template run<kG, true>([&](int i) { return IO<T>::ld(u0 - i * kCh); },
                                          [&](int i) {
                                            if constexpr (kFused) return IO<T>::rnd(__uint_as_float(dv[kG - 1 - i]));
                                            else return IO<T>::ld(dl0 - i * kCh);
                                          },
                                          [&](int i) { return b0 - i * 2 * kN; },
                                          [&](int i, float y) { IO<T>::st(y0 - i * kCh, y); });
            if constexpr (kFused) {
              tmem_ld4_wait(dn);
#pragma unroll
              for (int i = 0; i < kG; ++i) dv[i] = dn[i];
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

// One kernel launch for a given (T, KS); MODE and POWER_A are runtime here.  mode | 8 asks for the 3-slot ring
// (16-bit I/O without the fused dt_proj, single-level and chained launches).
template <typename T, int KS>
void launch_ks(bool pw, int mode, dim3 grid, cudaStream_t stream, const MaskedParams<T> &P, const MaskedMaps &M) {
  const bool short_ring = (mode & 8) != 0;
  mode &= 7;
  if constexpr (sizeof(T) == 2 && KS == 0) {
    if (short_ring && mode == 2) {
      if (pw) masked_scan_kernel<T, true, 2, KS, true><<<grid, kCh, 0, stream>>>(P, M);
      else masked_scan_kernel<T, false, 2, KS, true><<<grid, kCh, 0, stream>>>(P, M);
      return;
    }
    if (short_ring && mode == 0) {
      if (pw) masked_scan_kernel<T, true, 0, KS, true><<<grid, kCh, 0, stream>>>(P, M);
      else masked_scan_kernel<T, false, 0, KS, true><<<grid, kCh, 0, stream>>>(P, M);
      return;
    }
  }
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

}  // namespace actk
