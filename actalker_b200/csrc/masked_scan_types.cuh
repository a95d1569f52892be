// Shared declarations of the fused masked scan: parameter blocks, tensor-map bundle and the launch function whose
// instantiations live in masked_scan_ks*.cu (one translation unit per dt-rank slab count, compiled in parallel).
#pragma once
#include <cuda.h>

#include "scan_core.cuh"

namespace actk {

constexpr int kCh = 64;  // channels per CTA == threads per CTA
constexpr int kT = 16;   // time steps per tile
constexpr int kGroup = 4;  // steps software-pipelined together (ChannelScan::run); see kG in the kernel
#ifndef ACTK_CHAIN_POLL_NS
#define ACTK_CHAIN_POLL_NS 200
#endif
#ifndef ACTK_LEAN_GROUP
#define ACTK_LEAN_GROUP 8   // lean kernel, 4-slot ring: steps per software-pipelined group
#endif
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
  const float *bc32, *bc32_tail;   // lean kernel: fp32 copy of x_dbl's B|C columns, (Bp, n_sel | n_tail, 4N)
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
  int long_first;   // 1: a two-branch launch hands out the LONGER branch's sequences first (grid z / ticket order), so the short
                    // branch's CTAs fill the tail instead of opening the launch (partial masks: 513 vs 1551 selected tokens)
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
  CUtensorMap bc32, ydir32;   // lean kernel: fp32 B|C boxes (2N wide), per-warp y boxes (32 channels wide)
};
struct alignas(64) MaskedMaps {
  BranchMaps m[2];
};


// One kernel launch for a given (T, KS); MODE and POWER_A are runtime here.  Defined in masked_scan_kernel.cuh.
template <typename T, int KS>
void launch_ks(bool pw, int mode, dim3 grid, cudaStream_t stream, const MaskedParams<T> &P, const MaskedMaps &M);

// Lean kernel (masked_scan_lean.cu): 16-bit I/O, idx == iota, D % 64 == 0, fp32 B|C given; plain or chained launch.
template <typename T>
void launch_lean(bool pw, bool chain, bool short_ring, dim3 grid, cudaStream_t stream, const MaskedParams<T> &P,
                 const MaskedMaps &M);

}  // namespace actk
