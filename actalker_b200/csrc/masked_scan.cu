// Fused masked bidirectional selective scan (C-ABI entry actk_masked_scan_fwd).
//
// Replaces, for one call of SS2D_cond_v10.forward (reference src/models/base/mamba_layer.py):
//   :1963/:1974  gather of the mask-selected tokens           -> rows are fetched through idx[] by the TMA producer
//   :1965-1967   cat([selected, id, cond])                    -> tail rows come from a second base pointer
//   :1508-1519   HSCANS_dynamic identity encode + flip + cat  -> direction 1 walks the same rows downwards
//   :1532-1538   selective_scan_fn (bias, softplus, scan, D)  -> ChannelScan::step, fp32 state in registers
//   :1969-1970   slice [:n_sel] + index_put_ scatter          -> y of position p is stored to latent row idx[p]
// The direction sum (:1542-1547) and branch sum (:1983) need the reference's rounding points and are done by
// actk_merge_layernorm_fwd, which reads the two per-direction outputs written here.
//
// Work decomposition (B200: 148 SMs): one CTA = 64 channels x one (batch, branch, direction); 2 compute warps
// (one thread per channel, 16 states in registers) + 1 producer warp.  The producer warp stages kT-step tiles
// of u / delta / B|C rows into a kStages-deep shared-memory ring with bulk async copies (TMA, SASS UBLKCP)
// completing on mbarriers, converts the B|C rows to fp32 once per CTA, and publishes the scatter rows.
// Config 2 (B'=25, D=640, 2 branches x 2 directions) gives 1000 CTAs = 6.8 per SM, all co-resident.
#include "scan_core.cuh"

namespace actk {

constexpr int kCh = 64;      // channels per CTA
constexpr int kT = 16;       // time steps per staged tile
constexpr int kStages = 4;   // ring depth
constexpr int kThreads = kCh + 32;

template <typename T>
struct BranchDev {
  const T *xz, *tail, *xdbl, *xdbl_tail, *delta;
  const int *idx;
  const float *A, *Dskip, *dt_bias;
  T *ydir;
  int n_sel, n_tail;
};
template <typename T>
struct MaskedParams {
  BranchDev<T> br[2];
  int first_branch;
  int Bp, L, D, xw;
};

template <typename T>
struct alignas(16) Stage {
  T u[kT][kCh];
  T dt[kT][kCh];
  float bc[kT][2 * kN];                                  // fp32 B|C, what the compute warps read
  T bc_raw[sizeof(T) == 4 ? 1 : kT][2 * kN];             // 16-bit landing zone (unused for fp32 I/O)
  int row[kT];                                           // latent row to scatter to, -1 = tail token (dropped)
};

template <typename T, bool POWER_A>
__global__ void __launch_bounds__(kThreads) masked_scan_kernel(const __grid_constant__ MaskedParams<T> P) {
  __shared__ Stage<T> st[kStages];
  __shared__ alignas(8) uint64_t full_bar[kStages], ready_bar[kStages], empty_bar[kStages];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int d0 = blockIdx.x * kCh;
  const int b = blockIdx.y;
  const int bi = P.first_branch + (blockIdx.z >> 1);
  const int k = blockIdx.z & 1;
  const BranchDev<T> br = P.br[bi];   // by value: keeps the fields in registers instead of indexed constant loads
  const int n_sel = br.n_sel, n_tail = br.n_tail;
  const int Lp = n_sel + n_tail;
  const int D = P.D, L = P.L;
  const int nch = min(kCh, D - d0);   // last channel block may be partial (D % 8 == 0 keeps rows 16-byte granular)
  const int ntiles = (Lp + kT - 1) / kT;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 32);
      mbar_init(&empty_bar[s], kCh / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kCh / 32) {
    // ------------------------------------------------------------------ producer warp
    const uint32_t kRowBytes = nch * sizeof(T);
    constexpr uint32_t kBcBytes = 2 * kN * sizeof(T);
    const int r = lane & (kT - 1);
    auto issue = [&](int tile) {
      const int s = tile % kStages;
      const uint32_t ph = (tile / kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      const int p0 = tile * kT;
      const int nrows = min(kT, Lp - p0);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], nrows * (2 * kRowBytes + kBcBytes));
      __syncwarp();
      if (r < nrows) {
        const int p = p0 + r;
        const int l = k ? Lp - 1 - p : p;
        if (lane < kT) {
          const T *usrc, *bsrc;
          int row = -1;
          if (l < n_sel) {
            row = __ldg(br.idx + l);
            const size_t tok = (size_t)b * L + row;
            usrc = br.xz + tok * D + d0;
            bsrc = br.xdbl + tok * P.xw + k * 2 * kN;
          } else {
            const size_t tok = (size_t)b * n_tail + (l - n_sel);
            usrc = br.tail + tok * D + d0;
            bsrc = br.xdbl_tail + tok * P.xw + k * 2 * kN;
          }
          bulk_g2s(&st[s].u[r][0], usrc, kRowBytes, &full_bar[s]);
          if (sizeof(T) == 4)
            bulk_g2s(&st[s].bc[r][0], bsrc, kBcBytes, &full_bar[s]);
          else
            bulk_g2s(&st[s].bc_raw[r][0], bsrc, kBcBytes, &full_bar[s]);
          st[s].row[r] = row;
        } else {
          const T *dsrc = br.delta + (((size_t)b * Lp + l) * 2 + k) * D + d0;
          bulk_g2s(&st[s].dt[r][0], dsrc, kRowBytes, &full_bar[s]);
        }
      }
    };
    const int pre = min(kStages - 1, ntiles);
    for (int t = 0; t < pre; ++t) issue(t);
    for (int t = 0; t < ntiles; ++t) {
      // publish tile t first (it landed while the compute warps were busy with earlier tiles) ...
      const int s = t % kStages;
      const uint32_t ph = (t / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      if (sizeof(T) == 2) {
        // 16 rows x 2 halves: each lane widens 16 values of one row to fp32
        const int rr = lane >> 1, half = lane & 1;
        const uint4 *src = reinterpret_cast<const uint4 *>(&st[s].bc_raw[rr][half * kN]);
        float *dst = &st[s].bc[rr][half * kN];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w = src[v];
          const T *e = reinterpret_cast<const T *>(&w);
          float4 lo = make_float4(IO<T>::ld(e + 0), IO<T>::ld(e + 1), IO<T>::ld(e + 2), IO<T>::ld(e + 3));
          float4 hi = make_float4(IO<T>::ld(e + 4), IO<T>::ld(e + 5), IO<T>::ld(e + 6), IO<T>::ld(e + 7));
          reinterpret_cast<float4 *>(dst)[2 * v] = lo;
          reinterpret_cast<float4 *>(dst)[2 * v + 1] = hi;
        }
      }
      mbar_arrive(&ready_bar[s]);
      // ... then refill the stage the compute warps released last (blocks until they are done with tile t-1)
      if (t + kStages - 1 < ntiles) issue(t + kStages - 1);
    }
  } else {
    // ------------------------------------------------------------------ compute warps: one channel per thread
    const int c = tid;
    const bool live = c < nch;
    const int ch = k * D + d0 + (live ? c : 0);
    ChannelScan<POWER_A> cs;
    cs.init(br.A + (size_t)ch * kN, br.Dskip[ch], br.dt_bias[ch]);
    T *ybase = br.ydir + ((size_t)k * P.Bp + b) * L * D + d0 + c;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      const uint32_t ph = (t / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      mbar_wait(&ready_bar[s], ph);
      const int nrows = live ? min(kT, Lp - t * kT) : 0;
      if (nrows == kT) {
#pragma unroll 4
        for (int r = 0; r < kT; ++r) {
          float y = cs.template step<true>(IO<T>::ld(&st[s].u[r][c]), IO<T>::ld(&st[s].dt[r][c]), st[s].bc[r]);
          const int row = st[s].row[r];
          if (row >= 0) IO<T>::st(ybase + (size_t)row * D, y);
        }
      } else {
        for (int r = 0; r < nrows; ++r) {
          float y = cs.template step<true>(IO<T>::ld(&st[s].u[r][c]), IO<T>::ld(&st[s].dt[r][c]), st[s].bc[r]);
          const int row = st[s].row[r];
          if (row >= 0) IO<T>::st(ybase + (size_t)row * D, y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
}

template <typename T>
static int launch_masked(const actk_masked_scan_args *a, cudaStream_t stream) {
  MaskedParams<T> P;
  P.Bp = a->Bp; P.L = a->L; P.D = a->D; P.xw = a->xw;
  for (int i = 0; i < 2; ++i) {
    const actk_branch_args &s = a->br[i < a->n_branches ? i : 0];
    BranchDev<T> &d = P.br[i];
    d.xz = (const T *)s.xz; d.tail = (const T *)s.tail; d.xdbl = (const T *)s.xdbl;
    d.xdbl_tail = (const T *)s.xdbl_tail; d.delta = (const T *)s.delta; d.idx = s.idx;
    d.A = s.A; d.Dskip = s.Dskip; d.dt_bias = s.dt_bias; d.ydir = (T *)s.ydir;
    d.n_sel = s.n_sel; d.n_tail = s.n_tail;
  }
  // group consecutive live branches that share an A kind into one launch (better tail balance)
  int i = 0;
  while (i < a->n_branches) {
    if (a->br[i].n_sel == 0) { ++i; continue; }
    int j = i + 1;
    while (j < a->n_branches && a->br[j].n_sel > 0 && a->br[j].a_kind == a->br[i].a_kind) ++j;
    P.first_branch = i;
    dim3 grid((a->D + kCh - 1) / kCh, a->Bp, 2 * (j - i));
    if (a->br[i].a_kind == ACTK_A_POWER)
      masked_scan_kernel<T, true><<<grid, kThreads, 0, stream>>>(P);
    else
      masked_scan_kernel<T, false><<<grid, kThreads, 0, stream>>>(P);
    ACTK_CUDA_OK(cudaGetLastError());
    i = j;
  }
  return ACTK_OK;
}

static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

}  // namespace actk

using namespace actk;

extern "C" int actk_masked_scan_fwd(const actk_masked_scan_args *a, void *stream) {
  if (!a) ACTK_FAIL(ACTK_ERR_BAD_ARG, "actk_masked_scan_fwd: args is NULL");
  if (a->dtype < ACTK_F32 || a->dtype > ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "masked_scan: dtype=%d", a->dtype);
  if (a->n_branches < 1 || a->n_branches > 2) ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: n_branches=%d", a->n_branches);
  if (a->N != kN) ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "masked_scan: d_state=%d, this build has %d", a->N, kN);
  if (a->Bp <= 0 || a->L <= 0 || a->D <= 0 || a->xw < 4 * kN)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: Bp=%d L=%d D=%d xw=%d", a->Bp, a->L, a->D, a->xw);
  if (a->Bp > 65535) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "masked_scan: Bp=%d exceeds grid.y", a->Bp);
  const int es = a->dtype == ACTK_F32 ? 4 : 2;
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
    if (!s.xz || !s.xdbl || !s.delta || !s.idx || !s.A || !s.Dskip || !s.dt_bias || !s.ydir ||
        (s.n_tail > 0 && (!s.tail || !s.xdbl_tail)))
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "masked_scan: branch %d has a NULL pointer", i);
    if (misaligned(s.xz) || misaligned(s.tail) || misaligned(s.xdbl) || misaligned(s.xdbl_tail) ||
        misaligned(s.delta) || misaligned(s.ydir))
      ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "masked_scan: branch %d has a pointer not aligned to 16 bytes", i);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case ACTK_F32: return launch_masked<float>(a, st);
    case ACTK_F16: return launch_masked<__half>(a, st);
    default: return launch_masked<__nv_bfloat16>(a, st);
  }
}
