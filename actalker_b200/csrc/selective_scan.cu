// Operator-contract selective scan (C-ABI entry actk_selective_scan_fwd): the drop-in for
// mamba_ssm.ops.selective_scan_interface.selective_scan_fn as the reference calls it
// (src/models/base/mamba_layer.py:1532-1538): channels-first (batch, dim, seqlen) tensors with seqlen
// contiguous, grouped B/C (batch, groups, dstate, seqlen), fp32 A/D/delta_bias, optional z gate.
//
// Same per-channel register recurrence as the fused layer kernel (scan_core.cuh).  Because the contract is
// channels-first with arbitrary seqlen (L' = 5217 is odd, so rows are not 16-byte aligned and bulk/TMA copies
// are not legal), tiles of 64 channels x 32 steps are transposed through padded shared memory with coalesced
// 64-byte row segments on the global side; the loads of tile i+1 are staged in registers while tile i is scanned.
// dstate == 16 takes this kernel; other dstate <= 64 a plain one.
#include "scan_core.cuh"

namespace actk {

constexpr int kOpCh = 64;   // channels per CTA (one thread each)
constexpr int kOpT = 32;    // steps per tile == lanes of the transposing loads
constexpr int kBcPitch = 2 * kN + 4;  // fp32 row pitch of the B|C tile: keeps rows 16-byte aligned, 4-way conflicts on fill

struct OpParams {
  actk_scan_args a;
  int ch_per_group, blocks_per_group;
};

template <typename T>
struct OpTile {
  static constexpr int kPad = 4 / sizeof(T);       // 2-byte: +2, 4-byte: +1 -> conflict-free transposes
  T u[kOpT][kOpCh + kPad];
  T dt[kOpT][kOpCh + kPad];
  T zy[kOpT][kOpCh + kPad];                        // z on the way in, y on the way out
  alignas(16) float bc[kOpT][kBcPitch];
};

template <typename T, bool POWER_A, bool SOFTPLUS, bool HAS_Z>
__global__ void __launch_bounds__(kOpCh) selective_scan_kernel(const __grid_constant__ OpParams P) {
  __shared__ OpTile<T> tile;
  const actk_scan_args &a = P.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.x / P.blocks_per_group;
  const int c0 = g * P.ch_per_group + (blockIdx.x % P.blocks_per_group) * kOpCh;   // first channel of this CTA
  const int c_end = min((g + 1) * P.ch_per_group, a.dim);
  const int nch = min(kOpCh, c_end - c0);
  const int b = blockIdx.y;
  const int L = a.seqlen;

  const T *u = (const T *)a.u + (size_t)b * a.u_sb;
  const T *dl = (const T *)a.delta + (size_t)b * a.delta_sb;
  const T *z = HAS_Z ? (const T *)a.z + (size_t)b * a.z_sb : nullptr;
  T *out = (T *)a.out + (size_t)b * a.out_sb;
  const T *Bg = (const T *)a.B + (size_t)b * a.B_sb + (size_t)g * a.B_sg;
  const T *Cg = (const T *)a.C + (size_t)b * a.C_sb + (size_t)g * a.C_sg;

  const bool live = tid < nch;
  const int ch = c0 + (live ? tid : 0);
  ChannelScan<POWER_A, sizeof(T) == 2> cs;   // 16-bit I/O: softplus without the small-argument series (scan_core.cuh)
  cs.init(a.A + (size_t)ch * kN, a.D ? a.D[ch] : 0.f, a.delta_bias ? a.delta_bias[ch] : 0.f);

  // Register-staged tile pipeline: every global load of tile i+1 is issued (as ~80 independent LDGs per thread)
  // before tile i is scanned, so DRAM latency hides behind the compute instead of stalling each row copy
  // (the first version waited load -> STS row by row: 52 % of warp samples sat on long_scoreboard).
  constexpr int kRows = kOpCh / 2;   // channel rows moved by each of the 2 warps
  constexpr int kBcRows = kN / 2;    // B rows (and C rows) moved by each warp
  T pu[kRows], pd[kRows], pz[HAS_Z ? kRows : 1], pb[kBcRows], pc[kBcRows];
  auto fetch = [&](int t0) {
    const bool ok = t0 + lane < L;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int cc = warp + 2 * i;
      const bool v = ok && cc < nch;
      pu[i] = v ? u[(size_t)(c0 + cc) * a.u_sd + t0 + lane] : T{};
      pd[i] = v ? dl[(size_t)(c0 + cc) * a.delta_sd + t0 + lane] : T{};
      if (HAS_Z) pz[i] = v ? z[(size_t)(c0 + cc) * a.z_sd + t0 + lane] : T{};
    }
#pragma unroll
    for (int i = 0; i < kBcRows; ++i) {
      const int n = warp + 2 * i;
      pb[i] = ok ? Bg[(size_t)n * a.B_sn + t0 + lane] : T{};
      pc[i] = ok ? Cg[(size_t)n * a.C_sn + t0 + lane] : T{};
    }
  };
  auto stash = [&]() {   // registers -> time-major shared tile (padded: conflict-free)
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int cc = warp + 2 * i;
      tile.u[lane][cc] = pu[i];
      tile.dt[lane][cc] = pd[i];
      if (HAS_Z) tile.zy[lane][cc] = pz[i];
    }
#pragma unroll
    for (int i = 0; i < kBcRows; ++i) {
      const int n = warp + 2 * i;
      tile.bc[lane][n] = IO<T>::f(pb[i]);
      tile.bc[lane][kN + n] = IO<T>::f(pc[i]);
    }
  };

  fetch(0);
  for (int t0 = 0; t0 < L; t0 += kOpT) {
    const int nt = min(kOpT, L - t0);
    stash();
    __syncthreads();
    if (t0 + kOpT < L) fetch(t0 + kOpT);
    // ---- scan the tile: one channel per thread, 4 steps software-pipelined
    if (live) {
      const T *us = &tile.u[0][tid], *ds = &tile.dt[0][tid];
      T *zy = &tile.zy[0][tid];
      constexpr int kPitch = kOpCh + OpTile<T>::kPad;
      int r = 0;
#ifndef ACTK_OP_GROUP
#define ACTK_OP_GROUP 8   // 8-step groups: 1.306 -> 1.257 ms for one reference-shaped call (4: ACTK_OP_GROUP=4)
#endif
      for (; r + ACTK_OP_GROUP <= nt; r += ACTK_OP_GROUP) {
        cs.template run<ACTK_OP_GROUP, SOFTPLUS>([&](int i) { return IO<T>::ld(us + (r + i) * kPitch); },
                                     [&](int i) { return IO<T>::ld(ds + (r + i) * kPitch); },
                                     [&](int i) { return tile.bc[r + i]; },
                                     [&](int i, float y) {
                                       if (HAS_Z) y *= silu(IO<T>::ld(zy + (r + i) * kPitch));
                                       IO<T>::st(zy + (r + i) * kPitch, y);
                                     });
      }
      for (; r < nt; ++r) {
        float y = cs.template step<SOFTPLUS>(IO<T>::ld(us + r * kPitch), IO<T>::ld(ds + r * kPitch), tile.bc[r]);
        if (HAS_Z) y *= silu(IO<T>::ld(zy + r * kPitch));
        IO<T>::st(zy + r * kPitch, y);
      }
    }
    __syncthreads();
    // ---- drain y: lanes along time again, coalesced 64-byte row segments
    if (lane < nt)
      for (int cc = warp; cc < nch; cc += kOpCh / 32) out[(size_t)(c0 + cc) * a.out_sd + t0 + lane] = tile.zy[lane][cc];
    __syncthreads();
  }
  if (a.last_state && live) {
    float *hs = a.last_state + ((size_t)b * a.dim + ch) * kN;
#pragma unroll
    for (int j = 0; j < kN / 2; ++j) upk(cs.h[j], hs[2 * j], hs[2 * j + 1]);
  }
}

// Plain kernel for dstate != 16 (1..64): one thread per (batch, channel), direct global access.
constexpr int kMaxGenericN = 64;
template <typename T>
__global__ void selective_scan_generic_kernel(const __grid_constant__ actk_scan_args a) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (ch >= a.dim) return;
  const int N = a.dstate, L = a.seqlen;
  const int g = ch / (a.dim / a.groups);
  const T *u = (const T *)a.u + (size_t)b * a.u_sb + (size_t)ch * a.u_sd;
  const T *dl = (const T *)a.delta + (size_t)b * a.delta_sb + (size_t)ch * a.delta_sd;
  const T *z = a.z ? (const T *)a.z + (size_t)b * a.z_sb + (size_t)ch * a.z_sd : nullptr;
  T *out = (T *)a.out + (size_t)b * a.out_sb + (size_t)ch * a.out_sd;
  const T *Bg = (const T *)a.B + (size_t)b * a.B_sb + (size_t)g * a.B_sg;
  const T *Cg = (const T *)a.C + (size_t)b * a.C_sb + (size_t)g * a.C_sg;
  float h[kMaxGenericN], al[kMaxGenericN];
  for (int n = 0; n < N; ++n) { h[n] = 0.f; al[n] = a.A[(size_t)ch * N + n] * kLog2e; }
  const float Dd = a.D ? a.D[ch] : 0.f, bias = a.delta_bias ? a.delta_bias[ch] : 0.f;
  for (int l = 0; l < L; ++l) {
    float uv = IO<T>::ld(u + l);
    float dt = IO<T>::ld(dl + l) + bias;
    if (a.delta_softplus) dt = softplus20(dt);
    float x = dt * uv, y = 0.f;
    for (int n = 0; n < N; ++n) {
      h[n] = ex2(dt * al[n]) * h[n] + x * IO<T>::ld(Bg + (size_t)n * a.B_sn + l);
      y = fmaf(IO<T>::ld(Cg + (size_t)n * a.C_sn + l), h[n], y);
    }
    y = fmaf(Dd, uv, y);
    if (z) y *= silu(IO<T>::ld(z + l));
    IO<T>::st(out + l, y);
  }
  if (a.last_state)
    for (int n = 0; n < N; ++n) a.last_state[((size_t)b * a.dim + ch) * N + n] = h[n];
}

template <typename T>
int launch_op(const actk_scan_args *a, cudaStream_t stream) {
  if (a->dstate != kN) {
    dim3 grid((a->dim + 63) / 64, a->batch);
    selective_scan_generic_kernel<T><<<grid, 64, 0, stream>>>(*a);
    ACTK_CUDA_OK(cudaGetLastError());
    return ACTK_OK;
  }
  OpParams P;
  P.a = *a;
  P.ch_per_group = a->dim / a->groups;
  P.blocks_per_group = (P.ch_per_group + kOpCh - 1) / kOpCh;
  dim3 grid(a->groups * P.blocks_per_group, a->batch);
  const bool pw = a->a_kind == ACTK_A_POWER, sp = a->delta_softplus != 0, hz = a->z != nullptr;
#define ACTK_OP_LAUNCH(PW, SP, HZ) selective_scan_kernel<T, PW, SP, HZ><<<grid, kOpCh, 0, stream>>>(P)
  if (pw) {
    if (sp) { if (hz) ACTK_OP_LAUNCH(true, true, true); else ACTK_OP_LAUNCH(true, true, false); }
    else    { if (hz) ACTK_OP_LAUNCH(true, false, true); else ACTK_OP_LAUNCH(true, false, false); }
  } else {
    if (sp) { if (hz) ACTK_OP_LAUNCH(false, true, true); else ACTK_OP_LAUNCH(false, true, false); }
    else    { if (hz) ACTK_OP_LAUNCH(false, false, true); else ACTK_OP_LAUNCH(false, false, false); }
  }
#undef ACTK_OP_LAUNCH
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

// This file is compiled four times (build.py): once per I/O dtype with -DACTK_TU_DTYPE=0/1/2, which instantiates the
// kernels of that dtype only, and once without it for the C-ABI entry — the instantiations build in parallel.
#if defined(ACTK_TU_DTYPE)
#if ACTK_TU_DTYPE == 0
template int launch_op<float>(const actk_scan_args *, cudaStream_t);
#elif ACTK_TU_DTYPE == 1
template int launch_op<__half>(const actk_scan_args *, cudaStream_t);
#else
template int launch_op<__nv_bfloat16>(const actk_scan_args *, cudaStream_t);
#endif
#else
extern template int launch_op<float>(const actk_scan_args *, cudaStream_t);
extern template int launch_op<__half>(const actk_scan_args *, cudaStream_t);
extern template int launch_op<__nv_bfloat16>(const actk_scan_args *, cudaStream_t);
#endif

}  // namespace actk

#ifndef ACTK_TU_DTYPE
using namespace actk;

extern "C" int actk_selective_scan_fwd(const actk_scan_args *a, void *stream) {
  if (!a) ACTK_FAIL(ACTK_ERR_BAD_ARG, "actk_selective_scan_fwd: args is NULL");
  if (a->dtype < ACTK_F32 || a->dtype > ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "selective_scan: dtype=%d", a->dtype);
  if (a->batch <= 0 || a->dim <= 0 || a->groups <= 0 || a->dstate <= 0 || a->seqlen <= 0)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "selective_scan: batch=%d dim=%d groups=%d dstate=%d seqlen=%d", a->batch, a->dim,
              a->groups, a->dstate, a->seqlen);
  if (a->dim % a->groups != 0)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "selective_scan: dim=%d is not a multiple of groups=%d", a->dim, a->groups);
  if (a->dstate > kMaxGenericN) ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "selective_scan: dstate=%d > %d", a->dstate, kMaxGenericN);
  if (a->batch > 65535) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "selective_scan: batch=%d exceeds grid.y", a->batch);
  if (!a->u || !a->delta || !a->A || !a->B || !a->C || !a->out)
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "selective_scan: u, delta, A, B, C and out are required");
  if (a->a_kind != ACTK_A_GENERAL && a->a_kind != ACTK_A_POWER)
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "selective_scan: a_kind=%d", a->a_kind);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case ACTK_F32: return launch_op<float>(a, st);
    case ACTK_F16: return launch_op<__half>(a, st);
    default: return launch_op<__nv_bfloat16>(a, st);
  }
}
#endif  // !ACTK_TU_DTYPE
