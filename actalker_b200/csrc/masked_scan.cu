// Host side of the fused masked bidirectional selective scan (C-ABI entry actk_masked_scan_fwd): argument checks,
// tensor maps, launch-shape selection, the inter-chunk carry kernel and the dt_proj weight packer.  The scan kernel
// itself is in masked_scan_kernel.cuh and is instantiated by masked_scan_ks{0,2,3,5}.cu.
#include <cuda.h>
#include <string.h>

#include "masked_scan_types.cuh"

namespace actk {

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

static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

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
  // Lean kernel (masked_scan_lean.cu): the caller asks for it by passing the fp32 B|C copies; taken when every live
  // branch has them and runs under an all-ones mask, with 16-bit I/O, full channel blocks, delta tensors, single level.
  bool lean = es == 2 && ks == 0 && a->D % kCh == 0 && nseg == 1;
  {
    bool any = false;
    for (int i = 0; i < a->n_branches && lean; ++i) {
      const actk_branch_args &s = a->br[i];
      if (s.n_sel == 0) continue;
      any = true;
      if (s.n_sel != a->L || !s.bc32 || (s.n_tail > 0 && !s.bc32_tail) || misaligned(s.bc32) || misaligned(s.bc32_tail)) lean = false;
    }
    lean = lean && any;
  }
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
    d.bc32 = s.bc32; d.bc32_tail = s.bc32_tail;
    if (P.tma_ok && i < a->n_branches && s.n_sel > 0) {
      const uint64_t Lp = (uint64_t)s.n_sel + s.n_tail;
      int rc;
      if ((rc = make_map(&M.m[i].xz, dt, es, s.xz, a->D, a->L, a->Bp, kCh))) return rc;
      if ((rc = make_map(&M.m[i].xdbl, dt, es, s.xdbl, a->xw, (uint64_t)s.n_sel, a->Bp, 2 * kN))) return rc;
      if (ks) {
        if ((rc = make_map_dtin(&M.m[i].xdbl_dt, dt, es, s.xdbl, a->xw, (uint64_t)s.n_sel, a->Bp, 16 * ks))) return rc;
      } else if ((rc = make_map(&M.m[i].delta, dt, es, s.delta, 2ull * a->D, (uint64_t)s.n_sel, a->Bp, kCh))) return rc;
      if ((rc = make_map(&M.m[i].ydir, dt, es, s.ydir, a->D, a->L, 2ull * a->Bp, kCh))) return rc;
      if (lean) {
        if ((rc = make_map(&M.m[i].bc32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s.bc32, 4 * kN, (uint64_t)s.n_sel, a->Bp, 2 * kN))) return rc;
        if ((rc = make_map(&M.m[i].ydir32, dt, es, s.ydir, a->D, a->L, 2ull * a->Bp, kCh / 2))) return rc;
      }
    }
  }
  // group consecutive live branches that share an A kind into one launch (better tail balance)
  int i = 0;
  while (i < a->n_branches) {
    if (a->br[i].n_sel == 0) { ++i; continue; }
    int j = i + 1;
    while (j < a->n_branches && a->br[j].n_sel > 0 && a->br[j].a_kind == a->br[i].a_kind) ++j;
    P.first_branch = i;
    P.long_first = (j - i == 2 && a->br[i + 1].n_sel + a->br[i + 1].n_tail > a->br[i].n_sel + a->br[i].n_tail) ? 1 : 0;
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
      if constexpr (sizeof(T) == 2) {
        if (lean) {
          launch_lean<T>(pw, true, P.nq > 7 * 148, dim3(nblocks), stream, P, M);
          ACTK_CUDA_OK(cudaGetLastError());
          i = j;
          continue;
        }
      }
      launch_any<T>(ks, pw, 2 | (P.nq > 7 * 148 ? 8 : 0), dim3(nblocks), stream, P, M);
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
    if constexpr (sizeof(T) == 2) {
      if (lean) {
        launch_lean<T>(pw, false, (long long)grid.x * grid.y * grid.z > 7 * 148, grid, stream, P, M);
        ACTK_CUDA_OK(cudaGetLastError());
        i = j;
        continue;
      }
    }
    // single level: more sequence-CTAs than 7 per SM can hold -> the 3-slot ring's 8th CTA per SM pays
    launch_any<T>(ks, pw, (nseg == 1 && (long long)grid.x * grid.y * grid.z > 7 * 148) ? 8 : 0, grid, stream, P, M);
    ACTK_CUDA_OK(cudaGetLastError());
    i = j;
  }
  return ACTK_OK;
}

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
