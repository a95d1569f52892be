// C-ABI housekeeping: error text, version, A-structure probe, algorithmic-bytes helper.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace actk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Single CTA: every thread checks a strided slice, the block ANDs the verdicts and thread 0 publishes it.
__global__ void __launch_bounds__(1024) a_structure_kernel(const float *__restrict__ A, int dim, int N, float rel_tol,
                                                           int *flag) {
  int ok = 1;
  for (int i = threadIdx.x; i < dim * N; i += blockDim.x) {
    const int d = i / N, n = i % N;
    const float want = (float)(n + 1) * A[(size_t)d * N];
    if (!(fabsf(A[i] - want) <= rel_tol * fabsf(want))) ok = 0;
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) *flag = ok ? ACTK_A_POWER : ACTK_A_GENERAL;
}

// Row gather: dst[b][p][:] = src[b][idx[p]][:], rows of row_bytes (a multiple of 16).  One warp per output row, 16 bytes
// per lane and trip; grid-stride over rows.  Replaces torch's index_select on the partial-mask path (mamba_layer.py:1963,
// 1974: xz[:, idx, :]), whose generic gather kernel reached 1.3 TB/s on B200 (0.21 ms of a 1.66 ms layer call at config 3's
// rectangle masks).
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4 *__restrict__ src, const int *__restrict__ idx,
                                                          uint4 *__restrict__ dst, long long rows_out, int n_idx, int rows_src,
                                                          int row_vec) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows_out; r += nwarps) {
    const long long b = r / n_idx;
    const int p = (int)(r - b * n_idx);
    const uint4 *s = src + (b * rows_src + __ldg(idx + p)) * row_vec;
    uint4 *d = dst + r * row_vec;
    for (int v = lane; v < row_vec; v += 32) d[v] = __ldg(s + v);
  }
}

}  // namespace actk

using namespace actk;

extern "C" int actk_gather_rows(const void *src, const int *idx, void *dst, int batch, int rows_src, int n_idx, long long row_bytes,
                                void *stream) {
  if (!src || !idx || !dst) ACTK_FAIL(ACTK_ERR_BAD_ARG, "gather_rows: NULL pointer");
  if (batch <= 0 || rows_src <= 0 || n_idx < 0 || row_bytes <= 0) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "gather_rows: batch=%d rows_src=%d n_idx=%d row_bytes=%lld", batch, rows_src, n_idx, row_bytes);
  if (row_bytes % 16 || (reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15)
    ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "gather_rows: rows of %lld bytes / pointers must be multiples of 16 bytes", row_bytes);
  if (n_idx == 0) return ACTK_OK;
  const long long rows_out = (long long)batch * n_idx;
  long long blocks = (rows_out + 7) / 8;            // 8 warps per block
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4 *>(src), idx, static_cast<uint4 *>(dst), rows_out, n_idx, rows_src, (int)(row_bytes / 16));
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

extern "C" int actk_abi_version(void) { return ACTK_ABI_VERSION; }
extern "C" int actk_sm_arch(void) { return 100; }
extern "C" const char *actk_last_error(void) { return g_err; }

extern "C" int actk_a_structure(const float *A, int dim, int dstate, float rel_tol, int *flag_dev, void *stream) {
  if (!A || !flag_dev) ACTK_FAIL(ACTK_ERR_BAD_ARG, "actk_a_structure: NULL pointer");
  if (dim <= 0 || dstate <= 0) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "actk_a_structure: dim=%d dstate=%d", dim, dstate);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  a_structure_kernel<<<1, 1024, 0, st>>>(A, dim, dstate, rel_tol, flag_dev);
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

// ---- gather buffers of the fused push all-gather (multi-GPU channel sharding, one process per GPU) -----------------
// Plain cudaMalloc allocations exported / imported with CUDA IPC: the importer maps a peer's buffer into ITS device's
// address space with peer access enabled, so the merge kernel can store into it over NVLink.
extern "C" int actk_peer_buffer_alloc(long long bytes, void **ptr) {
  if (!ptr || bytes <= 0) ACTK_FAIL(ACTK_ERR_BAD_ARG, "peer_buffer_alloc: bytes=%lld", bytes);
  ACTK_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
  return ACTK_OK;
}
extern "C" int actk_peer_buffer_free(void *ptr) {
  ACTK_CUDA_OK(cudaFree(ptr));
  return ACTK_OK;
}
extern "C" int actk_peer_buffer_export(void *ptr, unsigned char *handle64) {
  if (!ptr || !handle64) ACTK_FAIL(ACTK_ERR_BAD_ARG, "peer_buffer_export: NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  ACTK_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return ACTK_OK;
}
extern "C" int actk_peer_buffer_open(const unsigned char *handle64, void **ptr) {
  if (!ptr || !handle64) ACTK_FAIL(ACTK_ERR_BAD_ARG, "peer_buffer_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  ACTK_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ACTK_OK;
}
extern "C" int actk_peer_buffer_close(void *ptr) {
  ACTK_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return ACTK_OK;
}

extern "C" long long actk_scan_algorithmic_bytes(int batch, int seqlen, int dim, int groups, int dstate, int elsize) {
  // u, delta read + y written at (batch, dim, seqlen); B, C read at (batch, groups, dstate, seqlen);
  // A, D, delta_bias (fp32) once.  SURVEY.md §8(d).
  const long long act = (long long)batch * seqlen * (3LL * dim + 2LL * groups * dstate) * elsize;
  const long long par = (long long)dim * (dstate + 2) * 4;
  return act + par;
}
