// Where should the scan's warp-uniform B|C row come from?  (B200, sm_100a)
//
// per channel-step every lane needs the step's 32 fp32 B|C values in registers; today they arrive as 8
// broadcast LDS.128, and those loads ADD to the packed FMA work of the step instead of hiding under it (159 = 95 + 64
// SMSP-cycles per warp-step).  Tensor memory is the other on-chip store a warp can fill registers from: one
// tcgen05.ld.32x32b.x32 delivers 32 registers per lane (lane = tensor-memory lane), and a row replicated over all 128 lanes
// (what an M = 128 MMA with a ones-column A operand writes) is a broadcast.  This microbenchmark times one "step" = 24
// packed FMAs that consume 32 freshly loaded registers, with the registers filled
//   0: not at all (the FMAs alone), 1: by 8 broadcast LDS.128, 2: by one tcgen05.ld.x32, 3: by two tcgen05.ld.x16,
//   4: by eight tcgen05.ld.x4,
// the next step's fill issued before the current step's FMAs (double-buffered registers), at 8 and 16 warps per SM.
// Prints SMSP-cycles per warp-step.  Tensor memory holds whatever it held: only timing is measured.
#include <cstdio>

#include "../actalker_b200/csrc/common.cuh"
namespace actk {
void set_error(const char *, ...) {}
}
using namespace actk;

__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tm_wait(uint32_t *r) {   // ties all 32 registers to the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

template <int SRC>
__device__ __forceinline__ void fill(uint32_t *r, const float *sm, uint32_t tmem, int step) {
  if (SRC == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 x = *reinterpret_cast<const float4 *>(&sm[((step & 15) * 32 + j * 4)]);
      r[4 * j] = __float_as_uint(x.x); r[4 * j + 1] = __float_as_uint(x.y); r[4 * j + 2] = __float_as_uint(x.z); r[4 * j + 3] = __float_as_uint(x.w);
    }
  }
  const uint32_t col = (uint32_t)(step & 3) * 32;       // 128 columns hold four steps' rows
  if (SRC == 2) tm_ld32(tmem + col, r);
  if (SRC == 3) { tm_ld16(tmem + col, r); tm_ld16(tmem + col + 16, r + 16); }
  if (SRC == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) tm_ld4(tmem + col + 4 * j, r + 4 * j);
  }
}

// 24 packed FMAs on 8 accumulator pairs, every loaded register used as an operand
__device__ __forceinline__ void consume(uint64_t (&v)[8], const uint32_t *r, uint64_t b2) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t p0 = ((uint64_t)r[4 * j + 1] << 32) | r[4 * j], p1 = ((uint64_t)r[4 * j + 3] << 32) | r[4 * j + 2];
    v[j] = fma2(v[j], p0, b2);
    v[(j + 3) & 7] = fma2(v[(j + 3) & 7], p1, b2);
    v[(j + 5) & 7] = fma2(v[(j + 5) & 7], b2, p0);
  }
}

template <int SRC>
__global__ void __launch_bounds__(128) step_kernel(float *out, const float *in, int iters) {
  __shared__ alignas(16) float sm[16 * 32];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) sm[i] = in[i];
  if (threadIdx.x < 32) {   // 128 columns: four CTAs per SM can hold their allocation at once
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);     // this warp's lane quadrant
  uint64_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = pk(in[i] + threadIdx.x, in[i] + 1.f);
  const uint64_t b2 = pk(in[17], in[18] + 1e-3f);
  uint32_t ra[32], rb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) ra[i] = rb[i] = __float_as_uint(in[i & 15]);
  if (SRC != 0) { fill<SRC>(ra, sm, tmem, 0); if (SRC >= 2) tm_wait(ra); }
  for (int it = 0; it < iters; it += 2) {
    if (SRC != 0) fill<SRC>(rb, sm, tmem, it + 1);
    consume(v, ra, b2);
    if (SRC >= 2) tm_wait(rb);
    if (SRC != 0) fill<SRC>(ra, sm, tmem, it + 2);
    consume(v, rb, b2);
    if (SRC >= 2) tm_wait(ra);
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float lo, hi; upk(v[i], lo, hi); r += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
}

template <int SRC>
static void run(const char *name, float *out, const float *in) {
  const int iters = 4096;
  for (int warps_per_sm : {4, 8, 16}) {
    const int ctas = 148 * warps_per_sm / 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    step_kernel<SRC><<<ctas, 128>>>(out, in, iters);
    cudaEventRecord(a);
    step_kernel<SRC><<<ctas, 128>>>(out, in, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double wsteps = (double)ctas * 4 * iters;
    printf("%-44s warps/SM %2d: %7.2f SMSP-cycles per warp-step\n", name, warps_per_sm, ms * 1e-3 * 1.965e9 / (wsteps / (148 * 4)));
  }
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 148 * 8 * 128 * 4 * 4);
  cudaMemset(in, 0, 4096);
  run<0>("24 FFMA2 alone", out, in);
  run<1>("24 FFMA2 + 8 broadcast LDS.128", out, in);
  run<2>("24 FFMA2 + 1 tcgen05.ld.32x32b.x32", out, in);
  run<3>("24 FFMA2 + 2 tcgen05.ld.32x32b.x16", out, in);
  run<4>("24 FFMA2 + 8 tcgen05.ld.32x32b.x4", out, in);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
