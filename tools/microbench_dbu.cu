// What would the scan's step cost if dBu[n] = (dt*u) * B[n] came out of tensor memory?  (B200, sm_100a)
//
// Today a channel-step loads the 32 fp32 B|C values (8 broadcast LDS.128) and forms dBu with 8 FMUL2.  The proposal: one
// tcgen05.mma per group of 8 steps computes D[channel][(t, n)] = X[channel][t'] * Bx[t'][(t, n)] (Bx = block-diagonal
// expansion of the B rows, X = dt*u split into three bf16 terms), and the scan thread reads its 16 dBu values per step with
// tcgen05.ld (lane = channel).  This microbenchmark times the CONSUMER side only — prologue (softplus), 16 decays, then
//   0: the step as it is (8 LDS.128, 16 FMUL2-lanes for dBu),
//   1: dBu from tensor memory (one tcgen05.ld.x16 per step, issued one step ahead), C still from shared memory (4 LDS.128)
// with 128-thread CTAs (lane = channel, what an M = 128 MMA fills) at 4 ... 16 warps per SM.  The producer side of the
// proposal (splitting dt*u, the MMA, two hand-overs per group) is NOT in it: the figures bound what the redesign can reach.
// Tensor memory holds whatever it held: only timing is measured.
#include <cstdio>
#include <vector>

#include "../actalker_b200/csrc/scan_core.cuh"

namespace actk {
void set_error(const char *, ...) {}
}
using namespace actk;

__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tm_wait16(uint32_t *r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// apply() of ChannelScan with dBu given: h = p*h + dBu, y += C*h
template <typename CS>
__device__ __forceinline__ float apply_dbu(CS &cs, const uint64_t (&p)[8], const StepIn &s, const uint32_t *dbu, const float *c16) {
  const ulonglong2 *c2 = reinterpret_cast<const ulonglong2 *>(c16);
  uint64_t ya = pk(0.f, 0.f), yb = pk(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const ulonglong2 Cq = c2[q];
    const uint64_t d0 = ((uint64_t)dbu[4 * q + 1] << 32) | dbu[4 * q], d1 = ((uint64_t)dbu[4 * q + 3] << 32) | dbu[4 * q + 2];
    cs.h[2 * q] = fma2(p[2 * q], cs.h[2 * q], d0);
    cs.h[2 * q + 1] = fma2(p[2 * q + 1], cs.h[2 * q + 1], d1);
    ya = fma2(Cq.x, cs.h[2 * q], ya);
    yb = fma2(Cq.y, cs.h[2 * q + 1], yb);
  }
  float y0, y1;
  upk(add2(ya, yb), y0, y1);
  return fmaf(cs.dskip, s.u, y0 + y1);
}

template <int SRC>
__global__ void __launch_bounds__(128) step_kernel(const float *A, const float *in, float *out, int steps) {
  __shared__ alignas(16) float bc[16][32];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = in[i] * 0.01f;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  ChannelScan<false> cs;
  cs.init(A + (threadIdx.x % 64) * 16, 1.0f, -2.0f);
  const float u = in[threadIdx.x], d = in[512 + threadIdx.x];
  float acc = 0.f;
  constexpr int G = 8;          // steps per group, as in the kernel
  uint32_t dbu[2][16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dbu[0][i] = dbu[1][i] = __float_as_uint(1e-3f * i);
  for (int s0 = 0; s0 < steps; s0 += G) {
    StepIn s[G];
#pragma unroll
    for (int i = 0; i < G; ++i) s[i] = cs.template prologue<true>(u + 1e-3f * i, d + 1e-6f * (float)s0 + 1e-3f * i);   // eight different dt, independent of the results
    uint64_t p[2][8];
    cs.decay(s[0].dt, p[0]);
    if (SRC == 1) { tm_ld16(tmem, dbu[0]); tm_wait16(dbu[0]); }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      if (i + 1 < G) {
        cs.decay(s[i + 1].dt, p[(i + 1) & 1]);
        if (SRC == 1) tm_ld16(tmem + 16 * ((i + 1) & 7), dbu[(i + 1) & 1]);
      }
      if (SRC == 0) {
        acc += cs.apply(p[i & 1], s[i], bc[(s0 + i) & 15]);
      } else {
        acc += apply_dbu(cs, p[i & 1], s[i], dbu[i & 1], &bc[(s0 + i) & 15][16]);
        if (i + 1 < G) tm_wait16(dbu[(i + 1) & 1]);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
}

template <int SRC>
static void run(const char *name, const float *A, const float *in, float *out) {
  const int steps = 4096;
  printf("%s\n", name);
  for (int warps_per_sm : {4, 8, 12, 16}) {
    const int ctas = 148 * warps_per_sm / 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    step_kernel<SRC><<<ctas, 128>>>(A, in, out, steps);
    cudaEventRecord(a);
    step_kernel<SRC><<<ctas, 128>>>(A, in, out, steps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_steps = (double)ctas * 4 * steps;
    printf("  warps/SM %2d: %.1f SMSP-cycles per warp-step\n", warps_per_sm, ms * 1e-3 * 1.965e9 / (warp_steps / (148 * 4)));
  }
}

int main() {
  float *A, *in, *out;
  std::vector<float> hA(64 * 16), hin(1024);
  for (int d = 0; d < 64; ++d)
    for (int n = 0; n < 16; ++n) hA[d * 16 + n] = -(n + 1.0f) * (1.0f + 0.01f * d) * (1.0f + 0.003f * n * n);   // general A
  for (int i = 0; i < 1024; ++i) hin[i] = 0.5f + 0.001f * i;
  cudaMalloc(&A, hA.size() * 4);
  cudaMalloc(&in, hin.size() * 4);
  cudaMalloc(&out, 148 * 64 * 128 * 4);
  cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(in, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice);
  run<0>("general step as it is (8 LDS.128 B|C, dBu on the FMA pipe), 8-step groups", A, in, out);
  run<1>("general step with dBu from tensor memory (tcgen05.ld.x16 per step, 4 LDS.128 for C)", A, in, out);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
