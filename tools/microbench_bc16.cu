// Microbenchmark: does reading the step's B|C row as 16-bit values (4 broadcast LDS.128 instead of 8) and widening them
// in registers beat the fp32 row the scan kernels read today?  The per-step time of the scan fits
//     41 packed + 13 scalar FMA-pipe operations (95 cycles) + 8 x LDS.128 (8 cycles each) = 159 SMSP-cycles per warp-step
// almost exactly (profiles/r02_lean_scan.txt), i.e. the broadcast loads do not overlap the FMA pipe — either because they
// hold the dispatch port or because their 4 KB of register writes per warp-step compete for the register file's write
// port.  If it is the port, widening in registers (32 more register writes) buys nothing; this measures it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -I. tools/microbench_bc16.cu -o /tmp/mb16 && /tmp/mb16
#include <cstdio>
#include <vector>

#include "../actalker_b200/csrc/scan_core.cuh"

namespace actk {
void set_error(const char *, ...) {}
}
using namespace actk;

// WIDEN 0: fp32 row (the kernels' form).  1: bf16 row, (w << 16, w & 0xffff0000).  2: bf16 row, two PRMT.
template <int WIDEN>
__device__ __forceinline__ float apply_bc(ChannelScan<false, true> &cs, const uint64_t (&p)[kN / 2], const StepIn &s,
                                          const void *row) {
  if (WIDEN == 0) return cs.apply(p, s, static_cast<const float *>(row));
  const uint4 *r4 = static_cast<const uint4 *>(row);   // 64 bytes: B0..B15 | C0..C15 as bf16
  uint64_t x2 = pk(s.x, s.x);
  uint64_t ya = pk(0.f, 0.f), yb = pk(0.f, 0.f);
  auto widen = [](uint32_t w) -> uint64_t {
    uint32_t lo, hi;
    if (WIDEN == 1) { lo = w << 16; hi = w & 0xffff0000u; }
    else { lo = __byte_perm(w, 0, 0x1044); hi = __byte_perm(w, 0, 0x3244); }
    return pk(__uint_as_float(lo), __uint_as_float(hi));
  };
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint4 Bq = r4[half], Cq = r4[2 + half];   // 8 B values, 8 C values
    const uint32_t bw[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, cw[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n2 = half * 4 + j;                   // state pair
      cs.h[n2] = fma2(p[n2], cs.h[n2], mul2(x2, widen(bw[j])));
      if (j & 1) yb = fma2(widen(cw[j]), cs.h[n2], yb); else ya = fma2(widen(cw[j]), cs.h[n2], ya);
    }
  }
  float y0, y1;
  upk(add2(ya, yb), y0, y1);
  return fmaf(cs.dskip, s.u, y0 + y1);
}

template <int WIDEN>
__global__ void __launch_bounds__(64) step_kernel(const float *A, const float *in, float *out, int steps) {
  __shared__ alignas(16) float bc[16][32];   // WIDEN != 0 reads the first 64 bytes of a row as 32 bf16 values
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = in[i] * 0.01f;
  __syncthreads();
  ChannelScan<false, true> cs;
  cs.init(A + (threadIdx.x % 64) * 16, 1.0f, -2.0f);
  float u = in[threadIdx.x], d = in[512 + threadIdx.x];
  float acc = 0.f;
  for (int s = 0; s < steps; s += 8) {
    // the kernels' schedule: 8 prologues, then decay(i + 1) ahead of apply(i)
    StepIn si[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) si[i] = cs.template prologue<true>(u + 1e-3f * i, d + acc * 1e-9f + 1e-3f * i);
    uint64_t p[2][kN / 2];
    cs.decay(si[0].dt, p[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i + 1 < 8) cs.decay(si[i + 1].dt, p[(i + 1) & 1]);
      acc += apply_bc<WIDEN>(cs, p[i & 1], si[i], bc[(s + i) & 15]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIDEN>
static void run(const char *name, const float *A, const float *in, float *out) {
  const int steps = 4096;
  printf("%s\n", name);
  for (int warps_per_sm : {8, 14, 16, 28}) {
    int ctas = 148 * warps_per_sm / 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    step_kernel<WIDEN><<<ctas, 64>>>(A, in, out, steps);
    cudaEventRecord(a);
    step_kernel<WIDEN><<<ctas, 64>>>(A, in, out, steps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double warp_steps = (double)ctas * 2 * steps;
    printf("  warps/SM %2d: %.3f ms, %.1f SMSP-cycles per warp-step\n", warps_per_sm, ms,
           ms * 1e-3 * 1.965e9 / (warp_steps / (148 * 4)));
  }
}

int main() {
  float *A, *in, *out;
  std::vector<float> hA(64 * 16), hin(1024);
  for (int d = 0; d < 64; ++d)
    for (int n = 0; n < 16; ++n) hA[d * 16 + n] = -(n + 1.0f) * (1.0f + 0.013f * d + 0.001f * n * n);
  for (int i = 0; i < 1024; ++i) hin[i] = 0.5f + 0.001f * i;
  cudaMalloc(&A, hA.size() * 4);
  cudaMalloc(&in, hin.size() * 4);
  cudaMalloc(&out, 148 * 64 * 64 * 4);
  cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(in, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice);
  run<0>("general step, fp32 B|C row (8 LDS.128)", A, in, out);
  run<1>("general step, bf16 B|C row (4 LDS.128) + SHL / LOP3 widening", A, in, out);
  run<2>("general step, bf16 B|C row (4 LDS.128) + PRMT widening", A, in, out);
  return cudaDeviceSynchronize() != cudaSuccess;
}
