// Microbenchmark: throughput of the scan's per-step instruction mix (ChannelScan::prologue/decay/apply) with the
// B|C rows in shared memory and no global traffic, as a function of resident warps per SM.  Answers two
// questions the kernel design depends on: what is the pipe-bound step rate of this mix on B200, and how many
// warps per scheduler does it take to reach it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -I. tools/microbench_step.cu -o /tmp/mb && /tmp/mb
#include <cstdio>
#include <vector>

#include "../actalker_b200/csrc/scan_core.cuh"
#include "scan_core2.cuh"

namespace actk {
void set_error(const char *, ...) {}
}
using namespace actk;

template <bool POWER_A, int MODE>  // MODE 0: full step, 1: apply only (decay hoisted), 2: prologue+decay only
__global__ void __launch_bounds__(256) step_kernel(const float *A, const float *in, float *out, int steps) {
  __shared__ alignas(16) float bc[16][32];
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = in[i] * 0.01f;
  __syncthreads();
  ChannelScan<POWER_A> cs;
  cs.init(A + (threadIdx.x % 64) * 16, 1.0f, -2.0f);
  float u = in[threadIdx.x], d = in[512 + threadIdx.x];
  float acc = 0.f;
  uint64_t p[8];
  cs.decay(0.01f, p);
  for (int s = 0; s < steps; s += 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ui = u + 1e-3f * i, di = d + acc * 1e-9f;
      if (MODE == 0) {
        acc += cs.template step<true>(ui, di, bc[(s + i) & 15]);
      } else if (MODE == 1) {
        StepIn si{0.01f, ui * 0.01f, ui};
        acc += cs.apply(p, si, bc[(s + i) & 15]);
      } else {
        StepIn si = cs.template prologue<true>(ui, di);
        cs.decay(si.dt, p);
        float lo, hi;
        upk(p[7], lo, hi);
        acc += lo + hi + si.x;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// two channels per thread, one 32-thread CTA per 64 channels (the layout under evaluation)
template <bool POWER_A, int NPOLY>
__global__ void __launch_bounds__(32) step2_kernel(const float *A, const float *in, float *out, int steps) {
  __shared__ alignas(16) float bc[16][32];
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = in[i] * 0.01f;
  __syncwarp();
  ChannelScan2<POWER_A, NPOLY> cs;
  cs.init(A + (2 * threadIdx.x) * 16, A + (2 * threadIdx.x + 1) * 16, 1.0f, 1.0f, -2.0f, -2.1f);
  float u = in[threadIdx.x], d = in[512 + threadIdx.x];
  float acc = 0.f;
  for (int s = 0; s < steps; s += 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ui = u + 1e-3f * i, di = d + acc * 1e-9f;
      float y0, y1;
      upk(cs.template step<true>(ui, ui + 0.5f, di, di - 0.25f, bc[(s + i) & 15]), y0, y1);
      acc += y0 + y1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <bool POWER_A, int NPOLY>
static void run2(const char *name, const float *A, const float *in, float *out) {
  const int steps = 4096;
  printf("%s\n", name);
  for (int warps_per_sm : {4, 8, 12, 16}) {
    int ctas = 148 * warps_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    step2_kernel<POWER_A, NPOLY><<<ctas, 32>>>(A, in, out, steps);
    cudaEventRecord(a);
    step2_kernel<POWER_A, NPOLY><<<ctas, 32>>>(A, in, out, steps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double warp_steps = (double)ctas * steps;
    double cyc = ms * 1e-3 * 1.965e9 / (warp_steps / (148 * 4));
    printf("  warps/SM %2d: %.3f ms, %.1f SMSP-cycles per warp-step (= %.1f per 32 channel-steps), %.2f G chan-steps/s\n",
           warps_per_sm, ms, cyc, cyc / 2, warp_steps * 64 / (ms * 1e-3) / 1e9);
  }
}

template <bool POWER_A, int MODE>
static void run(const char *name, const float *A, const float *in, float *out) {
  const int steps = 4096;
  printf("%s\n", name);
  for (int warps_per_sm : {4, 8, 12, 14, 16, 24, 32, 48}) {
    // 64-thread CTAs like the real kernel; warps_per_sm/2 CTAs per SM
    int ctas = 148 * warps_per_sm / 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    step_kernel<POWER_A, MODE><<<ctas, 64>>>(A, in, out, steps);
    cudaEventRecord(a);
    step_kernel<POWER_A, MODE><<<ctas, 64>>>(A, in, out, steps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double warp_steps = (double)ctas * 2 * steps;
    double cyc_per_step_smsp = ms * 1e-3 * 1.965e9 / (warp_steps / (148 * 4));
    printf("  warps/SM %2d: %.3f ms, %.1f SMSP-cycles per warp-step, %.2f G chan-steps/s\n", warps_per_sm, ms,
           cyc_per_step_smsp, warp_steps * 32 / (ms * 1e-3) / 1e9);
  }
}

int main(int argc, char **) {
  float *A, *in, *out;
  std::vector<float> hA(64 * 16), hin(1024);
  for (int d = 0; d < 64; ++d)
    for (int n = 0; n < 16; ++n) hA[d * 16 + n] = -(n + 1.0f) * (1.0f + 0.01f * d);
  for (int i = 0; i < 1024; ++i) hin[i] = 0.5f + 0.001f * i;
  cudaMalloc(&A, hA.size() * 4);
  cudaMalloc(&in, hin.size() * 4);
  cudaMalloc(&out, 148 * 64 * 64 * 4);
  cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(in, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice);
  if (argc > 1) {
    run<true, 0>("power, full step", A, in, out);
    run<false, 0>("general, full step", A, in, out);
    run<true, 1>("apply only (24 packed FMA ops + 8 LDS.128)", A, in, out);
  }
  run2<true, 0>("power, 2 channels/thread", A, in, out);
  run2<false, 0>("general, 2 channels/thread, 0 poly states", A, in, out);
  run2<false, 2>("general, 2 channels/thread, 2 poly states", A, in, out);
  run2<false, 3>("general, 2 channels/thread, 3 poly states", A, in, out);
  run2<false, 4>("general, 2 channels/thread, 4 poly states", A, in, out);
  run2<false, 6>("general, 2 channels/thread, 6 poly states", A, in, out);
  printf("config-2 needs 2000 warps x 5217 steps = %.2f M warp-steps per launch\n", 2000 * 5217 / 1e6);
  return cudaDeviceSynchronize() != cudaSuccess;
}
