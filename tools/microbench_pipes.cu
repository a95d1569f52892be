// Microbenchmark of the raw instruction rates the scan depends on (B200, sm_100a): scalar FFMA, packed FFMA2,
// FMUL2 with a broadcast operand, MUFU.EX2, and broadcast LDS.128 — each as 16 independent chains per thread so
// that only pipe throughput (not latency) is measured.  Prints SMSP-cycles per warp-instruction.
#include <cstdio>

#include "../actalker_b200/csrc/common.cuh"
namespace actk {
void set_error(const char *, ...) {}
}
using namespace actk;

template <int OP>
__global__ void __launch_bounds__(128) rate_kernel(float *out, const float *in, int iters) {
  __shared__ alignas(16) float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  float f[16];
  uint64_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { f[i] = in[i] + threadIdx.x; v[i] = pk(f[i], f[i] + 1.f); }
  const float a = in[17], b = in[18];
  const uint64_t a2 = pk(a, a), b2 = pk(b, b + 1e-3f);
  float acc = 0.f;
  int n[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) n[i] = threadIdx.x * (i + 1);
  const int kx = __float_as_int(a) | 5, ky = __float_as_int(b) | 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) f[i] = fmaf(f[i], a, b);
      if (OP == 1) v[i] = fma2(v[i], a2, b2);
      if (OP == 2) v[i] = mul2(v[i], a2);
      if (OP == 3) f[i] = ex2(f[i]);
      if (OP == 4) {
        float4 x = *reinterpret_cast<const float4 *>(&sm[((it * 16 + i) * 4) & 1020]);
        acc += x.x + x.w;   // 2 FADD per LDS.128 (needed to keep the load alive)
      }
      // mixes: does a packed op hold the dispatch port for both of its pipe cycles?
      if (OP == 6) { v[i] = fma2(v[i], a2, b2); n[i] = (n[i] ^ kx) + ky; }                   // FFMA2 + LOP3 + IADD
      if (OP == 7) { v[i] = fma2(v[i], a2, b2); if ((i & 3) == 0) f[i] = ex2(f[i]); }          // 4 FFMA2 : 1 MUFU
      if (OP == 8) { f[i] = fmaf(f[i], a, b); f[(i + 8) & 15] = fmaf(f[(i + 8) & 15], a, b); n[i] = (n[i] ^ kx) + ky; }
      if (OP == 9) { f[i] = ex2(f[i]); n[i] = (n[i] ^ kx) + ky; n[(i + 8) & 15] = (n[(i + 8) & 15] ^ ky) + kx; }  // MUFU + 4 ALU
      if (OP == 10) {  // 3 FFMA2 : 1 LDS.128 (apply ratio), no FADD on the loaded data except one xor-reduce
        v[i] = fma2(v[i], a2, b2);
        if ((i % 3) == 0) {
          float4 x = *reinterpret_cast<const float4 *>(&sm[((it * 16 + i) * 4) & 1020]);
          n[i] ^= __float_as_int(x.x) ^ __float_as_int(x.w);
        }
      }
      if (OP == 11) { v[i] = fma2(v[i], a2, b2); if ((i & 1) == 0) f[i] = ex2(f[i]); n[i] = (n[i] ^ kx) + ky; }  // 2 FFMA2 : 1 MUFU : 4 ALU
      if (OP == 12) { n[i] = (n[i] ^ kx) + ky; }                                               // LOP3 + IADD alone: the ALU rate
      if (OP == 13) {  // ex2.approx.ftz.f16x2: two half-precision exponentials per MUFU instruction?
        uint32_t x = (uint32_t)n[i];
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x));
        n[i] = (int)x;
      }
      if (OP == 14) {  // ex2.approx.ftz.bf16x2
        uint32_t x = (uint32_t)n[i];
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x));
        n[i] = (int)x;
      }
      if (OP == 15) { v[i] = fma2(v[i], a2, b2); n[i] = __funnelshift_l(n[i], kx, 7); }        // FFMA2 + SHF: two pipes, one issue port
      if (OP == 16) { f[i] = fmaf(f[i], a, b); n[i] = __funnelshift_l(n[i], kx, 7); }          // FFMA + SHF
      if (OP == 5) {  // the apply() pattern: FMUL2 (bcast) + FFMA2 + FFMA2 per state pair, operands from registers
        uint64_t d = mul2(a2, v[(i + 1) & 15]);
        v[i] = fma2(b2, v[i], d);
        v[(i + 8) & 15] = fma2(a2, v[i], v[(i + 8) & 15]);
      }
    }
  }
  float r = acc;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float lo, hi; upk(v[i], lo, hi); r += f[i] + lo + hi + __int_as_float(n[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
static void run(const char *name, int inst_per_iter, float *out, const float *in) {
  const int iters = 2048;
  for (int warps_per_sm : {4, 16, 32}) {
    int ctas = 148 * warps_per_sm / 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    rate_kernel<OP><<<ctas, 128>>>(out, in, iters);
    cudaEventRecord(a);
    rate_kernel<OP><<<ctas, 128>>>(out, in, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double winst = (double)ctas * 4 * iters * inst_per_iter;
    printf("%-34s warps/SM %2d: %.2f SMSP-cycles per warp-instruction\n", name, warps_per_sm,
           ms * 1e-3 * 1.965e9 / (winst / (148 * 4)));
  }
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 148 * 8 * 128 * 4 * 4);
  cudaMemset(in, 0, 4096);
  run<0>("FFMA (scalar)", 16, out, in);
  run<1>("FFMA2 (packed)", 16, out, in);
  run<2>("FMUL2 (packed)", 16, out, in);
  run<3>("MUFU.EX2", 16, out, in);
  run<4>("LDS.128 broadcast (+2 FADD)", 16, out, in);
  run<5>("FMUL2+FFMA2+FFMA2 triple", 48, out, in);
  // per GROUP of instructions (cycles for the whole group): a sum of the parts means the dispatch port is shared
  run<6>("group{FFMA2,LOP3,IADD}", 16, out, in);
  run<7>("group{4 FFMA2,1 MUFU}", 4, out, in);
  run<8>("group{2 FFMA,LOP3,IADD}", 16, out, in);
  run<9>("group{MUFU,2 LOP3,2 IADD}", 16, out, in);
  run<10>("group{3 FFMA2,1 LDS.128,~2 LOP3}", 5, out, in);
  run<11>("group{2 FFMA2,1 MUFU,2 LOP3,2 IADD}", 8, out, in);
  run<12>("group{LOP3,IADD}", 16, out, in);
  run<13>("MUFU.EX2 f16x2 (2 exps/instr)", 16, out, in);
  run<14>("MUFU.EX2 bf16x2 (2 exps/instr)", 16, out, in);
  run<15>("group{FFMA2,SHF}", 16, out, in);
  run<16>("group{FFMA,SHF}", 16, out, in);
  return cudaDeviceSynchronize() != cudaSuccess;
}
