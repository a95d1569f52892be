// Register-resident selective-scan recurrence for ONE channel and all 16 states.
//
//   dt  = softplus(delta + dt_bias)                       (mamba_layer.py:1532-1538: delta_softplus=True)
//   h_n = exp(dt * A_n) * h_n + dt * B_n * u              
//   y   = sum_n C_n * h_n + D * u
//
// One thread owns one channel: the 16 states live in 8 packed fp32x2 registers, so the whole update is
// 8 x {mul.f32x2, fma.f32x2, fma.f32x2} with no cross-lane traffic, and y needs no shuffle reduction.
// B_n / C_n of the step are shared by every channel of the (batch, direction) and are read from shared
// memory as fp32 (broadcast LDS.128).  The scan over time is sequential per thread (parallelism comes from the
// batch x branch x direction x channel axes).
//
// The step is split in three stages so that callers can software-pipeline them across consecutive time steps
// (the only true loop-carried dependency is one FFMA2 per state pair):
//   prologue(u, delta)  -> dt (bias + softplus), x = dt*u                      [2 MUFU, serial chain ~150 cycles]
//   decay(dt)           -> exp(dt*A_n) for the 16 states, packed               [POWER_A: 2 MUFU + multiply tree;
//                                                                               general: 16 MUFU]
//   apply(p, x, u, bc)  -> h update, y                                          [24 packed FMA-pipe ops]
//
// POWER_A: A[d][n] == (n+1)*A[d][0] (S4D-real init) -> exp(dt*A_n) = r^(n+1).
#pragma once
#include "common.cuh"

namespace actk {

#ifndef ACTK_POLY_STATES
#define ACTK_POLY_STATES 2  // states per step whose exp runs on the FMA pipe instead of the MUFU unit (general-A path);
                            // round 1, whole pairs on B200 at config 2 with chain mode: 0 -> 1.53 ms, 2 -> 1.50 ms,
                            // 4 -> 1.59 ms, 6 -> 1.73 ms.  Odd counts take the scalar polynomial for one state.
#endif
constexpr int kPolyStates = ACTK_POLY_STATES;
constexpr int kPolyPairs = kPolyStates / 2;     // whole pairs: packed polynomial
constexpr bool kPolyOdd = (kPolyStates & 1) != 0;  // plus one state of the next pair: scalar polynomial

struct StepIn {
  float dt, x, u;
};

#ifndef ACTK_FAST_SOFTPLUS16
#define ACTK_FAST_SOFTPLUS16 1   // 16-bit I/O: softplus without the small-argument series (see softplus20_io16)
#endif

#ifndef ACTK_SOFTPLUS_POLY
#define ACTK_SOFTPLUS_POLY 0     // 16-bit I/O: log1p on the FMA pipe (softplus20_io16_poly): one MUFU per softplus instead of two
#endif

// IO16: the activations are 16-bit tensors (selects the softplus form; all arithmetic stays fp32)
template <bool POWER_A, bool IO16 = false>
struct ChannelScan {
  uint64_t h[kN / 2];
  uint64_t a2[POWER_A ? 1 : kN / 2];  // general: A[d][n]*log2e, packed pairs
  float a0;                           // A[d][0]*log2e
  float dskip, bias;

  __device__ __forceinline__ void init(const float *__restrict__ A_row, float D, float dt_bias) {
    a0 = A_row[0] * kLog2e;
    if (!POWER_A) {
#pragma unroll
      for (int j = 0; j < kN / 2; ++j) a2[j] = pk(A_row[2 * j] * kLog2e, A_row[2 * j + 1] * kLog2e);
    }
#pragma unroll
    for (int j = 0; j < kN / 2; ++j) h[j] = pk(0.f, 0.f);
    dskip = D;
    bias = dt_bias;
  }

  template <bool SOFTPLUS>
  __device__ __forceinline__ StepIn prologue(float u, float delta_raw) const {
    StepIn s;
    float dt = delta_raw + bias;
    if (SOFTPLUS) dt = (IO16 && ACTK_FAST_SOFTPLUS16) ? (ACTK_SOFTPLUS_POLY ? softplus20_io16_poly(dt) : softplus20_io16(dt))
                                                      : softplus20(dt);
    s.dt = dt;
    s.u = u;
    s.x = dt * u;
    return s;
  }

  // decay factors exp(dt*A_n) for the 16 states, packed
  __device__ __forceinline__ void decay(float dt, uint64_t (&p)[kN / 2]) const {
    if (POWER_A) {
      float t = dt * a0;
      float r = ex2(t), r8 = ex2(8.0f * t);
      float r2 = r * r, r4 = r2 * r2;
      uint64_t q2 = pk(r2, r2), q4 = pk(r4, r4), q8 = pk(r8, r8);
      p[0] = pk(r, r2);
      p[1] = mul2(p[0], q2);
      p[2] = mul2(p[0], q4);
      p[3] = mul2(p[1], q4);
#pragma unroll
      for (int j = 0; j < 4; ++j) p[4 + j] = mul2(p[j], q8);
    } else {
      // 16 exponentials per step: the MUFU unit (16/clk/SM) is the wall, the FMA pipe has slack, so the last
      // kPolyPairs state pairs take the polynomial route (see ex2_poly2) and the rest ex2.approx.
      uint64_t d2 = pk(dt, dt);
#pragma unroll
      for (int j = 0; j < kN / 2; ++j) {
        const uint64_t t2 = mul2(d2, a2[j]);
        if (j >= kN / 2 - kPolyPairs) {
          p[j] = ex2_poly2(t2);
        } else if (kPolyOdd && j == kN / 2 - kPolyPairs - 1) {
          float lo, hi;
          upk(t2, lo, hi);
          p[j] = pk(ex2(lo), ex2_poly1(hi));
        } else {
          float lo, hi;
          upk(t2, lo, hi);
          p[j] = pk(ex2(lo), ex2(hi));
        }
      }
    }
  }

  // State update + output of one step. `bc`: 32 fp32 in shared memory, B[0..15] then C[0..15].
  // Returns sum_n C_n h_n + D*u (fp32, unrounded).
  __device__ __forceinline__ float apply(const uint64_t (&p)[kN / 2], const StepIn &s, const float *__restrict__ bc) {
    uint64_t x2 = pk(s.x, s.x);
    const ulonglong2 *bc2 = reinterpret_cast<const ulonglong2 *>(bc);
    uint64_t ya = pk(0.f, 0.f), yb = pk(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kN / 4; ++q) {
      ulonglong2 Bq = bc2[q];
      ulonglong2 Cq = bc2[kN / 4 + q];
      h[2 * q] = fma2(p[2 * q], h[2 * q], mul2(x2, Bq.x));
      h[2 * q + 1] = fma2(p[2 * q + 1], h[2 * q + 1], mul2(x2, Bq.y));
      ya = fma2(Cq.x, h[2 * q], ya);
      yb = fma2(Cq.y, h[2 * q + 1], yb);
    }
    float y0, y1;
    upk(add2(ya, yb), y0, y1);
    return fmaf(dskip, s.u, y0 + y1);
  }

  // State-only update (no C, no y): the chunk-summary pass of the two-level scan.
  __device__ __forceinline__ void apply_state(const uint64_t (&p)[kN / 2], const StepIn &s, const float *__restrict__ bc) {
    uint64_t x2 = pk(s.x, s.x);
    const ulonglong2 *bc2 = reinterpret_cast<const ulonglong2 *>(bc);
#pragma unroll
    for (int q = 0; q < kN / 4; ++q) {
      ulonglong2 Bq = bc2[q];
      h[2 * q] = fma2(p[2 * q], h[2 * q], mul2(x2, Bq.x));
      h[2 * q + 1] = fma2(p[2 * q + 1], h[2 * q + 1], mul2(x2, Bq.y));
    }
  }

  // Pipelined state-only run (chunk-summary pass): returns the sum of dt over the NSTEP steps.
  template <int NSTEP, bool SOFTPLUS, typename LdU, typename LdD, typename Bc>
  __device__ __forceinline__ float run_state(LdU ld_u, LdD ld_d, Bc bc) {
    StepIn s[NSTEP];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) { s[i] = prologue<SOFTPLUS>(ld_u(i), ld_d(i)); sum += s[i].dt; }
    uint64_t p[2][kN / 2];
    decay(s[0].dt, p[0]);
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) {
      if (i + 1 < NSTEP) decay(s[i + 1].dt, p[(i + 1) & 1]);
      apply_state(p[i & 1], s[i], bc(i));
    }
    return sum;
  }

  // Unpipelined convenience form (operator-contract kernel, ragged tails).
  template <bool SOFTPLUS>
  __device__ __forceinline__ float step(float u, float delta_raw, const float *__restrict__ bc) {
    StepIn s = prologue<SOFTPLUS>(u, delta_raw);
    uint64_t p[kN / 2];
    decay(s.dt, p);
    return apply(p, s, bc);
  }

  // Software-pipelined run of NSTEP consecutive steps.  A warp issues in order, so latency is only hidden by
  // independent work that is adjacent in the instruction stream: all prologues first (independent softplus chains
  // that overlap each other), then decay(i+1) is issued ahead of apply(i) so the MUFU latency of the next step is
  // covered by the 24 packed FMA ops of the current one.  (Measured on B200, config 2: general-A 1.96 -> 1.85 ms;
  // deeper lookahead across groups cost registers and gained nothing, see DESIGN.md.)
  //   ld_u(i), ld_d(i) -> raw fp32 inputs of step i;  bc(i) -> B|C pointer;  out(i, y) consumes the result.
  template <int NSTEP, bool SOFTPLUS, typename LdU, typename LdD, typename Bc, typename Out>
  __device__ __forceinline__ void run(LdU ld_u, LdD ld_d, Bc bc, Out out) {
    StepIn s[NSTEP];
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) s[i] = prologue<SOFTPLUS>(ld_u(i), ld_d(i));
    uint64_t p[2][kN / 2];
    decay(s[0].dt, p[0]);
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) {
      if (i + 1 < NSTEP) decay(s[i + 1].dt, p[(i + 1) & 1]);
      out(i, apply(p[i & 1], s[i], bc(i)));
    }
  }
};

}  // namespace actk
