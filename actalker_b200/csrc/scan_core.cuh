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
// batch x branch x direction x channel axes); long sequences are cut into chunks whose carries are
// resolved by scan_carry (two-level scan), see masked_scan.cu.
//
// POWER_A: A[d][n] == (n+1)*A[d][0] (S4D-real init) -> exp(dt*A_n) = r^(n+1): 2 MUFU + a multiply tree
// instead of 16 MUFU per step.  The general path issues one ex2 per state.
#pragma once
#include "common.cuh"

namespace actk {

template <bool POWER_A>
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
      uint64_t d2 = pk(dt, dt);
#pragma unroll
      for (int j = 0; j < kN / 2; ++j) {
        float lo, hi;
        upk(mul2(d2, a2[j]), lo, hi);
        p[j] = pk(ex2(lo), ex2(hi));
      }
    }
  }

  // One time step. `bc` points at 32 fp32 in shared memory: B[0..15] then C[0..15].
  // Returns sum_n C_n h_n + D*u (fp32, unrounded).
  template <bool SOFTPLUS>
  __device__ __forceinline__ float step(float u, float delta_raw, const float *__restrict__ bc) {
    float dt = delta_raw + bias;
    if (SOFTPLUS) dt = softplus20(dt);
    uint64_t p[kN / 2];
    decay(dt, p);
    float x = dt * u;
    uint64_t x2 = pk(x, x);
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
    return fmaf(dskip, u, y0 + y1);
  }

  // State-only step for the chunk-summary pass of the two-level scan (no C, no y).
  template <bool SOFTPLUS>
  __device__ __forceinline__ float step_state(float u, float delta_raw, const float *__restrict__ bc) {
    float dt = delta_raw + bias;
    if (SOFTPLUS) dt = softplus20(dt);
    uint64_t p[kN / 2];
    decay(dt, p);
    float x = dt * u;
    uint64_t x2 = pk(x, x);
    const ulonglong2 *bc2 = reinterpret_cast<const ulonglong2 *>(bc);
#pragma unroll
    for (int q = 0; q < kN / 4; ++q) {
      ulonglong2 Bq = bc2[q];
      h[2 * q] = fma2(p[2 * q], h[2 * q], mul2(x2, Bq.x));
      h[2 * q + 1] = fma2(p[2 * q + 1], h[2 * q + 1], mul2(x2, Bq.y));
    }
    return dt;
  }
};

}  // namespace actk
