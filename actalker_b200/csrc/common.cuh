// Shared device helpers for the sm_100a kernels: dtype traits, packed fp32x2 math, MUFU wrappers,
// mbarrier + bulk-async-copy (TMA) PTX.  No torch, no CUTLASS: plain CUDA + inline PTX.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/actalker_b200.h"

namespace actk {

constexpr int kN = ACTK_DSTATE;  // 16 states per channel
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ----------------------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
#define ACTK_FAIL(code, ...)  \
  do {                        \
    set_error(__VA_ARGS__);   \
    return (code);            \
  } while (0)
#define ACTK_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) ACTK_FAIL(ACTK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// ----------------------------------------------------------------------------- dtype traits
template <typename T> struct IO;
template <> struct IO<float> {
  static constexpr bool is_bf16 = false;
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float ld(const float *p) { return *p; }
  static __device__ __forceinline__ float f(float v) { return v; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
  static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
  // two adjacent elements (8-byte aligned)
  static __device__ __forceinline__ void ld2(const float *p, float &a, float &b) {
    float2 v = *reinterpret_cast<const float2 *>(p);
    a = v.x; b = v.y;
  }
  static __device__ __forceinline__ void st2(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
  }
};
template <> struct IO<__half> {
  static constexpr bool is_bf16 = false;
  static __device__ __forceinline__ __half zero() { return __float2half_rn(0.f); }
  static __device__ __forceinline__ float ld(const __half *p) { return __half2float(*p); }
  static __device__ __forceinline__ float f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
  static __device__ __forceinline__ void st(__half *p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ void ld2(const __half *p, float &a, float &b) {
    float2 v = __half22float2(*reinterpret_cast<const __half2 *>(p));
    a = v.x; b = v.y;
  }
  static __device__ __forceinline__ void st2(__half *p, float a, float b) {
    *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(a, b);
  }
};
template <> struct IO<__nv_bfloat16> {
  static constexpr bool is_bf16 = true;
  static __device__ __forceinline__ __nv_bfloat16 zero() { return __float2bfloat16_rn(0.f); }
  static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) {
    return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t *>(p)) << 16);
  }
  static __device__ __forceinline__ float f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void ld2(const __nv_bfloat16 *p, float &a, float &b) {
    const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
  }
  static __device__ __forceinline__ void st2(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
  }
};

// ----------------------------------------------------------------------------- MUFU wrappers
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// torch.nn.functional.softplus(x) with beta=1, threshold=20 (the rule selective_scan_fn applies when
// delta_softplus=True): x > 20 ? x : log1p(exp(x)).  For small e = exp(x) the 1+e rounding of a plain
// log(1+e) would cost relative accuracy, so the three-term series takes over below 2^-6.
__device__ __forceinline__ float softplus20(float x) {
  float e = ex2(x * kLog2e);
  float big = lg2(1.0f + e) * kLn2;
  float small = e * (1.0f + e * (-0.5f + e * 0.33333334f));
  float r = e < 0.015625f ? small : big;
  return x > 20.0f ? x : r;
}
// The same rule without the series: log1p(e) as ln2 * lg2(1 + e).  Rounding 1 + e costs up to 2^-24 ABSOLUTE on dt
// (6e-8), i.e. 6e-5 relative at dt = 1e-3 — invisible next to the 2^-9 (bf16) / 2^-12 (fp16) relative rounding the raw
// delta already carries with 16-bit I/O, which is the only place it is used (fp32 I/O keeps softplus20): five issue
// slots fewer per channel-step in an issue-bound loop.
__device__ __forceinline__ float softplus20_io16(float x) {
  float e = ex2(x * kLog2e);
  float r = lg2(1.0f + e) * kLn2;
  return x > 20.0f ? x : r;
}
// One MUFU instead of two: softplus(x) = max(x, 0) + log1p(e), e = exp(-|x|) in (0, 1], with log1p(e) = e * q(e), q a
// degree-5 minimax polynomial of log1p(e)/e on [0, 1] constrained to q(0) = 1 (max relative error 9.6e-6 on dt over the
// whole range — six times tighter than softplus20_io16 at small dt — and x > 20 returns x exactly: log1p(e^-20) is below
// half an ulp of 20).  The scan is bound by the MUFU pipe (16 per channel-step, ~80 % busy) while the FMA pipe has slack:
// this trades the lg2 for five FFMA + one FMUL.  16-bit I/O only, like softplus20_io16.
__device__ __forceinline__ float softplus20_io16_poly(float x) {
  const float e = ex2(-fabsf(x) * kLog2e);
  float q = -0.02473430335521698f;
  q = fmaf(q, e, 0.10359206795692444f);
  q = fmaf(q, e, -0.21240372955799103f);
  q = fmaf(q, e, 0.3262259364128113f);
  q = fmaf(q, e, -0.49953946471214294f);
  q = fmaf(q, e, 1.0f);
  return fmaf(q, e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + ex2(-x * kLog2e)); }

// ----------------------------------------------------------------------------- packed fp32x2 (sm_100+)
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(uint64_t v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^t for two values t <= 0 on the FMA pipe instead of the MUFU unit (16 ex2/clk/SM is the wall of the
// general-A scan): Cody-Waite split t = n + f with the 1.5*2^23 magic add, degree-5 polynomial for 2^f on
// [-0.5, 0.5] constrained to p(0) = 1 (max rel. error 1.9e-7, the same class as ex2.approx), exponent inserted
// with one integer shift-add.  8 packed FMA-pipe ops + 4 ALU ops per pair.
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t t2) {
  float tl, th;
  upk(t2, tl, th);
  const uint64_t t = pk(fmaxf(tl, -126.0f), fmaxf(th, -126.0f));
  const uint64_t s = add2(t, pk(12582912.0f, 12582912.0f));           // low mantissa bits now hold round(t)
  const uint64_t nf = add2(s, pk(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(nf, pk(-1.0f, -1.0f), t);
  uint64_t p = pk(0.001326472731307149f, 0.001326472731307149f);
  p = fma2(p, f, pk(0.009671512991189957f, 0.009671512991189957f));
  p = fma2(p, f, pk(0.05550733581185341f, 0.05550733581185341f));
  p = fma2(p, f, pk(0.24022242426872253f, 0.24022242426872253f));
  p = fma2(p, f, pk(0.6931470036506653f, 0.6931470036506653f));
  p = fma2(p, f, pk(1.0f, 1.0f));
  float pl, ph, sl, sh;
  upk(p, pl, ph);
  upk(s, sl, sh);
  return pk(__int_as_float(__float_as_int(pl) + (__float_as_int(sl) << 23)),
            __int_as_float(__float_as_int(ph) + (__float_as_int(sh) << 23)));
}

// Scalar form of ex2_poly2 for ONE value (same split, same polynomial): 11 issue slots instead of one MUFU.  With the
// MUFU pipe at 8 cycles per warp instruction and the issue port at ~1.1 per scalar / 2.3 per packed instruction, moving an
// odd number of the 16 per-step exponentials lets the two limits meet (scan_core.cuh, ACTK_POLY_STATES).
__device__ __forceinline__ float ex2_poly1(float t) {
  t = fmaxf(t, -126.0f);
  const float s = t + 12582912.0f;
  const float f = t - (s - 12582912.0f);
  float p = 0.001326472731307149f;
  p = fmaf(p, f, 0.009671512991189957f);
  p = fmaf(p, f, 0.05550733581185341f);
  p = fmaf(p, f, 0.24022242426872253f);
  p = fmaf(p, f, 0.6931470036506653f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(s) << 23));
}

// ----------------------------------------------------------------------------- mbarrier / TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// The wait sleeps in hardware until the phase completes or the hint expires; the default hint is so short
// that an idle producer warp re-polls every ~45 cycles and takes issue slots from the compute warps.
#ifndef ACTK_SUSPEND_NS
#define ACTK_SUSPEND_NS 20000
#endif
constexpr uint32_t kSuspendHintNs = ACTK_SUSPEND_NS;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(kSuspendHintNs)
      : "memory");
}
// global -> shared bulk async copy (TMA unit, SASS UBLKCP); size % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 3-D tiled TMA load (SASS UTMALDG): box described by the tensor map, coordinates innermost first.
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst_smem, const void *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// 3-D tiled TMA store (SASS UTMASTG), tracked by the bulk async-group of the issuing thread.
__device__ __forceinline__ void tma_store_3d(const void *tmap, int c0, int c1, int c2, const void *src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap),
               "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src_smem))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmap_prefetch(const void *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 16-byte Ampere-style async copy (SASS LDGSTS) for ragged gather tiles, completing on an mbarrier.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace actk
