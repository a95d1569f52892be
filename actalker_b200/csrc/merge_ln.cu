// Direction merge + branch sum + LayerNorm (C-ABI entry actk_merge_layernorm_fwd).
//
// Reference (src/models/base/mamba_layer.py): :1542-1547 y = y_fwd + flip(y_bwd) per branch; :1970/:1981 the
// result overwrites the selected rows of the in_proj output, other rows keep in_proj's value; :1983 xz2 + xz1;
// :1984 out_norm.  One warp per latent-token row; a row is read once with 128-bit loads, kept in registers for
// the two-pass mean/variance, and written once.  Every intermediate the reference materialises as a `dtype`
// tensor is rounded at the same point here.
#include "common.cuh"

namespace actk {

struct MergeParams {
  actk_merge_ln_args a;
};

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&v)[8]) {
  if (sizeof(T) == 4) {
    float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 w = *reinterpret_cast<const uint4 *>(p);
    const T *e = reinterpret_cast<const T *>(&w);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = IO<T>::ld(e + i);
  }
}
// raw 8-element vector (16 bytes of 16-bit data, 32 bytes of fp32) held in registers between load and use
template <typename T>
struct Raw8 {
  uint4 w[sizeof(T) == 4 ? 2 : 1];
};
template <typename T>
__device__ __forceinline__ Raw8<T> ldraw8(const T *p) {
  Raw8<T> r;
  r.w[0] = __ldcs(reinterpret_cast<const uint4 *>(p));
  if (sizeof(T) == 4) r.w[sizeof(T) == 4 ? 1 : 0] = __ldcs(reinterpret_cast<const uint4 *>(p) + 1);
  return r;
}
template <typename T>
__device__ __forceinline__ void unpack8(const Raw8<T> &r, float (&v)[8]) {
  if (sizeof(T) == 4) {
    const float *f = reinterpret_cast<const float *>(&r.w[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f[i];
  } else {
    const T *e = reinterpret_cast<const T *>(&r.w[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = IO<T>::ld(e + i);
  }
}
// Packed 16-bit add / multiply with ONE rounding to nearest-even per element (add.rn.bf16x2 / .f16x2, SASS HADD2 / HMUL2):
// exactly what the reference's `a + b` / `a * w` on two 16-bit tensors produce (the fp32 sum or product of two 8- or
// 11-bit significands is exact, then rounded once) — at a quarter of the instructions of widen, add, round, re-widen.
template <typename T> __device__ __forceinline__ uint32_t add2_16(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t add2_16<__nv_bfloat16>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <> __device__ __forceinline__ uint32_t add2_16<__half>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <> __device__ __forceinline__ uint32_t add2_16<float>(uint32_t a, uint32_t) { return a; }   // never used
template <typename T> __device__ __forceinline__ uint32_t mul2_16(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t mul2_16<__nv_bfloat16>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <> __device__ __forceinline__ uint32_t mul2_16<__half>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <> __device__ __forceinline__ uint32_t mul2_16<float>(uint32_t a, uint32_t) { return a; }   // never used
template <typename T>
__device__ __forceinline__ uint4 add8_16(uint4 a, uint4 b) {
  return make_uint4(add2_16<T>(a.x, b.x), add2_16<T>(a.y, b.y), add2_16<T>(a.z, b.z), add2_16<T>(a.w, b.w));
}
template <typename T>
__device__ __forceinline__ uint4 mul8_16(uint4 a, uint32_t w2) {
  return make_uint4(mul2_16<T>(a.x, w2), mul2_16<T>(a.y, w2), mul2_16<T>(a.z, w2), mul2_16<T>(a.w, w2));
}

// One vector of a row through the merge: per branch the direction sum (selected rows) or the in_proj value, the optional
// row weight, then the branch sum — every step rounded where the reference holds a `dtype` tensor.
template <typename T>
__device__ __forceinline__ void merge_vec(const Raw8<T> (&raw)[2][2], const bool (&sel)[2], const actk_merge_ln_args &a, int l,
                                          float (&acc)[8]) {
  if constexpr (sizeof(T) == 2) {
    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      if (br >= a.n_branches) break;
      uint4 t = raw[br][0].w[0];
      if (sel[br]) {
        t = add8_16<T>(t, raw[br][1].w[0]);
        if (a.row_weight[br]) {   // multiplicative region blend of SS2D_cond_v8 / v9 (mamba_layer.py:1777-1797)
          const uint32_t w = *reinterpret_cast<const uint16_t *>((const T *)a.row_weight[br] + l);
          t = mul8_16<T>(t, w | (w << 16));
        }
      }
      m = br == 0 ? t : add8_16<T>(t, m);
    }
    const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if constexpr (IO<T>::is_bf16) {
        acc[2 * j] = __uint_as_float(mw[j] << 16);
        acc[2 * j + 1] = __uint_as_float(mw[j] & 0xffff0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&mw[j]));
        acc[2 * j] = f.x;
        acc[2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      if (br >= a.n_branches) break;
      float t[8];
      unpack8<T>(raw[br][0], t);
      if (sel[br]) {
        float g[8];
        unpack8<T>(raw[br][1], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = IO<T>::rnd(t[e] + g[e]);
        if (a.row_weight[br]) {
          const float w = IO<T>::ld((const T *)a.row_weight[br] + l);
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = IO<T>::rnd(t[e] * w);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = br == 0 ? t[e] : IO<T>::rnd(t[e] + acc[e]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void store8(T *p, const float (&v)[8]) {
  if (sizeof(T) == 4) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 w;
    T *e = reinterpret_cast<T *>(&w);
#pragma unroll
    for (int i = 0; i < 8; ++i) IO<T>::st(e + i, v[i]);
    *reinterpret_cast<uint4 *>(p) = w;
  }
}

// 6 CTAs per SM (<= 80 registers) for the common VPL = 4: one register more costs a whole CTA of loads in flight
// (measured: 81 registers -> 5 CTAs -> 176 us instead of 160 us at config 2).
#ifndef ACTK_MERGE_MINB
#define ACTK_MERGE_MINB 6
#endif
template <typename T, int VPL>
__global__ void __launch_bounds__(128, (VPL <= 4 && sizeof(T) == 2) ? ACTK_MERGE_MINB : 1) merge_ln_kernel(const __grid_constant__ MergeParams P) {
  const actk_merge_ln_args &a = P.a;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long rows = (long long)a.Bp * a.L;
  if (row >= rows) return;
  const int l = (int)(row % a.L);
  const int D = a.D, nvec = D >> 3;
  const size_t off = (size_t)row * D;
  const size_t dir1 = (size_t)rows * D;   // ydir[1] - ydir[0]

  // The row flags are uniform over the warp, so every load address is known up front: all 16-byte loads of a chunk
  // of up to 4 vectors per lane (x 2 branches x 2 directions) are issued before any arithmetic — the kernel is a
  // pure HBM stream and was latency-bound with two loads in flight per lane (186 -> see DESIGN.md us at config 2).
  bool sel[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) sel[br] = br < a.n_branches && a.selected[br][l] != 0;
  constexpr int CH = VPL < 4 ? VPL : 4;
  float x[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i0 = 0; i0 < VPL; i0 += CH) {
    Raw8<T> raw[CH][2][2];
#pragma unroll
    for (int ii = 0; ii < CH; ++ii) {
      const int v = lane + 32 * (i0 + ii);
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        if (v < nvec && br < a.n_branches) {
          const T *p = (sel[br] ? (const T *)a.ydir[br] : (const T *)a.xz[br]) + off + 8 * v;
          raw[ii][br][0] = ldraw8(p);
          if (sel[br]) raw[ii][br][1] = ldraw8(p + dir1);
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < CH; ++ii) {
      const int i = i0 + ii, v = lane + 32 * i;
      if (v < nvec) {
        float acc[8];
        merge_vec<T>(raw[ii], sel, a, l, acc);
#pragma unroll
        for (int e = 0; e < 8; ++e) { x[i][e] = acc[e]; sum += acc[e]; }
      }
    }
  }
  if (!a.layernorm) {   // channel-sharded path: emit the merged sums, LayerNorm follows the all-gather
    if (a.n_peers > 0) {
      // fused push all-gather: this rank's slice goes straight into every rank's gather buffer over NVLink (16-byte
      // peer stores), in the (part, row, slice) layout gathered_ln_kernel reads — no separate collective launch
      const size_t poff = ((size_t)a.my_part * rows + row) * D;
      for (int p = 0; p < a.n_peers; ++p) {
        T *dst = (T *)a.peer_out[p] + poff;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (lane + 32 * i < nvec) store8(dst + 8 * (lane + 32 * i), x[i]);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) store8((T *)a.out + off + 8 * (lane + 32 * i), x[i]);
    return;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec)
#pragma unroll
      for (int e = 0; e < 8; ++e) { float d = x[i][e] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / D + a.eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float g[8], bt[8], o[8];
      load8((const T *)a.gamma + 8 * v, g);
      load8((const T *)a.beta + 8 * v, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((x[i][e] - mean) * rstd, g[e], bt[e]);
      store8((T *)a.out + off + 8 * v, o);
    }
  }
}

// merge_ln_kernel with WPR = 2 or 4 warps sharing a row: the UNet's wide layers (D = 1280 / 2560) would otherwise hold 5 / 10 vectors per
// lane and tensor in registers (234 registers, two CTAs per SM: 1.9 TB/s at D = 2560, profiles/r02_merge_ln_wide_ncu.txt);
// with the row split every width runs the 3-vector schedule of D = 640 and the statistics cross the warps through
// shared memory.  Lane `lane` of part `part` owns vectors lane + 32 * (part + WPR * i).  (A separate kernel: folding
// WPR = 1 into this body moved the D = 640 instantiation across the 80-register line of its 6-CTA launch bound.)
template <typename T, int VPL, int WPR>
__global__ void __launch_bounds__(128, (VPL <= 4 && sizeof(T) == 2) ? ACTK_MERGE_MINB - 1 : 1) merge_ln_split_kernel(const __grid_constant__ MergeParams P) {
  const actk_merge_ln_args &a = P.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = WPR == 1 ? 0 : warp % WPR;
  const long long rows = (long long)a.Bp * a.L;
  const long long row_raw = WPR == 1 ? (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)
                                     : (long long)blockIdx.x * (4 / WPR) + warp / WPR;
  if (WPR == 1 && row_raw >= rows) return;
  // WPR > 1: partner warps of a row must reach the barriers, so a row past the end recomputes the last one and stores nothing
  const bool valid = WPR == 1 ? true : row_raw < rows;
  const long long row = (WPR == 1 || valid) ? row_raw : rows - 1;
  const int l = (int)(row % a.L);
  const int D = a.D, nvec = D >> 3;
  const size_t off = (size_t)row * D;
  const size_t dir1 = (size_t)rows * D;   // ydir[1] - ydir[0]

  // The row flags are uniform over the warp, so every load address is known up front: all 16-byte loads of a chunk
  // of up to 4 vectors per lane (x 2 branches x 2 directions) are issued before any arithmetic — the kernel is a
  // pure HBM stream and was latency-bound with two loads in flight per lane (186 -> see DESIGN.md us at config 2).
  bool sel[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) sel[br] = br < a.n_branches && a.selected[br][l] != 0;
  constexpr int CH = VPL < 4 ? VPL : 4;
  float x[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i0 = 0; i0 < VPL; i0 += CH) {
    Raw8<T> raw[CH][2][2];
#pragma unroll
    for (int ii = 0; ii < CH; ++ii) {
      const int v = lane + 32 * (part + WPR * (i0 + ii));
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        if (v < nvec && br < a.n_branches) {
          const T *p = (sel[br] ? (const T *)a.ydir[br] : (const T *)a.xz[br]) + off + 8 * v;
          raw[ii][br][0] = ldraw8(p);
          if (sel[br]) raw[ii][br][1] = ldraw8(p + dir1);
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < CH; ++ii) {
      const int i = i0 + ii, v = lane + 32 * (part + WPR * i);
      if (v < nvec) {
        float acc[8];
        merge_vec<T>(raw[ii], sel, a, l, acc);
#pragma unroll
        for (int e = 0; e < 8; ++e) { x[i][e] = acc[e]; sum += acc[e]; }
      }
    }
  }
  if (!a.layernorm) {   // channel-sharded path: emit the merged sums, LayerNorm follows the all-gather
    if (a.n_peers > 0) {
      // fused push all-gather: this rank's slice goes straight into every rank's gather buffer over NVLink (16-byte
      // peer stores), in the (part, row, slice) layout gathered_ln_kernel reads — no separate collective launch
      const size_t poff = ((size_t)a.my_part * rows + row) * D;
      if (!valid) return;
      for (int p = 0; p < a.n_peers; ++p) {
        T *dst = (T *)a.peer_out[p] + poff;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * (part + WPR * i);
          if (v < nvec) store8(dst + 8 * v, x[i]);
        }
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * (part + WPR * i);
      if (v < nvec && valid) store8((T *)a.out + off + 8 * v, x[i]);
    }
    return;
  }
  __shared__ float red_sum[WPR > 1 ? 4 : 1], red_sq[WPR > 1 ? 4 : 1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (WPR > 1) {     // the row's parts meet in shared memory
    if (lane == 0) red_sum[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int j = 0; j < WPR; ++j) sum += red_sum[warp / WPR * WPR + j];
  }
  const float mean = sum / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * (part + WPR * i) < nvec)
#pragma unroll
      for (int e = 0; e < 8; ++e) { float d = x[i][e] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (WPR > 1) {
    if (lane == 0) red_sq[warp] = sq;
    __syncthreads();
    sq = 0.f;
#pragma unroll
    for (int j = 0; j < WPR; ++j) sq += red_sq[warp / WPR * WPR + j];
  }
  const float rstd = rsqrtf(sq / D + a.eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * (part + WPR * i);
    if (v < nvec && valid) {
      float g[8], bt[8], o[8];
      load8((const T *)a.gamma + 8 * v, g);
      load8((const T *)a.beta + 8 * v, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((x[i][e] - mean) * rstd, g[e], bt[e]);
      store8((T *)a.out + off + 8 * v, o);
    }
  }
}

// LayerNorm over channel slices gathered from P ranks: in (P, rows, Ds) -> out (rows, P*Ds).
// One warp per row; slice p of a row sits at in + (p*rows + row)*Ds (the layout NCCL all-gather produces).
template <typename T, int VPL>
__global__ void __launch_bounds__(128) gathered_ln_kernel(const T *__restrict__ in, const T *__restrict__ gamma,
                                                          const T *__restrict__ beta, T *__restrict__ out, int parts,
                                                          long long rows, int Ds, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int D = parts * Ds, nvec = D >> 3;
  float x[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      const int ch = 8 * v, p = ch / Ds;
      load8(in + ((size_t)p * rows + row) * Ds + (ch - p * Ds), x[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += x[i][e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec)
#pragma unroll
      for (int e = 0; e < 8; ++e) { float d = x[i][e] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / D + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float g[8], bt[8], o[8];
      load8(gamma + 8 * v, g);
      load8(beta + 8 * v, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((x[i][e] - mean) * rstd, g[e], bt[e]);
      store8(out + (size_t)row * D + 8 * v, o);
    }
  }
}

template <typename T>
int launch_gathered(const void *in, const void *gamma, const void *beta, void *out, int parts, long long rows,
                           int Ds, float eps, cudaStream_t stream) {
  const int vpl = ((parts * Ds >> 3) + 31) / 32;
  const unsigned grid = (unsigned)((rows + 3) / 4);
#define ACTK_GLN(V) gathered_ln_kernel<T, V><<<grid, 128, 0, stream>>>((const T *)in, (const T *)gamma, (const T *)beta, (T *)out, parts, rows, Ds, eps)
  if (vpl <= 4) ACTK_GLN(4);
  else if (vpl <= 8) ACTK_GLN(8);
  else if (vpl <= 16) ACTK_GLN(16);
  else ACTK_GLN(32);
#undef ACTK_GLN
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

template <typename T>
int launch_merge(const actk_merge_ln_args *a, cudaStream_t stream) {
  MergeParams P;
  P.a = *a;
  const int nvec = a->D >> 3;
  const long long rows = (long long)a->Bp * a->L;
  // warps per row: as few as keep <= 4 vectors per lane (D <= 1024: 1, <= 2048: 2, else 4)
  const int wpr = nvec <= 128 ? 1 : (nvec <= 256 ? 2 : 4);
  const int vpl = (nvec + 32 * wpr - 1) / (32 * wpr);
  const unsigned grid = (unsigned)((rows * wpr + 3) / 4);
  if (wpr == 1) merge_ln_kernel<T, 4><<<grid, 128, 0, stream>>>(P);
  else if (wpr == 2) merge_ln_split_kernel<T, 4, 2><<<grid, 128, 0, stream>>>(P);
  else if (vpl <= 4) merge_ln_split_kernel<T, 4, 4><<<grid, 128, 0, stream>>>(P);
  else merge_ln_split_kernel<T, 8, 4><<<grid, 128, 0, stream>>>(P);   // D up to 8192
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

// Compiled once per I/O dtype with -DACTK_TU_DTYPE=0/1/2 (kernel instantiations) and once without it (C-ABI entries).
#define ACTK_LN_INST(KW, T)                                                                                              \
  KW template int launch_gathered<T>(const void *, const void *, const void *, void *, int, long long, int, float, cudaStream_t); \
  KW template int launch_merge<T>(const actk_merge_ln_args *, cudaStream_t);
#if defined(ACTK_TU_DTYPE)
#if ACTK_TU_DTYPE == 0
ACTK_LN_INST(, float)
#elif ACTK_TU_DTYPE == 1
ACTK_LN_INST(, __half)
#else
ACTK_LN_INST(, __nv_bfloat16)
#endif
#else
ACTK_LN_INST(extern, float)
ACTK_LN_INST(extern, __half)
ACTK_LN_INST(extern, __nv_bfloat16)
#endif
#undef ACTK_LN_INST

}  // namespace actk

#ifndef ACTK_TU_DTYPE
using namespace actk;

extern "C" int actk_merge_layernorm_fwd(const actk_merge_ln_args *a, void *stream) {
  if (!a) ACTK_FAIL(ACTK_ERR_BAD_ARG, "actk_merge_layernorm_fwd: args is NULL");
  if (a->dtype < ACTK_F32 || a->dtype > ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "merge_ln: dtype=%d", a->dtype);
  if (a->n_branches < 1 || a->n_branches > 2) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln: n_branches=%d", a->n_branches);
  if (a->Bp <= 0 || a->L <= 0 || a->D <= 0 || a->D % 8 != 0 || a->D > 8192)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "merge_ln: Bp=%d L=%d D=%d (D must be a multiple of 8, <= 8192)", a->Bp, a->L, a->D);
  if (a->n_peers < 0 || a->n_peers > 8 || (a->n_peers > 0 && (a->layernorm || a->my_part < 0 || a->my_part >= a->n_peers)))
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln: n_peers=%d my_part=%d layernorm=%d (push gather needs layernorm == 0, my_part < n_peers <= 8)",
              a->n_peers, a->my_part, a->layernorm);
  for (int p = 0; p < a->n_peers; ++p)
    if (!a->peer_out[p] || (reinterpret_cast<uintptr_t>(a->peer_out[p]) & 15))
      ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln: peer_out[%d] is NULL or not aligned to 16 bytes", p);
  if ((!a->out && a->n_peers == 0) || (a->layernorm && (!a->gamma || !a->beta)))
    ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln: out (and gamma, beta when layernorm != 0) are required");
  for (int i = 0; i < a->n_branches; ++i) {
    if (!a->xz[i] || !a->ydir[i] || !a->selected[i]) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln: branch %d has a NULL pointer", i);
    if ((reinterpret_cast<uintptr_t>(a->xz[i]) | reinterpret_cast<uintptr_t>(a->ydir[i])) & 15)
      ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "merge_ln: branch %d pointer not aligned to 16 bytes", i);
  }
  if ((reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->gamma) | reinterpret_cast<uintptr_t>(a->beta)) & 15)
    ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "merge_ln: out/gamma/beta not aligned to 16 bytes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case ACTK_F32: return launch_merge<float>(a, st);
    case ACTK_F16: return launch_merge<__half>(a, st);
    default: return launch_merge<__nv_bfloat16>(a, st);
  }
}

extern "C" int actk_gathered_layernorm_fwd(const void *in, int parts, long long rows, int Ds, const void *gamma,
                                           const void *beta, float eps, void *out, int dtype, void *stream) {
  if (dtype < ACTK_F32 || dtype > ACTK_BF16) ACTK_FAIL(ACTK_ERR_BAD_DTYPE, "gathered_ln: dtype=%d", dtype);
  if (!in || !gamma || !beta || !out) ACTK_FAIL(ACTK_ERR_BAD_ARG, "gathered_ln: NULL pointer");
  if (parts <= 0 || rows <= 0 || Ds <= 0 || Ds % 8 != 0 || (long long)parts * Ds > 8192)
    ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "gathered_ln: parts=%d rows=%lld Ds=%d (Ds %% 8 == 0, parts*Ds <= 8192)", parts, rows, Ds);
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
       reinterpret_cast<uintptr_t>(beta)) & 15)
    ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "gathered_ln: pointer not aligned to 16 bytes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case ACTK_F32: return launch_gathered<float>(in, gamma, beta, out, parts, rows, Ds, eps, st);
    case ACTK_F16: return launch_gathered<__half>(in, gamma, beta, out, parts, rows, Ds, eps, st);
    default: return launch_gathered<__nv_bfloat16>(in, gamma, beta, out, parts, rows, Ds, eps, st);
  }
}
#endif  // !ACTK_TU_DTYPE
