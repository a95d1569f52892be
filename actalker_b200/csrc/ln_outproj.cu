// Direction merge + branch sum + LayerNorm + out_proj in ONE kernel (C-ABI entry actk_merge_ln_outproj_fwd) —
// "LN + out_proj is a natural epilogue ... dense => tensor cores (tcgen05)".
//
// Reference (src/models/base/mamba_layer.py): :1542-1547 y = y_fwd + flip(y_bwd); :1970/:1981 scatter over the in_proj
// output; :1983 xz2 + xz1; :1984 out_norm; :1985 out_proj (nn.Linear(D, d_model, bias=False)).
// The unfused path writes the normalised (B'L, D) tensor and cuBLAS reads it back; here a 128-row tile is merged and
// normalised straight into shared memory as the tensor cores' A operand and multiplied by W_out (streamed from L2
// in 32-column slabs by TMA) with tcgen05.mma, fp32 accumulators in tensor memory, one rounding to the output dtype.
// Same rounding points as the unfused path (every tensor the reference materialises is rounded where it does).
//
//   one persistent CTA per SM, 256 threads, ~205 KB of shared memory:
//     phase A  8 warps x 4 rows at a time: 16-byte loads of the (up to) four scan outputs of a row, two-pass LayerNorm in
//              fp32 over 8 lanes, 16-byte stores into the A tile in the 128-byte-swizzled K-major layout (conflict-free)
//     phase B  one thread: TMA of W slabs (64-byte swizzle, double buffered) + 2 x tcgen05.mma (M=128, N=d_model/2, K=16)
//              per 16 columns; completion through tcgen05.commit on mbarriers
//     phase C  tcgen05.ld of the accumulator rows (lane = row), convert, 16-byte stores of the output rows
// Shapes: 16-bit I/O, D % 64 == 0 and D <= 640 (A tile in shared memory), d_model % 32 == 0 and <= 512 (TMEM columns).
// Other shapes use actk_merge_layernorm_fwd + a library GEMM (the caller decides; this entry returns UNSUPPORTED).
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace actk {

constexpr int kRows = 128;        // rows per tile == MMA M
constexpr int kLnThreads = 256;
constexpr int kWSlab = 32;        // W columns (K) per TMA slab: 64-byte rows, SWIZZLE_64B

struct LnOutParams {
  actk_merge_ln_args a;           // merge + LayerNorm inputs (a.out unused)
  void *out;                      // (rows, N) output
  int N;                          // d_model
};

__device__ __forceinline__ uint64_t sw_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

template <typename T>
__device__ __forceinline__ void unpack8v(const uint4 &w, float (&v)[8]) {
  const T *e = reinterpret_cast<const T *>(&w);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = IO<T>::ld(e + i);
}
template <typename T>
__device__ __forceinline__ uint4 pack8v(const float (&v)[8]) {
  uint4 w;
  T *e = reinterpret_cast<T *>(&w);
#pragma unroll
  for (int i = 0; i < 8; ++i) IO<T>::st(e + i, v[i]);
  return w;
}

// KB = D / 64 k-blocks of the A tile.
template <typename T, int KB>
__global__ void __launch_bounds__(kLnThreads, 1) ln_outproj_kernel(const __grid_constant__ LnOutParams P,
                                                                  const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms want 1024-byte alignment
  uint8_t *smem = smem_raw + (smem0 - smem_u32(smem_raw));
  constexpr int D = 64 * KB;
  constexpr uint32_t kABytes = (uint32_t)KB * kRows * 128;            // A tile: [k-block][row][128 B]
  const int N = P.N;
  const uint32_t w_bytes = (uint32_t)N * kWSlab * sizeof(T);          // one W slab: [n][64 B]
  const uint32_t a_addr = smem0, w_addr = smem0 + kABytes;            // W double buffer follows A (512-aligned)
  const uint32_t bar0 = w_addr + 2 * w_bytes;                          // full[2], empty[2], acc
  const uint32_t full_bar = bar0, empty_bar = bar0 + 16, acc_bar = bar0 + 32;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + (bar0 + 40 - smem0));

  const actk_merge_ln_args &a = P.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long rows = (long long)a.Bp * a.L;
  const int tiles = (int)((rows + kRows - 1) / kRows);
  const size_t dir1 = (size_t)rows * D;

  if (tid == 0) {
    mbar_init_a(full_bar, 1); mbar_init_a(full_bar + 8, 1);
    mbar_init_a(empty_bar, 1); mbar_init_a(empty_bar + 8, 1);
    mbar_init_a(acc_bar, 1);
    mbar_fence_init();
    tmap_prefetch(&wmap);
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  constexpr int kSlabs = D / kWSlab;                                  // W slabs per tile
  const uint32_t idesc = (1u << 4) | ((IO<T>::is_bf16 ? 1u : 0u) << 7) | ((IO<T>::is_bf16 ? 1u : 0u) << 10) |
                         ((uint32_t)(N / 2 >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t slab_count = 0;    // W slabs issued so far by thread 0 (ring position and phases)
  uint32_t tile_count = 0;

  // phase A geometry: lane group g = lane / 8 takes one of 4 rows, j = lane % 8 the 16-byte piece within a 128-byte row
  const int g = lane >> 3, j = lane & 7;

  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_count) {
    const long long row0 = (long long)tile * kRows;
    // ---- thread 0: the first two W slabs travel while the rows are normalised (they do not depend on the tile)
    if (tid == 0) {
      for (int c = 0; c < 2; ++c) {
        const uint32_t s = slab_count & 1, use = slab_count >> 1;
        if (use > 0) mbar_wait_a(empty_bar + 8 * s, (use - 1) & 1);
        mbar_expect_tx_a(full_bar + 8 * s, w_bytes);
        tma_load_2d(w_addr + s * w_bytes, &wmap, c * kWSlab, 0, full_bar + 8 * s);
        tma_load_2d(w_addr + s * w_bytes + (uint32_t)(N / 2) * kWSlab * sizeof(T), &wmap, c * kWSlab, N / 2, full_bar + 8 * s);
        ++slab_count;
      }
    }
    // ---- phase A: merge + LayerNorm into the A tile
#pragma unroll 1
    for (int pass = 0; pass < kRows / 32; ++pass) {
      const int rt = pass * 32 + warp * 4 + g;                         // row within the tile
      const long long row = row0 + rt;
      const bool valid = row < rows;
      const int l = valid ? (int)(row % a.L) : 0;
      const size_t off = (size_t)(valid ? row : 0) * D;
      bool sel[2];
#pragma unroll
      for (int br = 0; br < 2; ++br) sel[br] = br < a.n_branches && a.selected[br][l] != 0;
      float x[KB][8];
      float sum = 0.f;
#pragma unroll
      for (int i0 = 0; i0 < KB; i0 += 5) {
        uint4 raw[5][2][2];
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          const int v = (i0 + ii) * 8 + j;
#pragma unroll
          for (int br = 0; br < 2; ++br) {
            if (i0 + ii < KB && br < a.n_branches) {
              const T *p = (sel[br] ? (const T *)a.ydir[br] : (const T *)a.xz[br]) + off + 8 * v;
              raw[ii][br][0] = __ldcs(reinterpret_cast<const uint4 *>(p));
              if (sel[br]) raw[ii][br][1] = __ldcs(reinterpret_cast<const uint4 *>(p + dir1));
            }
          }
        }
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          const int i = i0 + ii;
          if (i < KB) {
            float acc[8];
#pragma unroll
            for (int br = 0; br < 2; ++br) {
              if (br >= a.n_branches) break;
              float t[8];
              unpack8v<T>(raw[ii][br][0], t);
              if (sel[br]) {
                float q[8];
                unpack8v<T>(raw[ii][br][1], q);
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = IO<T>::rnd(t[e] + q[e]);
                if (a.row_weight[br]) {
                  const float w = IO<T>::ld((const T *)a.row_weight[br] + l);
#pragma unroll
                  for (int e = 0; e < 8; ++e) t[e] = IO<T>::rnd(t[e] * w);
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] = br == 0 ? t[e] : IO<T>::rnd(t[e] + acc[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { x[i][e] = acc[e]; sum += acc[e]; }
          }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / D;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < KB; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = x[i][e] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / D + a.eps);
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        const int v = i * 8 + j;
        float gm[8], bt[8], o[8];
        unpack8v<T>(__ldg(reinterpret_cast<const uint4 *>((const T *)a.gamma + 8 * v)), gm);
        unpack8v<T>(__ldg(reinterpret_cast<const uint4 *>((const T *)a.beta + 8 * v)), bt);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = valid ? fmaf((x[i][e] - mean) * rstd, gm[e], bt[e]) : 0.f;
        // A tile, K-major, 128-byte swizzle: k-block i, row rt, 16-byte piece j at ((j ^ (rt & 7)) * 16)
        *reinterpret_cast<uint4 *>(smem + (size_t)i * (kRows * 128) + (size_t)rt * 128 + ((j ^ (rt & 7)) << 4)) = pack8v<T>(o);
      }
    }
    fence_proxy_async();                 // generic-proxy writes of the A tile -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- phase B: one thread streams W and issues the MMAs
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t first = slab_count - 2;       // ring position of this tile's slab 0
      for (int c = 0; c < kSlabs; ++c) {
        const uint32_t n = first + c, s = n & 1, use = n >> 1;
        mbar_wait_a(full_bar + 8 * s, use & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int ks = 0; ks < kWSlab / 16; ++ks) {
          const int kk = c * kWSlab + ks * 16;     // absolute k of this 16-wide step
          const uint64_t ad = sw_desc(a_addr + (uint32_t)(kk >> 6) * (kRows * 128) + (uint32_t)(kk & 63) * sizeof(T), 1024, 2);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t bd = sw_desc(w_addr + s * w_bytes + (uint32_t)h * (N / 2) * kWSlab * sizeof(T) + ks * 32, 512, 4);
            umma_ss(tmem + h * (N / 2), ad, bd, idesc, (c | ks) != 0);
          }
        }
        umma_commit_a(empty_bar + 8 * s);          // slab s may be overwritten once these MMAs have read it
        if (c + 2 < kSlabs) {                      // refill the slot with slab c + 2
          const uint32_t n2 = slab_count, s2 = n2 & 1, use2 = n2 >> 1;
          mbar_wait_a(empty_bar + 8 * s2, (use2 - 1) & 1);
          mbar_expect_tx_a(full_bar + 8 * s2, w_bytes);
          tma_load_2d(w_addr + s2 * w_bytes, &wmap, (c + 2) * kWSlab, 0, full_bar + 8 * s2);
          tma_load_2d(w_addr + s2 * w_bytes + (uint32_t)(N / 2) * kWSlab * sizeof(T), &wmap, (c + 2) * kWSlab, N / 2,
                      full_bar + 8 * s2);
          ++slab_count;
        }
      }
      umma_commit_a(acc_bar);                      // accumulators complete
    }
    __syncwarp();
    // ---- phase C: accumulator rows -> output
    mbar_wait_a(acc_bar, tile_count & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const int q = warp & 3, h = warp >> 2;                           // TMEM lane quadrant, column half
      const long long row = row0 + q * 32 + lane;
      T *orow = (T *)P.out + (size_t)(row < rows ? row : 0) * N + h * (N / 2);
      for (int c0 = 0; c0 < N / 2; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * (N / 2) + c0), r);
        if (row < rows) {
          float f0[8], f1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { f0[e] = __uint_as_float(r[e]); f1[e] = __uint_as_float(r[8 + e]); }
          *reinterpret_cast<uint4 *>(orow + c0) = pack8v<T>(f0);
          *reinterpret_cast<uint4 *>(orow + c0 + 8) = pack8v<T>(f1);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();        // TMEM and the A tile are free for the next tile
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T, int KB>
static int launch_ln_outproj(const actk_merge_ln_args *a, const void *w_out, void *out, int N, int dtype, cudaStream_t stream) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  if (!fn) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available from this driver");
  const int D = a->D;
  CUtensorMap wmap;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)D * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)kWSlab, (cuuint32_t)(N / 2)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&wmap, dtype == ACTK_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void *>(w_out), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ACTK_FAIL(ACTK_ERR_CUDA, "cuTensorMapEncodeTiled (W_out) failed with CUresult %d", (int)r);
  LnOutParams P;
  P.a = *a; P.out = out; P.N = N;
  const size_t smem = 1024 + (size_t)KB * kRows * 128 + 2 * (size_t)N * kWSlab * sizeof(T) + 64;
  auto kern = ln_outproj_kernel<T, KB>;
  ACTK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  ACTK_CUDA_OK(cudaGetDevice(&dev));
  ACTK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long rows = (long long)a->Bp * a->L;
  const int tiles = (int)((rows + kRows - 1) / kRows);
  kern<<<tiles < sms ? tiles : sms, kLnThreads, smem, stream>>>(P, wmap);
  ACTK_CUDA_OK(cudaGetLastError());
  return ACTK_OK;
}

}  // namespace actk

using namespace actk;

extern "C" int actk_merge_ln_outproj_supported(int D, int d_model, int dtype) {
  return (dtype == ACTK_F16 || dtype == ACTK_BF16) && D == 640 && d_model % 32 == 0 && d_model >= 32 && d_model <= 512;
}

extern "C" int actk_merge_ln_outproj_fwd(const actk_merge_ln_args *a, const void *w_out, void *out, int d_model, void *stream) {
  if (!a || !w_out || !out) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln_outproj: NULL pointer");
  if (!actk_merge_ln_outproj_supported(a->D, d_model, a->dtype))
    ACTK_FAIL(ACTK_ERR_UNSUPPORTED, "merge_ln_outproj: D=%d d_model=%d dtype=%d (built for 16-bit, D = 640, d_model %% 32 == 0, <= 512)",
              a->D, d_model, a->dtype);
  if (a->n_branches < 1 || a->n_branches > 2) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln_outproj: n_branches=%d", a->n_branches);
  if (a->Bp <= 0 || a->L <= 0) ACTK_FAIL(ACTK_ERR_BAD_SHAPE, "merge_ln_outproj: Bp=%d L=%d", a->Bp, a->L);
  if (!a->gamma || !a->beta) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln_outproj: gamma / beta are required");
  for (int i = 0; i < a->n_branches; ++i) {
    if (!a->xz[i] || !a->ydir[i] || !a->selected[i]) ACTK_FAIL(ACTK_ERR_BAD_ARG, "merge_ln_outproj: branch %d has a NULL pointer", i);
    if ((reinterpret_cast<uintptr_t>(a->xz[i]) | reinterpret_cast<uintptr_t>(a->ydir[i])) & 15)
      ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "merge_ln_outproj: branch %d pointer not aligned to 16 bytes", i);
  }
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w_out) | reinterpret_cast<uintptr_t>(a->gamma) |
       reinterpret_cast<uintptr_t>(a->beta)) & 15)
    ACTK_FAIL(ACTK_ERR_BAD_ALIGN, "merge_ln_outproj: out / w_out / gamma / beta not aligned to 16 bytes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == ACTK_F16) return launch_ln_outproj<__half, 10>(a, w_out, out, d_model, a->dtype, st);
  return launch_ln_outproj<__nv_bfloat16, 10>(a, w_out, out, d_model, a->dtype, st);
}
