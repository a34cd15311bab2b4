// Encoder-side GEMM: C = act(A * W^T + bias) (+ residual), fp16 operands, fp32 accumulation in TMEM.
//
// Replaces the dense layers of the reference's CoreML encoder (exported at whisper_to_cml.py:10-23 from upstream
// AudioEncoder): conv1/conv2 (as implicit GEMMs over overlapping-row tensor maps), the fused QKV / out / MLP linears,
// and the decoder's cross-attention K/V projection of the audio features.
//
// Kernel (gemm_tc_kernel): persistent, one CTA per SM looping over 128 x BN output tiles (BN = 256 when N allows), 320 threads:
//   warp 0    TMA producer: cp.async.bulk.tensor (128B swizzle) of a 128x64 A tile and a BNx64 W tile per stage
//   warp 1    TMEM allocator + MMA issuer: 4 x tcgen05.mma (M=128, N=BN, K=16) per stage, tcgen05.commit frees the stage;
//             two accumulator buffers in TMEM, so tile i+1 is multiplied while tile i is drained
//   warps 2-9 epilogue: tcgen05.ld the fp32 accumulator (one TMEM lane = one output row per thread), bias / GELU /
//             residual in registers; the results leave through a swizzled shared-memory staging tile per warp and
//             cp.async.bulk.tensor stores ([32 rows][128 bytes] boxes; rows past the end of a batch are clipped by the TMA
//             unit), instead of 16-byte stores that touch 32 different lines per instruction
#include <cuda.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "ops.cuh"
#include "ptx.cuh"

namespace wb {

constexpr int kBM = 128, kBK = 64;
constexpr int kGemmThreads = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue

struct GemmEpi {
  const float* bias;
  const float* res;
  __half* c16;
  float* c32;
  int N, K, rows, gelu, res_mode, ldc, c_row_off;
  long long c_batch_rows;
  int tiles_m, tiles_n, n_tiles;   // per batch: tiles_m x tiles_n; n_tiles = n_batch * tiles_m * tiles_n
  int tma_store;                   // outputs through shared memory + tensor stores (tmC16 / tmC32)
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageOutOff = kStages * kStageBytes;   // [8 epilogue warps][32 rows][128 B], 1024-byte aligned, SWIZZLE_128B
  static constexpr int kBarOff = kStageOutOff + 8 * 4096;
  static constexpr int kTotal = kBarOff + 256 + 1024;   // barriers + slack for the 1024-byte alignment
};

// exact-erf GELU evaluated with the Abramowitz-Stegun 7.1.26 rational form (|erf error| <= 1.5e-7, far below the fp16
// rounding of the stored result): ~18 instructions instead of erff's two divergent branches, which matters because the
// GELU epilogue of the MLP GEMM is what paces that kernel
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, __expf(-z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
// The same formula on a pair of values: packed fp32 arithmetic (FFMA2 / FMUL2 halve the instruction count) and single-MUFU
// reciprocal / exp2 (rcp.approx, ex2.approx: ~1 ulp, far below the fp16 rounding of the stored result). The GELU epilogue
// issues more instructions per tile than the tile's MMAs take cycles, so this is what paces the MLP1 and conv GEMMs.
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = __fmul2_rn(ax, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 dn = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(dn.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(dn.y));
  float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  // exp(-z^2) = exp2(-z^2 * log2(e))
  const float2 w = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.44269504088896340736f, -1.44269504088896340736f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(w.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(w.y));
  const float2 pt = __fmul2_rn(p, t);
  const float2 erf_abs = __ffma2_rn(make_float2(-pt.x, -pt.y), e, make_float2(1.0f, 1.0f));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  // 0.5 x (1 + sign(x) erf_abs) = hx + |hx| erf_abs
  return __ffma2_rn(make_float2(fabsf(hx.x), fabsf(hx.y)), erf_abs, hx);
}

// Persistent: one CTA per SM loops over output tiles (n fastest, so the CTAs running at the same time share A tiles in
// L2). The TMA->MMA shared-memory ring runs continuously across tiles; the fp32 accumulator is double-buffered in TMEM
// (2 x BN columns), so the epilogue of tile i (8 warps: TMEM lane quadrant x column half) overlaps the MMAs of tile i+1.
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmC16,
                                                                  const __grid_constant__ CUtensorMap tmC32, GemmEpi ep) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;     // [2] accumulator buffer ready for the epilogue
  uint64_t* tempty = tfull + 2;          // [2] accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (ep.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (ep.tma_store && ep.c16) ptx::prefetch_tensormap(&tmC16);
    if (ep.tma_store && ep.c32) ptx::prefetch_tensormap(&tmC32);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 2 * BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ep.n_tiles; tile += gridDim.x) {
        const int nt = tile % ep.tiles_n, mt = (tile / ep.tiles_n) % ep.tiles_m, b = tile / (ep.tiles_n * ep.tiles_m);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], S::kStageBytes);
          unsigned char* sa = smem + s * S::kStageBytes;
          ptx::tma_load_3d(sa, &tmA, &full[s], kb * kBK, mt * kBM, b);
          ptx::tma_load_2d(sa + S::kABytes, &tmB, &full[s], kb * kBK, nt * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(kBM, BN);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < ep.n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1u, tph = (lt >> 1) & 1u;
        ptx::mbar_wait(&tempty[buf], tph ^ 1u);   // epilogue has drained this accumulator buffer
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * S::kStageBytes);
          const uint64_t adesc = ptx::umma_desc_sw128_kmajor(sa);
          const uint64_t bdesc = ptx::umma_desc_sw128_kmajor(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in 16-byte units
            ptx::umma_f16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty[s]);
        }
        ptx::umma_commit(&tfull[buf]);
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); warps 2-5 take the first column half, 6-9 the second
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int kChunks = BN / 64;   // 32-column chunks per warp
    // staging tile of this warp: row = lane, 128 bytes per row, 16-byte chunk j of a row at position j ^ (row & 7) (SWIZZLE_128B)
    unsigned char* stage = smem + S::kStageOutOff + (warp - 2) * 4096;
    unsigned char* srow = stage + lane * 128;
    const int sw = lane & 7;
    bool store_pending = false;        // a tensor store of this warp may still be reading `stage`
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < ep.n_tiles; tile += gridDim.x, ++lt) {
      const int nt = tile % ep.tiles_n, mt = (tile / ep.tiles_n) % ep.tiles_m, b = tile / (ep.tiles_n * ep.tiles_m);
      const uint32_t buf = lt & 1u, tph = (lt >> 1) & 1u;
      const int t0 = mt * kBM + q * 32;
      const int t = t0 + lane;
      const bool row_ok = t < ep.rows;
      const long long crow = (long long)b * ep.c_batch_rows + ep.c_row_off + t;
      ptx::mbar_wait(&tfull[buf], tph);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col = half * (BN / 2) + c * 32;
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + buf * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
        ptx::tmem_ld_wait();
        if (c == kChunks - 1) {   // everything this warp needs from the buffer is in registers: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
        }
        const int nb = nt * BN + col;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + nb + j));
            f[j] += bb.x, f[j + 1] += bb.y, f[j + 2] += bb.z, f[j + 3] += bb.w;
          }
        }
        if (ep.gelu) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 g = gelu_erf_fast2(make_float2(f[j], f[j + 1]));
            f[j] = g.x, f[j + 1] = g.y;
          }
        }
        if (ep.res_mode && row_ok) {
          const float* rp = ep.res + (ep.res_mode == 1 ? crow * ep.ldc : (long long)t * ep.N) + nb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 rr = *reinterpret_cast<const float4*>(rp + j);
            f[j] += rr.x, f[j + 1] += rr.y, f[j + 2] += rr.z, f[j + 3] += rr.w;
          }
        }
        if (ep.tma_store) {
          // (exactly one of c32 / c16 is set on this path) rows past the end of the batch hold finite garbage - the A rows
          // there were zero-filled - and the tensor store clips them
          if (ep.c32) {   // 32 fp32 columns = one 128-byte row of the staging tile
            if (store_pending) {
              if (lane == 0) ptx::bulk_wait_read0();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_3d(&tmC32, stage, nb, t0, b);
              ptx::bulk_commit();
            }
            store_pending = true;
          }
          if (ep.c16) {   // 32 fp16 columns = half a row: chunks 0-3 (even c) or 4-7 (odd c); stored once the row is complete
            if ((c & 1) == 0 && store_pending) {
              if (lane == 0) ptx::bulk_wait_read0();
              __syncwarp();
              store_pending = false;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __half2 h0 = __floats2half2_rn(f[8 * j], f[8 * j + 1]), h1 = __floats2half2_rn(f[8 * j + 2], f[8 * j + 3]);
              const __half2 h2 = __floats2half2_rn(f[8 * j + 4], f[8 * j + 5]), h3 = __floats2half2_rn(f[8 * j + 6], f[8 * j + 7]);
              uint4 u;
              u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
              u.z = *reinterpret_cast<const uint32_t*>(&h2), u.w = *reinterpret_cast<const uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(srow + ((((c & 1) * 4 + j) ^ sw) << 4)) = u;
            }
            if (c & 1) {
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                ptx::tma_store_3d(&tmC16, stage, nb - 32, t0, b);
                ptx::bulk_commit();
              }
              store_pending = true;
            }
          }
        } else if (row_ok) {
          if (ep.c32) {
            float* cp = ep.c32 + crow * ep.ldc + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
          if (ep.c16) {
            __half* cp = ep.c16 + crow * ep.ldc + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __half2 h0 = __floats2half2_rn(f[j], f[j + 1]), h1 = __floats2half2_rn(f[j + 2], f[j + 3]);
              __half2 h2 = __floats2half2_rn(f[j + 4], f[j + 5]), h3 = __floats2half2_rn(f[j + 6], f[j + 7]);
              uint4 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
              u.z = *reinterpret_cast<uint32_t*>(&h2), u.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(cp + j) = u;
            }
          }
        }
      }
    }
    if (store_pending && lane == 0) ptx::bulk_wait0();   // the last tiles are in global memory before the CTA retires
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct GemmContext {
  PFN_encodeTiled encode = nullptr;
  std::map<std::tuple<const void*, long long, long long, long long, long long, long long, int>, CUtensorMap> cache;
};

GemmContext* gemm_context_create() {
  GemmContext* c = new GemmContext();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    c->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return c;
}
void gemm_context_destroy(GemmContext* c) { delete c; }

// fp16 tensor (k, rows, batch) with element strides (1, row_stride, batch_stride); box (64, box_rows, 1); 128B swizzle
static int get_tmap(GemmContext* ctx, const void* ptr, long long K, long long rows, long long nb, long long row_stride,
                    long long batch_stride, int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, K, rows, nb, row_stride, batch_stride, box_rows);
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) {
    *out = it->second;
    return 0;
  }
  if (!ctx->encode) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return -2;
  }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = ctx->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (K=%lld rows=%lld nb=%lld rs=%lld bs=%lld)", (int)r, K, rows, nb,
              row_stride, batch_stride);
    return -2;
  }
  ctx->cache[key] = m;
  *out = m;
  return 0;
}

// fp16 / fp32 output (n, rows, batch) with element strides (1, ldc, batch_rows * ldc); box (128 bytes, 32 rows, 1); 128B swizzle
static int get_store_tmap(GemmContext* ctx, const void* ptr, bool f32, long long N, long long rows, long long nb, long long ldc,
                          long long batch_rows, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, N, rows, nb, ldc, batch_rows, f32 ? -32 : -16);
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) {
    *out = it->second;
    return 0;
  }
  if (!ctx->encode) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return -2;
  }
  const cuuint64_t esz = f32 ? 4 : 2;
  const long long brows = (nb > 1 && batch_rows > 0) ? batch_rows : rows;   // a single batch: any valid stride
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t strides[2] = {(cuuint64_t)ldc * esz, (cuuint64_t)brows * (cuuint64_t)ldc * esz};
  cuuint32_t box[3] = {(cuuint32_t)(128 / esz), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = ctx->encode(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims,
                                 strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(C) failed: %d (N=%lld rows=%lld nb=%lld ldc=%lld)", (int)r, N, rows, nb, ldc);
    return -2;
  }
  ctx->cache[key] = m;
  *out = m;
  return 0;
}

int gemm_get_tmap(GemmContext* ctx, const void* ptr, long long K, long long rows, long long nb, long long row_stride,
                  long long batch_stride, int box_rows, void* out128) {
  return get_tmap(ctx, ptr, K, rows, nb, row_stride, batch_stride, box_rows, reinterpret_cast<CUtensorMap*>(out128));
}

template <int BN>
static int launch_tc(GemmContext* ctx, const GemmDesc& d, const GemmEpi& ep, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc = get_tmap(ctx, d.a, d.K, d.rows, d.n_batch, d.a_row_stride, d.a_batch_stride, kBM, &tmA);
  if (rc) return rc;
  // W as a 3-D map with a unit batch, so both operands share one encoder; the kernel loads it with 2-D coordinates
  cuuint64_t dims[2] = {(cuuint64_t)d.K, (cuuint64_t)d.N};
  cuuint64_t strides[1] = {(cuuint64_t)d.K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  auto key = std::make_tuple((const void*)d.w, (long long)d.K, (long long)d.N, -1LL, (long long)d.K, 0LL, BN);
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) {
    tmB = it->second;
  } else {
    if (!ctx->encode) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return -2;
    }
    const CUresult r = ctx->encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(d.w), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(W) failed: %d (N=%d K=%d)", (int)r, d.N, d.K);
      return -2;
    }
    ctx->cache[key] = tmB;
  }
  // output maps for the tensor-store epilogue: (columns, rows of a batch, batches); fp16 boxes [32][64], fp32 boxes [32][32]
  // (128 bytes per row either way); one output only, whole 64-column pairs per warp (BN / 64 even)
  static int tma_store_env = -1;
  if (tma_store_env < 0) {
    const char* e = getenv("WB_GEMM_TMA_STORE");   // development: 0 = per-thread 16-byte stores
    tma_store_env = (e && e[0] == '0') ? 0 : 1;
  }
  CUtensorMap tmC16{}, tmC32{};
  GemmEpi e2 = ep;
  e2.tma_store = (tma_store_env && ((d.c16 != nullptr) != (d.c32 != nullptr)) && (BN / 64) % 2 == 0) ? 1 : 0;
  if (e2.tma_store) {
    const bool f32 = d.c32 != nullptr;
    rc = get_store_tmap(ctx, f32 ? (const void*)(d.c32 + (long long)d.c_row_off * d.ldc) : (const void*)(d.c16 + (long long)d.c_row_off * d.ldc),
                        f32, d.N, d.rows, d.n_batch, d.ldc, d.c_batch_rows, f32 ? &tmC32 : &tmC16);
    if (rc) return rc;
  }
  auto kern = gemm_tc_kernel<BN>;
  static bool attr_set_dev[kMaxDevices] = {}; bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    WB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::kTotal));
    attr_set = true;
  }
  e2.tiles_n = d.N / BN, e2.tiles_m = (d.rows + kBM - 1) / kBM, e2.n_tiles = e2.tiles_n * e2.tiles_m * d.n_batch;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const int grid = e2.n_tiles < n_sm ? e2.n_tiles : n_sm;   // persistent: one CTA per SM
  kern<<<grid, kGemmThreads, GemmSmem<BN>::kTotal, st>>>(tmA, tmB, tmC16, tmC32, e2);
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gemm(GemmContext* ctx, const GemmDesc& d, cudaStream_t st, int64_t* launches) {
  if (d.N % 128 != 0 || d.K % 8 != 0 || d.a_row_stride % 8 != 0 || d.a_batch_stride % 8 != 0 || d.ldc % 8 != 0) {
    set_error("launch_gemm: unsupported shape N=%d K=%d (N%%128, K%%8, strides%%8 required)", d.N, d.K);
    return -1;
  }
  GemmEpi ep;
  ep.bias = d.bias, ep.res = d.res, ep.c16 = d.c16, ep.c32 = d.c32;
  ep.N = d.N, ep.K = d.K, ep.rows = d.rows, ep.gelu = d.gelu, ep.res_mode = d.res_mode, ep.ldc = d.ldc;
  ep.c_row_off = d.c_row_off, ep.c_batch_rows = d.c_batch_rows;
  if (launches) *launches += 1;
  static int force_bn = -1;   // WB_GEMM_BN=128|256: development override of the tile width
  if (force_bn < 0) {
    const char* e = getenv("WB_GEMM_BN");
    force_bn = e ? atoi(e) : 0;
  }
  if ((d.N % 256 == 0 && force_bn != 128) || force_bn == 256) {
    if (d.N % 256 == 0) return launch_tc<256>(ctx, d, ep, st);
  }
  return launch_tc<128>(ctx, d, ep, st);
}

}  // namespace wb
