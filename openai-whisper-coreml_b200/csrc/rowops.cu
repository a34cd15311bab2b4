// Row-wise helpers: LayerNorm (fp32 in, fp16/fp32 out), dtype conversion, seeded synthetic weight fill.
#include "ops.cuh"

namespace wb {

// One warp per row; d % 128 == 0 (every Whisper width is), d <= 1280. Two-pass statistics in registers, fp32,
// biased variance, eps = 1e-5 — upstream whisper's LayerNorm (fp32 even under fp16 inference).
template <int NV>   // float4 per lane = d / 128
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int M, __half* __restrict__ o16,
                                                        float* __restrict__ o32) {
  constexpr int d = NV * 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = xr[lane + 32 * j];
    s += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  const float mean = warp_sum(s) * (1.0f / d);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
    q += a * a + b * b + c * c + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), bb = *reinterpret_cast<const float4*>(beta + c);
    const float y0 = (v[j].x - mean) * rstd * g.x + bb.x, y1 = (v[j].y - mean) * rstd * g.y + bb.y;
    const float y2 = (v[j].z - mean) * rstd * g.z + bb.z, y3 = (v[j].w - mean) * rstd * g.w + bb.w;
    if (o16) {
      __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(o16 + (size_t)row * d + c) = u;
    }
    if (o32) *reinterpret_cast<float4*>(o32 + (size_t)row * d + c) = make_float4(y0, y1, y2, y3);
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, int M, int d, __half* out16, float* out32,
                     cudaStream_t st, int64_t* launches) {
  const int grid = (M + 7) / 8;
  switch (d / 128) {
#define WB_LN_CASE(NV) \
  case NV:             \
    layernorm_kernel<NV><<<grid, 256, 0, st>>>(x, gamma, beta, M, out16, out32); \
    break;
    WB_LN_CASE(1) WB_LN_CASE(2) WB_LN_CASE(3) WB_LN_CASE(4) WB_LN_CASE(5) WB_LN_CASE(6) WB_LN_CASE(8) WB_LN_CASE(10)
#undef WB_LN_CASE
    default:
      set_error("layernorm: unsupported width %d", d);
      return -1;
  }
  if (d % 128) {
    set_error("layernorm: width %d not a multiple of 128", d);
    return -1;
  }
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half2float(in[i]);
}
int launch_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t st, int64_t* launches) {
  if (n == 0) return 0;
  f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_f16_to_f32(const __half* in, float* out, size_t n, cudaStream_t st, int64_t* launches) {
  if (n == 0) return 0;
  f16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// splitmix64 -> two uniforms -> Box-Muller normal. value = offset + scale * N(0,1)
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void fill_random_kernel(void* ptr, size_t n, int is_f16, float scale, float offset, uint64_t seed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = splitmix64(seed * 0x100000001b3ull + i);
  const float u1 = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777217.0f);
  const float u2 = (float)(uint32_t)((r >> 8) & 0xffffff) * (1.0f / 16777216.0f);
  const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  const float v = offset + scale * z;
  if (is_f16)
    reinterpret_cast<__half*>(ptr)[i] = __float2half_rn(v);
  else
    reinterpret_cast<float*>(ptr)[i] = v;
}
int launch_fill_random(void* ptr, size_t n, int is_f16, float scale, float offset, uint64_t seed, cudaStream_t st,
                       int64_t* launches) {
  if (n == 0) return 0;
  fill_random_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ptr, n, is_f16, scale, offset, seed);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// Position-weighted 64-bit checksum of a byte range (size a multiple of 4): sum over 32-bit words w_i of w_i * (2i + 1)
// mod 2^64. Integer adds commute, so the value does not depend on the order the atomics land in. Used to verify that
// every rank holds the same weight arena after the load-time broadcast.
__global__ void checksum64_kernel(const uint32_t* __restrict__ words, size_t n, unsigned long long* __restrict__ acc) {
  unsigned long long s = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += (unsigned long long)words[i] * (2ull * i + 1ull);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}
int launch_checksum64(const void* ptr, size_t bytes, unsigned long long* acc, cudaStream_t st, int64_t* launches) {
  WB_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(*acc), st));
  const size_t n = bytes / 4;
  if (n) checksum64_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(ptr), n, acc);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
