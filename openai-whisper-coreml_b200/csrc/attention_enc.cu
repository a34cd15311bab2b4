// Encoder self-attention (non-causal, T = 1500, d_head = 64): softmax(q k^T / sqrt(64)) v per (chunk, head).
// Upstream MultiHeadAttention scales q and k by d_head^-0.25 each; the product of the two scalings is the exact power
// of two 1/8, folded here into the exponent (fp32 softmax, as upstream).
//
// Flash-style streaming over 64-key tiles: one CTA = 128 queries of one (chunk, head), 8 warps x 16 query rows;
// K/V tiles double-buffered in shared memory with cp.async; S = Q K^T and O += P V on mma.sync m16n8k16 (fp16 in,
// fp32 accumulate); online softmax in registers with quad shuffles. Input is the fused QKV GEMM output
// [B*T][3d] (q | k | v), output [B*T][d] fp16.
#include "ops.cuh"
#include "ptx.cuh"

namespace wb {

constexpr int kAttBM = 128, kAttBN = 64, kAttD = 64, kAttLd = 72;   // smem row stride (halves): 144 B, ldmatrix conflict-free
constexpr int kAttThreads = 256;
constexpr int kAttSmem = (kAttBM + 4 * kAttBN) * kAttLd * 2;

__global__ void __launch_bounds__(kAttThreads, 2) encoder_attention_kernel(const __half* __restrict__ qkv, int T, int d,
                                                                           __half* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* sQ = reinterpret_cast<__half*>(smem_raw);
  __half* sK = sQ + kAttBM * kAttLd;          // [2][64][72]
  __half* sV = sK + 2 * kAttBN * kAttLd;      // [2][64][72]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kAttBM, h = blockIdx.y, b = blockIdx.z;
  const size_t ld = (size_t)3 * d;
  const __half* base = qkv + (size_t)b * T * ld;
  const __half* gq = base + (size_t)h * kAttD;
  const __half* gk = base + d + (size_t)h * kAttD;
  const __half* gv = base + 2 * d + (size_t)h * kAttD;
  const int nt = (T + kAttBN - 1) / kAttBN;

  // Q tile: 128 rows x 8 chunks
  for (int c = tid; c < kAttBM * 8; c += kAttThreads) {
    const int r = c >> 3, ch = c & 7;
    const bool ok = q0 + r < T;
    ptx::cp_async_16(ptx::smem_u32(sQ + r * kAttLd + ch * 8), gq + (size_t)(ok ? q0 + r : 0) * ld + ch * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * kAttBN;
    for (int c = tid; c < kAttBN * 8; c += kAttThreads) {
      const int r = c >> 3, ch = c & 7;
      const bool ok = k0 + r < T;
      const size_t off = (size_t)(ok ? k0 + r : 0) * ld + ch * 8;
      ptx::cp_async_16(ptx::smem_u32(sK + (buf * kAttBN + r) * kAttLd + ch * 8), gk + off, ok);
      ptx::cp_async_16(ptx::smem_u32(sV + (buf * kAttBN + r) * kAttLd + ch * 8), gv + off, ok);
    }
  };
  load_kv(0, 0);
  ptx::cp_async_commit();

  const int grp = lane >> 2, tq = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  uint32_t aq[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl = 0.125f * 1.44269504088896340736f;

  for (int j = 0; j < nt; ++j) {
    if (j + 1 < nt) {
      load_kv(j + 1, (j + 1) & 1);
      ptx::cp_async_commit();
      ptx::cp_async_wait<1>();
    } else {
      ptx::cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ptx::ldmatrix_x4(aq[kk], ptx::smem_u32(sQ + (warp * 16 + (mi & 1) * 8 + r8) * kAttLd + kk * 16 + (mi >> 1) * 8));
    }
    const __half* bK = sK + (j & 1) * kAttBN * kAttLd;
    const __half* bV = sV + (j & 1) * kAttBN * kAttLd;

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int half_k = 0; half_k < 2; ++half_k) {
        uint32_t kb[4];
        ptx::ldmatrix_x4(kb, ptx::smem_u32(bK + (n * 8 + r8) * kAttLd + half_k * 32 + mi * 8));
        const uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
        ptx::mma_16816(s[n], aq[half_k * 2], b0);
        ptx::mma_16816(s[n], aq[half_k * 2 + 1], b1);
      }
    }
    // mask keys beyond T (last tile only)
    const int k0 = j * kAttBN;
    if (k0 + kAttBN > T) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int key = k0 + n * 8 + 2 * tq;
        if (key >= T) s[n][0] = s[n][2] = -INFINITY;
        if (key + 1 >= T) s[n][1] = s[n][3] = -INFINITY;
      }
    }
    // online softmax; rows grp (e = 0,1) and grp + 8 (e = 2,3)
    uint32_t pa[4][4];
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * rh], s[n][2 * rh + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[rh], mx);
      const float corr = exp2f((m_run[rh] - m_new) * sl);
      const float mb = m_new * sl;
      float rs = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = exp2f(s[n][2 * rh] * sl - mb), p1 = exp2f(s[n][2 * rh + 1] * sl - mb);
        rs += p0 + p1;
        __half2 hp = __floats2half2_rn(p0, p1);
        // A fragment of P for k-step n/2: reg (rh) for keys of the even tile, reg (2 + rh) for the odd tile
        pa[n >> 1][(n & 1) * 2 + rh] = *reinterpret_cast<uint32_t*>(&hp);
      }
      l_run[rh] = l_run[rh] * corr + rs;
      m_run[rh] = m_new;
#pragma unroll
      for (int n = 0; n < 8; ++n) o[n][2 * rh] *= corr, o[n][2 * rh + 1] *= corr;
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t vb[4];
        ptx::ldmatrix_x4_trans(vb, ptx::smem_u32(bV + (kk * 16 + (mi & 1) * 8 + r8) * kAttLd + (jd + (mi >> 1)) * 8));
        const uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
        ptx::mma_16816(o[jd], pa[kk], b0);
        ptx::mma_16816(o[jd + 1], pa[kk], b1);
      }
    }
    __syncthreads();
  }
  // finalise and store
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    float l = l_run[rh];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;
    const int row = q0 + warp * 16 + grp + rh * 8;
    if (row < T) {
      __half* op = out + ((size_t)b * T + row) * d + (size_t)h * kAttD + 2 * tq;
#pragma unroll
      for (int n = 0; n < 8; ++n) *reinterpret_cast<__half2*>(op + n * 8) = __floats2half2_rn(o[n][2 * rh] * inv, o[n][2 * rh + 1] * inv);
    }
  }
}

int launch_encoder_attention(const __half* qkv, int B, int T, int n_head, __half* out, cudaStream_t st, int64_t* launches) {
  static bool attr_set = false;
  if (!attr_set) {
    WB_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set = true;
  }
  dim3 grid((T + kAttBM - 1) / kAttBM, n_head, B);
  encoder_attention_kernel<<<grid, kAttThreads, kAttSmem, st>>>(qkv, T, n_head * kAttD, out);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
