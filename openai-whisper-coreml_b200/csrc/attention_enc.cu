// Encoder self-attention (non-causal, T = 1500, d_head = 64): softmax(q k^T / sqrt(64)) v per (chunk, head).
// Upstream MultiHeadAttention scales q and k by d_head^-0.25 each; the product of the two scalings is the exact power
// of two 1/8, folded here into the exponent (fp32 softmax, as upstream).
//
// Flash-style streaming over 64-key tiles: one CTA = 128 queries of one (chunk, head), 8 warps x 16 query rows;
// K/V tiles double-buffered in shared memory with cp.async; S = Q K^T and O += P V on mma.sync m16n8k16 (fp16 in,
// fp32 accumulate); online softmax in registers with quad shuffles. Input is the fused QKV GEMM output
// [B*T][3d] (q | k | v), output [B*T][d] fp16.
#include <cuda.h>

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

// ---- the same attention on the tensor cores of sm_100a: tcgen05 + TMEM + TMA ------------------------------------------------------------
// One CTA = 256 queries (two 128-row tiles) of one (chunk, head), one CTA per SM, 384 threads in three warpgroups:
//   warp 0     TMA producer: the two Q tiles once, then [128 keys][64] K and V boxes of the fused QKV matrix through a 3-stage
//              ring (128-byte swizzle; rows past T are zero-filled by the TMA unit). K/V are read once per 256 queries.
//   warp 1     MMA issuer (+ TMEM allocation, all 512 columns): per key tile and query tile w, S_w = Q_w K^T (4 x tcgen05.mma
//              M128 N128 K16, K-major operands) and PV_w = P_w V (8 x M128 N64 K16: P K-major from shared memory, V MN-major —
//              rows are keys, the 64 head columns contiguous, exactly as the QKV GEMM wrote them). S_w(j+1) is issued as soon
//              as the softmax warpgroup has S_w(j) in registers, so the tensor pipe works under the exponentials.
//   warps 4-7 / 8-11   softmax warpgroup of query tile 0 / 1, one query row per thread (TMEM lane = row): the whole S row with
//              one burst of tcgen05.ld (128 registers; setmaxnreg moves the register budget from warpgroup 0 to these two),
//              row maximum, exp2 in the log2 domain, P as fp16 into the swizzled K-major A tiles, fence.proxy.async; after the
//              PV MMA  O = O * corr + PV  in registers (the output accumulator never needs rescaling in TMEM).
constexpr int kEtThreads = 384;
constexpr int kEtTile = 128 * 128;                       // bytes of one [128 rows][64 halves] tile
constexpr int kEtStages = 3;
constexpr int kEtSmem = kEtTile * (2 + 2 * kEtStages + 4) + 256;   // Q0 Q1, ring of (K, V), P0 P1 (2 k-tiles each), barriers

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kEtThreads, 1) encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, int T, int d,
                                                                            __half* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char et_dyn[];   // no static shared memory in this kernel: the window starts 1024-aligned
  unsigned char* smem = et_dyn;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();           // SWIZZLE_128B tiles need it
  unsigned char* sQ = smem;                                   // [2][tile]
  unsigned char* ring = sQ + 2 * kEtTile;                     // [stages][K tile | V tile]
  unsigned char* sP = ring + 2 * kEtStages * kEtTile;         // [2 query tiles][2 k-tiles][128 rows][128 B]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sP + 4 * kEtTile);
  uint64_t* kv_full = q_full + 1;                             // [stages]
  uint64_t* kv_empty = kv_full + kEtStages;                   // [stages]
  uint64_t* s_full = kv_empty + kEtStages;                    // [2] S_w(j) is in TMEM
  uint64_t* s_free = s_full + 2;                              // [2] S_w(j) is in registers: the buffer may be overwritten
  uint64_t* p_full = s_free + 2;                              // [2] P_w(j) is in shared memory, PV_w(j-1) has been read
  uint64_t* pv_full = p_full + 2;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (T + 127) / 128;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQKV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kEtStages; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 128);
      ptx::mbar_init(&p_full[i], 128);
      ptx::mbar_init(&pv_full[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, 2 * kEtTile);
      ptx::tma_load_3d(sQ, &tmQKV, q_full, h * 64, q0, b);
      ptx::tma_load_3d(sQ + kEtTile, &tmQKV, q_full, h * 64, q0 + 128, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kEtStages;
        ptx::mbar_wait(&kv_empty[s], (((uint32_t)(j / kEtStages)) & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * kEtTile);
        ptx::tma_load_3d(ring + s * 2 * kEtTile, &tmQKV, &kv_full[s], d + h * 64, j * 128, b);
        ptx::tma_load_3d(ring + s * 2 * kEtTile + kEtTile, &tmQKV, &kv_full[s], 2 * d + h * 64, j * 128, b);
      }
    } else if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_f16(128, 128);
      constexpr uint32_t idesc_pv = ptx::umma_idesc_f16(128, 64) | (1u << 16);   // B (= V) is MN-major
      ptx::mbar_wait(q_full, 0);
      auto issue_s = [&](int j, int w) {
        const int s = j % kEtStages;
        const uint64_t qdesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sQ + w * kEtTile));
        const uint64_t kdesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(ring + s * 2 * kEtTile));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_f16(tmem_base + w * 128, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0);
        ptx::umma_commit(&s_full[w]);
      };
      ptx::mbar_wait(&kv_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0, 0);
      issue_s(0, 1);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kEtStages;
        if (j + 1 < n_kv) {   // S(j+1) runs on the tensor pipe while the softmax warpgroups work on S(j)
          ptx::mbar_wait(&kv_full[(j + 1) % kEtStages], ((uint32_t)((j + 1) / kEtStages)) & 1u);
          for (int w = 0; w < 2; ++w) {
            ptx::mbar_wait(&s_free[w], (uint32_t)j & 1u);
            ptx::tc_fence_after();
            issue_s(j + 1, w);
          }
        }
        const uint64_t vdesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(ring + s * 2 * kEtTile + kEtTile));
        for (int w = 0; w < 2; ++w) {
          ptx::mbar_wait(&p_full[w], (uint32_t)j & 1u);
          ptx::tc_fence_after();
          const uint64_t pdesc0 = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sP + (2 * w) * kEtTile));
          const uint64_t pdesc1 = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sP + (2 * w + 1) * kEtTile));
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // A: 16 keys = 32 bytes along K inside a 64-key swizzle atom (+2), next atom for k >= 4; B (MN-major): 16 key rows = 2048 bytes (+128)
            const uint64_t pd = (k < 4 ? pdesc0 : pdesc1) + (uint64_t)(2 * (k & 3));
            ptx::umma_f16(tmem_base + 256 + w * 64, pd, vdesc + (uint64_t)(128 * k), idesc_pv, k != 0);
          }
          ptx::umma_commit(&pv_full[w]);
        }
        ptx::umma_commit(&kv_empty[s]);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int w = (warp >> 2) - 1;                           // query tile of this warpgroup
    const int q4 = warp & 3;                                 // TMEM lane quadrant of this warp
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + w * 128 + lane_off, tmem_pv = tmem_base + 256 + w * 64 + lane_off;
    const float c = 0.125f * 1.44269504088896340736f;        // (d_head^-0.25)^2 = 1/8, log2 domain
    float m = -INFINITY, l = 0.f;
    float2 o[32];                                            // packed pairs: FFMA2 / FADD2 halve the fp32 instruction count
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = make_float2(0.f, 0.f);
    const float2 c2 = make_float2(c, c);
    unsigned char* prow = sP + (2 * w) * kEtTile + row * 128;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = T - j * 128;                         // keys of this tile that exist (>= 128: all)
      ptx::mbar_wait(&s_full[w], (uint32_t)j & 1u);
      ptx::tc_fence_after();
      uint32_t v[128];
#pragma unroll
      for (int hc = 0; hc < 4; ++hc) ptx::tmem_ld_32x32(tmem_s + hc * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[hc * 32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s_free[w]);                          // the MMA warp may overwrite S_w with the next tile's scores
      if (valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) v[i] = 0xff800000u;                // -inf: exp2 gives an exact 0
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[i])), mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[i + 2])), mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
      }
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      const float mc = m_new * c;
      const float corr = ex2_approx(m * c - mc);             // m = -inf: 0
      const float2 nmc2 = make_float2(-mc, -mc);
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int hc = 0; hc < 4; ++hc) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[hc * 32 + i]), __uint_as_float(v[hc * 32 + i + 1])), c2, nmc2);
          const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          sum2 = __fadd2_rn(sum2, pp);
          const __half2 hp = __floats2half2_rn(pp.x, pp.y);
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        // 32 keys = four 16-byte chunks of k-tile hc / 2, chunk index xor (row & 7) (SWIZZLE_128B, K-major)
        unsigned char* pt = prow + (hc >> 1) * kEtTile;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (hc & 1) * 4 + ch;
          *reinterpret_cast<uint4*>(pt + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
      }
      l = l * corr + (sum2.x + sum2.y);
      m = m_new;
      ptx::fence_proxy_async();                              // P: generic-proxy writes -> tensor-core (async) proxy
      ptx::mbar_arrive(&p_full[w]);
      ptx::mbar_wait(&pv_full[w], (uint32_t)j & 1u);
      ptx::tc_fence_after();
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t pv[32];
        ptx::tmem_ld_32x32(tmem_pv + hc * 32, pv);
        ptx::tmem_ld_wait();
        const float2 corr2 = make_float2(corr, corr);
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          o[hc * 16 + (i >> 1)] = __ffma2_rn(o[hc * 16 + (i >> 1)], corr2, make_float2(__uint_as_float(pv[i]), __uint_as_float(pv[i + 1])));
      }
      ptx::tc_fence_before();
    }
    const int q = q0 + w * 128 + row;
    if (q < T) {
      const float inv = 1.f / l;
      __half* dst = out + ((size_t)b * T + q) * d + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        const int k = i >> 1;
        const __half2 h0 = __floats2half2_rn(o[k].x * inv, o[k].y * inv), h1 = __floats2half2_rn(o[k + 1].x * inv, o[k + 1].y * inv);
        const __half2 h2 = __floats2half2_rn(o[k + 2].x * inv, o[k + 2].y * inv), h3 = __floats2half2_rn(o[k + 3].x * inv, o[k + 3].y * inv);
        uint4 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
        u.z = *reinterpret_cast<const uint32_t*>(&h2), u.w = *reinterpret_cast<const uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

int launch_encoder_attention(GemmContext* tmaps, const __half* qkv, int B, int T, int n_head, __half* out, cudaStream_t st, int64_t* launches) {
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("WB_ENC_ATTN_TC");
    use_tc = (e && e[0] == '0') ? 0 : 1;
  }
  const int d = n_head * kAttD;
  dim3 grid((T + kAttBM - 1) / kAttBM, n_head, B);
  if (use_tc && tmaps) {
    grid.x = (T + 255) / 256;
    CUtensorMap tm;
    const int rc = gemm_get_tmap(tmaps, qkv, 3 * d, T, B, 3 * d, (long long)T * 3 * d, 128, &tm);
    if (rc) return rc;
    static bool attr_tc_dev[kMaxDevices] = {}; bool& attr_tc = attr_tc_dev[current_device_slot()];
    if (!attr_tc) {
      WB_CUDA_OK(cudaFuncSetAttribute(encoder_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEtSmem));
      attr_tc = true;
    }
    encoder_attention_tc_kernel<<<grid, kEtThreads, kEtSmem, st>>>(tm, T, d, out);
    if (launches) *launches += 1;
    WB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  static bool attr_set_dev[kMaxDevices] = {}; bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    WB_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set = true;
  }
  encoder_attention_kernel<<<grid, kAttThreads, kAttSmem, st>>>(qkv, T, d, out);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
