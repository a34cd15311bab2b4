// Decoder step kernels. Replaces `decoderModel.prediction(x_1:xa:)` (Whisper.swift:36; exported from upstream
// TextDecoder at whisper_to_cml.py:25-43) for one new token per sequence, against a persistent HBM KV cache — the
// reference re-projects the cross-attention K/V of all 1500 positions on every call and has no cache at all.
//
// A step is HBM-bound: per sequence it streams the cross-attention K/V of every layer (L*2*1500*d fp16) and, once per
// batch, the decoder weights. Kernels:
//   embed_kernel          token + learned positional embedding -> fp32 residual stream x [Mb][d]; advances cur_len
//   skinny_gemm_kernel    y[Mb][N] = act(f(x)[Mb][K] W[N][K]^T + b): weight-streaming GEMM for Mb <= 40 rows. f is a fused
//                         input transform (LayerNorm of x / merge of attention split partials / plain), the epilogue is
//                         fused too (GELU, residual +=, QKV scatter straight into the self-attention cache). Weights are
//                         read exactly once with 16-byte coalesced loads directly into mma.sync A fragments (the k index
//                         inside a 32-wide block is permuted identically for both operands, so no shuffle is needed).
//   attn_decode_kernel    one query per (sequence, head) over the cached K/V rows: each warp streams whole [d]-wide rows
//                         (all heads at once, 16 B per lane), 8-lane shuffle dot products, online softmax in fp32,
//                         split over rows across CTAs; partial (m, l, acc) merged by the consumer GEMM's input stage
//   sample_greedy_kernel  logit filters, arg-max, log-softmax of the chosen token, EOT forcing, token append
#include "ops.cuh"
#include "ptx.cuh"

namespace wb {

constexpr float kLog2e = 1.44269504088896340736f;

// ---- embed ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) embed_kernel(const int32_t* __restrict__ tokens, int tokens_ld,
                                                     const __half* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                                     int Mb, int d, int V, float* __restrict__ x, DecodeState* state) {
  const int p = state->cur_len;
  for (int i = threadIdx.x; i < Mb * d; i += blockDim.x) {
    const int b = i / d, c = i - b * d;
    int tok = tokens[(size_t)b * tokens_ld + p];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    x[i] = __half2float(tok_emb[(size_t)tok * d + c]) + pos_emb[(size_t)p * d + c];
  }
  __syncthreads();
  if (threadIdx.x == 0) state->cur_len = p + 1;
}

// ---- skinny GEMM -----------------------------------------------------------------------------------------------------------
constexpr int kSkThreads = 256;
constexpr int kSkKC = 2048;   // activation columns staged in shared memory at a time

struct SkinnyArgs {
  SkinnyDesc d;
  int strips_per_cta;   // 1, 2, 4 or 8 strips of 16 weight rows; the 8 warps split K 8/strips ways
};

template <int MT>
__global__ void __launch_bounds__(kSkThreads) skinny_gemm_kernel(SkinnyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SkinnyDesc& p = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3;
  const int S = a.strips_per_cta, KS = 8 / S;
  const int strip = warp % S, kslice = warp / S;
  const int KC = p.K < kSkKC ? p.K : kSkKC;
  const int xs_stride = KC * 2 + 64;                       // bytes; (stride/16) % 8 == 4 -> conflict-free LDS.128
  unsigned char* xs = smem_raw;
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)MT * 8 * xs_stride);   // [8 warps][16][MT*8]

  const int n_cta = blockIdx.x * S * 16;
  int n_g = n_cta + strip * 16 + grp, n_g8 = n_g + 8;
  n_g = n_g < p.N ? n_g : p.N - 1;
  n_g8 = n_g8 < p.N ? n_g8 : p.N - 1;
  const __half* wrow0 = p.w + (size_t)n_g * p.K + tq * 8;
  const __half* wrow1 = p.w + (size_t)n_g8 * p.K + tq * 8;

  float acc[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;

  for (int kc0 = 0; kc0 < p.K; kc0 += KC) {
    const int kc = (p.K - kc0) < KC ? (p.K - kc0) : KC;
    if (kc0) __syncthreads();
    // ---- input stage: build xs[MT*8][kc] fp16 --------------------------------------------------------------------------
    for (int r = warp; r < MT * 8; r += 8) {
      __half* xr = reinterpret_cast<__half*>(xs + (size_t)r * xs_stride);
      if (r >= p.Mb) {
        for (int c = lane * 8; c < kc; c += 256) *reinterpret_cast<uint4*>(xr + c) = make_uint4(0, 0, 0, 0);
      } else if (p.in_mode == SKINNY_IN_F16) {
        const __half* src = reinterpret_cast<const __half*>(p.in) + (size_t)r * p.K + kc0;
        for (int c = lane * 8; c < kc; c += 256) *reinterpret_cast<uint4*>(xr + c) = *reinterpret_cast<const uint4*>(src + c);
      } else if (p.in_mode == SKINNY_IN_F32) {
        const float* src = reinterpret_cast<const float*>(p.in) + (size_t)r * p.K + kc0;
        for (int c = lane * 4; c < kc; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(xr + c) = u;
        }
      } else if (p.in_mode == SKINNY_IN_LN) {
        // LayerNorm over the full row (K = d <= kSkKC): fp32 statistics, two passes over L2-resident data
        const float* src = reinterpret_cast<const float*>(p.in) + (size_t)r * p.K;
        float s = 0.f;
        for (int c = lane * 4; c < p.K; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          s += v.x + v.y + v.z + v.w;
        }
        const float mean = warp_sum(s) / (float)p.K;
        float q = 0.f;
        for (int c = lane * 4; c < p.K; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          const float e0 = v.x - mean, e1 = v.y - mean, e2 = v.z - mean, e3 = v.w - mean;
          q += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)p.K + 1e-5f);
        for (int c = lane * 4; c < p.K; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          const float4 g = *reinterpret_cast<const float4*>(p.ln_g + c), bb = *reinterpret_cast<const float4*>(p.ln_b + c);
          __half2 h0 = __floats2half2_rn((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y);
          __half2 h1 = __floats2half2_rn((v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(xr + c) = u;
        }
      } else {   // SKINNY_IN_ATTN: merge the row-split partials of attn_decode_kernel (K = d, single chunk)
        const int H = p.n_head, NS = p.n_split;
        const float* ml = p.part_ml + (size_t)r * NS * H * 2;
        const float* pa = p.part_acc + (size_t)r * NS * p.K;
        for (int c = lane * 4; c < p.K; c += 128) {
          const int h = c >> 6;
          float M = -INFINITY;
          for (int s2 = 0; s2 < NS; ++s2) M = fmaxf(M, ml[(s2 * H + h) * 2]);
          float L = 0.f;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s2 = 0; s2 < NS; ++s2) {
            const float m = ml[(s2 * H + h) * 2];
            const float wgt = (m == -INFINITY) ? 0.f : exp2f(m - M);
            L += wgt * ml[(s2 * H + h) * 2 + 1];
            const float4 v = *reinterpret_cast<const float4*>(pa + (size_t)s2 * p.K + c);
            o.x += wgt * v.x, o.y += wgt * v.y, o.z += wgt * v.z, o.w += wgt * v.w;
          }
          const float inv = 1.0f / L;
          __half2 h0 = __floats2half2_rn(o.x * inv, o.y * inv), h1 = __floats2half2_rn(o.z * inv, o.w * inv);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(xr + c) = u;
        }
      }
    }
    __syncthreads();
    // ---- stream this warp's weight rows over its K slice of the chunk ------------------------------------------------------
    const int nblk = kc / 32;
    const int blk0 = (kslice * nblk) / KS, blk1 = ((kslice + 1) * nblk) / KS;
    const unsigned char* xl = xs + (size_t)grp * xs_stride + tq * 16;
    int blk = blk0;
    for (; blk + 4 <= blk1; blk += 4) {
      uint4 wa[4], wb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        wa[u] = ptx::ldg_nc_16(wrow0 + kc0 + (blk + u) * 32);
        wb[u] = ptx::ldg_nc_16(wrow1 + kc0 + (blk + u) * 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t a0[4] = {wa[u].x, wb[u].x, wa[u].y, wb[u].y}, a1[4] = {wa[u].z, wb[u].z, wa[u].w, wb[u].w};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint4 xb = *reinterpret_cast<const uint4*>(xl + (size_t)mt * 8 * xs_stride + (blk + u) * 64);
          const uint32_t b0[2] = {xb.x, xb.y}, b1[2] = {xb.z, xb.w};
          ptx::mma_16816(acc[mt], a0, b0);
          ptx::mma_16816(acc[mt], a1, b1);
        }
      }
    }
    for (; blk < blk1; ++blk) {
      const uint4 wa = ptx::ldg_nc_16(wrow0 + kc0 + blk * 32), wb = ptx::ldg_nc_16(wrow1 + kc0 + blk * 32);
      const uint32_t a0[4] = {wa.x, wb.x, wa.y, wb.y}, a1[4] = {wa.z, wb.z, wa.w, wb.w};
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint4 xb = *reinterpret_cast<const uint4*>(xl + (size_t)mt * 8 * xs_stride + blk * 64);
        const uint32_t b0[2] = {xb.x, xb.y}, b1[2] = {xb.z, xb.w};
        ptx::mma_16816(acc[mt], a0, b0);
        ptx::mma_16816(acc[mt], a1, b1);
      }
    }
  }
  // ---- cross-warp (K split) reduction and epilogue -------------------------------------------------------------------------
  constexpr int MB8 = MT * 8;
  float* myred = red + (size_t)warp * 16 * MB8;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    myred[grp * MB8 + mt * 8 + 2 * tq] = acc[mt][0];
    myred[grp * MB8 + mt * 8 + 2 * tq + 1] = acc[mt][1];
    myred[(grp + 8) * MB8 + mt * 8 + 2 * tq] = acc[mt][2];
    myred[(grp + 8) * MB8 + mt * 8 + 2 * tq + 1] = acc[mt][3];
  }
  __syncthreads();
  const int rows_cta = S * 16;
  const int pos = (p.out_mode == SKINNY_OUT_QKV) ? p.state->cur_len - 1 : 0;
  const int dq = p.N / 3;
  for (int idx = tid; idx < rows_cta * p.Mb; idx += kSkThreads) {
    const int b = idx / rows_cta, rr = idx - b * rows_cta;
    const int st = rr >> 4, r = rr & 15;
    const int n = n_cta + rr;
    if (n >= p.N) continue;
    float v = 0.f;
    for (int ks = 0; ks < KS; ++ks) v += red[(size_t)(ks * S + st) * 16 * MB8 + r * MB8 + b];
    if (p.bias) v += p.bias[n];
    if (p.gelu) v = gelu_erf(v);
    switch (p.out_mode) {
      case SKINNY_OUT_F16:
        reinterpret_cast<__half*>(p.out)[(size_t)b * p.N + n] = __float2half_rn(v);
        break;
      case SKINNY_OUT_F32:
        reinterpret_cast<float*>(p.out)[(size_t)b * p.N + n] = v;
        break;
      case SKINNY_OUT_RESID:
        reinterpret_cast<float*>(p.out)[(size_t)b * p.N + n] += v;
        break;
      default:   // SKINNY_OUT_QKV
        if (n < dq)
          p.q32[(size_t)b * dq + n] = v;
        else if (n < 2 * dq)
          p.kcache[((size_t)b * p.n_ctx + pos) * dq + (n - dq)] = __float2half_rn(v);
        else
          p.vcache[((size_t)b * p.n_ctx + pos) * dq + (n - 2 * dq)] = __float2half_rn(v);
        break;
    }
  }
}

int launch_skinny_gemm(const SkinnyDesc& d, cudaStream_t st, int64_t* launches) {
  if (d.Mb < 1 || d.Mb > 40 || d.K % 128 != 0 || d.N < 16) {
    set_error("skinny_gemm: unsupported shape Mb=%d N=%d K=%d", d.Mb, d.N, d.K);
    return -1;
  }
  if ((d.in_mode == SKINNY_IN_LN || d.in_mode == SKINNY_IN_ATTN) && d.K > kSkKC) {
    set_error("skinny_gemm: fused LN / attention-merge input needs K <= %d", kSkKC);
    return -1;
  }
  const int strips = (d.N + 15) / 16;
  int S = 1;
  while (S < 8 && strips / S > 296) S *= 2;
  SkinnyArgs a{d, S};
  const int MT = (d.Mb + 7) / 8;
  const int KC = d.K < kSkKC ? d.K : kSkKC;
  const size_t smem = (size_t)MT * 8 * (KC * 2 + 64) + (size_t)8 * 16 * MT * 8 * 4;
  const int grid = (strips + S - 1) / S;
#define WB_SK_CASE(M)                                                                                             \
  case M: {                                                                                                       \
    static size_t smem_set = 0;                                                                                   \
    if (smem > smem_set) {                                                                                        \
      WB_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                            \
    }                                                                                                             \
    skinny_gemm_kernel<M><<<grid, kSkThreads, smem, st>>>(a);                                                    \
  } break;
  switch (MT) {
    WB_SK_CASE(1) WB_SK_CASE(2) WB_SK_CASE(3) WB_SK_CASE(4) WB_SK_CASE(5)
    default:
      set_error("skinny_gemm: Mb too large");
      return -1;
  }
#undef WB_SK_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- decode attention --------------------------------------------------------------------------------------------------------
constexpr int kAdThreads = 256;

template <int NJ, int RB>
__global__ void __launch_bounds__(kAdThreads) attn_decode_kernel(AttnDecodeDesc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.n_head, d = p.d;
  float* wm = reinterpret_cast<float*>(smem_raw);   // [8][H]
  float* wl = wm + 8 * H;                           // [8][H]
  float* wacc = wl + 8 * H;                         // [8][d]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, b = blockIdx.y;
  const int n_rows = p.n_rows_fixed > 0 ? p.n_rows_fixed : p.state->cur_len;
  const size_t slab = (size_t)(b / p.kv_share) * p.n_ctx * d;
  const __half* K = p.k + slab;
  const __half* V = p.v + slab;
  const int n_chunks = d >> 3;
  const float sl = 0.125f * kLog2e;   // (d_head^-0.25)^2 = 1/8 exactly; scores kept in the log2 domain

  float qf[NJ][8], acc[NJ][8], m[NJ], l[NJ];
  bool valid[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = lane + 32 * j;
    valid[j] = c < n_chunks;
    m[j] = -INFINITY, l[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[j][e] = 0.f;
      qf[j][e] = valid[j] ? p.q[(size_t)b * d + c * 8 + e] * sl : 0.f;
    }
  }
  const int n_units = (n_rows + RB - 1) / RB;
  for (int u = split * 8 + warp; u < n_units; u += p.n_split * 8) {
    const int r0 = u * RB;
    uint4 kr[RB][NJ], vr[RB][NJ];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int row = (r0 + i < n_rows) ? r0 + i : n_rows - 1;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (valid[j]) {
          kr[i][j] = ptx::ldg_nc_16(K + (size_t)row * d + (lane + 32 * j) * 8);
          vr[i][j] = ptx::ldg_nc_16(V + (size_t)row * d + (lane + 32 * j) * 8);
        } else {
          kr[i][j] = make_uint4(0, 0, 0, 0);
          vr[i][j] = make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float sc[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const __half2* kh = reinterpret_cast<const __half2*>(&kr[i][j]);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 kf = __half22float2(kh[e]);
          s = fmaf(qf[j][2 * e], kf.x, s);
          s = fmaf(qf[j][2 * e + 1], kf.y, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        sc[i] = (r0 + i < n_rows) ? s : -INFINITY;
      }
      float mx = sc[0];
#pragma unroll
      for (int i = 1; i < RB; ++i) mx = fmaxf(mx, sc[i]);
      const float m_new = fmaxf(m[j], mx);
      const float corr = exp2f(m[j] - m_new);
      float pr[RB], ps = 0.f;
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        pr[i] = exp2f(sc[i] - m_new);
        ps += pr[i];
      }
      l[j] = l[j] * corr + ps;
      m[j] = m_new;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] *= corr;
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const __half2* vh = reinterpret_cast<const __half2*>(&vr[i][j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 vf = __half22float2(vh[e]);
          acc[j][2 * e] = fmaf(pr[i], vf.x, acc[j][2 * e]);
          acc[j][2 * e + 1] = fmaf(pr[i], vf.y, acc[j][2 * e + 1]);
        }
      }
    }
  }
  // per-warp partials -> shared
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (valid[j]) {
      const int c = lane + 32 * j;
      if ((c & 7) == 0) {
        wm[warp * H + (c >> 3)] = m[j];
        wl[warp * H + (c >> 3)] = l[j];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) wacc[(size_t)warp * d + c * 8 + e] = acc[j][e];
    }
  }
  __syncthreads();
  float* out_acc = p.part_acc + ((size_t)b * p.n_split + split) * d;
  float* out_ml = p.part_ml + ((size_t)b * p.n_split + split) * H * 2;
  for (int c = tid; c < d; c += kAdThreads) {
    const int h = c >> 6;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, wm[w * H + h]);
    float L = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mw = wm[w * H + h];
      const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      L += wgt * wl[w * H + h];
      A += wgt * wacc[(size_t)w * d + c];
    }
    out_acc[c] = A;
    if ((c & 63) == 0) {
      out_ml[h * 2] = M;
      out_ml[h * 2 + 1] = L;
    }
  }
}

int launch_attn_decode(const AttnDecodeDesc& p, cudaStream_t st, int64_t* launches) {
  if (p.d % 64 != 0 || p.d / 64 != p.n_head || p.d > 1280 || p.kv_share < 1) {
    set_error("attn_decode: unsupported d=%d heads=%d", p.d, p.n_head);
    return -1;
  }
  const int NJ = (p.d / 8 + 31) / 32;
  const size_t smem = (size_t)(16 * p.n_head + 8 * p.d) * 4;
  dim3 grid(p.n_split, p.Mb);
#define WB_AD_CASE(J, R)                                                                                          \
  case J: {                                                                                                       \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      WB_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel<J, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024)); \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    attn_decode_kernel<J, R><<<grid, kAdThreads, smem, st>>>(p);                                                 \
  } break;
  switch (NJ) {
    WB_AD_CASE(1, 4) WB_AD_CASE(2, 4) WB_AD_CASE(3, 2) WB_AD_CASE(4, 2) WB_AD_CASE(5, 2)
    default:
      set_error("attn_decode: width too large");
      return -1;
  }
#undef WB_AD_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_embed(const int32_t* tokens, int tokens_ld, const __half* tok_emb, const float* pos_emb, int Mb, int d, int V,
                 float* x, DecodeState* state, cudaStream_t st, int64_t* launches) {
  embed_kernel<<<1, 1024, 0, st>>>(tokens, tokens_ld, tok_emb, pos_emb, Mb, d, V, x, state);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- greedy sampling ---------------------------------------------------------------------------------------------------------
// Upstream DecodingTask with temperature 0: SuppressBlank at the first sampled position, SuppressTokens,
// argmax, sum_logprobs += logprob * (previous token != eot), sequences that ended keep emitting eot.
__global__ void __launch_bounds__(1024) sample_greedy_kernel(SampleDesc p) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float s_max;
  __shared__ int s_arg;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* row = p.logits + (size_t)b * p.V;
  const int cur_len = p.state->cur_len;   // tokens in the context, including the one just consumed
  for (int i = tid; i < p.n_suppress; i += blockDim.x) {
    const int id = p.suppress[i];
    if (id >= 0 && id < p.V) row[id] = -INFINITY;
  }
  if (cur_len == p.n_initial) {
    for (int i = tid; i < p.n_suppress_begin; i += blockDim.x) {
      const int id = p.suppress_begin[i];
      if (id >= 0 && id < p.V) row[id] = -INFINITY;
    }
  }
  __syncthreads();
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int i = tid; i < p.V; i += blockDim.x) {
    const float v = row[i];
    if (v > best || (v == best && i < arg)) best = v, arg = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > best || (ov == best && oi < arg)) best = ov, arg = oi;
  }
  if (lane == 0) s_val[warp] = best, s_idx[warp] = arg;
  __syncthreads();
  if (warp == 0) {
    best = s_val[lane], arg = s_idx[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ov > best || (ov == best && oi < arg)) best = ov, arg = oi;
    }
    if (lane == 0) s_max = best, s_arg = arg;
  }
  __syncthreads();
  const float mx = s_max;
  float se = 0.f;
  for (int i = tid; i < p.V; i += blockDim.x) se += expf(row[i] - mx);
  se = warp_sum(se);
  __syncthreads();
  if (lane == 0) s_val[warp] = se;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < 32; ++w) tot += s_val[w];
    const float logprob = -logf(tot);   // chosen logit equals the maximum
    int32_t* trow = p.tokens + (size_t)b * p.tokens_ld;
    const bool ended = trow[cur_len - 1] == p.eot;
    if (!ended) p.sum_logprob[b] += logprob;
    const int next = ended ? p.eot : s_arg;
    trow[cur_len] = next;
    p.done[b] = next == p.eot;
  }
}

int launch_sample_greedy(const SampleDesc& d, cudaStream_t st, int64_t* launches) {
  sample_greedy_kernel<<<d.Mb, 1024, 0, st>>>(d);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
