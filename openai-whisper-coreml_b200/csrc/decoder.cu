// Decoder step kernels. Replaces `decoderModel.prediction(x_1:xa:)` (Whisper.swift:36; exported from upstream
// TextDecoder at whisper_to_cml.py:25-43) for one new token per sequence, against a persistent HBM KV cache — the
// reference re-projects the cross-attention K/V of all 1500 positions on every call and has no cache at all.
//
// A step is HBM-bound: per sequence it streams the cross-attention K/V of every layer (L*2*1500*d fp16) and, once per
// batch, the decoder weights. Kernels:
//   skinny_gemm_kernel    y[Mb][N] = act(f(x)[Mb][K] W[N][K]^T + b): weight-streaming GEMM for Mb <= 40 rows. f is a fused
//                         input transform (LayerNorm of x / merge of attention split partials / plain), the epilogue is
//                         fused too (GELU, residual +=, QKV scatter straight into the self-attention cache). Weights are
//                         read exactly once with 16-byte coalesced loads directly into mma.sync A fragments (the k index
//                         inside a 32-wide block is permuted identically for both operands, so no shuffle is needed).
//   attn_decode_kernel    one query per (sequence, head) over the cached K/V rows: each warp streams whole [d]-wide rows
//                         (all heads at once, 16 B per lane), 8-lane shuffle dot products, online softmax in fp32,
//                         split over rows across CTAs; the last CTA of a sequence to finish merges the split partials
//   step_finish_kernel    merges the per-CTA (max, argmax, sum-exp) partials the logits GEMM epilogue produced (logit
//                         filters already applied there), EOT forcing, log-prob accumulation, token append, then embeds
//                         the next token (+ learned position) into the residual stream and advances the position
#include "ops.cuh"
#include "ptx.cuh"

namespace wb {

constexpr float kLog2e = 1.44269504088896340736f;

// Everything that changes from kernel to kernel inside a decode step (residual stream, q, attention outputs, partials,
// tokens, DecodeState) is read through L2 (.cg): with programmatic dependent launch a kernel can share an SM — and its L1 —
// with its still-running predecessor, so L1 may hold lines the predecessor fetched before another SM rewrote them.
// Weights and the cross-attention K/V are constant during a decode and use the non-coherent path.
__device__ __forceinline__ int ld_state(const int* p) { return __ldcg(p); }

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

  // Weights do not depend on the previous kernel: fetch this warp's first blocks before waiting for it (PDL) and
  // before the input stage, so their DRAM latency overlaps both.
  const int KC0 = p.K < KC ? p.K : KC;
  const int nblk0 = KC0 / 32;
  const int pb0 = (kslice * nblk0) / KS, pb1 = ((kslice + 1) * nblk0) / KS;
  uint4 pwa[4], pwb[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int blk = (pb0 + u < pb1) ? pb0 + u : pb0;
    pwa[u] = ptx::ldg_nc_16(wrow0 + blk * 32);
    pwb[u] = ptx::ldg_nc_16(wrow1 + blk * 32);
  }
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();

  for (int kc0 = 0; kc0 < p.K; kc0 += KC) {
    const int kc = (p.K - kc0) < KC ? (p.K - kc0) : KC;
    if (kc0) __syncthreads();
    // ---- input stage: build xs[MT*8][kc] fp16 --------------------------------------------------------------------------
    if (p.in_mode == SKINNY_IN_F16) {
      const __half* src = reinterpret_cast<const __half*>(p.in);
      const int cpr = kc >> 3;                              // 16-byte chunks per row
      for (int i = tid; i < MT * 8 * cpr; i += kSkThreads) {
        const int r = i / cpr, c = (i - r * cpr) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < p.Mb) v = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * p.K + kc0 + c));
        *reinterpret_cast<uint4*>(xs + (size_t)r * xs_stride + c * 2) = v;
      }
    } else {
      // LayerNorm of the fp32 residual stream (K = d, one chunk). 8 threads per row, 32 rows per pass: every row of the
      // batch is in flight at once; two passes over L2 (statistics, then normalise). fp32, biased variance, eps 1e-5.
      const int sub = tid & 7;
      for (int r = tid >> 3; r < MT * 8; r += kSkThreads / 8) {     // trip count is warp-uniform (MT*8 is a multiple of 8)
        __half* xr = reinterpret_cast<__half*>(xs + (size_t)r * xs_stride);
        const bool act = r < p.Mb;                                   // padding rows compute on row 0 and store zeros
        const float* src = reinterpret_cast<const float*>(p.in) + (size_t)(act ? r : 0) * p.K;
        float s = 0.f, q = 0.f;                                      // one pass: sum and sum of squares (fp32)
#pragma unroll 4
        for (int c = sub * 4; c < p.K; c += 32) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src + c));
          s += (v.x + v.y) + (v.z + v.w);
          q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float mean = s / (float)p.K;
        q = fmaxf(q / (float)p.K - mean * mean, 0.f);
        const float rstd = act ? rsqrtf(q + 1e-5f) : 0.f;
#pragma unroll 4
        for (int c = sub * 4; c < p.K; c += 32) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src + c));
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + c)), bb = __ldg(reinterpret_cast<const float4*>(p.ln_b + c));
          const float ab = act ? 1.f : 0.f;
          __half2 h0 = __floats2half2_rn((v.x - mean) * rstd * g.x + ab * bb.x, (v.y - mean) * rstd * g.y + ab * bb.y);
          __half2 h1 = __floats2half2_rn((v.z - mean) * rstd * g.z + ab * bb.z, (v.w - mean) * rstd * g.w + ab * bb.w);
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
    if (kc0 == 0) {   // the prefetched blocks
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (blk0 + u < blk1) {
          const uint32_t a0[4] = {pwa[u].x, pwb[u].x, pwa[u].y, pwb[u].y}, a1[4] = {pwa[u].z, pwb[u].z, pwa[u].w, pwb[u].w};
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint4 xb = *reinterpret_cast<const uint4*>(xl + (size_t)mt * 8 * xs_stride + (blk0 + u) * 64);
            const uint32_t b0[2] = {xb.x, xb.y}, b1[2] = {xb.z, xb.w};
            ptx::mma_16816(acc[mt], a0, b0);
            ptx::mma_16816(acc[mt], a1, b1);
          }
        }
      }
      blk = blk0 + 4 < blk1 ? blk0 + 4 : blk1;
    }
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
  if (p.out_mode == SKINNY_OUT_LOGITS) {
    // rows_cta == 128 (S == 8, no K split). Warp per sequence: filter, optional store, CTA-local (max, argmax, sum-exp).
    const bool first = ld_state(&p.state->cur_len) + 1 == p.n_initial;
    for (int b = warp; b < p.Mb; b += 8) {
      float v[4];
      float best = -INFINITY;
      int arg = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = lane + 32 * i, n = n_cta + rr;
        float x = -INFINITY;
        if (n < p.N) {
          x = red[(size_t)(rr >> 4) * 16 * MB8 + (rr & 15) * MB8 + b];
          const unsigned char mk = p.mask ? p.mask[n] : 0;
          if (mk == 1 || (mk == 2 && first)) x = -INFINITY;
          if (p.out) reinterpret_cast<float*>(p.out)[(size_t)b * p.N + n] = x;
        }
        v[i] = x;
        if (x > best) best = x, arg = n;     // ascending n: first maximum wins
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oi < arg)) best = ov, arg = oi;
      }
      float se = 0.f;
      if (best > -INFINITY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) se += expf(v[i] - best);
      }
      se = warp_sum(se);
      if (lane == 0) {
        float* pp = p.part_logits + ((size_t)b * gridDim.x + blockIdx.x) * 4;
        pp[0] = best, pp[1] = __int_as_float(arg), pp[2] = se;
      }
    }
    return;
  }
  const int pos = (p.out_mode == SKINNY_OUT_QKV) ? ld_state(&p.state->cur_len) : 0;
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
        {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)b * p.N + n;
        *o = __ldcg(o) + v;
      }
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

static bool use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// launch with the programmatic-dependent-launch attribute: the kernel may start while its predecessor drains; every
// kernel launched this way calls griddepcontrol.wait before touching anything the predecessor wrote
template <typename Kern, typename Arg>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg& arg) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, arg);
}

int skinny_logits_ctas(int N) { return ((N + 15) / 16 + 7) / 8; }

int launch_skinny_gemm(const SkinnyDesc& d, cudaStream_t st, int64_t* launches) {
  if (d.Mb < 1 || d.Mb > 40 || d.K % 128 != 0 || d.N < 16) {
    set_error("skinny_gemm: unsupported shape Mb=%d N=%d K=%d", d.Mb, d.N, d.K);
    return -1;
  }
  if (d.in_mode == SKINNY_IN_LN && d.K > kSkKC) {
    set_error("skinny_gemm: fused LayerNorm input needs K <= %d", kSkKC);
    return -1;
  }
  const int strips = (d.N + 15) / 16;
  int S = 1;
  while (S < 8 && strips / S > 296) S *= 2;
  if (d.out_mode == SKINNY_OUT_LOGITS) S = 8;
  SkinnyArgs a{d, S};
  const int MT = (d.Mb + 7) / 8;
  const int KC = d.K < kSkKC ? d.K : kSkKC;
  const size_t smem = (size_t)MT * 8 * (KC * 2 + 64) + (size_t)8 * 16 * MT * 8 * 4;
  const int grid = (strips + S - 1) / S;
  cudaError_t le = cudaSuccess;
#define WB_SK_CASE(M)                                                                                             \
  case M: {                                                                                                       \
    static size_t smem_set = 0;                                                                                   \
    if (smem > smem_set) {                                                                                        \
      WB_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                            \
    }                                                                                                             \
    le = launch_pdl(skinny_gemm_kernel<M>, dim3(grid), dim3(kSkThreads), smem, st, a);                            \
  } break;
  switch (MT) {
    WB_SK_CASE(1) WB_SK_CASE(2) WB_SK_CASE(3) WB_SK_CASE(4) WB_SK_CASE(5)
    default:
      set_error("skinny_gemm: Mb too large");
      return -1;
  }
#undef WB_SK_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// ---- decode attention --------------------------------------------------------------------------------------------------------
constexpr int kAdThreads = 256;

template <int NJ, int RB>
__global__ void __launch_bounds__(kAdThreads) attn_decode_kernel(AttnDecodeDesc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_last;
  const int H = p.n_head, d = p.d;
  float* wm = reinterpret_cast<float*>(smem_raw);   // [8][H]
  float* wl = wm + 8 * H;                           // [8][H]
  float* wacc = wl + 8 * H;                         // [8][d]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, b = blockIdx.y;
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  const int n_rows = p.n_rows_fixed > 0 ? p.n_rows_fixed : ld_state(&p.state->cur_len) + 1;
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
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f, qf[j][e] = 0.f;
    if (valid[j]) {
      const float4 q0 = __ldcg(reinterpret_cast<const float4*>(p.q + (size_t)b * d + c * 8));
      const float4 q1 = __ldcg(reinterpret_cast<const float4*>(p.q + (size_t)b * d + c * 8 + 4));
      qf[j][0] = q0.x * sl, qf[j][1] = q0.y * sl, qf[j][2] = q0.z * sl, qf[j][3] = q0.w * sl;
      qf[j][4] = q1.x * sl, qf[j][5] = q1.y * sl, qf[j][6] = q1.z * sl, qf[j][7] = q1.w * sl;
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
    if (p.n_split == 1) {
      p.out16[(size_t)b * d + c] = __float2half_rn(A / L);
    } else {
      out_acc[c] = A;
      if ((c & 63) == 0) {
        out_ml[h * 2] = M;
        out_ml[h * 2 + 1] = L;
      }
    }
  }
  if (p.n_split == 1) return;
  // the last CTA of this sequence to arrive merges the splits (fixed split order -> deterministic result)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int ticket = atomicAdd(&p.counters[b], 1);
    s_last = ticket == p.n_split - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* all_ml = p.part_ml + (size_t)b * p.n_split * H * 2;
  const float* all_acc = p.part_acc + (size_t)b * p.n_split * d;
  for (int c = tid; c < d; c += kAdThreads) {
    const int h = c >> 6;
    float M = -INFINITY;
    for (int s2 = 0; s2 < p.n_split; ++s2) M = fmaxf(M, __ldcg(all_ml + (s2 * H + h) * 2));
    float L = 0.f, A = 0.f;
    for (int s2 = 0; s2 < p.n_split; ++s2) {
      const float ms = __ldcg(all_ml + (s2 * H + h) * 2);
      const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
      L += wgt * __ldcg(all_ml + (s2 * H + h) * 2 + 1);
      A += wgt * __ldcg(all_acc + (size_t)s2 * d + c);
    }
    p.out16[(size_t)b * d + c] = __float2half_rn(A / L);
  }
  if (tid == 0) p.counters[b] = 0;   // ready for the next launch (graph replay)
}

int launch_attn_decode(const AttnDecodeDesc& p, cudaStream_t st, int64_t* launches) {
  if (p.d % 64 != 0 || p.d / 64 != p.n_head || p.d > 1280 || p.kv_share < 1 || p.n_split < 1) {
    set_error("attn_decode: unsupported d=%d heads=%d", p.d, p.n_head);
    return -1;
  }
  const int NJ = (p.d / 8 + 31) / 32;
  const size_t smem = (size_t)(16 * p.n_head + 8 * p.d) * 4;
  dim3 grid(p.n_split, p.Mb);
  cudaError_t le = cudaSuccess;
#define WB_AD_CASE(J, R)                                                                                          \
  case J: {                                                                                                       \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      WB_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel<J, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024)); \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    le = launch_pdl(attn_decode_kernel<J, R>, grid, dim3(kAdThreads), smem, st, p);                               \
  } break;
  switch (NJ) {
    WB_AD_CASE(1, 4) WB_AD_CASE(2, 4) WB_AD_CASE(3, 2) WB_AD_CASE(4, 2) WB_AD_CASE(5, 2)
    default:
      set_error("attn_decode: width too large");
      return -1;
  }
#undef WB_AD_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// ---- end of step: sample (optional), embed the next token, advance -------------------------------------------------------------
// Upstream DecodingTask with temperature 0: the logit filters were applied in the logits GEMM epilogue;
// here argmax over the CTA partials, sum_logprobs += logprob * (previous token != eot), sequences that ended keep
// emitting eot. Then x[b] = token_embedding[next] + positional_embedding[cur_len + 1] for the next step.
__global__ void __launch_bounds__(256) step_finish_kernel(FinishDesc p) {
  __shared__ float s_val[8], s_sum[8];
  __shared__ int s_idx[8];
  __shared__ int s_tok;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  const int cur = ld_state(&p.state->cur_len);   // index of the token this step consumed (-1 before the first step)
  int32_t* trow = p.tokens + (size_t)b * p.tokens_ld;
  if (p.sample) {
    const float* pp = p.part_logits + (size_t)b * p.n_part * 4;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int i = tid; i < p.n_part; i += 256) {
      const float v = __ldcg(pp + i * 4);
      const int a = __float_as_int(__ldcg(pp + i * 4 + 1));
      if (v > best || (v == best && a < arg)) best = v, arg = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ov > best || (ov == best && oi < arg)) best = ov, arg = oi;
    }
    if (lane == 0) s_val[warp] = best, s_idx[warp] = arg;
    __syncthreads();
    best = s_val[0], arg = s_idx[0];
#pragma unroll
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > best || (s_val[w] == best && s_idx[w] < arg)) best = s_val[w], arg = s_idx[w];
    float se = 0.f;
    for (int i = tid; i < p.n_part; i += 256) {
      const float mi = __ldcg(pp + i * 4);
      if (mi > -INFINITY) se += __ldcg(pp + i * 4 + 2) * expf(mi - best);
    }
    se = warp_sum(se);
    if (lane == 0) s_sum[warp] = se;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += s_sum[w];
      const float logprob = -logf(tot);   // the chosen logit is the maximum
      const bool ended = __ldcg(trow + cur) == p.eot;
      if (!ended) p.sum_logprob[b] = __ldcg(p.sum_logprob + b) + logprob;
      const int next = ended ? p.eot : arg;
      trow[cur + 1] = next;
      p.done[b] = next == p.eot;
      s_tok = next;
    }
  } else if (tid == 0) {
    s_tok = __ldcg(trow + cur + 1);
  }
  __syncthreads();
  const int np = cur + 1;
  if (np < p.n_ctx) {
    int tok = s_tok;
    tok = tok < 0 ? 0 : (tok >= p.V ? p.V - 1 : tok);
    const __half* e = p.tok_emb + (size_t)tok * p.d;
    const float* pe = p.pos_emb + (size_t)np * p.d;
    for (int c = tid; c < p.d; c += 256) p.x[(size_t)b * p.d + c] = __half2float(e[c]) + pe[c];
  }
  // every CTA has read cur_len by now only once all have arrived: the last one advances it
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int ticket = atomicAdd(&p.state->arrive, 1);
    if (ticket == (int)gridDim.x - 1) {
      p.state->arrive = 0;
      p.state->cur_len = cur + 1;
    }
  }
}

int launch_step_finish(const FinishDesc& d, cudaStream_t st, int64_t* launches) {
  const cudaError_t le = launch_pdl(step_finish_kernel, dim3(d.Mb), dim3(256), 0, st, d);
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

}  // namespace wb
