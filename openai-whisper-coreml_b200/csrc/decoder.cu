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
//   attn_decode_head_kernel  one CTA per (sequence, head): K/V streamed by TMA through a shared-memory ring, scores and
//                         P*V on mma.sync, fp32 online softmax; no row split, so no partials and no merge
//   step_finish_kernel    merges the per-CTA (max, argmax, sum-exp) partials the logits GEMM epilogue produced (logit
//                         filters already applied there), EOT forcing, log-prob accumulation, token append, then embeds
//                         the next token (+ learned position) into the residual stream and advances the position
#include <cuda.h>

#include "ops.cuh"
#include "ptx.cuh"

namespace wb {

constexpr float kLog2e = 1.44269504088896340736f;

// L2 residency policy of the decode step (ptx.cuh): 1 = weights evict_last + cross-attention K/V stream evict_first, 0 = no
// hints. Set per device by decoder_set_l2_mode() (env WB_L2_POLICY, default 1).
__constant__ int c_l2_mode = 1;
__device__ __forceinline__ uint64_t weight_policy() { return ptx::l2_policy(c_l2_mode ? 1 : 0); }
__device__ __forceinline__ uint64_t stream_policy() { return ptx::l2_policy(c_l2_mode ? 2 : 0); }
__device__ __forceinline__ void weight_prefetch_l2(const void* p) {
  if (c_l2_mode)
    ptx::prefetch_l2_evict_last(p);
  else
    ptx::prefetch_l2(p);
}
// The embedding matrix (53 MB at base.en) is streamed once per step by the logits kernel and does not survive the 590 MB
// cross K/V stream of the next step whatever its policy (in-step ncu: 52.7 MB of DRAM reads per launch either way), but
// loaded evict_last it pushes the weights of decoder layers 0-1 out of L2 (2-4 MB of DRAM reads in their block kernels).
// WB_EMB_KEEP8 = eighths of its row groups loaded evict_last, the rest evict_first like the K/V stream; default 0
// (measured 61.74 -> 61.59 ms per 225-step decode).
__constant__ int c_emb_keep8 = 0;
int decoder_set_l2_mode(int mode) {
  WB_CUDA_OK(cudaMemcpyToSymbol(c_l2_mode, &mode, sizeof(int)));
  int keep = 0;
  if (const char* e = getenv("WB_EMB_KEEP8")) keep = atoi(e);
  keep = keep < 0 ? 0 : (keep > 8 ? 8 : keep);
  WB_CUDA_OK(cudaMemcpyToSymbol(c_emb_keep8, &keep, sizeof(int)));
  return 0;
}

// Everything that changes from kernel to kernel inside a decode step (residual stream, q, attention outputs, partials,
// tokens, DecodeState) is read through L2 (.cg): with programmatic dependent launch a kernel can share an SM — and its L1 —
// with its still-running predecessor, so L1 may hold lines the predecessor fetched before another SM rewrote them.
// Weights and the cross-attention K/V are constant during a decode and use the non-coherent path.
__device__ __forceinline__ int ld_state(const int* p) { return __ldcg(p); }

// Development tracing: block (0,0) thread 0 of every decode kernel stamps %globaltimer at entry and exit (WB_TRACE=1).
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct TraceScope {
  unsigned long long* rec = nullptr;
  __device__ __forceinline__ TraceScope(const DecodeState* st, int id) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && st) {
      unsigned long long* tr = st->trace;
      if (tr) {
        const int i = atomicAdd(const_cast<int*>(&st->trace_n), 1);
        if (i < 16384) {
          rec = tr + (size_t)i * 8;
          rec[0] = (unsigned long long)id;
          rec[1] = globaltimer();
        }
      }
    }
  }
  __device__ __forceinline__ void end() {
    if (rec) rec[2] = globaltimer();
  }
  __device__ __forceinline__ void mark(int k) {
    if (rec) rec[k] = globaltimer();
  }
};

// ---- skinny GEMM -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_x4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
constexpr int kSkThreads = 256;
constexpr int kSkKC = 2048;   // activation columns staged in shared memory at a time

struct SkinnyArgs {
  SkinnyDesc d;
  int strips_per_cta;   // 1, 2, 4 or 8 strips of 16 weight rows; the 8 warps split K 8/strips ways
  int kc;               // activation columns staged at a time (multiple of 256): the whole K where Mb x K fits shared memory
};

template <int MT>
__global__ void __launch_bounds__(kSkThreads) skinny_gemm_kernel(SkinnyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SkinnyDesc& p = a.d;
  TraceScope trace(p.state, 100 + p.out_mode * 10 + p.in_mode);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3;
  const int S = a.strips_per_cta, KS = 8 / S;
  const int strip = warp % S, kslice = warp / S;
  const int KC = p.K < a.kc ? p.K : a.kc;
  const int xs_stride = KC * 2 + 64;                       // bytes; (stride/16) % 8 == 4 -> conflict-free LDS.128
  unsigned char* xs = smem_raw;
  constexpr int RS = MT * 8 + 1;                                                  // padded row stride of red: conflict-free epilogue
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)MT * 8 * xs_stride);   // [8 warps][16][RS]
  float* sbias = red + 8 * 16 * RS;                                           // [128] bias of this CTA's rows
  float* sg = sbias + 128;                                                        // [K] LayerNorm gamma (LN input mode)
  float* sb = sg + p.K;                                                           // [K] LayerNorm beta

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
  constexpr int kPre = 8;
  const uint64_t wpol = weight_policy();
  uint4 pwa[kPre], pwb[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int blk = (pb0 + u < pb1) ? pb0 + u : pb0;
    pwa[u] = ptx::ldg_nc_16(wrow0 + blk * 32, wpol);
    pwb[u] = ptx::ldg_nc_16(wrow1 + blk * 32, wpol);
  }
  // constants (bias, LayerNorm affine) are staged in shared memory before the wait as well: keeping their global loads
  // out of the input stage and of the epilogue matters — interleaved with shared-memory stores the compiler cannot batch
  // them, and each one cost an L2 round trip (the LN input stage measured 8-10 us before this, ~2 us after)
  for (int i = tid; i < S * 16; i += kSkThreads) sbias[i] = (p.bias && n_cta + i < p.N) ? __ldg(p.bias + n_cta + i) : 0.f;
  if (p.in_mode == SKINNY_IN_LN) {
    for (int i = tid * 4; i < p.K; i += kSkThreads * 4) {
      *reinterpret_cast<float4*>(sg + i) = __ldg(reinterpret_cast<const float4*>(p.ln_g + i));
      *reinterpret_cast<float4*>(sb + i) = __ldg(reinterpret_cast<const float4*>(p.ln_b + i));
    }
  }
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  __syncthreads();   // the staged constants are read by other threads in the input stage
  trace.mark(3);
  const int pos = (p.out_mode == SKINNY_OUT_QKV) ? ld_state(&p.state->cur_len) : 0;   // cache row of this step (prefetched)
  // residual-add epilogue: fetch the old values now, off the critical path (N = d: 16 rows x Mb <= 640 values per CTA)
  float resid[3] = {0.f, 0.f, 0.f};
  const bool resid_pre = p.out_mode == SKINNY_OUT_RESID && S == 1;
  if (resid_pre) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int idx = tid + j * kSkThreads;
      const int b = idx >> 4, n = n_cta + (idx & 15);
      if (b < p.Mb && n < p.N) resid[j] = __ldcg(reinterpret_cast<const float*>(p.out) + (size_t)b * p.N + n);
    }
  }

  for (int kc0 = 0; kc0 < p.K; kc0 += KC) {
    const int kc = (p.K - kc0) < KC ? (p.K - kc0) : KC;
    if (kc0) __syncthreads();
    // ---- input stage: build xs[MT*8][kc] fp16 --------------------------------------------------------------------------
    if (p.in_mode == SKINNY_IN_F16) {
      // warp w copies rows w, w+8, ...; lanes stride over the 16-byte chunks of a row; MT x 4 independent loads in flight
      const __half* src = reinterpret_cast<const __half*>(p.in);
      const int cpr = kc >> 3;                              // 16-byte chunks per row
      for (int c0 = lane; c0 < cpr; c0 += 128) {
        uint4 v[MT][4];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int r = warp + 8 * j;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + 32 * u;
            v[j][u] = make_uint4(0, 0, 0, 0);
            if (c < cpr && r < p.Mb) v[j][u] = ptx::ldg_nc_16(src + (size_t)r * p.K + kc0 + c * 8);
          }
        }
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int r = warp + 8 * j;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + 32 * u;
            if (c < cpr) *reinterpret_cast<uint4*>(xs + (size_t)r * xs_stride + c * 16) = v[j][u];
          }
        }
      }
    } else {
      // LayerNorm of the fp32 residual stream (K = d, one chunk). 8 threads per row, 32 rows per pass: every row of the
      // batch is in flight at once; two passes over L2 (statistics, then normalise). fp32, biased variance, eps 1e-5.
      const int sub = tid & 7;
      for (int r = tid >> 3; r < MT * 8; r += kSkThreads / 8) {     // trip count is warp-uniform (MT*8 is a multiple of 8)
        __half* xr = reinterpret_cast<__half*>(xs + (size_t)r * xs_stride);
        const bool act = r < p.Mb;                                   // padding rows compute on row 0 and store zeros
        const float* src = reinterpret_cast<const float*>(p.in) + (size_t)(act ? r : 0) * p.K;
        // statistics: batches of 16 independent float4 loads per thread (one batch covers K = 512)
        float s = 0.f, q = 0.f;
        float4 v[16];
        for (int c0 = 0; c0 < p.K; c0 += 512) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + (sub + 8 * i) * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < p.K) v[i] = ld_x4(src + c);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
          }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float mean = s / (float)p.K;
        const float var = fmaxf(q / (float)p.K - mean * mean, 0.f);
        const float rstd = act ? rsqrtf(var + 1e-5f) : 0.f;
        const float ab = act ? 1.f : 0.f;
        for (int c0 = 0; c0 < p.K; c0 += 512) {
          if (p.K > 512) {   // rows wider than one batch: reload (L2)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = c0 + (sub + 8 * i) * 4;
              if (c < p.K) v[i] = ld_x4(src + c);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + (sub + 8 * i) * 4;
            if (c < p.K) {
              const float4 g = *reinterpret_cast<const float4*>(sg + c), bb = *reinterpret_cast<const float4*>(sb + c);
              __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + ab * bb.x, (v[i].y - mean) * rstd * g.y + ab * bb.y);
              __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + ab * bb.z, (v[i].w - mean) * rstd * g.w + ab * bb.w);
              uint2 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(xr + c) = u;
            }
          }
        }
      }
    }
    __syncthreads();
    trace.mark(4);
    // ---- stream this warp's weight rows over its K slice of the chunk ------------------------------------------------------
    const int nblk = kc / 32;
    const int blk0 = (kslice * nblk) / KS, blk1 = ((kslice + 1) * nblk) / KS;
    const unsigned char* xl = xs + (size_t)grp * xs_stride + tq * 16;
    int blk = blk0;
    if (kc0 == 0) {   // the prefetched blocks
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
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
      blk = blk0 + kPre < blk1 ? blk0 + kPre : blk1;
    }
    for (; blk + 8 <= blk1; blk += 8) {
      uint4 wa[8], wb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        wa[u] = ptx::ldg_nc_16(wrow0 + kc0 + (blk + u) * 32, wpol);
        wb[u] = ptx::ldg_nc_16(wrow1 + kc0 + (blk + u) * 32, wpol);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
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
    if (blk < blk1) {   // remainder (< 8 blocks): still one batch of independent loads
      uint4 wa[8], wb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int bb = (blk + u < blk1) ? blk + u : blk;
        wa[u] = ptx::ldg_nc_16(wrow0 + kc0 + bb * 32, wpol);
        wb[u] = ptx::ldg_nc_16(wrow1 + kc0 + bb * 32, wpol);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (blk + u < blk1) {
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
    }
  }
  // ---- cross-warp (K split) reduction and epilogue -------------------------------------------------------------------------
  trace.mark(5);
  float* myred = red + (size_t)warp * 16 * RS;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    myred[grp * RS + mt * 8 + 2 * tq] = acc[mt][0];
    myred[grp * RS + mt * 8 + 2 * tq + 1] = acc[mt][1];
    myred[(grp + 8) * RS + mt * 8 + 2 * tq] = acc[mt][2];
    myred[(grp + 8) * RS + mt * 8 + 2 * tq + 1] = acc[mt][3];
  }
  __syncthreads();
  const int rows_cta = S * 16;
  const int dq = p.N / 3;
  if (resid_pre) {   // rows_cta == 16; same index map as the prefetch above
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int idx = tid + j * kSkThreads;
      const int b = idx >> 4, r = idx & 15, n = n_cta + r;
      if (b < p.Mb && n < p.N) {
        float v = 0.f;
        for (int ks = 0; ks < 8; ++ks) v += red[(size_t)ks * 16 * RS + r * RS + b];
        v += sbias[r];
        if (p.gelu) v = gelu_erf(v);
        reinterpret_cast<float*>(p.out)[(size_t)b * p.N + n] = resid[j] + v;
      }
    }
    trace.end();
    return;
  }
  for (int idx = tid; idx < rows_cta * p.Mb; idx += kSkThreads) {
    const int b = idx / rows_cta, rr = idx - b * rows_cta;
    const int st = rr >> 4, r = rr & 15;
    const int n = n_cta + rr;
    if (n >= p.N) continue;
    float v = 0.f;
    for (int ks = 0; ks < KS; ++ks) v += red[(size_t)(ks * S + st) * 16 * RS + r * RS + b];
    v += sbias[rr];
    if (p.gelu) v = gelu_erf(v);
    switch (p.out_mode) {
      case SKINNY_OUT_F16:
        reinterpret_cast<__half*>(p.out)[(size_t)b * p.N + n] = __float2half_rn(v);
        break;
      case SKINNY_OUT_F32:
        reinterpret_cast<float*>(p.out)[(size_t)b * p.N + n] = v;
        break;
      case SKINNY_OUT_RESID: {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)b * p.N + n;
        *o = __ldcg(o) + v;
      } break;
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
  trace.end();
}

// ---- logits GEMM: final LayerNorm + tied-embedding projection + logit filters + per-group (max, argmax, sum-exp) -----------------
// Persistent: one CTA per SM normalises the Mb rows once, then loops over 128-row groups of the [V][d] embedding matrix
// (the only weight-streaming GEMM of the step that is bandwidth- rather than latency-bound: V*d*2 = 53 MB for base.en).
// Warp w owns rows 16w..16w+15 of a group over the whole K; the next group's first blocks are requested before the
// epilogue of the current one.
template <int MT>
__global__ void __launch_bounds__(kSkThreads) logits_gemm_kernel(SkinnyDesc p, int n_groups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TraceScope trace(p.state, 141);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3;
  const int xs_stride = p.K * 2 + 64;
  unsigned char* xs = smem_raw;
  constexpr int MB8 = MT * 8;
  constexpr int RS = MB8 + 1;                                                  // padded: conflict-free transposed reads
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)MB8 * xs_stride);   // [128][RS]
  float* sg = red + 128 * RS;
  float* sb = sg + p.K;
  const int nblk = p.K / 32;
  constexpr int kPre = 8;

  auto row_ptr = [&](int g, int half) {
    int n = g * 128 + warp * 16 + grp + half * 8;
    n = n < p.N ? n : p.N - 1;
    return p.w + (size_t)n * p.K + tq * 8;
  };
  int g = blockIdx.x;
  const __half* wrow0 = row_ptr(g, 0);
  const __half* wrow1 = row_ptr(g, 1);
  const uint64_t wpol = weight_policy();
  uint4 pwa[kPre], pwb[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int blk = u < nblk ? u : 0;
    pwa[u] = ptx::ldg_nc_16(wrow0 + blk * 32, wpol);
    pwb[u] = ptx::ldg_nc_16(wrow1 + blk * 32, wpol);
  }
  for (int i = tid * 4; i < p.K; i += kSkThreads * 4) {
    *reinterpret_cast<float4*>(sg + i) = __ldg(reinterpret_cast<const float4*>(p.ln_g + i));
    *reinterpret_cast<float4*>(sb + i) = __ldg(reinterpret_cast<const float4*>(p.ln_b + i));
  }
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  __syncthreads();
  trace.mark(3);
  {   // LayerNorm of the residual stream into xs (same scheme as skinny_gemm_kernel)
    const int sub = tid & 7;
    for (int r = tid >> 3; r < MB8; r += kSkThreads / 8) {
      __half* xr = reinterpret_cast<__half*>(xs + (size_t)r * xs_stride);
      const bool act = r < p.Mb;
      const float* src = reinterpret_cast<const float*>(p.in) + (size_t)(act ? r : 0) * p.K;
      float s = 0.f, q = 0.f;
      float4 v[16];
      for (int c0 = 0; c0 < p.K; c0 += 512) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + (sub + 8 * i) * 4;
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < p.K) v[i] = ld_x4(src + c);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
          q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      q += __shfl_xor_sync(0xffffffffu, q, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      q += __shfl_xor_sync(0xffffffffu, q, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      q += __shfl_xor_sync(0xffffffffu, q, 4);
      const float mean = s / (float)p.K;
      const float var = fmaxf(q / (float)p.K - mean * mean, 0.f);
      const float rstd = act ? rsqrtf(var + 1e-5f) : 0.f;
      const float ab = act ? 1.f : 0.f;
      for (int c0 = 0; c0 < p.K; c0 += 512) {
        if (p.K > 512) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + (sub + 8 * i) * 4;
            if (c < p.K) v[i] = ld_x4(src + c);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + (sub + 8 * i) * 4;
          if (c < p.K) {
            const float4 gm = *reinterpret_cast<const float4*>(sg + c), bb = *reinterpret_cast<const float4*>(sb + c);
            __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * gm.x + ab * bb.x, (v[i].y - mean) * rstd * gm.y + ab * bb.y);
            __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * gm.z + ab * bb.z, (v[i].w - mean) * rstd * gm.w + ab * bb.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(xr + c) = u;
          }
        }
      }
    }
  }
  __syncthreads();
  trace.mark(4);
  const bool first = ld_state(&p.state->cur_len) + 1 == p.n_initial;
  const unsigned char* xl = xs + (size_t)grp * xs_stride + tq * 16;
  while (true) {
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
    unsigned char mk4[4];   // logit-filter classes of this lane's 4 rows of the group: fetched now, used in the epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = g * 128 + lane + 32 * i;
      mk4[i] = (p.mask && n < p.N) ? __ldg(p.mask + n) : 0;
    }
    auto mma_block = [&](const uint4& wa, const uint4& wb, int blk) {
      const uint32_t a0[4] = {wa.x, wb.x, wa.y, wb.y}, a1[4] = {wa.z, wb.z, wa.w, wb.w};
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint4 xb = *reinterpret_cast<const uint4*>(xl + (size_t)mt * 8 * xs_stride + blk * 64);
        const uint32_t b0[2] = {xb.x, xb.y}, b1[2] = {xb.z, xb.w};
        ptx::mma_16816(acc[mt], a0, b0);
        ptx::mma_16816(acc[mt], a1, b1);
      }
    };
#pragma unroll
    for (int u = 0; u < kPre; ++u)
      if (u < nblk) mma_block(pwa[u], pwb[u], u);
    for (int blk = kPre; blk < nblk; blk += 8) {
      uint4 wa[8], wb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int bb = blk + u < nblk ? blk + u : blk;
        wa[u] = ptx::ldg_nc_16(wrow0 + bb * 32, wpol);
        wb[u] = ptx::ldg_nc_16(wrow1 + bb * 32, wpol);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (blk + u < nblk) mma_block(wa[u], wb[u], blk + u);
    }
    const int g_next = g + gridDim.x;
    if (g_next < n_groups) {   // request the next group's first blocks before the epilogue
      wrow0 = row_ptr(g_next, 0);
      wrow1 = row_ptr(g_next, 1);
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int blk = u < nblk ? u : 0;
        pwa[u] = ptx::ldg_nc_16(wrow0 + blk * 32, wpol);
        pwb[u] = ptx::ldg_nc_16(wrow1 + blk * 32, wpol);
      }
    }
    // accumulators -> red[row in group][sequence]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      red[(warp * 16 + grp) * RS + mt * 8 + 2 * tq] = acc[mt][0];
      red[(warp * 16 + grp) * RS + mt * 8 + 2 * tq + 1] = acc[mt][1];
      red[(warp * 16 + grp + 8) * RS + mt * 8 + 2 * tq] = acc[mt][2];
      red[(warp * 16 + grp + 8) * RS + mt * 8 + 2 * tq + 1] = acc[mt][3];
    }
    __syncthreads();
    // warp per sequence: filter, optional store, group-local (max, argmax, sum-exp)
    for (int b = warp; b < p.Mb; b += 8) {
      float v[4];
      float best = -INFINITY;
      int arg = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = lane + 32 * i, n = g * 128 + rr;
        float x = -INFINITY;
        if (n < p.N) {
          x = red[rr * RS + b];
          const unsigned char mk = mk4[i];
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
      if (lane == 0) *reinterpret_cast<float4*>(p.part_logits + ((size_t)b * n_groups + g) * 4) = make_float4(best, __int_as_float(arg), se, 0.f);
    }
    if (g_next >= n_groups) break;
    g = g_next;
    __syncthreads();   // red is rewritten by the next group
  }
  trace.end();
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
template <typename Kern, typename... Args>
static cudaError_t launch_pdl_n(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
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

int skinny_logits_ctas(int N) { return ((N + 15) / 16 + 7) / 8; }   // 128-row groups (= partials per sequence) of the LOGITS mode
int logits_groups(int Mb, int N, int K) {
  (void)Mb, (void)K;
  return skinny_logits_ctas(N);
}

// ---- logits on the tensor cores: final LayerNorm + tied-embedding GEMM, swap-AB, tcgen05 + TMA ---------------------------------------
// logits[b][n] = LN(x[b]) . E[n]: the vocabulary is the M dimension (128 embedding rows per tile = one partial group), the
// sequences are N (NB = 16 / 32 / 48 columns). The [V][d] matrix is the only operand of the step that is bandwidth-bound
// (V*d*2 = 53 MB for base.en): it streams through a TMA ring (128 rows x 64 k per stage, 128-byte swizzle, rows past V
// zero-filled) that starts filling before the previous kernel has finished; the normalised activations are written once into
// shared memory in the same swizzled K-major layout; one thread issues tcgen05.mma (M128 N=NB K16) into a double-buffered
// TMEM accumulator; eight epilogue warps read it back (one vocabulary row per thread, half of the columns per warp), apply the
// logit filters, and reduce (max, argmax, sum-exp) per sequence over the 128 rows through a padded shared-memory transpose
// (the sequences of a warp are unrolled so that their shuffle chains overlap). Persistent: one CTA per SM.
constexpr int kLtThreads = 320;   // warp 0: TMA producer, warp 1: TMEM + MMA issuer, warps 2-9: LayerNorm, then epilogue
constexpr int kLtStageBytes = 128 * 64 * 2;

struct LogitsTcArgs {
  SkinnyDesc p;
  int n_groups, n_stages;
};

template <int NB, bool TS>   // TS: upstream ApplyTimestampRules among the logit filters (compiled out otherwise)
__global__ void __launch_bounds__(kLtThreads, 1) logits_tc_kernel(const __grid_constant__ CUtensorMap tmW, LogitsTcArgs a) {
  constexpr int kBufStride = NB <= 32 ? 32 : 64;               // TMEM columns per accumulator buffer
  constexpr int kTmemCols = 2 * kBufStride;
  constexpr int RS = NB + 1;                                   // padded row of the transpose buffer
  extern __shared__ unsigned char lt_dyn[];
  const SkinnyDesc& p = a.p;
  TraceScope trace(p.state, 142);
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lt_dyn) + 1023) & ~(uintptr_t)1023);
  const int nkb = p.K >> 6, n_stages = a.n_stages;
  unsigned char* ring = smem;                                            // [n_stages][128 rows][128 B]
  unsigned char* xb = ring + (size_t)n_stages * kLtStageBytes;           // [nkb][NB rows][128 B] LayerNorm(x), swizzled
  float* red = reinterpret_cast<float*>(xb + (size_t)nkb * NB * 128);    // [128][RS]
  float* sg = red + 128 * RS;                                            // [K] LayerNorm gamma, beta
  float* sb = sg + p.K;
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + p.K);                // [n_stages]
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;                                           // [2]
  uint64_t* tempty = tfull + 2;                                          // [2]
  uint64_t* bready = tempty + 2;                                         // activations written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bready + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmW);
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 8);
    }
    ptx::mbar_init(bready, 256);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ---- TMA producer: the embedding matrix does not depend on the previous kernel, so the ring fills during its tail -----------------
    if (lane == 0) {
      const uint64_t pol_keep = weight_policy(), pol_stream = stream_policy();
      const int keep_groups = (a.n_groups * c_emb_keep8) >> 3;
      uint32_t it = 0;
      for (int g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
        const uint64_t wpol = g < keep_groups ? pol_keep : pol_stream;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (it / n_stages) & 1u;
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], kLtStageBytes);
          ptx::tma_load_3d(ring + (size_t)s * kLtStageBytes, &tmW, &full[s], kb * 64, g * 128, 0, wpol);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(128, NB);
      ptx::mbar_wait(bready, 0);
      ptx::tc_fence_after();
      const uint32_t xb_addr = ptx::smem_u32(xb);
      uint32_t it = 0, lt = 0;
      for (int g = blockIdx.x; g < a.n_groups; g += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1u, tph = (lt >> 1) & 1u;
        ptx::mbar_wait(&tempty[buf], tph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + buf * kBufStride;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (it / n_stages) & 1u;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(ring + (size_t)s * kLtStageBytes));
          const uint64_t bdesc = ptx::umma_desc_sw128_kmajor(xb_addr + (uint32_t)kb * NB * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty[s]);
        }
        ptx::umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ---- warps 2-9: LayerNorm of the residual stream into the swizzled B tiles, then the epilogue of every group -------------------------
    const int et = tid - 64, ewarp = warp - 2;                 // 0..255, 0..7
    for (int i = et * 4; i < p.K; i += 256 * 4) {
      *reinterpret_cast<float4*>(sg + i) = __ldg(reinterpret_cast<const float4*>(p.ln_g + i));
      *reinterpret_cast<float4*>(sb + i) = __ldg(reinterpret_cast<const float4*>(p.ln_b + i));
    }
    ptx::grid_dep_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    trace.mark(3);
    {
      const int sub = et & 7;
      for (int r = et >> 3; r < NB; r += 32) {                 // 8 threads per row, 32 rows per pass (trip count warp-uniform)
        const bool act = r < p.Mb;
        const float* src = reinterpret_cast<const float*>(p.in) + (size_t)(act ? r : 0) * p.K;
        float sm = 0.f, q = 0.f;
        float4 v[16];
        for (int c0 = 0; c0 < p.K; c0 += 512) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + (sub + 8 * i) * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < p.K) v[i] = ld_x4(src + c);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
          }
        }
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        sm += __shfl_xor_sync(0xffffffffu, sm, 4);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float mean = sm / (float)p.K;
        const float var = fmaxf(q / (float)p.K - mean * mean, 0.f);
        const float rstd = act ? rsqrtf(var + 1e-5f) : 0.f;
        const float ab = act ? 1.f : 0.f;
        for (int c0 = 0; c0 < p.K; c0 += 512) {
          if (p.K > 512) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = c0 + (sub + 8 * i) * 4;
              if (c < p.K) v[i] = ld_x4(src + c);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + (sub + 8 * i) * 4;
            if (c < p.K) {
              const float4 gm = *reinterpret_cast<const float4*>(sg + c), bb = *reinterpret_cast<const float4*>(sb + c);
              __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * gm.x + ab * bb.x, (v[i].y - mean) * rstd * gm.y + ab * bb.y);
              __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * gm.z + ab * bb.z, (v[i].w - mean) * rstd * gm.w + ab * bb.w);
              uint2 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0), u.y = *reinterpret_cast<uint32_t*>(&h1);
              // element (row r, column c) of the K-major SWIZZLE_128B tile kt = c / 64: 16-byte chunk index xor (r & 7)
              const int kt = c >> 6, kk = c & 63;
              unsigned char* dst = xb + (size_t)kt * NB * 128 + r * 128 + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1));
              *reinterpret_cast<uint2*>(dst) = u;
            }
          }
        }
      }
    }
    ptx::fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor-core (async) proxy
    ptx::mbar_arrive(bready);
    trace.mark(4);
    const bool first = ld_state(&p.state->cur_len) + 1 == p.n_initial;
    // timestamp rules: per-sequence state of this step (written by the previous finish kernel) into shared memory
    int4* s_ts = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(tmem_slot) + 31) & ~(uintptr_t)15);
    constexpr bool ts_on = TS;
    if (ts_on) {
      if (et < NB) s_ts[et] = et < p.Mb ? __ldcg(p.ts_state + et) : make_int4(0, 0, 0, 0);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const int g_straddle = (ts_on && (p.ts_begin & 127)) ? (p.ts_begin >> 7) : -1;   // group with text and timestamp rows
    const int q4 = warp & 3, half = ewarp >> 2;                // TMEM lane quadrant this warp may read; its half of the columns
    const int j_lo = half * (NB / 2), j_hi = j_lo + NB / 2;
    constexpr int kSeqPerWarp = (NB + 7) / 8;
    uint32_t lt = 0;
    for (int g = blockIdx.x; g < a.n_groups; g += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u, tph = (lt >> 1) & 1u;
      const int row = q4 * 32 + lane, n = g * 128 + row;
      const unsigned char mk = (p.mask && n < p.N) ? __ldg(p.mask + n) : 0;
      const bool dead = n >= p.N || mk == 1 || (mk == 2 && first);
      ptx::mbar_wait(&tfull[buf], tph);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + buf * kBufStride + ((uint32_t)(q4 * 32) << 16), v);
      ptx::tmem_ld_wait();
      const bool is_ts = n >= p.ts_begin, below_eot = n < p.eot, late = first && n > p.ts_last_allowed;
      auto dead_for = [&](int j) {   // static filters, then the timestamp rules of sequence j
        if (!ts_on) return dead;
        const int4 st = s_ts[j];
        const bool dyn = is_ts ? ((st.x & 1) || n < st.y || late) : (first || ((st.x & 2) && below_eot));
        return dead || dyn;
      };
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < NB && j >= j_lo && j < j_hi) red[row * RS + j] = dead_for(j) ? -INFINITY : __uint_as_float(v[j]);
      if (NB > 32) {
        ptx::tmem_ld_32x32(tmem_base + buf * kBufStride + ((uint32_t)(q4 * 32) << 16) + 32u, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 + j < NB && 32 + j >= j_lo && 32 + j < j_hi) red[row * RS + 32 + j] = dead_for(32 + j) ? -INFINITY : __uint_as_float(v[j]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[buf]);           // the accumulator buffer is free for the group after next
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // warp per sequence (b = ewarp, ewarp + 8, ...): optional store, group-local (max, argmax, sum-exp); unrolled over the
      // sequences of this warp so that the independent shuffle chains overlap. `cls`: 0 = every row of the group,
      // 1 / 2 = only its text / timestamp rows (the group that straddles ts_begin is reduced once per class)
      auto reduce_group = [&](int cls, float* dst_base, int dst_stride, int dst_index) {
        float xv[kSeqPerWarp][4], best[kSeqPerWarp], se[kSeqPerWarp];
        int arg[kSeqPerWarp];
#pragma unroll
        for (int jb = 0; jb < kSeqPerWarp; ++jb) {
          const int b = ewarp + 8 * jb;
          best[jb] = -INFINITY, arg[jb] = 0x7fffffff;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = lane + 32 * i, nn = g * 128 + rr;
            float x = -INFINITY;
            if (nn < p.N && b < p.Mb) {
              x = red[rr * RS + b];
              if (cls != 2 && p.out) reinterpret_cast<float*>(p.out)[(size_t)b * p.N + nn] = x;
              if ((cls == 1 && nn >= p.ts_begin) || (cls == 2 && nn < p.ts_begin)) x = -INFINITY;
            }
            xv[jb][i] = x;
            if (x > best[jb]) best[jb] = x, arg[jb] = nn;      // ascending n: first maximum wins
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int jb = 0; jb < kSeqPerWarp; ++jb) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[jb], o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg[jb], o);
            if (ov > best[jb] || (ov == best[jb] && oi < arg[jb])) best[jb] = ov, arg[jb] = oi;
          }
        }
#pragma unroll
        for (int jb = 0; jb < kSeqPerWarp; ++jb) {
          se[jb] = 0.f;
          if (best[jb] > -INFINITY) {
#pragma unroll
            for (int i = 0; i < 4; ++i) se[jb] += expf(xv[jb][i] - best[jb]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int jb = 0; jb < kSeqPerWarp; ++jb) se[jb] += __shfl_xor_sync(0xffffffffu, se[jb], o);
        }
#pragma unroll
        for (int jb = 0; jb < kSeqPerWarp; ++jb) {
          const int b = ewarp + 8 * jb;
          if (lane == 0 && b < p.Mb)
            *reinterpret_cast<float4*>(dst_base + ((size_t)b * dst_stride + dst_index) * 4) = make_float4(best[jb], __int_as_float(arg[jb]), se[jb], 0.f);
        }
      };
      if (g == g_straddle) {
        reduce_group(1, p.part_logits, a.n_groups, g);
        reduce_group(2, p.part_extra, 1, 0);
      } else {
        reduce_group(0, p.part_logits, a.n_groups, g);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // red is rewritten by the next group
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  trace.end();
}

static int launch_logits_tc(const SkinnyDesc& d, cudaStream_t st, int64_t* launches) {
  const int NB = d.Mb <= 16 ? 16 : (d.Mb <= 32 ? 32 : (d.Mb <= 48 ? 48 : 64));
  const int nkb = d.K / 64, n_groups = skinny_logits_ctas(d.N);
  CUtensorMap tmW;
  const int rc = gemm_get_tmap(d.tmaps, d.w, d.K, d.N, 1, d.K, (long long)d.N * d.K, 128, &tmW);
  if (rc) return rc;
  const size_t fixed = (size_t)nkb * NB * 128 + (size_t)128 * (NB + 1) * 4 + (size_t)2 * d.K * 4 + 1280 + 1024;   // + barriers and rule states + alignment slack
  int n_stages = (int)(((size_t)200 * 1024 - fixed) / kLtStageBytes);
  n_stages = n_stages > 8 ? 8 : n_stages;
  if (n_stages < 2) {
    set_error("logits GEMM: K=%d too wide for the shared-memory ring", d.K);
    return -1;
  }
  const size_t smem = fixed + (size_t)n_stages * kLtStageBytes;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  LogitsTcArgs a{d, n_groups, n_stages};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_groups < n_sm ? n_groups : n_sm), cfg.blockDim = dim3(kLtThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = use_pdl() ? 1 : 0;
  cudaError_t le = cudaSuccess;
#define WB_LT_CASE(N_)                                                                                                 \
  case N_: {                                                                                                           \
    static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];                                                                                        \
    if (smem > smem_set) {                                                                                             \
      WB_CUDA_OK(cudaFuncSetAttribute(logits_tc_kernel<N_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      WB_CUDA_OK(cudaFuncSetAttribute(logits_tc_kernel<N_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                                 \
    }                                                                                                                  \
    le = d.ts_state ? cudaLaunchKernelEx(&cfg, logits_tc_kernel<N_, true>, tmW, a)                                     \
                    : cudaLaunchKernelEx(&cfg, logits_tc_kernel<N_, false>, tmW, a);                                   \
  } break;
  switch (NB) {
    WB_LT_CASE(16) WB_LT_CASE(32) WB_LT_CASE(48) WB_LT_CASE(64)
  }
#undef WB_LT_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

static bool use_logits_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WB_LOGITS_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_skinny_gemm(const SkinnyDesc& d, cudaStream_t st, int64_t* launches) {
  const bool to_tc = d.out_mode == SKINNY_OUT_LOGITS && d.in_mode == SKINNY_IN_LN && d.K <= kSkKC && use_logits_tc() && d.tmaps && d.K % 64 == 0;
  if (d.Mb < 1 || d.Mb > (to_tc ? kMaxSequences : kMaxSequencesWide) || d.K % 128 != 0 || d.N < 16) {
    set_error("skinny_gemm: unsupported shape Mb=%d N=%d K=%d", d.Mb, d.N, d.K);
    return -1;
  }
  if (d.in_mode == SKINNY_IN_LN && d.K > kSkKC) {
    set_error("skinny_gemm: fused LayerNorm input needs K <= %d", kSkKC);
    return -1;
  }
  if (d.out_mode == SKINNY_OUT_LOGITS) {
    if (d.in_mode != SKINNY_IN_LN || d.K > kSkKC) {
      set_error("logits GEMM: LayerNorm input with K <= %d required", kSkKC);
      return -1;
    }
    if (use_logits_tc() && d.tmaps && d.K % 64 == 0 && d.Mb <= 64) return launch_logits_tc(d, st, launches);
    if (d.ts_state) {
      set_error("logits GEMM: the timestamp rules are implemented by the tcgen05 kernel only (Mb <= 64, K %% 64 == 0)");
      return -1;
    }
    const int n_groups = skinny_logits_ctas(d.N);
    const int MTl = (d.Mb + 7) / 8;
    const size_t sm = (size_t)MTl * 8 * (d.K * 2 + 64) + (size_t)128 * (MTl * 8 + 1) * 4 + 16 + (size_t)2 * d.K * 4;
    static int n_sm = 0;
    if (!n_sm) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      if (n_sm <= 0) n_sm = 148;
    }
    const int grid = n_groups < n_sm ? n_groups : n_sm;
    cudaError_t le2 = cudaSuccess;
#define WB_LG_CASE(M)                                                                                                 \
  case M: {                                                                                                           \
    static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];                                                                                       \
    if (sm > smem_set) {                                                                                              \
      WB_CUDA_OK(cudaFuncSetAttribute(logits_gemm_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      smem_set = sm;                                                                                                  \
    }                                                                                                                 \
    cudaLaunchConfig_t cfg{};                                                                                         \
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kSkThreads), cfg.dynamicSmemBytes = sm, cfg.stream = st;           \
    cudaLaunchAttribute at[1];                                                                                        \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                    \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                                             \
    cfg.attrs = at, cfg.numAttrs = use_pdl() ? 1 : 0;                                                                 \
    le2 = cudaLaunchKernelEx(&cfg, logits_gemm_kernel<M>, d, n_groups);                                               \
  } break;
    switch (MTl) {
      WB_LG_CASE(1) WB_LG_CASE(2) WB_LG_CASE(3) WB_LG_CASE(4) WB_LG_CASE(5)
      default:
        set_error("logits GEMM: Mb too large");
        return -1;
    }
#undef WB_LG_CASE
    if (launches) *launches += 1;
    WB_CUDA_OK(le2);
    return 0;
  }
  const int strips = (d.N + 15) / 16;
  int S = 1;
  while (S < 8 && strips / S > 296) S *= 2;
  const int MT = (d.Mb + 7) / 8;
  // few sequences: a K = 4096 / 5120 activation block fits shared memory whole (8 rows x 5120 halves = 82 KB), and the chunk loop
  // with its two barriers per chunk goes away (large-v2 MLP2 at 8 sequences: 13.4 us in three chunks)
  int kc = kSkKC;
  if (d.in_mode == SKINNY_IN_F16) {
    const int budget = (170 * 1024) / (MT * 8) - 64;            // bytes per activation row
    kc = (budget / 2) / 256 * 256;
    kc = kc < kSkKC ? kSkKC : (kc > 8192 ? 8192 : kc);
  }
  SkinnyArgs a{d, S, kc};
  const int KC = d.K < kc ? d.K : kc;
  const size_t smem = (size_t)MT * 8 * (KC * 2 + 64) + (size_t)8 * 16 * (MT * 8 + 1) * 4 + 128 * 4 + (d.in_mode == SKINNY_IN_LN ? (size_t)2 * d.K * 4 : 0);
  const int grid = (strips + S - 1) / S;
  cudaError_t le = cudaSuccess;
#define WB_SK_CASE(M)                                                                                             \
  case M: {                                                                                                       \
    static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];                                                                                   \
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

// ---- KV-cache attention, one CTA per (sequence, head) ---------------------------------------------------------------------------
// No row split, hence no partials, no fence/atomic and no merge pass: a CTA streams the [n_rows][64] K and V slabs of its
// head through a shared-memory ring with TMA tensor copies (box 64 x 128 rows, 128-byte swizzle -> conflict-free ldmatrix,
// out-of-range rows zero-filled by the TMA unit) and writes the 64 outputs of its head. With HBM as the bottleneck every CTA
// progresses at the same rate, so the uneven 1-or-2 CTAs per SM placement costs nothing.
// 8 compute warps each own 16 of the 128 rows of a stage: S = K q^T and O += V^T p on mma.sync m16n8k16, q and p split into
// fp16 hi + lo parts (two MMAs each) so that only K and V themselves are fp16-rounded; fp32 online softmax in the log2 domain.
constexpr int kHaStageRows = 128;
constexpr int kHaThreads = 288;
constexpr int kHaMaxCluster = 8;                   // portable thread-block cluster size: the self block runs one CTA per head
constexpr int kHaTileBytes = kHaStageRows * 128;   // one K (or V) stage tile: 128 rows x 64 halves

struct HeadAttnArgs {
  const float* q;          // [Mb][d] query (self attention), or null when the query projection is fused:
  const float* x;          //   residual stream [Mb][d]; q = LayerNorm(x; ln_g, ln_b) Wq^T + bq computed by the CTA for its head
  const float* ln_g;
  const float* ln_b;
  const __half* wq;        //   [d][d]
  const float* bq;         //   [d]
  __half* out16;
  const DecodeState* state;
  int d, n_rows_fixed, kv_share, n_stages;
  int l2_prefetch_tiles;   // cross attention: stage tiles beyond the ring requested into L2 while q is still being computed
  int pdl_late;            // release the dependent kernel after the main loop instead of at entry
  int Mb;                  // sequences (attn_decode_multi_kernel: the last slab may hold fewer than kv_share)
};

// 8 weight rows x d of a [.][d] fp16 matrix against one fp32 vector in shared memory: lane l owns the 16-byte chunks
// l, l+32, ... of every row. Used by the fused query projection (prologue) of
// attn_decode_head_kernel; the rows are requested in two halves so that the first can be in flight across the wait.
template <int NJW>
__device__ __forceinline__ void head_rows_load(uint4 (&w)[4][NJW], const __half* wbase, int d, int lane, uint64_t wpol) {
  const int n_chunks = d >> 3;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < NJW; ++j) {
      const int c = lane + 32 * j;
      w[r][j] = c < n_chunks ? ptx::ldg_nc_16(wbase + (size_t)r * d + c * 8, wpol) : make_uint4(0, 0, 0, 0);
    }
}
template <int NJW>
__device__ __forceinline__ float head_rows_dot(const uint4 (&w0)[4][NJW], const uint4 (&w1)[4][NJW], const float* s_vec, int lane) {
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
#pragma unroll
  for (int j = 0; j < NJW; ++j) {
    const int c = lane + 32 * j;
    const float4 x0 = *reinterpret_cast<const float4*>(s_vec + c * 8), x1 = *reinterpret_cast<const float4*>(s_vec + c * 8 + 4);
    const float xs8[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint4 wv = r < 4 ? w0[r][j] : w1[r - 4][j];
      const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 wf = __half22float2(wh[e]);
        acc[r] = fmaf(wf.x, xs8[2 * e], acc[r]);
        acc[r] = fmaf(wf.y, xs8[2 * e + 1], acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = warp_sum(acc[r]);
  float mine = 0.f;   // lane r (< 8) keeps the dot product of row r
#pragma unroll
  for (int r = 0; r < 8; ++r) mine = lane == r ? acc[r] : mine;
  return mine;
}

template <int NJW>   // ceil(d / 256): 16-byte weight chunks per lane and row in the fused query projection
__global__ void __launch_bounds__(kHaThreads, NJW <= 3 ? 2 : 1) attn_decode_head_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV, HeadAttnArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  __shared__ __align__(16) float s_x[NJW * 256];   // normalised residual row (fused query projection)
  __shared__ float s_q[64], s_red[16];
  __shared__ float s_xred[8 * 66];                 // row split: (outputs[64], max, sum) per cluster rank, written by the peers
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int h = blockIdx.x, b = blockIdx.y;
  // Few sequences (a single window of a long-form stream, configs[1]): the rows of a (sequence, head) are split over the
  // gridDim.z CTAs of a cluster, whose partial (max, sum, outputs) meet in the rank-0 CTA over distributed shared memory
  const int n_split = (int)gridDim.z, z = (int)blockIdx.z;
  if (n_split > 1) ptx::cluster_arrive_release();   // paired with the wait in front of the remote stores below
  const bool fixed = a.n_rows_fixed > 0;
  TraceScope trace(a.state, fixed ? 201 : 200);
  const int n_stages = a.n_stages;
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 8);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (!a.pdl_late) ptx::grid_dep_launch();
  if (!fixed) ptx::grid_dep_sync();   // self attention: the row count and the newest K/V row come from the previous kernel
  const int n_rows = fixed ? a.n_rows_fixed : ld_state(&a.state->cur_len) + 1;
  const int n_tiles_all = (n_rows + kHaStageRows - 1) / kHaStageRows;
  const int tiles_per = (n_tiles_all + n_split - 1) / n_split;
  const int t_first = z * tiles_per;                                                   // this CTA's tiles: [t_first, t_first + n_tiles)
  const int n_tiles = n_tiles_all - t_first < tiles_per ? (n_tiles_all - t_first > 0 ? n_tiles_all - t_first : 0) : tiles_per;

  float o[4][4], m_run = -INFINITY, l_run = 0.f;

  if (warp == 8) {
    if (lane == 0) {
      const int slab = b / a.kv_share;
      const uint64_t kvpol = fixed ? stream_policy() : ptx::l2_policy(0);   // cross K/V: read once per step
      if (fixed) {   // the producer of a cross-attention CTA runs ahead of q: pull the tiles after the ring into L2 meanwhile
        const int pf_end = n_stages + a.l2_prefetch_tiles < n_tiles ? n_stages + a.l2_prefetch_tiles : n_tiles;
        for (int t = n_stages; t < pf_end; ++t) {
          ptx::tma_prefetch_l2_3d(&tmK, h * 64, (t_first + t) * kHaStageRows, slab);
          ptx::tma_prefetch_l2_3d(&tmV, h * 64, (t_first + t) * kHaStageRows, slab);
        }
      }
      for (int t = 0; t < n_tiles; ++t) {
        // pdl_late = k > 1: the dependent kernel is released when the producer reaches the k-th tile from the end, so that
        // its launch latency runs under the last tiles of the stream (any thread of the CTA can issue the release)
        if (a.pdl_late > 1 && t == (n_tiles > a.pdl_late ? n_tiles - a.pdl_late : 0)) ptx::grid_dep_launch();
        const int s = t % n_stages;
        const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * kHaTileBytes);
        unsigned char* dk = smem + (size_t)s * 2 * kHaTileBytes;
        ptx::tma_load_3d(dk, &tmK, &full_bar[s], h * 64, (t_first + t) * kHaStageRows, slab, kvpol);
        ptx::tma_load_3d(dk + kHaTileBytes, &tmV, &full_bar[s], h * 64, (t_first + t) * kHaStageRows, slab, kvpol);
      }
    }
  } else {
    const float sl = 0.125f * kLog2e;   // (d_head^-0.25)^2 = 1/8 exactly
    const float* qsrc;                  // 64 query values of this head
    if (a.wq) {
      // ---- fused LayerNorm + query projection for this (sequence, head): q_h = LN(x[b]) Wq[h*64..+64]^T + bq -----------------
      // Saves a whole kernel of the latency chain per layer. The weight rows of this warp (8 of the 64) and the LayerNorm
      // affine do not depend on the previous kernel and are requested before griddepcontrol.wait; meanwhile the producer
      // warp is already streaming K/V.
      const int d = a.d;
      const int ct = tid;                                   // 256 compute threads: columns ct, ct+256, ...
      const __half* wbase = a.wq + (size_t)(h * 64 + warp * 8) * d;
      uint4 w0[4][NJW];
      const uint64_t wpol = weight_policy();
      head_rows_load<NJW>(w0, wbase, d, lane, wpol);
      float gv[NJW], bv[NJW];
#pragma unroll
      for (int j = 0; j < NJW; ++j) {
        const int c = ct + 256 * j;
        gv[j] = c < d ? __ldg(a.ln_g + c) : 0.f;
        bv[j] = c < d ? __ldg(a.ln_b + c) : 0.f;
      }
      const float bias = lane < 8 ? __ldg(a.bq + h * 64 + warp * 8 + lane) : 0.f;
      ptx::grid_dep_sync();                                 // x comes from the previous kernel
      uint4 w1[4][NJW];
      head_rows_load<NJW>(w1, wbase + (size_t)4 * d, d, lane, wpol);
      float xv[NJW], sum = 0.f, sq = 0.f;
#pragma unroll
      for (int j = 0; j < NJW; ++j) {
        const int c = ct + 256 * j;
        xv[j] = c < d ? __ldcg(a.x + (size_t)b * d + c) : 0.f;
        sum += xv[j], sq += xv[j] * xv[j];
      }
      sum = warp_sum(sum), sq = warp_sum(sq);
      if (lane == 0) s_red[warp] = sum, s_red[8 + warp] = sq;
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 compute warps only
      float tsum = 0.f, tsq = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) tsum += s_red[w8], tsq += s_red[8 + w8];
      const float mean = tsum / (float)d;
      const float rstd = rsqrtf(fmaxf(tsq / (float)d - mean * mean, 0.f) + 1e-5f);
#pragma unroll
      for (int j = 0; j < NJW; ++j) {
        const int c = ct + 256 * j;
        if (c < NJW * 256) s_x[c] = c < d ? (xv[j] - mean) * rstd * gv[j] + bv[j] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mine = head_rows_dot<NJW>(w0, w1, s_x, lane);
      if (lane < 8) s_q[warp * 8 + lane] = mine + bias;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      qsrc = s_q;
    } else {
      if (fixed) ptx::grid_dep_sync();                      // q comes from the previous kernel
      qsrc = nullptr;
    }
    // q as B fragments (replicated over the 8 n columns), hi + lo fp16 parts, pre-scaled into the log2 domain
    uint32_t qh[4][2], ql[4][2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float2 q0, q1;
      if (qsrc) {
        q0 = *reinterpret_cast<const float2*>(qsrc + kk * 16 + 2 * tq);
        q1 = *reinterpret_cast<const float2*>(qsrc + kk * 16 + 2 * tq + 8);
      } else {
        const float* qp = a.q + (size_t)b * a.d + h * 64 + kk * 16 + 2 * tq;
        q0 = __ldcg(reinterpret_cast<const float2*>(qp)), q1 = __ldcg(reinterpret_cast<const float2*>(qp + 8));
      }
      const float v[4] = {q0.x * sl, q0.y * sl, q1.x * sl, q1.y * sl};
      __half hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __float2half_rn(v[e]);
        lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
      }
      __half2 h0 = __halves2half2(hi[0], hi[1]), h1 = __halves2half2(hi[2], hi[3]);
      __half2 l0 = __halves2half2(lo[0], lo[1]), l1 = __halves2half2(lo[2], lo[3]);
      qh[kk][0] = *reinterpret_cast<uint32_t*>(&h0), qh[kk][1] = *reinterpret_cast<uint32_t*>(&h1);
      ql[kk][0] = *reinterpret_cast<uint32_t*>(&l0), ql[kk][1] = *reinterpret_cast<uint32_t*>(&l1);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % n_stages;
      const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
      ptx::mbar_wait(&full_bar[s], ph);
      const int row0 = (t_first + t) * kHaStageRows + warp * 16;   // first of this warp's 16 rows
      if (row0 < n_rows) {                                 // warp-uniform
        const uint32_t sk = ptx::smem_u32(smem + (size_t)s * 2 * kHaTileBytes);
        const uint32_t sv = sk + kHaTileBytes;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        {
          const int R = warp * 16 + (mi & 1) * 8 + r8;     // tile row this lane addresses for ldmatrix
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t af[4];
            ptx::ldmatrix_x4(af, sk + R * 128 + (((kk * 2 + (mi >> 1)) ^ (R & 7)) << 4));
            ptx::mma_16816(c, af, qh[kk]);
            ptx::mma_16816(c, af, ql[kk]);
          }
        }
        const float s_lo = (row0 + grp < n_rows) ? c[0] : -INFINITY;
        const float s_hi = (row0 + grp + 8 < n_rows) ? c[2] : -INFINITY;
        float tmax = fmaxf(s_lo, s_hi);
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 4));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 8));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 16));
        if (tmax > m_run) {                                // warp-uniform (identical in every lane)
          const float corr = exp2f(m_run - tmax);
          m_run = tmax;
          l_run *= corr;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) o[mt][0] *= corr, o[mt][1] *= corr, o[mt][2] *= corr, o[mt][3] *= corr;
        }
        const float p_lo = exp2f(s_lo - m_run), p_hi = exp2f(s_hi - m_run);
        l_run += p_lo + p_hi;
        const float e0 = __shfl_sync(0xffffffffu, p_lo, (2 * tq) * 4), e1 = __shfl_sync(0xffffffffu, p_lo, (2 * tq + 1) * 4);
        const float e2 = __shfl_sync(0xffffffffu, p_hi, (2 * tq) * 4), e3 = __shfl_sync(0xffffffffu, p_hi, (2 * tq + 1) * 4);
        const __half2 ph0 = __floats2half2_rn(e0, e1), ph1 = __floats2half2_rn(e2, e3);
        const float2 f0 = __half22float2(ph0), f1 = __half22float2(ph1);
        const __half2 pl0 = __floats2half2_rn(e0 - f0.x, e1 - f0.y), pl1 = __floats2half2_rn(e2 - f1.x, e3 - f1.y);
        const uint32_t pbh[2] = {*reinterpret_cast<const uint32_t*>(&ph0), *reinterpret_cast<const uint32_t*>(&ph1)};
        const uint32_t pbl[2] = {*reinterpret_cast<const uint32_t*>(&pl0), *reinterpret_cast<const uint32_t*>(&pl1)};
        {
          const int R = warp * 16 + (mi >> 1) * 8 + r8;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            uint32_t af[4];
            ptx::ldmatrix_x4_trans(af, sv + R * 128 + (((mt * 2 + (mi & 1)) ^ (R & 7)) << 4));
            ptx::mma_16816(o[mt], af, pbh);
            ptx::mma_16816(o[mt], af, pbl);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();   // all stages consumed; reuse the ring for the cross-warp merge: [8][68] floats
  if (a.pdl_late == 1) ptx::grid_dep_launch();
  float* red = reinterpret_cast<float*>(smem);
  if (warp < 8) {
    float L = l_run;
    L += __shfl_xor_sync(0xffffffffu, L, 4);
    L += __shfl_xor_sync(0xffffffffu, L, 8);
    L += __shfl_xor_sync(0xffffffffu, L, 16);
    if (tq == 0) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        red[warp * 68 + mt * 16 + grp] = o[mt][0];
        red[warp * 68 + mt * 16 + grp + 8] = o[mt][2];
      }
      if (grp == 0) red[warp * 68 + 64] = m_run, red[warp * 68 + 65] = L;
    }
  }
  __syncthreads();
  float M = -INFINITY, L = 0.f, A = 0.f;
  if (tid < 64) {
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, red[w * 68 + 64]);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mw = red[w * 68 + 64];
      const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      L += wgt * red[w * 68 + 65];
      A += wgt * red[w * 68 + tid];
    }
  }
  if (n_split == 1) {
    if (tid < 64) a.out16[(size_t)b * a.d + h * 64 + tid] = __float2half_rn(A / L);
    trace.end();
    return;
  }
  // cluster merge: every CTA puts (A[64], M, L) of its row range into slot z of the rank-0 CTA's table (its own shared
  // array: the ring of rank 0 may still be receiving tiles); rank 0 folds the slots in rank order. A CTA without rows
  // contributes M = -inf, L = 0.
  float* xred = s_xred;                                         // [n_split][66]
  ptx::cluster_wait_acquire();                                  // every CTA of the cluster runs (arrival at entry)
  if (tid < 64) {
    const uint32_t dst = ptx::mapa(ptx::smem_u32(xred + z * 66), 0u);
    ptx::st_cluster_f32(dst + 4u * (uint32_t)tid, A);
    if (tid == 0) {
      ptx::st_cluster_f32(dst + 4u * 64u, M);
      ptx::st_cluster_f32(dst + 4u * 65u, L);
    }
  }
  ptx::cluster_arrive_release();
  ptx::cluster_wait_acquire();
  if (z == 0 && tid < 64) {
    float Mx = -INFINITY;
    for (int c = 0; c < n_split; ++c) Mx = fmaxf(Mx, xred[c * 66 + 64]);
    float Lx = 0.f, Ax = 0.f;
    for (int c = 0; c < n_split; ++c) {
      const float mc = xred[c * 66 + 64];
      const float wgt = (mc == -INFINITY) ? 0.f : exp2f(mc - Mx);
      Lx += wgt * xred[c * 66 + 65];
      Ax += wgt * xred[c * 66 + tid];
    }
    a.out16[(size_t)b * a.d + h * 64 + tid] = __float2half_rn(Ax / Lx);
  }
  trace.end();
}

// ---- KV-cache attention for sequences that share one K/V slab (the beams of a chunk, the best_of draws of a window) -------------------
// One CTA per (slab, head) instead of one per (sequence, head): the up to 8 queries are the 8 columns of every mma.sync that the
// single-query kernel fills with copies of one query, so the slab is streamed once (5 beams: a fifth of the L2 -> SM traffic and
// of the CTAs: configs[3]'s 480 CTAs were two waves on 296 slots). Per column online softmax: a lane owns columns 2 tq, 2 tq + 1 of
// rows grp, grp + 8; P goes back in as the B operand of O^T += V^T P through one shuffle pair per element.
template <int NJW>
__global__ void __launch_bounds__(kHaThreads, NJW <= 3 ? 2 : 1) attn_decode_multi_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV, HeadAttnArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  __shared__ __align__(16) float s_q[8 * 64];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int h = blockIdx.x, slab = blockIdx.y;
  // behind the ring: LayerNorm gamma / beta and the normalised rows of the slab's sequences (fused projection only)
  float* s_g = reinterpret_cast<float*>(smem + (size_t)a.n_stages * 2 * kHaTileBytes);
  float* s_b = s_g + NJW * 256;
  float* s_xm = s_b + NJW * 256;                             // [8][NJW * 256]
  const int b0 = slab * a.kv_share;
  const int nq = a.Mb - b0 < a.kv_share ? a.Mb - b0 : a.kv_share;
  TraceScope trace(a.state, 201);
  const int n_stages = a.n_stages;
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 8);
    }
    ptx::fence_mbar_init();
  }
  for (int i = tid; i < 8 * 64; i += kHaThreads) s_q[i] = 0.f;   // unused columns: zero queries
  __syncthreads();
  if (!a.pdl_late) ptx::grid_dep_launch();
  const int n_rows = a.n_rows_fixed;
  const int n_tiles = (n_rows + kHaStageRows - 1) / kHaStageRows;

  float o[4][4], m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;

  if (warp == 8) {
    if (lane == 0) {
      const uint64_t kvpol = stream_policy();
      for (int t = 0; t < n_tiles; ++t) {
        if (a.pdl_late > 1 && t == (n_tiles > a.pdl_late ? n_tiles - a.pdl_late : 0)) ptx::grid_dep_launch();
        const int s = t % n_stages;
        const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * kHaTileBytes);
        unsigned char* dk = smem + (size_t)s * 2 * kHaTileBytes;
        ptx::tma_load_3d(dk, &tmK, &full_bar[s], h * 64, t * kHaStageRows, slab, kvpol);
        ptx::tma_load_3d(dk + kHaTileBytes, &tmV, &full_bar[s], h * 64, t * kHaStageRows, slab, kvpol);
      }
    }
  } else {
    const float sl = 0.125f * kLog2e;   // (d_head^-0.25)^2 = 1/8 exactly
    if (a.wq) {
      // fused LayerNorm + query projection: warp s normalises the row of sequence s into shared memory (fp32), then every warp
      // multiplies its 8 weight rows - held in registers - with all nq rows
      const int d = a.d;
      const __half* wbase = a.wq + (size_t)(h * 64 + warp * 8) * d;
      uint4 w0[4][NJW];
      const uint64_t wpol = weight_policy();
      head_rows_load<NJW>(w0, wbase, d, lane, wpol);
      for (int i = tid * 4; i < d; i += 256 * 4) {
        *reinterpret_cast<float4*>(s_g + i) = __ldg(reinterpret_cast<const float4*>(a.ln_g + i));
        *reinterpret_cast<float4*>(s_b + i) = __ldg(reinterpret_cast<const float4*>(a.ln_b + i));
      }
      const float bias = lane < 8 ? __ldg(a.bq + h * 64 + warp * 8 + lane) : 0.f;
      ptx::grid_dep_sync();
      asm volatile("bar.sync 1, 256;" ::: "memory");        // gamma / beta staged
      if (warp < nq) {
        constexpr int kV4 = NJW * 2;                        // float4 per lane: NJW * 256 columns / (32 lanes * 4)
        float4 xv[kV4];
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < kV4; ++i) {
          const int c = (lane + 32 * i) * 4;
          xv[i] = c < d ? __ldcg(reinterpret_cast<const float4*>(a.x + (size_t)(b0 + warp) * d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
          sq += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
        }
        sum = warp_sum(sum), sq = warp_sum(sq);
        const float mean = sum / (float)d;
        const float rstd = rsqrtf(fmaxf(sq / (float)d - mean * mean, 0.f) + 1e-5f);
        float* dst = s_xm + (size_t)warp * (NJW * 256);
#pragma unroll
        for (int i = 0; i < kV4; ++i) {
          const int c = (lane + 32 * i) * 4;
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < d) {
            const float4 g = *reinterpret_cast<const float4*>(s_g + c), bb = *reinterpret_cast<const float4*>(s_b + c);
            r = make_float4((xv[i].x - mean) * rstd * g.x + bb.x, (xv[i].y - mean) * rstd * g.y + bb.y,
                            (xv[i].z - mean) * rstd * g.z + bb.z, (xv[i].w - mean) * rstd * g.w + bb.w);
          }
          *reinterpret_cast<float4*>(dst + c) = r;
        }
      }
      uint4 w1[4][NJW];                                     // second half of the rows: requested under the barrier
      head_rows_load<NJW>(w1, wbase + (size_t)4 * d, d, lane, wpol);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int s = 0; s < nq; ++s) {
        const float mine = head_rows_dot<NJW>(w0, w1, s_xm + (size_t)s * (NJW * 256), lane);
        if (lane < 8) s_q[s * 64 + warp * 8 + lane] = mine + bias;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    } else {
      ptx::grid_dep_sync();                                 // q comes from the previous kernel
      for (int i = tid; i < nq * 64; i += 256) s_q[i] = __ldcg(a.q + (size_t)(b0 + (i >> 6)) * a.d + h * 64 + (i & 63));
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // q as B fragments: column n = grp is query grp; hi + lo fp16 parts, pre-scaled into the log2 domain
    uint32_t qh[4][2], ql[4][2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float2 q0 = *reinterpret_cast<const float2*>(s_q + grp * 64 + kk * 16 + 2 * tq);
      const float2 q1 = *reinterpret_cast<const float2*>(s_q + grp * 64 + kk * 16 + 2 * tq + 8);
      const float v[4] = {q0.x * sl, q0.y * sl, q1.x * sl, q1.y * sl};
      __half hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __float2half_rn(v[e]);
        lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
      }
      __half2 h0 = __halves2half2(hi[0], hi[1]), h1 = __halves2half2(hi[2], hi[3]);
      __half2 lw0 = __halves2half2(lo[0], lo[1]), lw1 = __halves2half2(lo[2], lo[3]);
      qh[kk][0] = *reinterpret_cast<uint32_t*>(&h0), qh[kk][1] = *reinterpret_cast<uint32_t*>(&h1);
      ql[kk][0] = *reinterpret_cast<uint32_t*>(&lw0), ql[kk][1] = *reinterpret_cast<uint32_t*>(&lw1);
    }
    const int src0 = (2 * tq) * 4 + (grp >> 1), src1 = src0 + 4;   // lanes that hold rows 2 tq / 2 tq + 1 (and + 8) of column grp
    const bool odd = grp & 1;
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % n_stages;
      const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
      ptx::mbar_wait(&full_bar[s], ph);
      const int row0 = t * kHaStageRows + warp * 16;
      if (row0 < n_rows) {                                 // warp-uniform
        const uint32_t sk = ptx::smem_u32(smem + (size_t)s * 2 * kHaTileBytes);
        const uint32_t sv = sk + kHaTileBytes;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        {
          const int R = warp * 16 + (mi & 1) * 8 + r8;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t af[4];
            ptx::ldmatrix_x4(af, sk + R * 128 + (((kk * 2 + (mi >> 1)) ^ (R & 7)) << 4));
            ptx::mma_16816(c, af, qh[kk]);
            ptx::mma_16816(c, af, ql[kk]);
          }
        }
        const bool ok_lo = row0 + grp < n_rows, ok_hi = row0 + grp + 8 < n_rows;
        const float s00 = ok_lo ? c[0] : -INFINITY, s01 = ok_lo ? c[1] : -INFINITY;
        const float s10 = ok_hi ? c[2] : -INFINITY, s11 = ok_hi ? c[3] : -INFINITY;
        float t0 = fmaxf(s00, s10), t1 = fmaxf(s01, s11);
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {
          t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, off));
          t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, off));
        }
        const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);         // finite: row0 itself exists
        const float corr0 = exp2f(m0 - n0), corr1 = exp2f(m1 - n1);  // m = -inf: 0
        m0 = n0, m1 = n1;
        l0 *= corr0, l1 *= corr1;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) o[mt][0] *= corr0, o[mt][2] *= corr0, o[mt][1] *= corr1, o[mt][3] *= corr1;
        const float p00 = exp2f(s00 - m0), p01 = exp2f(s01 - m1), p10 = exp2f(s10 - m0), p11 = exp2f(s11 - m1);
        l0 += p00 + p10, l1 += p01 + p11;
        // B fragment of P (k = row, n = column grp): rows 2 tq, 2 tq + 1, 2 tq + 8, 2 tq + 9
        const float a00 = __shfl_sync(0xffffffffu, p00, src0), a01 = __shfl_sync(0xffffffffu, p01, src0);
        const float b00 = __shfl_sync(0xffffffffu, p00, src1), b01 = __shfl_sync(0xffffffffu, p01, src1);
        const float a10 = __shfl_sync(0xffffffffu, p10, src0), a11 = __shfl_sync(0xffffffffu, p11, src0);
        const float b10 = __shfl_sync(0xffffffffu, p10, src1), b11 = __shfl_sync(0xffffffffu, p11, src1);
        const float e0 = odd ? a01 : a00, e1 = odd ? b01 : b00, e2 = odd ? a11 : a10, e3 = odd ? b11 : b10;
        const __half2 ph0 = __floats2half2_rn(e0, e1), ph1 = __floats2half2_rn(e2, e3);
        const float2 f0 = __half22float2(ph0), f1 = __half22float2(ph1);
        const __half2 pl0 = __floats2half2_rn(e0 - f0.x, e1 - f0.y), pl1 = __floats2half2_rn(e2 - f1.x, e3 - f1.y);
        const uint32_t pbh[2] = {*reinterpret_cast<const uint32_t*>(&ph0), *reinterpret_cast<const uint32_t*>(&ph1)};
        const uint32_t pbl[2] = {*reinterpret_cast<const uint32_t*>(&pl0), *reinterpret_cast<const uint32_t*>(&pl1)};
        {
          const int R = warp * 16 + (mi >> 1) * 8 + r8;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            uint32_t af[4];
            ptx::ldmatrix_x4_trans(af, sv + R * 128 + (((mt * 2 + (mi & 1)) ^ (R & 7)) << 4));
            ptx::mma_16816(o[mt], af, pbh);
            ptx::mma_16816(o[mt], af, pbl);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();   // all stages consumed; the ring becomes the cross-warp merge table [8 warps][8 columns][68]
  if (a.pdl_late == 1) ptx::grid_dep_launch();
  float* red = reinterpret_cast<float*>(smem);
  if (warp < 8) {
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, off);
      l1 += __shfl_xor_sync(0xffffffffu, l1, off);
    }
    float* r0 = red + (size_t)(warp * 8 + 2 * tq) * 68;
    float* r1 = r0 + 68;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      r0[mt * 16 + grp] = o[mt][0], r0[mt * 16 + grp + 8] = o[mt][2];
      r1[mt * 16 + grp] = o[mt][1], r1[mt * 16 + grp + 8] = o[mt][3];
    }
    if (grp == 0) r0[64] = m0, r0[65] = l0, r1[64] = m1, r1[65] = l1;
  }
  __syncthreads();
  for (int idx = tid; idx < nq * 64; idx += kHaThreads) {
    const int col = idx >> 6, dim = idx & 63;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, red[(w * 8 + col) * 68 + 64]);
    float L = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mw = red[(w * 8 + col) * 68 + 64];
      const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      L += wgt * red[(w * 8 + col) * 68 + 65];
      A += wgt * red[(w * 8 + col) * 68 + dim];
    }
    a.out16[(size_t)(b0 + col) * a.d + h * 64 + dim] = __float2half_rn(A / L);
  }
  trace.end();
}

static int launch_attn_decode_head(const AttnDecodeDesc& p, cudaStream_t st, int64_t* launches) {
  CUtensorMap tmK, tmV;
  const long long nslab = (p.Mb + p.kv_share - 1) / p.kv_share;
  int rc = gemm_get_tmap(p.tmaps, p.k, p.d, p.n_ctx, nslab, p.d, (long long)p.n_ctx * p.d, kHaStageRows, &tmK);
  if (rc) return rc;
  rc = gemm_get_tmap(p.tmaps, p.v, p.d, p.n_ctx, nslab, p.d, (long long)p.n_ctx * p.d, kHaStageRows, &tmV);
  if (rc) return rc;
  HeadAttnArgs a{p.q, p.x, p.ln_g, p.ln_b, p.wq, p.bq, p.out16, p.state, p.d, p.n_rows_fixed, p.kv_share, 3, 0, 0, p.Mb};   // L2 prefetch measured slightly negative in-step: off
  static int pdl_xa = -1;
  if (pdl_xa < 0) {
    const char* e = getenv("WB_PDL_XA");
    pdl_xa = e ? atoi(e) : 1;
  }
  a.pdl_late = (p.n_rows_fixed > 0 && p.pdl_late_ok) ? pdl_xa : 0;
  const int ctas = p.n_head * p.Mb;
  if (p.n_rows_fixed <= 0 || ctas > 296) a.n_stages = 2;          // self attention: few rows; big grids: 3 CTAs per SM
  static int stages_env = -1;
  if (stages_env < 0) {
    const char* e = getenv("WB_HA_STAGES");
    stages_env = e ? atoi(e) : 0;
  }
  if (stages_env > 0) a.n_stages = stages_env;
  static int pf_env = -1;
  if (pf_env < 0) {
    const char* e = getenv("WB_HA_L2PF");
    pf_env = e ? atoi(e) + 1000 : 0;
  }
  if (pf_env >= 1000) a.l2_prefetch_tiles = pf_env - 1000;
  const size_t smem = (size_t)a.n_stages * 2 * kHaTileBytes + 1024;
  // row split for small grids (cross attention only: the row count is known on the host): the largest cluster size <= 8 that
  // divides the tiles evenly-ish and keeps the grid within one CTA per SM
  int n_split = 1;
  static int split_env = -1;
  if (split_env < 0) {
    const char* e = getenv("WB_HA_SPLIT");   // development: 0 = never split, n = force
    split_env = e ? atoi(e) : -2;
  }
  if (p.n_rows_fixed > 0 && split_env != 0) {
    const int tiles = (p.n_rows_fixed + kHaStageRows - 1) / kHaStageRows;
    if (split_env > 0) {
      n_split = split_env < 8 ? split_env : 8;
    } else {
      for (int c = 8; c >= 2; --c)
        if (ctas * c <= 148 && tiles >= 2 * c - 1 && (tiles + c - 1) / c * (c - 1) < tiles) {   // every CTA of the cluster gets rows
          n_split = c;
          break;
        }
    }
    if (n_split > tiles) n_split = tiles;
    if (n_split < 1) n_split = 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.n_head, p.Mb, n_split), cfg.blockDim = dim3(kHaThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  if (use_pdl()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  if (n_split > 1) {
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = 1, at[n_at].val.clusterDim.y = 1, at[n_at].val.clusterDim.z = (unsigned)n_split;
    ++n_at;
  }
  cfg.attrs = at, cfg.numAttrs = n_at;
  cudaError_t le = cudaSuccess;
  // sequences that share a slab (beam search, best_of): one CTA per (slab, head) with the sequences as MMA columns
  static int multi_env = -1;
  if (multi_env < 0) {
    const char* e = getenv("WB_HA_MULTI");   // development: 0 = one CTA per (sequence, head) also for shared slabs, 2 = shared slabs always
    multi_env = e ? atoi(e) : 1;
  }
  // ... when that still leaves enough CTAs to pull the stream: a single window with best_of draws (6-20 heads x 1 slab) stays on the
  // one-CTA-per-(sequence, head) kernel, whose grid the row split above widens
  if (multi_env && p.kv_share > 1 && p.kv_share <= 8 && p.n_rows_fixed > 0 && (multi_env >= 2 || p.n_head * (int)nslab >= 64)) {
    // (a row split over a cluster, as in the single-query kernel, measured slower here: 282 -> 394 ms per decode of configs[3])
    cudaLaunchConfig_t cm = cfg;
    cm.gridDim = dim3(p.n_head, (unsigned)nslab, 1);
    cudaLaunchAttribute am[1];
    am[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    am[0].val.programmaticStreamSerializationAllowed = 1;
    cm.attrs = am, cm.numAttrs = use_pdl() ? 1 : 0;
    HeadAttnArgs am_args = a;
    am_args.n_stages = stages_env > 0 ? stages_env : 3;
    const int njw = p.wq ? (p.d + 255) / 256 : 1;
    const size_t smem_m = (size_t)am_args.n_stages * 2 * kHaTileBytes + 1024 + (p.wq ? (size_t)10 * njw * 256 * 4 : 0);
    cm.dynamicSmemBytes = smem_m;
#define WB_HM_CASE(J)                                                                                                      \
  case J: {                                                                                                                \
    static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];                  \
    if (smem_m > smem_set) {                                                                                               \
      WB_CUDA_OK(cudaFuncSetAttribute(attn_decode_multi_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m)); \
      smem_set = smem_m;                                                                                                   \
    }                                                                                                                      \
    le = cudaLaunchKernelEx(&cm, attn_decode_multi_kernel<J>, tmK, tmV, am_args);                                          \
  } break;
    switch (njw) {
      WB_HM_CASE(1) WB_HM_CASE(2) WB_HM_CASE(3) WB_HM_CASE(4) WB_HM_CASE(5)
      default:
        set_error("attn_decode: unsupported width %d", p.d);
        return -1;
    }
#undef WB_HM_CASE
    if (launches) *launches += 1;
    WB_CUDA_OK(le);
    return 0;
  }
#define WB_HA_CASE(J)                                                                                                     \
  case J: {                                                                                                               \
    static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];                                                                                           \
    if (smem > smem_set) {                                                                                                \
      WB_CUDA_OK(cudaFuncSetAttribute(attn_decode_head_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                                    \
    }                                                                                                                     \
    le = cudaLaunchKernelEx(&cfg, attn_decode_head_kernel<J>, tmK, tmV, a);                                               \
  } break;
  // the template parameter sizes the register arrays of the fused query projection only: without it (self attention, or cross
  // attention behind a separate projection) the smallest instantiation runs - 96 registers, two CTAs per SM - whatever the
  // width (the wide instantiations hold 160 registers of weight rows and run one CTA per SM)
  switch (p.wq ? (p.d + 255) / 256 : 1) {
    WB_HA_CASE(1) WB_HA_CASE(2) WB_HA_CASE(3) WB_HA_CASE(4) WB_HA_CASE(5)
    default:
      set_error("attn_decode: unsupported width %d", p.d);
      return -1;
  }
#undef WB_HA_CASE
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

int launch_attn_decode(const AttnDecodeDesc& p, cudaStream_t st, int64_t* launches) {
  if (p.d % 64 != 0 || p.d / 64 != p.n_head || p.d > 1280 || p.kv_share < 1) {
    set_error("attn_decode: unsupported d=%d heads=%d", p.d, p.n_head);
    return -1;
  }
  return launch_attn_decode_head(p, st, launches);
}

// ---- self-attention block in one kernel: LayerNorm + QKV + cache append + attention + output projection + residual -----------------
// For models with n_head <= 8. Grid (head, group of kSbG sequences), launched as one thread-block cluster per group: CTA h
//   1. normalises the group's rows of the residual stream (fp32 LayerNorm -> fp16, as the skinny GEMM's input stage),
//   2. multiplies them with the 192 rows of Wqkv that belong to head h (q, k, v: 64 each; one 16-row strip per warp, weights
//      streamed once with 16-byte loads straight into mma.sync A fragments; the sequences are the 8 MMA columns),
//   3. appends k, v (fp16) to the cache and attends over the cached rows plus the new one (three warps per sequence,
//      16-byte K/V loads, fp32 online softmax in the log2 domain),
//   4. pushes its 64 attention outputs per sequence into the shared memory of every CTA of the cluster (DSMEM), and after
//      the cluster barrier owns columns h*64..h*64+63 of  x += attn Wo^T + bo  (strip x K-third per warp).
// Replaces three kernels of the latency chain (QKV GEMM, attention, output projection) with one. The footprint is kept small
// (64 CTAs at 32 sequences, ~36 KB shared memory, <= 96 registers) so that the CTAs of the following cross-attention kernel
// are resident and streaming K/V while this one runs.
constexpr int kSbThreads = 384;     // 12 warps
constexpr int kSbWarps = kSbThreads / 32;

struct SelfBlockArgs {
  float* x;                 // [Mb][d] residual stream, updated in place
  const float* ln_g;
  const float* ln_b;
  const __half* wqkv;       // [3d][d]
  const float* bqkv;        // [3d]
  const __half* wo;         // [d][d]
  const float* bo;          // [d]
  __half* kcache;           // [Mb][n_ctx][d]
  __half* vcache;
  int Mb, d, n_ctx;
  int pdl_point;            // where the dependent kernel may start: 0 at entry, 1 after the wait, 2 after QKV, 3 after attention
  const DecodeState* state;
};

struct SbPartial {   // online-softmax state of one lane / warp over the 8 columns it owns
  float m, l, o[8];
};
__device__ __forceinline__ void sb_merge_shfl(SbPartial& p, int off) {
  const float m2 = __shfl_xor_sync(0xffffffffu, p.m, off), l2 = __shfl_xor_sync(0xffffffffu, p.l, off);
  const float M = fmaxf(p.m, m2);
  const float w1 = p.m == -INFINITY ? 0.f : exp2f(p.m - M), w2 = m2 == -INFINITY ? 0.f : exp2f(m2 - M);
  p.l = p.l * w1 + l2 * w2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float o2 = __shfl_xor_sync(0xffffffffu, p.o[i], off);
    p.o[i] = p.o[i] * w1 + o2 * w2;
  }
  p.m = M;
}
__device__ __forceinline__ void sb_row(SbPartial& p, float score, const float (&v)[8]) {
  const float mn = fmaxf(p.m, score);
  const float corr = exp2f(p.m - mn), pr = exp2f(score - mn);   // p.m = -inf: corr = 0
  p.l = p.l * corr + pr;
#pragma unroll
  for (int i = 0; i < 8; ++i) p.o[i] = p.o[i] * corr + pr * v[i];
  p.m = mn;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    f[2 * e] = t.x, f[2 * e + 1] = t.y;
  }
}

template <int kSbG>   // sequences per CTA (2 or 4): 12 / kSbG warps share a sequence in the attention phase
__global__ void __maxnreg__(96) self_block_kernel(SelfBlockArgs a) {
  constexpr int WPS = kSbWarps / kSbG;
  extern __shared__ __align__(16) unsigned char sb_smem[];
  TraceScope trace(a.state, 210);
  // Distributed shared memory may only be written once the owning CTA is known to run: every CTA arrives here, and waits for
  // its peers right before the first remote store (phase 3 -> 4, microseconds later: the wait never blocks).
  ptx::cluster_arrive_release();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3;
  const int h = blockIdx.x, b0 = blockIdx.y * kSbG;
  const int d = a.d, nblk = d >> 5;
  const int XS = d * 2 + 64;                                   // bytes per activation row; (XS/16) % 8 == 4 -> conflict-free LDS.128
  unsigned char* xs = sb_smem;                                 // [8][XS] fp16 LayerNorm(x) rows (slots >= kSbG stay zero)
  unsigned char* sa = xs + 8 * XS;                             // [8][XS] fp16 attention outputs of all heads (written by the cluster)
  float* s_g = reinterpret_cast<float*>(sa + 8 * XS);          // [d] LayerNorm gamma
  float* s_b = s_g + d;                                        // [d] beta
  float* s_qkv = s_b + d;                                      // [3][kSbG][64] q, k, v of this head (fp32)
  float* s_part = s_qkv + 3 * kSbG * 64;                       // [kSbG][WPS][66] attention partials of the warps of a sequence
  float* s_red = s_part + kSbG * WPS * 66;                       // [3][kSbG][64] output-projection partials of the three K-thirds

  // ---- before the wait: everything that does not depend on the previous kernel -------------------------------------------------
  // warp w owns strip w of this head's 192 QKV rows: part = w / 4 (q, k, v), rows (w % 4) * 16 .. + 15 of the head
  const int part = warp >> 2, strip = warp & 3;
  const int row_lo = part * d + h * 64 + strip * 16 + grp;     // Wqkv row of accumulator rows grp / grp + 8
  const __half* wrow0 = a.wqkv + (size_t)row_lo * d + tq * 8;
  const __half* wrow1 = wrow0 + (size_t)8 * d;
  constexpr int kPre = 8;
  const uint64_t wpol = weight_policy();
  uint4 pwa[kPre], pwb[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int blk = u < nblk ? u : 0;
    pwa[u] = ptx::ldg_nc_16(wrow0 + blk * 32, wpol);
    pwb[u] = ptx::ldg_nc_16(wrow1 + blk * 32, wpol);
  }
  const float bias_lo = __ldg(a.bqkv + row_lo), bias_hi = __ldg(a.bqkv + row_lo + 8);
  // L2 hints for what is loaded later: the rest of this warp's QKV strip, its slice of Wo, and the cached K/V rows of the
  // group (the row count may be one step stale before the wait: it is only a hint)
#pragma unroll
  for (int u = kPre; u < 16; u += 2) {
    if (u < nblk) {
      weight_prefetch_l2(wrow0 + u * 32);
      weight_prefetch_l2(wrow1 + u * 32);
    }
  }
  {
    const int ostrip_h = warp & 3, kthird_h = warp >> 2;
    const int hb0 = (kthird_h * nblk) / 3, hb1 = ((kthird_h + 1) * nblk) / 3;
    const __half* orow = a.wo + (size_t)(h * 64 + ostrip_h * 16 + grp) * d + tq * 8;
#pragma unroll
    for (int u = 0; u < 6; u += 2) {
      if (hb0 + u < hb1) {
        weight_prefetch_l2(orow + (hb0 + u) * 32);
        weight_prefetch_l2(orow + (size_t)8 * d + (hb0 + u) * 32);
      }
    }
    const int n_hint = ld_state(&a.state->cur_len) + 1;
    const int total = n_hint > 0 ? kSbG * 2 * n_hint : 0;
    for (int i = tid; i < total; i += kSbThreads) {
      const int sq = i / (2 * n_hint), rem = i - sq * 2 * n_hint;
      const int kv = rem / n_hint, rr = rem - kv * n_hint;
      if (b0 + sq < a.Mb && rr < a.n_ctx)
        ptx::prefetch_l2((kv ? a.vcache : a.kcache) + ((size_t)(b0 + sq) * a.n_ctx + rr) * d + h * 64);
    }
  }
  for (int i = tid * 4; i < d; i += kSbThreads * 4) {
    *reinterpret_cast<float4*>(s_g + i) = __ldg(reinterpret_cast<const float4*>(a.ln_g + i));
    *reinterpret_cast<float4*>(s_b + i) = __ldg(reinterpret_cast<const float4*>(a.ln_b + i));
  }
  for (int i = tid; i < (8 - kSbG) * XS / 16; i += kSbThreads) {   // zero the padding slots of both activation tiles
    *reinterpret_cast<uint4*>(xs + kSbG * XS + i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sa + kSbG * XS + i * 16) = make_uint4(0, 0, 0, 0);
  }
  // The dependent kernel is the cross attention, whose CTAs start streaming K/V the moment they are resident: released too
  // early, that stream competes with the latency-bound loads of this kernel and of its predecessor (measured: the step is
  // slower with everything released at entry than with no programmatic launch at all).
  if (a.pdl_point == 0) ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  if (a.pdl_point == 1) ptx::grid_dep_launch();
  __syncthreads();
  trace.mark(3);
  const int pos = ld_state(&a.state->cur_len);                 // cache row of the token this step consumes

  // ---- 1. LayerNorm: warp s normalises row b0 + s (fp32 statistics, two passes over registers) ----------------------------------
  if (warp < kSbG) {
    const int b = b0 + warp;
    __half* xr = reinterpret_cast<__half*>(xs + warp * XS);
    const int n4 = d >> 2;
    float4 v[4];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < n4 && b < a.Mb) ? ld_x4(a.x + (size_t)b * d + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    sum = warp_sum(sum), sq = warp_sum(sq);
    const float mean = sum / (float)d;
    const float rstd = b < a.Mb ? rsqrtf(fmaxf(sq / (float)d - mean * mean, 0.f) + 1e-5f) : 0.f;
    const float ab = b < a.Mb ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < n4) {
        const float4 g = *reinterpret_cast<const float4*>(s_g + c * 4), bb = *reinterpret_cast<const float4*>(s_b + c * 4);
        const __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + ab * bb.x, (v[i].y - mean) * rstd * g.y + ab * bb.y);
        const __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + ab * bb.z, (v[i].w - mean) * rstd * g.w + ab * bb.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(xr + c * 4) = u;
      }
    }
  }
  __syncthreads();
  trace.mark(4);

  // ---- 2. QKV for head h: D[16 weight rows][8 sequence slots] per warp over the whole K ---------------------------------------
  {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned char* xl = xs + grp * XS + tq * 16;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      if (u < nblk) {
        const uint32_t a0[4] = {pwa[u].x, pwb[u].x, pwa[u].y, pwb[u].y}, a1[4] = {pwa[u].z, pwb[u].z, pwa[u].w, pwb[u].w};
        const uint4 xb = *reinterpret_cast<const uint4*>(xl + u * 64);
        const uint32_t bf0[2] = {xb.x, xb.y}, bf1[2] = {xb.z, xb.w};
        ptx::mma_16816(acc, a0, bf0);
        ptx::mma_16816(acc, a1, bf1);
      }
    }
    for (int blk = kPre; blk < nblk; blk += 8) {
      uint4 wa[8], wb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int bb = blk + u < nblk ? blk + u : blk;
        wa[u] = ptx::ldg_nc_16(wrow0 + bb * 32, wpol);
        wb[u] = ptx::ldg_nc_16(wrow1 + bb * 32, wpol);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (blk + u < nblk) {
          const uint32_t a0[4] = {wa[u].x, wb[u].x, wa[u].y, wb[u].y}, a1[4] = {wa[u].z, wb[u].z, wa[u].w, wb[u].w};
          const uint4 xb = *reinterpret_cast<const uint4*>(xl + (blk + u) * 64);
          const uint32_t bf0[2] = {xb.x, xb.y}, bf1[2] = {xb.z, xb.w};
          ptx::mma_16816(acc, a0, bf0);
          ptx::mma_16816(acc, a1, bf1);
        }
      }
    }
    // accumulator (row grp / grp+8, sequence slots 2tq, 2tq+1): + bias -> s_qkv; k and v also go to the cache (fp16)
    if (2 * tq < kSbG) {
      const int c_lo = strip * 16 + grp, c_hi = c_lo + 8;     // column inside the head
      const float v00 = acc[0] + bias_lo, v01 = acc[1] + bias_lo, v10 = acc[2] + bias_hi, v11 = acc[3] + bias_hi;
      float* dst = s_qkv + part * kSbG * 64;
      dst[(2 * tq) * 64 + c_lo] = v00, dst[(2 * tq + 1) * 64 + c_lo] = v01;
      dst[(2 * tq) * 64 + c_hi] = v10, dst[(2 * tq + 1) * 64 + c_hi] = v11;
      if (part > 0) {
        __half* cache = part == 1 ? a.kcache : a.vcache;
        const int bA = b0 + 2 * tq, bB = bA + 1;
        if (bA < a.Mb) {
          __half* r = cache + ((size_t)bA * a.n_ctx + pos) * d + h * 64;
          r[c_lo] = __float2half_rn(v00), r[c_hi] = __float2half_rn(v10);
        }
        if (bB < a.Mb) {
          __half* r = cache + ((size_t)bB * a.n_ctx + pos) * d + h * 64;
          r[c_lo] = __float2half_rn(v01), r[c_hi] = __float2half_rn(v11);
        }
      }
    }
  }
  __syncthreads();
  if (a.pdl_point == 2) ptx::grid_dep_launch();
  trace.mark(5);

  // ---- 3. attention: warps WPS*s .. WPS*s+WPS-1 share sequence slot s; lane = (row sub-index 0..3, 16-byte column chunk 0..7) -----------------
  {
    const int s = warp / WPS, sub = warp - s * WPS;
    const int b = b0 + s;
    const int rsub = lane >> 3, cc = lane & 7;
    const float sl = 0.125f * kLog2e;   // (d_head^-0.25)^2 = 1/8 exactly; log2 domain
    float qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = s_qkv[s * 64 + cc * 8 + i] * sl;
    SbPartial p;
    p.m = -INFINITY, p.l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) p.o[i] = 0.f;
    if (b < a.Mb) {                                            // warp-uniform
      const int per = (pos + WPS - 1) / WPS;                   // cached rows [0, pos) in WPS contiguous ranges
      const int r_begin = sub * per;
      const int r_end = pos < r_begin + per ? pos : r_begin + per;
      const __half* kb = a.kcache + (size_t)b * a.n_ctx * d + h * 64 + cc * 8;
      const __half* vb = a.vcache + (size_t)b * a.n_ctx * d + h * 64 + cc * 8;
      for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        uint4 kq[8], vq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = r0 + j * 4 + rsub;
          const int rc = r < r_end ? r : r_begin;              // clamped: the value is discarded below
          kq[j] = ptx::ldg_nc_16(kb + (size_t)rc * d);
          vq[j] = ptx::ldg_nc_16(vb + (size_t)rc * d);
        }
        // the 8 scores of this lane group first (independent chains), then ONE rescale of the running state for the batch:
        // a per-row online update serialises max -> exp2 -> accumulate eight times over
        float dots[8], bm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float kf[8];
          unpack8(kq[j], kf);
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) dot = fmaf(qv[i], kf[i], dot);
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          dot += __shfl_xor_sync(0xffffffffu, dot, 4);
          dots[j] = (r0 + j * 4 + rsub < r_end) ? dot : -INFINITY;
          bm = fmaxf(bm, dots[j]);
        }
        if (bm > -INFINITY) {                                  // uniform over the 8 lanes that share a row group
          const float mn = fmaxf(p.m, bm);
          const float corr = exp2f(p.m - mn);                  // p.m = -inf: 0
          p.l *= corr;
#pragma unroll
          for (int i = 0; i < 8; ++i) p.o[i] *= corr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pr = exp2f(dots[j] - mn);              // masked rows: exp2(-inf) = 0
            float vf[8];
            unpack8(vq[j], vf);
            p.l += pr;
#pragma unroll
            for (int i = 0; i < 8; ++i) p.o[i] = fmaf(pr, vf[i], p.o[i]);
          }
          p.m = mn;
        }
      }
      if (sub == 0) {   // the new row: k, v as the cache holds them (fp16-rounded), taken from shared memory
        float kf[8], vf[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          kf[i] = __half2float(__float2half_rn(s_qkv[(kSbG + s) * 64 + cc * 8 + i]));
          vf[i] = __half2float(__float2half_rn(s_qkv[(2 * kSbG + s) * 64 + cc * 8 + i]));
        }
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(qv[i], kf[i], dot);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        if (rsub == 0) sb_row(p, dot, vf);
      }
    }
    sb_merge_shfl(p, 8);
    sb_merge_shfl(p, 16);
    if (rsub == 0) {
      float* dst = s_part + (s * WPS + sub) * 66;
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[cc * 8 + i] = p.o[i];
      if (cc == 0) dst[64] = p.m, dst[65] = p.l;
    }
  }
  // this warp's slice of the output projection: strip (warp % 4) of the head's 64 output columns, K-third warp / 4.
  // Requested now so that the latency is covered by the merge and the cluster exchange.
  const int ostrip = warp & 3, kthird = warp >> 2;
  const int ob0 = (kthird * nblk) / 3, ob1 = ((kthird + 1) * nblk) / 3;
  constexpr int kOB = 6;                                       // blocks per K-third: ceil(16 / 3)
  uint4 owa[kOB], owb[kOB];
  {
    const __half* orow0 = a.wo + (size_t)(h * 64 + ostrip * 16 + grp) * d + tq * 8;
    const __half* orow1 = orow0 + (size_t)8 * d;
#pragma unroll
    for (int u = 0; u < kOB; ++u) {
      const int blk = ob0 + u < ob1 ? ob0 + u : ob0;
      owa[u] = ptx::ldg_nc_16(orow0 + blk * 32, wpol);
      owb[u] = ptx::ldg_nc_16(orow1 + blk * 32, wpol);
    }
  }
  // epilogue mapping: thread -> (sequence slot, output column); old residual and bias requested now as well
  const int es = tid >> 6, ec = tid & 63;
  float x_old = 0.f, bias_o = 0.f;
  if (tid < kSbG * 64) {
    bias_o = __ldg(a.bo + h * 64 + ec);
    if (b0 + es < a.Mb) x_old = __ldcg(a.x + (size_t)(b0 + es) * d + h * 64 + ec);
  }
  __syncthreads();
  if (a.pdl_point == 3) ptx::grid_dep_launch();
  trace.mark(6);
  ptx::cluster_wait_acquire();   // all CTAs of the cluster have started (arrival at kernel entry)
  if (tid < kSbG * 64) {   // merge the partials of (slot es, column ec) and push the result into every CTA of the cluster
    const float* pp = s_part + es * WPS * 66;
    float M = -INFINITY;
#pragma unroll
    for (int j = 0; j < WPS; ++j) M = fmaxf(M, pp[j * 66 + 64]);
    float L = 0.f, A = 0.f;
#pragma unroll
    for (int j = 0; j < WPS; ++j) {
      const float mj = pp[j * 66 + 64];
      const float wj = mj == -INFINITY ? 0.f : exp2f(mj - M);
      L += wj * pp[j * 66 + 65];
      A += wj * pp[j * 66 + ec];
    }
    const __half r = __float2half_rn(L > 0.f ? A / L : 0.f);
    const uint32_t local = ptx::smem_u32(sa + es * XS + (h * 64 + ec) * 2);
    const int n_cta = (int)gridDim.x;
    for (int c = 0; c < n_cta; ++c) ptx::st_cluster_u16(ptx::mapa(local, (uint32_t)c), __half_as_ushort(r));
  }
  ptx::cluster_arrive_release();
  ptx::cluster_wait_acquire();   // every head's outputs have landed in sa; no remote access after this point
  trace.mark(7);

  // ---- 4. output projection for columns h*64 .. h*64+63 ------------------------------------------------------------------------------
  {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned char* al = sa + grp * XS + tq * 16;
#pragma unroll
    for (int u = 0; u < kOB; ++u) {
      if (ob0 + u < ob1) {
        const uint32_t a0[4] = {owa[u].x, owb[u].x, owa[u].y, owb[u].y}, a1[4] = {owa[u].z, owb[u].z, owa[u].w, owb[u].w};
        const uint4 xb = *reinterpret_cast<const uint4*>(al + (ob0 + u) * 64);
        const uint32_t bf0[2] = {xb.x, xb.y}, bf1[2] = {xb.z, xb.w};
        ptx::mma_16816(acc, a0, bf0);
        ptx::mma_16816(acc, a1, bf1);
      }
    }
    if (2 * tq < kSbG) {
      float* dst = s_red + kthird * kSbG * 64;
      const int c_lo = ostrip * 16 + grp, c_hi = c_lo + 8;
      dst[(2 * tq) * 64 + c_lo] = acc[0], dst[(2 * tq + 1) * 64 + c_lo] = acc[1];
      dst[(2 * tq) * 64 + c_hi] = acc[2], dst[(2 * tq + 1) * 64 + c_hi] = acc[3];
    }
  }
  __syncthreads();
  if (tid < kSbG * 64 && b0 + es < a.Mb) {
    const float v = (s_red[tid] + s_red[kSbG * 64 + tid]) + s_red[2 * kSbG * 64 + tid];
    a.x[(size_t)(b0 + es) * d + h * 64 + ec] = x_old + (v + bias_o);
  }
  trace.end();
}

int self_block_supported(int n_head, int d) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("WB_SELF_BLOCK");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env && n_head >= 1 && n_head <= kHaMaxCluster && d == n_head * 64;
}

int launch_self_block(const SelfBlockDesc& p, cudaStream_t st, int64_t* launches) {
  if (!self_block_supported(p.n_head, p.d) || p.Mb < 1) {
    set_error("self_block: unsupported shape d=%d heads=%d Mb=%d", p.d, p.n_head, p.Mb);
    return -1;
  }
  static int pdl_point = -1, G = 0;
  if (pdl_point < 0) {
    const char* e = getenv("WB_PDL_SB");
    pdl_point = e ? atoi(e) : 2;
    e = getenv("WB_SB_G");
    G = (e && atoi(e) == 2) ? 2 : 4;
  }
  SelfBlockArgs a{p.x, p.ln_g, p.ln_b, p.wqkv, p.bqkv, p.wo, p.bo, p.kcache, p.vcache, p.Mb, p.d, p.n_ctx, pdl_point, p.state};
  const int XS = p.d * 2 + 64;
  const int wps = kSbWarps / G;
  const size_t smem = (size_t)16 * XS + (size_t)2 * p.d * 4 + (size_t)(3 * G * 64 + G * wps * 66 + 3 * G * 64) * 4;
  static size_t smem_set_dev[kMaxDevices] = {}; size_t& smem_set = smem_set_dev[current_device_slot()];
  if (smem > smem_set) {
    WB_CUDA_OK(cudaFuncSetAttribute(self_block_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WB_CUDA_OK(cudaFuncSetAttribute(self_block_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.n_head, (p.Mb + G - 1) / G), cfg.blockDim = dim3(kSbThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  at[n_at].id = cudaLaunchAttributeClusterDimension;
  at[n_at].val.clusterDim.x = (unsigned)p.n_head, at[n_at].val.clusterDim.y = 1, at[n_at].val.clusterDim.z = 1;
  ++n_at;
  if (use_pdl()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  cfg.attrs = at, cfg.numAttrs = n_at;
  const cudaError_t le = G == 2 ? cudaLaunchKernelEx(&cfg, self_block_kernel<2>, a) : cudaLaunchKernelEx(&cfg, self_block_kernel<4>, a);
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// ---- everything after cross attention in one kernel: output projection + residual + LayerNorm + MLP (GELU) + residual ---------------
// For d = 384 / 512. Grid (C, groups of 8 sequences), one thread-block cluster of C CTAs per group (C = 16 for d = 512: non-portable
// size, C = 8 otherwise). The 8 sequences of a group are the 8 columns of every mma.sync; weights stream once per group with
// 16-byte loads straight into A fragments. CTA r of a cluster
//   0. multiplies the cross-attention outputs a16 with rows r*OC .. of Wo (OC = d/C output columns), adds bias and the old
//      residual, and pushes this slice of x' into the shared memory of all C CTAs (DSMEM all-gather),
//   1. normalises the 8 rows of x' (every CTA needs the whole rows), fp32 statistics -> fp16,
//   2. computes its HS = 4d/C hidden units h = gelu(W1[r*HS ..] LN(x') + b1) -> fp16 in shared memory,
//   3. multiplies them with the matching K-slice of W2: a partial [8][d] of the MLP output,
//   4. after the second cluster barrier sums the C partials of its own OC columns in rank order (remote shared-memory reads,
//      deterministic), adds b2 and x' and writes x.
// Replaces three kernels of the latency chain (cross-attention output projection, MLP1, MLP2).
constexpr int kPbThreads = 256;

struct PostBlockArgs {
  float* x;                 // [Mb][d] residual stream, updated in place
  const __half* a16;        // [Mb][d] cross-attention outputs
  const __half* wo;         // [d][d]
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  const __half* w1;         // [4d][d]
  const float* b1;
  const __half* w2;         // [d][4d]
  const float* b2;
  int Mb;
  int pdl_point;            // where the dependent kernel may start: 0 at entry, 1 after phase 0, 2 after phase 2, 3 after phase 3
  int hints_after_wait;
  const DecodeState* state;
};

constexpr int pb_batch(int n) {   // largest divisor of n that is <= 8: blocks requested at once (16 registers each)
  return n % 8 == 0 ? 8 : (n % 7 == 0 ? 7 : (n % 6 == 0 ? 6 : (n % 5 == 0 ? 5 : (n % 4 == 0 ? 4 : (n % 3 == 0 ? 3 : (n % 2 == 0 ? 2 : 1))))));
}
template <int D, int C>
struct PbCfg {
  static constexpr int OC = D / C, HS = 4 * D / C, NBLK = D / 32;
  static constexpr int S0 = OC / 16, KS0 = 8 / S0, NB0 = NBLK / KS0;                 // phase 0: strips, K split, blocks per unit
  static constexpr int NB0b = pb_batch(NB0);                                          //   ... requested NB0b at a time
  static constexpr int SB = HS / 16, KSB = (SB % 8 == 0) ? 1 : 2, NBU = NBLK / KSB;  // phase 2
  static constexpr int UPW = SB * KSB / 8;                                           //   units per warp
  static constexpr int SCW = D / 16 / 8, NBC = HS / 32;                              // phase 3: strips per warp, blocks per strip
  static constexpr int XS = D * 2 + 64, HSS = HS * 2 + 64, PS = D + 4;               // row strides: bytes, bytes, floats
  static constexpr int NJ = (OC * 8 + 255) / 256;                                    // (slot, column) outputs per thread
  static constexpr int LNV = (D / 4 + 31) / 32;                                      // float4 per lane and row in the LayerNorm
  static constexpr int RED = (KS0 * 8 * OC > KSB * 8 * HS) ? KS0 * 8 * OC : KSB * 8 * HS;
  static constexpr size_t smem_used = (size_t)2 * 8 * XS + (size_t)8 * HSS + ((size_t)8 * D + 8 * PS + RED + 2 * D + HS) * 4;
  // Requested size: more than half of an SM's shared memory, so that two CTAs of this kernel never share an SM. Released
  // while the cross attention still occupies most SMs, the clusters were otherwise packed two CTAs per SM onto the few free
  // ones and every weight-streaming phase took twice as long (per-SM L2 bandwidth).
  static constexpr size_t smem = smem_used > (size_t)118 * 1024 ? smem_used : (size_t)118 * 1024;
  static_assert(OC % 16 == 0 && HS % 32 == 0 && NBLK % KS0 == 0 && NBLK % KSB == 0 && (SB * KSB) % 8 == 0 && (D / 16) % 8 == 0, "shape");
  static_assert(smem <= (size_t)227 * 1024, "shared memory");
};

// one warp: acc += W[16 rows][blocks blk0 .. blk0+NB) x B tile (8 slots); NB <= 8 blocks requested at once
template <int NB>
__device__ __forceinline__ void pb_load(uint4 (&wa)[NB], uint4 (&wb)[NB], const __half* wrow0, const __half* wrow1, int blk0, uint64_t wpol) {
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    wa[u] = ptx::ldg_nc_16(wrow0 + (blk0 + u) * 32, wpol);
    wb[u] = ptx::ldg_nc_16(wrow1 + (blk0 + u) * 32, wpol);
  }
}
template <int NB>
__device__ __forceinline__ void pb_mma(float (&acc)[4], const uint4 (&wa)[NB], const uint4 (&wb)[NB], const unsigned char* bl, int bblk0) {
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const uint32_t a0[4] = {wa[u].x, wb[u].x, wa[u].y, wb[u].y}, a1[4] = {wa[u].z, wb[u].z, wa[u].w, wb[u].w};
    const uint4 xb = *reinterpret_cast<const uint4*>(bl + (bblk0 + u) * 64);
    const uint32_t bf0[2] = {xb.x, xb.y}, bf1[2] = {xb.z, xb.w};
    ptx::mma_16816(acc, a0, bf0);
    ptx::mma_16816(acc, a1, bf1);
  }
}

template <int D, int C>
__global__ void __maxnreg__(D > 1024 ? 224 : 128) post_block_kernel(PostBlockArgs a) {
  using Cfg = PbCfg<D, C>;
  constexpr int OC = Cfg::OC, HS = Cfg::HS, XS = Cfg::XS, HSS = Cfg::HSS, PS = Cfg::PS;
  extern __shared__ __align__(16) unsigned char pb_smem[];
  TraceScope trace(a.state, 220);
  ptx::cluster_arrive_release();   // remote shared memory is written only after every CTA of the cluster is known to run (phase 0)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tq = lane & 3;
  const int r = blockIdx.x, b0 = blockIdx.y * 8;                // cluster rank, first sequence of the group
  unsigned char* as16 = pb_smem;                                // [8][XS] fp16 cross-attention outputs
  unsigned char* xs = as16 + 8 * XS;                            // [8][XS] fp16 LayerNorm(x')
  unsigned char* hs = xs + 8 * XS;                              // [8][HSS] fp16 hidden slice
  float* xp = reinterpret_cast<float*>(hs + 8 * HSS);           // [8][D] x' (all-gathered from the cluster)
  float* s_part = xp + 8 * D;                                   // [8][PS] MLP2 partial (read by the cluster)
  float* s_red = s_part + 8 * PS;                               // K-split partials of phases 0 and 2
  float* s_g = s_red + Cfg::RED;                                // [D]
  float* s_b = s_g + D;                                         // [D]
  float* s_b1 = s_b + D;                                        // [HS]

  // ---- before the wait ---------------------------------------------------------------------------------------------------------------
  const int strip0 = warp % Cfg::S0, kp0 = warp / Cfg::S0;
  const bool act0 = warp < Cfg::S0 * Cfg::KS0;
  const uint64_t wpol = weight_policy();
  uint4 wa0[Cfg::NB0b], wb0[Cfg::NB0b];
  const __half* w0row = a.wo + (size_t)(r * OC + strip0 * 16 + grp) * D + tq * 8;
  pb_load<Cfg::NB0b>(wa0, wb0, w0row, w0row + (size_t)8 * D, act0 ? kp0 * Cfg::NB0 : 0, wpol);
  for (int i = tid * 4; i < D; i += kPbThreads * 4) {
    *reinterpret_cast<float4*>(s_g + i) = __ldg(reinterpret_cast<const float4*>(a.ln_g + i));
    *reinterpret_cast<float4*>(s_b + i) = __ldg(reinterpret_cast<const float4*>(a.ln_b + i));
  }
  for (int i = tid; i < HS; i += kPbThreads) s_b1[i] = __ldg(a.b1 + r * HS + i);
  // L2 hints for this CTA's slices of W1 (HS rows x D) and W2 (D rows x HS), 128-byte lines. When this kernel is released
  // while the cross attention is still streaming, hints given before the wait are evicted again by that stream (measured:
  // phases 2 and 3 take twice as long), so they are given after the wait in that case.
  auto weight_hints = [&]() {
    for (int i = tid; i < HS * (D / 64); i += kPbThreads) {
      const int row = i / (D / 64), seg = i - row * (D / 64);
      weight_prefetch_l2(a.w1 + (size_t)(r * HS + row) * D + seg * 64);
    }
    for (int i = tid; i < D * (HS / 64); i += kPbThreads) {
      const int row = i / (HS / 64), seg = i - row * (HS / 64);
      weight_prefetch_l2(a.w2 + (size_t)row * (4 * D) + r * HS + seg * 64);
    }
  };
  if (!a.hints_after_wait) weight_hints();
  constexpr int NJ = Cfg::NJ;
  float bias_o[NJ], b2v[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = tid + j * kPbThreads;
    const int col = i % OC;
    bias_o[j] = i < OC * 8 ? __ldg(a.bo + r * OC + col) : 0.f;
    b2v[j] = i < OC * 8 ? __ldg(a.b2 + r * OC + col) : 0.f;
  }
  if (a.pdl_point == 0) ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  trace.mark(3);
  if (a.hints_after_wait) weight_hints();

  // ---- 0. x' slice = x + a16 Wo[r*OC ..]^T + bo, pushed to every CTA of the cluster ---------------------------------------------------
  for (int i = tid; i < 8 * (D / 8); i += kPbThreads) {          // a16 rows of the group -> as16
    const int slot = i / (D / 8), c = i - slot * (D / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (b0 + slot < a.Mb) v = __ldcg(reinterpret_cast<const uint4*>(a.a16 + (size_t)(b0 + slot) * D) + c);
    *reinterpret_cast<uint4*>(as16 + slot * XS + c * 16) = v;
  }
  float x_old[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = tid + j * kPbThreads;
    const int slot = i / OC, col = i - slot * OC;
    x_old[j] = (i < OC * 8 && b0 + slot < a.Mb) ? __ldcg(a.x + (size_t)(b0 + slot) * D + r * OC + col) : 0.f;
  }
  __syncthreads();
  if (act0) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    pb_mma<Cfg::NB0b>(acc, wa0, wb0, as16 + grp * XS + tq * 16, kp0 * Cfg::NB0);
#pragma unroll
    for (int bb = Cfg::NB0b; bb < Cfg::NB0; bb += Cfg::NB0b) {   // wider models: the rest of the unit's blocks
      pb_load<Cfg::NB0b>(wa0, wb0, w0row, w0row + (size_t)8 * D, kp0 * Cfg::NB0 + bb, wpol);
      pb_mma<Cfg::NB0b>(acc, wa0, wb0, as16 + grp * XS + tq * 16, kp0 * Cfg::NB0 + bb);
    }
    float* dst = s_red + kp0 * 8 * OC;
    const int c_lo = strip0 * 16 + grp, c_hi = c_lo + 8;
    dst[(2 * tq) * OC + c_lo] = acc[0], dst[(2 * tq + 1) * OC + c_lo] = acc[1];
    dst[(2 * tq) * OC + c_hi] = acc[2], dst[(2 * tq + 1) * OC + c_hi] = acc[3];
  }
  __syncthreads();
  ptx::cluster_wait_acquire();     // pairs with the arrival at kernel entry
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = tid + j * kPbThreads;
    if (i < OC * 8) {
      const int slot = i / OC, col = i - slot * OC;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < Cfg::KS0; ++k) v += s_red[k * 8 * OC + i];
      v = x_old[j] + (v + bias_o[j]);
      const uint32_t local = ptx::smem_u32(xp + slot * D + r * OC + col);
#pragma unroll
      for (int c = 0; c < C; ++c) ptx::st_cluster_f32(ptx::mapa(local, (uint32_t)c), v);
    }
  }
  // weights never depend on activations: the first batch of phase 2 is requested before the barrier and the LayerNorm
  constexpr int NBb = pb_batch(Cfg::NBU);                        // blocks per batch
  uint4 wa2[NBb], wb2[NBb];
  {
    const int strip = warp % Cfg::SB, kp = warp / Cfg::SB;
    const __half* w0 = a.w1 + (size_t)(r * HS + strip * 16 + grp) * D + tq * 8;
    pb_load<NBb>(wa2, wb2, w0, w0 + (size_t)8 * D, kp * Cfg::NBU, wpol);
  }
  ptx::cluster_arrive_release();
  ptx::cluster_wait_acquire();
  if (a.pdl_point == 1) ptx::grid_dep_launch();
  trace.mark(4);

  // ---- 1. LayerNorm of x' (warp w: row w) ---------------------------------------------------------------------------------------------
  {
    const bool live = b0 + warp < a.Mb;
    const float* xr = xp + warp * D;
    __half* dst = reinterpret_cast<__half*>(xs + warp * XS);
    constexpr int N4 = D / 4;
    float4 v[Cfg::LNV];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < Cfg::LNV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < N4 ? *reinterpret_cast<const float4*>(xr + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    sum = warp_sum(sum), sq = warp_sum(sq);
    const float mean = sum / (float)D;
    const float rstd = live ? rsqrtf(fmaxf(sq / (float)D - mean * mean, 0.f) + 1e-5f) : 0.f;
    const float ab = live ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < Cfg::LNV; ++i) {
      const int c = lane + 32 * i;
      if (c < N4) {
        const float4 g = *reinterpret_cast<const float4*>(s_g + c * 4), bb = *reinterpret_cast<const float4*>(s_b + c * 4);
        const __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + ab * bb.x, (v[i].y - mean) * rstd * g.y + ab * bb.y);
        const __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + ab * bb.z, (v[i].w - mean) * rstd * g.w + ab * bb.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(dst + c * 4) = u;
      }
    }
  }
  __syncthreads();
  trace.mark(5);

  // ---- 2. hidden slice: h = gelu(W1[r*HS ..] LN(x') + b1) -------------------------------------------------------------------------------
  constexpr int NPAIR = Cfg::SCW * Cfg::NBC;                     // phase 3: (strip, block) pairs of this warp
  const __half* w2base = a.w2 + (size_t)grp * (4 * D) + r * HS + tq * 8;
  auto load_pairs = [&](uint4 (&wa)[8], uint4 (&wb)[8], int p0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u < NPAIR ? p0 + u : NPAIR - 1;
      const int si = p / Cfg::NBC, blk = p % Cfg::NBC;
      const __half* w0 = w2base + (size_t)((warp + 8 * si) * 16) * (4 * D) + blk * 32;
      wa[u] = ptx::ldg_nc_16(w0, wpol);
      wb[u] = ptx::ldg_nc_16(w0 + (size_t)8 * (4 * D), wpol);
    }
  };
  uint4 wa3[8], wb3[8];
#pragma unroll
  for (int ui = 0; ui < Cfg::UPW; ++ui) {
    const int u = warp + 8 * ui;
    const int strip = u % Cfg::SB, kp = u / Cfg::SB;
    const __half* w0 = a.w1 + (size_t)(r * HS + strip * 16 + grp) * D + tq * 8;
    const __half* w1r = w0 + (size_t)8 * D;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int bb = 0; bb < Cfg::NBU; bb += NBb) {
      if (ui == 0 && bb == 0) {
        pb_mma<NBb>(acc, wa2, wb2, xs + grp * XS + tq * 16, kp * Cfg::NBU);
      } else {
        uint4 wa[NBb], wb[NBb];
        pb_load<NBb>(wa, wb, w0, w1r, kp * Cfg::NBU + bb, wpol);
        pb_mma<NBb>(acc, wa, wb, xs + grp * XS + tq * 16, kp * Cfg::NBU + bb);
      }
    }
    if (ui == Cfg::UPW - 1) load_pairs(wa3, wb3, 0);             // first batch of phase 3, in flight across the GELU pass
    float* dst = s_red + kp * 8 * HS;
    const int c_lo = strip * 16 + grp, c_hi = c_lo + 8;
    dst[(2 * tq) * HS + c_lo] = acc[0], dst[(2 * tq + 1) * HS + c_lo] = acc[1];
    dst[(2 * tq) * HS + c_hi] = acc[2], dst[(2 * tq + 1) * HS + c_hi] = acc[3];
  }
  __syncthreads();
  for (int i = tid; i < 8 * HS; i += kPbThreads) {
    const int slot = i / HS, j = i - slot * HS;
    float v = s_b1[j];
#pragma unroll
    for (int k = 0; k < Cfg::KSB; ++k) v += s_red[k * 8 * HS + i];
    reinterpret_cast<__half*>(hs + slot * HSS)[j] = __float2half_rn(gelu_erf(v));
  }
  __syncthreads();
  if (a.pdl_point == 2) ptx::grid_dep_launch();
  trace.mark(6);

  // ---- 3. partial MLP output over this CTA's hidden slice: strips warp, warp+8, ... of all D output rows --------------------------------
  {
    float acc[Cfg::SCW][4];
#pragma unroll
    for (int si = 0; si < Cfg::SCW; ++si) acc[si][0] = acc[si][1] = acc[si][2] = acc[si][3] = 0.f;
    const unsigned char* hl = hs + grp * HSS + tq * 16;
#pragma unroll
    for (int p0 = 0; p0 < NPAIR; p0 += 8) {
      if (p0 > 0) load_pairs(wa3, wb3, p0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (p0 + u < NPAIR) {
          const int si = (p0 + u) / Cfg::NBC, blk = (p0 + u) % Cfg::NBC;
          const uint32_t a0[4] = {wa3[u].x, wb3[u].x, wa3[u].y, wb3[u].y}, a1[4] = {wa3[u].z, wb3[u].z, wa3[u].w, wb3[u].w};
          const uint4 xb = *reinterpret_cast<const uint4*>(hl + blk * 64);
          const uint32_t bf0[2] = {xb.x, xb.y}, bf1[2] = {xb.z, xb.w};
          ptx::mma_16816(acc[si], a0, bf0);
          ptx::mma_16816(acc[si], a1, bf1);
        }
      }
    }
#pragma unroll
    for (int si = 0; si < Cfg::SCW; ++si) {
      const int c_lo = (warp + 8 * si) * 16 + grp, c_hi = c_lo + 8;
      s_part[(2 * tq) * PS + c_lo] = acc[si][0], s_part[(2 * tq + 1) * PS + c_lo] = acc[si][1];
      s_part[(2 * tq) * PS + c_hi] = acc[si][2], s_part[(2 * tq + 1) * PS + c_hi] = acc[si][3];
    }
  }
  ptx::cluster_arrive_release();
  ptx::cluster_wait_acquire();
  if (a.pdl_point == 3) ptx::grid_dep_launch();
  trace.mark(7);

  // ---- 4. x = x' + b2 + sum over the cluster of the partials (rank order), own OC columns -----------------------------------------------
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = tid + j * kPbThreads;
    if (i < OC * 8) {
      const int slot = i / OC, col = i - slot * OC;
      const uint32_t local = ptx::smem_u32(s_part + slot * PS + r * OC + col);
      float pv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) pv[c] = ptx::ld_cluster_f32(ptx::mapa(local, (uint32_t)c));
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) v += pv[c];
      if (b0 + slot < a.Mb) a.x[(size_t)(b0 + slot) * D + r * OC + col] = xp[slot * D + r * OC + col] + (v + b2v[j]);
    }
  }
  ptx::cluster_arrive_release();   // no CTA may exit (and free its shared memory) while a peer still reads its partial
  ptx::cluster_wait_acquire();
  trace.end();
}

// cluster size the post block uses for width d (0: unsupported). Sizes above 8 are non-portable: asked for once per device.
template <int D, int C>
static bool post_block_cluster_fits() {
  if (cudaFuncSetAttribute(post_block_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PbCfg<D, C>::smem) != cudaSuccess) return false;
  if (C > 8 && cudaFuncSetAttribute(post_block_kernel<D, C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return false;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, 1), cfg.blockDim = dim3(kPbThreads), cfg.dynamicSmemBytes = PbCfg<D, C>::smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  int nc = 0;
  return cudaOccupancyMaxActiveClusters(&nc, post_block_kernel<D, C>, &cfg) == cudaSuccess && nc >= 1;
}
static int post_block_cluster(int d) {
  static int cached[kMaxDevices][5];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < kMaxDevices; ++i)
      for (int j = 0; j < 5; ++j) cached[i][j] = -1;
    init = true;
  }
  const int slot = d == 384 ? 0 : d == 512 ? 1 : d == 768 ? 2 : d == 1024 ? 3 : d == 1280 ? 4 : -1;
  if (slot < 0) return 0;
  int& c = cached[current_device_slot()][slot];
  if (c < 0) {
    // d = 768 / 1024: measured at 40 sequences 1284 -> 1098 and 3629 -> 3502 us per step against the skinny-GEMM chain;
    // d = 1280 (ten CTAs per cluster, 1.3 MB of W1 per CTA): 6175 -> 6435, so off unless WB_POST_BLOCK_WIDE=2. 0 = all off.
    static int wide = -1;
    if (wide < 0) {
      const char* e = getenv("WB_POST_BLOCK_WIDE");
      wide = e ? atoi(e) : 1;
    }
    const char* e = getenv("WB_POST_CLUSTER");
    const int want = e ? atoi(e) : 16;
    switch (d) {
      case 384: c = 8; break;
      case 512: c = (want == 16 && post_block_cluster_fits<512, 16>()) ? 16 : 8; break;
      case 768: {
        // 16 CTAs of 48 output columns / 192 hidden units each (80 CTAs at 40 sequences) against 12 of 64 / 256 (60 CTAs): the
        // phases are bound by what one SM pulls from L2: 1030 -> 1009 us per step, configs[3] 250.8 -> 241.3 ms per decode
        const char* e768 = getenv("WB_POST_CLUSTER_768");
        const int want768 = e768 ? atoi(e768) : 16;
        c = !wide ? 0 : (want768 == 16 && post_block_cluster_fits<768, 16>()) ? 16 : (post_block_cluster_fits<768, 12>() ? 12 : 0);
      } break;
      case 1024: c = (wide && post_block_cluster_fits<1024, 16>()) ? 16 : 0; break;
      default: c = (wide >= 2 && post_block_cluster_fits<1280, 10>()) ? 10 : 0; break;
    }
    cudaGetLastError();
  }
  return c;
}

int post_block_supported(int n_head, int d) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("WB_POST_BLOCK");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env && d == n_head * 64 && post_block_cluster(d) > 0;
}

template <int D, int C>
static cudaError_t launch_post_block_t(const PostBlockArgs& a, int n_groups, cudaStream_t st) {
  static bool attr_set_dev[kMaxDevices] = {}; bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(post_block_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PbCfg<D, C>::smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
      e = cudaFuncSetAttribute(post_block_kernel<D, C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, n_groups), cfg.blockDim = dim3(kPbThreads), cfg.dynamicSmemBytes = PbCfg<D, C>::smem, cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  at[n_at].id = cudaLaunchAttributeClusterDimension;
  at[n_at].val.clusterDim.x = C, at[n_at].val.clusterDim.y = 1, at[n_at].val.clusterDim.z = 1;
  ++n_at;
  if (use_pdl()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  cfg.attrs = at, cfg.numAttrs = n_at;
  return cudaLaunchKernelEx(&cfg, post_block_kernel<D, C>, a);
}

int launch_post_block(const PostBlockDesc& p, cudaStream_t st, int64_t* launches) {
  const int C = post_block_supported(p.n_head, p.d) ? post_block_cluster(p.d) : 0;
  if (!C || p.Mb < 1) {
    set_error("post_block: unsupported shape d=%d heads=%d Mb=%d", p.d, p.n_head, p.Mb);
    return -1;
  }
  static int pdl_point = -1, hints_late = 0;
  if (pdl_point < 0) {
    const char* e = getenv("WB_PDL_PB");
    pdl_point = e ? atoi(e) : 3;
    e = getenv("WB_PB_HINTS_LATE");
    hints_late = e ? atoi(e) : 0;
  }
  PostBlockArgs a{p.x, p.a16, p.wo, p.bo, p.ln_g, p.ln_b, p.w1, p.b1, p.w2, p.b2, p.Mb, pdl_point, hints_late, p.state};
  const int n_groups = (p.Mb + 7) / 8;
  cudaError_t le;
  if (p.d == 384)
    le = launch_post_block_t<384, 8>(a, n_groups, st);
  else if (p.d == 768)
    le = C == 16 ? launch_post_block_t<768, 16>(a, n_groups, st) : launch_post_block_t<768, 12>(a, n_groups, st);
  else if (p.d == 1024)
    le = launch_post_block_t<1024, 16>(a, n_groups, st);
  else if (p.d == 1280)
    le = launch_post_block_t<1280, 10>(a, n_groups, st);
  else if (C == 16)
    le = launch_post_block_t<512, 16>(a, n_groups, st);
  else
    le = launch_post_block_t<512, 8>(a, n_groups, st);
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// ---- end of step: sample (optional), embed the next token, advance -------------------------------------------------------------
// Upstream DecodingTask with temperature 0: the logit filters were applied in the logits GEMM epilogue;
// here argmax over the CTA partials, sum_logprobs += logprob * (previous token != eot), sequences that ended keep
// emitting eot. Then x[b] = token_embedding[next] + positional_embedding[cur_len + 1] for the next step.
// online (max, argmax, sum-exp) accumulator over partial records
struct LseAcc {
  float best, se;
  int arg;
};
__device__ __forceinline__ void lse_fold(LseAcc& a, float m, int idx, float z) {
  if (m > a.best) {
    a.se = a.se * expf(a.best - m) + z;   // best = -inf first: se is 0
    a.best = m, a.arg = idx;
  } else if (m > -INFINITY) {
    a.se += z * expf(m - a.best);
    if (m == a.best && idx < a.arg) a.arg = idx;
  }
}
__device__ __forceinline__ void lse_merge(LseAcc& a, float om, int oa, float os) {
  const float nm = fmaxf(a.best, om);
  a.se = (a.best > -INFINITY ? a.se * expf(a.best - nm) : 0.f) + (om > -INFINITY ? os * expf(om - nm) : 0.f);
  if (om > a.best || (om == a.best && oa < a.arg)) a.arg = oa;
  a.best = nm;
}

__global__ void __launch_bounds__(256) step_finish_kernel(FinishDesc p) {
  __shared__ float s_val[2][8], s_sum[2][8];
  __shared__ int s_idx[2][8];
  __shared__ int s_tok, s_cur;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TraceScope trace(p.state, 300 + p.sample);
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  int32_t* trow = p.tokens + (size_t)b * p.tokens_ld;
  int prev_tok = 0;
  float slp_old = 0.f;
  if (tid == 0) {
    // one reader per CTA, then the arrival ticket: the last CTA to have read cur_len advances it (nothing later in this
    // kernel reads it again; the kernels of the next step see the new value)
    const int c = ld_state(&p.state->cur_len);   // index of the token this step consumed (-1 before the first step)
    s_cur = c;
    const int ticket = atomicAdd(&p.state->arrive, 1);
    if (ticket == (int)gridDim.x - 1) {
      p.state->arrive = 0;
      p.state->cur_len = c + 1;
    }
    if (p.sample) {
      prev_tok = __ldcg(trow + c);
      slp_old = __ldcg(p.sum_logprob + b);
    } else {
      s_tok = __ldcg(trow + c + 1);
    }
  }
  // the logits partials of this sequence: each thread folds its records into online (max, argmax, sum-exp) accumulators,
  // one for the text rows and (timestamp rules) one for the timestamp rows: groups >= ts_group0 and the extra record
  const bool ts_on = p.ts_state != nullptr;
  const int ts_group0 = ts_on ? p.ts_group0 : 0x7fffffff;
  LseAcc acc[2] = {{-INFINITY, 0.f, 0x7fffffff}, {-INFINITY, 0.f, 0x7fffffff}};
  if (p.sample) {
    for (int i0 = tid; i0 < p.n_part; i0 += 256 * 4) {
      float4 rec[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * 256;
        rec[j] = make_float4(-INFINITY, __int_as_float(0x7fffffff), 0.f, 0.f);
        if (i < p.n_part) rec[j] = __ldcg(reinterpret_cast<const float4*>(p.part_logits + ((size_t)b * p.n_part + i) * 4));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * 256;
        if (i >= ts_group0)
          lse_fold(acc[1], rec[j].x, __float_as_int(rec[j].y), rec[j].z);
        else
          lse_fold(acc[0], rec[j].x, __float_as_int(rec[j].y), rec[j].z);
      }
    }
    if (ts_on && tid == 0 && (p.ts_begin & 127)) {   // timestamp rows of the group that straddles ts_begin
      const float4 r = __ldcg(reinterpret_cast<const float4*>(p.part_extra + (size_t)b * 4));
      lse_fold(acc[1], r.x, __float_as_int(r.y), r.z);
    }
  }
  __syncthreads();
  const int cur = s_cur;
  const int np = cur + 1;
  float pe_v[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int c = tid + j * 256;
    pe_v[j] = (np < p.n_ctx && c < p.d) ? __ldg(p.pos_emb + (size_t)np * p.d + c) : 0.f;
  }
  if (p.sample) {
    const int n_cls = ts_on ? 2 : 1;
    for (int k = 0; k < n_cls; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, acc[k].best, o), os = __shfl_xor_sync(0xffffffffu, acc[k].se, o);
        const int oa = __shfl_xor_sync(0xffffffffu, acc[k].arg, o);
        lse_merge(acc[k], om, oa, os);
      }
      if (lane == 0) s_val[k][warp] = acc[k].best, s_idx[k][warp] = acc[k].arg, s_sum[k][warp] = acc[k].se;
    }
    __syncthreads();
    if (tid == 0) {
      LseAcc t[2];
      for (int k = 0; k < n_cls; ++k) {
        t[k] = LseAcc{s_val[k][0], s_sum[k][0], s_idx[k][0]};
        for (int w = 1; w < 8; ++w) lse_merge(t[k], s_val[k][w], s_idx[k][w], s_sum[k][w]);
      }
      float logprob;
      int A;
      if (p.chosen) {
        // temperature > 0: sample_rows_kernel drew the token from the filtered logits (mass rule included)
        A = __ldcg(p.chosen + b);
        logprob = __ldcg(p.chosen_logprob + b);
      } else if (!ts_on) {
        A = t[0].arg;
        logprob = -logf(t[0].se);   // the chosen logit is the maximum
      } else {
        // upstream ApplyTimestampRules, last rule: if the probability mass over the timestamps is above every text token,
        // the text tokens are suppressed:  logsumexp(ts) > max(text)  <=>  M_ts + log S_ts > M_text
        const float lse_ts = t[1].best > -INFINITY ? t[1].best + logf(t[1].se) : -INFINITY;
        if (lse_ts > t[0].best) {
          A = t[1].arg;
          logprob = -logf(t[1].se);
        } else {
          const bool text = t[0].best >= t[1].best;   // equal logits: the lower index (text) wins, as argmax does
          A = text ? t[0].arg : t[1].arg;
          LseAcc all = t[0];
          lse_merge(all, t[1].best, t[1].arg, t[1].se);
          logprob = -logf(all.se);
        }
      }
      const bool ended = prev_tok == p.eot;
      if (!ended) p.sum_logprob[b] = slp_old + logprob;
      const int next = ended ? p.eot : A;
      trow[cur + 1] = next;
      p.done[b] = next == p.eot;
      s_tok = next;
      if (ts_on) {   // rule state of the next step: sampled tokens are positions n_initial .. cur + 1
        int4 st = p.ts_state[b];
        const int len = cur + 2 - p.n_initial;
        const bool last_ts = next >= p.ts_begin;
        const bool penult_ts = len < 2 || prev_tok >= p.ts_begin;
        if (last_ts) st.z = next;                            // the newest timestamp of the sequence
        st.x = (last_ts && penult_ts ? 1 : 0) | (last_ts && !penult_ts ? 2 : 0);
        st.y = st.z ? ((last_ts && !penult_ts) ? st.z : st.z + 1) : 0;
        p.ts_state[b] = st;
      }
    }
    __syncthreads();
  }
  if (np < p.n_ctx) {
    int tok = s_tok;
    tok = tok < 0 ? 0 : (tok >= p.V ? p.V - 1 : tok);
    const __half* e = p.tok_emb + (size_t)tok * p.d;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int c = tid + j * 256;
      if (c < p.d) p.x[(size_t)b * p.d + c] = __half2float(__ldg(e + c)) + pe_v[j];
    }
  }
  trace.end();
}

int launch_step_finish(const FinishDesc& d, cudaStream_t st, int64_t* launches) {
  const cudaError_t le = launch_pdl(step_finish_kernel, dim3(d.Mb), dim3(256), 0, st, d);
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// ---- LayerNorm of the decoder rows, one warp per row (wider models) -----------------------------------------------------------------
// fp32 statistics in one pass (sum, sum of squares), as the fused input stage of the skinny GEMM computes them; gamma / beta are
// requested before the wait. d <= 1280: at most 10 float4 per lane.
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      int Mb, int d, __half* __restrict__ out16, const DecodeState* state) {
  TraceScope trace(state, 150);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int n4 = d >> 2;
  float4 g[10], bt[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < n4 ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = c < n4 ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ptx::grid_dep_launch();
  ptx::grid_dep_sync();
  if (row < Mb) {
    float4 v[10];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < n4 ? ld_x4(x + (size_t)row * d + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    sum = warp_sum(sum), sq = warp_sum(sq);
    const float mean = sum / (float)d;
    const float rstd = rsqrtf(fmaxf(sq / (float)d - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int c = lane + 32 * i;
      if (c < n4) {
        const __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g[i].x + bt[i].x, (v[i].y - mean) * rstd * g[i].y + bt[i].y);
        const __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g[i].z + bt[i].z, (v[i].w - mean) * rstd * g[i].w + bt[i].w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out16 + (size_t)row * d + c * 4) = u;
      }
    }
  }
  trace.end();
}
int launch_ln_rows(const float* x, const float* gamma, const float* beta, int Mb, int d, __half* out16, const DecodeState* state,
                   cudaStream_t st, int64_t* launches) {
  if (Mb < 1 || d % 4 != 0 || d > 1280) {
    set_error("ln_rows: unsupported shape Mb=%d d=%d", Mb, d);
    return -1;
  }
  const cudaError_t le = launch_pdl_n(ln_rows_kernel, dim3((Mb + 7) / 8), dim3(256), 0, st, x, gamma, beta, Mb, d, out16, state);
  if (launches) *launches += 1;
  WB_CUDA_OK(le);
  return 0;
}

// One thread spins on %globaltimer for `ns` nanoseconds (bounded): staggers the sub-batch chains of a decode so that the
// latency-bound layer-block kernels of one sub-batch run under the attention stream of the other instead of in lockstep.
__global__ void delay_kernel(unsigned long long ns) {
  const unsigned long long t0 = globaltimer();
  while (globaltimer() - t0 < ns) {
  }
}
int launch_delay(unsigned long long ns, cudaStream_t st, int64_t* launches) {
  if (ns > 1000000ull) ns = 1000000ull;
  delay_kernel<<<1, 1, 0, st>>>(ns);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- temperature sampling (upstream GreedyDecoder.update with temperature > 0) ---------------------------------------------------------
// Upstream draws `Categorical(logits = logits / temperature).sample()` from torch's global generator, which no other
// implementation can reproduce; this library defines the draw as the Gumbel-max form of the same distribution with a
// counter-based generator, so that a CPU restatement sees the same noise:
//   token = argmax_v ( logit_v / T + g_v ),  g_v = -log(-log(u_v)),  u_v = (top 23 bits of r_v + 0.5) / 2^23,
//   r_v = splitmix64(key + v),  key = splitmix64(seed ^ splitmix64((sample << 32) | position))
// sample = index of the sequence within the call, position = index of the token being drawn. The log-probability that goes
// into sum_logprobs is log_softmax(logits)[token] at temperature 1, as upstream.
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct SampleBest {
  float pv;    // perturbed value logit / T + g
  float lg;    // the logit itself
  int idx;
};
__device__ __forceinline__ void sample_take(SampleBest& a, float pv, float lg, int idx) {
  if (pv > a.pv || (pv == a.pv && idx < a.idx)) a.pv = pv, a.lg = lg, a.idx = idx;
}
// One CTA per sequence over the stored, filtered logits row. Text rows [0, ts_begin) and timestamp rows [ts_begin, V) keep
// separate (max, sum-exp, best perturbed) records so that the last timestamp rule (probability mass over the timestamps above
// every text token -> text suppressed; upstream ApplyTimestampRules) is applied here; ts_begin >= V turns it off.
__global__ void __launch_bounds__(256) sample_rows_kernel(const float* __restrict__ logits, int V, int ts_begin, float inv_temperature,
                                                          unsigned long long seed, const DecodeState* state, int32_t* __restrict__ chosen,
                                                          float* __restrict__ chosen_logprob) {
  __shared__ float s_m[2][8], s_s[2][8], s_pv[2][8], s_lg[2][8];
  __shared__ int s_ix[2][8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (size_t)b * V;
  const int position = ld_state(&state->cur_len) + 1;
  const unsigned long long key = splitmix64(seed ^ splitmix64(((unsigned long long)b << 32) | (unsigned int)position));
  float m[2] = {-INFINITY, -INFINITY}, ss[2] = {0.f, 0.f};
  SampleBest best[2] = {{-INFINITY, -INFINITY, 0x7fffffff}, {-INFINITY, -INFINITY, 0x7fffffff}};
  for (int i = tid; i < V; i += 256) {
    const float x = row[i];
    if (x == -INFINITY) continue;
    const int c = i >= ts_begin ? 1 : 0;
    if (x > m[c]) {
      ss[c] = ss[c] * expf(m[c] - x) + 1.0f;
      m[c] = x;
    } else {
      ss[c] += expf(x - m[c]);
    }
    const unsigned long long r = splitmix64(key + (unsigned long long)i);
    const float u = ((float)(unsigned int)(r >> 41) + 0.5f) * (1.0f / 8388608.0f);   // 23 bits: strictly inside (0, 1) in fp32
    const float g = -logf(-logf(u));
    sample_take(best[c], x * inv_temperature + g, x, i);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m[c], o), os = __shfl_xor_sync(0xffffffffu, ss[c], o);
      const float nm = fmaxf(m[c], om);
      ss[c] = (m[c] > -INFINITY ? ss[c] * expf(m[c] - nm) : 0.f) + (om > -INFINITY ? os * expf(om - nm) : 0.f);
      m[c] = nm;
      const float opv = __shfl_xor_sync(0xffffffffu, best[c].pv, o), olg = __shfl_xor_sync(0xffffffffu, best[c].lg, o);
      const int oix = __shfl_xor_sync(0xffffffffu, best[c].idx, o);
      sample_take(best[c], opv, olg, oix);
    }
    if (lane == 0) s_m[c][warp] = m[c], s_s[c][warp] = ss[c], s_pv[c][warp] = best[c].pv, s_lg[c][warp] = best[c].lg, s_ix[c][warp] = best[c].idx;
  }
  __syncthreads();
  if (tid == 0) {
    float M[2], S[2];
    SampleBest B2[2];
    for (int c = 0; c < 2; ++c) {
      M[c] = -INFINITY, S[c] = 0.f;
      B2[c] = SampleBest{-INFINITY, -INFINITY, 0x7fffffff};
      for (int w = 0; w < 8; ++w) M[c] = fmaxf(M[c], s_m[c][w]);
      for (int w = 0; w < 8; ++w) {
        S[c] += s_m[c][w] > -INFINITY ? s_s[c][w] * expf(s_m[c][w] - M[c]) : 0.f;
        sample_take(B2[c], s_pv[c][w], s_lg[c][w], s_ix[c][w]);
      }
    }
    const float lse_ts = M[1] > -INFINITY ? M[1] + logf(S[1]) : -INFINITY;
    int tok;
    float lp;
    if (lse_ts > M[0]) {                      // timestamps only
      tok = B2[1].idx, lp = B2[1].lg - lse_ts;
    } else {
      const float Mx = fmaxf(M[0], M[1]);
      const float Sx = (M[0] > -INFINITY ? S[0] * expf(M[0] - Mx) : 0.f) + (M[1] > -INFINITY ? S[1] * expf(M[1] - Mx) : 0.f);
      SampleBest all = B2[0];
      sample_take(all, B2[1].pv, B2[1].lg, B2[1].idx);
      tok = all.idx, lp = all.lg - (Mx + logf(Sx));
    }
    chosen[b] = tok, chosen_logprob[b] = lp;
  }
}
int launch_sample_rows(const float* logits, int Mb, int V, int ts_begin, float temperature, unsigned long long seed, const DecodeState* state,
                       int32_t* chosen, float* chosen_logprob, cudaStream_t st, int64_t* launches) {
  if (!(temperature > 0.f)) {
    set_error("sample_rows: temperature must be positive");
    return -1;
  }
  sample_rows_kernel<<<Mb, 256, 0, st>>>(logits, V, ts_begin, 1.0f / temperature, seed, state, chosen, chosen_logprob);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// softmax(logits row)[token] per sequence (upstream no_speech_probs: the unfiltered logits at the <|startoftranscript|> position)
__global__ void __launch_bounds__(256) row_token_prob_kernel(const float* __restrict__ logits, int V, int token, float* __restrict__ prob) {
  __shared__ float s_m[8], s_s[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (size_t)b * V;
  float m = -INFINITY, ss = 0.f;
  for (int i = tid; i < V; i += 256) {
    const float x = row[i];
    if (x > m) {
      ss = ss * expf(m - x) + 1.0f;
      m = x;
    } else if (x > -INFINITY) {
      ss += expf(x - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, ss, o);
    const float nm = fmaxf(m, om);
    ss = (m > -INFINITY ? ss * expf(m - nm) : 0.f) + (om > -INFINITY ? os * expf(om - nm) : 0.f);
    m = nm;
  }
  if (lane == 0) s_m[warp] = m, s_s[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    float M = -INFINITY, S = 0.f;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
    for (int w = 0; w < 8; ++w) S += s_m[w] > -INFINITY ? s_s[w] * expf(s_m[w] - M) : 0.f;
    prob[b] = expf(row[token] - M) / S;
  }
}
int launch_row_token_prob(const float* logits, int Mb, int V, int token, float* prob, cudaStream_t st, int64_t* launches) {
  if (token < 0 || token >= V) {
    set_error("row_token_prob: token %d outside the vocabulary", token);
    return -1;
  }
  row_token_prob_kernel<<<Mb, 256, 0, st>>>(logits, V, token, prob);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// state->cur_len = v (the no-speech probe re-runs the step of one prompt position: see decode_prompt in model.cu)
__global__ void set_cur_len_kernel(DecodeState* state, int v) { state->cur_len = v; }
int launch_set_cur_len(DecodeState* state, int v, cudaStream_t st, int64_t* launches) {
  set_cur_len_kernel<<<1, 1, 0, st>>>(state, v);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- beam search support -----------------------------------------------------------------------------------------------------------
// Top-k of one logits row (k <= 8) and its log-sum-exp: every thread keeps a sorted top-k of its strided elements and an
// online (max, sum-exp); the candidates are merged through shared memory by one warp. Output log-probabilities = logit - LSE
// (upstream: F.log_softmax(logits).topk(beam_size + 1)).
__global__ void __launch_bounds__(256) topk_logprobs_kernel(const float* __restrict__ logits, int V, int k, int ts_begin,
                                                            float* __restrict__ top_lp, int32_t* __restrict__ top_idx) {
  __shared__ float c_val[256 * 8];
  __shared__ int c_idx[256 * 8];
  __shared__ float s_m[8], s_s[8];
  __shared__ float s_mt[8];
  __shared__ int s_lo;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (size_t)b * V;
  int lo = 0;
  if (ts_begin < V) {
    // upstream ApplyTimestampRules, last rule: logsumexp(timestamps) > max(text)  <=>  M_ts + log S_ts > M_text  =>  text rows masked
    float mt = -INFINITY, ms = -INFINITY, ss = 0.f;
    for (int i = tid; i < ts_begin; i += 256) mt = fmaxf(mt, row[i]);
    for (int i = ts_begin + tid; i < V; i += 256) {
      const float x = row[i];
      if (x > ms) {
        ss = ss * expf(ms - x) + 1.0f;
        ms = x;
      } else if (x > -INFINITY) {
        ss += expf(x - ms);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
      const float om = __shfl_xor_sync(0xffffffffu, ms, o), os = __shfl_xor_sync(0xffffffffu, ss, o);
      const float nm = fmaxf(ms, om);
      ss = (ms > -INFINITY ? ss * expf(ms - nm) : 0.f) + (om > -INFINITY ? os * expf(om - nm) : 0.f);
      ms = nm;
    }
    if (lane == 0) s_mt[warp] = mt, s_m[warp] = ms, s_s[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      float MT = -INFINITY, M = -INFINITY;
      for (int w = 0; w < 8; ++w) MT = fmaxf(MT, s_mt[w]), M = fmaxf(M, s_m[w]);
      float S = 0.f;
      for (int w = 0; w < 8; ++w) S += s_m[w] > -INFINITY ? s_s[w] * expf(s_m[w] - M) : 0.f;
      const float lse_ts = M > -INFINITY ? M + logf(S) : -INFINITY;
      s_lo = lse_ts > MT ? ts_begin : 0;
    }
    __syncthreads();
    lo = s_lo;
    __syncthreads();   // s_m / s_s are reused below
  }
  float tv[8];
  int ti[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) tv[j] = -INFINITY, ti[j] = 0x7fffffff;
  float m = -INFINITY, ssum = 0.f;
  // eight independent loads per thread and batch: with one load per iteration the loop ran at one L2 / DRAM latency per element
  // (203 dependent round trips per thread: 132 us per launch at 40 rows, 12 % of a beam-search step)
  for (int i0 = lo + tid; i0 < V; i0 += 256 * 8) {
    float xs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * 256;
      xs[u] = i < V ? __ldcg(row + i) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float x = xs[u];
      const int i = i0 + u * 256;
      if (x > m) {
        ssum = ssum * expf(m - x) + 1.0f;
        m = x;
      } else if (x > -INFINITY) {
        ssum += expf(x - m);
      }
      if (x > tv[7]) {   // insert into the sorted list (descending; earlier index wins ties)
        tv[7] = x, ti[7] = i;
#pragma unroll
        for (int j = 7; j > 0; --j) {
          if (tv[j] > tv[j - 1]) {
            const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
            const int iv = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = iv;
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) c_val[tid * 8 + j] = tv[j], c_idx[tid * 8 + j] = ti[j];
  // block LSE
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, ssum, o);
    const float nm = fmaxf(m, om);
    ssum = (m > -INFINITY ? ssum * expf(m - nm) : 0.f) + (om > -INFINITY ? os * expf(om - nm) : 0.f);
    m = nm;
  }
  if (lane == 0) s_m[warp] = m, s_s[warp] = ssum;
  __syncthreads();
  if (warp == 0) {
    float M = -INFINITY;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
    float S = 0.f;
    for (int w = 0; w < 8; ++w) S += s_m[w] > -INFINITY ? s_s[w] * expf(s_m[w] - M) : 0.f;
    const float lse = M + logf(S);
    for (int r = 0; r < k; ++r) {
      float best = -INFINITY;
      int bi = 0x7fffffff, bslot = -1;
      for (int c = lane; c < 256 * 8; c += 32) {
        const float v = c_val[c];
        const int ix = c_idx[c];
        if (v > best || (v == best && ix < bi)) best = v, bi = ix, bslot = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o), os = __shfl_xor_sync(0xffffffffu, bslot, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi, bslot = os;
      }
      __syncwarp();   // every lane's scan of c_val is complete before lane 0 retires the winner (shuffles alone do not order shared memory)
      if (lane == 0) {
        top_lp[(size_t)b * 8 + r] = best - lse;
        top_idx[(size_t)b * 8 + r] = bi;
        if (bslot >= 0) c_val[bslot] = -INFINITY;
      }
      __syncwarp();
    }
  }
}

int launch_topk_logprobs(const float* logits, int Mb, int V, int k, int ts_begin, float* top_logprob, int32_t* top_index, cudaStream_t st,
                         int64_t* launches) {
  if (k < 1 || k > 8) {
    set_error("topk: k=%d out of range [1,8]", k);
    return -1;
  }
  topk_logprobs_kernel<<<Mb, 256, 0, st>>>(logits, V, k, ts_begin, top_logprob, top_index);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

// dst[l][b][0..cur_len] = src[l][source[b]][0..cur_len] for K and V of every layer (upstream rearrange_kv_cache)
struct ReorderArgs {
  const __half* src_k[32];
  const __half* src_v[32];
  __half* dst_k[32];
  __half* dst_v[32];
};
__global__ void __launch_bounds__(256) reorder_kv_kernel(ReorderArgs a, int n_ctx, int d, const int32_t* __restrict__ source,
                                                         const DecodeState* state) {
  const int b = blockIdx.x, l = blockIdx.y >> 1, is_v = blockIdx.y & 1;
  const int n_rows = ld_state(&state->cur_len) + 1;
  const int sb = source[b];
  const uint4* src = reinterpret_cast<const uint4*>((is_v ? a.src_v[l] : a.src_k[l]) + (size_t)sb * n_ctx * d);
  uint4* dst = reinterpret_cast<uint4*>((is_v ? a.dst_v[l] : a.dst_k[l]) + (size_t)b * n_ctx * d);
  const int n = n_rows * d / 8;
  for (int i = threadIdx.x; i < n; i += 256) dst[i] = src[i];
}

int launch_reorder_kv(const __half* const* src_k, const __half* const* src_v, __half* const* dst_k, __half* const* dst_v, int n_layer,
                      int Mb, int n_ctx, int d, const int32_t* source, const DecodeState* state, cudaStream_t st, int64_t* launches) {
  if (n_layer > 32) {
    set_error("reorder_kv: too many layers");
    return -1;
  }
  ReorderArgs a{};
  for (int l = 0; l < n_layer; ++l) a.src_k[l] = src_k[l], a.src_v[l] = src_v[l], a.dst_k[l] = dst_k[l], a.dst_v[l] = dst_v[l];
  reorder_kv_kernel<<<dim3(Mb, n_layer * 2), 256, 0, st>>>(a, n_ctx, d, source, state);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
