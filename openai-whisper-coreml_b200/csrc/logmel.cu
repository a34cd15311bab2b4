// Log-mel front end on the GPU: kernels + launchers. See logmel.cuh for the per-phase math and the reference map.
//
// Data layout in HBM
//   audio      f32 [B][480000] (throughput path)  or  f64 [B][480400] (legacy ABI, pads ignored on read)
//   logspec    T   [B][80][3000]   log10(max(mel,1e-10))            (intermediate, lib.rs:76)
//   gmax       ordered-uint per chunk                               (lib.rs:82-88)
//   out        T   [B][80][3000]   (max(l, g-8)+4)/4                (lib.rs:96, layout lib.rs:116-121)
//   melT       f16 [B][3002][80]   same values, time-major with one zero row either side: the A operand of the
//                                  conv1 implicit GEMM (row t of the im2col matrix = 240 contiguous halves at t*80)
//
// Algorithmic bytes per chunk (f32 path): 480000*4 read + 80*3000*4 written = 2.88 MB.
#include "logmel.cuh"

#include <math.h>
#include <string.h>

#include "mel80_sparse.inc"

namespace wb {

template <typename T>
struct OrderedMax;
template <>
struct OrderedMax<float> {
  using U = unsigned int;
  __device__ static U enc(float f) { return float_to_ordered(f); }
  __device__ static float dec(U u) { return ordered_to_float(u); }
};
template <>
struct OrderedMax<double> {
  using U = unsigned long long;
  __device__ static U enc(double f) {
    U u = (U)__double_as_longlong(f);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
  }
  __device__ static double dec(U u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
  }
};

// One CTA = F consecutive frames of one chunk.  grid = (3000 / F, B).  MINB CTAs per SM bound the registers.
template <typename T, int F, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) logmel_kernel(const T* __restrict__ audio, size_t chunk_stride, int chunk_off,
                                                    const LogmelTables<T>* __restrict__ gtab, T* __restrict__ logspec,
                                                    typename OrderedMax<T>::U* __restrict__ gmax, long long stream_samples) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LogmelSmem<T, F>& sm = *reinterpret_cast<LogmelSmem<T, F>*>(smem_raw);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, b = blockIdx.y;
  const T* clip = audio + (size_t)b * chunk_stride + chunk_off;

  // tables -> smem (8-byte copies)
  {
    using Smem = LogmelSmem<T, F>;
    static_assert(sizeof(LogmelTables<T>) % 8 == 0 && offsetof(Smem, tab) % 8 == 0, "table copy is done in 8-byte words");
    const uint2* src = reinterpret_cast<const uint2*>(gtab);
    uint2* dst = reinterpret_cast<uint2*>(&sm.tab);
    for (int i = tid; i < (int)(sizeof(LogmelTables<T>) / 8); i += NT) dst[i] = __ldg(src + i);
  }
  // samples of this tile, reflection folded into the index map (lib.rs:34-40); padded coordinate p0 + s
  const int p0 = tile * F * WB_HOP;
  if (stream_samples < 0) {
    // a tile that does not touch either reflected end is one contiguous run of the clip: 16-byte loads (float4 / double2),
    // 16-byte shared-memory stores (a pad group never splits a vector: 160 and 16 are multiples of the vector width)
    constexpr int VW = 16 / (int)sizeof(T);
    static_assert(tile_samples<F>() % VW == 0 && kSampGroup % VW == 0 && kSampPad % VW == 0, "vector sample load");
    const T* run = clip + (p0 - 200);
    if (p0 >= 200 && p0 - 200 + tile_samples<F>() <= WB_N_SAMPLES && (reinterpret_cast<uintptr_t>(run) & 15) == 0) {
      const uint4* src = reinterpret_cast<const uint4*>(run);
      for (int v = tid; v < tile_samples<F>() / VW; v += NT)
        *reinterpret_cast<uint4*>(&sm.region0[samp_index(v * VW)]) = __ldg(src + v);
    } else {
      for (int s = tid; s < tile_samples<F>(); s += NT) sm.region0[samp_index(s)] = clip[reflect_index(p0 + s)];
    }
  } else {
    // stream mode (upstream whisper/audio.py log_mel_spectrogram on a whole recording followed by 30 s of zeros): item b holds
    // frames [3000 b, 3000 b + 3000) of ONE signal of stream_samples samples; the reflection exists only at the start of the
    // signal, everything past its end reads as zero (the appended zeros, and their reflection)
    const long long i0 = (long long)b * WB_N_SAMPLES + p0 - 200;
    for (int s = tid; s < tile_samples<F>(); s += NT) {
      long long i = i0 + s;
      if (i < 0) i = -i;
      sm.region0[samp_index(s)] = i < stream_samples ? audio[i] : (T)0;
    }
  }
  __syncthreads();

  logmel_phase_a<T, F>(sm, tid);
  __syncthreads();
  for (int task = tid; task < F * 25; task += NT) logmel_phase_b<T, F>(sm, task);
  __syncthreads();
  // C1: a thread keeps its bin pair k and walks the frames, so that the slot arithmetic and the twiddle of a bin are computed once
  {
    constexpr int FS = NT / 100;                       // frames in flight
    if (tid < 100 * FS) {
      const int k = tid % 100 + 1;
      for (int fl = tid / 100; fl < F; fl += FS) logmel_phase_c1<T, F>(sm, fl, k);
    }
  }
  __syncthreads();

  // C2: a thread keeps its mel band and accumulates NF frames at once (one weight load per bin for all of them); the values go
  // through shared memory (the FFT buffer is free by now) so that the stores run along the frames of a band
  T vmax = (T)-1e30;
  T* stage = reinterpret_cast<T*>(sm.y);               // [80][F + 1]
  {
    constexpr int GF = NT / WB_N_MELS;                 // threads per band
    constexpr int NF = (F + GF - 1) / GF;              // frames per thread: fl = g + GF * n
    static_assert(sizeof(sm.y) >= sizeof(T) * WB_N_MELS * (F + 1), "stage fits the FFT buffer");
    if (tid < WB_N_MELS * GF) {
      const int i = tid / GF, g = tid - i * GF;
      T v[NF];
      logmel_phase_c2<T, F, NF>(sm, i, g, GF, v);
#pragma unroll
      for (int n = 0; n < NF; ++n) {
        const int fl = g + GF * n;
        if (fl < F) {
          stage[i * (F + 1) + fl] = v[n];
          vmax = v[n] > vmax ? v[n] : vmax;
        }
      }
    }
  }
  __syncthreads();
  for (int task = tid; task < F * WB_N_MELS; task += NT) {
    const int i = task / F, fl = task - i * F;
    logspec[((size_t)b * WB_N_MELS + i) * WB_N_FRAMES + tile * F + fl] = stage[i * (F + 1) + fl];
  }
  // block max -> one atomic per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, vmax, o);
    vmax = other > vmax ? other : vmax;
  }
  if ((tid & 31) == 0) sm.red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    T m = sm.red[0];
    for (int w = 1; w < NT / 32; ++w) m = sm.red[w] > m ? sm.red[w] : m;
    atomicMax(&gmax[stream_samples < 0 ? b : 0], OrderedMax<T>::enc(m));   // stream mode: one maximum for the whole signal
  }
}

// (max(l, g-8)+4)/4 (lib.rs:96).  One CTA = 120 frames x 80 mels of one chunk.  grid = (25, B).
// Writes the [80][3000] layout of the reference (optional) and the time-major fp16 layout for conv1 (optional).
#ifndef WB_NORM_FRAMES
#define WB_NORM_FRAMES 120
#endif
constexpr int kNormFrames = WB_NORM_FRAMES;
template <typename T>
__global__ void __launch_bounds__(256) logmel_normalize_kernel(const T* __restrict__ logspec,
                                                               const typename OrderedMax<T>::U* __restrict__ gmax,
                                                               T* __restrict__ out, __half* __restrict__ melT) {
  __shared__ __half tile[kNormFrames][WB_N_MELS + 1];   // odd row stride: lanes four frames apart land in 16 different banks
  const int b = blockIdx.y, f0 = blockIdx.x * kNormFrames;
  const T floor_v = OrderedMax<T>::dec(gmax[b]) - (T)8.0;
  // 16 bytes (four floats / two doubles) of a band per thread and iteration: rows of 3000 and tiles of kNormFrames frames keep
  // every vector aligned
  constexpr int VW = 16 / (int)sizeof(T), VPR = kNormFrames / VW;
  static_assert(kNormFrames % VW == 0 && WB_N_FRAMES % VW == 0, "vector width divides the tile and the row");
  struct alignas(16) Vec {
    T e[VW];
  };
  const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;   // a caller's device buffer may sit at any float boundary
#pragma unroll 2
  for (int idx = threadIdx.x; idx < WB_N_MELS * VPR; idx += 256) {
    const int i = idx / VPR, f = (idx - i * VPR) * VW;
    const size_t g = ((size_t)b * WB_N_MELS + i) * WB_N_FRAMES + f0 + f;
    Vec v = *reinterpret_cast<const Vec*>(logspec + g);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      T x = v.e[e];
      x = x > floor_v ? x : floor_v;
      v.e[e] = (x + (T)4.0) / (T)4.0;
    }
    if (out) {
      if (out_vec) {
        *reinterpret_cast<Vec*>(out + g) = v;
      } else {
#pragma unroll
        for (int e = 0; e < VW; ++e) out[g + e] = v.e[e];
      }
    }
    if (melT) {
#pragma unroll
      for (int e = 0; e < VW; ++e) tile[f + e][i] = __float2half_rn((float)v.e[e]);
    }
  }
  if (melT) {
    __syncthreads();
    __half* dst = melT + ((size_t)b * (WB_N_FRAMES + 2) + 1 + f0) * WB_N_MELS;
    for (int idx = threadIdx.x; idx < WB_N_MELS * kNormFrames; idx += 256) {
      const int f = idx / WB_N_MELS, i = idx - f * WB_N_MELS;
      dst[idx] = tile[f][i];
    }
  }
}

// mel [B][80][3000] f32 (already normalised, encoder.prediction's input) -> melT fp16 [B][3002][80]
__global__ void __launch_bounds__(256) mel_transpose_kernel(const float* __restrict__ mel, __half* __restrict__ melT) {
  __shared__ __half tile[kNormFrames][WB_N_MELS + 2];
  const int b = blockIdx.y, f0 = blockIdx.x * kNormFrames;
  for (int idx = threadIdx.x; idx < WB_N_MELS * kNormFrames; idx += 256) {
    const int i = idx / kNormFrames, f = idx - i * kNormFrames;
    tile[f][i] = __float2half_rn(mel[((size_t)b * WB_N_MELS + i) * WB_N_FRAMES + f0 + f]);
  }
  __syncthreads();
  __half* dst = melT + ((size_t)b * (WB_N_FRAMES + 2) + 1 + f0) * WB_N_MELS;
  for (int idx = threadIdx.x; idx < WB_N_MELS * kNormFrames; idx += 256) {
    const int f = idx / WB_N_MELS, i = idx - f * WB_N_MELS;
    dst[idx] = tile[f][i];
  }
}

template <typename U>
__global__ void fill_kernel(U* p, U v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- host side -----------------------------------------------------------------------------------------------------
template <typename T>
void build_logmel_tables(LogmelTables<T>& t) {
  const double PI = 3.14159265358979323846264338327950288;
  for (int i = 0; i < 400; ++i) t.window[i] = (T)((1.0 - cos(((double)i * 2.0 * PI) / 400.0)) / 2.0);   // lib.rs:26
  for (int k = 0; k < 200; ++k) t.tw200[k] = {(T)cos(2.0 * PI * k / 200.0), (T)(-sin(2.0 * PI * k / 200.0))};
  for (int k = 0; k <= 100; ++k) t.tw400[k] = {(T)cos(2.0 * PI * k / 400.0), (T)(-sin(2.0 * PI * k / 400.0))};
  for (int i = 0; i < MEL80_NNZ; ++i) {
    float f;
    const uint32_t bits = MEL80_W_BITS_H[i];
    memcpy(&f, &bits, 4);
    t.melw[i] = (T)f;                                                                                    // lib.rs:65
  }
  t.melw[MEL80_NNZ] = (T)0;
  for (int i = 0; i < 80; ++i) {
    t.mel_lo[i] = MEL80_LO_H[i];
    t.mel_cnt[i] = MEL80_CNT_H[i];
    t.mel_off[i] = MEL80_OFF_H[i];
  }
}
template void build_logmel_tables<float>(LogmelTables<float>&);
template void build_logmel_tables<double>(LogmelTables<double>&);

template <typename T>
struct LogmelCfg;
template <>
struct LogmelCfg<float> {
  // 15 frames x 256 threads, four CTAs (32 warps) per SM under a 64-register cap (36 bytes of spills): 46 KB of shared memory
  // per CTA. Measured for 32 chunks, whole front end: 30 frames x 256 threads (two CTAs per SM, 80 registers) 152 us,
  // 15 x 192 and 24 x 256 (24 warps) 132, 30 x 512 133, this 127 (profiles/r02_logmel_variants.log)
#ifndef WB_LM_F
#define WB_LM_F 15
#endif
#ifndef WB_LM_NT
#define WB_LM_NT 256
#endif
#ifndef WB_LM_MINB
#define WB_LM_MINB 4
#endif
  static constexpr int F = WB_LM_F, NT = WB_LM_NT, MINB = WB_LM_MINB;
};
template <>
struct LogmelCfg<double> {
  static constexpr int F = 15, NT = 128, MINB = 4;   // 128 registers, two CTAs per SM (92 KB of shared memory)
};

// Enqueue: logspec/gmax are scratch ([B][80][3000] T and [B] ordered). Returns 0 or -2 (error text recorded).
template <typename T>
int launch_logmel(const T* audio, size_t chunk_stride, int chunk_off, int B, const LogmelTables<T>* dtab, T* logspec,
                  void* gmax, T* out, __half* melT, cudaStream_t st, int64_t* launches) {
  using Cfg = LogmelCfg<T>;
  using U = typename OrderedMax<T>::U;
  static_assert(WB_N_FRAMES % Cfg::F == 0, "tile must divide 3000 frames");
  static_assert(Cfg::NT >= Cfg::F * 8, "phase A needs 8 threads per frame");
  const size_t smem = sizeof(LogmelSmem<T, Cfg::F>);
  auto kern = logmel_kernel<T, Cfg::F, Cfg::NT, Cfg::MINB>;
  WB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fill_kernel<U><<<(B + 127) / 128, 128, 0, st>>>((U*)gmax, (U)0, B);
  kern<<<dim3(WB_N_FRAMES / Cfg::F, B), Cfg::NT, smem, st>>>(audio, chunk_stride, chunk_off, dtab, logspec, (U*)gmax, -1ll);
  logmel_normalize_kernel<T><<<dim3(WB_N_FRAMES / kNormFrames, B), 256, 0, st>>>(logspec, (const U*)gmax, out, melT);
  if (launches) *launches += 3;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}
template int launch_logmel<float>(const float*, size_t, int, int, const LogmelTables<float>*, float*, void*, float*,
                                  __half*, cudaStream_t, int64_t*);
template int launch_logmel<double>(const double*, size_t, int, int, const LogmelTables<double>*, double*, void*,
                                   double*, __half*, cudaStream_t, int64_t*);

// ---- stream mode: the log-mel of a whole recording, then 3000-frame segments at arbitrary frame offsets ------------------------------
// logspec [W][80][3000] unnormalised, one maximum for the whole signal (upstream normalises with the global maximum).
int launch_logmel_stream(const float* audio, long long n_samples, int W, const LogmelTables<float>* dtab, float* logspec, void* gmax,
                         cudaStream_t st, int64_t* launches) {
  using Cfg = LogmelCfg<float>;
  using U = OrderedMax<float>::U;
  const size_t smem = sizeof(LogmelSmem<float, Cfg::F>);
  auto kern = logmel_kernel<float, Cfg::F, Cfg::NT, Cfg::MINB>;
  WB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fill_kernel<U><<<1, 32, 0, st>>>((U*)gmax, (U)0, 1);
  kern<<<dim3(WB_N_FRAMES / Cfg::F, W), Cfg::NT, smem, st>>>(audio, 0, 0, dtab, logspec, (U*)gmax, n_samples);
  if (launches) *launches += 2;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}
// frames [frame0, frame0 + 3000) of the stream -> one normalised segment: out [80][3000] f32 (optional) and melT f16 [3002][80]
// (optional); frames past the W windows are zero (upstream pad_or_trim of the mel segment)
__global__ void __launch_bounds__(256) logmel_segment_kernel(const float* __restrict__ logspec, const unsigned int* __restrict__ gmax, int W,
                                                             long long frame0, float* __restrict__ out, __half* __restrict__ melT) {
  __shared__ __half tile[kNormFrames][WB_N_MELS + 2];
  const int f0 = blockIdx.x * kNormFrames;
  const float floor_v = OrderedMax<float>::dec(gmax[0]) - 8.0f;
  for (int idx = threadIdx.x; idx < WB_N_MELS * kNormFrames; idx += 256) {
    const int i = idx / kNormFrames, f = idx - i * kNormFrames;
    const long long gf = frame0 + f0 + f;
    const long long w = gf / WB_N_FRAMES;
    float v = 0.f;
    if (w < W) {
      v = logspec[((size_t)w * WB_N_MELS + i) * WB_N_FRAMES + (int)(gf - w * WB_N_FRAMES)];
      v = v > floor_v ? v : floor_v;
      v = (v + 4.0f) / 4.0f;
    }
    if (out) out[(size_t)i * WB_N_FRAMES + f0 + f] = v;
    if (melT) tile[f][i] = __float2half_rn(v);
  }
  if (melT) {
    __syncthreads();
    __half* dst = melT + ((size_t)1 + f0) * WB_N_MELS;
    for (int idx = threadIdx.x; idx < WB_N_MELS * kNormFrames; idx += 256) {
      const int f = idx / WB_N_MELS, i = idx - f * WB_N_MELS;
      dst[idx] = tile[f][i];
    }
  }
}
int launch_logmel_segment(const float* logspec, const void* gmax, int W, long long frame0, float* out, __half* melT, cudaStream_t st,
                          int64_t* launches) {
  logmel_segment_kernel<<<WB_N_FRAMES / kNormFrames, 256, 0, st>>>(logspec, (const unsigned int*)gmax, W, frame0, out, melT);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_mel_transpose(const float* mel, int B, __half* melT, cudaStream_t st, int64_t* launches) {
  mel_transpose_kernel<<<dim3(WB_N_FRAMES / kNormFrames, B), 256, 0, st>>>(mel, melT);
  if (launches) *launches += 1;
  WB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wb
