// Shared helpers for the whisper_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define WB_N_SAMPLES 480000      // 16000 * 30            (stft/src/lib.rs:37,112)
#define WB_N_FFT 400             //                        (lib.rs:24)
#define WB_HOP 160               //                        (lib.rs:50)
#define WB_N_FRAMES 3000         //                        (lib.rs:52,62)
#define WB_N_MELS 80             //                        (lib.rs:60)
#define WB_N_BINS 201            //                        (lib.rs:51)
#define WB_N_AUDIO_CTX 1500      //                        (whisper_to_cml.py:29)
#define WB_HEAD_DIM 64           // every Whisper size has d_head = 64

namespace wb {

// error plumbing: kernels never throw; host code records the first CUDA error in a thread-local string.
void set_error(const char* fmt, ...);
const char* get_error();

#define WB_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      wb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

// Per-kernel launch attributes (dynamic shared-memory limit, non-portable cluster size) are per device: the "already set"
// caches of the launchers are indexed by the current device, so one process may hold handles on several GPUs.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d >= kMaxDevices) ? 0 : d;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact-erf GELU, as upstream (torch.nn.functional.gelu default)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// order-preserving float <-> uint mapping, for atomicMax on floats of either sign
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(unsigned int u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

}  // namespace wb
