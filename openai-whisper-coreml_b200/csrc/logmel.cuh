// Fused reflect-pad + periodic-Hann + 400-point real FFT + |.|^2 + 80-band mel + log10 for Whisper's 30 s chunks.
//
// Replaces, on the GPU, the hot loops of the reference's Rust crate (/root/reference/stft/src/lib.rs):
//   reflect      lib.rs:34-40   folded into the sample loader's index map (no padded copy is materialised)
//   window+fft   lib.rs:42-47   shared-memory mixed-radix butterflies: real-400 = complex-200 = 25 x 8, then split pass
//   power        lib.rs:54      computed pairwise for bins k and 200-k from the same two half-length outputs
//   mel          lib.rs:60-69   banded-sparse (391 non-zeros of 16080), bins ascending = the dense loop's order
//   log10/floor  lib.rs:76      log10(max(x,1e-10)); per-chunk max folded in with one atomicMax per CTA
// The chunk-wide max (lib.rs:82-88) is the only cross-frame dependency, so the clamp/scale (lib.rs:96) lives in a
// second, elementwise kernel (logmel_normalize_kernel) that also emits the layouts the consumers want.
//
// No cuFFT. Templated on the arithmetic type: float for the throughput path, double for the legacy f64 ABI.
//
// Every phase is a __host__ __device__ function of (tid), so tests/host_emulate_logmel.cu can run the exact same
// index math on the CPU (this container has no GPU) before any GPU time is spent.
#pragma once
#include "common.cuh"

namespace wb {

template <typename T>
struct Cpx {
  T re, im;
};
template <typename T>
__host__ __device__ __forceinline__ Cpx<T> cadd(Cpx<T> a, Cpx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T>
__host__ __device__ __forceinline__ Cpx<T> csub(Cpx<T> a, Cpx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <typename T>
__host__ __device__ __forceinline__ Cpx<T> cmul(Cpx<T> a, Cpx<T> b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T>
__host__ __device__ __forceinline__ Cpx<T> mul_neg_i(Cpx<T> a) { return {a.im, -a.re}; }   // a * (-i)

// Tables (host-built in double, see logmel.cu). One copy in global memory per arithmetic type.
template <typename T>
struct LogmelTables {
  T window[400];          // periodic Hann, lib.rs:26
  Cpx<T> tw200[200];      // exp(-2 pi i k / 200)
  Cpx<T> tw400[101];      // exp(-2 pi i k / 400), k = 0..100
  T melw[392];            // 391 non-zero mel weights (f32 values widened exactly, lib.rs:65)
  int mel_lo[80];
  int mel_cnt[80];
  int mel_off[80];
};

// ---- shared-memory geometry ------------------------------------------------------------------------------------------
// Samples are stored with 16 pad words after every 160 so that consecutive frames (hop 160 = 0 mod 32 banks) start
// 16 banks apart. Y holds the 200 complex points of each frame at slot (j*9 + t) (j = 0..24, t = 0..7): stride 9
// keeps phase-B's 8-value reads conflict-free; frame stride 232.
constexpr int kSampGroup = 160, kSampPad = 16;
constexpr int kYFrame = 232;      // complex slots per frame (>= 24*9+8 = 224)
constexpr int kPFrame = 201;      // power slots per frame

__host__ __device__ constexpr int samp_index(int s) { return s + (s / kSampGroup) * kSampPad; }
template <int F>
__host__ __device__ constexpr int tile_samples() { return (F - 1) * WB_HOP + WB_N_FFT; }
template <int F>
__host__ __device__ constexpr int samp_words() { return samp_index(tile_samples<F>() - 1) + 1; }

template <typename T, int F>
struct LogmelSmem {
  // region 0: samples during load/phase A, then the power spectrum (phase C)
  static constexpr int kRegion0 = (samp_words<F>() > F * kPFrame ? samp_words<F>() : F * kPFrame);
  alignas(16) T region0[kRegion0 + 2];
  alignas(16) Cpx<T> y[F * kYFrame];
  alignas(16) LogmelTables<T> tab;
  T red[32];
};

// padded index p in [0, 480400) -> sample of the unpadded clip (reflect, lib.rs:34-40; stft.swift:10-11 offsets)
__host__ __device__ __forceinline__ int reflect_index(int p) {
  int a = p - 200;
  if (a < 0) a = -a;                                   // padded[i] = padded[400-i]  ->  audio[200-i]
  if (a >= WB_N_SAMPLES) a = 2 * (WB_N_SAMPLES - 1) - a;   // padded[480200+i] = audio[479998-i]
  return a;
}

// ---- 5-point and 8-point DFT butterflies -------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ void dft5(Cpx<T>& x0, Cpx<T>& x1, Cpx<T>& x2, Cpx<T>& x3, Cpx<T>& x4) {
  const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;   // cos(2pi/5), cos(4pi/5)
  const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;    // sin(2pi/5), sin(4pi/5)
  const Cpx<T> t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const Cpx<T> a1 = {x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im};
  const Cpx<T> a2 = {x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im};
  const Cpx<T> b1 = {s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
  const Cpx<T> b2 = {s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
  x0 = {x0.re + t1.re + t2.re, x0.im + t1.im + t2.im};
  x1 = {a1.re + b1.im, a1.im - b1.re};      // a1 - i b1
  x4 = {a1.re - b1.im, a1.im + b1.re};      // a1 + i b1
  x2 = {a2.re + b2.im, a2.im - b2.re};
  x3 = {a2.re - b2.im, a2.im + b2.re};
}

template <typename T>
__host__ __device__ __forceinline__ void dft4(Cpx<T> b0, Cpx<T> b1, Cpx<T> b2, Cpx<T> b3,
                                              Cpx<T>& o0, Cpx<T>& o1, Cpx<T>& o2, Cpx<T>& o3) {
  const Cpx<T> c0 = cadd(b0, b2), c1 = csub(b0, b2), c2 = cadd(b1, b3), c3 = mul_neg_i(csub(b1, b3));
  o0 = cadd(c0, c2);
  o2 = csub(c0, c2);
  o1 = cadd(c1, c3);
  o3 = csub(c1, c3);
}

// X[a] = sum_t y[t] * exp(-2 pi i t a / 8), in place
template <typename T>
__host__ __device__ __forceinline__ void dft8(Cpx<T> (&y)[8]) {
  const T r = (T)0.70710678118654752440;
  const Cpx<T> e0 = cadd(y[0], y[4]), e1 = cadd(y[1], y[5]), e2 = cadd(y[2], y[6]), e3 = cadd(y[3], y[7]);
  const Cpx<T> d0 = csub(y[0], y[4]);
  Cpx<T> d1 = csub(y[1], y[5]), d2 = csub(y[2], y[6]), d3 = csub(y[3], y[7]);
  d1 = {(d1.re + d1.im) * r, (d1.im - d1.re) * r};        // * (1 - i)/sqrt2
  d2 = mul_neg_i(d2);                                     // * (-i)
  d3 = {(d3.im - d3.re) * r, (-d3.re - d3.im) * r};       // * (-1 - i)/sqrt2
  dft4(e0, e1, e2, e3, y[0], y[2], y[4], y[6]);
  dft4(d0, d1, d2, d3, y[1], y[3], y[5], y[7]);
}

// ---- phase A: 25-point DFT (5 x 5) over m of z[t + 8 m], then the 200-point twiddle ---------------------------------------
// thread (fl, t): fl = frame within tile, t = 0..7.  z[n] = (x[2n] w[2n], x[2n+1] w[2n+1]).
template <typename T, int F>
__host__ __device__ __forceinline__ void logmel_phase_a(LogmelSmem<T, F>& sm, int tid) {
  const int fl = tid >> 3, t = tid & 7;
  if (fl >= F) return;
  Cpx<T> v[25];
#pragma unroll
  for (int m = 0; m < 25; ++m) {
    const int n = t + 8 * m;
    const int s = samp_index(fl * WB_HOP + 2 * n);
    v[m] = {sm.region0[s] * sm.tab.window[2 * n], sm.region0[s + 1] * sm.tab.window[2 * n + 1]};
  }
  // m = m1 + 5 m2 ; j = q + 5 r.   step 1: DFT5 over m2 for each m1  -> u[m1][q] stored at v[m1 + 5 q]
#pragma unroll
  for (int m1 = 0; m1 < 5; ++m1) dft5(v[m1], v[m1 + 5], v[m1 + 10], v[m1 + 15], v[m1 + 20]);
  //   twiddle u[m1][q] *= w25^(m1 q) = w200^(8 m1 q)
#pragma unroll
  for (int m1 = 1; m1 < 5; ++m1)
#pragma unroll
    for (int q = 1; q < 5; ++q) v[m1 + 5 * q] = cmul(v[m1 + 5 * q], sm.tab.tw200[8 * m1 * q]);
  //   step 2: DFT5 over m1 for each q -> out[q + 5 r] lands at v[r + 5 q]
#pragma unroll
  for (int q = 0; q < 5; ++q) dft5(v[5 * q], v[5 * q + 1], v[5 * q + 2], v[5 * q + 3], v[5 * q + 4]);
  // 200-point twiddle w200^(t j), j = q + 5 r, and store to slot (j*9 + t)
#pragma unroll
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int j = q + 5 * r;
      Cpx<T> o = v[r + 5 * q];
      if (t != 0 && j != 0) o = cmul(o, sm.tab.tw200[t * j]);     // t*j <= 7*24 = 168 < 200
      sm.y[fl * kYFrame + j * 9 + t] = o;
    }
}

// ---- phase B: radix-8 across t for each (frame, j); Z[j + 25 a] written back in place at slot (j*9 + a) -------------
template <typename T, int F>
__host__ __device__ __forceinline__ void logmel_phase_b(LogmelSmem<T, F>& sm, int task) {
  if (task >= F * 25) return;
  const int fl = task / 25, j = task - fl * 25;
  Cpx<T>* p = &sm.y[fl * kYFrame + j * 9];
  Cpx<T> y[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) y[t] = p[t];
  dft8(y);
#pragma unroll
  for (int a = 0; a < 8; ++a) p[a] = y[a];
}

template <typename T, int F>
__host__ __device__ __forceinline__ Cpx<T> logmel_z(const LogmelSmem<T, F>& sm, int fl, int k) {   // Z[k], k in [0,200)
  const int a = k / 25, j = k - a * 25;
  return sm.y[fl * kYFrame + j * 9 + a];
}

// ---- phase C1: split pass + power, bins k and 200-k together (k = 1..100). Bins 0 and 200 carry zero mel weight ---------
template <typename T, int F>
__host__ __device__ __forceinline__ void logmel_phase_c1(LogmelSmem<T, F>& sm, int fl, int k) {   // frame fl, bin pair k in [1, 100]
  const Cpx<T> zk = logmel_z(sm, fl, k), zc = logmel_z(sm, fl, 200 - k == 200 ? 0 : 200 - k);
  const Cpx<T> e = {(T)0.5 * (zk.re + zc.re), (T)0.5 * (zk.im - zc.im)};       // (Z[k] + conj Z[200-k]) / 2
  const Cpx<T> d = {(T)0.5 * (zk.re - zc.re), (T)0.5 * (zk.im + zc.im)};       // (Z[k] - conj Z[200-k]) / 2
  const Cpx<T> wo = cmul(mul_neg_i(d), sm.tab.tw400[k]);                        // w400^k * (-i d)
  const Cpx<T> xp = cadd(e, wo), xm = csub(e, wo);                             // X[k], conj X[200-k]
  T* P = &sm.region0[fl * kPFrame];
  P[k] = xp.re * xp.re + xp.im * xp.im;
  P[200 - k] = xm.re * xm.re + xm.im * xm.im;
}

// ---- phase C2: banded mel + log10 floor; returns the value (for the max) ----------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ T wb_log10(T x);
template <>
__host__ __device__ __forceinline__ float wb_log10<float>(float x) { return log10f(x); }
template <>
__host__ __device__ __forceinline__ double wb_log10<double>(double x) { return log10(x); }

// Mel band i for the NF frames fl0, fl0 + step, ... (those below F): the weight of a bin is loaded once for all of them, and each
// frame's sum runs over the band's bins in ascending order, exactly as the dense loop of lib.rs:60-69 does
template <typename T, int F, int NF>
__host__ __device__ __forceinline__ void logmel_phase_c2(const LogmelSmem<T, F>& sm, int i, int fl0, int step, T (&out)[NF]) {
  const int lo = sm.tab.mel_lo[i], cnt = sm.tab.mel_cnt[i], off = sm.tab.mel_off[i];
  const T* P = &sm.region0[fl0 * kPFrame + lo];
  T sum[NF];
#pragma unroll
  for (int n = 0; n < NF; ++n) sum[n] = (T)0;
  for (int c = 0; c < cnt; ++c) {
    const T w = sm.tab.melw[off + c];
#pragma unroll
    for (int n = 0; n < NF; ++n)
      if (fl0 + n * step < F) sum[n] += P[n * step * kPFrame + c] * w;
  }
  const T fl10 = (T)1e-10;
#pragma unroll
  for (int n = 0; n < NF; ++n) out[n] = wb_log10<T>(sum[n] > fl10 ? sum[n] : fl10);
}

}  // namespace wb
