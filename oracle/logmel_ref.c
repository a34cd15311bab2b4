/*
 * oracle/logmel_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU f64 restatement of the reference's Rust `stft` crate (/root/reference/stft/src/lib.rs), used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker and the
 * reported CPU baseline. The CUDA product never links or calls this file.
 *
 * What follows the reference, line by line:
 *   window      lib.rs:26      periodic Hann  w[i] = (1 - cos(2*pi*i/400)) / 2
 *   reflect     lib.rs:34-40   a[i] = a[400-i];  a[480200+i] = a[200 + 479998 - i],  i in [0,200)   (in place)
 *   frame loop  lib.rs:52      i = 0,160,... while i < len-400  -> exactly 3000 frames for len = 480400
 *   fft         lib.rs:42-47   window-multiply 400 samples, unnormalised forward real DFT -> 201 bins
 *   power       lib.rs:54      |X[k]|^2  (Complex::norm_sqr = re*re + im*im)
 *   mel         lib.rs:60-69   mel[i][j] = sum_{k=0..200} P[k][j] * (f64)MELS[i*201+k], k ascending, dense
 *   log         lib.rs:76      log10(max(x, 1e-10))
 *   max         lib.rs:82-88   global max over all 80x3000 values of this chunk
 *   scale       lib.rs:96      (max(x, gmax - 8) + 4) / 4
 *   layout      lib.rs:116-121 out[i*3000 + j], i = mel row, j = frame
 *
 * PARITY PIN STATUS. The FFT arithmetic of the reference lives in crates.io `realfft 3.0.1` -> `rustfft 6.0.1`
 * (stft/Cargo.toml:11, Cargo.lock), which is NOT vendored in /root/reference and cannot be built here (no
 * cargo/rustc). It is an exact unnormalised forward DFT in f64; this file restates the published algorithm
 * realfft uses for even lengths (pack to a half-length complex FFT, mixed-radix Cooley-Tukey, split post-pass).
 * Any two f64 DFTs agree to ~1e-15 relative, so results are pinned to ~1e-13 on the final output, not bit-for-bit:
 * "parity unpinned at the rustfft boundary". Pins that DO exist and are tested (tests/test_oracle_logmel.py):
 * m80.npy sha256; zeros -> -1.5 everywhere; agreement with an independent numpy O(N^2)-free rfft formulation and
 * with torch.stft in f64 (<= 1e-12); the naive O(N^2) DFT in this file (logmel_ref_naive) vs the FFT path.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC). No reference sources are copied or compiled.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mel80_table.h"

#define N_FFT 400
#define HALF 200
#define HOP 160
#define N_BINS 201
#define N_FRAMES 3000
#define N_MELS 80
#define N_SAMPLES 480000
#define N_PADDED (N_SAMPLES + 400)

typedef struct { double re, im; } cpx;

static double g_window[N_FFT];
static cpx g_tw200[HALF];      /* exp(-2 pi i k / 200) */
static cpx g_tw400[N_BINS];    /* exp(-2 pi i k / 400) */
static double g_mels[N_MELS * N_BINS];
static int g_init = 0;

static void init_tables(void) {
  if (g_init) return;
  const double PI = 3.14159265358979323846264338327950288;
  for (int i = 0; i < N_FFT; ++i) {
    double a = ((double)i * 2.0 * PI) / 400.0;            /* lib.rs:26 — same operation order */
    g_window[i] = (1.0 - cos(a)) / 2.0;
  }
  for (int k = 0; k < HALF; ++k) {
    g_tw200[k].re = cos(2.0 * PI * k / 200.0);
    g_tw200[k].im = -sin(2.0 * PI * k / 200.0);
  }
  for (int k = 0; k < N_BINS; ++k) {
    g_tw400[k].re = cos(2.0 * PI * k / 400.0);
    g_tw400[k].im = -sin(2.0 * PI * k / 400.0);
  }
  for (int i = 0; i < N_MELS * N_BINS; ++i) {
    float f;
    uint32_t b = MEL80_BITS[i];
    memcpy(&f, &b, 4);
    g_mels[i] = (double)f;                                  /* lib.rs:65 `as f64` */
  }
  g_init = 1;
}

/* lib.rs:34-40 */
void logmel_ref_reflect(double *audio) {
  for (int i = 0; i < 200; ++i) {
    audio[i] = audio[400 - i];
    int j = 16000 * 30 + i + 200;
    audio[j] = audio[200 + (16000 * 30 - 2) - i];
  }
}

/* Stockham autosort mixed-radix DIF step (radix r in {2,5}); n = current sub-length, s = stride. */
static void stockham_step(int n, int s, int r, const cpx *x, cpx *y) {
  const int m = n / r;
  const int tw_step = HALF / n;                             /* twiddle index scale: w_n^p = w_200^(p*200/n) */
  for (int p = 0; p < m; ++p) {
    for (int q = 0; q < s; ++q) {
      cpx a[5];
      for (int k = 0; k < r; ++k) a[k] = x[q + s * (p + k * m)];
      for (int j = 0; j < r; ++j) {
        cpx acc = a[0];
        for (int k = 1; k < r; ++k) {
          const cpx w = g_tw200[((j * k) % r) * (HALF / r)];     /* w_r^(jk) */
          acc.re += a[k].re * w.re - a[k].im * w.im;
          acc.im += a[k].re * w.im + a[k].im * w.re;
        }
        const cpx t = g_tw200[(p * j * tw_step) % HALF];         /* w_n^(pj) */
        cpx o;
        o.re = acc.re * t.re - acc.im * t.im;
        o.im = acc.re * t.im + acc.im * t.re;
        y[q + s * (r * p + j)] = o;
      }
    }
  }
}

/* Unnormalised forward real DFT of 400 samples -> 201 complex bins (even-length trick as in realfft). */
static void rfft400(const double *x, cpx *out) {
  cpx a[HALF], b[HALF];
  for (int n = 0; n < HALF; ++n) { a[n].re = x[2 * n]; a[n].im = x[2 * n + 1]; }
  static const int radices[5] = {5, 5, 2, 2, 2};
  int n = HALF, s = 1;
  cpx *src = a, *dst = b;
  for (int st = 0; st < 5; ++st) {
    stockham_step(n, s, radices[st], src, dst);
    n /= radices[st];
    s *= radices[st];
    cpx *t = src; src = dst; dst = t;
  }
  const cpx *Z = src;
  for (int k = 0; k <= HALF; ++k) {
    const cpx zk = Z[k % HALF];
    const cpx zc = Z[(HALF - k) % HALF];                      /* conj applied below */
    const double er = 0.5 * (zk.re + zc.re), ei = 0.5 * (zk.im - zc.im);      /* even part  */
    const double dr = 0.5 * (zk.re - zc.re), di = 0.5 * (zk.im + zc.im);      /* (Z-conjZ')/2 */
    /* odd part = -i * d ; X = e + w400^k * odd */
    const double orr = di, oi = -dr;
    const cpx w = g_tw400[k];
    out[k].re = er + (orr * w.re - oi * w.im);
    out[k].im = ei + (orr * w.im + oi * w.re);
  }
}

static void dft400_naive(const double *x, cpx *out) {
  const double PI = 3.14159265358979323846264338327950288;
  for (int k = 0; k < N_BINS; ++k) {
    double re = 0.0, im = 0.0;
    for (int n = 0; n < N_FFT; ++n) {
      const int idx = (k * n) % N_FFT;
      const double ang = 2.0 * PI * idx / 400.0;
      re += x[n] * cos(ang);
      im -= x[n] * sin(ang);
    }
    out[k].re = re;
    out[k].im = im;
  }
}

/* lib.rs:49-102. `audio` is the 480400-sample reflected buffer. `out` is [80][3000]. */
static int spectrogram(const double *audio, double *out, int naive) {
  init_tables();
  double *power = (double *)malloc(sizeof(double) * N_BINS * N_FRAMES);   /* [201][3000], like Vec<Vec<f64>> */
  if (!power) return -1;
  int j = 0;
  for (int i = 0; i < N_PADDED - 400; i += HOP, ++j) {      /* lib.rs:52 */
    double in[N_FFT];
    cpx sp[N_BINS];
    for (int t = 0; t < N_FFT; ++t) in[t] = audio[i + t] * g_window[t];   /* lib.rs:43 */
    if (naive) dft400_naive(in, sp); else rfft400(in, sp);
    for (int k = 0; k < N_BINS; ++k) power[k * N_FRAMES + j] = sp[k].re * sp[k].re + sp[k].im * sp[k].im;
  }
  if (j != N_FRAMES) { free(power); return -2; }
  for (int i = 0; i < N_MELS; ++i) {                         /* lib.rs:60-69 — same loop nest and order */
    for (int jj = 0; jj < N_FRAMES; ++jj) {
      double sum = 0.0;
      for (int k = 0; k < N_BINS; ++k) sum += power[k * N_FRAMES + jj] * g_mels[i * N_BINS + k];
      out[i * N_FRAMES + jj] = sum;
    }
  }
  free(power);
  double gmax = -INFINITY;
  for (int i = 0; i < N_MELS * N_FRAMES; ++i) {              /* lib.rs:76, 82-88 */
    double v = out[i];
    v = log10(v > 1e-10 ? v : 1e-10);
    out[i] = v;
    if (v > gmax) gmax = v;
  }
  for (int i = 0; i < N_MELS * N_FRAMES; ++i) {              /* lib.rs:96 */
    double v = out[i];
    const double fl = gmax - 8.0;
    v = v > fl ? v : fl;
    out[i] = (v + 4.0) / 4.0;
  }
  return 0;
}

/* Same contract as the reference symbol `generate_spectrogram` (lib.rs:110-122; bridge.h:11):
 * audio[480400] in/out (pads overwritten by the reflection), output[240000] = [80][3000]. */
void logmel_ref_generate_spectrogram(double *audio, double *output) {
  logmel_ref_reflect(audio);
  spectrogram(audio, output, 0);
}

/* Same, with the O(N^2) DFT — used once in the tests to pin the FFT path. */
void logmel_ref_generate_spectrogram_naive(double *audio, double *output) {
  logmel_ref_reflect(audio);
  spectrogram(audio, output, 1);
}

/* Convenience for batched f32 callers (tests/bench): audio [B][480000] f32, out [B][80][3000] f64.
 * Follows the Swift wrapper stft.swift:8-19 (prepend/append 200 zeros) + ContentView.swift:57-60 (f32->f64). */
int logmel_ref_batch_f32(const float *audio, int B, double *out) {
  double *buf = (double *)malloc(sizeof(double) * N_PADDED);
  if (!buf) return -1;
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < 200; ++i) { buf[i] = 0.0; buf[N_SAMPLES + 200 + i] = 0.0; }
    for (int i = 0; i < N_SAMPLES; ++i) buf[200 + i] = (double)audio[(size_t)b * N_SAMPLES + i];
    logmel_ref_generate_spectrogram(buf, out + (size_t)b * N_MELS * N_FRAMES);
  }
  free(buf);
  return 0;
}
