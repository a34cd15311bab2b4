// Model instance: weight arena, workspaces, encoder forward, decoder step / greedy loop (CUDA-graph replayed), and the
// extern "C" entry points declared in include/whisper_b200.h.
//
// HBM layout (one handle = one GPU):
//   weight arena   one allocation, every tensor 256-byte aligned: fp16 matrices in [N][K] (K-major, the layout both the
//                  TMA-fed tcgen05 GEMM and the weight-streaming decoder GEMM read), q/k/v fused to [3d][d], conv weights
//                  permuted to [d_out][tap][c_in]; fp32 biases / LayerNorm / positional tables
//   encoder ws     melT f16 [B][3002][80] | x1 f16 [B][3001][d] | x f32 [B*1500][d] | h f16 | qkv f16 [.][3d] | att f16 |
//                  mlp f16 [.][4d] | xa f16 [B*1500][d]
//   cross K/V      f16 [L][B][1500][d] each  (persistent across decode steps; 2*L*1500*d*2 bytes per chunk)
//   self  K/V      f16 [L][Mb][n_text_ctx][d] each
//   decode ws      tokens i32 [Mb][n_text_ctx+1] | x f32 [Mb][d] | q f32 | attention split partials | mlp f16 [Mb][4d] |
//                  logits f32 [Mb][V] | DecodeState
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <stdio.h>
#include <zlib.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/whisper_b200.h"
#include "logmel.cuh"
#include "ops.cuh"

namespace wb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

size_t logmel_tables_bytes_f32() { return sizeof(LogmelTables<float>); }
size_t logmel_tables_bytes_f64() { return sizeof(LogmelTables<double>); }

struct LayerW {
  float *ln1_g, *ln1_b;
  __half* wqkv;
  float* bqkv;
  __half* wo;
  float* bo;
  float *ln2_g, *ln2_b;
  __half* w1;
  float* b1;
  __half* w2;
  float* b2;
  // decoder only
  float *lnc_g, *lnc_b;
  __half* wq_c;
  float* bq_c;
  __half* wk_c;
  __half* wv_c;
  float* bv_c;
  __half* wo_c;
  float* bo_c;
};

enum WKind { WK_F32 = 0, WK_F16 = 1, WK_CONV = 2 };
struct WEntry {
  int kind;
  void* dst;
  size_t numel;
  int conv_out, conv_in;   // WK_CONV: source [out][in][3] -> dest [out][3][in]
  bool set;
  float rnd_scale, rnd_offset;
};

struct Arena {
  unsigned char* base = nullptr;
  size_t size = 0, off = 0;
  bool measure = true;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = measure ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace wb

using namespace wb;

struct wb_handle {
  wb_dims dims;
  int max_batch, max_beams, device;
  cudaStream_t stream;
  bool own_stream;
  int64_t launches;
  bool weights_ready;
  int enc_batch;   // chunks whose features are resident (0 = none)

  // weights
  Arena arena;
  std::unordered_map<std::string, WEntry> wmap;
  __half *conv1_w, *conv2_w, *tok_emb;
  float *conv1_b, *conv2_b, *enc_pos, *dec_pos, *lnpost_g, *lnpost_b, *lnf_g, *lnf_b;
  std::vector<LayerW> enc, dec;

  // log-mel
  LogmelTables<float>* tab32;
  float* audio_dev;
  float* logspec;
  void* gmax;
  float* mel32;

  // encoder workspace
  Arena ws;
  __half *melT, *x1, *h16, *qkv16, *att16, *mlp16, *xa16;
  float *xenc, *xa32;
  std::vector<__half*> crossK, crossV;
  GemmContext* gemm;

  // decoder workspace
  int Mb_max;
  std::vector<__half*> selfK, selfV, selfK_alt, selfV_alt;   // *_alt: ping-pong target of the beam re-indexing
  float* top_lp;
  int32_t *top_idx, *beam_src;
  int32_t* tokens;
  int tokens_ld;
  float *xdec, *q32, *logits, *sum_logprob, *part_logits, *part_extra;
  int4* ts_state;   // timestamp-rule state per sequence (FinishDesc)
  int32_t* chosen;  // temperature sampling: drawn tokens / their log-probabilities / no-speech probabilities [Mb]
  float *chosen_lp, *no_speech_prob;
  // long-form (wb_transcribe_long): the whole recording and its unnormalised log-mel, allocated on demand
  float *long_audio, *long_logspec;
  size_t long_audio_cap, long_windows_cap;
  __half *dmlp16, *a16, *dln16;   // dln16: LayerNorm rows of the current use (wider models)
  int32_t* done;
  unsigned char* mask;
  int n_logit_ctas;
  DecodeState* state;
  unsigned long long* trace;   // WB_TRACE=1

  // greedy decode runs as up to 4 sub-batches of the chunk batch, each with its own stream, DecodeState and pair of CUDA
  // graphs: the latency-bound GEMM chain of one sub-batch overlaps the bandwidth-bound attention of another
  static constexpr int kMaxSub = 4;
  cudaStream_t sub_stream[kMaxSub];
  cudaEvent_t sub_ev[kMaxSub], fork_ev;
  cudaGraphExec_t g_step[kMaxSub], g_sample[kMaxSub];
  cudaGraphExec_t g_sample_n[kMaxSub];   // several sampling steps in one graph (fewer graph launches, dependent launch across steps)
  cudaGraphExec_t g_beam[2];             // beam search: the scored step (+ top-k), one graph per K/V buffer orientation
  const __half* beam_kv[2];
  int64_t nodes_beam;
  std::string beam_key;
  cudaGraphExec_t g_pair[3];             // interleaved pair of sub-batches: prompt step, one sampling step, sample_n sampling steps
  int64_t nodes_pair[3];
  cudaEvent_t pair_ev[2];
  int sample_n;                          // steps per g_sample_n graph
  int64_t nodes_step, nodes_sample;
  std::string graph_key;

  // host audio arrives in slabs on its own stream, so that all but the first slab's copy runs under the encoder (wb_transcribe)
  static constexpr int kMaxSlabs = 4;
  cudaStream_t copy_stream;
  cudaEvent_t copy_ev[kMaxSlabs], copy_fence;

  // host staging + timing
  int32_t* h_done;
  // beam search staging in pinned host memory: top-k values / indices, and - double-buffered by step parity, so that no second
  // synchronisation per step is needed - the source-beam map, the newest token column and the timestamp-rule states
  float* hb_top_lp;
  int32_t *hb_top_idx, *hb_src, *hb_col;
  int4* hb_ts;
  cudaEvent_t ev[4];
  float timings[4];
};

namespace wb {

static void add_entry(wb_handle* h, const std::string& name, int kind, void* dst, size_t numel, float rs, float ro,
                      int co = 0, int ci = 0) {
  h->wmap[name] = WEntry{kind, dst, numel, co, ci, false, rs, ro};
}

// Lays out every weight in the arena; called twice (measure, then assign).
static void layout_weights(wb_handle* h) {
  const wb_dims& D = h->dims;
  Arena& A = h->arena;
  const size_t d = D.n_audio_state, dt = D.n_text_state;
  const bool reg = !A.measure;
  h->conv1_w = A.take<__half>(d * 240);
  h->conv1_b = A.take<float>(d);
  h->conv2_w = A.take<__half>(d * 3 * d);
  h->conv2_b = A.take<float>(d);
  h->enc_pos = A.take<float>((size_t)D.n_audio_ctx * d);
  if (reg) {
    add_entry(h, "encoder.conv1.weight", WK_CONV, h->conv1_w, d * 240, 1.0f / sqrtf(240.f), 0, (int)d, 80);
    add_entry(h, "encoder.conv1.bias", WK_F32, h->conv1_b, d, 0.02f, 0);
    add_entry(h, "encoder.conv2.weight", WK_CONV, h->conv2_w, d * 3 * d, 1.0f / sqrtf(3.f * d), 0, (int)d, (int)d);
    add_entry(h, "encoder.conv2.bias", WK_F32, h->conv2_b, d, 0.02f, 0);
    add_entry(h, "encoder.positional_embedding", WK_F32, h->enc_pos, (size_t)D.n_audio_ctx * d, 0.3f, 0);
  }
  auto block = [&](const std::string& pre, LayerW& L, size_t n, bool cross) {
    const float ws = 0.7f / sqrtf((float)n), qs = 2.0f / sqrtf((float)n);
    L.ln1_g = A.take<float>(n), L.ln1_b = A.take<float>(n);
    L.wqkv = A.take<__half>(3 * n * n), L.bqkv = A.take<float>(3 * n);
    L.wo = A.take<__half>(n * n), L.bo = A.take<float>(n);
    L.ln2_g = A.take<float>(n), L.ln2_b = A.take<float>(n);
    L.w1 = A.take<__half>(4 * n * n), L.b1 = A.take<float>(4 * n);
    L.w2 = A.take<__half>(4 * n * n), L.b2 = A.take<float>(n);
    if (cross) {
      L.lnc_g = A.take<float>(n), L.lnc_b = A.take<float>(n);
      L.wq_c = A.take<__half>(n * n), L.bq_c = A.take<float>(n);
      L.wk_c = A.take<__half>(n * n);
      L.wv_c = A.take<__half>(n * n), L.bv_c = A.take<float>(n);
      L.wo_c = A.take<__half>(n * n), L.bo_c = A.take<float>(n);
    }
    if (!reg) return;
    add_entry(h, pre + ".attn_ln.weight", WK_F32, L.ln1_g, n, 0.1f, 1.0f);
    add_entry(h, pre + ".attn_ln.bias", WK_F32, L.ln1_b, n, 0.05f, 0);
    add_entry(h, pre + ".attn.query.weight", WK_F16, L.wqkv, n * n, qs, 0);
    add_entry(h, pre + ".attn.key.weight", WK_F16, L.wqkv + n * n, n * n, qs, 0);
    add_entry(h, pre + ".attn.value.weight", WK_F16, L.wqkv + 2 * n * n, n * n, ws, 0);
    add_entry(h, pre + ".attn.query.bias", WK_F32, L.bqkv, n, 0.02f, 0);
    add_entry(h, pre + ".attn.value.bias", WK_F32, L.bqkv + 2 * n, n, 0.02f, 0);
    add_entry(h, pre + ".attn.out.weight", WK_F16, L.wo, n * n, ws, 0);
    add_entry(h, pre + ".attn.out.bias", WK_F32, L.bo, n, 0.02f, 0);
    add_entry(h, pre + ".mlp_ln.weight", WK_F32, L.ln2_g, n, 0.1f, 1.0f);
    add_entry(h, pre + ".mlp_ln.bias", WK_F32, L.ln2_b, n, 0.05f, 0);
    add_entry(h, pre + ".mlp.0.weight", WK_F16, L.w1, 4 * n * n, ws, 0);
    add_entry(h, pre + ".mlp.0.bias", WK_F32, L.b1, 4 * n, 0.02f, 0);
    add_entry(h, pre + ".mlp.2.weight", WK_F16, L.w2, 4 * n * n, 0.7f / sqrtf(4.f * n), 0);
    add_entry(h, pre + ".mlp.2.bias", WK_F32, L.b2, n, 0.02f, 0);
    if (cross) {
      add_entry(h, pre + ".cross_attn_ln.weight", WK_F32, L.lnc_g, n, 0.1f, 1.0f);
      add_entry(h, pre + ".cross_attn_ln.bias", WK_F32, L.lnc_b, n, 0.05f, 0);
      add_entry(h, pre + ".cross_attn.query.weight", WK_F16, L.wq_c, n * n, qs, 0);
      add_entry(h, pre + ".cross_attn.query.bias", WK_F32, L.bq_c, n, 0.02f, 0);
      add_entry(h, pre + ".cross_attn.key.weight", WK_F16, L.wk_c, n * n, qs, 0);
      add_entry(h, pre + ".cross_attn.value.weight", WK_F16, L.wv_c, n * n, ws, 0);
      add_entry(h, pre + ".cross_attn.value.bias", WK_F32, L.bv_c, n, 0.02f, 0);
      add_entry(h, pre + ".cross_attn.out.weight", WK_F16, L.wo_c, n * n, ws, 0);
      add_entry(h, pre + ".cross_attn.out.bias", WK_F32, L.bo_c, n, 0.02f, 0);
    }
  };
  h->enc.resize(D.n_audio_layer);
  for (int i = 0; i < D.n_audio_layer; ++i) block("encoder.blocks." + std::to_string(i), h->enc[i], d, false);
  h->lnpost_g = A.take<float>(d), h->lnpost_b = A.take<float>(d);
  h->tok_emb = A.take<__half>((size_t)D.n_vocab * dt);
  h->dec_pos = A.take<float>((size_t)D.n_text_ctx * dt);
  h->dec.resize(D.n_text_layer);
  for (int i = 0; i < D.n_text_layer; ++i) block("decoder.blocks." + std::to_string(i), h->dec[i], dt, true);
  h->lnf_g = A.take<float>(dt), h->lnf_b = A.take<float>(dt);
  if (reg) {
    add_entry(h, "encoder.ln_post.weight", WK_F32, h->lnpost_g, d, 0.1f, 1.0f);
    add_entry(h, "encoder.ln_post.bias", WK_F32, h->lnpost_b, d, 0.05f, 0);
    add_entry(h, "decoder.token_embedding.weight", WK_F16, h->tok_emb, (size_t)D.n_vocab * dt, 0.05f, 0);
    add_entry(h, "decoder.positional_embedding", WK_F32, h->dec_pos, (size_t)D.n_text_ctx * dt, 0.1f, 0);
    add_entry(h, "decoder.ln.weight", WK_F32, h->lnf_g, dt, 0.1f, 1.0f);
    add_entry(h, "decoder.ln.bias", WK_F32, h->lnf_b, dt, 0.05f, 0);
  }
}

static void layout_workspace(wb_handle* h) {
  const wb_dims& D = h->dims;
  Arena& A = h->ws;
  const size_t B = h->max_batch, d = D.n_audio_state, dt = D.n_text_state, T = D.n_audio_ctx, Mb = h->Mb_max;
  h->tab32 = A.take<LogmelTables<float>>(1);
  h->audio_dev = A.take<float>(B * WB_N_SAMPLES);
  h->logspec = A.take<float>(B * WB_N_MELS * WB_N_FRAMES);
  h->gmax = A.take<unsigned long long>(B);
  h->mel32 = A.take<float>(B * WB_N_MELS * WB_N_FRAMES);
  h->melT = A.take<__half>(B * (WB_N_FRAMES + 2) * WB_N_MELS);
  h->x1 = A.take<__half>(B * (2 * T + 1) * d);
  h->xenc = A.take<float>(B * T * d);
  h->h16 = A.take<__half>(B * T * d);
  h->qkv16 = A.take<__half>(B * T * 3 * d);
  h->att16 = A.take<__half>(B * T * d);
  h->mlp16 = A.take<__half>(B * T * 4 * d);
  h->xa16 = A.take<__half>(B * T * d);
  h->xa32 = A.take<float>(B * T * d);
  h->crossK.resize(D.n_text_layer), h->crossV.resize(D.n_text_layer);
  for (int l = 0; l < D.n_text_layer; ++l) {
    h->crossK[l] = A.take<__half>(B * T * dt);
    h->crossV[l] = A.take<__half>(B * T * dt);
  }
  h->selfK.resize(D.n_text_layer), h->selfV.resize(D.n_text_layer);
  for (int l = 0; l < D.n_text_layer; ++l) {
    h->selfK[l] = A.take<__half>(Mb * D.n_text_ctx * dt);
    h->selfV[l] = A.take<__half>(Mb * D.n_text_ctx * dt);
  }
  if (h->max_beams > 1) {
    h->selfK_alt.resize(D.n_text_layer), h->selfV_alt.resize(D.n_text_layer);
    for (int l = 0; l < D.n_text_layer; ++l) {
      h->selfK_alt[l] = A.take<__half>(Mb * D.n_text_ctx * dt);
      h->selfV_alt[l] = A.take<__half>(Mb * D.n_text_ctx * dt);
    }
  }
  h->top_lp = A.take<float>(Mb * 8);
  h->top_idx = A.take<int32_t>(Mb * 8);
  h->beam_src = A.take<int32_t>(Mb);
  h->tokens_ld = D.n_text_ctx + 8;
  h->tokens = A.take<int32_t>(Mb * h->tokens_ld);
  h->xdec = A.take<float>(Mb * dt);
  h->q32 = A.take<float>(Mb * dt);
  h->dmlp16 = A.take<__half>(Mb * 4 * dt);
  h->dln16 = A.take<__half>(Mb * dt);
  h->logits = A.take<float>(Mb * (size_t)D.n_vocab);
  h->sum_logprob = A.take<float>(Mb);
  h->done = A.take<int32_t>(Mb);
  h->a16 = A.take<__half>(Mb * dt);
  h->mask = A.take<unsigned char>((size_t)D.n_vocab);
  h->n_logit_ctas = skinny_logits_ctas(D.n_vocab);
  h->part_logits = A.take<float>(Mb * (size_t)h->n_logit_ctas * 4);
  h->part_extra = A.take<float>(Mb * (size_t)4);
  h->ts_state = A.take<int4>(Mb);
  h->chosen = A.take<int32_t>(Mb);
  h->chosen_lp = A.take<float>(Mb);
  h->no_speech_prob = A.take<float>(Mb);
  h->state = A.take<DecodeState>(wb_handle::kMaxSub);
  h->trace = getenv("WB_TRACE") ? A.take<unsigned long long>(65536 * 8) : nullptr;
}

static int check_dims(const wb_dims& D) {
  if (D.n_mels != 80 || D.n_audio_ctx != 1500) return -1;
  if (D.n_audio_state <= 0 || D.n_audio_state % 128 || D.n_audio_state > 1280) return -1;
  if (D.n_text_state != D.n_audio_state) return -1;
  if (D.n_audio_head * 64 != D.n_audio_state || D.n_text_head * 64 != D.n_text_state) return -1;
  if (D.n_audio_layer < 1 || D.n_text_layer < 1 || D.n_vocab < 1000 || D.n_text_ctx < 8 || D.n_text_ctx > 448) return -1;
  return 0;
}

#define WB_TRY(expr)           \
  do {                         \
    const int _rc = (expr);    \
    if (_rc != 0) return _rc;  \
  } while (0)

// ---- encoder -------------------------------------------------------------------------------------------------------------
static int plain_gemm(wb_handle* h, const __half* a, int M, const __half* w, int N, int K, const float* bias, int gelu,
                      const float* res, __half* c16, float* c32) {
  GemmDesc g{};
  g.a = a, g.a_row_stride = K, g.a_batch_stride = (long long)M * K, g.rows = M, g.n_batch = 1;
  g.w = w, g.N = N, g.K = K, g.bias = bias, g.gelu = gelu, g.res_mode = res ? 1 : 0, g.res = res;
  g.c16 = c16, g.c32 = c32, g.ldc = N, g.c_batch_rows = 0, g.c_row_off = 0;
  return launch_gemm(h->gemm, g, h->stream, &h->launches);
}

// melT (already filled) -> xa16 / xa32 and the cross-attention K/V of every decoder layer, for chunks [b0, b0 + n) of the
// batch. Every buffer is chunk-major, so a slab is the same computation on offset pointers (results do not depend on the
// slab split: every output element is one K-ordered reduction).
static int encoder_forward(wb_handle* h, int b0, int n) {
  const wb_dims& D = h->dims;
  const int d = D.n_audio_state, T = D.n_audio_ctx, M = n * T;
  const size_t r0 = (size_t)b0 * T;   // first row of the slab in the [B*T][.] matrices
  cudaStream_t st = h->stream;
  __half* melT = h->melT + (size_t)b0 * (WB_N_FRAMES + 2) * WB_N_MELS;
  __half* x1 = h->x1 + (size_t)b0 * (2 * T + 1) * d;
  float* xenc = h->xenc + r0 * d;
  __half *h16 = h->h16 + r0 * d, *qkv16 = h->qkv16 + r0 * 3 * d, *att16 = h->att16 + r0 * d, *mlp16 = h->mlp16 + r0 * 4 * d;
  {   // conv1 + GELU: im2col row t = 240 contiguous halves at melT[b][t][0]
    GemmDesc g{};
    g.a = melT, g.a_row_stride = WB_N_MELS, g.a_batch_stride = (long long)(WB_N_FRAMES + 2) * WB_N_MELS;
    g.rows = WB_N_FRAMES, g.n_batch = n, g.w = h->conv1_w, g.N = d, g.K = 3 * WB_N_MELS, g.bias = h->conv1_b, g.gelu = 1;
    g.c16 = x1, g.ldc = d, g.c_batch_rows = 2 * T + 1, g.c_row_off = 1;
    WB_TRY(launch_gemm(h->gemm, g, st, &h->launches));
  }
  {   // conv2 (stride 2) + GELU + sinusoidal positions: row t = 3d contiguous halves at x1[b][2t][0]
    GemmDesc g{};
    g.a = x1, g.a_row_stride = 2 * d, g.a_batch_stride = (long long)(2 * T + 1) * d;
    g.rows = T, g.n_batch = n, g.w = h->conv2_w, g.N = d, g.K = 3 * d, g.bias = h->conv2_b, g.gelu = 1;
    g.res_mode = 2, g.res = h->enc_pos, g.c32 = xenc, g.ldc = d, g.c_batch_rows = T, g.c_row_off = 0;
    WB_TRY(launch_gemm(h->gemm, g, st, &h->launches));
  }
  for (int l = 0; l < D.n_audio_layer; ++l) {
    const LayerW& L = h->enc[l];
    WB_TRY(launch_layernorm(xenc, L.ln1_g, L.ln1_b, M, d, h16, nullptr, st, &h->launches));
    WB_TRY(plain_gemm(h, h16, M, L.wqkv, 3 * d, d, L.bqkv, 0, nullptr, qkv16, nullptr));
    WB_TRY(launch_encoder_attention(h->gemm, qkv16, n, T, D.n_audio_head, att16, st, &h->launches));
    WB_TRY(plain_gemm(h, att16, M, L.wo, d, d, L.bo, 0, xenc, nullptr, xenc));
    WB_TRY(launch_layernorm(xenc, L.ln2_g, L.ln2_b, M, d, h16, nullptr, st, &h->launches));
    WB_TRY(plain_gemm(h, h16, M, L.w1, 4 * d, d, L.b1, 1, nullptr, mlp16, nullptr));
    WB_TRY(plain_gemm(h, mlp16, M, L.w2, d, 4 * d, L.b2, 0, xenc, nullptr, xenc));
  }
  WB_TRY(launch_layernorm(xenc, h->lnpost_g, h->lnpost_b, M, d, h->xa16 + r0 * d, h->xa32 + r0 * d, st, &h->launches));
  return 0;
}
static int encoder_forward(wb_handle* h, int B) { return encoder_forward(h, 0, B); }

static int cross_kv(wb_handle* h, int b0, int n) {
  const wb_dims& D = h->dims;
  const int d = D.n_text_state, M = n * D.n_audio_ctx;
  const size_t off = (size_t)b0 * D.n_audio_ctx * d;
  for (int l = 0; l < D.n_text_layer; ++l) {
    const LayerW& L = h->dec[l];
    WB_TRY(plain_gemm(h, h->xa16 + off, M, L.wk_c, d, d, nullptr, 0, nullptr, h->crossK[l] + off, nullptr));
    WB_TRY(plain_gemm(h, h->xa16 + off, M, L.wv_c, d, d, L.bv_c, 0, nullptr, h->crossV[l] + off, nullptr));
  }
  if (b0 == 0 || h->enc_batch == b0) h->enc_batch = b0 + n;   // slabs arrive in order
  return 0;
}
static int cross_kv(wb_handle* h, int B) {
  h->enc_batch = 0;
  return cross_kv(h, 0, B);
}

// ---- decoder ---------------------------------------------------------------------------------------------------------------
struct StepOpts {
  int Mb, beams;
  int store_logits;   // write the fp32 logits [Mb][V] (teacher-forced / language-ID paths)
  int sample;         // run the logits GEMM with filters + partial argmax and sample in the finish kernel
  int n_initial, eot;
  int no_finish;      // beam search: the host picks the next tokens between the logits and the finish kernel
  int b0;             // first sequence of this sub-batch (Mb sequences starting at b0)
  int sub;            // sub-batch index: selects the stream and the DecodeState
  int timestamps;     // upstream ApplyTimestampRules among the logit filters (sampling steps only)
  int ts_begin, ts_last_allowed;
  int use_chosen;     // the finish kernel takes the tokens sample_rows_kernel drew (temperature > 0)
  // interleaved sub-batch pair (decode_steps_pair): layers [l_begin, l_end) of the step, the logits / finish tail only with
  // `tail`; the KV-cache kernel waits for xwait (the other sub-batch's previous KV-cache kernel) and records xrec behind itself
  int partial, l_begin, l_end, tail;
  cudaEvent_t xwait, xrec;
};

static cudaStream_t step_stream(wb_handle* h, const StepOpts& o) { return o.sub > 0 ? h->sub_stream[o.sub] : h->stream; }
static DecodeState* step_state(wb_handle* h, const StepOpts& o) { return h->state + o.sub; }

static int step_finish(wb_handle* h, const StepOpts& o, int sample) {
  const wb_dims& D = h->dims;
  const size_t b0 = o.b0;
  FinishDesc f{};
  f.Mb = o.Mb, f.V = D.n_vocab, f.d = D.n_text_state, f.n_ctx = D.n_text_ctx, f.sample = sample;
  f.n_part = logits_groups(o.Mb, D.n_vocab, D.n_text_state);
  f.part_logits = h->part_logits + b0 * h->n_logit_ctas * 4, f.eot = o.eot;
  f.tokens = h->tokens + b0 * h->tokens_ld, f.tokens_ld = h->tokens_ld;
  f.sum_logprob = h->sum_logprob + b0, f.done = h->done + b0, f.tok_emb = h->tok_emb, f.pos_emb = h->dec_pos;
  f.x = h->xdec + b0 * D.n_text_state, f.state = step_state(h, o);
  if (sample && o.timestamps) {
    f.ts_state = h->ts_state + b0, f.part_extra = h->part_extra + b0 * 4;
    f.ts_begin = o.ts_begin, f.ts_group0 = (o.ts_begin + 127) / 128, f.n_initial = o.n_initial;
  }
  if (sample && o.use_chosen) f.chosen = h->chosen + b0, f.chosen_logprob = h->chosen_lp + b0;
  return launch_step_finish(f, step_stream(h, o), &h->launches);
}

// One decoder step for the token at position state->cur_len of every sequence (its embedding is already in xdec).
static int decode_step(wb_handle* h, const StepOpts& o) {
  const wb_dims& D = h->dims;
  const int d = D.n_text_state, H = D.n_text_head, Mb = o.Mb;
  const size_t b0 = o.b0;
  cudaStream_t st = step_stream(h, o);
  DecodeState* state = step_state(h, o);
  float* xdec = h->xdec + b0 * d;
  float* q32 = h->q32 + b0 * d;
  __half* a16 = h->a16 + b0 * d;
  __half* dmlp16 = h->dmlp16 + b0 * 4 * d;
  __half* dln16 = h->dln16 + b0 * d;
  // wider models (the skinny-GEMM chain): LayerNorm as its own one-warp-per-row kernel in front of the GEMMs that consume it
  // (WB_LN_ROWS=0: fused into every CTA of the GEMM, as for the block-kernel path's fallbacks)
  static int ln_rows_env = -1;
  if (ln_rows_env < 0) {
    const char* e = getenv("WB_LN_ROWS");
    ln_rows_env = e ? atoi(e) : 1;
  }
  const bool ln_rows = ln_rows_env != 0 && d >= 768;
  const size_t self_off = b0 * (size_t)D.n_text_ctx * d;
  const size_t cross_off = (b0 / o.beams) * (size_t)D.n_audio_ctx * d;
  const int l_begin = o.partial ? o.l_begin : 0, l_end = o.partial ? o.l_end : D.n_text_layer;
  const bool tail = o.partial ? o.tail != 0 : true;
  for (int l = l_begin; l < l_end; ++l) {
    const LayerW& L = h->dec[l];
    AttnDecodeDesc a{};
    a.Mb = Mb, a.d = d, a.n_head = H, a.q = q32, a.k = h->selfK[l] + self_off, a.v = h->selfV[l] + self_off;
    a.n_ctx = D.n_text_ctx, a.n_rows_fixed = 0, a.kv_share = 1, a.state = state, a.out16 = a16, a.tmaps = h->gemm;
    SkinnyDesc so{};
    so.Mb = Mb, so.state = state, so.N = d, so.K = d, so.w = L.wo, so.bias = L.bo, so.in_mode = SKINNY_IN_F16, so.in = a16;
    so.out_mode = SKINNY_OUT_RESID, so.out = xdec;
    if (self_block_supported(H, d)) {
      // n_head <= 8: LN + QKV + cache append + attention + output projection + residual in one cluster kernel
      SelfBlockDesc sb{};
      sb.Mb = Mb, sb.d = d, sb.n_head = H, sb.n_ctx = D.n_text_ctx, sb.x = xdec, sb.ln_g = L.ln1_g, sb.ln_b = L.ln1_b;
      sb.wqkv = L.wqkv, sb.bqkv = L.bqkv, sb.wo = L.wo, sb.bo = L.bo;
      sb.kcache = h->selfK[l] + self_off, sb.vcache = h->selfV[l] + self_off, sb.state = state;
      WB_TRY(launch_self_block(sb, st, &h->launches));
    } else {
      // LN + fused QKV (K/V appended to the cache by the epilogue), attention, output projection
      SkinnyDesc s{};
      s.Mb = Mb, s.state = state;
      s.N = 3 * d, s.K = d, s.w = L.wqkv, s.bias = L.bqkv, s.in_mode = SKINNY_IN_LN, s.in = xdec, s.ln_g = L.ln1_g,
      s.ln_b = L.ln1_b, s.out_mode = SKINNY_OUT_QKV, s.q32 = q32, s.kcache = h->selfK[l] + self_off, s.vcache = h->selfV[l] + self_off,
      s.n_ctx = D.n_text_ctx;
      if (ln_rows) {   // the rows are normalised once, not by each of the GEMM's 144 .. 240 CTAs
        WB_TRY(launch_ln_rows(xdec, L.ln1_g, L.ln1_b, Mb, d, dln16, state, st, &h->launches));
        s.in_mode = SKINNY_IN_F16, s.in = dln16, s.ln_g = s.ln_b = nullptr;
      }
      WB_TRY(launch_skinny_gemm(s, st, &h->launches));
      WB_TRY(launch_attn_decode(a, st, &h->launches));
      WB_TRY(launch_skinny_gemm(so, st, &h->launches));
    }
    // cross attention: LayerNorm + query projection fused into the attention kernel (one kernel less per layer)
    AttnDecodeDesc c = a;
    c.k = h->crossK[l] + cross_off, c.v = h->crossV[l] + cross_off, c.n_ctx = D.n_audio_ctx, c.n_rows_fixed = D.n_audio_ctx;
    c.kv_share = o.beams;
    c.q = nullptr, c.x = xdec, c.ln_g = L.lnc_g, c.ln_b = L.lnc_b, c.wq = L.wq_c, c.bq = L.bq_c;
    c.pdl_late_ok = post_block_supported(H, d) ? 1 : 0;
    {
      // d >= 1024: the fused projection keeps 160 registers of weight rows per thread (one CTA per SM) and re-reads 128 / 164 KB
      // of Wq per (sequence, head); a separate LayerNorm + projection kernel in front of a two-CTAs-per-SM attention kernel
      // instead (WB_XA_FUSE_Q = 0 / 1 forces either)
      static int fuse_env = -2;
      if (fuse_env == -2) {
        const char* e = getenv("WB_XA_FUSE_Q");
        fuse_env = e ? atoi(e) : -1;
      }
      const bool fuse = fuse_env >= 0 ? fuse_env != 0 : d < 1024;
      if (!fuse) {
        SkinnyDesc sq{};
        sq.Mb = Mb, sq.state = state, sq.N = d, sq.K = d, sq.w = L.wq_c, sq.bias = L.bq_c, sq.in_mode = SKINNY_IN_LN, sq.in = xdec;
        sq.ln_g = L.lnc_g, sq.ln_b = L.lnc_b, sq.out_mode = SKINNY_OUT_F32, sq.out = q32;
        if (ln_rows) {
          WB_TRY(launch_ln_rows(xdec, L.lnc_g, L.lnc_b, Mb, d, dln16, state, st, &h->launches));
          sq.in_mode = SKINNY_IN_F16, sq.in = dln16, sq.ln_g = sq.ln_b = nullptr;
        }
        WB_TRY(launch_skinny_gemm(sq, st, &h->launches));
        c.q = q32, c.x = nullptr, c.wq = nullptr, c.bq = nullptr, c.ln_g = nullptr, c.ln_b = nullptr;
      }
    }
    if (o.xwait) WB_CUDA_OK(cudaStreamWaitEvent(st, o.xwait, 0));
    WB_TRY(launch_attn_decode(c, st, &h->launches));
    if (o.xrec) WB_CUDA_OK(cudaEventRecord(o.xrec, st));
    if (post_block_supported(H, d)) {
      // d = 384 / 512: output projection + residual + LayerNorm + MLP + residual in one cluster kernel
      PostBlockDesc pb{};
      pb.Mb = Mb, pb.d = d, pb.n_head = H, pb.x = xdec, pb.a16 = a16, pb.wo = L.wo_c, pb.bo = L.bo_c, pb.ln_g = L.ln2_g, pb.ln_b = L.ln2_b;
      pb.w1 = L.w1, pb.b1 = L.b1, pb.w2 = L.w2, pb.b2 = L.b2, pb.state = state;
      WB_TRY(launch_post_block(pb, st, &h->launches));
      continue;
    }
    SkinnyDesc sc = so;
    sc.w = L.wo_c, sc.bias = L.bo_c;
    WB_TRY(launch_skinny_gemm(sc, st, &h->launches));
    // MLP
    SkinnyDesc m1{};
    m1.Mb = Mb, m1.state = state, m1.N = 4 * d, m1.K = d, m1.w = L.w1, m1.bias = L.b1, m1.gelu = 1;
    m1.in_mode = SKINNY_IN_LN, m1.in = xdec, m1.ln_g = L.ln2_g, m1.ln_b = L.ln2_b, m1.out_mode = SKINNY_OUT_F16, m1.out = dmlp16;
    if (ln_rows) {
      WB_TRY(launch_ln_rows(xdec, L.ln2_g, L.ln2_b, Mb, d, dln16, state, st, &h->launches));
      m1.in_mode = SKINNY_IN_F16, m1.in = dln16, m1.ln_g = m1.ln_b = nullptr;
    }
    WB_TRY(launch_skinny_gemm(m1, st, &h->launches));
    SkinnyDesc m2{};
    m2.Mb = Mb, m2.state = state, m2.N = d, m2.K = 4 * d, m2.w = L.w2, m2.bias = L.b2;
    m2.in_mode = SKINNY_IN_F16, m2.in = dmlp16, m2.out_mode = SKINNY_OUT_RESID, m2.out = xdec;
    WB_TRY(launch_skinny_gemm(m2, st, &h->launches));
  }
  if (!tail) return 0;
  if (o.store_logits || o.sample) {
    // final LN + tied-embedding logits; filters and per-CTA (max, argmax, sum-exp) fused into the epilogue
    SkinnyDesc lg{};
    lg.Mb = Mb, lg.state = state, lg.N = D.n_vocab, lg.K = d, lg.w = h->tok_emb, lg.in_mode = SKINNY_IN_LN;
    lg.in = xdec, lg.ln_g = h->lnf_g, lg.ln_b = h->lnf_b, lg.out_mode = SKINNY_OUT_LOGITS;
    lg.out = o.store_logits ? h->logits + b0 * (size_t)D.n_vocab : nullptr, lg.mask = o.sample ? h->mask : nullptr, lg.n_initial = o.n_initial;
    lg.part_logits = h->part_logits + b0 * h->n_logit_ctas * 4;
    lg.tmaps = h->gemm;
    lg.ts_begin = 0x7fffffff, lg.ts_last_allowed = 0x7fffffff, lg.eot = o.eot;
    if (o.sample && o.timestamps) {
      lg.ts_state = h->ts_state + b0, lg.part_extra = h->part_extra + b0 * 4;
      lg.ts_begin = o.ts_begin, lg.ts_last_allowed = o.ts_last_allowed;
    }
    WB_TRY(launch_skinny_gemm(lg, st, &h->launches));
  }
  if (o.no_finish) return 0;
  return step_finish(h, o, o.sample);
}

// cur_len = -1, then embed the token at position 0 (which advances cur_len to 0)
static int reset_decode_state(wb_handle* h, const StepOpts& o) {
  static const DecodeState k_init_untraced{-1, 0, 0, 0, nullptr};
  DecodeState init = k_init_untraced;
  if (h->trace) init.trace = h->trace + (size_t)o.sub * 16384 * 8;   // a quarter of the trace buffer per sub-batch
  static thread_local DecodeState staged[wb_handle::kMaxSub];
  staged[o.sub] = init;                    // stays valid until the async copy has run
  WB_CUDA_OK(cudaMemcpyAsync(step_state(h, o), &staged[o.sub], sizeof(init), cudaMemcpyHostToDevice, step_stream(h, o)));
  return step_finish(h, o, 0);
}

static void destroy_graphs(wb_handle* h) {
  for (int i = 0; i < wb_handle::kMaxSub; ++i) {
    if (h->g_step[i]) cudaGraphExecDestroy(h->g_step[i]);
    if (h->g_sample[i]) cudaGraphExecDestroy(h->g_sample[i]);
    if (h->g_sample_n[i]) cudaGraphExecDestroy(h->g_sample_n[i]);
    h->g_step[i] = h->g_sample[i] = h->g_sample_n[i] = nullptr;
  }
  for (int i = 0; i < 3; ++i) {
    if (h->g_pair[i]) cudaGraphExecDestroy(h->g_pair[i]);
    h->g_pair[i] = nullptr;
  }
  for (int i = 0; i < 2; ++i) {
    if (h->g_beam[i]) cudaGraphExecDestroy(h->g_beam[i]);
    h->g_beam[i] = nullptr, h->beam_kv[i] = nullptr;
  }
  h->beam_key.clear();
  h->graph_key.clear();
}

static int capture(wb_handle* h, const StepOpts& o, cudaGraphExec_t* out, int64_t* nodes, int n_steps = 1) {
  cudaGraph_t g;
  const int64_t before = h->launches;
  cudaStream_t st = step_stream(h, o);
  WB_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  for (int k = 0; k < n_steps && rc == 0; ++k) rc = decode_step(h, o);
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  *nodes = h->launches - before;
  h->launches = before;
  if (rc) return rc;
  WB_CUDA_OK(e);
  WB_CUDA_OK(cudaGraphInstantiate(out, g, 0));
  WB_CUDA_OK(cudaGraphDestroy(g));
  return 0;
}

// Two sub-batches as ONE interleaved schedule (the three-kernel path): the bandwidth-bound KV-cache kernels of the two halves
// are ordered against each other (A.l, B.l, A.l+1, ...: each streams alone, at full HBM rate), and between two of them a
// half's latency-bound block kernels run under the other half's stream. Left to themselves, two sub-batch streams fall into
// lockstep (both in the same kind of kernel at the same time: tools/trace_sub.py) and gain nothing.
static int decode_steps_pair(wb_handle* h, const StepOpts& oa, const StepOpts& ob, int n_steps) {
  const int L = h->dims.n_text_layer;
  bool first = true;
  for (int k = 0; k < n_steps; ++k) {
    for (int l = 0; l < L; ++l) {
      StepOpts a = oa, b = ob;
      a.partial = b.partial = 1, a.l_begin = b.l_begin = l, a.l_end = b.l_end = l + 1, a.tail = b.tail = 0;
      a.xwait = first ? nullptr : h->pair_ev[1], a.xrec = h->pair_ev[0];
      b.xwait = h->pair_ev[0], b.xrec = h->pair_ev[1];
      WB_TRY(decode_step(h, a));
      WB_TRY(decode_step(h, b));
      first = false;
    }
    StepOpts a = oa, b = ob;
    a.partial = b.partial = 1, a.l_begin = b.l_begin = a.l_end = b.l_end = L, a.tail = b.tail = 1;
    WB_TRY(decode_step(h, a));
    WB_TRY(decode_step(h, b));
  }
  return 0;
}

static int capture_pair(wb_handle* h, const StepOpts& oa, const StepOpts& ob, cudaGraphExec_t* out, int64_t* nodes, int n_steps) {
  cudaGraph_t g;
  const int64_t before = h->launches;
  cudaStream_t sa = step_stream(h, oa), sb = step_stream(h, ob);
  WB_CUDA_OK(cudaStreamBeginCapture(sa, cudaStreamCaptureModeThreadLocal));
  cudaError_t e = cudaEventRecord(h->fork_ev, sa);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(sb, h->fork_ev, 0);
  int rc = e == cudaSuccess ? decode_steps_pair(h, oa, ob, n_steps) : -1;
  if (cudaEventRecord(h->sub_ev[1], sb) != cudaSuccess || cudaStreamWaitEvent(sa, h->sub_ev[1], 0) != cudaSuccess) rc = rc ? rc : -1;
  const cudaError_t ee = cudaStreamEndCapture(sa, &g);
  *nodes = h->launches - before;
  h->launches = before;
  if (rc) {
    if (e != cudaSuccess) set_error("pair capture: %s", cudaGetErrorString(e));
    return rc < 0 ? WB_ERR_CUDA : rc;
  }
  WB_CUDA_OK(ee);
  WB_CUDA_OK(cudaGraphInstantiate(out, g, 0));
  WB_CUDA_OK(cudaGraphDestroy(g));
  return 0;
}

}  // namespace wb

// ======================================================================================================================
//                                                   extern "C"
// ======================================================================================================================
extern "C" {

const char* wb_last_error(void) { return wb::get_error(); }
int wb_version(void) { return 100; }

static int select_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return WB_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (%d visible)", device, n);
    return WB_ERR_ARG;
  }
  WB_CUDA_OK(cudaSetDevice(device));
  return 0;
}

// Releases whatever a (possibly half-constructed) handle owns; every member is null / zero until it is created.
static void free_handle(wb_handle* h) {
  if (!h) return;
  destroy_graphs(h);
  if (h->gemm) gemm_context_destroy(h->gemm);
  for (int i = 0; i < 4; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 1; i < wb_handle::kMaxSub; ++i)
    if (h->sub_stream[i]) cudaStreamDestroy(h->sub_stream[i]);
  for (int i = 0; i < wb_handle::kMaxSub; ++i)
    if (h->sub_ev[i]) cudaEventDestroy(h->sub_ev[i]);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  for (int i = 0; i < 2; ++i)
    if (h->pair_ev[i]) cudaEventDestroy(h->pair_ev[i]);
  for (int i = 0; i < wb_handle::kMaxSlabs; ++i)
    if (h->copy_ev[i]) cudaEventDestroy(h->copy_ev[i]);
  if (h->copy_fence) cudaEventDestroy(h->copy_fence);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h_done) cudaFreeHost(h->h_done);
  if (h->hb_top_lp) cudaFreeHost(h->hb_top_lp);
  if (h->arena.base) cudaFree(h->arena.base);
  if (h->ws.base) cudaFree(h->ws.base);
  if (h->long_audio) cudaFree(h->long_audio);
  if (h->long_logspec) cudaFree(h->long_logspec);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int create_impl(wb_handle* h, void* stream) {
  h->own_stream = stream == nullptr;
  if (h->own_stream)
    WB_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  else
    h->stream = reinterpret_cast<cudaStream_t>(stream);
  // weights
  h->arena.measure = true, h->arena.off = 0;
  layout_weights(h);
  h->arena.size = (h->arena.off + 255) & ~(size_t)255;
  if (cudaMalloc(&h->arena.base, h->arena.size) != cudaSuccess) {
    h->arena.base = nullptr;
    set_error("wb_create: cudaMalloc(%zu) for weights failed", h->arena.size);
    return WB_ERR_NOMEM;
  }
  WB_CUDA_OK(cudaMemsetAsync(h->arena.base, 0, h->arena.size, h->stream));
  h->arena.measure = false, h->arena.off = 0;
  layout_weights(h);
  // workspace
  h->ws.measure = true, h->ws.off = 0;
  layout_workspace(h);
  h->ws.size = (h->ws.off + 255) & ~(size_t)255;
  if (cudaMalloc(&h->ws.base, h->ws.size) != cudaSuccess) {
    h->ws.base = nullptr;
    set_error("wb_create: cudaMalloc(%zu) for workspace failed", h->ws.size);
    return WB_ERR_NOMEM;
  }
  WB_CUDA_OK(cudaMemsetAsync(h->ws.base, 0, h->ws.size, h->stream));
  h->ws.measure = false, h->ws.off = 0;
  layout_workspace(h);
  {
    std::unique_ptr<LogmelTables<float>> t(new LogmelTables<float>());
    build_logmel_tables<float>(*t);
    WB_CUDA_OK(cudaMemcpyAsync(h->tab32, t.get(), sizeof(*t), cudaMemcpyHostToDevice, h->stream));
    WB_CUDA_OK(cudaStreamSynchronize(h->stream));   // `t` is pageable host memory
  }
  h->gemm = gemm_context_create();
  {
    const char* e = getenv("WB_L2_POLICY");
    WB_TRY(decoder_set_l2_mode(e ? atoi(e) : 1));
  }
  WB_CUDA_OK(cudaMallocHost(&h->h_done, sizeof(int32_t) * h->Mb_max));
  if (h->max_beams > 1) {   // one pinned block: [Mb*8] f32 | [Mb*8] i32 | 2 x [Mb] i32 | 2 x [Mb] i32 | 2 x [Mb] int4
    const size_t Mb = (size_t)h->Mb_max;
    unsigned char* blk = nullptr;
    WB_CUDA_OK(cudaMallocHost(&blk, Mb * 8 * 4 + Mb * 8 * 4 + 2 * Mb * 4 + 2 * Mb * 4 + 2 * Mb * 16));
    h->hb_top_lp = reinterpret_cast<float*>(blk);
    h->hb_top_idx = reinterpret_cast<int32_t*>(blk + Mb * 32);
    h->hb_src = h->hb_top_idx + Mb * 8;
    h->hb_col = h->hb_src + 2 * Mb;
    h->hb_ts = reinterpret_cast<int4*>(h->hb_col + 2 * Mb);
  }
  for (int i = 0; i < 4; ++i) WB_CUDA_OK(cudaEventCreate(&h->ev[i]));
  for (int i = 1; i < wb_handle::kMaxSub; ++i) WB_CUDA_OK(cudaStreamCreateWithFlags(&h->sub_stream[i], cudaStreamNonBlocking));
  for (int i = 0; i < wb_handle::kMaxSub; ++i) WB_CUDA_OK(cudaEventCreateWithFlags(&h->sub_ev[i], cudaEventDisableTiming));
  WB_CUDA_OK(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) WB_CUDA_OK(cudaEventCreateWithFlags(&h->pair_ev[i], cudaEventDisableTiming));
  WB_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < wb_handle::kMaxSlabs; ++i) WB_CUDA_OK(cudaEventCreateWithFlags(&h->copy_ev[i], cudaEventDisableTiming));
  WB_CUDA_OK(cudaEventCreateWithFlags(&h->copy_fence, cudaEventDisableTiming));
  // everything above (the zero fill the bqkv key-bias slice, the melT pad rows and x1 row 0 rely on, the tables) is complete
  // on the device before the handle is handed out, whatever stream later work uses
  WB_CUDA_OK(cudaDeviceSynchronize());
  return WB_OK;
}

int wb_create(const wb_dims* dims, int32_t max_batch, int32_t max_beams, int32_t device, void* stream, wb_handle** out) {
  if (!dims || !out || max_batch < 1 || max_beams < 1 || max_batch * max_beams > kMaxSequences) {
    set_error("wb_create: bad argument (max_batch * max_beams must be in [1, %d])", kMaxSequences);
    return WB_ERR_ARG;
  }
  if (check_dims(*dims)) {
    set_error("wb_create: unsupported model dimensions");
    return WB_ERR_ARG;
  }
  if (!self_block_supported(dims->n_text_head, dims->n_text_state) && max_batch * max_beams > kMaxSequencesWide) {
    set_error("wb_create: max_batch * max_beams must be in [1, %d] for this model width (%d up to d = 512)", kMaxSequencesWide, kMaxSequences);
    return WB_ERR_ARG;
  }
  WB_TRY(select_device(device));
  wb_handle* h = new wb_handle();   // value-initialised: every pointer null, every counter zero
  h->dims = *dims, h->max_batch = max_batch, h->max_beams = max_beams, h->device = device;
  h->Mb_max = max_batch * max_beams;
  h->sample_n = 1;
  const int rc = create_impl(h, stream);
  if (rc != WB_OK) {
    free_handle(h);
    return rc;
  }
  *out = h;
  return WB_OK;
}

int wb_destroy(wb_handle* h) {
  if (!h) return WB_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_handle(h);
  return WB_OK;
}

int wb_get_dims(const wb_handle* h, wb_dims* out) {
  if (!h || !out) return WB_ERR_ARG;
  *out = h->dims;
  return WB_OK;
}

int wb_set_weight(wb_handle* h, const char* name, const float* data, size_t numel) {
  if (!h || !name || !data) return WB_ERR_ARG;
  auto it = h->wmap.find(name);
  if (it == h->wmap.end()) {
    set_error("wb_set_weight: unknown tensor '%s'", name);
    return WB_ERR_ARG;
  }
  WEntry& e = it->second;
  if (numel != e.numel) {
    set_error("wb_set_weight: '%s' expects %zu elements, got %zu", name, e.numel, numel);
    return WB_ERR_ARG;
  }
  WB_CUDA_OK(cudaSetDevice(h->device));
  if (e.kind == WK_F32) {
    WB_CUDA_OK(cudaMemcpy(e.dst, data, numel * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    std::vector<__half> tmp(numel);
    if (e.kind == WK_F16) {
      for (size_t i = 0; i < numel; ++i) tmp[i] = __float2half_rn(data[i]);
    } else {   // [out][in][3] -> [out][3][in]
      const size_t ci = e.conv_in;
      for (size_t o = 0; o < (size_t)e.conv_out; ++o)
        for (size_t c = 0; c < ci; ++c)
          for (size_t t = 0; t < 3; ++t) tmp[(o * 3 + t) * ci + c] = __float2half_rn(data[(o * ci + c) * 3 + t]);
    }
    WB_CUDA_OK(cudaMemcpy(e.dst, tmp.data(), numel * sizeof(__half), cudaMemcpyHostToDevice));
  }
  e.set = true;
  return WB_OK;
}

int wb_weights_commit(wb_handle* h) {
  if (!h) return WB_ERR_ARG;
  for (auto& kv : h->wmap)
    if (!kv.second.set) {
      set_error("wb_weights_commit: tensor '%s' was never set", kv.first.c_str());
      return WB_ERR_STATE;
    }
  // wb_set_weight copies from pageable memory on the legacy stream, which does not order against the handle's
  // non-blocking stream: make every copy land before anything can read the arena
  WB_CUDA_OK(cudaSetDevice(h->device));
  WB_CUDA_OK(cudaDeviceSynchronize());
  h->weights_ready = true;
  return WB_OK;
}

int wb_init_random_weights(wb_handle* h, uint64_t seed) {
  if (!h) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  uint64_t i = 0;
  std::vector<std::string> names;
  for (auto& kv : h->wmap) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  for (auto& nm : names) {
    WEntry& e = h->wmap[nm];
    WB_TRY(launch_fill_random(e.dst, e.numel, e.kind != WK_F32, e.rnd_scale, e.rnd_offset, seed * 1000003ull + (++i), h->stream,
                              &h->launches));
    e.set = true;
  }
  // encoder positions: upstream sinusoids (same table as the oracle's sinusoids())
  const int T = h->dims.n_audio_ctx, d = h->dims.n_audio_state;
  std::vector<float> pos((size_t)T * d);
  const double inc = log(10000.0) / (d / 2 - 1);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < d / 2; ++c) {
      const float inv = expf((float)(-inc * c));
      pos[(size_t)t * d + c] = sinf((float)t * inv);
      pos[(size_t)t * d + d / 2 + c] = cosf((float)t * inv);
    }
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  WB_CUDA_OK(cudaMemcpy(h->enc_pos, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice));
  WB_CUDA_OK(cudaDeviceSynchronize());
  h->weights_ready = true;
  return WB_OK;
}

int wb_weight_arena(wb_handle* h, void** device_ptr, size_t* bytes) {
  if (!h || !device_ptr || !bytes) return WB_ERR_ARG;
  *device_ptr = h->arena.base;
  *bytes = h->arena.size;
  return WB_OK;
}

int wb_weights_mark_loaded(wb_handle* h) {
  if (!h) return WB_ERR_ARG;
  for (auto& kv : h->wmap) kv.second.set = true;
  WB_CUDA_OK(cudaSetDevice(h->device));
  WB_CUDA_OK(cudaDeviceSynchronize());   // the external fill (e.g. an NCCL broadcast on another stream) has landed
  h->weights_ready = true;
  return WB_OK;
}

int wb_get_weight(wb_handle* h, const char* name, float* data, size_t numel) {
  if (!h || !name || !data) return WB_ERR_ARG;
  auto it = h->wmap.find(name);
  if (it == h->wmap.end()) {
    set_error("wb_get_weight: unknown tensor '%s'", name);
    return WB_ERR_ARG;
  }
  const WEntry& e = it->second;
  if (numel != e.numel) {
    set_error("wb_get_weight: '%s' holds %zu elements, caller asked for %zu", name, e.numel, numel);
    return WB_ERR_ARG;
  }
  WB_CUDA_OK(cudaSetDevice(h->device));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (e.kind == WK_F32) {
    WB_CUDA_OK(cudaMemcpy(data, e.dst, numel * sizeof(float), cudaMemcpyDeviceToHost));
    return WB_OK;
  }
  std::vector<__half> tmp(numel);
  WB_CUDA_OK(cudaMemcpy(tmp.data(), e.dst, numel * sizeof(__half), cudaMemcpyDeviceToHost));
  if (e.kind == WK_F16) {
    for (size_t i = 0; i < numel; ++i) data[i] = __half2float(tmp[i]);
  } else {   // device [out][3][in] -> upstream [out][in][3]
    const size_t ci = e.conv_in;
    for (size_t o = 0; o < (size_t)e.conv_out; ++o)
      for (size_t c = 0; c < ci; ++c)
        for (size_t t = 0; t < 3; ++t) data[(o * ci + c) * 3 + t] = __half2float(tmp[(o * 3 + t) * ci + c]);
  }
  return WB_OK;
}

int wb_weight_count(const wb_handle* h) { return h ? (int)h->wmap.size() : -1; }

int wb_weight_info(const wb_handle* h, int32_t index, char* name_out, size_t name_cap, size_t* numel) {
  if (!h || index < 0 || (size_t)index >= h->wmap.size() || !name_out || name_cap == 0) return WB_ERR_ARG;
  std::vector<std::string> names;
  names.reserve(h->wmap.size());
  for (auto& kv : h->wmap) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  const std::string& nm = names[index];
  if (nm.size() + 1 > name_cap) {
    set_error("wb_weight_info: name buffer too small (%zu needed)", nm.size() + 1);
    return WB_ERR_ARG;
  }
  memcpy(name_out, nm.c_str(), nm.size() + 1);
  if (numel) *numel = h->wmap.at(nm).numel;
  return WB_OK;
}

int wb_weights_checksum(wb_handle* h, uint64_t* out) {
  if (!h || !out) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(h->gmax);   // scratch word of the workspace
  WB_TRY(launch_checksum64(h->arena.base, h->arena.size, acc, h->stream, &h->launches));
  unsigned long long v = 0;
  WB_CUDA_OK(cudaMemcpyAsync(&v, acc, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  *out = (uint64_t)v;
  return WB_OK;
}

// ---- log-mel -------------------------------------------------------------------------------------------------------------
static int check_batch(wb_handle* h, int32_t B) {
  if (!h) return WB_ERR_ARG;
  if (B < 1 || B > h->max_batch) {
    set_error("batch %d out of range [1, %d]", B, h->max_batch);
    return WB_ERR_ARG;
  }
  WB_CUDA_OK(cudaSetDevice(h->device));
  return 0;
}

int wb_logmel_dev(wb_handle* h, const float* audio_dev, int32_t B, float* out_dev) {
  WB_TRY(check_batch(h, B));
  if (!audio_dev || !out_dev) return WB_ERR_ARG;
  return launch_logmel<float>(audio_dev, WB_N_SAMPLES, 0, B, h->tab32, h->logspec, h->gmax, out_dev, nullptr, h->stream,
                              &h->launches);
}

int wb_logmel(wb_handle* h, const float* audio, int32_t B, float* out) {
  WB_TRY(check_batch(h, B));
  if (!audio || !out) return WB_ERR_ARG;
  WB_CUDA_OK(cudaMemcpyAsync(h->audio_dev, audio, (size_t)B * WB_N_SAMPLES * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  WB_TRY(wb_logmel_dev(h, h->audio_dev, B, h->mel32));
  WB_CUDA_OK(cudaMemcpyAsync(out, h->mel32, (size_t)B * WB_N_MELS * WB_N_FRAMES * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return WB_OK;
}

// ---- encoder -------------------------------------------------------------------------------------------------------------
static int need_weights(wb_handle* h) {
  if (!h->weights_ready) {
    set_error("weights not loaded (wb_set_weight + wb_weights_commit, or wb_init_random_weights)");
    return WB_ERR_STATE;
  }
  return 0;
}

int wb_encode_dev(wb_handle* h, const float* audio_dev, int32_t B) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_weights(h));
  if (!audio_dev) return WB_ERR_ARG;
  WB_CUDA_OK(cudaEventRecord(h->ev[0], h->stream));
  WB_TRY(launch_logmel<float>(audio_dev, WB_N_SAMPLES, 0, B, h->tab32, h->logspec, h->gmax, nullptr, h->melT, h->stream, &h->launches));
  WB_CUDA_OK(cudaEventRecord(h->ev[1], h->stream));
  WB_TRY(encoder_forward(h, B));
  WB_TRY(cross_kv(h, B));
  WB_CUDA_OK(cudaEventRecord(h->ev[2], h->stream));
  return WB_OK;
}

static int fetch_xa(wb_handle* h, int B, float* xa_out) {
  if (xa_out)
    WB_CUDA_OK(cudaMemcpyAsync(xa_out, h->xa32, (size_t)B * h->dims.n_audio_ctx * h->dims.n_audio_state * sizeof(float),
                               cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int wb_encode(wb_handle* h, const float* audio, int32_t B, float* xa_out) {
  WB_TRY(check_batch(h, B));
  if (!audio) return WB_ERR_ARG;
  WB_CUDA_OK(cudaMemcpyAsync(h->audio_dev, audio, (size_t)B * WB_N_SAMPLES * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  WB_TRY(wb_encode_dev(h, h->audio_dev, B));
  return fetch_xa(h, B, xa_out);
}

int wb_encode_mel(wb_handle* h, const float* mel, int32_t B, float* xa_out) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_weights(h));
  if (!mel) return WB_ERR_ARG;
  WB_CUDA_OK(cudaMemcpyAsync(h->mel32, mel, (size_t)B * WB_N_MELS * WB_N_FRAMES * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  WB_TRY(launch_mel_transpose(h->mel32, B, h->melT, h->stream, &h->launches));
  WB_TRY(encoder_forward(h, B));
  WB_TRY(cross_kv(h, B));
  return fetch_xa(h, B, xa_out);
}

int wb_set_audio_features(wb_handle* h, const float* xa, int32_t B) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_weights(h));
  if (!xa) return WB_ERR_ARG;
  const size_t n = (size_t)B * h->dims.n_audio_ctx * h->dims.n_audio_state;
  WB_CUDA_OK(cudaMemcpyAsync(h->xa32, xa, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  WB_TRY(launch_f32_to_f16(h->xa32, h->xa16, n, h->stream, &h->launches));
  WB_TRY(cross_kv(h, B));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return WB_OK;
}

// ---- decoder -------------------------------------------------------------------------------------------------------------
static int need_features(wb_handle* h, int B) {
  WB_TRY(need_weights(h));
  if (h->enc_batch < B) {
    set_error("audio features for %d chunk(s) are not resident (run wb_encode first; resident: %d)", B, h->enc_batch);
    return WB_ERR_STATE;
  }
  return 0;
}

int wb_get_audio_features(wb_handle* h, float* xa_out, int32_t B) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_features(h, B));
  if (!xa_out) return WB_ERR_ARG;
  return fetch_xa(h, B, xa_out);
}

int wb_decoder_logits(wb_handle* h, const int32_t* tokens, int32_t B, int32_t t, float* logits) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_features(h, B));
  if (!tokens || !logits || t < 1 || t > h->dims.n_text_ctx) {
    set_error("wb_decoder_logits: bad argument");
    return WB_ERR_ARG;
  }
  const int V = h->dims.n_vocab;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < t; ++i)
      if (tokens[b * t + i] < 0 || tokens[b * t + i] >= V) {
        set_error("wb_decoder_logits: token %d out of vocabulary", tokens[b * t + i]);
        return WB_ERR_ARG;
      }
  WB_CUDA_OK(cudaMemcpy2DAsync(h->tokens, h->tokens_ld * sizeof(int32_t), tokens, t * sizeof(int32_t), t * sizeof(int32_t), B,
                               cudaMemcpyHostToDevice, h->stream));
  StepOpts o{};
  o.Mb = B, o.beams = 1, o.store_logits = 1, o.sample = 0;
  WB_TRY(reset_decode_state(h, o));
  for (int i = 0; i < t; ++i) {
    WB_TRY(decode_step(h, o));
    WB_CUDA_OK(cudaMemcpy2DAsync(logits + (size_t)i * V, (size_t)t * V * sizeof(float), h->logits, (size_t)V * sizeof(float),
                                 (size_t)V * sizeof(float), B, cudaMemcpyDeviceToHost, h->stream));
  }
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return WB_OK;
}

int wb_decoder_logits_f32tok(wb_handle* h, const float* tokens, int32_t B, int32_t t, float* logits) {
  if (!tokens || B < 1 || t < 1) return WB_ERR_ARG;
  std::vector<int32_t> tk((size_t)B * t);
  for (size_t i = 0; i < tk.size(); ++i) tk[i] = (int32_t)lrintf(tokens[i]);   // Whisper.swift:34-35 passes 50258.0
  return wb_decoder_logits(h, tk.data(), B, t, logits);
}

int wb_detect_language(wb_handle* h, int32_t B, int32_t sot, int32_t lang0, int32_t* lang_idx) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_features(h, B));
  if (!lang_idx) return WB_ERR_ARG;
  if (sot <= 0) sot = 50258;      // Whisper.swift:35
  if (lang0 <= 0) lang0 = 50259;  // Whisper.swift:37
  const int V = h->dims.n_vocab;
  if (lang0 + 99 > V || sot >= V) {
    set_error("wb_detect_language: language tokens [%d,%d) outside the vocabulary (%d)", lang0, lang0 + 99, V);
    return WB_ERR_ARG;
  }
  std::vector<int32_t> tk(B, sot);
  WB_CUDA_OK(cudaMemcpy2DAsync(h->tokens, h->tokens_ld * sizeof(int32_t), tk.data(), sizeof(int32_t), sizeof(int32_t), B,
                               cudaMemcpyHostToDevice, h->stream));
  StepOpts o{};
  o.Mb = B, o.beams = 1, o.store_logits = 1, o.sample = 0;
  WB_TRY(reset_decode_state(h, o));
  WB_TRY(decode_step(h, o));
  std::vector<float> conf((size_t)B * 99);
  WB_CUDA_OK(cudaMemcpy2DAsync(conf.data(), 99 * sizeof(float), h->logits + lang0, (size_t)V * sizeof(float), 99 * sizeof(float), B,
                               cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int b = 0; b < B; ++b) {
    // Swift `max { $0.element < $1.element }` (Whisper.swift:38) replaces its running result only on a strict increase
    // (`if areInIncreasingOrder(result, e) { result = e }`): the FIRST maximal element wins ties, NaN never replaces
    int best = 0;
    for (int i = 1; i < 99; ++i)
      if (conf[b * 99 + best] < conf[b * 99 + i]) best = i;
    lang_idx[b] = best;
  }
  return WB_OK;
}

// Option checks shared by the greedy and the beam path (pointer / count consistency, lengths, vocabulary ranges)
static int validate_decode_opts(const wb_handle* h, int32_t B, const wb_decode_opts* opts) {
  const wb_dims& D = h->dims;
  const int n_init = opts->n_initial, total = n_init + opts->sample_len;
  if (n_init < 1 || opts->sample_len < 1 || total > D.n_text_ctx || !opts->initial_tokens) {
    set_error("wb_decode: bad options (n_initial=%d sample_len=%d: need n_initial >= 1, sample_len >= 1, n_initial + sample_len <= n_text_ctx=%d)",
              n_init, opts->sample_len, D.n_text_ctx);
    return WB_ERR_ARG;
  }
  if (opts->n_suppress < 0 || opts->n_suppress_begin < 0 || (opts->n_suppress > 0 && !opts->suppress) ||
      (opts->n_suppress_begin > 0 && !opts->suppress_begin)) {
    set_error("wb_decode: suppress / suppress_begin count without a list");
    return WB_ERR_ARG;
  }
  if (opts->eot < 0 || opts->eot >= D.n_vocab) {
    set_error("wb_decode: eot %d outside the vocabulary (%d)", opts->eot, D.n_vocab);
    return WB_ERR_ARG;
  }
  for (int i = 0; i < n_init; ++i)
    if (opts->initial_tokens[i] < 0 || opts->initial_tokens[i] >= D.n_vocab) {
      set_error("wb_decode: initial token %d outside the vocabulary (%d)", opts->initial_tokens[i], D.n_vocab);
      return WB_ERR_ARG;
    }
  if (opts->beam_size < 0 || opts->beam_size > 7) {
    set_error("wb_decode: beam_size %d out of range [0, 7]", opts->beam_size);
    return WB_ERR_ARG;
  }
  if (!(opts->temperature >= 0.f) || opts->best_of < 0 || (opts->temperature > 0.f && opts->beam_size > 1)) {
    set_error("wb_decode: temperature must be >= 0 (and 0 with beam search), best_of >= 0");
    return WB_ERR_ARG;
  }
  if (opts->no_speech_prob &&
      (opts->sot_index < 0 || opts->sot_index >= n_init || opts->no_speech < 0 || opts->no_speech >= D.n_vocab || opts->beam_size > 1)) {
    set_error("wb_decode: no_speech_prob needs 0 <= sot_index < n_initial, a valid no_speech token and greedy / sampled decoding");
    return WB_ERR_ARG;
  }
  if (opts->timestamps) {
    if (opts->timestamp_begin <= opts->eot || opts->timestamp_begin >= D.n_vocab || opts->no_timestamps < 0 ||
        opts->no_timestamps >= D.n_vocab || B > 64) {
      set_error("wb_decode: timestamp rules need eot < timestamp_begin < n_vocab, a valid no_timestamps token and <= 64 sequences");
      return WB_ERR_ARG;
    }
  }
  return 0;
}

// Upstream DecodingTask._main_loop `no_speech_probs`: softmax of the unfiltered logits at the <|startoftranscript|> position.
// The prompt step that consumes that token runs with the logits GEMM; if it is also the first sampling step (the whole
// prompt is [sot]), its state is rewound afterwards (the step updates the residual stream in place) and the token embedded
// again, so that the sampling step runs from the same state with the logit filters.
static int no_speech_probe_step(wb_handle* h, const StepOpts& plain, int no_speech, int pos, bool advance) {
  const int V = h->dims.n_vocab;
  StepOpts o = plain;
  o.store_logits = 1, o.sample = 0, o.no_finish = advance ? 0 : 1;
  WB_TRY(decode_step(h, o));
  WB_TRY(launch_row_token_prob(h->logits + (size_t)plain.b0 * V, plain.Mb, V, no_speech, h->no_speech_prob + plain.b0, step_stream(h, plain),
                               &h->launches));
  if (!advance) {
    WB_TRY(launch_set_cur_len(step_state(h, plain), pos - 1, step_stream(h, plain), &h->launches));
    WB_TRY(step_finish(h, plain, 0));
  }
  return 0;
}

static int decode_greedy(wb_handle* h, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens, float* sum_logprob) {
  const wb_dims& D = h->dims;
  const int n_init = opts->n_initial, total = n_init + opts->sample_len;
  const bool ts_on = opts->timestamps != 0;
  cudaStream_t st = h->stream;
  // token rows: sot sequence, then eot padding
  std::vector<int32_t> rows((size_t)B * h->tokens_ld, opts->eot);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < n_init; ++i) rows[(size_t)b * h->tokens_ld + i] = opts->initial_tokens[i];
  // logit filters as a per-token mask: 1 = SuppressTokens (every step), 2 = SuppressBlank (first sampled position only)
  std::vector<unsigned char> mask(D.n_vocab, 0);
  if (ts_on) mask[opts->no_timestamps] = 1;   // ApplyTimestampRules: <|notimestamps|> is never sampled
  for (int i = 0; i < opts->n_suppress_begin; ++i)
    if (opts->suppress_begin[i] >= 0 && opts->suppress_begin[i] < D.n_vocab) mask[opts->suppress_begin[i]] = 2;
  for (int i = 0; i < opts->n_suppress; ++i)
    if (opts->suppress[i] >= 0 && opts->suppress[i] < D.n_vocab) mask[opts->suppress[i]] = 1;
  WB_CUDA_OK(cudaMemcpyAsync(h->tokens, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemcpyAsync(h->mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemsetAsync(h->sum_logprob, 0, sizeof(float) * B, st));
  WB_CUDA_OK(cudaMemsetAsync(h->done, 0, sizeof(int32_t) * B, st));
  WB_CUDA_OK(cudaMemsetAsync(h->ts_state, 0, sizeof(int4) * B, st));
  WB_CUDA_OK(cudaStreamSynchronize(st));   // `rows` / `mask` are pageable host memory

  // sub-batches: the chunks are decoded as up to 4 independent groups on their own streams, so the latency-bound GEMM
  // chain of one group runs under the bandwidth-bound attention of another. Per-sequence arithmetic is unchanged.
  int nsb = 1;   // measured on B200 at B=32: 2 groups 81.9 ms vs 79.3 ms for one (kernels of different groups rarely co-reside); kept as a knob
  if (const char* e = getenv("WB_SUBBATCHES")) nsb = atoi(e);
  nsb = nsb < 1 ? 1 : (nsb > wb_handle::kMaxSub ? wb_handle::kMaxSub : nsb);
  if (nsb > B) nsb = B;
  const int per = (B + nsb - 1) / nsb;
  StepOpts plain[wb_handle::kMaxSub], samp[wb_handle::kMaxSub];
  for (int i = 0; i < nsb; ++i) {
    StepOpts o{};
    o.b0 = i * per, o.Mb = (B - o.b0) < per ? (B - o.b0) : per, o.sub = i;
    o.beams = 1, o.store_logits = 0, o.sample = 0, o.n_initial = n_init, o.eot = opts->eot;
    o.timestamps = ts_on ? 1 : 0, o.ts_begin = opts->timestamp_begin;
    o.ts_last_allowed = opts->max_initial_timestamp_index >= 0 ? opts->timestamp_begin + opts->max_initial_timestamp_index : 0x7fffffff;
    plain[i] = o;
    o.sample = 1;
    samp[i] = o;
  }

  // the host looks at the EOT flags every `interval` steps anyway: that many steps go into one graph
  const int interval = opts->eot_check_interval > 0 ? opts->eot_check_interval : 8;
  int multi = interval;
  if (const char* e = getenv("WB_GRAPH_STEPS")) multi = atoi(e);
  multi = multi < 1 ? 1 : (multi > 32 ? 32 : multi);
  const bool use_graph = getenv("WB_NO_GRAPH") == nullptr;
  // two sub-batches as one interleaved schedule (decode_steps_pair): graph replay only, the three-kernel decoder path
  bool pair = false;
  if (const char* e = getenv("WB_PAIR")) pair = e[0] == '1';
  pair = pair && nsb == 2 && use_graph && opts->no_speech_prob == nullptr;
  char key[128];
  snprintf(key, sizeof(key), "B%d i%d e%d s%d m%d t%d.%d.%d p%d", B, n_init, opts->eot, nsb, multi, ts_on ? 1 : 0,
           ts_on ? opts->timestamp_begin : 0, ts_on ? opts->max_initial_timestamp_index : 0, pair ? 1 : 0);
  if (use_graph && h->graph_key != key) {
    destroy_graphs(h);
    h->sample_n = multi;
    for (int i = 0; i < nsb; ++i) {
      // one eager pass of each variant first: sets function attributes and faults in code outside of capture
      WB_TRY(reset_decode_state(h, plain[i]));
      WB_TRY(decode_step(h, plain[i]));
      WB_TRY(decode_step(h, samp[i]));
      WB_CUDA_OK(cudaStreamSynchronize(step_stream(h, plain[i])));
      if (pair) continue;
      WB_TRY(capture(h, plain[i], &h->g_step[i], &h->nodes_step));
      WB_TRY(capture(h, samp[i], &h->g_sample[i], &h->nodes_sample));
      if (h->sample_n > 1) {
        int64_t nodes_n = 0;
        WB_TRY(capture(h, samp[i], &h->g_sample_n[i], &nodes_n, h->sample_n));
      }
    }
    if (pair) {
      WB_TRY(capture_pair(h, plain[0], plain[1], &h->g_pair[0], &h->nodes_pair[0], 1));
      WB_TRY(capture_pair(h, samp[0], samp[1], &h->g_pair[1], &h->nodes_pair[1], 1));
      h->g_pair[2] = nullptr, h->nodes_pair[2] = 0;
      if (h->sample_n > 1) WB_TRY(capture_pair(h, samp[0], samp[1], &h->g_pair[2], &h->nodes_pair[2], h->sample_n));
    }
    h->graph_key = key;
    // the eager passes wrote tokens and log-probs: restore
    WB_CUDA_OK(cudaMemcpyAsync(h->tokens, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    WB_CUDA_OK(cudaMemsetAsync(h->sum_logprob, 0, sizeof(float) * B, st));
    WB_CUDA_OK(cudaMemsetAsync(h->done, 0, sizeof(int32_t) * B, st));
    WB_CUDA_OK(cudaMemsetAsync(h->ts_state, 0, sizeof(int4) * B, st));
    WB_CUDA_OK(cudaStreamSynchronize(st));
  }

  WB_CUDA_OK(cudaEventRecord(h->ev[2], st));
  WB_CUDA_OK(cudaEventRecord(h->fork_ev, st));
  for (int i = 1; i < nsb; ++i) WB_CUDA_OK(cudaStreamWaitEvent(h->sub_stream[i], h->fork_ev, 0));
  for (int i = 0; i < nsb; ++i) WB_TRY(reset_decode_state(h, plain[i]));
  if (pair) {   // the pair graphs run from the main stream: the second sub-batch's reset joins it first
    WB_CUDA_OK(cudaEventRecord(h->sub_ev[1], h->sub_stream[1]));
    WB_CUDA_OK(cudaStreamWaitEvent(st, h->sub_ev[1], 0));
  }
  const bool probe = opts->no_speech_prob != nullptr;
  for (int k = 0; k + 1 < n_init; ++k) {
    if (pair) {
      WB_CUDA_OK(cudaGraphLaunch(h->g_pair[0], st));
      h->launches += h->nodes_pair[0];
      continue;
    }
    for (int i = 0; i < nsb; ++i) {
      if (probe && k == opts->sot_index) {
        WB_TRY(no_speech_probe_step(h, plain[i], opts->no_speech, k, true));
      } else if (use_graph) {
        WB_CUDA_OK(cudaGraphLaunch(h->g_step[i], step_stream(h, plain[i])));
        h->launches += h->nodes_step;
      } else {
        WB_TRY(decode_step(h, plain[i]));
      }
    }
  }
  if (probe && opts->sot_index == n_init - 1)
    for (int i = 0; i < nsb; ++i) WB_TRY(no_speech_probe_step(h, plain[i], opts->no_speech, n_init - 1, false));
  int steps = 0;
  // sub-batch i starts i * stagger microseconds late (again after every EOT poll, which re-aligns the streams)
  int stagger_us = 0;
  if (const char* e = getenv("WB_STAGGER_US")) stagger_us = atoi(e);
  bool restagger = nsb > 1 && stagger_us > 0;
  for (int s = 0; s < opts->sample_len;) {
    const int n = (use_graph && multi > 1 && s + multi <= opts->sample_len && (s % interval) + multi <= interval) ? multi : 1;
    if (restagger) {
      for (int i = 1; i < nsb; ++i) WB_TRY(launch_delay((unsigned long long)i * stagger_us * 1000ull, step_stream(h, samp[i]), &h->launches));
      restagger = false;
    }
    for (int i = 0; i < nsb; ++i) {
      if (pair) {
        if (i == 0) {
          WB_CUDA_OK(cudaGraphLaunch(n > 1 ? h->g_pair[2] : h->g_pair[1], st));
          h->launches += n > 1 ? h->nodes_pair[2] : h->nodes_pair[1];
        }
      } else if (use_graph) {
        WB_CUDA_OK(cudaGraphLaunch(n > 1 ? h->g_sample_n[i] : h->g_sample[i], step_stream(h, samp[i])));
        h->launches += h->nodes_sample * n;
      } else {
        WB_TRY(decode_step(h, samp[i]));
      }
    }
    s += n;
    steps += n;
    if (s % interval == 0 && s < opts->sample_len) {
      for (int i = 0; i < nsb; ++i)
        WB_CUDA_OK(cudaMemcpyAsync(h->h_done + plain[i].b0, h->done + plain[i].b0, sizeof(int32_t) * plain[i].Mb, cudaMemcpyDeviceToHost,
                                   pair ? st : step_stream(h, plain[i])));
      for (int i = 0; i < nsb; ++i) WB_CUDA_OK(cudaStreamSynchronize(step_stream(h, plain[i])));
      bool all = true;
      for (int b = 0; b < B; ++b) all = all && h->h_done[b];
      if (all) break;
      restagger = nsb > 1 && stagger_us > 0;
    }
  }
  for (int i = 1; i < nsb; ++i) {   // join: everything after this point is ordered after every sub-batch
    WB_CUDA_OK(cudaEventRecord(h->sub_ev[i], h->sub_stream[i]));
    WB_CUDA_OK(cudaStreamWaitEvent(st, h->sub_ev[i], 0));
  }
  WB_CUDA_OK(cudaEventRecord(h->ev[3], st));
  h->timings[3] = (float)(steps + n_init - 1);
  WB_CUDA_OK(cudaMemcpyAsync(rows.data(), h->tokens, rows.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  std::vector<float> slp(B);
  WB_CUDA_OK(cudaMemcpyAsync(slp.data(), h->sum_logprob, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
  if (probe) WB_CUDA_OK(cudaMemcpyAsync(opts->no_speech_prob, h->no_speech_prob, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
  WB_CUDA_OK(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b) {
    const int32_t* r = &rows[(size_t)b * h->tokens_ld];
    int len = total;
    for (int i = n_init; i < total; ++i)
      if (r[i] == opts->eot) {
        len = i + 1;
        break;
      }
    for (int i = 0; i < total; ++i) tokens_out[(size_t)b * total + i] = i < len ? r[i] : opts->eot;
    if (lens) lens[b] = len;
    if (sum_logprob) sum_logprob[b] = slp[b];
  }
  return WB_OK;
}

// Temperature > 0 (upstream GreedyDecoder with Categorical sampling, DecodingOptions.best_of, MaximumLikelihoodRanker): the step
// stores the filtered logits, sample_rows_kernel draws per sequence, the finish kernel takes the drawn token. best_of samples
// of a chunk are rows of the same batch that share the chunk's cross K/V. Runs eagerly: it is the fallback path of
// wb_transcribe_long, taken only for windows whose arg-max decode failed the quality thresholds.
static int decode_sampled(wb_handle* h, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens, float* sum_logprob) {
  const wb_dims& D = h->dims;
  const int G = opts->best_of > 1 ? opts->best_of : 1, Mb = B * G, V = D.n_vocab;
  const int n_init = opts->n_initial, total = n_init + opts->sample_len, eot = opts->eot;
  const bool ts_on = opts->timestamps != 0;
  if (G > h->max_beams || Mb > h->Mb_max || (ts_on && Mb > 64)) {
    set_error("wb_decode: best_of %d exceeds the handle's max_beams %d, or batch * best_of %d exceeds %d (64 with timestamp rules)", G,
              h->max_beams, Mb, h->Mb_max);
    return WB_ERR_ARG;
  }
  cudaStream_t st = h->stream;
  std::vector<int32_t> rows((size_t)Mb * h->tokens_ld, eot);
  for (int b = 0; b < Mb; ++b)
    for (int i = 0; i < n_init; ++i) rows[(size_t)b * h->tokens_ld + i] = opts->initial_tokens[i];
  std::vector<unsigned char> mask(V, 0);
  if (ts_on) mask[opts->no_timestamps] = 1;
  for (int i = 0; i < opts->n_suppress_begin; ++i)
    if (opts->suppress_begin[i] >= 0 && opts->suppress_begin[i] < V) mask[opts->suppress_begin[i]] = 2;
  for (int i = 0; i < opts->n_suppress; ++i)
    if (opts->suppress[i] >= 0 && opts->suppress[i] < V) mask[opts->suppress[i]] = 1;
  WB_CUDA_OK(cudaMemcpyAsync(h->tokens, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemcpyAsync(h->mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemsetAsync(h->sum_logprob, 0, sizeof(float) * Mb, st));
  WB_CUDA_OK(cudaMemsetAsync(h->done, 0, sizeof(int32_t) * Mb, st));
  WB_CUDA_OK(cudaMemsetAsync(h->ts_state, 0, sizeof(int4) * Mb, st));
  WB_CUDA_OK(cudaStreamSynchronize(st));   // `rows` / `mask` are pageable host memory

  StepOpts plain{};
  plain.Mb = Mb, plain.beams = G, plain.n_initial = n_init, plain.eot = eot;
  plain.timestamps = ts_on ? 1 : 0, plain.ts_begin = opts->timestamp_begin;
  plain.ts_last_allowed = opts->max_initial_timestamp_index >= 0 ? opts->timestamp_begin + opts->max_initial_timestamp_index : 0x7fffffff;
  StepOpts scored = plain, fin = plain;
  scored.sample = 1, scored.store_logits = 1, scored.no_finish = 1;
  fin.sample = 1, fin.use_chosen = 1;
  WB_CUDA_OK(cudaEventRecord(h->ev[2], st));
  WB_TRY(reset_decode_state(h, plain));
  const bool probe = opts->no_speech_prob != nullptr;
  for (int k = 0; k + 1 < n_init; ++k) {
    if (probe && k == opts->sot_index)
      WB_TRY(no_speech_probe_step(h, plain, opts->no_speech, k, true));
    else
      WB_TRY(decode_step(h, plain));
  }
  if (probe && opts->sot_index == n_init - 1) WB_TRY(no_speech_probe_step(h, plain, opts->no_speech, n_init - 1, false));
  const int interval = opts->eot_check_interval > 0 ? opts->eot_check_interval : 8;
  int steps = 0;
  for (int s = 0; s < opts->sample_len; ++s) {
    WB_TRY(decode_step(h, scored));
    WB_TRY(launch_sample_rows(h->logits, Mb, V, ts_on ? opts->timestamp_begin : 0x7fffffff, opts->temperature, opts->seed, h->state, h->chosen,
                              h->chosen_lp, st, &h->launches));
    WB_TRY(step_finish(h, fin, 1));
    ++steps;
    if ((s + 1) % interval == 0 && s + 1 < opts->sample_len) {
      WB_CUDA_OK(cudaMemcpyAsync(h->h_done, h->done, sizeof(int32_t) * Mb, cudaMemcpyDeviceToHost, st));
      WB_CUDA_OK(cudaStreamSynchronize(st));
      bool all = true;
      for (int b = 0; b < Mb; ++b) all = all && h->h_done[b];
      if (all) break;
    }
  }
  WB_CUDA_OK(cudaEventRecord(h->ev[3], st));
  h->timings[3] = (float)(steps + n_init - 1);
  WB_CUDA_OK(cudaMemcpyAsync(rows.data(), h->tokens, rows.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  std::vector<float> slp(Mb), nsp(Mb, 0.f);
  WB_CUDA_OK(cudaMemcpyAsync(slp.data(), h->sum_logprob, sizeof(float) * Mb, cudaMemcpyDeviceToHost, st));
  if (probe) WB_CUDA_OK(cudaMemcpyAsync(nsp.data(), h->no_speech_prob, sizeof(float) * Mb, cudaMemcpyDeviceToHost, st));
  WB_CUDA_OK(cudaStreamSynchronize(st));
  for (int a = 0; a < B; ++a) {
    // MaximumLikelihoodRanker without length penalty: sum_logprob / number of sampled tokens before the first eot
    int best = -1, best_len = 0;
    double best_score = 0.0;
    for (int j = 0; j < G; ++j) {
      const int32_t* r = &rows[(size_t)(a * G + j) * h->tokens_ld];
      int len = total;
      for (int i = n_init; i < total; ++i)
        if (r[i] == eot) {
          len = i + 1;
          break;
        }
      const int n_text = (len < total || r[total - 1] == eot ? len - 1 : len) - n_init;   // tokens before the first eot
      const double score = (double)slp[a * G + j] / (double)(n_text > 0 ? n_text : 1);
      if (best < 0 || score > best_score) best = j, best_score = score, best_len = len;
    }
    const int32_t* r = &rows[(size_t)(a * G + best) * h->tokens_ld];
    for (int i = 0; i < total; ++i) tokens_out[(size_t)a * total + i] = i < best_len ? r[i] : eot;
    if (lens) lens[a] = best_len;
    if (sum_logprob) sum_logprob[a] = slp[a * G + best];
    if (probe) opts->no_speech_prob[a] = nsp[a * G + best];
  }
  return WB_OK;
}

// Beam search (upstream whisper/decoding.py BeamSearchDecoder + MaximumLikelihoodRanker): the device runs the
// decoder step for n_audio*beam sequences (cross K/V shared by the beams of a chunk) and extracts the top beam+1
// log-probabilities per row; the candidate bookkeeping of a step is a few hundred scalar operations and runs on the host;
// the self-attention cache is re-indexed by source beam on the device.
static double now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int decode_beam(wb_handle* h, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens, float* sum_logprob) {
  const wb_dims& D = h->dims;
  const int beam = opts->beam_size, Mb = B * beam, K = beam + 1;
  const int n_init = opts->n_initial, total = n_init + opts->sample_len, eot = opts->eot;
  if (beam > h->max_beams || Mb > h->Mb_max || h->selfK_alt.empty()) {
    set_error("wb_decode: beam_size %d exceeds the handle's max_beams %d (or 7), or batch*beam %d exceeds %d", beam, h->max_beams, Mb, h->Mb_max);
    return WB_ERR_ARG;
  }
  cudaStream_t st = h->stream;
  std::vector<int32_t> rows((size_t)Mb * h->tokens_ld, eot);
  for (int b = 0; b < Mb; ++b)
    for (int i = 0; i < n_init; ++i) rows[(size_t)b * h->tokens_ld + i] = opts->initial_tokens[i];
  std::vector<unsigned char> mask(D.n_vocab, 0);
  for (int i = 0; i < opts->n_suppress_begin; ++i)
    if (opts->suppress_begin[i] >= 0 && opts->suppress_begin[i] < D.n_vocab) mask[opts->suppress_begin[i]] = 2;
  for (int i = 0; i < opts->n_suppress; ++i)
    if (opts->suppress[i] >= 0 && opts->suppress[i] < D.n_vocab) mask[opts->suppress[i]] = 1;
  const bool ts_on = opts->timestamps != 0;
  const int ts_begin = ts_on ? opts->timestamp_begin : 0x7fffffff;
  if (ts_on) mask[opts->no_timestamps] = 1;   // ApplyTimestampRules: <|notimestamps|> is never sampled
  WB_CUDA_OK(cudaMemcpyAsync(h->tokens, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemcpyAsync(h->mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaStreamSynchronize(st));

  StepOpts plain{}, scored{};
  plain.Mb = Mb, plain.beams = beam, plain.n_initial = n_init, plain.eot = eot;
  scored = plain;
  scored.sample = 1, scored.store_logits = 1, scored.no_finish = 1;   // filtered logits stored, host decides the next tokens
  // Timestamp rules: the per-sequence rule state the finish kernel keeps in greedy decoding is a function of the sequence, and
  // the sequences live on the host here: it is recomputed per step and uploaded in front of the scored step; the logits kernel
  // applies the per-row rules, the top-k kernel the probability-mass rule.
  scored.timestamps = ts_on ? 1 : 0, scored.ts_begin = opts->timestamp_begin;
  scored.ts_last_allowed = opts->max_initial_timestamp_index >= 0 ? opts->timestamp_begin + opts->max_initial_timestamp_index : 0x7fffffff;
  int4* const ts_host2 = h->hb_ts;   // [2][Mb], by step parity
  WB_CUDA_OK(cudaEventRecord(h->ev[2], st));
  WB_TRY(reset_decode_state(h, plain));
  for (int i = 0; i + 1 < n_init; ++i) WB_TRY(decode_step(h, plain));

  typedef std::vector<int32_t> Seq;
  std::vector<Seq> seqs(Mb, Seq(opts->initial_tokens, opts->initial_tokens + n_init));
  std::vector<float> slp(Mb, 0.f);
  std::vector<std::vector<std::pair<Seq, float>>> finished(B);   // insertion-ordered, unique keys (Python dict semantics)
  const size_t max_candidates = (size_t)beam;                    // round(beam_size * patience), patience = 1
  float* const top_lp = h->hb_top_lp;
  int32_t* const top_idx = h->hb_top_idx;
  int steps = 0;
  // The scored step (decoder step with stored filtered logits + top-k) replays from a CUDA graph: on the wide models it is 7
  // kernels per layer, and launched one by one the host, not the GPU, paces the step. The re-indexing swaps the two K/V buffer
  // sets, so there is one graph per orientation (keyed by the buffer the first layer reads), captured when first needed.
  const bool prof = getenv("WB_BEAM_PROF") != nullptr;   // development: where a beam step's wall time goes
  double t_wait = 0.0, t_host = 0.0, t_reorder = 0.0;
  char bkey[96];
  snprintf(bkey, sizeof(bkey), "B%d b%d i%d e%d t%d.%d.%d", B, beam, n_init, eot, ts_on ? 1 : 0, ts_on ? opts->timestamp_begin : 0,
           ts_on ? opts->max_initial_timestamp_index : 0);
  const bool beam_graph = getenv("WB_NO_GRAPH") == nullptr;
  if (h->beam_key != bkey) {
    for (int i = 0; i < 2; ++i) {
      if (h->g_beam[i]) cudaGraphExecDestroy(h->g_beam[i]);
      h->g_beam[i] = nullptr, h->beam_kv[i] = nullptr;
    }
    h->beam_key = bkey;
  }
  for (int s = 0; s < opts->sample_len; ++s) {
    if (ts_on) {
      int4* const ts_host = ts_host2 + (size_t)(s & 1) * Mb;
      for (int b = 0; b < Mb; ++b) {   // FinishDesc::ts_state from the sampled part of the sequence (empty at s = 0)
        const Seq& q = seqs[b];
        const int len = (int)q.size() - n_init;
        int z = 0;
        for (int i = n_init; i < (int)q.size(); ++i)
          if (q[i] >= ts_begin) z = q[i];
        const bool last_ts = len >= 1 && q.back() >= ts_begin;
        const bool penult_ts = len < 2 || q[q.size() - 2] >= ts_begin;
        int4 v;
        v.x = (last_ts && penult_ts ? 1 : 0) | (last_ts && !penult_ts ? 2 : 0);
        v.y = z ? ((last_ts && !penult_ts) ? z : z + 1) : 0;
        v.z = z, v.w = 0;
        ts_host[b] = v;
      }
      WB_CUDA_OK(cudaMemcpyAsync(h->ts_state, ts_host, (size_t)Mb * sizeof(int4), cudaMemcpyHostToDevice, st));
    }
    int slot = -1;
    if (beam_graph && s > 0) {   // the first scored step runs eagerly (function attributes are set outside of capture)
      for (int i = 0; i < 2; ++i)
        if (h->beam_kv[i] == h->selfK[0]) slot = i;
      if (slot < 0) {
        slot = h->beam_kv[0] ? 1 : 0;
        if (h->g_beam[slot]) cudaGraphExecDestroy(h->g_beam[slot]);
        h->g_beam[slot] = nullptr;
        cudaGraph_t g;
        const int64_t before = h->launches;
        WB_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = decode_step(h, scored);
        if (rc == 0) rc = launch_topk_logprobs(h->logits, Mb, D.n_vocab, K, ts_begin, h->top_lp, h->top_idx, st, &h->launches);
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        h->nodes_beam = h->launches - before;
        h->launches = before;
        if (rc) return rc;
        WB_CUDA_OK(ce);
        WB_CUDA_OK(cudaGraphInstantiate(&h->g_beam[slot], g, 0));
        WB_CUDA_OK(cudaGraphDestroy(g));
        h->beam_kv[slot] = h->selfK[0];
      }
      WB_CUDA_OK(cudaGraphLaunch(h->g_beam[slot], st));
      h->launches += h->nodes_beam;
    } else {
      WB_TRY(decode_step(h, scored));
      WB_TRY(launch_topk_logprobs(h->logits, Mb, D.n_vocab, K, ts_begin, h->top_lp, h->top_idx, st, &h->launches));
    }
    WB_CUDA_OK(cudaMemcpyAsync(top_lp, h->top_lp, (size_t)Mb * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    WB_CUDA_OK(cudaMemcpyAsync(top_idx, h->top_idx, (size_t)Mb * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    const double t_a = prof ? now_ms() : 0.0;
    WB_CUDA_OK(cudaStreamSynchronize(st));
    const double t_b = prof ? now_ms() : 0.0;
    t_wait += t_b - t_a;
    ++steps;
    std::vector<Seq> next_seqs;
    std::vector<float> next_slp;
    std::vector<int32_t> next_src;
    for (int a = 0; a < B; ++a) {
      // STEP 1: cumulative log-probabilities of the candidates, keyed by sequence (upstream's dict: a sequence reached twice keeps
      // its first position and takes the later score and source). A candidate is (source beam, token): two of them are the
      // same sequence iff the tokens agree and the source beams hold equal sequences, so the beams are classed once per step
      // (at the first step all of them are the prompt) and no candidate sequence is built before it is selected.
      struct Cand { int source, token, cls; float score; };
      int cls[8];
      for (int j = 0; j < beam; ++j) {
        cls[j] = j;
        for (int i = 0; i < j; ++i)
          if (cls[i] == i && seqs[a * beam + i] == seqs[a * beam + j]) {
            cls[j] = i;
            break;
          }
      }
      Cand cands[8 * 8];
      int n_cand = 0;
      for (int j = 0; j < beam; ++j) {
        const int idx = a * beam + j;
        for (int c = 0; c < K; ++c) {
          const int tok = top_idx[(size_t)idx * 8 + c];
          const float sc = slp[idx] + top_lp[(size_t)idx * 8 + c];
          bool found = false;
          for (int e = 0; e < n_cand; ++e)
            if (cands[e].cls == cls[j] && cands[e].token == tok) {
              cands[e].score = sc, cands[e].source = idx, found = true;
              break;
            }
          if (!found) cands[n_cand++] = Cand{idx, tok, cls[j], sc};
        }
      }
      // STEP 2: rank, keep the best `beam` unfinished sequences; sequences ending in eot are set aside
      std::stable_sort(cands, cands + n_cand, [](const Cand& x, const Cand& y) { return x.score > y.score; });
      std::vector<std::pair<Seq, float>> newly;
      int saved = 0;
      for (int e = 0; e < n_cand; ++e) {
        const Seq& from = seqs[cands[e].source];
        Seq sq;
        sq.reserve(from.size() + 1);
        sq.insert(sq.end(), from.begin(), from.end());
        sq.push_back(cands[e].token);
        if (cands[e].token == eot) {
          newly.emplace_back(std::move(sq), cands[e].score);
        } else {
          next_seqs.push_back(std::move(sq)), next_slp.push_back(cands[e].score), next_src.push_back(cands[e].source);
          if (++saved == beam) break;
        }
      }
      while (saved < beam) {   // cannot happen with beam+1 candidates per beam and one eot token, kept for safety
        next_seqs.push_back(next_seqs.back()), next_slp.push_back(-INFINITY), next_src.push_back(next_src.back());
        ++saved;
      }
      for (auto& nf : newly) {   // already in descending score order
        if (finished[a].size() >= max_candidates) break;
        bool found = false;
        for (auto& pf : finished[a])
          if (pf.first == nf.first) {
            pf.second = nf.second, found = true;
            break;
          }
        if (!found) finished[a].push_back(nf);
      }
    }
    seqs.swap(next_seqs), slp.swap(next_slp);
    if (prof) t_host += now_ms() - t_b;
    bool completed = true;
    for (int a = 0; a < B; ++a) completed = completed && finished[a].size() >= max_candidates;
    if (completed || s + 1 == opts->sample_len) break;
    // device side of the update: newest token column, cache re-indexing, then embed + advance
    bool identity = true;
    int32_t* const src = h->hb_src + (size_t)(s & 1) * Mb;        // read by the copies below while the next step's are written
    int32_t* const next_col = h->hb_col + (size_t)(s & 1) * Mb;
    for (int b = 0; b < Mb; ++b) {
      src[b] = next_src[b], next_col[b] = seqs[b].back();
      identity = identity && src[b] == b;
    }
    const int col = n_init + s;
    WB_CUDA_OK(cudaMemcpy2DAsync(h->tokens + col, h->tokens_ld * sizeof(int32_t), next_col, sizeof(int32_t), sizeof(int32_t), Mb,
                                 cudaMemcpyHostToDevice, st));
    if (!identity) {
      WB_CUDA_OK(cudaMemcpyAsync(h->beam_src, src, Mb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      WB_TRY(launch_reorder_kv(h->selfK.data(), h->selfV.data(), h->selfK_alt.data(), h->selfV_alt.data(), D.n_text_layer, Mb,
                               D.n_text_ctx, D.n_text_state, h->beam_src, h->state, st, &h->launches));
      h->selfK.swap(h->selfK_alt), h->selfV.swap(h->selfV_alt);
    }
    WB_TRY(step_finish(h, plain, 0));
    // no synchronisation here: the staging buffers alternate by step parity, and the copies of step s have run before the
    // synchronisation that follows the scored step of s + 1
    (void)t_reorder;
  }
  if (prof)
    fprintf(stderr, "[wb] beam search, %d steps: waiting for the scored step %.1f ms, host bookkeeping %.1f ms, waiting for re-index + embed %.1f ms\n",
            steps, t_wait, t_host, t_reorder);
  WB_CUDA_OK(cudaEventRecord(h->ev[3], st));
  h->timings[3] = (float)(steps + n_init - 1);
  // finalize: unfinished beams (eot appended) fill up to `beam` candidates; rank by sum_logprob / length
  for (int a = 0; a < B; ++a) {
    auto& fin = finished[a];
    if ((int)fin.size() < beam) {
      std::vector<int> order(beam);
      for (int j = 0; j < beam; ++j) order[j] = j;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return slp[a * beam + x] > slp[a * beam + y]; });
      for (int j : order) {
        Seq sq = seqs[a * beam + j];
        sq.push_back(eot);
        bool found = false;
        for (auto& pf : fin)
          if (pf.first == sq) {
            pf.second = slp[a * beam + j], found = true;
            break;
          }
        if (!found) fin.emplace_back(sq, slp[a * beam + j]);
        if ((int)fin.size() >= beam) break;
      }
    }
    int best = 0;
    float best_rank = -INFINITY;
    std::vector<Seq> trimmed(fin.size());
    for (size_t c = 0; c < fin.size(); ++c) {
      const Seq& sq = fin[c].first;
      size_t end = sq.size();
      for (size_t i = n_init; i < sq.size(); ++i)
        if (sq[i] == eot) {
          end = i;
          break;
        }
      trimmed[c].assign(sq.begin() + n_init, sq.begin() + end);
      const float rank = trimmed[c].empty() ? -INFINITY : fin[c].second / (float)trimmed[c].size();
      if (rank > best_rank) best_rank = rank, best = (int)c;
    }
    int len = n_init + (int)trimmed[best].size();
    for (int i = 0; i < total; ++i) {
      int32_t tk = eot;
      if (i < n_init) tk = opts->initial_tokens[i];
      else if (i < len) tk = trimmed[best][i - n_init];
      tokens_out[(size_t)a * total + i] = tk;
    }
    if (lens) lens[a] = len + 1 <= total ? len + 1 : total;
    if (sum_logprob) sum_logprob[a] = fin[best].second;
  }
  return WB_OK;
}

int wb_decode(wb_handle* h, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens, float* sum_logprob) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_features(h, B));
  if (!opts || !tokens_out) {
    set_error("wb_decode: null options or output");
    return WB_ERR_ARG;
  }
  WB_TRY(validate_decode_opts(h, B, opts));
  if (opts->beam_size > 1) return decode_beam(h, B, opts, tokens_out, lens, sum_logprob);
  if (opts->temperature > 0.f) return decode_sampled(h, B, opts, tokens_out, lens, sum_logprob);
  return decode_greedy(h, B, opts, tokens_out, lens, sum_logprob);
}

int wb_transcribe_dev(wb_handle* h, const float* audio_dev, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out,
                      int32_t* lens, float* sum_logprob) {
  WB_TRY(wb_encode_dev(h, audio_dev, B));
  WB_TRY(wb_decode(h, B, opts, tokens_out, lens, sum_logprob));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->timings[0] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->timings[1] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->timings[2] = ms;
  return WB_OK;
}

// Slab sizes for a host batch of B chunks: the encoder takes ~3.4x as long per chunk as the PCIe copy, so after a first slab
// (whose copy is the only exposed one) every later copy runs under the previous slab's compute. Measured at base.en, 32 chunks
// (gpurun_out/exp1_slabs.log): pinned host memory 68.5 -> 68.1 ms per step with {8, 24} (smaller first slabs lose more in
// the encoder's tile quantisation than their copy saves), pageable 73.5 -> 71.6 ms ({4, 12, 16}: 70.3 ms).
// WB_H2D_SLABS="a,b,c" (chunks per slab, the last one takes the rest) overrides; "0" = one copy.
static int slab_plan(int B, bool pageable, int (&slab)[wb_handle::kMaxSlabs]) {
  int n = 0, used = 0;
  if (const char* e = getenv("WB_H2D_SLABS")) {
    while (*e && n < wb_handle::kMaxSlabs - 1) {
      const int v = atoi(e);
      if (v <= 0 || used + v >= B) break;
      slab[n++] = v, used += v;
      while (*e && *e != ',') ++e;
      if (*e == ',') ++e;
    }
  } else if (pageable && B >= 16) {   // a pageable copy blocks the host and is slower: a smaller first slab starts the GPU earlier
    slab[n++] = B / 8, used += B / 8;
    slab[n++] = (3 * B) / 8, used += (3 * B) / 8;
  } else if (B >= 8) {
    slab[n++] = B / 4, used += B / 4;
  }
  slab[n++] = B - used;
  return n;
}

int wb_transcribe(wb_handle* h, const float* audio, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens,
                  float* sum_logprob) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_weights(h));
  if (!audio) return WB_ERR_ARG;
  int slab[wb_handle::kMaxSlabs];
  cudaPointerAttributes pa{};
  const bool pageable = cudaPointerGetAttributes(&pa, audio) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();   // an ordinary malloc'ed pointer may report an error on older drivers: not ours to keep
  const int n_slab = slab_plan(B, pageable, slab);
  // the copies may not overtake work already queued on the handle's stream that still reads audio_dev
  WB_CUDA_OK(cudaEventRecord(h->copy_fence, h->stream));
  WB_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->copy_fence, 0));
  WB_CUDA_OK(cudaEventRecord(h->ev[0], h->stream));
  h->enc_batch = 0;
  for (int s = 0, b0 = 0; s < n_slab; b0 += slab[s], ++s) {
    const size_t off = (size_t)b0 * WB_N_SAMPLES;
    WB_CUDA_OK(cudaMemcpyAsync(h->audio_dev + off, audio + off, (size_t)slab[s] * WB_N_SAMPLES * sizeof(float), cudaMemcpyHostToDevice,
                               h->copy_stream));
    WB_CUDA_OK(cudaEventRecord(h->copy_ev[s], h->copy_stream));
    WB_CUDA_OK(cudaStreamWaitEvent(h->stream, h->copy_ev[s], 0));
    WB_TRY(launch_logmel<float>(h->audio_dev + off, WB_N_SAMPLES, 0, slab[s], h->tab32, h->logspec + (size_t)b0 * WB_N_MELS * WB_N_FRAMES,
                                reinterpret_cast<unsigned int*>(h->gmax) + b0 /* OrderedMax<float>::U */, nullptr,
                                h->melT + (size_t)b0 * (WB_N_FRAMES + 2) * WB_N_MELS, h->stream, &h->launches));
    if (s == 0) WB_CUDA_OK(cudaEventRecord(h->ev[1], h->stream));   // timings[0]: the first slab's copy + log-mel
    WB_TRY(encoder_forward(h, b0, slab[s]));
    WB_TRY(cross_kv(h, b0, slab[s]));
  }
  WB_CUDA_OK(cudaEventRecord(h->ev[2], h->stream));
  WB_TRY(wb_decode(h, B, opts, tokens_out, lens, sum_logprob));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->timings[0] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->timings[1] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->timings[2] = ms;
  return WB_OK;
}

int64_t wb_launch_count(const wb_handle* h) { return h ? h->launches : -1; }

/* development: copy out the %globaltimer trace of the last decode (WB_TRACE=1); returns the number of records */
int wb_debug_trace_sub(wb_handle* h, int sub, unsigned long long* out, int max_records) {
  if (!h || !h->trace || !out || sub < 0 || sub >= wb_handle::kMaxSub) return 0;
  DecodeState st;
  if (cudaMemcpy(&st, h->state + sub, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  int n = st.trace_n < max_records ? st.trace_n : max_records;
  n = n < 16384 ? n : 16384;
  if (cudaMemcpy(out, h->trace + (size_t)sub * 16384 * 8, (size_t)n * 64, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return n;
}
int wb_debug_trace(wb_handle* h, unsigned long long* out, int max_records) { return wb_debug_trace_sub(h, 0, out, max_records); }
int wb_last_timings(const wb_handle* h, float out[4]) {
  if (!h || !out) return WB_ERR_ARG;
  for (int i = 0; i < 4; ++i) out[i] = h->timings[i];
  return WB_OK;
}
int wb_sync(wb_handle* h) {
  if (!h) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return WB_OK;
}

int wb_profile_cross_attention(wb_handle* h, int32_t B, int32_t reps, float* avg_ms, double* bytes_per_launch) {
  WB_TRY(check_batch(h, B));
  WB_TRY(need_features(h, B));
  if (!avg_ms || reps < 1) return WB_ERR_ARG;
  const wb_dims& D = h->dims;
  AttnDecodeDesc c{};
  c.Mb = B, c.d = D.n_text_state, c.n_head = D.n_text_head, c.q = nullptr, c.x = h->xdec;
  c.n_ctx = D.n_audio_ctx, c.n_rows_fixed = D.n_audio_ctx, c.kv_share = 1, c.state = h->state, c.out16 = h->a16, c.tmaps = h->gemm;
  // the kernel exactly as the decode step runs it: with the fused LayerNorm + query projection prologue
  for (int i = -3; i < reps; ++i) {   // 3 warm-up launches
    if (i == 0) WB_CUDA_OK(cudaEventRecord(h->ev[0], h->stream));
    const int l = ((i % D.n_text_layer) + D.n_text_layer) % D.n_text_layer;
    c.k = h->crossK[l], c.v = h->crossV[l];
    c.ln_g = h->dec[l].lnc_g, c.ln_b = h->dec[l].lnc_b, c.wq = h->dec[l].wq_c, c.bq = h->dec[l].bq_c;
    WB_TRY(launch_attn_decode(c, h->stream, &h->launches));
  }
  WB_CUDA_OK(cudaEventRecord(h->ev[1], h->stream));
  WB_CUDA_OK(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  WB_CUDA_OK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  *avg_ms = ms / reps;
  if (bytes_per_launch) *bytes_per_launch = (double)B * 2.0 * D.n_audio_ctx * D.n_text_state * 2.0;
  return WB_OK;
}

// ---- operator-level entry points for the parity tests ---------------------------------------------------------------------------
int wb_op_gemm(wb_handle* h, const void* A_f16, const void* W_f16, const float* bias, const float* residual, int32_t M, int32_t N,
               int32_t K, int32_t gelu, void* C, int32_t c_is_f32) {
  if (!h || !A_f16 || !W_f16 || !C) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  return plain_gemm(h, (const __half*)A_f16, M, (const __half*)W_f16, N, K, bias, gelu, residual, c_is_f32 ? nullptr : (__half*)C,
                    c_is_f32 ? (float*)C : nullptr);
}
int wb_op_layernorm(wb_handle* h, const float* x, const float* gamma, const float* beta, int32_t M, int32_t d, void* out_f16) {
  if (!h || !x || !out_f16) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  return launch_layernorm(x, gamma, beta, M, d, (__half*)out_f16, nullptr, h->stream, &h->launches);
}
int wb_op_attention(wb_handle* h, const void* qkv_f16, int32_t B, int32_t T, int32_t n_head, void* out_f16) {
  if (!h || !qkv_f16 || !out_f16) return WB_ERR_ARG;
  WB_CUDA_OK(cudaSetDevice(h->device));
  return launch_encoder_attention(h->gemm, (const __half*)qkv_f16, B, T, n_head, (__half*)out_f16, h->stream, &h->launches);
}

// ---- legacy f64 symbol (stft/src/lib.rs:110-122; bridge.h:11) --------------------------------------------------------------------
// Per-device state of the legacy path: the f64 tables (like the crate's lazy_static GENERATOR, lib.rs:11-13) and the
// scratch buffers, created once under a lock and reused by later calls (grown when a larger batch arrives).
struct LegacyState {
  std::mutex mu;
  LogmelTables<double>* tab = nullptr;
  double *audio = nullptr, *logspec = nullptr, *out = nullptr;
  unsigned long long* gmax = nullptr;
  int cap = 0;   // clips the scratch buffers hold
};
static LegacyState g_legacy[wb::kMaxDevices];

static int legacy_prepare(LegacyState& L, int B) {
  if (!L.tab) {
    std::unique_ptr<LogmelTables<double>> t(new LogmelTables<double>());
    build_logmel_tables<double>(*t);
    LogmelTables<double>* dptr = nullptr;
    WB_CUDA_OK(cudaMalloc(&dptr, sizeof(*t)));
    if (cudaMemcpy(dptr, t.get(), sizeof(*t), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(dptr);
      set_error("wb_generate_spectrogram_f64: table upload failed");
      return WB_ERR_CUDA;
    }
    L.tab = dptr;
  }
  if (B > L.cap) {
    cudaFree(L.audio), cudaFree(L.logspec), cudaFree(L.out), cudaFree(L.gmax);
    L.audio = L.logspec = L.out = nullptr, L.gmax = nullptr, L.cap = 0;
    const size_t na = (size_t)B * WB_N_SAMPLES_PADDED, no = (size_t)B * WB_N_MELS * WB_N_FRAMES;
    if (cudaMalloc(&L.audio, na * 8) != cudaSuccess || cudaMalloc(&L.logspec, no * 8) != cudaSuccess ||
        cudaMalloc(&L.out, no * 8) != cudaSuccess || cudaMalloc(&L.gmax, 8 * (size_t)B) != cudaSuccess) {
      cudaFree(L.audio), cudaFree(L.logspec), cudaFree(L.out), cudaFree(L.gmax);
      L.audio = L.logspec = L.out = nullptr, L.gmax = nullptr;
      cudaGetLastError();
      set_error("wb_generate_spectrogram_f64: cudaMalloc failed");
      return WB_ERR_NOMEM;
    }
    L.cap = B;
  }
  return 0;
}

int wb_generate_spectrogram_f64(double* audio, int32_t B, double* output) {
  if (!audio || !output || B < 1) {
    set_error("wb_generate_spectrogram_f64: bad argument");
    return WB_ERR_ARG;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return WB_ERR_CUDA;
  }
  const char* dev_env = getenv("WB_DEVICE");
  const int dev = dev_env ? atoi(dev_env) : 0;
  if (dev < 0 || dev >= n_dev || dev >= wb::kMaxDevices) {
    set_error("WB_DEVICE=%d out of range (%d visible)", dev, n_dev);
    return WB_ERR_ARG;
  }
  // the in-place reflection of the two 200-sample pads (lib.rs:34-40) is part of the contract: do it on the host buffer
  for (int b = 0; b < B; ++b) {
    double* a = audio + (size_t)b * WB_N_SAMPLES_PADDED;
    for (int i = 0; i < 200; ++i) {
      a[i] = a[400 - i];
      a[WB_N_SAMPLES + 200 + i] = a[200 + (WB_N_SAMPLES - 2) - i];
    }
  }
  int caller_dev = 0;
  const bool have_caller_dev = cudaGetDevice(&caller_dev) == cudaSuccess;   // restored on every exit path
  LegacyState& L = g_legacy[dev];
  int rc = WB_OK;
  {
    std::lock_guard<std::mutex> lock(L.mu);   // the reference's lazy_static is thread-safe; calls on one device serialise here
    do {
      if (cudaSetDevice(dev) != cudaSuccess) {
        set_error("wb_generate_spectrogram_f64: cudaSetDevice(%d) failed", dev);
        rc = WB_ERR_CUDA;
        break;
      }
      rc = legacy_prepare(L, B);
      if (rc) break;
      const size_t na = (size_t)B * WB_N_SAMPLES_PADDED, no = (size_t)B * WB_N_MELS * WB_N_FRAMES;
      if (cudaMemcpy(L.audio, audio, na * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("wb_generate_spectrogram_f64: H2D copy failed");
        rc = WB_ERR_CUDA;
        break;
      }
      rc = launch_logmel<double>(L.audio, WB_N_SAMPLES_PADDED, 200, B, L.tab, L.logspec, L.gmax, L.out, nullptr, nullptr, nullptr);
      if (rc) break;
      if (cudaMemcpy(output, L.out, no * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("wb_generate_spectrogram_f64: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = WB_ERR_CUDA;
      }
    } while (0);
  }
  if (have_caller_dev) cudaSetDevice(caller_dev);
  return rc;
}

void generate_spectrogram(double* audio, double* output) {
  const int rc = wb_generate_spectrogram_f64(audio, 1, output);
  if (rc != WB_OK) {
    fprintf(stderr, "generate_spectrogram: %s (status %d)\n", wb_last_error(), rc);
    abort();   // the reference panics across the FFI boundary (lib.rs .unwrap()); there is no status to return
  }
}

#include "checkpoint.inc"
#include "longform.inc"

}  // extern "C"
