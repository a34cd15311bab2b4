// Launcher declarations shared by model.cu / api.cu. Each launcher enqueues on `st`, returns 0 or a negative wb_status,
// and adds the number of kernels it launched to *launches.
#pragma once
#include "common.cuh"

namespace wb {

constexpr int kMaxSequences = 64;       // decoder rows per handle (chunks x beams) where the layer runs as the block kernels (tiny / base)
constexpr int kMaxSequencesWide = 40;   // ... on the skinny-GEMM path of the wider models: 5 MMA column tiles of 8

// ---- log-mel (logmel.cu) ---------------------------------------------------------------------------------------------
template <typename T>
struct LogmelTables;
template <typename T>
void build_logmel_tables(LogmelTables<T>& t);
template <typename T>
int launch_logmel(const T* audio, size_t chunk_stride, int chunk_off, int B, const LogmelTables<T>* dtab, T* logspec,
                  void* gmax, T* out, __half* melT, cudaStream_t st, int64_t* launches);
int launch_mel_transpose(const float* mel, int B, __half* melT, cudaStream_t st, int64_t* launches);
// stream mode (a whole recording followed by zeros, one global maximum): logspec [W][80][3000]; then one normalised 3000-frame
// segment starting at any frame of the stream
int launch_logmel_stream(const float* audio, long long n_samples, int W, const LogmelTables<float>* dtab, float* logspec, void* gmax,
                         cudaStream_t st, int64_t* launches);
int launch_logmel_segment(const float* logspec, const void* gmax, int W, long long frame0, float* out, __half* melT, cudaStream_t st,
                          int64_t* launches);
size_t logmel_tables_bytes_f32();
size_t logmel_tables_bytes_f64();

// ---- big GEMM (gemm.cu): tcgen05 + TMA ---------------------------------------------------------------------------------
// C[b][t][n] = act( sum_k A[b][t][k] * W[n][k] + bias[n] ) (+ residual)
//   A element (b,t,k) at a_ptr[b*a_batch_stride + t*a_row_stride + k]  (fp16; strides in elements, multiples of 8;
//     rows may overlap: the conv layers are expressed this way, k >= K reads as zero)
//   W fp16 [N][K] row-major (ldw = K)
//   C row (b,t) at row index  b*c_batch_rows + c_row_off + t  of a row-major [.][ldc] matrix
//   residual: mode 1 = fp32, same addressing as C;  mode 2 = fp32 [rows][N] indexed by t (positional table)
struct GemmDesc {
  const __half* a;
  long long a_row_stride, a_batch_stride;
  int rows, n_batch;   // rows per batch (t range), number of batches
  const __half* w;
  int N, K;
  const float* bias;   // [N] or null
  int gelu;
  int res_mode;        // 0 none, 1 same-as-C fp32, 2 per-t table
  const float* res;
  __half* c16;         // optional fp16 output
  float* c32;          // optional fp32 output
  int ldc;
  long long c_batch_rows;
  int c_row_off;
};
struct GemmContext;   // tensor-map cache + driver entry point
GemmContext* gemm_context_create();
void gemm_context_destroy(GemmContext*);
int launch_gemm(GemmContext* ctx, const GemmDesc& d, cudaStream_t st, int64_t* launches);
// Cached CUtensorMap (128 bytes, written to out128) of an fp16 tensor (k, rows, batch) with element strides
// (1, row_stride, batch_stride), box (64, box_rows, 1), 128-byte swizzle.
int gemm_get_tmap(GemmContext* ctx, const void* ptr, long long K, long long rows, long long nb, long long row_stride,
                  long long batch_stride, int box_rows, void* out128);

// ---- row-wise ops (rowops.cu) ------------------------------------------------------------------------------------------
int launch_layernorm(const float* x, const float* gamma, const float* beta, int M, int d, __half* out16, float* out32,
                     cudaStream_t st, int64_t* launches);
int launch_f32_to_f16(const float* in, __half* out, size_t n, cudaStream_t st, int64_t* launches);
int launch_f16_to_f32(const __half* in, float* out, size_t n, cudaStream_t st, int64_t* launches);
int launch_fill_random(void* ptr, size_t n, int is_f16, float scale, float offset, uint64_t seed, cudaStream_t st,
                       int64_t* launches);
int launch_checksum64(const void* ptr, size_t bytes, unsigned long long* acc, cudaStream_t st, int64_t* launches);

// ---- encoder attention (attention_enc.cu) ---------------------------------------------------------------------------------
// qkv fp16 [B*T][3d] (q | k | v, head h at columns h*64) -> out fp16 [B*T][d]; softmax(q k^T / 8) v, non-causal
// tmaps != null: tcgen05 kernel (TMA-fed, S and PV accumulators in TMEM); null: the mma.sync kernel
int launch_encoder_attention(GemmContext* tmaps, const __half* qkv, int B, int T, int n_head, __half* out, cudaStream_t st, int64_t* launches);

// ---- decoder step (decoder.cu) ---------------------------------------------------------------------------------------------
struct DecodeState {      // lives in device memory; read by every kernel of a step (CUDA-graph friendly)
  int cur_len;            // index of the token the current step consumes (-1 before the first embed)
  int arrive;             // arrival counter of step_finish_kernel's CTAs
  int trace_n, pad1;      // development tracing (WB_TRACE=1): number of records written
  unsigned long long* trace;   // [n][4] = {kernel id, %globaltimer at entry, at exit of block 0, 0}; null = off
};

// Input transform of a skinny GEMM (how the [Mb][K] fp16 activation tile in shared memory is produced)
enum SkinnyIn { SKINNY_IN_F16 = 0, SKINNY_IN_LN = 1 };
// Output transform
enum SkinnyOut { SKINNY_OUT_F16 = 0, SKINNY_OUT_F32 = 1, SKINNY_OUT_RESID = 2, SKINNY_OUT_QKV = 3, SKINNY_OUT_LOGITS = 4 };

struct SkinnyDesc {
  int Mb, N, K;
  const __half* w;        // [N][K]
  const float* bias;      // [N] or null
  int gelu;
  // input
  int in_mode;
  const void* in;         // F16: half [Mb][K]; LN: float [Mb][K]
  const float* ln_g;      // LN
  const float* ln_b;
  // output
  int out_mode;
  void* out;              // F16: half [Mb][N]; F32 / RESID: float [Mb][N] (RESID: +=); LOGITS: float [Mb][N] or null
  // QKV scatter: n < d -> q32[b][n];  d <= n < 2d -> kcache[(b*n_ctx + pos)*d + n-d];  else vcache;  pos = state->cur_len
  float* q32;
  __half* kcache;
  __half* vcache;
  int n_ctx;
  // LOGITS: filters + per-CTA (max, argmax, sum-exp) partials [Mb][n_cta][4]
  const unsigned char* mask;   // [N]: 1 = suppressed, 2 = suppressed at the first sampled position only; may be null
  int n_initial;
  float* part_logits;
  // LOGITS, timestamp rules (upstream ApplyTimestampRules; null = off): per-sequence rule state written by the finish kernel
  // {x: bit0 = timestamps forbidden, bit1 = text below eot forbidden; y: timestamps below y forbidden}; rows >= ts_begin
  // are timestamps; at the first sampled position only timestamps <= ts_last_allowed are allowed. The group that straddles
  // ts_begin reports its text rows in its partial and its timestamp rows in part_extra[b].
  const int4* ts_state;
  int ts_begin, ts_last_allowed, eot;
  float* part_extra;           // [Mb][4]
  GemmContext* tmaps;          // LOGITS: tensor-map cache (tcgen05 path); null selects the mma.sync kernel
  const DecodeState* state;
};
int launch_skinny_gemm(const SkinnyDesc& d, cudaStream_t st, int64_t* launches);
int skinny_logits_ctas(int N);   // upper bound on the row groups (= partials per sequence) of the LOGITS mode
int logits_groups(int Mb, int N, int K);   // actual number of row groups for this shape

// one query per (sequence, head) against rows [0, n_rows) of K/V [B][n_ctx][d] fp16 -> out16 [Mb][d] fp16
struct AttnDecodeDesc {
  int Mb, d, n_head;
  const float* q;         // [Mb][d] fp32, or null with the fused query projection below (cross attention)
  const float* x;         // residual stream [Mb][d]: q = LayerNorm(x; ln_g, ln_b) wq^T + bq, computed per (sequence, head)
  const float* ln_g;
  const float* ln_b;
  const __half* wq;       // [d][d]
  const float* bq;
  const __half* k;        // [Mb / kv_share][n_ctx][d]
  const __half* v;
  int n_ctx;              // allocated rows per sequence
  int n_rows_fixed;       // >0: fixed row count (cross attention, 1500); 0: rows = state->cur_len + 1 (self attention)
  int kv_share;           // sequences per K/V slab (beam search: beams of one chunk share the cross K/V); >= 1
  const DecodeState* state;
  __half* out16;          // [Mb][d]
  int pdl_late_ok;        // the successor is a block kernel that gains nothing from starting before this one's main loop ends
  GemmContext* tmaps;     // tensor-map cache
};
int launch_attn_decode(const AttnDecodeDesc& d, cudaStream_t st, int64_t* launches);

// whole self-attention block of a decoder layer in one kernel (n_head <= 8; one cluster per group of 4 sequences):
// x += Wo attn(LN(x) Wqkv^T + bqkv against the self-attention cache, new k/v appended at state->cur_len) + bo
struct SelfBlockDesc {
  int Mb, d, n_head, n_ctx;
  float* x;               // [Mb][d] residual stream, updated in place
  const float* ln_g;
  const float* ln_b;
  const __half* wqkv;     // [3d][d] (q | k | v)
  const float* bqkv;      // [3d]
  const __half* wo;       // [d][d]
  const float* bo;        // [d]
  __half* kcache;         // [Mb][n_ctx][d]
  __half* vcache;
  const DecodeState* state;
};
int self_block_supported(int n_head, int d);
int launch_self_block(const SelfBlockDesc& d, cudaStream_t st, int64_t* launches);

// everything after cross attention in one cluster kernel (d = 384 / 512, 8 sequences per cluster):
// x' = x + a16 wo^T + bo;  x = x' + w2 gelu(w1 LN(x') + b1) + b2
struct PostBlockDesc {
  int Mb, d, n_head;
  float* x;               // [Mb][d] residual stream, updated in place
  const __half* a16;      // [Mb][d] cross-attention outputs
  const __half* wo;
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  const __half* w1;       // [4d][d]
  const float* b1;
  const __half* w2;       // [d][4d]
  const float* b2;
  const DecodeState* state;
};
int post_block_supported(int n_head, int d);
int launch_post_block(const PostBlockDesc& d, cudaStream_t st, int64_t* launches);

struct FinishDesc {
  int Mb, V, d, n_ctx;
  int sample;                 // 0: only embed the next (already present) token and advance
  const float* part_logits;   // [Mb][n_part][4]
  int n_part;
  int eot;
  int32_t* tokens;            // [Mb][tokens_ld]
  int tokens_ld;
  float* sum_logprob;         // [Mb]
  int32_t* done;              // [Mb]: 1 if the sequence's newest token is eot
  const __half* tok_emb;
  const float* pos_emb;
  float* x;                   // [Mb][d] residual stream of the next step
  // timestamp rules (null = off): partial groups >= ts_group0 (and part_extra) are timestamp rows; the probability-mass rule
  // and the rule state for the next step are evaluated here
  int4* ts_state;
  const float* part_extra;
  int ts_begin, ts_group0, n_initial;
  // temperature > 0: the tokens drawn by sample_rows_kernel and their log-probabilities [Mb] (null: argmax of the partials)
  const int32_t* chosen;
  const float* chosen_logprob;
  DecodeState* state;
};
int launch_step_finish(const FinishDesc& d, cudaStream_t st, int64_t* launches);
// Gumbel-max draw from softmax(logits / temperature) per sequence over the stored filtered logits [Mb][V] (timestamp mass rule
// applied when ts_begin < V), with log_softmax(logits)[token]; and softmax(logits)[token] for one fixed token
int launch_sample_rows(const float* logits, int Mb, int V, int ts_begin, float temperature, unsigned long long seed, const DecodeState* state,
                       int32_t* chosen, float* chosen_logprob, cudaStream_t st, int64_t* launches);
int launch_set_cur_len(DecodeState* state, int v, cudaStream_t st, int64_t* launches);
int launch_row_token_prob(const float* logits, int Mb, int V, int token, float* prob, cudaStream_t st, int64_t* launches);
// L2 residency hints of the decode kernels on the current device: 1 = weights evict_last, cross K/V stream evict_first
int decoder_set_l2_mode(int mode);
int launch_delay(unsigned long long ns, cudaStream_t st, int64_t* launches);
// LayerNorm of the Mb rows of the residual stream into fp16 [Mb][d], once per use instead of once per CTA of the following
// skinny GEMM (wider models: a row of d = 768 .. 1280 floats is two or three dependent load batches per thread there)
int launch_ln_rows(const float* x, const float* gamma, const float* beta, int Mb, int d, __half* out16, const DecodeState* state,
                   cudaStream_t st, int64_t* launches);

// beam search support: per-row top-k (k <= 8) of the filtered logits with their log-softmax values, and the re-indexing
// of the self-attention K/V cache by source beam (upstream rearrange_kv_cache)
// ts_begin < V: the last rule of upstream ApplyTimestampRules first - rows below ts_begin (text) are dropped when the
// probability mass over the rows from ts_begin on (timestamps) exceeds every text row
int launch_topk_logprobs(const float* logits, int Mb, int V, int k, int ts_begin, float* top_logprob /*[Mb][8]*/, int32_t* top_index /*[Mb][8]*/,
                         cudaStream_t st, int64_t* launches);
int launch_reorder_kv(const __half* const* src_k, const __half* const* src_v, __half* const* dst_k, __half* const* dst_v, int n_layer,
                      int Mb, int n_ctx, int d, const int32_t* source, const DecodeState* state, cudaStream_t st, int64_t* launches);

}  // namespace wb
