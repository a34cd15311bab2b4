/*
 * whisper_b200.h — C ABI of libwhisper_b200.so: the B200 (sm_100a) replacement for the hot path of
 * tanmayb123/OpenAI-Whisper-CoreML:  audio -> log-mel -> encoder -> decoder -> token IDs.
 *
 * Every entry point cites the reference interface it replaces. Plain pointers and sizes only; no C++ or torch
 * types. Functions return 0 on success and a negative wb_status on failure (never throw, never abort);
 * wb_last_error() returns a thread-local message for the last failure. A handle owns all device memory of one
 * model instance on one GPU; distinct handles may be used from distinct threads, one handle may not.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point returns WB_ERR_CUDA.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_N_SAMPLES_PER_CHUNK 480000 /* 16000 * 30           stft/src/lib.rs:37,112 ; ContentView.swift:57-60 */
#define WB_N_SAMPLES_PADDED 480400    /* + 200 + 200          stft.swift:10-11 ; lib.rs:112                    */
#define WB_N_MELS_C 80                /*                      lib.rs:60 ; whisper_to_cml.py:13                 */
#define WB_N_FRAMES_C 3000            /*                      lib.rs:52,62 ; whisper_to_cml.py:13              */

typedef enum wb_status {
  WB_OK = 0,
  WB_ERR_ARG = -1,     /* bad argument (null pointer, size out of range, unknown weight name, ...) */
  WB_ERR_CUDA = -2,    /* CUDA runtime / driver error, or no device                                  */
  WB_ERR_STATE = -3,   /* call sequence error (weights not loaded, encode not run before decode ...) */
  WB_ERR_NOMEM = -4
} wb_status;

/* Upstream `ModelDimensions`; the reference fixes n_mels=80, n_audio_ctx=1500 (whisper_to_cml.py:13,29) and exports
 * "small" (d=768, :7). d_head is 64 for every Whisper size and is required here (state % head == 0, state/head == 64). */
typedef struct wb_dims {
  int32_t n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int32_t n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer;
} wb_dims;

/* Greedy / beam decoding options (north-star extension; semantics of upstream whisper/decoding.py DecodingTask,
 * see SURVEY.md §8c). Token lists are passed as data because the tokenizer vocabulary is not part of the model. */
typedef struct wb_decode_opts {
  const int32_t* initial_tokens; /* sot sequence, e.g. {50257,50362} (.en) or {50258,50259,50359,50363}           */
  int32_t n_initial;
  int32_t sample_len;            /* max sampled tokens; upstream default n_text_ctx/2 = 224                         */
  int32_t eot;                   /* 50256 (.en) / 50257 (multilingual)                                              */
  const int32_t* suppress;       /* SuppressTokens: -inf at every step                                              */
  int32_t n_suppress;
  const int32_t* suppress_begin; /* SuppressBlank: -inf at the first sampled position only                          */
  int32_t n_suppress_begin;
  int32_t beam_size;             /* 0 or 1 = greedy; 2..7 = beam search, patience 1 (needs max_beams >= beam_size)  */
  int32_t eot_check_interval;    /* how often (in steps) the host polls "all sequences ended"; 0 = default (8)      */
  /* upstream ApplyTimestampRules (DecodingOptions.without_timestamps = False, upstream's own default)                */
  int32_t timestamps;            /* 0 = off (the prompt then ends with <|notimestamps|>)                            */
  int32_t timestamp_begin;       /* first timestamp token <|0.00|>: 50363 (.en) / 50364 (multilingual)              */
  int32_t no_timestamps;         /* <|notimestamps|>: 50362 / 50363; never sampled when the rules are on            */
  int32_t max_initial_timestamp_index; /* first timestamp <= this many 0.02 s steps (upstream: 50); < 0 = no limit  */
  /* upstream DecodingOptions.temperature / best_of (greedy path only). The draw is defined as Gumbel-max with a counter-
   * based generator (see wb_transcribe_long), so that a CPU restatement sees the same noise: reproducible per seed.      */
  float temperature;             /* 0 = arg-max; > 0: one draw from softmax(logits / temperature) per step            */
  int32_t best_of;               /* temperature > 0: independent samples per chunk (<= max_beams, batch * best_of <=
                                    the handle's sequence capacity); the best sum_logprob / length is returned; 0 = 1  */
  uint64_t seed;                 /* generator key of this call                                                        */
  /* upstream DecodingTask no_speech_probs: softmax of the UNFILTERED logits after <|startoftranscript|>.                */
  int32_t no_speech;             /* <|nospeech|> token: 50361 (.en) / 50362 (multilingual)                            */
  int32_t sot_index;             /* index of <|startoftranscript|> in initial_tokens (> 0 with a prompt)              */
  float* no_speech_prob;         /* out [B], or null (then no_speech / sot_index are ignored)                         */
} wb_decode_opts;

typedef struct wb_handle wb_handle;

/* ---- legacy symbol -----------------------------------------------------------------------------------------------
 * Replaces `#[no_mangle] pub extern fn generate_spectrogram(audio: *mut f64, output: *mut f64)`
 * (stft/src/lib.rs:110-122), declared to Swift at Whisper/Whisper/bridge.h:11 and called at stft.swift:15.
 * Same contract: audio[480400] in/out — the two 200-sample pads are overwritten with the torch-style reflection
 * (lib.rs:34-40); output[240000] = [80][3000] normalised log-mel, out[i*3000+j]. Computed in f64 on GPU 0 (device
 * overridable with env WB_DEVICE) by the templated fused kernel. On a CUDA failure the process aborts with a
 * message, mirroring the Rust panic-across-FFI behaviour of the reference (there is no status to return). */
void generate_spectrogram(double* audio, double* output);

/* Same computation with a status code instead of abort, B clips at once: audio [B][480400] f64 in/out, output
 * [B][80][3000] f64. Host pointers. */
int wb_generate_spectrogram_f64(double* audio, int32_t B, double* output);

const char* wb_last_error(void);
int wb_version(void);

/* ---- model instance -----------------------------------------------------------------------------------------------
 * Replaces `Whisper.init()` (Whisper/Whisper/Whisper.swift:17-21: loads encoder.mlpackage and decoder.mlpackage).
 * `stream` is a cudaStream_t (NULL = a private non-blocking stream created by the handle); all work of the handle is
 * enqueued there. max_batch = maximum number of 30 s chunks per call; max_beams >= 1. */
int wb_create(const wb_dims* dims, int32_t max_batch, int32_t max_beams, int32_t device, void* stream, wb_handle** out);
int wb_destroy(wb_handle* h);
int wb_get_dims(const wb_handle* h, wb_dims* out);

/* Weights. Replaces `whisper.load_model("small")` + the two ct.convert exports (whisper_to_cml.py:6-8,10-43) which
 * bake the checkpoint into the .mlpackages. Tensors are addressed by their upstream openai-whisper state-dict names
 * ("encoder.blocks.0.attn.query.weight", "decoder.token_embedding.weight", ...), passed as host fp32 and stored on
 * the device as fp16 (matrices, embeddings) or fp32 (biases, LayerNorm, encoder positional table) in the layouts the
 * kernels want. wb_weights_commit() checks that every tensor was set. */
int wb_set_weight(wb_handle* h, const char* name, const float* data, size_t numel);
int wb_weights_commit(wb_handle* h);
/* Seeded synthetic weights generated on the device (benchmarks without a checkpoint). */
int wb_init_random_weights(wb_handle* h, uint64_t seed);
/* The packed device arena holding every weight (for a load-time NCCL broadcast from rank 0). */
int wb_weight_arena(wb_handle* h, void** device_ptr, size_t* bytes);
/* Mark weights as present after the arena was filled externally (e.g. by the broadcast). */
int wb_weights_mark_loaded(wb_handle* h);
/* Read a tensor back as host fp32 in its upstream layout (the inverse of wb_set_weight; fp16-stored tensors return their
 * rounded values). Makes device-generated weights (wb_init_random_weights) visible to a checker. */
int wb_get_weight(wb_handle* h, const char* name, float* data, size_t numel);
/* Enumerate the tensors of the model: count, then (name, element count) by index in sorted-name order. */
int wb_weight_count(const wb_handle* h);
int wb_weight_info(const wb_handle* h, int32_t index, char* name_out, size_t name_cap, size_t* numel);
/* Position-weighted 64-bit checksum of the packed arena, computed on the device (ranks compare it after the broadcast). */
int wb_weights_checksum(wb_handle* h, uint64_t* out);

/* ---- log-mel ------------------------------------------------------------------------------------------------------
 * Replaces `generateSpectrogram(audio:)` (Whisper/Whisper/stft.swift:8-19) for B clips of f32 PCM: audio [B][480000]
 * (the unpadded clip; the wrapper's 200+200 zero pad and the crate's reflection are folded into the kernel's index
 * map), out [B][80][3000] f32. `*_dev` variants take device pointers and only enqueue work on the handle's stream. */
int wb_logmel(wb_handle* h, const float* audio, int32_t B, float* out);
int wb_logmel_dev(wb_handle* h, const float* audio_dev, int32_t B, float* out_dev);

/* ---- encoder ------------------------------------------------------------------------------------------------------
 * wb_encode replaces `Whisper.encode(audio:)` (Whisper.swift:23-31): log-mel, then the encoder forward
 * (`encoderModel.prediction(x_1:).var_1385`, exported at whisper_to_cml.py:10-23). audio [B][480000] f32 host;
 * xa_out [B][1500][d] f32 host, or NULL to keep the features on the device only. The features and the
 * cross-attention K/V derived from them stay resident in the handle for the decode calls that follow.
 * wb_encode_mel replaces `encoderModel.prediction(x_1:)` alone: mel [B][80][3000] f32 host (already normalised). */
int wb_encode(wb_handle* h, const float* audio, int32_t B, float* xa_out);
int wb_encode_mel(wb_handle* h, const float* mel, int32_t B, float* xa_out);
int wb_encode_dev(wb_handle* h, const float* audio_dev, int32_t B);
/* Load externally computed audio features (decoder.prediction's `xa` argument, Whisper.swift:36): xa [B][1500][d]. */
int wb_set_audio_features(wb_handle* h, const float* xa, int32_t B);
/* The resident features of the last wb_encode* / wb_transcribe* call (`audioFeatures`, Whisper.swift:30): xa_out [B][1500][d]. */
int wb_get_audio_features(wb_handle* h, float* xa_out, int32_t B);

/* ---- decoder ------------------------------------------------------------------------------------------------------
 * wb_decoder_logits replaces `decoderModel.prediction(x_1: tokens, xa: audioFeatures).var_2217`
 * (Whisper.swift:36; exported at whisper_to_cml.py:25-43 with tokens (1,1)): teacher-forced logits of tokens [B][t]
 * against the resident audio features; logits [B][t][n_vocab] f32 host. (The reference passes the token as a float32
 * MLMultiArray, Whisper.swift:34-35; wb_decoder_logits_f32tok accepts that form.) */
int wb_decoder_logits(wb_handle* h, const int32_t* tokens, int32_t B, int32_t t, float* logits);
int wb_decoder_logits_f32tok(wb_handle* h, const float* tokens, int32_t B, int32_t t, float* logits);

/* Replaces `Whisper.decode(audioFeatures:)` (Whisper.swift:33-40): one decoder call on [sot], arg-max over the 99
 * language logits [lang0, lang0+99) with Swift `max(by:)` tie-breaking (the first maximal element: `max(by:)` replaces
 * its running result only on a strict increase, so NaN never wins either). sot/lang0 default to
 * 50258/50259 (Whisper.swift:35,37) when passed as 0. lang_idx [B] in 0..98. */
int wb_detect_language(wb_handle* h, int32_t B, int32_t sot, int32_t lang0, int32_t* lang_idx);

/* Greedy / beam decode of the resident features. tokens_out [B][n_initial + sample_len] (padded with eot), lens [B]
 * (initial tokens included, first eot included), sum_logprob [B]. Beam search returns, per chunk, the candidate with the
 * best sum_logprob / length (upstream MaximumLikelihoodRanker without length penalty). */
int wb_decode(wb_handle* h, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out, int32_t* lens, float* sum_logprob);
/* audio -> tokens in one call (= wb_encode + wb_decode). Host pointers. The batch is copied in slabs on a copy stream of the
 * handle (the first quarter, then the rest while the first slab is encoded; env WB_H2D_SLABS), so pinned host memory hides
 * most of the transfer; results do not depend on the split. */
int wb_transcribe(wb_handle* h, const float* audio, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out,
                  int32_t* lens, float* sum_logprob);
/* Same with the audio already on the device (throughput path; device pointer, results to host). */
int wb_transcribe_dev(wb_handle* h, const float* audio_dev, int32_t B, const wb_decode_opts* opts, int32_t* tokens_out,
                      int32_t* lens, float* sum_logprob);

/* ---- checkpoint files ------------------------------------------------------------------------
 * The reference bakes real weights into its .mlmodels at export time (whisper_to_cml.py:7 `whisper.load_model("small")`); here a
 * handle is filled from a safetensors file: F32 / F16 / BF16 tensors under upstream openai-whisper names or transformers names
 * (`model.decoder.layers.0.self_attn.q_proj.weight`; proj_out.weight and k_proj.bias have no upstream counterpart and are
 * ignored). wb_safetensors_read_dims derives the ten model dimensions from the tensor shapes (host only, no GPU needed), so
 * that wb_create can be called from the file alone. wb_load_safetensors sets every weight and commits; a missing tensor is an
 * error, except the encoder's sinusoidal positions, which are regenerated. */
int wb_safetensors_read_dims(const char* path, wb_dims* out);
int wb_load_safetensors(wb_handle* h, const char* path, int32_t* n_loaded);

/* ---- long-form transcription ---------------------------------------------------------------
 * The reference transcribes exactly one fixed 30 s window (Whisper/Whisper/ContentView.swift:57-62). For longer recordings
 * "transcribe" means upstream openai-whisper `transcribe()` (whisper/transcribe.py), restated here: log-mel of the whole
 * recording followed by 30 s of zeros with one global maximum; a 3000-frame window at `seek`; decode_with_fallback over a
 * temperature schedule (compression-ratio and average-log-probability thresholds, no-speech override); the no-speech skip;
 * segments cut at consecutive timestamp tokens; seek advanced to the last timestamp; the previous text as the next prompt
 * unless the window needed a temperature above 0.5. Not restated: word_timestamps, clip_timestamps,
 * hallucination_silence_threshold. One recording per call (the loop is sequential and data dependent); recordings shard
 * across handles / GPUs. A window's prompt + sampled tokens are capped at n_text_ctx (upstream: n_text_ctx + 1).
 *
 * Token table for the compression-ratio rule and the empty-text test (the vocabulary is not part of the model): token id
 * -> bytes, ids beyond the table (special tokens) decode to nothing. */
typedef struct wb_tokenizer wb_tokenizer;
wb_tokenizer* wb_tokenizer_create(const uint8_t* blob, const uint32_t* offsets /* [n_tokens + 1] */, int32_t n_tokens);
/* upstream whisper/assets/{gpt2,multilingual}.tiktoken: lines of "base64(token bytes) rank" */
wb_tokenizer* wb_tokenizer_load_tiktoken(const char* path);
void wb_tokenizer_destroy(wb_tokenizer* t);
int32_t wb_tokenizer_size(const wb_tokenizer* t);
/* Upstream Tokenizer.decode before the UTF-8 step: concatenated bytes of the ids below drop_from (pass timestamp_begin to drop
 * timestamps and special tokens). *len = total bytes; at most cap of them are written to out (out may be null). */
int wb_tokenizer_decode(const wb_tokenizer* t, const int32_t* ids, int32_t n, int32_t drop_from, uint8_t* out, size_t cap, size_t* len);
/* upstream whisper/utils.py compression_ratio of `raw.decode("utf-8", errors="replace").strip()`:
 * len(utf8) / len(zlib.compress(utf8)); *stripped_len (optional) = len(utf8). */
int wb_text_compression_ratio(const uint8_t* raw, size_t n, float* ratio, size_t* stripped_len);

typedef struct wb_long_opts {
  wb_decode_opts decode;           /* per-window options: initial_tokens = the sot sequence WITHOUT prompt, sot_index into it,
                                      no_speech, timestamps (upstream default: on), suppress lists, sample_len (0 = n_text_ctx/2),
                                      beam_size (temperature 0 only; no no_speech_prob then), best_of (temperature > 0), seed;
                                      temperature and no_speech_prob are set by the loop                                     */
  const float* temperatures;       /* fallback schedule; null = {0, 0.2, 0.4, 0.6, 0.8, 1.0}                                 */
  int32_t n_temperatures;
  float compression_ratio_threshold; /* upstream 2.4; NaN = rule off (also off without a tokenizer)                           */
  float logprob_threshold;         /* upstream -1.0; NaN = off                                                               */
  float no_speech_threshold;       /* upstream 0.6; NaN = off                                                                */
  int32_t condition_on_previous_text; /* upstream True                                                                      */
  const int32_t* initial_prompt;   /* tokenised initial_prompt (upstream: encode(" " + prompt.strip())), or null              */
  int32_t n_initial_prompt;
  int32_t sot_prev;                /* <|startofprev|>: 50360 (.en) / 50361 (multilingual)                                    */
  const wb_tokenizer* tokenizer;   /* may be null                                                                            */
  int32_t detect_language;         /* 1: arg-max language of the first window replaces initial_tokens[sot_index + 1]         */
  int32_t lang0;                   /* first language token (50259), used with detect_language                                */
  int32_t* detected_language;      /* out (optional): language index 0..98, or -1                                            */
} wb_long_opts;

typedef struct wb_segment {
  int32_t seek;                    /* mel frame the window started at                                                        */
  float start, end;                /* seconds                                                                                */
  int32_t token_begin, n_tokens;   /* range in the returned token stream (timestamp tokens included); n_tokens = 0 for a
                                      cleared segment (instantaneous or without text)                                        */
  float temperature, avg_logprob, compression_ratio, no_speech_prob;   /* of the window's accepted decode                   */
} wb_segment;

/* pcm: n_samples of 16 kHz mono f32 (host). Fails with WB_ERR_ARG (and the needed counts in n_segments / n_tokens) when the
 * output buffers are too small. */
int wb_transcribe_long(wb_handle* h, const float* pcm, int64_t n_samples, const wb_long_opts* opts, wb_segment* segments,
                       int32_t segment_cap, int32_t* n_segments, int32_t* tokens, int32_t token_cap, int32_t* n_tokens);
/* The loop's front end alone: upstream log_mel_spectrogram(pcm, padding = 30 s of zeros)[:, frame0 : frame0 + 3000], one global
 * maximum, reflection only at the start of the recording -> out [80][3000] f32 host (frames past the padded end are 0). */
int wb_logmel_long(wb_handle* h, const float* pcm, int64_t n_samples, int64_t frame0, float* out);
/* generator key the loop passes to wb_decode for window `seek` and schedule index i (for restatements / tests) */
uint64_t wb_call_seed(uint64_t seed, int64_t seek, int32_t temperature_index);

/* ---- introspection for tests / benchmarks ----------------------------------------------------------------------------- */
/* Number of kernels this library launched on the handle since creation (graph replays count their nodes). */
int64_t wb_launch_count(const wb_handle* h);
/* Device time (ms) of the last call's phases, measured with CUDA events on the handle's stream:
 * out[0]=logmel, out[1]=encoder (incl. cross K/V), out[2]=decode loop, out[3]=number of decode steps run. After wb_transcribe
 * (host audio, slabs) out[0] is the first slab's copy + log-mel and out[1] everything up to the last slab's cross K/V. */
int wb_last_timings(const wb_handle* h, float out[4]);
int wb_sync(wb_handle* h);
/* Times the decoder's KV-cache attention kernel alone on the resident cross-attention K/V of B chunks: `reps` launches
 * cycling over the decoder layers (so successive launches stream different HBM), CUDA events on the handle's stream.
 * avg_ms = mean device time per launch; bytes_per_launch = algorithmic K+V bytes one launch reads (B*2*1500*d*2). */
int wb_profile_cross_attention(wb_handle* h, int32_t B, int32_t reps, float* avg_ms, double* bytes_per_launch);

/* Low-level operator entry points used by the parity tests (device pointers, enqueue on the handle's stream).
 * C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual); A,W fp16; fp32 accumulate; C fp16 or fp32. */
int wb_op_gemm(wb_handle* h, const void* A_f16, const void* W_f16, const float* bias, const float* residual, int32_t M,
               int32_t N, int32_t K, int32_t gelu, void* C, int32_t c_is_f32);
int wb_op_layernorm(wb_handle* h, const float* x, const float* gamma, const float* beta, int32_t M, int32_t d,
                    void* out_f16);
/* Encoder self-attention on fused qkv [B*T][3d] fp16 -> out [B*T][d] fp16 (non-causal). */
int wb_op_attention(wb_handle* h, const void* qkv_f16, int32_t B, int32_t T, int32_t n_head, void* out_f16);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
