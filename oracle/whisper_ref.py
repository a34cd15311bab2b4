"""
oracle/whisper_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PyTorch fp32 CPU restatement of the encoder/decoder arithmetic the reference exports to CoreML
(/root/reference/whisper_to_cml.py:6-8,10-23,25-43: `whisper.load_model("small")`, `model.encoder` traced on
(1,80,3000), `model.decoder` traced on (tokens(1,1), audio(1,1500,768))) and of the language-ID decode the
reference's Swift performs on it (/root/reference/Whisper/Whisper/Whisper.swift:33-40).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.

PARITY UNPINNED. The arithmetic itself lives in the PyPI package `openai-whisper` (whisper/model.py,
whisper/decoding.py), which is imported by whisper_to_cml.py:1 but is neither vendored in /root/reference nor
version-pinned there (no requirements file; the Swift headers are dated 2022-09-26, i.e. the first public release),
and is not installed in this image. The reference holds no test or golden output for any encoder/decoder value.
This file therefore restates the *published* upstream algorithm:
  sinusoids()                  sin/cos table, log-timescale increment ln(10000)/(d/2-1)
  AudioEncoder.forward         gelu(conv1 k3 p1) -> gelu(conv2 k3 s2 p1) -> +sinusoids -> L pre-LN blocks -> ln_post
  MultiHeadAttention           query/value/out with bias, key without; q and k each scaled by d_head**-0.25;
                               softmax in fp32; additive -inf causal mask for decoder self-attention
  ResidualAttentionBlock       x += attn(attn_ln(x)); [x += cross_attn(cross_attn_ln(x), xa)]; x += mlp(mlp_ln(x))
  TextDecoder.forward          token_embedding[x] + positional_embedding[offset:offset+t]; blocks; ln;
                               logits = x @ token_embedding.T  (fp32)
  DecodingTask (greedy, t=0)   SuppressBlank at the first sampled position, SuppressTokens, argmax, EOT forcing,
                               sum_logprobs += logprob * (last != eot), stop when every sequence ended or after
                               sample_len steps
  detect_language              argmax over the language-token logits after one decoder call on [sot]
and is cross-checked, on identical seeded weights, against the independent implementation that IS in this image,
`transformers.WhisperForConditionalGeneration` (tests/test_oracle_whisper.py, tools/gen_golden.py): encoder
output, teacher-forced logits and greedy tokens agree to fp32 round-off. Token-ID layout for the language-ID path
is pinned by the reference itself: sot = 50258, language logits 50259...50357 (Whisper.swift:35,37).

State-dict keys follow upstream openai-whisper names (encoder.blocks.N.attn.query.weight, decoder.ln.weight, ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class ModelDims:
    """Upstream `ModelDimensions`. n_audio_ctx=1500 and n_mels=80 are pinned by whisper_to_cml.py:13,29."""
    n_mels: int = 80
    n_audio_ctx: int = 1500
    n_audio_state: int = 384
    n_audio_head: int = 6
    n_audio_layer: int = 4
    n_vocab: int = 51864
    n_text_ctx: int = 448
    n_text_state: int = 384
    n_text_head: int = 6
    n_text_layer: int = 4

    @property
    def is_multilingual(self) -> bool:
        return self.n_vocab == 51865


DIMS = {
    "tiny.en": ModelDims(80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4),
    "tiny": ModelDims(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base.en": ModelDims(80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6),
    "base": ModelDims(80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small.en": ModelDims(80, 1500, 768, 12, 12, 51864, 448, 768, 12, 12),
    "small": ModelDims(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium": ModelDims(80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v2": ModelDims(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
}


@dataclass(frozen=True)
class Vocab:
    """Special-token layout (upstream tokenizer.py). Multilingual values match Whisper.swift:35,37."""
    eot: int
    sot: int
    lang0: int          # first language token; 99 languages follow
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int

    @staticmethod
    def for_dims(dims: ModelDims) -> "Vocab":
        if dims.is_multilingual:
            return Vocab(50257, 50258, 50259, 50358, 50359, 50360, 50361, 50362, 50363, 50364)
        return Vocab(50256, 50257, 50258, 50357, 50358, 50359, 50360, 50361, 50362, 50363)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    assert channels % 2 == 0
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def weight_shapes(dims: ModelDims) -> Dict[str, tuple]:
    """Every tensor of the upstream state dict, in a fixed order (also the order of the packed weight file)."""
    d, dt = dims.n_audio_state, dims.n_text_state
    s: Dict[str, tuple] = {}
    s["encoder.conv1.weight"] = (d, dims.n_mels, 3)
    s["encoder.conv1.bias"] = (d,)
    s["encoder.conv2.weight"] = (d, d, 3)
    s["encoder.conv2.bias"] = (d,)
    s["encoder.positional_embedding"] = (dims.n_audio_ctx, d)

    def block(prefix: str, n: int, cross: bool):
        for nm in (["attn"] + (["cross_attn"] if cross else [])):
            s[f"{prefix}.{nm}.query.weight"] = (n, n)
            s[f"{prefix}.{nm}.query.bias"] = (n,)
            s[f"{prefix}.{nm}.key.weight"] = (n, n)
            s[f"{prefix}.{nm}.value.weight"] = (n, n)
            s[f"{prefix}.{nm}.value.bias"] = (n,)
            s[f"{prefix}.{nm}.out.weight"] = (n, n)
            s[f"{prefix}.{nm}.out.bias"] = (n,)
            s[f"{prefix}.{nm}_ln.weight"] = (n,)
            s[f"{prefix}.{nm}_ln.bias"] = (n,)
        s[f"{prefix}.mlp.0.weight"] = (4 * n, n)
        s[f"{prefix}.mlp.0.bias"] = (4 * n,)
        s[f"{prefix}.mlp.2.weight"] = (n, 4 * n)
        s[f"{prefix}.mlp.2.bias"] = (n,)
        s[f"{prefix}.mlp_ln.weight"] = (n,)
        s[f"{prefix}.mlp_ln.bias"] = (n,)

    for i in range(dims.n_audio_layer):
        block(f"encoder.blocks.{i}", d, False)
    s["encoder.ln_post.weight"] = (d,)
    s["encoder.ln_post.bias"] = (d,)
    s["decoder.token_embedding.weight"] = (dims.n_vocab, dt)
    s["decoder.positional_embedding"] = (dims.n_text_ctx, dt)
    for i in range(dims.n_text_layer):
        block(f"decoder.blocks.{i}", dt, True)
    s["decoder.ln.weight"] = (dt,)
    s["decoder.ln.bias"] = (dt,)
    return s


def random_weights(dims: ModelDims, seed: int = 0, fp16_round: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint (no real checkpoints exist offline). Non-trivial biases and LN affine
    parameters so that every term of the arithmetic is exercised. With fp16_round=True every value is exactly
    representable in fp16, so an fp16-weight implementation and this fp32 oracle hold identical weights."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    for name, shape in weight_shapes(dims).items():
        if name == "encoder.positional_embedding":
            t = sinusoids(*shape)
        elif name.endswith("_ln.weight") or name.endswith("ln_post.weight") or name == "decoder.ln.weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("_ln.bias") or name.endswith("ln_post.bias") or name == "decoder.ln.bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name == "decoder.token_embedding.weight":
            t = 0.05 * torch.randn(shape, generator=g)
        elif name == "decoder.positional_embedding":
            t = 0.1 * torch.randn(shape, generator=g)
        elif "conv" in name:
            fan_in = shape[1] * shape[2]
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        elif ".query." in name or ".key." in name:
            t = torch.randn(shape, generator=g) * (2.0 / math.sqrt(shape[1]))     # peaked (trained-like) attention
        else:
            fan_in = shape[1]
            t = torch.randn(shape, generator=g) * (0.7 / math.sqrt(fan_in))
        if fp16_round:
            t = t.to(torch.float16).to(torch.float32)
        w[name] = t.contiguous()
    return w


def _ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)


class WhisperRef:
    """fp32 restatement of upstream `Whisper` (encoder + decoder) over a plain state dict."""

    def __init__(self, dims: ModelDims, weights: Dict[str, torch.Tensor]):
        self.dims = dims
        self.w = {k: v.float() for k, v in weights.items()}
        missing = set(weight_shapes(dims)) - set(self.w)
        assert not missing, f"missing weights: {sorted(missing)[:4]}"
        n = dims.n_text_ctx
        self.mask = torch.full((n, n), float("-inf")).triu_(1)
        self.vocab = Vocab.for_dims(dims)

    # ---- attention ------------------------------------------------------------------------------------------
    def _proj_kv(self, p: str, src: torch.Tensor):
        k = F.linear(src, self.w[f"{p}.key.weight"])
        v = F.linear(src, self.w[f"{p}.value.weight"], self.w[f"{p}.value.bias"])
        return k, v

    def _attend(self, p: str, n_head: int, q_in, k, v, mask=None):
        q = F.linear(q_in, self.w[f"{p}.query.weight"], self.w[f"{p}.query.bias"])
        B, T, D = q.shape
        scale = (D // n_head) ** -0.25
        qh = q.view(B, T, n_head, -1).permute(0, 2, 1, 3) * scale
        kh = k.view(B, k.shape[1], n_head, -1).permute(0, 2, 3, 1) * scale
        vh = v.view(B, v.shape[1], n_head, -1).permute(0, 2, 1, 3)
        qk = qh @ kh
        if mask is not None:
            qk = qk + mask
        a = F.softmax(qk.float(), dim=-1)
        o = (a @ vh).permute(0, 2, 1, 3).flatten(start_dim=2)
        return F.linear(o, self.w[f"{p}.out.weight"], self.w[f"{p}.out.bias"])

    def _mlp(self, p: str, x):
        h = F.gelu(F.linear(x, self.w[f"{p}.mlp.0.weight"], self.w[f"{p}.mlp.0.bias"]))
        return F.linear(h, self.w[f"{p}.mlp.2.weight"], self.w[f"{p}.mlp.2.bias"])

    # ---- encoder (AudioEncoder.forward; exported at whisper_to_cml.py:10-23) -------------------------------------
    @torch.no_grad()
    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B,80,3000] (already normalised log-mel, Whisper.swift:25-29) -> xa [B,1500,d]."""
        w, dims = self.w, self.dims
        x = F.gelu(F.conv1d(mel.float(), w["encoder.conv1.weight"], w["encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, w["encoder.conv2.weight"], w["encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1)
        x = x + w["encoder.positional_embedding"]
        for i in range(dims.n_audio_layer):
            p = f"encoder.blocks.{i}"
            h = _ln(x, w[f"{p}.attn_ln.weight"], w[f"{p}.attn_ln.bias"])
            k, v = self._proj_kv(f"{p}.attn", h)
            x = x + self._attend(f"{p}.attn", dims.n_audio_head, h, k, v)
            x = x + self._mlp(p, _ln(x, w[f"{p}.mlp_ln.weight"], w[f"{p}.mlp_ln.bias"]))
        return _ln(x, w["encoder.ln_post.weight"], w["encoder.ln_post.bias"])

    # ---- decoder (TextDecoder.forward; exported at whisper_to_cml.py:25-43) ------------------------------------
    @torch.no_grad()
    def cross_kv(self, xa: torch.Tensor):
        return [self._proj_kv(f"decoder.blocks.{i}.cross_attn", xa.float()) for i in range(self.dims.n_text_layer)]

    @torch.no_grad()
    def decoder_logits(self, tokens: torch.Tensor, xa: torch.Tensor, cross=None, self_cache=None) -> torch.Tensor:
        """tokens [B,t] int64, xa [B,1500,d] -> logits [B,t,V] fp32. With `self_cache` (list of [k,v] per layer,
        mutated in place) only the new tokens are fed and the cache supplies the history (upstream kv_cache hooks)."""
        w, dims = self.w, self.dims
        offset = 0 if not self_cache or self_cache[0] is None else self_cache[0][0].shape[1]
        t = tokens.shape[-1]
        x = w["decoder.token_embedding.weight"][tokens] + w["decoder.positional_embedding"][offset:offset + t]
        if cross is None:
            cross = self.cross_kv(xa)
        for i in range(dims.n_text_layer):
            p = f"decoder.blocks.{i}"
            h = _ln(x, w[f"{p}.attn_ln.weight"], w[f"{p}.attn_ln.bias"])
            k, v = self._proj_kv(f"{p}.attn", h)
            if self_cache is not None:
                if self_cache[i] is not None:
                    k = torch.cat([self_cache[i][0], k], dim=1)
                    v = torch.cat([self_cache[i][1], v], dim=1)
                self_cache[i] = (k, v)
            n = k.shape[1]
            mask = self.mask[n - t:n, :n]
            x = x + self._attend(f"{p}.attn", dims.n_text_head, h, k, v, mask)
            h = _ln(x, w[f"{p}.cross_attn_ln.weight"], w[f"{p}.cross_attn_ln.bias"])
            x = x + self._attend(f"{p}.cross_attn", dims.n_text_head, h, cross[i][0], cross[i][1])
            x = x + self._mlp(p, _ln(x, w[f"{p}.mlp_ln.weight"], w[f"{p}.mlp_ln.bias"]))
        x = _ln(x, w["decoder.ln.weight"], w["decoder.ln.bias"])
        return (x @ w["decoder.token_embedding.weight"].t()).float()

    # ---- Whisper.decode (Whisper.swift:33-40): language ID ------------------------------------------------------
    @torch.no_grad()
    def detect_language(self, xa: torch.Tensor) -> torch.Tensor:
        """One decoder call on the single token `sot`; argmax over the 99 language logits. Returns [B] in 0..98."""
        v = self.vocab
        B = xa.shape[0]
        logits = self.decoder_logits(torch.full((B, 1), v.sot, dtype=torch.long), xa)[:, 0]
        return language_argmax(logits, v.lang0)

    # ---- greedy transcribe (upstream DecodingTask with temperature 0) ---------------------------------------------
    @torch.no_grad()
    def greedy(self, xa: torch.Tensor, opts: "DecodeOptions"):
        B = xa.shape[0]
        v = self.vocab
        init = list(opts.initial_tokens)
        tokens = torch.tensor([init] * B, dtype=torch.long)
        sample_begin = len(init)
        sum_logprobs = torch.zeros(B)
        cross = self.cross_kv(xa)
        cache: List[Optional[tuple]] = [None] * self.dims.n_text_layer
        all_logits = []
        for i in range(opts.sample_len):
            feed = tokens if i == 0 else tokens[:, -1:]
            logits = self.decoder_logits(feed, xa, cross, cache)[:, -1].clone()
            if tokens.shape[1] == sample_begin and len(opts.suppress_begin):
                logits[:, list(opts.suppress_begin)] = float("-inf")
            if len(opts.suppress):
                logits[:, list(opts.suppress)] = float("-inf")
            if opts.timestamps:
                logits = apply_timestamp_rules(logits, tokens, sample_begin, v, opts.max_initial_timestamp_index)
            all_logits.append(logits)
            nxt = logits.argmax(dim=-1)
            logprobs = F.log_softmax(logits.float(), dim=-1)
            cur = logprobs[torch.arange(B), nxt]
            sum_logprobs += cur * (tokens[:, -1] != v.eot)
            nxt[tokens[:, -1] == v.eot] = v.eot
            tokens = torch.cat([tokens, nxt[:, None]], dim=-1)
            if (tokens[:, -1] == v.eot).all() or tokens.shape[-1] > self.dims.n_text_ctx:
                break
        return tokens, sum_logprobs, torch.stack(all_logits, dim=1)


    # ---- beam search (upstream DecodingTask with beam_size, BeamSearchDecoder, MaximumLikelihoodRanker) -----------------
    @torch.no_grad()
    def beam_search(self, xa: torch.Tensor, opts: "DecodeOptions", beam_size: int = 5, patience: float = 1.0):
        """Restatement of upstream whisper/decoding.py: BeamSearchDecoder.update / finalize and the length-normalised
        ranking (length_penalty=None -> sum_logprob / len). UNPINNED beyond the published algorithm: no independent
        implementation of this exact procedure exists in the image (transformers' beam search is a different algorithm).
        Returns (best token list per audio incl. the sot sequence, without the final eot; best sum_logprob per audio)."""
        v = self.vocab
        n_audio = xa.shape[0]
        init = list(opts.initial_tokens)
        sample_begin = len(init)
        max_candidates = round(beam_size * patience)
        xa_rep = xa.repeat_interleave(beam_size, dim=0)
        cross = self.cross_kv(xa_rep)
        tokens = torch.tensor([init] * (n_audio * beam_size), dtype=torch.long)
        sum_logprobs = torch.zeros(n_audio * beam_size)
        cache: List[Optional[tuple]] = [None] * self.dims.n_text_layer
        finished_sequences = [dict() for _ in range(n_audio)]
        for i in range(opts.sample_len):
            feed = tokens if i == 0 else tokens[:, -1:]
            logits = self.decoder_logits(feed, xa_rep, cross, cache)[:, -1].clone()
            if tokens.shape[1] == sample_begin and len(opts.suppress_begin):
                logits[:, list(opts.suppress_begin)] = float("-inf")
            if len(opts.suppress):
                logits[:, list(opts.suppress)] = float("-inf")
            if opts.timestamps:                     # upstream logit_filters order: SuppressBlank, SuppressTokens, ApplyTimestampRules
                logits = apply_timestamp_rules(logits, tokens, sample_begin, v, opts.max_initial_timestamp_index)
            logprobs = F.log_softmax(logits.float(), dim=-1)
            next_tokens, source_indices, newly = [], [], []
            for a in range(n_audio):
                scores, sources, finished = {}, {}, {}
                for j in range(beam_size):
                    idx = a * beam_size + j
                    prefix = tokens[idx].tolist()
                    for logprob, token in zip(*logprobs[idx].topk(beam_size + 1)):
                        sequence = tuple(prefix + [token.item()])
                        scores[sequence] = (sum_logprobs[idx] + logprob).item()
                        sources[sequence] = idx
                saved = 0
                for sequence in sorted(scores, key=scores.get, reverse=True):
                    if sequence[-1] == v.eot:
                        finished[sequence] = scores[sequence]
                    else:
                        sum_logprobs[len(next_tokens)] = scores[sequence]
                        next_tokens.append(sequence)
                        source_indices.append(sources[sequence])
                        saved += 1
                        if saved == beam_size:
                            break
                newly.append(finished)
            tokens = torch.tensor(next_tokens, dtype=torch.long)
            src = torch.tensor(source_indices, dtype=torch.long)
            cache = [(k.index_select(0, src), vv.index_select(0, src)) for (k, vv) in cache]     # rearrange_kv_cache
            for prev, new in zip(finished_sequences, newly):
                for seq in sorted(new, key=new.get, reverse=True):
                    if len(prev) >= max_candidates:
                        break
                    prev[seq] = new[seq]
            completed = all(len(s) >= max_candidates for s in finished_sequences)
            if completed or tokens.shape[-1] > self.dims.n_text_ctx:
                break
        # finalize: add unfinished beams (with eot appended) when not enough sequences finished
        tokens = tokens.reshape(n_audio, beam_size, -1)
        slp = sum_logprobs.reshape(n_audio, beam_size)
        best_tokens, best_scores = [], []
        for a, sequences in enumerate(finished_sequences):
            if len(sequences) < beam_size:
                for j in list(np.argsort(slp[a].numpy()))[::-1]:
                    sequence = tokens[a, j].tolist() + [v.eot]
                    sequences[tuple(sequence)] = slp[a][j].item()
                    if len(sequences) >= beam_size:
                        break
            cands = [list(seq) for seq in sequences.keys()]
            vals = list(sequences.values())
            trimmed = [c[sample_begin:c.index(v.eot, sample_begin)] if v.eot in c[sample_begin:] else c[sample_begin:] for c in cands]
            ranks = [val / len(t) if len(t) else float("-inf") for val, t in zip(vals, trimmed)]   # MaximumLikelihoodRanker, no penalty
            k = int(np.argmax(ranks))
            best_tokens.append(init + trimmed[k])
            best_scores.append(vals[k])
        return best_tokens, best_scores


def language_argmax(logits: torch.Tensor, lang0: int = 50259) -> torch.Tensor:
    """Whisper.swift:37-38: `(50259...50357).map{...}.enumerated().max{ $0.element < $1.element }`.
    Swift's `Sequence.max(by:)` replaces its running result only when `areInIncreasingOrder(result, e)` holds, i.e. on a
    strict increase: the FIRST maximal element wins ties (all-equal logits -> index 0, "en"), and a NaN never replaces
    the running result."""
    conf = logits[..., lang0:lang0 + 99].float()
    best = torch.zeros(conf.shape[:-1], dtype=torch.long)
    cur = conf[..., 0].clone()
    for i in range(1, 99):
        take = cur < conf[..., i]
        best = torch.where(take, torch.full_like(best, i), best)
        cur = torch.where(take, conf[..., i], cur)
    return best


@dataclass
class DecodeOptions:
    initial_tokens: Sequence[int]
    sample_len: int = 224                      # upstream default n_text_ctx // 2
    suppress: Sequence[int] = field(default_factory=list)         # SuppressTokens (every step)
    suppress_begin: Sequence[int] = field(default_factory=list)   # SuppressBlank (first sampled position only)
    timestamps: bool = False                   # ApplyTimestampRules (upstream: unless without_timestamps)
    max_initial_timestamp_index: Optional[int] = 50   # upstream max_initial_timestamp = 1.0 s / 0.02 s per timestamp token

    @staticmethod
    def default_for(dims: ModelDims, sample_len: int = 224, language: int = 0, without_timestamps: bool = True) -> "DecodeOptions":
        """Upstream defaults. The non-speech symbol list needs the tokenizer vocab, which is not available offline; the
        special tokens upstream always suppresses are included. `without_timestamps=False` is upstream's own default:
        no <|notimestamps|> in the prompt and ApplyTimestampRules among the logit filters."""
        v = Vocab.for_dims(dims)
        if dims.is_multilingual:
            init = [v.sot, v.lang0 + language, v.transcribe]
        else:
            init = [v.sot]
        if without_timestamps:
            init.append(v.no_timestamps)
        suppress = sorted({v.sot, v.sot_prev, v.sot_lm, v.translate, v.transcribe, v.no_speech})
        return DecodeOptions(init, sample_len, suppress, [220, v.eot], timestamps=not without_timestamps)


def apply_timestamp_rules(logits: torch.Tensor, tokens: torch.Tensor, sample_begin: int, vocab: "Vocab",
                          max_initial_timestamp_index: Optional[int]) -> torch.Tensor:
    """Restatement of upstream whisper/decoding.py ApplyTimestampRules.apply (the version with the non-decreasing rule of
    openai/whisper PR 914): logits [B,V] of the next position given tokens [B,t] (prompt included). Pinned in
    tests/test_oracle_whisper.py against transformers' WhisperTimeStampLogitsProcessor, an independent implementation."""
    logits = logits.clone()
    ts_begin = vocab.timestamp_begin
    logits[:, vocab.no_timestamps] = float("-inf")           # handled by without_timestamps
    for k in range(tokens.shape[0]):                         # timestamps appear in pairs, except directly before EOT
        seq = tokens[k, sample_begin:].tolist()
        last_was_timestamp = len(seq) >= 1 and seq[-1] >= ts_begin
        penultimate_was_timestamp = len(seq) < 2 or seq[-2] >= ts_begin
        if last_was_timestamp:
            if penultimate_was_timestamp:                    # has to be non-timestamp
                logits[k, ts_begin:] = float("-inf")
            else:                                            # cannot be normal text tokens
                logits[k, :vocab.eot] = float("-inf")
        stamps = [t for t in seq if t >= ts_begin]
        if stamps:                                           # timestamps shouldn't decrease; segments have nonzero length
            last = stamps[-1] if (last_was_timestamp and not penultimate_was_timestamp) else stamps[-1] + 1
            logits[k, ts_begin:last] = float("-inf")
    if tokens.shape[1] == sample_begin:                      # first position: a timestamp, at most max_initial_timestamp
        logits[:, :ts_begin] = float("-inf")
        if max_initial_timestamp_index is not None:
            logits[:, ts_begin + max_initial_timestamp_index + 1:] = float("-inf")
    logprobs = F.log_softmax(logits.float(), dim=-1)         # probability mass over timestamps above every text token: timestamp
    for k in range(tokens.shape[0]):
        if logprobs[k, ts_begin:].logsumexp(dim=-1) > logprobs[k, :ts_begin].max():
            logits[k, :ts_begin] = float("-inf")
    return logits


# ---- independent cross-check: map these weights into transformers' Whisper ------------------------------------------
def to_hf(dims: ModelDims, weights: Dict[str, torch.Tensor]):
    """Build transformers.WhisperForConditionalGeneration holding exactly `weights` (upstream->HF name map)."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    cfg = WhisperConfig(
        vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_audio_state,
        encoder_layers=dims.n_audio_layer, encoder_attention_heads=dims.n_audio_head,
        decoder_layers=dims.n_text_layer, decoder_attention_heads=dims.n_text_head,
        encoder_ffn_dim=4 * dims.n_audio_state, decoder_ffn_dim=4 * dims.n_text_state,
        max_source_positions=dims.n_audio_ctx, max_target_positions=dims.n_text_ctx,
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        pad_token_id=0, bos_token_id=0, eos_token_id=0, decoder_start_token_id=0,
        suppress_tokens=None, begin_suppress_tokens=None,
    )
    cfg._attn_implementation = "eager"
    model = WhisperForConditionalGeneration(cfg).eval().float()

    def ren(k: str) -> str:
        k = k.replace("blocks.", "layers.")
        for a, b in ((".cross_attn_ln.", ".encoder_attn_layer_norm."), (".attn_ln.", ".self_attn_layer_norm."),
                     (".mlp_ln.", ".final_layer_norm."), (".cross_attn.", ".encoder_attn."), (".attn.", ".self_attn."),
                     (".query.", ".q_proj."), (".key.", ".k_proj."), (".value.", ".v_proj."), (".out.", ".out_proj."),
                     (".mlp.0.", ".fc1."), (".mlp.2.", ".fc2.")):
            k = k.replace(a, b)
        k = k.replace("encoder.ln_post.", "encoder.layer_norm.").replace("decoder.ln.", "decoder.layer_norm.")
        k = k.replace("encoder.positional_embedding", "encoder.embed_positions.weight")
        k = k.replace("decoder.positional_embedding", "decoder.embed_positions.weight")
        k = k.replace("decoder.token_embedding.", "decoder.embed_tokens.")
        return "model." + k

    sd = {ren(k): v.clone() for k, v in weights.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m for m in missing) or not missing, missing
    return model


def synth_audio(seed: int, kind: str = "noise", n: int = 480000) -> np.ndarray:
    """The synthetic clips SURVEY.md §8(d) config 1 names. float64, 16 kHz, 30 s."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    if kind == "noise":
        return rng.standard_normal(n) * 0.1
    if kind == "zeros":
        return np.zeros(n)
    if kind == "sine":
        return 0.5 * np.sin(2 * np.pi * 440.0 * t)
    if kind == "chirp":
        return 0.5 * np.sin(2 * np.pi * (7000.0 / 60.0) * t * t)
    if kind == "noise_then_zeros":
        a = np.zeros(n)
        a[:160000] = rng.standard_normal(160000) * 0.1
        return a
    if kind == "int16":
        a = rng.standard_normal(n) * 0.1 * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
        return np.round(a * 32768.0) / 32768.0
    if kind == "fullscale":
        return rng.uniform(-1.0, 1.0, n)
    raise ValueError(kind)


# ---- long-form transcription: upstream whisper/transcribe.py transcribe() ---------------------------
# PARITY UNPINNED like the rest of this file (openai-whisper is not in the image): restated from the published source of
# openai-whisper v20230314 ... v20231117 — `log_mel_spectrogram(audio, padding=N_SAMPLES)`, `decode_with_fallback`, the
# no-speech skip, the consecutive-timestamp segment cutter, `seek` arithmetic, the prompt reset above temperature 0.5 —
# without word_timestamps / clip_timestamps / hallucination_silence_threshold. The pieces that ARE pinned independently:
# the log-mel of a long signal against torch.stft here, the token table / compression ratio against tiktoken and Python's
# zlib (tests/test_host.py), the timestamp rules against transformers (tests/test_oracle_whisper.py).
#
# Sampling at temperature > 0: upstream draws `Categorical(logits=logits / temperature).sample()` from torch's global
# generator, which no second implementation can reproduce. Both this restatement and the CUDA path define the draw as
# the Gumbel-max form of the same distribution over a counter-based generator (include/whisper_b200.h, decoder.cu
# sample_rows_kernel), so that they see the same noise.
_M64 = (1 << 64) - 1


def splitmix64(x):
    """Vectorised over numpy uint64 arrays (wraps modulo 2**64) or a Python int."""
    if isinstance(x, np.ndarray):
        with np.errstate(over="ignore"):
            x = x + np.uint64(0x9E3779B97F4A7C15)
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return x ^ (x >> np.uint64(31))
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def call_seed(seed: int, seek: int, temperature_index: int) -> int:
    return splitmix64((splitmix64((seed ^ seek) & _M64) + temperature_index) & _M64)


def gumbel_noise(seed: int, sample: int, position: int, n_vocab: int) -> torch.Tensor:
    key = splitmix64((seed ^ splitmix64(((sample << 32) | position) & _M64)) & _M64)
    with np.errstate(over="ignore"):
        r = splitmix64(np.uint64(key) + np.arange(n_vocab, dtype=np.uint64))
    u = ((r >> np.uint64(41)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)   # 23 bits: strictly inside (0, 1) in fp32
    return torch.from_numpy(-np.log(-np.log(u))).float()


def log_mel_stream(audio: np.ndarray, mel_filters: np.ndarray, padding: int = 480000) -> torch.Tensor:
    """upstream whisper/audio.py log_mel_spectrogram(audio, padding=N_SAMPLES): [80, (n + padding) // 160] float32."""
    a = F.pad(torch.from_numpy(np.asarray(audio, dtype=np.float32)), (0, padding))
    stft = torch.stft(a, 400, 160, window=torch.hann_window(400), return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = torch.from_numpy(np.asarray(mel_filters, dtype=np.float32).reshape(80, 201)) @ magnitudes   # m80.npy holds 80 x 201
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def text_compression_ratio(raw: bytes) -> float:
    text = raw.decode("utf-8", errors="replace").strip().encode("utf-8")
    import zlib
    return len(text) / len(zlib.compress(text))


@dataclass
class WindowResult:
    tokens: List[int]
    avg_logprob: float
    no_speech_prob: float
    temperature: float
    compression_ratio: float
    min_margin: float = float("inf")   # smallest top-1 / top-2 gap of the (perturbed) logits over every draw of this decode


@torch.no_grad()
def decode_window(model: "WhisperRef", xa: torch.Tensor, opts: "DecodeOptions", prompt: Sequence[int], temperature: float, seed: int,
                  best_of: int = 0, sot_index: int = 0, table: Optional[List[bytes]] = None) -> WindowResult:
    """upstream DecodingTask.run for one window: [sot_prev] + prompt[-(n_ctx // 2 - 1):] + sot sequence, no_speech_probs at the
    sot position, GreedyDecoder at `temperature` with best_of samples, MaximumLikelihoodRanker, DecodingResult fields."""
    v, dims = model.vocab, model.dims
    n_ctx = dims.n_text_ctx
    init: List[int] = []
    if len(prompt):
        init = [v.sot_prev] + list(prompt)[-(n_ctx // 2 - 1):]
    sot_at = len(init) + sot_index
    init = init + list(opts.initial_tokens)
    sample_begin = len(init)
    sample_len = min(opts.sample_len, n_ctx - sample_begin)          # the CUDA path's cap (upstream: one more)
    G = best_of if (temperature > 0 and best_of > 1) else 1
    tokens = torch.tensor([init] * G, dtype=torch.long)
    xg = xa[:1].repeat_interleave(G, dim=0)
    cross = model.cross_kv(xg)
    cache: List[Optional[tuple]] = [None] * dims.n_text_layer
    sum_logprobs = torch.zeros(G)
    no_speech_prob = 0.0
    min_margin = float("inf")
    for i in range(sample_len):
        feed = tokens if i == 0 else tokens[:, -1:]
        all_logits = model.decoder_logits(feed, xg, cross, cache)
        if i == 0:
            no_speech_prob = float(all_logits[0, sot_at].float().softmax(dim=-1)[v.no_speech])
        logits = all_logits[:, -1].clone()
        if tokens.shape[1] == sample_begin and len(opts.suppress_begin):
            logits[:, list(opts.suppress_begin)] = float("-inf")
        if len(opts.suppress):
            logits[:, list(opts.suppress)] = float("-inf")
        if opts.timestamps:
            logits = apply_timestamp_rules(logits, tokens, sample_begin, v, opts.max_initial_timestamp_index)
        if temperature == 0:
            scored = logits
        else:
            position = tokens.shape[1]
            inv_t = np.float32(1.0) / np.float32(temperature)
            scored = torch.stack([logits[j] * float(inv_t) + gumbel_noise(seed, j, position, dims.n_vocab) for j in range(G)])
        nxt = scored.argmax(dim=-1)
        top2 = scored.topk(2, dim=-1).values
        live = tokens[:, -1] != v.eot
        if bool(live.any()):
            min_margin = min(min_margin, float((top2[:, 0] - top2[:, 1])[live].min()))
        logprobs = F.log_softmax(logits.float(), dim=-1)
        cur = logprobs[torch.arange(G), nxt]
        sum_logprobs += cur * (tokens[:, -1] != v.eot)
        nxt[tokens[:, -1] == v.eot] = v.eot
        tokens = torch.cat([tokens, nxt[:, None]], dim=-1)
        if (tokens[:, -1] == v.eot).all():
            break
    best, best_score, best_tokens = -1, 0.0, []
    for j in range(G):
        t = tokens[j, sample_begin:].tolist()
        t = t[:t.index(v.eot)] if v.eot in t else t
        score = float(sum_logprobs[j]) / max(len(t), 1)
        if best < 0 or score > best_score:
            best, best_score, best_tokens = j, score, t
    cr = 0.0
    if table is not None:
        drop = v.timestamp_begin if opts.timestamps else v.eot
        cr = text_compression_ratio(b"".join(table[t] for t in best_tokens if t < drop and t < len(table)))
    return WindowResult(best_tokens, float(sum_logprobs[best]) / (len(best_tokens) + 1), no_speech_prob, float(temperature), cr, min_margin)


@torch.no_grad()
def transcribe_seek(model: "WhisperRef", audio: np.ndarray, mel_filters: np.ndarray, opts: "DecodeOptions", *,
                    temperatures: Sequence[float] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0), compression_ratio_threshold: Optional[float] = 2.4,
                    logprob_threshold: Optional[float] = -1.0, no_speech_threshold: Optional[float] = 0.6,
                    condition_on_previous_text: bool = True, initial_prompt: Sequence[int] = (), table: Optional[List[bytes]] = None,
                    seed: int = 0, best_of: int = 0, sot_index: int = 0):
    """upstream transcribe(): returns (all_tokens without the initial prompt, segments as dicts, log of (seek, temperatures tried))."""
    v = model.vocab
    N_FRAMES, HOP, SR = 3000, 160, 16000
    mel = log_mel_stream(audio, mel_filters)
    content_frames = mel.shape[-1] - N_FRAMES
    input_stride = N_FRAMES // model.dims.n_audio_ctx
    time_precision = input_stride * HOP / SR
    ts_begin = v.timestamp_begin if opts.timestamps else 1 << 30
    use_cr = compression_ratio_threshold is not None and table is not None
    all_tokens: List[int] = list(initial_prompt)
    n_prompt0 = len(all_tokens)
    prompt_reset_since = 0
    segments, trace = [], []
    seek = 0
    while seek < content_frames:
        time_offset = float(seek * HOP / SR)
        mel_segment = mel[:, seek:seek + N_FRAMES]
        segment_size = min(N_FRAMES, content_frames - seek)
        segment_duration = segment_size * HOP / SR
        if mel_segment.shape[-1] < N_FRAMES:
            mel_segment = F.pad(mel_segment, (0, N_FRAMES - mel_segment.shape[-1]))
        xa = model.encode(mel_segment[None])
        prompt = all_tokens[prompt_reset_since:]
        tried = []
        worst = float("inf")
        for ti, t in enumerate(temperatures):
            res = decode_window(model, xa, opts, prompt, t, call_seed(seed, seek, ti), best_of, sot_index, table)
            tried.append(t)
            worst = min(worst, res.min_margin)
            needs_fallback = False
            if use_cr and res.compression_ratio > compression_ratio_threshold:
                needs_fallback = True
            if logprob_threshold is not None and res.avg_logprob < logprob_threshold:
                needs_fallback = True
            if no_speech_threshold is not None and res.no_speech_prob > no_speech_threshold:
                needs_fallback = False
            if not needs_fallback:
                break
        res.min_margin = worst            # over every decode of this window, the rejected ones included
        trace.append((seek, tried, res))
        tokens = res.tokens
        if no_speech_threshold is not None:
            should_skip = res.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and res.avg_logprob > logprob_threshold:
                should_skip = False
            if should_skip:
                seek += segment_size
                continue
        current = []

        def new_segment(start, end, toks):
            current.append({"seek": seek, "start": start, "end": end, "tokens": list(toks), "temperature": res.temperature,
                            "avg_logprob": res.avg_logprob, "compression_ratio": res.compression_ratio, "no_speech_prob": res.no_speech_prob})

        is_ts = [t >= ts_begin for t in tokens]
        single_timestamp_ending = is_ts[-2:] == [False, True]
        consecutive = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
        if consecutive:
            slices = list(consecutive)
            if single_timestamp_ending:
                slices.append(len(tokens))
            last_slice = 0
            for current_slice in slices:
                sliced = tokens[last_slice:current_slice]
                new_segment(time_offset + (sliced[0] - ts_begin) * time_precision, time_offset + (sliced[-1] - ts_begin) * time_precision, sliced)
                last_slice = current_slice
            if single_timestamp_ending:
                seek += segment_size
            else:
                seek += (tokens[last_slice - 1] - ts_begin) * input_stride
        else:
            duration = segment_duration
            stamps = [t for t in tokens if t >= ts_begin]
            if stamps and stamps[-1] != ts_begin:
                duration = (stamps[-1] - ts_begin) * time_precision
            new_segment(time_offset, time_offset + duration, tokens)
            seek += segment_size
        for s in current:
            text_tokens = [t for t in s["tokens"] if t < v.eot]
            if table is not None:
                has_text = b"".join(table[t] for t in text_tokens if t < len(table)).decode("utf-8", errors="replace").strip() != ""
            else:
                has_text = len(text_tokens) > 0
            # the CUDA path reports float32 seconds: compare start and end as it does
            if np.float32(s["start"]) == np.float32(s["end"]) or not has_text:
                s["tokens"] = []
        segments.extend(current)
        all_tokens.extend(t for s in current for t in s["tokens"])
        if not condition_on_previous_text or res.temperature > 0.5:
            prompt_reset_since = len(all_tokens)
    return all_tokens[n_prompt0:], segments, trace
