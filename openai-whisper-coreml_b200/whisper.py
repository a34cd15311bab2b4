"""ctypes binding of libwhisper_b200.so and the Python mirror of the reference's Swift API.

Reference interfaces mirrored here (file:line under /root/reference):
  generateSpectrogram(audio: [Double]) -> [Double]      Whisper/Whisper/stft.swift:8-19
  Whisper.init() throws                                 Whisper/Whisper/Whisper.swift:17-21
  Whisper.encode(audio: [Double]) -> MLMultiArray       Whisper/Whisper/Whisper.swift:23-31
  Whisper.decode(audioFeatures:)  (prints a language)   Whisper/Whisper/Whisper.swift:33-40
  input contract: pad/truncate to 480000 samples        Whisper/Whisper/ContentView.swift:57-60
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
N_SAMPLES = 480000
N_MELS = 80
N_FRAMES = 3000

# Whisper.swift:12 — the 99 language codes, index = token id - 50259
LANGUAGES = ("en,zh,de,es,ru,ko,fr,ja,pt,tr,pl,ca,nl,ar,sv,it,id,hi,fi,vi,iw,uk,el,ms,cs,ro,da,hu,ta,no,th,ur,hr,bg,lt,la,"
             "mi,ml,cy,sk,te,fa,lv,bn,sr,az,sl,kn,et,mk,br,eu,is,hy,ne,mn,bs,kk,sq,sw,gl,mr,pa,si,km,sn,yo,so,af,oc,ka,be,"
             "tg,sd,gu,am,yi,lo,uz,fo,ht,ps,tk,nn,mt,sa,lb,my,bo,tl,mg,as,tt,haw,ln,ha,ba,jw,su").split(",")


class WhisperB200Error(RuntimeError):
    pass


class _Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
                 "n_vocab", "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer")]


class _DecodeOpts(ctypes.Structure):
    _fields_ = [("initial_tokens", ctypes.POINTER(ctypes.c_int32)), ("n_initial", ctypes.c_int32),
                ("sample_len", ctypes.c_int32), ("eot", ctypes.c_int32),
                ("suppress", ctypes.POINTER(ctypes.c_int32)), ("n_suppress", ctypes.c_int32),
                ("suppress_begin", ctypes.POINTER(ctypes.c_int32)), ("n_suppress_begin", ctypes.c_int32),
                ("beam_size", ctypes.c_int32), ("eot_check_interval", ctypes.c_int32),
                ("timestamps", ctypes.c_int32), ("timestamp_begin", ctypes.c_int32), ("no_timestamps", ctypes.c_int32),
                ("max_initial_timestamp_index", ctypes.c_int32),
                ("temperature", ctypes.c_float), ("best_of", ctypes.c_int32), ("seed", ctypes.c_uint64),
                ("no_speech", ctypes.c_int32), ("sot_index", ctypes.c_int32), ("no_speech_prob", ctypes.POINTER(ctypes.c_float))]


class _LongOpts(ctypes.Structure):
    _fields_ = [("decode", _DecodeOpts), ("temperatures", ctypes.POINTER(ctypes.c_float)), ("n_temperatures", ctypes.c_int32),
                ("compression_ratio_threshold", ctypes.c_float), ("logprob_threshold", ctypes.c_float),
                ("no_speech_threshold", ctypes.c_float), ("condition_on_previous_text", ctypes.c_int32),
                ("initial_prompt", ctypes.POINTER(ctypes.c_int32)), ("n_initial_prompt", ctypes.c_int32),
                ("sot_prev", ctypes.c_int32), ("tokenizer", ctypes.c_void_p), ("detect_language", ctypes.c_int32),
                ("lang0", ctypes.c_int32), ("detected_language", ctypes.POINTER(ctypes.c_int32))]


class _Segment(ctypes.Structure):
    _fields_ = [("seek", ctypes.c_int32), ("start", ctypes.c_float), ("end", ctypes.c_float), ("token_begin", ctypes.c_int32),
                ("n_tokens", ctypes.c_int32), ("temperature", ctypes.c_float), ("avg_logprob", ctypes.c_float),
                ("compression_ratio", ctypes.c_float), ("no_speech_prob", ctypes.c_float)]


@dataclass(frozen=True)
class ModelDims:
    """Upstream ModelDimensions (whisper_to_cml.py:13,29 pin n_mels=80, n_audio_ctx=1500)."""
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


DIMS: Dict[str, ModelDims] = {
    "tiny.en": ModelDims(80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4),
    "tiny": ModelDims(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base.en": ModelDims(80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6),
    "base": ModelDims(80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small.en": ModelDims(80, 1500, 768, 12, 12, 51864, 448, 768, 12, 12),
    "small": ModelDims(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium": ModelDims(80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v2": ModelDims(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
}


@dataclass
class DecodeOptions:
    """Greedy decoding options (upstream DecodingOptions subset; token lists are data)."""
    initial_tokens: Sequence[int]
    eot: int
    sample_len: int = 224
    suppress: Sequence[int] = field(default_factory=list)
    suppress_begin: Sequence[int] = field(default_factory=list)
    beam_size: int = 0
    eot_check_interval: int = 8
    timestamps: bool = False                 # upstream ApplyTimestampRules (DecodingOptions.without_timestamps = False); greedy only
    timestamp_begin: int = 0                 # <|0.00|>
    no_timestamps: int = 0                   # <|notimestamps|>
    max_initial_timestamp_index: int = 50    # upstream max_initial_timestamp = 1.0 s; < 0: no limit
    temperature: float = 0.0                 # > 0: Gumbel-max draw from softmax(logits / temperature) (whisper_b200.h)
    best_of: int = 0                         # temperature > 0: samples per chunk, best sum_logprob / length wins
    seed: int = 0
    no_speech: int = -1                      # <|nospeech|>; with want_no_speech_prob the decode also returns its probability
    sot_index: int = 0                       # index of <|startoftranscript|> in initial_tokens
    sot_prev: int = 0                        # <|startofprev|> (long-form prompts)

    @staticmethod
    def default_for(dims: ModelDims, sample_len: int = 224, language: int = 0, without_timestamps: bool = True) -> "DecodeOptions":
        if dims.is_multilingual:
            eot, sot, lang0, translate, transcribe, sot_lm, sot_prev, no_speech, no_ts = (
                50257, 50258, 50259, 50358, 50359, 50360, 50361, 50362, 50363)
            init = [sot, lang0 + language, transcribe]
        else:
            eot, sot, translate, transcribe, sot_lm, sot_prev, no_speech, no_ts = (
                50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362)
            init = [sot]
        if without_timestamps:
            init.append(no_ts)
        suppress = sorted({sot, sot_prev, sot_lm, translate, transcribe, no_speech})
        return DecodeOptions(init, eot, sample_len, suppress, [220, eot], timestamps=not without_timestamps,
                             timestamp_begin=no_ts + 1, no_timestamps=no_ts, no_speech=no_speech, sot_index=0, sot_prev=sot_prev)


_HF_RULES = (("layers.", "blocks."), (".encoder_attn_layer_norm.", ".cross_attn_ln."), (".self_attn_layer_norm.", ".attn_ln."),
             (".final_layer_norm.", ".mlp_ln."), (".encoder_attn.", ".cross_attn."), (".self_attn.", ".attn."),
             (".q_proj.", ".query."), (".k_proj.", ".key."), (".v_proj.", ".value."), (".out_proj.", ".out."),
             (".fc1.", ".mlp.0."), (".fc2.", ".mlp.2."))


def hf_to_upstream_name(key: str) -> Optional[str]:
    """Maps a transformers WhisperForConditionalGeneration state-dict key to the upstream openai-whisper name the C ABI
    uses (SURVEY.md §8c weight-name map). Returns None for tensors that have no upstream counterpart (the tied
    `proj_out.weight`, the always-zero `k_proj.bias`)."""
    if key == "proj_out.weight" or key.endswith("k_proj.bias"):
        return None
    k = key[len("model."):] if key.startswith("model.") else key
    for a, b in _HF_RULES:
        k = k.replace(a, b)
    k = k.replace("encoder.layer_norm.", "encoder.ln_post.").replace("decoder.layer_norm.", "decoder.ln.")
    k = k.replace("encoder.embed_positions.weight", "encoder.positional_embedding")
    k = k.replace("decoder.embed_positions.weight", "decoder.positional_embedding")
    k = k.replace("decoder.embed_tokens.", "decoder.token_embedding.")
    return k


def split_windows(pcm: Sequence[float], window: int = N_SAMPLES) -> np.ndarray:
    """Cuts a long 16 kHz stream into consecutive fixed 30 s windows, zero-padding the last one (BASELINE config 5: a
    30 min stream = 60 independent windows; each window follows the pad/truncate contract of ContentView.swift:57-60)."""
    a = np.asarray(pcm, dtype=np.float32).reshape(-1)
    n = max(1, -(-a.shape[0] // window))
    out = np.zeros((n, window), dtype=np.float32)
    out.reshape(-1)[: a.shape[0]] = a
    return out


_LIB: Optional[ctypes.CDLL] = None
_SYMBOLS = {
    "generate_spectrogram": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "wb_generate_spectrogram_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "wb_last_error": (ctypes.c_char_p, []),
    "wb_version": (ctypes.c_int, []),
    "wb_create": (ctypes.c_int, [ctypes.POINTER(_Dims), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_void_p)]),
    "wb_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "wb_get_dims": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(_Dims)]),
    "wb_set_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]),
    "wb_weights_commit": (ctypes.c_int, [ctypes.c_void_p]),
    "wb_init_random_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "wb_weight_arena": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "wb_weights_mark_loaded": (ctypes.c_int, [ctypes.c_void_p]),
    "wb_get_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]),
    "wb_weight_count": (ctypes.c_int, [ctypes.c_void_p]),
    "wb_weight_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "wb_weights_checksum": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "wb_logmel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "wb_logmel_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "wb_encode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "wb_encode_mel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "wb_encode_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]),
    "wb_set_audio_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]),
    "wb_get_audio_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]),
    "wb_decoder_logits": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_decoder_logits_f32tok": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_detect_language": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts), ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]),
    "wb_transcribe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts),
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wb_transcribe_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts),
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wb_safetensors_read_dims": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_Dims)]),
    "wb_load_safetensors": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32)]),
    "wb_tokenizer_create": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]),
    "wb_tokenizer_load_tiktoken": (ctypes.c_void_p, [ctypes.c_char_p]),
    "wb_tokenizer_destroy": (None, [ctypes.c_void_p]),
    "wb_tokenizer_size": (ctypes.c_int32, [ctypes.c_void_p]),
    "wb_tokenizer_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "wb_text_compression_ratio": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_float),
                                                 ctypes.POINTER(ctypes.c_size_t)]),
    "wb_transcribe_long": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(_LongOpts),
                                          ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p,
                                          ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "wb_call_seed": (ctypes.c_uint64, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32]),
    "wb_logmel_long": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "wb_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "wb_last_timings": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "wb_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "wb_profile_cross_attention": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "wb_op_gemm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]),
    "wb_op_layernorm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_void_p]),
    "wb_op_attention": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p]),
}


def library_path() -> str:
    return os.environ.get("WHISPER_B200_LIB", os.path.join(_HERE, "libwhisper_b200.so"))


def exported_symbols() -> List[str]:
    return sorted(_SYMBOLS)


def load_library() -> ctypes.CDLL:
    """Loads libwhisper_b200.so; raises (no fallback) when it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise WhisperB200Error(f"{path} not found: build it with `python openai-whisper-coreml_b200/build.py` "
                                   "(or __graft_entry__.build()); there is no CPU fallback")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().wb_last_error()
        raise WhisperB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def pad_or_trim(audio: Sequence[float]) -> np.ndarray:
    """ContentView.swift:57-60: zero-pad or truncate a clip to exactly 480000 samples ([Float] -> [Double])."""
    a = np.asarray(audio, dtype=np.float64).reshape(-1)
    out = np.zeros(N_SAMPLES, dtype=np.float64)
    n = min(a.shape[0], N_SAMPLES)
    out[:n] = a[:n]
    return out


def generateSpectrogram(audio: Sequence[float]) -> np.ndarray:
    """stft.swift:8-19: prepend/append 200 zeros, call the C symbol generate_spectrogram, return 80*3000 doubles.
    The computation runs in f64 on the GPU (the crate's path is f64)."""
    a = np.asarray(audio, dtype=np.float64).reshape(-1)
    if a.shape[0] != N_SAMPLES:
        raise ValueError(f"generateSpectrogram expects {N_SAMPLES} samples (got {a.shape[0]}); see pad_or_trim")
    buf = np.zeros(N_SAMPLES + 400, dtype=np.float64)          # stft.swift:10-11
    buf[200:200 + N_SAMPLES] = a
    result = np.zeros(N_MELS * N_FRAMES, dtype=np.float64)     # stft.swift:12
    lib = load_library()
    _check(lib.wb_generate_spectrogram_f64(_ptr(buf), 1, _ptr(result)), "generate_spectrogram")
    return result


def read_checkpoint_dims(path: str) -> ModelDims:
    """Model dimensions of a safetensors checkpoint, from its tensor shapes (wb_safetensors_read_dims; host only)."""
    d = _Dims()
    _check(load_library().wb_safetensors_read_dims(path.encode(), ctypes.byref(d)), "wb_safetensors_read_dims")
    return ModelDims(*[getattr(d, f[0]) for f in _Dims._fields_])


def bytes_to_unicode() -> Dict[int, str]:
    """GPT-2's reversible byte <-> printable-character map (the alphabet of vocab.json / merges.txt)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


_GPT2_PAT = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


class Tokenizer:
    """Byte-level BPE vocabulary of Whisper (upstream whisper/tokenizer.py over tiktoken).

    Detokenisation (token ids -> bytes -> text, compression ratio) lives behind the C ABI (wb_tokenizer_*), because the
    long-form loop needs it natively; encoding (text -> ids, used only for prompts) is the published BPE merge loop here.
    Sources: an upstream `.tiktoken` rank file, a transformers `vocab.json` (+ the byte map above), or a dict of ranks."""

    def __init__(self, ranks: Dict[bytes, int]):
        self._lib = load_library()
        n = max(ranks.values()) + 1
        table: List[bytes] = [b""] * n
        for tok, r in ranks.items():
            table[r] = tok
        self.ranks = dict(ranks)
        self.table = table
        blob = b"".join(table)
        offs = np.zeros(n + 1, dtype=np.uint32)
        np.cumsum([len(t) for t in table], out=offs[1:])
        buf = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(1, dtype=np.uint8)
        self._h = ctypes.c_void_p(self._lib.wb_tokenizer_create(_ptr(buf), _ptr(offs), n))
        if not self._h:
            raise WhisperB200Error("wb_tokenizer_create failed: " + (self._lib.wb_last_error() or b"").decode())

    @classmethod
    def from_tiktoken(cls, path: str) -> "Tokenizer":
        import base64
        ranks = {}
        with open(path, "rb") as f:
            for line in f:
                if line.strip():
                    tok, rank = line.split()
                    ranks[base64.b64decode(tok)] = int(rank)
        return cls(ranks)

    @classmethod
    def from_vocab_json(cls, path: str) -> "Tokenizer":
        import json
        inv = {c: b for b, c in bytes_to_unicode().items()}
        with open(path, "r", encoding="utf-8") as f:
            vocab = json.load(f)
        ranks = {}
        for tok, idx in vocab.items():
            if all(ch in inv for ch in tok):            # special tokens (<|...|>) are not byte-level entries
                ranks[bytes(inv[ch] for ch in tok)] = idx
        return cls(ranks)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.wb_tokenizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    def __len__(self) -> int:
        return int(self._lib.wb_tokenizer_size(self._h))

    def decode_bytes(self, ids: Sequence[int], drop_from: int = 1 << 30) -> bytes:
        a = np.asarray(list(ids), dtype=np.int32)
        n = ctypes.c_size_t(0)
        _check(self._lib.wb_tokenizer_decode(self._h, _ptr(a) if a.size else None, a.size, drop_from, None, 0, ctypes.byref(n)),
               "wb_tokenizer_decode")
        out = np.zeros(max(1, n.value), dtype=np.uint8)
        _check(self._lib.wb_tokenizer_decode(self._h, _ptr(a) if a.size else None, a.size, drop_from, _ptr(out), out.size,
                                             ctypes.byref(n)), "wb_tokenizer_decode")
        return out[:n.value].tobytes()

    def decode(self, ids: Sequence[int], drop_from: int = 1 << 30) -> str:
        return self.decode_bytes(ids, drop_from).decode("utf-8", errors="replace")

    def compression_ratio(self, ids: Sequence[int], drop_from: int = 1 << 30) -> float:
        raw = np.frombuffer(self.decode_bytes(ids, drop_from) or b"\0", dtype=np.uint8)
        n = len(self.decode_bytes(ids, drop_from))
        r = ctypes.c_float(0)
        _check(self._lib.wb_text_compression_ratio(_ptr(raw), n, ctypes.byref(r), None), "wb_text_compression_ratio")
        return float(r.value)

    def encode(self, text: str) -> List[int]:
        """tiktoken's encode_ordinary: GPT-2 pre-tokenisation, then lowest-rank-first merges per piece."""
        import regex
        out: List[int] = []
        for piece in regex.findall(_GPT2_PAT, text):
            parts = [bytes([b]) for b in piece.encode("utf-8")]
            while len(parts) > 1:
                best, at = None, -1
                for i in range(len(parts) - 1):
                    r = self.ranks.get(parts[i] + parts[i + 1])
                    if r is not None and (best is None or r < best):
                        best, at = r, i
                if best is None:
                    break
                parts[at:at + 2] = [parts[at] + parts[at + 1]]
            out.extend(self.ranks[p] for p in parts)
        return out


class Whisper:
    """Mirror of `struct Whisper` (Whisper.swift:11-41). `Whisper()` there loads the two CoreML packages with real
    `small` weights; here the model size is a parameter and weights come from a state dict (upstream key names) or are
    seeded synthetic (no checkpoint exists offline)."""

    LANGUAGES = LANGUAGES

    def __init__(self, model: str | ModelDims = "small", weights: Optional[Dict[str, np.ndarray]] = None,
                 seed: int = 0, max_batch: int = 1, max_beams: int = 1, device: int = 0, stream: int = 0):
        self.dims = DIMS[model] if isinstance(model, str) else model
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        d = _Dims(*[getattr(self.dims, f[0]) for f in _Dims._fields_])
        _check(self._lib.wb_create(ctypes.byref(d), max_batch, max_beams, device, ctypes.c_void_p(stream),
                                   ctypes.byref(self._h)), "wb_create")
        self.max_batch = max_batch
        if weights is not None:
            self.load_state_dict(weights)
        elif seed is not None:
            _check(self._lib.wb_init_random_weights(self._h, seed), "wb_init_random_weights")

    @classmethod
    def from_checkpoint(cls, path: str, **kw) -> "Whisper":
        """A model from a checkpoint file (whisper_to_cml.py:7 loads one by name from the network). `.safetensors` (upstream or
        transformers tensor names; F32 / F16 / BF16) is read natively by the C ABI; an upstream `.pt` ({"dims", "model_state_dict"})
        is unpickled with torch and handed over as a state dict."""
        if path.endswith(".safetensors"):
            w = cls(read_checkpoint_dims(path), seed=None, **kw)
            n = ctypes.c_int32(0)
            _check(w._lib.wb_load_safetensors(w._h, path.encode(), ctypes.byref(n)), "wb_load_safetensors")
            w.n_loaded = n.value
            return w
        import torch
        ck = torch.load(path, map_location="cpu", weights_only=True)
        sd = ck["model_state_dict"] if "model_state_dict" in ck else ck
        d = ck.get("dims") if isinstance(ck, dict) else None
        dims = ModelDims(**{k: int(v) for k, v in d.items()}) if d else None
        if dims is None:
            raise WhisperB200Error(f"{path}: no 'dims' entry; save as safetensors (dimensions are derived from the shapes)")
        return cls(dims, weights=sd, **kw)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.wb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------------------------
    def load_state_dict(self, weights: Dict[str, object]) -> None:
        for name, t in weights.items():
            a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
            a = np.ascontiguousarray(a, dtype=np.float32)
            _check(self._lib.wb_set_weight(self._h, name.encode(), _ptr(a), a.size), f"wb_set_weight({name})")
        _check(self._lib.wb_weights_commit(self._h), "wb_weights_commit")

    def load_hf_state_dict(self, weights: Dict[str, object]) -> None:
        """Same, from a transformers Whisper checkpoint (safetensors / state_dict key names)."""
        mapped = {}
        for k, t in weights.items():
            u = hf_to_upstream_name(k)
            if u is not None:
                mapped[u] = t
        self.load_state_dict(mapped)

    def weight_arena(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.wb_weight_arena(self._h, ctypes.byref(p), ctypes.byref(n)), "wb_weight_arena")
        return p.value, n.value

    def mark_weights_loaded(self) -> None:
        _check(self._lib.wb_weights_mark_loaded(self._h), "wb_weights_mark_loaded")

    def weight_names(self) -> Dict[str, int]:
        """Upstream state-dict name -> element count of every tensor of the model."""
        out: Dict[str, int] = {}
        buf = ctypes.create_string_buffer(256)
        n = ctypes.c_size_t()
        for i in range(int(self._lib.wb_weight_count(self._h))):
            _check(self._lib.wb_weight_info(self._h, i, buf, 256, ctypes.byref(n)), "wb_weight_info")
            out[buf.value.decode()] = int(n.value)
        return out

    def get_weight(self, name: str) -> np.ndarray:
        """Reads a tensor back as flat host fp32 in its upstream layout (fp16-stored tensors return their rounded values)."""
        n = self.weight_names()[name]
        a = np.empty(n, dtype=np.float32)
        _check(self._lib.wb_get_weight(self._h, name.encode(), _ptr(a), n), f"wb_get_weight({name})")
        return a

    def state_dict(self) -> Dict[str, np.ndarray]:
        """Every tensor as the device holds it (flat fp32): makes device-generated weights visible to a checker."""
        return {k: self.get_weight(k) for k in self.weight_names()}

    def weights_checksum(self) -> int:
        v = ctypes.c_uint64()
        _check(self._lib.wb_weights_checksum(self._h, ctypes.byref(v)), "wb_weights_checksum")
        return int(v.value)

    # ---- Whisper.encode (Whisper.swift:23-31) ------------------------------------------------------------------------
    def _as_batch(self, audio) -> np.ndarray:
        a = np.asarray(audio)
        if a.ndim == 1:
            a = a[None, :]
        if a.shape[-1] != N_SAMPLES:
            raise ValueError(f"audio must have {N_SAMPLES} samples per chunk (got {a.shape[-1]}); see pad_or_trim")
        return np.ascontiguousarray(a, dtype=np.float32)

    def logmel(self, audio) -> np.ndarray:
        a = self._as_batch(audio)
        out = np.empty((a.shape[0], N_MELS, N_FRAMES), dtype=np.float32)
        _check(self._lib.wb_logmel(self._h, _ptr(a), a.shape[0], _ptr(out)), "wb_logmel")
        return out

    def encode(self, audio, return_features: bool = True) -> Optional[np.ndarray]:
        """audio: [480000] (or [B,480000]) samples -> audio features [B,1500,d] f32 (kept resident for decode)."""
        a = self._as_batch(audio)
        B = a.shape[0]
        out = np.empty((B, self.dims.n_audio_ctx, self.dims.n_audio_state), dtype=np.float32) if return_features else None
        _check(self._lib.wb_encode(self._h, _ptr(a), B, _ptr(out) if out is not None else None), "wb_encode")
        return out

    def encode_mel(self, mel) -> np.ndarray:
        """`encoderModel.prediction(x_1:)` alone (Whisper.swift:29): mel [B,80,3000] f32 -> [B,1500,d]."""
        m = np.ascontiguousarray(np.asarray(mel, dtype=np.float32).reshape(-1, N_MELS, N_FRAMES))
        out = np.empty((m.shape[0], self.dims.n_audio_ctx, self.dims.n_audio_state), dtype=np.float32)
        _check(self._lib.wb_encode_mel(self._h, _ptr(m), m.shape[0], _ptr(out)), "wb_encode_mel")
        return out

    # ---- decoder.prediction / Whisper.decode (Whisper.swift:33-40) --------------------------------------------------------
    def decoder_logits(self, tokens, audio_features=None) -> np.ndarray:
        """`decoderModel.prediction(x_1: tokens, xa: audioFeatures).var_2217`: tokens [B,t] -> logits [B,t,V] f32."""
        if audio_features is not None:
            self.set_audio_features(audio_features)
        tk = np.ascontiguousarray(np.asarray(tokens).reshape(np.asarray(tokens).shape[0], -1), dtype=np.int32)
        B, t = tk.shape
        out = np.empty((B, t, self.dims.n_vocab), dtype=np.float32)
        _check(self._lib.wb_decoder_logits(self._h, _ptr(tk), B, t, _ptr(out)), "wb_decoder_logits")
        return out

    def set_audio_features(self, xa) -> None:
        x = np.ascontiguousarray(np.asarray(xa, dtype=np.float32).reshape(-1, self.dims.n_audio_ctx, self.dims.n_audio_state))
        _check(self._lib.wb_set_audio_features(self._h, _ptr(x), x.shape[0]), "wb_set_audio_features")

    def detect_language(self, B: int = 1, sot: int = 50258, lang0: int = 50259) -> np.ndarray:
        out = np.empty(B, dtype=np.int32)
        _check(self._lib.wb_detect_language(self._h, B, sot, lang0, _ptr(out)), "wb_detect_language")
        return out

    def decode(self, audioFeatures=None, quiet: bool = False) -> List[str]:
        """Whisper.decode(audioFeatures:): one decoder call on [50258], arg-max over logits 50259...50357, print the
        language code (Whisper.swift:34-39). Returns the codes as well (the Swift returns Void)."""
        B = 1
        if audioFeatures is not None:
            x = np.asarray(audioFeatures, dtype=np.float32).reshape(-1, self.dims.n_audio_ctx, self.dims.n_audio_state)
            B = x.shape[0]
            self.set_audio_features(x)
        idx = self.detect_language(B)
        codes = [LANGUAGES[i] for i in idx]
        if not quiet:
            for c in codes:
                print(c)
        return codes

    # ---- transcribe (north-star extension) ----------------------------------------------------------------------------
    def _opts(self, o: DecodeOptions):
        init = np.asarray(list(o.initial_tokens), dtype=np.int32)
        sup = np.asarray(list(o.suppress), dtype=np.int32)
        supb = np.asarray(list(o.suppress_begin), dtype=np.int32)
        c = _DecodeOpts(init.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), init.size, o.sample_len, o.eot,
                        sup.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), sup.size,
                        supb.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), supb.size, o.beam_size, o.eot_check_interval,
                        1 if o.timestamps else 0, o.timestamp_begin, o.no_timestamps, o.max_initial_timestamp_index,
                        float(o.temperature), int(o.best_of), int(o.seed) & 0xFFFFFFFFFFFFFFFF, int(o.no_speech), int(o.sot_index),
                        ctypes.POINTER(ctypes.c_float)())
        return c, (init, sup, supb)

    def decode_tokens(self, B: int, opts: DecodeOptions):
        """wb_decode on the resident features: greedy (beam_size <= 1) or beam search (beam_size > 1; per chunk the best
        candidate by sum_logprob / length). Returns (tokens [B, n_init+sample_len], lens [B], sum_logprob [B])."""
        c, keep = self._opts(opts)
        total = len(opts.initial_tokens) + opts.sample_len
        tokens = np.empty((B, total), dtype=np.int32)
        lens = np.empty(B, dtype=np.int32)
        slp = np.empty(B, dtype=np.float32)
        _check(self._lib.wb_decode(self._h, B, ctypes.byref(c), _ptr(tokens), _ptr(lens), _ptr(slp)), "wb_decode")
        return tokens, lens, slp

    def decode_with_no_speech(self, B: int, opts: DecodeOptions):
        """decode_tokens plus upstream's no_speech_probs: softmax of the unfiltered logits at opts.sot_index, at opts.no_speech.
        Returns (tokens, lens, sum_logprob, no_speech_prob [B])."""
        c, keep = self._opts(opts)
        nsp = np.zeros(B, dtype=np.float32)
        c.no_speech_prob = nsp.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        total = len(opts.initial_tokens) + opts.sample_len
        tokens = np.empty((B, total), dtype=np.int32)
        lens = np.empty(B, dtype=np.int32)
        slp = np.empty(B, dtype=np.float32)
        _check(self._lib.wb_decode(self._h, B, ctypes.byref(c), _ptr(tokens), _ptr(lens), _ptr(slp)), "wb_decode")
        return tokens, lens, slp, nsp

    def greedy(self, B: int, opts: DecodeOptions):
        """Greedy decode of the resident features (rejects beam options: use decode_tokens / beam_search for those)."""
        if opts.beam_size > 1:
            raise ValueError("greedy() was given beam_size > 1; use decode_tokens() or beam_search()")
        return self.decode_tokens(B, opts)

    def beam_search(self, B: int, opts: DecodeOptions, beam_size: int = 5):
        """Beam-search decode (upstream BeamSearchDecoder, patience 1); needs max_beams >= beam_size at construction."""
        import dataclasses
        return self.decode_tokens(B, dataclasses.replace(opts, beam_size=beam_size))

    def transcribe(self, audio, opts: Optional[DecodeOptions] = None):
        a = self._as_batch(audio)
        opts = opts or DecodeOptions.default_for(self.dims)
        c, keep = self._opts(opts)
        B = a.shape[0]
        total = len(opts.initial_tokens) + opts.sample_len
        tokens = np.empty((B, total), dtype=np.int32)
        lens = np.empty(B, dtype=np.int32)
        slp = np.empty(B, dtype=np.float32)
        _check(self._lib.wb_transcribe(self._h, _ptr(a), B, ctypes.byref(c), _ptr(tokens), _ptr(lens), _ptr(slp)),
               "wb_transcribe")
        return tokens, lens, slp

    def transcribe_long(self, pcm, opts: Optional[DecodeOptions] = None):
        """A stream longer than 30 s as independent fixed windows, max_batch at a time: (tokens [n_windows, L], lens)."""
        win = split_windows(pcm)
        toks, lens = [], []
        for i in range(0, win.shape[0], self.max_batch):
            t, l, _ = self.transcribe(win[i:i + self.max_batch], opts)
            toks.append(t)
            lens.append(l)
        return np.concatenate(toks, axis=0), np.concatenate(lens, axis=0)

    def logmel_long(self, pcm, frame0: int = 0) -> np.ndarray:
        """upstream log_mel_spectrogram(pcm, padding=N_SAMPLES)[:, frame0:frame0 + 3000] -> [80, 3000] f32."""
        a = np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1))
        out = np.empty((N_MELS, N_FRAMES), dtype=np.float32)
        _check(self._lib.wb_logmel_long(self._h, _ptr(a), a.shape[0], frame0, _ptr(out)), "wb_logmel_long")
        return out

    def transcribe_seek(self, pcm, opts: Optional[DecodeOptions] = None, *, temperatures: Optional[Sequence[float]] = None,
                        compression_ratio_threshold: Optional[float] = 2.4, logprob_threshold: Optional[float] = -1.0,
                        no_speech_threshold: Optional[float] = 0.6, condition_on_previous_text: bool = True,
                        initial_prompt: Optional[Sequence[int]] = None, tokenizer: Optional["Tokenizer"] = None,
                        detect_language: bool = False, max_segments: int = 0) -> dict:
        """Upstream whisper `transcribe()` on one recording of any length (wb_transcribe_long: seek loop, temperature fallback,
        thresholds, previous text as prompt). `opts` holds the per-window options (default: upstream's, timestamps on);
        a threshold of None turns its rule off. Returns {"tokens", "segments": [dict], "language": index or None}."""
        a = np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1))
        opts = opts or DecodeOptions.default_for(self.dims, without_timestamps=False)
        c, keep = self._opts(opts)
        lo = _LongOpts()
        lo.decode = c
        t = None
        if temperatures is not None:
            t = np.asarray(list(temperatures), dtype=np.float32)
            lo.temperatures = t.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            lo.n_temperatures = t.size
        nan = float("nan")
        lo.compression_ratio_threshold = nan if compression_ratio_threshold is None else compression_ratio_threshold
        lo.logprob_threshold = nan if logprob_threshold is None else logprob_threshold
        lo.no_speech_threshold = nan if no_speech_threshold is None else no_speech_threshold
        lo.condition_on_previous_text = 1 if condition_on_previous_text else 0
        ip = np.asarray(list(initial_prompt or []), dtype=np.int32)
        if ip.size:
            lo.initial_prompt = ip.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
            lo.n_initial_prompt = ip.size
        lo.sot_prev = opts.sot_prev
        lo.tokenizer = tokenizer.handle if tokenizer is not None else None
        lo.detect_language = 1 if detect_language else 0
        lo.lang0 = 50259
        lang = ctypes.c_int32(-1)
        lo.detected_language = ctypes.pointer(lang)
        n_windows = a.shape[0] // N_SAMPLES + 2
        seg_cap = max_segments or 64 * n_windows
        tok_cap = 256 * n_windows + 256
        segs = (_Segment * seg_cap)()
        tokens = np.empty(tok_cap, dtype=np.int32)
        n_seg, n_tok = ctypes.c_int32(0), ctypes.c_int32(0)
        _check(self._lib.wb_transcribe_long(self._h, _ptr(a), a.shape[0], ctypes.byref(lo), segs, seg_cap, ctypes.byref(n_seg),
                                            _ptr(tokens), tok_cap, ctypes.byref(n_tok)), "wb_transcribe_long")
        out = []
        for i in range(n_seg.value):
            g = segs[i]
            out.append({"id": i, "seek": g.seek, "start": g.start, "end": g.end,
                        "tokens": tokens[g.token_begin:g.token_begin + g.n_tokens].tolist(), "temperature": g.temperature,
                        "avg_logprob": g.avg_logprob, "compression_ratio": g.compression_ratio, "no_speech_prob": g.no_speech_prob})
            if tokenizer is not None:
                out[-1]["text"] = tokenizer.decode(out[-1]["tokens"], drop_from=opts.eot)
        res = {"tokens": tokens[:n_tok.value].copy(), "segments": out, "language": lang.value if lang.value >= 0 else None}
        if tokenizer is not None:
            res["text"] = tokenizer.decode(res["tokens"].tolist(), drop_from=opts.eot)
        return res

    def transcribe_dev(self, audio_dev_ptr: int, B: int, opts: DecodeOptions):
        c, keep = self._opts(opts)
        total = len(opts.initial_tokens) + opts.sample_len
        tokens = np.empty((B, total), dtype=np.int32)
        lens = np.empty(B, dtype=np.int32)
        slp = np.empty(B, dtype=np.float32)
        _check(self._lib.wb_transcribe_dev(self._h, ctypes.c_void_p(audio_dev_ptr), B, ctypes.byref(c), _ptr(tokens),
                                           _ptr(lens), _ptr(slp)), "wb_transcribe_dev")
        return tokens, lens, slp

    # ---- introspection ------------------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._lib.wb_launch_count(self._h))

    def last_timings(self) -> np.ndarray:
        t = np.zeros(4, dtype=np.float32)
        _check(self._lib.wb_last_timings(self._h, _ptr(t)), "wb_last_timings")
        return t

    def profile_cross_attention(self, B: int, reps: int = 60):
        """(mean ms per launch, algorithmic bytes per launch) of the KV-cache attention kernel on the resident cross K/V."""
        ms, by = ctypes.c_float(), ctypes.c_double()
        _check(self._lib.wb_profile_cross_attention(self._h, B, reps, ctypes.byref(ms), ctypes.byref(by)),
               "wb_profile_cross_attention")
        return float(ms.value), float(by.value)

    def sync(self) -> None:
        _check(self._lib.wb_sync(self._h), "wb_sync")

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h
