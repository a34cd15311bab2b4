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
                ("max_initial_timestamp_index", ctypes.c_int32)]


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
                             timestamp_begin=no_ts + 1, no_timestamps=no_ts)


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
    "wb_decoder_logits": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_decoder_logits_f32tok": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_detect_language": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "wb_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts), ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]),
    "wb_transcribe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts),
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wb_transcribe_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_DecodeOpts),
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
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
                        1 if o.timestamps else 0, o.timestamp_begin, o.no_timestamps, o.max_initial_timestamp_index)
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
