"""whisper_b200 — B200-native replacement for the hot path of tanmayb123/OpenAI-Whisper-CoreML.

Host-side mirror (Python, over the C ABI in include/whisper_b200.h) of the reference's Swift surface:
    generateSpectrogram(audio:)      Whisper/Whisper/stft.swift:8-19
    struct Whisper { init, encode(audio:), decode(audioFeatures:) }    Whisper/Whisper/Whisper.swift:11-41
plus the north-star extension Whisper.transcribe (greedy decode over a persistent KV cache).

All compute happens in libwhisper_b200.so (hand-written sm_100a CUDA). There is no CPU fallback: importing works
anywhere, but every compute call raises if the library is missing or no CUDA device is present.
"""
from .whisper import (DIMS, LANGUAGES, DecodeOptions, ModelDims, Tokenizer, Whisper, WhisperB200Error, bytes_to_unicode,
                      generateSpectrogram, hf_to_upstream_name, library_path, load_library, pad_or_trim, read_checkpoint_dims, split_windows)

__all__ = ["DIMS", "LANGUAGES", "DecodeOptions", "ModelDims", "Tokenizer", "Whisper", "WhisperB200Error", "bytes_to_unicode", "generateSpectrogram",
           "hf_to_upstream_name", "library_path", "load_library", "pad_or_trim", "read_checkpoint_dims", "split_windows"]
