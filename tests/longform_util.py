"""Scenario shared by the long-form tests (GPU parity) and tools/preview_longform.py (oracle alone, CPU): seeded weights
shaped so that upstream's transcribe() loop takes every branch on a short synthetic recording.

With plain seeded weights the logits are close to uniform over 51864 tokens: no sequence ever ends, nothing is ever
"silence". Three rows of the token embedding are scaled up so that the events the loop reacts to do happen, at rates that
depend on the audio of the window: <|endoftext|> (sequences end after a handful of tokens), <|nospeech|> (some windows
are skipped) and the text rows (the timestamp probability-mass rule falls both ways)."""
from __future__ import annotations

from typing import List

import numpy as np


def synthetic_table(n: int) -> List[bytes]:
    """token id -> bytes for ids below eot: short lower-case words, some with a leading space, a few non-ASCII and a few
    that are only white space, so that strip(), the UTF-8 replacement rule and the compression ratio all matter."""
    out = []
    for i in range(n):
        if i % 53 == 0:
            out.append(b" ")
        elif i % 41 == 0:
            out.append("é".encode("utf-8")[: 1 + (i // 41) % 2])       # sometimes half a code point
        else:
            out.append((b" " if i % 3 == 0 else b"") + bytes([97 + i % 26]) * (1 + i % 4))
    return out


def scenario(ref, name: str = "tiny.en", seed: int = 31, text_scale: float = 1.8, eot_scale: float = -2.2, ns_scale: float = 14.0):
    dims = ref.DIMS[name]
    v = ref.Vocab.for_dims(dims)
    weights = ref.random_weights(dims, seed=seed)
    E = weights["decoder.token_embedding.weight"]
    E[:v.eot] *= text_scale
    E[v.eot] *= eot_scale
    E[v.no_speech] *= ns_scale
    weights["decoder.token_embedding.weight"] = E.half().float()
    return dims, v, weights


def recording(seconds: float, seed: int = 5) -> np.ndarray:
    """Noise bursts of varying level with silent stretches: the encoder output, and with it every decision, changes from
    window to window."""
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    t = np.arange(n) / 16000.0
    env = 0.02 + 0.12 * (0.5 + 0.5 * np.sin(2 * np.pi * t / 7.3)) * (np.sin(2 * np.pi * t / 23.0) > -0.3)
    return (rng.standard_normal(n) * env).astype(np.float32)
