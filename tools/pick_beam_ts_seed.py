#!/usr/bin/env python3
"""Development (CPU): a (seed, text_scale) for tests/test_gpu_parity.py::test_beam_search_with_timestamp_rules whose oracle result
is stable under logit noise of the size of the stated tolerance and samples both token classes (see pick_beam_seed.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import whisper_ref as ref  # noqa: E402

B, BEAM, STEPS, NOISE = 2, 3, 12, 2e-2
dims = ref.DIMS["tiny"]
v = ref.Vocab.for_dims(dims)
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 0, 60):
    weights = ref.random_weights(dims, seed=seed)
    weights["decoder.token_embedding.weight"][:v.eot] *= 1.8
    oracle = ref.WhisperRef(dims, weights)
    xa = (torch.randn(B, 1500, dims.n_audio_state, generator=torch.Generator().manual_seed(200 + seed)) * 0.7).half().float()
    opts = ref.DecodeOptions.default_for(dims, sample_len=STEPS, without_timestamps=False)
    base, scores = oracle.beam_search(xa, opts, beam_size=BEAM)
    body = [t[len(opts.initial_tokens):] for t in base]
    mixed = all(any(x >= v.timestamp_begin for x in t) and any(x < v.eot for x in t) for t in body)
    plain = oracle.decoder_logits
    ok = mixed
    for trial in range(3 if ok else 0):
        g = torch.Generator().manual_seed(1000 * seed + trial)

        def noisy(*a, **k):
            out = plain(*a, **k)
            return out + torch.randn(out.shape, generator=g) * NOISE

        oracle.decoder_logits = noisy
        got, _ = oracle.beam_search(xa, opts, beam_size=BEAM)
        oracle.decoder_logits = plain
        if got != base:
            ok = False
            break
    print(f"seed {seed}: mixed={mixed} {'robust' if ok else 'fragile'} {body}", flush=True)
    if ok:
        break
