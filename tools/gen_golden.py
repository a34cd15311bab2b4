#!/usr/bin/env python3
"""Generates tests/golden/*.npz — known-answer vectors from implementations INDEPENDENT of oracle/ and of the CUDA code.

The reference holds no test vectors for this path and neither of its runtimes exists in this image
(no cargo/rustc for stft/, no `whisper`/coremltools for the models). The independent implementations available here:
  * log-mel: torch.stft (f64, periodic Hann, center=True reflect) + the reference's own mel fixture m80.npy —
    the formulation upstream whisper/audio.py uses and the Rust crate re-implements (stft/src/lib.rs:22-102)
  * encoder/decoder: transformers.WhisperForConditionalGeneration holding the seeded weights of
    oracle/whisper_ref.random_weights (same arithmetic as upstream whisper, different code base)
Run in this container: python tools/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import whisper_ref as ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SMALL = ref.ModelDims(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2)       # d_head = 64 like every Whisper size
SMALL_ML = ref.ModelDims(80, 1500, 128, 2, 2, 51865, 448, 128, 2, 2)


def logmel_torch_f64(audio: np.ndarray) -> np.ndarray:
    mel = torch.from_numpy(np.load(os.path.join(GOLD, "m80.npy")).reshape(80, 201).astype(np.float64))
    x = torch.from_numpy(audio.astype(np.float64))
    st = torch.stft(x, 400, 160, window=torch.hann_window(400, dtype=torch.float64), return_complex=True)   # center, reflect
    mag = st[..., :-1].abs() ** 2
    spec = torch.clamp(mel @ mag, min=1e-10).log10()
    spec = torch.maximum(spec, spec.max() - 8.0)
    return ((spec + 4.0) / 4.0).numpy()


def main():
    rng = np.random.default_rng(1234)
    idx = np.sort(rng.choice(240000, 4096, replace=False))
    out = {"index": idx}
    for kind in ["noise", "sine", "chirp", "noise_then_zeros", "int16", "fullscale", "zeros"]:
        a = ref.synth_audio(11, kind)
        out[kind] = logmel_torch_f64(a).reshape(-1)[idx]
    np.savez_compressed(os.path.join(GOLD, "logmel_torch_f64.npz"), **out)

    torch.manual_seed(0)
    for tag, dims in (("small_en", SMALL), ("small_ml", SMALL_ML)):
        w = ref.random_weights(dims, seed=3)
        hf = ref.to_hf(dims, w)
        audio = ref.synth_audio(21, "noise")
        mel = torch.from_numpy(logmel_torch_f64(audio)).float()[None]
        with torch.no_grad():
            xa = hf.model.encoder(mel).last_hidden_state
            toks = torch.tensor([[50257, 50362, 1000, 2000, 3000, 40000]]) if not dims.is_multilingual else torch.tensor(
                [[50258, 50259, 50359, 50363, 1000, 2000]])
            logits = hf(input_features=mel, decoder_input_ids=toks).logits
            # greedy, no filters, 12 steps, via HF forward
            seq = toks[:, :2].clone()
            for _ in range(12):
                nxt = hf(input_features=mel, decoder_input_ids=seq).logits[:, -1].argmax(-1)
                seq = torch.cat([seq, nxt[:, None]], dim=1)
        cols = np.concatenate([np.arange(0, 51864, 97), np.arange(50250, dims.n_vocab)])
        np.savez_compressed(os.path.join(GOLD, f"whisper_{tag}_hf.npz"), xa_rows=xa[0, ::75].numpy(), tokens=toks.numpy(),
                            logit_cols=cols, logits=logits[0][:, cols].numpy(), greedy=seq.numpy(),
                            lang=np.array([int(logits[0, 0, 50259:50358].argmax())]) if dims.is_multilingual else np.array([-1]))
        print(tag, "xa", tuple(xa.shape), "greedy", seq.tolist())


if __name__ == "__main__":
    main()
