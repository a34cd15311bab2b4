#!/usr/bin/env python3
"""Development: per-launch time of the KV-cache attention kernel alone for several batch sizes (CTAs = 8 x B at base.en):
how much HBM bandwidth a given number of resident CTAs pulls with the ring depth WB_HA_STAGES."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")
model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
w = wbm.Whisper(model, seed=0, max_batch=32)
audio = (np.random.default_rng(0).standard_normal((32, 480000)) * 0.1).astype(np.float32)
w.encode(audio, return_features=False)
for B in (1, 2, 4, 8, 12, 16, 18, 24, 32):
    ms, by = w.profile_cross_attention(B, 240)
    print(f"stages={os.environ.get('WB_HA_STAGES', 'default')} B={B:2d} ctas={B * w.dims.n_text_head:3d} us={ms * 1e3:7.2f} GB/s={by / ms / 1e6:8.1f} per-CTA GB/s={by / ms / 1e6 / (B * w.dims.n_text_head):6.1f}")
w.close()
