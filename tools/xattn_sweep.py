#!/usr/bin/env python3
"""Times the KV-cache attention kernel alone (wb_profile_cross_attention) at the bench configuration."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")
model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
w = wbm.Whisper(model, seed=0, max_batch=B)
w.encode((np.random.default_rng(0).standard_normal((B, 480000)) * 0.1).astype(np.float32), return_features=False)
ms, by = w.profile_cross_attention(B, 240)
print(f"{model} B={B} env={ {k: v for k, v in os.environ.items() if k.startswith('WB_')} }: {ms*1e3:.2f} us/launch  {by/ms/1e6:.0f} GB/s  frac={by/ms/1e6/6541.5:.3f}")
w.close()
