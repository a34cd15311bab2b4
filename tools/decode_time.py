#!/usr/bin/env python3
"""Development: device time of the decode phase (wb_last_timings) of a greedy transcribe, for A/B runs of environment switches.
usage: python tools/decode_time.py [model=base.en] [chunks=32] [sample_len=224] [reps=4]; prints a digest of the tokens too."""
import ctypes, hashlib, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")
model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 224
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
w = wbm.Whisper(model, seed=0, max_batch=B)
o = wbm.DecodeOptions.default_for(wbm.DIMS[model], sample_len=n)
o.suppress = list(o.suppress) + [o.eot]
audio = (np.random.default_rng(0).standard_normal((B, 480000)) * 0.1).astype(np.float32)
lib = wbm.load_library()
t = (ctypes.c_float * 4)()
dec = []
for i in range(2 + reps):
    r = w.transcribe(audio, o)
    lib.wb_last_timings(w.handle, t)
    if i >= 2:
        dec.append(t[2])
toks = np.asarray(r[0])
print(f"decode ms min {min(dec):.3f} avg {sum(dec) / len(dec):.3f}  ({min(dec) * 1e3 / t[3]:.1f} us/step over {t[3]:.0f} steps)  tokens sha {hashlib.sha256(toks.tobytes()).hexdigest()[:12]}")
w.close()
