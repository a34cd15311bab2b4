#!/usr/bin/env python3
"""Development: run-to-run determinism of a 36-sequence tiny.en greedy decode (ragged last group), with the decoded tokens
compared between repeated runs and between hand-off modes (env WB_HANDOFF_FLAGS is read once per process)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
model = sys.argv[2] if len(sys.argv) > 2 else "tiny.en"
w = wbm.Whisper(model, seed=2, max_batch=B)
d = wbm.DIMS[model].n_audio_state
xa = (np.random.default_rng(9).standard_normal((B, 1500, d)) * 0.7).astype(np.float32)
w.set_audio_features(xa)
o = wbm.DecodeOptions.default_for(wbm.DIMS[model], sample_len=40)
o.suppress = list(o.suppress) + [o.eot]
runs = [w.greedy(B, o) for _ in range(6)]
for i, (t, l, s) in enumerate(runs[1:], 1):
    same = np.array_equal(t, runs[0][0])
    bad = np.nonzero((t != runs[0][0]).any(1))[0].tolist()
    print(f"run {i}: tokens identical to run 0: {same}; differing sequences {bad}; max |d slp| {np.abs(s - runs[0][2]).max():.3e}")
np.save(os.path.join(ROOT, "gpurun_out", f"det36_{os.environ.get('WB_HANDOFF_FLAGS', '1')}.npy"), runs[0][0])
w.close()
