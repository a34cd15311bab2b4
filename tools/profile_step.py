#!/usr/bin/env python3
"""Short, profiler-friendly run of the hot path: one warm pass (graph capture), then — between cudaProfilerStart/Stop —
one pass of log-mel + encoder + cross K/V + a few decode steps at the bench configuration (base.en, 32 chunks).
Use under `ncu --profile-from-start off ...`."""
import importlib
import os
import sys

import numpy as np
import torch

os.environ.setdefault("WB_H2D_SLABS", "0")   # one copy, one encoder pass over the whole batch: kernel shapes as DESIGN.md quotes them

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")

model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
beam = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = wbm.Whisper(model, seed=0, max_batch=B, max_beams=max(beam, 1))
o = wbm.DecodeOptions.default_for(wbm.DIMS[model], sample_len=steps)
o.beam_size = beam
o.suppress = list(o.suppress) + [o.eot]
audio = (np.random.default_rng(0).standard_normal((B, 480000)) * 0.1).astype(np.float32)
w.transcribe(audio, o)
torch.cuda.synchronize()
torch.cuda.profiler.start()
w.transcribe(audio, o)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", w.launch_count(), "timings", w.last_timings().tolist())
w.close()
