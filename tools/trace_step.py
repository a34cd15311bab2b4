#!/usr/bin/env python3
"""Development: per-kernel %globaltimer timeline of decode steps (WB_TRACE=1 must be set)."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WB_TRACE"] = "1"
wbm = importlib.import_module("openai-whisper-coreml_b200")
model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
w = wbm.Whisper(model, seed=0, max_batch=B)
o = wbm.DecodeOptions.default_for(wbm.DIMS[model], sample_len=steps)
o.suppress = list(o.suppress) + [o.eot]
audio = (np.random.default_rng(0).standard_normal((B, 480000)) * 0.1).astype(np.float32)
w.transcribe(audio, o)
w.transcribe(audio, o)
lib = wbm.load_library()
lib.wb_debug_trace.restype = ctypes.c_int
buf = np.zeros((65536, 8), dtype=np.uint64)
n = lib.wb_debug_trace(w.handle, buf.ctypes.data_as(ctypes.c_void_p), 65536)
rec = buf[:n]
names = {210: "self_block", 220: "post_block", 230: "layer_block", 200: "self_attn", 201: "cross_attn", 202: "cross_stream", 300: "finish", 301: "finish+sample", 131: "QKV(LN)", 120: "out/mlp2(f16,resid)", 111: "Q(LN,f32)",
         101: "mlp1(LN,f16)", 141: "logits(LN)", 142: "logits_tc", 150: "LN rows", 130: "QKV(f16)", 110: "Q(f16,f32)", 100: "mlp1(f16,f16)"}
# last full step: records between the last two finish kernels
fin = [i for i in range(n) if rec[i, 0] in (300, 301)]
s, e = fin[-2] + 1, fin[-1] + 1
t0 = int(rec[s, 1])
prev_end = t0
print(f"{n} records; last step has {e - s} kernels")
for i in range(s, e):
    kid, st, en = int(rec[i, 0]), int(rec[i, 1]), int(rec[i, 2])
    marks = " ".join(f"{(int(rec[i, k]) - t0) / 1e3:7.2f}" if rec[i, k] else "      -" for k in (3, 6, 7, 4, 5))
    print(f"{names.get(kid, kid):22s} start {((st - t0) / 1e3):8.2f} end {((en - t0) / 1e3):8.2f}  after_prev_end {((en - prev_end) / 1e3):6.2f} | wait/ld+sum/shfl/prologue/mainloop done at {marks}")
    prev_end = en
print("step total us", (int(rec[e - 1, 2]) - t0) / 1e3)
w.close()
