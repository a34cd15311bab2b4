#!/usr/bin/env python3
"""Development: merged %globaltimer timeline of the sub-batch streams of one decode step (WB_TRACE=1, WB_SUBBATCHES=n).
Every record is block (0,0) of a kernel: id, entry, exit. Prints the kernels of all sub-batches inside the window of
sub-batch 0's last full step, ordered by entry time."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WB_TRACE"] = "1"
wbm = importlib.import_module("openai-whisper-coreml_b200")
model = sys.argv[1] if len(sys.argv) > 1 else "base.en"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
nsb = int(os.environ.get("WB_SUBBATCHES", "1"))
w = wbm.Whisper(model, seed=0, max_batch=B)
o = wbm.DecodeOptions.default_for(wbm.DIMS[model], sample_len=steps)
o.suppress = list(o.suppress) + [o.eot]
audio = (np.random.default_rng(0).standard_normal((B, 480000)) * 0.1).astype(np.float32)
w.transcribe(audio, o)
w.transcribe(audio, o)
lib = wbm.load_library()
lib.wb_debug_trace_sub.restype = ctypes.c_int
names = {210: "self_block", 220: "post_block", 230: "layer_block", 200: "self_attn", 201: "cross_attn", 202: "cross_stream",
         300: "finish", 301: "finish+sample", 141: "logits(LN)", 142: "logits_tc"}
recs = []
for s in range(nsb):
    buf = np.zeros((16384, 8), dtype=np.uint64)
    n = lib.wb_debug_trace_sub(w.handle, s, buf.ctypes.data_as(ctypes.c_void_p), 16384)
    recs.append(buf[:n].astype(np.int64))
r0 = recs[0]
fin = [i for i in range(len(r0)) if r0[i, 0] in (300, 301)]
t0, t1 = int(r0[fin[-3], 2]), int(r0[fin[-1], 2])   # two steps of sub-batch 0
rows = []
for s, r in enumerate(recs):
    for i in range(len(r)):
        if t0 <= r[i, 1] <= t1:
            rows.append((int(r[i, 1]), int(r[i, 2]), s, int(r[i, 0])))
rows.sort()
print(f"window {(t1 - t0) / 1e3:.1f} us = two steps of sub-batch 0; {len(rows)} kernels of {nsb} sub-batches")
for st, en, s, kid in rows:
    print(f"{'    ' * s}sub{s} {names.get(kid, kid):14s} {(st - t0) / 1e3:8.2f} -> {(en - t0) / 1e3:8.2f}  ({(en - st) / 1e3:6.2f})")
w.close()
