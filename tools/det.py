import importlib, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wbm = importlib.import_module("openai-whisper-coreml_b200")
w = wbm.Whisper("base.en", seed=0, max_batch=32)
o = wbm.DecodeOptions.default_for(wbm.DIMS["base.en"], sample_len=224)
o.suppress = list(o.suppress) + [o.eot]
audio = np.stack([(np.random.default_rng(1000 + i).standard_normal(480000) * 0.1).astype(np.float32) for i in range(32)])
runs = [w.transcribe(audio, o) for _ in range(4)]
for r in runs[1:]:
    neq = (r[0] != runs[0][0])
    print("token mismatches", int(neq.sum()), "slp equal", bool(np.array_equal(r[2], runs[0][2])), "max token", int(r[0].max()))
