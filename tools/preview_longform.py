#!/usr/bin/env python3
"""Development (CPU): runs the ORACLE's long-form loop on the test scenario and prints which branches each window took,
to pick scenario constants that cover the loop before GPU time is spent."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import whisper_ref as ref
import longform_util as lu
import torch
torch.set_num_threads(8)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 75.0
kw = dict(a.split("=") for a in sys.argv[2:])
dims, v, weights = lu.scenario(ref, **{k: float(x) for k, x in kw.items() if k in ("text_scale", "eot_scale", "ns_scale")})
model = ref.WhisperRef(dims, weights)
mel_filters = np.load(os.path.join(ROOT, "tests", "golden", "m80.npy"))
audio = lu.recording(secs)
opts = ref.DecodeOptions.default_for(dims, sample_len=int(kw.get("sample_len", 40)), without_timestamps=False)
table = lu.synthetic_table(v.eot)
t0 = time.time()
toks, segs, trace = ref.transcribe_seek(model, audio, mel_filters, opts, temperatures=(0.0, 0.6, 1.0), logprob_threshold=float(kw.get("lp", -9.0)),
                                        compression_ratio_threshold=float(kw.get("cr", 1.35)), no_speech_threshold=float(kw.get("ns", 0.5)),
                                        table=table, seed=3, best_of=int(kw.get("best_of", 2)))
print(f"{time.time() - t0:.1f} s; {len(toks)} tokens, {len(segs)} segments")
for seek, tried, res in trace:
    print(f"seek {seek:6d} temps {tried} n_tok {len(res.tokens):3d} avg_lp {res.avg_logprob:7.3f} nsp {res.no_speech_prob:.3f} cr {res.compression_ratio:.2f} "
          f"ts {[t - v.timestamp_begin for t in res.tokens if t >= v.timestamp_begin][:8]}")
for s in segs:
    print(f"  seg seek {s['seek']} {s['start']:.2f}-{s['end']:.2f} n {len(s['tokens'])} T {s['temperature']}")
