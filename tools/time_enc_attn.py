#!/usr/bin/env python3
"""Development: wall-clock time of the encoder attention kernel alone (no profiler): B x H x T at base.en's shape."""
import ctypes, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")
B, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 1500, int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = H * 64
w = wbm.Whisper("tiny.en", seed=0, max_batch=1)
lib = wbm.load_library()
qkv = (torch.randn(B * T, 3 * d, device="cuda") * 1.5).half()
o = torch.empty((B * T, d), dtype=torch.float16, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
torch.cuda.synchronize()
for reps in (3, 20):
    t0 = time.perf_counter()
    for _ in range(reps):
        assert lib.wb_op_attention(w.handle, p(qkv), B, T, H, p(o)) == 0
    w.sync()
    dt = (time.perf_counter() - t0) / reps
flop = 4.0 * B * H * T * T * 64
print(f"encoder attention B={B} H={H}: {dt * 1e6:.1f} us per launch, {flop / dt / 1e12:.0f} TFLOP/s")
w.close()
