"""Times the log-mel front end alone (wb_logmel_dev, device PCM in, [B][80][3000] f32 out) and prints a digest of the output, so
that compile-time variants of the kernel (WHISPER_B200_LIB=<variant .so>) can be compared for speed and bit identity.
usage: python tools/logmel_time.py [chunks=32] [reps=50]"""
import hashlib
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wbm = importlib.import_module("openai-whisper-coreml_b200")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    w = wbm.Whisper("tiny.en", seed=0, max_batch=B, device=0, stream=stream.cuda_stream)
    g = torch.Generator().manual_seed(5)
    audio = (torch.randn(B, 480000, generator=g) * 0.1).to(dev)
    out = torch.empty((B, 80, 3000), dtype=torch.float32, device=dev)
    lib = wbm.load_library()
    for _ in range(5):
        assert lib.wb_logmel_dev(w.handle, audio.data_ptr(), B, out.data_ptr()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        lib.wb_logmel_dev(w.handle, audio.data_ptr(), B, out.data_ptr())
    e1.record(stream)
    e1.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    nbytes = B * (480000 * 4 + 80 * 3000 * 4)
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"{os.path.basename(wbm.library_path()):24s} chunks {B}  {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s  sha {digest}")
