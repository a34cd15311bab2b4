"""Two GPUs, one process per GPU under torchrun: the load-time weight broadcast, the rank verification of
bench.py (arena checksums all-gathered, rank r re-decodes rank 0's first chunk) and the sharded job of BASELINE configs[4]
(60 windows partitioned over the ranks, token rows all-gathered, compared with a single-GPU run). Skipped on one-GPU boxes."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_bench_verifies_broadcast_and_sharded_job():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "3",
           "--configs", "large_v2_60w", "--extra-steps", "1", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]          # rank 0 alone prints
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["ranks_verified"] == 2 and line["weights_broadcast_bytes"] > 0
    assert line["scaling"] == "weak" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 32 * 480000 * 4
    job = line["extra_configs"]["large_v2_60w"]
    assert job["windows_per_rank"] == [30, 30] and job["tokens_identical_to_single_gpu_run"] is True
