#!/usr/bin/env python3
"""bench.py — audio-seconds per second (RTF x) of the Whisper hot path on synthetic 16 kHz audio.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs headline|all|a,b,...]
  (N > 1: launched by torchrun, one rank per GPU; reads RANK / LOCAL_RANK / WORLD_SIZE.)

Headline (the line's `value`, BASELINE.json configs[2], the configuration the metric is quoted on): a step = one pass of
the hot path (log-mel -> encoder -> cross K/V -> 224-token greedy decode) over 32 synthetic 30 s chunks per GPU on
Whisper-base.en dimensions with seeded random weights. Weak scaling: 32 chunks per GPU.

Prints ONE JSON line (rank 0): value = device-resident-input throughput; e2e = the same through the C ABI with pinned
HOST audio (H2D inside the timed region), e2e_pageable = with an ordinary (pageable) host buffer; roofline = the decoder's
KV-cache attention kernel against the measured HBM peak; cpu_baseline = the oracle port (restatement of the Rust stft +
upstream PyTorch whisper) on the host cores; extra_configs = the other BASELINE.json configurations (tiny.en B=1, small
5-beam B=8, large-v2 60 windows sharded over the ranks) plus strong scaling and sample_len=100 variants of the headline,
each with its own ms_per_step, e2e and roofline kernel; ranks_verified = ranks whose weight arena checksum and re-decode
of rank 0's first chunk agree (N > 1).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "base.en"
BATCH = 32
SAMPLE_LEN = 224
METRIC = "audio-sec/sec (RTFx) Whisper-base.en 30 s chunks"
UNIT = "audio-s/s"

# BASELINE.json configs[1..4] made concrete. "base_b32" is the headline.
CONFIGS = {
    "tiny_b1": dict(model="tiny.en", batch=1, beam=0, sample_len=224, what="configs[1]: Whisper-tiny.en, one 30 s chunk, greedy, 1 GPU"),
    "base_b32": dict(model="base.en", batch=32, beam=0, sample_len=224, what="configs[2]: Whisper-base.en, 32 chunks per GPU, greedy"),
    "small_beam5_b8": dict(model="small", batch=8, beam=5, sample_len=224, what="configs[3]: Whisper-small multilingual, 5-beam, batch 8, 1 GPU"),
    "large_v2_60w": dict(model="large-v2", batch=40, beam=0, sample_len=224, windows=60,
                         what="configs[4]: Whisper-large-v2, 30 min = 60 x 30 s windows sharded over the ranks, greedy"),
    # not a BASELINE configuration: the headline model at the handle's largest batch (the latency-bound block kernels of a
    # decode step cost the same for 64 sequences as for 32)
    "base_b64": dict(model="base.en", batch=64, beam=0, sample_len=224,
                     what="beyond BASELINE: Whisper-base.en, 64 chunks per GPU (the handle's limit), greedy, 1 GPU"),
}


def workload_string(chunks: int) -> str:
    return (f"whisper-{MODEL} dims, batch={chunks} x 30 s chunks per GPU, greedy {SAMPLE_LEN} tokens (EOT suppressed so "
            "every sequence decodes the full length), seeded random weights")


def _logmel_worker(args):
    import ctypes

    import numpy as np
    so, seed = args
    olib = ctypes.CDLL(so)
    a = (np.random.default_rng(seed).standard_normal(480000) * 0.1).astype(np.float32)
    mel = np.zeros((1, 80, 3000))
    t0 = time.perf_counter()
    olib.logmel_ref_batch_f32(a.ctypes.data_as(ctypes.c_void_p), 1, mel.ctypes.data_as(ctypes.c_void_p))
    return time.perf_counter() - t0


def cpu_logmel_legs(threads: int, n: int = 8):
    """the Rust crate's log-mel restated in C f64, one thread (as the crate runs) and one chunk per process
    on every host core. Returns chunks per second for both."""
    import multiprocessing as mp
    so = os.path.join(ROOT, "oracle", "liblogmel_ref.so")
    one = [_logmel_worker((so, 1000 + i)) for i in range(min(n, 4))]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(threads) as pool:
        pool.map(_logmel_worker, [(so, 1000 + i) for i in range(threads)])          # warm: process start-up, page-in
        t0 = time.perf_counter()
        pool.map(_logmel_worker, [(so, 2000 + i) for i in range(threads * 2)])
        dt = time.perf_counter() - t0
    return 1.0 / (sum(one) / len(one)), threads * 2 / dt


def cpu_reference_rtf(n_chunks: int, steps: int, warmup: int, threads: int, budget_s: float = 150.0):
    """Times the reference's CPU path — restated: C f64 log-mel (single thread, as the Rust crate) + PyTorch fp32
    KV-cached greedy whisper with `threads` threads — on n_chunks chunks per step. The one place oracle/ is executed
    as the thing measured (bench contract ④). Stops early (never before one timed step) once `budget_s` is spent, and
    returns how many warm-up / timed steps actually ran."""
    import ctypes

    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import whisper_ref as ref

    so = os.path.join(ROOT, "oracle", "liblogmel_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    olib = ctypes.CDLL(so)
    torch.set_num_threads(threads)
    dims = ref.DIMS[MODEL]
    model = ref.WhisperRef(dims, ref.random_weights(dims, seed=0))
    opts = ref.DecodeOptions.default_for(dims, sample_len=SAMPLE_LEN)
    opts.suppress = list(opts.suppress) + [model.vocab.eot]   # full 224 tokens, like the GPU arm
    audio = np.stack([ref.synth_audio(1000 + i, "noise") for i in range(n_chunks)]).astype(np.float32)
    times, ran_warm = [], 0
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        spent = time.perf_counter() - t_start
        if it < warmup and spent > budget_s / 3:
            continue                                            # skip the remaining warm-up passes
        if times and spent > budget_s:
            break
        t0 = time.perf_counter()
        mel = np.zeros((n_chunks, 80, 3000))
        assert olib.logmel_ref_batch_f32(audio.ctypes.data_as(ctypes.c_void_p), n_chunks, mel.ctypes.data_as(ctypes.c_void_p)) == 0
        xa = model.encode(torch.from_numpy(mel).float())
        model.greedy(xa, opts)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        else:
            ran_warm += 1
    sec = sum(times) / len(times)
    return n_chunks * 30.0 / sec, sec, len(times), ran_warm


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def synth_chunks(first_seed: int, n: int, pinned: bool = True):
    """Chunk c of a job is N(0, 0.1^2) seeded first_seed + c."""
    import numpy as np
    import torch
    host = torch.empty((n, 480000), dtype=torch.float32)
    if pinned:
        host = host.pin_memory()
    for i in range(n):
        g = np.random.default_rng(first_seed + i)
        host[i] = torch.from_numpy((g.standard_normal(480000) * 0.1).astype(np.float32))
    return host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-chunks", type=int, default=0, help="chunks per CPU pass (0: 32 for --impl reference, 4 for the cpu_baseline leg)")
    ap.add_argument("--configs", default="all", help="'headline', 'all' or a comma list of extra configurations: " + ",".join(CONFIGS))
    ap.add_argument("--extra-steps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    config = {"workload": workload_string(BATCH), "chunks_per_gpu": BATCH,
              "sample_len": SAMPLE_LEN, "cache": "per-step working set (590 MB cross K/V + weights) exceeds the 126 MB L2; no flush needed",
              "parallelism": f"dp{world} (chunk-sharded, no data-path collective)"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        chunks = args.cpu_chunks or BATCH                       # the stated configuration: 32 chunks per pass
        config["workload"] = workload_string(chunks)
        config["chunks_per_gpu"] = chunks
        rtf, sec, ran, ran_warm = cpu_reference_rtf(chunks, max(args.steps, 1), args.warmup, threads)
        lm1, lmn = cpu_logmel_legs(threads)
        line = {"impl": "reference", "metric": METRIC, "value": rtf, "unit": UNIT, "n_gpus": args.gpus, "steps": ran,
                "warmup": ran_warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rtf, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{chunks} chunks of the same workload per step, {ran} timed step(s) after {ran_warm} warm-up (steps are cut to "
                                           "a ~150 s budget, never the batch): C f64 log-mel restatement of the Rust stft, 1 thread; PyTorch fp32 "
                                           "restatement of upstream whisper, all threads",
                                 "logmel_chunks_per_s_1thread": lm1, "logmel_chunks_per_s_all_cores": lmn},
                "e2e": {"value": rtf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import numpy as np
    import torch

    wbm = importlib.import_module("openai-whisper-coreml_b200")
    sharding = importlib.import_module("openai-whisper-coreml_b200.sharding")

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst; kernel timed alone)" if peaks else "fallback 6650 (B200_PROFILING.md)"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        """k calls of fn between barrier + synchronize on both sides; device time (CUDA events on the handle's stream), max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def make_handle(model, max_batch, max_beams=1):
        w = wbm.Whisper(model, seed=(0 if rank == 0 else None), max_batch=max_batch, max_beams=max_beams, device=local_rank,
                        stream=stream.cuda_stream)
        nbytes = 0
        if world > 1:
            nbytes = sharding.broadcast_weights(w, dev, src=0)   # NCCL over NVLink, load time only
        return w, nbytes

    def roofline_of(w, B, reps=120):
        k_ms, k_bytes = w.profile_cross_attention(B, reps)
        achieved = k_bytes / (k_ms * 1e-3) / 1e9
        return {"kernel": "attn_decode_head_kernel (decoder cross-attention over the persistent KV cache, one CTA per sequence x head)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "us_per_launch": k_ms * 1e3}

    def logmel_leg(wh, audio_ptr, B, reps=30):
        """The log-mel front end alone: device f32 PCM in, [B][80][3000] f32 out, CUDA events on the handle's stream.
        Algorithmic bytes per chunk: 480000 * 4 in + 80 * 3000 * 4 out = 2.88 MB."""
        out = torch.empty((B, 80, 3000), dtype=torch.float32, device=dev)
        lib = wbm.load_library()
        for _ in range(3):
            assert lib.wb_logmel_dev(wh.handle, audio_ptr, B, out.data_ptr()) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            lib.wb_logmel_dev(wh.handle, audio_ptr, B, out.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = B * (480000 * 4 + 80 * 3000 * 4)
        return {"chunks": B, "us": ms * 1e3, "GB/s": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak}

    with torch.cuda.stream(stream):
        # ------------------------------------------------------------------------------------------------ headline: base.en B=32
        w, bcast_bytes = make_handle(MODEL, BATCH)
        dims = wbm.DIMS[MODEL]
        opts = wbm.DecodeOptions.default_for(dims, sample_len=SAMPLE_LEN)
        opts.suppress = list(opts.suppress) + [opts.eot]
        host = synth_chunks(1000 + rank * BATCH, BATCH)
        audio_dev = host.to(dev, non_blocking=True)
        stream.synchronize()
        host_np = host.numpy()
        host_pageable = np.array(host_np)                        # an ordinary malloc'ed buffer, as a Swift / C caller would pass

        def step_dev():
            return w.transcribe_dev(audio_dev.data_ptr(), BATCH, opts)

        def step_host():
            return w.transcribe(host_np, opts)

        def step_pageable():
            return w.transcribe(host_pageable, opts)

        # multi-GPU correctness, outside the timed region: identical arenas, and every rank decodes rank 0's first chunk to
        # the tokens rank 0 got (a broken broadcast or a rank-dependent kernel path would still scale linearly)
        ranks_verified = 1
        if world > 1:
            sharding.verify_ranks(w, dev, rank, world)
            probe = synth_chunks(1000, 1)
            t_probe, _, _ = w.transcribe(probe.numpy(), opts)
            t0 = torch.from_numpy(t_probe.astype(np.int32)).to(dev)
            ref0 = t0.clone()
            dist.broadcast(ref0, src=0)
            ok = torch.tensor([1 if torch.equal(t0, ref0) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.SUM)
            ranks_verified = int(ok.item())
            assert ranks_verified == world, f"only {ranks_verified} of {world} ranks reproduce rank 0's tokens"

        for _ in range(warmup):
            step_dev()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = w.launch_count()
        ms = timed(step_dev, args.steps)
        launches = w.launch_count() - l0
        phase = w.last_timings().tolist()
        for _ in range(2):
            step_host()
        ms_e2e = timed(step_host, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        step_pageable()
        ms_pageable = timed(step_pageable, max(2, args.steps // 2)) / max(2, args.steps // 2)
        tokens, lens, slp = step_dev()
        # dominant kernel: KV-cache attention over the resident cross K/V, timed alone with CUDA events on the same stream
        roof = roofline_of(w, BATCH)
        logmel_legs = [logmel_leg(w, audio_dev.data_ptr(), BATCH)]
        if world == 1 and args.configs != "headline":
            w64, _ = make_handle("tiny.en", 64)                   # the widest batch a handle takes (SURVEY asks for 256)
            a64 = synth_chunks(7000, 64).to(dev, non_blocking=True)
            stream.synchronize()
            logmel_legs.append(logmel_leg(w64, a64.data_ptr(), 64))
            w64.close()
            del a64

        extras = {}
        want = [] if args.configs == "headline" else (list(CONFIGS) if args.configs == "all" else args.configs.split(","))
        K2 = max(1, args.extra_steps)

        def run_extra(name, fn_dev, fn_host, audio_s_total, wh, B_roof, note=None):
            for _ in range(3):
                fn_dev()
            m = timed(fn_dev, K2) / K2
            ph = wh.last_timings().tolist()
            fn_host()
            me = timed(fn_host, K2) / K2
            r = {"what": name, "value": audio_s_total / (m * 1e-3), "unit": UNIT, "ms_per_step": m, "steps": K2, "warmup": 3,
                 "e2e": {"value": audio_s_total / (me * 1e-3), "unit": UNIT, "ms_per_step": me},
                 "phase_ms": {"logmel": ph[0], "encoder_and_cross_kv": ph[1], "decode": ph[2], "decode_steps": ph[3]},
                 "roofline": roofline_of(wh, B_roof, 60)}
            if note:
                r["note"] = note
            return r

        # headline variants: sample_len = 100 (speech-typical) and strong scaling (32 chunks in total, 32 / G per GPU)
        if "base_b32" in want:
            o100 = wbm.DecodeOptions.default_for(dims, sample_len=100)
            o100.suppress = list(o100.suppress) + [o100.eot]
            extras["base_b32_sample_len_100"] = run_extra(
                "headline with max_new_tokens = 100", lambda: w.transcribe_dev(audio_dev.data_ptr(), BATCH, o100),
                lambda: w.transcribe(host_np, o100), world * BATCH * 30.0, w, BATCH)
            if world > 1 and BATCH % world == 0:
                bs = BATCH // world
                extras["base_b32_strong_scaling"] = run_extra(
                    f"32 chunks in total, {bs} per GPU (strong scaling)", lambda: w.transcribe_dev(audio_dev.data_ptr(), bs, opts),
                    lambda: w.transcribe(host_np[:bs], opts), BATCH * 30.0, w, bs)
                extras["base_b32_strong_scaling"]["scaling"] = "strong"
        w.close()

        for name in want:
            if name == "base_b32":
                continue
            c = CONFIGS[name]
            cd = wbm.DIMS[c["model"]]
            if "windows" not in c:
                if world > 1:
                    continue                                      # the 1-GPU configurations are measured at N = 1 only
                beams = max(1, c["beam"])
                wh, _ = make_handle(c["model"], c["batch"], beams)
                oc = wbm.DecodeOptions.default_for(cd, sample_len=c["sample_len"])
                oc.suppress = list(oc.suppress) + [oc.eot]
                oc.beam_size = c["beam"]
                hc = synth_chunks(3000, c["batch"])
                dc = hc.to(dev, non_blocking=True)
                stream.synchronize()
                hn = hc.numpy()
                B = c["batch"]
                extras[name] = run_extra(c["what"], lambda: wh.transcribe_dev(dc.data_ptr(), B, oc), lambda: wh.transcribe(hn, oc),
                                         B * 30.0, wh, B,
                                         note=("EOT suppressed: no beam ever finishes, all 224 steps run with 40 live sequences; "
                                               "cross K/V shared per chunk (8 slabs, not 40)") if c["beam"] > 1 else None)
                wh.close()
            else:
                # config 5: one 30-minute stream = 60 windows (seed 5), partitioned over the ranks, token rows all-gathered
                n_win = c["windows"]
                parts = sharding.partition(n_win, world)
                mb = min(c["batch"], max(e - s for s, e in parts))
                wh, _ = make_handle(c["model"], mb)
                oc = wbm.DecodeOptions.default_for(cd, sample_len=c["sample_len"])
                oc.suppress = list(oc.suppress) + [oc.eot]
                pcm = (np.random.default_rng(5).standard_normal(n_win * 480000) * 0.1).astype(np.float32)
                windows = wbm.split_windows(pcm)
                result = {}

                def job():
                    result["tokens"], result["lens"] = sharding.transcribe_windows_sharded(wh, windows, oc, rank, world, device=dev)

                for _ in range(2):
                    job()
                m = timed(job, K2) / K2
                s0, e0 = parts[rank]
                r = {"what": c["what"], "value": n_win * 30.0 / (m * 1e-3), "unit": UNIT, "ms_per_step": m, "steps": K2, "warmup": 2,
                     "scaling": "strong", "windows_per_rank": [e - s for s, e in parts],
                     "e2e": {"value": n_win * 30.0 / (m * 1e-3), "unit": UNIT, "ms_per_step": m,
                             "h2d_bytes_per_step": int((e0 - s0) * 480000 * 4), "note": "the job takes host windows and returns gathered host tokens: value is e2e"},
                     "roofline": roofline_of(wh, (e0 - s0) - ((e0 - s0 - 1) // mb) * mb, 30)}   # the features of the last pass are resident
                if world > 1:
                    # token identity with a single-GPU run: rank 0 transcribes all 60 windows alone and compares
                    if rank == 0:
                        solo_t, solo_l = sharding.transcribe_windows_sharded(wh, windows, oc, 0, 1)
                        same = bool(np.array_equal(solo_t, result["tokens"]) and np.array_equal(solo_l, result["lens"]))
                        assert same, "gathered tokens of the sharded job differ from the single-GPU run"
                        r["tokens_identical_to_single_gpu_run"] = same
                    barrier()
                extras[name] = r
                wh.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("attn_decode_head_cross_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roof["traffic"] = traffic
    ms_step = ms / args.steps
    ms_step_e2e = ms_e2e / args.steps
    audio_s = world * BATCH * 30.0
    line = {"metric": METRIC, "value": audio_s / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": audio_s / (ms_step_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_np.nbytes),
                    "d2h_bytes_per_step": int(tokens.nbytes + lens.nbytes + slp.nbytes), "ms_per_step": ms_step_e2e},
            "e2e_pageable": {"value": audio_s / (ms_pageable * 1e-3), "unit": UNIT, "ms_per_step": ms_pageable,
                             "note": "same call with an ordinary (non-pinned) host buffer"},
            "gpu_launches": int(launches),
            "phase_ms": {"logmel": phase[0], "encoder_and_cross_kv": phase[1], "decode": phase[2], "decode_steps": phase[3]},
            "roofline": roof,
            "logmel": {"kernel": "logmel_kernel + logmel_normalize_kernel (f32, device in / out)", "bound": "instruction issue, not HBM (DESIGN.md section 4)",
                       "algorithmic_bytes_per_chunk": 2880000, "runs": logmel_legs},
            "weights_broadcast_bytes": bcast_bytes, "ranks_verified": ranks_verified, "extra_configs": extras}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        chunks = args.cpu_chunks or 4
        rtf, sec, _, _ = cpu_reference_rtf(chunks, 1, 0, threads)
        lm1, lmn = cpu_logmel_legs(threads)
        line["cpu_baseline"] = {"value": rtf, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{chunks} chunks of the same workload, one pass ({sec:.1f} s): C f64 log-mel restatement of "
                                          "the Rust stft (1 thread) + PyTorch fp32 restatement of upstream whisper (all threads); "
                                          "`--impl reference` runs the full 32-chunk batch",
                                "logmel_chunks_per_s_1thread": lm1, "logmel_chunks_per_s_all_cores": lmn}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
