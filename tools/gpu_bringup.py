#!/usr/bin/env python3
"""GPU bring-up: runs every kernel against its reference and prints error magnitudes (no early exit).
Development tool (the graded parity tests are tests/test_gpu_*.py). Uses oracle/ as the checker only."""
import ctypes
import importlib
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
wbm = importlib.import_module("openai-whisper-coreml_b200")
import whisper_ref as ref  # noqa: E402

RESULTS = []


def section(name):
    def deco(fn):
        def run():
            t0 = time.time()
            try:
                msg = fn()
                RESULTS.append((name, "PASS" if msg is None or not str(msg).startswith("FAIL") else "FAIL", msg))
            except Exception as e:  # noqa: BLE001
                traceback.print_exc()
                RESULTS.append((name, "ERROR", repr(e)))
            print(f"[{name}] {RESULTS[-1][1]} {RESULTS[-1][2]}  ({time.time() - t0:.1f}s)", flush=True)
        return run
    return deco


def oracle_logmel(audio64):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liblogmel_ref.so"))
    B = audio64.shape[0]
    out = np.zeros((B, 80, 3000))
    for b in range(B):
        buf = np.zeros(480400)
        buf[200:480200] = audio64[b]
        lib.logmel_ref_generate_spectrogram(buf.ctypes.data_as(ctypes.c_void_p), out[b].ctypes.data_as(ctypes.c_void_p))
    return out


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


@section("logmel_f32")
def t_logmel():
    w = wbm.Whisper("tiny.en", seed=1, max_batch=4)
    kinds = ["noise", "sine", "chirp", "zeros"]
    a = np.stack([ref.synth_audio(1 + i, k) for i, k in enumerate(kinds)]).astype(np.float32)
    got = w.logmel(a)
    want = oracle_logmel(a.astype(np.float64))
    err = [float(np.abs(got[i] - want[i]).max()) for i in range(4)]
    w.close()
    return ("FAIL " if max(err) > 2e-4 else "") + f"max|d| per clip {dict(zip(kinds, err))}"


@section("logmel_f64_legacy")
def t_legacy():
    a = ref.synth_audio(3, "noise")
    got = wbm.generateSpectrogram(a).reshape(80, 3000)
    want = oracle_logmel(a[None])[0]
    e = float(np.abs(got - want).max())
    return ("FAIL " if e > 1e-12 else "") + f"max|d| {e:.3e}"


def gemm_case(w, M, N, K, gelu, use_bias, use_res, c32):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    bias = torch.randn(N, generator=g).cuda() if use_bias else None
    res = torch.randn(M, N, generator=g).cuda() if use_res else None
    C = torch.full((M, N), float("nan"), dtype=torch.float32 if c32 else torch.float16, device="cuda")
    lib = wbm.load_library()
    rc = lib.wb_op_gemm(w.handle, ptr(A), ptr(W), ptr(bias) if use_bias else None, ptr(res) if use_res else None, M, N, K,
                        int(gelu), ptr(C), int(c32))
    assert rc == 0, lib.wb_last_error()
    w.sync()
    want = A.float() @ W.float().t()
    if use_bias:
        want = want + bias
    if gelu:
        want = torch.nn.functional.gelu(want)
    if use_res:
        want = want + res
    err = (C.float() - want).abs().max().item()
    scale = want.abs().max().item()
    return err, scale


def run_gemm_cases(w):
    out = []
    bad = False
    for (M, N, K, gelu, b, r, c32) in [(128, 128, 64, 0, 0, 0, 1), (128, 128, 512, 0, 0, 0, 1), (300, 256, 512, 0, 1, 0, 1),
                                       (1500, 384, 384, 1, 1, 0, 0), (3000, 512, 2048, 0, 1, 1, 1), (257, 1536, 512, 0, 1, 0, 0),
                                       (200, 128, 240, 1, 1, 0, 0)]:
        err, scale = gemm_case(w, M, N, K, gelu, b, r, c32)
        tol = 2e-3 * max(scale, 1.0) if c32 else 4e-3 * max(scale, 1.0)
        ok = err == err and err <= tol
        bad |= not ok
        out.append(f"{M}x{N}x{K}{'g' if gelu else ''}{'b' if b else ''}{'r' if r else ''}{'/f32' if c32 else '/f16'}:{err:.2e}{'' if ok else '(!)'}")
    return ("FAIL " if bad else "") + " ".join(out)


@section("gemm_tcgen05")
def t_gemm_tc():
    w = wbm.Whisper("tiny.en", seed=1)
    r = run_gemm_cases(w)
    w.close()
    return r


@section("layernorm")
def t_ln():
    w = wbm.Whisper("tiny.en", seed=1)
    lib = wbm.load_library()
    errs = []
    for d in (384, 512, 768, 1280):
        x = torch.randn(1000, d, device="cuda") * 2 + 0.3
        g = torch.randn(d, device="cuda")
        b = torch.randn(d, device="cuda")
        o = torch.empty(1000, d, dtype=torch.float16, device="cuda")
        assert lib.wb_op_layernorm(w.handle, ptr(x), ptr(g), ptr(b), 1000, d, ptr(o)) == 0
        w.sync()
        want = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
        errs.append((o.float() - want).abs().max().item())
    w.close()
    return ("FAIL " if max(errs) > 2e-2 else "") + f"max|d| {errs}"


@section("encoder_attention")
def t_att():
    w = wbm.Whisper("tiny.en", seed=1)
    lib = wbm.load_library()
    msgs = []
    bad = False
    for (B, T, H) in [(1, 64, 2), (2, 200, 6), (2, 1500, 6), (1, 1500, 8)]:
        d = H * 64
        qkv = (torch.randn(B * T, 3 * d, device="cuda") * 1.5).half()
        o = torch.full((B * T, d), float("nan"), dtype=torch.float16, device="cuda")
        assert lib.wb_op_attention(w.handle, ptr(qkv), B, T, H, ptr(o)) == 0
        w.sync()
        q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.view(B, T, 3, d).unbind(2)]
        want = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
        want = want.transpose(1, 2).reshape(B * T, d)
        e = (o.float() - want).abs().max().item()
        bad |= not (e < 1e-2)
        msgs.append(f"B{B}T{T}H{H}:{e:.2e}")
    w.close()
    return ("FAIL " if bad else "") + " ".join(msgs)


def model_case(name, B, n_tok, sample_len):
    dims_ref = ref.DIMS[name]
    weights = ref.random_weights(dims_ref, seed=0)
    oracle = ref.WhisperRef(dims_ref, weights)
    w = wbm.Whisper(name, weights=weights, max_batch=B)
    audio = np.stack([ref.synth_audio(100 + i, "noise") for i in range(B)])
    mel = torch.from_numpy(oracle_logmel(audio)).float()
    xa_ref = oracle.encode(mel)
    xa = torch.from_numpy(w.encode(audio.astype(np.float32)))
    e_xa = (xa - xa_ref).abs().max().item()
    rel_xa = ((xa - xa_ref).norm() / xa_ref.norm()).item()
    # teacher-forced logits
    g = torch.Generator().manual_seed(5)
    toks = torch.randint(0, 50000, (B, n_tok), generator=g)
    lg_ref = oracle.decoder_logits(toks, xa_ref)
    lg = torch.from_numpy(w.decoder_logits(toks.numpy()))
    e_lg = (lg - lg_ref).abs().max().item()
    rel_lg = ((lg - lg_ref).norm() / lg_ref.norm()).item()
    # greedy
    opts_ref = ref.DecodeOptions.default_for(dims_ref, sample_len=sample_len)
    tok_ref, slp_ref, _ = oracle.greedy(xa_ref, opts_ref)
    o = wbm.DecodeOptions.default_for(wbm.DIMS[name], sample_len=sample_len)
    tok, lens, slp = w.greedy(B, o)
    n = tok_ref.shape[1]
    same = bool((torch.from_numpy(tok[:, :n].astype(np.int64)) == tok_ref).all())
    first_div = -1
    if not same:
        neq = (torch.from_numpy(tok[:, :n].astype(np.int64)) != tok_ref).any(0).nonzero()
        first_div = int(neq[0])
    lang = None
    if dims_ref.is_multilingual:
        lang = (w.detect_language(B).tolist(), oracle.detect_language(xa_ref).tolist())
    w.close()
    bad = e_xa > 5e-2 or rel_lg > 5e-3 or not same
    return (("FAIL " if bad else "") + f"xa max|d| {e_xa:.3e} rel {rel_xa:.2e}; logits max|d| {e_lg:.3e} rel {rel_lg:.2e}; "
            f"greedy identical={same} first_div={first_div} slp d={np.abs(slp - slp_ref.numpy()).max():.3e} lang={lang}")


@section("model_tiny.en_tc")
def t_model_tiny():
    return model_case("tiny.en", 2, 5, 12)


@section("model_tiny_multilingual")
def t_model_tiny_ml():
    return model_case("tiny", 1, 3, 6)


@section("bench_base.en_B32")
def t_bench():
    w = wbm.Whisper("base.en", seed=0, max_batch=32)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["base.en"], sample_len=224)
    o.suppress = list(o.suppress) + [o.eot]
    audio = (np.random.default_rng(0).standard_normal((32, 480000)) * 0.1).astype(np.float32)
    out = []
    for it in range(3):
        t0 = time.time()
        tok, lens, slp = w.transcribe(audio, o)
        dt = time.time() - t0
        out.append(f"{dt * 1e3:.1f}ms timings={w.last_timings().tolist()}")
    w.close()
    return " | ".join(out) + f" RTF={32 * 30 / dt:.0f}x"


if __name__ == "__main__":
    only = sys.argv[1:]
    print(torch.cuda.get_device_name(0), flush=True)
    table = {"logmel": t_logmel, "legacy": t_legacy, "gemm_tc": t_gemm_tc, "ln": t_ln, "att": t_att,
             "tiny_tc": t_model_tiny, "tiny_ml": t_model_tiny_ml, "bench": t_bench}
    for k, fn in table.items():
        if not only or k in only:
            fn()
    print("\n==== summary ====")
    for name, status, msg in RESULTS:
        print(f"{status:6s} {name}: {msg}")
