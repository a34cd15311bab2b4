"""GPU parity tests proper: every call goes through the C ABI (ctypes) into the hand-written CUDA kernels and is
compared with the oracle on the same seeded inputs, with the committed golden vectors, and — at BASELINE.json's full
size — through size-independent properties (determinism, batch-slot independence).

Tolerances (stated once, SURVEY.md §8c / BASELINE.md §3):
  log-mel fp32 kernel   max|d| <= 2e-4 on the normalised output        f64 legacy ABI  <= 1e-12
  encoder features      max|d| <= 5e-2, rel-L2 <= 5e-3 (fp16 operands, fp32 accumulate / LayerNorm / softmax)
  decoder logits        max|d| <= 5e-2, rel-L2 <= 5e-3 (teacher-forced)
  greedy token IDs      identical to the oracle's
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_util as pu  # noqa: E402

pytestmark = pytest.mark.gpu

KINDS = ["noise", "sine", "chirp", "noise_then_zeros", "int16", "fullscale", "zeros"]
TOL_MEL32, TOL_MEL64, TOL_ABS, TOL_REL = 2e-4, 1e-12, 5e-2, 5e-3


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def tiny(wbm, ref):
    dims = ref.DIMS["tiny.en"]
    weights = ref.random_weights(dims, seed=0)
    w = wbm.Whisper("tiny.en", weights=weights, max_batch=4)
    yield w, ref.WhisperRef(dims, weights)
    w.close()


# ---- log-mel -----------------------------------------------------------------------------------------
def test_logmel_f32_all_clip_kinds(tiny, ref, oracle_logmel):
    w, _ = tiny
    for i in range(0, len(KINDS), 4):
        kinds = KINDS[i:i + 4]
        a = np.stack([ref.synth_audio(11, k) for k in kinds]).astype(np.float32)
        got = w.logmel(a)
        for j, k in enumerate(kinds):
            assert np.abs(got[j] - oracle_logmel(a[j].astype(np.float64))).max() <= TOL_MEL32, k


def test_logmel_f32_vs_golden_torch_stft(tiny, ref, golden_dir):
    w, _ = tiny
    g = np.load(os.path.join(golden_dir, "logmel_torch_f64.npz"))
    for k in ["noise", "chirp", "int16"]:
        got = w.logmel(ref.synth_audio(11, k).astype(np.float32))[0].reshape(-1)[g["index"]]
        assert np.abs(got - g[k]).max() <= TOL_MEL32 + 1e-5     # + fp32 rounding of the input clip itself


@pytest.mark.parametrize("kind", ["noise", "chirp", "zeros"])
def test_legacy_f64_symbol(wbm, ref, oracle_logmel, kind):
    a = ref.synth_audio(3, kind)
    got = wbm.generateSpectrogram(a).reshape(80, 3000)
    assert np.abs(got - oracle_logmel(a)).max() <= TOL_MEL64


def test_legacy_symbol_buffer_protocol_and_pad_mutation(wbm):
    """stft.swift:10-15 + lib.rs:34-40: pads are overwritten in place with the reflection."""
    lib = wbm.load_library()
    buf = np.zeros(480400)
    buf[200:480200] = np.arange(480000, dtype=np.float64) * 1e-6
    out = np.zeros(240000)
    lib.generate_spectrogram(buf.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert np.allclose(buf[:200], np.arange(200, 0, -1) * 1e-6, rtol=0, atol=0)
    assert np.allclose(buf[480200:], (479998 - np.arange(200)) * 1e-6, rtol=0, atol=0)
    assert np.isfinite(out).all() and out.max() <= 2.0


def test_logmel_nan_samples_take_the_floor_like_the_reference(tiny, wbm, ref, oracle_logmel):
    """lib.rs:76 `x.max(1e-10)` is Rust's f64::max: a NaN operand is ignored, so every frame whose window holds a NaN sample
    becomes log10(1e-10) = -10 in all 80 bands before the maximum is taken (lib.rs:82-88 never sees a NaN and does not panic).
    The kernels' `sum > 1e-10 ? sum : 1e-10` keeps that, in f64 (legacy symbol) and in f32."""
    a = ref.synth_audio(5, "noise")
    a[100000] = np.nan
    a[300007] = np.nan
    want = oracle_logmel(a)
    assert np.isfinite(want).all()
    hit = [f for f in range(3000) if any(f * 160 - 200 <= s < f * 160 + 200 for s in (100000, 300007))]
    assert len(hit) == 6 and all(np.ptp(want[:, f]) == 0 for f in hit)          # three frames per NaN sample, all bands at the floor
    got64 = wbm.generateSpectrogram(a).reshape(80, 3000)
    assert np.isfinite(got64).all() and np.abs(got64 - want).max() <= TOL_MEL64
    w, _ = tiny
    got32 = w.logmel(a.astype(np.float32))[0]
    assert np.isfinite(got32).all() and np.abs(got32 - want).max() <= TOL_MEL32


def test_logmel_batch_slot_independence(tiny, ref):
    w, _ = tiny
    a = np.stack([ref.synth_audio(30 + i, "noise") for i in range(4)]).astype(np.float32)
    full = w.logmel(a)
    for i in range(4):
        assert np.array_equal(w.logmel(a[i])[0], full[i])


# ---- operators (rows a11/a13: GEMM, LayerNorm, attention) against fp32 torch --------------------------------------------------
@pytest.mark.parametrize("M,N,K,gelu,bias,res,c32", [(128, 128, 64, 0, 0, 0, 1), (300, 256, 512, 0, 1, 0, 1), (1500, 384, 384, 1, 1, 0, 0),
                                                      (3000, 512, 2048, 0, 1, 1, 1), (257, 1536, 512, 0, 1, 0, 0), (200, 128, 240, 1, 1, 0, 0),
                                                      (48000, 512, 512, 0, 1, 1, 1)])
def test_gemm_tcgen05(tiny, wbm, M, N, K, gelu, bias, res, c32):
    w, _ = tiny
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    r = torch.randn(M, N, generator=g).cuda() if res else None
    C = torch.full((M, N), float("nan"), dtype=torch.float32 if c32 else torch.float16, device="cuda")
    lib = wbm.load_library()
    assert lib.wb_op_gemm(w.handle, ptr(A), ptr(W), ptr(b) if bias else None, ptr(r) if res else None, M, N, K, gelu, ptr(C), c32) == 0
    w.sync()
    want = A.float() @ W.float().t()
    want = want + b if bias else want
    want = torch.nn.functional.gelu(want) if gelu else want
    want = want + r if res else want
    tol = (2e-3 if c32 else 4e-3) * max(1.0, want.abs().max().item())
    assert (C.float() - want).abs().max().item() <= tol


def test_gemm_linearity_at_full_size(tiny, wbm):
    """Size-independent property at BASELINE size (M = 32*1500): C(A1 + A2) == C(A1) + C(A2) up to fp32 accumulation."""
    w, _ = tiny
    M, N, K = 48000, 512, 512
    g = torch.Generator().manual_seed(1)
    A1 = (torch.randint(-8, 9, (M, K), generator=g).float() / 8).half().cuda()      # exactly representable sums
    A2 = (torch.randint(-8, 9, (M, K), generator=g).float() / 8).half().cuda()
    W = (torch.randint(-8, 9, (N, K), generator=g).float() / 16).half().cuda()
    lib = wbm.load_library()
    outs = []
    for A in (A1, A2, (A1 + A2)):
        C = torch.empty((M, N), dtype=torch.float32, device="cuda")
        assert lib.wb_op_gemm(w.handle, ptr(A), ptr(W), None, None, M, N, K, 0, ptr(C), 1) == 0
        w.sync()
        outs.append(C)
    assert torch.equal(outs[0] + outs[1], outs[2])       # all partial sums are exact in fp32 for these inputs


@pytest.mark.parametrize("d", [384, 512, 768, 1024, 1280])
def test_layernorm(tiny, wbm, d):
    w, _ = tiny
    x = torch.randn(777, d, device="cuda") * 2 + 0.3
    g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    o = torch.empty(777, d, dtype=torch.float16, device="cuda")
    assert wbm.load_library().wb_op_layernorm(w.handle, ptr(x), ptr(g), ptr(b), 777, d, ptr(o)) == 0
    w.sync()
    want = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
    assert (o.float() - want).abs().max().item() <= 1e-2


@pytest.mark.parametrize("B,T,H", [(1, 64, 2), (2, 200, 6), (2, 1500, 6), (1, 1500, 20), (1, 129, 1)])
def test_encoder_attention(tiny, wbm, B, T, H):
    w, _ = tiny
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, device="cuda") * 1.5).half()
    o = torch.full((B * T, d), float("nan"), dtype=torch.float16, device="cuda")
    assert wbm.load_library().wb_op_attention(w.handle, ptr(qkv), B, T, H, ptr(o)) == 0
    w.sync()
    q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.view(B, T, 3, d).unbind(2)]
    want = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(B * T, d)
    assert (o.float() - want).abs().max().item() <= 1e-2


# ---- encoder / decoder vs the oracle (rows a10-a14) -----------------------------------------------------------------------------
def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_encode_matches_oracle(tiny, ref, oracle_logmel):
    w, oracle = tiny
    audio = np.stack([ref.synth_audio(100 + i, k) for i, k in enumerate(["noise", "int16"])])
    mel = torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float()
    want = oracle.encode(mel)
    got = torch.from_numpy(w.encode(audio.astype(np.float32)))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL
    got2 = torch.from_numpy(w.encode_mel(mel.numpy()))            # encoder.prediction(x_1:) alone
    assert (got2 - want).abs().max().item() <= TOL_ABS and _rel(got2, want) <= TOL_REL


@pytest.mark.parametrize("t", [1, 5, 37])
def test_decoder_logits_match_oracle(tiny, ref, oracle_logmel, t):
    w, oracle = tiny
    audio = np.stack([ref.synth_audio(100 + i, "noise") for i in range(2)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    w.encode(audio.astype(np.float32), return_features=False)
    toks = torch.randint(0, 50000, (2, t), generator=torch.Generator().manual_seed(t))
    want = oracle.decoder_logits(toks, xa_ref)
    got = torch.from_numpy(w.decoder_logits(toks.numpy()))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL


def test_decoder_on_external_features_f32_token(tiny, ref, wbm):
    """decoder.prediction(x_1: [[50258.0]], xa:) with features supplied by the caller (Whisper.swift:34-36)."""
    w, oracle = tiny
    xa = torch.randn(1, 1500, 384, generator=torch.Generator().manual_seed(9))
    w.set_audio_features(xa.numpy())
    lib = wbm.load_library()
    tok = np.array([[50257.0]], dtype=np.float32)
    out = np.empty((1, 1, 51864), dtype=np.float32)
    assert lib.wb_decoder_logits_f32tok(w.handle, tok.ctypes.data_as(ctypes.c_void_p), 1, 1, out.ctypes.data_as(ctypes.c_void_p)) == 0
    want = oracle.decoder_logits(torch.tensor([[50257]]), xa.half().float())
    assert (torch.from_numpy(out) - want).abs().max().item() <= TOL_ABS


def test_greedy_tokens_identical_to_oracle(tiny, ref, wbm, oracle_logmel):
    w, oracle = tiny
    audio = np.stack([ref.synth_audio(200 + i, "noise") for i in range(3)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    opts_ref = ref.DecodeOptions.default_for(oracle.dims, sample_len=24)
    tok_ref, slp_ref, _ = oracle.greedy(xa_ref, opts_ref)
    tok, lens, slp = w.transcribe(audio.astype(np.float32), wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=24))
    n = tok_ref.shape[1]
    mism = (torch.from_numpy(tok[:, :n].astype(np.int64)) != tok_ref).any(0).nonzero()
    assert mism.numel() == 0, f"first divergence at position {int(mism[0])}"
    assert np.abs(slp - slp_ref.numpy()).max() <= 0.05


def test_eot_forcing_and_lengths(tiny, ref, wbm):
    """With everything but {eot, 1234} suppressed the decoder must pick one of the two each step; after the first EOT the
    row is EOT-padded, the length includes it, and the result equals the oracle's."""
    w, oracle = tiny
    xa = torch.randn(2, 1500, 384, generator=torch.Generator().manual_seed(4)).half().float()
    w.set_audio_features(xa.numpy())
    eot = oracle.vocab.eot
    sup = [i for i in range(51864) if i not in (eot, 1234)]
    tok_ref, slp_ref, _ = oracle.greedy(xa, ref.DecodeOptions([50257, 50362], sample_len=10, suppress=sup))
    o = wbm.DecodeOptions([50257, 50362], eot, sample_len=10, suppress=sup, eot_check_interval=2)
    tok, lens, slp = w.greedy(2, o)
    for b in range(2):
        r = tok_ref[b].tolist()
        r = r + [eot] * (12 - len(r))
        assert tok[b].tolist() == r
        body = r[2:]
        assert lens[b] == (2 + body.index(eot) + 1 if eot in body else 12)
    assert np.abs(slp - slp_ref.numpy()).max() <= 0.05


@pytest.mark.parametrize("tag", ["small_en", "small_ml"])
def test_against_golden_hf_vectors(wbm, ref, golden_dir, small_dims, small_dims_ml, tag):
    dims = small_dims if tag == "small_en" else small_dims_ml
    g = np.load(os.path.join(golden_dir, f"whisper_{tag}_hf.npz"))
    pd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    w = wbm.Whisper(pd, weights=ref.random_weights(dims, seed=3), max_batch=1)
    xa = w.encode(ref.synth_audio(21, "noise").astype(np.float32))
    assert np.abs(xa[0, ::75] - g["xa_rows"]).max() <= TOL_ABS
    lg = w.decoder_logits(g["tokens"])[0][:, g["logit_cols"]]
    assert np.abs(lg - g["logits"]).max() <= TOL_ABS
    init = [int(t) for t in g["greedy"][0, :2]]
    tok, _, _ = w.greedy(1, wbm.DecodeOptions(init, eot=dims.n_vocab - 1, sample_len=12))
    assert np.array_equal(tok, g["greedy"])
    if dims.is_multilingual:
        assert int(w.detect_language(1)[0]) == int(g["lang"][0])
    w.close()


@pytest.mark.parametrize("d,heads,vocab", [(768, 12, 51865), (1024, 16, 51864), (1280, 20, 51865)])
def test_wider_models_two_layers(wbm, ref, oracle_logmel, d, heads, vocab):
    """small / medium / large widths with 2 layers each: exercises every width-dependent code path."""
    dims = ref.ModelDims(80, 1500, d, heads, 2, vocab, 448, d, heads, 2)
    weights = ref.random_weights(dims, seed=1)
    oracle = ref.WhisperRef(dims, weights)
    w = wbm.Whisper(wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__]), weights=weights, max_batch=2)
    audio = np.stack([ref.synth_audio(400 + i, "noise") for i in range(2)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    xa = torch.from_numpy(w.encode(audio.astype(np.float32)))
    assert (xa - xa_ref).abs().max().item() <= TOL_ABS and _rel(xa, xa_ref) <= TOL_REL
    toks = torch.randint(0, 50000, (2, 4), generator=torch.Generator().manual_seed(d))
    want = oracle.decoder_logits(toks, xa_ref)
    got = torch.from_numpy(w.decoder_logits(toks.numpy()))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=10)
    tok_ref, _, _ = oracle.greedy(xa_ref, opts_ref)
    tok, _, _ = w.greedy(2, wbm.DecodeOptions.default_for(wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__]), sample_len=10))
    n = tok_ref.shape[1]
    mism = (torch.from_numpy(tok[:, :n].astype(np.int64)) != tok_ref).any(0).nonzero()
    assert mism.numel() == 0, f"first divergence at position {int(mism[0])}"
    w.close()


@pytest.mark.parametrize("B", [1, 5, 11])
def test_base_width_block_kernels(wbm, ref, oracle_logmel, B):
    """d = 512 / 8 heads (BASELINE config 3's width) with 2 layers: the decoder runs as cluster kernels there (self block:
    groups of 4 sequences; post block: groups of 8, 16 CTAs per cluster). Batch sizes that leave ragged groups."""
    dims = ref.ModelDims(80, 1500, 512, 8, 2, 51864, 448, 512, 8, 2)
    weights = ref.random_weights(dims, seed=3)
    oracle = ref.WhisperRef(dims, weights)
    wd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    w = wbm.Whisper(wd, weights=weights, max_batch=B)
    audio = np.stack([ref.synth_audio(700 + i, "noise") for i in range(B)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    w.encode(audio.astype(np.float32))
    toks = torch.randint(0, 50000, (B, 9), generator=torch.Generator().manual_seed(B))
    want = oracle.decoder_logits(toks, xa_ref)
    got = torch.from_numpy(w.decoder_logits(toks.numpy()))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=40)
    tok_ref, slp_ref, _ = oracle.greedy(xa_ref, opts_ref)
    tok, _, slp = w.greedy(B, wbm.DecodeOptions.default_for(wd, sample_len=40))
    n = tok_ref.shape[1]
    mism = (torch.from_numpy(tok[:, :n].astype(np.int64)) != tok_ref).any(0).nonzero()
    assert mism.numel() == 0, f"first divergence at position {int(mism[0])}"
    assert np.allclose(slp, slp_ref.numpy(), rtol=2e-3, atol=5e-2)
    w.close()


def test_whisper_decode_language_id(wbm, ref, oracle_logmel, capsys):
    """Whisper.decode(audioFeatures:) (Whisper.swift:33-40) on the multilingual vocabulary."""
    dims = ref.DIMS["tiny"]
    weights = ref.random_weights(dims, seed=0)
    oracle = ref.WhisperRef(dims, weights)
    w = wbm.Whisper("tiny", weights=weights, max_batch=2)
    audio = np.stack([ref.synth_audio(300 + i, "noise") for i in range(2)])
    xa = w.encode(audio.astype(np.float32))
    want = oracle.detect_language(torch.from_numpy(xa)).tolist()
    codes = w.decode(xa)
    assert codes == [wbm.LANGUAGES[i] for i in want]
    assert capsys.readouterr().out.split() == codes
    w.close()


# ---- BASELINE.json full size: base.en, 32 chunks, 224 tokens — size-independent properties -------------------------------------
def test_full_size_determinism_and_slot_independence(wbm):
    w = wbm.Whisper("base.en", seed=0, max_batch=32)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["base.en"], sample_len=224)
    o.suppress = list(o.suppress) + [o.eot]
    audio = np.stack([(np.random.default_rng(1000 + i).standard_normal(480000) * 0.1).astype(np.float32) for i in range(32)])
    t1, l1, s1 = w.transcribe(audio, o)
    t2, l2, s2 = w.transcribe(audio, o)
    assert np.array_equal(t1, t2) and np.array_equal(s1, s2)                 # run-to-run bit identity
    assert (l1 == 226).all() and (t1[:, 2:] != o.eot).all()
    perm = np.roll(np.arange(32), 5)
    t3, _, s3 = w.transcribe(audio[perm], o)                                   # a chunk's result does not depend on its slot
    assert np.array_equal(t3, t1[perm])
    t4, _, _ = w.transcribe(audio[:3], o)                                      # ... nor on the batch size
    assert np.array_equal(t4[:, :40], t1[:3, :40])
    assert w.launch_count() > 0
    w.close()


def test_long_stream_as_windows_and_hf_checkpoint_names(wbm, ref):
    """BASELINE config 5 shape: a stream longer than 30 s = independent windows; weights loaded under HF key names."""
    dims = ref.DIMS["tiny.en"]
    weights = ref.random_weights(dims, seed=0)
    hf_sd = {k: v for k, v in ref.to_hf(dims, weights).state_dict().items()}
    w = wbm.Whisper("tiny.en", seed=None, max_batch=2)
    w.load_hf_state_dict(hf_sd)
    pcm = np.concatenate([ref.synth_audio(500 + i, "noise") for i in range(3)])[: 480000 * 2 + 123456].astype(np.float32)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=8)
    toks, lens = w.transcribe_long(pcm, o)
    assert toks.shape == (3, 10)
    win = wbm.split_windows(pcm)
    w2 = wbm.Whisper("tiny.en", weights=weights, max_batch=3)
    t2, l2, _ = w2.transcribe(win, o)
    assert np.array_equal(toks, t2) and np.array_equal(lens, l2)
    w.close(), w2.close()


@pytest.mark.parametrize("multilingual,beam,B", [(False, 5, 2), (True, 3, 1)])
def test_beam_search_matches_oracle(wbm, ref, small_dims, small_dims_ml, oracle_logmel, multilingual, beam, B):
    """BASELINE config 4 (beam decode) at test size: upstream BeamSearchDecoder semantics, best candidate per chunk."""
    dims = small_dims_ml if multilingual else small_dims
    weights = ref.random_weights(dims, seed=3)
    oracle = ref.WhisperRef(dims, weights)
    pd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    w = wbm.Whisper(pd, weights=weights, max_batch=B, max_beams=beam)
    audio = np.stack([ref.synth_audio(600 + i, "noise") for i in range(B)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    w.encode(audio.astype(np.float32), return_features=False)
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=9)
    want_tokens, want_scores = oracle.beam_search(xa_ref, opts_ref, beam_size=beam)
    o = wbm.DecodeOptions.default_for(pd, sample_len=9)
    o.beam_size = beam
    tok, lens, slp = w.decode_tokens(B, o)
    for b in range(B):
        n = len(want_tokens[b])
        assert tok[b, :n].tolist() == want_tokens[b], (b, tok[b].tolist(), want_tokens[b])
        assert (tok[b, n:] == o.eot).all()
        assert abs(float(slp[b]) - want_scores[b]) <= 0.05
    # greedy still works on the same handle afterwards (cache pointers were ping-ponged)
    o.beam_size = 0
    g_ref, _, _ = oracle.greedy(xa_ref, opts_ref)
    g, _, _ = w.decode_tokens(B, o)
    assert np.array_equal(g[:, :g_ref.shape[1]].astype(np.int64), g_ref.numpy())
    w.close()


def test_beam_search_on_block_kernels(wbm, ref, oracle_logmel):
    """Beam search at base width (d = 512): the decoder step runs as the cluster kernels with chunks x beams sequences
    (groups of 4 / 8 that straddle chunks), the cross K/V shared per chunk and the self K/V cache re-indexed every step."""
    dims = ref.ModelDims(80, 1500, 512, 8, 2, 51864, 448, 512, 8, 2)
    weights = ref.random_weights(dims, seed=5)
    oracle = ref.WhisperRef(dims, weights)
    pd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    B, beam = 3, 5
    w = wbm.Whisper(pd, weights=weights, max_batch=B, max_beams=beam)
    audio = np.stack([ref.synth_audio(800 + i, "noise") for i in range(B)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    w.encode(audio.astype(np.float32), return_features=False)
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=9)
    want_tokens, want_scores = oracle.beam_search(xa_ref, opts_ref, beam_size=beam)
    o = wbm.DecodeOptions.default_for(pd, sample_len=9)
    o.beam_size = beam
    tok, lens, slp = w.decode_tokens(B, o)
    for b in range(B):
        n = len(want_tokens[b])
        assert tok[b, :n].tolist() == want_tokens[b], (b, tok[b].tolist(), want_tokens[b])
        assert abs(float(slp[b]) - want_scores[b]) <= 0.05
    w.close()


@pytest.mark.parametrize("d,seed", [(768, 7), (1024, 32)])
def test_beam_search_at_small_width(wbm, ref, d, seed):
    """BASELINE config 4's shape (whisper_to_cml.py:7 exports `small`: d = 768, 12 heads, multilingual vocabulary) with 2
    layers: 8 chunks x 5 beams = 40 sequences on the decoder path of the wide models (LayerNorm rows kernel, skinny GEMMs, post
    block), the cross K/V shared by the beams of a chunk (one CTA per (chunk, head)), the graph-replayed scored step; and
    medium's width (d = 1024, 16 heads), where the query projection is a kernel of its own. Weights / features: seeds whose
    oracle result is stable under logit noise of 2e-2 (tools/pick_beam_seed.py), since whole token lists are compared."""
    dims = ref.ModelDims(80, 1500, d, d // 64, 2, 51865, 448, d, d // 64, 2)
    B, beam = 8, 5
    weights = ref.random_weights(dims, seed=seed)
    oracle = ref.WhisperRef(dims, weights)
    pd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    w = wbm.Whisper(pd, weights=weights, max_batch=B, max_beams=beam)
    xa = (torch.randn(B, 1500, d, generator=torch.Generator().manual_seed(100 + seed)) * 0.7).half().float()
    w.set_audio_features(xa.numpy())
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=9)
    want_tokens, want_scores = oracle.beam_search(xa, opts_ref, beam_size=beam)
    o = wbm.DecodeOptions.default_for(pd, sample_len=9)
    o.beam_size = beam
    tok, lens, slp = w.decode_tokens(B, o)
    for b in range(B):
        n = len(want_tokens[b])
        assert tok[b, :n].tolist() == want_tokens[b], (b, tok[b].tolist(), want_tokens[b])
        assert abs(float(slp[b]) - want_scores[b]) <= 0.05
    tok2, _, slp2 = w.decode_tokens(B, o)                       # second call replays the captured graphs
    assert np.array_equal(tok, tok2) and np.array_equal(slp, slp2)
    w.close()


@pytest.mark.parametrize("name,B,text_scale", [("tiny.en", 3, 1.0), ("tiny.en", 3, 1.8), ("tiny", 2, 1.8)])
def test_greedy_with_timestamp_rules_matches_oracle(wbm, ref, oracle_logmel, name, B, text_scale):
    """Upstream's default decoding (without_timestamps=False): ApplyTimestampRules among the logit filters — first token a
    timestamp <= 1 s, timestamps in pairs and non-decreasing, timestamp probability mass above every text token forces a
    timestamp. With seeded random weights the mass of the 1501 timestamp logits always wins (text_scale 1.0: a timestamp
    whenever one is allowed); scaling the text rows of the embedding makes the mass rule fall either way from step to step."""
    dims = ref.DIMS[name]
    v = ref.Vocab.for_dims(dims)
    weights = ref.random_weights(dims, seed=11)
    weights["decoder.token_embedding.weight"][:v.eot] *= text_scale
    oracle = ref.WhisperRef(dims, weights)
    w = wbm.Whisper(name, weights=weights, max_batch=B)
    audio = np.stack([ref.synth_audio(900 + i, "noise") for i in range(B)])
    xa_ref = oracle.encode(torch.from_numpy(np.stack([oracle_logmel(a) for a in audio])).float())
    w.encode(audio.astype(np.float32), return_features=False)
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=48, without_timestamps=False)
    tok_ref, slp_ref, _ = oracle.greedy(xa_ref, opts_ref)
    o = wbm.DecodeOptions.default_for(wbm.DIMS[name], sample_len=48, without_timestamps=False)
    tok, lens, slp = w.greedy(B, o)
    n = tok_ref.shape[1]
    got = torch.from_numpy(tok[:, :n].astype(np.int64))
    n_init = len(opts_ref.initial_tokens)
    same = np.ones(B, dtype=bool)
    for b in range(B):
        diff = (got[b] != tok_ref[b]).nonzero()
        if diff.numel() == 0:
            continue
        # a stream may leave the oracle's only at a tie for the fp16 path: the oracle's own gap between its choice and the GPU's,
        # on the filtered logits of that step, below parity_util.TOL_TIE; everything before that point is identical
        p = int(diff[0])
        prefix = tok_ref[b:b + 1, :p]
        lg = oracle.decoder_logits(prefix, xa_ref[b:b + 1])[:, -1].clone()
        if p == n_init:
            lg[:, list(opts_ref.suppress_begin)] = float("-inf")
        lg[:, list(opts_ref.suppress)] = float("-inf")
        lg = ref.apply_timestamp_rules(lg, prefix, n_init, v, opts_ref.max_initial_timestamp_index)
        gap = float(lg[0, int(tok_ref[b, p])] - lg[0, int(got[b, p])])
        print(f"\n[timestamps {name} x{text_scale}] sequence {b} leaves the oracle at position {p}: oracle gap {gap:.4f}")
        assert 0.0 <= gap <= pu.TOL_TIE, f"sequence {b}, position {p}: {got[b].tolist()} vs {tok_ref[b].tolist()} (gap {gap})"
        same[b] = False
        # ... and from there on every choice is graded against the oracle given the GPU's own prefix (teacher-forced)
        row = got[b:b + 1]
        all_lg = oracle.decoder_logits(row[:, :-1], xa_ref[b:b + 1])
        for q in range(p + 1, n):
            lq = all_lg[:, q - 1].clone()
            lq[:, list(opts_ref.suppress)] = float("-inf")
            lq = ref.apply_timestamp_rules(lq, row[:, :q], n_init, v, opts_ref.max_initial_timestamp_index)
            gq = float(lq[0].max() - lq[0, int(row[0, q])])
            assert 0.0 <= gq <= pu.TOL_TIE, f"sequence {b}, position {q} (after the tie at {p}): gap {gq}"
    assert same.any()
    assert np.allclose(slp[same], slp_ref.numpy()[same], rtol=2e-3, atol=5e-2)
    body = tok_ref[:, len(opts_ref.initial_tokens):]
    assert (body[:, 0] >= v.timestamp_begin).all() and (body[:, 0] <= v.timestamp_begin + 50).all()   # the rules did act
    assert (body >= v.timestamp_begin).any(1).all() and (body < v.eot).any()                            # both classes sampled
    # the same handle without the rules afterwards: nothing of the rule state leaks (bit-identical to a fresh handle)
    o2 = wbm.DecodeOptions.default_for(wbm.DIMS[name], sample_len=8)
    g, _, gs = w.greedy(B, o2)
    w2 = wbm.Whisper(name, weights=weights, max_batch=B)
    w2.encode(audio.astype(np.float32), return_features=False)
    g2, _, gs2 = w2.greedy(B, o2)
    assert np.array_equal(g, g2) and np.array_equal(gs, gs2)
    w.close(), w2.close()


def test_beam_search_with_timestamp_rules(wbm, ref):
    """ApplyTimestampRules under beam search (upstream's default options with beam_size set): the per-sequence rule state is
    rebuilt from the host-side sequences every step, the logits kernel applies the per-row rules to all chunks x beams rows,
    the top-k kernel the probability-mass rule before the log-softmax. Seed / text scale from tools/pick_beam_ts_seed.py
    (both token classes sampled, result stable under logit noise of 2e-2)."""
    dims = ref.DIMS["tiny"]
    v = ref.Vocab.for_dims(dims)
    seed, B, beam = 2, 2, 3
    weights = ref.random_weights(dims, seed=seed)
    weights["decoder.token_embedding.weight"][:v.eot] *= 1.8
    oracle = ref.WhisperRef(dims, weights)
    w = wbm.Whisper("tiny", weights=weights, max_batch=B, max_beams=beam)
    xa = (torch.randn(B, 1500, dims.n_audio_state, generator=torch.Generator().manual_seed(200 + seed)) * 0.7).half().float()
    w.set_audio_features(xa.numpy())
    opts_ref = ref.DecodeOptions.default_for(dims, sample_len=12, without_timestamps=False)
    want_tokens, want_scores = oracle.beam_search(xa, opts_ref, beam_size=beam)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny"], sample_len=12, without_timestamps=False)
    o.beam_size = beam
    tok, lens, slp = w.decode_tokens(B, o)
    n_init = len(o.initial_tokens)
    for b in range(B):
        n = len(want_tokens[b])
        assert tok[b, :n].tolist() == want_tokens[b], (b, tok[b].tolist(), want_tokens[b])
        assert abs(float(slp[b]) - want_scores[b]) <= 0.05
        body = want_tokens[b][n_init:]
        assert body[0] >= v.timestamp_begin and any(t < v.eot for t in body)        # the rules acted, text was sampled too
    # the rules off again on the same handle: identical to a fresh handle (no rule state left behind, graphs re-keyed)
    o2 = wbm.DecodeOptions.default_for(wbm.DIMS["tiny"], sample_len=6)
    o2.beam_size = beam
    t1, _, s1 = w.decode_tokens(B, o2)
    w2 = wbm.Whisper("tiny", weights=weights, max_batch=B, max_beams=beam)
    w2.set_audio_features(xa.numpy())
    t2, _, s2 = w2.decode_tokens(B, o2)
    assert np.array_equal(t1, t2) and np.array_equal(s1, s2)
    w.close(), w2.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_handles_on_two_devices_in_one_process(wbm, ref):
    """One process, one handle per GPU (the launch-attribute caches of the kernels are per device): same tokens on both."""
    dims = ref.DIMS["tiny.en"]
    weights = ref.random_weights(dims, seed=0)
    audio = np.stack([ref.synth_audio(40 + i, "noise") for i in range(2)]).astype(np.float32)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=12)
    res = []
    for dev in (0, 1):
        w = wbm.Whisper("tiny.en", weights=weights, max_batch=2, device=dev)
        res.append(w.transcribe(audio, o))
        w.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][2], res[1][2])


def test_thirty_six_sequences_tiny(wbm, ref, oracle_logmel):
    """33..48 sequences: the logits kernel runs with 48 MMA columns (two TMEM loads per row), the block kernels with ragged
    last groups (36 = 9 x 4 = 4 x 8 + 4). Features are set directly so that the oracle only has to run the decoder."""
    dims = ref.DIMS["tiny.en"]
    weights = ref.random_weights(dims, seed=2)
    oracle = ref.WhisperRef(dims, weights)
    B = 36
    w = wbm.Whisper("tiny.en", weights=weights, max_batch=B)
    xa = torch.randn(B, 1500, dims.n_audio_state, generator=torch.Generator().manual_seed(9)) * 0.7
    w.set_audio_features(xa.numpy())
    toks = torch.randint(0, 50000, (B, 3), generator=torch.Generator().manual_seed(4))
    want = oracle.decoder_logits(toks, xa)
    got = torch.from_numpy(w.decoder_logits(toks.numpy()))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL
    tok_ref, slp_ref, _ = oracle.greedy(xa, ref.DecodeOptions.default_for(dims, sample_len=10))
    tok, _, slp = w.greedy(B, wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=10))
    n = tok_ref.shape[1]
    differ = (torch.from_numpy(tok[:, :n].astype(np.int64)) != tok_ref).any(1)
    # a stream may leave the oracle's only at a tie (fp16 operands): every choice is graded against the oracle given the GPU's prefix
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=10)
    rep = pu.teacher_forced_check(oracle, xa, tok, len(o.initial_tokens), o.suppress, o.suppress_begin, o.eot)
    assert rep.bad == 0 and rep.ties >= int(differ.sum()) and int(differ.sum()) <= 2, rep.line()
    same = (~differ).numpy()
    assert np.allclose(slp[same], slp_ref.numpy()[same], rtol=2e-3, atol=5e-2)
    w.close()


def test_sixty_four_sequences_base_width(wbm, ref):
    """The handle's upper limit where a layer runs as the block kernels: 64 sequences = 16 self-block clusters, 8 post-block
    clusters, the logits kernel with 64 MMA columns (two full TMEM loads per row). Wider models stay at 40 (skinny-GEMM path)."""
    dims = ref.ModelDims(80, 1500, 512, 8, 2, 51864, 448, 512, 8, 2)
    weights = ref.random_weights(dims, seed=8)
    oracle = ref.WhisperRef(dims, weights)
    wd = wbm.ModelDims(*[getattr(dims, f) for f in dims.__dataclass_fields__])
    B = 64
    w = wbm.Whisper(wd, weights=weights, max_batch=B)
    xa = (torch.randn(B, 1500, 512, generator=torch.Generator().manual_seed(19)) * 0.7).half().float()
    w.set_audio_features(xa.numpy())
    toks = torch.randint(0, 50000, (B, 3), generator=torch.Generator().manual_seed(5))
    want = oracle.decoder_logits(toks, xa)
    got = torch.from_numpy(w.decoder_logits(toks.numpy()))
    assert (got - want).abs().max().item() <= TOL_ABS and _rel(got, want) <= TOL_REL
    o = wbm.DecodeOptions.default_for(wd, sample_len=12)
    tok, _, slp = w.greedy(B, o)
    rep = pu.teacher_forced_check(oracle, xa, tok, len(o.initial_tokens), o.suppress, o.suppress_begin, o.eot)
    print("\n[d=512 x 2 layers, 64 sequences x 12] " + rep.line())
    assert rep.decisions == B * 12 and rep.bad == 0 and rep.ties <= 3, rep.line()
    w.close()
    with pytest.raises(wbm.WhisperB200Error, match="max_batch"):
        wbm.Whisper("small", max_batch=41)
    with pytest.raises(wbm.WhisperB200Error, match="max_batch"):
        wbm.Whisper("tiny.en", max_batch=65)


def test_error_paths(tiny, wbm):
    w, _ = tiny
    lib = wbm.load_library()
    with pytest.raises(wbm.WhisperB200Error, match="out of range"):
        w.logmel(np.zeros((5, 480000), dtype=np.float32))                       # max_batch = 4
    with pytest.raises(wbm.WhisperB200Error, match="unknown tensor"):
        w.load_state_dict({"encoder.nope": np.zeros(3, dtype=np.float32)})
    fresh = wbm.Whisper("tiny.en", seed=None)
    with pytest.raises(wbm.WhisperB200Error, match="weights not loaded"):
        fresh.encode(np.zeros(480000, dtype=np.float32))
    fresh.close()
    w2 = wbm.Whisper("tiny.en", seed=1)
    with pytest.raises(wbm.WhisperB200Error, match="not resident"):
        w2.decoder_logits(np.array([[1]]))
    w2.close()
