"""GPU parity of the long-form path: the stream log-mel, temperature sampling with best_of and the
no-speech probability, and upstream's transcribe() loop (wb_transcribe_long) against oracle/whisper_ref.py transcribe_seek.

Every call goes through the C ABI. The scenario (tests/longform_util.py) is shaped so that the loop takes its branches:
fallback over the temperature schedule, the silence override of the fallback, the no-speech skip, several segments per
window with seek advanced to the last timestamp, whole-window advances, cleared segments, prompt reset after a hot window."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import longform_util as lu  # noqa: E402

pytestmark = pytest.mark.gpu


def _mel_filters(golden_dir):
    return np.load(os.path.join(golden_dir, "m80.npy"))


def test_stream_logmel_matches_upstream_restatement(wbm, ref, golden_dir):
    """log_mel_spectrogram(audio, padding=N_SAMPLES): reflection only at the start of the recording, zeros after its end, one
    maximum for the whole recording; windows at arbitrary frame offsets, also across the 3000-frame item boundary and into
    the zero padding."""
    w = wbm.Whisper("tiny.en", seed=1, max_batch=1)
    audio = lu.recording(70.37, seed=8)
    want = ref.log_mel_stream(audio, _mel_filters(golden_dir)).numpy()
    content = want.shape[1] - 3000
    assert content == audio.shape[0] // 160
    for frame0 in (0, 1, 1234, 2999, 3000, 4037, content - 1500, content):
        got = w.logmel_long(audio, frame0)
        ref_seg = want[:, frame0:frame0 + 3000]
        assert got.shape == (80, 3000)
        err = np.abs(got[:, :ref_seg.shape[1]] - ref_seg).max()
        assert err <= 2e-4, f"frame0 {frame0}: max|d| {err}"
        assert (got[:, ref_seg.shape[1]:] == 0).all()
    # a short clip (< 30 s) and the chunk path afterwards (the stream maximum lives in the same scratch word)
    short = lu.recording(3.21, seed=9)
    got = w.logmel_long(short, 0)
    assert np.abs(got - ref.log_mel_stream(short, _mel_filters(golden_dir)).numpy()[:, :3000]).max() <= 2e-4
    chunk = np.zeros(480000, dtype=np.float32)
    assert np.allclose(w.logmel(chunk), -1.5)
    w.close()


@pytest.mark.parametrize("name,best_of,temperature,n_prompt", [("tiny.en", 1, 0.7, 0), ("tiny.en", 3, 1.0, 5), ("tiny", 2, 0.4, 0)])
def test_temperature_sampling_and_no_speech_prob_match_oracle(wbm, ref, name, best_of, temperature, n_prompt):
    """wb_decode with temperature > 0: Gumbel-max draws over the counter-based generator, best_of samples sharing the cross
    K/V, MaximumLikelihoodRanker, timestamp rules (mass rule inside the sampler); no_speech_prob from the unfiltered logits
    at the sot position — with a prompt in front of the sot sequence, and when sot is the last initial token."""
    dims, v, weights = lu.scenario(ref, name)
    oracle = ref.WhisperRef(dims, weights)
    w = wbm.Whisper(name, weights=weights, max_batch=1, max_beams=max(best_of, 1))
    xa = (torch.randn(1, 1500, dims.n_audio_state, generator=torch.Generator().manual_seed(40 + best_of)) * 0.7).half().float()
    w.set_audio_features(xa.numpy())
    prompt = [1000 + 7 * i for i in range(n_prompt)]
    for seed in (1, 2, 3):
        o_ref = ref.DecodeOptions.default_for(dims, sample_len=24, without_timestamps=False)
        res = ref.decode_window(oracle, xa, o_ref, prompt, temperature, seed, best_of, 0, None)
        o = wbm.DecodeOptions.default_for(wbm.DIMS[name], sample_len=24, without_timestamps=False)
        if n_prompt:
            o.initial_tokens = [o.sot_prev] + prompt + list(o.initial_tokens)
            o.sot_index = n_prompt + 1
        o.temperature, o.best_of, o.seed = temperature, best_of, seed
        tok, lens, slp, nsp = w.decode_with_no_speech(1, o)
        n0 = len(o.initial_tokens)
        got = tok[0, n0:lens[0]].tolist()
        got = got[:got.index(o.eot)] if o.eot in got else got
        if got != res.tokens:   # allowed only where the oracle's own top-2 gap of the perturbed logits is a tie for the fp16 path
            assert res.min_margin <= TOL_TIE, f"seed {seed}: {got} vs {res.tokens} (oracle's smallest gap {res.min_margin:.4f})"
            print(f"\n[sampling] seed {seed}: tie (oracle gap {res.min_margin:.4f})")
            continue
        assert abs(slp[0] / (len(got) + 1) - res.avg_logprob) <= 2e-2
        assert abs(nsp[0] - res.no_speech_prob) <= 2e-3 + 2e-2 * res.no_speech_prob
    # temperature 0 through the same entry point: the arg-max path, same no-speech probability
    o.temperature, o.best_of = 0.0, 0
    tok0, lens0, slp0, nsp0 = w.decode_with_no_speech(1, o)
    res0 = ref.decode_window(oracle, xa, o_ref, prompt, 0.0, 0, 0, 0, None)
    got0 = tok0[0, n0:lens0[0]].tolist()
    got0 = got0[:got0.index(o.eot)] if o.eot in got0 else got0
    if got0 != res0.tokens:   # as above: only where the oracle's own top-2 gap is a tie for the fp16 path
        assert res0.min_margin <= TOL_TIE, f"{got0} vs {res0.tokens} (oracle's smallest gap {res0.min_margin:.4f})"
        print(f"\n[arg-max] tie (oracle gap {res0.min_margin:.4f})")
    assert abs(nsp0[0] - res0.no_speech_prob) <= 2e-3 + 2e-2 * res0.no_speech_prob
    w.close()


TOL_TIE = 2e-2   # as tests/parity_util.py: a top-1 / top-2 gap of the oracle's (perturbed) logits below this is a tie for the fp16 path


def _compare(res, toks_ref, segs_ref, trace):
    """Segments and tokens identical to the oracle's. Every decision of the loop is an arg-max over (perturbed) logits the CUDA
    path computes with fp16 operands, so a draw whose top-2 gap in the oracle is below TOL_TIE may fall the other way and the
    recordings' paths part from that window on: then everything before that window must be identical, and the oracle's own
    smallest gap inside that window (over every decode of it, rejected ones included) must indeed be a tie."""
    n_same = 0
    for g, r in zip(res["segments"], segs_ref):
        if g["seek"] != r["seek"] or g["tokens"] != r["tokens"] or g["temperature"] != pytest.approx(r["temperature"]):
            break
        n_same += 1
    if n_same < len(segs_ref) or len(res["segments"]) != len(segs_ref):
        cands = [x["seek"] for x in (res["segments"][n_same:n_same + 1] + segs_ref[n_same:n_same + 1])]
        w = min(cands)
        margins = {s: r.min_margin for s, _, r in trace}
        assert w in margins, f"paths part at a window the oracle never visited (seek {w})"
        assert margins[w] <= TOL_TIE, f"window at seek {w} differs although the oracle's smallest gap there is {margins[w]:.4f}"
        assert all(g["seek"] < w for g in res["segments"][:n_same])
        print(f"\n[long-form] tie at seek {w} (oracle gap {margins[w]:.4f}): {n_same} of {len(segs_ref)} segments compared")
        segs_ref = segs_ref[:n_same]
        res = {"segments": res["segments"][:n_same], "tokens": None}
    else:
        assert res["tokens"].tolist() == toks_ref
    for g, r in zip(res["segments"], segs_ref):
        assert g["seek"] == r["seek"] and g["tokens"] == r["tokens"], (g, r)
        assert abs(g["start"] - r["start"]) < 1e-3 and abs(g["end"] - r["end"]) < 1e-3
        assert g["temperature"] == pytest.approx(r["temperature"])
        assert abs(g["avg_logprob"] - r["avg_logprob"]) <= 2e-2
        assert abs(g["compression_ratio"] - r["compression_ratio"]) <= 1e-5
        assert abs(g["no_speech_prob"] - r["no_speech_prob"]) <= 2e-3 + 2e-2 * r["no_speech_prob"]


def test_transcribe_seek_matches_oracle(wbm, ref, golden_dir):
    """110 s of synthetic audio through upstream's loop on tiny.en dims: tokens, segments (seek, times, temperatures,
    thresholds' inputs) identical to the oracle's restatement; the branches the scenario is built for did occur."""
    dims, v, weights = lu.scenario(ref, "tiny.en")
    oracle = ref.WhisperRef(dims, weights)
    audio = lu.recording(110.0)
    table = lu.synthetic_table(v.eot)
    temps = (0.0, 0.6, 1.0)
    kw = dict(temperatures=temps, logprob_threshold=-4.5, compression_ratio_threshold=2.0, no_speech_threshold=0.85)
    o_ref = ref.DecodeOptions.default_for(dims, sample_len=40, without_timestamps=False)
    toks_ref, segs_ref, trace = ref.transcribe_seek(oracle, audio, _mel_filters(golden_dir), o_ref, table=table, seed=3, best_of=2, **kw)
    w = wbm.Whisper("tiny.en", weights=weights, max_batch=1, max_beams=2)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny.en"], sample_len=40, without_timestamps=False)
    o.best_of, o.seed = 2, 3
    res = w.transcribe_seek(audio, o, tokenizer=_table_tokenizer(wbm, table), **kw)
    print("\n[long-form tiny.en 110 s] windows (seek, temperatures tried, tokens): " +
          str([(s, t, len(r.tokens)) for s, t, r in trace]))
    _compare(res, toks_ref, segs_ref, trace)
    tried = [t for _, t, _ in trace]
    assert any(len(t) > 1 for t in tried) and any(len(t) == 1 for t in tried)              # fallback taken and not taken
    assert len({s["seek"] for s in segs_ref}) < len(segs_ref)                              # several segments in one window
    assert any(b - a not in (0, 3000) for (a, _, _), (b, _, _) in zip(trace, trace[1:]))   # seek advanced to a timestamp
    assert any(r.no_speech_prob > 0.85 and len(t) == 1 and r.compression_ratio > 2.0 for _, t, r in trace)   # silence overrides the fallback
    # arg-max only, tighter log-probability threshold: windows that look like silence and are not confidently text are skipped
    kw1 = dict(temperatures=(0.0,), logprob_threshold=-3.7, compression_ratio_threshold=2.0, no_speech_threshold=0.85)
    toks1, segs1, trace1 = ref.transcribe_seek(oracle, audio, _mel_filters(golden_dir), o_ref, table=table, seed=3, **kw1)
    res1 = w.transcribe_seek(audio, o, tokenizer=_table_tokenizer(wbm, table), **kw1)
    _compare(res1, toks1, segs1, trace1)
    assert len(segs1) and len(trace1) > len({s["seek"] for s in segs1})                    # a window was skipped as silence
    # without a tokenizer (no compression-ratio rule), without conditioning on the previous text, with an initial prompt
    kw2 = dict(temperatures=(0.0, 1.0), logprob_threshold=-3.0, compression_ratio_threshold=None, no_speech_threshold=None,
               condition_on_previous_text=False)
    prompt = [2000, 2001, 2002]
    toks2, segs2, trace2 = ref.transcribe_seek(oracle, audio[: 16000 * 50], _mel_filters(golden_dir), o_ref, table=None, seed=10, best_of=2,
                                               initial_prompt=prompt, **kw2)
    o.seed = 10
    res2 = w.transcribe_seek(audio[: 16000 * 50], o, initial_prompt=prompt, **kw2)
    _compare(res2, toks2, segs2, trace2)
    assert all(len(t) == 2 for _, t, _ in trace2)                                          # every window fell back to temperature 1
    w.close()


def _table_tokenizer(wbm, table):
    """Tokenizer over a table with duplicate byte strings: built through the C entry point directly (ranks are ids)."""
    import ctypes
    t = wbm.Tokenizer.__new__(wbm.Tokenizer)
    t._lib = wbm.load_library()
    blob = np.frombuffer(b"".join(table), dtype=np.uint8)
    offs = np.zeros(len(table) + 1, dtype=np.uint32)
    np.cumsum([len(x) for x in table], out=offs[1:])
    t._h = ctypes.c_void_p(t._lib.wb_tokenizer_create(blob.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), len(table)))
    t.table, t.ranks = list(table), {}
    assert t._h
    return t


def test_transcribe_seek_multilingual_with_language_detection(wbm, ref, golden_dir):
    dims, v, weights = lu.scenario(ref, "tiny", seed=33)
    oracle = ref.WhisperRef(dims, weights)
    audio = lu.recording(47.0, seed=12)
    mel = ref.log_mel_stream(audio, _mel_filters(golden_dir))
    lang = int(oracle.detect_language(oracle.encode(mel[None, :, :3000]))[0])
    o_ref = ref.DecodeOptions.default_for(dims, sample_len=32, language=lang, without_timestamps=False)
    kw = dict(temperatures=(0.0, 0.8), logprob_threshold=-4.5, compression_ratio_threshold=None, no_speech_threshold=0.85)
    toks_ref, segs_ref, trace = ref.transcribe_seek(oracle, audio, _mel_filters(golden_dir), o_ref, seed=4, **kw)
    w = wbm.Whisper("tiny", weights=weights, max_batch=1)
    o = wbm.DecodeOptions.default_for(wbm.DIMS["tiny"], sample_len=32, language=0, without_timestamps=False)
    o.seed = 4
    res = w.transcribe_seek(audio, o, detect_language=True, **kw)
    assert res["language"] == lang
    _compare(res, toks_ref, segs_ref, trace)
    w.close()
