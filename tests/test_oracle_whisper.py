"""CPU: pins oracle/whisper_ref.py (restatement of upstream whisper's AudioEncoder/TextDecoder/greedy) against the
independent transformers implementation — live on small dims and through the committed golden vectors."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def small(ref, small_dims):
    w = ref.random_weights(small_dims, seed=3)
    return ref.WhisperRef(small_dims, w), w


def _mel(ref, oracle_logmel, seed=21):
    return torch.from_numpy(oracle_logmel(ref.synth_audio(seed, "noise"))).float()[None]


def test_restatement_matches_hf_live(ref, small_dims, small, oracle_logmel):
    model, w = small
    hf = ref.to_hf(small_dims, w)
    mel = _mel(ref, oracle_logmel)
    xa = model.encode(mel)
    with torch.no_grad():
        xa_hf = hf.model.encoder(mel).last_hidden_state
        toks = torch.tensor([[50257, 50362, 7, 8, 9]])
        lg_hf = hf(input_features=mel, decoder_input_ids=toks).logits
    assert (xa - xa_hf).abs().max() <= 1e-4
    assert (model.decoder_logits(toks, xa) - lg_hf).abs().max() <= 1e-4


@pytest.mark.parametrize("tag", ["small_en", "small_ml"])
def test_restatement_matches_golden(ref, small_dims, small_dims_ml, golden_dir, oracle_logmel, tag):
    dims = small_dims if tag == "small_en" else small_dims_ml
    g = np.load(os.path.join(golden_dir, f"whisper_{tag}_hf.npz"))
    model = ref.WhisperRef(dims, ref.random_weights(dims, seed=3))
    mel = _mel(ref, oracle_logmel)
    xa = model.encode(mel)
    assert np.abs(xa[0, ::75].numpy() - g["xa_rows"]).max() <= 1e-4
    toks = torch.from_numpy(g["tokens"])
    lg = model.decoder_logits(toks, xa)[0][:, g["logit_cols"]].numpy()
    assert np.abs(lg - g["logits"]).max() <= 1e-4
    opts = ref.DecodeOptions(list(g["greedy"][0, :2]), sample_len=12)          # no filters, like the fixture
    out, _, _ = model.greedy(xa, opts)
    assert np.array_equal(out.numpy(), g["greedy"])
    if dims.is_multilingual:
        assert int(model.detect_language(xa)[0]) == int(g["lang"][0])


def test_kv_cache_equals_uncached(small, ref, oracle_logmel):
    model, _ = small
    xa = model.encode(_mel(ref, oracle_logmel))
    toks = torch.tensor([[50257, 50362, 11, 12, 13, 14]])
    full = model.decoder_logits(toks, xa)
    cache = [None] * model.dims.n_text_layer
    cross = model.cross_kv(xa)
    step = torch.cat([model.decoder_logits(toks[:, i:i + 1], xa, cross, cache) for i in range(toks.shape[1])], dim=1)
    assert (full - step).abs().max() <= 1e-4


def test_language_argmax_is_first_max(ref):
    """Whisper.swift:38 `max { $0.element < $1.element }`: Swift's Sequence.max(by:) replaces its running result only on a
    strict increase, so the FIRST maximal element wins ties and NaN never replaces it (all-equal logits print "en")."""
    lg = torch.zeros(1, 51865)
    assert int(ref.language_argmax(lg)[0]) == 0
    lg[0, 50259 + 4] = 3.0
    lg[0, 50259 + 17] = 3.0
    assert int(ref.language_argmax(lg)[0]) == 4
    lg[0, 50259 + 98] = 5.0
    assert int(ref.language_argmax(lg)[0]) == 98
    lg[0, 50259 + 50] = float("nan")
    assert int(ref.language_argmax(lg)[0]) == 98
    lg2 = torch.full((1, 51865), float("nan"))
    assert int(ref.language_argmax(lg2)[0]) == 0


def test_eot_forcing_and_logprob_mask(ref, small_dims):
    """After a sequence samples EOT every later token is EOT and its log-probs stop accumulating."""
    dims = small_dims
    model = ref.WhisperRef(dims, ref.random_weights(dims, seed=3))
    xa = torch.zeros(2, 1500, dims.n_audio_state)
    eot = model.vocab.eot
    keep = {eot, 1234}
    opts = ref.DecodeOptions([50257, 50362], sample_len=6, suppress=[i for i in range(dims.n_vocab) if i not in keep])
    toks, slp, _ = model.greedy(xa, opts)
    for row in toks.tolist():
        body = row[2:]
        if eot in body:
            k = body.index(eot)
            assert all(t == eot for t in body[k:])
    assert torch.isfinite(slp).all()


def test_beam_size_one_reduces_to_greedy(ref, small_dims):
    """Sanity pin of the beam restatement: with one beam and no EOT in reach it must follow the greedy path and
    accumulate the same log-probabilities."""
    model = ref.WhisperRef(small_dims, ref.random_weights(small_dims, seed=3))
    xa = torch.randn(2, 1500, small_dims.n_audio_state, generator=torch.Generator().manual_seed(2))
    opts = ref.DecodeOptions.default_for(small_dims, sample_len=6)
    opts.suppress = list(opts.suppress) + [model.vocab.eot]
    g, slp, _ = model.greedy(xa, opts)
    bt, bs = model.beam_search(xa, opts, beam_size=1)
    for b in range(2):
        assert bt[b] == g[b].tolist()
        assert abs(bs[b] - float(slp[b])) <= 1e-3


def test_beam_finds_at_least_greedy_likelihood(ref, small_dims):
    model = ref.WhisperRef(small_dims, ref.random_weights(small_dims, seed=3))
    xa = torch.randn(1, 1500, small_dims.n_audio_state, generator=torch.Generator().manual_seed(5))
    opts = ref.DecodeOptions.default_for(small_dims, sample_len=6)
    opts.suppress = list(opts.suppress) + [model.vocab.eot]
    _, slp, _ = model.greedy(xa, opts)
    _, bs = model.beam_search(xa, opts, beam_size=4)
    assert bs[0] >= float(slp[0]) - 1e-4


@pytest.mark.parametrize("multilingual", [False, True])
def test_timestamp_rules_match_hf_logits_processor(ref, multilingual):
    """The oracle's ApplyTimestampRules restatement against transformers' WhisperTimeStampLogitsProcessor (an independent
    implementation of the same upstream rules) on random logits and token histories that exercise every branch: first
    position, text after a timestamp pair, single timestamp, timestamp mass above / below the best text token."""
    from types import SimpleNamespace
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor

    dims = ref.DIMS["tiny" if multilingual else "tiny.en"]
    v = ref.Vocab.for_dims(dims)
    V, ts = dims.n_vocab, v.timestamp_begin
    g = torch.Generator().manual_seed(7 + int(multilingual))
    begin = 3 if multilingual else 1
    cfg = SimpleNamespace(no_timestamps_token_id=v.no_timestamps, eos_token_id=v.eot, bos_token_id=v.eot, max_initial_timestamp_index=50)
    hf = WhisperTimeStampLogitsProcessor(cfg, begin_index=begin)
    prompt = list(range(v.sot, v.sot + begin))
    histories = [[], [ts + 3], [ts + 3, 100], [ts + 3, 100, 200, ts + 40], [ts + 3, 100, ts + 40, ts + 40],
                 [ts, 11, 12, ts + 7, ts + 7, 13], [500, 600], [ts + 1499]]
    for hist in histories:
        for boost in (0.0, 12.0):                           # boost: put the probability mass on the timestamps
            tokens = torch.tensor([prompt + hist, prompt + hist], dtype=torch.long)
            logits = torch.randn(2, V, generator=g) * 3.0
            logits[1, ts:] += boost
            want = hf(tokens, logits)
            got = ref.apply_timestamp_rules(logits, tokens, begin, v, 50)
            assert torch.equal(torch.isinf(got), torch.isinf(want)), (hist, boost)
            assert torch.equal(got[~torch.isinf(got)], want[~torch.isinf(want)])
            assert torch.equal(got.argmax(-1), want.argmax(-1))
