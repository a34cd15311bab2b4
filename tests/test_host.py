"""CPU: host-side logic of the Swift-API mirror and the chunk sharding (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pad_or_trim_follows_contentview(wbm):
    """ContentView.swift:57-60: input = zeros(480000); copy min(n, 480000) samples."""
    short = wbm.pad_or_trim(np.ones(160000, dtype=np.float32))
    assert short.shape == (480000,) and short.dtype == np.float64
    assert short[:160000].sum() == 160000 and not short[160000:].any()
    long = wbm.pad_or_trim(np.arange(500000))
    assert long.shape == (480000,) and long[-1] == 479999
    assert not wbm.pad_or_trim([]).any()


def test_languages_table(wbm):
    assert len(wbm.LANGUAGES) == 99                      # Whisper.swift:12, indexed by token - 50259 (:37)
    assert wbm.LANGUAGES[0] == "en" and wbm.LANGUAGES[1] == "zh" and wbm.LANGUAGES[98] == "su"
    assert len(set(wbm.LANGUAGES)) == 99


def test_dims_match_oracle(wbm, ref):
    for name, d in ref.DIMS.items():
        p = wbm.DIMS[name]
        assert all(getattr(p, f) == getattr(d, f) for f in d.__dataclass_fields__)


@pytest.mark.parametrize("name", ["tiny.en", "small"])
def test_default_decode_options_match_oracle(wbm, ref, name):
    o = wbm.DecodeOptions.default_for(wbm.DIMS[name])
    r = ref.DecodeOptions.default_for(ref.DIMS[name])
    assert list(o.initial_tokens) == list(r.initial_tokens)
    assert list(o.suppress) == list(r.suppress) and list(o.suppress_begin) == list(r.suppress_begin)
    assert o.eot == ref.Vocab.for_dims(ref.DIMS[name]).eot and o.sample_len == r.sample_len == 224
    # upstream's own default (without_timestamps=False): no <|notimestamps|> in the prompt, timestamp rules on
    ot = wbm.DecodeOptions.default_for(wbm.DIMS[name], without_timestamps=False)
    rt = ref.DecodeOptions.default_for(ref.DIMS[name], without_timestamps=False)
    v = ref.Vocab.for_dims(ref.DIMS[name])
    assert list(ot.initial_tokens) == list(rt.initial_tokens) == list(r.initial_tokens)[:-1]
    assert ot.timestamps and rt.timestamps and not o.timestamps and not r.timestamps
    assert (ot.timestamp_begin, ot.no_timestamps, ot.max_initial_timestamp_index) == (v.timestamp_begin, v.no_timestamps, rt.max_initial_timestamp_index)


def test_generate_spectrogram_validates_length(wbm):
    with pytest.raises(ValueError):
        wbm.generateSpectrogram(np.zeros(1000))


def test_hf_name_map_covers_every_upstream_tensor(wbm, ref, small_dims):
    """transformers checkpoint keys -> upstream names (SURVEY §8c map), checked on a live HF model."""
    hf = ref.to_hf(small_dims, ref.random_weights(small_dims, seed=3))
    mapped = {wbm.hf_to_upstream_name(k) for k in hf.state_dict().keys()} - {None}
    assert mapped == set(ref.weight_shapes(small_dims).keys())


def test_split_windows(wbm):
    pcm = np.arange(480000 * 2 + 1000, dtype=np.float32)
    w = wbm.split_windows(pcm)
    assert w.shape == (3, 480000)
    assert w[1, 0] == 480000 and w[2, 999] == 480000 * 2 + 999 and not w[2, 1000:].any()
    assert wbm.split_windows([]).shape == (1, 480000)


def test_partition(wbm):
    from importlib import import_module
    sh = import_module("openai-whisper-coreml_b200.sharding")
    parts = sh.partition(60, 8)                           # BASELINE config 5: 60 windows on 8 GPUs
    assert [e - s for s, e in parts] == [8, 8, 8, 8, 7, 7, 7, 7]
    assert parts[0][0] == 0 and parts[-1][1] == 60 and all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert sh.partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert sh.partition(0, 2) == [(0, 0), (0, 0)]
    with pytest.raises(ValueError):
        sh.partition(4, 0)


def _gloo_worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from importlib import import_module
    sh = import_module("openai-whisper-coreml_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s, e = sh.partition(n_total, world)[rank]
    # "decode" chunk c -> token row [c, c+1, ...]; length c % 5 + 1
    toks = torch.stack([torch.arange(c, c + 6, dtype=torch.int32) for c in range(s, e)]) if e > s else torch.zeros((0, 6), dtype=torch.int32)
    lens = torch.tensor([c % 5 + 1 for c in range(s, e)], dtype=torch.int32)
    all_t, all_l = sh.gather_tokens(toks, lens, n_total, world, rank)
    q.put((rank, all_t.tolist(), all_l.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gather_tokens_world_size_2_gloo(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + n_total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_t = [list(range(c, c + 6)) for c in range(n_total)]
    want_l = [c % 5 + 1 for c in range(n_total)]
    for _, t, l in got:
        assert t == want_t and l == want_l
