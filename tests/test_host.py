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


# ---- token table, text rules of the long-form loop (host code of the C ABI: no GPU needed) ----------------------------------
def _synthetic_ranks():
    """256 byte tokens + BPE merges learned from a small multilingual corpus (no vocabulary file exists offline)."""
    import collections
    corpus = ("the quick brown fox jumps over the lazy dog. The theme of the thesis: thé naïve café — 東京 tokyo! 12345 12 123\n\n"
              "  indented   text's it's we'll they've I'm he'd 3.14 <tag> end\t\ttabs ")
    ranks = {bytes([b]): b for b in range(256)}
    seqs = [[bytes([b]) for b in w.encode()] for w in corpus.split(" ")]
    for _ in range(80):
        cnt = collections.Counter()
        for s in seqs:
            for a, b in zip(s, s[1:]):
                cnt[a + b] += 1
        cnt = {k: c for k, c in cnt.items() if k not in ranks}
        if not cnt:
            break
        best = max(cnt, key=lambda k: (cnt[k], k))
        ranks[best] = len(ranks)
        for s in seqs:
            i = 0
            while i < len(s) - 1:
                if s[i] + s[i + 1] == best:
                    s[i:i + 2] = [best]
                else:
                    i += 1
    return corpus, ranks


def test_tokenizer_against_tiktoken(wbm, tmp_path):
    """Encode (Python merge loop) and decode (C ABI) against tiktoken, the library upstream whisper's tokenizer is built on,
    constructed offline from the same ranks; the .tiktoken file reader of the C ABI; vocab.json with GPT-2's byte alphabet."""
    import base64
    import ctypes
    import json
    tiktoken = pytest.importorskip("tiktoken")
    corpus, ranks = _synthetic_ranks()
    pat = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""
    enc = tiktoken.Encoding("synthetic", pat_str=pat, mergeable_ranks=ranks, special_tokens={})
    t = wbm.Tokenizer(ranks)
    assert len(t) == len(ranks)
    for text in [corpus, " hello the theme", "東京the", "", "   ", "it's 12 o'clock", corpus[::-1]]:
        a, b = enc.encode_ordinary(text), t.encode(text)
        assert a == b, text
        assert t.decode(b) == text == enc.decode(a)
    ids = t.encode(corpus)
    assert t.decode(ids, drop_from=256) == enc.decode([i for i in ids if i < 256])        # the timestamp / special-token filter
    assert t.decode(ids[:3] + [10 ** 6, -1] + ids[3:]) == corpus                           # ids outside the table decode to nothing
    f = tmp_path / "synthetic.tiktoken"
    f.write_bytes(b"".join(base64.b64encode(tok) + b" " + str(r).encode() + b"\n" for tok, r in ranks.items()))
    t2 = wbm.Tokenizer.from_tiktoken(str(f))
    assert t2.table == t.table
    lib = wbm.load_library()
    h = ctypes.c_void_p(lib.wb_tokenizer_load_tiktoken(str(f).encode()))
    assert h and lib.wb_tokenizer_size(h) == len(ranks)
    a = np.asarray(ids, dtype=np.int32)
    n = ctypes.c_size_t(0)
    out = np.zeros(4096, dtype=np.uint8)
    assert lib.wb_tokenizer_decode(h, a.ctypes.data_as(ctypes.c_void_p), a.size, 1 << 30, out.ctypes.data_as(ctypes.c_void_p), out.size,
                                   ctypes.byref(n)) == 0
    assert out[:n.value].tobytes() == corpus.encode()
    lib.wb_tokenizer_destroy(h)
    assert not lib.wb_tokenizer_load_tiktoken(str(tmp_path / "missing").encode()) and b"cannot open" in lib.wb_last_error()
    bad = tmp_path / "bad.tiktoken"
    bad.write_bytes(b"not base64 !!\n")
    assert not lib.wb_tokenizer_load_tiktoken(str(bad).encode())
    b2u = wbm.bytes_to_unicode()
    from transformers.convert_slow_tokenizer import bytes_to_unicode as hf_b2u
    assert b2u == hf_b2u()                                                                 # an independent copy of the byte alphabet
    vj = tmp_path / "vocab.json"
    vocab = {"".join(b2u[b] for b in tok): r for tok, r in ranks.items()}
    vocab["<|endoftext|>"] = len(ranks)
    vj.write_text(json.dumps(vocab, ensure_ascii=False), encoding="utf-8")
    t3 = wbm.Tokenizer.from_vocab_json(str(vj))
    assert t3.table[:len(ranks)] == t.table and t3.decode(ids) == corpus


def test_compression_ratio_and_text_rules_against_python(wbm):
    """upstream compression_ratio(tokenizer.decode(tokens).strip()): the C ABI's UTF-8 decoding with replacement, Unicode
    strip and zlib deflate against CPython's, on random byte strings rich in invalid sequences and white space."""
    import random
    import zlib
    t = wbm.Tokenizer({bytes([b]): b for b in range(256)})
    rng = random.Random(0)
    special = b" \t\n\x0b\x0c\r\x1c\x1f\x85\xc2\xa0\xc2\x85\xe1\x9a\x80\xe2\x80\x83\xe2\x80\xa8\xe2\x81\x9f\xe3\x80\x80abc\xf0\x9f\x98\x80\xed\xa0\x80\xc0\xf5\xe0\x80\xf4\x90"
    for i in range(3000):
        raw = bytes(rng.choice([rng.randrange(256), rng.choice(special)]) for _ in range(rng.randrange(0, 48)))
        text = raw.decode("utf-8", errors="replace").strip().encode("utf-8")
        want = len(text) / len(zlib.compress(text))
        assert abs(t.compression_ratio(list(raw)) - want) < 1e-6, raw
    rep = list(b"again and again and " * 40)
    text = bytes(rep).decode().strip().encode()
    assert abs(t.compression_ratio(rep) - len(text) / len(zlib.compress(text))) < 1e-5 and t.compression_ratio(rep) > 2.4
    assert t.compression_ratio([]) == 0.0


def test_oracle_sampling_generator(ref):
    """The counter-based generator both sides draw from: splitmix64 known answers (Vigna's reference sequence from state 0),
    uniform range strictly inside (0, 1) in fp32, Gumbel statistics."""
    x, outs = 0, []
    for _ in range(3):                                           # the reference generator advances its state by the golden gamma
        outs.append(ref.splitmix64(x))
        x = (x + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    assert outs == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    arr = ref.splitmix64(np.array([0, 0x9E3779B97F4A7C15], dtype=np.uint64))
    assert [int(a) for a in arr] == outs[:2]
    g = ref.gumbel_noise(7, 1, 3, 51864)
    assert torch.isfinite(g).all() and abs(float(g.mean()) - 0.5772) < 0.02 and abs(float(g.var()) - np.pi ** 2 / 6) < 0.05
    assert not torch.equal(g, ref.gumbel_noise(7, 2, 3, 51864)) and not torch.equal(g, ref.gumbel_noise(7, 1, 4, 51864))
    assert ref.call_seed(5, 300, 2) != ref.call_seed(5, 300, 1) != ref.call_seed(5, 301, 1)


def test_call_seed_matches_the_c_abi(wbm, ref):
    lib = wbm.load_library()
    for seed, seek, ti in [(0, 0, 0), (3, 1804, 2), (2 ** 63 + 11, 179999, 5)]:
        assert lib.wb_call_seed(seed, seek, ti) == ref.call_seed(seed, seek, ti)


def test_oracle_stream_logmel_first_window_against_chunk_oracle(ref, oracle_logmel, golden_dir):
    """The long-form front end restated with torch.stft agrees with the chunk oracle (lib.rs restatement, f64) wherever the
    two definitions coincide: a 30 s clip whose last 2 s are silent — the stream's zero padding and the chunk's reflection
    then see the same samples, and both normalise with the same maximum."""
    a = ref.synth_audio(3, "noise")
    a[-32000:] = 0.0
    got = ref.log_mel_stream(a.astype(np.float32), np.load(os.path.join(golden_dir, "m80.npy")))[:, :3000].numpy()
    want = oracle_logmel(a.astype(np.float32).astype(np.float64))
    assert np.abs(got - want).max() < 2e-4


# ---- checkpoint files (host part: the safetensors header reader and the shape -> dimensions rule) ---------------------------
def _write_checkpoints(ref, tmp_path, dims, seed=1):
    from safetensors.torch import save_file
    w = ref.random_weights(dims, seed=seed)
    up = tmp_path / "upstream.safetensors"
    save_file({k: v.contiguous() for k, v in w.items()}, str(up))
    hf = ref.to_hf(dims, w)
    sd = {k: v.detach().clone().contiguous().half() for k, v in hf.state_dict().items() if k != "proj_out.weight"}
    hfp = tmp_path / "hf_f16.safetensors"
    save_file(sd, str(hfp), metadata={"format": "pt", "note": "nested \"quotes\" and a brace } in metadata"})
    return w, str(up), str(hfp)


def test_safetensors_dims_from_shapes(wbm, ref, tmp_path):
    """A file written by the `safetensors` package (an independent writer) under upstream names in F32 and under transformers
    names in F16: the ten model dimensions come out of the tensor shapes; malformed files are errors with a message."""
    import ctypes
    dims = ref.ModelDims(80, 1500, 128, 2, 2, 51865, 448, 128, 2, 3)
    _, up, hfp = _write_checkpoints(ref, tmp_path, dims)
    for path in (up, hfp):
        got = wbm.read_checkpoint_dims(path)
        assert tuple(getattr(got, f) for f in got.__dataclass_fields__) == tuple(getattr(dims, f) for f in dims.__dataclass_fields__)
    lib = wbm.load_library()
    raw = open(up, "rb").read()
    cases = {"short": raw[:5], "hdr_len": (10 ** 12).to_bytes(8, "little") + raw[8:200], "not_json": (16).to_bytes(8, "little") + b"[1, 2, 3]       ",
             "truncated": raw[:len(raw) // 2], "empty": (2).to_bytes(8, "little") + b"{}"}
    from importlib import import_module
    D = import_module("openai-whisper-coreml_b200.whisper")._Dims
    for tag, blob in cases.items():
        p = tmp_path / f"{tag}.safetensors"
        p.write_bytes(blob)
        assert lib.wb_safetensors_read_dims(str(p).encode(), ctypes.byref(D())) == -1, tag
        assert b"safetensors" in lib.wb_last_error()
    with pytest.raises(wbm.WhisperB200Error, match="cannot open"):
        wbm.read_checkpoint_dims(str(tmp_path / "missing.safetensors"))
