"""CPU: pins oracle/logmel_ref.c (restatement of /root/reference/stft/src/lib.rs) and the CUDA kernel's index math."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KINDS = ["noise", "sine", "chirp", "noise_then_zeros", "int16", "fullscale"]


def test_mel_fixture_is_the_reference_fixture(golden_dir):
    raw = open(os.path.join(golden_dir, "m80.npy"), "rb").read()
    assert hashlib.sha256(raw).hexdigest() == "3cd88ccebda3c0589c05574824909c8fd04fd456c6a5a992fd92652ae6de9580"
    m = np.load(os.path.join(golden_dir, "m80.npy")).reshape(80, 201)
    assert m.dtype == np.float32 and int((m != 0).sum()) == 391
    assert not m[:, 0].any() and not m[:, 200].any()          # DC and Nyquist carry no weight
    for row in m:                                             # every row is one contiguous band (the kernel's sparse form)
        nz = np.nonzero(row)[0]
        assert nz.size and nz[-1] - nz[0] + 1 == nz.size


def test_zeros_give_minus_one_point_five(oracle_logmel):
    assert np.all(oracle_logmel(np.zeros(480000)) == -1.5)    # (max(-10, -10-8)+4)/4, lib.rs:76,96


def test_fft_path_matches_naive_dft(oracle_logmel, ref):
    a = ref.synth_audio(2, "noise")
    assert np.abs(oracle_logmel(a) - oracle_logmel(a, naive=True)).max() <= 1e-12


def _numpy_logmel(a, golden_dir):
    """Independent vectorised formulation: reflect pad, frames, numpy rfft, power, mel, log, clamp."""
    mel = np.load(os.path.join(golden_dir, "m80.npy")).reshape(80, 201).astype(np.float64)
    p = np.concatenate([a[200:0:-1], a, a[-2:-202:-1]])
    idx = np.arange(3000)[:, None] * 160 + np.arange(400)[None, :]
    w = (1.0 - np.cos(2.0 * np.pi * np.arange(400) / 400.0)) / 2.0
    power = np.abs(np.fft.rfft(p[idx] * w, axis=1)) ** 2
    lg = np.log10(np.maximum(mel @ power.T, 1e-10))
    return (np.maximum(lg, lg.max() - 8.0) + 4.0) / 4.0


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_vs_numpy_rfft(oracle_logmel, ref, golden_dir, kind):
    a = ref.synth_audio(11, kind)
    assert np.abs(oracle_logmel(a) - _numpy_logmel(a, golden_dir)).max() <= 1e-12


@pytest.mark.parametrize("kind", KINDS + ["zeros"])
def test_oracle_vs_golden_torch_stft(oracle_logmel, ref, golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "logmel_torch_f64.npz"))
    got = oracle_logmel(ref.synth_audio(11, kind)).reshape(-1)[g["index"]]
    assert np.abs(got - g[kind]).max() <= 1e-12


def test_reflect_and_inplace_mutation(oracle_lib):
    buf = np.zeros(480400)
    buf[200:480200] = np.arange(480000, dtype=np.float64)     # ramp: value == unpadded index
    out = np.zeros(240000)
    oracle_lib.logmel_ref_generate_spectrogram(buf.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(buf[:200], np.arange(200, 0, -1))                  # a[i] = a[400-i]      (lib.rs:36)
    assert np.array_equal(buf[480200:], 479998 - np.arange(200))             # tail mirror          (lib.rs:37-38)


def test_sine_lands_in_expected_mel_row(oracle_logmel, golden_dir):
    t = np.arange(480000) / 16000.0
    out = oracle_logmel(0.5 * np.sin(2 * np.pi * 1000.0 * t))
    mel = np.load(os.path.join(golden_dir, "m80.npy")).reshape(80, 201)
    assert int(out[:, 1500].argmax()) == int(mel[:, 25].argmax())            # 1000 Hz = bin 25 of a 400-point FFT at 16 kHz


def test_batch_f32_entry_matches(oracle_lib, oracle_logmel, ref):
    a = np.stack([ref.synth_audio(5, "noise"), ref.synth_audio(6, "sine")]).astype(np.float32)
    out = np.zeros((2, 80, 3000))
    assert oracle_lib.logmel_ref_batch_f32(a.ctypes.data_as(ctypes.c_void_p), 2, out.ctypes.data_as(ctypes.c_void_p)) == 0
    for b in range(2):
        assert np.array_equal(out[b], oracle_logmel(a[b].astype(np.float64)))


# ---- the CUDA kernel's phase functions, executed on the CPU (tests/emu) --------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, "tests", "emu", "libemu_logmel.so")
    if not os.path.exists(so):
        import __graft_entry__ as ge
        ge.build_emulator()
    return ctypes.CDLL(so)


@pytest.mark.parametrize("kind", ["noise", "chirp", "noise_then_zeros", "zeros"])
def test_kernel_emulation_matches_oracle(emu, oracle_logmel, ref, kind):
    a = ref.synth_audio(1, kind)
    a32 = a.astype(np.float32)
    o32 = np.zeros(240000, np.float32)
    emu.emu_logmel_f32(a32.ctypes.data_as(ctypes.c_void_p), o32.ctypes.data_as(ctypes.c_void_p))
    assert np.abs(o32 - oracle_logmel(a32.astype(np.float64)).reshape(-1)).max() <= 2e-4     # fp32 tolerance, SURVEY §8c
    o64 = np.zeros(240000)
    emu.emu_logmel_f64(a.ctypes.data_as(ctypes.c_void_p), o64.ctypes.data_as(ctypes.c_void_p))
    assert np.abs(o64 - oracle_logmel(a).reshape(-1)).max() <= 1e-12


def test_nan_samples_are_floored_not_propagated(oracle_logmel, ref):
    """Rust's f64::max (lib.rs:76) drops a NaN operand: frames that hold a NaN sample come out at the floor, nothing panics."""
    a = ref.synth_audio(5, "noise")
    clean = oracle_logmel(a.copy())
    a[100000] = np.nan
    out = oracle_logmel(a)
    assert np.isfinite(out).all()
    frames = [f for f in range(3000) if f * 160 - 200 <= 100000 < f * 160 + 200]
    assert len(frames) == 3
    floor = clean.max() - 2.0                 # ((M - 8) + 4) / 4 with the normalised maximum (M + 4) / 4
    assert np.allclose(out[:, frames], floor, atol=1e-12)
    keep = np.ones(3000, dtype=bool)
    keep[frames] = False
    assert np.array_equal(out[:, keep], clean[:, keep])
