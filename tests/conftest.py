import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def wbm():
    """The product package (host mirror of the Swift API over the C ABI)."""
    return importlib.import_module("openai-whisper-coreml_b200")


@pytest.fixture(scope="session")
def ref():
    """oracle/whisper_ref.py — the checker."""
    import whisper_ref
    return whisper_ref


@pytest.fixture(scope="session")
def oracle_lib():
    so = os.path.join(ROOT, "oracle", "liblogmel_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.logmel_ref_batch_f32.restype = ctypes.c_int
    return lib


@pytest.fixture(scope="session")
def oracle_logmel(oracle_lib):
    def fn(audio64: np.ndarray, naive: bool = False) -> np.ndarray:
        """audio64 [480000] f64 -> [80,3000] f64 through the same buffer protocol as stft.swift:10-15."""
        buf = np.zeros(480400)
        buf[200:480200] = audio64
        out = np.zeros(240000)
        f = oracle_lib.logmel_ref_generate_spectrogram_naive if naive else oracle_lib.logmel_ref_generate_spectrogram
        f(buf.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        return out.reshape(80, 3000)
    return fn


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def small_dims(ref):
    return ref.ModelDims(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2)


@pytest.fixture(scope="session")
def small_dims_ml(ref):
    return ref.ModelDims(80, 1500, 128, 2, 2, 51865, 448, 128, 2, 2)
