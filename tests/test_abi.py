"""CPU: the C-ABI library loads and exports every symbol include/whisper_b200.h declares; argument / no-device errors
are reported as status codes (no compute is attempted without a GPU)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int|void)\s+(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_symbol():
    syms = _declared_symbols()
    assert "generate_spectrogram" in syms            # Whisper/Whisper/bridge.h:11
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(wbm):
    lib = ctypes.CDLL(wbm.library_path())
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header(wbm):
    from importlib import import_module
    bound = set(import_module("openai-whisper-coreml_b200.whisper").exported_symbols())
    assert set(_declared_symbols()) <= bound


def test_integration_notes_name_every_entry_point():
    """INTEGRATION.md is the maintainer's map from the reference's call sites to the C ABI: no declared symbol is left out of it."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in _declared_symbols() if s not in text]
    assert not missing, missing


def test_swift_mirror_only_uses_declared_names():
    """The Swift side of the boundary (not compiled here: no toolchain) must at least bind names the header declares."""
    header = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    declared = set(re.findall(r"\bwb_[a-z0-9_]+\b", header)) | {"generate_spectrogram"}
    sdir = os.path.join(ROOT, "openai-whisper-coreml_b200", "swift")
    used = set()
    for f in os.listdir(sdir):
        if f.endswith(".swift"):
            used |= set(re.findall(r"\bwb_[a-z0-9_]+\b", open(os.path.join(sdir, f)).read()))
    assert used and used <= declared, sorted(used - declared)
    assert "generate_spectrogram" in open(os.path.join(sdir, "stft.swift")).read()


def test_library_has_no_libcuda_or_torch_dependency(wbm):
    out = subprocess.run(["ldd", wbm.library_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "torch" not in out and "cublas" not in out and "cufft" not in out


def test_version_and_bad_arguments(wbm):
    lib = wbm.load_library()
    assert lib.wb_version() >= 100
    assert lib.wb_create(None, 1, 1, 0, None, None) == -1                      # WB_ERR_ARG
    assert b"bad argument" in lib.wb_last_error()
    assert lib.wb_destroy(None) == -1
    assert lib.wb_launch_count(None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_device_is_an_error_not_a_fallback(wbm):
    with pytest.raises(wbm.WhisperB200Error, match="no CUDA device"):
        wbm.Whisper("tiny.en")
    with pytest.raises(wbm.WhisperB200Error, match="no CUDA device"):
        wbm.generateSpectrogram(np.zeros(480000))
    lib = wbm.load_library()
    buf, out = np.zeros(480400), np.zeros(240000)
    assert lib.wb_generate_spectrogram_f64(buf.ctypes.data_as(ctypes.c_void_p), 1, out.ctypes.data_as(ctypes.c_void_p)) == -2


def test_missing_library_fails_loudly(wbm, monkeypatch, tmp_path):
    from importlib import import_module
    mod = import_module("openai-whisper-coreml_b200.whisper")
    monkeypatch.setattr(mod, "_LIB", None)
    monkeypatch.setenv("WHISPER_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(wbm.WhisperB200Error, match="no CPU fallback"):
        mod.load_library()


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (contract ③)."""
    pkg = os.path.join(ROOT, "openai-whisper-coreml_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "whisper_ref" not in txt and "logmel_ref" not in txt and "oracle/" not in txt.replace("oracle/ ", ""), f


def test_struct_layouts_match_the_header(wbm, tmp_path):
    """The ctypes mirrors of wb_dims / wb_decode_opts / wb_long_opts / wb_segment have exactly the layout a C compiler gives the header's structs
    (what a Swift / cgo / JNI binding generated from include/whisper_b200.h would see): sizes and every field offset."""
    from importlib import import_module
    mod = import_module("openai-whisper-coreml_b200.whisper")
    mirrors = (("wb_dims", mod._Dims), ("wb_decode_opts", mod._DecodeOpts), ("wb_long_opts", mod._LongOpts), ("wb_segment", mod._Segment))
    fields = {st: [n for n, _ in cls._fields_] for st, cls in mirrors}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "whisper_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for n in names:
            lines.append(f'  printf("{st}.{n} %zu\\n", offsetof({st}, {n}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in mirrors:
        assert int(got[st]) == ctypes.sizeof(cls), st
        for n in fields[st]:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, (st, n)
