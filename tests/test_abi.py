"""CPU: the C-ABI library loads and exports exactly what include/b200codec.h declares.
No compute calls (there is no GPU here and no CPU fallback in the library)."""

import ctypes
import os
import re

import pytest

from tts_max_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200codec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:codec|enc)_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.commonpath([ROOT, _lib.LIB_PATH]) == ROOT


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in b200codec.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures out of sync with the header"


def test_no_libcuda_or_torch_link_dependency():
    # the boundary is a plain C ABI: no torch types, and it must load without a driver
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    assert "libcuda.so" not in out


def test_config_struct_matches_header():
    assert ctypes.sizeof(_lib.B200CodecConfig) == 16 * 4
    text = open(os.path.join(ROOT, "include", "b200codec.h")).read()
    assert f"#define B200CODEC_ABI_VERSION {_lib.ABI_VERSION}" in text


def test_error_reporting_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.B200CodecConfig(abi_version=_lib.ABI_VERSION, sample_rate=16000, hop_length=320, n_upsample=0,
                               precision=0, device=0, hidden_dim=1024, depth=12, heads=16, vq_dim=2048)
    h = ctypes.c_void_p()
    rc = lib.b200codec_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value  # fails loudly: no CPU fallback
    assert len(lib.b200codec_last_error()) > 0
    cfg.hop_length = 300
    assert lib.b200codec_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    with pytest.raises(ValueError):
        _lib.check(1)  # "Current hop length ..." maps to the reference's ValueError
