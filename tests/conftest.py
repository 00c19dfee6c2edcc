"""Shared fixtures. `-m "not gpu"` runs on the CPU-only build container; `-m gpu` on a B200."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_decode_seed0.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN_PATH))


@pytest.fixture(scope="session")
def state_dict(golden):
    """The deterministic weights the golden vectors were generated with (seed 0, perturbed)."""
    from oracle import weights

    sd = weights.make_state_dict(seed=0, perturb=True)
    fp = weights.fingerprint(sd)
    want = float(golden["weights_fingerprint"])
    assert abs(fp - want) <= 1e-9 * max(1.0, abs(want)), (
        f"torch CPU RNG drift: weights fingerprint {fp!r} != golden {want!r}; regenerate tests/golden "
        "with oracle/make_golden.py in the build container"
    )
    return sd


@pytest.fixture(scope="session")
def gpu_decoders(state_dict):
    """One B200 decoder per operand precision, loaded with the golden weights."""
    from tts_max_b200.codec import decoder

    out = {}
    for prec in ("bf16", "fp16"):
        d = decoder.Decoder(16000, 320, None, None, precision=prec)
        d.load_state_dict(state_dict)
        d.to("cuda").eval()
        out[prec] = d
    return out
