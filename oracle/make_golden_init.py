"""Generates tests/golden/reference_init_stats.npz: summary statistics of the reference's OWN random
initialisation (TEST INFRASTRUCTURE; run in the build container only, it needs /root/reference).

    python oracle/make_golden_init.py

Builds the unmodified reference `Decoder` (tts/core/codec/decoder.py:17-67 -> `Generator.__init__`,
`init_weights`, `reset_parameters`, decoder_modules.py:403-433, 13-16, 463-464) under a few seeds, for
the xcodec2 config and the 48 kHz upsampler config, and stores per state-dict tensor
(mean, std, min, max) averaged over the seeds plus the shape. `tests/test_host_logic.py` checks
`tts_max_b200.codec.decoder.random_init_state_dict` -- which restates the distributions, not the RNG
stream -- against them.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEEDS = (0, 1, 2)
CONFIGS = {"xcodec2": (16000, 320, None, None), "48k": (48000, 160, [3, 2], [7, 6])}


def main() -> None:
    from tts.core.codec import decoder as ref_decoder  # the reference, unmodified

    out = {}
    for name, (sr, hop, ups, ks) in CONFIGS.items():
        acc: dict[str, list[np.ndarray]] = {}
        for seed in SEEDS:
            torch.manual_seed(seed)
            sd = ref_decoder.Decoder(sr, hop, ups, ks).state_dict()
            for k, v in sd.items():
                v = v.detach().double()
                acc.setdefault(k, []).append(np.array([v.mean(), v.std() if v.numel() > 1 else 0.0, v.min(), v.max()]))
                out[f"{name}/shape/{k}"] = np.array(v.shape, dtype=np.int64)
        out[f"{name}/keys"] = np.array(list(acc.keys()))
        for k, rows in acc.items():
            out[f"{name}/stats/{k}"] = np.mean(np.stack(rows), axis=0)
    path = os.path.join(GOLDEN, "reference_init_stats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", sum(1 for k in out if "/stats/" in k), "tensors")


if __name__ == "__main__":
    main()
