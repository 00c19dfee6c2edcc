"""Golden vectors for the 48 kHz upsampler variant, from the UNMODIFIED reference.

    python oracle/make_golden_48k.py        (build container only: needs /root/reference)

Config = example/configs/codec_training_config.json:24-37 (sample_rate 48000, hop_length 160,
upsample_factors [3, 2], kernel_sizes [7, 6]). The weights (oracle.weights, seed 0, perturbed) are saved
in the tts-max checkpoint layout -- the only layout that can carry `upsampler.*` (decoder.py:112-119) --
and loaded through the reference's own decoding.create().
"""

from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

from oracle import weights  # noqa: E402

CFG = {"model_type": "", "sample_rate": 48000, "token_rate": 50, "hop_length": 160,
       "upsample_factors": [3, 2], "kernel_sizes": [7, 6]}


def main() -> None:
    from tts.core.codec import decoding as ref_decoding

    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict(seed=0, perturb=True, hop=160, upsample_factors=[3, 2], kernel_sizes=[7, 6])
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "model_config.json"), "w") as f:
            json.dump(CFG, f)
        path = os.path.join(tmp, "ckpt.pt")
        torch.save(weights.to_ttsmax_checkpoint(sd), path)
        ref = ref_decoding.create(path, device="cpu")
    ref_sd = ref._decoder.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state-dict key order differs from the reference"
    assert all(torch.equal(ref_sd[k], sd[k]) for k in sd)

    g = torch.Generator().manual_seed(4321)
    out = {"weights_fingerprint": np.float64(weights.fingerprint(sd))}
    for name, T in (("u29", 29), ("u3", 3), ("u1", 1)):
        ids = torch.randint(0, 65536, (T,), generator=g)
        wav = ref.decode(ids)
        assert wav.shape == (1, 960 * T)
        out[f"{name}_ids"] = ids.numpy().astype(np.int64)
        out[f"{name}_wav"] = wav.numpy().astype(np.float32)
    ids_b = torch.randint(0, 65536, (2, 16), generator=g)
    cap = {}
    dec = ref._decoder
    hooks = [
        dec.decoder.backbone.register_forward_hook(lambda m, i, o: cap.__setitem__("backbone", o.detach())),
        dec.upsampler.upsample_layers[0].register_forward_hook(lambda m, i, o: cap.__setitem__("up0", o.detach())),
        dec.upsampler.resnet_blocks[0].register_forward_hook(lambda m, i, o: cap.__setitem__("res0", o.detach())),
        dec.upsampler.upsample_layers[1].register_forward_hook(lambda m, i, o: cap.__setitem__("up1", o.detach())),
        dec.upsampler.register_forward_hook(lambda m, i, o: cap.__setitem__("upsampled", o.detach())),
        dec.decoder.head.out.register_forward_hook(lambda m, i, o: cap.__setitem__("head_linear", o.detach())),
    ]
    with torch.no_grad():
        wav_b = dec(ids_b)
    for h in hooks:
        h.remove()
    out["b2x16_ids"] = ids_b.numpy().astype(np.int64)
    out["b2x16_wav"] = wav_b.numpy().astype(np.float32)                      # (2, 1, 15360)
    out["b2x16_up0"] = cap["up0"].numpy().astype(np.float32)                 # (2, 512, 48)
    out["b2x16_res0"] = cap["res0"].numpy().astype(np.float32)               # (2, 512, 48)
    out["b2x16_up1"] = cap["up1"].numpy().astype(np.float32)                 # (2, 256, 96)
    out["b2x16_upsampled"] = cap["upsampled"].numpy().astype(np.float32)     # (2, 96, 1024)
    out["b2x16_head_linear"] = cap["head_linear"].numpy().astype(np.float32)  # (2, 96, 642)
    path = os.path.join(ROOT, "tests", "golden", "reference_decode_48k_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
