"""Generates tests/golden/reference_encoder_seed0.npz by running the UNMODIFIED reference encoder modules
(TEST INFRASTRUCTURE; build container only -- it needs /root/reference):

    python oracle/make_golden_encoder.py

`tts.core.codec.encoder.Encoder.__init__` downloads w2v-BERT from the HuggingFace hub, so the golden is built
from the reference's own sub-modules exactly as `Encoder.__init__` / `forward` wire them
(tts/core/codec/encoder.py:28-44, 58-78): `encoder_modules.AcousticEncoder`, `encoder_modules.SemanticEncoder`,
`torch.nn.Linear(2048, 2048)`, with the w2v-BERT hidden state replaced by a seeded random tensor. Weights are
the deterministic ones of oracle/encoder_oracle.py, loaded with strict `load_state_dict`.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import encoder_oracle as E  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    from tts.core.codec import encoder_modules as ref  # the reference, unmodified

    torch.set_num_threads(os.cpu_count() or 1)
    sd = E.make_state_dict(seed=0)
    ac = ref.AcousticEncoder(num_generator_features=48, initial_conv_kernel_size=7, final_conv_kernel_size=3,
                             up_ratios=[2, 2, 4, 4, 5], dilations=(1, 3, 9), output_dim=1024)
    se = ref.SemanticEncoder(input_channels=1024, output_channels=1024, encode_channels=1024, kernel_size=3)
    fusion = torch.nn.Linear(2048, 2048)
    ac_sd = {k[len("acoustic_encoder."):]: v for k, v in sd.items() if k.startswith("acoustic_encoder.")}
    assert list(ac.state_dict().keys()) == list(ac_sd.keys()), "acoustic_encoder key order differs from the reference"
    ac.load_state_dict(ac_sd, strict=True)
    se_sd = {k[len("semantic_encoder."):]: v for k, v in sd.items() if k.startswith("semantic_encoder.")}
    assert list(se.state_dict().keys()) == list(se_sd.keys())
    se.load_state_dict(se_sd, strict=True)
    fusion.load_state_dict({"weight": sd["fusion_layer.weight"], "bias": sd["fusion_layer.bias"]}, strict=True)
    for m in (ac, se, fusion):
        m.eval()
    # the reference's own filter buffers are what the oracle's restated filter must reproduce
    ref_filter = ac.state_dict()["conv_blocks.1.block.0.block.0.upsample.filter"]
    fresh = ref.AcousticEncoder(48, 7, 3, [2, 2, 4, 4, 5], (1, 3, 9), 1024).state_dict()["conv_final_block.0.downsample.lowpass.filter"]
    assert torch.equal(fresh, sd["acoustic_encoder.conv_final_block.0.downsample.lowpass.filter"]), "kaiser filter restatement"

    g = torch.Generator().manual_seed(4321)
    out = {"filter": ref_filter.reshape(-1).numpy()}
    for name, B, T in (("b2x12", 2, 12), ("b1x50", 1, 50)):
        wav = 0.3 * torch.randn(B, 1, 320 * T, generator=g)
        w2v = torch.randn(B, T, 1024, generator=g)
        cap = {}
        hooks = [ac.conv_blocks[0].register_forward_hook(lambda m, i, o: cap.__setitem__("conv0", o.detach()))]
        for i in range(1, 6):
            hooks.append(ac.conv_blocks[i].register_forward_hook(lambda m, inp, o, i=i: cap.__setitem__(f"block{i}", o.detach())))
        with torch.no_grad():
            acoustic = ac(wav).transpose(1, 2)                                  # encoder.py:60-61
            semantic = se(w2v.transpose(1, 2))                                  # :63-64
            hidden = torch.cat([semantic, acoustic], dim=1)                     # :66-68
            hidden = fusion(hidden.transpose(1, 2)).transpose(1, 2)             # :69
        for h in hooks:
            h.remove()
        out[f"{name}_wav"] = wav.numpy()
        out[f"{name}_w2v"] = w2v.numpy()
        out[f"{name}_acoustic"] = acoustic.numpy()      # (B, 1024, T)
        out[f"{name}_semantic"] = semantic.numpy()      # (B, 1024, T)
        out[f"{name}_hidden"] = hidden.numpy()          # (B, 2048, T)
        if name == "b2x12":
            out[f"{name}_conv0"] = cap["conv0"].numpy()                 # (B, 48, S)
            for i in (1, 3, 5):
                out[f"{name}_block{i}"] = cap[f"block{i}"].numpy()
        # the restatement, checked right here as well
        st = {}
        o_hidden = E.encoder_hidden(sd, wav, w2v, stages=st)
        err = (o_hidden - hidden).abs().max().item() / hidden.abs().max().item()
        print(f"{name}: hidden {tuple(hidden.shape)} |max| {hidden.abs().max():.3f}  oracle rel err {err:.2e}")
        assert err < 1e-5
    path = os.path.join(GOLDEN, "reference_encoder_seed0.npz")
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
